// Micro-benchmark: what bounds RANDOM 32-byte reads on this box?  Sweeps the footprint (inside / outside the TLB reach, inside /
// outside L2) and the L2 fetch granularity (cudaLimitMaxL2FetchGranularity: 32 / 64 / 128 bytes).  If the access rate rises
// sharply for footprints of a few hundred MB the ceiling is address translation, not DRAM; if it rises with a 32-byte fetch
// granularity it is DRAM over-fetch.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scratch/randsweep scratch/randsweep.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
struct __align__(32) R8 { float4 a, b; };
__device__ __forceinline__ void ld256(const void* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
// WINDOW = 0: every access anywhere in the buffer.  WINDOW = 1: a CTA's accesses stay inside its own `win` sectors (page locality
// per CTA: the C4 alternative where a CTA's replicas own a compact slab of state).
__global__ void rnd(const R8* p, size_t n, int iters, size_t win, float* out) {
  uint64_t h = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 1;
  float s = 0;
  size_t base = 0, span = n;
  if (win) {
    span = win;
    uint64_t hb = (blockIdx.x + 1) * 0xD6E8FEB86659FD93ull;
    hb ^= hb >> 32;
    base = (size_t)__umul64hi(hb * 0x9E3779B97F4A7C15ull, (uint64_t)(n - win));
  }
  for (int it = 0; it < iters; ++it) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 29;
    size_t i = base + (size_t)__umul64hi(h, (uint64_t)span);
    float4 a, b;
    ld256(p + i, a, b);
    s += a.x + b.w;
  }
  if (s == 123.456f) out[0] = s;
}
int main() {
  size_t cap = (size_t)48 << 30;
  void* d; if (cudaMalloc(&d, cap) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(d, 1, cap);
  float* out; cudaMalloc(&out, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 200, blocks = 148 * 8;
  for (size_t gran : {(size_t)0, (size_t)32, (size_t)128}) {
    if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    printf("L2 fetch granularity: asked %zu, device reports %zu\n", gran, got);
    for (double gb : {0.25, 1.0, 4.0, 16.0, 48.0}) {
      size_t n = (size_t)(gb * (1ull << 30)) / 32;
      for (size_t win_mb : {(size_t)0, (size_t)2, (size_t)32}) {
        size_t win = win_mb * (1ull << 20) / 32;
        if (win >= n) continue;
        rnd<<<blocks, 256>>>((const R8*)d, n, iters, win, out); cudaDeviceSynchronize();
        cudaEventRecord(e0); rnd<<<blocks, 256>>>((const R8*)d, n, iters, win, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double acc = (double)blocks * 256 * iters;
        printf("  footprint %6.2f GB  per-CTA window %3zu MB : %8.3f ms  %.3e accesses/s  %7.1f GB/s (%s)\n", gb, win_mb, ms,
               acc / (ms * 1e-3), acc * 32 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  return 0;
}
