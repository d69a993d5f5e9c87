// Micro-benchmark: achievable HBM throughput for RANDOM 32-byte / 128-byte reads over a 24 GB buffer (the C4 regime).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
struct __align__(32) R8 { float4 a, b; };
__device__ __forceinline__ void ld256(const void* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
// each thread: `iters` dependent-free random sector reads (address from a hash), `chunk` consecutive sectors each
template <int CHUNK, int DEP>
__global__ void rnd(const R8* p, size_t n, int iters, float* out) {
  uint64_t h = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 1;
  float s = 0;
  for (int it = 0; it < iters; ++it) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 29;
    size_t i = (size_t)__umul64hi(h, (uint64_t)(n - CHUNK)) & ~(size_t)(CHUNK - 1);  // multiply-shift: no 64-bit division
    float4 a[CHUNK], b[CHUNK];
#pragma unroll
    for (int c = 0; c < CHUNK; ++c) ld256(p + i + c, a[c], b[c]);
#pragma unroll
    for (int c = 0; c < CHUNK; ++c) s += a[c].x + b[c].w;
    if (DEP) h += (uint64_t)(s > 1e30f);  // make the next address depend on the loaded data
  }
  if (s == 123.456f) out[0] = s;
}
int main() {
  size_t bytes = (size_t)24 << 30;
  void* d; if (cudaMalloc(&d, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(d, 1, bytes);
  float* out; cudaMalloc(&out, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  size_t n = bytes / 32;
  auto time = [&](const char* name, auto launch, double moved) {
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-60s %8.3f ms  %8.1f GB/s  %.3e accesses/s (%s)\n", name, ms, moved / ms / 1e6, moved / 32 / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
  };
  int iters = 200;
  for (int occ : {2048, 1024, 512}) {
    int blocks = 148 * occ / 256;
    char nm[128];
    snprintf(nm, 128, "random 32B, independent, %d thr/SM", occ);
    time(nm, [&] { rnd<1, 0><<<blocks, 256>>>((const R8*)d, n, iters, out); }, (double)blocks * 256 * iters * 32);
    snprintf(nm, 128, "random 32B, dependent chain, %d thr/SM", occ);
    time(nm, [&] { rnd<1, 1><<<blocks, 256>>>((const R8*)d, n, iters, out); }, (double)blocks * 256 * iters * 32);
    snprintf(nm, 128, "random 128B (4 sectors), dependent chain, %d thr/SM", occ);
    time(nm, [&] { rnd<4, 1><<<blocks, 256>>>((const R8*)d, n, iters, out); }, (double)blocks * 256 * iters * 128);
  }
  return 0;
}
