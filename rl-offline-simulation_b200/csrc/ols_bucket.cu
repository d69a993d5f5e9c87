// Bucketing of the logged transitions into per-state FIFO queues (CSR), on the GPU.
//
// Replaces offsim4rl/evaluators/psrs.py:19-30 `PSRS.reset_sampler`:
//   :22     init_queue = rows with info['t'] == 0, in log order        -> init_* kernels
//   :26     sorted(buffer, key=z) (stable) + groupby -> queues[z]      -> histogram / scan / scatter
//   :23,29-30  default_rng(seed).shuffle(...) of init_queue and of every queue
//                                                                      -> shuffle_perm / gather kernels
//
// The sort is a STABLE counting sort in one scatter pass.  The rows are cut into T contiguous
// tiles, one tile per CTA.  Pass 1 builds the tile's histogram in shared memory (all warps, atomics:
// order is irrelevant for counting); a column scan over the [tile][bucket] table gives every tile its
// first slot in every queue; pass 2 re-walks the tile in rounds of 8 batches (one per warp, consecutive rows):
// every warp loads its batch's columns and ranks equal keys with __match_any_sync (lower lane = earlier row), then
// the warps take their slots from the tile's running per-queue counters in shared memory one after the other, in
// log order, separated by __syncthreads (8 cheap barriers per 1024 rows - no spinning, no fences), and finally all
// warps scatter concurrently.  Each row's packed record (which carries the
// row index, i.e. the queue order) goes straight to its final slot as full 32-byte sectors.  Algorithmic
// traffic per row: z twice (8 B), the log columns once (~13 + 4*nA B), the record once (rec_bytes).
#include <limits.h>

#include "ols_common.cuh"

namespace {

constexpr unsigned kFull = 0xffffffffu;

struct BucketPlan {
  int ctas;
  int tiles;          // = ctas (one tile per CTA)
  int64_t tile_rows;  // multiple of the 64-row batch
  size_t smem_bytes;  // per CTA: one counter per queue
};

constexpr int kBucketThreads = 256;  // 8 warps per CTA
#ifndef OLS_BATCH_ROWS
#define OLS_BATCH_ROWS 128
#endif
constexpr int kBatchRows = OLS_BATCH_ROWS;  // rows a warp holds per round (U x 32)

// One counter array of n_buckets entries per CTA; as many CTAs per SM as shared memory and the thread
// limit allow, and a grid that is a multiple of the SM count.
inline bool make_plan(int64_t n_rows, int n_buckets, BucketPlan* p) {
  const size_t smem_cap = 200 * 1024;
  const size_t per_cta = (size_t)n_buckets * sizeof(uint32_t) + 64;
  if (per_cta > smem_cap) return false;
  int ctas_per_sm = (int)(smem_cap / per_cta);
  if (ctas_per_sm > 8) ctas_per_sm = 8;  // 2048 threads / 256
  const int64_t max_tiles = (int64_t)OLS_SM_COUNT * ctas_per_sm;
  int64_t want = (n_rows + 4095) / 4096;  // small logs: no tile shorter than ~4096 rows
  if (want < 1) want = 1;
  const int64_t tiles = want < max_tiles ? want : max_tiles;
  p->ctas = (int)tiles;
  p->tiles = (int)tiles;
  int64_t tr = (n_rows + tiles - 1) / tiles;
  tr = (tr + kBatchRows - 1) / kBatchRows * kBatchRows;
  if (tr < kBatchRows) tr = kBatchRows;
  p->tile_rows = tr;
  p->smem_bytes = per_cta;
  return true;
}

struct Scratch {
  uint32_t* tile_hist;  // [tiles][n_buckets]
  uint32_t* totals;     // [n_buckets]
  uint32_t* invalid;    // [1] rows whose bucket is out of range (skipped)
};

inline size_t scratch_bytes_for(const BucketPlan& p, int n_buckets) {
  size_t a = ((size_t)p.tiles * n_buckets + n_buckets + 64) * sizeof(uint32_t);
  return (a + 255) / 256 * 256;
}
inline Scratch carve(void* base, const BucketPlan& p, int n_buckets) {
  Scratch s;
  s.tile_hist = (uint32_t*)base;
  s.totals = s.tile_hist + (size_t)p.tiles * n_buckets;
  s.invalid = s.totals + n_buckets;
  return s;
}

// ---------------------------------------------------------------------------------------- pass 1
__global__ void __launch_bounds__(kBucketThreads) tile_histogram_kernel(const int32_t* __restrict__ z, int64_t n_rows,
                                                                         int z_offset, int n_buckets, int64_t tile_rows,
                                                                         uint32_t* __restrict__ tile_hist,
                                                                         uint32_t* __restrict__ invalid) {
  extern __shared__ uint32_t s_hist[];
  for (int b = threadIdx.x; b < n_buckets; b += blockDim.x) s_hist[b] = 0;
  __syncthreads();
  const int64_t lo = (int64_t)blockIdx.x * tile_rows;
  const int64_t hi = min(lo + tile_rows, n_rows);
  uint32_t bad = 0;
  for (int64_t base = lo + threadIdx.x; base < hi; base += 4 * kBucketThreads) {
    int key[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = base + (int64_t)j * kBucketThreads;
      key[j] = (i < hi) ? __ldg(z + i) + z_offset : INT_MIN;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (key[j] >= 0 && key[j] < n_buckets) atomicAdd(&s_hist[key[j]], 1u);
      else if (key[j] != INT_MIN) ++bad;  // id outside [-z_offset, n_buckets - z_offset): row is dropped
    }
  }
  __syncthreads();
  uint32_t* out = tile_hist + (size_t)blockIdx.x * n_buckets;
  for (int b = threadIdx.x; b < n_buckets; b += blockDim.x) out[b] = s_hist[b];
  if (bad) atomicAdd(invalid, bad);
}

// ---------------------------------------------------------------------------------------- scans
// Column scan of the [tiles][n_buckets] table: block = 32 buckets (x) by 32 tile stripes (y).
__global__ void __launch_bounds__(1024) tile_scan_kernel(uint32_t* __restrict__ tile_hist, int tiles, int n_buckets,
                                                         uint32_t* __restrict__ totals) {
  __shared__ uint32_t part[32][33];
  const int b = blockIdx.x * 32 + threadIdx.x;
  const int y = threadIdx.y;
  const int per = (tiles + 31) / 32;
  const int t0 = y * per, t1 = min(t0 + per, tiles);
  uint32_t sum = 0;
  if (b < n_buckets)
    for (int t = t0; t < t1; ++t) sum += tile_hist[(size_t)t * n_buckets + b];
  part[y][threadIdx.x] = sum;
  __syncthreads();
  if (y == 0) {
    uint32_t run = 0;
    for (int k = 0; k < 32; ++k) {
      uint32_t v = part[k][threadIdx.x];
      part[k][threadIdx.x] = run;
      run += v;
    }
    if (b < n_buckets) totals[b] = run;
  }
  __syncthreads();
  if (b < n_buckets) {
    uint32_t run = part[y][threadIdx.x];
    for (int t = t0; t < t1; ++t) {
      size_t idx = (size_t)t * n_buckets + b;
      uint32_t v = tile_hist[idx];
      tile_hist[idx] = run;
      run += v;
    }
  }
}

// Exclusive scan of the bucket totals -> CSR offsets (single CTA; n_buckets is at most ~51k).
__global__ void __launch_bounds__(1024) offsets_kernel(const uint32_t* __restrict__ totals, int n_buckets,
                                                        int32_t* __restrict__ offsets) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_buckets; base += 1024) {
    int i = base + threadIdx.x;
    uint32_t v = (i < n_buckets) ? totals[i] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(kFull, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_sums[lane];
      uint32_t winc = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(kFull, winc, d);
        if (lane >= d) winc += o;
      }
      warp_sums[lane] = winc - w;  // exclusive
    }
    __syncthreads();
    uint32_t excl = carry + warp_sums[warp] + inc - v;
    if (i < n_buckets) offsets[i] = (int32_t)excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n_buckets] = (int32_t)carry;
}

// ---------------------------------------------------------------------------------------- pass 2
template <int NV>
struct RowData {
  float4 p[NV];
  float r;
  int zn;
  uint32_t meta;
};

template <int NV>
__device__ __forceinline__ void load_row(const ols_log_t& log, int64_t i, RowData<NV>& d) {
  const int nA = log.n_actions;
  float pv[4 * NV];
#pragma unroll
  for (int j = 0; j < 4 * NV; ++j) pv[j] = 1.0f;  // padding entries hold 1 (see accept_filter)
  if (log.p_is_f64) {
    const double* p = (const double*)log.p_log + i * nA;
#pragma unroll
    for (int j = 0; j < 4 * NV; ++j)
      if (j < nA) pv[j] = (float)__ldg(p + j);
  } else if (NV == 1 && nA == 4) {
    float4 v = __ldg((const float4*)log.p_log + i);
    pv[0] = v.x; pv[1] = v.y; pv[2] = v.z; pv[3] = v.w;
  } else if (NV == 1 && nA == 2) {
    float2 v = __ldg((const float2*)log.p_log + i);
    pv[0] = v.x; pv[1] = v.y;
  } else {
    const float* p = (const float*)log.p_log + i * nA;
#pragma unroll
    for (int j = 0; j < 4 * NV; ++j)
      if (j < nA) pv[j] = __ldg(p + j);
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) d.p[v] = make_float4(pv[4 * v], pv[4 * v + 1], pv[4 * v + 2], pv[4 * v + 3]);
  d.r = __ldg(log.reward + i);
  d.zn = __ldg(log.z_next + i);
  float mn = pv[0], mx = pv[0];
#pragma unroll
  for (int j = 1; j < 4 * NV; ++j)
    if (j < nA) {
      mn = fminf(mn, pv[j]);
      mx = fmaxf(mx, pv[j]);
    }
  const int tame = (mn >= 9.094947e-13f) && (mx <= 1.0995116e12f);  // [2^-40, 2^40]; false for NaN
  d.meta = ols_pack_meta(__ldg(log.action + i), __ldg(log.terminal + i), tame);
}

#ifndef OLS_SCATTER_MINB
#define OLS_SCATTER_MINB 1
#endif
template <int NV>
__global__ void __launch_bounds__(kBucketThreads, OLS_SCATTER_MINB) scatter_pack_kernel(ols_log_t log, ols_queues_t q, int64_t tile_rows,
                                                                       const uint32_t* __restrict__ tile_hist) {
  extern __shared__ uint32_t s_next_slot[];   // [n_buckets] next free slot of every queue for this tile
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarps = kBucketThreads / 32;
  const int nB = q.n_buckets;
  const uint32_t* th = tile_hist + (size_t)blockIdx.x * nB;
  for (int b = threadIdx.x; b < nB; b += blockDim.x) s_next_slot[b] = (uint32_t)__ldg(q.offsets + b) + th[b];
  __syncthreads();
  const int64_t lo = (int64_t)blockIdx.x * tile_rows;
  const int64_t hi = min(lo + tile_rows, log.n_rows);
  const uint32_t lt_mask = (1u << lane) - 1u;
  constexpr int U = kBatchRows / 32;
  // The tile is consumed in rounds of kWarps batches; warp w holds the w-th batch of the round (kBatchRows consecutive
  // rows), so log order = (round, warp, j, lane).  The loop count is uniform over the CTA: barriers inside.
  const int n_rounds = (int)((hi - lo + (int64_t)kWarps * kBatchRows - 1) / ((int64_t)kWarps * kBatchRows));
  for (int rd = 0; rd < n_rounds; ++rd) {
    const int64_t base = lo + ((int64_t)rd * kWarps + warp) * kBatchRows;
    int key[U];
    RowData<NV> row[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int64_t i = base + j * 32 + lane;
      key[j] = -1;
      if (i < hi) {
        // the row's columns are fetched together with its key, not after it: one memory latency per round, not two
        const int b = __ldg(log.z + i) + q.z_offset;
        load_row<NV>(log, i, row[j]);
        key[j] = (b >= 0 && b < nB) ? b : -1;
      }
    }
    // lanes holding the same key, lowest lane = earliest row; invalid lanes form their own group
    unsigned peers[U];
#pragma unroll
    for (int j = 0; j < U; ++j) peers[j] = __match_any_sync(kFull, key[j]);
    // ---- the warps take their slots strictly in log order: one barrier-separated turn per warp
    uint32_t first[U];
#pragma unroll 1
    for (int w = 0; w < kWarps; ++w) {
      if (warp == w) {
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const int leader = __ffs(peers[j]) - 1;
          uint32_t f = 0;
          if (key[j] >= 0 && lane == leader) {
            f = s_next_slot[key[j]];
            s_next_slot[key[j]] = f + __popc(peers[j]);
          }
          __syncwarp();  // the next j may hit the same counter from another lane
          first[j] = __shfl_sync(kFull, f, leader);
        }
      }
      __syncthreads();
    }
    // ---- scattered stores
#pragma unroll
    for (int j = 0; j < U; ++j) {
      if (key[j] >= 0) {
        const uint32_t pos = first[j] + __popc(peers[j] & lt_mask);
        const int64_t i = base + j * 32 + lane;
        float4* dst = (float4*)((char*)q.records + (size_t)pos * q.rec_bytes);
        float4 tail;
        tail.x = row[j].r;
        tail.y = __int_as_float(row[j].zn);
        tail.z = __int_as_float((int32_t)i);  // the row index: reading it along a queue gives the queue order
        tail.w = __uint_as_float(row[j].meta);
        if (NV == 1) {
            // a 32-byte record is one sector: ONE 256-bit store, i.e. one full-sector write transaction instead of two
            // half-sector ones (the scatter is bound by the number of isolated write transactions, not by bytes)
            const float4 p0 = row[j].p[0];
            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(p0.x), "f"(p0.y), "f"(p0.z), "f"(p0.w),
                         "f"(tail.x), "f"(tail.y), "f"(tail.z), "f"(tail.w)
                         : "memory");
        } else {
#pragma unroll
          for (int v = 0; v < NV; ++v) dst[v] = row[j].p[v];
          dst[NV] = tail;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------- hot records
// One streaming pass over the packed queues: 32 bytes in, 8 bytes out per record (n_actions <= 4).
__global__ void __launch_bounds__(256) hot_pack_kernel(ols_queues_t q) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q.n_rows; i += stride) {
    const float4* rp = (const float4*)((const char*)q.records + (size_t)i * 32);
    const float4 p = __ldg(rp);
    const float4 tail = __ldg(rp + 1);
    const uint32_t meta = __float_as_uint(tail.w);
    if (q.hot_kind == OLS_HOT_TAIL)
      ((uint2*)q.hot)[i] = ols_hot_tail_pack(ols_meta_action(meta), tail.x, __float_as_int(tail.y), ols_meta_terminal(meta));
    else
      ((uint2*)q.hot)[i] = ols_hot_pack(p, ols_meta_action(meta), q.n_actions);
  }
}

// Per queue: does every record hold the distribution of the first one (bitwise)?  One CTA per queue (grid-strided).  The same
// pass writes the hot records (when asked to), so the packed queues are read once for both.
template <bool WRITE_HOT>
__global__ void __launch_bounds__(256) queue_dist_kernel(ols_queues_t q) {
  __shared__ int s_diff;
  for (int b = blockIdx.x; b < q.n_buckets; b += gridDim.x) {
    const int lo = __ldg(q.offsets + b), hi = __ldg(q.offsets + b + 1);
    if (threadIdx.x == 0) s_diff = 0;
    __syncthreads();
    float4 p0 = make_float4(1.f, 1.f, 1.f, 1.f);
    int diff = 0;
    if (hi > lo) {
      p0 = __ldg((const float4*)((const char*)q.records + (size_t)lo * 32));
      for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const float4* rp = (const float4*)((const char*)q.records + (size_t)i * 32);
        const float4 p = __ldg(rp);
        diff |= (__float_as_uint(p.x) != __float_as_uint(p0.x)) | (__float_as_uint(p.y) != __float_as_uint(p0.y)) |
                (__float_as_uint(p.z) != __float_as_uint(p0.z)) | (__float_as_uint(p.w) != __float_as_uint(p0.w));
        if (WRITE_HOT) {
          const float4 tail = __ldg(rp + 1);
          const uint32_t meta = __float_as_uint(tail.w);
          if (q.hot_kind == OLS_HOT_TAIL)
            ((uint2*)q.hot)[i] = ols_hot_tail_pack(ols_meta_action(meta), tail.x, __float_as_int(tail.y), ols_meta_terminal(meta));
          else
            ((uint2*)q.hot)[i] = ols_hot_pack(p, ols_meta_action(meta), q.n_actions);
        }
      }
    }
    if (diff) s_diff = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
      const float v[4] = {p0.x, p0.y, p0.z, p0.w};
      bool tame = true;
      for (int j = 0; j < q.n_actions && j < 4; ++j) tame = tame && (v[j] >= 9.094947e-13f && v[j] <= 1.0995116e12f);
      float* out = q.queue_dist + 8 * (size_t)b;
      out[0] = p0.x; out[1] = p0.y; out[2] = p0.z; out[3] = p0.w;
      ((int*)out)[4] = (s_diff ? 0 : 1) | (tame ? 2 : 0);
      ((int*)out)[5] = ((int*)out)[6] = ((int*)out)[7] = 0;
      if (s_diff && q.const_info) atomicAdd(q.const_info, 1);
    }
    __syncthreads();
  }
}

inline int pack_hot(const ols_queues_t* q, cudaStream_t stream) {
  if (q->n_rows == 0 || (q->hot == nullptr && q->queue_dist == nullptr)) {
    if (q->const_info) OLS_CUDA_CHECK(cudaMemsetAsync(q->const_info, 0, sizeof(int32_t), stream));
    return OLS_OK;
  }
  OLS_REQUIRE(q->n_actions <= 4 && q->rec_bytes == 32, OLS_ERR_UNSUPPORTED, "hot records need n_actions <= 4");
  OLS_REQUIRE(q->hot_kind == OLS_HOT_PROBS || q->hot_kind == OLS_HOT_TAIL, OLS_ERR_INVALID, "bad hot_kind");
  if (q->queue_dist) {
    OLS_REQUIRE((uintptr_t)q->queue_dist % 32 == 0, OLS_ERR_INVALID, "queue_dist must be 32-byte aligned");
    if (q->const_info) OLS_CUDA_CHECK(cudaMemsetAsync(q->const_info, 0, sizeof(int32_t), stream));
    int blocks = q->n_buckets < OLS_SM_COUNT * 8 ? q->n_buckets : OLS_SM_COUNT * 8;
    // queues of a few thousand records and more: one pass checks the distributions and writes the hot records
    const bool fused = q->hot != nullptr && q->n_rows / q->n_buckets >= 1024;
    if (fused) {
      OLS_REQUIRE((uintptr_t)q->hot % 8 == 0, OLS_ERR_INVALID, "hot records must be 8-byte aligned");
      queue_dist_kernel<true><<<blocks, 256, 0, stream>>>(*q);
    } else {
      queue_dist_kernel<false><<<blocks, 256, 0, stream>>>(*q);
    }
    OLS_LAUNCH_CHECK();
    if (fused) return OLS_OK;
  }
  if (q->hot) {
    OLS_REQUIRE(q->hot_kind != OLS_HOT_TAIL || q->queue_dist, OLS_ERR_INVALID, "OLS_HOT_TAIL needs queue_dist");
    OLS_REQUIRE((uintptr_t)q->hot % 8 == 0, OLS_ERR_INVALID, "hot records must be 8-byte aligned");
    int blocks = ols_div_up(q->n_rows, 256);
    if (blocks > OLS_SM_COUNT * 16) blocks = OLS_SM_COUNT * 16;
    hot_pack_kernel<<<blocks, 256, 0, stream>>>(*q);
    OLS_LAUNCH_CHECK();
  }
  return OLS_OK;
}

// ---------------------------------------------------------------------------------------- initial rows
// Stable compaction of the rows with step == 0: per-CTA counts -> scan -> ordered write.
constexpr int kInitThreads = 256;

__global__ void __launch_bounds__(kInitThreads) init_count_kernel(const int32_t* __restrict__ step, int64_t n_rows,
                                                                   int64_t chunk, uint32_t* __restrict__ counts) {
  __shared__ uint32_t wsum[kInitThreads / 32];
  const int64_t lo = (int64_t)blockIdx.x * chunk, hi = min(lo + chunk, n_rows);
  uint32_t c = 0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += kInitThreads) c += (__ldg(step + i) == 0);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(kFull, c, d);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < kInitThreads / 32; ++w) t += wsum[w];
    counts[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024) init_scan_kernel(uint32_t* __restrict__ counts, int n, int64_t* __restrict__ total) {
  // single CTA, n <= a few thousand
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    int i = base + threadIdx.x;
    uint32_t v = (i < n) ? counts[i] : 0u, inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(kFull, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_sums[lane], winc = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(kFull, winc, d);
        if (lane >= d) winc += o;
      }
      warp_sums[lane] = winc - w;
    }
    __syncthreads();
    uint32_t excl = carry + warp_sums[warp] + inc - v;
    if (i < n) counts[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = (int64_t)carry;
}

__global__ void all_rows_prefix_kernel(uint32_t* __restrict__ prefix, int n_chunks, uint32_t chunk, int64_t n_rows,
                                       int64_t* __restrict__ total) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_chunks) prefix[c] = (uint32_t)c * chunk;
  if (c == 0) *total = n_rows;
}

__global__ void __launch_bounds__(kInitThreads) init_write_kernel(ols_log_t log, ols_queues_t q, int64_t chunk,
                                                                   const uint32_t* __restrict__ prefix) {
  __shared__ uint32_t wsum[kInitThreads / 32];
  __shared__ uint32_t running;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t lo = (int64_t)blockIdx.x * chunk, hi = min(lo + chunk, log.n_rows);
  if (threadIdx.x == 0) running = prefix[blockIdx.x];
  __syncthreads();
  for (int64_t base = lo; base < hi; base += kInitThreads) {
    int64_t i = base + threadIdx.x;
    bool flag = (i < hi) && (log.step == nullptr || __ldg(log.step + i) == 0);
    unsigned ballot = __ballot_sync(kFull, flag);
    if (lane == 0) wsum[warp] = __popc(ballot);
    __syncthreads();
    uint32_t before = running;
    for (int w = 0; w < warp; ++w) before += wsum[w];
    if (flag) {
      uint32_t pos = before + __popc(ballot & ((1u << lane) - 1u));
      int zz = __ldg(log.z + i);
      q.init_row[pos] = (int32_t)i;
      q.init_z[pos] = zz;
      q.init_s[pos] = log.s ? __ldg(log.s + i) : zz;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t t = 0;
      for (int w = 0; w < kInitThreads / 32; ++w) t += wsum[w];
      running += t;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------- seeded shuffle
// np.random.default_rng(seed).shuffle(x) on a length-n list: for i = n-1 .. 1: j = interval(i);
// swap(x[i], x[j]); interval() is NumPy's masked rejection on buffered 32-bit halves of PCG64 output
// (64-bit draws when i > 2^32-1, which cannot occur with int32 row indices).  One warp replays one queue, 32 steps
// per round:
//  * draws: PCG64 is an LCG, so lane l jumps straight to output n0 + l (state = A_l * base + C_l, constants built once):
//    a round has the next 64 raw 32-bit draws in registers.  Which draw serves which step depends on the rejections
//    before it; that triangular system (slot t needs the accept count of slots < t) is solved by ballot/popc
//    fixed-point iteration — exact after at most as many passes as there are slots, two or three in practice;
//  * swaps: the round's swaps run as one parallel round of loads and stores when they touch pairwise distinct slots
//    (the rule for long queues), else lane 0 applies them in order.
// perm[] holds absolute slot indices, pre-filled with iota; segment `seg` covers [lo, lo+len).
__global__ void iota_kernel(int32_t* __restrict__ perm, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    perm[i] = (int32_t)i;
}

constexpr int kShuffleWarps = 4;

__global__ void __launch_bounds__(32 * kShuffleWarps) shuffle_perm_kernel(const int32_t* __restrict__ offsets, int n_buckets,
                                                                          int64_t n_rows, int64_t n_init, uint64_t seed,
                                                                          int32_t* __restrict__ perm) {
  __shared__ uint32_t s_j[kShuffleWarps][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seg = blockIdx.x * kShuffleWarps + warp;
  if (seg > n_buckets) return;  // warp-uniform
  int64_t lo, len;
  if (seg < n_buckets) {
    lo = offsets[seg];
    len = offsets[seg + 1] - lo;
  } else {  // the initial-state list lives after the n_rows queue slots
    lo = n_rows;
    len = n_init;
  }
  if (len < 2) return;
  ols_u128 base, inc;  // base = generator state before output n0 (identical in every lane)
  ols_pcg64_seed(seed, &base, &inc);
  // lane constants: state before output n0 + lane + 1, i.e. after lane + 1 steps = A * base + C
  const ols_u128 mult = {OLS_PCG_MULT_HI, OLS_PCG_MULT_LO};
  ols_u128 A = mult, Cc = inc;
  for (int k = 0; k < lane; ++k) {
    Cc = ols_add128(ols_mul128(Cc, mult), inc);
    A = ols_mul128(A, mult);
  }
  const unsigned lt = (1u << lane) - 1u;
  uint64_t n0 = 0;    // outputs fully consumed so far
  bool half = false;  // the low half of output n0 is consumed as well
  int32_t* x = perm + lo;
  uint32_t* jbuf = s_j[warp];
  uint32_t i = (uint32_t)len - 1;  // steps i, i-1, ..., 1
  while (i >= 1) {
    const uint32_t W = i < 32u ? i : 32u;  // at most this many steps in the round: i - k, k = 0 .. W-1
    // ---- the next 64 raw draws: slot 2l = low half, slot 2l+1 = high half of output n0 + l
    const ols_u128 st = ols_add128(ols_mul128(base, A), Cc);
    const uint64_t out = ols_pcg_output(st);
    const uint32_t v_lo = (uint32_t)out, v_hi = (uint32_t)(out >> 32);
    const bool ok_lo = !(lane == 0 && half);
    // ---- which draw serves which step: slot t is accepted iff (v & mask(i_t)) <= i_t with i_t = i - #accepted before t
    unsigned b_lo = 0, b_hi = 0;
    uint32_t c_lo = 0, c_hi = 0, r_lo = 0, r_hi = 0;
    for (;;) {
      c_lo = (uint32_t)(__popc(b_lo & lt) + __popc(b_hi & lt));
      bool a_lo = false;
      if (ok_lo && c_lo < W) {
        const uint32_t it = i - c_lo;
        r_lo = v_lo & (0xffffffffu >> __clz(it));
        a_lo = r_lo <= it;
      }
      c_hi = c_lo + (a_lo ? 1u : 0u);
      bool a_hi = false;
      if (c_hi < W) {
        const uint32_t it = i - c_hi;
        r_hi = v_hi & (0xffffffffu >> __clz(it));
        a_hi = r_hi <= it;
      }
      const unsigned n_lo = __ballot_sync(kFull, a_lo), n_hi = __ballot_sync(kFull, a_hi);
      const bool same = (n_lo == b_lo) && (n_hi == b_hi);
      b_lo = n_lo;
      b_hi = n_hi;
      if (same) break;
    }
    const uint32_t T = (uint32_t)(__popc(b_lo) + __popc(b_hi));  // steps served by this batch of draws (<= W)
    if ((b_lo >> lane) & 1u) jbuf[c_lo] = r_lo;
    if ((b_hi >> lane) & 1u) jbuf[c_hi] = r_hi;
    // raw draws consumed: through the last accepted one when the round is complete, all 64 otherwise
    uint32_t used = 64;
    if (T == W) {
      const unsigned any = b_lo | b_hi;
      const int last = 31 - __clz(any);
      used = 2u * (uint32_t)last + (((b_hi >> last) & 1u) ? 2u : 1u);
    }
    const uint32_t adv = used >> 1;  // whole outputs consumed; an odd count leaves the low half of the next one used
    const uint64_t st_hi = __shfl_sync(kFull, st.hi, adv > 0 ? (int)adv - 1 : 0);
    const uint64_t st_lo = __shfl_sync(kFull, st.lo, adv > 0 ? (int)adv - 1 : 0);
    if (adv > 0) {
      base.hi = st_hi;
      base.lo = st_lo;
    }
    n0 += adv;
    half = (used & 1u) != 0u;
    __syncwarp();
    if (T > 0) {
      const bool mine = (uint32_t)lane < T;
      const uint32_t my_i = i - (uint32_t)lane;
      const uint32_t my_j = mine ? jbuf[lane] : 0xffffffffu - (uint32_t)lane;  // idle lanes: distinct dummies
      const bool self = mine && my_j == my_i;  // swap with itself
      // independent iff every partner lies below the round's own slots (i-T, i] and no two partners coincide
      const bool below = !mine || self || my_j + T <= i;
      const unsigned peers = __match_any_sync(kFull, self ? 0xffffff00u - (uint32_t)lane : my_j);
      const bool alone = (peers == (1u << lane));
      if (__all_sync(kFull, below && alone)) {
        if (mine && !self) {
          const int32_t a = x[my_i], b = x[my_j];
          x[my_i] = b;
          x[my_j] = a;
        }
      } else if (lane == 0) {
        for (uint32_t k = 0; k < T; ++k) {
          const uint32_t ii = i - k, j = jbuf[k];
          const int32_t a = x[ii], b = x[j];
          x[ii] = b;
          x[j] = a;
        }
      }
      __syncwarp();  // the round's stores precede the next round's loads; jbuf is free again
      i -= T;
    }
  }
}

__global__ void __launch_bounds__(256) gather_queues_kernel(ols_queues_t src, ols_queues_t dst,
                                                             const int32_t* __restrict__ perm) {
  const int chunks = src.rec_bytes / 16;
  const int64_t total = src.n_rows * chunks;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t slot = e / chunks;
    int c = (int)(e % chunks);
    int32_t from = __ldg(perm + slot);
    const float4* s = (const float4*)((const char*)src.records + (size_t)from * src.rec_bytes);
    float4* d = (float4*)((char*)dst.records + (size_t)slot * dst.rec_bytes);
    d[c] = __ldg(s + c);
  }
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < src.n_init; k += (int64_t)gridDim.x * blockDim.x) {
    int32_t from = __ldg(perm + src.n_rows + k) - (int32_t)src.n_rows;
    dst.init_row[k] = __ldg(src.init_row + from);
    dst.init_z[k] = __ldg(src.init_z + from);
    dst.init_s[k] = __ldg(src.init_s + from);
  }
}

int check_queues(const ols_queues_t* q, const char* who) {
  OLS_REQUIRE(q != nullptr, OLS_ERR_INVALID, "%s: null queues", who);
  OLS_REQUIRE(q->n_rows >= 0 && q->n_rows < (1ll << 31), OLS_ERR_INVALID, "%s: n_rows must be in [0, 2^31)", who);
  OLS_REQUIRE(q->n_actions >= 1 && q->n_actions <= OLS_MAX_ACTIONS, OLS_ERR_UNSUPPORTED,
              "%s: n_actions must be in [1, %d]", who, OLS_MAX_ACTIONS);
  OLS_REQUIRE(q->n_buckets >= 1, OLS_ERR_INVALID, "%s: n_buckets must be >= 1", who);
  OLS_REQUIRE(q->rec_bytes == ols_record_bytes(q->n_actions), OLS_ERR_INVALID, "%s: rec_bytes must be %d", who,
              ols_record_bytes(q->n_actions));
  OLS_REQUIRE(q->offsets != nullptr, OLS_ERR_INVALID, "%s: null offsets", who);
  OLS_REQUIRE(q->n_rows == 0 || q->records, OLS_ERR_INVALID, "%s: null records", who);
  OLS_REQUIRE((uintptr_t)q->records % 32 == 0, OLS_ERR_INVALID, "%s: records must be 32-byte aligned", who);
  return OLS_OK;
}

}  // namespace

extern "C" int32_t ols_record_bytes(int32_t n_actions) {
  if (n_actions < 1 || n_actions > OLS_MAX_ACTIONS) return 0;
  return 16 * (1 + (n_actions + 3) / 4);
}

extern "C" size_t ols_bucket_scratch_bytes(int64_t n_rows, int32_t n_buckets) {
  BucketPlan p;
  if (n_rows < 0 || n_buckets < 1 || !make_plan(n_rows, n_buckets, &p)) return 0;
  size_t a = scratch_bytes_for(p, n_buckets);
  // the initial-row compaction reuses the same scratch: one counter per 4096-row chunk
  size_t b = (size_t)((n_rows + 4095) / 4096 + 64) * sizeof(uint32_t);
  return a > b ? a : b;
}

extern "C" int ols_bucket(const ols_log_t* log, const ols_queues_t* q, void* scratch, size_t scratch_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OLS_REQUIRE(log != nullptr, OLS_ERR_INVALID, "ols_bucket: null log");
  int rc = check_queues(q, "ols_bucket");
  if (rc) return rc;
  OLS_REQUIRE(log->n_rows == q->n_rows && log->n_actions == q->n_actions, OLS_ERR_INVALID,
              "ols_bucket: log and queues disagree on n_rows / n_actions");
  BucketPlan plan;
  OLS_REQUIRE(make_plan(q->n_rows, q->n_buckets, &plan), OLS_ERR_UNSUPPORTED,
              "ols_bucket: n_buckets = %d needs more than 200 KB of shared memory for one tile's counters", q->n_buckets);
  OLS_REQUIRE(scratch != nullptr && scratch_bytes >= scratch_bytes_for(plan, q->n_buckets), OLS_ERR_INVALID,
              "ols_bucket: scratch too small (%zu < %zu)", scratch_bytes, scratch_bytes_for(plan, q->n_buckets));
  if (q->n_rows > 0) {
    OLS_REQUIRE(log->z && log->z_next && log->action && log->reward && log->terminal && log->p_log, OLS_ERR_INVALID,
                "ols_bucket: null log column");
    OLS_REQUIRE((uintptr_t)log->p_log % 16 == 0, OLS_ERR_INVALID, "ols_bucket: p_log must be 16-byte aligned");
  }
  Scratch s = carve(scratch, plan, q->n_buckets);
  OLS_CUDA_CHECK(cudaMemsetAsync(s.invalid, 0, sizeof(uint32_t), stream));
  const int threads = kBucketThreads;
  if (plan.smem_bytes > 48 * 1024) {
    OLS_CUDA_CHECK(cudaFuncSetAttribute(tile_histogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes));
  }
  tile_histogram_kernel<<<plan.ctas, threads, plan.smem_bytes, stream>>>(log->z, q->n_rows, q->z_offset, q->n_buckets,
                                                                        plan.tile_rows, s.tile_hist, s.invalid);
  OLS_LAUNCH_CHECK();
  tile_scan_kernel<<<ols_div_up(q->n_buckets, 32), dim3(32, 32), 0, stream>>>(s.tile_hist, plan.tiles, q->n_buckets, s.totals);
  OLS_LAUNCH_CHECK();
  offsets_kernel<<<1, 1024, 0, stream>>>(s.totals, q->n_buckets, q->offsets);
  OLS_LAUNCH_CHECK();
  if (q->n_rows == 0) return OLS_OK;
  const int nv = (q->n_actions + 3) / 4;
#define OLS_SCATTER(NV)                                                                                              \
  do {                                                                                                               \
    if (plan.smem_bytes > 48 * 1024)                                                                                 \
      OLS_CUDA_CHECK(cudaFuncSetAttribute(scatter_pack_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                          (int)plan.smem_bytes));                                                    \
    scatter_pack_kernel<NV><<<plan.ctas, threads, plan.smem_bytes, stream>>>(*log, *q, plan.tile_rows, s.tile_hist); \
  } while (0)
  switch (nv) {
    case 1: OLS_SCATTER(1); break;
    case 2: OLS_SCATTER(2); break;
    case 3: OLS_SCATTER(3); break;
    default: OLS_SCATTER(4); break;
  }
#undef OLS_SCATTER
  OLS_LAUNCH_CHECK();
  return pack_hot(q, stream);
}

extern "C" int ols_pack_hot(const ols_queues_t* q, void* stream_) {
  int rc = check_queues(q, "ols_pack_hot");
  if (rc) return rc;
  return pack_hot(q, (cudaStream_t)stream_);
}

extern "C" int ols_init_queue(const ols_log_t* log, const ols_queues_t* q, int64_t* n_init_dev, void* scratch,
                              size_t scratch_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OLS_REQUIRE(log && q && n_init_dev, OLS_ERR_INVALID, "ols_init_queue: null argument");
  OLS_REQUIRE(log->n_rows >= 0 && log->n_rows < (1ll << 31), OLS_ERR_INVALID, "ols_init_queue: bad n_rows");
  const int64_t chunk = 4096;
  const int n_chunks = (int)((log->n_rows + chunk - 1) / chunk);
  OLS_REQUIRE(scratch && scratch_bytes >= (size_t)(n_chunks + 1) * sizeof(uint32_t), OLS_ERR_INVALID,
              "ols_init_queue: scratch too small");
  if (log->n_rows == 0) {
    OLS_CUDA_CHECK(cudaMemsetAsync(n_init_dev, 0, sizeof(int64_t), stream));
    return OLS_OK;
  }
  OLS_REQUIRE(log->z && q->init_row && q->init_z && q->init_s, OLS_ERR_INVALID, "ols_init_queue: null buffer");
  uint32_t* counts = (uint32_t*)scratch;
  if (log->step) {
    init_count_kernel<<<n_chunks, kInitThreads, 0, stream>>>(log->step, log->n_rows, chunk, counts);
    OLS_LAUNCH_CHECK();
    init_scan_kernel<<<1, 1024, 0, stream>>>(counts, n_chunks, n_init_dev);
    OLS_LAUNCH_CHECK();
  } else {
    // data.py:72 — no steps column: every row is an initial state, chunk c starts at slot c * chunk
    all_rows_prefix_kernel<<<ols_div_up(n_chunks, 256), 256, 0, stream>>>(counts, n_chunks, (uint32_t)chunk, log->n_rows,
                                                                         n_init_dev);
    OLS_LAUNCH_CHECK();
  }
  init_write_kernel<<<n_chunks, kInitThreads, 0, stream>>>(*log, *q, chunk, counts);
  OLS_LAUNCH_CHECK();
  return OLS_OK;
}

extern "C" int ols_shuffle_queues(const ols_queues_t* src, const ols_queues_t* dst, uint64_t seed, int32_t* perm_scratch,
                                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_queues(src, "ols_shuffle_queues(src)");
  if (rc) return rc;
  rc = check_queues(dst, "ols_shuffle_queues(dst)");
  if (rc) return rc;
  OLS_REQUIRE(src->n_rows == dst->n_rows && src->n_buckets == dst->n_buckets && src->n_actions == dst->n_actions &&
                  src->n_init == dst->n_init,
              OLS_ERR_INVALID, "ols_shuffle_queues: src and dst shapes differ");
  OLS_REQUIRE(src->records != dst->records, OLS_ERR_INVALID, "ols_shuffle_queues: dst must not alias src (except offsets)");
  const int64_t total = src->n_rows + src->n_init;
  OLS_REQUIRE(total < (1ll << 31), OLS_ERR_UNSUPPORTED, "ols_shuffle_queues: n_rows + n_init must be < 2^31");
  if (total == 0) return OLS_OK;
  OLS_REQUIRE(perm_scratch != nullptr, OLS_ERR_INVALID, "ols_shuffle_queues: null perm_scratch");
  OLS_REQUIRE(src->n_init == 0 || (src->init_row && dst->init_row && src->init_z && dst->init_z && src->init_s && dst->init_s),
              OLS_ERR_INVALID, "ols_shuffle_queues: null init buffers");
  if (dst->offsets != src->offsets)
    OLS_CUDA_CHECK(cudaMemcpyAsync(dst->offsets, src->offsets, sizeof(int32_t) * (src->n_buckets + 1),
                                   cudaMemcpyDeviceToDevice, stream));
  int blocks = ols_div_up(total, 256);
  if (blocks > OLS_SM_COUNT * 8) blocks = OLS_SM_COUNT * 8;
  iota_kernel<<<blocks, 256, 0, stream>>>(perm_scratch, total);
  OLS_LAUNCH_CHECK();
  shuffle_perm_kernel<<<ols_div_up(src->n_buckets + 1, kShuffleWarps), 32 * kShuffleWarps, 0, stream>>>(src->offsets, src->n_buckets, src->n_rows,
                                                                              src->n_init, seed, perm_scratch);
  OLS_LAUNCH_CHECK();
  gather_queues_kernel<<<blocks, 256, 0, stream>>>(*src, *dst, perm_scratch);
  OLS_LAUNCH_CHECK();
  return pack_hot(dst, stream);
}
