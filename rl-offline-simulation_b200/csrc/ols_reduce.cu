// Return / learning-curve evaluators as warp-shuffle segmented reductions.
//
// The reference returns per-run arrays (`Gs`, `lengths`: offsim4rl/evaluators/psrs.py:180-185,271)
// and averages them over runs when plotting (offsim4rl/utils/vis_utils.py:15-21).  Here the runs are
// replicas: Gs is [n_replicas][ep_cap] with a ragged valid length ep_index[r] per row.
//   * ols_reduce_curves: per episode index, sum / sum of squares / count over replicas.  Two
//     deterministic stages: partial[slab][e] over slabs of replicas, then a fixed-order fold.
//   * ols_replica_totals: one warp per replica reduces that replica's own segment with shuffles.
#include "ols_common.cuh"

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kSlabRows = 256;  // replicas per partial slab

// block = 32 episodes (x) by 8 replica stripes (y); consecutive x -> consecutive addresses of one row
__global__ void __launch_bounds__(256) curve_partial_kernel(const double* __restrict__ Gs, const int32_t* __restrict__ ep_index,
                                                            int64_t n_replicas, int ep_cap, int n_episodes,
                                                            double* __restrict__ partial) {
  __shared__ double s_sum[8][33], s_sq[8][33], s_cnt[8][33];
  const int e = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * kSlabRows;
  const int64_t r1 = min(r0 + kSlabRows, n_replicas);
  double sum = 0.0, sq = 0.0, cnt = 0.0;
  if (e < n_episodes) {
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      if (__ldg(ep_index + r) > e) {
        double g = __ldg(Gs + r * ep_cap + e);
        sum += g;
        sq = fma(g, g, sq);
        cnt += 1.0;
      }
    }
  }
  s_sum[threadIdx.y][threadIdx.x] = sum;
  s_sq[threadIdx.y][threadIdx.x] = sq;
  s_cnt[threadIdx.y][threadIdx.x] = cnt;
  __syncthreads();
  if (threadIdx.y == 0 && e < n_episodes) {
    for (int y = 1; y < 8; ++y) {
      sum += s_sum[y][threadIdx.x];
      sq += s_sq[y][threadIdx.x];
      cnt += s_cnt[y][threadIdx.x];
    }
    double* out = partial + ((size_t)blockIdx.y * 3) * n_episodes;
    out[e] = sum;
    out[n_episodes + e] = sq;
    out[2 * n_episodes + e] = cnt;
  }
}

// block = 32 episodes (x) by 32 slab stripes (y): stripe y folds slabs y, y + 32, ... in order, then the stripes are folded in
// order — a fixed summation tree, so the curve is deterministic (and independent of the launch geometry)
__global__ void __launch_bounds__(1024, 1) curve_fold_kernel(const double* __restrict__ partial, int n_slabs, int n_episodes,
                                                          double* __restrict__ sum, double* __restrict__ sumsq,
                                                          double* __restrict__ count) {
  __shared__ double s_a[32][33], s_b[32][33], s_c[32][33];
  const int e = blockIdx.x * 32 + threadIdx.x;
  double a = 0.0, b = 0.0, c = 0.0;
  if (e < n_episodes) {
    for (int s = threadIdx.y; s < n_slabs; s += 32) {
      const double* p = partial + ((size_t)s * 3) * n_episodes;
      a += p[e];
      b += p[n_episodes + e];
      c += p[2 * n_episodes + e];
    }
  }
  s_a[threadIdx.y][threadIdx.x] = a;
  s_b[threadIdx.y][threadIdx.x] = b;
  s_c[threadIdx.y][threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.y == 0 && e < n_episodes) {
#pragma unroll 4
    for (int y = 1; y < 32; ++y) {
      a += s_a[y][threadIdx.x];
      b += s_b[y][threadIdx.x];
      c += s_c[y][threadIdx.x];
    }
    sum[e] = a;
    sumsq[e] = b;
    count[e] = c;
  }
}

// one warp per replica: lanes stride over the replica's episodes, then a shuffle tree folds the warp
__global__ void __launch_bounds__(256) replica_totals_kernel(const double* __restrict__ Gs, const int32_t* __restrict__ lengths,
                                                             const int32_t* __restrict__ ep_index,
                                                             const int32_t* __restrict__ ep_len_count, int64_t n_replicas,
                                                             int ep_cap, double* __restrict__ ret_sum,
                                                             int64_t* __restrict__ len_sum) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n_replicas; r += n_warps) {
    const int ng = min(__ldg(ep_index + r), ep_cap), nl = min(__ldg(ep_len_count + r), ep_cap);
    double g = 0.0;
    long long l = 0;
    for (int e = lane; e < ng; e += 32) g += __ldg(Gs + r * ep_cap + e);
    for (int e = lane; e < nl; e += 32) l += __ldg(lengths + r * ep_cap + e);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      g += __shfl_xor_sync(kFull, g, d);
      l += __shfl_xor_sync(kFull, l, d);
    }
    if (lane == 0) {
      ret_sum[r] = g;
      len_sum[r] = l;
    }
  }
}

}  // namespace

extern "C" size_t ols_curve_scratch_bytes(int64_t n_replicas, int32_t n_episodes) {
  if (n_replicas < 0 || n_episodes < 0) return 0;
  size_t slabs = (size_t)((n_replicas + kSlabRows - 1) / kSlabRows);
  return (slabs * 3 * (size_t)n_episodes + 8) * sizeof(double);
}

extern "C" int ols_reduce_curves(const double* Gs, const int32_t* ep_index, int64_t n_replicas, int32_t ep_cap,
                                 int32_t n_episodes, double* sum, double* sumsq, double* count, void* scratch,
                                 size_t scratch_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OLS_REQUIRE(n_replicas >= 0 && n_episodes >= 0 && ep_cap >= n_episodes, OLS_ERR_INVALID, "ols_reduce_curves: bad sizes");
  if (n_episodes == 0) return OLS_OK;
  OLS_REQUIRE(sum && sumsq && count, OLS_ERR_INVALID, "ols_reduce_curves: null output");
  const int n_slabs = (int)((n_replicas + kSlabRows - 1) / kSlabRows);
  if (n_slabs > 0) {
    OLS_REQUIRE(Gs && ep_index, OLS_ERR_INVALID, "ols_reduce_curves: null input");
    OLS_REQUIRE(scratch && scratch_bytes >= ols_curve_scratch_bytes(n_replicas, n_episodes), OLS_ERR_INVALID,
                "ols_reduce_curves: scratch too small");
    OLS_REQUIRE(n_slabs <= 65535, OLS_ERR_UNSUPPORTED, "ols_reduce_curves: more than 16.7M replicas in one call");
    dim3 grid(ols_div_up(n_episodes, 32), n_slabs);
    curve_partial_kernel<<<grid, dim3(32, 8), 0, stream>>>(Gs, ep_index, n_replicas, ep_cap, n_episodes, (double*)scratch);
    OLS_LAUNCH_CHECK();
  }
  if (n_episodes > 0)
    curve_fold_kernel<<<ols_div_up(n_episodes, 32), dim3(32, 32), 0, stream>>>((const double*)scratch, n_slabs, n_episodes, sum, sumsq, count);
  OLS_LAUNCH_CHECK();
  return OLS_OK;
}

extern "C" int ols_replica_totals(const double* Gs, const int32_t* lengths, const int32_t* ep_index,
                                  const int32_t* ep_len_count, int64_t n_replicas, int32_t ep_cap, double* ret_sum,
                                  int64_t* len_sum, void* stream) {
  OLS_REQUIRE(n_replicas >= 0 && ep_cap >= 0, OLS_ERR_INVALID, "ols_replica_totals: bad sizes");
  if (n_replicas == 0) return OLS_OK;
  OLS_REQUIRE(Gs && lengths && ep_index && ep_len_count && ret_sum && len_sum, OLS_ERR_INVALID, "ols_replica_totals: null pointer");
  int64_t blocks = (n_replicas * 32 + 255) / 256;
  if (blocks > OLS_SM_COUNT * 8) blocks = OLS_SM_COUNT * 8;
  replica_totals_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(Gs, lengths, ep_index, ep_len_count, n_replicas, ep_cap,
                                                                       ret_sum, len_sum);
  OLS_LAUNCH_CHECK();
  return OLS_OK;
}
