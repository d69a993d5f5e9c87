// State encoders: observation -> integer state id (the PSRS queue key).
//
// Replaces, for the PSRS path, `encoder.encode(observations)` as called at
// offsim4rl/evaluators/per_state_rejection.py:30-31:
//   * CartpoleBoxEncoder.get_box / encode      offsim4rl/encoders/heuristic.py:20-60,65-71
//   * HOMEREncoder.encode over obs_encoder      offsim4rl/encoders/homer.py:159-168, models.py:15-19
//   * the MyGridNaviCoords grid discretiser     offsim4rl/envs/gridworld.py:205-211,123-124
// All three are streaming, HBM-bound kernels: one row per thread, 16-byte loads where the row
// shape allows, ids written coalesced.
#include "ols_common.cuh"

namespace {

constexpr int kThreads = 256;

inline int grid_for(int64_t n, int per_block) {
  int64_t blocks = (n + per_block - 1) / per_block;
  int64_t cap = (int64_t)OLS_SM_COUNT * 16;  // grid-stride beyond 16 resident CTAs per SM
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

// ---------------------------------------------------------------------------------------- grid
__global__ void __launch_bounds__(kThreads) encode_grid_kernel(const float* __restrict__ obs, int64_t n, int d_obs,
                                                                int cells, int32_t* __restrict__ ids) {
  const float fc = (float)cells;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int id = 0, mul = 1;
    if (d_obs == 2) {
      float2 o = __ldg(reinterpret_cast<const float2*>(obs) + i);
      int kx = (int)floorf(o.x * fc), ky = (int)floorf(o.y * fc);
      kx = min(max(kx, 0), cells - 1);
      ky = min(max(ky, 0), cells - 1);
      id = kx + cells * ky;
    } else {
      for (int d = 0; d < d_obs; ++d) {
        int k = (int)floorf(__ldg(obs + i * d_obs + d) * fc);
        k = min(max(k, 0), cells - 1);
        id += k * mul;
        mul *= cells;
      }
    }
    ids[i] = id;
  }
}

// ---------------------------------------------------------------------------------------- cartpole boxes
// The thresholds are the reference's python-float constants; comparing in the observation dtype means
// rounding each constant to that dtype first (NumPy >= 2 weak-scalar rule).
template <typename T>
__device__ __forceinline__ int cartpole_box(T x, T x_dot, T theta, T theta_dot) {
  const T ONE = (T)0.0174532, SIX = (T)0.1047192, TWELVE = (T)0.2094384, FIFTY = (T)0.87266;
  if (x < (T)-2.4 || x > (T)2.4 || theta < -TWELVE || theta > TWELVE) return -1;
  int box = (x < (T)-0.8) ? 0 : ((x < (T)0.8) ? 1 : 2);
  box += (x_dot < (T)-0.5) ? 0 : ((x_dot < (T)0.5) ? 3 : 6);
  if (theta < -SIX) box += 0;
  else if (theta < -ONE) box += 9;
  else if (theta < (T)0) box += 18;
  else if (theta < ONE) box += 27;
  else if (theta < SIX) box += 36;
  else box += 45;
  box += (theta_dot < -FIFTY) ? 0 : ((theta_dot < FIFTY) ? 54 : 108);
  return box;
}

__global__ void __launch_bounds__(kThreads) encode_box_f32_kernel(const float4* __restrict__ obs, int64_t n,
                                                                   int32_t* __restrict__ ids) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float4 o = __ldg(obs + i);
    ids[i] = cartpole_box<float>(o.x, o.y, o.z, o.w);
  }
}

__global__ void __launch_bounds__(kThreads) encode_box_f64_kernel(const double2* __restrict__ obs, int64_t n,
                                                                   int32_t* __restrict__ ids) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double2 a = __ldg(obs + 2 * i), b = __ldg(obs + 2 * i + 1);
    ids[i] = cartpole_box<double>(a.x, a.y, b.x, b.y);
  }
}

// ---------------------------------------------------------------------------------------- latent-state MLP
// One row per thread.  The weights are staged once per CTA in shared memory; the hidden vector lives in shared memory
// as h[j][thread] (bank = thread, conflict-free).  The output layer is evaluated EIGHT latent states at a time: W2 is
// staged transposed and padded, sW2t[j][k], so the eight weights of one hidden unit arrive as two 128-bit broadcast
// loads and feed eight independent FMA chains per h[j] load — 3 shared-memory instructions per 8 FMAs where the
// one-output-at-a-time loop issued 16 (it was shared-memory-bound at 11 % of the FP32 FMA rate).
// Arithmetic per output is unchanged: a fixed-order float32 FMA chain (bias first, inputs ascending), so the ids are a
// pure function of the inputs, and argmax keeps the FIRST maximum like torch.max / np.argmax.
constexpr int kMlpThreads = 128;
constexpr int kMlpKB = 8;  // latent states per block of the output layer

__host__ __device__ inline int mlp_padded_latent(int n_latent) { return (n_latent + kMlpKB - 1) / kMlpKB * kMlpKB; }

// AUDIT: also write, per row, the margin between the two largest logits and sum_j |h_j| — what a caller needs to bound the
// effect of the accumulation order on the arg-max (ols_encode_mlp_audit)
template <bool AUDIT>
__global__ void __launch_bounds__(kMlpThreads) encode_mlp_kernel(const float* __restrict__ obs, int64_t n, int d_obs,
                                                                  const float* __restrict__ W1, const float* __restrict__ b1,
                                                                  int hidden, const float* __restrict__ W2,
                                                                  const float* __restrict__ b2, int n_latent, float slope,
                                                                  int32_t* __restrict__ ids, float* __restrict__ margin,
                                                                  float* __restrict__ hnorm) {
  extern __shared__ __align__(16) float smem[];
  const int KP = mlp_padded_latent(n_latent);
  float* sW2t = smem;                      // [hidden][KP], 16-byte aligned rows
  float* sb2 = sW2t + hidden * KP;         // [KP]
  float* sW1 = sb2 + KP;                   // [hidden][d_obs]
  float* sb1 = sW1 + hidden * d_obs;       // [hidden]
  float* sx = sb1 + hidden;                // [d_obs][threads]
  float* sh = sx + d_obs * kMlpThreads;    // [hidden][threads]
  for (int i = threadIdx.x; i < hidden * d_obs; i += blockDim.x) sW1[i] = W1[i];
  for (int i = threadIdx.x; i < hidden; i += blockDim.x) sb1[i] = b1[i];
  for (int i = threadIdx.x; i < hidden * KP; i += blockDim.x) {
    const int j = i / KP, k = i % KP;
    sW2t[i] = (k < n_latent) ? W2[k * hidden + j] : 0.0f;
  }
  for (int i = threadIdx.x; i < KP; i += blockDim.x) sb2[i] = (i < n_latent) ? b2[i] : 0.0f;
  __syncthreads();
  const int t = threadIdx.x;
  for (int64_t base = blockIdx.x * (int64_t)kMlpThreads; base < n; base += (int64_t)gridDim.x * kMlpThreads) {
    // coalesced load of the CTA's observation tile, transposed into sx[d][thread]
    int64_t tile_elems = min((int64_t)kMlpThreads, n - base) * d_obs;
    for (int64_t e = t; e < tile_elems; e += kMlpThreads) {
      int row = (int)(e / d_obs), d = (int)(e % d_obs);
      sx[d * kMlpThreads + row] = __ldg(obs + base * d_obs + e);
    }
    __syncthreads();
    int64_t i = base + t;
    if (i < n) {
      float hn = 0.0f;
      for (int j = 0; j < hidden; ++j) {
        float acc = sb1[j];
        for (int d = 0; d < d_obs; ++d) acc = __fmaf_rn(sx[d * kMlpThreads + t], sW1[j * d_obs + d], acc);
        const float h = acc > 0.0f ? acc : __fmul_rn(acc, slope);
        sh[j * kMlpThreads + t] = h;
        if (AUDIT) hn += fabsf(h);
      }
      int best = 0;
      float best_v = 0.0f, second_v = -INFINITY;
      for (int kb = 0; kb < KP; kb += kMlpKB) {
        float acc[kMlpKB];
        const float4 ba = *(const float4*)(sb2 + kb), bb = *(const float4*)(sb2 + kb + 4);
        acc[0] = ba.x; acc[1] = ba.y; acc[2] = ba.z; acc[3] = ba.w;
        acc[4] = bb.x; acc[5] = bb.y; acc[6] = bb.z; acc[7] = bb.w;
        const float* w = sW2t + kb;
#pragma unroll 4
        for (int j = 0; j < hidden; ++j) {
          const float h = sh[j * kMlpThreads + t];
          const float4 wa = *(const float4*)(w + j * KP), wb = *(const float4*)(w + j * KP + 4);
          acc[0] = __fmaf_rn(h, wa.x, acc[0]);
          acc[1] = __fmaf_rn(h, wa.y, acc[1]);
          acc[2] = __fmaf_rn(h, wa.z, acc[2]);
          acc[3] = __fmaf_rn(h, wa.w, acc[3]);
          acc[4] = __fmaf_rn(h, wb.x, acc[4]);
          acc[5] = __fmaf_rn(h, wb.y, acc[5]);
          acc[6] = __fmaf_rn(h, wb.z, acc[6]);
          acc[7] = __fmaf_rn(h, wb.w, acc[7]);
        }
#pragma unroll
        for (int u = 0; u < kMlpKB; ++u) {
          const int k = kb + u;
          if (k < n_latent && (k == 0 || acc[u] > best_v)) {  // strict >: first maximum wins
            if (AUDIT && k > 0) second_v = best_v;
            best_v = acc[u];
            best = k;
          } else if (AUDIT && k < n_latent && acc[u] > second_v) {
            second_v = acc[u];
          }
        }
      }
      ids[i] = best;
      if (AUDIT) {
        margin[i] = best_v - second_v;  // +inf with a single latent state
        hnorm[i] = hn;
      }
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" int ols_encode_grid(const float* obs, int64_t n, int32_t d_obs, int32_t cells, int32_t* ids, void* stream) {
  OLS_REQUIRE(n >= 0 && d_obs >= 1 && cells >= 1, OLS_ERR_INVALID, "ols_encode_grid: bad sizes");
  if (n == 0) return OLS_OK;
  OLS_REQUIRE(obs && ids, OLS_ERR_INVALID, "ols_encode_grid: null pointer");
  int64_t states = 1;
  for (int d = 0; d < d_obs; ++d) {
    states *= cells;
    OLS_REQUIRE(states < (1ll << 31), OLS_ERR_INVALID, "ols_encode_grid: cells^d_obs overflows int32");
  }
  OLS_REQUIRE(d_obs != 2 || ((uintptr_t)obs % 8 == 0), OLS_ERR_INVALID, "ols_encode_grid: obs must be 8-byte aligned");
  encode_grid_kernel<<<grid_for(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(obs, n, d_obs, cells, ids);
  OLS_LAUNCH_CHECK();
  return OLS_OK;
}

extern "C" int ols_encode_cartpole_box(const void* obs, int32_t obs_is_f64, int64_t n, int32_t* ids, void* stream) {
  OLS_REQUIRE(n >= 0, OLS_ERR_INVALID, "ols_encode_cartpole_box: negative n");
  if (n == 0) return OLS_OK;
  OLS_REQUIRE(obs && ids, OLS_ERR_INVALID, "ols_encode_cartpole_box: null pointer");
  OLS_REQUIRE((uintptr_t)obs % 16 == 0, OLS_ERR_INVALID, "ols_encode_cartpole_box: obs must be 16-byte aligned");
  if (obs_is_f64)
    encode_box_f64_kernel<<<grid_for(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>((const double2*)obs, n, ids);
  else
    encode_box_f32_kernel<<<grid_for(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>((const float4*)obs, n, ids);
  OLS_LAUNCH_CHECK();
  return OLS_OK;
}

static int encode_mlp_impl(const float* obs, int64_t n, int32_t d_obs, const float* W1, const float* b1, int32_t hidden,
                           const float* W2, const float* b2, int32_t n_latent, float negative_slope, int32_t* ids, float* margin,
                           float* hnorm, void* stream, const char* who) {
  OLS_REQUIRE(n >= 0 && d_obs >= 1 && hidden >= 1 && n_latent >= 1, OLS_ERR_INVALID, "%s: bad sizes", who);
  if (n == 0) return OLS_OK;
  OLS_REQUIRE(obs && W1 && b1 && W2 && b2 && ids, OLS_ERR_INVALID, "%s: null pointer", who);
  const int KP = mlp_padded_latent(n_latent);
  size_t smem = sizeof(float) * ((size_t)hidden * d_obs + hidden + (size_t)KP * hidden + KP +
                                 (size_t)(d_obs + hidden) * kMlpThreads);
  OLS_REQUIRE(smem <= 227 * 1024, OLS_ERR_UNSUPPORTED, "%s: weights + activations need %zu B of shared memory (> 227 KB)", who, smem);
  const bool audit = margin != nullptr;
  if (smem > 48 * 1024) {
    if (audit) OLS_CUDA_CHECK(cudaFuncSetAttribute(encode_mlp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else OLS_CUDA_CHECK(cudaFuncSetAttribute(encode_mlp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  int64_t blocks = (n + kMlpThreads - 1) / kMlpThreads;
  int grid = (int)(blocks < OLS_SM_COUNT * 4 ? blocks : OLS_SM_COUNT * 4);
  if (audit)
    encode_mlp_kernel<true><<<grid, kMlpThreads, smem, (cudaStream_t)stream>>>(obs, n, d_obs, W1, b1, hidden, W2, b2, n_latent,
                                                                               negative_slope, ids, margin, hnorm);
  else
    encode_mlp_kernel<false><<<grid, kMlpThreads, smem, (cudaStream_t)stream>>>(obs, n, d_obs, W1, b1, hidden, W2, b2, n_latent,
                                                                                negative_slope, ids, nullptr, nullptr);
  OLS_LAUNCH_CHECK();
  return OLS_OK;
}

extern "C" int ols_encode_mlp(const float* obs, int64_t n, int32_t d_obs, const float* W1, const float* b1,
                              int32_t hidden, const float* W2, const float* b2, int32_t n_latent, float negative_slope,
                              int32_t* ids, void* stream) {
  return encode_mlp_impl(obs, n, d_obs, W1, b1, hidden, W2, b2, n_latent, negative_slope, ids, nullptr, nullptr, stream, "ols_encode_mlp");
}

extern "C" int ols_encode_mlp_audit(const float* obs, int64_t n, int32_t d_obs, const float* W1, const float* b1, int32_t hidden,
                                    const float* W2, const float* b2, int32_t n_latent, float negative_slope, int32_t* ids,
                                    float* margin, float* hnorm, void* stream) {
  OLS_REQUIRE(n == 0 || (margin && hnorm), OLS_ERR_INVALID, "ols_encode_mlp_audit: null margin / hnorm");
  return encode_mlp_impl(obs, n, d_obs, W1, b1, hidden, W2, b2, n_latent, negative_slope, ids, margin, hnorm, stream, "ols_encode_mlp_audit");
}
