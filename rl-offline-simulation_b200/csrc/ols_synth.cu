// Synthetic logged data of the benchmark shape, generated on the device.
// Counter-based (Philox4x32-10 keyed by the seed, counter = row), so every GPU of a job regenerates
// the identical log without communication.  Not a reference function: it stands in for the data
// collection scripts (examples/*/random_agent_rollout.py) that are out of scope.
#include "ols_common.cuh"

namespace {

__device__ __forceinline__ uint8_t row_terminal(uint64_t seed, int64_t i, double p_done) {
  return ols_philox_uniform(seed, (uint64_t)i, 3u, 1) < p_done ? 1 : 0;
}

__global__ void __launch_bounds__(256) synth_kernel(ols_log_t log, const float* __restrict__ pb_state, int n_states,
                                                    uint64_t seed, double p_done) {
  const int nA = log.n_actions;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < log.n_rows; i += (int64_t)gridDim.x * blockDim.x) {
    ols_u4 c = {(uint32_t)i, (uint32_t)((uint64_t)i >> 32), 0x5EEDu, 0u};
    ols_u4 o = ols_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const int z = (int)(((uint64_t)o.x * (uint64_t)n_states) >> 32);
    const int zn = (int)(((uint64_t)o.y * (uint64_t)n_states) >> 32);
    const float ua = (float)(o.z >> 8) * (1.0f / 16777216.0f);
    const float* pb = pb_state + (size_t)z * nA;
    float cdf = 0.f;
    int a = nA - 1;
    float* pl = (float*)log.p_log + i * nA;
    for (int j = 0; j < nA; ++j) {
      float pj = __ldg(pb + j);
      pl[j] = pj;
      cdf += pj;
      if (a == nA - 1 && ua < cdf) a = j;
    }
    // reward ~ N(0,1) by Box-Muller on two more uniforms
    const double u1 = ols_philox_uniform(seed, (uint64_t)i, 3u, 2);
    const double u2 = ols_philox_uniform(seed, (uint64_t)i, 3u, 3);
    const float rw = (float)(sqrt(-2.0 * log1p(-u1)) * cospi(2.0 * u2));
    const uint8_t term = row_terminal(seed, i, p_done);
    ((int32_t*)log.z)[i] = z;
    ((int32_t*)log.z_next)[i] = zn;
    if (log.s) ((int32_t*)log.s)[i] = z;
    if (log.s_next) ((int32_t*)log.s_next)[i] = zn;
    ((int32_t*)log.action)[i] = a;
    ((float*)log.reward)[i] = rw;
    ((uint8_t*)log.terminal)[i] = term;
    if (log.step) ((int32_t*)log.step)[i] = (i == 0 || row_terminal(seed, i - 1, p_done)) ? 0 : 1;
  }
}

}  // namespace

extern "C" int ols_synth_log(const ols_log_t* log, const float* pb_state, int32_t n_states, uint64_t seed, double p_done,
                             void* stream) {
  OLS_REQUIRE(log && pb_state, OLS_ERR_INVALID, "ols_synth_log: null argument");
  OLS_REQUIRE(log->n_rows >= 0 && log->n_rows < (1ll << 31) && n_states >= 1, OLS_ERR_INVALID, "ols_synth_log: bad sizes");
  OLS_REQUIRE(log->n_actions >= 1 && log->n_actions <= OLS_MAX_ACTIONS && !log->p_is_f64, OLS_ERR_UNSUPPORTED,
              "ols_synth_log: float32 p_log with at most %d actions", OLS_MAX_ACTIONS);
  if (log->n_rows == 0) return OLS_OK;
  OLS_REQUIRE(log->z && log->z_next && log->action && log->reward && log->terminal && log->p_log, OLS_ERR_INVALID,
              "ols_synth_log: null column");
  int64_t blocks = (log->n_rows + 255) / 256;
  if (blocks > OLS_SM_COUNT * 16) blocks = OLS_SM_COUNT * 16;
  synth_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(*log, pb_state, n_states, seed, p_done);
  OLS_LAUNCH_CHECK();
  return OLS_OK;
}
