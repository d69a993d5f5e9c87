// Instantiations of the PSRS step/run kernels for records with 3 x 16 B of logged distribution
// (n_actions in [9, 12]); split per width so the translation units compile in parallel.
#include "ols_sim_kernels.cuh"

cudaError_t ols_launch_run_nv3(const ols_queues_t& q, const ols_replicas_t& r, const ols_run_t& c,
                               const ols_sim::RunLaunch& L, int ppl) {
  return ols_sim::launch_run_nv<3>(q, r, c, L, ppl);
}

cudaError_t ols_launch_step_nv3(OLS_STEP_ARGS) {
  return ols_sim::launch_step_nv<3>(q, r, mask, p_new, p_new_is_f64, p_stride, reject_mode, compute_mode, lanes, out, grid,
                                    stream, act);
}
