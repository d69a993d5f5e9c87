// Host side of the simulator entry points: argument checks, launch geometry, shared-memory staging
// plan, dispatch to the per-record-width instantiations (ols_sim_nv*.cu), plus the two small
// per-replica kernels (init / reset).  Reference interfaces replaced: see ols_sim_kernels.cuh.
#include "ols_sim_kernels.cuh"

using namespace ols_sim;

// ols_learner.cu, ols_eval.cu
cudaError_t ols_launch_learner(const ols_queues_t&, const ols_replicas_t&, const ols_run_t&, const ols_sim::RunLaunch&);
cudaError_t ols_launch_eval(const ols_queues_t&, const ols_replicas_t&, const ols_run_t&, cudaStream_t);

namespace {

__global__ void __launch_bounds__(kThreads) replicas_init_kernel(ols_queues_t q, ols_replicas_t R, const uint64_t* seeds) {
  const int64_t n = R.n_replicas;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    R.init_cursor[r] = 0;
    R.cur_z[r] = 0;
    R.cur_s[r] = 0;
    R.status[r] = OLS_IDLE;
    R.draws[r] = 0;
    R.ties[r] = 0;
    R.pops[r] = 0;
    R.accepted[r] = 0;
    R.ep_index[r] = 0;
    R.ep_len_count[r] = 0;
    R.ep_t[r] = 0;
    R.ep_open[r] = 0;
    R.ep_G[r] = 0.0;
    R.ep_disc[r] = 1.0;
    if (R.rng_mode == OLS_RNG_PCG64) {
      ols_u128 st, inc;
      ols_pcg64_seed(seeds ? seeds[r] : R.seed, &st, &inc);  // np.random.default_rng(seed), psrs.py:20
      uint64_t* p = R.pcg_state + 4 * r;
      p[0] = st.hi;
      p[1] = st.lo;
      p[2] = inc.hi;
      p[3] = inc.lo;
    }
  }
}

__global__ void __launch_bounds__(kThreads) reset_kernel(ols_queues_t q, ols_replicas_t R, const uint8_t* mask) {
  const int64_t n = R.n_replicas;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    if (mask && !mask[r]) continue;
    int ic = R.init_cursor[r], z = R.cur_z[r], s = R.cur_s[r];
    const bool ok = reset_one(q, R, R.ring_mode ? ols_ring_replica_hash(R.ring_seed, (uint64_t)(R.replica_id0 + r)) : 0u, ic, z, s);
    R.init_cursor[r] = ic;
    R.cur_z[r] = z;
    R.cur_s[r] = s;
    R.status[r] = ok ? OLS_RUNNING : OLS_INIT_EMPTY;
    R.ep_open[r] = 0;
  }
}

int check_sim(const ols_queues_t* q, const ols_replicas_t* r, const char* who) {
  OLS_REQUIRE(q && r, OLS_ERR_INVALID, "%s: null argument", who);
  OLS_REQUIRE(q->n_actions >= 1 && q->n_actions <= OLS_MAX_ACTIONS, OLS_ERR_UNSUPPORTED, "%s: n_actions must be in [1, %d]",
              who, OLS_MAX_ACTIONS);
  OLS_REQUIRE(q->rec_bytes == ols_record_bytes(q->n_actions), OLS_ERR_INVALID, "%s: bad rec_bytes", who);
  OLS_REQUIRE(q->n_buckets >= 1 && q->offsets, OLS_ERR_INVALID, "%s: bad queues", who);
  OLS_REQUIRE((uintptr_t)q->records % 32 == 0, OLS_ERR_INVALID, "%s: records must be 32-byte aligned", who);
  OLS_REQUIRE(r->n_replicas >= 0, OLS_ERR_INVALID, "%s: negative n_replicas", who);
  if (r->n_replicas == 0) return OLS_OK;
  OLS_REQUIRE(r->init_cursor && r->cur_z && r->cur_s && r->status && r->draws && r->ties && r->pops &&
                  r->accepted && r->ep_index && r->ep_len_count && r->ep_t && r->ep_open && r->ep_G && r->ep_disc,
              OLS_ERR_INVALID, "%s: null replica buffer", who);
  OLS_REQUIRE(r->rng_mode == OLS_RNG_FEED || r->rng_mode == OLS_RNG_PCG64 || r->rng_mode == OLS_RNG_PHILOX, OLS_ERR_INVALID,
              "%s: bad rng_mode", who);
  OLS_REQUIRE(r->rng_mode != OLS_RNG_PCG64 || r->pcg_state, OLS_ERR_INVALID, "%s: PCG64 mode needs pcg_state", who);
  OLS_REQUIRE(r->rng_mode != OLS_RNG_FEED || r->u_feed || r->u_feed_len == 0, OLS_ERR_INVALID, "%s: FEED mode needs u_feed", who);
  OLS_REQUIRE(r->tie_mode == OLS_RNG_FEED || r->tie_mode == OLS_RNG_PHILOX, OLS_ERR_INVALID, "%s: bad tie_mode", who);
  return OLS_OK;
}

bool valid_lanes(int g) { return g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32; }

inline int align16(int64_t x) { return (int)((x + 15) / 16 * 16); }

// Shared-memory staging: each lane group of a CTA gets one slot for its replica's cursor table and,
// if that still fits, its Q table.  Budget 96 KB per CTA keeps two CTAs resident per SM.
StagePlan make_stage_plan(const ols_queues_t* q, const ols_run_t* c) {
  StagePlan p = {0, 0, 0};
  if (!c->stage_smem || c->state_aos) return p;  // the AoS entries are read straight from HBM, one sector per state
  const int groups = kThreads / c->lanes;
  const int64_t budget = 96 * 1024 - (int64_t)sizeof(Cold) * groups;
  const int64_t cb = align16((int64_t)q->n_buckets * 4);
  const int64_t qb = (c->agent == OLS_AGENT_FIXED_PI)
                         ? 0
                         : align16((int64_t)c->n_states * q->n_actions * (c->q_is_f64 ? 8 : 4));
  if (groups * (cb + qb) <= budget) {
    p.cursor_bytes = (int)cb;
    p.q_bytes = (int)qb;
  } else if (groups * cb <= budget) {
    p.cursor_bytes = (int)cb;
  }
  p.slot_bytes = p.cursor_bytes + p.q_bytes;
  return p;
}

}  // namespace

extern "C" int ols_replicas_init(const ols_queues_t* q, const ols_replicas_t* r, const uint64_t* seeds, void* stream) {
  int rc = check_sim(q, r, "ols_replicas_init");
  if (rc) return rc;
  if (r->n_replicas == 0) return OLS_OK;
  if (r->cursor)  // NULL when the block only ever runs with ols_run_t.state_aos (the caller zeroes that table)
    OLS_CUDA_CHECK(cudaMemsetAsync(r->cursor, 0, sizeof(int32_t) * (size_t)r->n_replicas * q->n_buckets, (cudaStream_t)stream));
  int blocks = ols_div_up(r->n_replicas, kThreads);
  if (blocks > OLS_SM_COUNT * 8) blocks = OLS_SM_COUNT * 8;
  replicas_init_kernel<<<blocks, kThreads, 0, (cudaStream_t)stream>>>(*q, *r, seeds);
  OLS_LAUNCH_CHECK();
  return OLS_OK;
}

extern "C" int ols_reset(const ols_queues_t* q, const ols_replicas_t* r, const uint8_t* mask, void* stream) {
  int rc = check_sim(q, r, "ols_reset");
  if (rc) return rc;
  if (r->n_replicas == 0) return OLS_OK;
  OLS_REQUIRE(q->n_init == 0 || (q->init_z && q->init_s), OLS_ERR_INVALID, "ols_reset: null init arrays");
  int blocks = ols_div_up(r->n_replicas, kThreads);
  if (blocks > OLS_SM_COUNT * 8) blocks = OLS_SM_COUNT * 8;
  reset_kernel<<<blocks, kThreads, 0, (cudaStream_t)stream>>>(*q, *r, mask);
  OLS_LAUNCH_CHECK();
  return OLS_OK;
}

static int step_impl(const ols_queues_t* q, const ols_replicas_t* r, const uint8_t* mask, const void* p_new, int32_t p_new_is_f64,
                     int64_t p_stride, int32_t reject_mode, int32_t compute_mode, int32_t lanes, const ols_step_out_t* out,
                     const int32_t* act, void* stream, const char* who) {
  int rc = check_sim(q, r, who);
  if (rc) return rc;
  OLS_REQUIRE(valid_lanes(lanes), OLS_ERR_INVALID, "%s: lanes must be 1, 2, 4, 8, 16 or 32", who);
  OLS_REQUIRE(reject_mode == OLS_REJECT_DEFAULT || reject_mode == OLS_REJECT_ACCEPT_ALL, OLS_ERR_INVALID, "%s: bad reject_mode", who);
  OLS_REQUIRE(compute_mode == OLS_COMPUTE_F64 || compute_mode == OLS_COMPUTE_F32, OLS_ERR_INVALID, "%s: bad compute_mode", who);
  if (r->n_replicas == 0) return OLS_OK;
  OLS_REQUIRE(r->cursor, OLS_ERR_INVALID, "%s: null cursor table", who);
  OLS_REQUIRE((p_new || act) && out && out->row && out->action && out->reward && out->terminal && out->s_next && out->z_from,
              OLS_ERR_INVALID, "%s: null input / output buffer", who);
  OLS_REQUIRE(p_stride == 0 || p_stride >= q->n_actions, OLS_ERR_INVALID, "%s: p_stride must be 0 or >= n_actions", who);
  OLS_REQUIRE(q->records, OLS_ERR_INVALID, "%s: queues not bucketed", who);
  const int grid = (int)((r->n_replicas * lanes + kThreads - 1) / kThreads);  // one lane group per replica
  cudaError_t e;
  switch ((q->n_actions + 3) / 4) {
    case 1: e = ols_launch_step_nv1(*q, *r, mask, p_new, p_new_is_f64, p_stride, reject_mode, compute_mode, lanes, *out, grid, (cudaStream_t)stream, act); break;
    case 2: e = ols_launch_step_nv2(*q, *r, mask, p_new, p_new_is_f64, p_stride, reject_mode, compute_mode, lanes, *out, grid, (cudaStream_t)stream, act); break;
    case 3: e = ols_launch_step_nv3(*q, *r, mask, p_new, p_new_is_f64, p_stride, reject_mode, compute_mode, lanes, *out, grid, (cudaStream_t)stream, act); break;
    default: e = ols_launch_step_nv4(*q, *r, mask, p_new, p_new_is_f64, p_stride, reject_mode, compute_mode, lanes, *out, grid, (cudaStream_t)stream, act); break;
  }
  OLS_REQUIRE(e == cudaSuccess, OLS_ERR_CUDA, "%s: launch failed: %s", who, cudaGetErrorString(e));
  ols_note_launch(1);
  return OLS_OK;
}

extern "C" int ols_step(const ols_queues_t* q, const ols_replicas_t* r, const uint8_t* mask, const void* p_new,
                        int32_t p_new_is_f64, int64_t p_stride, int32_t reject_mode, int32_t compute_mode, int32_t lanes,
                        const ols_step_out_t* out, void* stream) {
  OLS_REQUIRE(p_new != nullptr || (r && r->n_replicas == 0), OLS_ERR_INVALID, "ols_step: null p_new");
  return step_impl(q, r, mask, p_new, p_new_is_f64, p_stride, reject_mode, compute_mode, lanes, out, nullptr, stream, "ols_step");
}

extern "C" int ols_step_action(const ols_queues_t* q, const ols_replicas_t* r, const uint8_t* mask, const int32_t* action,
                               const ols_step_out_t* out, void* stream) {
  OLS_REQUIRE(action != nullptr || (r && r->n_replicas == 0), OLS_ERR_INVALID, "ols_step_action: null action");
  OLS_REQUIRE(q && q->n_buckets % (q->n_actions > 0 ? q->n_actions : 1) == 0, OLS_ERR_INVALID,
              "ols_step_action: queues must be keyed by (z, a): n_buckets = n_state_keys * n_actions");
  return step_impl(q, r, mask, nullptr, 0, 0, OLS_REJECT_ACCEPT_ALL, OLS_COMPUTE_F64, 1, out, action, stream, "ols_step_action");
}

extern "C" size_t ols_run_sched_bytes(int64_t n_replicas, int32_t lanes) {
  if (n_replicas < 0 || !valid_lanes(lanes)) return 0;
  const int threads = (lanes == 1 && OLS_G1_THREADS < kThreads) ? OLS_G1_THREADS : kThreads;  // the smallest CTA a run may use
  return sizeof(int32_t) * (size_t)((n_replicas * lanes + threads - 1) / threads + 1);
}

extern "C" int ols_run(const ols_queues_t* q, const ols_replicas_t* r, const ols_run_t* c, void* stream) {
  int rc = check_sim(q, r, "ols_run");
  if (rc) return rc;
  OLS_REQUIRE(c != nullptr, OLS_ERR_INVALID, "ols_run: null run config");
  OLS_REQUIRE(c->agent == OLS_AGENT_FIXED_PI || c->agent == OLS_AGENT_EXPSARSA || c->agent == OLS_AGENT_QLEARN, OLS_ERR_INVALID,
              "ols_run: bad agent");
  OLS_REQUIRE(valid_lanes(c->lanes), OLS_ERR_INVALID, "ols_run: lanes must be 1, 2, 4, 8, 16 or 32");
  OLS_REQUIRE(c->n_states >= 1 && c->n_episodes >= 0 && c->ep_cap >= c->n_episodes + 1, OLS_ERR_INVALID,
              "ols_run: need n_states >= 1 and ep_cap >= n_episodes + 1");
  OLS_REQUIRE(c->reject_mode == OLS_REJECT_DEFAULT || c->reject_mode == OLS_REJECT_ACCEPT_ALL, OLS_ERR_INVALID, "ols_run: bad reject_mode");
  OLS_REQUIRE(c->compute_mode == OLS_COMPUTE_F64 || c->compute_mode == OLS_COMPUTE_F32, OLS_ERR_INVALID, "ols_run: bad compute_mode");
  OLS_REQUIRE(c->pops_per_lane == 0 || c->pops_per_lane == 1 || c->pops_per_lane == 2 || c->pops_per_lane == 4, OLS_ERR_INVALID,
              "ols_run: pops_per_lane must be 0, 1, 2 or 4");
  OLS_REQUIRE(!(c->pops_per_lane >= 2 && r->rng_mode == OLS_RNG_PCG64), OLS_ERR_UNSUPPORTED,
              "ols_run: PCG64 replay examines one record per lane per round");
  if (r->n_replicas == 0) return OLS_OK;
  OLS_REQUIRE(c->Gs && c->lengths, OLS_ERR_INVALID, "ols_run: null Gs / lengths");
  OLS_REQUIRE(q->records, OLS_ERR_INVALID, "ols_run: queues not bucketed");
  OLS_REQUIRE(q->n_init == 0 || (q->init_z && q->init_s), OLS_ERR_INVALID, "ols_run: null init arrays");
  OLS_REQUIRE(r->cursor || c->state_aos, OLS_ERR_INVALID, "ols_run: null cursor table");
  if (c->state_aos) {
    OLS_REQUIRE(c->agent != OLS_AGENT_FIXED_PI && !c->q_is_f64 && q->n_actions <= 4 && q->z_offset == 0 &&
                    c->n_states == q->n_buckets && q->s_next == nullptr && r->rng_mode != OLS_RNG_PCG64 &&
                    c->reject_mode == OLS_REJECT_DEFAULT && (c->pops_per_lane == 0 || c->pops_per_lane == 2),
                OLS_ERR_UNSUPPORTED,
                "ols_run: state_aos needs a float32 learner, <= 4 actions, z_offset 0, n_states == n_buckets, s == z, "
                "Philox/fed uniforms, default reject rule, 2 records per lane");
    OLS_REQUIRE((uintptr_t)c->state_aos % 32 == 0 && c->alpha, OLS_ERR_INVALID, "ols_run: state_aos must be 32-byte aligned");
  } else if (c->agent != OLS_AGENT_FIXED_PI) {
    OLS_REQUIRE(c->Q && c->alpha, OLS_ERR_INVALID, "ols_run: learner needs Q and alpha");
  }
  if (c->agent == OLS_AGENT_QLEARN) OLS_REQUIRE(c->epsilon, OLS_ERR_INVALID, "ols_run: qlearn needs epsilon");
  if (c->agent != OLS_AGENT_QLEARN) OLS_REQUIRE(c->pi, OLS_ERR_INVALID, "ols_run: fixed-policy agents need pi");
  OLS_REQUIRE((!c->trace_row && !c->td_error) || c->trace_cap > 0, OLS_ERR_INVALID, "ols_run: trace taps need trace_cap > 0");
  RunLaunch L;
  // TAIL hot records serve the one-thread-per-replica float32 Q-learner only (ols_learner.cu); every other kernel pops from
  // the full records (PROBS hot records: the lane-group AoS learners and fixed-policy evaluation, see ols_sim_kernels.cuh)
  const bool hot_tail = q->hot && q->hot_kind == OLS_HOT_TAIL;
  const bool fast_learner = hot_tail && q->queue_dist && !q->p_log64 && c->agent == OLS_AGENT_QLEARN && c->lanes == 1 && c->state_aos &&
                            r->rng_mode == OLS_RNG_PHILOX;
  // ... and fixed-policy evaluation (run_kernel<..., HOT = 2>)
  const bool tail_eval = hot_tail && q->queue_dist && !q->p_log64 && c->agent == OLS_AGENT_FIXED_PI && r->rng_mode == OLS_RNG_PHILOX &&
                         c->reject_mode == OLS_REJECT_DEFAULT;
  ols_queues_t q_plain = *q;
  if (hot_tail && !fast_learner && !tail_eval) {
    q_plain.hot = nullptr;
    q = &q_plain;
  }
  OLS_REQUIRE(c->batch_lanes >= 0 && c->batch_lanes <= 32, OLS_ERR_INVALID, "ols_run: batch_lanes must be in [0, 32]");
  // (the one-lane hot-record learners are built for OLS_MINB_G1 CTAs of OLS_G1_THREADS per SM, see their launch bounds)
  const bool g1_hot = c->lanes == 1 && c->state_aos && q->hot && r->rng_mode == OLS_RNG_PHILOX;
  L.threads = g1_hot ? OLS_G1_THREADS : kThreads;
  const int n_blocks = (int)((r->n_replicas * c->lanes + L.threads - 1) / L.threads);  // one lane group per replica
  L.plan = make_stage_plan(q, c);
  // ---- time slicing (see run_kernel): only a budgeted run has slices to hand out
  OLS_REQUIRE(c->slices >= -1 && c->slices <= 4096, OLS_ERR_INVALID, "ols_run: slices must be -1 (auto), 0/1 (off) or 2..4096");
  OLS_REQUIRE(c->slices <= 1 || (c->sched && c->max_steps > 0), OLS_ERR_INVALID,
              "ols_run: time slicing needs max_steps > 0 and the sched scratch");
  int slices = c->slices > 1 ? c->slices : 1;
  if (c->slices == -1 && c->sched && c->max_steps > 0 && L.plan.slot_bytes == 0) {
    // resident CTAs per wave under the kernels' launch bounds; slice when the last wave would run part-empty
    const int resident = OLS_SM_COUNT * (q->n_actions <= 4 ? (fast_learner ? OLS_LEARNER_MINB : g1_hot ? OLS_MINB_G1 : OLS_MINB) : 2);
    const int full = n_blocks / resident, rest = n_blocks % resident;
    // the one-lane learner is also sliced when every block is resident at once: its 6.6 CTAs per SM are 6 on some SMs and 7
    // on others, and slices re-deal the blocks to whichever SM has room (C4: 207.9 ms unsliced, 187.8 ms in 16 slices)
    if ((full >= 1 && rest != 0 && rest < resident * 9 / 10) || (g1_hot && n_blocks > OLS_SM_COUNT)) {
      int64_t k = c->max_steps / 512;  // at least 512 accepted transitions per slice
      slices = (int)(k < 1 ? 1 : k > 16 ? 16 : k);
    }
  }
  ols_run_t cfg = *c;
  cfg.slices = slices;
  c = &cfg;
  if (slices > 1) OLS_CUDA_CHECK(cudaMemsetAsync(cfg.sched, 0, sizeof(int32_t) * (size_t)(n_blocks + 1), (cudaStream_t)stream));
  L.grid = n_blocks * slices;
  L.smem = (sizeof(Cold) + (size_t)L.plan.slot_bytes) * (L.threads / c->lanes);
  L.keys = ols_philox_make_keys(r->seed);
  L.stream = (cudaStream_t)stream;
  int ppl = c->pops_per_lane;
  if (ppl == 0) ppl = (r->rng_mode == OLS_RNG_PCG64 || c->reject_mode == OLS_REJECT_ACCEPT_ALL) ? 1 : 2;
  cudaError_t e;
  if (tail_eval && c->lanes == 32 && slices <= 1 && !q->s_next && r->cursor) {
    // the streaming evaluator: one warp per replica, warp-uniform control flow (ols_eval.cu)
    e = ols_launch_eval(*q, *r, *c, (cudaStream_t)stream);
    OLS_REQUIRE(e == cudaSuccess, OLS_ERR_CUDA, "ols_run: launch failed: %s", cudaGetErrorString(e));
    return OLS_OK;
  }
  if (fast_learner) {
    L.smem = 0;
    e = ols_launch_learner(*q, *r, *c, L);
    OLS_REQUIRE(e == cudaSuccess, OLS_ERR_CUDA, "ols_run: launch failed: %s", cudaGetErrorString(e));
    return OLS_OK;
  }
  switch ((q->n_actions + 3) / 4) {
    case 1: e = ols_launch_run_nv1(*q, *r, *c, L, ppl); break;
    case 2: e = ols_launch_run_nv2(*q, *r, *c, L, ppl); break;
    case 3: e = ols_launch_run_nv3(*q, *r, *c, L, ppl); break;
    default: e = ols_launch_run_nv4(*q, *r, *c, L, ppl); break;
  }
  OLS_REQUIRE(e == cudaSuccess, OLS_ERR_CUDA, "ols_run: launch failed: %s", cudaGetErrorString(e));
  ols_note_launch(1);
  return OLS_OK;
}
