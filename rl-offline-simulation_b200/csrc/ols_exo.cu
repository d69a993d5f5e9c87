// PSRS with exogenous state: two queue families popped in lock-step.
//
// Replaces offsim4rl/evaluators/psrs.py:59-117 `PSRS_Exo.step` (and supports its `reset`): an observation o splits
// into an endogenous state s and an exogenous state x (`o_split_func`, psrs.py:64,75,103).  The logged transitions are
// bucketed twice — by s (records carry a, r, s', done, p_b; psrs.py:82) and by x (records carry x'; psrs.py:83).  One
// pop consumes the head of BOTH queues (psrs.py:104-109) and is accepted by the usual rule on the s-record
// (psrs.py:112-115); the next observation combines s' of the accepted s-record with x' of the x-record popped with it
// (psrs.py:116).  The simulation ends when either queue is empty.
//
// Same lane-group machinery as ols_step: lane k examines the k-th pending pair with the uniform of draw (draws + k).
#include "ols_sim_kernels.cuh"

using namespace ols_sim;

namespace {

template <int NV, int RNG>
__global__ void __launch_bounds__(kThreads) exo_step_kernel(ols_queues_t qs, ols_queues_t qx, ols_replicas_t R, int32_t* cursor_x,
                                                             int32_t* cur_x, const uint8_t* mask, const void* p_new,
                                                             int p_new_is_f64, int64_t p_stride, int compute_mode, int G,
                                                             ols_step_out_t out, int32_t* out_x_next, ols_philox_keys PK) {
  constexpr int NA = 4 * NV;
  constexpr int kRec = 16 * (NV + 1);
  const Lane L = make_lane(G);
  const int nA = qs.n_actions;
  const int64_t r = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) * (32 / G) + (L.lane / G);
  const bool exists = r < R.n_replicas;
  const int64_t rr = exists ? r : 0;
  const uint64_t rid = (uint64_t)(R.replica_id0 + rr);
  const int64_t prow = rr * p_stride;
  auto pn_fn = [=](int j) -> double {
    return p_new_is_f64 ? __ldg((const double*)p_new + prow + j) : (double)__ldg((const float*)p_new + prow + j);
  };
  Policy<NA> P;
  finish_policy<NA>(P, pn_fn, nA);
  int status = OLS_IDLE, s_cur = 0, x_cur = 0;
  uint64_t draws = 0;
  ols_u128 pcg_s = {0, 0}, pcg_inc = {0, 0};
  bool active = false;
  if (exists) {
    status = R.status[r];
    s_cur = R.cur_z[r];
    x_cur = cur_x[r];
    draws = R.draws[r];
    if (RNG == RNGT_PCG64) {
      const uint64_t* p = R.pcg_state + 4 * r;
      pcg_s.hi = p[0]; pcg_s.lo = p[1]; pcg_inc.hi = p[2]; pcg_inc.lo = p[3];
    }
    active = (!mask || mask[r]) && (status == OLS_RUNNING || status == OLS_QUEUE_EMPTY);
  }
  const bool was_active = active;
  int32_t* tab_s = R.cursor + rr * (int64_t)qs.n_buckets;
  int32_t* tab_x = cursor_x + rr * (int64_t)qx.n_buckets;
  Bucket Bs, Bx;
  Bs.b = Bx.b = -1;
  Bs.lo = Bs.len = Bs.cur = Bx.lo = Bx.len = Bx.cur = 0;
  Bs.start = Bx.start = 0;
  int res = kPopIdle;
  if (active && (!load_bucket(qs, R, 0u, tab_s, s_cur, Bs) || !load_bucket(qx, R, 0u, tab_x, x_cur, Bx))) {
    res = OLS_NO_QUEUE;  // KeyError: psrs.py:104,106 index s_queues[s] / x_queues[x]
    active = false;
  }
  uint32_t n_pop = 0;
  float4 win_s = make_float4(0.f, 0.f, 0.f, 0.f);
  int win_xn = 0;
  while (__any_sync(kFull, active)) {
    const int remaining = min(Bs.len - Bs.cur, Bx.len - Bx.cur);
    if (active && remaining <= 0) {
      res = OLS_QUEUE_EMPTY;  // psrs.py:104-107
      active = false;
    }
    const int n_eval = active ? min(G, remaining) : 0;
    const bool valid = L.sub < n_eval;
    float4 pv[NV], tail = make_float4(0.f, 0.f, 0.f, 0.f);
    int xn = 0;
#pragma unroll
    for (int v = 0; v < NV; ++v) pv[v] = make_float4(1.f, 1.f, 1.f, 1.f);
    if (valid) {
      const float4* rp = (const float4*)((const char*)qs.records + (size_t)(uint32_t)(Bs.lo + Bs.cur + L.sub) * kRec);
#pragma unroll
      for (int v = 0; v < NV; ++v) pv[v] = __ldg(rp + v);
      tail = __ldg(rp + NV);
      const float4* xp = (const float4*)((const char*)qx.records + (size_t)(uint32_t)(Bx.lo + Bx.cur + L.sub) * kRec);
      xn = __float_as_int(__ldg(xp + NV).y);  // the x-queue record's "z_next" field is x'
    }
    double u = 0.0;
    bool lack = false;
    ols_u128 stepped = pcg_s;
    if (valid) {
      const uint64_t d = draws + (uint64_t)L.sub;
      if (RNG == RNGT_FEED) {
        if ((int64_t)d < R.u_feed_len) u = __ldg(R.u_feed + rr * R.u_stride + (int64_t)d);
        else lack = true;
      } else if (RNG == RNGT_PCG64) {
        for (int t = 0; t <= L.sub; ++t) stepped = ols_pcg_step(stepped, pcg_inc);
        u = ols_u64_to_unit(ols_pcg_output(stepped));
      } else {
        u = ols_stream_uniform(PK, rid, 0u, d);
      }
    }
    bool acc = false;
    if (valid && !lack) {
      const uint32_t meta = __float_as_uint(tail.w);
      const int a = ols_meta_action(meta);
      float plf[NA];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        plf[4 * v + 0] = pv[v].x; plf[4 * v + 1] = pv[v].y; plf[4 * v + 2] = pv[v].z; plf[4 * v + 3] = pv[v].w;
      }
      const float uf = (float)u;
      int dec = accept_filter<NA>(P, plf, a, uf, P.ok && uf >= 8.6736174e-19f && ols_meta_tame(meta) != 0);
      if (dec == 2) {
        const double* p64 = qs.p_log64 ? qs.p_log64 + (int64_t)__float_as_int(tail.z) * nA : nullptr;
        dec = (int)accept_slow(pn_fn, (const float*)((const char*)qs.records + (size_t)(uint32_t)(Bs.lo + Bs.cur + L.sub) * kRec),
                               p64, nA, a, u, compute_mode);
      }
      acc = (dec == 1);  // psrs.py:114 `u <= thr` == !(u > thr)
    }
    const unsigned ba = (__ballot_sync(kFull, acc) >> L.gbase) & L.gmask;
    const unsigned bl = (__ballot_sync(kFull, lack) >> L.gbase) & L.gmask;
    const int first_acc = ba ? __ffs(ba) - 1 : (1 << 20), first_lack = bl ? __ffs(bl) - 1 : (1 << 20);
    int consumed = 0;
    bool took = false;
    if (active) {
      if (first_acc < first_lack) { consumed = first_acc + 1; took = true; }
      else if (first_lack < (1 << 20)) { consumed = first_lack; res = OLS_FEED_EMPTY; active = false; }
      else consumed = n_eval;
    }
    const int src = L.gbase + (first_acc < (1 << 20) ? first_acc : 0);
    float4 sel = tail;
    sel.x = __shfl_sync(kFull, sel.x, src);
    sel.y = __shfl_sync(kFull, sel.y, src);
    sel.z = __shfl_sync(kFull, sel.z, src);
    sel.w = __shfl_sync(kFull, sel.w, src);
    const int sel_xn = __shfl_sync(kFull, xn, src);
    if (RNG == RNGT_PCG64) {
      const int adv_src = L.gbase + (consumed > 0 ? consumed - 1 : 0);
      const uint64_t nh = __shfl_sync(kFull, stepped.hi, adv_src);
      const uint64_t nl = __shfl_sync(kFull, stepped.lo, adv_src);
      if (consumed > 0) { pcg_s.hi = nh; pcg_s.lo = nl; }
    }
    Bs.cur += consumed;
    Bx.cur += consumed;
    n_pop += (uint32_t)consumed;
    draws += (uint64_t)consumed;
    if (took) {
      win_s = sel;
      win_xn = sel_xn;
      res = OLS_RUNNING;
      active = false;
    }
  }
  if (exists && L.sub == 0) {
    if (was_active) {
      if (Bs.b >= 0) tab_s[Bs.b] = Bs.cur;
      if (Bx.b >= 0) tab_x[Bx.b] = Bx.cur;
      R.draws[r] = draws;
      R.pops[r] += n_pop;
      if (RNG == RNGT_PCG64) {
        R.pcg_state[4 * r] = pcg_s.hi;
        R.pcg_state[4 * r + 1] = pcg_s.lo;
      }
      if (res == OLS_RUNNING) {
        const uint32_t meta = __float_as_uint(win_s.w);
        R.status[r] = OLS_RUNNING;
        R.cur_z[r] = __float_as_int(win_s.y);  // s'
        cur_x[r] = win_xn;                     // x'
        R.accepted[r] += 1;
        out.row[r] = __float_as_int(win_s.z);
        out.action[r] = ols_meta_action(meta);
        out.reward[r] = win_s.x;
        out.terminal[r] = (uint8_t)ols_meta_terminal(meta);
        out.s_next[r] = __float_as_int(win_s.y);
        out.z_from[r] = s_cur;
        out_x_next[r] = win_xn;
      } else {
        R.status[r] = res;
        out.row[r] = -1;
      }
    } else if (!mask || mask[r]) {
      out.row[r] = -1;
    }
  }
}

__global__ void reseed_kernel(ols_replicas_t R, const uint64_t* seeds, const uint8_t* mask) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < R.n_replicas; r += (int64_t)gridDim.x * blockDim.x) {
    if (mask && !mask[r]) continue;
    ols_u128 st, inc;
    ols_pcg64_seed(seeds ? seeds[r] : R.seed, &st, &inc);
    uint64_t* p = R.pcg_state + 4 * r;
    p[0] = st.hi; p[1] = st.lo; p[2] = inc.hi; p[3] = inc.lo;
    R.draws[r] = 0;
  }
}

// ---------------------------------------------------------------------------------------- fused multi-step run
// evalMC_psrs (psrs.py:241-271) over a PSRS_Exo environment (psrs.py:59-117) in ONE launch, one thread per replica: every
// episode starts with reset(seed = episode index) — the PCG64 generator is re-seeded there (psrs.py:92), with
// reseed_base[r] + episode —, the agent's policy row is pi[o] for the observation o = s * x_mult + x (the linear
// o_combine_func; anything else stays on the host loop), and a step pops the heads of the s- and the x-queue in lock-step until
// the s-record is accepted.  With Philox uniforms (the batched default) there is no re-seeding: the replica's stream runs on.
template <int NV, int RNG>
__global__ void __launch_bounds__(128) exo_run_kernel(ols_queues_t qs, ols_queues_t qx, ols_replicas_t R, int32_t* cursor_x, int32_t* cur_x,
                                                      ols_exo_run_t cfg, ols_philox_keys PK) {
  constexpr int NA = 4 * NV;
  constexpr int kRec = 16 * (NV + 1);
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= R.n_replicas) return;
  const int nA = qs.n_actions;
  const uint64_t rid = (uint64_t)(R.replica_id0 + r);
  int status = R.status[r];
  if (!(status == OLS_IDLE || status == OLS_RUNNING)) return;
  int s_cur = R.cur_z[r], x_cur = cur_x[r], init_cursor = R.init_cursor[r];
  int ep_index = R.ep_index[r], lc = R.ep_len_count[r], ep_t = R.ep_t[r], ep_open = R.ep_open[r];
  double G = R.ep_G[r], disc = R.ep_disc[r];
  uint64_t draws = R.draws[r];
  unsigned long long accepted = R.accepted[r], pops = 0;
  long long done_now = 0;
  ols_u128 pcg_s = {0, 0}, pcg_inc = {0, 0};
  if (RNG == RNGT_PCG64) {
    const uint64_t* p = R.pcg_state + 4 * r;
    pcg_s.hi = p[0]; pcg_s.lo = p[1]; pcg_inc.hi = p[2]; pcg_inc.lo = p[3];
  }
  int32_t* tab_s = R.cursor + r * (int64_t)qs.n_buckets;
  int32_t* tab_x = cursor_x + r * (int64_t)qx.n_buckets;
  status = OLS_RUNNING;
  for (;;) {
    if (!ep_open) {  // psrs.py:246-253 / 91-97
      if (ep_index >= cfg.n_episodes) { status = OLS_DONE; break; }
      if (RNG == RNGT_PCG64) {
        ols_pcg64_seed((cfg.reseed_base ? cfg.reseed_base[r] : 0ull) + (uint64_t)ep_index, &pcg_s, &pcg_inc);
        draws = 0;
      }
      if ((int64_t)init_cursor >= qs.n_init) { status = OLS_INIT_EMPTY; break; }
      s_cur = __ldg(qs.init_z + init_cursor);
      x_cur = __ldg(qx.init_z + init_cursor);
      init_cursor += 1;
      ep_open = 1; ep_t = 0; G = 0.0; disc = 1.0;
    }
    const long long o = (long long)s_cur * cfg.x_mult + x_cur;
    if (o < 0 || o >= cfg.n_obs) { status = OLS_NO_QUEUE; break; }  // IndexError on pi[S]
    const void* pi_tab = cfg.pi;
    const int is64 = cfg.pi_is_f64;
    const int64_t off = o * nA;
    auto pn_fn = [=](int j) -> double {
      return is64 ? __ldg((const double*)pi_tab + off + j) : (double)__ldg((const float*)pi_tab + off + j);
    };
    Policy<NA> P;
    finish_policy<NA>(P, pn_fn, nA);
    // ---- env.step(p): psrs.py:99-117
    Bucket Bs, Bx;
    if (!load_bucket(qs, R, 0u, tab_s, s_cur, Bs) || !load_bucket(qx, R, 0u, tab_x, x_cur, Bx)) { status = OLS_NO_QUEUE; break; }  // KeyError
    float4 win = make_float4(0.f, 0.f, 0.f, 0.f);
    int win_xn = 0;
    bool took = false;
    while (!took) {
      if (Bs.cur >= Bs.len || Bx.cur >= Bx.len) { status = OLS_QUEUE_EMPTY; break; }  // psrs.py:104-107
      const float4* rp = (const float4*)((const char*)qs.records + (size_t)(uint32_t)(Bs.lo + Bs.cur) * kRec);
      float plf[NA];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 pv = __ldg(rp + v);
        plf[4 * v + 0] = pv.x; plf[4 * v + 1] = pv.y; plf[4 * v + 2] = pv.z; plf[4 * v + 3] = pv.w;
      }
      const float4 tail = __ldg(rp + NV);
      const int xn = __float_as_int(__ldg((const float4*)((const char*)qx.records + (size_t)(uint32_t)(Bx.lo + Bx.cur) * kRec) + NV).y);
      double u;
      if (RNG == RNGT_PCG64) {
        pcg_s = ols_pcg_step(pcg_s, pcg_inc);
        u = ols_u64_to_unit(ols_pcg_output(pcg_s));
      } else {
        u = ols_stream_uniform(PK, rid, 0u, draws);
      }
      Bs.cur += 1;
      Bx.cur += 1;
      draws += 1;
      pops += 1;
      const uint32_t meta = __float_as_uint(tail.w);
      const int a = ols_meta_action(meta);
      const float uf = (float)u;
      int dec = accept_filter<NA>(P, plf, a, uf, P.ok && uf >= 8.6736174e-19f && ols_meta_tame(meta) != 0);
      if (dec == 2) {
        const double* p64 = qs.p_log64 ? qs.p_log64 + (int64_t)__float_as_int(tail.z) * nA : nullptr;
        dec = (int)accept_slow(pn_fn, (const float*)rp, p64, nA, a, u, cfg.compute_mode);
      }
      if (dec == 1) {  // psrs.py:114 `u <= thr`
        took = true;
        win = tail;
        win_xn = xn;
      }
    }
    tab_s[Bs.b] = Bs.cur;
    tab_x[Bx.b] = Bx.cur;
    if (!took) break;
    const uint32_t meta = __float_as_uint(win.w);
    const double Rw = (double)win.x;
    if (cfg.trace_row && (long long)accepted < cfg.trace_cap) cfg.trace_row[r * cfg.trace_cap + (long long)accepted] = __float_as_int(win.z);
    s_cur = __float_as_int(win.y);
    x_cur = win_xn;
    accepted += 1;
    done_now += 1;
    G = __dadd_rn(G, __dmul_rn(disc, Rw));
    disc = __dmul_rn(disc, cfg.gamma);
    ep_t += 1;
    if (ols_meta_terminal(meta)) {
      if (ep_index < cfg.ep_cap) cfg.Gs[r * (int64_t)cfg.ep_cap + ep_index] = G;
      if (lc < cfg.ep_cap) cfg.lengths[r * (int64_t)cfg.ep_cap + lc] = ep_t;
      ep_index += 1;
      lc += 1;
      ep_open = 0;
    }
    if (cfg.max_steps > 0 && done_now >= cfg.max_steps) break;  // budget: stays RUNNING, resumable
  }
  if (status == OLS_QUEUE_EMPTY && ep_open) {  // psrs.py:265-269: the cut episode's length is appended, its return is not
    if (lc < cfg.ep_cap) cfg.lengths[r * (int64_t)cfg.ep_cap + lc] = ep_t;
    lc += 1;
    ep_open = 0;
  }
  R.status[r] = status;
  R.cur_z[r] = s_cur;
  R.cur_s[r] = s_cur;
  cur_x[r] = x_cur;
  R.init_cursor[r] = init_cursor;
  R.draws[r] = draws;
  R.pops[r] += pops;
  R.accepted[r] = accepted;
  R.ep_index[r] = ep_index;
  R.ep_len_count[r] = lc;
  R.ep_t[r] = ep_t;
  R.ep_open[r] = ep_open;
  R.ep_G[r] = G;
  R.ep_disc[r] = disc;
  if (RNG == RNGT_PCG64) {
    uint64_t* p = R.pcg_state + 4 * r;
    p[0] = pcg_s.hi; p[1] = pcg_s.lo; p[2] = pcg_inc.hi; p[3] = pcg_inc.lo;
  }
}

template <int NV>
cudaError_t launch_exo_run(const ols_queues_t& qs, const ols_queues_t& qx, const ols_replicas_t& r, int32_t* cursor_x, int32_t* cur_x,
                           const ols_exo_run_t& c, cudaStream_t stream) {
  const int grid = ols_div_up(r.n_replicas, 128);
  const ols_philox_keys K = ols_philox_make_keys(r.seed);
  if (r.rng_mode == OLS_RNG_PCG64) exo_run_kernel<NV, RNGT_PCG64><<<grid, 128, 0, stream>>>(qs, qx, r, cursor_x, cur_x, c, K);
  else exo_run_kernel<NV, RNGT_PHILOX><<<grid, 128, 0, stream>>>(qs, qx, r, cursor_x, cur_x, c, K);
  return cudaGetLastError();
}

template <int NV>
cudaError_t launch_exo(const ols_queues_t& qs, const ols_queues_t& qx, const ols_replicas_t& r, int32_t* cursor_x, int32_t* cur_x,
                       const uint8_t* mask, const void* p_new, int p64, int64_t p_stride, int compute_mode, int lanes,
                       const ols_step_out_t& out, int32_t* out_x_next, int grid, cudaStream_t stream) {
  const ols_philox_keys K = ols_philox_make_keys(r.seed);
#define OLS_GO(RNG) \
  exo_step_kernel<NV, RNG><<<grid, kThreads, 0, stream>>>(qs, qx, r, cursor_x, cur_x, mask, p_new, p64, p_stride, compute_mode, lanes, out, out_x_next, K)
  if (r.rng_mode == OLS_RNG_PCG64) OLS_GO(RNGT_PCG64);
  else if (r.rng_mode == OLS_RNG_FEED) OLS_GO(RNGT_FEED);
  else OLS_GO(RNGT_PHILOX);
#undef OLS_GO
  return cudaGetLastError();
}

}  // namespace

extern "C" int ols_step_exo(const ols_queues_t* qs, const ols_queues_t* qx, const ols_replicas_t* r, int32_t* cursor_x,
                            int32_t* cur_x, const uint8_t* mask, const void* p_new, int32_t p_new_is_f64, int64_t p_stride,
                            int32_t compute_mode, int32_t lanes, const ols_step_out_t* out, int32_t* out_x_next, void* stream) {
  OLS_REQUIRE(qs && qx && r && out, OLS_ERR_INVALID, "ols_step_exo: null argument");
  OLS_REQUIRE(qs->n_actions >= 1 && qs->n_actions <= OLS_MAX_ACTIONS && qx->n_actions == qs->n_actions, OLS_ERR_INVALID,
              "ols_step_exo: both queue families must share n_actions in [1, %d]", OLS_MAX_ACTIONS);
  OLS_REQUIRE(qs->n_rows == qx->n_rows, OLS_ERR_INVALID, "ols_step_exo: the two bucketings must cover the same log");
  OLS_REQUIRE(lanes == 1 || lanes == 2 || lanes == 4 || lanes == 8 || lanes == 16 || lanes == 32, OLS_ERR_INVALID,
              "ols_step_exo: lanes must be 1, 2, 4, 8, 16 or 32");
  OLS_REQUIRE(compute_mode == OLS_COMPUTE_F64 || compute_mode == OLS_COMPUTE_F32, OLS_ERR_INVALID, "ols_step_exo: bad compute_mode");
  if (r->n_replicas == 0) return OLS_OK;
  OLS_REQUIRE(qs->records && qx->records && qs->offsets && qx->offsets, OLS_ERR_INVALID, "ols_step_exo: queues not bucketed");
  OLS_REQUIRE(r->cursor && cursor_x && r->cur_z && cur_x && r->status && r->draws && r->pops && r->accepted && p_new &&
                  out->row && out->action && out->reward && out->terminal && out->s_next && out->z_from && out_x_next,
              OLS_ERR_INVALID, "ols_step_exo: null buffer");
  OLS_REQUIRE(r->rng_mode != OLS_RNG_PCG64 || r->pcg_state, OLS_ERR_INVALID, "ols_step_exo: PCG64 mode needs pcg_state");
  OLS_REQUIRE(p_stride == 0 || p_stride >= qs->n_actions, OLS_ERR_INVALID, "ols_step_exo: p_stride must be 0 or >= n_actions");
  const int grid = (int)((r->n_replicas * lanes + kThreads - 1) / kThreads);
  cudaError_t e;
  switch ((qs->n_actions + 3) / 4) {
    case 1: e = launch_exo<1>(*qs, *qx, *r, cursor_x, cur_x, mask, p_new, p_new_is_f64, p_stride, compute_mode, lanes, *out, out_x_next, grid, (cudaStream_t)stream); break;
    case 2: e = launch_exo<2>(*qs, *qx, *r, cursor_x, cur_x, mask, p_new, p_new_is_f64, p_stride, compute_mode, lanes, *out, out_x_next, grid, (cudaStream_t)stream); break;
    case 3: e = launch_exo<3>(*qs, *qx, *r, cursor_x, cur_x, mask, p_new, p_new_is_f64, p_stride, compute_mode, lanes, *out, out_x_next, grid, (cudaStream_t)stream); break;
    default: e = launch_exo<4>(*qs, *qx, *r, cursor_x, cur_x, mask, p_new, p_new_is_f64, p_stride, compute_mode, lanes, *out, out_x_next, grid, (cudaStream_t)stream); break;
  }
  OLS_REQUIRE(e == cudaSuccess, OLS_ERR_CUDA, "ols_step_exo: launch failed: %s", cudaGetErrorString(e));
  return OLS_OK;
}

extern "C" int ols_reseed(const ols_replicas_t* r, const uint64_t* seeds, const uint8_t* mask, void* stream) {
  OLS_REQUIRE(r != nullptr, OLS_ERR_INVALID, "ols_reseed: null argument");
  if (r->n_replicas == 0) return OLS_OK;
  OLS_REQUIRE(r->rng_mode == OLS_RNG_PCG64 && r->pcg_state && r->draws, OLS_ERR_INVALID, "ols_reseed: needs PCG64 mode and its state");
  int blocks = ols_div_up(r->n_replicas, 256);
  if (blocks > OLS_SM_COUNT * 8) blocks = OLS_SM_COUNT * 8;
  reseed_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*r, seeds, mask);
  OLS_LAUNCH_CHECK();
  return OLS_OK;
}

extern "C" int ols_run_exo(const ols_queues_t* qs, const ols_queues_t* qx, const ols_replicas_t* r, int32_t* cursor_x, int32_t* cur_x,
                           const ols_exo_run_t* c, void* stream) {
  OLS_REQUIRE(qs && qx && r && c, OLS_ERR_INVALID, "ols_run_exo: null argument");
  OLS_REQUIRE(qs->n_actions >= 1 && qs->n_actions <= OLS_MAX_ACTIONS && qx->n_actions == qs->n_actions, OLS_ERR_INVALID,
              "ols_run_exo: both queue families must share n_actions in [1, %d]", OLS_MAX_ACTIONS);
  OLS_REQUIRE(qs->n_rows == qx->n_rows && qs->n_init == qx->n_init, OLS_ERR_INVALID, "ols_run_exo: the two bucketings must cover the same log");
  OLS_REQUIRE(qs->z_offset == 0 && qx->z_offset == 0, OLS_ERR_UNSUPPORTED, "ols_run_exo: state ids must be non-negative");
  OLS_REQUIRE(c->n_episodes >= 0 && c->ep_cap >= c->n_episodes + 1 && c->n_obs >= 1 && c->x_mult >= 1, OLS_ERR_INVALID, "ols_run_exo: bad sizes");
  OLS_REQUIRE(c->compute_mode == OLS_COMPUTE_F64 || c->compute_mode == OLS_COMPUTE_F32, OLS_ERR_INVALID, "ols_run_exo: bad compute_mode");
  OLS_REQUIRE(r->rng_mode == OLS_RNG_PCG64 || r->rng_mode == OLS_RNG_PHILOX, OLS_ERR_UNSUPPORTED, "ols_run_exo: PCG64 or Philox uniforms");
  OLS_REQUIRE(!c->trace_row || c->trace_cap > 0, OLS_ERR_INVALID, "ols_run_exo: the row trace needs trace_cap > 0");
  if (r->n_replicas == 0) return OLS_OK;
  OLS_REQUIRE(qs->records && qx->records && qs->offsets && qx->offsets && (qs->n_init == 0 || (qs->init_z && qx->init_z)), OLS_ERR_INVALID,
              "ols_run_exo: queues not bucketed");
  OLS_REQUIRE(r->cursor && cursor_x && r->cur_z && r->cur_s && cur_x && r->status && r->draws && r->pops && r->accepted && r->init_cursor &&
                  r->ep_index && r->ep_len_count && r->ep_t && r->ep_open && r->ep_G && r->ep_disc && c->pi && c->Gs && c->lengths,
              OLS_ERR_INVALID, "ols_run_exo: null buffer");
  OLS_REQUIRE(r->rng_mode != OLS_RNG_PCG64 || r->pcg_state, OLS_ERR_INVALID, "ols_run_exo: PCG64 mode needs pcg_state");
  cudaError_t e;
  switch ((qs->n_actions + 3) / 4) {
    case 1: e = launch_exo_run<1>(*qs, *qx, *r, cursor_x, cur_x, *c, (cudaStream_t)stream); break;
    case 2: e = launch_exo_run<2>(*qs, *qx, *r, cursor_x, cur_x, *c, (cudaStream_t)stream); break;
    case 3: e = launch_exo_run<3>(*qs, *qx, *r, cursor_x, cur_x, *c, (cudaStream_t)stream); break;
    default: e = launch_exo_run<4>(*qs, *qx, *r, cursor_x, cur_x, *c, (cudaStream_t)stream); break;
  }
  OLS_REQUIRE(e == cudaSuccess, OLS_ERR_CUDA, "ols_run_exo: launch failed: %s", cudaGetErrorString(e));
  ols_note_launch(1);
  return OLS_OK;
}
