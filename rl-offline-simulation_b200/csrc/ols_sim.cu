// The batched PSRS simulator: reset / step / fused learner runs over many independent replicas.
//
// Replaces, per replica, offsim4rl/evaluators/psrs.py:
//   :32-37   PSRS.reset                      -> reset_one()
//   :39-51   PSRS.step (pop until accepted)  -> pop_until_accept()
//   :53-57   PSRS._default_reject            -> accept_rule()
//   :119-185 qlearn_psrs    (+ agents/tabular.py:4-5,24-32 epsilon_greedy_policy / _random_argmax)
//   :187-239 expSARSA_psrs
//   :241-271 evalMC_psrs                     -> run_kernel<NV, AGENT, QT>
//
// Execution model.  A replica is served by a lane group of G = 1..32 lanes of one warp.  The group
// keeps the replica's registers (state, counters, episode accumulators) redundantly in every lane.
// One pop round: lane k loads the 16B-aligned record at queue position cursor+k (two or more 16-byte
// vector loads; consecutive lanes read consecutive records, i.e. one coalesced segment), evaluates
// the accept rule with uniform number draws+k, and a warp ballot finds the first accepting lane.
// Because uniforms are indexed by draw count, this is exactly the reference's sequential loop.
// Cursor and draw counter advance by (first accepting lane + 1), or by the whole round.
#include "ols_common.cuh"

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kThreads = 256;

struct Lane {
  int lane;    // lane in warp
  int G;       // lanes per group
  int sub;     // lane in group
  int gbase;   // first lane of the group
  unsigned gmask;  // (1 << G) - 1
};

__device__ __forceinline__ Lane make_lane(int G) {
  Lane L;
  L.lane = threadIdx.x & 31;
  L.G = G;
  L.sub = L.lane & (G - 1);
  L.gbase = L.lane - L.sub;
  L.gmask = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
  return L;
}

// ------------------------------------------------------------------------------------------ accept rule
// psrs.py:53-57:  M = (p_new / p_log).max();  u = rng.random();  reject = u > p_new[a] / p_log[a] / M
// NumPy semantics kept: IEEE division (no fast-math), left-to-right, NaN-propagating max, accept on
// !(u > thr) so a NaN threshold accepts and an infinite M gives thr = 0.
template <int NA>
__device__ __forceinline__ bool accept_rule_f64(const double (&pn)[NA], const double (&pl)[NA], int nA, int a, double u) {
  double m = __ddiv_rn(pn[0], pl[0]);
  double pna = pn[0], pla = pl[0];
#pragma unroll
  for (int j = 1; j < NA; ++j) {
    if (j < nA) {
      double qj = __ddiv_rn(pn[j], pl[j]);
      if (!(m != m) && (qj != qj || qj > m)) m = qj;
      if (j == a) {
        pna = pn[j];
        pla = pl[j];
      }
    }
  }
  double thr = __ddiv_rn(__ddiv_rn(pna, pla), m);
  return !(u > thr);
}

// Both distributions float32: NumPy keeps the ratio, the max and the threshold in float32 and (NumPy >= 2)
// compares the python-float u as a float32.
template <int NA>
__device__ __forceinline__ bool accept_rule_f32(const double (&pn)[NA], const double (&pl)[NA], int nA, int a, double u) {
  float m = __fdiv_rn((float)pn[0], (float)pl[0]);
  float pna = (float)pn[0], pla = (float)pl[0];
#pragma unroll
  for (int j = 1; j < NA; ++j) {
    if (j < nA) {
      float qj = __fdiv_rn((float)pn[j], (float)pl[j]);
      if (!(m != m) && (qj != qj || qj > m)) m = qj;
      if (j == a) {
        pna = (float)pn[j];
        pla = (float)pl[j];
      }
    }
  }
  float thr = __fdiv_rn(__fdiv_rn(pna, pla), m);
  return !(__double2float_rn(u) > thr);
}

// ------------------------------------------------------------------------------------------ per-replica registers
struct Rep {
  int status;
  int cur_z, cur_s;
  int init_cursor;
  uint64_t draws, ties, pops, accepted;
  ols_u128 pcg_s, pcg_inc;
};

__device__ __forceinline__ void load_rep(const ols_replicas_t& R, int64_t r, Rep& s) {
  s.status = R.status[r];
  s.cur_z = R.cur_z[r];
  s.cur_s = R.cur_s[r];
  s.init_cursor = R.init_cursor[r];
  s.draws = R.draws[r];
  s.ties = R.ties[r];
  s.pops = R.pops[r];
  s.accepted = R.accepted[r];
  if (R.rng_mode == OLS_RNG_PCG64) {
    const uint64_t* p = R.pcg_state + 4 * r;
    s.pcg_s.hi = p[0];
    s.pcg_s.lo = p[1];
    s.pcg_inc.hi = p[2];
    s.pcg_inc.lo = p[3];
  }
}

__device__ __forceinline__ void store_rep(const ols_replicas_t& R, int64_t r, const Rep& s) {
  R.status[r] = s.status;
  R.cur_z[r] = s.cur_z;
  R.cur_s[r] = s.cur_s;
  R.init_cursor[r] = s.init_cursor;
  R.draws[r] = s.draws;
  R.ties[r] = s.ties;
  R.pops[r] = s.pops;
  R.accepted[r] = s.accepted;
  if (R.rng_mode == OLS_RNG_PCG64) {
    uint64_t* p = R.pcg_state + 4 * r;
    p[0] = s.pcg_s.hi;
    p[1] = s.pcg_s.lo;
  }
}

// psrs.py:32-37 — pop the next initial (z, s); the ring start (if enabled) rotates the shared list.
__device__ __forceinline__ bool reset_one(const ols_queues_t& q, const ols_replicas_t& R, int64_t r, Rep& s) {
  if ((int64_t)s.init_cursor >= q.n_init) {
    s.status = OLS_INIT_EMPTY;
    return false;
  }
  int64_t pos = s.init_cursor;
  if (R.ring_mode) {
    uint64_t h = ols_mix64(R.ring_seed ^ ols_mix64((uint64_t)(R.replica_id0 + r) + 0x51ED270B1ull));
    pos += (int64_t)(((h >> 32) * (uint64_t)q.n_init) >> 32);
    if (pos >= q.n_init) pos -= q.n_init;
  }
  s.cur_z = __ldg(q.init_z + pos);
  s.cur_s = __ldg(q.init_s + pos);
  s.init_cursor += 1;
  s.status = OLS_RUNNING;
  return true;
}

// ------------------------------------------------------------------------------------------ the pop loop
struct PopResult {
  bool accepted;
  int status;     // OLS_RUNNING when accepted, else why not
  float reward;
  int z_next, s_next;
  uint32_t meta;
  int32_t slot;   // CSR slot of the accepted record
};

// Warp-convergent: every lane of the warp calls this; `active` is group-uniform.
template <int NV>
__device__ __forceinline__ void pop_until_accept(const ols_queues_t& q, const ols_replicas_t& R, const Lane& L, bool active,
                                                 int64_t r, const double (&pn)[4 * NV], int reject_mode, int compute_mode,
                                                 Rep& s, PopResult& res) {
  constexpr int NA = 4 * NV;
  const int nA = q.n_actions;
  res.accepted = false;
  res.status = OLS_RUNNING;
  res.reward = 0.f;
  res.z_next = res.s_next = 0;
  res.meta = 0;
  res.slot = -1;

  int lo = 0, len = 0, cur = 0;
  uint32_t start = 0;
  int32_t* cur_ptr = nullptr;
  if (active) {
    int b = s.cur_z + q.z_offset;
    if (b < 0 || b >= q.n_buckets) {
      res.status = OLS_NO_QUEUE;
      active = false;
    } else {
      lo = __ldg(q.offsets + b);
      len = __ldg(q.offsets + b + 1) - lo;
      if (len == 0) {
        res.status = OLS_NO_QUEUE;  // dict KeyError in the reference: z never logged as a from-state
        active = false;
      } else {
        cur_ptr = R.cursor + r * (int64_t)q.n_buckets + b;
        cur = *cur_ptr;
        if (R.ring_mode) start = ols_ring_start(R.ring_seed, (uint64_t)(R.replica_id0 + r), (uint32_t)b, (uint32_t)len);
      }
    }
  }
  const bool wrote = active;

  while (__any_sync(kFull, active)) {
    int remaining = len - cur;
    if (active && remaining <= 0) {
      res.status = OLS_QUEUE_EMPTY;  // psrs.py:44-45
      active = false;
    }
    const int n_eval = active ? min(L.G, remaining) : 0;
    bool acc = false, lacking = false;
    float rec_r = 0.f;
    int rec_zn = 0, rec_sn = 0;
    uint32_t rec_meta = 0;
    int32_t slot = -1;
    ols_u128 stepped = s.pcg_s;
    if (L.sub < n_eval) {
      uint32_t pos = start + (uint32_t)cur + (uint32_t)L.sub;
      if (pos >= (uint32_t)len) pos -= (uint32_t)len;
      slot = lo + (int32_t)pos;
      const float4* rp = (const float4*)((const char*)q.records + (size_t)slot * q.rec_bytes);
      float4 pv[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) pv[v] = __ldg(rp + v);
      float4 tail = __ldg(rp + NV);
      rec_r = tail.x;
      rec_zn = __float_as_int(tail.y);
      rec_sn = __float_as_int(tail.z);
      rec_meta = __float_as_uint(tail.w);
      if (reject_mode == OLS_REJECT_ACCEPT_ALL) {
        acc = true;  // trivial_baselines.py:8-10: no uniform is drawn
      } else {
        const uint64_t d = s.draws + (uint64_t)L.sub;
        double u = 0.0;
        if (R.rng_mode == OLS_RNG_FEED) {
          if ((int64_t)d < R.u_feed_len) u = __ldg(R.u_feed + r * R.u_stride + (int64_t)d);
          else lacking = true;
        } else if (R.rng_mode == OLS_RNG_PCG64) {
          for (int k = 0; k <= L.sub; ++k) stepped = ols_pcg_step(stepped, s.pcg_inc);
          u = ols_u64_to_unit(ols_pcg_output(stepped));
        } else {
          u = ols_philox_uniform(R.seed, (uint64_t)(R.replica_id0 + r), 0u, d);
        }
        if (!lacking) {
          double pl[NA];
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            pl[4 * v + 0] = (double)pv[v].x;
            pl[4 * v + 1] = (double)pv[v].y;
            pl[4 * v + 2] = (double)pv[v].z;
            pl[4 * v + 3] = (double)pv[v].w;
          }
          if (q.p_log64 != nullptr) {  // float64 log: the exact distribution lives in row order
            const double* p64 = q.p_log64 + (int64_t)__ldg(q.order + slot) * nA;
#pragma unroll
            for (int j = 0; j < NA; ++j)
              if (j < nA) pl[j] = __ldg(p64 + j);
          }
          const int a = ols_meta_action(rec_meta);
          acc = (compute_mode == OLS_COMPUTE_F32) ? accept_rule_f32<NA>(pn, pl, nA, a, u)
                                                  : accept_rule_f64<NA>(pn, pl, nA, a, u);
        }
      }
    }
    const unsigned b_acc = __ballot_sync(kFull, acc);
    const unsigned b_lack = __ballot_sync(kFull, lacking);
    const unsigned g_acc = (b_acc >> L.gbase) & L.gmask;
    const unsigned g_lack = (b_lack >> L.gbase) & L.gmask;
    const int first_acc = g_acc ? (__ffs(g_acc) - 1) : 32;
    const int first_lack = g_lack ? (__ffs(g_lack) - 1) : 32;
    int consumed = 0;
    bool took = false;
    if (active) {
      if (first_acc < first_lack) {
        consumed = first_acc + 1;
        took = true;
      } else if (first_lack < 32) {
        consumed = first_lack;
        res.status = OLS_FEED_EMPTY;
        active = false;
      } else {
        consumed = n_eval;
      }
    }
    // winner's record -> every lane of its group (shuffles are executed by the whole warp)
    const int src = L.gbase + (first_acc < 32 ? first_acc : 0);
    const float w_r = __shfl_sync(kFull, rec_r, src);
    const int w_zn = __shfl_sync(kFull, rec_zn, src);
    const int w_sn = __shfl_sync(kFull, rec_sn, src);
    const uint32_t w_meta = __shfl_sync(kFull, rec_meta, src);
    const int32_t w_slot = __shfl_sync(kFull, slot, src);
    if (R.rng_mode == OLS_RNG_PCG64) {
      const int adv_src = L.gbase + (consumed > 0 ? consumed - 1 : 0);
      const uint64_t nh = __shfl_sync(kFull, stepped.hi, adv_src);
      const uint64_t nl = __shfl_sync(kFull, stepped.lo, adv_src);
      if (consumed > 0 && reject_mode != OLS_REJECT_ACCEPT_ALL) {
        s.pcg_s.hi = nh;
        s.pcg_s.lo = nl;
      }
    }
    cur += consumed;
    s.pops += (uint64_t)consumed;
    if (reject_mode != OLS_REJECT_ACCEPT_ALL) s.draws += (uint64_t)consumed;
    if (took) {
      res.accepted = true;
      res.reward = w_r;
      res.z_next = w_zn;
      res.s_next = w_sn;
      res.meta = w_meta;
      res.slot = w_slot;
      active = false;
    }
  }
  // every lane of the group stores the same value: each lane later re-reads its own store
  if (wrote) *cur_ptr = cur;
}

// ------------------------------------------------------------------------------------------ kernels: init / reset / step
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
      ols_pcg64_seed(seeds ? seeds[r] : R.seed, &st, &inc);  // psrs.py:20
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
    Rep s;
    s.init_cursor = R.init_cursor[r];
    s.cur_z = R.cur_z[r];
    s.cur_s = R.cur_s[r];
    s.status = R.status[r];
    reset_one(q, R, r, s);
    R.init_cursor[r] = s.init_cursor;
    R.cur_z[r] = s.cur_z;
    R.cur_s[r] = s.cur_s;
    R.status[r] = s.status;
    R.ep_open[r] = 0;
  }
}

template <int NV>
__global__ void __launch_bounds__(kThreads) step_kernel(ols_queues_t q, ols_replicas_t R, const uint8_t* mask,
                                                         const void* p_new, int p_new_is_f64, int64_t p_stride,
                                                         int reject_mode, int compute_mode, int G, ols_step_out_t out) {
  constexpr int NA = 4 * NV;
  const Lane L = make_lane(G);
  const int gpw = 32 / G;
  const int64_t n_warp_blocks = (R.n_replicas + gpw - 1) / gpw;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t wb = warp0; wb < n_warp_blocks; wb += n_warps) {
    const int64_t r = wb * gpw + (L.lane / G);
    const bool exists = r < R.n_replicas;
    Rep s;
    bool active = false;
    double pn[NA];
#pragma unroll
    for (int j = 0; j < NA; ++j) pn[j] = 0.0;
    if (exists) {
      load_rep(R, r, s);
      const bool selected = !mask || mask[r];
      // a replica whose queue ran dry may be stepped again (it reports QUEUE_EMPTY again, psrs.py:44-45)
      active = selected && (s.status == OLS_RUNNING || s.status == OLS_QUEUE_EMPTY);
      if (active) {
#pragma unroll
        for (int j = 0; j < NA; ++j)
          if (j < q.n_actions)
            pn[j] = p_new_is_f64 ? __ldg((const double*)p_new + r * p_stride + j)
                                 : (double)__ldg((const float*)p_new + r * p_stride + j);
      }
    } else {
      s.status = OLS_IDLE; s.cur_z = s.cur_s = 0; s.init_cursor = 0;
      s.draws = s.ties = s.pops = s.accepted = 0;
      s.pcg_s = {0, 0}; s.pcg_inc = {0, 0};
    }
    const bool was_active = active;
    const int z_from = s.cur_z;
    PopResult res;
    pop_until_accept<NV>(q, R, L, active, r, pn, reject_mode, compute_mode, s, res);
    if (exists && L.sub == 0) {
      if (was_active) {
        if (res.accepted) {
          s.status = OLS_RUNNING;
          s.cur_z = res.z_next;   // psrs.py:49-50
          s.cur_s = res.s_next;
          s.accepted += 1;
          out.row[r] = __ldg(q.order + res.slot);
          out.action[r] = ols_meta_action(res.meta);
          out.reward[r] = res.reward;
          out.terminal[r] = (uint8_t)ols_meta_terminal(res.meta);
          out.s_next[r] = res.s_next;
          out.z_from[r] = z_from;
        } else {
          s.status = res.status;
          out.row[r] = -1;
        }
        store_rep(R, r, s);
      } else if (!mask || mask[r]) {
        out.row[r] = -1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ fused learner runs
template <typename QT>
struct QMath;
template <>
struct QMath<double> {
  // float64 parity path: the reference's expressions, one IEEE operation at a time, never contracted
  static __device__ __forceinline__ double td(double R, double gamma, double boot, double q) {
    return __dsub_rn(__dadd_rn(R, __dmul_rn(gamma, boot)), q);   // R + gamma * boot - Q[S,A]   (psrs.py:165,223)
  }
  static __device__ __forceinline__ double upd(double q, double alpha, double td) {
    return __dadd_rn(q, __dmul_rn(alpha, td));                   // Q[S,A] + alpha * (...)      (psrs.py:168,223)
  }
};
template <>
struct QMath<float> {
  static __device__ __forceinline__ float td(float R, float gamma, float boot, float q) { return R + gamma * boot - q; }
  static __device__ __forceinline__ float upd(float q, float alpha, float td) { return q + alpha * td; }
};

template <int NV, int AGENT, typename QT>
__global__ void __launch_bounds__(kThreads) run_kernel(ols_queues_t q, ols_replicas_t R, ols_run_t cfg) {
  constexpr int NA = 4 * NV;
  const int nA = q.n_actions;
  const int nS = cfg.n_states;
  const Lane L = make_lane(cfg.lanes);
  const int gpw = 32 / cfg.lanes;
  const int64_t n_warp_blocks = (R.n_replicas + gpw - 1) / gpw;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;

  for (int64_t wb = warp0; wb < n_warp_blocks; wb += n_warps) {
    const int64_t r = wb * gpw + (L.lane / cfg.lanes);
    const bool exists = r < R.n_replicas;
    Rep s;
    int ep_index = 0, ep_len_count = 0, ep_t = 0, ep_open = 0;
    double ep_G = 0.0, ep_disc = 1.0;
    int64_t steps = 0;
    bool live = false;
    if (exists) {
      load_rep(R, r, s);
      ep_index = R.ep_index[r];
      ep_len_count = R.ep_len_count[r];
      ep_t = R.ep_t[r];
      ep_open = R.ep_open[r];
      ep_G = R.ep_G[r];
      ep_disc = R.ep_disc[r];
      // a replica is live until it terminates; IDLE/RUNNING replicas (re)start here
      live = (s.status == OLS_IDLE || s.status == OLS_RUNNING);
    } else {
      s.status = OLS_IDLE; s.cur_z = s.cur_s = 0; s.init_cursor = 0;
      s.draws = s.ties = s.pops = s.accepted = 0;
      s.pcg_s = {0, 0}; s.pcg_inc = {0, 0};
    }
    QT* Qr = (AGENT == OLS_AGENT_FIXED_PI) ? nullptr : ((QT*)cfg.Q + r * (int64_t)nS * nA);
    double* Gs = cfg.Gs + r * (int64_t)cfg.ep_cap;
    int32_t* lens = cfg.lengths + r * (int64_t)cfg.ep_cap;

    while (__any_sync(kFull, live)) {
      // ---- episode start: psrs.py:149-156 / 205-212 / 246-253
      if (live && !ep_open) {
        if (ep_index >= cfg.n_episodes) {
          s.status = OLS_DONE;
          live = false;
        } else if (!reset_one(q, R, r, s)) {
          live = false;  // INIT_EMPTY: terminate without appending (psrs.py:153-155)
        } else {
          ep_open = 1;
          ep_t = 0;
          ep_G = 0.0;
          ep_disc = 1.0;
        }
      }
      // ---- the agent's action distribution for the current state
      double pn[NA];
#pragma unroll
      for (int j = 0; j < NA; ++j) pn[j] = 0.0;
      const int S = s.cur_s;
      if (live && (S < 0 || S >= nS)) {  // IndexError on Q[S] / pi[S] in the reference
        s.status = OLS_NO_QUEUE;
        live = false;
      }
      if (live) {
        if (AGENT == OLS_AGENT_QLEARN) {
          // tabular.py:24-32 epsilon_greedy_policy on the single row Q[[S], :] (psrs.py:158)
          const double eps = __ldg(cfg.epsilon + ep_index);
          const double base = __ddiv_rn(eps, (double)nA);                       // ones * eps / nA
          const double top = __dadd_rn(__dsub_rn(1.0, eps), base);              // 1 - eps + eps / nA
          QT row[NA];
          QT mx = Qr[(int64_t)S * nA];
          row[0] = mx;
#pragma unroll
          for (int j = 1; j < NA; ++j)
            if (j < nA) {
              row[j] = Qr[(int64_t)S * nA + j];
              if (row[j] > mx) mx = row[j];
            }
          int ncand = 0;
#pragma unroll
          for (int j = 0; j < NA; ++j)
            if (j < nA && row[j] == mx) ++ncand;
          int pick = 0;
          if (ncand > 1) {  // tabular.py:4-5 _random_argmax: a draw is consumed only on ties
            double v = 0.0;
            bool have = true;
            if (R.tie_mode == OLS_RNG_FEED) {
              if ((int64_t)s.ties < R.tie_feed_len) v = __ldg(R.tie_feed + r * R.tie_stride + (int64_t)s.ties);
              else have = false;
            } else {
              v = ols_philox_uniform(R.seed, (uint64_t)(R.replica_id0 + r), 1u, s.ties);
            }
            if (!have) {
              s.status = OLS_FEED_EMPTY;
              live = false;
            } else {
              s.ties += 1;
              pick = min(ncand - 1, (int)(v * (double)ncand));
            }
          }
          int seen = 0, a_star = 0;
#pragma unroll
          for (int j = 0; j < NA; ++j)
            if (j < nA && row[j] == mx) {
              if (seen == pick) a_star = j;
              ++seen;
            }
#pragma unroll
          for (int j = 0; j < NA; ++j)
            if (j < nA) pn[j] = (j == a_star) ? top : base;
        } else {
          // psrs.py:214,255 — p = pi[S]
#pragma unroll
          for (int j = 0; j < NA; ++j)
            if (j < nA)
              pn[j] = cfg.pi_is_f64 ? __ldg((const double*)cfg.pi + (int64_t)S * nA + j)
                                    : (double)__ldg((const float*)cfg.pi + (int64_t)S * nA + j);
        }
      }
      // ---- env.step(p): psrs.py:39-57
      PopResult res;
      pop_until_accept<NV>(q, R, L, live, r, pn, cfg.reject_mode, cfg.compute_mode, s, res);
      // ---- learner update and episode bookkeeping
      if (live) {
        if (res.accepted) {
          const int A = ols_meta_action(res.meta);
          const int S2 = res.s_next;
          const int done = ols_meta_terminal(res.meta);
          const double Rw = (double)res.reward;
          if (AGENT != OLS_AGENT_FIXED_PI) {
            if (S2 < 0 || S2 >= nS) {
              s.status = OLS_NO_QUEUE;
              live = false;
            } else {
              const double alpha = __ldg(cfg.alpha + ep_index);
              QT boot;
              if (AGENT == OLS_AGENT_QLEARN) {
                boot = Qr[(int64_t)S2 * nA];                                   // Q[S_].max()
#pragma unroll
                for (int j = 1; j < NA; ++j)
                  if (j < nA) {
                    QT v = Qr[(int64_t)S2 * nA + j];
                    if (v > boot) boot = v;
                  }
              } else {
                double acc = 0.0;                                               // Q[S_] @ pi[S_]
#pragma unroll
                for (int j = 0; j < NA; ++j)
                  if (j < nA) {
                    double pj = cfg.pi_is_f64 ? __ldg((const double*)cfg.pi + (int64_t)S2 * nA + j)
                                              : (double)__ldg((const float*)cfg.pi + (int64_t)S2 * nA + j);
                    acc = fma((double)Qr[(int64_t)S2 * nA + j], pj, acc);
                  }
                boot = (QT)acc;
              }
              const QT qsa = Qr[(int64_t)S * nA + A];
              const QT td = QMath<QT>::td((QT)Rw, (QT)cfg.gamma, boot, qsa);
              const QT qn = QMath<QT>::upd(qsa, (QT)alpha, td);
              Qr[(int64_t)S * nA + A] = qn;  // all lanes of the group store the same value (see cursor)
              if (L.sub == 0 && cfg.td_error && (int64_t)s.accepted < cfg.trace_cap)
                cfg.td_error[r * cfg.trace_cap + (int64_t)s.accepted] = (double)td;
            }
          }
          if (live) {
            if (cfg.trace_row && L.sub == 0 && (int64_t)s.accepted < cfg.trace_cap)
              cfg.trace_row[r * cfg.trace_cap + (int64_t)s.accepted] = __ldg(q.order + res.slot);
            s.cur_z = res.z_next;
            s.cur_s = S2;
            s.accepted += 1;
            ep_G = __dadd_rn(ep_G, __dmul_rn(ep_disc, Rw));   // G + gamma**t * R   (psrs.py:171)
            ep_disc = __dmul_rn(ep_disc, cfg.gamma);
            ep_t += 1;
            steps += 1;
            if (done) {  // psrs.py:176-178 / 265-269
              if (L.sub == 0) {
                if (ep_index < cfg.ep_cap) Gs[ep_index] = ep_G;
                if (ep_len_count < cfg.ep_cap) lens[ep_len_count] = ep_t;
              }
              ep_index += 1;
              ep_len_count += 1;
              ep_open = 0;
            }
            if (cfg.max_steps > 0 && steps >= cfg.max_steps) live = false;  // budget: stays resumable
          }
        } else {
          // exhaustion in mid-episode (psrs.py:160-162): qlearn/expSARSA still append the partial
          // return (:177-178), evalMC appends only the length (:265-269)
          s.status = res.status;
          live = false;
          if (res.status == OLS_QUEUE_EMPTY) {
            if (L.sub == 0) {
              if (ep_len_count < cfg.ep_cap) lens[ep_len_count] = ep_t;
              if (AGENT != OLS_AGENT_FIXED_PI && ep_index < cfg.ep_cap) Gs[ep_index] = ep_G;
            }
            ep_len_count += 1;
            if (AGENT != OLS_AGENT_FIXED_PI) ep_index += 1;
            ep_open = 0;
          }
        }
      }
    }
    if (exists && L.sub == 0) {
      store_rep(R, r, s);
      R.ep_index[r] = ep_index;
      R.ep_len_count[r] = ep_len_count;
      R.ep_t[r] = ep_t;
      R.ep_open[r] = ep_open;
      R.ep_G[r] = ep_G;
      R.ep_disc[r] = ep_disc;
    }
  }
}

int check_sim(const ols_queues_t* q, const ols_replicas_t* r, const char* who) {
  OLS_REQUIRE(q && r, OLS_ERR_INVALID, "%s: null argument", who);
  OLS_REQUIRE(q->n_actions >= 1 && q->n_actions <= OLS_MAX_ACTIONS, OLS_ERR_UNSUPPORTED, "%s: n_actions must be in [1, %d]",
              who, OLS_MAX_ACTIONS);
  OLS_REQUIRE(q->rec_bytes == ols_record_bytes(q->n_actions), OLS_ERR_INVALID, "%s: bad rec_bytes", who);
  OLS_REQUIRE(q->n_buckets >= 1 && q->offsets, OLS_ERR_INVALID, "%s: bad queues", who);
  OLS_REQUIRE(r->n_replicas >= 0, OLS_ERR_INVALID, "%s: negative n_replicas", who);
  if (r->n_replicas == 0) return OLS_OK;
  OLS_REQUIRE(r->cursor && r->init_cursor && r->cur_z && r->cur_s && r->status && r->draws && r->ties && r->pops &&
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

int sim_grid(int64_t n_replicas, int lanes) {
  int64_t warps = (n_replicas * lanes + 31) / 32;
  int64_t blocks = (warps * 32 + kThreads - 1) / kThreads;
  int64_t cap = (int64_t)OLS_SM_COUNT * 8;  // 8 x 256 threads = 2048 resident threads per SM
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace

extern "C" int ols_replicas_init(const ols_queues_t* q, const ols_replicas_t* r, const uint64_t* seeds, void* stream) {
  int rc = check_sim(q, r, "ols_replicas_init");
  if (rc) return rc;
  if (r->n_replicas == 0) return OLS_OK;
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

extern "C" int ols_step(const ols_queues_t* q, const ols_replicas_t* r, const uint8_t* mask, const void* p_new,
                        int32_t p_new_is_f64, int64_t p_stride, int32_t reject_mode, int32_t compute_mode, int32_t lanes,
                        const ols_step_out_t* out, void* stream) {
  int rc = check_sim(q, r, "ols_step");
  if (rc) return rc;
  OLS_REQUIRE(valid_lanes(lanes), OLS_ERR_INVALID, "ols_step: lanes must be 1, 2, 4, 8, 16 or 32");
  OLS_REQUIRE(reject_mode == OLS_REJECT_DEFAULT || reject_mode == OLS_REJECT_ACCEPT_ALL, OLS_ERR_INVALID, "ols_step: bad reject_mode");
  OLS_REQUIRE(compute_mode == OLS_COMPUTE_F64 || compute_mode == OLS_COMPUTE_F32, OLS_ERR_INVALID, "ols_step: bad compute_mode");
  if (r->n_replicas == 0) return OLS_OK;
  OLS_REQUIRE(p_new && out && out->row && out->action && out->reward && out->terminal && out->s_next && out->z_from,
              OLS_ERR_INVALID, "ols_step: null p_new / output buffer");
  OLS_REQUIRE(p_stride == 0 || p_stride >= q->n_actions, OLS_ERR_INVALID, "ols_step: p_stride must be 0 or >= n_actions");
  OLS_REQUIRE(q->order && q->records, OLS_ERR_INVALID, "ols_step: queues not bucketed");
  const int grid = sim_grid(r->n_replicas, lanes);
  const int nv = (q->n_actions + 3) / 4;
#define OLS_STEP(NV)                                                                                                  \
  step_kernel<NV><<<grid, kThreads, 0, (cudaStream_t)stream>>>(*q, *r, mask, p_new, p_new_is_f64, p_stride, reject_mode, \
                                                               compute_mode, lanes, *out)
  switch (nv) {
    case 1: OLS_STEP(1); break;
    case 2: OLS_STEP(2); break;
    case 3: OLS_STEP(3); break;
    default: OLS_STEP(4); break;
  }
#undef OLS_STEP
  OLS_LAUNCH_CHECK();
  return OLS_OK;
}

template <int NV>
static int launch_run(const ols_queues_t* q, const ols_replicas_t* r, const ols_run_t* c, int grid, cudaStream_t stream) {
#define OLS_RUN(AG, QT) run_kernel<NV, AG, QT><<<grid, kThreads, 0, stream>>>(*q, *r, *c)
  if (c->agent == OLS_AGENT_FIXED_PI) OLS_RUN(OLS_AGENT_FIXED_PI, float);
  else if (c->agent == OLS_AGENT_EXPSARSA) { if (c->q_is_f64) OLS_RUN(OLS_AGENT_EXPSARSA, double); else OLS_RUN(OLS_AGENT_EXPSARSA, float); }
  else { if (c->q_is_f64) OLS_RUN(OLS_AGENT_QLEARN, double); else OLS_RUN(OLS_AGENT_QLEARN, float); }
#undef OLS_RUN
  OLS_LAUNCH_CHECK();
  return OLS_OK;
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
  if (r->n_replicas == 0) return OLS_OK;
  OLS_REQUIRE(c->Gs && c->lengths, OLS_ERR_INVALID, "ols_run: null Gs / lengths");
  OLS_REQUIRE(q->order && q->records, OLS_ERR_INVALID, "ols_run: queues not bucketed");
  OLS_REQUIRE(q->n_init == 0 || (q->init_z && q->init_s), OLS_ERR_INVALID, "ols_run: null init arrays");
  if (c->agent != OLS_AGENT_FIXED_PI) OLS_REQUIRE(c->Q && c->alpha, OLS_ERR_INVALID, "ols_run: learner needs Q and alpha");
  if (c->agent == OLS_AGENT_QLEARN) OLS_REQUIRE(c->epsilon, OLS_ERR_INVALID, "ols_run: qlearn needs epsilon");
  if (c->agent != OLS_AGENT_QLEARN) OLS_REQUIRE(c->pi, OLS_ERR_INVALID, "ols_run: fixed-policy agents need pi");
  OLS_REQUIRE((!c->trace_row && !c->td_error) || c->trace_cap > 0, OLS_ERR_INVALID, "ols_run: trace taps need trace_cap > 0");
  const int grid = sim_grid(r->n_replicas, c->lanes);
  switch ((q->n_actions + 3) / 4) {
    case 1: return launch_run<1>(q, r, c, grid, (cudaStream_t)stream);
    case 2: return launch_run<2>(q, r, c, grid, (cudaStream_t)stream);
    case 3: return launch_run<3>(q, r, c, grid, (cudaStream_t)stream);
    default: return launch_run<4>(q, r, c, grid, (cudaStream_t)stream);
  }
}
