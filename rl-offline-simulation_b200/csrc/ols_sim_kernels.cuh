// Device code of the batched PSRS simulator (included by ols_sim_nv*.cu, one per record width).
//
// Replaces, per replica, offsim4rl/evaluators/psrs.py:
//   :32-37   PSRS.reset                      -> reset_one()
//   :39-51   PSRS.step (pop until accepted)  -> pop_until_accept()
//   :53-57   PSRS._default_reject            -> accept_exact_* (+ a rigorous float32 pre-filter)
//   :119-185 qlearn_psrs    (+ agents/tabular.py:4-5,24-32 epsilon_greedy_policy / _random_argmax)
//   :187-239 expSARSA_psrs
//   :241-271 evalMC_psrs                     -> run_kernel<NV, AGENT, QT, RNG, PPL>
//
// Execution model.  A replica is served by a lane group of G = 1..32 lanes of one warp; every lane of
// the group holds the replica's registers redundantly.  One pop ROUND examines G*PPL consecutive queue
// records: lane k loads PPL 16B-aligned records (16-byte vector loads; consecutive lanes read
// consecutive records, i.e. one coalesced segment), evaluates the accept rule with the uniform of
// that pop's draw index, and warp ballots find the first accepting pop in queue order.  Because the
// uniforms are indexed by draw count (counter-based Philox, a fed array, or a stepped PCG64), this
// is exactly the reference's sequential loop; cursor and draw counter advance by (first accepted
// pop + 1) or by the whole round.
//
// Accept rule.  The reference decides `u > p_new[a]/p_log[a]/M`, M = max(p_new/p_log), in float64
// (float32 when both inputs are float32).  Evaluating that needs nA+2 IEEE divisions per pop; instead
// each pop first gets a float32 ESTIMATE of the threshold (relative error < 2^-19 under the range
// guards below) and is decided by it whenever u is outside a 2^-16 relative band around the estimate;
// inside the band, or when any guard fails (zero/huge/tiny probabilities, NaN, inf), the exact
// rule runs.  The decision is therefore always the exact rule's decision.
#pragma once

#include "ols_common.cuh"

namespace ols_sim {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kThreads = 256;
#ifndef OLS_MINB
#define OLS_MINB 2
#endif

enum { RNGT_FEED = OLS_RNG_FEED, RNGT_PCG64 = OLS_RNG_PCG64, RNGT_PHILOX = OLS_RNG_PHILOX };

struct Lane {
  int lane, G, sub, gbase;
  unsigned gmask;
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

// ------------------------------------------------------------------------------------------ exact accept rule
// psrs.py:53-57 with NumPy semantics: IEEE division (never fast-math), left-to-right, NaN-propagating
// max, accept on !(u > thr) so a NaN threshold accepts and an infinite M gives thr = 0.
template <int NA>
__device__ __forceinline__ bool accept_exact_f64(const double (&pn)[NA], const double (&pl)[NA], int nA, int a, double u) {
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

// Both distributions float32: NumPy keeps ratio, max and threshold in float32 and (NumPy >= 2) compares
// the python-float u as a float32.
template <int NA>
__device__ __forceinline__ bool accept_exact_f32(const double (&pn)[NA], const double (&pl)[NA], int nA, int a, double u) {
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

// Float32 estimate of the threshold.  Returns 0 = reject, 1 = accept, 2 = undecided (run the exact rule).
// Guards: every logged probability in [2^-40, 2^40] and the policy's entries 0 or in [2^-40, 1] with a
// positive maximum (pn_ok, checked once per step) keep all intermediates normal, so each approximate
// division (2 ulp) and each float conversion (0.5 ulp) stays a relative perturbation; five of them
// compound to < 2^-19, well inside the 2^-16 band.
template <int NA>
__device__ __forceinline__ int accept_filter(const float (&pnf)[NA], const float (&plf)[NA], int nA, int a, float uf,
                                             bool pn_ok) {
  float mn = plf[0], mx = plf[0];
  float q = __fdividef(pnf[0], plf[0]);
  float m = q, qa = q;
#pragma unroll
  for (int j = 1; j < NA; ++j) {
    if (j < nA) {
      mn = fminf(mn, plf[j]);
      mx = fmaxf(mx, plf[j]);
      q = __fdividef(pnf[j], plf[j]);
      m = fmaxf(m, q);
      qa = (j == a) ? q : qa;
    }
  }
  const bool guards = pn_ok && (mn >= 9.094947e-13f) && (mx <= 1.0995116e12f);  // 2^-40, 2^40 (false for NaN)
  const float thr = __fdividef(qa, m);
  const float lo = thr * (1.0f - 1.52587890625e-5f), hi = thr * (1.0f + 1.52587890625e-5f);  // 2^-16 band
  int dec = 2;
  if (uf > hi) dec = 0;
  if (uf < lo) dec = 1;
  // u below 2^-60 (incl. 0) is not resolved in float32; a zero policy entry gives thr = 0 exactly: reject iff u > 0
  if (!(uf >= 8.6736174e-19f)) dec = 2;
  return guards ? dec : 2;
}

// ------------------------------------------------------------------------------------------ per-replica registers
struct Rep {
  int status;
  int cur_z, cur_s;
  int init_cursor;
  uint64_t draws, ties, pops, accepted;
  ols_u128 pcg_s, pcg_inc;  // PCG64 mode only
};

template <int RNG>
__device__ __forceinline__ void load_rep(const ols_replicas_t& R, int64_t r, Rep& s) {
  s.status = R.status[r];
  s.cur_z = R.cur_z[r];
  s.cur_s = R.cur_s[r];
  s.init_cursor = R.init_cursor[r];
  s.draws = R.draws[r];
  s.ties = R.ties[r];
  s.pops = R.pops[r];
  s.accepted = R.accepted[r];
  if (RNG == RNGT_PCG64) {
    const uint64_t* p = R.pcg_state + 4 * r;
    s.pcg_s.hi = p[0];
    s.pcg_s.lo = p[1];
    s.pcg_inc.hi = p[2];
    s.pcg_inc.lo = p[3];
  }
}

template <int RNG>
__device__ __forceinline__ void store_rep(const ols_replicas_t& R, int64_t r, const Rep& s) {
  R.status[r] = s.status;
  R.cur_z[r] = s.cur_z;
  R.cur_s[r] = s.cur_s;
  R.init_cursor[r] = s.init_cursor;
  R.draws[r] = s.draws;
  R.ties[r] = s.ties;
  R.pops[r] = s.pops;
  R.accepted[r] = s.accepted;
  if (RNG == RNGT_PCG64) {
    uint64_t* p = R.pcg_state + 4 * r;
    p[0] = s.pcg_s.hi;
    p[1] = s.pcg_s.lo;
  }
}

__device__ __forceinline__ void blank_rep(Rep& s) {
  s.status = OLS_IDLE;
  s.cur_z = s.cur_s = 0;
  s.init_cursor = 0;
  s.draws = s.ties = s.pops = s.accepted = 0;
  s.pcg_s.hi = s.pcg_s.lo = s.pcg_inc.hi = s.pcg_inc.lo = 0;
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
  int status;  // OLS_RUNNING when accepted, else why not
  float reward;
  int z_next, s_next;
  uint32_t meta;
  int32_t slot;  // CSR slot of the accepted record
};

// The policy as the pop loop needs it: the float32 image of pi(.|s) and the guard bit of the filter.
// The exact float64 entries are only needed when the filter is undecided; they are re-derived there by
// the caller-supplied functor (epsilon-greedy: from eps and a*, fixed policies: re-read from the table).
template <int NA>
struct Policy {
  float pnf[NA];
  bool ok;
};

template <int NA>
__device__ __forceinline__ void finish_policy(Policy<NA>& P, const double (&pn)[NA], int nA) {
  float mx = 0.f;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < NA; ++j) {
    P.pnf[j] = (j < nA) ? (float)pn[j] : 0.f;
    if (j < nA) {
      ok = ok && (P.pnf[j] == 0.f || (P.pnf[j] >= 9.094947e-13f && P.pnf[j] <= 1.0f));
      mx = fmaxf(mx, P.pnf[j]);
    }
  }
  P.ok = ok && (mx > 0.f);
}

// Exact decision for one pop (cold path, a real call so that its register needs do not weigh on the hot
// loop): re-derive pi in float64, fetch the float64 log row if any.  Everything is passed BY VALUE — a
// reference to a kernel parameter struct would force the whole struct into local memory.
template <int NA, typename PnFn>
__device__ __noinline__ bool accept_slow(PnFn pn_fn, const double* p_log64, const int32_t* order, int nA, float4 pl0, float4 pl1,
                                        float4 pl2, float4 pl3, int32_t slot, int a, double u, int compute_mode) {
  double pn[NA], pl[NA];
  pn_fn(pn);
  const float4 pv[4] = {pl0, pl1, pl2, pl3};
#pragma unroll
  for (int v = 0; v < NA / 4; ++v) {
    pl[4 * v + 0] = (double)pv[v].x;
    pl[4 * v + 1] = (double)pv[v].y;
    pl[4 * v + 2] = (double)pv[v].z;
    pl[4 * v + 3] = (double)pv[v].w;
  }
  if (p_log64 != nullptr) {  // float64 log: the exact distribution lives in row order
    const double* p64 = p_log64 + (int64_t)__ldg(order + slot) * nA;
#pragma unroll
    for (int j = 0; j < NA; ++j)
      if (j < nA) pl[j] = __ldg(p64 + j);
  }
  return (compute_mode == OLS_COMPUTE_F32) ? accept_exact_f32<NA>(pn, pl, nA, a, u) : accept_exact_f64<NA>(pn, pl, nA, a, u);
}

// Warp-convergent: every lane of the warp calls this; `active` is group-uniform.
// cur_tab: the replica's cursor table (global or staged in shared memory).
template <int NV, int RNG, int PPL, typename PnFn>
__device__ __forceinline__ void pop_until_accept(const ols_queues_t& q, const ols_replicas_t& R, const Lane& L, bool active,
                                                 int64_t r, const Policy<4 * NV>& P, const PnFn& pn_fn, int reject_mode,
                                                 int compute_mode, int32_t* cur_tab, Rep& s, PopResult& res) {
  constexpr int NA = 4 * NV;
  static_assert(RNG != RNGT_PCG64 || PPL == 1, "PCG64 replay is sequential per lane group: one pop per lane");
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
    const int b = s.cur_z + q.z_offset;
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
        cur_ptr = cur_tab + b;
        cur = *cur_ptr;
        if (R.ring_mode) start = ols_ring_start(R.ring_seed, (uint64_t)(R.replica_id0 + r), (uint32_t)b, (uint32_t)len);
      }
    }
  }
  const bool wrote = active;
  const bool accept_all = (reject_mode == OLS_REJECT_ACCEPT_ALL);

  while (__any_sync(kFull, active)) {
    const int remaining = len - cur;
    if (active && remaining <= 0) {
      res.status = OLS_QUEUE_EMPTY;  // psrs.py:44-45
      active = false;
    }
    // Philox yields two uniforms per block: with PPL == 2 the round is aligned to an even draw index and
    // the slot before `draws` (if draws is odd) is masked out, so each lane needs exactly one block.
    const int align = (RNG == RNGT_PHILOX && PPL == 2 && !accept_all) ? (int)(s.draws & 1) : 0;
    int n_eval = active ? min(L.G * PPL - align, remaining) : 0;
    if (accept_all) n_eval = min(n_eval, 1);

    bool acc[PPL], lack[PPL];
    float rec_r[PPL];
    int rec_zn[PPL], rec_sn[PPL];
    uint32_t rec_meta[PPL];
    int32_t rec_slot[PPL];
    float4 pv[PPL][NV];
    ols_u128 stepped = s.pcg_s;
    const int k0 = L.sub * PPL - align;  // queue offset (from the cursor) of this lane's first pop

    // ---- issue all record loads of the round first
#pragma unroll
    for (int i = 0; i < PPL; ++i) {
      const int k = k0 + i;
      acc[i] = lack[i] = false;
      rec_r[i] = 0.f;
      rec_zn[i] = rec_sn[i] = 0;
      rec_meta[i] = 0;
      rec_slot[i] = -1;
      if (k >= 0 && k < n_eval) {
        uint32_t pos = start + (uint32_t)cur + (uint32_t)k;
        if (pos >= (uint32_t)len) pos -= (uint32_t)len;
        rec_slot[i] = lo + (int32_t)pos;
        const float4* rp = (const float4*)((const char*)q.records + (size_t)rec_slot[i] * q.rec_bytes);
#pragma unroll
        for (int v = 0; v < NV; ++v) pv[i][v] = __ldg(rp + v);
        const float4 tail = __ldg(rp + NV);
        rec_r[i] = tail.x;
        rec_zn[i] = __float_as_int(tail.y);
        rec_sn[i] = __float_as_int(tail.z);
        rec_meta[i] = __float_as_uint(tail.w);
      }
    }
    // ---- uniforms of this lane's pops
    double u[PPL];
#pragma unroll
    for (int i = 0; i < PPL; ++i) u[i] = 0.0;
    if (!accept_all) {
      if (RNG == RNGT_PHILOX && PPL == 2) {
        const uint64_t d0 = (s.draws - (uint64_t)align) + 2ull * (uint64_t)L.sub;  // even
        if (k0 + 1 >= 0 && k0 < n_eval) {
          ols_u4 c;
          const uint64_t blk = d0 >> 1, sid = (uint64_t)(R.replica_id0 + r);
          c.x = (uint32_t)blk;
          c.y = (uint32_t)(blk >> 32);
          c.z = (uint32_t)sid;
          c.w = (uint32_t)(sid >> 32);
          const ols_u4 o = ols_philox4x32_10(c, (uint32_t)R.seed, (uint32_t)(R.seed >> 32));
          u[0] = ols_u64_to_unit(((uint64_t)o.y << 32) | o.x);
          if (PPL > 1) u[PPL - 1] = ols_u64_to_unit(((uint64_t)o.w << 32) | o.z);
        }
      } else {
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
          const int k = k0 + i;
          if (k >= 0 && k < n_eval) {
            const uint64_t d = s.draws + (uint64_t)k;
            if (RNG == RNGT_FEED) {
              if ((int64_t)d < R.u_feed_len) u[i] = __ldg(R.u_feed + r * R.u_stride + (int64_t)d);
              else lack[i] = true;
            } else if (RNG == RNGT_PCG64) {
              // PCG64 always runs with PPL == 1: pop k needs the group's generator stepped k + 1 times
              for (int t = 0; t <= k; ++t) stepped = ols_pcg_step(stepped, s.pcg_inc);
              u[i] = ols_u64_to_unit(ols_pcg_output(stepped));
            } else {
              u[i] = ols_philox_uniform(R.seed, (uint64_t)(R.replica_id0 + r), 0u, d);
            }
          }
        }
      }
    }
    // ---- accept rule per pop
#pragma unroll
    for (int i = 0; i < PPL; ++i) {
      const int k = k0 + i;
      if (k >= 0 && k < n_eval && !lack[i]) {
        if (accept_all) {
          acc[i] = true;  // trivial_baselines.py:8-10: no uniform is drawn
        } else {
          const int a = ols_meta_action(rec_meta[i]);
          float plf[NA];
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            plf[4 * v + 0] = pv[i][v].x;
            plf[4 * v + 1] = pv[i][v].y;
            plf[4 * v + 2] = pv[i][v].z;
            plf[4 * v + 3] = pv[i][v].w;
          }
          int dec = accept_filter<NA>(P.pnf, plf, nA, a, (float)u[i], P.ok && q.p_log64 == nullptr);
          if (dec == 2) {
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            dec = (int)accept_slow<NA>(pn_fn, q.p_log64, q.order, nA, pv[i][0], NV > 1 ? pv[i][NV > 1 ? 1 : 0] : z4,
                                       NV > 2 ? pv[i][NV > 2 ? 2 : 0] : z4, NV > 3 ? pv[i][NV > 3 ? 3 : 0] : z4, rec_slot[i], a,
                                       u[i], compute_mode);
          }
          acc[i] = (dec == 1);
        }
      }
    }
    // ---- first accepting pop of each group, in queue order
    int first_acc = 1 << 20, first_lack = 1 << 20;
#pragma unroll
    for (int i = 0; i < PPL; ++i) {
      const unsigned ba = (__ballot_sync(kFull, acc[i]) >> L.gbase) & L.gmask;
      if (ba) first_acc = min(first_acc, (__ffs(ba) - 1) * PPL + i - align);
      if (RNG == RNGT_FEED) {
        const unsigned bl = (__ballot_sync(kFull, lack[i]) >> L.gbase) & L.gmask;
        if (bl) first_lack = min(first_lack, (__ffs(bl) - 1) * PPL + i - align);
      }
    }
    int consumed = 0;
    bool took = false;
    if (active) {
      if (first_acc < first_lack) {
        consumed = first_acc + 1;
        took = true;
      } else if (first_lack < (1 << 20)) {
        consumed = first_lack;
        res.status = OLS_FEED_EMPTY;
        active = false;
      } else {
        consumed = n_eval;
      }
    }
    // winner's record -> every lane of its group (shuffles are executed by the whole warp)
    const int wk = (first_acc < (1 << 20)) ? first_acc + align : 0;
    const int src = L.gbase + wk / PPL;
    const int wi = wk % PPL;
    float sel_r = rec_r[0];
    int sel_zn = rec_zn[0], sel_sn = rec_sn[0], sel_slot = rec_slot[0];
    uint32_t sel_meta = rec_meta[0];
#pragma unroll
    for (int i = 1; i < PPL; ++i) {
      // each lane offers the candidate its group's winner index points at
      if (wi == i) {
        sel_r = rec_r[i];
        sel_zn = rec_zn[i];
        sel_sn = rec_sn[i];
        sel_meta = rec_meta[i];
        sel_slot = rec_slot[i];
      }
    }
    const float w_r = __shfl_sync(kFull, sel_r, src);
    const int w_zn = __shfl_sync(kFull, sel_zn, src);
    const int w_sn = __shfl_sync(kFull, sel_sn, src);
    const uint32_t w_meta = __shfl_sync(kFull, sel_meta, src);
    const int32_t w_slot = __shfl_sync(kFull, sel_slot, src);
    if (RNG == RNGT_PCG64) {
      // the generator advanced by `consumed` draws: that is the stepped state of the lane that evaluated
      // pop (consumed - 1); PCG64 runs with PPL == 1, so that lane is gbase + consumed - 1
      const int adv_src = L.gbase + (consumed > 0 ? consumed - 1 : 0);
      const uint64_t nh = __shfl_sync(kFull, stepped.hi, adv_src);
      const uint64_t nl = __shfl_sync(kFull, stepped.lo, adv_src);
      if (consumed > 0 && !accept_all) {
        s.pcg_s.hi = nh;
        s.pcg_s.lo = nl;
      }
    }
    cur += consumed;
    s.pops += (uint64_t)consumed;
    if (!accept_all) s.draws += (uint64_t)consumed;
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

// ------------------------------------------------------------------------------------------ kernels: step
template <int NV, int RNG>
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
    blank_rep(s);
    bool active = false;
    const int nA = q.n_actions;
    const int64_t prow = exists ? r * p_stride : 0;
    auto pn_fn = [=](double (&pn)[NA]) {
#pragma unroll
      for (int j = 0; j < NA; ++j)
        pn[j] = (j < nA) ? (p_new_is_f64 ? __ldg((const double*)p_new + prow + j) : (double)__ldg((const float*)p_new + prow + j))
                         : 0.0;
    };
    Policy<NA> P;
    if (exists) {
      load_rep<RNG>(R, r, s);
      const bool selected = !mask || mask[r];
      // a replica whose queue ran dry may be stepped again (it reports QUEUE_EMPTY again, psrs.py:44-45)
      active = selected && (s.status == OLS_RUNNING || s.status == OLS_QUEUE_EMPTY);
    }
    {
      double pn[NA];
      pn_fn(pn);
      finish_policy<NA>(P, pn, nA);
    }
    const bool was_active = active;
    const int z_from = s.cur_z;
    PopResult res;
    pop_until_accept<NV, RNG, 1>(q, R, L, active, r, P, pn_fn, reject_mode, compute_mode,
                                 R.cursor + (exists ? r : 0) * (int64_t)q.n_buckets, s, res);
    if (exists && L.sub == 0) {
      if (was_active) {
        if (res.accepted) {
          s.status = OLS_RUNNING;
          s.cur_z = res.z_next;  // psrs.py:49-50
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
        store_rep<RNG>(R, r, s);
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
    return __dsub_rn(__dadd_rn(R, __dmul_rn(gamma, boot)), q);  // R + gamma * boot - Q[S,A]   (psrs.py:165,223)
  }
  static __device__ __forceinline__ double upd(double q, double alpha, double td) {
    return __dadd_rn(q, __dmul_rn(alpha, td));  // Q[S,A] + alpha * (...)      (psrs.py:168,223)
  }
};
template <>
struct QMath<float> {
  static __device__ __forceinline__ float td(float R, float gamma, float boot, float q) { return R + gamma * boot - q; }
  static __device__ __forceinline__ float upd(float q, float alpha, float td) { return q + alpha * td; }
};

// Shared-memory staging plan of one CTA (host-computed, see ols_sim.cu): each lane group owns a slot of
// `slot_bytes` holding its replica's cursor table and, when it fits, its Q table.
struct StagePlan {
  int cursor_bytes;  // 0 = cursors stay in global memory
  int q_bytes;       // 0 = Q stays in global memory
  int slot_bytes;
};

template <int NV, int AGENT, typename QT, int RNG, int PPL>
__global__ void __launch_bounds__(kThreads, (NV == 1 ? OLS_MINB : 2)) run_kernel(ols_queues_t q, ols_replicas_t R, ols_run_t cfg,
                                                                           StagePlan plan) {
  constexpr int NA = 4 * NV;
  extern __shared__ __align__(16) unsigned char s_stage[];
  const int nA = q.n_actions;
  const int nS = cfg.n_states;
  const int nB = q.n_buckets;
  const Lane L = make_lane(cfg.lanes);
  // one lane group per replica, one launch-time warp per 32/G replicas (the grid covers all of them)
  const int64_t r = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) * (32 / cfg.lanes) + (L.lane / cfg.lanes);
  const bool exists = r < R.n_replicas;
  const int64_t rr = exists ? r : 0;
  unsigned char* my_slot = s_stage + (size_t)(threadIdx.x / cfg.lanes) * plan.slot_bytes;

  Rep s;
  blank_rep(s);
  int ep_index = 0, ep_len_count = 0, ep_t = 0, ep_open = 0;
  double ep_G = 0.0, ep_disc = 1.0;
  int steps = 0;
  bool live = false;
  if (exists) {
    load_rep<RNG>(R, r, s);
    ep_index = R.ep_index[r];
    ep_len_count = R.ep_len_count[r];
    ep_t = R.ep_t[r];
    ep_open = R.ep_open[r];
    ep_G = R.ep_G[r];
    ep_disc = R.ep_disc[r];
    live = (s.status == OLS_IDLE || s.status == OLS_RUNNING);  // terminated replicas stay terminated
  }
  const uint64_t accepted0 = s.accepted;  // lifetime counters: 32-bit deltas inside the launch
  uint32_t n_acc = 0, n_pop = 0;
  int32_t* cur_tab = R.cursor + rr * (int64_t)nB;
  QT* Qr = (AGENT == OLS_AGENT_FIXED_PI) ? nullptr : ((QT*)cfg.Q + rr * (int64_t)nS * nA);
  // ---- stage the replica's tables in shared memory (the group's lanes copy cooperatively)
  if (plan.cursor_bytes) {
    int32_t* sc = (int32_t*)my_slot;
    if (exists)
      for (int i = L.sub; i < nB; i += L.G) sc[i] = cur_tab[i];
    cur_tab = sc;
  }
  if (AGENT != OLS_AGENT_FIXED_PI && plan.q_bytes) {
    QT* sq = (QT*)(my_slot + plan.cursor_bytes);
    if (exists)
      for (int i = L.sub; i < nS * nA; i += L.G) sq[i] = Qr[i];
    Qr = sq;
  }
  if (plan.cursor_bytes || plan.q_bytes) __syncwarp();

  QT row[NA];  // Q[S, :] of the current state, carried over from the bootstrap read of the previous step
#pragma unroll
  for (int j = 0; j < NA; ++j) row[j] = (QT)0;
  bool row_valid = false;
  int a_star = 0;

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
        row_valid = false;
      }
    }
    // ---- the agent's action distribution for the current state
    const int S = s.cur_s;
    if (live && (S < 0 || S >= nS)) {  // IndexError on Q[S] / pi[S] in the reference
      s.status = OLS_NO_QUEUE;
      live = false;
    }
    if (AGENT == OLS_AGENT_QLEARN && live) {
      // tabular.py:24-32 epsilon_greedy_policy on the single row Q[[S], :] (psrs.py:158)
      if (!row_valid) {
#pragma unroll
        for (int j = 0; j < NA; ++j)
          if (j < nA) row[j] = Qr[(int64_t)S * nA + j];
      }
      QT mx = row[0];
#pragma unroll
      for (int j = 1; j < NA; ++j)
        if (j < nA && row[j] > mx) mx = row[j];
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
      int seen = 0;
      a_star = 0;
#pragma unroll
      for (int j = 0; j < NA; ++j)
        if (j < nA && row[j] == mx) {
          if (seen == pick) a_star = j;
          ++seen;
        }
    }
    // exact float64 pi(.|S): needed up front for its float image, and again on the (rare) slow path
    const int a_now = a_star;
    const double* eps_ptr = (AGENT == OLS_AGENT_QLEARN) ? cfg.epsilon + ep_index : nullptr;
    const void* pi_row = (AGENT == OLS_AGENT_QLEARN) ? nullptr
                         : (cfg.pi_is_f64 ? (const void*)((const double*)cfg.pi + (int64_t)(live ? S : 0) * nA)
                                          : (const void*)((const float*)cfg.pi + (int64_t)(live ? S : 0) * nA));
    const int pi_is_f64 = cfg.pi_is_f64;
    auto pn_fn = [=](double (&pn)[NA]) {
      if (AGENT == OLS_AGENT_QLEARN) {
        const double eps = __ldg(eps_ptr);
        const double base = __ddiv_rn(eps, (double)nA);           // ones * eps / nA          (tabular.py:27)
        const double top = __dadd_rn(__dsub_rn(1.0, eps), base);  // 1 - eps + eps / nA       (tabular.py:30)
#pragma unroll
        for (int j = 0; j < NA; ++j) pn[j] = (j < nA) ? ((j == a_now) ? top : base) : 0.0;
      } else {
        // psrs.py:214,255 — p = pi[S]
#pragma unroll
        for (int j = 0; j < NA; ++j)
          pn[j] = (j < nA) ? (pi_is_f64 ? __ldg((const double*)pi_row + j) : (double)__ldg((const float*)pi_row + j)) : 0.0;
      }
    };
    Policy<NA> P;
    {
      double pn[NA];
      pn_fn(pn);
      finish_policy<NA>(P, pn, nA);
    }
    // ---- env.step(p): psrs.py:39-57
    PopResult res;
    const uint64_t pops_before = s.pops;
    pop_until_accept<NV, RNG, PPL>(q, R, L, live, r, P, pn_fn, cfg.reject_mode, cfg.compute_mode, cur_tab, s, res);
    n_pop += (uint32_t)(s.pops - pops_before);
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
            QT nrow[NA];
#pragma unroll
            for (int j = 0; j < NA; ++j) nrow[j] = (j < nA) ? Qr[(int64_t)S2 * nA + j] : (QT)0;
            QT boot;
            if (AGENT == OLS_AGENT_QLEARN) {
              boot = nrow[0];  // Q[S_].max()
#pragma unroll
              for (int j = 1; j < NA; ++j)
                if (j < nA && nrow[j] > boot) boot = nrow[j];
            } else {
              double acc = 0.0;  // Q[S_] @ pi[S_]
#pragma unroll
              for (int j = 0; j < NA; ++j)
                if (j < nA) {
                  double pj = cfg.pi_is_f64 ? __ldg((const double*)cfg.pi + (int64_t)S2 * nA + j)
                                            : (double)__ldg((const float*)cfg.pi + (int64_t)S2 * nA + j);
                  acc = fma((double)nrow[j], pj, acc);
                }
              boot = (QT)acc;
            }
            QT qsa;
            if (AGENT == OLS_AGENT_QLEARN) {
              qsa = row[0];
#pragma unroll
              for (int j = 1; j < NA; ++j)
                if (j == A) qsa = row[j];
            } else {
              qsa = Qr[(int64_t)S * nA + A];
            }
            const QT td = QMath<QT>::td((QT)Rw, (QT)cfg.gamma, boot, qsa);
            const QT qn = QMath<QT>::upd(qsa, (QT)alpha, td);
            Qr[(int64_t)S * nA + A] = qn;  // all lanes of the group store the same value (see cursor)
            if (L.sub == 0 && cfg.td_error && (int64_t)(accepted0 + n_acc) < cfg.trace_cap)
              cfg.td_error[r * cfg.trace_cap + (int64_t)(accepted0 + n_acc)] = (double)td;
            if (AGENT == OLS_AGENT_QLEARN) {
              // the bootstrap row becomes the next step's policy row (patched if S' == S)
#pragma unroll
              for (int j = 0; j < NA; ++j) row[j] = (S2 == S && j == A) ? qn : nrow[j];
              row_valid = true;
            }
          }
        }
        if (live) {
          if (cfg.trace_row && L.sub == 0 && (int64_t)(accepted0 + n_acc) < cfg.trace_cap)
            cfg.trace_row[r * cfg.trace_cap + (int64_t)(accepted0 + n_acc)] = __ldg(q.order + res.slot);
          s.cur_z = res.z_next;
          s.cur_s = S2;
          n_acc += 1;
          ep_G = __dadd_rn(ep_G, __dmul_rn(ep_disc, Rw));  // G + gamma**t * R   (psrs.py:171)
          ep_disc = __dmul_rn(ep_disc, cfg.gamma);
          ep_t += 1;
          steps += 1;
          if (done) {  // psrs.py:176-178 / 265-269
            if (L.sub == 0) {
              if (ep_index < cfg.ep_cap) cfg.Gs[rr * (int64_t)cfg.ep_cap + ep_index] = ep_G;
              if (ep_len_count < cfg.ep_cap) cfg.lengths[rr * (int64_t)cfg.ep_cap + ep_len_count] = ep_t;
            }
            ep_index += 1;
            ep_len_count += 1;
            ep_open = 0;
          }
          if (cfg.max_steps > 0 && (int64_t)steps >= cfg.max_steps) live = false;  // budget: stays resumable
        }
      } else {
        // exhaustion in mid-episode (psrs.py:160-162): qlearn/expSARSA still append the partial
        // return (:177-178), evalMC appends only the length (:265-269)
        s.status = res.status;
        live = false;
        if (res.status == OLS_QUEUE_EMPTY) {
          if (L.sub == 0) {
            if (ep_len_count < cfg.ep_cap) cfg.lengths[rr * (int64_t)cfg.ep_cap + ep_len_count] = ep_t;
            if (AGENT != OLS_AGENT_FIXED_PI && ep_index < cfg.ep_cap) cfg.Gs[rr * (int64_t)cfg.ep_cap + ep_index] = ep_G;
          }
          ep_len_count += 1;
          if (AGENT != OLS_AGENT_FIXED_PI) ep_index += 1;
          ep_open = 0;
        }
      }
    }
  }
  // ---- write the staged tables back
  if (plan.cursor_bytes || plan.q_bytes) __syncwarp();
  if (plan.cursor_bytes && exists) {
    const int32_t* sc = (const int32_t*)my_slot;
    int32_t* g_cur = R.cursor + rr * (int64_t)nB;
    for (int i = L.sub; i < nB; i += L.G) g_cur[i] = sc[i];
  }
  if (AGENT != OLS_AGENT_FIXED_PI && plan.q_bytes && exists) {
    const QT* sq = (const QT*)(my_slot + plan.cursor_bytes);
    QT* g_Q = (QT*)cfg.Q + rr * (int64_t)nS * nA;
    for (int i = L.sub; i < nS * nA; i += L.G) g_Q[i] = sq[i];
  }
  if (exists && L.sub == 0) {
    s.accepted = accepted0 + n_acc;
    store_rep<RNG>(R, r, s);
    R.ep_index[r] = ep_index;
    R.ep_len_count[r] = ep_len_count;
    R.ep_t[r] = ep_t;
    R.ep_open[r] = ep_open;
    R.ep_G[r] = ep_G;
    R.ep_disc[r] = ep_disc;
  }
}

// ------------------------------------------------------------------------------------------ per-NV launchers
struct RunLaunch {
  int grid;
  size_t smem;
  StagePlan plan;
  cudaStream_t stream;
};

template <int NV, int AGENT, typename QT>
inline cudaError_t launch_run_rng(const ols_queues_t& q, const ols_replicas_t& r, const ols_run_t& c, const RunLaunch& L,
                                  int ppl) {
#define OLS_GO(RNG, PPL)                                                                                               \
  do {                                                                                                                 \
    auto k = run_kernel<NV, AGENT, QT, RNG, PPL>;                                                                      \
    if (L.smem > 48 * 1024) {                                                                                          \
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);               \
      if (e != cudaSuccess) return e;                                                                                  \
    }                                                                                                                  \
    k<<<L.grid, kThreads, L.smem, L.stream>>>(q, r, c, L.plan);                                                        \
  } while (0)
  if (r.rng_mode == OLS_RNG_PCG64) OLS_GO(RNGT_PCG64, 1);
  else if (r.rng_mode == OLS_RNG_FEED) { if (ppl == 2) OLS_GO(RNGT_FEED, 2); else OLS_GO(RNGT_FEED, 1); }
  else { if (ppl == 2) OLS_GO(RNGT_PHILOX, 2); else OLS_GO(RNGT_PHILOX, 1); }
#undef OLS_GO
  return cudaGetLastError();
}

template <int NV>
inline cudaError_t launch_run_nv(const ols_queues_t& q, const ols_replicas_t& r, const ols_run_t& c, const RunLaunch& L,
                                 int ppl) {
  if (c.agent == OLS_AGENT_FIXED_PI) return launch_run_rng<NV, OLS_AGENT_FIXED_PI, float>(q, r, c, L, ppl);
  if (c.agent == OLS_AGENT_EXPSARSA)
    return c.q_is_f64 ? launch_run_rng<NV, OLS_AGENT_EXPSARSA, double>(q, r, c, L, ppl)
                      : launch_run_rng<NV, OLS_AGENT_EXPSARSA, float>(q, r, c, L, ppl);
  return c.q_is_f64 ? launch_run_rng<NV, OLS_AGENT_QLEARN, double>(q, r, c, L, ppl)
                    : launch_run_rng<NV, OLS_AGENT_QLEARN, float>(q, r, c, L, ppl);
}

template <int NV>
inline cudaError_t launch_step_nv(const ols_queues_t& q, const ols_replicas_t& r, const uint8_t* mask, const void* p_new,
                                  int p_new_is_f64, int64_t p_stride, int reject_mode, int compute_mode, int lanes,
                                  const ols_step_out_t& out, int grid, cudaStream_t stream) {
#define OLS_GO(RNG) \
  step_kernel<NV, RNG><<<grid, kThreads, 0, stream>>>(q, r, mask, p_new, p_new_is_f64, p_stride, reject_mode, compute_mode, lanes, out)
  if (r.rng_mode == OLS_RNG_PCG64) OLS_GO(RNGT_PCG64);
  else if (r.rng_mode == OLS_RNG_FEED) OLS_GO(RNGT_FEED);
  else OLS_GO(RNGT_PHILOX);
#undef OLS_GO
  return cudaGetLastError();
}

}  // namespace ols_sim

// implemented in ols_sim_nv{1,2,3,4}.cu
cudaError_t ols_launch_run_nv1(const ols_queues_t&, const ols_replicas_t&, const ols_run_t&, const ols_sim::RunLaunch&, int ppl);
cudaError_t ols_launch_run_nv2(const ols_queues_t&, const ols_replicas_t&, const ols_run_t&, const ols_sim::RunLaunch&, int ppl);
cudaError_t ols_launch_run_nv3(const ols_queues_t&, const ols_replicas_t&, const ols_run_t&, const ols_sim::RunLaunch&, int ppl);
cudaError_t ols_launch_run_nv4(const ols_queues_t&, const ols_replicas_t&, const ols_run_t&, const ols_sim::RunLaunch&, int ppl);
#define OLS_STEP_ARGS                                                                                                      \
  const ols_queues_t &q, const ols_replicas_t &r, const uint8_t *mask, const void *p_new, int p_new_is_f64, int64_t p_stride, \
      int reject_mode, int compute_mode, int lanes, const ols_step_out_t &out, int grid, cudaStream_t stream
cudaError_t ols_launch_step_nv1(OLS_STEP_ARGS);
cudaError_t ols_launch_step_nv2(OLS_STEP_ARGS);
cudaError_t ols_launch_step_nv3(OLS_STEP_ARGS);
cudaError_t ols_launch_step_nv4(OLS_STEP_ARGS);
