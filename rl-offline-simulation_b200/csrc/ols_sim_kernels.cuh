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

#include <type_traits>

#include "ols_common.cuh"

namespace ols_sim {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kThreads = 256;
#ifndef OLS_MINB
#define OLS_MINB 4
#endif
// one lane per replica (GT == 1): no register is held redundantly, so the kernel can trade occupancy for registers in finer
// steps with small CTAs: 7 CTAs of 128 threads per SM = 896 replicas in flight at 72 registers (C4: 187.8 ms; 3 CTAs of 256
// at 80 registers: 196.9 ms; 4 of 256 at 64: 204-208 ms).  Every other kernel of the simulator runs kThreads per CTA.
#ifndef OLS_MINB_G1
#define OLS_MINB_G1 7
#endif
#ifndef OLS_G1_THREADS
#define OLS_G1_THREADS 128
#endif
// the one-thread-per-replica learner over TAIL hot records (ols_learner.cu) needs ~96 registers: 5 CTAs of 128 threads per SM
// without spills beat 7 with them (C4 kernel 156 ms against 162 ms; 6: 162 ms, 4: 182 ms — spilled registers travel through
// an L1 that the random record stream keeps flushing)
#ifndef OLS_LEARNER_MINB
#define OLS_LEARNER_MINB 5
#endif

// RNGT_NONE: the accept-all baselines (trivial_baselines.py:8-10,22-24) draw no uniform at all
enum { RNGT_FEED = OLS_RNG_FEED, RNGT_PCG64 = OLS_RNG_PCG64, RNGT_PHILOX = OLS_RNG_PHILOX, RNGT_NONE = 3 };

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
//
// This is the COLD path (the float32 filter below decides all but ~1e-5 of the pops).  It is a real call
// with few scalar arguments and rolled loops, so that its register needs stay small and do not weigh on
// the pop loop: pi(j|s) comes from the caller's functor, the logged distribution is re-read from memory
// (rec = the packed float32 record, p64 = the row of the float64 log, or NULL).
template <typename PnFn>
__device__ __noinline__ bool accept_slow(PnFn pn_fn, const float* rec, const double* p64, int nA, int a, double u,
                                        int compute_mode) {
  if (compute_mode == OLS_COMPUTE_F32) {
    // both distributions float32: NumPy keeps ratio, max and threshold in float32 and (NumPy >= 2) compares
    // the python-float u as a float32
    float m = __fdiv_rn((float)pn_fn(0), __ldg(rec));
#pragma unroll 1
    for (int j = 1; j < nA; ++j) {
      const float qj = __fdiv_rn((float)pn_fn(j), __ldg(rec + j));
      if (!(m != m) && (qj != qj || qj > m)) m = qj;
    }
    const float thr = __fdiv_rn(__fdiv_rn((float)pn_fn(a), __ldg(rec + a)), m);
    return !(__double2float_rn(u) > thr);
  }
  double m = __ddiv_rn(pn_fn(0), p64 ? __ldg(p64) : (double)__ldg(rec));
#pragma unroll 1
  for (int j = 1; j < nA; ++j) {
    const double qj = __ddiv_rn(pn_fn(j), p64 ? __ldg(p64 + j) : (double)__ldg(rec + j));
    if (!(m != m) && (qj != qj || qj > m)) m = qj;
  }
  const double thr = __ddiv_rn(__ddiv_rn(pn_fn(a), p64 ? __ldg(p64 + a) : (double)__ldg(rec + a)), m);
  return !(u > thr);
}

// Float32 estimate of the threshold.  Returns 0 = reject, 1 = accept, 2 = undecided (run the exact rule).
//
// With ipn[j] = 1 / pi(j|s) (+inf where pi is 0) and t_j = p_log[j] * ipn[j] = 1 / ratio_j:
//     1/M = min_j t_j,     thr = (1/M) / t_a          (= 0 when pi(a|s) = 0, since t_a = +inf)
// so a pop costs nA multiplies, nA-1 mins, one reciprocal and one multiply.  Guards: the record's logged
// probabilities lie in [2^-40, 2^40] (the "tame" bit, set when the record was packed) and pi's entries are
// 0 or in [2^-40, 1] with a positive maximum (P.ok, checked when the policy is formed).  Under them every
// intermediate is a normal float or +inf, each operation is a relative perturbation of at most 1 ulp, and
// the estimate is within 2^-20 of the exact threshold; the band is 2^-16.
template <int NA>
struct Policy {
  float ipn[NA];
  bool ok;
  // pop_round takes a policy SOURCE and materialises the reciprocals right before the accept rule, after the round's
  // loads and uniforms have been issued; a plain Policy is its own source
  __device__ __forceinline__ const Policy& get() const { return *this; }
};

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int NA>
__device__ __forceinline__ int accept_filter(const Policy<NA>& P, const float (&plf)[NA], int a, float uf, bool guards,
                                             float band = 1.52587890625e-5f /* 2^-16 */) {
  // padding entries (j >= n_actions) hold p = 1 in the record and ipn = +inf in the policy: t_j = +inf never wins the min
  float t = plf[0] * P.ipn[0];
  float inv_m = t, ta = t;
#pragma unroll
  for (int j = 1; j < NA; ++j) {
    t = plf[j] * P.ipn[j];
    inv_m = fminf(inv_m, t);
    ta = (j == a) ? t : ta;
  }
  const float thr = inv_m * rcp_approx(ta);
  int dec = 2;
  if (uf > thr * (1.0f + band)) dec = 0;
  if (uf < thr * (1.0f - band)) dec = 1;
  return guards ? dec : 2;
}

template <int NA, typename PnFn>
__device__ __forceinline__ void finish_policy(Policy<NA>& P, const PnFn& pn_fn, int nA) {
  float mx = 0.f;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < NA; ++j) {
    const float pj = (j < nA) ? (float)pn_fn(j) : 0.f;
    if (j < nA) {
      ok = ok && (pj == 0.f || (pj >= 9.094947e-13f && pj <= 1.0f));
      mx = fmaxf(mx, pj);
    }
    P.ipn[j] = (pj == 0.f) ? __int_as_float(0x7f800000) : __frcp_rn(pj);  // +inf for pi = 0 and for padding entries
  }
  P.ok = ok && (mx > 0.f);
}

// ------------------------------------------------------------------------------------------ replica state
// psrs.py:32-37 — pop the next initial (z, s); the ring start (if enabled) rotates the shared list.
// Returns false (INIT_EMPTY) when the list is exhausted.
__device__ __forceinline__ bool reset_one(const ols_queues_t& q, const ols_replicas_t& R, uint32_t h_rep, int& init_cursor,
                                          int& cur_z, int& cur_s) {
  if ((int64_t)init_cursor >= q.n_init) return false;
  int64_t pos = init_cursor;
  if (R.ring_mode) {
    pos += (int64_t)ols_ring_init_start(h_rep, (uint32_t)q.n_init);
    if (pos >= q.n_init) pos -= q.n_init;
  }
  cur_z = __ldg(q.init_z + pos);
  cur_s = __ldg(q.init_s + pos);
  init_cursor += 1;
  return true;
}

// The queue a replica is about to pop from: CSR range, ring start and the replica's cursor into it.
struct Bucket {
  int b;           // bucket index, -1 = none
  int lo, len;     // CSR slots [lo, lo + len)
  int cur;         // records consumed so far by this replica
  uint32_t start;  // ring offset (0 unless ring_mode)
};

// psrs.py:44 — `self.queues[z]`: a z that was never logged as a from-state has no queue (KeyError).
// A replica's learner state in array-of-structures form: one 32-byte entry per state,
//     { int32 cursor; int32 pad[3]; float q[4]; }
// so that the queue cursor and the Q row of a state share ONE DRAM sector.  A step reads the next state's entry with a
// single 256-bit load (cursor for the pops + Q row for the bootstrap and the next policy) and dirties only the current
// state's entry (cursor and Q[S, A]): two random sector accesses per transition instead of four.  Used by the fused
// float32 learners when the agent's observation id is the state id (the reference's shipped configs).
struct AosEntry {
  int cursor;
  float q[4];
};
// An entry is live only if its epoch word equals the run's epoch: a restart bumps the epoch instead of clearing the
// table (40 GB per GPU at the production shape), stale entries read as cursor 0 / Q 0, and every write stores the
// whole entry (one fully covered sector).
__device__ __forceinline__ AosEntry load_aos(const unsigned char* aos, int state, int epoch) {
  float4 a, b;
  asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(aos + (size_t)state * 32));
  AosEntry e;
  const bool fresh = (__float_as_int(a.y) == epoch);
  e.cursor = fresh ? __float_as_int(a.x) : 0;
  e.q[0] = fresh ? b.x : 0.f; e.q[1] = fresh ? b.y : 0.f; e.q[2] = fresh ? b.z : 0.f; e.q[3] = fresh ? b.w : 0.f;
  return e;
}
__device__ __forceinline__ void store_aos(unsigned char* aos, int state, int cursor, int epoch, float4 qrow) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(aos + (size_t)state * 32), "f"(__int_as_float(cursor)),
               "f"(__int_as_float(epoch)), "f"(0.f), "f"(0.f), "f"(qrow.x), "f"(qrow.y), "f"(qrow.z), "f"(qrow.w)
               : "memory");
}

__device__ __forceinline__ bool load_bucket(const ols_queues_t& q, const ols_replicas_t& R, uint32_t h_rep, const int32_t* cur_tab,
                                            int z, Bucket& B, bool have_cursor = false, int cursor_value = 0) {
  B.b = z + q.z_offset;
  B.lo = B.len = B.cur = 0;
  B.start = 0;
  if (B.b < 0 || B.b >= q.n_buckets) {
    B.b = -1;
    return false;
  }
  B.lo = __ldg(q.offsets + B.b);
  B.len = __ldg(q.offsets + B.b + 1) - B.lo;
  if (B.len == 0) {
    B.b = -1;
    return false;
  }
  B.cur = have_cursor ? cursor_value : cur_tab[B.b];
  if (R.ring_mode) B.start = ols_ring_start(h_rep, (uint32_t)B.b, (uint32_t)B.len);
  return true;
}

// ------------------------------------------------------------------------------------------ the pop loop
// The policy as the pop loop needs it is `Policy` above: float reciprocals of pi(.|s) plus the guard bit.
// The exact float64 entries are only needed when the filter is undecided; they are re-derived there by
// the caller-supplied functor (epsilon-greedy: from eps and a*, fixed policies: re-read from the table).

__device__ __forceinline__ void load_record32(const float4* rp, float4& p, float4& tail) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(p.x), "=f"(p.y), "=f"(p.z), "=f"(p.w), "=f"(tail.x), "=f"(tail.y), "=f"(tail.z), "=f"(tail.w)
               : "l"(rp));
}

// Return codes of pop_round besides the OLS_* statuses
constexpr int kPopIdle = -1;      // the group did not take part in this round
constexpr int kPopContinue = -2;  // every record examined was rejected; the queue is not exhausted yet

// ONE pop round, warp-convergent (every lane of the warp calls it; `active` is group-uniform).  Lane groups
// are decoupled: each executes its own round against its own bucket, and the caller loops.  Outcomes per
// group: OLS_RUNNING = a record was accepted (its tail is returned; tail.z is the log row), kPopContinue,
// OLS_QUEUE_EMPTY (psrs.py:44-45, detected when a round starts on a drained queue), OLS_FEED_EMPTY.
// B.cur, draws, n_pop (and the PCG64 state) advance by the records consumed.
template <int NV, int RNG, int PPL, typename PolT, typename PnFn>
__device__ __forceinline__ int pop_round(const ols_queues_t& q, const ols_replicas_t& R, int G, int sub, int gbase,
                                         unsigned gmask, bool active, int64_t r, uint64_t rid, const PolT& pol,
                                         PnFn pn_fn, const ols_philox_keys& PK, int compute_mode, Bucket& B, uint64_t& draws,
                                         uint32_t& n_pop, ols_u128& pcg_s, const ols_u128& pcg_inc, float4& win_tail,
                                         int prefetch = 0) {
  constexpr int NA = 4 * NV;
  constexpr int kRec = 16 * (NV + 1);  // record bytes
  constexpr bool accept_all = (RNG == RNGT_NONE);
  // Philox (protocol v2, ols_common.cuh): one block of the stream carries the high words of four consecutive draws.  A
  // round is aligned so that a lane's PPL pops fall inside ONE block (PPL = 4: a whole block, PPL = 2: half of one); the
  // slots before `draws` are masked out.
  constexpr bool blocked = (RNG == RNGT_PHILOX);
  static_assert(RNG != RNGT_PCG64 || PPL == 1, "PCG64 replay is sequential per lane group: one pop per lane");
  static_assert(!accept_all || PPL == 1, "accept-all examines one record per round");
  typedef typename std::conditional<RNG == RNGT_PHILOX, float, double>::type ustore_t;
  const int nA = q.n_actions;
  int result = active ? kPopContinue : kPopIdle;

  const int remaining = B.len - B.cur;
  if (active && remaining <= 0) {
    result = OLS_QUEUE_EMPTY;
    active = false;
  }
  const int align = blocked ? (int)(draws & (uint64_t)(PPL - 1)) : 0;
  const int n_eval = active ? min(accept_all ? 1 : G * PPL - align, remaining) : 0;
  const int k0 = sub * PPL - align;  // queue offset (from the cursor) of this lane's first pop; -1 = masked slot

  // ---- issue all record loads of the round first
  float4 pv[PPL][NV], tail[PPL];
  bool valid[PPL];
#pragma unroll
  for (int i = 0; i < PPL; ++i) {
    valid[i] = (unsigned)(k0 + i) < (unsigned)n_eval;  // 0 <= k < n_eval
    tail[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int v = 0; v < NV; ++v) pv[i][v] = make_float4(1.f, 1.f, 1.f, 1.f);
    if (valid[i]) {
      uint32_t pos = B.start + (uint32_t)(B.cur + k0 + i);
      if (pos >= (uint32_t)B.len) pos -= (uint32_t)B.len;
      const float4* rp = (const float4*)((const char*)q.records + (size_t)(uint32_t)(B.lo + (int32_t)pos) * kRec);
      if (NV == 1) {
        // a 4-action record is 32 bytes, 32-byte aligned: one 256-bit load (LDG.E.256, new with sm_100) instead of
        // two 128-bit ones halves the L1 wavefronts of the gather
        load_record32(rp, pv[i][0], tail[i]);
      } else {
#pragma unroll
        for (int v = 0; v < NV; ++v) pv[i][v] = __ldg(rp + v);
        tail[i] = __ldg(rp + NV);
      }
    }
  }
  // ---- long reject chains: pull the records this lane would examine `prefetch` rounds ahead into L2
  if (prefetch > 0 && active) {
    const int kp = k0 + prefetch * G * PPL;
#pragma unroll
    for (int i = 0; i < PPL; ++i) {
      if ((unsigned)(kp + i) < (unsigned)remaining) {
        uint32_t pos = B.start + (uint32_t)(B.cur + kp + i);
        if (pos >= (uint32_t)B.len) pos -= (uint32_t)B.len;
        asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)q.records + (size_t)(uint32_t)(B.lo + (int32_t)pos) * kRec));
      }
    }
  }
  // ---- uniforms of this lane's pops: ustore_t is a float image for Philox (the exact double is re-derived on the
  // slow path), the exact double otherwise
  ustore_t u[PPL];
  bool lack[PPL], tiny[PPL];
#pragma unroll
  for (int i = 0; i < PPL; ++i) {
    u[i] = (ustore_t)0;
    lack[i] = tiny[i] = false;
  }
  ols_u128 stepped = pcg_s;
  if (blocked) {
    bool any = false;
#pragma unroll
    for (int i = 0; i < PPL; ++i) any = any || valid[i];
    if (any) {
      const uint64_t d0 = draws + (uint64_t)(int64_t)k0;      // this lane's first slot: a multiple of PPL
      const ols_u4 o = ols_stream_block(PK, rid, 0u, d0 >> 2);
#pragma unroll
      for (int i = 0; i < PPL; ++i) {
        const uint32_t hi = (PPL == 4) ? ols_u4_word(o, i) : ols_u4_word(o, (int)(d0 & 3) + i);
        u[i] = (ustore_t)ols_stream_float(PK, rid, 0u, d0 + (uint64_t)i, hi, tiny[i]);
      }
    }
  } else if (!accept_all) {
#pragma unroll
    for (int i = 0; i < PPL; ++i) {
      if (valid[i]) {
        const int k = k0 + i;
        const uint64_t d = draws + (uint64_t)k;
        if (RNG == RNGT_FEED) {
          if ((int64_t)d < R.u_feed_len) u[i] = (ustore_t)__ldg(R.u_feed + r * R.u_stride + (int64_t)d);
          else lack[i] = true;
        } else if (RNG == RNGT_PCG64) {
          // PCG64 always runs with PPL == 1: pop k needs the group's generator stepped k + 1 times
          for (int t = 0; t <= k; ++t) stepped = ols_pcg_step(stepped, pcg_inc);
          u[i] = (ustore_t)ols_u64_to_unit(ols_pcg_output(stepped));
        }
        tiny[i] = !((float)u[i] >= 8.6736174e-19f);  // u < 2^-60 (incl. 0): not resolved by the float32 filter
      }
    }
  }
  // ---- accept rule per pop
  const Policy<NA> P = pol.get();
  bool acc[PPL];
#pragma unroll
  for (int i = 0; i < PPL; ++i) {
    acc[i] = false;
    if (accept_all) {
      acc[i] = valid[i];  // trivial_baselines.py:8-10: no uniform is drawn
    } else if (valid[i] && !lack[i]) {
      const uint32_t meta = __float_as_uint(tail[i].w);
      const int a = ols_meta_action(meta);
      float plf[NA];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        plf[4 * v + 0] = pv[i][v].x;
        plf[4 * v + 1] = pv[i][v].y;
        plf[4 * v + 2] = pv[i][v].z;
        plf[4 * v + 3] = pv[i][v].w;
      }
      int dec = accept_filter<NA>(P, plf, a, (float)u[i], P.ok && !tiny[i] && ols_meta_tame(meta) != 0);
      if (dec == 2) {
        const int k = k0 + i;
        uint32_t pos = B.start + (uint32_t)(B.cur + k);
        if (pos >= (uint32_t)B.len) pos -= (uint32_t)B.len;
        const double ue = (RNG == RNGT_PHILOX) ? ols_stream_uniform(PK, rid, 0u, draws + (uint64_t)(int64_t)k) : (double)u[i];
        const double* p64 = q.p_log64 ? q.p_log64 + (int64_t)__float_as_int(tail[i].z) * nA : nullptr;
        dec = (int)accept_slow(pn_fn, (const float*)((const char*)q.records + (size_t)(uint32_t)(B.lo + (int32_t)pos) * kRec),
                               p64, nA, a, ue, compute_mode);
      }
      acc[i] = (dec == 1);
    }
  }
  // ---- first accepting pop of each group, in queue order
  int first_acc = 1 << 20, first_lack = 1 << 20;
  if (PPL == 2) {
    const unsigned g_any = (__ballot_sync(kFull, acc[0] || acc[PPL - 1]) >> gbase) & gmask;
    const unsigned g_a0 = (__ballot_sync(kFull, acc[0]) >> gbase) & gmask;
    if (g_any) {
      const int w = __ffs(g_any) - 1;
      first_acc = 2 * w + (((g_a0 >> w) & 1u) ? 0 : 1) - align;
    }
  } else {
    // each lane's earliest accepting pop, then the earliest lane of the group
    int mine = PPL;
#pragma unroll
    for (int i = PPL - 1; i >= 0; --i) mine = acc[i] ? i : mine;
    const unsigned g_any = (__ballot_sync(kFull, mine < PPL) >> gbase) & gmask;
    const int w = g_any ? __ffs(g_any) - 1 : 0;
    const int wi = __shfl_sync(kFull, mine, gbase + w);
    if (g_any) first_acc = w * PPL + wi - align;
  }
  if (RNG == RNGT_FEED) {
#pragma unroll
    for (int i = 0; i < PPL; ++i) {
      const unsigned bl = (__ballot_sync(kFull, lack[i]) >> gbase) & gmask;
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
      result = OLS_FEED_EMPTY;
    } else {
      consumed = n_eval;
    }
  }
  // winner's record tail -> every lane of its group (shuffles are executed by the whole warp)
  const int wk = (first_acc < (1 << 20)) ? first_acc + align : 0;
  const int src = gbase + wk / PPL;
  float4 sel = tail[0];
#pragma unroll
  for (int i = 1; i < PPL; ++i)
    if ((wk % PPL) == i) sel = tail[i];
  sel.x = __shfl_sync(kFull, sel.x, src);
  sel.y = __shfl_sync(kFull, sel.y, src);
  sel.z = __shfl_sync(kFull, sel.z, src);
  sel.w = __shfl_sync(kFull, sel.w, src);
  if (RNG == RNGT_PCG64) {
    // the generator advanced by `consumed` draws: that is the stepped state of the lane that evaluated
    // pop (consumed - 1); PCG64 runs with PPL == 1, so that lane is gbase + consumed - 1
    const int adv_src = gbase + (consumed > 0 ? consumed - 1 : 0);
    const uint64_t nh = __shfl_sync(kFull, stepped.hi, adv_src);
    const uint64_t nl = __shfl_sync(kFull, stepped.lo, adv_src);
    if (consumed > 0) {
      pcg_s.hi = nh;
      pcg_s.lo = nl;
    }
  }
  if (took) {
    win_tail = sel;
    result = OLS_RUNNING;
  }
  B.cur += consumed;
  n_pop += (uint32_t)consumed;
  if (!accept_all) draws += (uint64_t)consumed;
  return result;
}

// ONE pop round over HOT records (ols_queues_t.hot: 8 bytes per record, see ols_common.cuh): every lane examines four
// consecutive records = 32 bytes, a lane pair one 64-byte DRAM atom, where the full records cost 128 bytes for half as
// many pops.  Same contract as pop_round<1, RNGT_PHILOX, 4>.  Decision ladder per pop: half-precision estimate with a
// 2^-9 band -> float32 estimate from the full record with the 2^-16 band -> exact rule; the full record of the ACCEPTED
// pop (reward, next state, row, terminal) is fetched once the winner is known, by every lane of the group (one sector).
template <typename PolT, typename PnFn>
__device__ __forceinline__ int pop_round_hot(const ols_queues_t& q, int G, int sub, int gbase, unsigned gmask, bool active,
                                             uint64_t rid, const PolT& pol, PnFn pn_fn, const ols_philox_keys& PK,
                                             int compute_mode, Bucket& B, uint64_t& draws, float4& win_tail) {
  constexpr int PPL = 4, NA = 4, kRec = 32;
  const int nA = q.n_actions;
  int result = active ? kPopContinue : kPopIdle;
  const int remaining = B.len - B.cur;
  if (active && remaining <= 0) {
    result = OLS_QUEUE_EMPTY;
    active = false;
  }
  const int align = (int)(draws & 3);  // rounds are aligned to a multiple of four draws: each lane needs exactly ONE stream block
  const int n_eval = active ? min(G * PPL - align, remaining) : 0;
  const int k0 = sub * PPL - align;
  uint2 h[PPL];
  bool valid[PPL];
#pragma unroll
  for (int i = 0; i < PPL; ++i) {
    valid[i] = (unsigned)(k0 + i) < (unsigned)n_eval;
    h[i] = make_uint2(0u, 0u);
    if (valid[i]) {
      uint32_t pos = B.start + (uint32_t)(B.cur + k0 + i);
      if (pos >= (uint32_t)B.len) pos -= (uint32_t)B.len;
      h[i] = __ldg((const uint2*)q.hot + (size_t)(uint32_t)(B.lo + (int32_t)pos));
    }
  }
  float u[PPL];
  bool tiny[PPL];
#pragma unroll
  for (int i = 0; i < PPL; ++i) {
    u[i] = 0.f;
    tiny[i] = false;
  }
  if (valid[0] || valid[1] || valid[2] || valid[3]) {
    const uint64_t d0 = draws + (uint64_t)(int64_t)k0;  // a multiple of four
    const ols_u4 o = ols_stream_block(PK, rid, 0u, d0 >> 2);
#pragma unroll
    for (int i = 0; i < PPL; ++i) u[i] = ols_stream_float(PK, rid, 0u, d0 + (uint64_t)i, ols_u4_word(o, i), tiny[i]);
  }
  const Policy<NA> P = pol.get();
  bool acc[PPL];
#pragma unroll
  for (int i = 0; i < PPL; ++i) {
    acc[i] = false;
    if (valid[i]) {
      float plf[NA];
      int a;
      bool tame;
      ols_hot_unpack(h[i], plf, a, tame);
      int dec = accept_filter<NA>(P, plf, a, u[i], P.ok && !tiny[i] && tame, 1.953125e-3f /* 2^-9 */);
      if (dec == 2) {
        const int k = k0 + i;
        uint32_t pos = B.start + (uint32_t)(B.cur + k);
        if (pos >= (uint32_t)B.len) pos -= (uint32_t)B.len;
        const float4* rp = (const float4*)((const char*)q.records + (size_t)(uint32_t)(B.lo + (int32_t)pos) * kRec);
        float4 pf, tl;
        load_record32(rp, pf, tl);
        const uint32_t meta = __float_as_uint(tl.w);
        const float full[NA] = {pf.x, pf.y, pf.z, pf.w};
        dec = accept_filter<NA>(P, full, ols_meta_action(meta), u[i], P.ok && !tiny[i] && ols_meta_tame(meta) != 0);
        if (dec == 2) {
          const double ue = ols_stream_uniform(PK, rid, 0u, draws + (uint64_t)(int64_t)k);
          const double* p64 = q.p_log64 ? q.p_log64 + (int64_t)__float_as_int(tl.z) * nA : nullptr;
          dec = (int)accept_slow(pn_fn, (const float*)rp, p64, nA, ols_meta_action(meta), ue, compute_mode);
        }
      }
      acc[i] = (dec == 1);
    }
  }
  // ---- first accepting pop of the group, in queue order
  int first_acc = 1 << 20;
  {
    int mine = PPL;
#pragma unroll
    for (int i = PPL - 1; i >= 0; --i) mine = acc[i] ? i : mine;
    if (G == 1) {  // one lane per replica (a compile-time fact in the production kernel): nothing to vote on
      if (mine < PPL) first_acc = mine - align;
    } else {
      const unsigned g_any = (__ballot_sync(kFull, mine < PPL) >> gbase) & gmask;
      const int w = g_any ? __ffs(g_any) - 1 : 0;
      const int wi = __shfl_sync(kFull, mine, gbase + w);
      if (g_any) first_acc = w * PPL + wi - align;
    }
  }
  int consumed = 0;
  bool took = false;
  if (active) {
    if (first_acc < (1 << 20)) {
      consumed = first_acc + 1;
      took = true;
    } else {
      consumed = n_eval;
    }
  }
  if (took) {
    uint32_t pos = B.start + (uint32_t)(B.cur + first_acc);
    if (pos >= (uint32_t)B.len) pos -= (uint32_t)B.len;
    float4 pf;
    load_record32((const float4*)((const char*)q.records + (size_t)(uint32_t)(B.lo + (int32_t)pos) * kRec), pf, win_tail);
    result = OLS_RUNNING;
  }
  B.cur += consumed;
  draws += (uint64_t)consumed;
  return result;
}

// ---- logs whose behaviour distribution is a function of the state (every queue holds ONE distribution, ols_queues_t.queue_dist)
// With t_j = p_j / pi_j and inv_m = min_j t_j the threshold of a record with action a is thr_a = inv_m / t_a (psrs.py:55-57
// rearranged): it depends on the record only through its ACTION.  A pop's uniform is known to the pre-filter through its
// high word hi (u in [hi, hi + 1) * 2^-32).  Per STEP and action the band around the float32 estimate of thr_a (within 2^-19
// relative of the exact value under the guards; band 2^-15) becomes two 32-bit limits:
//     hi <  lim.x  ==>  u <  thr_a: accept          hi >  lim.y  ==>  u >  thr_a: reject          else: the exact rule
// so a pop costs an 8-byte shared-memory load and two integer compares.  A failed guard (untamed queue distribution, policy
// entry outside {0} U [2^-40, 1]) sets lim = {0, 2^32 - 1}: every pop goes to the exact rule.
__device__ __forceinline__ void make_limits(float4 p, bool ok, const float (&ipn)[4], uint2 (&lim)[4]) {
  // padding entries (j >= n_actions) hold p = 1 in the queue's distribution and 1/pi = +inf: t_j = +inf never wins the min
  const float t[4] = {p.x * ipn[0], p.y * ipn[1], p.z * ipn[2], p.w * ipn[3]};
  const float inv_m = fminf(fminf(t[0], t[1]), fminf(t[2], t[3]));
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const float thr = __fdiv_rn(inv_m, t[a]);  // 0 when pi_a = 0 (t_a = +inf); NaN only when a guard failed; <= 1 otherwise
    // in units of 2^-32, rounded outward, with two units of slack on either side for the truncation of u to hi; the
    // float32 products are exact to 2^-24 relative, which the 2^-15 factors absorb (band 2^-16 + rounding)
    const float lo = __fmul_rd(thr, 4294836224.0f /* (1 - 2^-15) 2^32 */), hi = __fmul_ru(thr, 4295098368.0f /* (1 + 2^-15) 2^32 */);
    const bool fine = ok && (thr == thr);
    const unsigned lo32 = __float2uint_rd(lo), hi32 = __float2uint_ru(hi);  // saturating conversions
    lim[a].x = !fine ? 0u : (lo32 > 2u ? lo32 - 2u : 0u);
    lim[a].y = !fine ? 0xffffffffu : (hi32 < 0xfffffffdu ? hi32 + 2u : 0xffffffffu);
  }
}

// ONE pop round over TAIL hot records (OLS_HOT_TAIL: { reward; next state | action | terminal }, 8 bytes per record): same
// contract as pop_round_hot, but the accept rule is the per-step limits `lim` (the group's Cold block) applied to the
// record's action — no per-record arithmetic — and the accepted pop's fields come from the hot record itself; the full
// record is read only inside the band (the exact rule needs nothing of it but the queue's distribution) or when the caller
// needs the log row (`need_row`: row trace, or agent observation ids distinct from the state ids).
template <typename PnFn>
__device__ __forceinline__ int pop_round_tail(const ols_queues_t& q, int G, int sub, int gbase, unsigned gmask, bool active,
                                              uint64_t rid, const volatile uint2* lim, PnFn pn_fn, const ols_philox_keys& PK,
                                              int compute_mode, bool need_row, Bucket& B, uint64_t& draws, float4& win_tail) {
  constexpr int PPL = 4, kRec = 32;
  const int nA = q.n_actions;
  int result = active ? kPopContinue : kPopIdle;
  const int remaining = B.len - B.cur;
  if (active && remaining <= 0) {
    result = OLS_QUEUE_EMPTY;
    active = false;
  }
  const int align = (int)(draws & 3);  // rounds are aligned to a multiple of four draws: each lane needs exactly ONE stream block
  const int n_eval = active ? min(G * PPL - align, remaining) : 0;
  const int k0 = sub * PPL - align;
  uint2 h[PPL];
  bool valid[PPL];
#pragma unroll
  for (int i = 0; i < PPL; ++i) {
    valid[i] = (unsigned)(k0 + i) < (unsigned)n_eval;
    h[i] = make_uint2(0u, 0u);
    if (valid[i]) {
      uint32_t pos = B.start + (uint32_t)(B.cur + k0 + i);
      if (pos >= (uint32_t)B.len) pos -= (uint32_t)B.len;
      h[i] = __ldg((const uint2*)q.hot + (size_t)(uint32_t)(B.lo + (int32_t)pos));
    }
  }
  ols_u4 o = {0u, 0u, 0u, 0u};
  if (valid[0] || valid[1] || valid[2] || valid[3]) o = ols_stream_block(PK, rid, 0u, (draws + (uint64_t)(int64_t)k0) >> 2);
  int mine = PPL;
#pragma unroll
  for (int i = PPL - 1; i >= 0; --i) {
    const unsigned a = (h[i].y >> 29) & 3u;
    const uint2 lm = make_uint2(lim[a].x, lim[a].y);
    const uint32_t hi = ols_u4_word(o, i);
    int dec = (!valid[i] || hi > lm.y) ? 0 : (hi < lm.x) ? 1 : 2;
    if (dec == 2) {
      const double ue = ols_stream_uniform(PK, rid, 0u, draws + (uint64_t)(int64_t)(k0 + i));
      dec = (int)accept_slow(pn_fn, q.queue_dist + 8 * (size_t)B.b, nullptr, nA, (int)a, ue, compute_mode);
    }
    mine = (dec == 1) ? i : mine;
  }
  // ---- first accepting pop of the group, in queue order; its hot record goes to every lane of the group
  int first_acc = 1 << 20;
  uint2 won = (mine == 0) ? h[0] : (mine == 1) ? h[1] : (mine == 2) ? h[2] : h[3];
  if (G == 1) {
    if (mine < PPL) first_acc = mine - align;
  } else {
    const unsigned g_any = (__ballot_sync(kFull, mine < PPL) >> gbase) & gmask;
    const int w = g_any ? __ffs(g_any) - 1 : 0;
    const int wi = __shfl_sync(kFull, mine, gbase + w);
    won.x = __shfl_sync(kFull, won.x, gbase + w);
    won.y = __shfl_sync(kFull, won.y, gbase + w);
    if (g_any) first_acc = w * PPL + wi - align;
  }
  int consumed = 0;
  bool took = false;
  if (active) {
    if (first_acc < (1 << 20)) {
      consumed = first_acc + 1;
      took = true;
    } else {
      consumed = n_eval;
    }
  }
  if (took) {
    const int z2 = ((int)(won.y << 3)) >> 3;
    const int action = (int)((won.y >> 29) & 3u), term = (int)(won.y >> 31);
    int row = -1;
    if (need_row) {
      uint32_t pos = B.start + (uint32_t)(B.cur + first_acc);
      if (pos >= (uint32_t)B.len) pos -= (uint32_t)B.len;
      row = __ldg((const int32_t*)((const char*)q.records + (size_t)(uint32_t)(B.lo + (int32_t)pos) * kRec) + 6);
    }
    win_tail = make_float4(__uint_as_float(won.x), __int_as_float(z2), __int_as_float(row), __uint_as_float(ols_pack_meta(action, term, 1)));
    result = OLS_RUNNING;
  }
  B.cur += consumed;
  draws += (uint64_t)consumed;
  return result;
}

// ------------------------------------------------------------------------------------------ kernels: step
template <int NV, int RNG>
__global__ void __launch_bounds__(kThreads) step_kernel(ols_queues_t q, ols_replicas_t R, const uint8_t* mask,
                                                         const void* p_new, int p_new_is_f64, int64_t p_stride,
                                                         int compute_mode, int G, ols_step_out_t out, ols_philox_keys PK,
                                                         const int32_t* act) {
  constexpr int NA = 4 * NV;
  const Lane L = make_lane(G);
  const int nA = q.n_actions;
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
  int status = OLS_IDLE, cur_z = 0;
  uint64_t draws = 0;
  ols_u128 pcg_s = {0, 0}, pcg_inc = {0, 0};
  bool active = false;
  if (exists) {
    status = R.status[r];
    cur_z = R.cur_z[r];
    draws = R.draws[r];
    if (RNG == RNGT_PCG64) {
      const uint64_t* p = R.pcg_state + 4 * r;
      pcg_s.hi = p[0];
      pcg_s.lo = p[1];
      pcg_inc.hi = p[2];
      pcg_inc.lo = p[3];
    }
    // a replica whose queue ran dry may be stepped again (it reports QUEUE_EMPTY again, psrs.py:44-45)
    active = (!mask || mask[r]) && (status == OLS_RUNNING || status == OLS_QUEUE_EMPTY);
  }
  const bool was_active = active;
  int32_t* cur_tab = R.cursor + rr * (int64_t)q.n_buckets;
  Bucket B;
  B.b = -1;
  B.lo = B.len = B.cur = 0;
  B.start = 0;
  int res = kPopIdle;
  // queue_evaluator.py:107,125 — queues keyed by (z, a): the key is (z + z_offset) * n_actions + a, and the agent's
  // own action picks the queue (act != NULL); otherwise the state alone does
  const int key_z = act ? ((cur_z + q.z_offset) * q.n_actions + (exists ? __ldg(act + rr) : 0) - q.z_offset) : cur_z;
  const bool act_ok = !act || ((unsigned)__ldg(act + rr) < (unsigned)q.n_actions);
  if (active && (!act_ok || !load_bucket(q, R, ols_ring_replica_hash(R.ring_seed, rid), cur_tab, key_z, B))) {
    res = OLS_NO_QUEUE;
    active = false;
  }
  uint32_t n_pop = 0;
  float4 tail = make_float4(0.f, 0.f, 0.f, 0.f);
  while (__any_sync(kFull, active)) {
    const int rc = pop_round<NV, RNG, 1>(q, R, L.G, L.sub, L.gbase, L.gmask, active, rr, rid, P, pn_fn, PK, compute_mode, B,
                                         draws, n_pop, pcg_s, pcg_inc, tail);
    if (active && rc != kPopContinue) {
      res = rc;
      active = false;
    }
  }
  if (exists && L.sub == 0) {
    if (was_active) {
      if (B.b >= 0) cur_tab[B.b] = B.cur;
      R.draws[r] = draws;
      R.pops[r] += n_pop;
      if (RNG == RNGT_PCG64) {
        R.pcg_state[4 * r] = pcg_s.hi;
        R.pcg_state[4 * r + 1] = pcg_s.lo;
      }
      if (res == OLS_RUNNING) {
        const uint32_t meta = __float_as_uint(tail.w);
        R.status[r] = OLS_RUNNING;
        R.cur_z[r] = __float_as_int(tail.y);  // psrs.py:49-50
        const int32_t row = __float_as_int(tail.z);
        const int32_t s2 = q.s_next ? __ldg(q.s_next + row) : __float_as_int(tail.y);
        R.cur_s[r] = s2;
        R.accepted[r] += 1;
        out.row[r] = row;
        out.action[r] = ols_meta_action(meta);
        out.reward[r] = tail.x;
        out.terminal[r] = (uint8_t)ols_meta_terminal(meta);
        out.s_next[r] = s2;
        out.z_from[r] = cur_z;
      } else {
        R.status[r] = res;
        out.row[r] = -1;
      }
    } else if (!mask || mask[r]) {
      out.row[r] = -1;
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
  // float32 production path: the same expressions with every operand rounded to float32 first (gamma, alpha(ep) and the
  // reward are float32 images) and one IEEE float32 operation at a time — the order is part of the contract
  // (oracle/ols_oracle.c `q_is_f32`), never left to the compiler's FMA contraction
  static __device__ __forceinline__ float td(float R, float gamma, float boot, float q) {
    return __fsub_rn(__fadd_rn(R, __fmul_rn(gamma, boot)), q);
  }
  static __device__ __forceinline__ float upd(float q, float alpha, float td) { return __fadd_rn(q, __fmul_rn(alpha, td)); }
};

// Shared-memory plan of one CTA (host-computed, see ols_sim.cu).  Every lane group owns
//   * a Cold block: the replica's per-episode registers, touched once per accepted transition, kept out
//     of the register file so the pop rounds run at high occupancy;
//   * optionally a staging slot holding its replica's cursor table and, when it fits, its Q table.
struct StagePlan {
  int cursor_bytes;  // 0 = cursors stay in global memory
  int q_bytes;       // 0 = Q stays in global memory
  int slot_bytes;
};

struct __align__(16) Cold {
  double ep_G, ep_disc;
  unsigned long long ties;
  unsigned long long accepted0;
  int ep_index, ep_len_count, ep_t, ep_open;
  int init_cursor, status, n_acc, pn_ok;
  float ipn_base, ipn_top;  // 1 / (eps/nA) and 1 / (1 - eps + eps/nA) of the open episode (QLEARN)
  int eps_episode;          // the episode index these were computed for (-1 = none)
  int budget;               // accepted transitions this launch (or time slice) may still add; 0 = unlimited
  float qrow[4];            // AoS Q-learning: Q[S, :] of the current state (the bootstrap row of the previous step)
  unsigned h_rep;           // ring hash of the replica
  int pad[3];
  uint2 lim[4];             // TAIL hot records: per action, accept below .x / reject above .y (high word of the uniform)
};

// 128-bit accesses to a group's Cold block (the struct is volatile for the scalar fields)
__device__ __forceinline__ float4 cold_ld4(const volatile float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"((unsigned)__cvta_generic_to_shared((const void*)p))
               : "memory");
  return v;
}
__device__ __forceinline__ void cold_st4(volatile float* p, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"((unsigned)__cvta_generic_to_shared((const void*)p)), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// epsilon-greedy policy source (tabular.py:24-32): two distinct probabilities per episode, their float reciprocals
// cached in the group's Cold block; which action gets the larger one is the only per-step state.
template <int NA>
struct EpsGreedySource {
  const volatile Cold* C;
  int a_star, nA;
  __device__ __forceinline__ Policy<NA> get() const {
    Policy<NA> P;
    const float ib = C->ipn_base, it = C->ipn_top;
#pragma unroll
    for (int j = 0; j < NA; ++j) P.ipn[j] = (j >= nA) ? __int_as_float(0x7f800000) : ((j == a_star) ? it : ib);
    P.ok = C->pn_ok != 0;
    return P;
  }
};

// The run loop.  Lane groups of a warp are DECOUPLED: every iteration each live group executes one pop round
// against its own queue; a group whose round accepted a record runs the learner update and prepares its next
// step in the same iteration (predicated), so no group waits for the longest reject chain in its warp.
template <int NV, int AGENT, typename QT, int RNG, int PPL, bool AOS = false, int GT = 0, int HOT = 0>
__global__ void __launch_bounds__((GT == 1 ? OLS_G1_THREADS : kThreads), (NV == 1 ? (GT == 1 ? OLS_MINB_G1 : OLS_MINB) : 2)) run_kernel(ols_queues_t q, ols_replicas_t R, ols_run_t cfg,
                                                                                  StagePlan plan, ols_philox_keys PK) {
  constexpr int NA = 4 * NV;
  // AoS Q-learning keeps the policy row Q[S, :] in the group's Cold block (shared memory) between the bootstrap read of
  // one step and the action choice of the next, and re-forms the float32 policy from the episode's two cached
  // reciprocals every round: neither is held in registers across the pop rounds (they were spilled to local memory).
  constexpr bool QCARRY = AOS && AGENT == OLS_AGENT_QLEARN;
  static_assert(!AOS || (NV == 1 && sizeof(QT) == 4 && AGENT != OLS_AGENT_FIXED_PI), "AoS state: float32 learners, <= 4 actions");
  static_assert(!HOT || (NV == 1 && RNG == RNGT_PHILOX && PPL == 4), "hot records: <= 4 actions, Philox uniforms, 4 per lane");
  static_assert(HOT != 2 || (AGENT == OLS_AGENT_FIXED_PI && !AOS), "TAIL hot records: fixed-policy evaluation here, the Q-learner in ols_learner.cu");
  extern __shared__ __align__(16) unsigned char s_mem[];
  const int nA = q.n_actions;
  const int nS = cfg.n_states;
  const int G = GT ? GT : cfg.lanes;  // GT: lane-group width known at compile time (the production widths 2 and 4)
  const int lane = threadIdx.x & 31;
  const int sub = lane & (G - 1);
  const int gbase = lane - sub;
  const unsigned gmask = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
  // ---- the CTA's unit of work.  Plain launch: replica block blockIdx.x, budget cfg.max_steps.  Time-sliced launch
  // (cfg.slices > 1): the grid holds slices x blocks CTAs; a CTA draws a ticket u and serves slice u / blocks of
  // replica block u % blocks once that block's previous slice has been published.  Tickets rise with start time,
  // so the awaited slice always belongs to a CTA that is already running: no deadlock, and the SMs stay full when
  // the blocks do not divide evenly into waves (977 blocks over 592 resident CTAs = 1.65 waves cost 2 unsliced).
  __shared__ int s_ticket;
  int blk = blockIdx.x;
  long long budget64 = cfg.max_steps;
  if (cfg.slices > 1) {
    if (threadIdx.x == 0) s_ticket = atomicAdd(cfg.sched, 1);
    __syncthreads();
    const int n_blk = gridDim.x / cfg.slices;
    blk = s_ticket % n_blk;
    const int slice = s_ticket / n_blk;
    const long long per = (cfg.max_steps + cfg.slices - 1) / cfg.slices;
    budget64 = min(per, cfg.max_steps - (long long)slice * per);  // <= 0: nothing left, just pass the token on
    if (threadIdx.x == 0) {
      const volatile int* turn = (const volatile int*)cfg.sched + 1 + blk;
      while (*turn != slice) __nanosleep(256);
      __threadfence();  // acquire: the previous slice's replica state (written on another SM) is visible below
    }
    __syncthreads();
  }
  // one lane group per replica, one warp per 32/G replicas (the blocks cover all of them)
  const bool no_budget = cfg.slices > 1 && budget64 <= 0;
  const int budget0 = (int)max(0ll, min(budget64, 0x7fffffffll));  // n_acc is an int: a launch adds < 2^31 transitions
  const int64_t r = ((blk * (int64_t)blockDim.x + threadIdx.x) >> 5) * (32 / G) + (lane / G);
  const bool exists = r < R.n_replicas;
  const int64_t rr = exists ? r : 0;
  const uint64_t rid = (uint64_t)(R.replica_id0 + rr);
  const uint32_t h_rep = R.ring_mode ? ols_ring_replica_hash(R.ring_seed, rid) : 0u;
  const int group = threadIdx.x / G;
  volatile Cold* C = (volatile Cold*)s_mem + group;
  unsigned char* my_slot = s_mem + sizeof(Cold) * (blockDim.x / G) + (size_t)group * plan.slot_bytes;

  // hot registers
  int cur_z = 0, cur_s = 0;
  uint64_t draws = 0;
  uint32_t n_pop = 0;  // only read back by the accept-all baselines; otherwise pops == uniforms drawn
  ols_u128 pcg_s = {0, 0}, pcg_inc = {0, 0};
  bool live = false;
  // every lane of the group writes the same values into the group's Cold block
  if (!exists) C->status = OLS_IDLE;  // (one write per field and lane: the lanes of a group are not in lockstep)
  C->n_acc = 0;
  C->eps_episode = -1;
  C->budget = budget0;
  C->h_rep = h_rep;
  if (exists) {
    const int status = __ldcg(R.status + r);
    cur_z = __ldcg(R.cur_z + r);
    cur_s = AOS ? cur_z : __ldcg(R.cur_s + r);  // AoS state: the observation id IS the queue key
    draws = __ldcg(R.draws + r);
    if (RNG == RNGT_PCG64) {
      const uint64_t* p = R.pcg_state + 4 * r;
      pcg_s.hi = __ldcg(p);
      pcg_s.lo = __ldcg(p + 1);
      pcg_inc.hi = __ldcg(p + 2);
      pcg_inc.lo = __ldcg(p + 3);
    }
    C->status = status;
    C->init_cursor = __ldcg(R.init_cursor + r);
    C->ties = __ldcg(R.ties + r);
    C->accepted0 = __ldcg(R.accepted + r);
    C->ep_index = __ldcg(R.ep_index + r);
    C->ep_len_count = __ldcg(R.ep_len_count + r);
    C->ep_t = __ldcg(R.ep_t + r);
    C->ep_open = __ldcg(R.ep_open + r);
    C->ep_G = __ldcg(R.ep_G + r);
    C->ep_disc = __ldcg(R.ep_disc + r);
    live = (status == OLS_IDLE || status == OLS_RUNNING) && !no_budget;  // terminated replicas stay terminated
  }
  int32_t* cur_tab = AOS ? nullptr : R.cursor + rr * (int64_t)q.n_buckets;
  QT* Qr = (AGENT == OLS_AGENT_FIXED_PI || AOS) ? nullptr : ((QT*)cfg.Q + rr * (int64_t)nS * nA);
  unsigned char* aos = AOS ? (unsigned char*)cfg.state_aos + (size_t)rr * (size_t)nS * 32 : nullptr;
  // ---- stage the replica's tables in shared memory (the group's lanes copy cooperatively)
  if (!AOS && plan.cursor_bytes) {
    int32_t* sc = (int32_t*)my_slot;
    if (exists)
      for (int i = sub; i < q.n_buckets; i += G) sc[i] = cur_tab[i];
    cur_tab = sc;
  }
  if (!AOS && AGENT != OLS_AGENT_FIXED_PI && plan.q_bytes) {
    QT* sq = (QT*)(my_slot + plan.cursor_bytes);
    if (exists)
      for (int i = sub; i < nS * nA; i += G) sq[i] = Qr[i];
    Qr = sq;
  }
  __syncwarp();

  QT row[NA];  // Q[S, :] of the current state, carried over from the bootstrap read of the previous step
#pragma unroll
  for (int j = 0; j < NA; ++j) row[j] = (QT)0;
  bool row_valid = false;
  Bucket B;  // the queue of the current state, loaded as soon as the state is known
  B.b = -1;
  B.lo = B.len = B.cur = 0;
  B.start = 0;
  bool bucket_valid = false;
  bool need_step = true;  // the group has to (re)start a step: episode start, policy
  int a_star = 0;
  Policy<NA> P;
#pragma unroll
  for (int j = 0; j < NA; ++j) P.ipn[j] = 1.f;
  P.ok = false;

  // Lanes of one group execute the step code redundantly and keep shared state (the Cold block, the cursor table,
  // the Q table) in memory they all write.  The programming model does not promise that they run in lockstep, so
  // every block below is phased: all reads of shared state -> __syncwarp(gm) -> all writes (identical values from
  // every lane) -> __syncwarp(gm).  gm is the group's own lane mask.
  const unsigned gm = gmask << gbase;

  while (__any_sync(kFull, live)) {
    if (live && need_step) {
      need_step = false;
      // ---------------- phase 1: reads
      const int ep_open0 = C->ep_open, ep_index0 = C->ep_index;
      int ic = C->init_cursor;
      const unsigned long long ties0 = C->ties;
      const int eps_episode0 = C->eps_episode;
      int new_status = -1;       // >= 0: terminate with this status
      bool started = false;      // a new episode starts here
      // ---- episode start: psrs.py:149-156 / 205-212 / 246-253
      if (!ep_open0) {
        if (ep_index0 >= cfg.n_episodes) {
          new_status = OLS_DONE;
        } else if (!reset_one(q, R, C->h_rep, ic, cur_z, cur_s)) {
          new_status = OLS_INIT_EMPTY;  // terminate without appending (psrs.py:153-155)
        } else {
          started = true;
          row_valid = false;
          bucket_valid = false;
        }
      }
      if (AOS) cur_s = cur_z;
      if (new_status < 0 && (cur_s < 0 || cur_s >= nS)) new_status = OLS_NO_QUEUE;  // IndexError on Q[S] / pi[S]
      QT prow[NA];         // Q[S, :] for the action choice
      bool fresh = false;  // prow was read from memory in this block
#pragma unroll
      for (int j = 0; j < NA; ++j) prow[j] = QCARRY ? (QT)0 : row[j];
      if (AOS) {
        if (new_status < 0 && !bucket_valid) {
          // episode start: one 256-bit load brings the state's cursor and Q row
          const AosEntry e = load_aos(aos, cur_s, cfg.state_epoch);
          if (!load_bucket(q, R, C->h_rep, nullptr, cur_z, B, true, e.cursor)) new_status = OLS_NO_QUEUE;
#pragma unroll
          for (int j = 0; j < NA; ++j) prow[j] = (j < nA) ? (QT)e.q[j & 3] : (QT)-INFINITY;
          fresh = true;
        }
      } else if (new_status < 0 && !bucket_valid && !load_bucket(q, R, C->h_rep, cur_tab, cur_z, B))
        new_status = OLS_NO_QUEUE;  // KeyError in the reference (psrs.py:44); no partial episode is appended
      // ---- the agent's action distribution for the current state
      bool used_tie = false;
      float ib = 0.f, it = 0.f;
      int pn_ok = 0;
      bool new_eps = false;
      uint2 step_lim[4];
      bool have_lim = false;
#pragma unroll
      for (int a = 0; a < 4; ++a) step_lim[a] = make_uint2(0u, 0xffffffffu);
      if (AGENT == OLS_AGENT_QLEARN && new_status < 0) {
        // tabular.py:24-32 epsilon_greedy_policy on the single row Q[[S], :] (psrs.py:158)
        if (!row_valid && !fresh) {
          fresh = true;
          if (AOS) {
            const AosEntry e = load_aos(aos, cur_s, cfg.state_epoch);
#pragma unroll
            for (int j = 0; j < NA; ++j) prow[j] = (j < nA) ? (QT)e.q[j & 3] : (QT)-INFINITY;
          } else {
#pragma unroll
            for (int j = 0; j < NA; ++j) prow[j] = (j < nA) ? Qr[(int64_t)cur_s * nA + j] : -INFINITY;  // padding never wins
          }
        } else if (QCARRY && !fresh) {
          const float4 v = cold_ld4(C->qrow);
          prow[0] = (QT)v.x;
          prow[1 % NA] = (QT)v.y;
          prow[2 % NA] = (QT)v.z;
          prow[3 % NA] = (QT)v.w;
        }
        QT mx = prow[0];
#pragma unroll
        for (int j = 1; j < NA; ++j) mx = (prow[j] > mx) ? prow[j] : mx;
        unsigned cand = 0;  // bit j set: action j attains the maximum
#pragma unroll
        for (int j = 0; j < NA; ++j) cand |= (prow[j] == mx) ? (1u << j) : 0u;
        const int ncand = __popc(cand);
        if (ncand > 1) {  // tabular.py:4-5 _random_argmax: a draw is consumed only on ties
          double v = 0.0;
          bool have = true;
          if (R.tie_mode == OLS_RNG_FEED) {
            if ((int64_t)ties0 < R.tie_feed_len) v = __ldg(R.tie_feed + rr * R.tie_stride + (int64_t)ties0);
            else have = false;
          } else {
            // stream 1, high word only: v = hi * 2^-32, so pick = floor(hi * ncand / 2^32) exactly
            v = (double)ols_u4_word(ols_stream_block(PK, rid, 1u, ties0 >> 2), (int)(ties0 & 3)) * 2.3283064365386963e-10;
          }
          if (!have) {
            new_status = OLS_FEED_EMPTY;
          } else {
            used_tie = true;
            const int pick = min(ncand - 1, (int)(v * (double)ncand));
            if (NA == 4) {  // drop the `pick` lowest candidates, branch-free for the 4-bit mask
              const unsigned c1 = cand & (cand - 1), c2 = c1 & (c1 - 1), c3 = c2 & (c2 - 1);
              cand = (pick == 0) ? cand : (pick == 1) ? c1 : (pick == 2) ? c2 : c3;
            } else {
              for (int t = 0; t < pick; ++t) cand &= cand - 1;
            }
          }
        }
        a_star = __ffs(cand) - 1;
        // epsilon-greedy has two distinct probabilities per episode (tabular.py:27,30): their float reciprocals and
        // the guard bit are cached in the Cold block when an episode's epsilon is first used
        if (eps_episode0 != ep_index0) {
          new_eps = true;
          const double eps = __ldg(cfg.epsilon + ep_index0);
          const double based = __ddiv_rn(eps, (double)nA);
          const float base = (float)based, top = (float)__dadd_rn(__dsub_rn(1.0, eps), based);
          const bool okb = (base == 0.f) || (base >= 9.094947e-13f && base <= 1.0f);
          const bool okt = (top == 0.f) || (top >= 9.094947e-13f && top <= 1.0f);
          ib = (base == 0.f) ? __int_as_float(0x7f800000) : __frcp_rn(base);
          it = (top == 0.f) ? __int_as_float(0x7f800000) : __frcp_rn(top);
          pn_ok = (okb && okt && (base > 0.f || top > 0.f)) ? 1 : 0;
        } else {
          ib = C->ipn_base;
          it = C->ipn_top;
          pn_ok = C->pn_ok;
        }
        // the float32 policy itself is re-formed from (ib, it, a*) before every pop round, see below
      } else if (new_status < 0) {
        const void* pi_tab = cfg.pi;
        const int is64 = cfg.pi_is_f64;
        const int64_t off = (int64_t)cur_s * nA;
        finish_policy<NA>(P, [=](int j) -> double {
          return is64 ? __ldg((const double*)pi_tab + off + j) : (double)__ldg((const float*)pi_tab + off + j);
        }, nA);
        if constexpr (HOT == 2) {
          // the accept rule of this step as per-action limits: the queue's one distribution against pi[S]
          const float4 qd = __ldg((const float4*)q.queue_dist + 2 * (size_t)B.b);
          const bool tame = (__ldg((const int*)q.queue_dist + 8 * (size_t)B.b + 4) & 2) != 0;
          const float ipn4[4] = {P.ipn[0], P.ipn[1 % NA], P.ipn[2 % NA], P.ipn[3 % NA]};
          make_limits(qd, P.ok && tame, ipn4, step_lim);
          have_lim = true;
        }
      }
      __syncwarp(gm);
      // ---------------- phase 2: writes
      if (QCARRY) {
        if (fresh) cold_st4(C->qrow, make_float4((float)prow[0], (float)prow[1 % NA], (float)prow[2 % NA], (float)prow[3 % NA]));
      } else {
#pragma unroll
        for (int j = 0; j < NA; ++j) row[j] = prow[j];
      }
      if (fresh) row_valid = true;
      if (started) {
        C->init_cursor = ic;
        C->status = OLS_RUNNING;
        C->ep_open = 1;
        C->ep_t = 0;
        C->ep_G = 0.0;
        C->ep_disc = 1.0;
      }
      if (used_tie) C->ties = ties0 + 1;
      if (have_lim) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          C->lim[a].x = step_lim[a].x;
          C->lim[a].y = step_lim[a].y;
        }
      }
      if (new_eps) {
        C->ipn_base = ib;
        C->ipn_top = it;
        C->pn_ok = pn_ok;
        C->eps_episode = ep_index0;
      }
      if (new_status >= 0) {
        C->status = new_status;
        live = false;
      }
      __syncwarp(gm);
    }
    // exact float64 pi(.|S) for the cold path: re-derived from (eps, a*) or re-read from the table when needed
    const int a_now = a_star, S_now = AOS ? cur_z : cur_s;
    const double* eps_tab = cfg.epsilon;
    const void* pi_tab = cfg.pi;
    const int pi_is_f64 = cfg.pi_is_f64;
    auto pn_fn = [=](int j) -> double {
      if (AGENT == OLS_AGENT_QLEARN) {
        const double eps = __ldg(eps_tab + C->ep_index);
        const double base = __ddiv_rn(eps, (double)nA);           // ones * eps / nA          (tabular.py:27)
        const double top = __dadd_rn(__dsub_rn(1.0, eps), base);  // 1 - eps + eps / nA       (tabular.py:30)
        return (j == a_now) ? top : base;
      }
      // psrs.py:214,255 — p = pi[S]
      const int64_t off = (int64_t)S_now * nA + j;
      return pi_is_f64 ? __ldg((const double*)pi_tab + off) : (double)__ldg((const float*)pi_tab + off);
    };
    // ---- env.step(p), one round of it: psrs.py:39-57
    float4 tail = make_float4(0.f, 0.f, 0.f, 0.f);
    int res;
    if constexpr (HOT == 2) {
      res = pop_round_tail(q, G, sub, gbase, gmask, live, rid, C->lim, pn_fn, PK, cfg.compute_mode,
                           cfg.trace_row != nullptr || q.s_next != nullptr, B, draws, tail);
    } else if constexpr (HOT == 1) {
      if (AGENT == OLS_AGENT_QLEARN) {
        EpsGreedySource<NA> src;
        src.C = C;
        src.a_star = a_star;
        src.nA = nA;
        res = pop_round_hot(q, G, sub, gbase, gmask, live, rid, src, pn_fn, PK, cfg.compute_mode, B, draws, tail);
      } else {
        res = pop_round_hot(q, G, sub, gbase, gmask, live, rid, P, pn_fn, PK, cfg.compute_mode, B, draws, tail);
      }
    } else if (AGENT == OLS_AGENT_QLEARN) {
      EpsGreedySource<NA> src;
      src.C = C;
      src.a_star = a_star;
      src.nA = nA;
      res = pop_round<NV, RNG, PPL>(q, R, G, sub, gbase, gmask, live, rr, rid, src, pn_fn, PK, cfg.compute_mode, B, draws, n_pop,
                                    pcg_s, pcg_inc, tail, cfg.prefetch);
    } else {
      res = pop_round<NV, RNG, PPL>(q, R, G, sub, gbase, gmask, live, rr, rid, P, pn_fn, PK, cfg.compute_mode, B, draws, n_pop,
                                    pcg_s, pcg_inc, tail, cfg.prefetch);
    }
    // ---- learner update and episode bookkeeping for the groups whose round ended the step
    if (live && res != kPopContinue) {
      // ---------------- phase 1: reads of shared state, loads, arithmetic
      const int ep_index = C->ep_index, lc = C->ep_len_count, t_old = C->ep_t, n_acc_old = C->n_acc;
      const double G_old = C->ep_G, disc = C->ep_disc;
      const long long acc_idx = (long long)C->accepted0 + n_acc_old;
      const int b_done = B.b, cur_done = B.cur;  // the queue just popped from and its new cursor
      int new_status = -1;
      bool accepted = (res == OLS_RUNNING), q_write = false, close_ep = false, append_G = false;
      const int S = AOS ? cur_z : cur_s;
      const int budget = C->budget;
      const uint32_t meta = __float_as_uint(tail.w);
      const int A = ols_meta_action(meta);
      const int Z2 = __float_as_int(tail.y);
      const int32_t row_idx = __float_as_int(tail.z);
      const int done = ols_meta_terminal(meta);
      const double Rw = (double)tail.x;
      int S2 = 0;
      QT qn = (QT)0, td = (QT)0;
      float4 carry = make_float4(0.f, 0.f, 0.f, 0.f);  // QCARRY: the next step's policy row
      // AoS: Q[S, :] as it stands before this step's update; the entry of S is rewritten whole in phase 2
      float4 srow = make_float4(0.f, 0.f, 0.f, 0.f);
      if (QCARRY) {
        srow = cold_ld4(C->qrow);
      } else if (AOS) {
        if (S >= 0 && S < nS) {
          const AosEntry eS = load_aos(aos, S, cfg.state_epoch);
          srow = make_float4(eS.q[0], eS.q[1], eS.q[2], eS.q[3]);
        }
      }
      double G_new = G_old;
      int t_new = t_old;
      if (accepted) {
        S2 = (!AOS && q.s_next) ? __ldg(q.s_next + row_idx) : Z2;
        // the next state is known: fetch its queue head now, ahead of the learner's own memory traffic.  (The cursor
        // of the queue just popped is still the old one in memory; a self-loop patches it from the register.)
        AosEntry e2;
        e2.cursor = 0;
        if (AOS && S2 >= 0 && S2 < nS) e2 = load_aos(aos, S2, cfg.state_epoch);  // the next state's cursor AND Q row: one sector
        const bool next_ok = done ? false : load_bucket(q, R, C->h_rep, cur_tab, Z2, B, AOS, e2.cursor);
        if (next_ok && B.b == b_done) B.cur = cur_done;
        bucket_valid = next_ok;
        need_step = true;
        if (AGENT != OLS_AGENT_FIXED_PI) {
          if (S2 < 0 || S2 >= nS) {
            new_status = OLS_NO_QUEUE;
          } else {
            const double alpha = __ldg(cfg.alpha + ep_index);
            QT nrow[NA];
#pragma unroll
            for (int j = 0; j < NA; ++j) nrow[j] = (j < nA) ? (AOS ? (QT)e2.q[j & 3] : Qr[(int64_t)S2 * nA + j]) : (QT)-INFINITY;
            QT boot;
            if (AGENT == OLS_AGENT_QLEARN) {
              boot = nrow[0];  // Q[S_].max()
#pragma unroll
              for (int j = 1; j < NA; ++j) boot = (nrow[j] > boot) ? nrow[j] : boot;
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
            if (AOS) {
              qsa = (QT)((A == 0) ? srow.x : (A == 1) ? srow.y : (A == 2) ? srow.z : srow.w);
            } else if (AGENT == OLS_AGENT_QLEARN) {
              qsa = row[0];
#pragma unroll
              for (int j = 1; j < NA; ++j)
                if (j == A) qsa = row[j];
            } else {
              qsa = Qr[(int64_t)S * nA + A];
            }
            td = QMath<QT>::td((QT)Rw, (QT)cfg.gamma, boot, qsa);
            qn = QMath<QT>::upd(qsa, (QT)alpha, td);
            q_write = true;
            if (AGENT == OLS_AGENT_QLEARN) {
              // the bootstrap row becomes the next step's policy row (patched if S' == S)
              if (QCARRY) {
                const bool self = (S2 == S);
                carry = make_float4((float)((self && A == 0) ? qn : nrow[0]), (float)((self && A == 1) ? qn : nrow[1 % NA]),
                                    (float)((self && A == 2) ? qn : nrow[2 % NA]), (float)((self && A == 3) ? qn : nrow[3 % NA]));
              } else {
#pragma unroll
                for (int j = 0; j < NA; ++j) row[j] = (S2 == S && j == A) ? qn : nrow[j];
              }
              row_valid = true;
            }
          }
        }
        if (new_status < 0) {
          G_new = __dadd_rn(G_old, __dmul_rn(disc, Rw));  // G + gamma**t * R   (psrs.py:171)
          t_new = t_old + 1;
          if (done) {  // psrs.py:176-178 / 265-269
            close_ep = true;
            append_G = true;
          } else if (!next_ok) {
            new_status = OLS_NO_QUEUE;  // the next state has no queue: KeyError on the reference's next step (psrs.py:44)
          }
        }
      } else {
        // exhaustion in mid-episode (psrs.py:160-162): qlearn/expSARSA still append the partial return (:177-178),
        // evalMC appends only the length (:265-269)
        new_status = res;
        if (res == OLS_QUEUE_EMPTY) {
          close_ep = true;
          append_G = (AGENT != OLS_AGENT_FIXED_PI);
        }
      }
      __syncwarp(gm);
      // ---------------- phase 2: writes (every lane of the group writes the same values)
      // an accepted transition is committed unless the learner found its next observation out of range
      const bool commit = accepted && !(AGENT != OLS_AGENT_FIXED_PI && !q_write);
      if (b_done >= 0) {
        if (AOS) {
          // b_done == S (the observation id is the queue key): cursor and Q[S, A] go out as one 32-byte store
          if (commit && q_write) {
            srow.x = (A == 0) ? (float)qn : srow.x;
            srow.y = (A == 1) ? (float)qn : srow.y;
            srow.z = (A == 2) ? (float)qn : srow.z;
            srow.w = (A == 3) ? (float)qn : srow.w;
          }
          store_aos(aos, b_done, cur_done, cfg.state_epoch, srow);
        } else {
          cur_tab[b_done] = cur_done;
        }
      }
      if (commit) {
        if (q_write) {
          if (!AOS) Qr[(int64_t)S * nA + A] = qn;
          if (QCARRY) cold_st4(C->qrow, carry);
          if (sub == 0 && cfg.td_error && acc_idx < cfg.trace_cap) cfg.td_error[rr * cfg.trace_cap + acc_idx] = (double)td;
        }
        if (cfg.trace_row && sub == 0 && acc_idx < cfg.trace_cap) cfg.trace_row[rr * cfg.trace_cap + acc_idx] = row_idx;
        cur_z = Z2;
        cur_s = S2;
        C->n_acc = n_acc_old + 1;
        C->ep_G = G_new;
        C->ep_disc = __dmul_rn(disc, cfg.gamma);
        C->ep_t = t_new;
        if (budget > 0 && n_acc_old + 1 >= budget) live = false;  // budget: stays resumable
      }
      if (close_ep) {
        if (sub == 0) {
          if (append_G && ep_index < cfg.ep_cap) cfg.Gs[rr * (int64_t)cfg.ep_cap + ep_index] = G_new;
          if (lc < cfg.ep_cap) cfg.lengths[rr * (int64_t)cfg.ep_cap + lc] = t_new;
        }
        if (append_G) C->ep_index = ep_index + 1;
        C->ep_len_count = lc + 1;
        C->ep_open = 0;
      }
      if (new_status >= 0) {
        C->status = new_status;
        live = false;
      }
      __syncwarp(gm);
    }
  }
  // ---- write the staged tables back
  __syncwarp();
  if (!AOS && plan.cursor_bytes && exists) {
    const int32_t* sc = (const int32_t*)my_slot;
    int32_t* g_cur = R.cursor + rr * (int64_t)q.n_buckets;
    for (int i = sub; i < q.n_buckets; i += G) g_cur[i] = sc[i];
  }
  if (!AOS && AGENT != OLS_AGENT_FIXED_PI && plan.q_bytes && exists) {
    const QT* sq = (const QT*)(my_slot + plan.cursor_bytes);
    QT* g_Q = (QT*)cfg.Q + rr * (int64_t)nS * nA;
    for (int i = sub; i < nS * nA; i += G) g_Q[i] = sq[i];
  }
  if (exists && sub == 0) {
    R.status[r] = C->status;
    R.cur_z[r] = cur_z;
    R.cur_s[r] = AOS ? cur_z : cur_s;
    R.init_cursor[r] = C->init_cursor;
    const uint64_t draws0 = __ldcg(R.draws + r);
    R.draws[r] = draws;
    R.ties[r] = C->ties;
    R.pops[r] = __ldcg(R.pops + r) + ((RNG == RNGT_NONE) ? (unsigned long long)n_pop : (unsigned long long)(draws - draws0));
    R.accepted[r] = C->accepted0 + (unsigned long long)C->n_acc;
    if (RNG == RNGT_PCG64) {
      R.pcg_state[4 * r] = pcg_s.hi;
      R.pcg_state[4 * r + 1] = pcg_s.lo;
    }
    R.ep_index[r] = C->ep_index;
    R.ep_len_count[r] = C->ep_len_count;
    R.ep_t[r] = C->ep_t;
    R.ep_open[r] = C->ep_open;
    R.ep_G[r] = C->ep_G;
    R.ep_disc[r] = C->ep_disc;
  }
  if (cfg.slices > 1) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();  // release: every thread's state writes precede the token
      const int n_blk = gridDim.x / cfg.slices;
      *((volatile int*)cfg.sched + 1 + s_ticket % n_blk) = s_ticket / n_blk + 1;
    }
  }
}

// ------------------------------------------------------------------------------------------ per-NV launchers
struct RunLaunch {
  int grid;
  int threads;  // CTA size: kThreads, or OLS_G1_THREADS for the one-lane hot-record learner
  size_t smem;
  StagePlan plan;
  ols_philox_keys keys;
  cudaStream_t stream;
};

// the AoS-state variant exists for the production shape only: <= 4 actions, float32 Q, Philox or fed uniforms, 2 records/lane
template <int NV, int AGENT, typename QT>
inline cudaError_t launch_run_aos(const ols_queues_t& q, const ols_replicas_t& r, const ols_run_t& c, const RunLaunch& L) {
  if constexpr (NV == 1 && sizeof(QT) == 4 && AGENT != OLS_AGENT_FIXED_PI) {
    auto kf = run_kernel<NV, AGENT, QT, RNGT_FEED, 2, true>;
    auto kp = run_kernel<NV, AGENT, QT, RNGT_PHILOX, 2, true>;
    auto kp2 = run_kernel<NV, AGENT, QT, RNGT_PHILOX, 2, true, 2>;
    auto kp4 = run_kernel<NV, AGENT, QT, RNGT_PHILOX, 2, true, 4>;
    auto k = (r.rng_mode == OLS_RNG_FEED) ? kf : (c.lanes == 2 ? kp2 : c.lanes == 4 ? kp4 : kp);
    if (q.hot != nullptr && r.rng_mode == OLS_RNG_PHILOX) {  // hot records: 4 x 8 bytes per lane and round
      auto kh = run_kernel<NV, AGENT, QT, RNGT_PHILOX, 4, true, 0, true>;
      auto kh1 = run_kernel<NV, AGENT, QT, RNGT_PHILOX, 4, true, 1, true>;
      auto kh2 = run_kernel<NV, AGENT, QT, RNGT_PHILOX, 4, true, 2, true>;
      k = (c.lanes == 2) ? kh2 : (c.lanes == 1) ? kh1 : kh;
    }
    if (L.smem > 40 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
      if (e != cudaSuccess) return e;
    }
    k<<<L.grid, L.threads, L.smem, L.stream>>>(q, r, c, L.plan, L.keys);
    return cudaGetLastError();
  } else {
    return cudaErrorInvalidValue;
  }
}

template <int NV, int AGENT, typename QT>
inline cudaError_t launch_run_rng(const ols_queues_t& q, const ols_replicas_t& r, const ols_run_t& c, const RunLaunch& L,
                                  int ppl) {
  if (c.state_aos != nullptr) return launch_run_aos<NV, AGENT, QT>(q, r, c, L);
#define OLS_GO(RNG, PPL)                                                                                               \
  do {                                                                                                                 \
    auto k = run_kernel<NV, AGENT, QT, RNG, PPL>;                                                                      \
    if (L.smem > 40 * 1024) {                                                                                          \
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);               \
      if (e != cudaSuccess) return e;                                                                                  \
    }                                                                                                                  \
    k<<<L.grid, kThreads, L.smem, L.stream>>>(q, r, c, L.plan, L.keys);                                                        \
  } while (0)
  if constexpr (NV == 1 && AGENT == OLS_AGENT_FIXED_PI) {
    // fixed-policy evaluation over hot records (the streaming-reject regime reads 8 instead of 32 bytes per pop)
    if (q.hot != nullptr && r.rng_mode == OLS_RNG_PHILOX && c.reject_mode == OLS_REJECT_DEFAULT && (ppl == 2 || ppl == 4)) {
      auto kt = run_kernel<NV, AGENT, QT, RNGT_PHILOX, 4, false, 0, 2>;
      auto kp = run_kernel<NV, AGENT, QT, RNGT_PHILOX, 4, false, 0, 1>;
      auto k = (q.hot_kind == OLS_HOT_TAIL) ? kt : kp;
      if (L.smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
        if (e != cudaSuccess) return e;
      }
      k<<<L.grid, kThreads, L.smem, L.stream>>>(q, r, c, L.plan, L.keys);
      return cudaGetLastError();
    }
  }
  if (c.reject_mode == OLS_REJECT_ACCEPT_ALL) OLS_GO(RNGT_NONE, 1);
  else if (r.rng_mode == OLS_RNG_PCG64) OLS_GO(RNGT_PCG64, 1);
  else if (r.rng_mode == OLS_RNG_FEED) { if (ppl == 4) OLS_GO(RNGT_FEED, 4); else if (ppl == 2) OLS_GO(RNGT_FEED, 2); else OLS_GO(RNGT_FEED, 1); }
  else { if (ppl == 4) OLS_GO(RNGT_PHILOX, 4); else if (ppl == 2) OLS_GO(RNGT_PHILOX, 2); else OLS_GO(RNGT_PHILOX, 1); }
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
                                  const ols_step_out_t& out, int grid, cudaStream_t stream, const int32_t* act) {
#define OLS_GO(RNG) \
  step_kernel<NV, RNG><<<grid, kThreads, 0, stream>>>(q, r, mask, p_new, p_new_is_f64, p_stride, compute_mode, lanes, out, \
                                                      ols_philox_make_keys(r.seed), act)
  if (reject_mode == OLS_REJECT_ACCEPT_ALL) OLS_GO(RNGT_NONE);
  else if (r.rng_mode == OLS_RNG_PCG64) OLS_GO(RNGT_PCG64);
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
      int reject_mode, int compute_mode, int lanes, const ols_step_out_t &out, int grid, cudaStream_t stream, const int32_t *act
cudaError_t ols_launch_step_nv1(OLS_STEP_ARGS);
cudaError_t ols_launch_step_nv2(OLS_STEP_ARGS);
cudaError_t ols_launch_step_nv3(OLS_STEP_ARGS);
cudaError_t ols_launch_step_nv4(OLS_STEP_ARGS);
