// The streaming evaluator: psrs.py:241-271 `evalMC_psrs` over psrs.py:32-57, ONE WARP per replica, for the rejection-heavy
// regime (BASELINE.json configs[4]: hundreds of rejected pops per accepted transition) on a log whose behaviour distribution
// is a function of the state (TAIL hot records + queue_dist, see ols_learner.cu / make_limits).  Same semantics, statement
// for statement, as run_kernel<1, FIXED_PI, float, PHILOX, 4, AOS = 0, GT = 0, HOT = 2> with lanes = 32 (the tests compare
// both with oracle/ols_oracle.c draw for draw); what differs is that everything but the per-record decision is
// warp-uniform here: no lane groups, no shared-memory replica block, no phased reads and writes — a pop round is
//     4 x LDG.64 (the lane's four consecutive 8-byte records)  +  one Philox4x32-7 block (the high words of its four draws)
//     +  4 x { action -> limits (one LDS.64), two integer compares }  +  one ballot / two shuffles
// for 128 records, i.e. about one warp-instruction per popped record: the loop is bound by the 8 bytes per record it reads.
#include "ols_sim_kernels.cuh"

namespace ols_sim {

constexpr int kEvalThreads = 256;  // 8 replicas per CTA
#ifndef OLS_EVAL_MINB
#define OLS_EVAL_MINB 4
#endif

// The cold end of a lane's four slots: those whose high word fell inside the band are decided by the exact rule psrs.py:53-57 on
// the queue's distribution and the policy row pi[S], in queue order; returns the lane's first accepting slot (4 = none).
__device__ __noinline__ int resolve_slots(const float* dist, const void* pi_row, int pi_is_f64, uint64_t seed, uint64_t rid, uint64_t d0,
                                          unsigned decs, unsigned acts, int nA, int compute_mode) {
  for (int i = 0; i < 4; ++i) {
    unsigned dec = (decs >> (2 * i)) & 3u;
    if (dec == 2u) {
      const double ue = ols_u53_to_unit(ols_stream_u53_seed(seed, rid, 0u, d0 + (uint64_t)i));
      auto pn_fn = [=](int j) -> double {
        return pi_is_f64 ? __ldg((const double*)pi_row + j) : (double)__ldg((const float*)pi_row + j);
      };
      dec = accept_slow(pn_fn, dist, nullptr, nA, (int)((acts >> (2 * i)) & 3u), ue, compute_mode) ? 1u : 0u;
    }
    if (dec == 1u) return i;
  }
  return 4;
}

__global__ void __launch_bounds__(kEvalThreads, OLS_EVAL_MINB) eval_kernel(ols_queues_t q, ols_replicas_t R, ols_run_t cfg, ols_philox_keys PK) {
  constexpr int NA = 4;
  __shared__ uint2 s_lim[kEvalThreads / 32][4];  // per replica and action: accept below .x / reject above .y (make_limits)
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * (kEvalThreads / 32) + wib;
  if (r >= R.n_replicas) return;  // warp-uniform
  const int nA = q.n_actions, nS = cfg.n_states;
  const uint64_t rid = (uint64_t)(R.replica_id0 + r);
  const uint32_t h_rep = R.ring_mode ? ols_ring_replica_hash(R.ring_seed, rid) : 0u;
  const uint2* const hot = (const uint2*)q.hot;
  int status = __ldcg(R.status + r);
  if (!(status == OLS_IDLE || status == OLS_RUNNING)) return;  // terminated replicas stay terminated
  int cur_z = __ldcg(R.cur_z + r), init_cursor = __ldcg(R.init_cursor + r);
  int ep_index = __ldcg(R.ep_index + r), lc = __ldcg(R.ep_len_count + r), ep_t = __ldcg(R.ep_t + r), ep_open = __ldcg(R.ep_open + r);
  double G = __ldcg(R.ep_G + r), disc = __ldcg(R.ep_disc + r);
  uint64_t draws = __ldcg(R.draws + r);
  int n_acc = 0;  // (a launch adds < 2^31 transitions per replica)
  int32_t* const cur_tab = R.cursor + r * (int64_t)q.n_buckets;
  Bucket B;
  B.b = -1;
  B.lo = B.len = B.cur = 0;
  B.start = 0;
  bool bucket_valid = false;
  const int pf_rounds = cfg.prefetch;
  for (;;) {
    // ---------------- episode start and the policy of the current state (psrs.py:246-255)
    if (!ep_open) {
      if (ep_index >= cfg.n_episodes) { status = OLS_DONE; break; }
      int cs_unused = 0;
      if (!reset_one(q, R, h_rep, init_cursor, cur_z, cs_unused)) { status = OLS_INIT_EMPTY; break; }  // no partial episode is appended
      status = OLS_RUNNING;
      ep_open = 1; ep_t = 0; G = 0.0; disc = 1.0;
      bucket_valid = false;
    }
    if (cur_z < 0 || cur_z >= nS) { status = OLS_NO_QUEUE; break; }                                     // IndexError on pi[S]
    if (!bucket_valid && !load_bucket(q, R, h_rep, cur_tab, cur_z, B)) { status = OLS_NO_QUEUE; break; }  // KeyError (psrs.py:44)
    const void* pi_tab = cfg.pi;
    const int pi_is_f64 = cfg.pi_is_f64;
    const int64_t off = (int64_t)cur_z * nA;
    auto pn_fn = [=](int j) -> double {
      return pi_is_f64 ? __ldg((const double*)pi_tab + off + j) : (double)__ldg((const float*)pi_tab + off + j);
    };
    {
      Policy<NA> P;
      finish_policy<NA>(P, pn_fn, nA);
      const float4 qd = __ldg((const float4*)q.queue_dist + 2 * (size_t)B.b);
      const bool tame = (__ldg((const int*)q.queue_dist + 8 * (size_t)B.b + 4) & 2) != 0;
      const float ipn4[4] = {P.ipn[0], P.ipn[1], P.ipn[2], P.ipn[3]};
      uint2 lim[4];
      make_limits(qd, P.ok && tame, ipn4, lim);
      __syncwarp();
      if (lane < 4) s_lim[wib][lane] = lane == 0 ? lim[0] : lane == 1 ? lim[1] : lane == 2 ? lim[2] : lim[3];
      __syncwarp();
    }
    // ---------------- env.step(p): pop rounds of 128 records until one is accepted (psrs.py:39-57)
    uint2 won = make_uint2(0u, 0u);
    int first_acc = -1;
    bool empty = false;
    for (;;) {
      const int remaining = B.len - B.cur;
      if (remaining <= 0) { empty = true; break; }  // psrs.py:44-45
      const int align = (int)(draws & 3);           // rounds are aligned to four draws: one stream block per lane
      const int n_eval = min(128 - align, remaining);
      const int k0 = 4 * lane - align;
      uint32_t pos = B.start + (uint32_t)(B.cur + k0);  // (k0 = -align..-1 only in lane 0: masked slots)
      uint2 h[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t p = pos + (uint32_t)i;
        if (p >= (uint32_t)B.len) p -= (uint32_t)B.len;
        h[i] = ((unsigned)(k0 + i) < (unsigned)n_eval) ? __ldg(hot + (size_t)(uint32_t)(B.lo + (int32_t)p)) : make_uint2(0u, 0u);
      }
      if (pf_rounds > 0) {
        // the reject chains of this regime are hundreds of records long: pull the span a later round will examine toward L2
        const int kp = k0 + pf_rounds * 128;
        if (kp + 3 < remaining) {
          uint32_t pp = B.start + (uint32_t)(B.cur + kp);
          if (pp >= (uint32_t)B.len) pp -= (uint32_t)B.len;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(hot + (size_t)(uint32_t)(B.lo + (int32_t)pp)));
        }
      }
      const ols_u4 o = ols_stream_block(PK, rid, 0u, ((draws - (uint64_t)align) >> 2) + (uint64_t)lane);
      // 2 bits per slot: 0 reject, 1 accept, 2 undecided (inside the band: 6e-5 of the pops)
      unsigned decs = 0, acts = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const unsigned a = (h[i].y >> 29) & 3u;
        const uint2 lm = s_lim[wib][a];
        const uint32_t hi = ols_u4_word(o, i);
        const unsigned dec = ((unsigned)(k0 + i) >= (unsigned)n_eval || hi > lm.y) ? 0u : (hi < lm.x) ? 1u : 2u;
        decs |= dec << (2 * i);
        acts |= a << (2 * i);
      }
      int mine = decs ? (__ffs(decs) - 1) >> 1 : 4;  // the lane's first slot that is not a plain reject
      if (mine < 4 && ((decs >> (2 * mine)) & 3u) == 2u) {
        const void* pi_row = pi_is_f64 ? (const void*)((const double*)pi_tab + off) : (const void*)((const float*)pi_tab + off);
        mine = resolve_slots(q.queue_dist + 8 * (size_t)B.b, pi_row, pi_is_f64, ols_keys_seed(PK), rid, draws + (uint64_t)(int64_t)k0, decs,
                             acts, nA, cfg.compute_mode);
      }
      const unsigned any = __ballot_sync(kFull, mine < 4);
      if (any) {
        const int w = __ffs(any) - 1;
        const int wi = __shfl_sync(kFull, mine, w);
        const uint2 mw = (mine == 0) ? h[0] : (mine == 1) ? h[1] : (mine == 2) ? h[2] : h[3];
        won.x = __shfl_sync(kFull, mw.x, w);
        won.y = __shfl_sync(kFull, mw.y, w);
        first_acc = 4 * w + wi - align;
        B.cur += first_acc + 1;
        draws += (uint64_t)(first_acc + 1);
        break;
      }
      B.cur += n_eval;
      draws += (uint64_t)n_eval;
    }
    // ---------------- the outcome of the step (psrs.py:256-269)
    const int b_done = B.b, cur_done = B.cur;
    if (lane == 0) cur_tab[b_done] = cur_done;
    if (empty) {  // exhaustion in mid-episode: evalMC appends the length only (psrs.py:265-269)
      status = OLS_QUEUE_EMPTY;
      if (lane == 0 && lc < cfg.ep_cap) cfg.lengths[r * (int64_t)cfg.ep_cap + lc] = ep_t;
      lc += 1;
      ep_open = 0;
      break;
    }
    const int Z2 = ((int)(won.y << 3)) >> 3;
    const int done = (int)(won.y >> 31);
    const double Rw = (double)__uint_as_float(won.x);
    if (cfg.trace_row && lane == 0) {
      const long long acc_idx = (long long)__ldcg(R.accepted + r) + n_acc;
      if (acc_idx < cfg.trace_cap) {
        // the row index lives in the full record only: a parity tap, off in production
        uint32_t p = B.start + (uint32_t)(cur_done - 1);
        if (p >= (uint32_t)B.len) p -= (uint32_t)B.len;
        cfg.trace_row[r * cfg.trace_cap + acc_idx] = __ldg((const int32_t*)((const char*)q.records + (size_t)(uint32_t)(B.lo + (int32_t)p) * 32) + 6);
      }
    }
    __syncwarp();  // lane 0's cursor store precedes the next bucket's cursor load (a self-loop reads it back)
    bool next_ok = false;
    if (!done) {
      next_ok = load_bucket(q, R, h_rep, cur_tab, Z2, B);
      if (next_ok && B.b == b_done) B.cur = cur_done;
    }
    bucket_valid = next_ok;
    cur_z = Z2;
    n_acc += 1;
    G = __dadd_rn(G, __dmul_rn(disc, Rw));  // G + gamma**t * R   (psrs.py:262)
    disc = __dmul_rn(disc, cfg.gamma);
    ep_t += 1;
    if (done) {  // psrs.py:265-269
      if (lane == 0) {
        if (ep_index < cfg.ep_cap) cfg.Gs[r * (int64_t)cfg.ep_cap + ep_index] = G;
        if (lc < cfg.ep_cap) cfg.lengths[r * (int64_t)cfg.ep_cap + lc] = ep_t;
      }
      ep_index += 1;
      lc += 1;
      ep_open = 0;
    } else if (!next_ok) {
      status = OLS_NO_QUEUE;  // the next state has no queue: KeyError on the reference's next step (psrs.py:44)
      break;
    }
    if (cfg.max_steps > 0 && (long long)n_acc >= cfg.max_steps) break;  // budget: stays RUNNING, resumable
  }
  if (lane == 0) {
    R.status[r] = status;
    R.cur_z[r] = cur_z;
    R.cur_s[r] = cur_z;
    R.init_cursor[r] = init_cursor;
    R.pops[r] = __ldcg(R.pops + r) + (unsigned long long)(draws - __ldcg(R.draws + r));
    R.draws[r] = draws;
    R.accepted[r] = __ldcg(R.accepted + r) + (unsigned long long)n_acc;
    R.ep_index[r] = ep_index;
    R.ep_len_count[r] = lc;
    R.ep_t[r] = ep_t;
    R.ep_open[r] = ep_open;
    R.ep_G[r] = G;
    R.ep_disc[r] = disc;
  }
}

}  // namespace ols_sim

cudaError_t ols_launch_eval(const ols_queues_t& q, const ols_replicas_t& r, const ols_run_t& c, cudaStream_t stream) {
  const int per = ols_sim::kEvalThreads / 32;
  const int grid = (int)((r.n_replicas + per - 1) / per);
  ols_sim::eval_kernel<<<grid, ols_sim::kEvalThreads, 0, stream>>>(q, r, c, ols_philox_make_keys(r.seed));
  ols_note_launch(1);
  return cudaGetLastError();
}
