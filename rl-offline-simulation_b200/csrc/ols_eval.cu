// The streaming evaluator: psrs.py:241-271 `evalMC_psrs` over psrs.py:32-57, ONE WARP per replica, for the rejection-heavy
// regime (BASELINE.json configs[4]: hundreds of rejected pops per accepted transition) on a log whose behaviour distribution
// is a function of the state (TAIL hot records + queue_dist, see ols_learner.cu / make_limits).  Same semantics, statement
// for statement, as run_kernel<1, FIXED_PI, float, PHILOX, 4, AOS = 0, GT = 0, HOT = 2> with lanes = 32 (the tests compare
// both with oracle/ols_oracle.c draw for draw); what differs is that everything but the per-record decision is
// warp-uniform here: no lane groups, no shared-memory replica block, no phased reads and writes — a pop round is
//     U x one 256-bit load (the lane's 32-byte sectors, four 8-byte records each)  +  U Philox4x32-7 blocks (the high words of its
//     draws; the words that belong to the neighbour's sector travel by shuffle)  +  per record { action -> reject limit (one
//     LDS.32), one compare }  +  one ballot per sector row
// for 128 U records: 1.7 warp-instructions per popped record at U = 2 (profiles/r04g_ncu_eval_kernel_c5.txt).
#include "ols_sim_kernels.cuh"

namespace ols_sim {

constexpr int kEvalThreads = 256;  // 8 replicas per CTA
#ifndef OLS_EVAL_MINB
#define OLS_EVAL_MINB 4
#endif
#ifndef OLS_EVAL_U
#define OLS_EVAL_U 2  // 32-byte sectors per lane and round (measured: 1 -> 6.5 ms, 2 -> 6.1 ms, 4 -> 7.4+ ms on C5)
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

template <int U>
__global__ void __launch_bounds__(kEvalThreads, OLS_EVAL_MINB) eval_kernel(ols_queues_t q, ols_replicas_t R, ols_run_t cfg, ols_philox_keys PK) {
  constexpr int NA = 4;
  __shared__ __align__(32) uint2 s_lim[kEvalThreads / 32][4];  // per replica and action: accept below .x / reject above .y (make_limits)
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
  const unsigned lim_y = (unsigned)__cvta_generic_to_shared(&s_lim[wib][0]) + 4u;  // + 8 a: the reject limit of action a (32-byte aligned row)
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
      // the step's limits (finish_policy + make_limits, same arithmetic), one action per lane: lane j & 3 holds action j
      const int j = lane & 3;
      const float pj = (j < nA) ? (float)pn_fn(j) : 0.f;
      const bool okj = (j >= nA) || pj == 0.f || (pj >= 9.094947e-13f && pj <= 1.0f);
      const float ipn = (pj == 0.f) ? __int_as_float(0x7f800000) : __frcp_rn(pj);  // +inf for pi = 0 and for padding entries
      const float pq = __ldg(q.queue_dist + 8 * (size_t)B.b + j);
      const bool tame = (__ldg((const int*)q.queue_dist + 8 * (size_t)B.b + 4) & 2) != 0;
      const bool ok = __all_sync(kFull, okj) && __any_sync(kFull, pj > 0.f) && tame;
      const float t = pq * ipn;
      float inv_m = fminf(t, __shfl_xor_sync(kFull, t, 1));   // (all t > 0 or +inf when ok: the order of the min is immaterial)
      inv_m = fminf(inv_m, __shfl_xor_sync(kFull, inv_m, 2));
      const float thr = __fdiv_rn(inv_m, t);
      const float lo = __fmul_rd(thr, 4294836224.0f), hi = __fmul_ru(thr, 4295098368.0f);
      const bool fine = ok && (thr == thr);
      const unsigned lo32 = __float2uint_rd(lo), hi32 = __float2uint_ru(hi);
      uint2 lim;
      lim.x = !fine ? 0u : (lo32 > 2u ? lo32 - 2u : 0u);
      lim.y = !fine ? 0xffffffffu : (hi32 < 0xfffffffdu ? hi32 + 2u : 0xffffffffu);
      __syncwarp();
      if (lane < 4) s_lim[wib][lane] = lim;
      __syncwarp();
    }
    // ---------------- env.step(p): pop rounds of 128 * U records until one is accepted (psrs.py:39-57)
    // A round examines one CONTIGUOUS piece of the queue's ring, and its lanes are aligned to the MEMORY: lane L reads the
    // 32-byte sectors (A + 4 (L + 32 u)) .. + 3, u < U, each with one 256-bit load (A = the round's first record rounded down to
    // a sector).  The draws are aligned differently (record k of the round consumes draw `draws + k`, stream block
    // (draws + k) >> 2), so a lane computes the block that starts inside its sector and takes the rest from its neighbour.
    uint2 won = make_uint2(0u, 0u);
    bool empty = false;
    for (;;) {
      const int remaining = B.len - B.cur;
      if (remaining <= 0) { empty = true; break; }  // psrs.py:44-45
      uint32_t pos0 = B.start + (uint32_t)B.cur;    // ring position of the next record
      if (pos0 >= (uint32_t)B.len) pos0 -= (uint32_t)B.len;
      const uint32_t P = (uint32_t)B.lo + pos0;     // its index in the hot array
      const int mis = (int)(P & 3u);
      const long long D1 = (long long)draws - mis;  // the draw the sector's first record would consume (>= -3)
      const int s = (int)(D1 & 3);                  // where the sectors start inside the stream blocks
      // up to the ring's wrap, and as far as the lanes' blocks reach (the last lane has no neighbour)
      const int n_eval = min(min(remaining, B.len - (int)pos0), 128 * U - mis - s);
      const uint2* const base = hot + (size_t)(P - (uint32_t)mis);
      uint2 h[U][4];
      ols_u4 o[U];
      bool sv[U];  // the lane's sector holds records of this round
      const bool all_in = (long long)(P - (uint32_t)mis) + 128 * U <= q.n_rows;  // (warp-uniform) no sector runs past the array's end
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int g = 4 * (lane + 32 * u);  // the sector's first record, relative to A
        const int k0 = g - mis;             // ... relative to the round's first record
        sv[u] = k0 + 3 >= 0 && k0 < n_eval;
#pragma unroll
        for (int i = 0; i < 4; ++i) h[u][i] = make_uint2(0u, 0u);
        if (sv[u]) {
          if (all_in) {
            asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(h[u][0].x), "=r"(h[u][0].y), "=r"(h[u][1].x), "=r"(h[u][1].y), "=r"(h[u][2].x), "=r"(h[u][2].y),
                           "=r"(h[u][3].x), "=r"(h[u][3].y)
                         : "l"(base + g));
          } else {  // near the array's end: record by record
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if ((unsigned)(k0 + i) < (unsigned)n_eval) h[u][i] = __ldg(base + g + i);
          }
        }
      }
      if (pf_rounds > 0) {
        // the reject chains of this regime are hundreds of records long: pull the span a later round will examine toward L2
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int gp = 4 * (lane + 32 * u) + pf_rounds * (128 * U - 4);
          if (gp - mis < min(remaining, B.len - (int)pos0)) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + gp));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        o[u] = ols_stream_block(PK, rid, 0u, (uint64_t)((D1 >> 2) + (long long)(lane + 32 * u)));
      int mine[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        uint32_t w0 = o[u].x, w1 = o[u].y, w2 = o[u].z, w3 = o[u].w;
        if (s != 0) {  // warp-uniform
          // the neighbour's block: lane L + 1's of the same u; for the last lane, lane 0's of the next u
          uint32_t n0, n1, n2;
          if (u + 1 < U) {
            const ols_u4 nx = o[u + 1 < U ? u + 1 : u];
            n0 = __shfl_sync(kFull, lane == 0 ? nx.x : o[u].x, (lane + 1) & 31);
            n1 = __shfl_sync(kFull, lane == 0 ? nx.y : o[u].y, (lane + 1) & 31);
            n2 = __shfl_sync(kFull, lane == 0 ? nx.z : o[u].z, (lane + 1) & 31);
          } else {
            n0 = __shfl_down_sync(kFull, o[u].x, 1);
            n1 = __shfl_down_sync(kFull, o[u].y, 1);
            n2 = __shfl_down_sync(kFull, o[u].z, 1);
          }
          if (s == 1) { w0 = o[u].y; w1 = o[u].z; w2 = o[u].w; w3 = n0; }
          else if (s == 2) { w0 = o[u].z; w1 = o[u].w; w2 = n0; w3 = n1; }
          else { w0 = o[u].w; w1 = n0; w2 = n1; w3 = n2; }
        }
        const uint32_t wv[4] = {w0, w1, w2, w3};
        const int k0 = 4 * (lane + 32 * u) - mis;
        // the common case first: is any of the four slots NOT plainly rejected?  One 4-byte shared-memory load (its address
        // one shift and one and-or away from the record) and one compare each; slots outside the round are sorted out below.
        uint32_t ly[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ly[i]) : "r"(lim_y | ((h[u][i].y >> 26) & 0x18u)));
        const bool some = sv[u] & ((wv[0] <= ly[0]) | (wv[1] <= ly[1]) | (wv[2] <= ly[2]) | (wv[3] <= ly[3]));
        int m = 4;
        if (some) {
          // 2 bits per slot: 0 reject, 1 accept, 2 undecided (inside the band: 6e-5 of the pops)
          unsigned decs = 0, acts = 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const unsigned a = (h[u][i].y >> 29) & 3u;
            const uint2 lm = s_lim[wib][a];
            const uint32_t hi = wv[i];
            const unsigned dec = ((unsigned)(k0 + i) >= (unsigned)n_eval || hi > lm.y) ? 0u : (hi < lm.x) ? 1u : 2u;
            decs |= dec << (2 * i);
            acts |= a << (2 * i);
          }
          m = decs ? (__ffs(decs) - 1) >> 1 : 4;  // the lane's first slot that is not a plain reject
          if (m < 4 && ((decs >> (2 * m)) & 3u) == 2u) {
            const void* pi_row = pi_is_f64 ? (const void*)((const double*)pi_tab + off) : (const void*)((const float*)pi_tab + off);
            m = resolve_slots(q.queue_dist + 8 * (size_t)B.b, pi_row, pi_is_f64, ols_keys_seed(PK), rid, draws + (uint64_t)(int64_t)k0,
                              decs, acts, nA, cfg.compute_mode);
          }
        }
        mine[u] = m;
      }
      int first_acc = -1;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (first_acc >= 0) break;
        const unsigned any = __ballot_sync(kFull, mine[u] < 4);
        if (any) {
          const int w = __ffs(any) - 1;
          const int wi = __shfl_sync(kFull, mine[u], w);
          const int m = mine[u];
          const uint2 mw = (m == 0) ? h[u][0] : (m == 1) ? h[u][1] : (m == 2) ? h[u][2] : h[u][3];
          won.x = __shfl_sync(kFull, mw.x, w);
          won.y = __shfl_sync(kFull, mw.y, w);
          first_acc = 4 * (w + 32 * u) + wi - mis;
        }
      }
      if (first_acc >= 0) {
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
  ols_sim::eval_kernel<OLS_EVAL_U><<<grid, ols_sim::kEvalThreads, 0, stream>>>(q, r, c, ols_philox_make_keys(r.seed));
  ols_note_launch(1);
  return cudaGetLastError();
}
