// The production learner: psrs.py:119-185 `qlearn_psrs` + agents/tabular.py:4-5,24-32 epsilon-greedy, ONE THREAD per
// replica, for the shape the benchmark names (BASELINE.json configs[3]): float32 Q, <= 4 actions, the agent's observation
// id is the state id, AoS replica state {cursor, epoch, Q row} per (replica, state), Philox streams (protocol v2), and a log
// whose behaviour distribution is a function of the state: every queue holds ONE logged distribution (checked bit for bit
// when the log is bucketed: ols_queues_t.queue_dist / hot_kind = OLS_HOT_TAIL).  Same semantics, statement for statement,
// as run_kernel<1, QLEARN, float, PHILOX, 4, AOS, 1, HOT> in ols_sim_kernels.cuh (the tests compare both with
// oracle/ols_oracle.c draw for draw); what differs is the schedule:
//
//   * the accept rule psrs.py:53-57 depends on a record only through its ACTION once the queue's distribution is known, so
//     the per-record work of the float32 pre-filter (four products, the min, the reciprocal) moves out of the pop loop: per
//     STEP and action, the band around the estimated threshold as two 32-bit limits for the high word of the uniform
//     (make_limits; same error budget as accept_filter); per POP, one shared-memory load and two integer compares, and the
//     exact rule inside the band.
//   * 8-byte TAIL hot records { reward; next state | action | terminal }: everything a pop needs, rejected or accepted.
//     The 32-byte record is no longer on the path (only the optional row trace reads it), so a transition's dependent DRAM
//     chain is   hot span -> AoS entry of the next state   instead of   hot span -> record -> AoS entry.
//   * one stream block per four draws, carried from round to round (a round examines four VALID records whatever the
//     draw counter's alignment; the block holding the next draw is kept in shared memory), ~0.75 seven-round Philox blocks
//     per round instead of two ten-round ones.
//   * batched updates: a lane that has an outcome (accepted record / exhausted queue) parks it — its next state's AoS
//     entry is prefetched into L2 at that moment — and the warp runs the update + next-action block only once
//     `batch_lanes` lanes hold one, so that block's instructions serve more replicas and the parked lanes' DRAM latency
//     overlaps the other lanes' pop rounds.
//   * per-replica cold registers live in shared memory as structure-of-arrays (bank-conflict free; one owner thread per
//     slot, so no phasing / __syncwarp between reads and writes).
#include "ols_sim_kernels.cuh"

namespace ols_sim {

constexpr int kLT = OLS_G1_THREADS;  // threads (= replicas) per CTA

struct LearnerCold {
  double ep_G[kLT], ep_disc[kLT];
  unsigned long long ties[kLT];
  float4 qrow[kLT];   // Q[S, :] of the current state (the bootstrap row of the previous step)
  float4 qdist[kLT];  // the logged distribution of the current state's queue
  uint2 lim[4][kLT];  // per action: accept below .x, reject above .y (high word of the uniform), see make_limits
  uint2 park[kLT];    // a parked outcome: reward bits, next state | action | terminal
  uint4 blk[kLT];     // the rejection stream's block that holds draw `draws` (block index draws >> 2)
  uint4 tblk[kLT];    // the tie-break stream's block that holds tie `ties` (block index ties >> 2)
  int ep_index[kLT], ep_len_count[kLT], ep_t[kLT], ep_open[kLT], init_cursor[kLT], status[kLT], n_acc[kLT], pn_ok[kLT],
      eps_episode[kLT], qtame[kLT];
  float ipn_base[kLT], ipn_top[kLT];
  unsigned h_rep[kLT];
};

// The cold end of a pop round: slots whose high word fell inside the band are decided by the exact rule psrs.py:53-57 on
// the queue's distribution, in queue order; returns the first accepting slot (4 = none).
__device__ __noinline__ int resolve_round(const float* dist, const double* eps_ptr, uint64_t seed, uint64_t rid, uint64_t draws,
                                          unsigned decs, unsigned acts, int a_star, int nA, int compute_mode) {
  for (int i = 0; i < 4; ++i) {
    unsigned dec = (decs >> (2 * i)) & 3u;
    if (dec == 2u) {
      const double ue = ols_u53_to_unit(ols_stream_u53_seed(seed, rid, 0u, draws + (uint64_t)i));
      auto pn_fn = [=](int j) -> double {
        const double eps = __ldg(eps_ptr);
        const double base = __ddiv_rn(eps, (double)nA);           // ones * eps / nA          (tabular.py:27)
        const double top = __dadd_rn(__dsub_rn(1.0, eps), base);  // 1 - eps + eps / nA       (tabular.py:30)
        return (j == a_star) ? top : base;
      };
      dec = accept_slow(pn_fn, dist, nullptr, nA, (int)((acts >> (2 * i)) & 3u), ue, compute_mode) ? 1u : 0u;
    }
    if (dec == 1u) return i;
  }
  return 4;
}

__global__ void __launch_bounds__(kLT, OLS_LEARNER_MINB) learner_kernel(ols_queues_t q, ols_replicas_t R, ols_run_t cfg, ols_philox_keys PK) {
  constexpr int NA = 4;
  __shared__ LearnerCold C;
  __shared__ int s_ticket;
  const int nA = q.n_actions;
  const int nS = cfg.n_states;
  const int tid = threadIdx.x;
  // ---- the CTA's unit of work: a replica block, or one time slice of one (see run_kernel)
  int blk = blockIdx.x;
  long long budget64 = cfg.max_steps;
  if (cfg.slices > 1) {
    if (tid == 0) s_ticket = atomicAdd(cfg.sched, 1);
    __syncthreads();
    const int n_blk = gridDim.x / cfg.slices;
    blk = s_ticket % n_blk;
    const int slice = s_ticket / n_blk;
    const long long per = (cfg.max_steps + cfg.slices - 1) / cfg.slices;
    budget64 = min(per, cfg.max_steps - (long long)slice * per);
    if (tid == 0) {
      const volatile int* turn = (const volatile int*)cfg.sched + 1 + blk;
      while (*turn != slice) __nanosleep(256);
      __threadfence();  // acquire: the previous slice's replica state (written on another SM) is visible below
    }
    __syncthreads();
  }
  const bool no_budget = cfg.slices > 1 && budget64 <= 0;
  const int64_t r = (int64_t)blk * kLT + tid;
  const bool exists = r < R.n_replicas;
  const int64_t rr = exists ? r : 0;
  const uint64_t rid = (uint64_t)(R.replica_id0 + rr);
  unsigned char* const aos = (unsigned char*)cfg.state_aos + (size_t)rr * (size_t)nS * 32;
  const uint2* const hot = (const uint2*)q.hot;

  int cur_z = 0;
  uint64_t draws = 0;
  bool live = false;
  C.n_acc[tid] = 0;
  C.eps_episode[tid] = -1;
  const int budget = (int)max(0ll, min(budget64, 0x7fffffffll));  // accepted transitions this launch / slice may add (0: no limit)
  C.h_rep[tid] = R.ring_mode ? ols_ring_replica_hash(R.ring_seed, rid) : 0u;
  C.status[tid] = OLS_IDLE;
  if (exists) {
    const int status = __ldcg(R.status + r);
    cur_z = __ldcg(R.cur_z + r);
    draws = __ldcg(R.draws + r);
    C.status[tid] = status;
    C.init_cursor[tid] = __ldcg(R.init_cursor + r);
    C.ties[tid] = __ldcg(R.ties + r);
    C.ep_index[tid] = __ldcg(R.ep_index + r);
    C.ep_len_count[tid] = __ldcg(R.ep_len_count + r);
    C.ep_t[tid] = __ldcg(R.ep_t + r);
    C.ep_open[tid] = __ldcg(R.ep_open + r);
    C.ep_G[tid] = __ldcg(R.ep_G + r);
    C.ep_disc[tid] = __ldcg(R.ep_disc + r);
    live = (status == OLS_IDLE || status == OLS_RUNNING) && !no_budget;  // terminated replicas stay terminated
    const ols_u4 o = ols_stream_block(PK, rid, 0u, draws >> 2);
    C.blk[tid] = make_uint4(o.x, o.y, o.z, o.w);
    if (R.tie_mode != OLS_RNG_FEED) {
      const ols_u4 t = ols_stream_block(PK, rid, 1u, C.ties[tid] >> 2);
      C.tblk[tid] = make_uint4(t.x, t.y, t.z, t.w);
    }
  }

  // the queue of the current state
  int b_cur = -1, b_lo = 0, b_len = 0, b_pos = 0;  // bucket, CSR range, records consumed by this replica
  uint32_t b_start = 0;                            // ring offset
  bool bucket_valid = false, row_valid = false;
  int a_star = 0;
  // a parked outcome: what the round that ended the step found
  bool pend = false;
  int p_res = 0;           // OLS_RUNNING = a record was accepted, else OLS_QUEUE_EMPTY
  uint32_t p_slot = 0;     // queue slot of the accepted record (for the row trace)
  bool need_step = live;

  const int batch = cfg.batch_lanes > 0 ? min(cfg.batch_lanes, 32) : 20;

  // Stage the next state of a parked, accepted transition: its AoS entry (cursor + Q row), its queue and (to shared memory, for
  // the next step block) that queue's distribution; the hot span the next step will pop first is pulled toward L2.  Runs right before
  // the update (the entry was prefetched at the accept, so for lanes parked a round or more the loads hit L2).  All loads are
  // issued before any is consumed.  Tried and dropped: staging one loop iteration after the accept, while the lane is still
  // parked (the extra divergent block costs more issue slots than the hidden latency buys: 165 ms against 150 ms), and
  // prefetching the span a rejected round would continue with (173 ms).
  auto stage_next = [&](uint32_t zw, float4& nq, int4& nb, unsigned& nstart) {
    const int Z2 = ((int)(zw << 3)) >> 3;
    const bool inr = (Z2 >= 0 && Z2 < nS), want_q = inr && !(zw >> 31);
    float4 ea = make_float4(0.f, 0.f, 0.f, 0.f), eb = ea, qd = ea;
    int lo = 0, hi = 0, qf = 0;
    if (inr)
      asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=f"(ea.x), "=f"(ea.y), "=f"(ea.z), "=f"(ea.w), "=f"(eb.x), "=f"(eb.y), "=f"(eb.z), "=f"(eb.w)
                   : "l"(aos + (size_t)Z2 * 32));
    if (want_q) {
      lo = __ldg(q.offsets + Z2);
      hi = __ldg(q.offsets + Z2 + 1);
      qd = __ldg((const float4*)q.queue_dist + 2 * (size_t)Z2);
      qf = __ldg((const int*)q.queue_dist + 8 * (size_t)Z2 + 4);
    }
    const bool fresh = (__float_as_int(ea.y) == cfg.state_epoch);  // stale entries read as cursor 0 / Q 0
    const int cur = fresh ? __float_as_int(ea.x) : 0;
    nq = fresh ? eb : make_float4(0.f, 0.f, 0.f, 0.f);
    const int len = hi - lo;
    const bool next_ok = want_q && len > 0;   // psrs.py:44 — a state never logged as a from-state has no queue
    unsigned start = 0;
    if (next_ok) {
      if (R.ring_mode) start = ols_ring_start(C.h_rep[tid], (uint32_t)Z2, (uint32_t)len);
      C.qdist[tid] = qd;
      C.qtame[tid] = qf & 2;
      uint32_t pos = start + (uint32_t)cur;
      if (pos >= (uint32_t)len) pos -= (uint32_t)len;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(hot + (size_t)(uint32_t)(lo + (int)pos)));
    }
    nb = make_int4(next_ok ? Z2 : -1, next_ok ? lo : 0, next_ok ? len : 0, next_ok ? cur : 0);
    nstart = start;
  };

  while (__any_sync(kFull, live)) {
    // =========================================================== the step block: episode start + the agent's action choice
    if (live && need_step) {
      need_step = false;
      const int ep_index0 = C.ep_index[tid];
      int new_status = -1;
      float4 prow = make_float4(0.f, 0.f, 0.f, 0.f);
      bool fresh = false;
      // ---- episode start: psrs.py:149-156
      if (!C.ep_open[tid]) {
        int ic = C.init_cursor[tid], cs_unused = 0;
        if (ep_index0 >= cfg.n_episodes) {
          new_status = OLS_DONE;
        } else if (!reset_one(q, R, C.h_rep[tid], ic, cur_z, cs_unused)) {
          new_status = OLS_INIT_EMPTY;  // terminate without appending (psrs.py:153-155)
        } else {
          row_valid = false;
          bucket_valid = false;
          C.init_cursor[tid] = ic;
          C.status[tid] = OLS_RUNNING;
          C.ep_open[tid] = 1;
          C.ep_t[tid] = 0;
          C.ep_G[tid] = 0.0;
          C.ep_disc[tid] = 1.0;
        }
      }
      if (new_status < 0 && (cur_z < 0 || cur_z >= nS)) new_status = OLS_NO_QUEUE;  // IndexError on Q[S]
      if (new_status < 0 && !bucket_valid) {
        // one 256-bit load brings the state's cursor and Q row
        const AosEntry e = load_aos(aos, cur_z, cfg.state_epoch);
        Bucket B;
        if (!load_bucket(q, R, C.h_rep[tid], nullptr, cur_z, B, true, e.cursor)) new_status = OLS_NO_QUEUE;  // KeyError (psrs.py:44)
        b_cur = B.b; b_lo = B.lo; b_len = B.len; b_pos = B.cur; b_start = B.start;
        prow = make_float4(e.q[0], nA > 1 ? e.q[1] : -INFINITY, nA > 2 ? e.q[2] : -INFINITY, nA > 3 ? e.q[3] : -INFINITY);
        fresh = true;
        if (new_status < 0) {
          C.qdist[tid] = __ldg((const float4*)q.queue_dist + 2 * (size_t)b_cur);
          C.qtame[tid] = __ldg((const int*)q.queue_dist + 8 * (size_t)b_cur + 4) & 2;
        }
      }
      if (new_status < 0) {
        // tabular.py:24-32 epsilon_greedy_policy on the single row Q[[S], :] (psrs.py:158)
        if (!fresh) prow = C.qrow[tid];
        const float mx = fmaxf(fmaxf(prow.x, prow.y), fmaxf(prow.z, prow.w));  // no NaN in a Q row that started finite
        unsigned cand = (prow.x == mx ? 1u : 0u) | (prow.y == mx ? 2u : 0u) | (prow.z == mx ? 4u : 0u) | (prow.w == mx ? 8u : 0u);
        const int ncand = __popc(cand);
        if (ncand > 1) {  // tabular.py:4-5 _random_argmax: a draw is consumed only on ties
          const unsigned long long t0 = C.ties[tid];
          double v = 0.0;
          bool have = true;
          if (R.tie_mode == OLS_RNG_FEED) {
            if ((int64_t)t0 < R.tie_feed_len) v = __ldg(R.tie_feed + rr * R.tie_stride + (int64_t)t0);
            else have = false;
          } else {
            // stream 1, high word only (v = hi * 2^-32); four ties per block, the block carried in shared memory
            const uint4 tb = C.tblk[tid];
            const int w = (int)(t0 & 3);
            v = (double)(w == 0 ? tb.x : w == 1 ? tb.y : w == 2 ? tb.z : tb.w) * 2.3283064365386963e-10;
            if (w == 3) {
              const ols_u4 t = ols_stream_block(PK, rid, 1u, (t0 >> 2) + 1);
              C.tblk[tid] = make_uint4(t.x, t.y, t.z, t.w);
            }
          }
          if (!have) {
            new_status = OLS_FEED_EMPTY;
          } else {
            C.ties[tid] = t0 + 1;
            const int pick = min(ncand - 1, (int)(v * (double)ncand));
            const unsigned c1 = cand & (cand - 1), c2 = c1 & (c1 - 1), c3 = c2 & (c2 - 1);
            cand = (pick == 0) ? cand : (pick == 1) ? c1 : (pick == 2) ? c2 : c3;
          }
        }
        a_star = __ffs(cand) - 1;
        if (C.eps_episode[tid] != ep_index0) {  // two distinct probabilities per episode (tabular.py:27,30)
          const double eps = __ldg(cfg.epsilon + ep_index0);
          const double based = __ddiv_rn(eps, (double)nA);
          const float base = (float)based, top = (float)__dadd_rn(__dsub_rn(1.0, eps), based);
          const bool okb = (base == 0.f) || (base >= 9.094947e-13f && base <= 1.0f);
          const bool okt = (top == 0.f) || (top >= 9.094947e-13f && top <= 1.0f);
          C.ipn_base[tid] = (base == 0.f) ? __int_as_float(0x7f800000) : __frcp_rn(base);
          C.ipn_top[tid] = (top == 0.f) ? __int_as_float(0x7f800000) : __frcp_rn(top);
          C.pn_ok[tid] = (okb && okt && (base > 0.f || top > 0.f)) ? 1 : 0;
          C.eps_episode[tid] = ep_index0;
        }
        // the accept rule of this step as per-action limits (make_limits, ols_sim_kernels.cuh): the queue's one distribution
        // against the epsilon-greedy policy — two cached reciprocals and which action holds the larger probability
        uint2 lim[4];
        const float inf = __int_as_float(0x7f800000), ib = C.ipn_base[tid], it = C.ipn_top[tid];
        const float ipn4[4] = {(a_star == 0) ? it : ib, (nA <= 1) ? inf : (a_star == 1) ? it : ib, (nA <= 2) ? inf : (a_star == 2) ? it : ib,
                               (nA <= 3) ? inf : (a_star == 3) ? it : ib};
        make_limits(C.qdist[tid], C.qtame[tid] != 0 && C.pn_ok[tid] != 0, ipn4, lim);
#pragma unroll
        for (int a = 0; a < 4; ++a) C.lim[a][tid] = lim[a];
      }
      if (fresh) {
        C.qrow[tid] = prow;
        row_valid = true;
      }
      if (new_status >= 0) {
        C.status[tid] = new_status;
        live = false;
      }
    }


    // =========================================================== one pop round: psrs.py:39-57 over four queue records
    if (live && !pend) {
      const int remaining = b_len - b_pos;
      if (remaining <= 0) {
        pend = true;
        p_res = OLS_QUEUE_EMPTY;  // psrs.py:44-45
      } else {
        const int n_eval = min(4, remaining);
        // ---- the four hot records (one 32-byte span unless the ring wraps inside it)
        uint32_t pos0 = b_start + (uint32_t)b_pos;
        if (pos0 >= (uint32_t)b_len) pos0 -= (uint32_t)b_len;
        uint2 h[4];
        if (pos0 + 3u < (uint32_t)b_len) {
          const uint2* hp = hot + (size_t)(uint32_t)(b_lo + (int)pos0);
#pragma unroll
          for (int i = 0; i < 4; ++i) h[i] = __ldg(hp + i);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t pos = pos0 + (uint32_t)i;
            if (pos >= (uint32_t)b_len) pos -= (uint32_t)b_len;
            h[i] = (i < n_eval) ? __ldg(hot + (size_t)(uint32_t)(b_lo + (int)pos)) : make_uint2(0u, 0u);
          }
        }
        // ---- uniforms: draw d + i is word (d & 3) + i of the carried block, continued in the next block
        const int a0 = (int)(draws & 3);
        const uint4 o0 = C.blk[tid];
        ols_u4 o1 = {0u, 0u, 0u, 0u};
        if (a0 != 0) o1 = ols_stream_block(PK, rid, 0u, (draws >> 2) + 1);
        uint32_t hi[4];
        hi[0] = a0 == 0 ? o0.x : a0 == 1 ? o0.y : a0 == 2 ? o0.z : o0.w;
        hi[1] = a0 == 0 ? o0.y : a0 == 1 ? o0.z : a0 == 2 ? o0.w : o1.x;
        hi[2] = a0 == 0 ? o0.z : a0 == 1 ? o0.w : a0 == 2 ? o1.x : o1.y;
        hi[3] = a0 == 0 ? o0.w : a0 == 1 ? o1.x : a0 == 2 ? o1.y : o1.z;
        // ---- accept rule per record: u * t_a against the band around inv_m; inside the band (or when a guard fails, or for
        // the 2.4e-4 of the draws below 2^-12, whose float image needs the low word) the exact rule on the queue's distribution
        // 2 bits per slot: 0 reject, 1 accept, 2 undecided
        unsigned decs = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint2 lm = C.lim[(h[i].y >> 29) & 3u][tid];
          const unsigned dec = (i >= n_eval || hi[i] > lm.y) ? 0u : (hi[i] < lm.x) ? 1u : 2u;
          decs |= dec << (2 * i);
        }
        // the first slot that is not a plain reject ends the round if it is an accept; an undecided one (3e-5 of the pops)
        // sends the round to the exact rule, out of line
        int first = decs ? (__ffs(decs) - 1) >> 1 : 4;
        if (first < 4 && ((decs >> (2 * first)) & 3u) == 2u) {
          const unsigned acts = ((h[0].y >> 29) & 3u) | (((h[1].y >> 29) & 3u) << 2) | (((h[2].y >> 29) & 3u) << 4) | (((h[3].y >> 29) & 3u) << 6);
          first = resolve_round(q.queue_dist + 8 * (size_t)b_cur, cfg.epsilon + C.ep_index[tid], ols_keys_seed(PK), rid, draws, decs, acts,
                                a_star, nA, cfg.compute_mode);
        }
        const uint2 won = (first == 0) ? h[0] : (first == 1) ? h[1] : (first == 2) ? h[2] : h[3];
        const int consumed = (first < 4) ? first + 1 : n_eval;
        if (first < 4) {
          pend = true;
          p_res = OLS_RUNNING;
          uint32_t pos = pos0 + (uint32_t)first;
          if (pos >= (uint32_t)b_len) pos -= (uint32_t)b_len;
          p_slot = (uint32_t)b_lo + pos;
          // the next state is known: pull its AoS entry (cursor + Q row) toward L2 while other lanes keep popping
          C.park[tid] = won;
          const int z2 = ((int)(won.y << 3)) >> 3;
          if (z2 >= 0 && z2 < nS) asm volatile("prefetch.global.L2 [%0];" ::"l"(aos + (size_t)z2 * 32));
        }
        // ---- carry the block that holds the next draw
        const uint64_t d_next = draws + (uint64_t)consumed;
        if ((d_next >> 2) != (draws >> 2)) {
          ols_u4 keep = o1;
          if (a0 == 0) keep = ols_stream_block(PK, rid, 0u, d_next >> 2);  // the round consumed the whole carried block
          C.blk[tid] = make_uint4(keep.x, keep.y, keep.z, keep.w);
        }
        draws = d_next;
        b_pos += consumed;
      }
    }

    // =========================================================== the update block, once enough lanes hold an outcome
    const unsigned m_live = __ballot_sync(kFull, live);
    const unsigned m_pend = __ballot_sync(kFull, live && pend);
    const int n_live = __popc(m_live);
    const int want = max(1, min(batch, n_live) * 1);
    if (__popc(m_pend) >= min(want, n_live) && live && pend) {
      pend = false;
      const int ep_index = C.ep_index[tid], lc = C.ep_len_count[tid], t_old = C.ep_t[tid], n_acc_old = C.n_acc[tid];
      const double G_old = C.ep_G[tid], disc = C.ep_disc[tid];
      // index of this transition in the replica's lifetime (for the optional taps only)
      const long long acc_idx = (cfg.trace_row || cfg.td_error) ? (long long)__ldcg(R.accepted + r) + n_acc_old : 0;
      const int b_done = b_cur, cur_done = b_pos;  // the queue just popped from and its new cursor
      const int S = cur_z;
      int new_status = -1;
      const bool accepted = (p_res == OLS_RUNNING);
      bool q_write = false, close_ep = false, append_G = false;
      float4 srow = C.qrow[tid];  // Q[S, :] as it stands before this step's update; the entry of S is rewritten whole below
      float4 carry = make_float4(0.f, 0.f, 0.f, 0.f);
      float qn = 0.f, td = 0.f;
      double G_new = G_old;
      int t_new = t_old;
      const uint2 won = C.park[tid];
      const uint32_t p_rw = won.x, p_zw = won.y;
      const int Z2 = ((int)(p_zw << 3)) >> 3;
      const int done = (int)(p_zw >> 31);
      const int A = (int)((p_zw >> 29) & 3u);
      if (accepted) {
        const double Rw = (double)__uint_as_float(p_rw);
        float4 nq;
        int4 nb;
        unsigned nstart;
        stage_next(p_zw, nq, nb, nstart);
        const bool next_ok = nb.x >= 0;
        if (!done) {
          b_cur = nb.x; b_lo = nb.y; b_len = nb.z; b_pos = nb.w; b_start = nstart;
        }
        if (next_ok && b_cur == b_done) b_pos = cur_done;  // a self-loop: the cursor in memory is still the old one
        bucket_valid = next_ok;
        need_step = true;
        if (Z2 < 0 || Z2 >= nS) {
          new_status = OLS_NO_QUEUE;
        } else {
          const double alpha = __ldg(cfg.alpha + ep_index);
          const float n0 = nq.x, n1 = nA > 1 ? nq.y : -INFINITY, n2 = nA > 2 ? nq.z : -INFINITY, n3 = nA > 3 ? nq.w : -INFINITY;
          float boot = n0;  // Q[S_].max(), left to right
          boot = (n1 > boot) ? n1 : boot;
          boot = (n2 > boot) ? n2 : boot;
          boot = (n3 > boot) ? n3 : boot;
          const float qsa = (A == 0) ? srow.x : (A == 1) ? srow.y : (A == 2) ? srow.z : srow.w;
          td = QMath<float>::td((float)Rw, (float)cfg.gamma, boot, qsa);  // psrs.py:165
          qn = QMath<float>::upd(qsa, (float)alpha, td);                  // psrs.py:168
          q_write = true;
          const bool self = (Z2 == S);  // the bootstrap row becomes the next step's policy row (patched if S' == S)
          carry = make_float4((self && A == 0) ? qn : n0, (self && A == 1) ? qn : n1, (self && A == 2) ? qn : n2,
                              (self && A == 3) ? qn : n3);
          row_valid = true;
          G_new = __dadd_rn(G_old, __dmul_rn(disc, Rw));  // G + gamma**t * R   (psrs.py:171)
          t_new = t_old + 1;
          if (done) {  // psrs.py:176-178
            close_ep = true;
            append_G = true;
          } else if (!next_ok) {
            new_status = OLS_NO_QUEUE;  // the next state has no queue: KeyError on the reference's next step (psrs.py:44)
          }
        }
      } else {
        // exhaustion in mid-episode (psrs.py:160-162): the partial return is still appended (:177-178)
        new_status = p_res;
        close_ep = true;
        append_G = true;
      }
      const bool commit = accepted && q_write;
      if (b_done >= 0) {
        // b_done == S (the observation id is the queue key): cursor and Q[S, A] go out as one 32-byte store
        if (commit) {
          srow.x = (A == 0) ? qn : srow.x;
          srow.y = (A == 1) ? qn : srow.y;
          srow.z = (A == 2) ? qn : srow.z;
          srow.w = (A == 3) ? qn : srow.w;
        }
        store_aos(aos, b_done, cur_done, cfg.state_epoch, srow);
      }
      if (commit) {
        C.qrow[tid] = carry;
        if (cfg.td_error && acc_idx < cfg.trace_cap) cfg.td_error[rr * cfg.trace_cap + acc_idx] = (double)td;
        if (cfg.trace_row && acc_idx < cfg.trace_cap) {
          // the row index lives in the full record only: a parity tap, off in production
          const int32_t row_idx = __ldg((const int32_t*)((const char*)q.records + (size_t)p_slot * 32) + 6);
          cfg.trace_row[rr * cfg.trace_cap + acc_idx] = row_idx;
        }
        cur_z = Z2;
        C.n_acc[tid] = n_acc_old + 1;
        C.ep_G[tid] = G_new;
        C.ep_disc[tid] = __dmul_rn(disc, cfg.gamma);
        C.ep_t[tid] = t_new;
        if (budget > 0 && n_acc_old + 1 >= budget) live = false;  // budget: stays resumable
      }
      if (close_ep) {
        if (append_G && ep_index < cfg.ep_cap) cfg.Gs[rr * (int64_t)cfg.ep_cap + ep_index] = G_new;
        if (lc < cfg.ep_cap) cfg.lengths[rr * (int64_t)cfg.ep_cap + lc] = t_new;
        if (append_G) C.ep_index[tid] = ep_index + 1;
        C.ep_len_count[tid] = lc + 1;
        C.ep_open[tid] = 0;
      }
      if (new_status >= 0) {
        C.status[tid] = new_status;
        live = false;
      }
    }
  }

  if (exists) {
    R.status[r] = C.status[tid];
    R.cur_z[r] = cur_z;
    R.cur_s[r] = cur_z;
    R.init_cursor[r] = C.init_cursor[tid];
    R.pops[r] = __ldcg(R.pops + r) + (unsigned long long)(draws - __ldcg(R.draws + r));
    R.draws[r] = draws;
    R.ties[r] = C.ties[tid];
    R.accepted[r] = __ldcg(R.accepted + r) + (unsigned long long)C.n_acc[tid];
    R.ep_index[r] = C.ep_index[tid];
    R.ep_len_count[r] = C.ep_len_count[tid];
    R.ep_t[r] = C.ep_t[tid];
    R.ep_open[r] = C.ep_open[tid];
    R.ep_G[r] = C.ep_G[tid];
    R.ep_disc[r] = C.ep_disc[tid];
  }
  if (cfg.slices > 1) {
    __syncthreads();
    if (tid == 0) {
      __threadfence();  // release: every thread's state writes precede the token
      const int n_blk = gridDim.x / cfg.slices;
      *((volatile int*)cfg.sched + 1 + s_ticket % n_blk) = s_ticket / n_blk + 1;
    }
  }
}

}  // namespace ols_sim

cudaError_t ols_launch_learner(const ols_queues_t& q, const ols_replicas_t& r, const ols_run_t& c, const ols_sim::RunLaunch& L) {
  ols_sim::learner_kernel<<<L.grid, ols_sim::kLT, 0, L.stream>>>(q, r, c, L.keys);
  ols_note_launch(1);
  return cudaGetLastError();
}
