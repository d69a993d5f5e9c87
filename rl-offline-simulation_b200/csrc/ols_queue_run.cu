// Fused multi-step driver of the QueueEvaluator: offsim4rl/evaluators/queue_evaluator.py:113-131 run for many steps in ONE
// launch, one thread per replica.
//
//   :113-119 reset       pop the next initial (z, s); None when the list is exhausted
//   :121-131 step(a)     pop the head of queue (z, a) — queues keyed by (z, a) (:107), no rejection; None when it is empty,
//                        KeyError when (z, a) was never logged
//
// The reference has no driver loop for this simulator (its gym wrapper is stepped by an Agent, queue_evaluator.py:80-88);
// the loop here is the one of its tests / of evalMC_psrs (psrs.py:241-271) with the action chosen by the AGENT: either a
// fed action list (per replica or shared: parity runs, agents whose actions were computed elsewhere) or a fixed policy
// pi[s] sampled with the replica's own counter-based stream (stream 1 of protocol v2, ols_common.cuh: draw t picks
// a = #{ j : cumsum_j(pi[s]) <= v }, v = hi * 2^-32, the cumulative sum in float64 from action 0 up, clipped to nA - 1).
// Returns are accumulated like evalMC_psrs: G += gamma^t R; a completed episode appends G and its length, an episode cut by
// an empty queue appends its length only.  Every step is one random 32-byte record read plus the replica's cursor: the
// kernel is bound by DRAM latency, which one thread per replica and a full grid hide.
#include "ols_sim_kernels.cuh"

namespace {

using namespace ols_sim;

__global__ void __launch_bounds__(256) queue_run_kernel(ols_queues_t q, ols_replicas_t R, ols_queue_run_t cfg, ols_philox_keys PK) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= R.n_replicas) return;
  const int nA = q.n_actions;
  const uint64_t rid = (uint64_t)(R.replica_id0 + r);
  const uint32_t h_rep = R.ring_mode ? ols_ring_replica_hash(R.ring_seed, rid) : 0u;
  int status = R.status[r];
  if (!(status == OLS_IDLE || status == OLS_RUNNING)) return;  // terminated replicas stay terminated
  int cur_z = R.cur_z[r], cur_s = R.cur_s[r], init_cursor = R.init_cursor[r];
  int ep_index = R.ep_index[r], lc = R.ep_len_count[r], ep_t = R.ep_t[r], ep_open = R.ep_open[r];
  double G = R.ep_G[r], disc = R.ep_disc[r];
  unsigned long long n_draw = R.ties[r], accepted = R.accepted[r];
  long long done_now = 0;
  int32_t* cur_tab = R.cursor + r * (int64_t)q.n_buckets;
  status = OLS_RUNNING;
  for (;;) {
    if (!ep_open) {  // psrs.py:246-253 shape: an episode starts with reset()
      if (ep_index >= cfg.n_episodes) { status = OLS_DONE; break; }
      if (!reset_one(q, R, h_rep, init_cursor, cur_z, cur_s)) { status = OLS_INIT_EMPTY; break; }
      ep_open = 1; ep_t = 0; G = 0.0; disc = 1.0;
    }
    // ---- the agent's action
    int a;
    if (cfg.actions) {
      if ((int64_t)n_draw >= cfg.actions_len) { status = OLS_FEED_EMPTY; break; }
      a = __ldg(cfg.actions + r * cfg.a_stride + (int64_t)n_draw);
    } else {
      if (cur_s < 0 || cur_s >= cfg.n_states) { status = OLS_NO_QUEUE; break; }  // IndexError on pi[s]
      const double v = (double)ols_u4_word(ols_stream_block(PK, rid, 1u, n_draw >> 2), (int)(n_draw & 3)) * 2.3283064365386963e-10;
      double cum = 0.0;
      a = 0;
      for (int j = 0; j < nA; ++j) {
        const int64_t off = (int64_t)cur_s * nA + j;
        cum = __dadd_rn(cum, cfg.pi_is_f64 ? __ldg((const double*)cfg.pi + off) : (double)__ldg((const float*)cfg.pi + off));
        a += (cum <= v) ? 1 : 0;
      }
      a = min(a, nA - 1);
    }
    n_draw += 1;
    // ---- env.step(a): queue_evaluator.py:121-131
    Bucket B;
    if ((unsigned)a >= (unsigned)nA ||
        !load_bucket(q, R, h_rep, cur_tab, (cur_z + q.z_offset) * nA + a - q.z_offset, B)) { status = OLS_NO_QUEUE; break; }  // KeyError
    if (B.cur >= B.len) { status = OLS_QUEUE_EMPTY; break; }                                                                // step -> None
    uint32_t pos = B.start + (uint32_t)B.cur;
    if (pos >= (uint32_t)B.len) pos -= (uint32_t)B.len;
    const float4* rp = (const float4*)((const char*)q.records + (size_t)(uint32_t)(B.lo + (int32_t)pos) * q.rec_bytes);
    const float4 tail = __ldg(rp + (q.rec_bytes / 16 - 1));
    cur_tab[B.b] = B.cur + 1;
    const uint32_t meta = __float_as_uint(tail.w);
    const int32_t row = __float_as_int(tail.z);
    const double Rw = (double)tail.x;
    if (cfg.trace_row && (long long)accepted < cfg.trace_cap) cfg.trace_row[r * cfg.trace_cap + (long long)accepted] = row;
    cur_z = __float_as_int(tail.y);
    cur_s = q.s_next ? __ldg(q.s_next + row) : cur_z;
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
  if (status == OLS_NO_QUEUE) n_draw -= 1;  // the reference raises inside step(): that action does not count as taken
  R.status[r] = status;
  R.cur_z[r] = cur_z;
  R.cur_s[r] = cur_s;
  R.init_cursor[r] = init_cursor;
  R.ties[r] = n_draw;
  R.pops[r] += (unsigned long long)done_now;
  R.accepted[r] = accepted;
  R.ep_index[r] = ep_index;
  R.ep_len_count[r] = lc;
  R.ep_t[r] = ep_t;
  R.ep_open[r] = ep_open;
  R.ep_G[r] = G;
  R.ep_disc[r] = disc;
}

}  // namespace

extern "C" int ols_run_actions(const ols_queues_t* q, const ols_replicas_t* r, const ols_queue_run_t* c, void* stream) {
  OLS_REQUIRE(q && r && c, OLS_ERR_INVALID, "ols_run_actions: null argument");
  OLS_REQUIRE(q->n_actions >= 1 && q->n_actions <= OLS_MAX_ACTIONS && q->rec_bytes == ols_record_bytes(q->n_actions), OLS_ERR_INVALID,
              "ols_run_actions: bad queues");
  OLS_REQUIRE(q->n_buckets % q->n_actions == 0, OLS_ERR_INVALID,
              "ols_run_actions: queues must be keyed by (z, a): n_buckets = n_state_keys * n_actions");
  OLS_REQUIRE(c->n_episodes >= 0 && c->ep_cap >= c->n_episodes + 1, OLS_ERR_INVALID, "ols_run_actions: need ep_cap >= n_episodes + 1");
  OLS_REQUIRE((c->actions != nullptr) != (c->pi != nullptr), OLS_ERR_INVALID, "ols_run_actions: give either an action list or a policy");
  OLS_REQUIRE(!c->actions || (c->actions_len >= 0 && (c->a_stride == 0 || c->a_stride >= c->actions_len)), OLS_ERR_INVALID,
              "ols_run_actions: a_stride must be 0 (shared list) or >= actions_len");
  OLS_REQUIRE(!c->pi || c->n_states >= 1, OLS_ERR_INVALID, "ols_run_actions: a policy needs n_states");
  OLS_REQUIRE(!c->trace_row || c->trace_cap > 0, OLS_ERR_INVALID, "ols_run_actions: the row trace needs trace_cap > 0");
  if (r->n_replicas == 0) return OLS_OK;
  OLS_REQUIRE(r->cursor && r->init_cursor && r->cur_z && r->cur_s && r->status && r->ties && r->pops && r->accepted && r->ep_index &&
                  r->ep_len_count && r->ep_t && r->ep_open && r->ep_G && r->ep_disc,
              OLS_ERR_INVALID, "ols_run_actions: null replica buffer");
  OLS_REQUIRE(c->Gs && c->lengths && q->records && q->offsets, OLS_ERR_INVALID, "ols_run_actions: null buffer");
  const int grid = ols_div_up(r->n_replicas, 256);
  queue_run_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*q, *r, *c, ols_philox_make_keys(r->seed));
  OLS_LAUNCH_CHECK();
  return OLS_OK;
}
