/*
 * ols_oracle.c — plain-C restatement of the reference's PSRS hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library; the
 * product never does (it has no CPU path).  The semantics are those of
 * /root/reference/offsim4rl/evaluators/psrs.py (PSRS.reset :32-37, PSRS.step :39-51,
 * _default_reject :53-57, qlearn_psrs :119-185, expSARSA_psrs :187-239, evalMC_psrs :241-271) and
 * agents/tabular.py:4-5,24-32 (epsilon-greedy with random arg-max), restated on CSR queues with one
 * cursor per queue exactly like oracle/psrs_oracle.py, against which (and through it against the
 * golden vectors of the unmodified reference) this file is checked in tests/test_oracle_c.py.
 * Parity status: pinned (transitively, via the golden vectors).
 *
 * Uniform streams: either harness-fed arrays or Philox4x32-7 (Salmon et al., SC'11) with the same
 * (seed, replica, draw) -> value mapping the CUDA path documents, so that CPU and GPU replicas can be
 * compared draw for draw at sizes the python oracle cannot reach.  Replicas are independent: the
 * replicas are self-scheduled over POSIX threads (this is also the multi-core CPU baseline).
 *
 * Build: see oracle/Makefile (gcc -O2 -pthread -ffp-contract=off; no fast-math, IEEE division).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>

enum { ST_RUNNING = 0, ST_DONE = 1, ST_QUEUE_EMPTY = 2, ST_INIT_EMPTY = 3, ST_NO_QUEUE = 4, ST_FEED_EMPTY = 5, ST_IDLE = 6 };
enum { AGENT_FIXED_PI = 0, AGENT_EXPSARSA = 1, AGENT_QLEARN = 2 };
enum { RNG_FEED = 0, RNG_PHILOX = 2 };

/* ------------------------------------------------------------------------------------------ uniform streams */
/* Philox4x32 (Salmon et al., SC'11) with 7 rounds.  Stream protocol v2, as documented in
 * rl-offline-simulation_b200/csrc/ols_common.cuh: draw d of stream `which` of replica `stream` under key `seed` is
 *     blk = d >> 2, w = d & 3
 *     hi  = Philox4x32-7({blk_lo, blk_hi, stream_lo, stream_hi ^ (which << 28)},       seed)[w]
 *     lo  = Philox4x32-7({blk_lo, blk_hi, stream_lo, stream_hi ^ ((which + 2) << 28)}, seed)[w] >> 11
 *     u   = ((hi << 21) | lo) * 2^-53                       (rejection uniforms: which = 0)
 * and an arg-max tie-break uses the high word alone: v = hi * 2^-32 (which = 1).                              */
static void philox4x32_7(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int i = 0; i < 7; ++i) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

static uint32_t stream_word(uint64_t seed, uint64_t stream, uint32_t which, uint64_t d) {
  uint64_t blk = d >> 2;
  uint32_t c[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)stream, (uint32_t)(stream >> 32) ^ (which << 28)};
  philox4x32_7(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  return c[d & 3];
}

static double philox_uniform(uint64_t seed, uint64_t stream, uint32_t which, uint64_t d) {
  uint64_t hi = stream_word(seed, stream, which, d), lo = stream_word(seed, stream, which + 2u, d) >> 11;
  return (double)((hi << 21) | lo) * (1.0 / 9007199254740992.0);
}

static double philox_tie(uint64_t seed, uint64_t stream, uint64_t t) {
  return (double)stream_word(seed, stream, 1u, t) * (1.0 / 4294967296.0);
}

/* ring offsets (replica decorrelation): 32-bit murmur3-finaliser hashing, as documented in DESIGN.md §6 */
static uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
static uint32_t ring_replica_hash(uint64_t ring_seed, uint64_t replica) {
  uint32_t h = fmix32((uint32_t)(replica >> 32));
  h = fmix32((uint32_t)replica ^ h);
  h = fmix32((uint32_t)(ring_seed >> 32) ^ h);
  return fmix32((uint32_t)ring_seed ^ h);
}
static uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static uint32_t ring_start(uint64_t ring_seed, uint64_t replica, uint32_t bucket, uint32_t len) {
  return mulhi32(fmix32(ring_replica_hash(ring_seed, replica) ^ ((bucket + 1u) * 0x9E3779B1u)), len);
}
static int64_t init_ring_start(uint64_t ring_seed, uint64_t replica, int64_t n_init) {
  return (int64_t)mulhi32(fmix32(ring_replica_hash(ring_seed, replica) ^ 0x85EBCA6Bu), (uint32_t)n_init);
}

/* ------------------------------------------------------------------------------------------ bucketing */
/* psrs.py:26 — stable grouping of rows by from-state id; returns the number of out-of-range rows. */
int64_t oracle_bucket(const int32_t* z, int64_t n, int32_t n_buckets, int32_t z_offset, int32_t* offsets, int32_t* order) {
  int64_t bad = 0;
  int64_t* count = (int64_t*)calloc((size_t)n_buckets + 1, sizeof(int64_t));
  for (int64_t i = 0; i < n; ++i) {
    int64_t b = (int64_t)z[i] + z_offset;
    if (b < 0 || b >= n_buckets) { ++bad; continue; }
    count[b + 1]++;
  }
  for (int32_t b = 0; b < n_buckets; ++b) count[b + 1] += count[b];
  for (int32_t b = 0; b <= n_buckets; ++b) offsets[b] = (int32_t)count[b];
  for (int64_t i = 0; i < n; ++i) {
    int64_t b = (int64_t)z[i] + z_offset;
    if (b < 0 || b >= n_buckets) continue;
    order[count[b]++] = (int32_t)i;
  }
  free(count);
  return bad;
}

/* psrs.py:22 — rows with step == 0, in log order (all rows when there is no step column). */
int64_t oracle_init_rows(const int32_t* step, int64_t n, int32_t* init_row) {
  int64_t k = 0;
  for (int64_t i = 0; i < n; ++i)
    if (!step || step[i] == 0) init_row[k++] = (int32_t)i;
  return k;
}

/* ------------------------------------------------------------------------------------------ accept rule */
/* psrs.py:53-57 in float64: NaN-propagating max, left-to-right division, accept on !(u > thr). */
static int accepts_f64(const double* pn, const double* pl, int nA, int a, double u) {
  double m = pn[0] / pl[0];
  for (int j = 1; j < nA; ++j) {
    double q = pn[j] / pl[j];
    if (!(m != m) && (q != q || q > m)) m = q;
  }
  double thr = pn[a] / pl[a] / m;
  return !(u > thr);
}
/* both distributions float32: float32 ratio/max/threshold, u compared as float32 (NumPy >= 2) */
static int accepts_f32(const double* pn, const double* pl, int nA, int a, double u) {
  float m = (float)pn[0] / (float)pl[0];
  for (int j = 1; j < nA; ++j) {
    float q = (float)pn[j] / (float)pl[j];
    if (!(m != m) && (q != q || q > m)) m = q;
  }
  float thr = (float)pn[a] / (float)pl[a] / m;
  return !((float)u > thr);
}

/* ------------------------------------------------------------------------------------------ run */
typedef struct oracle_sim {
  /* log columns, row order */
  int64_t n_rows;
  int32_t n_actions, n_buckets, z_offset, p_is_f64;
  const int32_t *z, *z_next, *s, *s_next, *action;
  const float* reward;
  const uint8_t* terminal;
  const void* p_log;
  /* queues */
  const int32_t *offsets, *order, *init_row;
  int64_t n_init;
  /* replicas */
  int64_t n_replicas, replica_id0;
  int32_t rng_mode, tie_mode;
  uint64_t seed;
  const double* u_feed; int64_t u_feed_len, u_stride;
  const double* tie_feed; int64_t tie_feed_len, tie_stride;
  int32_t ring_mode; uint64_t ring_seed;
  /* agent */
  int32_t agent, n_states, accept_all, compute_f32, pi_is_f64;
  int32_t n_episodes, ep_cap;
  int64_t max_steps;
  double gamma;
  const double *alpha, *epsilon;
  const void* pi;
  /* outputs */
  double* Q;            /* [R][nS][nA] (may be NULL for FIXED_PI) */
  double* Gs;           /* [R][ep_cap] */
  int32_t* lengths;     /* [R][ep_cap] */
  int32_t *n_returns, *n_lengths, *status;
  int64_t *accepted, *pops, *draws, *ties;
  int32_t* trace_row;   /* [R][trace_cap] or NULL */
  double* td_error;     /* [R][trace_cap] or NULL */
  int64_t trace_cap;
  uint64_t* trace_hash; /* [R] order-sensitive hash of the accepted rows (always written) */
  /* float32 learner (the GPU production mode): Q entries are float32 values (held in the double array, exactly), and the
   * update psrs.py:165-168 runs as four separately rounded float32 operations on float32 images of R, gamma, alpha:
   *     td = ((R + gamma * boot) - Q[S,A]);   Q[S,A] = Q[S,A] + alpha * td
   * (no contraction: -ffp-contract=off).  The CUDA side states the same order with __fmul_rn/__fadd_rn/__fsub_rn.     */
  int32_t q_is_f32;
  int32_t _pad0;
  const int64_t* replica_ids; /* [R] global replica ids (keys of the Philox streams and ring offsets), or NULL =
                                 replica_id0 + r: lets a test run a SAMPLE of a large fleet's replicas                */
} oracle_sim_t;

static double p_at(const oracle_sim_t* S, int64_t row, int j) {
  return S->p_is_f64 ? ((const double*)S->p_log)[row * S->n_actions + j] : (double)((const float*)S->p_log)[row * S->n_actions + j];
}
static double pi_at(const oracle_sim_t* S, int s, int j) {
  return S->pi_is_f64 ? ((const double*)S->pi)[(int64_t)s * S->n_actions + j] : (double)((const float*)S->pi)[(int64_t)s * S->n_actions + j];
}

static void run_replica(const oracle_sim_t* S, int64_t r, int32_t* cursor) {
  const int nA = S->n_actions, nS = S->n_states;
  const uint64_t rid = S->replica_ids ? (uint64_t)S->replica_ids[r] : (uint64_t)(S->replica_id0 + r);
  double* Q = S->Q ? S->Q + r * (int64_t)nS * nA : NULL;
  double* Gs = S->Gs + r * (int64_t)S->ep_cap;
  int32_t* lens = S->lengths + r * (int64_t)S->ep_cap;
  memset(cursor, 0, sizeof(int32_t) * (size_t)S->n_buckets);
  int64_t init_cursor = 0, accepted = 0, pops = 0, steps = 0;
  uint64_t draws = 0, ties = 0, hash = 0xcbf29ce484222325ull;
  int ep_index = 0, n_len = 0, status = ST_RUNNING;
  const int64_t init_off = (S->ring_mode && S->n_init > 0) ? init_ring_start(S->ring_seed, rid, S->n_init) : 0;
  double pn[16], pl[16];

  while (status == ST_RUNNING) {
    if (ep_index >= S->n_episodes) { status = ST_DONE; break; }
    /* reset: psrs.py:32-37 */
    if (init_cursor >= S->n_init) { status = ST_INIT_EMPTY; break; }
    int64_t ipos = init_cursor + init_off;
    if (ipos >= S->n_init) ipos -= S->n_init;
    int32_t irow = S->init_row[ipos];
    ++init_cursor;
    int cz = S->z[irow], cs = S->s ? S->s[irow] : S->z[irow];
    double G = 0.0, disc = 1.0;
    int t = 0, done = 0, budget_hit = 0;
    while (!done) {
      if (cs < 0 || cs >= nS) { status = ST_NO_QUEUE; break; }
      /* the agent's distribution */
      if (S->agent == AGENT_QLEARN) {
        const double eps = S->epsilon[ep_index];
        const double base = eps / (double)nA, top = (1.0 - eps) + base;   /* tabular.py:24-32 */
        const double* row = Q + (int64_t)cs * nA;
        double mx = row[0];
        for (int j = 1; j < nA; ++j) if (row[j] > mx) mx = row[j];
        int ncand = 0;
        for (int j = 0; j < nA; ++j) ncand += (row[j] == mx);
        int pick = 0;
        if (ncand > 1) {                                                  /* tabular.py:4-5 */
          double v;
          if (S->tie_mode == RNG_FEED) {
            if ((int64_t)ties >= S->tie_feed_len) { status = ST_FEED_EMPTY; break; }
            v = S->tie_feed[r * S->tie_stride + (int64_t)ties];
          } else v = philox_tie(S->seed, rid, ties);
          ++ties;
          pick = (int)(v * (double)ncand);
          if (pick > ncand - 1) pick = ncand - 1;
        }
        int seen = 0, a_star = 0;
        for (int j = 0; j < nA; ++j) if (row[j] == mx) { if (seen == pick) a_star = j; ++seen; }
        for (int j = 0; j < nA; ++j) pn[j] = (j == a_star) ? top : base;
      } else {
        for (int j = 0; j < nA; ++j) pn[j] = pi_at(S, cs, j);
      }
      /* env.step: psrs.py:39-51 */
      const int b = cz + S->z_offset;
      if (b < 0 || b >= S->n_buckets || S->offsets[b + 1] == S->offsets[b]) { status = ST_NO_QUEUE; break; }
      const int32_t lo = S->offsets[b], len = S->offsets[b + 1] - lo;
      const uint32_t start = S->ring_mode ? ring_start(S->ring_seed, rid, (uint32_t)b, (uint32_t)len) : 0u;
      int32_t row = -1;
      for (;;) {
        if (cursor[b] >= len) { status = ST_QUEUE_EMPTY; break; }
        uint32_t pos = start + (uint32_t)cursor[b];
        if (pos >= (uint32_t)len) pos -= (uint32_t)len;
        int32_t cand = S->order[lo + (int32_t)pos];
        if (S->accept_all) { cursor[b]++; ++pops; row = cand; break; }
        double u;
        if (S->rng_mode == RNG_FEED) {
          if ((int64_t)draws >= S->u_feed_len) { status = ST_FEED_EMPTY; break; }
          u = S->u_feed[r * S->u_stride + (int64_t)draws];
        } else u = philox_uniform(S->seed, rid, 0u, draws);
        cursor[b]++; ++pops; ++draws;
        for (int j = 0; j < nA; ++j) pl[j] = p_at(S, cand, j);
        const int a = S->action[cand];
        if (S->compute_f32 ? accepts_f32(pn, pl, nA, a, u) : accepts_f64(pn, pl, nA, a, u)) { row = cand; break; }
      }
      if (row < 0) break;
      /* learner update */
      const int A = S->action[row];
      const int s2 = S->s_next ? S->s_next[row] : S->z_next[row];
      const double R = (double)S->reward[row];
      done = S->terminal[row] != 0;
      if (S->agent != AGENT_FIXED_PI) {
        if (s2 < 0 || s2 >= nS) { status = ST_NO_QUEUE; break; }
        const double alpha = S->alpha[ep_index];
        double boot;
        const double* nrow = Q + (int64_t)s2 * nA;
        if (S->agent == AGENT_QLEARN) {
          boot = nrow[0];
          for (int j = 1; j < nA; ++j) if (nrow[j] > boot) boot = nrow[j];
        } else {
          boot = 0.0;
          for (int j = 0; j < nA; ++j) boot = fma(nrow[j], pi_at(S, s2, j), boot);
          if (S->q_is_f32) boot = (double)(float)boot;
        }
        double* q = Q + (int64_t)cs * nA + A;
        double td;
        if (S->q_is_f32) {
          const float Rf = (float)R, gf = (float)S->gamma, af = (float)alpha, bf = (float)boot, qf = (float)*q;
          const float gb = gf * bf;
          const float tgt = Rf + gb;
          const float tdf = tgt - qf;                                      /* psrs.py:165, float32 */
          const float step = af * tdf;
          *q = (double)(qf + step);                                        /* psrs.py:168, float32 */
          td = (double)tdf;
        } else {
          td = R + S->gamma * boot - *q;                                   /* psrs.py:165 */
          *q = *q + alpha * td;                                            /* psrs.py:168 */
        }
        if (S->td_error && accepted < S->trace_cap) S->td_error[r * S->trace_cap + accepted] = td;
      }
      if (S->trace_row && accepted < S->trace_cap) S->trace_row[r * S->trace_cap + accepted] = row;
      hash = (hash ^ (uint64_t)(uint32_t)row) * 0x100000001b3ull;
      cz = S->z_next[row];
      cs = s2;
      ++accepted;
      G = G + disc * R;                                                    /* psrs.py:171 */
      disc = disc * S->gamma;
      ++t;
      ++steps;
      if (!done) {
        /* The accepted transition leads to a state without a queue: the reference's NEXT env.step raises KeyError (psrs.py:44)
         * and the whole run is lost; the replica ends here (committed transition, nothing appended), whatever the budget says. */
        const int b2 = cz + S->z_offset;
        if (b2 < 0 || b2 >= S->n_buckets || S->offsets[b2 + 1] == S->offsets[b2]) { status = ST_NO_QUEUE; break; }
      }
      if (S->max_steps > 0 && steps >= S->max_steps) { budget_hit = 1; break; }
    }
    if (status == ST_NO_QUEUE || status == ST_FEED_EMPTY) break;          /* error: nothing is appended */
    if (done) {                                                            /* psrs.py:176-178 / 265-269 */
      if (n_len < S->ep_cap) lens[n_len] = t;
      ++n_len;
      if (ep_index < S->ep_cap) Gs[ep_index] = G;
      ++ep_index;
    } else if (status == ST_QUEUE_EMPTY) {
      /* exhaustion in mid-episode: qlearn/expSARSA still append the partial return (psrs.py:160-162,177),
       * evalMC appends only the length (psrs.py:265-269) */
      if (n_len < S->ep_cap) lens[n_len] = t;
      ++n_len;
      if (S->agent != AGENT_FIXED_PI) {
        if (ep_index < S->ep_cap) Gs[ep_index] = G;
        ++ep_index;
      }
      break;
    }
    if (budget_hit) break;                                                 /* step budget: an open episode is not appended */
  }
  S->n_returns[r] = ep_index;
  S->n_lengths[r] = n_len;
  S->status[r] = status;
  S->accepted[r] = accepted;
  S->pops[r] = pops;
  S->draws[r] = (int64_t)draws;
  S->ties[r] = (int64_t)ties;
  S->trace_hash[r] = hash;
}

/* Runs every replica from a fresh state on n_threads POSIX threads (dynamic self-scheduling over
 * replicas through one atomic counter); returns the number of threads used. */
typedef struct worker_arg {
  const oracle_sim_t* S;
  atomic_llong* next;
} worker_arg_t;

static void* worker(void* p) {
  worker_arg_t* w = (worker_arg_t*)p;
  int32_t* cursor = (int32_t*)malloc(sizeof(int32_t) * (size_t)w->S->n_buckets);
  for (;;) {
    long long r = atomic_fetch_add(w->next, 1);
    if (r >= w->S->n_replicas) break;
    run_replica(w->S, r, cursor);
  }
  free(cursor);
  return NULL;
}

int oracle_run(const oracle_sim_t* S, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if ((int64_t)n_threads > S->n_replicas) n_threads = S->n_replicas > 0 ? (int)S->n_replicas : 1;
  atomic_llong next = 0;
  worker_arg_t arg = {S, &next};
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
  int started = 0;
  for (int i = 1; i < n_threads; ++i)
    if (pthread_create(&th[started], NULL, worker, &arg) == 0) ++started;
  worker(&arg);
  for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
  free(th);
  return started + 1;
}

size_t oracle_sim_bytes(void) { return sizeof(oracle_sim_t); }

/* ------------------------------------------------------------------------------------------ encoders */
/* encoders/models.py:15-19 + homer.py:159-168: bias, then fmaf over inputs in index order, first arg-max */
void oracle_mlp_ids(const float* x, int64_t n, int dO, const float* W1, const float* b1, int H, const float* W2, const float* b2,
                    int nZ, float slope, int32_t* ids) {
  float* h = (float*)malloc(sizeof(float) * (size_t)H);
  for (int64_t i = 0; i < n; ++i) {
    for (int j = 0; j < H; ++j) {
      float acc = b1[j];
      for (int d = 0; d < dO; ++d) acc = fmaf(x[i * dO + d], W1[j * dO + d], acc);
      h[j] = acc > 0.0f ? acc : acc * slope;
    }
    int best = 0;
    float bv = 0.0f;
    for (int k = 0; k < nZ; ++k) {
      float acc = b2[k];
      for (int j = 0; j < H; ++j) acc = fmaf(h[j], W2[k * H + j], acc);
      if (k == 0 || acc > bv) { bv = acc; best = k; }
    }
    ids[i] = best;
  }
  free(h);
}
