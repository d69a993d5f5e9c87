/*
 * ols_b200.h — C ABI of the B200-native Per-State Rejection Sampling (PSRS) engine.
 *
 * This is the drop-in boundary for ONE path of microsoft/rl-offline-simulation (package
 * `offsim4rl`): offline learner simulation by per-state rejection sampling.  Every entry
 * point names the reference interface it replaces (paths relative to the reference's
 * `offsim4rl/` directory).  The reference is pure Python and has no FFI of its own; the
 * binding a maintainer would add is a ctypes stub (shown in INTEGRATION.md, shipped in
 * rl-offline-simulation_b200/_abi.py).
 *
 * Conventions
 *   - plain C: pointers, sizes and POD structs only; no C++ or torch types.
 *   - every pointer is a DEVICE pointer (cudaMalloc'd by the caller, e.g. a torch CUDA tensor's
 *     data_ptr) unless its comment says HOST.  The library allocates nothing and keeps no state:
 *     the caller owns all buffers, including scratch (sizes via the *_scratch_bytes queries).
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous and stream-ordered.
 *   - every function returns OLS_OK or an OLS_ERR_* code and never throws; ols_last_error()
 *     gives the thread-local message of the last failure.
 *   - row indices, queue offsets and cursors are int32: a log holds < 2^31 transitions.
 *   - integer/byte results (ids, offsets, order, accepted rows) are bit-exact with the reference;
 *     the accept rule runs in IEEE float64 (or float32, see OLS_COMPUTE_F32) exactly as NumPy does.
 */
#ifndef OLS_B200_H
#define OLS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OLS_ABI_VERSION 2

/* ---- return codes ------------------------------------------------------------------------- */
enum {
  OLS_OK = 0,
  OLS_ERR_INVALID = 1,      /* bad argument (null pointer, size out of range, misaligned buffer) */
  OLS_ERR_CUDA = 2,         /* a CUDA runtime call or kernel launch failed                        */
  OLS_ERR_UNSUPPORTED = 3   /* valid request outside what the kernels cover (e.g. n_actions > 16) */
};

/* ---- per-replica status (evaluators/psrs.py return conventions) -------------------------- */
enum {
  OLS_RUNNING = 0,        /* holds a current state, can step                                      */
  OLS_DONE = 1,           /* a fused driver reached n_episodes (psrs.py:149)                       */
  OLS_QUEUE_EMPTY = 2,    /* psrs.py:44-45 — step() returned (None,)*4: the state's queue ran dry  */
  OLS_INIT_EMPTY = 3,     /* psrs.py:33-35 — reset() returned None: no initial state left          */
  OLS_NO_QUEUE = 4,       /* psrs.py:44 raises KeyError — state never logged as a from-state.  The fused
                             drivers set it AT the accepted transition that leads into such a state (the
                             transition is committed and counted, no episode is appended, a step budget
                             ending on the same transition does not hide it); the reference would raise
                             on its next env.step and lose the whole run                               */
  OLS_FEED_EMPTY = 5,     /* harness-fed uniform / tie stream exhausted (parity mode only)         */
  OLS_IDLE = 6            /* fresh replica: reset() has not been called yet                        */
};

/* ---- options ------------------------------------------------------------------------------ */
enum { OLS_REJECT_DEFAULT = 0,   /* psrs.py:53-57 `_default_reject`                                */
       OLS_REJECT_ACCEPT_ALL = 1 /* trivial_baselines.py:8-10,22-24 `_reject -> False`             */ };
enum { OLS_COMPUTE_F64 = 0,      /* NumPy promotion when either distribution is float64            */
       OLS_COMPUTE_F32 = 1       /* both p_new and p_log float32: ratio, max and compare in float32 */ };
enum { OLS_RNG_FEED = 0,         /* u = u_feed[replica*stride + draw]  (harness-fed, parity tests)  */
       OLS_RNG_PCG64 = 1,        /* np.random.default_rng(seed).random() replayed bit for bit        */
       OLS_RNG_PHILOX = 2        /* Philox4x32-10 keyed by (seed, replica id), counter = draw index  */ };
enum { OLS_HOT_PROBS = 0, OLS_HOT_TAIL = 1 };   /* ols_queues_t.hot_kind */
enum { OLS_AGENT_FIXED_PI = 0,   /* psrs.py:241-271 evalMC_psrs                                     */
       OLS_AGENT_EXPSARSA = 1,   /* psrs.py:187-239 expSARSA_psrs                                   */
       OLS_AGENT_QLEARN = 2      /* psrs.py:119-185 qlearn_psrs + tabular.py:24-32 epsilon-greedy    */ };

/* ---- the logged transitions, one column per field (data.py:100-118 `experience`) ---------- *
 * Row i is the transition (s_i, a_i, r_i, s'_i, done_i, p_b(.|s_i)); z / z_next are the integer
 * state ids that key the queues (encoder output, per_state_rejection.py:30-31), s / s_next the
 * integer observation the tabular agent indexes Q with (psrs.py:36,49-50; == z in the configs the
 * reference ships).                                                                            */
typedef struct ols_log {
  int64_t n_rows;
  int32_t n_actions;          /* <= 16                                                           */
  int32_t p_is_f64;           /* dtype of p_log: 0 float32, 1 float64                            */
  const int32_t* z;           /* [n_rows] from-state id, in [-z_offset, n_buckets - z_offset)     */
  const int32_t* z_next;      /* [n_rows]                                                        */
  const int32_t* s;           /* [n_rows] or NULL (= z)                                          */
  const int32_t* s_next;      /* [n_rows] or NULL (= z_next)                                     */
  const int32_t* action;      /* [n_rows] in [0, n_actions)                                      */
  const float* reward;        /* [n_rows]                                                        */
  const uint8_t* terminal;    /* [n_rows] 0/1                                                    */
  const void* p_log;          /* [n_rows][n_actions] logged action distribution                  */
  const int32_t* step;        /* [n_rows] or NULL: row is an initial state iff step == 0
                                 (psrs.py:22; data.py:72 — no column means every row is)          */
} ols_log_t;

/* ---- the bucketed log: per-state FIFO queues in CSR form (psrs.py:22-30) ------------------ *
 * Queue b holds slots offsets[b] .. offsets[b+1]-1; slot k carries the packed record of one logged
 * transition.  Record layout (rec_bytes = 16 * (1 + ceil(n_actions / 4)), 32-byte aligned base):
 *     float    p[4 * ceil(n_actions/4)]   logged distribution, padded with 1.0 (rounded to float32
 *                                         when the log is float64; the exact path then reads
 *                                         p_log64[row])
 *     float    reward
 *     int32_t  z_next
 *     int32_t  row                        index of the transition in the log (the queue ORDER:
 *                                         reading this field along a queue gives env.queues[z])
 *     uint32_t meta                       bits 0..7 action, bit 8 terminal, bit 9 "tame" (all
 *                                         logged probabilities in [2^-40, 2^40])                 */
typedef struct ols_queues {
  int64_t n_rows;
  int64_t n_init;             /* number of initial rows                                          */
  int32_t n_actions;
  int32_t n_buckets;          /* number of queue keys                                            */
  int32_t z_offset;           /* bucket = z + z_offset (1 when the encoder can emit -1,
                                 encoders/heuristic.py:23-24)                                    */
  int32_t rec_bytes;
  int32_t* offsets;           /* [n_buckets + 1]                                                 */
  void* records;              /* [n_rows * rec_bytes], 32-byte aligned                           */
  int32_t* init_row;          /* [n_init] initial rows in queue order (psrs.py:22-23)            */
  int32_t* init_z;            /* [n_init] their z                                                */
  int32_t* init_s;            /* [n_init] their s                                                */
  const int32_t* s_next;      /* NULL (= z_next), or the agent-visible next observation id per ROW
                                 when it differs from the state id (psrs.py:36,49-50)            */
  const double* p_log64;      /* NULL, or the float64 log distribution [n_rows][n_actions] in ROW
                                 order when it is not exactly float32-representable              */
  void* hot;                  /* NULL, or [n_rows] 8-byte HOT records in queue order (n_actions <= 4): half-precision
                                 images of the record's probabilities plus its action — all the accept rule needs for
                                 all but ~5e-4 of the pops, so the fused run reads 8 instead of 32 bytes per rejected
                                 record.  Filled by ols_bucket / ols_shuffle_queues when non-NULL.               */
  int32_t hot_kind;           /* what the 8-byte hot records hold:
                                 OLS_HOT_PROBS (0)  the logged probabilities as halves + the action (above): any log;
                                 OLS_HOT_TAIL  (1)  { float reward; uint32 z_next (bits 0..28, two's complement) | action << 29
                                                    | terminal << 31 } — for a log whose behaviour distribution is a function
                                                    of the state, i.e. every queue holds ONE logged distribution (queue_dist):
                                                    the accept rule then depends on a record only through its action, and
                                                    the one-thread-per-replica float32 Q-learner (ols_run with lanes = 1,
                                                    state_aos, Philox uniforms) never reads the 32-byte record.  Every other
                                                    kernel ignores TAIL hot records and pops from the full records.        */
  int32_t _pad;
  float* queue_dist;          /* NULL, or [n_buckets][8] written by ols_bucket / ols_shuffle_queues: floats 0..3 the logged
                                 distribution of the queue's FIRST record (padded with 1.0), word 4 (int32) flags: bit 0 =
                                 every record of the queue holds exactly that distribution (bitwise), bit 1 = tame (all
                                 in [2^-40, 2^40]); words 5..7 zero.  Empty queues count as constant.                      */
  int32_t* const_info;        /* NULL, or [1] written with queue_dist: the number of queues that are NOT constant — hot_kind
                                 OLS_HOT_TAIL is only valid when it is 0 (and p_log64 is NULL)                             */
} ols_queues_t;

/* ---- a block of independent simulation replicas ------------------------------------------ *
 * One replica == one reference `PSRS` object over the same bucketed log (its own queue cursors,
 * initial-state cursor, current (z, s) and uniform stream).                                    */
typedef struct ols_replicas {
  int64_t n_replicas;
  int64_t replica_id0;        /* global id of replica 0 of this block (multi-GPU sharding; keys
                                 the Philox streams and the ring offsets)                        */
  int32_t* cursor;            /* [n_replicas][n_buckets] records consumed from each queue        */
  int32_t* init_cursor;       /* [n_replicas] initial states consumed                            */
  int32_t* cur_z;             /* [n_replicas]                                                    */
  int32_t* cur_s;             /* [n_replicas]                                                    */
  int32_t* status;            /* [n_replicas] OLS_RUNNING ...                                    */
  uint64_t* draws;            /* [n_replicas] uniforms consumed (one per pop, psrs.py:56)        */
  uint64_t* ties;             /* [n_replicas] tie-break values consumed (tabular.py:4-5)         */
  uint64_t* pops;             /* [n_replicas] records popped                                     */
  uint64_t* accepted;         /* [n_replicas] transitions accepted                               */
  /* episode registers of the fused drivers; they make ols_run resumable across calls          */
  int32_t* ep_index;          /* [n_replicas] episodes appended to Gs so far                     */
  int32_t* ep_len_count;      /* [n_replicas] entries written to `lengths` (psrs.py:265)         */
  int32_t* ep_t;              /* [n_replicas] step inside the open episode                       */
  int32_t* ep_open;           /* [n_replicas] 1 while an episode is in progress                  */
  double* ep_G;               /* [n_replicas] discounted return of the open episode              */
  double* ep_disc;            /* [n_replicas] gamma^t of the open episode                        */
  /* uniform stream of the rejection rule                                                      */
  int32_t rng_mode;           /* OLS_RNG_*                                                       */
  int32_t tie_mode;           /* OLS_RNG_FEED or OLS_RNG_PHILOX                                  */
  uint64_t seed;              /* Philox key                                                      */
  uint64_t* pcg_state;        /* [n_replicas][4] {state_hi, state_lo, inc_hi, inc_lo} (PCG64)    */
  const double* u_feed;       /* FEED: uniforms; element [replica * u_stride + draw]             */
  int64_t u_feed_len;
  int64_t u_stride;           /* 0: one stream shared by all replicas                            */
  const double* tie_feed;
  int64_t tie_feed_len;
  int64_t tie_stride;
  /* optional decorrelation of replicas that share one queue order (not in the reference; 0 in
   * parity runs): queue b of replica r is consumed as a ring starting at hash(ring_seed, r, b)
   * mod len_b, the initial-state list at hash(ring_seed, r) mod n_init.                        */
  int32_t ring_mode;          /* 0 off, 1 on                                                     */
  int32_t _pad;
  uint64_t ring_seed;
} ols_replicas_t;

/* ---- batched step: outputs of one accepted transition per replica ------------------------- */
typedef struct ols_step_out {
  int32_t* row;               /* [n_replicas] accepted log row, -1 when nothing was accepted     */
  int32_t* action;            /* [n_replicas] info['a'] (psrs.py:51)                             */
  float* reward;              /* [n_replicas]                                                    */
  uint8_t* terminal;          /* [n_replicas]                                                    */
  int32_t* s_next;            /* [n_replicas]                                                    */
  int32_t* z_from;            /* [n_replicas] info['z']                                          */
} ols_step_out_t;

/* ---- fused learner / evaluator run -------------------------------------------------------- */
typedef struct ols_run {
  int32_t agent;              /* OLS_AGENT_*                                                     */
  int32_t n_states;           /* rows of Q and pi (env.nS, psrs.py:139)                          */
  int32_t q_is_f64;           /* dtype of Q: 0 float32 (production), 1 float64 (parity)          */
  int32_t pi_is_f64;          /* dtype of pi                                                     */
  int32_t reject_mode;        /* OLS_REJECT_*                                                    */
  int32_t compute_mode;       /* OLS_COMPUTE_* (F32 only meaningful with a float32 pi)           */
  int32_t lanes;              /* lane-group width per replica: 1,2,4,8,16 or 32                  */
  int32_t n_episodes;         /* stop a replica after this many episodes                         */
  int32_t ep_cap;             /* row length of Gs / lengths (>= n_episodes + 1)                  */
  int32_t stage_smem;         /* 1: stage the replica's cursor table (and Q) in shared memory
                                 when it fits; 0: operate on global memory                       */
  int32_t pops_per_lane;      /* records each lane examines per round: 1, 2, 4, or 0 = auto (2 with
                                 Philox/fed uniforms, 1 with PCG64 replay)                       */
  int32_t prefetch;           /* rounds ahead whose records are prefetched into L2 (0 = off; useful for
                                 long reject chains, i.e. lanes >= 16)                           */
  int32_t slices;             /* time slicing of a budgeted run (max_steps > 0): 0/1 off; k > 1 = the budget is
                                 served in k slices per replica block (one CTA: 256 threads; 128 for the one-lane
                                 hot-record learner), handed to the SMs through a ticket counter, so a replica
                                 count that is not a whole number of waves keeps every SM busy; -1 = let the
                                 library decide.  Results do not depend on it.                                 */
  int32_t state_epoch;        /* state_aos: only entries whose epoch word equals this value are live, all others
                                 read as cursor 0 / Q 0 — a restart bumps the epoch instead of clearing the table */
  int32_t batch_lanes;        /* one-thread-per-replica learner (OLS_HOT_TAIL): the warp runs the update block once this many
                                 of its 32 replicas hold an outcome (1..32; 0 = library default, 20).  Results do not
                                 depend on it.                                                                    */
  int32_t _pad1;
  int64_t max_steps;          /* > 0: every replica stops after this many accepted transitions in
                                 THIS call and stays resumable (not in the reference)            */
  double gamma;
  const double* alpha;        /* [n_episodes] step size per episode (psrs.py:134-136), QLEARN/EXPSARSA */
  const double* epsilon;      /* [n_episodes] exploration per episode (psrs.py:129-132), QLEARN  */
  const void* pi;             /* [n_states][n_actions] fixed policy, FIXED_PI/EXPSARSA           */
  void* Q;                    /* [n_replicas][n_states][n_actions], QLEARN/EXPSARSA              */
  void* state_aos;            /* optional, instead of r->cursor and Q: [n_replicas][n_states] entries of 32 bytes
                                 { int32 cursor; int32 epoch; int32 pad[2]; float q[4]; } — cursor and Q row of a state in ONE
                                 DRAM sector.  Requires n_actions <= 4, float32 Q, z_offset == 0, n_states ==
                                 n_buckets, no distinct s_next, Philox or fed uniforms (2 records per lane).      */
  int32_t* sched;             /* device scratch of ols_run_sched_bytes() for slices != 0/1 (zeroed by ols_run);
                                 NULL = never slice                                              */
  double* Gs;                 /* [n_replicas][ep_cap] discounted return per episode              */
  int32_t* lengths;           /* [n_replicas][ep_cap] episode lengths                            */
  /* optional parity taps (NULL to disable)                                                    */
  int32_t* trace_row;         /* [n_replicas][trace_cap] accepted rows, indexed by `accepted`    */
  double* td_error;           /* [n_replicas][trace_cap] TD errors (psrs.py:165), QLEARN         */
  int64_t trace_cap;
} ols_run_t;

/* ---- fused QueueEvaluator run (queue_evaluator.py:113-131 for many steps in one launch) -------------------------- */
typedef struct ols_queue_run {
  int32_t n_states;           /* rows of pi (sampled-policy mode)                                 */
  int32_t pi_is_f64;
  int32_t n_episodes;         /* stop a replica after this many completed episodes                */
  int32_t ep_cap;             /* row length of Gs / lengths (>= n_episodes + 1)                   */
  int64_t max_steps;          /* > 0: stop after this many transitions in THIS call (resumable)   */
  double gamma;
  const void* pi;             /* [n_states][n_actions] fixed policy the agent samples its actions from, or NULL */
  const int32_t* actions;     /* or: the agent's actions, element [replica * a_stride + k] for its k-th step() call */
  int64_t actions_len;
  int64_t a_stride;           /* 0: one list shared by all replicas                               */
  double* Gs;                 /* [n_replicas][ep_cap] discounted return per COMPLETED episode     */
  int32_t* lengths;           /* [n_replicas][ep_cap] episode lengths (also the one cut by an empty queue) */
  int32_t* trace_row;         /* [n_replicas][trace_cap] log rows served, or NULL                 */
  int64_t trace_cap;
} ols_queue_run_t;

/* ---- fused PSRS_Exo evaluation (evalMC_psrs, psrs.py:241-271, over psrs.py:59-117) ------------------------------ */
typedef struct ols_exo_run {
  int32_t n_obs;              /* rows of pi: observations o = s * x_mult + x                      */
  int32_t x_mult;             /* the linear o_combine_func (psrs.py:64): o = s * x_mult + x       */
  int32_t pi_is_f64;
  int32_t compute_mode;       /* OLS_COMPUTE_*                                                    */
  int32_t n_episodes;
  int32_t ep_cap;             /* row length of Gs / lengths (>= n_episodes + 1)                   */
  int64_t max_steps;          /* > 0: stop after this many transitions in THIS call (resumable)   */
  double gamma;
  const void* pi;             /* [n_obs][n_actions] the evaluated policy                          */
  const uint64_t* reseed_base;/* PCG64: episode e of replica r re-seeds with reseed_base[r] + e (psrs.py:92 with the
                                 drivers' reset(seed=episode)); NULL = e, the reference's own seeds                 */
  double* Gs;                 /* [n_replicas][ep_cap] discounted return per COMPLETED episode     */
  int32_t* lengths;           /* [n_replicas][ep_cap]                                             */
  int32_t* trace_row;         /* [n_replicas][trace_cap] accepted log rows, or NULL               */
  int64_t trace_cap;
} ols_exo_run_t;

/* ---- library ------------------------------------------------------------------------------ */
int ols_abi_version(void);
const char* ols_last_error(void);                          /* HOST string, thread-local          */
/* sizeof() of each struct as compiled, indexed 0..6 = log, queues, replicas, step_out, run, queue_run, exo_run —
 * lets a binding verify its mirror of the layouts.                                            */
size_t ols_struct_bytes(int which);
/* kernels this library has launched in this process so far (every entry point counts its own launches; memsets and the
 * caller's kernels are not included) — lets a benchmark report how many of its launches were this library's.          */
uint64_t ols_kernel_launches(void);

/* ---- state encoders: observation -> integer state id (the queue key) ---------------------- */
/* Grid discretiser for MyGridNaviCoords observations (envs/gridworld.py:205-211,123-124):
 * id = sum_d min(floor(cells * obs[d]), cells-1) * cells^d, float32 arithmetic.                */
int ols_encode_grid(const float* obs, int64_t n, int32_t d_obs, int32_t cells, int32_t* ids, void* stream);
/* encoders/heuristic.py:20-60 `CartpoleBoxEncoder.get_box` (+ :65-71 `encode`): 162 boxes, -1 on
 * failure; comparisons in the observation dtype (obs_is_f64).  obs is [n][4].                  */
int ols_encode_cartpole_box(const void* obs, int32_t obs_is_f64, int64_t n, int32_t* ids, void* stream);
/* encoders/homer.py:159-168 `HOMEREncoder.encode` over encoders/models.py:15-19 `obs_encoder`:
 * argmax_k( W2 · leaky_relu(W1 · x + b1, slope) + b2 ), float32 FMA in index order, first max.
 * W1 [hidden][d_obs], W2 [n_latent][hidden], row-major as torch.nn.Linear stores them.         */
int ols_encode_mlp(const float* obs, int64_t n, int32_t d_obs, const float* W1, const float* b1, int32_t hidden,
                   const float* W2, const float* b2, int32_t n_latent, float negative_slope, int32_t* ids,
                   void* stream);

/* The same ids, plus what a caller needs to AUDIT them against another float32 implementation of the same network (torch's
 * sgemm sums in a different order, so rows whose two largest logits are closer than the rounding error of a dot product may
 * get another id there): margin[i] = largest - second largest logit, hnorm[i] = sum_j |h_j| of the hidden vector.        */
int ols_encode_mlp_audit(const float* obs, int64_t n, int32_t d_obs, const float* W1, const float* b1, int32_t hidden,
                         const float* W2, const float* b2, int32_t n_latent, float negative_slope, int32_t* ids, float* margin,
                         float* hnorm, void* stream);

/* ---- bucketing: psrs.py:19-30 `PSRS.reset_sampler` ---------------------------------------- */
int32_t ols_record_bytes(int32_t n_actions);
size_t ols_bucket_scratch_bytes(int64_t n_rows, int32_t n_buckets);
/* Stable counting sort of the rows by z (psrs.py:26: sorted + groupby) fused with record packing.
 * Fills q->offsets and q->records (log order inside every queue).  q's scalar fields and buffer
 * pointers are set by the caller.  A row whose key z + z_offset falls outside [0, n_buckets) is
 * DROPPED (the reference would open a new dict entry for it, psrs.py:26): offsets[n_buckets] is the
 * number of rows actually bucketed, so a caller detects drops by comparing it with n_rows (the
 * Python host does, and raises).                                                               */
int ols_bucket(const ols_log_t* log, const ols_queues_t* q, void* scratch, size_t scratch_bytes, void* stream);
/* psrs.py:22 — initial rows (step == 0) in log order -> q->init_row/init_z/init_s; the count is
 * written to *n_init_dev (device int64).  scratch: ols_bucket_scratch_bytes(n_rows, 1).        */
int ols_init_queue(const ols_log_t* log, const ols_queues_t* q, int64_t* n_init_dev, void* scratch,
                   size_t scratch_bytes, void* stream);
/* psrs.py:23,29-30 — replay `np.random.default_rng(seed).shuffle` on the initial-state list and on
 * every queue (a fresh generator each): dst gets the permuted records/init_* of src (src stays
 * the log-order CSR; dst->offsets may alias src->offsets).  perm_scratch: n_rows + n_init int32.  */
int ols_shuffle_queues(const ols_queues_t* src, const ols_queues_t* dst, uint64_t seed, int32_t* perm_scratch,
                       void* stream);
/* (Re)build the derived arrays q points to — hot (in the format q->hot_kind names), queue_dist, const_info — from
 * q->offsets and q->records.  ols_bucket and ols_shuffle_queues call it themselves; a caller uses it to switch the hot
 * format of a bucketed log (e.g. to OLS_HOT_PROBS after const_info reported a queue with several distributions).        */
int ols_pack_hot(const ols_queues_t* q, void* stream);

/* ---- replicas ----------------------------------------------------------------------------- */
/* Zero cursors/counters, status = OLS_IDLE, and (PCG64 mode) seed every replica's generator with
 * seeds[r] (device array, NULL = r->seed for all) exactly as np.random.default_rng does.       */
int ols_replicas_init(const ols_queues_t* q, const ols_replicas_t* r, const uint64_t* seeds, void* stream);
/* psrs.py:32-37 `PSRS.reset` for the replicas with mask[r] != 0 (mask NULL = all): pops the next
 * initial state into (cur_z, cur_s) or sets OLS_INIT_EMPTY.                                     */
int ols_reset(const ols_queues_t* q, const ols_replicas_t* r, const uint8_t* mask, void* stream);
/* psrs.py:39-57 `PSRS.step` + `_default_reject` for every RUNNING replica with mask[r] != 0:
 * p_new is [n_replicas][n_actions] (p_stride = n_actions) or one shared row (p_stride = 0),
 * float64 or float32 (p_new_is_f64).                                                           */
int ols_step(const ols_queues_t* q, const ols_replicas_t* r, const uint8_t* mask, const void* p_new,
             int32_t p_new_is_f64, int64_t p_stride, int32_t reject_mode, int32_t compute_mode, int32_t lanes,
             const ols_step_out_t* out, void* stream);
/* queue_evaluator.py:121-131 `QueueEvaluator_impl.step(a)`: queues keyed by (z, a), no rejection — every RUNNING
 * replica with mask[r] != 0 pops the head of queue (cur_z, action[r]).  The queues must have been bucketed on the
 * combined key (z + z_offset) * n_actions + a, i.e. q->n_buckets = n_state_keys * n_actions (queue_evaluator.py:107).
 * OLS_NO_QUEUE = the reference's KeyError for a (z, a) pair that was never logged.                            */
int ols_step_action(const ols_queues_t* q, const ols_replicas_t* r, const uint8_t* mask, const int32_t* action,
                    const ols_step_out_t* out, void* stream);
/* psrs.py:99-117 `PSRS_Exo.step`: observations split into an endogenous state s and an exogenous state x; the log is
 * bucketed twice (qs keyed by s, qx keyed by x, both over the same rows; the x-queue record's z_next field holds x').
 * One pop consumes the head of both queues; the accept rule runs on the s-record; the step returns s' of the accepted
 * s-record (out->s_next) and x' of the x-record popped with it (out_x_next).  r->cursor / r->cur_z are the s-side
 * cursor table and current s; cursor_x [n_replicas][qx->n_buckets] and cur_x [n_replicas] the x side.            */
int ols_step_exo(const ols_queues_t* qs, const ols_queues_t* qx, const ols_replicas_t* r, int32_t* cursor_x, int32_t* cur_x,
                 const uint8_t* mask, const void* p_new, int32_t p_new_is_f64, int64_t p_stride, int32_t compute_mode,
                 int32_t lanes, const ols_step_out_t* out, int32_t* out_x_next, void* stream);
/* evalMC_psrs (psrs.py:241-271) over a PSRS_Exo environment as ONE launch, one independent run per replica: reset(seed =
 * episode) with the generator re-seeded (PCG64 mode), policy row pi[s * x_mult + x], lock-step pops of the s- and x-queue
 * (ols_step_exo) until accepted, returns accumulated per episode.  Ring offsets are not applied (as in ols_step_exo).      */
int ols_run_exo(const ols_queues_t* qs, const ols_queues_t* qx, const ols_replicas_t* r, int32_t* cursor_x, int32_t* cur_x,
                const ols_exo_run_t* run, void* stream);
/* psrs.py:92 (PSRS_Exo.reset), queue_evaluator.py:114 — `rejection_sampling_rng = default_rng(seed)` again: re-seed the
 * PCG64 state of the replicas with mask[r] != 0 from seeds[r] (NULL = r->seed) and restart their draw counters.     */
int ols_reseed(const ols_replicas_t* r, const uint64_t* seeds, const uint8_t* mask, void* stream);
/* psrs.py:119-271 `qlearn_psrs` / `expSARSA_psrs` / `evalMC_psrs`, one independent run per replica,
 * fused with the agent update and the per-episode return evaluator.                            */
int ols_run(const ols_queues_t* q, const ols_replicas_t* r, const ols_run_t* run, void* stream);
/* queue_evaluator.py:113-131 as a fused loop, one independent run per replica: reset, then step(a) with the agent's own
 * actions (a fed list, or sampled from a fixed policy with the replica's counter-based stream) until the episode ends, and
 * again; returns are accumulated as in evalMC_psrs (psrs.py:241-271).  The queues must be keyed by (z, a) as for
 * ols_step_action.  r->ties counts the actions taken (the index into the list / the stream), r->pops the records served.
 * Statuses: DONE (n_episodes reached), INIT_EMPTY, QUEUE_EMPTY (step returned None), NO_QUEUE (KeyError), FEED_EMPTY
 * (action list used up), RUNNING (budget).                                                                              */
int ols_run_actions(const ols_queues_t* q, const ols_replicas_t* r, const ols_queue_run_t* run, void* stream);
/* bytes of ols_run_t.sched for a run over n_replicas replicas with the given lane-group width   */
size_t ols_run_sched_bytes(int64_t n_replicas, int32_t lanes);

/* ---- evaluators --------------------------------------------------------------------------- */
size_t ols_curve_scratch_bytes(int64_t n_replicas, int32_t n_episodes);
/* Learning curve over replicas: for every episode index e < n_episodes, over the replicas with
 * ep_index[r] > e: sum[e] = Σ Gs[r][e], sumsq[e] = Σ Gs[r][e]^2, count[e] (as double) — the
 * mean ± std curve of utils/vis_utils.py:15-21 on the Gs of psrs.py:180-185.  Deterministic.    */
int ols_reduce_curves(const double* Gs, const int32_t* ep_index, int64_t n_replicas, int32_t ep_cap,
                      int32_t n_episodes, double* sum, double* sumsq, double* count, void* scratch,
                      size_t scratch_bytes, void* stream);
/* Per-replica segment reduction: ret_sum[r] = Σ_e Gs[r][e], len_sum[r] = Σ_e lengths[r][e] over
 * that replica's own episodes (segments of ragged length ep_index[r] / ep_len_count[r]).        */
int ols_replica_totals(const double* Gs, const int32_t* lengths, const int32_t* ep_index,
                       const int32_t* ep_len_count, int64_t n_replicas, int32_t ep_cap, double* ret_sum,
                       int64_t* len_sum, void* stream);

/* ---- synthetic logged data of the benchmark shape -------------------------- */
/* Row i from Philox(seed, i): z, z_next ~ U{0..n_states-1}; a ~ pb_state[z]; r ~ N(0,1);
 * terminal ~ Bernoulli(p_done); step = 0 iff i == 0 or terminal[i-1].  Writes the columns of
 * `log` (its pointers must be writable; p_log float32).                                         */
int ols_synth_log(const ols_log_t* log, const float* pb_state, int32_t n_states, uint64_t seed, double p_done,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OLS_B200_H */
