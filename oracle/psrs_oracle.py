"""CPU restatement of the reference's Per-State Rejection Sampling path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) — never imported by the product.

The reference keeps python lists of tuples and pops their heads; this restatement
keeps the same *semantics* on flat arrays: a stable bucketing of the log into CSR
queues (``offsets``/``order``), one cursor per queue, one cursor into the list of
initial rows.  Every function cites the reference lines it follows
(paths relative to /root/reference/offsim4rl).  The arithmetic of the accept rule
is done with NumPy scalars/arrays of the *same dtypes* the reference would see, so
promotion (float64 unless both sides are float32), NaN/inf behaviour and the
NumPy-2 weak-scalar comparison are inherited rather than re-derived.

Parity is pinned by tests/test_oracle_golden.py against vectors produced by the
unmodified reference (tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np

from . import np_random

# Per-replica status codes (shared vocabulary with include/ols_b200.h).
RUNNING = 0        # replica holds a current state and can step
DONE = 1           # the episode/step budget given to a driver was reached
QUEUE_EMPTY = 2    # evaluators/psrs.py:44-45 — the needed queue ran dry
INIT_EMPTY = 3     # evaluators/psrs.py:33-35 — no initial state left
NO_QUEUE = 4       # evaluators/psrs.py:44 raises KeyError: state never seen as a from-state


# --------------------------------------------------------------------------- bucketing
def bucket(z, n_buckets, z_offset=0):
    """Stable grouping of rows by from-state id.

    evaluators/psrs.py:26 — ``sorted(buffer, key=z)`` (python's sort is stable) followed by
    ``itertools.groupby``: queue ``z`` holds its rows in log order.
    Returns CSR ``offsets[n_buckets+1]`` and ``order[N]`` (row index at each queue slot).
    """
    b = np.asarray(z, dtype=np.int64) + z_offset
    if b.size and (b.min() < 0 or b.max() >= n_buckets):
        raise ValueError("state id outside [−z_offset, n_buckets − z_offset)")
    counts = np.bincount(b, minlength=n_buckets)
    offsets = np.zeros(n_buckets + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    order = np.argsort(b, kind="stable").astype(np.int64)
    return offsets, order


def init_rows(steps, n):
    """evaluators/psrs.py:22 — rows whose ``info['t'] == 0``, in log order.

    data.py:72 — with no ``steps`` column every row has ``t = 0``.
    """
    if steps is None:
        return np.arange(n, dtype=np.int64)
    return np.nonzero(np.asarray(steps) == 0)[0].astype(np.int64)


def apply_seed(offsets, order, init, seed):
    """evaluators/psrs.py:23,29-30 — shuffle ``init_queue`` and every queue, each with a FRESH
    ``default_rng(seed)``; the permutation therefore depends on ``(seed, len)`` only."""
    order = np.array(order, copy=True)
    init = np.array(init, copy=True)
    cache = {}

    def perm(n):
        if n not in cache:
            cache[n] = np.asarray(np_random.shuffle_permutation(seed, n), dtype=np.int64)
        return cache[n]

    if init.size:
        init = init[perm(init.size)]
    for b in range(len(offsets) - 1):
        lo, hi = int(offsets[b]), int(offsets[b + 1])
        if hi - lo > 1:
            order[lo:hi] = order[lo:hi][perm(hi - lo)]
    return order, init


# --------------------------------------------------------------------------- accept rule
def threshold(p_new, p_log, a):
    """evaluators/psrs.py:54-57 — ``p_new[a] / p_log[a] / M`` with ``M = (p_new / p_log).max()``.

    Evaluated by NumPy on the caller's dtypes: left-to-right division, NaN-propagating max.
    """
    a = int(a)
    with np.errstate(divide="ignore", invalid="ignore"):
        M = (p_new / p_log).max()
        return p_new[a] / p_log[a] / M


def rejects(u, p_new, p_log, a):
    """evaluators/psrs.py:57 — ``u > thr`` (accept on ``u <= thr``; NaN threshold accepts).

    ``u`` is a python float, ``thr`` a NumPy scalar: under NumPy >= 2 the comparison happens in
    thr's dtype (float32 when both distributions are float32)."""
    return bool(u > threshold(p_new, p_log, a))


class FeedUniforms:
    """Harness-fed uniform stream, indexed by draw count (one draw per pop, psrs.py:56)."""

    def __init__(self, values):
        self.values = np.asarray(values, dtype=np.float64)
        self.k = 0

    def random(self):
        v = float(self.values[self.k])
        self.k += 1
        return v


class Pcg64Uniforms:
    """``np.random.default_rng(seed).random()`` restated (psrs.py:20)."""

    def __init__(self, seed):
        self._g = np_random.PCG64(seed)
        self.k = 0

    def random(self):
        self.k += 1
        return self._g.random()


class FeedTies:
    """Tie-break stream for ``_random_argmax`` (agents/tabular.py:4-5): a value v in [0,1) picks
    candidate ``min(k-1, floor(v*k))``; consumed only when there is more than one candidate."""

    def __init__(self, values):
        self.values = np.asarray(values, dtype=np.float64)
        self.k = 0

    def pick(self, candidates):
        k = len(candidates)
        if k == 1:
            return int(candidates[0])
        v = float(self.values[self.k])
        self.k += 1
        return int(candidates[min(k - 1, int(v * k))])


# --------------------------------------------------------------------------- the simulator
class OraclePSRS:
    """One replica ≡ one reference ``PSRS`` instance over the (immutable, shared) bucketed log.

    evaluators/psrs.py:5-57.  ``ring_start[b]`` (default 0) rotates queue ``b`` before consumption;
    it mirrors the product's optional per-(replica, state) decorrelation and is 0 in parity runs.
    """

    def __init__(self, z, z_next, s, s_next, action, reward, terminal, p_log, steps=None,
                 n_buckets=None, z_offset=0, seed=None, uniforms=None, accept_all=False,
                 ring_start=None, init_start=0):
        self.z = np.asarray(z)
        self.z_next = np.asarray(z_next)
        self.s = np.asarray(s)
        self.s_next = np.asarray(s_next)
        self.action = np.asarray(action)
        self.reward = np.asarray(reward)
        self.terminal = np.asarray(terminal)
        self.p_log = np.asarray(p_log)
        n = self.z.shape[0]
        self.z_offset = z_offset
        if n_buckets is None:
            n_buckets = int(max(self.z.max(initial=-z_offset), self.z_next.max(initial=-z_offset))) + z_offset + 1
        self.n_buckets = n_buckets
        self.offsets, self.order = bucket(self.z, n_buckets, z_offset)
        self.init = init_rows(steps, n)
        if seed is not None:
            self.order, self.init = apply_seed(self.offsets, self.order, self.init, seed)
        self.uniforms = uniforms if uniforms is not None else Pcg64Uniforms(0 if seed is None else seed)
        self.accept_all = accept_all
        self.cursor = np.zeros(n_buckets, dtype=np.int64)
        self.ring_start = np.zeros(n_buckets, dtype=np.int64) if ring_start is None else np.asarray(ring_start, dtype=np.int64)
        self.init_cursor = 0
        self.init_start = init_start
        self.cur_z = None
        self.cur_s = None
        self.status = INIT_EMPTY
        self.pops = 0
        self.accepted = 0

    # evaluators/psrs.py:32-37
    def reset(self):
        if self.init_cursor >= self.init.size:
            self.cur_s = None
            self.status = INIT_EMPTY
            return None
        row = int(self.init[(self.init_start + self.init_cursor) % self.init.size])
        self.init_cursor += 1
        self.cur_z = int(self.z[row])
        self.cur_s = self.s[row]
        self.status = RUNNING
        return self.cur_s

    # evaluators/psrs.py:39-51
    def step(self, p_new):
        """Returns the accepted row index, or None on exhaustion (state unchanged, psrs.py:44-45)."""
        b = self.cur_z + self.z_offset
        if b < 0 or b >= self.n_buckets or self.offsets[b + 1] == self.offsets[b]:
            self.status = NO_QUEUE          # dict lookup KeyError in the reference
            raise KeyError(self.cur_z)
        lo = int(self.offsets[b])
        ln = int(self.offsets[b + 1]) - lo
        while True:
            c = int(self.cursor[b])
            if c >= ln:
                self.status = QUEUE_EMPTY
                return None
            row = int(self.order[lo + (int(self.ring_start[b]) + c) % ln])
            self.cursor[b] = c + 1
            self.pops += 1
            if self.accept_all:
                reject = False              # trivial_baselines.py:8-10,22-24
            else:
                u = self.uniforms.random()
                reject = rejects(u, p_new, self.p_log[row], self.action[row])
            if not reject:
                break
        self.cur_s = self.s_next[row]
        self.cur_z = int(self.z_next[row])
        self.accepted += 1
        return row


class OraclePSRSExo:
    """evaluators/psrs.py:59-117 — endogenous queues keyed by s, exogenous queues keyed by x, popped in lock-step;
    ``reset(seed)`` re-creates the rejection generator (psrs.py:92).  ``step`` returns ``(row_s, x_next)`` or None."""

    def __init__(self, s, x, s_next, x_next, action, p_log, steps, n_s, n_x, seed=None):
        self.s, self.x, self.s_next, self.x_next = (np.asarray(v) for v in (s, x, s_next, x_next))
        self.action, self.p_log = np.asarray(action), np.asarray(p_log)
        n = self.s.shape[0]
        self.off_s, self.ord_s = bucket(self.s, n_s)
        self.off_x, self.ord_x = bucket(self.x, n_x)
        self.init = init_rows(steps, n)
        if seed is not None:
            self.ord_s, self.init = apply_seed(self.off_s, self.ord_s, self.init, seed)         # psrs.py:79,86-87
            self.ord_x, _ = apply_seed(self.off_x, self.ord_x, np.zeros(0, dtype=np.int64), seed)  # psrs.py:88-89
        self.cur_s_tab = np.zeros(n_s, dtype=np.int64)
        self.cur_x_tab = np.zeros(n_x, dtype=np.int64)
        self.init_cursor = 0
        self.cs = self.cx = None
        self.rng = None

    def reset(self, seed):                                                                       # psrs.py:91-97
        self.rng = Pcg64Uniforms(seed)
        if self.init_cursor >= self.init.size:
            return None
        row = int(self.init[self.init_cursor])
        self.init_cursor += 1
        self.cs, self.cx = int(self.s[row]), int(self.x[row])
        return row

    def step(self, p_new):                                                                       # psrs.py:99-117
        bs, bx = self.cs, self.cx
        if self.off_s[bs + 1] == self.off_s[bs] or self.off_x[bx + 1] == self.off_x[bx]:
            raise KeyError((bs, bx))
        while True:
            cs, cx = int(self.cur_s_tab[bs]), int(self.cur_x_tab[bx])
            if cs >= self.off_s[bs + 1] - self.off_s[bs] or cx >= self.off_x[bx + 1] - self.off_x[bx]:
                return None
            row_s = int(self.ord_s[self.off_s[bs] + cs])
            row_x = int(self.ord_x[self.off_x[bx] + cx])
            self.cur_s_tab[bs] += 1
            self.cur_x_tab[bx] += 1
            if not rejects(self.rng.random(), p_new, self.p_log[row_s], self.action[row_s]):
                break
        self.cs, self.cx = int(self.s_next[row_s]), int(self.x_next[row_x])
        return row_s, self.cx


class OracleQueueEvaluator:
    """evaluators/queue_evaluator.py:90-131 — queues keyed by ``(z, a)`` (:107), the agent's own action picks the queue
    (:125-127), no rejection.  ``step`` returns the row index, None on an empty queue, KeyError for an unlogged pair."""

    def __init__(self, z, z_next, s, s_next, action, steps=None, n_state_keys=None, n_actions=None, z_offset=0, seed=None):
        self.z, self.z_next = np.asarray(z), np.asarray(z_next)
        self.s, self.s_next = np.asarray(s), np.asarray(s_next)
        self.action = np.asarray(action).astype(np.int64)
        n = self.z.shape[0]
        self.nA, self.z_offset = n_actions, z_offset
        key = (self.z.astype(np.int64) + z_offset) * n_actions + self.action
        self.offsets, self.order = bucket(key, n_state_keys * n_actions, 0)
        self.init = init_rows(steps, n)
        if seed is not None:
            self.order, self.init = apply_seed(self.offsets, self.order, self.init, seed)     # :104,110-111
        self.cursor = np.zeros(n_state_keys * n_actions, dtype=np.int64)
        self.init_cursor = 0
        self.cur_z = self.cur_s = None

    def reset(self):                                                                              # :113-119
        if self.init_cursor >= self.init.size:
            self.cur_s = None
            return None
        row = int(self.init[self.init_cursor])
        self.init_cursor += 1
        self.cur_z, self.cur_s = int(self.z[row]), self.s[row]
        return self.cur_s

    def step(self, a):                                                                            # :121-131
        b = (self.cur_z + self.z_offset) * self.nA + int(a)
        lo, hi = int(self.offsets[b]), int(self.offsets[b + 1])
        if hi == lo:
            raise KeyError((self.cur_z, int(a)))
        c = int(self.cursor[b])
        if c >= hi - lo:
            return None
        row = int(self.order[lo + c])
        self.cursor[b] = c + 1
        self.cur_s, self.cur_z = self.s_next[row], int(self.z_next[row])
        return row


# --------------------------------------------------------------------------- tabular policies
def epsilon_greedy_row(q_row, epsilon, ties):
    """agents/tabular.py:24-32 for a single row: ``eps/nA`` everywhere, ``1-eps+eps/nA`` at a
    uniformly random argmax (tabular.py:4-5)."""
    q_row = np.asarray(q_row, dtype=np.float64)
    nA = q_row.shape[0]
    pi = np.ones_like(q_row) * epsilon / nA
    cand = np.where(q_row == np.max(q_row))[0]
    a_star = ties.pick(cand)
    pi[a_star] = 1 - epsilon + epsilon / nA
    return pi


def _schedule(x):
    return x if callable(x) else (lambda episode: x)


def qlearn(env: OraclePSRS, n_episodes, gamma, alpha, epsilon, ties, nS, nA, q_init=None,
           max_steps=None, record=True, q_dtype=np.float64):
    """evaluators/psrs.py:119-185 with ``behavior_policy = epsilon_greedy_policy``.

    No terminal mask on the bootstrap (psrs.py:165,168); a partial last episode IS appended to Gs
    (psrs.py:160-162,177); ``gamma ** t`` is a fresh power (psrs.py:171).  ``max_steps`` (not in the
    reference) stops after that many accepted transitions, mirroring the product's step budget.
    ``q_dtype=np.float32``: the float32 learner — the same two update lines on a float32 table; with float32 rewards NumPy
    >= 2 evaluates them as separately rounded float32 operations with gamma and alpha rounded to float32 (what the
    reference's own lines do on a float32 table, tests/golden/make_golden.py ``F32Table``); the policy row is still formed
    from the float64 image of the Q row.
    """
    eps_f, alpha_f = _schedule(epsilon), _schedule(alpha)
    Q = np.zeros((nS, nA), dtype=q_dtype) if q_init is None else np.array(q_init, dtype=q_dtype, copy=True)
    Gs, lengths, td_errors, trace = [], [], [], []
    episode, terminate, total = 0, False, 0
    while episode < n_episodes and not terminate:
        G, t = 0, 0
        S = env.reset()
        if S is None:
            break
        S = int(S)
        done = False
        while not done:
            p = epsilon_greedy_row(np.asarray(Q[S], dtype=np.float64), eps_f(episode), ties)
            row = env.step(p)
            if row is None:
                terminate = True
                break
            A, R, S_ = int(env.action[row]), env.reward[row], int(env.s_next[row])
            done = bool(env.terminal[row])
            td = R + gamma * Q[S_].max() - Q[S, A]
            Q[S, A] = Q[S, A] + alpha_f(episode) * (R + gamma * Q[S_].max() - Q[S, A])
            if record:
                td_errors.append(td)
                trace.append(row)
            S = S_
            G = G + (gamma ** t) * R
            t += 1
            total += 1
            if max_steps is not None and total >= max_steps:
                terminate = True
                break
        Gs.append(G)
        lengths.append(t)
        episode += 1
    return Q, dict(Gs=np.array(Gs, dtype=np.float64), lengths=np.array(lengths, dtype=np.int64),
                   TD_errors=np.array(td_errors, dtype=np.float64), trace=np.array(trace, dtype=np.int64))


def expsarsa(env: OraclePSRS, n_episodes, pi, gamma, alpha, nS, nA, q_init=None, record=True):
    """evaluators/psrs.py:187-239 — fixed ``pi[nS, nA]``, target ``Q[S'] @ pi[S']`` (psrs.py:223)."""
    alpha_f = _schedule(alpha)
    pi = np.asarray(pi)
    Q = np.zeros((nS, nA)) if q_init is None else np.array(q_init, dtype=np.float64, copy=True)
    Gs, lengths, trace = [], [], []
    episode, terminate = 0, False
    while episode < n_episodes and not terminate:
        G, t = 0, 0
        S = env.reset()
        if S is None:
            break
        S = int(S)
        done = False
        while not done:
            row = env.step(pi[S])
            if row is None:
                terminate = True
                break
            A, R, S_ = int(env.action[row]), env.reward[row], int(env.s_next[row])
            done = bool(env.terminal[row])
            Q[S, A] = Q[S, A] + alpha_f(episode) * (R + gamma * (Q[S_] @ pi[S_]) - Q[S, A])
            if record:
                trace.append(row)
            S = S_
            G = G + (gamma ** t) * R
            t += 1
        Gs.append(G)
        lengths.append(t)
        episode += 1
    return Q, dict(Gs=np.array(Gs, dtype=np.float64), lengths=np.array(lengths, dtype=np.int64),
                   trace=np.array(trace, dtype=np.int64))


def evalmc(env: OraclePSRS, n_episodes, pi, gamma, record=True):
    """evaluators/psrs.py:241-271 — ``lengths`` gets every episode, ``Gs`` only completed ones
    (psrs.py:265-269)."""
    pi = np.asarray(pi)
    Gs, lengths, trace = [], [], []
    episode, terminate = 0, False
    while episode < n_episodes and not terminate:
        G, t = 0, 0
        S = env.reset()
        if S is None:
            break
        S = int(S)
        done = False
        while not done:
            row = env.step(pi[S])
            if row is None:
                terminate = True
                break
            R = env.reward[row]
            done = bool(env.terminal[row])
            if record:
                trace.append(row)
            S = int(env.s_next[row])
            G = G + (gamma ** t) * R
            t += 1
        lengths.append(t)
        if done:
            Gs.append(G)
            episode += 1
    return (np.array(Gs, dtype=np.float64), np.array(lengths, dtype=np.int64),
            np.array(trace, dtype=np.int64))


# --------------------------------------------------------------------------- encoders
_ONE_DEGREE = 0.0174532
_SIX_DEGREES = 0.1047192
_TWELVE_DEGREES = 0.2094384
_FIFTY_DEGREES = 0.87266


def cartpole_box(observations):
    """encoders/heuristic.py:20-60,65-71 — the 162-box discretiser, −1 outside the failure bounds.

    Comparisons run in the observation dtype with the python-float constants rounded to it (NumPy >= 2
    weak-scalar rule), which is what ``np.float32(x) < -2.4`` does in the reference.
    """
    obs = np.asarray(observations)
    dt = obs.dtype.type

    def c(v):
        return dt(v)

    x, xd, th, thd = obs[:, 0], obs[:, 1], obs[:, 2], obs[:, 3]
    fail = (x < c(-2.4)) | (x > c(2.4)) | (th < c(-_TWELVE_DEGREES)) | (th > c(_TWELVE_DEGREES))
    box = np.where(x < c(-0.8), 0, np.where(x < c(0.8), 1, 2))
    box = box + np.where(xd < c(-0.5), 0, np.where(xd < c(0.5), 3, 6))
    box = box + np.where(th < c(-_SIX_DEGREES), 0,
                         np.where(th < c(-_ONE_DEGREE), 9,
                                  np.where(th < c(0), 18,
                                           np.where(th < c(_ONE_DEGREE), 27,
                                                    np.where(th < c(_SIX_DEGREES), 36, 45)))))
    box = box + np.where(thd < c(-_FIFTY_DEGREES), 0, np.where(thd < c(_FIFTY_DEGREES), 54, 108))
    return np.where(fail, -1, box).astype(np.int64)


def grid_ids(observations, cells=5):
    """Grid discretiser for ``MyGridNaviCoords`` observations (envs/gridworld.py:205-211 emits
    ``cell/cells + 0.1 + U(-0.1, 0.1)``; gridworld.py:123-124 numbers states ``x + cells*y``).

    Integer math on the float32 product: ``min(int(cells*o), cells-1)`` per coordinate.
    """
    obs = np.asarray(observations, dtype=np.float32)
    k = np.floor(obs * np.float32(cells)).astype(np.int64)
    k = np.clip(k, 0, cells - 1)
    z = np.zeros(obs.shape[0], dtype=np.int64)
    mul = 1
    for d in range(obs.shape[1]):
        z += k[:, d] * mul
        mul *= cells
    return z


def _fmaf(a, b, c):
    """Correctly rounded float32 fma(a, b, c), vectorised: the product of two float32 is exact in
    float64; the float64 sum is taken to round-to-odd (TwoSum residual decides) so that the final
    rounding to float32 equals a single rounding of the exact a*b+c."""
    p = a.astype(np.float64) * b.astype(np.float64)
    c = c.astype(np.float64)
    s = p + c
    bb = s - p
    err = (p - (s - bb)) + (c - bb)
    odd = (s.view(np.int64) & 1) == 1
    need = (err != 0) & ~odd & np.isfinite(s)
    s_adj = s.copy()
    s_adj[need & (err > 0)] = np.nextafter(s[need & (err > 0)], np.inf)
    s_adj[need & (err < 0)] = np.nextafter(s[need & (err < 0)], -np.inf)
    return s_adj.astype(np.float32)


def mlp_logits(observations, W1, b1, W2, b2, negative_slope=0.01):
    """encoders/models.py:15-19 — Linear(dO,H) → LeakyReLU(0.01) → Linear(H,nZ) in float32, with the
    product's accumulation order: ``acc = bias; acc = fmaf(x_i, w_i, acc)`` for i ascending."""
    x = np.asarray(observations, dtype=np.float32)
    W1 = np.asarray(W1, dtype=np.float32)
    W2 = np.asarray(W2, dtype=np.float32)
    n, dO = x.shape
    H, nZ = W1.shape[0], W2.shape[0]
    h = np.empty((n, H), dtype=np.float32)
    for j in range(H):
        acc = np.full(n, np.float32(b1[j]), dtype=np.float32)
        for i in range(dO):
            acc = _fmaf(x[:, i], np.full(n, W1[j, i], dtype=np.float32), acc)
        h[:, j] = np.where(acc > 0, acc, acc * np.float32(negative_slope))
    out = np.empty((n, nZ), dtype=np.float32)
    for k in range(nZ):
        acc = np.full(n, np.float32(b2[k]), dtype=np.float32)
        for j in range(H):
            acc = _fmaf(h[:, j], np.full(n, W2[k, j], dtype=np.float32), acc)
        out[:, k] = acc
    return out


def mlp_ids(observations, W1, b1, W2, b2, negative_slope=0.01):
    """encoders/homer.py:159-168 — ``argmax(log_softmax(logits))`` ≡ first ``argmax(logits)``."""
    return np.argmax(mlp_logits(observations, W1, b1, W2, b2, negative_slope), axis=1).astype(np.int64)


# --------------------------------------------------------------------------- evaluators
def learning_curve(Gs_list):
    """Per-episode-index sum, sum of squares and count over replicas (the mean±std learning curve of
    utils/vis_utils.py:15-21 restated on the ``Gs`` arrays of psrs.py:180-185)."""
    E = max((len(g) for g in Gs_list), default=0)
    s = np.zeros(E)
    s2 = np.zeros(E)
    n = np.zeros(E, dtype=np.int64)
    for g in Gs_list:
        g = np.asarray(g, dtype=np.float64)
        s[:len(g)] += g
        s2[:len(g)] += g * g
        n[:len(g)] += 1
    return s, s2, n
