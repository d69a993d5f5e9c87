"""Scalar (one-replica) PSRS environment and learner drivers with the reference's API.

Mirrors offsim4rl/evaluators/psrs.py: ``PSRS(buffer, nS, nA, reject_func)`` with ``reset_sampler`` /
``reset`` / ``step`` (psrs.py:5-57) and the drivers ``qlearn_psrs`` / ``expSARSA_psrs`` / ``evalMC_psrs``
(psrs.py:119-271).  Every simulation step runs on the GPU (``ols_step`` with one replica, PCG64 replay
so that ``reset_sampler(seed=k)`` reproduces the reference's own run bit for bit); the drivers run the
whole learning loop in ONE kernel launch (``ols_run``) when given the stock policies.  The batched
vector form lives in ``offsim4rl_b200.BatchedPSRS``.
"""
from __future__ import annotations

import secrets

import numpy as np
import torch

from .. import _abi as abi
from ..agents import tabular as _tabular
from ..engine import BatchedPSRS, BucketedLog

_REJECT_NAMES = {None: "default", "default": "default", "accept_all": "accept_all"}


def _is_int_like(x):
    return isinstance(x, (int, np.integer)) or (isinstance(x, np.ndarray) and x.ndim == 0 and np.issubdtype(x.dtype, np.integer))


class PSRS(object):
    """Rejection sampler that acts as an environment (reference: psrs.py:5-57).

    ``buffer`` holds legacy tuples ``(s, a, r, s', done, p, info)`` with ``info['z']``, ``info['z_next']``
    (or ``'next_z'``) and ``info['t']`` (psrs.py:8,17,22).  ``reject_func``: ``None``/``"default"`` for psrs.py:53-57 and
    ``"accept_all"`` for the trivial baselines (trivial_baselines.py:8-10) run inside the CUDA step; any other callable
    ``reject_func(p_new, p_log, a) -> bool`` (psrs.py:11,47) is honoured through a HOST-driven loop — one single-record pop
    on the GPU per call of the function — which is API compatibility, not a fast path (and keeps such an env out of the
    fused drivers).  ``env.rejection_sampling_rng`` is then the reference's ``default_rng(seed)`` for the callable to use.
    """

    def __init__(self, buffer, nS=25, nA=5, reject_func=None, device=None):
        if not (callable(reject_func) or reject_func in _REJECT_NAMES):
            raise ValueError("reject_func must be None, 'default', 'accept_all' or a callable (p_new, p_log, a) -> bool")
        rows = list(buffer)                               # the reference also materialises it (psrs.py:17)
        n = len(rows)
        z = np.fromiter((r[6]["z"] for r in rows), dtype=np.int64, count=n)
        z_next = np.fromiter((r[6]["z_next"] if "z_next" in r[6] else r[6]["next_z"] for r in rows), dtype=np.int64, count=n)
        t = np.fromiter((r[6]["t"] for r in rows), dtype=np.int64, count=n)
        p_log = np.stack([np.asarray(r[5]) for r in rows]) if n else np.zeros((0, nA), dtype=np.float32)
        self._init_from_columns(
            z=z, z_next=z_next, action=np.array([r[1] for r in rows]), reward=np.array([r[2] for r in rows]),
            terminal=np.array([bool(r[4]) for r in rows]), p_log=p_log, steps=t,
            s=[r[0] for r in rows], s_next=[r[3] for r in rows], nS=nS, nA=nA,
            reject=reject_func if callable(reject_func) else _REJECT_NAMES[reject_func], device=device)

    @classmethod
    def from_arrays(cls, z, z_next, action, reward, terminal, p_log, steps=None, s=None, s_next=None, nS=25, nA=None,
                    reject="default", device=None):
        """Build from column arrays directly (no per-row Python objects)."""
        self = cls.__new__(cls)
        p_log = np.asarray(p_log)
        self._init_from_columns(z=z, z_next=z_next, action=np.asarray(action),
                                reward=np.asarray(reward), terminal=np.asarray(terminal), p_log=p_log, steps=steps,
                                s=s, s_next=s_next, nS=nS, nA=nA if nA is not None else p_log.shape[1], reject=reject,
                                device=device)
        return self

    def _init_from_columns(self, z, z_next, action, reward, terminal, p_log, steps, s, s_next, nS, nA, reject, device):
        self.nS, self.nA = nS, nA
        n = len(z)
        if p_log.ndim != 2 or (n and p_log.shape[1] != nA):
            raise ValueError("logged action distributions must be [n_rows, nA]")
        self._action, self._reward, self._terminal, self._p_log = action, reward, terminal, p_log
        # state ids may arrive as device tensors (an encoder's encode_device): bucket them in place and keep a
        # host copy for the scalar API's bookkeeping
        to_host = lambda a: (a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)).astype(np.int64)
        z_in, z_next_in = z, z_next
        self._z, self._z_next = to_host(z), to_host(z_next)
        # observations handed back to the caller are the original python objects; the device keeps an integer
        # copy when they are integers (the tabular drivers index Q with them, psrs.py:158,165)
        self._s_rows = self._z if s is None else s
        self._s_next_rows = self._z_next if s_next is None else s_next
        s_int = s_next_int = None
        if s is not None and n and all(_is_int_like(x) for x in s) and all(_is_int_like(x) for x in s_next):
            s_int = np.fromiter((int(x) for x in s), dtype=np.int64, count=n)
            s_next_int = np.fromiter((int(x) for x in s_next), dtype=np.int64, count=n)
        self._log = BucketedLog(z=z_in if isinstance(z_in, torch.Tensor) else self._z,
                                z_next=z_next_in if isinstance(z_next_in, torch.Tensor) else self._z_next,
                                action=action, reward=reward, terminal=terminal,
                                p_log=p_log, s=s_int, s_next=s_next_int, steps=steps, device=device)
        # a Python callable decides on the host, one popped record at a time; the device then serves single pops
        self._host_reject = reject if callable(reject) else None
        self._reject = "accept_all" if callable(reject) else reject
        self.rejection_sampling_rng = None
        self._sim = None
        self.s = self.z = None
        self.reset_sampler()
        self.reset()

    # ------------------------------------------------------------------ psrs.py:19-30
    def reset_sampler(self, seed=None):
        if seed is None:
            seed = secrets.randbits(63)          # default_rng(None): fresh OS entropy (psrs.py:13,20)
        self._seed = int(seed)
        if self._host_reject is not None:
            self.rejection_sampling_rng = np.random.default_rng(self._seed)       # psrs.py:20, for the callable's own use
        self._cur = self._log.seeded(self._seed)
        self._init_rows_host = self._cur.init_row[: self._cur.n_init].cpu().numpy()
        if self._sim is None:
            self._sim = BatchedPSRS(self._cur, 1, rng="pcg64", seed=self._seed, reject=self._reject, lanes=4)
        else:
            self._sim.seed = self._seed
            self._sim.u_feed = None
            self._sim.rng_mode = abi.RNG_PCG64
            if self._sim.pcg_state is None:
                self._sim.pcg_state = torch.zeros((1, 4), dtype=torch.int64, device=self._sim.device)
            self._sim.restart(self._cur)
            if self.z is not None:
                # the reference keeps the current (z, s) across reset_sampler (psrs.py:19-30 do not touch them)
                self._sim.cur_z.fill_(int(self.z))
                if _is_int_like(self.s):
                    self._sim.cur_s.fill_(int(self.s))
                self._sim.status.fill_(abi.RUNNING)
        self._n_resets = 0

    def feed_uniforms(self, values):
        """Parity hook: replace the rejection uniforms by a given stream (what a harness does to the
        reference with ``env.rejection_sampling_rng = Feed(values)``); draw k uses ``values[k]``."""
        self._sim.u_feed = torch.as_tensor(np.asarray(values, dtype=np.float64), device=self._sim.device)
        self._sim.rng_mode = abi.RNG_FEED
        self._sim.draws.zero_()

    def feed_ties(self, values):
        """Parity hook for the fused Q-learning driver: tie-break k of ``_random_argmax`` (tabular.py:4-5) picks candidate
        ``min(n - 1, int(values[k] * n))`` — what a harness does to the reference by patching ``np.random.choice``."""
        self._sim.tie_feed = torch.as_tensor(np.asarray(values, dtype=np.float64), device=self._sim.device)
        self._sim.tie_mode = abi.RNG_FEED
        self._sim.ties.zero_()

    # ------------------------------------------------------------------ psrs.py:32-37
    def reset(self, seed=None):
        _, status = self._sim.reset()
        if int(status[0]) != abi.RUNNING:
            self.s = None
            return None
        row = int(self._init_rows_host[self._n_resets])
        self._n_resets += 1
        self.z = int(self._z[row])
        self.s = self._s_rows[row]
        return self.s

    # ------------------------------------------------------------------ psrs.py:39-51
    def step(self, p_new):
        p = np.asarray(p_new)
        if p.dtype not in (np.float32, np.float64):
            p = p.astype(np.float64)
        while True:
            out = self._sim.step(p)               # with a host callable: ONE record popped, accepted unconditionally
            row = int(out["row"][0])
            if row < 0:
                status = int(self._sim.status[0])
                if status == abi.NO_QUEUE:
                    raise KeyError(self.z)        # psrs.py:44 — queues[z] for a z never logged as a from-state
                if status == abi.FEED_EMPTY:
                    raise IndexError("the fed uniform stream is exhausted")
                return None, None, None, None     # psrs.py:44-45
            if self._host_reject is None or not self._host_reject(p_new, self._p_log[row], self._action[row]):
                break
            # psrs.py:46-48 — rejected: the record is gone, the state is unchanged; pop again
            self._sim.cur_z.fill_(int(self.z))
            if _is_int_like(self.s):
                self._sim.cur_s.fill_(int(self.s))
            self._sim.accepted -= 1
        z_from = self.z
        self.s = self._s_next_rows[row]
        self.z = int(self._z_next[row])
        return self.s, self._reward[row], self._terminal[row], {"z": z_from, "a": self._action[row], "p": self._p_log[row]}

    # ------------------------------------------------------------------ introspection (public attrs in the reference)
    @property
    def queues(self):
        """``{z: [row indices still queued]}`` (the reference's ``env.queues``, psrs.py:26)."""
        off = self._cur.offsets.cpu().numpy()
        order = self._cur.order.cpu().numpy()
        cur = self._sim.cursor[0].cpu().numpy()
        out = {}
        for b in range(self._cur.n_buckets):
            if off[b + 1] > off[b]:
                out[b - self._cur.z_offset] = order[off[b] + cur[b]: off[b + 1]].tolist()
        return out

    @property
    def init_queue(self):
        return [(int(self._z[r]), self._s_rows[r]) for r in self._init_rows_host[self._n_resets:]]

    @property
    def pops(self):
        return int(self._sim.pops[0])

    def _sync_from_device(self):
        self._n_resets = int(self._sim.init_cursor[0])
        self.z = int(self._sim.cur_z[0])
        st = int(self._sim.status[0])
        self.s = None if st == abi.INIT_EMPTY else int(self._sim.cur_s[0])
        if st == abi.DONE:
            # the reference's env is an ordinary env after a driver returns: a scalar step() continues from the last state
            self._sim.status.fill_(abi.RUNNING)


class PSRS_Exo(object):
    """Rejection sampler with an exogenous state (reference: psrs.py:59-117).

    An observation ``o`` splits into an endogenous state ``s`` and an exogenous state ``x`` (``o_split_func``); the
    log is bucketed by ``s`` and, independently, by ``x``; one pop consumes the head of both queues and the usual
    accept rule runs on the s-record; the next observation is ``o_combine_func(s', x')``.  ``s`` and ``x`` must be
    integers.  ``buffer`` rows are ``(o, a, r, o', done, p, info)`` with ``info['t']``.
    """

    def __init__(self, buffer, nO=25, nA=5, o_split_func=lambda o: (o, 0), o_combine_func=lambda s, x: s, device=None):
        rows = list(buffer)
        n = len(rows)
        self.nS = self.nO = nO
        self.nA = nA
        self.o_split_func, self.o_combine_func = o_split_func, o_combine_func
        split = [o_split_func(r[0]) for r in rows]
        split_next = [o_split_func(r[3]) for r in rows]
        s = np.array([int(a) for a, _ in split], dtype=np.int64)
        x = np.array([int(b) for _, b in split], dtype=np.int64)
        s2 = np.array([int(a) for a, _ in split_next], dtype=np.int64)
        x2 = np.array([int(b) for _, b in split_next], dtype=np.int64)
        self._o_rows = [r[0] for r in rows]
        self._x = x
        self._action = np.array([r[1] for r in rows])
        self._reward = np.array([r[2] for r in rows])
        self._terminal = np.array([bool(r[4]) for r in rows])
        self._p_log = np.stack([np.asarray(r[5]) for r in rows]) if n else np.zeros((0, nA), dtype=np.float32)
        t = np.fromiter((r[6]["t"] for r in rows), dtype=np.int64, count=n)
        common = dict(action=self._action, reward=self._reward, terminal=self._terminal, p_log=self._p_log, steps=t, device=device)
        self._log_s = BucketedLog(z=s, z_next=s2, **common)      # psrs.py:82
        self._log_x = BucketedLog(z=x, z_next=x2, **common)      # psrs.py:83
        self._sim = None
        self.o = self.s = None
        self.reset_sampler()
        self.reset()

    def reset_sampler(self, seed=None):                                     # psrs.py:77-89
        if seed is None:
            seed = secrets.randbits(63)
        self._cur_s, self._cur_x = self._log_s.seeded(int(seed)), self._log_x.seeded(int(seed))
        self._init_rows_host = self._cur_s.init_row[: self._cur_s.n_init].cpu().numpy()
        if self._sim is None:
            self._sim = BatchedPSRS(self._cur_s, 1, rng="pcg64", seed=int(seed), lanes=4)
            dev = self._sim.device
            self._cursor_x = torch.zeros((1, self._log_x.n_buckets), dtype=torch.int32, device=dev)
            self._x_state = torch.zeros(1, dtype=torch.int32, device=dev)
            self._out_x = torch.zeros(1, dtype=torch.int32, device=dev)
        else:
            self._sim.restart(self._cur_s)
            self._cursor_x.zero_()
            if self.o is not None:
                s_now, x_now = self.o_split_func(self.o)
                self._sim.cur_z.fill_(int(s_now))
                self._x_state.fill_(int(x_now))
                self._sim.status.fill_(abi.RUNNING)
        self._n_resets = 0

    def reset(self, seed=None):                                             # psrs.py:91-97
        import ctypes as C
        from ..engine import _ptr, _stream_ptr
        if seed is None:
            seed = secrets.randbits(63)
        sim = self._sim
        sim.seed = int(seed)
        with torch.cuda.device(sim.device):
            r = sim.replicas_struct()
            abi.check(sim.lib.ols_reseed(C.byref(r), _ptr(None), _ptr(None), _stream_ptr()))   # default_rng(seed) again
        _, status = sim.reset()
        if int(status[0]) != abi.RUNNING:
            self.s = None
            return None
        row = int(self._init_rows_host[self._n_resets])
        self._n_resets += 1
        self.o = self._o_rows[row]
        self._x_state.fill_(int(self._x[row]))
        return self.o

    def step(self, p_new):                                                  # psrs.py:99-117
        import ctypes as C
        from ..engine import _as_tensor, _ptr, _stream_ptr
        sim = self._sim
        with torch.cuda.device(sim.device):
            p = _as_tensor(np.asarray(p_new), sim.device)
            if p.dtype not in (torch.float32, torch.float64):
                p = p.to(torch.float64)
            compute = abi.COMPUTE_F32 if (p.dtype == torch.float32 and self._log_s.p_source_f32) else abi.COMPUTE_F64
            qs, qx, r = self._cur_s.queues_struct(), self._cur_x.queues_struct(), sim.replicas_struct()
            out = abi.StepOut(row=_ptr(sim.out_row), action=_ptr(sim.out_action), reward=_ptr(sim.out_reward),
                              terminal=_ptr(sim.out_terminal), s_next=_ptr(sim.out_s_next), z_from=_ptr(sim.out_z_from))
            abi.check(sim.lib.ols_step_exo(C.byref(qs), C.byref(qx), C.byref(r), _ptr(self._cursor_x), _ptr(self._x_state),
                                           _ptr(None), _ptr(p), int(p.dtype == torch.float64), 0, compute, sim.lanes,
                                           C.byref(out), _ptr(self._out_x), _stream_ptr()))
        row = int(sim.out_row[0])
        if row < 0:
            if int(sim.status[0]) == abi.NO_QUEUE:
                raise KeyError(self.o_split_func(self.o))
            return None, None, None, None
        s_from = int(sim.out_z_from[0])
        self.o = self.o_combine_func(int(sim.out_s_next[0]), int(self._out_x[0]))
        return self.o, self._reward[row], self._terminal[row], {"s": s_from, "a": self._action[row], "p": self._p_log[row]}


# --------------------------------------------------------------------------------------------------- drivers
def _schedule(x):
    return x if callable(x) else (lambda episode: x)


def _fusable(env):
    return isinstance(env, PSRS) and env._host_reject is None


def _rows_to_memory(env, rows):
    return [(int(env._z[r]) if env._s_rows is env._z else env._s_rows[r], env._action[r], env._reward[r],
             env._s_next_rows[r], env._terminal[r], None, {"z": int(env._z[r]), "a": env._action[r], "p": env._p_log[r]})
            for r in rows]


def _raise_on_error(env, res):
    st = int(res.status[0])
    if st == abi.NO_QUEUE:
        raise KeyError("a simulated transition led to a state that was never logged as a from-state (psrs.py:44)")
    if st == abi.FEED_EMPTY:
        raise IndexError("the fed uniform / tie stream is exhausted")


def qlearn_psrs(env, n_episodes, behavior_policy, gamma, alpha=0.1, epsilon=1.0, Q_init=None, save_Q=0):
    """Tabular Q-learning against a PSRS environment (reference: psrs.py:119-185).

    With ``behavior_policy = agents.tabular.epsilon_greedy_policy`` and ``save_Q == 0`` the whole run is one
    fused kernel launch (float64 Q, the reference's operation order); any other policy callable falls back to
    a host-driven loop whose every ``env.step`` is still the CUDA step.  As in the reference, every call counts its
    episodes from 0 and keeps consuming the queues where the previous call (or scalar stepping) left them.  Rewards
    travel as float32 inside the packed records: the fused path computes Q / TD / Gs from the float32 image of a
    float64 reward (~1e-8 relative, inside the 1e-5 contract); ``memory`` tuples carry ``p = None``."""
    if _fusable(env) and behavior_policy is _tabular.epsilon_greedy_policy and not save_Q:
        Q0 = np.zeros((env.nS, env.nA)) if Q_init is None else np.asarray(Q_init, dtype=np.float64)
        Qd = torch.as_tensor(Q0.reshape(1, env.nS, env.nA).copy(), device=env._sim.device)
        cap = max(1, env._log.n_rows)
        first = int(env._sim.accepted[0])
        env._sim.begin_driver_call()
        res = env._sim.run_qlearn(env.nS, n_episodes, gamma, alpha=alpha, epsilon=epsilon, Q=Qd, trace_cap=cap)
        _raise_on_error(env, res)
        env._sync_from_device()
        n = int(res.n_returns[0])
        rows = res.trace_row[0, first:int(res.accepted[0])].cpu().numpy()
        return Qd[0].cpu().numpy(), {
            "Gs": res.Gs[0, :n].cpu().numpy(), "Qs": np.array([Q0]),
            "TD_errors": res.td_error[0, first:int(res.accepted[0])].cpu().numpy(), "memory": _rows_to_memory(env, rows)}

    eps_f, alpha_f = _schedule(epsilon), _schedule(alpha)
    Q = np.zeros((env.nS, env.nA)) if Q_init is None else np.array(Q_init, dtype=float)
    out = dict(Gs=[], Qs=[Q.copy()], TD_errors=[], memory=[])
    for episode in range(n_episodes):
        S = env.reset(seed=episode)
        if S is None:
            break
        G, discount, exhausted, done = 0, 1.0, False, False
        while not done:
            p = behavior_policy(Q[[S], :], dict(epsilon=eps_f(episode)))[0]
            S2, R, done, info = env.step(p)
            if S2 is None:
                exhausted = True
                break
            A = info["a"]
            delta = R + gamma * Q[S2].max() - Q[S, A]
            out["memory"].append((S, A, R, S2, done, p, info))
            out["TD_errors"].append(delta)
            Q[S, A] += alpha_f(episode) * delta
            G += discount * R
            discount *= gamma
            S = S2
            if save_Q:
                out["Qs"].append(Q.copy())
        out["Gs"].append(G)                        # a partial last episode is appended too (psrs.py:160-162,177)
        if exhausted:
            break
    return Q, {"Gs": np.array(out["Gs"]), "Qs": np.array(out["Qs"]), "TD_errors": np.array(out["TD_errors"]),
               "memory": out["memory"]}


def expSARSA_psrs(env, n_episodes, pi, gamma, alpha=0.1, Q_init=None, save_Q=0):
    """Expected SARSA for a fixed policy ``pi[nS, nA]`` (reference: psrs.py:187-239)."""
    pi = np.asarray(pi)
    if _fusable(env) and not save_Q:
        Q0 = np.zeros((env.nS, env.nA)) if Q_init is None else np.asarray(Q_init, dtype=np.float64)
        Qd = torch.as_tensor(Q0.reshape(1, env.nS, env.nA).copy(), device=env._sim.device)
        first = int(env._sim.accepted[0])
        env._sim.begin_driver_call()
        res = env._sim.run_expsarsa(env.nS, n_episodes, pi, gamma, alpha=alpha, Q=Qd, trace_cap=max(1, env._log.n_rows))
        _raise_on_error(env, res)
        env._sync_from_device()
        n = int(res.n_returns[0])
        rows = res.trace_row[0, first:int(res.accepted[0])].cpu().numpy()
        return Qd[0].cpu().numpy(), {"Gs": res.Gs[0, :n].cpu().numpy(), "Qs": np.array([Q0]),
                                     "memory": _rows_to_memory(env, rows)}

    alpha_f = _schedule(alpha)
    Q = np.zeros((env.nS, env.nA)) if Q_init is None else np.array(Q_init, dtype=float)
    Gs, Qs, memory = [], [Q.copy()], []
    for episode in range(n_episodes):
        S = env.reset(seed=episode)
        if S is None:
            break
        G, discount, exhausted, done = 0, 1.0, False, False
        while not done:
            S2, R, done, info = env.step(pi[S])
            if S2 is None:
                exhausted = True
                break
            A = info["a"]
            memory.append((S, A, R, S2, done, pi[S], info))
            Q[S, A] += alpha_f(episode) * (R + gamma * (Q[S2] @ pi[S2]) - Q[S, A])
            G += discount * R
            discount *= gamma
            S = S2
            if save_Q:
                Qs.append(Q.copy())
        Gs.append(G)
        if exhausted:
            break
    return Q, {"Gs": np.array(Gs), "Qs": np.array(Qs), "memory": memory}


def evalMC_psrs(env, n_episodes, pi, gamma):
    """Monte-Carlo evaluation of a fixed policy: returns of COMPLETED episodes and the lengths of all
    simulated ones (reference: psrs.py:241-271)."""
    pi = np.asarray(pi)
    if _fusable(env):
        env._sim.begin_driver_call()
        res = env._sim.run_evalmc(env.nS, n_episodes, pi, gamma)
        _raise_on_error(env, res)
        env._sync_from_device()
        return (res.Gs[0, :int(res.n_returns[0])].cpu().numpy(),
                res.lengths[0, :int(res.n_lengths[0])].cpu().numpy().astype(np.int64))
    Gs, lengths = [], []
    while len(Gs) < n_episodes:
        S = env.reset(seed=len(Gs))
        if S is None:
            break
        G, discount, t, done = 0, 1.0, 0, False
        while not done:
            S2, R, done, info = env.step(pi[S])
            if S2 is None:
                break
            G += discount * R
            discount *= gamma
            t += 1
            S = S2
        lengths.append(t)
        if not done:
            break
        Gs.append(G)
    return np.array(Gs), np.array(lengths)
