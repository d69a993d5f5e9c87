"""ctypes front-end of oracle/ols_oracle.c (the C restatement).  TEST INFRASTRUCTURE ONLY.

Used for (i) parity checks at sizes the pure-python oracle cannot reach (it shares the Philox and
ring-offset definitions with the CUDA path, so CPU and GPU replicas can be compared draw for draw) and
(ii) the multi-core CPU baseline of bench.py.  Never imported by the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libols_oracle.so")

FIXED_PI, EXPSARSA, QLEARN = 0, 1, 2

p, i32, i64, u64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double


class Sim(C.Structure):
    _fields_ = [
        ("n_rows", i64), ("n_actions", i32), ("n_buckets", i32), ("z_offset", i32), ("p_is_f64", i32),
        ("z", p), ("z_next", p), ("s", p), ("s_next", p), ("action", p), ("reward", p), ("terminal", p), ("p_log", p),
        ("offsets", p), ("order", p), ("init_row", p), ("n_init", i64),
        ("n_replicas", i64), ("replica_id0", i64), ("rng_mode", i32), ("tie_mode", i32), ("seed", u64),
        ("u_feed", p), ("u_feed_len", i64), ("u_stride", i64), ("tie_feed", p), ("tie_feed_len", i64), ("tie_stride", i64),
        ("ring_mode", i32), ("ring_seed", u64),
        ("agent", i32), ("n_states", i32), ("accept_all", i32), ("compute_f32", i32), ("pi_is_f64", i32),
        ("n_episodes", i32), ("ep_cap", i32), ("max_steps", i64), ("gamma", f64), ("alpha", p), ("epsilon", p), ("pi", p),
        ("Q", p), ("Gs", p), ("lengths", p), ("n_returns", p), ("n_lengths", p), ("status", p),
        ("accepted", p), ("pops", p), ("draws", p), ("ties", p), ("trace_row", p), ("td_error", p), ("trace_cap", i64),
        ("trace_hash", p), ("q_is_f32", i32), ("_pad0", i32), ("replica_ids", p),
    ]


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "ols_oracle.c")):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        lib = C.CDLL(LIB)
        lib.oracle_run.restype = C.c_int
        lib.oracle_run.argtypes = [C.POINTER(Sim), C.c_int]
        lib.oracle_bucket.restype = i64
        lib.oracle_bucket.argtypes = [p, i64, i32, i32, p, p]
        lib.oracle_init_rows.restype = i64
        lib.oracle_init_rows.argtypes = [p, i64, p]
        lib.oracle_sim_bytes.restype = C.c_size_t
        lib.oracle_mlp_ids.restype = None
        lib.oracle_mlp_ids.argtypes = [p, i64, C.c_int, p, p, C.c_int, p, p, C.c_int, C.c_float, p]
        assert lib.oracle_sim_bytes() == C.sizeof(Sim), "oracle struct layout mismatch"
        _lib = lib
    return _lib


def _c(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def _p(a):
    return C.c_void_p(0 if a is None else a.ctypes.data)


class Log:
    """Log columns + log-order CSR queues, built by the C restatement."""

    def __init__(self, z, z_next, action, reward, terminal, p_log, s=None, s_next=None, steps=None, n_buckets=None,
                 z_offset=0):
        lib = load()
        self.z, self.z_next = _c(z, np.int32), _c(z_next, np.int32)
        self.s, self.s_next = _c(s, np.int32), _c(s_next, np.int32)
        self.action = _c(np.asarray(action).astype(np.int64), np.int32)
        self.reward, self.terminal = _c(reward, np.float32), _c(np.asarray(terminal).astype(np.uint8), np.uint8)
        p_log = np.asarray(p_log)
        self.p_is_f64 = int(p_log.dtype == np.float64)
        self.p_source_f32 = p_log.dtype == np.float32
        self.p_log = _c(p_log, np.float64 if self.p_is_f64 else np.float32)
        self.steps = _c(steps, np.int32)
        n = self.z.shape[0]
        self.n_rows, self.n_actions, self.z_offset = n, self.p_log.shape[1], z_offset
        if n_buckets is None:
            n_buckets = int(max(self.z.max(initial=0), self.z_next.max(initial=0))) + z_offset + 1
        self.n_buckets = n_buckets
        self.offsets = np.zeros(n_buckets + 1, dtype=np.int32)
        self.order = np.zeros(max(n, 1), dtype=np.int32)
        bad = lib.oracle_bucket(_p(self.z), n, n_buckets, z_offset, _p(self.offsets), _p(self.order))
        assert bad == 0, "state id out of range"
        self.init_row = np.zeros(max(n, 1), dtype=np.int32)
        self.n_init = int(lib.oracle_init_rows(_p(self.steps), n, _p(self.init_row)))


def run(log: Log, n_replicas, agent, n_states, n_episodes, gamma, alpha=None, epsilon=None, pi=None, seed=0,
        replica_id0=0, u_feed=None, tie_feed=None, ring=False, ring_seed=0, accept_all=False, compute_f32=None,
        max_steps=0, trace_cap=0, ep_cap=None, threads=1, order=None, init_row=None, want_q=True, q_f32=False,
        replica_ids=None):
    """Run ``n_replicas`` fresh replicas; returns a dict of numpy arrays (+ ``threads`` actually used).

    ``q_f32``: the float32 learner of the GPU production path (Q entries are float32 values, the update is four
    separately rounded float32 operations); ``Q`` is then returned as float32.  ``replica_ids``: global ids of the
    replicas to run (a sample of a larger fleet), default ``replica_id0 + arange(n_replicas)``."""
    lib = load()
    R, nA = int(n_replicas), log.n_actions
    ep_cap = int(ep_cap or n_episodes + 1)

    def sched(x):
        if x is None:
            return None
        if callable(x):
            return np.array([float(x(e)) for e in range(n_episodes + 1)])
        if np.ndim(x) == 0:
            return np.full(n_episodes + 1, float(x))
        x = np.asarray(x, dtype=np.float64)
        return np.concatenate([x, np.full(max(0, n_episodes + 1 - x.size), x[-1])])

    alpha, epsilon = sched(alpha), sched(epsilon)
    pi_a, pi_is_f64 = None, 0
    if pi is not None:
        pi = np.asarray(pi)
        pi_is_f64 = int(pi.dtype != np.float32)
        pi_a = _c(pi, np.float64 if pi_is_f64 else np.float32)
    if compute_f32 is None:
        compute_f32 = bool(pi is not None and not pi_is_f64 and log.p_source_f32)
    u_feed, tie_feed = _c(u_feed, np.float64), _c(tie_feed, np.float64)

    def feed_shape(a):
        if a is None:
            return 0, 0
        return (a.shape[0], 0) if a.ndim == 1 else (a.shape[1], a.shape[1])

    ulen, ustride = feed_shape(u_feed)
    tlen, tstride = feed_shape(tie_feed)
    rid_a = None if replica_ids is None else _c(replica_ids, np.int64)
    if rid_a is not None and rid_a.shape[0] != R:
        raise ValueError("replica_ids must hold n_replicas ids")
    order = log.order if order is None else _c(order, np.int32)
    init_row = log.init_row if init_row is None else _c(init_row, np.int32)
    out = dict(
        Q=np.zeros((R, n_states, nA)) if (agent != FIXED_PI and want_q) else None,
        Gs=np.zeros((R, ep_cap)), lengths=np.zeros((R, ep_cap), dtype=np.int32),
        n_returns=np.zeros(R, dtype=np.int32), n_lengths=np.zeros(R, dtype=np.int32), status=np.zeros(R, dtype=np.int32),
        accepted=np.zeros(R, dtype=np.int64), pops=np.zeros(R, dtype=np.int64), draws=np.zeros(R, dtype=np.int64),
        ties=np.zeros(R, dtype=np.int64),
        trace_row=np.full((R, trace_cap), -1, dtype=np.int32) if trace_cap else None,
        td_error=np.zeros((R, trace_cap)) if (trace_cap and agent == QLEARN) else None,
        trace_hash=np.zeros(R, dtype=np.uint64))
    if agent != FIXED_PI and not want_q:
        raise ValueError("learners need Q")
    S = Sim(n_rows=log.n_rows, n_actions=nA, n_buckets=log.n_buckets, z_offset=log.z_offset, p_is_f64=log.p_is_f64,
            z=_p(log.z), z_next=_p(log.z_next), s=_p(log.s), s_next=_p(log.s_next), action=_p(log.action),
            reward=_p(log.reward), terminal=_p(log.terminal), p_log=_p(log.p_log), offsets=_p(log.offsets),
            order=_p(order), init_row=_p(init_row), n_init=log.n_init, n_replicas=R, replica_id0=replica_id0,
            rng_mode=0 if u_feed is not None else 2, tie_mode=0 if tie_feed is not None else 2, seed=seed,
            u_feed=_p(u_feed), u_feed_len=ulen, u_stride=ustride, tie_feed=_p(tie_feed), tie_feed_len=tlen,
            tie_stride=tstride, ring_mode=int(ring), ring_seed=ring_seed, agent=agent, n_states=n_states,
            accept_all=int(accept_all), compute_f32=int(compute_f32), pi_is_f64=pi_is_f64, n_episodes=n_episodes,
            ep_cap=ep_cap, max_steps=max_steps, gamma=gamma, alpha=_p(alpha), epsilon=_p(epsilon), pi=_p(pi_a),
            Q=_p(out["Q"]), Gs=_p(out["Gs"]), lengths=_p(out["lengths"]), n_returns=_p(out["n_returns"]),
            n_lengths=_p(out["n_lengths"]), status=_p(out["status"]), accepted=_p(out["accepted"]), pops=_p(out["pops"]),
            draws=_p(out["draws"]), ties=_p(out["ties"]), trace_row=_p(out["trace_row"]), td_error=_p(out["td_error"]),
            trace_cap=trace_cap, trace_hash=_p(out["trace_hash"]), q_is_f32=int(bool(q_f32)), _pad0=0, replica_ids=_p(rid_a))
    out["threads"] = lib.oracle_run(C.byref(S), int(threads))
    if q_f32 and out["Q"] is not None:
        q32 = out["Q"].astype(np.float32)
        assert np.array_equal(q32.astype(np.float64), out["Q"]), "float32 learner left a non-float32 Q entry"
        out["Q"] = q32
    return out


def trace_hash(rows):
    """The order-sensitive FNV-1a style hash the C oracle keeps per replica, for a row sequence."""
    h = 0xcbf29ce484222325
    for r in rows:
        h = ((h ^ (int(r) & 0xFFFFFFFF)) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def mlp_ids(x, W1, b1, W2, b2, slope=0.01):
    lib = load()
    x, W1, b1, W2, b2 = (_c(a, np.float32) for a in (x, W1, b1, W2, b2))
    ids = np.zeros(x.shape[0], dtype=np.int32)
    lib.oracle_mlp_ids(_p(x), x.shape[0], x.shape[1], _p(W1), _p(b1), W1.shape[0], _p(W2), _p(b2), W2.shape[0],
                       C.c_float(slope), _p(ids))
    return ids.astype(np.int64)
