"""Tabular policies of the PSRS path (offsim4rl/agents/tabular.py:4-32).

These are the host-side, NumPy-facing forms: ``policy(Q_rows, args) -> pi[rows, nA]`` float64, the
signature ``qlearn_psrs`` takes as ``behavior_policy`` (psrs.py:122,158).  When ``qlearn_psrs`` is given
``epsilon_greedy_policy`` it does not call it per step: the same rule runs fused in the CUDA kernel
(``ols_run`` with OLS_AGENT_QLEARN).  Tie-breaks use NumPy's global generator like the reference."""
import numpy as np


def _random_argmax(x):
    best = np.flatnonzero(x == np.max(x))
    return best[0] if best.size == 1 else np.random.choice(best)


def _argmax_rows(Q):
    return np.fromiter((_random_argmax(row) for row in Q), dtype=np.int64, count=len(Q))


def uniformly_random_policy(Q, args=None):
    return np.full_like(Q, 1.0 / Q.shape[1], dtype=np.float64)


def greedy_policy(Q, args=None):
    pi = np.zeros_like(Q, dtype=np.float64)
    pi[np.arange(len(Q)), _argmax_rows(Q)] = 1
    return pi


def soft_greedy_policy(Q, args=None):
    pi = np.isclose(Q, Q.max(axis=1, keepdims=True)).astype(np.float64)
    return pi / pi.sum(axis=1, keepdims=True)


def epsilon_greedy_policy(Q, args):
    eps = args["epsilon"]
    nA = Q.shape[1]
    pi = np.ones_like(Q, dtype=np.float64) * eps / nA
    pi[np.arange(len(Q)), _argmax_rows(Q)] = 1 - eps + eps / nA
    return pi
