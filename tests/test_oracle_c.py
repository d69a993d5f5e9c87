"""The C restatement (oracle/ols_oracle.c) against the python oracle and the reference's golden vectors."""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import psrs_oracle as O
from oracle import np_random

DRIVER_CASES = ["drivers_f64", "drivers_f32", "drivers_p64_a5", "drivers_edge"]


def _clog(g, seed=None):
    steps = g["t"] if ("with_t" not in g.files or bool(g["with_t"])) else None
    nS = g["pi"].shape[0]
    log = CO.Log(z=g["z"], z_next=g["z_next"], action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"],
                 steps=steps, n_buckets=nS)
    order, init = log.order, log.init_row
    if seed is not None:
        order, init = O.apply_seed(log.offsets, log.order[:log.n_rows], log.init_row[:log.n_init], seed)
    return log, order, init


@pytest.mark.parametrize("case", DRIVER_CASES)
@pytest.mark.parametrize("tag", ["const", "sched"])
def test_c_qlearn_matches_reference_golden(golden, case, tag):
    g = golden(case)
    nS, nA = g["pi"].shape
    log, order, init = _clog(g, int(g["sampler_seed"]))
    eps, alpha = (0.2, 0.1) if tag == "const" else (g["ql_sched_eps"], g["ql_sched_alpha"])
    out = CO.run(log, 2, CO.QLEARN, nS, 25, gamma=0.9, alpha=alpha, epsilon=eps, u_feed=g["U"], tie_feed=g["T"],
                 trace_cap=log.n_rows, order=order, init_row=init, threads=2)
    n = g[f"ql_{tag}_trace"].size
    for r in range(2):
        assert out["accepted"][r] == n
        assert np.array_equal(out["trace_row"][r, :n], g[f"ql_{tag}_trace"])
        assert out["draws"][r] == int(g[f"ql_{tag}_draws"]) and out["ties"][r] == int(g[f"ql_{tag}_ties"])
        assert np.array_equal(out["Q"][r], g[f"ql_{tag}_Q"])
        assert np.array_equal(out["td_error"][r, :n], g[f"ql_{tag}_TD"])
        assert out["trace_hash"][r] == CO.trace_hash(g[f"ql_{tag}_trace"])
        ne = out["n_returns"][r]
        assert ne == g[f"ql_{tag}_Gs"].size
        np.testing.assert_allclose(out["Gs"][r, :ne], g[f"ql_{tag}_Gs"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("case", DRIVER_CASES)
@pytest.mark.parametrize("tag", ["const", "sched"])
def test_c_float32_learner_matches_reference_on_a_float32_table(golden, case, tag):
    """The float32 learner (the GPU production mode): the reference's own update lines psrs.py:165-168 run on a float32 Q
    table (tests/golden/make_golden.py `F32Table`) — Q, TD errors, accepted rows, draw and tie counts bit for bit."""
    g, g32 = golden(case), golden(case + "_f32q")
    nS, nA = g["pi"].shape
    log, order, init = _clog(g, int(g["sampler_seed"]))
    eps, alpha = (0.2, 0.1) if tag == "const" else (g["ql_sched_eps"], g["ql_sched_alpha"])
    out = CO.run(log, 2, CO.QLEARN, nS, 25, gamma=0.9, alpha=alpha, epsilon=eps, u_feed=g["U"], tie_feed=g["T"],
                 trace_cap=log.n_rows, order=order, init_row=init, threads=2, q_f32=True)
    n = g32[f"ql_{tag}_trace"].size
    assert g32[f"ql_{tag}_Q"].dtype == np.float32 and out["Q"].dtype == np.float32
    for r in range(2):
        assert out["accepted"][r] == n
        assert np.array_equal(out["trace_row"][r, :n], g32[f"ql_{tag}_trace"])
        assert out["draws"][r] == int(g32[f"ql_{tag}_draws"]) and out["ties"][r] == int(g32[f"ql_{tag}_ties"])
        assert np.array_equal(out["Q"][r], g32[f"ql_{tag}_Q"])
        assert np.array_equal(out["td_error"][r, :n].astype(np.float32), g32[f"ql_{tag}_TD"])
        assert np.array_equal(out["td_error"][r, :n], g32[f"ql_{tag}_TD"].astype(np.float64))
        ne = out["n_returns"][r]
        assert ne == g32[f"ql_{tag}_Gs"].size
        np.testing.assert_allclose(out["Gs"][r, :ne], g32[f"ql_{tag}_Gs"], rtol=1e-5, atol=1e-5)
    # the float32 and float64 learners really differ on these cases (the fixture is not vacuous)
    assert not np.array_equal(g32[f"ql_{tag}_Q"].astype(np.float64), g[f"ql_{tag}_Q"])


@pytest.mark.parametrize("case", DRIVER_CASES + ["drivers_float_a"])
def test_c_evalmc_and_expsarsa_match_reference_golden(golden, case):
    g = golden(case)
    nS, nA = g["pi"].shape
    log, order, init = _clog(g, int(g["sampler_seed"]))
    out = CO.run(log, 1, CO.FIXED_PI, nS, 40, gamma=0.97, pi=g["pi"], u_feed=g["U"], trace_cap=log.n_rows, order=order,
                 init_row=init)
    n = g["mc_trace"].size
    assert np.array_equal(out["trace_row"][0, :n], g["mc_trace"]) and out["accepted"][0] == n
    assert np.array_equal(out["lengths"][0, :out["n_lengths"][0]], g["mc_lengths"])
    assert out["n_returns"][0] == g["mc_Gs"].size
    if "es_trace" in g.files:
        out = CO.run(log, 1, CO.EXPSARSA, nS, 25, gamma=0.95, alpha=0.2, pi=g["pi"], u_feed=g["U"], trace_cap=log.n_rows,
                     order=order, init_row=init)
        n = g["es_trace"].size
        assert np.array_equal(out["trace_row"][0, :n], g["es_trace"])
        np.testing.assert_allclose(out["Q"][0], g["es_Q"], rtol=1e-12, atol=1e-14)


def test_c_bucketing_equals_python_oracle():
    rng = np.random.default_rng(0)
    z = rng.integers(-1, 40, 5000); zn = rng.integers(-1, 40, 5000)
    steps = rng.integers(0, 3, 5000)
    log = CO.Log(z=z, z_next=zn, action=np.zeros(5000), reward=np.zeros(5000), terminal=np.zeros(5000),
                 p_log=np.full((5000, 2), 0.5, dtype=np.float32), steps=steps, n_buckets=41, z_offset=1)
    off, order = O.bucket(z, 41, 1)
    assert np.array_equal(log.offsets, off) and np.array_equal(log.order, order)
    init = O.init_rows(steps, 5000)
    assert log.n_init == init.size and np.array_equal(log.init_row[:init.size], init)


def test_c_philox_ring_replicas_equal_python_oracle_with_same_streams():
    """Philox-driven C replicas == python-oracle replicas fed the same numbers (ring offsets mirrored)."""
    import ctypes as C
    rng = np.random.default_rng(3)
    N, nS, nA, R = 4000, 12, 4, 5
    z = rng.integers(0, nS, N); zn = rng.integers(0, nS, N)
    pb = rng.dirichlet(np.ones(nA), size=nS).clip(0.05); pb /= pb.sum(1, keepdims=True)
    p_log = pb[z].astype(np.float32)
    a = (rng.random(N)[:, None] > np.cumsum(pb[z], 1)).sum(1).clip(0, nA - 1)
    r = rng.normal(size=N).astype(np.float32); done = rng.random(N) < 0.1
    t = np.where(np.r_[True, done[:-1]], 0, 1)
    log = CO.Log(z=z, z_next=zn, action=a, reward=r, terminal=done, p_log=p_log, steps=t, n_buckets=nS)
    out = CO.run(log, R, CO.QLEARN, nS, 20, gamma=0.9, alpha=0.1, epsilon=0.3, seed=77, replica_id0=1000, ring=True,
                 ring_seed=5, trace_cap=N, threads=3)
    lib = CO.load()
    # restate the streams in python from the documented Philox / ring definitions, via the python oracle
    from tests.helpers import philox_uniform, philox_tie, ring_start, init_ring_start
    for rep in range(R):
        rid = 1000 + rep
        U = [philox_uniform(77, rid, 0, d) for d in range(int(out["draws"][rep]) + 8)]
        T = [philox_tie(77, rid, d) for d in range(int(out["ties"][rep]) + 8)]
        lens = np.diff(log.offsets)
        starts = [ring_start(5, rid, b, int(lens[b])) if lens[b] else 0 for b in range(nS)]
        env = O.OraclePSRS(z=z, z_next=zn, s=z, s_next=zn, action=a, reward=r, terminal=done, p_log=p_log, steps=t,
                           n_buckets=nS, uniforms=O.FeedUniforms(U), ring_start=starts,
                           init_start=init_ring_start(5, rid, log.n_init))
        Q, info = O.qlearn(env, 20, gamma=0.9, alpha=0.1, epsilon=0.3, ties=O.FeedTies(T), nS=nS, nA=nA)
        n = info["trace"].size
        assert out["accepted"][rep] == n and np.array_equal(out["trace_row"][rep, :n], info["trace"])
        assert np.array_equal(out["Q"][rep], Q)


def test_c_agent_ids_sentinel_offset_and_budget_equal_python_oracle():
    """The checker semantics tests/test_gpu_configs.py leans on: agent ids distinct from the queue keys (s = z + 1), the
    CartPole sentinel as bucket 0 (z_offset = 1, z' = -1 on terminal rows) and the step budget — C replicas against
    python-oracle replicas fed the same Philox numbers."""
    from tests.helpers import philox_uniform, philox_tie
    rng = np.random.default_rng(8)
    N, nB, nA, R = 6000, 9, 2, 4
    z = rng.integers(0, nB - 1, N)
    zn = rng.integers(-1, nB - 1, N)
    done = zn < 0
    pb = rng.dirichlet(np.ones(nA), size=nB).clip(0.1); pb /= pb.sum(1, keepdims=True)
    p_log = pb[z].astype(np.float32)
    a = (rng.random(N) > pb[z, 0]).astype(np.int64)
    r = rng.normal(size=N).astype(np.float32)
    t = np.where(np.r_[True, done[:-1]], 0, 1)
    s, sn, nS = z + 1, zn + 1, nB
    log = CO.Log(z=z, z_next=zn, s=s, s_next=sn, action=a, reward=r, terminal=done, p_log=p_log, steps=t, n_buckets=nB, z_offset=1)
    out = CO.run(log, R, CO.QLEARN, nS, 50, gamma=0.9, alpha=0.2, epsilon=0.2, seed=5, replica_id0=40, max_steps=120,
                 trace_cap=200, threads=2)
    assert (out["accepted"] == 120).all()
    for rep in range(R):
        rid = 40 + rep
        U = [philox_uniform(5, rid, 0, d) for d in range(int(out["draws"][rep]) + 8)]
        T = [philox_tie(5, rid, d) for d in range(int(out["ties"][rep]) + 8)]
        env = O.OraclePSRS(z=z, z_next=zn, s=s, s_next=sn, action=a, reward=r, terminal=done, p_log=p_log, steps=t,
                           n_buckets=nB, z_offset=1, uniforms=O.FeedUniforms(U))
        Q, info = O.qlearn(env, 50, gamma=0.9, alpha=0.2, epsilon=0.2, ties=O.FeedTies(T), nS=nS, nA=nA, max_steps=120)
        n = info["trace"].size
        assert n == 120 and np.array_equal(out["trace_row"][rep, :n], info["trace"])
        assert np.array_equal(out["Q"][rep], Q)


def test_c_mlp_equals_python_fma_emulation(golden):
    g = golden("encoders")
    for tag in ["mlp_grid", "mlp_cart", "mlp_wide"]:
        x, W1, b1, W2, b2 = (g[f"{tag}_{k}"] for k in ["x", "W1", "b1", "W2", "b2"])
        assert np.array_equal(CO.mlp_ids(x, W1, b1, W2, b2), O.mlp_ids(x, W1, b1, W2, b2))


def test_c_transition_into_a_state_without_queue_ends_the_replica_at_once():
    """psrs.py:44: `self.queues[self.z]` raises KeyError on the step AFTER a transition into a state that was never logged as
    a from-state, and the reference's whole run is lost.  The restatement (like the device kernels) ends the replica at the
    transition itself: committed and counted, nothing appended, and the same whether or not a step budget ends there too."""
    # state 0 -> state 1 (logged, two rows) -> state 2 (never a from-state); all rows initial, uniform behaviour and target
    z = np.array([0, 1, 1], dtype=np.int64)
    zn = np.array([1, 2, 2], dtype=np.int64)
    a = np.zeros(3, dtype=np.int64)
    r = np.array([1.0, 2.0, 3.0], dtype=np.float32)
    done = np.zeros(3, dtype=bool)
    p = np.full((3, 2), 0.5, dtype=np.float32)
    log = CO.Log(z=z, z_next=zn, action=a, reward=r, terminal=done, p_log=p, steps=np.array([0, 1, 2]), n_buckets=3)
    pi = np.full((3, 2), 0.5)
    U = np.zeros((1, 16))                                   # u = 0 accepts every pop
    for budget in (0, 2, 5):
        out = CO.run(log, 1, CO.FIXED_PI, 3, 4, gamma=1.0, pi=pi, u_feed=U, trace_cap=8, max_steps=budget, ep_cap=4, threads=1)
        assert out["status"][0] == 4                        # NO_QUEUE
        assert out["accepted"][0] == 2 and list(out["trace_row"][0, :2]) == [0, 1]
        assert out["n_returns"][0] == 0 and out["n_lengths"][0] == 0
    out = CO.run(log, 1, CO.FIXED_PI, 3, 4, gamma=1.0, pi=pi, u_feed=U, trace_cap=8, max_steps=1, ep_cap=4, threads=1)
    assert out["status"][0] == 0 and out["accepted"][0] == 1    # the budget ended the run one step earlier: RUNNING
