"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden vectors."""
import os

import numpy as np
import pytest

from oracle import psrs_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import offsim4rl_b200
    return offsim4rl_b200


def make_log(seed, N, nS, nA, ep_len=8, p_dtype=np.float32, skew=False, neg_key=False):
    rng = np.random.default_rng(seed)
    if skew:
        w = 1.0 / np.arange(1, nS + 1) ** 1.1
        z = rng.choice(nS, size=N, p=w / w.sum())
    else:
        z = rng.integers(0, nS, size=N)
    zn = rng.integers(0, nS, size=N)
    if neg_key:
        z = z - 1
        zn = zn - 1
    pb = rng.dirichlet(np.ones(nA), size=nS + 1).clip(0.05)
    pb /= pb.sum(1, keepdims=True)
    p_log = pb[z].astype(p_dtype)
    u = rng.random(N)
    a = (u[:, None] > np.cumsum(pb[z], axis=1)).sum(1).clip(0, nA - 1)
    r = rng.normal(size=N).astype(np.float32)
    done = rng.random(N) < 1.0 / ep_len
    t = np.zeros(N, dtype=np.int64)
    start = np.r_[True, done[:-1]]
    idx = np.arange(N)
    t = idx - np.maximum.accumulate(np.where(start, idx, 0))
    return dict(z=z.astype(np.int64), z_next=zn.astype(np.int64), a=a.astype(np.int64), r=r, done=done, t=t, p_log=p_log)


def device_log(ob, g, **kw):
    kw.setdefault("steps", g["t"])
    return ob.BucketedLog(z=g["z"], z_next=g["z_next"], action=g["a"], reward=g["r"], terminal=g["done"],
                          p_log=g["p_log"], **kw)


def oracle_env(g, **kw):
    kw.setdefault("steps", g["t"])
    return O.OraclePSRS(z=g["z"], z_next=g["z_next"], s=g["z"], s_next=g["z_next"], action=g["a"], reward=g["r"],
                        terminal=g["done"], p_log=g["p_log"], **kw)


# --------------------------------------------------------------------------------------- bucketing
@pytest.mark.parametrize("N,nS,nA,kw", [
    (0, 3, 2, {}), (1, 1, 2, {}), (33, 4, 4, {}), (5000, 25, 5, {}), (100000, 1000, 4, {}),
    (300000, 10000, 4, {}), (200000, 162, 2, dict(neg_key=True)), (150000, 50, 4, dict(skew=True)),
    (70000, 7, 9, {}), (40000, 3, 16, {}),
])
def test_bucketing_is_stable_counting_sort(ob, N, nS, nA, kw):
    g = make_log(1, N, nS, nA, **kw)
    z_offset = 1 if kw.get("neg_key") else 0
    log = device_log(ob, g, n_buckets=nS + z_offset, z_offset=z_offset)
    offsets, order = O.bucket(g["z"], nS + z_offset, z_offset)
    assert np.array_equal(log.offsets.cpu().numpy(), offsets)
    assert np.array_equal(log.order.cpu().numpy()[:N], order)                       # psrs.py:26
    init = O.init_rows(g["t"], N)
    assert log.n_init == init.size
    assert np.array_equal(log.init_row.cpu().numpy()[:init.size], init)              # psrs.py:22
    assert np.array_equal(log.init_z.cpu().numpy()[:init.size], g["z"][init])
    # packed records carry the row's fields in queue order
    if N:
        nv = (nA + 3) // 4
        rec = log.records.cpu().numpy().reshape(N, 4 * nv + 4)
        assert np.array_equal(rec[:, :nA], g["p_log"][order])
        assert np.all(rec[:, nA:4 * nv] == 1)                        # padding entries hold 1.0
        assert np.array_equal(rec[:, 4 * nv], g["r"][order])
        tail = rec[:, 4 * nv + 1:].copy().view(np.int32)
        assert np.array_equal(tail[:, 0], g["z_next"][order])
        assert np.array_equal(tail[:, 1], order)                     # the row index field IS the queue order
        assert np.array_equal(tail[:, 2] & 0xFF, g["a"][order])
        assert np.array_equal((tail[:, 2] >> 8) & 1, g["done"][order].astype(np.int32))


def test_no_steps_column_makes_every_row_initial(ob):
    g = make_log(2, 9000, 12, 3)
    log = device_log(ob, g, steps=None, n_buckets=12)
    assert log.n_init == 9000
    assert np.array_equal(log.init_row.cpu().numpy(), np.arange(9000))               # data.py:72


@pytest.mark.parametrize("seed", [0, 3, 98052, 2**40 + 17])
def test_seeded_shuffle_replays_numpy(ob, seed):
    g = make_log(3, 6000, 30, 4, skew=True)
    log = device_log(ob, g, n_buckets=30)
    sh = log.seeded(seed)
    order, init = O.apply_seed(*O.bucket(g["z"], 30), O.init_rows(g["t"], 6000), seed)
    assert np.array_equal(sh.order.cpu().numpy(), order)                             # psrs.py:29-30
    assert np.array_equal(sh.init_row.cpu().numpy()[:init.size], init)               # psrs.py:23
    assert np.array_equal(sh.init_z.cpu().numpy()[:init.size], g["z"][init])
    rec = sh.records.cpu().numpy().reshape(6000, 8)
    assert np.array_equal(rec[:, :4], g["p_log"][order])
    # the log-order base is untouched
    assert np.array_equal(log.order.cpu().numpy(), O.bucket(g["z"], 30)[1])


def test_seeded_queues_match_reference_golden(ob, golden):
    g = golden("seeded_walk")
    log = ob.BucketedLog(z=g["z"], z_next=g["z_next"], action=g["a"], reward=g["r"], terminal=g["done"],
                         p_log=g["p_log"], steps=g["t"], n_buckets=9)
    for seed in [0, 3, 98052]:
        sh = log.seeded(seed)
        assert np.array_equal(sh.offsets.cpu().numpy(), g[f"offsets_{seed}"])
        assert np.array_equal(sh.order.cpu().numpy(), g[f"order_{seed}"])
        n0 = sh.n_init
        assert np.array_equal(sh.init_z.cpu().numpy()[:n0], g[f"init_zs_{seed}"][:, 0])


# --------------------------------------------------------------------------------------- step API
@pytest.mark.parametrize("lanes", [1, 2, 4, 8, 32])
def test_batched_step_matches_oracle_under_fed_uniforms(ob, lanes):
    import torch
    g = make_log(4, 8000, 15, 4)
    R = 37
    rng = np.random.default_rng(9)
    U = rng.random((R, 40000))
    log = device_log(ob, g, n_buckets=15)
    sim = ob.BatchedPSRS(log, R, rng="feed", u_feed=U, lanes=lanes)
    envs = [oracle_env(g, n_buckets=15, uniforms=O.FeedUniforms(U[r])) for r in range(R)]
    pis = rng.dirichlet(np.ones(4), size=R)
    s_dev, st = sim.reset()
    s_or = [e.reset() for e in envs]
    assert np.array_equal(s_dev.cpu().numpy(), np.array(s_or))
    for it in range(120):
        out = sim.step(pis)
        rows = out["row"].cpu().numpy()
        status = out["status"].cpu().numpy()
        need_reset = np.zeros(R, dtype=np.uint8)
        for r, e in enumerate(envs):
            if e.status not in (O.RUNNING, O.QUEUE_EMPTY):
                continue
            row = e.step(pis[r])
            if row is None:
                assert rows[r] == -1 and status[r] == O.QUEUE_EMPTY
            else:
                assert rows[r] == row, (it, r)
                if g["done"][row]:
                    need_reset[r] = 1
                    e.reset()
        assert np.array_equal(sim.draws.cpu().numpy(), [e.uniforms.k for e in envs])
        assert np.array_equal(sim.pops.cpu().numpy(), [e.pops for e in envs])
        if need_reset.any():
            sim.reset(mask=need_reset)
        assert np.array_equal(sim.cur_s.cpu().numpy()[[e.status == O.RUNNING for e in envs]],
                              [int(e.cur_s) for e in envs if e.status == O.RUNNING])
    torch.cuda.synchronize()


@pytest.mark.parametrize("seed", [0, 3, 98052])
def test_pcg64_walk_reproduces_reference_run_bit_for_bit(ob, golden, seed):
    """reset_sampler(seed) + the reference's OWN default_rng(seed) uniforms, no injection."""
    g = golden("seeded_walk")
    log = ob.BucketedLog(z=g["z"], z_next=g["z_next"], action=g["a"], reward=g["r"], terminal=g["done"],
                         p_log=g["p_log"], steps=g["t"], n_buckets=9).seeded(seed)
    sim = ob.BatchedPSRS(log, 3, rng="pcg64", seed=seed, lanes=4)      # three identical replicas
    trace, resets = [], []
    s, _ = sim.reset()
    resets.append(int(s[0]))
    while True:
        out = sim.step(g["p_new"])
        row = int(out["row"][1])
        assert int(out["row"][0]) == row == int(out["row"][2])
        if row < 0:
            break
        trace.append(row)
        if g["done"][row]:
            s, st = sim.reset()
            if int(st[0]) == O.INIT_EMPTY:
                resets.append(-1)
                break
            resets.append(int(s[0]))
    assert np.array_equal(trace, g[f"trace_{seed}"])
    assert np.array_equal(resets, g[f"resets_{seed}"])
    assert int(sim.pops[0]) == int(g[f"pops_{seed}"])


# --------------------------------------------------------------------------------------- fused drivers vs golden
DRIVER_CASES = ["drivers_f64", "drivers_f32", "drivers_p64_a5", "drivers_edge"]


def _golden_log(ob, g):
    steps = g["t"] if bool(g["with_t"]) else None
    nS = g["pi"].shape[0]
    return ob.BucketedLog(z=g["z"], z_next=g["z_next"], action=g["a"], reward=g["r"], terminal=g["done"],
                          p_log=g["p_log"], steps=steps, n_buckets=nS).seeded(int(g["sampler_seed"]))


@pytest.mark.parametrize("case", DRIVER_CASES)
@pytest.mark.parametrize("tag", ["const", "sched"])
@pytest.mark.parametrize("lanes,ppl,stage", [(1, 1, False), (4, 2, True), (32, 2, False), (2, 2, True), (8, 1, True), (1, 4, True),
                                             (16, 4, False)])
def test_fused_qlearn_matches_reference(ob, golden, case, tag, lanes, ppl, stage):
    import torch
    g = golden(case)
    nS, nA = g["pi"].shape
    N = g["z"].shape[0]
    sim = ob.BatchedPSRS(_golden_log(ob, g), 2, rng="feed", u_feed=g["U"], tie_feed=g["T"], lanes=lanes,
                         pops_per_lane=ppl, stage_smem=stage)
    eps, alpha = (0.2, 0.1) if tag == "const" else (g["ql_sched_eps"], g["ql_sched_alpha"])
    res = sim.run_qlearn(nS, 25, gamma=0.9, alpha=alpha, epsilon=eps, q_dtype=torch.float64, trace_cap=N)
    n = g[f"ql_{tag}_trace"].size
    for r in range(2):
        assert int(res.accepted[r]) == n
        assert np.array_equal(res.trace_row[r, :n].cpu().numpy(), g[f"ql_{tag}_trace"])      # bit-exact accepted rows
        assert int(sim.draws[r]) == int(g[f"ql_{tag}_draws"])
        assert int(sim.ties[r]) == int(g[f"ql_{tag}_ties"])
        assert np.array_equal(res.Q[r].cpu().numpy(), g[f"ql_{tag}_Q"])                      # same IEEE op order
        assert np.array_equal(res.td_error[r, :n].cpu().numpy(), g[f"ql_{tag}_TD"])
        ne = int(res.n_returns[r])
        assert ne == g[f"ql_{tag}_Gs"].size
        np.testing.assert_allclose(res.Gs[r, :ne].cpu().numpy(), g[f"ql_{tag}_Gs"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("case", DRIVER_CASES)
@pytest.mark.parametrize("tag", ["const", "sched"])
@pytest.mark.parametrize("lanes,ppl,aos", [(4, 0, False), (1, 1, False), (32, 2, False), (1, 0, True), (2, 0, True), (8, 0, True)])
def test_fused_qlearn_float32_q_matches_reference_bit_for_bit(ob, golden, case, tag, lanes, ppl, aos):
    """The float32 learner (the production Q dtype) against the reference's own update lines run on a float32 table
    (tests/golden/make_golden.py `F32Table`): accepted rows, draw/tie counts, Q and TD errors bit for bit — in the plain
    layout and in the AoS replica-state layout the benchmark uses."""
    import torch
    g, g32 = golden(case), golden(case + "_f32q")
    nS, nA = g["pi"].shape
    N = g["z"].shape[0]
    if aos and nA > 4:
        pytest.skip("AoS state holds <= 4 actions")
    sim = ob.BatchedPSRS(_golden_log(ob, g), 2, rng="feed", u_feed=g["U"], tie_feed=g["T"], lanes=lanes, pops_per_lane=ppl,
                         aos_state=aos)
    eps, alpha = (0.2, 0.1) if tag == "const" else (g["ql_sched_eps"], g["ql_sched_alpha"])
    res = sim.run_qlearn(nS, 25, gamma=0.9, alpha=alpha, epsilon=eps, q_dtype=torch.float32, trace_cap=N)
    n = g32[f"ql_{tag}_trace"].size
    assert res.Q.dtype == torch.float32
    for r in range(2):
        assert int(res.accepted[r]) == n
        assert np.array_equal(res.trace_row[r, :n].cpu().numpy(), g32[f"ql_{tag}_trace"])
        assert int(sim.draws[r]) == int(g32[f"ql_{tag}_draws"]) and int(sim.ties[r]) == int(g32[f"ql_{tag}_ties"])
        assert np.array_equal(res.Q[r].cpu().numpy(), g32[f"ql_{tag}_Q"])                    # same float32 op order
        assert np.array_equal(res.td_error[r, :n].cpu().numpy(), g32[f"ql_{tag}_TD"].astype(np.float64))
        ne = int(res.n_returns[r])
        assert ne == g32[f"ql_{tag}_Gs"].size
        np.testing.assert_allclose(res.Gs[r, :ne].cpu().numpy(), g32[f"ql_{tag}_Gs"], rtol=1e-5, atol=1e-5)
    # north_star's tolerance bar against the float64 reference run, where the float32 trajectory did not fork from it
    if np.array_equal(g32[f"ql_{tag}_trace"], g[f"ql_{tag}_trace"]):
        np.testing.assert_allclose(res.Q[0].cpu().numpy(), g[f"ql_{tag}_Q"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", DRIVER_CASES)
def test_fused_expsarsa_matches_reference(ob, golden, case):
    import torch
    g = golden(case)
    nS, nA = g["pi"].shape
    N = g["z"].shape[0]
    sim = ob.BatchedPSRS(_golden_log(ob, g), 2, rng="feed", u_feed=g["U"], lanes=4, pops_per_lane=2)
    res = sim.run_expsarsa(nS, 25, pi=g["pi"], gamma=0.95, alpha=0.2, q_dtype=torch.float64, trace_cap=N)
    n = g["es_trace"].size
    assert int(res.accepted[1]) == n
    assert np.array_equal(res.trace_row[1, :n].cpu().numpy(), g["es_trace"])
    assert int(sim.draws[1]) == int(g["es_draws"])
    np.testing.assert_allclose(res.Q[1].cpu().numpy(), g["es_Q"], rtol=1e-12, atol=1e-14)
    ne = int(res.n_returns[1])
    assert ne == g["es_Gs"].size
    np.testing.assert_allclose(res.Gs[1, :ne].cpu().numpy(), g["es_Gs"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("case", DRIVER_CASES + ["drivers_float_a"])
@pytest.mark.parametrize("lanes,ppl", [(1, 1), (8, 2), (16, 1), (2, 2)])
def test_fused_evalmc_matches_reference(ob, golden, case, lanes, ppl):
    g = golden(case)
    nS, nA = g["pi"].shape
    N = g["z"].shape[0]
    sim = ob.BatchedPSRS(_golden_log(ob, g), 2, rng="feed", u_feed=g["U"], lanes=lanes, pops_per_lane=ppl)
    res = sim.run_evalmc(nS, 40, pi=g["pi"], gamma=0.97, trace_cap=N)
    n = g["mc_trace"].size
    assert int(res.accepted[0]) == n
    assert np.array_equal(res.trace_row[0, :n].cpu().numpy(), g["mc_trace"])
    assert int(sim.draws[0]) == int(g["mc_draws"])
    ne, nl = int(res.n_returns[0]), int(res.n_lengths[0])
    assert ne == g["mc_Gs"].size and nl == g["mc_lengths"].size                   # psrs.py:265-269
    assert np.array_equal(res.lengths[0, :nl].cpu().numpy(), g["mc_lengths"])
    np.testing.assert_allclose(res.Gs[0, :ne].cpu().numpy(), g["mc_Gs"], rtol=1e-5, atol=1e-5)


def test_chunked_runs_equal_one_run(ob):
    """max_steps budgets make ols_run resumable; chunking must be invisible."""
    import torch
    g = make_log(7, 20000, 20, 4)
    R = 50
    rng = np.random.default_rng(1)
    U, T = rng.random((R, 60000)), rng.random((R, 60000))
    log = device_log(ob, g, n_buckets=20)
    one = ob.BatchedPSRS(log, R, rng="feed", u_feed=U, tie_feed=T, lanes=4)
    ra = one.run_qlearn(20, 40, gamma=0.95, alpha=0.1, epsilon=0.1, q_dtype=torch.float64, trace_cap=20000)
    two = ob.BatchedPSRS(log, R, rng="feed", u_feed=U, tie_feed=T, lanes=4)
    rb = None
    for _ in range(200):
        rb = two.run_qlearn(20, 40, gamma=0.95, alpha=0.1, epsilon=0.1, q_dtype=torch.float64, trace_cap=20000,
                            max_steps=17, outputs=rb)
        if bool((rb.status != O.RUNNING).all()):
            break
    assert torch.equal(ra.status, rb.status) and torch.equal(ra.accepted, rb.accepted)
    assert torch.equal(ra.trace_row, rb.trace_row) and torch.equal(ra.Q, rb.Q)
    assert torch.equal(ra.Gs, rb.Gs) and torch.equal(ra.lengths, rb.lengths)
    # and it matches the oracle replica by replica
    for r in (0, R - 1):
        env = oracle_env(g, n_buckets=20, uniforms=O.FeedUniforms(U[r]))
        Q, info = O.qlearn(env, 40, gamma=0.95, alpha=0.1, epsilon=0.1, ties=O.FeedTies(T[r]), nS=20, nA=4)
        n = info["trace"].size
        assert np.array_equal(ra.trace_row[r, :n].cpu().numpy(), info["trace"])
        assert np.array_equal(ra.Q[r].cpu().numpy(), Q)


def test_statuses_and_edge_cases(ob):
    import torch
    # state 2 is reachable but never a from-state -> KeyError in the reference -> NO_QUEUE here
    z = np.array([0, 0, 1, 1]); zn = np.array([1, 2, 0, 0])
    log = ob.BucketedLog(z=z, z_next=zn, action=np.zeros(4), reward=np.ones(4), terminal=np.zeros(4),
                         p_log=np.full((4, 2), 0.5, dtype=np.float32), steps=np.array([0, 1, 1, 1]), n_buckets=3)
    sim = ob.BatchedPSRS(log, 1, rng="philox", seed=1, reject="accept_all")
    s, st = sim.reset()
    assert int(st[0]) == O.RUNNING and int(s[0]) == 0
    out = sim.step(np.array([0.5, 0.5]))
    assert int(out["row"][0]) == 0 and int(out["s_next"][0]) == 1
    out = sim.step(np.array([0.5, 0.5]))
    assert int(out["row"][0]) == 2
    out = sim.step(np.array([0.5, 0.5]))
    assert int(out["row"][0]) == 1 and int(out["s_next"][0]) == 2
    out = sim.step(np.array([0.5, 0.5]))
    assert int(out["row"][0]) == -1 and int(out["status"][0]) == O.NO_QUEUE
    s, st = sim.reset()
    assert int(st[0]) == O.INIT_EMPTY                                               # psrs.py:33-35
    assert int(sim.draws[0]) == 0                                                   # accept-all draws nothing
    # empty log
    empty = ob.BucketedLog(z=np.zeros(0), z_next=np.zeros(0), action=np.zeros(0), reward=np.zeros(0),
                           terminal=np.zeros(0), p_log=np.zeros((0, 2), dtype=np.float32), n_buckets=2)
    sim = ob.BatchedPSRS(empty, 4)
    s, st = sim.reset()
    assert (st == O.INIT_EMPTY).all()
    torch.cuda.synchronize()


# --------------------------------------------------------------------------------------- encoders
def test_cartpole_box_encoder_bit_exact(ob, golden):
    from offsim4rl_b200.encoders.heuristic import CartpoleBoxEncoder
    g = golden("encoders")
    enc = CartpoleBoxEncoder()
    assert np.array_equal(enc.encode(g["box_obs32"]), g["box_ids32"])
    assert np.array_equal(enc.encode(g["box_obs64"]), g["box_ids64"])


@pytest.mark.parametrize("tag", ["mlp_grid", "mlp_cart", "mlp_wide"])
def test_mlp_encoder_matches_fma_oracle_bit_exact(ob, golden, tag):
    from offsim4rl_b200.encoders.homer import HOMEREncoder
    g = golden("encoders")
    x, W1, b1, W2, b2 = (g[f"{tag}_{k}"] for k in ["x", "W1", "b1", "W2", "b2"])
    enc = HOMEREncoder.from_weights(W1, b1, W2, b2)
    ids = enc.encode(x)
    assert np.array_equal(ids, O.mlp_ids(x, W1, b1, W2, b2))
    # against torch (the reference's own forward): equal except float32 near-ties (see the oracle test)
    assert (ids != g[f"{tag}_ids"]).sum() <= 2
    # ... and every such row is flagged by the audit (margin of the two largest logits within the rounding-error bound),
    # which flags few rows: a caller can list exactly the ids that depend on the accumulation order
    ids2, near = enc.audit_device(x)
    near = near.cpu().numpy()
    assert np.array_equal(ids2.cpu().numpy(), ids)
    assert near[ids != g[f"{tag}_ids"]].all() and near.mean() < 0.02


def test_grid_encoder(ob):
    from offsim4rl_b200.encoders.heuristic import GridEncoder
    rng = np.random.default_rng(0)
    cell = rng.integers(0, 5, size=(50000, 2))
    obs = (cell / 5 + 0.1 + rng.uniform(-0.1, 0.1, size=cell.shape)).astype(np.float32)   # gridworld.py:205-211
    ids = GridEncoder(cells=5).encode(obs)
    assert np.array_equal(ids, O.grid_ids(obs, 5))
    assert (ids == cell[:, 0] + 5 * cell[:, 1]).mean() > 0.999                             # gridworld.py:123-124


# --------------------------------------------------------------------------------------- evaluators
def test_learning_curve_and_totals(ob):
    import torch
    g = make_log(11, 50000, 30, 4)
    R = 1000
    log = device_log(ob, g, n_buckets=30)
    sim = ob.BatchedPSRS(log, R, rng="philox", seed=5, lanes=4)
    pi = np.full((30, 4), 0.25)
    res = sim.run_evalmc(30, 60, pi=pi, gamma=0.99)
    s, s2, c = sim.learning_curve(res, 60, all_reduce=False)
    Gs = res.Gs.cpu().numpy(); ne = res.n_returns.cpu().numpy()
    es, es2, ec = O.learning_curve([Gs[r, :ne[r]] for r in range(R)])
    k = es.size
    np.testing.assert_allclose(s.cpu().numpy()[:k], es, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(s2.cpu().numpy()[:k], es2, rtol=1e-12, atol=1e-12)
    assert np.array_equal(c.cpu().numpy()[:k], ec) and np.all(c.cpu().numpy()[k:] == 0)
    ret, ln = sim.replica_totals(res)
    nl = res.n_lengths.cpu().numpy(); L = res.lengths.cpu().numpy()
    np.testing.assert_allclose(ret.cpu().numpy(), [Gs[r, :ne[r]].sum() for r in range(R)], rtol=1e-12, atol=1e-12)
    assert np.array_equal(ln.cpu().numpy(), [L[r, :nl[r]].sum() for r in range(R)])
    assert int(ln.sum()) == int(res.accepted.sum())


# --------------------------------------------------------------------------------------- production streams vs the C oracle
@pytest.mark.parametrize("agent", ["qlearn", "evalmc"])
@pytest.mark.parametrize("lanes,ppl", [(4, 2), (32, 2), (8, 1), (1, 2), (1, 4), (2, 4)])
@pytest.mark.parametrize("ring", [False, True])
def test_philox_replicas_match_c_oracle(ob, agent, lanes, ppl, ring):
    """Counter-based uniforms + ring offsets: thousands of replicas compared with oracle/ols_oracle.c, which
    consumes the same (seed, replica, draw) -> uniform mapping strictly sequentially."""
    import torch
    from oracle import c_oracle as CO
    g = make_log(21, 200000, 300, 4, ep_len=25)
    R, nS = 3000, 300
    log = device_log(ob, g, n_buckets=nS)
    sim = ob.BatchedPSRS(log, R, rng="philox", seed=4242, ring=ring, ring_seed=7, replica_id0=10**6, lanes=lanes,
                         pops_per_lane=ppl)
    clog = CO.Log(z=g["z"], z_next=g["z_next"], action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"],
                  steps=g["t"], n_buckets=nS)
    pi = np.random.default_rng(5).dirichlet(np.ones(4), size=nS)
    if agent == "qlearn":
        res = sim.run_qlearn(nS, 30, gamma=0.95, alpha=0.2, epsilon=0.15, q_dtype=torch.float64, trace_cap=4000)
        ref = CO.run(clog, R, CO.QLEARN, nS, 30, gamma=0.95, alpha=0.2, epsilon=0.15, seed=4242, replica_id0=10**6,
                     ring=ring, ring_seed=7, trace_cap=4000, threads=8)
    else:
        res = sim.run_evalmc(nS, 30, pi=pi, gamma=0.95, trace_cap=4000)
        ref = CO.run(clog, R, CO.FIXED_PI, nS, 30, gamma=0.95, pi=pi, seed=4242, replica_id0=10**6, ring=ring,
                     ring_seed=7, trace_cap=4000, threads=8)
    assert np.array_equal(res.accepted.cpu().numpy(), ref["accepted"])
    assert np.array_equal(res.pops.cpu().numpy(), ref["pops"])
    assert np.array_equal(sim.draws.cpu().numpy(), ref["draws"])
    assert np.array_equal(res.status.cpu().numpy(), ref["status"])
    assert np.array_equal(res.trace_row.cpu().numpy(), ref["trace_row"])                 # bit-exact accepted rows
    assert np.array_equal(res.n_returns.cpu().numpy(), ref["n_returns"])
    assert np.array_equal(res.lengths.cpu().numpy(), ref["lengths"])
    np.testing.assert_allclose(res.Gs.cpu().numpy(), ref["Gs"], rtol=1e-12, atol=1e-12)
    if agent == "qlearn":
        assert np.array_equal(sim.ties.cpu().numpy(), ref["ties"])
        assert np.array_equal(res.Q.cpu().numpy(), ref["Q"])


PROD_LAYOUTS = [
    # lanes, aos_state, hot_records, slices, R, budget
    (1, True, "tail", 8, 3000, 900),        # the benchmarked kernel: one thread per replica, TAIL hot records, AoS state, time slices
    (1, True, "tail", -1, 20000, 1100),     # ... with the library's own slicing decision (more CTAs than SMs -> 2 slices)
    (1, True, "tail", 0, 777, 300),         # ... unsliced, a ragged last warp
    (1, True, True, 8, 3000, 900),          # the one-lane learner over 8-byte hot records
    (2, True, True, 5, 3000, 700),
    (2, True, "tail", 5, 3000, 700),        # TAIL hot records are ignored by every other kernel
    (32, True, True, 0, 1500, 400),
    (4, True, False, 3, 3000, 500),         # AoS state over the 32-byte records
    (4, False, False, 0, 3000, 500),        # plain layout
]


@pytest.mark.parametrize("lanes,aos,hot,slices,R,budget", PROD_LAYOUTS)
@pytest.mark.parametrize("ring", [True, False])
def test_float32_production_learner_matches_c_oracle(ob, lanes, aos, hot, slices, R, budget, ring):
    """The configuration bench.py times (float32 Q, Philox uniforms, 8-byte hot records, AoS replica state, one lane per
    replica, time slices, a step budget) against oracle/ols_oracle.c in its float32-learner mode (pinned to the reference's
    update lines by tests/test_oracle_c.py): every accepted row, every counter, the returns and the Q tables."""
    import torch
    from oracle import c_oracle as CO
    g = make_log(77, 200000, 300, 4, ep_len=25)
    nS = 300
    log = device_log(ob, g, n_buckets=nS, hot_records=hot)
    sim = ob.BatchedPSRS(log, R, rng="philox", seed=991, ring=ring, ring_seed=13, replica_id0=5 * 10**5, lanes=lanes,
                         aos_state=aos, slices=slices)
    res = sim.run_qlearn(nS, 511, gamma=0.99, alpha=0.1, epsilon=0.1, q_dtype=torch.float32, max_steps=budget, ep_cap=512,
                         trace_cap=budget)
    clog = CO.Log(z=g["z"], z_next=g["z_next"], action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"],
                  steps=g["t"], n_buckets=nS)
    ref = CO.run(clog, R, CO.QLEARN, nS, 511, gamma=0.99, alpha=0.1, epsilon=0.1, seed=991, replica_id0=5 * 10**5, ring=ring,
                 ring_seed=13, max_steps=budget, ep_cap=512, trace_cap=budget, threads=16, q_f32=True)
    assert np.array_equal(res.accepted.cpu().numpy(), ref["accepted"])
    assert np.array_equal(res.pops.cpu().numpy(), ref["pops"]) and np.array_equal(sim.draws.cpu().numpy(), ref["draws"])
    assert np.array_equal(sim.ties.cpu().numpy(), ref["ties"])
    assert np.array_equal(res.trace_row.cpu().numpy(), ref["trace_row"])                 # bit-exact accepted rows
    # a replica stopped by the budget stays RUNNING on the GPU (resumable); the oracle reports the same replicas as RUNNING
    assert np.array_equal(res.status.cpu().numpy(), ref["status"])
    assert np.array_equal(res.n_returns.cpu().numpy(), ref["n_returns"])
    nl = ref["n_lengths"]
    assert np.array_equal(res.n_lengths.cpu().numpy(), nl)
    mask = np.arange(512)[None, :] < nl[:, None]
    assert np.array_equal(res.lengths.cpu().numpy()[mask], ref["lengths"][mask])
    mask = np.arange(512)[None, :] < ref["n_returns"][:, None]
    np.testing.assert_allclose(res.Gs.cpu().numpy()[mask], ref["Gs"][mask], rtol=1e-12, atol=1e-12)
    assert np.array_equal(res.Q.cpu().numpy(), ref["Q"])                                 # float32 Q tables, bit for bit
    sample = [0, 17, R - 1]
    assert np.array_equal(res.q_rows(sample).cpu().numpy(), ref["Q"][sample])
    assert int((ref["accepted"] == budget).sum()) > 0.9 * R                             # the budget, not exhaustion, ended the runs


def test_tail_hot_records_need_one_distribution_per_queue(ob):
    """hot_records='tail' is for logs whose behaviour distribution is a function of the state: the device checks every queue
    bitwise (queue_dist), and a log that does not qualify falls back to the probability hot records — results unchanged."""
    import torch
    from offsim4rl_b200 import _abi as abi
    g = make_log(91, 60000, 40, 4, ep_len=20)
    nS = 40
    log = device_log(ob, g, n_buckets=nS, hot_records="tail")
    assert log.hot_kind == abi.HOT_TAIL
    off, order = O.bucket(g["z"], nS)
    qd = log.queue_dist.cpu().numpy()
    first = order[off[:-1]]
    assert np.array_equal(qd[:, :4], g["p_log"][first])                               # the queue's distribution, bit for bit
    assert np.array_equal(qd[:, 4].view(np.int32), np.full(nS, 3))                    # constant + tame
    w = log.hot.view(torch.int32).view(-1, 2).cpu().numpy()
    assert np.array_equal(w[:, 0].view(np.float32), g["r"][order])
    assert np.array_equal(w[:, 1] & 0x1FFFFFFF, g["z_next"][order]) and np.array_equal((w[:, 1] >> 29) & 3, g["a"][order])
    assert np.array_equal((w[:, 1] >> 31) & 1, g["done"][order].astype(np.int32))
    assert log.seeded(3).hot_kind == abi.HOT_TAIL
    # one record with another distribution: the log no longer qualifies
    g2 = dict(g)
    g2["p_log"] = g["p_log"].copy()
    g2["p_log"][12345] = np.array([0.4, 0.3, 0.2, 0.1], dtype=np.float32)
    log2 = device_log(ob, g2, n_buckets=nS, hot_records="tail")
    assert log2.hot_kind == abi.HOT_PROBS and log2.queue_dist is None
    plain = device_log(ob, g2, n_buckets=nS)
    runs = []
    for lg, aos in ((log2, True), (plain, False)):
        sim = ob.BatchedPSRS(lg, 500, rng="philox", seed=8, ring=True, ring_seed=2, lanes=1, aos_state=aos)
        runs.append(sim.run_qlearn(nS, 30, gamma=0.9, alpha=0.1, epsilon=0.2, max_steps=200, trace_cap=200))
    assert torch.equal(runs[0].trace_row, runs[1].trace_row) and torch.equal(runs[0].Q, runs[1].Q)


def test_accept_filter_never_changes_a_decision(ob):
    """Adversarial uniforms: u placed exactly on, and one ulp around, the exact threshold of the record it
    will meet — the float32 pre-filter must defer to the exact rule there."""
    rng = np.random.default_rng(3)
    nS, nA, N = 6, 4, 3000
    g = make_log(31, N, nS, nA)
    # a single replica walks the log in order; build its uniform stream pop by pop from the oracle's thresholds
    pi = rng.dirichlet(np.ones(nA), size=nS)
    env = oracle_env(g, n_buckets=nS, uniforms=O.FeedUniforms(np.zeros(1)))
    U = []
    # simulate with the oracle, choosing each u adversarially
    env.reset()
    order = env.order; offs = env.offsets
    cursors = np.zeros(nS, dtype=np.int64)
    z = env.cur_z; trace = []
    init_i = 1
    steps = 0
    while steps < 1500:
        lo, hi = offs[z], offs[z + 1]
        if cursors[z] >= hi - lo:
            break
        row = order[lo + cursors[z]]; cursors[z] += 1
        thr = float(O.threshold(pi[z], g["p_log"][row], g["a"][row]))
        choice = rng.integers(0, 5)
        u = [thr, np.nextafter(thr, 0.0), np.nextafter(thr, 2.0), thr * (1 - 2.0**-17), thr * (1 + 2.0**-17)][choice]
        u = float(min(max(u, 0.0), np.nextafter(1.0, 0.0)))
        U.append(u)
        if not (u > thr):
            trace.append(row); steps += 1
            z = int(g["z_next"][row])
            if g["done"][row]:
                if init_i >= env.init.size:
                    break
                z = int(g["z"][env.init[init_i]]); init_i += 1
    U = np.array(U + [0.5] * 64)
    log = device_log(ob, g, n_buckets=nS)
    for lanes, ppl in [(1, 1), (4, 2), (32, 2), (1, 4), (8, 4)]:
        sim = ob.BatchedPSRS(log, 1, rng="feed", u_feed=U, lanes=lanes, pops_per_lane=ppl)
        res = sim.run_evalmc(nS, 10**6 // 1000, pi=pi, gamma=1.0, trace_cap=N, max_steps=len(trace), ep_cap=2000)
        got = res.trace_row[0, :len(trace)].cpu().numpy()
        assert np.array_equal(got, trace), (lanes, ppl)


def test_agent_observation_ids_distinct_from_state_ids(ob):
    """psrs.py:36,49-50 — the agent indexes Q with the observation s, the queues are keyed by z."""
    import torch
    from oracle import c_oracle as CO
    g = make_log(41, 30000, 12, 4)
    rng = np.random.default_rng(0)
    # observation = (z, a noise bit): two observations per latent state
    s = g["z"] * 2 + rng.integers(0, 2, 30000)
    s_next = g["z_next"] * 2 + rng.integers(0, 2, 30000)
    log = ob.BucketedLog(z=g["z"], z_next=g["z_next"], action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"],
                         s=s, s_next=s_next, steps=g["t"], n_buckets=12)
    sim = ob.BatchedPSRS(log, 200, rng="philox", seed=3, lanes=4)
    res = sim.run_qlearn(24, 20, gamma=0.9, alpha=0.1, epsilon=0.2, q_dtype=torch.float64, trace_cap=3000)
    clog = CO.Log(z=g["z"], z_next=g["z_next"], action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"], s=s,
                  s_next=s_next, steps=g["t"], n_buckets=12)
    ref = CO.run(clog, 200, CO.QLEARN, 24, 20, gamma=0.9, alpha=0.1, epsilon=0.2, seed=3, trace_cap=3000, threads=4)
    assert np.array_equal(res.trace_row.cpu().numpy(), ref["trace_row"])
    assert np.array_equal(res.Q.cpu().numpy(), ref["Q"])
    assert np.array_equal(sim.cur_s.cpu().numpy()[ref["status"] == O.DONE] % 2 >= 0, np.ones((ref["status"] == O.DONE).sum(), bool))


def test_learning_curves_agree_with_oracle_under_independent_rngs_ks(ob):
    """north_star: aggregate learning curves must agree under a two-sample KS test when each side uses its OWN RNG.
    GPU: Philox streams, ring-offset decorrelation.  Oracle: NumPy's PCG64 (the reference's generator family), the
    same protocol mirrored on its side."""
    from scipy.stats import ks_2samp
    from tests.helpers import ring_start, init_ring_start
    g = make_log(51, 40000, 10, 4, ep_len=12)
    nS = 10
    pi = np.random.default_rng(2).dirichlet(np.ones(4) * 2, size=nS)
    log = device_log(ob, g, n_buckets=nS)
    R = 4000
    sim = ob.BatchedPSRS(log, R, rng="philox", seed=99, ring=True, ring_seed=11, lanes=4)
    res = sim.run_evalmc(nS, 6, pi=pi, gamma=0.95)
    ne = res.n_returns.cpu().numpy()
    Gs, L = res.Gs.cpu().numpy(), res.lengths.cpu().numpy()
    lens = np.diff(O.bucket(g["z"], nS)[0])
    n_init = O.init_rows(g["t"], 40000).size
    oG, oL = [], []
    for i in range(500):
        rid = 10**7 + i
        env = oracle_env(g, n_buckets=nS, uniforms=O.Pcg64Uniforms(5000 + i),
                         ring_start=[ring_start(11, rid, b, int(lens[b])) for b in range(nS)],
                         init_start=init_ring_start(11, rid, n_init))
        G, ln, _ = O.evalmc(env, 6, pi=pi, gamma=0.95, record=False)
        oG.append(G); oL.append(ln)
    for e in range(6):
        a = Gs[ne > e, e]
        b = np.array([x[e] for x in oG if len(x) > e])
        assert a.size > 3000 and b.size > 400
        assert ks_2samp(a, b).pvalue > 1e-3, f"episode {e}: return distributions differ"
        la = L[ne > e, e]
        lb = np.array([x[e] for x in oL if len(x) > e])
        assert ks_2samp(la, lb).pvalue > 1e-3, f"episode {e}: length distributions differ"
    # and the reduced learning curve is the mean of those per-replica returns
    s, s2, c = sim.learning_curve(res, 6, all_reduce=False)
    np.testing.assert_allclose((s / c).cpu().numpy(), [Gs[ne > e, e].mean() for e in range(6)], rtol=1e-10)


@pytest.mark.parametrize("layout", ["production", "plain_f64"])
def test_learner_curve_agrees_with_the_reference_itself_under_independent_rngs_ks(ob, golden, layout):
    """north_star's third bar for the LEARNER: tests/golden/learner_curves.npz holds the per-episode returns of 320 runs of
    the UNMODIFIED reference `qlearn_psrs` (psrs.py:119-185), each with the reference's own generators (NumPy PCG64
    rejection uniforms, MT19937 tie-breaks); here 6000 GPU replicas learn on the same log with Philox streams (float32 Q,
    TAIL hot records, AoS state, one thread per replica: the benchmarked kernel — or the float64 plain layout).  Per episode index: two-sample KS on the
    returns, and the means of the learning curve within 4 standard errors."""
    import torch
    from scipy.stats import ks_2samp
    g = golden("learner_curves")
    nS, nA = g["pb_state"].shape
    E, R = int(g["n_episodes"]), 6000
    z = g["z"].astype(np.int64)
    prod = layout == "production"
    log = ob.BucketedLog(z=z, z_next=g["z_next"].astype(np.int64), action=g["a"].astype(np.int64), reward=g["r"], terminal=g["done"],
                         p_log=g["pb_state"][z], steps=g["t"], n_buckets=nS, hot_records="tail" if prod else False)
    # GPU replica ids continue after the ids whose data order the reference runs followed (10**7 + i, i < 320)
    sim = ob.BatchedPSRS(log, R, rng="philox", seed=424242, ring=True, ring_seed=int(g["ring_seed"]),
                         replica_id0=int(g["replica_id0"]) + 1000, lanes=1 if prod else 4, aos_state=prod)
    res = sim.run_qlearn(nS, E, gamma=float(g["gamma"]), alpha=float(g["alpha"]), epsilon=float(g["epsilon"]),
                         q_dtype=torch.float32 if prod else torch.float64)
    ne = res.n_returns.cpu().numpy()
    Gs = res.Gs.cpu().numpy()
    ref, ref_n = g["ref_Gs"], g["ref_n_returns"]
    assert bool((ne == E).all()) and bool((ref_n == E).all())
    means = []
    for e in range(E):
        a, b = Gs[:, e], ref[:, e]
        assert ks_2samp(a, b).pvalue > 1e-3, f"episode {e}: return distributions differ"
        se = np.sqrt(a.var() / a.size + b.var() / b.size)
        assert abs(a.mean() - b.mean()) < 4 * se, f"episode {e}: learning-curve means differ"
        means.append(a.mean())
    # the curve rises: the learner learns (rewards depend on (state, action)), on both sides alike
    assert np.mean(means[-4:]) > np.mean(means[:2]) + 1.0 and ref[:, -4:].mean() > ref[:, :2].mean() + 1.0
    # and the reduced learning curve is the mean of those per-replica returns, its spread the reference's
    s, s2, c = sim.learning_curve(res, E, all_reduce=False)
    np.testing.assert_allclose((s / c).cpu().numpy(), means, rtol=1e-10)
    std_gpu = np.sqrt((s2 / c - (s / c) ** 2).cpu().numpy())
    np.testing.assert_allclose(std_gpu, ref.std(0), rtol=0.25)


@pytest.mark.parametrize("agent", ["qlearn", "expsarsa"])
@pytest.mark.parametrize("lanes,ring", [(2, True), (4, False), (32, True), (1, False)])
def test_aos_state_layout_is_bitwise_equivalent(ob, agent, lanes, ring):
    """Cursor + Q row in one 32-byte entry per (replica, state) is only a layout: accepted rows, Q tables, returns and
    counters must be identical to the run with the separate cursor / Q arrays, also across resumed (budgeted) launches."""
    import torch
    g = make_log(61, 120000, 200, 4, ep_len=20)
    nS, R = 200, 2500
    log = device_log(ob, g, n_buckets=nS)
    pi = np.random.default_rng(9).dirichlet(np.ones(4), size=nS)
    outs = []
    for aos in (False, True):
        sim = ob.BatchedPSRS(log, R, rng="philox", seed=77, ring=ring, ring_seed=3, lanes=lanes, aos_state=aos)
        res = None
        for _ in range(3):          # three budgeted launches: the state must carry over in either layout
            if agent == "qlearn":
                res = sim.run_qlearn(nS, 40, gamma=0.9, alpha=0.1, epsilon=0.1, max_steps=150, trace_cap=500, outputs=res)
            else:
                res = sim.run_expsarsa(nS, 40, pi=pi, gamma=0.9, alpha=0.1, max_steps=150, trace_cap=500, outputs=res)
        outs.append((res, sim.draws.clone(), sim.ties.clone()))
    (a, da, ta), (b, db, tb) = outs
    assert torch.equal(a.trace_row, b.trace_row) and torch.equal(a.accepted, b.accepted) and torch.equal(a.pops, b.pops)
    assert torch.equal(da, db) and torch.equal(ta, tb) and torch.equal(a.status, b.status)
    assert torch.equal(a.Q, b.Q.contiguous()) and torch.equal(a.Gs, b.Gs) and torch.equal(a.lengths, b.lengths)
    assert int(a.accepted.sum()) > 0.9 * R * 450


def test_aos_state_rejects_unsupported_shapes(ob):
    import torch
    g = make_log(62, 5000, 10, 4)
    log = device_log(ob, g, n_buckets=10)
    sim = ob.BatchedPSRS(log, 4, aos_state=True)
    with pytest.raises(ValueError):
        sim.run_qlearn(10, 5, gamma=0.9, q_dtype=torch.float64)
    with pytest.raises(RuntimeError):
        sim.step(np.full(4, 0.25))
    with pytest.raises(RuntimeError):
        sim.run_evalmc(10, 5, pi=np.full((10, 4), 0.25), gamma=0.9)


@pytest.mark.parametrize("agent,aos", [("qlearn", True), ("qlearn", False), ("expsarsa", True), ("evalmc", False)])
@pytest.mark.parametrize("lanes,slices", [(2, 16), (4, 3), (32, 7)])
def test_time_slices_are_invisible(ob, agent, aos, lanes, slices):
    """A budgeted run served in time slices (ticket scheduler inside the kernel, state handed from CTA to CTA through
    HBM) must produce exactly the run of one launch per replica block: rows, Q, returns, counters."""
    import torch
    g = make_log(63, 120000, 200, 4, ep_len=20)
    nS, R = 200, 3000
    log = device_log(ob, g, n_buckets=nS)
    pi = np.random.default_rng(10).dirichlet(np.ones(4), size=nS)
    outs = []
    for k in (0, slices):
        sim = ob.BatchedPSRS(log, R, rng="philox", seed=5, ring=True, ring_seed=1, lanes=lanes, aos_state=aos, slices=k)
        res = None
        for budget in (333, 100):      # 333 is not a multiple of the slice count; the second launch resumes
            if agent == "qlearn":
                res = sim.run_qlearn(nS, 60, gamma=0.9, alpha=0.1, epsilon=0.1, max_steps=budget, trace_cap=500, outputs=res)
            elif agent == "expsarsa":
                res = sim.run_expsarsa(nS, 60, pi=pi, gamma=0.9, alpha=0.1, max_steps=budget, trace_cap=500, outputs=res)
            else:
                res = sim.run_evalmc(nS, 60, pi=pi, gamma=0.9, max_steps=budget, trace_cap=500, outputs=res)
        outs.append((res, sim.draws.clone(), sim.ties.clone(), sim.cur_z.clone(), sim.ep_t.clone()))
    (a, da, ta, za, ea), (b, db, tb, zb, eb) = outs
    assert torch.equal(a.trace_row, b.trace_row) and torch.equal(a.accepted, b.accepted) and torch.equal(a.pops, b.pops)
    assert torch.equal(da, db) and torch.equal(ta, tb) and torch.equal(a.status, b.status) and torch.equal(za, zb)
    assert torch.equal(ea, eb) and torch.equal(a.Gs, b.Gs) and torch.equal(a.lengths, b.lengths)
    if agent != "evalmc":
        assert torch.equal(a.Q.contiguous(), b.Q.contiguous())
    assert int(a.accepted.sum()) > 0.9 * R * 433


@pytest.mark.parametrize("agent", ["qlearn", "expsarsa"])
def test_one_lane_production_kernel_sliced_equals_two_lane_unsliced(ob, agent):
    """The C4 production path — one lane per replica over hot records and AoS state, small CTAs, served in automatic time
    slices because the launch holds more CTAs than SMs — against the two-lane kernel in a single unsliced launch."""
    import torch
    g = make_log(91, 600000, 300, 4, ep_len=20)
    nS, R = 300, 40000                              # 313 CTAs of 128 threads > 148 SMs: the automatic rule slices
    log = device_log(ob, g, n_buckets=nS, hot_records=True)
    pi = np.random.default_rng(14).dirichlet(np.ones(4), size=nS)
    outs = []
    for lanes, slices in ((1, -1), (2, 0), (1, 5)):
        sim = ob.BatchedPSRS(log, R, rng="philox", seed=31, ring=True, ring_seed=4, lanes=lanes, aos_state=True, slices=slices)
        res = None
        for budget in (1100, 300):                  # 1100 accepted steps -> two automatic slices; the second launch resumes
            if agent == "qlearn":
                res = sim.run_qlearn(nS, 200, gamma=0.9, alpha=0.1, epsilon=0.1, max_steps=budget, trace_cap=64, outputs=res)
            else:
                res = sim.run_expsarsa(nS, 200, pi=pi, gamma=0.9, alpha=0.1, max_steps=budget, trace_cap=64, outputs=res)
        outs.append((res, sim.draws.clone(), sim.ties.clone(), sim.cur_z.clone(), sim.ep_t.clone()))
    a, da, ta, za, ea = outs[1]
    for b, db, tb, zb, eb in (outs[0], outs[2]):
        assert torch.equal(a.trace_row, b.trace_row) and torch.equal(a.accepted, b.accepted) and torch.equal(a.pops, b.pops)
        assert torch.equal(da, db) and torch.equal(ta, tb) and torch.equal(a.status, b.status) and torch.equal(za, zb)
        assert torch.equal(ea, eb) and torch.equal(a.Gs, b.Gs) and torch.equal(a.lengths, b.lengths)
        assert torch.equal(a.Q.contiguous(), b.Q.contiguous())
    assert int(a.accepted.sum()) > 0.9 * R * 1400


def test_time_slices_need_a_budget(ob):
    g = make_log(64, 5000, 10, 4)
    log = device_log(ob, g, n_buckets=10)
    sim = ob.BatchedPSRS(log, 64, slices=4)
    res = sim.run_qlearn(10, 5, gamma=0.9)            # no max_steps: runs unsliced
    assert int(res.accepted.sum()) > 0
    import ctypes as C
    from offsim4rl_b200 import _abi as abi
    assert abi.load().ols_run_sched_bytes(1000, 2) == 4 * (8 + 1)
    assert abi.load().ols_run_sched_bytes(1000, 3) == 0


def test_aos_restart_and_initial_q(ob):
    """The AoS table is never cleared: a restart makes its entries stale by bumping the epoch.  A restarted block must
    behave like a new one, and a caller-supplied initial Q like the separate-array layout with the same Q."""
    import torch
    g = make_log(65, 60000, 100, 4, ep_len=15)
    nS, R = 100, 700
    log = device_log(ob, g, n_buckets=nS)
    Q0 = torch.from_numpy(np.random.default_rng(3).normal(size=(R, nS, 4)).astype(np.float32)).cuda()

    def run(sim, Q=None):
        return sim.run_qlearn(nS, 30, gamma=0.9, alpha=0.2, epsilon=0.1, max_steps=200, trace_cap=200, Q=Q)

    used = ob.BatchedPSRS(log, R, rng="philox", seed=9, ring=True, lanes=2, aos_state=True)
    run(used, Q0.clone())                  # dirty the table, different epoch and different Q
    for Qinit in (None, Q0):
        used.restart()
        a = run(used, None if Qinit is None else Qinit.clone())
        ref = ob.BatchedPSRS(log, R, rng="philox", seed=9, ring=True, lanes=2, aos_state=False)
        b = run(ref, None if Qinit is None else Qinit.clone())
        assert torch.equal(a.trace_row, b.trace_row) and torch.equal(a.Q, b.Q) and torch.equal(a.Gs, b.Gs)
        assert torch.equal(used.draws, ref.draws) and torch.equal(a.pops, b.pops)


# --------------------------------------------------------------------------------------- hot records
def _expected_hot(p_log, action, nA):
    """numpy restatement of ols_hot_pack (csrc/ols_common.cuh)."""
    p = np.ones((p_log.shape[0], 4), dtype=np.float32)
    p[:, :nA] = p_log
    in_range = (p >= np.float32(2.0 ** -13)) & (p <= 1.0)
    tame = in_range[:, :nA].all(1) & (action < 4)
    codes = np.where(in_range, p, np.float32(1.0)).astype(np.float16).view(np.uint16).astype(np.uint64) & 0x7FFF
    a = action.astype(np.uint64)
    lo = codes[:, 0] | (codes[:, 1] << 16) | ((a & 1) << 15) | (((a >> 1) & 1) << 31)
    hi = codes[:, 2] | (codes[:, 3] << 16) | (tame.astype(np.uint64) << 15)
    return lo | (hi << 32)


@pytest.mark.parametrize("nA,seed", [(4, None), (3, 5), (2, None)])
def test_hot_records_encode_the_queue_records(ob, nA, seed):
    g = make_log(70, 30000, 50, nA)
    g["p_log"][::7, 0] = 1e-5                 # below the half range: not hot-tame
    g["p_log"][3::11, nA - 1] = 0.0
    log = device_log(ob, g, n_buckets=50, hot_records=True).seeded(seed)
    order = log.order.cpu().numpy()
    want = _expected_hot(g["p_log"][order], g["a"][order], nA)
    got = log.hot.cpu().numpy().view(np.uint64)[: order.size]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("agent", ["qlearn", "expsarsa"])
@pytest.mark.parametrize("lanes,ring,seeded", [(2, True, None), (4, False, 3), (32, True, None), (1, True, 9)])
def test_hot_records_never_change_a_run(ob, agent, lanes, ring, seeded):
    """The 8-byte hot records only pre-decide pops whose half-precision threshold estimate is far from the uniform; every
    other pop goes to the full record.  Rows, Q, returns and all counters must equal the run on the 32-byte records,
    including logs with probabilities outside the half range, zero probabilities and float64-only values."""
    import torch
    g = make_log(71, 150000, 200, 4, ep_len=20, p_dtype=np.float64)
    g["p_log"][::13, 1] = 3e-6                 # not hot-tame
    g["p_log"][5::17, 2] = 0.0                  # infinite ratios on the exact path
    nS, R = 200, 3000
    pi = np.random.default_rng(12).dirichlet(np.ones(4), size=nS)
    pi[::9, 0] = 0.0
    pi /= pi.sum(1, keepdims=True)
    outs = []
    for hot in (False, True):
        log = device_log(ob, g, n_buckets=nS, hot_records=hot).seeded(seeded)
        assert (log.hot is not None) == hot
        sim = ob.BatchedPSRS(log, R, rng="philox", seed=21, ring=ring, ring_seed=2, lanes=lanes, aos_state=True)
        res = None
        for budget in (260, 140):
            if agent == "qlearn":
                res = sim.run_qlearn(nS, 50, gamma=0.9, alpha=0.1, epsilon=0.15, max_steps=budget, trace_cap=400, outputs=res)
            else:
                res = sim.run_expsarsa(nS, 50, pi=pi, gamma=0.9, alpha=0.1, max_steps=budget, trace_cap=400, outputs=res)
        outs.append((res, sim.draws.clone(), sim.ties.clone()))
    (a, da, ta), (b, db, tb) = outs
    assert torch.equal(a.trace_row, b.trace_row) and torch.equal(a.accepted, b.accepted) and torch.equal(a.pops, b.pops)
    assert torch.equal(da, db) and torch.equal(ta, tb) and torch.equal(a.status, b.status)
    assert torch.equal(a.Q, b.Q) and torch.equal(a.Gs, b.Gs) and torch.equal(a.lengths, b.lengths)
    assert int(a.accepted.sum()) > 0.5 * R * 400


def _fuzz_config(k):
    rng = np.random.default_rng(1000 + k)
    pick = lambda xs: xs[int(rng.integers(0, len(xs)))]
    c = dict(nA=pick([2, 3, 4, 4, 4, 5]), nS=pick([1, 2, 5, 17, 64]), N=pick([1, 2, 3, 5, 37, 300, 5001, 20011]), skew=bool(rng.integers(0, 2)),
             ep_len=pick([3, 10]), near_det=bool(rng.integers(0, 3) == 0), agent=pick(["qlearn32", "qlearn32", "qlearn64", "evalmc", "expsarsa32"]),
             lanes=pick([1, 1, 2, 4, 8, 32]), hot=pick([False, True, "tail", "tail"]), aos=bool(rng.integers(0, 2)), slices=pick([0, 0, 3]),
             ring=bool(rng.integers(0, 2)), budget=pick([0, 5, 40]), R=pick([1, 33, 200]), seed=int(rng.integers(0, 2**40)),
             pi_zeros=bool(rng.integers(0, 3) == 0))
    if c["agent"] in ("qlearn64", "evalmc"):
        c["aos"] = False                      # the AoS table holds float32 Q rows of a learner
    if c["budget"] == 0:
        c["slices"] = 0                       # time slices hand out a budget
    if c["nA"] > 4:
        c["hot"], c["aos"] = False, False     # hot records and the AoS table are built for <= 4 actions
    return c


@pytest.mark.parametrize("k", range(int(os.environ.get("OLS_FUZZ_N", "240"))))     # OLS_FUZZ_N=3000 for a longer hunt
def test_randomized_small_configurations_match_c_oracle(ob, k):
    """Differential fuzzing of the fused drivers against oracle/ols_oracle.c over small random shapes: 1-20011 rows, 1-64 states,
    2-4 actions, every lane-group width, hot-record kind, state layout, slicing, ring and budget setting the API accepts."""
    import torch
    from oracle import c_oracle as CO
    c = _fuzz_config(k)
    nA, nS, N, R = c["nA"], c["nS"], c["N"], c["R"]
    rng = np.random.default_rng(c["seed"] % 2**32)
    g = make_log(int(c["seed"] % 1000), N, nS, nA, ep_len=c["ep_len"], skew=c["skew"])
    if c["near_det"]:
        pb = np.full((nS, nA), 0.02 / (nA - 1), dtype=np.float32)
        pb[np.arange(nS), rng.integers(0, nA, nS)] = 0.98
        g["p_log"] = pb[g["z"]]
        g["a"] = (rng.random(N)[:, None] > np.cumsum(pb[g["z"]].astype(np.float64), 1)).sum(1).clip(0, nA - 1).astype(np.int64)
    pi = rng.dirichlet(np.ones(nA), size=nS)
    if c["pi_zeros"]:                         # exact zeros in the target policy: thr = 0 for those actions
        pi[::2, 0] = 0.0
        pi /= pi.sum(1, keepdims=True)
    n_ep, ep_cap, cap = 20, 32, 64
    try:
        log = device_log(ob, g, n_buckets=nS, hot_records=c["hot"])
        sim = ob.BatchedPSRS(log, R, rng="philox", seed=c["seed"], ring=c["ring"], ring_seed=3, replica_id0=12345, lanes=c["lanes"],
                             aos_state=c["aos"], slices=c["slices"])
        kw = dict(gamma=0.95, max_steps=c["budget"], ep_cap=ep_cap, trace_cap=cap)
        if c["agent"].startswith("qlearn"):
            qd = torch.float32 if c["agent"].endswith("32") else torch.float64
            res = sim.run_qlearn(nS, n_ep, alpha=0.2, epsilon=0.15, q_dtype=qd, **kw)
        elif c["agent"] == "evalmc":
            res = sim.run_evalmc(nS, n_ep, pi=pi, **kw)
        else:
            res = sim.run_expsarsa(nS, n_ep, pi=pi, alpha=0.2, q_dtype=torch.float32, **kw)
    except (ValueError, RuntimeError) as e:
        if "CUDA" in str(e) or "cuda" in str(e):
            raise
        pytest.skip("combination refused by the API: %s" % str(e)[:80])
    clog = CO.Log(z=g["z"], z_next=g["z_next"], action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"],
                  steps=g["t"], n_buckets=nS)
    okw = dict(gamma=0.95, seed=c["seed"], replica_id0=12345, ring=c["ring"], ring_seed=3, max_steps=c["budget"], ep_cap=ep_cap,
               trace_cap=cap, threads=4)
    if c["agent"].startswith("qlearn"):
        ref = CO.run(clog, R, CO.QLEARN, nS, n_ep, alpha=0.2, epsilon=0.15, q_f32=c["agent"].endswith("32"), **okw)
    elif c["agent"] == "evalmc":
        ref = CO.run(clog, R, CO.FIXED_PI, nS, n_ep, pi=pi, **okw)
    else:
        ref = CO.run(clog, R, CO.EXPSARSA, nS, n_ep, pi=pi, alpha=0.2, q_f32=True, **okw)
    tag = str(c)
    for name, got in (("accepted", res.accepted), ("pops", res.pops), ("draws", sim.draws), ("status", res.status),
                      ("n_returns", res.n_returns), ("n_lengths", res.n_lengths)):
        assert np.array_equal(got.cpu().numpy(), ref[name]), (name, tag)
    if c["agent"] != "evalmc":
        assert np.array_equal(sim.ties.cpu().numpy(), ref["ties"]), ("ties", tag)
    n_tr = np.minimum(ref["accepted"], cap)
    mask = np.arange(cap)[None, :] < n_tr[:, None]
    assert np.array_equal(res.trace_row.cpu().numpy()[mask], ref["trace_row"][mask]), ("trace", tag)
    mask = np.arange(ep_cap)[None, :] < np.minimum(ref["n_lengths"], ep_cap)[:, None]
    assert np.array_equal(res.lengths.cpu().numpy()[mask], ref["lengths"][mask]), ("lengths", tag)
    mask = np.arange(ep_cap)[None, :] < np.minimum(ref["n_returns"], ep_cap)[:, None]
    np.testing.assert_allclose(res.Gs.cpu().numpy()[mask], ref["Gs"][mask], rtol=1e-12, atol=1e-12, err_msg=tag)
    if c["agent"] != "evalmc":
        assert np.array_equal(res.Q.cpu().numpy(), ref["Q"]), ("Q", tag)


@pytest.mark.parametrize("lanes,aos,hot,slices", [(1, True, "tail", 4), (1, True, True, 0), (4, False, False, 0), (32, True, "tail", 3)])
def test_float32_learner_matches_c_oracle_on_ragged_queues(ob, lanes, aos, hot, slices):
    """The learners on a log with Zipf-distributed queue lengths (one record to tens of thousands), states that were never
    logged as from-states (the reference's KeyError: OLS_NO_QUEUE at the transition that leads there) and queues that run
    dry in mid-episode (partial returns are appended, psrs.py:160-162,177) — every counter, status, row, return and Q table
    against oracle/ols_oracle.c in its float32-learner mode."""
    import torch
    from oracle import c_oracle as CO
    from offsim4rl_b200 import _abi as abi
    N, nS, R, budget = 100003, 400, 1500, 300
    rng = np.random.default_rng(5)
    w = 1.0 / np.arange(1, nS + 1) ** 1.8
    w /= w.sum()
    g = make_log(102, N, nS, 4, ep_len=15)
    z = rng.choice(nS, size=N, p=w).astype(np.int64)
    pb = rng.dirichlet(np.ones(4), size=nS).clip(0.05)
    pb = (pb / pb.sum(1, keepdims=True)).astype(np.float32)                    # one logged distribution per state
    g["z"], g["z_next"] = z, rng.choice(nS, size=N, p=w).astype(np.int64)
    g["p_log"] = pb[z]
    g["a"] = (rng.random(N)[:, None] > np.cumsum(pb[z].astype(np.float64), 1)).sum(1).clip(0, 3).astype(np.int64)
    log = device_log(ob, g, n_buckets=nS, hot_records=hot)
    assert hot != "tail" or log.hot_kind == abi.HOT_TAIL
    sim = ob.BatchedPSRS(log, R, rng="philox", seed=313, ring=True, ring_seed=17, replica_id0=3 * 10**5, lanes=lanes,
                         aos_state=aos, slices=slices)
    res = sim.run_qlearn(nS, 255, gamma=0.99, alpha=0.1, epsilon=0.1, q_dtype=torch.float32, max_steps=budget, ep_cap=256,
                         trace_cap=budget)
    clog = CO.Log(z=g["z"], z_next=g["z_next"], action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"],
                  steps=g["t"], n_buckets=nS)
    ref = CO.run(clog, R, CO.QLEARN, nS, 255, gamma=0.99, alpha=0.1, epsilon=0.1, seed=313, replica_id0=3 * 10**5, ring=True,
                 ring_seed=17, max_steps=budget, ep_cap=256, trace_cap=budget, threads=16, q_f32=True)
    assert np.array_equal(res.accepted.cpu().numpy(), ref["accepted"])
    assert np.array_equal(res.pops.cpu().numpy(), ref["pops"]) and np.array_equal(sim.draws.cpu().numpy(), ref["draws"])
    st_gpu, st_ref = res.status.cpu().numpy(), ref["status"]
    bad = np.flatnonzero(st_gpu != st_ref)
    assert bad.size == 0, (bad[:10], st_gpu[bad[:10]], st_ref[bad[:10]], ref["accepted"][bad[:10]])
    assert np.array_equal(sim.ties.cpu().numpy(), ref["ties"])
    assert np.array_equal(res.trace_row.cpu().numpy(), ref["trace_row"])
    assert np.array_equal(res.n_returns.cpu().numpy(), ref["n_returns"])
    nl = ref["n_lengths"]
    assert np.array_equal(res.n_lengths.cpu().numpy(), nl)
    mask = np.arange(256)[None, :] < nl[:, None]
    assert np.array_equal(res.lengths.cpu().numpy()[mask], ref["lengths"][mask])
    mask = np.arange(256)[None, :] < ref["n_returns"][:, None]
    np.testing.assert_allclose(res.Gs.cpu().numpy()[mask], ref["Gs"][mask], rtol=1e-12, atol=1e-12)
    assert np.array_equal(res.Q.cpu().numpy(), ref["Q"])
    assert (st_ref == abi.NO_QUEUE).sum() > 0 and (st_ref == abi.QUEUE_EMPTY).sum() > 0 and (st_ref == abi.RUNNING).sum() > 0


@pytest.mark.parametrize("ring", [True, False])
@pytest.mark.parametrize("N,nS", [(100003, 400), (4099, 37)])
def test_streaming_evaluator_matches_c_oracle_on_ragged_queues(ob, ring, N, nS):
    """eval_kernel (one warp per replica over TAIL hot records; rounds aligned to the 32-byte sectors of the hot array, stream
    words aligned to blocks of four draws) on the geometry that stresses that alignment: Zipf-distributed queue lengths from one
    record to thousands, a row count that is not a multiple of four (the array's last sector is partial), ring offsets that
    wrap inside a sector, long reject chains that exhaust queues in mid-episode, resumed launches — against
    oracle/ols_oracle.c draw for draw."""
    import torch
    from oracle import c_oracle as CO
    from offsim4rl_b200 import _abi as abi
    rng = np.random.default_rng(N)
    w = 1.0 / np.arange(1, nS + 1) ** 1.8
    w /= w.sum()
    g = make_log(101, N, nS, 4, ep_len=15)
    g["z"] = rng.choice(nS, size=N, p=w).astype(np.int64)
    g["z_next"] = rng.choice(nS, size=N, p=w).astype(np.int64)
    d = 1e-2
    pb = np.full((nS, 4), d / 3, dtype=np.float32)
    pb[np.arange(nS), rng.integers(0, 4, nS)] = 1 - d
    g["p_log"] = pb[g["z"]]
    g["a"] = (rng.random(N)[:, None] > np.cumsum(pb[g["z"]], 1)).sum(1).clip(0, 3).astype(np.int64)
    pi = np.full((nS, 4), 0.25)
    pi[::5] = [0.7, 0.1, 0.1, 0.1]
    pi[3::11] = [0.0, 0.5, 0.5, 0.0]                                   # exact zeros: thr = 0 for two actions
    R = 1200
    log = device_log(ob, g, n_buckets=nS, hot_records="tail")
    assert log.hot_kind == abi.HOT_TAIL
    lens = np.bincount(g["z"], minlength=nS)
    assert lens[lens > 0].min() <= 3 and lens.max() > 300 and N % 4 != 0
    sim = ob.BatchedPSRS(log, R, rng="philox", seed=2024, ring=ring, ring_seed=11, replica_id0=7 * 10**5, lanes=32)
    res = None
    for budget in (7, 40):                                             # the second launch resumes in mid-episode
        res = sim.run_evalmc(nS, 63, pi=pi, gamma=0.97, max_steps=budget, ep_cap=64, trace_cap=47, outputs=res)
    clog = CO.Log(z=g["z"], z_next=g["z_next"], action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"],
                  steps=g["t"], n_buckets=nS)
    ref = CO.run(clog, R, CO.FIXED_PI, nS, 63, gamma=0.97, pi=pi, seed=2024, replica_id0=7 * 10**5, ring=ring, ring_seed=11,
                 max_steps=47, ep_cap=64, trace_cap=47, threads=8)
    assert np.array_equal(res.accepted.cpu().numpy(), ref["accepted"])
    assert np.array_equal(res.pops.cpu().numpy(), ref["pops"]) and np.array_equal(sim.draws.cpu().numpy(), ref["draws"])
    st_gpu, st_ref = res.status.cpu().numpy(), ref["status"]
    bad = np.flatnonzero(st_gpu != st_ref)
    assert bad.size == 0, (bad[:10], st_gpu[bad[:10]], st_ref[bad[:10]], ref["accepted"][bad[:10]], sim.cur_z.cpu().numpy()[bad[:10]])
    assert np.array_equal(res.trace_row.cpu().numpy(), ref["trace_row"])                 # bit-exact accepted rows
    assert np.array_equal(res.n_returns.cpu().numpy(), ref["n_returns"])
    nl = ref["n_lengths"]
    assert np.array_equal(res.n_lengths.cpu().numpy(), nl)
    mask = np.arange(64)[None, :] < nl[:, None]
    assert np.array_equal(res.lengths.cpu().numpy()[mask], ref["lengths"][mask])
    mask = np.arange(64)[None, :] < ref["n_returns"][:, None]
    np.testing.assert_allclose(res.Gs.cpu().numpy()[mask], ref["Gs"][mask], rtol=1e-12, atol=1e-12)
    st = ref["status"]
    assert (st == abi.QUEUE_EMPTY).sum() > 0 and int(ref["pops"].sum()) > 10 * int(ref["accepted"].sum())


@pytest.mark.parametrize("lanes,ring,seeded,pi_dtype", [(32, True, None, np.float64), (8, False, 4, np.float32), (2, True, None, np.float64),
                                                        (16, True, 7, np.float32)])
def test_hot_records_never_change_an_evaluation(ob, lanes, ring, seeded, pi_dtype):
    """Fixed-policy evaluation (evalMC_psrs) over hot records, incl. the near-deterministic behaviour policy of the
    rejection-heavy configuration (long reject chains), float32 and float64 policies: identical rows, returns, lengths."""
    import torch
    nS, nA, N, R = 60, 4, 200000, 1500
    rng = np.random.default_rng(33)
    g = make_log(72, N, nS, nA, ep_len=12, p_dtype=np.float32)
    d = 1e-2
    pb = np.full((nS, nA), d / 3, dtype=np.float32)
    pb[np.arange(nS), rng.integers(0, nA, nS)] = 1 - d
    g["p_log"] = pb[g["z"]]
    g["a"] = (rng.random(N)[:, None] > np.cumsum(pb[g["z"]], 1)).sum(1).clip(0, nA - 1)
    pi = np.full((nS, nA), 0.25, dtype=pi_dtype)
    pi[::7] = np.array([0.7, 0.1, 0.1, 0.1], dtype=pi_dtype)
    from offsim4rl_b200 import _abi as abi
    outs = []
    for hot, trace in ((False, 64), (True, 64), ("tail", 64), ("tail", 0)):     # trace 0: the TAIL path never reads a full record
        log = device_log(ob, g, n_buckets=nS, hot_records=hot).seeded(seeded)
        assert hot != "tail" or log.hot_kind == abi.HOT_TAIL                    # the behaviour distribution is a function of the state
        sim = ob.BatchedPSRS(log, R, rng="philox", seed=8, ring=ring, ring_seed=5, lanes=lanes)
        res = None
        for budget in (25, 15):
            res = sim.run_evalmc(nS, 40, pi=pi, gamma=0.95, max_steps=budget, trace_cap=trace, outputs=res)
        outs.append((res, sim.draws.clone()))
    (a, da) = outs[0]
    for b, db in outs[1:]:
        assert b.trace_row is None or torch.equal(a.trace_row, b.trace_row)
        assert torch.equal(a.accepted, b.accepted) and torch.equal(a.pops, b.pops)
        assert torch.equal(da, db) and torch.equal(a.status, b.status) and torch.equal(a.Gs, b.Gs) and torch.equal(a.lengths, b.lengths)
    assert int(a.accepted.sum()) > 0.5 * R * 40 and int(a.pops.sum()) > 20 * int(a.accepted.sum())
