"""BASELINE.json configs 0-2 at their named shapes, end to end through the public API: observations -> device encoder ->
bucketing -> batched fused run, every replica compared with the C oracle (oracle/ols_oracle.c) draw for draw.
(configs 3 and 4 — the 1e8-row logs — are in tests/test_gpu_fullsize.py.)

  C1  continuous_grid (gridworld.py:205-211 observations, 25 states, 5 actions), grid discretiser, ONE replica, Q-learning
  C2  cartpole-shaped log, CartpoleBoxEncoder ids (162 boxes + the -1 sentinel), 2 actions, 10k replicas
  C3  the C1 log with latent-state MLP ids (dO = 2, H = 64, nZ = 25 / 50), 1M transitions, 64k replicas
"""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import psrs_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import offsim4rl_b200
    return offsim4rl_b200


def grid_log(n, seed=0, cells=5):
    """A random-agent rollout of MyGridNaviCoords in column form: uniform behaviour policy over 5 actions
    (examples/continuous_grid/random_agent_rollout.py:65), reward -0.1 / +1 at the goal cell (gridworld.py:104-110),
    observation = cell / 5 + 0.1 + U(-0.1, 0.1) (gridworld.py:205-211)."""
    rng = np.random.default_rng(seed)
    x, y = rng.integers(0, cells, n), rng.integers(0, cells, n)
    a = rng.integers(0, 5, n)
    dx = np.array([0, 1, -1, 0, 0])[a]
    dy = np.array([0, 0, 0, 1, -1])[a]
    x2, y2 = np.clip(x + dx, 0, cells - 1), np.clip(y + dy, 0, cells - 1)
    done = (x2 == cells - 1) & (y2 == cells - 1)

    def observe(cx, cy):
        o = np.stack([cx, cy], 1) / cells + 0.1 + rng.uniform(-0.1, 0.1, (n, 2))
        return o.astype(np.float32)

    r = np.where(done, 1.0, -0.1).astype(np.float32)
    t = np.where(rng.random(n) < 0.05, 0, 1).astype(np.int64)      # 5 % of the rows open an episode
    t[0] = 0
    return dict(obs=observe(x, y), next_obs=observe(x2, y2), cell=x + cells * y, next_cell=x2 + cells * y2, a=a, r=r, done=done,
                t=t, p_log=np.full((n, 5), 0.2, dtype=np.float32))


def compare_runs(res, sim, ref, learner):
    assert np.array_equal(res.accepted.cpu().numpy(), ref["accepted"])
    assert np.array_equal(res.pops.cpu().numpy(), ref["pops"])
    assert np.array_equal(sim.draws.cpu().numpy(), ref["draws"])
    assert np.array_equal(res.status.cpu().numpy(), ref["status"])
    assert np.array_equal(res.trace_row.cpu().numpy(), ref["trace_row"])          # bit-exact accepted rows
    assert np.array_equal(res.n_returns.cpu().numpy(), ref["n_returns"])
    assert np.array_equal(res.lengths.cpu().numpy(), ref["lengths"])
    np.testing.assert_allclose(res.Gs.cpu().numpy(), ref["Gs"], rtol=1e-12, atol=1e-12)
    if learner:
        assert np.array_equal(sim.ties.cpu().numpy(), ref["ties"])
        assert np.array_equal(res.Q.cpu().numpy(), ref["Q"])                      # float64 Q tables, bit for bit


def test_c1_continuous_grid_single_replica_qlearning(ob):
    import torch
    from offsim4rl_b200.encoders import GridEncoder
    N, nS, nA = 1_000_000, 25, 5
    g = grid_log(N, seed=1)
    enc = GridEncoder(cells=5)
    z, zn = enc.encode_device(g["obs"]), enc.encode_device(g["next_obs"])
    z_h, zn_h = z.cpu().numpy().astype(np.int64), zn.cpu().numpy().astype(np.int64)
    assert np.array_equal(z_h, O.grid_ids(g["obs"])) and np.array_equal(zn_h, O.grid_ids(g["next_obs"]))
    assert (z_h == g["cell"]).mean() > 0.9999                         # ground truth = the logged info['z'], up to boundary rounding
    log = ob.BucketedLog(z=z, z_next=zn, action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"], steps=g["t"],
                         n_buckets=nS)
    clog = CO.Log(z=z_h, z_next=zn_h, action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"], steps=g["t"], n_buckets=nS)
    assert np.array_equal(log.offsets.cpu().numpy(), clog.offsets) and np.array_equal(log.order.cpu().numpy(), clog.order[:N])
    for lanes in (1, 32):
        sim = ob.BatchedPSRS(log, 1, rng="philox", seed=99, lanes=lanes)
        res = sim.run_qlearn(nS, 2000, gamma=0.99, alpha=0.1, epsilon=0.1, q_dtype=torch.float64, trace_cap=20000)
        ref = CO.run(clog, 1, CO.QLEARN, nS, 2000, gamma=0.99, alpha=0.1, epsilon=0.1, seed=99, trace_cap=20000, threads=1)
        assert int(res.accepted[0]) > 5000
        compare_runs(res, sim, ref, learner=True)


def test_c2_cartpole_boxes_10k_replicas(ob):
    import torch
    from offsim4rl_b200.encoders import CartpoleBoxEncoder
    N, nA, R = 1_000_000, 2, 10_000
    rng = np.random.default_rng(2)
    lim = np.array([2.4, 1.5, 0.2094384, 1.5])
    obs = rng.uniform(-lim, lim, (N, 4)).astype(np.float32)
    next_obs = rng.uniform(-lim * [1.04, 1, 1.03, 1], lim * [1.04, 1, 1.03, 1], (N, 4)).astype(np.float32)
    enc = CartpoleBoxEncoder()
    z, zn = enc.encode_device(obs), enc.encode_device(next_obs)
    z_h, zn_h = z.cpu().numpy().astype(np.int64), zn.cpu().numpy().astype(np.int64)
    assert np.array_equal(z_h, O.cartpole_box(obs)) and np.array_equal(zn_h, O.cartpole_box(next_obs))
    assert z_h.min() >= 0 and zn_h.min() == -1 and z_h.max() == 161
    done = zn_h < 0                                                    # CartPole terminates outside the failure bounds
    pb = rng.dirichlet(np.ones(nA), size=162).clip(0.1)
    pb /= pb.sum(1, keepdims=True)
    p_log = pb[z_h].astype(np.float32)
    a = (rng.random(N) > pb[z_h, 0]).astype(np.int64)
    r = np.ones(N, dtype=np.float32)
    start = np.r_[True, done[:-1]]
    idx = np.arange(N)
    t = idx - np.maximum.accumulate(np.where(start, idx, 0))
    # the tabular agent indexes its table by box + 1 (163 rows: the sentinel is row 0), the queues by the box id itself
    s_h, sn_h = z_h + 1, zn_h + 1
    nS = 163
    log = ob.BucketedLog(z=z, z_next=zn, s=s_h, s_next=sn_h, action=a, reward=r, terminal=done, p_log=p_log, steps=t,
                         n_buckets=163, z_offset=1)
    clog = CO.Log(z=z_h, z_next=zn_h, s=s_h, s_next=sn_h, action=a, reward=r, terminal=done, p_log=p_log, steps=t, n_buckets=163,
                  z_offset=1)
    assert np.array_equal(log.offsets.cpu().numpy(), clog.offsets) and np.array_equal(log.order.cpu().numpy(), clog.order[:N])
    pi = rng.dirichlet(np.ones(nA), size=nS)
    kw = dict(seed=7, ring=True, ring_seed=3)
    sim = ob.BatchedPSRS(log, R, rng="philox", lanes=8, **kw)
    res = sim.run_evalmc(nS, 20, pi=pi, gamma=0.99, trace_cap=800)
    ref = CO.run(clog, R, CO.FIXED_PI, nS, 20, gamma=0.99, pi=pi, trace_cap=800, threads=8, **kw)
    assert int(res.accepted.sum()) > 20 * R
    compare_runs(res, sim, ref, learner=False)
    sim = ob.BatchedPSRS(log, R, rng="philox", lanes=8, **kw)
    res = sim.run_qlearn(nS, 20, gamma=0.99, alpha=0.1, epsilon=0.1, q_dtype=torch.float64, trace_cap=800)
    ref = CO.run(clog, R, CO.QLEARN, nS, 20, gamma=0.99, alpha=0.1, epsilon=0.1, trace_cap=800, threads=8, **kw)
    compare_runs(res, sim, ref, learner=True)


@pytest.mark.parametrize("n_latent", [25, 50])
def test_c3_latent_state_mlp_64k_replicas(ob, n_latent):
    import torch
    from offsim4rl_b200.encoders import HOMEREncoder
    N, nA, R, H = 1_000_000, 5, 65_536, 64
    g = grid_log(N, seed=3)
    # weights of the reference's architecture (models.py:15-19), trained for a moment on the cell labels so that every latent
    # state is in use (a random-init net of this shape maps the unit square to 3-6 latent states); the training is a fixture,
    # plain torch on the device, and is not part of the path under test
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(2, H), torch.nn.LeakyReLU(), torch.nn.Linear(H, n_latent)).cuda()
    xs = torch.from_numpy(g["obs"][:50_000]).cuda()
    label = g["cell"][:50_000] if n_latent == 25 else g["cell"][:50_000] * 2 + (g["obs"][:50_000, 0] * 5 % 1 > 0.5)
    ys = torch.from_numpy(label).cuda()
    opt = torch.optim.Adam(net.parameters(), lr=0.02)
    for _ in range(300):
        opt.zero_grad()
        torch.nn.functional.cross_entropy(net(xs), ys).backward()
        opt.step()
    W1, b1, W2, b2 = (p.detach().cpu().numpy() for p in net.parameters())
    enc = HOMEREncoder.from_weights(W1, b1, W2, b2)
    z, zn = enc.encode_device(g["obs"]), enc.encode_device(g["next_obs"])
    z_h, zn_h = z.cpu().numpy().astype(np.int64), zn.cpu().numpy().astype(np.int64)
    # FMA-order oracle, bit-exact (a slice: the scalar C restatement takes ~5 us per row)
    assert np.array_equal(z_h[:250_000], CO.mlp_ids(g["obs"][:250_000], W1, b1, W2, b2))
    assert np.array_equal(zn_h[-100_000:], CO.mlp_ids(g["next_obs"][-100_000:], W1, b1, W2, b2))
    with torch.no_grad():                                              # and the torch forward the reference runs (near ties only)
        t_ids = torch.argmax(torch.log_softmax(net(torch.from_numpy(g["obs"]).cuda()), -1), -1).cpu().numpy()
    assert (t_ids != z_h).mean() < 1e-3
    assert len(np.unique(z_h)) >= n_latent // 2
    nS = n_latent
    log = ob.BucketedLog(z=z, z_next=zn, action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"], steps=g["t"],
                         n_buckets=nS)
    clog = CO.Log(z=z_h, z_next=zn_h, action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"], steps=g["t"], n_buckets=nS)
    assert np.array_equal(log.offsets.cpu().numpy(), clog.offsets) and np.array_equal(log.order.cpu().numpy(), clog.order[:N])
    kw = dict(seed=11, ring=True, ring_seed=5)
    sim = ob.BatchedPSRS(log, R, rng="philox", lanes=4, **kw)
    res = sim.run_qlearn(nS, 30, gamma=0.99, alpha=0.1, epsilon=0.1, q_dtype=torch.float64, max_steps=150, trace_cap=150)
    ref = CO.run(clog, R, CO.QLEARN, nS, 30, gamma=0.99, alpha=0.1, epsilon=0.1, max_steps=150, trace_cap=150, threads=8, **kw)
    assert int(res.accepted.sum()) > 100 * R
    compare_runs(res, sim, ref, learner=True)
