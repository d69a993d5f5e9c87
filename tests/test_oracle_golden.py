"""Pin the oracle (oracle/psrs_oracle.py) to vectors produced by the unmodified reference.

The fixtures come from tests/golden/make_golden.py (run once in the build container against
/root/reference).  Everything here is CPU-only.
"""
import numpy as np
import pytest

from oracle import psrs_oracle as O


def _env(g, **kw):
    steps = g["t"] if ("with_t" not in g.files or bool(g["with_t"])) else None
    return O.OraclePSRS(z=g["z"], z_next=g["z_next"], s=g["z"], s_next=g["z_next"], action=g["a"],
                        reward=g["r"], terminal=g["done"], p_log=g["p_log"], steps=steps, **kw)


# --------------------------------------------------------------------------- reset_sampler(seed) + own RNG
@pytest.mark.parametrize("seed", [0, 3, 98052])
def test_seeded_queues_and_walk(golden, seed):
    g = golden("seeded_walk")
    nS = 9
    env = _env(g, n_buckets=nS, seed=seed)
    assert np.array_equal(env.offsets, g[f"offsets_{seed}"])
    assert np.array_equal(env.order, g[f"order_{seed}"])                     # psrs.py:26,29-30
    init = g[f"init_zs_{seed}"]
    assert np.array_equal(g["z"][env.init], init[:, 0])                       # psrs.py:22-23
    # walk with the restated default_rng(seed).random() stream (psrs.py:20,56)
    p_new = g["p_new"]
    trace, resets = [], []
    s = env.reset()
    resets.append(s)
    while s is not None:
        row = env.step(p_new)
        if row is None:
            break
        trace.append(row)
        s = env.cur_s
        if g["done"][row]:
            s = env.reset()
            resets.append(-1 if s is None else s)
    assert np.array_equal(trace, g[f"trace_{seed}"])
    assert np.array_equal(np.array(resets, dtype=np.int64), g[f"resets_{seed}"])
    assert env.pops == int(g[f"pops_{seed}"])


# --------------------------------------------------------------------------- drivers under injected streams
DRIVER_CASES = ["drivers_f64", "drivers_f32", "drivers_p64_a5", "drivers_edge"]


@pytest.mark.parametrize("case", DRIVER_CASES)
@pytest.mark.parametrize("tag", ["const", "sched"])
def test_qlearn(golden, case, tag):
    g = golden(case)
    nS, nA = g["pi"].shape
    env = _env(g, n_buckets=nS, seed=int(g["sampler_seed"]), uniforms=O.FeedUniforms(g["U"]))
    ties = O.FeedTies(g["T"])
    if tag == "const":
        eps, alpha = 0.2, 0.1
    else:
        eps = lambda e: float(g["ql_sched_eps"][e])
        alpha = lambda e: float(g["ql_sched_alpha"][e])
    with np.errstate(all="ignore"):
        Q, info = O.qlearn(env, 25, gamma=0.9, alpha=alpha, epsilon=eps, ties=ties, nS=nS, nA=nA)
    assert np.array_equal(info["trace"], g[f"ql_{tag}_trace"])
    assert env.uniforms.k == int(g[f"ql_{tag}_draws"])
    assert ties.k == int(g[f"ql_{tag}_ties"])
    assert np.array_equal(Q, g[f"ql_{tag}_Q"])                               # same float64 op order
    assert np.array_equal(info["Gs"], g[f"ql_{tag}_Gs"])
    assert np.array_equal(info["TD_errors"], g[f"ql_{tag}_TD"])


@pytest.mark.parametrize("case", DRIVER_CASES)
def test_expsarsa(golden, case):
    g = golden(case)
    nS, nA = g["pi"].shape
    env = _env(g, n_buckets=nS, seed=int(g["sampler_seed"]), uniforms=O.FeedUniforms(g["U"]))
    with np.errstate(all="ignore"):
        Q, info = O.expsarsa(env, 25, pi=g["pi"], gamma=0.95, alpha=0.2, nS=nS, nA=nA)
    assert np.array_equal(info["trace"], g["es_trace"])
    assert env.uniforms.k == int(g["es_draws"])
    assert np.array_equal(Q, g["es_Q"])
    assert np.array_equal(info["Gs"], g["es_Gs"])


@pytest.mark.parametrize("case", DRIVER_CASES + ["drivers_float_a"])
def test_evalmc(golden, case):
    g = golden(case)
    nS, nA = g["pi"].shape
    env = _env(g, n_buckets=nS, seed=int(g["sampler_seed"]), uniforms=O.FeedUniforms(g["U"]))
    with np.errstate(all="ignore"):
        Gs, lengths, trace = O.evalmc(env, 40, pi=g["pi"], gamma=0.97)
    assert np.array_equal(trace, g["mc_trace"])
    assert env.uniforms.k == int(g["mc_draws"])
    assert np.array_equal(Gs, g["mc_Gs"])
    assert np.array_equal(lengths, g["mc_lengths"])


def test_accept_rule_table(golden):
    tab = golden("accept_rule")["table"]
    assert tab.shape[0] > 400
    with np.errstate(all="ignore"):
        for pn0, pn1, pl0, pl1, a, n32, l32, u, rej in tab:
            p_new = np.array([pn0, pn1], dtype=np.float32 if n32 else np.float64)
            p_log = np.array([pl0, pl1], dtype=np.float32 if l32 else np.float64)
            assert O.rejects(float(u), p_new, p_log, int(a)) == bool(rej)


# --------------------------------------------------------------------------- encoders
def test_cartpole_box_ids(golden):
    g = golden("encoders")
    assert np.array_equal(O.cartpole_box(g["box_obs32"]), g["box_ids32"])     # encoders/heuristic.py:20-71
    assert np.array_equal(O.cartpole_box(g["box_obs64"]), g["box_ids64"])
    assert (g["box_ids32"] == -1).any() and g["box_ids32"].max() <= 161


@pytest.mark.parametrize("tag", ["mlp_grid", "mlp_cart", "mlp_wide"])
def test_mlp_ids_match_torch_except_near_ties(golden, tag):
    """The oracle fixes the accumulation order (bias, then fmaf in index order); torch's sgemm uses its
    own.  The ids must agree wherever the float64 top-2 logit gap is not within rounding noise."""
    g = golden("encoders")
    x, W1, b1, W2, b2 = (g[f"{tag}_{k}"] for k in ["x", "W1", "b1", "W2", "b2"])
    logits = O.mlp_logits(x, W1, b1, W2, b2)
    assert np.allclose(logits, g[f"{tag}_logits"], rtol=1e-5, atol=1e-6)
    ids = np.argmax(logits, axis=1)
    bad = np.nonzero(ids != g[f"{tag}_ids"])[0]
    h = x.astype(np.float64) @ W1.T.astype(np.float64) + b1
    h = np.where(h > 0, h, 0.01 * h)
    l64 = h @ W2.T.astype(np.float64) + b2
    top2 = np.sort(l64, axis=1)[:, -2:]
    gap = top2[:, 1] - top2[:, 0]
    assert bad.size <= 2 and np.all(gap[bad] < 1e-5)


def test_fmaf_emulation_is_correctly_rounded():
    from fractions import Fraction
    rng = np.random.default_rng(0)
    a = rng.normal(size=300).astype(np.float32)
    b = rng.normal(size=300).astype(np.float32)
    c = (-(a.astype(np.float64) * b) + rng.normal(size=300) * 1e-7).astype(np.float32)   # heavy cancellation
    got = O._fmaf(a, b, c)
    for i in range(300):
        exact = Fraction(float(a[i])) * Fraction(float(b[i])) + Fraction(float(c[i]))
        lo = np.float32(float(exact))          # float(Fraction) is correctly rounded to f64; not enough alone
        cands = {np.float32(lo), np.nextafter(np.float32(lo), np.float32(np.inf)), np.nextafter(np.float32(lo), np.float32(-np.inf))}
        best = min(cands, key=lambda v: (abs(Fraction(float(v)) - exact), int(np.float32(v).view(np.int32)) & 1))
        assert got[i] == best


# --------------------------------------------------------------------------- reference's own behavioural tests
def test_reference_gridworld_scenario(golden):
    """tests/test_psrs.py:9-34 of the reference: > 40 simulated steps from a 100-transition gridworld
    log under a uniform target; the per-seed counts are reproduced exactly."""
    g = golden("gridworld")
    counts = []
    for seed in range(20):
        env = O.OraclePSRS(z=g["z"], z_next=g["z_next"], s=g["s"], s_next=g["s_next"], action=g["a"], reward=g["r"],
                           terminal=g["done"], p_log=g["p"], steps=g["t"], n_buckets=25, seed=seed)
        obs = env.reset()
        n = -1
        while obs is not None:
            row = env.step(np.array([0.2] * 5))
            n += 1
            if row is None:
                break
            obs = env.cur_s
            if g["done"][row]:
                obs = env.reset()
        counts.append(n)
    assert np.array_equal(counts, g["steps_simulated_by_seed"])
    assert min(counts) > 40


# --------------------------------------------------------------------------- QueueEvaluator
@pytest.mark.parametrize("seed", [0, 4])
def test_queue_evaluator_oracle_matches_reference(golden, seed):
    g = golden("queue_evaluator")
    env = O.OracleQueueEvaluator(z=g["z"], z_next=g["z_next"], s=g["z"], s_next=g["z_next"], action=g["a"], steps=g["t"],
                                 n_state_keys=6, n_actions=3, seed=seed)
    assert np.array_equal(env.offsets, g[f"offsets_{seed}"]) and np.array_equal(env.order, g[f"order_{seed}"])
    trace, resets, k, status = [], [], 0, "exhausted"
    s = env.reset()
    resets.append(s)
    try:
        while s is not None:
            row = env.step(g["acts"][k]); k += 1
            if row is None:
                break
            trace.append(row)
            s = env.cur_s
            if g["done"][row]:
                s = env.reset()
                resets.append(-1 if s is None else s)
    except KeyError:
        status = "keyerror"
    assert np.array_equal(trace, g[f"trace_{seed}"]) and np.array_equal(np.array(resets, dtype=np.int64), g[f"resets_{seed}"])
    assert status == str(g[f"status_{seed}"]) and k == int(g[f"nacts_{seed}"])


# --------------------------------------------------------------------------- PSRS_Exo
@pytest.mark.parametrize("seed", [1, 6])
def test_psrs_exo_oracle_matches_reference(golden, seed):
    g = golden("psrs_exo")
    nX = 3
    env = O.OraclePSRSExo(s=g["z"], x=g["x"], s_next=g["z_next"], x_next=g["x_next"], action=g["a"], p_log=g["p_log"],
                          steps=g["t"], n_s=5, n_x=nX, seed=seed)
    trace, obs, episode = [], [], 0
    row = env.reset(episode)
    obs.append(-1 if row is None else int(g["z"][row]) * nX + int(g["x"][row]))
    while row is not None:
        res = env.step(g["p_new"])
        if res is None:
            break
        row_s, x_next = res
        trace.append(row_s)
        obs.append(int(g["z_next"][row_s]) * nX + x_next)
        row = row_s
        if g["done"][row_s]:
            episode += 1
            row = env.reset(episode)
            obs.append(-1 if row is None else int(g["z"][row]) * nX + int(g["x"][row]))
    assert np.array_equal(trace, g[f"trace_{seed}"]) and np.array_equal(obs, g[f"obs_{seed}"])


@pytest.mark.parametrize("case", DRIVER_CASES)
@pytest.mark.parametrize("tag", ["const", "sched"])
def test_qlearn_float32_table(golden, case, tag):
    """The float32 learner: the reference's own update lines on a float32 Q table (make_golden.py `F32Table`)."""
    g, g32 = golden(case), golden(case + "_f32q")
    nS, nA = g["pi"].shape
    env = _env(g, n_buckets=nS, seed=int(g["sampler_seed"]), uniforms=O.FeedUniforms(g["U"]))
    ties = O.FeedTies(g["T"])
    eps, alpha = (0.2, 0.1) if tag == "const" else ((lambda e: 1.0 / (1 + e)), (lambda e: 0.5 / (1 + 0.1 * e)))
    with np.errstate(all="ignore"):
        Q, info = O.qlearn(env, 25, gamma=0.9, alpha=alpha, epsilon=eps, ties=ties, nS=nS, nA=nA, q_dtype=np.float32)
    assert Q.dtype == np.float32 and np.array_equal(Q, g32[f"ql_{tag}_Q"])
    assert np.array_equal(info["trace"], g32[f"ql_{tag}_trace"])
    assert np.array_equal(info["TD_errors"], g32[f"ql_{tag}_TD"].astype(np.float64))
    assert env.uniforms.k == int(g32[f"ql_{tag}_draws"]) and ties.k == int(g32[f"ql_{tag}_ties"])


def test_back_to_back_drivers_on_one_env(golden):
    """Several drivers on one env without reset_sampler (tests/golden/back_to_back.npz): every call counts its episodes from
    0 and keeps consuming the queues and the streams where the previous call stopped."""
    g = golden("back_to_back")
    nS, nA = g["pi"].shape
    env = _env(g, n_buckets=nS, seed=int(g["sampler_seed"]), uniforms=O.FeedUniforms(g["U"]))
    ties = O.FeedTies(g["T"])
    k1, k2, k3, k4, k5 = (int(v) for v in g["marks"])
    eps_list, alpha_list = list(g["eps_list"]), list(g["alpha_list"])
    Q1, i1 = O.qlearn(env, 10, gamma=0.9, alpha=lambda e: alpha_list[e], epsilon=lambda e: eps_list[e], ties=ties, nS=nS, nA=nA)
    assert np.array_equal(i1["trace"], g["trace"][:k1]) and np.array_equal(Q1, g["Q1"]) and np.array_equal(i1["TD_errors"], g["TD1"])
    assert env.uniforms.k == int(g["draws"][0]) and ties.k == int(g["ties"][0])
    Gs2, len2, tr2 = O.evalmc(env, 12, pi=g["pi"], gamma=0.95)
    assert np.array_equal(tr2, g["trace"][k1:k2]) and np.array_equal(len2, g["len2"])
    np.testing.assert_allclose(Gs2, g["Gs2"], rtol=1e-5, atol=1e-5)
    assert int(env.cur_s) == int(g["s_mid"])
    row = env.step(g["pi"][int(env.cur_s)])
    assert row == int(g["trace"][k2]) and int(env.cur_s) == int(g["s3"]) and env.uniforms.k == int(g["draws"][2])
    Q4, i4 = O.qlearn(env, 6, gamma=0.9, alpha=0.2, epsilon=0.1, ties=ties, nS=nS, nA=nA, q_init=Q1)
    assert np.array_equal(i4["trace"], g["trace"][k3:k4]) and np.array_equal(Q4, g["Q4"])
    Q5, i5 = O.expsarsa(env, 5, pi=g["pi"], gamma=0.9, alpha=0.2, nS=nS, nA=nA)
    assert np.array_equal(i5["trace"], g["trace"][k4:k5])
    np.testing.assert_allclose(Q5, g["Q5"], rtol=1e-12, atol=1e-14)
    assert env.uniforms.k == int(g["draws"][4])
