"""Differential tests against the UNMODIFIED reference, executed live next to the GPU run.

`__graft_entry__.build()` installs the reference under `baseline/_ref` (git-ignored; it travels to the GPU box with the
snapshot, `/root/reference` does not) and these tests load its `evaluators/psrs.py` and `agents/tabular.py` by file path, as
`bench.py --impl reference` does.  Where the committed fixtures under tests/golden/ pin a handful of shapes, this draws random
small logs and runs BOTH sides on them:

  reference:  PSRS(rows).reset_sampler(seed) -> qlearn_psrs / expSARSA_psrs / evalMC_psrs with its own PCG64 generator for the
              queue shuffles and the rejection uniforms (psrs.py:20-30,56); only `np.random.choice` (the legacy global generator
              behind `_random_argmax`, tabular.py:4-5) is replaced by a recorded stream, which the device reads as its tie feed
  device:     BucketedLog(...).seeded(seed) + BatchedPSRS(rng="pcg64", seed=seed): the library replays SeedSequence -> PCG64 ->
              Fisher-Yates / random() itself

and compares accepted rows, draw and tie counts, Q tables and TD errors bit for bit, returns within 1e-5.  Skipped when
`baseline/_ref` is absent."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref", "offsim4rl")


@pytest.fixture(scope="module")
def ref():
    mods = []
    for rel in ("evaluators/psrs.py", "agents/tabular.py"):
        path = os.path.join(REF_DIR, rel)
        if not os.path.exists(path):
            pytest.skip("baseline/_ref is not installed (python -c 'import __graft_entry__ as g; g.build()' where /root/reference exists)")
        spec = importlib.util.spec_from_file_location("offsim4rl_live_" + os.path.basename(rel)[:-3], path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


@pytest.fixture(scope="module")
def ob():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import offsim4rl_b200
    return offsim4rl_b200


class TieFeed:
    """Stands in for `np.random.choice` inside `_random_argmax` (agents/tabular.py:4-5): candidate int(v * n) of a recorded v."""

    def __init__(self, values):
        self.values, self.k = values, 0

    def __call__(self, candidates):
        n = len(candidates)
        if n == 1:
            return candidates[0]
        v = float(self.values[self.k])
        self.k += 1
        return candidates[min(n - 1, int(v * n))]


class F32Table(np.ndarray):
    """A Q_init whose `.copy().astype(float)` (psrs.py:141) stays float32, so the reference's own update lines run in float32
    (NumPy >= 2 promotion): the arithmetic of the production learner.  Nothing in the reference is modified."""

    def astype(self, dtype, *a, **k):
        return np.array(self, dtype=np.float32, copy=True, subok=False)


def random_log(rng, N, nS, nA, p_dtype):
    z = rng.integers(0, nS, size=N)
    if N >= nS:
        z[rng.permutation(N)[:nS]] = np.arange(nS)          # every state has a queue: the reference raises KeyError otherwise
    zn = rng.integers(0, nS, size=N)
    pb = rng.dirichlet(np.ones(nA), size=nS).clip(0.05)
    pb /= pb.sum(1, keepdims=True)
    if rng.integers(0, 2):                                  # per-ROW distributions (TAIL hot records do not apply), else per state
        p_log = rng.dirichlet(np.ones(nA), size=N).clip(0.05)
        p_log = (p_log / p_log.sum(1, keepdims=True)).astype(p_dtype)
        a = (rng.random(N)[:, None] > np.cumsum(p_log.astype(np.float64), 1)).sum(1).clip(0, nA - 1)
    else:
        p_log = pb[z].astype(p_dtype)
        a = (rng.random(N)[:, None] > np.cumsum(pb[z], 1)).sum(1).clip(0, nA - 1)
    done = rng.random(N) < 1.0 / int(rng.integers(2, 9))
    idx = np.arange(N)
    t = idx - np.maximum.accumulate(np.where(np.r_[True, done[:-1]], idx, 0))
    return dict(z=z.astype(np.int64), z_next=zn.astype(np.int64), a=a.astype(np.int64), r=rng.normal(size=N).astype(np.float32),
                done=done, t=t, p_log=p_log)


def legacy_tuples(g, with_t):
    rows, id2row = [], {}
    for i in range(len(g["z"])):
        p = np.array(g["p_log"][i], copy=True)              # a distinct object per row: its id() names the row in `memory`
        id2row[id(p)] = i
        info = {"z": int(g["z"][i]), "z_next": int(g["z_next"][i]), "t": int(g["t"][i]) if with_t else 0}
        rows.append((int(g["z"][i]), g["a"][i], g["r"][i], int(g["z_next"][i]), bool(g["done"][i]), p, info))
    return rows, id2row


def make_case(k):
    rng = np.random.default_rng(7000 + k)
    pick = lambda xs: xs[int(rng.integers(0, len(xs)))]
    c = dict(k=k, nA=pick([2, 3, 4, 4, 5]), nS=pick([1, 3, 9, 20]))
    c["N"] = max(pick([1, 7, 60, 400, 1500]), c["nS"])
    c["with_t"] = bool(rng.integers(0, 4) != 0)
    p_dtype = pick([np.float32, np.float32, np.float64])
    c["seed"] = int(rng.integers(0, 2**31))
    c["agent"] = pick(["qlearn64", "qlearn32", "qlearn32", "expsarsa", "evalmc"])
    c["lanes"], c["hot"] = pick([1, 2, 4, 8, 32]), pick([False, True, "tail"])
    c["aos"] = False          # (the AoS replica state takes Philox / fed uniforms only; its parity is pinned through the C oracle)
    rng.integers(0, 2)
    c["sched"] = bool(rng.integers(0, 2))
    g = random_log(rng, c["N"], c["nS"], c["nA"], p_dtype)
    T = rng.random(16 * c["N"] + 64)
    pi = rng.dirichlet(np.ones(c["nA"]), size=c["nS"]).astype(pick([np.float64, np.float32]))
    return c, g, T, pi


E = 12
EPS = lambda e: 1.0 / (1 + e)
ALPHA = lambda e: 0.5 / (1 + 0.1 * e)


def reference_run(ref, c, g, T, pi):
    """The reference, unmodified, with its own generators; returns what the device run is compared with."""
    ref_psrs, ref_tabular = ref
    nS, nA, N, agent = c["nS"], c["nA"], c["N"], c["agent"]
    eps, alpha = (EPS, ALPHA) if c["sched"] else (0.2, 0.1)
    rows, id2row = legacy_tuples(g, c["with_t"])
    env = ref_psrs.PSRS(rows, nS=nS, nA=nA)
    env.reset_sampler(seed=c["seed"])
    ties, saved = TieFeed(T), np.random.choice
    np.random.choice = ties
    trace, out = [], {}
    try:
        if agent.startswith("qlearn"):
            if agent == "qlearn32":
                policy = lambda Q, args: ref_tabular.epsilon_greedy_policy(np.asarray(Q, dtype=np.float64), args)
                Q_ref, info = ref_psrs.qlearn_psrs(env, n_episodes=E, behavior_policy=policy, gamma=0.9, alpha=alpha, epsilon=eps,
                                                   Q_init=np.zeros((nS, nA), np.float32).view(F32Table))
                assert Q_ref.dtype == np.float32
            else:
                Q_ref, info = ref_psrs.qlearn_psrs(env, n_episodes=E, behavior_policy=ref_tabular.epsilon_greedy_policy, gamma=0.9,
                                                   alpha=alpha, epsilon=eps)
            trace = [id2row[id(m[6]["p"])] for m in info["memory"]]
            out.update(Q=Q_ref, TD=np.asarray(info["TD_errors"], dtype=np.float64), Gs=np.asarray(info["Gs"], dtype=np.float64))
        elif agent == "expsarsa":
            Q_ref, info = ref_psrs.expSARSA_psrs(env, n_episodes=E, pi=pi, gamma=0.95, alpha=0.2)
            trace = [id2row[id(m[6]["p"])] for m in info["memory"]]
            out.update(Q=Q_ref, Gs=np.asarray(info["Gs"], dtype=np.float64))
        else:
            orig_step = env.step

            def traced_step(p):
                res = orig_step(p)
                if res[0] is not None:
                    trace.append(id2row[id(res[3]["p"])])
                return res
            env.step = traced_step
            Gs_ref, len_ref = ref_psrs.evalMC_psrs(env, n_episodes=E, pi=pi, gamma=0.97)
            out.update(Gs=np.asarray(Gs_ref, dtype=np.float64), lengths=np.asarray(len_ref))
    finally:
        np.random.choice = saved
    out.update(trace=np.array(trace, dtype=np.int64), ties=ties.k, pops=N - sum(len(q) for q in env.queues.values()))
    return out


@pytest.mark.parametrize("k", range(int(os.environ.get("OLS_LIVE_FUZZ_N", "120"))))
def test_fused_drivers_equal_the_live_reference(ob, ref, k):
    import torch
    c, g, T, pi = make_case(k)
    nS, nA, N, agent, with_t, seed = c["nS"], c["nA"], c["N"], c["agent"], c["with_t"], c["seed"]
    want = reference_run(ref, c, g, T, pi)

    # ---- the device: the library replays the reference's generators itself
    log = ob.BucketedLog(z=g["z"], z_next=g["z_next"], action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"],
                         steps=g["t"] if with_t else None, n_buckets=nS, hot_records=c["hot"] if nA <= 4 else False).seeded(seed)
    sim = ob.BatchedPSRS(log, 2, rng="pcg64", seed=seed, tie_feed=T, lanes=c["lanes"], aos_state=c["aos"])
    tag = str(c)
    n = want["trace"].size
    if agent.startswith("qlearn"):
        e_arr = np.array([EPS(e) for e in range(E)]) if c["sched"] else 0.2
        a_arr = np.array([ALPHA(e) for e in range(E)]) if c["sched"] else 0.1
        res = sim.run_qlearn(nS, E, gamma=0.9, alpha=a_arr, epsilon=e_arr, q_dtype=torch.float32 if agent == "qlearn32" else torch.float64,
                             trace_cap=max(N, 1))
    elif agent == "expsarsa":
        res = sim.run_expsarsa(nS, E, pi=pi, gamma=0.95, alpha=0.2, q_dtype=torch.float64, trace_cap=max(N, 1))
    else:
        res = sim.run_evalmc(nS, E, pi=pi, gamma=0.97, trace_cap=max(N, 1))
    for r in range(2):                                       # two replicas of the same seed: identical runs
        assert int(res.accepted[r]) == n, tag
        assert np.array_equal(res.trace_row[r, :n].cpu().numpy(), want["trace"]), tag
        assert int(res.pops[r]) == want["pops"], tag
        if agent.startswith("qlearn"):
            assert int(sim.ties[r]) == want["ties"], tag
            assert np.array_equal(res.Q[r].cpu().numpy(), want["Q"]), tag                       # bit for bit, float64 and float32
            assert np.array_equal(res.td_error[r, :n].cpu().numpy(), want["TD"]), tag
        if agent == "expsarsa":
            np.testing.assert_allclose(res.Q[r].cpu().numpy(), want["Q"], rtol=1e-12, atol=1e-14, err_msg=str(tag))
        if agent != "evalmc":
            ne = int(res.n_returns[r])
            assert ne == want["Gs"].size, tag
            np.testing.assert_allclose(res.Gs[r, :ne].cpu().numpy(), want["Gs"], rtol=1e-5, atol=1e-5,
                                       err_msg=str(tag))
        else:
            ne, nl = int(res.n_returns[r]), int(res.n_lengths[r])
            assert ne == want["Gs"].size and nl == want["lengths"].size, tag                           # psrs.py:265-269
            assert np.array_equal(res.lengths[r, :nl].cpu().numpy(), want["lengths"]), tag
            np.testing.assert_allclose(res.Gs[r, :ne].cpu().numpy(), want["Gs"], rtol=1e-5, atol=1e-5,
                                       err_msg=str(tag))


# ------------------------------------------------------------------------------------------ PSRS_Exo (psrs.py:59-117)
def make_exo_case(k):
    rng = np.random.default_rng(9000 + k)
    pick = lambda xs: xs[int(rng.integers(0, len(xs)))]
    c = dict(k=k, nA=pick([2, 3, 4]), nS=pick([1, 3, 5]), nX=pick([1, 2, 4]), seed=int(rng.integers(0, 2**31)))
    c["N"] = max(pick([7, 60, 400, 900]), c["nS"] * c["nX"])
    N, nS, nX, nA = c["N"], c["nS"], c["nX"], c["nA"]
    g = random_log(rng, N, nS, nA, np.float32)
    x, x_next = rng.integers(0, nX, N), rng.integers(0, nX, N)
    x[rng.permutation(N)[:nX]] = np.arange(nX)              # every exogenous state has a queue
    g["x"], g["x_next"] = x.astype(np.int64), x_next.astype(np.int64)
    pi = rng.dirichlet(np.ones(nA), size=nS * nX)
    return c, g, pi


def reference_exo_run(ref, c, g, pi, n_episodes):
    ref_psrs, _ = ref
    nS, nX, nA, N = c["nS"], c["nX"], c["nA"], c["N"]
    o, o_next = g["z"] * nX + g["x"], g["z_next"] * nX + g["x_next"]
    rows = [(int(o[i]), int(g["a"][i]), g["r"][i], int(o_next[i]), bool(g["done"][i]), np.array(g["p_log"][i], copy=True),
             {"t": int(g["t"][i])}) for i in range(N)]
    env = ref_psrs.PSRS_Exo(rows, nO=nS * nX, nA=nA, o_split_func=lambda v: divmod(v, nX), o_combine_func=lambda s, xx: s * nX + xx)
    id2row = {id(e[5]): i for i, e in enumerate(env.raw_buffer)}     # the reference deep-copies the buffer (psrs.py:65)
    env.reset_sampler(seed=c["seed"])
    trace, orig_step = [], env.step

    def traced_step(p):
        res = orig_step(p)
        if res[0] is not None:
            trace.append(id2row[id(res[3]["p"])])
        return res
    env.step = traced_step
    Gs, lengths = ref_psrs.evalMC_psrs(env, n_episodes=n_episodes, pi=pi, gamma=0.9)
    return dict(trace=np.array(trace, dtype=np.int64), Gs=np.asarray(Gs, dtype=np.float64), lengths=np.asarray(lengths))


@pytest.mark.parametrize("k", range(30))
def test_fused_psrs_exo_evaluation_equals_the_live_reference(ob, ref, k):
    """BatchedPSRSExo.run_evalmc (ols_run_exo) against the reference's evalMC_psrs over its PSRS_Exo: endogenous and exogenous
    queues popped in lock-step, the rejection generator re-seeded with the episode index at every reset (psrs.py:92,246)."""
    c, g, pi = make_exo_case(k)
    nS, nX, N, seed, R, E = c["nS"], c["nX"], c["N"], c["seed"], 3, 15
    want = reference_exo_run(ref, c, g, pi, E)
    common = dict(action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"], steps=g["t"])
    log_s = ob.BucketedLog(z=g["z"], z_next=g["z_next"], n_buckets=nS, **common).seeded(seed)
    log_x = ob.BucketedLog(z=g["x"], z_next=g["x_next"], n_buckets=nX, **common).seeded(seed)
    sim = ob.BatchedPSRSExo(log_s, log_x, R, seed=seed)
    res = sim.run_evalmc(nS * nX, E, pi, gamma=0.9, x_mult=nX, trace_cap=max(N, 1))
    n = want["trace"].size
    for r in range(R):
        assert int(res.accepted[r]) == n, c
        assert np.array_equal(res.trace_row[r, :n].cpu().numpy(), want["trace"]), c
        ne, nl = int(res.n_returns[r]), int(res.n_lengths[r])
        assert ne == want["Gs"].size and nl == want["lengths"].size, c
        assert np.array_equal(res.lengths[r, :nl].cpu().numpy(), want["lengths"]), c
        np.testing.assert_allclose(res.Gs[r, :ne].cpu().numpy(), want["Gs"], rtol=1e-5, atol=1e-5, err_msg=str(c))


# ------------------------------------------------------------------------------------------ the scalar API (psrs.py:5-57)
@pytest.mark.parametrize("k", range(20))
def test_scalar_psrs_walk_equals_the_live_reference(ob, ref, k):
    """`offsim4rl_b200.evaluators.psrs.PSRS` next to the reference's `PSRS` on the same legacy tuples: reset_sampler(seed),
    then reset / step with a fresh random distribution at every step until the initial states or a queue run out — every
    return value compared (agent-visible ids distinct from the state ids, psrs.py:36,49-50)."""
    from offsim4rl_b200.evaluators.psrs import PSRS
    ref_psrs, _ = ref
    rng = np.random.default_rng(11000 + k)
    pick = lambda xs: xs[int(rng.integers(0, len(xs)))]
    nA, nS = pick([2, 3, 5]), pick([1, 4, 11])
    N = max(pick([3, 40, 300]), nS)
    g = random_log(rng, N, nS, nA, pick([np.float32, np.float64]))
    seed = int(rng.integers(0, 2**31))

    def tuples():
        return [(int(g["z"][i]) * 7 + 3, g["a"][i], g["r"][i], int(g["z_next"][i]) * 7 + 3, bool(g["done"][i]), np.array(g["p_log"][i], copy=True),
                 {"z": int(g["z"][i]), "z_next": int(g["z_next"][i]), "t": int(g["t"][i])}) for i in range(N)]
    a, b = ref_psrs.PSRS(tuples(), nS=nS, nA=nA), PSRS(tuples(), nS=nS, nA=nA)
    a.reset_sampler(seed=seed)
    b.reset_sampler(seed=seed)
    sa, sb = a.reset(), b.reset()
    steps = 0
    while True:
        assert sa == sb
        if sa is None:
            break
        p = rng.dirichlet(np.ones(nA))
        oa, ob_ = a.step(p), b.step(p)
        steps += 1
        if oa[0] is None:
            assert ob_[0] is None and ob_ == (None, None, None, None)
            break
        assert oa[0] == ob_[0] and oa[1] == ob_[1] and bool(oa[2]) == bool(ob_[2])
        assert oa[3]["z"] == ob_[3]["z"] and oa[3]["a"] == ob_[3]["a"] and np.array_equal(oa[3]["p"], ob_[3]["p"])
        sa, sb = oa[0], ob_[0]
        if oa[2]:
            sa, sb = a.reset(), b.reset()
    assert steps > 0


# ------------------------------------------------------------------------------------------ QueueEvaluator (queue_evaluator.py:90-131)
@pytest.fixture(scope="module")
def ref_queue_impl():
    """The reference's QueueEvaluator_impl.  Its module imports gym and two package modules at the top for the wrapper class
    next to it; empty stand-ins satisfy those imports (the impl class itself uses numpy and itertools only)."""
    import sys
    import types
    path = os.path.join(REF_DIR, "evaluators", "queue_evaluator.py")
    if not os.path.exists(path):
        pytest.skip("baseline/_ref has no evaluators/queue_evaluator.py (re-run __graft_entry__.build() where /root/reference exists)")
    stubs = {"gym": types.ModuleType("gym"), "offsim4rl": types.ModuleType("offsim4rl"), "offsim4rl.core": types.ModuleType("offsim4rl.core"),
             "offsim4rl.data": types.ModuleType("offsim4rl.data")}
    stubs["gym"].spaces = types.SimpleNamespace(Discrete=type("Discrete", (), {}))
    stubs["offsim4rl.core"].RevealedRandomnessEnv = type("RevealedRandomnessEnv", (), {})
    stubs["offsim4rl.data"].OfflineDataset = type("OfflineDataset", (), {})
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("offsim4rl_live_queue_evaluator", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod.QueueEvaluator_impl


@pytest.mark.parametrize("k", range(20))
def test_queue_evaluator_equals_the_live_reference(ob, ref_queue_impl, k):
    """The reference's QueueEvaluator_impl, the package's scalar class and the fused BatchedPSRS.run_actions on the same random
    log, sampler seed and action sequence: every return value of the scalar walk, the served rows, lengths and returns of the
    fused run."""
    from offsim4rl_b200.evaluators.queue_evaluator import QueueEvaluator_impl
    from offsim4rl_b200 import _abi as abi
    rng = np.random.default_rng(13000 + k)
    pick = lambda xs: xs[int(rng.integers(0, len(xs)))]
    nA, nS = pick([2, 3, 4]), pick([1, 4, 9])
    N = max(pick([30, 200, 1200]), nS * nA)
    g = random_log(rng, N, nS, nA, np.float32)
    cover = rng.permutation(N)[:nS * nA]                     # every (z, a) pair has a queue: the reference raises KeyError otherwise
    g["z"][cover], g["a"][cover] = np.repeat(np.arange(nS), nA), np.tile(np.arange(nA), nS)
    seed = int(rng.integers(0, 2**31))
    acts = rng.integers(0, nA, 4 * N)

    def tuples():
        return [(int(g["z"][i]) * 5 + 1, int(g["a"][i]), g["r"][i], int(g["z_next"][i]) * 5 + 1, bool(g["done"][i]), np.array(g["p_log"][i], copy=True),
                 {"z": int(g["z"][i]), "next_z": int(g["z_next"][i]), "t": int(g["t"][i])}) for i in range(N)]
    rows = tuples()
    id2row = {id(r[5]): i for i, r in enumerate(rows)}
    a, b = ref_queue_impl(rows, nS=nS, nA=nA), QueueEvaluator_impl(tuples(), nS=nS, nA=nA)
    a.reset_sampler(seed=seed)
    b.reset_sampler(seed=seed)
    trace, lens, Gs, n_act = [], [], [], 0
    sa, sb = a.reset(), b.reset()
    G, t = 0.0, 0
    while True:
        assert sa == sb
        if sa is None:
            break
        act = int(acts[n_act]); n_act += 1
        oa, ob_ = a.step(act), b.step(act)
        if oa[0] is None:
            assert ob_ == (None, None, None, None)
            lens.append(t)                                   # (an episode cut short: evalMC-style bookkeeping keeps the length only)
            break
        assert oa[0] == ob_[0] and oa[1] == ob_[1] and bool(oa[2]) == bool(ob_[2])
        assert oa[3]["z"] == ob_[3]["z"] and oa[3]["next_z"] == ob_[3]["next_z"] and oa[3]["a"] == ob_[3]["a"]
        assert np.array_equal(oa[3]["p"], ob_[3]["p"])
        trace.append(id2row[id(oa[3]["p"])])
        G += 0.9 ** t * float(oa[1]); t += 1
        sa, sb = oa[0], ob_[0]
        if oa[2]:
            lens.append(t); Gs.append(G)
            G, t = 0.0, 0
            sa, sb = a.reset(), b.reset()
    # ---- the fused run: one launch for the whole walk
    log = ob.BucketedLog(z=g["z"], z_next=g["z_next"], action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"], steps=g["t"],
                         n_buckets=nS, key_by_action=True).seeded(seed)
    sim = ob.BatchedPSRS(log, 2, rng="philox", seed=1, reject="accept_all")
    res = sim.run_actions(N, gamma=0.9, actions=acts, trace_cap=N, ep_cap=N + 1)
    for r in range(2):
        n = int(res.accepted[r])
        assert n == len(trace) and np.array_equal(res.trace_row[r, :n].cpu().numpy(), np.array(trace, dtype=np.int64))
        assert int(sim.ties[r]) == n_act
        assert int(res.status[r]) in (abi.QUEUE_EMPTY, abi.INIT_EMPTY)
        ne, nl = int(res.n_returns[r]), int(res.n_lengths[r])
        assert ne == len(Gs) and nl == len(lens)
        assert np.array_equal(res.lengths[r, :nl].cpu().numpy(), np.array(lens))
        np.testing.assert_allclose(res.Gs[r, :ne].cpu().numpy(), np.array(Gs), rtol=1e-12, atol=1e-12)


# ------------------------------------------------------------------------------------------ CartpoleBoxEncoder (heuristic.py:10-71)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cartpole_box_encoder_equals_the_live_reference(ob, dtype):
    """ols_encode_cartpole_box against the reference's own `get_box` loop, in float32 and float64 observations: 20000 random
    rows over (and beyond) the encoder's range plus every threshold constant, its float32 image and their neighbours on either
    side in both precisions (comparisons run in the observation dtype, NumPy >= 2 semantics: SURVEY A.9)."""
    from offsim4rl_b200.encoders import CartpoleBoxEncoder
    path = os.path.join(REF_DIR, "encoders", "heuristic.py")
    if not os.path.exists(path):
        pytest.skip("baseline/_ref has no encoders/heuristic.py (re-run __graft_entry__.build() where /root/reference exists)")
    spec = importlib.util.spec_from_file_location("offsim4rl_live_heuristic", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(3)
    lim = np.array([2.6, 1.5, 0.23, 1.6])
    obs = rng.uniform(-lim, lim, (20000, 4))
    consts = [2.4, 0.8, 0.5, 0.0174532, 0.1047192, 0.2094384, 0.87266, 0.0]
    edge = []
    for c in consts:
        for sgn in (1.0, -1.0):
            v64 = np.float64(sgn * c)
            v32 = np.float64(np.float32(v64))
            for v in (v64, np.nextafter(v64, np.inf), np.nextafter(v64, -np.inf), v32,
                      np.float64(np.nextafter(np.float32(v64), np.float32(np.inf))), np.float64(np.nextafter(np.float32(v64), np.float32(-np.inf)))):
                edge.append(v)
    edge = np.array(edge)
    rows = rng.uniform(-lim * 0.3, lim * 0.3, (4 * edge.size, 4))
    for col in range(4):
        rows[col * edge.size:(col + 1) * edge.size, col] = edge
    obs = np.concatenate([obs, rows]).astype(dtype)
    want = mod.CartpoleBoxEncoder().encode(obs)
    got = CartpoleBoxEncoder().encode(obs)
    assert np.array_equal(got, want)
    assert (want == -1).sum() > 0 and np.unique(want).size > 100
