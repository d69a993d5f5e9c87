"""Generate the golden vectors in tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (it needs /root/reference):

    python tests/golden/make_golden.py

The reference (microsoft/rl-offline-simulation, package ``offsim4rl``) is pure Python.  ``gym``,
``h5py``, ``matplotlib`` and ``tensorboardX`` are absent here, so minimal in-memory stand-ins are
injected into ``sys.modules`` before import — none of them is on the PSRS path; the evaluators,
the tabular policies and the encoders run their own, unmodified code.  Harness techniques
(SURVEY.md Appendix C): accepted rows are recovered from ``id(info['p'])`` (every row gets its own
``p_log`` ndarray), uniforms are injected by replacing ``env.rejection_sampling_rng`` with a feed
object exposing ``.random()``, tie-breaks by patching ``np.random.choice`` while the reference's
``epsilon_greedy_policy`` runs.

Nothing at test time reads /root/reference: the fixtures written here are committed.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------- stand-ins
def _install_stubs():
    gym = types.ModuleType("gym")

    class Env:
        metadata = {}

        def seed(self, seed=None):
            return [seed]

    class Wrapper(Env):
        def __init__(self, env=None, new_step_api=False):
            self.env = env

    class Space:
        pass

    class Discrete(Space):
        def __init__(self, n):
            self.n = n
            self.shape = ()
            self.dtype = np.int64

        def contains(self, x):
            return 0 <= int(x) < self.n

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    spaces = types.ModuleType("gym.spaces")
    spaces.Discrete, spaces.Box, spaces.Space = Discrete, Box, Space
    gym.Env, gym.Wrapper, gym.Space, gym.spaces = Env, Wrapper, Space, spaces
    envs = types.ModuleType("gym.envs")
    reg = types.ModuleType("gym.envs.registration")
    reg.register = lambda **kw: None
    envs.registration = reg
    gym.envs = envs
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.envs": envs, "gym.envs.registration": reg})

    for name in ["h5py", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.cm",
                 "matplotlib.colors", "tensorboardX", "seaborn", "pygame"]:
        m = types.ModuleType(name)
        sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]
    sys.modules["matplotlib.patches"].Rectangle = object
    sys.modules["tensorboardX"].SummaryWriter = lambda *a, **k: types.SimpleNamespace(
        add_scalar=lambda *a, **k: None, add_histogram=lambda *a, **k: None, close=lambda: None)
    return gym


gym = _install_stubs()
sys.path.insert(0, REF)
from offsim4rl.evaluators import psrs as ref_psrs                     # noqa: E402
from offsim4rl.evaluators.per_state_rejection import PerStateRejectionSampling  # noqa: E402
from offsim4rl.evaluators import trivial_baselines as ref_baselines  # noqa: E402
from offsim4rl.agents import tabular as ref_tabular                  # noqa: E402
from offsim4rl.encoders.heuristic import CartpoleBoxEncoder          # noqa: E402
from offsim4rl.data import OfflineDataset, ProbDistribution          # noqa: E402


# --------------------------------------------------------------------------- harness
_HELPERS = None


def _helpers():
    """The repo's tests/helpers.py, loaded by FILE PATH: /root/reference is first on sys.path and has its own `tests`
    package, which would shadow `import tests.helpers`."""
    global _HELPERS
    if _HELPERS is None:
        import importlib.util
        spec = importlib.util.spec_from_file_location("_repo_test_helpers", os.path.join(os.path.dirname(OUT), "helpers.py"))
        _HELPERS = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_HELPERS)
    return _HELPERS


class F32Table(np.ndarray):
    """A Q_init whose `.copy().astype(float)` (psrs.py:141) stays float32: the reference's OWN update lines
    (psrs.py:165-168) then run on a float32 table, i.e. under NumPy >= 2 promotion as separately rounded float32
    operations `(R + gamma * max) - Q[S,A]`, `Q[S,A] + alpha * td` with gamma and alpha rounded to float32 — the
    arithmetic of the GPU production learner.  Nothing in the reference is modified."""

    def astype(self, dtype, *a, **k):
        return np.array(self, dtype=np.float32, copy=True, subok=False)


def f64_policy(policy):
    """The reference's policy evaluated on the float64 image of the (float32) Q row, so that p_new and the accept rule
    stay float64 as in the float64 runs (the arg-max is unchanged: the values are exact)."""
    return lambda Q, args: policy(np.asarray(Q, dtype=np.float64), args)


class Feed:
    def __init__(self, values):
        self.values = values
        self.k = 0

    def random(self):
        v = float(self.values[self.k])
        self.k += 1
        return v


class TieFeed:
    """Stands in for ``np.random.choice`` inside ``_random_argmax`` (agents/tabular.py:4-5)."""

    def __init__(self, values):
        self.values = values
        self.k = 0

    def __call__(self, candidates):
        n = len(candidates)
        if n == 1:
            return candidates[0]
        v = float(self.values[self.k])
        self.k += 1
        return candidates[min(n - 1, int(v * n))]


def make_log(rng, N, nS, nA, p_dtype=np.float32, r_dtype=np.float32, a_dtype=np.int64, ep_len=8,
             zero_prob_rows=0):
    z = rng.integers(0, nS, size=N)
    z_next = rng.integers(0, nS, size=N)
    pb_state = rng.dirichlet(np.ones(nA), size=nS)
    pb_state = np.maximum(pb_state, 0.05)
    pb_state /= pb_state.sum(1, keepdims=True)
    p_log = pb_state[z].astype(p_dtype)
    a = np.array([rng.choice(nA, p=pb_state[zi]) for zi in z]).astype(a_dtype)
    r = rng.normal(size=N).astype(r_dtype)
    done = rng.random(N) < 1.0 / ep_len
    t = np.zeros(N, dtype=np.int64)
    for i in range(1, N):
        t[i] = 0 if done[i - 1] else t[i - 1] + 1
    for k in range(zero_prob_rows):
        i = int(rng.integers(0, N))
        j = int(rng.integers(0, nA))
        p_log[i, j] = 0           # logged probability 0 -> inf / NaN ratios
    return dict(z=z.astype(np.int64), z_next=z_next.astype(np.int64), a=a, r=r, done=done, t=t, p_log=p_log)


def legacy_tuples(log, with_t=True):
    rows, id2row = [], {}
    for i in range(len(log["z"])):
        p = np.array(log["p_log"][i], copy=True)       # a distinct object per row
        id2row[id(p)] = i
        info = {"z": int(log["z"][i]), "z_next": int(log["z_next"][i])}
        if with_t:
            info["t"] = int(log["t"][i])
        else:
            info["t"] = 0
        rows.append((int(log["z"][i]), log["a"][i], log["r"][i], int(log["z_next"][i]), bool(log["done"][i]), p, info))
    return rows, id2row


def dump_queues(env, id2row, nS):
    offsets = np.zeros(nS + 1, dtype=np.int64)
    order = []
    for zz in range(nS):
        q = env.queues.get(zz, [])
        offsets[zz + 1] = offsets[zz] + len(q)
        order.extend(id2row[id(row[7])] for row in q)
    init = np.array([[zz, ss] for zz, ss in env.init_queue], dtype=np.int64).reshape(-1, 2)
    return offsets, np.array(order, dtype=np.int64), init


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path} ({os.path.getsize(path)} bytes)")


# --------------------------------------------------------------------------- cases
def case_seeded_walk():
    """reset_sampler(seed) queue/init order + a walk driven by the reference's OWN generator."""
    rng = np.random.default_rng(1234)
    N, nS, nA = 700, 9, 3
    log = make_log(rng, N, nS, nA)
    out = {k: v for k, v in log.items()}
    p_new = np.array([0.6, 0.3, 0.1])
    out["p_new"] = p_new
    seeds = [0, 3, 98052]
    out["seeds"] = np.array(seeds)
    for seed in seeds:
        rows, id2row = legacy_tuples(log)
        env = ref_psrs.PSRS(rows, nS=nS, nA=nA)
        env.reset_sampler(seed=seed)
        offsets, order, init = dump_queues(env, id2row, nS)
        out[f"offsets_{seed}"] = offsets
        out[f"order_{seed}"] = order
        out[f"init_zs_{seed}"] = init
        trace, resets = [], []
        s = env.reset()
        resets.append(s)
        while s is not None:
            s, r, done, info = env.step(p_new)
            if s is None:
                break
            trace.append(id2row[id(info["p"])])
            if done:
                s = env.reset()
                resets.append(-1 if s is None else s)
        out[f"trace_{seed}"] = np.array(trace, dtype=np.int64)
        out[f"resets_{seed}"] = np.array(resets, dtype=np.int64)
        out[f"pops_{seed}"] = np.array(N - sum(len(q) for q in env.queues.values()))
    save("seeded_walk", **out)


def _drivers_case(name, N, nS, nA, p_dtype, pi_dtype, zero_prob_rows=0, a_dtype=np.int64, seed=11, with_t=True,
                  only_mc=False):
    rng = np.random.default_rng(seed)
    log = make_log(rng, N, nS, nA, p_dtype=p_dtype, a_dtype=a_dtype, zero_prob_rows=zero_prob_rows)
    out = {k: v for k, v in log.items()}
    U = rng.random(8 * N)
    T = rng.random(8 * N)
    out["U"], out["T"] = U, T
    out["with_t"] = np.array(with_t)
    out["sampler_seed"] = np.array(5)

    def fresh():
        rows, id2row = legacy_tuples(log, with_t=with_t)
        env = ref_psrs.PSRS(rows, nS=nS, nA=nA)
        env.reset_sampler(seed=5)
        feed = Feed(U)
        env.rejection_sampling_rng = feed
        return env, id2row, feed

    # --- qlearn_psrs with the reference's epsilon_greedy_policy, constant and scheduled eps/alpha
    # (qlearn_psrs / expSARSA_psrs index Q[S, A] and so need an integer action column)
    for tag, eps, alpha in [] if only_mc else [("const", 0.2, 0.1),
                            ("sched", (lambda e: 1.0 / (1 + e)), (lambda e: 0.5 / (1 + 0.1 * e)))]:
        env, id2row, feed = fresh()
        ties = TieFeed(T)
        saved = np.random.choice
        np.random.choice = ties
        try:
            Q, info = ref_psrs.qlearn_psrs(env, n_episodes=25, behavior_policy=ref_tabular.epsilon_greedy_policy,
                                           gamma=0.9, alpha=alpha, epsilon=eps)
        finally:
            np.random.choice = saved
        out[f"ql_{tag}_Q"] = Q
        out[f"ql_{tag}_Gs"] = info["Gs"]
        out[f"ql_{tag}_TD"] = info["TD_errors"]
        out[f"ql_{tag}_trace"] = np.array([id2row[id(m[6]["p"])] for m in info["memory"]], dtype=np.int64)
        out[f"ql_{tag}_draws"] = np.array(feed.k)
        out[f"ql_{tag}_ties"] = np.array(ties.k)
    # --- the same two runs on a float32 Q table (F32Table): pins the float32 learner's operation order to the reference's code
    out32 = {}
    for tag, eps, alpha in [] if (only_mc or log["r"].dtype != np.float32) else [
            ("const", 0.2, 0.1), ("sched", (lambda e: 1.0 / (1 + e)), (lambda e: 0.5 / (1 + 0.1 * e)))]:
        env, id2row, feed = fresh()
        ties = TieFeed(T)
        saved = np.random.choice
        np.random.choice = ties
        try:
            Q, info = ref_psrs.qlearn_psrs(env, n_episodes=25, behavior_policy=f64_policy(ref_tabular.epsilon_greedy_policy),
                                           gamma=0.9, alpha=alpha, epsilon=eps, Q_init=np.zeros((nS, nA), np.float32).view(F32Table))
        finally:
            np.random.choice = saved
        assert Q.dtype == np.float32 and info["TD_errors"].dtype == np.float32
        out32[f"ql_{tag}_Q"] = Q
        out32[f"ql_{tag}_Gs"] = np.asarray(info["Gs"], dtype=np.float64)
        out32[f"ql_{tag}_TD"] = info["TD_errors"]
        out32[f"ql_{tag}_trace"] = np.array([id2row[id(m[6]["p"])] for m in info["memory"]], dtype=np.int64)
        out32[f"ql_{tag}_draws"] = np.array(feed.k)
        out32[f"ql_{tag}_ties"] = np.array(ties.k)
    if out32:
        save(name + "_f32q", **out32)
    out["ql_sched_eps"] = np.array([1.0 / (1 + e) for e in range(25)])
    out["ql_sched_alpha"] = np.array([0.5 / (1 + 0.1 * e) for e in range(25)])

    # --- expSARSA_psrs / evalMC_psrs with a fixed random policy
    pi = rng.dirichlet(np.ones(nA), size=nS).astype(pi_dtype)
    out["pi"] = pi
    if not only_mc:
        env, id2row, feed = fresh()
        Q, info = ref_psrs.expSARSA_psrs(env, n_episodes=25, pi=pi, gamma=0.95, alpha=0.2)
        out["es_Q"], out["es_Gs"] = Q, info["Gs"]
        out["es_trace"] = np.array([id2row[id(m[6]["p"])] for m in info["memory"]], dtype=np.int64)
        out["es_draws"] = np.array(feed.k)

    env, id2row, feed = fresh()
    # evalMC_psrs does not return the memory: trace the accepted rows through step()
    trace = []
    orig_step = env.step

    def traced_step(p):
        res = orig_step(p)
        if res[0] is not None:
            trace.append(id2row[id(res[3]["p"])])
        return res
    env.step = traced_step
    Gs, lengths = ref_psrs.evalMC_psrs(env, n_episodes=40, pi=pi, gamma=0.97)
    out["mc_Gs"], out["mc_lengths"] = Gs, lengths
    out["mc_trace"] = np.array(trace, dtype=np.int64)
    out["mc_draws"] = np.array(feed.k)
    save(name, **out)


def case_drivers():
    _drivers_case("drivers_f64", N=900, nS=10, nA=4, p_dtype=np.float32, pi_dtype=np.float64)
    # both distributions float32 -> the reference's accept rule runs in float32
    _drivers_case("drivers_f32", N=900, nS=10, nA=4, p_dtype=np.float32, pi_dtype=np.float32, seed=12)
    # float64 logged probabilities that are not float32-representable, 5 actions (gridworld shape)
    _drivers_case("drivers_p64_a5", N=800, nS=25, nA=5, p_dtype=np.float64, pi_dtype=np.float64, seed=13)
    # logged zero probabilities (inf / NaN thresholds), no steps column (every row is an initial state)
    _drivers_case("drivers_edge", N=600, nS=6, nA=2, p_dtype=np.float32, pi_dtype=np.float64,
                  zero_prob_rows=40, seed=14, with_t=False)
    # float action column (HDF5Dataset.concatenate forces float64, data.py:166): PSRS.step does int(a)
    _drivers_case("drivers_float_a", N=400, nS=5, nA=3, p_dtype=np.float32, pi_dtype=np.float64,
                  a_dtype=np.float64, seed=15, only_mc=True)


def case_accept_rule():
    """_default_reject on hand-made corner cases (thresholds at 0, 1, inf, NaN; u on the boundary)."""
    rows = []
    env = ref_psrs.PSRS([(0, 0, 0.0, 0, False, np.array([0.5, 0.5]), {"z": 0, "z_next": 0, "t": 0})], nS=1, nA=2)
    cases = []
    f32, f64 = np.float32, np.float64
    base = [
        ([0.5, 0.5], [0.5, 0.5], 0), ([0.9, 0.1], [0.5, 0.5], 1), ([0.9, 0.1], [0.5, 0.5], 0),
        ([0.3, 0.7], [0.0, 1.0], 1), ([0.3, 0.7], [0.0, 1.0], 0), ([0.0, 1.0], [0.0, 1.0], 0),
        ([0.0, 1.0], [0.0, 1.0], 1), ([1.0, 0.0], [0.25, 0.75], 1), ([0.2, 0.8], [0.7, 0.3], 0),
        ([1 / 3, 2 / 3], [0.1, 0.9], 0), ([1 / 3, 2 / 3], [0.1, 0.9], 1),
    ]
    us = [0.0, 1e-300, 0.1, 0.25, 0.5, 0.9999999999999999, float(np.float32(0.3)) + 1e-10,
          0.10714285714285714, 0.107142857142857, 0.03703703703703704, 0.037037037037037035]
    for pn, pl, a in base:
        for dn, dl in [(f64, f64), (f64, f32), (f32, f64), (f32, f32)]:
            for u in us:
                p_new, p_log = np.array(pn, dtype=dn), np.array(pl, dtype=dl)
                env.rejection_sampling_rng = Feed([u])
                with np.errstate(all="ignore"):
                    rej = bool(env._default_reject(p_new, p_log, a))
                rows.append((pn[0], pn[1], pl[0], pl[1], a, dn is f32, dl is f32, u, rej))
    arr = np.array(rows, dtype=np.float64)
    del cases
    save("accept_rule", table=arr)


def case_wrapper():
    """PerStateRejectionSampling + trivial baselines over an OfflineDataset (gym-style boundary)."""
    rng = np.random.default_rng(77)
    out = {}
    # (1) Discrete observations, no encoder, no 'steps' column (every row an initial state)
    N, nS, nA = 300, 8, 4
    log = make_log(rng, N, nS, nA)
    ds = OfflineDataset(
        observation_space=gym.spaces.Discrete(nS), action_space=gym.spaces.Discrete(nA),
        action_dist_type=ProbDistribution.Discrete,
        observations=log["z"], actions=log["a"], action_distributions=log["p_log"], rewards=log["r"],
        next_observations=log["z_next"], terminals=log["done"])
    for k in ["z", "a", "p_log", "r", "z_next", "done"]:
        out["disc_" + k] = log[k]
    U = rng.random(4 * N)
    out["disc_U"] = U
    target = np.array([0.4, 0.3, 0.2, 0.1])
    out["disc_target"] = target
    for api in (False, True):
        env = PerStateRejectionSampling(ds, new_step_api=api)
        env.reset_sampler(seed=2)
        env._impl.rejection_sampling_rng = Feed(U)
        obs = env.reset()
        acts, nobs, rews, dones = [], [], [], []
        first = obs
        while obs is not None:
            res = env.step_dist(target)
            if res[0] is None:
                assert len(res) == (6 if api else 5)
                break
            a, o, r, d = res[0], res[1], res[2], res[3]
            if api:
                assert res[4] is False
            acts.append(a); nobs.append(o); rews.append(r); dones.append(d)
            obs = o
            if d:
                obs = env.reset()
                nobs.append(-1 if obs is None else obs)
        tag = "disc_new" if api else "disc_old"
        out[tag + "_first"] = np.array(first)
        out[tag + "_a"] = np.array(acts)
        out[tag + "_obs"] = np.array(nobs)
        out[tag + "_r"] = np.array(rews)
        out[tag + "_d"] = np.array(dones)

    # (2) continuous CartPole-like observations + CartpoleBoxEncoder (ids include −1), with steps
    N = 500
    obs = np.stack([rng.uniform(-2.6, 2.6, N), rng.normal(0, 0.7, N), rng.uniform(-0.23, 0.23, N),
                    rng.normal(0, 1.0, N)], axis=1).astype(np.float32)
    nxt = np.roll(obs, -1, axis=0).copy()
    a = rng.integers(0, 2, N)
    pl = np.stack([np.full(N, 0.35), np.full(N, 0.65)], 1).astype(np.float32)
    r = np.ones(N, dtype=np.float32)
    done = rng.random(N) < 0.1
    steps = np.zeros(N, dtype=np.int64)
    for i in range(1, N):
        steps[i] = 0 if done[i - 1] else steps[i - 1] + 1
    ds = OfflineDataset(
        observation_space=gym.spaces.Box(low=-10, high=10, shape=(4,)), action_space=gym.spaces.Discrete(2),
        action_dist_type=ProbDistribution.Discrete,
        observations=obs, actions=a, action_distributions=pl, rewards=r, next_observations=nxt,
        terminals=done, steps=steps)
    out.update(cp_obs=obs, cp_next=nxt, cp_a=a, cp_p=pl, cp_r=r, cp_done=done, cp_steps=steps)
    U = rng.random(6 * N)
    out["cp_U"] = U
    target = np.array([0.5, 0.5])
    for cls, tag in [(PerStateRejectionSampling, "psrs"), (ref_baselines.FollowObservationOnly, "followobs"),
                     (ref_baselines.FollowActionOnly, "followact"), (ref_baselines.ServeRandomTransitions, "random")]:
        if cls in (PerStateRejectionSampling, ref_baselines.FollowObservationOnly):
            env = cls(ds, num_states=162, encoder=CartpoleBoxEncoder())
        else:
            env = cls(ds)
        env.reset_sampler(seed=9)
        env._impl.rejection_sampling_rng = Feed(U)
        o = env.reset()
        first = o
        acts, nobs, dones, status = [], [], [], "exhausted"
        try:
            while o is not None:
                res = env.step_dist(target)
                if res[0] is None:
                    break
                acts.append(res[0]); nobs.append(res[1]); dones.append(res[3])
                o = res[1]
                if res[3]:
                    o = env.reset()
                    if o is not None:
                        nobs.append(o)
        except KeyError:
            status = "keyerror"       # psrs.py:44 — next state never seen as a from-state
        out[f"cp_{tag}_first"] = np.array(first)
        out[f"cp_{tag}_a"] = np.array(acts)
        out[f"cp_{tag}_obs"] = np.array(nobs, dtype=np.float32).reshape(-1, 4)
        out[f"cp_{tag}_d"] = np.array(dones)
        out[f"cp_{tag}_status"] = np.array(status)
        out[f"cp_{tag}_draws"] = np.array(env._impl.rejection_sampling_rng.k)
    save("wrapper", **out)


def case_encoders():
    import torch
    from offsim4rl.encoders.homer import HOMEREncoder
    rng = np.random.default_rng(5)
    out = {}
    # CartpoleBoxEncoder, float32 and float64 observations, with values exactly on thresholds
    N = 4000
    obs = np.stack([rng.uniform(-2.6, 2.6, N), rng.normal(0, 0.7, N), rng.uniform(-0.23, 0.23, N),
                    rng.normal(0, 1.0, N)], axis=1)
    edges = [-2.4, 2.4, -0.8, 0.8, -0.5, 0.5, -0.1047192, -0.0174532, 0.0, 0.0174532, 0.1047192,
             -0.2094384, 0.2094384, -0.87266, 0.87266]
    k = 0
    for col in range(4):
        for e in edges:
            obs[k, col] = e
            k += 1
    obs32 = obs.astype(np.float32)
    obs64 = obs32.astype(np.float64)       # float32-valued thresholds seen as float64
    obs64[N // 2:] = obs[N // 2:]
    enc = CartpoleBoxEncoder()
    out["box_obs32"], out["box_ids32"] = obs32, enc.encode(obs32)
    out["box_obs64"], out["box_ids64"] = obs64, enc.encode(obs64)

    # HOMER obs_encoder inference: Linear(dO,H) -> LeakyReLU -> Linear(H,nZ) -> argmax(log_softmax)
    for tag, dO, H, nZ, n in [("mlp_grid", 2, 64, 25, 4000), ("mlp_cart", 4, 16, 10, 4000), ("mlp_wide", 2, 64, 50, 2000)]:
        torch.manual_seed(0)
        h = HOMEREncoder(obs_dim=dO, action_dim=5, latent_size=nZ, hidden_size=H)
        h.model.eval()
        x = rng.uniform(0, 1, size=(n, dO)).astype(np.float32)
        ids = h.encode(x)
        sd = h.model.obs_encoder.state_dict()
        out[tag + "_x"] = x
        out[tag + "_ids"] = ids
        out[tag + "_W1"] = sd["0.weight"].numpy().copy()
        out[tag + "_b1"] = sd["0.bias"].numpy().copy()
        out[tag + "_W2"] = sd["2.weight"].numpy().copy()
        out[tag + "_b2"] = sd["2.bias"].numpy().copy()
        with torch.no_grad():
            out[tag + "_logits"] = h.model.obs_encoder(torch.tensor(x)).numpy()
    save("encoders", **out)


def case_gridworld():
    """tests/test_psrs.py:9-34 scenario: the reference's own gridworld log and simulated-step count."""
    from offsim4rl.envs import gridworld
    env = gridworld.MyGridNavi(num_cells=5, num_steps=10, seed=0)
    env.reset_task(np.array([4, 4]))
    md = ref_tabular.rollout(env, 10, ref_tabular.epsilon_greedy_policy(np.zeros((env.nS, env.nA)), dict(epsilon=1)),
                             gamma=1, seed=0, show_tqdm=False)
    mem = md["memory"]
    out = dict(
        s=np.array([m[0] for m in mem]), a=np.array([m[1] for m in mem]), r=np.array([m[2] for m in mem]),
        s_next=np.array([m[3] for m in mem]), done=np.array([m[4] for m in mem]),
        p=np.array([m[5] for m in mem]), z=np.array([m[6]["z"] for m in mem]),
        z_next=np.array([m[6]["z_next"] for m in mem]), t=np.array([m[6]["t"] for m in mem]))
    counts = []
    for seed in range(20):
        e = ref_psrs.PSRS(mem, nS=25, nA=5)
        e.reset_sampler(seed=seed)
        obs = e.reset()
        n = -1
        while obs is not None:
            obs, r, done, info = e.step(np.array([0.2] * 5))
            n += 1
            if done:
                obs = e.reset()
        counts.append(n)
    out["steps_simulated_by_seed"] = np.array(counts)
    save("gridworld", **out)


def case_queue_evaluator():
    """QueueEvaluator_impl (queues keyed by (z, a), no rejection) driven by a fixed action sequence."""
    from offsim4rl.evaluators.queue_evaluator import QueueEvaluator_impl
    rng = np.random.default_rng(31)
    N, nS, nA = 500, 6, 3
    log = make_log(rng, N, nS, nA)
    out = {k: v for k, v in log.items()}
    acts = rng.integers(0, nA, size=4 * N)
    out["acts"] = acts
    for seed in (0, 4):
        rows, id2row = legacy_tuples(log)
        rows = [(s, a, r, s_, d, p, {**info, "next_z": info["z_next"]}) for s, a, r, s_, d, p, info in rows]
        env = QueueEvaluator_impl(rows, nS=nS, nA=nA)
        env.reset_sampler(seed=seed)
        order, offsets = [], [0]
        for z in range(nS):
            for a in range(nA):
                qq = env.queues.get((z, a), [])
                order.extend(id2row[id(e[7])] for e in qq)
                offsets.append(offsets[-1] + len(qq))
        out[f"order_{seed}"], out[f"offsets_{seed}"] = np.array(order), np.array(offsets)
        trace, resets, k, status = [], [], 0, "exhausted"
        s = env.reset()
        resets.append(s)
        try:
            while s is not None:
                s, r, done, info = env.step(acts[k]); k += 1
                if s is None:
                    break
                trace.append(id2row[id(info["p"])])
                if done:
                    s = env.reset()
                    resets.append(-1 if s is None else s)
        except KeyError:
            status = "keyerror"
        out[f"trace_{seed}"], out[f"resets_{seed}"] = np.array(trace, dtype=np.int64), np.array(resets, dtype=np.int64)
        out[f"status_{seed}"], out[f"nacts_{seed}"] = np.array(status), np.array(k)
    save("queue_evaluator", **out)


def case_psrs_exo():
    """PSRS_Exo (psrs.py:59-117): o = 3*s + x, split/combine by divmod; the reference's own generator, re-seeded by
    reset(seed=episode)."""
    rng = np.random.default_rng(41)
    N, nS, nX, nA = 600, 5, 3, 3
    log = make_log(rng, N, nS, nA, ep_len=6)
    x = rng.integers(0, nX, N)
    x_next = rng.integers(0, nX, N)
    o, o_next = log["z"] * nX + x, log["z_next"] * nX + x_next
    out = {k: v for k, v in log.items()}
    out.update(x=x, x_next=x_next)
    p_new = np.array([0.5, 0.2, 0.3])
    out["p_new"] = p_new
    rows, id2row = [], {}
    for i in range(N):
        p = np.array(log["p_log"][i], copy=True)
        id2row[id(p)] = i
        rows.append((int(o[i]), int(log["a"][i]), log["r"][i], int(o_next[i]), bool(log["done"][i]), p, {"t": int(log["t"][i])}))
    env = ref_psrs.PSRS_Exo(rows, nO=nS * nX, nA=nA, o_split_func=lambda v: divmod(v, nX), o_combine_func=lambda s, xx: s * nX + xx)
    # the reference deep-copies the buffer (psrs.py:65): recover rows through the copied p_log objects
    id2row = {id(e[5]): i for i, e in enumerate(env.raw_buffer)}
    for seed in (1, 6):
        env.reset_sampler(seed=seed)
        trace, obs, episode, status = [], [], 0, "exhausted"
        ob = env.reset(seed=episode)
        obs.append(-1 if ob is None else ob)
        try:
            while ob is not None:
                ob, r, done, info = env.step(p_new)
                if ob is None:
                    break
                trace.append(id2row[id(info["p"])])
                obs.append(ob)
                if done:
                    episode += 1
                    ob = env.reset(seed=episode)
                    obs.append(-1 if ob is None else ob)
        except KeyError:
            status = "keyerror"
        out[f"trace_{seed}"], out[f"obs_{seed}"] = np.array(trace, dtype=np.int64), np.array(obs, dtype=np.int64)
        out[f"status_{seed}"] = np.array(status)
    save("psrs_exo", **out)


def case_recorder():
    """utils/dataset_utils.py:34-136 `record_dataset_in_memory` on the toy environment of tests/helpers.py: the column
    layout (`infos/<key>` flattening, first-step key rule, episode ids with several workers) that the ingest consumes."""
    ToyChainEnv, ToyAgent = _helpers().ToyChainEnv, _helpers().ToyAgent
    from offsim4rl.utils.dataset_utils import record_dataset_in_memory
    out = {}
    for tag, kw in {"w0": dict(seed=5, worker_id=0, workers_num=1), "w2of3": dict(seed=8, worker_id=2, workers_num=3)}.items():
        np.random.seed(123)
        agent = ToyAgent(ProbDistribution.Discrete)
        ds = record_dataset_in_memory(ToyChainEnv(), agent, 300, new_step_api=True, include_infos=True, **kw)
        for k, v in ds.experience.items():
            out[f"{tag}/{k}".replace("/", "__")] = np.asarray(v)
        out[f"{tag}__committed"] = np.array(agent.committed)
        out[f"{tag}__ended"] = np.array(agent.ended, dtype=np.float64)
    save("recorder", **out)


def case_back_to_back():
    """Several drivers on ONE env without reset_sampler in between (learn, then evaluate, then learn on): every call counts
    its episodes from 0 and keeps consuming the queues and the uniform stream where the previous call stopped."""
    rng = np.random.default_rng(808)
    N, nS, nA = 6000, 8, 3
    log = make_log(rng, N, nS, nA, ep_len=6)
    out = {k: v for k, v in log.items()}
    U, T = rng.random(2 * N), rng.random(N // 4)
    pi = rng.dirichlet(np.ones(nA), size=nS)
    out.update(U=U, T=T, pi=pi, sampler_seed=np.array(7))
    rows, id2row = legacy_tuples(log)
    env = ref_psrs.PSRS(rows, nS=nS, nA=nA)
    env.reset_sampler(seed=7)
    feed = Feed(U)
    env.rejection_sampling_rng = feed
    trace = []
    orig_step = env.step

    def traced_step(p):
        res = orig_step(p)
        if res[0] is not None:
            trace.append(id2row[id(res[3]["p"])])
        return res
    env.step = traced_step
    ties = TieFeed(T)
    saved = np.random.choice
    np.random.choice = ties
    try:
        eps_list, alpha_list = [0.5 / (1 + e) for e in range(10)], [0.3] * 10
        Q1, i1 = ref_psrs.qlearn_psrs(env, n_episodes=10, behavior_policy=ref_tabular.epsilon_greedy_policy, gamma=0.9,
                                      alpha=lambda e: alpha_list[e], epsilon=lambda e: eps_list[e])   # defined on [0, 10) only
        k1, d1, t1 = len(trace), feed.k, ties.k
        Gs2, len2 = ref_psrs.evalMC_psrs(env, n_episodes=12, pi=pi, gamma=0.95)
        k2, d2 = len(trace), feed.k
        s_mid = env.s
        # a scalar step in between, then learning resumes from the learnt table
        s3, r3, done3, info3 = env.step(pi[env.s])
        k3, d3 = len(trace), feed.k
        Q4, i4 = ref_psrs.qlearn_psrs(env, n_episodes=6, behavior_policy=ref_tabular.epsilon_greedy_policy, gamma=0.9,
                                      alpha=0.2, epsilon=0.1, Q_init=Q1)
        k4, d4, t4 = len(trace), feed.k, ties.k
        Q5, i5 = ref_psrs.expSARSA_psrs(env, n_episodes=5, pi=pi, gamma=0.9, alpha=0.2)
        k5, d5 = len(trace), feed.k
    finally:
        np.random.choice = saved
    out.update(trace=np.array(trace, dtype=np.int64), marks=np.array([k1, k2, k3, k4, k5]), draws=np.array([d1, d2, d3, d4, d5]),
               ties=np.array([t1, t4]), eps_list=np.array(eps_list), alpha_list=np.array(alpha_list),
               Q1=Q1, Gs1=i1["Gs"], TD1=i1["TD_errors"], Gs2=Gs2, len2=len2, s_mid=np.array(s_mid), s3=np.array(s3),
               Q4=Q4, Gs4=i4["Gs"], Q5=Q5, Gs5=i5["Gs"])
    save("back_to_back", **out)


def case_learner_curves():
    """north_star's third bar: learning curves of the LEARNER under each side's own RNG.  The unmodified `qlearn_psrs`
    (psrs.py:119-185) with the reference's epsilon-greedy policy runs 320 times; run i re-seeds the reference's own
    generators (`reset_sampler(seed=5000+i)` -> rejection uniforms from NumPy's PCG64; `np.random.seed(i)` -> MT19937
    tie-breaks) and follows the data order of GPU replica 10**7 + i: log order in every queue, consumed as a ring from the
    documented per-(replica, state) offset (`env.queues` / `env.init_queue` are public lists)."""
    H = _helpers()
    rng = np.random.default_rng(2024)
    N, nS, nA, E, RUNS, RING_SEED = 20000, 6, 4, 12, 320, 11
    log = make_log(rng, N, nS, nA, ep_len=12)
    # rewards that depend on (state, action), so that what the learner has learnt changes the returns it collects
    r_table = rng.normal(size=(nS, nA))
    log["r"] = (r_table[log["z"], log["a"]] + 0.3 * rng.normal(size=N)).astype(np.float32)
    pb_rows = {int(z): log["p_log"][i] for i, z in enumerate(log["z"])}
    out = dict(z=log["z"].astype(np.uint8), z_next=log["z_next"].astype(np.uint8), a=log["a"].astype(np.uint8), r=log["r"],
               done=log["done"], t=log["t"].astype(np.int32), pb_state=np.stack([pb_rows[z] for z in range(nS)]),
               ring_seed=np.array(RING_SEED), replica_id0=np.array(10**7), n_episodes=np.array(E),
               gamma=np.array(0.95), alpha=np.array(0.2), epsilon=np.array(0.3))
    assert np.array_equal(out["pb_state"][log["z"]], log["p_log"])
    rows, id2row = legacy_tuples(log)
    lens = np.bincount(log["z"], minlength=nS)
    n_init = int((log["t"] == 0).sum())
    Gs = np.full((RUNS, E), np.nan)
    n_ret = np.zeros(RUNS, dtype=np.int64)
    acc = np.zeros(RUNS, dtype=np.int64)
    for i in range(RUNS):
        rid = 10**7 + i
        env = ref_psrs.PSRS(rows, nS=nS, nA=nA)
        env.reset_sampler(seed=5000 + i)
        for zz, q in env.queues.items():
            q.sort(key=lambda e: id2row[id(e[7])])                      # back to log order ...
            o = H.ring_start(RING_SEED, rid, zz, int(lens[zz]))
            q[:] = q[o:] + q[:o]                                         # ... consumed as a ring from the replica's offset
        first = [(int(log["z"][j]), int(log["z"][j])) for j in np.nonzero(log["t"] == 0)[0]]
        o = H.init_ring_start(RING_SEED, rid, n_init)
        env.init_queue[:] = first[o:] + first[:o]
        np.random.seed(i)
        Q, info = ref_psrs.qlearn_psrs(env, n_episodes=E, behavior_policy=ref_tabular.epsilon_greedy_policy, gamma=0.95,
                                       alpha=0.2, epsilon=0.3)
        g = np.asarray(info["Gs"], dtype=np.float64)
        Gs[i, :g.size] = g
        n_ret[i] = g.size
        acc[i] = len(info["memory"])
    out.update(ref_Gs=Gs, ref_n_returns=n_ret, ref_accepted=acc)
    save("learner_curves", **out)


if __name__ == "__main__":
    only = sys.argv[1:]                   # e.g. `make_golden.py learner_curves` regenerates one fixture
    if only:
        for name in only:
            globals()["case_" + name]()
        sys.exit(0)
    case_learner_curves()
    case_recorder()
    case_seeded_walk()
    case_drivers()
    case_accept_rule()
    case_wrapper()
    case_encoders()
    case_gridworld()
    case_queue_evaluator()
    case_psrs_exo()
