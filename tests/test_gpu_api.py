"""GPU tests of the reference-facing API: PSRS / PerStateRejectionSampling / drivers, against the golden
vectors of the unmodified reference and the reference's own behavioural tests."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import offsim4rl_b200
    return offsim4rl_b200


def _legacy(g, with_t=True, keys=("z", "z_next")):
    rows = []
    for i in range(len(g["z"])):
        info = {keys[0]: int(g["z"][i]), keys[1]: int(g["z_next"][i]), "t": int(g["t"][i]) if with_t else 0}
        rows.append((int(g["z"][i]), g["a"][i], g["r"][i], int(g["z_next"][i]), bool(g["done"][i]), g["p_log"][i], info))
    return rows


# ----------------------------------------------------------------------------- PSRS scalar API
@pytest.mark.parametrize("seed", [0, 3, 98052])
def test_psrs_seeded_walk_equals_reference_run(ob, golden, seed):
    """PSRS(buffer) + reset_sampler(seed) + step(p): same queues, same RNG, same accepted rows as the reference."""
    from offsim4rl_b200.evaluators.psrs import PSRS
    g = golden("seeded_walk")
    env = PSRS(_legacy(g, keys=("z", "next_z") if seed == 3 else ("z", "z_next")), nS=9, nA=3)
    env.reset_sampler(seed=seed)
    order = np.concatenate([env.queues.get(z, []) for z in range(9)]).astype(np.int64)
    assert np.array_equal(order, g[f"order_{seed}"])
    assert np.array_equal(np.array([z for z, s in env.init_queue]), g[f"init_zs_{seed}"][:, 0])
    rewards, resets, n = [], [], 0
    s = env.reset()
    resets.append(s)
    expect = g[f"trace_{seed}"]
    while s is not None:
        s, r, done, info = env.step(g["p_new"])
        if s is None:
            assert (r, done, info) == (None, None, None)
            break
        row = int(expect[n]); n += 1
        assert s == g["z_next"][row] and r == g["r"][row] and done == g["done"][row]
        assert info["z"] == g["z"][row] and info["a"] == g["a"][row] and info["p"] is not None
        assert np.array_equal(info["p"], g["p_log"][row])
        if done:
            s = env.reset()
            resets.append(-1 if s is None else s)
    assert n == expect.size
    assert np.array_equal(np.array(resets, dtype=np.int64), g[f"resets_{seed}"])
    assert env.pops == int(g[f"pops_{seed}"])


def test_reference_test_psrs_scenario(ob, golden):
    """tests/test_psrs.py:9-34 of the reference on its own gridworld log: > 40 simulated steps, and per-seed
    counts identical to the reference's."""
    from offsim4rl_b200.evaluators import psrs
    g = golden("gridworld")
    mem = [(int(g["s"][i]), int(g["a"][i]), float(g["r"][i]), int(g["s_next"][i]), bool(g["done"][i]), g["p"][i],
            {"z": int(g["z"][i]), "z_next": int(g["z_next"][i]), "t": int(g["t"][i])}) for i in range(len(g["z"]))]
    env_psrs = psrs.PSRS(mem, nS=25, nA=5)
    counts = []
    for seed in [None] + list(range(20)):
        if seed is not None:
            env_psrs.reset_sampler(seed=seed)
        obs = env_psrs.reset()
        steps_simulated = -1
        while obs is not None:
            obs, r, done, info = env_psrs.step(np.array([0.2, 0.2, 0.2, 0.2, 0.2]))
            steps_simulated += 1
            if done:
                obs = env_psrs.reset()
        assert steps_simulated > 40
        if seed is not None:
            counts.append(steps_simulated)
    assert np.array_equal(counts, g["steps_simulated_by_seed"])


@pytest.mark.parametrize("case", ["drivers_f64", "drivers_p64_a5", "drivers_edge"])
def test_drivers_with_reference_signatures(ob, golden, case):
    from offsim4rl_b200.evaluators.psrs import PSRS, qlearn_psrs, expSARSA_psrs, evalMC_psrs
    from offsim4rl_b200.agents import tabular
    g = golden(case)
    nS, nA = g["pi"].shape
    with_t = bool(g["with_t"])

    def fresh():
        env = PSRS(_legacy(g, with_t), nS=nS, nA=nA)
        env.reset_sampler(seed=int(g["sampler_seed"]))
        env.feed_uniforms(g["U"])
        return env

    # expSARSA / evalMC have no tie-breaks: identical to the reference's outputs
    Q, info = expSARSA_psrs(fresh(), n_episodes=25, pi=g["pi"], gamma=0.95, alpha=0.2)
    np.testing.assert_allclose(Q, g["es_Q"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(info["Gs"], g["es_Gs"], rtol=1e-5, atol=1e-5)
    assert len(info["memory"]) == g["es_trace"].size
    assert [m[6]["a"] for m in info["memory"]] == [g["a"][r] for r in g["es_trace"]]
    Gs, lengths = evalMC_psrs(fresh(), n_episodes=40, pi=g["pi"], gamma=0.97)
    assert np.array_equal(lengths, g["mc_lengths"])
    np.testing.assert_allclose(Gs, g["mc_Gs"], rtol=1e-5, atol=1e-5)
    # qlearn: the tie stream is the library's own here, so compare structure and the generic host loop
    env = fresh()
    Q, info = qlearn_psrs(env, n_episodes=25, behavior_policy=tabular.epsilon_greedy_policy, gamma=0.9, alpha=0.1, epsilon=0.2)
    assert Q.shape == (nS, nA) and info["Gs"].ndim == 1 and info["TD_errors"].size == len(info["memory"])
    assert info["Qs"].shape == (1, nS, nA)
    # a custom policy callable goes through the host loop (every step still on the GPU)
    np.random.seed(0)
    env = fresh()
    Q2, info2 = qlearn_psrs(env, n_episodes=5, behavior_policy=lambda Qr, a: tabular.epsilon_greedy_policy(Qr, a), gamma=0.9,
                            alpha=0.1, epsilon=0.2)
    assert Q2.shape == (nS, nA) and len(info2["Gs"]) <= 5 and info2["TD_errors"].size == len(info2["memory"])


def test_back_to_back_drivers_on_one_env_match_reference(ob, golden):
    """Learn, evaluate, take a scalar step, learn on from the learnt table, then expected SARSA — all on ONE env without
    reset_sampler, as the unmodified reference did when tests/golden/back_to_back.npz was recorded: every driver call counts
    its episodes from 0 and keeps consuming the queues and the uniform stream where the previous call stopped."""
    from offsim4rl_b200.evaluators.psrs import PSRS, qlearn_psrs, expSARSA_psrs, evalMC_psrs
    from offsim4rl_b200.agents import tabular
    g = golden("back_to_back")
    nS, nA = g["pi"].shape
    env = PSRS(_legacy(g), nS=nS, nA=nA)
    env.reset_sampler(seed=int(g["sampler_seed"]))
    env.feed_uniforms(g["U"])
    env.feed_ties(g["T"])
    k1, k2, k3, k4, k5 = (int(v) for v in g["marks"])
    eps_list, alpha_list = list(g["eps_list"]), list(g["alpha_list"])

    def same_rows(memory, lo, hi):
        rows = g["trace"][lo:hi]
        assert len(memory) == hi - lo
        assert [int(m[1]) for m in memory] == [int(g["a"][r]) for r in rows]
        assert [float(m[2]) for m in memory] == [float(g["r"][r]) for r in rows]

    # schedules defined on [0, n_episodes) only, like the reference's lazy calls (psrs.py:158,168)
    Q1, i1 = qlearn_psrs(env, n_episodes=10, behavior_policy=tabular.epsilon_greedy_policy, gamma=0.9,
                         alpha=lambda e: alpha_list[e], epsilon=lambda e: eps_list[e])
    same_rows(i1["memory"], 0, k1)
    assert np.array_equal(Q1, g["Q1"]) and np.array_equal(i1["TD_errors"], g["TD1"])
    np.testing.assert_allclose(i1["Gs"], g["Gs1"], rtol=1e-5, atol=1e-5)
    assert int(env._sim.draws[0]) == int(g["draws"][0]) and int(env._sim.ties[0]) == int(g["ties"][0])
    Gs2, len2 = evalMC_psrs(env, n_episodes=12, pi=g["pi"], gamma=0.95)
    assert np.array_equal(len2, g["len2"]) and Gs2.size == g["Gs2"].size
    np.testing.assert_allclose(Gs2, g["Gs2"], rtol=1e-5, atol=1e-5)
    assert int(env._sim.draws[0]) == int(g["draws"][1]) and int(env.s) == int(g["s_mid"])
    s3, r3, d3, info3 = env.step(g["pi"][env.s])                              # the env is an ordinary env in between
    assert int(s3) == int(g["s3"]) and int(env._sim.draws[0]) == int(g["draws"][2])
    assert float(r3) == float(g["r"][g["trace"][k2]])
    Q4, i4 = qlearn_psrs(env, n_episodes=6, behavior_policy=tabular.epsilon_greedy_policy, gamma=0.9, alpha=0.2, epsilon=0.1,
                         Q_init=Q1)
    same_rows(i4["memory"], k3, k4)
    assert np.array_equal(Q4, g["Q4"]) and i4["Gs"].size == g["Gs4"].size
    assert int(env._sim.draws[0]) == int(g["draws"][3]) and int(env._sim.ties[0]) == int(g["ties"][1])
    Q5, i5 = expSARSA_psrs(env, n_episodes=5, pi=g["pi"], gamma=0.9, alpha=0.2)
    same_rows(i5["memory"], k4, k5)
    np.testing.assert_allclose(Q5, g["Q5"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(i5["Gs"], g["Gs5"], rtol=1e-5, atol=1e-5)
    assert int(env._sim.draws[0]) == int(g["draws"][4])


def test_generic_host_loop_equals_fused_kernel(ob, golden):
    """The host-driven loop (arbitrary policy callables) and the fused kernel implement the same algorithm."""
    from offsim4rl_b200.evaluators.psrs import PSRS, expSARSA_psrs, evalMC_psrs
    g = golden("drivers_f64")
    nS, nA = g["pi"].shape

    class Unfusable(PSRS):
        pass

    def fresh(cls):
        env = cls(_legacy(g), nS=nS, nA=nA)
        env.reset_sampler(seed=5)
        env.feed_uniforms(g["U"])
        return env
    import offsim4rl_b200.evaluators.psrs as mod
    saved = mod._fusable
    try:
        mod._fusable = lambda env: False
        Qh, ih = expSARSA_psrs(fresh(PSRS), 25, g["pi"], 0.95, alpha=0.2)
        Gh, Lh = evalMC_psrs(fresh(PSRS), 40, g["pi"], 0.97)
    finally:
        mod._fusable = saved
    np.testing.assert_allclose(Qh, g["es_Q"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(ih["Gs"], g["es_Gs"], rtol=1e-5, atol=1e-5)
    assert np.array_equal(Lh, g["mc_lengths"])
    np.testing.assert_allclose(Gh, g["mc_Gs"], rtol=1e-5, atol=1e-5)


def test_unseen_state_raises_keyerror_like_the_reference(ob):
    from offsim4rl_b200.evaluators.psrs import PSRS
    p = np.array([0.5, 0.5])
    rows = [(0, 0, 1.0, 2, False, p, {"z": 0, "z_next": 2, "t": 0})]       # state 2 is never a from-state
    env = PSRS(rows, nS=3, nA=2)
    env.reset_sampler(seed=0)
    assert env.reset() == 0
    s, r, done, info = env.step(p)
    assert s == 2
    with pytest.raises(KeyError):
        env.step(p)
    assert env.reset() is None                                              # psrs.py:33-35
    with pytest.raises(ValueError):
        PSRS(rows, nS=3, nA=2, reject_func="no such rule")


def test_python_reject_callbacks_run_through_the_host_loop(ob, golden):
    """psrs.py:11,47 `reject_func` and per_state_rejection.py:97-98 `_reject` are hooks for arbitrary Python: they are served
    by a host-driven loop (one single-record pop on the GPU per call).  With the reference's own rule written as a callable
    over the fed uniform stream, the run equals the golden run of the unmodified reference (which decided on the device
    path's behalf in every other test)."""
    from offsim4rl_b200.evaluators.psrs import PSRS, evalMC_psrs, qlearn_psrs
    from offsim4rl_b200.agents import tabular
    g = golden("drivers_f64")
    nS, nA = g["pi"].shape
    U, k = g["U"], [0]

    def reject(p_new, p_log, a):                       # psrs.py:53-57 on the harness's uniforms
        a = int(a)
        M = (p_new / p_log).max()
        u = U[k[0]]
        k[0] += 1
        return u > p_new[a] / p_log[a] / M

    env = PSRS(_legacy(g), nS=nS, nA=nA, reject_func=reject)
    env.reset_sampler(seed=int(g["sampler_seed"]))
    assert env.rejection_sampling_rng is not None
    Gs, lengths = evalMC_psrs(env, n_episodes=40, pi=g["pi"], gamma=0.97)
    assert np.array_equal(lengths, g["mc_lengths"]) and k[0] == int(g["mc_draws"])
    np.testing.assert_allclose(Gs, g["mc_Gs"], rtol=1e-5, atol=1e-5)
    assert int(env._sim.accepted[0]) == g["mc_trace"].size and env.pops == int(g["mc_draws"])
    # the gym-style wrapper with an overridden _reject hook
    from offsim4rl_b200.evaluators import PerStateRejectionSampling
    from offsim4rl_b200.data import OfflineDataset, ProbDistribution
    from offsim4rl_b200 import spaces
    w = golden("wrapper")
    calls = [0]

    class AcceptEveryOther(PerStateRejectionSampling):
        def _reject(self, p_new, p_log, a):
            calls[0] += 1
            return calls[0] % 2 == 0

    ds = OfflineDataset(observation_space=spaces.Discrete(8), action_space=spaces.Discrete(4), action_dist_type=ProbDistribution.Discrete,
                        observations=w["disc_z"], actions=w["disc_a"], action_distributions=w["disc_p_log"], rewards=w["disc_r"],
                        next_observations=w["disc_z_next"], terminals=w["disc_done"])
    envw = AcceptEveryOther(ds)
    envw.reset_sampler(seed=2)
    obs, served = envw.reset(), 0
    while obs is not None and served < 20:
        res = envw.step_dist(w["disc_target"])
        if res[0] is None:
            break
        served += 1
        obs = envw.reset() if res[3] else res[1]
    assert served > 0 and calls[0] >= served and envw._impl.pops == calls[0]


# ----------------------------------------------------------------------------- gym-style wrapper
def _disc_dataset(ob, g):
    from offsim4rl_b200.data import OfflineDataset, ProbDistribution
    from offsim4rl_b200 import spaces
    return OfflineDataset(observation_space=spaces.Discrete(8), action_space=spaces.Discrete(4),
                          action_dist_type=ProbDistribution.Discrete, observations=g["disc_z"], actions=g["disc_a"],
                          action_distributions=g["disc_p_log"], rewards=g["disc_r"], next_observations=g["disc_z_next"],
                          terminals=g["disc_done"])


@pytest.mark.parametrize("api", [False, True])
def test_wrapper_discrete_matches_reference(ob, golden, api):
    from offsim4rl_b200.evaluators import PerStateRejectionSampling
    g = golden("wrapper")
    env = PerStateRejectionSampling(_disc_dataset(ob, g), new_step_api=api)
    env.reset_sampler(seed=2)
    env._impl.feed_uniforms(g["disc_U"])
    tag = "disc_new" if api else "disc_old"
    obs = env.reset()
    assert obs == g[tag + "_first"] and np.shape(obs) == ()
    acts, nobs, rews, dones = [], [], [], []
    while obs is not None:
        res = env.step_dist(g["disc_target"])
        if res[0] is None:
            assert res == (None,) * (6 if api else 5)
            break
        if api:
            assert res[4] is False and isinstance(res[5], dict)
        acts.append(res[0]); nobs.append(res[1]); rews.append(res[2]); dones.append(res[3])
        obs = res[1]
        if res[3]:
            obs = env.reset()
            nobs.append(-1 if obs is None else obs)
    assert np.array_equal(acts, g[tag + "_a"]) and np.array_equal(nobs, g[tag + "_obs"])
    assert np.array_equal(rews, g[tag + "_r"]) and np.array_equal(dones, g[tag + "_d"])


@pytest.mark.parametrize("tag", ["psrs", "followobs", "followact", "random"])
@pytest.mark.parametrize("device_encoder", [True, False])
def test_wrapper_cartpole_encoder_and_baselines_match_reference(ob, golden, tag, device_encoder):
    from offsim4rl_b200.data import OfflineDataset, ProbDistribution
    from offsim4rl_b200 import spaces
    from offsim4rl_b200.encoders import CartpoleBoxEncoder
    from offsim4rl_b200.evaluators import (PerStateRejectionSampling, FollowObservationOnly, FollowActionOnly,
                                           ServeRandomTransitions)
    g = golden("wrapper")
    ds = OfflineDataset(observation_space=spaces.Box(low=-10, high=10, shape=(4,)), action_space=spaces.Discrete(2),
                        action_dist_type=ProbDistribution.Discrete, observations=g["cp_obs"], actions=g["cp_a"],
                        action_distributions=g["cp_p"], rewards=g["cp_r"], next_observations=g["cp_next"],
                        terminals=g["cp_done"], steps=g["cp_steps"])
    base = CartpoleBoxEncoder()
    enc = base
    if not device_encoder:
        class HostEncoder:                      # a user-supplied encoder with only .encode (numpy in, ids out)
            def encode(self, o):
                return base.encode(o)
        enc = HostEncoder()
    cls = {"psrs": PerStateRejectionSampling, "followobs": FollowObservationOnly, "followact": FollowActionOnly,
           "random": ServeRandomTransitions}[tag]
    env = cls(ds, num_states=162, encoder=enc) if tag in ("psrs", "followobs") else cls(ds)
    env.reset_sampler(seed=9)
    env._impl.feed_uniforms(g["cp_U"])
    o = env.reset()
    assert np.array_equal(o, g[f"cp_{tag}_first"])
    acts, nobs, dones, status = [], [], [], "exhausted"
    try:
        while o is not None:
            res = env.step_dist(np.array([0.5, 0.5]))
            if res[0] is None:
                break
            acts.append(res[0]); nobs.append(res[1]); dones.append(res[3])
            o = res[1]
            if res[3]:
                o = env.reset()
                if o is not None:
                    nobs.append(o)
    except KeyError:
        status = "keyerror"
    assert status == str(g[f"cp_{tag}_status"])
    assert np.array_equal(acts, g[f"cp_{tag}_a"]) and np.array_equal(dones, g[f"cp_{tag}_d"])
    assert np.array_equal(np.array(nobs, dtype=np.float32).reshape(-1, 4), g[f"cp_{tag}_obs"])
    assert int(env._impl._sim.draws[0]) == int(g[f"cp_{tag}_draws"])


def test_reference_test_per_state_rejection_and_baselines(ob):
    """tests/test_per_state_rejection.py:7-20 and tests/test_trivial_baselines.py:8-31 of the reference."""
    from offsim4rl_b200.data import OfflineDataset, ProbDistribution
    from offsim4rl_b200 import spaces
    from offsim4rl_b200.evaluators import (PerStateRejectionSampling, FollowObservationOnly, FollowActionOnly,
                                           ServeRandomTransitions)
    dataset = OfflineDataset(
        observation_space=spaces.Discrete(25), action_space=spaces.Discrete(4), action_dist_type=ProbDistribution.Discrete,
        observations=np.array([0, 5], dtype=np.int64), actions=np.array([0, 1], dtype=np.int64),
        action_distributions=np.full((2, 4), fill_value=0.25, dtype=np.float32), rewards=np.array([0.0, 1.0], dtype=np.float32),
        next_observations=np.array([5, 7], dtype=np.int64), terminals=np.array([False, True], dtype=bool))
    uniform = np.full((4,), 0.25)
    for cls in (PerStateRejectionSampling, FollowObservationOnly, FollowActionOnly, ServeRandomTransitions):
        env = cls(dataset)
        obs = env.reset()
        assert obs is None or np.shape(obs) == tuple()
        env.step_dist(uniform)
        with pytest.raises(NotImplementedError):
            env.step(0)


def test_wrapper_validation_errors(ob):
    from offsim4rl_b200.data import OfflineDataset, ProbDistribution
    from offsim4rl_b200 import spaces
    from offsim4rl_b200.encoders import CartpoleBoxEncoder
    from offsim4rl_b200.evaluators import PerStateRejectionSampling
    obs = np.zeros((3, 4), dtype=np.float32)
    kw = dict(action_dist_type=ProbDistribution.Discrete, observations=obs, actions=np.zeros(3, dtype=np.int64),
              action_distributions=np.full((3, 2), 0.5, dtype=np.float32), rewards=np.zeros(3, dtype=np.float32),
              next_observations=obs, terminals=np.zeros(3, dtype=bool))
    box = OfflineDataset(observation_space=spaces.Box(-1, 1, (4,)), action_space=spaces.Discrete(2), **kw)
    with pytest.raises(ValueError, match="only supports discrete observation"):
        PerStateRejectionSampling(box)                                           # per_state_rejection.py:16-17
    with pytest.raises(ValueError, match="both need"):
        PerStateRejectionSampling(box, num_states=162)                           # :19-20
    cont = OfflineDataset(observation_space=spaces.Box(-1, 1, (4,)), action_space=spaces.Box(-1, 1, (1,)), **kw)
    with pytest.raises(ValueError, match="discrete action"):
        PerStateRejectionSampling(cont, num_states=162, encoder=CartpoleBoxEncoder())   # :22-25
    with pytest.raises(ValueError, match="Missing required key"):
        OfflineDataset(observation_space=spaces.Discrete(2), action_space=spaces.Discrete(2),
                       action_dist_type=ProbDistribution.Discrete, observations=obs)


def test_torch_distribution_is_accepted(ob, golden):
    import torch
    from offsim4rl_b200.evaluators import PerStateRejectionSampling
    g = golden("wrapper")
    env = PerStateRejectionSampling(_disc_dataset(ob, g))
    env.reset_sampler(seed=2)
    env._impl.feed_uniforms(g["disc_U"])
    env.reset()
    dist = torch.distributions.Categorical(probs=torch.tensor(g["disc_target"]))
    a, o, r, d, info = env.step_dist(dist)
    assert a == g["disc_old_a"][0] and o == g["disc_old_obs"][0]


# ----------------------------------------------------------------------------- QueueEvaluator
@pytest.mark.parametrize("seed", [0, 4])
def test_queue_evaluator_matches_reference(ob, golden, seed):
    from offsim4rl_b200.evaluators.queue_evaluator import QueueEvaluator_impl
    g = golden("queue_evaluator")
    rows = [(int(g["z"][i]), g["a"][i], g["r"][i], int(g["z_next"][i]), bool(g["done"][i]), g["p_log"][i],
             {"z": int(g["z"][i]), "next_z": int(g["z_next"][i]), "t": int(g["t"][i])}) for i in range(len(g["z"]))]
    env = QueueEvaluator_impl(rows, nS=6, nA=3)
    env.reset_sampler(seed=seed)
    assert np.array_equal(env._cur.offsets.cpu().numpy(), g[f"offsets_{seed}"])
    assert np.array_equal(env._cur.order.cpu().numpy(), g[f"order_{seed}"])          # queue_evaluator.py:107,110-111
    expect = g[f"trace_{seed}"]
    n, resets, k, status = 0, [], 0, "exhausted"
    s = env.reset()
    resets.append(s)
    try:
        while s is not None:
            s, r, done, info = env.step(g["acts"][k]); k += 1
            if s is None:
                break
            row = int(expect[n]); n += 1
            assert s == g["z_next"][row] and r == g["r"][row] and done == g["done"][row]
            assert info["z"] == g["z"][row] and info["next_z"] == g["z_next"][row] and info["a"] == g["a"][row]
            if done:
                s = env.reset()
                resets.append(-1 if s is None else s)
    except KeyError:
        status = "keyerror"
    assert n == expect.size and np.array_equal(np.array(resets, dtype=np.int64), g[f"resets_{seed}"])
    assert status == str(g[f"status_{seed}"]) and k == int(g[f"nacts_{seed}"])


@pytest.mark.parametrize("seed", [0, 4])
def test_fused_queue_evaluator_run_matches_reference(ob, golden, seed):
    """BatchedPSRS.run_actions (ols_run_actions: reset / step(a) for many steps in ONE launch) under the fixed action sequence
    the unmodified reference's QueueEvaluator_impl was driven with: same served rows, same end of run, same action count."""
    import torch
    from offsim4rl_b200 import _abi as abi
    g = golden("queue_evaluator")
    N, nS, nA = len(g["z"]), 6, 3
    log = ob.BucketedLog(z=g["z"], z_next=g["z_next"], action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"],
                         steps=g["t"], n_buckets=nS, key_by_action=True).seeded(seed)
    sim = ob.BatchedPSRS(log, 3, rng="philox", seed=1, reject="accept_all")
    res = sim.run_actions(N, gamma=0.9, actions=g["acts"], trace_cap=N)
    expect = g[f"trace_{seed}"]
    want_status = abi.NO_QUEUE if str(g[f"status_{seed}"]) == "keyerror" else None
    for r in range(3):
        n = int(res.accepted[r])
        assert n == expect.size and np.array_equal(res.trace_row[r, :n].cpu().numpy(), expect)
        assert int(sim.ties[r]) == int(g[f"nacts_{seed}"])                       # actions taken, incl. the one that found an empty queue
        st = int(res.status[r])
        assert st == want_status if want_status is not None else st in (abi.QUEUE_EMPTY, abi.INIT_EMPTY)
        # returns of the completed episodes and all lengths, from the served rows
        ends = np.nonzero(g["done"][expect])[0]
        lens = np.diff(np.r_[-1, ends])
        ne = int(res.n_returns[r])
        assert ne == ends.size and np.array_equal(res.lengths[r, :ne].cpu().numpy(), lens)
        Gs, k = [], 0
        for L in lens:
            Gs.append(sum(0.9 ** t * float(g["r"][expect[k + t]]) for t in range(L))); k += L
        np.testing.assert_allclose(res.Gs[r, :ne].cpu().numpy(), Gs, rtol=1e-12, atol=1e-12)
    # chunked == one launch
    sim2 = ob.BatchedPSRS(log, 3, rng="philox", seed=1, reject="accept_all")
    out = None
    for _ in range(N):
        out = sim2.run_actions(N, gamma=0.9, actions=g["acts"], trace_cap=N, max_steps=7, outputs=out)
        if bool((out.status != abi.RUNNING).all()):
            break
    assert torch.equal(out.trace_row, res.trace_row) and torch.equal(out.Gs, res.Gs) and torch.equal(out.status, res.status)


def test_fused_queue_evaluator_run_with_a_sampled_policy_matches_the_oracle(ob):
    """The agent samples its own actions from a fixed policy with the replica's counter-based stream (protocol in
    csrc/ols_queue_run.cu, restated here with tests/helpers.py): the python oracle's QueueEvaluator, driven with the same
    actions, serves the same rows — thousands of replicas, ring offsets off (the oracle class has none)."""
    from oracle import psrs_oracle as O
    from tests.helpers import stream_word
    from offsim4rl_b200 import _abi as abi
    rng = np.random.default_rng(12)
    N, nS, nA, R = 30000, 12, 4, 300
    z, zn = rng.integers(0, nS, N), rng.integers(0, nS, N)
    a = rng.integers(0, nA, N)
    done = rng.random(N) < 0.1
    t = np.where(np.r_[True, done[:-1]], 0, 1)
    r = rng.normal(size=N).astype(np.float32)
    pi = rng.dirichlet(np.ones(nA), size=nS)
    log = ob.BucketedLog(z=z, z_next=zn, action=a, reward=r, terminal=done, p_log=np.full((N, nA), 0.25, np.float32), steps=t,
                         n_buckets=nS, key_by_action=True)
    sim = ob.BatchedPSRS(log, R, rng="philox", seed=77, replica_id0=4000, reject="accept_all")
    res = sim.run_actions(25, gamma=0.95, pi=pi, trace_cap=2000)
    cum = np.cumsum(pi, axis=1)
    for rep in (0, 7, R - 1):
        env = O.OracleQueueEvaluator(z, zn, z, zn, a, steps=t, n_state_keys=nS, n_actions=nA)
        trace, k, episodes, status = [], 0, 0, None
        while episodes < 25:
            s = env.reset()
            if s is None:
                status = abi.INIT_EMPTY
                break
            while True:
                v = stream_word(77, 4000 + rep, 1, k) * (1.0 / 4294967296.0)
                act = min(int((cum[int(env.cur_s)] <= v).sum()), nA - 1)
                k += 1
                row = env.step(act)
                if row is None:
                    status = abi.QUEUE_EMPTY
                    break
                trace.append(row)
                if done[row]:
                    episodes += 1
                    break
            if status is not None:
                break
        status = abi.DONE if status is None else status
        n = int(res.accepted[rep])
        assert n == len(trace) and np.array_equal(res.trace_row[rep, :n].cpu().numpy(), trace)
        assert int(res.status[rep]) == status and int(sim.ties[rep]) == k


def test_batched_queue_evaluator_and_dataset_front_end(ob, golden):
    import torch
    from oracle import psrs_oracle as O
    from offsim4rl_b200.data import OfflineDataset, ProbDistribution
    from offsim4rl_b200 import spaces
    from offsim4rl_b200.evaluators import QueueEvaluator
    g = golden("queue_evaluator")
    # batched: R replicas, each with its own action sequence, against independent oracle instances (log order)
    R = 16
    rng = np.random.default_rng(0)
    log = ob.BucketedLog(z=g["z"], z_next=g["z_next"], action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"],
                         steps=g["t"], n_buckets=6, key_by_action=True)
    sim = ob.BatchedPSRS(log, R, reject="accept_all")
    envs = [O.OracleQueueEvaluator(z=g["z"], z_next=g["z_next"], s=g["z"], s_next=g["z_next"], action=g["a"], steps=g["t"],
                                   n_state_keys=6, n_actions=3) for _ in range(R)]
    sim.reset()
    for e in envs:
        e.reset()
    alive = np.ones(R, dtype=bool)
    for it in range(60):
        acts = rng.integers(0, 3, R)
        out = sim.step_action(acts, mask=alive.astype(np.uint8))
        rows = out["row"].cpu().numpy()
        for r, e in enumerate(envs):
            if not alive[r]:
                continue
            try:
                row = e.step(acts[r])
            except KeyError:
                row = None
                assert int(out["status"][r]) == O.NO_QUEUE
            if row is None:
                assert rows[r] == -1
                alive[r] = False
            else:
                assert rows[r] == row
    # dataset front end (queue_evaluator.py:8-88)
    ds = OfflineDataset(observation_space=spaces.Discrete(6), action_space=spaces.Discrete(3),
                        action_dist_type=ProbDistribution.Discrete, observations=g["z"], actions=g["a"],
                        action_distributions=g["p_log"], rewards=g["r"], next_observations=g["z_next"], terminals=g["done"],
                        steps=g["t"])
    qe = QueueEvaluator(ds)
    qe.reset_sampler(seed=0)
    obs = qe.reset()
    assert obs == g["resets_0"][0]
    nxt, r, done, info = qe.step(int(g["acts"][0]))
    assert nxt == g["z_next"][g["trace_0"][0]]
    a, nxt, r, done, info = qe.step_dist(torch.distributions.Categorical(probs=torch.tensor([0.0, 1.0, 0.0])))
    assert a == 1


# ----------------------------------------------------------------------------- PSRS_Exo
@pytest.mark.parametrize("seed", [1, 6])
def test_psrs_exo_matches_reference(ob, golden, seed):
    """Two queue families popped in lock-step, the generator re-seeded by reset(seed=episode): same observations
    and accepted rows as the reference's PSRS_Exo running on its own NumPy generator."""
    from offsim4rl_b200.evaluators.psrs import PSRS_Exo
    g = golden("psrs_exo")
    nX, N = 3, len(g["z"])
    o, o_next = g["z"] * nX + g["x"], g["z_next"] * nX + g["x_next"]
    rows = [(int(o[i]), int(g["a"][i]), g["r"][i], int(o_next[i]), bool(g["done"][i]), g["p_log"][i], {"t": int(g["t"][i])})
            for i in range(N)]
    env = PSRS_Exo(rows, nO=15, nA=3, o_split_func=lambda v: divmod(v, nX), o_combine_func=lambda s, x: s * nX + x)
    env.reset_sampler(seed=seed)
    expect, obs, n, episode = g[f"trace_{seed}"], [], 0, 0
    ob_ = env.reset(seed=episode)
    obs.append(-1 if ob_ is None else ob_)
    while ob_ is not None:
        ob_, r, done, info = env.step(g["p_new"])
        if ob_ is None:
            break
        row = int(expect[n]); n += 1
        assert r == g["r"][row] and done == g["done"][row] and info["a"] == g["a"][row] and info["s"] == g["z"][row]
        obs.append(ob_)
        if done:
            episode += 1
            ob_ = env.reset(seed=episode)
            obs.append(-1 if ob_ is None else ob_)
    assert n == expect.size and np.array_equal(np.array(obs, dtype=np.int64), g[f"obs_{seed}"])


@pytest.mark.parametrize("seed", [1, 6])
def test_batched_psrs_exo_matches_reference(ob, golden, seed):
    """BatchedPSRSExo — the vector form of PSRS_Exo: five replicas driven like the reference run of the golden file (reset with
    seed = episode index re-seeds each generator, psrs.py:92), every one serving the golden's rows and observations."""
    import torch
    from offsim4rl_b200 import _abi as abi
    g = golden("psrs_exo")
    nX, N, R = 3, len(g["z"]), 5
    common = dict(action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"], steps=g["t"])
    log_s = ob.BucketedLog(z=g["z"], z_next=g["z_next"], **common).seeded(seed)
    log_x = ob.BucketedLog(z=g["x"], z_next=g["x_next"], **common).seeded(seed)
    sim = ob.BatchedPSRSExo(log_s, log_x, R, seed=seed)
    expect, want_obs = g[f"trace_{seed}"], g[f"obs_{seed}"]
    episode = torch.zeros(R, dtype=torch.int64)
    s, x, status = sim.reset(seeds=episode.numpy())
    obs = [[int(s[r]) * nX + int(x[r]) if int(status[r]) == abi.RUNNING else -1] for r in range(R)]
    rows = [[] for _ in range(R)]
    alive = (status == abi.RUNNING).cpu()
    for _ in range(4 * N):
        if not bool(alive.any()):
            break
        out = sim.step(g["p_new"], mask=alive.to(torch.uint8))
        row, term = out["row"].cpu(), out["terminal"].cpu().bool()
        s2, x2 = out["s_next"].cpu(), out["x_next"].cpu()
        stepped = alive & (row >= 0)
        for r in torch.nonzero(stepped).flatten().tolist():
            rows[r].append(int(row[r]))
            obs[r].append(int(s2[r]) * nX + int(x2[r]))
        alive = stepped.clone()
        ended = stepped & term
        if bool(ended.any()):
            episode += ended.long()
            s, x, status = sim.reset(seeds=episode.numpy(), mask=ended.to(torch.uint8))
            for r in torch.nonzero(ended).flatten().tolist():
                ok = int(status[r]) == abi.RUNNING
                obs[r].append(int(s[r]) * nX + int(x[r]) if ok else -1)
                alive[r] = ok
    for r in range(R):
        assert np.array_equal(rows[r], expect) and np.array_equal(obs[r], want_obs)


@pytest.mark.parametrize("seed", [1, 6])
def test_fused_psrs_exo_evaluation_matches_reference(ob, golden, seed):
    """BatchedPSRSExo.run_evalmc (ols_run_exo: evalMC_psrs over PSRS_Exo in ONE launch, the generator re-seeded with the
    episode index at every reset) serves the rows of the reference's own PSRS_Exo run (golden psrs_exo: a policy with the same
    row for every observation), returns and lengths following from them; chunked launches equal one launch."""
    import torch
    from offsim4rl_b200 import _abi as abi
    g = golden("psrs_exo")
    nX, nS, nA, N, R = 3, 5, 3, len(g["z"]), 4
    common = dict(action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"], steps=g["t"])
    log_s = ob.BucketedLog(z=g["z"], z_next=g["z_next"], **common).seeded(seed)
    log_x = ob.BucketedLog(z=g["x"], z_next=g["x_next"], **common).seeded(seed)
    pi = np.tile(g["p_new"], (nS * nX, 1))
    expect = g[f"trace_{seed}"]
    sim = ob.BatchedPSRSExo(log_s, log_x, R, seed=seed)
    res = sim.run_evalmc(nS * nX, N, pi, gamma=0.9, x_mult=nX, trace_cap=N)
    ends = np.nonzero(g["done"][expect])[0]
    lens = np.diff(np.r_[-1, ends])
    for r in range(R):
        n = int(res.accepted[r])
        assert n == expect.size and np.array_equal(res.trace_row[r, :n].cpu().numpy(), expect)
        assert int(res.status[r]) in (abi.QUEUE_EMPTY, abi.INIT_EMPTY)
        ne = int(res.n_returns[r])
        assert ne == ends.size and np.array_equal(res.lengths[r, :ne].cpu().numpy(), lens)
        Gs, k = [], 0
        for L in lens:
            Gs.append(sum(0.9 ** t * float(g["r"][expect[k + t]]) for t in range(L))); k += L
        np.testing.assert_allclose(res.Gs[r, :ne].cpu().numpy(), Gs, rtol=1e-12, atol=1e-12)
    sim2 = ob.BatchedPSRSExo(log_s, log_x, R, seed=seed)
    out = None
    for _ in range(N):
        out = sim2.run_evalmc(nS * nX, N, pi, gamma=0.9, x_mult=nX, trace_cap=N, max_steps=5, outputs=out)
        if bool((out.status != abi.RUNNING).all()):
            break
    assert torch.equal(out.trace_row, res.trace_row) and torch.equal(out.Gs, res.Gs) and torch.equal(out.status, res.status)


# --------------------------------------------------------------------------- ingest
@pytest.mark.gpu
def test_recorded_dataset_flows_into_the_simulator(ob, tmp_path):
    """record -> save -> load -> device columns -> bucketing -> fused run, against the python oracle on the same
    columns (fed uniforms and tie draws: bit-exact rows, float64 Q)."""
    import torch
    from tests.helpers import ToyChainEnv, ToyAgent
    from offsim4rl_b200.utils import record_dataset_in_memory
    from offsim4rl_b200.data import OfflineDataset, ProbDistribution
    from oracle import psrs_oracle as O
    np.random.seed(11)
    ds = record_dataset_in_memory(ToyChainEnv(n=7), ToyAgent(ProbDistribution.Discrete), 4000, seed=2)
    ds.save_npz(tmp_path / "toy.npz")
    ds = OfflineDataset.load_npz(tmp_path / "toy.npz")
    nS, nA = 7, 3
    log = ob.BucketedLog.from_dataset(ds, n_buckets=nS)
    cols = ds.device_columns()
    rng = np.random.default_rng(4)
    U, T = rng.random(20000), rng.random(5000)
    sim = ob.BatchedPSRS(log, 3, rng="feed", u_feed=U, tie_feed=T, lanes=4)
    res = sim.run_qlearn(nS, 30, gamma=0.9, alpha=0.1, epsilon=0.2, q_dtype=torch.float64, trace_cap=4000)
    env = O.OraclePSRS(z=cols["z"], z_next=cols["z_next"], s=cols["z"], s_next=cols["z_next"], action=cols["action"],
                       reward=cols["reward"], terminal=cols["terminal"], p_log=cols["p_log"], steps=cols["steps"],
                       n_buckets=nS, uniforms=O.FeedUniforms(U))
    Q, info = O.qlearn(env, 30, gamma=0.9, alpha=0.1, epsilon=0.2, ties=O.FeedTies(T), nS=nS, nA=nA)
    n = len(info["trace"])
    assert n > 100
    for r in range(3):
        assert int(res.accepted[r]) == n and np.array_equal(res.trace_row[r, :n].cpu().numpy(), info["trace"])
        assert np.array_equal(res.Q[r].cpu().numpy(), Q)
    # ids handed over as device tensors (an encoder's output) give the same queues
    log2 = ob.BucketedLog.from_dataset(ds, n_buckets=nS, z=torch.as_tensor(cols["z"]).cuda(), z_next=torch.as_tensor(cols["z_next"]).cuda())
    assert torch.equal(log2.offsets, log.offsets) and torch.equal(log2.order, log.order)
