"""BASELINE.json's full sizes (C4: 1e8 transitions, 10k states, 4 actions, 125k replicas per GPU; C5: 1000 states,
near-deterministic behaviour policy).  Two kinds of checks:
  * against the ORACLE: oracle/ols_oracle.c (float32-learner mode, pinned to the reference by tests/test_oracle_c.py)
    buckets the host copy of the same 1e8-row log and runs a few hundred SAMPLED global replica ids; the production
    kernels (hot records, AoS state, one lane per replica, time slices) must match them draw for draw — accepted rows,
    pops, ties, returns and the float32 Q tables bit for bit;
  * size-independent properties over ALL replicas: the bucketing is a stable sorting permutation that packed every
    field of every row, hot records decode to the records, runs conserve their counters, are idempotent and invariant
    under replica sharding, and every accelerated layout reproduces the plain one bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, NS, NA = 100_000_000, 10_000, 4


@pytest.fixture(scope="module")
def ob():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    if torch.cuda.get_device_properties(0).total_memory < 120 * 2 ** 30:
        pytest.skip("full-size tests need a 180 GB B200")
    import offsim4rl_b200 as ob
    return ob


def _behaviour(nS, nA, delta=None, seed=0):
    rng = np.random.default_rng(seed)
    if delta is None:
        pb = np.maximum(rng.dirichlet(np.ones(nA), size=nS), 0.05)
        return pb / pb.sum(1, keepdims=True)
    pb = np.full((nS, nA), delta / (nA - 1))
    pb[np.arange(nS), rng.integers(0, nA, nS)] = 1.0 - delta
    return pb


@pytest.fixture(scope="module")
def c4(ob):
    import torch
    cols = ob.synth_columns(N, NS, _behaviour(NS, NA), seed=0, device=torch.device("cuda:0"))
    log = ob.BucketedLog(z=cols["z"], z_next=cols["z_next"], action=cols["action"], reward=cols["reward"],
                         terminal=cols["terminal"], p_log=cols["p_log"], steps=cols["steps"], n_buckets=NS, z_offset=0,
                         hot_records=True)
    return cols, log


def _host_log(cols, nS):
    """The same log on the host, bucketed by the C oracle (4.5 GB of host memory at 1e8 rows)."""
    from oracle import c_oracle as CO
    h = {k: v.cpu().numpy() for k, v in cols.items()}
    return CO.Log(z=h["z"], z_next=h["z_next"], action=h["action"], reward=h["reward"], terminal=h["terminal"], p_log=h["p_log"],
                  steps=h["steps"], n_buckets=nS)


@pytest.fixture(scope="module")
def c4_host(c4):
    return _host_log(c4[0], NS)


def test_c4_production_learner_matches_c_oracle_on_sampled_replicas(ob, c4, c4_host):
    """The benchmarked C4 configuration at full size (1e8 rows, 10k states, 125k replicas on this GPU, float32 Q, hot
    records, AoS state, one lane per replica, auto time slices) against the C oracle on 320 sampled global replica ids."""
    import torch
    from oracle import c_oracle as CO
    cols, log8 = c4
    clog = c4_host
    # the benchmark's own layout: TAIL hot records (bench.py) — the synthetic behaviour policy is a function of the state
    from offsim4rl_b200 import _abi as abi
    log = ob.BucketedLog(z=cols["z"], z_next=cols["z_next"], action=cols["action"], reward=cols["reward"],
                         terminal=cols["terminal"], p_log=cols["p_log"], steps=cols["steps"], n_buckets=NS, z_offset=0,
                         hot_records="tail")
    assert log.hot_kind == abi.HOT_TAIL and torch.equal(log.records, log8.records)
    w = log.hot.view(torch.int32).view(-1, 2)[:N]
    rec = log.records.view(torch.int32).view(-1, 8)[:N]
    assert torch.equal(w[:, 0], rec[:, 4])                                                 # reward
    assert torch.equal(w[:, 1] & 0x1FFFFFFF, rec[:, 5]) and torch.equal((w[:, 1] >> 29) & 3, rec[:, 7] & 0xFF)
    assert torch.equal((w[:, 1] >> 31) & 1, (rec[:, 7] >> 8) & 1)
    first = log.offsets[:-1].long()
    assert torch.equal(log.queue_dist[:, :4].view(torch.int32), rec[first, :4])           # every queue's one distribution
    assert bool((log.queue_dist[:, 4].view(torch.int32) == 3).all())
    del w, rec
    assert np.array_equal(log.offsets.cpu().numpy(), clog.offsets)
    assert np.array_equal(log.order[:4096].cpu().numpy(), clog.order[:4096]) and np.array_equal(log.order[-4096:].cpu().numpy(), clog.order[N - 4096:N])
    R, T, id0 = 125_000, 300, 3 * 125_000                      # the shard of rank 3 of 8
    sim = ob.BatchedPSRS(log, R, rng="philox", seed=1234, ring=True, ring_seed=99, replica_id0=id0, lanes=1, aos_state=True)
    res = sim.run_qlearn(NS, 511, gamma=0.99, alpha=0.1, epsilon=0.1, max_steps=T, ep_cap=512, trace_cap=T)
    sample = np.unique(np.r_[0, R - 1, np.random.default_rng(5).integers(0, R, 320)])
    ref = CO.run(clog, sample.size, CO.QLEARN, NS, 511, gamma=0.99, alpha=0.1, epsilon=0.1, seed=1234, ring=True, ring_seed=99,
                 max_steps=T, ep_cap=512, trace_cap=T, threads=16, q_f32=True, replica_ids=id0 + sample)
    idx = torch.as_tensor(sample, device=res.Gs.device)
    assert np.array_equal(res.trace_row[idx].cpu().numpy(), ref["trace_row"])            # bit-exact accepted rows
    assert np.array_equal(res.accepted[idx].cpu().numpy(), ref["accepted"]) and bool((ref["accepted"] == T).all())
    assert np.array_equal(res.pops[idx].cpu().numpy(), ref["pops"]) and np.array_equal(sim.ties[idx].cpu().numpy(), ref["ties"])
    assert np.array_equal(res.n_returns[idx].cpu().numpy(), ref["n_returns"])
    mask = np.arange(512)[None, :] < ref["n_returns"][:, None]
    np.testing.assert_allclose(res.Gs[idx].cpu().numpy()[mask], ref["Gs"][mask], rtol=1e-12, atol=1e-12)
    assert np.array_equal(res.lengths[idx].cpu().numpy()[mask], ref["lengths"][mask])
    assert np.array_equal(res.q_rows(sample).cpu().numpy(), ref["Q"])                    # float32 Q tables, bit for bit


def test_c4_bucketing_is_a_stable_sort_that_packed_every_field(ob, c4):
    import torch
    cols, log = c4
    off, order = log.offsets.long(), log.order.long()
    assert int(off[0]) == 0 and int(off[-1]) == N and bool((off[1:] >= off[:-1]).all())
    # a permutation of 0..N-1: exact sum and a multiplicative checksum
    idx = torch.arange(N, device=order.device)
    assert int(order.sum()) == N * (N - 1) // 2
    assert int(((order * 2654435761) & 0xFFFFFFFF).sum()) == int(((idx * 2654435761) & 0xFFFFFFFF).sum())
    zs = cols["z"].long()[order]
    assert bool(((zs[1:] > zs[:-1]) | ((zs[1:] == zs[:-1]) & (order[1:] > order[:-1]))).all())      # sorted AND stable
    assert torch.equal(torch.bincount(cols["z"].long(), minlength=NS), off[1:] - off[:-1])
    del zs, idx
    rec = log.records.view(torch.int32).view(-1, 8)[:N]
    assert torch.equal(rec[:, 5], cols["z_next"][order])                                          # next state
    assert torch.equal(rec[:, 4].view(torch.float32), cols["reward"][order])
    assert torch.equal(rec[:, :4].view(torch.float32), cols["p_log"][order])
    meta = rec[:, 7]
    assert torch.equal(meta & 0xFF, cols["action"][order]) and torch.equal((meta >> 8) & 1, cols["terminal"][order].int())
    # initial states: the rows with step 0, in log order
    first = torch.nonzero(cols["steps"] == 0).flatten().int()
    assert log.n_init == first.numel() and torch.equal(log.init_row[: log.n_init], first)


def test_c4_hot_records_decode_to_the_records(ob, c4):
    import torch
    cols, log = c4
    words = log.hot.view(torch.int32).view(-1, 2)[:N]
    lo, hi = words[:, 0], words[:, 1]
    rec = log.records.view(torch.int32).view(-1, 8)[:N]
    action = ((lo >> 15) & 1) | (((lo >> 31) & 1) << 1)
    assert torch.equal(action, rec[:, 7] & 0xFF)
    assert bool((((hi >> 15) & 1) == 1).all())                       # every synthetic probability is in the half range
    halves = torch.stack([lo & 0x7FFF, (lo >> 16) & 0x7FFF, hi & 0x7FFF, (hi >> 16) & 0x7FFF], 1).to(torch.int16).view(torch.float16)
    p = rec[:, :4].view(torch.float32)
    assert torch.equal(halves, p.to(torch.float16))                  # round-to-nearest-even images
    assert float(((halves.float() - p).abs() / p).max()) <= 2.0 ** -11


def _run(ob, log, R, id0, budget, lanes=1, **kw):
    # lanes = 1 over hot records and AoS state is the production (bench.py) kernel
    sim = ob.BatchedPSRS(log, R, rng="philox", seed=1234, ring=True, ring_seed=99, replica_id0=id0, lanes=lanes, **kw)
    res = sim.run_qlearn(NS, 511, gamma=0.99, alpha=0.1, epsilon=0.1, max_steps=budget, ep_cap=512)
    return sim, res


def test_c4_run_conserves_is_idempotent_and_shards(ob, c4):
    import torch
    cols, log = c4
    R, T = 125_000, 300
    sim, res = _run(ob, log, R, 0, T, aos_state=True)
    acc, pops = res.accepted.clone(), res.pops.clone()
    from offsim4rl_b200 import _abi as abi
    assert bool((acc == T).all()) and bool((res.status == abi.RUNNING).all())   # every replica spent its whole budget, still RUNNING
    assert torch.equal(sim.draws, pops) and bool((pops >= acc).all())
    # the lengths of the closed episodes plus the open episode's clock account for every accepted transition
    nl = res.n_lengths.long()
    closed = (res.lengths.long() * (torch.arange(512, device=nl.device)[None, :] < nl[:, None])).sum(1)
    assert torch.equal(closed + sim.ep_t.long() * sim.ep_open.long(), acc.long())
    assert bool(torch.isfinite(res.Gs).all())
    gs_sum, len_sum = res.Gs.sum(1).clone(), res.lengths.sum(1).clone()
    q_check = res.Q.sum((1, 2)).clone()
    del sim, res
    torch.cuda.empty_cache()
    # idempotent, and independent of the accelerated layouts: two lanes per replica, separate cursor/Q arrays, 32-byte records,
    # no time slices
    plain = ob.BucketedLog.__new__(ob.BucketedLog)
    plain.__dict__.update(log.__dict__)
    plain.hot = None
    sim2, res2 = _run(ob, plain, R, 0, T, lanes=2, aos_state=False, slices=0)
    assert torch.equal(res2.accepted, acc) and torch.equal(res2.pops, pops)
    assert torch.equal(res2.Gs.sum(1), gs_sum) and torch.equal(res2.lengths.sum(1), len_sum)
    assert torch.equal(res2.Q.sum((1, 2)), q_check)
    del sim2, res2
    torch.cuda.empty_cache()
    # sharding: two ranks' worth of replicas, keyed by global replica id
    half = R // 2
    parts = []
    for id0 in (0, half):
        s, r = _run(ob, log, half, id0, T, aos_state=True)
        parts.append((r.pops.clone(), r.Gs.sum(1).clone()))
        del s, r
        torch.cuda.empty_cache()
    assert torch.equal(torch.cat([p[0] for p in parts]), pops) and torch.equal(torch.cat([p[1] for p in parts]), gs_sum)


def test_c5_evaluation_over_hot_records_at_full_size(ob):
    import torch
    nS, R, T = 1000, 32_000, 20
    cols = ob.synth_columns(N, nS, _behaviour(nS, NA, delta=1e-3), seed=0, device=torch.device("cuda:0"))
    pi = np.full((nS, NA), 0.25)
    outs = []
    for hot in ("tail", True, False):
        log = ob.BucketedLog(z=cols["z"], z_next=cols["z_next"], action=cols["action"], reward=cols["reward"],
                             terminal=cols["terminal"], p_log=cols["p_log"], steps=cols["steps"], n_buckets=nS, z_offset=0,
                             hot_records=hot)
        sim = ob.BatchedPSRS(log, R, rng="philox", seed=4321, ring=True, ring_seed=7, lanes=32)
        res = sim.run_evalmc(nS, 511, pi, 0.99, max_steps=T, ep_cap=512, trace_cap=T)
        outs.append((res.pops.clone(), res.accepted.clone(), res.trace_row.clone(), res.Gs.sum(1).clone()))
        del sim, res, log
        torch.cuda.empty_cache()
    for other in outs[1:]:                                                           # TAIL, PROBS and no hot records agree
        for a, b in zip(outs[0], other):
            assert torch.equal(a, b)
    # against the C oracle on sampled replicas (the streaming reject chains: ~750 pops per accepted transition)
    from oracle import c_oracle as CO
    clog = _host_log(cols, nS)
    sample = np.unique(np.r_[0, R - 1, np.random.default_rng(6).integers(0, R, 256)])
    ref = CO.run(clog, sample.size, CO.FIXED_PI, nS, 511, gamma=0.99, pi=pi, seed=4321, ring=True, ring_seed=7, max_steps=T,
                 ep_cap=512, trace_cap=T, threads=16, replica_ids=sample)
    idx = torch.as_tensor(sample, device=outs[0][0].device)
    assert np.array_equal(outs[0][2][idx].cpu().numpy(), ref["trace_row"])               # bit-exact accepted rows
    assert np.array_equal(outs[0][0][idx].cpu().numpy(), ref["pops"]) and np.array_equal(outs[0][1][idx].cpu().numpy(), ref["accepted"])
    np.testing.assert_allclose(outs[0][3][idx].cpu().numpy(), ref["Gs"].sum(1), rtol=1e-12, atol=1e-12)
    pops, acc = outs[0][0].double().sum(), outs[0][1].double().sum()
    assert bool((outs[0][1] == T).all()) and 500 < float(pops / acc) < 1100          # one accept per ~750 pops (M = 750)


@pytest.mark.parametrize("seed", [0, 98052])
def test_c4_seeded_orders_are_numpy_shuffles(ob, c4, seed):
    """reset_sampler(seed) at full size: the initial-state list (2e6 entries) and the queues are exactly what
    np.random.default_rng(seed).shuffle makes of the log-order lists (psrs.py:23,29-30)."""
    cols, log = c4
    s = log.seeded(seed)
    init = log.init_row[: log.n_init].cpu().numpy().copy()
    np.random.default_rng(seed).shuffle(init)
    assert np.array_equal(s.init_row[: s.n_init].cpu().numpy(), init)
    off = log.offsets.cpu().numpy()
    base_order, new_order = log.order, s.order
    for b in (0, 1, 4999, NS - 1):
        q = base_order[off[b]:off[b + 1]].cpu().numpy().copy()
        np.random.default_rng(seed).shuffle(q)
        assert np.array_equal(new_order[off[b]:off[b + 1]].cpu().numpy(), q)
    # hot records follow their records
    rec = s.records.view(s.records.dtype).view(-1, 8)
    import torch
    words = s.hot.view(torch.int32).view(-1, 2)[:N]
    assert torch.equal(((words[:, 0] >> 15) & 1) | (((words[:, 0] >> 31) & 1) << 1), s.records.view(torch.int32).view(-1, 8)[:N, 7] & 0xFF)


def test_long_queues_seeded_order_is_numpy_shuffle(ob):
    """A few very long queues (2.5e6 records each): the regime where the swap rounds of the shuffle run in parallel."""
    import torch
    n, nS = 10_000_000, 4
    cols = ob.synth_columns(n, nS, _behaviour(nS, NA), seed=1, device=torch.device("cuda:0"))
    log = ob.BucketedLog(z=cols["z"], z_next=cols["z_next"], action=cols["action"], reward=cols["reward"],
                         terminal=cols["terminal"], p_log=cols["p_log"], steps=cols["steps"], n_buckets=nS, z_offset=0)
    s = log.seeded(7)
    off = log.offsets.cpu().numpy()
    for b in range(nS):
        q = log.order[off[b]:off[b + 1]].cpu().numpy().copy()
        np.random.default_rng(7).shuffle(q)
        assert np.array_equal(s.order[off[b]:off[b + 1]].cpu().numpy(), q)
