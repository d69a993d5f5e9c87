"""N > 1 host path on CPU: replica sharding + the final curve reduction over a world_size-2 gloo group.
Each rank simulates its shard of replicas with the C oracle (same Philox streams as the GPU path, keyed by
GLOBAL replica id), reduces its (sum, sumsq, count) accumulators, and the all-reduced curve must equal the
single-process curve over all replicas."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_log():
    rng = np.random.default_rng(8)
    N, nS, nA = 20000, 15, 4
    z = rng.integers(0, nS, N); zn = rng.integers(0, nS, N)
    pb = rng.dirichlet(np.ones(nA), size=nS).clip(0.05); pb /= pb.sum(1, keepdims=True)
    a = (rng.random(N)[:, None] > np.cumsum(pb[z], 1)).sum(1).clip(0, nA - 1)
    done = rng.random(N) < 0.1
    return dict(z=z, z_next=zn, action=a, reward=rng.normal(size=N).astype(np.float32), terminal=done,
                p_log=pb[z].astype(np.float32), steps=np.where(np.r_[True, done[:-1]], 0, 1)), nS


def _curve(id0, n_local, E=12):
    from oracle import c_oracle as CO, psrs_oracle as O
    cols, nS = _make_log()
    log = CO.Log(n_buckets=nS, **cols)
    out = CO.run(log, n_local, CO.QLEARN, nS, E, gamma=0.9, alpha=0.1, epsilon=0.2, seed=77, replica_id0=id0, ring=True,
                 ring_seed=3)
    s, s2, c = O.learning_curve([out["Gs"][r, :out["n_returns"][r]] for r in range(n_local)])
    acc = np.zeros((3, E + 1))
    acc[0, :s.size], acc[1, :s.size], acc[2, :s.size] = s, s2, c
    return acc


def _worker(rank, world, port, total, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import offsim4rl_b200.parallel as par
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    id0, n_local = par.replica_shard(total, rank, world)
    acc = torch.from_numpy(_curve(id0, n_local))
    par.all_reduce_sum_(acc)
    if rank == 0:
        q.put(acc.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _ingest_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import offsim4rl_b200.parallel as par
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cols, _ = _make_log()
    cols = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in cols.items()}
    n = cols["z"].shape[0]
    lo, hi = par.row_shard(n, rank, world)
    shards = {k: v[lo:hi].contiguous() for k, v in cols.items()}
    full = {k: torch.empty_like(v) for k, v in cols.items()}
    par.all_gather_columns(shards, full)
    ok = all(torch.equal(full[k], cols[k]) for k in cols)
    if rank == 1:
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharded_ingest_rebuilds_every_column():
    """Sharded ingest: each rank contributes its row shard of every log column (1-D and [n, nA]); after the all-gather every
    rank holds the whole log, bit for bit."""
    import offsim4rl_b200.parallel as par
    assert par.row_shard(100, 1, 4) == (25, 50)
    with pytest.raises(ValueError):
        par.row_shard(10, 0, 3)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ingest_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    assert q.get(timeout=240) is True
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # no process group: the shard is the column
    a = {"x": torch.arange(6).reshape(3, 2)}
    b = {"x": torch.empty(3, 2, dtype=torch.int64)}
    par.all_gather_columns(a, b)
    assert torch.equal(a["x"], b["x"])


def test_replica_shard_covers_all_ids_once():
    import offsim4rl_b200.parallel as par
    for total in (0, 1, 7, 1000, 10**6 + 3):
        for world in (1, 2, 3, 8):
            blocks = [par.replica_shard(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(n for _, n in blocks) == total
            for (a, n), (b, _) in zip(blocks, blocks[1:]):
                assert a + n == b
            assert max(n for _, n in blocks) - min(n for _, n in blocks) <= 1
    with pytest.raises(ValueError):
        par.replica_shard(10, 2, 2)


@pytest.mark.timeout(300)
def test_two_rank_gloo_curve_reduction_equals_single_process():
    total, world = 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _curve(0, total)
    assert np.array_equal(got[2], want[2])                       # counts per episode index
    np.testing.assert_allclose(got[:2], want[:2], rtol=1e-12, atol=1e-12)
    import offsim4rl_b200.parallel as par
    mean, std = par.curve_stats(*(torch.from_numpy(x) for x in got))
    assert mean.shape == (13,) and bool((std >= 0).all())
