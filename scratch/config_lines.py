"""Throughput lines for BASELINE.json configs 0-2 at their named shapes (C1 one replica, C2 10k replicas, C3 64k replicas).
These configs are parity-test cases (tests/test_gpu_configs.py), not the bench headline; this script only TIMES the same
pipelines — observations -> device encoder -> bucketing -> fused run — with CUDA events and prints one JSON line each.

    python scratch/config_lines.py > gpurun_out/config_lines.jsonl
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import offsim4rl_b200 as ob                                      # noqa: E402
from offsim4rl_b200.encoders import CartpoleBoxEncoder, GridEncoder, HOMEREncoder   # noqa: E402
from tests.test_gpu_configs import grid_log                      # noqa: E402

PEAK = 6547.2
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps=5, warm=2):
    ms = []
    for i in range(reps + warm):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ms.append(e0.elapsed_time(e1))
    return float(np.mean(ms)), out


def line(name, sim, run, encode_ms, bucket_ms, rec_bytes, extra=None):
    def go():
        sim.restart()
        return run()
    ms, res = timed(go)
    acc, pops = float(sim.accepted.sum()), float(sim.pops.sum())
    out = {"config": name, "replicas": sim.R, "accepted_per_s": acc / ms * 1e3, "pops_per_s": pops / ms * 1e3, "kernel_plus_restart_ms": ms,
           "accepted_per_launch": acc, "encode_ms": encode_ms, "bucket_ms": bucket_ms,
           "gather_GBps_contract": pops * rec_bytes / ms / 1e6, "frac_of_hbm_peak": pops * rec_bytes / ms / 1e6 / PEAK}
    out.update(extra or {})
    print(json.dumps(out), flush=True)


def main():
    dev = torch.device("cuda:0")
    # ---- C1: continuous_grid, grid discretiser, ONE replica, Q-learning (the reference's own CPU-runnable case)
    N = 1_000_000
    g = grid_log(N, seed=1)
    enc = GridEncoder(cells=5)
    obs, nobs = torch.from_numpy(g["obs"]).to(dev), torch.from_numpy(g["next_obs"]).to(dev)
    e_ms, (z, zn) = timed(lambda: (enc.encode_device(obs), enc.encode_device(nobs)))
    mk = lambda **kw: ob.BucketedLog(z=z, z_next=zn, action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"], steps=g["t"],
                                     n_buckets=25, **kw)
    b_ms, log = timed(mk, reps=3, warm=1)
    sim = ob.BatchedPSRS(log, 1, rng="philox", seed=1, lanes=32)
    line("c1: continuous_grid, grid-discretised states, 1M transitions, 1 replica, Q-learning (100 episodes)", sim,
         lambda: sim.run_qlearn(25, 100, gamma=0.99, alpha=0.1, epsilon=0.1), e_ms, b_ms, log.rec_bytes,
         {"note": "one replica = one lane group: latency of a single dependent chain (one launch, no per-step host work)"})
    sim = ob.BatchedPSRS(log, 65_536, rng="philox", seed=1, ring=True, lanes=4)
    line("c1 log, 64k replicas, Q-learning (budget 500 steps)", sim,
         lambda: sim.run_qlearn(25, 511, gamma=0.99, alpha=0.1, epsilon=0.1, max_steps=500, ep_cap=512), e_ms, b_ms, log.rec_bytes)
    # ---- C3: the same log with latent-state MLP ids (dO = 2, H = 64, nZ = 25), 64k replicas
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(2, 64), torch.nn.LeakyReLU(), torch.nn.Linear(64, 25)).to(dev)
    xs, ys = obs[:50_000], torch.from_numpy(g["cell"][:50_000]).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=0.02)
    for _ in range(300):                                          # a fixture: make every latent state occur (not part of the path)
        opt.zero_grad()
        torch.nn.functional.cross_entropy(net(xs), ys).backward()
        opt.step()
    W1, b1, W2, b2 = (p.detach() for p in net.parameters())
    menc = HOMEREncoder.from_weights(W1, b1, W2, b2)
    e_ms, (z, zn) = timed(lambda: (menc.encode_device(obs), menc.encode_device(nobs)))
    b_ms, log = timed(mk, reps=3, warm=1)
    sim = ob.BatchedPSRS(log, 65_536, rng="philox", seed=11, ring=True, ring_seed=5, lanes=4)
    line("c3: continuous_grid with latent-state MLP encoder (H=64, nZ=25), 1M transitions, 64k replicas, Q-learning (budget 500 steps)",
         sim, lambda: sim.run_qlearn(25, 511, gamma=0.99, alpha=0.1, epsilon=0.1, max_steps=500, ep_cap=512), e_ms, b_ms, log.rec_bytes,
         {"encode_rows": 2 * N, "encode_GFLOPs": 2 * N * 2 * (2 * 64 + 64 * 25) / e_ms / 1e6})
    del sim, log
    # ---- C2: cartpole-shaped log, CartpoleBoxEncoder ids (162 boxes + sentinel), 2 actions, 10k replicas
    rng = np.random.default_rng(2)
    lim = np.array([2.4, 1.5, 0.2094384, 1.5])
    o1 = torch.from_numpy(rng.uniform(-lim, lim, (N, 4)).astype(np.float32)).to(dev)
    o2 = torch.from_numpy(rng.uniform(-lim * [1.04, 1, 1.03, 1], lim * [1.04, 1, 1.03, 1], (N, 4)).astype(np.float32)).to(dev)
    benc = CartpoleBoxEncoder()
    e_ms, (z, zn) = timed(lambda: (benc.encode_device(o1), benc.encode_device(o2)))
    z_h, zn_h = z.cpu().numpy().astype(np.int64), zn.cpu().numpy().astype(np.int64)
    done = zn_h < 0
    pb = rng.dirichlet(np.ones(2), size=162).clip(0.1)
    pb /= pb.sum(1, keepdims=True)
    a = (rng.random(N) > pb[z_h, 0]).astype(np.int64)
    idx = np.arange(N)
    t = idx - np.maximum.accumulate(np.where(np.r_[True, done[:-1]], idx, 0))
    mk2 = lambda: ob.BucketedLog(z=z, z_next=zn, s=z_h + 1, s_next=zn_h + 1, action=a, reward=np.ones(N, np.float32), terminal=done,
                                 p_log=pb[z_h].astype(np.float32), steps=t, n_buckets=163, z_offset=1)
    b_ms, log = timed(mk2, reps=3, warm=1)
    pi = rng.dirichlet(np.ones(2), size=163)
    for lanes in (8, 1):
        sim = ob.BatchedPSRS(log, 10_000, rng="philox", seed=7, ring=True, ring_seed=3, lanes=lanes)
        line("c2: cartpole PSRS with discretised-state encoder, 1M transitions, 10k replicas, Q-learning (budget 2000 steps), lanes=%d" % lanes,
             sim, lambda: sim.run_qlearn(163, 511, gamma=0.99, alpha=0.1, epsilon=0.1, max_steps=2000, ep_cap=512), e_ms, b_ms, log.rec_bytes,
             {"note": "10k replicas cannot fill 148 SMs (79 warps at lanes=1): bound by the latency of each replica's dependent chain"})
        sim = ob.BatchedPSRS(log, 10_000, rng="philox", seed=7, ring=True, ring_seed=3, lanes=lanes)
        line("c2, fixed-policy evaluation (evalMC), lanes=%d" % lanes, sim,
             lambda: sim.run_evalmc(163, 511, pi=pi, gamma=0.99, max_steps=2000, ep_cap=512), e_ms, b_ms, log.rec_bytes)


if __name__ == "__main__":
    main()
