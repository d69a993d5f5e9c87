"""Scratch experiment: does the H2D upload of the next log overlap with the fused run kernel?  Times both alone and together."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import offsim4rl_b200 as ob

dev = torch.device("cuda:0")
N, nS, nA, R, T = 100_000_000, 10_000, 4, 125_000, 10_000
rng = np.random.default_rng(0)
pb = rng.dirichlet(np.ones(nA), size=nS); pb = np.maximum(pb, 0.05); pb /= pb.sum(1, keepdims=True)
cols = ob.synth_columns(N, nS, pb, seed=0, device=dev)
host = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in cols.items()}
for k, v in cols.items():
    host[k].copy_(v)
log = ob.BucketedLog(z=cols["z"], z_next=cols["z_next"], action=cols["action"], reward=cols["reward"], terminal=cols["terminal"],
                     p_log=cols["p_log"], steps=cols["steps"], n_buckets=nS, z_offset=0, device=dev, hot_records=True)
dst = {k: torch.empty_like(v) for k, v in cols.items()}
sim = ob.BatchedPSRS(log, R, rng="philox", seed=1, ring=True, ring_seed=9, lanes=2, aos_state=True)
copy_stream = torch.cuda.Stream(device=dev)
res = None

def run():
    global res
    sim.restart()
    res = sim.run_qlearn(nS, 511, 0.99, alpha=0.1, epsilon=0.1, max_steps=T, ep_cap=512, outputs=res)

def ev():
    return torch.cuda.Event(enable_timing=True)

def copy(stream):
    with torch.cuda.stream(stream):
        a, b = ev(), ev()
        a.record(stream)
        for k, v in host.items():
            dst[k].copy_(v, non_blocking=True)
        b.record(stream)
    return a, b

for _ in range(2):
    run()
torch.cuda.synchronize()
a, b = copy(copy_stream); torch.cuda.synchronize(); print("copy alone ms", a.elapsed_time(b))
r0, r1 = ev(), ev(); r0.record(); run(); r1.record(); torch.cuda.synchronize(); print("run alone ms", r0.elapsed_time(r1))
for order in ("copy_first", "run_first"):
    r0, r1 = ev(), ev()
    if order == "copy_first":
        a, b = copy(copy_stream)
    r0.record(); run(); r1.record()
    if order == "run_first":
        a, b = copy(copy_stream)
    torch.cuda.synchronize()
    print(order, "copy ms", a.elapsed_time(b), "run ms", r0.elapsed_time(r1), "span ms", min(a, r0, key=lambda e: 0) and max(r0.elapsed_time(b), r0.elapsed_time(r1)))
