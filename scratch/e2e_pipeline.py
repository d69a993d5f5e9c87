"""Scratch experiment: the e2e step loop with the next upload issued behind the run launch; host and device timings per phase."""
import os, sys, time
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
del cols
dev_cols = [{k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in host.items()} for _ in range(2)]
copy_stream = torch.cuda.Stream(device=dev)
sim = None
res = None
n_up = 0
mode = sys.argv[1] if len(sys.argv) > 1 else "pipe"

def ev():
    return torch.cuda.Event(enable_timing=True)

def upload():
    global n_up
    c = dev_cols[n_up % 2]; n_up += 1
    with torch.cuda.stream(copy_stream):
        a, b = ev(), ev()
        a.record(copy_stream)
        for k, v in host.items():
            c[k].copy_(v, non_blocking=True)
        b.record(copy_stream)
    return c, (a, b)

def step(cur, prefetch):
    global sim, res
    t0 = time.perf_counter()
    c, (a, b) = cur
    torch.cuda.current_stream().wait_event(b)
    e = [ev() for _ in range(4)]
    e[0].record()
    lg = ob.BucketedLog(z=c["z"], z_next=c["z_next"], action=c["action"], reward=c["reward"], terminal=c["terminal"],
                        p_log=c["p_log"], steps=c["steps"], n_buckets=nS, z_offset=0, device=dev, hot_records=True)
    t1 = time.perf_counter()
    e[1].record()
    if sim is None:
        sim = ob.BatchedPSRS(lg, R, rng="philox", seed=1, ring=True, ring_seed=9, lanes=2, aos_state=True)
    else:
        sim.restart(lg)
    res = sim.run_qlearn(nS, 511, 0.99, alpha=0.1, epsilon=0.1, max_steps=T, ep_cap=512, outputs=res)
    e[2].record()
    t2 = time.perf_counter()
    nxt = upload() if prefetch else None
    if mode == "serial" and nxt is not None:
        copy_stream.synchronize()
    t3 = time.perf_counter()
    s, s2, cnt = sim.learning_curve(res, 511)
    outs = [s.cpu(), s2.cpu(), cnt.cpu(), sim.accepted.sum().cpu()]
    e[3].record()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print("host ms: ctor %.1f launch %.1f upload-enqueue %.1f tail %.1f total %.1f | dev ms: bucket %.1f run %.1f curve %.1f copy %.1f" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t4 - t0),
        e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), a.elapsed_time(b)))
    return nxt

cur = upload()
for i in range(5):
    cur = step(cur, True)
