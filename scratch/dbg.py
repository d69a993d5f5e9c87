import sys; sys.path.insert(0,'.')
import numpy as np, torch
import offsim4rl_b200 as ob
from oracle import c_oracle as CO
from tests.test_gpu_engine import make_log, device_log
g = make_log(21, 200000, 300, 4, ep_len=25)
R, nS = 3000, 300
log = device_log(ob, g, n_buckets=nS)
for lanes, ppl in [(4,2),(1,1)]:
    sim = ob.BatchedPSRS(log, R, rng="philox", seed=4242, ring=False, ring_seed=7, replica_id0=10**6, lanes=lanes, pops_per_lane=ppl)
    clog = CO.Log(z=g["z"], z_next=g["z_next"], action=g["a"], reward=g["r"], terminal=g["done"], p_log=g["p_log"], steps=g["t"], n_buckets=nS)
    res = sim.run_qlearn(nS, 30, gamma=0.95, alpha=0.2, epsilon=0.15, q_dtype=torch.float64, trace_cap=4000)
    ref = CO.run(clog, R, CO.QLEARN, nS, 30, gamma=0.95, alpha=0.2, epsilon=0.15, seed=4242, replica_id0=10**6, ring=False, ring_seed=7, trace_cap=4000, threads=8)
    acc = res.accepted.cpu().numpy()
    bad = np.nonzero(acc != ref["accepted"])[0]
    print(lanes, ppl, "bad replicas", bad.size, bad[:10])
    tr = res.trace_row.cpu().numpy()
    nb = 0
    for r in range(R):
        n = min(acc[r], ref["accepted"][r])
        d = np.nonzero(tr[r,:n] != ref["trace_row"][r,:n])[0]
        if d.size:
            nb += 1
            if nb <= 3:
                k = d[0]
                print(" replica", r, "first diff at step", k, "gpu row", tr[r,k], "ref row", ref["trace_row"][r,k], "ties", int(sim.ties[r]), ref["ties"][r], "draws", int(sim.draws[r]), ref["draws"][r])
    print(" replicas with trace diffs:", nb)
