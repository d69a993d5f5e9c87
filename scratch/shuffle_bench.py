"""Times BucketedLog.seeded(seed) (ols_shuffle_queues: the reference's reset_sampler(seed) permutation replay) at full size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import offsim4rl_b200 as ob
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
nS = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
dev = torch.device("cuda:0")
pb = np.random.default_rng(0).dirichlet(np.ones(4), size=nS)
cols = ob.synth_columns(N, nS, pb, seed=0, device=dev)
log = ob.BucketedLog(z=cols["z"], z_next=cols["z_next"], action=cols["action"], reward=cols["reward"], terminal=cols["terminal"],
                     p_log=cols["p_log"], steps=cols["steps"], n_buckets=nS, z_offset=0, device=dev, hot_records=True)
for seed in (1, 2, 3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s = log.seeded(seed)
    torch.cuda.synchronize(); print("seeded ms", N, nS, round(1e3 * (time.perf_counter() - t0), 2), "n_init", log.n_init)
    del s
