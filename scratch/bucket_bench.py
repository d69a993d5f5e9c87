"""Times ols_bucket + ols_init_queue on a synthetic log (scratch tool; A/B of library variants via OLS_B200_LIB)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import offsim4rl_b200 as ob

def main():
    N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    nS = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
    nA = 4
    pb = np.random.default_rng(0).dirichlet(np.ones(nA), size=nS)
    dev = torch.device("cuda:0")
    cols = ob.synth_columns(N, nS, pb, seed=0, device=dev)
    log = ob.BucketedLog(z=cols["z"], z_next=cols["z_next"], action=cols["action"], reward=cols["reward"],
                         terminal=cols["terminal"], p_log=cols["p_log"], steps=cols["steps"], n_buckets=nS, z_offset=0,
                         device=dev, keep_columns=True)
    log.time_bucketing(2)
    print("bucket_ms", N, nS, round(log.time_bucketing(5), 3), "order_checksum", int(log.order[:: max(N // 1000, 1)].sum()))

main()
