#!/usr/bin/env python
"""bench.py — PSRS offline-learner simulation throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c5|c2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one complete simulation job over the bucketed log: fresh replicas (cursors, Q, counters
zeroed), the fused PSRS + epsilon-greedy Q-learning run for `--budget` accepted transitions per replica,
and the learning-curve reduction (plus its NCCL all-reduce when N > 1).  Replicas shard across GPUs
(weak scaling: `--replicas` per GPU), the bucketed log is replicated per GPU, no data-path collective.

Printed JSON (one line, rank 0): metric = accepted transitions/s over all GPUs, `roofline` for the
fused run kernel (algorithmic bytes = pops x record bytes), `e2e` (host log -> H2D -> encode-free
bucketing -> run -> curve -> D2H every step), `cpu_baseline` (the C restatement of the reference loop on
this box's host cores, bounded sample), clocks, launches.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[3] sharded weak: 100M-transition / 10k-state / 4-action log, 1M replicas over 8 GPUs
    "c4": dict(rows=100_000_000, states=10_000, actions=4, replicas=125_000, budget=10_000, delta=None, lanes=1,
               desc="synthetic 100M-transition log, 10k states, 4 actions, eps-greedy Q-learning, 125k replicas/GPU (1M over 8)"),
    # configs[4]: rejection-heavy stress, near-deterministic behaviour vs uniform target, 256k replicas over 8 GPUs
    "c5": dict(rows=100_000_000, states=1_000, actions=4, replicas=32_000, budget=100, delta=1e-3, lanes=32,
               desc="rejection-heavy: 100M-transition log, 1000 states, p_b=(1-d,d/3,d/3,d/3) d=1e-3, uniform target, 32k replicas/GPU"),
    # configs[1] shape: cartpole-box sized state space, 10k replicas on one GPU
    "c2": dict(rows=1_000_000, states=162, actions=2, replicas=10_000, budget=2_000, delta=None, lanes=8,
               desc="cartpole-shaped: 1M-transition log, 162 states, 2 actions, eps-greedy Q-learning, 10k replicas"),
}
# DRAM bytes (read + write) of the dominant kernels, measured by `ncu --set full` on the build that is benchmarked and kept
# in profiles/ncu_traffic.json (bytes per accepted transition for the c4 learner, per pop for the c5 evaluation, with the
# capture each number comes from).  Used to report `roofline.traffic` per launch.
def load_traffic():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return {}


EPS, ALPHA, GAMMA = 0.1, 0.1, 0.99
EP_CAP = 512


def behaviour_table(n_states, n_actions, delta, seed=0):
    rng = np.random.default_rng(seed)
    if delta is None:
        pb = rng.dirichlet(np.ones(n_actions), size=n_states)
        pb = np.maximum(pb, 0.05)
        pb /= pb.sum(1, keepdims=True)
    else:  # near-deterministic: (1-d, d/3, ...) permuted per state (a strictly deterministic p_b accepts nothing)
        pb = np.full((n_states, n_actions), delta / (n_actions - 1))
        pb[np.arange(n_states), rng.integers(0, n_actions, n_states)] = 1.0 - delta
    return pb.astype(np.float32)


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    """One `nvidia-smi -lms` process for the whole run.  It is started BEFORE the warm-up (its start-up attaches to the
    driver and can stall launches for tens of ms) and only the samples that arrive between begin() and stop() count."""

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx, self.t_begin = [], None, gpu_index, 0.0

    def begin(self):
        self.t_begin = time.monotonic()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t >= self.t_begin]
        self.rows = rows if rows else [r for t, r in self.rows[-2:]]
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- CPU legs
def numpy_log(rows, states, actions, delta, seed=0, p_done=1 / 50, chunk=10_000_000):
    """The synthetic log of the benchmark shape generated on the HOST (the reference arm touches no GPU), in chunks so that
    the 1e8-row log needs no multi-GB temporaries."""
    pb = behaviour_table(states, actions, delta)
    cum = np.cumsum(pb, 1)
    out = dict(z=np.empty(rows, np.int32), z_next=np.empty(rows, np.int32), action=np.empty(rows, np.int32),
               reward=np.empty(rows, np.float32), terminal=np.empty(rows, np.uint8), p_log=np.empty((rows, actions), np.float32),
               steps=np.empty(rows, np.int32))
    prev_done = True
    for k, lo in enumerate(range(0, rows, chunk)):
        n = min(chunk, rows - lo)
        rng = np.random.default_rng([seed + 1, k])
        z = rng.integers(0, states, n).astype(np.int32)
        sl = slice(lo, lo + n)
        out["z"][sl] = z
        out["z_next"][sl] = rng.integers(0, states, n)
        out["action"][sl] = (rng.random(n)[:, None] > cum[z]).sum(1).clip(0, actions - 1)
        done = rng.random(n) < p_done
        out["reward"][sl] = rng.normal(size=n)
        out["terminal"][sl] = done
        out["p_log"][sl] = pb[z]
        out["steps"][sl] = np.where(np.r_[prev_done, done[:-1]], 0, 1)
        prev_done = bool(done[-1]) if n else prev_done
    return out


def cpu_leg(cols, wl, seconds, threads, log=None):
    """Time the C restatement of the reference's loop (oracle/ols_oracle.c) on `threads` host threads.  Returns the
    measurement and the bucketed host log (so that repeated legs bucket once; bucketing is untimed)."""
    from oracle import c_oracle as CO
    if log is None:
        log = CO.Log(z=cols["z"], z_next=cols["z_next"], action=cols["action"], reward=cols["reward"], terminal=cols["terminal"],
                     p_log=cols["p_log"], steps=cols["steps"], n_buckets=wl["states"])
    R = threads * 4
    budget = min(wl["budget"], 10_000)
    uniform = np.full((wl["states"], wl["actions"]), 1.0 / wl["actions"])
    total, t0, calls = 0, time.perf_counter(), 0
    pops = 0
    while True:
        if wl["delta"] is None:
            out = CO.run(log, R, CO.QLEARN, wl["states"], EP_CAP - 1, GAMMA, alpha=ALPHA, epsilon=EPS, seed=calls, max_steps=budget,
                         ring=True, ring_seed=calls, threads=threads, ep_cap=EP_CAP)
        else:
            out = CO.run(log, R, CO.FIXED_PI, wl["states"], EP_CAP - 1, GAMMA, pi=uniform, seed=calls, max_steps=budget, ring=True,
                         ring_seed=calls, threads=threads, ep_cap=EP_CAP)
        total += int(out["accepted"].sum())
        pops += int(out["pops"].sum())
        calls += 1
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
    return dict(value=total / dt, pops_per_s=pops / dt, seconds=dt, accepted=total, threads=out["threads"], replicas=R * calls,
                budget=budget, rows=log.n_rows), log


REF_DIR = os.path.join(ROOT, "baseline", "_ref", "offsim4rl")


def load_reference():
    """The UNMODIFIED reference's evaluators/psrs.py and agents/tabular.py from baseline/_ref (installed by
    __graft_entry__.build(); they import numpy / copy / itertools / tqdm only), loaded by file path because the package's
    __init__ imports gym.  None when the install is absent."""
    import importlib.util
    mods = []
    for rel in ("evaluators/psrs.py", "agents/tabular.py"):
        path = os.path.join(REF_DIR, rel)
        if not os.path.exists(path):
            return None
        spec = importlib.util.spec_from_file_location("offsim4rl_ref_" + os.path.basename(rel)[:-3], path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


def reference_python_leg(cols, wl, seconds, seed=0, rows=1_000_000, max_runs=6):
    """The reference's OWN loop, unmodified: `PSRS` (psrs.py:5-57) driven by `qlearn_psrs` (psrs.py:119-185) with
    `epsilon_greedy_policy` (tabular.py:24-32) — or `evalMC_psrs` (psrs.py:241-271) for the fixed-policy workload — on one
    core, over the first `rows` rows of the same synthetic log (the reference holds ~0.75 KB of python objects per row).
    Building the env is untimed; a run ends when a needed queue is exhausted (psrs.py:44-45) or at the deadline, which is
    injected through the env's public `step` attribute (an exhausted-queue answer), not by touching reference code."""
    ref = load_reference()
    if ref is None:
        return None
    ref_psrs, ref_tabular = ref
    n = min(len(cols["z"]), rows)
    nS, nA = wl["states"], wl["actions"]
    z, zn = cols["z"][:n].tolist(), cols["z_next"][:n].tolist()
    p_rows = list(cols["p_log"][:n])
    buf = [(zi, ai, ri, z2, di, pi, {"z": zi, "z_next": z2, "t": ti})
           for zi, ai, ri, z2, di, pi, ti in zip(z, cols["action"][:n].tolist(), list(cols["reward"][:n]), zn,
                                                 cols["terminal"][:n].astype(bool).tolist(), p_rows, cols["steps"][:n].tolist())]
    env = ref_psrs.PSRS(buf, nS=nS, nA=nA)
    timed, steps, runs = 0.0, 0, 0
    while timed < seconds and runs < max_runs:     # every run re-buckets and re-shuffles the log (reset_sampler, ~1.5 s, untimed)
        env.reset_sampler(seed=seed * 1000 + runs)
        deadline = time.perf_counter() + (seconds - timed)
        orig = type(env).step.__get__(env)
        count = [0]

        def step(p, _orig=orig, _count=count):
            if time.perf_counter() > deadline:
                return None, None, None, None
            out = _orig(p)
            _count[0] += out[0] is not None
            return out
        env.step = step
        np.random.seed(seed * 1000 + runs)
        t0 = time.perf_counter()
        if wl["delta"] is None:
            ref_psrs.qlearn_psrs(env, n_episodes=10**9, behavior_policy=ref_tabular.epsilon_greedy_policy, gamma=GAMMA, alpha=ALPHA,
                                 epsilon=EPS)
        else:
            ref_psrs.evalMC_psrs(env, n_episodes=10**9, pi=np.full((nS, nA), 1.0 / nA), gamma=GAMMA)
        timed += time.perf_counter() - t0
        steps += count[0]
        runs += 1
        del env.step
    return dict(value=steps / max(timed, 1e-9), seconds=timed, rows=n, runs=runs, accepted=steps)


def reference_python_multiprocess(cols, wl, seconds, procs, workload, rows=1_000_000):
    """The reference has no multi-process mode for PSRS (MPI is only used by its PPO agent); the closest equivalent is P
    independent single-replica runs in P separate processes, rates summed.  Plain subprocesses
    with a hard timeout: no fork of the CUDA-initialised parent, no pool."""
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "log.npz")
        np.savez(path, **{k: v[:rows] for k, v in cols.items()})
        cmd = [sys.executable, os.path.abspath(__file__), "--ref-worker", path, "--py-seconds", str(seconds), "--workload", workload]
        ps = [subprocess.Popen(cmd + ["--ref-seed", str(i + 1)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
              for i in range(procs)]
        vals = []
        deadline = time.time() + seconds * 6 + 240
        for pr in ps:
            try:
                out, _ = pr.communicate(timeout=max(1.0, deadline - time.time()))
                vals.append(json.loads(out.strip().splitlines()[-1])["value"])
            except Exception:
                pr.kill()
    if len(vals) != procs:
        return None
    return dict(value=float(sum(vals)), processes=procs, seconds=seconds, rows=rows)


def python_port_leg(cols, wl, seconds):
    """The pure-python restatement (interpreter-bound like the reference itself), one core.  A PSRS run ends when
    any needed queue runs dry (psrs.py:44-45), which on the truncated log happens after a few hundred steps, so
    fresh environments are built (untimed) and only the stepping is timed."""
    from oracle import psrs_oracle as O
    n = min(len(cols["z"]), 1_000_000)
    rng = np.random.default_rng(0)

    class Ties:
        def pick(self, cand):
            return int(cand[rng.integers(0, len(cand))]) if len(cand) > 1 else int(cand[0])
    timed, steps, runs = 0.0, 0, 0
    while timed < seconds and runs < 200:
        env = O.OraclePSRS(z=cols["z"][:n], z_next=cols["z_next"][:n], s=cols["z"][:n], s_next=cols["z_next"][:n],
                           action=cols["action"][:n], reward=cols["reward"][:n], terminal=cols["terminal"][:n], p_log=cols["p_log"][:n],
                           steps=cols["steps"][:n], n_buckets=wl["states"], uniforms=O.Pcg64Uniforms(runs), init_start=runs * 997)
        t0 = time.perf_counter()
        _, info = O.qlearn(env, 10_000, GAMMA, ALPHA, EPS, Ties(), wl["states"], wl["actions"], record=False)
        timed += time.perf_counter() - t0
        steps += int(info["lengths"].sum())
        runs += 1
    return dict(value=steps / max(timed, 1e-9), seconds=timed, rows=n, runs=runs)


def python_port_multiprocess(cols, wl, seconds, procs, workload):
    """The reference has no multi-process mode for PSRS (MPI is only used by its PPO agent); the closest equivalent is P
    independent single-replica runs in P separate processes, rates summed.  Plain subprocesses
    with a hard timeout: no fork of the CUDA-initialised parent, no pool."""
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "log.npz")
        np.savez(path, **{k: v[:1_000_000] for k, v in cols.items()})
        cmd = [sys.executable, os.path.abspath(__file__), "--py-worker", path, "--py-seconds", str(seconds), "--workload", workload]
        ps = [subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for _ in range(procs)]
        vals = []
        deadline = time.time() + seconds * 6 + 90
        for pr in ps:
            try:
                out, _ = pr.communicate(timeout=max(1.0, deadline - time.time()))
                vals.append(json.loads(out.strip().splitlines()[-1])["value"])
            except Exception:
                pr.kill()
    if len(vals) != procs:
        return None
    return dict(value=float(sum(vals)), processes=procs, seconds=seconds)


def run_reference(args, wl, rank, world):
    """--impl reference: the CPU implementation of the path on this box's host cores, same metric and config as the GPU
    arm; rank 0 only.  `value` = oracle/ols_oracle.c (the C restatement of the reference's loop; the reference itself is
    python and cannot hold the 1e8-row log) on ALL host threads over the FULL synthetic log of the workload; beside it the
    UNMODIFIED reference (baseline/_ref: `PSRS` + `qlearn_psrs` / `evalMC_psrs`) timed on 1 process and on P processes over
    the first 1e6 rows."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rows = wl["rows"]
    t0 = time.perf_counter()
    cols = numpy_log(rows, wl["states"], wl["actions"], wl["delta"])
    gen_s = time.perf_counter() - t0
    per_step = 4.0
    vals = []
    clog = None
    for i in range(args.warmup + args.steps):
        leg, clog = cpu_leg(cols, wl, per_step if i >= args.warmup else 1.0, threads, log=clog)
        if i >= args.warmup:
            vals.append(leg)
    acc = sum(v["accepted"] for v in vals)
    secs = sum(v["seconds"] for v in vals)
    value = acc / secs
    del clog
    ref1 = refp = None
    if not args.skip_cpu:
        ref1 = reference_python_leg(cols, wl, 6.0)
        refp = reference_python_multiprocess(cols, wl, 6.0, threads, args.workload) if ref1 is not None else None
    sample = (f"oracle/ols_oracle.c (C restatement of psrs.py PSRS.step + qlearn/evalMC loop), {vals[-1]['threads']} threads, the full "
              f"{rows}-row synthetic log (host-generated in {gen_s:.0f} s, untimed), {vals[-1]['budget']} accepted steps per replica, "
              f"~{per_step:.0f} s per step")
    print(json.dumps({
        "impl": "reference", "metric": "simulated accepted transitions/sec", "value": value, "unit": "transitions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, len(vals)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload + ": " + wl["desc"], "rows": rows, "states": wl["states"], "actions": wl["actions"]},
        "cpu_baseline": {"value": value, "unit": "transitions/s", "cores": vals[-1]["threads"], "kind": "port", "sample": sample,
                         "pops_per_s": sum(v["pops_per_s"] * v["seconds"] for v in vals) / secs,
                         "reference_python": None if ref1 is None else {
                             "kind": "reference", "what": "UNMODIFIED offsim4rl evaluators/psrs.py + agents/tabular.py from baseline/_ref, "
                                                          "first 1e6 rows of the same log, env construction untimed",
                             "one_process": ref1, "processes": refp}},
        "e2e": {"value": value, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def gather_stress_leg(ob, torch, dev, hbm_peak, launches=5, warm=3, hot="tail"):
    """Second half of the BASELINE metric ("PSRS gather HBM GB/s vs peak"): the rejection-heavy configuration
    (BASELINE.json configs[4], workload c5) is the streaming-gather regime.  Rank 0, one GPU, a few launches."""
    wl = WORKLOADS["c5"]
    N, nS, nA, R, T = wl["rows"], wl["states"], wl["actions"], wl["replicas"], wl["budget"]
    cols = ob.synth_columns(N, nS, behaviour_table(nS, nA, wl["delta"]), seed=0, device=dev)
    log = ob.BucketedLog(z=cols["z"], z_next=cols["z_next"], action=cols["action"], reward=cols["reward"],
                         terminal=cols["terminal"], p_log=cols["p_log"], steps=cols["steps"], n_buckets=nS, z_offset=0, device=dev,
                         hot_records=hot)
    del cols
    sim = ob.BatchedPSRS(log, R, rng="philox", seed=4321, ring=True, ring_seed=7, lanes=wl["lanes"])
    pi = torch.full((nS, nA), 1.0 / nA, dtype=torch.float64, device=dev)
    res, ms, pops = None, [], 0.0
    for i in range(launches + warm):
        sim.restart()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = sim.run_evalmc(nS, EP_CAP - 1, pi, GAMMA, max_steps=T, ep_cap=EP_CAP, outputs=res)
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:               # the first launches are warm-up
            ms.append(e0.elapsed_time(e1))
            pops = float(sim.pops.sum())
    t = sum(ms) / len(ms) / 1e3
    gbs = pops * log.rec_bytes / t / 1e9
    read_per_pop = 8 if hot else log.rec_bytes
    kind = {False: "2 full records of 32 bytes", True: "4 PROBS hot records of 8 bytes", "probs": "4 PROBS hot records of 8 bytes",
            "tail": "8 TAIL hot records of 8 bytes (two 32-byte sectors)"}[hot]
    tr = load_traffic().get("c5_" + ("tail" if hot == "tail" else "probs" if hot else "full"), {})
    return {"workload": "c5: " + wl["desc"],
            "kernel": ("eval_kernel (streaming fixed-policy evaluation, one warp per replica x %s per lane and round)" if hot == "tail"
                       else "run_kernel (fixed-policy evaluation, lanes=32 x %s per round)") % kind,
            "bytes_read_per_rejected_pop": read_per_pop, "pops_per_s": pops / t,
            "accepted_per_s": float(sim.accepted.sum()) / t, "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
            "frac_note": "contract bytes (SURVEY 8d: 32 B per popped record) / time / measured HBM copy peak; the kernel itself reads "
                         "%d B per popped record: see dram_achieved / dram_frac" % read_per_pop,
            "dram_achieved": pops * read_per_pop / t / 1e9, "dram_frac": pops * read_per_pop / t / 1e9 / hbm_peak,
            "kernel_ms": 1e3 * t, "launches": launches,
            "traffic": (tr["dram_bytes_per_pop"] * pops) if tr else None, "traffic_source": tr.get("capture"),
            "algorithmic_bytes_per_launch": pops * log.rec_bytes}


def queue_evaluator_leg(ob, torch, dev, hbm_peak, steps=200, launches=5, warm=2):
    """the QueueEvaluator (queues keyed by (z, a), the agent's own action picks the queue, no rejection;
    queue_evaluator.py:113-131) as a FUSED run — `ols_run_actions`: 1M replicas, each resets and steps `steps` times per
    launch with actions sampled from a uniform policy by its own counter-based stream.  Rank 0, one GPU, a second or two."""
    N, nS, nA, R = 20_000_000, 1000, 4, 1_000_000
    cols = ob.synth_columns(N, nS, behaviour_table(nS, nA, None), seed=3, device=dev)
    log = ob.BucketedLog(z=cols["z"], z_next=cols["z_next"], action=cols["action"], reward=cols["reward"], terminal=cols["terminal"],
                         p_log=cols["p_log"], steps=cols["steps"], n_buckets=nS, z_offset=0, device=dev, key_by_action=True)
    del cols
    sim = ob.BatchedPSRS(log, R, rng="philox", seed=5, ring=True, ring_seed=11, lanes=1, reject="accept_all")
    pi = torch.full((nS, nA), 1.0 / nA, dtype=torch.float64, device=dev)
    res, ms, n = None, [], 0.0
    for i in range(launches + warm):
        sim.restart()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = sim.run_actions(EP_CAP - 1, GAMMA, pi=pi, max_steps=steps, ep_cap=EP_CAP, outputs=res)
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ms.append(e0.elapsed_time(e1))
            n = float(sim.accepted.sum())
    t = sum(ms) / len(ms) / 1e3
    per = log.rec_bytes + 8                      # the record + the replica's cursor (read, write)
    return {"row": "QueueEvaluator reset/step over (z, a)-keyed queues (queue_evaluator.py:113-131), fused multi-step run (ols_run_actions)",
            "rows": N, "states": nS, "actions": nA, "replicas": R, "steps_per_launch": steps, "value": n / t, "unit": "transitions/s",
            "launches": launches, "kernel_ms": 1e3 * t, "achieved": n * per / t / 1e9, "peak": hbm_peak, "roofline_unit": "GB/s",
            "frac": n * per / t / 1e9 / hbm_peak, "algorithmic_bytes_per_transition": per,
            "note": "one launch per %d simulated steps of every replica; one random 32-byte record read and one random cursor "
                    "read-modify-write per transition (isolated DRAM accesses, not streaming)" % steps}


# ------------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int)
    ap.add_argument("--states", type=int)
    ap.add_argument("--replicas", type=int, help="replicas per GPU")
    ap.add_argument("--budget", type=int, help="accepted transitions per replica per step")
    ap.add_argument("--lanes", type=int)
    ap.add_argument("--prefetch", type=int, default=-1)
    ap.add_argument("--ppl", type=int, default=0, help="records per lane per round (0 = auto)")
    ap.add_argument("--slices", type=int, default=-1, help="time slices of the fused run: -1 auto, 0 off, k forced")
    ap.add_argument("--hot", choices=["probs", "tail"], help="hot record format (default: tail for the one-lane Q-learner and for evaluation, else probs)")
    ap.add_argument("--batch-lanes", type=int, default=0, help="one-lane learner: lanes per update batch (0 = library default)")
    ap.add_argument("--no-hot", action="store_true", help="pop from the 32-byte records only (no 8-byte hot records)")
    ap.add_argument("--no-aos", action="store_true", help="separate cursor / Q arrays instead of one 32-byte entry per (replica, state)")
    ap.add_argument("--no-ring", action="store_true", help="all replicas consume the queues from slot 0 (L2-resident gather)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-stress", action="store_true", help="skip the c5 streaming-gather measurement appended to a c4 run")
    ap.add_argument("--e2e-steps", type=int, default=0, help="end-to-end steps (0 = --steps)")
    ap.add_argument("--py-worker", help=argparse.SUPPRESS)      # internal: one process of the multi-process python baseline
    ap.add_argument("--py-seconds", type=float, default=4.0, help=argparse.SUPPRESS)
    ap.add_argument("--ref-worker", help=argparse.SUPPRESS)     # internal: one process of the multi-process reference baseline
    ap.add_argument("--ref-seed", type=int, default=1, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.py_worker:
        cols = dict(np.load(args.py_worker))
        print(json.dumps(python_port_leg(cols, dict(WORKLOADS[args.workload]), args.py_seconds)))
        return
    if args.ref_worker:
        cols = dict(np.load(args.ref_worker))
        print(json.dumps(reference_python_leg(cols, dict(WORKLOADS[args.workload]), args.py_seconds, seed=args.ref_seed)))
        return
    if args.e2e_steps <= 0:
        args.e2e_steps = args.steps
    wl = dict(WORKLOADS[args.workload])
    for k in ("rows", "states", "replicas", "budget", "lanes"):
        if getattr(args, k) is not None:
            wl[k] = getattr(args, k)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist
    import offsim4rl_b200 as ob

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

    N, nS, nA, R, T = wl["rows"], wl["states"], wl["actions"], wl["replicas"], wl["budget"]
    pb = behaviour_table(nS, nA, wl["delta"])
    qlearn = wl["delta"] is None
    uniform_pi = torch.full((nS, nA), 1.0 / nA, dtype=torch.float64, device=dev)   # resident: no per-step upload of the policy

    # ---- setup (untimed): synthetic log generated on the device, bucketed once
    cols = ob.synth_columns(N, nS, pb, seed=0, device=dev)
    torch.cuda.synchronize()
    # learners: hot records need the AoS state and give nothing when the state fits L2 (c2); fixed-policy evaluation (c5): always
    use_hot = nA <= 4 and not args.no_hot and ((qlearn and not args.no_aos and args.workload == "c4") or not qlearn)
    # the one-thread-per-replica Q-learner reads TAIL hot records (reward, next state, action, terminal: the synthetic
    # behaviour distribution is a function of the state, so the accept rule needs only the action); evaluation PROBS ones
    hot_kind = (args.hot or ("tail" if (not qlearn or wl["lanes"] == 1) else "probs")) if use_hot else False
    log = ob.BucketedLog(z=cols["z"], z_next=cols["z_next"], action=cols["action"], reward=cols["reward"],
                         terminal=cols["terminal"], p_log=cols["p_log"], steps=cols["steps"], n_buckets=nS, z_offset=0, device=dev,
                         keep_columns=True, hot_records=hot_kind)
    from offsim4rl_b200 import _abi as _abi_mod
    hot_in_effect = None if log.hot is None else ("tail" if log.hot_kind == _abi_mod.HOT_TAIL else "probs")
    default_shape = args.ppl == 0 and args.lanes is None and args.batch_lanes == 0
    bucket_ms = log.time_bucketing(3)       # ols_bucket + ols_init_queue, CUDA events on the launching stream
    log.drop_columns()
    rec_bytes = log.rec_bytes
    bucket_bytes = N * (12 + 17 + 4 * nA + rec_bytes)   # z twice + order, log columns once, record written once
    sim = ob.BatchedPSRS(log, R, rng="philox", seed=1234, ring=not args.no_ring, ring_seed=99, replica_id0=rank * R, lanes=wl["lanes"], prefetch=args.prefetch, pops_per_lane=args.ppl,
                         aos_state=(qlearn and nA <= 4 and not args.no_aos), slices=args.slices, batch_lanes=args.batch_lanes)
    use_aos = qlearn and nA <= 4 and not args.no_aos
    Q = torch.zeros((R, nS, nA), dtype=torch.float32, device=dev) if (qlearn and not use_aos) else None
    res = None

    k_ev = {"run": []}

    def one_step(timed):
        nonlocal res
        ea = torch.cuda.Event(enable_timing=True); ez = torch.cuda.Event(enable_timing=True)
        ea.record()
        sim.restart()
        if Q is not None:
            Q.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        if qlearn:
            res = sim.run_qlearn(nS, EP_CAP - 1, GAMMA, alpha=ALPHA, epsilon=EPS, Q=Q, max_steps=T, ep_cap=EP_CAP, outputs=res)
        else:
            res = sim.run_evalmc(nS, EP_CAP - 1, uniform_pi, GAMMA, max_steps=T, ep_cap=EP_CAP, outputs=res)
        e1.record()
        curve = sim.learning_curve(res, EP_CAP - 1)
        ez.record()
        if timed:
            k_ev["run"].append((e0, e1))
            k_ev.setdefault("restart", []).append((ea, e0))
            k_ev.setdefault("reduce", []).append((e1, ez))
        return curve

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("OLS_BENCH_NO_SAMPLER"):     # (the switch exists to measure what the sampler itself costs)
        sampler.start()
    acc_total = torch.zeros(2, dtype=torch.float64, device=dev)
    for _ in range(args.warmup):
        one_step(False)
        acc_total[0] += sim.accepted.sum()      # every kernel of the timed loop runs once before it (CUDA loads modules lazily:
        acc_total[1] += sim.pops.sum()          # the first launch of a kernel costs 10-40 ms and would land in the first timed step)
    barrier()
    sampler.begin()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    acc_total.zero_()
    launches0 = int(ob._abi.load().ols_kernel_launches())
    t0.record()
    for _ in range(args.steps):
        one_step(True)
        acc_total[0] += sim.accepted.sum()
        acc_total[1] += sim.pops.sum()
    t1.record()
    launches_timed = int(ob._abi.load().ols_kernel_launches()) - launches0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    run_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in k_ev["run"]) / max(1, len(k_ev["run"]))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(acc_total, op=dist.ReduceOp.SUM)
    accepted, pops = float(acc_total[0]), float(acc_total[1])
    secs = float(elapsed_ms) / 1e3
    value = accepted / secs
    # roofline of the dominant kernel (fused run kernel) on this rank
    my_pops = float(sim.pops.sum())
    last_accepted = float(sim.accepted.sum())
    run_s = float(run_ms) / 1e3
    gather_gbs = my_pops * rec_bytes / run_s / 1e9
    agent_gbs = last_accepted * (4 * nA + 4 + 8 if qlearn else 8) / run_s / 1e9

    # ---- e2e: host-resident log -> H2D -> bucket -> run -> curve -> D2H, every step
    e2e = None
    if not args.skip_e2e:
        # The job's input is ONE host-resident log.  With N ranks each rank keeps 1/N of the rows in pinned host memory,
        # uploads that shard over its own PCIe link and the shards are all-gathered over NVLink (parallel.all_gather_columns):
        # the host side moves 3.7 GB per step in total instead of 3.7 GB per rank.
        from offsim4rl_b200 import parallel as par
        sharded = world > 1 and N % world == 0
        lo_r, hi_r = par.row_shard(N, rank, world) if sharded else (0, N)
        host = {k: torch.empty(v[lo_r:hi_r].shape, dtype=v.dtype, pin_memory=True) for k, v in cols.items()}
        for k, v in cols.items():
            host[k].copy_(v[lo_r:hi_r])
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        full_host = None
        if rank == 0 and not args.skip_cpu and world == 1:
            full_host = host                                           # the CPU baseline below reads the same host copy
        col_meta = {k: (v.shape, v.dtype) for k, v in cols.items()}
        del cols, log
        sim.log = None
        torch.cuda.synchronize()
        e_acc, d2h = 0.0, 0

        copy_stream = torch.cuda.Stream(device=dev)
        dev_cols = [{k: torch.empty(shp, dtype=dt, device=dev) for k, (shp, dt) in col_meta.items()} for _ in range(2)]
        dev_shard = [{k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in host.items()} for _ in range(2)] if sharded else None
        n_up = 0

        def upload():
            """H2D of one step's log columns (this rank's row shard) from pinned host memory, then the all-gather of the
            shards, into one of two resident column sets, on the copy stream.  (A set is reused two steps later; every step
            ends with a blocking D2H read, so it is free by then.)"""
            nonlocal n_up
            c = dev_cols[n_up % 2]
            with torch.cuda.stream(copy_stream):
                if sharded:
                    sh = dev_shard[n_up % 2]
                    for k, v in host.items():
                        sh[k].copy_(v, non_blocking=True)
                    par.all_gather_columns(sh, c)
                else:
                    for k, v in host.items():
                        c[k].copy_(v, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            n_up += 1
            return c, ev

        def e2e_step(cur, prefetch_next):
            nonlocal res, d2h
            c, ev = cur
            torch.cuda.current_stream().wait_event(ev)
            lg = ob.BucketedLog(z=c["z"], z_next=c["z_next"], action=c["action"], reward=c["reward"], terminal=c["terminal"],
                                p_log=c["p_log"], steps=c["steps"], n_buckets=nS, z_offset=0, device=dev, hot_records=hot_kind)
            sim.restart(lg)
            if Q is not None:
                Q.zero_()
            if qlearn:
                res = sim.run_qlearn(nS, EP_CAP - 1, GAMMA, alpha=ALPHA, epsilon=EPS, Q=Q, max_steps=T, ep_cap=EP_CAP, outputs=res)
            else:
                res = sim.run_evalmc(nS, EP_CAP - 1, uniform_pi, GAMMA, max_steps=T, ep_cap=EP_CAP, outputs=res)
            # double-buffered input: the NEXT step's columns are uploaded by the copy engine while this step's run executes
            nxt = upload() if prefetch_next else None
            s, s2, cnt = sim.learning_curve(res, EP_CAP - 1)
            outs = [s.cpu(), s2.cpu(), cnt.cpu(), sim.accepted.sum().cpu()]
            d2h = sum(o.numel() * o.element_size() for o in outs)
            return float(outs[3]), nxt

        e2e_step(upload(), False)  # warm the allocator
        e2e_step(upload(), False)
        barrier()
        w0 = time.perf_counter()
        cur = upload()             # every timed step's upload happens inside the timed region
        for i in range(args.e2e_steps):
            acc_i, cur = e2e_step(cur, i + 1 < args.e2e_steps)
            e_acc += acc_i
        barrier()
        e_ms = torch.tensor([(time.perf_counter() - w0) * 1e3], dtype=torch.float64, device=dev)
        e_tot = torch.tensor([e_acc], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(e_tot, op=dist.ReduceOp.SUM)
        e2e = {"value": float(e_tot) / (float(e_ms) / 1e3), "unit": "transitions/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": float(e_ms) / args.e2e_steps, "steps": args.e2e_steps,
               "h2d_note": ("per rank: its 1/%d row shard of the 7 log columns; the shards are all-gathered over NVLink (NCCL), so the "
                            "host moves the %.1f GB log once per step in total" % (world, h2d * world / 1e9)) if sharded
                           else "per rank: the 7 log columns of the whole log",
               "d2h_note": "the learning curve (sum, sumsq, count per episode) and the accepted count; Q tables stay on the device",
               "pipeline": "per step: H2D of the 7 log columns (pinned host, copy stream) -> bucketing -> fused run -> curve reduce -> "
                           "D2H of the curve; the H2D of step i+1 is issued behind the run launch of step i (double-buffered "
                           "columns); all %d uploads are inside the timed region" % args.e2e_steps}
        cpu_cols = None if full_host is None else {k: v.numpy() for k, v in full_host.items()}   # the full log, already on the host
    else:
        cpu_cols = {k: v[:1_000_000].cpu().numpy() for k, v in cols.items()}

    stress = next_rows = None
    if rank == 0 and args.workload == "c4" and not args.skip_stress:
        sim = Q = res = None        # release the c4 state (25 GB) before building the c5 log
        torch.cuda.empty_cache()
        try:
            stress = gather_stress_leg(ob, torch, dev, hbm_peak, hot=False if args.no_hot else (args.hot or "tail"))
        except Exception as exc:   # a side leg must never take the headline line down
            stress = {"error": repr(exc)}
        torch.cuda.empty_cache()
        try:
            next_rows = {"queue_evaluator": queue_evaluator_leg(ob, torch, dev, hbm_peak)}
        except Exception as exc:
            next_rows = {"queue_evaluator": {"error": repr(exc)}}
    if rank == 0:
        cpu = None
        if not args.skip_cpu and world == 1 and cpu_cols is not None:
            threads = os.cpu_count() or 1
            leg, _clog = cpu_leg(cpu_cols, wl, args.cpu_seconds, threads)
            del _clog
            small = {k: v[:1_000_000] for k, v in cpu_cols.items()}
            ref1 = reference_python_leg(small, wl, 5.0)
            refp = reference_python_multiprocess(small, wl, 5.0, threads, args.workload) if ref1 is not None else None
            py = python_port_leg(small, wl, 3.0) if (qlearn and ref1 is None) else None
            cpu = {"value": leg["value"], "unit": "transitions/s", "cores": leg["threads"], "kind": "port",
                   "sample": (f"oracle/ols_oracle.c (C restatement of the reference's PSRS + learner loop), {leg['threads']} threads, the same "
                              f"synthetic log ({leg['rows']} rows), {leg['replicas']} replicas x {leg['budget']} accepted steps, "
                              f"{leg['seconds']:.1f} s"),
                   "pops_per_s": leg["pops_per_s"],
                   "reference_python": None if ref1 is None else {
                       "kind": "reference", "what": "UNMODIFIED offsim4rl evaluators/psrs.py + agents/tabular.py (baseline/_ref): PSRS + "
                                                    "qlearn_psrs / evalMC_psrs, first 1e6 rows of the same log, env construction untimed",
                       "one_process": ref1, "processes": refp},
                   "python_port_1core": None if py is None else {"value": py["value"], "rows": py["rows"], "seconds": py["seconds"]}}
        tail_on = hot_in_effect == "tail"
        traffic = load_traffic().get(("c4_" if qlearn else "c5_") + (hot_in_effect or "full"), {}) if default_shape else {}
        per_unit = traffic.get("dram_bytes_per_accepted" if qlearn else "dram_bytes_per_pop")
        units = last_accepted if qlearn else my_pops
        read_per_pop = 8 if hot_in_effect else rec_bytes
        roofline = {
            "bound": "hbm",
            "kernel": ("learner_kernel (one thread per replica: PSRS pops over TAIL hot records + epsilon-greedy Q-learning update, ols_learner.cu)"
                       if (qlearn and tail_on and wl["lanes"] == 1) else
                       "eval_kernel (one warp per replica: streaming PSRS pops over TAIL hot records + return evaluator, ols_eval.cu)"
                       if (not qlearn and tail_on and wl["lanes"] == 32) else "run_kernel (fused PSRS step + agent update / return evaluator)"),
            "achieved": gather_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gather_gbs / hbm_peak,
            "frac_note": "contract bytes (SURVEY 8d: 32 B per popped record) / kernel time / measured HBM copy peak",
            "traffic": per_unit * units if per_unit else None,
            "traffic_note": ("DRAM read+write bytes per launch = %s of the same build (%s) x this launch's count"
                             % ("ncu bytes per accepted transition" if qlearn else "ncu bytes per pop", traffic.get("capture"))) if per_unit else None,
            "algorithmic_bytes_per_launch": my_pops * rec_bytes, "peak_source": peak_src,
            "algorithmic_bytes_per_pop": rec_bytes, "bytes_read_per_pop_by_the_kernel": read_per_pop,
            "dram_achieved": (per_unit * units / run_s / 1e9) if per_unit else None,
            "dram_frac": (per_unit * units / run_s / 1e9 / hbm_peak) if per_unit else None,
            "agent_state_GBps": agent_gbs, "kernel_ms": float(run_ms),
            "dram_sectors_per_s": (traffic["dram_sectors_per_accepted"] * last_accepted / run_s)
                                  if (qlearn and traffic.get("dram_sectors_per_accepted")) else None}
        out = {
            "metric": "simulated accepted transitions/sec", "value": value, "unit": "transitions/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "dtype_note": "accept rule decided exactly as float64 (float32 pre-filter + exact fallback); Q table float32",
            "data": "synthetic",
            "config": {"workload": args.workload + ": " + wl["desc"], "rows": N, "states": nS, "actions": nA,
                       "replicas_per_gpu": R, "accepted_per_replica_per_step": T, "lanes": wl["lanes"],
                       "time_slices": ("auto (16 per replica block: the one-lane learners are sliced whenever the launch holds more CTAs than SMs, the others when their blocks are not a whole number of waves)" if args.slices == -1 else args.slices),
                       "hot_records": {"tail": "8-byte TAIL records {reward; next state | action | terminal}: the behaviour distribution is a function "
                                               "of the state (checked bitwise per queue on the device), so the accept rule needs only the "
                                               "record's action; per-step accept/reject limits decide all but ~6e-5 of the pops",
                                       "probs": "8-byte half-precision images of the records decide all but ~5e-4 of the pops; the 32-byte record "
                                                "is read for the accepted pop and for the undecided ones", None: "off"}[hot_in_effect],
                       "replica_state": ("one 32-byte entry {cursor, Q row} per (replica, state)" if use_aos else "separate cursor and Q arrays"),
                       "queue_order": "log order" + ("" if args.no_ring else " + per-(replica,state) ring offset"),
                       "l2": "working set (records %.1f GB + replica state %.1f GB) >> 126 MB L2; no flush" % (
                           N * rec_bytes / 1e9, (R * nS * 32 / 1e9) if use_aos else (R * nS * 4 / 1e9 + (R * nS * nA * 4 / 1e9 if qlearn else 0.0)))},
            "pops_per_s": pops / secs,
            "roofline": roofline,
            "bucketing": {"rows": N, "ms": bucket_ms, "GBps": bucket_bytes / bucket_ms / 1e6, "frac": bucket_bytes / bucket_ms / 1e6 / hbm_peak,
                          "note": "ols_bucket + ols_init_queue kernels; bytes = z twice + order + log columns + packed record per row"},
            "step_breakdown_ms": {k: sum(a.elapsed_time(b) for a, b in v) / max(1, len(v)) for k, v in k_ev.items()},
            "step_ms_each": [round(a[0].elapsed_time(b[0]), 3) for a, b in zip(k_ev["restart"], k_ev["restart"][1:] + [(t1, None)])],
            "gather_stress": stress, "next_rows": next_rows,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches_timed),
            "gpu_launches_note": "kernels of libols_b200.so launched inside the timed region, counted by the library (ols_kernel_launches)",
            "clocks": clocks,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
