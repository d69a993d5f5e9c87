"""ptxas -v of a forced build -> profiles/r04_ptxas_registers.txt (registers, spills, static shared memory per kernel)."""
import re, subprocess, sys
txt = subprocess.run([sys.executable, "rl-offline-simulation_b200/build.py", "--force", "-v"], capture_output=True, text=True).stderr
rows, cur, stack = [], None, ("0", "0", "0")
for line in txt.split("\n"):
    m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur:
        stack = (m.group(1), m.group(2), m.group(3))
        continue
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        sm = re.search(r"(\d+) bytes smem", line)
        rows.append((cur, m.group(1), stack, sm.group(1) if sm else "0"))
        cur = None
dem = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.strip().split("\n")
out = ["# round 2 (final), shipped build (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo): ptxas -v per kernel",
       "# registers | stack frame / spill stores / spill loads (bytes) | static smem | kernel",
       "# occupancy follows from __launch_bounds__: learner_kernel 5 CTAs x 128 threads per SM at 96 registers; run_kernel<.., GT=1, HOT=1> 7 x 128",
       "# at 72; eval_kernel<2> 4 x 256 at 64 (its stack frame is spill space: ~6 LDL per 256-record round); the other run_kernel instantiations\n# 4 x 256 at 64 (<= 4 actions) or 2 x 256 (wider records)."]
want = ["learner_kernel", "eval_kernel", "queue_run", "exo_run", "scatter_pack_kernel<1>", "hot_pack", "queue_dist", "tile_histogram", "encode_mlp", "curve_",
        "run_kernel<1, 2, float, 2, 4, true, 1, 1>", "run_kernel<1, 0, float, 2, 4, false, 0, 2>", "run_kernel<1, 0, float, 2, 4, false, 0, 1>",
        "run_kernel<1, 2, double, 0, 2, false, 0, 0>", "step_kernel<1, 2>", "shuffle_perm", "exo_step_kernel<1, 1>", "encode_box", "encode_grid"]
n = 0
for (mang, regs, stack, sm), d in zip(rows, dem):
    d = re.sub(r"\(ols_queues.*|\(ols_log.*|\(const .*|\(unsigned.*|\(int.*", "", d).replace("(anonymous namespace)::", "").replace("ols_sim::", "")
    d = d.replace("(int)", "").replace("(bool)", "")
    n += 1
    if any(k in d for k in want):
        out.append(f"{regs:>4s} | {stack[0]:>4s} / {stack[1]:>4s} / {stack[2]:>4s} | {sm:>6s} | {d[:150]}")
out.append(f"# {n} kernels in the library; max registers {max(int(r[1]) for r in rows)}")
open("profiles/r04_ptxas_registers.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
