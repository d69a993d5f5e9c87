"""SASS of the two benchmarked kernels of the shipped library -> profiles/r04_sass_benchmarked_kernels.txt (opcode histogram,
memory / sync / conversion instructions: the 256-bit accesses, shuffles, votes the design names)."""
import collections, re, subprocess
SO = "rl-offline-simulation_b200/libols_b200.so"
KERNELS = [("learner_kernel (C4: one thread per replica, TAIL hot records, AoS state)", "_ZN7ols_sim14learner_kernelE10ols_queues12ols_replicas7ols_run15ols_philox_keys"),
           ("eval_kernel<2> (C5: one warp per replica, streaming evaluation over TAIL hot records)", "_ZN7ols_sim11eval_kernelILi2EEEv10ols_queues12ols_replicas7ols_run15ols_philox_keys")]
out = ["# round 2 (final), shipped libols_b200.so (sm_100a only): SASS of the two benchmarked kernels — opcode histogram and the memory instructions",
       "# cuobjdump -sass -fun <kernel> rl-offline-simulation_b200/libols_b200.so", ""]
for title, sym in KERNELS:
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", sym, SO], capture_output=True, text=True).stdout
    ops, mem = collections.Counter(), collections.Counter()
    for line in txt.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        full = m.group(1)
        ops[full.split(".")[0]] += 1
        if re.match(r"(LDG|STG|LDS|STS|LDL|STL|ATOM|RED|BAR|SHFL|VOTE|MATCH|REDUX|CCTL|F2I|I2F|MUFU|LDGSTS|MEMBAR|ERRBAR)", full):
            mem[full] += 1
    n = sum(ops.values())
    out.append(f"## {title}: {n} instructions")
    out.append("opcodes: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(28)))
    out.append("memory / sync / conversion instructions: " + ", ".join(f"{k} x{v}" for k, v in sorted(mem.items())))
    out.append("")
open("profiles/r04_sass_benchmarked_kernels.txt", "w").write("\n".join(out))
print("\n".join(out)[:3000])
