"""Map an ncu report's per-SASS-instruction counts to source lines using nvdisasm line info of the matching .so."""
import csv, re, collections, subprocess, sys, os
rep, so, kernel_substr = os.path.abspath(sys.argv[1]), os.path.abspath(sys.argv[2]), sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
ELF = "/tmp/ols_elf"; os.makedirs(ELF, exist_ok=True)
for f in os.listdir(ELF):          # (stale cubins of another build would be matched by name)
    os.remove(os.path.join(ELF, f))
subprocess.run(f"cd {ELF} && cuobjdump -xelf all {so} > /dev/null 2>&1", shell=True)
src_csv = subprocess.run(f"ncu -i {rep} --page source --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(src_csv.splitlines()))
kname = rows[0][1]
hi = [i for i, r in enumerate(rows) if 'Instructions Executed' in r][0]
h = rows[hi]; ci = h.index('Instructions Executed'); sm = h.index('# Samples')
base = int(rows[hi + 1][0], 16)
# find cubin + mangled name
target = None
for f in os.listdir(ELF):
    if not f.endswith(".cubin"): continue
    txt = subprocess.run(f"cuobjdump -elf " + ELF + f"/{f} | grep -o '\\.text\\.[A-Za-z0-9_$]*'", shell=True, capture_output=True, text=True).stdout
    for name in set(txt.split()):
        if kernel_substr in name:
            target = (f, name[len(".text."):])
if not target:
    print("kernel not found"); sys.exit(1)
dis = subprocess.run("nvdisasm -g " + ELF + f"/{target[0]}", shell=True, capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.strip().startswith('.section') and '.text.' + target[1] in l][0]
end = [i for i in range(start + 1, len(dis)) if dis[i].strip().startswith('.section')][0]
cur = None; a2l = {}
for ln in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m and cur: a2l[int(m.group(1), 16)] = cur
agg = collections.Counter(); aggs = collections.Counter(); tot = tots = 0
for r in rows[hi + 1:]:
    try: n = float(r[ci]); a = int(r[0], 16) - base; s = float(r[sm] or 0)
    except Exception: continue
    k = a2l.get(a, ('?', 0)); agg[k] += n; aggs[k] += s; tot += n; tots += s
srcs = {}
print(kname[:150]); print("total warp-instructions %.4e" % tot)
for (f, l), n in agg.most_common(top):
    if f not in srcs:
        pth = os.path.join("/root/repo/rl-offline-simulation_b200/csrc", f)
        srcs[f] = open(pth).read().split("\n") if os.path.exists(pth) else []
    text = srcs[f][l - 1].strip()[:100] if 0 < l <= len(srcs[f]) else ''
    print(f"{100*n/tot:5.1f}% i {100*aggs[(f,l)]/max(tots,1):5.1f}% s {f[:14]}:{l:4d} {text}")
