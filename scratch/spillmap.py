"""List the local-memory (spill) slots of a kernel and the source lines that store / reload them."""
import subprocess, re, os, collections, sys
d = '/root/repo/gpurun_out/elf3'
kn = sys.argv[1] if len(sys.argv) > 1 else 'run_kernelILi1ELi2EfLi2ELi2ELb1'
for f in os.listdir(d):
    if not f.endswith('.cubin'): continue
    txt = subprocess.run(f"cuobjdump -elf {d}/{f} | grep -o '\\.text\\.[A-Za-z0-9_$]*'", shell=True, capture_output=True, text=True).stdout
    names = [n for n in set(txt.split()) if kn in n]
    if not names: continue
    dis = subprocess.run(f"nvdisasm -g {d}/{f}", shell=True, capture_output=True, text=True).stdout.split('\n')
    start = [i for i, l in enumerate(dis) if l.strip().startswith('.section') and names[0] in l][0]
    end = [i for i in range(start + 1, len(dis)) if dis[i].strip().startswith('.section')][0]
    cur = None; out = collections.defaultdict(list)
    for ln in dis[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
        if m and ('STL' in m.group(2) or 'LDL' in m.group(2)):
            off = re.search(r'\[R1(\+0x[0-9a-f]+)?\]', m.group(2))
            out[off.group(1) or '+0x0'].append(('STL' if 'STL' in m.group(2) else 'LDL', cur))
    for k in sorted(out, key=lambda x: int(x, 16)):
        st = [c for o, c in out[k] if o == 'STL']; ld = [c for o, c in out[k] if o == 'LDL']
        print(k, 'ST@', sorted(set(l for f, l in st)), 'LD@', sorted(set(l for f, l in ld)))
    break
