#!/bin/bash
# usage: scratch/sweep.sh <tag> "<args1>" "<args2>" ... : short c4 bench runs (no CPU / e2e / stress legs), one JSON line each
tag=$1; shift
mkdir -p gpurun_out
: > gpurun_out/${tag}_sweep.jsonl
for a in "$@"; do
  echo "### $a" >> gpurun_out/${tag}_sweep.jsonl
  # an item may start with VAR=value settings (e.g. OLS_B200_LIB=variants/libols_b200_mb6.so)
  envs=""; rest="$a"
  while [[ "$rest" == [A-Z_]*=* ]]; do
    envs="$envs ${rest%% *}"
    if [[ "$rest" == *" "* ]]; then rest="${rest#* }"; else rest=""; fi
  done
  env $envs timeout 300 python bench.py --steps 3 --warmup 2 --skip-cpu --skip-e2e --skip-stress $rest >> gpurun_out/${tag}_sweep.jsonl 2>> gpurun_out/${tag}_sweep.err
done
python - <<PY
import json
for line in open('gpurun_out/${tag}_sweep.jsonl'):
    if line.startswith('###'): print(line.strip()); continue
    try:
        d=json.loads(line); print('   value %.3e  kernel_ms %.2f  pops/s %.3e' % (d['value'], d['roofline']['kernel_ms'], d['pops_per_s']))
    except Exception as e: print('   ?', line[:200])
PY
