#!/bin/bash
# One gpurun call of round 2: GPU tests, the default bench line, ncu captures of the shipped run kernels.
# usage: scratch/gpu_call.sh <tag> [steps...]   steps: test bench ncu_c4 ncu_c5 launches
tag=$1; shift
mkdir -p gpurun_out
for step in "$@"; do
  case $step in
    test) timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" ;;
    bench) timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" ;;
    ncu_c4) timeout 900 ncu --set full --clock-control none --import-source on -k regex:"run_kernel|learner_kernel" -s 1 -c 1 -f -o gpurun_out/${tag}_ncu_c4 \
        python bench.py --steps 1 --warmup 1 --budget 1024 --slices 16 --skip-cpu --skip-e2e --skip-stress > gpurun_out/${tag}_ncu_c4.log 2>&1; echo "ncu_c4 rc=$?" ;;
    ncu_c5) timeout 900 ncu --set full --clock-control none --import-source on -k regex:"run_kernel|eval_kernel" -s 1 -c 1 -f -o gpurun_out/${tag}_ncu_c5 \
        python bench.py --workload c5 --steps 1 --warmup 1 --skip-cpu --skip-e2e > gpurun_out/${tag}_ncu_c5.log 2>&1; echo "ncu_c5 rc=$?" ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
        python bench.py --steps 2 --warmup 1 --budget 500 --skip-cpu --skip-stress > gpurun_out/${tag}_launches.log 2>&1; echo "launches rc=$?" ;;
    *) echo "unknown step $step" ;;
  esac
done
tail -5 gpurun_out/${tag}_pytest.log 2>/dev/null
