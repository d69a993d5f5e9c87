#!/bin/bash
# Builds variants of the one-lane hot-record learner (CTA size x CTAs per SM) next to the production library, for an A/B on the
# GPU box:   scratch/sweep_g1_builds.sh            (here, no GPU needed: nvcc cross-compiles)
#            gpurun -- 'for so in scratch/libols_g1_*.so; do echo $so; OLS_B200_LIB=$PWD/$so python bench.py --steps 3 --warmup 3 \
#                       --skip-cpu --skip-e2e --skip-stress | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); \
#                       print(d[\"value\"], d[\"roofline\"][\"kernel_ms\"])"; done'
# Measured in round 1 (C4, kernel ms): 256x4 @64 regs 204-208, 256x3 @80 regs 196.9, 128x7 @72 regs 187.8 (production).
set -e
cd "$(dirname "$0")/.."
python rl-offline-simulation_b200/build.py > /dev/null
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include"
OTHERS=$(ls build/obj/*.o | grep -v "ols_sim_nv1.o\|ols_sim.o")
for v in "64 14" "64 16" "128 6" "128 8" "256 3"; do
  set -- $v
  D="-DOLS_G1_THREADS=$1 -DOLS_MINB_G1=$2"
  nvcc $F $D -c rl-offline-simulation_b200/csrc/ols_sim_nv1.cu -o /tmp/g1_nv1_$1_$2.o &
  nvcc $F $D -c rl-offline-simulation_b200/csrc/ols_sim.cu -o /tmp/g1_sim_$1_$2.o &
  wait
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o scratch/libols_g1_$1x$2.so /tmp/g1_nv1_$1_$2.o /tmp/g1_sim_$1_$2.o $OTHERS
  echo "built scratch/libols_g1_$1x$2.so"
done
