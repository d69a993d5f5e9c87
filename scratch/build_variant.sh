#!/bin/bash
# usage: scratch/build_variant.sh <name> <nvcc extra flags...> : builds variants/libols_b200_<name>.so, then restores the default build
name=$1; shift
mkdir -p variants
OLS_NVCC_EXTRA="$*" python rl-offline-simulation_b200/build.py --force > /dev/null && cp rl-offline-simulation_b200/libols_b200.so variants/libols_b200_$name.so
python rl-offline-simulation_b200/build.py --force > /dev/null
ls -la variants/libols_b200_$name.so
