"""Run another script of this repo with cudaLimitMaxL2FetchGranularity set first:
    python scratch/with_l2_granularity.py 32 bench.py --steps 3 --warmup 3 --skip-cpu ...
(experiment: do the isolated 32-byte accesses of the C4 run cost less DRAM traffic with a smaller L2 fetch granularity?)"""
import ctypes
import glob
import os
import runpy
import sys

import torch

torch.cuda.init()
torch.zeros(1, device="cuda")
lib = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
rt = ctypes.CDLL(lib[0] if lib else "libcudart.so")
gran = int(sys.argv[1])
v = ctypes.c_size_t()
rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(gran))  # cudaLimitMaxL2FetchGranularity = 0x05
rt.cudaDeviceGetLimit(ctypes.byref(v), 5)
print(f"cudaLimitMaxL2FetchGranularity: asked {gran}, rc {rc}, device reports {v.value}", file=sys.stderr)
sys.argv = sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
