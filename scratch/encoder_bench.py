"""Times the three state encoders on device-resident observations (scratch tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import offsim4rl_b200 as ob
from offsim4rl_b200.encoders import CartpoleBoxEncoder, GridEncoder, HOMEREncoder

dev = torch.device("cuda:0")

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

N = 100_000_000
obs4 = (torch.rand((N, 4), device=dev) - 0.5) * torch.tensor([4.0, 3.0, 0.4, 3.0], device=dev)
ms = timeit(lambda: CartpoleBoxEncoder().encode_device(obs4))
print("cartpole_box rows %d ms %.3f GB/s %.0f" % (N, ms, N * (16 + 4) / ms / 1e6))
del obs4
obs2 = torch.rand((N, 2), device=dev)
ms = timeit(lambda: GridEncoder(5).encode_device(obs2))
print("grid rows %d ms %.3f GB/s %.0f" % (N, ms, N * (8 + 4) / ms / 1e6))
g = torch.Generator(device="cpu").manual_seed(0)
for nz in (25, 50):
    W1, b1 = torch.randn(64, 2, generator=g), torch.randn(64, generator=g)
    W2, b2 = torch.randn(nz, 64, generator=g), torch.randn(nz, generator=g)
    enc = HOMEREncoder.from_weights(W1.numpy(), b1.numpy(), W2.numpy(), b2.numpy(), device=dev)
    ms = timeit(lambda: enc.encode_device(obs2))
    flops = 2.0 * N * (2 * 64 + 64 * nz)
    print("mlp nz %d rows %d ms %.3f GB/s %.0f TFLOP/s(fp32 FMA) %.2f" % (nz, N, ms, N * 12 / ms / 1e6, flops / ms / 1e9))
