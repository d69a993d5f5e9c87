"""offsim4rl_b200 — B200-native Per-State Rejection Sampling (PSRS) offline learner simulation.

A drop-in for ONE path of microsoft/rl-offline-simulation (``offsim4rl``): the PSRS environment
(``evaluators.psrs.PSRS``, ``evaluators.per_state_rejection.PerStateRejectionSampling``), its state
encoders, tabular policies and learner drivers — plus a batched vector form (``BatchedPSRS``) that
runs thousands to millions of independent replicas in one kernel.  Host code is Python/PyTorch;
all computation is hand-written sm_100a CUDA behind the C ABI in ``include/ols_b200.h``.

Submodules that touch the device import lazily, so ``import offsim4rl_b200`` works on a machine
without a GPU; constructing anything does not.
"""
__version__ = "0.1.0"

from . import _abi  # noqa: F401  (ctypes mirror of the C ABI; loading the library is lazy)


def __getattr__(name):
    if name in ("BucketedLog", "BatchedPSRS", "BatchedPSRSExo", "RunResult", "synth_columns"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
