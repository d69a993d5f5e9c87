"""ctypes binding of include/ols_b200.h — the stub a maintainer of the reference would add.

The library is CUDA-only: there is no CPU implementation behind these symbols and no fallback.
Importing this module needs the built ``libols_b200.so`` (``python rl-offline-simulation_b200/build.py``);
a missing library raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# OLS_B200_LIB overrides the library path (used to A/B differently tuned builds of the same ABI)
LIB_PATH = os.environ.get("OLS_B200_LIB") or os.path.join(HERE, "libols_b200.so")

ABI_VERSION = 2

# return codes
OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = 0, 1, 2, 3
# replica status
RUNNING, DONE, QUEUE_EMPTY, INIT_EMPTY, NO_QUEUE, FEED_EMPTY, IDLE = range(7)
STATUS_NAMES = ["RUNNING", "DONE", "QUEUE_EMPTY", "INIT_EMPTY", "NO_QUEUE", "FEED_EMPTY", "IDLE"]
# options
REJECT_DEFAULT, REJECT_ACCEPT_ALL = 0, 1
COMPUTE_F64, COMPUTE_F32 = 0, 1
RNG_FEED, RNG_PCG64, RNG_PHILOX = 0, 1, 2
AGENT_FIXED_PI, AGENT_EXPSARSA, AGENT_QLEARN = 0, 1, 2
HOT_PROBS, HOT_TAIL = 0, 1

p = C.c_void_p
i32, i64, u64, f64 = C.c_int32, C.c_int64, C.c_uint64, C.c_double


class Log(C.Structure):
    _fields_ = [("n_rows", i64), ("n_actions", i32), ("p_is_f64", i32), ("z", p), ("z_next", p), ("s", p),
                ("s_next", p), ("action", p), ("reward", p), ("terminal", p), ("p_log", p), ("step", p)]


class Queues(C.Structure):
    _fields_ = [("n_rows", i64), ("n_init", i64), ("n_actions", i32), ("n_buckets", i32), ("z_offset", i32),
                ("rec_bytes", i32), ("offsets", p), ("records", p), ("init_row", p), ("init_z", p),
                ("init_s", p), ("s_next", p), ("p_log64", p), ("hot", p), ("hot_kind", i32), ("_pad", i32), ("queue_dist", p), ("const_info", p)]


class Replicas(C.Structure):
    _fields_ = [("n_replicas", i64), ("replica_id0", i64), ("cursor", p), ("init_cursor", p), ("cur_z", p),
                ("cur_s", p), ("status", p), ("draws", p), ("ties", p), ("pops", p), ("accepted", p),
                ("ep_index", p), ("ep_len_count", p), ("ep_t", p), ("ep_open", p), ("ep_G", p), ("ep_disc", p),
                ("rng_mode", i32), ("tie_mode", i32), ("seed", u64), ("pcg_state", p), ("u_feed", p),
                ("u_feed_len", i64), ("u_stride", i64), ("tie_feed", p), ("tie_feed_len", i64), ("tie_stride", i64),
                ("ring_mode", i32), ("_pad", i32), ("ring_seed", u64)]


class StepOut(C.Structure):
    _fields_ = [("row", p), ("action", p), ("reward", p), ("terminal", p), ("s_next", p), ("z_from", p)]


class Run(C.Structure):
    _fields_ = [("agent", i32), ("n_states", i32), ("q_is_f64", i32), ("pi_is_f64", i32), ("reject_mode", i32),
                ("compute_mode", i32), ("lanes", i32), ("n_episodes", i32), ("ep_cap", i32), ("stage_smem", i32),
                ("pops_per_lane", i32), ("prefetch", i32), ("slices", i32), ("state_epoch", i32), ("batch_lanes", i32), ("_pad1", i32), ("max_steps", i64), ("gamma", f64), ("alpha", p), ("epsilon", p), ("pi", p), ("Q", p), ("state_aos", p), ("sched", p), ("Gs", p),
                ("lengths", p), ("trace_row", p), ("td_error", p), ("trace_cap", i64)]


class QueueRun(C.Structure):
    _fields_ = [("n_states", i32), ("pi_is_f64", i32), ("n_episodes", i32), ("ep_cap", i32), ("max_steps", i64), ("gamma", f64),
                ("pi", p), ("actions", p), ("actions_len", i64), ("a_stride", i64), ("Gs", p), ("lengths", p), ("trace_row", p),
                ("trace_cap", i64)]


class ExoRun(C.Structure):
    _fields_ = [("n_obs", i32), ("x_mult", i32), ("pi_is_f64", i32), ("compute_mode", i32), ("n_episodes", i32), ("ep_cap", i32),
                ("max_steps", i64), ("gamma", f64), ("pi", p), ("reseed_base", p), ("Gs", p), ("lengths", p), ("trace_row", p),
                ("trace_cap", i64)]


STRUCTS = [Log, Queues, Replicas, StepOut, Run, QueueRun, ExoRun]

# every symbol include/ols_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "ols_abi_version": (C.c_int, []),
    "ols_last_error": (C.c_char_p, []),
    "ols_struct_bytes": (C.c_size_t, [C.c_int]),
    "ols_kernel_launches": (C.c_uint64, []),
    "ols_encode_grid": (C.c_int, [p, i64, i32, i32, p, p]),
    "ols_encode_cartpole_box": (C.c_int, [p, i32, i64, p, p]),
    "ols_encode_mlp": (C.c_int, [p, i64, i32, p, p, i32, p, p, i32, C.c_float, p, p]),
    "ols_encode_mlp_audit": (C.c_int, [p, i64, i32, p, p, i32, p, p, i32, C.c_float, p, p, p, p]),
    "ols_record_bytes": (i32, [i32]),
    "ols_bucket_scratch_bytes": (C.c_size_t, [i64, i32]),
    "ols_bucket": (C.c_int, [C.POINTER(Log), C.POINTER(Queues), p, C.c_size_t, p]),
    "ols_init_queue": (C.c_int, [C.POINTER(Log), C.POINTER(Queues), p, p, C.c_size_t, p]),
    "ols_shuffle_queues": (C.c_int, [C.POINTER(Queues), C.POINTER(Queues), u64, p, p]),
    "ols_pack_hot": (C.c_int, [C.POINTER(Queues), p]),
    "ols_replicas_init": (C.c_int, [C.POINTER(Queues), C.POINTER(Replicas), p, p]),
    "ols_reset": (C.c_int, [C.POINTER(Queues), C.POINTER(Replicas), p, p]),
    "ols_step": (C.c_int, [C.POINTER(Queues), C.POINTER(Replicas), p, p, i32, i64, i32, i32, i32,
                           C.POINTER(StepOut), p]),
    "ols_step_action": (C.c_int, [C.POINTER(Queues), C.POINTER(Replicas), p, p, C.POINTER(StepOut), p]),
    "ols_step_exo": (C.c_int, [C.POINTER(Queues), C.POINTER(Queues), C.POINTER(Replicas), p, p, p, p, i32, i64, i32, i32,
                               C.POINTER(StepOut), p, p]),
    "ols_run_exo": (C.c_int, [C.POINTER(Queues), C.POINTER(Queues), C.POINTER(Replicas), p, p, C.POINTER(ExoRun), p]),
    "ols_reseed": (C.c_int, [C.POINTER(Replicas), p, p, p]),
    "ols_run": (C.c_int, [C.POINTER(Queues), C.POINTER(Replicas), C.POINTER(Run), p]),
    "ols_run_actions": (C.c_int, [C.POINTER(Queues), C.POINTER(Replicas), C.POINTER(QueueRun), p]),
    "ols_run_sched_bytes": (C.c_size_t, [i64, i32]),
    "ols_curve_scratch_bytes": (C.c_size_t, [i64, i32]),
    "ols_reduce_curves": (C.c_int, [p, p, i64, i32, i32, p, p, p, p, C.c_size_t, p]),
    "ols_replica_totals": (C.c_int, [p, p, p, p, i64, i32, p, p, p]),
    "ols_synth_log": (C.c_int, [C.POINTER(Log), p, i32, u64, f64, p]),
}


class OlsError(RuntimeError):
    """A C-ABI call returned a non-zero status."""

    def __init__(self, code, message):
        super().__init__(f"ols_b200 error {code}: {message}")
        self.code = code


_lib = None


def load():
    """dlopen the CUDA library (once) and type every entry point.  No fallback: raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. Build it with `python rl-offline-simulation_b200/build.py` "
            "(needs nvcc; sm_100a). There is no CPU fallback for the PSRS engine.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the build disagree
        fn.restype = res
        fn.argtypes = args
    if lib.ols_abi_version() != ABI_VERSION:
        raise ImportError(f"libols_b200.so has ABI {lib.ols_abi_version()}, binding expects {ABI_VERSION}")
    for k, st in enumerate(STRUCTS):
        if lib.ols_struct_bytes(k) != C.sizeof(st):
            raise ImportError(f"struct layout mismatch for {st.__name__}: "
                              f"library {lib.ols_struct_bytes(k)} B, binding {C.sizeof(st)} B")
    _lib = lib
    return lib


def check(code):
    if code != OK:
        raise OlsError(code, load().ols_last_error().decode("utf-8", "replace"))
