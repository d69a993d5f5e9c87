"""Host side of the B200 PSRS engine: device buffers (torch) + calls through the C ABI.

PyTorch is plumbing here — it owns device memory and streams and nothing else; every computation on
the path runs in the hand-written kernels of ``libols_b200.so``.  There is no CPU fallback: without a
CUDA device (or without the built library) construction raises.

Vocabulary follows the reference (offsim4rl): a *log* of transitions is *bucketed* into per-state
*queues*; a *replica* is one independent reference ``PSRS`` object (own cursors, own uniform stream)
over the shared, immutable bucketed log.
"""
from __future__ import annotations

import ctypes as C
import secrets
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _abi as abi


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise RuntimeError("offsim4rl_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device(device if device is not None else "cuda")


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _as_tensor(x, device, dtype=None):
    """numpy / torch / list -> contiguous device tensor (dtype conversion on the device)."""
    if isinstance(x, torch.Tensor):
        t = x.to(device, non_blocking=True)
    else:
        a = np.asarray(x)
        if a.dtype == np.bool_:
            a = a.astype(np.uint8)
        t = torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)            # float -> int truncates toward zero, like the reference's int(a) (psrs.py:54)
    return t.contiguous()


# ----------------------------------------------------------------------------------------------- bucketed log
class BucketedLog:
    """The logged transitions bucketed into per-state FIFO queues, resident in HBM.

    Mirrors what ``PSRS.__init__`` + ``reset_sampler`` build (offsim4rl/evaluators/psrs.py:7-30):
    ``offsets``/``order``/``records`` are the CSR queues, ``init_*`` the list of initial states.
    ``base`` keeps the log-order queues; :meth:`seeded` derives the permutation ``reset_sampler(seed)``
    would produce without re-bucketing.
    """

    @classmethod
    def from_dataset(cls, dataset, n_buckets=None, z=None, z_next=None, device=None, **kw):
        """Ingest an ``OfflineDataset`` (in memory, ``load_npz`` or ``load_hdf5``): state ids from ``z``/``z_next`` (an
        encoder's output, host or device), else the recorded ``infos/z`` columns, else discrete observations
        (``OfflineDataset.device_columns``); the columns are uploaded once and bucketed on the GPU."""
        if z is not None and not isinstance(z, np.ndarray):        # ids already on the device stay there
            cols = dataset.device_columns(z=np.zeros(len(dataset), np.int32), z_next=np.zeros(len(dataset), np.int32))
            cols["z"], cols["z_next"] = z, z_next
        else:
            cols = dataset.device_columns(z=z, z_next=z_next)
        return cls(z=cols["z"], z_next=cols["z_next"], action=cols["action"], reward=cols["reward"], terminal=cols["terminal"],
                   p_log=cols["p_log"], steps=cols["steps"], n_buckets=n_buckets, device=device, **kw)

    def __init__(self, z, z_next, action, reward, terminal, p_log, s=None, s_next=None, steps=None,
                 n_buckets=None, z_offset=None, device=None, keep_columns=False, key_by_action=False, hot_records=False):
        dev = _require_cuda(device)
        self.device = dev
        lib = abi.load()
        with torch.cuda.device(dev):
            z = _as_tensor(z, dev, torch.int32)
            z_next = _as_tensor(z_next, dev, torch.int32)
            n = int(z.shape[0])
            if n >= 2 ** 31:
                raise ValueError("a log holds at most 2^31 - 1 transitions")
            action = _as_tensor(action, dev, torch.int32)
            reward = _as_tensor(reward, dev, torch.float32)
            terminal = _as_tensor(terminal, dev, torch.uint8)
            p = _as_tensor(p_log, dev)
            # NumPy promotion of the accept rule depends on the dtype the LOG was given in
            self.p_source_f32 = (p.dtype == torch.float32)
            if p.dim() != 2 or p.shape[0] != n:
                raise ValueError("p_log must be [n_rows, n_actions]")
            n_actions = int(p.shape[1])
            p_log64 = None
            if p.dtype == torch.float64:
                p32 = p.to(torch.float32)
                if n and not bool((p32.to(torch.float64) == p).all()):
                    p_log64 = p                 # exact float64 distribution kept for the accept rule
                p = p32 if p_log64 is None else p
            elif p.dtype != torch.float32:
                p = p.to(torch.float32)
            s = None if s is None else _as_tensor(s, dev, torch.int32)
            s_next = None if s_next is None else _as_tensor(s_next, dev, torch.int32)
            steps = None if steps is None else _as_tensor(steps, dev, torch.int32)
            for name, col in [("z_next", z_next), ("action", action), ("reward", reward), ("terminal", terminal)]:
                if col.shape[0] != n:
                    raise ValueError(f"column {name} has {col.shape[0]} rows, expected {n}")

            if n:
                lo = int(min(z.min().item(), z_next.min().item()))
                hi = int(max(z.max().item(), z_next.max().item()))
                if action.min().item() < 0 or action.max().item() >= n_actions:
                    raise ValueError("logged action outside [0, n_actions)")
            else:
                lo, hi = 0, -1
            if z_offset is None:
                z_offset = max(0, -lo)          # encoders/heuristic.py:23-24 emits -1
            if lo + z_offset < 0:
                raise ValueError(f"state id {lo} is below -z_offset")
            if n_buckets is None:
                n_buckets = hi + z_offset + 1
            n_buckets = max(int(n_buckets), hi + z_offset + 1, 1)

            # queue_evaluator.py:107 — QueueEvaluator keys its queues by (z, a): bucket on (z + z_offset) * nA + a
            self.key_by_action = bool(key_by_action)
            self.n_state_keys = n_buckets
            z_key, key_offset = z, int(z_offset)
            if key_by_action:
                z_key = (z + int(z_offset)) * n_actions + action
                key_offset = 0
                n_buckets = n_buckets * n_actions
            self.n_rows, self.n_actions = n, n_actions
            self.n_buckets, self.z_offset = n_buckets, int(z_offset)
            self.rec_bytes = int(lib.ols_record_bytes(n_actions))
            if self.rec_bytes == 0:
                raise ValueError(f"n_actions = {n_actions} is outside [1, 16]")
            self.p_log64 = p_log64

            log = abi.Log(n_rows=n, n_actions=n_actions, p_is_f64=int(p.dtype == torch.float64), z=_ptr(z),
                          z_next=_ptr(z_next), s=_ptr(s), s_next=_ptr(s_next), action=_ptr(action),
                          reward=_ptr(reward), terminal=_ptr(terminal), p_log=_ptr(p), step=_ptr(steps))
            self.offsets = torch.empty(n_buckets + 1, dtype=torch.int32, device=dev)
            self.s_next = s_next            # kept resident only when the agent's observation id differs from z
            self.records = torch.empty(max(n, 1) * self.rec_bytes // 4, dtype=torch.float32, device=dev)
            # 8-byte hot records next to the 32-byte ones (n_actions <= 4): what the fused runs read per popped record.
            # True / "probs": the logged probabilities as halves + the action (any log).  "tail": reward, next state, action
            # and terminal flag — for a log whose behaviour distribution is a function of the state (every queue holds ONE
            # distribution; checked bitwise on the device), where the accept rule needs only the record's action; falls back
            # to "probs" when the log does not qualify.
            if hot_records not in (False, True, None, "probs", "tail"):
                raise ValueError("hot_records must be False, True / 'probs', or 'tail'")
            want_hot = bool(hot_records) and n_actions <= 4
            self.hot_kind = abi.HOT_TAIL if (want_hot and hot_records == "tail" and p_log64 is None) else abi.HOT_PROBS
            self.hot = torch.empty(max(n, 1), dtype=torch.int64, device=dev) if want_hot else None
            self.queue_dist = torch.zeros((n_buckets, 8), dtype=torch.float32, device=dev) if self.hot_kind == abi.HOT_TAIL else None
            self._const_info = torch.zeros(1, dtype=torch.int32, device=dev) if self.hot_kind == abi.HOT_TAIL else None
            init_row = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
            init_z = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
            init_s = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
            sb = int(lib.ols_bucket_scratch_bytes(n, n_buckets))
            if sb == 0:
                raise ValueError(f"n_buckets = {n_buckets} is beyond what one bucketing pass supports")
            scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
            n_init_dev = torch.zeros(1, dtype=torch.int64, device=dev)
            log_key = abi.Log(n_rows=n, n_actions=n_actions, p_is_f64=int(p.dtype == torch.float64), z=_ptr(z_key),
                              z_next=_ptr(z_next), s=_ptr(s), s_next=_ptr(s_next), action=_ptr(action),
                              reward=_ptr(reward), terminal=_ptr(terminal), p_log=_ptr(p), step=_ptr(steps))
            q = abi.Queues(n_rows=n, n_init=0, n_actions=n_actions, n_buckets=n_buckets, z_offset=key_offset,
                           rec_bytes=self.rec_bytes, offsets=_ptr(self.offsets),
                           records=_ptr(self.records), init_row=_ptr(init_row), init_z=_ptr(init_z),
                           init_s=_ptr(init_s), s_next=_ptr(s_next), p_log64=_ptr(p_log64), hot=_ptr(self.hot),
                           hot_kind=self.hot_kind, _pad=0, queue_dist=_ptr(self.queue_dist), const_info=_ptr(self._const_info))
            abi.check(lib.ols_bucket(C.byref(log_key), C.byref(q), _ptr(scratch), sb, _stream_ptr()))
            abi.check(lib.ols_init_queue(C.byref(log), C.byref(q), _ptr(n_init_dev), _ptr(scratch), sb, _stream_ptr()))
            self.n_init = int(n_init_dev.item())        # one sync: construction is not the step path
            if self.hot_kind == abi.HOT_TAIL and int(self._const_info.item()) != 0:
                # some queue holds several distributions: the accept rule needs each record's own probabilities
                self.hot_kind, self.queue_dist, self._const_info = abi.HOT_PROBS, None, None
                q.hot_kind, q.queue_dist, q.const_info = abi.HOT_PROBS, None, None
                abi.check(lib.ols_pack_hot(C.byref(q), _stream_ptr()))
            dropped = n - int(self.offsets[-1].item())   # ols_bucket skips rows whose key is outside [0, n_buckets)
            if dropped:
                raise ValueError(f"{dropped} rows have a state id outside [-z_offset, n_buckets - z_offset)")
            self.init_row = init_row[: max(self.n_init, 1)]
            self.init_z = init_z[: max(self.n_init, 1)]
            self.init_s = init_s[: max(self.n_init, 1)]
            self.seed = None
            self._base = None           # a log-order log is its own base (no self-reference: the buffers must die by refcount)
            self._columns = dict(z=z_key, z_next=z_next, action=action, reward=reward, terminal=terminal, p_log=p,
                                 s=s, s_next=s_next, steps=steps) if keep_columns else None
            self._scratch = scratch if keep_columns else None

    def time_bucketing(self, iters: int = 3):
        """Re-run the bucketing kernels (ols_bucket + ols_init_queue) on the kept columns and return the mean
        device time per run in ms (CUDA events on the launching stream).  Needs ``keep_columns=True``."""
        if self._columns is None:
            raise RuntimeError("time_bucketing needs BucketedLog(..., keep_columns=True)")
        lib = abi.load()
        c = self._columns
        with torch.cuda.device(self.device):
            log = abi.Log(n_rows=self.n_rows, n_actions=self.n_actions, p_is_f64=int(c["p_log"].dtype == torch.float64),
                          z=_ptr(c["z"]), z_next=_ptr(c["z_next"]), s=_ptr(c["s"]), s_next=_ptr(c["s_next"]),
                          action=_ptr(c["action"]), reward=_ptr(c["reward"]), terminal=_ptr(c["terminal"]),
                          p_log=_ptr(c["p_log"]), step=_ptr(c["steps"]))
            q = self.queues_struct()
            n_init_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
            sb = self._scratch.numel()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                abi.check(lib.ols_bucket(C.byref(log), C.byref(q), _ptr(self._scratch), sb, _stream_ptr()))
                abi.check(lib.ols_init_queue(C.byref(log), C.byref(q), _ptr(n_init_dev), _ptr(self._scratch), sb, _stream_ptr()))
            e1.record()
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    def drop_columns(self):
        self._columns = None
        self._scratch = None

    # ------------------------------------------------------------------ views
    @property
    def base(self) -> "BucketedLog":
        """The log-order queues this log was derived from (itself for a freshly bucketed log)."""
        return self if self._base is None else self._base

    def queues_struct(self) -> abi.Queues:
        return abi.Queues(n_rows=self.n_rows, n_init=self.n_init, n_actions=self.n_actions, n_buckets=self.n_buckets,
                          z_offset=self.z_offset, rec_bytes=self.rec_bytes, offsets=_ptr(self.offsets),
                          records=_ptr(self.records), init_row=_ptr(self.init_row), init_z=_ptr(self.init_z),
                          init_s=_ptr(self.init_s), s_next=_ptr(self.s_next), p_log64=_ptr(self.p_log64),
                          hot=_ptr(getattr(self, "hot", None)), hot_kind=int(getattr(self, "hot_kind", 0)), _pad=0,
                          queue_dist=_ptr(getattr(self, "queue_dist", None)), const_info=_ptr(getattr(self, "_const_info", None)))

    @property
    def order(self) -> torch.Tensor:
        """Row index at every queue slot (the reference's queue order): a strided view of the records' row field."""
        words = self.rec_bytes // 4
        return self.records.view(torch.int32).view(-1, words)[: max(self.n_rows, 1), words - 2]

    def seeded(self, seed: Optional[int]) -> "BucketedLog":
        """Queues as ``reset_sampler(seed)`` leaves them (psrs.py:23,29-30).  ``None`` = log order."""
        if seed is None:
            return self.base
        if not (0 <= int(seed) < 2 ** 64):
            raise ValueError("seed must be in [0, 2^64)")
        base = self.base
        lib = abi.load()
        out = object.__new__(BucketedLog)
        out.__dict__.update(base.__dict__)
        with torch.cuda.device(base.device):
            out.records = torch.empty_like(base.records)
            out.hot = None if base.hot is None else torch.empty_like(base.hot)
            out.queue_dist = None if base.queue_dist is None else torch.empty_like(base.queue_dist)
            out.init_row = torch.empty_like(base.init_row)
            out.init_z = torch.empty_like(base.init_z)
            out.init_s = torch.empty_like(base.init_s)
            perm = torch.empty(base.n_rows + base.n_init + 1, dtype=torch.int32, device=base.device)
            src, dst = base.queues_struct(), out.queues_struct()
            abi.check(lib.ols_shuffle_queues(C.byref(src), C.byref(dst), C.c_uint64(int(seed)), _ptr(perm), _stream_ptr()))
        out.seed = int(seed)
        out._base = base
        return out

    def queue_rows(self, z: int, a: Optional[int] = None) -> torch.Tensor:
        """Row indices of queue ``z`` (or ``(z, a)``) in queue order (the reference's ``env.queues[z]``)."""
        b = (z + self.z_offset) * self.n_actions + a if self.key_by_action else z + self.z_offset
        lo, hi = int(self.offsets[b].item()), int(self.offsets[b + 1].item())
        return self.order[lo:hi]


# ----------------------------------------------------------------------------------------------- replicas
@dataclass
class RunResult:
    """Outputs of a fused learner run, as device tensors (reference: psrs.py:180-185, 235-239, 271).

    ``Q`` is [R, nS, nA]; with the AoS replica state it is gathered from the live entries on first access.  The
    per-replica counters are snapshots taken when the run returned.  ``trace_row`` / ``td_error`` are indexed by the replica's
    LIFETIME accepted count (since the last ``restart``): a resumed run continues the same trace, so pass the previous
    result back through ``outputs=`` when chunking (fresh buffers would start at the resumed index, not at 0)."""
    Gs: torch.Tensor                     # [R, ep_cap] discounted returns; row r valid up to n_returns[r]
    lengths: torch.Tensor                # [R, ep_cap] episode lengths; row r valid up to n_lengths[r]
    n_returns: torch.Tensor              # [R]
    n_lengths: torch.Tensor              # [R]
    status: torch.Tensor                 # [R]
    accepted: torch.Tensor               # [R] lifetime accepted transitions
    pops: torch.Tensor                   # [R] lifetime popped records
    trace_row: Optional[torch.Tensor]    # [R, trace_cap] accepted log rows (parity tap)
    td_error: Optional[torch.Tensor]     # [R, trace_cap]
    _Q: Optional[torch.Tensor] = None
    _q_fn: Optional[object] = None

    @property
    def Q(self) -> Optional[torch.Tensor]:
        if self._Q is None and self._q_fn is not None:
            self._Q = self._q_fn(None)
        return self._Q

    def q_rows(self, replicas) -> Optional[torch.Tensor]:
        """Q tables [k, nS, nA] of the listed replicas only (at the production shape the whole fleet's Q is 20 GB; a
        parity check of a sample gathers 160 KB per replica)."""
        idx = torch.as_tensor(replicas, dtype=torch.long, device=self.Gs.device)
        if self._Q is not None:
            return self._Q[idx]
        return None if self._q_fn is None else self._q_fn(idx)


class BatchedPSRS:
    """R independent PSRS simulators over one bucketed log — the batched vector form.

    Scalar reference API per replica: ``reset`` = psrs.py:32-37, ``step`` = psrs.py:39-57; fused
    drivers ``run_qlearn`` / ``run_expsarsa`` / ``run_evalmc`` = psrs.py:119-271.

    ``lanes`` = width of the lane group that serves one replica; results never depend on it.  Measured on B200
    (DESIGN.md 4.3): ``lanes=1`` with ``hot_records=True`` logs and ``aos_state=True`` when there are enough replicas
    to fill the GPU (>= ~1e5) and reject chains are short (the C4 learner: 188.6 ms per 1.25e9 transitions);
    ``lanes=32`` for long reject chains (C5: the round streams 128 hot records); 4-8 in between or for few replicas.
    """

    def __init__(self, log: BucketedLog, n_replicas: int, rng: str = "philox", seed: Optional[int] = 0,
                 seeds=None, u_feed=None, tie_feed=None, reject: str = "default", ring: bool = False,
                 ring_seed: int = 0, replica_id0: int = 0, lanes: int = 4, pops_per_lane: int = 0,
                 stage_smem: bool = True, prefetch: int = -1, aos_state: bool = False, slices: int = -1, batch_lanes: int = 0):
        self.lib = abi.load()
        self.log = log
        dev = log.device
        self.device = dev
        R = int(n_replicas)
        self.R = R
        self.reject_mode = {"default": abi.REJECT_DEFAULT, "accept_all": abi.REJECT_ACCEPT_ALL}[reject]
        self.lanes = 1 if self.reject_mode == abi.REJECT_ACCEPT_ALL else int(lanes)
        self.pops_per_lane = int(pops_per_lane)     # 0 = auto (2 with Philox / fed uniforms)
        self.stage_smem = bool(stage_smem)          # stage cursor table (+Q) in shared memory when it fits
        self.aos_state = bool(aos_state)            # cursor + Q row of a state in one 32-byte entry (fused float32 learners)
        self.aos = None
        self._epoch = 1                             # live epoch of the AoS entries (a zeroed table holds epoch 0: all stale)
        self.slices = int(slices)                   # time slices of a budgeted fused run: -1 auto, 0/1 off, k > 1 forced
        self._sched = None
        self.batch_lanes = int(batch_lanes)         # one-thread-per-replica learner ("tail" hot records): lanes per update batch
        self.prefetch = int(prefetch)               # rounds ahead prefetched into L2 (<= 0: off; measured neutral on B200)
        self.rng_mode = {"feed": abi.RNG_FEED, "pcg64": abi.RNG_PCG64, "philox": abi.RNG_PHILOX}[rng]
        if seed is None:
            seed = secrets.randbits(63)           # psrs.py:13 — default_rng(None) draws OS entropy
        self.seed = int(seed)
        with torch.cuda.device(dev):
            z32 = lambda: torch.zeros(max(R, 1), dtype=torch.int32, device=dev)
            z64 = lambda: torch.zeros(max(R, 1), dtype=torch.int64, device=dev)
            self.cursor = None if aos_state else torch.empty((max(R, 1), log.n_buckets), dtype=torch.int32, device=dev)
            self.init_cursor, self.cur_z, self.cur_s, self.status = z32(), z32(), z32(), z32()
            self.draws, self.ties, self.pops, self.accepted = z64(), z64(), z64(), z64()
            self.ep_index, self.ep_len_count, self.ep_t, self.ep_open = z32(), z32(), z32(), z32()
            self.ep_G = torch.zeros(max(R, 1), dtype=torch.float64, device=dev)
            self.ep_disc = torch.ones(max(R, 1), dtype=torch.float64, device=dev)
            self.pcg_state = (torch.zeros((max(R, 1), 4), dtype=torch.int64, device=dev)
                              if self.rng_mode == abi.RNG_PCG64 else None)
            self.u_feed = None if u_feed is None else _as_tensor(u_feed, dev, torch.float64)
            self.tie_feed = None if tie_feed is None else _as_tensor(tie_feed, dev, torch.float64)
            if self.rng_mode == abi.RNG_FEED and self.u_feed is None:
                raise ValueError("rng='feed' needs u_feed")
            self.tie_mode = abi.RNG_FEED if self.tie_feed is not None else abi.RNG_PHILOX
            self._seeds = None if seeds is None else _as_tensor(np.asarray(seeds, dtype=np.uint64).view(np.int64), dev)
            self.ring, self.ring_seed, self.replica_id0 = bool(ring), int(ring_seed), int(replica_id0)
            # step outputs
            self.out_row, self.out_action, self.out_s_next, self.out_z_from = z32(), z32(), z32(), z32()
            self.out_reward = torch.zeros(max(R, 1), dtype=torch.float32, device=dev)
            self.out_terminal = torch.zeros(max(R, 1), dtype=torch.uint8, device=dev)
        self.restart()

    # ------------------------------------------------------------------ structs
    def _feed_shape(self, t):
        if t is None:
            return 0, 0
        if t.dim() == 1:
            return int(t.shape[0]), 0
        return int(t.shape[1]), int(t.stride(0))

    def replicas_struct(self) -> abi.Replicas:
        ulen, ustride = self._feed_shape(self.u_feed)
        tlen, tstride = self._feed_shape(self.tie_feed)
        return abi.Replicas(
            n_replicas=self.R, replica_id0=self.replica_id0, cursor=_ptr(self.cursor), init_cursor=_ptr(self.init_cursor),
            cur_z=_ptr(self.cur_z), cur_s=_ptr(self.cur_s), status=_ptr(self.status), draws=_ptr(self.draws),
            ties=_ptr(self.ties), pops=_ptr(self.pops), accepted=_ptr(self.accepted), ep_index=_ptr(self.ep_index),
            ep_len_count=_ptr(self.ep_len_count), ep_t=_ptr(self.ep_t), ep_open=_ptr(self.ep_open), ep_G=_ptr(self.ep_G),
            ep_disc=_ptr(self.ep_disc), rng_mode=self.rng_mode, tie_mode=self.tie_mode, seed=self.seed,
            pcg_state=_ptr(self.pcg_state), u_feed=_ptr(self.u_feed), u_feed_len=ulen, u_stride=ustride,
            tie_feed=_ptr(self.tie_feed), tie_feed_len=tlen, tie_stride=tstride, ring_mode=int(self.ring), _pad=0,
            ring_seed=self.ring_seed)

    # ------------------------------------------------------------------ lifecycle
    def restart(self, log: Optional[BucketedLog] = None):
        """Fresh simulators (cursors 0, counters 0, generators re-seeded) — ``reset_sampler`` for the
        replica side; pass ``log`` to switch to another queue order (``log.seeded(k)``)."""
        if log is not None:
            if self.log is not None and (log.n_buckets != self.log.n_buckets or log.n_actions != self.log.n_actions):
                raise ValueError("restart(log): shape differs from the log this block was built for")
            self.log = log
        with torch.cuda.device(self.device):
            q, r = self.log.queues_struct(), self.replicas_struct()
            abi.check(self.lib.ols_replicas_init(C.byref(q), C.byref(r), _ptr(self._seeds), _stream_ptr()))
            if self.aos is not None:
                self._epoch += 1            # every entry of the previous run is stale now: cursors 0, Q 0
                if self._epoch >= 2 ** 31 - 1:
                    self.aos.zero_()
                    self._epoch = 1

    def reset(self, mask=None):
        """psrs.py:32-37 for every replica (or those with ``mask``): returns ``(s, status)`` tensors."""
        with torch.cuda.device(self.device):
            m = None if mask is None else _as_tensor(mask, self.device, torch.uint8)
            q, r = self.log.queues_struct(), self.replicas_struct()
            abi.check(self.lib.ols_reset(C.byref(q), C.byref(r), _ptr(m), _stream_ptr()))
        return self.cur_s, self.status

    def _need_cursor_table(self, what):
        if self.cursor is None:
            raise RuntimeError(f"{what} needs the separate cursor table: build the block with aos_state=False")

    def step(self, p_new, mask=None, compute: Optional[str] = None):
        """psrs.py:39-57 for every RUNNING replica.  ``p_new``: [nA] shared or [R, nA] per replica.

        Returns a dict of device tensors: ``row`` (accepted log row, -1 if none), ``action``, ``reward``,
        ``terminal``, ``s_next``, ``z_from``, ``status``.
        """
        self._need_cursor_table("step()")
        with torch.cuda.device(self.device):
            p = _as_tensor(p_new, self.device)
            if p.dtype not in (torch.float32, torch.float64):
                p = p.to(torch.float64)
            if p.shape[-1] != self.log.n_actions:
                raise ValueError("p_new must have n_actions entries per row")
            stride = 0 if p.dim() == 1 else int(p.stride(0))
            if p.dim() == 2 and p.shape[0] != self.R:
                raise ValueError("p_new must be [n_actions] or [n_replicas, n_actions]")
            if compute is None:
                # NumPy promotion: float32 arithmetic only if BOTH distributions are float32
                compute = "f32" if (p.dtype == torch.float32 and self.log.p_source_f32) else "f64"
            cm = abi.COMPUTE_F32 if compute == "f32" else abi.COMPUTE_F64
            m = None if mask is None else _as_tensor(mask, self.device, torch.uint8)
            q, r = self.log.queues_struct(), self.replicas_struct()
            out = abi.StepOut(row=_ptr(self.out_row), action=_ptr(self.out_action), reward=_ptr(self.out_reward),
                              terminal=_ptr(self.out_terminal), s_next=_ptr(self.out_s_next), z_from=_ptr(self.out_z_from))
            abi.check(self.lib.ols_step(C.byref(q), C.byref(r), _ptr(m), _ptr(p), int(p.dtype == torch.float64), stride,
                                        self.reject_mode, cm, self.lanes, C.byref(out), _stream_ptr()))
        return dict(row=self.out_row, action=self.out_action, reward=self.out_reward, terminal=self.out_terminal,
                    s_next=self.out_s_next, z_from=self.out_z_from, status=self.status)

    def step_action(self, actions, mask=None):
        """queue_evaluator.py:121-131 for every RUNNING replica: pop the head of queue ``(z, actions[r])`` — no
        rejection.  Needs a log bucketed with ``key_by_action=True``.  Returns the same dict as :meth:`step`."""
        if not self.log.key_by_action:
            raise ValueError("step_action needs BucketedLog(..., key_by_action=True)")
        self._need_cursor_table("step_action()")
        with torch.cuda.device(self.device):
            a = _as_tensor(actions, self.device, torch.int32)
            if a.dim() == 0:
                a = a.expand(self.R).contiguous()
            if a.shape[0] != self.R:
                raise ValueError("actions must hold one action per replica")
            m = None if mask is None else _as_tensor(mask, self.device, torch.uint8)
            q, r = self.log.queues_struct(), self.replicas_struct()
            out = abi.StepOut(row=_ptr(self.out_row), action=_ptr(self.out_action), reward=_ptr(self.out_reward),
                              terminal=_ptr(self.out_terminal), s_next=_ptr(self.out_s_next), z_from=_ptr(self.out_z_from))
            abi.check(self.lib.ols_step_action(C.byref(q), C.byref(r), _ptr(m), _ptr(a), C.byref(out), _stream_ptr()))
        return dict(row=self.out_row, action=self.out_action, reward=self.out_reward, terminal=self.out_terminal,
                    s_next=self.out_s_next, z_from=self.out_z_from, status=self.status)

    # ------------------------------------------------------------------ fused drivers
    def begin_driver_call(self):
        """A reference driver (psrs.py:119-271) counts its episodes from 0 on EVERY call and keeps consuming the queues
        where the previous call left them: clear the episode registers and re-arm the replicas a previous fused run left
        DONE / QUEUE_EMPTY (the run kernel starts every episode with `reset`, which reports INIT_EMPTY by itself).
        The batched API does not call this: there a second `run_*` RESUMES the first (chunked runs)."""
        with torch.cuda.device(self.device):
            for t in (self.ep_index, self.ep_len_count, self.ep_t, self.ep_open):
                t.zero_()
            self.ep_G.zero_()
            self.ep_disc.fill_(1.0)
            self.status.fill_(abi.RUNNING)

    def _schedule(self, x, n_episodes):
        if callable(x):
            # the reference calls a schedule lazily, for episode < n_episodes only (psrs.py:158,168): a schedule defined on
            # [0, n_episodes) must work; the kernel never reads entry n_episodes before it terminates with DONE
            vals = [float(x(e)) for e in range(n_episodes)]
            vals = vals + [vals[-1] if vals else 0.0]
        elif np.ndim(x) == 0:
            # constant schedules are kept on the device: building them again would put a blocking host-to-device copy (and so
            # a pipeline bubble) in front of every run
            key = (float(x), int(n_episodes))
            cache = self.__dict__.setdefault("_sched_cache", {})
            if key not in cache:
                if len(cache) > 64:
                    cache.clear()
                cache[key] = torch.full((n_episodes + 1,), float(x), dtype=torch.float64, device=self.device)
            return cache[key]
        else:
            vals = [float(v) for v in x]
            if len(vals) < n_episodes:
                raise ValueError("schedule shorter than n_episodes")
            vals = vals + [vals[-1]] * (n_episodes + 1 - len(vals))
        return torch.tensor(vals, dtype=torch.float64, device=self.device)

    def _run(self, agent, n_states, n_episodes, gamma, alpha=None, epsilon=None, pi=None, Q=None,
             q_dtype=torch.float32, max_steps=0, trace_cap=0, compute=None, ep_cap=None, outputs=None):
        dev = self.device
        R, nA = self.R, self.log.n_actions
        with torch.cuda.device(dev):
            ep_cap = int(ep_cap or (n_episodes + 1))
            use_aos = self.aos_state and agent != abi.AGENT_FIXED_PI
            if self.aos_state and not use_aos:
                self._need_cursor_table("a fixed-policy evaluation")
            if use_aos:
                if q_dtype != torch.float32 or nA > 4 or n_states != self.log.n_buckets or self.log.z_offset != 0 \
                        or self.log.s_next is not None or self.rng_mode == abi.RNG_PCG64 or self.reject_mode != abi.REJECT_DEFAULT:
                    raise ValueError("aos_state needs a float32 learner, <= 4 actions, n_states == n_buckets, z_offset 0, "
                                     "s == z, Philox/fed uniforms and the default reject rule")
                if self.aos is None:
                    self.aos = torch.zeros((max(R, 1), n_states, 8), dtype=torch.float32, device=dev)
                if Q is not None:                                       # caller-supplied initial Q: make every entry live
                    words = self.aos.view(torch.int32)
                    words[:, :, 0] = torch.where(words[:, :, 1] == self._epoch, words[:, :, 0], torch.zeros_like(words[:, :, 0]))
                    words[:, :, 1] = self._epoch
                    self.aos[:, :, 4:4 + nA] = Q.to(torch.float32)
                Q = None
            if outputs is None:
                Gs = torch.zeros((max(R, 1), ep_cap), dtype=torch.float64, device=dev)
                lengths = torch.zeros((max(R, 1), ep_cap), dtype=torch.int32, device=dev)
                trace_row = torch.full((max(R, 1), trace_cap), -1, dtype=torch.int32, device=dev) if trace_cap else None
                td = torch.zeros((max(R, 1), trace_cap), dtype=torch.float64, device=dev) if (trace_cap and agent == abi.AGENT_QLEARN) else None
                if agent != abi.AGENT_FIXED_PI and Q is None and not use_aos:
                    Q = torch.zeros((max(R, 1), n_states, nA), dtype=q_dtype, device=dev)
            else:
                Gs, lengths, trace_row, td = outputs.Gs, outputs.lengths, outputs.trace_row, outputs.td_error
                Q = None if use_aos else outputs.Q
            pi_t = None
            pi_is_f64 = 0
            if pi is not None:
                pi_t = _as_tensor(pi, dev)
                if pi_t.dtype not in (torch.float32, torch.float64):
                    pi_t = pi_t.to(torch.float64)
                pi_is_f64 = int(pi_t.dtype == torch.float64)
                if tuple(pi_t.shape) != (n_states, nA):
                    raise ValueError("pi must be [n_states, n_actions]")
            if compute is None:
                compute = "f32" if (pi_t is not None and not pi_is_f64 and self.log.p_source_f32) else "f64"
            a_t = None if alpha is None else self._schedule(alpha, n_episodes)
            e_t = None if epsilon is None else self._schedule(epsilon, n_episodes)
            if self._sched is None and self.slices not in (0, 1):
                n_sched = int(self.lib.ols_run_sched_bytes(R, self.lanes)) // 4
                self._sched = torch.zeros(max(n_sched, 1), dtype=torch.int32, device=dev)
            cfg = abi.Run(agent=agent, n_states=n_states, state_epoch=self._epoch, batch_lanes=self.batch_lanes, _pad1=0, slices=self.slices if max_steps > 0 else 0, sched=_ptr(self._sched), q_is_f64=int(Q is not None and Q.dtype == torch.float64),
                          state_aos=_ptr(self.aos if use_aos else None),
                          pi_is_f64=pi_is_f64, reject_mode=self.reject_mode,
                          compute_mode=abi.COMPUTE_F32 if compute == "f32" else abi.COMPUTE_F64, lanes=self.lanes,
                          n_episodes=int(n_episodes), ep_cap=ep_cap, stage_smem=int(self.stage_smem),
                          pops_per_lane=0 if self.rng_mode == abi.RNG_PCG64 else self.pops_per_lane,
                          prefetch=max(self.prefetch, 0),
                          max_steps=int(max_steps),
                          gamma=float(gamma), alpha=_ptr(a_t), epsilon=_ptr(e_t), pi=_ptr(pi_t), Q=_ptr(None if use_aos else Q), Gs=_ptr(Gs),
                          lengths=_ptr(lengths), trace_row=_ptr(trace_row), td_error=_ptr(td), trace_cap=int(trace_cap))
            q, r = self.log.queues_struct(), self.replicas_struct()
            abi.check(self.lib.ols_run(C.byref(q), C.byref(r), C.byref(cfg), _stream_ptr()))
        q_fn = None
        if use_aos:
            aos, epoch = self.aos, self._epoch

            def q_fn(idx):
                tab = aos if idx is None else aos[idx]
                live = (tab.view(torch.int32)[:, :, 1] == epoch).unsqueeze(-1)
                return torch.where(live, tab[:, :, 4:4 + nA], torch.zeros((), dtype=torch.float32, device=aos.device))
        # the per-replica counters are SNAPSHOTS: a later run / restart on this block does not change a result already
        # handed out (Gs, lengths and the trace buffers are the result's own unless passed back through `outputs=`)
        with torch.cuda.device(dev):
            return RunResult(_Q=Q, _q_fn=q_fn, Gs=Gs, lengths=lengths, n_returns=self.ep_index.clone(),
                             n_lengths=self.ep_len_count.clone(), status=self.status.clone(), accepted=self.accepted.clone(),
                             pops=self.pops.clone(), trace_row=trace_row, td_error=td)

    def run_qlearn(self, n_states, n_episodes, gamma, alpha=0.1, epsilon=1.0, Q=None, q_dtype=torch.float32,
                   max_steps=0, trace_cap=0, ep_cap=None, outputs=None) -> RunResult:
        """psrs.py:119-185 with ``behavior_policy = epsilon_greedy_policy`` (tabular.py:24-32), fused."""
        return self._run(abi.AGENT_QLEARN, n_states, n_episodes, gamma, alpha=alpha, epsilon=epsilon, Q=Q,
                         q_dtype=q_dtype, max_steps=max_steps, trace_cap=trace_cap, ep_cap=ep_cap, outputs=outputs)

    def run_expsarsa(self, n_states, n_episodes, pi, gamma, alpha=0.1, Q=None, q_dtype=torch.float32, max_steps=0,
                     trace_cap=0, compute=None, ep_cap=None, outputs=None) -> RunResult:
        """psrs.py:187-239, fused."""
        return self._run(abi.AGENT_EXPSARSA, n_states, n_episodes, gamma, alpha=alpha, pi=pi, Q=Q, q_dtype=q_dtype,
                         max_steps=max_steps, trace_cap=trace_cap, compute=compute, ep_cap=ep_cap, outputs=outputs)

    def run_evalmc(self, n_states, n_episodes, pi, gamma, max_steps=0, trace_cap=0, compute=None, ep_cap=None,
                   outputs=None) -> RunResult:
        """psrs.py:241-271, fused."""
        return self._run(abi.AGENT_FIXED_PI, n_states, n_episodes, gamma, pi=pi, max_steps=max_steps,
                         trace_cap=trace_cap, compute=compute, ep_cap=ep_cap, outputs=outputs)

    def run_actions(self, n_episodes, gamma, actions=None, pi=None, max_steps=0, trace_cap=0, ep_cap=None, outputs=None) -> RunResult:
        """queue_evaluator.py:113-131 as ONE launch for many steps: every replica resets, steps with the agent's own actions
        until its episode ends, and again.  ``actions``: [n_steps] shared or [R, n_steps] per replica (the k-th ``step()``
        call of a replica takes element k), or ``pi`` [n_states, nA]: the agent samples from a fixed policy with the
        replica's counter-based stream.  Needs ``BucketedLog(..., key_by_action=True)``.  Returns / lengths as evalMC_psrs."""
        if not self.log.key_by_action:
            raise ValueError("run_actions needs BucketedLog(..., key_by_action=True)")
        if (actions is None) == (pi is None):
            raise ValueError("give either actions or pi")
        self._need_cursor_table("run_actions()")
        dev, R, nA = self.device, self.R, self.log.n_actions
        with torch.cuda.device(dev):
            ep_cap = int(ep_cap or (n_episodes + 1))
            if outputs is None:
                Gs = torch.zeros((max(R, 1), ep_cap), dtype=torch.float64, device=dev)
                lengths = torch.zeros((max(R, 1), ep_cap), dtype=torch.int32, device=dev)
                trace_row = torch.full((max(R, 1), trace_cap), -1, dtype=torch.int32, device=dev) if trace_cap else None
            else:
                Gs, lengths, trace_row = outputs.Gs, outputs.lengths, outputs.trace_row
            a_t = pi_t = None
            a_len = a_stride = n_states = pi_is_f64 = 0
            if actions is not None:
                a_t = _as_tensor(actions, dev, torch.int32)
                if a_t.dim() == 2 and a_t.shape[0] != R:
                    raise ValueError("actions must be [n_steps] or [n_replicas, n_steps]")
                a_len, a_stride = (int(a_t.shape[0]), 0) if a_t.dim() == 1 else (int(a_t.shape[1]), int(a_t.stride(0)))
            else:
                pi_t = _as_tensor(pi, dev)
                if pi_t.dtype not in (torch.float32, torch.float64):
                    pi_t = pi_t.to(torch.float64)
                if pi_t.dim() != 2 or pi_t.shape[1] != nA:
                    raise ValueError("pi must be [n_states, n_actions]")
                n_states, pi_is_f64 = int(pi_t.shape[0]), int(pi_t.dtype == torch.float64)
            cfg = abi.QueueRun(n_states=n_states, pi_is_f64=pi_is_f64, n_episodes=int(n_episodes), ep_cap=ep_cap, max_steps=int(max_steps),
                               gamma=float(gamma), pi=_ptr(pi_t), actions=_ptr(a_t), actions_len=a_len, a_stride=a_stride, Gs=_ptr(Gs),
                               lengths=_ptr(lengths), trace_row=_ptr(trace_row), trace_cap=int(trace_cap))
            q, r = self.log.queues_struct(), self.replicas_struct()
            abi.check(self.lib.ols_run_actions(C.byref(q), C.byref(r), C.byref(cfg), _stream_ptr()))
            return RunResult(Gs=Gs, lengths=lengths, n_returns=self.ep_index.clone(), n_lengths=self.ep_len_count.clone(),
                             status=self.status.clone(), accepted=self.accepted.clone(), pops=self.pops.clone(), trace_row=trace_row,
                             td_error=None)

    # ------------------------------------------------------------------ evaluators
    def learning_curve(self, result: RunResult, n_episodes: int, all_reduce: bool = True):
        """Per-episode ``(sum, sumsq, count)`` of the returns over replicas — over ALL ranks when
        ``torch.distributed`` is initialised (one small all-reduce, the job's only collective)."""
        dev = self.device
        with torch.cuda.device(dev):
            out = torch.zeros((3, max(n_episodes, 1)), dtype=torch.float64, device=dev)
            sb = int(self.lib.ols_curve_scratch_bytes(self.R, n_episodes))
            scratch = torch.empty(max(sb, 8), dtype=torch.uint8, device=dev)
            abi.check(self.lib.ols_reduce_curves(_ptr(result.Gs), _ptr(result.n_returns), self.R, int(result.Gs.shape[1]),
                                                 int(n_episodes), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(scratch),
                                                 sb, _stream_ptr()))
            if all_reduce:
                from .parallel import all_reduce_sum_
                all_reduce_sum_(out)
        return out[0, :n_episodes], out[1, :n_episodes], out[2, :n_episodes]

    def replica_totals(self, result: RunResult):
        """Per-replica total return and total simulated length (segmented reduction over its episodes)."""
        dev = self.device
        with torch.cuda.device(dev):
            ret = torch.zeros(max(self.R, 1), dtype=torch.float64, device=dev)
            ln = torch.zeros(max(self.R, 1), dtype=torch.int64, device=dev)
            abi.check(self.lib.ols_replica_totals(_ptr(result.Gs), _ptr(result.lengths), _ptr(result.n_returns),
                                                  _ptr(result.n_lengths), self.R, int(result.Gs.shape[1]), _ptr(ret),
                                                  _ptr(ln), _stream_ptr()))
        return ret, ln


class BatchedPSRSExo:
    """R independent ``PSRS_Exo`` simulators (psrs.py:59-117) — the batched vector form: observations split into an
    endogenous state s and an exogenous state x; ``log_s`` / ``log_x`` are the SAME rows bucketed by s and by x (x' in the
    ``z_next`` column of ``log_x``); one pop consumes the head of both queues, the accept rule runs on the s-record.
    ``reset(seeds)`` re-seeds each replica's PCG64 generator like ``PSRS_Exo.reset(seed)`` (psrs.py:92) and pops its next
    initial (s, x); ``step(p_new)`` is psrs.py:99-117 for every running replica."""

    def __init__(self, log_s: BucketedLog, log_x: BucketedLog, n_replicas: int, seed: int = 0, lanes: int = 4):
        if log_s.n_rows != log_x.n_rows or log_s.n_init != log_x.n_init:
            raise ValueError("log_s and log_x must bucket the same rows")
        self.log_s, self.log_x = log_s, log_x
        self.sim = BatchedPSRS(log_s, n_replicas, rng="pcg64", seed=seed, lanes=lanes)
        dev, R = self.sim.device, self.sim.R
        with torch.cuda.device(dev):
            self.cursor_x = torch.zeros((max(R, 1), log_x.n_buckets), dtype=torch.int32, device=dev)
            self.cur_x = torch.zeros(max(R, 1), dtype=torch.int32, device=dev)
            self.out_x_next = torch.zeros(max(R, 1), dtype=torch.int32, device=dev)

    def restart(self, log_s: Optional[BucketedLog] = None, log_x: Optional[BucketedLog] = None):
        """``reset_sampler`` for the replica side: fresh cursors; pass the ``seeded(k)`` logs to switch queue order."""
        if log_s is not None:
            self.log_s, self.log_x = log_s, log_x
        self.sim.restart(self.log_s)
        self.cursor_x.zero_()

    def reset(self, seeds=None, mask=None):
        """psrs.py:91-97 for every replica (or those with ``mask``): ``seeds`` [R] (or one int) re-seed the generators."""
        sim = self.sim
        with torch.cuda.device(sim.device):
            m = None if mask is None else _as_tensor(mask, sim.device, torch.uint8)
            sd = None
            if seeds is not None:
                arr = np.full(sim.R, int(seeds), dtype=np.uint64) if np.ndim(seeds) == 0 else np.asarray(seeds, dtype=np.uint64)
                sd = _as_tensor(arr.view(np.int64), sim.device)
            r = sim.replicas_struct()
            abi.check(sim.lib.ols_reseed(C.byref(r), _ptr(sd), _ptr(m), _stream_ptr()))
            before = sim.init_cursor.clone()
            sim.reset(mask)
            took = sim.init_cursor > before                      # replicas that popped an initial state: its x comes with it
            idx = (sim.init_cursor - 1).clamp(min=0).long()
            self.cur_x.copy_(torch.where(took, self.log_x.init_z[idx.clamp(max=max(self.log_x.n_init - 1, 0))], self.cur_x))
        return sim.cur_z, self.cur_x, sim.status

    def step(self, p_new, mask=None):
        """psrs.py:99-117 for every RUNNING replica; returns the dict of :meth:`BatchedPSRS.step` plus ``x_next``."""
        sim = self.sim
        with torch.cuda.device(sim.device):
            p = _as_tensor(p_new, sim.device)
            if p.dtype not in (torch.float32, torch.float64):
                p = p.to(torch.float64)
            stride = 0 if p.dim() == 1 else int(p.stride(0))
            compute = abi.COMPUTE_F32 if (p.dtype == torch.float32 and self.log_s.p_source_f32) else abi.COMPUTE_F64
            m = None if mask is None else _as_tensor(mask, sim.device, torch.uint8)
            qs, qx, r = self.log_s.queues_struct(), self.log_x.queues_struct(), sim.replicas_struct()
            out = abi.StepOut(row=_ptr(sim.out_row), action=_ptr(sim.out_action), reward=_ptr(sim.out_reward),
                              terminal=_ptr(sim.out_terminal), s_next=_ptr(sim.out_s_next), z_from=_ptr(sim.out_z_from))
            abi.check(sim.lib.ols_step_exo(C.byref(qs), C.byref(qx), C.byref(r), _ptr(self.cursor_x), _ptr(self.cur_x), _ptr(m), _ptr(p),
                                           int(p.dtype == torch.float64), stride, compute, sim.lanes, C.byref(out),
                                           _ptr(self.out_x_next), _stream_ptr()))
        return dict(row=sim.out_row, action=sim.out_action, reward=sim.out_reward, terminal=sim.out_terminal, s_next=sim.out_s_next,
                    x_next=self.out_x_next, z_from=sim.out_z_from, status=sim.status)

    def run_evalmc(self, n_obs, n_episodes, pi, gamma, x_mult, reseed_base=None, max_steps=0, trace_cap=0, ep_cap=None,
                   outputs=None) -> RunResult:
        """evalMC_psrs (psrs.py:241-271) over these PSRS_Exo simulators as ONE launch: the agent's policy row is
        ``pi[s * x_mult + x]`` (the linear ``o_combine_func``); every episode re-seeds the replica's PCG64 generator with
        ``reseed_base[r] + episode`` (``None``: ``episode``, what the reference's drivers do through ``reset(seed=episode)``)."""
        sim = self.sim
        dev, R, nA = sim.device, sim.R, self.log_s.n_actions
        with torch.cuda.device(dev):
            ep_cap = int(ep_cap or (n_episodes + 1))
            if outputs is None:
                Gs = torch.zeros((max(R, 1), ep_cap), dtype=torch.float64, device=dev)
                lengths = torch.zeros((max(R, 1), ep_cap), dtype=torch.int32, device=dev)
                trace_row = torch.full((max(R, 1), trace_cap), -1, dtype=torch.int32, device=dev) if trace_cap else None
            else:
                Gs, lengths, trace_row = outputs.Gs, outputs.lengths, outputs.trace_row
            pi_t = _as_tensor(pi, dev)
            if pi_t.dtype not in (torch.float32, torch.float64):
                pi_t = pi_t.to(torch.float64)
            if tuple(pi_t.shape) != (n_obs, nA):
                raise ValueError("pi must be [n_obs, n_actions]")
            compute = abi.COMPUTE_F32 if (pi_t.dtype == torch.float32 and self.log_s.p_source_f32) else abi.COMPUTE_F64
            rb = None if reseed_base is None else _as_tensor(np.asarray(reseed_base, dtype=np.uint64).view(np.int64), dev)
            cfg = abi.ExoRun(n_obs=int(n_obs), x_mult=int(x_mult), pi_is_f64=int(pi_t.dtype == torch.float64), compute_mode=compute,
                             n_episodes=int(n_episodes), ep_cap=ep_cap, max_steps=int(max_steps), gamma=float(gamma), pi=_ptr(pi_t),
                             reseed_base=_ptr(rb), Gs=_ptr(Gs), lengths=_ptr(lengths), trace_row=_ptr(trace_row), trace_cap=int(trace_cap))
            qs, qx, r = self.log_s.queues_struct(), self.log_x.queues_struct(), sim.replicas_struct()
            abi.check(sim.lib.ols_run_exo(C.byref(qs), C.byref(qx), C.byref(r), _ptr(self.cursor_x), _ptr(self.cur_x), C.byref(cfg),
                                          _stream_ptr()))
            return RunResult(Gs=Gs, lengths=lengths, n_returns=sim.ep_index.clone(), n_lengths=sim.ep_len_count.clone(),
                             status=sim.status.clone(), accepted=sim.accepted.clone(), pops=sim.pops.clone(), trace_row=trace_row,
                             td_error=None)


# ----------------------------------------------------------------------------------------------- synthetic data
def synth_columns(n_rows, n_states, pb_state, seed=0, p_done=1.0 / 50, device=None):
    """Device-generated synthetic log of the benchmark shape: returns the column dict
    ``BucketedLog`` takes.  ``pb_state`` is the per-state behaviour policy [n_states, n_actions]."""
    dev = _require_cuda(device)
    lib = abi.load()
    with torch.cuda.device(dev):
        pb = _as_tensor(pb_state, dev, torch.float32)
        nA = int(pb.shape[1])
        cols = dict(
            z=torch.empty(n_rows, dtype=torch.int32, device=dev), z_next=torch.empty(n_rows, dtype=torch.int32, device=dev),
            action=torch.empty(n_rows, dtype=torch.int32, device=dev), reward=torch.empty(n_rows, dtype=torch.float32, device=dev),
            terminal=torch.empty(n_rows, dtype=torch.uint8, device=dev),
            p_log=torch.empty((n_rows, nA), dtype=torch.float32, device=dev),
            steps=torch.empty(n_rows, dtype=torch.int32, device=dev))
        log = abi.Log(n_rows=n_rows, n_actions=nA, p_is_f64=0, z=_ptr(cols["z"]), z_next=_ptr(cols["z_next"]), s=_ptr(None),
                      s_next=_ptr(None), action=_ptr(cols["action"]), reward=_ptr(cols["reward"]),
                      terminal=_ptr(cols["terminal"]), p_log=_ptr(cols["p_log"]), step=_ptr(cols["steps"]))
        abi.check(lib.ols_synth_log(C.byref(log), _ptr(pb), int(n_states), C.c_uint64(int(seed)), float(p_done), _stream_ptr()))
    return cols
