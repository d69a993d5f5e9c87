"""Queue evaluator: queues keyed by ``(z, a)``, the agent's own sampled action picks the queue, no rejection
(reference: offsim4rl/evaluators/queue_evaluator.py:8-131).  Same bucketing and cursor machinery as PSRS with the
combined key ``(z + z_offset) * nA + a``; each step is one ``ols_step_action`` call."""
from __future__ import annotations

import secrets

import numpy as np

from .. import _abi as abi
from ..engine import BatchedPSRS, BucketedLog
from ..spaces import is_discrete
from .per_state_rejection import PerStateRejectionSampling


class QueueEvaluator_impl(object):
    """reference: queue_evaluator.py:90-131.  ``buffer``: legacy tuples ``(s, a, r, s', done, p, info)`` with
    ``info['z']``, ``info['next_z']``, ``info['t']``."""

    def __init__(self, buffer, nS=25, nA=5, device=None):
        rows = list(buffer)
        n = len(rows)
        self._init_columns(
            z=np.fromiter((r[6]["z"] for r in rows), dtype=np.int64, count=n),
            z_next=np.fromiter((r[6]["next_z"] if "next_z" in r[6] else r[6]["z_next"] for r in rows), dtype=np.int64, count=n),
            action=np.array([int(r[1]) for r in rows], dtype=np.int64), reward=np.array([r[2] for r in rows]),
            terminal=np.array([bool(r[4]) for r in rows]),
            p_log=np.stack([np.asarray(r[5]) for r in rows]) if n else np.zeros((0, nA), dtype=np.float32),
            steps=np.fromiter((r[6]["t"] for r in rows), dtype=np.int64, count=n),
            s=[r[0] for r in rows], s_next=[r[3] for r in rows], nS=nS, nA=nA, device=device)

    @classmethod
    def from_arrays(cls, z, z_next, action, reward, terminal, p_log, steps=None, s=None, s_next=None, nS=25, nA=None, device=None):
        self = cls.__new__(cls)
        p_log = np.asarray(p_log)
        self._init_columns(z=z, z_next=z_next, action=np.asarray(action), reward=np.asarray(reward), terminal=np.asarray(terminal),
                           p_log=p_log, steps=steps, s=s, s_next=s_next, nS=nS, nA=nA if nA is not None else p_log.shape[1],
                           device=device)
        return self

    def _init_columns(self, z, z_next, action, reward, terminal, p_log, steps, s, s_next, nS, nA, device):
        import torch
        self.nS, self.nA = nS, nA
        to_host = lambda a: (a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)).astype(np.int64)
        self._z, self._z_next = to_host(z), to_host(z_next)
        self._action, self._reward, self._terminal, self._p_log = action, reward, terminal, p_log
        self._s_rows = self._z if s is None else s
        self._s_next_rows = self._z_next if s_next is None else s_next
        self._log = BucketedLog(z=z if isinstance(z, torch.Tensor) else self._z,
                                z_next=z_next if isinstance(z_next, torch.Tensor) else self._z_next, action=action, reward=reward,
                                terminal=terminal, p_log=p_log, steps=steps, device=device, key_by_action=True)
        self._sim = None
        self.s = self.z = None
        self.reset_sampler()
        self.reset()

    def reset_sampler(self, seed=None):                                     # queue_evaluator.py:102-111
        if seed is None:
            seed = secrets.randbits(63)
        self._cur = self._log.seeded(int(seed))
        self._init_rows_host = self._cur.init_row[: self._cur.n_init].cpu().numpy()
        if self._sim is None:
            self._sim = BatchedPSRS(self._cur, 1, rng="philox", seed=int(seed), reject="accept_all")
        else:
            self._sim.restart(self._cur)
            if self.z is not None:
                self._sim.cur_z.fill_(int(self.z))
                self._sim.status.fill_(abi.RUNNING)
        self._n_resets = 0

    def reset(self, seed=None):                                             # queue_evaluator.py:113-119
        _, status = self._sim.reset()
        if int(status[0]) != abi.RUNNING:
            self.s = None
            return None
        row = int(self._init_rows_host[self._n_resets])
        self._n_resets += 1
        self.z, self.s = int(self._z[row]), self._s_rows[row]
        return self.s

    def step(self, a):                                                      # queue_evaluator.py:121-131
        a = int(a)
        out = self._sim.step_action(np.array([a], dtype=np.int32))
        row = int(out["row"][0])
        if row < 0:
            if int(self._sim.status[0]) == abi.NO_QUEUE:
                raise KeyError((self.z, a))                                 # queues[z, a] never logged
            return None, None, None, None
        z_from = self.z
        self.s, self.z = self._s_next_rows[row], int(self._z_next[row])
        return self.s, self._reward[row], self._terminal[row], {"z": z_from, "next_z": self.z, "a": self._action[row],
                                                               "p": self._p_log[row]}


class QueueEvaluator:
    """Gym-style front end over an ``OfflineDataset`` (reference: queue_evaluator.py:8-88)."""

    def __init__(self, dataset, num_states=None, encoder=None, device=None):
        if not is_discrete(dataset.observation_space) and num_states is None and encoder is None:
            raise ValueError("QueueEvaluator only supports discrete observation spaces")
        if (num_states is None or encoder is None) and (num_states != encoder):
            raise ValueError("num_states and encoder either both need to be None, or both need to be specified")
        if not is_discrete(dataset.action_space):
            raise ValueError("QueueEvaluator currently only supports discrete action spaces")
        self._dataset = dataset
        exp = dataset.experience
        obs, next_obs = exp["observations"], exp["next_observations"]
        if encoder is not None:
            zs, next_zs = PerStateRejectionSampling._encode(encoder, obs), PerStateRejectionSampling._encode(encoder, next_obs)
        else:
            zs, next_zs = np.asarray(obs), np.asarray(next_obs)
        nA = dataset.action_space.n
        p_log = exp["action_distributions"] if "action_distributions" in exp else np.full((len(exp["actions"]), nA), 1.0 / nA, np.float32)
        self._impl = QueueEvaluator_impl.from_arrays(
            z=zs, z_next=next_zs, action=exp["actions"], reward=exp["rewards"], terminal=exp["terminals"], p_log=p_log,
            steps=exp.get("steps"), nS=num_states if num_states is not None else 25, nA=nA, device=device)
        self._impl._s_rows, self._impl._s_next_rows = obs, next_obs
        if self._impl.s is not None:
            self._impl.s = obs[int(self._impl._init_rows_host[self._impl._n_resets - 1])]

    @property
    def observation_space(self):
        return self._dataset.observation_space

    @property
    def action_space(self):
        return self._dataset.action_space

    def reset_sampler(self, seed=None):
        return self._impl.reset_sampler(seed=seed)

    def reset(self, seed=None):
        return self._impl.reset(seed=seed)

    def step(self, action):
        next_obs, r, done, info = self._impl.step(action)
        if next_obs is None:
            return None, None, None, None
        return next_obs, r, done, info

    def step_dist(self, action_dist):
        a = action_dist.sample()                                            # queue_evaluator.py:84
        next_obs, r, done, info = self._impl.step(a)
        if next_obs is None:
            return None, None, None, None, None
        return info["a"], next_obs, r, done, info
