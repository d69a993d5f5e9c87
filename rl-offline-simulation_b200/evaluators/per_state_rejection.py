"""Gym-style PSRS environment over an ``OfflineDataset`` — the drop-in boundary of the path
(reference: offsim4rl/evaluators/per_state_rejection.py:7-98).

Same constructor, properties, ``reset_sampler`` / ``reset`` / ``step_dist`` and error behaviour.  The
observations are encoded to integer state ids on the GPU (an encoder exposing ``encode_device``) or by
the caller's own ``encoder.encode``; the dataset columns go to the device as arrays — no per-row tuples.
"""
from __future__ import annotations

import numpy as np

from ..core import RevealedRandomnessEnv
from ..spaces import is_discrete
from .psrs import PSRS


class PerStateRejectionSampling(RevealedRandomnessEnv):
    # "default": psrs.py:53-57, inside the CUDA step.  Subclasses that accept every transition set "accept_all" (also on the
    # device).  A subclass that OVERRIDES `_reject` with Python code (per_state_rejection.py:97-98 is a hook) is served by the
    # host-driven single-pop loop of `PSRS` — API compatibility, not a fast path.
    _reject_mode = "default"

    def __init__(self, dataset, num_states=None, encoder=None, new_step_api=False, device=None):
        if not is_discrete(dataset.observation_space) and num_states is None and encoder is None:
            raise ValueError("PerStateRejectionSampling only supports discrete observation spaces")
        if (num_states is None or encoder is None) and (num_states != encoder):
            raise ValueError("num_states and encoder either both need to be None, or both need to be specified")
        if not is_discrete(dataset.action_space):
            raise ValueError("PerStateRejectionSampling currently only supports discrete action spaces")
        overridden = type(self)._reject is not PerStateRejectionSampling._reject
        self._dataset = dataset
        exp = dataset.experience
        obs, next_obs = exp["observations"], exp["next_observations"]
        if encoder is not None:                                      # per_state_rejection.py:29-31
            zs, next_zs = self._encode(encoder, obs), self._encode(encoder, next_obs)
        else:                                                        # discrete observations are the states (:33-35)
            zs, next_zs = np.asarray(obs), np.asarray(next_obs)
        if "action_distributions" not in exp:
            raise ValueError("PerStateRejectionSampling needs the logged action_distributions")
        self.new_step_api = new_step_api
        nA = dataset.action_space.n
        self._obs, self._next_obs = obs, next_obs
        self._impl = PSRS.from_arrays(
            z=zs, z_next=next_zs, action=exp["actions"], reward=exp["rewards"], terminal=exp["terminals"],
            p_log=exp["action_distributions"], steps=exp.get("steps"), s=None, s_next=None,
            nS=num_states if num_states is not None else 25, nA=nA, reject=self._reject if overridden else self._reject_mode,
            device=device)
        # the environment hands back the logged observations themselves, not their ids
        self._impl._s_rows, self._impl._s_next_rows = obs, next_obs
        if self._impl.s is not None:
            self._impl.s = obs[int(self._impl._init_rows_host[self._impl._n_resets - 1])]

    @staticmethod
    def _encode(encoder, observations):
        if hasattr(encoder, "encode_device"):
            return encoder.encode_device(observations)               # stays on the GPU
        return np.asarray(encoder.encode(observations))

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
        raise NotImplementedError(
            f"{self.__class__.__name__} does not support step(). To implement Per-State Rejection "
            "Sampling efficiently, your agent needs to reveal its action distribution via the step_dist() method instead.")

    def step_dist(self, action_dist):
        if hasattr(action_dist, "probs") and not isinstance(action_dist, np.ndarray):   # torch.distributions.Distribution
            action_dist = action_dist.probs.detach().cpu().numpy()
        next_obs, r, done, info = self._impl.step(action_dist)
        dones = [done, False] if self.new_step_api else [done]       # terminated, truncated
        if next_obs is None:
            return (None,) * (6 if self.new_step_api else 5)
        return (info["a"], next_obs, r, *dones, info)

    def _reject(self, p_new, p_log, a) -> bool:
        """Host restatement of the default rule for callers that probe it; the simulation itself applies the
        rule inside the CUDA step."""
        a = int(a)
        M = (p_new / p_log).max()
        u = np.random.default_rng().random()
        return u > p_new[a] / p_log[a] / M
