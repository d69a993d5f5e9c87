"""Env protocol of the path (offsim4rl/core.py:7-45): an environment whose ``step_dist`` takes the
agent's action DISTRIBUTION.  No gym dependency: the class only fixes the method names."""
from __future__ import annotations


class RevealedRandomnessEnv:
    """Gym-style environment with ``step_dist(action_dist) -> (action, obs, reward, done[, truncated], info)``."""

    def reset(self, seed=None):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def step_dist(self, action_dist):
        raise NotImplementedError
