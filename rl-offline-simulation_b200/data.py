"""Column-store of logged transitions — the part of offsim4rl/data.py the PSRS path consumes
(data.py:16-30 ``ProbDistribution``, :33-35 ``Transition``, :38-79 + :100-118 ``OfflineDataset``).
HDF5 persistence (data.py:81-98,120-175) is outside the path and not provided."""
from __future__ import annotations

import enum
import logging
from collections import namedtuple


class ProbDistribution(enum.Enum):
    """How the logged action distribution is described (data.py:16-30)."""
    NoProbability = 0
    LoggedActionOnly = 1
    Discrete = 2
    TorchDistribution = 3


Transition = namedtuple(
    "Transition",
    ["episode_id", "step", "observation", "action", "action_distribution", "reward", "next_observation", "terminal", "info"])

_REQUIRED = ("observations", "actions", "rewards", "next_observations", "terminals")


class OfflineDataset:
    """``experience`` keys — required: observations, actions, rewards, next_observations, terminals;
    optional: episode_ids, steps, action_distributions, infos (data.py:45-58)."""

    def __init__(self, observation_space, action_space, action_dist_type, **experience):
        self._validate_experience(experience)
        self.observation_space = observation_space
        self.action_space = action_space
        self.action_dist_type = action_dist_type
        self.experience = experience

    def __len__(self):
        return self.experience["observations"].shape[0]

    def iterate_row_tuples(self):
        e = self.experience
        for i in range(len(self)):
            yield Transition(
                e["episode_ids"][i] if "episode_ids" in e else None,
                e["steps"][i] if "steps" in e else 0,            # data.py:72: no steps -> every row is t = 0
                e["observations"][i], e["actions"][i],
                e["action_distributions"][i] if "action_distributions" in e else None,
                e["rewards"][i], e["next_observations"][i], e["terminals"][i],
                e["infos"][i] if "infos" in e else {})

    @staticmethod
    def _validate_experience(experience):
        for k in _REQUIRED:
            if k not in experience:
                raise ValueError(f"Missing required key {k} in experience")
        if experience["observations"].shape != experience["next_observations"].shape:
            raise ValueError("Shapes in observations and next_observations do not match")
        n = experience["observations"].shape[0]
        for k, v in experience.items():
            if len(v) != n:
                raise ValueError(f"Length of {k} ({len(v)}) does not match length of observations ({n})")
        if "steps" not in experience:
            logging.warning("Missing steps in experience. Algorithms may need to assume all states can be initial states...")
        if "episode_ids" not in experience:
            logging.warning("Missing episode_ids in experience. Some algorithms may not be compatible with this dataset.")
