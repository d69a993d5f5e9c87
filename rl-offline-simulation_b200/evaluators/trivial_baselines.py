"""Trivial baselines to compare PSRS with (reference: offsim4rl/evaluators/trivial_baselines.py:6-29)."""
from .per_state_rejection import PerStateRejectionSampling


class _DummyEncoder:
    """Maps every observation to state 0 (trivial_baselines.py:26-29)."""

    def encode(self, observation):
        return [0] * observation.shape[0]


class FollowObservationOnly(PerStateRejectionSampling):
    """Follow the observation queues but accept every transition (trivial_baselines.py:6-10)."""
    _reject_mode = "accept_all"


class FollowActionOnly(PerStateRejectionSampling):
    """Reject on the action probabilities but treat all observations alike (trivial_baselines.py:12-15)."""

    def __init__(self, dataset, **kwargs):
        super().__init__(dataset, num_states=1, encoder=_DummyEncoder(), **kwargs)


class ServeRandomTransitions(PerStateRejectionSampling):
    """Just serve random transitions (trivial_baselines.py:17-24)."""
    _reject_mode = "accept_all"

    def __init__(self, dataset, **kwargs):
        super().__init__(dataset, num_states=1, encoder=_DummyEncoder(), **kwargs)
