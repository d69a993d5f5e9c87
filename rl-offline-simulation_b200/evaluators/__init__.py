"""Offline learner simulation evaluators (reference: offsim4rl/evaluators)."""
from . import psrs  # noqa: F401
from .per_state_rejection import PerStateRejectionSampling  # noqa: F401
from .psrs import PSRS, PSRS_Exo, evalMC_psrs, expSARSA_psrs, qlearn_psrs  # noqa: F401
from .trivial_baselines import FollowActionOnly, FollowObservationOnly, ServeRandomTransitions  # noqa: F401
from .queue_evaluator import QueueEvaluator, QueueEvaluator_impl  # noqa: F401
