"""Action-distribution helpers of the data-collection side (offsim4rl/utils/prob_utils.py:10-29)."""
import numpy as np

from ..data import ProbDistribution
from ..spaces import is_discrete


def get_uniform_dist(action_space, action_dist_type=ProbDistribution.Discrete):
    if not is_discrete(action_space):
        raise NotImplementedError(f"Unsupported action_space: {action_space}")
    n = int(action_space.n)
    if action_dist_type == ProbDistribution.Discrete:
        return np.full((n,), 1.0 / n)
    if action_dist_type == ProbDistribution.TorchDistribution:
        import torch
        return torch.distributions.Categorical(probs=torch.ones(n) / n)
    raise NotImplementedError(f"Unsupported action_dist_type: {action_dist_type} for action_space: {action_space}")


def sample_dist(prob_dist):
    """One action from a probability vector (legacy global NumPy stream, as the reference) or a torch distribution."""
    if isinstance(prob_dist, np.ndarray):
        return np.random.choice(len(prob_dist), p=prob_dist)
    if hasattr(prob_dist, "sample"):
        return prob_dist.sample().numpy()
    raise NotImplementedError(f"Unsupported prob_dist: {prob_dist}")
