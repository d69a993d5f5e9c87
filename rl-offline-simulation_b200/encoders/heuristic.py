"""Heuristic (integer-math) state encoders on the device.

``CartpoleBoxEncoder`` keeps the reference's interface (offsim4rl/encoders/heuristic.py:10-71):
``encode(observations[N, 4]) -> int64 ids[N]`` in {-1, 0..161}, and ``get_box`` for one observation.
``GridEncoder`` discretises MyGridNaviCoords observations (offsim4rl/envs/gridworld.py:205-211,123-124).
Both also expose ``encode_device`` (torch CUDA tensor in, int32 CUDA tensor out, no host copy), which
``PerStateRejectionSampling`` uses when it builds the queues.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _abi as abi
from ..engine import _as_tensor, _ptr, _require_cuda, _stream_ptr


class CartpoleBoxEncoder:
    """Heuristic encoder based on Sutton's pole.c boxes (reference: encoders/heuristic.py:10-60)."""

    N_BOXES = 162
    ONE_DEGREE = 0.0174532
    SIX_DEGREES = 0.1047192
    TWELVE_DEGREES = 0.2094384
    FIFTY_DEGREES = 0.87266

    def __init__(self, device=None):
        self._device = device

    def encode_device(self, observations) -> torch.Tensor:
        dev = _require_cuda(self._device)
        with torch.cuda.device(dev):
            obs = _as_tensor(observations, dev)
            if obs.dtype not in (torch.float32, torch.float64):
                obs = obs.to(torch.float64)
            if obs.dim() != 2 or obs.shape[1] != 4:
                raise ValueError("CartpoleBoxEncoder expects observations of shape [N, 4]")
            ids = torch.empty(obs.shape[0], dtype=torch.int32, device=dev)
            abi.check(abi.load().ols_encode_cartpole_box(_ptr(obs), int(obs.dtype == torch.float64), obs.shape[0],
                                                         _ptr(ids), _stream_ptr()))
        return ids

    def encode(self, observations):
        return self.encode_device(observations).cpu().numpy().astype(np.int64)

    @staticmethod
    def get_box(x, x_dot, theta, theta_dot):
        obs = np.array([[x, x_dot, theta, theta_dot]])
        return int(CartpoleBoxEncoder().encode(obs)[0])


class GridEncoder:
    """``id = sum_d min(floor(cells * obs[d]), cells - 1) * cells**d`` (x + cells*y for 2-D coordinates)."""

    def __init__(self, cells=5, device=None):
        self.cells = int(cells)
        self._device = device

    def encode_device(self, observations) -> torch.Tensor:
        dev = _require_cuda(self._device)
        with torch.cuda.device(dev):
            obs = _as_tensor(observations, dev, torch.float32)
            if obs.dim() != 2:
                raise ValueError("GridEncoder expects observations of shape [N, d]")
            ids = torch.empty(obs.shape[0], dtype=torch.int32, device=dev)
            abi.check(abi.load().ols_encode_grid(_ptr(obs), obs.shape[0], obs.shape[1], self.cells, _ptr(ids), _stream_ptr()))
        return ids

    def encode(self, observations):
        return self.encode_device(observations).cpu().numpy().astype(np.int64)
