"""Latent-state MLP encoder, inference only (reference: offsim4rl/encoders/homer.py:159-168 over
offsim4rl/encoders/models.py:15-19).

``obs_encoder = Linear(dO, H) -> LeakyReLU(0.01) -> Linear(H, nZ)``; ``encode`` returns
``argmax(log_softmax(logits))`` == first ``argmax(logits)``.  Training the encoder (homer.py:35-157)
is an offline pre-step outside the PSRS path: this class consumes trained weights, either a
``state_dict`` of the reference's ``EncoderModel`` / its ``obs_encoder`` or raw arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _abi as abi
from ..engine import _as_tensor, _ptr, _require_cuda, _stream_ptr


class HOMEREncoder:
    def __init__(self, obs_dim=None, action_dim=None, latent_size=None, hidden_size=None, model_path=None,
                 log_dir=None, device=None):
        self._device = device
        self.obs_dim, self.latent_size, self.hidden_size = obs_dim, latent_size, hidden_size
        self.negative_slope = 0.01          # nn.LeakyReLU() default, models.py:17
        self._w = None
        if model_path:
            self.load(model_path)

    # -- weights ---------------------------------------------------------------------------------
    @classmethod
    def from_weights(cls, W1, b1, W2, b2, negative_slope=0.01, device=None):
        enc = cls(device=device)
        enc._set(W1, b1, W2, b2)
        enc.negative_slope = float(negative_slope)
        return enc

    def _set(self, W1, b1, W2, b2):
        dev = _require_cuda(self._device)
        W1, b1, W2, b2 = (_as_tensor(w, dev, torch.float32) for w in (W1, b1, W2, b2))
        H, dO = W1.shape
        nZ = W2.shape[0]
        if b1.shape != (H,) or W2.shape != (nZ, H) or b2.shape != (nZ,):
            raise ValueError("weights do not form Linear(dO,H) -> Linear(H,nZ)")
        self.obs_dim, self.hidden_size, self.latent_size = dO, H, nZ
        self._w = (W1, b1, W2, b2)

    def load_state_dict(self, sd):
        """Accepts the reference's EncoderModel state_dict (keys ``obs_encoder.0.weight`` ...) or the
        ``obs_encoder`` Sequential's own (``0.weight`` ...)."""
        pre = "obs_encoder." if any(k.startswith("obs_encoder.") for k in sd) else ""
        self._set(sd[pre + "0.weight"], sd[pre + "0.bias"], sd[pre + "2.weight"], sd[pre + "2.bias"])

    def load(self, model_path):
        self.load_state_dict(torch.load(model_path, map_location="cpu"))       # models.py:71-72

    def train(self, *args, **kwargs):
        raise NotImplementedError("training the HOMER encoder is outside the PSRS path; train with the reference "
                                  "(offsim4rl/encoders/homer.py:35-157) and load the weights here")

    # -- inference -------------------------------------------------------------------------------
    def encode_device(self, observations) -> torch.Tensor:
        if self._w is None:
            raise ValueError("Model not initialized. Either train a new model for the encoder or load an existing one.")
        dev = _require_cuda(self._device)
        W1, b1, W2, b2 = self._w
        with torch.cuda.device(dev):
            obs = _as_tensor(observations, dev, torch.float32)          # torch.tensor(obs, dtype=torch.float) homer.py:163
            if obs.dim() != 2 or obs.shape[1] != self.obs_dim:
                raise ValueError(f"expected observations of shape [N, {self.obs_dim}]")
            ids = torch.empty(obs.shape[0], dtype=torch.int32, device=dev)
            abi.check(abi.load().ols_encode_mlp(_ptr(obs), obs.shape[0], self.obs_dim, _ptr(W1), _ptr(b1), self.hidden_size,
                                                _ptr(W2), _ptr(b2), self.latent_size, C.c_float(self.negative_slope),
                                                _ptr(ids), _stream_ptr()))
        return ids

    def encode(self, observations):
        return self.encode_device(observations).cpu().numpy().astype(np.int64)

    def audit_device(self, observations):
        """``(ids, near_tie)`` on the device: ``near_tie[i]`` marks the rows whose two largest logits are closer than a
        forward-error bound of the float32 dot products — the only rows where another float32 implementation of the same
        network (the reference's torch forward, whose sgemm sums in another order) can return a different id.  The ids are
        this package's own (bias first, then ``fmaf`` over the inputs in index order; bit-exact against the oracle)."""
        if self._w is None:
            raise ValueError("Model not initialized. Either train a new model for the encoder or load an existing one.")
        dev = _require_cuda(self._device)
        W1, b1, W2, b2 = self._w
        with torch.cuda.device(dev):
            obs = _as_tensor(observations, dev, torch.float32)
            if obs.dim() != 2 or obs.shape[1] != self.obs_dim:
                raise ValueError(f"expected observations of shape [N, {self.obs_dim}]")
            n = obs.shape[0]
            ids = torch.empty(n, dtype=torch.int32, device=dev)
            margin = torch.empty(n, dtype=torch.float32, device=dev)
            hnorm = torch.empty(n, dtype=torch.float32, device=dev)
            abi.check(abi.load().ols_encode_mlp_audit(_ptr(obs), n, self.obs_dim, _ptr(W1), _ptr(b1), self.hidden_size, _ptr(W2),
                                                      _ptr(b2), self.latent_size, C.c_float(self.negative_slope), _ptr(ids),
                                                      _ptr(margin), _ptr(hnorm), _stream_ptr()))
            # |error of a logit| <= gamma_{H+1} (|b2| + sum_j |h_j| |W2_kj|) + sum_j |dh_j| |W2_kj|, with the hidden units
            # themselves off by |dh_j| <= gamma_{dO+1} (|b1_j| + sum_d |x_d| |W1_jd|); both implementations err, hence the 2
            eps = 2.0 ** -24
            w2max, b2max = W2.abs().max(), b2.abs().max()
            layer1 = (obs.abs() @ W1.abs().t() + b1.abs()).sum(1)
            bound = 2.0 * eps * 1.01 * ((self.hidden_size + 1) * (b2max + hnorm * w2max) + (self.obs_dim + 1) * layer1 * w2max)
            near = margin <= 2.0 * bound            # the DIFFERENCE of two logits
        return ids, near
