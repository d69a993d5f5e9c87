"""Column-store of logged transitions — the part of offsim4rl/data.py the PSRS path consumes
(data.py:16-30 ``ProbDistribution``, :33-35 ``Transition``, :38-79 + :100-118 ``OfflineDataset``) and its
persistence (data.py:81-98 ``save_hdf5`` / ``load_hdf5``, :120-147 ``HDF5Dataset``).

The HDF5 layout is the reference's (one gzip dataset per experience key, the two spaces and the distribution
type as pickled ``np.void`` attributes), so files written by either package load in the other.  ``h5py`` is an
optional dependency: without it the HDF5 entry points raise ``ImportError`` and ``save_npz`` / ``load_npz`` keep
the same content in a NumPy archive.  ``device_columns`` is the ingest step of the simulator: the columns the
bucketing kernel consumes, as contiguous arrays of the dtypes ``ols_log_t`` takes."""
from __future__ import annotations

import enum
import logging
import pickle
from collections import namedtuple

import numpy as np


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

    # ------------------------------------------------------------------ persistence (data.py:81-98)
    @classmethod
    def load_hdf5(cls, path, group_name=None):
        return HDF5Dataset(path, group_name)

    def save_hdf5(self, path, group_name=None):
        h5py = _require_h5py()
        with h5py.File(path, "w") as fout:
            group = fout.create_group(group_name) if group_name else fout
            for name in ("observation_space", "action_space", "action_dist_type"):
                group.attrs.create(name, np.void(pickle.dumps(getattr(self, name))))
            for k, v in self.experience.items():
                group.create_dataset(k, data=v, compression="gzip")

    def save_npz(self, path):
        """Same content as ``save_hdf5`` in a NumPy archive (for hosts without h5py)."""
        meta = {name: np.void(pickle.dumps(getattr(self, name))) for name in ("observation_space", "action_space", "action_dist_type")}
        np.savez_compressed(path, **{f"__attr__{k}": v for k, v in meta.items()},
                            **{k.replace("/", "__slash__"): np.asarray(v) for k, v in self.experience.items()})

    @classmethod
    def load_npz(cls, path):
        with np.load(path, allow_pickle=False) as f:
            attrs = {k[len("__attr__"):]: pickle.loads(f[k].tobytes()) for k in f.files if k.startswith("__attr__")}
            exp = {k.replace("__slash__", "/"): f[k] for k in f.files if not k.startswith("__attr__")}
        return cls(attrs.get("observation_space"), attrs.get("action_space"),
                   attrs.get("action_dist_type", ProbDistribution.NoProbability), **exp)

    # ------------------------------------------------------------------ ingest: columns for ols_log_t
    def device_columns(self, z=None, z_next=None, info_prefix="infos/"):
        """The log columns of the PSRS path, ready for ``BucketedLog``: ``z``/``z_next`` int32 state ids, ``action`` int32,
        ``reward`` float32, ``terminal`` uint8, ``p_log`` [N, nA] in its own float dtype, ``steps`` int32 or None.

        State ids come from, in this order: the ``z`` / ``z_next`` arguments (e.g. an encoder's output), the recorded
        ``infos/z`` and ``infos/z_next`` (or ``infos/next_z``, psrs.py:17) columns that ``record_dataset_in_memory``
        writes for environments which reveal their state, or — for discrete observation spaces — the observations."""
        e = self.experience
        if "action_distributions" not in e:
            raise ValueError("the PSRS path needs the logged action_distributions")

        def pick(arg, *names):
            if arg is not None:
                return np.asarray(arg)
            for n in names:
                if info_prefix + n in e:
                    return np.asarray(e[info_prefix + n])
            return None

        zc, znc = pick(z, "z"), pick(z_next, "z_next", "next_z")
        if zc is None or znc is None:
            obs, nobs = np.asarray(e["observations"]), np.asarray(e["next_observations"])
            if obs.ndim != 1 or not np.issubdtype(obs.dtype, np.integer):
                raise ValueError("no state ids: pass z / z_next, record infos/z and infos/z_next, or use discrete observations")
            zc, znc = obs, nobs
        p = np.asarray(e["action_distributions"])
        if p.dtype not in (np.float32, np.float64):
            p = p.astype(np.float64)
        return dict(z=np.ascontiguousarray(zc, dtype=np.int32), z_next=np.ascontiguousarray(znc, dtype=np.int32),
                    action=np.ascontiguousarray(e["actions"], dtype=np.int32),
                    reward=np.ascontiguousarray(e["rewards"], dtype=np.float32),
                    terminal=np.ascontiguousarray(np.asarray(e["terminals"]).astype(bool), dtype=np.uint8),
                    p_log=np.ascontiguousarray(p),
                    steps=None if "steps" not in e else np.ascontiguousarray(e["steps"], dtype=np.int32))


def _require_h5py():
    try:
        import h5py
    except ImportError as exc:            # not in this image; the NPZ twins work everywhere
        raise ImportError("h5py is not installed: use OfflineDataset.save_npz / load_npz, or install h5py for HDF5 files") from exc
    return h5py


class HDF5Dataset(OfflineDataset):
    """An ``OfflineDataset`` whose experience columns are the open file's datasets (data.py:120-147)."""

    def __init__(self, path, group_name=None):
        h5py = _require_h5py()
        self._fin = h5py.File(path, "r")
        group = self._fin.get(group_name, default=self._fin) if group_name else self._fin
        attr = lambda name, default=None: (pickle.loads(group.attrs[name].tobytes()) if name in group.attrs else default)
        super().__init__(attr("observation_space"), attr("action_space"),
                         attr("action_dist_type", ProbDistribution.NoProbability), **{k: group[k] for k in group})

    def close(self):
        fin, self._fin = getattr(self, "_fin", None), None
        if fin is not None:
            fin.close()

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.close()

    def __del__(self):
        self.close()
