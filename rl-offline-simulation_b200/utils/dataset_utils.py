"""Collecting a log for offline simulation (offsim4rl/utils/dataset_utils.py:16-32, :34-136).

``record_dataset_in_memory`` drives an ``Agent`` through a gym-style environment and returns the columns the
PSRS path ingests, in the reference's key layout: ``episode_ids, steps, observations, actions,
action_distributions, rewards, next_observations, terminals`` plus one ``infos/<key>`` column per scalar/array
entry of the FIRST step's ``info`` dict (e.g. ``infos/z`` for environments that reveal their state).  This is host
code by nature (it calls the user's environment); the simulator's own work starts at
``OfflineDataset.device_columns``."""
from __future__ import annotations

import logging
from collections import defaultdict

import numpy as np

from ..data import OfflineDataset
from .prob_utils import sample_dist

_INFO_TYPES = (int, float, np.ndarray)


def get_keys(h5file):
    import h5py
    keys = []
    h5file.visititems(lambda name, item: keys.append(name) if isinstance(item, h5py.Dataset) else None)
    return keys


def load_h5_dataset(h5path):
    """Every dataset of an HDF5 file as a dict of arrays (scalars as 0-d reads)."""
    from ..data import _require_h5py
    h5py = _require_h5py()
    out = {}
    with h5py.File(h5path, "r") as f:
        for k in get_keys(f):
            try:
                out[k] = f[k][:]
            except ValueError:
                out[k] = f[k][()]
    return out


def record_dataset_in_memory(env, agent, num_samples, seed=None, worker_id=0, workers_num=1, new_step_api=True,
                             include_infos=True):
    cols = defaultdict(list)
    episode_id, t, reward = worker_id, 0, None
    obs = env.reset(seed=seed)
    info_keys = None
    for _ in range(num_samples):
        action_dist = agent.begin_episode(obs) if t == 0 else agent.step(reward, obs)
        action = sample_dist(action_dist)
        agent.commit_action(action)
        result = env.step(action)
        if len(result) == 4:            # old gym API: `done` cannot tell termination from truncation
            if new_step_api:
                raise ValueError("new_step_api is enabled, but the environment seems return just 'done' instead of "
                                 "'terminated' and 'truncated'. Use different environment or set new_step_api to False.")
            next_obs, reward, terminated, info = result
            truncated = False
        else:
            next_obs, reward, terminated, truncated, info = result
        row = dict(episode_ids=episode_id, steps=t, observations=obs, actions=action, action_distributions=action_dist,
                   rewards=reward, next_observations=next_obs, terminals=terminated)
        if include_infos:
            if info_keys is None:       # only keys present (with a supported type) in the first step are recorded
                info_keys = [k for k, v in info.items() if isinstance(v, _INFO_TYPES)]
            for k in info_keys:
                if k not in info:
                    row[f"infos/{k}"] = None
                elif isinstance(info[k], _INFO_TYPES):
                    row[f"infos/{k}"] = info[k]
                else:
                    logging.warning(f"Example's info {k} is not of a supported type. Skipping.")
        for k, v in row.items():
            cols[k].append(v)
        if terminated or truncated:
            agent.end_episode(reward, truncated=truncated)
            obs = env.reset()
            t = 0
            episode_id += workers_num
        else:
            obs, t = next_obs, t + 1
    return OfflineDataset(observation_space=env.observation_space, action_space=env.action_space,
                          action_dist_type=agent.action_dist_type, **{k: np.array(v) for k, v in cols.items()})
