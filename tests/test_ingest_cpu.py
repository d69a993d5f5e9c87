"""SURVEY §8f rank 4 — ingest adapters: the recorder's column layout against vectors produced by the unmodified
reference (tests/golden/make_golden.py::case_recorder), persistence round trips, and the columns handed to the
bucketing kernel.  CPU-only."""
import pickle

import numpy as np
import pytest

from tests.helpers import ToyChainEnv, ToyAgent


@pytest.fixture(scope="module")
def pkg():
    import offsim4rl_b200 as ob        # the package imports without CUDA; only kernels need the library
    return ob


@pytest.mark.parametrize("tag,kw", [("w0", dict(seed=5, worker_id=0, workers_num=1)), ("w2of3", dict(seed=8, worker_id=2, workers_num=3))])
def test_recorder_matches_reference_columns(pkg, golden, tag, kw):
    from offsim4rl_b200.utils import record_dataset_in_memory
    from offsim4rl_b200.data import ProbDistribution
    g = golden("recorder")
    np.random.seed(123)
    agent = ToyAgent(ProbDistribution.Discrete)
    ds = record_dataset_in_memory(ToyChainEnv(), agent, 300, new_step_api=True, include_infos=True, **kw)
    want = {k[len(tag) + 2:].replace("__", "/"): g[k] for k in g.files if k.startswith(tag + "__")}
    committed, ended = want.pop("committed"), want.pop("ended")
    assert set(ds.experience) == set(want)                     # infos/label (str) and infos/late (absent at first) are skipped
    for k, v in want.items():
        got = np.asarray(ds.experience[k])
        assert got.dtype == v.dtype and np.array_equal(got, v), k
    assert np.array_equal(agent.committed, committed) and np.array_equal(np.array(agent.ended, dtype=np.float64), ended)
    assert ds.action_dist_type == ProbDistribution.Discrete and len(ds) == 300


def test_recorder_rejects_old_api_when_new_is_required(pkg):
    from offsim4rl_b200.utils import record_dataset_in_memory
    from offsim4rl_b200.data import ProbDistribution

    class OldApi(ToyChainEnv):
        def step(self, a):
            o, r, term, trunc, info = super().step(a)
            return o, r, term or trunc, info

    with pytest.raises(ValueError):
        record_dataset_in_memory(OldApi(), ToyAgent(ProbDistribution.Discrete), 5, seed=1, new_step_api=True)
    np.random.seed(0)
    ds = record_dataset_in_memory(OldApi(), ToyAgent(ProbDistribution.Discrete), 40, seed=1, new_step_api=False, include_infos=False)
    assert not any(k.startswith("infos/") for k in ds.experience) and ds.experience["terminals"].any()


def _recorded(pkg):
    from offsim4rl_b200.utils import record_dataset_in_memory
    from offsim4rl_b200.data import ProbDistribution
    np.random.seed(7)
    return record_dataset_in_memory(ToyChainEnv(), ToyAgent(ProbDistribution.Discrete), 500, seed=3)


def test_npz_round_trip_and_hdf5_gate(pkg, tmp_path):
    from offsim4rl_b200.data import OfflineDataset
    ds = _recorded(pkg)
    path = tmp_path / "log.npz"
    ds.save_npz(path)
    back = OfflineDataset.load_npz(path)
    assert set(back.experience) == set(ds.experience)
    for k in ds.experience:
        assert np.array_equal(back.experience[k], ds.experience[k]) and back.experience[k].dtype == np.asarray(ds.experience[k]).dtype
    assert back.action_dist_type == ds.action_dist_type and back.action_space.n == 3
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            ds.save_hdf5(tmp_path / "log.h5")
        with pytest.raises(ImportError):
            OfflineDataset.load_hdf5(tmp_path / "log.h5")
    else:
        ds.save_hdf5(tmp_path / "log.h5")
        with OfflineDataset.load_hdf5(tmp_path / "log.h5") as h:
            assert np.array_equal(h.experience["actions"][:], ds.experience["actions"])
            blob = h._fin.attrs["action_dist_type"]           # the reference's attribute encoding (data.py:96-98)
            assert pickle.loads(blob.tobytes()) == ds.action_dist_type


def test_hdf5_paths_execute_against_a_stand_in_h5py(pkg, tmp_path, monkeypatch):
    """h5py is not in this image, so save_hdf5 / load_hdf5 / HDF5Dataset would never run here: execute them against the
    stand-in module of tests/helpers.py (the same h5py calls the reference's data.py:81-98,120-147 makes) — root group and a
    named group, the pickled-np.void attribute encoding, context-manager close — and bucket-ready columns from the file."""
    import sys
    from tests.helpers import make_fake_h5py
    from offsim4rl_b200.data import OfflineDataset, HDF5Dataset
    try:
        import h5py  # noqa: F401
    except ImportError:
        monkeypatch.setitem(sys.modules, "h5py", make_fake_h5py())
    ds = _recorded(pkg)
    for group in (None, "run0"):
        path = tmp_path / f"log_{group}.h5"
        ds.save_hdf5(path, group_name=group)
        with OfflineDataset.load_hdf5(path, group_name=group) as h:
            assert isinstance(h, HDF5Dataset) and set(h.experience) == set(ds.experience)
            for k in ds.experience:
                assert np.array_equal(np.asarray(h.experience[k][:]), np.asarray(ds.experience[k]))
            assert h.action_dist_type == ds.action_dist_type and h.action_space.n == ds.action_space.n
            assert h.observation_space.shape == ds.observation_space.shape
            cols = h.device_columns()
            assert np.array_equal(cols["z"], ds.experience["infos/z"]) and cols["p_log"].shape == (500, 3)
        assert h._fin is None                                      # closed by the context manager


def test_device_columns_sources_and_dtypes(pkg):
    from offsim4rl_b200.data import OfflineDataset, ProbDistribution
    ds = _recorded(pkg)
    cols = ds.device_columns()                                 # ids from the recorded infos/z, infos/z_next
    assert np.array_equal(cols["z"], ds.experience["infos/z"]) and np.array_equal(cols["z_next"], ds.experience["infos/z_next"])
    assert [cols[k].dtype for k in ("z", "z_next", "action", "reward", "terminal", "steps")] == \
        [np.int32, np.int32, np.int32, np.float32, np.uint8, np.int32]
    assert cols["p_log"].dtype == np.float64 and cols["p_log"].shape == (500, 3) and cols["p_log"].flags.c_contiguous
    ids = (np.arange(500) % 4).astype(np.int64)
    assert np.array_equal(ds.device_columns(z=ids, z_next=ids[::-1])["z_next"], ids[::-1])
    # `next_z` spelling (psrs.py:17) and discrete observations as ids
    e = dict(ds.experience)
    e["infos/next_z"] = e.pop("infos/z_next")
    assert np.array_equal(OfflineDataset(ds.observation_space, ds.action_space, ProbDistribution.Discrete, **e).device_columns()["z_next"], cols["z_next"])
    disc = OfflineDataset(None, ds.action_space, ProbDistribution.Discrete, observations=cols["z"].astype(np.int64),
                          next_observations=cols["z_next"].astype(np.int64), actions=e["actions"], rewards=e["rewards"],
                          terminals=e["terminals"], action_distributions=e["action_distributions"])
    c2 = disc.device_columns()
    assert np.array_equal(c2["z"], cols["z"]) and c2["steps"] is None
    e.pop("infos/z")
    with pytest.raises(ValueError):
        OfflineDataset(ds.observation_space, ds.action_space, ProbDistribution.Discrete, **e).device_columns()
    e.pop("action_distributions")
    with pytest.raises(ValueError):
        OfflineDataset(ds.observation_space, ds.action_space, ProbDistribution.Discrete, **e).device_columns(z=ids, z_next=ids)
