"""Pure-python restatements of the product's counter-based streams (Philox4x32-7 stream protocol v2, ring offsets) as
documented in include/ols_b200.h / DESIGN.md — shared by CPU and GPU tests."""

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF


import numpy as np


def philox4x32(ctr, k0, k1, rounds=7):
    c = list(ctr)
    for _ in range(rounds):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c[3] ^ k1) & M32, p0 & M32]
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c


def stream_word(seed, stream, which, d):
    """Word d & 3 of block d >> 2 of stream `which` (protocol v2, csrc/ols_common.cuh): Philox4x32-7."""
    blk = d >> 2
    c = philox4x32([blk & M32, (blk >> 32) & M32, stream & M32, ((stream >> 32) & M32) ^ ((which << 28) & M32)],
                   seed & M32, (seed >> 32) & M32)
    return c[d & 3]


def philox_uniform(seed, stream, which, d):
    """Rejection uniform: 32 high bits from stream `which`, 21 low bits from stream `which + 2`."""
    hi, lo = stream_word(seed, stream, which, d), stream_word(seed, stream, which + 2, d) >> 11
    return ((hi << 21) | lo) * (1.0 / 9007199254740992.0)


def philox_tie(seed, stream, t):
    """Arg-max tie-break value: the high word of stream 1 alone."""
    return stream_word(seed, stream, 1, t) * (1.0 / 4294967296.0)


def fmix32(h):
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & M32
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & M32
    h ^= h >> 16
    return h


def ring_replica_hash(ring_seed, replica):
    h = fmix32((replica >> 32) & M32)
    h = fmix32((replica & M32) ^ h)
    h = fmix32(((ring_seed >> 32) & M32) ^ h)
    return fmix32((ring_seed & M32) ^ h)


def ring_start(ring_seed, replica, bucket, length):
    h = fmix32(ring_replica_hash(ring_seed, replica) ^ (((bucket + 1) * 0x9E3779B1) & M32))
    return (h * length) >> 32


def init_ring_start(ring_seed, replica, n_init):
    return (fmix32(ring_replica_hash(ring_seed, replica) ^ 0x85EBCA6B) * n_init) >> 32


# --------------------------------------------------------------------------- toy environment / agent for the recorder
class _Space:
    def __init__(self, n=None, shape=None):
        self.n, self.shape = n, shape


class ToyChainEnv:
    """Deterministic (given its seed) new-API gym-style environment: a ring of `n` cells, observation = 2 floats,
    info reveals the cell id `z`, a float, an array, a string (unsupported type) and a key that is absent at first."""

    def __init__(self, n=7, horizon=9):
        self.n, self.horizon = n, horizon
        self.observation_space, self.action_space = _Space(shape=(2,)), _Space(n=3)
        self.rng = np.random.default_rng(0)
        self.cell = self.t = 0
        self.calls = 0

    def _obs(self):
        return np.array([self.cell / self.n, self.t / self.horizon], dtype=np.float32)

    def reset(self, seed=None):
        if seed is not None:
            self.rng = np.random.default_rng(seed)
        self.cell, self.t = int(self.rng.integers(self.n)), 0
        return self._obs()

    def step(self, action):
        self.calls += 1
        z = self.cell
        self.cell = (self.cell + int(action) - 1 + (1 if self.rng.random() < 0.2 else 0)) % self.n
        self.t += 1
        terminated = self.cell == 0
        truncated = self.t >= self.horizon
        info = {"z": z, "z_next": self.cell, "noise": float(self.rng.normal()), "vec": np.array([z, self.cell]), "label": "x"}
        if self.calls > 1:
            info["late"] = 1.0
        return self._obs(), float(self.cell == 0) - 0.1, terminated, truncated, info


class ToyAgent:
    """Observation-dependent fixed distributions; remembers what the recorder told it."""

    def __init__(self, dist_type):
        self.action_dist_type = dist_type
        self.committed, self.ended = [], []

    def _dist(self, obs):
        w = np.array([1.0 + obs[0], 1.0, 2.0 - obs[1]])
        return w / w.sum()

    def begin_episode(self, observation):
        return self._dist(observation)

    def step(self, prev_reward, observation):
        return self._dist(observation)

    def commit_action(self, action):
        self.committed.append(int(action))

    def end_episode(self, reward, truncated=False):
        self.ended.append((float(reward), bool(truncated)))


# --------------------------------------------------------------------------- a stand-in for h5py (absent from this image)
def make_fake_h5py():
    """A minimal module with the part of h5py's API that OfflineDataset.save_hdf5 / HDF5Dataset (and the reference's
    data.py:81-98,120-147) use: File(path, mode) as a context manager, create_group / get, group.attrs.create / [] / in,
    create_dataset(name, data=, compression=), iteration over member names and [] returning something sliceable.  Files
    are pickles; the point is to EXECUTE the HDF5 code paths of the package, which the missing dependency otherwise gates."""
    import pickle
    import types

    class Attrs(dict):
        def create(self, name, data):
            self[name] = np.void(bytes(data.tobytes()))

    class Group:
        def __init__(self):
            self.attrs, self.members = Attrs(), {}

        def create_group(self, name):
            self.members[name] = Group()
            return self.members[name]

        def create_dataset(self, name, data=None, compression=None):
            assert compression in (None, "gzip")
            self.members[name] = np.array(data)
            return self.members[name]

        def get(self, name, default=None):
            return self.members.get(name, default)

        def __getitem__(self, name):
            return self.members[name]

        def __iter__(self):
            return iter(self.members)

        def __contains__(self, name):
            return name in self.members

        def plain(self):
            return {"attrs": {k: v.tobytes() for k, v in self.attrs.items()},
                    "members": {k: (v.plain() if isinstance(v, Group) else v) for k, v in self.members.items()}}

        def fill(self, tree):
            for k, v in tree["attrs"].items():
                self.attrs[k] = np.void(v)
            for k, v in tree["members"].items():
                self.members[k] = Group().fill(v) if isinstance(v, dict) else v
            return self

    class File(Group):
        def __init__(self, path, mode="r"):
            super().__init__()
            self.path, self.mode, self.closed = str(path), mode, False
            if mode == "r":
                with open(self.path, "rb") as f:
                    self.fill(pickle.load(f))

        def close(self):
            if not self.closed and self.mode == "w":
                with open(self.path, "wb") as f:
                    pickle.dump(self.plain(), f)
            self.closed = True

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            self.close()

    mod = types.ModuleType("h5py")
    mod.File, mod.Group = File, Group
    mod.__fake__ = True
    return mod
