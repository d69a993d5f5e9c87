"""The oracle's NumPy-RNG restatement (oracle/np_random.py) against NumPy itself."""
import numpy as np
import pytest

from oracle import np_random as R

SEEDS = [0, 1, 5, 99, 98052, 2**32 - 1, 2**32, 2**64 + 12345, 2**127 + 3]


@pytest.mark.parametrize("seed", SEEDS)
def test_seed_sequence_and_state(seed):
    assert list(np.random.SeedSequence(seed).generate_state(4, np.uint64)) == R.seed_sequence_state64(seed)
    st = np.random.default_rng(seed).bit_generator.state["state"]
    mine = R.PCG64(seed)
    assert (st["state"], st["inc"]) == (mine.state, mine.inc)


@pytest.mark.parametrize("seed", SEEDS)
def test_uniform_stream(seed):
    g = np.random.default_rng(seed)
    assert [g.random() for _ in range(257)] == R.uniform_stream(seed, 257)


@pytest.mark.parametrize("seed", SEEDS[:5])
@pytest.mark.parametrize("n", [0, 1, 2, 3, 7, 31, 32, 33, 100, 1000, 4097])
def test_shuffle_permutation(seed, n):
    x = list(range(n))
    np.random.default_rng(seed).shuffle(x)           # list path, as evaluators/psrs.py:23,30
    assert x == R.shuffle_permutation(seed, n)


def test_buffered_uint32_is_kept_between_calls():
    g, m = np.random.default_rng(4), R.PCG64(4)
    x = list(range(11)); g.shuffle(x)
    assert x == m.shuffle_permutation(11)
    assert g.random() == m.random()
    y = list(range(6)); g.shuffle(y)
    assert y == m.shuffle_permutation(6)
