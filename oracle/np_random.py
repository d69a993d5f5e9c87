"""Restatement of the NumPy random machinery the reference PSRS path relies on.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

The reference calls, on its hot path,
  * ``np.random.default_rng(seed)``            (offsim4rl/evaluators/psrs.py:20,23,30)
  * ``Generator.shuffle(list)``                (psrs.py:23,30)
  * ``Generator.random()``                     (psrs.py:56)
NumPy is a third-party dependency that is not vendored in /root/reference and is
not pinned in its requirements.txt; the container has NumPy 2.3.5.  The published
algorithms restated here: SeedSequence (O'Neill's seed_seq_fe128 variant, as
documented in numpy/random/bit_generator.pyx), PCG64 = PCG XSL-RR 128/64
(pcg-random.org; numpy/random/src/pcg64), ``random()`` = 53-bit mantissa from the
upper bits of one 64-bit draw, and ``shuffle`` = descending Fisher-Yates with the
masked-rejection bounded integer (numpy/random/_generator.pyx ``_shuffle_raw`` +
``random_interval`` in distributions.c).  Everything is plain Python integers so
it can be diffed against NumPy bit for bit (tests/test_oracle_rng.py).
"""
from __future__ import annotations

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF
M128 = (1 << 128) - 1

_INIT_A = 0x43B0D7E5
_MULT_A = 0x931E8875
_INIT_B = 0x8B51F9DD
_MULT_B = 0x58F38DED
_MIX_L = 0xCA01F9DD
_MIX_R = 0x4973F715
_XSHIFT = 16

PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645


def _entropy_words(seed: int):
    """Non-negative python int -> little-endian uint32 words ([0] for 0)."""
    if seed < 0:
        raise ValueError("seed must be non-negative")
    words = []
    while True:
        words.append(seed & M32)
        seed >>= 32
        if seed == 0:
            break
    return words


def seed_sequence_pool(seed: int):
    """SeedSequence(seed).pool  (4 x uint32)."""
    ent = _entropy_words(seed)
    hash_const = [_INIT_A]

    def hashmix(v):
        v = (v ^ hash_const[0]) & M32
        hash_const[0] = (hash_const[0] * _MULT_A) & M32
        v = (v * hash_const[0]) & M32
        v ^= v >> _XSHIFT
        return v

    def mix(x, y):
        r = ((_MIX_L * x) - (_MIX_R * y)) & M32
        r ^= r >> _XSHIFT
        return r

    pool = [0, 0, 0, 0]
    for i in range(4):
        pool[i] = hashmix(ent[i] if i < len(ent) else 0)
    for i_src in range(4):
        for i_dst in range(4):
            if i_src != i_dst:
                pool[i_dst] = mix(pool[i_dst], hashmix(pool[i_src]))
    for i_src in range(4, len(ent)):
        for i_dst in range(4):
            pool[i_dst] = mix(pool[i_dst], hashmix(ent[i_src]))
    return pool


def seed_sequence_state64(seed: int, n_words64: int = 4):
    """SeedSequence(seed).generate_state(n_words64, uint64)."""
    pool = seed_sequence_pool(seed)
    hb = _INIT_B
    out32 = []
    for i in range(2 * n_words64):
        d = pool[i % 4]
        d ^= hb
        hb = (hb * _MULT_B) & M32
        d = (d * hb) & M32
        d ^= d >> _XSHIFT
        out32.append(d)
    return [out32[2 * i] | (out32[2 * i + 1] << 32) for i in range(n_words64)]


class PCG64:
    """``np.random.PCG64(seed)`` for a non-negative integer seed."""

    def __init__(self, seed: int):
        w = seed_sequence_state64(seed, 4)
        initstate = (w[0] << 64) | w[1]
        initseq = (w[2] << 64) | w[3]
        self.inc = ((initseq << 1) | 1) & M128
        self.state = 0
        self._step()
        self.state = (self.state + initstate) & M128
        self._step()
        self.has_uint32 = False
        self.uinteger = 0

    def _step(self):
        self.state = (self.state * PCG_MULT + self.inc) & M128

    def next_uint64(self) -> int:
        self._step()
        hi = self.state >> 64
        lo = self.state & M64
        x = hi ^ lo
        rot = self.state >> 122
        return ((x >> rot) | (x << ((-rot) & 63))) & M64

    def next_uint32(self) -> int:
        if self.has_uint32:
            self.has_uint32 = False
            return self.uinteger
        n = self.next_uint64()
        self.has_uint32 = True
        self.uinteger = n >> 32
        return n & M32

    def random(self) -> float:
        """``Generator.random()``: (next_uint64 >> 11) * 2**-53."""
        return (self.next_uint64() >> 11) * (1.0 / 9007199254740992.0)

    def interval(self, mx: int) -> int:
        """``random_interval(bitgen, max)``: uniform integer in [0, max]."""
        if mx == 0:
            return 0
        mask = mx
        mask |= mask >> 1
        mask |= mask >> 2
        mask |= mask >> 4
        mask |= mask >> 8
        mask |= mask >> 16
        mask |= mask >> 32
        if mx <= M32:
            while True:
                v = self.next_uint32() & mask
                if v <= mx:
                    return v
        while True:
            v = self.next_uint64() & mask
            if v <= mx:
                return v

    def shuffle_permutation(self, n: int):
        """Indices ``perm`` such that ``Generator.shuffle(x)`` leaves ``x_new[k] = x_old[perm[k]]``."""
        perm = list(range(n))
        for i in range(n - 1, 0, -1):
            j = self.interval(i)
            perm[i], perm[j] = perm[j], perm[i]
        return perm


def shuffle_permutation(seed: int, n: int):
    """Permutation applied by ``np.random.default_rng(seed).shuffle`` to a length-n sequence."""
    return PCG64(seed).shuffle_permutation(n)


def uniform_stream(seed: int, count: int):
    """First ``count`` values of ``np.random.default_rng(seed).random()``."""
    g = PCG64(seed)
    return [g.random() for _ in range(count)]
