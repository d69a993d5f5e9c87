"""The round geometry of the streaming evaluator (csrc/ols_eval.cu), restated in Python and checked against the sequential
definition it must reproduce (psrs.py:39-51: pop the queue front to back, one uniform per pop):

    record i of a queue visit sits at hot[lo + (start + cur + i) mod len] and consumes draw `draws + i`
    (stream protocol v2: block (draws + i) >> 2, word (draws + i) & 3).

The kernel walks the visit in rounds of up to 128 U records whose lanes are aligned to the 32-byte sectors of the hot array
while the draws are aligned to their own blocks of four; a lane takes the words its block lacks from the next lane (or, for the
last lane, from lane 0's block of the next sector row).  The GPU suite checks the result draw for draw against the C oracle;
this is the index arithmetic on its own, over random and extreme geometries, on the CPU."""
import numpy as np
import pytest


def rounds(lo, start, length, cur, draws, U):
    """Yield, per round, the list of (hot index, draw, source) triples the kernel examines, in queue order."""
    while True:
        remaining = length - cur
        if remaining <= 0:
            return
        pos0 = start + cur
        if pos0 >= length:
            pos0 -= length
        P = lo + pos0
        mis = P & 3
        D1 = draws - mis                       # >= -3
        s = D1 & 3                             # (two's complement, as on the device)
        n_eval = min(remaining, length - pos0, 128 * U - mis - s)
        assert n_eval > 0, "a round must make progress"
        A = P - mis
        out = []
        for u in range(U):
            for lane in range(32):
                g = 4 * (lane + 32 * u)
                k0 = g - mis
                block = (D1 >> 2) + lane + 32 * u          # the block this lane computes for this sector row
                for i in range(4):
                    k = k0 + i
                    if not (0 <= k < n_eval):
                        continue
                    c = s + i                               # word index counted from the lane's own block
                    if c < 4:
                        src = (block, c)
                    else:                                   # the neighbour's block
                        assert lane < 31 or u + 1 < U, "the last lane of the last row has no neighbour"
                        src = (block + 1, c - 4)
                    out.append((A + g + i, draws + k, src))
        out.sort(key=lambda t: t[1])
        yield out, n_eval
        cur += n_eval
        draws += n_eval


@pytest.mark.parametrize("U", [1, 2, 4])
def test_rounds_reproduce_the_sequential_visit(U):
    rng = np.random.default_rng(U)
    cases = [(0, 0, 1, 0, 0), (3, 0, 5, 0, 7), (1, 4, 5, 0, 2), (7, 999, 1000, 0, 3), (5, 0, 1000, 997, 1)]
    for _ in range(300):
        length = int(rng.integers(1, 1500))
        cases.append((int(rng.integers(0, 10_000)), int(rng.integers(0, length)), length, int(rng.integers(0, length)),
                      int(rng.integers(0, 1 << 40))))
    for lo, start, length, cur, draws in cases:
        seen = []
        for out, n_eval in rounds(lo, start, length, cur, draws, U):
            assert len(out) == n_eval
            seen.extend(out)
        want = length - cur
        assert len(seen) == want
        for i, (idx, d, (block, word)) in enumerate(seen):
            assert idx == lo + (start + cur + i) % length        # queue order, ring wrap included
            assert lo <= idx < lo + length
            assert d == draws + i                                # one draw per pop, in order
            assert (block, word) == (d >> 2, d & 3)              # the word the protocol assigns to that draw


def test_misalignment_is_constant_along_a_chain_of_full_rounds():
    # what lets the kernel treat `s` as a property of the visit: a full round advances cursor and draws by the same amount
    for U in (1, 2):
        for mis0 in range(4):
            for d0 in range(8):
                lo, start, length, cur = mis0, 0, 100_000, 0
                ss = []
                for out, n_eval in rounds(lo, start, length, cur, d0, U):
                    first_idx, first_draw, _ = out[0]
                    ss.append((first_draw - (first_idx & 3)) & 3)
                    if len(ss) == 5:
                        break
                assert len(set(ss)) == 1
