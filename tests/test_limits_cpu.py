"""The per-step accept/reject limits of the TAIL hot-record kernels (csrc/ols_sim_kernels.cuh `make_limits`), restated in
NumPy with the same float32 operations and directed roundings, checked against the exact rule psrs.py:53-57 in float64:

    hi <  lim.x  must imply  u <  thr   for EVERY u whose high word is hi  (u in [hi, hi + 1) * 2^-32)
    hi >  lim.y  must imply  u >  thr

so the integer compares of the pop loop can never take a decision the exact rule would not take; whatever falls between the
limits goes to the exact rule on the device.  The GPU suite checks the same thing end to end (draw-for-draw equality with the
C oracle over ~1e8 pops); this is the arithmetic on its own, over millions of random and extreme cases, on the CPU."""
import numpy as np

F32 = np.float32
LO_SCALE, HI_SCALE = F32(4294836224.0), F32(4295098368.0)       # (1 - 2^-15) 2^32 and (1 + 2^-15) 2^32, exact in float32


def _mul_dir(a, b, up):
    """float32 product rounded toward -inf / +inf (`__fmul_rd` / `__fmul_ru`): the float64 product of two float32 is exact."""
    exact = a.astype(np.float64) * b.astype(np.float64)
    r = exact.astype(F32)
    with np.errstate(over="ignore", invalid="ignore"):
        wrong = (r.astype(np.float64) < exact) if up else (r.astype(np.float64) > exact)
        step = np.nextafter(r, F32(np.inf) if up else F32(-np.inf))
    return np.where(wrong, step, r)


def make_limits(p, pi):
    """p [n, 4] float32 logged distribution of the queue, pi [n, 4] float64 agent distribution -> (lo32, hi32) uint64 [n, 4]."""
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        pj = pi.astype(F32)
        ipn = np.where(pj == 0, F32(np.inf), F32(1) / pj)                       # __frcp_rn
        t = p * ipn                                                             # float32, round to nearest
        inv_m = t.min(axis=1, keepdims=True)
        thr = inv_m / t                                                         # __fdiv_rn
        lo = _mul_dir(thr, np.broadcast_to(LO_SCALE, thr.shape), up=False)
        hi = _mul_dir(thr, np.broadcast_to(HI_SCALE, thr.shape), up=True)
        lo32 = np.clip(np.floor(lo.astype(np.float64)), 0, 4294967295).astype(np.uint64)     # saturating __float2uint_rd
        hi32 = np.clip(np.ceil(hi.astype(np.float64)), 0, 4294967295).astype(np.uint64)      # saturating __float2uint_ru
    fine = ~np.isnan(thr)
    lo32 = np.where(fine, np.where(lo32 > 2, lo32 - 2, 0), 0)
    hi32 = np.where(fine, np.where(hi32 < 0xFFFFFFFD, hi32 + 2, 0xFFFFFFFF), 0xFFFFFFFF)
    return lo32, hi32


def exact_thresholds(p, pi):
    """psrs.py:55-57 in float64: thr_a = (pi_a / p_a) / max_j (pi_j / p_j), for every action a."""
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = pi / p.astype(np.float64)
        return ratio / ratio.max(axis=1, keepdims=True)


def _check(p, pi):
    lo32, hi32 = make_limits(p, pi)
    thr = exact_thresholds(p, pi)
    # accept below lo32: the LARGEST u with high word lo32 - 1 is just under lo32 * 2^-32
    assert np.all(lo32.astype(np.float64) * 2.0 ** -32 <= thr), "an accept limit lies above the exact threshold"
    # reject above hi32: the SMALLEST u with high word hi32 + 1 is (hi32 + 1) * 2^-32
    rej = hi32 < 0xFFFFFFFF
    assert np.all(((hi32.astype(np.float64) + 1) * 2.0 ** -32 > thr)[rej]), "a reject limit lies below the exact threshold"
    return lo32, hi32, thr


def test_limits_bracket_the_exact_threshold_random_policies():
    rng = np.random.default_rng(0)
    n = 400_000
    pb = np.maximum(rng.dirichlet(np.ones(4), size=n), 0.05)
    p = (pb / pb.sum(1, keepdims=True)).astype(F32)
    for pi in (rng.dirichlet(np.ones(4), size=n),                                               # arbitrary policies
               np.full((n, 4), 0.25),                                                           # uniform target (C5)
               np.where(np.eye(4)[rng.integers(0, 4, n)] > 0, 1 - 0.1 + 0.025, 0.025)):         # epsilon-greedy, eps = 0.1 (C4)
        lo32, hi32, thr = _check(p, pi)
        # the band is narrow: what goes to the exact rule is ~2^-14 of the unit interval per action
        width = (hi32.astype(np.float64) - lo32.astype(np.float64)) * 2.0 ** -32
        assert np.all(width <= thr * 2.0 ** -13 + 8 * 2.0 ** -32)


def test_limits_bracket_the_exact_threshold_extreme_cases():
    rng = np.random.default_rng(1)
    n = 300_000
    # tame distributions span 2^-40 .. 2^40 per entry (unnormalised on purpose), policies 2^-40 .. 1 with exact zeros
    p = np.exp2(rng.uniform(-40, 40, (n, 4))).astype(F32)
    pi = np.exp2(rng.uniform(-40, 0, (n, 4)))
    pi[rng.random((n, 4)) < 0.2] = 0.0
    pi[pi.max(1) == 0, 0] = 1.0                                    # (an all-zero policy fails the guard and never gets here)
    lo32, hi32, thr = _check(p, pi)
    zero = pi == 0
    assert np.all(lo32[zero] == 0) and np.all(hi32[zero] <= 2)     # pi_a = 0: thr = 0, only u = 0 can accept (exact rule)
    # the action that attains the max ratio has thr = 1: never rejected by the limits
    top = thr == 1.0
    assert np.all(hi32[top] == 0xFFFFFFFF)
    # near-deterministic behaviour policy (C5), uniform target
    d = 1e-3
    pb = np.full((1000, 4), d / 3)
    pb[np.arange(1000), rng.integers(0, 4, 1000)] = 1 - d
    _check(pb.astype(F32), np.full((1000, 4), 0.25))


def test_a_failed_guard_sends_every_pop_to_the_exact_rule():
    p = np.array([[0.5, 0.5, 1.0, 1.0]], dtype=F32)
    lo32, hi32 = make_limits(p, np.zeros((1, 4)))                   # inv_m = +inf, thr = NaN
    assert np.all(lo32 == 0) and np.all(hi32 == 0xFFFFFFFF)
