"""CPU emulation of the device's per-pair arithmetic (swiftest_b200/csrc/kick_math.cuh): the integer high-word
manipulations, the one-compare fast-path test and the six-instruction inverse-cube refinement are restated bit for bit in
numpy (MUFU.RSQ replaced by a correctly rounded float32 1/sqrt perturbed by its documented error bound) and checked over
the whole exponent range.  This pins the constants and the error budget of the fast path without a GPU; the kernels
themselves are compared with the oracle in tests/test_gpu_parity.py.
"""
import numpy as np

SEED_HI_MIN = 0x38100000
SEED_HI_MAX = 0x47F00000


def hi_word(x):
    return (np.asarray(x, np.float64).view(np.uint64) >> np.uint64(32)).astype(np.uint32)


def seed_threshold(rlim2):
    t = (int(hi_word(np.float64(rlim2))) + 1) & 0xFFFFFFFF
    t = min(max(t, SEED_HI_MIN), SEED_HI_MAX)
    return np.uint32(t), np.uint32(SEED_HI_MAX - t)


def rcube_seeded(r2, thr, span, upper=True, rng=None):
    """kick_math.cuh::rcube_seeded; returns (value, hy)."""
    r2 = np.asarray(r2, np.float64)
    hi = hi_word(r2)
    with np.errstate(over="ignore"):
        ok = ((hi - thr) < span) if upper else (hi >= thr)                     # unsigned wrap-around compare
        fb = ((hi << np.uint32(3)) - np.uint32(0xC0000000)).astype(np.uint32)  # exponent re-biased by 1023-127
    f = fb.view(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        y0f = (np.float32(1.0) / np.sqrt(f.astype(np.float64))).astype(np.float32)
    if rng is not None:  # MUFU.RSQ: 2^-22.x relative error; emulate the worst case with a random 2^-22 perturbation
        y0f = (y0f.astype(np.float64) * (1.0 + rng.uniform(-2.0 ** -22, 2.0 ** -22, y0f.shape))).astype(np.float32)
    yb = y0f.view(np.uint32)
    hy = np.where(ok, (yb >> np.uint32(3)) + np.uint32(0x38000000), np.uint32(0)).astype(np.uint32)
    s = ((hy.astype(np.uint64) << np.uint64(32)) | yb.astype(np.uint64)).view(np.float64)
    with np.errstate(over="ignore", invalid="ignore", under="ignore"):
        s2 = s * s
        e = 1.0 - r2 * s2          # one rounding more than the device's fma: pessimistic
        s3 = s2 * s
        q = 1.875 * e + 1.5
        se = s3 * e
        val = se * q + s3
    return val, hy


def test_seed_constants_cover_exactly_the_normal_float_range():
    # smallest r2 taken: 2^-126 (smallest normal float); first r2 refused at the top: 2^128
    thr, span = seed_threshold(0.0)
    for r2, want in ((2.0 ** -126, True), (np.nextafter(2.0 ** -126, 0.0), False), (np.nextafter(2.0 ** 128, 0.0), True),
                     (2.0 ** 128, False), (0.0, False), (5e-324, False), (np.inf, False), (np.nan, False), (1.0, True)):
        _, hy = rcube_seeded(np.array([r2]), thr, span)
        assert (hy[0] != 0) == want, r2
    # the re-biased float has the same value as r2 truncated to 20 mantissa bits
    r2 = np.array([3.7, 1e-20, 1e20, 2.0 ** -100])
    hi = hi_word(r2)
    f = ((hi << np.uint32(3)) - np.uint32(0xC0000000)).astype(np.uint32).view(np.float32).astype(np.float64)
    assert np.all(f <= r2) and np.all(f > r2 * (1 - 2.0 ** -19))


def test_radius_threshold_is_conservative_and_tight():
    """A pair passes the single compare only if r2 > rlim2 (never a false accept); it may refuse pairs up to one high-word
    step (2^-20 relative) above rlim2: those go to the exact redo path."""
    rng = np.random.default_rng(1)
    for _ in range(200):
        rlim2 = 10.0 ** rng.uniform(-12, 4)
        thr, span = seed_threshold(rlim2)
        r2 = rlim2 * (1.0 + rng.uniform(-1e-5, 1e-5, 4000))
        _, hy = rcube_seeded(r2, thr, span)
        acc = hy != 0
        assert not np.any(acc & ~(r2 > rlim2))
        assert np.all(acc[r2 > rlim2 * (1 + 2.0 ** -19)])


def test_inverse_cube_error_budget_over_the_whole_range():
    rng = np.random.default_rng(2)
    thr, span = seed_threshold(0.0)
    r2 = 2.0 ** rng.uniform(-125.9, 127.9, 400000)
    val, hy = rcube_seeded(r2, thr, span, rng=rng)
    assert np.all(hy != 0)
    exact = r2.astype(np.longdouble) ** np.longdouble(-1.5)
    rel = np.abs((val.astype(np.longdouble) - exact) / exact).astype(np.float64)
    assert rel.max() < 1.5e-15          # against the 1e-12 parity bar of the accelerations
    # the variant without the upper range test accepts everything above the threshold
    val2, hy2 = rcube_seeded(r2, thr, span, upper=False, rng=None)
    assert np.all(hy2 != 0)


def test_rejected_pairs_contribute_exactly_zero():
    thr, span = seed_threshold(1.0)        # radius sum 1: everything with r2 <= 1 is refused
    r2 = np.array([0.0, 5e-324, 1e-300, 0.3, 1.0, 2.0 ** 128, 1e300])
    val, hy = rcube_seeded(r2, thr, span)
    assert np.all(hy == 0) and np.all(val == 0.0)
