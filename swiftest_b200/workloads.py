"""Synthetic inputs of the shapes BASELINE.json names (numpy only; no compute-path code lives here).

* `disk(n, seed)`            SyMBA planetesimal disk in the style of examples/Chambers2013/initial_conditions.py
                             (:52-120): two-slope semi-major-axis profile on [0.3, 2] AU (pdf ~ a^2 below 0.7 AU,
                             ~ a^-1/2 above), e ~ Rayleigh(0.01), inc ~ Rayleigh(0.005 rad), uniform angles,
                             equal masses with a fixed total disk mass, rho = 3000 kg/m^3 radii, Hill radii.
* `tp_cloud(n, seed)`        test particles with the element distributions of tests/test_swiftest.py:94-99
                             extended to a in [0.5, 40] AU.
* `fixture(name)`            the reference's ASCII initial-condition fixtures, parsed once by
                             tests/golden/gen_golden.py into tests/golden/*.npz.

Units: AU, year, solar mass (GM_sun = 39.476926408897626, examples/Swifter_Swiftest/108pl_50tp/cb.in).
"""
import os

import numpy as np

GMSUN = 39.476926408897626
MSUN_KG = 1.988409870698051e30       # MU2KG of the 108pl_50tp fixture
AU_M = 149597870700.0                # DU2M
_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def kepler_E(M, e, iters=60):
    """Eccentric anomaly from mean anomaly (Newton iterations, vectorised)."""
    E = np.where(e < 0.8, M, np.pi * np.ones_like(M))
    for _ in range(iters):
        E = E - (E - e * np.sin(E) - M) / (1.0 - e * np.cos(E))
    return E


def el2xv(mu, a, e, inc, capom, omega, capm):
    """Elliptic orbital elements (radians) -> heliocentric Cartesian position and velocity, arrays (n,3)."""
    a, e, inc, capom, omega, capm = map(np.asarray, (a, e, inc, capom, omega, capm))
    E = kepler_E(np.mod(capm, 2 * np.pi), e)
    cE, sE = np.cos(E), np.sin(E)
    b = a * np.sqrt(1.0 - e * e)
    xp, yp = a * (cE - e), b * sE                     # perifocal position
    n = np.sqrt(mu / a ** 3)
    rr = a * (1.0 - e * cE)
    vxp, vyp = -a * a * n * sE / rr, a * b * n * cE / rr
    co, so, cO, sO, ci, si = np.cos(omega), np.sin(omega), np.cos(capom), np.sin(capom), np.cos(inc), np.sin(inc)
    r11, r12 = cO * co - sO * so * ci, -cO * so - sO * co * ci
    r21, r22 = sO * co + cO * so * ci, -sO * so + cO * co * ci
    r31, r32 = so * si, co * si
    x = np.stack([r11 * xp + r12 * yp, r21 * xp + r22 * yp, r31 * xp + r32 * yp], axis=-1)
    v = np.stack([r11 * vxp + r12 * vyp, r21 * vxp + r22 * vyp, r31 * vxp + r32 * vyp], axis=-1)
    return np.ascontiguousarray(x), np.ascontiguousarray(v)


def _two_slope_a(rng, n, a_in=0.3, a_brk=0.7, a_out=2.0):
    """Inverse-CDF sample of pdf ~ a^2 on [a_in,a_brk], ~ a_brk^2.5 * a^-0.5 on [a_brk,a_out] (continuous)."""
    w1 = (a_brk ** 3 - a_in ** 3) / 3.0
    w2 = a_brk ** 2.5 * 2.0 * (np.sqrt(a_out) - np.sqrt(a_brk))
    u = rng.uniform(0.0, w1 + w2, n)
    lo = np.cbrt(3.0 * np.minimum(u, w1) + a_in ** 3)
    hi = (np.maximum(u - w1, 0.0) / (2.0 * a_brk ** 2.5) + np.sqrt(a_brk)) ** 2
    return np.where(u < w1, lo, hi)


def disk(n, seed=3031179, total_mass=7.84e-6, sort_by_mass=True):
    """SyMBA planetesimal disk of n fully interacting bodies.  Returns a dict of numpy arrays:
    rh, vh (n,3); Gmass, radius, rhill, mu (n,); plus dt (6.0875/365.25 y) and nplm = n."""
    rng = np.random.default_rng(seed)
    a = _two_slope_a(rng, n)
    e = rng.rayleigh(0.01, n)
    inc = rng.rayleigh(0.005, n)
    capom, omega, capm = (rng.uniform(0.0, 2 * np.pi, n) for _ in range(3))
    m = (total_mass / n) * (1.0 + 1e4 * rng.uniform(-np.finfo(float).eps, np.finfo(float).eps, n))
    if sort_by_mass:  # the reference keeps massive bodies sorted by mass, descending (swiftest_util.f90:1709)
        order = np.argsort(-m, kind="stable")
        a, e, inc, capom, omega, capm, m = (q[order] for q in (a, e, inc, capom, omega, capm, m))
    Gm = GMSUN * m
    rh, vh = el2xv(GMSUN + Gm, a, e, inc, capom, omega, capm)
    radius = (3.0 * m * MSUN_KG / (4.0 * np.pi * 3000.0)) ** (1.0 / 3.0) / AU_M
    rhill = a * (m / 3.0) ** (1.0 / 3.0)
    return dict(rh=rh, vh=vh, Gmass=Gm, radius=radius, rhill=rhill, mu=GMSUN + Gm, a=a, dt=6.0875 / 365.25, nplm=n,
                n=n)


def tp_cloud(n, seed=123, a_lo=0.5, a_hi=40.0):
    """Test particles: a~U(a_lo,a_hi) AU, e~U(0,0.2), inc~U(0,10 deg), angles uniform."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(a_lo, a_hi, n)
    e = rng.uniform(0.0, 0.2, n)
    inc = np.deg2rad(rng.uniform(0.0, 10.0, n))
    capom, omega, capm = (rng.uniform(0.0, 2 * np.pi, n) for _ in range(3))
    rh, vh = el2xv(GMSUN, a, e, inc, capom, omega, capm)
    return dict(rh=rh, vh=vh, mu=np.full(n, GMSUN), n=n)


def fixture(name):
    """'108pl_50tp' or '8pl_0tp' -> dict with cb_Gmass, pl_{Gmass,rhill,radius,rh,vh}, tp_{rh,vh} and params."""
    z = np.load(os.path.join(_GOLDEN, f"fixture_{name}.npz"))
    return {k: z[k] for k in z.files}


def planets8_year_units():
    """Sun + 8 planets of the 8pl_0tp fixture converted from (AU, day, GM in AU^3/day^2) to (AU, year)."""
    f = fixture("8pl_0tp")
    k = 365.25
    return dict(rh=f["pl_rh"].copy(), vh=f["pl_vh"] * k, Gmass=f["pl_Gmass"] * k * k, radius=f["pl_radius"].copy(),
                rhill=f["pl_rhill"].copy(), cb_Gmass=float(f["cb_Gmass"]) * k * k)
