"""Generates the committed fixtures under tests/golden/ from the reference tree (run in the build container only;
/root/reference does not exist on the GPU box and nothing reads it at test/bench time).

  python tests/golden/gen_golden.py

1. fixture_108pl_50tp.npz, fixture_8pl_0tp.npz
     the reference's ASCII initial conditions examples/Swifter_Swiftest/{108pl_50tp,8pl_0tp}/ parsed into arrays
     (format: swiftest/swiftest_io.f90:3319-3356: "name Gmass rhill" / "radius" / "x y z" / "vx vy vz").
2. drift_kepler_ref.npz
     REFERENCE-GENERATED golden vectors for the Kepler drift: the reference's own Python two-body code
     (swiftest/tool.py el2xv_one + danby, imported from /root/reference with a stub for the uninstalled xarray)
     evaluates the same elliptic orbit at mean anomaly M0 and M0 + n*dt.  The drift oracle must carry the first
     state into the second (tests/test_oracle.py, tolerance 1e-11 relative: the Python solver stops at 1e-14).
3. xv2aeq_ref.npz
     REFERENCE-GENERATED vectors for swiftest_orbel_xv2aeq (the orbital elements inside collision_check_one): a and e
     from the reference's Python xv2el_one on 200 random elliptic relative orbits.
4. el2xv_ref.npz
     REFERENCE-GENERATED states from el2xv_one for the element ranges of the synthetic workloads.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _tokens(path):
    toks = []
    for line in open(path):
        line = line.split("!")[0].strip()
        if line:
            toks.extend(line.split())
    return toks


def parse_pl(path):
    t = _tokens(path)
    n = int(t[0])
    p = 1
    names, Gm, rhill, radius, rh, vh = [], [], [], [], [], []
    for _ in range(n):
        names.append(t[p]); Gm.append(float(t[p + 1])); rhill.append(float(t[p + 2])); p += 3
        radius.append(float(t[p])); p += 1
        rh.append([float(q) for q in t[p:p + 3]]); p += 3
        vh.append([float(q) for q in t[p:p + 3]]); p += 3
    assert p == len(t), (p, len(t))
    return names, np.array(Gm), np.array(rhill), np.array(radius), np.array(rh), np.array(vh)


def parse_tp(path):
    t = _tokens(path)
    n = int(t[0])
    p = 1
    rh, vh = [], []
    for _ in range(n):
        p += 1
        rh.append([float(q) for q in t[p:p + 3]]); p += 3
        vh.append([float(q) for q in t[p:p + 3]]); p += 3
    assert p == len(t)
    return np.array(rh).reshape(n, 3), np.array(vh).reshape(n, 3)


def parse_cb(path):
    t = _tokens(path)
    return float(t[1]), float(t[2])


def write_fixture(name, cb_file, pl_file, tp_file, extra):
    d = os.path.join(REF, "examples", "Swifter_Swiftest", name)
    cbG, cbR = parse_cb(os.path.join(d, cb_file))
    names, Gm, rhill, radius, rh, vh = parse_pl(os.path.join(d, pl_file))
    trh, tvh = parse_tp(os.path.join(d, tp_file))
    np.savez(os.path.join(OUT, f"fixture_{name}.npz"), cb_Gmass=cbG, cb_radius=cbR, pl_Gmass=Gm, pl_rhill=rhill,
             pl_radius=radius, pl_rh=rh, pl_vh=vh, tp_rh=trh, tp_vh=tvh, pl_names=np.array(names), **extra)
    print(name, "npl", len(Gm), "ntp", len(trh))


def load_reference_tool():
    sys.modules.setdefault("xarray", types.ModuleType("xarray"))  # tool.py imports it at module level only
    spec = importlib.util.spec_from_file_location("ref_tool", os.path.join(REF, "swiftest", "tool.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def write_drift_golden():
    tool = load_reference_tool()
    rng = np.random.default_rng(20231002)
    mu = 39.476926408897626
    rows = []
    # (a range, e range, dt): small-dm/small-e (kepmd fast path), larger steps and eccentricities (kepu paths)
    cases = [((0.3, 40.0), (0.0, 0.3), 0.01, 120), ((0.3, 2.0), (0.0, 0.2), 6.0875 / 365.25, 60),
             ((0.2, 5.0), (0.3, 0.95), 0.05, 80), ((0.05, 1.0), (0.0, 0.9), 0.2, 60)]
    for (alo, ahi), (elo, ehi), dt, cnt in cases:
        for _ in range(cnt):
            a = rng.uniform(alo, ahi)
            e = rng.uniform(elo, ehi)
            inc, Om, om, M0 = rng.uniform(0, 60), rng.uniform(0, 360), rng.uniform(0, 360), rng.uniform(0, 360)
            n = np.sqrt(mu / a ** 3)
            M1 = M0 + np.rad2deg(n * dt)
            r0, v0 = tool.el2xv_one(mu, a, e, inc, Om, om, M0)
            r1, v1 = tool.el2xv_one(mu, a, e, inc, Om, om, M1)
            rows.append(np.concatenate([[mu, dt, a, e], r0, v0, r1, v1]))
    rows = np.array(rows)
    np.savez(os.path.join(OUT, "drift_kepler_ref.npz"), mu=rows[:, 0], dt=rows[:, 1], a=rows[:, 2], e=rows[:, 3],
             x0=rows[:, 4:7], v0=rows[:, 7:10], x1=rows[:, 10:13], v1=rows[:, 13:16])
    print("drift golden", rows.shape)


def write_el2xv_golden():
    """REFERENCE-GENERATED states for the synthetic-workload generator: el2xv_one (swiftest/tool.py:221-343) on the
    element ranges of the disk / test-particle workloads (swiftest_b200/workloads.py restates it, vectorised)."""
    tool = load_reference_tool()
    rng = np.random.default_rng(4711)
    mu = 39.476926408897626
    rows = []
    for _ in range(150):
        a, e = rng.uniform(0.3, 40.0), rng.uniform(0.0, 0.3)
        inc, Om, om, M = rng.uniform(0, 30), rng.uniform(0, 360), rng.uniform(0, 360), rng.uniform(0, 360)
        r, v = tool.el2xv_one(mu, a, e, inc, Om, om, M)
        rows.append(np.concatenate([[a, e, inc, Om, om, M], r, v]))
    rows = np.array(rows)
    np.savez(os.path.join(OUT, "el2xv_ref.npz"), mu=mu, elements_deg=rows[:, :6], r=rows[:, 6:9], v=rows[:, 9:12])
    print("el2xv golden", rows.shape)


def write_xv2aeq_golden():
    """REFERENCE-GENERATED vectors for swiftest_orbel_xv2aeq (used by collision_check_one): the reference's Python
    xv2el_one (swiftest/tool.py:377-455) gives a and e of the relative orbit; q = a(1-e)."""
    tool = load_reference_tool()
    rng = np.random.default_rng(7002)
    rows = []
    for _ in range(200):
        mu = 10.0 ** rng.uniform(-8, 2)
        a = 10.0 ** rng.uniform(-3, 2)
        e = rng.uniform(0.0, 0.98)
        inc, Om, om, M = rng.uniform(0, 180), rng.uniform(0, 360), rng.uniform(0, 360), rng.uniform(0, 360)
        r, v = tool.el2xv_one(mu, a, e, inc, Om, om, M)
        el = tool.xv2el_one(mu, np.asarray(r), np.asarray(v))
        rows.append(np.concatenate([[mu], r, v, [float(el[0]), float(el[1])]]))
    rows = np.array(rows)
    np.savez(os.path.join(OUT, "xv2aeq_ref.npz"), mu=rows[:, 0], r=rows[:, 1:4], v=rows[:, 4:7], a=rows[:, 7], e=rows[:, 8])
    print("xv2aeq golden", rows.shape)


if __name__ == "__main__":
    write_fixture("108pl_50tp", "cb.in", "pl.swiftest.in", "tp.swiftest.in",
                  dict(dt=0.005, GMTINY=2.1554293571575797e-06))
    write_fixture("8pl_0tp", "cb.swiftest.in", "pl.swiftest.in", "tp.swiftest.in", dict(dt=1.0))
    write_drift_golden()
    write_xv2aeq_golden()
    write_el2xv_golden()
