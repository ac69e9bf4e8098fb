"""The compiled C++ host mirror (swiftest_b200/host/swiftest_host.hpp: swiftest_pl/tp, symba_pl type-bound procedures and
the generic swiftest_kick_getacch_int_all overloads) driven through one SyMBA-style step and compared with the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from swiftest_b200 import workloads as W

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "swiftest_b200", "lib", "host_selftest")


def _run(tmp_path, pl, tp, dt, flat, gmtiny, cbG):
    npl, ntp = len(pl["Gmass"]), len(tp["rh"])
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<4i3d", npl, ntp, int(flat), int(gmtiny is not None), dt, gmtiny or -1.0, cbG))
        for a in (pl["rh"], pl["vb"], pl["Gmass"], pl["radius"], pl["rhill"], tp["rh"], tp["vb"]):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    subprocess.check_call([EXE, fin, fout])
    raw = open(fout, "rb").read()
    counts = np.frombuffer(raw, np.int64, 4)
    off = 32
    out = {"counts": counts}
    for key, dt_, n in (("p1", np.int32, counts[0]), ("p2", np.int32, counts[0]), ("t1", np.int32, counts[1]),
                        ("t2", np.int32, counts[1]), ("pl_ah", np.float64, 3 * npl), ("tp_ah", np.float64, 3 * ntp),
                        ("pl_rh", np.float64, 3 * npl), ("pl_vb", np.float64, 3 * npl)):
        nb = int(n) * np.dtype(dt_).itemsize
        out[key] = np.frombuffer(raw[off:off + nb], dt_)
        off += nb
    assert off == len(raw)
    return out


@pytest.mark.parametrize("flat", [False, True])
@pytest.mark.parametrize("gmtiny", [None, "split"])
def test_symba_style_step_through_cpp_host_mirror(tmp_path, oracle, flat, gmtiny):
    assert os.path.exists(EXE), "build with __graft_entry__.build()"
    n, ntp = 2500, 4000
    d = W.disk(n, seed=31)
    t = W.tp_cloud(ntp, seed=32, a_lo=0.3, a_hi=2.0)
    rhill = d["rhill"] * 3.0  # enough encounters to exercise the subtract
    gm = float(np.sort(d["Gmass"])[::-1][799]) if gmtiny else None  # 800 fully interacting bodies
    pl = dict(rh=d["rh"], vb=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=rhill)
    out = _run(tmp_path, pl, dict(rh=t["rh"], vb=t["vh"]), d["dt"], flat, gm, W.GMSUN)
    nplm = int((d["Gmass"] >= gm).sum()) if gmtiny else n
    assert out["counts"][2] == nplm
    renc = oracle.set_renc(rhill, 0)
    # pl-pl encounters: plpl or the merged plplm list
    if nplm == n:
        r1, r2, _ = oracle.encounter_plpl(d["rh"], d["vh"], renc, d["dt"])
    else:
        r1, r2, _ = oracle.encounter_plplm(d["rh"][:nplm], d["vh"][:nplm], d["rh"][nplm:], d["vh"][nplm:], renc[:nplm],
                                           renc[nplm:], d["dt"], merged=True)
    assert len(r1) > 5 and np.array_equal(out["p1"], r1) and np.array_equal(out["p2"], r2)
    q1, q2, _ = oracle.encounter_pltp(d["rh"], d["vh"], t["rh"], t["vh"], renc, d["dt"])
    assert np.array_equal(out["t1"], q1) and np.array_equal(out["t2"], q2)
    # accelerations: all interactions minus the encounter pairs (F1)
    ah = oracle.kick_tri_pl(d["rh"], d["Gmass"], d["radius"], np.zeros((n, 3)), nplm=nplm)
    ah = oracle.symba_kick_subtract_enc(r1, r2, d["rh"], d["Gmass"], d["radius"], ah)
    scale = oracle.kick_tri_abs_scale(d["rh"], d["Gmass"], d["radius"], nplm=nplm)
    assert np.max(np.abs(out["pl_ah"].reshape(n, 3) - ah) / scale) < 1e-12
    at = oracle.kick_all_tp(t["rh"], d["rh"], d["Gmass"], np.ones(ntp, np.int32), np.zeros((ntp, 3)))
    assert np.max(np.abs(out["tp_ah"].reshape(ntp, 3) - at)) <= 1e-12 * np.abs(at).max()
    xr, vr, fl = oracle.drift_all(W.GMSUN, d["rh"], d["vh"], d["dt"])
    assert not fl.any() and out["counts"][3] % 10 == 0
    assert np.array_equal(out["pl_rh"].reshape(n, 3), xr) and np.array_equal(out["pl_vb"].reshape(n, 3), vr)
