"""`Context`: numpy-facing mirror of the reference's array-level interfaces over the C ABI.

Method names follow the reference procedures they stand for (swiftest/swiftest_kick.f90,
swiftest/swiftest_drift.f90, encounter/encounter_check.f90, symba/symba_kick.f90); argument meaning and the
early-return / error behaviour are the reference's.  Every method runs CUDA kernels through
libswiftest_cuda.so -- there is no CPU path here.
"""
import ctypes as C

import numpy as np

from ._lib import SwcuError, load

PL, TP = 0, 1
LOOP_TRIANGULAR, LOOP_FLAT, LOOP_AUTO = 0, 1, 2
FAM_PLPL, FAM_PLTP, FAM_DRIFT, FAM_SWEEP, FAM_ALLGATHER = range(5)

_f64, _i32 = np.float64, np.int32


def _vec3(a, n=None, name="array"):
    a = np.ascontiguousarray(a, dtype=_f64)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError(f"{name} must have shape (n,3) (Fortran r(3,n))")
    if n is not None and a.shape[0] != n:
        raise ValueError(f"{name} has {a.shape[0]} bodies, expected {n}")
    return a


def _vec(a, n=None, dt=_f64, name="array"):
    a = np.ascontiguousarray(a, dtype=dt)
    if a.ndim != 1 or (n is not None and a.shape[0] != n):
        raise ValueError(f"{name} must be a vector of length {n}")
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data


class Context:
    """One GPU context (swcu_create).  Use as a context manager or call close()."""

    def __init__(self, device=0):
        self._L = load()
        h = C.c_void_p()
        rc = self._L.swcu_create(int(device), C.byref(h))
        if rc != 0:
            raise SwcuError({5: "no sm_100 (B200) GPU visible: swiftest_b200 has no CPU fallback",
                             2: f"bad device index {device}"}.get(rc, f"swcu_create failed with status {rc}"))
        self._h = h
        self.device = device

    # ---- plumbing ----
    def _ck(self, rc):
        if rc != 0:
            msg = self._L.swcu_last_error(self._h)
            raise SwcuError(f"status {rc}: {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "_h", None):
            self._L.swcu_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        self._ck(self._L.swcu_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        self._ck(self._L.swcu_synchronize(self._h))

    def device_info(self):
        sm, cc, mem = C.c_int32(), C.c_int32(), C.c_int64()
        self._ck(self._L.swcu_device_info(self._h, C.byref(sm), C.byref(cc), C.byref(mem)))
        return {"sm_count": sm.value, "cc": cc.value, "mem_bytes": mem.value}

    def launch_count(self):
        return int(self._L.swcu_launch_count(self._h))

    # ---- tier 1: array-level, host arrays (numpy) ----
    def kick_getacch_int_all_tri_pl(self, npl, nplm, r, Gmass, radius, acc):
        """swiftest_kick_getacch_int_all_tri_{rad,norad}_pl (kick.f90:165-371); radius=None -> norad.
        acc (npl,3) float64 C-contiguous is updated in place."""
        r, Gmass = _vec3(r, npl, "r"), _vec(Gmass, npl, name="Gmass")
        radius = None if radius is None else _vec(radius, npl, name="radius")
        self._inplace3(acc, npl)
        self._ck(self._L.swcu_kick_getacch_int_all_tri_pl(self._h, npl, nplm, _ptr(r), _ptr(Gmass), _ptr(radius),
                                                          _ptr(acc)))
        return acc

    def kick_getacch_int_all_flat_pl(self, npl, nplpl, k_plpl, r, Gmass, radius, acc):
        """swiftest_kick_getacch_int_all_flat_{rad,norad}_pl (kick.f90:69-162).  k_plpl=None: canonical flattened
        pairs 1..nplpl; else an (nplpl,2) int32 array of 1-based pairs (Fortran k_plpl(2,nplpl))."""
        r, Gmass = _vec3(r, npl, "r"), _vec(Gmass, npl, name="Gmass")
        radius = None if radius is None else _vec(radius, npl, name="radius")
        if k_plpl is not None:
            k_plpl = np.ascontiguousarray(k_plpl, dtype=_i32)
            if k_plpl.ndim != 2 or k_plpl.shape != (nplpl, 2):
                raise ValueError("k_plpl must have shape (nplpl,2)")
        self._inplace3(acc, npl)
        self._ck(self._L.swcu_kick_getacch_int_all_flat_pl(self._h, npl, int(nplpl), _ptr(k_plpl), _ptr(r), _ptr(Gmass),
                                                           _ptr(radius), _ptr(acc)))
        return acc

    def kick_getacch_int_all_tp(self, ntp, npl, rtp, rpl, GMpl, lmask, acc):
        """swiftest_kick_getacch_int_all_tp (kick.f90:374-415)."""
        rtp, rpl, GMpl = _vec3(rtp, ntp, "rtp"), _vec3(rpl, npl, "rpl"), _vec(GMpl, npl, name="GMpl")
        lmask = _vec(lmask, ntp, _i32, "lmask")
        self._inplace3(acc, ntp)
        self._ck(self._L.swcu_kick_getacch_int_all_tp(self._h, ntp, npl, _ptr(rtp), _ptr(rpl), _ptr(GMpl), _ptr(lmask),
                                                      _ptr(acc)))
        return acc

    def symba_kick_subtract_encounters(self, npl, index1, index2, rh, Gmass, radius, ah):
        """The encounter-pair removal of symba_kick_getacch_pl (symba_kick.f90:59-70): ah -= flat_rad(list)."""
        i1, i2 = _vec(index1, dt=_i32), _vec(index2, dt=_i32)
        rh, Gmass, radius = _vec3(rh, npl), _vec(Gmass, npl), _vec(radius, npl)
        self._inplace3(ah, npl)
        self._ck(self._L.swcu_symba_kick_subtract_encounters(self._h, npl, len(i1), _ptr(i1), _ptr(i2), _ptr(rh),
                                                             _ptr(Gmass), _ptr(radius), _ptr(ah)))
        return ah

    def drift_all(self, mu, x, v, n, dt, lmask, iflag, lgr=False, inv_c2=0.0):
        """swiftest_drift_all (drift.f90:60-108): x, v (n,3) and iflag (n,) int32 are updated in place."""
        mu = _vec(mu, n, name="mu")
        lmask = _vec(lmask, n, _i32, "lmask")
        self._inplace3(x, n)
        self._inplace3(v, n)
        if not (isinstance(iflag, np.ndarray) and iflag.dtype == _i32 and iflag.flags.c_contiguous and iflag.shape == (n,)):
            raise ValueError("iflag must be a C-contiguous int32 vector of length n")
        self._ck(self._L.swcu_drift_all(self._h, n, _ptr(mu), _ptr(x), _ptr(v), float(dt), int(bool(lgr)), float(inv_c2),
                                        _ptr(lmask), _ptr(iflag)))
        return x, v, iflag

    def _fetch(self, nenc):
        i1, i2, lv = np.empty(nenc, _i32), np.empty(nenc, _i32), np.empty(nenc, _i32)
        self._ck(self._L.swcu_encounter_fetch(self._h, nenc, _ptr(i1), _ptr(i2), _ptr(lv)))
        return nenc, i1, i2, lv.astype(bool)

    def encounter_check_all_sort_and_sweep_plpl(self, npl, r, v, renc, dt):
        """encounter_check.f90:143-192 -> (nenc, index1, index2, lvdotr), 1-based, lexicographic order."""
        r, v, renc = _vec3(r, npl), _vec3(v, npl), _vec(renc, npl)
        n = C.c_int64()
        self._ck(self._L.swcu_encounter_check_all_sort_and_sweep_plpl(self._h, npl, _ptr(r), _ptr(v), _ptr(renc),
                                                                      float(dt), C.byref(n)))
        return self._fetch(n.value)

    def encounter_check_all_sort_and_sweep_pltp(self, npl, ntp, rpl, vpl, rtp, vtp, rencpl, dt):
        """encounter_check.f90:261-326."""
        rpl, vpl, rtp, vtp = _vec3(rpl, npl), _vec3(vpl, npl), _vec3(rtp, ntp), _vec3(vtp, ntp)
        rencpl = _vec(rencpl, npl)
        n = C.c_int64()
        self._ck(self._L.swcu_encounter_check_all_sort_and_sweep_pltp(self._h, npl, ntp, _ptr(rpl), _ptr(vpl), _ptr(rtp),
                                                                      _ptr(vtp), _ptr(rencpl), float(dt), C.byref(n)))
        return self._fetch(n.value)

    def encounter_check_all_sort_and_sweep_plplm(self, nplm, nplt, rplm, vplm, rplt, vplt, rencm, renct, dt):
        """encounter_check.f90:195-258 (index2 counts within the plt block)."""
        a = [_vec3(rplm, nplm), _vec3(vplm, nplm), _vec3(rplt, nplt), _vec3(vplt, nplt), _vec(rencm, nplm),
             _vec(renct, nplt)]
        n = C.c_int64()
        self._ck(self._L.swcu_encounter_check_all_sort_and_sweep_plplm(self._h, nplm, nplt, *[_ptr(q) for q in a],
                                                                       float(dt), C.byref(n)))
        return self._fetch(n.value)

    def encounter_check_all_plplm(self, nplm, nplt, rplm, vplm, rplt, vplt, rencm, renct, dt):
        """encounter_check.f90:42-109 with SORTSWEEP: merged plpl + plm-plt list, index2 shifted by nplm."""
        a = [_vec3(rplm, nplm), _vec3(vplm, nplm), _vec3(rplt, nplt), _vec3(vplt, nplt), _vec(rencm, nplm),
             _vec(renct, nplt)]
        n = C.c_int64()
        self._ck(self._L.swcu_encounter_check_all_plplm(self._h, nplm, nplt, *[_ptr(q) for q in a], float(dt),
                                                        C.byref(n)))
        return self._fetch(n.value)

    def encounter_stats(self):
        a, b = C.c_int64(), C.c_int64()
        self._ck(self._L.swcu_encounter_stats(self._h, C.byref(a), C.byref(b)))
        return {"nbox_total": a.value, "emitted": b.value}

    # ---- tier 2: device-resident populations ----
    def body_sync(self, kind, n, nplm=0, r=None, v=None, Gmass=None, radius=None, rhill=None, mu=None, lmask=None,
                  generation=0):
        r = None if r is None else _vec3(r, n)
        v = None if v is None else _vec3(v, n)
        arrs = [None if q is None else _vec(q, n) for q in (Gmass, radius, rhill, mu)]
        lmask = None if lmask is None else _vec(lmask, n, _i32)
        self._ck(self._L.swcu_body_sync(self._h, kind, n, nplm, _ptr(r), _ptr(v), *[_ptr(q) for q in arrs], _ptr(lmask),
                                        int(generation)))

    def body_put(self, kind, r=None, v=None, a=None, lmask=None):
        r, v, a = [None if q is None else _vec3(q) for q in (r, v, a)]
        lmask = None if lmask is None else _vec(lmask, dt=_i32)
        self._ck(self._L.swcu_body_put(self._h, kind, _ptr(r), _ptr(v), _ptr(a), _ptr(lmask)))

    def body_put_range(self, kind, i0, i1, r=None, v=None):
        """Refresh bodies [i0, i1) from host arrays of that length (a rank's slice)."""
        r, v = [None if q is None else _vec3(q, i1 - i0) for q in (r, v)]
        self._ck(self._L.swcu_body_put_range(self._h, kind, int(i0), int(i1), _ptr(r), _ptr(v)))

    def body_get_range(self, kind, i0, i1, r=True, v=True, a=True, out=None):
        """Read bodies [i0, i1) back; `out` may supply preallocated (pinned) arrays keyed 'r','v','a'."""
        out, res = dict(out or {}), {}
        for key, want in (("r", r), ("v", v), ("a", a)):
            if want:
                res[key] = out.get(key) if out.get(key) is not None else np.empty((i1 - i0, 3), _f64)
        self._ck(self._L.swcu_body_get_range(self._h, kind, int(i0), int(i1), _ptr(res.get("r")), _ptr(res.get("v")),
                                             _ptr(res.get("a"))))
        return res

    def body_put_range_async(self, kind, i0, i1, r=None, v=None):
        """Enqueue the refresh of bodies [i0, i1) from PINNED host arrays (kept alive by the caller until io_wait)."""
        for q in (r, v):
            if q is not None and not (q.dtype == _f64 and q.flags.c_contiguous and q.shape == (i1 - i0, 3)):
                raise ValueError("async arrays must be C-contiguous float64 of shape (i1-i0, 3)")
        self._ck(self._L.swcu_body_put_range_async(self._h, kind, int(i0), int(i1), _ptr(r), _ptr(v)))

    def body_get_range_async(self, kind, i0, i1, out):
        """Enqueue the read-back of bodies [i0, i1) into the PINNED arrays out['r'|'v'|'a'] (defined after io_wait)."""
        for q in out.values():
            if not (q.dtype == _f64 and q.flags.c_contiguous and q.shape == (i1 - i0, 3)):
                raise ValueError("async arrays must be C-contiguous float64 of shape (i1-i0, 3)")
        self._ck(self._L.swcu_body_get_range_async(self._h, kind, int(i0), int(i1), _ptr(out.get("r")), _ptr(out.get("v")),
                                                   _ptr(out.get("a"))))

    def io_wait(self):
        self._ck(self._L.swcu_io_wait(self._h))

    def body_count(self, kind):
        n, nplm, gen = C.c_int32(), C.c_int32(), C.c_uint64()
        self._ck(self._L.swcu_body_count(self._h, kind, C.byref(n), C.byref(nplm), C.byref(gen)))
        return n.value, nplm.value, gen.value

    def body_get(self, kind, r=True, v=True, a=True, iflag=False, out=None):
        """Read resident arrays back.  `out` may supply preallocated (pinned) arrays keyed 'r','v','a','iflag'."""
        n = self.body_count(kind)[0]
        out = dict(out or {})
        res = {}
        for key, want in (("r", r), ("v", v), ("a", a)):
            if want:
                res[key] = out.get(key) if out.get(key) is not None else np.empty((n, 3), _f64)
        if iflag:
            res["iflag"] = out.get("iflag") if out.get("iflag") is not None else np.empty(n, _i32)
        self._ck(self._L.swcu_body_get(self._h, kind, _ptr(res.get("r")), _ptr(res.get("v")), _ptr(res.get("a")),
                                       _ptr(res.get("iflag"))))
        return res

    def body_zero_accel(self, kind):
        self._ck(self._L.swcu_body_zero_accel(self._h, kind))

    def pl_accel_int(self, loop_variant=LOOP_TRIANGULAR, lclose=True):
        self._ck(self._L.swcu_pl_accel_int(self._h, loop_variant, int(bool(lclose))))

    def tp_accel_int(self):
        self._ck(self._L.swcu_tp_accel_int(self._h))

    def pl_set_renc(self, irec):
        self._ck(self._L.swcu_pl_set_renc(self._h, irec))

    def body_drift(self, kind, dt, lgr=False, inv_c2=0.0, want_nfail=True):
        nf = C.c_int32()
        self._ck(self._L.swcu_body_drift(self._h, kind, float(dt), int(bool(lgr)), float(inv_c2),
                                         C.byref(nf) if want_nfail else None))
        return nf.value

    def whm_tp_step(self, dt, ah0, want_nfail=True):
        """Fused whm_step_tp on the resident test particles (kick dt/2 with the kept ah, drift dt, new ah at the resident
        planets' positions + ah0, kick dt/2).  Returns the number of particles whose drift failed."""
        ah0 = None if ah0 is None else _vec(ah0, 3, name="ah0")   # None: the value whm_step_pl left on the device
        nf = C.c_int32()
        self._ck(self._L.swcu_whm_tp_step(self._h, float(dt), _ptr(ah0), C.byref(nf) if want_nfail else None))
        return nf.value

    def whm_step_pl(self, GMcb, dt, loop_variant=LOOP_AUTO, lclose=True, lfirst=False, want_nfail=True):
        """whm_step_pl on the resident planets (Jacobi chains, ah0+ah1+ah2+accel_int, kick, drift, kick)."""
        nf = C.c_int32()
        self._ck(self._L.swcu_whm_step_pl(self._h, float(GMcb), float(dt), loop_variant, int(bool(lclose)), int(bool(lfirst)),
                                          C.byref(nf) if want_nfail else None))
        return nf.value

    def whm_tp_first_accel(self):
        """ah of the resident test particles for their first WHM step (planets at their current = begin positions)."""
        self._ck(self._L.swcu_whm_tp_first_accel(self._h))

    def whm_get_jacobi(self):
        n = self.body_count(PL)[0]
        xj, vj = np.empty((n, 3), _f64), np.empty((n, 3), _f64)
        self._ck(self._L.swcu_whm_get_jacobi(self._h, _ptr(xj), _ptr(vj)))
        return xj, vj

    def body_kick_velocity(self, kind, dt):
        self._ck(self._L.swcu_body_kick_velocity(self._h, kind, float(dt)))

    def pl_encounter_check(self, dt, fetch=True):
        n = C.c_int64()
        self._ck(self._L.swcu_pl_encounter_check(self._h, float(dt), C.byref(n)))
        return self._fetch(n.value) if fetch else n.value

    def pl_encounter_check_triangular(self, dt, fetch=True):
        n = C.c_int64()
        self._ck(self._L.swcu_pl_encounter_check_triangular(self._h, float(dt), C.byref(n)))
        return self._fetch(n.value) if fetch else n.value

    def tp_encounter_check_triangular(self, dt, fetch=True):
        n = C.c_int64()
        self._ck(self._L.swcu_tp_encounter_check_triangular(self._h, float(dt), C.byref(n)))
        return self._fetch(n.value) if fetch else n.value

    def tp_encounter_check(self, dt, fetch=True):
        n = C.c_int64()
        self._ck(self._L.swcu_tp_encounter_check(self._h, float(dt), C.byref(n)))
        return self._fetch(n.value) if fetch else n.value

    # ---- the O(N) glue of the democratic-heliocentric step (helio/helio_step.f90, swiftest_util.f90:363-485) ----
    def _vec3_out(self, fn, *args, want=True):
        out = np.zeros(3, _f64)
        self._ck(fn(self._h, *args, _ptr(out) if want else None))
        return out if want else None

    def pl_vh2vb(self, GMcb, want=True):
        """swiftest_util_coord_vh2vb_pl: vb = vh + vbcb on the resident planets; returns vbcb."""
        return self._vec3_out(self._L.swcu_pl_vh2vb, float(GMcb), want=want)

    def pl_vb2vh(self, GMcb, want=True):
        """swiftest_util_coord_vb2vh_pl: vh = vb - vbcb; returns vbcb."""
        return self._vec3_out(self._L.swcu_pl_vb2vh, float(GMcb), want=want)

    def pl_lindrift(self, GMcb, dt, lbeg, want=True):
        """helio_drift_linear_pl: rh += pt*dt with pt = sum(Gm*vb, lmask)/GMcb; returns pt (kept as ptbeg/ptend)."""
        return self._vec3_out(self._L.swcu_pl_lindrift, float(GMcb), float(dt), int(bool(lbeg)), want=want)

    def tp_lindrift(self, dt, lbeg):
        self._ck(self._L.swcu_tp_lindrift(self._h, float(dt), int(bool(lbeg))))

    def cb_set_pt(self, ptbeg=None, ptend=None):
        ptbeg = None if ptbeg is None else _vec(ptbeg, 3)
        ptend = None if ptend is None else _vec(ptend, 3)
        self._ck(self._L.swcu_cb_set_pt(self._h, _ptr(ptbeg), _ptr(ptend)))

    def cb_get_pt(self):
        b, e = np.zeros(3, _f64), np.zeros(3, _f64)
        self._ck(self._L.swcu_cb_get_pt(self._h, _ptr(b), _ptr(e)))
        return b, e

    def tp_vh2vb(self, lbeg=True):
        self._ck(self._L.swcu_tp_vh2vb(self._h, int(bool(lbeg))))

    def tp_vb2vh(self, lbeg=False):
        self._ck(self._L.swcu_tp_vb2vh(self._h, int(bool(lbeg))))

    def body_kick_vb(self, kind, dt, lbeg):
        self._ck(self._L.swcu_body_kick_vb(self._h, kind, float(dt), int(bool(lbeg))))

    def body_drift_vb(self, kind, GMcb, dt, want_nfail=True):
        nf = C.c_int32()
        self._ck(self._L.swcu_body_drift_vb(self._h, kind, float(GMcb), float(dt), C.byref(nf) if want_nfail else None))
        return nf.value

    def body_put_vb(self, kind, vb):
        vb = _vec3(vb, self.body_count(kind)[0])
        self._ck(self._L.swcu_body_put_vb(self._h, kind, _ptr(vb)))

    def body_set_active(self, kind, lactive):
        """status /= INACTIVE per body (None: all active); only swiftest_util_coord_vb2vh_pl filters on it."""
        if lactive is None:
            self._ck(self._L.swcu_body_set_active(self._h, kind, None))
        else:
            la = np.ascontiguousarray(lactive, dtype=np.int32)
            self._ck(self._L.swcu_body_set_active(self._h, kind, _ptr(la)))

    def body_get_vb(self, kind, vb=True, rbeg=False, rend=False):
        n = self.body_count(kind)[0]
        res = {k: np.empty((n, 3), _f64) for k, w in (("vb", vb), ("rbeg", rbeg), ("rend", rend)) if w}
        self._ck(self._L.swcu_body_get_vb(self._h, kind, _ptr(res.get("vb")), _ptr(res.get("rbeg")), _ptr(res.get("rend"))))
        return res

    def helio_step_pl(self, GMcb, dt, loop_variant=LOOP_AUTO, lclose=True, lfirst=False, want_nfail=True):
        """helio_step_pl on the resident planets; nothing but the drift-failure count crosses PCIe."""
        nf = C.c_int32()
        self._ck(self._L.swcu_helio_step_pl(self._h, float(GMcb), float(dt), loop_variant, int(bool(lclose)),
                                            int(bool(lfirst)), C.byref(nf) if want_nfail else None))
        return nf.value

    def helio_step_tp(self, GMcb, dt, lfirst=False, want_nfail=True):
        """helio_step_tp as one kernel; call after helio_step_pl of the same step (uses its rbeg/rend/ptbeg/ptend)."""
        nf = C.c_int32()
        self._ck(self._L.swcu_helio_step_tp(self._h, float(GMcb), float(dt), int(bool(lfirst)),
                                            C.byref(nf) if want_nfail else None))
        return nf.value

    # ---- triangular encounter checks, discard, SyMBA list check ----
    def encounter_check_all_triangular_plpl(self, npl, r, v, renc, dt):
        r, v, renc = _vec3(r, npl), _vec3(v, npl), _vec(renc, npl)
        n = C.c_int64()
        self._ck(self._L.swcu_encounter_check_all_triangular_plpl(self._h, npl, _ptr(r), _ptr(v), _ptr(renc), float(dt),
                                                                  C.byref(n)))
        return self._fetch(n.value)

    def encounter_check_all_triangular_pltp(self, npl, ntp, rpl, vpl, rtp, vtp, rencpl, dt):
        rpl, vpl, rtp, vtp = _vec3(rpl, npl), _vec3(vpl, npl), _vec3(rtp, ntp), _vec3(vtp, ntp)
        rencpl = _vec(rencpl, npl)
        n = C.c_int64()
        self._ck(self._L.swcu_encounter_check_all_triangular_pltp(self._h, npl, ntp, _ptr(rpl), _ptr(vpl), _ptr(rtp),
                                                                  _ptr(vtp), _ptr(rencpl), float(dt), C.byref(n)))
        return self._fetch(n.value)

    def encounter_check_all_triangular_plplm(self, nplm, nplt, rplm, vplm, rplt, vplt, rencm, renct, dt):
        rplm, vplm, rplt, vplt = _vec3(rplm, nplm), _vec3(vplm, nplm), _vec3(rplt, nplt), _vec3(vplt, nplt)
        rencm, renct = _vec(rencm, nplm), _vec(renct, nplt)
        n = C.c_int64()
        self._ck(self._L.swcu_encounter_check_all_triangular_plplm(self._h, nplm, nplt, _ptr(rplm), _ptr(vplm), _ptr(rplt),
                                                                   _ptr(vplt), _ptr(rencm), _ptr(renct), float(dt),
                                                                   C.byref(n)))
        return self._fetch(n.value)

    def discard_pl_tp(self, rtp, vtp, lactive, rpl, vpl, radius, dt):
        """swiftest_discard_pl_tp: returns (iplanet[ntp] with the 1-based discarding planet or 0, number discarded)."""
        rtp, vtp = _vec3(rtp), _vec3(vtp)
        ntp = rtp.shape[0]
        rpl, vpl = _vec3(rpl), _vec3(vpl)
        npl = rpl.shape[0]
        radius = _vec(radius, npl)
        lactive = None if lactive is None else _vec(lactive, ntp, _i32)
        ipl = np.zeros(ntp, _i32)
        nd = C.c_int32()
        self._ck(self._L.swcu_discard_pl_tp(self._h, ntp, npl, _ptr(rtp), _ptr(vtp), _ptr(lactive), _ptr(rpl), _ptr(vpl),
                                            _ptr(radius), float(dt), _ptr(ipl), C.byref(nd)))
        return ipl, nd.value

    def tp_discard_pl(self, dt, want_iplanet=True):
        """swiftest_discard_pl_tp on the resident populations: (iplanet or None, number discarded)."""
        ntp = self.body_count(TP)[0]
        ipl = np.zeros(ntp, _i32) if want_iplanet else None
        nd = C.c_int32()
        self._ck(self._L.swcu_tp_discard_pl(self._h, float(dt), _ptr(ipl), C.byref(nd)))
        return ipl, nd.value

    def symba_encounter_check_list(self, index1, index2, lencmask, r1, v1, renc1, radius1, dt, r2=None, v2=None,
                                   renc2=None, radius2=None, lvdotr=None):
        """Pair loop of symba_encounter_check_list_plpl (r2 None) / _pltp.  Returns (lencounter, lvdotr, nfound)."""
        index1, index2 = _vec(index1, dt=_i32), _vec(index2, dt=_i32)
        nenc = len(index1)
        lencmask = None if lencmask is None else _vec(lencmask, nenc, _i32)
        r1, v1 = _vec3(r1), _vec3(v1)
        n1 = r1.shape[0]
        renc1, radius1 = _vec(renc1, n1), _vec(radius1, n1)
        n2 = 0
        if r2 is not None:
            r2, v2 = _vec3(r2), _vec3(v2)
            n2 = r2.shape[0]
            renc2 = None if renc2 is None else _vec(renc2, n2)
            radius2 = None if radius2 is None else _vec(radius2, n2)
        lenc = np.zeros(nenc, _i32)
        lvd = np.zeros(nenc, _i32) if lvdotr is None else _vec(lvdotr, nenc, _i32).copy()
        nf = C.c_int64()
        self._ck(self._L.swcu_symba_encounter_check_list(
            self._h, nenc, _ptr(index1), _ptr(index2), _ptr(lencmask), n1, _ptr(r1), _ptr(v1), _ptr(renc1), _ptr(radius1),
            n2, _ptr(r2), _ptr(v2), _ptr(renc2), _ptr(radius2), float(dt), _ptr(lenc), _ptr(lvd), C.byref(nf)))
        return lenc, lvd, nf.value

    def symba_kick_list_plpl(self, index1, index2, lactive, levelg, rh, rhill, Gmass, dt, irec, sgn, vb):
        """symba_kick_list_plpl: returns (vb after the kick, final lgoodlevel mask)."""
        index1, index2 = _vec(index1, dt=_i32), _vec(index2, dt=_i32)
        nenc = len(index1)
        rh = _vec3(rh)
        npl = rh.shape[0]
        lactive = None if lactive is None else _vec(lactive, nenc, _i32)
        levelg, rhill, Gmass = _vec(levelg, npl, _i32), _vec(rhill, npl), _vec(Gmass, npl)
        vb = _vec3(vb, npl).copy()
        lgood = np.zeros(nenc, _i32)
        self._ck(self._L.swcu_symba_kick_list_plpl(self._h, nenc, _ptr(index1), _ptr(index2), _ptr(lactive), npl,
                                                   _ptr(levelg), _ptr(rh), _ptr(rhill), _ptr(Gmass), float(dt), int(irec),
                                                   int(sgn), _ptr(vb), _ptr(lgood)))
        return vb, lgood

    def symba_kick_list_pltp(self, index1, index2, lactive, levelg_pl, levelg_tp, rh_pl, rhill, Gmass, rh_tp, dt, irec,
                             sgn, vb_tp):
        index1, index2 = _vec(index1, dt=_i32), _vec(index2, dt=_i32)
        nenc = len(index1)
        rh_pl, rh_tp = _vec3(rh_pl), _vec3(rh_tp)
        npl, ntp = rh_pl.shape[0], rh_tp.shape[0]
        lactive = None if lactive is None else _vec(lactive, nenc, _i32)
        levelg_pl, levelg_tp = _vec(levelg_pl, npl, _i32), _vec(levelg_tp, ntp, _i32)
        rhill, Gmass = _vec(rhill, npl), _vec(Gmass, npl)
        vb = _vec3(vb_tp, ntp).copy()
        lgood = np.zeros(nenc, _i32)
        self._ck(self._L.swcu_symba_kick_list_pltp(self._h, nenc, _ptr(index1), _ptr(index2), _ptr(lactive), npl, ntp,
                                                   _ptr(levelg_pl), _ptr(levelg_tp), _ptr(rh_pl), _ptr(rhill), _ptr(Gmass),
                                                   _ptr(rh_tp), float(dt), int(irec), int(sgn), _ptr(vb), _ptr(lgood)))
        return vb, lgood

    def collision_check_list(self, index1, index2, lmask, lvdotr, r1, v1, Gmass1, radius1, dt, r2=None, v2=None):
        """Pair loop of collision_check_plpl (r2 None) / _pltp.  Returns (lcollision, lclosest, ncollision)."""
        index1, index2 = _vec(index1, dt=_i32), _vec(index2, dt=_i32)
        nenc = len(index1)
        lmask = None if lmask is None else _vec(lmask, nenc, _i32)
        lvdotr = _vec(lvdotr, nenc, _i32)
        r1, v1 = _vec3(r1), _vec3(v1)
        n1 = r1.shape[0]
        Gmass1, radius1 = _vec(Gmass1, n1), _vec(radius1, n1)
        n2 = 0
        if r2 is not None:
            r2, v2 = _vec3(r2), _vec3(v2)
            n2 = r2.shape[0]
        lcol, lclo = np.zeros(nenc, _i32), np.zeros(nenc, _i32)
        nc = C.c_int64()
        self._ck(self._L.swcu_collision_check_list(self._h, nenc, _ptr(index1), _ptr(index2), _ptr(lmask), _ptr(lvdotr), n1,
                                                   _ptr(r1), _ptr(v1), _ptr(Gmass1), _ptr(radius1), n2, _ptr(r2), _ptr(v2),
                                                   float(dt), _ptr(lcol), _ptr(lclo), C.byref(nc)))
        return lcol, lclo, nc.value

    # ---- the same list loops on the resident populations (tier 2) ----
    def pl_symba_kick_list(self, index1, index2, lactive, levelg, dt, irec, sgn, want_lgood=True):
        """symba_kick_list_plpl on the resident pl: vb is kicked on the device.  Returns lgood (or None)."""
        index1, index2 = _vec(index1, dt=_i32), _vec(index2, dt=_i32)
        nenc = len(index1)
        lactive = None if lactive is None else _vec(lactive, nenc, _i32)
        levelg = _vec(levelg, dt=_i32)
        lgood = np.zeros(nenc, _i32) if want_lgood else None
        self._ck(self._L.swcu_pl_symba_kick_list(self._h, nenc, _ptr(index1), _ptr(index2), _ptr(lactive), _ptr(levelg),
                                                 float(dt), int(irec), int(sgn), _ptr(lgood)))
        return lgood

    def tp_symba_kick_list(self, index1, index2, lactive, levelg_pl, levelg_tp, dt, irec, sgn, want_lgood=True):
        index1, index2 = _vec(index1, dt=_i32), _vec(index2, dt=_i32)
        nenc = len(index1)
        lactive = None if lactive is None else _vec(lactive, nenc, _i32)
        levelg_pl, levelg_tp = _vec(levelg_pl, dt=_i32), _vec(levelg_tp, dt=_i32)
        lgood = np.zeros(nenc, _i32) if want_lgood else None
        self._ck(self._L.swcu_tp_symba_kick_list(self._h, nenc, _ptr(index1), _ptr(index2), _ptr(lactive), _ptr(levelg_pl),
                                                 _ptr(levelg_tp), float(dt), int(irec), int(sgn), _ptr(lgood)))
        return lgood

    def body_symba_encounter_check_list(self, kind, index1, index2, lencmask, dt, lvdotr=None):
        """Pair loop of symba_encounter_check_list_plpl (kind PL) / _pltp (kind TP) on the resident populations."""
        index1, index2 = _vec(index1, dt=_i32), _vec(index2, dt=_i32)
        nenc = len(index1)
        lencmask = None if lencmask is None else _vec(lencmask, nenc, _i32)
        lenc = np.zeros(nenc, _i32)
        lvd = np.zeros(nenc, _i32) if lvdotr is None else _vec(lvdotr, nenc, _i32).copy()
        nf = C.c_int64()
        self._ck(self._L.swcu_body_symba_encounter_check_list(self._h, int(kind), nenc, _ptr(index1), _ptr(index2),
                                                              _ptr(lencmask), float(dt), _ptr(lenc), _ptr(lvd), C.byref(nf)))
        return lenc, lvd, nf.value

    def body_collision_check_list(self, kind, index1, index2, lmask, lvdotr, dt):
        index1, index2 = _vec(index1, dt=_i32), _vec(index2, dt=_i32)
        nenc = len(index1)
        lmask = None if lmask is None else _vec(lmask, nenc, _i32)
        lvdotr = _vec(lvdotr, nenc, _i32)
        lcol, lclo = np.zeros(nenc, _i32), np.zeros(nenc, _i32)
        nc = C.c_int64()
        self._ck(self._L.swcu_body_collision_check_list(self._h, int(kind), nenc, _ptr(index1), _ptr(index2), _ptr(lmask),
                                                        _ptr(lvdotr), float(dt), _ptr(lcol), _ptr(lclo), C.byref(nc)))
        return lcol, lclo, nc.value

    # ---- energy and momentum (swiftest_util.f90:1172-1394) ----
    def util_get_potential_energy(self, npl, lmask, GMcb, Gmass, mass, rb):
        lmask = None if lmask is None else _vec(lmask, npl, _i32)
        Gmass, mass, rb = _vec(Gmass, npl), _vec(mass, npl), _vec3(rb, npl)
        pe = C.c_double()
        self._ck(self._L.swcu_util_get_potential_energy(self._h, npl, _ptr(lmask), float(GMcb), _ptr(Gmass), _ptr(mass),
                                                        _ptr(rb), C.byref(pe)))
        return pe.value

    def util_get_energy_and_momentum(self, npl, lmask, GMcb, mass_cb, rbcb, vbcb, Gmass, mass, radius, rb, vb,
                                     lclose=True):
        """Returns dict(ke_orbit, pe, be, te, L_orbit(3), GMtot)."""
        lmask = None if lmask is None else _vec(lmask, npl, _i32)
        Gmass, mass, radius = _vec(Gmass, npl), _vec(mass, npl), _vec(radius, npl)
        rb, vb, rbcb, vbcb = _vec3(rb, npl), _vec3(vb, npl), _vec(rbcb, 3), _vec(vbcb, 3)
        out = np.zeros(8, _f64)
        self._ck(self._L.swcu_util_get_energy_and_momentum(
            self._h, npl, _ptr(lmask), float(GMcb), float(mass_cb), _ptr(rbcb), _ptr(vbcb), _ptr(Gmass), _ptr(mass),
            _ptr(radius), _ptr(rb), _ptr(vb), int(bool(lclose)), _ptr(out)))
        return dict(ke_orbit=out[0], pe=out[1], be=out[2], te=out[3], L_orbit=out[4:7].copy(), GMtot=out[7])

    # ---- multi-GPU ----
    def comm_unique_id(self):
        buf = (C.c_char * 128)()
        self._ck(self._L.swcu_comm_unique_id(self._h, buf))
        return bytes(buf)

    def comm_init(self, nranks, rank, unique_id):
        buf = (C.c_char * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        self._ck(self._L.swcu_comm_init(self._h, nranks, rank, buf))

    def comm_finalize(self):
        self._ck(self._L.swcu_comm_finalize(self._h))

    def pl_set_slice(self, i0, i1):
        self._ck(self._L.swcu_pl_set_slice(self._h, i0, i1))

    def pl_allgather(self, with_v=False):
        self._ck(self._L.swcu_pl_allgather(self._h, int(bool(with_v))))

    # peer-memory (CUDA IPC) fused step
    P2P_HANDLE_BYTES = 8 * 64

    def p2p_export(self):
        buf = (C.c_char * self.P2P_HANDLE_BYTES)()
        self._ck(self._L.swcu_p2p_export(self._h, buf))
        return bytes(buf)

    def p2p_import(self, nranks, rank, all_handles):
        """all_handles: the concatenation of every rank's p2p_export() bytes, in rank order."""
        if len(all_handles) != nranks * self.P2P_HANDLE_BYTES:
            raise ValueError("all_handles must hold nranks * 512 bytes")
        buf = (C.c_char * len(all_handles)).from_buffer_copy(all_handles)
        self._ck(self._L.swcu_p2p_import(self._h, nranks, rank, buf))

    def p2p_close(self):
        self._ck(self._L.swcu_p2p_close(self._h))

    def pl_kick_drift_p2p(self, dt, lclose=True, want_nfail=True):
        """One fused multi-GPU step of the resident pl population (third-law gravity on this rank's block pairs, then
        reduce-scatter + vb += ah*dt + Kepler drift + allgather in one kernel over NVLink peer memory)."""
        nf = C.c_int32()
        self._ck(self._L.swcu_pl_kick_drift_p2p(self._h, int(bool(lclose)), float(dt), C.byref(nf) if want_nfail else None))
        return nf.value

    # ---- measurement ----
    def timer_start(self):
        self._ck(self._L.swcu_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double()
        self._ck(self._L.swcu_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def timer_lap_begin(self):
        self._ck(self._L.swcu_timer_lap_begin(self._h))

    def timer_lap_end(self):
        self._ck(self._L.swcu_timer_lap_end(self._h))

    def timer_laps(self, each=0):
        """(total ms, count[, per-lap ms]) of the laps since the last call; synchronises."""
        ms, n = C.c_double(), C.c_int32()
        buf = (C.c_double * each)() if each else None
        self._ck(self._L.swcu_timer_laps(self._h, C.byref(ms), C.byref(n), buf, int(each)))
        if each:
            return ms.value, n.value, [buf[k] for k in range(min(each, n.value))]
        return ms.value, n.value

    def enable_kernel_timing(self, on=True):
        """False/0 off, True/1 last launch group per family, 2 accumulate every launch group (kernel_ms_accumulated)."""
        self._ck(self._L.swcu_enable_kernel_timing(self._h, int(on)))

    def kernel_ms_accumulated(self, family):
        ms, n = C.c_double(), C.c_int32()
        self._ck(self._L.swcu_kernel_ms_accumulated(self._h, family, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def last_kernel_ms(self, family):
        ms = C.c_double()
        self._ck(self._L.swcu_last_kernel_ms(self._h, family, C.byref(ms)))
        return ms.value

    def encounter_direct_count(self):
        """(pl-tp sweeps answered without the sort, how many of those had to be repeated on the sort path)."""
        a, b = C.c_int64(), C.c_int64()
        self._ck(self._L.swcu_encounter_direct_count(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def encounter_bucket_fallbacks(self):
        """Sort-and-sweep calls that were repeated with the radix sort (clump of equal radii / non-finite extent)."""
        a = C.c_int64()
        self._ck(self._L.swcu_encounter_bucket_fallbacks(self._h, C.byref(a)))
        return int(a.value)

    def step_graph_replays(self):
        """helio_step_pl calls that were replayed as one CUDA graph launch since create."""
        a = C.c_int64()
        self._ck(self._L.swcu_step_graph_replays(self._h, C.byref(a)))
        return int(a.value)

    def flat_redo_count(self):
        """Chunks the third-law gravity kernel rolled back and redid with the IEEE expression since create."""
        n = C.c_uint64()
        self._ck(self._L.swcu_flat_redo_count(self._h, C.byref(n)))
        return int(n.value)

    def probe_fp64_peak(self):
        t = C.c_double()
        self._ck(self._L.swcu_probe_fp64_peak(self._h, C.byref(t)))
        return t.value

    def probe_hbm_copy(self, nbytes=1 << 30):
        g = C.c_double()
        self._ck(self._L.swcu_probe_hbm_copy(self._h, int(nbytes), C.byref(g)))
        return g.value

    def flush_l2(self):
        self._ck(self._L.swcu_flush_l2(self._h))

    # ---- helpers ----
    @staticmethod
    def _inplace3(a, n):
        if not (isinstance(a, np.ndarray) and a.dtype == _f64 and a.flags.c_contiguous and a.shape == (n, 3)):
            raise ValueError("in/out arrays must be C-contiguous float64 of shape (n,3)")
