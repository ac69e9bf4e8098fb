"""ctypes binding of oracle/libswiftest_oracle.so (the CPU restatement of the reference loops).

TEST INFRASTRUCTURE ONLY.  PARITY: pinned bit for bit to the reference's own Fortran statements executed by
oracle/f90interp.py (tests/golden/fortran_*.npz), drift also to the reference's Python two-body propagation
(see swiftest_oracle.h).  Arrays follow the Fortran layout r(3,n) == numpy shape (n,3) C-order.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f64 = np.float64
_i32 = np.int32


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _c(a, dt=_f64):
    return np.ascontiguousarray(a, dtype=dt)


def build(native=True):
    subprocess.check_call(["make", "-s", "-C", _HERE, "all" if native else "libswiftest_oracle.so"])


class Oracle:
    def __init__(self, native=False):
        name = "libswiftest_oracle_native.so" if native else "libswiftest_oracle.so"
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        d, i32, i64, p = C.c_double, C.c_int32, C.c_int64, C.c_void_p
        L.swo_nplplm.restype = i64
        L.swo_nplplm.argtypes = [i64, i64]
        for n in ("swo_encounter_sas_plpl", "swo_encounter_tri_plpl"):
            getattr(L, n).restype = i64
            getattr(L, n).argtypes = [i32, p, p, p, d]
        for n in ("swo_encounter_sas_pltp", "swo_encounter_tri_pltp"):
            getattr(L, n).restype = i64
            getattr(L, n).argtypes = [i32, i32, p, p, p, p, p, d]
        for n in ("swo_encounter_sas_plplm", "swo_encounter_all_plplm", "swo_encounter_tri_plplm"):
            getattr(L, n).restype = i64
            getattr(L, n).argtypes = [i32, i32, p, p, p, p, p, p, d]
        L.swo_encounter_last_nbox_total.restype = i64
        L.swo_encounter_fetch.argtypes = [p, p, p]
        L.swo_kick_flat_rad_pl.argtypes = [i32, i64, p, p, p, p, p]
        L.swo_kick_flat_norad_pl.argtypes = [i32, i64, p, p, p, p]
        L.swo_kick_tri_rad_pl.argtypes = [i32, i32, p, p, p, p]
        L.swo_kick_tri_norad_pl.argtypes = [i32, i32, p, p, p]
        L.swo_kick_tri_abs_scale.argtypes = [i32, i32, p, p, p, C.c_int, p]
        L.swo_kick_all_tp.argtypes = [i32, i32, p, p, p, p, p]
        L.swo_symba_kick_subtract_enc.argtypes = [i32, i64, p, p, p, p, p, p]
        L.swo_coord_vh2vb_pl.argtypes = [i32, d, p, p, p, p]
        L.swo_coord_vb2vh_pl.argtypes = [i32, d, p, p, p, p, p]
        L.swo_coord_vh2vb_tp.argtypes = [i32, p, p, p, p]
        L.swo_coord_vb2vh_tp.argtypes = [i32, p, p, p, p]
        L.swo_coord_h2b_pl.argtypes = [i32, d, p, p, p, p, p, p, p, p]
        L.swo_helio_drift_linear_pl.argtypes = [i32, d, p, p, p, d, p, p]
        L.swo_helio_drift_linear_tp.argtypes = [i32, p, p, d, p]
        L.swo_helio_kick_vb.argtypes = [i32, p, p, d, p]
        L.swo_helio_step_pl.argtypes = [i32, d, p, p, C.c_int, p, p, d, p, p, p, p, p, p, p, p, p, p]
        L.swo_helio_step_tp.argtypes = [i32, i32, d, p, p, p, p, p, p, p, d, p, p, p, p, p]
        L.swo_whm_kick_getacch_ah0.argtypes = [i32, p, p, p]
        L.swo_whm_set_mu_eta.argtypes = [i32, d, p, p, p, p]
        L.swo_whm_coord_h2j.argtypes = [i32, p, p, p, p, p, p]
        L.swo_whm_coord_j2h.argtypes = [i32, p, p, p, p, p, p]
        L.swo_whm_coord_vh2vj.argtypes = [i32, p, p, p, p]
        L.swo_whm_kick_getacch_pl.argtypes = [i32, d, p, p, C.c_int, p, p, p, p]
        L.swo_whm_step_pl.argtypes = [i32, d, p, p, C.c_int, p, p, p, p, d, p, p, p, p, p, p, p, p]
        L.swo_whm_step_tp.argtypes = [i32, i32, d, p, p, p, p, p, d, p, p, p, p]
        L.swo_symba_kick_list_plpl.argtypes = [i64, p, p, p, i32, p, p, p, p, d, i32, i32, p, p, p]
        L.swo_symba_kick_list_pltp.argtypes = [i64, p, p, p, i32, i32, p, p, p, p, p, p, d, i32, i32, p, p, p]
        L.swo_collision_check_list.restype = i64
        L.swo_collision_check_list.argtypes = [i64, p, p, p, p, p, p, p, p, p, p, p, p, d, p, p]
        L.swo_orbel_xv2aeq.argtypes = [d, d, d, d, d, d, d, p, p, p]
        L.swo_discard_pl_tp.restype = i32
        L.swo_discard_pl_tp.argtypes = [i32, i32, p, p, p, p, p, p, d, p]
        L.swo_symba_encounter_check_list.restype = i64
        L.swo_symba_encounter_check_list.argtypes = [i64, p, p, p, p, p, p, p, p, p, p, p, d, p, p]
        L.swo_get_potential_energy_tri.argtypes = [i32, p, d, p, p, p, p]
        L.swo_get_potential_energy_flat.argtypes = [i32, i64, p, p, d, p, p, p, p]
        L.swo_get_energy_and_momentum.argtypes = [i32, p, d, d, p, p, p, p, p, p, p, C.c_int, C.c_int, p]
        L.swo_omp_kick_flat_rad_pl.argtypes = [i32, i32, p, p, p, p]
        L.swo_omp_kick_tri_rad_pl.argtypes = [i32, i32, p, p, p, p]
        L.swo_omp_kick_tri_rad_pl_rows.argtypes = [i32, i32, i32, i32, p, p, p, p]
        L.swo_omp_kick_all_tp.argtypes = [i32, i32, p, p, p, p, p]
        L.swo_drift_all.argtypes = [p, p, p, i32, C.c_int, d, d, p, p]
        L.swo_omp_drift_all.argtypes = [p, p, p, i32, C.c_int, d, d, p, p]
        L.swo_drift_branch.restype = i32
        L.swo_drift_branch.argtypes = [d] * 8
        L.swo_symba_set_renc.argtypes = [i32, p, i32, p]
        L.swo_encounter_check_one.argtypes = [d] * 8 + [p, p]
        L.swo_flatten_k_to_ij.argtypes = [i32, i64, p, p]
        L.swo_flatten_ij_to_k.argtypes = [i32, i32, i32, p]
        L.swo_omp_max_threads.restype = C.c_int

    # ---------------- gravity ----------------
    def nplplm(self, npl, nplm):
        return int(self.lib.swo_nplplm(npl, nplm))

    def kick_flat_pl(self, r, Gmass, radius, acc, nplpl=None, k_plpl=None):
        """swiftest_kick_getacch_int_all_flat_{rad,norad}_pl; radius=None selects norad. Returns new acc."""
        r, Gmass, acc = _c(r), _c(Gmass), _c(acc).copy()
        npl = len(Gmass)
        kp = None
        if k_plpl is not None:
            k_plpl = _c(k_plpl, _i32)
            nplpl = k_plpl.shape[0]
            kp = k_plpl.ctypes.data
        elif nplpl is None:
            nplpl = npl * (npl - 1) // 2
        if radius is None:
            self.lib.swo_kick_flat_norad_pl(npl, nplpl, kp, r.ctypes.data, Gmass.ctypes.data, acc.ctypes.data)
        else:
            radius = _c(radius)
            self.lib.swo_kick_flat_rad_pl(npl, nplpl, kp, r.ctypes.data, Gmass.ctypes.data, radius.ctypes.data,
                                          acc.ctypes.data)
        return acc

    def kick_tri_pl(self, r, Gmass, radius, acc, nplm=None):
        r, Gmass, acc = _c(r), _c(Gmass), _c(acc).copy()
        npl = len(Gmass)
        nplm = npl if nplm is None else nplm
        if radius is None:
            self.lib.swo_kick_tri_norad_pl(npl, nplm, r.ctypes.data, Gmass.ctypes.data, acc.ctypes.data)
        else:
            radius = _c(radius)
            self.lib.swo_kick_tri_rad_pl(npl, nplm, r.ctypes.data, Gmass.ctypes.data, radius.ctypes.data,
                                         acc.ctypes.data)
        return acc

    def kick_tri_abs_scale(self, r, Gmass, radius, nplm=None):
        r, Gmass = _c(r), _c(Gmass)
        npl = len(Gmass)
        nplm = npl if nplm is None else nplm
        scale = np.zeros((npl, 3))
        rad = _c(radius) if radius is not None else np.zeros(npl)
        self.lib.swo_kick_tri_abs_scale(npl, nplm, r.ctypes.data, Gmass.ctypes.data, rad.ctypes.data,
                                        int(radius is not None), scale.ctypes.data)
        return scale

    def kick_all_tp(self, rtp, rpl, GMpl, lmask, acc):
        rtp, rpl, GMpl, acc = _c(rtp), _c(rpl), _c(GMpl), _c(acc).copy()
        lmask = _c(lmask, _i32)
        self.lib.swo_kick_all_tp(len(rtp), len(GMpl), rtp.ctypes.data, rpl.ctypes.data, GMpl.ctypes.data,
                                 lmask.ctypes.data, acc.ctypes.data)
        return acc

    def symba_kick_subtract_enc(self, index1, index2, rh, Gmass, radius, ah):
        rh, Gmass, radius, ah = _c(rh), _c(Gmass), _c(radius), _c(ah).copy()
        i1, i2 = _c(index1, _i32), _c(index2, _i32)
        self.lib.swo_symba_kick_subtract_enc(len(Gmass), len(i1), i1.ctypes.data, i2.ctypes.data, rh.ctypes.data,
                                             Gmass.ctypes.data, radius.ctypes.data, ah.ctypes.data)
        return ah

    def omp_kick_tri_rad_pl_rows(self, r, Gmass, radius, acc, nplm, i0, i1):
        self.lib.swo_omp_kick_tri_rad_pl_rows(len(Gmass), nplm, i0, i1, r.ctypes.data, Gmass.ctypes.data,
                                              radius.ctypes.data, acc.ctypes.data)

    def omp_kick_flat_rad_pl(self, r, Gmass, radius, acc, nplm):
        self.lib.swo_omp_kick_flat_rad_pl(len(Gmass), nplm, r.ctypes.data, Gmass.ctypes.data, radius.ctypes.data,
                                          acc.ctypes.data)

    def omp_kick_all_tp(self, rtp, rpl, GMpl, lmask, acc):
        self.lib.swo_omp_kick_all_tp(len(rtp), len(GMpl), rtp.ctypes.data, rpl.ctypes.data, GMpl.ctypes.data,
                                     lmask.ctypes.data, acc.ctypes.data)

    def omp_threads(self):
        return int(self.lib.swo_omp_max_threads())

    def omp_set_threads(self, n=None):
        """Thread count of the OpenMP loops; default = the CPUs this process may run on (ignores OMP_NUM_THREADS, which
        torchrun sets to 1 for its workers)."""
        import os
        n = len(os.sched_getaffinity(0)) if n is None else int(n)
        self.lib.swo_omp_set_num_threads(n)
        return self.omp_threads()

    def use_omp_kick(self, on=True):
        """Steppers: full-row pl-pl kick through the OpenMP row loop (bit-identical to the serial loop)."""
        self.lib.swo_use_omp_kick(int(bool(on)))

    # ---------------- drift ----------------
    def drift_all(self, mu, x, v, dt, lmask=None, lgr=False, inv_c2=0.0, omp=False):
        """swiftest_drift_all. Returns (x, v, iflag) new arrays."""
        x, v = _c(x).copy(), _c(v).copy()
        n = len(x)
        mu = np.full(n, mu, dtype=_f64) if np.isscalar(mu) else _c(mu)
        lmask = np.ones(n, _i32) if lmask is None else _c(lmask, _i32)
        iflag = np.zeros(n, _i32)
        fn = self.lib.swo_omp_drift_all if omp else self.lib.swo_drift_all
        fn(mu.ctypes.data, x.ctypes.data, v.ctypes.data, n, int(lgr), float(inv_c2), float(dt), lmask.ctypes.data,
           iflag.ctypes.data)
        return x, v, iflag

    def drift_branch(self, mu, x, v, dt):
        x, v = _c(x), _c(v)
        mu = np.full(len(x), mu) if np.isscalar(mu) else mu
        return np.array([self.lib.swo_drift_branch(float(mu[i]), *map(float, x[i]), *map(float, v[i]), float(dt))
                         for i in range(len(x))], dtype=_i32)

    # ---------------- encounters ----------------
    def set_renc(self, rhill, irec):
        rhill = _c(rhill)
        renc = np.empty_like(rhill)
        self.lib.swo_symba_set_renc(len(rhill), rhill.ctypes.data, irec, renc.ctypes.data)
        return renc

    def _fetch(self, nenc):
        i1, i2, lv = np.empty(nenc, _i32), np.empty(nenc, _i32), np.empty(nenc, _i32)
        if nenc:
            self.lib.swo_encounter_fetch(i1.ctypes.data, i2.ctypes.data, lv.ctypes.data)
        return i1, i2, lv

    def encounter_plpl(self, r, v, renc, dt, triangular=False):
        r, v, renc = _c(r), _c(v), _c(renc)
        fn = self.lib.swo_encounter_tri_plpl if triangular else self.lib.swo_encounter_sas_plpl
        nenc = fn(len(renc), r.ctypes.data, v.ctypes.data, renc.ctypes.data, float(dt))
        return self._fetch(nenc)

    def encounter_pltp(self, rpl, vpl, rtp, vtp, renc, dt, triangular=False):
        rpl, vpl, rtp, vtp, renc = _c(rpl), _c(vpl), _c(rtp), _c(vtp), _c(renc)
        fn = self.lib.swo_encounter_tri_pltp if triangular else self.lib.swo_encounter_sas_pltp
        nenc = fn(len(renc), len(rtp), rpl.ctypes.data, vpl.ctypes.data, rtp.ctypes.data, vtp.ctypes.data,
                  renc.ctypes.data, float(dt))
        return self._fetch(nenc)

    def encounter_plplm(self, rplm, vplm, rplt, vplt, rencm, renct, dt, merged=False, triangular=False):
        a = [_c(q) for q in (rplm, vplm, rplt, vplt, rencm, renct)]
        fn = self.lib.swo_encounter_all_plplm if merged else self.lib.swo_encounter_sas_plplm
        if triangular:
            fn = self.lib.swo_encounter_tri_plplm
        nenc = fn(len(a[4]), len(a[5]), *[q.ctypes.data for q in a], float(dt))
        return self._fetch(nenc)

    def nbox_total(self):
        return int(self.lib.swo_encounter_last_nbox_total())

    def encounter_check_one(self, xr, yr, zr, vxr, vyr, vzr, renc, dt):
        a, b = C.c_int32(0), C.c_int32(0)
        self.lib.swo_encounter_check_one(xr, yr, zr, vxr, vyr, vzr, renc, dt, C.byref(a), C.byref(b))
        return bool(a.value), bool(b.value)

    # ---- swiftest_oracle_step.c: integrator glue and energy sums ----
    @staticmethod
    def _a(x):
        return None if x is None else x.ctypes.data

    def coord_vh2vb_pl(self, GMcb, Gmass, vh):
        Gmass, vh = _c(Gmass), _c(vh)
        vb, vbcb = np.zeros_like(vh), np.zeros(3)
        self.lib.swo_coord_vh2vb_pl(len(Gmass), GMcb, self._a(Gmass), self._a(vh), self._a(vb), self._a(vbcb))
        return vb, vbcb

    def coord_vb2vh_pl(self, GMcb, Gmass, vb, lactive=None):
        Gmass, vb = _c(Gmass), _c(vb)
        la = None if lactive is None else _c(lactive, _i32)
        vh, vbcb = np.zeros_like(vb), np.zeros(3)
        self.lib.swo_coord_vb2vh_pl(len(Gmass), GMcb, self._a(Gmass), self._a(la), self._a(vb), self._a(vh), self._a(vbcb))
        return vh, vbcb

    def coord_h2b_pl(self, GMcb, Gmass, rh, vh, lactive=None):
        Gmass, rh, vh = _c(Gmass), _c(rh), _c(vh)
        la = None if lactive is None else _c(lactive, _i32)
        rb, vb, rbcb, vbcb = np.zeros_like(rh), np.zeros_like(vh), np.zeros(3), np.zeros(3)
        self.lib.swo_coord_h2b_pl(len(Gmass), GMcb, self._a(Gmass), self._a(la), self._a(rh), self._a(vh), self._a(rb),
                                  self._a(vb), self._a(rbcb), self._a(vbcb))
        return rb, vb, rbcb, vbcb

    def helio_drift_linear_pl(self, GMcb, Gmass, vb, rh, dt, lmask=None):
        Gmass, vb, rh = _c(Gmass), _c(vb), _c(rh).copy()
        lm = None if lmask is None else _c(lmask, _i32)
        pt = np.zeros(3)
        self.lib.swo_helio_drift_linear_pl(len(Gmass), GMcb, self._a(Gmass), self._a(vb), self._a(lm), dt, self._a(rh),
                                           self._a(pt))
        return rh, pt

    def helio_kick_vb(self, ah, vb, dt, lmask=None):
        ah, vb = _c(ah), _c(vb).copy()
        lm = None if lmask is None else _c(lmask, _i32)
        self.lib.swo_helio_kick_vb(len(ah), self._a(lm), self._a(ah), dt, self._a(vb))
        return vb

    def helio_step_pl(self, st, GMcb, Gmass, radius, dt, lflat=False, lmask=None):
        """st: dict with rh, vh, vb, lfirst (updated in place).  Returns iflag."""
        n = len(Gmass)
        Gmass = _c(Gmass)
        radius = None if radius is None else _c(radius)
        lm = None if lmask is None else _c(lmask, _i32)
        for k in ("rh", "vh", "vb"):
            st[k] = _c(st[k])
        for k in ("ah", "rbeg", "rend"):
            st[k] = np.zeros((n, 3))
        for k in ("ptbeg", "ptend", "vbcb"):
            st.setdefault(k, np.zeros(3))
        lf = C.c_int32(int(st.get("lfirst", True)))
        iflag = np.zeros(n, _i32)
        self.lib.swo_helio_step_pl(n, GMcb, self._a(Gmass), self._a(radius), int(lflat), self._a(lm), C.byref(lf), dt,
                                   self._a(st["rh"]), self._a(st["vh"]), self._a(st["vb"]), self._a(st["ah"]),
                                   self._a(st["rbeg"]), self._a(st["rend"]), self._a(st["ptbeg"]), self._a(st["ptend"]),
                                   self._a(st["vbcb"]), self._a(iflag))
        st["lfirst"] = bool(lf.value)
        return iflag

    def helio_step_tp(self, st, pl, GMcb, GMpl, dt, lmask=None):
        """st: tp state dict (rh, vh, vb, lfirst); pl: the planets' state dict after helio_step_pl of the same step."""
        n = len(st["rh"])
        GMpl = _c(GMpl)
        lm = None if lmask is None else _c(lmask, _i32)
        for k in ("rh", "vh", "vb"):
            st[k] = _c(st[k])
        st["ah"] = np.zeros((n, 3))
        lf = C.c_int32(int(st.get("lfirst", True)))
        iflag = np.zeros(n, _i32)
        self.lib.swo_helio_step_tp(n, len(GMpl), GMcb, self._a(GMpl), self._a(pl["rbeg"]), self._a(pl["rend"]),
                                   self._a(pl["ptbeg"]), self._a(pl["ptend"]), self._a(lm), C.byref(lf), dt,
                                   self._a(st["rh"]), self._a(st["vh"]), self._a(st["vb"]), self._a(st["ah"]),
                                   self._a(iflag))
        st["lfirst"] = bool(lf.value)
        return iflag

    def whm_kick_getacch_ah0(self, mu, rhp):
        mu, rhp = _c(mu), _c(rhp)
        out = np.zeros(3)
        self.lib.swo_whm_kick_getacch_ah0(len(mu), self._a(mu), self._a(rhp), self._a(out))
        return out

    def get_potential_energy(self, GMcb, Gmass, mass, rb, lmask=None, flat=False):
        Gmass, mass, rb = _c(Gmass), _c(mass), _c(rb)
        lm = None if lmask is None else _c(lmask, _i32)
        n = len(Gmass)
        pe = C.c_double()
        if flat:
            self.lib.swo_get_potential_energy_flat(n, n * (n - 1) // 2, None, self._a(lm), GMcb, self._a(Gmass),
                                                   self._a(mass), self._a(rb), C.byref(pe))
        else:
            self.lib.swo_get_potential_energy_tri(n, self._a(lm), GMcb, self._a(Gmass), self._a(mass), self._a(rb),
                                                  C.byref(pe))
        return pe.value

    def get_energy_and_momentum(self, GMcb, mass_cb, rbcb, vbcb, Gmass, mass, radius, rb, vb, lmask=None, lclose=True,
                                flat=False):
        Gmass, mass, radius, rb, vb = _c(Gmass), _c(mass), _c(radius), _c(rb), _c(vb)
        rbcb, vbcb = _c(rbcb), _c(vbcb)
        lm = None if lmask is None else _c(lmask, _i32)
        out = np.zeros(8)
        self.lib.swo_get_energy_and_momentum(len(Gmass), self._a(lm), GMcb, mass_cb, self._a(rbcb), self._a(vbcb),
                                             self._a(Gmass), self._a(mass), self._a(radius), self._a(rb), self._a(vb),
                                             int(lclose), int(flat), self._a(out))
        return dict(ke_orbit=out[0], pe=out[1], be=out[2], te=out[3], L_orbit=out[4:7].copy(), GMtot=out[7])

    def discard_pl_tp(self, rtp, vtp, lactive, rpl, vpl, radius, dt):
        rtp, vtp, rpl, vpl, radius = _c(rtp), _c(vtp), _c(rpl), _c(vpl), _c(radius)
        la = None if lactive is None else _c(lactive, _i32)
        ipl = np.zeros(len(rtp), _i32)
        nd = self.lib.swo_discard_pl_tp(len(rtp), len(rpl), self._a(rtp), self._a(vtp), self._a(la), self._a(rpl),
                                        self._a(vpl), self._a(radius), float(dt), self._a(ipl))
        return ipl, int(nd)

    def symba_encounter_check_list(self, index1, index2, lencmask, r1, v1, renc1, radius1, dt, r2=None, v2=None,
                                   renc2=None, radius2=None, lvdotr=None):
        index1, index2 = _c(index1, _i32), _c(index2, _i32)
        nenc = len(index1)
        lm = None if lencmask is None else _c(lencmask, _i32)
        r1, v1, renc1, radius1 = _c(r1), _c(v1), _c(renc1), _c(radius1)
        if r2 is None:
            r2, v2, renc2, radius2 = r1, v1, renc1, radius1
        else:
            r2, v2 = _c(r2), _c(v2)
            renc2 = None if renc2 is None else _c(renc2)
            radius2 = None if radius2 is None else _c(radius2)
        lenc = np.zeros(nenc, _i32)
        lvd = np.zeros(nenc, _i32) if lvdotr is None else _c(lvdotr, _i32).copy()
        n = self.lib.swo_symba_encounter_check_list(nenc, self._a(index1), self._a(index2), self._a(lm), self._a(r1),
                                                    self._a(v1), self._a(renc1), self._a(radius1), self._a(r2), self._a(v2),
                                                    self._a(renc2), self._a(radius2), float(dt), self._a(lenc), self._a(lvd))
        return lenc, lvd, int(n)

    def symba_kick_list_plpl(self, index1, index2, lactive, levelg, rh, rhill, Gmass, dt, irec, sgn, vb, ah=None):
        index1, index2 = _c(index1, _i32), _c(index2, _i32)
        la = None if lactive is None else _c(lactive, _i32)
        levelg, rh, rhill, Gmass = _c(levelg, _i32), _c(rh), _c(rhill), _c(Gmass)
        vb = _c(vb).copy()
        ah = np.full_like(vb, 123.0) if ah is None else _c(ah).copy()
        lgood = np.zeros(len(index1), _i32)
        self.lib.swo_symba_kick_list_plpl(len(index1), self._a(index1), self._a(index2), self._a(la), len(rhill),
                                          self._a(levelg), self._a(rh), self._a(rhill), self._a(Gmass), float(dt), int(irec),
                                          int(sgn), self._a(vb), self._a(ah), self._a(lgood))
        return vb, lgood, ah

    def symba_kick_list_pltp(self, index1, index2, lactive, levelg_pl, levelg_tp, rh_pl, rhill, Gmass, rh_tp, dt, irec,
                             sgn, vb_tp):
        index1, index2 = _c(index1, _i32), _c(index2, _i32)
        la = None if lactive is None else _c(lactive, _i32)
        levelg_pl, levelg_tp = _c(levelg_pl, _i32), _c(levelg_tp, _i32)
        rh_pl, rhill, Gmass, rh_tp = _c(rh_pl), _c(rhill), _c(Gmass), _c(rh_tp)
        vb = _c(vb_tp).copy()
        ah = np.full_like(vb, 123.0)
        lgood = np.zeros(len(index1), _i32)
        self.lib.swo_symba_kick_list_pltp(len(index1), self._a(index1), self._a(index2), self._a(la), len(rhill), len(rh_tp),
                                          self._a(levelg_pl), self._a(levelg_tp), self._a(rh_pl), self._a(rhill),
                                          self._a(Gmass), self._a(rh_tp), float(dt), int(irec), int(sgn), self._a(vb),
                                          self._a(ah), self._a(lgood))
        return vb, lgood, ah

    def collision_check_list(self, index1, index2, lmask, lvdotr, r1, v1, Gmass1, radius1, dt, r2=None, v2=None):
        index1, index2 = _c(index1, _i32), _c(index2, _i32)
        lm = None if lmask is None else _c(lmask, _i32)
        lvdotr = _c(lvdotr, _i32)
        r1, v1, Gmass1, radius1 = _c(r1), _c(v1), _c(Gmass1), _c(radius1)
        if r2 is None:
            r2, v2, g2, rad2 = r1, v1, Gmass1, radius1
        else:
            r2, v2, g2, rad2 = _c(r2), _c(v2), None, None
        lcol, lclo = np.zeros(len(index1), _i32), np.zeros(len(index1), _i32)
        n = self.lib.swo_collision_check_list(len(index1), self._a(index1), self._a(index2), self._a(lm), self._a(lvdotr),
                                              self._a(r1), self._a(v1), self._a(Gmass1), self._a(radius1), self._a(r2),
                                              self._a(v2), self._a(g2), self._a(rad2), float(dt), self._a(lcol), self._a(lclo))
        return lcol, lclo, int(n)

    def orbel_xv2aeq(self, mu, r, v):
        a, e, q = C.c_double(), C.c_double(), C.c_double()
        self.lib.swo_orbel_xv2aeq(mu, r[0], r[1], r[2], v[0], v[1], v[2], C.byref(a), C.byref(e), C.byref(q))
        return a.value, e.value, q.value

    # ---- swiftest_oracle_whm.c: Wisdom-Holman step ----
    def whm_set_mu_eta(self, GMcb, Gmass):
        Gmass = _c(Gmass)
        n = len(Gmass)
        mu, eta, muj = np.zeros(n), np.zeros(n), np.zeros(n)
        self.lib.swo_whm_set_mu_eta(n, GMcb, self._a(Gmass), self._a(mu), self._a(eta), self._a(muj))
        return mu, eta, muj

    def whm_coord_h2j(self, Gmass, eta, rh, vh):
        Gmass, eta, rh, vh = _c(Gmass), _c(eta), _c(rh), _c(vh)
        xj, vj = np.zeros_like(rh), np.zeros_like(vh)
        self.lib.swo_whm_coord_h2j(len(Gmass), self._a(Gmass), self._a(eta), self._a(rh), self._a(vh), self._a(xj), self._a(vj))
        return xj, vj

    def whm_coord_j2h(self, Gmass, eta, xj, vj):
        Gmass, eta, xj, vj = _c(Gmass), _c(eta), _c(xj), _c(vj)
        rh, vh = np.zeros_like(xj), np.zeros_like(vj)
        self.lib.swo_whm_coord_j2h(len(Gmass), self._a(Gmass), self._a(eta), self._a(xj), self._a(vj), self._a(rh), self._a(vh))
        return rh, vh

    def whm_step_pl(self, st, GMcb, Gmass, radius, dt, lflat=False, lmask=None):
        """st: dict with rh, vh (+ xj, vj, ah, eta, muj, lfirst created on the first call); updated in place."""
        Gmass = _c(Gmass)
        n = len(Gmass)
        radius = None if radius is None else _c(radius)
        lm = None if lmask is None else _c(lmask, _i32)
        if "eta" not in st:
            _, st["eta"], st["muj"] = self.whm_set_mu_eta(GMcb, Gmass)
            st["xj"], st["vj"], st["ah"] = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
            st.setdefault("lfirst", True)
        for k in ("rh", "vh", "xj", "vj", "ah"):
            st[k] = _c(st[k])
        st["rbeg"], st["rend"] = np.zeros((n, 3)), np.zeros((n, 3))
        lf = C.c_int32(int(st["lfirst"]))
        iflag = np.zeros(n, _i32)
        self.lib.swo_whm_step_pl(n, GMcb, self._a(Gmass), self._a(radius), int(lflat), self._a(lm), self._a(st["eta"]),
                                 self._a(st["muj"]), C.byref(lf), dt, self._a(st["rh"]), self._a(st["vh"]), self._a(st["xj"]),
                                 self._a(st["vj"]), self._a(st["ah"]), self._a(st["rbeg"]), self._a(st["rend"]),
                                 self._a(iflag))
        st["lfirst"] = bool(lf.value)
        return iflag

    def whm_step_tp(self, st, pl, GMcb, GMpl, dt, lmask=None):
        """st: tp dict (rh, vh, ah, lfirst); pl: planets' dict after whm_step_pl of the same step (rbeg, rend)."""
        n = len(st["rh"])
        GMpl = _c(GMpl)
        lm = None if lmask is None else _c(lmask, _i32)
        st.setdefault("ah", np.zeros((n, 3)))
        st.setdefault("lfirst", True)
        for k in ("rh", "vh", "ah"):
            st[k] = _c(st[k])
        lf = C.c_int32(int(st["lfirst"]))
        iflag = np.zeros(n, _i32)
        self.lib.swo_whm_step_tp(n, len(GMpl), GMcb, self._a(GMpl), self._a(pl["rbeg"]), self._a(pl["rend"]), self._a(lm),
                                 C.byref(lf), dt, self._a(st["rh"]), self._a(st["vh"]), self._a(st["ah"]), self._a(iflag))
        st["lfirst"] = bool(lf.value)
        return iflag


_cache = {}


def load(native=False):
    if native not in _cache:
        _cache[native] = Oracle(native)
    return _cache[native]
