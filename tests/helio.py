"""A democratic-heliocentric kick-drift-kick stepper over a pluggable hot-path backend (test infrastructure).

Mirrors helio_step_pl (reference helio/helio_step.f90:37-78): lindrift(dt/2), kick(dt/2), drift(dt), kick(dt/2),
lindrift(dt/2), with vh2vb / vb2vh (swiftest_util.f90:363-459), helio_kick_vb_pl (helio_kick.f90:91-132) and
helio_drift_linear_pl (helio_drift.f90:129-160).  The O(N) glue runs in numpy; the hot path (pl%accel_int and
pl%drift) goes through the backend: the CPU oracle, or the CUDA library through its C ABI.
"""
import numpy as np


class OracleBackend:
    def __init__(self, oracle):
        self.o = oracle

    def accel_int(self, rh, Gmass, radius, ah):
        ah[:] = self.o.kick_tri_pl(rh, Gmass, radius, ah)

    def drift(self, mu, rh, vb, dt):
        x, v, fl = self.o.drift_all(mu, rh, vb, dt)
        rh[:], vb[:] = x, v
        return fl


class GpuBackend:
    """Tier-1 (host pointer) C-ABI calls, exactly what the Fortran submodule bodies would issue."""

    def __init__(self, ctx):
        self.c = ctx

    def accel_int(self, rh, Gmass, radius, ah):
        n = len(Gmass)
        self.c.kick_getacch_int_all_tri_pl(n, n, rh, Gmass, radius, ah)

    def drift(self, mu, rh, vb, dt):
        n = len(mu)
        fl = np.zeros(n, np.int32)
        self.c.drift_all(mu, rh, vb, n, dt, np.ones(n, np.int32), fl)
        return fl


class HelioSystem:
    def __init__(self, cb_Gmass, Gmass, rh, vh, radius, backend):
        self.Gcb = float(cb_Gmass)
        self.Gm = np.ascontiguousarray(Gmass, dtype=np.float64)
        self.rh = np.ascontiguousarray(rh, dtype=np.float64).copy()
        self.vh = np.ascontiguousarray(vh, dtype=np.float64).copy()
        self.radius = np.ascontiguousarray(radius, dtype=np.float64)
        self.backend = backend
        self.n = len(self.Gm)
        self.ah = np.zeros((self.n, 3))
        self.vb = np.zeros((self.n, 3))
        self.vbcb = np.zeros(3)
        self.lfirst = True

    # swiftest_util_coord_vh2vb_pl (swiftest_util.f90:440-459)
    def vh2vb(self):
        Gmtot = self.Gcb + self.Gm.sum()
        self.vbcb = -(self.Gm[:, None] * self.vh).sum(0) / Gmtot
        self.vb = self.vh + self.vbcb

    # swiftest_util_coord_vb2vh_pl (swiftest_util.f90:363-395)
    def vb2vh(self):
        self.vbcb = -(self.Gm[:, None] * self.vb).sum(0) / self.Gcb
        self.vh = self.vb - self.vbcb

    def lindrift(self, dt):
        pt = (self.Gm[:, None] * self.vb).sum(0) / self.Gcb
        self.rh += pt * dt

    def kick(self, dt):
        self.ah[:] = 0.0
        self.backend.accel_int(self.rh, self.Gm, self.radius, self.ah)
        self.vb += self.ah * dt

    def drift(self, dt):
        mu = np.full(self.n, self.Gcb)
        fl = self.backend.drift(mu, self.rh, self.vb, dt)
        assert not np.any(fl), "Danby drift failed"

    def step(self, dt):
        dth = 0.5 * dt
        if self.lfirst:
            self.vh2vb()
            self.lfirst = False
        self.lindrift(dth)
        self.kick(dth)
        self.drift(dt)
        self.kick(dth)
        self.lindrift(dth)
        self.vb2vh()

    def energy_and_momentum(self):
        """Total energy and angular momentum in barycentric coordinates, in units of G (masses are G*m)."""
        Gmtot = self.Gcb + self.Gm.sum()
        rcb = -(self.Gm[:, None] * self.rh).sum(0) / Gmtot
        vcb = -(self.Gm[:, None] * self.vh).sum(0) / Gmtot
        rb, vb = self.rh + rcb, self.vh + vcb
        ke = 0.5 * (self.Gm * (vb ** 2).sum(1)).sum() + 0.5 * self.Gcb * (vcb ** 2).sum()
        pe = -(self.Gcb * self.Gm / np.linalg.norm(self.rh, axis=1)).sum()
        d = self.rh[:, None, :] - self.rh[None, :, :]
        dist = np.linalg.norm(d, axis=2)
        iu = np.triu_indices(self.n, 1)
        pe -= (self.Gm[iu[0]] * self.Gm[iu[1]] / dist[iu]).sum()
        L = (self.Gm[:, None] * np.cross(rb, vb)).sum(0) + self.Gcb * np.cross(rcb, vcb)
        return ke + pe, L
