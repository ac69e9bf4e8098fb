/*
 * swiftest_oracle_step.c -- CPU restatement of the O(N) glue around the hot path and of the energy sums
 * (SURVEY.md section 8f, ranks 1 and 2).  TEST INFRASTRUCTURE ONLY, see swiftest_oracle.h.
 *
 * PARITY STATUS: PINNED bit for bit (round 2) to the reference's own Fortran executed by oracle/f90interp.py:
 * helio_step_pl/_tp over consecutive steps, energy and momentum (flat and triangular potential loops),
 * symba_kick_list_plpl/_pltp, symba_encounter_check_list_*, collision_check_one, discard_pl_close
 * (tests/golden/fortran_steps.npz, fortran_lists.npz; tests/test_oracle_fortran_goldens.py).  swo_discard_pl_tp's
 * bookkeeping loop and swo_collision_check_list's wrapper loop are restated by hand around pinned predicates.
 * swo_orbel_xv2aeq is also pinned to a and e from the reference's Python xv2el_one (swiftest/tool.py:377-455;
 * tests/golden/xv2aeq_ref.npz) at 1e-11; the system-level pin is the conservation thresholds of
 * tests/test_swiftest.py:119-121, which tests/test_oracle.py re-checks on the restated democratic-heliocentric step.
 *
 * Every sum below runs in the order the reference's serial loop / `sum` intrinsic runs it (first to last index unless
 * noted), one IEEE operation per Fortran operation (-ffp-contract=off).  `norm2` is restated as sqrt(x*x+y*y+z*z)
 * (libgfortran's norm2 rescales to avoid overflow and may differ in the last place; the CUDA path is compared with a
 * tolerance for every quantity that passes through a long sum).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "swiftest_oracle.h"

/* swiftest_util_coord_vh2vb_pl, swiftest/swiftest_util.f90:424-459 (no mask: every body counts) */
void swo_coord_vh2vb_pl(int32_t npl, double GMcb, const double *Gmass, const double *vh, double *vb, double *vbcb)
{
    if (npl <= 0) return;
    double s = 0.0;
    for (int32_t i = 0; i < npl; ++i) s += Gmass[i];
    const double Gmtot = GMcb + s;
    vbcb[0] = vbcb[1] = vbcb[2] = 0.0;
    for (int32_t i = 0; i < npl; ++i)
        for (int k = 0; k < 3; ++k) vbcb[k] = vbcb[k] - Gmass[i] * vh[3 * i + k];
    for (int k = 0; k < 3; ++k) vbcb[k] = vbcb[k] / Gmtot;
    for (int32_t i = 0; i < npl; ++i)
        for (int k = 0; k < 3; ++k) vb[3 * i + k] = vh[3 * i + k] + vbcb[k];
}

/* swiftest_util_coord_vb2vh_pl, swiftest_util.f90:363-395: the loop runs from npl DOWN to 1 and divides every term
 * by GMcb; lactive(i) = status(i) /= INACTIVE */
void swo_coord_vb2vh_pl(int32_t npl, double GMcb, const double *Gmass, const int32_t *lactive, const double *vb,
                        double *vh, double *vbcb)
{
    if (npl <= 0) return;
    vbcb[0] = vbcb[1] = vbcb[2] = 0.0;
    for (int32_t i = npl - 1; i >= 0; --i) {
        if (lactive && !lactive[i]) continue;
        for (int k = 0; k < 3; ++k) vbcb[k] = vbcb[k] - Gmass[i] * vb[3 * i + k] / GMcb;
    }
    for (int32_t i = 0; i < npl; ++i)
        for (int k = 0; k < 3; ++k) vh[3 * i + k] = vb[3 * i + k] - vbcb[k];
}

/* swiftest_util_coord_vh2vb_tp :462-485 and _vb2vh_tp :398-421 */
void swo_coord_vh2vb_tp(int32_t ntp, const int32_t *lmask, const double *vbcb, const double *vh, double *vb)
{
    for (int32_t i = 0; i < ntp; ++i) {
        if (lmask && !lmask[i]) continue;
        for (int k = 0; k < 3; ++k) vb[3 * i + k] = vh[3 * i + k] + vbcb[k];
    }
}
void swo_coord_vb2vh_tp(int32_t ntp, const int32_t *lmask, const double *vbcb, const double *vb, double *vh)
{
    for (int32_t i = 0; i < ntp; ++i) {
        if (lmask && !lmask[i]) continue;
        for (int k = 0; k < 3; ++k) vh[3 * i + k] = vb[3 * i + k] - vbcb[k];
    }
}

/* swiftest_util_coord_h2b_pl :228-265 (position and velocity; inactive bodies skipped) */
void swo_coord_h2b_pl(int32_t npl, double GMcb, const double *Gmass, const int32_t *lactive, const double *rh,
                      const double *vh, double *rb, double *vb, double *rbcb, double *vbcb)
{
    if (npl <= 0) return;
    double Gmtot = GMcb, xt[3] = {0, 0, 0}, vt[3] = {0, 0, 0};
    for (int32_t i = 0; i < npl; ++i) {
        if (lactive && !lactive[i]) continue;
        Gmtot = Gmtot + Gmass[i];
        for (int k = 0; k < 3; ++k) {
            xt[k] = xt[k] + Gmass[i] * rh[3 * i + k];
            vt[k] = vt[k] + Gmass[i] * vh[3 * i + k];
        }
    }
    for (int k = 0; k < 3; ++k) {
        rbcb[k] = -xt[k] / Gmtot;
        vbcb[k] = -vt[k] / Gmtot;
    }
    for (int32_t i = 0; i < npl; ++i) {
        if (lactive && !lactive[i]) continue;
        for (int k = 0; k < 3; ++k) {
            rb[3 * i + k] = rh[3 * i + k] + rbcb[k];
            vb[3 * i + k] = vh[3 * i + k] + vbcb[k];
        }
    }
}

/* helio_drift_linear_pl, helio/helio_drift.f90:129-165 (+ _one :82-103, _all :106-126) */
void swo_helio_drift_linear_pl(int32_t npl, double GMcb, const double *Gmass, const double *vb, const int32_t *lmask,
                               double dt, double *rh, double *pt)
{
    if (npl <= 0) return;
    for (int k = 0; k < 3; ++k) {
        double s = 0.0;
        for (int32_t i = 0; i < npl; ++i)
            if (!lmask || lmask[i]) s += Gmass[i] * vb[3 * i + k];
        pt[k] = s / GMcb;
    }
    for (int32_t i = 0; i < npl; ++i) {
        if (lmask && !lmask[i]) continue;
        for (int k = 0; k < 3; ++k) rh[3 * i + k] = rh[3 * i + k] + pt[k] * dt;
    }
}

/* helio_drift_linear_tp, helio_drift.f90:168-200 */
void swo_helio_drift_linear_tp(int32_t ntp, const int32_t *lmask, const double *pt, double dt, double *rh)
{
    for (int32_t i = 0; i < ntp; ++i) {
        if (lmask && !lmask[i]) continue;
        for (int k = 0; k < 3; ++k) rh[3 * i + k] = rh[3 * i + k] + pt[k] * dt;
    }
}

/* the velocity update of helio_kick_vb_pl / _tp, helio/helio_kick.f90:120-128, 160-165 */
void swo_helio_kick_vb(int32_t n, const int32_t *lmask, const double *ah, double dt, double *vb)
{
    for (int32_t i = 0; i < n; ++i) {
        if (lmask && !lmask[i]) continue;
        for (int k = 0; k < 3; ++k) vb[3 * i + k] = vb[3 * i + k] + ah[3 * i + k] * dt;
    }
}

/* helio_kick_vb_pl, helio_kick.f90:91-132: ah = 0; helio_kick_getacch_pl (:14-54, interaction term only);
 * set_beg_end; vb += ah*dt */
static void helio_kick_vb_pl(int32_t npl, const double *Gmass, const double *radius, int lflat, const int32_t *lmask,
                             double dt, const double *rh, double *vb, double *ah, double *rsave)
{
    memset(ah, 0, sizeof(double) * 3 * (size_t)npl);
    if (lflat) { /* swiftest_kick_getacch_int_pl, swiftest_kick.f90:29-33 */
        const int64_t nplpl = (int64_t)npl * (npl - 1) / 2;
        if (radius) swo_kick_flat_rad_pl(npl, nplpl, NULL, rh, Gmass, radius, ah);
        else swo_kick_flat_norad_pl(npl, nplpl, NULL, rh, Gmass, ah);
    } else { /* :35-39 */
        if (radius && swo_omp_kick_enabled()) swo_omp_kick_tri_rad_pl(npl, npl, rh, Gmass, radius, ah);
        else if (radius) swo_kick_tri_rad_pl(npl, npl, rh, Gmass, radius, ah);
        else swo_kick_tri_norad_pl(npl, npl, rh, Gmass, ah);
    }
    if (rsave) memcpy(rsave, rh, sizeof(double) * 3 * (size_t)npl);
    swo_helio_kick_vb(npl, lmask, ah, dt, vb);
}

/* helio_step_pl, helio/helio_step.f90:37-78 (no GR, no oblateness, no user force).
 * radius == NULL selects the norad variants (param%lclose false).  *lfirst is cleared after the first call. */
void swo_helio_step_pl(int32_t npl, double GMcb, const double *Gmass, const double *radius, int lflat,
                       const int32_t *lmask, int32_t *lfirst, double dt, double *rh, double *vh, double *vb,
                       double *ah, double *rbeg, double *rend, double *ptbeg, double *ptend, double *vbcb,
                       int32_t *iflag)
{
    if (npl <= 0) return;
    const double dth = 0.5 * dt;
    if (*lfirst) {
        swo_coord_vh2vb_pl(npl, GMcb, Gmass, vh, vb, vbcb);
        *lfirst = 0;
    }
    swo_helio_drift_linear_pl(npl, GMcb, Gmass, vb, lmask, dth, rh, ptbeg);
    helio_kick_vb_pl(npl, Gmass, radius, lflat, lmask, dth, rh, vb, ah, rbeg);
    { /* helio_drift_body, helio_drift.f90:14-54: mu(:) = cb%Gmass */
        double *mu = (double *)malloc(sizeof(double) * (size_t)npl);
        int32_t *mask1 = NULL;
        if (!lmask) {
            mask1 = (int32_t *)malloc(sizeof(int32_t) * (size_t)npl);
            for (int32_t i = 0; i < npl; ++i) mask1[i] = 1;
        }
        for (int32_t i = 0; i < npl; ++i) mu[i] = GMcb;
        for (int32_t i = 0; i < npl; ++i) iflag[i] = 0;
        swo_drift_all(mu, rh, vb, npl, 0, 0.0, dt, lmask ? lmask : mask1, iflag);
        free(mu);
        free(mask1);
    }
    helio_kick_vb_pl(npl, Gmass, radius, lflat, lmask, dth, rh, vb, ah, rend);
    swo_helio_drift_linear_pl(npl, GMcb, Gmass, vb, lmask, dth, rh, ptend);
    swo_coord_vb2vh_pl(npl, GMcb, Gmass, NULL, vb, vh, vbcb);
}

/* helio_step_tp, helio_step.f90:81-123 with helio_kick_vb_tp (helio_kick.f90:135-169) and helio_kick_getacch_tp
 * (:57-88, interaction term only): the begin kick uses pl%rbeg, the end kick pl%rend */
void swo_helio_step_tp(int32_t ntp, int32_t npl, double GMcb, const double *GMpl, const double *rbeg,
                       const double *rend, const double *ptbeg, const double *ptend, const int32_t *lmask,
                       int32_t *lfirst, double dt, double *rh, double *vh, double *vb, double *ah, int32_t *iflag)
{
    if (ntp <= 0) return;
    const double dth = 0.5 * dt;
    double m[3];
    int32_t *mask1 = NULL;
    if (!lmask) {
        mask1 = (int32_t *)malloc(sizeof(int32_t) * (size_t)ntp);
        for (int32_t i = 0; i < ntp; ++i) mask1[i] = 1;
        lmask = mask1;
    }
    if (*lfirst) {
        for (int k = 0; k < 3; ++k) m[k] = -ptbeg[k];
        swo_coord_vh2vb_tp(ntp, lmask, m, vh, vb);
        *lfirst = 0;
    }
    swo_helio_drift_linear_tp(ntp, lmask, ptbeg, dth, rh);
    memset(ah, 0, sizeof(double) * 3 * (size_t)ntp);
    if (npl > 0) swo_kick_all_tp(ntp, npl, rh, rbeg, GMpl, lmask, ah);
    swo_helio_kick_vb(ntp, lmask, ah, dth, vb);
    {
        double *mu = (double *)malloc(sizeof(double) * (size_t)ntp);
        for (int32_t i = 0; i < ntp; ++i) mu[i] = GMcb;
        for (int32_t i = 0; i < ntp; ++i) iflag[i] = 0;
        swo_drift_all(mu, rh, vb, ntp, 0, 0.0, dt, lmask, iflag);
        free(mu);
    }
    memset(ah, 0, sizeof(double) * 3 * (size_t)ntp);
    if (npl > 0) swo_kick_all_tp(ntp, npl, rh, rend, GMpl, lmask, ah);
    swo_helio_kick_vb(ntp, lmask, ah, dth, vb);
    swo_helio_drift_linear_tp(ntp, lmask, ptend, dth, rh);
    for (int k = 0; k < 3; ++k) m[k] = -ptend[k];
    swo_coord_vb2vh_tp(ntp, lmask, m, vb, vh);
    free(mask1);
}

/* whm_kick_getacch_ah0, whm/whm_kick.f90:124-149 */
void swo_whm_kick_getacch_ah0(int32_t n, const double *mu, const double *rhp, double *ah0)
{
    ah0[0] = ah0[1] = ah0[2] = 0.0;
    for (int32_t i = 0; i < n; ++i) {
        const double x = rhp[3 * i], y = rhp[3 * i + 1], z = rhp[3 * i + 2];
        const double r2 = x * x + y * y + z * z;
        const double ir3h = 1.0 / (r2 * sqrt(r2));
        const double fac = mu[i] * ir3h;
        ah0[0] = ah0[0] - fac * x;
        ah0[1] = ah0[1] - fac * y;
        ah0[2] = ah0[2] - fac * z;
    }
}

/* swiftest_util_get_potential_energy_triangular, swiftest_util.f90:1344-1394 */
void swo_get_potential_energy_tri(int32_t npl, const int32_t *lmask, double GMcb, const double *Gmass,
                                  const double *mass, const double *rb, double *pe_out)
{
    double pe = 0.0;
    for (int32_t i = 0; i < npl; ++i) {
        if (lmask && !lmask[i]) continue;
        double row = 0.0; /* sum(pepl(i+1:npl), lmask(i+1:npl)) */
        for (int32_t j = i + 1; j < npl; ++j) {
            if (lmask && !lmask[j]) continue;
            const double dx = rb[3 * i] - rb[3 * j], dy = rb[3 * i + 1] - rb[3 * j + 1], dz = rb[3 * i + 2] - rb[3 * j + 2];
            row += -(Gmass[i] * mass[j]) / sqrt(dx * dx + dy * dy + dz * dz);
        }
        pe = pe + row;
    }
    double cb = 0.0; /* sum(pecb(1:npl), lmask(1:npl)), pecb(i) = -GMcb*mass(i)/norm2(rb(:,i)) */
    for (int32_t i = 0; i < npl; ++i) {
        if (lmask && !lmask[i]) continue;
        const double x = rb[3 * i], y = rb[3 * i + 1], z = rb[3 * i + 2];
        cb += -GMcb * mass[i] / sqrt(x * x + y * y + z * z);
    }
    *pe_out = pe + cb;
}

/* swiftest_util_get_potential_energy_flat, swiftest_util.f90:1291-1341; k_plpl == NULL: canonical flattened order */
void swo_get_potential_energy_flat(int32_t npl, int64_t nplpl, const int32_t *k_plpl, const int32_t *lmask,
                                   double GMcb, const double *Gmass, const double *mass, const double *rb,
                                   double *pe_out)
{
    double pp = 0.0;
    for (int64_t k = 1; k <= nplpl; ++k) {
        int32_t i, j;
        if (k_plpl) {
            i = k_plpl[2 * (k - 1)];
            j = k_plpl[2 * (k - 1) + 1];
        } else {
            swo_flatten_k_to_ij(npl, k, &i, &j);
        }
        --i;
        --j;
        if (lmask && !(lmask[i] && lmask[j])) continue;
        const double dx = rb[3 * i] - rb[3 * j], dy = rb[3 * i + 1] - rb[3 * j + 1], dz = rb[3 * i + 2] - rb[3 * j + 2];
        pp += -(Gmass[i] * mass[j]) / sqrt(dx * dx + dy * dy + dz * dz);
    }
    double cb = 0.0;
    for (int32_t i = 0; i < npl; ++i) {
        if (lmask && !lmask[i]) continue;
        const double x = rb[3 * i], y = rb[3 * i + 1], z = rb[3 * i + 2];
        cb += -GMcb * mass[i] / sqrt(x * x + y * y + z * z);
    }
    *pe_out = pp + cb;
}

/* swiftest_util_get_energy_and_momentum_system, swiftest_util.f90:1172-1288, lrotation = .false., no oblateness.
 * out[0] ke_orbit, out[1] pe, out[2] be, out[3] te, out[4..6] L_orbit, out[7] GMtot */
void swo_get_energy_and_momentum(int32_t npl, const int32_t *lmask, double GMcb, double mass_cb, const double *rbcb,
                                 const double *vbcb, const double *Gmass, const double *mass, const double *radius,
                                 const double *rb, const double *vb, int lclose, int lflat, double *out)
{
    double gms = 0.0;
    for (int32_t i = 0; i < npl; ++i)
        if (!lmask || lmask[i]) gms += Gmass[i];
    out[7] = GMcb + gms;
    const double kecb = mass_cb * (vbcb[0] * vbcb[0] + vbcb[1] * vbcb[1] + vbcb[2] * vbcb[2]);
    double Lcb[3];
    Lcb[0] = mass_cb * (rbcb[1] * vbcb[2] - rbcb[2] * vbcb[1]);
    Lcb[1] = mass_cb * (rbcb[2] * vbcb[0] - rbcb[0] * vbcb[2]);
    Lcb[2] = mass_cb * (rbcb[0] * vbcb[1] - rbcb[1] * vbcb[0]);
    double ke = 0.0, L[3] = {0, 0, 0};
    for (int32_t i = 0; i < npl; ++i) {
        if (lmask && !lmask[i]) continue;
        const double *r = rb + 3 * i, *v = vb + 3 * i;
        const double h0 = r[1] * v[2] - r[2] * v[1];
        const double h1 = r[2] * v[0] - r[0] * v[2];
        const double h2 = r[0] * v[1] - r[1] * v[0];
        L[0] += mass[i] * h0;
        L[1] += mass[i] * h1;
        L[2] += mass[i] * h2;
        ke += mass[i] * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    }
    double pe;
    if (lflat) swo_get_potential_energy_flat(npl, (int64_t)npl * (npl - 1) / 2, NULL, lmask, GMcb, Gmass, mass, rb, &pe);
    else swo_get_potential_energy_tri(npl, lmask, GMcb, Gmass, mass, rb, &pe);
    out[0] = 0.5 * (kecb + ke);
    out[1] = pe;
    for (int k = 0; k < 3; ++k) out[4 + k] = Lcb[k] + L[k];
    double be = 0.0;
    if (lclose)
        for (int32_t i = 0; i < npl; ++i)
            if (!lmask || lmask[i]) be += -3 * Gmass[i] * mass[i] / (5 * radius[i]);
    out[2] = be;
    out[3] = out[0] + 0.0 + out[1] + out[2];
}

/* swiftest_discard_pl_close, swiftest/swiftest_discard.f90:295-337 */
void swo_discard_pl_close(const double *dx, const double *dv, double dt, double r2crit, int32_t *iflag, double *r2min)
{
    const double r2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
    *r2min = r2;
    if (r2 <= r2crit) {
        *iflag = 1;
    } else {
        const double vdotr = dx[0] * dv[0] + dx[1] * dv[1] + dx[2] * dv[2];
        if (vdotr > 0.0) {
            *iflag = 0;
        } else {
            const double v2 = dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2];
            const double tmin = -vdotr / v2;
            double m;
            if (tmin < dt) m = r2 - vdotr * vdotr / v2;
            else m = r2 + 2 * vdotr * dt + v2 * (dt * dt);
            m = m < r2 ? m : r2; /* min(r2min, r2) */
            *r2min = m;
            *iflag = (m <= r2crit) ? 1 : 0;
        }
    }
}

/* swiftest_discard_pl_tp, swiftest_discard.f90:244-292: for every ACTIVE test particle the FIRST planet (ascending j)
 * whose discard_pl_close test fires; iplanet[i] = that j (1-based) or 0.  Returns the number of particles discarded. */
int32_t swo_discard_pl_tp(int32_t ntp, int32_t npl, const double *rtp, const double *vtp, const int32_t *lactive,
                          const double *rpl, const double *vpl, const double *radius, double dt, int32_t *iplanet)
{
    int32_t nd = 0;
    for (int32_t i = 0; i < ntp; ++i) {
        iplanet[i] = 0;
        if (lactive && !lactive[i]) continue;
        for (int32_t j = 0; j < npl; ++j) {
            double dx[3], dv[3], r2min;
            int32_t isp;
            for (int k = 0; k < 3; ++k) {
                dx[k] = rtp[3 * i + k] - rpl[3 * j + k];
                dv[k] = vtp[3 * i + k] - vpl[3 * j + k];
            }
            swo_discard_pl_close(dx, dv, dt, radius[j] * radius[j], &isp, &r2min);
            if (isp != 0) {
                iplanet[i] = j + 1;
                ++nd;
                break;
            }
        }
    }
    return nd;
}

/* The pair loop of symba_encounter_check_list_plpl / _pltp, symba/symba_encounter_check.f90:122-137 / 197-211:
 * for the pairs of the mask: xr = r2(j) - r1(i), vr = v2(j) - v1(i), rcrit = renc1(i) + renc2(j) (renc2 == NULL: test
 * particles, 0), encounter_check_one, then drop physically overlapping pairs (rji2 > (radius1(i) + radius2(j))**2 must
 * hold; radius2 == NULL: 0).  lencounter[k] is written for every k (0 outside the mask), lvdotr[k] only inside it.
 * Returns the number of encounters. */
int64_t swo_symba_encounter_check_list(int64_t nenc, const int32_t *index1, const int32_t *index2,
                                       const int32_t *lencmask, const double *r1, const double *v1, const double *renc1,
                                       const double *radius1, const double *r2, const double *v2, const double *renc2,
                                       const double *radius2, double dt, int32_t *lencounter, int32_t *lvdotr)
{
    int64_t n = 0;
    for (int64_t k = 0; k < nenc; ++k) {
        lencounter[k] = 0;
        if (lencmask && !lencmask[k]) continue;
        const int32_t i = index1[k] - 1, j = index2[k] - 1;
        const double xr = r2[3 * j] - r1[3 * i], yr = r2[3 * j + 1] - r1[3 * i + 1], zr = r2[3 * j + 2] - r1[3 * i + 2];
        const double vxr = v2[3 * j] - v1[3 * i], vyr = v2[3 * j + 1] - v1[3 * i + 1], vzr = v2[3 * j + 2] - v1[3 * i + 2];
        const double rcrit12 = renc1[i] + (renc2 ? renc2[j] : 0.0);
        int32_t lenc, lvd;
        swo_encounter_check_one(xr, yr, zr, vxr, vyr, vzr, rcrit12, dt, &lenc, &lvd);
        lvdotr[k] = lvd;
        if (lenc) {
            const double rl = radius1[i] + (radius2 ? radius2[j] : 0.0);
            const double rlim2 = rl * rl;
            const double rji2 = xr * xr + yr * yr + zr * zr;
            lenc = rji2 > rlim2;
        }
        lencounter[k] = lenc;
        n += lenc ? 1 : 0;
    }
    return n;
}

/* x**n with an integer variable exponent as libgfortran evaluates it (_gfortran_pow_r8_i4: binary exponentiation) */
double swo_pow_r8_i4(double a, int32_t b)
{
    double pw = 1.0, x = a;
    int32_t n = b;
    if (n != 0) {
        uint32_t u;
        if (n < 0) {
            u = (uint32_t)(-n);
            x = pw / x;
        } else {
            u = (uint32_t)n;
        }
        for (;;) {
            if (u & 1u) pw *= x;
            u >>= 1;
            if (u) x *= x;
            else break;
        }
    }
    return pw;
}

#define SWO_RHSCALE 6.5     /* symba_module.f90:22 */
#define SWO_RSHELL 0.48075  /* symba_module.f90:23 */

/* the force factor of symba_kick_list_plpl / _pltp for one pair (symba/symba_kick.f90:180-201, 284-304):
 * returns 0 and *fac when the pair is kicked at this level, 1 when it lies inside the inner shell (r2 < rim1) */
static int symba_list_fac(double rhsum, double r2, int32_t irecl, double *fac)
{
    const double ri = (rhsum * rhsum) * (SWO_RHSCALE * SWO_RHSCALE) * swo_pow_r8_i4(SWO_RSHELL, 2 * irecl);
    const double rim1 = ri * (SWO_RSHELL * SWO_RSHELL);
    if (r2 < rim1) {
        *fac = 0.0;
        return 1;
    }
    if (r2 < ri) {
        const double ris = sqrt(ri);
        const double r = sqrt(r2);
        const double rr = (ris - r) / (ris * (1.0 - SWO_RSHELL));
        *fac = pow(r2, -1.5) * (1.0 - 3 * (rr * rr) + 2 * (rr * rr * rr));
    } else {
        *fac = 1.0 / (r2 * sqrt(r2));
    }
    return 0;
}

/* symba_kick_list_plpl, symba/symba_kick.f90:126-231.  lactive(k) = (status(k) == ACTIVE).  vb and ah are updated in
 * place exactly as the serial reference does; lgood (optional) returns the final lgoodlevel mask. */
void swo_symba_kick_list_plpl(int64_t nenc, const int32_t *index1, const int32_t *index2, const int32_t *lactive,
                              int32_t npl, const int32_t *levelg, const double *rh, const double *rhill,
                              const double *Gmass, double dt, int32_t irec, int32_t sgn, double *vb, double *ah,
                              int32_t *lgood_out)
{
    if (nenc == 0 || npl == 0) return;
    int32_t *lgood = (int32_t *)malloc(sizeof(int32_t) * (size_t)nenc);
    const int32_t irm1 = irec - 1;
    const int32_t irecl = (sgn < 0) ? irec - 1 : irec;
    int64_t ngood = 0;
    for (int64_t k = 0; k < nenc; ++k) {
        const int32_t i = index1[k] - 1, j = index2[k] - 1;
        lgood[k] = (levelg[i] >= irm1) && (levelg[j] >= irm1) && (!lactive || lactive[k]);
        ngood += lgood[k];
    }
    if (ngood > 0) {
        for (int64_t k = 0; k < nenc; ++k) {
            if (!lgood[k]) continue;
            const int32_t i = index1[k] - 1, j = index2[k] - 1;
            for (int c = 0; c < 3; ++c) ah[3 * i + c] = ah[3 * j + c] = 0.0;
        }
        for (int64_t k = 0; k < nenc; ++k) {
            if (!lgood[k]) continue;
            const int32_t i = index1[k] - 1, j = index2[k] - 1;
            double dx[3], fac;
            for (int c = 0; c < 3; ++c) dx[c] = rh[3 * j + c] - rh[3 * i + c];
            const double r2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
            if (symba_list_fac(rhill[i] + rhill[j], r2, irecl, &fac)) {
                lgood[k] = 0;
                continue;
            }
            const double faci = fac * Gmass[i], facj = fac * Gmass[j];
            for (int c = 0; c < 3; ++c) {
                ah[3 * i + c] = ah[3 * i + c] + facj * dx[c];
                ah[3 * j + c] = ah[3 * j + c] - faci * dx[c];
            }
        }
        const double sdt = sgn * dt;
        for (int64_t k = 0; k < nenc; ++k) {
            if (!lgood[k]) continue;
            const int32_t i = index1[k] - 1, j = index2[k] - 1;
            for (int c = 0; c < 3; ++c) {
                vb[3 * i + c] = vb[3 * i + c] + sdt * ah[3 * i + c];
                vb[3 * j + c] = vb[3 * j + c] + sdt * ah[3 * j + c];
                ah[3 * i + c] = 0.0;
                ah[3 * j + c] = 0.0;
            }
        }
    }
    if (lgood_out) memcpy(lgood_out, lgood, sizeof(int32_t) * (size_t)nenc);
    free(lgood);
}

/* symba_kick_list_pltp, symba_kick.f90:234-337: only the test particle (index2) is kicked */
void swo_symba_kick_list_pltp(int64_t nenc, const int32_t *index1, const int32_t *index2, const int32_t *lactive,
                              int32_t npl, int32_t ntp, const int32_t *levelg_pl, const int32_t *levelg_tp,
                              const double *rh_pl, const double *rhill, const double *Gmass, const double *rh_tp,
                              double dt, int32_t irec, int32_t sgn, double *vb_tp, double *ah_tp, int32_t *lgood_out)
{
    if (nenc == 0 || npl == 0 || ntp == 0) return;
    int32_t *lgood = (int32_t *)malloc(sizeof(int32_t) * (size_t)nenc);
    const int32_t irm1 = irec - 1;
    const int32_t irecl = (sgn < 0) ? irec - 1 : irec;
    int64_t ngood = 0;
    for (int64_t k = 0; k < nenc; ++k) {
        const int32_t i = index1[k] - 1, j = index2[k] - 1;
        lgood[k] = (levelg_pl[i] >= irm1) && (levelg_tp[j] >= irm1) && (!lactive || lactive[k]);
        ngood += lgood[k];
    }
    if (ngood > 0) {
        for (int64_t k = 0; k < nenc; ++k)
            if (lgood[k])
                for (int c = 0; c < 3; ++c) ah_tp[3 * (index2[k] - 1) + c] = 0.0;
        for (int64_t k = 0; k < nenc; ++k) {
            if (!lgood[k]) continue;
            const int32_t i = index1[k] - 1, j = index2[k] - 1;
            double dx[3], fac;
            for (int c = 0; c < 3; ++c) dx[c] = rh_tp[3 * j + c] - rh_pl[3 * i + c];
            const double r2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
            if (symba_list_fac(rhill[i], r2, irecl, &fac)) {
                lgood[k] = 0;
                continue;
            }
            const double faci = fac * Gmass[i];
            for (int c = 0; c < 3; ++c) ah_tp[3 * j + c] = ah_tp[3 * j + c] - faci * dx[c];
        }
        const double sdt = sgn * dt;
        for (int64_t k = 0; k < nenc; ++k) {
            if (!lgood[k]) continue;
            const int32_t j = index2[k] - 1;
            for (int c = 0; c < 3; ++c) {
                vb_tp[3 * j + c] = vb_tp[3 * j + c] + sdt * ah_tp[3 * j + c];
                ah_tp[3 * j + c] = 0.0;
            }
        }
    }
    if (lgood_out) memcpy(lgood_out, lgood, sizeof(int32_t) * (size_t)nenc);
    free(lgood);
}

/* swiftest_orbel_xv2aeq, swiftest/swiftest_orbel.f90:700-764 */
#define SWO_TINYVALUE 4.0e-15 /* swiftest_orbel.f90:11 */
void swo_orbel_xv2aeq(double mu, double rx, double ry, double rz, double vx, double vy, double vz, double *a, double *e,
                      double *q)
{
    *a = 0.0;
    *e = 0.0;
    *q = 0.0;
    const double r = sqrt(rx * rx + ry * ry + rz * rz);
    const double v2 = vx * vx + vy * vy + vz * vz;
    const double hx = ry * vz - rz * vy, hy = rz * vx - rx * vz, hz = rx * vy - ry * vx;
    const double h2 = hx * hx + hy * hy + hz * hz;
    if (h2 < 2.2250738585072014e-308) return; /* tiny(h2) */
    const double energy = 0.5 * v2 - mu / r;
    int type; /* -1 ellipse, 0 parabola, 1 hyperbola */
    double fac = 0.0;
    if (fabs(energy * r / mu) < sqrt(SWO_TINYVALUE)) {
        type = 0;
    } else {
        *a = -0.5 * mu / energy;
        if (*a < 0.0) {
            fac = -h2 / (mu * *a);
            type = (fac > SWO_TINYVALUE) ? 1 : 0;
        } else {
            type = -1;
        }
    }
    if (type == -1) {
        fac = 1.0 - h2 / (mu * *a);
        if (fac > SWO_TINYVALUE) *e = sqrt(fac);
        *q = *a * (1.0 - *e);
    } else if (type == 0) {
        *a = 0.5 * h2 / mu;
        *e = 1.0;
        *q = *a;
    } else {
        *e = sqrt(1.0 + fac);
        *q = *a * (1.0 - *e);
    }
}

/* collision_check_one, collision/collision_check.f90:15-58 */
void swo_collision_check_one(double xr, double yr, double zr, double vxr, double vyr, double vzr, double Gmtot,
                             double rlim, double dt, int32_t lvdotr, int32_t *lcollision, int32_t *lclosest)
{
    const double r2 = xr * xr + yr * yr + zr * zr;
    const double rlim2 = rlim * rlim;
    *lclosest = 0;
    if (r2 <= rlim2) {
        *lcollision = 1;
    } else {
        *lcollision = 0;
        const double vdotr = xr * vxr + yr * vyr + zr * vzr;
        if (lvdotr && (vdotr > 0.0)) {
            const double tcr2 = r2 / (vxr * vxr + vyr * vyr + vzr * vzr);
            const double dt2 = dt * dt;
            if (tcr2 <= dt2) {
                double a, e, q;
                swo_orbel_xv2aeq(Gmtot, xr, yr, zr, vxr, vyr, vzr, &a, &e, &q);
                *lcollision = (q < rlim);
            }
            *lclosest = !*lcollision;
        }
    }
}

/* the pair loop of collision_check_plpl (:96-110) / _pltp (:213-223): xr = r1(i) - r2(j), vr = v1(i) - v2(j);
 * plpl (Gmass2/radius2 given): rlim = radius1(i) + radius2(j), Gmtot = Gmass1(i) + Gmass2(j);
 * pltp (Gmass2 == NULL): rlim = radius1(i), Gmtot = Gmass1(i).  lclosest is cleared for every k first (:93). */
int64_t swo_collision_check_list(int64_t nenc, const int32_t *index1, const int32_t *index2, const int32_t *lmask,
                                 const int32_t *lvdotr, const double *r1, const double *v1, const double *Gmass1,
                                 const double *radius1, const double *r2, const double *v2, const double *Gmass2,
                                 const double *radius2, double dt, int32_t *lcollision, int32_t *lclosest)
{
    int64_t n = 0;
    for (int64_t k = 0; k < nenc; ++k) {
        lcollision[k] = 0;
        lclosest[k] = 0;
        if (lmask && !lmask[k]) continue;
        const int32_t i = index1[k] - 1, j = index2[k] - 1;
        const double xr = r1[3 * i] - r2[3 * j], yr = r1[3 * i + 1] - r2[3 * j + 1], zr = r1[3 * i + 2] - r2[3 * j + 2];
        const double vxr = v1[3 * i] - v2[3 * j], vyr = v1[3 * i + 1] - v2[3 * j + 1], vzr = v1[3 * i + 2] - v2[3 * j + 2];
        const double rlim = Gmass2 ? radius1[i] + radius2[j] : radius1[i];
        const double Gmtot = Gmass2 ? Gmass1[i] + Gmass2[j] : Gmass1[i];
        swo_collision_check_one(xr, yr, zr, vxr, vyr, vzr, Gmtot, rlim, dt, lvdotr[k], &lcollision[k], &lclosest[k]);
        n += lcollision[k] ? 1 : 0;
    }
    return n;
}
