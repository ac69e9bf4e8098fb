/*
 * swiftest_oracle_whm.c -- CPU restatement of the Wisdom-Holman (WHM) step around the hot path: Jacobi coordinate
 * changes, the ah0/ah1/ah2 terms, the planet and test-particle kick-drift-kick (BASELINE.json configs[1]).
 * TEST INFRASTRUCTURE ONLY, see swiftest_oracle.h.  PARITY STATUS: PINNED bit for bit (round 2): whm_step_pl + whm_step_tp
 * of the reference, executed from its Fortran source by oracle/f90interp.py over consecutive steps (tri and flat loops,
 * masked bodies; tests/golden/fortran_steps.npz, tests/test_oracle_fortran_goldens.py), plus the two-body / conservation
 * properties in tests/test_oracle.py.
 *
 * Reference (paths relative to src/): whm/whm_step.f90:37-100, whm/whm_kick.f90:14-314, whm/whm_coord.f90:14-115,
 * whm/whm_drift.f90:14-58, whm/whm_util.f90:117-198, swiftest/swiftest_util.f90:2121-2149.
 * The loops that carry a running sum (h2j, j2h, vh2vj, ah2, eta) are serial in the reference and restated serially.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "swiftest_oracle.h"

/* whm_util_set_mu_eta_pl, whm_util.f90:175-198 (+ swiftest_util_set_mu_pl): mu = GMcb + Gm, eta = running mass,
 * muj = GMcb * eta(i)/eta(i-1) */
void swo_whm_set_mu_eta(int32_t npl, double GMcb, const double *Gmass, double *mu, double *eta, double *muj)
{
    if (npl <= 0) return;
    for (int32_t i = 0; i < npl; ++i) mu[i] = GMcb + Gmass[i];
    eta[0] = GMcb + Gmass[0];
    muj[0] = eta[0];
    for (int32_t i = 1; i < npl; ++i) {
        eta[i] = eta[i - 1] + Gmass[i];
        muj[i] = GMcb * eta[i] / eta[i - 1];
    }
}

/* whm_coord_h2j_pl, whm_coord.f90:14-46 */
void swo_whm_coord_h2j(int32_t npl, const double *Gmass, const double *eta, const double *rh, const double *vh,
                       double *xj, double *vj)
{
    if (npl <= 0) return;
    double sumx[3] = {0, 0, 0}, sumv[3] = {0, 0, 0};
    for (int c = 0; c < 3; ++c) {
        xj[c] = rh[c];
        vj[c] = vh[c];
    }
    for (int32_t i = 1; i < npl; ++i)
        for (int c = 0; c < 3; ++c) {
            sumx[c] = sumx[c] + Gmass[i - 1] * rh[3 * (i - 1) + c];
            sumv[c] = sumv[c] + Gmass[i - 1] * vh[3 * (i - 1) + c];
            const double cap = sumx[c] / eta[i - 1], capv = sumv[c] / eta[i - 1];
            xj[3 * i + c] = rh[3 * i + c] - cap;
            vj[3 * i + c] = vh[3 * i + c] - capv;
        }
}

/* whm_coord_j2h_pl, whm_coord.f90:49-80 */
void swo_whm_coord_j2h(int32_t npl, const double *Gmass, const double *eta, const double *xj, const double *vj,
                       double *rh, double *vh)
{
    if (npl <= 0) return;
    double sumx[3] = {0, 0, 0}, sumv[3] = {0, 0, 0};
    for (int c = 0; c < 3; ++c) {
        rh[c] = xj[c];
        vh[c] = vj[c];
    }
    for (int32_t i = 1; i < npl; ++i)
        for (int c = 0; c < 3; ++c) {
            sumx[c] = sumx[c] + Gmass[i - 1] * xj[3 * (i - 1) + c] / eta[i - 1];
            sumv[c] = sumv[c] + Gmass[i - 1] * vj[3 * (i - 1) + c] / eta[i - 1];
            rh[3 * i + c] = xj[3 * i + c] + sumx[c];
            vh[3 * i + c] = vj[3 * i + c] + sumv[c];
        }
}

/* whm_coord_vh2vj_pl, whm_coord.f90:83-113 */
void swo_whm_coord_vh2vj(int32_t npl, const double *Gmass, const double *eta, const double *vh, double *vj)
{
    if (npl <= 0) return;
    double sumv[3] = {0, 0, 0};
    for (int c = 0; c < 3; ++c) vj[c] = vh[c];
    for (int32_t i = 1; i < npl; ++i)
        for (int c = 0; c < 3; ++c) {
            sumv[c] = sumv[c] + Gmass[i - 1] * vh[3 * (i - 1) + c];
            vj[3 * i + c] = vh[3 * i + c] - sumv[c] / eta[i - 1];
        }
}

/* whm_util_set_ir3j, whm_util.f90:117-140 */
static void whm_set_ir3(int32_t npl, const double *rh, const double *xj, double *ir3h, double *ir3j)
{
    for (int32_t i = 0; i < npl; ++i) {
        double r2 = rh[3 * i] * rh[3 * i] + rh[3 * i + 1] * rh[3 * i + 1] + rh[3 * i + 2] * rh[3 * i + 2];
        double ir = 1.0 / sqrt(r2);
        ir3h[i] = ir / r2;
        r2 = xj[3 * i] * xj[3 * i] + xj[3 * i + 1] * xj[3 * i + 1] + xj[3 * i + 2] * xj[3 * i + 2];
        ir = 1.0 / sqrt(r2);
        ir3j[i] = ir / r2;
    }
}

/* whm_kick_getacch_pl, whm_kick.f90:14-67 (no oblateness, GR or user force): ah += ah0 + ah1 + ah2 + interaction term */
void swo_whm_kick_getacch_pl(int32_t npl, double GMcb, const double *Gmass, const double *radius, int lflat,
                             const int32_t *lmask, const double *rh, const double *xj, double *ah)
{
    if (npl <= 0) return;
    double *ir3h = (double *)malloc(sizeof(double) * (size_t)npl), *ir3j = (double *)malloc(sizeof(double) * (size_t)npl);
    whm_set_ir3(npl, rh, xj, ir3h, ir3j);
    double ah0[3];
    swo_whm_kick_getacch_ah0(npl - 1, Gmass + 1, rh + 3, ah0); /* bodies 2..npl (:33) */
    for (int32_t i = 0; i < npl; ++i)
        for (int c = 0; c < 3; ++c) ah[3 * i + c] = ah[3 * i + c] + ah0[c];
    /* ah1, :150-172 */
    for (int32_t i = 1; i < npl; ++i) {
        if (lmask && !lmask[i]) continue;
        for (int c = 0; c < 3; ++c) {
            const double ah1j = xj[3 * i + c] * ir3j[i], ah1h = rh[3 * i + c] * ir3h[i];
            ah[3 * i + c] = ah[3 * i + c] + GMcb * (ah1j - ah1h);
        }
    }
    /* ah2, :175-205: running sum over the Jacobi chain */
    {
        double ah2o[3] = {0, 0, 0}, etaj = GMcb;
        for (int32_t i = 1; i < npl; ++i) {
            if (lmask && !lmask[i]) continue;
            etaj = etaj + Gmass[i - 1];
            const double fac = Gmass[i] * GMcb * ir3j[i] / etaj;
            for (int c = 0; c < 3; ++c) {
                const double ah2 = ah2o[c] + fac * xj[3 * i + c];
                ah[3 * i + c] = ah[3 * i + c] + ah2;
                ah2o[c] = ah2;
            }
        }
    }
    /* pl%accel_int (:38) */
    if (lflat) {
        const int64_t nplpl = (int64_t)npl * (npl - 1) / 2;
        if (radius) swo_kick_flat_rad_pl(npl, nplpl, NULL, rh, Gmass, radius, ah);
        else swo_kick_flat_norad_pl(npl, nplpl, NULL, rh, Gmass, ah);
    } else {
        if (radius) swo_kick_tri_rad_pl(npl, npl, rh, Gmass, radius, ah);
        else swo_kick_tri_norad_pl(npl, npl, rh, Gmass, ah);
    }
    free(ir3h);
    free(ir3j);
}

/* whm_step_pl, whm_step.f90:37-69 with whm_kick_vh_pl (whm_kick.f90:208-262) and whm_drift_pl (whm_drift.f90:14-58).
 * State: rh, vh, xj, vj, ah (kept between steps), eta, muj (swo_whm_set_mu_eta), *lfirst; rbeg / rend receive the
 * positions the two kicks were evaluated at.  lmask may be NULL. */
void swo_whm_step_pl(int32_t npl, double GMcb, const double *Gmass, const double *radius, int lflat,
                     const int32_t *lmask, const double *eta, const double *muj, int32_t *lfirst, double dt, double *rh,
                     double *vh, double *xj, double *vj, double *ah, double *rbeg, double *rend, int32_t *iflag)
{
    if (npl <= 0) return;
    const double dth = 0.5 * dt;
    const size_t nb = sizeof(double) * 3 * (size_t)npl;
    int32_t *mask1 = NULL;
    if (!lmask) {
        mask1 = (int32_t *)malloc(sizeof(int32_t) * (size_t)npl);
        for (int32_t i = 0; i < npl; ++i) mask1[i] = 1;
        lmask = mask1;
    }
    /* kick(beg): the accelerations of the previous end-of-step are reused unless this is the first step */
    if (*lfirst) {
        swo_whm_coord_h2j(npl, Gmass, eta, rh, vh, xj, vj);
        memset(ah, 0, nb);
        swo_whm_kick_getacch_pl(npl, GMcb, Gmass, radius, lflat, lmask, rh, xj, ah);
        *lfirst = 0;
    }
    memcpy(rbeg, rh, nb);
    for (int32_t i = 0; i < npl; ++i)
        if (lmask[i])
            for (int c = 0; c < 3; ++c) vh[3 * i + c] = vh[3 * i + c] + ah[3 * i + c] * dth;
    swo_whm_coord_vh2vj(npl, Gmass, eta, vh, vj);
    for (int32_t i = 0; i < npl; ++i) iflag[i] = 0;
    swo_drift_all(muj, xj, vj, npl, 0, 0.0, dt, lmask, iflag);
    swo_whm_coord_j2h(npl, Gmass, eta, xj, vj, rh, vh);
    /* kick(end) */
    memset(ah, 0, nb);
    swo_whm_kick_getacch_pl(npl, GMcb, Gmass, radius, lflat, lmask, rh, xj, ah);
    memcpy(rend, rh, nb);
    for (int32_t i = 0; i < npl; ++i)
        if (lmask[i])
            for (int c = 0; c < 3; ++c) vh[3 * i + c] = vh[3 * i + c] + ah[3 * i + c] * dth;
    free(mask1);
}

/* whm_kick_getacch_tp, whm_kick.f90:70-121: ah += ah0(planets at rbeg or rend) + direct terms */
static void whm_getacch_tp(int32_t ntp, int32_t npl, const double *GMpl, const double *rpl, const int32_t *lmask,
                           const double *rh, double *ah)
{
    if (ntp == 0 || npl == 0) return;
    double ah0[3];
    swo_whm_kick_getacch_ah0(npl, GMpl, rpl, ah0);
    for (int32_t i = 0; i < ntp; ++i)
        if (lmask[i])
            for (int c = 0; c < 3; ++c) ah[3 * i + c] = ah[3 * i + c] + ah0[c];
    swo_kick_all_tp(ntp, npl, rh, rpl, GMpl, lmask, ah);
}

/* whm_step_tp, whm_step.f90:72-100 with whm_kick_vh_tp (whm_kick.f90:265-314): rbeg / rend are the planets' positions
 * at the two kicks (left by swo_whm_step_pl of the same step); tp%mu = GMcb */
void swo_whm_step_tp(int32_t ntp, int32_t npl, double GMcb, const double *GMpl, const double *rbeg, const double *rend,
                     const int32_t *lmask, int32_t *lfirst, double dt, double *rh, double *vh, double *ah, int32_t *iflag)
{
    if (ntp <= 0) return;
    const double dth = 0.5 * dt;
    int32_t *mask1 = NULL;
    if (!lmask) {
        mask1 = (int32_t *)malloc(sizeof(int32_t) * (size_t)ntp);
        for (int32_t i = 0; i < ntp; ++i) mask1[i] = 1;
        lmask = mask1;
    }
    if (*lfirst) {
        for (int32_t i = 0; i < ntp; ++i)
            if (lmask[i])
                for (int c = 0; c < 3; ++c) ah[3 * i + c] = 0.0;
        whm_getacch_tp(ntp, npl, GMpl, rbeg, lmask, rh, ah);
        *lfirst = 0;
    }
    for (int32_t i = 0; i < ntp; ++i)
        if (lmask[i])
            for (int c = 0; c < 3; ++c) vh[3 * i + c] = vh[3 * i + c] + ah[3 * i + c] * dth;
    {
        double *mu = (double *)malloc(sizeof(double) * (size_t)ntp);
        for (int32_t i = 0; i < ntp; ++i) {
            mu[i] = GMcb;
            iflag[i] = 0;
        }
        swo_drift_all(mu, rh, vh, ntp, 0, 0.0, dt, lmask, iflag);
        free(mu);
    }
    for (int32_t i = 0; i < ntp; ++i)
        if (lmask[i])
            for (int c = 0; c < 3; ++c) ah[3 * i + c] = 0.0;
    whm_getacch_tp(ntp, npl, GMpl, rend, lmask, rh, ah);
    for (int32_t i = 0; i < ntp; ++i)
        if (lmask[i])
            for (int c = 0; c < 3; ++c) vh[3 * i + c] = vh[3 * i + c] + ah[3 * i + c] * dth;
    free(mask1);
}
