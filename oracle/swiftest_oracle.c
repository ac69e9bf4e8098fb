/*
 * swiftest_oracle.c -- CPU restatement of Swiftest's force-and-drift hot path (see swiftest_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY; never linked into or called by the CUDA product path.
 * PARITY: PINNED bit for bit to the outputs of the reference's own Fortran statements, executed from
 * /root/reference/src by the interpreter oracle/f90interp.py (tests/golden/fortran_*.npz,
 * tests/test_oracle_fortran_goldens.py; the reference cannot be compiled here, see swiftest_oracle.h);
 * drift additionally pinned to the reference's Python el2xv/xv2el propagation.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared (oracle/Makefile).
 * All "file:line" citations are relative to /root/reference/src.
 */
#include "swiftest_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* globals/globals_module.f90:33-38,135 */
static const double PIBY2 = 1.570796326794896619231321691639751442099;
static const double PI3BY2 = 4.712388980384689857693965074919254326296;
static const double TWOPI = 6.283185307179586476925286766559005768394;
static const double THIRD = 0.333333333333333333333333333333333333333;
static const double SIXTH = 0.166666666666666666666666666666666666667;
#define VSMALL (sqrt(DBL_MIN)) /* sqrt(TINY(1._DP)) */
/* encounter/encounter_module.f90:21 */
static const double RSWEEP_FACTOR = 1.1;
/* swiftest/swiftest_drift.f90:12-17 */
static const double E2MAX = 0.36, DM2MAX = 0.16, E2DM2MAX = 0.0016, DANBYB = 1.0e-13;
#define NLAG1 50
#define NLAG2 40
/* symba/symba_module.f90:22-23 */
static const double RHSCALE = 6.5, RSHELL = 0.48075;

/* ======================================================================================================
 * Gravity
 * ==================================================================================================== */

/* swiftest_kick.f90:418-446 swiftest_kick_getacch_int_one_pl */
void swo_kick_one_pl(double rji2, double xr, double yr, double zr, double Gmi, double Gmj, double *axi, double *ayi,
                     double *azi, double *axj, double *ayj, double *azj)
{
    double irij3 = 1.0 / (rji2 * sqrt(rji2));
    double faci = Gmi * irij3;
    double facj = Gmj * irij3;
    *axi = *axi + facj * xr;
    *ayi = *ayi + facj * yr;
    *azi = *azi + facj * zr;
    *axj = *axj - faci * xr;
    *ayj = *ayj - faci * yr;
    *azj = *azj - faci * zr;
}

/* swiftest_kick.f90:449-470 swiftest_kick_getacch_int_one_tp */
void swo_kick_one_tp(double rji2, double xr, double yr, double zr, double GMpl, double *ax, double *ay, double *az)
{
    double fac = GMpl / (rji2 * sqrt(rji2));
    *ax = *ax - fac * xr;
    *ay = *ay - fac * yr;
    *az = *az - fac * zr;
}

/* symba/symba_util.f90:202 */
int64_t swo_nplplm(int64_t npl, int64_t nplm) { return nplm * npl - nplm * (nplm + 1) / 2; }

/* swiftest_util.f90:1031-1053 */
void swo_flatten_ij_to_k(int32_t n, int32_t i, int32_t j, int64_t *k)
{
    int64_t i8 = i, j8 = j, n8 = n;
    *k = (i8 - 1) * n8 - i8 * (i8 - 1) / 2 + (j8 - i8);
}

/* swiftest_util.f90:1056-1087 */
void swo_flatten_k_to_ij(int32_t n, int64_t k, int32_t *i, int32_t *j)
{
    int64_t n8 = n;
    int64_t kp = n8 * (n8 - 1) / 2 - k;
    int64_t p = (int64_t)floor((sqrt(1.0 + 8.0 * (double)kp) - 1.0) / 2.0);
    int64_t i8 = n8 - 1 - p;
    int64_t j8 = k - (n8 - 1) * (n8 - 2) / 2 + p * (p + 1) / 2 + 1;
    *i = (int32_t)i8;
    *j = (int32_t)j8;
}

/* shared body of the two flat variants: swiftest_kick.f90:92-112 (rad) / 140-159 (norad), serial k order */
static void kick_flat(int32_t npl, int64_t nplpl, const int32_t *k_plpl, const double *r, const double *Gmass,
                      const double *radius, double *acc)
{
    double *ahi = (double *)calloc((size_t)3 * npl, sizeof(double));
    double *ahj = (double *)calloc((size_t)3 * npl, sizeof(double));
    int32_t ci = 1, cj = 1; /* running (i,j) of the canonical order when k_plpl is NULL */
    for (int64_t k = 1; k <= nplpl; ++k) {
        int32_t i, j;
        if (k_plpl) {
            i = k_plpl[2 * (k - 1)];
            j = k_plpl[2 * (k - 1) + 1];
        } else {
            cj += 1;
            if (cj > npl) {
                ci += 1;
                cj = ci + 1;
            }
            i = ci;
            j = cj;
        }
        const double *ri = r + 3 * (size_t)(i - 1), *rj = r + 3 * (size_t)(j - 1);
        double rx = rj[0] - ri[0];
        double ry = rj[1] - ri[1];
        double rz = rj[2] - ri[2];
        double rji2 = rx * rx + ry * ry + rz * rz;
        int go = 1;
        if (radius) {
            double rlim = radius[i - 1] + radius[j - 1];
            double rlim2 = rlim * rlim;
            go = (rji2 > rlim2);
        }
        if (go)
            swo_kick_one_pl(rji2, rx, ry, rz, Gmass[i - 1], Gmass[j - 1], &ahi[3 * (size_t)(i - 1)],
                            &ahi[3 * (size_t)(i - 1) + 1], &ahi[3 * (size_t)(i - 1) + 2], &ahj[3 * (size_t)(j - 1)],
                            &ahj[3 * (size_t)(j - 1) + 1], &ahj[3 * (size_t)(j - 1) + 2]);
    }
    for (size_t q = 0; q < (size_t)3 * npl; ++q) acc[q] = acc[q] + ahi[q] + ahj[q];
    free(ahi);
    free(ahj);
}

/* swiftest_kick.f90:69-115 */
void swo_kick_flat_rad_pl(int32_t npl, int64_t nplpl, const int32_t *k_plpl, const double *r, const double *Gmass,
                          const double *radius, double *acc)
{
    kick_flat(npl, nplpl, k_plpl, r, Gmass, radius, acc);
}

/* swiftest_kick.f90:118-162 */
void swo_kick_flat_norad_pl(int32_t npl, int64_t nplpl, const int32_t *k_plpl, const double *r, const double *Gmass,
                            double *acc)
{
    kick_flat(npl, nplpl, k_plpl, r, Gmass, NULL, acc);
}

/* shared body of the two triangular variants: swiftest_kick.f90:165-271 (rad) / 274-371 (norad) */
static void kick_tri(int32_t npl, int32_t nplm, const double *r, const double *Gmass, const double *radius,
                     double *acc)
{
    int32_t nplt = npl - nplm;
    int lmtiny = (nplt > nplm);
    if (lmtiny) { /* :189-217 upper triangle with ahi/ahj reduction */
        double *ahi = (double *)calloc((size_t)3 * npl, sizeof(double));
        double *ahj = (double *)calloc((size_t)3 * npl, sizeof(double));
        for (int32_t i = 1; i <= nplm; ++i) {
            for (int32_t j = i + 1; j <= npl; ++j) {
                const double *ri = r + 3 * (size_t)(i - 1), *rj = r + 3 * (size_t)(j - 1);
                double rx = rj[0] - ri[0];
                double ry = rj[1] - ri[1];
                double rz = rj[2] - ri[2];
                double rji2 = rx * rx + ry * ry + rz * rz;
                int go = 1;
                if (radius) {
                    double rlim = radius[i - 1] + radius[j - 1];
                    go = (rji2 > rlim * rlim);
                }
                if (go)
                    swo_kick_one_pl(rji2, rx, ry, rz, Gmass[i - 1], Gmass[j - 1], &ahi[3 * (size_t)(i - 1)],
                                    &ahi[3 * (size_t)(i - 1) + 1], &ahi[3 * (size_t)(i - 1) + 2],
                                    &ahj[3 * (size_t)(j - 1)], &ahj[3 * (size_t)(j - 1) + 1],
                                    &ahj[3 * (size_t)(j - 1) + 2]);
            }
        }
        for (size_t q = 0; q < (size_t)3 * npl; ++q) acc[q] = acc[q] + ahi[q] + ahj[q];
        free(ahi);
        free(ahj);
    } else { /* :218-265 full rows, ascending j, straight into acc */
        for (int32_t i = 1; i <= nplm; ++i) {
            const double *ri = r + 3 * (size_t)(i - 1);
            double *ai = acc + 3 * (size_t)(i - 1);
            for (int32_t j = 1; j <= npl; ++j) {
                if (j == i) continue;
                const double *rj = r + 3 * (size_t)(j - 1);
                double rx = rj[0] - ri[0];
                double ry = rj[1] - ri[1];
                double rz = rj[2] - ri[2];
                double rji2 = rx * rx + ry * ry + rz * rz;
                int go = 1;
                if (radius) {
                    double rlim = radius[i - 1] + radius[j - 1];
                    go = (rji2 > rlim * rlim);
                }
                if (go) {
                    double fac = Gmass[j - 1] / (rji2 * sqrt(rji2));
                    ai[0] = ai[0] + fac * rx;
                    ai[1] = ai[1] + fac * ry;
                    ai[2] = ai[2] + fac * rz;
                }
            }
        }
        if (nplt > 0) {
            for (int32_t i = nplm + 1; i <= npl; ++i) {
                const double *ri = r + 3 * (size_t)(i - 1);
                double *ai = acc + 3 * (size_t)(i - 1);
                for (int32_t j = 1; j <= nplm; ++j) {
                    const double *rj = r + 3 * (size_t)(j - 1);
                    double rx = rj[0] - ri[0];
                    double ry = rj[1] - ri[1];
                    double rz = rj[2] - ri[2];
                    double rji2 = rx * rx + ry * ry + rz * rz;
                    int go = 1;
                    if (radius) {
                        double rlim = radius[i - 1] + radius[j - 1];
                        go = (rji2 > rlim * rlim);
                    }
                    if (go) {
                        double fac = Gmass[j - 1] / (rji2 * sqrt(rji2));
                        ai[0] = ai[0] + fac * rx;
                        ai[1] = ai[1] + fac * ry;
                        ai[2] = ai[2] + fac * rz;
                    }
                }
            }
        }
    }
}

void swo_kick_tri_rad_pl(int32_t npl, int32_t nplm, const double *r, const double *Gmass, const double *radius,
                         double *acc)
{
    kick_tri(npl, nplm, r, Gmass, radius, acc);
}

void swo_kick_tri_norad_pl(int32_t npl, int32_t nplm, const double *r, const double *Gmass, double *acc)
{
    kick_tri(npl, nplm, r, Gmass, NULL, acc);
}

/* Sum over the same interactions of |fac*r_component|: the per-component magnitude against which the
 * north-star 1e-12 relative tolerance is stated (summation order differs between implementations). */
void swo_kick_tri_abs_scale(int32_t npl, int32_t nplm, const double *r, const double *Gmass, const double *radius,
                            int lrad, double *scale)
{
#pragma omp parallel for schedule(static)
    for (int32_t i = 1; i <= npl; ++i) {
        const double *ri = r + 3 * (size_t)(i - 1);
        double s0 = 0, s1 = 0, s2 = 0;
        int32_t jmax = (i <= nplm) ? npl : nplm;
        for (int32_t j = 1; j <= jmax; ++j) {
            if (j == i) continue;
            const double *rj = r + 3 * (size_t)(j - 1);
            double rx = rj[0] - ri[0], ry = rj[1] - ri[1], rz = rj[2] - ri[2];
            double rji2 = rx * rx + ry * ry + rz * rz;
            if (lrad) {
                double rlim = radius[i - 1] + radius[j - 1];
                if (!(rji2 > rlim * rlim)) continue;
            }
            double fac = Gmass[j - 1] / (rji2 * sqrt(rji2));
            s0 += fabs(fac * rx);
            s1 += fabs(fac * ry);
            s2 += fabs(fac * rz);
        }
        scale[3 * (size_t)(i - 1)] = s0;
        scale[3 * (size_t)(i - 1) + 1] = s1;
        scale[3 * (size_t)(i - 1) + 2] = s2;
    }
}

/* swiftest_kick.f90:374-415 */
void swo_kick_all_tp(int32_t ntp, int32_t npl, const double *rtp, const double *rpl, const double *GMpl,
                     const int32_t *lmask, double *acc)
{
    for (int32_t i = 1; i <= ntp; ++i) {
        if (!lmask[i - 1]) continue;
        const double *ri = rtp + 3 * (size_t)(i - 1);
        double *ai = acc + 3 * (size_t)(i - 1);
        for (int32_t j = 1; j <= npl; ++j) {
            const double *rj = rpl + 3 * (size_t)(j - 1);
            double rx = ri[0] - rj[0];
            double ry = ri[1] - rj[1];
            double rz = ri[2] - rj[2];
            double rji2 = rx * rx + ry * ry + rz * rz;
            swo_kick_one_tp(rji2, rx, ry, rz, GMpl[j - 1], &ai[0], &ai[1], &ai[2]);
        }
    }
}

/* symba/symba_kick.f90:59-70 : compute the encounter pairs again with flat_rad and subtract */
void swo_symba_kick_subtract_enc(int32_t npl, int64_t nenc, const int32_t *index1, const int32_t *index2,
                                 const double *rh, const double *Gmass, const double *radius, double *ah)
{
    if (nenc <= 0) return;
    double *ah_enc = (double *)calloc((size_t)3 * npl, sizeof(double));
    int32_t *k_enc = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)nenc);
    for (int64_t k = 0; k < nenc; ++k) {
        k_enc[2 * k] = index1[k];
        k_enc[2 * k + 1] = index2[k];
    }
    swo_kick_flat_rad_pl(npl, nenc, k_enc, rh, Gmass, radius, ah_enc);
    for (size_t q = 0; q < (size_t)3 * npl; ++q) ah[q] = ah[q] - ah_enc[q];
    free(ah_enc);
    free(k_enc);
}

/* ---------------- reference-shaped OpenMP loops (timed CPU baseline only) ---------------- */

/* thread count of the OpenMP loops, set explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers, which is not the
 * configuration the reference's OpenMP build runs in (bench.py passes the size of the process's CPU affinity set) */
void swo_omp_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* steppers (swiftest_oracle_step.c): run the full-row pl-pl kick through the OpenMP row loop below.  The reference's
 * loop IS an OpenMP do over i with every row summed by one thread in ascending j (swiftest_kick.f90:219-240), so the
 * result is bit-identical to the serial restatement for any thread count. */
static int swo_parallel_kick = 0;
void swo_use_omp_kick(int on) { swo_parallel_kick = on; }
int swo_omp_kick_enabled(void) { return swo_parallel_kick; }

int swo_omp_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* swiftest_kick.f90:95-112 loop shape: static split of the k range, thread-private ahi/ahj(3,npl) reduction.
 * The pair (i,j) of each k is generated on the fly (the reference's k_plpl table would be 40 GB at npl=1e5). */
void swo_omp_kick_flat_rad_pl(int32_t npl, int32_t nplm, const double *r, const double *Gmass, const double *radius,
                              double *acc)
{
    int64_t nplpl = swo_nplplm(npl, nplm);
    double *ahi = (double *)calloc((size_t)3 * npl, sizeof(double));
    double *ahj = (double *)calloc((size_t)3 * npl, sizeof(double));
#pragma omp parallel
    {
        double *pi = (double *)calloc((size_t)3 * npl, sizeof(double));
        double *pj = (double *)calloc((size_t)3 * npl, sizeof(double));
        int nt = 1, tid = 0;
#ifdef _OPENMP
        nt = omp_get_num_threads();
        tid = omp_get_thread_num();
#endif
        int64_t chunk = (nplpl + nt - 1) / nt;
        int64_t k0 = 1 + chunk * tid, k1 = k0 + chunk - 1;
        if (k1 > nplpl) k1 = nplpl;
        if (k0 <= k1) {
            int32_t i, j;
            swo_flatten_k_to_ij(npl, k0, &i, &j);
            /* guard the float sqrt in k_to_ij */
            int64_t kk;
            swo_flatten_ij_to_k(npl, i, j, &kk);
            while (kk > k0) { if (--j <= i) { --i; j = npl; } swo_flatten_ij_to_k(npl, i, j, &kk); }
            while (kk < k0) { if (++j > npl) { ++i; j = i + 1; } swo_flatten_ij_to_k(npl, i, j, &kk); }
            for (int64_t k = k0; k <= k1; ++k) {
                const double *ri = r + 3 * (size_t)(i - 1), *rj = r + 3 * (size_t)(j - 1);
                double rx = rj[0] - ri[0], ry = rj[1] - ri[1], rz = rj[2] - ri[2];
                double rji2 = rx * rx + ry * ry + rz * rz;
                double rlim = radius[i - 1] + radius[j - 1];
                if (rji2 > rlim * rlim)
                    swo_kick_one_pl(rji2, rx, ry, rz, Gmass[i - 1], Gmass[j - 1], &pi[3 * (size_t)(i - 1)],
                                    &pi[3 * (size_t)(i - 1) + 1], &pi[3 * (size_t)(i - 1) + 2],
                                    &pj[3 * (size_t)(j - 1)], &pj[3 * (size_t)(j - 1) + 1],
                                    &pj[3 * (size_t)(j - 1) + 2]);
                if (++j > npl) { ++i; j = i + 1; }
            }
        }
#pragma omp critical
        {
            for (size_t q = 0; q < (size_t)3 * npl; ++q) { ahi[q] += pi[q]; ahj[q] += pj[q]; }
        }
        free(pi);
        free(pj);
    }
    for (size_t q = 0; q < (size_t)3 * npl; ++q) acc[q] = acc[q] + ahi[q] + ahj[q];
    free(ahi);
    free(ahj);
}

/* swiftest_kick.f90:219-240 full-row branch, rows [i0,i1) (0-based half-open) of the first block, schedule(static) */
void swo_omp_kick_tri_rad_pl_rows(int32_t npl, int32_t nplm, int32_t i0, int32_t i1, const double *r,
                                  const double *Gmass, const double *radius, double *acc)
{
#pragma omp parallel for schedule(static)
    for (int32_t i = i0 + 1; i <= i1; ++i) {
        const double *ri = r + 3 * (size_t)(i - 1);
        double *ai = acc + 3 * (size_t)(i - 1);
        int32_t jmax = (i <= nplm) ? npl : nplm;
        double a0 = ai[0], a1 = ai[1], a2 = ai[2];
        double radi = radius[i - 1];
        for (int32_t j = 1; j <= jmax; ++j) {
            if (j == i) continue;
            const double *rj = r + 3 * (size_t)(j - 1);
            double rx = rj[0] - ri[0], ry = rj[1] - ri[1], rz = rj[2] - ri[2];
            double rji2 = rx * rx + ry * ry + rz * rz;
            double rlim = radi + radius[j - 1];
            if (rji2 > rlim * rlim) {
                double fac = Gmass[j - 1] / (rji2 * sqrt(rji2));
                a0 = a0 + fac * rx;
                a1 = a1 + fac * ry;
                a2 = a2 + fac * rz;
            }
        }
        ai[0] = a0;
        ai[1] = a1;
        ai[2] = a2;
    }
}

void swo_omp_kick_tri_rad_pl(int32_t npl, int32_t nplm, const double *r, const double *Gmass, const double *radius,
                             double *acc)
{
    swo_omp_kick_tri_rad_pl_rows(npl, nplm, 0, npl, r, Gmass, radius, acc);
}

/* swiftest_kick.f90:394-412 (the reference reduces the whole acc array per thread; rows are independent so
 * the arithmetic per tp is identical) */
void swo_omp_kick_all_tp(int32_t ntp, int32_t npl, const double *rtp, const double *rpl, const double *GMpl,
                         const int32_t *lmask, double *acc)
{
#pragma omp parallel for schedule(static)
    for (int32_t i = 1; i <= ntp; ++i) {
        if (!lmask[i - 1]) continue;
        const double *ri = rtp + 3 * (size_t)(i - 1);
        double *ai = acc + 3 * (size_t)(i - 1);
        for (int32_t j = 1; j <= npl; ++j) {
            const double *rj = rpl + 3 * (size_t)(j - 1);
            double rx = ri[0] - rj[0], ry = ri[1] - rj[1], rz = ri[2] - rj[2];
            double rji2 = rx * rx + ry * ry + rz * rz;
            swo_kick_one_tp(rji2, rx, ry, rz, GMpl[j - 1], &ai[0], &ai[1], &ai[2]);
        }
    }
}

/* ======================================================================================================
 * Drift
 * ==================================================================================================== */

/* swiftest_orbel.f90:147-172 */
void swo_orbel_scget(double angle, double *sx, double *cx)
{
    int32_t nper = (int32_t)(angle / TWOPI);
    double x = angle - nper * TWOPI;
    if (x < 0.0) x = x + TWOPI;
    *sx = sin(x);
    *cx = sqrt(1.0 - (*sx) * (*sx));
    if ((x > PIBY2) && (x < PI3BY2)) *cx = -(*cx);
}

/* swiftest_drift.f90:536-580 ; x is inout in the reference and is restored before return */
void swo_drift_kepu_stumpff(double *xio, double *c0, double *c1, double *c2, double *c3)
{
    double x = *xio;
    int32_t n = 0;
    const double xm = 0.1;
    while (fabs(x) >= xm) {
        n = n + 1;
        x = x / 4.0;
    }
    *c2 = (1.0 - x * (1.0 - x * (1.0 - x * (1.0 - x * (1.0 - x * (1.0 - x / 182.0) / 132.0) / 90.0) / 56.0) / 30.0) /
                     12.0) /
          2.0;
    *c3 = (1.0 - x * (1.0 - x * (1.0 - x * (1.0 - x * (1.0 - x * (1.0 - x / 210.0) / 156.0) / 110.0) / 72.0) / 42.0) /
                     20.0) /
          6.0;
    *c1 = 1.0 - x * (*c3);
    *c0 = 1.0 - x * (*c2);
    if (n != 0) {
        for (int32_t i = n; i >= 1; --i) {
            *c3 = (*c2 + (*c0) * (*c3)) / 4.0;
            *c2 = (*c1) * (*c1) / 2.0;
            *c1 = (*c0) * (*c1);
            *c0 = 2 * (*c0) * (*c0) - 1.0;
            x = x * 4;
        }
    }
    *xio = x;
}

/* swiftest_drift.f90:234-276 */
void swo_drift_kepmd(double dm, double es, double ec, double *xo, double *so, double *co)
{
    const double a0 = 39916800.0, a1 = 6652800.0, a2 = 332640.0, a3 = 7920.0, a4 = 110.0;
    double dx, fac1, fac2, q, y, f, fp, fpp, fppp, x, s, c;
    fac1 = 1.0 / (1.0 - ec);
    q = fac1 * dm;
    fac2 = es * es * fac1 - ec / 3.0;
    x = q * (1.0 - 0.5 * fac1 * q * (es - q * fac2));
    y = x * x;
    s = x * (a0 - y * (a1 - y * (a2 - y * (a3 - y * (a4 - y))))) / a0;
    c = sqrt(1.0 - s * s);
    f = x - ec * s + es * (1.0 - c) - dm;
    fp = 1.0 - ec * c + es * s;
    fpp = ec * s + es * c;
    fppp = ec * c - es * s;
    dx = -f / fp;
    dx = -f / (fp + dx * fpp / 2.0);
    dx = -f / (fp + dx * fpp / 2.0 + dx * dx * fppp * SIXTH);
    x = x + dx;
    y = x * x;
    s = x * (a0 - y * (a1 - y * (a2 - y * (a3 - y * (a4 - y))))) / a0;
    c = sqrt(1.0 - s * s);
    *xo = x;
    *so = s;
    *co = c;
}

/* swiftest_drift.f90:307-334 */
static void kepu_fchk(double dt, double r0, double mu, double alpha, double u, double s, double *f)
{
    double x, c0, c1, c2, c3;
    x = s * s * alpha;
    swo_drift_kepu_stumpff(&x, &c0, &c1, &c2, &c3);
    c1 = c1 * s;
    c2 = c2 * (s * s);
    c3 = c3 * (s * s * s);
    *f = r0 * c1 + u * c2 + mu * c3 - dt;
}

/* swiftest_drift.f90:486-533 */
static void kepu_p3solve(double dt, double r0, double mu, double alpha, double u, double *s, int32_t *iflag)
{
    double denom, a0, a1, a2, q, r, sq2, sq, p1, p2;
    denom = (mu - alpha * r0) * SIXTH;
    a2 = 0.5 * u / denom;
    a1 = r0 / denom;
    a0 = -dt / denom;
    q = (a1 - a2 * a2 * THIRD) * THIRD;
    r = (a1 * a2 - 3 * a0) * SIXTH - (a2 * a2 * a2) / 27.0;
    sq2 = q * q * q + r * r;
    if (sq2 >= 0.0) {
        sq = sqrt(sq2);
        if ((r + sq) <= 0.0)
            p1 = -pow(-(r + sq), THIRD);
        else
            p1 = pow(r + sq, THIRD);
        if ((r - sq) <= 0.0)
            p2 = -pow(-(r - sq), THIRD);
        else
            p2 = pow(r - sq, THIRD);
        *iflag = 0;
        *s = p1 + p2 - a2 * THIRD;
    } else {
        *iflag = 1;
        *s = 0.0;
    }
}

/* swiftest_drift.f90:337-378 */
static void kepu_guess(double dt, double r0, double mu, double alpha, double u, double *s)
{
    const double thresh = 0.4, danbyk = 0.85;
    int32_t iflag;
    double y, sy, cy, sigma, es, x, a, en, ec, e;
    if (alpha > 0.0) {
        if (dt / r0 <= thresh) {
            *s = dt / r0 - (dt * dt * u) / (2.0 * r0 * r0 * r0);
        } else {
            a = mu / alpha;
            en = sqrt(mu / (a * a * a));
            ec = 1.0 - r0 / a;
            es = u / (en * a * a);
            e = sqrt(ec * ec + es * es);
            y = en * dt - es;
            swo_orbel_scget(y, &sy, &cy);
            sigma = copysign(1.0, es * cy + ec * sy);
            x = y + sigma * danbyk * e;
            *s = x / sqrt(alpha);
        }
    } else {
        kepu_p3solve(dt, r0, mu, alpha, u, s, &iflag);
        if (iflag != 0) *s = dt / r0;
    }
}

/* swiftest_drift.f90:435-483 */
static void kepu_new(double *s, double dt, double r0, double mu, double alpha, double u, double *fp, double *c1,
                     double *c2, double *c3, int32_t *iflag)
{
    double x, c0, ds, f, fpp, fppp, fdt;
    for (int32_t nc = 0; nc <= 6; ++nc) {
        x = (*s) * (*s) * alpha;
        swo_drift_kepu_stumpff(&x, &c0, c1, c2, c3);
        *c1 = (*c1) * (*s);
        *c2 = (*c2) * (*s) * (*s);
        *c3 = (*c3) * (*s) * (*s) * (*s);
        f = r0 * (*c1) + u * (*c2) + mu * (*c3) - dt;
        *fp = r0 * c0 + u * (*c1) + mu * (*c2);
        fpp = (-r0 * alpha + mu) * (*c1) + u * c0;
        fppp = (-r0 * alpha + mu) * c0 - u * alpha * (*c1);
        ds = -f / (*fp);
        ds = -f / (*fp + ds * fpp / 2.0);
        ds = -f / (*fp + ds * fpp / 2.0 + ds * ds * fppp / 6.0);
        *s = *s + ds;
        fdt = f / dt;
        if (fdt * fdt < DANBYB * DANBYB) {
            *iflag = 0;
            return;
        }
    }
    *iflag = 1;
}

/* swiftest_drift.f90:381-432 */
static void kepu_lag(double *s, double dt, double r0, double mu, double alpha, double u, double *fp, double *c1,
                     double *c2, double *c3, int32_t *iflag)
{
    const int32_t ln = 5;
    int32_t ncmax;
    double x, fpp, ds, c0, f, fdt;
    if (alpha < 0.0)
        ncmax = NLAG2;
    else
        ncmax = NLAG1;
    for (int32_t nc = 0; nc <= ncmax; ++nc) {
        x = (*s) * (*s) * alpha;
        swo_drift_kepu_stumpff(&x, &c0, c1, c2, c3);
        *c1 = (*c1) * (*s);
        *c2 = (*c2) * (*s) * (*s);
        *c3 = (*c3) * (*s) * (*s) * (*s);
        f = r0 * (*c1) + u * (*c2) + mu * (*c3) - dt;
        *fp = r0 * c0 + u * (*c1) + mu * (*c2);
        fpp = (-r0 * alpha + mu) * (*c1) + u * c0;
        ds = -ln * f /
             (*fp + copysign(1.0, *fp) * sqrt(fabs((ln - 1.0) * (ln - 1.0) * (*fp) * (*fp) - (ln - 1.0) * ln * f * fpp)));
        *s = *s + ds;
        fdt = f / dt;
        if (fdt * fdt < DANBYB * DANBYB) {
            *iflag = 0;
            return;
        }
    }
    *iflag = 2;
}

/* swiftest_drift.f90:279-304 */
void swo_drift_kepu(double dt, double r0, double mu, double alpha, double u, double *fp, double *c1, double *c2,
                    double *c3, int32_t *iflag)
{
    double s, st, fo, fn;
    kepu_guess(dt, r0, mu, alpha, u, &s);
    st = s;
    kepu_new(&s, dt, r0, mu, alpha, u, fp, c1, c2, c3, iflag);
    if (*iflag != 0) {
        kepu_fchk(dt, r0, mu, alpha, u, st, &fo);
        kepu_fchk(dt, r0, mu, alpha, u, s, &fn);
        if (fabs(fo) < fabs(fn)) s = st;
        kepu_lag(&s, dt, r0, mu, alpha, u, fp, c1, c2, c3, iflag);
    }
}

/* swiftest_drift.f90:141-231 */
void swo_drift_dan(double mu, double *rx0, double *ry0, double *rz0, double *vx0, double *vy0, double *vz0,
                   double dt0, int32_t *iflag)
{
    double rx, ry, rz, vx, vy, vz, dt;
    double f, g, fdot, gdot, c1, c2, c3, u, alpha, fp, r0;
    double v0s, a, asq, en, dm, ec, es, esq, xkep, fchk, s, c;

    *iflag = 0;
    dt = dt0;
    r0 = sqrt((*rx0) * (*rx0) + (*ry0) * (*ry0) + (*rz0) * (*rz0));
    v0s = (*vx0) * (*vx0) + (*vy0) * (*vy0) + (*vz0) * (*vz0);
    u = (*rx0) * (*vx0) + (*ry0) * (*vy0) + (*rz0) * (*vz0);
    alpha = 2 * mu / r0 - v0s;
    if (alpha > 0.0) {
        a = mu / alpha;
        asq = a * a;
        en = sqrt(mu / (a * asq));
        ec = 1.0 - r0 / a;
        es = u / (en * asq);
        esq = ec * ec + es * es;
        dm = dt * en - (int32_t)(dt * en / TWOPI) * TWOPI;
        dt = dm / en;
        if ((esq < E2MAX) && (dm * dm < DM2MAX) && (esq * (dm * dm) < E2DM2MAX)) {
            swo_drift_kepmd(dm, es, ec, &xkep, &s, &c);
            fchk = (xkep - ec * s + es * (1.0 - c) - dm);
            if (fchk * fchk > DANBYB * DANBYB) {
                *iflag = 1;
                return;
            }
            fp = 1.0 - ec * c + es * s;
            f = a / r0 * (c - 1.0) + 1.0;
            g = dt + (s - xkep) / en;
            fdot = -(a / (r0 * fp)) * en * s;
            gdot = (c - 1.0) / fp + 1.0;
            rx = (*rx0) * f + (*vx0) * g;
            ry = (*ry0) * f + (*vy0) * g;
            rz = (*rz0) * f + (*vz0) * g;
            vx = (*rx0) * fdot + (*vx0) * gdot;
            vy = (*ry0) * fdot + (*vy0) * gdot;
            vz = (*rz0) * fdot + (*vz0) * gdot;
            *rx0 = rx;
            *ry0 = ry;
            *rz0 = rz;
            *vx0 = vx;
            *vy0 = vy;
            *vz0 = vz;
            *iflag = 0;
            return;
        }
    }

    swo_drift_kepu(dt, r0, mu, alpha, u, &fp, &c1, &c2, &c3, iflag);
    if (*iflag == 0) {
        f = 1.0 - mu / r0 * c2;
        g = dt - mu * c3;
        fdot = -mu / (fp * r0) * c1;
        gdot = 1.0 - mu / fp * c2;
        rx = (*rx0) * f + (*vx0) * g;
        ry = (*ry0) * f + (*vy0) * g;
        rz = (*rz0) * f + (*vz0) * g;
        vx = (*rx0) * fdot + (*vx0) * gdot;
        vy = (*ry0) * fdot + (*vy0) * gdot;
        vz = (*rz0) * fdot + (*vz0) * gdot;
        *rx0 = rx;
        *ry0 = ry;
        *rz0 = rz;
        *vx0 = vx;
        *vy0 = vy;
        *vz0 = vz;
    }
}

/* diagnostic: which branch of drift_dan / kepu_guess a body takes (for test bookkeeping of bit-exact coverage) */
int32_t swo_drift_branch(double mu, double rx, double ry, double rz, double vx, double vy, double vz, double dt)
{
    double r0 = sqrt(rx * rx + ry * ry + rz * rz);
    double v0s = vx * vx + vy * vy + vz * vz;
    double u = rx * vx + ry * vy + rz * vz;
    double alpha = 2 * mu / r0 - v0s;
    if (alpha > 0.0) {
        double a = mu / alpha, asq = a * a, en = sqrt(mu / (a * asq));
        double ec = 1.0 - r0 / a, es = u / (en * asq), esq = ec * ec + es * es;
        double dm = dt * en - (int32_t)(dt * en / TWOPI) * TWOPI;
        dt = dm / en;
        if ((esq < E2MAX) && (dm * dm < DM2MAX) && (esq * (dm * dm) < E2DM2MAX)) return 0;
        return (dt / r0 <= 0.4) ? 1 : 2;
    }
    return 3;
}

/* swiftest_drift.f90:111-138 */
void swo_drift_one(double mu, double *rx, double *ry, double *rz, double *vx, double *vy, double *vz, double dt,
                   int32_t *iflag)
{
    swo_drift_dan(mu, rx, ry, rz, vx, vy, vz, dt, iflag);
    if (*iflag != 0) {
        double dttmp = 0.1 * dt;
        for (int32_t i = 1; i <= 10; ++i) {
            swo_drift_dan(mu, rx, ry, rz, vx, vy, vz, dttmp, iflag);
            if (*iflag != 0) break;
        }
    }
}

/* GR step-size dilation, swiftest_drift.f90:84-97.  norm2 is restated as sqrt(x^2+y^2+z^2). */
static double drift_dtp(int lgr, double inv_c2, double dt, double mu, const double *x, const double *v)
{
    if (!lgr) return dt;
    double rmag = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    double vmag2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    double energy = 0.5 * vmag2 - mu / rmag;
    return dt * (1.0 + 3 * inv_c2 * energy);
}

/* swiftest_drift.f90:60-108 (serial loop, as in the reference) */
void swo_drift_all(const double *mu, double *x, double *v, int32_t n, int lgr, double inv_c2, double dt,
                   const int32_t *lmask, int32_t *iflag)
{
    if (n == 0) return;
    for (int32_t i = 0; i < n; ++i) {
        if (!lmask[i]) continue;
        double dtp = drift_dtp(lgr, inv_c2, dt, mu[i], x + 3 * (size_t)i, v + 3 * (size_t)i);
        swo_drift_one(mu[i], &x[3 * (size_t)i], &x[3 * (size_t)i + 1], &x[3 * (size_t)i + 2], &v[3 * (size_t)i],
                      &v[3 * (size_t)i + 1], &v[3 * (size_t)i + 2], dtp, &iflag[i]);
    }
}

/* the same loop threaded (the reference's loop is serial; reported separately) */
void swo_omp_drift_all(const double *mu, double *x, double *v, int32_t n, int lgr, double inv_c2, double dt,
                       const int32_t *lmask, int32_t *iflag)
{
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < n; ++i) {
        if (!lmask[i]) continue;
        double dtp = drift_dtp(lgr, inv_c2, dt, mu[i], x + 3 * (size_t)i, v + 3 * (size_t)i);
        swo_drift_one(mu[i], &x[3 * (size_t)i], &x[3 * (size_t)i + 1], &x[3 * (size_t)i + 2], &v[3 * (size_t)i],
                      &v[3 * (size_t)i + 1], &v[3 * (size_t)i + 2], dtp, &iflag[i]);
    }
}

/* ======================================================================================================
 * Encounter detection
 * ==================================================================================================== */

/* encounter_check.f90:573-621 */
void swo_encounter_check_one(double xr, double yr, double zr, double vxr, double vyr, double vzr, double renc,
                             double dt, int32_t *lencounter, int32_t *lvdotr)
{
    double r2crit, r2min, r2, v2, vdotr, tmin;
    r2 = xr * xr + yr * yr + zr * zr;
    r2crit = renc * renc;
    if (r2 > r2crit) {
        vdotr = vxr * xr + vyr * yr + vzr * zr;
        if (vdotr > 0.0) {
            r2min = r2;
        } else {
            v2 = vxr * vxr + vyr * vyr + vzr * vzr;
            if (v2 <= VSMALL) {
                r2min = r2;
            } else {
                tmin = -vdotr / v2;
                if (tmin < dt)
                    r2min = r2 - vdotr * vdotr / v2;
                else
                    r2min = r2 + 2 * vdotr * dt + v2 * (dt * dt);
            }
        }
    } else {
        vdotr = -1.0;
        r2min = r2;
    }
    *lvdotr = (vdotr < 0.0);
    *lencounter = *lvdotr && (r2min <= r2crit);
}

/* symba/symba_util.f90:245-267 */
void swo_symba_set_renc(int32_t npl, const double *rhill, int32_t irec, double *renc)
{
    double rshell_irec = 1.0;
    for (int32_t i = 1; i <= irec; ++i) rshell_irec = rshell_irec * RSHELL;
    for (int32_t i = 0; i < npl; ++i) renc[i] = rhill[i] * RHSCALE * rshell_irec;
}

/* ---- result buffer (stands in for the Fortran allocatable intent(out) arrays) ---- */
static uint64_t *g_keys = NULL;
static int64_t g_nkeys = 0, g_cap = 0;
static int64_t g_nbox_total = 0;

static void keys_reset(void) { g_nkeys = 0; }
static void keys_push(int32_t i1, int32_t i2)
{
    if (g_nkeys == g_cap) {
        g_cap = g_cap ? 2 * g_cap : 1024;
        g_keys = (uint64_t *)realloc(g_keys, sizeof(uint64_t) * (size_t)g_cap);
    }
    g_keys[g_nkeys++] = ((uint64_t)(uint32_t)i1 << 32) | (uint32_t)i2;
}
static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}
/* encounter_check.f90:676-760 remove_duplicates: sort by index1, sort index2 inside each group, drop equal
 * neighbours.  The outcome is the lexicographically sorted set of distinct (index1,index2). */
static void keys_sort_unique(void)
{
    if (g_nkeys == 0) return;
    qsort(g_keys, (size_t)g_nkeys, sizeof(uint64_t), cmp_u64);
    int64_t m = 1;
    for (int64_t k = 1; k < g_nkeys; ++k)
        if (g_keys[k] != g_keys[m - 1]) g_keys[m++] = g_keys[k];
    g_nkeys = m;
}

void swo_encounter_fetch(int32_t *index1, int32_t *index2, int32_t *lvdotr)
{
    for (int64_t k = 0; k < g_nkeys; ++k) {
        index1[k] = (int32_t)(g_keys[k] >> 32);
        index2[k] = (int32_t)(g_keys[k] & 0xffffffffu);
        if (lvdotr) lvdotr[k] = 1; /* lencounter = lvdotr .and. ... so every emitted pair has lvdotr true (F4) */
    }
}

int64_t swo_encounter_last_nbox_total(void) { return g_nbox_total; }

/* encounter_check.f90:763-792 sort_aabb_1D.  util_sort (base_module.f90:1361-1472) is an unstable quicksort
 * seeded with the previous call's permutation; only the relative order of EQUAL extents depends on it.
 * The oracle fixes that order as ascending extent-array position (begin endpoints 1..n before end endpoints
 * n+1..2n), i.e. a stable sort of the concatenated [rmin, rmax] array. */
typedef struct {
    double key;
    int32_t idx; /* 1..2n position in the extent array */
} ext_t;
static int cmp_ext(const void *a, const void *b)
{
    const ext_t *x = (const ext_t *)a, *y = (const ext_t *)b;
    if (x->key < y->key) return -1;
    if (x->key > y->key) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}
static void sort_aabb_1d(int32_t n, const double *rmin, const double *rmax, int32_t *ind, int64_t *ibeg,
                         int64_t *iend)
{
    ext_t *e = (ext_t *)malloc(sizeof(ext_t) * 2 * (size_t)n);
    for (int32_t i = 0; i < n; ++i) {
        e[i].key = rmin[i];
        e[i].idx = i + 1;
        e[n + i].key = rmax[i];
        e[n + i].idx = n + i + 1;
    }
    qsort(e, 2 * (size_t)n, sizeof(ext_t), cmp_ext);
    for (int64_t k = 1; k <= 2 * (int64_t)n; ++k) {
        int32_t i = e[k - 1].idx;
        ind[k - 1] = i;
        if (i <= n)
            ibeg[i - 1] = k;
        else
            iend[i - n - 1] = k;
    }
    free(e);
}

/* extents, encounter_check.f90:180-185 (norm2 restated as sqrt(x^2+y^2+z^2)) */
static void extents(int32_t n, const double *r, const double *renc, double *rmin, double *rmax)
{
    for (int32_t i = 0; i < n; ++i) {
        const double *ri = r + 3 * (size_t)i;
        double rmag = sqrt(ri[0] * ri[0] + ri[1] * ri[1] + ri[2] * ri[2]);
        double w = renc ? RSWEEP_FACTOR * renc[i] : RSWEEP_FACTOR * 0.0;
        rmax[i] = rmag + w;
        rmin[i] = rmag - w;
    }
}

/* encounter_check.f90:905-988 sweep_aabb_single_list (+ :329-381 sweep_one) */
static void sweep_single(int32_t n, const int32_t *ind, const int64_t *ibeg, const int64_t *iend, const double *r,
                         const double *v, const double *renc, double dt)
{
    for (int32_t i = 1; i <= n; ++i) {
        if (!((ibeg[i - 1] + 1) < (iend[i - 1] - 1))) continue; /* loverlap, :951 (F3) */
        int64_t kb = ibeg[i - 1] + 1, ke = iend[i - 1] - 1;
        g_nbox_total += ke - kb + 1;
        const double *ri = r + 3 * (size_t)(i - 1), *vi = v + 3 * (size_t)(i - 1);
        for (int64_t k = kb; k <= ke; ++k) {
            int32_t j = ind[k - 1] > n ? ind[k - 1] - n : ind[k - 1]; /* ext_ind :937-941 */
            const double *rj = r + 3 * (size_t)(j - 1), *vj = v + 3 * (size_t)(j - 1);
            double xr = rj[0] - ri[0], yr = rj[1] - ri[1], zr = rj[2] - ri[2];
            double vxr = vj[0] - vi[0], vyr = vj[1] - vi[1], vzr = vj[2] - vi[2];
            double renc12 = renc[i - 1] + renc[j - 1];
            int32_t lenc, lvd;
            swo_encounter_check_one(xr, yr, zr, vxr, vyr, vzr, renc12, dt, &lenc, &lvd);
            if (lenc) {
                if (i > j) /* :976-983 swap so that index1 < index2 */
                    keys_push(j, i);
                else
                    keys_push(i, j);
            }
        }
    }
    keys_sort_unique(); /* :985 */
}

/* encounter_check.f90:795-902 sweep_aabb_double_list */
static void sweep_double(int32_t n1, int32_t n2, const int32_t *ind, const int64_t *ibeg, const int64_t *iend,
                         const double *r1, const double *v1, const double *r2, const double *v2, const double *renc1,
                         const double *renc2, double dt)
{
    int32_t ntot = n1 + n2;
    for (int32_t i = 1; i <= ntot; ++i) {
        if (!((ibeg[i - 1] + 1) < (iend[i - 1] - 1))) continue; /* loverlap :828 */
        int64_t kb = ibeg[i - 1] + 1, ke = iend[i - 1] - 1;
        g_nbox_total += ke - kb + 1;
        int in1 = (i <= n1);
        int32_t ii = in1 ? i : i - n1;
        const double *ri = in1 ? r1 + 3 * (size_t)(ii - 1) : r2 + 3 * (size_t)(ii - 1);
        const double *vi = in1 ? v1 + 3 * (size_t)(ii - 1) : v2 + 3 * (size_t)(ii - 1);
        double renci = in1 ? renc1[ii - 1] : (renc2 ? renc2[ii - 1] : 0.0);
        for (int64_t k = kb; k <= ke; ++k) {
            int32_t e = ind[k - 1] > ntot ? ind[k - 1] - ntot : ind[k - 1];
            int jl1 = (e <= n1); /* llist1 :834 */
            if (jl1 == in1) continue; /* lgood mask: only bodies of the other list :873,:890 */
            int32_t j = jl1 ? e : e - n1;
            const double *rj = jl1 ? r1 + 3 * (size_t)(j - 1) : r2 + 3 * (size_t)(j - 1);
            const double *vj = jl1 ? v1 + 3 * (size_t)(j - 1) : v2 + 3 * (size_t)(j - 1);
            double rencj = jl1 ? renc1[j - 1] : (renc2 ? renc2[j - 1] : 0.0);
            double xr = rj[0] - ri[0], yr = rj[1] - ri[1], zr = rj[2] - ri[2];
            double vxr = vj[0] - vi[0], vyr = vj[1] - vi[1], vzr = vj[2] - vi[2];
            double renc12 = renci + rencj;
            int32_t lenc, lvd;
            swo_encounter_check_one(xr, yr, zr, vxr, vyr, vzr, renc12, dt, &lenc, &lvd);
            if (lenc) {
                if (in1)
                    keys_push(ii, j); /* index1 = list-1 body, index2 = list-2 body */
                else
                    keys_push(j, ii);
            }
        }
    }
    keys_sort_unique(); /* :899 */
}

/* encounter_check.f90:143-192 */
int64_t swo_encounter_sas_plpl(int32_t npl, const double *r, const double *v, const double *renc, double dt)
{
    keys_reset();
    g_nbox_total = 0;
    if (npl == 0) return 0;
    double *rmin = (double *)malloc(sizeof(double) * (size_t)npl), *rmax = (double *)malloc(sizeof(double) * (size_t)npl);
    int32_t *ind = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)npl);
    int64_t *ibeg = (int64_t *)malloc(sizeof(int64_t) * (size_t)npl), *iend = (int64_t *)malloc(sizeof(int64_t) * (size_t)npl);
    extents(npl, r, renc, rmin, rmax);
    sort_aabb_1d(npl, rmin, rmax, ind, ibeg, iend);
    sweep_single(npl, ind, ibeg, iend, r, v, renc, dt);
    free(rmin); free(rmax); free(ind); free(ibeg); free(iend);
    return g_nkeys;
}

static int64_t sas_double(int32_t n1, int32_t n2, const double *r1, const double *v1, const double *r2,
                          const double *v2, const double *renc1, const double *renc2, double dt)
{
    keys_reset();
    g_nbox_total = 0;
    if (n1 == 0 || n2 == 0) return 0;
    int32_t ntot = n1 + n2;
    double *rmin = (double *)malloc(sizeof(double) * (size_t)ntot), *rmax = (double *)malloc(sizeof(double) * (size_t)ntot);
    int32_t *ind = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)ntot);
    int64_t *ibeg = (int64_t *)malloc(sizeof(int64_t) * (size_t)ntot), *iend = (int64_t *)malloc(sizeof(int64_t) * (size_t)ntot);
    extents(n1, r1, renc1, rmin, rmax);
    extents(n2, r2, renc2, rmin + n1, rmax + n1);
    sort_aabb_1d(ntot, rmin, rmax, ind, ibeg, iend);
    sweep_double(n1, n2, ind, ibeg, iend, r1, v1, r2, v2, renc1, renc2, dt);
    free(rmin); free(rmax); free(ind); free(ibeg); free(iend);
    return g_nkeys;
}

/* encounter_check.f90:261-326 (test particles have renc = 0) */
int64_t swo_encounter_sas_pltp(int32_t npl, int32_t ntp, const double *rpl, const double *vpl, const double *rtp,
                               const double *vtp, const double *rencpl, double dt)
{
    return sas_double(npl, ntp, rpl, vpl, rtp, vtp, rencpl, NULL, dt);
}

/* encounter_check.f90:195-258 */
int64_t swo_encounter_sas_plplm(int32_t nplm, int32_t nplt, const double *rplm, const double *vplm,
                                const double *rplt, const double *vplt, const double *rencm, const double *renct,
                                double dt)
{
    return sas_double(nplm, nplt, rplm, vplm, rplt, vplt, rencm, renct, dt);
}

/* encounter_check.f90:42-109 with lencounter_sas_plpl true.  The reference re-sorts the merged list by index1
 * only (unstable); the canonical order used everywhere in this repo is lexicographic (index1,index2). */
int64_t swo_encounter_all_plplm(int32_t nplm, int32_t nplt, const double *rplm, const double *vplm,
                                const double *rplt, const double *vplt, const double *rencm, const double *renct,
                                double dt)
{
    int64_t n1 = swo_encounter_sas_plpl(nplm, rplm, vplm, rencm, dt);
    int64_t nbox1 = g_nbox_total;
    uint64_t *first = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n1 > 0 ? n1 : 1));
    memcpy(first, g_keys, sizeof(uint64_t) * (size_t)n1);
    int64_t n2 = swo_encounter_sas_plplm(nplm, nplt, rplm, vplm, rplt, vplt, rencm, renct, dt);
    for (int64_t k = 0; k < n2; ++k) g_keys[k] += (uint64_t)(uint32_t)nplm; /* :93 shift index2 */
    for (int64_t k = 0; k < n1; ++k) keys_push((int32_t)(first[k] >> 32), (int32_t)(first[k] & 0xffffffffu));
    free(first);
    g_nbox_total += nbox1;
    if (g_nkeys) qsort(g_keys, (size_t)g_nkeys, sizeof(uint64_t), cmp_u64);
    return g_nkeys;
}

/* encounter_check.f90:436-475 (+ :384-433) all pairs i<j */
int64_t swo_encounter_tri_plpl(int32_t npl, const double *r, const double *v, const double *renc, double dt)
{
    keys_reset();
    for (int32_t i = 1; i <= npl; ++i) {
        const double *ri = r + 3 * (size_t)(i - 1), *vi = v + 3 * (size_t)(i - 1);
        for (int32_t j = i + 1; j <= npl; ++j) {
            const double *rj = r + 3 * (size_t)(j - 1), *vj = v + 3 * (size_t)(j - 1);
            double xr = rj[0] - ri[0], yr = rj[1] - ri[1], zr = rj[2] - ri[2];
            double vxr = vj[0] - vi[0], vyr = vj[1] - vi[1], vzr = vj[2] - vi[2];
            double renc12 = renc[i - 1] + renc[j - 1];
            int32_t lenc, lvd;
            swo_encounter_check_one(xr, yr, zr, vxr, vyr, vzr, renc12, dt, &lenc, &lvd);
            if (lenc) keys_push(i, j);
        }
    }
    return g_nkeys;
}

/* encounter_check.f90:525-570 */
int64_t swo_encounter_tri_pltp(int32_t npl, int32_t ntp, const double *rpl, const double *vpl, const double *rtp,
                               const double *vtp, const double *rencpl, double dt)
{
    keys_reset();
    for (int32_t i = 1; i <= npl; ++i) {
        const double *ri = rpl + 3 * (size_t)(i - 1), *vi = vpl + 3 * (size_t)(i - 1);
        for (int32_t j = 1; j <= ntp; ++j) {
            const double *rj = rtp + 3 * (size_t)(j - 1), *vj = vtp + 3 * (size_t)(j - 1);
            double xr = rj[0] - ri[0], yr = rj[1] - ri[1], zr = rj[2] - ri[2];
            double vxr = vj[0] - vi[0], vyr = vj[1] - vi[1], vzr = vj[2] - vi[2];
            double renc12 = rencpl[i - 1] + 0.0;
            int32_t lenc, lvd;
            swo_encounter_check_one(xr, yr, zr, vxr, vyr, vzr, renc12, dt, &lenc, &lvd);
            if (lenc) keys_push(i, j);
        }
    }
    return g_nkeys;
}

/* encounter_check.f90:475-522: plm x plt double loop, index1 = plm index, index2 = plt index (caller shifts by nplm) */
int64_t swo_encounter_tri_plplm(int32_t nplm, int32_t nplt, const double *rplm, const double *vplm, const double *rplt,
                                const double *vplt, const double *rencm, const double *renct, double dt)
{
    keys_reset();
    for (int32_t i = 1; i <= nplm; ++i) {
        const double *ri = rplm + 3 * (size_t)(i - 1), *vi = vplm + 3 * (size_t)(i - 1);
        for (int32_t j = 1; j <= nplt; ++j) {
            const double *rj = rplt + 3 * (size_t)(j - 1), *vj = vplt + 3 * (size_t)(j - 1);
            double xr = rj[0] - ri[0], yr = rj[1] - ri[1], zr = rj[2] - ri[2];
            double vxr = vj[0] - vi[0], vyr = vj[1] - vi[1], vzr = vj[2] - vi[2];
            double renc12 = rencm[i - 1] + renct[j - 1];
            int32_t lenc, lvd;
            swo_encounter_check_one(xr, yr, zr, vxr, vyr, vzr, renc12, dt, &lenc, &lvd);
            if (lenc) keys_push(i, j);
        }
    }
    return g_nkeys;
}
