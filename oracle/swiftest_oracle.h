/*
 * swiftest_oracle.h -- CPU restatement of Swiftest's force-and-drift hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and there only as the checker / reported CPU baseline.  The CUDA product path never calls it.
 *
 * PARITY STATUS: the reference (Swiftest 2023.10.2, Modern Fortran) holds no function-level
 * golden vectors or known-answer tests for kick / sweep / drift, and no
 * Fortran compiler exists in the build container or on the GPU box (profiles/r02_fortran_probe.txt), so
 * the reference cannot be COMPILED here (there is no oracle/_ref).  Round 2 pins the restatement by
 * EXECUTING THE REFERENCE'S OWN FORTRAN STATEMENTS with a Fortran-subset interpreter
 * (oracle/f90interp.py; generator tests/golden/gen_golden_fortran.py reads /root/reference/src):
 *   - kick (flat/tri, rad/norad, every nplm branch, explicit pair lists, tp), sort-and-sweep and triangular
 *     encounter checks (plpl, pltp, plplm, merged list; util_sort quicksorts, dedupe, F3 quirk), drift (all
 *     solver branches, GR, failures): PINNED, BIT FOR BIT, tests/test_oracle_fortran_goldens.py.
 *   - whole helio and WHM steps, energy/momentum, SyMBA list kernels, collision/discard predicates
 *     (swiftest_oracle_step.c, swiftest_oracle_whm.c): PINNED the same way.
 *   - drift additionally agrees with the reference's own Python two-body propagation (swiftest/tool.py
 *     xv2el_one/el2xv_one, imported from /root/reference by tests/golden/gen_golden.py) at 1e-11.
 *   What the interpreter cannot show is compiler-specific code generation (FMA contraction, reassociation
 *   under the reference's -ffast-math release flags, OpenMP reduction order): the pinned semantics are
 *   "one IEEE operation per Fortran operation, serial loop order", the same contract as -ffp-contract=off here.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference/src).
 * Arrays use the Fortran memory layout: r(3,n) column-major == C r[3*i + {0,1,2}].
 * Indices crossing this interface are 1-based exactly as in the reference.
 * Compile with -ffp-contract=off: one IEEE-754 double operation per Fortran operation, left to right.
 */
#ifndef SWIFTEST_ORACLE_H
#define SWIFTEST_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- gravity: swiftest/swiftest_kick.f90 ---- */
void swo_kick_one_pl(double rji2, double xr, double yr, double zr, double Gmi, double Gmj,
                     double *axi, double *ayi, double *azi, double *axj, double *ayj, double *azj);
void swo_kick_one_tp(double rji2, double xr, double yr, double zr, double GMpl, double *ax, double *ay, double *az);
/* k_plpl == NULL means the canonical flattened upper-triangular order of swiftest_util_flatten_eucl_plpl */
void swo_kick_flat_rad_pl(int32_t npl, int64_t nplpl, const int32_t *k_plpl, const double *r, const double *Gmass,
                          const double *radius, double *acc);
void swo_kick_flat_norad_pl(int32_t npl, int64_t nplpl, const int32_t *k_plpl, const double *r, const double *Gmass,
                            double *acc);
void swo_kick_tri_rad_pl(int32_t npl, int32_t nplm, const double *r, const double *Gmass, const double *radius,
                         double *acc);
void swo_kick_tri_norad_pl(int32_t npl, int32_t nplm, const double *r, const double *Gmass, double *acc);
void swo_kick_all_tp(int32_t ntp, int32_t npl, const double *rtp, const double *rpl, const double *GMpl,
                     const int32_t *lmask, double *acc);
/* symba/symba_kick.f90:59-70 : ah -= flat_rad(encounter list) */
void swo_symba_kick_subtract_enc(int32_t npl, int64_t nenc, const int32_t *index1, const int32_t *index2,
                                 const double *rh, const double *Gmass, const double *radius, double *ah);
int64_t swo_nplplm(int64_t npl, int64_t nplm);
void swo_flatten_k_to_ij(int32_t n, int64_t k, int32_t *i, int32_t *j);
void swo_flatten_ij_to_k(int32_t n, int32_t i, int32_t j, int64_t *k);
/* per-component sum of |term| over the interactions the tri kernels evaluate: the scale of the 1e-12 tolerance */
void swo_kick_tri_abs_scale(int32_t npl, int32_t nplm, const double *r, const double *Gmass, const double *radius,
                            int lrad, double *scale);

/* reference-shaped OpenMP loops, used ONLY as the timed CPU baseline (bench.py) */
void swo_omp_kick_flat_rad_pl(int32_t npl, int32_t nplm, const double *r, const double *Gmass, const double *radius,
                              double *acc);
void swo_omp_kick_tri_rad_pl(int32_t npl, int32_t nplm, const double *r, const double *Gmass, const double *radius,
                             double *acc);
/* row-sampled variant: only rows [i0,i1) of the full-row branch (bounded CPU sample of a big system) */
void swo_omp_kick_tri_rad_pl_rows(int32_t npl, int32_t nplm, int32_t i0, int32_t i1, const double *r,
                                  const double *Gmass, const double *radius, double *acc);
void swo_omp_kick_all_tp(int32_t ntp, int32_t npl, const double *rtp, const double *rpl, const double *GMpl,
                         const int32_t *lmask, double *acc);
void swo_omp_drift_all(const double *mu, double *x, double *v, int32_t n, int lgr, double inv_c2, double dt,
                       const int32_t *lmask, int32_t *iflag);
int swo_omp_max_threads(void);
void swo_omp_set_num_threads(int n);
void swo_use_omp_kick(int on);
int swo_omp_kick_enabled(void);

/* ---- drift: swiftest/swiftest_drift.f90, swiftest/swiftest_orbel.f90:147-172 ---- */
void swo_drift_all(const double *mu, double *x, double *v, int32_t n, int lgr, double inv_c2, double dt,
                   const int32_t *lmask, int32_t *iflag);
void swo_drift_one(double mu, double *rx, double *ry, double *rz, double *vx, double *vy, double *vz, double dt,
                   int32_t *iflag);
void swo_drift_dan(double mu, double *rx0, double *ry0, double *rz0, double *vx0, double *vy0, double *vz0,
                   double dt0, int32_t *iflag);
void swo_drift_kepmd(double dm, double es, double ec, double *x, double *s, double *c);
void swo_drift_kepu(double dt, double r0, double mu, double alpha, double u, double *fp, double *c1, double *c2,
                    double *c3, int32_t *iflag);
void swo_drift_kepu_stumpff(double *x, double *c0, double *c1, double *c2, double *c3);
void swo_orbel_scget(double angle, double *sx, double *cx);
/* which branch drift_dan takes for a body: 0 kepmd fast path, 1 kepu elliptic small-dt guess,
 * 2 kepu elliptic Danby guess (uses sin), 3 kepu hyperbolic (uses x**(1/3)) */
int32_t swo_drift_branch(double mu, double rx, double ry, double rz, double vx, double vy, double vz, double dt);

/* ---- encounters: encounter/encounter_check.f90 ---- */
void swo_encounter_check_one(double xr, double yr, double zr, double vxr, double vyr, double vzr, double renc,
                             double dt, int32_t *lencounter, int32_t *lvdotr);
void swo_symba_set_renc(int32_t npl, const double *rhill, int32_t irec, double *renc);
/* Results are returned in library-owned buffers, canonical (index1,index2) lexicographic order, 1-based.
 * Call swo_encounter_fetch afterwards to copy them out (mirrors the Fortran allocatable intent(out)). */
int64_t swo_encounter_sas_plpl(int32_t npl, const double *r, const double *v, const double *renc, double dt);
int64_t swo_encounter_sas_pltp(int32_t npl, int32_t ntp, const double *rpl, const double *vpl, const double *rtp,
                               const double *vtp, const double *rencpl, double dt);
int64_t swo_encounter_sas_plplm(int32_t nplm, int32_t nplt, const double *rplm, const double *vplm,
                                const double *rplt, const double *vplt, const double *rencm, const double *renct,
                                double dt);
/* encounter_check_all_plplm :42-109 with sort-and-sweep: plpl on the plm block + plm x plt, index2 shifted by nplm */
int64_t swo_encounter_all_plplm(int32_t nplm, int32_t nplt, const double *rplm, const double *vplm,
                                const double *rplt, const double *vplt, const double *rencm, const double *renct,
                                double dt);
/* triangular (all-pairs) variants :436-570 -- also the brute-force superset checker for the sweep */
int64_t swo_encounter_tri_plpl(int32_t npl, const double *r, const double *v, const double *renc, double dt);
int64_t swo_encounter_tri_pltp(int32_t npl, int32_t ntp, const double *rpl, const double *vpl, const double *rtp,
                               const double *vtp, const double *rencpl, double dt);
int64_t swo_encounter_tri_plplm(int32_t nplm, int32_t nplt, const double *rplm, const double *vplm, const double *rplt,
                                const double *vplt, const double *rencm, const double *renct, double dt);
void swo_encounter_fetch(int32_t *index1, int32_t *index2, int32_t *lvdotr);
/* statistics of the last sort-and-sweep call: sum_i nbox_i over loverlap bodies (broad-phase candidates) */
int64_t swo_encounter_last_nbox_total(void);

/* ---- the O(N) glue around the hot path and the energy sums (swiftest_oracle_step.c) ---- */
void swo_coord_vh2vb_pl(int32_t npl, double GMcb, const double *Gmass, const double *vh, double *vb, double *vbcb);
void swo_coord_vb2vh_pl(int32_t npl, double GMcb, const double *Gmass, const int32_t *lactive, const double *vb,
                        double *vh, double *vbcb);
void swo_coord_vh2vb_tp(int32_t ntp, const int32_t *lmask, const double *vbcb, const double *vh, double *vb);
void swo_coord_vb2vh_tp(int32_t ntp, const int32_t *lmask, const double *vbcb, const double *vb, double *vh);
void swo_coord_h2b_pl(int32_t npl, double GMcb, const double *Gmass, const int32_t *lactive, const double *rh,
                      const double *vh, double *rb, double *vb, double *rbcb, double *vbcb);
void swo_helio_drift_linear_pl(int32_t npl, double GMcb, const double *Gmass, const double *vb, const int32_t *lmask,
                               double dt, double *rh, double *pt);
void swo_helio_drift_linear_tp(int32_t ntp, const int32_t *lmask, const double *pt, double dt, double *rh);
void swo_helio_kick_vb(int32_t n, const int32_t *lmask, const double *ah, double dt, double *vb);
void swo_helio_step_pl(int32_t npl, double GMcb, const double *Gmass, const double *radius, int lflat,
                       const int32_t *lmask, int32_t *lfirst, double dt, double *rh, double *vh, double *vb,
                       double *ah, double *rbeg, double *rend, double *ptbeg, double *ptend, double *vbcb,
                       int32_t *iflag);
void swo_helio_step_tp(int32_t ntp, int32_t npl, double GMcb, const double *GMpl, const double *rbeg,
                       const double *rend, const double *ptbeg, const double *ptend, const int32_t *lmask,
                       int32_t *lfirst, double dt, double *rh, double *vh, double *vb, double *ah, int32_t *iflag);
void swo_discard_pl_close(const double *dx, const double *dv, double dt, double r2crit, int32_t *iflag, double *r2min);
int32_t swo_discard_pl_tp(int32_t ntp, int32_t npl, const double *rtp, const double *vtp, const int32_t *lactive,
                          const double *rpl, const double *vpl, const double *radius, double dt, int32_t *iplanet);
int64_t swo_symba_encounter_check_list(int64_t nenc, const int32_t *index1, const int32_t *index2,
                                       const int32_t *lencmask, const double *r1, const double *v1, const double *renc1,
                                       const double *radius1, const double *r2, const double *v2, const double *renc2,
                                       const double *radius2, double dt, int32_t *lencounter, int32_t *lvdotr);
double swo_pow_r8_i4(double a, int32_t b);
void swo_symba_kick_list_plpl(int64_t nenc, const int32_t *index1, const int32_t *index2, const int32_t *lactive,
                              int32_t npl, const int32_t *levelg, const double *rh, const double *rhill,
                              const double *Gmass, double dt, int32_t irec, int32_t sgn, double *vb, double *ah,
                              int32_t *lgood_out);
void swo_symba_kick_list_pltp(int64_t nenc, const int32_t *index1, const int32_t *index2, const int32_t *lactive,
                              int32_t npl, int32_t ntp, const int32_t *levelg_pl, const int32_t *levelg_tp,
                              const double *rh_pl, const double *rhill, const double *Gmass, const double *rh_tp,
                              double dt, int32_t irec, int32_t sgn, double *vb_tp, double *ah_tp, int32_t *lgood_out);
void swo_orbel_xv2aeq(double mu, double rx, double ry, double rz, double vx, double vy, double vz, double *a, double *e,
                      double *q);
void swo_collision_check_one(double xr, double yr, double zr, double vxr, double vyr, double vzr, double Gmtot,
                             double rlim, double dt, int32_t lvdotr, int32_t *lcollision, int32_t *lclosest);
int64_t swo_collision_check_list(int64_t nenc, const int32_t *index1, const int32_t *index2, const int32_t *lmask,
                                 const int32_t *lvdotr, const double *r1, const double *v1, const double *Gmass1,
                                 const double *radius1, const double *r2, const double *v2, const double *Gmass2,
                                 const double *radius2, double dt, int32_t *lcollision, int32_t *lclosest);
void swo_whm_kick_getacch_ah0(int32_t n, const double *mu, const double *rhp, double *ah0);
void swo_get_potential_energy_tri(int32_t npl, const int32_t *lmask, double GMcb, const double *Gmass,
                                  const double *mass, const double *rb, double *pe);
void swo_get_potential_energy_flat(int32_t npl, int64_t nplpl, const int32_t *k_plpl, const int32_t *lmask,
                                   double GMcb, const double *Gmass, const double *mass, const double *rb, double *pe);
/* out[0] ke_orbit, out[1] pe, out[2] be, out[3] te, out[4..6] L_orbit, out[7] GMtot */
void swo_get_energy_and_momentum(int32_t npl, const int32_t *lmask, double GMcb, double mass_cb, const double *rbcb,
                                 const double *vbcb, const double *Gmass, const double *mass, const double *radius,
                                 const double *rb, const double *vb, int lclose, int lflat, double *out);

/* ---- Wisdom-Holman step around the hot path (swiftest_oracle_whm.c) ---- */
void swo_whm_set_mu_eta(int32_t npl, double GMcb, const double *Gmass, double *mu, double *eta, double *muj);
void swo_whm_coord_h2j(int32_t npl, const double *Gmass, const double *eta, const double *rh, const double *vh,
                       double *xj, double *vj);
void swo_whm_coord_j2h(int32_t npl, const double *Gmass, const double *eta, const double *xj, const double *vj,
                       double *rh, double *vh);
void swo_whm_coord_vh2vj(int32_t npl, const double *Gmass, const double *eta, const double *vh, double *vj);
void swo_whm_kick_getacch_pl(int32_t npl, double GMcb, const double *Gmass, const double *radius, int lflat,
                             const int32_t *lmask, const double *rh, const double *xj, double *ah);
void swo_whm_step_pl(int32_t npl, double GMcb, const double *Gmass, const double *radius, int lflat,
                     const int32_t *lmask, const double *eta, const double *muj, int32_t *lfirst, double dt, double *rh,
                     double *vh, double *xj, double *vj, double *ah, double *rbeg, double *rend, int32_t *iflag);
void swo_whm_step_tp(int32_t ntp, int32_t npl, double GMcb, const double *GMpl, const double *rbeg, const double *rend,
                     const int32_t *lmask, int32_t *lfirst, double dt, double *rh, double *vh, double *ah, int32_t *iflag);

#ifdef __cplusplus
}
#endif
#endif
