// whm_kernels.cu -- the Wisdom-Holman planet step on the device-resident pl population (SURVEY.md section 8f rank 1,
// VERDICT r1 "missing" #6): Jacobi coordinate changes, the ah0 / ah1 / ah2 terms, kick - drift - kick.
//
// Reference (paths relative to src/):
//   whm_util_set_mu_eta_pl     whm/whm_util.f90:175-198      eta = running mass, muj = GMcb*eta(i)/eta(i-1)
//   whm_coord_h2j_pl           whm/whm_coord.f90:14-46       xj(i) = rh(i) - sum_{k<i} Gm_k rh_k / eta(i-1)
//   whm_coord_j2h_pl           whm/whm_coord.f90:49-80       rh(i) = xj(i) + sum_{k<i} Gm_k xj_k / eta(k)
//   whm_coord_vh2vj_pl         whm/whm_coord.f90:83-113
//   whm_kick_getacch_pl        whm/whm_kick.f90:14-67        ah = ah0 + ah1 + ah2 + pl%accel_int
//   whm_kick_getacch_ah0/1/2   whm/whm_kick.f90:124-205      ah2 carries a running sum along the Jacobi chain
//   whm_kick_vh_pl             whm/whm_kick.f90:208-262      first step: h2j + accelerations at the begin positions
//   whm_drift_pl               whm/whm_drift.f90:14-58       Danby drift of (xj, vj) with mu = muj
//   whm_step_pl                whm/whm_step.f90:37-69
//
// The chains (eta, h2j, j2h, vh2vj, ah0, ah2) are SERIAL in the reference: body i needs the running sum over the bodies
// before it.  They are restated as single-thread loops (one kernel each, loads software-pipelined by the compiler):
// identical summation order, so with --fmad=false (this file) every chain is bit-identical to the CPU restatement.  A WHM
// run has a handful to a few hundred massive bodies, so these loops cost microseconds; the O(N^2) part of the step is
// pl%accel_int (kick_kernels.cu / kick_flat_kernels.cu) and the per-body part the Kepler drift (drift_kernels.cu).
#include "swcu_internal.cuh"

namespace swcu {
namespace {

struct V3 {
    double *x, *y, *z;
};
struct CV3 {
    const double *x, *y, *z;
};

__global__ void whm_set_mu_eta_kernel(int n, double gmcb, const double *__restrict__ gm, double *__restrict__ eta,
                                      double *__restrict__ muj)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double e = gmcb + gm[0];
    eta[0] = e;
    muj[0] = e;
    for (int i = 1; i < n; ++i) {
        const double en = e + gm[i];
        eta[i] = en;
        muj[i] = gmcb * en / e;
        e = en;
    }
}

// mode 0: h2j (positions and velocities), 1: vh2vj (velocities only)
__global__ void whm_h2j_kernel(int n, int mode, const double *__restrict__ gm, const double *__restrict__ eta, CV3 rh, CV3 vh,
                               V3 xj, V3 vj)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double sx0 = 0.0, sx1 = 0.0, sx2 = 0.0, sv0 = 0.0, sv1 = 0.0, sv2 = 0.0;
    if (mode == 0) {
        xj.x[0] = rh.x[0];
        xj.y[0] = rh.y[0];
        xj.z[0] = rh.z[0];
    }
    vj.x[0] = vh.x[0];
    vj.y[0] = vh.y[0];
    vj.z[0] = vh.z[0];
    for (int i = 1; i < n; ++i) {
        const double g = gm[i - 1], e = eta[i - 1];
        if (mode == 0) {
            sx0 = sx0 + g * rh.x[i - 1];
            sx1 = sx1 + g * rh.y[i - 1];
            sx2 = sx2 + g * rh.z[i - 1];
            xj.x[i] = rh.x[i] - sx0 / e;
            xj.y[i] = rh.y[i] - sx1 / e;
            xj.z[i] = rh.z[i] - sx2 / e;
        }
        sv0 = sv0 + g * vh.x[i - 1];
        sv1 = sv1 + g * vh.y[i - 1];
        sv2 = sv2 + g * vh.z[i - 1];
        vj.x[i] = vh.x[i] - sv0 / e;
        vj.y[i] = vh.y[i] - sv1 / e;
        vj.z[i] = vh.z[i] - sv2 / e;
    }
}

__global__ void whm_j2h_kernel(int n, const double *__restrict__ gm, const double *__restrict__ eta, CV3 xj, CV3 vj, V3 rh,
                               V3 vh)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double sx0 = 0.0, sx1 = 0.0, sx2 = 0.0, sv0 = 0.0, sv1 = 0.0, sv2 = 0.0;
    rh.x[0] = xj.x[0];
    rh.y[0] = xj.y[0];
    rh.z[0] = xj.z[0];
    vh.x[0] = vj.x[0];
    vh.y[0] = vj.y[0];
    vh.z[0] = vj.z[0];
    for (int i = 1; i < n; ++i) {
        const double g = gm[i - 1], e = eta[i - 1];
        sx0 = sx0 + g * xj.x[i - 1] / e;
        sx1 = sx1 + g * xj.y[i - 1] / e;
        sx2 = sx2 + g * xj.z[i - 1] / e;
        sv0 = sv0 + g * vj.x[i - 1] / e;
        sv1 = sv1 + g * vj.y[i - 1] / e;
        sv2 = sv2 + g * vj.z[i - 1] / e;
        rh.x[i] = xj.x[i] + sx0;
        rh.y[i] = xj.y[i] + sx1;
        rh.z[i] = xj.z[i] + sx2;
        vh.x[i] = vj.x[i] + sv0;
        vh.y[i] = vj.y[i] + sv1;
        vh.z[i] = vj.z[i] + sv2;
    }
}

// whm_kick_getacch_ah0 (whm_kick.f90:124-149) over bodies [first, n): out = -sum Gm_i r_i / |r_i|^3 (serial order)
__global__ void whm_ah0_kernel(int first, int n, const double *__restrict__ gm, CV3 r, double *__restrict__ out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int i = first; i < n; ++i) {
        const double x = r.x[i], y = r.y[i], z = r.z[i];
        const double r2 = x * x + y * y + z * z;
        const double ir3h = 1.0 / (r2 * sqrt(r2));
        const double fac = gm[i] * ir3h;
        a0 = a0 - fac * x;
        a1 = a1 - fac * y;
        a2 = a2 - fac * z;
    }
    out[0] = a0;
    out[1] = a1;
    out[2] = a2;
}

// ah(i) = ((ah(i) + ah0) + GMcb*(xj*ir3j - rh*ir3h)) for i >= 1 under the mask (ah0 for every body), whm_kick.f90:33-35,150-172
// and the per-body factor of the ah2 chain, fac(i) = Gm(i)*GMcb*ir3j(i)/etaj is formed in the chain kernel below
__global__ void whm_ah01_kernel(int n, double gmcb, const int32_t *__restrict__ lmask, CV3 rh, CV3 xj,
                                const double *__restrict__ ah0, double *__restrict__ ir3j_out, V3 ah)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a0 = ah.x[i] + ah0[0], a1 = ah.y[i] + ah0[1], a2 = ah.z[i] + ah0[2];
    // whm_util_set_ir3j (whm_util.f90:117-140): ir = 1/sqrt(r2) ; ir3 = ir/r2
    const double hx = rh.x[i], hy = rh.y[i], hz = rh.z[i];
    double r2 = hx * hx + hy * hy + hz * hz;
    double ir = 1.0 / sqrt(r2);
    const double ir3h = ir / r2;
    const double jx = xj.x[i], jy = xj.y[i], jz = xj.z[i];
    r2 = jx * jx + jy * jy + jz * jz;
    ir = 1.0 / sqrt(r2);
    const double ir3j = ir / r2;
    ir3j_out[i] = ir3j;
    if (i >= 1 && lmask[i] != 0) {
        a0 = a0 + gmcb * (jx * ir3j - hx * ir3h);
        a1 = a1 + gmcb * (jy * ir3j - hy * ir3h);
        a2 = a2 + gmcb * (jz * ir3j - hz * ir3h);
    }
    ah.x[i] = a0;
    ah.y[i] = a1;
    ah.z[i] = a2;
}

// whm_kick_getacch_ah2 (whm_kick.f90:175-205): ah2(i) = ah2(i-1) + Gm(i)*GMcb*ir3j(i)/etaj * xj(i), etaj running over
// the masked bodies
__global__ void whm_ah2_kernel(int n, double gmcb, const int32_t *__restrict__ lmask, const double *__restrict__ gm,
                               const double *__restrict__ ir3j, CV3 xj, V3 ah)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double o0 = 0.0, o1 = 0.0, o2 = 0.0, etaj = gmcb;
    for (int i = 1; i < n; ++i) {
        if (lmask[i] == 0) continue;
        etaj = etaj + gm[i - 1];
        const double fac = gm[i] * gmcb * ir3j[i] / etaj;
        o0 = o0 + fac * xj.x[i];
        o1 = o1 + fac * xj.y[i];
        o2 = o2 + fac * xj.z[i];
        ah.x[i] = ah.x[i] + o0;
        ah.y[i] = ah.y[i] + o1;
        ah.z[i] = ah.z[i] + o2;
    }
}

// vh = vh + ah*dt under the mask (whm_kick.f90:255-259): one multiply, one add (this file is compiled --fmad=false)
__global__ void whm_kick_vh_kernel(int n, const int32_t *__restrict__ lmask, double dt, CV3 ah, V3 vh)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || lmask[i] == 0) return;
    vh.x[i] = vh.x[i] + ah.x[i] * dt;
    vh.y[i] = vh.y[i] + ah.y[i] * dt;
    vh.z[i] = vh.z[i] + ah.z[i] * dt;
}

// a(i) = a(i) + c under the mask (whm_kick.f90:98-101)
__global__ void whm_add_const_kernel(int n, const int32_t *__restrict__ lmask, const double *__restrict__ c, V3 a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || lmask[i] == 0) return;
    a.x[i] = a.x[i] + c[0];
    a.y[i] = a.y[i] + c[1];
    a.z[i] = a.z[i] + c[2];
}

V3 v3(DevBuf &x, DevBuf &y, DevBuf &z) { return V3{x.as<double>(), y.as<double>(), z.as<double>()}; }
CV3 cv3(const DevBuf &x, const DevBuf &y, const DevBuf &z) { return CV3{x.as<double>(), y.as<double>(), z.as<double>()}; }

int ensure_whm(swcu_context *ctx, Body &pl, double gmcb)
{
    auto &W = ctx->whm;
    const size_t nb = sizeof(double) * (size_t)(pl.n > 0 ? pl.n : 1);
    DevBuf *all[] = {&W.xjx, &W.xjy, &W.xjz, &W.vjx, &W.vjy, &W.vjz, &W.eta, &W.muj, &W.ir3j};
    for (DevBuf *d : all) SWCU_CUDA(ctx, d->ensure(nb));
    SWCU_TRY(ensure_step_state(ctx));
    SWCU_TRY(ensure_helio(ctx, pl));  // rbeg / rend copies live in the helio buffers (b*, e*)
    if (W.generation != pl.generation || W.n != pl.n || W.gmcb != gmcb) {
        whm_set_mu_eta_kernel<<<1, 32, 0, ctx->stream>>>(pl.n, gmcb, pl.Gm.as<double>(), W.eta.as<double>(), W.muj.as<double>());
        SWCU_KERNEL_CHECK(ctx);
        W.generation = pl.generation;
        W.n = pl.n;
        W.gmcb = gmcb;
    }
    return SWCU_OK;
}

// whm_kick_getacch_pl without oblateness, GR or user force: ah = 0 + ah0 + ah1 + ah2 + interaction term
int whm_getacch_pl(swcu_context *ctx, Body &pl, double gmcb, int variant, int lclose)
{
    auto &W = ctx->whm;
    const int n = pl.n;
    SWCU_TRY(fill3_f64(ctx, pl.ax.as<double>(), pl.ay.as<double>(), pl.az.as<double>(), 0.0, n));
    double *ah0 = ctx->cbs.as<double>() + CBS_AH0PL;
    whm_ah0_kernel<<<1, 32, 0, ctx->stream>>>(1, n, pl.Gm.as<double>(), cv3(pl.rx, pl.ry, pl.rz), ah0);  // bodies 2..npl (:33)
    SWCU_KERNEL_CHECK(ctx);
    whm_ah01_kernel<<<cdiv(n, 128), 128, 0, ctx->stream>>>(n, gmcb, pl.lmask.as<int32_t>(), cv3(pl.rx, pl.ry, pl.rz),
                                                          cv3(W.xjx, W.xjy, W.xjz), ah0, W.ir3j.as<double>(),
                                                          v3(pl.ax, pl.ay, pl.az));
    SWCU_KERNEL_CHECK(ctx);
    whm_ah2_kernel<<<1, 32, 0, ctx->stream>>>(n, gmcb, pl.lmask.as<int32_t>(), pl.Gm.as<double>(), W.ir3j.as<double>(),
                                              cv3(W.xjx, W.xjy, W.xjz), v3(pl.ax, pl.ay, pl.az));
    SWCU_KERNEL_CHECK(ctx);
    return pl_accel_int(ctx, variant, lclose);
}

}  // namespace

// whm_step_pl (whm_step.f90:37-69) on the resident pl population: rh, vh in r*, v*; xj, vj and the accelerations are kept
// on the device between steps; pl%rbeg / pl%rend are left in b* / e* and whm_kick_getacch_ah0 of ALL planets at rend in
// cbs[CBS_AH0TP] for the test-particle step that follows (whm_kick.f90:91-93)
int whm_step_pl(swcu_context *ctx, double gmcb, double dt, int variant, int lclose, int lfirst, int32_t *nfail)
{
    Body &pl = ctx->pl;
    auto &W = ctx->whm;
    if (nfail) *nfail = 0;
    if (pl.n == 0) return SWCU_OK;
    const int n = pl.n;
    const size_t nb = sizeof(double) * (size_t)n;
    const double dth = 0.5 * dt;
    SWCU_TRY(ensure_whm(ctx, pl, gmcb));
    // small systems: the whole step is ONE launch (drift_kernels.cu::whm_step_pl_small_kernel), bit-identical to the
    // reference's statement order for both loop variants; SWCU_WHM_FUSED=0 keeps the multi-launch form below
    static const bool fused_ok = !(getenv("SWCU_WHM_FUSED") && atoi(getenv("SWCU_WHM_FUSED")) == 0);
    if (fused_ok && n <= whm_small_max() && pl.nplm == n && pl.slice0 == 0 && pl.slice1 == n && ctx->tune_variant < 0) {
        const int flat = variant == SWCU_LOOP_FLAT || (variant == SWCU_LOOP_AUTO && n >= 128);
        SWCU_TRY(whm_step_pl_small(ctx, pl, gmcb, dt, flat, lclose, lfirst));
        W.ah0tp_valid = true;
        if (nfail) {
            SWCU_CUDA(ctx, cudaMemcpyAsync(nfail, ctx->scratch64.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        return SWCU_OK;
    }
    CV3 rh = cv3(pl.rx, pl.ry, pl.rz), vh = cv3(pl.vx, pl.vy, pl.vz);
    V3 xj = v3(W.xjx, W.xjy, W.xjz), vj = v3(W.vjx, W.vjy, W.vjz);
    if (lfirst) {  // whm_kick_vh_pl :236-243
        whm_h2j_kernel<<<1, 32, 0, ctx->stream>>>(n, 0, pl.Gm.as<double>(), W.eta.as<double>(), rh, vh, xj, vj);
        SWCU_KERNEL_CHECK(ctx);
        SWCU_TRY(whm_getacch_pl(ctx, pl, gmcb, variant, lclose));
    }
    DevBuf *rb[] = {&pl.bx, &pl.by, &pl.bz}, *re[] = {&pl.ex, &pl.ey, &pl.ez}, *rr[] = {&pl.rx, &pl.ry, &pl.rz};
    for (int k = 0; k < 3; ++k) SWCU_CUDA(ctx, cudaMemcpyAsync(rb[k]->p, rr[k]->p, nb, cudaMemcpyDeviceToDevice, ctx->stream));
    whm_kick_vh_kernel<<<cdiv(n, 128), 128, 0, ctx->stream>>>(n, pl.lmask.as<int32_t>(), dth, cv3(pl.ax, pl.ay, pl.az),
                                                            v3(pl.vx, pl.vy, pl.vz));  // vh += ah*dth
    SWCU_KERNEL_CHECK(ctx);
    whm_h2j_kernel<<<1, 32, 0, ctx->stream>>>(n, 1, pl.Gm.as<double>(), W.eta.as<double>(), rh, vh, xj, vj);  // vh2vj
    SWCU_KERNEL_CHECK(ctx);
    SWCU_TRY(drift_arrays(ctx, n, W.muj.as<double>(), W.xjx.as<double>(), W.xjy.as<double>(), W.xjz.as<double>(),
                          W.vjx.as<double>(), W.vjy.as<double>(), W.vjz.as<double>(), pl.lmask.as<int32_t>(),
                          pl.iflag.as<int32_t>(), dt));
    whm_j2h_kernel<<<1, 32, 0, ctx->stream>>>(n, pl.Gm.as<double>(), W.eta.as<double>(), cv3(W.xjx, W.xjy, W.xjz),
                                              cv3(W.vjx, W.vjy, W.vjz), v3(pl.rx, pl.ry, pl.rz), v3(pl.vx, pl.vy, pl.vz));
    SWCU_KERNEL_CHECK(ctx);
    SWCU_TRY(whm_getacch_pl(ctx, pl, gmcb, variant, lclose));  // kick(end)
    for (int k = 0; k < 3; ++k) SWCU_CUDA(ctx, cudaMemcpyAsync(re[k]->p, rr[k]->p, nb, cudaMemcpyDeviceToDevice, ctx->stream));
    whm_kick_vh_kernel<<<cdiv(n, 128), 128, 0, ctx->stream>>>(n, pl.lmask.as<int32_t>(), dth, cv3(pl.ax, pl.ay, pl.az),
                                                            v3(pl.vx, pl.vy, pl.vz));
    SWCU_KERNEL_CHECK(ctx);
    // ah0 of the test-particle kick at the end-of-step planets: all npl planets
    whm_ah0_kernel<<<1, 32, 0, ctx->stream>>>(0, n, pl.Gm.as<double>(), cv3(pl.rx, pl.ry, pl.rz),
                                              ctx->cbs.as<double>() + CBS_AH0TP);
    SWCU_KERNEL_CHECK(ctx);
    W.ah0tp_valid = true;
    if (nfail) {  // the drift kernel counted its failures in scratch64[0]
        SWCU_CUDA(ctx, cudaMemcpyAsync(nfail, ctx->scratch64.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return SWCU_OK;
}

// first test-particle step of a WHM run (whm_kick_vh_tp :288-295): ah = 0 + ah0(planets now) + direct terms, planets at
// their CURRENT resident positions -- call it before the planets' first step moves them
int whm_tp_first_accel(swcu_context *ctx)
{
    Body &tp = ctx->tp, &pl = ctx->pl;
    if (tp.n == 0 || pl.n == 0) return SWCU_OK;
    SWCU_TRY(ensure_step_state(ctx));
    double *ah0 = ctx->cbs.as<double>() + CBS_AH0TP;
    whm_ah0_kernel<<<1, 32, 0, ctx->stream>>>(0, pl.n, pl.Gm.as<double>(), cv3(pl.rx, pl.ry, pl.rz), ah0);
    SWCU_KERNEL_CHECK(ctx);
    SWCU_TRY(fill3_f64(ctx, tp.ax.as<double>(), tp.ay.as<double>(), tp.az.as<double>(), 0.0, tp.n));
    whm_add_const_kernel<<<cdiv(tp.n, 256), 256, 0, ctx->stream>>>(tp.n, tp.lmask.as<int32_t>(), ah0, v3(tp.ax, tp.ay, tp.az));
    SWCU_KERNEL_CHECK(ctx);
    KickProblem k;
    k.xi = tp.rx.as<double>(); k.yi = tp.ry.as<double>(); k.zi = tp.rz.as<double>(); k.radi = nullptr;
    k.row0 = 0; k.row1 = tp.n;
    k.xj = pl.rx.as<double>(); k.yj = pl.ry.as<double>(); k.zj = pl.rz.as<double>(); k.gmj = pl.Gm.as<double>(); k.radj = nullptr;
    k.col0 = 0; k.col1 = pl.n;
    k.diag = false;
    k.lmask = tp.lmask.as<int32_t>();
    k.ax = tp.ax.as<double>(); k.ay = tp.ay.as<double>(); k.az = tp.az.as<double>();
    return kick_rows(ctx, k, FAM_PLTP);
}

int whm_get_jacobi(swcu_context *ctx, double *xj, double *vj)
{
    Body &pl = ctx->pl;
    auto &W = ctx->whm;
    if (!W.xjx.p || W.n != pl.n) return fail(ctx, SWCU_ERR_STATE, "whm_get_jacobi: no WHM step has run on this population");
    if (xj) SWCU_TRY(download_vec3(ctx, xj, pl.n, 0, W.xjx, W.xjy, W.xjz));
    if (vj) SWCU_TRY(download_vec3(ctx, vj, pl.n, 1, W.vjx, W.vjy, W.vjz));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SWCU_OK;
}

}  // namespace swcu
