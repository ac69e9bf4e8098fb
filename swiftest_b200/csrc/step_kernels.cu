// step_kernels.cu -- the O(N) glue of the democratic-heliocentric step on the device-resident populations
// (SURVEY.md section 8f rank 1): vh2vb / vb2vh, the linear drift with its sum of Gm*vb, the velocity kick.
//
// Reference (paths relative to src/):
//   swiftest_util_coord_vh2vb_pl   swiftest/swiftest_util.f90:424-459     vbcb = -sum(Gm*vh)/GMtot ; vb = vh + vbcb
//   swiftest_util_coord_vb2vh_pl   swiftest/swiftest_util.f90:363-395     vbcb = -sum_{i=npl..1}(Gm*vb/GMcb) ; vh = vb - vbcb
//   swiftest_util_coord_v*2v*_tp   swiftest/swiftest_util.f90:398-421,462-485
//   helio_drift_linear_pl / _tp    helio/helio_drift.f90:129-200          pt = sum(Gm*vb, lmask)/GMcb ; rh += pt*dt
//   helio_kick_vb_pl / _tp         helio/helio_kick.f90:91-169            ah = 0 ; accel ; set_beg_end ; vb += ah*dt
//   helio_step_pl                  helio/helio_step.f90:37-78
//
// Compiled with --fmad=false: every element-wise update is one multiply and one add as in the reference, so the
// glue is bit-identical to the CPU restatement whenever the sums are (n <= SERIAL_SUM_MAX, see reduce.cuh).
// The central-body scalars (vbcb, ptbeg, ptend) never leave the device during a step: the reduction kernel finishes
// them and the element-wise kernel that follows reads them from ctx->cbs.
#include "reduce.cuh"
#include "swcu_internal.cuh"

namespace swcu {
namespace {

struct GmvTerm {  // terms Gm*v (optionally each divided by `div`), and Gm itself as the fourth component
    const double *gm, *vx, *vy, *vz;
    const int32_t *lmask;
    double div;
    bool use_div;
    __device__ bool operator()(int i, double *t) const
    {
        if (lmask && lmask[i] == 0) return false;
        const double g = gm[i];
        double a = g * vx[i], b = g * vy[i], c = g * vz[i];
        if (use_div) {
            a = a / div;
            b = b / div;
            c = c / div;
        }
        t[0] = a;
        t[1] = b;
        t[2] = c;
        t[3] = g;
        return true;
    }
};

enum FinOp { FIN_VH2VB = 0, FIN_VB2VH = 1, FIN_PT = 2 };

struct CbFin {
    double *cbs;
    double gmcb;
    int op;
    int slot;  // FIN_PT: CBS_PTBEG or CBS_PTEND
    __device__ void operator()(const double *s) const
    {
        for (int k = 0; k < 4; ++k) cbs[CBS_SUM + k] = s[k];
        if (op == FIN_VH2VB) {
            const double gmtot = gmcb + s[3];
            cbs[CBS_GMTOT] = gmtot;
            for (int k = 0; k < 3; ++k) cbs[CBS_VBCB + k] = (0.0 - s[k]) / gmtot;
        } else if (op == FIN_VB2VH) {
            for (int k = 0; k < 3; ++k) cbs[CBS_VBCB + k] = 0.0 - s[k];
        } else {
            for (int k = 0; k < 3; ++k) cbs[slot + k] = s[k] / gmcb;
        }
    }
};

// out = in + c*scale for the bodies of the mask (c: three doubles in device memory)
__global__ void add_vec3_kernel(int n, const int32_t *__restrict__ lmask, const double *__restrict__ c3, double scale,
                                const double *ix, const double *iy, const double *iz, double *ox, double *oy, double *oz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (lmask && lmask[i] == 0) return;
    const double c0 = c3[0] * scale, c1 = c3[1] * scale, c2 = c3[2] * scale;
    ox[i] = ix[i] + c0;
    oy[i] = iy[i] + c1;
    oz[i] = iz[i] + c2;
}

// vb += ah*dt for the bodies of the mask, and (optionally) keep a copy of the positions the kick was evaluated at
__global__ void kick_vb_save_kernel(int n, const int32_t *__restrict__ lmask, double dt, const double *__restrict__ ax,
                                    const double *__restrict__ ay, const double *__restrict__ az, double *__restrict__ wx,
                                    double *__restrict__ wy, double *__restrict__ wz, const double *__restrict__ rx,
                                    const double *__restrict__ ry, const double *__restrict__ rz, double *__restrict__ sx,
                                    double *__restrict__ sy, double *__restrict__ sz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (sx) {
        sx[i] = rx[i];
        sy[i] = ry[i];
        sz[i] = rz[i];
    }
    if (lmask[i] == 0) return;
    wx[i] = wx[i] + ax[i] * dt;
    wy[i] = wy[i] + ay[i] * dt;
    wz[i] = wz[i] + az[i] * dt;
}

// masked: 0 none, 1 lmask, 2 the active flags (status /= INACTIVE; all active unless swcu_body_set_active loaded them)
int reduce_gmv(swcu_context *ctx, const Body &b, const double *vx, const double *vy, const double *vz, int masked,
               bool use_div, bool reverse, double gmcb, int op, int slot)
{
    SWCU_TRY(ensure_step_state(ctx));
    const int32_t *mask = masked == 1 ? b.lmask.as<int32_t>() : (masked == 2 && b.has_active ? b.lactive.as<int32_t>() : nullptr);
    GmvTerm term{b.Gm.as<double>(), vx, vy, vz, mask, gmcb, use_div};
    CbFin fin{ctx->cbs.as<double>(), gmcb, op, slot};
    if (b.n <= SERIAL_SUM_MAX) {
        sum_serial_kernel<4><<<1, SERIAL_THREADS, 0, ctx->stream>>>(b.n, reverse, term, fin);
    } else {
        const int g = sum_grid(b.n);
        double *partials = ctx->sumbuf.as<double>();
        unsigned *ticket = reinterpret_cast<unsigned *>(partials + (size_t)SUM_MAX_CTAS * 8);
        sum_tree_kernel<4><<<g, SUM_THREADS, 0, ctx->stream>>>(b.n, term, fin, partials, ticket);
    }
    SWCU_KERNEL_CHECK(ctx);
    return SWCU_OK;
}

int add_vec3(swcu_context *ctx, int n, const int32_t *lmask, const double *c3, double scale, const double *ix,
             const double *iy, const double *iz, double *ox, double *oy, double *oz)
{
    if (n <= 0) return SWCU_OK;
    add_vec3_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(n, lmask, c3, scale, ix, iy, iz, ox, oy, oz);
    SWCU_KERNEL_CHECK(ctx);
    return SWCU_OK;
}

}  // namespace

int ensure_step_state(swcu_context *ctx)
{
    if (!ctx->cbs.p) {
        SWCU_CUDA(ctx, ctx->cbs.ensure(sizeof(double) * CBS_DOUBLES));
        SWCU_CUDA(ctx, cudaMemsetAsync(ctx->cbs.p, 0, sizeof(double) * CBS_DOUBLES, ctx->stream));
    }
    if (!ctx->sumbuf.p) {
        const size_t bytes = sizeof(double) * (size_t)SUM_MAX_CTAS * 8 + 64;
        SWCU_CUDA(ctx, ctx->sumbuf.ensure(bytes));
        SWCU_CUDA(ctx, cudaMemsetAsync(ctx->sumbuf.p, 0, bytes, ctx->stream));
    }
    return SWCU_OK;
}

// second velocity (vb) and the begin/end position copies; vb starts as a copy of v
int ensure_helio(swcu_context *ctx, Body &b)
{
    if (b.helio_ready) return SWCU_OK;
    const size_t nb = sizeof(double) * (size_t)(b.n > 0 ? b.n : 1);
    DevBuf *all[] = {&b.wx, &b.wy, &b.wz, &b.bx, &b.by, &b.bz, &b.ex, &b.ey, &b.ez};
    for (DevBuf *d : all) SWCU_CUDA(ctx, d->ensure(nb));
    const DevBuf *src[] = {&b.vx, &b.vy, &b.vz, &b.rx, &b.ry, &b.rz, &b.rx, &b.ry, &b.rz};
    for (int k = 0; k < 9; ++k)
        SWCU_CUDA(ctx, cudaMemcpyAsync(all[k]->p, src[k]->p, nb, cudaMemcpyDeviceToDevice, ctx->stream));
    b.helio_ready = true;
    return SWCU_OK;
}

int pl_vh2vb(swcu_context *ctx, double gmcb)
{
    Body &pl = ctx->pl;
    if (pl.n == 0) return SWCU_OK;  // swiftest_util.f90:438
    SWCU_TRY(ensure_helio(ctx, pl));
    SWCU_TRY(reduce_gmv(ctx, pl, pl.vx.as<double>(), pl.vy.as<double>(), pl.vz.as<double>(), 0, false, false, gmcb,
                        FIN_VH2VB, 0));
    return add_vec3(ctx, pl.n, nullptr, ctx->cbs.as<double>() + CBS_VBCB, 1.0, pl.vx.as<double>(), pl.vy.as<double>(),
                    pl.vz.as<double>(), pl.wx.as<double>(), pl.wy.as<double>(), pl.wz.as<double>());
}

int pl_vb2vh(swcu_context *ctx, double gmcb)
{
    Body &pl = ctx->pl;
    if (pl.n == 0) return SWCU_OK;  // swiftest_util.f90:377
    SWCU_TRY(ensure_helio(ctx, pl));
    // swiftest_util.f90:377: `if (pl%status(i) /= INACTIVE)` -- the status, not lmask
    SWCU_TRY(reduce_gmv(ctx, pl, pl.wx.as<double>(), pl.wy.as<double>(), pl.wz.as<double>(), 2, true, true, gmcb,
                        FIN_VB2VH, 0));
    return add_vec3(ctx, pl.n, nullptr, ctx->cbs.as<double>() + CBS_VBCB, -1.0, pl.wx.as<double>(), pl.wy.as<double>(),
                    pl.wz.as<double>(), pl.vx.as<double>(), pl.vy.as<double>(), pl.vz.as<double>());
}

int pl_lindrift(swcu_context *ctx, double gmcb, double dt, int lbeg)
{
    Body &pl = ctx->pl;
    if (pl.n == 0) return SWCU_OK;  // helio_drift.f90:144
    SWCU_TRY(ensure_helio(ctx, pl));
    const int slot = lbeg ? CBS_PTBEG : CBS_PTEND;
    SWCU_TRY(reduce_gmv(ctx, pl, pl.wx.as<double>(), pl.wy.as<double>(), pl.wz.as<double>(), 1, false, false, gmcb,
                        FIN_PT, slot));
    return add_vec3(ctx, pl.n, pl.lmask.as<int32_t>(), ctx->cbs.as<double>() + slot, dt, pl.rx.as<double>(),
                    pl.ry.as<double>(), pl.rz.as<double>(), pl.rx.as<double>(), pl.ry.as<double>(), pl.rz.as<double>());
}

int tp_lindrift(swcu_context *ctx, double dt, int lbeg)
{
    Body &tp = ctx->tp;
    if (tp.n == 0) return SWCU_OK;
    SWCU_TRY(ensure_step_state(ctx));
    const int slot = lbeg ? CBS_PTBEG : CBS_PTEND;
    return add_vec3(ctx, tp.n, tp.lmask.as<int32_t>(), ctx->cbs.as<double>() + slot, dt, tp.rx.as<double>(),
                    tp.ry.as<double>(), tp.rz.as<double>(), tp.rx.as<double>(), tp.ry.as<double>(), tp.rz.as<double>());
}

int tp_vh2vb(swcu_context *ctx, int lbeg)
{
    Body &tp = ctx->tp;
    if (tp.n == 0) return SWCU_OK;
    SWCU_TRY(ensure_step_state(ctx));
    SWCU_TRY(ensure_helio(ctx, tp));
    const int slot = lbeg ? CBS_PTBEG : CBS_PTEND;  // vbcb = -pt: vb = vh + (-pt)
    return add_vec3(ctx, tp.n, tp.lmask.as<int32_t>(), ctx->cbs.as<double>() + slot, -1.0, tp.vx.as<double>(),
                    tp.vy.as<double>(), tp.vz.as<double>(), tp.wx.as<double>(), tp.wy.as<double>(), tp.wz.as<double>());
}

int tp_vb2vh(swcu_context *ctx, int lbeg)
{
    Body &tp = ctx->tp;
    if (tp.n == 0) return SWCU_OK;
    SWCU_TRY(ensure_step_state(ctx));
    SWCU_TRY(ensure_helio(ctx, tp));
    const int slot = lbeg ? CBS_PTBEG : CBS_PTEND;  // vh = vb - (-pt) = vb + pt
    return add_vec3(ctx, tp.n, tp.lmask.as<int32_t>(), ctx->cbs.as<double>() + slot, 1.0, tp.wx.as<double>(),
                    tp.wy.as<double>(), tp.wz.as<double>(), tp.vx.as<double>(), tp.vy.as<double>(), tp.vz.as<double>());
}

int kick_vb_save(swcu_context *ctx, Body &b, double dt, int save)
{
    if (b.n == 0) return SWCU_OK;
    SWCU_TRY(ensure_helio(ctx, b));
    double *sx = nullptr, *sy = nullptr, *sz = nullptr;
    if (save == 1) {
        sx = b.bx.as<double>(); sy = b.by.as<double>(); sz = b.bz.as<double>();
    } else if (save == 2) {
        sx = b.ex.as<double>(); sy = b.ey.as<double>(); sz = b.ez.as<double>();
    }
    kick_vb_save_kernel<<<cdiv(b.n, 256), 256, 0, ctx->stream>>>(
        b.n, b.lmask.as<int32_t>(), dt, b.ax.as<double>(), b.ay.as<double>(), b.az.as<double>(), b.wx.as<double>(),
        b.wy.as<double>(), b.wz.as<double>(), b.rx.as<double>(), b.ry.as<double>(), b.rz.as<double>(), sx, sy, sz);
    SWCU_KERNEL_CHECK(ctx);
    return SWCU_OK;
}

// helio_step_pl (helio_step.f90:37-78) on the resident pl population, nothing crosses PCIe but the failure count
int helio_step_pl(swcu_context *ctx, double gmcb, double dt, int variant, int lclose, int lfirst, int32_t *nfail)
{
    Body &pl = ctx->pl;
    if (nfail) *nfail = 0;
    if (pl.n == 0) return SWCU_OK;
    const double dth = 0.5 * dt;
    SWCU_TRY(ensure_helio(ctx, pl));
    // small systems: the whole step is ONE launch (drift_kernels.cu::helio_step_pl_small_kernel), bit-identical to the
    // reference's statement order for both loop variants; SWCU_HELIO_FUSED=0 keeps the multi-launch form below
    static const bool fused_ok = !(getenv("SWCU_HELIO_FUSED") && atoi(getenv("SWCU_HELIO_FUSED")) == 0);
    if (fused_ok && pl.n <= whm_small_max() && pl.nplm == pl.n && pl.slice0 == 0 && pl.slice1 == pl.n && ctx->tune_variant < 0) {
        const int flat = variant == SWCU_LOOP_FLAT || (variant == SWCU_LOOP_AUTO && pl.n >= 128);
        SWCU_TRY(ensure_step_state(ctx));
        SWCU_TRY(helio_step_pl_small(ctx, pl, gmcb, dt, flat, lclose, lfirst));
        if (nfail) {
            SWCU_CUDA(ctx, cudaMemcpyAsync(nfail, ctx->scratch64.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        return SWCU_OK;
    }
    auto body = [&]() -> int {  // the ~21 launches of a step (after the first)
        SWCU_TRY(pl_lindrift(ctx, gmcb, dth, 1));
        for (int half = 0; half < 2; ++half) {
            // helio_kick_vb_pl: ah = 0, interaction accelerations, set_beg_end, vb += ah*dth
            SWCU_TRY(fill3_f64(ctx, pl.ax.as<double>(), pl.ay.as<double>(), pl.az.as<double>(), 0.0, pl.n));
            SWCU_TRY(pl_accel_int(ctx, variant, lclose));
            SWCU_TRY(kick_vb_save(ctx, pl, dth, half == 0 ? 1 : 2));
            if (half == 0) SWCU_TRY(drift_bodies(ctx, pl, 0, pl.n, dt, 0, 0.0, nullptr, 1, gmcb));
        }
        SWCU_TRY(pl_lindrift(ctx, gmcb, dth, 0));
        return pl_vb2vh(ctx, gmcb);
    };
    // Steps after the first are the same ~21 launches with the same arguments (the two force evaluations flip the guard
    // parity of flat_prologue_kernel back to where it was): the second such step is captured into a CUDA graph and every
    // later one is a single cudaGraphLaunch -- at npl = 1e3 ... 1e4 the stream-ordered launches cost as much as the kernels.
    // Not with kernel timing on (events per launch group), not on a multi-GPU slice, SWCU_STEP_GRAPH=0 switches it off.
    const char *ge = getenv("SWCU_STEP_GRAPH");  // read per call: tests switch it at run time
    const bool graph_env = !(ge && atoi(ge) == 0);
    auto &G = ctx->helio_graph;
    const bool graph_ok = graph_env && !lfirst && !G.disabled && ctx->kernel_timing == 0 && !ctx->p2p.ready &&
                          pl.slice0 == 0 && pl.slice1 == pl.n && !getenv("SWCU_FLAT_TRACE");
    if (lfirst) SWCU_TRY(pl_vh2vb(ctx, gmcb));
    if (!graph_ok) {
        SWCU_TRY(body());
    } else {
        const bool same = G.n == pl.n && G.nplm == pl.nplm && G.variant == variant && G.lclose == lclose &&
                          G.tune_variant == ctx->tune_variant && G.tune_nsplit == ctx->tune_nsplit && G.gmcb == gmcb &&
                          G.dt == dt && G.generation == pl.generation && G.p0 == pl.rx.p && G.stream == ctx->stream &&
                          G.has_active == pl.has_active && G.tune_ib == ctx->tune_ib;
        if (same && G.exec) {
            SWCU_CUDA(ctx, cudaGraphLaunch(G.exec, ctx->stream));
            ctx->launches += G.launches;
            ++G.replays;
        } else if (same && G.warm >= 1) {  // every lazy allocation has happened: capture this step
            const long long l0 = ctx->launches;
            cudaGraph_t g = nullptr;
            int rc = SWCU_OK;
            if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                rc = body();
                const cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
                if (rc == SWCU_OK && e == cudaSuccess && g && cudaGraphInstantiate(&G.exec, g, 0) == cudaSuccess) {
                    G.launches = ctx->launches - l0;
                    cudaGraphDestroy(g);
                    SWCU_CUDA(ctx, cudaGraphLaunch(G.exec, ctx->stream));
                } else {  // something in the sequence cannot be captured: never try again, run the step the plain way
                    if (g) cudaGraphDestroy(g);
                    G.exec = nullptr;
                    G.disabled = true;
                    (void)cudaGetLastError();
                    ctx->launches = l0;
                    SWCU_TRY(body());
                }
            } else {
                G.disabled = true;
                (void)cudaGetLastError();
                SWCU_TRY(body());
            }
        } else {
            if (!same) {
                if (G.exec) cudaGraphExecDestroy(G.exec);
                G.exec = nullptr;
                G.n = pl.n, G.nplm = pl.nplm, G.variant = variant, G.lclose = lclose, G.tune_variant = ctx->tune_variant;
                G.tune_nsplit = ctx->tune_nsplit, G.gmcb = gmcb, G.dt = dt, G.generation = pl.generation, G.p0 = pl.rx.p;
                G.stream = ctx->stream, G.has_active = pl.has_active, G.tune_ib = ctx->tune_ib;
                G.warm = 0;
            }
            SWCU_TRY(body());
            ++G.warm;
        }
    }
    if (nfail) {  // the drift kernel counted its failures in scratch64[0]
        SWCU_CUDA(ctx, cudaMemcpyAsync(nfail, ctx->scratch64.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return SWCU_OK;
}

}  // namespace swcu
