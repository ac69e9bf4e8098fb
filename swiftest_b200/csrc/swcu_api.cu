// swcu_api.cu -- the C ABI of libswiftest_cuda.so (include/swiftest_cuda.h): context management, the
// host-pointer (tier 1) entry points, the device-resident (tier 2) entry points, and measurement helpers.
#include "swcu_internal.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

using namespace swcu;

#define SWCU_VERSION_NUMBER 100

namespace {

int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// ---------------- probes ----------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double seed)
{
    // 16 independent DFMA chains per thread, register resident
    double a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = seed + 1e-3 * (threadIdx.x + k);
    const double m = 1.0000000001, c = 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = fma(a[k], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
    if (s == 12345.6789) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true; keeps the chains alive
}

__global__ void flush_kernel(double2 *buf, size_t n)
{
    const double2 z = make_double2(0.0, 0.0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = z;
}

// second pass of the flush: stream a different region through L2 with loads, so that the dirty lines the write pass left
// are written back BEFORE the measured kernel starts (otherwise their eviction -- up to 126 MB of HBM writes, ~19 us --
// is charged to whatever runs next) and L2 ends up holding clean lines only
__global__ void flush_read_kernel(const double2 *buf, size_t n, double2 *sink)
{
    double ax = 0.0, ay = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double2 v = __ldcg(buf + i);
        ax += v.x, ay += v.y;
    }
    if (ax == 1.2345e300 && ay == -1.2345e300) *sink = make_double2(ax, ay);  // never true: keeps the loads alive
}

int check_ctx(swcu_context *ctx)
{
    if (!ctx) return SWCU_ERR_ARG;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return fail(ctx, SWCU_ERR_CUDA, "cudaSetDevice(%d): %s", ctx->device, cudaGetErrorString(e));
    return SWCU_OK;
}

Body &body_of(swcu_context *ctx, int kind) { return kind == SWCU_TP ? ctx->tp : ctx->pl; }

// upload a double array or fill it with a default
int put_or_fill(swcu_context *ctx, const double *h, int n, DevBuf &d, double dflt)
{
    SWCU_CUDA(ctx, d.ensure(sizeof(double) * (size_t)std::max(n, 1)));
    if (n <= 0) return SWCU_OK;
    if (h) {
        SWCU_CUDA(ctx, cudaMemcpyAsync(d.p, h, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        return SWCU_OK;
    }
    return fill_f64(ctx, d.as<double>(), dflt, n);
}

int put_or_fill_vec3(swcu_context *ctx, const double *h, int n, int slot, DevBuf &x, DevBuf &y, DevBuf &z)
{
    if (h) return upload_vec3(ctx, h, n, slot, x, y, z);
    SWCU_CUDA(ctx, x.ensure(sizeof(double) * (size_t)std::max(n, 1)));
    SWCU_CUDA(ctx, y.ensure(sizeof(double) * (size_t)std::max(n, 1)));
    SWCU_CUDA(ctx, z.ensure(sizeof(double) * (size_t)std::max(n, 1)));
    SWCU_TRY(fill_f64(ctx, x.as<double>(), 0.0, n));
    SWCU_TRY(fill_f64(ctx, y.as<double>(), 0.0, n));
    return fill_f64(ctx, z.as<double>(), 0.0, n);
}

SweepList sweep_list(const Body &b, int off, int n, bool with_renc)
{
    SweepList l;
    l.x = b.rx.as<double>() + off;
    l.y = b.ry.as<double>() + off;
    l.z = b.rz.as<double>() + off;
    l.vx = b.vx.as<double>() + off;
    l.vy = b.vy.as<double>() + off;
    l.vz = b.vz.as<double>() + off;
    l.renc = with_renc ? b.renc.as<double>() + off : nullptr;
    l.n = n;
    return l;
}

// stage a tier-1 population: positions (+ velocities, renc) into a scratch Body
int stage_population(swcu_context *ctx, Body &b, int n, const double *r, const double *v, const double *renc)
{
    SWCU_TRY(ensure_body(ctx, b, n));
    b.n = n;
    b.nplm = n;
    SWCU_TRY(upload_vec3(ctx, r, n, 0, b.rx, b.ry, b.rz));
    if (v) SWCU_TRY(upload_vec3(ctx, v, n, 1, b.vx, b.vy, b.vz));
    if (renc) SWCU_TRY(put_or_fill(ctx, renc, n, b.renc, 0.0));
    return SWCU_OK;
}

}  // namespace

// ======================================================================================================
// context
// ======================================================================================================
extern "C" int swcu_version(void) { return SWCU_VERSION_NUMBER; }

extern "C" int swcu_create(int device, swcu_context **out)
{
    if (!out) return SWCU_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return SWCU_ERR_NOGPU;
    if (device < 0 || device >= ndev) return SWCU_ERR_ARG;
    swcu_context *ctx = new (std::nothrow) swcu_context;
    if (!ctx) return SWCU_ERR_CUDA;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&ctx->prop, device) != cudaSuccess) {
        delete ctx;
        return SWCU_ERR_CUDA;
    }
    if (ctx->prop.major < 10 && !getenv("SWCU_ALLOW_ANY_ARCH")) {  // the library holds sm_100a code only
        delete ctx;
        return SWCU_ERR_NOGPU;
    }
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return SWCU_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    for (int f = 0; f < FAM_COUNT; ++f) {
        cudaEventCreate(&ctx->fam_ev0[f]);
        cudaEventCreate(&ctx->fam_ev1[f]);
    }
    ctx->tune_ib = env_int("SWCU_KICK_IB", 0);
    ctx->tune_nsplit = env_int("SWCU_KICK_NSPLIT", 0);
    ctx->tune_variant = env_int("SWCU_KICK_VARIANT", -1);
    *out = ctx;
    return SWCU_OK;
}

extern "C" int swcu_destroy(swcu_context *ctx)
{
    if (!ctx) return SWCU_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    swcu_p2p_close(ctx);
    ctx->p2p.F.release();
    ctx->p2p.flags.release();
    comm_release(ctx);
    ctx->pl.release();
    ctx->tp.release();
    ctx->s_pl.release();
    ctx->s_tp.release();
    for (auto &b : ctx->stage) b.release();
    for (auto &b : ctx->istage) b.release();
    ctx->partial.release();
    ctx->scratch64.release();
    ctx->flush.release();
    ctx->cbs.release();
    ctx->sumbuf.release();
    for (auto &b : ctx->lists) b.release();
    ctx->flat_trace.release();
    if (ctx->aio.ready) {
        cudaStreamDestroy(ctx->aio.h2d);
        cudaStreamDestroy(ctx->aio.d2h);
        for (int b = 0; b < 2; ++b) {
            cudaEventDestroy(ctx->aio.h2d_done[b]);
            cudaEventDestroy(ctx->aio.unpacked[b]);
            cudaEventDestroy(ctx->aio.packed[b]);
            cudaEventDestroy(ctx->aio.d2h_done[b]);
            ctx->aio.in[b].release();
            ctx->aio.out[b].release();
        }
        ctx->aio.ready = false;
    }
    ctx->flat_blockrad.release();
    ctx->flat_guard.release();
    {
        auto &W = ctx->whm;
        DevBuf *wb[] = {&W.xjx, &W.xjy, &W.xjz, &W.vjx, &W.vjy, &W.vjz, &W.eta, &W.muj, &W.ir3j};
        for (DevBuf *b : wb) b->release();
    }
    ctx->flat_redo.release();
    ctx->tp_discard.release();
    ctx->sendbuf.release();
    ctx->recvbuf.release();
    auto &E = ctx->enc;
    DevBuf *eb[] = {&E.keys_in, &E.keys_out, &E.vals_in, &E.vals_out, &E.cub_tmp, &E.cx, &E.cy, &E.cz, &E.cvx, &E.cvy,
                    &E.cvz, &E.crenc, &E.sx, &E.sy, &E.sz, &E.svx, &E.svy, &E.svz, &E.srenc, &E.sbody, &E.ibeg, &E.iend,
                    &E.nchunk, &E.choff, &E.owner, &E.cand, &E.cand_sorted, &E.uniq, &E.counters, &E.out1, &E.out2, &E.merged, &E.abase, &E.boxcnt, &E.bk_hist, &E.bk_offs};
    for (DevBuf *b : eb) b->release();
    if (E.h_counters) cudaFreeHost(E.h_counters);
    E.h_counters = nullptr;
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    for (int f = 0; f < FAM_COUNT; ++f) {
        for (cudaEvent_t e : ctx->fam_log0[f]) cudaEventDestroy(e);
        for (cudaEvent_t e : ctx->fam_log1[f]) cudaEventDestroy(e);
    }
    for (int f = 0; f < FAM_COUNT; ++f) {
        cudaEventDestroy(ctx->fam_ev0[f]);
        cudaEventDestroy(ctx->fam_ev1[f]);
    }
    if (ctx->helio_graph.exec) cudaGraphExecDestroy(ctx->helio_graph.exec);
    ctx->helio_graph.exec = nullptr;
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return SWCU_OK;
}

extern "C" const char *swcu_last_error(const swcu_context *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int swcu_set_stream(swcu_context *ctx, void *cuda_stream)
{
    SWCU_TRY(check_ctx(ctx));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return SWCU_OK;
}

extern "C" int swcu_synchronize(swcu_context *ctx)
{
    SWCU_TRY(check_ctx(ctx));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return p2p_check_error(ctx);  // reports a peer-memory exchange that timed out since the last check
}

extern "C" int swcu_device_info(swcu_context *ctx, int32_t *sm_count, int32_t *cc, int64_t *mem_bytes)
{
    if (!ctx) return SWCU_ERR_ARG;
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (cc) *cc = ctx->prop.major * 10 + ctx->prop.minor;
    if (mem_bytes) *mem_bytes = (int64_t)ctx->prop.totalGlobalMem;
    return SWCU_OK;
}

extern "C" int64_t swcu_launch_count(const swcu_context *ctx) { return ctx ? ctx->launches : 0; }

// ======================================================================================================
// tier 1: host pointers
// ======================================================================================================
extern "C" int swcu_kick_getacch_int_all_tri_pl(swcu_context *ctx, int32_t npl, int32_t nplm, const double *r,
                                                const double *Gmass, const double *radius, double *acc)
{
    SWCU_TRY(check_ctx(ctx));
    if (npl < 0 || nplm < 0 || nplm > npl) return fail(ctx, SWCU_ERR_ARG, "tri_pl: bad npl=%d nplm=%d", npl, nplm);
    if (npl == 0) return SWCU_OK;
    if (!r || !Gmass || !acc) return fail(ctx, SWCU_ERR_ARG, "tri_pl: null array");
    Body &b = ctx->s_pl;
    SWCU_TRY(ensure_body(ctx, b, npl));
    b.n = npl;
    b.nplm = nplm;
    SWCU_TRY(upload_vec3(ctx, r, npl, 0, b.rx, b.ry, b.rz));
    SWCU_TRY(upload_vec3(ctx, acc, npl, 1, b.ax, b.ay, b.az));
    SWCU_TRY(put_or_fill(ctx, Gmass, npl, b.Gm, 0.0));
    if (radius) SWCU_TRY(put_or_fill(ctx, radius, npl, b.radius, 0.0));
    SWCU_TRY(kick_pl_tri(ctx, b, radius != nullptr, 0, npl));
    SWCU_TRY(download_vec3(ctx, acc, npl, 1, b.ax, b.ay, b.az));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SWCU_OK;
}

extern "C" int swcu_kick_getacch_int_all_flat_pl(swcu_context *ctx, int32_t npl, int64_t nplpl, const int32_t *k_plpl,
                                                 const double *r, const double *Gmass, const double *radius, double *acc)
{
    SWCU_TRY(check_ctx(ctx));
    if (npl < 0 || nplpl < 0) return fail(ctx, SWCU_ERR_ARG, "flat_pl: bad npl=%d nplpl=%lld", npl, (long long)nplpl);
    if (npl == 0 || nplpl == 0) return SWCU_OK;
    if (!r || !Gmass || !acc) return fail(ctx, SWCU_ERR_ARG, "flat_pl: null array");
    Body &b = ctx->s_pl;
    SWCU_TRY(ensure_body(ctx, b, npl));
    b.n = npl;
    SWCU_TRY(upload_vec3(ctx, r, npl, 0, b.rx, b.ry, b.rz));
    SWCU_TRY(put_or_fill(ctx, Gmass, npl, b.Gm, 0.0));
    if (radius) SWCU_TRY(put_or_fill(ctx, radius, npl, b.radius, 0.0));

    if (k_plpl == nullptr) {
        // canonical flattened upper triangle: nplpl = nplm*npl - nplm*(nplm+1)/2 (symba_util.f90:202)
        int64_t nplm = -1;
        {
            // smallest nplm in [0,npl] whose pair count equals nplpl (count is strictly increasing on [0,npl-1])
            int64_t lo = 0, hi = npl;
            while (lo < hi) {
                const int64_t mid = (lo + hi) / 2;
                const int64_t cnt = mid * npl - mid * (mid + 1) / 2;
                if (cnt < nplpl) lo = mid + 1; else hi = mid;
            }
            if (lo * (int64_t)npl - lo * (lo + 1) / 2 == nplpl) nplm = lo;
        }
        if (nplm < 0)
            return fail(ctx, SWCU_ERR_ARG,
                        "flat_pl: nplpl=%lld is not nplm*npl-nplm*(nplm+1)/2 for any nplm (npl=%d); pass k_plpl",
                        (long long)nplpl, npl);
        if (nplm == npl - 1) nplm = npl;  // all pairs: the last body has no extra row of its own
        b.nplm = (int)nplm;
        SWCU_TRY(upload_vec3(ctx, acc, npl, 1, b.ax, b.ay, b.az));
        SWCU_TRY(kick_pl_flat(ctx, b, radius != nullptr, (int)nplm));
        SWCU_TRY(download_vec3(ctx, acc, npl, 1, b.ax, b.ay, b.az));
    } else {
        // explicit pair table: ahi/ahj accumulate from zero, then acc = acc + (ahi + ahj) (kick.f90:92-112)
        for (int64_t k = 0; k < nplpl; ++k)  // the reference trusts its own table; a foreign caller gets a checked error
            if (k_plpl[2 * k] < 1 || k_plpl[2 * k] > npl || k_plpl[2 * k + 1] < 1 || k_plpl[2 * k + 1] > npl)
                return fail(ctx, SWCU_ERR_ARG, "flat_pl: k_plpl(:,%lld) = (%d,%d) out of range 1..%d", (long long)k + 1,
                            k_plpl[2 * k], k_plpl[2 * k + 1], npl);
        SWCU_CUDA(ctx, ctx->istage[0].ensure(sizeof(int32_t) * 2 * (size_t)nplpl));
        SWCU_CUDA(ctx, ctx->istage[1].ensure(sizeof(int32_t) * 2 * (size_t)nplpl));
        // de-interleave k_plpl(2,nplpl) on the host side of the copy: two strided copies
        SWCU_CUDA(ctx, cudaMemcpy2DAsync(ctx->istage[0].p, sizeof(int32_t), k_plpl, 2 * sizeof(int32_t), sizeof(int32_t),
                                         (size_t)nplpl, cudaMemcpyHostToDevice, ctx->stream));
        SWCU_CUDA(ctx, cudaMemcpy2DAsync(ctx->istage[1].p, sizeof(int32_t), k_plpl + 1, 2 * sizeof(int32_t),
                                         sizeof(int32_t), (size_t)nplpl, cudaMemcpyHostToDevice, ctx->stream));
        SWCU_TRY(fill_f64(ctx, b.vx.as<double>(), 0.0, npl));
        SWCU_TRY(fill_f64(ctx, b.vy.as<double>(), 0.0, npl));
        SWCU_TRY(fill_f64(ctx, b.vz.as<double>(), 0.0, npl));
        SWCU_TRY(kick_pair_list(ctx, b, radius != nullptr, nplpl, ctx->istage[0].as<int32_t>(), ctx->istage[1].as<int32_t>(),
                                b.vx.as<double>(), b.vy.as<double>(), b.vz.as<double>()));
        SWCU_TRY(upload_vec3(ctx, acc, npl, 1, b.ax, b.ay, b.az));
        SWCU_TRY(axpy3(ctx, 1.0, b.vx.as<double>(), b.vy.as<double>(), b.vz.as<double>(), b.ax.as<double>(),
                       b.ay.as<double>(), b.az.as<double>(), nullptr, npl));
        SWCU_TRY(download_vec3(ctx, acc, npl, 1, b.ax, b.ay, b.az));
    }
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SWCU_OK;
}

extern "C" int swcu_symba_kick_subtract_encounters(swcu_context *ctx, int32_t npl, int64_t nenc, const int32_t *index1,
                                                   const int32_t *index2, const double *rh, const double *Gmass,
                                                   const double *radius, double *ah)
{
    SWCU_TRY(check_ctx(ctx));
    if (npl <= 0 || nenc <= 0) return SWCU_OK;  // symba_kick.f90:54,59
    if (!index1 || !index2 || !rh || !Gmass || !radius || !ah) return fail(ctx, SWCU_ERR_ARG, "symba subtract: null array");
    for (int64_t k = 0; k < nenc; ++k)
        if (index1[k] < 1 || index1[k] > npl || index2[k] < 1 || index2[k] > npl)
            return fail(ctx, SWCU_ERR_ARG, "symba subtract: pair %lld = (%d,%d) out of range 1..%d", (long long)k + 1,
                        index1[k], index2[k], npl);
    Body &b = ctx->s_pl;
    SWCU_TRY(ensure_body(ctx, b, npl));
    b.n = npl;
    SWCU_TRY(upload_vec3(ctx, rh, npl, 0, b.rx, b.ry, b.rz));
    SWCU_TRY(upload_vec3(ctx, ah, npl, 1, b.ax, b.ay, b.az));
    SWCU_TRY(put_or_fill(ctx, Gmass, npl, b.Gm, 0.0));
    SWCU_TRY(put_or_fill(ctx, radius, npl, b.radius, 0.0));
    SWCU_TRY(upload_arr(ctx, index1, sizeof(int32_t) * (size_t)nenc, ctx->istage[0]));
    SWCU_TRY(upload_arr(ctx, index2, sizeof(int32_t) * (size_t)nenc, ctx->istage[1]));
    SWCU_TRY(fill_f64(ctx, b.vx.as<double>(), 0.0, npl));  // ah_enc(:,:) = 0
    SWCU_TRY(fill_f64(ctx, b.vy.as<double>(), 0.0, npl));
    SWCU_TRY(fill_f64(ctx, b.vz.as<double>(), 0.0, npl));
    SWCU_TRY(kick_pair_list(ctx, b, true, nenc, ctx->istage[0].as<int32_t>(), ctx->istage[1].as<int32_t>(),
                            b.vx.as<double>(), b.vy.as<double>(), b.vz.as<double>()));
    SWCU_TRY(axpy3(ctx, -1.0, b.vx.as<double>(), b.vy.as<double>(), b.vz.as<double>(), b.ax.as<double>(),
                   b.ay.as<double>(), b.az.as<double>(), nullptr, npl));  // ah = ah - ah_enc
    SWCU_TRY(download_vec3(ctx, ah, npl, 1, b.ax, b.ay, b.az));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SWCU_OK;
}

extern "C" int swcu_kick_getacch_int_all_tp(swcu_context *ctx, int32_t ntp, int32_t npl, const double *rtp,
                                            const double *rpl, const double *GMpl, const int32_t *lmask, double *acc)
{
    SWCU_TRY(check_ctx(ctx));
    if (ntp < 0 || npl < 0) return fail(ctx, SWCU_ERR_ARG, "all_tp: bad ntp=%d npl=%d", ntp, npl);
    if (ntp == 0 || npl == 0) return SWCU_OK;  // kick.f90:61
    if (!rtp || !rpl || !GMpl || !lmask || !acc) return fail(ctx, SWCU_ERR_ARG, "all_tp: null array");
    Body &t = ctx->s_tp, &p = ctx->s_pl;
    SWCU_TRY(ensure_body(ctx, t, ntp));
    SWCU_TRY(ensure_body(ctx, p, npl));
    t.n = ntp;
    p.n = npl;
    SWCU_TRY(upload_vec3(ctx, rtp, ntp, 0, t.rx, t.ry, t.rz));
    SWCU_TRY(upload_vec3(ctx, acc, ntp, 1, t.ax, t.ay, t.az));
    SWCU_TRY(upload_arr(ctx, lmask, sizeof(int32_t) * (size_t)ntp, t.lmask));
    SWCU_TRY(upload_vec3(ctx, rpl, npl, 2, p.rx, p.ry, p.rz));
    SWCU_TRY(put_or_fill(ctx, GMpl, npl, p.Gm, 0.0));
    KickProblem k;
    k.xi = t.rx.as<double>(); k.yi = t.ry.as<double>(); k.zi = t.rz.as<double>(); k.radi = nullptr;
    k.row0 = 0; k.row1 = ntp;
    k.xj = p.rx.as<double>(); k.yj = p.ry.as<double>(); k.zj = p.rz.as<double>(); k.gmj = p.Gm.as<double>(); k.radj = nullptr;
    k.col0 = 0; k.col1 = npl;
    k.diag = false;
    k.lmask = t.lmask.as<int32_t>();
    k.ax = t.ax.as<double>(); k.ay = t.ay.as<double>(); k.az = t.az.as<double>();
    SWCU_TRY(kick_rows(ctx, k, FAM_PLTP));
    SWCU_TRY(download_vec3(ctx, acc, ntp, 1, t.ax, t.ay, t.az));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SWCU_OK;
}

extern "C" int swcu_drift_all(swcu_context *ctx, int32_t n, const double *mu, double *x, double *v, double dt, int32_t lgr,
                              double inv_c2, const int32_t *lmask, int32_t *iflag)
{
    SWCU_TRY(check_ctx(ctx));
    if (n < 0) return fail(ctx, SWCU_ERR_ARG, "drift_all: bad n=%d", n);
    if (n == 0) return SWCU_OK;  // drift.f90:81
    if (!mu || !x || !v || !lmask || !iflag) return fail(ctx, SWCU_ERR_ARG, "drift_all: null array");
    Body &b = ctx->s_tp;
    SWCU_TRY(ensure_body(ctx, b, n));
    b.n = n;
    SWCU_TRY(upload_vec3(ctx, x, n, 0, b.rx, b.ry, b.rz));
    SWCU_TRY(upload_vec3(ctx, v, n, 1, b.vx, b.vy, b.vz));
    SWCU_TRY(put_or_fill(ctx, mu, n, b.mu, 0.0));
    SWCU_TRY(upload_arr(ctx, lmask, sizeof(int32_t) * (size_t)n, b.lmask));
    SWCU_TRY(upload_arr(ctx, iflag, sizeof(int32_t) * (size_t)n, b.iflag));  // entries with lmask false keep their value
    SWCU_TRY(drift_bodies(ctx, b, 0, n, dt, lgr, inv_c2, nullptr));
    SWCU_TRY(download_vec3(ctx, x, n, 0, b.rx, b.ry, b.rz));
    SWCU_TRY(download_vec3(ctx, v, n, 1, b.vx, b.vy, b.vz));
    SWCU_CUDA(ctx, cudaMemcpyAsync(iflag, b.iflag.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SWCU_OK;
}

extern "C" int swcu_encounter_check_all_sort_and_sweep_plpl(swcu_context *ctx, int32_t npl, const double *r,
                                                            const double *v, const double *renc, double dt, int64_t *nenc)
{
    SWCU_TRY(check_ctx(ctx));
    if (!nenc || npl < 0) return fail(ctx, SWCU_ERR_ARG, "sas_plpl: bad argument");
    *nenc = 0;
    ctx->enc.nenc = 0;
    ctx->enc.result = nullptr;
    if (npl == 0) return SWCU_OK;
    if (!r || !v || !renc) return fail(ctx, SWCU_ERR_ARG, "sas_plpl: null array");
    SWCU_TRY(stage_population(ctx, ctx->s_pl, npl, r, v, renc));
    return encounter_sweep(ctx, sweep_list(ctx->s_pl, 0, npl, true), nullptr, dt, nenc);
}

extern "C" int swcu_encounter_check_all_sort_and_sweep_pltp(swcu_context *ctx, int32_t npl, int32_t ntp,
                                                            const double *rpl, const double *vpl, const double *rtp,
                                                            const double *vtp, const double *rencpl, double dt,
                                                            int64_t *nenc)
{
    SWCU_TRY(check_ctx(ctx));
    if (!nenc || npl < 0 || ntp < 0) return fail(ctx, SWCU_ERR_ARG, "sas_pltp: bad argument");
    *nenc = 0;
    ctx->enc.nenc = 0;
    ctx->enc.result = nullptr;
    if (npl == 0 || ntp == 0) return SWCU_OK;
    if (!rpl || !vpl || !rtp || !vtp || !rencpl) return fail(ctx, SWCU_ERR_ARG, "sas_pltp: null array");
    SWCU_TRY(stage_population(ctx, ctx->s_pl, npl, rpl, vpl, rencpl));
    SWCU_TRY(stage_population(ctx, ctx->s_tp, ntp, rtp, vtp, nullptr));
    SweepList l2 = sweep_list(ctx->s_tp, 0, ntp, false);
    return encounter_sweep(ctx, sweep_list(ctx->s_pl, 0, npl, true), &l2, dt, nenc);
}

extern "C" int swcu_encounter_check_all_sort_and_sweep_plplm(swcu_context *ctx, int32_t nplm, int32_t nplt,
                                                             const double *rplm, const double *vplm, const double *rplt,
                                                             const double *vplt, const double *rencm, const double *renct,
                                                             double dt, int64_t *nenc)
{
    SWCU_TRY(check_ctx(ctx));
    if (!nenc || nplm < 0 || nplt < 0) return fail(ctx, SWCU_ERR_ARG, "sas_plplm: bad argument");
    *nenc = 0;
    ctx->enc.nenc = 0;
    ctx->enc.result = nullptr;
    if (nplm == 0 || nplt == 0) return SWCU_OK;
    if (!rplm || !vplm || !rplt || !vplt || !rencm || !renct) return fail(ctx, SWCU_ERR_ARG, "sas_plplm: null array");
    SWCU_TRY(stage_population(ctx, ctx->s_pl, nplm, rplm, vplm, rencm));
    SWCU_TRY(stage_population(ctx, ctx->s_tp, nplt, rplt, vplt, renct));
    SweepList l2 = sweep_list(ctx->s_tp, 0, nplt, true);
    return encounter_sweep(ctx, sweep_list(ctx->s_pl, 0, nplm, true), &l2, dt, nenc);
}

extern "C" int swcu_encounter_check_all_plplm(swcu_context *ctx, int32_t nplm, int32_t nplt, const double *rplm,
                                              const double *vplm, const double *rplt, const double *vplt,
                                              const double *rencm, const double *renct, double dt, int64_t *nenc)
{
    SWCU_TRY(check_ctx(ctx));
    if (!nenc || nplm < 0 || nplt < 0) return fail(ctx, SWCU_ERR_ARG, "all_plplm: bad argument");
    *nenc = 0;
    ctx->enc.nenc = 0;
    ctx->enc.result = nullptr;
    if (nplm == 0) return SWCU_OK;
    if (!rplm || !vplm || !rencm) return fail(ctx, SWCU_ERR_ARG, "all_plplm: null array");
    SWCU_TRY(stage_population(ctx, ctx->s_pl, nplm, rplm, vplm, rencm));
    if (nplt == 0) return encounter_sweep(ctx, sweep_list(ctx->s_pl, 0, nplm, true), nullptr, dt, nenc);
    if (!rplt || !vplt || !renct) return fail(ctx, SWCU_ERR_ARG, "all_plplm: null array");
    SWCU_TRY(stage_population(ctx, ctx->s_tp, nplt, rplt, vplt, renct));
    return encounter_merge_plplm(ctx, sweep_list(ctx->s_pl, 0, nplm, true), sweep_list(ctx->s_tp, 0, nplt, true), dt, nenc);
}

// ======================================================================================================
// tier 2: device-resident populations
// ======================================================================================================
extern "C" int swcu_body_sync(swcu_context *ctx, int32_t kind, int32_t n, int32_t nplm, const double *r, const double *v,
                              const double *Gmass, const double *radius, const double *rhill, const double *mu,
                              const int32_t *lmask, uint64_t generation)
{
    SWCU_TRY(check_ctx(ctx));
    if (n < 0 || (kind != SWCU_PL && kind != SWCU_TP)) return fail(ctx, SWCU_ERR_ARG, "body_sync: bad kind/n");
    Body &b = body_of(ctx, kind);
    if (kind == SWCU_PL && (nplm < 0 || nplm > n)) return fail(ctx, SWCU_ERR_ARG, "body_sync: bad nplm=%d (npl=%d)", nplm, n);
    if (b.valid && b.generation == generation && b.n == n) return SWCU_OK;  // nothing changed on the host side
    // Peers hold CUDA-IPC mappings of the pl arrays and of F (sized by n): a population of another size would leave them
    // pointing at freed or too-short allocations.  The caller closes the mapping, re-syncs and exports/imports again.
    if (kind == SWCU_PL && ctx->p2p.ready && n != b.n)
        return fail(ctx, SWCU_ERR_STATE, "body_sync: npl changes %d -> %d while peer buffers are mapped; call swcu_p2p_close, "
                                         "sync, then swcu_p2p_export/import again", b.n, n);
    SWCU_TRY(ensure_body(ctx, b, n));
    b.helio_ready = false;  // vb, rbeg, rend are re-derived after a re-upload
    b.n = n;
    b.nplm = (kind == SWCU_PL) ? nplm : 0;
    b.slice0 = 0;
    b.slice1 = n;
    SWCU_TRY(put_or_fill_vec3(ctx, r, n, 0, b.rx, b.ry, b.rz));
    SWCU_TRY(put_or_fill_vec3(ctx, v, n, 1, b.vx, b.vy, b.vz));
    SWCU_TRY(put_or_fill_vec3(ctx, nullptr, n, 2, b.ax, b.ay, b.az));
    SWCU_TRY(put_or_fill(ctx, Gmass, n, b.Gm, 0.0));
    SWCU_TRY(put_or_fill(ctx, radius, n, b.radius, 0.0));
    SWCU_TRY(put_or_fill(ctx, rhill, n, b.rhill, 0.0));
    SWCU_TRY(put_or_fill(ctx, mu, n, b.mu, 0.0));
    SWCU_TRY(fill_f64(ctx, b.renc.as<double>(), 0.0, n));
    if (lmask)
        SWCU_TRY(upload_arr(ctx, lmask, sizeof(int32_t) * (size_t)n, b.lmask));
    else
        SWCU_TRY(fill_i32(ctx, b.lmask.as<int32_t>(), 1, n));
    SWCU_TRY(fill_i32(ctx, b.iflag.as<int32_t>(), 0, n));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // host arrays may be reused by the caller right away
    b.generation = generation;
    b.has_active = false;  // a new population: all active until swcu_body_set_active says otherwise
    b.valid = true;
    return SWCU_OK;
}

extern "C" int swcu_body_put(swcu_context *ctx, int32_t kind, const double *r, const double *v, const double *a,
                             const int32_t *lmask)
{
    SWCU_TRY(check_ctx(ctx));
    Body &b = body_of(ctx, kind);
    if (!b.valid) return fail(ctx, SWCU_ERR_STATE, "body_put: population not resident (call swcu_body_sync first)");
    if (r) SWCU_TRY(upload_vec3(ctx, r, b.n, 0, b.rx, b.ry, b.rz));
    if (v) SWCU_TRY(upload_vec3(ctx, v, b.n, 1, b.vx, b.vy, b.vz));
    if (a) SWCU_TRY(upload_vec3(ctx, a, b.n, 2, b.ax, b.ay, b.az));
    if (lmask) SWCU_TRY(upload_arr(ctx, lmask, sizeof(int32_t) * (size_t)b.n, b.lmask));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SWCU_OK;
}

extern "C" int swcu_body_get(swcu_context *ctx, int32_t kind, double *r, double *v, double *a, int32_t *iflag)
{
    SWCU_TRY(check_ctx(ctx));
    Body &b = body_of(ctx, kind);
    if (!b.valid) return fail(ctx, SWCU_ERR_STATE, "body_get: population not resident");
    if (r) SWCU_TRY(download_vec3(ctx, r, b.n, 0, b.rx, b.ry, b.rz));
    if (v) SWCU_TRY(download_vec3(ctx, v, b.n, 1, b.vx, b.vy, b.vz));
    if (a) SWCU_TRY(download_vec3(ctx, a, b.n, 2, b.ax, b.ay, b.az));
    if (iflag && b.n > 0)
        SWCU_CUDA(ctx, cudaMemcpyAsync(iflag, b.iflag.p, sizeof(int32_t) * (size_t)b.n, cudaMemcpyDeviceToHost, ctx->stream));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return p2p_check_error(ctx);
}

// slice forms: the host arrays hold bodies [i0, i1) only (3*(i1-i0) doubles each)
extern "C" int swcu_body_put_range(swcu_context *ctx, int32_t kind, int32_t i0, int32_t i1, const double *r, const double *v)
{
    SWCU_TRY(check_ctx(ctx));
    Body &b = body_of(ctx, kind);
    if (!b.valid) return fail(ctx, SWCU_ERR_STATE, "body_put_range: population not resident");
    if (i0 < 0 || i1 < i0 || i1 > b.n) return fail(ctx, SWCU_ERR_ARG, "body_put_range: bad range [%d,%d) of %d", i0, i1, b.n);
    const int m = i1 - i0;
    if (m == 0) return SWCU_OK;
    const size_t bytes = sizeof(double) * 3 * (size_t)m;
    const double *src[2] = {r, v};
    double *dst[2][3] = {{b.rx.as<double>(), b.ry.as<double>(), b.rz.as<double>()},
                         {b.vx.as<double>(), b.vy.as<double>(), b.vz.as<double>()}};
    for (int k = 0; k < 2; ++k) {
        if (!src[k]) continue;
        SWCU_CUDA(ctx, ctx->stage[k].ensure(bytes));
        SWCU_CUDA(ctx, cudaMemcpyAsync(ctx->stage[k].p, src[k], bytes, cudaMemcpyHostToDevice, ctx->stream));
        SWCU_TRY(aos_to_soa3(ctx, ctx->stage[k].as<double>(), dst[k][0] + i0, dst[k][1] + i0, dst[k][2] + i0, m));
    }
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SWCU_OK;
}

extern "C" int swcu_body_get_range(swcu_context *ctx, int32_t kind, int32_t i0, int32_t i1, double *r, double *v, double *a)
{
    SWCU_TRY(check_ctx(ctx));
    Body &b = body_of(ctx, kind);
    if (!b.valid) return fail(ctx, SWCU_ERR_STATE, "body_get_range: population not resident");
    if (i0 < 0 || i1 < i0 || i1 > b.n) return fail(ctx, SWCU_ERR_ARG, "body_get_range: bad range [%d,%d) of %d", i0, i1, b.n);
    const int m = i1 - i0;
    if (m == 0) return SWCU_OK;
    const size_t bytes = sizeof(double) * 3 * (size_t)m;
    double *dsth[3] = {r, v, a};
    const double *srcd[3][3] = {{b.rx.as<double>(), b.ry.as<double>(), b.rz.as<double>()},
                                {b.vx.as<double>(), b.vy.as<double>(), b.vz.as<double>()},
                                {b.ax.as<double>(), b.ay.as<double>(), b.az.as<double>()}};
    for (int k = 0; k < 3; ++k) {
        if (!dsth[k]) continue;
        SWCU_CUDA(ctx, ctx->stage[k].ensure(bytes));
        SWCU_TRY(soa_to_aos3(ctx, srcd[k][0] + i0, srcd[k][1] + i0, srcd[k][2] + i0, ctx->stage[k].as<double>(), m));
        SWCU_CUDA(ctx, cudaMemcpyAsync(dsth[k], ctx->stage[k].p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return p2p_check_error(ctx);
}

// ---- asynchronous slice I/O: the PCIe copies run on their own streams and overlap the compute stream's kernels ----
static int aio_setup(swcu_context *ctx)
{
    auto &A = ctx->aio;
    if (A.ready) return SWCU_OK;
    SWCU_CUDA(ctx, cudaStreamCreateWithFlags(&A.h2d, cudaStreamNonBlocking));
    SWCU_CUDA(ctx, cudaStreamCreateWithFlags(&A.d2h, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
        SWCU_CUDA(ctx, cudaEventCreateWithFlags(&A.h2d_done[b], cudaEventDisableTiming));
        SWCU_CUDA(ctx, cudaEventCreateWithFlags(&A.unpacked[b], cudaEventDisableTiming));
        SWCU_CUDA(ctx, cudaEventCreateWithFlags(&A.packed[b], cudaEventDisableTiming));
        SWCU_CUDA(ctx, cudaEventCreateWithFlags(&A.d2h_done[b], cudaEventDisableTiming));
    }
    A.ready = true;
    return SWCU_OK;
}

// Enqueue only: the H2D copies go to the copy stream (they overlap whatever the compute stream is running), the
// transposing kernels to the compute stream behind everything queued so far.  r, v must be page-locked and stay
// unchanged until swcu_io_wait / swcu_synchronize.
extern "C" int swcu_body_put_range_async(swcu_context *ctx, int32_t kind, int32_t i0, int32_t i1, const double *r,
                                         const double *v)
{
    SWCU_TRY(check_ctx(ctx));
    Body &b = body_of(ctx, kind);
    if (!b.valid) return fail(ctx, SWCU_ERR_STATE, "body_put_range_async: population not resident");
    if (i0 < 0 || i1 < i0 || i1 > b.n) return fail(ctx, SWCU_ERR_ARG, "body_put_range_async: bad range [%d,%d) of %d", i0, i1, b.n);
    const int m = i1 - i0;
    if (m == 0 || (!r && !v)) return SWCU_OK;
    SWCU_TRY(aio_setup(ctx));
    auto &A = ctx->aio;
    const int k = (int)(A.put_seq++ & 1);
    const size_t bytes = sizeof(double) * 3 * (size_t)m;
    if (A.in[k].cap < 2 * bytes + 256) {  // (re)allocation: nothing may still be using the old buffer
        SWCU_CUDA(ctx, cudaStreamSynchronize(A.h2d));
        SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        SWCU_CUDA(ctx, A.in[k].ensure(2 * bytes));
    }
    SWCU_CUDA(ctx, cudaStreamWaitEvent(A.h2d, A.unpacked[k], 0));  // the previous contents of this buffer were consumed
    double *st = A.in[k].as<double>();
    if (r) SWCU_CUDA(ctx, cudaMemcpyAsync(st, r, bytes, cudaMemcpyHostToDevice, A.h2d));
    if (v) SWCU_CUDA(ctx, cudaMemcpyAsync(st + 3 * (size_t)m, v, bytes, cudaMemcpyHostToDevice, A.h2d));
    SWCU_CUDA(ctx, cudaEventRecord(A.h2d_done[k], A.h2d));
    SWCU_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, A.h2d_done[k], 0));
    if (r) SWCU_TRY(aos_to_soa3(ctx, st, b.rx.as<double>() + i0, b.ry.as<double>() + i0, b.rz.as<double>() + i0, m));
    if (v) SWCU_TRY(aos_to_soa3(ctx, st + 3 * (size_t)m, b.vx.as<double>() + i0, b.vy.as<double>() + i0, b.vz.as<double>() + i0, m));
    SWCU_CUDA(ctx, cudaEventRecord(A.unpacked[k], ctx->stream));
    return SWCU_OK;
}

// Enqueue only: the transposing kernels run on the compute stream where the call is made (after the step that produced
// the values), the D2H copies on the copy stream behind them.  r, v, a must be page-locked; their contents are defined
// after swcu_io_wait / swcu_synchronize.
extern "C" int swcu_body_get_range_async(swcu_context *ctx, int32_t kind, int32_t i0, int32_t i1, double *r, double *v,
                                         double *a)
{
    SWCU_TRY(check_ctx(ctx));
    Body &b = body_of(ctx, kind);
    if (!b.valid) return fail(ctx, SWCU_ERR_STATE, "body_get_range_async: population not resident");
    if (i0 < 0 || i1 < i0 || i1 > b.n) return fail(ctx, SWCU_ERR_ARG, "body_get_range_async: bad range [%d,%d) of %d", i0, i1, b.n);
    const int m = i1 - i0;
    if (m == 0 || (!r && !v && !a)) return SWCU_OK;
    SWCU_TRY(aio_setup(ctx));
    auto &A = ctx->aio;
    const int k = (int)(A.get_seq++ & 1);
    const size_t bytes = sizeof(double) * 3 * (size_t)m;
    if (A.out[k].cap < 3 * bytes + 256) {
        SWCU_CUDA(ctx, cudaStreamSynchronize(A.d2h));
        SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        SWCU_CUDA(ctx, A.out[k].ensure(3 * bytes));
    }
    SWCU_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, A.d2h_done[k], 0));  // the previous copy out of this buffer is done
    double *st = A.out[k].as<double>();
    double *dsth[3] = {r, v, a};
    const double *srcd[3][3] = {{b.rx.as<double>(), b.ry.as<double>(), b.rz.as<double>()},
                                {b.vx.as<double>(), b.vy.as<double>(), b.vz.as<double>()},
                                {b.ax.as<double>(), b.ay.as<double>(), b.az.as<double>()}};
    for (int q = 0; q < 3; ++q)
        if (dsth[q]) SWCU_TRY(soa_to_aos3(ctx, srcd[q][0] + i0, srcd[q][1] + i0, srcd[q][2] + i0, st + 3 * (size_t)m * q, m));
    SWCU_CUDA(ctx, cudaEventRecord(A.packed[k], ctx->stream));
    SWCU_CUDA(ctx, cudaStreamWaitEvent(A.d2h, A.packed[k], 0));
    for (int q = 0; q < 3; ++q)
        if (dsth[q]) SWCU_CUDA(ctx, cudaMemcpyAsync(dsth[q], st + 3 * (size_t)m * q, bytes, cudaMemcpyDeviceToHost, A.d2h));
    SWCU_CUDA(ctx, cudaEventRecord(A.d2h_done[k], A.d2h));
    return SWCU_OK;
}

// all asynchronous transfers issued so far are complete (host buffers of the gets are filled, of the puts reusable)
extern "C" int swcu_io_wait(swcu_context *ctx)
{
    SWCU_TRY(check_ctx(ctx));
    if (ctx->aio.ready) {
        SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->aio.h2d));
        SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->aio.d2h));
    } else {
        SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return p2p_check_error(ctx);
}

extern "C" int swcu_body_count(swcu_context *ctx, int32_t kind, int32_t *n, int32_t *nplm, uint64_t *generation)
{
    if (!ctx) return SWCU_ERR_ARG;
    Body &b = body_of(ctx, kind);
    if (n) *n = b.valid ? b.n : 0;
    if (nplm) *nplm = b.valid ? b.nplm : 0;
    if (generation) *generation = b.generation;
    return SWCU_OK;
}

extern "C" int swcu_body_zero_accel(swcu_context *ctx, int32_t kind)
{
    SWCU_TRY(check_ctx(ctx));
    Body &b = body_of(ctx, kind);
    if (!b.valid) return fail(ctx, SWCU_ERR_STATE, "zero_accel: population not resident");
    return fill3_f64(ctx, b.ax.as<double>(), b.ay.as<double>(), b.az.as<double>(), 0.0, b.n);
}

namespace swcu {
int pl_accel_int(swcu_context *ctx, int loop_variant, int lclose)
{
    Body &pl = ctx->pl;
    if (pl.n == 0) return SWCU_OK;
    int variant = ctx->tune_variant >= 0 ? ctx->tune_variant : loop_variant;
    // AUTO by measurement (scripts/crossover_scan.py, profiles/r02_crossover.md): the third-law kernel is faster from
    // npl = 128 up (0.88x the full-row time at 128, 0.52x at 512, 0.29x at 1e4, 0.61x at 1e5); below that both are launch
    // bound within 5 % of each other and the bitwise-reproducible full-row kernel is taken
    if (variant == SWCU_LOOP_AUTO) variant = (pl.n >= 128) ? SWCU_LOOP_FLAT : SWCU_LOOP_TRIANGULAR;
    if (variant == SWCU_LOOP_FLAT) return kick_pl_flat(ctx, pl, lclose != 0, pl.nplm);
    return kick_pl_tri(ctx, pl, lclose != 0, pl.slice0, pl.slice1);
}
}  // namespace swcu

extern "C" int swcu_pl_accel_int(swcu_context *ctx, int32_t loop_variant, int32_t lclose)
{
    SWCU_TRY(check_ctx(ctx));
    if (!ctx->pl.valid) return fail(ctx, SWCU_ERR_STATE, "pl_accel_int: pl population not resident");
    return pl_accel_int(ctx, loop_variant, lclose);
}

extern "C" int swcu_tp_accel_int(swcu_context *ctx)
{
    SWCU_TRY(check_ctx(ctx));
    Body &tp = ctx->tp, &pl = ctx->pl;
    if (!tp.valid || !pl.valid) return fail(ctx, SWCU_ERR_STATE, "tp_accel_int: tp and pl populations must be resident");
    if (tp.n == 0 || pl.n == 0) return SWCU_OK;  // kick.f90:61
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

extern "C" int swcu_pl_set_renc(swcu_context *ctx, int32_t irec)
{
    SWCU_TRY(check_ctx(ctx));
    if (!ctx->pl.valid) return fail(ctx, SWCU_ERR_STATE, "pl_set_renc: pl population not resident");
    if (irec < 0) return fail(ctx, SWCU_ERR_ARG, "pl_set_renc: irec=%d", irec);
    return set_renc(ctx, ctx->pl, irec);
}

extern "C" int swcu_body_drift(swcu_context *ctx, int32_t kind, double dt, int32_t lgr, double inv_c2, int32_t *nfail)
{
    SWCU_TRY(check_ctx(ctx));
    Body &b = body_of(ctx, kind);
    if (!b.valid) return fail(ctx, SWCU_ERR_STATE, "body_drift: population not resident");
    const int i0 = (kind == SWCU_PL) ? b.slice0 : 0, i1 = (kind == SWCU_PL) ? b.slice1 : b.n;
    return drift_bodies(ctx, b, i0, i1, dt, lgr, inv_c2, nfail);
}

extern "C" int swcu_whm_tp_step(swcu_context *ctx, double dt, const double *ah0, int32_t *nfail)
{
    SWCU_TRY(check_ctx(ctx));
    if (!ctx->tp.valid || !ctx->pl.valid) return fail(ctx, SWCU_ERR_STATE, "whm_tp_step: tp and pl populations must be resident");
    if (ctx->tp.n == 0 || ctx->pl.n == 0) return SWCU_OK;  // whm_kick.f90:91
    return whm_tp_step(ctx, ctx->tp, ctx->pl, dt, ah0, nfail);
}

extern "C" int swcu_whm_step_pl(swcu_context *ctx, double GMcb, double dt, int32_t loop_variant, int32_t lclose, int32_t lfirst,
                                int32_t *nfail)
{
    SWCU_TRY(check_ctx(ctx));
    if (!ctx->pl.valid) return fail(ctx, SWCU_ERR_STATE, "whm_step_pl: pl population not resident");
    return whm_step_pl(ctx, GMcb, dt, loop_variant, lclose, lfirst, nfail);
}

extern "C" int swcu_whm_tp_first_accel(swcu_context *ctx)
{
    SWCU_TRY(check_ctx(ctx));
    if (!ctx->tp.valid || !ctx->pl.valid) return fail(ctx, SWCU_ERR_STATE, "whm_tp_first_accel: tp and pl populations must be resident");
    return whm_tp_first_accel(ctx);
}

extern "C" int swcu_whm_get_jacobi(swcu_context *ctx, double *xj, double *vj)
{
    SWCU_TRY(check_ctx(ctx));
    if (!ctx->pl.valid) return fail(ctx, SWCU_ERR_STATE, "whm_get_jacobi: pl population not resident");
    return whm_get_jacobi(ctx, xj, vj);
}

extern "C" int swcu_body_kick_velocity(swcu_context *ctx, int32_t kind, double dt)
{
    SWCU_TRY(check_ctx(ctx));
    Body &b = body_of(ctx, kind);
    if (!b.valid) return fail(ctx, SWCU_ERR_STATE, "kick_velocity: population not resident");
    return axpy3(ctx, dt, b.ax.as<double>(), b.ay.as<double>(), b.az.as<double>(), b.vx.as<double>(), b.vy.as<double>(),
                 b.vz.as<double>(), b.lmask.as<int32_t>(), b.n);
}

extern "C" int swcu_pl_encounter_check(swcu_context *ctx, double dt, int64_t *nenc)
{
    SWCU_TRY(check_ctx(ctx));
    Body &pl = ctx->pl;
    if (!nenc) return fail(ctx, SWCU_ERR_ARG, "pl_encounter_check: null nenc");
    *nenc = 0;
    ctx->enc.nenc = 0;
    ctx->enc.result = nullptr;
    if (!pl.valid) return fail(ctx, SWCU_ERR_STATE, "pl_encounter_check: pl population not resident");
    if (pl.n == 0) return SWCU_OK;
    const int nplm = pl.nplm, nplt = pl.n - pl.nplm;
    if (nplt == 0) return encounter_sweep(ctx, sweep_list(pl, 0, pl.n, true), nullptr, dt, nenc);  // symba_encounter_check.f90:45-46
    if (nplm == 0) return SWCU_OK;
    return encounter_merge_plplm(ctx, sweep_list(pl, 0, nplm, true), sweep_list(pl, nplm, nplt, true), dt, nenc);
}

extern "C" int swcu_tp_encounter_check(swcu_context *ctx, double dt, int64_t *nenc)
{
    SWCU_TRY(check_ctx(ctx));
    Body &pl = ctx->pl, &tp = ctx->tp;
    if (!nenc) return fail(ctx, SWCU_ERR_ARG, "tp_encounter_check: null nenc");
    *nenc = 0;
    ctx->enc.nenc = 0;
    ctx->enc.result = nullptr;
    if (!pl.valid || !tp.valid) return fail(ctx, SWCU_ERR_STATE, "tp_encounter_check: populations not resident");
    if (pl.n == 0 || tp.n == 0) return SWCU_OK;
    SweepList l2 = sweep_list(tp, 0, tp.n, false);
    return encounter_sweep(ctx, sweep_list(pl, 0, pl.n, true), &l2, dt, nenc);
}

// the same two checks with the all-pairs predicate instead of the sweep (ENCOUNTER_CHECK TRIANGULAR, encounter_check.f90:436-570)
extern "C" int swcu_pl_encounter_check_triangular(swcu_context *ctx, double dt, int64_t *nenc)
{
    SWCU_TRY(check_ctx(ctx));
    Body &pl = ctx->pl;
    if (!nenc) return fail(ctx, SWCU_ERR_ARG, "pl_encounter_check_triangular: null nenc");
    *nenc = 0;
    ctx->enc.nenc = 0;
    ctx->enc.result = nullptr;
    if (!pl.valid) return fail(ctx, SWCU_ERR_STATE, "pl_encounter_check_triangular: pl population not resident");
    if (pl.n == 0) return SWCU_OK;
    const int nplm = pl.nplm, nplt = pl.n - pl.nplm;
    if (nplt == 0) return encounter_triangular(ctx, sweep_list(pl, 0, pl.n, true), nullptr, dt, nenc);
    if (nplm == 0) return SWCU_OK;
    return encounter_merge_plplm(ctx, sweep_list(pl, 0, nplm, true), sweep_list(pl, nplm, nplt, true), dt, nenc, true);
}

extern "C" int swcu_tp_encounter_check_triangular(swcu_context *ctx, double dt, int64_t *nenc)
{
    SWCU_TRY(check_ctx(ctx));
    Body &pl = ctx->pl, &tp = ctx->tp;
    if (!nenc) return fail(ctx, SWCU_ERR_ARG, "tp_encounter_check_triangular: null nenc");
    *nenc = 0;
    ctx->enc.nenc = 0;
    ctx->enc.result = nullptr;
    if (!pl.valid || !tp.valid) return fail(ctx, SWCU_ERR_STATE, "tp_encounter_check_triangular: populations not resident");
    if (pl.n == 0 || tp.n == 0) return SWCU_OK;
    SweepList l2 = sweep_list(tp, 0, tp.n, false);
    return encounter_triangular(ctx, sweep_list(pl, 0, pl.n, true), &l2, dt, nenc);
}

// ======================================================================================================
// tier 2: the O(N) glue of the democratic-heliocentric step
// ======================================================================================================
namespace {
int fetch_cbs(swcu_context *ctx, int slot, double *out3)
{
    if (!out3) return SWCU_OK;
    SWCU_CUDA(ctx, cudaMemcpyAsync(out3, ctx->cbs.as<double>() + slot, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SWCU_OK;
}
int need_pl(swcu_context *ctx, const char *who)
{
    SWCU_TRY(check_ctx(ctx));
    if (!ctx->pl.valid) return fail(ctx, SWCU_ERR_STATE, "%s: pl population not resident", who);
    return SWCU_OK;
}
int need_tp(swcu_context *ctx, const char *who)
{
    SWCU_TRY(check_ctx(ctx));
    if (!ctx->tp.valid) return fail(ctx, SWCU_ERR_STATE, "%s: tp population not resident", who);
    return SWCU_OK;
}
}  // namespace

extern "C" int swcu_pl_vh2vb(swcu_context *ctx, double GMcb, double *vbcb)
{
    SWCU_TRY(need_pl(ctx, "pl_vh2vb"));
    if (ctx->pl.n == 0) return SWCU_OK;
    SWCU_TRY(pl_vh2vb(ctx, GMcb));
    return fetch_cbs(ctx, CBS_VBCB, vbcb);
}

extern "C" int swcu_pl_vb2vh(swcu_context *ctx, double GMcb, double *vbcb)
{
    SWCU_TRY(need_pl(ctx, "pl_vb2vh"));
    if (ctx->pl.n == 0) return SWCU_OK;
    SWCU_TRY(pl_vb2vh(ctx, GMcb));
    return fetch_cbs(ctx, CBS_VBCB, vbcb);
}

extern "C" int swcu_pl_lindrift(swcu_context *ctx, double GMcb, double dt, int32_t lbeg, double *pt)
{
    SWCU_TRY(need_pl(ctx, "pl_lindrift"));
    if (ctx->pl.n == 0) return SWCU_OK;
    SWCU_TRY(pl_lindrift(ctx, GMcb, dt, lbeg));
    return fetch_cbs(ctx, lbeg ? CBS_PTBEG : CBS_PTEND, pt);
}

extern "C" int swcu_cb_set_pt(swcu_context *ctx, const double *ptbeg, const double *ptend)
{
    SWCU_TRY(check_ctx(ctx));
    SWCU_TRY(ensure_step_state(ctx));
    if (ptbeg) SWCU_CUDA(ctx, cudaMemcpyAsync(ctx->cbs.as<double>() + CBS_PTBEG, ptbeg, 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (ptend) SWCU_CUDA(ctx, cudaMemcpyAsync(ctx->cbs.as<double>() + CBS_PTEND, ptend, 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SWCU_OK;
}

extern "C" int swcu_cb_get_pt(swcu_context *ctx, double *ptbeg, double *ptend)
{
    SWCU_TRY(check_ctx(ctx));
    SWCU_TRY(ensure_step_state(ctx));
    SWCU_TRY(fetch_cbs(ctx, CBS_PTBEG, ptbeg));
    return fetch_cbs(ctx, CBS_PTEND, ptend);
}

extern "C" int swcu_tp_lindrift(swcu_context *ctx, double dt, int32_t lbeg)
{
    SWCU_TRY(need_tp(ctx, "tp_lindrift"));
    return tp_lindrift(ctx, dt, lbeg);
}

extern "C" int swcu_tp_vh2vb(swcu_context *ctx, int32_t lbeg)
{
    SWCU_TRY(need_tp(ctx, "tp_vh2vb"));
    return tp_vh2vb(ctx, lbeg);
}

extern "C" int swcu_tp_vb2vh(swcu_context *ctx, int32_t lbeg)
{
    SWCU_TRY(need_tp(ctx, "tp_vb2vh"));
    return tp_vb2vh(ctx, lbeg);
}

extern "C" int swcu_body_kick_vb(swcu_context *ctx, int32_t kind, double dt, int32_t lbeg)
{
    SWCU_TRY(check_ctx(ctx));
    Body &b = body_of(ctx, kind);
    if (!b.valid) return fail(ctx, SWCU_ERR_STATE, "body_kick_vb: population not resident");
    return kick_vb_save(ctx, b, dt, kind == SWCU_PL ? (lbeg ? 1 : 2) : 0);
}

extern "C" int swcu_body_drift_vb(swcu_context *ctx, int32_t kind, double GMcb, double dt, int32_t *nfail)
{
    SWCU_TRY(check_ctx(ctx));
    Body &b = body_of(ctx, kind);
    if (!b.valid) return fail(ctx, SWCU_ERR_STATE, "body_drift_vb: population not resident");
    SWCU_TRY(ensure_helio(ctx, b));
    return drift_bodies(ctx, b, 0, b.n, dt, 0, 0.0, nfail, 1, GMcb);
}

extern "C" int swcu_body_put_vb(swcu_context *ctx, int32_t kind, const double *vb)
{
    SWCU_TRY(check_ctx(ctx));
    Body &b = body_of(ctx, kind);
    if (!b.valid) return fail(ctx, SWCU_ERR_STATE, "body_put_vb: population not resident");
    if (!vb) return fail(ctx, SWCU_ERR_ARG, "body_put_vb: null array");
    SWCU_TRY(ensure_helio(ctx, b));
    SWCU_TRY(upload_vec3(ctx, vb, b.n, 1, b.wx, b.wy, b.wz));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SWCU_OK;
}

extern "C" int swcu_body_set_active(swcu_context *ctx, int32_t kind, const int32_t *lactive)
{
    SWCU_TRY(check_ctx(ctx));
    Body &b = body_of(ctx, kind);
    if (!b.valid) return fail(ctx, SWCU_ERR_STATE, "body_set_active: population not resident");
    if (!lactive) {
        b.has_active = false;  // every body active again
        return SWCU_OK;
    }
    SWCU_TRY(upload_arr(ctx, lactive, sizeof(int32_t) * (size_t)b.n, b.lactive));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    b.has_active = true;
    return SWCU_OK;
}

extern "C" int swcu_body_get_vb(swcu_context *ctx, int32_t kind, double *vb, double *rbeg, double *rend)
{
    SWCU_TRY(check_ctx(ctx));
    Body &b = body_of(ctx, kind);
    if (!b.valid) return fail(ctx, SWCU_ERR_STATE, "body_get_vb: population not resident");
    SWCU_TRY(ensure_helio(ctx, b));
    if (vb) SWCU_TRY(download_vec3(ctx, vb, b.n, 1, b.wx, b.wy, b.wz));
    if (rbeg) SWCU_TRY(download_vec3(ctx, rbeg, b.n, 0, b.bx, b.by, b.bz));
    if (rend) SWCU_TRY(download_vec3(ctx, rend, b.n, 2, b.ex, b.ey, b.ez));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SWCU_OK;
}

extern "C" int swcu_helio_step_pl(swcu_context *ctx, double GMcb, double dt, int32_t loop_variant, int32_t lclose,
                                  int32_t lfirst, int32_t *nfail)
{
    SWCU_TRY(need_pl(ctx, "helio_step_pl"));
    if (!(GMcb > 0.0)) return fail(ctx, SWCU_ERR_ARG, "helio_step_pl: GMcb must be positive");
    return helio_step_pl(ctx, GMcb, dt, loop_variant, lclose, lfirst, nfail);
}

extern "C" int swcu_helio_step_tp(swcu_context *ctx, double GMcb, double dt, int32_t lfirst, int32_t *nfail)
{
    SWCU_TRY(need_tp(ctx, "helio_step_tp"));
    if (!ctx->pl.valid) return fail(ctx, SWCU_ERR_STATE, "helio_step_tp: pl population not resident");
    if (!(GMcb > 0.0)) return fail(ctx, SWCU_ERR_ARG, "helio_step_tp: GMcb must be positive");
    if (nfail) *nfail = 0;
    if (ctx->tp.n == 0) return SWCU_OK;  // helio_step.f90:98
    SWCU_TRY(ensure_step_state(ctx));
    SWCU_TRY(ensure_helio(ctx, ctx->tp));
    SWCU_TRY(ensure_helio(ctx, ctx->pl));
    return helio_tp_step(ctx, ctx->tp, ctx->pl, GMcb, dt, lfirst, nfail);
}

// ======================================================================================================
// tier 1: energy and momentum of the massive bodies
// ======================================================================================================
namespace {
int stage_energy(swcu_context *ctx, int32_t npl, const int32_t *lmask, const double *Gmass, const double *mass,
                 const double *radius, const double *rb, const double *vb)
{
    Body &b = ctx->s_pl;
    SWCU_TRY(stage_population(ctx, b, npl, rb, vb, nullptr));
    SWCU_TRY(put_or_fill(ctx, Gmass, npl, b.Gm, 0.0));
    SWCU_TRY(put_or_fill(ctx, mass, npl, b.mu, 0.0));
    SWCU_TRY(put_or_fill(ctx, radius, npl, b.radius, 1.0));
    if (lmask)
        SWCU_TRY(upload_arr(ctx, lmask, sizeof(int32_t) * (size_t)npl, b.lmask));
    else
        SWCU_TRY(fill_i32(ctx, b.lmask.as<int32_t>(), 1, npl));
    return SWCU_OK;
}
}  // namespace

extern "C" int swcu_util_get_potential_energy(swcu_context *ctx, int32_t npl, const int32_t *lmask, double GMcb,
                                              const double *Gmass, const double *mass, const double *rb, double *pe)
{
    SWCU_TRY(check_ctx(ctx));
    if (!pe || npl < 0) return fail(ctx, SWCU_ERR_ARG, "get_potential_energy: bad argument");
    *pe = 0.0;
    if (npl == 0) return SWCU_OK;
    if (!Gmass || !mass || !rb) return fail(ctx, SWCU_ERR_ARG, "get_potential_energy: null array");
    SWCU_TRY(stage_energy(ctx, npl, lmask, Gmass, mass, nullptr, rb, nullptr));
    double s[8];
    SWCU_TRY(energy_and_momentum(ctx, ctx->s_pl, GMcb, 0, true, s));
    *pe = -s[1] - s[2];
    return SWCU_OK;
}

extern "C" int swcu_util_get_energy_and_momentum(swcu_context *ctx, int32_t npl, const int32_t *lmask, double GMcb,
                                                 double mass_cb, const double *rbcb, const double *vbcb,
                                                 const double *Gmass, const double *mass, const double *radius,
                                                 const double *rb, const double *vb, int32_t lclose, double *out8)
{
    SWCU_TRY(check_ctx(ctx));
    if (!out8 || npl < 0 || !rbcb || !vbcb) return fail(ctx, SWCU_ERR_ARG, "get_energy_and_momentum: bad argument");
    if (npl > 0 && (!Gmass || !mass || !rb || !vb || (lclose && !radius)))
        return fail(ctx, SWCU_ERR_ARG, "get_energy_and_momentum: null array");
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (npl > 0) {
        SWCU_TRY(stage_energy(ctx, npl, lmask, Gmass, mass, radius, rb, vb));
        SWCU_TRY(energy_and_momentum(ctx, ctx->s_pl, GMcb, lclose, false, s));
    }
    // swiftest_util.f90:1206-1209, 1269-1283
    const double kecb = mass_cb * (vbcb[0] * vbcb[0] + vbcb[1] * vbcb[1] + vbcb[2] * vbcb[2]);
    out8[0] = 0.5 * (kecb + s[0]);
    out8[1] = -s[1] - s[2];
    out8[2] = lclose ? -s[3] : 0.0;
    out8[3] = out8[0] + 0.0 + out8[1] + out8[2];
    out8[4] = mass_cb * (rbcb[1] * vbcb[2] - rbcb[2] * vbcb[1]) + s[4];
    out8[5] = mass_cb * (rbcb[2] * vbcb[0] - rbcb[0] * vbcb[2]) + s[5];
    out8[6] = mass_cb * (rbcb[0] * vbcb[1] - rbcb[1] * vbcb[0]) + s[6];
    out8[7] = GMcb + s[7];
    return SWCU_OK;
}

// ======================================================================================================
// tier 1: triangular encounter checks, pl-tp discard, SyMBA list check
// ======================================================================================================
extern "C" int swcu_encounter_check_all_triangular_plpl(swcu_context *ctx, int32_t npl, const double *r, const double *v,
                                                        const double *renc, double dt, int64_t *nenc)
{
    SWCU_TRY(check_ctx(ctx));
    if (!nenc || npl < 0) return fail(ctx, SWCU_ERR_ARG, "tri_plpl: bad argument");
    *nenc = 0;
    ctx->enc.nenc = 0;
    ctx->enc.result = nullptr;
    if (npl == 0) return SWCU_OK;
    if (!r || !v || !renc) return fail(ctx, SWCU_ERR_ARG, "tri_plpl: null array");
    SWCU_TRY(stage_population(ctx, ctx->s_pl, npl, r, v, renc));
    return encounter_triangular(ctx, sweep_list(ctx->s_pl, 0, npl, true), nullptr, dt, nenc);
}

extern "C" int swcu_encounter_check_all_triangular_pltp(swcu_context *ctx, int32_t npl, int32_t ntp, const double *rpl,
                                                        const double *vpl, const double *rtp, const double *vtp,
                                                        const double *rencpl, double dt, int64_t *nenc)
{
    SWCU_TRY(check_ctx(ctx));
    if (!nenc || npl < 0 || ntp < 0) return fail(ctx, SWCU_ERR_ARG, "tri_pltp: bad argument");
    *nenc = 0;
    ctx->enc.nenc = 0;
    ctx->enc.result = nullptr;
    if (npl == 0 || ntp == 0) return SWCU_OK;
    if (!rpl || !vpl || !rtp || !vtp || !rencpl) return fail(ctx, SWCU_ERR_ARG, "tri_pltp: null array");
    SWCU_TRY(stage_population(ctx, ctx->s_pl, npl, rpl, vpl, rencpl));
    SWCU_TRY(stage_population(ctx, ctx->s_tp, ntp, rtp, vtp, nullptr));
    SweepList l2 = sweep_list(ctx->s_tp, 0, ntp, false);
    return encounter_triangular(ctx, sweep_list(ctx->s_pl, 0, npl, true), &l2, dt, nenc);
}

extern "C" int swcu_encounter_check_all_triangular_plplm(swcu_context *ctx, int32_t nplm, int32_t nplt, const double *rplm,
                                                         const double *vplm, const double *rplt, const double *vplt,
                                                         const double *rencm, const double *renct, double dt, int64_t *nenc)
{
    SWCU_TRY(check_ctx(ctx));
    if (!nenc || nplm < 0 || nplt < 0) return fail(ctx, SWCU_ERR_ARG, "tri_plplm: bad argument");
    *nenc = 0;
    ctx->enc.nenc = 0;
    ctx->enc.result = nullptr;
    if (nplm == 0 || nplt == 0) return SWCU_OK;
    if (!rplm || !vplm || !rplt || !vplt || !rencm || !renct) return fail(ctx, SWCU_ERR_ARG, "tri_plplm: null array");
    SWCU_TRY(stage_population(ctx, ctx->s_pl, nplm, rplm, vplm, rencm));
    SWCU_TRY(stage_population(ctx, ctx->s_tp, nplt, rplt, vplt, renct));
    SweepList l2 = sweep_list(ctx->s_tp, 0, nplt, true);
    return encounter_triangular(ctx, sweep_list(ctx->s_pl, 0, nplm, true), &l2, dt, nenc);
}

extern "C" int swcu_discard_pl_tp(swcu_context *ctx, int32_t ntp, int32_t npl, const double *rtp, const double *vtp,
                                  const int32_t *lactive, const double *rpl, const double *vpl, const double *radius,
                                  double dt, int32_t *iplanet, int32_t *ndiscard)
{
    SWCU_TRY(check_ctx(ctx));
    if (ntp < 0 || npl < 0 || !iplanet) return fail(ctx, SWCU_ERR_ARG, "discard_pl_tp: bad argument");
    if (ndiscard) *ndiscard = 0;
    if (ntp == 0) return SWCU_OK;
    if (npl == 0) {
        for (int32_t i = 0; i < ntp; ++i) iplanet[i] = 0;
        return SWCU_OK;
    }
    if (!rtp || !vtp || !rpl || !vpl || !radius) return fail(ctx, SWCU_ERR_ARG, "discard_pl_tp: null array");
    SWCU_TRY(stage_population(ctx, ctx->s_pl, npl, rpl, vpl, nullptr));
    SWCU_TRY(put_or_fill(ctx, radius, npl, ctx->s_pl.radius, 0.0));
    SWCU_TRY(stage_population(ctx, ctx->s_tp, ntp, rtp, vtp, nullptr));
    if (lactive)
        SWCU_TRY(upload_arr(ctx, lactive, sizeof(int32_t) * (size_t)ntp, ctx->s_tp.lmask));
    else
        SWCU_TRY(fill_i32(ctx, ctx->s_tp.lmask.as<int32_t>(), 1, ntp));
    int32_t nd = 0;
    SWCU_TRY(discard_pl_tp(ctx, ctx->s_tp, ctx->s_pl, ctx->s_tp.lmask.as<int32_t>(), dt, ctx->s_tp.iflag.as<int32_t>(), &nd));
    SWCU_CUDA(ctx, cudaMemcpyAsync(iplanet, ctx->s_tp.iflag.p, sizeof(int32_t) * (size_t)ntp, cudaMemcpyDeviceToHost, ctx->stream));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ndiscard) *ndiscard = nd;
    return SWCU_OK;
}

// tier 2 of the above: resident test particles against resident planets (tp%rh, tp%vh, pl%rh, pl%vh, pl%radius in HBM);
// only the count, and the planet indices when there is something to discard, cross PCIe
extern "C" int swcu_tp_discard_pl(swcu_context *ctx, double dt, int32_t *iplanet, int32_t *ndiscard)
{
    SWCU_TRY(check_ctx(ctx));
    Body &pl = ctx->pl, &tp = ctx->tp;
    if (ndiscard) *ndiscard = 0;
    if (!pl.valid || !tp.valid) return fail(ctx, SWCU_ERR_STATE, "tp_discard_pl: populations not resident");
    if (tp.n == 0) return SWCU_OK;
    SWCU_CUDA(ctx, ctx->tp_discard.ensure(sizeof(int32_t) * (size_t)tp.n));
    int32_t nd = 0;
    const int32_t *d_active = tp.has_active ? tp.lactive.as<int32_t>() : tp.lmask.as<int32_t>();
    SWCU_TRY(discard_pl_tp(ctx, tp, pl, d_active, dt, ctx->tp_discard.as<int32_t>(), &nd));
    if (ndiscard) *ndiscard = nd;
    if (iplanet && (nd > 0 || pl.n == 0)) {
        SWCU_CUDA(ctx, cudaMemcpyAsync(iplanet, ctx->tp_discard.p, sizeof(int32_t) * (size_t)tp.n, cudaMemcpyDeviceToHost,
                                       ctx->stream));
        SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else if (iplanet) {
        memset(iplanet, 0, sizeof(int32_t) * (size_t)tp.n);  // nobody is discarded: nothing to read back
    }
    return SWCU_OK;
}

extern "C" int swcu_symba_encounter_check_list(swcu_context *ctx, int64_t nenc, const int32_t *index1,
                                               const int32_t *index2, const int32_t *lencmask, int32_t n1, const double *r1,
                                               const double *v1, const double *renc1, const double *radius1, int32_t n2,
                                               const double *r2, const double *v2, const double *renc2,
                                               const double *radius2, double dt, int32_t *lencounter, int32_t *lvdotr,
                                               int64_t *nfound)
{
    SWCU_TRY(check_ctx(ctx));
    if (nenc < 0 || n1 < 0 || n2 < 0) return fail(ctx, SWCU_ERR_ARG, "symba_encounter_check_list: bad argument");
    if (nfound) *nfound = 0;
    if (nenc == 0) return SWCU_OK;  // symba_encounter_check.f90:105
    if (!index1 || !index2 || !r1 || !v1 || !renc1 || !radius1 || !lencounter || !lvdotr || n1 == 0)
        return fail(ctx, SWCU_ERR_ARG, "symba_encounter_check_list: null array");
    const bool two = (n2 > 0);
    if (two && (!r2 || !v2)) return fail(ctx, SWCU_ERR_ARG, "symba_encounter_check_list: null second list");
    const int32_t nmax2 = two ? n2 : n1;
    for (int64_t k = 0; k < nenc; ++k) {  // the reference trusts its own lists; a foreign caller gets a checked error
        if (lencmask && !lencmask[k]) continue;
        if (index1[k] < 1 || index1[k] > n1 || index2[k] < 1 || index2[k] > nmax2)
            return fail(ctx, SWCU_ERR_ARG, "symba_encounter_check_list: pair %lld = (%d,%d) out of range", (long long)k,
                        index1[k], index2[k]);
    }
    SWCU_TRY(stage_population(ctx, ctx->s_pl, n1, r1, v1, renc1));
    SWCU_TRY(put_or_fill(ctx, radius1, n1, ctx->s_pl.radius, 0.0));
    SweepList l1 = sweep_list(ctx->s_pl, 0, n1, true), l2 = l1;
    const double *d_rad2 = ctx->s_pl.radius.as<double>();
    if (two) {
        SWCU_TRY(stage_population(ctx, ctx->s_tp, n2, r2, v2, renc2));
        l2 = sweep_list(ctx->s_tp, 0, n2, renc2 != nullptr);
        d_rad2 = nullptr;
        if (radius2) {
            SWCU_TRY(put_or_fill(ctx, radius2, n2, ctx->s_tp.radius, 0.0));
            d_rad2 = ctx->s_tp.radius.as<double>();
        }
    }
    const size_t ib = sizeof(int32_t) * (size_t)nenc;
    SWCU_CUDA(ctx, ctx->istage[0].ensure(2 * ib));
    SWCU_CUDA(ctx, ctx->istage[1].ensure(3 * ib));
    int32_t *d_i1 = ctx->istage[0].as<int32_t>(), *d_i2 = d_i1 + nenc;
    int32_t *d_mask = ctx->istage[1].as<int32_t>(), *d_lenc = d_mask + nenc, *d_lvd = d_lenc + nenc;
    SWCU_CUDA(ctx, cudaMemcpyAsync(d_i1, index1, ib, cudaMemcpyHostToDevice, ctx->stream));
    SWCU_CUDA(ctx, cudaMemcpyAsync(d_i2, index2, ib, cudaMemcpyHostToDevice, ctx->stream));
    if (lencmask) SWCU_CUDA(ctx, cudaMemcpyAsync(d_mask, lencmask, ib, cudaMemcpyHostToDevice, ctx->stream));
    SWCU_CUDA(ctx, cudaMemcpyAsync(d_lvd, lvdotr, ib, cudaMemcpyHostToDevice, ctx->stream));  // kept outside the mask
    SWCU_TRY(symba_check_list(ctx, nenc, d_i1, d_i2, lencmask ? d_mask : nullptr, l1, ctx->s_pl.radius.as<double>(), l2,
                              d_rad2, dt, d_lenc, d_lvd, nfound));
    SWCU_CUDA(ctx, cudaMemcpyAsync(lencounter, d_lenc, ib, cudaMemcpyDeviceToHost, ctx->stream));
    SWCU_CUDA(ctx, cudaMemcpyAsync(lvdotr, d_lvd, ib, cudaMemcpyDeviceToHost, ctx->stream));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SWCU_OK;
}

// tier 2 of the above: the pair loop on the resident populations -- pl%rh, pl%vb, pl%renc (swcu_pl_set_renc), pl%radius and
// tp%rh, tp%vb stay on the device; a recursion level moves the pair list, the mask and the flags only.
// kind = SWCU_PL: symba_encounter_check_list_plpl (:88-140), SWCU_TP: _pltp (:163-214).
extern "C" int swcu_body_symba_encounter_check_list(swcu_context *ctx, int32_t kind, int64_t nenc, const int32_t *index1,
                                                    const int32_t *index2, const int32_t *lencmask, double dt,
                                                    int32_t *lencounter, int32_t *lvdotr, int64_t *nfound)
{
    SWCU_TRY(check_ctx(ctx));
    if (nfound) *nfound = 0;
    if (kind != SWCU_PL && kind != SWCU_TP) return fail(ctx, SWCU_ERR_ARG, "body_symba_encounter_check_list: kind");
    const bool two = (kind == SWCU_TP);
    Body &pl = ctx->pl, &tp = ctx->tp;
    if (!pl.valid || (two && !tp.valid))
        return fail(ctx, SWCU_ERR_STATE, "body_symba_encounter_check_list: population not resident");
    if (nenc < 0) return fail(ctx, SWCU_ERR_ARG, "body_symba_encounter_check_list: bad argument");
    if (nenc == 0) return SWCU_OK;  // symba_encounter_check.f90:105
    if (!index1 || !index2 || !lencounter || !lvdotr || pl.n == 0)
        return fail(ctx, SWCU_ERR_ARG, "body_symba_encounter_check_list: null array");
    const int32_t nmax2 = two ? tp.n : pl.n;
    for (int64_t k = 0; k < nenc; ++k) {
        if (lencmask && !lencmask[k]) continue;
        if (index1[k] < 1 || index1[k] > pl.n || index2[k] < 1 || index2[k] > nmax2)
            return fail(ctx, SWCU_ERR_ARG, "body_symba_encounter_check_list: pair %lld = (%d,%d) out of range", (long long)k,
                        index1[k], index2[k]);
    }
    SWCU_TRY(ensure_helio(ctx, pl));
    if (two) SWCU_TRY(ensure_helio(ctx, tp));
    auto vb_list = [](Body &b, bool with_renc) {
        SweepList l = sweep_list(b, 0, b.n, with_renc);
        l.vx = b.wx.as<double>(), l.vy = b.wy.as<double>(), l.vz = b.wz.as<double>();
        return l;
    };
    const SweepList l1 = vb_list(pl, true), l2 = two ? vb_list(tp, false) : l1;
    const size_t ib = sizeof(int32_t) * (size_t)nenc;
    SWCU_CUDA(ctx, ctx->istage[0].ensure(2 * ib));
    SWCU_CUDA(ctx, ctx->istage[1].ensure(3 * ib));
    int32_t *d_i1 = ctx->istage[0].as<int32_t>(), *d_i2 = d_i1 + nenc;
    int32_t *d_mask = ctx->istage[1].as<int32_t>(), *d_lenc = d_mask + nenc, *d_lvd = d_lenc + nenc;
    SWCU_CUDA(ctx, cudaMemcpyAsync(d_i1, index1, ib, cudaMemcpyHostToDevice, ctx->stream));
    SWCU_CUDA(ctx, cudaMemcpyAsync(d_i2, index2, ib, cudaMemcpyHostToDevice, ctx->stream));
    if (lencmask) SWCU_CUDA(ctx, cudaMemcpyAsync(d_mask, lencmask, ib, cudaMemcpyHostToDevice, ctx->stream));
    SWCU_CUDA(ctx, cudaMemcpyAsync(d_lvd, lvdotr, ib, cudaMemcpyHostToDevice, ctx->stream));  // kept outside the mask
    SWCU_TRY(symba_check_list(ctx, nenc, d_i1, d_i2, lencmask ? d_mask : nullptr, l1, pl.radius.as<double>(), l2,
                              two ? nullptr : pl.radius.as<double>(), dt, d_lenc, d_lvd, nfound));
    SWCU_CUDA(ctx, cudaMemcpyAsync(lencounter, d_lenc, ib, cudaMemcpyDeviceToHost, ctx->stream));
    SWCU_CUDA(ctx, cudaMemcpyAsync(lvdotr, d_lvd, ib, cudaMemcpyDeviceToHost, ctx->stream));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SWCU_OK;
}

// ======================================================================================================
// measurement helpers
// ======================================================================================================
extern "C" int swcu_timer_start(swcu_context *ctx)
{
    SWCU_TRY(check_ctx(ctx));
    SWCU_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    return SWCU_OK;
}

extern "C" int swcu_timer_stop(swcu_context *ctx, double *elapsed_ms)
{
    SWCU_TRY(check_ctx(ctx));
    SWCU_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    SWCU_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    SWCU_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (elapsed_ms) *elapsed_ms = ms;
    return SWCU_OK;
}

// Laps: an event pair per bracketed region, logged without synchronising; swcu_timer_laps sums them.  bench.py times
// the K steps of a run this way with the L2 flush BETWEEN the laps.
extern "C" int swcu_timer_lap_begin(swcu_context *ctx)
{
    SWCU_TRY(check_ctx(ctx));
    if (ctx->lap_used >= ctx->lap0.size()) {
        cudaEvent_t a, b;
        SWCU_CUDA(ctx, cudaEventCreate(&a));
        SWCU_CUDA(ctx, cudaEventCreate(&b));
        ctx->lap0.push_back(a);
        ctx->lap1.push_back(b);
    }
    SWCU_CUDA(ctx, cudaEventRecord(ctx->lap0[ctx->lap_used], ctx->stream));
    return SWCU_OK;
}

extern "C" int swcu_timer_lap_end(swcu_context *ctx)
{
    SWCU_TRY(check_ctx(ctx));
    if (ctx->lap_used >= ctx->lap0.size()) return fail(ctx, SWCU_ERR_STATE, "timer_lap_end without timer_lap_begin");
    SWCU_CUDA(ctx, cudaEventRecord(ctx->lap1[ctx->lap_used], ctx->stream));
    ctx->lap_used++;
    return SWCU_OK;
}

// total (and optionally every lap) in ms since the last call; synchronises the stream and clears the log
extern "C" int swcu_timer_laps(swcu_context *ctx, double *total_ms, int32_t *count, double *each_ms, int32_t each_cap)
{
    SWCU_TRY(check_ctx(ctx));
    if (!total_ms) return SWCU_ERR_ARG;
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double sum = 0.0;
    for (size_t k = 0; k < ctx->lap_used; ++k) {
        float t = 0.f;
        SWCU_CUDA(ctx, cudaEventElapsedTime(&t, ctx->lap0[k], ctx->lap1[k]));
        sum += t;
        if (each_ms && (int32_t)k < each_cap) each_ms[k] = t;
    }
    *total_ms = sum;
    if (count) *count = (int32_t)ctx->lap_used;
    ctx->lap_used = 0;
    return p2p_check_error(ctx);
}

extern "C" int swcu_enable_kernel_timing(swcu_context *ctx, int32_t on)
{
    if (!ctx) return SWCU_ERR_ARG;
    ctx->kernel_timing = on;
    if (on == 2)
        for (int f = 0; f < FAM_COUNT; ++f) ctx->fam_log_used[f] = 0;
    return SWCU_OK;
}

extern "C" int swcu_kernel_ms_accumulated(swcu_context *ctx, int32_t family, double *total_ms, int32_t *count)
{
    SWCU_TRY(check_ctx(ctx));
    if (family < 0 || family >= FAM_COUNT || !total_ms) return fail(ctx, SWCU_ERR_ARG, "kernel_ms_accumulated: bad family");
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double sum = 0.0;
    const size_t n = ctx->fam_log_used[family];
    for (size_t k = 0; k < n; ++k) {
        float t = 0.f;
        SWCU_CUDA(ctx, cudaEventElapsedTime(&t, ctx->fam_log0[family][k], ctx->fam_log1[family][k]));
        sum += t;
    }
    *total_ms = sum;
    if (count) *count = (int32_t)n;
    return SWCU_OK;
}

extern "C" int swcu_last_kernel_ms(swcu_context *ctx, int32_t family, double *ms)
{
    SWCU_TRY(check_ctx(ctx));
    if (family < 0 || family >= FAM_COUNT || !ms) return fail(ctx, SWCU_ERR_ARG, "last_kernel_ms: bad family");
    if (ctx->fam_pending[family]) {
        SWCU_CUDA(ctx, cudaEventSynchronize(ctx->fam_ev1[family]));
        float t = 0.f;
        SWCU_CUDA(ctx, cudaEventElapsedTime(&t, ctx->fam_ev0[family], ctx->fam_ev1[family]));
        ctx->fam_ms[family] = t;
        ctx->fam_pending[family] = false;
    }
    *ms = ctx->fam_ms[family];
    return SWCU_OK;
}

extern "C" int swcu_probe_fp64_peak(swcu_context *ctx, double *tflops)
{
    SWCU_TRY(check_ctx(ctx));
    if (!tflops) return SWCU_ERR_ARG;
    const int blocks = ctx->prop.multiProcessorCount * 8, threads = 256, iters = 20000;
    SWCU_CUDA(ctx, ctx->flush.ensure(sizeof(double) * (size_t)blocks * threads));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        SWCU_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->flush.as<double>(), iters, 1.0 + rep);
        SWCU_KERNEL_CHECK(ctx);
        SWCU_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        SWCU_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        SWCU_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double flops = 2.0 * 16.0 * (double)iters * (double)blocks * threads;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    *tflops = best;
    return SWCU_OK;
}

extern "C" int swcu_step_graph_replays(swcu_context *ctx, int64_t *count)
{
    SWCU_TRY(check_ctx(ctx));
    if (count) *count = ctx->helio_graph.replays;
    return SWCU_OK;
}

extern "C" int swcu_flat_redo_count(swcu_context *ctx, uint64_t *chunks)
{
    SWCU_TRY(check_ctx(ctx));
    if (!chunks) return SWCU_ERR_ARG;
    *chunks = 0;
    if (!ctx->flat_redo.p) return SWCU_OK;  // the third-law kernel has not run yet
    SWCU_CUDA(ctx, cudaMemcpyAsync(chunks, ctx->flat_redo.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SWCU_OK;
}

extern "C" int swcu_probe_hbm_copy(swcu_context *ctx, int64_t bytes, double *gbs)
{
    SWCU_TRY(check_ctx(ctx));
    if (!gbs || bytes <= 0) return SWCU_ERR_ARG;
    DevBuf a, b;
    SWCU_CUDA(ctx, a.ensure((size_t)bytes));
    SWCU_CUDA(ctx, b.ensure((size_t)bytes));
    SWCU_CUDA(ctx, cudaMemsetAsync(a.p, 1, (size_t)bytes, ctx->stream));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        SWCU_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        SWCU_CUDA(ctx, cudaMemcpyAsync(b.p, a.p, (size_t)bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        SWCU_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        SWCU_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        SWCU_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (rep > 0) best = std::max(best, 2.0 * (double)bytes / (ms * 1e-3) / 1e9);
    }
    a.release();
    b.release();
    *gbs = best;
    return SWCU_OK;
}

extern "C" int swcu_flush_l2(swcu_context *ctx)
{
    SWCU_TRY(check_ctx(ctx));
    static const int mib = env_int("SWCU_FLUSH_MIB", 160);  // 160 MiB = 168 MB = 1.33 x the 126 MB L2
    const size_t bytes = (size_t)(mib < 128 ? 128 : mib) << 20;
    const bool fresh = ctx->flush.cap < 2 * bytes + 256;
    SWCU_CUDA(ctx, ctx->flush.ensure(2 * bytes));
    if (fresh) SWCU_CUDA(ctx, cudaMemsetAsync(ctx->flush.p, 0, 2 * bytes, ctx->stream));
    double2 *w = ctx->flush.as<double2>();
    const size_t n2 = bytes / sizeof(double2);
    flush_kernel<<<ctx->prop.multiProcessorCount * 8, 256, 0, ctx->stream>>>(w, n2);  // write 1.33 x L2 ...
    SWCU_KERNEL_CHECK(ctx);
    static const int clean = env_int("SWCU_FLUSH_CLEAN", 1);
    if (clean) {  // ... then read another 1.33 x L2: the dirty lines are gone before the next kernel starts
        flush_read_kernel<<<ctx->prop.multiProcessorCount * 8, 256, 0, ctx->stream>>>(w + n2, n2, w);
        SWCU_KERNEL_CHECK(ctx);
    }
    return SWCU_OK;
}
