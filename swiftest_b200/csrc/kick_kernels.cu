// kick_kernels.cu -- FP64 pairwise gravity on sm_100a.
//
// Replaces the OpenMP loops of swiftest_kick.f90 (reference, relative to /root/reference/src):
//   full rows      swiftest_kick_getacch_int_all_tri_{rad,norad}_pl   kick.f90:165-271, 274-371
//   pl -> tp       swiftest_kick_getacch_int_all_tp                    kick.f90:374-415
//   pair list      swiftest_kick_getacch_int_all_flat_rad_pl on k_plpl kick.f90:69-115 (SyMBA list, symba_kick.f90:61-70)
//
// Design (FP64-FMA bound, no tensor cores -- the pair force is not a contraction):
//   * one kernel, "rows x columns": each thread keeps IB row bodies (position, radius, 3 accumulators) in
//     registers; column bodies are staged as SoA tiles in shared memory by 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier, 2 stages) and read with warp-broadcast LDS.128;
//   * 1/r^3 comes from an FP32 MUFU.RSQ seed on float(r^2) refined by one third-order Newton step in FP64
//     (relative error ~1e-16); operands outside the normal FP32 range (or r^2 == 0) take the IEEE
//     1/(r2*sqrt(r2)) path, so the result is valid for every double input;
//   * a thread walks its columns in ascending j exactly like the reference's full-row loop; with one column
//     split the accumulators start from the incoming acc, reproducing the reference summation order;
//   * columns are split over gridDim.y to fill 148 SMs evenly; partial sums are reduced in a fixed order
//     (deterministic, no atomics).
#include "swcu_internal.cuh"
#include "kick_math.cuh"

#include <algorithm>
#include <cstdlib>

namespace swcu {
namespace {

constexpr int KNT = 128;     // threads per CTA
constexpr int KTJ = 256;     // column bodies per shared-memory tile
constexpr int KSTAGES = 2;   // TMA pipeline depth

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct KickArgs {
    const double *xi, *yi, *zi, *radi;
    int row0, row1;
    const double *xj, *yj, *zj, *gmj, *radj;
    int col0, col1;
    const double *radmax;     // device scalars of the columns: [0] max radius, [1] max |coordinate|
    const double *rowmax;     // same for the rows ([1] used)
    int diag;                 // rows and columns index the same population
    const int32_t *lmask;
    double *ax, *ay, *az;     // final accumulators (used directly when gridDim.y == 1)
    double *px, *py, *pz;     // partial sums [gridDim.y][pstride]
    int64_t pstride;
};

// One column body against the IB row bodies of this thread: 17 FP64 + ~9 other instructions per evaluation, one
// basic block.  Pairs the seeded path cannot take (r^2 == 0 on the diagonal, r^2 outside the FP32 exponent range, or
// not safely outside the sum of radii -- kick_math.cuh) contribute exactly zero and raise `bad`; the caller redoes
// those few once per tile (redo_tile).
template <int IB, bool UPPER>
__device__ __forceinline__ void eval_column(const double (&xi)[IB], const double (&yi)[IB], const double (&zi)[IB],
                                            const unsigned (&thr)[IB], const unsigned (&span)[IB], double xj, double yj,
                                            double zj, double gmj, double (&ax)[IB], double (&ay)[IB], double (&az)[IB],
                                            unsigned &hymin)
{
#pragma unroll
    for (int b = 0; b < IB; ++b) {
        const double dx = xj - xi[b];
        const double dy = yj - yi[b];
        const double dz = zj - zi[b];
        const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
        unsigned hy;
        const double y3 = rcube_seeded<UPPER>(r2, thr[b], span[b], hy);
        if (hy == 0u) hymin = 0u;  // (a predicate OR measured 4 % faster here than an integer min)
        const double f = gmj * y3;
        ax[b] = fma(f, dx, ax[b]);
        ay[b] = fma(f, dy, ay[b]);
        az[b] = fma(f, dz, az[b]);
    }
}

// Rare path: the evaluations of one tile that the seeded path skipped, with the reference's own IEEE expression
// fac = Gm_j / (r2*sqrt(r2)) and exact radius test rji2 > (radius_i + radius_j)**2 (kick.f90:227-237).
template <int IB>
__device__ __noinline__ void redo_tile(const KickArgs &a, const double *sx, const double *sy, const double *sz,
                                       const double *sg, int cnt, int jbase, const double (&xi)[IB],
                                       const double (&yi)[IB], const double (&zi)[IB], const unsigned (&thr)[IB],
                                       const unsigned (&span)[IB], const int (&rowid)[IB], double (&ax)[IB],
                                       double (&ay)[IB], double (&az)[IB])
{
    for (int jj = 0; jj < cnt; ++jj) {
#pragma unroll
        for (int b = 0; b < IB; ++b) {
            const double dx = sx[jj] - xi[b], dy = sy[jj] - yi[b], dz = sz[jj] - zi[b];
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            if (seed_ok(r2, thr[b], span[b])) continue;  // already added by the fast path
            if (a.diag && rowid[b] == jbase + jj) continue;
            if (a.radi != nullptr) {
                const double rl = a.radi[min(rowid[b], a.row1 - 1)] + a.radj[jbase + jj];
                if (!(r2 > rl * rl)) continue;
            }
            const double fac = sg[jj] / (r2 * sqrt(r2));
            ax[b] = fma(fac, dx, ax[b]);
            ay[b] = fma(fac, dy, ay[b]);
            az[b] = fma(fac, dz, az[b]);
        }
    }
}

template <int IB>
__global__ void __launch_bounds__(KNT) kick_rows_kernel(const __grid_constant__ KickArgs a)
{
    __shared__ __align__(128) double sm[KSTAGES][4][KTJ];
    __shared__ __align__(8) uint64_t full[KSTAGES];

    const int tid = threadIdx.x;
    const int ntile_total = (a.col1 - a.col0 + KTJ - 1) / KTJ;
    const int tiles_per = (ntile_total + (int)gridDim.y - 1) / (int)gridDim.y;
    const int t0 = (int)blockIdx.y * tiles_per;
    const int t1 = min(ntile_total, t0 + tiles_per);
    const bool direct = (gridDim.y == 1);
    const double radmax = (a.radi != nullptr) ? a.radmax[0] : 0.0;
    // rows and columns all below 2^62 in magnitude: r^2 cannot overflow a float and the range test loses its upper end
    const bool coords_safe = a.radmax[1] < COORD_SAFE_MAX && a.rowmax[1] < COORD_SAFE_MAX;

    double xi[IB], yi[IB], zi[IB], ax[IB], ay[IB], az[IB];
    unsigned thr[IB], span[IB];
    int rowid[IB];
#pragma unroll
    for (int b = 0; b < IB; ++b) {
        rowid[b] = a.row0 + (int)blockIdx.x * (KNT * IB) + b * KNT + tid;
        const int ic = min(rowid[b], a.row1 - 1);
        xi[b] = a.xi[ic];
        yi[b] = a.yi[ic];
        zi[b] = a.zi[ic];
        double rl2 = 0.0;
        if (a.radi != nullptr) {
            const double rl = a.radi[ic] + radmax;
            rl2 = rl * rl;
        }
        seed_threshold(rl2, thr[b], span[b]);
        // one column split: start from the incoming acceleration so the sum runs in the reference's order
        ax[b] = direct ? a.ax[ic] : 0.0;
        ay[b] = direct ? a.ay[ic] : 0.0;
        az[b] = direct ? a.az[ic] : 0.0;
    }

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < KSTAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int t, int s) {
        const int j = a.col0 + t * KTJ;
        const int cnt = min(KTJ, a.col1 - j);
        const uint32_t bytes = (uint32_t)((cnt * 8 + 15) & ~15);
        mbar_expect_tx(&full[s], bytes * 4u);
        bulk_g2s(&sm[s][0][0], a.xj + j, bytes, &full[s]);
        bulk_g2s(&sm[s][1][0], a.yj + j, bytes, &full[s]);
        bulk_g2s(&sm[s][2][0], a.zj + j, bytes, &full[s]);
        bulk_g2s(&sm[s][3][0], a.gmj + j, bytes, &full[s]);
    };

    if (tid == 0 && t0 < t1) issue(t0, 0);

    for (int t = t0; t < t1; ++t) {
        const int it = t - t0;
        const int s = it & 1;
        // stage s^1 was consumed in iteration it-1; every thread has passed the barrier that ended it
        if (tid == 0 && t + 1 < t1) issue(t + 1, s ^ 1);
        mbar_wait(&full[s], (uint32_t)((it >> 1) & 1));

        const int jbase = a.col0 + t * KTJ;
        const int cnt = min(KTJ, a.col1 - jbase);
        const double *sx = sm[s][0], *sy = sm[s][1], *sz = sm[s][2], *sg = sm[s][3];
        unsigned bad = 0xffffffffu;  // 0 <=> some evaluation of this tile was rejected by the fast-path test
        int jj = 0;
        if (coords_safe) {
#pragma unroll 1
            for (; jj + 1 < cnt; jj += 2) {
                const double2 x2 = *reinterpret_cast<const double2 *>(sx + jj);
                const double2 y2 = *reinterpret_cast<const double2 *>(sy + jj);
                const double2 z2 = *reinterpret_cast<const double2 *>(sz + jj);
                const double2 g2 = *reinterpret_cast<const double2 *>(sg + jj);
                eval_column<IB, false>(xi, yi, zi, thr, span, x2.x, y2.x, z2.x, g2.x, ax, ay, az, bad);
                eval_column<IB, false>(xi, yi, zi, thr, span, x2.y, y2.y, z2.y, g2.y, ax, ay, az, bad);
            }
        }
        for (; jj < cnt; ++jj) eval_column<IB, true>(xi, yi, zi, thr, span, sx[jj], sy[jj], sz[jj], sg[jj], ax, ay, az, bad);
        if (__builtin_expect(bad == 0u, 0)) redo_tile<IB>(a, sx, sy, sz, sg, cnt, jbase, xi, yi, zi, thr, span, rowid, ax, ay, az);
        __syncthreads();
    }

#pragma unroll
    for (int b = 0; b < IB; ++b) {
        const int i = rowid[b];
        if (i >= a.row1) continue;
        if (a.lmask != nullptr && a.lmask[i] == 0) continue;
        if (direct) {
            a.ax[i] = ax[b];
            a.ay[i] = ay[b];
            a.az[i] = az[b];
        } else {
            const int64_t o = (int64_t)blockIdx.y * a.pstride + (i - a.row0);
            a.px[o] = ax[b];
            a.py[o] = ay[b];
            a.pz[o] = az[b];
        }
    }
}

// acc[row0+i] += sum_s partial[s][i], s ascending (fixed order => bitwise reproducible)
__global__ void kick_reduce_partials_kernel(const double *px, const double *py, const double *pz, int64_t pstride,
                                            int nsplit, int row0, int nrows, const int32_t *lmask, double *ax,
                                            double *ay, double *az)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    if (lmask != nullptr && lmask[row0 + i] == 0) return;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int s = 0; s < nsplit; ++s) {
        sx += px[(int64_t)s * pstride + i];
        sy += py[(int64_t)s * pstride + i];
        sz += pz[(int64_t)s * pstride + i];
    }
    ax[row0 + i] += sx;
    ay[row0 + i] += sy;
    az[row0 + i] += sz;
}

// Explicit pair list (kick.f90:99-109 on a k_plpl table): one thread per pair, the reference's own IEEE
// expression, native FP64 atomics into a zeroed accumulator.  Used for the short SyMBA encounter list.
template <bool RAD>
__global__ void kick_pair_list_kernel(int64_t npairs, const int32_t *i1, const int32_t *i2, const double *x,
                                      const double *y, const double *z, const double *gm, const double *rad,
                                      double *ex, double *ey, double *ez)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npairs) return;
    const int i = i1[k] - 1, j = i2[k] - 1;
    const double rx = x[j] - x[i], ry = y[j] - y[i], rz = z[j] - z[i];
    const double rji2 = fma(rz, rz, fma(ry, ry, rx * rx));
    if (RAD) {
        const double rl = rad[i] + rad[j];
        if (!(rji2 > rl * rl)) return;
    }
    const double irij3 = 1.0 / (rji2 * sqrt(rji2));
    const double faci = gm[i] * irij3, facj = gm[j] * irij3;
    atomicAdd(&ex[i], facj * rx);
    atomicAdd(&ey[i], facj * ry);
    atomicAdd(&ez[i], facj * rz);
    atomicAdd(&ex[j], -(faci * rx));
    atomicAdd(&ey[j], -(faci * ry));
    atomicAdd(&ez[j], -(faci * rz));
}

__global__ void axpy3_kernel(double alpha, const double *x0, const double *x1, const double *x2, double *y0, double *y1,
                             double *y2, const int32_t *lmask, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (lmask != nullptr && lmask[i] == 0) return;
    y0[i] = fma(alpha, x0[i], y0[i]);
    y1[i] = fma(alpha, x1[i], y1[i]);
    y2[i] = fma(alpha, x2[i], y2[i]);
}

// pl -> tp with a handful of planets (swiftest_kick_getacch_int_all_tp, kick.f90:394-412): pure streaming over the
// test particles, HBM bound (76 B per tp).  The planets (<= TP_SMALL_NPL) sit in shared memory; one thread per tp, two
// tps in flight per thread for memory-level parallelism; no tile pipeline, no barrier after the prologue.
constexpr int TP_SMALL_NPL = 64;
__global__ void __launch_bounds__(256) kick_tp_small_kernel(int ntp, int npl, const double *__restrict__ xt,
                                                            const double *__restrict__ yt, const double *__restrict__ zt,
                                                            const double *__restrict__ xp, const double *__restrict__ yp,
                                                            const double *__restrict__ zp, const double *__restrict__ gp,
                                                            const int32_t *__restrict__ lmask, double *__restrict__ ax,
                                                            double *__restrict__ ay, double *__restrict__ az)
{
    __shared__ double4 pl[TP_SMALL_NPL];
    if (threadIdx.x < npl) pl[threadIdx.x] = make_double4(xp[threadIdx.x], yp[threadIdx.x], zp[threadIdx.x], gp[threadIdx.x]);
    __syncthreads();
    unsigned thr, span;
    seed_threshold(0.0, thr, span);
    const int i0 = blockIdx.x * 512 + threadIdx.x;
    int idx[2] = {i0, i0 + 256};
    double x[2], y[2], z[2], a0[2], a1[2], a2[2];
    bool on[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        on[q] = idx[q] < ntp && lmask[idx[q]] != 0;
        const int ic = min(idx[q], ntp - 1);
        x[q] = xt[ic];
        y[q] = yt[ic];
        z[q] = zt[ic];
        a0[q] = ax[ic];
        a1[q] = ay[ic];
        a2[q] = az[ic];
    }
    unsigned hymin = 0xffffffffu;
    for (int j = 0; j < npl; ++j) {
        const double4 p = pl[j];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double dx = p.x - x[q], dy = p.y - y[q], dz = p.z - z[q];
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            unsigned hy;
            const double y3 = rcube_seeded(r2, thr, span, hy);
            hymin = min(hymin, hy);
            const double f = p.w * y3;
            a0[q] = fma(f, dx, a0[q]);
            a1[q] = fma(f, dy, a1[q]);
            a2[q] = fma(f, dz, a2[q]);
        }
    }
    if (__builtin_expect(hymin == 0u, 0)) {  // a tp on top of a planet or coordinates outside the FP32 exponent range
        for (int j = 0; j < npl; ++j) {
            const double4 p = pl[j];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const double dx = p.x - x[q], dy = p.y - y[q], dz = p.z - z[q];
                const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                if (seed_ok(r2, thr, span)) continue;
                const double f = p.w / (r2 * sqrt(r2));  // kick.f90:464
                a0[q] = fma(f, dx, a0[q]);
                a1[q] = fma(f, dy, a1[q]);
                a2[q] = fma(f, dz, a2[q]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (on[q]) {
            ax[idx[q]] = a0[q];
            ay[idx[q]] = a1[q];
            az[idx[q]] = a2[q];
        }
    }
}

int choose_nsplit(int nrowblocks, int ntiles, int nsm)
{
    int best = 1;
    double bestcost = 1e300;
    const int maxs = std::min(ntiles, 96);
    for (int ns = 1; ns <= maxs; ++ns) {
        const int64_t total = (int64_t)nrowblocks * ns;
        const int tiles_per = (ntiles + ns - 1) / ns;
        double cost = (double)((total + nsm - 1) / nsm) * tiles_per;
        if (total < 2 * (int64_t)nsm) cost *= 1.0 + 0.25 * (double)(2 * nsm - total) / (2.0 * nsm);  // few warps/SM
        cost += 0.02 * ns;  // partial-sum traffic, and prefer the smaller split on ties
        if (cost < bestcost - 1e-9) {
            bestcost = cost;
            best = ns;
        }
    }
    return best;
}

template <int IB>
int launch_rows(swcu_context *ctx, const KickProblem &p, int nsplit_override)
{
    const int nrows = p.row1 - p.row0, ncols = p.col1 - p.col0;
    const int nrb = cdiv(nrows, KNT * IB);
    const int ntiles = cdiv(ncols, KTJ);
    int ns = nsplit_override > 0 ? std::min(nsplit_override, ntiles) : choose_nsplit(nrb, ntiles, ctx->prop.multiProcessorCount);
    // every split must own at least one tile
    while (ns > 1 && (int64_t)(ns - 1) * cdiv(ntiles, ns) >= ntiles) --ns;

    KickArgs a;
    a.xi = p.xi; a.yi = p.yi; a.zi = p.zi; a.radi = p.radi;
    a.row0 = p.row0; a.row1 = p.row1;
    a.xj = p.xj; a.yj = p.yj; a.zj = p.zj; a.gmj = p.gmj; a.radj = p.radj;
    a.col0 = p.col0; a.col1 = p.col1;
    a.lmask = p.lmask;
    a.ax = p.ax; a.ay = p.ay; a.az = p.az;
    a.px = a.py = a.pz = nullptr;
    a.pstride = 0;
    if (ns > 1) {
        const int64_t stride = ((int64_t)nrows + 31) & ~int64_t(31);
        SWCU_CUDA(ctx, ctx->partial.ensure(sizeof(double) * 3 * (size_t)stride * ns));
        a.px = ctx->partial.as<double>();
        a.py = a.px + stride * ns;
        a.pz = a.py + stride * ns;
        a.pstride = stride;
    }
    a.diag = p.diag ? 1 : 0;
    // {max radius, max |coordinate|} of the columns and, when they are a different population (pl -> tp), of the rows
    SWCU_TRY(max_radius(ctx, p.radj ? p.radj + p.col0 : nullptr, p.xj + p.col0, p.yj + p.col0, p.zj + p.col0, ncols, 1,
                        &a.radmax));
    a.rowmax = a.radmax;
    if (p.xi != p.xj)
        SWCU_TRY(max_radius(ctx, nullptr, p.xi + p.row0, p.yi + p.row0, p.zi + p.row0, nrows, 4, &a.rowmax));
    const dim3 grid(nrb, ns), block(KNT);
    kick_rows_kernel<IB><<<grid, block, 0, ctx->stream>>>(a);
    SWCU_KERNEL_CHECK(ctx);
    if (ns > 1) {
        kick_reduce_partials_kernel<<<cdiv(nrows, 256), 256, 0, ctx->stream>>>(a.px, a.py, a.pz, a.pstride, ns, p.row0,
                                                                               nrows, p.lmask, p.ax, p.ay, p.az);
        SWCU_KERNEL_CHECK(ctx);
    }
    return SWCU_OK;
}

}  // namespace

int kick_rows(swcu_context *ctx, const KickProblem &p, int family)
{
    const int nrows = p.row1 - p.row0, ncols = p.col1 - p.col0;
    if (nrows <= 0 || ncols <= 0) return SWCU_OK;
    if (p.col0 % 2 != 0) return fail(ctx, SWCU_ERR_ARG, "kick_rows: column range must start at an even index");
    FamTimer ft(ctx, family);
    if (!p.diag && p.radi == nullptr && ncols <= TP_SMALL_NPL && p.lmask != nullptr && p.row0 == 0 && ctx->tune_ib == 0) {
        kick_tp_small_kernel<<<cdiv(nrows, 512), 256, 0, ctx->stream>>>(nrows, ncols, p.xi, p.yi, p.zi, p.xj + p.col0,
                                                                       p.yj + p.col0, p.zj + p.col0, p.gmj + p.col0,
                                                                       p.lmask, p.ax, p.ay, p.az);
        SWCU_KERNEL_CHECK(ctx);
        return SWCU_OK;
    }
    int ib = ctx->tune_ib;
    if (ib != 1 && ib != 2 && ib != 4) {
        // rows per thread, by measurement (scripts/tri_ib_scan.sh, profiles/r02_crossover.md): one row per thread is the
        // fastest up to ~4e4 rows (more CTAs, better tail), two rows per thread win by 1 % from ~6e4 rows; four never do
        ib = (ncols > 64 && nrows >= 60000) ? 2 : 1;
    }
    switch (ib) {
        case 4: return launch_rows<4>(ctx, p, ctx->tune_nsplit);
        case 2: return launch_rows<2>(ctx, p, ctx->tune_nsplit);
        default: return launch_rows<1>(ctx, p, ctx->tune_nsplit);
    }
}

// swiftest_kick_getacch_int_all_tri_*_pl restricted to rows [row0,row1): rows below nplm see every column,
// rows from nplm up see the first nplm columns (kick.f90:221-239 and :245-263).  The lmtiny branch (:189-217)
// evaluates the same interactions in another order, so it maps onto the same two launches.
int kick_pl_tri(swcu_context *ctx, Body &pl, bool lrad, int row0, int row1)
{
    const int npl = pl.n, nplm = pl.nplm;
    KickProblem p;
    p.xi = p.xj = pl.rx.as<double>();
    p.yi = p.yj = pl.ry.as<double>();
    p.zi = p.zj = pl.rz.as<double>();
    p.radi = p.radj = lrad ? pl.radius.as<double>() : nullptr;
    p.gmj = pl.Gm.as<double>();
    p.diag = true;
    p.lmask = nullptr;
    p.ax = pl.ax.as<double>();
    p.ay = pl.ay.as<double>();
    p.az = pl.az.as<double>();
    // block 1: rows [row0, min(row1,nplm)) x columns [0,npl)
    p.row0 = row0;
    p.row1 = std::min(row1, nplm);
    p.col0 = 0;
    p.col1 = npl;
    if (p.row1 > p.row0) SWCU_TRY(kick_rows(ctx, p, FAM_PLPL));
    // block 2: rows [max(row0,nplm), row1) x columns [0,nplm)
    p.row0 = std::max(row0, nplm);
    p.row1 = row1;
    p.col0 = 0;
    p.col1 = nplm;
    if (p.row1 > p.row0 && nplm > 0) SWCU_TRY(kick_rows(ctx, p, FAM_PLPL));
    return SWCU_OK;
}

int kick_pair_list(swcu_context *ctx, const Body &pl, bool lrad, int64_t nenc, const int32_t *d_i1, const int32_t *d_i2,
                   double *ex, double *ey, double *ez)
{
    if (nenc <= 0) return SWCU_OK;
    const int nb = cdiv(nenc, 256);
    if (lrad)
        kick_pair_list_kernel<true><<<nb, 256, 0, ctx->stream>>>(nenc, d_i1, d_i2, pl.rx.as<double>(), pl.ry.as<double>(),
                                                               pl.rz.as<double>(), pl.Gm.as<double>(),
                                                               pl.radius.as<double>(), ex, ey, ez);
    else
        kick_pair_list_kernel<false><<<nb, 256, 0, ctx->stream>>>(nenc, d_i1, d_i2, pl.rx.as<double>(), pl.ry.as<double>(),
                                                                pl.rz.as<double>(), pl.Gm.as<double>(), nullptr, ex, ey,
                                                                ez);
    SWCU_KERNEL_CHECK(ctx);
    return SWCU_OK;
}

int axpy3(swcu_context *ctx, double alpha, const double *x0, const double *x1, const double *x2, double *y0, double *y1,
          double *y2, const int32_t *lmask, int n)
{
    if (n <= 0) return SWCU_OK;
    axpy3_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(alpha, x0, x1, x2, y0, y1, y2, lmask, n);
    SWCU_KERNEL_CHECK(ctx);
    return SWCU_OK;
}

}  // namespace swcu
