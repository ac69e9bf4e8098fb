// kick_flat_kernels.cu -- Newton's-third-law ("flat") FP64 pl-pl gravity on sm_100a.
//
// Replaces swiftest_kick_getacch_int_all_flat_{rad,norad}_pl (reference swiftest/swiftest_kick.f90:69-162) over the
// canonical flattened upper triangle (swiftest_util.f90:1090-1130) restricted to i <= nplm
// (nplplm, symba/symba_util.f90:202), and the lmtiny branch of the triangular variants (kick.f90:189-217).
// The 8-byte-per-pair k_plpl table (40 GB at npl = 1e5) is never built: (i,j) come from tile coordinates.
//
// Design: each unordered pair is evaluated ONCE (26 FP64-pipe instructions instead of 2 x 22 for the full-row
// kernel).  Bodies are cut into blocks of 128.  A warp keeps one block resident -- every lane owns 4 "i" bodies
// (position, Gm, radius, accumulators) in registers -- and meets another block 32 bodies at a time: each lane also
// holds ONE travelling "j" body with its own accumulator, and the 32 travelling bodies rotate around the warp with
// SHFL so that after 32 steps every j has met all 128 i.  The j-side sums then leave the warp with 3 coalesced FP64
// RED.ADD per body.  (A broadcast + butterfly-reduction variant was measured 30% slower: 10 SHFL + 3.8 extra DADD per
// warp-level pair evaluation against 4 SHFL here.)  Block pairs are assigned cyclically (block I meets I+1 .. I+(nb-1)/2 mod nb) so every block
// owns the same amount of work, and the (I,k) items are cut into equal consecutive runs, one per resident warp.
// Like the reference's OpenMP reduction(+:ahi,ahj) the summation order is not fixed (FP64 atomics); the full-row
// kernel (kick_kernels.cu) is the bitwise-reproducible variant.
#include "swcu_internal.cuh"
#include "kick_math.cuh"

#include <algorithm>

namespace swcu {
namespace {

constexpr int FIB = 4;          // i bodies per lane
constexpr int FT = 32 * FIB;    // bodies per block
constexpr int FWARPS = 4;       // warps (independent work units) per CTA

struct FlatArgs {
    const double *x, *y, *z, *gm, *rad;
    int n, nplm;
    int nb, nbm;          // blocks in total / blocks that own rows (cover [0,nplm))
    int Km, evenm;        // cyclic half-range among the owner blocks, and whether nbm is even
    long long total_items, items_per_unit;
    double *fx, *fy, *fz; // zero-initialised accumulation target
};

// number of items owned by block I: diagonal + cyclic partners among owner blocks + all non-owner blocks
__device__ __host__ __forceinline__ int items_of(int I, int nb, int nbm, int Km, int evenm)
{
    return 1 + Km + ((evenm && I < nbm / 2) ? 1 : 0) + (nb - nbm);
}

__device__ __forceinline__ void item_decode(long long t, const FlatArgs &a, int &I, int &J, bool &diag)
{
    const int base = 1 + a.Km + (a.nb - a.nbm);
    int k;
    if (a.evenm) {
        const long long big = (long long)(base + 1) * (a.nbm / 2);
        if (t < big) {
            I = (int)(t / (base + 1));
            k = (int)(t % (base + 1));
        } else {
            t -= big;
            I = a.nbm / 2 + (int)(t / base);
            k = (int)(t % base);
        }
    } else {
        I = (int)(t / base);
        k = (int)(t % base);
    }
    const int kmI = a.Km + ((a.evenm && I < a.nbm / 2) ? 1 : 0);
    diag = (k == 0);
    if (k == 0)
        J = I;
    else if (k <= kmI)
        J = (I + k) % a.nbm;
    else
        J = a.nbm + (k - 1 - kmI);
}

// Rare path of block_pair: the pairs of one 32-body chunk that the FP32-seeded evaluation skipped (r^2 == 0, denormal
// or > FLT_MAX), with the reference's IEEE expression irij3 = 1/(r2*sqrt(r2)) (kick.f90:435); all index masks applied.
template <bool RAD>
__device__ __noinline__ void redo_chunk(const FlatArgs &a, int jbase, bool diag, int lane, const double (&xi)[FIB],
                                        const double (&yi)[FIB], const double (&zi)[FIB], const double (&gmi)[FIB],
                                        const double (&radi)[FIB], const int (&idx_i)[FIB], double (&axi)[FIB],
                                        double (&ayi)[FIB], double (&azi)[FIB])
{
    for (int s = 0; s < 32; ++s) {
        const int jcur = jbase + ((lane + s) & 31);
        if (jcur >= a.n) continue;
        const double xj = a.x[jcur], yj = a.y[jcur], zj = a.z[jcur], gmj = a.gm[jcur];
        const double radj = RAD ? a.rad[jcur] : 0.0;
#pragma unroll
        for (int b = 0; b < FIB; ++b) {
            const double dx = xj - xi[b], dy = yj - yi[b], dz = zj - zi[b];
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            bool ok;
            (void)rsqrt_newton(r2, ok);
            if (ok || idx_i[b] == jcur || idx_i[b] >= a.n) continue;
            if (!((idx_i[b] < a.nplm) || (jcur < a.nplm))) continue;
            if (RAD) {
                const double rl = radi[b] + radj;
                if (!(r2 > rl * rl)) continue;
            }
            const double irij3 = 1.0 / (r2 * sqrt(r2));
            const double fj = gmj * irij3, fi = gmi[b] * irij3;
            axi[b] = fma(fj, dx, axi[b]);
            ayi[b] = fma(fj, dy, ayi[b]);
            azi[b] = fma(fj, dz, azi[b]);
            if (!diag) {
                atomicAdd(a.fx + jcur, -(fi * dx));
                atomicAdd(a.fy + jcur, -(fi * dy));
                atomicAdd(a.fz + jcur, -(fi * dz));
            }
        }
    }
}

// One block pair: block I resident in registers (4 bodies per lane), block J streamed 32 bodies at a time.
// The 32 column bodies of a chunk are loaded coalesced, one per lane, and then TRAVEL: every lane evaluates its 4
// pairs against the body it currently holds, adds the reaction to that body's travelling accumulator, and passes
// body + accumulator to its neighbour (SHFL).  After 32 steps every column body has met all 128 row bodies and is
// back in the lane that loaded it, so the chunk ends with one coalesced RED.ADD.F64 per component.
// CHECKED adds the index masks needed by diagonal blocks, the ragged last block and blocks that straddle nplm.
template <bool RAD, bool CHECKED>
__device__ __forceinline__ void block_pair(const FlatArgs &a, int J, bool diag, int lane, const double (&xi)[FIB],
                                           const double (&yi)[FIB], const double (&zi)[FIB], const double (&gmi)[FIB],
                                           const double (&radi)[FIB], const int (&idx_i)[FIB], double (&axi)[FIB],
                                           double (&ayi)[FIB], double (&azi)[FIB])
{
    const int src = (lane + 1) & 31;
#pragma unroll 1
    for (int c = 0; c < FIB; ++c) {
        const int jbase = J * FT + c * 32;
        const int jidx = jbase + lane;
        const int jc = min(jidx, a.n - 1);
        double xj = a.x[jc], yj = a.y[jc], zj = a.z[jc], gmj = a.gm[jc];
        double radj = RAD ? a.rad[jc] : 0.0;
        double ajx = 0.0, ajy = 0.0, ajz = 0.0;
        bool bad = false;
#pragma unroll 2
        for (int s = 0; s < 32; ++s) {
            const int jcur = jbase + ((lane + s) & 31);  // index of the body this lane currently holds
            double dx[FIB], dy[FIB], dz[FIB], r2[FIB], y3[FIB];
            bool okb[FIB];
#pragma unroll
            for (int b = 0; b < FIB; ++b) {
                dx[b] = xj - xi[b];
                dy[b] = yj - yi[b];
                dz[b] = zj - zi[b];
                r2[b] = fma(dz[b], dz[b], fma(dy[b], dy[b], dx[b] * dx[b]));
                const double y = rsqrt_newton(r2[b], okb[b]);
                const double y2 = y * y;
                y3[b] = y * y2;
                bad = bad || !okb[b];
            }
            // the travelling position can move on as soon as the differences are formed
            const double xn = __shfl_sync(0xffffffffu, xj, src);
            const double yn = __shfl_sync(0xffffffffu, yj, src);
            const double zn = __shfl_sync(0xffffffffu, zj, src);
#pragma unroll
            for (int b = 0; b < FIB; ++b) {
                bool use = okb[b];
                if (RAD) {
                    const double rl = radi[b] + radj;
                    use = use && (r2[b] > rl * rl);
                }
                if (CHECKED) {
                    use = use && (idx_i[b] != jcur) && (jcur < a.n) && (idx_i[b] < a.n) &&
                          ((idx_i[b] < a.nplm) || (jcur < a.nplm));
                }
                const double w = use ? y3[b] : 0.0;
                const double fj = gmj * w;          // acts on i
                double fi = gmi[b] * w;             // acts on j
                if (CHECKED) fi = diag ? 0.0 : fi;  // a diagonal block visits (i,j) and (j,i)
                axi[b] = fma(fj, dx[b], axi[b]);
                ayi[b] = fma(fj, dy[b], ayi[b]);
                azi[b] = fma(fj, dz[b], azi[b]);
                ajx = fma(-fi, dx[b], ajx);
                ajy = fma(-fi, dy[b], ajy);
                ajz = fma(-fi, dz[b], ajz);
            }
            xj = xn;
            yj = yn;
            zj = zn;
            gmj = __shfl_sync(0xffffffffu, gmj, src);
            if (RAD) radj = __shfl_sync(0xffffffffu, radj, src);
            ajx = __shfl_sync(0xffffffffu, ajx, src);
            ajy = __shfl_sync(0xffffffffu, ajy, src);
            ajz = __shfl_sync(0xffffffffu, ajz, src);
        }
        // 32 rotations: every travelling body is back in the lane that loaded it
        if (!diag && jidx < a.n) {
            atomicAdd(a.fx + jidx, ajx);
            atomicAdd(a.fy + jidx, ajy);
            atomicAdd(a.fz + jidx, ajz);
        }
        // rare: pairs whose r^2 could not use the FP32 seed (skipped above) are added with the IEEE expression
        if (__builtin_expect(bad, 0)) redo_chunk<RAD>(a, jbase, diag, lane, xi, yi, zi, gmi, radi, idx_i, axi, ayi, azi);
    }
}

template <bool RAD>
__global__ void __launch_bounds__(32 * FWARPS) kick_flat_kernel(const FlatArgs a)
{
    const int lane = threadIdx.x & 31;
    const long long unit = (long long)blockIdx.x * FWARPS + (threadIdx.x >> 5);
    long long t = unit * a.items_per_unit;
    const long long t_end = min(a.total_items, t + a.items_per_unit);
    if (t >= t_end) return;

    double xi[FIB], yi[FIB], zi[FIB], gmi[FIB], radi[FIB], axi[FIB], ayi[FIB], azi[FIB];
    int idx_i[FIB];
    int Icur = -1;

    auto flush = [&]() {
        if (Icur < 0) return;
#pragma unroll
        for (int b = 0; b < FIB; ++b) {
            if (idx_i[b] < a.n) {
                atomicAdd(a.fx + idx_i[b], axi[b]);
                atomicAdd(a.fy + idx_i[b], ayi[b]);
                atomicAdd(a.fz + idx_i[b], azi[b]);
            }
        }
    };

    for (; t < t_end; ++t) {
        int I, J;
        bool diag;
        item_decode(t, a, I, J, diag);
        if (I != Icur) {
            flush();
            Icur = I;
#pragma unroll
            for (int b = 0; b < FIB; ++b) {
                idx_i[b] = I * FT + b * 32 + lane;
                const int ic = min(idx_i[b], a.n - 1);
                xi[b] = a.x[ic];
                yi[b] = a.y[ic];
                zi[b] = a.z[ic];
                gmi[b] = a.gm[ic];
                radi[b] = RAD ? a.rad[ic] : 0.0;
                axi[b] = ayi[b] = azi[b] = 0.0;
            }
        }
        const bool checked = diag || I == a.nb - 1 || J == a.nb - 1 || (a.nplm < a.n && (I == a.nbm - 1 || J == a.nbm - 1));
        if (checked)
            block_pair<RAD, true>(a, J, diag, lane, xi, yi, zi, gmi, radi, idx_i, axi, ayi, azi);
        else
            block_pair<RAD, false>(a, J, diag, lane, xi, yi, zi, gmi, radi, idx_i, axi, ayi, azi);
    }
    flush();
}

}  // namespace

// Flat (third-law) variant over the canonical flattened triangle restricted to i < nplm_rows (0-based).
int kick_pl_flat(swcu_context *ctx, Body &pl, bool lrad, int nplm_rows)
{
    const int n = pl.n;
    if (n <= 1 || nplm_rows <= 0) return SWCU_OK;
    FamTimer ft(ctx, FAM_PLPL);
    FlatArgs a;
    a.x = pl.rx.as<double>();
    a.y = pl.ry.as<double>();
    a.z = pl.rz.as<double>();
    a.gm = pl.Gm.as<double>();
    a.rad = lrad ? pl.radius.as<double>() : nullptr;
    a.n = n;
    a.nplm = std::min(nplm_rows, n);
    a.nb = cdiv(n, FT);
    a.nbm = cdiv(a.nplm, FT);
    a.Km = (a.nbm - 1) / 2;
    a.evenm = (a.nbm % 2 == 0) ? 1 : 0;
    long long total = 0;
    for (int I = 0; I < a.nbm; ++I) total += items_of(I, a.nb, a.nbm, a.Km, a.evenm);
    a.total_items = total;

    // zeroed accumulation target, then acc += F
    const size_t stride = ((size_t)n + 31) & ~size_t(31);
    SWCU_CUDA(ctx, ctx->partial.ensure(sizeof(double) * 3 * stride));
    a.fx = ctx->partial.as<double>();
    a.fy = a.fx + stride;
    a.fz = a.fy + stride;
    SWCU_CUDA(ctx, cudaMemsetAsync(a.fx, 0, sizeof(double) * 3 * stride, ctx->stream));

    int occ = 1;
    if (lrad)
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kick_flat_kernel<true>, 32 * FWARPS, 0);
    else
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kick_flat_kernel<false>, 32 * FWARPS, 0);
    occ = std::max(1, occ);
    const long long max_units = (long long)ctx->prop.multiProcessorCount * occ * FWARPS;  // one resident wave
    long long units = std::min(max_units, total);
    if (ctx->tune_nsplit > 0) units = std::min<long long>(total, (long long)ctx->tune_nsplit * FWARPS);
    a.items_per_unit = (total + units - 1) / units;
    units = (total + a.items_per_unit - 1) / a.items_per_unit;
    const int grid = cdiv(units, FWARPS);
    if (lrad)
        kick_flat_kernel<true><<<grid, 32 * FWARPS, 0, ctx->stream>>>(a);
    else
        kick_flat_kernel<false><<<grid, 32 * FWARPS, 0, ctx->stream>>>(a);
    SWCU_KERNEL_CHECK(ctx);
    return axpy3(ctx, 1.0, a.fx, a.fy, a.fz, pl.ax.as<double>(), pl.ay.as<double>(), pl.az.as<double>(), nullptr, n);
}

}  // namespace swcu
