// kick_flat_kernels.cu -- Newton's-third-law ("flat") FP64 pl-pl gravity on sm_100a.
//
// Replaces swiftest_kick_getacch_int_all_flat_{rad,norad}_pl (reference swiftest/swiftest_kick.f90:69-162) over the
// canonical flattened upper triangle (swiftest_util.f90:1090-1130) restricted to i <= nplm
// (nplplm, symba/symba_util.f90:202), and the lmtiny branch of the triangular variants (kick.f90:189-217).
// The 8-byte-per-pair k_plpl table (40 GB at npl = 1e5) is never built: (i,j) come from tile coordinates.
//
// Design: each unordered pair is evaluated ONCE (20 FP64 instructions instead of 2 x 16 for the full-row kernel).
// Bodies are cut into blocks of 128.  A warp keeps one block resident -- every lane owns 4 "i" bodies (position, Gm,
// accumulators) in registers -- and meets another block 32 "j" bodies at a time.  The chunk is staged in the warp's
// private shared-memory tile; at step s lane l works on column body (l+s) mod 32, so all 32 lanes read different
// words (conflict free) and no two lanes ever update the same j in the same step.  The reaction on j is kept in a
// travelling accumulator that moves to the neighbouring lane after every step (3 SHFL.64), or in the shared tile
// (template switch, chosen by measurement); after 32 steps every j has met all 128 i and the chunk ends with one
// coalesced RED.ADD.F64 per component.  Block pairs are assigned cyclically (block I meets I+1 .. I+(nb-1)/2 mod nb)
// so every block owns the same amount of work; a persistent grid of warps claims runs of 2 consecutive (I,k) items
// from a global counter (dynamic scheduling; a static one-wave split proved fragile, see kick_flat_kernel).  Like the reference's OpenMP reduction(+:ahi,ahj) the summation order is not fixed (FP64 atomics);
// the full-row kernel (kick_kernels.cu) is the bitwise-reproducible variant.
#include "swcu_internal.cuh"
#include "kick_math.cuh"

#include <algorithm>
#include <cstdlib>

namespace swcu {
namespace {

constexpr int FIB = 4;          // i bodies per lane
constexpr int FT = 32 * FIB;    // bodies per block
constexpr int FWARPS = 4;       // warps (independent work units) per CTA
#ifndef FLAT_MIN_CTAS
#define FLAT_MIN_CTAS 3
#endif

// ordering of a lane's accumulator store before its neighbour's load of the same slot in the next step
// (a compiler-only fence was tried and is NOT sufficient: it produced wrong sums in blocks that take the masked path)
#define FLAT_STEP_FENCE() __syncwarp()

struct FlatArgs {
    const double *x, *y, *z, *gm, *rad;
    const double *radmax;  // device scalars: [0] max radius over all bodies, [1] max |coordinate|
    int n, nplm;
    int nb, nbm;          // blocks in total / blocks that own rows (cover [0,nplm))
    int Km, evenm;        // cyclic half-range among the owner blocks, and whether nbm is even
    long long total_items;
    long long item0, item1;  // the run of items this rank owns (multi-GPU: pair slices), [0,total) on one GPU
    int quantum;             // items a warp claims per visit to the work counter (coarse phase)
    // graded schedule: claims [ph_q[k], ph_q[k+1]) take ph_sz[k] consecutive 32-column chunks each, starting at chunk
    // ph_u[k] of this launch's run; sizes shrink towards the end of the run (whole items ... one chunk)
    long long ph_q[6], ph_u[6];
    int ph_sz[5];
    unsigned long long *counter;  // work counter (dynamic scheduling); in peer-memory mode ONE counter shared by all GPUs
    unsigned long long counter_base;  // value of the counter at which this launch's first quantum sits
    int system_scope;             // 1: the counter lives in (possibly remote) peer memory -> system-scope atomics
    long long first_warp, total_warps;  // this launch's first global warp id and the warps of all sharers together
    double *fx, *fy, *fz; // zero-initialised accumulation target
    unsigned long long *trace;  // development aid (SWCU_FLAT_TRACE): per warp {start, end} %globaltimer, items done
};

// every array has the same 16-byte stride so one byte offset addresses all four
struct __align__(16) WarpTile {
    double2 xy[32];
    double2 zg[32];   // z, Gm
    double2 axy[32];  // reaction accumulators (ACC_SMEM variant)
    double2 az[32];   // .x used
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

// Rare path: the pairs of one 32-body chunk that the seeded evaluation skipped (r^2 == 0, denormal, > FLT_MAX or not
// safely outside the radii), with the reference's IEEE expression irij3 = 1/(r2*sqrt(r2)) (kick.f90:435), the exact
// radius test (kick.f90:106-107) and all index masks.
template <bool RAD>
__device__ __noinline__ void redo_chunk(const FlatArgs &a, int jbase, bool diag, int lane, const double (&xi)[FIB],
                                        const double (&yi)[FIB], const double (&zi)[FIB], const double (&gmi)[FIB],
                                        const unsigned (&thr)[FIB], const unsigned (&span)[FIB], const int (&idx_i)[FIB],
                                        double (&axi)[FIB], double (&ayi)[FIB], double (&azi)[FIB])
{
    for (int s = 0; s < 32; ++s) {
        const int jcur = jbase + ((lane + s) & 31);
        if (jcur >= a.n) continue;
        const double xj = a.x[jcur], yj = a.y[jcur], zj = a.z[jcur], gmj = a.gm[jcur];
#pragma unroll
        for (int b = 0; b < FIB; ++b) {
            const double dx = xj - xi[b], dy = yj - yi[b], dz = zj - zi[b];
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            if (seed_ok(r2, thr[b], span[b])) continue;  // already done by the fast path
            if (idx_i[b] == jcur || idx_i[b] >= a.n) continue;
            if (!((idx_i[b] < a.nplm) || (jcur < a.nplm))) continue;
            if (RAD) {
                const double rl = a.rad[idx_i[b]] + a.rad[jcur];
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

// One step of block_pair: lane l meets column body `slot`; everything in the warp tile is addressed with the single byte
// offset off = 16*slot.  The column body of the NEXT step is fetched first so its LDS latency hides behind the math.
template <bool RAD, bool CHECKED, bool ACC_SMEM>
__device__ __forceinline__ void flat_step(const FlatArgs &a, char *wb, unsigned off, unsigned off_next, const double2 xy,
                                          const double2 zg, double2 &nxy, double2 &nzg, int jbase, bool diag, int src,
                                          const double (&xi)[FIB], const double (&yi)[FIB], const double (&zi)[FIB],
                                          const double (&gmi)[FIB], const unsigned (&thr)[FIB],
                                          const unsigned (&span)[FIB], const int (&idx_i)[FIB], double (&axi)[FIB],
                                          double (&ayi)[FIB], double (&azi)[FIB], double &ajx, double &ajy, double &ajz,
                                          unsigned &hymin)
{
    nxy = *reinterpret_cast<const double2 *>(wb + off_next);
    nzg = *reinterpret_cast<const double2 *>(wb + 512 + off_next);
    if (ACC_SMEM) {
        const double2 t2 = *reinterpret_cast<const double2 *>(wb + 1024 + off);
        ajx = t2.x;
        ajy = t2.y;
        ajz = *reinterpret_cast<const double *>(wb + 1536 + off);
    }
    const int jcur = jbase + (int)(off >> 4);
    unsigned hy[FIB];
#pragma unroll
    for (int b = 0; b < FIB; ++b) {
        const double dx = xy.x - xi[b];
        const double dy = xy.y - yi[b];
        const double dz = zg.x - zi[b];
        const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
        double y3;
        if (CHECKED) {
            const bool m = (idx_i[b] != jcur) && (jcur < a.n) && (idx_i[b] < a.n) &&
                           ((idx_i[b] < a.nplm) || (jcur < a.nplm));
            // a masked pair never contributes (always-failing test: the result is exactly zero) and must not send
            // the chunk to redo_chunk either: every chunk of a diagonal block holds a self pair per lane
            y3 = rcube_seeded<true>(r2, thr[b], m ? span[b] : 0u, hy[b]);
            hy[b] = m ? hy[b] : 0xffffffffu;
        } else {
            y3 = rcube_seeded<false>(r2, thr[b], span[b], hy[b]);  // the caller guarantees |coordinates| < 2^62
        }
        const double fj = zg.y * y3;  // acts on i
        double fi = gmi[b] * y3;      // acts on j
        if (CHECKED) fi = diag ? 0.0 : fi;  // a diagonal block visits (i,j) and (j,i)
        axi[b] = fma(fj, dx, axi[b]);
        ayi[b] = fma(fj, dy, ayi[b]);
        azi[b] = fma(fj, dz, azi[b]);
        ajx = fma(-fi, dx, ajx);
        ajy = fma(-fi, dy, ajy);
        ajz = fma(-fi, dz, ajz);
    }
    static_assert(FIB == 4, "the seed-word minimum below is written for 4 row bodies per lane");
    hymin = __vimin3_u32(__vimin3_u32(hymin, hy[0], hy[1]), hy[2], hy[3]);  // 2 x VIMNMX3 for 4 pairs
    if (ACC_SMEM) {
        *reinterpret_cast<double2 *>(wb + 1024 + off) = make_double2(ajx, ajy);
        *reinterpret_cast<double *>(wb + 1536 + off) = ajz;
        // The neighbour lane reads this slot in the next step.  The warp is converged here (no divergent branch inside
        // the chunk loop) and a warp's shared-memory accesses are performed in program order, so a compiler-level fence
        // is sufficient; the full __syncwarp() stays at the chunk boundaries.
        FLAT_STEP_FENCE();
    } else {
        ajx = __shfl_sync(0xffffffffu, ajx, src);
        ajy = __shfl_sync(0xffffffffu, ajy, src);
        ajz = __shfl_sync(0xffffffffu, ajz, src);
    }
}

// One block pair: block I resident in registers, block J streamed through the warp's shared tile 32 bodies at a time.
// CHECKED adds the index masks needed by diagonal blocks, the ragged last block and blocks that straddle nplm.
template <bool RAD, bool CHECKED, bool ACC_SMEM>
__device__ __forceinline__ void block_pair(const FlatArgs &a, WarpTile &w, int J, bool diag, int lane,
                                           const double (&xi)[FIB], const double (&yi)[FIB], const double (&zi)[FIB],
                                           const double (&gmi)[FIB], const unsigned (&thr)[FIB],
                                           const unsigned (&span)[FIB], const int (&idx_i)[FIB], double (&axi)[FIB],
                                           double (&ayi)[FIB], double (&azi)[FIB], int c0 = 0, int c1 = FIB)
{
    const int src = (lane + 1) & 31;
    // prefetch the first chunk
    int jc = min(J * FT + c0 * 32 + lane, a.n - 1);
    double nx = a.x[jc], ny = a.y[jc], nz = a.z[jc], ng = a.gm[jc];
#pragma unroll 1
    for (int c = c0; c < c1; ++c) {
        const int jbase = J * FT + c * 32;
        const int jidx = jbase + lane;
        __syncwarp();  // every lane is done reading the previous chunk
        w.xy[lane] = make_double2(nx, ny);
        w.zg[lane] = make_double2(nz, ng);
        if (ACC_SMEM) {
            w.axy[lane] = make_double2(0.0, 0.0);
            w.az[lane] = make_double2(0.0, 0.0);
        }
        __syncwarp();
        if (c + 1 < c1) {  // next chunk's loads fly while this one is computed
            jc = min(jbase + 32 + lane, a.n - 1);
            nx = a.x[jc];
            ny = a.y[jc];
            nz = a.z[jc];
            ng = a.gm[jc];
        }
        double ajx = 0.0, ajy = 0.0, ajz = 0.0;
        unsigned hymin = 0xffffffffu;  // running minimum of the seed words: 0 <=> some pair was rejected
        char *wb = reinterpret_cast<char *>(&w);
        // two steps per iteration with two named register sets (A, B): no register copies for the prefetch
        unsigned off = (unsigned)lane << 4;
        double2 axy_ = *reinterpret_cast<const double2 *>(wb + off);
        double2 azg_ = *reinterpret_cast<const double2 *>(wb + 512 + off);
        double2 bxy_, bzg_;
#pragma unroll 1
        for (int s = 0; s < 32; s += 2) {
            const unsigned off1 = (off + 16u) & 0x1f0u, off2 = (off + 32u) & 0x1f0u;
            flat_step<RAD, CHECKED, ACC_SMEM>(a, wb, off, off1, axy_, azg_, bxy_, bzg_, jbase, diag, src, xi, yi, zi, gmi,
                                              thr, span, idx_i, axi, ayi, azi, ajx, ajy, ajz, hymin);
            flat_step<RAD, CHECKED, ACC_SMEM>(a, wb, off1, off2, bxy_, bzg_, axy_, azg_, jbase, diag, src, xi, yi, zi, gmi,
                                              thr, span, idx_i, axi, ayi, azi, ajx, ajy, ajz, hymin);
            off = off2;
        }
        const bool bad = (hymin == 0u);
        if (ACC_SMEM) {
            const double2 t2 = w.axy[lane];
            ajx = t2.x;
            ajy = t2.y;
            ajz = w.az[lane].x;
        }
        // the accumulator of column body jbase+lane is now in this lane
        if (!diag && jidx < a.n) {
            atomicAdd(a.fx + jidx, ajx);
            atomicAdd(a.fy + jidx, ajy);
            atomicAdd(a.fz + jidx, ajz);
        }
        if (__builtin_expect(bad, 0))
            redo_chunk<RAD>(a, jbase, diag, lane, xi, yi, zi, gmi, thr, span, idx_i, axi, ayi, azi);
    }
}

template <bool RAD, bool ACC_SMEM>
__global__ void __launch_bounds__(32 * FWARPS, FLAT_MIN_CTAS) kick_flat_kernel(const FlatArgs a)
{
    __shared__ WarpTile tiles[FWARPS];
    const int lane = threadIdx.x & 31;
    WarpTile &w = tiles[threadIdx.x >> 5];
    double xi[FIB], yi[FIB], zi[FIB], gmi[FIB], axi[FIB], ayi[FIB], azi[FIB];
    unsigned thr[FIB], span[FIB];
    int idx_i[FIB];
    int Icur = -1;
    const double radmax = RAD ? a.radmax[0] : 0.0;
    const bool coords_safe = a.radmax[1] < COORD_SAFE_MAX;  // else every block pair takes the fully checked path
    unsigned long long *tr = a.trace ? a.trace + 4 * ((size_t)blockIdx.x * FWARPS + (threadIdx.x >> 5)) : nullptr;
    unsigned long long ndone = 0;
    if (tr && lane == 0) {
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        tr[0] = t0;
    }

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

    // Dynamic scheduling: a warp claims `quantum` consecutive items at a time from a global counter.  (A static split
    // sized to exactly one resident wave was measured to be fragile: when the tail of the previous kernel still
    // occupies an SM at launch, one CTA is left over, runs alone after the others and doubles the kernel time.)
    // The first quantum of every warp is static (its global warp id): no claim storm on the counter at launch, which
    // with eight GPUs sharing one counter cost ~0.1 ms.  Claimed values then map to quanta total_warps, total_warps+1, ...
    // The claim for the NEXT quantum is issued before the current one is processed, so the round trip of the atomic
    // (a couple of microseconds over NVLink when the counter lives on another GPU) hides behind ~15 us of arithmetic.
    auto claim = [&]() -> unsigned long long {
        unsigned long long c = 0;
        if (lane == 0)
            c = (a.system_scope ? atomicAdd_system(a.counter, 1ull) : atomicAdd(a.counter, 1ull)) - a.counter_base +
                (unsigned long long)a.total_warps;
        return c;  // valid in lane 0 only; broadcast when it is consumed
    };
    // Claim q covers chunk units [u0, u1) of this launch's run (4 units = the 32-column chunks of one block pair).  The
    // bulk of the run goes out in whole items; towards the end the claims shrink (4, 2, 1 chunks) so that the kernel
    // does not end on a ragged edge: three warps share an SMSP and the scheduler does not serve them equally (traced:
    // 361..987 chunks per warp over one launch), so a block pair claimed late by a slow warp can take 100-200 us.
    // Measured idle at the end of the launch: 76 us per warp without the grading; that is 7 % of the kernel once eight
    // GPUs share the work.
    const long long unit0 = a.item0 * FIB, unit1 = a.item1 * FIB;
    auto range = [&](unsigned long long qq, long long &u0, long long &u1) -> bool {
        if (qq >= (unsigned long long)a.ph_q[5]) return false;
        int k = 0;
#pragma unroll
        for (int j = 1; j < 5; ++j) k += (qq >= (unsigned long long)a.ph_q[j]) ? 1 : 0;
        u0 = unit0 + a.ph_u[k] + (long long)(qq - (unsigned long long)a.ph_q[k]) * a.ph_sz[k];
        u1 = min(u0 + a.ph_sz[k], unit0 + a.ph_u[k + 1]);  // the last claim of a phase may be short
        return u0 < unit1;
    };
    unsigned long long q = (unsigned long long)(a.first_warp + (long long)blockIdx.x * FWARPS + (threadIdx.x >> 5));
    unsigned long long qnext = claim();
    for (;;) {
        long long u, u_end;
        if (!range(q, u, u_end)) {
            if (q >= (unsigned long long)a.total_warps) break;  // the failed claim that ends this warp
            // (a warp whose static quantum lies beyond the range falls through to its prefetched claim)
            q = __shfl_sync(0xffffffffu, qnext, 0);
            if (!range(q, u, u_end)) break;
            qnext = claim();
        }
        while (u < u_end) {
            const long long t = u / FIB;
            const int c0 = (int)(u - t * FIB);
            const int c1 = (int)min((long long)FIB, c0 + (u_end - u));
            u += c1 - c0;
            ndone += c1 - c0;
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
                    double rl2 = 0.0;
                    if (RAD) {
                        const double rl = a.rad[ic] + radmax;
                        rl2 = rl * rl;
                    }
                    seed_threshold(rl2, thr[b], span[b]);
                    axi[b] = ayi[b] = azi[b] = 0.0;
                }
            }
            const bool checked = !coords_safe || diag || I == a.nb - 1 || J == a.nb - 1 ||
                                 (a.nplm < a.n && (I == a.nbm - 1 || J == a.nbm - 1));
            if (checked)
                block_pair<RAD, true, ACC_SMEM>(a, w, J, diag, lane, xi, yi, zi, gmi, thr, span, idx_i, axi, ayi, azi, c0, c1);
            else
                block_pair<RAD, false, ACC_SMEM>(a, w, J, diag, lane, xi, yi, zi, gmi, thr, span, idx_i, axi, ayi, azi, c0, c1);
        }
        q = __shfl_sync(0xffffffffu, qnext, 0);
        long long dummy0, dummy1;
        if (!range(q, dummy0, dummy1)) break;  // that was this warp's one failed claim
        qnext = claim();
    }
    if (tr && lane == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        tr[1] = t1;
        tr[2] = ndone;
    }
    flush();
    if (tr && lane == 0) {
        unsigned long long t2;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2));
        tr[3] = t2;
    }
}

// out[0] = max |radius[i]| (0 when radius == nullptr), out[1] = max over bodies of max(|x|,|y|,|z|); the bit patterns of
// non-negative doubles order like unsigned integers, so one integer atomicMax per warp does the reduction
__global__ void max_radius_coord_kernel(const double *radius, const double *x, const double *y, const double *z, int n,
                                        unsigned long long *out)
{
    unsigned long long mr = 0ull, mc = 0ull;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (radius != nullptr) mr = max(mr, (unsigned long long)__double_as_longlong(fabs(radius[i])));
        const double c = fmax(fabs(x[i]), fmax(fabs(y[i]), fabs(z[i])));
        mc = max(mc, (unsigned long long)__double_as_longlong(c));  // NaN orders above every finite value: unsafe
    }
    for (int o = 16; o > 0; o >>= 1) {
        mr = max(mr, __shfl_xor_sync(0xffffffffu, mr, o));
        mc = max(mc, __shfl_xor_sync(0xffffffffu, mc, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out, mr);
        atomicMax(out + 1, mc);
    }
}

}  // namespace

// device scalars {max radius, max |coordinate|} of a population into two 8-byte slots of ctx->scratch64
int max_radius(swcu_context *ctx, const double *radius, const double *x, const double *y, const double *z, int n,
               int slot, const double **d_out)
{
    SWCU_CUDA(ctx, ctx->scratch64.ensure(128));
    unsigned long long *d = ctx->scratch64.as<unsigned long long>() + slot;  // slot 1: columns / whole population, 4: rows
    SWCU_CUDA(ctx, cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), ctx->stream));
    if (n > 0) {
        max_radius_coord_kernel<<<std::min(cdiv(n, 256), 512), 256, 0, ctx->stream>>>(radius, x, y, z, n, d);
        SWCU_KERNEL_CHECK(ctx);
    }
    *d_out = reinterpret_cast<const double *>(d);
    return SWCU_OK;
}

// Flat (third-law) variant over the canonical flattened triangle restricted to i < nplm_rows (0-based).
// With several ranks every rank evaluates an equal run of the (I,k) items, the partial accelerations are summed with
// one allreduce of 3*npl doubles and added to ah on every rank (all ranks then hold the same ah for all bodies).
int kick_pl_flat(swcu_context *ctx, Body &pl, bool lrad, int nplm_rows, bool reduce)
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
    SWCU_TRY(max_radius(ctx, a.rad, a.x, a.y, a.z, n, 1, &a.radmax));
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
    if (!reduce) {  // peer-memory mode: accumulate into the exported buffer, the caller reduces across ranks
        if (!ctx->p2p.ready || ctx->p2p.stride != stride) return fail(ctx, SWCU_ERR_STATE, "kick_pl_flat: p2p buffers not set up for npl=%d", n);
        a.fx = ctx->p2p.F.as<double>();
    } else {
        SWCU_CUDA(ctx, ctx->partial.ensure(sizeof(double) * 3 * stride));
        a.fx = ctx->partial.as<double>();
    }
    a.fy = a.fx + stride;
    a.fz = a.fy + stride;
    SWCU_CUDA(ctx, cudaMemsetAsync(a.fx, 0, sizeof(double) * 3 * stride, ctx->stream));

    // measured on B200 at npl = 1e5 (profiles/r01_fp64_pipe.md): shared-memory accumulators beat the SHFL-travelling ones
    static const bool acc_smem = getenv("SWCU_FLAT_ACC_SMEM") ? atoi(getenv("SWCU_FLAT_ACC_SMEM")) != 0 : true;
    void (*kern)(const FlatArgs) = lrad ? (acc_smem ? kick_flat_kernel<true, true> : kick_flat_kernel<true, false>)
                                        : (acc_smem ? kick_flat_kernel<false, true> : kick_flat_kernel<false, false>);
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * FWARPS, 0);
    occ = std::max(1, occ);
    a.item0 = 0;
    a.item1 = total;
    // NCCL mode: balanced consecutive runs of items per rank.  Peer-memory mode: all ranks claim quanta from ONE counter
    // in rank 0's memory (system-scope atomics over NVLink), so the GPUs finish together whatever their speed.
    const int nr = reduce ? ctx->nranks : 1, rk = reduce ? ctx->rank : 0;
    if (nr > 1) {  // balanced consecutive runs, like swcu_partition
        const long long q = total / nr, r = total % nr;
        a.item0 = rk * q + std::min<long long>(rk, r);
        a.item1 = a.item0 + q + (rk < r ? 1 : 0);
    }
    const long long mine = a.item1 - a.item0;
    const long long warps_max = (long long)ctx->prop.multiProcessorCount * occ * FWARPS;
    // Items per coarse claim.  Every claim switches the resident row block (12 REDs + 20 loads + thresholds, with only
    // three warps per SMSP to hide it): measured at npl = 1e5 with the graded end of the schedule, 1 item 7.90 ms,
    // 2 items 7.73, 4 items 7.67, 6 items 7.65 (but slower at 7e4 bodies).  Fewer items per warp -> smaller claims.
    const int sharers = reduce ? 1 : ctx->p2p.nranks;
    a.quantum = 4;
    if (mine < 32 * warps_max * sharers) a.quantum = 2;
    if (mine < 16 * warps_max * sharers) a.quantum = 1;
    if (ctx->tune_nsplit > 0) a.quantum = ctx->tune_nsplit;
    SWCU_CUDA(ctx, ctx->scratch64.ensure(128));
    a.counter = ctx->scratch64.as<unsigned long long>() + 3;
    a.counter_base = 0;
    a.system_scope = 0;
    // persistent grid: every SM filled to its occupancy (or fewer CTAs when there is little work).  The last
    // `fine` chunks per warp (of all sharers) are handed out one 32-column chunk at a time (see the kernel's `range`).
    // chunks per warp (of all sharers) handed out in 1-, 2-, 4- and 8-chunk claims at the end of the run.  A claim of s
    // chunks can take ~2 s chunk-times on a warp the scheduler disfavours, so everything finer than s has to last that
    // long: 2 s chunks per warp for size s.
    static int fine[4] = {2, 4, 8, 16};
    static bool fine_read = false;
    if (!fine_read) {
        fine_read = true;
        if (const char *e = getenv("SWCU_FLAT_FINE")) sscanf(e, "%d,%d,%d,%d", &fine[0], &fine[1], &fine[2], &fine[3]);
    }
    // Eight GPUs on one counter: the single-chunk phase would ask for ~28k claims within ~25 us (> 1 G claims/s to one
    // address over NVLink; 0.6 G/s was measured harmless at four GPUs), so it is left out there unless overridden.
    const bool skip_single = (sharers >= 8) && !getenv("SWCU_FLAT_FINE");
    auto split_quanta = [&](long long items, long long warps_all) -> long long {
        const long long U = items * FIB, G = (long long)a.quantum * FIB;
        long long left = U;
        long long nq[5] = {0, 0, 0, 0, 0};  // phases in run order: G, 8, 4, 2, 1 chunks
        const int sz[5] = {(int)G, 8, 4, 2, 1};
        for (int k = 4; k >= 1; --k) {  // carve the fine phases off the end of the run
            if (sz[k] >= G) continue;
            if (k == 4 && skip_single) continue;
            long long want = std::min<long long>(left, warps_all * fine[4 - k]);
            want -= want % sz[k];
            nq[k] = want / sz[k];
            left -= want;
        }
        // the coarse phase takes what is left; a remainder that is not a multiple of G goes out as one more claim
        nq[0] = (left + G - 1) / G;
        long long qacc = 0, uacc = 0;
        for (int k = 0; k < 5; ++k) {
            a.ph_q[k] = qacc;
            a.ph_u[k] = uacc;
            a.ph_sz[k] = sz[k];
            qacc += nq[k];
            uacc += (k == 0) ? left : nq[k] * sz[k];
        }
        a.ph_q[5] = qacc;
        a.ph_u[5] = uacc;
        return qacc;
    };
    const long long nquanta = split_quanta(mine, warps_max * sharers);
    long long units = std::max<long long>(1, std::min<long long>(warps_max, nquanta));
    if (!reduce && ctx->p2p.nranks > 1) {
        // shared counter: never reset (a fast rank must not see a stale zero); every launch consumes exactly
        // max(nquanta - total_warps, 0) successful claims + one failed claim per warp of every rank, so all ranks agree
        // on the base of each epoch
        units = warps_max;  // identical grids on all ranks keep that sum predictable
        const unsigned long long per_epoch = (unsigned long long)std::max<long long>(nquanta, units * ctx->p2p.nranks);
        a.counter = reinterpret_cast<unsigned long long *>(ctx->p2p.peer[0][7]) + 40;
        a.counter_base = ctx->p2p.epoch * per_epoch;  // epoch counts completed steps (incremented after the kick)
        a.system_scope = 1;
        a.first_warp = units * ctx->p2p.rank;
        a.total_warps = units * ctx->p2p.nranks;
    } else {
        a.first_warp = 0;
        a.total_warps = units;
        SWCU_CUDA(ctx, cudaMemsetAsync(a.counter, 0, sizeof(unsigned long long), ctx->stream));
    }
    const int grid = cdiv(units, FWARPS);
    a.trace = nullptr;
    static const char *trace_path = getenv("SWCU_FLAT_TRACE");
    DevBuf &trace_buf = ctx->flat_trace;
    if (trace_path) {
        SWCU_CUDA(ctx, trace_buf.ensure(sizeof(unsigned long long) * 4 * (size_t)grid * FWARPS));
        SWCU_CUDA(ctx, cudaMemsetAsync(trace_buf.p, 0, sizeof(unsigned long long) * 4 * (size_t)grid * FWARPS, ctx->stream));
        a.trace = trace_buf.as<unsigned long long>();
    }
    kern<<<grid, 32 * FWARPS, 0, ctx->stream>>>(a);
    SWCU_KERNEL_CHECK(ctx);
    if (trace_path) {  // development aid: dump the per-warp timeline of this launch (overwrites the file)
        std::vector<unsigned long long> h((size_t)4 * grid * FWARPS);
        SWCU_CUDA(ctx, cudaMemcpyAsync(h.data(), trace_buf.p, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (FILE *f = fopen(trace_path, "wb")) {
            fwrite(h.data(), sizeof(unsigned long long), h.size(), f);
            fclose(f);
        }
    }
    if (!reduce) return SWCU_OK;
    if (ctx->nranks > 1) SWCU_TRY(comm_allreduce_sum(ctx, a.fx, 3 * stride));
    return axpy3(ctx, 1.0, a.fx, a.fy, a.fz, pl.ax.as<double>(), pl.ay.as<double>(), pl.az.as<double>(), nullptr, n);
}

}  // namespace swcu
