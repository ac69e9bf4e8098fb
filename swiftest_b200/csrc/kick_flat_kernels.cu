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
// private shared-memory tile; at step s lane l works on the column body in slot l+s (the 32 bodies are stored twice, so
// there is no wrap-around and every tile address is "lane base + immediate"): all 32 lanes read different words
// (conflict free) and no two lanes ever update the same j in the same step.  The reaction on j is accumulated in the
// tile; after 32 steps every j has met all 128 i and the chunk ends with one coalesced RED.ADD.F64 per component.
//
// Pair arithmetic: r^-3 from a MUFU.RSQ64H seed (one instruction from the high word of r^2 to the high word of an FP64
// seed) refined in six FP64 instructions (kick_math.cuh).  The fast path carries NO per-pair test: the step loop only
// keeps the running minimum of the high words of r^2 (one 3-input integer min per two pairs); when a chunk ends with
// that minimum below the block pair's threshold (first high word safely outside (max radius of I + max radius of J)^2
// and inside the range the refinement handles), the chunk is ROLLED BACK -- the row accumulators are restored from a
// snapshot taken at its start, the column accumulators are dropped -- and recomputed with the reference's IEEE
// expression and exact radius test (redo_chunk).  Hot loop: 20 FP64 + 3.3 other instructions per pair.
//
// Block pairs are assigned cyclically (block I meets I+1 .. I+(nb-1)/2 mod nb) so every block owns the same amount of
// work; a persistent grid of warps claims runs of consecutive (I,k) items from a global counter (dynamic scheduling; a
// static one-wave split proved fragile, see kick_flat_kernel).  Column bodies travel global -> staging area (cp.async,
// one chunk ahead, across block pairs) -> tile.  Like the reference's OpenMP reduction(+:ahi,ahj) the summation order is
// not fixed (FP64 atomics); the full-row kernel (kick_kernels.cu) is the bitwise-reproducible variant.
#include "swcu_internal.cuh"
#include "kick_math.cuh"

#include <algorithm>
#include <cstdlib>

namespace swcu {
namespace {

constexpr int FIB = 4;          // i bodies per lane
constexpr int FT = 32 * FIB;    // bodies per block
constexpr int FWARPS = 4;       // warps (independent work units) per CTA
// steps per iteration of the step loop (all tile offsets become immediates) and CTAs per SM are template parameters of
// the kernel: the combination in use was chosen by measurement (profiles/r02_kick_flat.md)

struct FlatArgs {
    const double *x, *y, *z, *gm, *rad;
    const double *coordmax;  // device scalar: max |coordinate| of this launch's positions (set by flat_prologue_kernel)
    const double *blockrad;  // max radius of every block of FT bodies (radius-checked variant)
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
    unsigned long long *counter;  // work counter of this GPU (dynamic scheduling), zeroed before every launch
    long long total_warps;        // warps of this launch: claims 0..total_warps-1 are static (the warp's own id)
    double *fx, *fy, *fz; // zero-initialised accumulation target
    unsigned long long *redo_count;  // chunks that went through the exact path (diagnostic, may be null)
    unsigned long long *trace;  // development aid (SWCU_FLAT_TRACE): per warp {start, end} %globaltimer, items done
};

// A warp's private tile.  The 32 column bodies of a chunk are stored TWICE (slots k and k+32) and the reaction
// accumulators have 64 slots as well: at step s lane l works on slot l+s (no wrap-around), so every tile address of the
// unrolled step loop is "lane base + immediate" and the loop carries one pointer.  Slots k and k+32 of the accumulators
// are two partial sums of column body k, added when the chunk ends.  Every array has a 16-byte stride.
struct __align__(16) WarpTile {
    double2 xy[64];
    double2 zg[64];       // z, Gm
    double2 axy[64];      // reaction accumulators
    double2 az[64];       // .x used
    double2 save[6][32];  // the lane's 12 row accumulators at the start of the chunk (restored if the chunk is redone)
    double stage[4][32];  // x, y, z, Gm of the NEXT chunk's column bodies, filled by cp.async while this chunk is computed
};
constexpr unsigned T_ZG = 1024, T_AXY = 2048, T_AZ = 3072;

// number of items owned by block I: diagonal + cyclic partners among owner blocks + all non-owner blocks
__device__ __host__ __forceinline__ int items_of(int I, int nb, int nbm, int Km, int evenm)
{
    return 1 + Km + ((evenm && I < nbm / 2) ? 1 : 0) + (nb - nbm);
}

// (32-bit arithmetic: the launcher refuses populations whose item count does not fit an int)
__device__ __forceinline__ void item_decode(int t, const FlatArgs &a, int &I, int &J, bool &diag)
{
    const int base = 1 + a.Km + (a.nb - a.nbm);
    int k;
    if (a.evenm) {
        const int big = (base + 1) * (a.nbm / 2);
        if (t < big) {
            I = t / (base + 1);
            k = t - I * (base + 1);
        } else {
            t -= big;
            const int d = t / base;
            I = a.nbm / 2 + d;
            k = t - d * base;
        }
    } else {
        I = t / base;
        k = t - I * base;
    }
    const int kmI = a.Km + ((a.evenm && I < a.nbm / 2) ? 1 : 0);
    diag = (k == 0);
    if (k == 0)
        J = I;
    else if (k <= kmI) {
        J = I + k;
        J -= (J >= a.nbm) ? a.nbm : 0;
    } else
        J = a.nbm + (k - 1 - kmI);
}

// Rare path: one whole 32-column chunk of a block pair with the reference's IEEE expression
// irij3 = 1/(r2*sqrt(r2)) (kick.f90:435), the exact radius test (kick.f90:106-107) and all index masks.  Taken when the
// seeded evaluation of the chunk met a pair it may not handle (r^2 not safely outside the radii, zero, tiny or not
// finite); the seeded results of that chunk are discarded.  Reads everything from global memory by index and adds
// straight into the accumulation target, so it shares no registers with the caller (no stack frame in the hot kernel).
// (Scalar arguments only: a reference to the kernel's parameter block would force a copy of it into local memory.)
template <bool RAD>
__device__ __noinline__ void redo_chunk(const double *__restrict__ x, const double *__restrict__ y,
                                        const double *__restrict__ z, const double *__restrict__ gm,
                                        const double *__restrict__ rad, double *fx, double *fy, double *fz, int n,
                                        int nplm, int I, int jbase, bool diag, int lane)
{
#pragma unroll 1
    for (int b = 0; b < FIB; ++b) {
        const int i = I * FT + b * 32 + lane;
        if (i >= n) continue;
        const double xi = x[i], yi = y[i], zi = z[i], gmi = gm[i];
        const double ri = RAD ? rad[i] : 0.0;
        double ax = 0.0, ay = 0.0, az = 0.0;
#pragma unroll 1
        for (int s = 0; s < 32; ++s) {
            const int j = jbase + ((lane + s) & 31);
            if (j >= n || j == i) continue;
            if (!((i < nplm) || (j < nplm))) continue;
            const double dx = x[j] - xi, dy = y[j] - yi, dz = z[j] - zi;
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            if (RAD) {
                const double rl = ri + rad[j];
                if (!(r2 > rl * rl)) continue;
            }
            const double irij3 = 1.0 / (r2 * sqrt(r2));
            const double fj = gm[j] * irij3, fi = gmi * irij3;
            ax = fma(fj, dx, ax);
            ay = fma(fj, dy, ay);
            az = fma(fj, dz, az);
            if (!diag) {  // a diagonal block visits (i,j) and (j,i)
                atomicAdd(fx + j, -(fi * dx));
                atomicAdd(fy + j, -(fi * dy));
                atomicAdd(fz + j, -(fi * dz));
            }
        }
        atomicAdd(fx + i, ax);
        atomicAdd(fy + i, ay);
        atomicAdd(fz + i, az);
    }
}

// The reaction accumulators are read by one lane and were written by its neighbour one step earlier; a __syncwarp() at
// the end of every step orders the hand-off.  The accesses are inline PTX so that their shared-window address is "lane
// base + immediate" (one pointer for the whole unrolled loop); volatile keeps them in program order for ptxas as well.
template <int OFF>
__device__ __forceinline__ double2 lds_acc2(unsigned addr)
{
    double2 v;
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ double lds_acc1(unsigned addr)
{
    double v;
    asm volatile("ld.volatile.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts_acc2(unsigned addr, double x, double y)
{
    asm volatile("st.volatile.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(addr), "n"(OFF), "d"(x), "d"(y));
}
template <int OFF>
__device__ __forceinline__ void sts_acc1(unsigned addr, double x)
{
    asm volatile("st.volatile.shared.f64 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "d"(x));
}

// One step of a chunk: lane l meets the column body in slot l+s.  `pc` / `pa` address slot l + s0 (generic pointer for
// the column data, shared-window address for the volatile accumulator accesses), K = s - s0 is a compile-time constant.
// The column body of the NEXT step is fetched first so that its LDS latency hides behind the arithmetic.
template <bool CHECKED, bool PRE, int K>
__device__ __forceinline__ void flat_step(const FlatArgs &a, const char *pc, unsigned pa, double2 &cxy, double2 &czg,
                                          int jslot0, int jbase, bool diag, int ibase, const double (&xi)[FIB],
                                          const double (&yi)[FIB], const double (&zi)[FIB], const double (&gmi)[FIB],
                                          double (&axi)[FIB], double (&ayi)[FIB], double (&azi)[FIB], unsigned &himin)
{
    double2 xy, zg;
    if (PRE) {  // the column body of the NEXT step is fetched first: its LDS latency hides behind this step's arithmetic
        xy = cxy, zg = czg;
        cxy = *reinterpret_cast<const double2 *>(pc + 16 * (K + 1));
        czg = *reinterpret_cast<const double2 *>(pc + T_ZG + 16 * (K + 1));
    } else {
        xy = *reinterpret_cast<const double2 *>(pc + 16 * K);
        zg = *reinterpret_cast<const double2 *>(pc + T_ZG + 16 * K);
    }
    const double2 t2 = lds_acc2<T_AXY + 16 * K>(pa);
    double ajx = t2.x, ajy = t2.y;
    double ajz = lds_acc1<T_AZ + 16 * K>(pa);
    const int jcur = jbase + ((jslot0 + K) & 31);  // CHECKED only
    // The four pairs of the step are written stage by stage (all differences, all seeds, all refinements ...): four
    // independent dependency chains side by side, which is the order the issue port wants (an FP64 result is usable
    // ~8 cycles after issue, an FP64 instruction issues every 2).
    unsigned hi[FIB];
    double dx[FIB], dy[FIB], dz[FIB], r2[FIB], s[FIB];
#pragma unroll
    for (int b = 0; b < FIB; ++b) {
        dx[b] = xy.x - xi[b];
        dy[b] = xy.y - yi[b];
        dz[b] = zg.x - zi[b];
    }
#pragma unroll
    for (int b = 0; b < FIB; ++b) {
        const double r2xy = fma(dy[b], dy[b], dx[b] * dx[b]);
        r2[b] = fma(dz[b], dz[b], r2xy);
        bool m = true;
        if (CHECKED) {
            // a masked pair contributes exactly zero (seed 0) and is left out of the running minimum: every chunk of a
            // diagonal block holds one self pair per lane
            const int i = ibase + 32 * b;
            m = (i != jcur) && (jcur < a.n) && (i < a.n) && ((i < a.nplm) || (jcur < a.nplm));
        }
        s[b] = rsq64h_seed<CHECKED>(r2[b], r2xy, m, hi[b]);
    }
    double s2[FIB], e[FIB], s3[FIB], y3[FIB];
#pragma unroll
    for (int b = 0; b < FIB; ++b) s2[b] = s[b] * s[b];
#pragma unroll
    for (int b = 0; b < FIB; ++b) {
        e[b] = fma(-r2[b], s2[b], 1.0);
        s3[b] = s2[b] * s[b];
    }
#pragma unroll
    for (int b = 0; b < FIB; ++b) {
        const double q = fma(1.875, e[b], 1.5);
        const double se = s3[b] * e[b];
        y3[b] = fma(se, q, s3[b]);
    }
#pragma unroll
    for (int b = 0; b < FIB; ++b) {
        const double fj = zg.y * y3[b];  // acts on i
        double fi = gmi[b] * y3[b];      // acts on j
        if (CHECKED) fi = diag ? 0.0 : fi;  // a diagonal block visits (i,j) and (j,i)
        axi[b] = fma(fj, dx[b], axi[b]);
        ayi[b] = fma(fj, dy[b], ayi[b]);
        azi[b] = fma(fj, dz[b], azi[b]);
        ajx = fma(-fi, dx[b], ajx);
        ajy = fma(-fi, dy[b], ajy);
        ajz = fma(-fi, dz[b], ajz);
    }
    static_assert(FIB == 4, "the high-word minimum below is written for 4 row bodies per lane");
    himin = __vimin3_u32(__vimin3_u32(himin, hi[0], hi[1]), hi[2], hi[3]);  // 2 x VIMNMX3 for 4 pairs
    sts_acc2<T_AXY + 16 * K>(pa, ajx, ajy);
    sts_acc1<T_AZ + 16 * K>(pa, ajz);
    // The neighbouring lane reads these slots in the next step.  The barrier costs nothing measurable here (6.85 vs
    // 6.84 ms: the loop is bound by the FP64 pipe, not by issue slots) and makes the hand-off formally ordered
    // (compute-sanitizer racecheck: 0 hazards; without it 40 intra-warp warnings).
    __syncwarp();
}

template <bool CHECKED, bool PRE, int K>
struct StepSeq {
    template <class... A>
    static __device__ __forceinline__ void run(A &&...args)
    {
        StepSeq<CHECKED, PRE, K - 1>::run(args...);
        flat_step<CHECKED, PRE, K - 1>(args...);
    }
};
template <bool CHECKED, bool PRE>
struct StepSeq<CHECKED, PRE, 0> {
    template <class... A>
    static __device__ __forceinline__ void run(A &&...) {}
};

// One block pair: block I resident in registers, block J streamed through the warp's shared tile 32 bodies at a time.
// CHECKED adds the index masks needed by diagonal blocks, the ragged last block and blocks that straddle nplm.
// `thr` = first high word of r^2 that is safely outside (max radius of I + max radius of J)^2 and inside the range the
// seeded refinement handles; a chunk in which any unmasked pair falls below it is rolled back and redone exactly.
__device__ __forceinline__ void cp_async8(void *smem_dst, const double *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}

template <bool RAD, bool CHECKED, int FUNROLL, bool PRE>
__device__ __forceinline__ void block_pair(const FlatArgs &a, WarpTile &w, int I, int J, bool diag, int lane, unsigned thr,
                                           bool force_exact, const double (&xi)[FIB], const double (&yi)[FIB],
                                           const double (&zi)[FIB], const double (&gmi)[FIB], double (&axi)[FIB],
                                           double (&ayi)[FIB], double (&azi)[FIB], int c0, int c1, int &fetched_j,
                                           int next_jfirst)
{
    const int ibase = I * FT + lane;
    // the column bodies of a chunk travel global -> w.stage (cp.async, issued one chunk ahead) -> tile; nothing of the
    // next chunk is held in registers while this one is computed
    auto fetch = [&](int jfirst) {
        const int jc = min(jfirst + lane, a.n - 1);
        cp_async8(&w.stage[0][lane], a.x + jc);
        cp_async8(&w.stage[1][lane], a.y + jc);
        cp_async8(&w.stage[2][lane], a.z + jc);
        cp_async8(&w.stage[3][lane], a.gm + jc);
    };
    // `fetched_j`: first column of the chunk that is in flight into w.stage (the previous block pair fetches the first
    // chunk of this one: `next_jfirst`), so the global-memory latency is exposed only at the start of the launch
    if (fetched_j != J * FT + c0 * 32) fetch(J * FT + c0 * 32);
    char *wb = reinterpret_cast<char *>(&w);
    const unsigned wa = (unsigned)__cvta_generic_to_shared(wb);
#pragma unroll 1
    for (int c = c0; c < c1; ++c) {
        const int jbase = J * FT + c * 32;
        const int jidx = jbase + lane;
        asm volatile("cp.async.wait_all;" ::: "memory");
        {
            const double nx = w.stage[0][lane], ny = w.stage[1][lane], nz = w.stage[2][lane], ng = w.stage[3][lane];
            __syncwarp();  // every lane is done with the previous chunk
            w.xy[lane] = w.xy[lane + 32] = make_double2(nx, ny);
            w.zg[lane] = w.zg[lane + 32] = make_double2(nz, ng);
        }
        w.axy[lane] = w.axy[lane + 32] = make_double2(0.0, 0.0);
        w.az[lane] = w.az[lane + 32] = make_double2(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            w.save[k][lane] = make_double2(axi[2 * k], axi[2 * k + 1]);
            w.save[2 + k][lane] = make_double2(ayi[2 * k], ayi[2 * k + 1]);
            w.save[4 + k][lane] = make_double2(azi[2 * k], azi[2 * k + 1]);
        }
        // the next chunk's loads (of this block pair, or the first chunk of the next one) fly while this one is computed
        if (c + 1 < c1) {
            fetch(jbase + 32);
        } else {
            fetched_j = next_jfirst;
            if (next_jfirst >= 0) fetch(next_jfirst);
        }
        __syncwarp();
        unsigned himin = 0xffffffffu;  // running minimum of the high words of r^2 over the chunk's unmasked pairs
        const char *pc = wb + 16 * lane;
        unsigned pa = wa + 16 * lane;
        double2 cxy = make_double2(0.0, 0.0), czg = cxy;
        if (PRE) {
            cxy = *reinterpret_cast<const double2 *>(pc);
            czg = *reinterpret_cast<const double2 *>(pc + T_ZG);
        }
#pragma unroll 1
        for (int s0 = 0; s0 < 32; s0 += FUNROLL) {
            StepSeq<CHECKED, PRE, FUNROLL>::run(a, pc, pa, cxy, czg, lane + s0, jbase, diag, ibase, xi, yi, zi, gmi, axi, ayi,
                                           azi, himin);
            pc += 16 * FUNROLL;
            pa += 16 * FUNROLL;
        }
        __syncwarp();  // all accumulator stores of the chunk are done (and ordered before the plain loads below)
        const bool bad = __any_sync(0xffffffffu, himin < thr) || force_exact;
        if (__builtin_expect(bad, 0)) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const double2 sx = w.save[k][lane], sy = w.save[2 + k][lane], sz = w.save[4 + k][lane];
                axi[2 * k] = sx.x, axi[2 * k + 1] = sx.y;
                ayi[2 * k] = sy.x, ayi[2 * k + 1] = sy.y;
                azi[2 * k] = sz.x, azi[2 * k + 1] = sz.y;
            }
            redo_chunk<RAD>(a.x, a.y, a.z, a.gm, a.rad, a.fx, a.fy, a.fz, a.n, a.nplm, I, jbase, diag, lane);
            if (a.redo_count && lane == 0) atomicAdd(a.redo_count, 1ull);
        } else if (!diag && jidx < a.n) {
            // the two partial sums of column body jbase+lane
            const double2 p0 = w.axy[lane], p1 = w.axy[lane + 32];
            const double z0 = w.az[lane].x, z1 = w.az[lane + 32].x;
            atomicAdd(a.fx + jidx, p0.x + p1.x);
            atomicAdd(a.fy + jidx, p0.y + p1.y);
            atomicAdd(a.fz + jidx, z0 + z1);
        }
    }
}

template <bool RAD, int CTAS, int FUNROLL, bool PRE>
__global__ void __launch_bounds__(32 * FWARPS, CTAS) kick_flat_kernel(const FlatArgs a)
{
    __shared__ WarpTile tiles[FWARPS];
    const int lane = threadIdx.x & 31;
    WarpTile &w = tiles[threadIdx.x >> 5];
    double xi[FIB], yi[FIB], zi[FIB], gmi[FIB], axi[FIB], ayi[FIB], azi[FIB];
    int Icur = -1;
    int fetched_j = -1;  // first column of the chunk whose cp.async copies are in flight into the warp's staging area
    double radI = 0.0;
    // coordinates beyond COORD_SAFE_MAX could overflow r^2: every chunk then takes the exact path
    const bool force_exact = !(a.coordmax[0] < COORD_SAFE_MAX_F64);
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
            const int i = Icur * FT + b * 32 + lane;
            if (i < a.n) {
                atomicAdd(a.fx + i, axi[b]);
                atomicAdd(a.fy + i, ayi[b]);
                atomicAdd(a.fz + i, azi[b]);
            }
        }
    };

    // Dynamic scheduling: a warp claims `quantum` consecutive items at a time from a global counter.  (A static split
    // sized to exactly one resident wave was measured to be fragile: when the tail of the previous kernel still
    // occupies an SM at launch, one CTA is left over, runs alone after the others and doubles the kernel time.)
    // The first quantum of every warp is static (its warp id): no claim storm on the counter at launch.  Claimed values
    // then map to quanta total_warps, total_warps+1, ...  The claim for the NEXT quantum is issued before the current
    // one is processed, so the round trip of the atomic hides behind the arithmetic.
    auto claim = [&]() -> unsigned long long {
        unsigned long long c = 0;
        if (lane == 0) c = atomicAdd(a.counter, 1ull) + (unsigned long long)a.total_warps;
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
    unsigned long long q = (unsigned long long)((long long)blockIdx.x * FWARPS + (threadIdx.x >> 5));
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
            const int t = (int)(u / FIB);
            const int c0 = (int)(u - (long long)t * FIB);
            const int c1 = (int)min((long long)FIB, c0 + (u_end - u));
            u += c1 - c0;
            ndone += c1 - c0;
            int I, J;
            bool diag;
            item_decode(t, a, I, J, diag);
            // first column of the work that follows (next item of this claim, or the start of the prefetched claim)
            int next_jfirst = -1;
            {
                long long un = u, un_end = u_end;
                if (un >= un_end) {
                    const unsigned long long qn = __shfl_sync(0xffffffffu, qnext, 0);
                    if (!range(qn, un, un_end)) un = -1;
                }
                if (un >= 0) {
                    const int tn = (int)(un / FIB);
                    int In, Jn;
                    bool dn;
                    item_decode(tn, a, In, Jn, dn);
                    next_jfirst = Jn * FT + (int)(un - (long long)tn * FIB) * 32;
                }
            }
            if (I != Icur) {
                flush();
                Icur = I;
#pragma unroll
                for (int b = 0; b < FIB; ++b) {
                    const int ic = min(I * FT + b * 32 + lane, a.n - 1);
                    xi[b] = a.x[ic];
                    yi[b] = a.y[ic];
                    zi[b] = a.z[ic];
                    gmi[b] = a.gm[ic];
                    axi[b] = ayi[b] = azi[b] = 0.0;
                }
                radI = RAD ? a.blockrad[I] : 0.0;
            }
            unsigned thr = RSQ64H_HI_MIN;
            if (RAD) {
                const double rl = radI + a.blockrad[J];
                thr = rsq64h_threshold(rl * rl);
            }
            const bool checked = diag || I == a.nb - 1 || J == a.nb - 1 ||
                                 (a.nplm < a.n && (I == a.nbm - 1 || J == a.nbm - 1));
            if (checked)
                block_pair<RAD, true, FUNROLL, PRE>(a, w, I, J, diag, lane, thr, force_exact, xi, yi, zi, gmi, axi, ayi, azi,
                                                    c0, c1, fetched_j, next_jfirst);
            else
                block_pair<RAD, false, FUNROLL, PRE>(a, w, I, J, diag, lane, thr, force_exact, xi, yi, zi, gmi, axi, ayi, azi,
                                                     c0, c1, fetched_j, next_jfirst);
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

// Everything the third-law kernel needs prepared, in ONE launch (it used to be two memsets and two kernels, each a
// launch gap on a stream that otherwise holds a 0.9 ms kernel when eight GPUs share the work):
//   * the accumulation target F (3*stride doubles) zeroed,
//   * guard[parity] = max |coordinate| (integer max of the bit patterns) -- guard[1-parity] is zeroed for the NEXT
//     launch, so no separate reset is needed (the launcher alternates `parity`),
//   * blockrad[B] = max radius of block B (radius-checked variant),
//   * the work counter reset.
__global__ void __launch_bounds__(256) flat_prologue_kernel(double2 *F, size_t nF2, const double *radius, const double *x,
                                                            const double *y, const double *z, int n, int nb, double *blockrad,
                                                            unsigned long long *guard, int parity,
                                                            unsigned long long *counter)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (size_t k = tid; k < nF2; k += nth) F[k] = make_double2(0.0, 0.0);
    unsigned long long mc = 0ull;
    for (size_t i = tid; i < (size_t)n; i += nth) {
        const double c = fmax(fabs(x[i]), fmax(fabs(y[i]), fabs(z[i])));
        mc = max(mc, (unsigned long long)__double_as_longlong(c));  // NaN orders above every finite value: unsafe
    }
    for (int o = 16; o > 0; o >>= 1) mc = max(mc, __shfl_xor_sync(0xffffffffu, mc, o));
    if ((threadIdx.x & 31) == 0 && mc != 0ull) atomicMax(guard + parity, mc);
    if (tid == 0) {
        guard[1 - parity] = 0ull;
        *counter = 0ull;
    }
    if (radius != nullptr) {  // one warp per block of FT bodies
        const int lane = threadIdx.x & 31;
        for (int B = (int)(tid >> 5); B < nb; B += (int)(nth >> 5)) {
            double m = 0.0;
            for (int k = lane; k < FT; k += 32) {
                const int i = B * FT + k;
                if (i < n) m = fmax(m, fabs(radius[i]));
            }
            for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) blockrad[B] = m;
        }
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
    a.n = n;
    a.nplm = std::min(nplm_rows, n);
    a.nb = cdiv(n, FT);
    a.blockrad = nullptr;
    if (lrad) {  // per-block radius bound: the fast-path threshold of a block pair follows the bodies that are in it
        SWCU_CUDA(ctx, ctx->flat_blockrad.ensure(sizeof(double) * a.nb));
        a.blockrad = ctx->flat_blockrad.as<double>();
    }
    a.nbm = cdiv(a.nplm, FT);
    a.Km = (a.nbm - 1) / 2;
    a.evenm = (a.nbm % 2 == 0) ? 1 : 0;
    long long total = 0;
    for (int I = 0; I < a.nbm; ++I) total += items_of(I, a.nb, a.nbm, a.Km, a.evenm);
    a.total_items = total;
    if (total * FIB >= (1ll << 31)) return fail(ctx, SWCU_ERR_ARG, "kick_pl_flat: npl=%d exceeds the third-law kernel's 32-bit work index", n);

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

    // (CTAs per SM, steps per loop iteration, column prefetch) = (3, 2, 1) by measurement at npl = 1e5
    // (profiles/r02_kick_flat.md: 2, 3 and 4 CTAs/SM and 2/4/8 steps all land within 2 % of each other -- the hot loop
    // sits at the rate the FP64 pipe sustains for three-register operands).  SWCU_FLAT_CFG selects the alternates that
    // are kept compiled for re-measurement: 381 (8 steps per iteration), 480 (4 CTAs/SM at 128 registers).
    static const int cfg = getenv("SWCU_FLAT_CFG") ? atoi(getenv("SWCU_FLAT_CFG")) : 321;
    void (*kern)(const FlatArgs);
    switch (cfg) {
#define FLAT_CFG(C, U, P) \
    case C * 100 + U * 10 + P: kern = lrad ? kick_flat_kernel<true, C, U, P> : kick_flat_kernel<false, C, U, P>; break;
        FLAT_CFG(3, 8, 1) FLAT_CFG(4, 8, 0)
        default: kern = lrad ? kick_flat_kernel<true, 3, 2, 1> : kick_flat_kernel<false, 3, 2, 1>;
#undef FLAT_CFG
    }
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * FWARPS, 0);
    occ = std::max(1, occ);
    a.item0 = 0;
    a.item1 = total;
    // Several GPUs (peer-memory or NCCL mode alike): every rank owns a balanced consecutive run of items and schedules it
    // with its OWN counter, exactly like a single GPU schedules the whole triangle.  (Round 1 let all GPUs claim from one
    // counter in rank 0's memory over NVLink; at eight GPUs the claim traffic to one address and the ragged end cost 9 %
    // of the gravity time, and the shared counter's epoch arithmetic was fragile -- ADVICE r1.  Identical GPUs at
    // identical clocks finish equal shares within the same ~2 % as the warps of one GPU do.)
    int nr = reduce ? ctx->nranks : ctx->p2p.nranks, rk = reduce ? ctx->rank : ctx->p2p.rank;
    // development aid: time the share one of SWCU_FLAT_EMULATE_RANKS ranks would get on a single GPU (the result is
    // then incomplete, only the launch time means something; the kernel has no cross-GPU interaction any more)
    static const int emulate = getenv("SWCU_FLAT_EMULATE_RANKS") ? atoi(getenv("SWCU_FLAT_EMULATE_RANKS")) : 0;
    if (nr == 1 && emulate > 1) {
        nr = emulate;
        rk = emulate / 2;
    }
    if (nr > 1) {  // balanced consecutive runs, like swcu_partition
        const long long q = total / nr, r = total % nr;
        a.item0 = rk * q + std::min<long long>(rk, r);
        a.item1 = a.item0 + q + (rk < r ? 1 : 0);
    }
    const long long mine = a.item1 - a.item0;
    const long long warps_max = (long long)ctx->prop.multiProcessorCount * occ * FWARPS;
    // Items per coarse claim.  Every claim switches the resident row block (12 REDs + 16 loads, with only three warps
    // per SMSP to hide it): measured at npl = 1e5 with the graded end of the schedule, 1 item 7.90 ms, 2 items 7.73,
    // 4 items 7.67, 6 items 7.65 (but slower at 7e4 bodies).  Fewer items per warp -> smaller claims.
    a.quantum = 4;
    if (mine < 32 * warps_max) a.quantum = 2;
    if (mine < 16 * warps_max) a.quantum = 1;
    if (ctx->tune_nsplit > 0) a.quantum = ctx->tune_nsplit;
    SWCU_CUDA(ctx, ctx->scratch64.ensure(128));
    a.counter = ctx->scratch64.as<unsigned long long>() + 3;
    if (!ctx->flat_redo.p) {
        SWCU_CUDA(ctx, ctx->flat_redo.ensure(sizeof(unsigned long long)));
        SWCU_CUDA(ctx, cudaMemsetAsync(ctx->flat_redo.p, 0, sizeof(unsigned long long), ctx->stream));
    }
    a.redo_count = ctx->flat_redo.as<unsigned long long>();  // cumulative over launches (swcu_flat_redo_count)
    // persistent grid: every SM filled to its occupancy (or fewer CTAs when there is little work).  The last chunks per
    // warp are handed out in 8-, 4-, 2- and 1-chunk claims at the end of the run.  A claim of s chunks can take ~2 s
    // chunk-times on a warp the scheduler disfavours, so everything finer than s has to last that long: 2 s chunks per
    // warp for size s.
    static int fine[4] = {2, 4, 8, 16};
    static bool fine_read = false;
    if (!fine_read) {
        fine_read = true;
        if (const char *e = getenv("SWCU_FLAT_FINE")) sscanf(e, "%d,%d,%d,%d", &fine[0], &fine[1], &fine[2], &fine[3]);
    }
    auto split_quanta = [&](long long items, long long warps_all) -> long long {
        const long long U = items * FIB, G = (long long)a.quantum * FIB;
        long long left = U;
        long long nq[5] = {0, 0, 0, 0, 0};  // phases in run order: G, 8, 4, 2, 1 chunks
        const int sz[5] = {(int)G, 8, 4, 2, 1};
        for (int k = 4; k >= 1; --k) {  // carve the fine phases off the end of the run
            if (sz[k] >= G) continue;
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
    const long long nquanta = split_quanta(mine, warps_max);
    const long long units = std::max<long long>(1, std::min<long long>(warps_max, nquanta));
    a.total_warps = units;
    // one prologue launch: F = 0, max |coordinate| guard, block radius bounds, work counter = 0
    if (!ctx->flat_guard.p) {
        SWCU_CUDA(ctx, ctx->flat_guard.ensure(2 * sizeof(unsigned long long)));
        SWCU_CUDA(ctx, cudaMemsetAsync(ctx->flat_guard.p, 0, 2 * sizeof(unsigned long long), ctx->stream));
        ctx->flat_parity = 0;
    }
    const int parity = ctx->flat_parity;
    ctx->flat_parity ^= 1;
    a.coordmax = reinterpret_cast<const double *>(ctx->flat_guard.as<unsigned long long>() + parity);
    flat_prologue_kernel<<<2 * ctx->prop.multiProcessorCount, 256, 0, ctx->stream>>>(
        reinterpret_cast<double2 *>(a.fx), 3 * stride / 2, a.rad, a.x, a.y, a.z, n, a.nb, ctx->flat_blockrad.as<double>(),
        ctx->flat_guard.as<unsigned long long>(), parity, a.counter);
    SWCU_KERNEL_CHECK(ctx);
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
        std::string path = trace_path;
        if (nr > 1) path += ".rank" + std::to_string(rk);
        if (FILE *f = fopen(path.c_str(), "wb")) {
            fwrite(h.data(), sizeof(unsigned long long), h.size(), f);
            fclose(f);
        }
    }
    if (!reduce) return SWCU_OK;
    if (ctx->nranks > 1) SWCU_TRY(comm_allreduce_sum(ctx, a.fx, 3 * stride));
    return axpy3(ctx, 1.0, a.fx, a.fy, a.fz, pl.ax.as<double>(), pl.ay.as<double>(), pl.az.as<double>(), nullptr, n);
}

}  // namespace swcu
