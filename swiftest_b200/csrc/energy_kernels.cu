// energy_kernels.cu -- potential energy, kinetic energy and angular momentum of the massive bodies
// (the other O(N^2) loop of the reference, run at every output when ENERGY is on).
//
// Reference (src/swiftest/swiftest_util.f90):
//   swiftest_util_get_energy_and_momentum_system  :1172-1288   ke_orbit, L_orbit, be, te (lrotation = .false.)
//   swiftest_util_get_potential_energy_flat       :1291-1341   pe = sum_k -(Gm_i*m_j)/|rb_i - rb_j| + sum_i -GMcb*m_i/|rb_i|
//   swiftest_util_get_potential_energy_triangular :1344-1394   same sum, row by row
// Both reference variants add the same terms (in a thread-count dependent order: OpenMP reductions); one kernel serves
// both.  The result is compared with the CPU restatement to a relative 1e-13 of sum|terms|, the tolerance written in
// tests/test_gpu_parity.py.
//
// Pair kernel: blocks of PE_T = 256 bodies, cyclic block pairing (row block bi against column blocks bi, bi+1, ...,
// bi + nb/2 mod nb: every unordered block pair exactly once, equal work per row block).  A thread keeps one row body in
// registers and walks the column block staged in shared memory as (x, y, z, m); 1/r comes from the FP64 MUFU.RSQ64H seed
// and one third-order Newton step of kick_math.cuh (12 FP64 + ~5 other instructions per pair; round 1 used the FP32 seed:
// 12 + ~10).  No per-pair test: the loop keeps the running minimum of the high words of r2; coincident bodies (r2 below
// 2^-600) or a coordinate beyond 2^500 in the tile make the thread redo its tile row with IEEE sqrt and divide.
// Per-CTA partial sums are folded by a fixed tree: same bits run to run, independent of the SM count.
#include "kick_math.cuh"
#include "reduce.cuh"
#include "swcu_internal.cuh"

namespace swcu {
namespace {

constexpr int PE_T = 256;

__global__ void __launch_bounds__(PE_T) pe_pairs_kernel(int n, int nb, const double *__restrict__ x,
                                                        const double *__restrict__ y, const double *__restrict__ z,
                                                        const double *__restrict__ gm, const double *__restrict__ mass,
                                                        const int32_t *__restrict__ lmask, double *__restrict__ partials)
{
    __shared__ double4 col[PE_T];
    __shared__ double wsum[PE_T / 32];
    const int bi = blockIdx.y, c = blockIdx.x;
    const int bj = (bi + c) % nb;
    const int cta = blockIdx.y * gridDim.x + blockIdx.x;
    // with an even block count the opposite pairing (c == nb/2) appears from both sides: keep the one with bi < bj
    const bool dead = (nb % 2 == 0) && (c == nb / 2) && (bi >= bj);
    double s = 0.0;
    if (!dead) {
        const int j = bj * PE_T + threadIdx.x;
        const bool jon = j < n && lmask[j] != 0;
        const double4 cj = jon ? make_double4(x[j], y[j], z[j], mass[j]) : make_double4(0.0, 0.0, 0.0, 0.0);
        col[threadIdx.x] = cj;
        const int i = bi * PE_T + threadIdx.x;
        const bool ion = i < n && lmask[i] != 0;
        const double xi = ion ? x[i] : 0.0, yi = ion ? y[i] : 0.0, zi = ion ? z[i] : 0.0;
        const double gi = ion ? gm[i] : 0.0;
        // the FP64 seed (MUFU.RSQ64H) needs r2 in [2^-600, inf): every |coordinate| of the tile below 2^500 (NaN fails the
        // test), and the running minimum of the high words of r2 below catches coincident bodies
        const bool wild = !(fabs(cj.x) < COORD_SAFE_MAX_F64 && fabs(cj.y) < COORD_SAFE_MAX_F64 && fabs(cj.z) < COORD_SAFE_MAX_F64 &&
                            fabs(xi) < COORD_SAFE_MAX_F64 && fabs(yi) < COORD_SAFE_MAX_F64 && fabs(zi) < COORD_SAFE_MAX_F64);
        const int unsafe = __syncthreads_or(wild ? 1 : 0);
        unsigned hmin = 0xffffffffu;
        // diagonal block: only j > i; the columns up to and including the thread's own are skipped
        const int jbeg = (c == 0) ? threadIdx.x + 1 : 0;
        double s0 = 0.0, s1 = 0.0;
        int jj = jbeg;
        for (; jj + 1 < PE_T; jj += 2) {
            const double4 p = col[jj], q = col[jj + 1];
            const double dx0 = p.x - xi, dy0 = p.y - yi, dz0 = p.z - zi;
            const double dx1 = q.x - xi, dy1 = q.y - yi, dz1 = q.z - zi;
            const double u0 = fma(dy0, dy0, dx0 * dx0), u1 = fma(dy1, dy1, dx1 * dx1);
            const double r0 = fma(dz0, dz0, u0);
            const double r1 = fma(dz1, dz1, u1);
            unsigned h0, h1;
            const double y0 = rsqrt_rsq64h(r0, u0, h0);
            const double y1 = rsqrt_rsq64h(r1, u1, h1);
            hmin = min(hmin, min(h0, h1));
            s0 = fma(p.w, y0, s0);
            s1 = fma(q.w, y1, s1);
        }
        if (jj < PE_T) {
            const double4 p = col[jj];
            const double dx0 = p.x - xi, dy0 = p.y - yi, dz0 = p.z - zi;
            const double u0 = fma(dy0, dy0, dx0 * dx0);
            const double r0 = fma(dz0, dz0, u0);
            unsigned h0;
            const double y0 = rsqrt_rsq64h(r0, u0, h0);
            hmin = min(hmin, h0);
            s0 = fma(p.w, y0, s0);
        }
        s = s0 + s1;
        if (unsafe || hmin < RSQ64H_HI_MIN) {  // a pair the seed cannot take: redo this thread's tile row with IEEE arithmetic
            s = 0.0;
            for (int k = jbeg; k < PE_T; ++k) {
                const double4 p = col[k];
                if (p.w == 0.0) continue;  // padding and masked-out columns carry zero mass
                const double dx0 = p.x - xi, dy0 = p.y - yi, dz0 = p.z - zi;
                const double r0 = fma(dz0, dz0, fma(dy0, dy0, dx0 * dx0));
                s += p.w / sqrt(r0);
            }
        }
        s = gi == 0.0 ? 0.0 : gi * s;  // a masked-out row contributes nothing (and hides any inf from its own redo)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = wsum[0];
        for (int w = 1; w < PE_T / 32; ++w) t += wsum[w];
        partials[cta] = t;
    }
}

struct PartialTerm {  // folds the per-CTA partials of pe_pairs_kernel
    const double *p;
    __device__ bool operator()(int i, double *t) const
    {
        t[0] = p[i];
        return true;
    }
};
struct StoreFin {
    double *out;
    int k;
    __device__ void operator()(const double *s) const
    {
        for (int c = 0; c < k; ++c) out[c] = s[c];
    }
};

// per-body terms of swiftest_util.f90:1214-1225 (+ the central-body potential and the binding energy):
// [0] m v.v  [1..3] m (r x v)  [4] Gm  [5] GMcb m/|r|  [6] 3 Gm m/(5 R)
struct BodyTerm {
    const double *x, *y, *z, *vx, *vy, *vz, *gm, *mass, *radius;
    const int32_t *lmask;
    double gmcb;
    bool lclose, pe_only;
    __device__ bool operator()(int i, double *t) const
    {
        if (lmask[i] == 0) return false;
        const double m = mass[i];
        const double rx = x[i], ry = y[i], rz = z[i];
        t[5] = gmcb * m / sqrt(rx * rx + ry * ry + rz * rz);
        t[4] = gm[i];
        if (pe_only) {
            t[0] = t[1] = t[2] = t[3] = t[6] = 0.0;
            return true;
        }
        const double ux = vx[i], uy = vy[i], uz = vz[i];
        t[0] = m * (ux * ux + uy * uy + uz * uz);
        t[1] = m * (ry * uz - rz * uy);
        t[2] = m * (rz * ux - rx * uz);
        t[3] = m * (rx * uy - ry * ux);
        t[6] = lclose ? 3 * gm[i] * m / (5 * radius[i]) : 0.0;
        return true;
    }
};

template <int K, class Term>
int run_sum(swcu_context *ctx, int n, Term term, double *d_out)
{
    StoreFin fin{d_out, K};
    if (n <= SERIAL_SUM_MAX) {
        sum_serial_kernel<K><<<1, SERIAL_THREADS, 0, ctx->stream>>>(n, false, term, fin);
    } else {
        double *partials = ctx->sumbuf.as<double>();
        unsigned *ticket = reinterpret_cast<unsigned *>(partials + (size_t)SUM_MAX_CTAS * 8);
        sum_tree_kernel<K><<<sum_grid(n), SUM_THREADS, 0, ctx->stream>>>(n, term, fin, partials, ticket);
    }
    SWCU_KERNEL_CHECK(ctx);
    return SWCU_OK;
}

}  // namespace

// b: rb in r*, vb in v*, Gmass in Gm, mass in mu, radius in radius, mask in lmask.
// out8 (host): [0] sum m v.v  [1] pair sum Gm_i m_j/r_ij  [2] sum GMcb m/|r|  [3] sum 3 Gm m/(5R)  [4..6] sum m (r x v)
//              [7] sum Gm
int energy_and_momentum(swcu_context *ctx, Body &b, double gmcb, int lclose, bool pe_only, double *out8)
{
    for (int k = 0; k < 8; ++k) out8[k] = 0.0;
    const int n = b.n;
    if (n <= 0) return SWCU_OK;
    SWCU_TRY(ensure_step_state(ctx));
    double *d_e = ctx->cbs.as<double>() + CBS_ENERGY;  // [0..6] body sums, [7] pair sum
    const int nb = cdiv(n, PE_T);
    const int ncol = nb / 2 + 1;
    const size_t ncta = (size_t)nb * ncol;
    SWCU_CUDA(ctx, ctx->partial.ensure(sizeof(double) * ncta));
    {
        FamTimer ft(ctx, FAM_PLPL);
        pe_pairs_kernel<<<dim3(ncol, nb), PE_T, 0, ctx->stream>>>(n, nb, b.rx.as<double>(), b.ry.as<double>(),
                                                                  b.rz.as<double>(), b.Gm.as<double>(), b.mu.as<double>(),
                                                                  b.lmask.as<int32_t>(), ctx->partial.as<double>());
        SWCU_KERNEL_CHECK(ctx);
        SWCU_TRY(run_sum<1>(ctx, (int)ncta, PartialTerm{ctx->partial.as<double>()}, d_e + 7));
    }
    BodyTerm bt{b.rx.as<double>(), b.ry.as<double>(), b.rz.as<double>(), b.vx.as<double>(), b.vy.as<double>(),
                b.vz.as<double>(), b.Gm.as<double>(), b.mu.as<double>(), b.radius.as<double>(), b.lmask.as<int32_t>(),
                gmcb, lclose != 0, pe_only};
    SWCU_TRY(run_sum<7>(ctx, n, bt, d_e));
    double h[8];
    SWCU_CUDA(ctx, cudaMemcpyAsync(h, d_e, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    out8[0] = h[0];
    out8[1] = h[7];
    out8[2] = h[5];
    out8[3] = h[6];
    out8[4] = h[1];
    out8[5] = h[2];
    out8[6] = h[3];
    out8[7] = h[4];
    return SWCU_OK;
}

}  // namespace swcu
