// list_kernels.cu -- the SyMBA recursion's encounter-list kernels, host-pointer tier.
//
// Reference (paths relative to src/):
//   symba_kick_list_plpl / _pltp     symba/symba_kick.f90:126-337   kick the bodies of the pairs of this recursion level
//   collision_check_plpl / _pltp     collision/collision_check.f90:61-250 (pair loops :96-110, :213-223)
//   collision_check_one              collision/collision_check.f90:15-58
//   swiftest_orbel_xv2aeq            swiftest/swiftest_orbel.f90:700-764
//
// symba_kick_list_* is a SERIAL loop in the reference: every pair adds into ah(i) and ah(j) in list order, then every
// body of a surviving pair gets vb += sgn*dt*ah once.  To reproduce the serial sums bit for bit the device builds the
// body -> pairs adjacency with one key sort ((body << 32) | k, so a body's pairs come out in list order) and one
// thread per body adds its contributions in that order.  The per-pair factor is computed once per pair by a first
// kernel.  Compiled with --fmad=false; the only libm call is pow(r2, -1.5) inside the shell (2 ulp on the device).
#include "swcu_internal.cuh"

#include <cmath>
#include <cub/cub.cuh>

namespace swcu {
namespace {

constexpr double RHSCALE = 6.5, RSHELL = 0.48075;  // symba_module.f90:22-23
constexpr double TINYVALUE = 4.0e-15;              // swiftest_orbel.f90:11
constexpr unsigned long long NOKEY = ~0ull;

// a 3-vector array in either layout: the caller's Fortran r(3,n) (AoS, stride 3: tier 1) or the resident SoA arrays of a
// population (stride 1: tier 2).  The kernels below are the same for both tiers.
struct V3 {
    const double *x, *y, *z;
    int s;
};
struct V3m {
    double *x, *y, *z;
    int s;
};
inline V3 aos(const double *r) { return V3{r, r + 1, r + 2, 3}; }
inline V3m aos_m(double *r) { return V3m{r, r + 1, r + 2, 3}; }
inline V3 soa(const DevBuf &x, const DevBuf &y, const DevBuf &z) { return V3{x.as<double>(), y.as<double>(), z.as<double>(), 1}; }
inline V3m soa_m(DevBuf &x, DevBuf &y, DevBuf &z) { return V3m{x.as<double>(), y.as<double>(), z.as<double>(), 1}; }

// x**n with an integer variable exponent as libgfortran evaluates it (_gfortran_pow_r8_i4)
__host__ __device__ inline double pow_r8_i4(double a, int b)
{
    double pw = 1.0, x = a;
    if (b != 0) {
        unsigned u;
        if (b < 0) {
            u = (unsigned)(-b);
            x = pw / x;
        } else {
            u = (unsigned)b;
        }
        for (;;) {
            if (u & 1u) pw *= x;
            u >>= 1;
            if (u) x *= x;
            else break;
        }
    }
    return pw;
}

// symba_kick.f90:180-201 / 284-304: false when the pair lies inside the inner shell (r2 < rim1)
__device__ __forceinline__ bool symba_list_fac(double rhsum, double r2, double shell2, double &fac)
{
    const double ri = (rhsum * rhsum) * (RHSCALE * RHSCALE) * shell2;  // shell2 = RSHELL**(2*irecl)
    const double rim1 = ri * (RSHELL * RSHELL);
    if (r2 < rim1) {
        fac = 0.0;
        return false;
    }
    if (r2 < ri) {
        const double ris = sqrt(ri);
        const double r = sqrt(r2);
        const double rr = (ris - r) / (ris * (1.0 - RSHELL));
        fac = pow(r2, -1.5) * (1.0 - 3 * (rr * rr) + 2 * (rr * rr * rr));
    } else {
        fac = 1.0 / (r2 * sqrt(r2));
    }
    return true;
}

// per pair: level mask, separation, force factor; emits the half-edge keys of the surviving pairs
// r1/r2: positions of list 1 / list 2 (r2 == r1 for pl-pl); flag: 0 not at this level, 1 kicked, 2 inner shell
__global__ void symba_pair_kernel(long long nenc, const int32_t *__restrict__ index1, const int32_t *__restrict__ index2,
                                  const int32_t *__restrict__ lactive, const int32_t *__restrict__ levelg1,
                                  const int32_t *__restrict__ levelg2, V3 r1, V3 r2,
                                  const double *__restrict__ rhill1, int plpl, int irm1,
                                  double shell2, double *__restrict__ pfac, double *__restrict__ pdx,
                                  int32_t *__restrict__ flag, unsigned long long *__restrict__ keys)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nenc) return;
    const int i = index1[k] - 1, j = index2[k] - 1;
    bool good = (levelg1[i] >= irm1) && (levelg2[j] >= irm1);
    if (lactive) good = good && (lactive[k] != 0);
    int fl = 0;
    unsigned long long k1 = NOKEY, k2 = NOKEY;
    if (good) {
        const double dx = r2.x[r2.s * j] - r1.x[r1.s * i], dy = r2.y[r2.s * j] - r1.y[r1.s * i],
                     dz = r2.z[r2.s * j] - r1.z[r1.s * i];
        const double rr2 = dx * dx + dy * dy + dz * dz;
        const double rhsum = plpl ? rhill1[i] + rhill1[j] : rhill1[i];
        double fac;
        if (symba_list_fac(rhsum, rr2, shell2, fac)) {
            fl = 1;
            pfac[k] = fac;
            pdx[3 * k] = dx;
            pdx[3 * k + 1] = dy;
            pdx[3 * k + 2] = dz;
            if (plpl) k1 = ((unsigned long long)(unsigned)i << 32) | (unsigned long long)k;
            k2 = ((unsigned long long)(unsigned)j << 32) | (unsigned long long)k;
        } else {
            fl = 2;
        }
    }
    flag[k] = fl;
    if (plpl) {
        keys[2 * k] = k1;
        keys[2 * k + 1] = k2;
    } else {
        keys[k] = k2;
    }
}

// one thread per sorted half-edge; the first half-edge of a body walks the body's run in list order
__global__ void symba_body_kernel(long long nkeys, const unsigned long long *__restrict__ keys,
                                  const int32_t *__restrict__ index1, const double *__restrict__ gm1,
                                  const int32_t *__restrict__ index2, const double *__restrict__ pfac,
                                  const double *__restrict__ pdx, int plpl, double sdt, V3m vb)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nkeys) return;
    const unsigned long long key = keys[p];
    if (key == NOKEY) return;
    const unsigned body = (unsigned)(key >> 32);
    if (p > 0 && (unsigned)(keys[p - 1] >> 32) == body) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (long long q = p; q < nkeys; ++q) {
        const unsigned long long kq = keys[q];
        if (kq == NOKEY || (unsigned)(kq >> 32) != body) break;
        const long long k = (long long)(kq & 0xffffffffull);
        const int i = index1[k] - 1, j = index2[k] - 1;
        const double fac = pfac[k];
        const double dx = pdx[3 * k], dy = pdx[3 * k + 1], dz = pdx[3 * k + 2];
        if (plpl && (unsigned)i == body) {  // ah(i) = ah(i) + facj * dx
            const double facj = fac * gm1[j];
            a0 = a0 + facj * dx;
            a1 = a1 + facj * dy;
            a2 = a2 + facj * dz;
        } else {  // ah(j) = ah(j) - faci * dx
            const double faci = fac * gm1[i];
            a0 = a0 - faci * dx;
            a1 = a1 - faci * dy;
            a2 = a2 - faci * dz;
        }
    }
    const size_t o = (size_t)vb.s * body;
    vb.x[o] = vb.x[o] + sdt * a0;
    vb.y[o] = vb.y[o] + sdt * a1;
    vb.z[o] = vb.z[o] + sdt * a2;
}

// swiftest_orbel_xv2aeq: only q is needed here
__device__ double orbel_xv2aeq_q(double mu, double rx, double ry, double rz, double vx, double vy, double vz)
{
    double a = 0.0, e = 0.0, q = 0.0;
    const double r = sqrt(rx * rx + ry * ry + rz * rz);
    const double v2 = vx * vx + vy * vy + vz * vz;
    const double hx = ry * vz - rz * vy, hy = rz * vx - rx * vz, hz = rx * vy - ry * vx;
    const double h2 = hx * hx + hy * hy + hz * hz;
    if (h2 < 2.2250738585072014e-308) return q;  // tiny(h2)
    const double energy = 0.5 * v2 - mu / r;
    int type;  // -1 ellipse, 0 parabola, 1 hyperbola
    double fac = 0.0;
    if (fabs(energy * r / mu) < sqrt(TINYVALUE)) {
        type = 0;
    } else {
        a = -0.5 * mu / energy;
        if (a < 0.0) {
            fac = -h2 / (mu * a);
            type = (fac > TINYVALUE) ? 1 : 0;
        } else {
            type = -1;
        }
    }
    if (type == -1) {
        fac = 1.0 - h2 / (mu * a);
        if (fac > TINYVALUE) e = sqrt(fac);
        q = a * (1.0 - e);
    } else if (type == 0) {
        a = 0.5 * h2 / mu;
        q = a;
    } else {
        e = sqrt(1.0 + fac);
        q = a * (1.0 - e);
    }
    return q;
}

// collision_check_one over the pairs of the mask; xr = r1(i) - r2(j), vr = v1(i) - v2(j)
__global__ void collision_check_kernel(long long nenc, const int32_t *__restrict__ index1,
                                       const int32_t *__restrict__ index2, const int32_t *__restrict__ lmask,
                                       const int32_t *__restrict__ lvdotr, V3 r1, V3 v1,
                                       const double *__restrict__ gm1, const double *__restrict__ rad1, V3 r2, V3 v2,
                                       const double *__restrict__ gm2,
                                       const double *__restrict__ rad2, double dt, int32_t *__restrict__ lcollision,
                                       int32_t *__restrict__ lclosest, unsigned long long *__restrict__ count)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nenc) return;
    int lcol = 0, lclo = 0;
    if (!lmask || lmask[k] != 0) {
        const int i = index1[k] - 1, j = index2[k] - 1;
        const double xr = r1.x[r1.s * i] - r2.x[r2.s * j], yr = r1.y[r1.s * i] - r2.y[r2.s * j],
                     zr = r1.z[r1.s * i] - r2.z[r2.s * j];
        const double vxr = v1.x[v1.s * i] - v2.x[v2.s * j], vyr = v1.y[v1.s * i] - v2.y[v2.s * j],
                     vzr = v1.z[v1.s * i] - v2.z[v2.s * j];
        const double rlim = gm2 ? rad1[i] + rad2[j] : rad1[i];
        const double gmtot = gm2 ? gm1[i] + gm2[j] : gm1[i];
        const double rr2 = xr * xr + yr * yr + zr * zr;
        const double rlim2 = rlim * rlim;
        if (rr2 <= rlim2) {
            lcol = 1;
        } else {
            const double vdotr = xr * vxr + yr * vyr + zr * vzr;
            if (lvdotr[k] != 0 && vdotr > 0.0) {
                const double tcr2 = rr2 / (vxr * vxr + vyr * vyr + vzr * vzr);
                const double dt2 = dt * dt;
                if (tcr2 <= dt2) lcol = (orbel_xv2aeq_q(gmtot, xr, yr, zr, vxr, vyr, vzr) < rlim) ? 1 : 0;
                lclo = lcol ? 0 : 1;
            }
        }
    }
    lcollision[k] = lcol;
    lclosest[k] = lclo;
    if (lcol) atomicAdd(count, 1ull);
}

int use_ctx(swcu_context *ctx)
{
    if (!ctx) return SWCU_ERR_ARG;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return fail(ctx, SWCU_ERR_CUDA, "cudaSetDevice(%d): %s", ctx->device, cudaGetErrorString(e));
    return SWCU_OK;
}

template <class T> int put(swcu_context *ctx, DevBuf &d, const T *h, size_t n)
{
    SWCU_CUDA(ctx, d.ensure(sizeof(T) * (n > 0 ? n : 1)));
    if (n > 0) SWCU_CUDA(ctx, cudaMemcpyAsync(d.p, h, sizeof(T) * n, cudaMemcpyHostToDevice, ctx->stream));
    return SWCU_OK;
}

int check_indices(swcu_context *ctx, const char *who, int64_t nenc, const int32_t *index1, const int32_t *index2, int32_t n1,
                  int32_t n2)
{
    for (int64_t k = 0; k < nenc; ++k)  // the reference trusts its own lists; a foreign caller gets a checked error
        if (index1[k] < 1 || index1[k] > n1 || index2[k] < 1 || index2[k] > n2)
            return fail(ctx, SWCU_ERR_ARG, "%s: pair %lld = (%d,%d) out of range", who, (long long)k, index1[k], index2[k]);
    return SWCU_OK;
}

// the device part of symba_kick_list_*: everything already on the device (tier 1 uploads it, tier 2 has it resident);
// L[10..15] are scratch, the per-pair flags are left in L[12]
int symba_kick_list_core(swcu_context *ctx, bool plpl, int64_t nenc, const int32_t *d_i1, const int32_t *d_i2,
                         const int32_t *d_lactive, const int32_t *d_lev1, V3 r1, const double *d_rhill1, const double *d_gm1,
                         const int32_t *d_lev2, V3 r2, double dt, int32_t irec, int32_t sgn, V3m vb)
{
    auto &L = ctx->lists;
    const size_t ne = (size_t)nenc, nkeys = plpl ? 2 * ne : ne;
    SWCU_CUDA(ctx, L[10].ensure(sizeof(double) * ne));                  // pfac
    SWCU_CUDA(ctx, L[11].ensure(sizeof(double) * 3 * ne));              // pdx
    SWCU_CUDA(ctx, L[12].ensure(sizeof(int32_t) * ne));                 // flag
    SWCU_CUDA(ctx, L[13].ensure(sizeof(unsigned long long) * nkeys));   // keys
    SWCU_CUDA(ctx, L[14].ensure(sizeof(unsigned long long) * nkeys));   // sorted keys
    const int irm1 = irec - 1;
    const int irecl = (sgn < 0) ? irec - 1 : irec;
    const double shell2 = pow_r8_i4(RSHELL, 2 * irecl);
    const double sdt = sgn * dt;
    FamTimer ft(ctx, FAM_PLPL);
    symba_pair_kernel<<<cdiv(nenc, 256), 256, 0, ctx->stream>>>(nenc, d_i1, d_i2, d_lactive, d_lev1, d_lev2, r1, r2, d_rhill1,
                                                               plpl ? 1 : 0, irm1, shell2, L[10].as<double>(),
                                                               L[11].as<double>(), L[12].as<int32_t>(),
                                                               L[13].as<unsigned long long>());
    SWCU_KERNEL_CHECK(ctx);
    size_t tmp = 0;
    SWCU_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tmp, L[13].as<unsigned long long>(), L[14].as<unsigned long long>(),
                                                  (long long)nkeys, 0, 64, ctx->stream));
    SWCU_CUDA(ctx, L[15].ensure(tmp));
    SWCU_CUDA(ctx, cub::DeviceRadixSort::SortKeys(L[15].p, tmp, L[13].as<unsigned long long>(), L[14].as<unsigned long long>(),
                                                  (long long)nkeys, 0, 64, ctx->stream));
    ctx->launches += 2;
    symba_body_kernel<<<cdiv((long long)nkeys, 256), 256, 0, ctx->stream>>>((long long)nkeys, L[14].as<unsigned long long>(),
                                                                           d_i1, d_gm1, d_i2, L[10].as<double>(),
                                                                           L[11].as<double>(), plpl ? 1 : 0, sdt, vb);
    SWCU_KERNEL_CHECK(ctx);
    return SWCU_OK;
}

// per-pair flags of the last kick (L[12]) -> lgoodlevel of the reference
int fetch_lgood(swcu_context *ctx, int64_t nenc, int32_t *lgood)
{
    std::vector<int32_t> flag;
    if (lgood) {
        flag.resize((size_t)nenc);
        SWCU_CUDA(ctx, cudaMemcpyAsync(flag.data(), ctx->lists[12].p, sizeof(int32_t) * (size_t)nenc, cudaMemcpyDeviceToHost,
                                       ctx->stream));
    }
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (lgood)
        for (size_t k = 0; k < (size_t)nenc; ++k) lgood[k] = (flag[k] == 1) ? 1 : 0;
    return SWCU_OK;
}

// tier 1: shared body of the two host-pointer kick-list entry points
int symba_kick_list(swcu_context *ctx, bool plpl, int64_t nenc, const int32_t *index1, const int32_t *index2,
                    const int32_t *lactive, int32_t n1, const int32_t *levelg1, const double *r1, const double *rhill1,
                    const double *gm1, int32_t n2, const int32_t *levelg2, const double *r2, double dt, int32_t irec,
                    int32_t sgn, double *vb, int32_t *lgood)
{
    auto &L = ctx->lists;
    const size_t ne = (size_t)nenc;
    const int32_t nvb = plpl ? n1 : n2;
    SWCU_TRY(put(ctx, L[0], index1, ne));
    SWCU_TRY(put(ctx, L[1], index2, ne));
    if (lactive) SWCU_TRY(put(ctx, L[2], lactive, ne));
    SWCU_TRY(put(ctx, L[3], levelg1, (size_t)n1));
    SWCU_TRY(put(ctx, L[4], r1, 3 * (size_t)n1));
    SWCU_TRY(put(ctx, L[5], rhill1, (size_t)n1));
    SWCU_TRY(put(ctx, L[6], gm1, (size_t)n1));
    if (!plpl) {
        SWCU_TRY(put(ctx, L[7], levelg2, (size_t)n2));
        SWCU_TRY(put(ctx, L[8], r2, 3 * (size_t)n2));
    }
    SWCU_TRY(put(ctx, L[9], vb, 3 * (size_t)nvb));
    SWCU_TRY(symba_kick_list_core(ctx, plpl, nenc, L[0].as<int32_t>(), L[1].as<int32_t>(),
                                  lactive ? L[2].as<int32_t>() : nullptr, L[3].as<int32_t>(), aos(L[4].as<double>()),
                                  L[5].as<double>(), L[6].as<double>(), plpl ? L[3].as<int32_t>() : L[7].as<int32_t>(),
                                  aos(plpl ? L[4].as<double>() : L[8].as<double>()), dt, irec, sgn, aos_m(L[9].as<double>())));
    SWCU_CUDA(ctx, cudaMemcpyAsync(vb, L[9].p, sizeof(double) * 3 * (size_t)nvb, cudaMemcpyDeviceToHost, ctx->stream));
    return fetch_lgood(ctx, nenc, lgood);
}

// tier 2: the same kick on the resident populations -- positions, Hill radii, masses and the barycentric velocities
// stay on the device, a recursion level moves the pair list and the level arrays only
int symba_kick_list_resident(swcu_context *ctx, bool plpl, int64_t nenc, const int32_t *index1, const int32_t *index2,
                             const int32_t *lactive, const int32_t *levelg_pl, const int32_t *levelg_tp, double dt,
                             int32_t irec, int32_t sgn, int32_t *lgood)
{
    auto &L = ctx->lists;
    Body &pl = ctx->pl, &tp = ctx->tp;
    const size_t ne = (size_t)nenc;
    SWCU_TRY(put(ctx, L[0], index1, ne));
    SWCU_TRY(put(ctx, L[1], index2, ne));
    if (lactive) SWCU_TRY(put(ctx, L[2], lactive, ne));
    SWCU_TRY(put(ctx, L[3], levelg_pl, (size_t)pl.n));
    if (!plpl) SWCU_TRY(put(ctx, L[7], levelg_tp, (size_t)tp.n));
    Body &kicked = plpl ? pl : tp;
    SWCU_TRY(ensure_helio(ctx, kicked));
    const V3 rpl = soa(pl.rx, pl.ry, pl.rz);
    SWCU_TRY(symba_kick_list_core(ctx, plpl, nenc, L[0].as<int32_t>(), L[1].as<int32_t>(),
                                  lactive ? L[2].as<int32_t>() : nullptr, L[3].as<int32_t>(), rpl, pl.rhill.as<double>(),
                                  pl.Gm.as<double>(), plpl ? L[3].as<int32_t>() : L[7].as<int32_t>(),
                                  plpl ? rpl : soa(tp.rx, tp.ry, tp.rz), dt, irec, sgn,
                                  soa_m(kicked.wx, kicked.wy, kicked.wz)));
    if (!lgood) return SWCU_OK;  // nothing to read back: the call stays asynchronous
    return fetch_lgood(ctx, nenc, lgood);
}

// the device part of the collision pair loop; results in L[10] (lcollision, lclosest) and the hit count
int collision_check_core(swcu_context *ctx, int64_t nenc, const int32_t *d_i1, const int32_t *d_i2, const int32_t *d_lmask,
                         const int32_t *d_lvdotr, V3 r1, V3 v1, const double *d_gm1, const double *d_rad1, V3 r2, V3 v2,
                         const double *d_gm2, const double *d_rad2, double dt, int32_t *lcollision, int32_t *lclosest,
                         int64_t *ncollision)
{
    auto &L = ctx->lists;
    const size_t ne = (size_t)nenc;
    SWCU_CUDA(ctx, L[10].ensure(sizeof(int32_t) * 2 * ne));
    SWCU_CUDA(ctx, L[11].ensure(64));
    int32_t *d_col = L[10].as<int32_t>(), *d_clo = d_col + ne;
    unsigned long long *d_count = L[11].as<unsigned long long>();
    SWCU_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), ctx->stream));
    {
        FamTimer ft(ctx, FAM_SWEEP);
        collision_check_kernel<<<cdiv(nenc, 256), 256, 0, ctx->stream>>>(nenc, d_i1, d_i2, d_lmask, d_lvdotr, r1, v1, d_gm1,
                                                                        d_rad1, r2, v2, d_gm2, d_rad2, dt, d_col, d_clo,
                                                                        d_count);
        SWCU_KERNEL_CHECK(ctx);
    }
    unsigned long long h = 0;
    SWCU_CUDA(ctx, cudaMemcpyAsync(lcollision, d_col, sizeof(int32_t) * ne, cudaMemcpyDeviceToHost, ctx->stream));
    SWCU_CUDA(ctx, cudaMemcpyAsync(lclosest, d_clo, sizeof(int32_t) * ne, cudaMemcpyDeviceToHost, ctx->stream));
    SWCU_CUDA(ctx, cudaMemcpyAsync(&h, d_count, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ncollision) *ncollision = (int64_t)h;
    return SWCU_OK;
}

}  // namespace
}  // namespace swcu

using namespace swcu;

extern "C" int swcu_symba_kick_list_plpl(swcu_context *ctx, int64_t nenc, const int32_t *index1, const int32_t *index2,
                                         const int32_t *lactive, int32_t npl, const int32_t *levelg, const double *rh,
                                         const double *rhill, const double *Gmass, double dt, int32_t irec, int32_t sgn,
                                         double *vb, int32_t *lgood)
{
    SWCU_TRY(use_ctx(ctx));
    if (nenc < 0 || npl < 0 || nenc > 0x7fffffffll) return fail(ctx, SWCU_ERR_ARG, "symba_kick_list_plpl: bad argument");
    if (nenc == 0 || npl == 0) return SWCU_OK;  // symba_kick.f90:148, :153
    if (!index1 || !index2 || !levelg || !rh || !rhill || !Gmass || !vb)
        return fail(ctx, SWCU_ERR_ARG, "symba_kick_list_plpl: null array");
    SWCU_TRY(check_indices(ctx, "symba_kick_list_plpl", nenc, index1, index2, npl, npl));
    return symba_kick_list(ctx, true, nenc, index1, index2, lactive, npl, levelg, rh, rhill, Gmass, npl, levelg, rh, dt, irec,
                           sgn, vb, lgood);
}

extern "C" int swcu_symba_kick_list_pltp(swcu_context *ctx, int64_t nenc, const int32_t *index1, const int32_t *index2,
                                         const int32_t *lactive, int32_t npl, int32_t ntp, const int32_t *levelg_pl,
                                         const int32_t *levelg_tp, const double *rh_pl, const double *rhill,
                                         const double *Gmass, const double *rh_tp, double dt, int32_t irec, int32_t sgn,
                                         double *vb_tp, int32_t *lgood)
{
    SWCU_TRY(use_ctx(ctx));
    if (nenc < 0 || npl < 0 || ntp < 0 || nenc > 0x7fffffffll)
        return fail(ctx, SWCU_ERR_ARG, "symba_kick_list_pltp: bad argument");
    if (nenc == 0 || npl == 0 || ntp == 0) return SWCU_OK;  // symba_kick.f90:256, :263
    if (!index1 || !index2 || !levelg_pl || !levelg_tp || !rh_pl || !rhill || !Gmass || !rh_tp || !vb_tp)
        return fail(ctx, SWCU_ERR_ARG, "symba_kick_list_pltp: null array");
    SWCU_TRY(check_indices(ctx, "symba_kick_list_pltp", nenc, index1, index2, npl, ntp));
    return symba_kick_list(ctx, false, nenc, index1, index2, lactive, npl, levelg_pl, rh_pl, rhill, Gmass, ntp, levelg_tp,
                           rh_tp, dt, irec, sgn, vb_tp, lgood);
}

extern "C" int swcu_collision_check_list(swcu_context *ctx, int64_t nenc, const int32_t *index1, const int32_t *index2,
                                         const int32_t *lmask, const int32_t *lvdotr, int32_t n1, const double *r1,
                                         const double *v1, const double *Gmass1, const double *radius1, int32_t n2,
                                         const double *r2, const double *v2, double dt, int32_t *lcollision,
                                         int32_t *lclosest, int64_t *ncollision)
{
    SWCU_TRY(use_ctx(ctx));
    if (ncollision) *ncollision = 0;
    if (nenc < 0 || n1 < 0 || n2 < 0) return fail(ctx, SWCU_ERR_ARG, "collision_check_list: bad argument");
    if (nenc == 0) return SWCU_OK;  // collision_check.f90:84
    if (!index1 || !index2 || !lvdotr || !r1 || !v1 || !Gmass1 || !radius1 || !lcollision || !lclosest || n1 == 0)
        return fail(ctx, SWCU_ERR_ARG, "collision_check_list: null array");
    const bool two = n2 > 0;  // pl-tp form: the second list has no mass and no radius
    if (two && (!r2 || !v2)) return fail(ctx, SWCU_ERR_ARG, "collision_check_list: null second list");
    SWCU_TRY(check_indices(ctx, "collision_check_list", nenc, index1, index2, n1, two ? n2 : n1));
    auto &L = ctx->lists;
    const size_t ne = (size_t)nenc;
    SWCU_TRY(put(ctx, L[0], index1, ne));
    SWCU_TRY(put(ctx, L[1], index2, ne));
    if (lmask) SWCU_TRY(put(ctx, L[2], lmask, ne));
    SWCU_TRY(put(ctx, L[3], lvdotr, ne));
    SWCU_TRY(put(ctx, L[4], r1, 3 * (size_t)n1));
    SWCU_TRY(put(ctx, L[5], v1, 3 * (size_t)n1));
    SWCU_TRY(put(ctx, L[6], Gmass1, (size_t)n1));
    SWCU_TRY(put(ctx, L[7], radius1, (size_t)n1));
    if (two) {
        SWCU_TRY(put(ctx, L[8], r2, 3 * (size_t)n2));
        SWCU_TRY(put(ctx, L[9], v2, 3 * (size_t)n2));
    }
    return collision_check_core(ctx, nenc, L[0].as<int32_t>(), L[1].as<int32_t>(), lmask ? L[2].as<int32_t>() : nullptr,
                                L[3].as<int32_t>(), aos(L[4].as<double>()), aos(L[5].as<double>()), L[6].as<double>(),
                                L[7].as<double>(), aos(two ? L[8].as<double>() : L[4].as<double>()),
                                aos(two ? L[9].as<double>() : L[5].as<double>()), two ? nullptr : L[6].as<double>(),
                                two ? nullptr : L[7].as<double>(), dt, lcollision, lclosest, ncollision);
}

// ---- tier 2: the list kernels on the resident populations (pl%rh, pl%vb, tp%rh, tp%vb stay in HBM) ----------------
namespace {
int need_resident(swcu_context *ctx, const char *who, bool with_tp)
{
    SWCU_TRY(use_ctx(ctx));
    if (!ctx->pl.valid) return fail(ctx, SWCU_ERR_STATE, "%s: pl population not resident", who);
    if (with_tp && !ctx->tp.valid) return fail(ctx, SWCU_ERR_STATE, "%s: tp population not resident", who);
    return SWCU_OK;
}
}  // namespace

extern "C" int swcu_pl_symba_kick_list(swcu_context *ctx, int64_t nenc, const int32_t *index1, const int32_t *index2,
                                       const int32_t *lactive, const int32_t *levelg, double dt, int32_t irec, int32_t sgn,
                                       int32_t *lgood)
{
    SWCU_TRY(need_resident(ctx, "pl_symba_kick_list", false));
    if (nenc < 0 || nenc > 0x7fffffffll) return fail(ctx, SWCU_ERR_ARG, "pl_symba_kick_list: bad argument");
    if (nenc == 0 || ctx->pl.n == 0) return SWCU_OK;  // symba_kick.f90:148, :153
    if (!index1 || !index2 || !levelg) return fail(ctx, SWCU_ERR_ARG, "pl_symba_kick_list: null array");
    SWCU_TRY(check_indices(ctx, "pl_symba_kick_list", nenc, index1, index2, ctx->pl.n, ctx->pl.n));
    return symba_kick_list_resident(ctx, true, nenc, index1, index2, lactive, levelg, nullptr, dt, irec, sgn, lgood);
}

extern "C" int swcu_tp_symba_kick_list(swcu_context *ctx, int64_t nenc, const int32_t *index1, const int32_t *index2,
                                       const int32_t *lactive, const int32_t *levelg_pl, const int32_t *levelg_tp, double dt,
                                       int32_t irec, int32_t sgn, int32_t *lgood)
{
    SWCU_TRY(need_resident(ctx, "tp_symba_kick_list", true));
    if (nenc < 0 || nenc > 0x7fffffffll) return fail(ctx, SWCU_ERR_ARG, "tp_symba_kick_list: bad argument");
    if (nenc == 0 || ctx->pl.n == 0 || ctx->tp.n == 0) return SWCU_OK;  // symba_kick.f90:256, :263
    if (!index1 || !index2 || !levelg_pl || !levelg_tp) return fail(ctx, SWCU_ERR_ARG, "tp_symba_kick_list: null array");
    SWCU_TRY(check_indices(ctx, "tp_symba_kick_list", nenc, index1, index2, ctx->pl.n, ctx->tp.n));
    return symba_kick_list_resident(ctx, false, nenc, index1, index2, lactive, levelg_pl, levelg_tp, dt, irec, sgn, lgood);
}

extern "C" int swcu_body_collision_check_list(swcu_context *ctx, int32_t kind, int64_t nenc, const int32_t *index1,
                                              const int32_t *index2, const int32_t *lmask, const int32_t *lvdotr, double dt,
                                              int32_t *lcollision, int32_t *lclosest, int64_t *ncollision)
{
    const bool two = (kind == SWCU_TP);
    SWCU_TRY(need_resident(ctx, "body_collision_check_list", two));
    if (ncollision) *ncollision = 0;
    if (kind != SWCU_PL && kind != SWCU_TP) return fail(ctx, SWCU_ERR_ARG, "body_collision_check_list: kind");
    if (nenc < 0) return fail(ctx, SWCU_ERR_ARG, "body_collision_check_list: bad argument");
    if (nenc == 0) return SWCU_OK;  // collision_check.f90:84
    if (!index1 || !index2 || !lvdotr || !lcollision || !lclosest)
        return fail(ctx, SWCU_ERR_ARG, "body_collision_check_list: null array");
    Body &pl = ctx->pl, &tp = ctx->tp;
    SWCU_TRY(check_indices(ctx, "body_collision_check_list", nenc, index1, index2, pl.n, two ? tp.n : pl.n));
    auto &L = ctx->lists;
    const size_t ne = (size_t)nenc;
    SWCU_TRY(put(ctx, L[0], index1, ne));
    SWCU_TRY(put(ctx, L[1], index2, ne));
    if (lmask) SWCU_TRY(put(ctx, L[2], lmask, ne));
    SWCU_TRY(put(ctx, L[3], lvdotr, ne));
    SWCU_TRY(ensure_helio(ctx, pl));
    if (two) SWCU_TRY(ensure_helio(ctx, tp));
    const V3 r1 = soa(pl.rx, pl.ry, pl.rz), v1 = soa(pl.wx, pl.wy, pl.wz);
    return collision_check_core(ctx, nenc, L[0].as<int32_t>(), L[1].as<int32_t>(), lmask ? L[2].as<int32_t>() : nullptr,
                                L[3].as<int32_t>(), r1, v1, pl.Gm.as<double>(), pl.radius.as<double>(),
                                two ? soa(tp.rx, tp.ry, tp.rz) : r1, two ? soa(tp.wx, tp.wy, tp.wz) : v1,
                                two ? nullptr : pl.Gm.as<double>(), two ? nullptr : pl.radius.as<double>(), dt, lcollision,
                                lclosest, ncollision);
}
