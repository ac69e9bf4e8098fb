// encounter_kernels.cu -- sort-and-sweep close-encounter detection on the device.
//
// Replaces (reference, relative to /root/reference/src/encounter/encounter_check.f90):
//   encounter_check_all_sort_and_sweep_plpl :143-192, _plplm :195-258, _pltp :261-326,
//   encounter_check_sort_aabb_1D :763-792, encounter_check_sweep_aabb_single_list :905-988 and
//   _double_list :795-902, encounter_check_all_sweep_one :329-381, encounter_check_one :573-621,
//   encounter_check_collapse_ragged_list :624-673, encounter_check_remove_duplicates :676-760,
//   and the merge in encounter_check_all_plplm :42-109; symba_util_set_renc (symba/symba_util.f90:245-267).
//
// The broad phase is ONE-dimensional on heliocentric distance |r| -/+ 1.1*renc, body i is swept
// only when more than one foreign endpoint lies strictly inside its interval (F3), lvdotr is always true (F4).
//
// Pipeline (all HBM/latency bound integer + compare work, no tensor cores):
//   K7  extents + concatenated population                     elementwise
//   K8  device radix sort of the 2N (extent, endpoint-id) pairs (CUB DeviceRadixSort, stable => ties are
//       ordered by position in the [rmin,rmax] array, the order the oracle fixes)
//   K9  ibeg/iend scatter + gather of r,v,renc into sorted-endpoint order (SoA)
//   K10 chunked sweep: every body's interval is cut into chunks of SWEEP_CHUNK endpoints, a prefix sum
//       assigns chunks to warps (one Jupiter-sized interval over 1e6 test particles spreads over the whole
//       GPU), lanes stream the gathered records, evaluate encounter_check_one and append hits with a
//       warp-aggregated atomic
//   K11 canonical order: radix sort of the 64-bit (index1<<32|index2) keys + unique
//
// THIS FILE IS COMPILED WITH --fmad=false so that the predicate is the same sequence of individually
// rounded IEEE operations as the oracle: the pair list is bit-exact, not merely close.
#include "swcu_internal.cuh"

#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cub/cub.cuh>
#include <vector>

namespace swcu {
namespace {

constexpr double RSWEEP_FACTOR = 1.1;          // encounter_module.f90:21
constexpr double RHSCALE = 6.5, RSHELL = 0.48075;  // symba_module.f90:22-23
constexpr int SWEEP_CHUNK = 1024;

struct ListDev {
    const double *x, *y, *z, *vx, *vy, *vz, *renc;
    int n;
};

// order-preserving map double -> uint64 (the radix sort's own key transform) and back
__device__ __forceinline__ unsigned long long sortable(double x)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double unsortable(unsigned long long u)
{
    const unsigned long long b = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)b);
}

// mm[0] = min, mm[1] = max of all extents (sortable form), mm[2] != 0: a non-finite extent was seen (bucket sort refuses)
__device__ __forceinline__ void extent_minmax(unsigned long long *mm, double lo, double hi, bool valid)
{
    unsigned long long a = valid ? sortable(lo) : ~0ull, b = valid ? sortable(hi) : 0ull;
    if (valid && hi < lo) {  // negative renc
        const unsigned long long t = a;
        a = b, b = t;
    }
    const bool bad = valid && !(isfinite(lo) && isfinite(hi));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a = min(a, __shfl_xor_sync(0xffffffffu, a, o));
        b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&mm[0], a);
        atomicMax(&mm[1], b);
    }
    if (bad) atomicOr(reinterpret_cast<unsigned int *>(&mm[2]), 1u);
}

// K7: extents (encounter_check.f90:180-185, 237-251, 305-319) and the concatenated copy of both lists
__global__ void extent_kernel(ListDev l1, ListDev l2, int ntot, double *__restrict__ cx, double *__restrict__ cy,
                              double *__restrict__ cz, double *__restrict__ cvx, double *__restrict__ cvy,
                              double *__restrict__ cvz, double *__restrict__ crenc, double *__restrict__ keys,
                              int *__restrict__ vals, unsigned long long *__restrict__ mm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntot) {
        if (mm) extent_minmax(mm, 0.0, 0.0, false);  // the whole warp takes part in the reduction
        return;
    }
    const bool in1 = i < l1.n;
    const ListDev &l = in1 ? l1 : l2;
    const int q = in1 ? i : i - l1.n;
    const double x = l.x[q], y = l.y[q], z = l.z[q];
    const double renc = l.renc ? l.renc[q] : 0.0;
    cx[i] = x;
    cy[i] = y;
    cz[i] = z;
    cvx[i] = l.vx[q];
    cvy[i] = l.vy[q];
    cvz[i] = l.vz[q];
    crenc[i] = renc;
    const double rmag = sqrt(x * x + y * y + z * z);
    const double w = RSWEEP_FACTOR * renc;
    keys[i] = rmag - w;         // rmin -> begin endpoint id i
    keys[ntot + i] = rmag + w;  // rmax -> end endpoint id ntot+i
    vals[i] = i;
    vals[ntot + i] = ntot + i;
    if (mm) extent_minmax(mm, rmag - w, rmag + w, true);
}

// ---------------------------------------------------------------------------------------------------------------------
// K8': bucket sort of the 2N (extent, endpoint id) pairs -- the same result as the stable radix sort (a stable sort by key IS
// the sort by (key, position), and the endpoint id is the position), in 4 launches instead of the 10 of an 8-pass radix
// sort whose passes are latency bound at these sizes (10 us per pass at 2e4 keys: 88 of the 146 us of kernel time of the
// npl = 1e4 sweep).  Buckets are equal slices of [min, max] of the extents (the map key -> bucket is a chain of rounded,
// hence monotone, operations); a CTA sorts one bucket in shared memory with a bitonic network on (key, id) and does K9's
// work for its endpoints right away.  A bucket beyond the shared-memory capacity (a clump of equal radii) or a non-finite
// extent raises a flag: the buffers stay in-bounds, the call repeats itself with the radix sort.
constexpr int BUCKET_CAP = 2048;      // endpoints a CTA can sort
constexpr int BUCKET_TARGET = 128;    // average endpoints per bucket
constexpr int BUCKET_MAXNB = 1 << 16;
constexpr int BUCKET_MAXKEYS = 1 << 19;  // beyond ~5e5 keys the 8-pass radix sort is faster (2e6 keys: 0.61 vs 0.69 ms per sweep)

struct BucketMap {
    double kmin, scale;
    int nb;
    __device__ __forceinline__ int of(double key) const
    {
        int b = (int)((key - kmin) * scale);  // NaN -> 0; monotone in key
        return min(max(b, 0), nb - 1);
    }
};
__device__ __forceinline__ BucketMap bucket_map(const unsigned long long *mm, int nb)
{
    BucketMap m;
    const double kmin = unsortable(mm[0]), kmax = unsortable(mm[1]);
    m.kmin = kmin;
    m.scale = (kmax > kmin) ? (double)nb / (kmax - kmin) : 0.0;
    m.nb = nb;
    return m;
}

__global__ void bucket_hist_kernel(const double *__restrict__ keys, int next, const unsigned long long *__restrict__ mm,
                                   int nb, int *__restrict__ hist)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= next) return;
    atomicAdd(&hist[bucket_map(mm, nb).of(keys[k])], 1);
}

// one CTA: offs = exclusive scan of hist (offs[nb] = total), cursor = offs; flags a bucket beyond the capacity
__global__ void __launch_bounds__(1024) bucket_scan_kernel(const int *__restrict__ hist, int nb, int *__restrict__ offs,
                                                           int *__restrict__ cursor, unsigned long long *__restrict__ mm)
{
    __shared__ int part[1024];
    const int t = threadIdx.x, per = (nb + 1023) / 1024;
    const int b0 = min(t * per, nb), b1 = min(b0 + per, nb);
    int sum = 0, big = 0;
    for (int b = b0; b < b1; ++b) {
        sum += hist[b];
        big |= hist[b] > BUCKET_CAP;
    }
    part[t] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = t >= o ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int run = part[t] - sum;
    for (int b = b0; b < b1; ++b) {
        offs[b] = run;
        cursor[b] = run;
        run += hist[b];
    }
    if (t == 1023) offs[nb] = part[1023];
    if (big) atomicOr(reinterpret_cast<unsigned int *>(&mm[2]), 2u);
}

__global__ void bucket_scatter_kernel(const double *__restrict__ keys, int next, const unsigned long long *__restrict__ mm,
                                      int nb, int *__restrict__ cursor, double *__restrict__ bkeys, int *__restrict__ bids)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= next) return;
    const double key = keys[k];
    const int pos = atomicAdd(&cursor[bucket_map(mm, nb).of(key)], 1);
    bkeys[pos] = key;
    bids[pos] = k;  // vals_in[k] == k
}

// one CTA per bucket: bitonic sort on (key, id), then K9 for the bucket's endpoints (encounter_check.f90:778-789, :937-949)
__global__ void __launch_bounds__(128) bucket_sort_kernel(const int *__restrict__ offs, const double *__restrict__ bkeys,
                                                          const int *__restrict__ bids, int ntot,
                                                          const double *__restrict__ cx, const double *__restrict__ cy,
                                                          const double *__restrict__ cz, const double *__restrict__ cvx,
                                                          const double *__restrict__ cvy, const double *__restrict__ cvz,
                                                          const double *__restrict__ crenc, int *__restrict__ ibeg,
                                                          int *__restrict__ iend, double *__restrict__ sx,
                                                          double *__restrict__ sy, double *__restrict__ sz,
                                                          double *__restrict__ svx, double *__restrict__ svy,
                                                          double *__restrict__ svz, double *__restrict__ srenc,
                                                          int *__restrict__ sbody)
{
    __shared__ double sk[BUCKET_CAP];
    __shared__ int si[BUCKET_CAP];
    const int o0 = offs[blockIdx.x], cnt = offs[blockIdx.x + 1] - o0;
    if (cnt == 0) return;
    const bool sortit = cnt <= BUCKET_CAP;  // otherwise pass the bucket through unsorted (flag already raised)
    int m = 1;
    if (sortit) {
        while (m < cnt) m <<= 1;
        for (int k = threadIdx.x; k < m; k += blockDim.x) {
            sk[k] = k < cnt ? bkeys[o0 + k] : INFINITY;
            si[k] = k < cnt ? bids[o0 + k] : 0x7fffffff;
        }
        __syncthreads();
        for (int size = 2; size <= m; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int k = threadIdx.x; k < (m >> 1); k += blockDim.x) {
                    const int lo = ((k & ~(stride - 1)) << 1) | (k & (stride - 1)), hi = lo + stride;  // stride is a power of 2
                    const bool up = ((lo & size) == 0);
                    const double ka = sk[lo], kb = sk[hi];
                    const int ia = si[lo], ib = si[hi];
                    const bool a_after_b = (ka > kb) || (ka == kb && ia > ib);
                    if (a_after_b == up) {
                        sk[lo] = kb, sk[hi] = ka;
                        si[lo] = ib, si[hi] = ia;
                    }
                }
                __syncthreads();
            }
        }
    }
    for (int r = threadIdx.x; r < cnt; r += blockDim.x) {
        const int id = sortit ? si[r] : bids[o0 + r];
        const int k = o0 + r;
        const int body = id < ntot ? id : id - ntot;
        if (id < ntot)
            ibeg[body] = k;
        else
            iend[body] = k;
        sx[k] = cx[body];
        sy[k] = cy[body];
        sz[k] = cz[body];
        svx[k] = cvx[body];
        svy[k] = cvy[body];
        svz[k] = cvz[body];
        srenc[k] = crenc[body];
        sbody[k] = body;
    }
}

// K9: encounter_check.f90:778-789 (ibeg/iend) and :937-949 (gather into sorted order)
__global__ void endpoint_kernel(int ntot, const int *__restrict__ sorted_id, const double *__restrict__ cx,
                                const double *__restrict__ cy, const double *__restrict__ cz,
                                const double *__restrict__ cvx, const double *__restrict__ cvy,
                                const double *__restrict__ cvz, const double *__restrict__ crenc, int *__restrict__ ibeg,
                                int *__restrict__ iend, double *__restrict__ sx, double *__restrict__ sy,
                                double *__restrict__ sz, double *__restrict__ svx, double *__restrict__ svy,
                                double *__restrict__ svz, double *__restrict__ srenc, int *__restrict__ sbody)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 2 * ntot) return;
    const int id = sorted_id[k];
    const int body = id < ntot ? id : id - ntot;
    if (id < ntot)
        ibeg[body] = k;
    else
        iend[body] = k;
    sx[k] = cx[body];
    sy[k] = cy[body];
    sz[k] = cz[body];
    svx[k] = cvx[body];
    svy[k] = cvy[body];
    svz[k] = cvz[body];
    srenc[k] = crenc[body];
    sbody[k] = body;
}

// loverlap (:828,:951) and the number of sweep chunks of every body; nchunk has ntot+1 entries (last = 0)
__global__ void chunk_count_kernel(int ntot, const int *__restrict__ ibeg, const int *__restrict__ iend,
                                   int *__restrict__ nchunk, unsigned long long *__restrict__ nbox_total)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    long long nb = 0;
    if (i < ntot) {
        const int b = ibeg[i], e = iend[i];
        if ((b + 1) < (e - 1)) nb = e - b - 1;  // endpoints b+1 .. e-1
        nchunk[i] = (int)((nb + SWEEP_CHUNK - 1) / SWEEP_CHUNK);
    } else if (i == ntot) {
        nchunk[i] = 0;
    }
    for (int o = 16; o > 0; o >>= 1) nb += __shfl_down_sync(0xffffffffu, nb, o);
    if ((threadIdx.x & 31) == 0 && nb > 0) atomicAdd(nbox_total, (unsigned long long)nb);
}

// chunk -> body map for the sweep (bodies own consecutive chunk ids from the prefix sum); chunks beyond the capacity of
// the map fall back to a binary search in the sweep kernel
__global__ void chunk_owner_kernel(int ntot, const int *__restrict__ nchunk, const int *__restrict__ choff,
                                   int *__restrict__ owner, int cap, unsigned long long *__restrict__ total_chunks)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *total_chunks = (unsigned long long)choff[ntot];  // read back with the other counters in one copy
    if (i >= ntot) return;
    const int c0 = choff[i], nc = nchunk[i];
    for (int q = 0; q < nc && c0 + q < cap; ++q) owner[c0 + q] = i;
}

// encounter_check_one, encounter_check.f90:591-618.
// The exact expressions (two IEEE divisions) are only evaluated for pairs that can possibly be an encounter: for
// vdotr <= 0 both branches of r2min satisfy r2min >= r2 + 2*vdotr*dt (tmin < dt implies vdotr^2/v2 < -vdotr*dt), so
// r2 + 2*vdotr*dt > r2crit + 1e-9*r2 proves "no encounter": every quantity involved is computed to a few ulp of r2
// (the largest term), and the margin is a million times that whatever the ratio r2 / r2crit (the all-pairs checks see
// pairs a thousand encounter radii apart).  The decision is therefore identical to the reference expression.
__device__ __forceinline__ bool check_one(double xr, double yr, double zr, double vxr, double vyr, double vzr,
                                          double renc, double dt, double vsmall)
{
    const double r2 = xr * xr + yr * yr + zr * zr;
    const double r2crit = renc * renc;
    if (!(r2 > r2crit)) return true;  // vdotr = -1, r2min = r2 <= r2crit  (:612-615)
    const double vdotr = vxr * xr + vyr * yr + vzr * zr;
    if (vdotr > 0.0) return false;    // lvdotr false (:596-597, :617)
    if (r2 + 2.0 * vdotr * dt > r2crit + 1e-9 * r2) return false;  // conservative: cannot come close enough
    double r2min;
    const double v2 = vxr * vxr + vyr * vyr + vzr * vzr;
    if (v2 <= vsmall) {
        r2min = r2;
    } else {
        const double tmin = -vdotr / v2;
        if (tmin < dt)
            r2min = r2 - vdotr * vdotr / v2;
        else
            r2min = r2 + 2 * vdotr * dt + v2 * (dt * dt);
    }
    const bool lvdotr = (vdotr < 0.0);
    return lvdotr && (r2min <= r2crit);
}

// K10: the sweep.  One warp per chunk of one body's interval.
__global__ void __launch_bounds__(128) sweep_kernel(int ntot, int n1, int single, const int *__restrict__ choff,
                                                    const int *__restrict__ owner, int owner_cap,
                                                    const int *__restrict__ ibeg, const int *__restrict__ iend,
                                                    const double *__restrict__ cx, const double *__restrict__ cy,
                                                    const double *__restrict__ cz, const double *__restrict__ cvx,
                                                    const double *__restrict__ cvy, const double *__restrict__ cvz,
                                                    const double *__restrict__ crenc, const double *__restrict__ sx,
                                                    const double *__restrict__ sy, const double *__restrict__ sz,
                                                    const double *__restrict__ svx, const double *__restrict__ svy,
                                                    const double *__restrict__ svz, const double *__restrict__ srenc,
                                                    const int *__restrict__ sbody, double dt, double vsmall,
                                                    unsigned long long *__restrict__ cand, unsigned long long cap,
                                                    unsigned long long *__restrict__ count, int b2)
{
    // keys are packed (index1 << b2) | index2 with b2 = bits of the largest index2: the key sort then runs over
    // bits(index1) + b2 bits (34 at npl = 1e5) instead of 64
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int total = choff[ntot];
    for (int c = warp; c < total; c += nwarps) {
        // body i with choff[i] <= c < choff[i+1]
        int i;
        if (c < owner_cap) {
            i = owner[c];
        } else {
            int lo = 0, hi = ntot - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (choff[mid + 1] <= c)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            i = lo;
        }
        const int kb = ibeg[i] + 1 + (c - choff[i]) * SWEEP_CHUNK;
        const int ke = min(kb + SWEEP_CHUNK, iend[i]);
        const double xi = cx[i], yi = cy[i], zi = cz[i];
        const double vxi = cvx[i], vyi = cvy[i], vzi = cvz[i];
        const double renci = crenc[i];
        const bool in1 = i < n1;
        // two candidates per lane per iteration: their eight streams of loads are issued together
        for (int k0 = kb; k0 < ke; k0 += 64) {
            bool hit[2];
            unsigned long long key[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int k = k0 + 32 * u + lane;
                hit[u] = false;
                key[u] = 0ull;
                if (k < ke) {
                    const int j = sbody[k];
                    bool good = true;
                    if (!single) good = ((j < n1) != in1);  // only bodies of the other list (:873,:890)
                    if (good) {
                        const double xr = sx[k] - xi, yr = sy[k] - yi, zr = sz[k] - zi;
                        const double vxr = svx[k] - vxi, vyr = svy[k] - vyi, vzr = svz[k] - vzi;
                        const double renc12 = renci + srenc[k];
                        hit[u] = check_one(xr, yr, zr, vxr, vyr, vzr, renc12, dt, vsmall);
                        if (hit[u]) {
                            unsigned a, b;
                            if (single) {  // :976-983 index1 < index2
                                a = (unsigned)min(i, j) + 1u;
                                b = (unsigned)max(i, j) + 1u;
                            } else if (in1) {  // index1 = list-1 body, index2 = list-2 body
                                a = (unsigned)i + 1u;
                                b = (unsigned)(j - n1) + 1u;
                            } else {
                                a = (unsigned)j + 1u;
                                b = (unsigned)(i - n1) + 1u;
                            }
                            key[u] = ((unsigned long long)a << b2) | b;
                        }
                    }
                }
            }
            const unsigned m0 = __ballot_sync(0xffffffffu, hit[0]);
            const unsigned m1 = __ballot_sync(0xffffffffu, hit[1]);
            if (m0 | m1) {
                unsigned long long base = 0ull;
                const int n0 = __popc(m0);
                if (lane == 0) base = atomicAdd(count, (unsigned long long)(n0 + __popc(m1)));
                base = __shfl_sync(0xffffffffu, base, 0);
                const unsigned below = (1u << lane) - 1u;
                if (hit[0]) {
                    const unsigned long long pos = base + __popc(m0 & below);
                    if (pos < cap) cand[pos] = key[0];
                }
                if (hit[1]) {
                    const unsigned long long pos = base + n0 + __popc(m1 & below);
                    if (pos < cap) cand[pos] = key[1];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// pl-tp sweep without the sort (few massive bodies, many test particles: BASELINE configs[1], 8 planets + 1e6 tp).
//
// A test particle has renc = 0, so its two endpoints carry the same key K = |r_tp| (:305-319) and it never sweeps from
// its own side; what the 2(npl+ntp)-key sort decides is only WHICH tp endpoints lie strictly between the two endpoints
// of planet i in the stably sorted sequence.  With ties resolved by position in the [rmin(1:ntot), rmax(1:ntot)] array
// (planet begin i < tp begin npl+q < planet end ntot+i < tp end ntot+npl+q) that set is known without sorting:
//     tp begin inside  <=>  rmin_i <= K <= rmax_i          tp end inside  <=>  rmin_i <= K <  rmax_i
// so nbox_i = (#tp with rmin_i <= K <= rmax_i) + (#tp with rmin_i <= K < rmax_i) + (planet endpoints inside, counted
// with the same tie rule), planet i is swept iff nbox_i >= 2 (F3, :828), and its candidates are those tp -- twice each in
// the reference's ragged list, once after remove_duplicates.  One pass over the particles replaces the radix sort of
// 2e6 (key, id) pairs, the gather into sorted order and the chunked sweep: 48 B per particle instead of ~400.
// The one situation in which the sorted sequence holds more than this -- a particle whose |r| EQUALS a planet's rmax
// bit for bit (then the particle's own degenerate interval contains that planet's end point and the reference may sweep
// from the particle's side, and the planet sees one of the particle's endpoints only) -- raises a flag and the call
// falls back to the sort path, as do NaN extents.  nbox_total counts the planets' boxes; three or more particles with
// bit-identical |r| would add their (pairless) boxes in the reference's count, not here.
struct PlRec {
    double rmin, rmax, x, y, z, vx, vy, vz, renc;
};

constexpr int PLTP_T = 256;               // threads per CTA
constexpr int PLTP_MAXPL = 128;

// encounter_check_one (:591-618) with ONE divergent branch: the three early exits of check_one are evaluated as
// predicates (a handful of multiplies), only pairs that survive them -- approaching, outside renc, and not excluded by the
// conservative bound -- take the path with the two IEEE divisions.  Same decisions as check_one for every input.
__device__ __forceinline__ bool check_one_flat(double xr, double yr, double zr, double vxr, double vyr, double vzr,
                                               double renc, double dt, double vsmall)
{
    const double r2 = xr * xr + yr * yr + zr * zr;
    const double r2crit = renc * renc;
    const double vdotr = vxr * xr + vyr * yr + vzr * zr;
    const bool inside = !(r2 > r2crit);                                            // (:612-615): encounter
    bool hit = inside;
    if (!inside && !(vdotr > 0.0) && !(r2 + 2.0 * vdotr * dt > r2crit + 1e-9 * r2)) {
        double r2min;
        const double v2 = vxr * vxr + vyr * vyr + vzr * vzr;
        if (v2 <= vsmall) {
            r2min = r2;
        } else {
            const double tmin = -vdotr / v2;
            if (tmin < dt)
                r2min = r2 - vdotr * vdotr / v2;
            else
                r2min = r2 + 2 * vdotr * dt + v2 * (dt * dt);
        }
        hit = (vdotr < 0.0) && (r2min <= r2crit);
    }
    return hit;
}

// shared memory of pltp_direct_kernel<PER>: per planet -- record, 3 counters, hits per (slot, warp), hit bits per thread;
// fixed -- the chunk's coordinates (6 doubles per particle) and the candidate queue (PLTP_GROUP entries per particle)
constexpr int PLTP_GROUP = 8;  // planets per candidate round
template <int PER>
constexpr size_t pltp_shmem_per_planet()
{
    return sizeof(PlRec) + 3 * sizeof(unsigned int) + sizeof(unsigned short) * PER * (PLTP_T / 32) + PLTP_T;
}
template <int PER>
constexpr size_t pltp_shmem_fixed()
{
    return (size_t)PLTP_T * PER * (6 * sizeof(double) + PLTP_GROUP * sizeof(unsigned int));
}

// Pass A.  Persistent CTAs; a CTA takes the PER*T particles [b*CHUNK, (b+1)*CHUNK) of chunk b, thread t brings particles
// b*CHUNK + u*T + t, u < PER (coalesced loads, the next chunk prefetched into registers while this one is evaluated).
//   1. every thread tests its particles' |r| against the extents of a group of planets (two compares per pair) and pushes
//      the pairs inside an interval onto a queue in shared memory;
//   2. the queue is evaluated DENSELY -- lane k takes entry k, whatever particle and planet it names -- so the predicate
//      runs on full warps (evaluating it thread-per-particle ran it for every warp in which one lane was inside an
//      interval: 7 times the instructions, ncu in profiles/r02_pltp_direct_ncu.txt); hits set a bit per (planet, particle);
//   3. from the bits: hits per (planet, slot, warp), their prefix in particle order, and the hits leave the CTA ORDERED by
//      (planet, particle) into a region of the arena claimed with one atomic.
// cnt[i*nb + b] = hits of planet i in chunk b, box[i*nb + b] = endpoints of the chunk inside planet i's interval (CTA 0
// adds the planets' endpoints inside each other's interval to chunk 0), abase[b] = where the chunk's region starts.
template <int PER>
__global__ void __launch_bounds__(PLTP_T, 3) pltp_direct_kernel(ListDev pl, ListDev tp, int nb, double dt, double vsmall,
                                                                unsigned long long *__restrict__ arena,
                                                                unsigned long long cap,
                                                                unsigned long long *__restrict__ count,
                                                                unsigned int *__restrict__ box, int *__restrict__ flag,
                                                                int *__restrict__ cnt, unsigned long long *__restrict__ abase)
{
    constexpr int NW = PLTP_T / 32, CHUNK = PLTP_T * PER;
    extern __shared__ unsigned char sh_raw[];
    const int n1 = pl.n, n2 = tp.n;
    PlRec *spl = reinterpret_cast<PlRec *>(sh_raw);                      // n1 records
    double *scx = reinterpret_cast<double *>(spl + n1);                  // the chunk: x, y, z, vx, vy, vz
    double *scy = scx + CHUNK, *scz = scy + CHUNK, *scvx = scz + CHUNK, *scvy = scvx + CHUNK, *scvz = scvy + CHUNK;
    unsigned int *sq = reinterpret_cast<unsigned int *>(scvz + CHUNK);   // candidate queue: (planet << 16) | slot
    unsigned int *sbox = sq + CHUNK * PLTP_GROUP;                        // n1: endpoints of this chunk inside planet i
    unsigned int *shit = sbox + n1;                                      // n1: hits of planet i in this chunk
    unsigned int *soff = shit + n1;                                      // n1: start of planet i's run in the region
    unsigned short *swtot = reinterpret_cast<unsigned short *>(soff + n1);  // n1 x PER x NW: hits per (slot, warp),
                                                                            // later their exclusive prefix per planet
    unsigned char *shb = reinterpret_cast<unsigned char *>(swtot + n1 * PER * NW);  // n1 x T: hit bits per thread
    unsigned int *shbw = reinterpret_cast<unsigned int *>(shb);
    __shared__ unsigned long long s_base;
    __shared__ unsigned int s_qn[2];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

    for (int i = t; i < n1; i += PLTP_T) {
        const double x = pl.x[i], y = pl.y[i], z = pl.z[i];
        const double renc = pl.renc ? pl.renc[i] : 0.0;
        const double rmag = sqrt(x * x + y * y + z * z);
        const double w = RSWEEP_FACTOR * renc;
        PlRec r;
        r.rmin = rmag - w;
        r.rmax = rmag + w;
        r.x = x, r.y = y, r.z = z, r.vx = pl.vx[i], r.vy = pl.vy[i], r.vz = pl.vz[i], r.renc = renc;
        spl[i] = r;
        sbox[i] = 0u;
        shit[i] = 0u;
        if (blockIdx.x == 0 && (r.rmin != r.rmin || r.rmax != r.rmax)) atomicOr(flag, 1);
    }
    if (t < 2) s_qn[t] = 0u;
    double nx[PER], ny[PER], nz[PER], nvx[PER], nvy[PER], nvz[PER];
    auto fetch = [&](int chunk) {
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const long long q = (long long)chunk * CHUNK + u * PLTP_T + t;
            nx[u] = ny[u] = nz[u] = nvx[u] = nvy[u] = nvz[u] = 0.0;
            if (chunk < nb && q < n2) {
                nx[u] = tp.x[q], ny[u] = tp.y[q], nz[u] = tp.z[q];
                nvx[u] = tp.vx[q], nvy[u] = tp.vy[q], nvz[u] = tp.vz[q];
            }
        }
    };
    fetch(blockIdx.x);
    const unsigned below = (1u << lane) - 1u;
    unsigned round = 0u;
    for (int chunk = blockIdx.x; chunk < nb; chunk += gridDim.x) {
        const long long q0 = (long long)chunk * CHUNK + t;
        double K[PER];
        bool valid[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int sl = u * PLTP_T + t;
            scx[sl] = nx[u], scy[sl] = ny[u], scz[sl] = nz[u], scvx[sl] = nvx[u], scvy[sl] = nvy[u], scvz[sl] = nvz[u];
            valid[u] = q0 + u * PLTP_T < n2;
            K[u] = sqrt(nx[u] * nx[u] + ny[u] * ny[u] + nz[u] * nz[u]);  // rmin = rmax = |r| -/+ 1.1 * 0
        }
        fetch(chunk + gridDim.x);
        for (int w = t; w < n1 * (PLTP_T / 4); w += PLTP_T) shbw[w] = 0u;
        __syncthreads();  // planet records, zeroed counters and bits, the chunk in shared memory
        if (chunk == 0) {  // the planets' own endpoints inside each other's interval, same tie rule as the stable sort
            for (int i = t; i < n1; i += PLTP_T) {
                const double lo = spl[i].rmin, hi = spl[i].rmax;
                unsigned int c = 0;
                for (int j = 0; j < n1; ++j) {
                    if (j == i) continue;
                    const double bj = spl[j].rmin, ej = spl[j].rmax;
                    // begin of j sits at array position j, end of j at ntot + j; i's own endpoints at i and ntot + i
                    if ((bj > lo || (bj == lo && j > i)) && bj <= hi) ++c;
                    if (ej >= lo && (ej < hi || (ej == hi && j < i))) ++c;
                }
                if (c) atomicAdd(&sbox[i], c);
            }
        }
        for (int g0 = 0; g0 < n1; g0 += PLTP_GROUP, ++round) {
            unsigned int *qn = &s_qn[round & 1u];
            const int g1 = min(n1, g0 + PLTP_GROUP);
            for (int i = g0; i < g1; ++i) {
                const double rmin = spl[i].rmin, rmax = spl[i].rmax;
#pragma unroll
                for (int u = 0; u < PER; ++u) {
                    if (valid[u] && K[u] >= rmin && K[u] <= rmax) {  // begin endpoint inside planet i's interval
                        const bool in_e = K[u] < rmax;               // end endpoint inside
                        atomicAdd(&sbox[i], in_e ? 2u : 1u);
                        if (!in_e) atomicOr(flag, 2);                // |r_tp| == rmax_i: the sort path decides
                        sq[atomicAdd(qn, 1u)] = ((unsigned)i << 16) | (unsigned)(u * PLTP_T + t);
                    }
                }
            }
            __syncthreads();
            const unsigned nq = *qn;
            if (t == 0) s_qn[(round + 1u) & 1u] = 0u;  // the other counter is idle until the barrier below
            for (unsigned e = t; e < nq; e += PLTP_T) {
                const unsigned code = sq[e];
                const int i = (int)(code >> 16), sl = (int)(code & 0xffffu);
                const PlRec &p = spl[i];
                if (check_one_flat(scx[sl] - p.x, scy[sl] - p.y, scz[sl] - p.z, scvx[sl] - p.vx, scvy[sl] - p.vy,
                                   scvz[sl] - p.vz, p.renc + 0.0, dt, vsmall)) {
                    const int th = sl % PLTP_T, u = sl / PLTP_T;
                    atomicOr(&shbw[(i * PLTP_T + th) >> 2], 1u << (((th & 3) << 3) + u));
                }
            }
            __syncthreads();
        }
        // hits per (planet, slot, warp) from the hit bits, their exclusive prefix in particle order, the planet's total
        for (int i = warp; i < n1; i += NW) {
            unsigned run = 0u;
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                for (int w = 0; w < NW; ++w) {
                    const unsigned c = __popc(__ballot_sync(0xffffffffu, (shb[i * PLTP_T + w * 32 + lane] >> u) & 1u));
                    if (lane == 0) swtot[(i * PER + u) * NW + w] = (unsigned short)run;
                    run += c;
                }
            }
            if (lane == 0) shit[i] = run;
        }
        __syncthreads();
        if (warp == 0) {  // runs of the planets inside this chunk's region; one claim of the arena; counts for pass B
            unsigned int carry = 0u;
            for (int i0 = 0; i0 < n1; i0 += 32) {
                const int i = i0 + lane;
                const unsigned int c = i < n1 ? shit[i] : 0u;
                unsigned int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                if (i < n1) {
                    soff[i] = carry + incl - c;
                    cnt[(size_t)i * nb + chunk] = (int)c;
                    box[(size_t)i * nb + chunk] = sbox[i];  // summed by pass B (15 000 atomics on one line cost 40 us)
                }
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) {
                unsigned long long base = 0ull;
                if (carry) base = atomicAdd(count, (unsigned long long)carry);
                abase[chunk] = base;
                s_base = base;
            }
        }
        __syncthreads();
        const unsigned long long base = s_base;
        for (int i = 0; i < n1; ++i) {
            if (shit[i] == 0u) continue;
            const unsigned hb = shb[i * PLTP_T + t];
            if (!__any_sync(0xffffffffu, hb != 0u)) continue;
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const unsigned m = __ballot_sync(0xffffffffu, (hb >> u) & 1u);
                if ((hb >> u) & 1u) {
                    const unsigned long long pos = base + soff[i] + swtot[(i * PER + u) * NW + warp] + __popc(m & below);
                    if (pos < cap)
                        arena[pos] = ((unsigned long long)(i + 1) << 32) | (unsigned long long)(q0 + u * PLTP_T + 1);
                }
            }
        }
        __syncthreads();  // everybody is done with this chunk's counters
        for (int i = t; i < n1; i += PLTP_T) sbox[i] = 0u, shit[i] = 0u;
    }
}

// Pass B: one warp per chunk moves the chunk's runs to their place in the canonical list: planet-major, chunks in order
// (offs = exclusive scan of cnt), already ascending in the particle index inside a chunk.  CTAs beyond the chunks add up
// the box counts of one planet each.
__global__ void __launch_bounds__(256) pltp_place_kernel(int n1, int nb, int nplace, const int *__restrict__ cnt,
                                                         const int *__restrict__ offs,
                                                         const unsigned long long *__restrict__ abase,
                                                         const unsigned long long *__restrict__ arena,
                                                         unsigned long long cap, unsigned long long *__restrict__ out,
                                                         const unsigned int *__restrict__ box,
                                                         unsigned long long *__restrict__ nbc)
{
    const int lane = threadIdx.x & 31;
    if ((int)blockIdx.x >= nplace) {
        __shared__ unsigned long long part[8];
        const int i = blockIdx.x - nplace;
        unsigned long long sum = 0ull;
        for (int b = threadIdx.x; b < nb; b += blockDim.x) sum += box[(size_t)i * nb + b];
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) part[threadIdx.x >> 5] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) sum += part[w];
            nbc[i] = sum;
            if (i == 0) nbc[-3] = (unsigned long long)offs[(size_t)n1 * nb];  // counters[5]: hits placed (read back with the rest)
        }
        return;
    }
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= nb) return;
    unsigned long long src = abase[b];
    for (int i0 = 0; i0 < n1; i0 += 32) {
        const int i = i0 + lane;
        const int c = i < n1 ? cnt[(size_t)i * nb + b] : 0;
        const int d = i < n1 ? offs[(size_t)i * nb + b] : 0;
        if (!__any_sync(0xffffffffu, c != 0)) continue;
        for (int j = 0; j < 32 && i0 + j < n1; ++j) {
            const int cj = __shfl_sync(0xffffffffu, c, j);
            const int dj = __shfl_sync(0xffffffffu, d, j);
            for (int k = lane; k < cj; k += 32)
                if (src + k < cap && (unsigned long long)dj + k < cap) out[(size_t)dj + k] = arena[src + k];  // overflow: rerun
            src += (unsigned long long)cj;
        }
    }
}

// K11 for short candidate lists, on the device and without a host round trip in the middle: one CTA sorts up to
// CAND_SMALL keys in shared memory (bitonic), drops equal neighbours and leaves the list and its length where
// canonical_order would; longer lists set flag = 1 and the host runs the radix sort + unique after all.
constexpr int CAND_SMALL = 4096;
__global__ void __launch_bounds__(1024) finalize_small_kernel(const unsigned long long *__restrict__ cand,
                                                              unsigned long long cap,
                                                              const unsigned long long *__restrict__ count,
                                                              unsigned long long *__restrict__ uniq, int *__restrict__ nuniq,
                                                              unsigned long long *__restrict__ flag, int b2)
{
    __shared__ unsigned long long sk[CAND_SMALL];
    __shared__ int part[1024];
    const unsigned long long n64 = *count;
    const int t = threadIdx.x;
    if (n64 > (unsigned long long)CAND_SMALL || n64 > cap) {
        if (t == 0) *flag = 1ull, *nuniq = 0;
        return;
    }
    const int n = (int)n64;
    int m = 1;
    while (m < n) m <<= 1;
    for (int k = t; k < m; k += 1024) sk[k] = k < n ? cand[k] : ~0ull;
    __syncthreads();
    for (int size = 2; size <= m; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int k = t; k < (m >> 1); k += 1024) {
                const int lo = ((k & ~(stride - 1)) << 1) | (k & (stride - 1)), hi = lo + stride;  // stride is a power of 2
                const unsigned long long a = sk[lo], b = sk[hi];
                if ((a > b) == ((lo & size) == 0)) sk[lo] = b, sk[hi] = a;
            }
            __syncthreads();
        }
    // heads of runs of equal keys, in the CAND_SMALL / 1024 = 4 consecutive slots of this thread
    constexpr int PER = CAND_SMALL / 1024;
    int heads = 0;
    for (int q = 0; q < PER; ++q) {
        const int k = t * PER + q;
        heads += (k < n && (k == 0 || sk[k] != sk[k - 1])) ? 1 : 0;
    }
    part[t] = heads;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = t >= o ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int pos = part[t] - heads;
    for (int q = 0; q < PER; ++q) {
        const int k = t * PER + q;
        if (k < n && (k == 0 || sk[k] != sk[k - 1])) uniq[pos++] = ((sk[k] >> b2) << 32) | (sk[k] & ((1ull << b2) - 1ull));
    }
    if (t == 1023) *nuniq = part[1023], *flag = 0ull;
}

__global__ void unpack_keys_kernel(const unsigned long long *__restrict__ keys, long long n, int32_t *__restrict__ i1,
                                   int32_t *__restrict__ i2)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    i1[k] = (int32_t)(keys[k] >> 32);
    i2[k] = (int32_t)(keys[k] & 0xffffffffull);
}

__global__ void shift_index2_kernel(unsigned long long *keys, long long n, unsigned long long shift)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) keys[k] += shift;
}

__global__ void set_renc_kernel(int n, const double *__restrict__ rhill, double rshell_irec, double *__restrict__ renc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) renc[i] = rhill[i] * RHSCALE * rshell_irec;
}

// ---------------------------------------------------------------------------------------------------------------------
// Triangular (all-pairs) variants, encounter_check_all_triangular_plpl / _plplm / _pltp (:436-570) with
// encounter_check_all_triangular_one (:384-433): the predicate on every pair, no broad phase.  One row body per thread,
// column bodies staged in shared memory TRI_T at a time, hits appended with a warp-aggregated atomic; the canonical
// (index1, index2) order comes from the same key sort as the sweep.  O(n1*n2) FP64 compare work: this is the
// `ENCOUNTER_CHECK TRIANGULAR` option of the reference and the superset checker for the sweep.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TRI_T = 128;
constexpr int TRI_COLS = 2048;

__global__ void __launch_bounds__(TRI_T) tri_check_kernel(ListDev a, ListDev b, int single, double dt, double vsmall,
                                                          unsigned long long *__restrict__ cand, unsigned long long cap,
                                                          unsigned long long *__restrict__ count)
{
    __shared__ double sx[TRI_T], sy[TRI_T], sz[TRI_T], svx[TRI_T], svy[TRI_T], svz[TRI_T], sr[TRI_T];
    const ListDev &c = single ? a : b;
    const int row0 = blockIdx.x * TRI_T;
    const int i = row0 + threadIdx.x;
    const bool ion = i < a.n;
    const int ic = ion ? i : a.n - 1;
    const double xi = a.x[ic], yi = a.y[ic], zi = a.z[ic], vxi = a.vx[ic], vyi = a.vy[ic], vzi = a.vz[ic];
    const double renci = a.renc ? a.renc[ic] : 0.0;
    const int c0 = blockIdx.y * TRI_COLS, c1 = min(c.n, c0 + TRI_COLS);
    const int lane = threadIdx.x & 31;
    for (int t0 = c0; t0 < c1; t0 += TRI_T) {
        if (single && t0 + TRI_T - 1 <= row0) continue;  // every column of this tile has j <= every row of the CTA
        __syncthreads();
        {
            const int j = min(t0 + (int)threadIdx.x, c.n - 1);
            sx[threadIdx.x] = c.x[j];
            sy[threadIdx.x] = c.y[j];
            sz[threadIdx.x] = c.z[j];
            svx[threadIdx.x] = c.vx[j];
            svy[threadIdx.x] = c.vy[j];
            svz[threadIdx.x] = c.vz[j];
            sr[threadIdx.x] = c.renc ? c.renc[j] : 0.0;
        }
        __syncthreads();
        const int jn = min(TRI_T, c1 - t0);
        for (int jj = 0; jj < jn; ++jj) {
            const int j = t0 + jj;
            bool hit = false;
            if (ion && (!single || j > i)) {
                const double xr = sx[jj] - xi, yr = sy[jj] - yi, zr = sz[jj] - zi;
                const double vxr = svx[jj] - vxi, vyr = svy[jj] - vyi, vzr = svz[jj] - vzi;
                const double renc12 = renci + sr[jj];
                hit = check_one(xr, yr, zr, vxr, vyr, vzr, renc12, dt, vsmall);
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) {
                unsigned long long base = 0;
                if (lane == __ffs(m) - 1) base = atomicAdd(count, (unsigned long long)__popc(m));
                base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                if (hit) {
                    const unsigned long long slot = base + __popc(m & ((1u << lane) - 1u));
                    if (slot < cap) cand[slot] = ((unsigned long long)(unsigned)(i + 1) << 32) | (unsigned)(j + 1);
                }
            }
        }
    }
}

// swiftest_discard_pl_close (swiftest/swiftest_discard.f90:295-337)
__device__ __forceinline__ int discard_pl_close(double dx, double dy, double dz, double dvx, double dvy, double dvz,
                                                double dt, double r2crit)
{
    const double r2 = dx * dx + dy * dy + dz * dz;
    if (r2 <= r2crit) return 1;
    const double vdotr = dx * dvx + dy * dvy + dz * dvz;
    if (vdotr > 0.0) return 0;
    // Exact conservative pre-rejection (as in check_one): for vdotr <= 0 both branches below satisfy
    // min(r2min, r2) >= r2 + 2*vdotr*dt, so a pair whose bound clears r2crit by 1e-9 of r2 -- a million times the rounding
    // of either side, whatever the ratio r2 / r2crit -- is kept without the two IEEE divisions.
    if (r2 + 2.0 * vdotr * dt > r2crit + 1e-9 * r2) return 0;
    const double v2 = dvx * dvx + dvy * dvy + dvz * dvz;
    const double tmin = -vdotr / v2;
    double r2min;
    if (tmin < dt)
        r2min = r2 - vdotr * vdotr / v2;
    else
        r2min = r2 + 2 * vdotr * dt + v2 * (dt * dt);
    r2min = (r2min < r2) ? r2min : r2;  // min(r2min, r2); a NaN (v2 == 0) falls back to r2 like the intrinsic
    return (r2min <= r2crit) ? 1 : 0;
}

// swiftest_discard_pl_tp (swiftest_discard.f90:244-292): for every active test particle the first planet, in ascending
// index order, that it is or will be too close to within dt.  iplanet = 1-based planet index or 0.
__global__ void __launch_bounds__(128) discard_pl_tp_kernel(int ntp, int npl, const double *__restrict__ tx,
                                                            const double *__restrict__ ty, const double *__restrict__ tz,
                                                            const double *__restrict__ tvx, const double *__restrict__ tvy,
                                                            const double *__restrict__ tvz, const int32_t *__restrict__ lactive,
                                                            const double *__restrict__ px, const double *__restrict__ py,
                                                            const double *__restrict__ pz, const double *__restrict__ pvx,
                                                            const double *__restrict__ pvy, const double *__restrict__ pvz,
                                                            const double *__restrict__ radius, double dt,
                                                            int32_t *__restrict__ iplanet, int *__restrict__ ndiscard)
{
    __shared__ double sx[128], sy[128], sz[128], svx[128], svy[128], svz[128], sr2[128];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = i < ntp && lactive[i] != 0;
    const int ic = min(i, ntp - 1);
    const double xi = tx[ic], yi = ty[ic], zi = tz[ic], vxi = tvx[ic], vyi = tvy[ic], vzi = tvz[ic];
    int found = 0;
    for (int t0 = 0; t0 < npl; t0 += 128) {
        if (__syncthreads_and(found != 0 || !on)) break;  // nobody in this CTA is still looking
        {
            const int j = min(t0 + (int)threadIdx.x, npl - 1);
            sx[threadIdx.x] = px[j];
            sy[threadIdx.x] = py[j];
            sz[threadIdx.x] = pz[j];
            svx[threadIdx.x] = pvx[j];
            svy[threadIdx.x] = pvy[j];
            svz[threadIdx.x] = pvz[j];
            const double r = radius[j];
            sr2[threadIdx.x] = r * r;
        }
        __syncthreads();
        const int jn = min(128, npl - t0);
        if (on && found == 0) {
            for (int jj = 0; jj < jn; ++jj) {
                if (discard_pl_close(xi - sx[jj], yi - sy[jj], zi - sz[jj], vxi - svx[jj], vyi - svy[jj], vzi - svz[jj], dt,
                                     sr2[jj])) {
                    found = t0 + jj + 1;
                    break;
                }
            }
        }
    }
    if (i < ntp) {
        iplanet[i] = found;
        if (found) atomicAdd(ndiscard, 1);
    }
}

// encounter_check_one with both outputs (lencounter, lvdotr), :573-621
__device__ __forceinline__ bool check_one_full(double xr, double yr, double zr, double vxr, double vyr, double vzr,
                                               double renc, double dt, double vsmall, bool &lvdotr)
{
    const double r2 = xr * xr + yr * yr + zr * zr;
    const double r2crit = renc * renc;
    if (!(r2 > r2crit)) {  // vdotr = -1, r2min = r2
        lvdotr = true;
        return true;
    }
    const double vdotr = vxr * xr + vyr * yr + vzr * zr;
    lvdotr = (vdotr < 0.0);
    if (vdotr > 0.0) return false;
    double r2min;
    const double v2 = vxr * vxr + vyr * vyr + vzr * vzr;
    if (v2 <= vsmall) {
        r2min = r2;
    } else {
        const double tmin = -vdotr / v2;
        if (tmin < dt)
            r2min = r2 - vdotr * vdotr / v2;
        else
            r2min = r2 + 2 * vdotr * dt + v2 * (dt * dt);
    }
    return lvdotr && (r2min <= r2crit);
}

// the pair loop of symba_encounter_check_list_plpl / _pltp (symba/symba_encounter_check.f90:122-137, 197-211)
__global__ void symba_check_list_kernel(long long nenc, const int32_t *__restrict__ index1,
                                        const int32_t *__restrict__ index2, const int32_t *__restrict__ lencmask, ListDev a,
                                        const double *__restrict__ radius1, ListDev b, const double *__restrict__ radius2,
                                        double dt, double vsmall, int32_t *__restrict__ lencounter,
                                        int32_t *__restrict__ lvdotr, unsigned long long *__restrict__ count)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nenc) return;
    if (lencmask && lencmask[k] == 0) {
        lencounter[k] = 0;
        return;
    }
    const int i = index1[k] - 1, j = index2[k] - 1;
    const double xr = b.x[j] - a.x[i], yr = b.y[j] - a.y[i], zr = b.z[j] - a.z[i];
    const double vxr = b.vx[j] - a.vx[i], vyr = b.vy[j] - a.vy[i], vzr = b.vz[j] - a.vz[i];
    const double rcrit12 = a.renc[i] + (b.renc ? b.renc[j] : 0.0);
    bool lvd;
    bool lenc = check_one_full(xr, yr, zr, vxr, vyr, vzr, rcrit12, dt, vsmall, lvd);
    lvdotr[k] = lvd ? 1 : 0;
    if (lenc) {  // physically overlapping bodies are ignored (:131-135)
        const double rl = radius1[i] + (radius2 ? radius2[j] : 0.0);
        const double rlim2 = rl * rl;
        const double rji2 = xr * xr + yr * yr + zr * zr;
        lenc = rji2 > rlim2;
    }
    lencounter[k] = lenc ? 1 : 0;
    if (lenc) atomicAdd(count, 1ull);
}

ListDev to_dev(const SweepList &l)
{
    ListDev d;
    d.x = l.x; d.y = l.y; d.z = l.z; d.vx = l.vx; d.vy = l.vy; d.vz = l.vz; d.renc = l.renc; d.n = l.n;
    return d;
}

}  // namespace

int set_renc(swcu_context *ctx, Body &pl, int irec)
{
    if (pl.n <= 0) return SWCU_OK;
    double rshell_irec = 1.0;
    for (int i = 1; i <= irec; ++i) rshell_irec = rshell_irec * RSHELL;  // symba_util.f90:259-262
    set_renc_kernel<<<cdiv(pl.n, 256), 256, 0, ctx->stream>>>(pl.n, pl.rhill.as<double>(), rshell_irec,
                                                             pl.renc.as<double>());
    SWCU_KERNEL_CHECK(ctx);
    return SWCU_OK;
}

// Sort-and-sweep of one list (l2 == nullptr) or two lists.  Leaves the sorted unique keys in ctx->enc.uniq.
// K11 canonical order + duplicate removal (:976-985, :703-757) of the ncand keys in E.cand
static int bits_for(unsigned long long v)
{
    int b = 1;
    while ((v >> b) != 0ull) ++b;
    return b;
}

// packed (index1 << b2 | index2) keys -> the canonical (index1 << 32 | index2) keys the fetch expects, in place
__global__ void expand_keys_kernel(unsigned long long *__restrict__ keys, const int *__restrict__ n, int b2)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= *n) return;
    const unsigned long long c = keys[k];
    keys[k] = ((c >> b2) << 32) | (c & ((1ull << b2) - 1ull));
}

// keys packed with b2 < 32 are sorted over b1 + b2 bits and expanded afterwards; b2 == 32: canonical keys, all 64 bits
int canonical_order(swcu_context *ctx, long long ncand, int64_t *nenc_out, int b1 = 32, int b2 = 32)
{
    auto &E = ctx->enc;
    const int end_bit = (b2 >= 32) ? 64 : b1 + b2;
    int *d_nuniq = reinterpret_cast<int *>(E.counters.as<unsigned long long>() + 2);
    SWCU_CUDA(ctx, E.cand_sorted.ensure(sizeof(unsigned long long) * ncand));
    SWCU_CUDA(ctx, E.uniq.ensure(sizeof(unsigned long long) * ncand));
    size_t tmp_k = 0, tmp_u = 0;
    SWCU_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tmp_k, E.cand.as<unsigned long long>(),
                                                  E.cand_sorted.as<unsigned long long>(), ncand, 0, end_bit, ctx->stream));
    SWCU_CUDA(ctx, cub::DeviceSelect::Unique(nullptr, tmp_u, E.cand_sorted.as<unsigned long long>(),
                                             E.uniq.as<unsigned long long>(), d_nuniq, ncand, ctx->stream));
    SWCU_CUDA(ctx, E.cub_tmp.ensure(std::max(tmp_k, tmp_u)));
    SWCU_CUDA(ctx, cub::DeviceRadixSort::SortKeys(E.cub_tmp.p, tmp_k, E.cand.as<unsigned long long>(),
                                                  E.cand_sorted.as<unsigned long long>(), ncand, 0, end_bit, ctx->stream));
    SWCU_CUDA(ctx, cub::DeviceSelect::Unique(E.cub_tmp.p, tmp_u, E.cand_sorted.as<unsigned long long>(),
                                             E.uniq.as<unsigned long long>(), d_nuniq, ncand, ctx->stream));
    ctx->launches += 6;
    if (b2 < 32) {
        expand_keys_kernel<<<cdiv(ncand, 256), 256, 0, ctx->stream>>>(E.uniq.as<unsigned long long>(), d_nuniq, b2);
        SWCU_KERNEL_CHECK(ctx);
    }
    int h_nuniq = 0;
    SWCU_CUDA(ctx, cudaMemcpyAsync(&h_nuniq, d_nuniq, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    E.nenc = h_nuniq;
    E.result = E.uniq.as<unsigned long long>();
    *nenc_out = h_nuniq;
    return SWCU_OK;
}

// pl-tp sweep of a few massive bodies over many particles without the sort (see pltp_direct_kernel).  *fell_back is set
// when the call has to be decided by the sorted sequence after all; the caller then runs the sort path.
static int pltp_per()  // particles per thread of pltp_direct_kernel: 1, 2 (default, by measurement) or 4
{
    const char *pe = getenv("SWCU_PLTP_PER");
    const int per = pe ? atoi(pe) : 2;
    return per == 4 ? 4 : per == 1 ? 1 : 2;
}

int encounter_pltp_direct(swcu_context *ctx, const SweepList &l1, const SweepList &l2, double dt, int64_t *nenc_out,
                          bool *fell_back)
{
    auto &E = ctx->enc;
    *fell_back = false;
    const int n1 = l1.n, n2 = l2.n;
    const int nb = cdiv(n2, PLTP_T * pltp_per());
    const size_t ncnt = (size_t)n1 * nb;
    if (ncnt + 1 > (size_t)0x7fffffff) {
        *fell_back = true;
        return SWCU_OK;
    }
    FamTimer ft(ctx, FAM_SWEEP);
    // counters: [0] hits claimed in the arena, [4] flag word, [8 .. 8+n1) box counts of the planets
    SWCU_CUDA(ctx, E.counters.ensure(sizeof(unsigned long long) * (8 + PLTP_MAXPL)));
    SWCU_CUDA(ctx, E.nchunk.ensure(sizeof(int) * (ncnt + 1)));   // cnt
    SWCU_CUDA(ctx, E.choff.ensure(sizeof(int) * (ncnt + 1)));    // offs
    SWCU_CUDA(ctx, E.abase.ensure(sizeof(unsigned long long) * (size_t)nb));
    SWCU_CUDA(ctx, E.boxcnt.ensure(sizeof(unsigned int) * ncnt));
    unsigned long long *d_count = E.counters.as<unsigned long long>();
    int *d_flag = reinterpret_cast<int *>(d_count + 4);
    unsigned long long *d_nbc = d_count + 8;
    int *d_cnt = E.nchunk.as<int>(), *d_offs = E.choff.as<int>();
    size_t tmp_scan = 0;
    SWCU_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, d_cnt, d_offs, (int)ncnt + 1, ctx->stream));
    SWCU_CUDA(ctx, E.cub_tmp.ensure(tmp_scan));
    const ListDev a = to_dev(l1), b = to_dev(l2);
    if (E.cand_cap < (size_t)n2 / 4 + 65536) E.cand_cap = (size_t)n2 / 4 + 65536;
    const double vsmall = std::sqrt(DBL_MIN);  // globals_module.f90:135
    const int per = pltp_per();
    const size_t shmem = per == 4   ? n1 * pltp_shmem_per_planet<4>() + pltp_shmem_fixed<4>()
                         : per == 2 ? n1 * pltp_shmem_per_planet<2>() + pltp_shmem_fixed<2>()
                                    : n1 * pltp_shmem_per_planet<1>() + pltp_shmem_fixed<1>();
    if (!E.direct_attr_set) {  // per context: the attribute belongs to the device
        SWCU_CUDA(ctx, cudaFuncSetAttribute(pltp_direct_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(PLTP_MAXPL * pltp_shmem_per_planet<1>() + pltp_shmem_fixed<1>())));
        SWCU_CUDA(ctx, cudaFuncSetAttribute(pltp_direct_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(PLTP_MAXPL * pltp_shmem_per_planet<2>() + pltp_shmem_fixed<2>())));
        SWCU_CUDA(ctx, cudaFuncSetAttribute(pltp_direct_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(PLTP_MAXPL * pltp_shmem_per_planet<4>() + pltp_shmem_fixed<4>())));
        E.direct_attr_set = true;
    }
    const int grid = std::min(nb, ctx->prop.multiProcessorCount * 3);  // persistent CTAs, 3 per SM (launch bounds)
    if (!E.h_counters)
        SWCU_CUDA(ctx, cudaHostAlloc((void **)&E.h_counters, (8 + PLTP_MAXPL) * sizeof(unsigned long long), cudaHostAllocDefault));
    const unsigned long long *h = E.h_counters;  // pinned: one read-back of [0] hits, [4] flags, [5] hits placed, [8..) boxes
    int h_total = 0;
    for (int attempt = 0; attempt < 3; ++attempt) {
        SWCU_CUDA(ctx, E.cand.ensure(sizeof(unsigned long long) * E.cand_cap));
        SWCU_CUDA(ctx, E.uniq.ensure(sizeof(unsigned long long) * E.cand_cap));
        SWCU_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(unsigned long long) * (8 + (size_t)n1), ctx->stream));
        SWCU_CUDA(ctx, cudaMemsetAsync(d_cnt + ncnt, 0, sizeof(int), ctx->stream));
        auto *kern = per == 4 ? pltp_direct_kernel<4> : per == 2 ? pltp_direct_kernel<2> : pltp_direct_kernel<1>;
        kern<<<grid, PLTP_T, shmem, ctx->stream>>>(a, b, nb, dt, vsmall, E.cand.as<unsigned long long>(),
                                                   (unsigned long long)E.cand_cap, d_count, E.boxcnt.as<unsigned int>(),
                                                   d_flag, d_cnt, E.abase.as<unsigned long long>());
        SWCU_KERNEL_CHECK(ctx);
        SWCU_CUDA(ctx, cub::DeviceScan::ExclusiveSum(E.cub_tmp.p, tmp_scan, d_cnt, d_offs, (int)ncnt + 1, ctx->stream));
        ctx->launches += 1;
        pltp_place_kernel<<<cdiv(nb, 8) + n1, 256, 0, ctx->stream>>>(
            n1, nb, cdiv(nb, 8), d_cnt, d_offs, E.abase.as<unsigned long long>(), E.cand.as<unsigned long long>(),
            (unsigned long long)E.cand_cap, E.uniq.as<unsigned long long>(), E.boxcnt.as<unsigned int>(), d_nbc);
        SWCU_KERNEL_CHECK(ctx);
        SWCU_CUDA(ctx, cudaMemcpyAsync(E.h_counters, d_count, sizeof(unsigned long long) * (8 + (size_t)n1),
                                       cudaMemcpyDeviceToHost, ctx->stream));
        SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        h_total = (int)h[5];
        if ((int)(h[4] & 0xffffffffull) != 0) {
            *fell_back = true;
            return SWCU_OK;
        }
        if (h[0] <= E.cand_cap) break;
        E.cand_cap = (size_t)(h[0] + h[0] / 4 + 1024);  // the arena was too small: grow it and run again
        if (attempt == 2) return fail(ctx, SWCU_ERR_STATE, "pl-tp encounter check: candidate buffer overflow persists");
    }
    int64_t nbox = 0;
    for (int i = 0; i < n1; ++i)
        if (h[8 + i] >= 2ull) nbox += (int64_t)h[8 + i];  // loverlap (:828): ibeg + 1 < iend - 1
    const long long nhit = (long long)h[0];
    if ((long long)h_total != nhit) return fail(ctx, SWCU_ERR_STATE, "pl-tp encounter check: %lld hits claimed, %d placed",
                                                nhit, h_total);
    E.nbox_total = nbox;
    E.nemitted = 2 * nhit;  // the reference's ragged list holds every hit once per endpoint
    if (nhit == 0) return SWCU_OK;
    E.nenc = nhit;
    E.result = E.uniq.as<unsigned long long>();
    *nenc_out = nhit;
    return SWCU_OK;
}

static int encounter_sweep_impl(swcu_context *ctx, const SweepList &l1, const SweepList *l2, double dt, int64_t *nenc_out,
                                bool allow_direct, bool allow_bucket);

int encounter_sweep(swcu_context *ctx, const SweepList &l1, const SweepList *l2, double dt, int64_t *nenc_out)
{
    return encounter_sweep_impl(ctx, l1, l2, dt, nenc_out, true, true);
}

static int encounter_sweep_impl(swcu_context *ctx, const SweepList &l1, const SweepList *l2, double dt, int64_t *nenc_out,
                                bool allow_direct, bool allow_bucket)
{
    auto &E = ctx->enc;
    E.nenc = 0;
    E.result = nullptr;
    E.nbox_total = 0;
    E.nemitted = 0;
    *nenc_out = 0;
    const int n1 = l1.n, n2 = l2 ? l2->n : 0;
    const bool single = (l2 == nullptr);
    if (n1 == 0 || (!single && n2 == 0)) return SWCU_OK;  // :168, :225, :291
    if (allow_direct && !single && l2->renc == nullptr) {  // pl-tp with few massive bodies: no sort needed
        const char *dm = getenv("SWCU_PLTP_DIRECT_MAX");  // read per call: tests switch the path at run time
        const int direct_max = std::min(dm ? atoi(dm) : PLTP_MAXPL, PLTP_MAXPL);
        if (n1 <= direct_max) {
            bool fell_back = false;
            ++E.direct_calls;
            SWCU_TRY(encounter_pltp_direct(ctx, l1, *l2, dt, nenc_out, &fell_back));
            if (!fell_back) return SWCU_OK;
            E.nenc = 0, E.result = nullptr, E.nbox_total = 0, E.nemitted = 0, *nenc_out = 0;
            ++E.direct_fallbacks;
        }
    }
    const int ntot = n1 + n2;
    const int next = 2 * ntot;
    FamTimer ft(ctx, FAM_SWEEP);

    const size_t db = sizeof(double), ib = sizeof(int);
    DevBuf *body_arrays[] = {&E.cx, &E.cy, &E.cz, &E.cvx, &E.cvy, &E.cvz, &E.crenc};
    for (DevBuf *d : body_arrays) SWCU_CUDA(ctx, d->ensure(db * ntot));
    DevBuf *sorted_arrays[] = {&E.sx, &E.sy, &E.sz, &E.svx, &E.svy, &E.svz, &E.srenc};
    for (DevBuf *d : sorted_arrays) SWCU_CUDA(ctx, d->ensure(db * next));
    SWCU_CUDA(ctx, E.sbody.ensure(ib * next));
    SWCU_CUDA(ctx, E.keys_in.ensure(db * next));
    SWCU_CUDA(ctx, E.keys_out.ensure(db * next));
    SWCU_CUDA(ctx, E.vals_in.ensure(ib * next));
    SWCU_CUDA(ctx, E.vals_out.ensure(ib * next));
    SWCU_CUDA(ctx, E.ibeg.ensure(ib * ntot));
    SWCU_CUDA(ctx, E.iend.ensure(ib * ntot));
    SWCU_CUDA(ctx, E.nchunk.ensure(ib * (ntot + 1)));
    SWCU_CUDA(ctx, E.choff.ensure(ib * (ntot + 1)));
    SWCU_CUDA(ctx, E.counters.ensure(sizeof(unsigned long long) * (8 + PLTP_MAXPL)));
    if (!E.h_counters) SWCU_CUDA(ctx, cudaHostAlloc((void **)&E.h_counters, (8 + PLTP_MAXPL) * sizeof(unsigned long long), cudaHostAllocDefault));
    unsigned long long *d_count = E.counters.as<unsigned long long>();       // [0] candidates emitted
    unsigned long long *d_nbox = d_count + 1;                                 // [1] sum nbox
    unsigned long long *d_small = d_count + 8;                                // [8] 0: finalize_small_kernel made the list
    // [2] unique count (canonical_order), [3] hits of the list check
    SWCU_CUDA(ctx, cudaMemsetAsync(d_count, 0, 32, ctx->stream));

    ListDev a = to_dev(l1), b;
    if (l2) b = to_dev(*l2); else { b = a; b.n = 0; }
    // bucket sort (4 launches, K9 fused) unless switched off or the population is large enough for the radix sort to win
    const char *bs = getenv("SWCU_SWEEP_BUCKET");
    const bool bucket = allow_bucket && (bs ? atoi(bs) != 0 : true) && next <= BUCKET_MAXKEYS;
    int nbk = 1;
    while (nbk < BUCKET_MAXNB && (long long)nbk * BUCKET_TARGET < next) nbk <<= 1;
    unsigned long long *d_mm = d_count + 5;  // [5] min, [6] max, [7] flags of the bucket sort
    if (bucket) {
        SWCU_CUDA(ctx, cudaMemsetAsync(d_mm, 0xff, sizeof(unsigned long long), ctx->stream));          // min = ~0
        SWCU_CUDA(ctx, cudaMemsetAsync(d_mm + 1, 0, 2 * sizeof(unsigned long long), ctx->stream));    // max = 0, flags = 0
        SWCU_CUDA(ctx, E.bk_hist.ensure(ib * (size_t)(2 * nbk + 1)));
        SWCU_CUDA(ctx, E.bk_offs.ensure(ib * (size_t)(nbk + 1)));
        SWCU_CUDA(ctx, cudaMemsetAsync(E.bk_hist.p, 0, ib * (size_t)nbk, ctx->stream));
    }
    extent_kernel<<<cdiv(ntot, 256), 256, 0, ctx->stream>>>(a, b, ntot, E.cx.as<double>(), E.cy.as<double>(),
                                                           E.cz.as<double>(), E.cvx.as<double>(), E.cvy.as<double>(),
                                                           E.cvz.as<double>(), E.crenc.as<double>(),
                                                           E.keys_in.as<double>(), E.vals_in.as<int>(),
                                                           bucket ? d_mm : nullptr);
    SWCU_KERNEL_CHECK(ctx);

    size_t tmp_sort = 0, tmp_scan = 0;
    SWCU_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, E.keys_in.as<double>(), E.keys_out.as<double>(),
                                                   E.vals_in.as<int>(), E.vals_out.as<int>(), next, 0, 64, ctx->stream));
    SWCU_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, E.nchunk.as<int>(), E.choff.as<int>(), ntot + 1,
                                                 ctx->stream));
    SWCU_CUDA(ctx, E.cub_tmp.ensure(std::max(tmp_sort, tmp_scan)));
    if (bucket) {
        int *hist = E.bk_hist.as<int>(), *cursor = hist + nbk, *offs = E.bk_offs.as<int>();
        bucket_hist_kernel<<<cdiv(next, 256), 256, 0, ctx->stream>>>(E.keys_in.as<double>(), next, d_mm, nbk, hist);
        SWCU_KERNEL_CHECK(ctx);
        bucket_scan_kernel<<<1, 1024, 0, ctx->stream>>>(hist, nbk, offs, cursor, d_mm);
        SWCU_KERNEL_CHECK(ctx);
        bucket_scatter_kernel<<<cdiv(next, 256), 256, 0, ctx->stream>>>(E.keys_in.as<double>(), next, d_mm, nbk, cursor,
                                                                       E.keys_out.as<double>(), E.vals_out.as<int>());
        SWCU_KERNEL_CHECK(ctx);
        bucket_sort_kernel<<<nbk, 128, 0, ctx->stream>>>(
            offs, E.keys_out.as<double>(), E.vals_out.as<int>(), ntot, E.cx.as<double>(), E.cy.as<double>(),
            E.cz.as<double>(), E.cvx.as<double>(), E.cvy.as<double>(), E.cvz.as<double>(), E.crenc.as<double>(),
            E.ibeg.as<int>(), E.iend.as<int>(), E.sx.as<double>(), E.sy.as<double>(), E.sz.as<double>(), E.svx.as<double>(),
            E.svy.as<double>(), E.svz.as<double>(), E.srenc.as<double>(), E.sbody.as<int>());
        SWCU_KERNEL_CHECK(ctx);
    } else {
        SWCU_CUDA(ctx, cub::DeviceRadixSort::SortPairs(E.cub_tmp.p, tmp_sort, E.keys_in.as<double>(),
                                                       E.keys_out.as<double>(), E.vals_in.as<int>(), E.vals_out.as<int>(),
                                                       next, 0, 64, ctx->stream));
        ctx->launches += 4;  // CUB's onesweep: histogram + 3..4 passes (counted as library launches of ours)

        endpoint_kernel<<<cdiv(next, 256), 256, 0, ctx->stream>>>(
            ntot, E.vals_out.as<int>(), E.cx.as<double>(), E.cy.as<double>(), E.cz.as<double>(), E.cvx.as<double>(),
            E.cvy.as<double>(), E.cvz.as<double>(), E.crenc.as<double>(), E.ibeg.as<int>(), E.iend.as<int>(),
            E.sx.as<double>(), E.sy.as<double>(), E.sz.as<double>(), E.svx.as<double>(), E.svy.as<double>(),
            E.svz.as<double>(), E.srenc.as<double>(), E.sbody.as<int>());
        SWCU_KERNEL_CHECK(ctx);
    }

    chunk_count_kernel<<<cdiv(ntot + 1, 256), 256, 0, ctx->stream>>>(ntot, E.ibeg.as<int>(), E.iend.as<int>(),
                                                                    E.nchunk.as<int>(), d_nbox);
    SWCU_KERNEL_CHECK(ctx);
    SWCU_CUDA(ctx, cub::DeviceScan::ExclusiveSum(E.cub_tmp.p, tmp_scan, E.nchunk.as<int>(), E.choff.as<int>(), ntot + 1,
                                                 ctx->stream));
    ctx->launches += 1;

    // chunk -> body map, sized from the previous call (grown after this one if it was too small)
    if (E.owner_cap < (size_t)ntot + 1024) E.owner_cap = (size_t)ntot + 1024;
    SWCU_CUDA(ctx, E.owner.ensure(ib * E.owner_cap));
    const int owner_cap = (int)std::min<size_t>(E.owner_cap, 0x7fffffff);
    chunk_owner_kernel<<<cdiv(ntot, 256), 256, 0, ctx->stream>>>(ntot, E.nchunk.as<int>(), E.choff.as<int>(),
                                                                E.owner.as<int>(), owner_cap, d_count + 9);
    SWCU_KERNEL_CHECK(ctx);

    if (E.cand_cap < (size_t)4 * ntot + 65536) E.cand_cap = (size_t)4 * ntot + 65536;
    const double vsmall = std::sqrt(DBL_MIN);  // globals_module.f90:135
    const int b1 = bits_for((unsigned long long)n1), b2 = bits_for((unsigned long long)(single ? n1 : n2));
    const char *sbe = getenv("SWCU_SWEEP_CTAS");
    const int sweep_blocks = ctx->prop.multiProcessorCount * (sbe ? atoi(sbe) : 8);  // 12 and 16 CTAs per SM measured no faster: the kernel is bound by the L2 -> SM path (40 B per clock and SM)
    unsigned long long h_counts[2] = {0, 0}, h_small[1] = {1};
    int h_total_chunks = 0, h_nuniq = 0;
    for (int attempt = 0; attempt < 3; ++attempt) {
        SWCU_CUDA(ctx, E.cand.ensure(sizeof(unsigned long long) * E.cand_cap));
        SWCU_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), ctx->stream));
        sweep_kernel<<<sweep_blocks, 128, 0, ctx->stream>>>(
            ntot, n1, single ? 1 : 0, E.choff.as<int>(), E.owner.as<int>(), owner_cap, E.ibeg.as<int>(), E.iend.as<int>(), E.cx.as<double>(),
            E.cy.as<double>(), E.cz.as<double>(), E.cvx.as<double>(), E.cvy.as<double>(), E.cvz.as<double>(),
            E.crenc.as<double>(), E.sx.as<double>(), E.sy.as<double>(), E.sz.as<double>(), E.svx.as<double>(),
            E.svy.as<double>(), E.svz.as<double>(), E.srenc.as<double>(), E.sbody.as<int>(), dt, vsmall,
            E.cand.as<unsigned long long>(), (unsigned long long)E.cand_cap, d_count, b2);
        SWCU_KERNEL_CHECK(ctx);
        SWCU_CUDA(ctx, E.uniq.ensure(sizeof(unsigned long long) * CAND_SMALL));
        finalize_small_kernel<<<1, 1024, 0, ctx->stream>>>(E.cand.as<unsigned long long>(), (unsigned long long)E.cand_cap,
                                                           d_count, E.uniq.as<unsigned long long>(),
                                                           reinterpret_cast<int *>(d_count + 2), d_small, b2);
        SWCU_KERNEL_CHECK(ctx);
        // one read-back: [0] candidates, [1] sum nbox, [2] unique count, [7] bucket-sort flags, [8] short-list flag, [9] chunks
        SWCU_CUDA(ctx, cudaMemcpyAsync(E.h_counters, d_count, 10 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                       ctx->stream));
        SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        h_counts[0] = E.h_counters[0], h_counts[1] = E.h_counters[1];
        h_nuniq = (int)(E.h_counters[2] & 0xffffffffull);
        h_small[0] = E.h_counters[8];
        h_total_chunks = (int)E.h_counters[9];
        const unsigned long long h_bflag = bucket ? E.h_counters[7] : 0ull;
        if (h_bflag != 0ull) {  // a clump of equal radii or a non-finite extent: the radix sort decides
            ++E.bucket_fallbacks;
            return encounter_sweep_impl(ctx, l1, l2, dt, nenc_out, false, false);
        }
        if (h_counts[0] <= E.cand_cap) break;
        E.cand_cap = (size_t)(h_counts[0] + h_counts[0] / 4 + 1024);  // overflow: grow and sweep again
        if (attempt == 2) return fail(ctx, SWCU_ERR_STATE, "encounter sweep: candidate buffer overflow persists");
    }
    if ((size_t)h_total_chunks > E.owner_cap) E.owner_cap = (size_t)h_total_chunks + (size_t)h_total_chunks / 4;
    E.nbox_total = (int64_t)h_counts[1];
    E.nemitted = (int64_t)h_counts[0];
    const long long ncand = (long long)h_counts[0];
    if (ncand == 0) return SWCU_OK;
    if (h_small[0] == 0ull) {  // short list: sorted and deduplicated on the device already
        E.nenc = h_nuniq;
        E.result = E.uniq.as<unsigned long long>();
        *nenc_out = h_nuniq;
        return SWCU_OK;
    }
    return canonical_order(ctx, ncand, nenc_out, b1, b2);
}

// encounter_check_all_plplm (:42-109): plpl on the fully interacting block, then plm x plt with index2 shifted
// by nplm; the two lists are disjoint, the canonical order is the lexicographic sort of their union.
int encounter_merge_plplm(swcu_context *ctx, const SweepList &plm, const SweepList &plt, double dt, int64_t *nenc_out,
                          bool triangular)
{
    auto &E = ctx->enc;
    int64_t n_a = 0, n_b = 0;
    auto check = triangular ? encounter_triangular : encounter_sweep;  // ENCOUNTER_CHECK TRIANGULAR / SORTSWEEP
    SWCU_TRY(check(ctx, plm, nullptr, dt, &n_a));
    const int64_t nbox_a = E.nbox_total, nem_a = E.nemitted;
    SWCU_CUDA(ctx, E.out1.ensure(sizeof(unsigned long long) * (size_t)(n_a > 0 ? n_a : 1)));
    if (n_a > 0)
        SWCU_CUDA(ctx, cudaMemcpyAsync(E.out1.p, E.result, sizeof(unsigned long long) * n_a, cudaMemcpyDeviceToDevice,
                                       ctx->stream));
    SWCU_TRY(check(ctx, plm, &plt, dt, &n_b));
    E.nbox_total += nbox_a;
    E.nemitted += nem_a;
    const int64_t n = n_a + n_b;
    *nenc_out = n;
    E.nenc = n;
    E.result = nullptr;
    if (n == 0) return SWCU_OK;
    SWCU_CUDA(ctx, E.merged.ensure(sizeof(unsigned long long) * 2 * (size_t)n));
    unsigned long long *m_in = E.merged.as<unsigned long long>(), *m_out = m_in + n;
    if (n_a > 0)
        SWCU_CUDA(ctx, cudaMemcpyAsync(m_in, E.out1.p, sizeof(unsigned long long) * n_a, cudaMemcpyDeviceToDevice,
                                       ctx->stream));
    if (n_b > 0) {
        SWCU_CUDA(ctx, cudaMemcpyAsync(m_in + n_a, E.uniq.p, sizeof(unsigned long long) * n_b, cudaMemcpyDeviceToDevice,
                                       ctx->stream));
        shift_index2_kernel<<<cdiv(n_b, 256), 256, 0, ctx->stream>>>(m_in + n_a, n_b, (unsigned long long)plm.n);
        SWCU_KERNEL_CHECK(ctx);
    }
    size_t tmp_k = 0;
    SWCU_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tmp_k, m_in, m_out, n, 0, 64, ctx->stream));
    SWCU_CUDA(ctx, E.cub_tmp.ensure(tmp_k));
    SWCU_CUDA(ctx, cub::DeviceRadixSort::SortKeys(E.cub_tmp.p, tmp_k, m_in, m_out, n, 0, 64, ctx->stream));
    ctx->launches += 4;
    E.result = m_out;
    return SWCU_OK;
}

// encounter_check_all_triangular_plpl (l2 == nullptr) / _pltp / _plplm (:436-570); results like encounter_sweep
int encounter_triangular(swcu_context *ctx, const SweepList &l1, const SweepList *l2, double dt, int64_t *nenc_out)
{
    auto &E = ctx->enc;
    E.nenc = 0;
    E.result = nullptr;
    E.nbox_total = 0;
    E.nemitted = 0;
    *nenc_out = 0;
    const int n1 = l1.n, n2 = l2 ? l2->n : 0;
    const bool single = (l2 == nullptr);
    if (n1 == 0 || (!single && n2 == 0)) return SWCU_OK;
    FamTimer ft(ctx, FAM_SWEEP);
    SWCU_CUDA(ctx, E.counters.ensure(64));
    unsigned long long *d_count = E.counters.as<unsigned long long>();
    SWCU_CUDA(ctx, cudaMemsetAsync(d_count, 0, 32, ctx->stream));
    ListDev a = to_dev(l1), b;
    if (l2) b = to_dev(*l2); else { b = a; b.n = 0; }
    const int ncols = single ? n1 : n2;
    const dim3 grid(cdiv(n1, TRI_T), cdiv(ncols, TRI_COLS));
    if (E.cand_cap < 65536) E.cand_cap = 65536;
    const double vsmall = std::sqrt(DBL_MIN);  // globals_module.f90:135
    unsigned long long h_count = 0;
    for (int attempt = 0; attempt < 3; ++attempt) {
        SWCU_CUDA(ctx, E.cand.ensure(sizeof(unsigned long long) * E.cand_cap));
        SWCU_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), ctx->stream));
        tri_check_kernel<<<grid, TRI_T, 0, ctx->stream>>>(a, b, single ? 1 : 0, dt, vsmall, E.cand.as<unsigned long long>(),
                                                         (unsigned long long)E.cand_cap, d_count);
        SWCU_KERNEL_CHECK(ctx);
        SWCU_CUDA(ctx, cudaMemcpyAsync(&h_count, d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (h_count <= E.cand_cap) break;
        E.cand_cap = (size_t)(h_count + h_count / 4 + 1024);  // overflow: grow and run again
        if (attempt == 2) return fail(ctx, SWCU_ERR_STATE, "triangular encounter check: candidate buffer overflow persists");
    }
    E.nemitted = (int64_t)h_count;
    E.nbox_total = (int64_t)n1 * ncols;
    if (h_count == 0) return SWCU_OK;
    return canonical_order(ctx, (long long)h_count, nenc_out);
}

// swiftest_discard_pl_tp on device arrays (tp: positions/velocities/mask; pl: positions/velocities/radius)
int discard_pl_tp(swcu_context *ctx, const Body &tp, const Body &pl, const int32_t *d_lactive, double dt,
                  int32_t *d_iplanet, int32_t *ndiscard)
{
    if (ndiscard) *ndiscard = 0;
    if (tp.n == 0) return SWCU_OK;
    if (pl.n == 0) return fill_i32(ctx, d_iplanet, 0, tp.n);
    SWCU_CUDA(ctx, ctx->scratch64.ensure(128));
    int *d_n = ctx->scratch64.as<int>();
    SWCU_CUDA(ctx, cudaMemsetAsync(d_n, 0, sizeof(int), ctx->stream));
    {
        FamTimer ft(ctx, FAM_PLTP);
        discard_pl_tp_kernel<<<cdiv(tp.n, 128), 128, 0, ctx->stream>>>(
            tp.n, pl.n, tp.rx.as<double>(), tp.ry.as<double>(), tp.rz.as<double>(), tp.vx.as<double>(), tp.vy.as<double>(),
            tp.vz.as<double>(), d_lactive, pl.rx.as<double>(), pl.ry.as<double>(), pl.rz.as<double>(), pl.vx.as<double>(),
            pl.vy.as<double>(), pl.vz.as<double>(), pl.radius.as<double>(), dt, d_iplanet, d_n);
        SWCU_KERNEL_CHECK(ctx);
    }
    if (ndiscard) {
        SWCU_CUDA(ctx, cudaMemcpyAsync(ndiscard, d_n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return SWCU_OK;
}

// symba_encounter_check_list_* pair loop on device arrays; radius2 / l2.renc may be null (test particles)
int symba_check_list(swcu_context *ctx, int64_t nenc, const int32_t *d_i1, const int32_t *d_i2, const int32_t *d_mask,
                     const SweepList &l1, const double *d_radius1, const SweepList &l2, const double *d_radius2, double dt,
                     int32_t *d_lenc, int32_t *d_lvdotr, int64_t *nfound)
{
    if (nfound) *nfound = 0;
    if (nenc <= 0) return SWCU_OK;
    auto &E = ctx->enc;
    SWCU_CUDA(ctx, E.counters.ensure(64));
    unsigned long long *d_count = E.counters.as<unsigned long long>() + 3;
    SWCU_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), ctx->stream));
    {
        FamTimer ft(ctx, FAM_SWEEP);
        symba_check_list_kernel<<<cdiv(nenc, 256), 256, 0, ctx->stream>>>(nenc, d_i1, d_i2, d_mask, to_dev(l1), d_radius1,
                                                                        to_dev(l2), d_radius2, dt, std::sqrt(DBL_MIN),
                                                                        d_lenc, d_lvdotr, d_count);
        SWCU_KERNEL_CHECK(ctx);
    }
    unsigned long long h = 0;
    SWCU_CUDA(ctx, cudaMemcpyAsync(&h, d_count, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (nfound) *nfound = (int64_t)h;
    return SWCU_OK;
}

}  // namespace swcu

// ---- C ABI pieces that only touch encounter state ----
extern "C" int swcu_encounter_fetch(swcu_context *ctx, int64_t nenc, int32_t *index1, int32_t *index2, int32_t *lvdotr)
{
    using namespace swcu;
    if (!ctx) return SWCU_ERR_ARG;
    auto &E = ctx->enc;
    if (E.nenc < 0) return fail(ctx, SWCU_ERR_STATE, "swcu_encounter_fetch: no encounter check result pending");
    if (nenc != E.nenc) return fail(ctx, SWCU_ERR_ARG, "swcu_encounter_fetch: nenc=%lld but the last check found %lld",
                                    (long long)nenc, (long long)E.nenc);
    if (nenc == 0) return SWCU_OK;
    SWCU_CUDA(ctx, E.out1.ensure(sizeof(int32_t) * (size_t)nenc));
    SWCU_CUDA(ctx, E.out2.ensure(sizeof(int32_t) * (size_t)nenc));
    unpack_keys_kernel<<<cdiv(nenc, 256), 256, 0, ctx->stream>>>(E.result, nenc, E.out1.as<int32_t>(), E.out2.as<int32_t>());
    SWCU_KERNEL_CHECK(ctx);
    if (index1) SWCU_CUDA(ctx, cudaMemcpyAsync(index1, E.out1.p, sizeof(int32_t) * nenc, cudaMemcpyDeviceToHost, ctx->stream));
    if (index2) SWCU_CUDA(ctx, cudaMemcpyAsync(index2, E.out2.p, sizeof(int32_t) * nenc, cudaMemcpyDeviceToHost, ctx->stream));
    SWCU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (lvdotr)
        for (int64_t k = 0; k < nenc; ++k) lvdotr[k] = 1;  // lencounter = lvdotr .and. ... (:617-618): always true
    return SWCU_OK;
}

extern "C" int swcu_encounter_direct_count(swcu_context *ctx, int64_t *direct, int64_t *fallbacks)
{
    if (!ctx) return SWCU_ERR_ARG;
    if (direct) *direct = ctx->enc.direct_calls;
    if (fallbacks) *fallbacks = ctx->enc.direct_fallbacks;
    return SWCU_OK;
}

extern "C" int swcu_encounter_bucket_fallbacks(swcu_context *ctx, int64_t *count)
{
    if (!ctx) return SWCU_ERR_ARG;
    if (count) *count = ctx->enc.bucket_fallbacks;
    return SWCU_OK;
}

extern "C" int swcu_encounter_stats(swcu_context *ctx, int64_t *nbox_total, int64_t *ncandidates_emitted)
{
    if (!ctx) return SWCU_ERR_ARG;
    if (nbox_total) *nbox_total = ctx->enc.nbox_total;
    if (ncandidates_emitted) *ncandidates_emitted = ctx->enc.nemitted;
    return SWCU_OK;
}
