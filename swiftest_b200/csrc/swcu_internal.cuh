// swcu_internal.cuh -- context, device buffers and launch helpers shared by the translation units of
// libswiftest_cuda.so.  Nothing here is visible through the C ABI (include/swiftest_cuda.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/swiftest_cuda.h"

namespace swcu {

// Growable device allocation.  Capacity only grows; every array is padded so that bulk (TMA) copies rounded
// up to 16 bytes and clamped row loads never leave the allocation.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        bytes += 256;
        if (bytes <= cap) return cudaSuccess;
        size_t want = bytes + bytes / 4;
        want = (want + 255) & ~size_t(255);
        if (p) {
            cudaError_t e = cudaFree(p);
            p = nullptr;
            cap = 0;
            if (e != cudaSuccess) return e;
        }
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return e;
        cap = want;
        return cudaSuccess;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// One resident body population, structure-of-arrays in HBM.
struct Body {
    int n = 0;
    int nplm = 0;
    uint64_t generation = ~uint64_t(0);
    bool valid = false;
    int slice0 = 0, slice1 = 0;  // rows this rank owns (multi-GPU); [0,n) on one GPU
    DevBuf rx, ry, rz, vx, vy, vz, ax, ay, az;
    DevBuf Gm, radius, rhill, renc, mu;
    DevBuf lmask, iflag;  // int32
    // status /= INACTIVE of the reference (swcu_body_set_active); not loaded = every body active.  Distinct from lmask:
    // swiftest_util_coord_vb2vh_pl filters on status, helio_drift_linear_pl and the kicks on lmask
    DevBuf lactive;
    bool has_active = false;
    // democratic-heliocentric integrators keep two velocities: v* above is vh, w* is vb (allocated on first use);
    // b* / e* are the planet positions at the begin / end kick (pl%rbeg, pl%rend: swiftest_util.f90:2057-2080)
    DevBuf wx, wy, wz, bx, by, bz, ex, ey, ez;
    bool helio_ready = false;
    void release()
    {
        DevBuf *all[] = {&rx, &ry, &rz, &vx, &vy, &vz, &ax, &ay, &az, &Gm, &radius, &rhill, &renc, &mu, &lmask, &iflag,
                         &wx, &wy, &wz, &bx, &by, &bz, &ex, &ey, &ez, &lactive};
        helio_ready = false;
        has_active = false;
        for (DevBuf *b : all) b->release();
        valid = false;
        n = 0;
    }
};

enum KernelFamily { FAM_PLPL = 0, FAM_PLTP = 1, FAM_DRIFT = 2, FAM_SWEEP = 3, FAM_ALLGATHER = 4, FAM_COUNT = 5 };

struct NcclApi;  // comm.cu

}  // namespace swcu

struct swcu_context {
    int device = 0;
    cudaDeviceProp prop;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;

    swcu::Body pl, tp;

    // scratch populations for the tier-1 (host pointer) entry points, kept apart from the resident ones
    swcu::Body s_pl, s_tp;

    // staging (AoS images of host arrays) and generic scratch
    swcu::DevBuf stage[6];
    swcu::DevBuf istage[2];
    swcu::DevBuf partial;  // j-split partial accelerations
    swcu::DevBuf scratch64;

    // encounter state (persists across calls like the reference's `save`d bounding box, encounter_check.f90:163)
    struct Enc {
        swcu::DevBuf keys_in, keys_out, vals_in, vals_out, cub_tmp;
        swcu::DevBuf cx, cy, cz, cvx, cvy, cvz, crenc;          // concatenated sweep population (double list)
        swcu::DevBuf sx, sy, sz, svx, svy, svz, srenc, sbody;   // gathered into sorted-endpoint order
        swcu::DevBuf ibeg, iend, nchunk, choff, owner;
        size_t owner_cap = 0;  // entries of the chunk -> body map
        swcu::DevBuf cand, cand_sorted, uniq, counters;
        swcu::DevBuf out1, out2;
        size_t cand_cap = 0;   // pairs
        int64_t nenc = -1;     // result of the last check (-1: none pending)
        int64_t nbox_total = 0, nemitted = 0;
        // second result buffer for the plplm merge
        swcu::DevBuf merged;
        // pl-tp without the sort: planet records, per-planet box counts, how often the sort path had to decide
        swcu::DevBuf abase, boxcnt;
        // bucket sort of the extents: histogram + cursors, offsets; how often the radix sort had to take over
        swcu::DevBuf bk_hist, bk_offs;
        int64_t bucket_fallbacks = 0;
        unsigned long long *h_counters = nullptr;  // pinned: the sweep's counters come back in one copy
        int64_t direct_calls = 0, direct_fallbacks = 0;
        bool direct_attr_set = false;
        const unsigned long long *result = nullptr;  // device pointer to the final sorted unique keys
    } enc;

    // timing
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int kernel_timing = 0;  // 0 off, 1 last launch group per family, 2 accumulate every launch group
    cudaEvent_t fam_ev0[swcu::FAM_COUNT] = {}, fam_ev1[swcu::FAM_COUNT] = {};
    bool fam_pending[swcu::FAM_COUNT] = {};
    double fam_ms[swcu::FAM_COUNT] = {};
    // accumulating mode (kernel_timing == 2): one event pair per launch group, read without stalling the stream
    std::vector<cudaEvent_t> fam_log0[swcu::FAM_COUNT], fam_log1[swcu::FAM_COUNT];
    size_t fam_log_used[swcu::FAM_COUNT] = {};

    swcu::DevBuf flush;
    // asynchronous slice I/O (swcu_body_put_range_async / _get_range_async): dedicated copy streams, double-buffered
    // AoS staging, events that order copy stream <-> compute stream
    struct AsyncIO {
        cudaStream_t h2d = nullptr, d2h = nullptr;
        swcu::DevBuf in[2], out[2];
        cudaEvent_t h2d_done[2] = {}, unpacked[2] = {}, packed[2] = {}, d2h_done[2] = {};
        unsigned long long put_seq = 0, get_seq = 0;
        bool ready = false;
    } aio;
    std::vector<cudaEvent_t> lap0, lap1;  // swcu_timer_lap_begin/_end
    size_t lap_used = 0;

    // central-body scalars of the integrator glue, on the device (step_kernels.cu): doubles
    // [0..3] raw sums of the last reduction, [4..6] vbcb, [8..10] ptbeg, [12..14] ptend, [16] GMtot, [20..27] energy sums
    swcu::DevBuf cbs;
    swcu::DevBuf sumbuf;  // per-CTA partials of the tree reductions + ticket counter (zeroed when allocated)
    // Wisdom-Holman planet step (whm_kernels.cu): Jacobi coordinates, running masses, per-body 1/|xj|^3
    struct Whm {
        swcu::DevBuf xjx, xjy, xjz, vjx, vjy, vjz, eta, muj, ir3j;
        uint64_t generation = ~uint64_t(0);
        int n = -1;
        double gmcb = 0.0;
        bool ah0tp_valid = false;  // cbs[CBS_AH0TP] holds whm_kick_getacch_ah0 of the planets for the tp step
    } whm;
    swcu::DevBuf lists[16];  // staging of the encounter-list kernels (list_kernels.cu)
    swcu::DevBuf flat_blockrad; // max radius per block of 128 bodies (third-law kernel)
    swcu::DevBuf tp_discard;  // iplanet of the last swcu_tp_discard_pl
    swcu::DevBuf flat_guard; // 2 x u64: max |coordinate| bit pattern of the current / next launch (flat_prologue_kernel)
    int flat_parity = 0;
    // whole multi-launch steps replayed as a CUDA graph (step_kernels.cu): captured on the second step with the same key
    struct StepGraph {
        cudaGraphExec_t exec = nullptr;
        int n = -1, nplm = -1, variant = 0, lclose = 0, tune_variant = 0, tune_nsplit = 0, tune_ib = 0;
        bool has_active = false;
        double gmcb = 0.0, dt = 0.0;
        uint64_t generation = 0;
        const void *p0 = nullptr;
        cudaStream_t stream = nullptr;
        long long launches = 0;
        int warm = 0;
        bool disabled = false;
        int64_t replays = 0;
    } helio_graph;
    swcu::DevBuf flat_redo;  // u64 count of chunks the third-law kernel redid exactly (zeroed when allocated)
    swcu::DevBuf flat_trace; // per-warp timeline of the third-law kernel (development aid, SWCU_FLAT_TRACE)

    // multi-GPU
    swcu::NcclApi *nccl = nullptr;
    void *comm = nullptr;
    int nranks = 1, rank = 0;
    swcu::DevBuf sendbuf, recvbuf;

    // peer-memory exchange (fused reduce + update + allgather over NVLink, comm_p2p / drift_kernels.cu)
    struct P2P {
        bool ready = false;
        int nranks = 1, rank = 0;
        size_t stride = 0;              // doubles per component in F
        swcu::DevBuf F, flags;          // local: partial accelerations [3][stride]; flags [2][16] u64 + error word
        void *peer[8][SWCU_P2P_NBUF];   // mapped pointers of every rank's exported buffers (own rank: local pointers)
        unsigned long long epoch = 0;
        unsigned long long *h_err = nullptr;  // pinned host copy of the error word flags[32], refreshed after every step
    } p2p;

    // tuning overrides (environment: SWCU_KICK_IB, SWCU_KICK_NSPLIT, SWCU_KICK_VARIANT)
    int tune_ib = 0, tune_nsplit = 0, tune_variant = -1;
};

namespace swcu {

int fail(swcu_context *ctx, int code, const char *fmt, ...);

#define SWCU_CUDA(ctx, call)                                                                       \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return swcu::fail((ctx), SWCU_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, \
                              cudaGetErrorString(e__));                                            \
    } while (0)

#define SWCU_TRY(expr)                 \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != SWCU_OK) return rc__; \
    } while (0)

#define SWCU_KERNEL_CHECK(ctx)                 \
    do {                                       \
        (ctx)->launches++;                     \
        SWCU_CUDA((ctx), cudaGetLastError());  \
    } while (0)

struct FamTimer {  // brackets a launch group with events when kernel timing is enabled
    swcu_context *c;
    int fam;
    size_t slot = 0;
    FamTimer(swcu_context *ctx, int family) : c(ctx), fam(family)
    {
        if (c->kernel_timing == 1) {
            cudaEventRecord(c->fam_ev0[fam], c->stream);
        } else if (c->kernel_timing == 2) {
            slot = c->fam_log_used[fam]++;
            if (slot >= c->fam_log0[fam].size()) {
                cudaEvent_t a, b;
                cudaEventCreate(&a);
                cudaEventCreate(&b);
                c->fam_log0[fam].push_back(a);
                c->fam_log1[fam].push_back(b);
            }
            cudaEventRecord(c->fam_log0[fam][slot], c->stream);
        }
    }
    ~FamTimer()
    {
        if (c->kernel_timing == 1) {
            cudaEventRecord(c->fam_ev1[fam], c->stream);
            c->fam_pending[fam] = true;
        } else if (c->kernel_timing == 2) {
            cudaEventRecord(c->fam_log1[fam][slot], c->stream);
        }
    }
};

inline int cdiv(int64_t a, int64_t b) { return int((a + b - 1) / b); }

// ---- layout conversion (host Fortran AoS r(3,n)  <->  device SoA) : layout_kernels.cu ----
int aos_to_soa3(swcu_context *ctx, const double *d_aos, double *x, double *y, double *z, int n);
int soa_to_aos3(swcu_context *ctx, const double *x, const double *y, const double *z, double *d_aos, int n);
int upload_vec3(swcu_context *ctx, const double *h_aos, int n, int stage_slot, DevBuf &x, DevBuf &y, DevBuf &z);
int download_vec3(swcu_context *ctx, double *h_aos, int n, int stage_slot, const DevBuf &x, const DevBuf &y,
                  const DevBuf &z);
int upload_arr(swcu_context *ctx, const void *h, size_t bytes, DevBuf &d);
int fill_f64(swcu_context *ctx, double *d, double value, int n);
int fill3_f64(swcu_context *ctx, double *a, double *b, double *c, double value, int n);
int fill_i32(swcu_context *ctx, int32_t *d, int32_t value, int n);
int ensure_body(swcu_context *ctx, Body &b, int n);

// ---- gravity : kick_kernels.cu ----
struct KickProblem {
    // rows (bodies that receive acceleration)
    const double *xi, *yi, *zi, *radi;
    int row0, row1;  // half-open
    // columns (bodies that exert it)
    const double *xj, *yj, *zj, *gmj, *radj;
    int col0, col1;
    bool diag;             // rows and columns index the same population: skip i == j
    const int32_t *lmask;  // optional row mask
    double *ax, *ay, *az;  // accumulated in place (+=)
};
int kick_rows(swcu_context *ctx, const KickProblem &p, int family);
int kick_pl_tri(swcu_context *ctx, Body &pl, bool lrad, int row0, int row1);
int kick_pl_flat(swcu_context *ctx, Body &pl, bool lrad, int nplm_rows, bool reduce = true);
int max_radius(swcu_context *ctx, const double *radius, const double *x, const double *y, const double *z, int n,
               int slot, const double **d_out);
int kick_pair_list(swcu_context *ctx, const Body &pl, bool lrad, int64_t nenc, const int32_t *d_i1, const int32_t *d_i2,
                   double *ex, double *ey, double *ez);
int axpy3(swcu_context *ctx, double alpha, const double *x0, const double *x1, const double *x2, double *y0,
          double *y1, double *y2, const int32_t *lmask, int n);

// ---- drift : drift_kernels.cu ----
// vsel = 0 drifts (r, v) with the per-body mu array; vsel = 1 drifts (r, w) = (rh, vb) with mu = mu_scalar = GMcb, the
// democratic-heliocentric pair (helio_drift.f90:38-40)
int drift_bodies(swcu_context *ctx, Body &b, int i0, int i1, double dt, int lgr, double inv_c2, int32_t *nfail,
                 int vsel = 0, double mu_scalar = 0.0);
int helio_tp_step(swcu_context *ctx, Body &tp, const Body &pl, double gmcb, double dt, int lfirst, int32_t *nfail);
// Danby drift of arbitrary SoA arrays with a per-body mu (WHM: Jacobi coordinates with muj); failures counted in scratch64[0]
int drift_arrays(swcu_context *ctx, int n, const double *mu, double *x, double *y, double *z, double *vx, double *vy,
                 double *vz, const int32_t *lmask, int32_t *iflag, double dt);

// ---- integrator glue : step_kernels.cu ----
constexpr int CBS_SUM = 0, CBS_VBCB = 4, CBS_PTBEG = 8, CBS_PTEND = 12, CBS_GMTOT = 16, CBS_ENERGY = 20, CBS_AH0PL = 28,
              CBS_AH0TP = 32, CBS_DOUBLES = 40;
int ensure_step_state(swcu_context *ctx);
int ensure_helio(swcu_context *ctx, Body &b);
int pl_vh2vb(swcu_context *ctx, double gmcb);
int pl_vb2vh(swcu_context *ctx, double gmcb);
int pl_lindrift(swcu_context *ctx, double gmcb, double dt, int lbeg);
int tp_lindrift(swcu_context *ctx, double dt, int lbeg);
int tp_vh2vb(swcu_context *ctx, int lbeg);  // vbcb = -ptbeg / -ptend (helio_step.f90:103,118)
int tp_vb2vh(swcu_context *ctx, int lbeg);
int kick_vb_save(swcu_context *ctx, Body &b, double dt, int save /*0 none, 1 rbeg, 2 rend*/);
int helio_step_pl(swcu_context *ctx, double gmcb, double dt, int variant, int lclose, int lfirst, int32_t *nfail);
// ---- Wisdom-Holman planet step : whm_kernels.cu ----
int whm_step_pl(swcu_context *ctx, double gmcb, double dt, int variant, int lclose, int lfirst, int32_t *nfail);
int whm_tp_first_accel(swcu_context *ctx);
// the whole planet step in one launch for small systems (drift_kernels.cu: it shares the drift device functions)
int whm_small_max();
int whm_step_pl_small(swcu_context *ctx, Body &pl, double gmcb, double dt, int flat, int lclose, int lfirst);
int helio_step_pl_small(swcu_context *ctx, Body &pl, double gmcb, double dt, int flat, int lclose, int lfirst);
int whm_get_jacobi(swcu_context *ctx, double *xj, double *vj);
int pl_accel_int(swcu_context *ctx, int loop_variant, int lclose);  // swcu_api.cu

// ---- energy and momentum : energy_kernels.cu ----
// positions / velocities in the s_pl scratch population (rb in r*, vb in v*), mass in s_pl.mu, Gmass in s_pl.Gm
int energy_and_momentum(swcu_context *ctx, Body &b, double gmcb, int lclose, bool pe_only, double *out8);
// ah0 == nullptr: use whm_kick_getacch_ah0 left in cbs[CBS_AH0TP] by whm_step_pl of the same step
int whm_tp_step(swcu_context *ctx, Body &tp, const Body &pl, double dt, const double *ah0, int32_t *nfail);

// ---- encounters : encounter_kernels.cu ----
struct SweepList {
    const double *x, *y, *z, *vx, *vy, *vz, *renc;  // renc may be null (test particles: 0)
    int n;
};
int encounter_sweep(swcu_context *ctx, const SweepList &l1, const SweepList *l2, double dt, int64_t *nenc);
int encounter_merge_plplm(swcu_context *ctx, const SweepList &plm, const SweepList &plt, double dt, int64_t *nenc,
                          bool triangular = false);
int set_renc(swcu_context *ctx, Body &pl, int irec);
int encounter_triangular(swcu_context *ctx, const SweepList &l1, const SweepList *l2, double dt, int64_t *nenc);
int discard_pl_tp(swcu_context *ctx, const Body &tp, const Body &pl, const int32_t *d_lactive, double dt,
                  int32_t *d_iplanet, int32_t *ndiscard);
int symba_check_list(swcu_context *ctx, int64_t nenc, const int32_t *d_i1, const int32_t *d_i2, const int32_t *d_mask,
                     const SweepList &l1, const double *d_radius1, const SweepList &l2, const double *d_radius2, double dt,
                     int32_t *d_lenc, int32_t *d_lvdotr, int64_t *nfound);

// ---- comm : comm.cu ----
int comm_allgather_pl(swcu_context *ctx, int with_v);
int comm_allreduce_sum(swcu_context *ctx, double *buf, size_t count);
int p2p_step_after_kick(swcu_context *ctx, double dt, int32_t *nfail);  // drift_kernels.cu
int p2p_check_error(swcu_context *ctx);  // comm.cu: SWCU_ERR_STATE if a peer-memory exchange timed out (clears the word)
void comm_release(swcu_context *ctx);

}  // namespace swcu
