// kick_math.cuh -- the per-pair arithmetic shared by the gravity kernels.
//
// Measured on B200 (scripts/fp64_microbench.cu, profiles/r01_fp64_pipe.md): an FP64 instruction holds the SMSP issue
// port for 2 cycles and nothing co-issues in its shadow, every other instruction costs 1 cycle.  The cost of a pair
// evaluation is therefore 2*N_fp64 + N_other issue cycles and BOTH counts are minimised here:
//   * 1/r^3 from an FP32 MUFU.RSQ seed refined directly in six FP64 instructions (rcube_seeded) instead of IEEE
//     sqrt + divide (~60); the potential-energy kernel needs 1/r and uses the five-instruction rsqrt_seeded;
//   * the double<->float conversions are integer moves on the high word (no F2F, which issues at quarter rate);
//   * ONE unsigned compare on the high word of r^2 decides "usable by the fast path": r^2 is a normal float, nonzero,
//     finite AND safely outside the sum of radii.  Pairs that fail contribute exactly zero and are redone by the
//     caller with the reference's IEEE expression and exact radius test (rare: diagonal, overlapping bodies,
//     coordinates outside the FP32 exponent range).
#pragma once
#include <stdint.h>

namespace swcu {

constexpr unsigned SEED_HI_MIN = 0x38100000u;  // high word of the smallest r^2 that maps to a normal float (2^-126)
constexpr unsigned SEED_HI_MAX = 0x47F00000u;  // high word of 2^128: first r^2 that overflows a float

// Threshold pair for the fast-path test of one row body: ok <=> (hi(r2) - thr) <u span.
// rlim2 = (radius_i + max radius of any column)^2; pass rlim2 = 0 for the variants without a radius check.
__device__ __forceinline__ void seed_threshold(double rlim2, unsigned &thr, unsigned &span)
{
    unsigned t = (unsigned)__double2hiint(rlim2) + 1u;  // first high word that guarantees r2 > rlim2
    t = max(t, SEED_HI_MIN);
    t = min(t, SEED_HI_MAX);
    thr = t;
    span = SEED_HI_MAX - t;
}

// y ~ r2^(-1/2) to ~2e-17 relative when ok; when not, y is a denormal (callers treat the pair as rejected and redo it).
// Seed: the top 20 mantissa bits of r2 re-biased into a float, MUFU.RSQ, top 20 bits widened back (relative error
// < 2^-19), then y = y0*(1 + e/2 + 3e^2/8) with e = 1 - r2*y0^2 (error ~ 5/16 e^3 < 2^-55).
// `hy` returns the selected high word of the seed (0 when the pair was rejected): callers that only need to know whether
// ANY pair of a tile was rejected keep the running minimum of hy (one 3-input integer min per two pairs) instead of a
// predicate per pair.
//
// UPPER = false drops the upper end of the range test (r2 < 2^128), saving the subtract: one compare hi >= thr.  It may
// only be used when the caller has established that no r2 can reach 2^128, i.e. every |coordinate| < 2^62
// (COORD_SAFE_MAX; the gravity launchers compute max|coordinate| on the device next to the max radius).
constexpr double COORD_SAFE_MAX = 4611686018427387904.0;  // 2^62: r2 <= 3*(2*2^62)^2 < 2^128

template <bool UPPER = true>
__device__ __forceinline__ double rsqrt_seeded(double r2, unsigned thr, unsigned span, unsigned &hy)
{
    const unsigned hi = (unsigned)__double2hiint(r2);
    const bool ok = UPPER ? ((hi - thr) < span) : (hi >= thr);
    const unsigned fb = (hi << 3) - 0xC0000000u;  // (hi - 0x38000000) << 3 : exponent re-biased by 1023-127
    float y0f;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0f) : "f"(__uint_as_float(fb)));
    const unsigned yb = __float_as_uint(y0f);
    hy = ok ? ((yb >> 3) + 0x38000000u) : 0u;
    // the low word is whatever is at hand (the seed bits): it perturbs a valid seed by < 2^-20 and leaves a rejected
    // one a denormal (< 2^-1042) whose cube underflows to exactly zero -- no register move to build the pair
    const double y0 = __hiloint2double((int)hy, (int)yb);
    const double t = r2 * y0;
    const double e = fma(-t, y0, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double ye = y0 * e;
    return fma(ye, p, y0);
}

// r2^(-3/2) directly: the quantity every gravity kernel needs.  Same seed and the same single compare as rsqrt_seeded,
// then with s2 = s*s and e = 1 - r2*s2:  r2^(-3/2) = s^3 (1 - e)^(-3/2) = s^3 (1 + e (3/2 + 15/8 e)) + O(35/16 e^3),
// six FP64 instructions (s2, e, s3, q, s3*e, fma) instead of five for y plus two for y^3 -- one issue slot pair less per
// pair evaluation.  |e| < 2^-18 so the truncation is < 2^-55; measured against long double on 2e5 random r2:
// max relative error 4.6e-16 (the y -> y^3 route: 6.9e-16).  A rejected pair has a denormal s: s2 underflows to 0,
// s3 = 0 and the result is exactly 0.
template <bool UPPER = true>
__device__ __forceinline__ double rcube_seeded(double r2, unsigned thr, unsigned span, unsigned &hy)
{
    const unsigned hi = (unsigned)__double2hiint(r2);
    const bool ok = UPPER ? ((hi - thr) < span) : (hi >= thr);
    const unsigned fb = (hi << 3) - 0xC0000000u;
    float y0f;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0f) : "f"(__uint_as_float(fb)));
    const unsigned yb = __float_as_uint(y0f);
    hy = ok ? ((yb >> 3) + 0x38000000u) : 0u;
    const double s = __hiloint2double((int)hy, (int)yb);
    const double s2 = s * s;
    const double e = fma(-r2, s2, 1.0);
    const double s3 = s2 * s;
    const double q = fma(1.875, e, 1.5);
    const double se = s3 * e;
    return fma(se, q, s3);
}

// ---- FP64 seed straight from the special-function unit (MUFU.RSQ64H, PTX rsqrt.approx.ftz.f64) ----
// One instruction maps the high word of r2 to the high word of a double s ~ r2^(-1/2) (low word zero): no re-biasing
// into a float and back (3 integer instructions per pair with the FP32 seed above) and no FP32 exponent range to guard.
// Measured on B200 over 4.2e6 samples (scripts/rsq64h_microbench.cu, profiles/r02_rsq64h.md): e = 1 - r2*s^2 within
// +-2^-19.05, so the same second-order refinement s^3 (1 + e (3/2 + 15/8 e)) leaves 35/16 e^3 < 2^-56; max relative
// error of r2^(-3/2) against long double 2.3e-16.  r2 = 0 / denormal give s = inf, r2 < 2^-680 overflows s^3 and
// r2 = inf gives NaN: callers keep r2 inside [2^-600, inf) with the running-minimum test below and a coordinate guard.
constexpr unsigned RSQ64H_HI_MIN = 0x1A700000u;        // high word of 2^-600
constexpr double COORD_SAFE_MAX_F64 = 0x1p500;         // |coordinate| below this: r2 <= 3*(2^501)^2 stays finite

// first high word of r2 that guarantees r2 > rlim2 (and r2 >= 2^-600)
__device__ __forceinline__ unsigned rsq64h_threshold(double rlim2)
{
    unsigned t = (unsigned)__double2hiint(rlim2) + 1u;
    t = max(t, RSQ64H_HI_MIN);
    return min(t, 0x7FE00000u);
}

// r2^(-3/2); `hi` returns the high word of r2 for the caller's running minimum (0xffffffff for a masked pair).  With
// MASKED a pair whose mask is false contributes exactly zero (seed high word 0: s is a denormal, s2 underflows to 0,
// e = 1, s3 = 0, result 0) whatever r2 is.  An unmasked pair below the caller's threshold returns garbage (possibly
// inf/NaN): the caller must discard the results of the whole tile and redo it with the reference's IEEE expression.
// `lo_donor` is any double that is dead after this call (the callers pass the partial sum dx^2+dy^2): MUFU.RSQ64H only
// produces a high word, and taking the low word from a dying register pair lets the seed be written in place -- no
// instruction to build the pair.  The donated low word perturbs the seed by < 2^-20: |e| < 2^-18.3, truncation error
// 35/16 e^3 < 6e-17.
template <bool MASKED>
__device__ __forceinline__ double rsq64h_seed(double r2, double lo_donor, bool m, unsigned &hi)
{
    double t;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(t) : "d"(r2));
    hi = (unsigned)__double2hiint(r2);
    int sh = __double2hiint(t);
    if (MASKED) {
        sh = m ? sh : 0;
        hi = m ? hi : 0xffffffffu;
    }
    return __hiloint2double(sh, __double2loint(lo_donor));
}

template <bool MASKED>
__device__ __forceinline__ double rcube_rsq64h(double r2, double lo_donor, bool m, unsigned &hi)
{
    const double s = rsq64h_seed<MASKED>(r2, lo_donor, m, hi);
    const double s2 = s * s;
    const double e = fma(-r2, s2, 1.0);
    const double s3 = s2 * s;
    const double q = fma(1.875, e, 1.5);
    const double se = s3 * e;
    return fma(se, q, s3);
}

// r2^(-1/2) from the same seed (potential energy): y = s (1 + e/2 + 3/8 e^2), e = 1 - r2 s^2; 5/16 e^3 < 2^-56
__device__ __forceinline__ double rsqrt_rsq64h(double r2, double lo_donor, unsigned &hi)
{
    const double s = rsq64h_seed<false>(r2, lo_donor, true, hi);
    const double t = r2 * s;
    const double e = fma(-t, s, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double se = s * e;
    return fma(se, p, s);
}

// Same test as rsqrt_seeded without the arithmetic (used by the redo paths to find the skipped pairs).
__device__ __forceinline__ bool seed_ok(double r2, unsigned thr, unsigned span)
{
    return ((unsigned)__double2hiint(r2) - thr) < span;
}

}  // namespace swcu
