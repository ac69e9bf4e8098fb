// kick_math.cuh -- FP64 inverse square root from an FP32 MUFU.RSQ seed (shared by the gravity kernels).
#pragma once

namespace swcu {

// y = r2^(-1/2) to ~1e-16 relative: seed y0 = rsqrt.approx.f32(float(r2)) (relative error < 2^-22), then one
// third-order Newton step  e = 1 - r2*y0^2,  y = y0*(1 + e/2 + 3e^2/8)  (error ~ 5/16 e^3).
// ok == false when float(r2) is not a normal finite positive number (r2 == 0, denormal, > FLT_MAX): the caller must
// then use the IEEE expression 1/(r2*sqrt(r2)) instead.
// The double<->float conversions are done with integer bit moves (exponent re-bias + funnel shift) instead of F2F:
// F2F.F32.F64 / F2F.F64.F32 issue at 1/4 of the DFMA rate and were measured to hold back the FP64 pipe
// (profiles/r01_kick_notes.md).
__device__ __forceinline__ double rsqrt_newton(double r2, bool &ok)
{
    const int hi = __double2hiint(r2);
    const unsigned lo = (unsigned)__double2loint(r2);
    // float(r2), truncated: exponent re-biased by 1023-127 = 896, top 23 mantissa bits kept
    const unsigned fb = __funnelshift_l(lo, (unsigned)(hi - 0x38000000), 3);
    // usable when the double exponent maps to a normal float exponent (1..254) and r2 is positive and finite
    ok = (unsigned)(hi - 0x38100000) < (unsigned)(0x47F00000 - 0x38100000);
    float y0f;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0f) : "f"(__uint_as_float(fb)));
    const unsigned yb = __float_as_uint(y0f);
    const double y0 = __hiloint2double((int)((yb >> 3) + 0x38000000u), (int)(yb << 29));  // exact widening
    const double t = r2 * y0;
    const double e = fma(-t, y0, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double ye = y0 * e;
    return fma(ye, p, y0);
}

}  // namespace swcu
