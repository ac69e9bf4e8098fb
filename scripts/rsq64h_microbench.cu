// rsq64h_microbench.cu -- accuracy and issue cost of MUFU.RSQ64H (PTX rsqrt.approx.ftz.f64) on the device
// (development aid; results in profiles/r02_rsq64h.md).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/rsq64h scripts/rsq64h_microbench.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ double rsq64h(double r2)
{
    double s;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(s) : "d"(r2));
    return s;
}

// out[i] = r2^-3/2 from the seed with the second-order refinement; e_out[i] = 1 - r2*s^2
__global__ void acc_kernel(const double *in, double *out, double *e_out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double r2 = in[i];
    const double s = rsq64h(r2);
    const double s2 = s * s;
    const double e = fma(-r2, s2, 1.0);
    const double s3 = s2 * s;
    const double q = fma(1.875, e, 1.5);
    const double se = s3 * e;
    out[i] = fma(se, q, s3);
    e_out[i] = e;
}

// NF DFMA chains + NM MUFU.RSQ64H + NI integer ops per iteration
template <int NF, int NM, int NI>
__global__ void mix_kernel(double *out, int iters, double s0)
{
    double a[NF];
    double m[NM > 0 ? NM : 1];
    unsigned u[NI > 0 ? NI : 1];
#pragma unroll
    for (int i = 0; i < NF; ++i) a[i] = s0 + i + threadIdx.x * 1e-3;
#pragma unroll
    for (int i = 0; i < (NM > 0 ? NM : 1); ++i) m[i] = 1.0 + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < (NI > 0 ? NI : 1); ++i) u[i] = threadIdx.x + i;
    const double b = 1.0 + 1e-9 * threadIdx.x, c = 1e-12 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NF; ++i) a[i] = fma(a[i], b, c);
#pragma unroll
        for (int i = 0; i < NM; ++i) m[i] = rsq64h(m[i]) + 0.0 * 0;  // dependent chain per slot; hi word only
#pragma unroll
        for (int i = 0; i < NI; ++i) u[i] = (u[i] ^ (u[i] >> 3)) + 0x9e3779b9u;  // 3 ALU ops
    }
    double s = 0;
    unsigned w = 0;
#pragma unroll
    for (int i = 0; i < NF; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < (NM > 0 ? NM : 1); ++i) s += m[i];
#pragma unroll
    for (int i = 0; i < (NI > 0 ? NI : 1); ++i) w ^= u[i];
    if (s == 123.456 || w == 0x12345u) out[0] = s + w;
}

template <class F>
static double timeit(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int nsm = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, nsm, clk);

    // ---- accuracy ----
    const int n = 1 << 22;
    std::vector<double> h(n), o(n), e(n);
    srand48(12345);
    for (int i = 0; i < n; ++i) {
        const double ex = (i & 1) ? (drand48() * 600.0 - 300.0) : (drand48() * 40.0 - 20.0);  // wide / typical range
        h[i] = (1.0 + drand48()) * pow(2.0, floor(ex));
    }
    // mantissa sweep just around the table breakpoints: r2 = 1 + k*2^-12 ... and the ends of the binade
    for (int k = 0; k < 8192; ++k) h[k] = 1.0 + k * (1.0 / 4096.0) + ((k & 1) ? 0x1p-21 * 0.999 : 0.0);
    double *din, *dout, *de;
    cudaMalloc(&din, n * 8);
    cudaMalloc(&dout, n * 8);
    cudaMalloc(&de, n * 8);
    cudaMemcpy(din, h.data(), n * 8, cudaMemcpyHostToDevice);
    acc_kernel<<<(n + 255) / 256, 256>>>(din, dout, de, n);
    cudaMemcpy(o.data(), dout, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(e.data(), de, n * 8, cudaMemcpyDeviceToHost);
    double emax = 0, emin = 0, rel = 0;
    for (int i = 0; i < n; ++i) {
        emax = fmax(emax, e[i]);
        emin = fmin(emin, e[i]);
        const long double ref = 1.0L / ((long double)h[i] * sqrtl((long double)h[i]));
        rel = fmax(rel, (double)fabsl(((long double)o[i] - ref) / ref));
    }
    printf("RSQ64H: e = 1 - r2*s^2 in [%.3e, %.3e] (2^%.2f); refined r^-3 max rel err %.3e over %d samples\n", emin, emax,
           log2(fmax(emax, -emin)), rel, n);
    // special values
    double sp[8] = {0.0, 4.9e-324, 2.2250738585072014e-308, 1e-310, INFINITY, NAN, 1.7e308, -1.0};
    cudaMemcpy(din, sp, 64, cudaMemcpyHostToDevice);
    acc_kernel<<<1, 8>>>(din, dout, de, 8);
    double so[8];
    cudaMemcpy(so, dout, 64, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 8; ++i) printf("  r2 = %-12.4g -> r^-3 = %g\n", sp[i], so[i]);

    // ---- issue cost ----
    const int iters = 20000;
    double *out;
    cudaMalloc(&out, 1024);
#define MIX(NF, NM, NI, WARPS)                                                                                       \
    {                                                                                                                \
        double ms = timeit([&] { mix_kernel<NF, NM, NI><<<nsm * 4, 32 * WARPS>>>(out, iters, 1.0); });               \
        double cyc = ms * 1e-3 * clk * 1e3 / ((double)iters * WARPS);                                                \
        printf("%2d DFMA + %d RSQ64H + %2d ALU per iter, %d warps/SMSP: %.2f cycles/iter/warp (model 2*NF+NM+NI = %d)\n", NF, NM, 3 * NI, \
               WARPS, cyc, 2 * NF + NM + 3 * NI);                                                                    \
    }
    MIX(20, 0, 0, 3) MIX(20, 1, 0, 3) MIX(20, 2, 0, 3) MIX(20, 4, 0, 3) MIX(20, 1, 1, 3) MIX(20, 1, 2, 3) MIX(10, 1, 0, 3) MIX(10, 2, 0, 3)
    MIX(5, 1, 0, 3) MIX(20, 1, 0, 1) MIX(20, 1, 0, 2) MIX(20, 1, 0, 4)
    return 0;
}
