// pair_microbench.cu -- issue-rate ceiling of the third-law pair arithmetic with NO memory traffic (development aid;
// results in profiles/r02_kick_flat.md).  Each lane owns 4 row bodies in registers and meets a synthetic column body per
// step; variants remove parts of the chain to see what the FP64 pipe sustains at 1..4 warps per SMSP.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/pairmb.bin scripts/pair_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int CTAS>
__global__ void __launch_bounds__(128, CTAS) k(double *out, int iters, double s0)
{
    double xi[4], yi[4], zi[4], gi[4], ax[4], ay[4], az[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        xi[b] = s0 + threadIdx.x * 1e-3 + b; yi[b] = 2 * s0 + b * 0.5 + threadIdx.x * 1e-4; zi[b] = 0.1 * b + threadIdx.x * 1e-5;
        gi[b] = 1e-9 * (b + 1); ax[b] = ay[b] = az[b] = 0.0;
    }
    double xj = 10.0 + s0, yj = -3.0, zj = 0.5, gj = 1e-9;
    double jx = 0, jy = 0, jz = 0;
    unsigned hm = 0xffffffffu;
    for (int it = 0; it < iters; ++it) {
        double dx[4], dy[4], dz[4], r2[4], s[4], y3[4];
        unsigned hi[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) { dx[b] = xj - xi[b]; dy[b] = yj - yi[b]; dz[b] = zj - zi[b]; }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double t2 = fma(dy[b], dy[b], dx[b] * dx[b]);
            r2[b] = fma(dz[b], dz[b], t2);
            hi[b] = __double2hiint(r2[b]);
            if (MODE == 1) {  // no MUFU: seed from integer ops on the high word (garbage value, same dependency shape)
                s[b] = __hiloint2double(0x5fe00000 - (hi[b] >> 1), __double2loint(t2));
            } else {
                double t;
                asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(t) : "d"(r2[b]));
                s[b] = __hiloint2double(__double2hiint(t), __double2loint(t2));
            }
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double s2 = s[b] * s[b];
            const double e = fma(-r2[b], s2, 1.0);
            const double s3 = s2 * s[b];
            const double q = fma(1.875, e, 1.5);
            const double se = s3 * e;
            y3[b] = fma(se, q, s3);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double fj = gj * y3[b];
            ax[b] = fma(fj, dx[b], ax[b]); ay[b] = fma(fj, dy[b], ay[b]); az[b] = fma(fj, dz[b], az[b]);
            if (MODE != 2) {  // MODE 2: no reaction on j (14 FP64 + ... per pair)
                const double fi = gi[b] * y3[b];
                jx = fma(-fi, dx[b], jx); jy = fma(-fi, dy[b], jy); jz = fma(-fi, dz[b], jz);
            }
        }
        hm = __vimin3_u32(__vimin3_u32(hm, hi[0], hi[1]), hi[2], hi[3]);
        xj += 1e-3; yj -= 1e-3; zj += 1e-4;  // 3 more FP64 per step (0.75 per pair), no memory
    }
    double r = jx + jy + jz + hm;
#pragma unroll
    for (int b = 0; b < 4; ++b) r += ax[b] + ay[b] + az[b];
    if (r == 123.456) out[0] = r;
}

template <class F>
static double timeit(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaEventRecord(e0); f(); cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int nsm = p.multiProcessorCount, iters = 20000;
    double *out; cudaMalloc(&out, 64);
    printf("device %s %d SMs %d kHz; cycles per pair per SMSP (FP64/pair: mode0 20.75, mode1 20.75 no MUFU, mode2 14.75 no reaction)\n", p.name, nsm, clk);
#define RUN(MODE, WARPS)                                                                                      \
    {                                                                                                         \
        /* WARPS warps per SMSP = 4*WARPS warps per SM, as CTAs of 128 threads */                              \
        double ms = timeit([&] { k<MODE, (WARPS > 3 ? WARPS : 3)><<<nsm * WARPS, 128>>>(out, iters, 1.0); });                           \
        double cyc = ms * 1e-3 * clk * 1e3 / ((double)iters * 4 * WARPS);                                      \
        printf("mode %d, %d warps/SMSP: %.2f cycles/pair\n", MODE, WARPS, cyc);                                \
    }
    RUN(0, 1) RUN(0, 2) RUN(0, 3) RUN(0, 4) RUN(0, 5) RUN(1, 3) RUN(1, 4) RUN(2, 3) RUN(2, 4) RUN(2, 5) RUN(2, 6)
    return 0;
}
