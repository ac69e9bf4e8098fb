// reduce.cuh -- deterministic K-component sums over a body population, used by the O(N) glue of the integrators
// (step_kernels.cu) and by the energy / momentum sums (energy_kernels.cu).
//
// Two shapes, chosen by n alone so that a given n always sums in the same order on every device:
//   * n <= SERIAL_SUM_MAX: one lane per component adds the terms one after the other in index order (or reverse) --
//     the order of the reference's serial loops and `sum` intrinsics, so the result is bit-identical to the CPU
//     restatement for the planetary systems the reference ships (8 and 108 bodies);
//   * larger n: fixed grid of min(ceil(n/256), 512) CTAs, grid-stride per thread, shuffle tree, one partial per CTA,
//     the last CTA to finish (ticket counter) folds the partials with the same tree.  Same bits run to run.
// The functor yields the K terms of body i (it applies its own mask by returning zeros) and `fin` consumes the K totals
// in one thread, so a reduction and its dependent scalar arithmetic (e.g. pt = S / GMcb) are one launch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace swcu {

constexpr int SERIAL_SUM_MAX = 1024;
constexpr int SUM_THREADS = 256;
constexpr int SUM_MAX_CTAS = 512;

template <int K> __device__ __forceinline__ double pick(const double (&t)[K], int lane)
{
    double v = t[0];
#pragma unroll
    for (int k = 1; k < K; ++k) v = (lane == k) ? t[k] : v;  // register select, no local-memory indexing
    return v;
}

template <int K, class Term, class Fin>
__global__ void __launch_bounds__(32) sum_serial_kernel(int n, bool reverse, Term term, Fin fin)
{
    const int lane = threadIdx.x;
    double s = 0.0;
    if (lane < K) {
        if (!reverse) {
            for (int i = 0; i < n; ++i) {
                double t[K];
                if (term(i, t)) s = s + pick<K>(t, lane);
            }
        } else {
            for (int i = n - 1; i >= 0; --i) {
                double t[K];
                if (term(i, t)) s = s + pick<K>(t, lane);
            }
        }
    }
    double tot[K];
#pragma unroll
    for (int k = 0; k < K; ++k) tot[k] = __shfl_sync(0xffffffffu, s, k);
    if (lane == 0) fin(tot);
}

template <int K> __device__ __forceinline__ void block_tree(double (&s)[K], double (*sm)[K])
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int k = 0; k < K; ++k) s[k] += __shfl_down_sync(0xffffffffu, s[k], o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0)
#pragma unroll
        for (int k = 0; k < K; ++k) sm[w][k] = s[k];
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double t = sm[0][k];
            for (int ww = 1; ww < SUM_THREADS / 32; ++ww) t += sm[ww][k];
            s[k] = t;
        }
    }
}

// partials: gridDim.x * K doubles; ticket: one zero-initialised unsigned that the kernel leaves at zero again
template <int K, class Term, class Fin>
__global__ void __launch_bounds__(SUM_THREADS) sum_tree_kernel(int n, Term term, Fin fin, double *partials, unsigned *ticket)
{
    __shared__ double sm[SUM_THREADS / 32][K];
    __shared__ bool last;
    double s[K];
#pragma unroll
    for (int k = 0; k < K; ++k) s[k] = 0.0;
    for (int i = blockIdx.x * SUM_THREADS + threadIdx.x; i < n; i += gridDim.x * SUM_THREADS) {
        double t[K];
        if (term(i, t)) {
#pragma unroll
            for (int k = 0; k < K; ++k) s[k] += t[k];
        }
    }
    block_tree<K>(s, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) partials[(size_t)blockIdx.x * K + k] = s[k];
        __threadfence();
        last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
#pragma unroll
    for (int k = 0; k < K; ++k) s[k] = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += SUM_THREADS)
#pragma unroll
        for (int k = 0; k < K; ++k) s[k] += __ldcg(&partials[(size_t)b * K + k]);
    block_tree<K>(s, sm);
    if (threadIdx.x == 0) {
        *ticket = 0u;
        fin(s);
    }
}

inline int sum_grid(int n)
{
    const int g = (n + SUM_THREADS - 1) / SUM_THREADS;
    return g < SUM_MAX_CTAS ? (g < 1 ? 1 : g) : SUM_MAX_CTAS;
}

}  // namespace swcu
