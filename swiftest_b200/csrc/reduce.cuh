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

// One CTA of SERIAL_THREADS threads.  The terms of a tile of bodies are evaluated in parallel into shared memory (the loads
// and the arithmetic of 256 bodies at once), then lane k of the first warp adds component k of the tile IN INDEX ORDER
// (or reverse) onto its running sum: the same additions in the same order as the one-lane-walks-global-memory form this
// replaces -- which paid a dependent global load per body, 115 us for 1000 bodies -- at the cost of the DADD chain alone.
constexpr int SERIAL_THREADS = 256;
template <int K, class Term, class Fin>
__global__ void __launch_bounds__(SERIAL_THREADS) sum_serial_kernel(int n, bool reverse, Term term, Fin fin)
{
    __shared__ double sm[K][SERIAL_THREADS];
    __shared__ unsigned char ok[SERIAL_THREADS];
    const int t = threadIdx.x;
    double s = 0.0;
    for (int base = 0; base < n; base += SERIAL_THREADS) {
        const int m = min(SERIAL_THREADS, n - base);  // bodies of this tile, in summation order q = 0 .. m-1
        if (t < m) {
            const int i = reverse ? (n - 1 - (base + t)) : (base + t);
            double tv[K];
            ok[t] = term(i, tv) ? 1 : 0;
#pragma unroll
            for (int k = 0; k < K; ++k) sm[k][t] = tv[k];
        }
        __syncthreads();
        if (t < K) {
            for (int q = 0; q < m; ++q)
                if (ok[q]) s = s + sm[t][q];
        }
        __syncthreads();
    }
    if (t < 32) {
        double tot[K];
#pragma unroll
        for (int k = 0; k < K; ++k) tot[k] = __shfl_sync(0xffffffffu, s, k);
        if (t == 0) fin(tot);
    }
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
