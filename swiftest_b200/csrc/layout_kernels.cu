// layout_kernels.cu -- conversion between the host's Fortran layout r(NDIM,n) (x1,y1,z1,x2,...) and the
// device-resident structure-of-arrays, plus small fill helpers.  Pure data movement (HBM bound).
#include "swcu_internal.cuh"

#include <stdarg.h>

namespace swcu {

int fail(swcu_context *ctx, int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

namespace {

// A warp moves 32 consecutive bodies: 96 contiguous doubles are read (or written) with unit stride through
// shared memory so both the AoS and the SoA side are fully coalesced.
__global__ void __launch_bounds__(256) aos_to_soa3_kernel(const double *__restrict__ aos, double *__restrict__ x,
                                                          double *__restrict__ y, double *__restrict__ z, int n)
{
    __shared__ double tile[3 * 256];
    const int base = blockIdx.x * 256;
    const int cnt = min(256, n - base);
    for (int q = threadIdx.x; q < 3 * cnt; q += 256) tile[q] = aos[(size_t)3 * base + q];
    __syncthreads();
    const int t = threadIdx.x;
    if (t < cnt) {
        x[base + t] = tile[3 * t];
        y[base + t] = tile[3 * t + 1];
        z[base + t] = tile[3 * t + 2];
    }
}

__global__ void __launch_bounds__(256) soa_to_aos3_kernel(const double *__restrict__ x, const double *__restrict__ y,
                                                          const double *__restrict__ z, double *__restrict__ aos, int n)
{
    __shared__ double tile[3 * 256];
    const int base = blockIdx.x * 256;
    const int cnt = min(256, n - base);
    const int t = threadIdx.x;
    if (t < cnt) {
        tile[3 * t] = x[base + t];
        tile[3 * t + 1] = y[base + t];
        tile[3 * t + 2] = z[base + t];
    }
    __syncthreads();
    for (int q = threadIdx.x; q < 3 * cnt; q += 256) aos[(size_t)3 * base + q] = tile[q];
}

__global__ void fill_f64_kernel(double *d, double v, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = v;
}
__global__ void fill3_f64_kernel(double *a, double *b, double *c, double v, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v, b[i] = v, c[i] = v;
}
__global__ void fill_i32_kernel(int32_t *d, int32_t v, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = v;
}

}  // namespace

int aos_to_soa3(swcu_context *ctx, const double *d_aos, double *x, double *y, double *z, int n)
{
    if (n <= 0) return SWCU_OK;
    aos_to_soa3_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_aos, x, y, z, n);
    SWCU_KERNEL_CHECK(ctx);
    return SWCU_OK;
}

int soa_to_aos3(swcu_context *ctx, const double *x, const double *y, const double *z, double *d_aos, int n)
{
    if (n <= 0) return SWCU_OK;
    soa_to_aos3_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(x, y, z, d_aos, n);
    SWCU_KERNEL_CHECK(ctx);
    return SWCU_OK;
}

int upload_vec3(swcu_context *ctx, const double *h_aos, int n, int slot, DevBuf &x, DevBuf &y, DevBuf &z)
{
    if (n <= 0) return SWCU_OK;
    const size_t bytes = sizeof(double) * 3 * (size_t)n;
    SWCU_CUDA(ctx, ctx->stage[slot].ensure(bytes));
    SWCU_CUDA(ctx, x.ensure(sizeof(double) * (size_t)n));
    SWCU_CUDA(ctx, y.ensure(sizeof(double) * (size_t)n));
    SWCU_CUDA(ctx, z.ensure(sizeof(double) * (size_t)n));
    SWCU_CUDA(ctx, cudaMemcpyAsync(ctx->stage[slot].p, h_aos, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return aos_to_soa3(ctx, ctx->stage[slot].as<double>(), x.as<double>(), y.as<double>(), z.as<double>(), n);
}

int download_vec3(swcu_context *ctx, double *h_aos, int n, int slot, const DevBuf &x, const DevBuf &y, const DevBuf &z)
{
    if (n <= 0) return SWCU_OK;
    const size_t bytes = sizeof(double) * 3 * (size_t)n;
    SWCU_CUDA(ctx, ctx->stage[slot].ensure(bytes));
    SWCU_TRY(soa_to_aos3(ctx, x.as<double>(), y.as<double>(), z.as<double>(), ctx->stage[slot].as<double>(), n));
    SWCU_CUDA(ctx, cudaMemcpyAsync(h_aos, ctx->stage[slot].p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return SWCU_OK;
}

int upload_arr(swcu_context *ctx, const void *h, size_t bytes, DevBuf &d)
{
    SWCU_CUDA(ctx, d.ensure(bytes));
    if (bytes) SWCU_CUDA(ctx, cudaMemcpyAsync(d.p, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return SWCU_OK;
}

int fill_f64(swcu_context *ctx, double *d, double value, int n)
{
    if (n <= 0) return SWCU_OK;
    fill_f64_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d, value, n);
    SWCU_KERNEL_CHECK(ctx);
    return SWCU_OK;
}

// three arrays of one vector quantity (ah = 0 before every force evaluation) in one launch
int fill3_f64(swcu_context *ctx, double *a, double *b, double *c, double value, int n)
{
    if (n <= 0) return SWCU_OK;
    fill3_f64_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(a, b, c, value, n);
    SWCU_KERNEL_CHECK(ctx);
    return SWCU_OK;
}

int fill_i32(swcu_context *ctx, int32_t *d, int32_t value, int n)
{
    if (n <= 0) return SWCU_OK;
    fill_i32_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d, value, n);
    SWCU_KERNEL_CHECK(ctx);
    return SWCU_OK;
}

// make every array of a population large enough for n bodies (contents undefined for new allocations)
int ensure_body(swcu_context *ctx, Body &b, int n)
{
    const size_t nb = sizeof(double) * (size_t)(n > 0 ? n : 1);
    DevBuf *f64[] = {&b.rx, &b.ry, &b.rz, &b.vx, &b.vy, &b.vz, &b.ax, &b.ay, &b.az, &b.Gm, &b.radius, &b.rhill, &b.renc, &b.mu};
    for (DevBuf *d : f64) SWCU_CUDA(ctx, d->ensure(nb));
    SWCU_CUDA(ctx, b.lmask.ensure(sizeof(int32_t) * (size_t)(n > 0 ? n : 1)));
    SWCU_CUDA(ctx, b.iflag.ensure(sizeof(int32_t) * (size_t)(n > 0 ? n : 1)));
    return SWCU_OK;
}

}  // namespace swcu
