// int8_quant.cu -- bit-exact GPU versions of attention_cpu/int8_quant.cpp:5-64.
//
// Convention (int8_quant.cpp): scale MULTIPLIES on quantise, DIVIDES on
// dequantise.  Bit-exactness rules: the product x*scale is a separately rounded
// fp32 multiply (__fmul_rn, never contracted into an FMA), rounding is half
// away from zero (roundf == std::round), the division is IEEE (__fdiv_rn).
// Out-of-int32-range products are UB in the reference (float->int32 cast);
// here they saturate to +-127/-128 and NaN maps to -128.
// HBM-bound byte work: 128-bit loads, 32/128-bit stores, grid = k * SM count.
#include "pa_common.cuh"

namespace pa {

__device__ __forceinline__ int quant1(float x, float scale) {
    float r = roundf(__fmul_rn(x, scale));
    r = fminf(127.f, fmaxf(-128.f, r));
    return (int)r;
}

__device__ __forceinline__ uint32_t quant4(float4 x, float scale) {
    return (uint32_t)(quant1(x.x, scale) & 0xff) | ((uint32_t)(quant1(x.y, scale) & 0xff) << 8) |
           ((uint32_t)(quant1(x.z, scale) & 0xff) << 16) | ((uint32_t)(quant1(x.w, scale) & 0xff) << 24);
}

__global__ void quantize_kernel(const float* __restrict__ x, int64_t n, float scale,
                                int8_t* __restrict__ q) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = n / 4;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    uint32_t* q4 = reinterpret_cast<uint32_t*>(q);
    for (int64_t i = tid; i < n4; i += stride) q4[i] = quant4(__ldg(x4 + i), scale);
    for (int64_t i = n4 * 4 + tid; i < n; i += stride) q[i] = (int8_t)quant1(x[i], scale);
}

// rows x dim with one scale per row; dim % 4 == 0 fast path, scalar otherwise.
__global__ void batch_quantize_kernel(const float* __restrict__ x, const float* __restrict__ scales,
                                      int64_t rows, int dim, int8_t* __restrict__ q) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if ((dim & 3) == 0) {
        const int d4 = dim >> 2;
        const int64_t n4 = rows * d4;
        const float4* x4 = reinterpret_cast<const float4*>(x);
        uint32_t* q4 = reinterpret_cast<uint32_t*>(q);
        for (int64_t i = tid; i < n4; i += stride) q4[i] = quant4(__ldg(x4 + i), __ldg(scales + i / d4));
    } else {
        const int64_t n = rows * dim;
        for (int64_t i = tid; i < n; i += stride) q[i] = (int8_t)quant1(x[i], scales[i / dim]);
    }
}

__global__ void dequantize_kernel(const int8_t* __restrict__ q, int64_t n, float scale,
                                  float* __restrict__ x) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = n / 4;
    const uint32_t* q4 = reinterpret_cast<const uint32_t*>(q);
    float4* x4 = reinterpret_cast<float4*>(x);
    for (int64_t i = tid; i < n4; i += stride) {
        uint32_t p = __ldg(q4 + i);
        float4 o;
        o.x = __fdiv_rn((float)(int8_t)(p & 0xff), scale);
        o.y = __fdiv_rn((float)(int8_t)((p >> 8) & 0xff), scale);
        o.z = __fdiv_rn((float)(int8_t)((p >> 16) & 0xff), scale);
        o.w = __fdiv_rn((float)(int8_t)(p >> 24), scale);
        x4[i] = o;
    }
    for (int64_t i = n4 * 4 + tid; i < n; i += stride) x[i] = __fdiv_rn((float)q[i], scale);
}

__global__ void batch_dequantize_kernel(const int8_t* __restrict__ q,
                                        const float* __restrict__ scales, int64_t rows, int dim,
                                        float* __restrict__ x) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if ((dim & 3) == 0) {
        const int d4 = dim >> 2;
        const int64_t n4 = rows * d4;
        const uint32_t* q4 = reinterpret_cast<const uint32_t*>(q);
        float4* x4 = reinterpret_cast<float4*>(x);
        for (int64_t i = tid; i < n4; i += stride) {
            uint32_t p = __ldg(q4 + i);
            float s = __ldg(scales + i / d4);
            float4 o;
            o.x = __fdiv_rn((float)(int8_t)(p & 0xff), s);
            o.y = __fdiv_rn((float)(int8_t)((p >> 8) & 0xff), s);
            o.z = __fdiv_rn((float)(int8_t)((p >> 16) & 0xff), s);
            o.w = __fdiv_rn((float)(int8_t)(p >> 24), s);
            x4[i] = o;
        }
    } else {
        const int64_t n = rows * dim;
        for (int64_t i = tid; i < n; i += stride) x[i] = __fdiv_rn((float)q[i], scales[i / dim]);
    }
}

// |x| max: warp shuffle -> smem -> one atomicMax per block on the float's bit pattern
// (non-negative floats order like unsigned ints).  *out must be zeroed first.
__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, uint32_t* __restrict__ out) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float m = 0.f;
    const int64_t n4 = ((uintptr_t)x % 16 == 0) ? n / 4 : 0;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (int64_t i = tid; i < n4; i += stride) {
        float4 v = __ldg(x4 + i);
        m = fmaxf(fmaxf(m, fabsf(v.x)), fmaxf(fabsf(v.y), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (int64_t i = n4 * 4 + tid; i < n; i += stride) m = fmaxf(m, fabsf(x[i]));
    m = warp_max(m);
    __shared__ float sm[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sm[w] = m;
    __syncthreads();
    if (w == 0) {
        m = (lane < (int)(blockDim.x >> 5)) ? sm[lane] : 0.f;
        m = warp_max(m);
        if (lane == 0) atomicMax(out, __float_as_uint(m));
    }
}

__global__ void scale_from_absmax_kernel(float* v) { *v = __fdiv_rn(127.f, __fadd_rn(*v, 1e-6f)); }

// One warp per row.
__global__ void batch_minmax_scale_kernel(const float* __restrict__ x, int64_t rows, int dim,
                                          float* __restrict__ scales) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = w0; r < rows; r += nw) {
        const float* xr = x + r * dim;
        float m = 0.f;
        for (int d = lane; d < dim; d += 32) m = fmaxf(m, fabsf(__ldg(xr + d)));
        m = warp_max(m);
        if (lane == 0) scales[r] = __fdiv_rn(127.f, __fadd_rn(m, 1e-6f));
    }
}

static int ew_grid(int64_t work_items, int sm_count) {
    int64_t b = (work_items + 255) / 256;
    int64_t cap = (int64_t)sm_count * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace pa

using namespace pa;

PA_API int pa_quantize_i8(const float* d_x, int64_t n, float scale, int8_t* d_q, pa_stream_t stream) {
    PA_CHECK_ARG(n >= 0);
    if (n == 0) return PA_OK;
    PA_CHECK_ARG(d_x && d_q && (uintptr_t)d_x % 16 == 0 && (uintptr_t)d_q % 4 == 0);
    const DeviceInfo& d = device_info();
    if (!d.ok) return PA_ERR_NO_DEVICE;
    quantize_kernel<<<ew_grid((n + 3) / 4, d.sm_count), 256, 0, as_stream(stream)>>>(d_x, n, scale, d_q);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_batch_quantize_i8(const float* d_x, const float* d_scales, int rows, int dim,
                                int8_t* d_q, pa_stream_t stream) {
    PA_CHECK_ARG(rows >= 0 && dim > 0);
    if (rows == 0) return PA_OK;
    PA_CHECK_ARG(d_x && d_scales && d_q && (uintptr_t)d_x % 16 == 0 && (uintptr_t)d_q % 4 == 0);
    const DeviceInfo& d = device_info();
    if (!d.ok) return PA_ERR_NO_DEVICE;
    batch_quantize_kernel<<<ew_grid(((int64_t)rows * dim + 3) / 4, d.sm_count), 256, 0, as_stream(stream)>>>(
        d_x, d_scales, rows, dim, d_q);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_absmax(const float* d_x, int64_t n, float* d_out, pa_stream_t stream) {
    PA_CHECK_ARG(d_out && n >= 0);
    cudaError_t e = cudaMemsetAsync(d_out, 0, sizeof(float), as_stream(stream));
    if (e != cudaSuccess) return (int)e;
    if (n == 0) return PA_OK;
    PA_CHECK_ARG(d_x);
    const DeviceInfo& d = device_info();
    if (!d.ok) return PA_ERR_NO_DEVICE;
    absmax_kernel<<<ew_grid((n + 3) / 4, d.sm_count), 256, 0, as_stream(stream)>>>(
        d_x, n, reinterpret_cast<uint32_t*>(d_out));
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_minmax_scale(const float* d_x, int64_t n, float* d_out, pa_stream_t stream) {
    PA_CHECK_ARG(n > 0);  // the reference dereferences min_element of an empty vector
    int st = pa_absmax(d_x, n, d_out, stream);
    if (st != PA_OK) return st;
    scale_from_absmax_kernel<<<1, 1, 0, as_stream(stream)>>>(d_out);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_batch_minmax_scale(const float* d_x, int rows, int dim, float* d_scales,
                                 pa_stream_t stream) {
    PA_CHECK_ARG(rows >= 0 && dim > 0);
    if (rows == 0) return PA_OK;
    PA_CHECK_ARG(d_x && d_scales);
    const DeviceInfo& d = device_info();
    if (!d.ok) return PA_ERR_NO_DEVICE;
    batch_minmax_scale_kernel<<<ew_grid((int64_t)rows * 32, d.sm_count), 256, 0, as_stream(stream)>>>(
        d_x, rows, dim, d_scales);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_dequantize_i8(const int8_t* d_q, int64_t n, float scale, float* d_x, pa_stream_t stream) {
    PA_CHECK_ARG(n >= 0);
    if (n == 0) return PA_OK;
    PA_CHECK_ARG(d_q && d_x && (uintptr_t)d_x % 16 == 0 && (uintptr_t)d_q % 4 == 0);
    const DeviceInfo& d = device_info();
    if (!d.ok) return PA_ERR_NO_DEVICE;
    dequantize_kernel<<<ew_grid((n + 3) / 4, d.sm_count), 256, 0, as_stream(stream)>>>(d_q, n, scale, d_x);
    PA_RETURN_LAUNCH_STATUS();
}

PA_API int pa_batch_dequantize_i8(const int8_t* d_q, const float* d_scales, int rows, int dim,
                                  float* d_x, pa_stream_t stream) {
    PA_CHECK_ARG(rows >= 0 && dim > 0);
    if (rows == 0) return PA_OK;
    PA_CHECK_ARG(d_q && d_scales && d_x && (uintptr_t)d_x % 16 == 0 && (uintptr_t)d_q % 4 == 0);
    const DeviceInfo& d = device_info();
    if (!d.ok) return PA_ERR_NO_DEVICE;
    batch_dequantize_kernel<<<ew_grid(((int64_t)rows * dim + 3) / 4, d.sm_count), 256, 0, as_stream(stream)>>>(
        d_q, d_scales, rows, dim, d_x);
    PA_RETURN_LAUNCH_STATUS();
}
