// aud_microbench.cu -- measured FP32 (non-tensor) peaks of the device the library runs on: the denominators
// of the fused kernel's FP32 roofline (SURVEY 8d asks for an FMA-loop figure measured on the box instead of
// SMs x lanes x 2 x clock).  Register-resident dependency chains, no memory traffic, timed with CUDA events.
//   kind 0: scalar FFMA                       (2 flop per lane-instruction)
//   kind 1: packed FFMA2  (fma.rn.f32x2)      (4 flop per lane-instruction; sm_100 only)
//   kind 2: scalar FADD : FMUL : FFMA = 2:1:1 (the instruction mix of an FFT butterfly; 5 flop per 4 instructions)
//   kind 3: the same mix as FADD2 / FMUL2 / FFMA2
#include <cuda_runtime.h>

#include "auditory_b200.h"
#include "aud_internal.h"

namespace aud {

constexpr int kChains = 8;

template <int KIND>
__global__ void __launch_bounds__(512) fp32_loop_kernel(float *out, int iters, float seed) {
    float2 a[kChains], b[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
        a[c] = make_float2(seed + (float)(threadIdx.x + c), seed - (float)c);
        b[c] = make_float2(1.0f + seed * (float)c, 0.5f + seed);
    }
    const float2 m = make_float2(0.999f + seed, 1.001f - seed);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) {
            if (KIND == 0) {
                a[c].x = fmaf(a[c].x, m.x, b[c].x); a[c].y = fmaf(a[c].y, m.y, b[c].y);
                b[c].x = fmaf(b[c].x, m.y, a[c].x); b[c].y = fmaf(b[c].y, m.x, a[c].y);
            } else if (KIND == 1) {
                a[c] = __ffma2_rn(a[c], m, b[c]);
                b[c] = __ffma2_rn(b[c], m, a[c]);
            } else if (KIND == 2) {
                const float s0 = a[c].x + b[c].x, s1 = a[c].y - b[c].y;
                const float p0 = s0 * m.x;
                a[c].x = fmaf(s1, m.y, p0); a[c].y = s0 - p0;
                const float s2 = b[c].x + a[c].y, s3 = b[c].y - a[c].x;
                const float p1 = s2 * m.y;
                b[c].x = fmaf(s3, m.x, p1); b[c].y = s2 - p1;
            } else {
                const float2 s = __fadd2_rn(a[c], b[c]);
                const float2 d = __fadd2_rn(a[c], make_float2(-b[c].x, -b[c].y));
                const float2 p = __fmul2_rn(s, m);
                a[c] = __ffma2_rn(d, m, p);
                b[c] = __fadd2_rn(s, make_float2(-p.x, -p.y));
                // 5 packed ops per chain and iteration would break the 2:1:1 ratio; one more add keeps it
                a[c] = __fadd2_rn(a[c], b[c]);
                const float2 p2 = __fmul2_rn(a[c], m);
                b[c] = __ffma2_rn(b[c], m, p2);
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < kChains; ++c) acc += a[c].x + a[c].y + b[c].x + b[c].y;
    if (acc == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;   // keep the chains alive
}

// flop per thread and iteration of each kind
static double flops_per_iter(int kind) {
    switch (kind) {
        case 0: return kChains * 4 * 2.0;                  // 4 FFMA
        case 1: return kChains * 2 * 4.0;                  // 2 FFMA2
        case 2: return kChains * (4 * 1.0 + 2 * 1.0 + 2 * 2.0);   // 4 FADD, 2 FMUL, 2 FFMA
        default: return kChains * (4 * 2.0 + 2 * 2.0 + 2 * 4.0);  // 4 FADD2, 2 FMUL2, 2 FFMA2
    }
}

}  // namespace aud

using namespace aud;

extern "C" AUD_API int32_t aud_measure_fp32(int32_t device, int32_t kind, double *tflops, double *ginst_per_s) {
    if (!tflops || kind < 0 || kind > 3) return fail(AUD_ERR_INVALID, "aud_measure_fp32: bad argument");
    cudaError_t e = cudaSetDevice(device);
    cudaDeviceProp prop{};
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return failf(AUD_ERR_CUDA, "no usable CUDA device %d: %s", device, cudaGetErrorString(e));
    float *d = nullptr;
    const int grid = prop.multiProcessorCount * 4, block = 512, iters = 4096;
    if ((e = cudaMalloc(&d, (size_t)grid * block * sizeof(float))) != cudaSuccess)
        return failf(AUD_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto launch = [&](int n) {
        switch (kind) {
            case 0: fp32_loop_kernel<0><<<grid, block>>>(d, n, 1e-6f); break;
            case 1: fp32_loop_kernel<1><<<grid, block>>>(d, n, 1e-6f); break;
            case 2: fp32_loop_kernel<2><<<grid, block>>>(d, n, 1e-6f); break;
            default: fp32_loop_kernel<3><<<grid, block>>>(d, n, 1e-6f); break;
        }
    };
    launch(256);   // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5 && e == cudaSuccess; ++rep) {
        cudaEventRecord(e0);
        launch(iters);
        cudaEventRecord(e1);
        e = cudaEventSynchronize(e1);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e == cudaSuccess && ms > 0.f) {
            const double fl = flops_per_iter(kind) * (double)iters * grid * block / (ms * 1e-3);
            if (fl > best) best = fl;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return failf(AUD_ERR_CUDA, "aud_measure_fp32 failed: %s", cudaGetErrorString(e));
    *tflops = best / 1e12;
    if (ginst_per_s) {
        const double flop_per_inst = (kind == 0) ? 2.0 : (kind == 1) ? 4.0 : (kind == 2) ? 10.0 / 8.0 : 20.0 / 8.0;
        *ginst_per_s = best / flop_per_inst / 1e9;   // lane-instructions per second
    }
    return AUD_OK;
}
