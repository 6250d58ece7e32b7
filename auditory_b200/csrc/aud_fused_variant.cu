// aud_fused_variant.cu -- one launch shape of the fused kernel per object file:
//   nvcc -DAUD_NW=12 -DAUD_NE=4 -DAUD_ER=1 -c aud_fused_variant.cu -o fused_12_4_1.o
// (see aud_launch.h for the list and the Makefile for the loop).
#include "aud_kernels.cuh"
#include "aud_launch.h"

#if !defined(AUD_NW) || !defined(AUD_NE) || !defined(AUD_ER)
#error "compile with -DAUD_NW=<fft warps> -DAUD_NE=<epilogue warps> -DAUD_ER=<0|1>"
#endif

#define AUD_CAT_(a, b, c) launch_fused_##a##_##b##_##c
#define AUD_CAT(a, b, c) AUD_CAT_(a, b, c)

namespace aud {

cudaError_t AUD_CAT(AUD_NW, AUD_NE, AUD_ER)(const KParams &kp, int grid, size_t smem, cudaStream_t st) {
    auto *fn = fused_features_kernel<AUD_NW, AUD_NE, (AUD_ER != 0)>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    fn<<<grid, (AUD_NW + AUD_NE) * 32, smem, st>>>(kp);
    return cudaGetLastError();
}

}  // namespace aud
