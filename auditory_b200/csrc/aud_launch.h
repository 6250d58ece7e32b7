// aud_launch.h -- the launch shapes of the fused kernel that are compiled in, one translation unit each
// (aud_fused_variant.cu built once per shape by the Makefile, in parallel).  A shape is
// (FFT warps, epilogue warps, records-by-epilogue); run_fused_once picks among exactly these.
#ifndef AUD_LAUNCH_H_
#define AUD_LAUNCH_H_

#include <cuda_runtime.h>

namespace aud {

struct KParams;

// X(NW, NE, ER): every shape the host side may select.  Plain log-mel and gabor launches run NW + 4 warps
// (ER = 1 only for plain log-mel, where the epilogue warps have time to write the frame-pair records);
// MFCC / smoothing / Energy launches run NW + 6.
#define AUD_FUSED_SHAPES(X) \
    X(12, 4, 1) X(12, 4, 0) X(10, 4, 1) X(10, 4, 0) X(8, 4, 1) X(8, 4, 0) X(6, 4, 1) X(6, 4, 0) \
    X(10, 6, 0) X(8, 6, 0) X(6, 6, 0)

#define AUD_DECLARE_SHAPE(NW, NE, ER) \
    cudaError_t launch_fused_##NW##_##NE##_##ER(const KParams &kp, int grid, size_t smem, cudaStream_t st);
AUD_FUSED_SHAPES(AUD_DECLARE_SHAPE)
#undef AUD_DECLARE_SHAPE

}  // namespace aud

#endif  // AUD_LAUNCH_H_
