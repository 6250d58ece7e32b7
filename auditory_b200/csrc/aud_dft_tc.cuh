// aud_dft_tc.cuh -- frame power for any window length on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same arithmetic as dft_power_kernel (aud_generic.cuh): |X[k]|^2 of every distinct frame as a folded real DFT
// (dft/dft.go:42-59; gonum CmplxFFT = forward unnormalised DFT),
//     Re X[k] = sum_h e[h] cos(2 pi h k / N),   Im X[k] = -/+ sum_h o[h] sin(2 pi h k / N),
//     e[h] = x[h] + x[N-h], o[h] = x[h] - x[N-h]   (h = 1 .. (N-1)/2; e[0] = x[0]; e[N/2] = x[N/2] for even N),
// as two GEMMs [frames x H] . [H x bins] on tcgen05.mma kind::f16 with FP16 operands and FP32 accumulators in
// tensor memory.  FP32 accuracy comes from splitting every operand in two FP16 slices (11 significant bits each,
// remainder exact) and issuing three slice products:
//     main accumulator   a0.c0
//     corr accumulator   a0.c1 + a1.c0                  (<= 2^-10 of main; the dropped a1.c1 is <= 2^-22)
// FP16 has a 5-bit exponent, so the samples are scaled: a small kernel first finds max |x| over the samples of every
// frame (finite values only) and the producers multiply by the power of two that brings it to [2^13, 2^14); the
// epilogue multiplies re and im by its inverse before squaring.  Samples more than 2^17 below their frame's maximum
// lose relative (not absolute) precision in the second slice -- an absolute error of 2^-40 of the maximum.  Table
// entries (|c| <= 1) are split on the host from float64; their second slice is exact to 2^-25 absolute.
// (Three BF16 slices with six products were the first parity-green version: same accuracy, twice the MMAs and
// 1.5x the operand bytes.  Two BF16 slices are not enough: log-mel errors of 5e-4 in quiet bands.)
// The tensor core adds into its FP32 accumulator with truncation, about half an ulp of the running sum per MMA;
// keeping the small terms in an accumulator of their own and K = 16 per instruction leaves H/16 truncations on the
// main sum, which measures the same as the FP32 FMA kernel (1.8e-6 of the peak power; a TF32 hi/lo version with one
// accumulator per output was 10x worse and failed the parity tests).  The epilogue adds main + corr in FP32.
//
// One CTA per SM, persistent over work items (128 frames x tn <= 128 bins), 22 warps:
//   warps 0-3    epilogue   tcgen05.ld the four accumulators (re/im x main/corr), unscale, re^2 + im^2, 128-bit stores
//   warps 4-19   operand A  8 frames each: load x[h], x[N-h] (coalesced along h; lane = two columns), fold, scale,
//                           split, store the 128-row x 64-column K-major SWIZZLE_128B FP16 blocks the MMA reads.
//                           Sixteen warps because a warp's chain load -> fold -> split -> store is serial and long:
//                           with eight warps of 16 frames the tensor pipe waited for operands most of the time.
//   warp 20      MMA        one thread: waits a stage, issues 4 k-steps x 3 MMAs (M128 N=tn K16), tcgen05.commit
//                           frees the stage / publishes the accumulators
//   warp 21      operand B  one thread: one 1-D TMA bulk copy of the pre-swizzled table block (2 slices) per stage
// Three stages of 64 KB (A: 2 slices x 16 KB; B: 2 slices x tn x 128 B), used round robin by the E/cos and O/sin
// blocks of successive k-blocks.  TMEM: 4 tn <= 512 columns.
#pragma once

#include <cuda_fp16.h>

#include "aud_generic.cuh"

namespace aud {
namespace tc {

constexpr int kTM = 128, kTNMax = 128, kTK = 64;
constexpr int kBlk = 128 * kTK * 2;                 // one 128-row x 64-column FP16 operand block: 16 KB
constexpr int kStage = 4 * kBlk, kStages = 3;       // A0 A1 | B0 B1 (B slices tn rows each)
constexpr int kEpiWarps = 4, kProdWarps = 16, kRowsPerWarp = kTM / kProdWarps;
constexpr int kThreads = (kEpiWarps + kProdWarps + 2) * 32;
constexpr int kTmemCols = 512;
constexpr size_t kSmemBytes = 1024 /*alignment slack*/ + kStages * (size_t)kStage + 80 /*barriers, TMEM slot*/ +
                              kTM * (sizeof(long long) + sizeof(int2) + sizeof(float));

struct TcParams {
    GParams g;
    const __half *tab;          // [cos/sin][n_nt][KB][slice 0..1][tn x 64, SWIZZLE_128B image]
    const float2 *row_scale;    // per frame row: power of two that brings its largest sample to [2^13, 2^14), and its inverse
    const int *row_job;         // per frame row of the scratch: its job (written by frame_scale_kernel; saves a binary search per row)
    int KB, n_nt, n_nt_tab, tn, n_items;   // n_nt: bin tiles computed (those somebody reads); n_nt_tab: bin tiles of the table
};

// byte offset of (row, 4-byte word w of the row) in a K-major SWIZZLE_128B block (1024-byte aligned):
// 8-row groups of 1024 bytes, 128-byte rows, 16-byte chunk index XOR (row & 7)
__host__ __device__ __forceinline__ uint32_t swz128(int row, int w) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((w >> 2) ^ (row & 7)) << 4) | ((w & 3) << 2)));
}

// shared-memory matrix descriptor: SWIZZLE_128B, K-major, 8-row group stride 1024 B, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor bit layout): FP32 accumulate, FP16 x FP16 (format 0), both K-major, M = 128
__host__ __device__ __forceinline__ uint32_t instr_desc(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc),
        "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void sts32(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

// (v0, v1) -> two words of packed FP16 pairs (v0 in the low half): v = s0 + s1 + O(2^-22 v), remainder exact
__device__ __forceinline__ void split2(float v0, float v1, uint32_t &s0, uint32_t &s1) {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(s0) : "f"(v1), "f"(v0));
    const float2 h = __half22float2(*reinterpret_cast<const __half2 *>(&s0));
    v0 -= h.x;
    v1 -= h.y;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(s1) : "f"(v1), "f"(v0));
}

// Largest finite |x| of every frame -> the power of two that makes its samples FP16 operands.  One CTA per job.
// Frames on a common hop grid (the usual case, P.dedupe): maxima of hop-sized blocks first (warp per block, coalesced),
// then every frame takes the maximum of the W = ceil(N / step) blocks it spans.  Otherwise a warp scans each frame.
// Also writes the frame-row -> job and segment -> job tables.
template <bool I16>
__global__ void __launch_bounds__(256) frame_scale_kernel(const __grid_constant__ GParams G, float2 *row_scale, int *row_job,
                                                          int *seg_job, float *blockmax, int W) {
    const KParams &P = G.k;
    const Job jb = P.jobs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, N = G.n_win;
    for (int f = tid; f < jb.nframes; f += 256) row_job[jb.frame_base + f] = (int)blockIdx.x;
    for (int c = tid; c < jb.nseg; c += 256) seg_job[jb.out_seg + c] = (int)blockIdx.x;
    auto sample_abs = [&](int i) -> float {
        if (i < 0 || i >= jb.utt_len) return 0.f;
        const float v = I16 ? (float)__ldg(static_cast<const short *>(P.wave) + jb.wave_off + i) * (1.0f / 32767.0f)
                            : __ldg(static_cast<const float *>(P.wave) + jb.wave_off + i);
        const float av = fabsf(v);
        return av <= 3.0e38f ? av : 0.f;   // NaN and Inf stay out of the maximum: they spoil their own frames only
    };
    auto to_scale = [](float m) -> float2 {
        int e = m > 0.f ? 13 - ilogbf(m) : 0;   // m * 2^e in [2^13, 2^14)
        e = max(-100, min(100, e));
        return make_float2(ldexpf(1.f, e), ldexpf(1.f, -e));
    };
    if (P.dedupe) {
        const int first0 = jb.seg0 * P.stride + P.add - P.border * P.step, nb = jb.nframes + W - 1;
        float *bm = blockmax + jb.frame_base + (size_t)blockIdx.x * (W - 1);
        for (int b = wid; b < nb; b += 8) {
            float m = 0.f;
            for (int i = lane; i < P.step; i += 32) m = fmaxf(m, sample_abs(first0 + b * P.step + i));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) bm[b] = m;
        }
        __syncthreads();
        for (int f = tid; f < jb.nframes; f += 256) {
            float m = 0.f;
            for (int b = f; b < f + W; ++b) m = fmaxf(m, bm[b]);
            row_scale[jb.frame_base + f] = to_scale(m);
        }
    } else {
        for (int f = wid; f < jb.nframes; f += 8) {
            const int c = f / P.S, i = f - c * P.S;
            const int first = (jb.seg0 + c) * P.stride + P.add + (i - P.border) * P.step;
            float m = 0.f;
            for (int n = lane; n < N; n += 32) m = fmaxf(m, sample_abs(first + n));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) row_scale[jb.frame_base + f] = to_scale(m);
        }
    }
}

template <bool I16>
__device__ __forceinline__ float ld_sample(const void *wave, long long idx) {
    if (I16) return (float)__ldg(static_cast<const short *>(wave) + idx) * (1.0f / 32767.0f);
    return __ldg(static_cast<const float *>(wave) + idx);
}

template <bool I16>
__global__ void __launch_bounds__(kThreads, 1) dft_power_tc_kernel(const __grid_constant__ TcParams T) {
    extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
    const uint32_t raw_s = smem_u32(tc_smem_raw);
    const uint32_t pad = ((raw_s + 1023u) & ~1023u) - raw_s;
    uint8_t *sm = tc_smem_raw + pad;
    const uint32_t sm_s = raw_s + pad;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + kStages * kStage);
    uint64_t *full = bars, *empty = bars + kStages, *acc_full = bars + 2 * kStages, *acc_empty = bars + 2 * kStages + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 2);
    long long *m_base = reinterpret_cast<long long *>(bars + 10);
    int2 *m_rng = reinterpret_cast<int2 *>(m_base + kTM);
    float *m_scale = reinterpret_cast<float *>(m_rng + kTM);

    const GParams &G = T.g;
    const KParams &P = G.k;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = G.n_win, H = G.bins, KB = T.KB, n_nt = T.n_nt, tn = T.tn;

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full[i], kProdWarps + 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, kEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kEpiWarps + kProdWarps) {   // the MMA warp owns the tensor-memory allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

    if (warp < kEpiWarps) {
        // ---- epilogue: lanes 32 q .. 32 q + 31 of each accumulator belong to warp q ----
        // accumulator columns: [re main | re corr | im main | im corr] x tn
        uint32_t ic = 0;
        for (int item = blockIdx.x; item < T.n_items; item += gridDim.x, ++ic) {
            const int mt = item / n_nt, nt = item - mt * n_nt;
            mbar_wait(acc_full, ic & 1);
            tc_fence_after();
            const int row = mt * kTM + warp * 32 + lane;
            float *dst = G.rawpow + (size_t)row * G.pitch + nt * tn;
            const float inv = row < G.total_frames ? T.row_scale[row].y : 0.f;   // undo the operand scale
            const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
            for (int cc = 0; cc < tn; cc += 8) {
                float rm[8], rc[8], im[8], ic2[8];
                tmem_ld8(ta + cc, rm);
                tmem_ld8(ta + tn + cc, rc);
                tmem_ld8(ta + 2 * tn + cc, im);
                tmem_ld8(ta + 3 * tn + cc, ic2);
                tmem_ld_wait();
                if (row < G.total_frames) {
#pragma unroll
                    for (int j = 0; j < 8; j += 4) {
                        float pw[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float re = (rm[j + u] + rc[j + u]) * inv, ii = (im[j + u] + ic2[j + u]) * inv;
                            pw[u] = fmaf(re, re, ii * ii);
                        }
                        *reinterpret_cast<float4 *>(dst + cc + j) = make_float4(pw[0], pw[1], pw[2], pw[3]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
    } else if (warp < kEpiWarps + kProdWarps) {
        // ---- operand A: this warp folds and splits rows 16 pw .. 16 pw + 15 of the item; lane = columns 2 lane, 2 lane + 1
        const int pw = warp - kEpiWarps, wr = pw * kRowsPerWarp;
        uint32_t st = 0, ph = 0;   // stage slot and its phase, in step with the MMA and table threads
        // word offset of this lane inside a 128-byte row: chunk (lane >> 2) is XORed with (row & 7) per row
        const uint32_t lane_w = (uint32_t)((lane & 3) << 2);
        for (int item = blockIdx.x; item < T.n_items; item += gridDim.x) {
            const int mt = item / n_nt;
            bool whole = false;   // this lane's row lies fully inside its utterance
            if (lane < kRowsPerWarp) {
                const int r = mt * kTM + wr + lane;
                long long base = 0;
                int nlo = 0, nhi = 0;
                float sc = 0.f;
                if (r < G.total_frames) {
                    const int ji = T.row_job[r];
                    const Job jb = P.jobs[ji];
                    sc = T.row_scale[r].x;
                    const int f = r - jb.frame_base;
                    if (f < jb.nframes) {
                        int first;
                        if (P.dedupe) first = jb.seg0 * P.stride + P.add - P.border * P.step + f * P.step;
                        else {
                            const int c = f / P.S, i = f - c * P.S;
                            first = (jb.seg0 + c) * P.stride + P.add + (i - P.border) * P.step;
                        }
                        base = jb.wave_off + first;
                        nlo = max(0, -first);                 // front padding (sndenv.go:443-450)
                        nhi = max(nlo, min(N, jb.utt_len - first));
                    }
                }
                m_base[wr + lane] = base;
                m_rng[wr + lane] = make_int2(nlo, nhi - nlo);
                m_scale[wr + lane] = sc;
                whole = (nlo == 0 && nhi == N);
            }
            const bool rows_whole = __all_sync(0xffffffffu, whole || lane >= kRowsPerWarp);
            __syncwarp();
            for (int kb = 0; kb < KB; ++kb) {
                const int h0 = kb * kTK + 2 * lane;
                // interior k-block: every column is a plain pair 1 <= h < N - h (no h = 0, no h = N/2, no h >= H)
                const bool interior = rows_whole && kb > 0 && (kb * kTK + kTK - 1) < (N + 1) / 2;
                float ev[kRowsPerWarp][2], ov[kRowsPerWarp][2];
                if (interior) {
#pragma unroll
                    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
                        const long long base = m_base[wr + rr];
                        const float a0 = ld_sample<I16>(P.wave, base + h0), a1 = ld_sample<I16>(P.wave, base + h0 + 1);
                        const float b0 = ld_sample<I16>(P.wave, base + N - h0), b1 = ld_sample<I16>(P.wave, base + N - h0 - 1);
                        const float sc = m_scale[wr + rr];
                        ev[rr][0] = (a0 + b0) * sc; ev[rr][1] = (a1 + b1) * sc;
                        ov[rr][0] = (a0 - b0) * sc; ov[rr][1] = (a1 - b1) * sc;
                    }
                } else {
#pragma unroll
                    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
                        const long long base = m_base[wr + rr];
                        const int2 rg = m_rng[wr + rr];
                        const float sc = m_scale[wr + rr];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int h = h0 + u, n2 = N - h;
                            const bool in = h < H, edge = (h == 0) || (2 * h == N);
                            const bool va = in && (unsigned)(h - rg.x) < (unsigned)rg.y;
                            const bool vb = in && !edge && (unsigned)(n2 - rg.x) < (unsigned)rg.y;
                            const float a = va ? ld_sample<I16>(P.wave, base + h) : 0.f;
                            const float b = vb ? ld_sample<I16>(P.wave, base + n2) : 0.f;
                            ev[rr][u] = (a + b) * sc;
                            ov[rr][u] = edge ? 0.f : (a - b) * sc;
                        }
                    }
                }
#pragma unroll
                for (int par = 0; par < 2; ++par) {
                    const uint32_t blk = sm_s + st * (uint32_t)kStage;
                    mbar_wait(&empty[st], ph ^ 1);
#pragma unroll
                    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
                        uint32_t s0, s1;
                        if (par == 0) split2(ev[rr][0], ev[rr][1], s0, s1);
                        else split2(ov[rr][0], ov[rr][1], s0, s1);
                        const uint32_t off = (uint32_t)(((wr + rr) >> 3) * 1024 + (rr & 7) * 128) +
                                             (uint32_t)(((lane >> 2) ^ (rr & 7)) << 4) + lane_w;
                        sts32(blk + off, s0);
                        sts32(blk + kBlk + off, s1);
                    }
                    fence_proxy_async();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[st]);
                    if (++st == kStages) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == kEpiWarps + kProdWarps) {
        // ---- MMA issue: one thread ----
        if (lane == 0) {
            const uint32_t idesc = instr_desc(tn);
            const uint32_t bsl = (uint32_t)tn * 128u;   // bytes of one B slice
            uint32_t st = 0, ph = 0, ic = 0;
            for (int item = blockIdx.x; item < T.n_items; item += gridDim.x, ++ic) {
                mbar_wait(acc_empty, (ic & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                    for (int par = 0; par < 2; ++par) {
                        mbar_wait(&full[st], ph);
                        tc_fence_after();
                        const uint32_t sa = sm_s + st * (uint32_t)kStage, sb = sa + 2 * kBlk;
                        const uint32_t d_main = tmem + (uint32_t)(par * 2 * tn), d_corr = d_main + (uint32_t)tn;
#pragma unroll
                        for (int k = 0; k < kTK / 16; ++k) {
                            const uint32_t ko = (uint32_t)k * 32u;
                            const uint64_t a0 = smem_desc(sa + ko), a1 = smem_desc(sa + kBlk + ko);
                            const uint64_t b0 = smem_desc(sb + ko), b1 = smem_desc(sb + bsl + ko);
                            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                            mma_f16(d_corr, a1, b0, idesc, acc);
                            mma_f16(d_corr, a0, b1, idesc, 1u);
                            mma_f16(d_main, a0, b0, idesc, acc);
                        }
                        mma_commit(&empty[st]);   // the stage is free once these MMAs have read it
                        if (++st == kStages) { st = 0; ph ^= 1; }
                    }
                }
                mma_commit(acc_full);
            }
        }
    } else {
        // ---- operand B: one thread, one bulk copy (two slices) per stage ----
        if (lane == 0) {
            const uint32_t bytes = 2u * (uint32_t)tn * 128u;
            uint32_t st = 0, ph = 0;
            for (int item = blockIdx.x; item < T.n_items; item += gridDim.x) {
                const int nt = item % n_nt;
                for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                    for (int par = 0; par < 2; ++par) {
                        mbar_wait(&empty[st], ph ^ 1);
                        mbar_expect_tx(&full[st], bytes);
                        const uint8_t *src = reinterpret_cast<const uint8_t *>(T.tab) + ((size_t)(par * T.n_nt_tab + nt) * KB + kb) * bytes;
                        tma_load_1d(sm + st * kStage + 2 * kBlk, src, bytes, &full[st]);
                        if (++st == kStages) { st = 0; ph ^= 1; }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kEpiWarps + kProdWarps) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
    }
}

}   // namespace tc
}   // namespace aud
