// aud_generic.cuh -- the same path for any window length (WinSamples != 400: other sample rates,
// other window durations).  The fused kernel in aud_kernels.cuh is specialised for the
// 400-sample window of the default 16 kHz configuration; this file is the general route behind the
// same C-ABI.  Two kernels, with the per-frame power spectrum as the only intermediate in HBM:
//
//   dft_power_kernel          |X[k]|^2 of every distinct frame (dft/dft.go:42-85; gonum CmplxFFT =
//                             forward unnormalised DFT) as a folded real DFT: with
//                             e[h] = x[h] + x[N-h], o[h] = x[h] - x[N-h]  (h = 1 .. (N-1)/2; e[0] = x[0],
//                             e[N/2] = x[N/2] for even N), Re X[k] = sum_h e[h] cos(2 pi h k / N) and
//                             Im X[k] = -sum_h o[h] sin(2 pi h k / N): a register-tiled FP32 product of the
//                             folded frames with an exact (float64-built) cos / sin table -- valid for
//                             prime lengths such as the 1103 samples of 25 ms at 44.1 kHz.
//   segment_features_kernel   per segment: banded mel sums of the raw power, the Prev/Cur recurrence
//                             over the steps (linear, so applied to the sums), ln, Energy, and the shared
//                             tile stage (cepstrum, deltas, gabor) of aud_kernels.cuh.
#pragma once

#include "aud_kernels.cuh"

namespace aud {

struct GParams {
    KParams k;              // geometry, scalars, wave / jobs / outputs (the fused kernel's fields that apply)
    int n_win, bins, pitch; // window length, n_win/2 + 1, row pitch of rawpow (multiple of 64; as wide as the bins computed)
    int tpitch;             // row pitch of the cos / sin tables (multiple of 64)
    int total_frames, njobs;
    const float *cos_t, *sin_t;   // [bins][pitch]: cos / sin(2 pi h k / n_win), zero for k >= bins
    const int *mel_lo, *mel_n;    // [n_mel] first bin and tap count of each filter
    const float *mel_w;           // [n_mel][mel_wpitch]
    int mel_wpitch;
    float *rawpow;                // [total_frames (+ padding)][pitch]
    int t_off[6];                 // float offsets of energy / mfcc / d1 / d2 / gabor tiles and of the DCT rows
    const int *seg_job;           // per output segment: its job (written by the tensor-core route's frame_scale_kernel), or null
    int stage_np, stage_wp;       // segment_features_kernel: pitch of the staged power rows (0: read them from global) and of the staged taps
};

// job that owns frame row `r` (jobs are sorted by frame_base)
__device__ __forceinline__ int job_of_frame(const Job *jobs, int njobs, int r) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].frame_base <= r) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}
// job that owns output segment `s` (jobs are sorted by out_seg)
__device__ __forceinline__ int job_of_segment(const Job *jobs, int njobs, long long s) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].out_seg <= s) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

constexpr int kGM = 64, kGN = 64, kGK = 16;   // frames x bins tile, folded samples per step

// grid (ceil(total_frames / 64), pitch / 64), 256 threads: thread (tx, ty) owns frames 4ty..4ty+3 and bins
// 4tx..4tx+3 of the tile.
__global__ void __launch_bounds__(256) dft_power_kernel(const __grid_constant__ GParams G) {
    __shared__ __align__(16) float es[kGK][kGM + 4], os[kGK][kGM + 4];
    __shared__ __align__(16) float cs[kGK][kGN], ss[kGK][kGN];
    __shared__ long long row_base[kGM];   // index of the frame's first sample in the wave buffer
    __shared__ int row_first[kGM], row_len[kGM];   // frame start relative to its utterance, utterance length (-1: no frame)
    const KParams &P = G.k;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int r0 = blockIdx.x * kGM, k0 = blockIdx.y * kGN;
    const int N = G.n_win, H = G.bins;

    if (tid < kGM) {
        const int r = r0 + tid;
        int first = 0, len = -1;
        long long base = 0;
        if (r < G.total_frames) {
            const Job jb = P.jobs[job_of_frame(P.jobs, G.njobs, r)];
            const int f = r - jb.frame_base;
            if (f < jb.nframes) {
                if (P.dedupe) first = jb.seg0 * P.stride + P.add - P.border * P.step + f * P.step;
                else {
                    const int c = f / P.S, i = f - c * P.S;
                    first = (jb.seg0 + c) * P.stride + P.add + (i - P.border) * P.step;
                }
                len = jb.utt_len;
                base = jb.wave_off + first;
            }
        }
        row_base[tid] = base; row_first[tid] = first; row_len[tid] = len;
    }
    __syncthreads();

    auto sample = [&](int row, int n) -> float {   // sample n of the tile's frame `row`, zero outside the utterance
        const int a = row_first[row] + n;
        if (a < 0 || a >= row_len[row]) return 0.f;   // front padding (sndenv.go:443-450) / no such frame
        const long long gi = row_base[row] + n;
        return P.in_i16 ? (float)__ldg(static_cast<const short *>(P.wave) + gi) * (1.0f / 32767.0f)
                        : __ldg(static_cast<const float *>(P.wave) + gi);
    };

    float re[4][4], im[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) re[r][c] = im[r][c] = 0.f;

    for (int h0 = 0; h0 < H; h0 += kGK) {
        // folded frames: consecutive threads walk h (coalesced), 16 rows per pass
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int hh = tid & 15, row = (tid >> 4) + 16 * p, h = h0 + hh;
            float e = 0.f, o = 0.f;
            if (h < H) {
                const float a = sample(row, h);
                if (h == 0 || 2 * h == N) e = a;
                else {
                    const float b = sample(row, N - h);
                    e = a + b;
                    o = a - b;
                }
            }
            es[hh][row] = e;
            os[hh][row] = o;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int kk = tid & 63, hh = (tid >> 6) + 4 * p, h = h0 + hh;
            float c = 0.f, s = 0.f;
            if (h < H) {
                c = __ldg(G.cos_t + (size_t)h * G.tpitch + k0 + kk);
                s = __ldg(G.sin_t + (size_t)h * G.tpitch + k0 + kk);
            }
            cs[hh][kk] = c;
            ss[hh][kk] = s;
        }
        __syncthreads();
#pragma unroll
        for (int hh = 0; hh < kGK; ++hh) {
            const float4 e4 = *reinterpret_cast<const float4 *>(&es[hh][4 * ty]);
            const float4 o4 = *reinterpret_cast<const float4 *>(&os[hh][4 * ty]);
            const float4 c4 = *reinterpret_cast<const float4 *>(&cs[hh][4 * tx]);
            const float4 s4 = *reinterpret_cast<const float4 *>(&ss[hh][4 * tx]);
            const float ev[4] = {e4.x, e4.y, e4.z, e4.w}, ov[4] = {o4.x, o4.y, o4.z, o4.w};
            const float cv[4] = {c4.x, c4.y, c4.z, c4.w}, sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    re[r][c] = fmaf(ev[r], cv[c], re[r][c]);
                    im[r][c] = fmaf(ov[r], sv[c], im[r][c]);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int row = r0 + 4 * ty + r;
        if (row < G.total_frames) {
            float4 pw;
            pw.x = fmaf(re[r][0], re[r][0], im[r][0] * im[r][0]);
            pw.y = fmaf(re[r][1], re[r][1], im[r][1] * im[r][1]);
            pw.z = fmaf(re[r][2], re[r][2], im[r][2] * im[r][2]);
            pw.w = fmaf(re[r][3], re[r][3], im[r][3] * im[r][3]);
            *reinterpret_cast<float4 *>(G.rawpow + (size_t)row * G.pitch + k0 + 4 * tx) = pw;
        }
    }
}

// One CTA per output segment, 128 threads; dynamic shared memory: the segment's tiles, the DCT rows and
// one done-list entry.
__global__ void __launch_bounds__(128) segment_features_kernel(const __grid_constant__ GParams G) {
    extern __shared__ __align__(16) float gsm[];
    const KParams &P = G.k;
    const int et = threadIdx.x, ENT = blockDim.x;
    const int S = P.S, M = P.n_mel, MS = M * S;
    const long long seg = blockIdx.x;
    const Job jb = P.jobs[G.seg_job ? G.seg_job[seg] : job_of_segment(P.jobs, G.njobs, seg)];
    const int c = (int)(seg - jb.out_seg);
    const int nv = valid_steps(jb.utt_len, P.add, P.stride, P.step, P.border, S, jb.seg0 + c, G.n_win);
    const float *rows = G.rawpow + (size_t)(jb.frame_base + c * P.seg_adv) * G.pitch;   // step i: rows + i * pitch

    TileSet t;
    t.mel = gsm;
    t.energy = gsm + G.t_off[0];
    t.mfcc = gsm + G.t_off[1];
    t.d1 = gsm + G.t_off[2];
    t.d2 = gsm + G.t_off[3];
    t.gab = gsm + G.t_off[4];
    float *dct_sm = gsm + G.t_off[5];
    float *gw_sm = dct_sm + P.dct_floats;
    int4 *done = reinterpret_cast<int4 *>(gw_sm + P.gw_floats);
    load_gabor_weights(P, gw_sm, et, ENT);

    if (P.want_mfcc) {
        const int M4 = ((M + 3) >> 2) << 2;
        for (int i = et; i < P.dct_floats; i += ENT) {
            const int k = i / M4, m = i - k * M4;
            dct_sm[i] = m < M ? P.dct[k * M + m] : 0.f;
        }
    }
    if (et == 0) done[0] = make_int4((int)seg, nv, 0, 0);
    // (a) banded sums of the raw linear power (mel.go:120-131); lanes walk the filters of one step.  The bins the bank
    // reads and the taps are first staged in shared memory with coalesced loads: straight from global memory every
    // lane of a warp reads its own sector (32 filters = 32 places in the row) and the L1 is what binds.
    float *prow = reinterpret_cast<float *>(done + 1);   // [S][stage_np]
    float *wsm = prow + (size_t)S * G.stage_np;          // [M][stage_wp]
    if (G.stage_np > 0) {
        const int p4 = G.stage_np >> 2, n4 = nv * p4;
        for (int i0 = et; i0 < n4; i0 += 8 * ENT) {   // eight independent 128-bit loads in flight per thread, then the stores
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int idx = i0 + k * ENT;
                if (idx < n4) {
                    const int r = idx / p4, c4 = idx - r * p4;
                    v[k] = __ldg(reinterpret_cast<const float4 *>(rows + (size_t)r * G.pitch) + c4);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int idx = i0 + k * ENT;
                if (idx < n4) reinterpret_cast<float4 *>(prow)[idx] = v[k];   // row r starts at r * p4 float4s
            }
        }
        for (int idx = et; idx < M * G.stage_wp; idx += ENT) {
            const int m = idx / G.stage_wp, q = idx - m * G.stage_wp;
            wsm[idx] = q < G.mel_wpitch ? __ldg(G.mel_w + (size_t)m * G.mel_wpitch + q) : 0.f;
        }
        __syncthreads();
        for (int idx = et; idx < MS; idx += ENT) {
            const int i = idx / M, m = idx - i * M;
            float sum = 0.f;
            if (i < nv) {
                const float *pr = prow + (size_t)i * G.stage_np + G.mel_lo[m];
                const float *w = wsm + m * G.stage_wp;
                const int n = G.mel_n[m];
                for (int q = 0; q < n; ++q) sum = fmaf(w[q], pr[q], sum);
            }
            t.mel[m * S + i] = sum;
        }
    } else {
        for (int idx = et; idx < MS; idx += ENT) {
            const int i = idx / M, m = idx - i * M;
            float sum = 0.f;
            if (i < nv) {
                const float *pr = rows + (size_t)i * G.pitch + G.mel_lo[m];
                const float *w = G.mel_w + (size_t)m * G.mel_wpitch;
                const int n = G.mel_n[m];
                for (int q = 0; q < n; ++q) sum = fmaf(__ldg(w + q), pr[q], sum);
            }
            t.mel[m * S + i] = sum;
        }
    }
    __syncthreads();
    // (b) P_s = Prev * P_{s-1} + Cur * p_s restarting at step 0 (dft.go:62-72), on the sums; then ln
    // (the recurrence is serial per filter, the logarithms are not: they are taken by all threads afterwards)
    for (int m = et; m < M; m += ENT) {
        float y = 0.f;
        for (int i = 0; i < nv; ++i) {
            const float x = t.mel[m * S + i];
            y = (i == 0) ? x : fmaf(P.prev, y, P.cur * x);   // also with Prev = 0: 0 * NaN carries a spoiled step on (dft.go:66-68)
            t.mel[m * S + i] = y;
        }
    }
    __syncthreads();
    for (int idx = et; idx < MS; idx += ENT) {
        const int m = idx / S, i = idx - m * S;
        t.mel[idx] = i < nv ? finish_mel(P, t.mel[idx]) : 0.f;
    }
    // (c) Energy[s] = sum over steps f of LogPowerSegment.Values[s*S + f]: bin s (sndenv.go:360-366)
    if (P.energy_bins > 0) {
        for (int sb = et; sb < S; sb += ENT) {
            float y = 0.f, esum = 0.f;
            if (P.comp_log_pow && sb < G.bins) {
                for (int i = 0; i < nv; ++i) {
                    const float x = rows[(size_t)i * G.pitch + sb];
                    y = (i == 0) ? x : fmaf(P.prev, y, P.cur * x);
                    const float qv = y + P.log_off;
                    esum += (qv == 0.f) ? P.log_min : (P.log1p_path ? log1pf(y) : logf(qv));
                }
            }
            if (P.o_energy) P.o_energy[(size_t)seg * S + sb] = esum;
            t.energy[sb] = esum;
        }
    }
    __syncthreads();
    finish_tiles(P, t, dct_sm, gw_sm, done, 1, et, ENT, P.o_mel != nullptr, [] { __syncthreads(); });
}

// agabor.Convolve as an operator of its own (agabor/gabor.go:225-315): one CTA per [n_mel][S] input tensor,
// which is staged in shared memory and handed to the tile stage with only the gabor part switched on.
__global__ void __launch_bounds__(128) gabor_convolve_kernel(const __grid_constant__ KParams P, const float *mel_in) {
    extern __shared__ __align__(16) float gsm[];
    const int et = threadIdx.x, ENT = blockDim.x, MS = P.n_mel * P.S;
    float *tile = gsm;
    float *gab = tile + ((MS + 3) & ~3);
    float *gw = gab + ((P.g_len + 3) & ~3);
    int4 *done = reinterpret_cast<int4 *>(gw + P.gw_floats);
    for (int i = et; i < MS; i += ENT) tile[i] = mel_in[(size_t)blockIdx.x * MS + i];
    load_gabor_weights(P, gw, et, ENT);
    if (et == 0) done[0] = make_int4((int)blockIdx.x, P.S, 0, 0);
    __syncthreads();
    const TileSet t{tile, nullptr, nullptr, nullptr, nullptr, gab};
    finish_tiles(P, t, nullptr, gw, done, 1, et, ENT, false, [] { __syncthreads(); });
}

// ------------------------------------------------- per-step operators (gaborview-style callers)
// mel.Params.FilterDft (mel/mel.go:120-153) for every step of a segment: power [bins][S] (PowerSegment layout, step
// fastest) -> mel [n_mel][S].  filt is the reference's [n_mel][n_mel + 2] table read with flat stride arithmetic.
struct MelOpParams {
    int bins, S, n_mel;
    float log_off, log_min;
    int renorm;
    float renorm_min, renorm_scale;
    const int *bin_pts;
    const float *filt, *power;
    float *mel;
};
__global__ void mel_filter_dft_kernel(const __grid_constant__ MelOpParams Q) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Q.n_mel * Q.S) return;
    const int m = idx / Q.S, step = idx - m * Q.S;
    const int lo = Q.bin_pts[m], hi = Q.bin_pts[m + 2];
    float sum = 0.f;
    for (int b = lo, fi = 0; b <= hi; ++b, ++fi) sum = fmaf(Q.filt[m * (Q.n_mel + 2) + fi], Q.power[(size_t)b * Q.S + step], sum);
    sum += Q.log_off;
    float val = (sum == 0.f) ? Q.log_min : logf(sum);
    if (Q.renorm) val = fminf(fmaxf(val - Q.renorm_min, 0.f) * Q.renorm_scale, 1.f);
    Q.mel[idx] = val;
}
// mel.Params.CepstrumDct (mel/mel.go:192-212) for every step: mel [n_mel][S] -> mfcc [n_coefs][S]; dct [n_coefs][n_mel]
// is the matrix of gonum's DCT.Transform; coefficient 0 becomes ln(1 + y0^2) (mel.go:203-204).
__global__ void cepstrum_dct_kernel(const float *mel, const float *dct, int n_mel, int S, int n_coefs, float *mfcc) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_coefs * S) return;
    const int k = idx / S, step = idx - k * S;
    float y = 0.f;
    for (int m = 0; m < n_mel; ++m) y = fmaf(dct[k * n_mel + m], mel[(size_t)m * S + step], y);
    mfcc[idx] = (k == 0) ? log1pf(y * y) : y;
}

// ------------------------------------------------- power / log-power outputs
// Parity / inspection path only (PowerSegment, LogPowerSegment: dft/dft.go:62-85):
// rebuilds the per-segment smoothed power from the raw per-frame power the
// fused kernel left in `rawpow`.  One CTA per job.
struct PowParams {
    int step, stride, S, border, add, seg_adv;
    int n_win, bins, pitch;   // window length, bins per frame, row pitch of rawpow
    float prev, cur, log_off, log_min;
    int comp_log_pow, log1p_path;
    const Job *jobs;
    const float *rawpow;
    float *o_power, *o_logpower;
};

__global__ void power_segments_kernel(const __grid_constant__ PowParams Q) {
    const Job jb = Q.jobs[blockIdx.x];
    for (int r = threadIdx.x; r < jb.nseg * Q.bins; r += blockDim.x) {
        const int c = r / Q.bins, k = r - c * Q.bins;
        const int nv = valid_steps(jb.utt_len, Q.add, Q.stride, Q.step, Q.border, Q.S, jb.seg0 + c, Q.n_win);
        const size_t base = ((size_t)(jb.out_seg + c) * Q.bins + k) * Q.S;
        float y = 0.f;
        for (int i = 0; i < Q.S; ++i) {
            float pw = 0.f, lp = 0.f;
            if (i < nv) {
                const float x = Q.rawpow[(size_t)(jb.frame_base + c * Q.seg_adv + i) * Q.pitch + k];
                y = (i == 0) ? x : fmaf(Q.prev, y, Q.cur * x);
                pw = y;
                if (Q.comp_log_pow) {
                    const float qv = y + Q.log_off;
                    lp = (qv == 0.f) ? Q.log_min : (Q.log1p_path ? log1pf(y) : logf(qv));
                }
            }
            if (Q.o_power) Q.o_power[base + i] = pw;
            if (Q.o_logpower) Q.o_logpower[base + i] = lp;
        }
    }
}

}  // namespace aud
