// aud_kernels.cuh -- device code of the fused waveform -> mel / MFCC / gabor
// path for sm_100a.  One CTA owns a *chunk*: a run of consecutive segments of
// one utterance.  It stages the chunk's waveform span in shared memory once,
// transforms every distinct frame of the span exactly once (frames shared by
// overlapping segments are not recomputed), and then finishes each segment
// (temporal smoothing scan, logs, DCT, deltas, gabor) from on-chip data; only
// final features go to HBM.
//
// Reference semantics implemented here (file:line under the reference tree):
//   frame extraction   sound/sndenv.go:438-478  (front zero pad, tail error -> rest of segment zero)
//   DFT + power        dft/dft.go:42-85         (rectangular window, length-WinSamples DFT, |X|^2,
//                                                Prev/Cur smoothing, ln(p + LogOffSet))
//   mel filter bank    mel/mel.go:120-153       (banded sums of LINEAR power, ln, LogMin on exact zero)
//   cepstrum           mel/mel.go:192-212       (DCT-I as a [n_coefs x n_mel] matrix)
//   energy / c0        sound/sndenv.go:360-372  (transposed indexing quirk)
//   deltas             sound/sndenv.go:378-432
//   gabor              agabor/gabor.go:225-315
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace aud {

constexpr int kN = 400;             // FFT length the fused kernel is specialised for
constexpr int kBins = kN / 2 + 1;   // 201
constexpr int kRS = 21;             // exchange buffer: slot(k1, n2) = k1 + kRS*n2 (float2 units)
constexpr int kPS = 426;            // per-pair stride of the exchange buffer (float2 units), >= 20*kRS
constexpr int kPairsPerWarp = 3;    // 10 lanes per frame pair, lanes 30/31 idle in the FFT passes
constexpr int kMelPitch = 33;       // row pitch of the per-frame raw mel sums
constexpr int kPowPitch = 208;      // row pitch of the raw-power scratch (debug / parity outputs)

struct Chunk {
    long long wave_off;   // index of the utterance's first sample in the wave buffer
    long long out_seg;    // global index of this chunk's first segment
    int utt_len;
    int seg0;             // first segment of the chunk within its utterance
    int nseg;
    int frame_base;       // first row of this chunk in the raw-power scratch
};

struct KParams {
    // geometry
    int step, stride, S, border, add;
    int seg_adv;          // frame slots between consecutive segments: stride/step if frames are shared, else S
    int dedupe;           // 1: stride % step == 0, frame slot f starts at f*step within the span
    int n_mel, n_coefs;
    int wave_cap;         // floats reserved for the staged span (+ kN zero tail) and, later, the output tiles
    int max_frames;       // frame slots per chunk
    int max_segs;         // segments per chunk
    int energy_bins;      // low bins kept per frame for Energy (0 = not needed)
    // dft.Params / mel.FilterBank scalars
    float prev, cur, log_off, log_min;
    int comp_log_pow, log1p_path;
    float mel_log_off, mel_log_min;
    int renorm;
    float renorm_min, renorm_scale;
    int do_mfcc, do_deltas, c0_energy;
    // gabor
    int g_on, g_nf, g_sx, g_sy, g_stx, g_sty, g_dims, g_by_time, g_nt, g_nfy, g_tmaxstrides, g_len;
    int g_str0, g_str1, g_str2;
    float g_gain;
    // tables (device)
    const float2 *tw;       // [400] e^{-2 pi i m / 400}
    const int *mel_start;   // [n_mel] first bin of each filter
    const int *mel_width;   // [n_mel] taps per filter
    const float *mel_taps;  // [mel_maxw][n_mel]
    int mel_maxw;
    const float *dct;       // [n_coefs][n_mel]
    const float *gabor;     // [nf][sy][sx]
    // io (device)
    const float *wave;
    const Chunk *chunks;
    int n_chunks;
    float *o_mel, *o_mfcc, *o_d1, *o_d2, *o_energy, *o_gabor;   // any may be NULL
    float *rawpow;          // [frame rows][kPowPitch] raw |X|^2, only when power / logpower are requested
};

// ------------------------------------------------------------------ DFT-20
// Prime-factor (Good-Thomas) 4 x 5 DFT on 20 complex values held in registers.
// Input natural order; after the call the value for output index k sits in
// register slot perm20(k).
__host__ __device__ constexpr int perm20(int k) { return (5 * (k % 4) + 4 * (k % 5)) % 20; }

__device__ __forceinline__ void dft4(float &r0, float &i0, float &r1, float &i1, float &r2, float &i2, float &r3,
                                     float &i3) {
    const float ar = r0 + r2, ai = i0 + i2, br = r0 - r2, bi = i0 - i2;
    const float cr = r1 + r3, ci = i1 + i3, dr = r1 - r3, di = i1 - i3;
    r0 = ar + cr; i0 = ai + ci;
    r1 = br + di; i1 = bi - dr;      // b - i d
    r2 = ar - cr; i2 = ai - ci;
    r3 = br - di; i3 = bi + dr;      // b + i d
}

__device__ __forceinline__ void dft5(float &r0, float &i0, float &r1, float &i1, float &r2, float &i2, float &r3,
                                     float &i3, float &r4, float &i4) {
    constexpr float C1 = 0.30901699437494742410f;    // cos(2 pi / 5)
    constexpr float C2 = -0.80901699437494742410f;   // cos(4 pi / 5)
    constexpr float S1 = 0.95105651629515357212f;    // sin(2 pi / 5)
    constexpr float S2 = 0.58778525229247312917f;    // sin(4 pi / 5)
    const float t1r = r1 + r4, t1i = i1 + i4, t2r = r2 + r3, t2i = i2 + i3;
    const float t3r = r1 - r4, t3i = i1 - i4, t4r = r2 - r3, t4i = i2 - i3;
    const float m1r = fmaf(C2, t2r, fmaf(C1, t1r, r0)), m1i = fmaf(C2, t2i, fmaf(C1, t1i, i0));
    const float m2r = fmaf(C1, t2r, fmaf(C2, t1r, r0)), m2i = fmaf(C1, t2i, fmaf(C2, t1i, i0));
    const float s1r = fmaf(S2, t4r, S1 * t3r), s1i = fmaf(S2, t4i, S1 * t3i);
    const float s2r = fmaf(-S1, t4r, S2 * t3r), s2i = fmaf(-S1, t4i, S2 * t3i);
    r0 = r0 + t1r + t2r; i0 = i0 + t1i + t2i;
    r1 = m1r + s1i; i1 = m1i - s1r;   // m1 - i s1
    r4 = m1r - s1i; i4 = m1i + s1r;   // m1 + i s1
    r2 = m2r + s2i; i2 = m2i - s2r;   // m2 - i s2
    r3 = m2r - s2i; i3 = m2i + s2r;   // m2 + i s2
}

__device__ __forceinline__ void dft20(float (&xr)[20], float (&xi)[20]) {
    // input slot n = (5a + 4b) % 20: size-4 transforms over a, then size-5 over b
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const int n0 = (4 * b) % 20, n1 = (5 + 4 * b) % 20, n2 = (10 + 4 * b) % 20, n3 = (15 + 4 * b) % 20;
        dft4(xr[n0], xi[n0], xr[n1], xi[n1], xr[n2], xi[n2], xr[n3], xi[n3]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int n0 = (5 * c) % 20, n1 = (5 * c + 4) % 20, n2 = (5 * c + 8) % 20, n3 = (5 * c + 12) % 20,
                  n4 = (5 * c + 16) % 20;
        dft5(xr[n0], xi[n0], xr[n1], xi[n1], xr[n2], xi[n2], xr[n3], xi[n3], xr[n4], xi[n4]);
    }
}

// exchange-buffer slot of spectrum index k (k = k1 + 20 k2 -> k1 + kRS*k2)
__device__ __forceinline__ int zslot(int k) { return (k % 20) + kRS * (k / 20); }

__device__ __forceinline__ long long floordiv(long long a, long long b) {   // b > 0
    long long q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
}

// number of leading steps of segment `seg` whose window lies inside the signal
// (sndenv.go:457-460: the first window that runs past the end aborts the rest)
__device__ __forceinline__ int valid_steps(const KParams &P, const Chunk &ck, int seg) {
    const long long room = (long long)ck.utt_len - kN - P.add - (long long)seg * P.stride;
    const long long last = floordiv(room, P.step) + P.border;   // largest valid step index
    if (last < 0) return 0;
    return last + 1 > P.S ? P.S : (int)(last + 1);
}

// ------------------------------------------------------------ fused kernel
template <int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, (NWARPS <= 8 ? 2 : 1)) fused_features_kernel(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s_wave = reinterpret_cast<float *>(smem_raw);
    float2 *s_tw = reinterpret_cast<float2 *>(s_wave + P.wave_cap);
    float2 *s_scr = s_tw + kN;
    float *s_melraw = reinterpret_cast<float *>(s_scr + NWARPS * kPairsPerWarp * kPS);
    float *s_lowpow = s_melraw + P.max_frames * kMelPitch;

    constexpr int NT = NWARPS * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = lane / 10, j = lane - 10 * q;     // frame pair within the warp's triple, column group
    const bool fft_lane = lane < 30;

    for (int i = tid; i < kN; i += NT) s_tw[i] = P.tw[i];

    for (int ch = blockIdx.x; ch < P.n_chunks; ch += gridDim.x) {
        const Chunk ck = P.chunks[ch];
        const int nframes = P.dedupe ? (ck.nseg - 1) * P.seg_adv + P.S : ck.nseg * P.S;
        const int span = (ck.nseg - 1) * P.stride + (P.S - 1) * P.step + kN;
        const long long a0 = (long long)ck.seg0 * P.stride + P.add - (long long)P.border * P.step;

        // ---- stage the span (zero outside the utterance, kN zeros after it)
        {
            const float *src = P.wave + ck.wave_off;
            for (int i = tid; i < span + kN; i += NT) {
                const long long a = a0 + i;
                float v = 0.f;
                if (i < span && a >= 0 && a < ck.utt_len) v = __ldg(src + a);
                s_wave[i] = v;
            }
        }
        __syncthreads();

        // ---- phase 1: every frame slot once: 2 real frames per complex 400-point FFT (20 x 20)
        const int npairs = (nframes + 1) >> 1;
        const int ntriples = (npairs + kPairsPerWarp - 1) / kPairsPerWarp;
        float2 *scr_w = s_scr + warp * kPairsPerWarp * kPS;
        for (int t = warp; t < ntriples; t += NWARPS) {
            if (fft_lane) {
                int pair = t * kPairsPerWarp + q;
                if (pair >= npairs) pair = npairs - 1;        // duplicate work, results discarded below
                const int fa = 2 * pair, fb = fa + 1;
                const int offA = P.dedupe ? fa * P.step : (fa / P.S) * P.stride + (fa % P.S) * P.step;
                const int offB = fb >= nframes ? span
                                               : (P.dedupe ? fb * P.step : (fb / P.S) * P.stride + (fb % P.S) * P.step);
                float2 *scr = scr_w + q * kPS;
                float xr[20], xi[20];
                // pass 1: columns n2 = j, j+10: DFT-20 over n1 of z[20 n1 + n2], twiddle W400^{n2 k1}
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    const int n2 = j + 10 * c;
                    const float *pa = s_wave + offA + n2;
                    const float *pb = s_wave + offB + n2;
#pragma unroll
                    for (int n1 = 0; n1 < 20; ++n1) {
                        xr[n1] = pa[20 * n1];
                        xi[n1] = pb[20 * n1];
                    }
                    dft20(xr, xi);
                    float2 *e = scr + kRS * n2;
#pragma unroll
                    for (int k1 = 0; k1 < 20; ++k1) {
                        float yr = xr[perm20(k1)], yi = xi[perm20(k1)];
                        if (k1 > 0) {
                            const float2 w = s_tw[n2 * k1];
                            const float tr = yr * w.x - yi * w.y;
                            yi = fmaf(yr, w.y, yi * w.x);
                            yr = tr;
                        }
                        e[k1] = make_float2(yr, yi);
                    }
                }
            }
            __syncwarp();
            if (fft_lane) {
                float2 *scr = scr_w + q * kPS;
                float xr[20], xi[20];
                // pass 2: rows k1 = j, j+10: DFT-20 over n2 -> Z[k1 + 20 k2], written back in place
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    float2 *e = scr + j + 10 * c;
#pragma unroll
                    for (int n2 = 0; n2 < 20; ++n2) {
                        const float2 v = e[kRS * n2];
                        xr[n2] = v.x;
                        xi[n2] = v.y;
                    }
                    dft20(xr, xi);
#pragma unroll
                    for (int k2 = 0; k2 < 20; ++k2) e[kRS * k2] = make_float2(xr[perm20(k2)], xi[perm20(k2)]);
                }
            }
            __syncwarp();
            // split the packed spectrum: |X_A[k]|^2, |X_B[k]|^2 from Z[k], Z[N-k]; stored over Z[k]
            for (int qq = 0; qq < kPairsPerWarp; ++qq) {
                const int pair = t * kPairsPerWarp + qq;
                if (pair >= npairs) break;
                float2 *scr = scr_w + qq * kPS;
                const int fa = 2 * pair;
                const bool has_b = fa + 1 < nframes;
                for (int k = lane; k < kBins; k += 32) {
                    const float2 a = scr[zslot(k)];
                    const float2 b = scr[zslot(k == 0 ? 0 : kN - k)];
                    const float ar = a.x + b.x, ai = a.y - b.y;
                    const float br = a.y + b.y, bi = b.x - a.x;
                    const float pa = 0.25f * fmaf(ar, ar, ai * ai);
                    const float pb = 0.25f * fmaf(br, br, bi * bi);
                    scr[zslot(k)] = make_float2(pa, pb);
                    if (k < P.energy_bins) {
                        s_lowpow[fa * P.energy_bins + k] = pa;
                        if (has_b) s_lowpow[(fa + 1) * P.energy_bins + k] = pb;
                    }
                    if (P.rawpow) {
                        P.rawpow[(size_t)(ck.frame_base + fa) * kPowPitch + k] = pa;
                        if (has_b) P.rawpow[(size_t)(ck.frame_base + fa + 1) * kPowPitch + k] = pb;
                    }
                }
            }
            __syncwarp();
            // mel filter bank on the raw power (smoothing is linear and is applied to the sums later)
            for (int m = lane; m < P.n_mel; m += 32) {
                const int b0 = P.mel_start[m], w = P.mel_width[m];
                for (int qq = 0; qq < kPairsPerWarp; ++qq) {
                    const int pair = t * kPairsPerWarp + qq;
                    if (pair >= npairs) break;
                    const float2 *scr = scr_w + qq * kPS;
                    float sa = 0.f, sb = 0.f;
                    int row = b0 % 20, col = b0 / 20;
                    for (int i = 0; i < w; ++i) {
                        const float wt = __ldg(P.mel_taps + i * P.n_mel + m);
                        const float2 p = scr[row + kRS * col];
                        sa = fmaf(wt, p.x, sa);
                        sb = fmaf(wt, p.y, sb);
                        if (++row == 20) { row = 0; ++col; }
                    }
                    const int fa = 2 * pair;
                    s_melraw[fa * kMelPitch + m] = sa;
                    if (fa + 1 < nframes) s_melraw[(fa + 1) * kMelPitch + m] = sb;
                }
            }
            __syncwarp();
        }
        __syncthreads();

        // ---- phase 2: per-segment epilogue on tiles that alias the (now dead) waveform span
        const int C = ck.nseg, S = P.S, M = P.n_mel, NC = P.n_coefs;
        float *t_mel = s_wave;                       // [C][M][S]
        float *t_energy = t_mel + P.max_segs * M * S;    // [C][S]
        float *t_mfcc = t_energy + P.max_segs * S;       // [C][NC][S]
        float *t_d1 = t_mfcc + P.max_segs * NC * S;      // [C][NC][S]
        float *t_d2 = t_d1 + P.max_segs * NC * S;        // [C][NC][S]
        float *t_gab = t_d2 + P.max_segs * NC * S;       // [C][g_len]

        // (a) mel rows: first-order smoothing recurrence over the steps of each segment, then ln
        for (int r = tid; r < C * M; r += NT) {
            const int c = r / M, m = r - c * M;
            const int nv = valid_steps(P, ck, ck.seg0 + c);
            float y = 0.f;
            float *dst = t_mel + (c * M + m) * S;
            for (int i = 0; i < S; ++i) {
                float val = 0.f;
                if (i < nv) {
                    const float x = s_melraw[(c * P.seg_adv + i) * kMelPitch + m];
                    y = (i == 0) ? x : fmaf(P.prev, y, P.cur * x);
                    const float s = y + P.mel_log_off;
                    val = (s == 0.f) ? P.mel_log_min : logf(s);
                    if (P.renorm) {
                        val -= P.renorm_min;
                        if (val < 0.f) val = 0.f;
                        val *= P.renorm_scale;
                        if (val > 1.f) val = 1.f;
                    }
                }
                dst[i] = val;
            }
        }
        // (b) Energy[s] = sum over steps f of LogPowerSegment.Values[s*S + f]  (bin s, transposed quirk)
        if (P.energy_bins > 0) {
            for (int r = tid; r < C * S; r += NT) {
                const int c = r / S, s = r - c * S;
                const int nv = valid_steps(P, ck, ck.seg0 + c);
                float y = 0.f, e = 0.f;
                if (P.comp_log_pow) {
                    for (int i = 0; i < nv; ++i) {
                        const float x = s_lowpow[(c * P.seg_adv + i) * P.energy_bins + s];
                        y = (i == 0) ? x : fmaf(P.prev, y, P.cur * x);
                        const float qv = y + P.log_off;
                        e += (qv == 0.f) ? P.log_min : (P.log1p_path ? log1pf(y) : logf(qv));
                    }
                }
                t_energy[c * S + s] = e;
            }
        }
        if (P.g_on)
            for (int r = tid; r < C * P.g_len; r += NT) t_gab[r] = 0.f;
        __syncthreads();

        // (c) cepstrum: DCT-I rows 0..NC-1 of the log-mel column of each step
        if (P.do_mfcc) {
            for (int r = tid; r < C * NC * S; r += NT) {
                const int c = r / (NC * S), rem = r - c * NC * S, k = rem / S, i = rem - k * S;
                const int nv = valid_steps(P, ck, ck.seg0 + c);
                float v = 0.f;
                if (k == 0 && P.c0_energy) {
                    v = t_energy[c * S + i];
                } else if (i < nv) {
                    const float *col = t_mel + c * M * S + i;
                    const float *drow = P.dct + k * M;
                    float acc = 0.f;
                    for (int m = 0; m < M; ++m) acc = fmaf(__ldg(drow + m), col[m * S], acc);
                    v = (k == 0) ? log1pf(acc * acc) : acc;
                }
                t_mfcc[(c * NC + k) * S + i] = v;
            }
        }
        // (e) gabor: strided valid correlation of every filter with the segment's mel tile
        if (P.g_on) {
            const int per_seg = P.g_nt * P.g_nfy * P.g_nf;
            for (int r = tid; r < C * per_seg; r += NT) {
                const int c = r / per_seg;
                int rem = r - c * per_seg;
                const int ti = rem / (P.g_nfy * P.g_nf);
                rem -= ti * P.g_nfy * P.g_nf;
                const int fi = rem / P.g_nf, flt = rem - fi * P.g_nf;
                const float *tile = t_mel + c * M * S + (fi * P.g_sty) * S + ti * P.g_stx;
                const float *gf = P.gabor + flt * P.g_sy * P.g_sx;
                float acc = 0.f;
                for (int ff = 0; ff < P.g_sy; ++ff)
                    for (int ft = 0; ft < P.g_sx; ++ft) {
                        float iv = tile[ff * S + ft];
                        if (iv != iv) iv = 0.5f;
                        acc = fmaf(__ldg(gf + ff * P.g_sx + ft), iv, acc);
                    }
                const bool pos = acc >= 0.f;
                const float act = P.g_gain * fabsf(acc);
                int on_off, off_off;
                if (P.g_dims == 2) {
                    const int x = P.g_by_time ? ti + P.g_tmaxstrides * flt : flt + ti * P.g_nf;
                    on_off = (2 * fi) * P.g_str0 + x;
                    off_off = on_off + P.g_str0;
                } else {
                    on_off = fi * P.g_str0 + ti * P.g_str1 + flt;
                    off_off = on_off + P.g_str2;
                }
                float *g = t_gab + c * P.g_len;
                g[on_off] = pos ? act : 0.f;
                g[off_off] = pos ? 0.f : act;
            }
        }
        __syncthreads();
        // (d) deltas and delta-deltas with the reference's accumulator quirk (prv/nxt carried across coefficients)
        if (P.do_mfcc && P.do_deltas) {
            for (int pass = 0; pass < 2; ++pass) {
                const float *src = pass == 0 ? t_mfcc : t_d1;
                float *dst = pass == 0 ? t_d1 : t_d2;
                for (int r = tid; r < C * S; r += NT) {
                    const int c = r / S, s = r - c * S;
                    float prv = 0.f, nxt = 0.f;
                    for (int k = 0; k < NC; ++k) {
                        const float *row = src + (c * NC + k) * S;
                        float nume = 0.f, d = 0.f;
                        for (int n = 1; n <= 2; ++n) {
                            const int sp = s - n < 0 ? 0 : s - n;
                            const int sn = s + n > S - 1 ? S - 1 : s + n;
                            prv += row[sp];
                            nxt += row[sn];
                            nume += (float)n * (nxt - prv);
                            d = nume / (float)(2 * n * n);
                        }
                        dst[(c * NC + k) * S + s] = d;
                    }
                }
                __syncthreads();
            }
        }

        // ---- coalesced stores of the finished tiles
        {
            const size_t seg = (size_t)ck.out_seg;
            if (P.o_mel) {
                float *dst = P.o_mel + seg * M * S;
                for (int i = tid; i < C * M * S; i += NT) dst[i] = t_mel[i];
            }
            if (P.o_energy) {
                float *dst = P.o_energy + seg * S;
                for (int i = tid; i < C * S; i += NT) dst[i] = t_energy[i];
            }
            if (P.do_mfcc) {
                if (P.o_mfcc) {
                    float *dst = P.o_mfcc + seg * NC * S;
                    for (int i = tid; i < C * NC * S; i += NT) dst[i] = t_mfcc[i];
                }
                if (P.do_deltas && P.o_d1) {
                    float *dst = P.o_d1 + seg * NC * S;
                    for (int i = tid; i < C * NC * S; i += NT) dst[i] = t_d1[i];
                }
                if (P.do_deltas && P.o_d2) {
                    float *dst = P.o_d2 + seg * NC * S;
                    for (int i = tid; i < C * NC * S; i += NT) dst[i] = t_d2[i];
                }
            }
            if (P.g_on && P.o_gabor) {
                float *dst = P.o_gabor + seg * P.g_len;
                for (int i = tid; i < C * P.g_len; i += NT) dst[i] = t_gab[i];
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------- power / log-power outputs
// Parity / inspection path only (PowerSegment, LogPowerSegment: dft/dft.go:62-85):
// rebuilds the per-segment smoothed power from the raw per-frame power the
// fused kernel left in `rawpow`.  One CTA per chunk.
struct PowParams {
    int step, stride, S, border, add, seg_adv;
    float prev, cur, log_off, log_min;
    int comp_log_pow, log1p_path;
    const Chunk *chunks;
    const float *rawpow;
    float *o_power, *o_logpower;
};

__global__ void power_segments_kernel(const __grid_constant__ PowParams Q) {
    const Chunk ck = Q.chunks[blockIdx.x];
    KParams P{};
    P.step = Q.step; P.stride = Q.stride; P.S = Q.S; P.border = Q.border; P.add = Q.add;
    for (int r = threadIdx.x; r < ck.nseg * kBins; r += blockDim.x) {
        const int c = r / kBins, k = r - c * kBins;
        const int nv = valid_steps(P, ck, ck.seg0 + c);
        const size_t base = ((size_t)(ck.out_seg + c) * kBins + k) * Q.S;
        float y = 0.f;
        for (int i = 0; i < Q.S; ++i) {
            float pw = 0.f, lp = 0.f;
            if (i < nv) {
                const float x = Q.rawpow[(size_t)(ck.frame_base + c * Q.seg_adv + i) * kPowPitch + k];
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
