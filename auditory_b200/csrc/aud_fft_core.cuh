// aud_fft_core.cuh -- the arithmetic core of the fused kernel's FFT warps: two real frames ride one complex
// 400-point FFT, factored 20 x 20 with in-register prime-factor DFT-20s, and the pair (Z[k], Z[N-k]) is split
// into |X_A[k]|^2, |X_B[k]|^2.  Reference semantics: dft/dft.go:42-85 (rectangular window, length-WinSamples
// forward DFT, power = re^2 + im^2).
//
// sm_100a: the butterflies run on packed FP32 pairs (FADD2 / FMUL2 / FFMA2, `add/mul/fma.rn.f32x2`): a lane's
// two independent DFT-20s -- two columns in pass 1, the row pair (u, 20 - u) in pass 2 -- share every
// instruction, which halves the issue slots of the transform.  The exchange buffer between the passes is laid
// out so that both sides move whole packed pairs with 128-bit accesses and no register shuffling:
//
//   pass 1  lane j of a frame pair owns columns c0 = 2j, c1 = 2j+1:  R[n1] = (x_A[20 n1 + c0], x_A[20 n1 + c1])
//           (one 8-byte window load), I[n1] likewise from frame B; DFT-20 over n1; twiddle by W400^{c k1};
//           for each row pair p = (u, v) -- (0, 10), (1, 19), (2, 18) ... (9, 11) -- one float4 per column:
//           E[p][epos(c)] = (Re Z1[u][c], Re Z1[v][c], Im Z1[u][c], Im Z1[v][c])
//   pass 2  the lane that owns row pair p loads E[p][*]: R[c] = (Re Z1[u][c], Re Z1[v][c]), I[c] = (Im, Im)
//           straight from the float4; DFT-20 over c; Z[u + 20 m] and Z[v + 20 m] are its outputs.
//
// Everything here is __host__ __device__ so that tests/cpp/fft_core_emul.cu can run one warp-round lane by lane
// on the CPU (same index arithmetic, same operation order) against a float64 DFT.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define AUD_HD __host__ __device__ __forceinline__

namespace aud {

constexpr int kN = 400;             // FFT length the fused kernel is specialised for
constexpr int kBins = kN / 2 + 1;   // 201
constexpr int kPairs = 3;           // frame pairs per warp and round (10 lanes each, lanes 30/31 idle in the FFT)
constexpr int kEPitch = 42;         // exchange buffer: float2 units per row pair (21 float4: 20 columns + 1 pad, odd so
                                    // that the eight lanes of a quarter-warp hit eight different 16-byte bank groups)
constexpr int kExchange = 10 + 10 * kEPitch;   // float2 units the exchange rows of a pair may span (largest pair offset + rows)
constexpr int kWinOff = 264;        // float2 offset of the next round's sample window inside a pair's scratch
                                    // (above the power buffer [0,219) and the parked rows [220,260))
constexpr int kPPitch = 21;         // padded natural order of the power buffer: index(k) = k + k/20
constexpr int kZPark = 220;         // where the lane of row pair (0, 10) parks its outputs (float2 index, 20 float4)

// ------------------------------------------------------------------ packed FP32 pairs
typedef float2 f2;
AUD_HD f2 add2(f2 a, f2 b) {
#ifdef __CUDA_ARCH__
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
AUD_HD f2 sub2(f2 a, f2 b) {
#ifdef __CUDA_ARCH__
    return __fadd2_rn(a, make_float2(-b.x, -b.y));   // SASS: FADD2 with a negated source, no extra instruction
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
AUD_HD f2 mul2(f2 a, f2 b) {
#ifdef __CUDA_ARCH__
    return __fmul2_rn(a, b);
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}
AUD_HD f2 fma2(f2 a, f2 b, f2 c) {
#ifdef __CUDA_ARCH__
    return __ffma2_rn(a, b, c);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
AUD_HD f2 bc2(float c) { return make_float2(c, c); }

// ------------------------------------------------------------------ DFT-20
// Prime-factor (Good-Thomas) 4 x 5 DFT on 20 complex values held in registers, two transforms at once (the .x
// and .y halves).  Input natural order; after the call the value for output index k sits in slot perm20(k).
__host__ __device__ constexpr int perm20(int k) { return (5 * (k % 4) + 4 * (k % 5)) % 20; }

AUD_HD void dft4(f2 &r0, f2 &i0, f2 &r1, f2 &i1, f2 &r2, f2 &i2, f2 &r3, f2 &i3) {
    const f2 ar = add2(r0, r2), ai = add2(i0, i2), br = sub2(r0, r2), bi = sub2(i0, i2);
    const f2 cr = add2(r1, r3), ci = add2(i1, i3), dr = sub2(r1, r3), di = sub2(i1, i3);
    r0 = add2(ar, cr); i0 = add2(ai, ci);
    r1 = add2(br, di); i1 = sub2(bi, dr);      // b - i d
    r2 = sub2(ar, cr); i2 = sub2(ai, ci);
    r3 = sub2(br, di); i3 = add2(bi, dr);      // b + i d
}

AUD_HD void dft5(f2 &r0, f2 &i0, f2 &r1, f2 &i1, f2 &r2, f2 &i2, f2 &r3, f2 &i3, f2 &r4, f2 &i4) {
    constexpr float C1 = 0.30901699437494742410f;    // cos(2 pi / 5)
    constexpr float C2 = -0.80901699437494742410f;   // cos(4 pi / 5)
    constexpr float S1 = 0.95105651629515357212f;    // sin(2 pi / 5)
    constexpr float S2 = 0.58778525229247312917f;    // sin(4 pi / 5)
    const f2 t1r = add2(r1, r4), t1i = add2(i1, i4), t2r = add2(r2, r3), t2i = add2(i2, i3);
    const f2 t3r = sub2(r1, r4), t3i = sub2(i1, i4), t4r = sub2(r2, r3), t4i = sub2(i2, i3);
    const f2 m1r = fma2(bc2(C2), t2r, fma2(bc2(C1), t1r, r0)), m1i = fma2(bc2(C2), t2i, fma2(bc2(C1), t1i, i0));
    const f2 m2r = fma2(bc2(C1), t2r, fma2(bc2(C2), t1r, r0)), m2i = fma2(bc2(C1), t2i, fma2(bc2(C2), t1i, i0));
    const f2 s1r = fma2(bc2(S2), t4r, mul2(bc2(S1), t3r)), s1i = fma2(bc2(S2), t4i, mul2(bc2(S1), t3i));
    const f2 s2r = fma2(bc2(-S1), t4r, mul2(bc2(S2), t3r)), s2i = fma2(bc2(-S1), t4i, mul2(bc2(S2), t3i));
    r0 = add2(add2(r0, t1r), t2r); i0 = add2(add2(i0, t1i), t2i);
    r1 = add2(m1r, s1i); i1 = sub2(m1i, s1r);   // m1 - i s1
    r4 = sub2(m1r, s1i); i4 = add2(m1i, s1r);   // m1 + i s1
    r2 = add2(m2r, s2i); i2 = sub2(m2i, s2r);   // m2 - i s2
    r3 = sub2(m2r, s2i); i3 = add2(m2i, s2r);   // m2 + i s2
}

AUD_HD void dft20(f2 (&xr)[20], f2 (&xi)[20]) {
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

// W400^k = exp(-2 pi i k / 400), k = 0..19 (function-local constants: they fold into FFMA / FMUL immediates)
AUD_HD float w400r(int k) {
    constexpr float t[20] = {1.f, 0.999876632f, 0.99950656f, 0.998889875f, 0.998026728f, 0.996917334f, 0.995561965f,
                             0.993960955f, 0.992114701f, 0.990023658f, 0.987688341f, 0.985109326f, 0.982287251f,
                             0.979222811f, 0.975916762f, 0.97236992f, 0.968583161f, 0.964557418f, 0.960293686f,
                             0.955793015f};
    return t[k];
}
AUD_HD float w400i(int k) {
    constexpr float t[20] = {-0.f, -0.0157073173f, -0.0314107591f, -0.0471064507f, -0.0627905195f, -0.0784590957f,
                             -0.0941083133f, -0.109734311f, -0.125333234f, -0.140901232f, -0.156434465f, -0.1719291f,
                             -0.187381315f, -0.202787295f, -0.218143241f, -0.233445364f, -0.248689887f, -0.26387305f,
                             -0.278991106f, -0.294040325f};
    return t[k];
}

// ------------------------------------------------------------------ exchange layout
// rows of row pair p: (0, 10) for p = 0, else (p, 20 - p)
__host__ __device__ constexpr int row_u(int p) { return p; }
__host__ __device__ constexpr int row_v(int p) { return p == 0 ? 10 : 20 - p; }
// float4 slot of column c inside a row pair: even columns first, so that the ten lanes of a frame pair write ten
// consecutive 16-byte chunks (conflict-free 128-bit stores) for c0 = 2j and again for c1 = 2j + 1
__host__ __device__ constexpr int epos(int c) { return (c & 1) * 10 + (c >> 1); }
// float2 offset of a pair's exchange rows inside its scratch: with a pair stride of 10 (mod 16) float2 the three
// pairs' rows then start 0 / 2 / 4 (mod 8) 16-byte chunks apart, which is what makes a quarter-warp that straddles
// two pairs conflict-free in both passes
__host__ __device__ constexpr int exch_off(int q) { return q == 1 ? 10 : q == 2 ? 4 : 0; }

// Pass-2 work assignment, lane -> (pair q2, row pair p), packed q2 << 5 | p.  Found by search (tools/pass2_assign_search.py)
// for a pair stride of 10 (mod 16) float2: every quarter-warp's row loads hit eight different bank groups, both halves
// of every half-warp's power stores hit sixteen different 8-byte banks, and the three p = 0 lanes (self-paired rows,
// parked for the cooperative step) share one quarter-warp.
#define AUD_PASS2_TABLE                                                                            \
    {0 << 5 | 0, 1 << 5 | 0, 2 << 5 | 0, 2 << 5 | 2, 1 << 5 | 1, 0 << 5 | 5, 1 << 5 | 5, 2 << 5 | 5,   \
     1 << 5 | 4, 2 << 5 | 3, 0 << 5 | 3, 1 << 5 | 7, 0 << 5 | 2, 0 << 5 | 8, 1 << 5 | 3, 2 << 5 | 8,   \
     2 << 5 | 1, 1 << 5 | 8, 1 << 5 | 2, 1 << 5 | 6, 0 << 5 | 1, 0 << 5 | 6, 0 << 5 | 7, 1 << 5 | 9,   \
     0 << 5 | 9, 2 << 5 | 7, 2 << 5 | 6, 2 << 5 | 9, 2 << 5 | 4, 0 << 5 | 4, 2 << 5 | 0, 2 << 5 | 0}
__device__ constexpr unsigned char kPass2Dev[32] = AUD_PASS2_TABLE;   // constant bank on the device
AUD_HD int pass2_assign(int lane) {
#ifdef __CUDA_ARCH__
    return kPass2Dev[lane];
#else
    constexpr unsigned char t[32] = AUD_PASS2_TABLE;
    return t[lane];
#endif
}

// ------------------------------------------------------------------ pass 1
// Z1[k1][c] = W400^{c k1} * DFT20_{n1}(z[20 n1 + c]) for the lane's columns c0 = 2j (.x halves) and c1 = 2j + 1
// (.y halves), written as packed row pairs.  tw[10 k1 + j] = (1/2) W400^{2 j k1}: the 1/2 of the real-pair split
// rides on the twiddles (exactly), so that |X|^2 = |Z[k] +- conj Z[N-k]|^2 needs no 1/4.
AUD_HD void twiddle_row(const f2 &yr, const f2 &yi, int k1, float2 w, float &re0, float &im0, float &re1, float &im1) {
    if (k1 == 0) {
        re0 = 0.5f * yr.x; im0 = 0.5f * yi.x; re1 = 0.5f * yr.y; im1 = 0.5f * yi.y;
        return;
    }
    const float vr = w.x * w400r(k1) - w.y * w400i(k1);        // (1/2) W400^{(2j+1) k1}
    const float vi = fmaf(w.x, w400i(k1), w.y * w400r(k1));
    re0 = yr.x * w.x - yi.x * w.y; im0 = fmaf(yr.x, w.y, yi.x * w.x);
    re1 = yr.y * vr - yi.y * vi;   im1 = fmaf(yr.y, vi, yi.y * vr);
}

AUD_HD void pass1_store(const f2 (&R)[20], const f2 (&I)[20], float2 *exq, const float2 *tw, int j) {
    float4 *e4 = reinterpret_cast<float4 *>(exq);
    // twiddles are fetched two row pairs (four rows) ahead of their use, so that the table loads' latency hides
    // behind the previous rows' arithmetic
    float2 wu[10], wv[10];
#pragma unroll
    for (int p = 0; p < 2; ++p) { wu[p] = tw[10 * row_u(p) + j]; wv[p] = tw[10 * row_v(p) + j]; }
#pragma unroll
    for (int p = 0; p < 10; ++p) {
        if (p + 2 < 10) { wu[p + 2] = tw[10 * row_u(p + 2) + j]; wv[p + 2] = tw[10 * row_v(p + 2) + j]; }
        const int u = row_u(p), v = row_v(p);
        float ru0, iu0, ru1, iu1, rv0, iv0, rv1, iv1;
        twiddle_row(R[perm20(u)], I[perm20(u)], u, wu[p], ru0, iu0, ru1, iu1);
        twiddle_row(R[perm20(v)], I[perm20(v)], v, wv[p], rv0, iv0, rv1, iv1);
        e4[p * (kEPitch / 2) + j] = make_float4(ru0, rv0, iu0, iv0);        // column c0: epos(2j) = j
        e4[p * (kEPitch / 2) + 10 + j] = make_float4(ru1, rv1, iu1, iv1);   // column c1: epos(2j+1) = 10 + j
    }
}

// ------------------------------------------------------------------ pass 2
AUD_HD void pass2_load(f2 (&R)[20], f2 (&I)[20], const float2 *ex2, int p) {
    const float4 *row = reinterpret_cast<const float4 *>(ex2) + p * (kEPitch / 2);
#pragma unroll
    for (int c = 0; c < 20; ++c) {
        const float4 v = row[epos(c)];
        R[c] = make_float2(v.x, v.y);
        I[c] = make_float2(v.z, v.w);
    }
}

// After dft20(R, I): A[m] = Z[u + 20 m] in the .x halves, B[m] = Z[(20 - u) + 20 m] in the .y halves (slot perm20(m)).
// |X_A|^2, |X_B|^2 of bin k from the pair (Z[k], Z[N-k]) = (A[m], B[19-m]); the formulas are symmetric in the pair, so
// m >= 10 yields the bins of the mirror column.  Stored in padded natural order P[k + k/20] = (A, B).  u = 1..9.
AUD_HD void pass2_power(const f2 (&R)[20], const f2 (&I)[20], float2 *pq, int u) {
#pragma unroll
    for (int m = 0; m < 20; ++m) {
        const float zr = R[perm20(m)].x, zi = I[perm20(m)].x;
        const float wr = R[perm20(19 - m)].y, wi = I[perm20(19 - m)].y;
        const f2 U = make_float2(zr + wr, zi + wi);   // (Re 2X_A, Re' 2X_B)
        const f2 V = make_float2(zi - wi, wr - zr);   // (Im 2X_A, Im' 2X_B)
        const int idx = (m < 10) ? u + kPPitch * m : (20 - u) + kPPitch * (19 - m);
        pq[idx] = fma2(U, U, mul2(V, V));
    }
}
// Row pair (0, 10): both rows pair with themselves (Z[20 m] with Z[400 - 20 m], Z[10 + 20 m] with Z[390 - 20 m]);
// the lane parks its outputs, park[m] = (Re A[m], Re B[m], Im A[m], Im B[m]), for the cooperative step below.
AUD_HD void pass2_park(const f2 (&R)[20], const f2 (&I)[20], float2 *pq) {
    float4 *pk = reinterpret_cast<float4 *>(pq + kZPark);
#pragma unroll
    for (int m = 0; m < 20; ++m) pk[m] = make_float4(R[perm20(m)].x, R[perm20(m)].y, I[perm20(m)].x, I[perm20(m)].y);
}
// One of the 21 self-paired bins of a frame pair: n <= 10 -> bin 20 n, n >= 11 -> bin 10 + 20 (n - 11).
AUD_HD void selfpair_item(float2 *pq, int n) {
    const float4 *pk = reinterpret_cast<const float4 *>(pq + kZPark);
    const bool row10 = n > 10;
    const int ia = row10 ? n - 11 : n;
    const int ib = row10 ? 30 - n : (n ? 20 - n : 0);
    const int idx = row10 ? 10 + kPPitch * (n - 11) : kPPitch * n;
    const float4 a = pk[ia], b = pk[ib];
    const float ar = row10 ? a.y : a.x, ai = row10 ? a.w : a.z, br = row10 ? b.y : b.x, bi = row10 ? b.w : b.z;
    const float xr = ar + br, xi = ai - bi, yr = ai + bi, yi = br - ar;
    pq[idx] = make_float2(fmaf(xr, xr, xi * xi), fmaf(yr, yr, yi * yi));
}

// ------------------------------------------------------------------ frame levels
// max |x| over a lane's samples of one frame, NaN-propagating (FMNMX3.NAN with |.| source modifiers): 0 <=> every
// sample is +-0; NaN / Inf <=> the frame holds a non-finite sample.  Returned as the bit pattern: non-negative
// floats and the canonical NaN order like unsigned integers, so lanes combine with an integer max.
AUD_HD float max3_nan(float m, float a, float b) {
#ifdef __CUDA_ARCH__
    float r;
    asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(m), "f"(fabsf(a)), "f"(fabsf(b)));
    return r;
#else
    if (m != m || a != a || b != b) return NAN;
    return fmaxf(m, fmaxf(fabsf(a), fabsf(b)));
#endif
}
AUD_HD float frame_peak(const f2 (&X)[20]) {
    float m0 = 0.f, m1 = 0.f;
#pragma unroll
    for (int n = 0; n < 20; n += 2) {
        m0 = max3_nan(m0, X[n].x, X[n].y);
        m1 = max3_nan(m1, X[n + 1].x, X[n + 1].y);
    }
    return max3_nan(m0, m1, 0.f);
}

// A frame rides alone (zero partner) when it is this much quieter than the frame it would share the transform
// with: the partner's float32 rounding noise (~1e-7 of ITS magnitude, in every bin) would otherwise show in the
// quiet frame's log-mel.  30 dB in peak level keeps that below 1e-5 relative.
constexpr float kAloneRatio = 32.f;

}  // namespace aud
