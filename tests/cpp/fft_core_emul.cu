// fft_core_emul.cu -- runs one warp-round of the fused kernel's FFT core (aud_fft_core.cuh) lane by lane on the
// CPU: the same index arithmetic, exchange layout and operation order as the device code, checked against a
// float64 DFT.  Test infrastructure for the -m "not gpu" suite (tests/test_fft_core_emul.py); host code only.
//   nvcc -std=c++17 -I auditory_b200/csrc -o fft_core_emul tests/cpp/fft_core_emul.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "aud_fft_core.cuh"

using namespace aud;

int main(int argc, char **argv) {
    const int ps = 554;                       // pair stride of the default geometry (10 mod 16)
    const unsigned seed = argc > 1 ? (unsigned)atoi(argv[1]) : 1u;
    srand(seed);
    std::vector<float2> scr((size_t)kPairs * ps, make_float2(NAN, NAN));   // NaN: any read of an unwritten slot shows
    std::vector<float2> tw(200);
    for (int k1 = 0; k1 < 20; ++k1)
        for (int j = 0; j < 10; ++j) {
            const double a = -2.0 * M_PI * (double)(k1 * 2 * j) / 400.0;
            tw[k1 * 10 + j] = make_float2(0.5f * (float)cos(a), 0.5f * (float)sin(a));
        }
    float xa[kPairs][400], xb[kPairs][400];
    for (int q = 0; q < kPairs; ++q)
        for (int n = 0; n < 400; ++n) {
            xa[q][n] = (float)rand() / RAND_MAX * 2.f - 1.f;
            xb[q][n] = ((float)rand() / RAND_MAX * 2.f - 1.f) * (q == 1 ? 0.05f : 1.f);
        }
    // pass 1
    static f2 R[30][20], I[30][20];
    for (int lane = 0; lane < 30; ++lane) {
        const int q = lane / 10, j = lane % 10;
        for (int n1 = 0; n1 < 20; ++n1) {
            R[lane][n1] = make_float2(xa[q][20 * n1 + 2 * j], xa[q][20 * n1 + 2 * j + 1]);
            I[lane][n1] = make_float2(xb[q][20 * n1 + 2 * j], xb[q][20 * n1 + 2 * j + 1]);
        }
        dft20(R[lane], I[lane]);
        pass1_store(R[lane], I[lane], scr.data() + q * ps + exch_off(q), tw.data(), j);
    }
    // pass 2: every (pair, row pair) must be owned by exactly one lane
    int owned[kPairs][10] = {};
    for (int lane = 0; lane < 30; ++lane) {
        const int q2 = pass2_assign(lane) >> 5, p = pass2_assign(lane) & 31;
        if (q2 >= kPairs || p >= 10) { printf("FAIL: bad assignment for lane %d\n", lane); return 1; }
        ++owned[q2][p];
        pass2_load(R[lane], I[lane], scr.data() + q2 * ps + exch_off(q2), p);
    }
    for (int q = 0; q < kPairs; ++q)
        for (int p = 0; p < 10; ++p)
            if (owned[q][p] != 1) { printf("FAIL: row pair (%d,%d) owned %d times\n", q, p, owned[q][p]); return 1; }
    // the exchange rows are dead from here on: poison them, as the next window / the power buffer overwrite them
    for (auto &v : scr) v = make_float2(NAN, NAN);
    for (int lane = 0; lane < 30; ++lane) {
        const int q2 = pass2_assign(lane) >> 5, p = pass2_assign(lane) & 31;
        dft20(R[lane], I[lane]);
        if (p == 0) pass2_park(R[lane], I[lane], scr.data() + q2 * ps);
        else pass2_power(R[lane], I[lane], scr.data() + q2 * ps, p);
    }
    for (int item = 0; item < 21 * kPairs; ++item) selfpair_item(scr.data() + (item / 21) * ps, item % 21);
    // reference
    double worst = 0.0;
    for (int q = 0; q < kPairs; ++q)
        for (int k = 0; k <= 200; ++k) {
            double ar = 0, ai = 0, br = 0, bi = 0;
            for (int n = 0; n < 400; ++n) {
                const double ang = -2.0 * M_PI * (double)((n * k) % 400) / 400.0;
                ar += xa[q][n] * cos(ang); ai += xa[q][n] * sin(ang);
                br += xb[q][n] * cos(ang); bi += xb[q][n] * sin(ang);
            }
            const double pa = ar * ar + ai * ai, pb = br * br + bi * bi;
            const float2 got = scr[(size_t)q * ps + k + k / 20];
            // float32 FFT: absolute error relative to the frame pair's spectral scale
            double scale = 0;
            for (int n = 0; n < 400; ++n) scale += (double)xa[q][n] * xa[q][n] + (double)xb[q][n] * xb[q][n];
            const double ea = fabs(got.x - pa) / scale, eb = fabs(got.y - pb) / scale;
            if (!(ea < 1e-5) || !(eb < 1e-5)) {
                printf("FAIL: pair %d bin %d: got (%g, %g) want (%g, %g)\n", q, k, got.x, got.y, pa, pb);
                return 1;
            }
            worst = fmax(worst, fmax(ea, eb));
        }
    printf("OK worst %.3e\n", worst);
    return 0;
}
