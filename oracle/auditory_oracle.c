/*
 * auditory_oracle.c -- float64 C twin of oracle/np_oracle.py: a CPU
 * restatement of emer/auditory's speech-feature path (sound.SndEnv ->
 * dft.Filter -> mel.FilterDft [-> mel.CepstrumDct] -> agabor.Convolve).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library, and
 * only as the checker or the timed CPU baseline.  The product path
 * (auditory_b200/) never links or calls it.
 *
 * PARITY UNPINNED.  The Go reference ships no tests or golden vectors for this
 * path and cannot be built here (no Go toolchain; gonum / etable are not
 * vendored).  See the header of np_oracle.py for what stands in.
 *
 * Third-party arithmetic restated here (gonum.org/v1/gonum v0.11.0,
 * go.mod:20): dsp/fourier.CmplxFFT.Coefficients = forward unnormalised DFT
 * (FFTPACK cfftf: mixed-radix Cooley-Tukey over the factors of n, 4 and 5
 * preferred), dsp/fourier.DCT.Transform = FFTPACK cost = unnormalised DCT-I.
 * The FFT below is an independent mixed-radix decimation-in-time
 * implementation of the same published algorithm, not a copy of FFTPACK.
 *
 * All file:line citations are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ params */
typedef struct orc_params {
    int32_t sample_rate;
    double win_ms, step_ms, segment_ms, stride_ms;
    int32_t border_steps;
    /* dft.Params (dft/dft.go:15-31) */
    int32_t comp_log_pow;
    double log_min, log_offset, prev_smooth, cur_smooth;
    /* mel.Params / FilterBank (mel/mel.go:16-66) */
    int32_t n_filters;
    double lo_hz, hi_hz, mel_log_off, mel_log_min;
    int32_t renorm;
    double renorm_min, renorm_max;
    int32_t mfcc, deltas, n_coefs;
    /* agabor.FilterSet (agabor/gabor.go:45-70) + SndEnv output shape */
    int32_t size_x, size_y, stride_x, stride_y;
    double gain;
    int32_t distribute;
    int32_t pools_y, pools_x, units_y, units_x;
    int32_t by_time;
    /* 1 = rebuild the FFT plan (factorisation + twiddles) for every frame as
     * dft/dft.go:45 does (fourier.NewCmplxFFT per call); 0 = build once. */
    int32_t rebuild_plan;
} orc_params;

typedef struct orc_gabor_spec { /* agabor/gabor.go:17-42 */
    int32_t off;
    double wave_len, orientation, sigma_width, sigma_length, phase_offset;
    int32_t circle_edge, circular;
} orc_gabor_spec;

/* sound/sndenv.go:522-524 MSecToSamples; Go math.Round = half away from zero */
ORC_API int32_t orc_msec_to_samples(double ms, int32_t rate) {
    return (int32_t)round(ms * 0.001 * (double)rate);
}

/* --------------------------------------------------------------------- FFT */
typedef struct orc_fft_plan {
    int n;
    int nfac;
    int fac[32];
    double *wr, *wi;   /* n roots of unity e^{-2 pi i k/n} */
    double *sr, *si;   /* scratch for generic-radix butterflies */
    double *tr, *ti;
} orc_fft_plan;

static void plan_init(orc_fft_plan *p, int n) {
    p->n = n;
    p->nfac = 0;
    int m = n;
    /* factor order preference mirrors FFTPACK's trial list 4,2,3,5,7,... */
    static const int tryf[4] = {4, 2, 3, 5};
    for (int t = 0; t < 4; t++)
        while (m % tryf[t] == 0 && m > 1) { p->fac[p->nfac++] = tryf[t]; m /= tryf[t]; }
    for (int f = 7; m > 1; f += 2)
        while (m % f == 0) { p->fac[p->nfac++] = f; m /= f; }
    p->wr = (double *)malloc(sizeof(double) * (size_t)n * 6);
    p->wi = p->wr + n;
    p->sr = p->wi + n;
    p->si = p->sr + n;
    p->tr = p->si + n;
    p->ti = p->tr + n;
    for (int k = 0; k < n; k++) {
        double a = -2.0 * M_PI * (double)k / (double)n;
        p->wr[k] = cos(a);
        p->wi[k] = sin(a);
    }
}
static void plan_free(orc_fft_plan *p) { free(p->wr); p->wr = NULL; }

/* recursive decimation in time: out[0..n) = DFT(in[0], in[stride], ...) */
static void fft_rec(const orc_fft_plan *P, const double *ir, const double *ii, int stride,
                    double *or_, double *oi, int n, int level) {
    if (n == 1) { or_[0] = ir[0]; oi[0] = ii[0]; return; }
    const int p = P->fac[level];
    const int m = n / p;
    for (int r = 0; r < p; r++)
        fft_rec(P, ir + (size_t)r * stride, ii + (size_t)r * stride, stride * p, or_ + r * m, oi + r * m, m, level + 1);
    const int tw = P->n / n;   /* twiddle step in the n-root table */
    double *tr = P->tr, *ti = P->ti;
    for (int k = 0; k < m; k++) {
        for (int r = 0; r < p; r++) {
            int widx = (int)(((long)r * k * tw) % P->n);
            double a = or_[r * m + k], b = oi[r * m + k];
            tr[r] = a * P->wr[widx] - b * P->wi[widx];
            ti[r] = a * P->wi[widx] + b * P->wr[widx];
        }
        if (p == 2) {
            or_[k] = tr[0] + tr[1];       oi[k] = ti[0] + ti[1];
            or_[k + m] = tr[0] - tr[1];   oi[k + m] = ti[0] - ti[1];
        } else if (p == 4) {
            double ar = tr[0] + tr[2], ai = ti[0] + ti[2];
            double br = tr[0] - tr[2], bi = ti[0] - ti[2];
            double cr = tr[1] + tr[3], ci = ti[1] + ti[3];
            double dr = tr[1] - tr[3], di = ti[1] - ti[3];
            or_[k] = ar + cr;             oi[k] = ai + ci;
            or_[k + m] = br + di;         oi[k + m] = bi - dr;      /* -i*d */
            or_[k + 2 * m] = ar - cr;     oi[k + 2 * m] = ai - ci;
            or_[k + 3 * m] = br - di;     oi[k + 3 * m] = bi + dr;
        } else {
            const int step = P->n / p;
            for (int q = 0; q < p; q++) {
                double sr = 0.0, si = 0.0;
                for (int r = 0; r < p; r++) {
                    int widx = (int)(((long)r * q * step) % P->n);
                    sr += tr[r] * P->wr[widx] - ti[r] * P->wi[widx];
                    si += tr[r] * P->wi[widx] + ti[r] * P->wr[widx];
                }
                P->sr[q] = sr; P->si[q] = si;
            }
            for (int q = 0; q < p; q++) { or_[k + q * m] = P->sr[q]; oi[k + q * m] = P->si[q]; }
        }
    }
}

ORC_API void orc_fft(const double *in_re, const double *in_im, int32_t n, double *out_re, double *out_im) {
    orc_fft_plan P;
    plan_init(&P, n);
    fft_rec(&P, in_re, in_im, 1, out_re, out_im, n, 0);
    plan_free(&P);
}

/* gonum fourier.DCT.Transform == FFTPACK cost (unnormalised DCT-I) */
ORC_API void orc_dct1(const double *x, int32_t n, double *y) {
    for (int k = 0; k < n; k++) {
        double s = x[0] + ((k & 1) ? -x[n - 1] : x[n - 1]);
        for (int j = 1; j < n - 1; j++) s += 2.0 * x[j] * cos(M_PI * (double)j * (double)k / (double)(n - 1));
        y[k] = s;
    }
}

/* --------------------------------------------------------------------- mel */
static double freq_to_mel(double f) { return 1127.0 * log(1.0 + f / 700.0); }       /* mel/mel.go:156 */
static double mel_to_freq(double m) { return 700.0 * (exp(m / 1127.0) - 1.0); }     /* mel/mel.go:161 */
static int freq_to_bin(double f, double nfft, double sr) { return (int)floor(((nfft + 1) * f) / sr); } /* :166 */

/* mel/mel.go:77-117 InitFilters.  filters is [n_filters][n_filters+2] flat;
 * returns -1 where the reference would panic (flat offset past the end). */
ORC_API int32_t orc_mel_init(const orc_params *p, int32_t dft_size, int32_t *binpts, double *hzpts, double *filters) {
    const int nf = p->n_filters, maxb = nf + 2;
    const double hi = freq_to_mel(p->hi_hz), lo = freq_to_mel(p->lo_hz);
    const double incr = (hi - lo) / (double)(nf + 1);
    for (int i = 0; i < nf + 2; i++) {
        double hz = mel_to_freq(lo + (double)i * incr);
        if (hzpts) hzpts[i] = hz;
        binpts[i] = freq_to_bin(hz, (double)dft_size, (double)p->sample_rate);
    }
    memset(filters, 0, sizeof(double) * (size_t)nf * maxb);
    for (int f = 0; f < nf; f++) {
        const int bmin = binpts[f], bctr = binpts[f + 1], bmax = binpts[f + 2];
        const double pkmin = (double)bctr - (double)bmin, pkmax = (double)bmax - (double)bctr;
        int fi = 0, b;
        for (b = bmin; b <= bctr; b++, fi++) {
            long off = (long)f * maxb + fi;
            if (off >= (long)nf * maxb) return -1;
            filters[off] = ((double)b - (double)bmin) / pkmin;
        }
        for (; b <= bmax; b++, fi++) {
            long off = (long)f * maxb + fi;
            if (off >= (long)nf * maxb) return -1;
            filters[off] = ((double)bmax - (double)b) / pkmax;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------- gabor */
/* agabor/gabor.go:89-222 ToTensor (with :73-86 Defaults, :329-336 Active).
 * filters must hold n_active*size_y*size_x doubles; returns n_active. */
ORC_API int32_t orc_gabor_to_tensor(const orc_params *p, const orc_gabor_spec *specs, int32_t nspecs, double *filters) {
    int nact = 0, nhf = 0, nvf = 0;
    for (int i = 0; i < nspecs; i++) if (!specs[i].off) nact++;
    if (p->distribute) {
        for (int i = 0; i < nspecs; i++) {
            if (specs[i].off) continue;
            if (specs[i].orientation == 0) nhf++;
            else if (specs[i].orientation == 90) nvf++;
        }
    } else { nhf = 1; nvf = 1; }
    const int sx = p->size_x, sy = p->size_y;
    const double radx = (double)sx / 2.0, rady = (double)sy / 2.0;
    const double ctrx = (double)(sx - 1) / 2.0, ctry = (double)(sy - 1) / 2.0;
    const double hinc = (double)(sy - 1) / (double)(nhf + 1), vinc = (double)(sx - 1) / (double)(nvf + 1);
    int hcnt = 0, vcnt = 0, ai = 0;
    for (int si = 0; si < nspecs; si++) {
        if (specs[si].off) continue;
        orc_gabor_spec f = specs[si];
        if (f.wave_len == 0) f.wave_len = 2;
        if (f.sigma_length == 0 && !f.circular) f.sigma_length = 0.5;
        if (f.sigma_width == 0) f.sigma_width = 0.5;
        const double twopin = (2.0 * M_PI) / f.wave_len;
        const double lnorm = 1.0 / (2.0 * f.sigma_length * f.sigma_length);
        const double wnorm = 1.0 / (2.0 * f.sigma_width * f.sigma_width);
        double hpos = 0, vpos = 0;
        if (p->distribute) {
            if (f.orientation == 0) { hpos = hinc * (double)(hcnt + 1); hcnt++; }
            if (f.orientation == 90) { vpos = vinc * (double)(vcnt + 1); vcnt++; }
        } else {
            hpos = hinc * (double)(hcnt + 1);
            vpos = vinc * (double)(vcnt + 1);
        }
        double *out = filters + (size_t)ai * sy * sx;
        if (!f.circular) {
            for (int y = 0; y < sy; y++)
                for (int x = 0; x < sx; x++) {
                    double xf = (double)x - ctrx, yf = (double)y - ctry;
                    if (f.orientation == 0) yf = (double)y - hpos;
                    if (f.orientation == 90) xf = (double)x - vpos;
                    double xfn = xf / radx, yfn = yf / rady;
                    double dist = hypot(xfn, yfn), val = 0;
                    if (!(f.circle_edge && dist > 1.0)) {
                        double rad = f.orientation * M_PI / 180;
                        double nx = xfn * cos(rad) - yfn * sin(rad);
                        double ny = yfn * cos(rad) + xfn * sin(rad);
                        double g = exp(-(wnorm * (nx * nx) + lnorm * (ny * ny)));
                        val = g * sin(twopin * ny + f.phase_offset);
                    }
                    out[y * sx + x] = val;
                }
        } else {
            const double norm = 1.0 / (2.0 * f.sigma_width * f.sigma_width);
            for (int y = 0; y < sy; y++)
                for (int x = 0; x < sx; x++) {
                    double xfn = ((double)x - ctrx) / radx, yfn = ((double)y - ctry) / rady;
                    double nx = xfn * xfn * norm, ny = yfn * yfn * norm;
                    out[y * sx + x] = -sqrt(nx + ny) * sin(twopin * nx * ny);
                }
        }
        ai++;
    }
    for (int i = 0; i < nact; i++) {     /* renorm each half: gabor.go:194-221 */
        double *f = filters + (size_t)i * sy * sx, ps = 0, ns = 0;
        for (int j = 0; j < sy * sx; j++) { if (f[j] > 0) ps += f[j]; else if (f[j] < 0) ns += f[j]; }
        const double pn = 1.0 / ps, nn = -1.0 / ns;
        for (int j = 0; j < sy * sx; j++) { if (f[j] > 0.0) f[j] *= pn; else if (f[j] < 0.0) f[j] *= nn; }
    }
    return nact;
}

/* agabor/gabor.go:225-315 Convolve.  mel is [M][S] row-major float64; out is
 * float32 with `ndims` (2 or 4) dims of `shape`, written through flat stride
 * arithmetic.  Returns 0 ok, 1 = reference logs and returns without writing,
 * -1 = reference would panic (index out of range). */
ORC_API int32_t orc_gabor_convolve(const double *mel, int32_t M, int32_t S, const double *filters, int32_t nf,
                                   int32_t sy, int32_t sx, int32_t stride_y, int32_t stride_x, double gain,
                                   float *out, int32_t ndims, const int32_t *shape, int32_t by_time) {
    if (S < sx) return 1;
    int tmax = 1, fmax = 1, tmaxstrides = 1;
    long ostr[4] = {0, 0, 0, 0}, olen = 1;
    if (ndims == 2) {
        int x = S - sx;
        if (!(x == 0 || x < stride_x)) tmax = x + 1;
        tmaxstrides = (S - sx) / stride_x + 1;
        int y = M - sy;
        if (!(y == 0 || y < stride_y)) fmax = y + 1;
    } else if (ndims == 4) {
        tmax = (int)fmin((double)(shape[1] * stride_x), (double)(S - stride_x));
        fmax = (int)fmin((double)(shape[0] * stride_y), (double)(M - stride_y));
    } else return 1;
    for (int d = ndims - 1; d >= 0; d--) { ostr[d] = olen; olen *= shape[d]; }
#define ORC_PUT(off_, v_) do { long o_ = (off_); if (o_ < 0 || o_ >= olen) return -1; out[o_] = (float)(v_); } while (0)
    int tidx = 0;
    for (int t = 0; t < tmax; t += stride_x, tidx++) {
        int fidx = 0;
        for (int f = 0; f < fmax; f += stride_y, fidx++) {
            for (int flt = 0; flt < nf; flt++) {
                double fsum = 0.0;
                for (int ff = 0; ff < sy; ff++)
                    for (int ft = 0; ft < sx; ft++) {
                        long moff = (long)(f + ff) * S + (t + ft);
                        if (moff >= (long)M * S) return -1;
                        double iv = mel[moff];
                        if (isnan(iv)) iv = .5;
                        fsum += filters[((size_t)flt * sy + ff) * sx + ft] * iv;
                    }
                const int pos = fsum >= 0.0;
                const double act = gain * fabs(fsum);
                if (ndims == 2) {
                    long y = (long)fidx * 2;
                    long x = by_time ? (tidx + (long)tmaxstrides * flt) : (flt + (long)tidx * nf);
                    ORC_PUT(y * ostr[0] + x, pos ? act : 0.0);
                    ORC_PUT((y + 1) * ostr[0] + x, pos ? 0.0 : act);
                } else {
                    long base = fidx * ostr[0] + tidx * ostr[1] + flt;
                    ORC_PUT(base, pos ? act : 0.0);
                    ORC_PUT(base + ostr[2], pos ? 0.0 : act);
                }
            }
        }
    }
#undef ORC_PUT
    return 0;
}

/* ------------------------------------------------------------------ SndEnv */
typedef struct orc_env {
    orc_params p;
    int win, step, seg_samples, stride, S, B, nf, ncoef;
    int32_t *binpts;
    double *mel_filters;            /* [nf][nf+2] */
    int gabor_nf;
    double *gabor;                  /* [gabor_nf][sy][sx] */
    int gabor_dims;                 /* 0 (none), 2 or 4 */
    int32_t gshape[4];
    long gabor_len;
    /* per-env scratch (one env per thread, like one SndEnv per goroutine) */
    double *window, *cre, *cim, *zero, *power, *logpower;
    double *power_seg, *logpower_seg, *mel_fbank, *mel_seg, *energy, *mfcc_seg, *deltas, *ddeltas, *dct_tmp;
    float *gabor_out;
    orc_fft_plan plan;
    int have_plan;
} orc_env;

ORC_API void orc_env_destroy(orc_env *e) {
    if (!e) return;
    free(e->binpts); free(e->mel_filters); free(e->gabor); free(e->window); free(e->power_seg);
    free(e->gabor_out);
    if (e->have_plan) plan_free(&e->plan);
    free(e);
}

/* sound/sndenv.go:195-267 Init.  Returns 0, or <0: -1 sample rate, -2 mel
 * table panic (F2), -3 bad gabor pools spec, -4 SegmentSteps > bins (F6). */
ORC_API int32_t orc_env_create(const orc_params *p, const orc_gabor_spec *specs, int32_t nspecs, orc_env **out) {
    *out = NULL;
    if (p->sample_rate <= 0) return -1;
    orc_env *e = (orc_env *)calloc(1, sizeof(orc_env));
    e->p = *p;
    e->win = orc_msec_to_samples(p->win_ms, p->sample_rate);
    e->step = orc_msec_to_samples(p->step_ms, p->sample_rate);
    e->seg_samples = orc_msec_to_samples(p->segment_ms, p->sample_rate);
    e->S = (int)round(p->segment_ms / p->step_ms) + 2 * p->border_steps;
    e->stride = orc_msec_to_samples(p->stride_ms, p->sample_rate);
    e->B = e->win / 2 + 1;
    e->nf = p->n_filters;
    e->ncoef = p->n_coefs;
    e->binpts = (int32_t *)calloc((size_t)e->nf + 2, sizeof(int32_t));
    e->mel_filters = (double *)calloc((size_t)e->nf * (e->nf + 2), sizeof(double));
    if (orc_mel_init(p, e->win, e->binpts, NULL, e->mel_filters) != 0) { orc_env_destroy(e); return -2; }
    int nact = 0;
    for (int i = 0; i < nspecs; i++) if (!specs[i].off) nact++;
    e->gabor_nf = nact;
    if (nact > 0) {
        e->gabor = (double *)calloc((size_t)nact * p->size_y * p->size_x, sizeof(double));
        orc_gabor_to_tensor(p, specs, nspecs, e->gabor);
        if (p->pools_x == 0 && p->pools_y == 0) {
            e->gabor_dims = 2; e->gshape[0] = p->units_y; e->gshape[1] = p->units_x;
            e->gabor_len = (long)p->units_y * p->units_x;
        } else if (p->pools_x > 0 && p->pools_y > 0) {
            e->gabor_dims = 4; e->gshape[0] = p->pools_y; e->gshape[1] = p->pools_x;
            e->gshape[2] = p->units_y; e->gshape[3] = p->units_x;
            e->gabor_len = (long)p->pools_y * p->pools_x * p->units_y * p->units_x;
        } else { orc_env_destroy(e); return -3; }
        e->gabor_out = (float *)calloc((size_t)e->gabor_len, sizeof(float));
    }
    if (e->S > e->B) { orc_env_destroy(e); return -4; }   /* energy loop would index past the tensor */
    const int N = e->win, B = e->B, S = e->S;
    e->window = (double *)calloc((size_t)N * 4 + (size_t)B * 2, sizeof(double));
    e->cre = e->window + N; e->cim = e->cre + N; e->zero = e->cim + N;
    e->power = e->zero + N; e->logpower = e->power + B;
    size_t tot = (size_t)B * S * 2 + (size_t)e->nf * 2 + (size_t)e->nf * S + S + (size_t)e->ncoef * S * 3;
    e->power_seg = (double *)calloc(tot, sizeof(double));
    e->logpower_seg = e->power_seg + (size_t)B * S;
    e->mel_fbank = e->logpower_seg + (size_t)B * S;
    e->dct_tmp = e->mel_fbank + e->nf;
    e->mel_seg = e->dct_tmp + e->nf;
    e->energy = e->mel_seg + (size_t)e->nf * S;
    e->mfcc_seg = e->energy + S;
    e->deltas = e->mfcc_seg + (size_t)e->ncoef * S;
    e->ddeltas = e->deltas + (size_t)e->ncoef * S;
    if (!p->rebuild_plan) { plan_init(&e->plan, N); e->have_plan = 1; }
    *out = e;
    return 0;
}

ORC_API int32_t orc_env_dims(const orc_env *e, int32_t *dims /* win, step, stride, S, B, nf, ncoef, gabor_nf, gabor_len, seg_samples */) {
    dims[0] = e->win; dims[1] = e->step; dims[2] = e->stride; dims[3] = e->S; dims[4] = e->B;
    dims[5] = e->nf; dims[6] = e->ncoef; dims[7] = e->gabor_nf; dims[8] = (int32_t)e->gabor_len; dims[9] = e->seg_samples;
    return 0;
}
ORC_API const int32_t *orc_env_binpts(const orc_env *e) { return e->binpts; }
ORC_API const double *orc_env_mel_filters(const orc_env *e) { return e->mel_filters; }
ORC_API const double *orc_env_gabor(const orc_env *e) { return e->gabor; }

/* sound/sndenv.go:263-265 SegCnt (Channels == 1); Go int division truncates */
ORC_API int32_t orc_env_seg_count(const orc_env *e, int32_t len) {
    int siglen = len - e->seg_samples;
    return siglen / e->stride + 1;
}

/* dft/dft.go:42-85 Filter + Power for one step */
static void dft_filter(orc_env *e, int step, const double *win) {
    const int N = e->win, B = e->B, S = e->S;
    const orc_params *p = &e->p;
    orc_fft_plan local, *P = &e->plan;
    if (p->rebuild_plan) { plan_init(&local, N); P = &local; }   /* dft.go:45 NewCmplxFFT per frame */
    fft_rec(P, win, e->zero, 1, e->cre, e->cim, N, 0);
    for (int k = 0; k < B; k++) {
        double rl = e->cre[k], im = e->cim[k];
        double powr = rl * rl + im * im;
        if (step > 0) powr = p->prev_smooth * e->power[k] + p->cur_smooth * powr;
        e->power[k] = powr;
        e->power_seg[(size_t)k * S + step] = powr;
        if (p->comp_log_pow) {
            powr += p->log_offset;
            double lp = (powr == 0) ? p->log_min : log(powr);
            e->logpower[k] = lp;
            e->logpower_seg[(size_t)k * S + step] = lp;
        }
    }
    if (p->rebuild_plan) plan_free(&local);
}

/* mel/mel.go:120-153 FilterDft for one step */
static void mel_filter_dft(orc_env *e, int step) {
    const orc_params *p = &e->p;
    const int nf = e->nf, maxb = nf + 2, S = e->S;
    const double rscale = 1.0 / (p->renorm_max - p->renorm_min);
    for (int flt = 0; flt < nf; flt++) {
        const int minb = e->binpts[flt], maxbin = e->binpts[flt + 2];
        double sum = 0.0;
        int fi = 0;
        for (int b = minb; b <= maxbin; b++, fi++) sum += e->mel_filters[(size_t)flt * maxb + fi] * e->power[b];
        sum += p->mel_log_off;
        double val = (sum == 0) ? p->mel_log_min : log(sum);
        if (p->renorm) {
            val -= p->renorm_min;
            if (val < 0.0) val = 0.0;
            val *= rscale;
            if (val > 1.0) val = 1.0;
        }
        e->mel_fbank[flt] = val;
        e->mel_seg[(size_t)flt * S + step] = val;
    }
}

/* mel/mel.go:192-212 CepstrumDct for one step */
static void cepstrum_dct(orc_env *e, int step) {
    orc_dct1(e->mel_fbank, e->nf, e->dct_tmp);
    double el0 = e->dct_tmp[0];
    e->dct_tmp[0] = log(1.0 + el0 * el0);
    for (int i = 0; i < e->ncoef; i++) e->mfcc_seg[(size_t)i * e->S + step] = e->dct_tmp[i];
}

/* sound/sndenv.go:380-404 / :407-431 */
static void deltas(const double *src, double *dst, int ncoef, int S) {
    for (int s = 0; s < S; s++) {
        double prv = 0.0, nxt = 0.0;
        for (int i = 0; i < ncoef; i++) {
            double nume = 0.0;
            for (int n = 1; n <= 2; n++) {
                int sprv = s - n, snxt = s + n;
                if (sprv < 0) sprv = 0;
                if (snxt > S - 1) snxt = S - 1;
                prv += src[(size_t)i * S + sprv];
                nxt += src[(size_t)i * S + snxt];
                nume += (double)n * (nxt - prv);
                dst[(size_t)i * S + s] = nume / (double)(2 * n * n);
            }
        }
    }
}

/* sound/sndenv.go:342-433 ProcessSegment (+ :438-478 ProcessStep/SndToWindow).
 * renorm: mel.InitFilters forces Renorm=false (mel.go:80); p->renorm is
 * honoured only so that tests can exercise the branch. */
static void process_segment(orc_env *e, const double *sig, int len, int segment, int add_samples) {
    const int N = e->win, B = e->B, S = e->S;
    memset(e->power, 0, sizeof(double) * B);
    memset(e->logpower, 0, sizeof(double) * B);
    memset(e->power_seg, 0, sizeof(double) * (size_t)B * S * 2);
    memset(e->energy, 0, sizeof(double) * S);
    memset(e->mel_seg, 0, sizeof(double) * (size_t)e->nf * S);
    if (e->p.mfcc) memset(e->mfcc_seg, 0, sizeof(double) * (size_t)e->ncoef * S);
    for (int s = 0; s < S; s++) {
        const int start = segment * e->stride + e->step * (s - e->p.border_steps) + add_samples;
        const int end = start + N;
        if (end > len) break;                                   /* sndenv.go:458-460, 355-358 */
        const double *win;
        if (start < 0) {
            int nz = (end <= 0) ? N : -start;
            memset(e->window, 0, sizeof(double) * nz);
            if (end > 0) memcpy(e->window + nz, sig, sizeof(double) * end);
            win = e->window;
        } else win = sig + start;
        dft_filter(e, s, win);
        mel_filter_dft(e, s);
        if (e->p.mfcc) cepstrum_dct(e, s);
    }
    for (int s = 0; s < S; s++) {                               /* :360-366, transposed indexing F6 */
        double en = 0.0;
        for (int f = 0; f < S; f++) en += e->logpower_seg[(size_t)s * S + f];
        e->energy[s] = en;
    }
    if (e->p.mfcc) {
        for (int s = 0; s < S; s++) e->mfcc_seg[s] = e->energy[s];
        if (e->p.deltas) {
            deltas(e->mfcc_seg, e->deltas, e->ncoef, S);
            deltas(e->deltas, e->ddeltas, e->ncoef, S);
        }
    }
}

typedef struct orc_outputs {   /* any pointer may be NULL; all [seg][...] row-major float64, gabor float32 */
    double *mel, *mfcc, *deltas, *delta_deltas, *energy, *power, *logpower;
    float *gabor;
} orc_outputs;

/* Process every segment of one signal (float64).  Returns SegCnt or <0. */
ORC_API int32_t orc_env_process(orc_env *e, const double *sig, int32_t len, int32_t add_ms, const orc_outputs *o) {
    const int nseg = orc_env_seg_count(e, len);
    const int add = orc_msec_to_samples((double)add_ms, e->p.sample_rate);
    const size_t S = e->S, B = e->B, nf = e->nf, nc = e->ncoef;
    for (int seg = 0; seg < nseg; seg++) {
        process_segment(e, sig, len, seg, add);
        if (o->mel) memcpy(o->mel + seg * nf * S, e->mel_seg, sizeof(double) * nf * S);
        if (o->energy) memcpy(o->energy + seg * S, e->energy, sizeof(double) * S);
        if (o->mfcc && e->p.mfcc) memcpy(o->mfcc + seg * nc * S, e->mfcc_seg, sizeof(double) * nc * S);
        if (o->deltas && e->p.mfcc && e->p.deltas) memcpy(o->deltas + seg * nc * S, e->deltas, sizeof(double) * nc * S);
        if (o->delta_deltas && e->p.mfcc && e->p.deltas) memcpy(o->delta_deltas + seg * nc * S, e->ddeltas, sizeof(double) * nc * S);
        if (o->power) memcpy(o->power + seg * B * S, e->power_seg, sizeof(double) * B * S);
        if (o->logpower) memcpy(o->logpower + seg * B * S, e->logpower_seg, sizeof(double) * B * S);
        if (e->gabor_nf > 0) {
            /* SndEnv.GborOutput persists across segments (never re-zeroed) */
            int rc = orc_gabor_convolve(e->mel_seg, e->nf, e->S, e->gabor, e->gabor_nf, e->p.size_y, e->p.size_x,
                                        e->p.stride_y, e->p.stride_x, e->p.gain, e->gabor_out, e->gabor_dims,
                                        e->gshape, e->p.by_time);
            if (rc < 0) return -5;
            if (o->gabor) memcpy(o->gabor + (size_t)seg * e->gabor_len, e->gabor_out, sizeof(float) * e->gabor_len);
        }
    }
    return nseg;
}

/* Batch driver for the CPU baseline: float32 utterances (widened to float64
 * exactly as the parity tests do), one env per worker thread (pthreads; like
 * one SndEnv per goroutine), utterances dealt to workers through an atomic
 * counter.  If o is non-NULL its pointers receive float32 copies laid out
 * [global_seg][...] using seg_base[u]; a checksum of the results is returned
 * through *checksum so the work cannot be optimised away.
 * Returns total segments or <0. */
typedef struct orc_outputs_f32 { float *mel, *mfcc, *energy, *gabor; } orc_outputs_f32;

typedef struct orc_job {
    const orc_params *p; const orc_gabor_spec *specs; int32_t nspecs;
    const float *wave; const int64_t *utt_off; const int32_t *utt_len; int32_t n_utt;
    int32_t add_ms; const int64_t *seg_base; const orc_outputs_f32 *o;
    int next;            /* atomic work counter */
    int err;
    int maxlen;
} orc_job;
typedef struct orc_worker { orc_job *job; int64_t total; double csum; } orc_worker;

static void *batch_worker(void *arg) {
    orc_worker *w = (orc_worker *)arg;
    orc_job *j = w->job;
    const orc_params *p = j->p;
    orc_env *e = NULL;
    int rc = orc_env_create(p, j->specs, j->nspecs, &e);
    if (rc != 0) { __atomic_store_n(&j->err, rc, __ATOMIC_RELAXED); return NULL; }
    const size_t S = e->S, nf = e->nf, nc = e->ncoef;
    const int add = orc_msec_to_samples((double)j->add_ms, p->sample_rate);
    double *sig = (double *)malloc(sizeof(double) * (size_t)(j->maxlen > 0 ? j->maxlen : 1));
    for (;;) {
        const int u = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (u >= j->n_utt) break;
        const int len = j->utt_len[u];
        const float *wv = j->wave + j->utt_off[u];
        for (int i = 0; i < len; i++) sig[i] = (double)wv[i];
        const int nseg = orc_env_seg_count(e, len);
        for (int seg = 0; seg < nseg; seg++) {
            process_segment(e, sig, len, seg, add);
            if (e->gabor_nf > 0)
                orc_gabor_convolve(e->mel_seg, e->nf, e->S, e->gabor, e->gabor_nf, p->size_y, p->size_x,
                                   p->stride_y, p->stride_x, p->gain, e->gabor_out, e->gabor_dims,
                                   e->gshape, p->by_time);
            w->csum += e->mel_seg[((size_t)seg * 7) % (nf * S)];
            if (j->o && j->seg_base) {
                const orc_outputs_f32 *o = j->o;
                const size_t g = (size_t)j->seg_base[u] + seg;
                if (o->mel) for (size_t i = 0; i < nf * S; i++) o->mel[g * nf * S + i] = (float)e->mel_seg[i];
                if (o->energy) for (size_t i = 0; i < S; i++) o->energy[g * S + i] = (float)e->energy[i];
                if (o->mfcc && p->mfcc) for (size_t i = 0; i < nc * S; i++) o->mfcc[g * nc * S + i] = (float)e->mfcc_seg[i];
                if (o->gabor && e->gabor_nf > 0) memcpy(o->gabor + g * e->gabor_len, e->gabor_out, sizeof(float) * e->gabor_len);
            }
        }
        if (nseg > 0) w->total += nseg;
    }
    free(sig);
    orc_env_destroy(e);
    return NULL;
}

ORC_API int64_t orc_batch_process_f32(const orc_params *p, const orc_gabor_spec *specs, int32_t nspecs,
                                      const float *wave, const int64_t *utt_off, const int32_t *utt_len,
                                      int32_t n_utt, int32_t add_ms, int32_t nthreads,
                                      const int64_t *seg_base, const orc_outputs_f32 *o, double *checksum) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    orc_job job = {p, specs, nspecs, wave, utt_off, utt_len, n_utt, add_ms, seg_base, o, 0, 0, 0};
    for (int u = 0; u < n_utt; u++) if (utt_len[u] > job.maxlen) job.maxlen = utt_len[u];
    orc_worker workers[256];
    pthread_t tids[256];
    for (int t = 0; t < nthreads; t++) { workers[t].job = &job; workers[t].total = 0; workers[t].csum = 0.0; }
    for (int t = 1; t < nthreads; t++) pthread_create(&tids[t], NULL, batch_worker, &workers[t]);
    batch_worker(&workers[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(tids[t], NULL);
    int64_t total = 0; double csum = 0.0;
    for (int t = 0; t < nthreads; t++) { total += workers[t].total; csum += workers[t].csum; }
    if (checksum) *checksum = csum;
    return job.err ? (int64_t)job.err : total;
}

ORC_API int32_t orc_online_cpus(void) { return (int32_t)sysconf(_SC_NPROCESSORS_ONLN); }
