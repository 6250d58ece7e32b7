// aud_tables.cpp -- host-side initialisers of the speech-feature path: the
// float64 arithmetic that SndEnv.Init / mel.InitFilters / agabor.ToTensor run
// once per configuration.  Pure C++, no CUDA.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "auditory_b200.h"
#include "aud_internal.h"

static const double kPi = 3.14159265358979323846264338327950288;

extern "C" {

// sound/sndenv.go:522-524.  Go's math.Round rounds half away from zero, which
// is what std::round does.
int32_t aud_msec_to_samples(double ms, int32_t sample_rate) {
    return static_cast<int32_t>(std::round(ms * 0.001 * static_cast<double>(sample_rate)));
}

// mel/mel.go:155-168
double aud_freq_to_mel(double freq) { return 1127.0 * std::log(1.0 + freq / 700.0); }
double aud_mel_to_freq(double mel) { return 700.0 * (std::exp(mel / 1127.0) - 1.0); }
int32_t aud_freq_to_bin(double freq, double n_fft, double sample_rate) {
    return static_cast<int32_t>(std::floor(((n_fft + 1) * freq) / sample_rate));
}

// mel/mel.go:77-117.  The table is [n_filters][n_filters+2] but is written
// through etensor's flat stride arithmetic, so a filter wider than a row
// continues into the next row (and is then partly overwritten by that row's
// own filter); past the end of the table the reference panics.
int32_t aud_mel_init_filters(int32_t dft_size, int32_t sample_rate, int32_t n_filters, double lo_hz, double hi_hz,
                             int32_t *bin_pts, double *hz_pts, double *filters) {
    if (!bin_pts || !filters || n_filters < 1 || dft_size < 2 || sample_rate <= 0)
        return aud::fail(AUD_ERR_INVALID, "aud_mel_init_filters: bad argument");
    const int npts = n_filters + 2;
    const double mel_hi = aud_freq_to_mel(hi_hz);
    const double mel_lo = aud_freq_to_mel(lo_hz);
    const double incr = (mel_hi - mel_lo) / static_cast<double>(n_filters + 1);
    for (int i = 0; i < npts; ++i) {
        const double hz = aud_mel_to_freq(mel_lo + static_cast<double>(i) * incr);
        if (hz_pts) hz_pts[i] = hz;
        bin_pts[i] = aud_freq_to_bin(hz, static_cast<double>(dft_size), static_cast<double>(sample_rate));
    }
    const int64_t table_len = static_cast<int64_t>(n_filters) * npts;
    std::memset(filters, 0, sizeof(double) * static_cast<size_t>(table_len));
    for (int f = 0; f < n_filters; ++f) {
        const int lo = bin_pts[f], mid = bin_pts[f + 1], hi = bin_pts[f + 2];
        const double rise = static_cast<double>(mid) - static_cast<double>(lo);
        const double fall = static_cast<double>(hi) - static_cast<double>(mid);
        int64_t at = static_cast<int64_t>(f) * npts;
        for (int bin = lo; bin <= hi; ++bin, ++at) {
            if (at >= table_len)
                return aud::fail(AUD_ERR_PANIC, "mel.InitFilters: filter table index out of range "
                                                "(the Go reference panics for this dft size / filter count)");
            filters[at] = (bin <= mid) ? (static_cast<double>(bin) - static_cast<double>(lo)) / rise
                                       : (static_cast<double>(hi) - static_cast<double>(bin)) / fall;
        }
    }
    return AUD_OK;
}

// agabor/gabor.go:73-222, 329-336
int32_t aud_gabor_to_tensor(const aud_gabor_spec *specs, int32_t n_specs, int32_t size_x, int32_t size_y,
                            int32_t distribute, double *filters) {
    if (n_specs < 0 || (n_specs > 0 && !specs) || size_x < 1 || size_y < 1)
        return aud::fail(AUD_ERR_INVALID, "aud_gabor_to_tensor: bad argument");
    std::vector<aud_gabor_spec> active;
    for (int i = 0; i < n_specs; ++i)
        if (!specs[i].off) active.push_back(specs[i]);
    if (active.empty()) return 0;
    if (!filters) return aud::fail(AUD_ERR_INVALID, "aud_gabor_to_tensor: filters is NULL");

    int n_horiz = 1, n_vert = 1;
    if (distribute) {
        n_horiz = n_vert = 0;
        for (const auto &s : active) {
            if (s.orientation == 0) ++n_horiz;
            else if (s.orientation == 90) ++n_vert;
        }
    }
    const double rad_x = size_x / 2.0, rad_y = size_y / 2.0;
    const double mid_x = (size_x - 1) / 2.0, mid_y = (size_y - 1) / 2.0;
    const double h_inc = static_cast<double>(size_y - 1) / static_cast<double>(n_horiz + 1);
    const double v_inc = static_cast<double>(size_x - 1) / static_cast<double>(n_vert + 1);
    int h_seen = 0, v_seen = 0;
    const size_t plane = static_cast<size_t>(size_x) * size_y;

    for (size_t idx = 0; idx < active.size(); ++idx) {
        aud_gabor_spec s = active[idx];
        if (s.wave_len == 0) s.wave_len = 2;                       // Filter.Defaults
        if (s.sigma_length == 0 && !s.circular) s.sigma_length = 0.5;
        if (s.sigma_width == 0) s.sigma_width = 0.5;
        const double k_wave = (2.0 * kPi) / s.wave_len;
        const double inv_len = 1.0 / (2.0 * s.sigma_length * s.sigma_length);
        const double inv_wid = 1.0 / (2.0 * s.sigma_width * s.sigma_width);
        double h_pos = 0, v_pos = 0;
        if (distribute) {
            if (s.orientation == 0) h_pos = h_inc * static_cast<double>(++h_seen);
            if (s.orientation == 90) v_pos = v_inc * static_cast<double>(++v_seen);
        } else {   // counters never advance in this branch of the reference
            h_pos = h_inc * static_cast<double>(h_seen + 1);
            v_pos = v_inc * static_cast<double>(v_seen + 1);
        }
        double *dst = filters + idx * plane;
        for (int y = 0; y < size_y; ++y) {
            for (int x = 0; x < size_x; ++x) {
                double v;
                if (!s.circular) {
                    const double dx = (s.orientation == 90) ? x - v_pos : x - mid_x;
                    const double dy = (s.orientation == 0) ? y - h_pos : y - mid_y;
                    const double ux = dx / rad_x, uy = dy / rad_y;
                    v = 0.0;
                    if (!(s.circle_edge && std::hypot(ux, uy) > 1.0)) {
                        const double th = s.orientation * kPi / 180;
                        const double rx = ux * std::cos(th) - uy * std::sin(th);
                        const double ry = uy * std::cos(th) + ux * std::sin(th);
                        v = std::exp(-(inv_wid * (rx * rx) + inv_len * (ry * ry))) * std::sin(k_wave * ry + s.phase_offset);
                    }
                } else {
                    const double ux = (x - mid_x) / rad_x, uy = (y - mid_y) / rad_y;
                    const double a = ux * ux * inv_wid, b = uy * uy * inv_wid;
                    v = -std::sqrt(a + b) * std::sin(k_wave * a * b);
                }
                dst[static_cast<size_t>(y) * size_x + x] = v;
            }
        }
    }
    // each lobe sums to +-1 (gabor.go:194-221)
    for (size_t idx = 0; idx < active.size(); ++idx) {
        double *f = filters + idx * plane;
        double pos = 0.0, neg = 0.0;
        for (size_t i = 0; i < plane; ++i) {
            if (f[i] > 0) pos += f[i];
            else if (f[i] < 0) neg += f[i];
        }
        const double pos_scale = 1.0 / pos, neg_scale = -1.0 / neg;
        for (size_t i = 0; i < plane; ++i) {
            if (f[i] > 0.0) f[i] *= pos_scale;
            else if (f[i] < 0.0) f[i] *= neg_scale;
        }
    }
    return static_cast<int32_t>(active.size());
}

// gonum fourier.DCT.Transform = FFTPACK cost:
//   y[k] = x[0] + (-1)^k x[n-1] + 2 sum_{j=1}^{n-2} x[j] cos(pi j k / (n-1))
void aud_dct1_matrix(int32_t n_mel, int32_t n_coefs, double *m) {
    for (int k = 0; k < n_coefs; ++k) {
        double *row = m + static_cast<size_t>(k) * n_mel;
        for (int j = 0; j < n_mel; ++j) {
            if (j == 0) row[j] = 1.0;
            else if (j == n_mel - 1) row[j] = (k & 1) ? -1.0 : 1.0;
            else row[j] = 2.0 * std::cos(kPi * static_cast<double>(j) * static_cast<double>(k) / static_cast<double>(n_mel - 1));
        }
    }
}

int32_t aud_params_defaults(aud_params *p, int32_t sample_rate, double win_ms, double step_ms, double segment_ms,
                            double stride_ms, int32_t border_steps) {
    if (!p) return aud::fail(AUD_ERR_INVALID, "aud_params_defaults: p is NULL");
    if (sample_rate <= 0) return aud::fail(AUD_ERR_INVALID, "sample rate <= 0");   // sndenv.go:197-201
    std::memset(p, 0, sizeof(*p));
    p->sample_rate = sample_rate;
    p->win_samples = aud_msec_to_samples(win_ms, sample_rate);
    p->step_samples = aud_msec_to_samples(step_ms, sample_rate);
    p->segment_samples = aud_msec_to_samples(segment_ms, sample_rate);
    p->stride_samples = aud_msec_to_samples(stride_ms, sample_rate);
    p->border_steps = border_steps;
    p->segment_steps = static_cast<int32_t>(std::round(segment_ms / step_ms)) + 2 * border_steps;
    p->comp_log_pow = 1;
    p->log_min = -100.0;
    p->log_offset = 1.0;
    p->prev_smooth = 0.0;
    p->cur_smooth = 1.0;
    p->n_mel = 32;
    p->mel_log_off = 0.0;
    p->mel_log_min = -10.0;
    p->renorm = 0;
    p->renorm_min = -6.0;
    p->renorm_scale = 1.0 / (4.0 - -6.0);
    p->mfcc = 1;
    p->n_coefs = 13;
    p->deltas = 1;
    p->mfcc_c0_energy = 1;
    p->gabor_nf = 0;
    p->gabor_gain = 1.0;
    p->gabor_out_dims = 2;
    return AUD_OK;
}

}  // extern "C"
