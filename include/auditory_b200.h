/*
 * auditory_b200.h -- C-ABI of the B200-native speech-feature path of
 * emer/auditory (reference v0.9.8; citations are file:line under the
 * reference tree).
 *
 * The reference is pure Go with no FFI: its boundary for this path is the
 * exported Go API (sound.SndEnv, dft.Params, mel.Params, agabor.FilterSet,
 * etensor in/out).  This header is what a cgo shim inside those packages binds
 * instead of running the Go loops (see INTEGRATION.md and go/).  Plain
 * pointers and sizes only; no CUDA or torch types.
 *
 * Conventions: every function returns AUD_OK (0) or a negative aud_status;
 * aud_last_error() gives the message for the calling thread.  The caller
 * owns every buffer it passes; the library never keeps a caller pointer after
 * a call returns.  One handle belongs to one (GPU, host thread) at a time;
 * handles on different GPUs may be used concurrently.
 */
#ifndef AUDITORY_B200_H_
#define AUDITORY_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AUD_API __attribute__((visibility("default")))

typedef enum aud_status {
    AUD_OK = 0,
    AUD_ERR_INVALID = -1,     /* bad argument */
    AUD_ERR_UNSUPPORTED = -2, /* valid for the reference, not implemented on the GPU path (e.g. windows > 4096 samples) */
    AUD_ERR_CUDA = -3,        /* CUDA runtime / driver failure (no CPU fallback exists) */
    AUD_ERR_NOMEM = -4,
    AUD_ERR_PANIC = -5        /* configuration for which the Go reference panics (index out of range) */
} aud_status;

/* ---------------------------------------------------------------------------
 * Host-side initialisers: the same float64 arithmetic the Go Init code does,
 * so the tables can be compared bit for bit.  No GPU needed.
 * ------------------------------------------------------------------------- */

/* sound.MSecToSamples (sound/sndenv.go:522-524): round-half-away(ms*0.001*rate) */
AUD_API int32_t aud_msec_to_samples(double ms, int32_t sample_rate);

/* mel.FreqToMel / MelToFreq / FreqToBin (mel/mel.go:155-168) */
AUD_API double aud_freq_to_mel(double freq);
AUD_API double aud_mel_to_freq(double mel);
AUD_API int32_t aud_freq_to_bin(double freq, double n_fft, double sample_rate);

/* mel.Params.InitFilters (mel/mel.go:77-117).  bin_pts[n_filters+2],
 * hz_pts[n_filters+2] (may be NULL), filters[n_filters*(n_filters+2)] with the
 * reference's flat stride arithmetic (wide filters spill into the next row).
 * AUD_ERR_PANIC when the reference would index past the table. */
AUD_API int32_t aud_mel_init_filters(int32_t dft_size, int32_t sample_rate, int32_t n_filters, double lo_hz,
                                     double hi_hz, int32_t *bin_pts, double *hz_pts, double *filters);

/* agabor.Filter (agabor/gabor.go:17-42) */
typedef struct aud_gabor_spec {
    int32_t off;
    double wave_len, orientation, sigma_width, sigma_length, phase_offset;
    int32_t circle_edge, circular;
} aud_gabor_spec;

/* agabor.ToTensor incl. Active and Filter.Defaults (agabor/gabor.go:73-222,
 * 329-336).  filters[n_active*size_y*size_x]; returns n_active (>= 0). */
AUD_API int32_t aud_gabor_to_tensor(const aud_gabor_spec *specs, int32_t n_specs, int32_t size_x, int32_t size_y,
                                    int32_t distribute, double *filters);

/* The matrix of gonum fourier.DCT.Transform (FFTPACK cost, unnormalised
 * DCT-I; mel/mel.go:198-202): m[k*n_mel+j], k < n_coefs. */
AUD_API void aud_dct1_matrix(int32_t n_mel, int32_t n_coefs, double *m);

/* ---------------------------------------------------------------------------
 * Parameters of one SndEnv (already converted to samples: SndEnv.Init,
 * sound/sndenv.go:195-267) plus dft.Params, mel.Params and the gabor
 * FilterSet / output tensor geometry.
 * ------------------------------------------------------------------------- */
typedef struct aud_params {
    int32_t sample_rate;
    int32_t win_samples;      /* Params.WinSamples: FFT length (dft/dft.go:42-47).  400 (25 ms at 16 kHz) runs the fused
                                 kernel; any other length up to 4096 runs the general DFT path (same results contract) */
    int32_t step_samples;     /* Params.StepSamples */
    int32_t segment_samples;  /* Params.SegmentSamples (used by SegCnt only) */
    int32_t stride_samples;   /* Params.StrideSamples */
    int32_t segment_steps;    /* Params.SegmentSteps, border steps included */
    int32_t border_steps;     /* Params.BorderSteps */
    /* dft.Params (dft/dft.go:15-31) */
    int32_t comp_log_pow;
    double log_min, log_offset, prev_smooth, cur_smooth;
    /* mel.Params / mel.FilterBank (mel/mel.go:16-66) */
    int32_t n_mel;
    double mel_log_off, mel_log_min;
    int32_t renorm;           /* InitFilters forces this off (mel.go:80) */
    double renorm_min, renorm_scale;
    int32_t mfcc, n_coefs, deltas;
    int32_t mfcc_c0_energy;   /* 1: SndEnv.ProcessSegment overwrites MFCC row 0 with Energy
                                 (sndenv.go:368-372); 0: CepstrumDct's ln(1+y0^2) (mel.go:203-204) */
    /* agabor.FilterSet (agabor/gabor.go:45-70); gabor_nf == 0 disables the stage */
    int32_t gabor_nf, gabor_size_x, gabor_size_y, gabor_stride_x, gabor_stride_y;
    double gabor_gain;
    int32_t gabor_out_dims;   /* 2 or 4: NumDims of the rawOut tensor (gabor.go:234-262) */
    int32_t gabor_shape[4];   /* its shape: [UnitsY,UnitsX] or [PoolsY,PoolsX,UnitsY,UnitsX] (sndenv.go:214-223) */
    int32_t gabor_by_time;
} aud_params;

/* Fill the sample counts from milliseconds exactly as SndEnv.Init does
 * (sndenv.go:202-207) and everything else with SndEnv.Defaults() /
 * DFT.Defaults() / Mel.Defaults() values (sndenv.go:64-71, dft.go:33-39,
 * mel.go:69-74,171-180), MFCC and Deltas ON as in the reference,
 * gabor disabled. */
AUD_API int32_t aud_params_defaults(aud_params *p, int32_t sample_rate, double win_ms, double step_ms,
                                    double segment_ms, double stride_ms, int32_t border_steps);

/* Shapes of the per-segment outputs for these parameters. */
typedef struct aud_dims {
    int32_t segment_steps;    /* S */
    int32_t n_bins;           /* win_samples/2+1 */
    int32_t n_mel, n_coefs;
    int64_t gabor_len;        /* product of gabor_shape, 0 if disabled */
} aud_dims;

typedef struct aud_handle aud_handle;

/* Create a pipeline on CUDA device `device`.  mel_bin_pts[n_mel+2] and
 * mel_filters[n_mel*(n_mel+2)] are mel.Params.BinPts and MelFilters.Values;
 * gabor_filters[gabor_nf*size_y*size_x] is FilterSet.Filters.Values (NULL when
 * gabor_nf == 0); dct[n_coefs*n_mel] may be NULL (built-in DCT-I).  Tables are
 * float64 as in Go and are narrowed to float32 once, here. */
AUD_API int32_t aud_create(const aud_params *p, const int32_t *mel_bin_pts, const double *mel_filters,
                           const double *gabor_filters, const double *dct, int32_t device, aud_handle **out);
AUD_API void aud_destroy(aud_handle *h);
AUD_API int32_t aud_get_dims(const aud_handle *h, aud_dims *d);

/* SegCnt of SndEnv.Init (sndenv.go:263-265) for a mono signal of n samples;
 * never negative. */
AUD_API int32_t aud_seg_count(const aud_handle *h, int32_t n_samples);
/* Sum of aud_seg_count over a batch; seg_base (may be NULL) receives
 * n_utt+1 prefix sums: utterance u owns segments [seg_base[u], seg_base[u+1]). */
AUD_API int64_t aud_total_segments(const aud_handle *h, const int32_t *utt_len, int32_t n_utt, int64_t *seg_base);

/* A batch of mono utterances stored in one float32 buffer.  utt_offset and
 * utt_len are HOST arrays in both entry points. */
typedef struct aud_batch {
    const float *wave;
    const int64_t *utt_offset; /* sample index of utterance u in wave */
    const int32_t *utt_len;    /* samples in utterance u */
    int32_t n_utt;
    int32_t add_samples;       /* MSecToSamples(add) of ProcessSegment(segment, add) (sndenv.go:440) */
} aud_batch;

/* Per-segment outputs, all [total_segments][...] row-major float32 in the
 * reference's per-segment tensor layout; NULL = not wanted.
 *   mel          [seg][n_mel][S]     MelFBankSegment   (sndenv.go:254)
 *   mfcc         [seg][n_coefs][S]   MFCCSegment       (sndenv.go:258)
 *   deltas, delta_deltas  same shape (sndenv.go:259-260, 378-432)
 *   energy       [seg][S]            Energy            (sndenv.go:255, 360-366)
 *   gabor        [seg][gabor_len]    GborOutput        (sndenv.go:214-219)
 *   power, logpower [seg][n_bins][S] PowerSegment / LogPowerSegment (sndenv.go:235-238)
 */
typedef struct aud_outputs {
    float *mel, *mfcc, *deltas, *delta_deltas, *energy, *gabor, *power, *logpower;
} aud_outputs;

/* Host buffers in, host buffers out -- the call the Go shim makes in place of the SndEnv.ProcessSegment /
 * ApplyGabor loops (sound/sndenv.go:342-433, 481-497).  Groups of utterances are pipelined over three streams
 * (copy in, compute, copy out) and the call returns when the outputs are complete.  Caller buffers that are already
 * page-locked (aud_host_alloc, cudaHostRegister) are copied from / to directly; large ordinary (pageable) buffers --
 * a Go slice, a numpy array -- are staged through the handle's own pinned bounce buffers (two per direction,
 * 8 MB chunks) by a few copy threads, so that the host copy of one chunk overlaps the DMA of the previous one
 * (aud_set_option "pin": 1 = page-lock the caller's buffers for the duration of the call instead, 2 = leave them to
 * the driver); small ones are left to the driver's staging.  No caller pointer is kept after the return. */
AUD_API int32_t aud_process_host(aud_handle *h, const aud_batch *b, const aud_outputs *o);

/* The same batch over several GPUs of one box: handles[g] was created with the same parameters on GPU g (any
 * distinct devices; two handles on one device also work).  Utterances are cut into n_handles contiguous blocks with
 * about equal numbers of segments; one host thread per handle runs aud_process_host on its block and writes its
 * own disjoint range of the caller's output tensors.  No collective (segments are independent: dft/dft.go:67,
 * sound/sndenv.go:343-351).  Returns the first failing GPU's status; aud_last_error() names it. */
AUD_API int32_t aud_process_host_multi(aud_handle *const *handles, int32_t n_handles, const aud_batch *b,
                                       const aud_outputs *o);
AUD_API int32_t aud_process_host_multi_i16(aud_handle *const *handles, int32_t n_handles, const int16_t *wave,
                                           const int64_t *utt_offset, const int32_t *utt_len, int32_t n_utt,
                                           int32_t add_samples, const aud_outputs *o);

/* Device buffers in and out (wave and every non-NULL output are device
 * pointers on the handle's GPU); the work is enqueued on `cuda_stream`
 * (a cudaStream_t, NULL = default stream) and the call does not synchronise. */
AUD_API int32_t aud_process_device(aud_handle *h, const aud_batch *b, const aud_outputs *o, void *cuda_stream);

/* The same two entry points for 16-bit PCM input, as decoded from a WAV file before
 * Wave.SoundToTensor / GetFloatAtIdx normalise it (sound/sound.go:116-141): sample = int16 / 0x7FFF,
 * applied on the GPU.  Halves the host-to-device bytes of the path. */
AUD_API int32_t aud_process_host_i16(aud_handle *h, const int16_t *wave, const int64_t *utt_offset,
                                     const int32_t *utt_len, int32_t n_utt, int32_t add_samples, const aud_outputs *o);
AUD_API int32_t aud_process_device_i16(aud_handle *h, const int16_t *wave, const int64_t *utt_offset,
                                       const int32_t *utt_len, int32_t n_utt, int32_t add_samples, const aud_outputs *o,
                                       void *cuda_stream);

/* agabor.Convolve as an operator of its own (agabor/gabor.go:225-315), for callers that hold a mel tensor
 * already (examples/gaborview/gbv.go:799-836): `n` input tensors [n_mel][steps] (row-major float32, host
 * memory), filters[nf*size_y*size_x] = FilterSet.Filters.Values, `out` = n raw output tensors of the given
 * 2-D / 4-D shape.  As in the reference, cells the convolution does not reach keep the values `out` came in
 * with, and nothing is written when the filter is wider than the input (gabor.go:226-229). */
AUD_API int32_t aud_gabor_convolve(int32_t device, const float *mel, int32_t n, int32_t n_mel, int32_t steps,
                                   const double *filters, int32_t nf, int32_t size_x, int32_t size_y, int32_t stride_x,
                                   int32_t stride_y, double gain, int32_t out_dims, const int32_t *out_shape,
                                   int32_t by_time, float *out);

/* ---------------------------------------------------------------------------
 * Per-step operators, for callers that drive dft.Filter / mel.FilterDft / mel.CepstrumDct themselves instead of
 * SndEnv.ProcessSegment (examples/gaborview/gbv.go:545-559, 627-641).  Each call covers every step of ONE segment
 * (the reference calls the operator once per step in a loop over the segment); host pointers in and out; tensors in
 * the reference's *Segment layouts (row-major, step fastest).  Not on the timed path: a temporary pipeline per call.
 * ------------------------------------------------------------------------- */

/* dft.Params (dft/dft.go:15-31) */
typedef struct aud_dft_params {
    int32_t comp_log_pow;
    double log_min, log_offset, prev_smooth, cur_smooth;
} aud_dft_params;

/* dft.Params.Filter (dft/dft.go:42-85) for steps 0 .. n_steps-1: windows[n_steps][win_samples] are the frames
 * SndToWindow produced; power_segment / log_power_segment [win_samples/2+1][n_steps] as PowerSegment /
 * LogPowerSegment (either may be NULL).  Step 0 is not smoothed, later steps are smoothed against the previous
 * step's smoothed power (dft.go:66-68).  win_samples <= 4096. */
AUD_API int32_t aud_dft_filter(int32_t device, const aud_dft_params *dp, const float *windows, int32_t n_steps,
                               int32_t win_samples, float *power_segment, float *log_power_segment);

/* mel.FilterBank scalars (mel/mel.go:16-44) */
typedef struct aud_mel_params {
    int32_t n_filters;
    double log_off, log_min;
    int32_t renorm;
    double renorm_min, renorm_scale;
} aud_mel_params;

/* mel.Params.FilterDft (mel/mel.go:120-153) for every step: power_segment[n_bins][n_steps] (already smoothed, as
 * dft.Filter leaves it) -> mel_segment[n_filters][n_steps]; bin_pts[n_filters+2] and filters[n_filters*(n_filters+2)]
 * are mel.Params.BinPts and the filter tensor's Values. */
AUD_API int32_t aud_mel_filter_dft(int32_t device, const aud_mel_params *mp, const int32_t *bin_pts, const double *filters,
                                   const float *power_segment, int32_t n_bins, int32_t n_steps, float *mel_segment);

/* mel.Params.CepstrumDct (mel/mel.go:192-212) for every step: mel_segment[n_filters][n_steps] ->
 * mfcc_segment[n_coefs][n_steps], coefficient 0 = ln(1 + y0^2) (mel.go:203-204; SndEnv overwrites it with Energy
 * afterwards, sndenv.go:368-372 -- that is the caller's step).  dct[n_coefs*n_filters] may be NULL (built-in DCT-I). */
AUD_API int32_t aud_cepstrum_dct(int32_t device, const float *mel_segment, int32_t n_filters, int32_t n_steps,
                                 int32_t n_coefs, const double *dct, float *mfcc_segment);

/* ---------------------------------------------------------------------------
 * After agabor.Convolve: SndEnv.ApplyNeighInhib + SndEnv.ApplyKwta (sound/sndenv.go:303-323, called from
 * ApplyGabor :481-497).  The algorithms are emer/vision v1.1.15 kwta.{KWTA,NeighInhib} and emer/leabra v1.1.48
 * fffb / nxx1 (go.mod:8-9) -- third-party code that is not in the reference tree: restated from the published FFFB /
 * noisy-XX1 equations, PARITY UNPINNED (the tests hold it against a float32 numpy twin).  Float32, same operation order as the Go
 * loops.  Host pointers in and out; a post-processing operator, not part of the fused kernel.
 * ------------------------------------------------------------------------- */
typedef struct aud_fffb_params {       /* leabra fffb.Params */
    int32_t on;
    float gi, ff, fb, fb_tau, max_vs_avg, ff0;
} aud_fffb_params;

typedef struct aud_kwta_params {       /* vision kwta.KWTA + kwta.NeighInhib + SndEnv.KwtaPool */
    int32_t on, iters;
    float del_act_thr;
    aud_fffb_params lay_fffb, pool_fffb;
    float xx1_thr, xx1_gain, xx1_nvar, xx1_vm_act_thr, xx1_sig_mult, xx1_sig_mult_pow, xx1_sig_gain, xx1_interp_range,
        xx1_gain_cor_range, xx1_gain_cor;                       /* leabra nxx1.Params */
    float act_tau;
    float gbar_e, gbar_l, gbar_i, gbar_k, erev_e, erev_l, erev_i, erev_k;   /* chans.Chans */
    int32_t pool_mode;                 /* SndEnv.KwtaPool: KWTAPool (1) or KWTALayer (0) */
    int32_t neigh_on;                  /* NeighInhib.On */
    float neigh_gi;
} aud_kwta_params;

/* KWTA.Defaults() / NeighInhib.Defaults() values; kwta on, layer mode, neighbour inhibition off. */
AUD_API void aud_kwta_defaults(aud_kwta_params *p);

/* gabor[n_tensors][len] are GborOutput tensors of shape[dims] (dims 2 or 4; NeighInhib and KWTAPool need 4:
 * [PoolsY, PoolsX, UnitsY, UnitsX]); kwta[n_tensors][len] receives GborKwta, ext_gi (may be NULL) ExtGi.
 * KWTAPool keeps each pool's feedback inhibition from one call to the next (SndEnv.Inhibs): seq_base[n_seq+1] gives
 * the runs of tensors that one SndEnv would have produced in order (one run per utterance; NULL: a single run). */
AUD_API int32_t aud_apply_kwta(int32_t device, const aud_kwta_params *kp, const float *gabor, int32_t n_tensors,
                               int32_t dims, const int32_t *shape, const int64_t *seq_base, int32_t n_seq, float *ext_gi,
                               float *kwta);

/* Page-locked host memory for callers that want to skip the per-call page-locking of large buffers. */
AUD_API void *aud_host_alloc(uint64_t bytes);
AUD_API void aud_host_free(void *p);

/* Number of kernels this handle has launched so far. */
AUD_API int64_t aud_launch_count(const aud_handle *h);
/* Tuning knobs of the fused kernel: "warps" (FFT warps per CTA), "epi" (epilogue warps), "job_segs" (segments per
 * job), "ctas" (persistent grid size), "groups" (utterance groups of the host-path copy/compute pipeline), "pin" (pageable caller
 * buffers: 0 staged through pinned bounce buffers, 1 page-locked for the call, 2 driver staging), "copy_threads" (host
 * threads of the staged copy including the caller, before the first staged call; measured best at the default 4);
 * 0 = auto.  "dft_tc": frame power of the general window-length route on the tensor cores (1, default) or by the FP32
 * FMA kernel (0). */
AUD_API int32_t aud_set_option(aud_handle *h, const char *name, int64_t value);

/* Measured FP32 (non-tensor) throughput of the device: register-resident dependency chains, CUDA-event timed.
 * kind 0: scalar FFMA; 1: packed FFMA2 (fma.rn.f32x2); 2: FADD:FMUL:FFMA = 2:1:1 (an FFT butterfly's mix);
 * 3: the same mix as FADD2 / FMUL2 / FFMA2.  tflops: achieved TFLOP/s; ginst_per_s (may be NULL): lane-instructions
 * per nanosecond.  The denominators of the FP32 roofline bench.py reports; no reference counterpart. */
AUD_API int32_t aud_measure_fp32(int32_t device, int32_t kind, double *tflops, double *ginst_per_s);

AUD_API const char *aud_last_error(void);
AUD_API int32_t aud_version(void);

#ifdef __cplusplus
}
#endif
#endif /* AUDITORY_B200_H_ */
