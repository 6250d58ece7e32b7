// aud_kernels.cuh -- device code of the fused waveform -> mel / MFCC / gabor
// path for sm_100a (persistent, streaming, warp-specialised design).
//
// One CTA per SM owns a list of *jobs* (runs of consecutive segments of one
// utterance) and walks the concatenated stream of their distinct frames in
// rounds of NWARPS x 3 frame pairs.  The CTA has two kinds of warps, coupled
// only through mbarriers (no CTA-wide barrier after the set-up):
//
// FFT warps -- each an independent engine over its share of the stream:
//   TMA window  : lanes 0..2 bulk-copy (cp.async.bulk + mbarrier) the 560-sample
//                 span of each of the warp's 3 frame pairs into the pair's
//                 scratch, one round ahead of use;
//   FFT         : two real frames ride one complex 400-point FFT, factored
//                 20 x 20 with in-register prime-factor DFT-20s (10 lanes per
//                 pair, 2 columns per lane, one shared-memory transpose);
//   power / mel : Z[k], Z[N-k] are split into |X_A|^2, |X_B|^2 and the banded
//                 mel sums of the RAW power (or their logs when there is no
//                 smoothing) go to a ring indexed by frame; arrive on full[R & 1].
// Epilogue warps -- finish the segments whose last frame landed in round R:
//   smoothing as a scan over steps (it is linear, so it commutes with the mel
//   sums), logs, Energy, DCT, deltas, gabor, and the stores of those final
//   features; arrive on empty[R & 1] so the FFT warps may reuse the ring slots.
// Frames shared by overlapping segments are transformed once.
//
// Reference semantics (file:line under the reference tree):
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

#include "aud_fft_core.cuh"

// Debug build (make debug -> libauditory_b200_dbg.so, -DAUD_DEBUG_CHECKS): compute-sanitizer is closed on the GPU pool
// this was developed on, so the kernel carries its own checks -- guard words between all shared-memory regions (verified
// when the CTA finishes) and asserts on the index arithmetic of the ring, the tiles, the done list, the job table and
// the bulk copies.  A failed check leaves its code in KParams::dbg; the host turns it into an error after the launch.
#ifdef AUD_DEBUG_CHECKS
#define AUD_CHECK(P, cond, code)                        \
    do {                                                \
        if (!(cond) && (P).dbg) atomicMax((P).dbg, (code)); \
    } while (0)
#define AUD_CANARY_BYTES 16
#else
#define AUD_CHECK(P, cond, code) \
    do {                         \
    } while (0)
#define AUD_CANARY_BYTES 0
#endif

namespace aud {

constexpr int kCanaryWord = 0x5afec0de;
constexpr int kSmemRegions = 17;    // regions carve_smem lays out

constexpr int kPowPitch = 208;      // row pitch of the raw-power scratch (parity / inspection outputs)
constexpr int kMaxJobs = 64;        // jobs per CTA
constexpr int kMaxDone = 96;        // segments that can complete in one round (<= frames per round)
constexpr int kMaxRanges = kMaxDone; // jobs that can complete segments in one round (each completes at least one)
constexpr int kDoneMeta = 2 + 4 * kMaxRanges + 2;   // ints per done-list header
constexpr int kRecRounds = 5;       // rounds of frame-pair records alive at once when the epilogue writes them four rounds ahead

struct Job {
    long long wave_off;   // index of the utterance's first sample in the wave buffer
    long long out_seg;    // global index of this job's first segment
    int utt_len;
    int seg0;             // first segment of the job within its utterance
    int nseg;
    int nframes;          // distinct frame slots of the job
    int pair_base;        // pairs of this CTA's earlier jobs (stream position of the job)
    int frame_base;       // first row of this job in the raw-power scratch
};

struct KParams {
    // geometry
    int step, stride, S, border, add;
    int seg_adv;          // frame slots between consecutive segments: stride/step if frames are shared, else S
    int dedupe;           // 1: frame slot f of a job starts f*step after the job's first frame
    int n_mel, n_coefs;
    int ps;               // per-pair scratch stride in float2 units (>= kExchange, >= kWinOff + window; == 10 mod 16
                          // so that the three pairs' windows sit 20 banks apart: conflict-free 8-byte loads)
    int win_len;          // floats copied per pair window in contiguous mode (step + 400)
    int contig;           // 1: frame B = frame A + step inside one window; 0: two 400-sample copies
    int ring;             // frame ring slots: >= 2 * frames per round + S (one barrier per round)
    int mel_pitch;        // row pitch of the frame ring: mel_ring_pitch(n_mel), odd and > n_mel
    int nosmooth;         // PrevSmooth == 0 && CurSmooth == 1: log-mel is per frame, phase 2 only gathers
    int rec_rounds;       // frame-pair record buffers: kRecRounds when the epilogue warps write them, else 0
    int energy_bins;      // low bins kept per frame for Energy (0 = not needed)
    int need_tiles;       // mfcc or gabor requested: phase 2 stages mel tiles in shared memory
    int tile_cap;         // segments that fit in the tile area
    int tile_floats;      // floats of the tile area (0 unless need_tiles)
    int t_off[5];         // float offsets of the energy / mfcc / delta / delta-delta / gabor tiles in it
    int dct_floats;       // n_coefs * ceil(n_mel/4)*4 when MFCC is requested, else 0
    int gw_floats;        // g_sy*g_sx * ceil(g_nf/8)*8 when gabor is requested, else 0
    // dft.Params / mel.FilterBank scalars
    float prev, cur, log_off, log_min;
    int comp_log_pow, log1p_path;
    float mel_log_off, mel_log_min;
    int renorm;
    float renorm_min, renorm_scale;
    int want_mfcc;        // Mel.MFCC and an MFCC / delta output requested
    int do_deltas, c0_energy;
    // gabor
    int g_on, g_nf, g_sx, g_sy, g_stx, g_sty, g_dims, g_by_time, g_nt, g_nfy, g_tmaxstrides, g_len;
    int g_str0, g_str1, g_str2;
    int g_keep;           // 1: cells Convolve does not write keep the caller's values (stand-alone operator)
    int g_direct;         // 1: dense 4-D output [PoolsY][PoolsX][2][8] covered by the positions: results go straight to global memory
    float g_gain;
    // tables (device)
    const float2 *tw2;      // [20][10] (1/2) W400^{2j*k1} at k1*10 + j
    const int *mel_start;   // [n_mel] even-aligned first padded power index of each filter
    const int *mel_quads;   // [n_mel] groups of 4 taps (zero padded) per filter
    const float *mel_taps;  // [slot][quad][lane][4]: taps of each lane's task, zero padded (pads, alignment slack, short rows)
    const int4 *mel_sched;  // [mel_tasks][32] {tap offset, power offset in the warp's scratch,
                            //   max quads of the slot << 24 | pair << 16 | filter (or -1: idle lane), 0}
    int mel_taps_len, mel_tasks;
    const float *dct;       // [n_coefs][n_mel]
    const float *gabor;     // [nf][sy][sx]
    // io (device)
    const void *wave;       // float32 samples, or int16 PCM when in_i16 (normalised by 1/0x7FFF: sound/sound.go:130-141)
    int in_i16;
    const Job *jobs;
    const int2 *cta_jobs;   // per CTA: [begin, end) into jobs
    float *o_mel, *o_mfcc, *o_d1, *o_d2, *o_energy, *o_gabor;   // any may be NULL
    float *rawpow;          // [frame rows][kPowPitch] raw |X|^2, only when power / logpower are requested
    int *dbg;               // debug builds: highest failed check code (0 = none); NULL otherwise
};

// The frame ring keeps copies of its first rows behind its last one, so that the S consecutive frames of a short
// segment can be read without a wrap test (the no-smoothing gather); segments longer than this take the wrap path.
constexpr int kMirror = 32;

// Row pitch of the per-frame mel ring: odd (conflict-free column walks) with at least one spare column.
__host__ __device__ inline int mel_ring_pitch(int n_mel) { return (n_mel + 1) | 1; }

// Bytes of dynamic shared memory the fused kernel needs (host and device agree through this).
__host__ __device__ inline size_t fused_smem_bytes(int nwarps, int ps, int mel_taps_len, int n_mel, int mel_tasks,
                                                   int ring, int energy_bins, size_t tile_floats, int rec_rounds) {
    size_t b = 0;
    b += (size_t)nwarps * kPairs * ps * 8;                 // per-pair scratch: exchange / power / next window
    b += (size_t)(kN / 2) * 8;                             // twiddles
    b += (size_t)(kN + 4) * 4;                             // zeros
    b += (size_t)mel_taps_len * 4;                         // taps
    b += 2 * (size_t)((n_mel + 3) & ~3) * 4;               // start, quads
    b += (size_t)mel_tasks * 32 * 16;                      // schedule
    b += (size_t)(((ring + kMirror) * mel_ring_pitch(n_mel) + 3) & ~3) * 4;   // mel ring + mirror rows
    b += (size_t)((ring * energy_bins + 3) & ~3) * 4;      // low-bin ring
    b += (size_t)((tile_floats + 3) & ~(size_t)3) * 4;     // phase-2 tiles + DCT rows (only when MFCC / gabor are requested)
    b += (size_t)nwarps * kPairs * rec_rounds * 32;        // frame-pair records
    b += (size_t)kMaxDone * 16 + (size_t)kDoneMeta * 4;    // done list + counts and ranges
    b += (size_t)((nwarps + 4 + 1) & ~1) * 8;              // mbarriers: per-warp windows, full[2], empty[2]
    b += (size_t)kMaxJobs * sizeof(Job);
    b += (size_t)kSmemRegions * AUD_CANARY_BYTES;          // guard words (debug builds)
    return b;
}

// ------------------------------------------------------------ small helpers
__host__ __device__ __forceinline__ long long floordiv(long long a, long long b) {   // b > 0
    long long q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
}

// number of leading steps of segment `seg` whose window lies inside the signal
// (sndenv.go:457-460: the first window that runs past the end aborts the rest)
__host__ __device__ __forceinline__ int valid_steps(int utt_len, int add, int stride, int step, int border, int S,
                                                    int seg, int win = kN) {
    const long long room = (long long)utt_len - win - add - (long long)seg * stride;
    const long long last = floordiv(room, step) + border;   // largest valid step index
    if (last < 0) return 0;
    return last + 1 > S ? S : (int)(last + 1);
}

__device__ __forceinline__ float ipowf(float b, int n) {   // b^n, n >= 0, by squaring
    float r = 1.f;
    while (n) {
        if (n & 1) r *= b;
        b *= b;
        n >>= 1;
    }
    return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the warp may sleep in hardware until the phase completes (or the
// hint runs out) instead of spinning through issue slots its neighbours on the scheduler need
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }


__device__ __forceinline__ int floordiv32(int a, int b) {   // b > 0
    const int q = a / b;
    return (a < 0 && q * b != a) ? q - 1 : q;
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Shared-memory map of the fused kernel (order mirrors fused_smem_bytes).
struct Smem {
    float2 *scr;       // [NWARPS][kPairs][ps]   exchange rows / power buffer / next round's sample window
    float2 *tw2;       // [200]
    float *zeros;      // [404]
    float *taps;       // [slot][quad][lane][4] mel taps in task order (mel_taps_len floats)
    int *mstart, *mquads;
    int4 *sched;
    float *rmel;       // [ring][mel_pitch]   per-frame mel sums (or ln mel without smoothing)
    float *rlow;       // [ring][energy_bins] per-frame low power bins
    float *tiles;      // phase-2 tiles (MFCC / gabor only)
    float *dct;        // [n_coefs][ceil(n_mel/4)*4] DCT-I rows, zero padded (MFCC only)
    float *gw;         // [g_sy*g_sx][ceil(g_nf/8)*8] gabor weights, tap-major, zero padded (gabor only)
    int4 *prec;        // [kRecRounds][NWARPS * kPairs][2] frame-pair records (see make_record)
    int4 *done;        // [kMaxDone] segments finished by the round being closed
    int *dmeta;        // counts + per-job ranges of the done list
    uint64_t *mbar;    // [NWARPS] window barriers, then full[2], empty[2]
    Job *jobs;
};

__device__ __forceinline__ Smem carve_smem(unsigned char *sp, const KParams &P, int nwarps) {
    Smem m;
    m.scr = reinterpret_cast<float2 *>(sp);      sp += AUD_CANARY_BYTES + (size_t)nwarps * kPairs * P.ps * 8;
    m.tw2 = reinterpret_cast<float2 *>(sp);      sp += AUD_CANARY_BYTES + (size_t)(kN / 2) * 8;
    m.zeros = reinterpret_cast<float *>(sp);     sp += AUD_CANARY_BYTES + (size_t)(kN + 4) * 4;
    m.taps = reinterpret_cast<float *>(sp);      sp += AUD_CANARY_BYTES + (size_t)P.mel_taps_len * 4;
    m.mstart = reinterpret_cast<int *>(sp);      sp += AUD_CANARY_BYTES + (size_t)((P.n_mel + 3) & ~3) * 4;
    m.mquads = reinterpret_cast<int *>(sp);      sp += AUD_CANARY_BYTES + (size_t)((P.n_mel + 3) & ~3) * 4;
    m.sched = reinterpret_cast<int4 *>(sp);      sp += AUD_CANARY_BYTES + (size_t)P.mel_tasks * 32 * 16;
    m.rmel = reinterpret_cast<float *>(sp);      sp += AUD_CANARY_BYTES + (size_t)(((P.ring + kMirror) * P.mel_pitch + 3) & ~3) * 4;
    m.rlow = reinterpret_cast<float *>(sp);      sp += AUD_CANARY_BYTES + (size_t)((P.ring * P.energy_bins + 3) & ~3) * 4;
    m.tiles = reinterpret_cast<float *>(sp);     sp += AUD_CANARY_BYTES + (size_t)((P.tile_floats + 3) & ~3) * 4;
    m.dct = reinterpret_cast<float *>(sp);       sp += AUD_CANARY_BYTES + (size_t)P.dct_floats * 4;
    m.gw = reinterpret_cast<float *>(sp);        sp += AUD_CANARY_BYTES + (size_t)P.gw_floats * 4;
    m.prec = reinterpret_cast<int4 *>(sp);       sp += AUD_CANARY_BYTES + (size_t)nwarps * kPairs * P.rec_rounds * 32;
    m.done = reinterpret_cast<int4 *>(sp);       sp += AUD_CANARY_BYTES + (size_t)kMaxDone * 16;
    m.dmeta = reinterpret_cast<int *>(sp);       sp += AUD_CANARY_BYTES + (size_t)kDoneMeta * 4;
    m.mbar = reinterpret_cast<uint64_t *>(sp);   sp += AUD_CANARY_BYTES + (size_t)((nwarps + 4 + 1) & ~1) * 8;
    m.jobs = reinterpret_cast<Job *>(sp);
    return m;
}

#ifdef AUD_DEBUG_CHECKS
// The guard words sit right behind every region: region r's guard starts at (start of region r + 1) - 16 bytes.
// which = 0: write them, 1: verify them (check codes 900 + region).
__device__ inline void canaries(const KParams &P, const Smem &m, int nwarps, int which) {
    const void *starts[kSmemRegions] = {m.tw2, m.zeros, m.taps, m.mstart, m.mquads, m.sched, m.rmel, m.rlow, m.tiles, m.dct, m.gw,
                                        m.prec, m.done, m.dmeta, m.mbar, m.jobs,
                                        reinterpret_cast<const unsigned char *>(m.jobs) + (size_t)kMaxJobs * sizeof(Job) + AUD_CANARY_BYTES};
    for (int r = 0; r < kSmemRegions; ++r) {
        int *g = reinterpret_cast<int *>(const_cast<unsigned char *>(static_cast<const unsigned char *>(starts[r])) - AUD_CANARY_BYTES);
        for (int w = 0; w < AUD_CANARY_BYTES / 4; ++w) {
            if (which == 0) g[w] = kCanaryWord;
            else AUD_CHECK(P, g[w] == kCanaryWord, 900 + r);
        }
    }
}
#endif

__device__ __forceinline__ int ring_slot(int rbase, int rel, int ring) {
    int sl = rbase + rel;
    if (sl < 0) sl += ring;
    if (sl >= ring) sl -= ring;
    return sl;
}

__device__ __forceinline__ float finish_mel(const KParams &P, float sum) {   // mel/mel.go:133-148
    sum += P.mel_log_off;
    float val = (sum == 0.f) ? P.mel_log_min : __logf(sum);
    if (P.renorm) val = fminf(fmaxf(val - P.renorm_min, 0.f) * P.renorm_scale, 1.f);
    return val;
}

// One frame pair of the CTA's stream.
struct PairInfo {
    int job;          // index into the CTA's job list, -1 = past the end of the stream
    int fa;           // frame slot of frame A inside the job (B = fa + 1)
    int startA;       // sample index of frame A relative to the utterance start (may be negative)
    int startB;
    int has_b;
};

// ------------------------------------------------------------ frame-pair records
// Where pair g of the CTA's stream comes from: {job, frame slot of frame A, B exists, start of A in its
// utterance} and {address of the bulk-copied part of its window (lo, hi), window samples [t0, t1) that part
// covers (t0 | t1 << 16), start of B}.  The 16-byte aligned part of the window that lies inside the utterance
// comes by one TMA bulk copy; the rest (front zero padding, ragged ends, odd alignments, non-contiguous
// frame pairs) is filled by the warp.  Records are computed by the epilogue warps four rounds ahead of use
// (and by everyone for the first four rounds), so the FFT warps only read them.
// `jp`: where to look for the job -- a caller that asks for increasing g passes its running job index (advanced
// here), a caller without state passes -1 and gets a binary search.
__device__ __forceinline__ void make_record(const KParams &P, const Smem &sm, int njobs, int total_pairs, int g,
                                            int4 &r0, int4 &r1, int &jp) {
    r0 = make_int4(-1, 0, 0, 0);
    r1 = make_int4(0, 0, 0, 0);
    if (g >= total_pairs) return;
    int lo;   // the last job whose first pair is <= g
    if (jp >= 0) {
        while (jp + 1 < njobs && sm.jobs[jp + 1].pair_base <= g) ++jp;
        lo = jp;
    } else {
        lo = 0;
        int hi = njobs - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (sm.jobs[mid].pair_base <= g) lo = mid;
            else hi = mid - 1;
        }
    }
    const Job &jb = sm.jobs[lo];
    const int S = P.S, fa = 2 * (g - jb.pair_base), fb = fa + 1;
    int startA, startB;
    if (P.dedupe) {
        startA = jb.seg0 * P.stride + P.add - P.border * P.step + fa * P.step;
        startB = startA + P.step;
    } else {
        const int ca = fa / S, ia = fa - ca * S, cb = fb / S, ib = fb - cb * S;
        startA = (jb.seg0 + ca) * P.stride + P.add + (ia - P.border) * P.step;
        startB = (jb.seg0 + cb) * P.stride + P.add + (ib - P.border) * P.step;
    }
    int t0 = 0, t1 = 0;
    unsigned long long src = 0;
    if (P.contig) {
        const int esz = P.in_i16 ? 2 : 4, al = 16 / esz, n = P.win_len;
        const long long s0 = jb.wave_off + startA;   // wave index of window sample 0 (may precede the utterance)
        if (((reinterpret_cast<uintptr_t>(P.wave) + (uintptr_t)(s0 * esz)) & 15) == 0) {
            const int v0 = max(0, -startA), v1 = min(n, jb.utt_len - startA);
            t0 = (v0 + al - 1) & ~(al - 1);
            t1 = v1 & ~(al - 1);
            if (t1 <= t0) t0 = t1 = 0;
            src = reinterpret_cast<uintptr_t>(P.wave) + (unsigned long long)((s0 + t0) * esz);
        }
    }
    r0 = make_int4(lo, fa, fb < jb.nframes ? 1 : 0, startA);
    r1 = make_int4((int)(unsigned)(src & 0xffffffffu), (int)(unsigned)(src >> 32), t0 | (t1 << 16), startB);
}
// records of one round, spread over `nthreads` cooperating threads
__device__ __forceinline__ void make_round_records(const KParams &P, const Smem &sm, int nwarps, int njobs, int total_pairs,
                                                   int round, int buf, int tid, int nthreads) {
    const int per_round = nwarps * kPairs;
    for (int pi = tid; pi < per_round; pi += nthreads) {
        int4 r0, r1;
        int stateless = -1;
        make_record(P, sm, njobs, total_pairs, round * per_round + pi, r0, r1, stateless);
        int4 *dst = sm.prec + ((size_t)buf * per_round + pi) * 2;
        dst[0] = r0;
        dst[1] = r1;
    }
}

// ------------------------------------------------------------ FFT warps
// One warp = an independent engine over its share of the CTA's frame-pair stream.
// EPIREC: the epilogue warps write the frame-pair records (plain log-mel launches, where they have time to
// spare); otherwise every FFT warp works out its own -- that code then stays out of the EPIREC kernels.
template <int NWARPS, bool EPIREC, bool MELPACK>
__device__ __forceinline__ void fft_role(const KParams &P, const Smem &sm, int warp, int lane, int njobs,
                                         int total_pairs, int rounds) {
    constexpr int FPR = 6 * NWARPS;   // frames per round
    float2 *scr_w = sm.scr + (size_t)warp * kPairs * P.ps;
    uint64_t *bar = &sm.mbar[warp];
    uint64_t *full = &sm.mbar[NWARPS], *empty = &sm.mbar[NWARPS + 2];
    const bool fft_lane = lane < 30;
    const int q = fft_lane ? lane / 10 : 2, j = fft_lane ? lane - 10 * q : 0;
    float2 *scr_q = scr_w + q * P.ps;
    float2 *exq = scr_q + exch_off(q);          // this pair's exchange rows (aud_fft_core.cuh)
    const int p2 = pass2_assign(lane);           // pass-2 assignment of this lane: pair q2, row pair p2p
    const int q2 = p2 >> 5, p2p = p2 & 31;
    float2 *scr_q2 = scr_w + q2 * P.ps;
    const float2 *ex2 = scr_q2 + exch_off(q2);

    // Without epilogue-written records (MFCC / gabor launches, where the epilogue warps have no time to spare):
    // lanes 0..2 each track one pair of the warp's triple themselves, resolve it and stage its window.
    int jp = 0;   // job pointer (per tracking lane), advanced monotonically
    auto resolve_own = [&](int R) {
        PairInfo x;
        x.job = -1; x.fa = 0; x.startA = 0; x.startB = 0; x.has_b = 0;
        const int g = (R * NWARPS + warp) * kPairs + lane;
        if (lane < kPairs && g < total_pairs) {
            while (jp + 1 < njobs && sm.jobs[jp + 1].pair_base <= g) ++jp;
            const Job &jb = sm.jobs[jp];
            x.job = jp;
            x.fa = 2 * (g - jb.pair_base);
            const int fb = x.fa + 1;
            x.has_b = fb < jb.nframes;
            if (P.dedupe) {
                x.startA = jb.seg0 * P.stride + P.add - P.border * P.step + x.fa * P.step;
                x.startB = x.startA + P.step;
            } else {
                const int S = P.S, ca = x.fa / S, ia = x.fa - ca * S, cb = fb / S, ib = fb - cb * S;
                x.startA = (jb.seg0 + ca) * P.stride + P.add + (ia - P.border) * P.step;
                x.startB = (jb.seg0 + cb) * P.stride + P.add + (ib - P.border) * P.step;
            }
        }
        return x;
    };
    // Stage the window of the lane's pair into the upper part of the pair's scratch (free once the
    // exchange rows have been consumed).  The 16-byte aligned part of the window that lies inside the
    // utterance comes by one TMA bulk copy; what is left (front zero padding at an utterance's first
    // frames, the ragged end, odd alignments, non-contiguous frame pairs) the warp fills itself.
    auto stage_own = [&](const PairInfo &pi) {
        const int esz = P.in_i16 ? 2 : 4;
        const int al = 16 / esz;                       // samples per 16 bytes
        const int n = P.contig ? P.win_len : 2 * kN;   // samples in the window
        int t0 = 0, t1 = 0;                            // [t0, t1): window samples the bulk copy brings
        const char *src = nullptr;
        const bool mine = lane < kPairs && pi.job >= 0;
        if (mine && P.contig) {
            const Job &jb = sm.jobs[pi.job];
            const long long s0 = jb.wave_off + pi.startA;   // wave index of window sample 0 (may precede the utterance)
            if (((reinterpret_cast<uintptr_t>(P.wave) + (uintptr_t)(s0 * esz)) & 15) == 0) {
                const int v0 = max(0, -pi.startA), v1 = min(n, jb.utt_len - pi.startA);
                t0 = (v0 + al - 1) & ~(al - 1);
                t1 = v1 & ~(al - 1);
                if (t1 <= t0) t0 = t1 = 0;
                src = static_cast<const char *>(P.wave) + (s0 + t0) * esz;
            }
        }
        const uint32_t bytes = (uint32_t)(t1 - t0) * esz;
        const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
        unsigned slow = __ballot_sync(0xffffffffu, mine && (t1 - t0) != n);
        if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar, total);
        }
        __syncwarp();
        if (bytes) tma_load_1d(reinterpret_cast<char *>(scr_w + lane * P.ps + kWinOff) + t0 * esz, src, bytes, bar);
        while (slow) {   // uniform; utterance edges only
            const int qq = __ffs(slow) - 1;
            slow &= slow - 1;
            const int job = __shfl_sync(0xffffffffu, pi.job, qq);
            const int sA = __shfl_sync(0xffffffffu, pi.startA, qq), sB = __shfl_sync(0xffffffffu, pi.startB, qq);
            const int hb = __shfl_sync(0xffffffffu, pi.has_b, qq);
            const int c0 = __shfl_sync(0xffffffffu, t0, qq), c1 = __shfl_sync(0xffffffffu, t1, qq);
            const Job &jb = sm.jobs[job];
            void *dstv = scr_w + qq * P.ps + kWinOff;
            auto put = [&](int i) {
                const bool second = !P.contig && i >= kN;
                const int a = second ? sB + (i - kN) : sA + i;
                const bool in = (!second || hb) && a >= 0 && a < jb.utt_len;
                if (P.in_i16) static_cast<short *>(dstv)[i] = in ? __ldg(static_cast<const short *>(P.wave) + jb.wave_off + a) : (short)0;
                else static_cast<float *>(dstv)[i] = in ? __ldg(static_cast<const float *>(P.wave) + jb.wave_off + a) : 0.f;
            };
            for (int i = lane; i < c0; i += 32) put(i);        // before the bulk part
            for (int i = c1 + lane; i < n; i += 32) put(i);    // after it (everything when there is none: c0 = c1 = 0)
        }
        __syncwarp();
    };

    // With records: lanes 0..2 each look after one pair of the warp's triple: they read its record and stage its window
    // into the upper part of the pair's scratch (free once the exchange rows have been consumed).
    auto record = [&](int buf, int which) {   // buf: round % rec_rounds, kept as a running counter
        return sm.prec[((size_t)buf * (NWARPS * kPairs) + warp * kPairs + lane) * 2 + which];
    };
    auto stage = [&](int round, int buf) {
        const int esz = P.in_i16 ? 2 : 4;
        const int n = P.contig ? P.win_len : 2 * kN;   // samples in the window
        int4 n0 = make_int4(-1, 0, 0, 0), n1 = make_int4(0, 0, 0, 0);
        if (lane < kPairs) {
            n0 = record(buf, 0);
            n1 = record(buf, 1);
        }
        const bool mine = n0.x >= 0;
        const int t0 = n1.z & 0xffff, t1 = n1.z >> 16;   // [t0, t1): window samples the bulk copy brings
        AUD_CHECK(P, !mine || (n0.x < njobs && t0 <= t1 && t1 <= n && (((t1 - t0) * esz) & 15) == 0 && ((t0 * esz) & 15) == 0), 30);
        const uint32_t bytes = (uint32_t)(t1 - t0) * esz;
        const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
        unsigned slow = __ballot_sync(0xffffffffu, mine && (t1 - t0) != n);
        if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar, total);
        }
        __syncwarp();
        if (bytes) {
            const unsigned long long src = (unsigned long long)(unsigned)n1.x | ((unsigned long long)(unsigned)n1.y << 32);
            tma_load_1d(reinterpret_cast<char *>(scr_w + lane * P.ps + kWinOff) + t0 * esz, reinterpret_cast<const void *>(src), bytes, bar);
        }
        while (slow) {   // uniform; utterance edges only
            const int qq = __ffs(slow) - 1;
            slow &= slow - 1;
            const int job = __shfl_sync(0xffffffffu, n0.x, qq);
            const int sA = __shfl_sync(0xffffffffu, n0.w, qq), sB = __shfl_sync(0xffffffffu, n1.w, qq);
            const int hb = __shfl_sync(0xffffffffu, n0.z, qq);
            const int c0 = __shfl_sync(0xffffffffu, t0, qq), c1 = __shfl_sync(0xffffffffu, t1, qq);
            const Job &jb = sm.jobs[job];
            void *dstv = scr_w + qq * P.ps + kWinOff;
            auto put = [&](int i) {
                const bool second = !P.contig && i >= kN;
                const int a = second ? sB + (i - kN) : sA + i;
                const bool in = (!second || hb) && a >= 0 && a < jb.utt_len;
                if (P.in_i16) static_cast<short *>(dstv)[i] = in ? __ldg(static_cast<const short *>(P.wave) + jb.wave_off + a) : (short)0;
                else static_cast<float *>(dstv)[i] = in ? __ldg(static_cast<const float *>(P.wave) + jb.wave_off + a) : 0.f;
            };
            for (int i = lane; i < c0; i += 32) put(i);        // before the bulk part
            for (int i = c1 + lane; i < n; i += 32) put(i);    // after it (everything when there is none: c0 = c1 = 0)
        }
        __syncwarp();
    };

    PairInfo own_cur, own_nxt;
    own_cur.job = -1; own_cur.fa = 0; own_cur.startA = 0; own_cur.startB = 0; own_cur.has_b = 0;
    if constexpr (EPIREC) stage(0, 0);
    else {
        own_cur = resolve_own(0);
        stage_own(own_cur);
    }
    own_nxt = own_cur;
    int rbase = 0;   // (R * FPR) % ring
    int rbuf = 0;    // R % kRecRounds

    for (int R = 0; R < rounds; ++R) {
        // {job, frame slot of A, B exists, start of A} and the start of B (samples relative to the utterance) of the
        // pair lane 0..2 looks after
        int4 cur = make_int4(-1, 0, 0, 0);
        int curB = 0;
        if constexpr (EPIREC) {
            if (lane < kPairs) {
                cur = record(rbuf, 0);
                curB = record(rbuf, 1).w;
            }
        } else {
            cur = make_int4(own_cur.job, own_cur.fa, own_cur.has_b, own_cur.startA);
            curB = own_cur.startB;
        }
        const int rnext = rbuf + 1 == kRecRounds ? 0 : rbuf + 1;
        const int my_job = __shfl_sync(0xffffffffu, cur.x, q);
        const int my_hasb = __shfl_sync(0xffffffffu, cur.z, q);
        const unsigned live = __ballot_sync(0xffffffffu, cur.x >= 0);   // bit qq: pair qq exists
        const int rel0 = 2 * kPairs * warp;   // first frame of this warp's triple, relative to the round
        const int rb6 = ring_slot(rbase, rel0, P.ring);   // its ring slot; the ring is a multiple of 6, so the six frames are contiguous

        f2 xr[20], xi[20];   // packed pairs: pass 1 (column 2j, column 2j+1), pass 2 (row u, row 20-u)
        mbar_wait(bar, (uint32_t)(R & 1));
        // Default hop (160 samples = 8 column strides of 20): frame B's columns n1 = 0..11 are frame A's
        // columns 8..19, so a full pair needs 28 loads per lane instead of 40.
        const bool shared_cols = P.contig && !P.in_i16 && P.step == 160 && __all_sync(0xffffffffu, my_job >= 0 && my_hasb);
        if (fft_lane && shared_cols) {
            const float2 *w2 = scr_q + kWinOff + j;
#pragma unroll
            for (int n = 0; n < 28; ++n) {
                const float2 v = w2[10 * n];
                if (n < 20) xr[n] = v;
                if (n >= 8) xi[n - 8] = v;
            }
        } else if (fft_lane) {
            const bool a_live = my_job >= 0, b_live = my_job >= 0 && my_hasb;
            const int offB = P.contig ? P.step : kN;   // frame B inside the window, in samples
            if (!P.in_i16) {
                const float *wq = reinterpret_cast<const float *>(scr_q + kWinOff);
                const float *pa = a_live ? wq : sm.zeros;
                const float *pb = b_live ? wq + offB : sm.zeros;
                if (!b_live || (offB & 1) == 0) {
                    const float2 *pa2 = reinterpret_cast<const float2 *>(pa) + j;
                    const float2 *pb2 = reinterpret_cast<const float2 *>(pb) + j;
#pragma unroll
                    for (int n1 = 0; n1 < 20; ++n1) { xr[n1] = pa2[10 * n1]; xi[n1] = pb2[10 * n1]; }
                } else {   // odd hop: frame B is not 8-byte aligned inside the window
#pragma unroll
                    for (int n1 = 0; n1 < 20; ++n1) {
                        xr[n1] = reinterpret_cast<const float2 *>(pa)[10 * n1 + j];
                        xi[n1] = make_float2(pb[20 * n1 + 2 * j], pb[20 * n1 + 2 * j + 1]);
                    }
                }
            } else {
                // int16 PCM window: two samples per 32-bit load, normalised like Wave.GetFloatAtIdx
                constexpr float kInv = 1.0f / 32767.0f;
                const short *wq = reinterpret_cast<const short *>(scr_q + kWinOff);
                const short *pa = a_live ? wq : reinterpret_cast<const short *>(sm.zeros);
                const short *pb = b_live ? wq + offB : reinterpret_cast<const short *>(sm.zeros);
                if (!b_live || (offB & 1) == 0) {
                    const short2 *pa2 = reinterpret_cast<const short2 *>(pa) + j;
                    const short2 *pb2 = reinterpret_cast<const short2 *>(pb) + j;
#pragma unroll
                    for (int n1 = 0; n1 < 20; ++n1) {
                        const short2 va = pa2[10 * n1], vb = pb2[10 * n1];
                        xr[n1] = make_float2((float)va.x * kInv, (float)va.y * kInv);
                        xi[n1] = make_float2((float)vb.x * kInv, (float)vb.y * kInv);
                    }
                } else {
#pragma unroll
                    for (int n1 = 0; n1 < 20; ++n1) {
                        const short2 va = reinterpret_cast<const short2 *>(pa)[10 * n1 + j];
                        xr[n1] = make_float2((float)va.x * kInv, (float)va.y * kInv);
                        xi[n1] = make_float2((float)pb[20 * n1 + 2 * j] * kInv, (float)pb[20 * n1 + 2 * j + 1] * kInv);
                    }
                }
            }
        }
        // ---- frame levels.  In the reference every frame is transformed alone (dft/dft.go:42-59); here two frames
        // share a complex FFT and each picks up the other's float32 rounding noise.  Per frame pair (bit qq):
        //   zero_*   the frame is exactly zero (front padding, digital silence): its mel sums are forced to the exact
        //            zero the reference tests for (mel.go:135);
        //   nonf_*   the frame holds a NaN / Inf sample: its spectrum is non-finite, as in the reference;
        //   alone_*  the frame is much quieter than its partner (kAloneRatio) or its partner is non-finite: it is
        //            transformed again with a zero partner in a second pass of this round.
        // packed: bits 0-2 frame A zero, 3-5 B zero, 6-8 A non-finite, 9-11 B non-finite, 12-14 A alone, 15-17 B alone
        unsigned lv = 0u;
        {
            unsigned pk_a = 0u, pk_b = 0u;
            if (fft_lane) {
                pk_a = __float_as_uint(frame_peak(xr));
                pk_b = __float_as_uint(frame_peak(xi));
            }
#pragma unroll
            for (int qq = 0; qq < kPairs; ++qq) {
                const unsigned ma = __reduce_max_sync(0xffffffffu, (fft_lane && q == qq) ? pk_a : 0u);
                const unsigned mb = __reduce_max_sync(0xffffffffu, (fft_lane && q == qq) ? pk_b : 0u);
                const bool za = ma == 0u, zb = mb == 0u, na = ma >= 0x7f800000u, nb = mb >= 0x7f800000u;
                const float fa = __uint_as_float(ma), fb = __uint_as_float(mb);
                const bool qa = !za && !na && (nb || fa * kAloneRatio < fb);
                const bool qb = !zb && !nb && (na || fb * kAloneRatio < fa);
                lv |= ((za ? 1u : 0u) | (zb ? 8u : 0u) | (na ? 64u : 0u) | (nb ? 512u : 0u) | (qa ? 4096u : 0u) | (qb ? 32768u : 0u)) << qq;
            }
        }
        const unsigned alone = ((lv >> 12) | (lv >> 15)) & 7u;   // pairs with a frame to redo (at most one frame of a pair)
        const int nrep = alone ? 2 : 1;
#pragma unroll 1
        for (int rep = 0; rep < nrep; ++rep) {
            if (rep == 1) {
                // second pass (rare: level steps, non-finite samples): the quiet frame of every flagged pair, alone.
                // Its window is gone (the exchange rows overwrote it), so the samples come from global memory.
                const int jb_i = __shfl_sync(0xffffffffu, cur.x, q);
                const int sA = __shfl_sync(0xffffffffu, cur.w, q), sB = __shfl_sync(0xffffffffu, curB, q);
                const bool mine = fft_lane && ((alone >> q) & 1u);
#pragma unroll
                for (int n1 = 0; n1 < 20; ++n1) { xr[n1] = make_float2(0.f, 0.f); xi[n1] = make_float2(0.f, 0.f); }
                if (mine) {
                    const Job &jb = sm.jobs[jb_i];
                    const int s0 = (((lv >> (15 + q)) & 1u) ? sB : sA) + 2 * j;
#pragma unroll
                    for (int n1 = 0; n1 < 20; ++n1) {
                        const int a0 = s0 + 20 * n1, a1 = a0 + 1;
                        float x0 = 0.f, x1 = 0.f;
                        if (P.in_i16) {
                            const short *wv = static_cast<const short *>(P.wave) + jb.wave_off;
                            if (a0 >= 0 && a0 < jb.utt_len) x0 = (float)__ldg(wv + a0) * (1.0f / 32767.0f);
                            if (a1 >= 0 && a1 < jb.utt_len) x1 = (float)__ldg(wv + a1) * (1.0f / 32767.0f);
                        } else {
                            const float *wv = static_cast<const float *>(P.wave) + jb.wave_off;
                            if (a0 >= 0 && a0 < jb.utt_len) x0 = __ldg(wv + a0);
                            if (a1 >= 0 && a1 < jb.utt_len) x1 = __ldg(wv + a1);
                        }
                        xr[n1] = make_float2(x0, x1);
                    }
                }
            }
            __syncwarp();   // the window is in registers; its space is the upper exchange rows from here on
            if (fft_lane) {
                // pass 1: columns 2j, 2j+1: DFT-20 over n1 of z[20 n1 + c], twiddle, packed row pairs to the exchange rows
                dft20(xr, xi);
                pass1_store(xr, xi, exq, sm.tw2, j);
            }
            __syncwarp();
            // pass 2: the lane owns row pair p2p of pair q2, so that Z[k] and its mirror Z[N - k] meet in one thread
            if (fft_lane) pass2_load(xr, xi, ex2, p2p);
            __syncwarp();   // every exchange row has been read: the scratch becomes power buffer + next window
            if (rep == nrep - 1 && R + 1 < rounds) {
                if constexpr (EPIREC) stage(R + 1, rnext);
                else {
                    own_nxt = resolve_own(R + 1);
                    stage_own(own_nxt);
                }
            }
            if (fft_lane) {
                dft20(xr, xi);
                if (p2p != 0) pass2_power(xr, xi, scr_q2, p2p);
                else pass2_park(xr, xi, scr_q2);   // rows 0 and 10 pair with themselves: cooperative step below
            } else {
                // lanes 30 / 31: zero the pad slots (index 20 mod 21) and the tail the last filter's quads reach
#pragma unroll
                for (int qq = 0; qq < kPairs; ++qq) {
                    float2 *pq = scr_w + qq * P.ps;
                    const int t0 = (lane - 30) * 5;
#pragma unroll
                    for (int t = 0; t < 5; ++t) pq[kPPitch * (t0 + t) + 20] = make_float2(0.f, 0.f);
#pragma unroll
                    for (int t = 0; t < 4; ++t) pq[211 + 4 * (lane - 30) + t] = make_float2(0.f, 0.f);   // 211..218
                }
            }
            __syncwarp();
            // ---- the self-paired rows 0 and 10 (21 bins per pair), all lanes: item = (pair, n)
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int item = lane + 32 * it;
                if (item < 21 * kPairs) selfpair_item(scr_w + (item / 21) * P.ps, item % 21);
            }
            __syncwarp();
            // the ring slots of this round were last used two rounds ago: wait until that round is finished
            if (rep == 0) mbar_wait(&empty[R & 1], (uint32_t)(((R >> 1) & 1) ^ 1));
            // In the second pass the lone frame's power sits in the .x halves; `tgt` says which frame of the pair it is.
            // ---- low bins for Energy, and the raw power rows of the parity / inspection outputs
            if (P.energy_bins > 0 || P.rawpow) {
#pragma unroll 1
                for (int qq = 0; qq < kPairs; ++qq) {
                    if (!(live & (1u << qq))) break;
                    if (rep == 1 && !((alone >> qq) & 1u)) continue;
                    const int tgt = (lv >> (15 + qq)) & 1u;
                    const float2 *pq = scr_w + qq * P.ps;
                    float *lowA = sm.rlow + (rb6 + 2 * qq) * P.energy_bins;
                    float *lowB = lowA + P.energy_bins;
                    for (int k = lane; k < P.energy_bins; k += 32) {
                        const float2 pv = pq[k + k / 20];
                        if (rep == 0) { lowA[k] = pv.x; lowB[k] = pv.y; }
                        else (tgt ? lowB : lowA)[k] = pv.x;
                    }
                    if (P.rawpow) {
                        const int job = __shfl_sync(0xffffffffu, cur.x, qq), fa = __shfl_sync(0xffffffffu, cur.y, qq);
                        const int hb = __shfl_sync(0xffffffffu, cur.z, qq);
                        float *rowA = P.rawpow + (size_t)(sm.jobs[job].frame_base + fa) * kPowPitch;
                        for (int k = lane; k < kBins; k += 32) {
                            const float2 pv = pq[k + k / 20];
                            if (rep == 0) {
                                rowA[k] = pv.x;
                                if (hb) rowA[kPowPitch + k] = pv.y;
                            } else {
                                rowA[tgt * kPowPitch + k] = pv.x;
                            }
                        }
                    }
                }
            }
            // ---- mel filter bank on the raw power.  Each lane runs one (pair, filter) task per slot; taps are
            // zero padded to whole quads and every lane of a slot runs the slot's longest loop (shorter rows
            // read weight 0 against finite power values).  Without smoothing the log is taken here, once per
            // frame; with it the raw sums go to the ring (smoothing is linear: phase 2 applies it to the sums).
            for (int t = 0; t < P.mel_tasks; ++t) {
                const int4 td = sm.sched[t * 32 + lane];   // {tap offset, power offset, code, -}
                const int task = td.z;
                const int qq = task < 0 ? 0 : (task >> 16) & 0xff, m = task < 0 ? 0 : task & 0xffff;
                const int nit = __shfl_sync(0xffffffffu, task, 0) >> 24;   // every task of a slot carries the slot's longest loop
                const bool on = task >= 0 && (live & (1u << qq));
                const float4 *wp = reinterpret_cast<const float4 *>(sm.taps) + td.x + lane;   // [slot][quad][lane]
                const float4 *pp = reinterpret_cast<const float4 *>(scr_w + td.y);
                AUD_CHECK(P, td.y >= 0 && td.y + 4 * nit <= kPairs * P.ps && (td.y & 1) == 0, 10);
                AUD_CHECK(P, nit >= 1 && 4 * (td.x + 32 * nit) <= P.mel_taps_len, 11);
                // four independent accumulator pairs (one per tap of a quad) keep the FMA chains short.  The loads of
                // quad it + 1 are issued before the FMAs of quad it (software pipelining): the warp then waits for
                // shared memory once per slot instead of once per quad.
                float sa, sb;
                float4 w0 = wp[0], p0 = pp[0], p1 = pp[1];   // every slot has at least one quad
                if constexpr (MELPACK) {
                    // (frame A, frame B) ride one packed FMA per tap: the power buffer keeps the pair side by side and
                    // FFMA2 takes the tap weight as a scalar that it applies to both halves
                    f2 c0 = make_float2(0.f, 0.f), c1 = c0, c2 = c0, c3 = c0;
#pragma unroll 2
                    for (int it = 1; it < nit; ++it) {
                        const float4 wn = wp[32 * it], p0n = pp[2 * it], p1n = pp[2 * it + 1];
                        c0 = fma2(make_float2(p0.x, p0.y), bc2(w0.x), c0);
                        c1 = fma2(make_float2(p0.z, p0.w), bc2(w0.y), c1);
                        c2 = fma2(make_float2(p1.x, p1.y), bc2(w0.z), c2);
                        c3 = fma2(make_float2(p1.z, p1.w), bc2(w0.w), c3);
                        w0 = wn; p0 = p0n; p1 = p1n;
                    }
                    c0 = fma2(make_float2(p0.x, p0.y), bc2(w0.x), c0);
                    c1 = fma2(make_float2(p0.z, p0.w), bc2(w0.y), c1);
                    c2 = fma2(make_float2(p1.x, p1.y), bc2(w0.z), c2);
                    c3 = fma2(make_float2(p1.z, p1.w), bc2(w0.w), c3);
                    const f2 cs = add2(add2(c0, c1), add2(c2, c3));
                    sa = cs.x; sb = cs.y;
                } else {
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
#pragma unroll 2
                    for (int it = 1; it < nit; ++it) {
                        const float4 wn = wp[32 * it], p0n = pp[2 * it], p1n = pp[2 * it + 1];
                        a0 = fmaf(w0.x, p0.x, a0); b0 = fmaf(w0.x, p0.y, b0);
                        a1 = fmaf(w0.y, p0.z, a1); b1 = fmaf(w0.y, p0.w, b1);
                        a2 = fmaf(w0.z, p1.x, a2); b2 = fmaf(w0.z, p1.y, b2);
                        a3 = fmaf(w0.w, p1.z, a3); b3 = fmaf(w0.w, p1.w, b3);
                        w0 = wn; p0 = p0n; p1 = p1n;
                    }
                    a0 = fmaf(w0.x, p0.x, a0); b0 = fmaf(w0.x, p0.y, b0);
                    a1 = fmaf(w0.y, p0.z, a1); b1 = fmaf(w0.y, p0.w, b1);
                    a2 = fmaf(w0.z, p1.x, a2); b2 = fmaf(w0.z, p1.y, b2);
                    a3 = fmaf(w0.w, p1.z, a3); b3 = fmaf(w0.w, p1.w, b3);
                    sa = (a0 + a1) + (a2 + a3); sb = (b0 + b1) + (b2 + b3);
                }
                if (on) {
                    const int slA = rb6 + 2 * qq;
                    AUD_CHECK(P, slA >= 0 && slA + 1 < P.ring && m < P.n_mel, 12);
                    float *rowA = sm.rmel + slA * P.mel_pitch + m;
                    float *rowB = rowA + P.mel_pitch;
                    const int mir = P.ring * P.mel_pitch;   // the first kMirror rows are kept twice
                    if (rep == 0) {
                        if ((lv >> qq) & 1u) sa = 0.f;         // exactly-zero frame -> exactly-zero sums
                        if ((lv >> (3 + qq)) & 1u) sb = 0.f;
                        if (P.nosmooth) { sa = finish_mel(P, sa); sb = finish_mel(P, sb); }
                        *rowA = sa;
                        *rowB = sb;
                        if (slA < kMirror) { rowA[mir] = sa; rowB[mir] = sb; }   // slA is even, kMirror too
                    } else if ((alone >> qq) & 1u) {
                        if (P.nosmooth) sa = finish_mel(P, sa);
                        float *row = ((lv >> (15 + qq)) & 1u) ? rowB : rowA;
                        *row = sa;
                        if (slA < kMirror) row[mir] = sa;
                    }
                }
            }
            __syncwarp();
        }
        // the spare column of a frame's ring row says whether the frame was non-finite: without smoothing the
        // epilogue needs it to carry the reference's `PrevSmooth*Power[k] + CurSmooth*p` (dft.go:66-68; 0 * NaN = NaN)
        // to the later steps of the segment
        if (lane < 2 * kPairs) {
            const unsigned nf = lv >> (((lane & 1) ? 9 : 6) + (lane >> 1));
            const float fl = (nf & 1u) ? __int_as_float(0x7fc00000) : 0.f;
            float *cell = sm.rmel + (rb6 + lane) * P.mel_pitch + P.n_mel;
            *cell = fl;
            if (rb6 + lane < kMirror) cell[P.ring * P.mel_pitch] = fl;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[R & 1]);   // release: this warp's frames of round R are in the ring
        rbase += FPR;
        if (rbase >= P.ring) rbase -= P.ring;
        rbuf = rnext;
        own_cur = own_nxt;
    }
}

// gabor weights [nf][sy][sx] (agabor.ToTensor layout) -> shared memory, tap-major with the filters of a
// tap padded to whole groups of 8, so that one thread reads 8 filters' weights of a tap as two float4
__device__ __forceinline__ void load_gabor_weights(const KParams &P, float *gw, int tid, int nthreads) {
    if (P.gw_floats <= 0) return;
    const int nfp = ((P.g_nf + 7) >> 3) << 3, taps = P.g_sy * P.g_sx;
    for (int i = tid; i < P.gw_floats; i += nthreads) {
        const int tp = i / nfp, f = i - tp * nfp;
        gw[i] = f < P.g_nf ? P.gabor[(size_t)f * taps + tp] : 0.f;
    }
}

// ------------------------------------------------------------ tile stage
// Everything that follows the log-mel tile of a finished segment: cepstrum (mel.go:192-212), Energy -> c0
// (sndenv.go:368-372), deltas (sndenv.go:378-432), gabor (gabor.go:225-315) and the stores of those
// outputs.  `t` holds nd segments' tiles in shared memory, done[dd] = {output segment, valid steps, -, -};
// the et-th of ENT cooperating threads calls it, esync() is their barrier.  Shared by the fused kernel's
// epilogue warps and by the general-window-length path.
struct TileSet {
    float *mel;      // [nd][M][S] log-mel
    float *energy;   // [nd][S]
    float *mfcc;     // [nd][NC][S]
    float *d1, *d2;  // [nd][NC][S]
    float *gab;      // [nd][g_len]
};

template <typename Sync>
__device__ __forceinline__ void finish_tiles(const KParams &P, const TileSet &t, const float *dct_sm, const float *gw_sm,
                                             const int4 *done, int nd, int et, int ENT, bool store_mel, Sync esync) {
    const int S = P.S, M = P.n_mel, NC = P.n_coefs, MS = M * S;
    // Dense 4-D gabor output [PoolsY][PoolsX][2][8]: a position's sixteen results (eight filters on, eight off) are
    // sixteen consecutive floats and the positions cover the whole tensor, so every thread stores its results
    // straight to global memory as four 128-bit stores -- no output tile, no zero fill, no copy-out pass.
    // (decided on the host, gabor_direct() in aud_api.cu, which then also leaves the output tile out of shared memory)
    const bool g_direct = P.g_on && P.g_direct;
    if (P.g_on && !g_direct) {
        if (P.g_keep)   // rawOut cells the loops below do not reach stay as the caller left them
            for (int r = et; r < nd * P.g_len; r += ENT) t.gab[r] = P.o_gabor[(size_t)done[r / P.g_len].x * P.g_len + r % P.g_len];
        else
            for (int r = et; r < nd * P.g_len; r += ENT) t.gab[r] = 0.f;
    }
    esync();
    // smoothed log-mel leaves through the tile (coalesced); one warp per segment, no index arithmetic
    const int wv = et >> 5, ln = et & 31, nwv = ENT >> 5;
    if (store_mel)
        for (int dd = wv; dd < nd; dd += nwv) {
            float *gout = P.o_mel + (size_t)done[dd].x * MS;
            const float *src = t.mel + (size_t)dd * MS;
            for (int e = ln; e < MS; e += 32) gout[e] = src[e];
        }
    // (c) cepstrum: two threads per (segment, step) column -- one for each half of the coefficients, so that
    // all the cooperating warps have work -- keep 32 log-mel values of the column in registers and run
    // their DCT-I rows over them (the matrix rows come from shared memory as broadcast 128-bit loads)
    if (P.want_mfcc) {
        const float4 *dct4 = reinterpret_cast<const float4 *>(dct_sm);
        const int M4 = (M + 3) >> 2;   // dct rows are padded to whole float4s
        const int kh = (NC + 1) >> 1;  // coefficients [0, kh) and [kh, NC)
        for (int r = et; r < 2 * nd * S; r += ENT) {
            const int half = r & 1, c = r >> 1;
            const int dd = c / S, i = c - dd * S;
            const int nv = done[dd].y;
            const int k0 = half ? kh : 0, k1 = half ? NC : kh;
            const float *col = t.mel + (size_t)dd * MS + i;
            float *mf = t.mfcc + (size_t)dd * NC * S + i;
            if (i < nv) {
                for (int k = k0; k < k1; ++k) mf[k * S] = 0.f;
                for (int m0 = 0; m0 < M; m0 += 32) {
                    float x[32];
#pragma unroll
                    for (int u = 0; u < 32; ++u) x[u] = (m0 + u < M) ? col[(m0 + u) * S] : 0.f;
                    for (int k = k0; k < k1; ++k) {
                        const float4 *drow = dct4 + k * M4 + (m0 >> 2);
                        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (m0 + 4 * u < M) {
                                const float4 d = drow[u];
                                a0 = fmaf(d.x, x[4 * u], a0); a1 = fmaf(d.y, x[4 * u + 1], a1);
                                a2 = fmaf(d.z, x[4 * u + 2], a2); a3 = fmaf(d.w, x[4 * u + 3], a3);
                            }
                        }
                        mf[k * S] += (a0 + a1) + (a2 + a3);
                    }
                }
                if (!half) {
                    const float y0 = mf[0];
                    mf[0] = log1pf(y0 * y0);   // mel.go:203-204
                }
            } else {
                for (int k = k0; k < k1; ++k) mf[k * S] = 0.f;
            }
            if (!half && P.c0_energy) mf[0] = t.energy[dd * S + i];   // sndenv.go:368-372 (every step)
        }
    }
    // (e) gabor: one thread per (segment, position, group of 8 filters): strided valid correlation of the
    // filters with the segment's mel tile (gabor.go:264-313); a tap's 8 weights are two broadcast float4s
    if (P.g_on) {
        const int ngrp = (P.g_nf + 7) >> 3;
        const int nfp4 = ngrp * 2;   // float4s per tap in gw_sm
        const int per_seg = P.g_nt * P.g_nfy * ngrp;
        for (int r = et; r < nd * per_seg; r += ENT) {
            const int dd = r / per_seg;
            int rem = r - dd * per_seg;
            const int ti = rem / (P.g_nfy * ngrp);
            rem -= ti * P.g_nfy * ngrp;
            const int fi = rem / ngrp, f0 = (rem - fi * ngrp) * 8;
            const float *trow = t.mel + (size_t)dd * MS + (fi * P.g_sty) * S + ti * P.g_stx;
            const float4 *w = reinterpret_cast<const float4 *>(gw_sm) + (f0 >> 2);
            // eight filters as four packed pairs: a tap's weights come as two 128-bit broadcast loads whose register
            // pairs feed FFMA2 directly; the mel value is duplicated into a pair once per tap
            f2 a01 = make_float2(0.f, 0.f), a23 = a01, a45 = a01, a67 = a01;
            for (int ff = 0; ff < P.g_sy; ++ff, trow += S) {
#pragma unroll 3
                for (int ft = 0; ft < P.g_sx; ++ft, w += nfp4) {
                    float iv = trow[ft];
                    if (iv != iv) iv = 0.5f;   // gabor.go:283-285
                    const float4 w0 = w[0], w1 = w[1];
                    const f2 v2 = make_float2(iv, iv);
                    a01 = fma2(make_float2(w0.x, w0.y), v2, a01); a23 = fma2(make_float2(w0.z, w0.w), v2, a23);
                    a45 = fma2(make_float2(w1.x, w1.y), v2, a45); a67 = fma2(make_float2(w1.z, w1.w), v2, a67);
                }
            }
            const float a[8] = {a01.x, a01.y, a23.x, a23.y, a45.x, a45.y, a67.x, a67.y};
            if (g_direct) {
                float on[8], off[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float act = P.g_gain * fabsf(a[u]);
                    const bool pos = a[u] >= 0.f;   // gabor.go:296-311: sign routes to the on or the off slot, the other is 0
                    on[u] = pos ? act : 0.f;
                    off[u] = pos ? 0.f : act;
                }
                float4 *g4 = reinterpret_cast<float4 *>(P.o_gabor + (size_t)done[dd].x * P.g_len + fi * P.g_str0 + ti * P.g_str1);
                g4[0] = make_float4(on[0], on[1], on[2], on[3]);
                g4[1] = make_float4(on[4], on[5], on[6], on[7]);
                g4[2] = make_float4(off[0], off[1], off[2], off[3]);
                g4[3] = make_float4(off[4], off[5], off[6], off[7]);
                continue;
            }
            float *g = t.gab + (size_t)dd * P.g_len;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int flt = f0 + u;
                if (flt < P.g_nf) {
                    const float acc = a[u];
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
                    g[on_off] = pos ? act : 0.f;
                    g[off_off] = pos ? 0.f : act;
                }
            }
        }
    }
    esync();
    // (d) deltas and delta-deltas with the reference's accumulator quirk
    if (P.want_mfcc && P.do_deltas) {
        for (int pass = 0; pass < 2; ++pass) {
            const float *src = pass == 0 ? t.mfcc : t.d1;
            float *dst = pass == 0 ? t.d1 : t.d2;
            for (int r = et; r < nd * S; r += ENT) {
                const int dd = r / S, s = r - dd * S;
                float prv = 0.f, nx = 0.f;
                for (int k = 0; k < NC; ++k) {
                    const float *rowp = src + ((size_t)dd * NC + k) * S;
                    float nume = 0.f, dv = 0.f;
                    for (int n = 1; n <= 2; ++n) {
                        const int sp2 = s - n < 0 ? 0 : s - n;
                        const int sn = s + n > S - 1 ? S - 1 : s + n;
                        prv += rowp[sp2];
                        nx += rowp[sn];
                        nume += (float)n * (nx - prv);
                        dv = nume / (float)(2 * n * n);
                    }
                    dst[((size_t)dd * NC + k) * S + s] = dv;
                }
            }
            esync();
        }
    }
    // stores of the tile-resident outputs: one warp per segment
    for (int dd = wv; dd < nd; dd += nwv) {
        const size_t seg = (size_t)done[dd].x;
        if (P.want_mfcc) {
            const int CS = NC * S;
            const size_t o = seg * CS, ti = (size_t)dd * CS;
            if (P.o_mfcc)
                for (int i = ln; i < CS; i += 32) P.o_mfcc[o + i] = t.mfcc[ti + i];
            if (P.do_deltas && P.o_d1)
                for (int i = ln; i < CS; i += 32) P.o_d1[o + i] = t.d1[ti + i];
            if (P.do_deltas && P.o_d2)
                for (int i = ln; i < CS; i += 32) P.o_d2[o + i] = t.d2[ti + i];
        }
        if (P.g_on && P.o_gabor && !g_direct)
            for (int i = ln; i < P.g_len; i += 32) P.o_gabor[seg * P.g_len + i] = t.gab[(size_t)dd * P.g_len + i];
    }
    esync();   // the tiles are reused by the next batch / round
}

// ------------------------------------------------------------ epilogue warps
// Finish the segments whose last frame landed in round R: smoothing scan, logs, Energy, DCT, deltas,
// gabor, stores.  `et` / ENT: thread index / thread count among the epilogue warps.
template <int NWARPS, int NEPI, bool EPIREC>
__device__ __forceinline__ void epilogue_role(const KParams &P, const Smem &sm, int et, int lane, int njobs,
                                              int total_pairs, int rounds) {
    constexpr int FPR = 6 * NWARPS, ENT = NEPI * 32;
    const int S = P.S, M = P.n_mel, MS = M * S;
    uint64_t *full = &sm.mbar[NWARPS], *empty = &sm.mbar[NWARPS + 2];
    // No-smoothing gather, fast path: a lane owns float4 number lane + 32 u of a segment's [M][S] tile (the order of
    // the output tensor) and knows, once and for all, where its four elements sit relative to the ring row of the
    // segment's first frame: step * pitch + filter.  Whole, finite segments then leave as 128-bit stores with no
    // index arithmetic.  Everything else (tail segments, non-finite frames, tiles that do not fit the table, long
    // segments) walks the tile element by element.
    constexpr int kG4 = 4;
    const bool fast_tile = (MS & 3) == 0 && MS <= 128 * kG4 && S <= kMirror && (reinterpret_cast<uintptr_t>(P.o_mel) & 15) == 0;
    int goff[kG4][4];
#pragma unroll
    for (int u = 0; u < kG4; ++u)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int e = 4 * (lane + 32 * u) + c;
            const int mm = e / S, ss = e - mm * S;
            goff[u][c] = (fast_tile && e < MS) ? ss * P.mel_pitch + mm : 0;
        }
    // walk of a lane over one [M][S] tile in steps of 32 elements (general path)
    constexpr int kGatherU = 7;
    const int ewarp = et >> 5;
    const int lane_m0 = lane / S, lane_i0 = lane - lane_m0 * S;
    const int lane_dm = 32 / S, lane_di = 32 - lane_dm * S;
    float *t_mel = sm.tiles;                       // [tile_cap][M][S]
    float *t_energy = sm.tiles + P.t_off[0];       // [tile_cap][S]
    float *t_mfcc = sm.tiles + P.t_off[1];         // [tile_cap][NC][S]
    float *t_d1 = sm.tiles + P.t_off[2];
    float *t_d2 = sm.tiles + P.t_off[3];
    float *t_gab = sm.tiles + P.t_off[4];          // [tile_cap][g_len]
    auto esync = [&]() {
        if (NEPI == 1) __syncwarp();
        else named_bar_sync(1, ENT);
    };
    int rbase = 0;
    int rbuf4 = 4;   // (R + 4) % kRecRounds: where the records written in round R go
    int jlo = 0;   // first job that may still complete segments (warp 0 of the role, lane 0)

    for (int R = 0; R < rounds; ++R) {
        const int F0 = R * FPR, F1 = F0 + FPR;
        // ---- the list of segments this round completes: {global output segment, valid steps, first ring
        // frame, job}; lane 0 of the role's first warp finds the per-job ranges, its lanes expand them
        if (et < 32) {
            int *rng = sm.dmeta + 2;   // [kMaxRanges][4]: job, first segment, segments before, count
            if (lane == 0) {
                int n = 0, total = 0;
                while (jlo < njobs && 2 * (sm.jobs[jlo].pair_base + ((sm.jobs[jlo].nframes + 1) >> 1)) <= F0) ++jlo;
                for (int jj = jlo; jj < njobs && n < kMaxRanges; ++jj) {
                    const Job &jb = sm.jobs[jj];
                    const int sb = 2 * jb.pair_base;
                    if (sb >= F1) break;
                    // segment c ends at stream frame sb + c*seg_adv + S - 1
                    int lo = floordiv32(F0 - sb - S + P.seg_adv, P.seg_adv);
                    int hi = floordiv32(F1 - sb - S, P.seg_adv);
                    if (lo < 0) lo = 0;
                    if (hi > jb.nseg - 1) hi = jb.nseg - 1;
                    if (total + (hi - lo + 1) > kMaxDone) hi = lo + (kMaxDone - total) - 1;   // cannot happen: <= 1 per frame
                    if (hi >= lo) {
                        rng[4 * n + 0] = jj; rng[4 * n + 1] = lo; rng[4 * n + 2] = total; rng[4 * n + 3] = hi - lo + 1;
                        total += hi - lo + 1;
                        ++n;
                    }
                }
                sm.dmeta[0] = n;
                sm.dmeta[1] = total;
            }
            __syncwarp();
            const int n = sm.dmeta[0], total = sm.dmeta[1];
            for (int d = lane; d < total; d += 32) {
                int rr = 0;
                while (rr + 1 < n && rng[4 * (rr + 1) + 2] <= d) ++rr;
                const int jj = rng[4 * rr], c = rng[4 * rr + 1] + (d - rng[4 * rr + 2]);
                const Job &jb = sm.jobs[jj];
                sm.done[d] = make_int4((int)(jb.out_seg + c),
                                       valid_steps(jb.utt_len, P.add, P.stride, P.step, P.border, S, jb.seg0 + c),
                                       2 * jb.pair_base + c * P.seg_adv, jj);
            }
        }
        esync();
        const int ndone = sm.dmeta[1];
        AUD_CHECK(P, ndone >= 0 && ndone <= kMaxDone && sm.dmeta[0] <= kMaxRanges, 21);
        mbar_wait(&full[R & 1], (uint32_t)((R >> 1) & 1));   // every FFT warp has delivered round R

        for (int d0 = 0; d0 < ndone; d0 += P.tile_cap) {
            const int nd = min(P.tile_cap, ndone - d0);
            // (a) log-mel tiles [M][S] of the finished segments
            if (P.nosmooth) {
                // the ring already holds ln(mel) per frame: gather it.  One warp per segment, lanes walking the
                // [M][S] tile linearly (the order of the output tensor: 128-byte coalesced stores), seven
                // independent elements in flight per lane.
                for (int dd = ewarp; dd < nd; dd += NEPI) {
                    const int4 en = sm.done[d0 + dd];   // {out segment, valid steps, first ring frame, job}
                    float *gout = P.o_mel + (size_t)en.x * MS;
                    float *tout = t_mel + dd * MS;
                    int b0 = rbase + (en.z - F0);       // ring slot of the segment's first frame, in [0, ring)
                    if (b0 < 0) b0 += P.ring;
                    if (b0 >= P.ring) b0 -= P.ring;
                    AUD_CHECK(P, b0 >= 0 && b0 < P.ring && dd < P.tile_cap && en.y >= 0 && en.y <= S, 20);
                    // The reference smooths from step 1 on with `PrevSmooth*Power[k] + CurSmooth*p` even when
                    // PrevSmooth is 0 (dft.go:66-68), and 0 * NaN is NaN: the steps after a non-finite frame are
                    // NaN in every filter until the segment ends.  `first_bad`: first such frame of this segment.
                    int first_bad = S;
                    for (int i0 = 0; i0 < en.y; i0 += 32) {
                        const int i = i0 + lane;
                        bool bad = false;
                        if (i < en.y) {
                            int sl = b0 + i;
                            if (sl >= P.ring) sl -= P.ring;
                            const float flag = sm.rmel[sl * P.mel_pitch + M];
                            bad = flag != flag;
                        }
                        const unsigned ball = __ballot_sync(0xffffffffu, bad);
                        if (ball) { first_bad = i0 + __ffs(ball) - 1; break; }
                    }
                    if (fast_tile && en.y == S && first_bad >= S) {
                        const float *base = sm.rmel + b0 * P.mel_pitch;   // rows b0 .. b0 + S - 1: no wrap (mirror rows)
#pragma unroll
                        for (int u = 0; u < kG4; ++u) {
                            const int f4 = lane + 32 * u;
                            if (4 * f4 < MS) {
                                const float4 v = make_float4(base[goff[u][0]], base[goff[u][1]], base[goff[u][2]], base[goff[u][3]]);
                                if (P.o_mel) reinterpret_cast<float4 *>(gout)[f4] = v;
                                if (P.need_tiles) reinterpret_cast<float4 *>(tout)[f4] = v;
                            }
                        }
                        continue;
                    }
                    int m = lane_m0, i = lane_i0;
                    for (int e0 = lane; e0 < MS; e0 += 32 * kGatherU) {
                        float v[kGatherU];
#pragma unroll
                        for (int u = 0; u < kGatherU; ++u) {
                            v[u] = 0.f;
                            if (e0 + 32 * u < MS && i < en.y) {
                                int sl = b0 + i;
                                if (sl >= P.ring) sl -= P.ring;
                                v[u] = sm.rmel[sl * P.mel_pitch + m];
                                if (i > first_bad) v[u] = __int_as_float(0x7fc00000);
                            }
                            m += lane_dm; i += lane_di;
                            if (i >= S) { i -= S; ++m; }
                        }
#pragma unroll
                        for (int u = 0; u < kGatherU; ++u) {
                            if (e0 + 32 * u < MS) {
                                if (P.o_mel) gout[e0 + 32 * u] = v[u];
                                if (P.need_tiles) tout[e0 + 32 * u] = v[u];
                            }
                        }
                    }
                }
            } else if (S <= 32) {
                // Prev/Cur smoothing, short segments: one thread per (segment, filter) row runs the recurrence
                // P_s = Prev * P_{s-1} + Cur * p_s (restarting at step 0) over the steps; rows are independent
                for (int row = et; row < nd * M; row += ENT) {
                    const int dd = row / M, m = row - dd * M;
                    const int4 en = sm.done[d0 + dd];
                    int b0 = rbase + (en.z - F0);
                    if (b0 < 0) b0 += P.ring;
                    float *trow = t_mel + dd * MS + m * S;
                    float y = 0.f;
                    for (int i = 0; i < S; ++i) {
                        float val = 0.f;
                        if (i < en.y) {
                            int sl = b0 + i;
                            if (sl >= P.ring) sl -= P.ring;
                            const float x = sm.rmel[sl * P.mel_pitch + m];
                            y = (i == 0) ? x : fmaf(P.prev, y, P.cur * x);
                            val = finish_mel(P, y);
                        }
                        trow[i] = val;
                    }
                }
            } else {
                // Prev/Cur smoothing, long segments: the same recurrence as a Kogge-Stone parallel scan over
                // the steps (lanes = steps, chunks of 32 chained through a carry)
                for (int row = ewarp; row < nd * M; row += NEPI) {
                    const int dd = row / M, m = row - dd * M;
                    const int4 en = sm.done[d0 + dd];
                    const int relf = en.z - F0;
                    float carry = 0.f;
                    for (int i0 = 0; i0 < S; i0 += 32) {
                        const int i = i0 + lane;
                        float x = 0.f;
                        if (i < en.y) x = sm.rmel[ring_slot(rbase, relf + i, P.ring) * P.mel_pitch + m];
                        float y = (i == 0) ? x : P.cur * x;
                        float pwr = P.prev;
#pragma unroll
                        for (int dlt = 1; dlt < 32; dlt <<= 1) {
                            const float up = __shfl_up_sync(0xffffffffu, y, dlt);
                            if (lane >= dlt) y = fmaf(pwr, up, y);
                            pwr *= pwr;
                        }
                        if (i0 > 0) y = fmaf(ipowf(P.prev, lane + 1), carry, y);
                        carry = __shfl_sync(0xffffffffu, y, 31);
                        if (i < S) t_mel[dd * MS + m * S + i] = (i < en.y) ? finish_mel(P, y) : 0.f;
                    }
                }
            }
            // (b) Energy[s] = sum over steps f of LogPowerSegment.Values[s*S + f]  (bin s: transposed quirk);
            // one thread per (segment, bin) row
            if (P.energy_bins > 0) {
                for (int row = et; row < nd * S; row += ENT) {
                    const int dd = row / S, sb = row - dd * S;
                    const int4 en = sm.done[d0 + dd];
                    int b0 = rbase + (en.z - F0);
                    if (b0 < 0) b0 += P.ring;
                    float y = 0.f, esum = 0.f;
                    if (P.comp_log_pow && sb < P.energy_bins) {   // rows past the spectrum exist only for per-step callers (S > bins)
                        for (int i = 0; i < en.y; ++i) {
                            int sl = b0 + i;
                            if (sl >= P.ring) sl -= P.ring;
                            const float x = sm.rlow[sl * P.energy_bins + sb];
                            y = (i == 0) ? x : fmaf(P.prev, y, P.cur * x);
                            const float qv = y + P.log_off;
                            // ln(y + LogOffSet): absolute error of __logf here (~1e-7) is far inside the 1e-4 bound
                            esum += (qv == 0.f) ? P.log_min : __logf(qv);
                        }
                    }
                    if (P.o_energy) P.o_energy[(size_t)en.x * S + sb] = esum;
                    if (P.need_tiles) t_energy[dd * S + sb] = esum;
                }
            }
            if (P.need_tiles) {
                const TileSet ts{t_mel, t_energy, t_mfcc, t_d1, t_d2, t_gab};
                finish_tiles(P, ts, sm.dct, sm.gw, sm.done + d0, nd, et, ENT, !P.nosmooth && P.o_mel != nullptr, esync);
            }
        }
        // frame-pair records four rounds ahead: an FFT warp reads round X's records while it works on round
        // X - 1, by which time it has waited for this arrive of round X - 4
        if constexpr (EPIREC) {
            if (R + 4 < rounds) make_round_records(P, sm, NWARPS, njobs, total_pairs, R + 4, rbuf4, et, ENT);
            rbuf4 = rbuf4 + 1 == kRecRounds ? 0 : rbuf4 + 1;
        }
        esync();   // everyone is done reading the ring (and the done list) for round R
        if (lane == 0) mbar_arrive(&empty[R & 1]);
        rbase += FPR;
        if (rbase >= P.ring) rbase -= P.ring;
    }
}

// ------------------------------------------------------------ fused kernel
// NWARPS FFT warps + NEPI epilogue warps, synchronised through two pairs of mbarriers
// (full[R & 1]: round R is in the ring; empty[R & 1]: round R has been consumed).  No CTA-wide
// barrier after the set-up, so the warps drift apart and keep the FP32 and shared-memory pipes busy
// at the same time.
template <int NWARPS, int NEPI, bool EPIREC>
__global__ void __launch_bounds__((NWARPS + NEPI) * 32, 1) fused_features_kernel(const __grid_constant__ KParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NT = (NWARPS + NEPI) * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Smem sm = carve_smem(smem_raw, P, NWARPS);

    // ---- one-time setup: tables, this CTA's jobs, barriers
    const int2 jr = P.cta_jobs[blockIdx.x];
    const int njobs = jr.y - jr.x;
    for (int i = tid; i < kN / 2; i += NT) sm.tw2[i] = P.tw2[i];
    for (int i = tid; i < kN + 4; i += NT) sm.zeros[i] = 0.f;
    for (int i = tid; i < P.mel_taps_len; i += NT) sm.taps[i] = P.mel_taps[i];
    for (int i = tid; i < P.n_mel; i += NT) { sm.mstart[i] = P.mel_start[i]; sm.mquads[i] = P.mel_quads[i]; }
    for (int i = tid; i < P.mel_tasks * 32; i += NT) sm.sched[i] = P.mel_sched[i];
    for (int i = tid; i < njobs; i += NT) sm.jobs[i] = P.jobs[jr.x + i];
    {
        const int M4 = ((P.n_mel + 3) >> 2) << 2;
        for (int i = tid; i < P.dct_floats; i += NT) {
            const int k = i / M4, m = i - k * M4;
            sm.dct[i] = m < P.n_mel ? P.dct[k * P.n_mel + m] : 0.f;
        }
    }
    load_gabor_weights(P, sm.gw, tid, NT);
    if (tid < NWARPS) mbar_init(&sm.mbar[tid], 1);
    if (tid == NWARPS || tid == NWARPS + 1) mbar_init(&sm.mbar[tid], NWARPS);   // full[2]
    if (tid == NWARPS + 2 || tid == NWARPS + 3) mbar_init(&sm.mbar[tid], NEPI); // empty[2]
    if (tid < NWARPS + 4) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#ifdef AUD_DEBUG_CHECKS
    if (tid == 0) canaries(P, sm, NWARPS, 0);
    AUD_CHECK(P, njobs >= 0 && njobs <= kMaxJobs, 1);
#endif
    __syncthreads();
    if (njobs == 0) return;

    const int total_pairs = sm.jobs[njobs - 1].pair_base + ((sm.jobs[njobs - 1].nframes + 1) >> 1);
    const int rounds = (total_pairs + NWARPS * kPairs - 1) / (NWARPS * kPairs);
    // records of the first four rounds (later ones come from the epilogue warps, four rounds ahead)
    if constexpr (EPIREC)
        for (int r = 0; r < 4 && r < rounds; ++r) make_round_records(P, sm, NWARPS, njobs, total_pairs, r, r, tid, NT);
    __syncthreads();
    // Register split (setmaxnreg works on aligned groups of four warps): with 12 FFT + 4 epilogue warps the three FFT
    // warpgroups take more than the 128 registers per thread the launch grants -- the packed DFT-20s keep 80 registers of
    // data live, and at 128 the loop state around them was spilled or recomputed every round -- and the epilogue
    // warpgroup gives back: 144 / 80, and 384 * 144 + 128 * 80 = 65536, the whole register file.  Only for plain log-mel
    // launches (EPIREC), whose epilogue is light: with gabor or MFCC tiles the epilogue warps are the critical path and
    // lose more from a smaller budget than the FFT warps gain (measured: 136 / 104 costs the gabor launch 2 %, 144 / 80 6 %).
    constexpr bool kSplitRegs = (NWARPS == 12 && NEPI == 4 && EPIREC);
    if (warp < NWARPS) {
        if constexpr (kSplitRegs) asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
        // the mel stage's packed-FMA form wins where measured (plain log-mel, MFCC launches), the scalar form for gabor launches
        fft_role<NWARPS, EPIREC, (EPIREC || NEPI == 6)>(P, sm, warp, lane, njobs, total_pairs, rounds);
    } else {
        if constexpr (kSplitRegs) asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
        epilogue_role<NWARPS, NEPI, EPIREC>(P, sm, tid - NWARPS * 32, lane, njobs, total_pairs, rounds);
    }
#ifdef AUD_DEBUG_CHECKS
    __syncthreads();
    if (tid == 0) canaries(P, sm, NWARPS, 1);
#endif
}

}  // namespace aud
