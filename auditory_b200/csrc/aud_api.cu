// aud_api.cu -- host side of the C-ABI (include/auditory_b200.h): parameter
// validation, table upload, chunk planning, launches and transfers.  There is
// deliberately no CPU compute path here: without a working CUDA device every
// processing entry point fails with AUD_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "auditory_b200.h"
#include "aud_internal.h"
#include "aud_kernels.cuh"
#include "aud_generic.cuh"
#include "aud_dft_tc.cuh"
#include "aud_launch.h"

namespace aud {

static thread_local std::string g_last_error;

int32_t fail(int32_t code, const char *msg) {
    g_last_error = msg ? msg : "";
    return code;
}
int32_t failf(int32_t code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define AUD_CUDA(call)                                                                               \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return aud::failf(AUD_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),  \
                              __FILE__, __LINE__);                                                   \
    } while (0)

// grow-only device buffer
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Staging of ordinary (pageable) caller memory: two rotating page-locked bounce buffers per direction and a few
// copy threads.  A chunk is copied caller -> bounce by the threads (memcpy, several GB/s each) and bounce -> GPU by
// the copy engine, so the host copy of chunk k + 1 overlaps the DMA of chunk k; the other direction mirrors it.
struct Stager {
    static constexpr size_t kChunk = 8u << 20;
    char *in[2] = {nullptr, nullptr}, *out[2] = {nullptr, nullptr};
    cudaEvent_t in_free[2] = {}, out_ready[2] = {};
    unsigned in_k = 0, out_k = 0;
    struct Slice { char *dst; const char *src; size_t n; };
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::vector<Slice> slices;
    size_t next = 0, finished = 0;
    bool stop = false;

    cudaError_t init() {
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
            e = cudaHostAlloc((void **)&in[i], kChunk, cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaHostAlloc((void **)&out[i], kChunk, cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&in_free[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&out_ready[i], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) return e;
        const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
        const int n = threads_wanted > 0 ? threads_wanted - 1 : (int)std::min(3u, hw / 2);   // plus the calling thread
        for (int i = 0; i < n; ++i) workers.emplace_back([this]() { work(); });
        return cudaSuccess;
    }
    int threads_wanted = 0;   // option "copy_threads": copy threads including the caller, 0 = default
    void work() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_work.wait(lk, [this]() { return stop || next < slices.size(); });
            if (stop) return;
            const Slice s = slices[next++];
            lk.unlock();
            std::memcpy(s.dst, s.src, s.n);
            lk.lock();
            if (++finished == slices.size()) cv_done.notify_all();
        }
    }
    // memcpy split over the copy threads and the caller; returns when every byte has been copied
    void copy(void *dst, const void *src, size_t n) {
        const size_t parts = workers.size() + 1, per = (n / parts + 4095) & ~(size_t)4095;
        if (n < (1u << 20) || workers.empty()) { std::memcpy(dst, src, n); return; }
        std::unique_lock<std::mutex> lk(mu);
        slices.clear();
        next = finished = 0;
        for (size_t a = per; a < n; a += per) slices.push_back({(char *)dst + a, (const char *)src + a, std::min(per, n - a)});
        lk.unlock();
        cv_work.notify_all();
        std::memcpy(dst, src, std::min(per, n));
        lk.lock();
        cv_done.wait(lk, [this]() { return finished == slices.size(); });
        slices.clear();
        next = finished = 0;
    }
    // caller memory -> device, through the bounce buffers, on stream `st`
    cudaError_t to_device(void *dev, const void *host, size_t bytes, cudaStream_t st) {
        for (size_t a = 0; a < bytes; a += kChunk) {
            const size_t n = std::min(kChunk, bytes - a);
            const int s = in_k++ & 1;
            cudaError_t e = cudaEventSynchronize(in_free[s]);   // the DMA that last read this bounce buffer is done
            if (e != cudaSuccess) return e;
            copy(in[s], (const char *)host + a, n);
            e = cudaMemcpyAsync((char *)dev + a, in[s], n, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaEventRecord(in_free[s], st);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    // device -> caller memory; the DMA of chunk k + 1 runs while the threads copy chunk k out of its bounce buffer
    cudaError_t to_host(void *host, const void *dev, size_t bytes, cudaStream_t st) {
        if (bytes == 0) return cudaSuccess;
        auto issue = [&](size_t a, int &slot) {
            slot = (int)(out_k++ & 1);
            cudaError_t e = cudaMemcpyAsync(out[slot], (const char *)dev + a, std::min(kChunk, bytes - a), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaEventRecord(out_ready[slot], st);
            return e;
        };
        int cur = 0, nxt = 0;
        cudaError_t e = issue(0, cur);
        for (size_t a = 0; a < bytes && e == cudaSuccess; a += kChunk) {
            if (a + kChunk < bytes) e = issue(a + kChunk, nxt);
            if (e != cudaSuccess) break;
            e = cudaEventSynchronize(out_ready[cur]);
            if (e != cudaSuccess) break;
            copy((char *)host + a, out[cur], std::min(kChunk, bytes - a));
            cur = nxt;
        }
        return e;
    }
    ~Stager() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_work.notify_all();
        for (auto &t : workers) t.join();
        for (int i = 0; i < 2; ++i) {
            if (in[i]) cudaFreeHost(in[i]);
            if (out[i]) cudaFreeHost(out[i]);
            if (in_free[i]) cudaEventDestroy(in_free[i]);
            if (out_ready[i]) cudaEventDestroy(out_ready[i]);
        }
    }
};

// A launch plan: the utterances split into jobs and the jobs dealt to the persistent CTAs.
struct Plan {
    std::vector<int64_t> off;   // utterance offsets RELATIVE to the batch's first utterance
    std::vector<int32_t> len;
    uint64_t hash = 0;          // of (relative offsets, lengths): cache look-ups compare arrays only on a hash hit
    int key = -1;
    uint64_t used = 0;
    std::vector<Job> jobs;
    std::vector<int2> cta_jobs;
    int64_t total_segs = 0, total_frames = 0;
    bool uploaded = false;
    DevBuf d_jobs, d_cta_jobs;
    ~Plan() { d_jobs.release(); d_cta_jobs.release(); }
};

}  // namespace aud

struct aud_handle {
    aud_params p{};
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    // derived
    int bins = 0;
    int dedupe = 0, seg_adv = 0;
    int energy_bins = 0;
    int64_t gabor_len = 0;
    int g_on = 0, g_nt = 0, g_nfy = 0, g_tmaxstrides = 1;
    // tuning (aud_set_option)
    int opt_job_segs = 0;         // segments per job, 0 = auto
    int opt_warps = 0;            // warps per CTA, 0 = largest that fits
    int opt_ctas = 0;             // CTAs in the persistent grid, 0 = one per SM
    int opt_epi = 0;              // epilogue warps (1, 2 or 4), 0 = auto
    int opt_groups = 0;           // utterance groups of the host-path pipeline, 0 = auto
    int opt_pin = 0;              // pageable caller buffers: 0 = auto (staged through pinned bounce buffers), 1 = page-lock for the call, 2 = leave to the driver
    // device tables
    aud::DevBuf d_tw, d_mel_start, d_mel_width, d_mel_taps, d_mel_sched, d_dct, d_gabor;
    int mel_taps_len = 0, mel_tasks = 0;
    int ps = 0, contig = 0, win_len = 0;   // pair-scratch geometry
    // general window lengths (WinSamples != 400): folded-DFT tables and plain per-filter taps
    int fused = 1;
    int g_pitch = 0, g_wpitch = 0, g_need = 0;   // g_need: power bins the mel bank reads (general route)
    aud::DevBuf d_cos, d_sin, d_gmel_lo, d_gmel_n, d_gmel_w;
    aud::DevBuf d_tc_scale;       // tensor-core route: per-frame operand scales, frame / segment -> job tables, block maxima
    aud::DevBuf d_tc_tab;         // tensor-core route: split TF32 cos / sin table blocks (aud_dft_tc.cuh)
    int tc_kb = 0, tc_nt = 0, tc_tn = 0;
    int opt_copy_threads = 0;     // host copy threads of the staged path (0 = default), takes effect before the first staged call
    int opt_dft_tc = 1;           // general route: 1 = tcgen05 folded DFT, 0 = FP32 SIMT folded DFT
    // plan cache: one entry per (batch geometry, launch shape), least recently used first out
    std::vector<aud::Plan *> plans;
    uint64_t plan_clock = 0;
    aud::DevBuf d_rawpow;
    aud::DevBuf d_dbg;            // debug builds: the kernel's failed-check code
    // host-path pipeline
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_in[8] = {}, ev_done[8] = {};
    // host-path buffers
    aud::DevBuf d_wave, d_out[8];
    cudaStream_t stream = nullptr;
    aud::Stager *stager = nullptr;   // bounce buffers + copy threads for pageable caller memory (created on first use)
    int64_t launches = 0;
};

namespace aud {

static int64_t seg_count(const aud_params &p, int32_t n) {
    // SndEnv.Init (sndenv.go:263-265), Channels()==1; Go integer division truncates toward zero
    const int64_t siglen = (int64_t)n - p.segment_samples;
    const int64_t cnt = siglen / p.stride_samples + 1;
    return cnt < 0 ? 0 : cnt;
}

struct Launch {
    int warps, ps, win_len, contig, ring, tile_cap, need_tiles, per_round;
    int t_off[5];
    size_t tile_floats, dct_floats, gw_floats, smem;
};

struct Needs {   // what this call asks of the epilogue
    bool tiles, energy, mfcc, deltas, gabor;
    bool gabor_direct;   // gabor results leave as 128-bit global stores: no output tile in shared memory
};

// Dense 4-D gabor output [PoolsY][PoolsX][2][8] whose cells are all written by the convolution's positions, and a
// 16-byte aligned output: every thread then stores its position's sixteen results (8 on, 8 off) straight to global memory.
static bool gabor_direct(const aud_handle *h, const float *o_gabor) {
    const aud_params &p = h->p;
    return h->g_on && o_gabor && p.gabor_out_dims == 4 && p.gabor_nf == 8 && p.gabor_shape[3] == 8 && p.gabor_shape[2] == 2 &&
           p.gabor_shape[1] == h->g_nt && h->gabor_len == (int64_t)16 * h->g_nt * h->g_nfy &&
           (reinterpret_cast<uintptr_t>(o_gabor) & 15) == 0 && (h->gabor_len % 4) == 0;
}

static Launch pick_launch(const aud_handle *h, int warps, const Needs &nd, int energy_bins, bool full_cap, int rec_rounds) {
    const aud_params &p = h->p;
    Launch L{};
    L.warps = warps;
    L.contig = h->contig;
    L.win_len = h->win_len;
    L.ps = h->ps;
    const int fpr = 6 * warps;
    L.ring = (2 * fpr + p.segment_steps + 1 + 5) / 6 * 6;   // two rounds of frames + the reach of a finishing segment;
                                                            // a multiple of 6: a warp's six frames never straddle the wrap
    L.need_tiles = nd.tiles ? 1 : 0;
    L.tile_cap = kMaxDone;
    const size_t base = fused_smem_bytes(warps, L.ps, h->mel_taps_len, p.n_mel, h->mel_tasks, L.ring, energy_bins, 0, rec_rounds);
    L.smem = base;
    if (nd.tiles) {
        const size_t S = p.segment_steps, MS = (size_t)p.n_mel * S, CS = (size_t)p.n_coefs * S;
        const size_t per_seg = MS + (nd.energy ? S : 0) + (nd.mfcc ? CS : 0) + (nd.mfcc && nd.deltas ? 2 * CS : 0) +
                               ((nd.gabor && !nd.gabor_direct) ? (size_t)h->gabor_len : 0);
        // segments a round can finish: one per seg_adv frames, plus one per job boundary inside the round
        const int per_round = std::min(kMaxDone, fpr / std::max(1, h->seg_adv) + 3);
        const size_t gwf = nd.gabor ? (size_t)p.gabor_size_x * p.gabor_size_y * ((p.gabor_nf + 7) / 8 * 8) : 0;
        const size_t dctf = (nd.mfcc ? (size_t)p.n_coefs * (((size_t)p.n_mel + 3) / 4 * 4) + 4 : 0) + gwf;
        const size_t avail0 = base < (size_t)h->max_smem_optin ? ((size_t)h->max_smem_optin - base) / 4 : 0;
        const size_t avail = avail0 > dctf ? avail0 - dctf : 0;
        int cap = (int)std::min<size_t>(per_round, avail / per_seg);
        if (full_cap && 2 * cap < per_round) cap = 0;   // first choice: tiles for at least half of what a round can finish
        L.per_round = per_round;
        L.tile_cap = cap;
        size_t off = (size_t)cap * MS;
        L.t_off[0] = (int)off; off += nd.energy ? (size_t)cap * S : 0;
        L.t_off[1] = (int)off; off += nd.mfcc ? (size_t)cap * CS : 0;
        L.t_off[2] = (int)off; off += (nd.mfcc && nd.deltas) ? (size_t)cap * CS : 0;
        L.t_off[3] = (int)off; off += (nd.mfcc && nd.deltas) ? (size_t)cap * CS : 0;
        L.t_off[4] = (int)off; off += (nd.gabor && !nd.gabor_direct) ? (size_t)cap * h->gabor_len : 0;
        L.tile_floats = off;
        L.dct_floats = nd.mfcc ? (size_t)p.n_coefs * (((size_t)p.n_mel + 3) / 4 * 4) : 0;
        L.smem = fused_smem_bytes(warps, L.ps, h->mel_taps_len, p.n_mel, h->mel_tasks, L.ring, energy_bins, ((off + 3) & ~(size_t)3) + L.dct_floats + gwf, rec_rounds);
        L.gw_floats = gwf;
    }
    return L;
}

// The launch shapes that are compiled in (aud_launch.h): (FFT warps, epilogue warps, records-by-epilogue).
static cudaError_t launch_shape(int nw, int ne, int er, const KParams &kp, int grid, size_t smem, cudaStream_t st) {
#define AUD_TRY_SHAPE(NW, NE, ER) \
    if (nw == NW && ne == NE && er == ER) return launch_fused_##NW##_##NE##_##ER(kp, grid, smem, st);
    AUD_FUSED_SHAPES(AUD_TRY_SHAPE)
#undef AUD_TRY_SHAPE
    return cudaErrorInvalidConfiguration;
}
static bool have_shape(int nw, int ne, int er) {
#define AUD_HAS_SHAPE(NW, NE, ER) \
    if (nw == NW && ne == NE && er == ER) return true;
    AUD_FUSED_SHAPES(AUD_HAS_SHAPE)
#undef AUD_HAS_SHAPE
    return false;
}

constexpr int32_t kPlanTooMany = -100;   // internal: the caller splits the batch and retries

// Split every utterance into jobs of about `job_segs` segments and deal the jobs, in order, to
// `n_cta` persistent CTAs so that each gets about the same number of frame pairs.
static int32_t build_plan(aud_handle *h, const aud_batch *b, int n_cta, int job_segs, Plan **out) {
    const int key = (n_cta + 1) * 4096 + job_segs;
    // A plan depends on the utterances' lengths and on their offsets relative to the first one only (the kernel gets
    // the first utterance's address as its wave base), so the runs of a large batch of equal-length utterances -- and
    // every later batch with the same geometry -- share one plan.  Look-up: FNV-1a hash, arrays compared on a hit.
    const int64_t off0 = b->n_utt > 0 ? b->utt_offset[0] : 0;
    uint64_t hash = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { hash = (hash ^ v) * 1099511628211ull; };
    for (int u = 0; u < b->n_utt; ++u) { mix((uint64_t)(b->utt_offset[u] - off0)); mix((uint64_t)(uint32_t)b->utt_len[u]); }
    for (Plan *pl : h->plans)
        if (pl->key == key && pl->hash == hash && (int)pl->len.size() == b->n_utt &&
            std::equal(pl->len.begin(), pl->len.end(), b->utt_len) &&
            std::equal(pl->off.begin(), pl->off.end(), b->utt_offset, [&](int64_t a, int64_t c) { return a == c - off0; })) {
            pl->used = ++h->plan_clock;
            *out = pl;
            return AUD_OK;
        }
    const aud_params &p = h->p;
    int64_t total_segs = 0;
    for (int u = 0; u < b->n_utt; ++u) {
        if (b->utt_len[u] < 0) return fail(AUD_ERR_INVALID, "negative utterance length");
        total_segs += seg_count(p, b->utt_len[u]);
    }
    auto frames_of = [&](int64_t nseg) { return h->dedupe ? (nseg - 1) * h->seg_adv + p.segment_steps : nseg * p.segment_steps; };

    std::vector<Job> best_jobs;
    std::vector<int2> best_cta;
    int64_t best_cost = INT64_MAX, best_frames = 0;
    std::vector<int> candidates;
    if (job_segs > 0) candidates.push_back(job_segs);
    else candidates = {1 << 20, 64, 32, 16, 8, 4};
    for (int J : candidates) {
        std::vector<Job> jobs;
        int64_t seg = 0, frames = 0, pairs_total = 0;
        for (int u = 0; u < b->n_utt; ++u) {
            const int64_t n = seg_count(p, b->utt_len[u]);
            if (n > 0) {
                const int64_t parts = (n + J - 1) / J;
                for (int64_t k = 0; k < parts; ++k) {
                    const int64_t s0 = n * k / parts, s1 = n * (k + 1) / parts;
                    Job jb{};
                    jb.wave_off = b->utt_offset[u] - off0;   // relative to the first utterance (see fill_kparams)
                    jb.out_seg = seg + s0;
                    jb.utt_len = b->utt_len[u];
                    jb.seg0 = (int)s0;
                    jb.nseg = (int)(s1 - s0);
                    const int64_t nf = frames_of(s1 - s0);
                    if (nf > (1 << 28) || frames > INT32_MAX - nf - 4096 || seg + s1 > INT32_MAX)
                        return fail(AUD_ERR_UNSUPPORTED, "batch too large for one call (frame / segment index overflow)");
                    jb.nframes = (int)nf;
                    jb.frame_base = (int)frames;
                    frames += nf;
                    pairs_total += (nf + 1) / 2;
                    jobs.push_back(jb);
                }
            }
            seg += n;
        }
        // contiguous split by cumulative pairs (n_cta <= 0: no dealing, the general path indexes jobs directly)
        const int nc = n_cta <= 0 ? 1 : (int)std::max<int64_t>(1, std::min<int64_t>(n_cta, (pairs_total + 29) / 30));
        std::vector<int2> cta(nc);
        size_t ji = 0;
        int64_t done_pairs = 0, worst = 0;
        bool too_many = false;
        for (int c = 0; c < nc; ++c) {
            const int64_t target = pairs_total * (c + 1) / nc;
            cta[c].x = (int)ji;
            int64_t mine = 0;
            while (ji < jobs.size()) {
                const int64_t pj = (jobs[ji].nframes + 1) / 2;
                // take the job if that leaves us closer to the target (the last CTA takes whatever is left)
                if (c + 1 < nc && done_pairs + mine + pj - target > target - (done_pairs + mine) && mine > 0) break;
                if (c + 1 < nc && done_pairs + mine >= target) break;
                jobs[ji].pair_base = (int)mine;
                mine += pj;
                ++ji;
            }
            cta[c].y = (int)ji;
            if (n_cta > 0 && cta[c].y - cta[c].x > kMaxJobs) too_many = true;
            done_pairs += mine;
            worst = std::max(worst, mine);
        }
        if (too_many) continue;
        if (worst < best_cost) {   // the bottleneck CTA decides the launch time; ties keep the larger jobs (less halo)
            best_cost = worst;
            best_jobs.swap(jobs);
            best_cta.swap(cta);
            best_frames = frames;
        }
    }
    if (best_cost == INT64_MAX)
        return fail(kPlanTooMany, "could not plan the batch: too many jobs per CTA");
    Plan *pl = nullptr;
    if (h->plans.size() >= 24) {   // evict the least recently used entry
        size_t lru = 0;
        for (size_t i = 1; i < h->plans.size(); ++i)
            if (h->plans[i]->used < h->plans[lru]->used) lru = i;
        pl = h->plans[lru];
        h->plans.erase(h->plans.begin() + lru);
        cudaDeviceSynchronize();   // a launch may still be reading the evicted plan's device buffers
        delete pl;
    }
    pl = new (std::nothrow) Plan();
    if (!pl) return fail(AUD_ERR_NOMEM, "out of host memory");
    pl->off.resize((size_t)b->n_utt);
    for (int u = 0; u < b->n_utt; ++u) pl->off[u] = b->utt_offset[u] - off0;
    pl->len.assign(b->utt_len, b->utt_len + b->n_utt);
    pl->hash = hash;
    pl->key = key;
    pl->used = ++h->plan_clock;
    pl->jobs.swap(best_jobs);
    pl->cta_jobs.swap(best_cta);
    pl->total_segs = total_segs;
    pl->total_frames = best_frames;
    h->plans.push_back(pl);
    *out = pl;
    return AUD_OK;
}

// The fields of KParams that do not depend on the launch shape.
static void fill_kparams(KParams &kp, const aud_handle *h, const aud_batch *b, const aud_outputs *o, const Plan *pl,
                         bool nosmooth, bool want_mfcc, int energy_bins, int in_i16) {
    const aud_params &p = h->p;
    kp.step = p.step_samples; kp.stride = p.stride_samples; kp.S = p.segment_steps; kp.border = p.border_steps;
    kp.add = b->add_samples;
    kp.seg_adv = h->seg_adv; kp.dedupe = h->dedupe;
    kp.n_mel = p.n_mel; kp.n_coefs = p.n_coefs;
    kp.nosmooth = nosmooth ? 1 : 0;
    kp.energy_bins = energy_bins;
    kp.prev = (float)p.prev_smooth; kp.cur = (float)p.cur_smooth;
    kp.log_off = (float)p.log_offset; kp.log_min = (float)p.log_min;
    kp.comp_log_pow = p.comp_log_pow; kp.log1p_path = (p.log_offset == 1.0);
    kp.mel_log_off = (float)p.mel_log_off; kp.mel_log_min = (float)p.mel_log_min;
    kp.renorm = p.renorm; kp.renorm_min = (float)p.renorm_min; kp.renorm_scale = (float)p.renorm_scale;
    kp.want_mfcc = want_mfcc ? 1 : 0; kp.do_deltas = p.deltas; kp.c0_energy = p.mfcc_c0_energy;
    kp.g_on = h->g_on; kp.g_nf = p.gabor_nf; kp.g_sx = p.gabor_size_x; kp.g_sy = p.gabor_size_y;
    kp.g_stx = p.gabor_stride_x; kp.g_sty = p.gabor_stride_y; kp.g_dims = p.gabor_out_dims;
    kp.g_by_time = p.gabor_by_time; kp.g_nt = h->g_nt; kp.g_nfy = h->g_nfy; kp.g_tmaxstrides = h->g_tmaxstrides;
    kp.g_len = (int)h->gabor_len;
    if (p.gabor_out_dims == 2) {
        kp.g_str0 = p.gabor_shape[1];
    } else {
        kp.g_str0 = p.gabor_shape[1] * p.gabor_shape[2] * p.gabor_shape[3];
        kp.g_str1 = p.gabor_shape[2] * p.gabor_shape[3];
        kp.g_str2 = p.gabor_shape[3];
    }
    kp.g_gain = (float)p.gabor_gain;
    kp.dct = (const float *)h->d_dct.p; kp.gabor = (const float *)h->d_gabor.p;
    // the plan's jobs address samples relative to the batch's first utterance (reinterpreted as int16 PCM when in_i16)
    kp.wave = reinterpret_cast<const char *>(b->wave) + (b->n_utt > 0 ? b->utt_offset[0] : 0) * (in_i16 ? 2 : 4);
    kp.in_i16 = in_i16;
    kp.jobs = (const Job *)pl->d_jobs.p; kp.cta_jobs = (const int2 *)pl->d_cta_jobs.p;
    kp.o_mel = o->mel; kp.o_mfcc = o->mfcc; kp.o_d1 = o->deltas; kp.o_d2 = o->delta_deltas;
    kp.o_energy = o->energy; kp.o_gabor = o->gabor;
}

static int32_t upload_plan(Plan *pl, cudaStream_t st) {
    if (pl->uploaded) return AUD_OK;
    AUD_CUDA(pl->d_jobs.reserve(pl->jobs.size() * sizeof(Job)));
    AUD_CUDA(pl->d_cta_jobs.reserve(pl->cta_jobs.size() * sizeof(int2)));
    AUD_CUDA(cudaMemcpyAsync(pl->d_jobs.p, pl->jobs.data(), pl->jobs.size() * sizeof(Job), cudaMemcpyHostToDevice, st));
    AUD_CUDA(cudaMemcpyAsync(pl->d_cta_jobs.p, pl->cta_jobs.data(), pl->cta_jobs.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
    // no synchronisation: a copy from pageable memory returns once the driver has staged the source, the vectors live as
    // long as the plan, and the launch that reads the tables follows in the same stream
    pl->uploaded = true;
    return AUD_OK;
}

static int32_t launch_power_outputs(aud_handle *h, const KParams &kp, const Plan *pl, const aud_outputs *o, int pitch,
                                    cudaStream_t st) {
    PowParams q{};
    q.step = kp.step; q.stride = kp.stride; q.S = kp.S; q.border = kp.border; q.add = kp.add; q.seg_adv = kp.seg_adv;
    q.n_win = h->p.win_samples; q.bins = h->bins; q.pitch = pitch;
    q.prev = kp.prev; q.cur = kp.cur; q.log_off = kp.log_off; q.log_min = kp.log_min;
    q.comp_log_pow = kp.comp_log_pow; q.log1p_path = kp.log1p_path;
    q.jobs = kp.jobs; q.rawpow = kp.rawpow; q.o_power = o->power; q.o_logpower = o->logpower;
    power_segments_kernel<<<(int)pl->jobs.size(), 256, 0, st>>>(q);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return failf(AUD_ERR_CUDA, "power_segments_kernel launch failed: %s", cudaGetErrorString(e));
    ++h->launches;
    return AUD_OK;
}

// WinSamples != 400: frame power by the folded-DFT kernel, then one CTA per segment (aud_generic.cuh).
static int32_t run_generic(aud_handle *h, const aud_batch *b, const aud_outputs *o, cudaStream_t st, int in_i16) {
    const aud_params &p = h->p;
    const bool want_mfcc = p.mfcc && (o->mfcc || o->deltas || o->delta_deltas);
    const bool nosmooth = (p.prev_smooth == 0.0 && p.cur_smooth == 1.0);
    const int energy_bins = (o->energy || (want_mfcc && p.mfcc_c0_energy)) ? h->energy_bins : 0;
    Plan *pl = nullptr;
    // jobs of at most 32 segments: the per-job kernels (frame scales) then see enough CTAs even for one long utterance, at
    // the price of recomputing the few frames two neighbouring jobs share
    int32_t rc = build_plan(h, b, 0, h->opt_job_segs > 0 ? h->opt_job_segs : 32, &pl);
    if (rc != AUD_OK) return rc;
    if (pl->jobs.empty()) return AUD_OK;
    rc = upload_plan(pl, st);
    if (rc != AUD_OK) return rc;
    // Only the bins somebody reads are transformed: the mel bank stops at HiHz (bin 200 of 552 at 44.1 kHz with the
    // default 8 kHz), Energy reads bins < SegmentSteps; PowerSegment / LogPowerSegment, when asked for, need them all.
    const int need_bins = (o->power || o->logpower) ? h->bins : std::min(h->bins, std::max(h->g_need, energy_bins));
    const int tc_tiles = (need_bins + h->tc_tn - 1) / h->tc_tn;
    // the scratch rows are as wide as what is computed (whole tiles of either kernel)
    const int pitch = std::min(h->g_pitch, (std::max(tc_tiles * h->tc_tn, (need_bins + kGN - 1) / kGN * kGN) + 63) / 64 * 64);
    AUD_CUDA(h->d_rawpow.reserve((size_t)(pl->total_frames + kGM) * pitch * sizeof(float)));

    GParams g{};
    fill_kparams(g.k, h, b, o, pl, nosmooth, want_mfcc, energy_bins, in_i16);
    g.n_win = p.win_samples; g.bins = h->bins; g.pitch = pitch; g.tpitch = h->g_pitch;
    g.total_frames = (int)pl->total_frames; g.njobs = (int)pl->jobs.size();
    g.cos_t = (const float *)h->d_cos.p; g.sin_t = (const float *)h->d_sin.p;
    g.mel_lo = (const int *)h->d_gmel_lo.p; g.mel_n = (const int *)h->d_gmel_n.p;
    g.mel_w = (const float *)h->d_gmel_w.p; g.mel_wpitch = h->g_wpitch;
    g.rawpow = (float *)h->d_rawpow.p;
    g.k.rawpow = g.rawpow;
    // tiles of one segment
    const size_t S = p.segment_steps, MS = (size_t)p.n_mel * S, CS = (size_t)p.n_coefs * S;
    const bool gab = h->g_on && o->gabor;
    size_t off = MS;
    g.t_off[0] = (int)off; off += S;
    g.t_off[1] = (int)off; off += want_mfcc ? CS : 0;
    g.t_off[2] = (int)off; off += (want_mfcc && p.deltas) ? CS : 0;
    g.t_off[3] = (int)off; off += (want_mfcc && p.deltas) ? CS : 0;
    g.t_off[4] = (int)off; off += gab ? (size_t)h->gabor_len : 0;
    off = (off + 3) & ~(size_t)3;
    g.t_off[5] = (int)off;
    g.k.dct_floats = want_mfcc ? p.n_coefs * ((p.n_mel + 3) / 4 * 4) : 0;
    off += (size_t)g.k.dct_floats;
    g.k.gw_floats = gab ? p.gabor_size_x * p.gabor_size_y * ((p.gabor_nf + 7) / 8 * 8) : 0;
    off += (size_t)g.k.gw_floats;
    size_t smem = off * sizeof(float) + 16;
    if (smem > (size_t)h->max_smem_optin)
        return failf(AUD_ERR_UNSUPPORTED, "segment geometry does not fit in shared memory (%zu bytes needed, %d available)", smem, h->max_smem_optin);
    g.k.need_tiles = 1;
    if (!gab) g.k.g_on = 0;   // the tile stage only runs the stages somebody asked for
    {   // the mel bank's power bins and taps staged in shared memory, if they fit beside the tiles (and leave room for
        // several CTAs per SM)
        const int np = (std::min(h->g_need, pitch) + 3) & ~3, wp = h->g_wpitch | 1;
        const size_t extra = ((size_t)S * np + (size_t)p.n_mel * wp) * sizeof(float);
        if (smem + extra <= (size_t)h->max_smem_optin / 2) {
            g.stage_np = np; g.stage_wp = wp;
            smem += extra;
        }
    }

    if (o->gabor && !h->g_on && h->gabor_len > 0)   // Convolve returned without writing (gabor.go:226-229)
        AUD_CUDA(cudaMemsetAsync(o->gabor, 0, (size_t)pl->total_segs * h->gabor_len * sizeof(float), st));

    cudaError_t e;
    if (h->opt_dft_tc) {
        tc::TcParams t{};
        t.g = g;
        t.tab = (const __half *)h->d_tc_tab.p;
        t.KB = h->tc_kb; t.n_nt_tab = h->tc_nt; t.tn = h->tc_tn;
        t.n_nt = tc_tiles;
        t.n_items = (int)((pl->total_frames + tc::kTM - 1) / tc::kTM) * t.n_nt;
        const unsigned grid = (unsigned)std::min(t.n_items, h->sm_count);
        // [total_frames] float2 scales, [total_frames] job of every frame row, [total_segs] job of every segment, block maxima
        const int W = (p.win_samples + p.step_samples - 1) / p.step_samples;
        const size_t nf = (size_t)pl->total_frames, nsg = (size_t)pl->total_segs, nbm = nf + pl->jobs.size() * (size_t)(W - 1) + 1;
        AUD_CUDA(h->d_tc_scale.reserve(nf * sizeof(float2) + (nf + nsg + nbm) * sizeof(int)));
        float2 *row_scale = (float2 *)h->d_tc_scale.p;
        int *row_job = (int *)(row_scale + nf), *seg_job = row_job + nf;
        float *blockmax = (float *)(seg_job + nsg);
        t.row_scale = row_scale;
        t.row_job = row_job;
        if (in_i16) tc::frame_scale_kernel<true><<<(unsigned)pl->jobs.size(), 256, 0, st>>>(g, row_scale, row_job, seg_job, blockmax, W);
        else tc::frame_scale_kernel<false><<<(unsigned)pl->jobs.size(), 256, 0, st>>>(g, row_scale, row_job, seg_job, blockmax, W);
        g.seg_job = seg_job;   // segment_features_kernel: no binary search per segment
        ++h->launches;
        auto kern = in_i16 ? tc::dft_power_tc_kernel<true> : tc::dft_power_tc_kernel<false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes);
        if (e == cudaSuccess) {
            kern<<<grid, tc::kThreads, tc::kSmemBytes, st>>>(t);
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) return failf(AUD_ERR_CUDA, "dft_power_tc_kernel launch failed: %s", cudaGetErrorString(e));
    } else {
        const dim3 grid1((unsigned)((pl->total_frames + kGM - 1) / kGM), (unsigned)((need_bins + kGN - 1) / kGN));
        dft_power_kernel<<<grid1, 256, 0, st>>>(g);
        e = cudaGetLastError();
        if (e != cudaSuccess) return failf(AUD_ERR_CUDA, "dft_power_kernel launch failed: %s", cudaGetErrorString(e));
    }
    ++h->launches;
    if (o->mel || o->energy || want_mfcc || gab) {
        e = cudaFuncSetAttribute(segment_features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) {
            segment_features_kernel<<<(unsigned)pl->total_segs, 128, smem, st>>>(g);
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) return failf(AUD_ERR_CUDA, "segment_features_kernel launch failed: %s", cudaGetErrorString(e));
        ++h->launches;
    }
    if (o->power || o->logpower) return launch_power_outputs(h, g.k, pl, o, pitch, st);
    return AUD_OK;
}

static int32_t run_fused_once(aud_handle *h, const aud_batch *b, const aud_outputs *o, cudaStream_t st, int in_i16) {
    const aud_params &p = h->p;
    const bool want_mfcc = p.mfcc && (o->mfcc || o->deltas || o->delta_deltas);
    // tiles stage everything that is not a plain gather of per-frame log-mel
    const bool nosmooth = (p.prev_smooth == 0.0 && p.cur_smooth == 1.0);
    const bool need_tiles = want_mfcc || (h->g_on && o->gabor) || !nosmooth;
    // Energy (and the low power bins it is built from) only when somebody consumes it
    const int energy_bins = (o->energy || (want_mfcc && p.mfcc_c0_energy)) ? h->energy_bins : 0;
    static const int kWarpChoices[] = {12, 10, 8, 6};
    const Needs needs{need_tiles, energy_bins > 0, want_mfcc, p.deltas != 0, h->g_on && o->gabor != nullptr, gabor_direct(h, o->gabor)};
    // Warp split, measured on the BASELINE batch: 12 FFT + 4 epilogue warps put three FFT warps and one
    // epilogue warp on each of the SM's four schedulers and win for plain log-mel and for gabor; the MFCC /
    // smoothing / Energy epilogue is heavier and wants 10 + 6.  FFT + epilogue warps stay within 16 (128
    // registers per thread without spills).  Only the shapes listed in aud_launch.h are compiled in.
    const bool light = nosmooth && !need_tiles && energy_bins == 0;
    const bool scan_or_mfcc = want_mfcc || !nosmooth || energy_bins > 0;
    const int nepi = h->opt_epi > 0 ? (h->opt_epi >= 5 ? 6 : 4) : (light ? 4 : scan_or_mfcc ? 6 : 4);
    Launch L{};
    bool found = false;
    int epirec = 0;
    for (int pass = 0; pass < 2 && !found; ++pass)   // first a plan whose tiles hold most of a round, then any plan
        for (int w : kWarpChoices) {
            if (h->opt_warps > 0 && w != h->opt_warps) continue;
            // frame-pair records: by the (idle) epilogue warps for plain log-mel, else by the FFT warps
            epirec = (light && have_shape(w, nepi, 1)) ? 1 : 0;
            if (!have_shape(w, nepi, epirec)) continue;
            L = pick_launch(h, w, needs, energy_bins, pass == 0, epirec ? kRecRounds : 0);
            if (L.smem <= (size_t)h->max_smem_optin && L.tile_cap >= 1 && 6 * w <= kMaxDone) { found = true; break; }
        }
    if (!found)
        return failf(AUD_ERR_UNSUPPORTED, "segment geometry does not fit in shared memory (%zu bytes needed, %d available)%s",
                     L.smem, h->max_smem_optin, h->opt_warps > 0 ? " with the requested warps option (12, 10, 8 or 6)" : "");
    const int n_cta = h->opt_ctas > 0 ? h->opt_ctas : h->sm_count;
    Plan *pl = nullptr;
    int32_t rc = build_plan(h, b, n_cta, h->opt_job_segs, &pl);
    if (rc != AUD_OK) return rc;
    if (pl->jobs.empty()) return AUD_OK;
    rc = upload_plan(pl, st);
    if (rc != AUD_OK) return rc;
    const bool want_pow = o->power || o->logpower;
    if (want_pow) AUD_CUDA(h->d_rawpow.reserve((size_t)(pl->total_frames + 2) * kPowPitch * sizeof(float)));

    KParams kp{};
    fill_kparams(kp, h, b, o, pl, nosmooth, want_mfcc, energy_bins, in_i16);
    kp.ps = L.ps; kp.win_len = L.win_len; kp.contig = L.contig; kp.ring = L.ring;
    kp.mel_pitch = mel_ring_pitch(p.n_mel);
    kp.need_tiles = L.need_tiles; kp.tile_cap = L.tile_cap; kp.tile_floats = (int)L.tile_floats;
    for (int i = 0; i < 5; ++i) kp.t_off[i] = L.t_off[i];
    kp.dct_floats = (int)L.dct_floats;
    kp.gw_floats = (int)L.gw_floats;
    kp.rec_rounds = epirec ? kRecRounds : 0;
    if (!needs.gabor) kp.g_on = 0;   // the tile stage only runs the stages somebody asked for (no gabor tile otherwise)
    kp.g_direct = needs.gabor_direct ? 1 : 0;
    kp.tw2 = (const float2 *)h->d_tw.p;
    kp.mel_start = (const int *)h->d_mel_start.p; kp.mel_quads = (const int *)h->d_mel_width.p;
    kp.mel_taps = (const float *)h->d_mel_taps.p; kp.mel_sched = (const int4 *)h->d_mel_sched.p;
    kp.mel_taps_len = h->mel_taps_len; kp.mel_tasks = h->mel_tasks;
    kp.rawpow = want_pow ? (float *)h->d_rawpow.p : nullptr;

    if (o->gabor && !h->g_on && h->gabor_len > 0)   // Convolve returned without writing (gabor.go:226-229)
        AUD_CUDA(cudaMemsetAsync(o->gabor, 0, (size_t)pl->total_segs * h->gabor_len * sizeof(float), st));

    const int grid = (int)pl->cta_jobs.size();
#ifdef AUD_DEBUG_CHECKS
    AUD_CUDA(h->d_dbg.reserve(sizeof(int)));
    AUD_CUDA(cudaMemsetAsync(h->d_dbg.p, 0, sizeof(int), st));
    kp.dbg = (int *)h->d_dbg.p;
#endif
    const cudaError_t e = launch_shape(L.warps, nepi, epirec, kp, grid, L.smem, st);
#ifdef AUD_DEBUG_CHECKS
    {
        int code = 0;
        AUD_CUDA(cudaStreamSynchronize(st));
        AUD_CUDA(cudaMemcpy(&code, h->d_dbg.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (code != 0) return failf(AUD_ERR_CUDA, "debug check %d failed in fused_features_kernel<%d,%d,%d>", code, L.warps, nepi, epirec);
    }
#endif
    if (e != cudaSuccess) return failf(AUD_ERR_CUDA, "fused_features_kernel launch failed: %s", cudaGetErrorString(e));
    ++h->launches;

    if (want_pow) return launch_power_outputs(h, kp, pl, o, kPowPitch, st);
    return AUD_OK;
}

// One launch handles at most kMaxJobs jobs per persistent CTA; larger batches (BASELINE config 4: 65,536
// utterances) are cut into runs of utterances, each with its own plan and its slice of the outputs.
static int32_t run_fused_range(aud_handle *h, const aud_batch *b, const aud_outputs *o, cudaStream_t st, int in_i16,
                               int u0, int u1, int64_t seg0) {
    const aud_params &p = h->p;
    const size_t S = p.segment_steps;
    const size_t per_seg[8] = {(size_t)p.n_mel * S, (size_t)p.n_coefs * S, (size_t)p.n_coefs * S, (size_t)p.n_coefs * S,
                               S, (size_t)h->gabor_len, (size_t)h->bins * S, (size_t)h->bins * S};
    const int n_cta = h->opt_ctas > 0 ? h->opt_ctas : h->sm_count;
    const int chunk = std::max(1, n_cta * 48);
    if (u1 - u0 > chunk) {
        for (int a = u0; a < u1; a += chunk) {
            const int e = std::min(u1, a + chunk);
            const int32_t rc = run_fused_range(h, b, o, st, in_i16, a, e, seg0);
            if (rc != AUD_OK) return rc;
            for (int u = a; u < e; ++u) seg0 += seg_count(p, b->utt_len[u]);
        }
        return AUD_OK;
    }
    aud_batch sb = *b;
    sb.utt_offset = b->utt_offset + u0;
    sb.utt_len = b->utt_len + u0;
    sb.n_utt = u1 - u0;
    aud_outputs so{};
    float *const *src = &o->mel;
    float **dst = &so.mel;
    for (int i = 0; i < 8; ++i) dst[i] = src[i] ? src[i] + (size_t)seg0 * per_seg[i] : nullptr;
    const int32_t rc = run_fused_once(h, &sb, &so, st, in_i16);
    if (rc != kPlanTooMany) return rc;
    if (u1 - u0 < 2) return fail(AUD_ERR_UNSUPPORTED, "could not plan the batch: one utterance needs too many jobs (raise the job_segs option)");
    // very uneven utterance lengths: halve the run and retry
    const int mid = u0 + (u1 - u0) / 2;
    int32_t rc2 = run_fused_range(h, b, o, st, in_i16, u0, mid, seg0);
    if (rc2 != AUD_OK) return rc2;
    for (int u = u0; u < mid; ++u) seg0 += seg_count(p, b->utt_len[u]);
    return run_fused_range(h, b, o, st, in_i16, mid, u1, seg0);
}

static int32_t run_device(aud_handle *h, const aud_batch *b, const aud_outputs *o, cudaStream_t st, int in_i16 = 0) {
    for (int u = 0; u < b->n_utt; ++u)
        if (b->utt_len[u] < 0) return fail(AUD_ERR_INVALID, "negative utterance length");
    if (!h->fused) return run_generic(h, b, o, st, in_i16);
    return run_fused_range(h, b, o, st, in_i16, 0, b->n_utt, 0);
}

// Geometry of agabor.Convolve for one [n_mel][segment_steps] input (agabor/gabor.go:231-262): positions
// along time and frequency, the 2-D by-time column stride, the output length, and the configurations the
// reference would panic on (or whose result depends on its loop order).
struct GaborGeom {
    int64_t len = 0;
    int on = 0, nt = 0, nfy = 0, tmaxstrides = 1;
};
static int32_t gabor_geometry(const aud_params &p, bool have_filters, GaborGeom *g) {
    int32_t rc = AUD_OK;
    {
        if (!have_filters) rc = fail(AUD_ERR_INVALID, "gabor_nf > 0 but gabor_filters is NULL");
        else if (p.gabor_size_x < 1 || p.gabor_size_y < 1 || p.gabor_stride_x < 1 || p.gabor_stride_y < 1)
            rc = fail(AUD_ERR_INVALID, "gabor size / stride must be positive");
        else if (p.gabor_out_dims != 2 && p.gabor_out_dims != 4)
            rc = fail(AUD_ERR_INVALID, "The output tensor should have 2 or 4 dimensions");   // gabor.go:260
        if (rc == AUD_OK) {
            int64_t len = 1;
            for (int d = 0; d < p.gabor_out_dims; ++d) {
                if (p.gabor_shape[d] < 0) rc = fail(AUD_ERR_INVALID, "negative gabor output dimension");
                len *= p.gabor_shape[d];
            }
            g->len = len;
        }
        if (rc == AUD_OK && p.segment_steps >= p.gabor_size_x) {   // else Convolve logs and returns (gabor.go:226-229)
            const int S = p.segment_steps, M = p.n_mel;
            int tmax = 1, fmax = 1;
            if (p.gabor_out_dims == 2) {
                const int x = S - p.gabor_size_x;
                if (!(x == 0 || x < p.gabor_stride_x)) tmax = x + 1;
                g->tmaxstrides = (S - p.gabor_size_x) / p.gabor_stride_x + 1;
                const int y = M - p.gabor_size_y;
                if (!(y == 0 || y < p.gabor_stride_y)) fmax = y + 1;
            } else {
                tmax = (int)std::min((double)p.gabor_shape[1] * p.gabor_stride_x, (double)(S - p.gabor_stride_x));
                fmax = (int)std::min((double)p.gabor_shape[0] * p.gabor_stride_y, (double)(M - p.gabor_stride_y));
            }
            g->nt = tmax <= 0 ? 0 : (tmax - 1) / p.gabor_stride_x + 1;
            g->nfy = fmax <= 0 ? 0 : (fmax - 1) / p.gabor_stride_y + 1;
            g->on = (g->nt > 0 && g->nfy > 0) ? 1 : 0;
            if (g->on) {
                const int64_t last_in = (int64_t)((g->nfy - 1) * p.gabor_stride_y + p.gabor_size_y - 1) * S +
                                        (g->nt - 1) * p.gabor_stride_x + p.gabor_size_x - 1;
                if (last_in >= (int64_t)M * S) rc = fail(AUD_ERR_PANIC, "agabor.Convolve reads past the mel tensor (reference panics)");
                std::vector<char> hit((size_t)g->len, 0);
                for (int ti = 0; ti < g->nt && rc == AUD_OK; ++ti)
                    for (int fi = 0; fi < g->nfy && rc == AUD_OK; ++fi)
                        for (int flt = 0; flt < p.gabor_nf; ++flt) {
                            int64_t on, off;
                            if (p.gabor_out_dims == 2) {
                                const int64_t x = p.gabor_by_time ? ti + (int64_t)g->tmaxstrides * flt : flt + (int64_t)ti * p.gabor_nf;
                                on = (int64_t)(2 * fi) * p.gabor_shape[1] + x;
                                off = on + p.gabor_shape[1];
                            } else {
                                const int64_t s2 = p.gabor_shape[3], s1 = s2 * p.gabor_shape[2], s0 = s1 * p.gabor_shape[1];
                                on = fi * s0 + ti * s1 + flt;
                                off = on + s2;
                            }
                            if (on < 0 || off < 0 || on >= g->len || off >= g->len) {
                                rc = fail(AUD_ERR_PANIC, "agabor.Convolve writes past the output tensor (reference panics)");
                                break;
                            }
                            if (hit[on] || hit[off]) {
                                rc = fail(AUD_ERR_UNSUPPORTED, "gabor output geometry maps two results to one cell (order-dependent in the reference)");
                                break;
                            }
                            hit[on] = hit[off] = 1;
                        }
            }
        }
    }
    return rc;
}

}  // namespace aud

using namespace aud;

extern "C" {

const char *aud_last_error(void) { return g_last_error.c_str(); }
#ifdef AUD_DEBUG_CHECKS
int32_t aud_version(void) { return -200; }   // negative: the debug-check build
#else
int32_t aud_version(void) { return 200; }
#endif

int32_t aud_create(const aud_params *pp, const int32_t *bin_pts, const double *mel_filters, const double *gabor_filters,
                   const double *dct, int32_t device, aud_handle **out) {
    if (!out) return fail(AUD_ERR_INVALID, "aud_create: out is NULL");
    *out = nullptr;
    if (!pp || !bin_pts || !mel_filters) return fail(AUD_ERR_INVALID, "aud_create: NULL parameter block or mel tables");
    const aud_params &p = *pp;
    if (p.sample_rate <= 0) return fail(AUD_ERR_INVALID, "sample rate <= 0");
    if (p.win_samples < 2 || p.step_samples < 1 || p.stride_samples < 1 || p.segment_steps < 1 || p.border_steps < 0 ||
        p.n_mel < 1 || p.segment_samples < 0)
        return fail(AUD_ERR_INVALID, "aud_create: non-positive window / step / stride / steps / filters");
    if (p.mfcc && (p.n_coefs < 1 || p.n_coefs > p.n_mel)) return fail(AUD_ERR_PANIC, "NCoefs must be in 1..NFilters (reference indexes past the DCT output)");
    const bool fused = p.win_samples == kN;   // the fused kernel is built for 25 ms at 16 kHz; other lengths take the general path
    if (p.win_samples > 4096)
        return failf(AUD_ERR_UNSUPPORTED, "WinSamples = %d: windows longer than 4096 samples are not supported", p.win_samples);
    const int bins = p.win_samples / 2 + 1;
    // SndEnv.ProcessSegment's Energy loop (sndenv.go:360-366) reads LogPowerSegment row s for every step s, MFCC or not,
    // smoothing or not: with more steps than spectrum rows the reference panics.  mfcc_c0_energy = 1 says the caller is
    // a SndEnv; per-step callers (gaborview's own loop, gbv.go:545-559) never run that loop and may use longer segments.
    if (p.mfcc_c0_energy && p.segment_steps > bins)
        return fail(AUD_ERR_PANIC, "SegmentSteps > WinSamples/2+1: SndEnv.ProcessSegment's Energy loop indexes past LogPowerSegment (reference panics)");

    // mel taps: the reference reads filters.Value({flt, fi}) = flat[flt*(n_mel+2)+fi] for fi < width.
    // The kernel keeps a frame's power in "padded natural order" (index k + k/20, one pad slot of
    // value 0 after every 20 bins), so the taps are laid out over those indices with weight 0 on pads.
    const int npts = p.n_mel + 2;
    std::vector<int> start(p.n_mel), quads(p.n_mel);
    std::vector<int> g_lo(p.n_mel), g_n(p.n_mel);   // general path: first bin and tap count per filter
    int max4 = 1, g_wpitch = 1, g_need = 1;   // g_need: bins [0, g_need) are all the mel bank reads
    for (int m = 0; m < p.n_mel; ++m) {
        const int lo = bin_pts[m], hi = bin_pts[m + 2];
        if (lo < 0 || hi >= bins) return fail(AUD_ERR_PANIC, "mel BinPts outside the power spectrum (reference panics)");
        const int nb = hi >= lo ? hi - lo + 1 : 0;
        if ((int64_t)m * npts + nb > (int64_t)p.n_mel * npts)
            return fail(AUD_ERR_PANIC, "mel filter table index out of range (reference panics)");
        g_lo[m] = lo; g_n[m] = nb;
        g_need = std::max(g_need, lo + nb);
        g_wpitch = std::max(g_wpitch, nb);
        const int lo_p = lo + lo / 20, hi_p = hi + hi / 20;
        start[m] = lo_p & ~1;                                  // 16-byte aligned (A, B) power pairs
        quads[m] = nb ? ((lo_p - start[m]) + (hi_p - lo_p + 1) + 3) / 4 : 0;
        max4 = std::max(max4, quads[m]);
    }
    int mel_pitch = 4 * max4;
    if ((mel_pitch / 4) % 2 == 0) mel_pitch += 4;              // odd number of 16-byte chunks per row: conflict-free
    std::vector<float> taps((size_t)p.n_mel * mel_pitch, 0.f);
    for (int m = 0; m < p.n_mel; ++m) {
        const int lo = bin_pts[m], hi = bin_pts[m + 2];
        for (int bin = lo; bin <= hi; ++bin)
            taps[(size_t)m * mel_pitch + ((bin + bin / 20) - start[m])] = (float)mel_filters[(size_t)m * npts + (bin - lo)];
    }
    // geometry of a pair's scratch (constant per handle): the sample window of the next round sits above
    // the power buffer, and the pair stride puts the three windows 20 banks apart (conflict-free 8-byte loads)
    const int dedupe0 = (p.stride_samples % p.step_samples == 0 && p.stride_samples / p.step_samples <= p.segment_steps) ? 1 : 0;
    const int contig = (dedupe0 && p.step_samples <= kN) ? 1 : 0;
    const int win_len = contig ? p.step_samples + kN : 2 * kN;
    int ps = std::max(kExchange, kWinOff + (win_len + 1) / 2);   // exchange rows (shifted by up to 10) and the window
    while (ps % 16 != 10) ++ps;

    // schedule of (pair, filter) tasks over the 32 lanes: widest first so that the lanes of one slot
    // run loops of similar length; inside a slot the tasks are then permuted so that the eight lanes
    // of every quarter-warp hit different 16-byte bank groups with their tap and power loads
    std::vector<std::pair<int, int>> tasks;   // (quads, pair << 16 | filter)
    for (int qq = 0; qq < kPairs; ++qq)
        for (int m = 0; m < p.n_mel; ++m) tasks.push_back({quads[m], (qq << 16) | m});
    std::stable_sort(tasks.begin(), tasks.end(), [](const auto &a, const auto &b2) { return a.first > b2.first; });
    const int mel_tasks = fused ? (int)((tasks.size() + 31) / 32) : 0;   // the general path has its own tables
    if (fused && (p.n_mel > 65535 || max4 > 127)) return fail(AUD_ERR_UNSUPPORTED, "mel filter bank too large for the task encoding");
    std::vector<int> sched((size_t)mel_tasks * 32 * 4, 0);   // int4 per (slot, lane): tap block, power offset, code, 0
    std::vector<float> taps_sl;                              // taps re-laid per slot
    for (int t = 0; t < mel_tasks; ++t) {
        int order[32];
        const int n_t = (int)std::min<size_t>(32, tasks.size() - (size_t)t * 32);
        for (int l = 0; l < 32; ++l) order[l] = l < n_t ? t * 32 + l : -1;
        const int slot_max = tasks[(size_t)t * 32].first;      // sorted: the first task of a slot is its longest
        // Bank groups.  A 128-bit load is served a quarter-warp at a time, conflict-free when its eight lanes
        // hit eight different 16-byte groups, i.e. when the tasks' power offsets differ mod 8 chunks.  A task
        // may start up to three chunks (6 taps of weight 0) early as long as it stays within the slot's loop
        // length, which moves its group; a small matching picks the shifts so that every group is used by at
        // most four tasks of the slot, then each quarter-warp takes one task of every group.
        int shift[32] = {0};
        {
            auto chunk0 = [&](int ti) {
                const int qq = tasks[ti].second >> 16, m = tasks[ti].second & 0xffff;
                return (qq * ps + start[m]) / 2;
            };
            auto can = [&](int ti, int k) {
                const int m = tasks[ti].second & 0xffff;
                if (start[m] - 2 * k < 0) return false;
                const int lo_p = bin_pts[m] + bin_pts[m] / 20, hi_p = bin_pts[m + 2] + bin_pts[m + 2] / 20;
                return ((lo_p - (start[m] - 2 * k)) + (hi_p - lo_p + 1) + 3) / 4 <= slot_max;
            };
            int owner[8][4];                     // group -> up to four tasks (lane-local indices)
            for (auto &o : owner) for (int &x : o) x = -1;
            int kof[32];
            for (int l = 0; l < 32; ++l) kof[l] = -1;
            // augmenting paths (Kuhn): task l tries its reachable groups; an occupied place may re-home its task
            std::vector<char> seen;
            std::function<bool(int)> place = [&](int l) -> bool {
                for (int k = 0; k < 4; ++k) {
                    if (!can(order[l], k)) continue;
                    const int g = ((chunk0(order[l]) - k) % 8 + 8) % 8;
                    for (int c = 0; c < 4; ++c) {
                        if (seen[g * 4 + c]) continue;
                        seen[g * 4 + c] = 1;
                        if (owner[g][c] < 0 || place(owner[g][c])) {
                            owner[g][c] = l;
                            kof[l] = k;
                            return true;
                        }
                    }
                }
                return false;
            };
            for (int l = 0; l < n_t; ++l) {
                seen.assign(32, 0);
                if (!place(l)) kof[l] = 0;       // no conflict-free home: keep its natural group
            }
            // lanes: quarter c takes owner[g][c] for every group g; unmatched tasks and idle lanes fill the gaps
            int lane_task[32];
            for (int &x : lane_task) x = -2;
            std::vector<char> used(32, 0);
            for (int g = 0; g < 8; ++g)
                for (int c = 0; c < 4; ++c)
                    if (owner[g][c] >= 0 && kof[owner[g][c]] >= 0) {
                        // owner[][] may hold stale entries of re-homed tasks: accept only a task whose final group is g
                        const int l = owner[g][c];
                        const int gl = ((chunk0(order[l]) - kof[l]) % 8 + 8) % 8;
                        if (gl == g && !used[l]) { lane_task[c * 8 + g] = l; used[l] = 1; }
                    }
            int free_lane = 0;
            for (int l = 0; l < 32; ++l) {
                if (used[l] || order[l] < 0) continue;
                while (lane_task[free_lane] != -2) ++free_lane;
                lane_task[free_lane] = l;
            }
            int new_order[32];
            for (int L = 0; L < 32; ++L) {
                const int l = lane_task[L];
                new_order[L] = l >= 0 ? order[l] : -1;
                shift[L] = l >= 0 ? std::max(0, kof[l]) : 0;
            }
            for (int L = 0; L < 32; ++L) order[L] = new_order[L];
        }
        for (int l = 0; l < 32; ++l) {
            int *d = &sched[((size_t)t * 32 + l) * 4];
            d[0] = (int)(taps_sl.size() / 4);          // float4 index of this slot's [quad][lane] block
            if (order[l] < 0) { d[1] = start[0]; d[2] = -1; continue; }
            const int qq = tasks[order[l]].second >> 16, m = tasks[order[l]].second & 0xffff;
            // every lane of a slot runs the slot's longest loop: shorter rows then read (weight 0) up to
            // 4*slot_max entries past their start, which must stay inside the zero-tailed power buffer [0, 219)
            if (start[m] - 2 * shift[l] + 4 * slot_max > 219)
                return fail(AUD_ERR_UNSUPPORTED, "mel filter bank geometry not supported by the fused kernel's task schedule");
            d[1] = qq * ps + start[m] - 2 * shift[l];
            d[2] = (slot_max << 24) | (qq << 16) | m;
        }
        // this slot's taps, [quad][lane][4], so that the 32 lanes read 32 consecutive 16-byte chunks
        const size_t base = taps_sl.size();
        taps_sl.resize(base + (size_t)slot_max * 32 * 4, 0.f);
        for (int l = 0; l < 32; ++l) {
            if (order[l] < 0) continue;
            const int m = tasks[order[l]].second & 0xffff;
            for (int i = 0; i < 4 * quads[m]; ++i) {   // the task's taps, behind 2*shift zeros
                const int o = i + 2 * shift[l];
                if (o < 4 * slot_max) taps_sl[base + ((size_t)(o / 4) * 32 + l) * 4 + (o % 4)] = taps[(size_t)m * mel_pitch + i];
            }
        }
    }

    aud_handle *h = new (std::nothrow) aud_handle();
    if (!h) return fail(AUD_ERR_NOMEM, "out of host memory");
    h->p = p;
    h->device = device;
    h->bins = bins;
    h->mel_taps_len = (int)taps_sl.size();
    h->ps = ps; h->contig = contig; h->win_len = win_len;
    h->mel_tasks = mel_tasks;
    h->fused = fused ? 1 : 0;
    h->g_wpitch = g_wpitch;
    h->g_need = g_need;
    // bin tiles of the tensor-core kernel: as few tiles of <= 128 bins as cover the spectrum, as narrow as that allows
    h->tc_nt = (bins + tc::kTNMax - 1) / tc::kTNMax;
    h->tc_tn = ((bins + h->tc_nt - 1) / h->tc_nt + 15) / 16 * 16;
    h->g_pitch = (h->tc_nt * h->tc_tn + 63) / 64 * 64;
    h->dedupe = (p.stride_samples % p.step_samples == 0) ? 1 : 0;
    h->seg_adv = h->dedupe ? p.stride_samples / p.step_samples : p.segment_steps;
    // frames are shared between segments only while the segments overlap or abut in slot space
    if (h->dedupe && h->seg_adv > p.segment_steps) { h->dedupe = 0; h->seg_adv = p.segment_steps; }
    const bool need_energy = true;   // Energy is an output in its own right
    h->energy_bins = (need_energy && p.comp_log_pow) ? std::min(p.segment_steps, bins) : 0;

    // gabor geometry (agabor/gabor.go:231-262) and the bounds the reference would panic on
    int32_t rc = AUD_OK;
    if (p.gabor_nf > 0) {
        GaborGeom gg{};
        rc = gabor_geometry(p, gabor_filters != nullptr, &gg);
        h->gabor_len = gg.len; h->g_on = gg.on; h->g_nt = gg.nt; h->g_nfy = gg.nfy; h->g_tmaxstrides = gg.tmaxstrides;
    }
    if (rc != AUD_OK) { delete h; return rc; }

    // device side
    cudaError_t e = cudaSetDevice(device);
    cudaDeviceProp prop{};
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        delete h;
        return failf(AUD_ERR_CUDA, "no usable CUDA device %d: %s (this library has no CPU fallback)", device, cudaGetErrorString(e));
    }
    h->sm_count = prop.multiProcessorCount;
    h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;

    std::vector<float2> tw(kN / 2);   // tw[k1*10 + j] = W400^{2 j k1} / 2
    for (int k1 = 0; k1 < 20; ++k1)
        for (int j = 0; j < 10; ++j) {
            const double a = -2.0 * 3.14159265358979323846264338327950288 * (double)(k1 * 2 * j) / (double)kN;
            tw[k1 * 10 + j] = make_float2(0.5f * (float)std::cos(a), 0.5f * (float)std::sin(a));   // the 1/2 of the real-pair split (exact)
        }
    std::vector<double> dct_d;
    if (!dct) {
        dct_d.resize((size_t)p.n_coefs * p.n_mel);
        aud_dct1_matrix(p.n_mel, p.n_coefs, dct_d.data());
        dct = dct_d.data();
    }
    std::vector<float> dct_f((size_t)std::max(p.n_coefs, 1) * p.n_mel);
    for (size_t i = 0; i < (size_t)p.n_coefs * p.n_mel; ++i) dct_f[i] = (float)dct[i];
    std::vector<float> gab_f((size_t)std::max(1, p.gabor_nf * p.gabor_size_x * p.gabor_size_y));
    for (size_t i = 0; i < (size_t)p.gabor_nf * p.gabor_size_x * p.gabor_size_y; ++i) gab_f[i] = (float)gabor_filters[i];

    auto up = [&](DevBuf &b, const void *src, size_t bytes) -> cudaError_t {
        cudaError_t e2 = b.reserve(bytes);
        if (e2 != cudaSuccess) return e2;
        return cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice);
    };
    e = up(h->d_tw, tw.data(), tw.size() * sizeof(float2));
    if (e == cudaSuccess) e = up(h->d_mel_start, start.data(), start.size() * sizeof(int));
    if (e == cudaSuccess) e = up(h->d_mel_width, quads.data(), quads.size() * sizeof(int));
    if (e == cudaSuccess) e = up(h->d_mel_taps, taps_sl.data(), taps_sl.size() * sizeof(float));
    if (e == cudaSuccess) e = up(h->d_mel_sched, sched.data(), sched.size() * sizeof(int));
    if (e == cudaSuccess) e = up(h->d_dct, dct_f.data(), dct_f.size() * sizeof(float));
    if (e == cudaSuccess && !fused) {
        // folded-DFT tables cos / sin(2 pi h k / N), h < bins, k < bins (zero beyond), built in float64 from
        // the exactly reduced index (h k) mod N; plain per-filter taps
        const int N = p.win_samples, pitch = h->g_pitch;
        std::vector<float> ct((size_t)bins * pitch, 0.f), st((size_t)bins * pitch, 0.f);
        for (int hh = 0; hh < bins; ++hh)
            for (int k = 0; k < bins; ++k) {
                const double a = 2.0 * 3.14159265358979323846264338327950288 * (double)(((int64_t)hh * k) % N) / (double)N;
                ct[(size_t)hh * pitch + k] = (float)std::cos(a);
                st[(size_t)hh * pitch + k] = (float)std::sin(a);
            }
        std::vector<float> gw((size_t)p.n_mel * g_wpitch, 0.f);
        for (int m = 0; m < p.n_mel; ++m)
            for (int q = 0; q < g_n[m]; ++q) gw[(size_t)m * g_wpitch + q] = (float)mel_filters[(size_t)m * npts + q];
        // the same tables for the tensor-core kernel: [cos/sin][bin tile][k-block][slice] blocks of tn bins x 64
        // folded samples in the shared-memory image the MMA reads (K-major, 128-byte swizzle), each entry split in
        // two FP16 slices from its float64 value
        const int tn = h->tc_tn, blkw = tn * 32;   // 32-bit words per slice block
        h->tc_kb = (bins + tc::kTK - 1) / tc::kTK;
        std::vector<uint16_t> tt((size_t)2 * h->tc_nt * h->tc_kb * 2 * blkw * 2, 0);
        for (int par = 0; par < 2; ++par)
            for (int nt = 0; nt < h->tc_nt; ++nt)
                for (int kb = 0; kb < h->tc_kb; ++kb) {
                    uint16_t *blk = tt.data() + (((size_t)(par * h->tc_nt + nt) * h->tc_kb + kb) * 2) * blkw * 2;
                    for (int n = 0; n < tn; ++n)
                        for (int c = 0; c < tc::kTK; ++c) {
                            const int k = nt * tn + n, hh = kb * tc::kTK + c;
                            if (k >= bins || hh >= bins) continue;
                            const double a = 2.0 * 3.14159265358979323846264338327950288 * (double)(((int64_t)hh * k) % N) / (double)N;
                            double v = par == 0 ? std::cos(a) : std::sin(a);
                            const size_t at = tc::swz128(n, c >> 1) / 2 + (c & 1);
                            for (int sl = 0; sl < 2; ++sl) {
                                const __half q = __double2half(v);   // round to nearest, subnormals kept
                                uint16_t bits;
                                std::memcpy(&bits, &q, 2);
                                blk[(size_t)sl * blkw * 2 + at] = bits;
                                v -= (double)__half2float(q);
                            }
                        }
                }
        e = up(h->d_cos, ct.data(), ct.size() * sizeof(float));
        if (e == cudaSuccess) e = up(h->d_sin, st.data(), st.size() * sizeof(float));
        if (e == cudaSuccess) e = up(h->d_tc_tab, tt.data(), tt.size() * sizeof(uint16_t));
        if (e == cudaSuccess) e = up(h->d_gmel_lo, g_lo.data(), g_lo.size() * sizeof(int));
        if (e == cudaSuccess) e = up(h->d_gmel_n, g_n.data(), g_n.size() * sizeof(int));
        if (e == cudaSuccess) e = up(h->d_gmel_w, gw.data(), gw.size() * sizeof(float));
    }
    if (e == cudaSuccess) e = up(h->d_gabor, gab_f.data(), gab_f.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking);
    for (int i = 0; i < 8 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        aud_destroy(h);
        return failf(AUD_ERR_CUDA, "device setup failed: %s", cudaGetErrorString(e));
    }
    *out = h;
    return AUD_OK;
}

void aud_destroy(aud_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (DevBuf *b : {&h->d_tw, &h->d_mel_start, &h->d_mel_width, &h->d_mel_taps, &h->d_mel_sched, &h->d_dct, &h->d_gabor,
                      &h->d_rawpow, &h->d_dbg, &h->d_wave, &h->d_cos, &h->d_sin, &h->d_tc_tab, &h->d_tc_scale, &h->d_gmel_lo, &h->d_gmel_n, &h->d_gmel_w})
        b->release();
    for (auto &b : h->d_out) b.release();
    for (aud::Plan *pl : h->plans) delete pl;
    for (int i = 0; i < 8; ++i) {
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
    }
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h->stager;
    delete h;
}

int32_t aud_get_dims(const aud_handle *h, aud_dims *d) {
    if (!h || !d) return fail(AUD_ERR_INVALID, "aud_get_dims: NULL argument");
    d->segment_steps = h->p.segment_steps;
    d->n_bins = h->bins;
    d->n_mel = h->p.n_mel;
    d->n_coefs = h->p.n_coefs;
    d->gabor_len = h->gabor_len;
    return AUD_OK;
}

int32_t aud_seg_count(const aud_handle *h, int32_t n_samples) {
    if (!h) return fail(AUD_ERR_INVALID, "aud_seg_count: NULL handle");
    return (int32_t)seg_count(h->p, n_samples);
}

int64_t aud_total_segments(const aud_handle *h, const int32_t *utt_len, int32_t n_utt, int64_t *seg_base) {
    if (!h || (n_utt > 0 && !utt_len) || n_utt < 0) return fail(AUD_ERR_INVALID, "aud_total_segments: bad argument");
    int64_t tot = 0;
    for (int u = 0; u < n_utt; ++u) {
        if (seg_base) seg_base[u] = tot;
        tot += seg_count(h->p, utt_len[u]);
    }
    if (seg_base) seg_base[n_utt] = tot;
    return tot;
}

static int32_t check_batch(const aud_handle *h, const aud_batch *b, const aud_outputs *o) {
    if (!h || !b || !o) return fail(AUD_ERR_INVALID, "NULL handle / batch / outputs");
    if (b->n_utt < 0 || (b->n_utt > 0 && (!b->wave || !b->utt_offset || !b->utt_len)))
        return fail(AUD_ERR_INVALID, "batch arrays are NULL");
    if ((o->mfcc || o->deltas || o->delta_deltas) && !h->p.mfcc) return fail(AUD_ERR_INVALID, "mfcc outputs requested but Mel.MFCC is off");
    if ((o->deltas || o->delta_deltas) && !h->p.deltas) return fail(AUD_ERR_INVALID, "delta outputs requested but Mel.Deltas is off");
    if (o->gabor && h->p.gabor_nf <= 0) return fail(AUD_ERR_INVALID, "gabor output requested but no gabor filters configured");
    if (o->logpower && !h->p.comp_log_pow) return fail(AUD_ERR_INVALID, "logpower requested but CompLogPow is off");
    return AUD_OK;
}

int32_t aud_process_device(aud_handle *h, const aud_batch *b, const aud_outputs *o, void *cuda_stream) {
    int32_t rc = check_batch(h, b, o);
    if (rc != AUD_OK) return rc;
    AUD_CUDA(cudaSetDevice(h->device));
    return run_device(h, b, o, (cudaStream_t)cuda_stream);
}

int32_t aud_process_device_i16(aud_handle *h, const int16_t *wave, const int64_t *utt_offset, const int32_t *utt_len,
                               int32_t n_utt, int32_t add_samples, const aud_outputs *o, void *cuda_stream) {
    aud_batch b{reinterpret_cast<const float *>(wave), utt_offset, utt_len, n_utt, add_samples};
    int32_t rc = check_batch(h, &b, o);
    if (rc != AUD_OK) return rc;
    AUD_CUDA(cudaSetDevice(h->device));
    return run_device(h, &b, o, (cudaStream_t)cuda_stream, 1);
}

// Caller memory that is not page-locked (a Go slice, a numpy array, malloc) would make every "asynchronous" copy
// of the pipeline below a synchronous, driver-staged one.  Large unpinned buffers are therefore page-locked for the
// duration of the call (cudaHostRegister) and released before it returns -- the library never keeps a caller
// pointer.  Buffers that are already pinned (aud_host_alloc, cudaHostAlloc, the caller's own cudaHostRegister) are
// used as they are; small ones are left to the driver's staging, which is cheaper than registering them.
struct HostPin {
    void *ptr = nullptr;
    static constexpr size_t kMinBytes = 2u << 20;
    void pin(const void *p, size_t bytes, bool read_only, int device) {
        if (!p || bytes < kMinBytes || ptr) return;
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return; }
        if (at.type != cudaMemoryTypeUnregistered) return;
        unsigned flags = cudaHostRegisterDefault;
        int ro = 0;
        if (read_only && cudaDeviceGetAttribute(&ro, cudaDevAttrHostRegisterReadOnlySupported, device) == cudaSuccess && ro)
            flags |= cudaHostRegisterReadOnly;
        if (cudaHostRegister(const_cast<void *>(p), bytes, flags) == cudaSuccess) ptr = const_cast<void *>(p);
        else cudaGetLastError();   // e.g. memory that cannot be locked: the copies fall back to driver staging
    }
    ~HostPin() {
        if (!ptr) return;
        cudaDeviceSynchronize();   // no copy may still be reading or writing the range (error paths return early)
        cudaHostUnregister(ptr);
    }
};

static int32_t process_host_impl(aud_handle *h, const aud_batch *b, const aud_outputs *o, int in_i16, bool pin_caller = true) {
    const size_t esz = in_i16 ? 2 : 4;
    const int64_t amask = in_i16 ? 7 : 3;   // samples per 16 bytes - 1
    int32_t rc = check_batch(h, b, o);
    if (rc != AUD_OK) return rc;
    if (b->n_utt == 0) return AUD_OK;
    AUD_CUDA(cudaSetDevice(h->device));
    // extent of the wave buffer actually referenced
    int64_t lo = INT64_MAX, hi = INT64_MIN;
    for (int u = 0; u < b->n_utt; ++u) {
        if (b->utt_len[u] <= 0) continue;
        lo = std::min<int64_t>(lo, b->utt_offset[u]);
        hi = std::max<int64_t>(hi, b->utt_offset[u] + b->utt_len[u]);
    }
    if (lo > hi) { lo = 0; hi = 0; }
    lo &= ~amask;   // keep the device copy 16-byte congruent with the caller's buffer (TMA windows)
    const size_t wbytes = (size_t)(hi - lo) * esz;
    AUD_CUDA(h->d_wave.reserve(std::max<size_t>(wbytes, 16)));

    std::vector<int64_t> seg_base((size_t)b->n_utt + 1);
    const int64_t nseg = aud_total_segments(h, b->utt_len, b->n_utt, seg_base.data());
    const aud_params &p = h->p;
    const size_t S = p.segment_steps;
    const size_t per_seg[8] = {(size_t)p.n_mel * S, (size_t)p.n_coefs * S, (size_t)p.n_coefs * S, (size_t)p.n_coefs * S,
                               S, (size_t)h->gabor_len, (size_t)h->bins * S, (size_t)h->bins * S};
    float *host[8] = {o->mel, o->mfcc, o->deltas, o->delta_deltas, o->energy, o->gabor, o->power, o->logpower};
    float *dev[8] = {};
    // How each caller buffer travels.  Page-locked memory (aud_host_alloc, cudaHostAlloc, cudaHostRegister): copied
    // from / to directly.  Ordinary pageable memory -- a Go slice, a numpy array: staged through the handle's pinned
    // bounce buffers by copy threads (default), or page-locked for the duration of the call (option pin = 1), or left
    // to the driver (pin = 2; also what small buffers get, where either set-up costs more than it saves).
    enum { kDirect, kStaged, kLocked };
    HostPin pin_in, pin_out[8];
    bool want_stager = false;
    auto classify = [&](const void *ptr, size_t bytes, bool read_only, HostPin &pin) {
        if (!ptr || bytes < HostPin::kMinBytes || h->opt_pin == 2) return (int)kDirect;
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return (int)kDirect; }
        if (at.type != cudaMemoryTypeUnregistered) return (int)kDirect;
        if (h->opt_pin == 1) {
            if (pin_caller) pin.pin(ptr, bytes, read_only, h->device);
            return (int)kLocked;
        }
        want_stager = true;
        return (int)kStaged;
    };
    const int in_mode = classify((const char *)b->wave + lo * (int64_t)esz, wbytes, true, pin_in);
    int out_mode[8] = {};
    for (int i = 0; i < 8; ++i) {
        if (!host[i]) continue;
        AUD_CUDA(h->d_out[i].reserve(std::max<size_t>((size_t)nseg * per_seg[i] * sizeof(float), 16)));
        dev[i] = (float *)h->d_out[i].p;
        out_mode[i] = classify(host[i], (size_t)nseg * per_seg[i] * sizeof(float), false, pin_out[i]);
    }
    if (want_stager && !h->stager) {
        h->stager = new (std::nothrow) Stager();
        if (!h->stager) return fail(AUD_ERR_NOMEM, "out of host memory");
        h->stager->threads_wanted = h->opt_copy_threads;
        AUD_CUDA(h->stager->init());
    }
    const char *d_wave0 = (const char *)h->d_wave.p - lo * (int64_t)esz;   // d_wave0 + k*esz mirrors sample k of b->wave

    // Pipeline over groups of utterances: H2D of group g+1 and D2H of group g-1 ride the two copy engines
    // while group g computes.  Utterances must be laid out in ascending order for the groups' extents to be
    // disjoint; otherwise (or for small batches) a single group is used.
    bool ascending = true;
    for (int u = 1; u < b->n_utt; ++u)
        if (b->utt_offset[u] < b->utt_offset[u - 1] + std::max(0, b->utt_len[u - 1])) ascending = false;
    int groups = 1;
    if (ascending && !o->power && !o->logpower && h->opt_groups != 1)
        groups = h->opt_groups > 0 ? std::min(h->opt_groups, 8) : (int)std::min<size_t>(8, wbytes / (24u << 20));
    groups = std::max(1, std::min(groups, b->n_utt));

    int64_t prev_s0 = 0, prev_s1 = 0;
    auto drain = [&](int64_t s0, int64_t s1) -> cudaError_t {
        for (int i = 0; i < 8; ++i)
            if (host[i] && per_seg[i] > 0 && out_mode[i] == kStaged) {
                // s_d2h already waits for the group's ev_done (enqueued in stream order before this call)
                cudaError_t e = h->stager->to_host(host[i] + (size_t)s0 * per_seg[i], dev[i] + (size_t)s0 * per_seg[i],
                                                   (size_t)(s1 - s0) * per_seg[i] * sizeof(float), h->s_d2h);
                if (e != cudaSuccess) return e;
            }
        return cudaSuccess;
    };
    int u0 = 0;
    for (int g = 0; g < groups; ++g) {
        // utterances [u0, u1): about 1/groups of the samples
        int u1 = u0;
        if (g + 1 == groups) u1 = b->n_utt;
        else {
            const int64_t target = lo + (hi - lo) * (g + 1) / groups;
            while (u1 < b->n_utt && b->utt_offset[u1] + b->utt_len[u1] <= target) ++u1;
            if (u1 == u0) u1 = std::min(b->n_utt, u0 + 1);
        }
        if (u1 == u0) continue;
        int64_t glo = INT64_MAX, ghi = INT64_MIN;
        for (int u = u0; u < u1; ++u) {
            if (b->utt_len[u] <= 0) continue;
            glo = std::min<int64_t>(glo, b->utt_offset[u]);
            ghi = std::max<int64_t>(ghi, b->utt_offset[u] + b->utt_len[u]);
        }
        if (glo <= ghi) {
            char *dst = const_cast<char *>(d_wave0) + glo * (int64_t)esz;
            const char *src = (const char *)b->wave + glo * (int64_t)esz;
            const size_t nb = (size_t)(ghi - glo) * esz;
            if (in_mode == kStaged) AUD_CUDA(h->stager->to_device(dst, src, nb, h->s_h2d));
            else AUD_CUDA(cudaMemcpyAsync(dst, src, nb, cudaMemcpyHostToDevice, h->s_h2d));
        }
        AUD_CUDA(cudaEventRecord(h->ev_in[g], h->s_h2d));
        AUD_CUDA(cudaStreamWaitEvent(h->stream, h->ev_in[g], 0));
        aud_batch db = *b;
        db.wave = reinterpret_cast<const float *>(d_wave0);
        db.utt_offset = b->utt_offset + u0;
        db.utt_len = b->utt_len + u0;
        db.n_utt = u1 - u0;
        const int64_t s0 = seg_base[u0], s1 = seg_base[u1];
        aud_outputs dout{};
        float **dp = &dout.mel;
        for (int i = 0; i < 8; ++i) dp[i] = dev[i] ? dev[i] + (size_t)s0 * per_seg[i] : nullptr;
        rc = run_device(h, &db, &dout, h->stream, in_i16);
        if (rc != AUD_OK) { cudaDeviceSynchronize(); return rc; }
        AUD_CUDA(cudaEventRecord(h->ev_done[g], h->stream));
        AUD_CUDA(cudaStreamWaitEvent(h->s_d2h, h->ev_done[g], 0));
        for (int i = 0; i < 8; ++i)
            if (host[i] && s1 > s0 && per_seg[i] > 0 && out_mode[i] != kStaged)
                AUD_CUDA(cudaMemcpyAsync(host[i] + (size_t)s0 * per_seg[i], dev[i] + (size_t)s0 * per_seg[i],
                                         (size_t)(s1 - s0) * per_seg[i] * sizeof(float), cudaMemcpyDeviceToHost, h->s_d2h));
        // staged outputs of the PREVIOUS group leave now, while this group computes (the calling thread blocks in
        // the drain, so it comes after this group's launch has been enqueued)
        if (prev_s1 > prev_s0) AUD_CUDA(drain(prev_s0, prev_s1));
        prev_s0 = s0; prev_s1 = s1;
        u0 = u1;
    }
    if (prev_s1 > prev_s0) AUD_CUDA(drain(prev_s0, prev_s1));
    AUD_CUDA(cudaStreamSynchronize(h->s_d2h));
    AUD_CUDA(cudaStreamSynchronize(h->stream));
    AUD_CUDA(cudaStreamSynchronize(h->s_h2d));
    return AUD_OK;
}

int32_t aud_process_host(aud_handle *h, const aud_batch *b, const aud_outputs *o) { return process_host_impl(h, b, o, 0); }

int32_t aud_process_host_i16(aud_handle *h, const int16_t *wave, const int64_t *utt_offset, const int32_t *utt_len,
                             int32_t n_utt, int32_t add_samples, const aud_outputs *o) {
    aud_batch b{reinterpret_cast<const float *>(wave), utt_offset, utt_len, n_utt, add_samples};
    return process_host_impl(h, &b, o, 1);
}

// One batch over several GPUs of the box: contiguous blocks of utterances with about equal numbers of
// segments, one host thread and one handle per GPU, every GPU writing its own disjoint range of the caller's output
// tensors.  No collective, no device-to-device traffic.
static int32_t process_host_multi_impl(aud_handle *const *hs, int32_t n, const aud_batch *b, const aud_outputs *o, int in_i16) {
    if (!hs || n < 1 || !b || !o) return fail(AUD_ERR_INVALID, "aud_process_host_multi: NULL / empty handle list, batch or outputs");
    for (int g = 0; g < n; ++g) {
        if (!hs[g]) return fail(AUD_ERR_INVALID, "aud_process_host_multi: NULL handle");
        if (std::memcmp(&hs[g]->p, &hs[0]->p, sizeof(aud_params)) != 0 || hs[g]->gabor_len != hs[0]->gabor_len)
            return fail(AUD_ERR_INVALID, "aud_process_host_multi: the handles were created with different parameters");
        for (int k = 0; k < g; ++k)
            if (hs[k] == hs[g]) return fail(AUD_ERR_INVALID, "aud_process_host_multi: the same handle is listed twice");
    }
    int32_t rc = check_batch(hs[0], b, o);
    if (rc != AUD_OK) return rc;
    if (b->n_utt == 0) return AUD_OK;
    if (n == 1) return process_host_impl(hs[0], b, o, in_i16);
    aud_handle *h0 = hs[0];
    const aud_params &p = h0->p;
    std::vector<int64_t> seg_base((size_t)b->n_utt + 1);
    const int64_t nseg = aud_total_segments(h0, b->utt_len, b->n_utt, seg_base.data());
    if (nseg < 0) return (int32_t)nseg;
    // block g ends where the running segment count first reaches (g + 1) / n of the total
    std::vector<int> cut((size_t)n + 1, 0);
    cut[n] = b->n_utt;
    for (int g = 1; g < n; ++g) {
        const int64_t target = nseg * g / n;
        int u = (int)(std::lower_bound(seg_base.begin(), seg_base.end(), target) - seg_base.begin());
        cut[g] = std::min(b->n_utt, std::max(u, cut[g - 1]));
    }
    const size_t S = p.segment_steps, esz = in_i16 ? 2 : 4;
    const size_t per_seg[8] = {(size_t)p.n_mel * S, (size_t)p.n_coefs * S, (size_t)p.n_coefs * S, (size_t)p.n_coefs * S,
                               S, (size_t)h0->gabor_len, (size_t)h0->bins * S, (size_t)h0->bins * S};
    // page-lock the caller's buffers once, here: the per-GPU calls then see pinned memory
    HostPin pin_in, pin_out[8];
    {
        int64_t lo = INT64_MAX, hi = INT64_MIN;
        for (int u = 0; u < b->n_utt; ++u) {
            if (b->utt_len[u] <= 0) continue;
            lo = std::min<int64_t>(lo, b->utt_offset[u]);
            hi = std::max<int64_t>(hi, b->utt_offset[u] + b->utt_len[u]);
        }
        // (only with option pin = 1: by default every GPU's thread stages its own block through its bounce buffers)
        if (h0->opt_pin == 1) {
            if (lo < hi) pin_in.pin((const char *)b->wave + lo * (int64_t)esz, (size_t)(hi - lo) * esz, true, h0->device);
            float *const *src = &o->mel;
            for (int i = 0; i < 8; ++i)
                if (src[i]) pin_out[i].pin(src[i], (size_t)nseg * per_seg[i] * sizeof(float), false, h0->device);
        }
    }
    std::vector<int32_t> rcs((size_t)n, AUD_OK);
    std::vector<std::string> msgs((size_t)n);
    std::vector<std::thread> th;
    for (int g = 0; g < n; ++g) {
        if (cut[g + 1] == cut[g]) continue;
        th.emplace_back([&, g]() {
            aud_batch sb = *b;
            sb.utt_offset = b->utt_offset + cut[g];
            sb.utt_len = b->utt_len + cut[g];
            sb.n_utt = cut[g + 1] - cut[g];
            aud_outputs so{};
            float *const *src = &o->mel;
            float **dst = &so.mel;
            for (int i = 0; i < 8; ++i) dst[i] = src[i] ? src[i] + (size_t)seg_base[cut[g]] * per_seg[i] : nullptr;
            rcs[g] = process_host_impl(hs[g], &sb, &so, in_i16, false);
            if (rcs[g] != AUD_OK) msgs[g] = g_last_error;   // the error text lives in this worker's thread-local slot
        });
    }
    for (auto &t : th) t.join();
    for (int g = 0; g < n; ++g)
        if (rcs[g] != AUD_OK) return failf(rcs[g], "GPU %d (handle %d): %s", hs[g]->device, g, msgs[g].c_str());
    return AUD_OK;
}

int32_t aud_process_host_multi(aud_handle *const *handles, int32_t n_handles, const aud_batch *b, const aud_outputs *o) {
    return process_host_multi_impl(handles, n_handles, b, o, 0);
}

int32_t aud_process_host_multi_i16(aud_handle *const *handles, int32_t n_handles, const int16_t *wave, const int64_t *utt_offset,
                                   const int32_t *utt_len, int32_t n_utt, int32_t add_samples, const aud_outputs *o) {
    aud_batch b{reinterpret_cast<const float *>(wave), utt_offset, utt_len, n_utt, add_samples};
    return process_host_multi_impl(handles, n_handles, &b, o, 1);
}

int32_t aud_gabor_convolve(int32_t device, const float *mel, int32_t n, int32_t n_mel, int32_t steps, const double *filters,
                           int32_t nf, int32_t size_x, int32_t size_y, int32_t stride_x, int32_t stride_y, double gain,
                           int32_t out_dims, const int32_t *out_shape, int32_t by_time, float *out) {
    if (!mel || !filters || !out || !out_shape) return fail(AUD_ERR_INVALID, "aud_gabor_convolve: NULL argument");
    if (n < 0 || n_mel < 1 || steps < 1 || nf < 1) return fail(AUD_ERR_INVALID, "aud_gabor_convolve: non-positive tensor count / shape / filter count");
    aud_params p{};
    p.n_mel = n_mel; p.segment_steps = steps;
    p.gabor_nf = nf; p.gabor_size_x = size_x; p.gabor_size_y = size_y; p.gabor_stride_x = stride_x; p.gabor_stride_y = stride_y;
    p.gabor_gain = gain; p.gabor_out_dims = out_dims; p.gabor_by_time = by_time;
    for (int d = 0; d < 4; ++d) p.gabor_shape[d] = (d < out_dims && out_dims <= 4) ? out_shape[d] : 0;
    GaborGeom g{};
    int32_t rc = gabor_geometry(p, true, &g);
    if (rc != AUD_OK) return rc;
    if (!g.on || n == 0 || g.len == 0) return AUD_OK;   // Convolve logs and returns without writing (gabor.go:226-229)
    cudaError_t e = cudaSetDevice(device);
    cudaDeviceProp prop{};
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return failf(AUD_ERR_CUDA, "no usable CUDA device %d: %s (this library has no CPU fallback)", device, cudaGetErrorString(e));

    KParams kp{};
    kp.S = steps; kp.n_mel = n_mel;
    kp.g_on = 1; kp.g_keep = 1; kp.g_nf = nf; kp.g_sx = size_x; kp.g_sy = size_y; kp.g_stx = stride_x; kp.g_sty = stride_y;
    kp.g_dims = out_dims; kp.g_by_time = by_time; kp.g_nt = g.nt; kp.g_nfy = g.nfy; kp.g_tmaxstrides = g.tmaxstrides;
    kp.g_len = (int)g.len; kp.g_gain = (float)gain;
    if (out_dims == 2) kp.g_str0 = out_shape[1];
    else { kp.g_str0 = out_shape[1] * out_shape[2] * out_shape[3]; kp.g_str1 = out_shape[2] * out_shape[3]; kp.g_str2 = out_shape[3]; }
    kp.gw_floats = size_x * size_y * ((nf + 7) / 8 * 8);
    const size_t MS = (size_t)n_mel * steps;
    const size_t smem = (((MS + 3) & ~(size_t)3) + (((size_t)g.len + 3) & ~(size_t)3) + (size_t)kp.gw_floats) * sizeof(float) + 16;
    if (smem > prop.sharedMemPerBlockOptin)
        return failf(AUD_ERR_UNSUPPORTED, "aud_gabor_convolve: tensor does not fit in shared memory (%zu bytes needed)", smem);
    // scratch of the stand-alone operator: grow-only device buffers kept per host thread and device, and the narrowed
    // filters kept as long as the caller passes the same ones (interactive callers re-run one FilterSet over many tensors)
    struct Scratch {
        int device = -1;
        DevBuf d_mel, d_w, d_out;
        std::vector<double> filters;
        ~Scratch() {   // thread exit; after CUDA has shut down the calls fail harmlessly
            if (device >= 0 && cudaSetDevice(device) == cudaSuccess) { d_mel.release(); d_w.release(); d_out.release(); }
        }
    };
    static thread_local Scratch sc;
    if (sc.device != device) {
        if (sc.device >= 0 && cudaSetDevice(sc.device) == cudaSuccess) { sc.d_mel.release(); sc.d_w.release(); sc.d_out.release(); }
        cudaSetDevice(device);
        sc.filters.clear();
        sc.device = device;
    }
    DevBuf &d_mel = sc.d_mel, &d_w = sc.d_w, &d_out = sc.d_out;
    auto done = [&](int32_t r) { return r; };
    const size_t nw = (size_t)nf * size_x * size_y;
    const bool same_filters = sc.filters.size() == nw && std::equal(sc.filters.begin(), sc.filters.end(), filters);
    e = d_mel.reserve((size_t)n * MS * sizeof(float));
    if (e == cudaSuccess) e = d_w.reserve(nw * sizeof(float));
    if (e == cudaSuccess) e = d_out.reserve((size_t)n * g.len * sizeof(float));
    if (e == cudaSuccess && !same_filters) {
        std::vector<float> wf(nw);
        for (size_t i = 0; i < nw; ++i) wf[i] = (float)filters[i];
        e = cudaMemcpy(d_w.p, wf.data(), nw * sizeof(float), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) sc.filters.assign(filters, filters + nw);
        else sc.filters.clear();
    }
    if (e == cudaSuccess) e = cudaMemcpy(d_mel.p, mel, (size_t)n * MS * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_out.p, out, (size_t)n * g.len * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gabor_convolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        kp.gabor = (const float *)d_w.p;
        kp.o_gabor = (float *)d_out.p;
        gabor_convolve_kernel<<<(unsigned)n, 128, smem>>>(kp, (const float *)d_mel.p);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out.p, (size_t)n * g.len * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return done(failf(AUD_ERR_CUDA, "aud_gabor_convolve failed: %s", cudaGetErrorString(e)));
    return done(AUD_OK);
}

// ---------------------------------------------------------------------------------------------------------------
// Per-step operators for callers that drive dft.Filter / mel.FilterDft / mel.CepstrumDct themselves
// (examples/gaborview/gbv.go:545-559, 627-641): each call covers every step of one segment.
// ---------------------------------------------------------------------------------------------------------------
int32_t aud_dft_filter(int32_t device, const aud_dft_params *dp, const float *windows, int32_t n_steps, int32_t win_samples,
                       float *power_segment, float *log_power_segment) {
    if (!dp || !windows) return fail(AUD_ERR_INVALID, "aud_dft_filter: NULL argument");
    if (n_steps < 1 || win_samples < 2) return fail(AUD_ERR_INVALID, "aud_dft_filter: non-positive step count / window length");
    if (!power_segment && !log_power_segment) return AUD_OK;
    if (log_power_segment && !dp->comp_log_pow) return fail(AUD_ERR_INVALID, "aud_dft_filter: log power requested but CompLogPow is off");
    // The steps of one segment are the frames of one "utterance" whose hop equals the window: the batch pipeline then
    // does exactly Filter's work -- FFT, |X|^2, Prev/Cur smoothing from step 1 on, ln(p + LogOffSet) -- for all of them.
    aud_params p{};
    p.sample_rate = 16000;   // not used by the transform
    p.win_samples = win_samples; p.step_samples = win_samples;
    p.segment_samples = n_steps * win_samples; p.stride_samples = n_steps * win_samples;
    p.segment_steps = n_steps; p.border_steps = 0;
    p.comp_log_pow = dp->comp_log_pow; p.log_min = dp->log_min; p.log_offset = dp->log_offset;
    p.prev_smooth = dp->prev_smooth; p.cur_smooth = dp->cur_smooth;
    p.n_mel = 1; p.mel_log_min = -10.0;   // a one-filter bank over bin 0: the mel stage is not asked for anything
    p.n_coefs = 1;
    const int32_t bin_pts[3] = {0, 0, 0};
    const double filt[3] = {0.0, 0.0, 0.0};
    aud_handle *h = nullptr;
    int32_t rc = aud_create(&p, bin_pts, filt, nullptr, nullptr, device, &h);
    if (rc != AUD_OK) return rc;
    const int64_t off = 0;
    const int32_t len = n_steps * win_samples;
    aud_batch b{windows, &off, &len, 1, 0};
    aud_outputs o{};
    o.power = power_segment;
    o.logpower = log_power_segment;
    rc = aud_process_host(h, &b, &o);
    aud_destroy(h);
    return rc;
}

int32_t aud_mel_filter_dft(int32_t device, const aud_mel_params *mp, const int32_t *bin_pts, const double *filters,
                           const float *power_segment, int32_t n_bins, int32_t n_steps, float *mel_segment) {
    if (!mp || !bin_pts || !filters || !power_segment || !mel_segment) return fail(AUD_ERR_INVALID, "aud_mel_filter_dft: NULL argument");
    const int nf = mp->n_filters;
    if (nf < 1 || n_bins < 1 || n_steps < 1) return fail(AUD_ERR_INVALID, "aud_mel_filter_dft: non-positive filter / bin / step count");
    for (int m = 0; m < nf; ++m) {
        const int lo = bin_pts[m], hi = bin_pts[m + 2];
        if (lo < 0 || hi >= n_bins) return fail(AUD_ERR_PANIC, "mel BinPts outside the power spectrum (reference panics)");
        if (hi >= lo && (int64_t)m * (nf + 2) + (hi - lo + 1) > (int64_t)nf * (nf + 2))
            return fail(AUD_ERR_PANIC, "mel filter table index out of range (reference panics)");
    }
    AUD_CUDA(cudaSetDevice(device));
    std::vector<float> ff((size_t)nf * (nf + 2));
    for (size_t i = 0; i < ff.size(); ++i) ff[i] = (float)filters[i];
    DevBuf d_bp, d_f, d_p, d_m;
    auto done = [&](int32_t r) { d_bp.release(); d_f.release(); d_p.release(); d_m.release(); return r; };
    cudaError_t e = d_bp.reserve((size_t)(nf + 2) * sizeof(int));
    if (e == cudaSuccess) e = d_f.reserve(ff.size() * sizeof(float));
    if (e == cudaSuccess) e = d_p.reserve((size_t)n_bins * n_steps * sizeof(float));
    if (e == cudaSuccess) e = d_m.reserve((size_t)nf * n_steps * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(d_bp.p, bin_pts, (size_t)(nf + 2) * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_f.p, ff.data(), ff.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_p.p, power_segment, (size_t)n_bins * n_steps * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        MelOpParams q{n_bins, n_steps, nf, (float)mp->log_off, (float)mp->log_min, mp->renorm, (float)mp->renorm_min,
                      (float)mp->renorm_scale, (const int *)d_bp.p, (const float *)d_f.p, (const float *)d_p.p, (float *)d_m.p};
        mel_filter_dft_kernel<<<(nf * n_steps + 127) / 128, 128>>>(q);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(mel_segment, d_m.p, (size_t)nf * n_steps * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return done(failf(AUD_ERR_CUDA, "aud_mel_filter_dft failed: %s", cudaGetErrorString(e)));
    return done(AUD_OK);
}

int32_t aud_cepstrum_dct(int32_t device, const float *mel_segment, int32_t n_filters, int32_t n_steps, int32_t n_coefs,
                         const double *dct, float *mfcc_segment) {
    if (!mel_segment || !mfcc_segment) return fail(AUD_ERR_INVALID, "aud_cepstrum_dct: NULL argument");
    if (n_filters < 1 || n_steps < 1) return fail(AUD_ERR_INVALID, "aud_cepstrum_dct: non-positive filter / step count");
    if (n_coefs < 1 || n_coefs > n_filters) return fail(AUD_ERR_PANIC, "NCoefs must be in 1..NFilters (reference indexes past the DCT output)");
    AUD_CUDA(cudaSetDevice(device));
    std::vector<double> dd;
    if (!dct) {
        dd.resize((size_t)n_coefs * n_filters);
        aud_dct1_matrix(n_filters, n_coefs, dd.data());
        dct = dd.data();
    }
    std::vector<float> df((size_t)n_coefs * n_filters);
    for (size_t i = 0; i < df.size(); ++i) df[i] = (float)dct[i];
    DevBuf d_d, d_m, d_c;
    auto done = [&](int32_t r) { d_d.release(); d_m.release(); d_c.release(); return r; };
    cudaError_t e = d_d.reserve(df.size() * sizeof(float));
    if (e == cudaSuccess) e = d_m.reserve((size_t)n_filters * n_steps * sizeof(float));
    if (e == cudaSuccess) e = d_c.reserve((size_t)n_coefs * n_steps * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(d_d.p, df.data(), df.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_m.p, mel_segment, (size_t)n_filters * n_steps * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        cepstrum_dct_kernel<<<(n_coefs * n_steps + 127) / 128, 128>>>((const float *)d_m.p, (const float *)d_d.p, n_filters, n_steps,
                                                                    n_coefs, (float *)d_c.p);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(mfcc_segment, d_c.p, (size_t)n_coefs * n_steps * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return done(failf(AUD_ERR_CUDA, "aud_cepstrum_dct failed: %s", cudaGetErrorString(e)));
    return done(AUD_OK);
}

void *aud_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        fail(AUD_ERR_CUDA, "cudaHostAlloc failed");
        return nullptr;
    }
    return p;
}
void aud_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int64_t aud_launch_count(const aud_handle *h) { return h ? h->launches : 0; }

int32_t aud_set_option(aud_handle *h, const char *name, int64_t value) {
    if (!h || !name) return fail(AUD_ERR_INVALID, "aud_set_option: NULL argument");
    const std::string n(name);
    if (n == "job_segs") h->opt_job_segs = (int)value;
    else if (n == "warps") h->opt_warps = (int)value;
    else if (n == "ctas") h->opt_ctas = (int)value;
    else if (n == "epi") h->opt_epi = (int)value;
    else if (n == "groups") h->opt_groups = (int)value;
    else if (n == "pin") h->opt_pin = (int)value;
    else if (n == "dft_tc") h->opt_dft_tc = value ? 1 : 0;
    else if (n == "copy_threads") h->opt_copy_threads = (int)std::max<int64_t>(0, std::min<int64_t>(64, value));
    else return failf(AUD_ERR_INVALID, "unknown option '%s'", name);

    return AUD_OK;
}

}  // extern "C"
