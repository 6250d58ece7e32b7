// aud_kwta.cu -- the step after agabor.Convolve in SndEnv.ApplyGabor (sound/sndenv.go:481-497): neighbour inhibition
// (ApplyNeighInhib, :303-311) and k-winners-take-all (ApplyKwta, :314-323) on the gabor output tensors.
//
// The algorithms live in third-party packages that are NOT in the reference tree (go.mod:8-9: emer/vision v1.1.15
// kwta, emer/leabra v1.1.48 fffb + nxx1), so this is a restatement of the published FFFB inhibition / noisy-XX1
// activation functions from memory of those versions: PARITY UNPINNED (the tests hold it against a float32 numpy twin of the same equations).
// Float32 throughout and the same order of operations as the Go loops: one thread walks one SndEnv's tensors in call
// order, because KWTAPool keeps per-pool feedback inhibition (SndEnv.Inhibs) from one call to the next; tensors of
// different sequences (utterances) run in parallel.  A post-processing stage, not part of the fused kernel's timed path.
#include <cuda_runtime.h>

#include <vector>

#include "auditory_b200.h"
#include "aud_internal.h"

namespace aud {

struct KwtaDev {
    aud_kwta_params p;
    // derived (kwta.KWTA.Update, nxx1.Params.Update, fffb.Params.Update)
    float sig_gain_nvar, sig_mult_eff, sig_val_at0, interp_val;
    float lay_fbdt, pool_fbdt, act_dt;
    float erev_sub_thr_i, erev_sub_thr_l, thr_sub_erev_e;
};

__host__ __device__ inline float xx1(const KwtaDev &k, float x) {
    x *= k.p.xx1_gain;
    return x / (x + 1.f);
}
__host__ __device__ inline float xx1_gain_cor(const KwtaDev &k, float x) {
    const float fact = (k.p.xx1_gain_cor_range - (x / k.p.xx1_nvar)) / k.p.xx1_gain_cor_range;
    if (fact < 0.f) return xx1(k, x);
    const float new_gain = k.p.xx1_gain * (1.f - k.p.xx1_gain_cor * fact);
    x *= new_gain;
    return x / (x + 1.f);
}
__device__ inline float noisy_xx1(const KwtaDev &k, float x) {
    if (x < 0.f) return k.sig_mult_eff / (1.f + expf(-(x * k.sig_gain_nvar)));
    if (x < k.p.xx1_interp_range) {
        const float interp = 1.f - ((k.p.xx1_interp_range - x) / k.p.xx1_interp_range);
        return k.sig_val_at0 + interp * k.interp_val;
    }
    return xx1_gain_cor(k, x);
}
__device__ inline float ff_inhib(const aud_fffb_params &fb, float avg_ge, float max_ge) {
    const float ff_netin = avg_ge + fb.max_vs_avg * (max_ge - avg_ge);
    return ff_netin > fb.ff0 ? fb.ff * (ff_netin - fb.ff0) : 0.f;
}
struct Inhib {   // fffb.Inhib
    float fbi, gi, ge_avg, ge_max, act_avg;
};
__device__ inline void fffb(const aud_fffb_params &fb, float fbdt, Inhib &inh) {
    if (!fb.on) { inh.fbi = 0.f; inh.gi = 0.f; return; }
    const float ffi = ff_inhib(fb, inh.ge_avg, inh.ge_max);
    const float fbi = fb.fb * inh.act_avg;
    inh.fbi += fbdt * (fbi - inh.fbi);
    inh.gi = fb.gi * (ffi + inh.fbi);
}
__device__ inline float ge_thr_from_g(const KwtaDev &k, float gi) {
    return (k.p.gbar_i * gi * k.erev_sub_thr_i + k.p.gbar_l * k.erev_sub_thr_l) / k.thr_sub_erev_e;
}

constexpr int kMaxPools = 64;   // pools per tensor in KWTAPool mode (PoolsY * PoolsX)

// One thread per sequence of tensors.  raw / ext / act: [n][len].  dims == 4: shape = {layY, layX, plY, plX}.
__global__ void kwta_kernel(const KwtaDev K, const float *raw, float *ext, float *act, const long long *seq_base, int n_seq,
                            int len, int lay_y, int lay_x, int pl_y, int pl_x, int four_d) {
    const int sq = blockIdx.x * blockDim.x + threadIdx.x;
    if (sq >= n_seq) return;
    const aud_kwta_params &p = K.p;
    const int lay_n = four_d ? lay_y * lay_x : 1, pl_n = four_d ? pl_y * pl_x : len;
    // SndEnv.Inhibs: every pool's fffb.Inhib is kept from call to call -- its feedback inhibition FBi and the average
    // activation its last iteration left (only the Ge statistics are recomputed at the start of a call)
    float pool_fbi[kMaxPools], pool_act_avg[kMaxPools];
    for (int i = 0; i < kMaxPools; ++i) pool_fbi[i] = pool_act_avg[i] = 0.f;
    for (long long t = seq_base[sq]; t < seq_base[sq + 1]; ++t) {
        const float *r = raw + (size_t)t * len;
        float *a = act + (size_t)t * len;
        float *e = ext ? ext + (size_t)t * len : nullptr;
        // ApplyNeighInhib (sndenv.go:303-311): NeighInhib.Inhib4 or zeros
        if (e) {
            for (int i = 0; i < len; ++i) e[i] = 0.f;
            if (p.neigh_on && four_d) {
                const int ox[4] = {0, -1, 1, -1}, oy[4] = {1, 1, 0, -1};
                for (int ly = 0; ly < lay_y; ++ly)
                    for (int lx = 0; lx < lay_x; ++lx)
                        for (int py = 0; py < pl_y; ++py)
                            for (int ang = 0; ang < 4 && ang < pl_x; ++ang) {
                                float gi = 0.f;
                                for (int sgn = 1; sgn >= -1; sgn -= 2) {
                                    const int nx = lx + sgn * ox[ang], ny = ly + sgn * oy[ang];
                                    if (nx >= 0 && nx < lay_x && ny >= 0 && ny < lay_y)
                                        gi = fmaxf(gi, p.neigh_gi * r[((ny * lay_x + nx) * pl_y + py) * pl_x + ang]);
                                }
                                e[((ly * lay_x + lx) * pl_y + py) * pl_x + ang] = gi;
                            }
            }
        }
        // ApplyKwta (sndenv.go:314-323): GborKwta starts as a copy of GborOutput
        for (int i = 0; i < len; ++i) a[i] = r[i];
        if (!p.on) continue;
        if (!p.pool_mode || !four_d) {
            // KWTA.KWTALayer: one level of inhibition over the whole tensor, fresh state every call
            Inhib inh{0.f, 0.f, 0.f, -3.4028235e38f, 0.f};
            float acc = 0.f;
            for (int i = 0; i < len; ++i) { acc += r[i]; inh.ge_max = fmaxf(inh.ge_max, r[i]); }
            inh.ge_avg = len > 0 ? acc / (float)len : acc;
            for (int cy = 0; cy < p.iters; ++cy) {
                fffb(p.lay_fffb, K.lay_fbdt, inh);
                float max_del = 0.f, sum = 0.f;
                for (int i = 0; i < len; ++i) {
                    const float gi = e ? inh.gi + e[i] : inh.gi;
                    const float nw = noisy_xx1(K, r[i] * p.gbar_e - ge_thr_from_g(K, gi));
                    const float del = K.act_dt * (nw - a[i]);
                    a[i] += del;
                    max_del = fmaxf(max_del, fabsf(del));
                    sum += a[i];
                }
                inh.act_avg = len > 0 ? sum / (float)len : sum;
                if (cy > 2 && max_del < p.del_act_thr) break;
            }
        } else {
            // KWTA.KWTAPool: layer inhibition over the pools' averages, pool inhibition inside each pool
            Inhib lay{0.f, 0.f, 0.f, -3.4028235e38f, 0.f};
            float pool_ge_avg[kMaxPools], pool_ge_max[kMaxPools];
            float lacc = 0.f;
            for (int pi = 0; pi < lay_n; ++pi) {
                float acc = 0.f, mx = -3.4028235e38f;
                for (int ui = 0; ui < pl_n; ++ui) { acc += r[pi * pl_n + ui]; mx = fmaxf(mx, r[pi * pl_n + ui]); }
                pool_ge_avg[pi] = pl_n > 0 ? acc / (float)pl_n : acc;
                pool_ge_max[pi] = mx;
                lacc += pool_ge_avg[pi];
                lay.ge_max = fmaxf(lay.ge_max, pool_ge_avg[pi]);
            }
            lay.ge_avg = lay_n > 0 ? lacc / (float)lay_n : lacc;
            for (int cy = 0; cy < p.iters; ++cy) {
                fffb(p.lay_fffb, K.lay_fbdt, lay);
                float max_del = 0.f, lsum = 0.f;
                for (int pi = 0; pi < lay_n; ++pi) {
                    Inhib pl{pool_fbi[pi], 0.f, pool_ge_avg[pi], pool_ge_max[pi], pool_act_avg[pi]};
                    fffb(p.pool_fffb, K.pool_fbdt, pl);
                    pool_fbi[pi] = pl.fbi;
                    const float gi_pool = fmaxf(lay.gi, pl.gi);
                    float sum = 0.f;
                    for (int ui = 0; ui < pl_n; ++ui) {
                        const int idx = pi * pl_n + ui;
                        float gi = gi_pool;
                        if (e) gi = fmaxf(gi, p.pool_fffb.gi * ff_inhib(p.pool_fffb, e[idx], e[idx]));
                        const float nw = noisy_xx1(K, r[idx] * p.gbar_e - ge_thr_from_g(K, gi));
                        const float del = K.act_dt * (nw - a[idx]);
                        a[idx] += del;
                        max_del = fmaxf(max_del, fabsf(del));
                        sum += a[idx];
                    }
                    pool_act_avg[pi] = pl_n > 0 ? sum / (float)pl_n : sum;
                    lsum += pool_act_avg[pi];
                }
                lay.act_avg = lay_n > 0 ? lsum / (float)lay_n : lsum;
                if (cy > 2 && max_del < p.del_act_thr) break;
            }
        }
    }
}

}  // namespace aud

using namespace aud;

extern "C" {

AUD_API void aud_kwta_defaults(aud_kwta_params *p) {
    if (!p) return;
    *p = aud_kwta_params{};
    p->on = 1; p->iters = 20; p->del_act_thr = 0.005f;                       // kwta.KWTA.Defaults
    const aud_fffb_params fb{1, 1.8f, 1.f, 1.f, 1.4f, 0.f, 0.1f};            // fffb.Params.Defaults
    p->lay_fffb = fb; p->pool_fffb = fb; p->pool_fffb.gi = 2.0f;
    p->xx1_thr = 0.5f; p->xx1_gain = 80.f; p->xx1_nvar = 0.01f;              // nxx1 defaults with kwta's Gain / NVar
    p->xx1_vm_act_thr = 0.01f; p->xx1_sig_mult = 0.33f; p->xx1_sig_mult_pow = 0.8f; p->xx1_sig_gain = 3.0f;
    p->xx1_interp_range = 0.01f; p->xx1_gain_cor_range = 10.f; p->xx1_gain_cor = 0.1f;
    p->act_tau = 3.f;
    p->gbar_e = 0.5f; p->gbar_l = 0.1f; p->gbar_i = 1.0f; p->gbar_k = 1.0f;
    p->erev_e = 1.0f; p->erev_l = 0.3f; p->erev_i = 0.25f; p->erev_k = 0.25f;
    p->pool_mode = 0;
    p->neigh_on = 0; p->neigh_gi = 0.6f;                                     // kwta.NeighInhib.Defaults (off in SndEnv.Defaults)
}

AUD_API int32_t aud_apply_kwta(int32_t device, const aud_kwta_params *kp, const float *gabor, int32_t n_tensors, int32_t dims,
                       const int32_t *shape, const int64_t *seq_base, int32_t n_seq, float *ext_gi, float *kwta) {
    if (!kp || !gabor || !shape || !kwta) return fail(AUD_ERR_INVALID, "aud_apply_kwta: NULL argument");
    if (n_tensors < 0 || (dims != 2 && dims != 4)) return fail(AUD_ERR_INVALID, "aud_apply_kwta: tensors must have 2 or 4 dimensions");
    int64_t len = 1;
    for (int d = 0; d < dims; ++d) {
        if (shape[d] < 1) return fail(AUD_ERR_INVALID, "aud_apply_kwta: non-positive dimension");
        len *= shape[d];
    }
    if (n_tensors == 0) return AUD_OK;
    const bool four_d = dims == 4;
    if (kp->on && kp->pool_mode && !four_d) return fail(AUD_ERR_PANIC, "KWTAPool needs a 4-D tensor (the reference indexes Dim(2), Dim(3))");
    if (kp->on && kp->pool_mode && (int64_t)shape[0] * shape[1] > kMaxPools)
        return failf(AUD_ERR_UNSUPPORTED, "aud_apply_kwta: more than %d pools per tensor", kMaxPools);
    if (kp->on && (kp->iters < 0 || kp->act_tau == 0.f || kp->lay_fffb.fb_tau == 0.f || kp->pool_fffb.fb_tau == 0.f))
        return fail(AUD_ERR_INVALID, "aud_apply_kwta: zero time constant / negative iteration count");
    // sequences: KWTAPool carries per-pool state from call to call of one SndEnv; every other mode is stateless,
    // so each tensor is its own sequence and they all run in parallel
    std::vector<long long> sb;
    const bool stateful = kp->on && kp->pool_mode;
    if (stateful && seq_base) {
        if (n_seq < 1 || seq_base[0] != 0 || seq_base[n_seq] != n_tensors) return fail(AUD_ERR_INVALID, "aud_apply_kwta: seq_base must run from 0 to n_tensors");
        for (int i = 0; i <= n_seq; ++i) {
            if (i && seq_base[i] < seq_base[i - 1]) return fail(AUD_ERR_INVALID, "aud_apply_kwta: seq_base must not decrease");
            sb.push_back(seq_base[i]);
        }
    } else if (stateful) {
        sb = {0, n_tensors};
    } else {
        sb.resize((size_t)n_tensors + 1);
        for (int i = 0; i <= n_tensors; ++i) sb[i] = i;
    }
    KwtaDev K{};
    K.p = *kp;
    K.sig_gain_nvar = kp->xx1_sig_gain / kp->xx1_nvar;
    K.sig_mult_eff = kp->xx1_sig_mult * powf(kp->xx1_gain * kp->xx1_nvar, kp->xx1_sig_mult_pow);
    K.sig_val_at0 = 0.5f * K.sig_mult_eff;
    K.interp_val = xx1_gain_cor(K, kp->xx1_interp_range) - K.sig_val_at0;
    K.lay_fbdt = 1.f / kp->lay_fffb.fb_tau;
    K.pool_fbdt = 1.f / kp->pool_fffb.fb_tau;
    K.act_dt = 1.f / kp->act_tau;
    K.erev_sub_thr_i = kp->erev_i - kp->xx1_thr;
    K.erev_sub_thr_l = kp->erev_l - kp->xx1_thr;
    K.thr_sub_erev_e = kp->xx1_thr - kp->erev_e;

    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return failf(AUD_ERR_CUDA, "no usable CUDA device %d: %s (this library has no CPU fallback)", device, cudaGetErrorString(e));
    const size_t bytes = (size_t)n_tensors * len * sizeof(float);
    float *d_raw = nullptr, *d_ext = nullptr, *d_act = nullptr;
    long long *d_sb = nullptr;
    auto done = [&](int32_t r) { cudaFree(d_raw); cudaFree(d_ext); cudaFree(d_act); cudaFree(d_sb); return r; };
    e = cudaMalloc(&d_raw, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&d_act, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&d_ext, bytes);   // ExtGi is an input of the kwta step even when the caller does not want it back
    if (e == cudaSuccess) e = cudaMalloc(&d_sb, sb.size() * sizeof(long long));
    if (e == cudaSuccess) e = cudaMemcpy(d_raw, gabor, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_sb, sb.data(), sb.size() * sizeof(long long), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        const int ns = (int)sb.size() - 1;
        kwta_kernel<<<(ns + 63) / 64, 64>>>(K, d_raw, d_ext, d_act, d_sb, ns, (int)len, four_d ? shape[0] : 1, four_d ? shape[1] : 1,
                                            four_d ? shape[2] : 1, four_d ? shape[3] : 1, four_d ? 1 : 0);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(kwta, d_act, bytes, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && ext_gi) e = cudaMemcpy(ext_gi, d_ext, bytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return done(failf(AUD_ERR_CUDA, "aud_apply_kwta failed: %s", cudaGetErrorString(e)));
    return done(AUD_OK);
}

}  // extern "C"
