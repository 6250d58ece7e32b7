"""GPU parity of the general-window-length path (WinSamples != 400: other sample rates and window
durations) against the float64 oracle, through the same C-ABI and the same tolerances as
the fused 16 kHz path."""
import numpy as np
import pytest

import auditory_b200 as ab
from auditory_b200 import synth
from oracle import c_oracle
from util import RTOL_GABOR, RTOL_LOG, assert_close
from test_gpu_parity import compare

pytestmark = pytest.mark.gpu
ALL = ["mel", "energy", "mfcc", "deltas", "delta_deltas", "gabor", "power", "logpower"]


def envs(sr, n_filters=32, hi_hz=None, mfcc=True, deltas=True, gabor=True, prev=0.0, win_ms=25.0, out4d=True, by_time=False):
    hi_hz = min(8000.0, sr / 2.0) if hi_hz is None else hi_hz   # Mel.Defaults: HiHz = 8000 (wider banks panic in the reference)
    se = ab.SndEnv(device=0)
    se.Defaults()
    se.SetSignal(np.zeros(sr, dtype=np.float32), sr)
    se.Params.WinMs = win_ms
    se.Mel.MFCC = mfcc
    se.Mel.Deltas = deltas
    se.Mel.FBank.NFilters = n_filters
    se.Mel.FBank.HiHz = hi_hz
    if gabor:
        synth.configure_processspeech_gabor(se, out4d=out4d, by_time=by_time)
    se.Init()
    se.DFT.PrevSmooth = prev
    se.DFT.CurSmooth = 1.0 - prev
    p = c_oracle.default_params(sample_rate=sr, win_ms=win_ms, n_filters=n_filters, hi_hz=hi_hz, mfcc=int(mfcc),
                                deltas=int(mfcc and deltas), prev_smooth=prev, cur_smooth=1.0 - prev)
    specs = []
    if gabor:
        c_oracle.with_processspeech_gabor(p, out4d=out4d, by_time=by_time)
        specs = c_oracle.processspeech_specs()
    return se, c_oracle.Env(p, specs)


def signal(sr, seconds, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(int(sr * seconds)) / sr
    x = 0.25 * np.sin(2 * np.pi * 440.0 * t + 1.0) + 0.2 * np.sin(2 * np.pi * 0.11 * sr * t) + 0.15 * np.sin(2 * np.pi * 0.31 * sr * t + 2.0)
    x += rng.normal(0.0, 0.05, t.size)
    return np.clip(x, -1.0, 1.0).astype(np.float32)


@pytest.mark.parametrize("sr,win,prev", [(44100, 1103, 0.0), (22050, 551, 0.3), (48000, 1200, 0.0), (11025, 276, 0.0)])
def test_other_sample_rates_all_outputs(sr, win, prev):
    """25 ms windows at the usual audio rates: prime (1103), 19 x 29 (551), highly composite (1200)."""
    se, orc = envs(sr, prev=prev)
    assert se.Params.WinSamples == win
    sig = signal(sr, 1.3, seed=sr)
    got = se.ProcessBatch(sig, [0], [sig.size], want=ALL)
    ref = orc.process(sig.astype(np.float64), want_power=True)
    assert got["power"].shape[1] == win // 2 + 1
    print(sr, compare(got, ref, ALL))
    assert np.array_equal(got["mfcc"][:, 0, :], got["energy"])


def test_8khz_narrow_bank_and_16khz_short_window():
    se, orc = envs(8000, n_filters=20, hi_hz=4000.0, gabor=False)
    assert se.Params.WinSamples == 200
    sig = signal(8000, 2.0, seed=8)
    names = ["mel", "energy", "mfcc", "deltas", "delta_deltas", "power", "logpower"]
    compare(se.ProcessBatch(sig, [0], [sig.size], want=names), orc.process(sig.astype(np.float64), want_power=True), names)
    # 20 ms window at 16 kHz: 320 samples, same rate as the fused path but the general route
    se, orc = envs(16000, win_ms=20.0, prev=0.3)
    assert se.Params.WinSamples == 320
    sig = synth.config1_signal()
    compare(se.ProcessBatch(sig, [0], [sig.size], want=ALL), orc.process(sig.astype(np.float64), want_power=True), ALL)


def test_ragged_batch_tail_offsets_and_int16():
    sr = 22050
    se, orc = envs(sr, deltas=False)
    rng = np.random.default_rng(3)
    lens = [int(sr * s) + int(rng.integers(0, 97)) for s in (0.35, 1.0, 0.1, 0.62, 0.02)]   # the last two are shorter than a segment / a window
    off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    wave = np.concatenate([signal(sr, n / sr + 0.01, seed=i)[:n] for i, n in enumerate(lens)])
    names = ["mel", "energy", "mfcc", "gabor"]
    for add_ms in (0, 7):
        got = se.ProcessBatch(wave, off, lens, want=names, add=add_ms)
        parts = [orc.process(wave[o:o + n].astype(np.float64), add_ms=add_ms) for o, n in zip(off, lens)]
        ref = {k: np.concatenate([p[k] for p in parts]) for k in names}
        assert got["mel"].shape[0] == ref["mel"].shape[0]
        compare(got, ref, names)
    # 16-bit PCM in, normalised on the GPU like Wave.GetFloatAtIdx (sound/sound.go:130-141)
    pcm = np.round(wave * 32767.0).astype(np.int16)
    got = se.pipeline().process_host(pcm, off, np.asarray(lens, np.int32), want=("mel",))
    parts = [orc.process(pcm[o:o + n].astype(np.float64) / 32767.0) for o, n in zip(off, lens)]
    assert_close(got["mel"], np.concatenate([p["mel"] for p in parts]).reshape(got["mel"].shape), RTOL_LOG, "mel(int16)")


def test_known_answers_prime_length():
    """Impulse -> flat power 1; DC -> bin 0 = N^2; silence -> ln 1 = 0 and mel = LogMin (mel.go:135)."""
    sr, n = 44100, 1103
    se, _ = envs(sr, mfcc=False, gabor=False)
    seg = se.Params.SegmentSteps
    x = np.zeros(sr, dtype=np.float32)
    got = se.ProcessBatch(x, [0], [x.size], want=["mel", "power", "logpower"])
    assert np.all(got["power"] == 0.0) and np.all(got["logpower"] == 0.0)
    assert np.all(got["mel"][:-1] == -10.0)            # exact-zero sums -> LogMin; the last segment's tail steps stay 0
    assert set(np.unique(got["mel"][-1])) <= {-10.0, 0.0}
    x[:] = 0.5
    got = se.ProcessBatch(x, [0], [x.size], want=["power"])
    # border steps of segment 0 are front-padded with zeros; step 2 onward see a full DC window
    dc = got["power"][1, :, 5]
    assert abs(dc[0] - (0.5 * n) ** 2) <= 1e-5 * (0.5 * n) ** 2 and np.all(dc[1:] <= 1e-6 * dc[0])
    assert seg == got["power"].shape[2]


@pytest.mark.parametrize("sr,win", [(48000, 1200), (44100, 1103), (8000, 200)])
def test_tensor_core_and_fp32_frame_power_agree(sr, win):
    """The general route's frame power runs on the tensor cores (tcgen05, two-slice FP16 split, aud_dft_tc.cuh); the FP32
    FMA kernel it replaced stays selectable (option dft_tc = 0).  Both against the oracle, and against each other:
    601 bins in five tiles of 128 (48 kHz), 552 in five of 112 (44.1 kHz, prime window), 101 in one (8 kHz)."""
    se, orc = envs(sr, hi_hz=min(8000.0, sr / 2.0) if sr > 8000 else 4000.0, n_filters=32 if sr > 8000 else 20, gabor=sr > 8000)
    assert se.Params.WinSamples == win
    sig = signal(sr, 1.7, seed=sr + 1)
    names = ["mel", "energy", "mfcc", "power", "logpower"] + (["gabor"] if sr > 8000 else [])
    ref = orc.process(sig.astype(np.float64), want_power=True)
    got = {}
    for mode in (1, 0):
        se.pipeline().set_option("dft_tc", mode)
        got[mode] = se.ProcessBatch(sig, [0], [sig.size], want=names)
        compare(got[mode], ref, names)
    se.pipeline().set_option("dft_tc", 1)
    peak = np.abs(ref["power"]).max()
    assert np.abs(got[1]["power"] - got[0]["power"]).max() <= 1e-5 * peak
    assert_close(got[1]["mel"], got[0]["mel"], RTOL_LOG, "mel tc vs fp32")


@pytest.mark.parametrize("mode", [1, 0])
def test_nan_sample_spoils_only_its_frames_general_route(mode):
    """dft/dft.go:53-59: every frame is transformed on its own, so a NaN sample reaches the frames that contain it
    (and, through the 0 * NaN of dft.go:66-68, the later steps of those segments) and nothing else."""
    sr = 44100
    se, orc = envs(sr, mfcc=False, gabor=False)
    se.pipeline().set_option("dft_tc", mode)
    sig = signal(sr, 1.3, seed=5)
    sig[30011] = np.nan
    names = ["mel", "power"]
    got = se.ProcessBatch(sig, [0], [sig.size], want=names)
    ref = orc.process(sig.astype(np.float64), want_power=True)
    assert np.isnan(ref["mel"]).any() and np.isfinite(ref["mel"]).any()
    assert_close(got["mel"], np.asarray(ref["mel"]).reshape(got["mel"].shape), RTOL_LOG, "mel")
    rp = np.asarray(ref["power"]).reshape(got["power"].shape)
    assert np.array_equal(np.isnan(got["power"]), np.isnan(rp))
    se.pipeline().set_option("dft_tc", 1)


def test_many_work_items_per_cta_general_route():
    """A batch large enough that every CTA of the persistent tensor-core kernel walks through several work items
    (stage / accumulator phases carried from item to item): 640 ragged utterances at 22.05 kHz, ~1400 items on 148 CTAs.
    Tensor-core route against the FP32 route on everything, against the oracle on sampled utterances."""
    sr = 22050
    se, orc = envs(sr, mfcc=False, gabor=False)
    rng = np.random.default_rng(11)
    lens = (sr * rng.uniform(0.6, 1.2, size=640)).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    base = signal(sr, 1.3, seed=2)
    wave = np.concatenate([np.roll(base, int(rng.integers(0, base.size)))[:n] * rng.uniform(0.05, 1.0) for n in lens]).astype(np.float32)
    pipe = se.pipeline()
    got = {}
    for mode in (1, 0):
        pipe.set_option("dft_tc", mode)
        got[mode] = pipe.process_host(wave, off, lens.astype(np.int32), want=("mel",))["mel"]
    pipe.set_option("dft_tc", 1)
    assert_close(got[1], got[0], RTOL_LOG, "mel tc vs fp32, 640 utterances")
    seg_base = pipe.seg_base(lens.astype(np.int32))
    for u in (0, 147, 148, 333, 639):
        ref = orc.process(wave[off[u]:off[u] + lens[u]].astype(np.float64))["mel"]
        g = got[1][seg_base[u]:seg_base[u + 1]]
        assert_close(g, np.asarray(ref).reshape(g.shape), RTOL_LOG, f"mel utt {u}")


def test_longest_window_general_route():
    """WinSamples = 4096 (256 ms at 16 kHz), the largest the C-ABI accepts: 2049 bins in 17 tiles, 33 k-blocks.
    The band is kept narrow so that the reference's mel table does not overflow (mel.go:96-115)."""
    se, orc = envs(16000, win_ms=256.0, hi_hz=300.0, mfcc=False, gabor=False)
    assert se.Params.WinSamples == 4096
    sig = signal(16000, 2.0, seed=4)
    names = ["mel", "power"]
    ref = orc.process(sig.astype(np.float64), want_power=True)
    for mode in (1, 0):
        se.pipeline().set_option("dft_tc", mode)
        compare(se.ProcessBatch(sig, [0], [sig.size], want=names), ref, names)
    se.pipeline().set_option("dft_tc", 1)


def test_operand_scaling_across_amplitudes_general_route():
    """The tensor-core kernel scales every frame's samples into FP16 range by a power of two (aud_dft_tc.cuh); utterances
    that differ by ten orders of magnitude in one batch must each keep their own accuracy, and a level step of 160 dB
    inside one utterance must not cost the quiet part its parity (every frame stands alone, dft/dft.go:42-59)."""
    sr = 44100
    se, orc = envs(sr, mfcc=False, gabor=False)
    base = signal(sr, 0.8, seed=9)
    step = base.copy()
    step[: step.size // 2] *= 1e-8          # -160 dB, then full scale
    utts = [base * np.float32(1e-6), base * np.float32(3e4), base, step]
    lens = np.array([u.size for u in utts], dtype=np.int32)
    off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    wave = np.concatenate(utts).astype(np.float32)
    pipe = se.pipeline()
    got = pipe.process_host(wave, off, lens, want=("mel",))["mel"]
    seg_base = pipe.seg_base(lens)
    for u, x in enumerate(utts):
        ref = orc.process(x.astype(np.float64))["mel"]
        g = got[seg_base[u]:seg_base[u + 1]]
        assert_close(g, np.asarray(ref).reshape(g.shape), RTOL_LOG, f"mel utt {u}")


def test_long_utterance_is_cut_into_jobs_general_route():
    """One utterance of 120 segments becomes four jobs of at most 32 segments on the general route (so that the per-job
    kernels see several CTAs); the frames two jobs share are computed twice, the segments must not notice."""
    sr = 22050
    se, orc = envs(sr, mfcc=False, gabor=False)
    sig = signal(sr, 12.0, seed=21)
    got = se.ProcessBatch(sig, [0], [sig.size], want=["mel"])["mel"]
    ref = np.asarray(orc.process(sig.astype(np.float64))["mel"]).reshape(got.shape)
    assert got.shape[0] >= 118
    assert_close(got, ref, RTOL_LOG, "mel, 12 s utterance")
