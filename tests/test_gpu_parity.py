"""GPU parity: the fused sm_100a kernel (called through the C-ABI) against the float64 oracle on
the same inputs.  Tolerances are BASELINE.json's: 1e-4 on log-power / log-mel / energy / MFCC and
1e-3 on gabor outputs, relative to max(1, |ref|) (see tests/util.py)."""
import os

import numpy as np
import pytest

import auditory_b200 as ab
from auditory_b200 import synth
from oracle import c_oracle
from util import RTOL_GABOR, RTOL_LOG, assert_close

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
ALL = ["mel", "energy", "mfcc", "deltas", "delta_deltas", "gabor", "power", "logpower"]


def make_env(mfcc=True, deltas=True, gabor=True, prev=0.0, cur=None, out4d=True, by_time=False, **sound):
    se = ab.SndEnv(device=0)
    se.Defaults()
    se.SetSignal(np.zeros(48000, dtype=np.float32), synth.SR)
    for k, v in sound.items():
        setattr(se.Params, k, v)
    se.Mel.MFCC = mfcc
    se.Mel.Deltas = deltas
    se.Kwta.On = False                 # Kwta / NeighInhib off (SndEnv.Defaults turns kwta on)
    if gabor:
        synth.configure_processspeech_gabor(se, out4d=out4d, by_time=by_time)
    se.Init()
    se.DFT.PrevSmooth = prev
    se.DFT.CurSmooth = (1.0 - prev) if cur is None else cur
    return se


def oracle_env(mfcc=True, deltas=True, gabor=True, prev=0.0, cur=None, out4d=True, by_time=False, **sound):
    names = {"WinMs": "win_ms", "StepMs": "step_ms", "SegmentMs": "segment_ms", "StrideMs": "stride_ms",
             "BorderSteps": "border_steps"}
    p = c_oracle.default_params(mfcc=int(mfcc), deltas=int(mfcc and deltas), prev_smooth=prev,
                                cur_smooth=(1.0 - prev) if cur is None else cur,
                                **{names[k]: v for k, v in sound.items()})
    specs = []
    if gabor:
        c_oracle.with_processspeech_gabor(p, out4d=out4d, by_time=by_time)
        specs = c_oracle.processspeech_specs()
    return c_oracle.Env(p, specs)


def oracle_batch(env, wave, off, ln, add_ms=0, want_power=False):
    parts = [env.process(wave[o:o + n].astype(np.float64), add_ms=add_ms, want_power=want_power)
             for o, n in zip(off, ln)]
    return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}


def compare(got, ref, names):
    worst = {}
    for n in names:
        g = got[n].astype(np.float64)
        r = ref[n].reshape(g.shape)
        if n == "power":
            # raw power spans many decades inside a frame: float32 FFT error is relative to the frame's peak
            scale = np.maximum(np.abs(r).max(axis=1, keepdims=True), 1.0)
            assert np.all(np.abs(g - r) <= 2e-5 * scale), "power"
        elif n in ("deltas", "delta_deltas"):
            # the reference's accumulators run across coefficients: compare against the tensor's scale
            assert np.abs(g - r).max() <= RTOL_LOG * max(1.0, np.abs(r).max()), n
        else:
            worst[n] = assert_close(g, r, RTOL_GABOR if n == "gabor" else RTOL_LOG, n)
    return worst


# ----------------------------------------------------------------------------- config 1
@pytest.mark.parametrize("prev", [0.0, 0.3])
def test_config1_all_outputs(prev):
    sig = synth.config1_signal()
    se = make_env(prev=prev)
    got = se.ProcessBatch(sig, [0], [sig.size], want=ALL)
    ref = oracle_env(prev=prev).process(sig.astype(np.float64), want_power=True)
    assert got["mel"].shape == (20, 32, 14) and got["gabor"].shape == (20, 256)
    print(compare(got, ref, ALL))
    # tail: the last segment's steps 10..13 run past the signal -> exact zeros (sndenv.go:355-358)
    assert np.all(got["mel"][19, :, 10:] == 0.0) and np.all(got["power"][19, :, 10:] == 0.0)
    # ... except MFCC row 0, which is Energy for every step (sndenv.go:368-372)
    assert np.array_equal(got["mfcc"][:, 0, :], got["energy"])
    # committed golden fixture (frozen oracle output)
    g = np.load(os.path.join(GOLDEN, "cfg1.npz" if prev == 0.0 else "cfg1_smooth.npz"))
    compare(got, {k: g[k] for k in ("mel", "energy", "mfcc", "gabor")}, ["mel", "energy", "mfcc", "gabor"])
    assert_close(got["logpower"][19], g["logpower_seg19"], RTOL_LOG, "logpower[19]")


def test_sndenv_call_shape_matches_reference_usage():
    """SndEnv.ProcessSegment(segment, add) + ApplyGabor(), one segment at a time (sndenv.go:342, 481)."""
    sig = synth.config1_signal()
    se = make_env(prev=0.0)
    se.SetSignal(sig, synth.SR)
    se.Init()
    ref = oracle_env().process(sig.astype(np.float64), want_power=True)
    assert se.SegCnt == 20
    for seg in (0, 7, 19):
        se.ProcessSegment(seg, 0)
        out = se.ApplyGabor()
        assert out.shape == (8, 2, 2, 8)
        assert_close(se.MelFBankSegment, ref["mel"][seg], RTOL_LOG, "MelFBankSegment")
        assert_close(se.MFCCSegment, ref["mfcc"][seg], RTOL_LOG, "MFCCSegment")
        assert_close(se.LogPowerSegment, ref["logpower"][seg], RTOL_LOG, "LogPowerSegment")
        assert_close(out.reshape(-1), ref["gabor"][seg], RTOL_GABOR, "GborOutput")


# ----------------------------------------------------------------------------- configs 2 / 3 (small batch) + goldens
def test_config2_config3_batch_and_goldens():
    wave, off, ln = synth.batch(4)
    for tag, kw, names in (("cfg2_mel", dict(mfcc=False, gabor=False), ["mel"]),
                           ("cfg3_mfcc_smooth", dict(mfcc=True, deltas=False, gabor=False, prev=0.3, cur=0.7), ["mel", "mfcc", "energy"])):
        se = make_env(**kw)
        got = se.ProcessBatch(wave, off, ln, want=names)
        ref = oracle_batch(oracle_env(**kw), wave, off, ln)
        assert got["mel"].shape[0] == 120
        compare(got, ref, names)
        g = np.load(os.path.join(GOLDEN, tag + ".npz"))
        for u in range(4):
            assert_close(got["mel"][30 * u:30 * u + 30], g[f"mel_{u}"], RTOL_LOG, f"{tag} mel_{u}")
            if "mfcc" in names:
                assert_close(got["mfcc"][30 * u:30 * u + 30], g[f"mfcc_{u}"], RTOL_LOG, f"{tag} mfcc_{u}")


# ----------------------------------------------------------------------------- ragged, empty, misaligned
@pytest.mark.parametrize("pack", ["tight", "aligned"])
def test_ragged_batch(pack):
    lens = np.array([48000, 16001, 1700, 999, 0, 33333, 2000, 48000, 1601], dtype=np.int32)
    off, pos = [], 5 if pack == "tight" else 0
    for n in lens:
        off.append(pos)
        pos += int(n) if pack == "tight" else (int(n) + 3) // 4 * 4 + 8
    off = np.array(off, dtype=np.int64)
    rng = np.random.default_rng(7)
    wave = rng.uniform(-1, 1, pos + 4).astype(np.float32)
    se = make_env(mfcc=True, deltas=True, gabor=True)
    got = se.ProcessBatch(wave, off, lens, want=ALL)
    ref = oracle_batch(oracle_env(), wave, off, lens, want_power=True)
    # SegCnt quirks: 999 samples -> 1 segment, 0 samples -> 0 (Go integer division, sndenv.go:263-265)
    assert got["mel"].shape[0] == 30 + 10 + 1 + 1 + 0 + 20 + 1 + 30 + 1
    compare(got, ref, ALL)


@pytest.mark.parametrize("add_ms", [20, 7, -5])
def test_add_offset(add_ms):
    sig = synth.config1_signal()
    se = make_env(mfcc=False, gabor=False)
    got = se.ProcessBatch(sig, [0], [sig.size], add=add_ms, want=["mel", "energy"])
    ref = oracle_env(mfcc=False, gabor=False).process(sig.astype(np.float64), add_ms=add_ms)
    compare(got, ref, ["mel", "energy"])


def test_odd_add_samples_takes_the_unaligned_path():
    sig = synth.config1_signal()
    se = make_env(mfcc=False, gabor=False)
    pipe = se.pipeline()
    got = pipe.process_host(sig, [0], [sig.size], want=["mel"], add_samples=113)
    # the oracle takes milliseconds; shift the signal instead: add moves every window 113 samples later
    shifted = np.concatenate([sig[113:], np.zeros(113, dtype=np.float32)])
    ref = oracle_env(mfcc=False, gabor=False).process(shifted.astype(np.float64))
    # compare the segments whose windows do not touch the signal's ends (front pad / tail differ by construction)
    assert_close(got["mel"][1:18], ref["mel"][1:18], RTOL_LOG, "mel (add=113 samples)")


# ----------------------------------------------------------------------------- other geometries
@pytest.mark.parametrize("sound", [
    dict(StrideMs=75.0),                 # stride not a multiple of the hop: every segment has its own frames
    dict(StrideMs=30.0),                 # heavy overlap: a frame is shared by up to 5 segments
    dict(StrideMs=300.0),                # gaps between segments
    dict(StepMs=12.5, SegmentMs=100.0),  # hop 200, S = 12
    dict(StepMs=5.0, BorderSteps=3),     # hop 80, S = 26: more than 20 low bins feed Energy
    dict(BorderSteps=0),
])
@pytest.mark.parametrize("prev", [0.0, 0.25])
def test_geometries(sound, prev):
    sig = synth.batch_utterance(3, seconds=1.7)
    kw = dict(mfcc=True, deltas=True, gabor=False, prev=prev, **sound)
    se = make_env(**kw)
    got = se.ProcessBatch(sig, [0], [sig.size], want=["mel", "energy", "mfcc", "deltas", "delta_deltas", "logpower"])
    ref = oracle_env(**kw).process(sig.astype(np.float64), want_power=True)
    assert got["mel"].shape == ref["mel"].shape
    compare(got, ref, ["mel", "energy", "mfcc", "deltas", "delta_deltas", "logpower"])


@pytest.mark.parametrize("prev", [0.0, 0.3])
def test_long_segment_parallel_scan(prev):
    """One long segment (S = 194 steps) per stride, gaborview-style: with smoothing this runs the
    Kogge-Stone scan over the steps, chained through carries (config 5's long-segment variant)."""
    sig = synth.long_signal(seconds=6.0)
    kw = dict(mfcc=False, gabor=False, prev=prev, SegmentMs=1900.0, StrideMs=1900.0)
    se = make_env(**kw)
    assert se.Params.SegmentSteps == 194
    got = se.ProcessBatch(sig, [0], [sig.size], want=["mel"])
    ref = oracle_env(**kw).process(sig.astype(np.float64))
    assert got["mel"].shape == (3, 32, 194)
    compare(got, ref, ["mel"])


def test_medium_segment_with_energy_and_mfcc():
    """S = 44 steps: Energy needs 44 low bins per frame (more than the 20 a lane group holds)."""
    sig = synth.long_signal(seconds=4.0)
    kw = dict(mfcc=True, deltas=True, gabor=False, prev=0.2, SegmentMs=400.0, StrideMs=200.0)
    se = make_env(**kw)
    assert se.Params.SegmentSteps == 44
    got = se.ProcessBatch(sig, [0], [sig.size], want=["mel", "energy", "mfcc", "deltas"])
    ref = oracle_env(**kw).process(sig.astype(np.float64))
    compare(got, ref, ["mel", "energy", "mfcc", "deltas"])


@pytest.mark.parametrize("by_time", [False, True])
def test_gabor_2d_layouts(by_time):
    sig = synth.config1_signal()
    kw = dict(mfcc=False, gabor=True, out4d=False, by_time=by_time)
    se = make_env(**kw)
    got = se.ProcessBatch(sig, [0], [sig.size], want=["mel", "gabor"])
    ref = oracle_env(**kw).process(sig.astype(np.float64))
    compare(got, ref, ["mel", "gabor"])
    assert np.count_nonzero(got["gabor"]) > 0


# ----------------------------------------------------------------------------- analytic inputs
def test_known_answers_on_gpu():
    se = make_env(mfcc=False, gabor=False)
    n = 32000
    silence = np.zeros(n, dtype=np.float32)
    impulse = silence.copy()
    impulse[1600 + 5] = 1.0
    t = np.arange(n) / 16000.0
    tone = (0.25 * np.sin(2 * np.pi * 1000.0 * t)).astype(np.float32)     # bin 25 exactly
    wave = np.concatenate([silence, impulse, tone])
    got = se.ProcessBatch(wave, [0, n, 2 * n], [n, n, n], want=["mel", "power", "logpower"])
    sil, imp, ton = got["mel"][:20], got["power"][20:40], got["power"][40:60]
    assert np.all(sil[:19] == -10.0)                                       # exact zero sum -> LogMin (mel.go:135)
    assert np.all(got["logpower"][:19] == 0.0)                             # ln(0 + 1)
    assert np.allclose(imp[1, :, 2], 1.0, atol=2e-6)                       # flat spectrum of an impulse
    assert abs(ton[4, 25, 4] - 2500.0) < 2500.0 * 1e-5
    assert np.delete(ton[4, :, 4], 25).max() < 2500.0 * 1e-9               # float32 leakage floor of a 400-point FFT
    ref = oracle_env(mfcc=False, gabor=False).process(tone.astype(np.float64))
    # the centred tone leaves most filters at the rounding floor: compare where the oracle sum is above it
    sums = np.exp(ref["mel"])
    mask = sums > 1e-3
    assert mask.sum() > 400
    assert np.all(np.abs(got["mel"][40:60][mask] - ref["mel"][mask]) <= RTOL_LOG * np.maximum(1, np.abs(ref["mel"][mask])))


def test_renorm_branch_and_raw_cepstrum_c0():
    """mel.FilterDft's Renorm clamp (dead after InitFilters, mel.go:80, but part of the operator) and
    CepstrumDct's own c0 = ln(1 + y0^2) (mel.go:203-204) when not driven by SndEnv."""
    sig = synth.config1_signal()
    se = make_env(mfcc=True, deltas=False, gabor=False)
    ap = se.aud_params()
    ap.renorm, ap.renorm_min, ap.renorm_scale = 1, -6.0, 0.1
    ap.mfcc_c0_energy = 0
    pipe = ab.Pipeline(ap, se.Mel.BinPts, se.MelFilters)
    got = pipe.process_host(sig, [0], [sig.size], want=["mel", "mfcc"])
    p = c_oracle.default_params(mfcc=1, deltas=0, renorm=1)
    ref = c_oracle.Env(p).process(sig.astype(np.float64))
    assert_close(got["mel"], ref["mel"], RTOL_LOG, "renormed mel")
    assert got["mel"].min() >= 0.0 and got["mel"].max() <= 1.0
    import scipy.fft
    y = scipy.fft.dct(ref["mel"], type=1, axis=1)[:, :13, :]
    y[:, 0, :] = np.log1p(y[:, 0, :] ** 2)
    y[19, :, 10:] = 0.0                                                    # tail steps never ran
    assert_close(got["mfcc"], y, RTOL_LOG, "mfcc with CepstrumDct c0")


# ----------------------------------------------------------------------------- tiling / scheduling invariance
def test_results_do_not_depend_on_the_launch_plan():
    wave, off, ln = synth.batch(6, seconds=2.3)
    se = make_env(mfcc=True, deltas=False, gabor=True, prev=0.2)
    pipe = se.pipeline()
    base = pipe.process_host(wave, off, ln, want=["mel", "mfcc", "gabor"])
    for opts in (dict(job_segs=1), dict(job_segs=3), dict(job_segs=7, warps=6), dict(warps=8, ctas=5), dict(epi=6), dict(warps=10, epi=4)):
        for k, v in opts.items():
            pipe.set_option(k, v)
        again = pipe.process_host(wave, off, ln, want=["mel", "mfcc", "gabor"])
        for k in base:
            assert np.array_equal(base[k], again[k]), (opts, k)
        for k in opts:
            pipe.set_option(k, 0)


def test_device_entry_point_matches_host_entry_point():
    import torch
    wave, off, ln = synth.batch(3)
    se = make_env(mfcc=False, gabor=False)
    pipe = se.pipeline()
    host = pipe.process_host(wave, off, ln, want=["mel"])
    d_wave = torch.from_numpy(wave).cuda()
    d_mel = torch.empty(host["mel"].shape, dtype=torch.float32, device="cuda")
    n0 = pipe.launch_count
    pipe.process_device(d_wave, off, ln, {"mel": d_mel})
    torch.cuda.synchronize()
    assert pipe.launch_count == n0 + 1
    assert np.array_equal(d_mel.cpu().numpy(), host["mel"])


# ----------------------------------------------------------------------------- config 5: ten minutes, streamed
def test_config5_ten_minute_waveform():
    sig = synth.long_signal(seconds=600.0)
    se = make_env(mfcc=False, gabor=True)
    got = se.ProcessBatch(sig, [0], [sig.size], want=["mel", "gabor"])
    assert got["mel"].shape == (6000, 32, 14)
    ref = oracle_env(mfcc=False, gabor=True).process(sig.astype(np.float64))
    compare(got, ref, ["mel", "gabor"])
    # segment boundaries: first (front pad) and last (tail) segments, and job / CTA seams everywhere in between
    assert np.all(got["mel"][5999, :, 10:] == 0.0)
    # the same waveform cut into two utterances overlapping by the border reproduces the interior segments bit for bit
    cut = 300 * 16000
    a = se.ProcessBatch(sig[:cut + 1600], [0], [cut + 1600], want=["mel"])["mel"]
    assert np.array_equal(a[1:2990], got["mel"][1:2990])


# ----------------------------------------------------------------------------- full size (configs[1]) properties
def test_full_batch_properties():
    wave, off, ln = synth.fast_batch(1024, seed=4321)
    se = make_env(mfcc=False, gabor=False)
    pipe = se.pipeline()
    full = pipe.process_host(wave, off, ln, want=["mel"])["mel"]
    assert full.shape == (30720, 32, 14) and np.isfinite(full).all()
    # determinism
    again = pipe.process_host(wave, off, ln, want=["mel"])["mel"]
    assert np.array_equal(full, again)
    # every utterance is independent of its batch: spot utterances alone reproduce their slice bit for bit
    for u in (0, 511, 1023):
        alone = pipe.process_host(wave[off[u]:off[u] + ln[u]].copy(), [0], [ln[u]], want=["mel"])["mel"]
        assert np.array_equal(alone, full[30 * u:30 * u + 30])
    # and a handful against the oracle
    env = oracle_env(mfcc=False, gabor=False)
    for u in (5, 700):
        ref = env.process(wave[off[u]:off[u] + ln[u]].astype(np.float64))["mel"]
        assert_close(full[30 * u:30 * u + 30], ref, RTOL_LOG, f"utterance {u}")


# ----------------------------------------------------------------------------- int16 PCM ingest
@pytest.mark.parametrize("pack", ["aligned", "tight"])
def test_int16_pcm_input(pack):
    """16-bit PCM in, normalised by 1/0x7FFF on the GPU exactly like Wave.GetFloatAtIdx (sound/sound.go:130-141)."""
    rng = np.random.default_rng(11)
    lens = np.array([48000, 16001, 999, 24000], dtype=np.int32)
    off, pos = [], 0 if pack == "aligned" else 3
    for n in lens:
        off.append(pos)
        pos += (int(n) + 7) // 8 * 8 if pack == "aligned" else int(n)
    off = np.array(off, dtype=np.int64)
    pcm = (rng.normal(0, 6000, pos + 8)).clip(-32768, 32767).astype(np.int16)
    se = make_env(mfcc=True, deltas=False, gabor=True)
    pipe = se.pipeline()
    got = pipe.process_host(pcm, off, lens, want=["mel", "mfcc", "energy", "gabor"])
    as_float = (pcm.astype(np.float64) / float(0x7FFF))
    ref = oracle_batch(oracle_env(mfcc=True, deltas=False, gabor=True), as_float, off, lens)
    compare(got, ref, ["mel", "mfcc", "energy", "gabor"])
    # and it agrees with the float32 entry point fed the normalised samples
    f32 = pipe.process_host(as_float.astype(np.float32), off, lens, want=["mel"])
    assert np.abs(f32["mel"] - got["mel"]).max() < 2e-4


def test_more_utterances_than_one_launch_holds():
    """BASELINE config 4 hands one GPU thousands of utterances: more jobs than the persistent CTAs hold in
    one launch, so the library cuts the batch into runs.  Same numbers as smaller calls, and the oracle's."""
    n_utt, n = 9000, 4800
    wave, off, ln = synth.fast_batch(n_utt, seed=77, seconds=n / synth.SR)
    se = make_env(mfcc=False, gabor=False)
    pipe = se.pipeline()
    got = pipe.process_host(wave, off, ln, want=("mel",))["mel"]
    assert got.shape == (3 * n_utt, 32, 14)
    a = pipe.process_host(wave, off[:4000], ln[:4000], want=("mel",))["mel"]
    b = pipe.process_host(wave, off[4000:], ln[4000:], want=("mel",))["mel"]
    assert np.array_equal(got, np.concatenate([a, b]))
    orc = oracle_env(mfcc=False, gabor=False)
    for u in (0, 7103, 7104, 8999):
        ref = orc.process(wave[off[u]:off[u] + n].astype(np.float64))["mel"]
        assert_close(got[3 * u:3 * u + 3], ref.reshape(3, 32, 14), RTOL_LOG, f"mel[utt {u}]")
    # the device entry point takes the same route
    import torch
    d_wave = torch.from_numpy(wave).cuda()
    d_out = torch.empty(got.shape, dtype=torch.float32, device="cuda")
    pipe.process_device(d_wave, off, ln, {"mel": d_out})
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), got)


def test_gabor_configured_but_not_requested():
    """A FilterSet is configured but the caller only wants MFCC (or only gabor): the stages nobody asked for
    must not run, and the ones asked for must not change."""
    sig = synth.config1_signal()
    se = make_env(prev=0.3)
    full = se.ProcessBatch(sig, [0], [sig.size], want=["mel", "mfcc", "gabor", "energy"])
    only_mfcc = se.ProcessBatch(sig, [0], [sig.size], want=["mfcc"])
    only_gabor = se.ProcessBatch(sig, [0], [sig.size], want=["gabor"])
    assert np.array_equal(only_mfcc["mfcc"], full["mfcc"]) and np.array_equal(only_gabor["gabor"], full["gabor"])


def test_empty_batches():
    """No utterances, or only utterances too short for a segment: zero segments, no launch, no error."""
    se = make_env(mfcc=False, gabor=False)
    pipe = se.pipeline()
    before = pipe.launch_count
    out = pipe.process_host(np.zeros(0, dtype=np.float32), [], [], want=("mel",))
    assert out["mel"].shape == (0, 32, 14)
    out = pipe.process_host(np.zeros(10, dtype=np.float32), [0, 5], np.array([0, 0], np.int32), want=("mel",))
    assert out["mel"].shape == (0, 32, 14) and pipe.launch_count == before
