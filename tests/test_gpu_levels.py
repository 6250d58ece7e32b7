"""GPU parity on signals with abrupt level changes.  The fused kernel rides two real frames on one complex
FFT; in the reference every frame stands alone (dft/dft.go:42-59), so a quiet frame next to a loud one, a
frame of digital silence next to speech, or a clean frame next to one that holds a NaN / Inf sample must
come out as if it had been transformed alone.  Tolerances are the BASELINE ones (tests/util.py)."""
import numpy as np
import pytest

from auditory_b200 import synth
from test_gpu_parity import compare, make_env, oracle_env
from util import RTOL_LOG, assert_close

pytestmark = pytest.mark.gpu
SR = synth.SR


def run_both(sig, names=("mel", "energy", "mfcc", "gabor"), **kw):
    se = make_env(**kw)
    got = se.ProcessBatch(sig, [0], [sig.size], want=list(names))
    ref = oracle_env(**kw).process(sig.astype(np.float64))
    return got, ref


@pytest.mark.parametrize("level_db", [-40, -60, -80, -100])
def test_quiet_noise_then_full_scale_burst(level_db):
    """Noise at level_db dBFS, a full-scale broadband burst that starts in the middle of a frame (and not on
    a hop boundary), then the quiet noise again: every quiet frame that is paired with a burst frame."""
    rng = np.random.default_rng(100 - level_db)
    n = 2 * SR
    sig = rng.normal(0.0, 10.0 ** (level_db / 20.0), n)
    for start, length in ((7 * 1600 + 333, 3000), (12 * 1600 + 160 * 3 + 77, 401), (17 * 1600 - 1, 4000)):
        sig[start:start + length] = rng.uniform(-1.0, 1.0, length)
    sig = sig.astype(np.float32)
    got, ref = run_both(sig, prev=0.0)
    print(level_db, compare(got, ref, ["mel", "energy", "mfcc", "gabor"]))


def test_quiet_burst_with_smoothing_and_int16():
    rng = np.random.default_rng(5)
    pcm = rng.normal(0.0, 3.0, 2 * SR)                       # about -80 dBFS in 16-bit PCM
    pcm[9000:9000 + 2500] = rng.uniform(-30000, 30000, 2500)
    pcm[20000:20000 + 170] = rng.uniform(-30000, 30000, 170)
    pcm = np.round(pcm).astype(np.int16)
    kw = dict(mfcc=True, deltas=False, gabor=True, prev=0.3)
    se = make_env(**kw)
    got = se.pipeline().process_host(pcm, [0], [pcm.size], want=["mel", "mfcc", "energy", "gabor"])
    ref = oracle_env(**kw).process(pcm.astype(np.float64) / float(0x7FFF))
    print(compare(got, ref, ["mel", "mfcc", "energy", "gabor"]))


def test_leading_silence_then_speech():
    """SndEnv.AdjustForSilence prepends exact zeros (sndenv.go:274-294): frames of digital silence, frames that
    straddle the onset, then a harmonic signal with a syllable-like envelope."""
    n0 = 5 * 1600 + 437
    t = np.arange(int(1.5 * SR)) / SR
    voiced = sum(a * np.sin(2 * np.pi * f * t) for f, a in ((140, 0.3), (280, 0.2), (420, 0.15), (2300, 0.05)))
    env = 0.5 * (1.0 - np.cos(2 * np.pi * 4.0 * t)) ** 2
    sig = np.concatenate([np.zeros(n0), voiced * env * 0.25]).astype(np.float32)
    got, ref = run_both(sig, prev=0.0)
    print(compare(got, ref, ["mel", "energy", "mfcc", "gabor"]))
    # frames that lie wholly inside the zeros: exactly LogMin (mel.go:135), in both
    assert np.all(got["mel"][1, :, 2:10] == -10.0) and np.all(ref["mel"][1, :, 2:10] == -10.0)


def test_impulse_train_over_a_noise_floor():
    rng = np.random.default_rng(9)
    sig = rng.normal(0.0, 10.0 ** (-70 / 20.0), 2 * SR)
    sig[500::777] = 0.9
    sig = sig.astype(np.float32)
    got, ref = run_both(sig, prev=0.0)
    print(compare(got, ref, ["mel", "energy", "mfcc", "gabor"]))


def frames_spoiled(index, nseg, S=14, border=2, step=160, stride=1600, win=400):
    """(segment, step) pairs the reference leaves non-finite for a NaN / Inf sample at `index`: the steps whose
    window holds the sample, and every later step of the same segment -- dft.Power smooths with
    `PrevSmooth*Power[k] + CurSmooth*p` from step 1 on (dft/dft.go:66-68), and 0 * NaN is NaN, so even with
    PrevSmooth = 0 the spoiled bins are carried to the end of the segment (step 0 of the next one starts clean)."""
    hit = np.zeros((nseg, S), dtype=bool)
    for seg in range(nseg):
        for s in range(S):
            start = seg * stride + (s - border) * step
            if start <= index < start + win:
                hit[seg, s:] = True
    return hit


@pytest.mark.parametrize("prev", [0.0, 0.4])
@pytest.mark.parametrize("bad", [np.nan, np.inf])
def test_one_non_finite_sample_stays_in_its_own_frames(bad, prev):
    """dft.FftReal copies one window per transform (dft/dft.go:53-59): a NaN / Inf sample spoils the frames
    whose window holds it (and, through the smoothing line, the rest of their segment) and no others."""
    sig = synth.config1_signal().copy()
    idx = 9 * 1600 + 523
    sig[idx] = bad
    se = make_env(mfcc=False, gabor=False, prev=prev)
    got = se.ProcessBatch(sig, [0], [sig.size], want=["mel", "energy"])
    ref = oracle_env(mfcc=False, gabor=False, prev=prev).process(sig.astype(np.float64))
    hit = frames_spoiled(idx, got["mel"].shape[0])
    g = np.moveaxis(got["mel"], 1, 2)          # [seg][step][filter]
    r = np.moveaxis(ref["mel"], 1, 2)
    assert not np.isfinite(r[hit]).any() and np.isfinite(r[~hit]).all(), "oracle: spoiled steps are what the rule says"
    assert not np.isfinite(g[hit]).any(), "gpu: spoiled steps must be non-finite in every filter"
    assert_close(g[~hit], r[~hit], RTOL_LOG, f"frames without the {bad} sample")
    # Energy sums ln(power + 1) of bin s over the steps (sndenv.go:360-366): spoiled wherever the segment is
    assert np.array_equal(np.isfinite(got["energy"]), np.isfinite(ref["energy"]))
    ok = np.isfinite(ref["energy"])
    assert_close(got["energy"][ok], ref["energy"][ok], RTOL_LOG, "energy of clean segments")


def test_nan_sample_through_gabor():
    """NaN mel values enter agabor.Convolve as 0.5 (agabor/gabor.go:283-285)."""
    sig = synth.config1_signal().copy()
    sig[4 * 1600 + 160 * 4 + 11] = np.nan
    got, ref = run_both(sig, names=("mel", "gabor"), mfcc=False, prev=0.0)
    nan_ref = np.isnan(ref["mel"])
    assert nan_ref.any() and np.array_equal(np.isnan(got["mel"]), nan_ref)
    assert_close(got["mel"][~nan_ref], ref["mel"][~nan_ref], RTOL_LOG, "mel outside the NaN frames")
    compare(got, ref, ["gabor"])


@pytest.mark.parametrize("n_filters", [36, 40, 64])
def test_more_than_32_mel_filters(n_filters):
    """The frame ring's row pitch follows NFilters (40 filters at 16 kHz is a common setting)."""
    import auditory_b200 as ab
    from oracle import c_oracle
    sig = synth.config1_signal()
    se = ab.SndEnv(device=0)
    se.Defaults()
    se.SetSignal(sig, SR)
    se.Mel.FBank.NFilters = n_filters
    se.Mel.Deltas = False
    synth.configure_processspeech_gabor(se)
    se.Init()
    se.DFT.PrevSmooth, se.DFT.CurSmooth = 0.25, 0.75
    got = se.ProcessBatch(sig, [0], [sig.size], want=["mel", "mfcc", "energy", "gabor"])
    p = c_oracle.default_params(n_filters=n_filters, mfcc=1, deltas=0, prev_smooth=0.25, cur_smooth=0.75)
    c_oracle.with_processspeech_gabor(p)
    ref = c_oracle.Env(p, c_oracle.processspeech_specs()).process(sig.astype(np.float64))
    assert got["mel"].shape == (20, n_filters, 14)
    print(compare(got, ref, ["mel", "mfcc", "energy", "gabor"]))
    se.DFT.PrevSmooth, se.DFT.CurSmooth = 0.0, 1.0
    got = se.ProcessBatch(sig, [0], [sig.size], want=["mel"])
    p.prev_smooth, p.cur_smooth = 0.0, 1.0
    ref = c_oracle.Env(p, c_oracle.processspeech_specs()).process(sig.astype(np.float64))
    print(compare(got, ref, ["mel"]))
