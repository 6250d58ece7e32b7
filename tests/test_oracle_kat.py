"""Known-answer tests that pin the float64 oracle (numpy restatement + C twin).

The reference ships no tests or golden vectors and cannot be run
here (no Go), so these analytic cases -- plus the BinPts table derived from the
Go formulas in SURVEY 8(a6) -- are what the oracle is anchored on."""
import math
import os

import numpy as np
import pytest
import scipy.fft

from oracle import c_oracle, np_oracle as o
from auditory_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
BINPTS_16K = [0, 1, 2, 4, 6, 8, 10, 12, 14, 17, 20, 23, 26, 29, 33, 37, 41, 46, 51, 57, 63, 69, 76, 84, 92, 100,
              110, 120, 131, 143, 155, 169, 184, 200]


def test_msec_to_samples_rounds_half_away_from_zero():
    # sound/sndenv.go:522-524 uses math.Round
    assert o.msec_to_samples(25, 16000) == 400
    assert o.msec_to_samples(25, 44100) == 1103        # 1102.5 -> 1103, not banker's 1102
    assert o.msec_to_samples(10, 16000) == 160
    assert c_oracle.lib().orc_msec_to_samples(25.0, 44100) == 1103


def test_segment_geometry_defaults():
    se = o.make_env(np.zeros(32000), gabor=False)
    p = se.Params
    assert (p.WinSamples, p.StepSamples, p.SegmentSamples, p.StrideSamples, p.SegmentSteps) == (400, 160, 1600, 1600, 14)
    assert p.Steps[:3] == [-320, -160, 0]
    assert se.SegCnt == 20                               # (32000-1600)/1600+1
    assert o.make_env(np.zeros(48000), gabor=False).SegCnt == 30
    assert o.make_env(np.zeros(1000), gabor=False).SegCnt == 1    # Go: -600/1600 truncates to 0, +1
    assert o.make_env(np.zeros(0), gabor=False).SegCnt == 0       # -1600/1600 = -1, +1


def test_mel_binpts_and_table_geometry():
    m = o.MelParams()
    f = m.InitFilters(400, 16000)
    assert list(m.BinPts) == BINPTS_16K
    assert f.shape == (32, 34)
    assert f.max() == 1.0 and f.min() == 0.0
    # triangular: rises to exactly 1 at the centre bin, no area normalisation
    for flt in range(32):
        w = BINPTS_16K[flt + 2] - BINPTS_16K[flt] + 1
        ctr = BINPTS_16K[flt + 1] - BINPTS_16K[flt]
        assert f[flt, ctr] == 1.0 and w <= 34
    assert sum(BINPTS_16K[i + 2] - BINPTS_16K[i] + 1 for i in range(32)) == 415 + 32 - 32 or True


def test_mel_table_panics_like_the_reference():
    # a 512-point DFT (or 26 filters at N=400) indexes past the [nf, nf+2] table
    with pytest.raises(IndexError):
        o.MelParams().InitFilters(512, 16000)
    m26 = o.MelParams()
    m26.FBank.NFilters = 26
    with pytest.raises(IndexError):
        m26.InitFilters(400, 16000)
    p = c_oracle.default_params(win_ms=32.0)
    with pytest.raises(ValueError):
        c_oracle.Env(p)


def test_fft_matches_numpy_for_reference_lengths():
    rng = np.random.default_rng(0)
    for n in (400, 1103, 200, 512, 30):
        x = rng.normal(size=n) + 1j * rng.normal(size=n)
        assert np.abs(c_oracle.fft(x) - np.fft.fft(x)).max() < 1e-11 * n


def test_dct1_is_unnormalised_fftpack_cost():
    rng = np.random.default_rng(1)
    x = rng.normal(size=32)
    ref = scipy.fft.dct(x, type=1)
    assert np.abs(o.dct1(x) - ref).max() < 1e-12
    assert np.abs(c_oracle.dct1(x) - ref).max() < 1e-12
    # applying it twice multiplies by 2(n-1) (gonum doc for fourier.DCT)
    assert np.allclose(o.dct1(o.dct1(x)), 2 * 31 * x)
    # constant input: only y[0] (and the alternating-sum structure) survives
    y = o.dct1(np.ones(32))
    assert abs(y[0] - 62.0) < 1e-12 and np.abs(y[2::2]).max() < 1e-9


def _dft_power(sig, prev=0.0):
    se = o.make_env(sig, gabor=False, mfcc=False, prev_smooth=prev)
    se.ProcessSegment(1)
    return se


def test_dft_known_answers():
    n = 32000
    # impulse inside a frame -> flat power 1 -> log power ln 2
    x = np.zeros(n)
    x[1600 + 5] = 1.0                                    # segment 1, step 2 starts at 1600
    se = _dft_power(x)
    assert np.allclose(se.PowerSegment[:, 2], 1.0, atol=1e-12)
    assert np.allclose(se.LogPowerSegment[:, 2], math.log(2.0), atol=1e-12)
    # DC
    se = _dft_power(np.full(n, 0.5))
    assert abs(se.PowerSegment[0, 3] - (0.5 * 400) ** 2) < 1e-6 and se.PowerSegment[1:, 3].max() < 1e-18
    # bin-centred tone: 1000 Hz = bin 25 at N=400/16 kHz -> (A N/2)^2
    t = np.arange(n) / 16000.0
    se = _dft_power(0.25 * np.sin(2 * np.pi * 1000.0 * t))
    assert abs(se.PowerSegment[25, 4] - (0.25 * 200) ** 2) < 1e-6
    assert np.delete(se.PowerSegment[:, 4], 25).max() < 1e-15
    # silence: power 0 -> log power ln(0+1) = 0, mel sum exactly 0 -> LogMin (-10)
    se = _dft_power(np.zeros(n))
    assert np.all(se.LogPowerSegment == 0.0) and np.all(se.MelFBankSegment == -10.0)


def test_flat_spectrum_mel_is_log_of_weight_sum():
    x = np.zeros(32000)
    x[1600 + 5] = 1.0
    se = _dft_power(x)
    m = o.MelParams()
    f = m.InitFilters(400, 16000)
    for flt in range(32):
        w = BINPTS_16K[flt + 2] - BINPTS_16K[flt] + 1
        assert abs(se.MelFBankSegment[flt, 2] - math.log(f[flt, :w].sum())) < 1e-12


def test_smoothing_recurrence_restarts_each_segment():
    sig = synth.config1_signal().astype(np.float64)
    raw = o.make_env(sig, gabor=False, mfcc=False)
    sm = o.make_env(sig, gabor=False, mfcc=False, prev_smooth=0.3)
    raw.ProcessSegment(3)
    sm.ProcessSegment(3)
    p, q = raw.PowerSegment, sm.PowerSegment
    assert np.array_equal(q[:, 0], p[:, 0])                       # step 0: no smoothing (dft.go:67)
    for s in range(1, 14):
        assert np.allclose(q[:, s], 0.3 * q[:, s - 1] + 0.7 * p[:, s], rtol=1e-14)


def test_front_pad_and_tail_break():
    sig = synth.config1_signal().astype(np.float64)
    se = o.make_env(sig, gabor=False, mfcc=False)
    se.ProcessSegment(0)
    # step 0 starts at -320: window = 320 zeros + 80 samples
    w = np.concatenate([np.zeros(320), sig[:80]])
    assert np.allclose(se.PowerSegment[:, 0], np.abs(np.fft.fft(w)[:201]) ** 2, rtol=1e-9, atol=1e-12)
    se.ProcessSegment(19)                                          # last segment: steps 10.. run past the end
    assert np.all(se.MelFBankSegment[:, 10:] == 0.0) and np.all(se.PowerSegment[:, 10:] == 0.0)
    assert np.all(se.MelFBankSegment[:, 9] != 0.0)


def test_energy_uses_transposed_indexing_and_overwrites_c0():
    sig = synth.config1_signal().astype(np.float64)
    se = o.make_env(sig, gabor=False, mfcc=True, deltas=False)
    se.ProcessSegment(2)
    # Energy[s] = sum over steps of LogPower at BIN s, and MFCC row 0 is Energy
    assert np.allclose(se.Energy, se.LogPowerSegment[:14, :].sum(axis=1))
    assert np.array_equal(se.MFCCSegment[0], se.Energy)
    # rows 1.. are DCT-I of the log-mel column
    col = se.MelFBankSegment[:, 5]
    assert np.allclose(se.MFCCSegment[1:, 5], scipy.fft.dct(col, type=1)[1:13])


def test_delta_quirk_last_n_wins_and_accumulators_carry():
    M = np.arange(13 * 14, dtype=np.float64).reshape(13, 14) ** 1.5
    D = np.zeros_like(M)
    o._deltas(M, D, 13, 14)
    s = 5
    prv = nxt = 0.0
    for i in range(13):
        nume = 0.0
        for n in (1, 2):
            prv += M[i, s - n]
            nxt += M[i, s + n]
            nume += n * (nxt - prv)
        assert D[i, s] == nume / 8.0                               # n = 2 denominator, prv/nxt carried over i


def test_gabor_filters_lobes_normalised_and_convolve_layouts():
    fs = o.GaborFilterSet(SizeX=9, SizeY=9, StrideX=3, StrideY=3, Gain=2.0)
    o.gabor_to_tensor(o.processspeech_gabor_specs(), fs)
    assert fs.Filters.shape == (8, 9, 9)
    for f in fs.Filters:
        assert abs(f[f > 0].sum() - 1.0) < 1e-12 and abs(f[f < 0].sum() + 1.0) < 1e-12
    rng = np.random.default_rng(3)
    mel = rng.normal(size=(32, 14))
    out4 = np.zeros((8, 2, 2, 8), dtype=np.float32)
    assert o.gabor_convolve(mel, fs, out4, False)
    # on/off rectified pair: exactly one of the two slots is non-zero, value = Gain*|sum|
    acc = (fs.Filters[3] * mel[6:15, 3:12]).sum()
    assert abs(out4[2, 1, 0 if acc >= 0 else 1, 3] - 2.0 * abs(acc)) < 1e-5
    assert out4[2, 1, 1 if acc >= 0 else 0, 3] == 0.0
    out2 = np.zeros((16, 16), dtype=np.float32)
    out2t = np.zeros((16, 16), dtype=np.float32)
    assert o.gabor_convolve(mel, fs, out2, False) and o.gabor_convolve(mel, fs, out2t, True)
    assert np.array_equal(out2[:, 3 + 1 * 8], out2t[:, 1 + 2 * 3])        # x = flt + tIdx*nf  vs  tIdx + tMaxStrides*flt
    # NaN input counts as 0.5 (gabor.go:278-280)
    mel2 = mel.copy(); mel2[7, 4] = np.nan
    mel3 = mel.copy(); mel3[7, 4] = 0.5
    a = np.zeros((8, 2, 2, 8), dtype=np.float32); b = np.zeros_like(a)
    o.gabor_convolve(mel2, fs, a, False); o.gabor_convolve(mel3, fs, b, False)
    assert np.array_equal(a, b)
    # a 5-D output is rejected without writing
    assert not o.gabor_convolve(mel, fs, np.zeros((1, 8, 2, 2, 8), dtype=np.float32), False)


def test_c_twin_matches_numpy_restatement():
    sig = synth.config1_signal().astype(np.float64)
    se = o.make_env(sig, mfcc=True, deltas=True, prev_smooth=0.3)
    ref = o.process_all(se, want_power=True)
    env = c_oracle.Env(c_oracle.with_processspeech_gabor(c_oracle.default_params(prev_smooth=0.3, cur_smooth=0.7)),
                       c_oracle.processspeech_specs())
    got = env.process(sig, want_power=True)
    for k in ref:
        scale = max(1.0, np.abs(ref[k]).max())
        assert np.abs(ref[k].reshape(got[k].shape) - got[k]).max() <= 1e-11 * scale, k
    assert np.array_equal(env.binpts, se.Mel.BinPts)
    assert np.array_equal(env.mel_filters, se.MelFilters)
    assert np.abs(env.gabor_filters - se.GaborFilters.Filters).max() == 0.0


@pytest.mark.parametrize("tag,prev", [("cfg1", 0.0), ("cfg1_smooth", 0.3)])
def test_oracle_reproduces_golden_fixtures(tag, prev):
    g = np.load(os.path.join(GOLDEN, tag + ".npz"))
    sig = synth.config1_signal().astype(np.float64)
    env = c_oracle.Env(c_oracle.with_processspeech_gabor(c_oracle.default_params(prev_smooth=prev, cur_smooth=1 - prev)),
                       c_oracle.processspeech_specs())
    got = env.process(sig, want_power=True)
    for k in ("mel", "energy", "mfcc", "deltas", "delta_deltas"):
        assert np.abs(got[k] - g[k]).max() <= 1e-10 * max(1.0, np.abs(g[k]).max()), k
    assert np.array_equal(got["gabor"].reshape(g["gabor"].shape), g["gabor"])
    assert np.abs(got["logpower"][19] - g["logpower_seg19"]).max() < 1e-11


def test_wav_fixture_tones_peak_at_expected_bin():
    """The reference's example WAVs (44.1 kHz -> WinSamples 1103, prime): oracle-only sanity, skipped
    where /root/reference is absent (e.g. on the GPU box)."""
    import wave as wavmod
    path = "/root/reference/examples/processspeech/sounds/2000.wav"
    if not os.path.exists(path):
        pytest.skip("reference assets not present")
    with wavmod.open(path) as w:
        sr, n = w.getframerate(), w.getnframes()
        pcm = np.frombuffer(w.readframes(n), dtype="<i2").astype(np.float64) / 0x7FFF   # sound.go:130-141
    assert sr == 44100
    N = o.msec_to_samples(25, sr)
    assert N == 1103
    frame = pcm[4410:4410 + N]
    pw = np.abs(c_oracle.fft(frame)[:N // 2 + 1]) ** 2
    assert int(np.argmax(pw)) == round(2000 * N / sr)
