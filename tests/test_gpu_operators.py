"""Stand-alone operators behind the C-ABI: agabor.Convolve on mel tensors the caller already
holds, as examples/gaborview does, against the oracle's restatement of agabor/gabor.go:225-315."""
import numpy as np
import pytest

import auditory_b200 as ab
from auditory_b200 import agabor, synth
from oracle import np_oracle
from test_gpu_parity import compare, make_env, oracle_batch, oracle_env
from util import RTOL_GABOR, assert_close

pytestmark = pytest.mark.gpu


def filter_set(size=9, stride=3, gain=2.0):
    fs = agabor.FilterSet(SizeX=size, SizeY=size, StrideX=stride, StrideY=stride, Gain=gain)
    agabor.ToTensor(synth.processspeech_gabor_specs(), fs)
    return fs


def oracle_convolve(mel, fs, shape, by_time, init=0.0):
    ofs = np_oracle.GaborFilterSet()
    ofs.SizeX, ofs.SizeY, ofs.StrideX, ofs.StrideY, ofs.Gain = fs.SizeX, fs.SizeY, fs.StrideX, fs.StrideY, fs.Gain
    ofs.Filters = np.asarray(fs.Filters, dtype=np.float64)
    out = np.full(shape, init, dtype=np.float32)
    wrote = np_oracle.gabor_convolve(mel.astype(np.float64), ofs, out, by_time)
    return out, wrote


@pytest.mark.parametrize("shape,by_time,steps", [((8, 2, 2, 8), False, 14), ((16, 16), False, 14), ((16, 48), True, 24),
                                                  ((16, 520), True, 200)])
def test_convolve_matches_oracle(shape, by_time, steps):
    rng = np.random.default_rng(steps)
    fs = filter_set()
    mel = rng.normal(-2.0, 3.0, (3, 32, steps)).astype(np.float32)
    mel[1, 5, 3] = np.nan                                       # gabor.go:283-285: NaN inputs count as 0.5
    out = np.full((3,) + shape, 7.0, dtype=np.float32)          # cells Convolve does not reach keep their values
    agabor.Convolve(mel, fs, out, by_time)
    for k in range(3):
        ref, wrote = oracle_convolve(mel[k], fs, shape, by_time, init=7.0)
        assert wrote
        assert_close(out[k], ref, RTOL_GABOR, f"Convolve[{k}]")
        assert np.array_equal(out[k] == 7.0, ref == 7.0)
    one = np.zeros(shape, dtype=np.float32)
    agabor.Convolve(mel[0], fs, one, by_time)
    ref, _ = oracle_convolve(mel[0], fs, shape, by_time)
    assert_close(one, ref, RTOL_GABOR, "Convolve (single tensor)")


def test_convolve_same_as_sndenv_applygabor():
    sig = synth.config1_signal()
    se = ab.SndEnv(device=0)
    se.Defaults()
    se.SetSignal(sig, synth.SR)
    se.Mel.MFCC = False
    synth.configure_processspeech_gabor(se)
    se.Init()
    got = se.ProcessBatch(sig, [0], [sig.size], want=["mel", "gabor"])
    out = np.zeros((got["mel"].shape[0], 8, 2, 2, 8), dtype=np.float32)
    agabor.Convolve(got["mel"], se.GaborFilters, out, False)
    assert np.array_equal(out.reshape(got["gabor"].shape), got["gabor"])


def test_convolve_edge_cases():
    fs = filter_set()
    narrow = np.ones((32, 5), dtype=np.float32)                 # filter wider than the input: logs and returns
    out = np.full((16, 16), 3.0, dtype=np.float32)
    agabor.Convolve(narrow, fs, out, False)
    assert np.all(out == 3.0)
    with pytest.raises(ab.AudError) as ei:                      # output tensor too small: the reference panics
        agabor.Convolve(np.ones((32, 14), dtype=np.float32), fs, np.zeros((2, 2), dtype=np.float32), False)
    assert ei.value.code == ab._lib.AUD_ERR_PANIC
    with pytest.raises(ab.AudError):
        agabor.Convolve(np.ones((32, 14), dtype=np.float32), fs, np.zeros((4, 4, 4), dtype=np.float32), False)


def test_process_segment_cache_follows_the_signal():
    """SndEnv.ProcessSegment serves segments from one batched GPU call per signal; a new Signal of the same length, or
    samples edited in place, must not be answered from the previous signal's features."""
    a = synth.config1_signal(seed=1)
    b = synth.config1_signal(seed=2)
    se = make_env(mfcc=False, gabor=True)
    se.SetSignal(a, synth.SR)
    se.Init()
    se.ProcessSegment(3, 0)
    mel_a = se.MelFBankSegment.copy()
    gab_a = se.ApplyGabor().copy()
    se.Signal = b.copy()                      # plain attribute assignment, as Go code does with se.Signal.Values
    se.ProcessSegment(3, 0)
    mel_b = se.MelFBankSegment.copy()
    assert not np.array_equal(mel_a, mel_b) and not np.array_equal(gab_a, se.ApplyGabor())
    se.Signal[3 * 1600:4 * 1600] = a[3 * 1600:4 * 1600]     # in-place edit of the samples segment 3 covers
    se.ProcessSegment(3, 0)
    assert not np.array_equal(se.MelFBankSegment, mel_b)
    se.Signal = a
    se.ProcessSegment(3, 0)
    assert np.array_equal(se.MelFBankSegment, mel_a)


def test_many_tiny_jobs_complete_in_one_round():
    """SegmentMs = StepMs with no border: one-step segments, so a round of 72 frames finishes dozens of segments that
    belong to dozens of one- and two-segment utterances (more than the 16 jobs per round the done list once held)."""
    rng = np.random.default_rng(21)
    lens = rng.integers(400, 1000, 300).astype(np.int32)
    off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    wave = rng.uniform(-0.7, 0.7, int(lens.sum())).astype(np.float32)
    kw = dict(mfcc=True, deltas=False, gabor=False, prev=0.3, SegmentMs=10.0, StrideMs=10.0, BorderSteps=0)
    se = make_env(**kw)
    got = se.ProcessBatch(wave, off, lens, want=["mel", "mfcc", "energy"])
    env = oracle_env(**kw)
    ref = oracle_batch(env, wave, off, lens)
    assert got["mel"].shape == ref["mel"].shape and got["mel"].shape[1:] == (32, 1) and got["mel"].shape[0] > 400
    compare(got, ref, ["mel", "mfcc", "energy"])


@pytest.mark.parametrize("sr,prev", [(16000, 0.0), (16000, 0.3), (44100, 0.2)])
def test_per_step_operators_follow_the_gaborview_loop(sr, prev):
    """examples/gaborview/gbv.go:545-559, 627-641: the caller cuts the windows itself and calls Dft.Filter ->
    Mel.FilterDft -> Mel.CepstrumDct once per step over one long segment.  The batched operators cover all steps of
    the segment in one call each; compared with the oracle's per-step restatement (dft.go:42-85, mel.go:120-212)."""
    from auditory_b200 import dft, mel
    rng = np.random.default_rng(sr + int(prev * 10))
    win, hop, steps = int(round(0.025 * sr)), int(round(0.010 * sr)), 57
    sig = (rng.uniform(-1, 1, hop * steps + win) * 0.4 + 0.2 * np.sin(np.arange(hop * steps + win) * 0.05)).astype(np.float32)
    windows = np.stack([sig[i * hop:i * hop + win] for i in range(steps)])
    # reference, step by step
    od = np_oracle.DftParams()
    od.Defaults()
    od.PrevSmooth, od.CurSmooth = prev, 1.0 - prev
    om = np_oracle.MelParams()
    om.FBank = np_oracle.MelFilterBank()
    om.MFCC, om.Deltas, om.NCoefs = True, False, 13
    filters = om.InitFilters(win, sr)
    bins, nf = win // 2 + 1, om.FBank.NFilters
    power, logp = np.zeros(bins), np.zeros(bins)
    pseg, lseg = np.zeros((bins, steps)), np.zeros((bins, steps))
    mseg, cseg, mfb = np.zeros((nf, steps)), np.zeros((13, steps)), np.zeros(nf)
    for s in range(steps):
        od.Filter(s, windows[s].astype(np.float64), win, power, logp, pseg, lseg)
        om.FilterDft(s, power, mseg, mfb, filters)
        om.CepstrumDct(s, mfb, cseg)
    # GPU, one call per operator
    gd = dft.Params()
    gd.Defaults()
    gd.PrevSmooth, gd.CurSmooth = prev, 1.0 - prev
    gm = mel.Params()
    gm.Defaults()
    gfilters = gm.InitFilters(win, sr)
    assert np.array_equal(gfilters, filters)
    g_pow, g_log = gd.FilterSegment(windows)
    scale = np.maximum(np.abs(pseg).max(axis=0, keepdims=True), 1.0)
    assert np.all(np.abs(g_pow - pseg) <= 2e-5 * scale), "PowerSegment"
    assert_close(g_log, lseg, 1e-4, "LogPowerSegment")
    g_mel = gm.FilterDftSegment(g_pow, gfilters)
    assert_close(g_mel, mseg, 1e-4, "MelFBankSegment")
    g_cep = gm.CepstrumDctSegment(g_mel)
    assert_close(g_cep, cseg, 1e-4, "MFCCSegment")
    # and the operators agree with themselves when fed the oracle's intermediate tensors
    assert_close(gm.FilterDftSegment(pseg.astype(np.float32), gfilters), mseg, 1e-4, "FilterDft on reference power")
    assert_close(gm.CepstrumDctSegment(mseg.astype(np.float32)), cseg, 1e-4, "CepstrumDct on reference mel")


@pytest.mark.parametrize("pool,neigh", [(False, False), (True, False), (True, True), (False, True)])
def test_neighbour_inhibition_and_kwta_after_gabor(pool, neigh):
    """SndEnv.ApplyGabor's tail (sndenv.go:303-323, 481-497): NeighInhib.Inhib4 then KWTALayer / KWTAPool on the gabor
    output of every segment, in segment order.  The oracle twin restates emer/vision kwta + leabra fffb / nxx1 from
    the published equations (third-party packages absent from the reference tree: parity unpinned)."""
    sig = synth.config1_signal()
    se = make_env(mfcc=False, gabor=True)
    se.SetSignal(sig, synth.SR)
    se.Init()
    se.Kwta.Defaults()
    se.KwtaPool = pool
    if neigh:
        se.NeighInhib.Defaults()
    outs, exts = [], []
    for seg in range(se.SegCnt):
        se.ProcessSegment(seg, 0)
        outs.append(se.ApplyGabor().copy())
        exts.append(se.ExtGi.copy())
    raw = se._cache["gabor"].reshape(-1, 8, 2, 2, 8)
    k, ni, inhibs = np_oracle.KWTA(), np_oracle.NeighInhib(), []
    with np.errstate(over="ignore"):
        for seg in range(se.SegCnt):
            ext = np.zeros((8, 2, 2, 8), dtype=np.float32)
            if neigh:
                ni.Inhib4(raw[seg], ext)
            act = raw[seg].copy()
            if pool:
                k.KWTAPool(raw[seg], act, inhibs, ext)
            else:
                k.KWTALayer(raw[seg], act, ext)
            assert np.array_equal(exts[seg], ext), ("ExtGi", seg)
            assert_close(outs[seg], act, 1e-4, f"GborKwta[{seg}] pool={pool} neigh={neigh}")
    assert any(o.max() > 0.2 for o in outs) and np.mean([np.mean(o > 0.05) for o in outs]) < 0.5   # sparse winners
