"""Host side of the C-ABI (no GPU): exported symbols, the initialisers that mirror
the Go Init code, parameter validation and the reference's panic cases."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import auditory_b200 as ab
from auditory_b200 import _lib, agabor, mel, synth
from oracle import np_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "auditory_b200.h")).read()
    declared = set(re.findall(r"AUD_API\s+[\w\s\*]+?\b(aud_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().aud_version() >= 100


def test_msec_to_samples_and_mel_scale():
    assert ab.MSecToSamples(25, 16000) == 400 and ab.MSecToSamples(25, 44100) == 1103
    assert mel.FreqToMel(1000.0) == o.freq_to_mel(1000.0)
    assert mel.MelToFreq(2000.0) == o.mel_to_freq(2000.0)
    assert mel.FreqToBin(8000.0, 400.0, 16000.0) == 200


def test_mel_init_filters_bit_exact_with_oracle():
    for n, sr, nf in ((400, 16000, 32), (1103, 44100, 32), (200, 8000, 20)):
        a, b = mel.Params(), o.MelParams()
        a.FBank.NFilters = b.FBank.NFilters = nf
        a.FBank.HiHz = b.FBank.HiHz = sr / 2
        try:
            fb = b.InitFilters(n, sr)
        except IndexError:
            with pytest.raises(ab.AudError) as ei:
                a.InitFilters(n, sr)
            assert ei.value.code == _lib.AUD_ERR_PANIC
            continue
        fa = a.InitFilters(n, sr)
        assert np.array_equal(a.BinPts, b.BinPts) and np.array_equal(a.HzPts, b.HzPts)
        assert np.array_equal(fa, fb, equal_nan=True)
        assert a.FBank.Renorm is False                       # mel.go:80


def test_mel_init_filters_panic_cases():
    with pytest.raises(ab.AudError) as ei:
        mel.Params().InitFilters(512, 16000)
    assert ei.value.code == _lib.AUD_ERR_PANIC


@pytest.mark.parametrize("distribute", [False, True])
def test_gabor_to_tensor_matches_oracle(distribute):
    specs_a = synth.processspeech_gabor_specs() + [agabor.Filter(Off=True), agabor.Filter(Orientation=30.0),
                                                   agabor.Filter(Circular=True, WaveLen=1.5, SigmaWidth=0.6)]
    specs_b = o.processspeech_gabor_specs() + [o.GaborFilter(Off=True), o.GaborFilter(Orientation=30.0),
                                               o.GaborFilter(Circular=True, WaveLen=1.5, SigmaWidth=0.6)]
    fa = agabor.FilterSet(SizeX=7, SizeY=9, StrideX=2, StrideY=3, Gain=1.5, Distribute=distribute)
    fb = o.GaborFilterSet(SizeX=7, SizeY=9, StrideX=2, StrideY=3, Gain=1.5, Distribute=distribute)
    agabor.ToTensor(specs_a, fa)
    o.gabor_to_tensor(specs_b, fb)
    assert fa.Filters.shape == fb.Filters.shape == (10, 9, 7)
    assert np.abs(fa.Filters - fb.Filters).max() <= 4e-16 * max(1.0, np.abs(fb.Filters).max())


def test_dct1_matrix_is_the_oracle_transform():
    m = np.zeros((13, 32))
    _lib.lib().aud_dct1_matrix(32, 13, m.ctypes.data)
    x = np.random.default_rng(0).normal(size=32)
    assert np.abs(m @ x - o.dct1(x)[:13]).max() < 1e-12


def test_params_defaults_match_sndenv_defaults():
    p = _lib.AudParams()
    _lib.check(_lib.lib().aud_params_defaults(C.byref(p), 16000, 25.0, 10.0, 100.0, 100.0, 2))
    assert (p.win_samples, p.step_samples, p.segment_samples, p.stride_samples, p.segment_steps) == (400, 160, 1600, 1600, 14)
    assert (p.log_offset, p.log_min, p.prev_smooth, p.cur_smooth) == (1.0, -100.0, 0.0, 1.0)
    assert (p.n_mel, p.mel_log_min, p.mfcc, p.deltas, p.n_coefs, p.renorm) == (32, -10.0, 1, 1, 13, 0)
    with pytest.raises(ab.AudError):
        _lib.check(_lib.lib().aud_params_defaults(C.byref(p), 0, 25.0, 10.0, 100.0, 100.0, 2))


def _env(**kw):
    se = ab.SndEnv()
    se.Defaults()
    se.SetSignal(np.zeros(32000, dtype=np.float32), kw.pop("sr", 16000))
    for k, v in kw.items():
        setattr(se.Params, k, v)
    return se


def test_sndenv_init_shapes_and_segcnt():
    se = _env()
    synth.configure_processspeech_gabor(se)
    se.Init()
    assert se.SegCnt == 20 and se.Params.Steps[:3] == [-320, -160, 0]
    assert se.MelFBankSegment.shape == (32, 14) and se.MFCCSegment.shape == (13, 14)
    assert se.GborOutput.shape == (8, 2, 2, 8) and se.PowerSegment.shape == (201, 14)
    se.DFT.PrevSmooth = 0.3
    se.Init()
    assert se.DFT.PrevSmooth == 0.0                          # Init calls DFT.Defaults()
    assert se.Pad(np.zeros(32100), 0.0).size == 32100 + 1600 - 160 - (32100 - 1600) % 1600 % 160


def test_create_rejects_what_the_reference_cannot_run():
    L = _lib.lib()

    def create(se):
        return ab.Pipeline(se.aud_params(), se.Mel.BinPts, se.MelFilters,
                           se.GaborFilters.Filters if se._n_gabor else None)

    # valid parameters: no GPU here, so creation must fail loudly with AUD_ERR_CUDA (never a CPU fallback)
    se = _env()
    se.Init()
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(ab.AudError) as ei:
            create(se)
        assert ei.value.code == _lib.AUD_ERR_CUDA and "no CPU fallback" in ei.value.msg
    # other window lengths take the general path: valid, but still no CPU fallback
    se = _env(sr=8000)
    se.Mel.FBank.HiHz = 4000.0
    se.Mel.FBank.NFilters = 20
    se.Init()
    assert se.Params.WinSamples == 200
    if not torch.cuda.is_available():
        with pytest.raises(ab.AudError) as ei:
            create(se)
        assert ei.value.code == _lib.AUD_ERR_CUDA
    # windows longer than 4096 samples are outside both kernels (documented)
    se = _env(sr=192000)
    se.Mel.FBank.HiHz = 96000.0
    se.Mel.FBank.NFilters = 300
    se.Mel.MFCC = False
    se.Init()
    assert se.Params.WinSamples == 4800
    with pytest.raises(ab.AudError) as ei:
        create(se)
    assert ei.value.code == _lib.AUD_ERR_UNSUPPORTED
    # SegmentSteps > WinSamples/2+1 with MFCC: ProcessSegment's Energy loop panics in Go
    se = _env(SegmentMs=2500.0)
    se.Init()
    with pytest.raises(ab.AudError) as ei:
        create(se)
    assert ei.value.code == _lib.AUD_ERR_PANIC
    # gabor filter larger than the mel tile in frequency: Convolve reads past the tensor
    se = _env()
    synth.configure_processspeech_gabor(se)
    se.GaborFilters.SizeY = 31
    se.GaborFilters.StrideY = 1
    se.GborOutPoolsY = 8
    se.Init()
    with pytest.raises(ab.AudError) as ei:
        create(se)
    assert ei.value.code == _lib.AUD_ERR_PANIC
    # NULL arguments
    assert L.aud_create(None, None, None, None, None, 0, C.byref(C.c_void_p())) == _lib.AUD_ERR_INVALID
    assert L.aud_seg_count(None, 10) == _lib.AUD_ERR_INVALID
