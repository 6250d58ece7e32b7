"""sound.Wave mirror (sound/sound.go:32-141): WAV decode and the int -> float normalisation the
reference applies before the feature path, checked against files written by Python's wave module."""
import struct
import wave

import numpy as np
import pytest

import auditory_b200 as ab


def write_wav(path, data, rate, channels, width):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(width)
        w.setframerate(rate)
        w.writeframes(data)


def test_pcm16_mono_and_interleaved_stereo_quirk(tmp_path):
    rng = np.random.default_rng(0)
    x = rng.integers(-32768, 32768, 4410, dtype=np.int16)
    write_wav(tmp_path / "m.wav", x.astype("<i2").tobytes(), 44100, 1, 2)
    w = ab.Wave()
    w.Load(str(tmp_path / "m.wav"))
    assert (w.SampleRate(), w.Channels(), w.NumFrames(), w.SourceBitDepth) == (44100, 1, 4410, 16)
    sig = w.SoundToTensor()
    assert sig.dtype == np.float64 and np.array_equal(sig, x.astype(np.float64) / 0x7FFF)     # sound.go:135-136
    assert w.GetFloatAtIdx(17) == float(x[17]) / 0x7FFF and np.array_equal(w.pcm16(), x)
    # stereo: SoundToTensor reads Data[i] for i < NumFrames, i.e. the first half of the interleaved samples
    write_wav(tmp_path / "s.wav", x[:4000].astype("<i2").tobytes(), 16000, 2, 2)
    w.Load(str(tmp_path / "s.wav"))
    assert (w.Channels(), w.NumFrames()) == (2, 2000)
    assert np.array_equal(w.SoundToTensor(), x[:2000].astype(np.float64) / 0x7FFF)


def test_pcm24_pcm32_pcm8_and_extra_chunks(tmp_path):
    v = np.array([0, 1, -1, 8388607, -8388608, 123456, -654321], dtype=np.int32)
    raw = b"".join(struct.pack("<i", int(a))[:3] for a in v)
    write_wav(tmp_path / "a.wav", raw, 8000, 1, 3)
    w = ab.Wave()
    w.Load(str(tmp_path / "a.wav"))
    assert np.array_equal(w.Data, v) and np.array_equal(w.SoundToTensor(), v / float(0x7FFFFF))
    v32 = np.array([0, 2147483647, -2147483648, 77], dtype=np.int32)
    write_wav(tmp_path / "b.wav", v32.astype("<i4").tobytes(), 8000, 1, 4)
    w.Load(str(tmp_path / "b.wav"))
    assert np.array_equal(w.SoundToTensor(), v32 / float(0x7FFFFFFF))
    write_wav(tmp_path / "c.wav", bytes([0, 127, 128, 255]), 8000, 1, 1)
    w.Load(str(tmp_path / "c.wav"))
    assert np.array_equal(w.SoundToTensor(), np.array([0, 127, 128, 255]) / float(0x7F))
    # a LIST chunk between fmt and data (common in the wild) is skipped
    body = open(tmp_path / "b.wav", "rb").read()
    i = body.index(b"data")
    patched = body[:i] + b"LIST" + struct.pack("<I", 5) + b"hello\x00" + body[i:]
    patched = patched[:4] + struct.pack("<I", len(patched) - 8) + patched[8:]
    (tmp_path / "d.wav").write_bytes(patched)
    w.Load(str(tmp_path / "d.wav"))
    assert np.array_equal(w.Data, v32)
    (tmp_path / "e.wav").write_bytes(b"not a wave file")
    with pytest.raises(ValueError):
        w.Load(str(tmp_path / "e.wav"))


def test_sndenv_totensor_pad_and_silence(tmp_path):
    x = (np.sin(np.arange(33000) * 0.05) * 12000).astype(np.int16)
    write_wav(tmp_path / "t.wav", x.astype("<i2").tobytes(), 16000, 1, 2)
    se = ab.SndEnv()
    se.Defaults()
    se.Sound.Load(str(tmp_path / "t.wav"))
    se.ToTensor()
    se.Init()
    assert se.SampleRate == 16000 and se.Signal.dtype == np.float32 and se.Signal.size == 33000
    assert se.SegCnt == (33000 - 1600) // 1600 + 1
    assert se.AdjustForSilence(30.0, 10.0) == 20 and se.Signal.size == 33000 + 320 and np.all(se.Signal[:320] == 0)
    assert se.AdjustForSilence(0.0, 20.0) == 20 and se.Signal.size == 33000
    assert se.AdjustForSilence(-1.0, 20.0) == 0 and se.Signal.size == 33000


@pytest.mark.gpu
def test_wav_44k_through_the_pipeline_float_and_pcm16(tmp_path):
    from oracle import c_oracle
    from util import RTOL_LOG, assert_close
    sr = 44100
    t = np.arange(int(0.6 * sr)) / sr
    x = np.round(9000 * np.sin(2 * np.pi * 2000.0 * t) + 3000 * np.sin(2 * np.pi * 440.0 * t + 1.0)).astype(np.int16)
    write_wav(tmp_path / "tone.wav", x.astype("<i2").tobytes(), sr, 1, 2)
    se = ab.SndEnv(device=0)
    se.Defaults()
    se.Mel.MFCC = False
    se.Sound.Load(str(tmp_path / "tone.wav"))
    se.ToTensor()
    se.Init()
    assert se.Params.WinSamples == 1103
    got = se.ProcessBatch(se.Signal, [0], [se.Signal.size], want=["mel"])["mel"]
    p = c_oracle.default_params(sample_rate=sr, mfcc=0, deltas=0)
    ref = c_oracle.Env(p, []).process(se.Signal.astype(np.float64))["mel"].reshape(got.shape)
    assert_close(got, ref, RTOL_LOG, "mel")
    # the 2 kHz tone lands in the filter whose centre bin is nearest to 2000 Hz
    bp = np.asarray(se.Mel.BinPts)
    k = int(np.floor((1103 + 1) * 2000.0 / sr))
    assert abs(int(bp[1 + int(got[1].mean(axis=1).argmax())]) - k) <= 2
    # 16-bit samples straight to the GPU: same features
    pcm = se.Sound.pcm16()
    got16 = se.pipeline().process_host(pcm, [0], np.asarray([pcm.size], np.int32), want=("mel",))["mel"]
    assert_close(got16, ref, RTOL_LOG, "mel (pcm16)")
