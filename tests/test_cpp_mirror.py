"""The C++ mirror of the Go API (include/auditory/auditory.hpp) compiles against the C-ABI and, on a
GPU, reproduces the Python mirror's results for the reference's call sequence."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "sndenv_demo")


def build():
    lib = os.path.join(ROOT, "auditory_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "sndenv_demo.cpp"), "-o", EXE,
                           "-L", lib, "-lauditory_b200", f"-Wl,-rpath,{lib}"])


def test_cpp_mirror_builds_and_inits_without_gpu():
    build()
    out = subprocess.run([EXE, "32000", "init-only"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["SegCnt", "20", "SegmentSteps", "14", "WinSamples", "400"]


def test_cpp_wave_load_matches_python_mirror(tmp_path):
    """sound.Wave.Load + SoundToTensor + SndEnv.Init from a 44.1 kHz 16-bit file, C++ against Python."""
    import wave
    import auditory_b200 as ab
    build()
    x = (np.sin(np.arange(50000) * 0.03) * 9000).astype(np.int16)
    fn = str(tmp_path / "t.wav")
    with wave.open(fn, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(44100)
        w.writeframes(x.astype("<i2").tobytes())
    out = subprocess.run([EXE, "0", "wav", fn], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    f = out.stdout.split()
    se = ab.SndEnv()
    se.Defaults()
    se.Sound.Load(fn)
    se.ToTensor()
    se.Init()
    assert [int(f[i]) for i in (1, 3, 5, 7, 9, 11)] == [44100, 1, 50000, 16, se.SegCnt, 1103]
    assert abs(float(f[13]) - float(se.Signal.astype(np.float64).sum())) < 1e-3


@pytest.mark.gpu
def test_cpp_mirror_matches_python_mirror():
    import auditory_b200 as ab
    from auditory_b200 import synth
    if not os.path.exists(EXE):
        build()
    out = subprocess.run([EXE, "32000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = [l.split() for l in out.stdout.strip().splitlines()[1:]]
    n = 32000
    s = np.uint32(12345)
    sig = np.zeros(n, dtype=np.float32)
    with np.errstate(over="ignore"):
        for i in range(n):
            s = np.uint32(s * np.uint32(1664525) + np.uint32(1013904223))
            sig[i] = np.float32(0.1) * (np.float32(int(s) >> 8) / np.float32(8388608.0) - np.float32(1.0)) + \
                np.float32(0.3 * np.sin(2.0 * np.pi * 1000.0 * i / 16000.0))
    se = ab.SndEnv()
    se.Defaults()
    se.SetSignal(sig, 16000)
    synth.configure_processspeech_gabor(se)
    se.Init()
    for l in lines:
        seg = int(l[1])
        se.ProcessSegment(seg, 0)
        g = se.ApplyGabor()
        assert abs(float(l[3]) - float(se.MelFBankSegment.astype(np.float64).sum())) < 2e-2
        assert abs(float(l[5]) - float(se.MFCCSegment.astype(np.float64).sum())) < 5e-2
        assert abs(float(l[7]) - float(g.astype(np.float64).sum())) < 2e-2
