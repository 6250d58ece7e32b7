#!/usr/bin/env python
"""Small cases for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): config 1 and a ragged batch
through the three launch shapes the library ships -- plain log-mel (12 FFT + 4 epilogue warps, records by the
epilogue), mel + gabor (12 + 4) and MFCC + deltas + smoothing + Energy (10 + 6) -- plus the int16 entry point,
the general window-length route and the stand-alone gabor operator.

  compute-sanitizer --tool racecheck python tools/sanitize_cases.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import auditory_b200 as ab
from auditory_b200 import synth


def env(mfcc, gabor, prev, sr=synth.SR):
    se = ab.SndEnv(device=0)
    se.Defaults()
    se.SetSignal(np.zeros(3 * sr, dtype=np.float32), sr)
    se.Mel.MFCC = mfcc
    se.Mel.Deltas = mfcc
    if gabor:
        synth.configure_processspeech_gabor(se)
    se.Init()
    se.DFT.PrevSmooth, se.DFT.CurSmooth = prev, 1.0 - prev
    return se


def ragged():
    rng = np.random.default_rng(3)
    lens = np.array([48000, 16001, 1700, 999, 0, 33333, 2000, 1601], dtype=np.int32)
    off, pos = [], 5
    for n in lens:
        off.append(pos)
        pos += int(n)
    wave = rng.uniform(-0.5, 0.5, pos + 8).astype(np.float32)
    wave[off[1] + 3000:off[1] + 9000] *= 1e-4          # a level step: the alone-frame path
    return wave, np.array(off, dtype=np.int64), lens


def main():
    sig = synth.config1_signal()
    wave, off, ln = ragged()
    total = 0.0
    for name, kw, want in (("mel", dict(mfcc=False, gabor=False, prev=0.0), ["mel"]),
                           ("gabor", dict(mfcc=False, gabor=True, prev=0.0), ["mel", "gabor"]),
                           ("mfcc", dict(mfcc=True, gabor=True, prev=0.3), ["mel", "mfcc", "deltas", "delta_deltas", "energy", "gabor"])):
        se = env(**kw)
        pipe = se.pipeline()
        a = pipe.process_host(sig, [0], [sig.size], want=want)
        b = pipe.process_host(wave, off, ln, want=want)
        pcm = np.round(wave * 20000).astype(np.int16)
        c = pipe.process_host(pcm, off, ln, want=want)
        total += float(sum(np.nansum(v) for d in (a, b, c) for v in d.values()))
        print(name, "ok", {k: v.shape for k, v in b.items()}, flush=True)
    # general route (44.1 kHz: 1103-sample window)
    se = env(False, False, 0.0, sr=44100)
    g = se.pipeline().process_host(np.random.default_rng(1).uniform(-1, 1, 44100).astype(np.float32), [0], [44100], want=["mel"])
    total += float(g["mel"].sum())
    print("generic ok", g["mel"].shape, "checksum", total)


if __name__ == "__main__":
    main()
