"""Synthetic inputs and stock configurations shared by tests/ and bench.py
(SURVEY 8d).  numpy only; nothing here touches the GPU or the oracle."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from . import agabor

SR = 16000


def processspeech_gabor_specs() -> List[agabor.Filter]:
    """Parameter values of examples/processspeech/processspeech.go:226-253:
    4 orientations x 2 phases, WaveLen 2, sigma 0.5, CircleEdge."""
    specs = []
    for orient in (0.0, 45.0, 90.0, 135.0):
        for ph in (0.0, 1.5708):
            specs.append(agabor.Filter(WaveLen=2.0, Orientation=orient, SigmaWidth=0.5, SigmaLength=0.5,
                                       PhaseOffset=ph, CircleEdge=True))
    return specs


def configure_processspeech_gabor(se, out4d: bool = True, by_time: bool = False) -> None:
    """Gabor set of config 1: 9x9, stride 3, gain 2; 4-D out [8,2,2,8] or 2-D [16,16]."""
    se.GaborSpecs = processspeech_gabor_specs()
    gf = se.GaborFilters
    gf.SizeX = gf.SizeY = 9
    gf.StrideX = gf.StrideY = 3
    gf.Gain = 2.0
    gf.Distribute = False
    se.ByTime = by_time
    if out4d:
        se.GborOutPoolsY, se.GborOutPoolsX, se.GborOutUnitsY, se.GborOutUnitsX = 8, 2, 2, 8
    else:
        se.GborOutPoolsY = se.GborOutPoolsX = 0
        se.GborOutUnitsY, se.GborOutUnitsX = 16, 16


def config1_signal(seed: int = 1234, seconds: float = 2.0, sr: int = SR) -> np.ndarray:
    """Config 1: three sines (440, 1800, 5200 Hz, amp 0.25, random phase) + N(0, 0.05^2), clipped."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    t = np.arange(n, dtype=np.float64) / sr
    sig = np.zeros(n)
    for fr in (440.0, 1800.0, 5200.0):
        sig += 0.25 * np.sin(2 * np.pi * fr * t + rng.uniform(0, 2 * np.pi))
    sig += rng.normal(0.0, 0.05, n)
    return np.clip(sig, -1.0, 1.0).astype(np.float32)


def batch_utterance(u: int, seed_base: int = 1000, seconds: float = 3.0, sr: int = SR) -> np.ndarray:
    """Configs 2-4: uniform(-1,1) noise x random gain (0.1-0.9) + one random sine."""
    rng = np.random.default_rng(seed_base + u)
    n = int(round(seconds * sr))
    gain = rng.uniform(0.1, 0.9)
    fr = rng.uniform(100.0, 7000.0)
    t = np.arange(n, dtype=np.float64) / sr
    sig = gain * rng.uniform(-1.0, 1.0, n) * 0.5 + 0.3 * np.sin(2 * np.pi * fr * t + rng.uniform(0, 2 * np.pi))
    return np.clip(sig, -1.0, 1.0).astype(np.float32)


def batch(n_utt: int, seed_base: int = 1000, seconds: float = 3.0, sr: int = SR) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(wave, utt_offset, utt_len) for n_utt equal-length utterances."""
    n = int(round(seconds * sr))
    wave = np.empty(n_utt * n, dtype=np.float32)
    for u in range(n_utt):
        wave[u * n:(u + 1) * n] = batch_utterance(u, seed_base, seconds, sr)
    return wave, np.arange(n_utt, dtype=np.int64) * n, np.full(n_utt, n, dtype=np.int32)


def fast_batch(n_utt: int, seed: int = 1000, seconds: float = 3.0, sr: int = SR) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Same statistics as batch() but generated in one vectorised pass (for the
    1024- and 65,536-utterance bench workloads, where per-utterance seeding
    would take minutes on the host)."""
    n = int(round(seconds * sr))
    rng = np.random.default_rng(seed)
    wave = rng.random((n_utt, n), dtype=np.float32)
    wave -= 0.5
    gain = rng.uniform(0.1, 0.9, (n_utt, 1)).astype(np.float32)
    wave *= gain
    fr = rng.uniform(100.0, 7000.0, (n_utt, 1)).astype(np.float32)
    ph = rng.uniform(0, 2 * np.pi, (n_utt, 1)).astype(np.float32)
    t = (np.arange(n, dtype=np.float32) / np.float32(sr))[None, :]
    step = max(1, 4096 // max(1, n // 4096 + 1))
    for i in range(0, n_utt, 64):
        wave[i:i + 64] += np.float32(0.3) * np.sin(np.float32(2 * np.pi) * fr[i:i + 64] * t + ph[i:i + 64])
    del step
    np.clip(wave, -1.0, 1.0, out=wave)
    return wave.reshape(-1), np.arange(n_utt, dtype=np.int64) * n, np.full(n_utt, n, dtype=np.int32)


def long_signal(seconds: float = 600.0, seed: int = 5, sr: int = SR) -> np.ndarray:
    """Config 5: chirp 100 -> 7000 Hz + noise."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    t = np.arange(n, dtype=np.float64) / sr
    f0, f1 = 100.0, 7000.0
    phase = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / seconds * t * t)
    sig = 0.5 * np.sin(phase) + rng.normal(0.0, 0.05, n)
    return np.clip(sig, -1.0, 1.0).astype(np.float32)
