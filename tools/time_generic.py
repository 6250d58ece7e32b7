#!/usr/bin/env python
"""Device-resident timing of the general-window-length path (not a bench: one-off numbers for DESIGN.md)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import auditory_b200 as ab
from auditory_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--sr", type=int, default=44100)
ap.add_argument("--n-utt", type=int, default=256)
ap.add_argument("--seconds", type=float, default=3.0)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--workload", default="mel")
a = ap.parse_args()
se = ab.SndEnv(device=0)
se.Defaults()
se.SetSignal(np.zeros(a.sr, dtype=np.float32), a.sr)
se.Mel.MFCC = a.workload == "mfcc"
se.Mel.Deltas = False
if a.workload == "gabor":
    synth.configure_processspeech_gabor(se)
se.Init()
want = {"mel": ["mel"], "mfcc": ["mel", "mfcc"], "gabor": ["mel", "gabor"]}[a.workload]
pipe = se.pipeline()
wave_h, off, ln = synth.fast_batch(a.n_utt, seed=1000, seconds=a.seconds, sr=a.sr)
nseg = int(pipe.seg_base(ln)[-1])
dev = torch.device("cuda", 0)
wave = torch.from_numpy(wave_h).to(dev)
outs = {n: torch.empty(pipe.out_shape(n, nseg), dtype=torch.float32, device=dev) for n in want}
for _ in range(3):
    pipe.process_device(wave, off, ln, outs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    pipe.process_device(wave, off, ln, outs)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(json.dumps({"sr": a.sr, "win": se.Params.WinSamples, "n_utt": a.n_utt, "workload": a.workload, "segments": nseg,
                  "ms": round(ms, 4), "audio_s_per_s": round(a.n_utt * a.seconds / (ms * 1e-3))}))
