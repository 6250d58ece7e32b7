#!/usr/bin/env python
"""Tuning sweep over the fused kernel's launch options on one GPU (not a bench:
numbers here only rank configurations)."""
import argparse, itertools, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from auditory_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="mel")
ap.add_argument("--warps", default="0")
ap.add_argument("--segs", default="0")
ap.add_argument("--ctas", default="0")
ap.add_argument("--epi", default="0")
ap.add_argument("--combos", default="", help="warps:job_segs:ctas:epi,...")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--n-utt", type=int, default=1024)
a = ap.parse_args()
se, want = bench.build_env(a.workload, 0)
pipe = se.pipeline()
wave_h, off, ln = synth.fast_batch(a.n_utt, seed=1000, seconds=3.0)
nseg = int(pipe.seg_base(ln)[-1])
dev = torch.device("cuda", 0)
waves = [torch.from_numpy(wave_h).to(dev)]
waves.append(torch.roll(waves[0], 48000))
outs = [{n: torch.empty(pipe.out_shape(n, nseg), dtype=torch.float32, device=dev) for n in want} for _ in range(2)]
ints = lambda s: [int(x) for x in s.split(",")]
combos = list(itertools.product(ints(a.warps), ints(a.segs), ints(a.ctas), ints(a.epi)))
if a.combos:
    combos = [tuple(int(x) for x in c.split(":")) for c in a.combos.split(",")]
for w, c, g, ep in combos:
    for k, v in (("warps", w), ("job_segs", c), ("ctas", g), ("epi", ep)):
        pipe.set_option(k, v)
    rec = {"warps": w, "job_segs": c, "ctas": g, "epi": ep}
    try:
        for i in range(3):
            pipe.process_device(waves[i & 1], off, ln, outs[i & 1])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            pipe.process_device(waves[i & 1], off, ln, outs[i & 1])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        rec.update(ms=round(ms, 4), audio_s_per_s=round(a.n_utt * 3.0 / ms * 1e3))
    except Exception as ex:
        rec.update(error=str(ex)[:160])
    print(json.dumps(rec), flush=True)
