#!/usr/bin/env python
"""Regenerate tests/golden/*.npz from the float64 oracle.

The reference (Go) ships no golden vectors and cannot run here, so these
fixtures freeze the ORACLE's outputs (numpy restatement, cross-checked against
the C twin) on the BASELINE configs' synthetic inputs.  They pin the oracle
against regressions and give the GPU tests a second, file-based target; they
do not pin the oracle to the Go code (parity unpinned, see oracle/np_oracle.py).
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import np_oracle as o
from auditory_b200 import synth

out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)

# config 1: one 2 s utterance, defaults (MFCC + deltas on), processspeech gabor, 4-D out
sig = synth.config1_signal()
for tag, prev in (("cfg1", 0.0), ("cfg1_smooth", 0.3)):
    se = o.make_env(sig.astype(np.float64), mfcc=True, deltas=True, prev_smooth=prev)
    r = o.process_all(se, want_power=True)
    keep = {k: r[k] for k in ("mel", "energy", "mfcc", "deltas", "delta_deltas", "gabor")}
    keep["logpower_seg0"] = r["logpower"][0]
    keep["logpower_seg19"] = r["logpower"][19]
    np.savez_compressed(os.path.join(out, tag + ".npz"), **keep)

# tables at defaults
se = o.make_env(sig.astype(np.float64))
np.savez_compressed(os.path.join(out, "tables.npz"), bin_pts=se.Mel.BinPts, mel_filters=se.MelFilters,
                    gabor=se.GaborFilters.Filters)

# configs 2/3: first 4 utterances of the batch generator
wave, off, ln = synth.batch(4)
for tag, kw in (("cfg2_mel", dict(mfcc=False, gabor=False)),
                ("cfg3_mfcc_smooth", dict(mfcc=True, deltas=False, gabor=False, prev_smooth=0.3, cur_smooth=0.7))):
    res = {}
    for u in range(4):
        se = o.make_env(wave[off[u]:off[u] + ln[u]].astype(np.float64), **kw)
        r = o.process_all(se, want_gabor=False)
        for k in ("mel", "mfcc"):
            if k in r:
                res[f"{k}_{u}"] = r[k].astype(np.float32)   # stored narrowed: 30 segments x 4 utterances
    np.savez_compressed(os.path.join(out, tag + ".npz"), **res)
print("golden fixtures written to", out)
