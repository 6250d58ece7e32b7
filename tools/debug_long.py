import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from auditory_b200 import synth
from test_gpu_parity import make_env, oracle_env
sig = synth.long_signal(seconds=600.0)
for gabor in (False, True):
    se = make_env(mfcc=False, gabor=gabor)
    got = se.ProcessBatch(sig, [0], [sig.size], want=["mel"] + (["gabor"] if gabor else []))
    ref = oracle_env(mfcc=False, gabor=gabor).process(sig.astype(np.float64))
    g, r = got["mel"].astype(np.float64), ref["mel"]
    bad = ~(np.abs(g - r) <= 1e-4 * np.maximum(1, np.abs(r)))
    print("gabor", gabor, "bad", bad.sum(), "of", bad.size, "nan", np.isnan(g).sum())
    segs = np.unique(np.argwhere(bad)[:, 0])
    print("bad segments", segs[:40], len(segs))
    if len(segs):
        s = segs[0]
        print("seg", s, "bad (filter, step):", np.argwhere(bad[s])[:20].tolist())
        print("got", g[s, :3, :], "\nref", r[s, :3, :])
