import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from test_gpu_generic import envs, signal
sr = 44100
se, orc = envs(sr, mfcc=False, gabor=False)
sig = np.tile(signal(sr, 10.0, seed=1), 60)   # 600 s
pipe = se.pipeline()
ln = np.array([sig.size], dtype=np.int32); off = np.zeros(1, dtype=np.int64)
for mode in (1, 0):
    pipe.set_option("dft_tc", mode)
    got = pipe.process_host(sig, off, ln, want=("mel",))
    t0 = time.perf_counter(); got = pipe.process_host(sig, off, ln, want=("mel",)); dt = time.perf_counter() - t0
    print(f"dft_tc {mode}: 600 s utterance in {dt*1e3:.1f} ms (host path), mel {got['mel'].shape}", flush=True)
ref = orc.process(sig[: sr * 3].astype(np.float64))["mel"]
n = ref.shape[0] - 1
print("max err first segments", float(np.abs(got["mel"][:n] - np.asarray(ref)[:n]).max()))
