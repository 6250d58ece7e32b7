"""General-route frame power: the tensor-core kernel against the FP32 SIMT kernel and the float64 oracle --
worst errors of power (relative to the frame peak) / logpower / mel, and the device time of a large batch.
Run on the GPU box: python tools/dft_tc_check.py [sample_rate]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch  # noqa: E402

from test_gpu_generic import envs, signal  # noqa: E402

sr = int(sys.argv[1]) if len(sys.argv) > 1 else 44100
se, orc = envs(sr)
sig = signal(sr, 1.3, seed=sr)
ref = orc.process(sig.astype(np.float64), want_power=True)
want = ["mel", "power", "logpower", "gabor", "mfcc"]
for mode in (0, 1):
    se.pipeline().set_option("dft_tc", mode)
    got = se.ProcessBatch(sig, [0], [sig.size], want=want)
    pw, rp = got["power"].astype(np.float64), np.asarray(ref["power"]).reshape(got["power"].shape)
    peak = rp.max(axis=1, keepdims=True)
    line = {"power/peak": float(np.abs(pw - rp).max() / peak.max()),
            "power_rel_to_frame_peak": float((np.abs(pw - rp) / peak).max())}
    for k in ("logpower", "mel", "mfcc", "gabor"):
        r = np.asarray(ref[k]).reshape(got[k].shape)
        line[k] = float(np.abs(got[k] - r).max())
    print("dft_tc", mode, {k: f"{v:.3g}" for k, v in line.items()}, flush=True)

# timing: 2048 utterances of 1.3 s, device resident
n_utt = int(os.environ.get("N_UTT", "2048"))
wave = np.tile(sig, n_utt)
off = np.arange(n_utt, dtype=np.int64) * sig.size
ln = np.full(n_utt, sig.size, dtype=np.int32)
pipe = se.pipeline()
dw = torch.from_numpy(wave).cuda()
nseg = int(pipe.seg_base(ln)[-1])
outs = {"mel": torch.empty(pipe.out_shape("mel", nseg), device="cuda")}
for mode in (0, 1):
    pipe.set_option("dft_tc", mode)
    for _ in range(2):
        pipe.process_device(dw, off, ln, outs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        pipe.process_device(dw, off, ln, outs)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"dft_tc {mode}: {dt * 1e3:.2f} ms per batch, {n_utt * sig.size / sr / dt:.4g} audio-s/s", flush=True)
