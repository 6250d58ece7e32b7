import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import bench
n_utt = 2048
ns = int(bench.SECONDS * bench.SR)
rng = np.random.default_rng(0)
wave = (rng.random(n_utt * ns, dtype=np.float32) - 0.5)
off = np.arange(n_utt, dtype=np.int64) * ns
ln = np.full(n_utt, ns, dtype=np.int32)
print("cores", os.cpu_count(), len(os.sched_getaffinity(0)))
for thr in (0, 2, 4, 6, 8, 12, 16):
    se, want = bench.build_env("gabor", 0)
    p = se.pipeline()
    p.set_option("copy_threads", thr)
    out = {}
    res = p.process_host(wave, off, ln, want=want)
    t0 = time.perf_counter()
    for _ in range(3):
        res = p.process_host(wave, off, ln, want=want, out=res)
    dt = (time.perf_counter() - t0) / 3
    print(f"copy_threads {thr}: {dt*1e3:.1f} ms, {n_utt * bench.SECONDS / dt:.4g} audio-s/s", flush=True)
    p.close()
