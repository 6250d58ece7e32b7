#!/usr/bin/env python
"""bench.py -- audio-seconds processed per second through the fused sm_100a speech-feature path
(BASELINE.json metric: "audio-sec processed/sec (mel+gabor) at 1/2/4/8 B200 vs Go CPU; roofline %").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload gabor|mel|mfcc]
  python bench.py --impl reference ...      # CPU arm: the oracle's C twin (the Go reference cannot run here)

Default workload = BASELINE configs[3], the configuration the metric is quoted on: mel + the multi-orientation
gabor FilterSet over 65,536 synthetic 3 s 16 kHz utterances, sharded by utterance across the N ranks with no
data-path collective ("scaling": "strong": the total is fixed, every rank owns 65,536 / N utterances).  A "step"
is one pass of the hot path over the rank's shard, device-resident in `value`, host buffers in and out in `e2e`.
`--workload mel` / `mfcc` run configs[1] / configs[2] (1024 utterances per rank); their kernel-only times are also
reported inside the default line under "other_workloads".
N > 1 is launched by torchrun, one rank per GPU.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

SECONDS = 3.0
SR = 16000
N_UTT_CONFIG4 = 65536          # BASELINE configs[3]: total over all ranks
N_UTT_SMALL = 1024             # BASELINE configs[1] / [2]: per rank
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # SMs x lanes x 2 x max SM clock (fallback only)


def alg_cost(workload: str):
    """Algorithmic bytes / flops per segment of S = 14 frames = 0.1 audio-s."""
    S, M, NC = 14, 32, 13
    bytes_in = 1600 * 4
    bytes_out = M * S * 4
    flops = S * (8644 + 1005 + 862)
    if workload == "mfcc":
        bytes_out += NC * S * 4
        flops += S * (832 + 201)
    if workload == "gabor":
        bytes_out += 8 * 2 * 2 * 8 * 4
        flops += 20736
    return bytes_in + bytes_out, flops


def build_env(workload: str, device: int, sr: int = SR):
    import auditory_b200 as ab
    from auditory_b200 import synth
    se = ab.SndEnv(device=device)
    se.Defaults()
    se.SampleRate = sr
    se.Signal = np.zeros(int(SECONDS * sr), dtype=np.float32)
    se.Mel.MFCC = workload == "mfcc"
    se.Mel.Deltas = False
    if workload == "gabor":
        synth.configure_processspeech_gabor(se)
    se.Init()
    if workload == "mfcc":                      # config 3: smoothing set after Init
        se.DFT.PrevSmooth, se.DFT.CurSmooth = 0.3, 0.7
    want = {"mel": ["mel"], "mfcc": ["mel", "mfcc"], "gabor": ["mel", "gabor"]}[workload]
    return se, want


def oracle_params(workload: str, rebuild_plan: int):
    from oracle import c_oracle
    p = c_oracle.default_params(mfcc=int(workload == "mfcc"), deltas=0, rebuild_plan=rebuild_plan)
    specs = []
    if workload == "mfcc":
        p.prev_smooth, p.cur_smooth = 0.3, 0.7
    if workload == "gabor":
        c_oracle.with_processspeech_gabor(p)
        specs = c_oracle.processspeech_specs()
    return p, specs


def workload_name(w: str) -> str:
    return {
        "gabor": "configs[3]: mel + multi-orientation gabor FilterSet (processspeech set: 9x9, stride 3, 8 filters) on "
                 "65,536 x 3 s 16 kHz utterances in total, sharded by utterance across the ranks",
        "mel": "configs[1]: 1024 x 3 s 16 kHz utterances per GPU, mel spectrogram only (SndEnv defaults, MFCC off)",
        "mfcc": "configs[2]: same batch as configs[1], MFCC (DCT-I, 13 coefs) + PrevSmooth 0.3 / CurSmooth 0.7",
    }[w]


class ClockSampler(threading.Thread):
    """Polls NVML SM clock / throttle reasons while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def committed_traffic(workload: str):
    """DRAM bytes per launch of the fused kernel from the committed ncu capture of this command
    (profiles/r02_traffic.json; None when that workload was not captured)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            t = json.load(f)
        t = t.get(workload)
        return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"]) if t else None
    except Exception:
        return None


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def fp32_peaks(device: int):
    """FMA-loop and butterfly-mix peaks measured on this GPU by the library's own microbenchmark (aud_measure_fp32)."""
    from auditory_b200 import _lib
    L = _lib.lib()
    out = {}
    for kind, name in ((0, "ffma"), (1, "ffma2_packed"), (2, "mix_fadd_fmul_ffma_2_1_1"), (3, "mix_packed")):
        t = C.c_double(0)
        _lib.check(L.aud_measure_fp32(device, kind, C.byref(t), None))
        out[name] = t.value
    return out


def pinned_copy_peaks(torch, dev, nbytes=256 << 20):
    """Pinned host <-> device copy bandwidth on this box (GB/s): each direction alone, then both at once."""
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def timed(fn, reps=4):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    def both():
        h2d()
        d2h()

    t_h2d, t_d2h, t_both = timed(h2d), timed(d2h), timed(both)
    return {"h2d_gbs": nbytes / t_h2d / 1e9, "d2h_gbs": nbytes / t_d2h / 1e9,
            "both_directions_gbs": 2 * nbytes / t_both / 1e9}


def time_device(torch, pipe, wave_d, off, ln, outs, steps, warmup, sync_all):
    for _ in range(warmup):
        pipe.process_device(wave_d, off, ln, outs)
    sync_all()
    l0 = pipe.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        pipe.process_device(wave_d, off, ln, outs)
    ev1.record()
    sync_all()
    return ev0.elapsed_time(ev1), pipe.launch_count - l0


def cpu_sample(workload: str, n_s: int):
    from auditory_b200 import synth
    return synth.fast_batch(n_s, seed=1000, seconds=SECONDS)


def run_reference(args, rank: int, world: int):
    """CPU arm: the oracle's C twin (the Go reference cannot be built or run here: no Go toolchain), all host
    threads, a bounded sample of the same workload per step."""
    if rank != 0:
        return
    from oracle import c_oracle
    cores = c_oracle.online_cpus()
    sample_utts = max(cores * 8, 64)
    wave, off, ln = cpu_sample(args.workload, sample_utts)
    p, specs = oracle_params(args.workload, rebuild_plan=1)
    for _ in range(args.warmup):
        c_oracle.batch_process_f32(p, specs, wave, off, ln, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_oracle.batch_process_f32(p, specs, wave, off, ln, nthreads=cores)
    dt = time.perf_counter() - t0
    val = sample_utts * SECONDS * args.steps / dt
    sample = (f"{sample_utts} x {SECONDS:g} s utterances of the workload per step (a rate metric: the per-utterance work "
              f"is identical), oracle C twin (float64, mixed-radix FFT, plan rebuilt per frame as dft/dft.go:45 does), "
              f"{cores} threads")
    line = {
        "impl": "reference", "metric": "audio-sec processed/sec", "value": val, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong" if args.workload == "gabor" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "cpu_sample_utterances_per_step": sample_utts},
        "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (default: 10 for the config-4 workload, 100 else)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gabor", choices=["gabor", "mel", "mfcc"])
    ap.add_argument("--utts", type=int, default=0, help="utterances in total (gabor) / per rank (mel, mfcc); default: the config's")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end leg (default 3 for config 4, else min(steps, 20))")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the other_workloads / peaks legs")
    ap.add_argument("--kernel-only", action="store_true", help="device-resident leg only (tuning runs; not a bench line)")
    ap.add_argument("--opt", action="append", default=[], help="name=value tuning option (aud_set_option)")
    args = ap.parse_args()
    big = args.workload == "gabor"
    if args.steps <= 0:
        args.steps = 10 if big else 100
    if args.impl == "ours":
        args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # pin this rank to the CPUs next to its GPU before any host buffer is allocated, so the pinned
    # staging memory is first-touched on the GPU's NUMA node (matters for the end-to-end leg at N > 1)
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(hnd, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    from auditory_b200 import _lib, synth
    se, want = build_env(args.workload, local_rank)
    pipe = se.pipeline()
    for kv in args.opt:
        k, v = kv.split("=")
        pipe.set_option(k, int(v))

    # ---- synthetic shard of this rank.  A base batch of 1024 utterances is generated on the host and copied to
    # the GPU; config 4's 65,536 / N utterances are built from it on the device (every copy rolled and rescaled, so
    # no two utterances are equal) -- 12.6 GB of float32 samples in total, far beyond L2.
    n_samp = int(SECONDS * SR)
    if big:
        total = args.utts or N_UTT_CONFIG4
        n_utt = total // world + (1 if rank < total % world else 0)
    else:
        total = (args.utts or N_UTT_SMALL) * world
        n_utt = args.utts or N_UTT_SMALL
    base_h, _, _ = synth.fast_batch(N_UTT_SMALL, seed=1000 + rank, seconds=SECONDS)
    base_d = torch.from_numpy(base_h).to(dev).view(N_UTT_SMALL, n_samp)
    wave_d = torch.empty((n_utt, n_samp), dtype=torch.float32, device=dev)
    for c0 in range(0, n_utt, N_UTT_SMALL):
        c = c0 // N_UTT_SMALL
        n = min(N_UTT_SMALL, n_utt - c0)
        wave_d[c0:c0 + n] = torch.roll(base_d[:n], 977 * c, dims=1) * (1.0 - 0.004 * (c % 100))
    wave_d = wave_d.view(-1)
    off = np.arange(n_utt, dtype=np.int64) * n_samp
    ln = np.full(n_utt, n_samp, dtype=np.int32)
    nseg = int(pipe.seg_base(ln)[-1])
    audio_s_per_step = n_utt * SECONDS
    outs = {n: torch.empty(pipe.out_shape(n, nseg), dtype=torch.float32, device=dev) for n in want}

    # ---- device-resident: the kernel(s) alone, CUDA events on the launching stream (torch's current stream)
    sampler = ClockSampler(local_rank)
    time_device(torch, pipe, wave_d, off, ln, outs, 0, args.warmup, sync_all)
    sampler.start()
    dev_ms, launches = time_device(torch, pipe, wave_d, off, ln, outs, args.steps, 0, sync_all)
    clocks = sampler.stop()

    if args.kernel_only:
        if rank == 0:
            print(json.dumps({"kernel_only": True, "workload": args.workload, "ms_per_step": dev_ms / args.steps,
                              "launches_per_step": launches / args.steps,
                              "value": n_utt * SECONDS * world / (dev_ms / args.steps * 1e-3), "clocks": clocks}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end: HOST buffers in and out through the C-ABI call a drop-in caller makes (aud_process_host);
    # every step copies the step's samples to the GPU and the features back, inside the timed region
    e2e_steps = args.e2e_steps or (3 if big else min(args.steps, 20))
    L = _lib.lib()
    ptrs = []

    def pinned(shape, ctype=C.c_float, nbytes_per=4):
        n = int(np.prod(shape))
        ptr = L.aud_host_alloc(max(n, 1) * nbytes_per)
        if not ptr:
            raise RuntimeError(_lib.last_error())
        ptrs.append(ptr)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,)).reshape(shape)

    wave_p = pinned((n_utt * n_samp,))
    wave_p[:] = wave_d.cpu().numpy()
    out_p = {n: pinned(pipe.out_shape(n, nseg)) for n in want}
    h2d = wave_p.nbytes
    d2h = sum(a.nbytes for a in out_p.values())

    def time_host(wave_arr, steps):
        for _ in range(2):
            pipe.process_host(wave_arr, off, ln, want=want, out=out_p)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(steps):
            pipe.process_host(wave_arr, off, ln, want=want, out=out_p)
        sync_all()
        return time.perf_counter() - t0

    e2e_s = time_host(wave_p, e2e_steps)
    assert np.isfinite(float(out_p["mel"].reshape(-1)[::9973].sum()))
    # the device-resident outputs and the host-path outputs are the same numbers
    probe = slice(0, min(nseg, 64))
    assert np.array_equal(outs["mel"][probe].cpu().numpy(), out_p["mel"][probe]), "device and host entry points disagree"

    # pageable caller memory (what a Go slice or a numpy array is): by default the library stages it through its own
    # pinned bounce buffers with a few copy threads; the two alternatives it offers are timed beside it
    pg_utts = min(n_utt, 2048)
    wave_pg = np.array(wave_p[:pg_utts * n_samp])          # ordinary malloc'd memory
    out_pg = {n: np.empty(pipe.out_shape(n, int(pipe.seg_base(ln[:pg_utts])[-1])), dtype=np.float32) for n in want}
    pageable = {}
    for mode, name in ((0, "staged (default)"), (1, "page-locked for the call"), (2, "driver staging")):
        pipe.set_option("pin", mode)
        for _ in range(2):
            pipe.process_host(wave_pg, off[:pg_utts], ln[:pg_utts], want=want, out=out_pg)
        t0 = time.perf_counter()
        pg_steps = 3
        for _ in range(pg_steps):
            pipe.process_host(wave_pg, off[:pg_utts], ln[:pg_utts], want=want, out=out_pg)
        pageable[name] = (time.perf_counter() - t0) / pg_steps
    pipe.set_option("pin", 0)
    pageable_s = pageable["staged (default)"]
    assert np.array_equal(out_pg["mel"][:64], out_p["mel"][:64]), "pageable and pinned host paths disagree"
    del wave_pg

    # ---- the same leg with 16-bit PCM input (what a WAV file holds before sound.Wave normalises it,
    # sound/sound.go:130-141): half the host-to-device bytes.  Reported beside the float32 number, not instead.
    pcm_p = pinned((n_utt * n_samp,), C.c_int16, 2)
    step_c = 1 << 24
    for a in range(0, pcm_p.size, step_c):
        pcm_p[a:a + step_c] = np.rint(wave_p[a:a + step_c] * 32767.0).astype(np.int16)
    e2e16_s = time_host(pcm_p, e2e_steps)

    # pinned-copy ceiling of the box: every rank copies at the same time (the ranks share the host's memory and
    # PCIe fabric), and the aggregate over the ranks is the denominator of the end-to-end roofline
    sync_all()
    copy_peaks = pinned_copy_peaks(torch, dev)
    if world > 1:
        t = torch.tensor([copy_peaks["h2d_gbs"], copy_peaks["d2h_gbs"], copy_peaks["both_directions_gbs"]], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        copy_peaks = {"h2d_gbs": float(t[0]), "d2h_gbs": float(t[1]), "both_directions_gbs": float(t[2])}
    for p_ in ptrs:
        L.aud_host_free(p_)
    ptrs.clear()

    # ---- the other configs' feature sets (kernel-only) and the FP32 peaks, rank 0 only and outside every timed region
    other = {}
    peaks = None
    if rank == 0 and not args.no_extra:
        peaks = fp32_peaks(local_rank)
        for w in ("mel", "mfcc", "gabor"):
            if w == args.workload and not big:
                continue
            se2, want2 = build_env(w, local_rank)
            p2 = se2.pipeline()
            n2 = N_UTT_SMALL
            off2 = np.arange(n2, dtype=np.int64) * n_samp
            ln2 = np.full(n2, n_samp, dtype=np.int32)
            nseg2 = int(p2.seg_base(ln2)[-1])
            # two rotating input copies so that consecutive steps never re-read L2-resident samples
            w_in = [base_d.reshape(-1), torch.roll(base_d.reshape(-1), n_samp)]
            o2 = {n: torch.empty(p2.out_shape(n, nseg2), dtype=torch.float32, device=dev) for n in want2}
            for i in range(5):
                p2.process_device(w_in[i & 1], off2, ln2, o2)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(50):
                p2.process_device(w_in[i & 1], off2, ln2, o2)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 50 * 1e3
            b_seg, f_seg = alg_cost(w)
            other[w] = {"workload": "1024 x 3 s utterances, " + {"mel": "mel only (configs[1])", "mfcc": "MFCC + smoothing (configs[2])",
                                                                 "gabor": "mel + gabor (configs[3] feature set)"}[w],
                        "launch_us": us, "audio_s_per_s": n2 * SECONDS / (us * 1e-6),
                        "fp32_frac": f_seg * nseg2 / (us * 1e-6) / 1e12 / peaks["ffma"],
                        "hbm_frac": b_seg * nseg2 / (us * 1e-6) / 1e9 / hbm_peak()[0]}
            p2.close()
        # the general window-length route (WinSamples != 400): 44.1 kHz, 1103-sample (prime) windows, mel + gabor --
        # frame power on the tensor cores (aud_dft_tc.cuh) and, beside it, the FP32 SIMT kernel it replaced
        sr3, n3 = 44100, 1024
        se3, want3 = build_env("gabor", local_rank, sr=sr3)
        p3 = se3.pipeline()
        ns3 = int(SECONDS * sr3)
        g3 = torch.Generator(device=dev).manual_seed(7)
        w3 = (torch.rand(n3 * ns3, generator=g3, device=dev) - 0.5) * 0.5
        off3 = np.arange(n3, dtype=np.int64) * ns3
        ln3 = np.full(n3, ns3, dtype=np.int32)
        nseg3 = int(p3.seg_base(ln3)[-1])
        o3 = {n: torch.empty(p3.out_shape(n, nseg3), dtype=torch.float32, device=dev) for n in want3}
        gen = {}
        for mode in (1, 0):
            p3.set_option("dft_tc", mode)
            for _ in range(2):
                p3.process_device(w3, off3, ln3, o3)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10 if mode else 3
            e0.record()
            for _ in range(reps):
                p3.process_device(w3, off3, ln3, o3)
            e1.record()
            torch.cuda.synchronize()
            gen[mode] = e0.elapsed_time(e1) / reps * 1e3
        other["general_44k1"] = {"workload": f"{n3} x 3 s utterances at 44.1 kHz (1103-sample windows), mel + gabor, general route",
                                 "launch_us": gen[1], "audio_s_per_s": n3 * SECONDS / (gen[1] * 1e-6),
                                 "frame_power": "tcgen05 FP16 two-slice folded DFT (aud_dft_tc.cuh)",
                                 "fp32_simt_launch_us": gen[0], "fp32_simt_audio_s_per_s": n3 * SECONDS / (gen[0] * 1e-6)}
        p3.close()

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, e2e16_s, pageable_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, e2e16_s, pageable_s = (float(x) for x in t)
        cnt = torch.tensor([n_utt, launches], dtype=torch.int64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        total_utts, total_launches = int(cnt[0]), int(cnt[1])
    else:
        total_utts, total_launches = n_utt, launches

    if rank == 0:
        ms_per_step = dev_ms / args.steps
        value = total_utts * SECONDS / (ms_per_step * 1e-3)
        e2e_val = total_utts * SECONDS * e2e_steps / e2e_s
        b_seg, f_seg = alg_cost(args.workload)
        hbm, hbm_src = hbm_peak()
        fp32_peak = peaks["ffma"] if peaks else FP32_NOMINAL_TFLOPS
        fp32_src = ("measured in this run: FFMA loop, aud_measure_fp32 kind 0" if peaks else
                    "148 SM x 128 lanes x 2 x 1.965 GHz (nominal; --no-extra skips the measurement)")
        # one step of this rank = `launches / steps` launches of the fused kernel, back to back on one stream
        launches_per_step = launches / args.steps
        segs_per_launch = nseg / launches_per_step
        launch_s = ms_per_step * 1e-3 / launches_per_step
        gbs = b_seg * segs_per_launch / launch_s / 1e9
        tfl = f_seg * segs_per_launch / launch_s / 1e12
        hbm_frac, fp32_frac = gbs / hbm, tfl / fp32_peak
        fp32_binds = fp32_frac >= hbm_frac       # the binding roof is the one with the larger fraction
        roof = {"bound": "fp32 (non-tensor FMA pipe)" if fp32_binds else "hbm",
                "achieved": tfl if fp32_binds else gbs, "peak": fp32_peak if fp32_binds else hbm,
                "unit": "TFLOP/s" if fp32_binds else "GB/s", "frac": max(hbm_frac, fp32_frac),
                "traffic": committed_traffic(args.workload),
                "peak_source": fp32_src if fp32_binds else hbm_src,
                "kernel": "fused_features_kernel", "avg_launch_us": launch_s * 1e6,
                "segments_per_launch": segs_per_launch,
                "algorithmic_flops_per_launch": f_seg * segs_per_launch,
                "algorithmic_bytes_per_launch": b_seg * segs_per_launch,
                "fp32": {"achieved": tfl, "peak": fp32_peak, "unit": "TFLOP/s", "frac": fp32_frac, "peak_source": fp32_src},
                "hbm": {"achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": hbm_frac, "peak_source": hbm_src}}
        if peaks:
            roof["fp32_peaks_measured_tflops"] = peaks
            roof["fp32"]["frac_of_butterfly_mix_peak"] = tfl / peaks["mix_packed"]
        h2d_all, d2h_all = h2d * world, d2h * world     # equal shards: every rank moves the same bytes
        e2e = {"value": e2e_val, "unit": "audio-s/s", "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all, "steps": e2e_steps,
               "api": "aud_process_host (C-ABI, host buffers in and out; caller buffers pinned with aud_host_alloc)"}
        if copy_peaks:
            # the copies of a step overlap on the two copy engines: the floor of a step is the slower direction
            # the copies of a step overlap on the two copy engines: the floor of a step is its slower direction at that
            # direction's own peak -- but never less than all the bytes at the rate the box sustains with both directions
            # busy at once (on an 8-GPU box the two directions share host memory bandwidth)
            t_dir = max(h2d_all / (copy_peaks["h2d_gbs"] * 1e9), d2h_all / (copy_peaks["d2h_gbs"] * 1e9))
            t_both = (h2d_all + d2h_all) / (copy_peaks["both_directions_gbs"] * 1e9)
            ideal = max(t_dir, t_both)
            got_gbs = (h2d_all + d2h_all) * e2e_steps / e2e_s / 1e9
            e2e["roofline"] = {"bound": "pcie (pinned host <-> device copies, all ranks copying at once)", "achieved": got_gbs,
                               "peak": (h2d_all + d2h_all) / ideal / 1e9, "unit": "GB/s",
                               "frac": got_gbs / ((h2d_all + d2h_all) / ideal / 1e9),
                               "aggregate_over_ranks": True, **copy_peaks}
        line = {
            "metric": "audio-sec processed/sec", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if big else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "utterances_total": total_utts,
                       "utterances_per_gpu": n_utt, "seconds": SECONDS, "sample_rate": SR, "segments_per_gpu": nseg,
                       "parallelism": f"utterance-shard x{world}, no data-path collective",
                       "l2": f"inputs larger than L2: {h2d / 1e9:.1f} GB of samples per GPU and step, no flush needed"},
            "roofline": roof,
            "e2e": e2e,
            "e2e_int16": {"value": total_utts * SECONDS * e2e_steps / e2e16_s, "unit": "audio-s/s",
                          "h2d_bytes_per_step": h2d // 2 * world, "d2h_bytes_per_step": d2h * world, "steps": e2e_steps,
                          "api": "aud_process_host_i16 (16-bit PCM in, normalised on the GPU; the recommended drop-in ingest)"},
            "e2e_pageable": {"value": world * pg_utts * SECONDS / pageable_s, "unit": "audio-s/s",
                             "utterances_per_step": pg_utts * world,
                             "api": "aud_process_host with ordinary (pageable) caller memory: staged through the library's pinned "
                                    "bounce buffers by copy threads",
                             "alternatives_audio_s_per_s_rank0": {k: pg_utts * SECONDS / v for k, v in pageable.items()}},
            "gpu_launches": int(total_launches),
            "clocks": clocks,
        }
        if other:
            line["other_workloads"] = other
        if world == 1 and not args.no_cpu:
            from oracle import c_oracle
            cores = c_oracle.online_cpus()
            res = {}
            n_s = max(64, cores * 8)
            wave_c, off_c, ln_c = cpu_sample(args.workload, n_s)
            for rb in (1, 0):
                p, specs = oracle_params(args.workload, rebuild_plan=rb)
                t0 = time.perf_counter()
                c_oracle.batch_process_f32(p, specs, wave_c, off_c, ln_c, nthreads=cores)
                res[rb] = n_s * SECONDS / (time.perf_counter() - t0)
            line["cpu_baseline"] = {
                "value": res[1], "unit": "audio-s/s", "cores": cores, "kind": "port",
                "sample": f"{n_s} x {SECONDS:g} s utterances of the same workload, oracle C twin (float64, FFT plan rebuilt per frame "
                          f"as dft/dft.go:45 does), one worker per host core",
                "value_plan_cached": res[0]}
            line["config"]["cpu_sample_utterances"] = n_s
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
