#!/usr/bin/env python
"""bench.py -- audio-seconds processed per second through the fused sm_100a
speech-feature path (BASELINE.json metric), with roofline, end-to-end and CPU
baseline figures.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload mel|mfcc|gabor]
  python bench.py --impl reference ...      # CPU oracle arm (the Go reference cannot run here)

A "step" is one pass of the hot path over one batch: 1024 synthetic 3 s 16 kHz
utterances per GPU (BASELINE configs[1]; --workload picks configs[2] / [3]
feature sets on the same batch).  N > 1 is launched by torchrun, one rank per
GPU; utterances are sharded by rank with no data-path collective (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

N_UTT = 1024
SECONDS = 3.0
SR = 16000
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # non-tensor FP32 peak at max SM clock


def alg_cost(workload: str):
    """Algorithmic bytes / flops per segment."""
    S, M, NC, B = 14, 32, 13, 201
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


def build_env(workload: str, device: int):
    import auditory_b200 as ab
    from auditory_b200 import synth
    se = ab.SndEnv(device=device)
    se.Defaults()
    se.SampleRate = SR
    se.Signal = np.zeros(int(SECONDS * SR), dtype=np.float32)
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


def measured_traffic(workload: str):
    """DRAM bytes per launch of the fused kernel from the committed ncu capture (None if not captured)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f)
        if t.get("workload") == workload:
            return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
    except Exception:
        pass
    return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference(args, rank: int, world: int):
    """CPU arm: the oracle's C twin (the Go reference cannot be built or run
    here), all host threads, a bounded sample of the same workload per step."""
    if rank != 0:
        return
    from auditory_b200 import synth
    from oracle import c_oracle
    cores = c_oracle.online_cpus()
    sample_utts = max(cores * 8, 64)
    wave, off, ln = synth.fast_batch(sample_utts, seed=1000, seconds=SECONDS)
    p, specs = oracle_params(args.workload, rebuild_plan=1)
    for _ in range(args.warmup):
        c_oracle.batch_process_f32(p, specs, wave, off, ln, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_oracle.batch_process_f32(p, specs, wave, off, ln, nthreads=cores)
    dt = time.perf_counter() - t0
    audio_s = sample_utts * SECONDS * args.steps
    val = audio_s / dt
    sample = (f"{sample_utts} of the {N_UTT} x {SECONDS:g} s utterances per step, oracle C twin (float64, mixed-radix "
              f"FFT, plan rebuilt per frame as dft/dft.go:45 does), {cores} threads")
    line = {
        "impl": "reference", "metric": "audio-sec processed/sec", "value": val, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "utterances_per_step": sample_utts},
        "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_name(w: str) -> str:
    return {
        "mel": "configs[1]: 1024 x 3 s 16 kHz utterances per GPU, mel spectrogram only (SndEnv defaults, MFCC off)",
        "mfcc": "configs[2]: same batch, MFCC (DCT-I, 13 coefs) + PrevSmooth 0.3 / CurSmooth 0.7",
        "gabor": "configs[3] feature set on the configs[1] batch: mel + processspeech gabor FilterSet (9x9, stride 3, 8 filters)",
    }[w]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mel", choices=["mel", "mfcc", "gabor"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end leg (default: min(steps, 20))")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--opt", action="append", default=[], help="name=value tuning option (aud_set_option)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from auditory_b200 import _lib, synth
    se, want = build_env(args.workload, local_rank)
    pipe = se.pipeline()
    for kv in args.opt:
        k, v = kv.split("=")
        pipe.set_option(k, int(v))

    # ---- synthetic batch, sharded by rank (each rank owns its own 1024 utterances)
    wave_h, off, ln = synth.fast_batch(N_UTT, seed=1000 + rank, seconds=SECONDS)
    nseg = int(pipe.seg_base(ln)[-1])
    audio_s_per_step = N_UTT * SECONDS

    # two rotating copies so consecutive steps never re-read L2-resident input (2 x 197 MB >> 126 MB L2)
    dev = torch.device("cuda", local_rank)
    waves = [torch.from_numpy(wave_h).to(dev), None]
    waves[1] = torch.roll(waves[0], int(SECONDS * SR))
    outs = [{n: torch.empty(pipe.out_shape(n, nseg), dtype=torch.float32, device=dev) for n in want} for _ in range(2)]

    def step(i):
        pipe.process_device(waves[i & 1], off, ln, outs[i & 1])

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = pipe.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    sync_all()
    dev_ms = ev0.elapsed_time(ev1)
    launches = pipe.launch_count - l0
    clocks = sampler.stop()

    # ---- end to end: pinned host buffers in and out through aud_process_host
    e2e_steps = args.e2e_steps or min(args.steps, 20)
    L = _lib.lib()
    import ctypes as C

    def pinned(shape, dtype=np.float32):
        n = int(np.prod(shape))
        ptr = L.aud_host_alloc(n * 4)
        if not ptr:
            raise RuntimeError(_lib.last_error())
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(n,)).reshape(shape)
        return arr, ptr

    wave_p, wave_ptr = pinned(wave_h.shape)
    wave_p[:] = wave_h
    out_p = {}
    ptrs = [wave_ptr]
    for n in want:
        out_p[n], p_ = pinned(pipe.out_shape(n, nseg))
        ptrs.append(p_)
    h2d = wave_p.nbytes
    d2h = sum(a.nbytes for a in out_p.values())
    for _ in range(3):
        pipe.process_host(wave_p, off, ln, want=want, out=out_p)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pipe.process_host(wave_p, off, ln, want=want, out=out_p)
    sync_all()
    e2e_s = time.perf_counter() - t0
    chk = float(out_p["mel"][::97].sum())
    assert np.isfinite(chk)

    # ---- the same leg with 16-bit PCM input (what a WAV file holds before sound.Wave normalises it):
    # half the host-to-device bytes.  Reported beside the float32 number, not instead of it.
    n16 = wave_h.size
    ptr16 = L.aud_host_alloc(n16 * 2)
    pcm_p = np.ctypeslib.as_array(C.cast(ptr16, C.POINTER(C.c_int16)), shape=(n16,))
    np.multiply(wave_h, 32767.0, out=wave_p)
    pcm_p[:] = wave_p.astype(np.int16)
    wave_p[:] = wave_h
    for _ in range(3):
        pipe.process_host(pcm_p, off, ln, want=want, out=out_p)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pipe.process_host(pcm_p, off, ln, want=want, out=out_p)
    sync_all()
    e2e16_s = time.perf_counter() - t0
    L.aud_host_free(ptr16)

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, e2e16_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, e2e16_s = float(t[0]), float(t[1]), float(t[2])

    for p_ in ptrs:
        L.aud_host_free(p_)

    if rank == 0:
        ms_per_step = dev_ms / args.steps
        value = world * audio_s_per_step / (ms_per_step * 1e-3)
        e2e_val = world * audio_s_per_step * e2e_steps / e2e_s
        b_seg, f_seg = alg_cost(args.workload)
        peak, peak_src = measured_peak()
        kern_s = ms_per_step * 1e-3          # one fused launch per step: the launch duration is the step
        gbs = b_seg * nseg / kern_s / 1e9
        tfl = f_seg * nseg / kern_s / 1e12
        line = {
            "metric": "audio-sec processed/sec", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "utterances_per_gpu": N_UTT, "seconds": SECONDS,
                       "sample_rate": SR, "segments_per_gpu": nseg, "parallelism": f"utterance-shard x{world}",
                       "l2": "inputs larger than L2: two rotating 197 MB input copies per GPU, no flush"},
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                         "traffic": measured_traffic(args.workload), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": b_seg * nseg,
                         "fp32": {"achieved": tfl, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s",
                                  "frac": tfl / FP32_PEAK_TFLOPS, "algorithmic_flops_per_launch": f_seg * nseg,
                                  "peak_source": "148 SM x 128 lanes x 2 x 1.965 GHz (non-tensor FP32)"}},
            "e2e": {"value": e2e_val, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "aud_process_host (pinned host buffers in and out)"},
            "e2e_int16": {"value": world * audio_s_per_step * e2e_steps / e2e16_s, "unit": "audio-s/s",
                          "h2d_bytes_per_step": n16 * 2, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                          "api": "aud_process_host_i16 (16-bit PCM in, normalised on the GPU)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            from oracle import c_oracle
            cores = c_oracle.online_cpus()
            res = {}
            for rb in (1, 0):
                p, specs = oracle_params(args.workload, rebuild_plan=rb)
                n_s = min(N_UTT, max(64, cores * 8))
                t0 = time.perf_counter()
                c_oracle.batch_process_f32(p, specs, wave_h, off[:n_s], ln[:n_s], nthreads=cores)
                res[rb] = n_s * SECONDS / (time.perf_counter() - t0)
            line["cpu_baseline"] = {
                "value": res[1], "unit": "audio-s/s", "cores": cores, "kind": "port",
                "sample": f"first {n_s} of the {N_UTT} utterances, oracle C twin (float64, FFT plan rebuilt per frame "
                          f"as dft/dft.go:45 does), one worker per host core",
                "value_plan_cached": res[0]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
