"""The host entry points a drop-in caller uses: pageable caller memory (page-locked by the library for the call),
the multi-GPU call (one host thread per handle, disjoint output ranges, no collective), and BASELINE config 4's
feature set at size through the device entry point."""
import ctypes as C

import numpy as np
import pytest

import auditory_b200 as ab
from auditory_b200 import _lib, synth
from test_gpu_parity import make_env, oracle_env
from util import RTOL_GABOR, RTOL_LOG, assert_close

pytestmark = pytest.mark.gpu


def second_pipeline(se, device):
    """Another handle with the same parameters (possibly on another GPU)."""
    return ab.Pipeline(se.aud_params(), se.Mel.BinPts, se.MelFilters,
                       se.GaborFilters.Filters if se._n_gabor else None, device=device)


@pytest.mark.parametrize("n_handles", [2, 3])
def test_multi_gpu_entry_point_matches_single(n_handles):
    """aud_process_host_multi with several handles -- on as many distinct GPUs as the box has (handles share a device
    when it has fewer: the sharding, threading and output ranges are what is under test)."""
    import torch
    ndev = torch.cuda.device_count()
    lens = np.array([48000, 16001, 1700, 999, 0, 33333, 2000, 48000, 1601, 24000, 48000], dtype=np.int32)
    off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    rng = np.random.default_rng(5)
    wave = rng.uniform(-0.5, 0.5, int(lens.sum())).astype(np.float32)
    se = make_env(mfcc=True, deltas=False, gabor=True, prev=0.2)
    pipes = [se.pipeline()] + [second_pipeline(se, g % ndev) for g in range(1, n_handles)]
    want = ["mel", "mfcc", "energy", "gabor"]
    single = pipes[0].process_host(wave, off, lens, want=want)
    multi = ab.process_host_multi(pipes, wave, off, lens, want=want)
    for k in want:
        assert np.array_equal(single[k], multi[k]), k
    pcm = np.round(wave * 20000).astype(np.int16)
    single16 = pipes[0].process_host(pcm, off, lens, want=["mel"])
    multi16 = ab.process_host_multi(pipes, pcm, off, lens, want=["mel"])
    assert np.array_equal(single16["mel"], multi16["mel"])
    # mismatching parameters are refused
    other = make_env(mfcc=False, gabor=False)
    with pytest.raises(ab.AudError):
        ab.process_host_multi([pipes[0], other.pipeline()], wave, off, lens, want=["mel"])
    for p in pipes[1:]:
        p.close()


def test_every_gpu_of_the_box():
    """One handle per GPU of the box (skipped on a single-GPU box; run with `gpurun --gpus 2`)."""
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs at least two GPUs")
    wave, off, ln = synth.fast_batch(512, seed=9)
    se = make_env(mfcc=False, gabor=True)
    pipes = [se.pipeline()] + [second_pipeline(se, g) for g in range(1, ndev)]
    single = pipes[0].process_host(wave, off, ln, want=["mel", "gabor"])
    multi = ab.process_host_multi(pipes, wave, off, ln, want=["mel", "gabor"])
    assert np.array_equal(single["mel"], multi["mel"]) and np.array_equal(single["gabor"], multi["gabor"])
    assert all(p.launch_count > 0 for p in pipes)


def test_pageable_and_pinned_caller_memory_agree():
    """Ordinary numpy memory (large: page-locked for the call; small: driver staging) and aud_host_alloc memory."""
    L = _lib.lib()
    wave, off, ln = synth.fast_batch(96, seed=3)          # 18 MB of samples: above the page-locking threshold
    se = make_env(mfcc=False, gabor=True)
    pipe = se.pipeline()
    pageable = pipe.process_host(wave, off, ln, want=["mel", "gabor"])
    ptr = L.aud_host_alloc(wave.nbytes)
    pinned_wave = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(wave.size,))
    pinned_wave[:] = wave
    pinned = pipe.process_host(pinned_wave, off, ln, want=["mel", "gabor"])
    L.aud_host_free(ptr)
    assert np.array_equal(pageable["mel"], pinned["mel"]) and np.array_equal(pageable["gabor"], pinned["gabor"])
    small = pipe.process_host(wave[:48000 * 2].copy(), off[:2], ln[:2], want=["mel"])
    assert np.array_equal(small["mel"], pageable["mel"][:60])
    # a read-only input array must work too (page-locking falls back to driver staging if it cannot be locked)
    ro = wave.copy()
    ro.setflags(write=False)
    again = pipe.process_host(ro, off, ln, want=["mel"])
    assert np.array_equal(again["mel"], pageable["mel"])


def test_large_pageable_batch_runs_the_staged_pipeline():
    """~100 MB of ordinary numpy memory: several utterance groups, every chunk through the pinned bounce buffers in both
    directions (and the two alternatives, option pin = 1 / 2), against page-locked caller buffers."""
    L = _lib.lib()
    wave, off, ln = synth.fast_batch(512, seed=17)
    se = make_env(mfcc=True, deltas=False, gabor=True, prev=0.3)
    pipe = se.pipeline()
    want = ["mel", "mfcc", "energy", "gabor"]
    ptr = L.aud_host_alloc(wave.nbytes)
    pinned_wave = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(wave.size,))
    pinned_wave[:] = wave
    ref = pipe.process_host(pinned_wave, off, ln, want=want)
    L.aud_host_free(ptr)
    for mode in (0, 1, 2):
        pipe.set_option("pin", mode)
        got = pipe.process_host(wave, off, ln, want=want)
        for k in want:
            assert np.array_equal(got[k], ref[k]), (mode, k)
    pipe.set_option("pin", 0)
    pcm = np.round(wave * 30000).astype(np.int16)
    a = pipe.process_host(pcm, off, ln, want=["mel"])
    pipe.set_option("pin", 2)
    b = pipe.process_host(pcm, off, ln, want=["mel"])
    pipe.set_option("pin", 0)
    assert np.array_equal(a["mel"], b["mel"])


@pytest.mark.parametrize("n_utt", [8192, 65536])
def test_config4_feature_set_at_size(n_utt):
    """BASELINE configs[3]: mel + gabor FilterSet over 8,192 and 65,536 x 3 s utterances through aud_process_device
    (many launches per call: a launch holds about 7,100 utterances).  Sampled utterances against the oracle, and bit
    for bit against the same utterances run alone."""
    import torch
    n_samp = 48000
    base_h, _, _ = synth.fast_batch(1024, seed=2000)
    dev = torch.device("cuda", 0)
    base_d = torch.from_numpy(base_h).to(dev).view(1024, n_samp)
    wave_d = torch.empty((n_utt, n_samp), dtype=torch.float32, device=dev)
    for c0 in range(0, n_utt, 1024):
        c = c0 // 1024
        wave_d[c0:c0 + 1024] = torch.roll(base_d, 977 * c, dims=1) * (1.0 - 0.004 * (c % 100))
    off = np.arange(n_utt, dtype=np.int64) * n_samp
    ln = np.full(n_utt, n_samp, dtype=np.int32)
    se = make_env(mfcc=False, gabor=True)
    pipe = se.pipeline()
    nseg = 30 * n_utt
    outs = {"mel": torch.empty((nseg, 32, 14), dtype=torch.float32, device=dev),
            "gabor": torch.empty((nseg, 256), dtype=torch.float32, device=dev)}
    n0 = pipe.launch_count
    pipe.process_device(wave_d.view(-1), off, ln, outs)
    torch.cuda.synchronize()
    assert pipe.launch_count - n0 >= n_utt // 7200
    env = oracle_env(mfcc=False, gabor=True)
    rng = np.random.default_rng(n_utt)
    picks = sorted({0, 7103, 7104, n_utt - 1, *rng.integers(0, n_utt, 6).tolist()})
    for u in picks:
        sig = wave_d[u].cpu().numpy()
        got_mel = outs["mel"][30 * u:30 * u + 30].cpu().numpy()
        got_gab = outs["gabor"][30 * u:30 * u + 30].cpu().numpy()
        ref = env.process(sig.astype(np.float64))
        assert_close(got_mel, ref["mel"], RTOL_LOG, f"config4[{n_utt}] mel utt {u}")
        assert_close(got_gab, ref["gabor"], RTOL_GABOR, f"config4[{n_utt}] gabor utt {u}")
        alone = pipe.process_host(sig, [0], [n_samp], want=["mel", "gabor"])
        assert np.array_equal(alone["mel"], got_mel) and np.array_equal(alone["gabor"], got_gab), u
    # a checksum of checksums over the whole output: every segment was written (no zeros left from the allocation)
    assert bool(torch.isfinite(outs["mel"]).all()) and float(outs["mel"].abs().sum(dim=(1, 2)).min()) > 0.0
