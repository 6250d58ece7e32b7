"""Multi-rank sharding logic on CPU: world_size-2 gloo processes (127.0.0.1).  The compute
stand-in is the oracle (there is no CPU product path); what is under test is the partition,
the segment ranges and the host-side gather."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from auditory_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_blocks_cover_everything_once():
    for n, w in ((1024, 8), (65536, 8), (10, 4), (3, 8), (0, 2)):
        blocks = [shard.utterance_block(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [e - b for b, e in blocks]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.utterance_block(4, 4, 4)


def test_balanced_blocks_follow_segment_counts():
    segs = [30] * 10 + [1] * 100 + [0] * 5 + [300]
    blocks = shard.balanced_blocks(segs, 4)
    assert blocks[0][0] == 0 and blocks[-1][1] == len(segs)
    assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
    work = [sum(segs[b:e]) for b, e in blocks]
    assert sum(work) == sum(segs) and max(work) <= 300 + 30


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import c_oracle
    from auditory_b200 import shard as sh, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_utt = 5
    lens = np.array([16000, 9000, 24000, 1700, 16000], dtype=np.int32)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    wave = np.concatenate([synth.batch_utterance(u, seconds=lens[u] / 16000.0) for u in range(n_utt)])
    env = c_oracle.Env(c_oracle.default_params(mfcc=0, deltas=0))
    segs = np.array([max(env.seg_count(int(x)), 0) for x in lens])
    seg_base = np.concatenate([[0], np.cumsum(segs)])
    block = sh.balanced_blocks(segs, world)[rank]
    rng = sh.segment_range(seg_base, block)
    local = np.concatenate([env.process(wave[offs[u]:offs[u] + lens[u]].astype(np.float64))["mel"]
                            for u in range(*block)] or [np.zeros((0, 32, 14))])
    full = sh.gather_outputs({"mel": local.astype(np.float32)}, rng, int(seg_base[-1]))
    if rank == 0:
        ref = np.concatenate([env.process(wave[offs[u]:offs[u] + lens[u]].astype(np.float64))["mel"]
                              for u in range(n_utt)]).astype(np.float32)
        np.save(os.path.join(tmp, "ok.npy"), np.array([np.array_equal(full["mel"], ref), full["mel"].shape[0]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_equals_single(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(tmp_path / "ok.npy")
    assert ok[0] == 1 and ok[1] == 10 + 5 + 15 + 1 + 10
