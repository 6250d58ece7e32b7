"""Utterance sharding across the GPUs of one box: utterances -- and
within one, segments -- are independent (smoothing restarts at step 0,
dft/dft.go:67; every *Segment tensor is zeroed per segment, sndenv.go:343-351),
so the path shards with no data-path collective.  Each rank owns a contiguous
block of utterances and writes a disjoint range of the output tensors; the
only communication is the host-side gather of those ranges."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np


def utterance_block(n_utt: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of rank's contiguous block; sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError("rank outside 0..world-1")
    base, extra = divmod(n_utt, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def balanced_blocks(seg_counts: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous blocks with about equal numbers of SEGMENTS (ragged batches):
    block r ends where the running total first reaches (r+1)/world of the work."""
    total = int(np.sum(seg_counts))
    cum = np.concatenate([[0], np.cumsum(seg_counts)])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cum, target, side="left"))
        # pick the boundary closest to the target
        if k > 0 and abs(cum[k - 1] - target) <= abs(cum[min(k, len(cum) - 1)] - target):
            k -= 1
        cuts.append(min(max(k, cuts[-1]), len(seg_counts)))
    cuts.append(len(seg_counts))
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def segment_range(seg_base: np.ndarray, block: Tuple[int, int]) -> Tuple[int, int]:
    """Global segment range owned by a block of utterances (seg_base: n_utt+1 prefix sums)."""
    return int(seg_base[block[0]]), int(seg_base[block[1]])


def gather_outputs(local: Dict[str, np.ndarray], seg_range: Tuple[int, int], total_segments: int,
                   group=None) -> Dict[str, np.ndarray]:
    """Host-side gather for the one-process-per-GPU layout (torchrun): every rank contributes its [seg_range) rows
    and receives the full tensors.  Rows travel as flat float32 tensors through torch.distributed's all_gather
    (gloo on host memory), padded to the largest block -- no pickling, so config 4's gigabytes of features pass.
    (Inside one process, aud_process_host_multi needs no gather at all: every GPU writes the caller's arrays.)"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.tensor([int(seg_range[0]), int(seg_range[1])], dtype=torch.int64)
    ranges = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(ranges, mine, group=group)
    ranges = [(int(r[0]), int(r[1])) for r in ranges]
    max_rows = max(e - b for b, e in ranges)
    out = {}
    for name in sorted(local):
        a = np.ascontiguousarray(local[name], dtype=np.float32)
        row = int(np.prod(a.shape[1:])) if a.ndim > 1 else 1
        if a.shape[0] != seg_range[1] - seg_range[0]:
            raise ValueError(f"'{name}' has {a.shape[0]} rows, its segment range has {seg_range[1] - seg_range[0]}")
        send = torch.zeros(max_rows * row, dtype=torch.float32)
        send[:a.size] = torch.from_numpy(a.reshape(-1))
        parts = [torch.empty(max_rows * row, dtype=torch.float32) for _ in range(world)]
        dist.all_gather(parts, send, group=group)
        full = np.zeros((total_segments,) + a.shape[1:], dtype=np.float32)
        for (b, e), part in zip(ranges, parts):
            full[b:e] = part[:(e - b) * row].numpy().reshape((e - b,) + a.shape[1:])
        out[name] = full
    return out
