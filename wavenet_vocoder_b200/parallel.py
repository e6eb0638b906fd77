# coding: utf-8
"""Multi-GPU host logic: utterances are independent (no cross-batch term anywhere on the path), so
a job shards by utterance, one process per GPU, with no data-path collective; the only exchange is
the gather of the finished waveforms.  (The reference loops batches on one device,
evaluate.py:162-204, padding every utterance of a batch to the longest, evaluate.py:56,215.)"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch


def shard_utterances(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Longest-first greedy partition of utterance indices over ``world`` ranks, balancing the total
    number of samples per rank (synthesis time is proportional to length).  Deterministic."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * world
    parts: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        parts[r].append(i)
        loads[r] += int(lengths[i])
    return parts


def tile_batches(indices: Sequence[int], lengths: Sequence[int], tile: int) -> List[List[int]]:
    """Group a rank's utterances (already longest first) into launches of at most ``tile`` utterances
    of similar length, so the padding to the longest member of a launch is small."""
    idx = sorted(indices, key=lambda i: (-int(lengths[i]), i))
    return [idx[k:k + tile] for k in range(0, len(idx), tile)]


def gather_waveforms(local: Dict[int, torch.Tensor], n_total: int, dst: int = 0, group=None, device=None):
    """Collect {utterance index: 1-D waveform} from every rank on ``dst`` (list indexed by utterance,
    None elsewhere).  Tensors are gathered as one padded block per rank (NCCL or gloo).  Every rank must
    use the same device type even when it holds no utterance: ``device`` defaults to the current CUDA
    device under the NCCL backend and to the CPU otherwise."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if device is None:
        if dist.get_backend(group) == "nccl":
            device = torch.device("cuda", torch.cuda.current_device())
        else:
            device = torch.device("cpu")
    dev = torch.device(device)
    local = {i: v.to(dev) for i, v in local.items()}
    meta = torch.tensor([len(local), max([v.numel() for v in local.values()], default=0)], device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    nmax = max(int(m[0]) for m in metas)
    lmax = max(int(m[1]) for m in metas)
    idx = torch.full((nmax,), -1, dtype=torch.int64, device=dev)
    lens = torch.zeros(nmax, dtype=torch.int64, device=dev)
    block = torch.zeros(nmax, max(lmax, 1), dtype=torch.float32, device=dev)
    for k, (i, v) in enumerate(sorted(local.items())):
        idx[k], lens[k] = i, v.numel()
        block[k, :v.numel()] = v.reshape(-1).float()
    outs = []
    for t in (idx, lens, block):
        bufs = [torch.zeros_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, bufs, dst=dst, group=group)
        outs.append(bufs)
    if rank != dst:
        return None
    result = [None] * n_total
    for r in range(world):
        for k in range(nmax):
            i = int(outs[0][r][k])
            if i >= 0:
                result[i] = outs[2][r][k, :int(outs[1][r][k])].clone()
    return result
