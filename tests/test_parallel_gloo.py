# coding: utf-8
"""N>1 host logic on CPU: 2 ranks over gloo shard a list of utterances, "synthesise" their share
(a deterministic stand-in: the kernel needs a GPU) and gather the waveforms on rank 0."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wavenet_vocoder_b200.parallel import gather_waveforms, shard_utterances, tile_batches

LENGTHS = [900, 120, 640, 640, 77, 1500, 300, 301, 50]


def fake_wave(i, n):
    return torch.sin(torch.arange(n, dtype=torch.float32) * 0.01 * (i + 1))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parts = shard_utterances(LENGTHS, world)
    local = {}
    for launch in tile_batches(parts[rank], LENGTHS, tile=4):
        assert len(launch) <= 4
        for i in launch:
            local[i] = fake_wave(i, LENGTHS[i])
    res = gather_waveforms(local, len(LENGTHS), dst=0)
    if rank == 0:
        ok = all(torch.equal(res[i], fake_wave(i, LENGTHS[i])) for i in range(len(LENGTHS)))
        q.put(ok)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_balance_and_determinism():
    parts = shard_utterances(LENGTHS, 2)
    assert sorted(parts[0] + parts[1]) == list(range(len(LENGTHS)))
    loads = [sum(LENGTHS[i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= max(LENGTHS) // 4
    assert parts == shard_utterances(LENGTHS, 2)
    assert shard_utterances(LENGTHS, 1) == [sorted(range(len(LENGTHS)), key=lambda i: (-LENGTHS[i], i))]
    parts8 = shard_utterances([100] * 64, 8)
    assert all(len(p) == 8 for p in parts8)               # BASELINE config 4: 64 utterances -> 8 per GPU


def test_two_rank_gloo_shard_and_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    for p_ in procs:
        p_.join(120)
        assert p_.exitcode == 0
    assert q.get(timeout=5) is True
