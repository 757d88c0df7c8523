"""world_size-2 gloo tests of the multi-GPU host logic (sharding, prototype broadcast, final gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from faster_orefsdet_b200 import dist_utils, synth
from faster_orefsdet_b200.modeling.prototypes import PrototypeBank, broadcast_bank


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _bank(C):
    return PrototypeBank([5 + i for i in range(C)], [synth.tensor((C, 7, 128), 1 + l) for l in range(3)],
                         synth.tensor((C, 128, 8, 8), 9), synth.tensor((C, 128), 10))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bank = _bank(3) if rank == 0 else None
        got = broadcast_bank(bank, torch.device("cpu"), 0)
        ref = _bank(3)
        ok = got.class_ids == ref.class_ids and torch.equal(got.support_mean, ref.support_mean) \
            and torch.equal(got.bias_cls, ref.bias_cls) and all(torch.equal(a, b) for a, b in zip(got.taps, ref.taps))
        # shard 11 images over 2 ranks like InferenceSampler, gather padded detections
        lo, hi = dist_utils.shard_range(11, rank, world)
        boxes = torch.arange(lo, hi, dtype=torch.float32).reshape(-1, 1, 1).expand(-1, 4, 4).contiguous()
        scores = boxes[..., 0].clone()
        classes = torch.zeros_like(scores, dtype=torch.int64)
        counts = torch.full((hi - lo,), 2, dtype=torch.int32)
        gb, gs, gc, gn = dist_utils.gather_detections(boxes, scores, classes, counts, 11)
        ok = ok and gb.shape == (11, 4, 4) and torch.equal(gb[:, 0, 0], torch.arange(11, dtype=torch.float32))
        ok = ok and torch.equal(gn, torch.full((11,), 2, dtype=torch.int32)) and gc.dtype == torch.int64
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_shard_range_matches_inference_sampler():
    # d2!/data/samplers/distributed_sampler.py:191-194: contiguous shards of ceil(n/world)
    for n, world in ((11, 2), (256, 8), (5, 8), (64, 1)):
        covered = []
        for r in range(world):
            lo, hi = dist_utils.shard_range(n, r, world)
            covered += list(range(lo, hi))
        assert covered == list(range(n))
