"""Multi-GPU plumbing of the hot path: query-batch sharding and the final gather.

The path shards over query images (each (image, class) problem is independent); the
only collectives are the once-per-episode prototype broadcast (modeling/prototypes.py) and one
fixed-size gather of padded detections at the end, replacing the reference's pickled gloo gather
(fewx/evaluation/coco_evaluation.py:131-137, d2!/utils/comm.py:177-217).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard like InferenceSampler (d2!/data/samplers/distributed_sampler.py:191-194)."""
    shard = (n - 1) // world + 1 if n > 0 else 0
    lo = min(shard * rank, n)
    return lo, min(shard * (rank + 1), n)


def gather_detections(boxes: torch.Tensor, scores: torch.Tensor, classes: torch.Tensor, counts: torch.Tensor,
                      total_images: int):
    """Every rank passes its padded block ([b,K,4], [b,K], [b,K] i64, [b] i32); every rank gets the
    blocks of all images in global order.  One all_gather of a fixed-size fp32 buffer."""
    world = dist.get_world_size()
    shard = (total_images - 1) // world + 1
    K = boxes.shape[1]
    packed = torch.zeros((shard, K, 7), dtype=torch.float32, device=boxes.device)
    b = boxes.shape[0]
    packed[:b, :, 0:4] = boxes
    packed[:b, :, 4] = scores
    packed[:b, :, 5] = classes.to(torch.float32)       # class indices are small integers: exact in fp32
    packed[:b, 0, 6] = counts.to(torch.float32)
    out = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(out, packed)
    allp = torch.cat(out, 0)[:total_images]
    return (allp[..., 0:4].contiguous(), allp[..., 4].contiguous(), allp[..., 5].to(torch.int64),
            allp[:, 0, 6].to(torch.int32))
