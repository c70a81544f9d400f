"""Batch/bucket sharding of inference and validation across the GPUs of one box.

Images are independent units, so ranks own whole bucket-pure batches and no collective touches the
data path; torch.distributed is used only to agree on timing (max over ranks) and to gather the
per-sample metrics after the run.  Bucket rule and the <=1 MP mix: reference
src/data_generation/prepare_rgba_buckets.py:33-39,75-96 and the histogram of test.ipynb:63-70
(SURVEY App. C)."""
from __future__ import annotations

import random
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

# (W, H, weight): the <= 1 048 576-pixel subset of the author's bucket histogram
BUCKET_MIX = [
    (1024, 1024, 2058520), (832, 1024, 274737), (1024, 832, 18288), (768, 512, 4582), (576, 576, 3924),
    (768, 576, 3517), (768, 768, 3182), (1024, 704, 2945), (1024, 640, 2731), (1024, 768, 2445), (704, 1024, 2064),
    (576, 896, 2033), (576, 768, 2022), (640, 640, 2022), (960, 768, 1466), (768, 1024, 1389), (576, 832, 1369),
    (1216, 768, 1333), (896, 576, 1314), (960, 640, 1311), (768, 1088, 1211), (640, 960, 1135), (1280, 704, 1021),
    (1024, 576, 1007), (704, 512, 986), (704, 704, 924), (768, 960, 815), (512, 512, 726), (1280, 768, 724),
    (512, 768, 651),
]


def sample_bucket_batches(num_batches: int, batch_size: int, seed: int = 1234) -> List[Tuple[int, int, int]]:
    """Bucket-pure batches (B, H, W) drawn in proportion to the bucket sizes (BucketBatchSampler,
    bucket_dataset.py:312-389)."""
    rng = random.Random(seed)
    weights = [w for _, _, w in BUCKET_MIX]
    picks = rng.choices(BUCKET_MIX, weights=weights, k=num_batches)
    return [(batch_size, h, w) for (w, h, _) in picks]


def batch_cost(shape: Tuple[int, int, int]) -> float:
    """Relative cost of one batch: conv work scales with pixels, mid-block attention with pixels^2
    (7.57 TFLOP per 1024^2 image of which 0.82 attention, Qwen arch)."""
    b, h, w = shape
    p = h * w / 1048576.0
    return b * (6.74 * p + 0.82 * p * p)


def assign_batches(shapes: Sequence[Tuple[int, int, int]], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first: batches sorted by descending cost go to the least-loaded rank.
    Returns, per rank, the indices into ``shapes`` it owns (deterministic, identical on every rank)."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(shapes)), key=lambda i: (-batch_cost(shapes[i]), i))
    load = [0.0] * world_size
    owned: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        owned[r].append(i)
        load[r] += batch_cost(shapes[i])
    return owned


def max_over_ranks(value: float, device: torch.device) -> float:
    """Job time = slowest rank's device time."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_per_sample(metric: torch.Tensor) -> torch.Tensor:
    """All ranks' per-sample metric vectors concatenated (ranks may own different sample counts).
    Called once after the run, outside any timed region (the reference gathers every batch)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return metric
    world = dist.get_world_size()
    n = torch.tensor([metric.numel()], dtype=torch.int64, device=metric.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    cap = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros(cap, dtype=metric.dtype, device=metric.device)
    pad[: metric.numel()] = metric.flatten()
    outs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[: int(s.item())] for o, s in zip(outs, sizes)])
