"""Batch sharding of one VkFFT configuration over the GPUs of a box (one process per GPU).

Batched transforms are independent units along the outermost (batch) stride (API guide :285-289), so a job splits
into contiguous slabs of `numberBatches` with NO data-path collective; torch.distributed is used only for the
barrier and the max-over-ranks time (NCCL on GPUs, gloo in the CPU tests).  The reference has no multi-GPU mode
(README.md:26-28 lists it as future work)."""
from dataclasses import replace
from typing import Tuple


def shard_batches(number_batches: int, world: int, rank: int) -> Tuple[int, int]:
    """contiguous slab [start, start+count) of the batch dimension owned by `rank`; slabs differ by at most 1"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(max(number_batches, 1), world)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def shard_configuration(cfg, world: int, rank: int):
    """per-rank copy of a VkFFTConfiguration: numberBatches cut to the rank's slab, device = local rank.
    Returns (cfg_rank, first_batch, byte_offset_of_slab) -- byte offset in the global (unsharded) buffer layout."""
    nb = cfg.numberBatches or 1
    start, count = shard_batches(nb, world, rank)
    per_batch = 1
    for s in cfg.size[: cfg.FFTdim]:
        per_batch *= s
    if cfg.performR2C:
        per_batch = per_batch // cfg.size[0] * (cfg.size[0] // 2 + 1)
    esz = (16 if cfg.doublePrecision else 8) if not cfg.performDCT else (8 if cfg.doublePrecision else 4)
    per_batch *= (cfg.coordinateFeatures or 1)
    out = replace(cfg, numberBatches=count, device=rank)
    return out, start, start * per_batch * esz


def max_over_ranks(value: float, dist=None) -> float:
    """max of a per-rank scalar (timings) over the job; identity when not distributed"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    import torch
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
