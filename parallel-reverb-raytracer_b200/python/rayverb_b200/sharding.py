"""Ray sharding across GPUs (one process per GPU, torch.distributed for the plumbing).

Rays are independent (kernel.cpp:319: one work-item per ray), so each rank traces a contiguous
slice of the direction set against its own replica of the scene.  What couples the shards:

  * the flatten length and the predelay are global: one all-reduce of (min non-zero, max) time;
  * image sources are de-duplicated across ALL rays with the lowest ray winning
    (rayverb.cpp:654-676): the per-shard unique lists are gathered and merged in rank order
    (contiguous slices => lowest rank == lowest ray), rank 0 keeps the merged list;
  * flattenImpulses is a sum, so the per-shard [channel][band][sample] histograms are combined with
    ONE reduce (NCCL over NVLink on GPUs); the band filters are linear, so rank 0 filters the sum.

`shard` is anything with render_times / raw_images / set_images / render_histogram (api.Trace on a
GPU; the CPU tests plug in an oracle-backed stand-in to check this logic under gloo).
"""
import math

import numpy as np
import torch
import torch.distributed as dist

from . import api
from .types import IMPULSE, NBANDS


def flatten_length(maxtime, predelay, fix, sr):
    """round(maxtime * sr) + 1 in float arithmetic, half away from zero (rayverb.cpp:54-57),
    after fixPredelay (rayverb.h:85-89)."""
    t = np.float32(maxtime)
    if fix:
        t = np.float32(t - np.float32(predelay)) if t > np.float32(predelay) else np.float32(0.0)
    return int(math.floor(float(np.float32(t * np.float32(sr))) + 0.5)) + 1


def shard_slice(nrays_total, rank, world):
    """Contiguous ray range [g*n/N, (g+1)*n/N) of rank g."""
    return (nrays_total * rank) // world, (nrays_total * (rank + 1)) // world


def global_times(mn, mx, device):
    """All-reduce the per-channel (min non-zero, max) attenuated times; returns (predelay, max[])."""
    t = torch.tensor([[(float(x) if x != 0 else math.inf) for x in mn], [-float(x) for x in mx]],
                     device=device, dtype=torch.float32)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    gmn = [0.0 if math.isinf(v) else v for v in t[0].tolist()]
    gmx = (-t[1]).tolist()
    predelay = 0.0
    for v in gmn:   # findPredelay's fold over channels (rayverb.h:47-66): zeros are ignored
        if predelay == 0:
            predelay = v
        elif v != 0:
            predelay = min(predelay, v)
    return predelay, gmx


def merge_images(img, keys, klen, rank, world):
    """Gather the per-shard unique image-source lists; rank 0 returns the global list (first
    occurrence in rank order wins), the others an empty one."""
    gathered = [None] * world
    dist.all_gather_object(gathered, (np.asarray(img), np.asarray(keys), np.asarray(klen)))
    if rank != 0:
        return np.zeros(0, dtype=IMPULSE)
    merged, _, _ = api.images_merge(np.concatenate([g[0] for g in gathered]), np.concatenate([g[1] for g in gathered]),
                                    np.concatenate([g[2] for g in gathered]))
    return merged


def sharded_histogram(shard, cfg, rank, world, device, sr, trim_predelay):
    """Stages 2-3 of the hot path on one shard + the collectives.  Returns (hist tensor on
    `device`, summed on rank 0; lengths; predelay)."""
    mn, mx = shard.render_times(cfg)
    predelay, gmx = global_times(mn, mx, device)
    lengths = [flatten_length(m, predelay, trim_predelay, sr) for m in gmx]
    img, keys, klen = shard.raw_images(with_keys=True)
    shard.set_images(merge_images(img, keys, klen, rank, world))
    hist = torch.zeros(sum(NBANDS * l for l in lengths), device=device, dtype=torch.float32)
    shard.render_histogram(cfg, predelay, lengths, hist.data_ptr())
    dist.reduce(hist, dst=0, op=dist.ReduceOp.SUM)
    return hist, lengths, predelay
