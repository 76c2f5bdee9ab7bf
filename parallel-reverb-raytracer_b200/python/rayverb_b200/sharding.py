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
from .types import IMPULSE, NBANDS, NUM_IMAGE_SOURCE


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


_REC = 64 + 4 * NUM_IMAGE_SOURCE + 4   # Impulse + key[10] + key length, bytes per image source


def merge_images(img, keys, klen, rank, world, device):
    """Gather the per-shard unique image-source lists (two tensor collectives: counts, then the
    padded records); rank 0 returns the global list (first occurrence in rank order wins), the
    others an empty one."""
    n = len(img)
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    mine = torch.tensor([n], dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts, mine)
    counts = counts.tolist()
    nmax = max(counts)
    if nmax == 0:
        return np.zeros(0, dtype=IMPULSE)
    rec = np.zeros((nmax, _REC), dtype=np.uint8)
    if n:
        rec[:n, :64] = np.ascontiguousarray(img, dtype=IMPULSE).view(np.uint8).reshape(n, 64)
        rec[:n, 64:64 + 4 * NUM_IMAGE_SOURCE] = np.ascontiguousarray(keys, dtype=np.uint32).reshape(n, -1).view(np.uint8)
        rec[:n, 64 + 4 * NUM_IMAGE_SOURCE:] = np.ascontiguousarray(klen, dtype=np.int32).reshape(n, 1).view(np.uint8)
    local = torch.from_numpy(rec).to(device)
    gathered = torch.empty((world, nmax, _REC), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(gathered, local.unsqueeze(0))
    if rank != 0:
        return np.zeros(0, dtype=IMPULSE)
    g = gathered.cpu().numpy()
    parts = [g[r, :counts[r]] for r in range(world)]
    allrec = np.ascontiguousarray(np.concatenate(parts))
    gi = np.ascontiguousarray(allrec[:, :64]).view(IMPULSE).reshape(-1)
    gk = np.ascontiguousarray(allrec[:, 64:64 + 4 * NUM_IMAGE_SOURCE]).view(np.uint32).reshape(-1, NUM_IMAGE_SOURCE)
    gl = np.ascontiguousarray(allrec[:, 64 + 4 * NUM_IMAGE_SOURCE:]).view(np.int32).reshape(-1)
    merged, _, _ = api.images_merge(gi, gk, gl)
    return merged


def sharded_histogram(shard, cfg, rank, world, device, sr, trim_predelay):
    """Stages 2-3 of the hot path on one shard + the collectives.  Returns (hist tensor on
    `device`, summed on rank 0; lengths; predelay)."""
    mn, mx = shard.render_times(cfg)
    predelay, gmx = global_times(mn, mx, device)
    lengths = [flatten_length(m, predelay, trim_predelay, sr) for m in gmx]
    img, keys, klen = shard.raw_images(with_keys=True)
    shard.set_images(merge_images(img, keys, klen, rank, world, device))
    hist = torch.zeros(sum(NBANDS * l for l in lengths), device=device, dtype=torch.float32)
    shard.render_histogram(cfg, predelay, lengths, hist.data_ptr())
    dist.reduce(hist, dst=0, op=dist.ReduceOp.SUM)
    return hist, lengths, predelay
