"""Ray sharding across GPUs (one process per GPU, torch.distributed for the plumbing).

Rays are independent (kernel.cpp:319: one work-item per ray), so each rank traces a contiguous
slice of the direction set against its own replica of the scene.  What couples the shards:

  * the flatten length and the predelay are global: one small all-gather of the per-shard
    (min non-zero, max) times, which also carries the image-source counts;
  * image sources are de-duplicated across ALL rays with the lowest ray winning
    (rayverb.cpp:654-676): the per-shard unique lists are gathered to rank 0 and merged in rank order
    (contiguous slices => lowest rank == lowest ray), rank 0 keeps the merged list;
  * flattenImpulses is a sum, so the per-shard [channel][band][sample] histograms are combined with
    ONE reduce (NCCL over NVLink on GPUs); the band filters are linear, so rank 0 filters the sum.

`shard` is anything with render_times / raw_images / set_images / render_histogram (api.Trace on a
GPU; the CPU tests plug in an oracle-backed stand-in to check this logic under gloo).
"""
import math
import time

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


def shard_bounds(nrays_total, world, weights=None):
    """Cut points b_0 = 0 <= b_1 <= ... <= b_world = n of the contiguous ray ranges.  Equal shares by
    default; `weights` gives relative shard sizes (see balance_weights)."""
    if weights is None:
        return [(nrays_total * r) // world for r in range(world + 1)]
    w = np.maximum(np.asarray(weights, dtype=np.float64), 0.0)
    if len(w) != world or not np.isfinite(w).all() or w.sum() <= 0:
        raise ValueError("weights must be one finite non-negative number per rank, not all zero")
    cum = np.concatenate([[0.0], np.cumsum(w)]) / w.sum()
    b = [int(round(c * nrays_total)) for c in cum]
    b[0], b[-1] = 0, nrays_total
    for i in range(1, world + 1):   # monotone
        b[i] = max(b[i], b[i - 1])
    return b


def shard_slice(nrays_total, rank, world, weights=None):
    """Contiguous ray range of rank g: [g*n/N, (g+1)*n/N) by default (contiguous slices keep
    "lowest ray wins" == "lowest rank wins" for the image-source merge)."""
    b = shard_bounds(nrays_total, world, weights)
    return b[rank], b[rank + 1]


def balance_weights(ms_per_ray, extra_ms, nrays_total):
    """Shard weights that equalise the finish time of the ranks when rank r carries a fixed extra cost
    extra_ms[r] on top of a common per-ray cost (rank 0 merges the image sources and filters the summed
    histogram, so it should trace fewer rays): solve c*n_r + x_r = T with sum n_r = N."""
    x = np.asarray(extra_ms, dtype=np.float64)
    world = len(x)
    c = float(ms_per_ray)
    if c <= 0:
        return [1.0] * world
    T = (c * nrays_total + x.sum()) / world
    n = np.maximum((T - x) / c, 0.0)
    return [float(v) for v in n]


def _fold_predelay(gmn):
    """findPredelay's fold over channels (rayverb.h:47-66): zeros are ignored."""
    predelay = 0.0
    for v in gmn:
        if predelay == 0:
            predelay = v
        elif v != 0:
            predelay = min(predelay, v)
    return predelay


def exchange_extrema(mn, mx, n_images, world, device):
    """ONE small all-gather carries everything the shards need from each other before the
    histogram: per-channel (min non-zero, max) attenuated times and the image-source count of
    every rank.  Returns (predelay, max time per channel, image counts per rank)."""
    nch = len(mn)
    mine = torch.tensor([(float(x) if x != 0 else math.inf) for x in mn] + [float(x) for x in mx] + [float(n_images)],
                        device=device, dtype=torch.float64)
    allv = torch.empty((world, 2 * nch + 1), device=device, dtype=torch.float64)
    dist.all_gather_into_tensor(allv, mine.unsqueeze(0))
    allv = allv.cpu().numpy()   # float32 values and small integers are exact in float64
    gmn = [0.0 if math.isinf(v) else float(v) for v in allv[:, :nch].min(axis=0)]
    gmx = [float(v) for v in allv[:, nch:2 * nch].max(axis=0)]
    counts = [int(v) for v in allv[:, 2 * nch]]
    return _fold_predelay(gmn), gmx, counts


_REC = 64 + 4 * NUM_IMAGE_SOURCE + 4   # Impulse + key[10] + key length, bytes per image source


def merge_images(img, keys, klen, counts, rank, world, device, timing=None):
    """Gather the per-shard unique image-source lists (one tensor collective over the padded
    records; `counts` came with exchange_extrema); rank 0 returns the global list (first occurrence
    in rank order wins), the others an empty one."""
    n = len(img)
    nmax = max(counts)
    if nmax == 0:
        return np.zeros(0, dtype=IMPULSE)
    rec = np.zeros((nmax, _REC), dtype=np.uint8)
    if n:
        rec[:n, :64] = np.ascontiguousarray(img, dtype=IMPULSE).view(np.uint8).reshape(n, 64)
        rec[:n, 64:64 + 4 * NUM_IMAGE_SOURCE] = np.ascontiguousarray(keys, dtype=np.uint32).reshape(n, -1).view(np.uint8)
        rec[:n, 64 + 4 * NUM_IMAGE_SOURCE:] = np.ascontiguousarray(klen, dtype=np.int32).reshape(n, 1).view(np.uint8)
    local = torch.from_numpy(rec).to(device)
    if rank == 0:
        gathered = torch.empty((world, nmax, _REC), dtype=torch.uint8, device=device)
        dist.gather(local.unsqueeze(0), list(gathered.unsqueeze(1).unbind(0)), dst=0)
    else:
        dist.gather(local.unsqueeze(0), None, dst=0)
        return np.zeros(0, dtype=IMPULSE)
    t_excl = time.perf_counter()
    g = gathered.cpu().numpy()
    allrec = np.ascontiguousarray(np.concatenate([g[r, :counts[r]] for r in range(world)]))
    gi = np.ascontiguousarray(allrec[:, :64]).view(IMPULSE).reshape(-1)
    gk = np.ascontiguousarray(allrec[:, 64:64 + 4 * NUM_IMAGE_SOURCE]).view(np.uint32).reshape(-1, NUM_IMAGE_SOURCE)
    gl = np.ascontiguousarray(allrec[:, 64 + 4 * NUM_IMAGE_SOURCE:]).view(np.int32).reshape(-1)
    merged, _, _ = api.images_merge(gi, gk, gl)
    if timing is not None:
        timing["exclusive_ms"] = timing.get("exclusive_ms", 0.0) + 1e3 * (time.perf_counter() - t_excl)
    return merged


def sharded_histogram(shard, cfg, rank, world, device, sr, trim_predelay, timing=None):
    """Stages 2-3 of the hot path on one shard + the collectives.  Returns (hist tensor on
    `device`, summed on rank 0; lengths; predelay).  `timing` (a dict) receives host wall-clock
    milliseconds: "local_ms" for the work every rank does on its own shard, "exclusive_ms" for what
    only this rank does (rank 0: the image-source merge) -- the inputs of balance_weights."""
    t0 = time.perf_counter()
    mn, mx = shard.render_times(cfg)
    img, keys, klen = shard.raw_images(with_keys=True)
    t1 = time.perf_counter()
    predelay, gmx, counts = exchange_extrema(mn, mx, len(img), world, device)
    lengths = [flatten_length(m, predelay, trim_predelay, sr) for m in gmx]
    shard.set_images(merge_images(img, keys, klen, counts, rank, world, device, timing))
    hist = torch.empty(sum(NBANDS * l for l in lengths), device=device, dtype=torch.float32)   # render_histogram zero-fills
    t2 = time.perf_counter()
    shard.render_histogram(cfg, predelay, lengths, hist.data_ptr())
    t3 = time.perf_counter()
    dist.reduce(hist, dst=0, op=dist.ReduceOp.SUM)
    if timing is not None:
        timing["local_ms"] = timing.get("local_ms", 0.0) + 1e3 * ((t1 - t0) + (t3 - t2))
    return hist, lengths, predelay
