"""Ray sharding across GPUs (one process per GPU, torch.distributed for the plumbing).

Rays are independent (kernel.cpp:319: one work-item per ray), so each rank traces a contiguous
slice of the direction set against its own replica of the scene.  What couples the shards:

  * the flatten length and the predelay are global, and image sources are de-duplicated across ALL
    rays with the lowest ray winning (rayverb.cpp:654-676).  ONE small all-gather carries both: every
    rank contributes its per-channel (min non-zero, max) attenuated times and its (short) list of
    unique image sources with their keys; rank 0 merges the lists in rank order (contiguous slices =>
    lowest rank == lowest ray) and keeps the result, the other ranks render no image sources;
  * flattenImpulses is a sum, so the per-shard [channel][band][sample] histograms are combined with
    ONE reduce (NCCL over NVLink on GPUs); the band filters are linear, so rank 0 filters the sum.

`shard` is anything with render_times / raw_images / set_images / render_histogram (api.Trace on a
GPU; the CPU tests plug in an oracle-backed stand-in to check this logic under gloo).

Streams: the library runs on its own non-blocking stream unless told otherwise, torch's collectives
are ordered against torch's current stream.  sharded_histogram therefore binds the library to torch's
current stream for the duration of the call, so that the histogram kernels, the reduce and whatever
the caller enqueues next (rvb_render_finish) execute in program order.
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


_REC = 64 + 4 * NUM_IMAGE_SOURCE + 4   # Impulse + key[10] + key length, bytes per image source
_CAP = 256                             # image sources per rank that ride in the first all-gather


def _pack(mn, mx, img, keys, klen, cap):
    """One rank's message: (2 nch + 1) float64 [mins (inf = none), maxs, image count] + cap records."""
    nch = len(mn)
    hdr = np.array([(float(x) if x != 0 else math.inf) for x in mn] + [float(x) for x in mx] + [float(len(img))], dtype=np.float64)
    buf = np.zeros(8 * (2 * nch + 1) + cap * _REC, dtype=np.uint8)
    buf[:hdr.nbytes] = hdr.view(np.uint8)
    n = min(len(img), cap)
    if n:
        rec = buf[hdr.nbytes:hdr.nbytes + n * _REC].reshape(n, _REC)
        rec[:, :64] = np.ascontiguousarray(img[:n], dtype=IMPULSE).view(np.uint8).reshape(n, 64)
        rec[:, 64:64 + 4 * NUM_IMAGE_SOURCE] = np.ascontiguousarray(keys[:n], dtype=np.uint32).reshape(n, -1).view(np.uint8)
        rec[:, 64 + 4 * NUM_IMAGE_SOURCE:] = np.ascontiguousarray(klen[:n], dtype=np.int32).reshape(n, 1).view(np.uint8)
    return buf


def _all_gather_bytes(buf, world, device):
    """all-gather of equal-sized uint8 host arrays -> [world, len(buf)] host array (one collective)."""
    local = torch.from_numpy(buf)
    if str(device) != "cpu":
        local = local.to(device, non_blocking=True)
    out = torch.empty((world, buf.size), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, local.unsqueeze(0))
    return out.cpu().numpy()


def exchange(mn, mx, img, keys, klen, rank, world, device, timing=None):
    """The one exchange before the histogram.  Returns (predelay, max time per channel, merged image
    list on rank 0 / empty elsewhere, merged key lengths)."""
    nch = len(mn)
    hb = 8 * (2 * nch + 1)
    g = _all_gather_bytes(_pack(mn, mx, img, keys, klen, _CAP), world, device)
    hdr = np.ascontiguousarray(g[:, :hb]).view(np.float64).reshape(world, 2 * nch + 1)   # float32 values and small integers are exact
    counts = [int(v) for v in hdr[:, 2 * nch]]
    if max(counts) > _CAP:   # long image lists (rare: many low-order reflections on a small scene): a second, larger round
        g = _all_gather_bytes(_pack(mn, mx, img, keys, klen, max(counts)), world, device)
    gmn = [0.0 if math.isinf(v) else float(v) for v in hdr[:, :nch].min(axis=0)]
    gmx = [float(v) for v in hdr[:, nch:2 * nch].max(axis=0)]
    predelay = _fold_predelay(gmn)
    if rank != 0 or sum(counts) == 0:
        return predelay, gmx, np.zeros(0, dtype=IMPULSE), np.zeros(0, dtype=np.int32)
    t_excl = time.perf_counter()
    allrec = np.ascontiguousarray(np.concatenate([g[r, hb:hb + counts[r] * _REC].reshape(counts[r], _REC) for r in range(world)]))
    gi = np.ascontiguousarray(allrec[:, :64]).view(IMPULSE).reshape(-1)
    gk = np.ascontiguousarray(allrec[:, 64:64 + 4 * NUM_IMAGE_SOURCE]).view(np.uint32).reshape(-1, NUM_IMAGE_SOURCE)
    gl = np.ascontiguousarray(allrec[:, 64 + 4 * NUM_IMAGE_SOURCE:]).view(np.int32).reshape(-1)
    merged, _, mlen = api.images_merge(gi, gk, gl)
    if timing is not None:
        timing["exclusive_ms"] = timing.get("exclusive_ms", 0.0) + 1e3 * (time.perf_counter() - t_excl)
    return predelay, gmx, merged, mlen


def sharded_histogram(shard, cfg, rank, world, device, sr, trim_predelay, timing=None, remove_direct=None):
    """Stages 2-3 of the hot path on one shard + the collectives.  Returns (hist tensor on
    `device`, summed on rank 0; lengths; predelay).  `timing` (a dict) receives host wall-clock
    milliseconds: "local_ms" for the work every rank does on its own shard, "exclusive_ms" for what
    only this rank does (rank 0: the image-source merge) -- the inputs of balance_weights.

    remove_direct (default: cfg's): drop the direct path (key {0}, rayverb.cpp:695-696).  The merged
    list handed back to the shard no longer carries keys, so the flag is applied here, on the merged
    list."""
    if remove_direct is None:
        remove_direct = bool(cfg.get("remove_direct", False) if isinstance(cfg, dict) else cfg.remove_direct)
    on_gpu = str(device) != "cpu"
    if on_gpu:   # one stream for the library's kernels and torch's collectives (see the module docstring)
        api.set_stream(torch.cuda.current_stream().cuda_stream)
    t0 = time.perf_counter()
    mn, mx = shard.render_times(cfg)
    img, keys, klen = shard.raw_images(with_keys=True)
    t1 = time.perf_counter()
    predelay, gmx, merged, mlen = exchange(mn, mx, img, keys, klen, rank, world, device, timing)
    if remove_direct and len(merged):
        merged = merged[mlen != 1]          # the only key of length 1 is {0}: the direct path
    lengths = [flatten_length(m, predelay, trim_predelay, sr) for m in gmx]
    shard.set_images(merged)
    hist = torch.empty(sum(NBANDS * l for l in lengths), device=device, dtype=torch.float32)   # render_histogram zero-fills
    t2 = time.perf_counter()
    shard.render_histogram(cfg, predelay, lengths, hist.data_ptr())
    t3 = time.perf_counter()
    dist.reduce(hist, dst=0, op=dist.ReduceOp.SUM)
    if timing is not None:
        timing["local_ms"] = timing.get("local_ms", 0.0) + 1e3 * ((t1 - t0) + (t3 - t2))
    return hist, lengths, predelay
