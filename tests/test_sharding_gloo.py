"""world_size-2 gloo test (CPU) of the ray-sharding logic in rayverb_b200/sharding.py: global
predelay/lengths, cross-shard image-source merge (lowest ray wins), one reduce of the per-shard
histograms.  The per-shard compute is the CPU oracle standing in for the GPU trace; the result must
equal the unsharded oracle pipeline."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rayverb_b200 import scenes, sharding

SR = 44100.0
SPEAKERS = [((-1, 0, -1), 0.5), ((1, 0, -1), 0.5)]
MIC, SRC, NREFL, NRAYS = (7, 2, 4), (2, 1.5, 2), 12, 1200


class OracleShard:
    """api.Trace look-alike backed by the CPU oracle."""

    def __init__(self, orc, tri, vtx, srf, dirs):
        self.orc = orc
        r = orc.raytrace(tri, vtx, srf, MIC, SRC, dirs, NREFL, debug=False)
        self.diffuse = r["diffuse"]
        self.images, self.keys, self.klen = orc.image_dedupe(r["image"], r["image_idx"])

    def _attenuated(self, images=None):
        imp = np.concatenate([self.diffuse, self.images if images is None else images])
        return [self.orc.attenuate_speaker(imp, MIC, d, c) for d, c in SPEAKERS]

    def render_times(self, cfg):
        # like rvb_render_times: cfg.remove_direct hides the shard's own direct path (key {0})
        images = self.images[self.klen != 1] if cfg.get("remove_direct") and len(self.images) == len(self.klen) else None
        att = self._attenuated(images)
        mn = [float(a["time"][a["time"] != 0].min()) if np.any(a["time"] != 0) else 0.0 for a in att]
        mx = [float(a["time"].max()) for a in att]
        return mn, mx

    def raw_images(self, with_keys=True):
        return self.images, self.keys.astype(np.uint32), self.klen

    def set_images(self, images):
        self.images = np.ascontiguousarray(images, dtype=self.orc.IMPULSE)

    def render_histogram(self, cfg, predelay, lengths, ptr):
        att = self._attenuated()
        total = sum(8 * l for l in lengths)
        out = np.ctypeslib.as_array((ctypes.c_float * total).from_address(ptr))
        o = 0
        for a, l in zip(att, lengths):
            if cfg["trim_predelay"]:
                self.orc.lib().orc_fix_predelay(a.ctypes.data_as(ctypes.c_void_p), len(a), ctypes.c_float(predelay))
            f = self.orc.flatten(a, SR)
            assert f.shape[1] <= l
            h = np.zeros((8, l), np.float32)
            h[:, :f.shape[1]] = f
            out[o:o + 8 * l] = h.ravel()
            o += 8 * l


def _worker(rank, world, port, trim_predelay, q, remove_direct=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    tri, vtx, srf = scenes.shoebox()
    dirs = scenes.sphere_directions(NRAYS, seed=77)
    b, e = sharding.shard_slice(NRAYS, rank, world)
    shard = OracleShard(orc, tri, vtx, srf, dirs[b:e])
    hist, lengths, predelay = sharding.sharded_histogram(shard, {"trim_predelay": trim_predelay, "remove_direct": remove_direct},
                                                         rank, world, "cpu", SR, trim_predelay)
    if rank == 0:
        q.put((hist.numpy().copy(), lengths, predelay, len(shard.images)))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(trim_predelay, orc, remove_direct=False):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, trim_predelay, q, remove_direct)) for r in range(2)]
    for p in procs:
        p.start()
    hist, lengths, predelay, nimg = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # unsharded oracle
    tri, vtx, srf = scenes.shoebox()
    dirs = scenes.sphere_directions(NRAYS, seed=77)
    r = orc.raytrace(tri, vtx, srf, MIC, SRC, dirs, NREFL, debug=False)
    images, _, _ = orc.image_dedupe(r["image"], r["image_idx"], remove_direct)
    assert nimg == len(images)                                   # merge == global dedupe (minus the direct path if asked)
    imp = np.concatenate([r["diffuse"], images])
    att = [orc.attenuate_speaker(imp, MIC, d, c) for d, c in SPEAKERS]
    if trim_predelay:
        pd = orc.fix_predelay(att)
        assert np.float32(pd) == np.float32(predelay)
    flat = [orc.flatten(a, SR) for a in att]
    assert lengths == [f.shape[1] for f in flat]
    o = 0
    for f in flat:
        h = hist[o:o + f.size].reshape(f.shape)
        o += f.size
        assert np.linalg.norm(h - f) <= 1e-6 * np.linalg.norm(f)  # same bins; float sum order differs


def test_two_shards_equal_unsharded(orc):
    _run(False, orc)


def test_two_shards_with_predelay(orc):
    _run(True, orc)


@pytest.mark.parametrize("trim_predelay", [False, True])
def test_two_shards_remove_direct(orc, trim_predelay):
    """remove_direct on the sharded path (rayverb.cpp:695-696): only rank 0's list carries key {0}, and it
    must not reach the histogram; the predelay then comes from the remaining impulses."""
    _run(trim_predelay, orc, remove_direct=True)


def test_shard_slices_cover_all_rays():
    for n in (0, 1, 7, 1 << 20):
        for w in (1, 2, 3, 8):
            cuts = [sharding.shard_slice(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))


def test_weighted_shards_cover_all_rays_and_balance():
    """Uneven contiguous shards (rank 0 also merges + filters, so it traces fewer rays): the cut
    points still tile [0, n) and the balance model equalises c * n_r + x_r."""
    for n in (1, 1000, 1 << 20):
        for w in ([1, 1], [0.5, 1, 1, 1], [0, 3, 1], [1e-9, 1]):
            b = sharding.shard_bounds(n, len(w), w)
            assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(len(w)))
    c, extra, n = 1.25e-4, [1.6, 0, 0, 0, 0, 0, 0, 0], 1 << 20
    wts = sharding.balance_weights(c, extra, n)
    b = sharding.shard_bounds(n, 8, wts)
    finish = [c * (b[r + 1] - b[r]) + extra[r] for r in range(8)]
    assert max(finish) - min(finish) < 2 * c                      # within a ray or two of each other
    assert b[1] - b[0] < n // 8 < b[2] - b[1]                      # rank 0 got the smaller shard
    assert sharding.balance_weights(0.0, extra, n) == [1.0] * 8   # degenerate measurement: equal shares
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, [0, 0])


def _uneven_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    tri, vtx, srf = scenes.shoebox()
    dirs = scenes.sphere_directions(NRAYS, seed=77)
    b, e = sharding.shard_slice(NRAYS, rank, world, [0.3, 1.0])
    shard = OracleShard(orc, tri, vtx, srf, dirs[b:e])
    timing = {}
    hist, lengths, predelay = sharding.sharded_histogram(shard, {"trim_predelay": False}, rank, world, "cpu", SR, False, timing)
    assert timing["local_ms"] > 0 and (rank != 0 or timing["exclusive_ms"] > 0)
    if rank == 0:
        q.put((hist.numpy().copy(), lengths, len(shard.images)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_uneven_shards_equal_even_shards(orc):
    """The sharded result does not depend on WHERE the contiguous cut is."""
    ctx = mp.get_context("spawn")
    outs = []
    for target, args in ((_uneven_worker, ()), (_worker, (False,))):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=target, args=(r, 2, port) + args + (q,)) for r in range(2)]
        for p in procs:
            p.start()
        outs.append(q.get(timeout=120))
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    (h_a, len_a, n_a), (h_b, len_b, _, n_b) = outs[0], outs[1]
    assert len_a == len_b and n_a == n_b
    assert np.linalg.norm(h_a - h_b) <= 1e-6 * np.linalg.norm(h_b)


def test_flatten_length_rounds_half_away_from_zero():
    assert sharding.flatten_length(0.0, 0.0, False, 44100.0) == 1
    assert sharding.flatten_length(2.5 / 1000.0, 0.0, False, 1000.0) == 4      # round(2.5) = 3, +1
    assert sharding.flatten_length(1.0, 0.25, True, 1000.0) == 751
    assert sharding.flatten_length(0.1, 0.25, True, 1000.0) == 1
