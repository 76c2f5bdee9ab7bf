"""Host wall-clock timeline of the sharded (multi-GPU) step on the C2 workload: where the time
between the end of the trace and the final IR goes.  Run under torchrun, one rank per GPU."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "parallel-reverb-raytracer_b200", "python"))
import numpy as np, torch, torch.distributed as dist
from rayverb_b200 import api, scenes, sharding
from rayverb_b200.types import NBANDS

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); api.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
api.set_stream(torch.cuda.current_stream().cuda_stream)
sys.path.insert(0, ROOT)
import bench
wl_name = sys.argv[1] if len(sys.argv) > 1 else "C3_hall_100ktri_1Mx128_stereo"
wl = bench.WORKLOADS[wl_name]
tri, vtx, srf = bench.build_scene(wl["scene"])
if wl["scaling"] == "weak":
    nrays, lo = wl["rays"], rank * wl["rays"]
    all_dirs = scenes.sphere_directions(nrays * world, seed=wl["seed"])
else:
    lo, hi = sharding.shard_slice(wl["rays"], rank, world)
    nrays = hi - lo
    all_dirs = scenes.sphere_directions(wl["rays"], seed=wl["seed"])
dirs = torch.from_numpy(all_dirs[lo:lo + nrays].copy()).cuda()
rt = api.Raytracer(wl["refl"], tri, vtx, srf)
kw = bench.render_kwargs(wl)
cfg, keep = api.make_render_config(**kw)
MIC, SRC = wl["jobs"][0]
def sync(): torch.cuda.synchronize()
for it in range(6):
    dist.barrier(); sync(); T = [time.perf_counter()]
    def mark(): sync(); T.append(time.perf_counter())
    tr = rt.raytrace(MIC, SRC, None, ray_offset=lo, device_ptr=dirs.data_ptr(), nrays=nrays); mark()
    mn, mx = tr.render_times(cfg); mark()
    img, keys, klen = tr.raw_images(with_keys=True); mark()
    predelay, gmx, merged, mlen = sharding.exchange(mn, mx, img, keys, klen, rank, world, "cuda"); mark()
    lengths = [sharding.flatten_length(m, predelay, False, 44100.0) for m in gmx]
    tr.set_images(merged); mark()
    hist = torch.empty(sum(NBANDS * l for l in lengths), device="cuda", dtype=torch.float32); mark()
    tr.render_histogram(cfg, predelay, lengths, hist.data_ptr()); mark()
    dist.reduce(hist, dst=0, op=dist.ReduceOp.SUM); mark()
    if rank == 0:
        ir = api.render_finish(cfg, hist.data_ptr(), lengths)
    mark()
    names = ["raytrace", "render_times", "raw_images(D2H)", "exchange(allgather+merge)", "set_images", "hist alloc", "render_histogram", "reduce(NCCL)", "finish(rank0)"]
    d = [1e3 * (T[i + 1] - T[i]) for i in range(len(T) - 1)]
    print(f"rank {rank} it {it} total {1e3*(T[-1]-T[0]):.2f} ms: " + ", ".join(f"{n} {v:.2f}" for n, v in zip(names, d)) + f" | images {len(img)} L {lengths}", flush=True)
dist.destroy_process_group()
