"""Host wall-clock breakdown of one bench step (debug aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "parallel-reverb-raytracer_b200", "python"))
import numpy as np, torch
from rayverb_b200 import api, scenes

torch.cuda.set_device(0); api.set_device(0)
api.set_stream(torch.cuda.current_stream().cuda_stream)
tri, vtx, srf = scenes.shoebox()
nrays = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dirs = torch.from_numpy(scenes.sphere_directions(nrays)).cuda()
rt = api.Raytracer(128, tri, vtx, srf)
kw = dict(speakers=[((0, 0, 1), 0.0)], sr=44100.0, ft="linkwitz_riley", lo_cutoff=45.0)
def sync(): torch.cuda.synchronize()
for it in range(6):
    sync(); t0 = time.perf_counter()
    tr = rt.raytrace((7, 2, 4), (2, 1.5, 2), None, device_ptr=dirs.data_ptr(), nrays=nrays)
    sync(); t1 = time.perf_counter()
    st = tr.stats()
    ir = tr.render_ir(**kw)
    sync(); t2 = time.perf_counter()
    rs = api.render_stats()
    print(f"iter {it}: raytrace {1e3*(t1-t0):.3f} ms (kernel {st['trace_kernel_ms']:.3f}, lib total {st['total_ms']:.3f})  render {1e3*(t2-t1):.3f} ms (lib total {rs['total_ms']:.3f}: keys {rs['keys_ms']:.3f} sort {rs['sort_ms']:.3f} reduce {rs['reduce_ms']:.3f} filter {rs['filter_ms']:.3f} finish {rs['finish_ms']:.3f})")
