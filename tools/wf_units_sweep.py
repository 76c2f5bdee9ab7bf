"""Steady-phase time of the wavefront bounce kernel on the C3 hall against the number of work units a ray
is cut into (RVB_TRACE_UNITS; 0 = the launcher's choice, 1 = whole rays), for several ray counts.
Usage: tools/wf_units_sweep.py [rays,rays,...] [units,units,...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "parallel-reverb-raytracer_b200", "python"))
import numpy as np, torch
from rayverb_b200 import api, scenes

torch.cuda.set_device(0); api.set_device(0)
api.set_stream(torch.cuda.current_stream().cuda_stream)
rays_list = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [131072, 262144, 1 << 20]
units_list = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 0, 4, 8, 16]
tri, vtx, srf = scenes.concert_hall()
rt = api.Raytracer(128, tri, vtx, srf)
alld = torch.from_numpy(scenes.sphere_directions(max(rays_list), seed=2015)).cuda()
print(f"minb={os.environ.get('RVB_WF_MINB', '5')} triangles={len(tri)}")
for n in rays_list:
    for u in units_list:
        best = (1e9, 0, 0)
        for it in range(4):
            tr = rt.raytrace((24, 1.5, 18), (20, 2.0, 4), None, device_ptr=alld.data_ptr(), nrays=n, flags=api.TRACE_UNITS(u))
            st = tr.stats()
            best = min(best, (st["steady_phase_ms"], st["image_phase_ms"], st["total_ms"]))
            tr.close()
        print(f"rays {n:8d} units {u:2d}: steady {best[0]:7.3f} ms  image {best[1]:6.3f}  total {best[2]:7.3f}", flush=True)
