"""BVH traversal statistics on the C3 hall (debug aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "parallel-reverb-raytracer_b200", "python"))
import numpy as np
from rayverb_b200 import api, scenes
scene = sys.argv[1] if len(sys.argv) > 1 else "hall"
nrays = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
nrefl = int(sys.argv[3]) if len(sys.argv) > 3 else 128
if scene == "hall":
    tri, vtx, srf = scenes.concert_hall(); mic, src = (24, 1.5, 18), (20, 2.0, 4)
else:
    tri, vtx, srf = scenes.cathedral(); mic, src = (15, 1.7, 60), (15, 2.0, 10)
t0 = time.time(); rt = api.Raytracer(nrefl, tri, vtx, srf); print("scene create", time.time() - t0, "s, triangles", len(tri))
dirs = scenes.sphere_directions(nrays)
for flags in (api.TRACE_NO_IMAGES, 0):
    for _ in range(2):
        tr = rt.raytrace(mic, src, dirs, flags=flags)
    st = tr.stats()
    q = st["queries"]
    print(f"flags={flags}: kernel {st['trace_kernel_ms']:.2f} ms, queries {q/1e6:.1f} M ({q/st['trace_kernel_ms']/1e6:.2f} G/s), "
          f"tri tests/query {st['tri_tests']/q:.2f}, box tests/query {st['box_tests']/q:.1f}, bounces {st['bounces']/1e6:.1f} M of {nrays*nrefl/1e6:.1f} M")
