"""Dump a scene + realistic ray queries (from the CPU oracle) for tools/bvh_lab/lab.cpp.
usage: python tools/bvh_lab/dump.py hall 384 40 /tmp/lab_hall"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "parallel-reverb-raytracer_b200", "python")); sys.path.insert(0, ROOT)
import numpy as np
from rayverb_b200 import scenes
from oracle import oracle as orc
name, nrays, nrefl, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
if name == "hall":
    tri, vtx, srf = scenes.concert_hall(); mic, src = (24, 1.5, 18), (20, 2.0, 4)
elif name == "cathedral":
    tri, vtx, srf = scenes.cathedral(); mic, src = (15, 1.7, 60), (15, 2.0, 10)
elif name == "room4k":
    tri, vtx = scenes.displaced_box(20.0, 8.0, 14.0, 4400, amp=0.1, fan=0.25, nsurfaces=4); srf = scenes.hall_surfaces(); mic, src = (12, 1.5, 9), (6, 2, 3)
else:
    raise SystemExit("scene?")
V = np.stack([vtx[tri["v0"].astype(np.int64), :3], vtx[tri["v1"].astype(np.int64), :3], vtx[tri["v2"].astype(np.int64), :3]], 1).astype(np.float32)
with open(out + ".tri", "wb") as f:
    np.array([len(tri)], dtype=np.uint32).tofile(f); V.reshape(-1).tofile(f)
dirs = scenes.sphere_directions(nrays, seed=7)
r = orc.raytrace(tri, vtx, srf, mic, src, dirs, nrefl)
X = r["diffuse"]["position"].reshape(nrays, nrefl, 4)[:, :, :3]
cnt = r["refl_count"]
mic = np.asarray(mic, np.float32)
q = []
for b in range(nrefl - 1):          # bounce-major, like a warp of consecutive rays at the same bounce
    for i in range(nrays):
        if cnt[i] <= b + 1:
            continue
        o = X[i, b]; d = X[i, b + 1] - o; d = d / np.linalg.norm(d)
        q.append((*o, *d, np.inf, 0.0))
        m = mic - o; mag = np.linalg.norm(m)
        q.append((*o, *(m / mag), mag, 1.0))
q = np.asarray(q, dtype=np.float32)
with open(out + ".qry", "wb") as f:
    np.array([len(q)], dtype=np.uint32).tofile(f); q.reshape(-1).tofile(f)
np.asarray(mic, np.float32).tofile(out + ".mic")
print("triangles", len(tri), "queries", len(q))
