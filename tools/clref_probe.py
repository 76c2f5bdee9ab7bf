import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "parallel-reverb-raytracer_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import cl_reference as clr, oracle as orc
from rayverb_b200 import scenes
print("available:", clr.available(), clr.why_unavailable())
print("build log:", clr.build_log()[:1500])
if not clr.available(): sys.exit(0)
tri, vtx, srf = scenes.diamond_room()
d = np.zeros((64, 4), np.float32); d[:, 2] = -1
d[:6, :3] = [(0, 0, -1), (0, 0, 1), (0, -1, 0), (0, 1, 0), (-1, 0, 0), (1, 0, 0)]
r = clr.raytrace(tri, vtx, srf, (0, 2, 0), (0, 2, 2), d, 128)
D = r["diffuse"].reshape(64, 128)
for i in range(6): print(D[i, 0]["position"][:3], D[i, 1]["position"][:3])
# compare with the CPU oracle on the shoebox
tri, vtx, srf = scenes.shoebox()
dirs = scenes.sphere_directions(2000, seed=2015)
a = clr.raytrace(tri, vtx, srf, (7, 2, 4), (2, 1.5, 2), dirs, 32)
b = orc.raytrace(tri, vtx, srf, (7, 2, 4), (2, 1.5, 2), dirs, 32)
pa = a["diffuse"]["position"].reshape(2000, 32, 4); pb = b["diffuse"]["position"].reshape(2000, 32, 4)
same = (pa.view(np.uint32) == pb.view(np.uint32)).all(axis=2)
print("bit-identical positions per bounce:", same.mean(axis=0).round(3))
print("max |pos diff| per bounce:", np.abs(pa - pb).max(axis=(0, 2)).round(6))
ta, tb = a["diffuse"]["time"], b["diffuse"]["time"]; print("time max abs diff", np.abs(ta - tb).max())
va, vb = a["diffuse"]["volume"], b["diffuse"]["volume"]; print("volume max rel diff", np.abs(va - vb).max() / np.abs(vb).max())
print("image idx equal:", np.array_equal(a["image_idx"], b["image_idx"]), (a["image_idx"] != b["image_idx"]).sum())
