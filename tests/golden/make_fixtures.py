"""Generates tests/golden/vault_scene.npz (the C1 scene) and tests/golden/c1_oracle.npz (oracle outputs on it).

Run in the build container only: it reads the reference's shipped assets
(/root/reference/demo/assets/test_models/vault.obj + materials/vault.json, the room named by
SURVEY.md 8(d) for config C1) through this repo's own OBJ loader and stores the resulting
Triangle[] / vertices / Surface[] arrays, because /root/reference does not exist on the GPU box.
The oracle outputs (hit triangles, reflection counts, IR) are stored next to them as golden
vectors for the C1 configuration: 1024 rays x 64 reflections, single speaker.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "parallel-reverb-raytracer_b200", "python"))

from oracle import oracle as orc  # noqa: E402
from rayverb_b200 import scenes  # noqa: E402

REF = "/root/reference/demo/assets"
C1 = dict(mic=(0, 1.75, 6), src=(0, 1.75, 0), nrays=1024, nrefl=64, seed=2015, sr=44100.0, lo_cutoff=60.0,
          speakers=[((0, 0, -1), 0.0)])


def main():
    tri, vtx, srf = scenes.load_obj(os.path.join(REF, "test_models", "vault.obj"), os.path.join(REF, "materials", "vault.json"))
    assert len(tri) == 3754
    np.savez_compressed(os.path.join(HERE, "vault_scene.npz"), triangles=tri, vertices=vtx, surfaces=srf)
    dirs = scenes.sphere_directions(C1["nrays"], seed=C1["seed"])
    r = orc.raytrace(tri, vtx, srf, C1["mic"], C1["src"], dirs, C1["nrefl"])
    img, keys, klen = orc.image_dedupe(r["image"], r["image_idx"])
    ir = orc.render_ir(tri, vtx, srf, C1["mic"], C1["src"], dirs, C1["nrefl"], speakers=C1["speakers"], sr=C1["sr"],
                       ft="linkwitz_riley", lo_cutoff=C1["lo_cutoff"], trim_predelay=True)
    np.savez_compressed(os.path.join(HERE, "c1_oracle.npz"), hit_tri=r["hit_tri"], refl_count=r["refl_count"],
                        ntests=np.int64(r["ntests"]), image_keys=keys, image_klen=klen,
                        image_time=img["time"], ir=ir[0],
                        first_bounce_position=r["diffuse"]["position"].reshape(C1["nrays"], C1["nrefl"], 4)[:, 0])
    print("triangles", len(tri), "images", len(img), "ir samples", len(ir[0]), "escaped rays", int((r["refl_count"] < C1["nrefl"]).sum()))


if __name__ == "__main__":
    main()
