"""GPU tests of the C++ host mirror: the reference's gtest suite re-expressed against the C++
classes (bin/host_tests) and the parallel_raytrace CLI end to end against the CPU oracle."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

from rayverb_b200 import api, scenes

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "parallel-reverb-raytracer_b200", "bin")
GOLD = os.path.join(ROOT, "tests", "golden")


def test_reference_gtests_against_cpp_classes():
    r = subprocess.run([os.path.join(BIN, "host_tests"), os.path.join(GOLD, "large_square.obj"), os.path.join(GOLD, "materials.json")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PASSED" in r.stdout


def read_pcm(path):
    raw = open(path, "rb").read()
    if raw[:4] == b"RIFF":
        nch, sr, bits = struct.unpack_from("<H", raw, 22)[0], struct.unpack_from("<I", raw, 24)[0], struct.unpack_from("<H", raw, 34)[0]
        n = struct.unpack_from("<I", raw, 40)[0]
        data = raw[44:44 + n]
        little = True
    else:
        assert raw[:4] == b"FORM" and raw[8:12] == b"AIFF" and raw[12:16] == b"COMM"
        nch, frames, bits = struct.unpack_from(">hIh", raw, 20)
        expo, mant = struct.unpack_from(">HQ", raw, 28)
        sr = int(round(mant * 2.0 ** (expo - 16383 - 63)))
        assert raw[38:42] == b"SSND"
        n = struct.unpack_from(">I", raw, 42)[0] - 8
        data = raw[54:54 + n]
        little = False
    b = bits // 8
    a = np.frombuffer(data, dtype=np.uint8).reshape(-1, b).astype(np.int64)
    if not little:
        a = a[:, ::-1]
    v = sum(a[:, k] << (8 * k) for k in range(b))
    v = np.where(v >= 1 << (bits - 1), v - (1 << bits), v)
    return v.reshape(-1, nch).T / float((1 << (bits - 1)) - 1), sr, bits


@pytest.mark.parametrize("ext,bits,model", [("wav", 16, "speakers"), ("aiff", 24, "hrtf")])
def test_cli_end_to_end_vs_oracle(orc, tmp_path, ext, bits, model):
    cfg = {"rays": 3000, "reflections": 24, "sample_rate": 44100, "bit_depth": bits, "source_position": [1.5, 1.5, 1.0],
           "mic_position": [4.5, 1.2, 1.2], "filter": "linkwitz_riley", "hipass": 60, "trim_predelay": True,
           "output_mode": "all", "normalize": True, "trim_tail": True}
    if model == "speakers":
        cfg["attenuation_model"] = {"speakers": [{"direction": [-1, 0, -1], "shape": 0.5}, {"direction": [1, 0, -1], "shape": 0.5}]}
    else:
        cfg["attenuation_model"] = {"hrtf": {"facing": [0, 0, -2], "up": [0, 3, 0]}}
    (tmp_path / "cfg.json").write_text(json.dumps(cfg))
    out = tmp_path / f"ir.{ext}"
    env = dict(os.environ, RAYVERB_SEED="1234", RAYVERB_DUMP_DIRECTIONS=str(tmp_path / "dirs.bin"), RAYVERB_DUMP_IR=str(tmp_path / "ir.bin"))
    r = subprocess.run([os.path.join(BIN, "parallel_raytrace"), str(tmp_path / "cfg.json"), os.path.join(GOLD, "l_room.obj"),
                        os.path.join(GOLD, "materials.json"), str(out)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    dirs = np.fromfile(tmp_path / "dirs.bin", dtype=np.float32).reshape(-1, 4)
    assert len(dirs) == 3000 and np.allclose(np.linalg.norm(dirs[:, :3], axis=1), 1, atol=1e-6)
    raw = open(tmp_path / "ir.bin", "rb").read()
    nch, n = struct.unpack_from("<QQ", raw, 0)
    ir = np.frombuffer(raw, dtype=np.float32, offset=16).reshape(nch, n)
    # the oracle on the same scene, directions and config
    tri, vtx, srf = scenes.load_obj(os.path.join(GOLD, "l_room.obj"), os.path.join(GOLD, "materials.json"))
    kw = dict(sr=44100.0, ft="linkwitz_riley", lo_cutoff=60.0, trim_predelay=True)
    if model == "speakers":
        kw["speakers"] = [((-1, 0, -1), 0.5), ((1, 0, -1), 0.5)]
    else:   # the CLI uses the synthetic built-in table (same formula as tests/hrtf_table.smooth_table)
        from hrtf_table import smooth_table
        kw["hrtf"] = dict(table=smooth_table(), facing=(0, 0, -1), up=(0, 1, 0))    # normalised at parse time
    want = orc.render_ir(tri, vtx, srf, cfg["mic_position"], cfg["source_position"], dirs, cfg["reflections"], **kw)
    assert nch == len(want) == 2
    for c in range(2):
        assert n == len(want[c])
        assert np.linalg.norm(ir[c] - want[c]) <= 1e-4 * np.linalg.norm(want[c])
    pcm, sr, got_bits = read_pcm(out)
    assert sr == 44100 and got_bits == bits and pcm.shape == (2, n)
    assert np.abs(pcm - ir).max() <= 1.0 / ((1 << (bits - 1)) - 1)


def test_cli_loads_hrtf_table_file(orc, tmp_path):
    """RAYVERB_HRTF_TABLE / HrtfAttenuator::loadTable (host/rayverb.cpp: readTable): the CLI renders with a
    table produced by the HRTF pipeline (rayverb_b200/hrtf_table.py) instead of the synthetic built-in,
    and a file of the wrong size is an error (exit code 1, like every std::runtime_error in
    cmd/main.cpp:299-316)."""
    from rayverb_b200 import hrtf_table as ht
    rng = np.random.default_rng(9)
    grid = [({"r": 1, "a": a, "e": e}, (0.2 + rng.random((2, 8))).tolist()) for a in range(0, 361, 30) for e in range(0, 181, 30)]
    table = ht.interpolate(grid)
    ht.write_table(str(tmp_path / "table.f32"), table)
    cfg = {"rays": 2000, "reflections": 16, "sample_rate": 44100, "bit_depth": 16, "source_position": [1.5, 1.5, 1.0],
           "mic_position": [4.5, 1.2, 1.2], "filter": "linkwitz_riley", "hipass": 60, "trim_predelay": False,
           "output_mode": "all", "normalize": True, "trim_tail": True,
           "attenuation_model": {"hrtf": {"facing": [0.3, 0, -1], "up": [0, 1, 0]}}}
    (tmp_path / "cfg.json").write_text(json.dumps(cfg))
    args = [os.path.join(BIN, "parallel_raytrace"), str(tmp_path / "cfg.json"), os.path.join(GOLD, "l_room.obj"),
            os.path.join(GOLD, "materials.json"), str(tmp_path / "ir.wav")]
    env = dict(os.environ, RAYVERB_SEED="77", RAYVERB_DUMP_DIRECTIONS=str(tmp_path / "dirs.bin"), RAYVERB_DUMP_IR=str(tmp_path / "ir.bin"),
               RAYVERB_HRTF_TABLE=str(tmp_path / "table.f32"))
    r = subprocess.run(args, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    dirs = np.fromfile(tmp_path / "dirs.bin", dtype=np.float32).reshape(-1, 4)
    raw = open(tmp_path / "ir.bin", "rb").read()
    nch, n = struct.unpack_from("<QQ", raw, 0)
    ir = np.frombuffer(raw, dtype=np.float32, offset=16).reshape(nch, n)
    tri, vtx, srf = scenes.load_obj(os.path.join(GOLD, "l_room.obj"), os.path.join(GOLD, "materials.json"))
    f = np.array([0.3, 0, -1]); f = f / np.linalg.norm(f)
    want = orc.render_ir(tri, vtx, srf, cfg["mic_position"], cfg["source_position"], dirs, cfg["reflections"], sr=44100.0,
                         ft="linkwitz_riley", lo_cutoff=60.0, hrtf=dict(table=table, facing=tuple(f), up=(0, 1, 0)))
    assert nch == 2
    for c in range(2):
        assert n == len(want[c])
        assert np.linalg.norm(ir[c] - want[c]) <= 1e-4 * np.linalg.norm(want[c])
    # the built-in table gives a different response: the file really was used
    r2 = subprocess.run(args, capture_output=True, text=True, env={k: v for k, v in env.items() if k != "RAYVERB_HRTF_TABLE"})
    assert r2.returncode == 0
    raw2 = open(tmp_path / "ir.bin", "rb").read()
    ir2 = np.frombuffer(raw2, dtype=np.float32, offset=16).reshape(struct.unpack_from("<QQ", raw2, 0))
    assert ir2.shape != ir.shape or np.linalg.norm(ir2 - ir) > 1e-2 * np.linalg.norm(ir)
    (tmp_path / "short.f32").write_bytes(b"\0" * 1000)
    r3 = subprocess.run(args, capture_output=True, text=True, env=dict(env, RAYVERB_HRTF_TABLE=str(tmp_path / "short.f32")))
    assert r3.returncode == 1 and "HRTF table" in (r3.stderr + r3.stdout)


def test_c1_golden_vectors():
    """C1: the shipped vault room (3754 triangles, fixture tests/golden/vault_scene.npz),
    1024 rays x 64 reflections, one speaker, against golden oracle vectors committed with the
    script that made them (tests/golden/make_fixtures.py)."""
    sc = np.load(os.path.join(GOLD, "vault_scene.npz"))
    g = np.load(os.path.join(GOLD, "c1_oracle.npz"))
    dirs = scenes.sphere_directions(1024, seed=2015)
    for accel in (api.ACCEL_BRUTE_FORCE, api.ACCEL_BVH):
        rt = api.Raytracer(64, sc["triangles"], sc["vertices"], sc["surfaces"], accel=accel)
        tr = rt.raytrace((0, 1.75, 6), (0, 1.75, 0), dirs, flags=api.TRACE_DEBUG)
        hit, cnt = tr.debug()
        assert np.array_equal(cnt, g["refl_count"]) and np.array_equal(hit, g["hit_tri"])
        d = tr.raw_diffuse().reshape(1024, 64)
        assert np.array_equal(d["position"][:, 0].view(np.uint32), g["first_bounce_position"].view(np.uint32))
        img, keys, klen = tr.raw_images(with_keys=True)
        assert np.array_equal(keys.astype(np.uint64), g["image_keys"]) and np.array_equal(img["time"], g["image_time"])
        ir = tr.render_ir(speakers=[((0, 0, -1), 0.0)], sr=44100.0, ft="linkwitz_riley", lo_cutoff=60.0, trim_predelay=True)
        assert len(ir.channels[0]) == len(g["ir"])
        assert np.linalg.norm(ir.channels[0] - g["ir"]) <= 1e-4 * np.linalg.norm(g["ir"])
        rt.close()
