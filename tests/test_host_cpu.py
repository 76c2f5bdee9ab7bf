"""CPU-only tests of the C++ host mirror: scene loader, config validation and CLI error
behaviour (cmd/main.cpp:104-239 of the reference: every failure prints a message and exits 1)."""
import json
import os
import subprocess

import numpy as np
import pytest

from rayverb_b200 import scenes, types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "parallel-reverb-raytracer_b200", "bin")
GOLD = os.path.join(ROOT, "tests", "golden")
CLI = os.path.join(BIN, "parallel_raytrace")

GOOD = {"rays": 64, "reflections": 4, "sample_rate": 44100, "bit_depth": 16, "source_position": [1, 1, 1],
        "mic_position": [2, 1, 1], "attenuation_model": {"speakers": [{"direction": [0, 0, 1], "shape": 0}]}}


def read_scene_dump(path):
    raw = open(path, "rb").read()
    T, V, S = np.frombuffer(raw, dtype=np.uint64, count=3)
    o = 24
    tri = np.frombuffer(raw, dtype=types.TRIANGLE, count=int(T), offset=o); o += int(T) * 32
    vtx = np.frombuffer(raw, dtype=np.float32, count=int(V) * 4, offset=o).reshape(-1, 4); o += int(V) * 16
    srf = np.frombuffer(raw, dtype=types.SURFACE, count=int(S), offset=o)
    return tri, vtx, srf


@pytest.mark.parametrize("obj", ["large_square.obj", "l_room.obj"])
def test_cpp_loader_matches_python_loader(tmp_path, obj):
    out = tmp_path / "scene.bin"
    subprocess.run([os.path.join(BIN, "dump_scene"), os.path.join(GOLD, obj), os.path.join(GOLD, "materials.json"), str(out)], check=True)
    tri, vtx, srf = read_scene_dump(out)
    ptri, pvtx, psrf = scenes.load_obj(os.path.join(GOLD, obj), os.path.join(GOLD, "materials.json"))
    assert np.array_equal(tri, ptri) and np.array_equal(vtx, pvtx) and np.array_equal(srf, psrf)
    assert len(srf) == 4                                    # default + 3 materials sorted by name
    assert np.allclose(srf["specular"][0], scenes.DEFAULT_SPECULAR)
    assert np.allclose(srf["specular"][1][:2], [0.80, 0.88])  # "brick" < "carpet" < "plaster"


def test_loader_semantics():
    tri, vtx, srf = scenes.load_obj(os.path.join(GOLD, "l_room.obj"), os.path.join(GOLD, "materials.json"))
    assert len(vtx) == 6 + 6 + 4 * 6                        # vertices are not shared between faces
    assert len(tri) == 4 + 4 + 2 * 6
    assert list(tri["surface"][:4]) == [2, 2, 2, 2]         # carpet
    assert list(tri["surface"][-6:]) == [0] * 6             # unknown material -> default surface
    # the concave floor is covered exactly once: triangle areas add up to the L's area
    def area(t):
        a, b, c = (vtx[int(t[k])][:3].astype(np.float64) for k in ("v0", "v1", "v2"))
        return 0.5 * np.linalg.norm(np.cross(b - a, c - a))
    assert abs(sum(area(t) for t in tri[:4]) - (6 * 2 + 3 * 3)) < 1e-9
    tri2, _, _ = scenes.load_obj(os.path.join(GOLD, "large_square.obj"), os.path.join(GOLD, "materials.json"))
    dtri, dvtx, _ = scenes.diamond_room()
    assert np.array_equal(tri2, dtri)


def run_cli(tmp_path, cfg, out_name="out.wav", model="large_square.obj", raw_cfg=None):
    cfg_path = tmp_path / "cfg.json"
    cfg_path.write_text(raw_cfg if raw_cfg is not None else json.dumps(cfg))
    return subprocess.run([CLI, str(cfg_path), os.path.join(GOLD, model), os.path.join(GOLD, "materials.json"),
                           str(tmp_path / out_name)], capture_output=True, text=True)


def test_cli_argument_and_file_checks(tmp_path):
    r = subprocess.run([CLI, "a", "b"], capture_output=True, text=True)
    assert r.returncode == 1 and "Command-line parameters are" in r.stderr
    r = subprocess.run([CLI, "/nonexistent.json", "x.obj", "m.json", str(tmp_path / "o.wav")], capture_output=True, text=True)
    assert r.returncode == 1 and "does not exist" in r.stderr
    r = subprocess.run([CLI, os.path.join(GOLD, "materials.json"), os.path.join(GOLD, "l_room.obj"), os.path.join(GOLD, "materials.json"),
                        "/nonexistent_dir/o.wav"], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot be written" in r.stderr


@pytest.mark.parametrize("mutate,message", [
    (lambda c: c.pop("rays"), "key rays not found in config object"),
    (lambda c: c.pop("attenuation_model"), "key attenuation_model not found in config object"),
    (lambda c: c.update(rays=1.5), "invalid value"),
    (lambda c: c.update(source_position=[0, 1]), "invalid value"),
    (lambda c: c.update(hipass=False), "invalid value"),                  # shipped configs with "hipass": false fail like this
    (lambda c: c.update(filter="butterworth"), "invalid value"),
    (lambda c: c.update(output_mode="everything"), "invalid value"),
    (lambda c: c.update(attenuation_model={"speakers": [], "hrtf": {"facing": [0, 0, 1], "up": [0, 1, 0]}}), "invalid value"),
    (lambda c: c.update(attenuation_model={"speakers": [{"direction": [0, 0, 1]}]}), "key shape not found in config object"),
    (lambda c: c.update(bit_depth=8), "Invalid bitdepth"),
])
def test_cli_config_validation(tmp_path, mutate, message):
    cfg = json.loads(json.dumps(GOOD))
    mutate(cfg)
    r = run_cli(tmp_path, cfg)
    assert r.returncode == 1, r.stderr
    assert message in r.stderr


def test_cli_parse_errors_and_extension(tmp_path):
    r = run_cli(tmp_path, None, raw_cfg='{"rays": 10,,}')
    assert r.returncode == 1 and "Encountered error while parsing config file" in r.stderr
    r = run_cli(tmp_path, None, raw_cfg='[1, 2, 3]')
    assert r.returncode == 1 and "Rayverb config must be stored in a JSON object" in r.stderr
    r = run_cli(tmp_path, GOOD, out_name="out.mp3")
    assert r.returncode == 1 and "Invalid output file extension" in r.stderr


def test_cli_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run_cli(tmp_path, GOOD)
    assert r.returncode == 1 and "no CUDA device" in r.stderr   # no CPU fallback


def test_diagnostic_ray_dump_format(tmp_path):
    """print_diagnostic (helpers.cpp:19-59): one JSON array per ray, {"position", "volume" = band mean}
    per reflection -- what viewer/viewer.pde reads."""
    out = tmp_path / "impulse.dump"
    subprocess.run([os.path.join(BIN, "dump_scene"), "--diagnostic", str(out)], check=True)
    lines = out.read_text().splitlines()
    assert len(lines) == 2
    for r, line in enumerate(lines):
        refl = json.loads(line)
        assert len(refl) == 3
        for j, e in enumerate(refl):
            i = 3 * r + j
            assert list(e.keys()) == ["position", "volume"]
            assert e["position"] == [float(i), 27.0, float(np.float32(-0.1) * np.float32(i))]
            assert e["volume"] == float(np.float32(0.125 * (i + 1) * (4 * 1.0 + 4 * -0.5)) / np.float32(8))
    assert "27.0" in lines[0]                                   # rapidjson writes integral doubles with a fraction
