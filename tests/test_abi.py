"""CPU-only: the C-ABI library loads and exports every symbol include/rayverb_b200.h declares;
compute calls fail loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from rayverb_b200 import api, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "rayverb_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rvb_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(api.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    L = ctypes.CDLL(api.LIB_PATH)
    declared = header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in the header but not exported"
    assert sorted(api.EXPORTS) == declared


def test_pod_layouts_match_reference_sizes():
    # rayverb/clstructs.h:13-58
    from rayverb_b200 import types
    assert types.TRIANGLE.itemsize == 32 and types.SURFACE.itemsize == 64
    assert types.IMPULSE.itemsize == 64 and types.IMPULSE.fields["position"][1] == 32 and types.IMPULSE.fields["time"][1] == 48
    assert types.ATTENUATED.itemsize == 64 and types.ATTENUATED.fields["time"][1] == 32
    assert types.SPEAKER.itemsize == 32 and types.SPEAKER.fields["coefficient"][1] == 16
    assert ctypes.sizeof(api.RenderConfig) == 96


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    tri, vtx, srf = scenes.shoebox()
    with pytest.raises(api.RayverbError):
        api.Raytracer(4, tri, vtx, srf)
    with pytest.raises(api.RayverbError):
        api.flatten_impulses(np.zeros(4, dtype=api.ATTENUATED), 44100.0)


def test_scene_validation_is_checked_before_the_device():
    tri, vtx, srf = scenes.shoebox()
    bad = tri.copy()
    bad["v1"][3] = 10_000
    with pytest.raises(api.RayverbError, match="out-of-range"):
        api.Raytracer(4, bad, vtx, srf)
