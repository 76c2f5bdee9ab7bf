"""BVH path (persistent wavefront kernel + microphone visibility map) against the GPU brute-force
kernel -- the intermediate oracle at sizes the CPU cannot sweep -- and against the CPU oracle on
larger subsamples of the C3 / C4 meshes.  Everything is compared bit for bit: both GPU paths run the
same device arithmetic on a superset / the full set of the triangles that can pass the reference's
test (kernel.cpp:63-88,167-192,274-296)."""
import numpy as np
import pytest

from rayverb_b200 import api, scenes
from test_gpu_parity import check_trace_against_oracle, _render_case

pytestmark = pytest.mark.gpu


def _both_paths(tri, vtx, srf, mic, src, dirs, nrefl, flags=api.TRACE_DEBUG):
    out = {}
    for name, accel in (("bf", api.ACCEL_BRUTE_FORCE), ("bvh", api.ACCEL_BVH)):
        rt = api.Raytracer(nrefl, tri, vtx, srf, accel=accel)
        tr = rt.raytrace(mic, src, dirs, flags=flags)
        out[name] = dict(debug=tr.debug(), diffuse=tr.raw_diffuse(), images=tr.raw_images(with_keys=True), stats=tr.stats())
        rt.close()
    return out["bf"], out["bvh"]


def _assert_identical(a, b):
    assert np.array_equal(a["debug"][1], b["debug"][1]), "reflection counts"
    assert np.array_equal(a["debug"][0], b["debug"][0]), "hit triangles"
    assert a["diffuse"].tobytes() == b["diffuse"].tobytes(), "diffuse impulses"
    (ia, ka, la), (ib, kb, lb) = a["images"], b["images"]
    assert np.array_equal(ka, kb) and np.array_equal(la, lb) and ia.tobytes() == ib.tobytes(), "image sources"
    assert a["stats"]["queries"] == b["stats"]["queries"]
    assert a["stats"]["bounces"] == b["stats"]["bounces"]
    assert a["stats"]["emitted"] == b["stats"]["emitted"]


def test_pillar_room_occluded_visibility_equals_bruteforce():
    """Free-standing pillars: a large share of the hit points does not see the microphone, so the
    visibility map's `blocked` answers are exercised, not just its `visible` ones."""
    tri, vtx, srf = scenes.pillar_room(4000)
    assert len(tri) <= 6452
    dirs = scenes.sphere_directions(16384, seed=31)
    a, b = _both_paths(tri, vtx, srf, (12.5, 1.4, 8.5), (2.0, 1.6, 2.5), dirs, 40)
    _assert_identical(a, b)
    st = b["stats"]
    assert st["micmap_cells"] > 0 and st["micmap_entries"] > 0
    skipped = 1.0 - st["emitted"] / st["bounces"]
    assert 0.05 < skipped < 0.95, skipped           # occlusion really happens in this scene


@pytest.mark.parametrize("mic", [(8.0, 0.11, 6.0), (0.12, 3.0, 0.2), (15.6, 5.8, 11.7)])
def test_microphone_next_to_a_wall_equals_bruteforce(mic):
    """Grazing visibility segments (microphone centimetres from the floor / a corner): the worst case
    for the map's conservative binning; also fills its near list."""
    tri, vtx, srf = scenes.pillar_room(3000, npillars=3)
    dirs = scenes.sphere_directions(6000, seed=5)
    a, b = _both_paths(tri, vtx, srf, mic, (4.0, 1.5, 3.0), dirs, 24)
    _assert_identical(a, b)


def test_no_images_flag_and_ragged_reflection_counts_equal_bruteforce():
    """RVB_TRACE_NO_IMAGES starts the wavefront kernel at reflection 0; 7 / 13 reflections leave a
    partial last slot group."""
    tri, vtx, srf = scenes.pillar_room(2500, npillars=2)
    dirs = scenes.sphere_directions(3000, seed=77)
    for nrefl in (7, 13):
        a, b = _both_paths(tri, vtx, srf, (12.5, 1.4, 8.5), (2.0, 1.6, 2.5), dirs, nrefl, flags=api.TRACE_DEBUG | api.TRACE_NO_IMAGES)
        _assert_identical(a, b)


def test_open_scene_rays_leave_the_mesh_equals_bruteforce():
    """A room without its ceiling: rays escape (kernel.cpp:372-375) at different reflections, which is
    what the wavefront kernel's lane refill compacts away."""
    tri, vtx = scenes.displaced_box(12.0, 5.0, 9.0, 3000, amp=0.05, nsurfaces=4)
    keep = ~((vtx[tri["v0"].astype(np.int64), 1] > 4.5) & (vtx[tri["v1"].astype(np.int64), 1] > 4.5) & (vtx[tri["v2"].astype(np.int64), 1] > 4.5))
    tri = tri[keep]
    dirs = scenes.sphere_directions(8192, seed=3)
    a, b = _both_paths(tri, vtx, scenes.hall_surfaces(), (9.0, 1.5, 6.0), (3.0, 1.5, 2.0), dirs, 32)
    _assert_identical(a, b)
    cnt = b["debug"][1]
    assert cnt.min() < cnt.max() and cnt.max() >= 10   # rays escape at different reflections


def test_c3_hall_2048_rays_x32_vs_oracle(orc):
    """C3 mesh (~100k triangles), 2048 rays x 32 reflections against the CPU brute-force oracle."""
    tri, vtx, srf = scenes.concert_hall()
    dirs = scenes.sphere_directions(2048, seed=2015)
    check_trace_against_oracle(orc, tri, vtx, srf, (24, 1.5, 18), (20, 2.0, 4), dirs, 32, api.ACCEL_AUTO)


def test_c4_cathedral_512_rays_x16_hrtf_vs_oracle(orc):
    """C4 mesh (~1M triangles), 512 rays x 16 reflections, HRTF output, against the CPU oracle."""
    tri, vtx, srf = scenes.cathedral()
    dirs = scenes.sphere_directions(512, seed=2016)
    mic, src = (15, 1.7, 60), (15, 2.0, 10)
    check_trace_against_oracle(orc, tri, vtx, srf, mic, src, dirs, 16, api.ACCEL_AUTO)
    _render_case(orc, tri, vtx, srf, mic, src, dirs, 16, api.ACCEL_AUTO,
                 hrtf=dict(table=scenes.synthetic_hrtf_table(), facing=(0, 0, -1), up=(0, 1, 0)))


@pytest.mark.parametrize("accel,ntri", [(api.ACCEL_BVH, 6000), (api.ACCEL_BRUTE_FORCE, 1500)])
def test_batched_trace_equals_individual_traces(accel, ntri):
    """rvb_trace_run_batch (parameter sweeps, cmd/main.cpp:241-297 once per pair): every pair of the batch
    gives, bit for bit, the trace and the impulse response of the same pair traced alone."""
    import torch
    tri, vtx, srf = scenes.pillar_room(ntri, npillars=3)
    nrefl, nrays = 24, 5000
    dirs = scenes.sphere_directions(nrays, seed=11)
    d_dirs = torch.from_numpy(dirs).to("cuda")
    pairs = [((12.5 - 1.1 * i, 1.4 + 0.2 * i, 8.5 - 0.6 * i), (2.0 + 0.9 * i, 1.6, 2.5 + 0.4 * i)) for i in range(7)]
    kw = dict(speakers=[((-1, 0, -1), 0.5), ((1, 0, -1), 0.5)], sr=44100.0, ft="linkwitz_riley", lo_cutoff=45.0)
    rt = api.Raytracer(nrefl, tri, vtx, srf, accel=accel)
    batch = rt.raytrace_batch(pairs, device_ptr=d_dirs.data_ptr(), nrays=nrays, flags=api.TRACE_DEBUG)
    assert len(batch) == len(pairs)
    for (mic, src), tb in zip(pairs, batch):
        one = rt.raytrace(mic, src, None, device_ptr=d_dirs.data_ptr(), nrays=nrays, flags=api.TRACE_DEBUG)
        (ha, ca), (hb, cb) = one.debug(), tb.debug()
        assert np.array_equal(ca, cb) and np.array_equal(ha, hb)
        assert one.raw_diffuse().tobytes() == tb.raw_diffuse().tobytes()
        (ia, ka, la), (ib, kb, lb) = one.raw_images(with_keys=True), tb.raw_images(with_keys=True)
        assert np.array_equal(ka, kb) and np.array_equal(la, lb) and ia.tobytes() == ib.tobytes()
        sa, sb = one.stats(), tb.stats()
        for k in ("queries", "tri_tests", "bounces", "emitted"):
            assert sa[k] == sb[k], k
        ir_a, ir_b = one.render_ir(**kw), tb.render_ir(**kw)
        for x, y in zip(ir_a.channels, ir_b.channels):
            assert x.tobytes() == y.tobytes()
        tb.close()
    rt.close()
