"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs, and against the reference's own gtest vectors.

Bar (BASELINE.json north_star / SURVEY.md 8d): hit triangle indices, reflection counts, positions
and times bit-exact; volumes within 2e-6 relative (pow is a library function on both sides);
every IR band within relative L2 <= 1e-4.
"""
import numpy as np
import pytest

from rayverb_b200 import api, scenes
from hrtf_table import smooth_table, test_table as make_test_table
from test_oracle_golden import assert_float_eq, kat_directions, _attenuation_inputs

pytestmark = pytest.mark.gpu

VOL_RTOL = 2e-6
IR_L2_TOL = 1e-4     # north_star: relative L2 <= 1e-4 per band


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / d if d > 0 else np.linalg.norm(a)


def assert_filter_close(got, want, what):
    """The chunk-parallel IIR runs the reference's own double recurrence inside every 16-sample
    chunk; the chunk start states come from shuffle scans (lanes, warps) and a fold over the tiles with
    powers of the state-transition matrix, which is associated differently from the serial filter (double precision, but with
    cancellation between large terms for the low bands' near-unit poles).  A start state that is off
    by ~1e-13 relative occasionally flips the float rounding of one output sample, and that one-ulp
    difference feeds the later sweeps of a two-pass / Linkwitz-Riley chain.  Bar: every sample
    within 2e-7 of the peak (about one ulp) and at most 1 % of the samples different at all."""
    got = np.asarray(got, np.float32); want = np.asarray(want, np.float32)
    mism = bits(got) != bits(want)
    assert mism.mean() <= 1e-2, (what, mism.mean())
    scale = np.abs(want).max()
    assert np.abs(got.astype(np.float64) - want).max() <= 2e-7 * scale, what
    assert rel_l2(got, want) <= 1e-7, what


def check_trace_against_oracle(orc, tri, vtx, srf, mic, src, dirs, nrefl, accel):
    rt = api.Raytracer(nrefl, tri, vtx, srf, accel=accel)
    tr = rt.raytrace(mic, src, dirs, flags=api.TRACE_DEBUG)
    ref = orc.raytrace(tri, vtx, srf, mic, src, dirs, nrefl)
    hit, cnt = tr.debug()
    assert np.array_equal(cnt, ref["refl_count"])
    assert np.array_equal(hit, ref["hit_tri"])
    d = tr.raw_diffuse()
    assert np.array_equal(bits(d["position"]), bits(ref["diffuse"]["position"]))
    assert np.array_equal(bits(d["time"]), bits(ref["diffuse"]["time"]))
    np.testing.assert_allclose(d["volume"], ref["diffuse"]["volume"], rtol=VOL_RTOL, atol=1e-30)
    assert np.array_equal(d["volume"] == 0, ref["diffuse"]["volume"] == 0)
    # image sources: same keys in the same (lexicographic) order, same winners
    img, keys, klen = tr.raw_images(with_keys=True)
    rimg, rkeys, rklen = orc.image_dedupe(ref["image"], ref["image_idx"])
    assert np.array_equal(klen, rklen)
    assert np.array_equal(keys.astype(np.uint64), rkeys)
    assert np.array_equal(bits(img["position"]), bits(rimg["position"]))
    assert np.array_equal(bits(img["time"]), bits(rimg["time"]))
    np.testing.assert_allclose(img["volume"], rimg["volume"], rtol=VOL_RTOL, atol=1e-30)
    img2 = tr.raw_images(remove_direct=True)
    rimg2, _, _ = orc.image_dedupe(ref["image"], ref["image_idx"], remove_direct=True)
    assert len(img2) == len(rimg2) == len(rimg) - 1
    st = tr.stats()
    if accel == api.ACCEL_BRUTE_FORCE:   # the brute-force path executes exactly the reference's tests
        direct_tests_skipped = (len(dirs) - 1) * len(tri)   # only ray 0 runs the source->mic query
        assert st["tri_tests"] == ref["ntests"] - direct_tests_skipped
    assert st["bounces"] == int(ref["refl_count"].sum())
    rt.close()
    return ref


def test_kat_impulse_directions():
    """tests/raytrace_tests.h:30-48 through the C ABI."""
    tri, vtx, srf = scenes.diamond_room()
    nrefl = 128
    rt = api.Raytracer(nrefl, tri, vtx, srf, accel=api.ACCEL_BRUTE_FORCE)
    rt.raytrace((0, 2, 0), (0, 2, 2), kat_directions(64 * 1000))
    d = rt.getRawDiffuse().reshape(-1, nrefl)
    first = [(0, 2, -27), (0, 2, 27), (0, 0, 2), (0, 27, 2), (-25, 2, 2), (25, 2, 2)]
    second = [(0, 0, 0), (0, 0, 0), (0, 27, 2), (0, 0, 2), (-25, 2, -2), (25, 2, -2)]
    for i in range(6):
        assert_float_eq(d[i, 0]["position"][:3], first[i])
        assert_float_eq(d[i, 1]["position"][:3], second[i])
    rt.close()


@pytest.mark.parametrize("coef,expect", [(0.0, [1, 1, 1, 1, 1, 1]), (0.5, [.5, .5, .5, .5, 0, 1]), (1.0, [0, 0, 0, 0, -1, 1])])
def test_kat_attenuate_speaker(orc, coef, expect):
    """tests/attenuation_tests.h:67-101 through the C ABI, plus bit parity with the oracle."""
    imp = _attenuation_inputs(1024 * 64)
    out = api.SpeakerAttenuator().attenuate(imp, (0, 0, 0), [((0, 0, 1), coef)])[0]
    for b in range(1, 8):
        assert_float_eq(out["volume"][:, 0], out["volume"][:, b])
    assert_float_eq(out["volume"][:6, 0], expect)
    if coef == 0.0:
        assert_float_eq(out["volume"][:, 0], np.ones(len(out)))
    assert np.array_equal(out["time"], imp["time"])
    ref = orc.attenuate_speaker(imp, (0, 0, 0), (0, 0, 1), coef)
    assert np.array_equal(bits(out["volume"]), bits(ref["volume"]))


@pytest.mark.parametrize("facing,order", [((0, 0, 1), (5, 4, 0, 1)), ((1, 0, 0), (1, 0, 5, 4)),
                                          ((0, 0, -1), (4, 5, 1, 0)), ((-1, 0, 0), (0, 1, 4, 5))])
def test_kat_hrtf(facing, order):
    """tests/hrtf_tests.cpp:42-85 through the C ABI."""
    table = make_test_table()
    imp = _attenuation_inputs(1000)
    out = api.HrtfAttenuator(table).attenuate(imp, (0, 0, 0), facing, (0, 1, 0), channels=(0,))[0]
    for cell, k in zip([(180, 90), (0, 90), (90, 90), (270, 90)], order):
        assert_float_eq(table[0, cell[0], cell[1]], out["volume"][k])


def test_attenuators_random_vs_oracle(orc):
    rng = np.random.default_rng(11)
    n = 20000
    imp = np.zeros(n, dtype=api.IMPULSE)
    imp["volume"] = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
    imp["volume"][rng.random(n) < 0.1] = 0
    imp["position"][:, :3] = rng.uniform(-20, 20, (n, 3)).astype(np.float32)
    imp["time"] = rng.uniform(0, 2, n).astype(np.float32)
    mic = (1.0, 1.5, -2.0)
    out = api.SpeakerAttenuator().attenuate(imp, mic, [((-1, 0, -1), 0.5), ((0.3, 0.2, 1), 0.8)])
    for o, (d, c) in zip(out, [((-1, 0, -1), 0.5), ((0.3, 0.2, 1), 0.8)]):
        ref = orc.attenuate_speaker(imp, mic, d, c)
        assert np.array_equal(bits(o["volume"]), bits(ref["volume"])) and np.array_equal(bits(o["time"]), bits(ref["time"]))
    table = smooth_table()
    facing, up = (0.6, 0.0, -0.8), (0, 1, 0)
    outs = api.HrtfAttenuator(table).attenuate(imp, mic, facing, up)
    for ch, o in enumerate(outs):
        ref = orc.attenuate_hrtf(imp, mic, table[ch], facing, up, ch)
        assert np.array_equal(bits(o["time"]), bits(ref["time"]))
        # atan2 is a library function on both sides: a 1-ulp difference can flip the degree
        # truncation at an integer boundary (documented near-tie class); allow a handful -- and every one
        # of them must be the NEIGHBOURING table cell (azimuth or elevation index off by one), not noise
        mism = np.flatnonzero(np.any(bits(o["volume"]) != bits(ref["volume"]), axis=1))
        assert len(mism) < 1e-3 * n
        tab = table[ch].reshape(360, 180, 8)
        for i in mism:
            gain = o["volume"][i] / imp["volume"][i]
            gref = ref["volume"][i] / imp["volume"][i]
            a, e = np.unravel_index(np.argmin(np.abs(tab - gref[None, None, :]).sum(-1)), (360, 180))
            neigh = [tab[(a + da) % 360, min(179, max(0, e + de))] for da in (-1, 0, 1) for de in (-1, 0, 1)]
            assert min(np.abs(nb - gain).max() for nb in neigh) <= 1e-5 * np.abs(gain).max(), (i, a, e)


def test_trace_shoebox_vs_oracle(orc):
    tri, vtx, srf = scenes.shoebox()
    dirs = scenes.sphere_directions(3000, seed=2015)
    check_trace_against_oracle(orc, tri, vtx, srf, (7, 2, 4), (2, 1.5, 2), dirs, 40, api.ACCEL_BRUTE_FORCE)


def test_trace_ragged_and_escaping_rays(orc):
    """Edge cases: ray count not a multiple of anything, zero directions, rays that leave the mesh
    (open scene: one wall removed), few reflections (< 9 image-source orders)."""
    tri, vtx, srf = scenes.shoebox()
    tri = tri[2:]                                  # drop the x=0 wall: rays escape
    dirs = scenes.sphere_directions(777, seed=3)
    dirs[5] = 0                                    # zero direction: no hit at all
    check_trace_against_oracle(orc, tri, vtx, srf, (7, 2, 4), (2, 1.5, 2), dirs, 5, api.ACCEL_BRUTE_FORCE)
    check_trace_against_oracle(orc, tri, vtx, srf, (7, 2, 4), (2, 1.5, 2), dirs[:1], 1, api.ACCEL_BRUTE_FORCE)


def test_trace_displaced_room_vs_oracle(orc):
    """A few hundred triangles with several surfaces: brute force from shared memory."""
    tri, vtx = scenes.displaced_box(8.0, 3.0, 5.0, 700, amp=0.08, fan=0.2, nsurfaces=4)
    srf = scenes.hall_surfaces()
    dirs = scenes.sphere_directions(1500, seed=9)
    check_trace_against_oracle(orc, tri, vtx, srf, (5.5, 1.2, 3.5), (2, 1.5, 2), dirs, 24, api.ACCEL_BRUTE_FORCE)


def test_warp_per_ray_equals_thread_per_ray():
    """Few rays over many triangles run one WARP per ray (lanes split the triangle records, shuffle
    argmin on (t, index)); many rays run one thread per ray.  Rays are independent, so the first 1500
    rays of a 120 000-ray trace must come out bit-identical when they are traced on their own."""
    tri, vtx = scenes.displaced_box(8.0, 3.0, 5.0, 700, amp=0.08, fan=0.2, nsurfaces=4)
    srf = scenes.hall_surfaces()
    dirs = scenes.sphere_directions(120_000, seed=9)
    mic, src, nrefl, k = (5.5, 1.2, 3.5), (2, 1.5, 2), 14, 1500
    out = []
    for n in (len(dirs), k):
        rt = api.Raytracer(nrefl, tri, vtx, srf, accel=api.ACCEL_BRUTE_FORCE)
        tr = rt.raytrace(mic, src, dirs[:n], flags=api.TRACE_DEBUG)
        hit, cnt = tr.debug()
        d = tr.raw_diffuse().reshape(n, nrefl)
        st = tr.stats()
        out.append((hit[:k].copy(), cnt[:k].copy(), d[:k].copy(), st))
        rt.close()
    (hit_a, cnt_a, d_a, st_a), (hit_b, cnt_b, d_b, st_b) = out
    assert np.array_equal(hit_a, hit_b) and np.array_equal(cnt_a, cnt_b)
    assert d_a.tobytes() == d_b.tobytes()
    # both forms execute exactly the reference's tests: queries x triangles
    assert st_a["tri_tests"] == st_a["queries"] * len(tri) and st_b["tri_tests"] == st_b["queries"] * len(tri)
    assert st_b["bounces"] == int(cnt_b.sum())


def _render_case(orc, tri, vtx, srf, mic, src, dirs, nrefl, accel, **kw):
    rt = api.Raytracer(nrefl, tri, vtx, srf, accel=accel)
    tr = rt.raytrace(mic, src, dirs)
    ir = tr.render_ir(keep_bands=True, **kw)
    okw = dict(kw)
    ref, flat, bands = orc.render_ir(tri, vtx, srf, mic, src, dirs, nrefl, return_bands=True, **okw)
    assert len(ir.channels) == len(ref)
    for c in range(len(ref)):
        assert ir.bands[c].shape == bands[c].shape           # bin indices: same flatten length
        for b in range(8):
            assert rel_l2(ir.bands[c][b], bands[c][b]) <= IR_L2_TOL, (c, b)
        assert len(ir.channels[c]) == len(ref[c])
        assert rel_l2(ir.channels[c], ref[c]) <= IR_L2_TOL
    ir2 = tr.render_ir(**kw)                                   # deterministic: bit-identical rerun
    for c in range(len(ref)):
        assert np.array_equal(bits(ir.channels[c]), bits(ir2.channels[c]))
    rt.close()
    return ir, ref


@pytest.mark.parametrize("ft", ["linkwitz_riley", "onepass", "twopass"])
def test_render_ir_stereo_speakers_vs_oracle(orc, ft):
    tri, vtx, srf = scenes.shoebox()
    dirs = scenes.sphere_directions(4096, seed=2015)
    _render_case(orc, tri, vtx, srf, (7, 2, 4), (2, 1.5, 2), dirs, 48, api.ACCEL_BRUTE_FORCE,
                 speakers=[((-1, 0, -1), 0.5), ((1, 0, -1), 0.5)], ft=ft, lo_cutoff=60.0, trim_predelay=True)


def test_render_ir_modes_vs_oracle(orc):
    tri, vtx, srf = scenes.shoebox()
    dirs = scenes.sphere_directions(2000, seed=4)
    for mode, rd in [("image_only", True), ("diffuse_only", False), ("all", True)]:
        _render_case(orc, tri, vtx, srf, (7, 2, 4), (2, 1.5, 2), dirs, 20, api.ACCEL_BRUTE_FORCE,
                     speakers=[((0, 0, 1), 0.0)], output_mode=mode, remove_direct=rd, do_normalize=False,
                     volume_scale=0.5, do_trim_tail=False)


def test_render_ir_hrtf_vs_oracle(orc):
    tri, vtx, srf = scenes.shoebox()
    dirs = scenes.sphere_directions(3000, seed=8)
    table = smooth_table()
    _render_case(orc, tri, vtx, srf, (7, 2, 4), (2, 1.5, 2), dirs, 32, api.ACCEL_BRUTE_FORCE,
                 hrtf=dict(table=table, facing=(0, 0, -1), up=(0, 1, 0)), trim_predelay=True)


def test_host_array_postprocessing_vs_oracle(orc):
    """rvb_flatten / rvb_filter / rvb_process against flattenImpulses / filter / process."""
    rng = np.random.default_rng(5)
    n = 50000
    att = np.zeros(n, dtype=api.ATTENUATED)
    att["volume"] = (rng.standard_normal((n, 8)) * np.exp(-rng.uniform(0, 6, (n, 1)))).astype(np.float32)
    att["time"] = rng.uniform(0, 0.4, n).astype(np.float32)
    att["time"][:10] = 0
    sr = 44100.0
    f = api.flatten_impulses(att, sr)
    rf = orc.flatten(att, sr)
    assert f.shape == rf.shape
    assert rel_l2(f, rf) < 1e-6
    empty = api.flatten_impulses(att[:0], sr)
    assert empty.shape == (8, 1)                      # round(0)+1 samples, all zero
    for ft in ("onepass", "twopass", "linkwitz_riley"):
        g = api.filter_bands(ft, rf[None], sr, 45.0)[0]
        rg = orc.filter_channel(ft, rf, sr, 45.0)
        assert_filter_close(g, rg, ft)
    out = api.process("linkwitz_riley", [rf, rf[:, :9000]], sr, True, 45.0, True, 1.0)
    ref = orc.process("linkwitz_riley", [rf, rf[:, :9000]], sr, True, 45.0, True, 1.0)
    for a, b in zip(out, ref):
        assert len(a) == len(b)
        assert_filter_close(a, b, "process")


@pytest.mark.parametrize("length", [1, 15, 17, 1000, 16384, 16385, 40000, 131072, 300001, 1_200_000])
def test_band_filters_across_tiles_and_cluster_sizes(orc, length):
    """The IIR kernel cuts a row into equal tiles of at most 16384 samples, the same number for each of
    the 1, 2, 4 or 8 CTAs of the row's thread-block cluster, and keeps up to 8 tiles per CTA between
    its two phases: rows shorter than a chunk, lengths around the tile edges, up to 19 tiles, and a row of
    1.2 M samples (more than 8 x 8 tiles: two rounds), two channels of different content, against the
    serial reference filter (Biquad / twopass / LinkwitzRiley, filters.cpp:156-266)."""
    rng = np.random.default_rng(length)
    nch = 2 if length < 1_000_000 else 1
    x = (rng.standard_normal((nch, 8, length)) * np.exp(-np.linspace(0, 8, length))[None, None, :]).astype(np.float32)
    x[nch - 1, :, ::7] = 0.0
    for ft in ("onepass", "twopass", "linkwitz_riley"):
        g = api.filter_bands(ft, x, 44100.0, 45.0)
        for c in range(nch):
            assert_filter_close(g[c], orc.filter_channel(ft, x[c], 44100.0, 45.0), (ft, length, c))


# ---- BVH path: must equal the brute-force loop (kernel.cpp:167-192) including its tie-break -----
def test_bvh_small_room_vs_oracle(orc):
    tri, vtx = scenes.displaced_box(8.0, 3.0, 5.0, 700, amp=0.08, fan=0.2, nsurfaces=4)
    srf = scenes.hall_surfaces()
    dirs = scenes.sphere_directions(1500, seed=9)
    check_trace_against_oracle(orc, tri, vtx, srf, (5.5, 1.2, 3.5), (2, 1.5, 2), dirs, 24, api.ACCEL_BVH)


def test_bvh_shoebox_and_diamond_vs_oracle(orc):
    """Large axis-aligned triangles (boxes as big as the scene, rays parallel to faces)."""
    tri, vtx, srf = scenes.shoebox()
    check_trace_against_oracle(orc, tri, vtx, srf, (7, 2, 4), (2, 1.5, 2), scenes.sphere_directions(2000, seed=1), 30, api.ACCEL_BVH)
    tri, vtx, srf = scenes.diamond_room()
    check_trace_against_oracle(orc, tri, vtx, srf, (0, 2, 0), (0, 2, 2), kat_directions(300), 20, api.ACCEL_BVH)


def test_bvh_equals_bruteforce_kernel_4k_triangles():
    """GPU brute force (shared-memory scene) is the intermediate oracle at sizes the CPU cannot
    sweep: same hits, counts, positions, times and bit-identical volumes (same device arithmetic)."""
    tri, vtx = scenes.displaced_box(20.0, 8.0, 14.0, 4400, amp=0.1, fan=0.25, nsurfaces=4)
    assert len(tri) <= 6452
    srf = scenes.hall_surfaces()
    dirs = scenes.sphere_directions(20000, seed=21)
    out = {}
    for name, accel in (("bf", api.ACCEL_BRUTE_FORCE), ("bvh", api.ACCEL_BVH)):
        rt = api.Raytracer(48, tri, vtx, srf, accel=accel)
        tr = rt.raytrace((12, 1.5, 9), (6, 2, 3), dirs, flags=api.TRACE_DEBUG)
        out[name] = (tr.debug(), tr.raw_diffuse(), tr.raw_images(with_keys=True), tr.stats())
        rt.close()
    (hit_a, cnt_a), d_a, (img_a, k_a, l_a), st_a = out["bf"]
    (hit_b, cnt_b), d_b, (img_b, k_b, l_b), st_b = out["bvh"]
    assert np.array_equal(cnt_a, cnt_b) and np.array_equal(hit_a, hit_b)
    assert d_a.tobytes() == d_b.tobytes()
    assert np.array_equal(k_a, k_b) and np.array_equal(l_a, l_b) and img_a.tobytes() == img_b.tobytes()
    assert st_b["tri_tests"] < st_a["tri_tests"] / 20          # the BVH actually culls
    assert st_a["queries"] == st_b["queries"]


@pytest.mark.parametrize("open_room", [False, True])
def test_wavefront_work_units_do_not_show(open_room):
    """The wavefront bounce kernel cuts a ray's steady-phase reflections into work units that different
    lanes (different SMs) execute one after the other, handing the ray state over through HBM
    (trace.cu: WfShared / SEG).  The hand-over record holds the same floats the registers held, so the
    trace must be bit-identical for whole rays (1 unit), for forced cuts -- with 20 000 rays every later
    unit is drawn while its predecessor still runs, which exercises the pending-unit path on nearly every
    lane -- and for the cut the launcher chooses itself once the rays exceed the resident lanes.  Open room:
    rays that leave the mesh void their later units."""
    tri, vtx = scenes.displaced_box(20.0, 8.0, 14.0, 4400, amp=0.1, fan=0.25, nsurfaces=4)
    if open_room:
        tri = tri[len(tri) // 6:]                  # drop most of one wall
    srf = scenes.hall_surfaces()
    mic, src = (12, 1.5, 9), (6, 2, 3)

    def run(dirs, nrefl, units, images=True):
        rt = api.Raytracer(nrefl, tri, vtx, srf, accel=api.ACCEL_BVH)
        flags = api.TRACE_DEBUG | api.TRACE_UNITS(units) | (0 if images else api.TRACE_NO_IMAGES)
        tr = rt.raytrace(mic, src, dirs, flags=flags)
        res = (tr.debug(), tr.raw_diffuse().tobytes(), tr.stats())
        rt.close()
        return res

    dirs = scenes.sphere_directions(20000, seed=21)
    (hit_w, cnt_w), d_w, st_w = run(dirs, 48, 1)
    if open_room:
        assert cnt_w.min() < 48 and cnt_w.max() == 48          # some rays escape, some do not
    for units in (2, 5, 16):
        (hit, cnt), d, st = run(dirs, 48, units)
        assert np.array_equal(cnt, cnt_w) and np.array_equal(hit, hit_w), units
        assert d == d_w, units
        for k in ("queries", "bounces", "tri_tests", "box_tests"):
            assert st[k] == st_w[k], (units, k)
    # no image phase: unit 0 starts from the source, the later ones from hand-over records
    (hit_a, cnt_a), d_a, _ = run(dirs[:5000], 30, 1, images=False)
    (hit_b, cnt_b), d_b, _ = run(dirs[:5000], 30, 7, images=False)
    assert np.array_equal(cnt_a, cnt_b) and np.array_equal(hit_a, hit_b) and d_a == d_b
    # more rays than resident lanes (148 SMs x 5 CTAs x 128): the launcher's own choice against whole rays
    dirs = scenes.sphere_directions(150_000, seed=5)
    (hit_a, cnt_a), d_a, _ = run(dirs, 41, 1)
    (hit_b, cnt_b), d_b, _ = run(dirs, 41, 0)
    assert np.array_equal(cnt_a, cnt_b) and np.array_equal(hit_a, hit_b) and d_a == d_b


def test_image_dedupe_few_candidates_kernel_equals_sort_path():
    """Up to 4096 valid image-source candidates are de-duplicated by one CTA (direct ranking in shared memory,
    images.cu: dedupe_small_kernel) instead of the radix-sort path.  Same keys, same order, same winners
    (the lowest ray of every sequence), same impulses: shoebox with many duplicates per sequence, a displaced
    room with several surfaces, an open room, and the direct path alone (zero reflections)."""
    cases = []
    tri, vtx, srf = scenes.shoebox()
    cases.append((tri, vtx, srf, (7, 2, 4), (2, 1.5, 2), scenes.sphere_directions(400, seed=4), 12, api.ACCEL_BRUTE_FORCE))
    cases.append((tri[2:], vtx, srf, (7, 2, 4), (2, 1.5, 2), scenes.sphere_directions(380, seed=5), 7, api.ACCEL_BRUTE_FORCE))
    cases.append((tri, vtx, srf, (7, 2, 4), (2, 1.5, 2), scenes.sphere_directions(64, seed=6), 0, api.ACCEL_BRUTE_FORCE))
    tri2, vtx2 = scenes.displaced_box(20.0, 8.0, 14.0, 4400, amp=0.1, fan=0.25, nsurfaces=4)
    cases.append((tri2, vtx2, scenes.hall_surfaces(), (12, 1.5, 9), (6, 2, 3), scenes.sphere_directions(30000, seed=7), 20, api.ACCEL_BVH))
    for tri, vtx, srf, mic, src, dirs, nrefl, accel in cases:
        out = []
        for flags in (0, api.TRACE_SORT_DEDUPE):
            rt = api.Raytracer(nrefl, tri, vtx, srf, accel=accel)
            tr = rt.raytrace(mic, src, dirs, flags=flags)
            img, keys, klen = tr.raw_images(with_keys=True)
            img_nd = tr.raw_images(remove_direct=True)
            out.append((img.tobytes(), keys.copy(), klen.copy(), img_nd.tobytes()))
            rt.close()
        a, b = out
        assert len(a[2]) > 0
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        assert a[0] == b[0] and a[3] == b[3]


def test_visibility_map_resolution_does_not_show():
    """The microphone visibility map (micmap.cu) pre-selects, per cube-map cell, the triangles that can block
    the segment hit point -> microphone; the reference's exact test then runs on those candidates.  Its
    resolution is chosen from the query count (4 ... 48 cells per triangle, up to 1024 x 1024 cells per face at
    C3's full size), so a 100k-triangle room with free-standing pillars (a third of the hit points do not see
    the microphone) is traced with the coarsest, the default and the finest map: hits, counts and impulses must
    be bit-identical (a cell that lost an occluder would turn a zero impulse into a non-zero one)."""
    tri, vtx, srf = scenes.pillar_room(100_000, npillars=6, pillar_grid=24)
    dirs = scenes.sphere_directions(100_000, seed=33)
    out = []
    for cells in (2, 0, 48):
        rt = api.Raytracer(48, tri, vtx, srf)
        tr = rt.raytrace((12.5, 1.4, 8.5), (2.0, 1.6, 2.5), dirs, flags=api.TRACE_DEBUG | api.TRACE_MAP_CELLS(cells))
        st = tr.stats()
        out.append((tr.debug(), tr.raw_diffuse().tobytes(), tr.raw_images(with_keys=True), st["micmap_cells"], st["emitted"], st["bounces"]))
        rt.close()
    assert out[0][3] < out[1][3] < out[2][3] and out[2][3] > 512       # three different grids, the last beyond 512 cells per edge
    skipped = 1.0 - out[0][4] / out[0][5]
    assert 0.05 < skipped < 0.95, skipped                               # occlusion really happens in this scene
    for o in out[1:]:
        assert np.array_equal(o[0][0], out[0][0][0]) and np.array_equal(o[0][1], out[0][0][1])
        assert o[1] == out[0][1] and o[4] == out[0][4]
        assert o[2][0].tobytes() == out[0][2][0].tobytes() and np.array_equal(o[2][1], out[0][2][1])


def test_bvh_hall_100k_subsample_vs_oracle(orc):
    """C3 mesh (~100k triangles): a ray subsample against the CPU brute-force oracle."""
    tri, vtx, srf = scenes.concert_hall()
    assert 95_000 < len(tri) < 105_000
    dirs = scenes.sphere_directions(48, seed=2015)
    check_trace_against_oracle(orc, tri, vtx, srf, (24, 1.5, 18), (20, 2.0, 4), dirs, 12, api.ACCEL_AUTO)


def test_c4_cathedral_1m_triangles_hrtf_subsample(orc):
    """C4 mesh (~1M triangles, BVH path) with HRTF output: a ray subsample end to end against the
    CPU brute-force oracle (hits bit-exact, IR within 1e-4 relative L2 per band)."""
    tri, vtx, srf = scenes.cathedral()
    assert 950_000 < len(tri) < 1_050_000
    dirs = scenes.sphere_directions(24, seed=2016)
    mic, src = (15, 1.7, 60), (15, 2.0, 10)
    ref = check_trace_against_oracle(orc, tri, vtx, srf, mic, src, dirs, 6, api.ACCEL_AUTO)
    assert ref["refl_count"].min() >= 1
    _render_case(orc, tri, vtx, srf, mic, src, dirs, 6, api.ACCEL_AUTO,
                 hrtf=dict(table=scenes.synthetic_hrtf_table(), facing=(0, 0, -1), up=(0, 1, 0)))


def test_synthetic_table_matches_test_helper():
    assert np.array_equal(scenes.synthetic_hrtf_table(), smooth_table())


def test_windowed_sinc_filter_vs_oracle(orc):
    """FILTER_TYPE_WINDOWED_SINC: host-array entry points and the fused render."""
    rng = np.random.default_rng(12)
    data = (rng.standard_normal((2, 8, 4000)) * np.exp(-np.arange(4000) / 700.0)).astype(np.float32)
    got = api.filter_bands("sinc", data, 44100.0, 45.0)
    assert got.shape == (2, 8, 4028)
    for c in range(2):
        want = orc.filter_channel_sinc(data[c], 44100.0, 45.0)
        assert np.array_equal(bits(got[c]), bits(want))          # same double sums in the same order
    out = api.process("sinc", [data[0], data[1]], 44100.0, True, 45.0, True, 1.0)
    ref = orc.process("sinc", [data[0], data[1]], 44100.0, True, 45.0, True, 1.0)
    for a, b in zip(out, ref):
        assert len(a) == len(b) and np.array_equal(bits(a), bits(b))
    tri, vtx, srf = scenes.shoebox()
    _render_case(orc, tri, vtx, srf, (7, 2, 4), (2, 1.5, 2), scenes.sphere_directions(1500, seed=6), 16, api.ACCEL_BRUTE_FORCE,
                 speakers=[((-1, 0, -1), 0.5), ((1, 0, -1), 0.5)], ft="sinc", lo_cutoff=60.0)


def test_empty_and_degenerate_inputs(orc):
    """Empty and degenerate inputs the reference's host code accepts: zero rays, zero reflections
    (only the direct path remains), a scene every ray leaves at once."""
    tri, vtx, srf = scenes.shoebox()
    mic, src = (7, 2, 4), (2, 1.5, 2)
    dirs = scenes.sphere_directions(257, seed=3)
    # zero rays
    rt = api.Raytracer(8, tri, vtx, srf, accel=api.ACCEL_BRUTE_FORCE)
    tr = rt.raytrace(mic, src, dirs[:0])
    assert len(tr.raw_diffuse()) == 0 and len(tr.raw_images()) == 0
    ir = tr.render_ir(speakers=[((0, 0, 1), 0.0)], do_normalize=False, do_trim_tail=False)
    assert len(ir.channels) == 1 and len(ir.channels[0]) == 1 and ir.channels[0][0] == 0     # round(0)+1 samples
    rt.close()
    # zero reflections: getRawDiffuse is empty, getRawImages holds the direct path only
    for accel in (api.ACCEL_BRUTE_FORCE, api.ACCEL_BVH):
        rt = api.Raytracer(0, tri, vtx, srf, accel=accel)
        tr = rt.raytrace(mic, src, dirs)
        ref = orc.raytrace(tri, vtx, srf, mic, src, dirs, 0)
        rimg, rkeys, rklen = orc.image_dedupe(ref["image"], ref["image_idx"])
        img, keys, klen = tr.raw_images(with_keys=True)
        assert len(tr.raw_diffuse()) == 0 and len(img) == len(rimg) == 1 and klen[0] == 1
        assert np.array_equal(bits(img["time"]), bits(rimg["time"]))
        np.testing.assert_allclose(img["volume"], rimg["volume"], rtol=VOL_RTOL)
        assert len(tr.raw_images(remove_direct=True)) == 0
        rt.close()
    # a single triangle far away: every ray escapes without a hit; the direct path is unobstructed
    far = tri[:1].copy()
    for accel in (api.ACCEL_BRUTE_FORCE,):
        check_trace_against_oracle(orc, far, vtx, srf, mic, src, dirs, 4, accel)
    # mic exactly at the source: zero-length direct path (normalize(0) = 0 everywhere)
    check_trace_against_oracle(orc, tri, vtx, srf, src, src, dirs[:64], 3, api.ACCEL_BRUTE_FORCE)


def test_errors_are_reported_not_swallowed():
    tri, vtx, srf = scenes.shoebox()
    rt = api.Raytracer(4, tri, vtx, srf)
    tr = rt.raytrace((7, 2, 4), (2, 1.5, 2), scenes.sphere_directions(64))
    with pytest.raises(api.RayverbError, match="RVB_TRACE_DEBUG"):
        tr.debug()
    with pytest.raises(api.RayverbError):
        tr.render_ir(speakers=[], ft="onepass")                      # speaker model needs a speaker
    with pytest.raises(api.RayverbError, match="exactly one"):
        tr.render_ir()
    big = scenes.displaced_box(20.0, 8.0, 14.0, 8000)[0]
    with pytest.raises(api.RayverbError, match="shared memory"):
        api.Raytracer(4, big, scenes.displaced_box(20.0, 8.0, 14.0, 8000)[1], scenes.hall_surfaces(), accel=api.ACCEL_BRUTE_FORCE)
    rt.close()
