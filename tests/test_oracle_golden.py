"""Pins the CPU oracle (oracle/rayverb_oracle.c) against every golden vector the reference's
own gtests hold for the hot path, and against the reference's own filters.cpp
compiled into oracle/_ref.  CPU only."""
import numpy as np
import pytest

from rayverb_b200 import scenes
from hrtf_table import test_table as make_test_table


def assert_float_eq(a, b):
    """gtest ASSERT_FLOAT_EQ: within 4 ULPs."""
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    ia = a.view(np.int32).astype(np.int64); ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7fffffff), ia); ib = np.where(ib < 0, -(ib & 0x7fffffff), ib)
    assert np.all(np.abs(ia - ib) <= 4), (a, b)


def kat_directions(n):
    d = np.zeros((n, 4), np.float32)
    d[:, 2] = -1                                   # filler {0,0,-1}  (raytrace_tests.cpp:24)
    d[:6, :3] = [(0, 0, -1), (0, 0, 1), (0, -1, 0), (0, 1, 0), (-1, 0, 0), (1, 0, 0)]
    return d


def test_raytracer_impulse_directions(orc):
    """tests/raytrace_tests.h:30-48 (RaytracerTest.ImpulseDirections)."""
    tri, vtx, srf = scenes.diamond_room()
    nrefl = 128
    r = orc.raytrace(tri, vtx, srf, (0, 2, 0), (0, 2, 2), kat_directions(4096 + 64), nrefl)
    d = r["diffuse"].reshape(-1, nrefl)
    first = [(0, 2, -27), (0, 2, 27), (0, 0, 2), (0, 27, 2), (-25, 2, 2), (25, 2, 2)]
    second = [(0, 0, 0), (0, 0, 0), (0, 27, 2), (0, 0, 2), (-25, 2, -2), (25, 2, -2)]
    for i in range(6):
        assert_float_eq(d[i, 0]["position"][:3], first[i])
        assert_float_eq(d[i, 1]["position"][:3], second[i])
    # rays 0/1 leave the mesh after one bounce (zero-filled slots); the others never escape
    assert list(r["refl_count"][:6]) == [1, 1, nrefl, nrefl, nrefl, nrefl]


def _attenuation_inputs(n, seed=1):
    rng = np.random.default_rng(seed)
    imp = np.zeros(n, dtype=orc_dtype_impulse())
    imp["volume"] = 1.0
    imp["position"][:6, :3] = [(-10, 0, 0), (10, 0, 0), (0, -10, 0), (0, 10, 0), (0, 0, -10), (0, 0, 10)]
    imp["time"] = rng.uniform(0, 100, n).astype(np.float32)
    return imp


def orc_dtype_impulse():
    from oracle import oracle
    return oracle.IMPULSE


@pytest.mark.parametrize("coef,expect", [(0.0, [1, 1, 1, 1, 1, 1]), (0.5, [.5, .5, .5, .5, 0, 1]),
                                         (1.0, [0, 0, 0, 0, -1, 1])])
def test_attenuate_speaker(orc, coef, expect):
    """tests/attenuation_tests.h:67-101 (AttenuateSpeaker0/1/2 and Timing)."""
    imp = _attenuation_inputs(1024 * 64)
    out = orc.attenuate_speaker(imp, (0, 0, 0), (0, 0, 1), coef)
    for b in range(1, 8):
        assert_float_eq(out["volume"][:, 0], out["volume"][:, b])
    assert_float_eq(out["volume"][:6, 0], expect)
    if coef == 0.0:   # filler impulses sit AT the mic: needs normalize(0) = 0
        assert_float_eq(out["volume"][:, 0], np.ones(len(out)))
    assert np.array_equal(out["time"], imp["time"])          # ASSERT_EQ, bit exact


@pytest.mark.parametrize("facing,order", [((0, 0, 1), (5, 4, 0, 1)), ((1, 0, 0), (1, 0, 5, 4)),
                                          ((0, 0, -1), (4, 5, 1, 0)), ((-1, 0, 0), (0, 1, 4, 5))])
def test_hrtf_config(orc, facing, order):
    """tests/hrtf_tests.cpp:42-85 (HrtfConfig0..3): [180][90],[0][90],[90][90],[270][90]."""
    table = make_test_table()
    imp = _attenuation_inputs(1000)
    out = orc.attenuate_hrtf(imp, (0, 0, 0), table[0], facing, (0, 1, 0), 0)
    for cell, k in zip([(180, 90), (0, 90), (90, 90), (270, 90)], order):
        assert_float_eq(table[0, cell[0], cell[1]], out["volume"][k])
    assert table[0, 180, 90, 0] == 180 and table[0, 270, 90, 1] == 90 and table[0, 350, 0, 0] == 230.0


@pytest.mark.parametrize("ft", ["onepass", "twopass", "linkwitz_riley"])
def test_filters_match_reference_build(orc, ft):
    """oracle filter restatement == the reference's own filters.cpp (oracle/_ref), bit for bit."""
    if not orc.ref_filters_available():
        pytest.skip("oracle/_ref not built (upstream tree absent at build time)")
    rng = np.random.default_rng(3)
    data = (rng.standard_normal((2, 8, 3001)) * rng.uniform(0, 1, (2, 8, 1))).astype(np.float32)
    data[:, :, 100:] *= np.exp(-np.arange(2901) / 500.0).astype(np.float32)
    for sr, lo in [(44100.0, 45.0), (48000.0, 60.0)]:
        ref = orc.ref_filter(ft, data, sr, lo)
        for c in range(2):
            mine = orc.filter_channel(ft, data[c], sr, lo)
            assert np.array_equal(mine.view(np.uint32), ref[c].view(np.uint32))


def test_flatten_process_semantics(orc):
    """rayverb.cpp:48-77,96-149: rounding half away from zero, per-channel length, trimTail's
    off-by-one (Q12), global normalise."""
    att = np.zeros(5, dtype=orc.ATTENUATED)
    sr = 1000.0
    att["time"] = [0.0005, 0.0015, 0.0025, 0.010, 0.0]     # -> samples 1 (0.5 rounds away), 2 (1.5->2), 3 (2.5->3), 10, 0
    att["volume"] = np.arange(1, 6, dtype=np.float32)[:, None]
    f = orc.flatten(att, sr)
    assert f.shape == (8, 11)
    assert list(f[0, :4]) == [5, 1, 2, 3] and f[0, 10] == 4
    assert orc.find_predelay([att]) == np.float32(0.0005)
    out = orc.process("onepass", [f], 44100.0, do_normalize=True, lo_cutoff=45.0, do_trim_tail=True)
    full = orc.process("onepass", [f], 44100.0, do_normalize=True, lo_cutoff=45.0, do_trim_tail=False)
    assert np.isclose(np.abs(full[0]).max(), 1.0)
    last = np.nonzero(np.abs(full[0]) >= 1e-5)[0].max()
    assert len(out[0]) == last                              # the last loud sample itself is dropped


def test_image_sources_match_shoebox_lattice(orc):
    """Image-source branch (kernel.cpp:379-457) + dedupe (rayverb.cpp:654-676) against the analytic
    image lattice of a shoebox: every emitted image distance is a lattice distance of its order."""
    lx, ly, lz = 10.0, 4.0, 6.0
    tri, vtx, srf = scenes.shoebox(lx, ly, lz)
    src = np.array([2, 1.5, 2.0]); mic = np.array([7, 2, 4.0])
    r = orc.raytrace(tri, vtx, srf, mic, src, scenes.sphere_directions(3000, seed=5), 6)
    img, keys, klen = orc.image_dedupe(r["image"], r["image_idx"])
    assert klen[0] == 1 and np.isclose(img[0]["time"] * 340.0, np.linalg.norm(src - mic), rtol=1e-6)

    def lattice(order):
        d = []
        rng_ = range(-order, order + 1)
        for i in rng_:
            for j in rng_:
                for k in rng_:
                    if abs(i) + abs(j) + abs(k) != order:
                        continue
                    p = [(i + (i % 2)) * L + (-1) ** (i % 2) * m for i, L, m in ((i, lx, mic[0]), (j, ly, mic[1]), (k, lz, mic[2]))]
                    d.append(np.linalg.norm(np.array(p) - src))
        return np.array(d)

    for order in range(1, 5):
        sel = (klen == order + 1) & np.all(keys[:, 1:order + 1] != 0, axis=1)
        dist = img["time"][sel] * 340.0
        lat = lattice(order)
        assert sel.sum() > 0
        assert all(np.min(np.abs(lat - x)) < 1e-3 for x in dist)
    assert (klen == 2).sum() == 6                            # all six first-order images found


def test_windowed_sinc_restatement_equals_fft_round_trip(orc):
    """The reference convolves through FFTW (r2c * r2c -> un-normalised c2r, filters.h:56-80); the
    oracle restates that as FFT_LENGTH x direct convolution.  Emulate the reference's algorithm
    with a double-precision FFT and compare."""
    def fftconv(a, b, n):
        return np.fft.irfft(np.fft.rfft(a, n) * np.fft.rfft(b, n), n) * n
    sr = 44100.0
    edges = [45.0, 175, 350, 700, 1400, 2800, 5600, 11200, 20000]
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((8, 3000)) * np.exp(-np.arange(3000) / 500.0)).astype(np.float32)
    y = orc.filter_channel_sinc(x, sr, 45.0)
    assert y.shape == (8, 3028)
    for b in range(8):
        k = orc.sinc_bandpass_kernel(sr, edges[b], edges[b + 1])
        ref = fftconv(k.astype(np.float64), x[b].astype(np.float64), 3028)
        assert np.linalg.norm(y[b] - ref) <= 1e-6 * np.linalg.norm(ref)
    # the 29-tap kernel itself is an FFT convolution of two 15-tap kernels with length 29 (x29)
    k = orc.sinc_bandpass_kernel(sr, 175.0, 350.0)
    assert k.shape == (29,) and abs(k[14]) == np.abs(k).max()
