"""The HRTF table pipeline (rayverb_b200/hrtf_table.py): a py3 restatement of the reference's offline
tooling hrtf_analysis/analyse_hrtf.py, checked against the reference's own synthetic test table
(hrtf_analysis/generate_test_hrtf_data.py, whose four probed cells tests/hrtf_tests.cpp:47-50 pins)
and, for the spectral analysis, against the formula evaluated directly on synthetic recordings."""
import os
import wave

import numpy as np
import pytest

from rayverb_b200 import hrtf_table as ht
from hrtf_table import test_table as reference_test_table


def _test_grid():
    """generate_test_hrtf_data.py:4-8."""
    return [({"r": 1, "a": a, "e": e}, [[a, e, 0, 0, 0, 0, 0, 0], [a, e, 0, 0, 0, 0, 0, 0]]) for a in range(0, 361, 15) for e in range(0, 181, 15)]


def test_interpolation_reproduces_the_reference_test_table():
    t = ht.interpolate(_test_grid())
    assert t.shape == (2, 360, 180, 8) and t.dtype == np.float32
    assert np.array_equal(t, reference_test_table())
    # the cells the reference's HrtfTest reads (tests/hrtf_tests.cpp:47-50)
    for a in (180, 0, 90, 270):
        assert list(t[0, a, 90, :2]) == [a, 90] and list(t[1, a, 90, :2]) == [a, 90]
    # the az = 360 corner resolves to the az = 0 entry: band 0 ramps 345 -> 0 above 345 degrees
    assert t[0, 352, 30, 0] == pytest.approx(345 * (360 - 352) / 15.0)


def test_band_energies_formula():
    rng = np.random.default_rng(3)
    sig, sr = rng.standard_normal(8192), 44100
    got = ht.band_energies(sig, sr)
    spec = np.abs(np.fft.rfft(sig)) ** 2
    edges = [f * 8192 // sr for f in ht.FREQUENCY_BOUNDARIES]       # integer bins, like Python 2's `/`
    assert edges[1] == 35 and edges[-1] == 3715
    want = [spec[i:j].mean() for i, j in zip(edges[:-1], edges[1:])]
    np.testing.assert_allclose(got, want, rtol=1e-12)
    # a pure tone lands in exactly one band
    tone = np.sin(2 * np.pi * 1000.0 * np.arange(8192) / sr)
    e = ht.band_energies(tone, sr)
    assert int(np.argmax(e)) == 3 and e[3] > 100 * max(e[:3] + e[4:])


def _write_wav(path, data, sr, width):
    with wave.open(path, "wb") as w:
        w.setnchannels(data.shape[1]); w.setsampwidth(width); w.setframerate(sr)
        if width == 2:
            w.writeframes((np.clip(data, -1, 1 - 2 ** -15) * 32768.0).astype("<i2").tobytes())
        else:
            v = (np.clip(data, -1, 1 - 2 ** -23) * float(1 << 23)).astype(np.int64) & 0xFFFFFF
            b = np.stack([v & 255, (v >> 8) & 255, (v >> 16) & 255], -1).astype(np.uint8)
            w.writeframes(b.tobytes())


def test_folder_to_loadable_table(tmp_path):
    """A folder of IRCAM-named stereo recordings -> table file -> read back: names are decoded, elevations
    remapped by (90 + 360 - e) % 360, each grid cell holds that recording's band energies."""
    rng = np.random.default_rng(5)
    folder = tmp_path / "listen"; folder.mkdir()
    sr, n = 44100, 512
    recs = {}
    for az in range(0, 360, 45):
        for el in (315, 330, 345, 0, 15, 30, 45, 60, 75, 90):       # IRCAM elevations: 315..345 are below the horizon
            data = 0.3 * rng.standard_normal((n, 2)) * np.array([1.0 + az / 360.0, 0.5 + el / 720.0])
            name = f"IRC_1002_C_R0195_T{az:03d}_P{el:03d}.wav"
            _write_wav(str(folder / name), data, sr, 2 if (az // 45) % 2 else 3)
            recs[(az, el)] = ht.read_wav(str(folder / name))[0]
    assert ht.decode_filename("IRC_1002_C_R0195_T045_P330.wav") == (195, 45, 330)
    with pytest.raises(RuntimeError):
        ht.decode_filename("notes.txt")
    out = tmp_path / "table.f32"
    assert ht.main([str(folder), str(out)]) == 0
    assert os.path.getsize(out) == 2 * 360 * 180 * 8 * 4
    t = ht.read_table(str(out))
    for (az, el), data in recs.items():
        e = (90 + 360 - el) % 360
        if az >= 360 or e >= 180:
            continue
        for ear in range(2):
            np.testing.assert_allclose(t[ear, az, e], np.float32(ht.band_energies(data[:, ear], sr)), rtol=1e-6)
    # between grid points the table is the bilinear blend of its four corners
    a, e = 100, 37     # corners: azimuth 90 / 135, elevation 30 / 45
    ar, er = (a - 90) / 45.0, (e - 30) / 15.0
    want = (t[0, 90, 30] * (1 - ar) + t[0, 135, 30] * ar) * (1 - er) + (t[0, 90, 45] * (1 - ar) + t[0, 135, 45] * ar) * er
    np.testing.assert_allclose(t[0, a, e], want, rtol=2e-6)
    with pytest.raises(ValueError):
        ht.write_table(str(tmp_path / "bad.f32"), np.zeros((2, 360, 180, 4)))
