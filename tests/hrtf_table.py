"""Regenerates the reference's synthetic HRTF test table (the blob tests/hrtf.cpp is missing
from the mount): a py3 restatement of hrtf_analysis/generate_test_hrtf_data.py:4-15 feeding
write_file in hrtf_analysis/analyse_hrtf.py:28-109 (15-degree grid whose value is
[azimuth, elevation, 0...], bilinearly resampled to 1 degree; the az=360 corner maps to the
az=0 entry)."""
import numpy as np


def test_table():
    grid = {}
    for a in range(0, 361, 15):
        for e in range(0, 181, 15):
            grid[(a, e)] = np.array([a, e, 0, 0, 0, 0, 0, 0], dtype=np.float64)
    az = sorted({k[0] for k in grid})
    el = sorted({k[1] for k in grid})

    def entry(a, e):                       # analyse_hrtf.py:34-39: matches a % 360
        return grid.get((int(a) % 360, int(e)), np.zeros(8))

    out = np.zeros((2, 360, 180, 8), dtype=np.float32)
    for a in range(360):
        a_min, a_max = 0.0, 360.0          # :41-51
        for t in az:
            if a < t <= a_max:
                a_max = t
            if a_min < t <= a:
                a_min = t
        for e in range(180):
            e_min, e_max = 0.0, 180.0      # :56-66
            for t in el:
                if e < t <= e_max:
                    e_max = t
                if e_min < t <= e:
                    e_min = t
            c00, c10 = entry(a_min, e_min), entry(a_max, e_min)
            c01, c11 = entry(a_min, e_max), entry(a_max, e_max)
            ar = (a - a_min) / float(a_max - a_min)
            er = (e - e_min) / float(e_max - e_min)
            a0 = c00 + (c10 - c00) * ar
            a1 = c01 + (c11 - c01) * ar
            out[:, a, e, :] = (a0 + (a1 - a0) * er).astype(np.float32)
    return out


def smooth_table(seed=7):
    """A smooth synthetic [2,360,180,8] table standing in for the unrecoverable real one
    (rayverb/hrtf.cpp needs the IRCAM Listen set): positive, ear-asymmetric, band-dependent."""
    a = np.deg2rad(np.arange(360, dtype=np.float64))[None, :, None, None]
    e = np.deg2rad(np.arange(180, dtype=np.float64))[None, None, :, None]
    b = np.arange(8, dtype=np.float64)[None, None, None, :]
    ear = np.array([-1.0, 1.0])[:, None, None, None]
    t = 0.55 + 0.35 * ear * np.sin(a) * np.sin(e) * (0.3 + 0.1 * b) + 0.08 * np.cos(2 * a + 0.3 * b) * np.sin(e)
    return np.ascontiguousarray(t, dtype=np.float32)
