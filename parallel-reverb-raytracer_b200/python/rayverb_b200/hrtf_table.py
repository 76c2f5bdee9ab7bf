"""HRTF table pipeline: measured head-related impulse responses -> the [2][360][180] table of
8-band gains that HrtfAttenuator reads (rayverb.h:255-257), as a file HrtfAttenuator::loadTable /
RAYVERB_HRTF_TABLE can load (2*360*180*8 float32, ear-major).

A Python 3 restatement of the reference's offline tooling, hrtf_analysis/analyse_hrtf.py:
  decode_filename   :12-26   IRCAM Listen names  IRC_<subject>_<C|R>_R<radius>_T<azimuth>_P<elevation>.wav
  analyse           :111-154 per ear: |rfft|^2 averaged over the bins of the eight bands with edges
                             0, 190, 380, 760, 1520, 3040, 6080, 12160, 20000 Hz (bin = f * nframes / sr in
                             Python 2 INTEGER arithmetic); elevation remapped to (90 + 360 - e) % 360
  interpolate       :28-109  (write_file) bilinear resampling of the sparse (azimuth, elevation) grid to
                             1 degree, corner lookup on azimuth % 360, missing corners = zeros
The reference's own table (rayverb/hrtf.cpp) was generated this way from the IRCAM Listen recordings,
which are not redistributable with it; with those WAV files in a folder,
    python -m rayverb_b200.hrtf_table <folder> table.f32
reproduces its inputs' analysis, and RAYVERB_HRTF_TABLE=table.f32 makes the CLI use it.
"""
import os
import sys
import wave

import numpy as np

FREQUENCY_BOUNDARIES = [0, 190, 380, 760, 1520, 3040, 6080, 12160, 20000]   # analyse_hrtf.py:10


def decode_filename(fname):
    """(radius, azimuth, elevation) from an IRCAM Listen file name (analyse_hrtf.py:12-26)."""
    stem, _ = os.path.splitext(os.path.basename(fname))
    parts = stem.split("_")
    if len(parts) != 6:
        raise RuntimeError("Filename isn't in the IRCAM Listen filename format")
    return int(parts[3][1:]), int(parts[4][1:]), int(parts[5][1:])


def band_energies(signal, sr):
    """Mean |rfft|^2 of one ear's impulse response over the eight bands (analyse_hrtf.py:131-148)."""
    signal = np.asarray(signal, dtype=np.float64)
    n = int(signal.size)
    bounds = [f * n // int(sr) for f in FREQUENCY_BOUNDARIES]        # Python 2 integer division
    fft = np.fft.rfft(signal)
    out = []
    for i, j in zip(bounds[:-1], bounds[1:]):
        band = fft[i:j]
        out.append(float(np.sum(np.abs(band) ** 2) / len(band)))     # an empty band divides by zero there too
    return out


def read_wav(path):
    """[nframes, nchannels] float64 in [-1, 1) and the sample rate (what audiolab's read_frames returns)."""
    with wave.open(path, "rb") as w:
        nch, width, sr, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(n)
    if width == 2:
        a = np.frombuffer(raw, dtype="<i2").astype(np.float64) / 32768.0
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int64)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        a = np.where(v >= 1 << 23, v - (1 << 24), v).astype(np.float64) / float(1 << 23)
    elif width == 4:
        a = np.frombuffer(raw, dtype="<i4").astype(np.float64) / float(1 << 31)
    else:
        raise RuntimeError(f"{path}: unsupported sample width {width}")
    return a.reshape(-1, nch), sr


def analyse_folder(folder):
    """[({'r','a','e'}, [left bands, right bands])] for every file of the folder (analyse_hrtf.py:116-154)."""
    out = []
    for fname in sorted(f for f in os.listdir(folder) if os.path.isfile(os.path.join(folder, f))):
        radius, azimuth, elevation = decode_filename(fname)
        elevation = (90 + 360 - elevation) % 360                     # :128
        data, sr = read_wav(os.path.join(folder, fname))
        if data.shape[1] != 2:
            raise RuntimeError(f"{fname}: expected a stereo (left / right ear) recording")
        out.append(({"r": radius, "a": azimuth, "e": elevation}, [band_energies(data[:, 0], sr), band_energies(data[:, 1], sr)]))
    return out


def interpolate(entries):
    """write_file's resampling (analyse_hrtf.py:28-109) -> float32 [2, 360, 180, 8]."""
    lookup = {}
    for meta, coeffs in entries:
        lookup.setdefault((int(meta["a"]), int(meta["e"])), np.asarray(coeffs, dtype=np.float64))   # get_entry returns the FIRST match
    azimuths = [int(m["a"]) for m, _ in entries]
    elevations = [int(m["e"]) for m, _ in entries]
    zeros = np.zeros((2, 8))

    def entry(a, e):                                                 # :34-39: matches a % 360
        return lookup.get((int(a) % 360, int(e)), zeros)

    out = np.zeros((2, 360, 180, 8), dtype=np.float32)
    for a in range(360):
        a_min, a_max = 0.0, 360.0                                    # :41-51
        for t in azimuths:
            if a < t <= a_max:
                a_max = t
            if a_min < t <= a:
                a_min = t
        for e in range(180):
            e_min, e_max = 0.0, 180.0                                # :56-66
            for t in elevations:
                if e < t <= e_max:
                    e_max = t
                if e_min < t <= e:
                    e_min = t
            c00, c10, c01, c11 = entry(a_min, e_min), entry(a_max, e_min), entry(a_min, e_max), entry(a_max, e_max)
            ar = (a - a_min) / float(a_max - a_min)
            er = (e - e_min) / float(e_max - e_min)
            a0 = c00 + (c10 - c00) * ar                              # :73-80
            a1 = c01 + (c11 - c01) * ar
            out[:, a, e, :] = (a0 + (a1 - a0) * er).astype(np.float32)
    return out


def write_table(path, table):
    """The file HrtfAttenuator::loadTable reads: 2*360*180*8 float32, [ear][azimuth][elevation][band]."""
    t = np.ascontiguousarray(table, dtype="<f4")
    if t.shape != (2, 360, 180, 8):
        raise ValueError("table must be [2, 360, 180, 8]")
    t.tofile(path)


def read_table(path):
    t = np.fromfile(path, dtype="<f4")
    if t.size != 2 * 360 * 180 * 8:
        raise ValueError(f"{path} must hold 2*360*180*8 float32 values")
    return t.reshape(2, 360, 180, 8)


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 2:
        print("usage: python -m rayverb_b200.hrtf_table <folder of IRCAM Listen wav files> <table.f32>", file=sys.stderr)
        return 2
    write_table(argv[1], interpolate(analyse_folder(argv[0])))
    return 0


if __name__ == "__main__":
    sys.exit(main())
