"""ctypes front-end of the CPU oracle (oracle/rayverb_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product path
(parallel-reverb-raytracer_b200/) never imports this module.

The POD dtypes below mirror rayverb/clstructs.h:13-58 of the reference
(Triangle 32 B, Surface 64 B, Impulse 64 B, AttenuatedImpulse 64 B, Speaker 32 B).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librayverb_oracle.so")
REF_FILTERS_PATH = os.path.join(HERE, "_ref", "libref_filters.so")

NUM_IMAGE_SOURCE = 10
NBANDS = 8

TRIANGLE = np.dtype([("surface", "<u8"), ("v0", "<u8"), ("v1", "<u8"), ("v2", "<u8")])
SURFACE = np.dtype([("specular", "<f4", 8), ("diffuse", "<f4", 8)])
IMPULSE = np.dtype([("volume", "<f4", 8), ("position", "<f4", 4), ("time", "<f4"), ("pad", "<f4", 3)])
ATTENUATED = np.dtype([("volume", "<f4", 8), ("time", "<f4"), ("pad", "<f4", 7)])
SPEAKER = np.dtype([("direction", "<f4", 4), ("coefficient", "<f4"), ("pad", "<f4", 3)])
assert TRIANGLE.itemsize == 32 and SURFACE.itemsize == 64 and IMPULSE.itemsize == 64
assert ATTENUATED.itemsize == 64 and SPEAKER.itemsize == 32

FILTER_TYPES = {"sinc": 0, "onepass": 1, "twopass": 2, "linkwitz_riley": 3}


def build(force=False):
    """Compile the oracle (and oracle/_ref when the upstream tree is mounted)."""
    src = os.path.join(HERE, "rayverb_oracle.c")
    stale = (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src)
    if force or stale or (os.path.isdir("/root/reference") and not os.path.exists(REF_FILTERS_PATH)):
        subprocess.run(["make", "-C", HERE, "-s"], check=True, stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp, u64, f32, i32 = C.c_void_p, C.c_uint64, C.c_float, C.c_int
        L.orc_raytrace.restype = u64
        L.orc_raytrace.argtypes = [vp, u64, vp, vp, vp, vp, vp, u64, u64, vp, vp, vp, vp, vp]
        L.orc_image_dedupe.restype = u64
        L.orc_image_dedupe.argtypes = [vp, vp, u64, i32, vp, vp, vp]
        L.orc_attenuate_speaker.restype = None
        L.orc_attenuate_speaker.argtypes = [vp, u64, vp, vp, vp]
        L.orc_attenuate_hrtf.restype = None
        L.orc_attenuate_hrtf.argtypes = [vp, u64, vp, vp, vp, vp, u64, vp]
        L.orc_find_predelay.restype = f32
        L.orc_find_predelay.argtypes = [vp, vp, u64]
        L.orc_fix_predelay.restype = None
        L.orc_fix_predelay.argtypes = [vp, u64, f32]
        L.orc_flatten_length.restype = u64
        L.orc_flatten_length.argtypes = [vp, u64, f32]
        L.orc_flatten.restype = None
        L.orc_flatten.argtypes = [vp, u64, f32, vp, u64]
        L.orc_filter_channel.restype = i32
        L.orc_filter_channel.argtypes = [i32, vp, u64, f32, f32]
        L.orc_process.restype = i32
        L.orc_process.argtypes = [i32, vp, vp, u64, f32, i32, f32, i32, f32, vp, vp]
        L.orc_bandpass_biquad_params.restype = None
        L.orc_bandpass_biquad_params.argtypes = [f32, f32, f32, vp]
        L.orc_linkwitz_riley_params.restype = None
        L.orc_linkwitz_riley_params.argtypes = [f32, f32, f32, vp]
        L.orc_sinc_bandpass_kernel.restype = None
        L.orc_sinc_bandpass_kernel.argtypes = [f32, f32, f32, vp]
        L.orc_filter_channel_sinc.restype = i32
        L.orc_filter_channel_sinc.argtypes = [vp, u64, u64, f32, f32, vp]
        L.orc_air_coefficient.restype = None
        L.orc_air_coefficient.argtypes = [vp]
        L.orc_sphere_point.restype = None
        L.orc_sphere_point.argtypes = [f32, f32, vp]
        L.orc_num_threads.restype = i32
        L.orc_set_num_threads.restype = None
        L.orc_set_num_threads.argtypes = [i32]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f3(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32)[:3])


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    """OpenMP threads of the oracle's ray loop (torchrun sets OMP_NUM_THREADS=1 for its ranks)."""
    lib().orc_set_num_threads(int(n))


def raytrace(triangles, vertices, surfaces, mic, source, directions, nreflections, debug=True):
    """kernel raytrace + the host group loop (kernel.cpp:304-503, rayverb.cpp:587-683).

    Returns dict(diffuse[nrays*nrefl] IMPULSE, image[nrays*10] IMPULSE, image_idx[nrays*10] u64,
    hit_tri[nrays, nrefl] u32 (0xffffffff = no bounce), refl_count[nrays] u32, ntests)."""
    tri = np.ascontiguousarray(triangles, dtype=TRIANGLE)
    vtx = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 4)
    srf = np.ascontiguousarray(surfaces, dtype=SURFACE)
    dirs = np.ascontiguousarray(directions, dtype=np.float32).reshape(-1, 4)
    nrays = dirs.shape[0]
    diffuse = np.zeros(nrays * nreflections, dtype=IMPULSE)
    image = np.zeros(nrays * NUM_IMAGE_SOURCE, dtype=IMPULSE)
    image_idx = np.zeros(nrays * NUM_IMAGE_SOURCE, dtype=np.uint64)
    hit = np.zeros((nrays, nreflections), dtype=np.uint32) if debug else None
    cnt = np.zeros(nrays, dtype=np.uint32) if debug else None
    mic, source = _f3(mic), _f3(source)
    nt = lib().orc_raytrace(_p(tri), len(tri), _p(vtx), _p(srf), _p(mic), _p(source), _p(dirs), nrays,
                            nreflections, _p(diffuse), _p(image), _p(image_idx),
                            _p(hit) if debug else None, _p(cnt) if debug else None)
    return dict(diffuse=diffuse, image=image, image_idx=image_idx, hit_tri=hit, refl_count=cnt, ntests=int(nt))


def image_dedupe(image, image_idx, remove_direct=False):
    """Host dedupe + getRawImages (rayverb.cpp:654-676,692-706)."""
    nrays = len(image) // NUM_IMAGE_SOURCE
    out = np.zeros(max(1, nrays * NUM_IMAGE_SOURCE), dtype=IMPULSE)
    keys = np.zeros((max(1, nrays * NUM_IMAGE_SOURCE), NUM_IMAGE_SOURCE), dtype=np.uint64)
    klen = np.zeros(max(1, nrays * NUM_IMAGE_SOURCE), dtype=np.int32)
    n = lib().orc_image_dedupe(_p(image), _p(image_idx), nrays, int(remove_direct), _p(out), _p(keys), _p(klen))
    return out[:n].copy(), keys[:n].copy(), klen[:n].copy()


def attenuate_speaker(impulses, mic, speaker_direction, coefficient):
    imp = np.ascontiguousarray(impulses, dtype=IMPULSE)
    spk = np.zeros(1, dtype=SPEAKER)
    spk["direction"][0, :3] = np.asarray(speaker_direction, dtype=np.float32)[:3]
    spk["coefficient"][0] = coefficient
    out = np.zeros(len(imp), dtype=ATTENUATED)
    mic = _f3(mic)
    lib().orc_attenuate_speaker(_p(imp), len(imp), _p(mic), _p(spk), _p(out))
    return out


def attenuate_hrtf(impulses, mic, table_one_ear, facing, up, channel):
    imp = np.ascontiguousarray(impulses, dtype=IMPULSE)
    # elevation index 180 reads into the next azimuth row (kernel.cpp:583); for azimuth 359 that
    # is one row past the table (undefined in the reference): defined here as a row of zeros
    tab = np.zeros((360 * 180 + 180, 8), dtype=np.float32)
    tab[:360 * 180] = np.asarray(table_one_ear, dtype=np.float32).reshape(360 * 180, 8)
    out = np.zeros(len(imp), dtype=ATTENUATED)
    mic, facing, up = _f3(mic), _f3(facing), _f3(up)
    lib().orc_attenuate_hrtf(_p(imp), len(imp), _p(mic), _p(tab), _p(facing), _p(up), int(channel), _p(out))
    return out


def find_predelay(channels):
    chans = [np.ascontiguousarray(c, dtype=ATTENUATED) for c in channels]
    ptrs = (C.c_void_p * len(chans))(*[c.ctypes.data for c in chans])
    cnts = np.array([len(c) for c in chans], dtype=np.uint64)
    return float(lib().orc_find_predelay(ptrs, _p(cnts), len(chans)))


def fix_predelay(channels):
    """fixPredelay(T&) (rayverb.h:92-97): in place on a list of ATTENUATED arrays."""
    pd = find_predelay(channels)
    for c in channels:
        lib().orc_fix_predelay(_p(c), len(c), C.c_float(pd))
    return pd


def flatten(attenuated, samplerate):
    """flattenImpulses for one channel -> float32 [8, L] (rayverb.cpp:48-77)."""
    a = np.ascontiguousarray(attenuated, dtype=ATTENUATED)
    n = int(lib().orc_flatten_length(_p(a), len(a), C.c_float(samplerate)))
    out = np.zeros((NBANDS, n), dtype=np.float32)
    lib().orc_flatten(_p(a), len(a), C.c_float(samplerate), _p(out), n)
    return out


SINC_KERNEL_LENGTH = 29


def sinc_bandpass_kernel(sr, lo, hi):
    k = np.zeros(SINC_KERNEL_LENGTH, dtype=np.float32)
    lib().orc_sinc_bandpass_kernel(C.c_float(sr), C.c_float(lo), C.c_float(hi), _p(k))
    return k


def filter_channel_sinc(data, sr, lo_cutoff, first_len=None):
    """BandpassWindowedSinc on one channel [8, L] -> [8, 29 + first_len - 1] (the data grows)."""
    d = np.ascontiguousarray(data, dtype=np.float32)
    first_len = d.shape[1] if first_len is None else int(first_len)
    out = np.zeros((NBANDS, SINC_KERNEL_LENGTH + first_len - 1), dtype=np.float32)
    rc = lib().orc_filter_channel_sinc(_p(d), d.shape[1], first_len, C.c_float(sr), C.c_float(lo_cutoff), _p(out))
    if rc:
        raise ValueError("a channel longer than the first one overflows the reference's FFT buffer")
    return out


def filter_channel(ft, data, sr, lo_cutoff):
    """RayverbFiltering::filter on one channel [8, L], returns a filtered copy."""
    if (FILTER_TYPES[ft] if isinstance(ft, str) else int(ft)) == 0:
        return filter_channel_sinc(data, sr, lo_cutoff)
    d = np.array(data, dtype=np.float32, order="C", copy=True)
    rc = lib().orc_filter_channel(FILTER_TYPES[ft] if isinstance(ft, str) else int(ft), _p(d), d.shape[1],
                                  C.c_float(sr), C.c_float(lo_cutoff))
    if rc:
        raise NotImplementedError("windowed-sinc filter is not part of the oracle")
    return d


def process(ft, flattened, sr, do_normalize=True, lo_cutoff=45.0, do_trim_tail=True, volume_scale=1.0):
    """process() (rayverb.cpp:125-149).  flattened: list of [8, L_c] arrays (copied).
    Returns a list of float32 channels, trimmed/resized like the reference."""
    data = [np.array(f, dtype=np.float32, order="C", copy=True) for f in flattened]
    ftid = FILTER_TYPES[ft] if isinstance(ft, str) else int(ft)
    if ftid == 0:   # windowed sinc: the channels grow to 29 + len(first) - 1; then mixdown / normalize / mul / trimTail
        first = data[0].shape[1]
        filt = [filter_channel_sinc(d, sr, lo_cutoff, first) for d in data]
        outs = []
        for f in filt:   # mixdown (rayverb.cpp:80-92): float adds in band order
            acc = np.zeros(f.shape[1], dtype=np.float32)
            for b in range(NBANDS):
                acc = (acc + f[b]).astype(np.float32)
            outs.append(acc)
        if do_normalize:
            m = np.float32(max(np.abs(o).max() for o in outs))
            fmul = np.float32(1.0 / np.float64(m))
            outs = [(o * fmul).astype(np.float32) for o in outs]
        if volume_scale != 1:
            outs = [(o * np.float32(volume_scale)).astype(np.float32) for o in outs]
        if do_trim_tail:
            n = 0
            for o in outs:
                loud = np.nonzero(np.abs(o) >= np.float32(0.00001))[0]
                n = max(n, int(loud[-1]) if len(loud) else -1)
            outs = [np.concatenate([o[:n], np.zeros(max(0, n - len(o)), np.float32)]) for o in outs]
        return outs
    lens = np.array([d.shape[1] for d in data], dtype=np.uint64)
    outs = [np.zeros(int(l), dtype=np.float32) for l in lens]
    dptr = (C.c_void_p * len(data))(*[d.ctypes.data for d in data])
    optr = (C.c_void_p * len(outs))(*[o.ctypes.data for o in outs])
    out_len = C.c_int64(-1)
    rc = lib().orc_process(FILTER_TYPES[ft] if isinstance(ft, str) else int(ft), dptr, _p(lens), len(data),
                           C.c_float(sr), int(do_normalize), C.c_float(lo_cutoff), int(do_trim_tail),
                           C.c_float(volume_scale), optr, C.byref(out_len))
    if rc:
        raise NotImplementedError("windowed-sinc filter is not part of the oracle")
    if out_len.value >= 0:  # vector::resize semantics: truncate or zero-extend
        n = out_len.value
        outs = [np.concatenate([o[:n], np.zeros(max(0, n - len(o)), np.float32)]) for o in outs]
    return outs


def bandpass_biquad_params(lo, hi, sr):
    c = np.zeros(5, dtype=np.float64)
    lib().orc_bandpass_biquad_params(C.c_float(lo), C.c_float(hi), C.c_float(sr), _p(c))
    return c


def linkwitz_riley_params(lo, hi, sr):
    c = np.zeros(10, dtype=np.float64)
    lib().orc_linkwitz_riley_params(C.c_float(lo), C.c_float(hi), C.c_float(sr), _p(c))
    return c


def air_coefficient():
    a = np.zeros(8, dtype=np.float32)
    lib().orc_air_coefficient(_p(a))
    return a


def render_ir(triangles, vertices, surfaces, mic, source, directions, nreflections, *, speakers=None,
              hrtf=None, sr=44100.0, ft="linkwitz_riley", lo_cutoff=45.0, do_normalize=True,
              do_trim_tail=True, volume_scale=1.0, trim_predelay=False, remove_direct=False,
              output_mode="all", return_bands=False):
    """The CLI pipeline cmd/main.cpp:241-336 end to end on the CPU.

    speakers: list of (direction, coefficient); hrtf: dict(table=[2,360,180,8], facing, up)."""
    tr = raytrace(triangles, vertices, surfaces, mic, source, directions, nreflections, debug=False)
    images, _, _ = image_dedupe(tr["image"], tr["image_idx"], remove_direct)
    if output_mode == "all":
        imp = np.concatenate([tr["diffuse"], images])
    elif output_mode == "image_only":
        imp = images
    else:
        imp = tr["diffuse"]
    if speakers is not None:
        att = [attenuate_speaker(imp, mic, d, c) for (d, c) in speakers]
    else:
        att = [attenuate_hrtf(imp, mic, hrtf["table"][ch], hrtf["facing"], hrtf["up"], ch) for ch in (0, 1)]
    if trim_predelay:
        fix_predelay(att)
    flat = [flatten(a, sr) for a in att]
    out = process(ft, flat, sr, do_normalize, lo_cutoff, do_trim_tail, volume_scale)
    if return_bands:
        if (FILTER_TYPES[ft] if isinstance(ft, str) else int(ft)) == 0:
            bands = [filter_channel_sinc(f, sr, lo_cutoff, flat[0].shape[1]) for f in flat]
        else:
            bands = [filter_channel(ft, f, sr, lo_cutoff) for f in flat]
        return out, flat, bands
    return out


# ---- oracle/_ref: the reference's own filters.cpp ------------------------------------------------
def ref_filters_available():
    return os.path.exists(REF_FILTERS_PATH)


def ref_filter(ft, data, sr, lo_cutoff):
    """RayverbFiltering::filter of the UNMODIFIED reference on [nch, 8, L] (copy)."""
    L = C.CDLL(REF_FILTERS_PATH)
    L.ref_filter.restype = C.c_int
    L.ref_filter.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_float, C.c_float]
    d = np.array(data, dtype=np.float32, order="C", copy=True)
    assert d.ndim == 3 and d.shape[1] == 8
    rc = L.ref_filter(FILTER_TYPES[ft] if isinstance(ft, str) else int(ft), _p(d), d.shape[0], d.shape[2],
                      C.c_float(sr), C.c_float(lo_cutoff))
    if rc:
        raise NotImplementedError
    return d
