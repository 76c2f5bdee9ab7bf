"""ctypes binding of librayverb_b200.so (include/rayverb_b200.h).

Mirrors the reference's host classes for the hot path -- Raytracer / SpeakerAttenuator /
HrtfAttenuator (rayverb/rayverb.h:136-308), flattenImpulses / process (rayverb.h:24-45) -- with
the same argument meaning and error behaviour (RuntimeError where the reference throws
std::runtime_error / cl::Error).  There is no CPU fallback: if the library or a B200 is missing,
every compute call raises.
"""
import ctypes as C
import os

import numpy as np

from .types import (ATTENUATED, FILTER_TYPES, IMPULSE, NBANDS, NUM_IMAGE_SOURCE, OUTPUT_MODES, SPEAKER, SURFACE,
                    TRIANGLE)

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB_PATH = os.path.join(PKG_ROOT, "librayverb_b200.so")

TRACE_DEBUG = 1
TRACE_NO_IMAGES = 2
TRACE_SORT_DEDUPE = 4


def TRACE_UNITS(n):
    """BVH scenes, testing: cut every ray's steady-phase reflections into n work units (RVB_TRACE_UNITS)."""
    return (int(n) & 0xff) << 8


def TRACE_MAP_CELLS(n):
    """BVH scenes, testing: cells per triangle of the microphone visibility map (RVB_TRACE_MAP_CELLS)."""
    return (int(n) & 0xff) << 16

ACCEL_AUTO, ACCEL_BRUTE_FORCE, ACCEL_BVH = 0, 1, 2
MODEL_SPEAKER, MODEL_HRTF = 0, 1


class RayverbError(RuntimeError):
    pass


class TraceStats(C.Structure):
    _fields_ = [("nrays", C.c_uint64), ("nreflections", C.c_uint64), ("tri_tests", C.c_uint64),
                ("box_tests", C.c_uint64), ("queries", C.c_uint64), ("bounces", C.c_uint64),
                ("trace_kernel_ms", C.c_float), ("total_ms", C.c_float), ("image_phase_ms", C.c_float),
                ("steady_phase_ms", C.c_float), ("steady_tri_tests", C.c_uint64), ("steady_box_tests", C.c_uint64),
                ("emitted", C.c_uint64), ("micmap_entries", C.c_uint64), ("micmap_cells", C.c_uint32),
                ("reserved_", C.c_uint32)]


class RenderStats(C.Structure):
    _fields_ = [("keys_ms", C.c_float), ("sort_ms", C.c_float), ("reduce_ms", C.c_float), ("filter_ms", C.c_float),
                ("finish_ms", C.c_float), ("total_ms", C.c_float), ("n_impulses", C.c_uint64)]


class RenderConfig(C.Structure):
    _fields_ = [("model", C.c_int), ("speakers", C.c_void_p), ("nspeakers", C.c_size_t), ("hrtf_table", C.c_void_p),
                ("facing", C.c_float * 3), ("up", C.c_float * 3), ("sample_rate", C.c_float), ("filter_type", C.c_int),
                ("lo_cutoff", C.c_float), ("do_normalize", C.c_int), ("volume_scale", C.c_float),
                ("do_trim_tail", C.c_int), ("trim_predelay", C.c_int), ("remove_direct", C.c_int),
                ("output_mode", C.c_int)]


# every symbol include/rayverb_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "rvb_last_error", "rvb_device_count", "rvb_set_device", "rvb_version", "rvb_set_stream", "rvb_launch_count", "rvb_trim_cache",
    "rvb_scene_create", "rvb_scene_destroy", "rvb_scene_accel",
    "rvb_trace_run", "rvb_trace_run_device", "rvb_trace_run_batch", "rvb_trace_destroy", "rvb_trace_diffuse_count", "rvb_trace_copy_diffuse",
    "rvb_trace_image_count", "rvb_trace_copy_images", "rvb_trace_set_images", "rvb_images_merge",
    "rvb_trace_copy_debug", "rvb_trace_get_stats",
    "rvb_attenuate_speaker", "rvb_attenuate_hrtf",
    "rvb_flatten", "rvb_filter", "rvb_filter_sinc", "rvb_process", "rvb_free",
    "rvb_render_ir", "rvb_render_times", "rvb_render_histogram", "rvb_render_finish",
    "rvb_ir_channels", "rvb_ir_length", "rvb_ir_copy", "rvb_ir_band_length", "rvb_ir_copy_bands", "rvb_ir_destroy",
    "rvb_render_get_stats", "rvb_render_keep_bands",
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RayverbError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                               "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.rvb_last_error.restype = C.c_char_p
        for name in EXPORTS:
            fn = getattr(L, name)
            if name not in ("rvb_last_error", "rvb_free", "rvb_launch_count"):
                fn.restype = C.c_int
        L.rvb_free.restype = None
        L.rvb_launch_count.restype = C.c_ulonglong
        L.rvb_set_stream.argtypes = [C.c_void_p, C.c_int]
        L.rvb_free.argtypes = [C.c_void_p]
        vp, sz, u64, f32, i32, u32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_float, C.c_int, C.c_uint
        L.rvb_device_count.argtypes = [vp]
        L.rvb_set_device.argtypes = [i32]
        L.rvb_scene_create.argtypes = [vp, sz, vp, sz, vp, sz, i32, vp]
        L.rvb_scene_destroy.argtypes = [vp]
        L.rvb_scene_accel.argtypes = [vp, vp]
        L.rvb_trace_run.argtypes = [vp, vp, vp, vp, sz, sz, u64, u32, vp]
        L.rvb_trace_run_device.argtypes = [vp, vp, vp, vp, sz, sz, u64, u32, vp]
        L.rvb_trace_run_batch.argtypes = [vp, sz, vp, vp, vp, sz, sz, u32, vp]
        L.rvb_trace_destroy.argtypes = [vp]
        L.rvb_trace_diffuse_count.argtypes = [vp, vp]
        L.rvb_trace_copy_diffuse.argtypes = [vp, vp]
        L.rvb_trace_image_count.argtypes = [vp, i32, vp]
        L.rvb_trace_copy_images.argtypes = [vp, i32, vp, vp, vp]
        L.rvb_trace_set_images.argtypes = [vp, vp, sz]
        L.rvb_images_merge.argtypes = [vp, vp, vp, sz, vp, vp, vp, vp]
        L.rvb_trace_copy_debug.argtypes = [vp, vp, vp]
        L.rvb_trace_get_stats.argtypes = [vp, vp]
        L.rvb_attenuate_speaker.argtypes = [vp, sz, vp, vp, vp]
        L.rvb_attenuate_hrtf.argtypes = [vp, sz, vp, vp, vp, vp, u64, vp]
        L.rvb_flatten.argtypes = [vp, sz, f32, vp, vp]
        L.rvb_filter.argtypes = [i32, vp, sz, sz, f32, f32]
        L.rvb_filter_sinc.argtypes = [vp, sz, sz, f32, f32, vp]
        L.rvb_process.argtypes = [i32, vp, vp, sz, f32, i32, f32, i32, f32, vp, vp]
        L.rvb_render_ir.argtypes = [vp, vp, vp]
        L.rvb_render_times.argtypes = [vp, vp, vp, vp, vp]
        L.rvb_render_histogram.argtypes = [vp, vp, f32, vp, vp]
        L.rvb_render_finish.argtypes = [vp, vp, vp, sz, vp]
        L.rvb_ir_channels.argtypes = [vp, vp]
        L.rvb_ir_length.argtypes = [vp, sz, vp]
        L.rvb_ir_copy.argtypes = [vp, sz, vp]
        L.rvb_ir_band_length.argtypes = [vp, sz, vp]
        L.rvb_ir_copy_bands.argtypes = [vp, sz, vp]
        L.rvb_ir_destroy.argtypes = [vp]
        L.rvb_render_get_stats.argtypes = [vp]
        L.rvb_render_keep_bands.argtypes = [i32]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RayverbError(f"rayverb_b200 error {rc}: {lib().rvb_last_error().decode()}")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f3(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(-1)[:3])


def device_count():
    n = C.c_int(0)
    _check(lib().rvb_device_count(C.byref(n)))
    return n.value


def set_device(dev):
    _check(lib().rvb_set_device(int(dev)))


def set_stream(cuda_stream_ptr):
    """Run this thread's calls on the given CUDA stream (None: the library's own stream)."""
    if cuda_stream_ptr is None:
        _check(lib().rvb_set_stream(None, 0))
    else:
        _check(lib().rvb_set_stream(C.c_void_p(int(cuda_stream_ptr)), 1))


def launch_count():
    return int(lib().rvb_launch_count())


def trim_cache():
    """rvb_trim_cache: give the idle cached device blocks back to the driver."""
    _check(lib().rvb_trim_cache())


def _filter_id(ft):
    return FILTER_TYPES[ft] if isinstance(ft, str) else int(ft)


class Ir:
    """Result of the fused render: per-channel samples (+ filtered bands before mixdown)."""

    def __init__(self, handle):
        self._h = handle
        L = lib()
        n = C.c_size_t(0)
        _check(L.rvb_ir_channels(handle, C.byref(n)))
        self.channels, self.bands = [], []
        for c in range(n.value):
            ln = C.c_size_t(0)
            _check(L.rvb_ir_length(handle, c, C.byref(ln)))
            a = np.zeros(ln.value, dtype=np.float32)
            _check(L.rvb_ir_copy(handle, c, _p(a)))
            self.channels.append(a)
            bl = C.c_size_t(0)
            _check(L.rvb_ir_band_length(handle, c, C.byref(bl)))
            b = np.zeros((NBANDS, bl.value), dtype=np.float32)
            _check(L.rvb_ir_copy_bands(handle, c, _p(b)))
            self.bands.append(b)
        L.rvb_ir_destroy(handle)
        self._h = None


def make_render_config(*, speakers=None, hrtf=None, sr=44100.0, ft="linkwitz_riley", lo_cutoff=45.0,
                       do_normalize=True, do_trim_tail=True, volume_scale=1.0, trim_predelay=False,
                       remove_direct=False, output_mode="all"):
    """Returns (RenderConfig, keepalive).  speakers: list of (direction, coefficient);
    hrtf: dict(table=[2,360,180,8] float32, facing, up)."""
    cfg = RenderConfig()
    keep = []
    if (speakers is None) == (hrtf is None):
        raise RayverbError("attenuation model must contain exactly one of speakers / hrtf")  # config.h:445-455
    if speakers is not None:
        spk = np.zeros(len(speakers), dtype=SPEAKER)
        for i, (d, c) in enumerate(speakers):
            spk["direction"][i, :3] = np.asarray(d, dtype=np.float32)[:3]
            spk["coefficient"][i] = c
        keep.append(spk)
        cfg.model, cfg.speakers, cfg.nspeakers = MODEL_SPEAKER, spk.ctypes.data, len(spk)
    else:
        tab = np.ascontiguousarray(hrtf["table"], dtype=np.float32).reshape(2, 360 * 180, 8)
        keep.append(tab)
        cfg.model, cfg.hrtf_table = MODEL_HRTF, tab.ctypes.data
        cfg.facing = (C.c_float * 3)(*[float(x) for x in _f3(hrtf["facing"])])
        cfg.up = (C.c_float * 3)(*[float(x) for x in _f3(hrtf["up"])])
    cfg.sample_rate, cfg.filter_type, cfg.lo_cutoff = float(sr), _filter_id(ft), float(lo_cutoff)
    cfg.do_normalize, cfg.volume_scale, cfg.do_trim_tail = int(do_normalize), float(volume_scale), int(do_trim_tail)
    cfg.trim_predelay, cfg.remove_direct = int(trim_predelay), int(remove_direct)
    cfg.output_mode = OUTPUT_MODES[output_mode] if isinstance(output_mode, str) else int(output_mode)
    return cfg, keep


class Trace:
    """Device-resident results of one Raytracer::raytrace call."""

    def __init__(self, handle, mic):
        self._h = handle
        self.mic = np.array(mic, dtype=np.float32)

    def close(self):
        if self._h:
            try:
                lib().rvb_trace_destroy(self._h)
            except TypeError:   # interpreter shutdown: module globals are already gone
                pass
            self._h = None

    __del__ = close

    def stats(self):
        s = TraceStats()
        _check(lib().rvb_trace_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}

    def raw_diffuse(self):
        """getRawDiffuse (rayverb.cpp:687-690): IMPULSE[nrays*nreflections], ray-major."""
        n = C.c_size_t(0)
        _check(lib().rvb_trace_diffuse_count(self._h, C.byref(n)))
        out = np.zeros(n.value, dtype=IMPULSE)
        _check(lib().rvb_trace_copy_diffuse(self._h, _p(out)))
        return out

    def raw_images(self, remove_direct=False, with_keys=False):
        """getRawImages (rayverb.cpp:692-706)."""
        n = C.c_size_t(0)
        _check(lib().rvb_trace_image_count(self._h, int(remove_direct), C.byref(n)))
        out = np.zeros(n.value, dtype=IMPULSE)
        keys = np.zeros((n.value, NUM_IMAGE_SOURCE), dtype=np.uint32)
        klen = np.zeros(n.value, dtype=np.int32)
        _check(lib().rvb_trace_copy_images(self._h, int(remove_direct), _p(out), _p(keys), _p(klen)))
        return (out, keys, klen) if with_keys else out

    def all_raw(self, remove_direct=False):
        """getAllRaw (rayverb.cpp:708-714): diffuse followed by images."""
        return np.concatenate([self.raw_diffuse(), self.raw_images(remove_direct)])

    def set_images(self, images):
        img = np.ascontiguousarray(images, dtype=IMPULSE)
        _check(lib().rvb_trace_set_images(self._h, _p(img), len(img)))

    def debug(self):
        st = self.stats()
        hit = np.zeros((st["nrays"], st["nreflections"]), dtype=np.uint32)
        cnt = np.zeros(st["nrays"], dtype=np.uint32)
        _check(lib().rvb_trace_copy_debug(self._h, _p(hit), _p(cnt)))
        return hit, cnt

    def render_ir(self, **kw):
        """Fused attenuate -> fixPredelay -> flattenImpulses -> process on the device."""
        keep_bands = kw.pop("keep_bands", False)
        cfg, keep = make_render_config(**kw)
        h = C.c_void_p()
        lib().rvb_render_keep_bands(int(keep_bands))
        _check(lib().rvb_render_ir(self._h, C.byref(cfg), C.byref(h)))
        del keep
        return Ir(h)

    # ---- staged form for ray shards on several GPUs ------------------------------------------
    def render_times(self, cfg):
        nch = 2 if cfg.model == MODEL_HRTF else max(1, int(cfg.nspeakers))   # one entry per channel is written
        mn = np.zeros(nch, dtype=np.float32)
        mx = np.zeros(nch, dtype=np.float32)
        n = C.c_size_t(0)
        _check(lib().rvb_render_times(self._h, C.byref(cfg), _p(mn), _p(mx), C.byref(n)))
        return mn[:n.value].copy(), mx[:n.value].copy()

    def render_histogram(self, cfg, predelay, lengths, d_hist_ptr):
        ln = np.asarray(lengths, dtype=np.uint64)
        _check(lib().rvb_render_histogram(self._h, C.byref(cfg), C.c_float(predelay), _p(ln), C.c_void_p(d_hist_ptr)))


def render_finish(cfg, d_hist_ptr, lengths, keep_bands=False):
    ln = np.asarray(lengths, dtype=np.uint64)
    h = C.c_void_p()
    lib().rvb_render_keep_bands(int(keep_bands))
    _check(lib().rvb_render_finish(C.byref(cfg), C.c_void_p(d_hist_ptr), _p(ln), len(ln), C.byref(h)))
    return Ir(h)


def render_stats():
    s = RenderStats()
    _check(lib().rvb_render_get_stats(C.byref(s)))
    return {k: getattr(s, k) for k, _ in s._fields_}


def images_merge(images, keys, klen):
    img = np.ascontiguousarray(images, dtype=IMPULSE)
    keys = np.ascontiguousarray(keys, dtype=np.uint32).reshape(-1, NUM_IMAGE_SOURCE)
    klen = np.ascontiguousarray(klen, dtype=np.int32)
    out = np.zeros(max(1, len(img)), dtype=IMPULSE)
    okeys = np.zeros((max(1, len(img)), NUM_IMAGE_SOURCE), dtype=np.uint32)
    olen = np.zeros(max(1, len(img)), dtype=np.int32)
    n = C.c_size_t(0)
    _check(lib().rvb_images_merge(_p(img), _p(keys), _p(klen), len(img), _p(out), _p(okeys), _p(olen), C.byref(n)))
    return out[:n.value].copy(), okeys[:n.value].copy(), olen[:n.value].copy()


class Raytracer:
    """Raytracer (rayverb.h:136-220): scene staged once on the device; raytrace() leaves the
    results on the device behind a Trace."""

    def __init__(self, nreflections, triangles, vertices, surfaces, verbose=False, accel=ACCEL_AUTO):
        self._h = None
        self.trace = None
        self.nreflections = int(nreflections)
        tri = np.ascontiguousarray(triangles, dtype=TRIANGLE)
        vtx = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 4)
        srf = np.ascontiguousarray(surfaces, dtype=SURFACE)
        self.ntriangles = len(tri)
        h = C.c_void_p()
        _check(lib().rvb_scene_create(_p(tri), len(tri), _p(vtx), len(vtx), _p(srf), len(srf), int(accel), C.byref(h)))
        self._h = h
        a = C.c_int(0)
        _check(lib().rvb_scene_accel(h, C.byref(a)))
        self.accel = a.value

    def close(self):
        if self.trace is not None:
            self.trace.close()
            self.trace = None
        if self._h:
            try:
                lib().rvb_scene_destroy(self._h)
            except TypeError:
                pass
            self._h = None

    __del__ = close

    def raytrace(self, micpos, source, directions, verbose=False, *, ray_offset=0, flags=0, device_ptr=None, nrays=None):
        """Raytracer::raytrace (rayverb.cpp:538-685).  `directions` is a host [n,4] float32 array,
        or pass device_ptr/nrays for directions already resident on the GPU."""
        if self.trace is not None:
            self.trace.close()
            self.trace = None
        mic, src = _f3(micpos), _f3(source)
        h = C.c_void_p()
        if device_ptr is not None:
            _check(lib().rvb_trace_run_device(self._h, _p(mic), _p(src), C.c_void_p(device_ptr), int(nrays),
                                              self.nreflections, int(ray_offset), int(flags), C.byref(h)))
        else:
            d = np.ascontiguousarray(directions, dtype=np.float32).reshape(-1, 4)
            _check(lib().rvb_trace_run(self._h, _p(mic), _p(src), _p(d), len(d), self.nreflections, int(ray_offset),
                                       int(flags), C.byref(h)))
        self.trace = Trace(h, mic)
        return self.trace

    def raytrace_batch(self, pairs, *, device_ptr, nrays, flags=0):
        """rvb_trace_run_batch: `pairs` = [(mic, source), ...] over the same device-resident directions,
        traced together.  Returns one Trace per pair (the caller closes them)."""
        mics = np.ascontiguousarray([_f3(m) for m, _ in pairs], dtype=np.float32)
        srcs = np.ascontiguousarray([_f3(s_) for _, s_ in pairs], dtype=np.float32)
        handles = (C.c_void_p * len(pairs))()
        _check(lib().rvb_trace_run_batch(self._h, len(pairs), _p(mics), _p(srcs), C.c_void_p(device_ptr), int(nrays),
                                         self.nreflections, int(flags), handles))
        return [Trace(C.c_void_p(handles[j]), mics[j]) for j in range(len(pairs))]

    def getRawDiffuse(self):
        return self.trace.raw_diffuse()

    def getRawImages(self, removeDirect=False):
        return self.trace.raw_images(removeDirect)

    def getAllRaw(self, removeDirect=False):
        return self.trace.all_raw(removeDirect)


class SpeakerAttenuator:
    """SpeakerAttenuator::attenuate (rayverb.h:282-308): one ATTENUATED array per speaker."""

    def attenuate(self, impulses, mic, speakers):
        imp = np.ascontiguousarray(impulses, dtype=IMPULSE)
        mic = _f3(mic)
        out = []
        for (d, c) in speakers:
            spk = np.zeros(1, dtype=SPEAKER)
            spk["direction"][0, :3] = np.asarray(d, dtype=np.float32)[:3]
            spk["coefficient"][0] = c
            o = np.zeros(len(imp), dtype=ATTENUATED)
            _check(lib().rvb_attenuate_speaker(_p(imp), len(imp), _p(mic), _p(spk), _p(o)))
            out.append(o)
        return out


class HrtfAttenuator:
    """HrtfAttenuator::attenuate (rayverb.h:237-280); the table is injectable like the
    reference's virtual getHrtfData()."""

    def __init__(self, table):
        self.table = np.ascontiguousarray(table, dtype=np.float32).reshape(2, 360 * 180, 8)

    def attenuate(self, impulses, mic, facing, up, channels=(0, 1)):
        imp = np.ascontiguousarray(impulses, dtype=IMPULSE)
        mic, facing, up = _f3(mic), _f3(facing), _f3(up)
        out = []
        for ch in channels:
            o = np.zeros(len(imp), dtype=ATTENUATED)
            tab = np.ascontiguousarray(self.table[ch])
            _check(lib().rvb_attenuate_hrtf(_p(imp), len(imp), _p(mic), _p(tab), _p(facing), _p(up), int(ch), _p(o)))
            out.append(o)
        return out


def flatten_impulses(attenuated, samplerate):
    """flattenImpulses for one channel (rayverb.cpp:48-77) -> float32 [8, L]."""
    a = np.ascontiguousarray(attenuated, dtype=ATTENUATED)
    ptr = C.c_void_p()
    n = C.c_size_t(0)
    _check(lib().rvb_flatten(_p(a), len(a), C.c_float(samplerate), C.byref(ptr), C.byref(n)))
    buf = (C.c_float * (NBANDS * n.value)).from_address(ptr.value)
    out = np.frombuffer(buf, dtype=np.float32).reshape(NBANDS, n.value).copy()
    lib().rvb_free(ptr)
    return out


def filter_bands(ft, data, sr, lo_cutoff):
    """RayverbFiltering::filter (filters.cpp:268-306) on [nch, 8, L]; returns a filtered copy."""
    d = np.array(data, dtype=np.float32, order="C", copy=True)
    assert d.ndim == 3 and d.shape[1] == NBANDS
    if _filter_id(ft) == 0:   # windowed sinc: the data grows by 28 samples
        out = np.zeros((d.shape[0], NBANDS, d.shape[2] + 28), dtype=np.float32)
        _check(lib().rvb_filter_sinc(_p(d), d.shape[0], d.shape[2], C.c_float(sr), C.c_float(lo_cutoff), _p(out)))
        return out
    _check(lib().rvb_filter(_filter_id(ft), _p(d), d.shape[0], d.shape[2], C.c_float(sr), C.c_float(lo_cutoff)))
    return d


def process(ft, flattened, sr, do_normalize=True, lo_cutoff=45.0, do_trim_tail=True, volume_scale=1.0):
    """process() (rayverb.cpp:125-149) on a list of [8, L_c] arrays."""
    data = [np.array(f, dtype=np.float32, order="C", copy=True) for f in flattened]
    lens = np.array([d.shape[1] for d in data], dtype=np.uint64)
    sinc = _filter_id(ft) == 0
    outs = [np.zeros(int(lens[0]) + 28 if sinc else int(l), dtype=np.float32) for l in lens]
    dptr = (C.c_void_p * len(data))(*[d.ctypes.data for d in data])
    optr = (C.c_void_p * len(outs))(*[o.ctypes.data for o in outs])
    out_len = C.c_int64(-1)
    _check(lib().rvb_process(_filter_id(ft), dptr, _p(lens), len(data), C.c_float(sr), int(do_normalize),
                             C.c_float(lo_cutoff), int(do_trim_tail), C.c_float(volume_scale), optr, C.byref(out_len)))
    if out_len.value >= 0:
        n = out_len.value
        outs = [np.concatenate([o[:n], np.zeros(max(0, n - len(o)), np.float32)]) for o in outs]
    return outs
