"""Runs the reference's OWN OpenCL kernels (the verbatim KERNEL_STRING, oracle/_ref/rayverb_kernel.cl)
on the GPU box through NVIDIA's OpenCL ICD -- a second oracle next to the CPU restatement.

TEST INFRASTRUCTURE ONLY (imported by tests/ and bench.py's reference legs; never by the product).

The image has the Khronos ICD loader (libOpenCL.so.1 from the CUDA toolkit) but no ICD
registration; the driver's libnvidia-opencl.so.1 is reachable with OCL_ICD_FILENAMES.  Nothing is
installed and nothing outside the repository is written.  available() is False when there is no
GPU / no vendor library / no extracted kernel text (e.g. in the CPU-only build container).

The host side below restates how the reference drives its kernels:
  raytrace   rayverb/rayverb.cpp:587-684 (zero-filled buffers, NDRange over the rays, AIR_COEFFICIENT)
  attenuate  rayverb/rayverb.cpp:856-892
  hrtf       rayverb/rayverb.cpp:765-818 (one ear's table flattened a*180+e)
with one simplification: all rays go in ONE NDRange instead of groups of 4096 (work-items are
independent, kernel.cpp:319), and the output buffer is zero-filled like the reference's.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KERNEL_PATH = os.path.join(HERE, "_ref", "rayverb_kernel.cl")
VENDOR_LIBS = ["/usr/lib/libnvidia-opencl.so.1", "/usr/local/nvidia/lib/libnvidia-opencl.so.1",
               "/usr/lib/x86_64-linux-gnu/libnvidia-opencl.so.1"]

TRIANGLE = np.dtype([("surface", "<u8"), ("v0", "<u8"), ("v1", "<u8"), ("v2", "<u8")])
SURFACE = np.dtype([("specular", "<f4", 8), ("diffuse", "<f4", 8)])
IMPULSE = np.dtype([("volume", "<f4", 8), ("position", "<f4", 4), ("time", "<f4"), ("pad", "<f4", 3)])
ATTENUATED = np.dtype([("volume", "<f4", 8), ("time", "<f4"), ("pad", "<f4", 7)])

CL_DEVICE_TYPE_GPU = 4
CL_MEM_READ_WRITE, CL_MEM_COPY_HOST_PTR = 1, 32
CL_PROGRAM_BUILD_LOG = 0x1183
CL_TRUE = 1

_state = {}


def _load():
    if "cl" in _state:
        return _state["cl"]
    _state["cl"] = None
    if not os.path.exists(KERNEL_PATH):
        _state["why"] = "oracle/_ref/rayverb_kernel.cl is missing (built only where /root/reference is mounted)"
        return None
    vendor = next((p for p in VENDOR_LIBS if os.path.exists(p)), None)
    if vendor is None:
        _state["why"] = "no libnvidia-opencl.so.1 on this machine"
        return None
    os.environ.setdefault("OCL_ICD_FILENAMES", vendor)
    try:
        cl = C.CDLL("libOpenCL.so.1")
    except OSError as e:
        _state["why"] = f"no OpenCL loader: {e}"
        return None
    vp = C.c_void_p
    cl.clCreateContext.restype = vp
    cl.clCreateCommandQueue.restype = vp
    cl.clCreateProgramWithSource.restype = vp
    cl.clCreateKernel.restype = vp
    cl.clCreateBuffer.restype = vp
    n = C.c_uint(0)
    if cl.clGetPlatformIDs(0, None, C.byref(n)) != 0 or n.value == 0:
        _state["why"] = "no OpenCL platform"
        return None
    plats = (vp * n.value)()
    cl.clGetPlatformIDs(n.value, plats, None)
    dev = vp()
    nd = C.c_uint(0)
    if cl.clGetDeviceIDs(vp(plats[0]), C.c_ulong(CL_DEVICE_TYPE_GPU), 1, C.byref(dev), C.byref(nd)) != 0 or nd.value == 0:
        _state["why"] = "no OpenCL GPU device"
        return None
    err = C.c_int(0)
    ctx = cl.clCreateContext(None, 1, C.byref(dev), None, None, C.byref(err))
    if err.value:
        _state["why"] = f"clCreateContext failed ({err.value})"
        return None
    q = cl.clCreateCommandQueue(vp(ctx), dev, C.c_ulong(0), C.byref(err))
    src = open(KERNEL_PATH, "rb").read()
    csrc = C.c_char_p(src)
    prog = cl.clCreateProgramWithSource(vp(ctx), 1, C.byref(csrc), None, C.byref(err))
    rc = cl.clBuildProgram(vp(prog), 1, C.byref(dev), None, None, None)     # no options, like cl_program.build(used_devices)
    logsz = C.c_size_t(0)
    cl.clGetProgramBuildInfo(vp(prog), dev, CL_PROGRAM_BUILD_LOG, 0, None, C.byref(logsz))
    log = C.create_string_buffer(max(1, logsz.value))
    cl.clGetProgramBuildInfo(vp(prog), dev, CL_PROGRAM_BUILD_LOG, logsz.value, log, None)
    _state["build_log"] = log.value.decode(errors="replace")
    if rc != 0:
        _state["why"] = f"the reference kernel string does not compile with this OpenCL compiler (rc {rc}): {_state['build_log'][:400]}"
        return None
    _state.update(cl=cl, ctx=ctx, queue=q, prog=prog, dev=dev)
    return cl


def available():
    return _load() is not None


def why_unavailable():
    _load()
    return _state.get("why", "")


def build_log():
    _load()
    return _state.get("build_log", "")


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"OpenCL error {rc} in {what}")


class _Buf:
    def __init__(self, arr=None, nbytes=None):
        cl, err = _state["cl"], C.c_int(0)
        if arr is not None:
            arr = np.ascontiguousarray(arr)
            self.h = cl.clCreateBuffer(C.c_void_p(_state["ctx"]), C.c_ulong(CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR), C.c_size_t(max(arr.nbytes, 1)),
                                       arr.ctypes.data_as(C.c_void_p), C.byref(err))
        else:
            zero = np.zeros(max(nbytes, 1), dtype=np.uint8)      # the reference zero-fills its outputs (rayverb.cpp:601-616)
            self.h = cl.clCreateBuffer(C.c_void_p(_state["ctx"]), C.c_ulong(CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR), C.c_size_t(zero.nbytes),
                                       zero.ctypes.data_as(C.c_void_p), C.byref(err))
        _check(err.value, "clCreateBuffer")

    def read(self, out):
        _check(_state["cl"].clEnqueueReadBuffer(C.c_void_p(_state["queue"]), C.c_void_p(self.h), CL_TRUE, C.c_size_t(0), C.c_size_t(out.nbytes),
                                                out.ctypes.data_as(C.c_void_p), 0, None, None), "clEnqueueReadBuffer")
        return out

    def __del__(self):
        try:
            _state["cl"].clReleaseMemObject(C.c_void_p(self.h))
        except Exception:
            pass


def _run(name, args, global_size):
    """args: list of _Buf | bytes (by-value argument).  Returns the wall time of launch + finish."""
    import time
    cl, err = _state["cl"], C.c_int(0)
    k = cl.clCreateKernel(C.c_void_p(_state["prog"]), name.encode(), C.byref(err))
    _check(err.value, f"clCreateKernel({name})")
    keep = []
    for i, a in enumerate(args):
        if isinstance(a, _Buf):
            h = C.c_void_p(a.h)
            keep.append(h)
            _check(cl.clSetKernelArg(C.c_void_p(k), i, C.c_size_t(8), C.byref(h)), f"arg {i}")
        else:
            b = C.create_string_buffer(a, len(a))
            keep.append(b)
            _check(cl.clSetKernelArg(C.c_void_p(k), i, C.c_size_t(len(a)), b), f"arg {i}")
    gs = C.c_size_t(global_size)
    _check(cl.clFinish(C.c_void_p(_state["queue"])), "clFinish")
    t0 = time.perf_counter()
    _check(cl.clEnqueueNDRangeKernel(C.c_void_p(_state["queue"]), C.c_void_p(k), 1, None, C.byref(gs), None, 0, None, None), f"NDRange({name})")
    _check(cl.clFinish(C.c_void_p(_state["queue"])), "clFinish")
    dt = time.perf_counter() - t0
    cl.clReleaseKernel(C.c_void_p(k))
    return dt


def _f3(v):
    a = np.zeros(4, dtype=np.float32)
    a[:3] = np.asarray(v, dtype=np.float32).reshape(-1)[:3]
    return a.tobytes()


def air_coefficient():
    """(VolumeType) {0.001 * -0.1, ...} narrowed to float (rayverb.cpp:632-641)."""
    return np.array([0.001 * -0.1, 0.001 * -0.2, 0.001 * -0.5, 0.001 * -1.1, 0.001 * -2.7, 0.001 * -9.4, 0.001 * -29.0, 0.001 * -60.0],
                    dtype=np.float64).astype(np.float32)


def raytrace(triangles, vertices, surfaces, mic, source, directions, nreflections, read_back=True):
    """kernel raytrace over all rays.  Returns dict(diffuse[nrays*nrefl], image[nrays*10], image_idx[nrays*10])."""
    assert available(), why_unavailable()
    tri = np.ascontiguousarray(triangles, dtype=TRIANGLE)
    vtx = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 4)
    srf = np.ascontiguousarray(surfaces, dtype=SURFACE)
    dirs = np.ascontiguousarray(directions, dtype=np.float32).reshape(-1, 4)
    n = len(dirs)
    b_dirs, b_tri, b_vtx, b_srf = _Buf(dirs), _Buf(tri), _Buf(vtx), _Buf(srf)
    b_imp, b_img, b_idx = _Buf(nbytes=n * nreflections * 64), _Buf(nbytes=n * 10 * 64), _Buf(nbytes=n * 10 * 8)
    # (directions, position, triangles, numtriangles, vertices, source, surfaces, impulses, image_source,
    #  image_source_index, outputOffset, AIR_COEFFICIENT)   kernel.cpp:304-317 / rayverb.cpp:619-642
    dt = _run("raytrace", [b_dirs, _f3(mic), b_tri, np.uint64(len(tri)).tobytes(), b_vtx, _f3(source), b_srf, b_imp, b_img, b_idx,
                           np.uint64(nreflections).tobytes(), air_coefficient().tobytes()], n)
    if not read_back:
        return dict(kernel_seconds=dt)
    return dict(diffuse=b_imp.read(np.zeros(n * nreflections, dtype=IMPULSE)), image=b_img.read(np.zeros(n * 10, dtype=IMPULSE)),
                image_idx=b_idx.read(np.zeros(n * 10, dtype=np.uint64)), kernel_seconds=dt)


def attenuate_speaker(impulses, mic, speaker_direction, coefficient):
    """kernel attenuate (kernel.cpp:515-535) on a zero-filled output buffer."""
    assert available(), why_unavailable()
    imp = np.ascontiguousarray(impulses, dtype=IMPULSE)
    spk = np.zeros(8, dtype=np.float32)
    spk[:3] = np.asarray(speaker_direction, dtype=np.float32)[:3]
    spk[4] = coefficient
    b_in, b_out = _Buf(imp), _Buf(nbytes=len(imp) * 64)
    _run("attenuate", [_f3(mic), b_in, b_out, spk.tobytes()], len(imp))
    return b_out.read(np.zeros(len(imp), dtype=ATTENUATED))


def attenuate_hrtf(impulses, mic, table_one_ear, facing, up, channel):
    """kernel hrtf (kernel.cpp:586-625); table padded by one row of zeros (elevation index 180)."""
    assert available(), why_unavailable()
    imp = np.ascontiguousarray(impulses, dtype=IMPULSE)
    tab = np.zeros((360 * 180 + 180, 8), dtype=np.float32)
    tab[:360 * 180] = np.asarray(table_one_ear, dtype=np.float32).reshape(360 * 180, 8)
    b_in, b_out, b_tab = _Buf(imp), _Buf(nbytes=len(imp) * 64), _Buf(tab)
    _run("hrtf", [_f3(mic), b_in, b_out, b_tab, _f3(facing), _f3(up), np.uint64(channel).tobytes()], len(imp))
    return b_out.read(np.zeros(len(imp), dtype=ATTENUATED))
