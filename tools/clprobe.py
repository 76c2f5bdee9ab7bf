import ctypes, os, sys
L = ctypes.CDLL("libOpenCL.so.1")
n = ctypes.c_uint(0)
rc = L.clGetPlatformIDs(0, None, ctypes.byref(n))
print(os.environ.get("OCL_ICD_VENDORS"), os.environ.get("OCL_ICD_FILENAMES"), "rc", rc, "platforms", n.value)
if rc == 0 and n.value:
    plats = (ctypes.c_void_p * n.value)()
    L.clGetPlatformIDs(n.value, plats, None)
    buf = ctypes.create_string_buffer(256)
    L.clGetPlatformInfo(ctypes.c_void_p(plats[0]), 0x0902, 256, buf, None); print("platform", buf.value)
    L.clGetPlatformInfo(ctypes.c_void_p(plats[0]), 0x0901, 256, buf, None); print("version", buf.value)
    nd = ctypes.c_uint(0)
    rc = L.clGetDeviceIDs(ctypes.c_void_p(plats[0]), 4, 0, None, ctypes.byref(nd)); print("gpu devices rc", rc, nd.value)
