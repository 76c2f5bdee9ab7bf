"""POD layouts shared with the C ABI (include/rayverb_b200.h).

Byte-identical to the reference's rayverb/clstructs.h:13-58: cl_float3 is 16 B
(xyz + pad), cl_float8 is 32 B, Triangle holds u64 indices.
"""
import numpy as np

NUM_IMAGE_SOURCE = 10   # clstructs.h:4
NBANDS = 8              # VolumeType = cl_float8, clstructs.h:13
SPEED_OF_SOUND = 340.0  # clstructs.h:5

TRIANGLE = np.dtype([("surface", "<u8"), ("v0", "<u8"), ("v1", "<u8"), ("v2", "<u8")])
SURFACE = np.dtype([("specular", "<f4", 8), ("diffuse", "<f4", 8)])
IMPULSE = np.dtype([("volume", "<f4", 8), ("position", "<f4", 4), ("time", "<f4"), ("pad", "<f4", 3)])
ATTENUATED = np.dtype([("volume", "<f4", 8), ("time", "<f4"), ("pad", "<f4", 7)])
SPEAKER = np.dtype([("direction", "<f4", 4), ("coefficient", "<f4"), ("pad", "<f4", 3)])

assert TRIANGLE.itemsize == 32 and SURFACE.itemsize == 64 and IMPULSE.itemsize == 64
assert ATTENUATED.itemsize == 64 and SPEAKER.itemsize == 32

# RayverbFiltering::FilterType, filters.h:196-201; JSON names config.h:326-330
FILTER_TYPES = {"sinc": 0, "onepass": 1, "twopass": 2, "linkwitz_riley": 3}
# OutputMode, config.h:342-345
OUTPUT_MODES = {"all": 0, "image_only": 1, "diffuse_only": 2}
