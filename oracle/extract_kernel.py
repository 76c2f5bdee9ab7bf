"""Writes oracle/_ref/rayverb_kernel.cl: the reference's OpenCL program text exactly as
KernelLoader::KERNEL_STRING assembles it at run time (rayverb/kernel.cpp:3-627): the two
injected #defines (NUM_IMAGE_SOURCE = 10, SPEED_OF_SOUND = std::to_string(340.0f) = "340.000000",
clstructs.h:4-5) followed by the raw string literal, unmodified.

Run in the build container only (oracle/Makefile `ref` target); the output lives in the
git-ignored oracle/_ref/ and travels to the GPU box, where NVIDIA's OpenCL ICD can compile it
(oracle/cl_reference.py).  The reference sources are never copied into the repository history.
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def main():
    src = open(os.path.join(REF, "rayverb", "kernel.cpp")).read()
    m = re.search(r'R"\((.*)\)"\s*\)\s*;', src, flags=re.S)
    if not m:
        raise SystemExit("kernel string not found")
    body = m.group(1)
    hdr = open(os.path.join(REF, "rayverb", "clstructs.h")).read()
    nimg = re.search(r"#define\s+NUM_IMAGE_SOURCE\s+(\d+)", hdr).group(1)
    sos = re.search(r"#define\s+SPEED_OF_SOUND\s+\(([\d.]+)f\)", hdr).group(1)
    text = "#define NUM_IMAGE_SOURCE %s\n#define SPEED_OF_SOUND %s\n%s" % (nimg, "%f" % float(sos), body)
    os.makedirs(os.path.join(HERE, "_ref"), exist_ok=True)
    out = os.path.join(HERE, "_ref", "rayverb_kernel.cl")
    open(out, "w").write(text)
    print("wrote %s (%d bytes)" % (out, len(text)))


if __name__ == "__main__":
    main()
