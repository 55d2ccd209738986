"""Builds libpumipic_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpumipic_b200.so")
OBJ = os.path.join(HERE, "_obj")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -fmad=false: never contract a*b+c, so geometry predicates are bit-identical to the reference's
# CPU arithmetic (DESIGN.md "Parity").  -cudart static: the library loads on a box without a GPU.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-Xcompiler", "-fPIC", "-cudart", "static",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _newer(src, dst, deps):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(d) > t for d in [src] + deps)


def build(verbose=False, force=False, extra_flags=(), lib=None, obj=None):
    """extra_flags/lib/obj let experiments build side-by-side variants (see tools/)."""
    global LIB, OBJ
    if lib:
        LIB = lib
    if obj:
        OBJ = obj
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))
    hdrs = (glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.hpp"))
            + glob.glob(os.path.join(ROOT, "include", "*.h")))
    objs, procs = [], []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s) + ".o")
        objs.append(o)
        if force or _newer(s, o, hdrs):
            cmd = ([NVCC] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else [])
                   + ["-c", s, "-o", o])
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for s, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s:\n%s\n" % (s, out))
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("libpumipic_b200.so: compilation failed")
    if force or procs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
               "-o", LIB] + objs + ["-ldl", "-lz"]
        subprocess.check_call(cmd)
    return LIB


CPP_DRIVERS = ("pseudo_push_and_search", "mirror_api")


def build_cpp_tests():
    """nvcc-compiled drivers that exercise the C++ API mirror (tests/cpp); the binaries travel to
    the GPU box with the snapshot.  Returns the path of the first one (the PIC-loop driver)."""
    out_dir = os.path.join(ROOT, "tests", "cpp", "_bin")
    os.makedirs(out_dir, exist_ok=True)
    hdr = os.path.join(HERE, "cpp", "pumipic_b200.hpp")
    outs = []
    for name in CPP_DRIVERS:
        src = os.path.join(ROOT, "tests", "cpp", name + ".cu")
        out = os.path.join(out_dir, name)
        outs.append(out)
        if _newer(src, out, [hdr, LIB, os.path.join(ROOT, "include", "pumipic_b200.h")]):
            subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O2",
                                   "-std=c++17", "--extended-lambda", "-fmad=false",
                                   "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "cpp"),
                                   src, "-o", out, "-L", HERE, "-lpumipic_b200",
                                   "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../../pumi-pic_b200"])
    return outs[0]


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
