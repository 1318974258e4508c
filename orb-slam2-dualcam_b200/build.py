"""Builds liborbslam2_dualcam_b200.so (the C-ABI library, sm_100a only) in-tree with nvcc.

    python -m orbslam2_dualcam_b200.build            (or __graft_entry__.build())

nvcc cross-compiles without a GPU.  -fmad=false keeps float arithmetic un-contracted so that the device code produces
the same bits as the CPU oracle (cv::fastAtan2, cvRound(x*b+y*a), glibc sincosf restatement).
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "liborbslam2_dualcam_b200.so")
SOURCES = ["orb_common.cu", "orb_extract.cu", "orb_match.cu", "orb_search.cu", "orb_bow.cu", "orb_ba.cu", "orb_gba.cu", "orb_pose.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-cudart", "static"]
# integer / bit-exact float paths: no FMA contraction; the FP64 bundle adjustment (1e-5 parity budget) keeps FMA
FILE_FLAGS = {"orb_ba.cu": [], "orb_gba.cu": [], "orb_pose.cu": []}
DEFAULT_FILE_FLAGS = ["-fmad=false"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(PKG), "include", "orbslam2_dualcam_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    for src in sources():
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
        if force or not os.path.exists(obj) or needs_build():
            cmd = ([_nvcc()] + NVCC_FLAGS + FILE_FLAGS.get(os.path.basename(src), DEFAULT_FILE_FLAGS) +
                   (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])
            subprocess.run(cmd, check=True)
        objs.append(obj)
    subprocess.run([_nvcc()] + NVCC_FLAGS + ["-shared", "-o", LIB] + objs + ["-ldl"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
