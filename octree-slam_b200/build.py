"""Builds libosl_b200.so in-tree with nvcc for sm_100a (no torch dependency in the library)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libosl_b200.so")
HOST = os.path.join(HERE, "host")
HOST_LIB = os.path.join(HERE, "libosl_host.so")
HOST_MAIN = os.path.join(HERE, "osl_main")
SOURCES = ["osl_capi.cu", "osl_integrate.cu", "osl_raycast.cu", "osl_extract.cu", "osl_image.cu", "osl_voxelize.cu", "osl_track.cu", "osl_replica.cu", "osl_voxelize_thin.cu", "osl_sort.cu"]
# the voxeliser is checked bit-exactly against a plain-C restatement: no FMA contraction there
EXTRA_FLAGS = {"osl_voxelize.cu": ["-fmad=false"], "osl_voxelize_thin.cu": ["-fmad=false"]}
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-Xcompiler", "-fno-fast-math"]


def _newest_source():
    t = 0
    for root, _, files in os.walk(CSRC):
        for f in files:
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    t = max(t, os.path.getmtime(os.path.join(HERE, "..", "include", "osl_b200.h")))
    return t


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source():
        build_host()
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for s in SOURCES:
        o = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(o)
        cmd = ([nvcc] + NVCC_FLAGS + EXTRA_FLAGS.get(s, []) + (["-Xptxas", "-v"] if verbose else []) +
               ["-c", os.path.join(CSRC, s), "-o", o])
        procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs +
                          ["-o", LIB, "-lcudart"])
    build_host(force=True)
    return LIB


def _cuda_home():
    nvcc = shutil.which(os.environ.get("NVCC", "nvcc")) or "/usr/local/cuda/bin/nvcc"
    return os.environ.get("CUDA_HOME", os.path.dirname(os.path.dirname(os.path.realpath(nvcc))))


def build_host(force=False):
    """C++ host side above the C ABI (plain g++, no CUDA code): libosl_host.so = the reference's world / rendering /
    sensor interface for the hot path; osl_main = the headless main.cpp:31-62 replay."""
    src = os.path.join(HOST, "osl_host.cpp")
    main_src = os.path.join(HOST, "osl_main.cpp")
    newest = max(os.path.getmtime(src), os.path.getmtime(main_src))
    if (not force and os.path.exists(HOST_LIB) and os.path.exists(HOST_MAIN)
            and min(os.path.getmtime(HOST_LIB), os.path.getmtime(HOST_MAIN)) >= newest):
        return HOST_LIB
    cuda = _cuda_home()
    inc = ["-I", os.path.join(HERE, "..", "include"), "-I", os.path.join(cuda, "include")]
    common = ["g++", "-std=c++17", "-O2", "-fPIC", "-Wall"] + inc
    link = ["-L", HERE, "-L", os.path.join(cuda, "lib64"), "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    subprocess.check_call(common + ["-shared", src, "-o", HOST_LIB] + link + ["-losl_b200", "-lcudart"])
    subprocess.check_call(common + [main_src, "-o", HOST_MAIN] + link + ["-losl_host", "-losl_b200", "-lcudart"])
    return HOST_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
