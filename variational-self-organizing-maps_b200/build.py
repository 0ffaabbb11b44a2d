"""Builds libvsom_b200.so (the C-ABI library of include/vsom_b200.h) in-tree with nvcc for sm_100a only."""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "libvsom_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # parity: the reference CPU build has no FMA; never let nvcc contract a*b+c (see csrc/common.cuh)
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
    "-Xptxas", "-v",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(PKG, "host", "*.cpp")) + glob.glob(os.path.join(REPO, "include", "*.hpp")) + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(REPO, "include", "*.h")) + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into lib/libvsom_b200.so.  Cross-compiles without a GPU."""
    if not force and up_to_date():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc_path(), *NVCC_FLAGS, "-I", os.path.join(REPO, "include"), "-I", CSRC, "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {os.path.basename(src)}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    # link the shared libstdc++ by name: this image's g++ would otherwise embed a private static copy
    # (its libstdc++.so symlink dangles), which clashes with the one already loaded into Python.
    stdcxx = ["-Xlinker", "-l:libstdc++.so.6"] if os.path.exists("/usr/lib/x86_64-linux-gnu/libstdc++.so.6") else []
    cmd = [nvcc_path(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", *stdcxx]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0:
        raise RuntimeError(f"link failed:\n{out.stdout}")
    with open(os.path.join(LIB_DIR, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    build_host()
    return LIB


HOST_LIB = os.path.join(LIB_DIR, "libvsom_host.so")
API_DRIVER = os.path.join(REPO, "tests", "cpp", "api_driver_b200")


def build_host() -> str:
    """The drop-in C++ classes (include/SOM.hpp, ...) on top of the C-ABI, and the API driver of tests/cpp."""
    cxx = shutil.which("g++") or "g++"
    eigen = "/usr/include/eigen3" if os.path.exists("/usr/include/eigen3/Eigen/Dense") else os.path.join(REPO, "include", "compat")
    flags = ["-std=c++20", "-O2", "-fPIC", "-ffp-contract=off", "-msse2", "-I", os.path.join(REPO, "include"), "-I", eigen]
    stdcxx = ["-nostdlib++", "-l:libstdc++.so.6"] if os.path.exists("/usr/lib/x86_64-linux-gnu/libstdc++.so.6") else []
    srcs = sorted(glob.glob(os.path.join(PKG, "host", "*.cpp")))
    cmd = [cxx, *flags, "-shared", "-o", HOST_LIB, *srcs, "-L", LIB_DIR, "-lvsom_b200", "-Wl,-rpath,$ORIGIN", *stdcxx]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0:
        raise RuntimeError(f"host library failed to build:\n{out.stdout}")
    drv = os.path.join(REPO, "tests", "cpp", "api_driver.cpp")
    if os.path.exists(drv):
        cmd = [cxx, *flags, drv, "-o", API_DRIVER, "-L", LIB_DIR, "-lvsom_host", "-lvsom_b200", f"-Wl,-rpath,{LIB_DIR}", *stdcxx]
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if out.returncode != 0:
            raise RuntimeError(f"api driver failed to build:\n{out.stdout}")
    return HOST_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
