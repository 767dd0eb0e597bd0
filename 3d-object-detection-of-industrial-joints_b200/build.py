"""Builds libb200reg.so (the CUDA library behind include/b200reg.h) for sm_100a, in-tree.

nvcc cross-compiles without a GPU.  --fmad=false: PCL/FLANN evaluate their float32 sums as separate
multiplies and adds; parity with the CPU evaluation needs the same sequence on the device.
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb200reg.so")
OBJ = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--fmad=false"] + (["-DB200_GC_TIMING"] if os.environ.get("B200_GC_TIMING") else []) + [
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-ccbin", "/usr/bin/g++",
]


def nvcc():
    for c in ("/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "b200reg.h"),
                                                                 os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    cc = nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        cmd = [cc] + NVCC_FLAGS + extra + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    objs = []
    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        for src, obj, r in ex.map(compile_one, sources()):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed on %s" % src)
            objs.append(obj)
    cmd = [cc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-o", OUT] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return OUT


ROOT = os.path.dirname(HERE)
APPS = os.path.join(ROOT, "apps")
APP_BIN = os.path.join(APPS, "bin")


def build_apps(force=False):
    """C++ host side above the C ABI: the harness programs in apps/ (PCL-style adapters from
    include/pcl_b200/), linked against libb200reg.so with an rpath relative to the binary."""
    so = build()
    os.makedirs(APP_BIN, exist_ok=True)
    hdr = os.path.join(ROOT, "include", "pcl_b200", "pcl_b200.h")
    outs = []
    for f in sorted(os.listdir(APPS)):
        if not f.endswith(".cpp"):
            continue
        src, out = os.path.join(APPS, f), os.path.join(APP_BIN, f[:-4])
        deps = [src, hdr, os.path.join(ROOT, "include", "b200reg.h"), so]
        if force or not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
            cmd = ["g++", "-O2", "-std=c++17", "-pthread", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"), src, "-o", out,
                   "-L" + HERE, "-lb200reg", "-Wl,-rpath,$ORIGIN/../../" + os.path.basename(HERE)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("g++ failed on %s" % src)
        outs.append(out)
    return outs


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_apps(force="--force" in sys.argv))
