"""Builds libubs_b200.so in-tree with nvcc for sm_100a (no torch headers, no pybind: a plain C-ABI library).

    python universal-beta-splatting_b200/build.py [--force] [--verbose]

The object files go to universal-beta-splatting_b200/build/, the library to
universal-beta-splatting_b200/ubs_b200/lib/libubs_b200.so (git-ignored, shipped to the GPU box by gpurun).
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB_DIR = os.path.join(HERE, "ubs_b200", "lib")
LIB = os.path.join(LIB_DIR, "libubs_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

# --use_fast_math matches the reference build (submodules/gsplat/cuda/_backend.py:93-99): FTZ, approximate
# division / sqrt and MUFU-based powf -- required for bit-identical radii / depths / tile lists.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "--use_fast_math", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/ubs_b200.h"]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src, verbose):
    obj = os.path.join(BUILD, src[:-3] + ".o")
    cmd = [NVCC, "-c", os.path.join(CSRC, src), "-o", obj] + NVCC_FLAGS
    if verbose:
        cmd += ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(BUILD, "stamp")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    if not os.path.exists(NVCC):
        if os.path.exists(LIB):
            return LIB  # GPU box without a need to rebuild
        raise RuntimeError("nvcc not found and %s is missing" % LIB)
    objs = []
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        for obj, log in ex.map(lambda s: _compile(s, verbose), _sources()):
            objs.append(obj)
            if verbose and log:
                sys.stderr.write(log)
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    open(stamp, "w").write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
