"""Builds articulated_pose_b200/libancsh_b200.so from csrc/*.cu for sm_100a (in-tree, so the .so travels
with the repository snapshot to the GPU box).  `python -m articulated_pose_b200.build [--force] [-v]`.

Every translation unit is compiled to its own object file (in parallel, only when it or a header changed) and the
objects are linked into the shared library; no relocatable device code is needed (kernels never call across files).
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libancsh_b200.so")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h")) + [__file__]


def _obj_of(src):
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    return _stale(LIB, sources() + _headers())


def _compile(src, verbose):
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", _obj_of(src)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return src, r.returncode, r.stdout


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    hdr = _headers()
    todo = [s for s in sources() if force or _stale(_obj_of(s), [s] + hdr)]
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), todo))
    for src, rc, out in results:
        if verbose and out:
            print("== " + os.path.basename(src))
            print(out)
        if rc != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
    for o in glob.glob(os.path.join(OBJ, "*.o")):      # objects of deleted sources
        if o not in [_obj_of(s) for s in sources()]:
            os.remove(o)
    cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + [_obj_of(s) for s in sources()] + ["-o", LIB]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
