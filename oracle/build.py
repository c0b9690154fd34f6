"""oracle/build.py -- TEST INFRASTRUCTURE ONLY: builds the CPU oracle and, when the reference
checkout is present (this container), the reference's own native code into oracle/_ref/.

  build_oracle()  gcc  oracle/pnpp_ref.c            -> oracle/liboracle_pnpp.so
  build_ref()     nvcc /root/reference/.../tf_sampling_g.cu + tf_grouping_g.cu + our extern "C" shim
                                                    -> oracle/_ref/libref_tfops.so   (GPU-side oracle
                                                       and the incumbent kernels bench.py can time)
                  g++  the TF-free loops threenn_cpu / threeinterpolate_cpu, extracted at build time
                       from /root/reference/.../tf_interpolate.cpp (the file itself needs TF headers)
                                                    -> oracle/_ref/libref_interp.so

Nothing under /root/reference is copied into the repository: oracle/_ref/ is git-ignored (build
products only) but NOT gpurun-ignored, so the .so files travel to the GPU box, where
/root/reference does not exist.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("ANCSH_REFERENCE_ROOT", "/root/reference")
REF_DIR = os.path.join(HERE, "_ref")
TFOPS = os.path.join(REF_ROOT, "pointnet_plusplus", "utils", "tf_ops")

ORACLE_SO = os.path.join(HERE, "liboracle_pnpp.so")
REF_TFOPS_SO = os.path.join(REF_DIR, "libref_tfops.so")
REF_INTERP_SO = os.path.join(REF_DIR, "libref_interp.so")


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_oracle(force=False):
    src = os.path.join(HERE, "pnpp_ref.c")
    if not force and _newer(ORACLE_SO, [src]):
        return ORACLE_SO
    # -ffp-contract=off: see the header of pnpp_ref.c.  -mfma only makes fmaf() a single instruction;
    # with contraction off no a*b+c is fused behind our back.
    _run(["gcc", "-O2", "-ftree-vectorize", "-mavx2", "-mfma", "-ffp-contract=off", "-fopenmp", "-fPIC",
          "-shared", "-o", ORACLE_SO, src, "-lm"])
    return ORACLE_SO


def _extract_function(text, name):
    m = re.search(r"^void\s+%s\s*\(" % re.escape(name), text, re.M)
    if not m:
        raise RuntimeError("function %s not found in reference source" % name)
    i = text.index("{", m.end())
    depth, j = 0, i
    while True:
        if text[j] == "{":
            depth += 1
        elif text[j] == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    return text[m.start(): j + 1]


def build_ref(force=False):
    """Returns the list of built artefacts; [] when the reference checkout is absent."""
    if not os.path.isdir(TFOPS):
        return []
    os.makedirs(REF_DIR, exist_ok=True)
    built = []
    samp = os.path.join(TFOPS, "sampling", "tf_sampling_g.cu")
    grp = os.path.join(TFOPS, "grouping", "tf_grouping_g.cu")
    shim = os.path.join(HERE, "ref_tfops_shim.cu")
    if force or not _newer(REF_TFOPS_SO, [samp, grp, shim]):
        # the reference builds with `nvcc -c -O2` and no -arch (tf_sampling_compile.sh:4); we keep -O2
        # and only add the sm_100a target.  Default fp contraction (fmad=true) as in the reference.
        _run(["nvcc", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared",
              "-o", REF_TFOPS_SO, samp, grp, shim])
    built.append(REF_TFOPS_SO)

    interp = os.path.join(TFOPS, "3d_interpolation", "tf_interpolate.cpp")
    if force or not _newer(REF_INTERP_SO, [interp, __file__]):
        text = open(interp).read()
        gen = os.path.join(REF_DIR, "ref_interp_extract.cpp")
        with open(gen, "w") as f:
            f.write("// GENERATED at build time from %s (git-ignored build intermediate)\n" % interp)
            f.write("#include <cstdio>\n#include <cmath>\n#include <algorithm>\nusing namespace std;\n")
            f.write(_extract_function(text, "threenn_cpu") + "\n")
            f.write(_extract_function(text, "threeinterpolate_cpu") + "\n")
            f.write('extern "C" void ref_three_nn(int b,int n,int m,const float*a,const float*c,float*d,int*i)'
                    "{threenn_cpu(b,n,m,a,c,d,i);}\n")
            f.write('extern "C" void ref_three_interpolate(int b,int m,int c,int n,const float*p,const int*i,'
                    "const float*w,float*o){threeinterpolate_cpu(b,m,c,n,p,i,w,o);}\n")
        # reference flags: g++ -std=c++11 -O2 (tf_interpolate_compile.sh); no -march, no fast-math
        _run(["g++", "-std=c++11", "-O2", "-fPIC", "-shared", "-o", REF_INTERP_SO, gen])
    built.append(REF_INTERP_SO)
    return built


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv))
    for p in build_ref(force="--force" in sys.argv):
        print(p)
