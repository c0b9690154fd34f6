"""tests/hostsim -- builds the pose math header for the host (g++) so tests can check it against numpy/scipy."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libpose_math_host.so")


def load():
    src = os.path.join(HERE, "pose_math_host.cpp")
    csrc = os.path.join(HERE, "..", "..", "articulated_pose_b200", "csrc")
    deps = [src] + [os.path.join(csrc, h) for h in ("pose_math.cuh", "lm_fast.cuh", "lm_tick.cuh")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in deps):
        # -ffp-contract=off: keep host rounding comparable with the device build's explicit operations
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", src,
                               "-o", SO])
    lib = ctypes.CDLL(SO)
    lib.hs_pair_scale.restype = ctypes.c_double
    return lib
