"""oracle/ref_loader.py -- TEST INFRASTRUCTURE ONLY (works only where /root/reference is mounted).

Imports the reference's OWN pose modules unmodified:
    evaluation/parallel_ancsh_pose.py   (ransac, *_estimator, *_verifier, objective_eval)
    lib/d3_utils.py                     (rotate_pts, scale_pts, transform_pts, rotate_points_with_rotvec)
    lib/aligning.py                     (estimateSimilarityUmeyama)
through stub modules for packages absent here (h5py, matplotlib, mpl_toolkits) and two shims that restore the
behaviour of the SciPy version the reference pins (scipy==1.3.1, requirements.txt:151):
    * Rotation.from_dcm / as_dcm were renamed from_matrix / as_matrix in SciPy 1.6;
    * least_squares(method='lm') used x_scale=1.0 (MINPACK mode 2, diag=1); SciPy >= 1.16 defaults to 'jac'.
Used to (a) validate oracle/pose_np.py and (b) mint tests/golden/pose_ref.npz (tests/golden/make_pose_golden.py).
"""
import contextlib
import functools
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("ANCSH_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "evaluation"))


_mods = None


def load():
    """Returns (parallel_ancsh_pose, d3_utils, aligning) reference modules."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REF_ROOT)
    for name in ("h5py", "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d",
                 "matplotlib.patches", "matplotlib.cm", "matplotlib.colors"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.setdefault("__path__", [])
            sys.modules[name] = m
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].use = lambda *a, **k: None
    from scipy.spatial.transform import Rotation
    if not hasattr(Rotation, "from_dcm"):
        Rotation.from_dcm = Rotation.from_matrix
        Rotation.as_dcm = Rotation.as_matrix
    for p in (REF_ROOT, os.path.join(REF_ROOT, "evaluation"), os.path.join(REF_ROOT, "lib")):
        if p not in sys.path:
            sys.path.insert(0, p)
    with open(os.devnull, "w") as dn, contextlib.redirect_stdout(dn):
        import evaluation.parallel_ancsh_pose as pap
        import lib.d3_utils as d3
    from scipy.optimize import least_squares
    pap.least_squares = functools.partial(least_squares, x_scale=1.0)
    pap.print = lambda *a, **k: None          # the reference prints per call (parallel_ancsh_pose.py:21)
    try:
        with open(os.devnull, "w") as dn, contextlib.redirect_stdout(dn):
            stub = types.ModuleType("lib.vis_utils")
            for fn in ("plot3d_pts", "plot_arrows", "plot_arrows_list", "plot_lines", "plot_imgs", "hist_show",
                       "plot2d_img", "visualize_mesh"):
                setattr(stub, fn, lambda *a, **k: None)
            sys.modules.setdefault("lib.vis_utils", stub)
            dstub = types.ModuleType("lib.data_utils")
            for fn in ("get_pickle", "load_pickle", "get_model_pts", "write_pointcloud", "get_urdf", "split_dataset",
                       "get_urdf_mobility", "get_test_group"):
                setattr(dstub, fn, lambda *a, **k: None)
            sys.modules.setdefault("lib.data_utils", dstub)
            import lib.aligning as al
    except Exception:   # aligning needs more of lib/*; only Umeyama is used and tests skip when it is missing
        al = None
    _mods = (pap, d3, al)
    return _mods


@contextlib.contextmanager
def injected_randint(index_stream):
    """Replace np.random.randint (the reference's unseeded global RNG, parallel_ancsh_pose.py:38,110-111) by a
    replay of recorded samples: each call pops the next (3,) int array from `index_stream` (an iterator)."""
    orig = np.random.randint

    def fake(high, size=None, **kw):
        idx = np.asarray(next(index_stream))
        assert idx.shape == (size,) or idx.shape == tuple(np.atleast_1d(size)), (idx.shape, size)
        assert (idx >= 0).all() and (idx < high).all()
        return idx.astype(np.int64)

    np.random.randint = fake
    try:
        yield
    finally:
        np.random.randint = orig
