"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls: no GPU here)."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    syms = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(h).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"^\s*(?:const\s+)?(?:int|char|void|size_t)\s*\*?\s*(ancsh_\w+)\s*\(", text, re.M):
            syms.add(m.group(1))
    return syms


def test_header_symbols_exported():
    lib = ctypes.CDLL(os.path.join(ROOT, "articulated_pose_b200", "libancsh_b200.so"))
    syms = _declared_symbols()
    assert len(syms) >= 9, syms
    missing = [s for s in sorted(syms) if not hasattr(lib, s)]
    assert not missing, missing


def test_version_string():
    from articulated_pose_b200 import _lib
    assert b"sm_100a" in _lib.ancsh_version()


def test_struct_sizes_match_header():
    """ctypes mirrors must have the C layout (2 pointers + 5 ints, padded to 8)."""
    from articulated_pose_b200 import _lib
    assert ctypes.sizeof(_lib.Layer) == 48
    assert ctypes.sizeof(_lib.Pred) == 88
    assert ctypes.sizeof(_lib.WsLayout) == 21 * 8
    assert ctypes.sizeof(_lib.Net) == 40 + 22 * 48 + 16


def test_product_does_not_import_oracle():
    """The product package must never route through the oracle (test infrastructure)."""
    for py in glob.glob(os.path.join(ROOT, "articulated_pose_b200", "**", "*.py"), recursive=True):
        src = open(py).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), py
        assert "liboracle" not in src, py


def test_argument_validation_needs_no_gpu():
    """Entry points reject bad sizes / NULL pointers with ANCSH_ERR_INVALID_ARG before touching the device (the reference's
    OP_REQUIRES -> InvalidArgument, tf_sampling.cpp:105), and empty batches are a no-op."""
    from articulated_pose_b200 import _lib
    INVALID = -1
    assert _lib.ancsh_box_iou_3d(-1, 50, None, None, None, None, None, None) == INVALID
    assert _lib.ancsh_box_iou_3d(4, 0, None, None, None, None, None, None) == INVALID
    assert _lib.ancsh_box_iou_3d(4, 50, None, None, None, None, None, None) == INVALID          # NULL boxes
    assert _lib.ancsh_box_iou_3d(0, 50, None, None, None, None, None, None) == _lib.OK          # empty batch
    assert _lib.ancsh_amodal_extent(2, 0, 3, None, None, None, None, None) == INVALID
    assert _lib.ancsh_amodal_extent(0, 1024, 3, None, None, None, None, None) == _lib.OK
    assert _lib.ancsh_joint_vote(2, 1024, 3, 5, 3, None, None, None, None, None, None, 0.2, None, None, None, None) == INVALID  # gn width
    assert _lib.ancsh_joint_vote(2, 1024, 1, 3, 3, None, None, None, None, None, None, 0.2, None, None, None, None) == INVALID  # K < 2
    assert _lib.ancsh_joint_vote(0, 1024, 3, 9, 3, None, None, None, None, None, None, 0.2, None, None, None, None) == _lib.OK
    assert _lib.ancsh_similarity_ransac(2, 512, 0, None, None, None, None, None, None, None, None, None, None, None, None) == INVALID
    assert _lib.ancsh_similarity_ransac(2, 512, 300, None, None, None, None, None, None, None, None, None, None, None, None) == INVALID
    assert _lib.ancsh_similarity_ransac(0, 512, 100, None, None, None, None, None, None, None, None, None, None, None, None) == _lib.OK
    assert _lib.ancsh_umeyama(3, 0, None, None, None, None, None, None, None) == INVALID
    assert _lib.ancsh_fps(2, 0, 4, None, None, None, None) != _lib.OK
