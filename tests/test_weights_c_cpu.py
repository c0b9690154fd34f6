"""The C-side weight import (ancsh_weights_pack, csrc/weights_host.cu) against its Python mirror
(articulated_pose_b200/weights.py) on the same TF-named variables: padded f32 matrices, biases, tensor-core images and
their scale exponents, bit for bit (the fc11_1 @ fc2_1 fold is a 128-term f64 sum whose order differs from BLAS: 1 ulp)."""
import ctypes

import numpy as np
import pytest

from articulated_pose_b200 import _lib, weights


def c_pack(w, K, mixed, early):
    names = list(w)
    arrs = [np.ascontiguousarray(w[k], np.float32) for k in names]
    n = len(names)
    c_names = (ctypes.c_char_p * n)(*[k.encode() for k in names])
    c_data = (ctypes.c_void_p * n)(*[a.ctypes.data for a in arrs])
    c_cnt = (ctypes.c_size_t * n)(*[a.size for a in arrs])
    h = ctypes.c_void_p()
    rc = _lib.ancsh_weights_pack(n, c_names, c_data, c_cnt, K, int(mixed), int(early), None, ctypes.byref(h))
    return rc, h, arrs


def views(h):
    cnt = ctypes.c_size_t()
    p = _lib.ancsh_packed_flat(h, ctypes.byref(cnt))
    flat = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), shape=(cnt.value,)).copy()
    p = _lib.ancsh_packed_tc(h, ctypes.byref(cnt))
    tc = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint16)), shape=(cnt.value,)).copy()
    return flat, tc


@pytest.mark.parametrize("K,mixed,early", [(3, True, True), (4, True, True), (2, False, False)])
def test_c_pack_equals_python_pack(K, mixed, early):
    w = weights.synthetic_weights(K, mixed, early, seed=11)
    rc, h, keep = c_pack(w, K, mixed, early)
    assert rc == 0
    try:
        L = weights.pack_network(w, K, mixed, early)
        flat_py, offs = weights.flatten_packed(L)
        tc_py, tc_offs = weights.flatten_tc_images(L)
        flat_c, tc_c = views(h)
        assert flat_c.shape == flat_py.shape and tc_c.shape == tc_py.shape
        for slot, pl in L.items():
            wo, bo, to = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
            dims = (ctypes.c_int * 5)()
            ex = ctypes.c_int()
            assert _lib.ancsh_packed_layer(h, slot.encode(), ctypes.byref(wo), ctypes.byref(bo), ctypes.byref(to), dims,
                                           ctypes.byref(ex)) == 0
            assert (wo.value, bo.value) == offs[slot], slot
            assert list(dims) == [pl.cin, pl.cout, pl.cin_pad, pl.cout_pad, pl.relu], slot
            Wc = flat_c[wo.value:wo.value + pl.W.size].reshape(pl.W.shape)
            bc = flat_c[bo.value:bo.value + pl.b.size]
            if slot == "nocs_heads" and early:
                np.testing.assert_allclose(Wc, pl.W, rtol=3e-7, atol=1e-9)
                np.testing.assert_allclose(bc, pl.b, rtol=3e-7, atol=1e-9)
                continue
            np.testing.assert_array_equal(Wc, pl.W, err_msg=slot)
            np.testing.assert_array_equal(bc, pl.b, err_msg=slot)
            if slot in tc_offs:
                assert to.value == tc_offs[slot] and ex.value == pl.tc_exp, slot
                n_img = (pl.cin_pad + 16) * pl.cout_pad * 2
                np.testing.assert_array_equal(tc_c[to.value:to.value + n_img], tc_py[to.value:to.value + n_img], err_msg=slot)
            else:
                assert to.value == ctypes.c_size_t(-1).value
    finally:
        _lib.ancsh_packed_destroy(h)


def test_c_pack_rejects_missing_or_misshapen_variables():
    w = weights.synthetic_weights(3, True, True, seed=1)
    bad = dict(w)
    del bad["SPFN/est_net/layer2/conv1/bn/gamma"]
    rc, h, _ = c_pack(bad, 3, True, True)
    assert rc == _lib.ERR_INVALID_ARG if hasattr(_lib, "ERR_INVALID_ARG") else rc == -1
    bad = dict(w)
    bad["SPFN/nocs_net/fc2_1/weights"] = bad["SPFN/nocs_net/fc2_1/weights"][..., :5]
    rc, h, _ = c_pack(bad, 3, True, True)
    assert rc == -1
    bad = dict(w)
    bad["SPFN/est_net/fc1/weights"] = np.full_like(bad["SPFN/est_net/fc1/weights"], np.inf)
    rc, h, _ = c_pack(bad, 3, True, True)
    assert rc == -3
