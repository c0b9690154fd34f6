import numpy as np

from articulated_pose_b200 import weights as W


def test_variable_inventory_matches_reference_graph():
    # 1,465,187 scalars for K=3 (ANCSH heads)
    shapes = W.variable_shapes(3)
    assert sum(int(np.prod(s)) for s in shapes.values()) == 1465187
    assert shapes["SPFN/est_net/layer2/conv0/weights"] == (1, 1, 131, 128)
    assert shapes["SPFN/est_net/fa_layer1/conv_0/weights"] == (1, 1, 1280, 256)
    assert shapes["SPFN/nocs_net/fc2_1/weights"] == (1, 128, 9)
    assert shapes["SPFN/joint_net/fc4_3/weights"] == (1, 128, 3)
    npcs = W.variable_shapes(3, mixed_pred=False, early_split_nocs=False)
    assert "SPFN/nocs_net/fc11_1/weights" not in npcs and npcs["SPFN/nocs_net/fc2_2/weights"] == (1, 128, 1)


def test_bn_fold_and_permutation():
    w = W.synthetic_weights(3, seed=3)
    L = W.pack_network(w, 3)
    rng = np.random.default_rng(0)
    # layer2/conv0: reference input order [xyz(3), feat(128)], packed order [feat(128), xyz(3)]
    x_ref = rng.normal(size=(5, 131))
    s = "SPFN/est_net/layer2/conv0"
    y = x_ref @ w[s + "/weights"].reshape(131, 128).astype(np.float64) + w[s + "/biases"]
    y = (y - w[s + "/bn/moving_mean"]) / np.sqrt(w[s + "/bn/moving_variance"].astype(np.float64) + 1e-3) \
        * w[s + "/bn/gamma"] + w[s + "/bn/beta"]
    pl = L["sa2[0]"]
    x_pk = np.zeros((5, pl.cin_pad))
    x_pk[:, :128] = x_ref[:, 3:]
    x_pk[:, 128:131] = x_ref[:, :3]
    y2 = x_pk @ pl.W.astype(np.float64) + pl.b
    assert pl.cin_pad == 144 and pl.cout_pad == 128
    np.testing.assert_allclose(y2[:, :128], y, rtol=1e-5, atol=1e-5)
    # folded fc11_1 -> fc2_1
    n = "SPFN/nocs_net/"
    h = rng.normal(size=(4, 128))
    mid = h @ w[n + "fc11_1/weights"].reshape(128, 128).astype(np.float64) + w[n + "fc11_1/biases"]
    ref = mid @ w[n + "fc2_1/weights"].reshape(128, 9).astype(np.float64) + w[n + "fc2_1/biases"]
    ph = L["nocs_heads"]
    got = h @ ph.W.astype(np.float64) + ph.b
    np.testing.assert_allclose(got[:, 3:12], ref, rtol=1e-5, atol=1e-5)
    assert ph.cout == 8 * 3 + 1 and ph.cout_pad == 64


def test_flatten_alignment():
    L = W.pack_network(W.synthetic_weights(4), 4)
    flat, offs = W.flatten_packed(L)
    for slot, (wo, bo) in offs.items():
        assert wo % 64 == 0 and bo % 64 == 0
        np.testing.assert_array_equal(flat[wo:wo + L[slot].W.size], L[slot].W.ravel())


def test_tc_image_layout_and_split():
    rng = np.random.default_rng(0)
    Wm = rng.normal(size=(32, 64)).astype(np.float32)
    img = W.tc_image(Wm)
    assert img.shape == (4, 2, 64, 8) and img.dtype == np.uint16
    hi = img[:, 0].view(np.float16).astype(np.float32).transpose(1, 0, 2).reshape(64, 32)     # [N][K]
    lo = img[:, 1].view(np.float16).astype(np.float32).transpose(1, 0, 2).reshape(64, 32)
    np.testing.assert_allclose(hi + lo / 2048.0, Wm.T, rtol=2 ** -20, atol=1e-7)   # hi + lo * 2^-11 carries ~22 mantissa bits
    assert np.abs(lo).max() <= np.abs(Wm).max() * 2.0                               # the lo piece is stored * 2^11
    flat, offs = W.flatten_tc_images(W.pack_network(W.synthetic_weights(3), 3))
    assert all(o % 128 == 0 for o in offs.values()) and set(offs) == set(W.TC_SLOTS)


def test_tc_image_bias_step():
    """The extra 16-deep k-step after cin_pad carries (fp16(b), fp16(b - fp16(b)), 0 ...): a GEMM whose operand has ones
    in those two columns adds the bias (csrc/net_lean.cu); the first cin_pad/8 slices equal the image without bias."""
    rng = np.random.default_rng(1)
    Wm = rng.normal(size=(32, 64)).astype(np.float32)
    b = rng.normal(size=(64,)).astype(np.float32)
    img = W.tc_image(Wm, b)
    assert img.shape == (6, 2, 64, 8)
    np.testing.assert_array_equal(img[:4], W.tc_image(Wm))
    step = img[4:, 0].view(np.float16).astype(np.float32)           # hi image of the bias step: [2][64][8]
    np.testing.assert_allclose(step[0, :, 0] + step[0, :, 1], b, rtol=2 ** -20, atol=1e-7)
    assert not step[0, :, 2:].any() and not step[1].any()
    assert not img[4:, 1].any()                                      # the lo image of the bias step is empty


def test_tc_image_scaling_keeps_small_weights_exact():
    """Per-layer power-of-two scale of the tensor-core images (the fp16 hi/lo split has a 5-bit exponent):
    hi + lo must reproduce W * 2^s to ~2^-21 of each weight whatever the layer's magnitude; unscaled, weights of 1e-4 keep
    ~12 bits only."""
    rng = np.random.default_rng(5)
    for mag in (1e-5, 1e-4, 1e-3, 0.3, 40.0, 3e3):
        Wm = (rng.standard_normal((32, 64)) * mag).astype(np.float32)
        b = (rng.standard_normal(64) * mag).astype(np.float32)
        s = W.tc_scale_exp(Wm, b)
        top = max(np.abs(Wm).max(), np.abs(b).max()) * 2.0 ** s
        assert 2 ** 13 <= top < 2 ** 14
        img = W.tc_image(Wm, b, s)                                   # [K/8][2][N][8], K = 32 + 16 (bias step)
        assert img.shape == (6, 2, 64, 8)
        hi = img[:, 0].view(np.float16).astype(np.float64).transpose(1, 0, 2).reshape(64, 48)
        lo = img[:, 1].view(np.float16).astype(np.float64).transpose(1, 0, 2).reshape(64, 48)
        rec = (hi + lo / 2048.0)[:, :32].T * 2.0 ** -s              # the lo piece is stored * 2^11 (csrc/tc_common.cuh)
        big = np.abs(Wm) > np.abs(Wm).max() * 2.0 ** -10
        assert np.abs(rec - Wm)[big].max() <= np.abs(Wm).max() * 2.0 ** -21
        np.testing.assert_allclose((hi + lo)[:, 32] * 2.0 ** -s + (hi + lo)[:, 33] * 2.0 ** -s, b, rtol=0, atol=np.abs(b).max() * 2.0 ** -21)
    # even an unscaled image of small weights is accurate now: the lo piece is stored * 2^11
    Wm = (rng.standard_normal((16, 64)) * 1e-4).astype(np.float32)
    img = W.tc_image(Wm, None, 0)
    rec = (img[:, 0].view(np.float16).astype(np.float64) + img[:, 1].view(np.float16).astype(np.float64) / 2048.0).transpose(1, 0, 2).reshape(64, 16).T
    assert np.abs(rec - Wm).max() <= np.abs(Wm).max() * 2.0 ** -20        # ... and so does the 2^11 scale of the lo piece alone
