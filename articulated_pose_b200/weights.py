"""Weight import for the ANCSH network: TF variable names -> BN-folded, padded layer matrices.

The import contract is the TF1 checkpoint's variable names:
  scopes from pointnet_plusplus/utils/pointnet_util.py:128,234 and
  pointnet_plusplus/architectures.py:65,70,75,79,82,86,90; variable names from
  pointnet_plusplus/utils/tf_util.py:164,174 (`weights`, `biases`) and :527-531
  (`bn/{beta,gamma,moving_mean,moving_variance}`); heads from lib/architecture.py:105-120,195-208.

No pretrained checkpoint ships with the reference (README.md:95-105), so `synthetic_weights`
provides seeded random variables of the right shapes for parity tests and benchmarks.
"""
from collections import OrderedDict

import numpy as np

BN_EPS = 1e-3  # tf.contrib.layers.batch_norm default, tf_util.py:527-531


def variable_shapes(n_parts, mixed_pred=True, early_split_nocs=True, prefix="SPFN"):
    """OrderedDict name -> shape of every variable the inference graph reads."""
    K = n_parts
    s = OrderedDict()

    def conv(scope, shape, bn):
        s[scope + "/weights"] = tuple(shape)
        s[scope + "/biases"] = (shape[-1],)
        if bn:
            for v in ("beta", "gamma", "moving_mean", "moving_variance"):
                s[scope + "/bn/" + v] = (shape[-1],)

    e = prefix + "/est_net/"
    for scope, dims in (("layer1", [3, 64, 64, 128]), ("layer2", [131, 128, 128, 256]),
                        ("layer3", [259, 256, 512, 1024])):
        for i in range(3):
            conv("%s%s/conv%d" % (e, scope, i), [1, 1, dims[i], dims[i + 1]], True)
    for scope, dims in (("fa_layer1", [1280, 256, 256]), ("fa_layer2", [384, 256, 128]),
                        ("fa_layer3", [131, 128, 128, 128])):
        for i in range(len(dims) - 1):
            conv("%s%s/conv_%d" % (e, scope, i), [1, 1, dims[i], dims[i + 1]], True)
    conv(e + "fc1", [1, 128, 128], True)

    n = prefix + "/nocs_net/"
    out_dims = [K, 3 * K] + ([K, 3 * K] if mixed_pred else []) + [1]      # lib/architecture.py:98-102
    for i, d in enumerate(out_dims):
        if early_split_nocs and i == 1:
            conv(n + "fc11_1", [1, 128, 128], False)
        conv(n + "fc2_%d" % i, [1, 128, d], False)

    j = prefix + "/joint_net/"
    conv(j + "fc3_0", [1, 128, 128], True)
    conv(j + "fc3_1", [1, 128, 128], True)
    for i, d in enumerate((3, 3, 1, 3)):                                   # n_max_parts=3 default, architecture.py:195
        conv(j + "fc4_%d" % i, [1, 128, d], False)
    return s


def synthetic_weights(n_parts, mixed_pred=True, early_split_nocs=True, seed=7, prefix="SPFN"):
    """Seeded variables: Xavier-uniform W, small biases, BN gamma~U(.5,1.5), beta~N(0,.1),
    mean~N(0,.1), var~U(.5,1.5)."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shape in variable_shapes(n_parts, mixed_pred, early_split_nocs, prefix).items():
        leaf = name.rsplit("/", 1)[-1]
        if leaf == "weights":
            fan_in, fan_out = shape[-2], shape[-1]
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            v = rng.uniform(-lim, lim, size=shape)
        elif leaf == "biases":
            v = rng.normal(0.0, 0.05, size=shape)
        elif leaf == "gamma":
            v = rng.uniform(0.5, 1.5, size=shape)
        elif leaf in ("beta", "moving_mean"):
            v = rng.normal(0.0, 0.1, size=shape)
        elif leaf == "moving_variance":
            v = rng.uniform(0.5, 1.5, size=shape)
        else:
            raise AssertionError(name)
        out[name] = v.astype(np.float32)
    return out


def _fold(weights, scope, bn):
    """conv + bias (+ inference BN) -> (W[cin,cout], b[cout]) in float64."""
    W = np.asarray(weights[scope + "/weights"], np.float64)
    W = W.reshape(-1, W.shape[-1])
    b = np.asarray(weights[scope + "/biases"], np.float64)
    if bn:
        g = np.asarray(weights[scope + "/bn/gamma"], np.float64)
        beta = np.asarray(weights[scope + "/bn/beta"], np.float64)
        mu = np.asarray(weights[scope + "/bn/moving_mean"], np.float64)
        var = np.asarray(weights[scope + "/bn/moving_variance"], np.float64)
        s = g / np.sqrt(var + BN_EPS)
        W = W * s[None, :]
        b = (b - mu) * s + beta
    return W, b


def _pad16(c):
    return (c + 15) // 16 * 16


def _pad_out(c):
    return 64 if c <= 64 else (c + 127) // 128 * 128


class PackedLayer:
    def __init__(self, W, b, relu):
        cin, cout = W.shape
        self.cin, self.cout, self.relu = cin, cout, int(relu)
        self.cin_pad, self.cout_pad = _pad16(cin), _pad_out(cout)
        self.W = np.zeros((self.cin_pad, self.cout_pad), np.float32)
        self.W[:cin, :cout] = W
        self.b = np.zeros((self.cout_pad,), np.float32)
        self.b[:cout] = b
        self.tc_exp = 0      # power-of-two scale of the tensor-core image (set by flatten_tc_images)


def pack_network(weights, n_parts, mixed_pred=True, early_split_nocs=True, prefix="SPFN"):
    """Returns OrderedDict slot -> PackedLayer, slots named after the fields of ancsh_net_t
    (include/ancsh_b200.h).  First-layer rows are permuted from the reference's [xyz, features]
    (pointnet_util.py:57,84) to the kernels' [features, xyz]."""
    K = n_parts
    e = prefix + "/est_net/"
    L = OrderedDict()

    def xyz_last(W):
        return np.concatenate([W[3:], W[:3]], axis=0)

    for lvl, scope in ((1, "layer1"), (2, "layer2"), (3, "layer3")):
        for i in range(3):
            W, b = _fold(weights, "%s%s/conv%d" % (e, scope, i), True)
            if i == 0:
                W = xyz_last(W)
            L["sa%d[%d]" % (lvl, i)] = PackedLayer(W, b, True)

    W, b = _fold(weights, e + "fa_layer1/conv_0", True)
    n_glob = L["sa3[2]"].cout                                   # interpolated part = global feature (1024)
    L["fp1_global"] = PackedLayer(W[:n_glob], b, False)
    L["fp1[0]"] = PackedLayer(W[n_glob:], np.zeros_like(b), True)   # bias comes per cloud from fp1_global
    W, b = _fold(weights, e + "fa_layer1/conv_1", True)
    L["fp1[1]"] = PackedLayer(W, b, True)
    for i in range(2):
        W, b = _fold(weights, e + "fa_layer2/conv_%d" % i, True)
        L["fp2[%d]" % i] = PackedLayer(W, b, True)
    for i in range(3):
        W, b = _fold(weights, e + "fa_layer3/conv_%d" % i, True)
        L["fp3[%d]" % i] = PackedLayer(W, b, True)
    W, b = _fold(weights, e + "fc1", True)
    L["fc1"] = PackedLayer(W, b, True)

    n = prefix + "/nocs_net/"
    out_dims = [K, 3 * K] + ([K, 3 * K] if mixed_pred else []) + [1]
    Ws, bs = [], []
    for i, _d in enumerate(out_dims):
        W, b = _fold(weights, n + "fc2_%d" % i, False)
        if early_split_nocs and i == 1:
            # fc11_1 is linear (activation_fn=None, lib/architecture.py:112): fold the pair into one matrix
            W1, b1 = _fold(weights, n + "fc11_1", False)
            b = b1 @ W + b
            W = W1 @ W
        Ws.append(W)
        bs.append(b)
    L["nocs_heads"] = PackedLayer(np.concatenate(Ws, axis=1), np.concatenate(bs), False)

    j = prefix + "/joint_net/"
    for i in range(2):
        W, b = _fold(weights, j + "fc3_%d" % i, True)
        L["fc3[%d]" % i] = PackedLayer(W, b, True)
    Ws, bs = zip(*[_fold(weights, j + "fc4_%d" % i, False) for i in range(4)])
    L["joint_heads"] = PackedLayer(np.concatenate(Ws, axis=1), np.concatenate(bs), False)
    return L


def f32_to_bf16_bits(x):
    """round-to-nearest-even f32 -> bf16 bit patterns (uint16)"""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    return ((u + (((u >> 16) & 1) + 0x7FFF)) >> 16).astype(np.uint16)


def bf16_bits_to_f32(b):
    return (b.astype(np.uint32) << 16).view(np.float32)


def tc_scale_exp(W, b=None):
    """Power-of-two scale exponent s of a layer's tensor-core image: W * 2^s (and b * 2^s) put the largest magnitude into
    [2^13, 2^14).  fp16 pieces are exact to 11 bits only while normal (>= 6.1e-5): unscaled BN-folded weights of ~1e-3
    lose the lo piece to subnormals (measured 1.7e-5 relative output error, 1.7e-4 at 1e-4); scaled, hi + lo keeps ~22 bits
    for every weight down to 2^-17 of the layer's largest.  The kernels undo the scale on the f32 accumulators
    (ancsh_layer_t::tc_descale = 2^-s, exact)."""
    m = float(np.abs(np.asarray(W, np.float32)).max(initial=0.0))
    if b is not None:
        m = max(m, float(np.abs(np.asarray(b, np.float32)).max(initial=0.0)))
    if not np.isfinite(m):
        raise ValueError("non-finite weights")
    if m == 0.0:
        return 0
    return int(np.clip(13 - np.floor(np.log2(m)), -100, 100))


def tc_image(W, b=None, scale_exp=0):
    """Tensor-core operand image of a padded weight matrix W [cin_pad][cout_pad] (f32), scaled by 2^scale_exp: W^T split into
    hi = fp16(w), lo = fp16(w - hi) (round to nearest even), tiled as [K/8][2][cout_pad][8] fp16 bit patterns
    (csrc/tc_common.cuh).  hi + lo carries ~22 significant bits of w.
    With a bias b [cout_pad], K = cin_pad + 16: the extra k-step holds the rows (fp16(b), fp16(b - fp16(b)), 0 ...), so
    that a GEMM whose operand has ones in those two columns adds the bias (csrc/net_lean.cu); kernels that use
    K = cin_pad never read it."""
    W = np.ldexp(np.asarray(W, np.float32), int(scale_exp)).astype(np.float32)
    if b is not None:
        b = np.ldexp(np.asarray(b, np.float32), int(scale_exp)).astype(np.float32)
        if np.abs(b).max(initial=0.0) >= 65504:
            raise ValueError("bias exceeds the fp16 range of the tensor-core path")
        b_hi = b.astype(np.float16).astype(np.float32)
        b_lo = (b - b_hi).astype(np.float16).astype(np.float32)      # a K row of the hi image (times the ones slab): NOT scaled
        extra = np.zeros((16, W.shape[1]), np.float32)
        extra[0], extra[1] = b_hi, b_lo
        W = np.concatenate([np.asarray(W, np.float32), extra], axis=0)
    K, N = W.shape
    Wt = np.ascontiguousarray(W.T, np.float32)                       # [N][K]
    if np.abs(Wt).max() >= 65504:
        raise ValueError("weights exceed the fp16 range of the tensor-core path")
    hi16 = Wt.astype(np.float16)
    lo16 = ((Wt - hi16.astype(np.float32)) * np.float32(2048.0)).astype(np.float16)      # lo piece stored * 2^11 (tc_common.cuh)
    hi, lo = hi16.view(np.uint16), lo16.view(np.uint16)
    img = np.stack([hi.reshape(N, K // 8, 8), lo.reshape(N, K // 8, 8)], axis=0)   # [2][N][K/8][8]
    return np.ascontiguousarray(img.transpose(2, 0, 1, 3))          # [K/8][2][N][8]


TC_SLOTS = ("sa1[0]", "sa1[1]", "sa1[2]", "sa2[0]", "sa2[1]", "sa2[2]", "sa3[0]", "sa3[1]", "sa3[2]", "fp1[0]", "fp1[1]", "fp2[0]", "fp2[1]", "fp3[0]", "fp3[1]", "fp3[2]",
            "fc1", "nocs_heads", "fc3[0]", "fc3[1]", "joint_heads")


def flatten_tc_images(layers):
    """uint16 buffer with the tensor-core images of the layers that run on tcgen05, 256-byte aligned offsets.
    Sets layers[slot].tc_exp (the power-of-two scale of each image, tc_scale_exp)."""
    off, offs, parts = 0, {}, []
    for slot in TC_SLOTS:
        layers[slot].tc_exp = tc_scale_exp(layers[slot].W, layers[slot].b)
        img = tc_image(layers[slot].W, layers[slot].b, layers[slot].tc_exp).ravel()
        offs[slot] = off
        parts.append((off, img))
        off = (off + img.size + 127) // 128 * 128
    flat = np.zeros((off,), np.uint16)
    for o, img in parts:
        flat[o:o + img.size] = img
    return flat, offs


def flatten_packed(layers):
    """Lay all W/b arrays out in one f32 buffer with 256-byte aligned offsets.
    Returns (flat float32 array, {slot: (w_offset_floats, b_offset_floats)})."""
    off = 0
    offs = {}
    for slot, pl in layers.items():
        wo = off
        off = (off + pl.W.size + 63) // 64 * 64
        bo = off
        off = (off + pl.b.size + 63) // 64 * 64
        offs[slot] = (wo, bo)
    flat = np.zeros((off,), np.float32)
    for slot, pl in layers.items():
        wo, bo = offs[slot]
        flat[wo:wo + pl.W.size] = pl.W.ravel()
        flat[bo:bo + pl.b.size] = pl.b
    return flat, offs


def fit_heads(weights, features, cls_gt, nocs_gt, n_parts, early_split_nocs=True, prefix="SPFN", ridge=1e-6):
    """Synthetic-but-meaningful heads: keep the random trunk, fit the linear segmentation and NOCS heads by ridge
    regression of the trunk feature `net` (AncshNet.features) onto ground truth of a few synthetic clouds.

    No checkpoint ships with the reference (README.md:95-105) and purely random heads put every point into one
    part, which would leave the pose stage without its real workload; a random-feature trunk with fitted linear
    read-outs gives network outputs with realistic part partitions.  Returns a new weights dict.

    features (M,128), cls_gt (M,), nocs_gt (M,3) part-normalised coordinates of each point's own part.
    """
    F = np.asarray(features, np.float64).reshape(-1, 128)
    cls_gt = np.asarray(cls_gt).reshape(-1).astype(int)
    nocs_gt = np.asarray(nocs_gt, np.float64).reshape(-1, 3)
    A = np.hstack([F, np.ones((F.shape[0], 1))])

    def solve(Asub, T):
        G = Asub.T @ Asub + ridge * Asub.shape[0] * np.eye(Asub.shape[1])
        X = np.linalg.solve(G, Asub.T @ T)
        return X[:-1], X[-1]

    out = OrderedDict((k, np.array(v, copy=True)) for k, v in weights.items())
    n = prefix + "/nocs_net/"
    T = np.full((F.shape[0], n_parts), -3.0)
    T[np.arange(F.shape[0]), cls_gt] = 3.0
    W, b = solve(A, T)
    out[n + "fc2_0/weights"] = W.reshape(1, 128, n_parts).astype(np.float32)
    out[n + "fc2_0/biases"] = b.astype(np.float32)
    Wn = np.zeros((128, 3 * n_parts))
    bn = np.zeros(3 * n_parts)
    y = np.clip(nocs_gt, 0.02, 0.98)
    logit = np.log(y / (1.0 - y))
    for j in range(n_parts):
        m = cls_gt == j
        if m.sum() > 130:
            Wn[:, 3 * j:3 * j + 3], bn[3 * j:3 * j + 3] = solve(A[m], logit[m])
    out[n + "fc2_1/weights"] = Wn.reshape(1, 128, 3 * n_parts).astype(np.float32)
    out[n + "fc2_1/biases"] = bn.astype(np.float32)
    if early_split_nocs and (n + "fc11_1/weights") in out:
        out[n + "fc11_1/weights"] = np.eye(128, dtype=np.float32).reshape(1, 128, 128)
        out[n + "fc11_1/biases"] = np.zeros(128, np.float32)
    return out
