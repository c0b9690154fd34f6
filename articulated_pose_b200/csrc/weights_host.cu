// weights_host.cu -- C side of the weight import (ancsh_weights_pack / ancsh_net_create), so that a host that
// is not Python can build an ancsh_net_t straight from the TF1 checkpoint's variables.
//
// Contract = the checkpoint's variable names: scopes from pointnet_plusplus/utils/pointnet_util.py:128,234 and
// pointnet_plusplus/architectures.py:65-90, variable names `weights` / `biases` (tf_util.py:164,174) and
// `bn/{beta,gamma,moving_mean,moving_variance}` (tf_util.py:527-531), heads from lib/architecture.py:105-120,195-208.
// What happens to them (same steps, same order of arithmetic as articulated_pose_b200/weights.py, which stays the
// Python mirror; tests/test_weights_c_cpu.py compares the two bit for bit on the f32 buffers):
//   * conv + bias + inference batch norm (eps 1e-3) folded in f64:  s = gamma / sqrt(var + eps), W *= s, b = (b - mean) * s + beta
//   * first-layer rows permuted from the reference's [xyz, features] (pointnet_util.py:57,84) to [features, xyz]
//   * fa_layer1/conv_0 split into the rows of the broadcast global feature (fp1_global) and the skip rows (fp1[0])
//   * head stacks concatenated column-wise, fc11_1 (linear) folded into fc2_1 (lib/architecture.py:112)
//   * zero padding to cin_pad % 16 == 0, cout_pad = 64 or a multiple of 128
//   * tensor-core images: W^T * 2^s split into hi = fp16(w), lo = fp16((w - hi) * 2^11), tiled [K/8][hi|lo][N][8], plus the
//     16-deep bias k-step (csrc/tc_common.cuh, csrc/net_lean.cu); s = 13 - floor(log2(max(|W|, |b|)))
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

constexpr double BN_EPS = 1e-3;

struct Mat {                       // row-major [rows][cols] f64
    int rows = 0, cols = 0;
    std::vector<double> v;
    double &at(int r, int c) { return v[(size_t)r * cols + c]; }
    double at(int r, int c) const { return v[(size_t)r * cols + c]; }
};

struct Packed {                    // one slot of ancsh_net_t
    std::string slot;
    int cin = 0, cout = 0, cin_pad = 0, cout_pad = 0, relu = 0, tc_exp = 0;
    std::vector<float> W, b;       // [cin_pad][cout_pad], [cout_pad]
    size_t w_off = 0, b_off = 0, tc_off = (size_t)-1;
};

int pad16(int c) { return (c + 15) / 16 * 16; }
int pad_out(int c) { return c <= 64 ? 64 : (c + 127) / 128 * 128; }

struct Vars {
    std::map<std::string, std::pair<const float *, size_t>> m;
    const float *get(const std::string &name, size_t count) const
    {
        auto it = m.find(name);
        if (it == m.end() || it->second.second != count) return nullptr;
        return it->second.first;
    }
    size_t count(const std::string &name) const
    {
        auto it = m.find(name);
        return it == m.end() ? 0 : it->second.second;
    }
};

// conv + bias (+ inference BN) -> W [cin][cout], b [cout] in f64 (weights.py::_fold)
bool fold(const Vars &V, const std::string &scope, int cout, bool bn, Mat &W, std::vector<double> &b)
{
    const size_t nw = V.count(scope + "/weights");
    if (nw == 0 || nw % (size_t)cout != 0) return false;
    const int cin = (int)(nw / cout);
    const float *w = V.get(scope + "/weights", nw), *bias = V.get(scope + "/biases", cout);
    if (!w || !bias) return false;
    W.rows = cin; W.cols = cout; W.v.resize(nw);
    b.resize(cout);
    std::vector<double> s(cout, 1.0);
    if (bn) {
        const float *g = V.get(scope + "/bn/gamma", cout), *beta = V.get(scope + "/bn/beta", cout);
        const float *mu = V.get(scope + "/bn/moving_mean", cout), *var = V.get(scope + "/bn/moving_variance", cout);
        if (!g || !beta || !mu || !var) return false;
        for (int c = 0; c < cout; ++c) {
            s[c] = (double)g[c] / std::sqrt((double)var[c] + BN_EPS);
            b[c] = ((double)bias[c] - (double)mu[c]) * s[c] + (double)beta[c];
        }
    } else {
        for (int c = 0; c < cout; ++c) b[c] = (double)bias[c];
    }
    for (int r = 0; r < cin; ++r)
        for (int c = 0; c < cout; ++c) W.v[(size_t)r * cout + c] = (double)w[(size_t)r * cout + c] * s[c];
    return true;
}

Packed make_packed(const std::string &slot, const Mat &W, const std::vector<double> &b, int row0, int rows, bool relu, bool zero_bias)
{
    Packed p;
    p.slot = slot; p.cin = rows; p.cout = W.cols; p.cin_pad = pad16(rows); p.cout_pad = pad_out(W.cols); p.relu = relu ? 1 : 0;
    p.W.assign((size_t)p.cin_pad * p.cout_pad, 0.f);
    p.b.assign(p.cout_pad, 0.f);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < W.cols; ++c) p.W[(size_t)r * p.cout_pad + c] = (float)W.at(row0 + r, c);
    if (!zero_bias)
        for (int c = 0; c < W.cols; ++c) p.b[c] = (float)b[c];
    return p;
}

// rows [xyz(3), features...] -> [features..., xyz(3)]
void xyz_last(Mat &W)
{
    Mat o = W;
    for (int r = 0; r < W.rows; ++r) {
        const int src = r < W.rows - 3 ? r + 3 : r - (W.rows - 3);
        for (int c = 0; c < W.cols; ++c) o.at(r, c) = W.at(src, c);
    }
    W = o;
}

int scale_exp(const Packed &p)
{
    float m = 0.f;
    for (float x : p.W) m = fmaxf(m, fabsf(x));
    for (float x : p.b) m = fmaxf(m, fabsf(x));
    if (!std::isfinite(m)) return 1000;                // rejected by the caller
    if (m == 0.f) return 0;
    int e = 13 - (int)std::floor(std::log2((double)m));
    return e < -100 ? -100 : (e > 100 ? 100 : e);
}

// [K/8][2][N][8] fp16 bit patterns of (W * 2^s)^T with the bias k-step appended (weights.py::tc_image)
bool tc_image(const Packed &p, std::vector<unsigned short> &img)
{
    const int K = p.cin_pad + 16, N = p.cout_pad;
    img.assign((size_t)(K / 8) * 2 * N * 8, 0);
    auto put = [&](int k, int n, float w) -> bool {
        if (!(fabsf(w) < 65504.f)) return false;
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn((w - __half2float(hi)) * 2048.f);
        const size_t base = ((size_t)(k / 8) * 2 * N + n) * 8 + (k % 8);
        img[base] = __half_as_ushort(hi);
        img[base + (size_t)N * 8] = __half_as_ushort(lo);
        return true;
    };
    for (int k = 0; k < p.cin_pad; ++k)
        for (int n = 0; n < N; ++n)
            if (!put(k, n, ldexpf(p.W[(size_t)k * N + n], p.tc_exp))) return false;
    for (int n = 0; n < N; ++n) {
        const float bs = ldexpf(p.b[n], p.tc_exp);
        if (!(fabsf(bs) < 65504.f)) return false;
        const float b_hi = __half2float(__float2half_rn(bs));
        const float b_lo = __half2float(__float2half_rn(bs - b_hi));
        if (!put(p.cin_pad, n, b_hi) || !put(p.cin_pad + 1, n, b_lo)) return false;
    }
    return true;
}

const char *TC_SLOTS[] = {"sa1[0]", "sa1[1]", "sa1[2]", "sa2[0]", "sa2[1]", "sa2[2]", "sa3[0]", "sa3[1]", "sa3[2]", "fp1[0]", "fp1[1]",
                          "fp2[0]", "fp2[1]", "fp3[0]", "fp3[1]", "fp3[2]", "fc1", "nocs_heads", "fc3[0]", "fc3[1]", "joint_heads"};

}  // namespace

struct ancsh_packed {
    std::vector<Packed> layers;                 // in the order of weights.py::pack_network
    std::vector<float> flat;                    // all W / b, 256-byte aligned offsets (weights.py::flatten_packed)
    std::vector<unsigned short> tc;             // all tensor-core images (weights.py::flatten_tc_images)
    float sa1_conv0[256];
    int has_sa1_conv0 = 0;
    int n_parts = 0, mixed = 0;
    const Packed *find(const char *slot) const
    {
        for (const Packed &p : layers)
            if (p.slot == slot) return &p;
        return nullptr;
    }
};

struct ancsh_net_handle {
    ancsh_net_t net;
    float *d_flat = nullptr;
    unsigned short *d_tc = nullptr;
    float sa1_conv0[256];
};

extern "C" {

int ancsh_weights_pack(int n_vars, const char *const *names, const float *const *data_host, const size_t *counts, int n_parts,
                       int mixed_pred, int early_split_nocs, const char *prefix, ancsh_packed_t **out)
{
    if (n_vars <= 0 || !names || !data_host || !counts || !out || n_parts < 1) return ANCSH_ERR_INVALID_ARG;
    Vars V;
    for (int i = 0; i < n_vars; ++i) {
        if (!names[i] || !data_host[i]) return ANCSH_ERR_INVALID_ARG;
        V.m[names[i]] = {data_host[i], counts[i]};
    }
    const std::string pre = prefix ? prefix : "SPFN";
    const std::string e = pre + "/est_net/", n = pre + "/nocs_net/", j = pre + "/joint_net/";
    const int K = n_parts;
    ancsh_packed *P = new ancsh_packed;
    P->n_parts = K; P->mixed = mixed_pred ? 1 : 0;
    Mat W;
    std::vector<double> b;
#define ANCSH_FOLD(scope, cout, bn)                                    \
    do {                                                               \
        if (!fold(V, (scope), (cout), (bn), W, b)) { delete P; return ANCSH_ERR_INVALID_ARG; } \
    } while (0)
    const int sa_dims[3][4] = {{3, 64, 64, 128}, {131, 128, 128, 256}, {259, 256, 512, 1024}};
    for (int lvl = 0; lvl < 3; ++lvl)
        for (int i = 0; i < 3; ++i) {
            char scope[256], slot[32];
            snprintf(scope, sizeof scope, "%slayer%d/conv%d", e.c_str(), lvl + 1, i);
            snprintf(slot, sizeof slot, "sa%d[%d]", lvl + 1, i);
            ANCSH_FOLD(scope, sa_dims[lvl][i + 1], true);
            if (W.rows != sa_dims[lvl][i]) { delete P; return ANCSH_ERR_INVALID_ARG; }
            if (i == 0) xyz_last(W);
            P->layers.push_back(make_packed(slot, W, b, 0, W.rows, true, false));
        }
    ANCSH_FOLD(e + "fa_layer1/conv_0", 256, true);
    if (W.rows != 1280) { delete P; return ANCSH_ERR_INVALID_ARG; }
    P->layers.push_back(make_packed("fp1_global", W, b, 0, 1024, false, false));
    P->layers.push_back(make_packed("fp1[0]", W, b, 1024, 256, true, true));       // bias comes per cloud from fp1_global
    ANCSH_FOLD(e + "fa_layer1/conv_1", 256, true);
    P->layers.push_back(make_packed("fp1[1]", W, b, 0, W.rows, true, false));
    const int fp2_out[2] = {256, 128};
    for (int i = 0; i < 2; ++i) {
        ANCSH_FOLD(e + "fa_layer2/conv_" + std::to_string(i), fp2_out[i], true);
        P->layers.push_back(make_packed("fp2[" + std::to_string(i) + "]", W, b, 0, W.rows, true, false));
    }
    for (int i = 0; i < 3; ++i) {
        ANCSH_FOLD(e + "fa_layer3/conv_" + std::to_string(i), 128, true);
        P->layers.push_back(make_packed("fp3[" + std::to_string(i) + "]", W, b, 0, W.rows, true, false));
    }
    ANCSH_FOLD(e + "fc1", 128, true);
    P->layers.push_back(make_packed("fc1", W, b, 0, W.rows, true, false));
    {
        // nocs_net heads [K, 3K] (+ [K, 3K] when mixed) + [1] concatenated column-wise; fc11_1 folded into fc2_1
        std::vector<int> dims = {K, 3 * K};
        if (mixed_pred) { dims.push_back(K); dims.push_back(3 * K); }
        dims.push_back(1);
        int total = 0;
        for (int d : dims) total += d;
        Mat H; H.rows = 128; H.cols = total; H.v.assign((size_t)128 * total, 0.0);
        std::vector<double> hb(total, 0.0);
        int c0 = 0;
        for (size_t i = 0; i < dims.size(); ++i) {
            ANCSH_FOLD(n + "fc2_" + std::to_string(i), dims[i], false);
            if (W.rows != 128) { delete P; return ANCSH_ERR_INVALID_ARG; }
            if (early_split_nocs && i == 1) {
                Mat W1; std::vector<double> b1;
                if (!fold(V, n + "fc11_1", 128, false, W1, b1) || W1.rows != 128) { delete P; return ANCSH_ERR_INVALID_ARG; }
                Mat W2 = W;                                               // W = W1 @ W2, b = b1 @ W2 + b2
                for (int c = 0; c < dims[i]; ++c) {
                    double acc = 0.0;
                    for (int k = 0; k < 128; ++k) acc += b1[k] * W2.at(k, c);
                    b[c] = acc + b[c];
                }
                for (int r = 0; r < 128; ++r)
                    for (int c = 0; c < dims[i]; ++c) {
                        double acc = 0.0;
                        for (int k = 0; k < 128; ++k) acc += W1.at(r, k) * W2.at(k, c);
                        W.at(r, c) = acc;
                    }
            }
            for (int r = 0; r < 128; ++r)
                for (int c = 0; c < dims[i]; ++c) H.at(r, c0 + c) = W.at(r, c);
            for (int c = 0; c < dims[i]; ++c) hb[c0 + c] = b[c];
            c0 += dims[i];
        }
        P->layers.push_back(make_packed("nocs_heads", H, hb, 0, 128, false, false));
    }
    for (int i = 0; i < 2; ++i) {
        ANCSH_FOLD(j + "fc3_" + std::to_string(i), 128, true);
        P->layers.push_back(make_packed("fc3[" + std::to_string(i) + "]", W, b, 0, W.rows, true, false));
    }
    {
        const int dims[4] = {3, 3, 1, 3};                                  // lib/architecture.py:195 (index head is 3 wide for every K)
        Mat H; H.rows = 128; H.cols = 10; H.v.assign((size_t)128 * 10, 0.0);
        std::vector<double> hb(10, 0.0);
        int c0 = 0;
        for (int i = 0; i < 4; ++i) {
            ANCSH_FOLD(j + "fc4_" + std::to_string(i), dims[i], false);
            if (W.rows != 128) { delete P; return ANCSH_ERR_INVALID_ARG; }
            for (int r = 0; r < 128; ++r)
                for (int c = 0; c < dims[i]; ++c) H.at(r, c0 + c) = W.at(r, c);
            for (int c = 0; c < dims[i]; ++c) hb[c0 + c] = b[c];
            c0 += dims[i];
        }
        P->layers.push_back(make_packed("joint_heads", H, hb, 0, 128, false, false));
    }
#undef ANCSH_FOLD
    // one f32 buffer, offsets aligned to 64 floats (weights.py::flatten_packed)
    size_t off = 0;
    for (Packed &p : P->layers) {
        p.w_off = off; off = (off + p.W.size() + 63) / 64 * 64;
        p.b_off = off; off = (off + p.b.size() + 63) / 64 * 64;
    }
    P->flat.assign(off, 0.f);
    for (const Packed &p : P->layers) {
        memcpy(P->flat.data() + p.w_off, p.W.data(), p.W.size() * sizeof(float));
        memcpy(P->flat.data() + p.b_off, p.b.data(), p.b.size() * sizeof(float));
    }
    // tensor-core images, offsets aligned to 128 halves (weights.py::flatten_tc_images)
    size_t toff = 0;
    std::vector<std::vector<unsigned short>> imgs;
    for (const char *slot : TC_SLOTS) {
        Packed *p = const_cast<Packed *>(P->find(slot));
        if (!p) { delete P; return ANCSH_ERR_INVALID_ARG; }
        p->tc_exp = scale_exp(*p);
        std::vector<unsigned short> img;
        if (p->tc_exp > 100 || !tc_image(*p, img)) { delete P; return ANCSH_ERR_UNSUPPORTED; }
        p->tc_off = toff;
        toff = (toff + img.size() + 127) / 128 * 128;
        imgs.push_back(std::move(img));
    }
    P->tc.assign(toff, 0);
    {
        size_t i = 0;
        for (const char *slot : TC_SLOTS) {
            const Packed *p = P->find(slot);
            memcpy(P->tc.data() + p->tc_off, imgs[i].data(), imgs[i].size() * sizeof(unsigned short));
            ++i;
        }
    }
    const Packed *l0 = P->find("sa1[0]");
    if (l0->cout_pad == 64 && l0->cin == 3) {
        for (int k = 0; k < 3; ++k)
            for (int c = 0; c < 64; ++c) P->sa1_conv0[k * 64 + c] = l0->W[(size_t)k * 64 + c];
        for (int c = 0; c < 64; ++c) P->sa1_conv0[192 + c] = l0->b[c];
        P->has_sa1_conv0 = 1;
    }
    *out = P;
    return ANCSH_OK;
}

void ancsh_packed_destroy(ancsh_packed_t *p) { delete p; }

const float *ancsh_packed_flat(const ancsh_packed_t *p, size_t *count)
{
    if (!p) return nullptr;
    if (count) *count = p->flat.size();
    return p->flat.data();
}

const unsigned short *ancsh_packed_tc(const ancsh_packed_t *p, size_t *count)
{
    if (!p) return nullptr;
    if (count) *count = p->tc.size();
    return p->tc.data();
}

int ancsh_packed_layer(const ancsh_packed_t *p, const char *slot, size_t *w_off, size_t *b_off, size_t *tc_off, int *dims5,
                       int *tc_exp)
{
    const Packed *l = p && slot ? p->find(slot) : nullptr;
    if (!l) return ANCSH_ERR_INVALID_ARG;
    if (w_off) *w_off = l->w_off;
    if (b_off) *b_off = l->b_off;
    if (tc_off) *tc_off = l->tc_off;
    if (dims5) { dims5[0] = l->cin; dims5[1] = l->cout; dims5[2] = l->cin_pad; dims5[3] = l->cout_pad; dims5[4] = l->relu; }
    if (tc_exp) *tc_exp = l->tc_exp;
    return ANCSH_OK;
}

int ancsh_net_create(const ancsh_packed_t *p, int nsample, int use_tensor_cores, ancsh_net_handle_t **out)
{
    if (!p || !out || nsample <= 0) return ANCSH_ERR_INVALID_ARG;
    ancsh_net_handle *h = new ancsh_net_handle;
    memset(&h->net, 0, sizeof h->net);
    if (cudaMalloc(&h->d_flat, p->flat.size() * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&h->d_tc, p->tc.size() * sizeof(unsigned short)) != cudaSuccess ||
        cudaMemcpy(h->d_flat, p->flat.data(), p->flat.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(h->d_tc, p->tc.data(), p->tc.size() * sizeof(unsigned short), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(h->d_flat); cudaFree(h->d_tc);
        delete h;
        return ANCSH_ERR_CUDA;
    }
    ancsh_net_t &n = h->net;
    n.use_tensor_cores = use_tensor_cores ? 1 : 0;
    n.n_parts = p->n_parts; n.mixed_pred = p->mixed;
    n.npoint1 = 512; n.nsample1 = nsample; n.radius1 = 0.2f;              // architectures.py:62-70
    n.npoint2 = 128; n.nsample2 = nsample; n.radius2 = 0.4f;
    n.tc_bias_step = 1;
    auto fill = [&](ancsh_layer_t &dst, const char *slot) {
        const Packed *l = p->find(slot);
        dst.W = h->d_flat + l->w_off; dst.b = h->d_flat + l->b_off;
        dst.W_tc = l->tc_off != (size_t)-1 ? (const void *)(h->d_tc + l->tc_off) : nullptr;
        dst.cin = l->cin; dst.cout = l->cout; dst.cin_pad = l->cin_pad; dst.cout_pad = l->cout_pad; dst.relu = l->relu;
        dst.tc_descale = ldexpf(1.f, -l->tc_exp);
    };
    char slot[32];
    for (int i = 0; i < 3; ++i) {
        snprintf(slot, sizeof slot, "sa1[%d]", i); fill(n.sa1[i], slot);
        snprintf(slot, sizeof slot, "sa2[%d]", i); fill(n.sa2[i], slot);
        snprintf(slot, sizeof slot, "sa3[%d]", i); fill(n.sa3[i], slot);
        snprintf(slot, sizeof slot, "fp3[%d]", i); fill(n.fp3[i], slot);
    }
    for (int i = 0; i < 2; ++i) {
        snprintf(slot, sizeof slot, "fp1[%d]", i); fill(n.fp1[i], slot);
        snprintf(slot, sizeof slot, "fp2[%d]", i); fill(n.fp2[i], slot);
        snprintf(slot, sizeof slot, "fc3[%d]", i); fill(n.fc3[i], slot);
    }
    fill(n.fp1_global, "fp1_global"); fill(n.fc1, "fc1"); fill(n.nocs_heads, "nocs_heads"); fill(n.joint_heads, "joint_heads");
    memcpy(h->sa1_conv0, p->sa1_conv0, sizeof h->sa1_conv0);
    n.sa1_conv0_host = p->has_sa1_conv0 ? h->sa1_conv0 : nullptr;
    *out = h;
    return ANCSH_OK;
}

const ancsh_net_t *ancsh_net_get(const ancsh_net_handle_t *h) { return h ? &h->net : nullptr; }

void ancsh_net_destroy(ancsh_net_handle_t *h)
{
    if (!h) return;
    cudaFree(h->d_flat);
    cudaFree(h->d_tc);
    delete h;
}

}  // extern "C"
