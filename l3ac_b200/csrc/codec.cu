// Step-level engine behind the C ABI: l3ac_create / l3ac_workspace_bytes / l3ac_encode / l3ac_decode / l3ac_destroy and the
// host-buffer calls l3ac_encode_host / l3ac_decode_host (include/l3ac_b200.h, "step-level interface").
//
// A handle owns what the operator-level entry points leave to the caller: the reference's checkpoint tensors folded and
// packed once (weight-norm w = g v / ||v||, l3ac/layers.py:17-18; tap-major conv weights; bf16 copies or split (hi, lo) bf16
// pairs; the GEGLU column interleave; the input-independent DynamicPositionBias table, l3ac/local_trans.py:30,43; GRN as a
// per-channel affine) and the LAUNCH SEQUENCE of the two hot-path methods,
//   L3AC.encode_audio  l3ac/__init__.py:108-114  (Codec.preprocess codec.py:79-84 -> Encoder modules.py:71-116 ->
//                      LocalEncoder / CompressedLocalEncoderWithCache local_trans.py:56-74,145-165 -> VQEmbed vq/__init__.py:25-30)
//   L3AC.decode_audio  l3ac/__init__.py:116-121  (VQEmbed.to_features vq/__init__.py:20-23 -> LocalDecoder /
//                      CompressedLocalDecoderWithCache local_trans.py:77-94,168-186 -> Decoder modules.py:135-201)
// in the product precision: encode side 3-term split-bf16, decode side bf16, fp32 accumulation and residual stream.
// Host code only -- every kernel is reached through the operator-level ABI of this same library, in the same order and with
// the same arguments as l3ac_b200/engine.py (which calls this file for the default precision).
//
// l3ac_encode / l3ac_decode allocate nothing and never synchronise: activations live in a caller-provided device workspace
// (first-fit sub-allocation, sized by a dry run of the same sequence: l3ac_workspace_bytes), all launches go to the caller's
// stream, so a step is capturable into a CUDA graph.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace {

constexpr float kCnEps = 1e-8f;      // ChannelNorm eps, l3ac/xtract/nn/utils.py:33
constexpr float kLnEps = 1e-5f;      // nn.LayerNorm default inside local_attention
constexpr int kFfPad = 352;          // FeedForward inner 341 -> multiple of 16 bf16 elements
constexpr int kHeads = 6;            // l3ac/local_trans.py:52
constexpr size_t kNone = (size_t)-1;
constexpr size_t kAlign = 512;

thread_local std::string g_last_error;

struct Err {
    int code;
    std::string msg;
};

[[noreturn]] void fail(int code, const std::string& msg) { throw Err{code, msg}; }

// ------------------------------------------------------------------------------------------------ checkpoint access
struct HostTensor {
    const float* data;
    long long numel;
};
struct Dict {
    std::unordered_map<std::string, HostTensor> map;
    std::string module;      // current "<module>." prefix
    const float* get(const std::string& key, long long numel) const {
        auto it = map.find(module + key);
        if (it == map.end()) fail(L3AC_EINVAL, "missing tensor " + module + key);
        if (it->second.numel != numel)
            fail(L3AC_EINVAL, "tensor " + module + key + " has " + std::to_string(it->second.numel) + " elements, expected " + std::to_string(numel));
        return it->second.data;
    }
    bool has(const std::string& key) const { return map.find(module + key) != map.end(); }
    std::vector<float> vec(const std::string& key, long long numel) const {
        const float* p = get(key, numel);
        return std::vector<float>(p, p + numel);
    }
    // weight-normed layer (l3ac/layers.py:17-18): g * v / ||v||, the norm over every dim but the first
    std::vector<float> folded(const std::string& prefix, int out, long long inner) const {
        if (has(prefix + ".weight")) return vec(prefix + ".weight", (long long)out * inner);
        const float* g = get(prefix + ".parametrizations.weight.original0", out);
        const float* v = get(prefix + ".parametrizations.weight.original1", (long long)out * inner);
        std::vector<float> w((size_t)out * inner);
        for (int o = 0; o < out; ++o) {
            double s = 0.0;
            for (long long i = 0; i < inner; ++i) s += (double)v[o * inner + i] * v[o * inner + i];
            const float sc = g[o] / (float)std::sqrt(s);
            for (long long i = 0; i < inner; ++i) w[o * inner + i] = v[o * inner + i] * sc;
        }
        return w;
    }
};

uint16_t bf16_rne(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
float bf16_to_float(uint16_t h) {
    const uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// (Co, Ci, k) conv weight -> (Co, k*Ci) GEMM weight with the tap index outermost
std::vector<float> taps_major(const std::vector<float>& w, int co, int ci, int k) {
    std::vector<float> o(w.size());
    for (int n = 0; n < co; ++n)
        for (int c = 0; c < ci; ++c)
            for (int s = 0; s < k; ++s) o[((size_t)n * k + s) * ci + c] = w[((size_t)n * ci + c) * k + s];
    return o;
}

// ------------------------------------------------------------------------------------------------ packed weights
struct Blob {                 // host image of the device weight blob; offsets become pointers after the upload
    std::vector<uint8_t> bytes;
    size_t put(const void* src, size_t n) {
        const size_t off = (bytes.size() + 255) & ~(size_t)255;
        bytes.resize(off + n);
        memcpy(bytes.data() + off, src, n);
        return off;
    }
    size_t f32(const std::vector<float>& v) { return put(v.data(), v.size() * 4); }
    size_t f32(const float* p, size_t n) { return put(p, n * 4); }
};

enum Kind { kF32 = L3AC_F32, kBf16 = L3AC_BF16, kSplit = L3AC_BF16X2 };

struct Lin {
    size_t w = kNone, w_lo = kNone, bias = kNone;
    int N = 0, Ktot = 0;
};

Lin pack_lin(Blob& b, const std::vector<float>& w, int N, int Ktot, const float* bias, int kind) {
    Lin l;
    l.N = N;
    l.Ktot = Ktot;
    std::vector<uint16_t> hi(w.size());
    for (size_t i = 0; i < w.size(); ++i) hi[i] = bf16_rne(w[i]);
    l.w = b.put(hi.data(), hi.size() * 2);
    if (kind == kSplit) {
        std::vector<uint16_t> lo(w.size());
        for (size_t i = 0; i < w.size(); ++i) lo[i] = bf16_rne(w[i] - bf16_to_float(hi[i]));
        l.w_lo = b.put(lo.data(), lo.size() * 2);
    }
    if (bias) l.bias = b.f32(bias, N);
    return l;
}

struct Unit {
    int C = 0;
    std::vector<float> h_dw, h_dwb, h_lnw, h_lnb;  // host copies of the prologue parameters (for plans built with a neighbour's)
    l3ac_convunit_plan* plan = nullptr;            // thin encode-side stages (C = 24 / 48): whole unit in one kernel
    l3ac_dwconv_plan* dw_plan = nullptr;           // bf16 units with C = 48 / 96: thread-per-row dwconv7 + LayerNorm
    size_t dw_w = kNone, dw_b = kNone, ln_w = kNone, ln_b = kNone, alpha = kNone, ialpha = kNone, scale = kNone, shift = kNone;
    Lin pw1, pw2;
};
struct EncStage {
    std::vector<Unit> units;
    int stride = 1, C_in = 0, C_out = 0;
    Lin down;
    size_t cn_w = kNone, cn_b = kNone;
};
struct Layer {
    size_t ln1_w, ln1_b, ln2_w, ln2_b;
    Lin qkv, out, ff1, ff2;
};
struct Trans {
    std::vector<Layer> layers;
    int window = 0;
    size_t table = kNone;
};
struct DecStage {
    std::vector<Unit> units;
    int stride = 1, C_in = 0, C_out = 0;
    size_t conv_w, conv_b, in_w, in_b, merge_w, merge_b;      // EnhanceBlock
    Lin up;
    size_t cn_w = kNone, cn_b = kNone;
    l3ac_updw_plan* updw = nullptr;      // Upsample + ChannelNorm fused with the next stage's first dwconv7 + LayerNorm
    l3ac_enhup_plan* enhup = nullptr;    // EnhanceBlock gate + the 1x1 up conv in one kernel (thin stages)
    std::vector<float> h_cnw, h_cnb;
};

}  // namespace

struct l3ac_codec {
    l3ac_codec_config cfg{};
    int dev = 0;
    int hop = 0, frame = 0;               // samples per token / per conv-encoder frame
    int dec_kind = L3AC_BF16;             // decode-side operand kind: bf16, or the split pair for precision "split"
    bool compressed = false;
    uint8_t* dweights = nullptr;
    l3ac_stem_plan* stem = nullptr;
    l3ac_tail_plan* tail = nullptr;
    std::vector<EncStage> enc_stages;
    std::vector<Unit> enc_last;
    Lin enc_out;
    Trans enc_frame, enc_token, dec_token, dec_frame;
    Lin enc_trans_down;
    size_t vq_w_in, vq_b_in, vq_w_out, vq_b_out;
    Lin dec_in;
    std::vector<DecStage> dec_stages;
    std::atomic<long long> launches{0};
    // state of the host-buffer calls (l3ac_encode_host / l3ac_decode_host): streams, staging buffers and workspaces, grown on demand
    static constexpr int kMaxStreams = 4;
    cudaStream_t streams[kMaxStreams] = {nullptr, nullptr, nullptr, nullptr};
    void* slot_ws[kMaxStreams] = {nullptr, nullptr, nullptr, nullptr};
    size_t slot_ws_bytes[kMaxStreams] = {0, 0, 0, 0};
    void* staging = nullptr;
    size_t staging_bytes = 0;

    template <typename T = float>
    const T* P(size_t off) const { return off == kNone ? nullptr : reinterpret_cast<const T*>(dweights + off); }
};

namespace {

// ------------------------------------------------------------------------------------------------ packing
Unit pack_unit(Blob& b, const Dict& d, const std::string& p, int C, int kind, bool thin_plan) {
    Unit u;
    u.C = C;
    const int H = 4 * C;
    std::vector<float> dw = d.folded(p + ".dw_conv", C, 7);                  // (C, 1, 7)
    std::vector<float> dw_t((size_t)7 * C);                                  // (7, C) tap-major
    for (int c = 0; c < C; ++c)
        for (int s = 0; s < 7; ++s) dw_t[(size_t)s * C + c] = dw[(size_t)c * 7 + s];
    const float* dw_b = d.get(p + ".dw_conv.bias", C);
    const float* ln_w = d.get(p + ".norm.weight", C);
    const float* ln_b = d.get(p + ".norm.bias", C);
    std::vector<float> w1 = d.folded(p + ".pw_conv1", H, C), w2 = d.folded(p + ".pw_conv2", C, H);
    const float* b1 = d.get(p + ".pw_conv1.bias", H);
    const float* b2 = d.get(p + ".pw_conv2.bias", C);
    const float* alpha = d.get(p + ".act.alpha", H);
    const float* gamma = d.get(p + ".grn.gamma", H);
    const float* beta = d.get(p + ".grn.beta", H);
    // GRN (l3ac/layers.py:112-115): n_x = g / (g + 1e-8) == 1 to within 1e-8 / g, so gamma * (x * n_x) + beta + x is the
    // per-channel affine (1 + gamma) x + beta
    std::vector<float> scale(H), ialpha(H);
    for (int i = 0; i < H; ++i) {
        scale[i] = 1.0f + gamma[i];
        ialpha[i] = 1.0f / (alpha[i] + 1e-8f);
    }
    if (thin_plan) {
        int rc = l3ac_convunit_plan_create(C, dw_t.data(), dw_b, ln_w, ln_b, kCnEps, w1.data(), b1, alpha, scale.data(), beta,
                                           w2.data(), b2, &u.plan);
        if (rc != 0) fail(rc, "l3ac_convunit_plan_create(" + p + ")");
        return u;
    }
    if (kind == kBf16 && (C == 48 || C == 96)) {
        int rc = l3ac_dwconv_plan_create(C, dw_t.data(), dw_b, ln_w, ln_b, kCnEps, &u.dw_plan);
        if (rc != 0) fail(rc, "l3ac_dwconv_plan_create(" + p + ")");
        u.h_dw = dw_t;
        u.h_dwb.assign(dw_b, dw_b + C);
        u.h_lnw.assign(ln_w, ln_w + C);
        u.h_lnb.assign(ln_b, ln_b + C);
    }
    u.dw_w = b.f32(dw_t);
    u.dw_b = b.f32(dw_b, C);
    u.ln_w = b.f32(ln_w, C);
    u.ln_b = b.f32(ln_b, C);
    u.alpha = b.f32(alpha, H);
    u.ialpha = b.f32(ialpha);
    u.scale = b.f32(scale);
    u.shift = b.f32(beta, H);
    u.pw1 = pack_lin(b, w1, H, C, b1, kind);
    u.pw2 = pack_lin(b, w2, C, H, b2, kind);
    return u;
}

Lin pack_conv(Blob& b, const Dict& d, const std::string& p, int co, int ci, int k, int kind) {
    std::vector<float> w = d.folded(p, co, (long long)ci * k);
    return pack_lin(b, k == 1 ? w : taps_major(w, co, ci, k), co, ci * k, d.get(p + ".bias", co), kind);
}

double silu(double x) { return x / (1.0 + std::exp(-x)); }

Trans pack_trans(Blob& b, const Dict& d, const std::string& p, int dim, int depth, int window, int kind) {
    Trans t;
    t.window = window;
    // DynamicPositionBias (l3ac/local_trans.py:30,43): the MLP input is the integer distance q_pos - k_pos in [0, 2w), so the
    // bias is a table f[h][d] independent of the input
    const int hdim = dim / 2;
    const std::string q = p + ".dynamic_pos_bias.mlp";
    const float *w0 = d.get(q + ".0.weight", hdim), *b0 = d.get(q + ".0.bias", hdim);
    const float *w2 = d.get(q + ".2.weight", (long long)hdim * hdim), *b2 = d.get(q + ".2.bias", hdim);
    const float *w4 = d.get(q + ".4.weight", (long long)kHeads * hdim), *b4 = d.get(q + ".4.bias", kHeads);
    std::vector<float> table((size_t)kHeads * 2 * window);
    std::vector<double> h0(hdim), h1(hdim);
    for (int dist = 0; dist < 2 * window; ++dist) {
        for (int i = 0; i < hdim; ++i) h0[i] = silu((double)w0[i] * dist + b0[i]);
        for (int i = 0; i < hdim; ++i) {
            double s = b2[i];
            for (int j = 0; j < hdim; ++j) s += (double)w2[(size_t)i * hdim + j] * h0[j];
            h1[i] = silu(s);
        }
        for (int hd = 0; hd < kHeads; ++hd) {
            double s = b4[hd];
            for (int j = 0; j < hdim; ++j) s += (double)w4[(size_t)hd * hdim + j] * h1[j];
            table[(size_t)hd * 2 * window + dist] = (float)s;
        }
    }
    t.table = b.f32(table);
    const int inner_att = kHeads * (dim / 4);
    const int ff_inner = (int)(dim * 4 * 2 / 3);
    if (ff_inner > kFfPad) fail(L3AC_EUNSUPPORTED, "FeedForward inner dimension > 352");
    for (int l = 0; l < depth; ++l) {
        const std::string a = p + ".layers." + std::to_string(l) + ".0", f = p + ".layers." + std::to_string(l) + ".1";
        Layer L;
        L.ln1_w = b.f32(d.get(a + ".norm.weight", dim), dim);
        L.ln1_b = b.f32(d.get(a + ".norm.bias", dim), dim);
        L.qkv = pack_lin(b, d.vec(a + ".to_qkv.weight", (long long)3 * inner_att * dim), 3 * inner_att, dim, nullptr, kind);
        L.out = pack_lin(b, d.vec(a + ".to_out.weight", (long long)dim * inner_att), dim, inner_att, nullptr, kind);
        L.ln2_w = b.f32(d.get(f + ".0.weight", dim), dim);
        L.ln2_b = b.f32(d.get(f + ".0.bias", dim), dim);
        // FeedForward: (2*inner, dim) = [value rows ; gate rows] -> interleaved (value_i, gate_i) row pairs, so that the GEGLU
        // is an epilogue of the GEMM; inner 341 zero-padded to 352
        const float* w1 = d.get(f + ".1.weight", (long long)2 * ff_inner * dim);
        std::vector<float> w1i((size_t)2 * kFfPad * dim, 0.0f);
        for (int i = 0; i < ff_inner; ++i) {
            memcpy(&w1i[(size_t)(2 * i) * dim], w1 + (size_t)i * dim, (size_t)dim * 4);
            memcpy(&w1i[(size_t)(2 * i + 1) * dim], w1 + (size_t)(ff_inner + i) * dim, (size_t)dim * 4);
        }
        const float* w2f = d.get(f + ".4.weight", (long long)dim * ff_inner);
        std::vector<float> w2p((size_t)dim * kFfPad, 0.0f);
        for (int n = 0; n < dim; ++n) memcpy(&w2p[(size_t)n * kFfPad], w2f + (size_t)n * ff_inner, (size_t)ff_inner * 4);
        L.ff1 = pack_lin(b, w1i, 2 * kFfPad, dim, nullptr, kind);
        L.ff2 = pack_lin(b, w2p, dim, kFfPad, nullptr, kind);
        t.layers.push_back(L);
    }
    return t;
}

void build(l3ac_codec* c, Dict& d) {
    const l3ac_codec_config& g = c->cfg;
    const int F = g.feature_dim;
    const int ns = g.n_encoder_stages, nd = g.n_decoder_stages;
    if (ns < 1 || ns > L3AC_MAX_STAGES || nd < 1 || nd > L3AC_MAX_STAGES || g.n_levels < 1 || g.n_levels > 8)
        fail(L3AC_EINVAL, "stage / level counts out of range");
    if (F != 128) fail(L3AC_EUNSUPPORTED, "feature_dim must be 128 (attention head dimension 32)");
    if (g.encoder_dims[0] != 24 || g.decoder_dims[nd - 1] != 24) fail(L3AC_EUNSUPPORTED, "first encoder / last decoder width must be 24");
    if (!g.en_coder_dynamic_pos) fail(L3AC_EUNSUPPORTED, "rotary position path: use the operator-level interface");
    if (g.en_coder_compress_rate < 1 || g.en_coder_window_size < 1) fail(L3AC_EINVAL, "en_coder parameters");
    if (g.precision != L3AC_PRECISION_BF16 && g.precision != L3AC_PRECISION_SPLIT) fail(L3AC_EINVAL, "precision must be L3AC_PRECISION_BF16 or L3AC_PRECISION_SPLIT");
    const int dk = g.precision == L3AC_PRECISION_SPLIT ? kSplit : kBf16;      // decode-side operand kind
    c->dec_kind = dk;
    c->compressed = g.en_coder_compress_rate != 1;
    c->frame = 1;
    for (int i = 0; i + 1 < ns; ++i) c->frame *= g.compress_rates[i];
    c->hop = c->frame * g.en_coder_compress_rate;
    Blob b;

    // ---- encoder (l3ac/modules.py:71-113; stem l3ac/tconv/__init__.py:8-27)
    d.module = "encoder.";
    {
        std::vector<float> bw(140), bb(20);
        for (int i = 0; i < 5; ++i) {
            const std::string p = "blocks.0.blocks." + std::to_string(i) + ".1";
            std::vector<float> w = d.folded(p, 4, 7);
            memcpy(&bw[(size_t)i * 28], w.data(), 28 * 4);
            memcpy(&bb[(size_t)i * 4], d.get(p + ".bias", 4), 16);
        }
        std::vector<float> w1 = d.folded("blocks.0.conv_1", 80, 20), w2 = d.folded("blocks.0.conv_2", 24, 81);
        int rc = l3ac_stem_plan_create(bw.data(), bb.data(), w1.data(), d.get("blocks.0.conv_1.bias", 80), w2.data(),
                                       d.get("blocks.0.conv_2.bias", 24), 24, &c->stem);
        if (rc != 0) fail(rc, "l3ac_stem_plan_create");
    }
    int blk = 1;
    for (int i = 0; i + 1 < ns; ++i) {
        EncStage st;
        st.stride = g.compress_rates[i];
        st.C_in = g.encoder_dims[i];
        st.C_out = g.encoder_dims[i + 1];
        if (st.stride < 1 || (st.stride * st.C_in) % 8) fail(L3AC_EUNSUPPORTED, "stride * channels must be a multiple of 8");
        for (int j = 0; j < g.encoder_depths[i]; ++j)
            st.units.push_back(pack_unit(b, d, "blocks." + std::to_string(blk) + "." + std::to_string(j) + ".module", st.C_in, kSplit,
                                         st.C_in == 24 || st.C_in == 48));
        ++blk;
        st.down = pack_conv(b, d, "blocks." + std::to_string(blk) + ".0", st.C_out, st.C_in, st.stride, kSplit);
        st.cn_w = b.f32(d.get("blocks." + std::to_string(blk) + ".1.weight", st.C_out), st.C_out);
        st.cn_b = b.f32(d.get("blocks." + std::to_string(blk) + ".1.bias", st.C_out), st.C_out);
        ++blk;
        c->enc_stages.push_back(std::move(st));
    }
    {
        const int C = g.encoder_dims[ns - 1];
        for (int j = 0; j < g.encoder_depths[ns - 1]; ++j)
            c->enc_last.push_back(pack_unit(b, d, "blocks." + std::to_string(blk) + "." + std::to_string(j) + ".module", C, kSplit,
                                            C == 24 || C == 48));
        c->enc_out = pack_conv(b, d, "blocks." + std::to_string(blk + 1), F, C, 3, kSplit);
    }
    // ---- en_encoder (l3ac/local_trans.py:56-74,145-165)
    d.module = "en_encoder.";
    const int w = g.en_coder_window_size, r = g.en_coder_compress_rate;
    if (c->compressed) {
        c->enc_frame = pack_trans(b, d, "down_trans.trans", F, 3 / 2, w * r, kSplit);
        c->enc_trans_down = pack_conv(b, d, "down_trans.down_layer", F, F, r, kSplit);
        c->enc_token = pack_trans(b, d, "local_trans", F, 3 - 3 / 2, w, kSplit);
    } else {
        c->enc_token = pack_trans(b, d, "local_trans", F, 1, w, kSplit);
    }
    // ---- quantizer (l3ac/vq/__init__.py:6-15)
    d.module = "quantizer.";
    const int D = g.n_levels;
    c->vq_w_in = b.f32(d.get("project_in.weight", (long long)D * F), (size_t)D * F);
    c->vq_b_in = b.f32(d.get("project_in.bias", D), D);
    c->vq_w_out = b.f32(d.get("project_out.weight", (long long)F * D), (size_t)F * D);
    c->vq_b_out = b.f32(d.get("project_out.bias", F), F);
    // ---- en_decoder (l3ac/local_trans.py:77-94,168-186)
    d.module = "en_decoder.";
    if (c->compressed) {
        if (g.en_coder_depth < 3) fail(L3AC_EINVAL, "en_coder_depth must be >= 3 with a compressed transformer");
        c->dec_token = pack_trans(b, d, "local_trans", F, g.en_coder_depth - 2, w, dk);
        c->dec_frame = pack_trans(b, d, "up_trans.trans", F, 2, w * r, dk);
    } else {
        c->dec_token = pack_trans(b, d, "local_trans", F, g.en_coder_depth, w, dk);
    }
    // ---- decoder (l3ac/modules.py:135-198; EnhanceBlock l3ac/tconv/__init__.py:30-38)
    d.module = "decoder.";
    c->dec_in = pack_conv(b, d, "blocks.0", g.decoder_dims[0], F, 3, dk);
    blk = 1;
    for (int i = 0; i + 1 < nd; ++i) {
        DecStage st;
        st.stride = g.decode_rates[i];
        st.C_in = g.decoder_dims[i];
        st.C_out = g.decoder_dims[i + 1];
        if (st.C_in % 8 || st.stride < 1) fail(L3AC_EUNSUPPORTED, "decoder widths must be multiples of 8");
        for (int j = 0; j < g.decoder_depths[i]; ++j)
            st.units.push_back(pack_unit(b, d, "blocks." + std::to_string(blk) + "." + std::to_string(j) + ".module", st.C_in, dk,
                                         dk == kSplit && (st.C_in == 24 || st.C_in == 48)));
        ++blk;
        const std::string e = "blocks." + std::to_string(blk);
        std::vector<float> cw(28), cb(4);
        for (int k = 0; k < 4; ++k) {
            std::vector<float> wk = d.folded(e + ".blocks." + std::to_string(k) + ".1", 1, 7);
            memcpy(&cw[(size_t)k * 7], wk.data(), 28);
            cb[k] = d.get(e + ".blocks." + std::to_string(k) + ".1.bias", 1)[0];
        }
        st.conv_w = b.f32(cw);
        st.conv_b = b.f32(cb);
        st.in_w = b.f32(d.get(e + ".merge_layer.0.weight", 4), 4);
        st.in_b = b.f32(d.get(e + ".merge_layer.0.bias", 4), 4);
        st.merge_w = b.f32(d.get(e + ".merge_layer.1.weight", (long long)st.C_in * 4), (size_t)st.C_in * 4);
        st.merge_b = b.f32(d.get(e + ".merge_layer.1.bias", st.C_in), st.C_in);
        ++blk;
        st.up = pack_conv(b, d, "blocks." + std::to_string(blk) + ".0", st.C_out, st.C_in, 1, dk);
        if (dk == kBf16 && ((st.C_in == 48 && st.C_out == 24) || (st.C_in == 96 && st.C_out == 48))) {
            const std::string up = "blocks." + std::to_string(blk) + ".0";
            std::vector<float> uw = d.folded(up, st.C_out, st.C_in);
            int rc = l3ac_enhup_plan_create(st.C_in, st.C_out, d.get(e + ".merge_layer.0.weight", 4), d.get(e + ".merge_layer.0.bias", 4),
                                            d.get(e + ".merge_layer.1.weight", (long long)st.C_in * 4), d.get(e + ".merge_layer.1.bias", st.C_in),
                                            uw.data(), d.get(up + ".bias", st.C_out), &st.enhup);
            if (rc != 0) fail(rc, "l3ac_enhup_plan_create");
        }
        st.cn_w = b.f32(d.get("blocks." + std::to_string(blk) + ".2.weight", st.C_out), st.C_out);
        st.cn_b = b.f32(d.get("blocks." + std::to_string(blk) + ".2.bias", st.C_out), st.C_out);
        st.h_cnw = d.vec("blocks." + std::to_string(blk) + ".2.weight", st.C_out);
        st.h_cnb = d.vec("blocks." + std::to_string(blk) + ".2.bias", st.C_out);
        ++blk;
        c->dec_stages.push_back(std::move(st));
    }
    for (size_t i = 0; i + 1 < c->dec_stages.size(); ++i) {      // up-layer tail + next unit's prologue in one kernel
        DecStage& st = c->dec_stages[i];
        const std::vector<Unit>& nxt = c->dec_stages[i + 1].units;
        if (dk != kBf16 || nxt.empty() || nxt[0].h_dw.empty() || (st.stride != 2 && st.stride != 3)) continue;
        const Unit& u0 = nxt[0];
        int rc = l3ac_updw_plan_create(st.C_out, st.stride, st.h_cnw.data(), st.h_cnb.data(), kCnEps, u0.h_dw.data(), u0.h_dwb.data(),
                                       u0.h_lnw.data(), u0.h_lnb.data(), kCnEps, &st.updw);
        if (rc != 0) fail(rc, "l3ac_updw_plan_create");
    }
    {   // three LegacyUnits (dilations 1, 3, 9) + Snake + Conv(24 -> 1, k7) + tanh, l3ac/modules.py:47-64,174-179,192-194
        const std::string p = "blocks." + std::to_string(blk) + ".block";
        std::vector<float> conv_w, conv_b, pw_w, pw_b, a0, a1;
        for (int j = 0; j < 3; ++j) {
            const std::string q = p + ".0." + std::to_string(j) + ".module.block";
            std::vector<float> cw = d.folded(q + ".1", 24, 24 * 7), pw = d.folded(q + ".3", 24, 24);
            conv_w.insert(conv_w.end(), cw.begin(), cw.end());
            pw_w.insert(pw_w.end(), pw.begin(), pw.end());
            const float* p0 = d.get(q + ".1.bias", 24);
            conv_b.insert(conv_b.end(), p0, p0 + 24);
            p0 = d.get(q + ".3.bias", 24);
            pw_b.insert(pw_b.end(), p0, p0 + 24);
            p0 = d.get(q + ".0.alpha", 24);
            a0.insert(a0.end(), p0, p0 + 24);
            p0 = d.get(q + ".2.alpha", 24);
            a1.insert(a1.end(), p0, p0 + 24);
        }
        std::vector<float> wf = d.folded(p + ".2", 1, 24 * 7), wf_t(7 * 24);      // (1, 24, 7) -> (7, 24)
        for (int ch = 0; ch < 24; ++ch)
            for (int s = 0; s < 7; ++s) wf_t[(size_t)s * 24 + ch] = wf[(size_t)ch * 7 + s];
        const int dil[3] = {1, 3, 9};
        int rc = l3ac_tail_plan_create(conv_w.data(), conv_b.data(), pw_w.data(), pw_b.data(), a0.data(), a1.data(), dil,
                                       d.get(p + ".1.alpha", 24), wf_t.data(), d.get(p + ".2.bias", 1)[0], 24, &c->tail);
        if (rc != 0) fail(rc, "l3ac_tail_plan_create");
    }
    cudaError_t e = cudaMalloc(&c->dweights, b.bytes.size());
    if (e != cudaSuccess) fail((int)e, "cudaMalloc(weights)");
    e = cudaMemcpy(c->dweights, b.bytes.data(), b.bytes.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) fail((int)e, "cudaMemcpy(weights)");
}

// ------------------------------------------------------------------------------------------------ workspace
// First-fit sub-allocation of the caller's workspace.  The launch sequence is a pure function of (B, T), so a dry run of the
// same sequence (no launches, a virtual arena) gives the exact high-water mark: that is l3ac_workspace_bytes.
struct Arena {
    uint8_t* base = nullptr;
    size_t cap = 0, peak = 0;
    struct Blk {
        size_t off, size;
        bool used;
    };
    std::vector<Blk> blks;
    Arena(void* p, size_t n) : base(static_cast<uint8_t*>(p)), cap(n) { blks.push_back({0, n, false}); }
    void* alloc(size_t n) {
        n = (n + kAlign - 1) & ~(kAlign - 1);
        if (n == 0) n = kAlign;
        for (size_t i = 0; i < blks.size(); ++i) {
            if (blks[i].used || blks[i].size < n) continue;
            if (blks[i].size > n) blks.insert(blks.begin() + i + 1, {blks[i].off + n, blks[i].size - n, false});
            blks[i].size = n;
            blks[i].used = true;
            if (blks[i].off + n > peak) peak = blks[i].off + n;
            return base + blks[i].off;
        }
        fail(L3AC_EINVAL, "workspace too small (see l3ac_workspace_bytes)");
    }
    void free(void* p) {
        if (!p) return;
        const size_t off = (size_t)(static_cast<uint8_t*>(p) - base);
        for (size_t i = 0; i < blks.size(); ++i) {
            if (blks[i].off != off || !blks[i].used) continue;
            blks[i].used = false;
            if (i + 1 < blks.size() && !blks[i + 1].used) {
                blks[i].size += blks[i + 1].size;
                blks.erase(blks.begin() + i + 1);
            }
            if (i > 0 && !blks[i - 1].used) {
                blks[i - 1].size += blks[i].size;
                blks.erase(blks.begin() + i);
            }
            return;
        }
    }
};

struct Act {                 // an activation: (B, T, C) rows of fp32, bf16 or a split (hi, lo) bf16 pair
    void* hi = nullptr;
    void* lo = nullptr;
    int kind = kF32;
    int B = 0, T = 0, C = 0;
    bool owned = true;       // lives in the arena (false: a caller's buffer)
    long long rows() const { return (long long)B * T; }
};

struct Run {
    l3ac_codec* c;
    Arena ar;
    cudaStream_t st;
    bool dry;
    long long launches = 0;
    Run(l3ac_codec* c_, void* ws, size_t n, cudaStream_t s, bool dry_) : c(c_), ar(ws, n), st(s), dry(dry_) {}

    Act make(int kind, int B, int T, int C) {
        Act a;
        a.kind = kind;
        a.B = B;
        a.T = T;
        a.C = C;
        const size_t n = (size_t)B * T * C;
        a.hi = ar.alloc(n * (kind == kF32 ? 4 : 2));
        if (kind == kSplit) a.lo = ar.alloc(n * 2);
        return a;
    }
    Act wrap(const void* p, int kind, int B, int T, int C) {
        Act a;
        a.hi = const_cast<void*>(p);
        a.kind = kind;
        a.B = B;
        a.T = T;
        a.C = C;
        a.owned = false;
        return a;
    }
    void drop(Act& a) {
        if (a.owned) {
            ar.free(a.hi);
            ar.free(a.lo);
        }
        a.hi = a.lo = nullptr;
    }
    void ok(int rc, const char* what) {
        if (rc != 0) fail(rc, what);
        ++launches;
    }

    // out[(b,t), n] = epi(bias[n] + sum_s sum_k a[b, t + shift_s, k] w[n, s*K + k]); `a` is viewed as B*T rows of K values
    Act gemm(const Act& a, const Lin& l, int B, int T, int K, int out_kind, int act = L3AC_ACT_NONE, int taps = 1, int shift0 = 0,
             const Unit* snake = nullptr, const Act* residual = nullptr) {
        if ((long long)B * T * K != a.rows() * a.C || l.Ktot != taps * K) fail(L3AC_EINVAL, "gemm shape mismatch");
        const int n_out = act == L3AC_ACT_GEGLU ? l.N / 2 : l.N;
        Act o = make(out_kind, B, T, n_out);
        if (dry) return o;
        l3ac_gemm_desc d{};
        d.A = a.hi;
        d.A_lo = a.kind == kSplit ? a.lo : nullptr;
        d.W = c->P<void>(l.w);
        d.W_lo = a.kind == kSplit ? c->P<void>(l.w_lo) : nullptr;
        d.bias = c->P(l.bias);
        if (snake) {
            d.alpha = c->P(snake->alpha);
            d.scale = c->P(snake->scale);
            d.shift = c->P(snake->shift);
        }
        d.residual = residual ? static_cast<const float*>(residual->hi) : nullptr;
        d.out = o.hi;
        d.out_lo = o.lo;
        d.lda = K;
        d.ldr = d.ldo = n_out;
        d.B = B;
        d.T = T;
        d.K = K;
        d.N = l.N;
        d.taps = taps;
        d.tap_shift0 = shift0;
        d.tap_step = 1;
        d.act = act;
        d.out_dtype = out_kind;
        ok(l3ac_gemm_bf16_tc(&d, st), "l3ac_gemm_bf16_tc");
        return o;
    }
    Act layernorm(const Act& x, size_t w, size_t b, float eps, int out_kind) {
        Act o = make(out_kind, x.B, x.T, x.C);
        if (!dry) ok(l3ac_layernorm(static_cast<const float*>(x.hi), x.rows(), x.C, c->P(w), c->P(b), eps, o.hi, o.lo, out_kind, st), "l3ac_layernorm");
        return o;
    }
    // fp32 -> the GEMM operand kind (bf16: the hi plane of the split pair IS bf16(x))
    Act as_operand(Act& x, int kind) {
        if (x.kind == kind) return x;
        Act o = make(kSplit, x.B, x.T, x.C);
        if (!dry) ok(l3ac_split_bf16(static_cast<const float*>(x.hi), x.rows() * x.C, o.hi, o.lo, st), "l3ac_split_bf16");
        drop(x);
        if (kind == kBf16) {
            ar.free(o.lo);
            o.lo = nullptr;
            o.kind = kBf16;
        }
        return o;
    }

    // Residual(ConvUnit), l3ac/modules.py:32-44.  Consumes x (fp32); the result is fp32 or, for the last unit before a GEMM
    // consumer on the encode side, the split pair written directly by the producing kernel.
    Act conv_unit(Act& x, const Unit& u, int act_kind, int out_kind, Act* ch0 = nullptr, Act* a_pre = nullptr) {
        const int B = x.B, T = x.T, C = x.C;
        if (u.plan) {
            Act o = make(out_kind, B, T, C);
            if (!dry) ok(l3ac_convunit_umma(u.plan, static_cast<const float*>(x.hi), B, T, o.hi, o.lo, out_kind, st), "l3ac_convunit_umma");
            drop(x);
            return o;
        }
        Act a = a_pre ? *a_pre : make(act_kind, B, T, C);
        if (a_pre) {
            a_pre->hi = nullptr;         // (dwconv7 + LayerNorm already done by the fused up-layer kernel; ownership moves here)
        } else if (!dry) {
            if (u.dw_plan && act_kind == kBf16 && B <= 65535)
                ok(l3ac_dwconv7_ln_plan(u.dw_plan, static_cast<const float*>(x.hi), B, T, a.hi, st), "l3ac_dwconv7_ln_plan");
            else
                ok(l3ac_dwconv7_ln(static_cast<const float*>(x.hi), B, T, C, c->P(u.dw_w), c->P(u.dw_b), c->P(u.ln_w), c->P(u.ln_b), kCnEps,
                                   a.hi, a.lo, act_kind, st), "l3ac_dwconv7_ln");
        }
        Act o;
        if (act_kind == kBf16 && out_kind == kF32 && C >= 16 && C <= 256 && C % 16 == 0) {
            o = make(kF32, B, T, C);       // fused MLP: the 4C hidden activation stays in TMEM / shared memory
            if (ch0) *ch0 = make(kF32, B, T, 1);      // + channel 0 as a compact plane for the EnhanceBlock statistics
            if (!dry)
                ok(l3ac_convunit_mlp_tc_ch0(a.hi, c->P<void>(u.pw1.w), c->P(u.pw1.bias), c->P(u.alpha), c->P(u.ialpha), c->P(u.scale),
                                            c->P(u.shift), c->P<void>(u.pw2.w), c->P(u.pw2.bias), static_cast<const float*>(x.hi),
                                            static_cast<float*>(o.hi), ch0 ? static_cast<float*>(ch0->hi) : nullptr, x.rows(), C, st),
                   "l3ac_convunit_mlp_tc");
            drop(a);
        } else {
            Act h = gemm(a, u.pw1, B, T, C, act_kind, L3AC_ACT_SNAKE, 1, 0, &u);
            drop(a);
            o = gemm(h, u.pw2, B, T, 4 * C, out_kind, L3AC_ACT_NONE, 1, 0, nullptr, &x);
            drop(h);
        }
        drop(x);
        return o;
    }

    // LocalTrans.forward, l3ac/local_trans.py:42-48 (LocalMHA prenorm + GEGLU FeedForward).  Consumes x.
    Act local_trans(Act x, const Trans& t, int kind) {
        const int B = x.B, T = x.T, D = x.C;
        const int inner = kHeads * 32;
        for (const Layer& L : t.layers) {
            Act a = layernorm(x, L.ln1_w, L.ln1_b, kLnEps, kind);
            Act qkv = gemm(a, L.qkv, B, T, D, kind);
            drop(a);
            Act o = make(kind, B, T, inner);
            if (!dry)
                ok(l3ac_local_attention_umma(qkv.hi, qkv.lo, c->P(t.table), B, T, kHeads, 32, t.window, o.hi, o.lo, kind, st),
                   "l3ac_local_attention_umma");
            drop(qkv);
            Act x1 = gemm(o, L.out, B, T, inner, kF32, L3AC_ACT_NONE, 1, 0, nullptr, &x);
            drop(o);
            drop(x);
            a = layernorm(x1, L.ln2_w, L.ln2_b, kLnEps, kind);
            Act gg = gemm(a, L.ff1, B, T, D, kind, L3AC_ACT_GEGLU);
            drop(a);
            x = gemm(gg, L.ff2, B, T, kFfPad, kF32, L3AC_ACT_NONE, 1, 0, nullptr, &x1);
            drop(gg);
            drop(x1);
        }
        return x;
    }

    void encode(const float* audio, int B, int T0, float* q_feature, int32_t* indices, float* level_indices) {
        const l3ac_codec_config& g = c->cfg;
        const int T = (T0 + c->hop - 1) / c->hop * c->hop;              // Codec.preprocess, l3ac/codec.py:79-84
        Act x;
        {
            Act in = wrap(audio, kF32, B, T, 1);
            if (T != T0) {
                in = make(kF32, B, T, 1);
                if (!dry) {
                    cudaError_t e = cudaMemsetAsync(in.hi, 0, (size_t)B * T * 4, st);
                    if (e == cudaSuccess)
                        e = cudaMemcpy2DAsync(in.hi, (size_t)T * 4, audio, (size_t)T0 * 4, (size_t)T0 * 4, B, cudaMemcpyDeviceToDevice, st);
                    if (e != cudaSuccess) fail((int)e, "pad");
                }
            }
            x = make(kF32, B, T, 24);
            if (!dry) ok(l3ac_stem_umma(c->stem, static_cast<const float*>(in.hi), B, T, static_cast<float*>(x.hi), st), "l3ac_stem_umma");
            drop(in);
        }
        for (const EncStage& s : c->enc_stages) {
            const int Tn = x.T / s.stride;
            for (size_t j = 0; j < s.units.size(); ++j) x = conv_unit(x, s.units[j], kSplit, j + 1 == s.units.size() ? kSplit : kF32);
            Act a = as_operand(x, kSplit);
            Act y = gemm(a, s.down, B, Tn, s.stride * s.C_in, kF32);      // Conv1d(k = s, stride = s) as a GEMM over (B, T/s, s*C)
            drop(a);
            x = layernorm(y, s.cn_w, s.cn_b, kCnEps, kF32);               // channels-first ChannelNorm
            drop(y);
        }
        for (size_t j = 0; j < c->enc_last.size(); ++j) x = conv_unit(x, c->enc_last[j], kSplit, j + 1 == c->enc_last.size() ? kSplit : kF32);
        {
            Act a = as_operand(x, kSplit);
            x = gemm(a, c->enc_out, B, a.T, a.C, kF32, L3AC_ACT_NONE, 3, -1);      // Conv1d(k3, pad 1)
            drop(a);
        }
        if (c->compressed) {
            x = local_trans(x, c->enc_frame, kSplit);
            const int r = g.en_coder_compress_rate;
            Act a = as_operand(x, kSplit);
            x = gemm(a, c->enc_trans_down, B, a.T / r, r * a.C, kF32);
            drop(a);
        }
        x = local_trans(x, c->enc_token, kSplit);
        Act q = q_feature ? wrap(q_feature, kF32, B, x.T, x.C) : make(kF32, B, x.T, x.C);
        if (!dry)
            ok(l3ac_fsq_quantize(static_cast<const float*>(x.hi), x.rows(), x.C, c->P(c->vq_w_in), c->P(c->vq_b_in), c->P(c->vq_w_out),
                                 c->P(c->vq_b_out), g.levels, g.n_levels, static_cast<float*>(q.hi), indices, level_indices, nullptr, st),
               "l3ac_fsq_quantize");
        drop(q);
        drop(x);
    }

    void decode(const void* indices, int indices_are_i64, const float* q_feature, int B, int T_tok, float* audio_out) {
        const l3ac_codec_config& g = c->cfg;
        const int F = g.feature_dim;
        Act x;
        if (q_feature) {
            x = wrap(q_feature, kF32, B, T_tok, F);
        } else {
            x = make(kF32, B, T_tok, F);
            if (!dry)
                ok(l3ac_fsq_dequantize(indices, indices_are_i64, (long long)B * T_tok, F, c->P(c->vq_w_out), c->P(c->vq_b_out), g.levels,
                                       g.n_levels, static_cast<float*>(x.hi), st), "l3ac_fsq_dequantize");
        }
        const int dk = c->dec_kind;
        x = local_trans(x, c->dec_token, dk);
        if (c->compressed) {                                                  // UpTransV2, l3ac/local_trans.py:123-126
            const int r = g.en_coder_compress_rate;
            Act y = make(kF32, B, x.T * r, F);
            if (!dry)
                ok(l3ac_upsample_linear_cn(static_cast<const float*>(x.hi), B, x.T, F, r, nullptr, nullptr, kCnEps, static_cast<float*>(y.hi), st),
                   "l3ac_upsample_linear_cn");
            drop(x);
            x = local_trans(y, c->dec_frame, dk);
        }
        {
            Act a = as_operand(x, dk);
            x = gemm(a, c->dec_in, B, a.T, F, kF32, L3AC_ACT_NONE, 3, -1);         // Conv1d(k3, pad 1)
            drop(a);
        }
        Act a_pre;
        for (const DecStage& s : c->dec_stages) {
            Act ch0;
            for (size_t j = 0; j < s.units.size(); ++j)
                x = conv_unit(x, s.units[j], dk, kF32, j + 1 == s.units.size() ? &ch0 : nullptr, (j == 0 && a_pre.hi) ? &a_pre : nullptr);
            const int T = x.T, C = x.C;
            // EnhanceBlock (l3ac/tconv/__init__.py:30-44): stats pass (partial sums + branch signals), then the streaming gate
            const long long np = l3ac_enhance_partials_floats(B, T);
            float* partials = static_cast<float*>(ar.alloc((size_t)np * 4));
            float* branches = static_cast<float*>(ar.alloc((size_t)B * T * 4 * 4));
            const bool fused_up = s.enhup && B <= 65535;               // thin stages: gate + 1x1 up conv in one kernel
            Act a, y;
            if (fused_up) y = make(kF32, B, T, s.C_out);
            else a = make(dk == kBf16 ? kBf16 : kF32, B, T, C);        // (split mode: fp32 out, split below)
            if (!dry) {
                const float* xp = static_cast<const float*>(x.hi);
                if (ch0.hi) ok(l3ac_enhance_stats(static_cast<const float*>(ch0.hi), B, T, 1, c->P(s.conv_w), c->P(s.conv_b), partials, branches, st), "l3ac_enhance_stats");
                else ok(l3ac_enhance_stats(xp, B, T, C, c->P(s.conv_w), c->P(s.conv_b), partials, branches, st), "l3ac_enhance_stats");
                if (fused_up)
                    ok(l3ac_enhance_up(s.enhup, xp, B, T, partials, branches, static_cast<float*>(y.hi), st), "l3ac_enhance_up");
                else
                    ok(l3ac_enhance_apply(xp, B, T, C, c->P(s.conv_w), c->P(s.conv_b), c->P(s.in_w), c->P(s.in_b), c->P(s.merge_w),
                                          c->P(s.merge_b), partials, branches, a.hi, a.kind, st), "l3ac_enhance_apply");
            }
            ar.free(partials);
            ar.free(branches);
            if (ch0.hi) drop(ch0);
            drop(x);
            if (!fused_up) {
                a = as_operand(a, dk);
                y = gemm(a, s.up, B, T, C, kF32);                                  // Conv1d 1x1
                drop(a);
            }
            x = make(kF32, B, T * s.stride, s.C_out);                              // Upsample(linear) + ChannelNorm
            if (s.updw && B <= 65535) {                                            // ... + the next unit's dwconv7 + LayerNorm
                a_pre = make(kBf16, B, T * s.stride, s.C_out);
                if (!dry)
                    ok(l3ac_upsample_cn_dwconv7_ln(s.updw, static_cast<const float*>(y.hi), B, T, static_cast<float*>(x.hi), a_pre.hi, st),
                       "l3ac_upsample_cn_dwconv7_ln");
            } else if (!dry) {
                ok(l3ac_upsample_linear_cn(static_cast<const float*>(y.hi), B, T, s.C_out, s.stride, c->P(s.cn_w), c->P(s.cn_b), kCnEps,
                                           static_cast<float*>(x.hi), st), "l3ac_upsample_linear_cn");
            }
            drop(y);
        }
        if (!dry)
            ok((dk == kSplit ? l3ac_decoder_tail_tc_split : l3ac_decoder_tail_tc)(c->tail, static_cast<const float*>(x.hi), B, x.T, audio_out, st),
               "l3ac_decoder_tail_tc");
        drop(x);
    }
};

template <typename Fn>
int guarded(Fn&& fn) {
    try {
        fn();
        return L3AC_OK;
    } catch (const Err& e) {
        g_last_error = e.msg;
        return e.code != 0 ? e.code : L3AC_EINVAL;
    } catch (const std::bad_alloc&) {
        g_last_error = "out of host memory";
        return L3AC_EINVAL;
    }
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

void cuda_ok(cudaError_t e, const char* what) {
    if (e != cudaSuccess) fail((int)e, what);
}

// balanced micro-batches of <= ~330 s of audio (the size the kernels were tuned at: DESIGN.md section 2)
std::vector<std::pair<int, int>> micro_batches(int B, long long samples_per_clip) {
    const long long per = std::max<long long>(1, (330LL * 16000) / std::max<long long>(samples_per_clip, 1));
    const int n = (int)((B + per - 1) / per);
    std::vector<std::pair<int, int>> out;
    int lo = 0;
    for (int i = 0; i < n; ++i) {
        const int hi = lo + B / n + (i < B % n ? 1 : 0);
        out.push_back({lo, hi});
        lo = hi;
    }
    return out;
}

void ensure_slot(l3ac_codec* c, int slot, size_t ws_bytes) {
    if (!c->streams[slot]) cuda_ok(cudaStreamCreateWithFlags(&c->streams[slot], cudaStreamNonBlocking), "cudaStreamCreate");
    if (c->slot_ws_bytes[slot] < ws_bytes) {
        if (c->slot_ws[slot]) cuda_ok(cudaFree(c->slot_ws[slot]), "cudaFree");
        c->slot_ws[slot] = nullptr;
        c->slot_ws_bytes[slot] = 0;
        cuda_ok(cudaMalloc(&c->slot_ws[slot], ws_bytes), "cudaMalloc(workspace)");
        c->slot_ws_bytes[slot] = ws_bytes;
    }
}

void ensure_staging(l3ac_codec* c, size_t bytes) {
    if (c->staging_bytes >= bytes) return;
    if (c->staging) cuda_ok(cudaFree(c->staging), "cudaFree");
    c->staging = nullptr;
    c->staging_bytes = 0;
    cuda_ok(cudaMalloc(&c->staging, bytes), "cudaMalloc(staging)");
    c->staging_bytes = bytes;
}

size_t up(size_t n) { return (n + kAlign - 1) & ~(kAlign - 1); }

}  // namespace

extern "C" const char* l3ac_last_error(void) { return g_last_error.c_str(); }

extern "C" int l3ac_create(const l3ac_codec_config* cfg, const l3ac_tensor* tensors, int n_tensors, l3ac_codec** out) {
    if (!cfg || !tensors || n_tensors <= 0 || !out) return L3AC_EINVAL;
    *out = nullptr;
    l3ac_codec* c = new l3ac_codec();
    int rc = guarded([&] {
        c->cfg = *cfg;
        cuda_ok(cudaGetDevice(&c->dev), "cudaGetDevice");
        Dict d;
        for (int i = 0; i < n_tensors; ++i) {
            if (!tensors[i].name || !tensors[i].data || tensors[i].numel <= 0) fail(L3AC_EINVAL, "bad tensor entry " + std::to_string(i));
            d.map[tensors[i].name] = HostTensor{tensors[i].data, tensors[i].numel};
        }
        build(c, d);
    });
    if (rc != L3AC_OK) {
        l3ac_destroy(c);
        return rc;
    }
    *out = c;
    return L3AC_OK;
}

extern "C" int l3ac_destroy(l3ac_codec* c) {
    if (!c) return L3AC_OK;
    DeviceGuard dg(c->dev);
    auto free_units = [](std::vector<Unit>& us) {
        for (Unit& u : us) {
            if (u.plan) l3ac_convunit_plan_destroy(u.plan);
            if (u.dw_plan) l3ac_dwconv_plan_destroy(u.dw_plan);
        }
    };
    for (EncStage& s : c->enc_stages) free_units(s.units);
    for (DecStage& s : c->dec_stages) {
        free_units(s.units);
        if (s.updw) l3ac_updw_plan_destroy(s.updw);
        if (s.enhup) l3ac_enhup_plan_destroy(s.enhup);
    }
    free_units(c->enc_last);
    if (c->stem) l3ac_stem_plan_destroy(c->stem);
    if (c->tail) l3ac_tail_plan_destroy(c->tail);
    if (c->dweights) cudaFree(c->dweights);
    for (int i = 0; i < l3ac_codec::kMaxStreams; ++i) {
        if (c->streams[i]) {
            cudaStreamSynchronize(c->streams[i]);
            cudaStreamDestroy(c->streams[i]);
        }
        if (c->slot_ws[i]) cudaFree(c->slot_ws[i]);
    }
    if (c->staging) cudaFree(c->staging);
    delete c;
    return L3AC_OK;
}

extern "C" int l3ac_hop_length(const l3ac_codec* c) { return c ? c->hop : L3AC_EINVAL; }

extern "C" long long l3ac_launch_count(const l3ac_codec* c) { return c ? c->launches.load() : 0; }

extern "C" long long l3ac_workspace_bytes(const l3ac_codec* c, int B, int T) {
    if (!c || B <= 0 || T <= 0) return L3AC_EINVAL;
    long long need = 0;
    int rc = guarded([&] {
        l3ac_codec* cc = const_cast<l3ac_codec*>(c);
        const int T_tok = (T + c->hop - 1) / c->hop;
        Run enc(cc, reinterpret_cast<void*>(kAlign), (size_t)1 << 50, nullptr, true);
        enc.encode(reinterpret_cast<const float*>(kAlign), B, T, nullptr, nullptr, nullptr);
        Run dec(cc, reinterpret_cast<void*>(kAlign), (size_t)1 << 50, nullptr, true);
        dec.decode(reinterpret_cast<const void*>(kAlign), 0, nullptr, B, T_tok, nullptr);
        need = (long long)std::max(enc.ar.peak, dec.ar.peak);
    });
    return rc == L3AC_OK ? need : rc;
}

extern "C" int l3ac_encode(l3ac_codec* c, const float* audio, int B, int T, void* workspace, long long workspace_bytes, float* q_feature,
                           int32_t* indices, float* level_indices, l3ac_stream_t stream) {
    if (!c || !audio || B <= 0 || T <= 0 || !workspace || workspace_bytes <= 0 || !indices) return L3AC_EINVAL;
    if ((reinterpret_cast<uintptr_t>(workspace) & 15) || (reinterpret_cast<uintptr_t>(audio) & 3)) return L3AC_EINVAL;
    return guarded([&] {
        Run r(c, workspace, (size_t)workspace_bytes, (cudaStream_t)stream, false);
        r.encode(audio, B, T, q_feature, indices, level_indices);
        c->launches += r.launches;
    });
}

extern "C" int l3ac_decode(l3ac_codec* c, const void* indices, int indices_are_i64, const float* q_feature, int B, int T_tok, void* workspace,
                           long long workspace_bytes, float* audio, l3ac_stream_t stream) {
    if (!c || (!indices && !q_feature) || B <= 0 || T_tok <= 0 || !workspace || workspace_bytes <= 0 || !audio) return L3AC_EINVAL;
    if ((reinterpret_cast<uintptr_t>(workspace) & 15) || (reinterpret_cast<uintptr_t>(audio) & 3)) return L3AC_EINVAL;
    return guarded([&] {
        Run r(c, workspace, (size_t)workspace_bytes, (cudaStream_t)stream, false);
        r.decode(indices, indices_are_i64, q_feature, B, T_tok, audio);
        c->launches += r.launches;
    });
}

// VQEmbed.forward / VQEmbed.to_features on their own (l3ac/vq/__init__.py:20-30), with the handle's quantiser weights.
extern "C" int l3ac_quantize(l3ac_codec* c, const float* trans_feature, int B, int T_tok, float* q_feature, int32_t* indices,
                             float* level_indices, l3ac_stream_t stream) {
    if (!c || !trans_feature || B <= 0 || T_tok <= 0 || !q_feature || !indices) return L3AC_EINVAL;
    const l3ac_codec_config& g = c->cfg;
    int rc = l3ac_fsq_quantize(trans_feature, (long long)B * T_tok, g.feature_dim, c->P(c->vq_w_in), c->P(c->vq_b_in), c->P(c->vq_w_out),
                               c->P(c->vq_b_out), g.levels, g.n_levels, q_feature, indices, level_indices, nullptr, stream);
    if (rc == L3AC_OK) c->launches += 1;
    return rc;
}

extern "C" int l3ac_dequantize(l3ac_codec* c, const void* indices, int indices_are_i64, int B, int T_tok, float* q_feature,
                               l3ac_stream_t stream) {
    if (!c || !indices || B <= 0 || T_tok <= 0 || !q_feature) return L3AC_EINVAL;
    const l3ac_codec_config& g = c->cfg;
    int rc = l3ac_fsq_dequantize(indices, indices_are_i64, (long long)B * T_tok, g.feature_dim, c->P(c->vq_w_out), c->P(c->vq_b_out), g.levels,
                                 g.n_levels, q_feature, stream);
    if (rc == L3AC_OK) c->launches += 1;
    return rc;
}

// Host-buffer calls: upload, run and download micro-batch by micro-batch on up to four internal streams, so that the copies of one
// micro-batch overlap the kernels of the others (pinned host memory makes the copies asynchronous; pageable memory works, slower).
extern "C" int l3ac_encode_host(l3ac_codec* c, const float* audio, int B, int T, int32_t* indices, float* q_feature) {
    if (!c || !audio || B <= 0 || T <= 0 || !indices) return L3AC_EINVAL;
    return guarded([&] {
        DeviceGuard dg(c->dev);
        const int F = c->cfg.feature_dim;
        const int T_tok = (T + c->hop - 1) / c->hop;
        const auto mbs = micro_batches(B, T);
        const int n_slots = (int)std::min<size_t>(mbs.size(), l3ac_codec::kMaxStreams);
        const int per = mbs[0].second - mbs[0].first;
        const long long ws = l3ac_workspace_bytes(c, per, T);
        if (ws < 0) fail((int)ws, g_last_error);
        for (int s = 0; s < n_slots; ++s) ensure_slot(c, s, (size_t)ws);
        const size_t a_bytes = up((size_t)B * T * 4), i_bytes = up((size_t)B * T_tok * 4), q_bytes = up((size_t)B * T_tok * F * 4);
        ensure_staging(c, a_bytes + i_bytes + q_bytes);
        float* d_audio = static_cast<float*>(c->staging);
        int32_t* d_idx = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(c->staging) + a_bytes);
        float* d_q = reinterpret_cast<float*>(static_cast<uint8_t*>(c->staging) + a_bytes + i_bytes);
        for (size_t i = 0; i < mbs.size(); ++i) {
            const int lo = mbs[i].first, n = mbs[i].second - mbs[i].first, s = (int)(i % n_slots);
            cudaStream_t st = c->streams[s];
            cuda_ok(cudaMemcpyAsync(d_audio + (size_t)lo * T, audio + (size_t)lo * T, (size_t)n * T * 4, cudaMemcpyHostToDevice, st), "H2D audio");
            Run r(c, c->slot_ws[s], c->slot_ws_bytes[s], st, false);
            r.encode(d_audio + (size_t)lo * T, n, T, d_q + (size_t)lo * T_tok * F, d_idx + (size_t)lo * T_tok, nullptr);
            c->launches += r.launches;
            cuda_ok(cudaMemcpyAsync(indices + (size_t)lo * T_tok, d_idx + (size_t)lo * T_tok, (size_t)n * T_tok * 4, cudaMemcpyDeviceToHost, st), "D2H indices");
            if (q_feature)
                cuda_ok(cudaMemcpyAsync(q_feature + (size_t)lo * T_tok * F, d_q + (size_t)lo * T_tok * F, (size_t)n * T_tok * F * 4,
                                        cudaMemcpyDeviceToHost, st), "D2H q_feature");
        }
        for (int s = 0; s < n_slots; ++s) cuda_ok(cudaStreamSynchronize(c->streams[s]), "cudaStreamSynchronize");
    });
}

extern "C" int l3ac_decode_host(l3ac_codec* c, const int32_t* indices, int B, int T_tok, float* audio) {
    if (!c || !indices || B <= 0 || T_tok <= 0 || !audio) return L3AC_EINVAL;
    return guarded([&] {
        DeviceGuard dg(c->dev);
        const long long T = (long long)T_tok * c->hop;
        if (T > 0x7fffffffLL) fail(L3AC_EINVAL, "clip too long");
        const auto mbs = micro_batches(B, T);
        const int n_slots = (int)std::min<size_t>(mbs.size(), l3ac_codec::kMaxStreams);
        const int per = mbs[0].second - mbs[0].first;
        const long long ws = l3ac_workspace_bytes(c, per, (int)T);
        if (ws < 0) fail((int)ws, g_last_error);
        for (int s = 0; s < n_slots; ++s) ensure_slot(c, s, (size_t)ws);
        const size_t a_bytes = up((size_t)B * T * 4), i_bytes = up((size_t)B * T_tok * 4);
        ensure_staging(c, a_bytes + i_bytes);
        float* d_audio = static_cast<float*>(c->staging);
        int32_t* d_idx = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(c->staging) + a_bytes);
        for (size_t i = 0; i < mbs.size(); ++i) {
            const int lo = mbs[i].first, n = mbs[i].second - mbs[i].first, s = (int)(i % n_slots);
            cudaStream_t st = c->streams[s];
            cuda_ok(cudaMemcpyAsync(d_idx + (size_t)lo * T_tok, indices + (size_t)lo * T_tok, (size_t)n * T_tok * 4, cudaMemcpyHostToDevice, st), "H2D indices");
            Run r(c, c->slot_ws[s], c->slot_ws_bytes[s], st, false);
            r.decode(d_idx + (size_t)lo * T_tok, 0, nullptr, n, T_tok, d_audio + (size_t)lo * T);
            c->launches += r.launches;
            cuda_ok(cudaMemcpyAsync(audio + (size_t)lo * T, d_audio + (size_t)lo * T, (size_t)n * T * 4, cudaMemcpyDeviceToHost, st), "D2H audio");
        }
        for (int s = 0; s < n_slots; ++s) cuda_ok(cudaStreamSynchronize(c->streams[s]), "cudaStreamSynchronize");
    });
}
