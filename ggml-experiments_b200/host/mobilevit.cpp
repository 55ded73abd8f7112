// mobilevit.cpp -- loader + batched graph builder + C ABI (include/mobilevit_b200.h).
// See mobilevit.h for how this relates to /root/reference/mobilevit/main.cpp.
#include "mobilevit.h"

#include <cmath>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "mobilevit_b200.h"

namespace mvit {

static const char * kRoot = "tf_mobile_vi_t_model/mobilevit";

// ------------------------------------------------------------------------------------------------------
// loader: read_all_weights (main.cpp:872-942) + assign_weights (main.cpp:218-312)
// ------------------------------------------------------------------------------------------------------
static bool read_i32(std::ifstream & f, int32_t & v) { return (bool)f.read(reinterpret_cast<char *>(&v), 4); }

static bool read_all_weights(model & m, std::ifstream & fin) {
    // record layout: int32 name_len, name, int32 n_dims, int32 dims[n_dims] (TF order), f32 data.
    // Unlike main.cpp:874-939 the loop ends at a clean EOF (no bogus trailing tensor).
    std::vector<float> staging;
    for (;;) {
        int32_t name_len = 0, n_dims = 0;
        if (!read_i32(fin, name_len)) break;
        if (name_len <= 0 || name_len > 4096) return false;
        std::string name((size_t)name_len, '\0');
        if (!fin.read(&name[0], name_len)) return false;
        if (!read_i32(fin, n_dims)) return false;
        // record-format extension (convert-tf-to-ggml.py:13-14 TODOs; weights.py): flags in the upper half of the n_dims word
        const bool disk_f16 = (n_dims & (1 << 16)) != 0;        // payload is f16
        const bool disk_t   = (n_dims & (1 << 17)) != 0;        // 2-D dense kernel stored (out, in)
        n_dims &= 0xFFFF;
        if (n_dims < 1 || n_dims > 4 || (disk_t && n_dims != 2)) return false;
        int32_t dims[4] = {1, 1, 1, 1};
        for (int i = 0; i < n_dims; i++)
            if (!read_i32(fin, dims[i]) || dims[i] <= 0) return false;
        if (disk_t) std::swap(dims[0], dims[1]);  // the tensor is created in the canonical (in, out) shape
        // main.cpp:887-889: every tensor whose name contains "convolution" is stored as F16
        const bool      is_f16 = name.find("convolution") != std::string::npos;
        const ggml_type type   = is_f16 ? GGML_TYPE_F16 : GGML_TYPE_F32;
        ggml_tensor *   t      = nullptr;  // ggml ne = TF dims reversed (main.cpp:905-917)
        switch (n_dims) {
            case 4: t = ggml_new_tensor_4d(m.ctx_w, type, dims[3], dims[2], dims[1], dims[0]); break;
            case 3: t = ggml_new_tensor_3d(m.ctx_w, GGML_TYPE_F32, dims[2], dims[1], dims[0]); break;
            case 2: t = ggml_new_tensor_2d(m.ctx_w, GGML_TYPE_F32, dims[1], dims[0]); break;
            default: t = ggml_new_tensor_1d(m.ctx_w, GGML_TYPE_F32, dims[0]); break;
        }
        const size_t count = (size_t)dims[0] * dims[1] * dims[2] * dims[3];
        staging.resize(count);
        if (disk_f16) {
            std::vector<ggml_fp16_t> half(count);
            if (!fin.read(reinterpret_cast<char *>(half.data()), (std::streamsize)(count * sizeof(ggml_fp16_t)))) return false;
            if (t->type == GGML_TYPE_F16 && !disk_t) {
                memcpy(t->data, half.data(), ggml_nbytes(t));  // the same bits the f32 file would have been rounded to
                ggml_set_name(t, name.size() > 60 ? name.c_str() + (name.size() - 60) : name.c_str());
                m.tensors[name] = t;
                m.total_weights += (int64_t)count;
                continue;
            }
            ggml_fp16_to_fp32_row(half.data(), staging.data(), (int)count);
        } else if (!fin.read(reinterpret_cast<char *>(staging.data()), (std::streamsize)(count * sizeof(float)))) {
            return false;
        }
        if (disk_t) {  // file rows are output features: back to (in, out), the layout the graph builder transposes in-graph
            std::vector<float> tr(count);
            const size_t n_in = (size_t)dims[0], n_out = (size_t)dims[1];
            for (size_t o = 0; o < n_out; o++)
                for (size_t i = 0; i < n_in; i++) tr[i * n_out + o] = staging[o * n_in + i];
            staging.swap(tr);
        }
        if (t->type == GGML_TYPE_F16) {
            ggml_fp32_to_fp16_row(staging.data(), (ggml_fp16_t *)t->data, (int)count);  // main.cpp:928-932
        } else {
            memcpy(t->data, staging.data(), ggml_nbytes(t));
        }
        ggml_set_name(t, name.size() > 60 ? name.c_str() + (name.size() - 60) : name.c_str());
        m.tensors[name] = t;
        m.total_weights += (int64_t)count;
    }
    return !m.tensors.empty();
}

static ggml_tensor * need(const model & m, const std::string & name) {
    auto it = m.tensors.find(name);
    if (it == m.tensors.end()) {  // main.cpp:225 tensors.at() -> std::out_of_range -> terminate
        fprintf(stderr, "mobilevit: weight file lacks tensor '%s'\n", name.c_str());
        abort();
    }
    return it->second;
}
static bool has(const model & m, const std::string & name) { return m.tensors.count(name) != 0; }

static void bind(const model & m, conv_layer & c, const std::string & path, bool use_normalization = true) {
    c.kernel = need(m, path + "/convolution/kernel:0");
    if (use_normalization) {
        c.gamma           = need(m, path + "/normalization/gamma:0");
        c.beta            = need(m, path + "/normalization/beta:0");
        c.moving_mean     = need(m, path + "/normalization/moving_mean:0");
        c.moving_variance = need(m, path + "/normalization/moving_variance:0");
    }
}
static void bind(const model & m, inverted_residual & r, const std::string & path, int strides) {
    r.strides = strides;
    bind(m, r.expand_1x1, path + "/expand_1x1");
    bind(m, r.conv_3x3, path + "/conv_3x3");
    bind(m, r.reduce_1x1, path + "/reduce_1x1");
}
static void bind(const model & m, transformer_layer & t, const std::string & p) {
    t.q_w    = need(m, p + "/attention/attention/query/kernel:0");  t.q_b    = need(m, p + "/attention/attention/query/bias:0");
    t.k_w    = need(m, p + "/attention/attention/key/kernel:0");    t.k_b    = need(m, p + "/attention/attention/key/bias:0");
    t.v_w    = need(m, p + "/attention/attention/value/kernel:0");  t.v_b    = need(m, p + "/attention/attention/value/bias:0");
    t.o_w    = need(m, p + "/attention/output/dense/kernel:0");     t.o_b    = need(m, p + "/attention/output/dense/bias:0");
    t.up_w   = need(m, p + "/intermediate/dense/kernel:0");         t.up_b   = need(m, p + "/intermediate/dense/bias:0");
    t.down_w = need(m, p + "/output/dense/kernel:0");               t.down_b = need(m, p + "/output/dense/bias:0");
    t.ln_before_g = need(m, p + "/layernorm_before/gamma:0");       t.ln_before_b = need(m, p + "/layernorm_before/beta:0");
    t.ln_after_g  = need(m, p + "/layernorm_after/gamma:0");        t.ln_after_b  = need(m, p + "/layernorm_after/beta:0");
}
static void bind(const model & m, vit_block & b, const std::string & path) {
    bind(m, b.downsampling, path + "/downsampling_layer", 2);  // main.cpp:397,434,471: ViT stages downsample by 2
    bind(m, b.conv_kxk, path + "/conv_kxk");
    bind(m, b.conv_1x1, path + "/conv_1x1", false);  // main.cpp:293: kernel only
    for (int j = 0; has(m, path + "/transformer/layer." + std::to_string(j) + "/attention/attention/query/kernel:0"); j++) {
        b.layers.emplace_back();
        bind(m, b.layers.back(), path + "/transformer/layer." + std::to_string(j));
    }
    b.ln_g = need(m, path + "/layernorm/gamma:0");
    b.ln_b = need(m, path + "/layernorm/beta:0");
    bind(m, b.conv_projection, path + "/conv_projection");
    bind(m, b.fusion, path + "/fusion");
}

bool model::load(const std::string & path) {
    std::ifstream fin(path, std::ios::binary);
    if (!fin) return false;  // main.cpp:316-318 only prints; a missing file is an error here
    fin.seekg(0, std::ios::end);
    const size_t file_bytes = (size_t)fin.tellg();
    fin.seekg(0, std::ios::beg);
    // weights arena: the reference uses a fixed 128 MiB (main.cpp:656); size it from the file instead
    ggml_init_params params = {file_bytes + (size_t)(8u << 20), nullptr, false};
    ctx_w                   = ggml_init(params);
    if (!ctx_w) return false;
    if (!read_all_weights(*this, fin)) return false;

    const std::string root = kRoot;
    bind(*this, conv_stem, root + "/conv_stem");
    // mobile_net_layer 1 and 2 (main.cpp:334-391): first block of layer.1 has stride 2, all others stride 1
    for (int i = 0; has(*this, root + "/encoder/layer.0/layer." + std::to_string(i) + "/expand_1x1/convolution/kernel:0"); i++) {
        layer_1.emplace_back();
        bind(*this, layer_1.back(), root + "/encoder/layer.0/layer." + std::to_string(i), 1);
    }
    for (int i = 0; has(*this, root + "/encoder/layer.1/layer." + std::to_string(i) + "/expand_1x1/convolution/kernel:0"); i++) {
        layer_2.emplace_back();
        bind(*this, layer_2.back(), root + "/encoder/layer.1/layer." + std::to_string(i), i == 0 ? 2 : 1);
    }
    bind(*this, layer_3, root + "/encoder/layer.2");
    bind(*this, layer_4, root + "/encoder/layer.3");
    bind(*this, layer_5, root + "/encoder/layer.4");
    bind(*this, conv_1x1_exp, root + "/conv_1x1_exp");
    // optional classification head (TFMobileViTForImageClassification's `classifier/{kernel,bias}:0`, kernel (in, out) like
    // every dense kernel of the file); the reference stops at the feature map (main.cpp:645), SURVEY 8f.1
    for (auto & kv : tensors) {
        const std::string & nm = kv.first;
        auto ends_with = [&](const char * suf) { const size_t n = strlen(suf); return nm.size() >= n && nm.compare(nm.size() - n, n, suf) == 0; };
        if (ends_with("classifier/kernel:0")) classifier_w = kv.second;
        if (ends_with("classifier/bias:0")) classifier_b = kv.second;
    }
    if (classifier_w && (!classifier_b || classifier_w->ne[1] != conv_1x1_exp.out_channels() || classifier_b->ne[0] != classifier_w->ne[0])) {
        fprintf(stderr, "mobilevit: classifier tensors do not fit the feature width\n");
        return false;
    }
    return true;
}

static void free_graph(forward_graph & g);
model::~model() {
    for (auto & kv : graphs) free_graph(kv.second);
    if (ctx_w) ggml_free(ctx_w);
}

// ------------------------------------------------------------------------------------------------------
// graph builder
// ------------------------------------------------------------------------------------------------------

// [1-d per-channel parameter] -> broadcast over an activation of shape like `like` (W,H,C,N)
static ggml_tensor * per_channel(ggml_context * ctx, ggml_tensor * p, ggml_tensor * like) {
    return ggml_repeat(ctx, ggml_cont_4d(ctx, p, 1, 1, p->ne[0], 1), like);
}

// conv (+ inference BatchNorm, eps 1e-5) (+ SiLU): main.cpp:771-852
ggml_tensor * conv_layer::forward(ggml_context * ctx, ggml_tensor * input, int stride, bool use_normalization,
                                  bool use_activation, bool depthwise) const {
    const int64_t oc = kernel->ne[0], ic = kernel->ne[1], kw = kernel->ne[2], kh = kernel->ne[3];
    const int     pad = (int)(kw - 1) / 2;
    // file layout (KH,KW,IC,OC) == ggml ne (OC,IC,KW,KH); ggml_conv_2d wants (KW,KH,IC,OC)
    ggml_tensor * w   = ggml_cont_4d(ctx, ggml_permute(ctx, kernel, 3, 2, 0, 1), kw, kh, ic, oc);
    ggml_tensor * out = depthwise ? ggml_conv_depthwise_2d(ctx, w, input, stride, stride, pad, pad, 1, 1)
                                  : ggml_conv_2d(ctx, w, input, stride, stride, pad, pad, 1, 1);
    if (use_normalization) {
        // ((x - mean) / sqrt(var + eps)) * gamma + beta, unfused f32 as in main.cpp:809-846
        out = ggml_sub(ctx, out, per_channel(ctx, moving_mean, out));
        out = ggml_div(ctx, out, ggml_sqrt(ctx, ggml_add(ctx, per_channel(ctx, moving_variance, out), ggml_new_f32(ctx, 1.0e-5f))));
        out = ggml_mul(ctx, out, per_channel(ctx, gamma, out));
        out = ggml_add(ctx, out, per_channel(ctx, beta, out));
    }
    if (use_activation) out = ggml_silu(ctx, out);
    return out;
}

// MobileNetV2 block: main.cpp:854-870.  The residual rule uses the real channel counts.
ggml_tensor * inverted_residual::forward(ggml_context * ctx, ggml_tensor * inp) const {
    ggml_tensor * x = expand_1x1.forward(ctx, inp, 1, true, true, false);
    x               = conv_3x3.forward(ctx, x, strides, true, true, true);
    x               = reduce_1x1.forward(ctx, x, 1, true, false, false);
    if (strides == 1 && expand_1x1.in_channels() == reduce_1x1.out_channels()) x = ggml_add(ctx, x, inp);
    return x;
}

// (W,H,C,N) -> (C, n_patches, ps*ps, N): main.cpp:721-747 with the batch carried in the slowest dim
ggml_tensor * unfolding(ggml_context * ctx, ggml_tensor * features, int ps) {
    const int64_t nw = features->ne[0], nh = features->ne[1], c = features->ne[2], n = features->ne[3];
    GGML_ASSERT(nw % ps == 0);
    GGML_ASSERT(nh % ps == 0);
    const int64_t npw = nw / ps, nph = nh / ps;
    ggml_tensor * p = ggml_reshape_4d(ctx, ggml_cont(ctx, features), ps, npw, ps, n * c * nph);
    p               = ggml_permute(ctx, p, 0, 2, 1, 3);  // (pw, ph, npw, n*c*nph)
    p               = ggml_reshape_4d(ctx, ggml_cont(ctx, p), ps * ps, npw * nph, c, n);
    return ggml_cont(ctx, ggml_permute(ctx, p, 2, 1, 0, 3));  // (c, n_patches, ps*ps, n)
}

// (C, n_patches, ps*ps*N) -> (W,H,C,N): main.cpp:750-768 (which assumes a square map and N=1)
ggml_tensor * folding(ggml_context * ctx, ggml_tensor * tokens, int ps, int npw, int nph, int batch) {
    const int64_t c = tokens->ne[0], n_patches = tokens->ne[1];
    GGML_ASSERT(n_patches == (int64_t)npw * nph);
    ggml_tensor * f = ggml_reshape_4d(ctx, tokens, c, n_patches, ps * ps, batch);
    f               = ggml_cont(ctx, ggml_permute(ctx, f, 2, 1, 0, 3));  // (ps*ps, n_patches, c, n)
    f               = ggml_reshape_4d(ctx, f, ps, ps, npw, (int64_t)nph * c * batch);
    f               = ggml_cont(ctx, ggml_permute(ctx, f, 0, 2, 1, 3));  // (ps, npw, ps, nph*c*n)
    return ggml_reshape_4d(ctx, f, (int64_t)ps * npw, (int64_t)ps * nph, c, batch);
}

static ggml_tensor * layer_norm(ggml_context * ctx, ggml_tensor * x, ggml_tensor * g, ggml_tensor * b, float eps) {
    // main.cpp:1002-1019: norm * gamma + beta, then cont
    ggml_tensor * y = ggml_mul(ctx, ggml_norm(ctx, x, eps), ggml_repeat(ctx, ggml_reshape_3d(ctx, g, g->ne[0], 1, 1), x));
    y               = ggml_add(ctx, y, ggml_repeat(ctx, ggml_reshape_3d(ctx, b, b->ne[0], 1, 1), x));
    return ggml_cont(ctx, y);
}

static ggml_tensor * dense(ggml_context * ctx, ggml_tensor * x, ggml_tensor * w, ggml_tensor * b) {
    // file kernel is (in,out) == ggml ne (out,in); mul_mat wants K=in fastest: main.cpp:1022-1035
    ggml_tensor * y = ggml_mul_mat(ctx, ggml_cont(ctx, ggml_permute(ctx, w, 1, 0, 2, 3)), x);
    return ggml_add(ctx, y, ggml_repeat(ctx, ggml_reshape_3d(ctx, b, b->ne[0], 1, 1), y));
}

static ggml_tensor * split_heads(ggml_context * ctx, ggml_tensor * x, int heads) {  // transpose_for_score, main.cpp:975-986
    GGML_ASSERT(x->ne[0] % heads == 0);
    x = ggml_reshape_4d(ctx, x, x->ne[0] / heads, heads, x->ne[1], x->ne[2]);
    return ggml_permute(ctx, x, 0, 2, 1, 3);  // (d, L, heads, B)
}

// pre-LN transformer layer: main.cpp:988-1172.  hidden: (C, L, B) with B = patch_area * N
ggml_tensor * transformer_layer::forward(ggml_context * ctx, ggml_tensor * hidden, float eps, int heads) const {
    const float   scale = sqrtf(1.0f * (float)(hidden->ne[0] / heads));  // main.cpp:999, used as a divisor (:1076)
    ggml_tensor * x     = layer_norm(ctx, hidden, ln_before_g, ln_before_b, eps);
    ggml_tensor * k     = split_heads(ctx, dense(ctx, x, k_w, k_b), heads);
    ggml_tensor * v     = split_heads(ctx, dense(ctx, x, v_w, v_b), heads);
    ggml_tensor * q     = split_heads(ctx, dense(ctx, x, q_w, q_b), heads);
    ggml_tensor * score = ggml_soft_max(ctx, ggml_div(ctx, ggml_mul_mat(ctx, k, q), ggml_new_f32(ctx, scale)));  // (L,L,h,B)
    ggml_tensor * att   = ggml_mul_mat(ctx, ggml_cont(ctx, ggml_permute(ctx, v, 1, 0, 2, 3)), score);             // (d,L,h,B)
    att                 = ggml_cont(ctx, ggml_permute(ctx, att, 0, 2, 1, 3));                                       // (d,h,L,B)
    att                 = ggml_reshape_3d(ctx, att, hidden->ne[0], hidden->ne[1], hidden->ne[2]);
    hidden              = ggml_add(ctx, hidden, dense(ctx, att, o_w, o_b));  // main.cpp:1095-1111
    ggml_tensor * y     = layer_norm(ctx, hidden, ln_after_g, ln_after_b, eps);
    y                   = ggml_silu(ctx, dense(ctx, y, up_w, up_b));     // main.cpp:1134-1148
    y                   = dense(ctx, y, down_w, down_b);                 // main.cpp:1151-1163
    return ggml_add(ctx, y, hidden);                                     // main.cpp:1165
}

// main.cpp:1174-1223
ggml_tensor * vit_block::forward(ggml_context * ctx, ggml_tensor * inp, const hparams & hp) const {
    ggml_tensor * residual = downsampling.forward(ctx, inp);
    ggml_tensor * x        = conv_kxk.forward(ctx, residual, 1, true, true, false);
    x                      = conv_1x1.forward(ctx, x, 1, false, false, false);
    const int ps = hp.patch_size;
    const int npw = (int)x->ne[0] / ps, nph = (int)x->ne[1] / ps, batch = (int)x->ne[3];
    x = unfolding(ctx, x, ps);                                              // (C, L, ps*ps, N)
    x = ggml_reshape_3d(ctx, x, x->ne[0], x->ne[1], x->ne[2] * x->ne[3]);   // (C, L, ps*ps*N)
    for (const transformer_layer & l : layers) x = l.forward(ctx, x, hp.layer_norm_eps, hp.num_attention_heads);
    x = layer_norm(ctx, x, ln_g, ln_b, hp.layer_norm_eps);
    x = folding(ctx, x, ps, npw, nph, batch);
    x = conv_projection.forward(ctx, x, 1, true, true, false);
    return fusion.forward(ctx, ggml_concat(ctx, residual, x), 1, true, true, false);
}

// extract_features' graph (main.cpp:604-646) for a batch.  `images_hwc` has ne=(3,W,H,N): the HWC->CHW copy
// the reference does on the host (main.cpp:627-634) is a permute+cont in the graph here.
ggml_tensor * model::build_forward(ggml_context * ctx, ggml_tensor * images_hwc, ggml_tensor ** pooled,
                                   std::vector<ggml_tensor *> * stages) const {
    auto tap = [&](ggml_tensor * t) { if (stages) stages->push_back(t); };
    ggml_tensor * x = ggml_cont(ctx, ggml_permute(ctx, images_hwc, 2, 0, 1, 3));  // (W,H,3,N)
    x               = conv_stem.forward(ctx, x, 2, true, true, false);
    tap(x);
    for (const inverted_residual & r : layer_1) x = r.forward(ctx, x);
    tap(x);
    for (const inverted_residual & r : layer_2) x = r.forward(ctx, x);
    tap(x);
    x = layer_3.forward(ctx, x, hp);
    tap(x);
    x = layer_4.forward(ctx, x, hp);
    tap(x);
    x = layer_5.forward(ctx, x, hp);
    tap(x);
    x = conv_1x1_exp.forward(ctx, x, 1, true, true, false);
    tap(x);
    if (pooled) *pooled = ggml_b200_pool_mean_hw(ctx, x);
    return x;
}

// Concurrent lanes (MVIT_LANES=S, default 1): a request of n images runs as S sub-batches on S streams at the same time.  Built to
// attack the launch-latency chain at small per-GPU batches (strong scaling: 32 images per GPU at 8 GPUs).  Per-image results do not
// depend on the batch they are computed in (bit for bit: the batch-independence tests), so a split changes nothing but the time --
// and measured on B200 it does not help: 32 images take 1.18 / 1.22 / 1.43 ms as 1 / 2 / 4 lanes (64: 1.77 / 1.80 / 2.05); with the
// fused transformer stages of round 2 (49 launches): 1.00 / 1.06 / 1.25 ms (64: 1.56 / 1.58 / 1.90), tests/lanes_probe.sh.  The
// persistent kernels of one lane already occupy every SM's shared memory and TMEM, so the lanes time-share the SMs instead of filling
// each other's gaps.  Kept as an option (and tested); off by default.
static int choose_lanes(int n, int h, int w) {
    (void)h; (void)w;
    int s = 1;
    if (const char * e = getenv("MVIT_LANES")) s = atoi(e);
    while (s > 1 && n % s) s--;
    return s < 1 ? 1 : s;
}

static void build_graph_body(const model & m, forward_graph & g, int n, int h, int w, bool debug_stages) {
    g.gf        = ggml_new_graph(g.ctx);
    g.input_hwc = ggml_new_tensor_4d(g.ctx, GGML_TYPE_F32, 3, w, h, n);
    ggml_set_name(g.input_hwc, "inp");
    ggml_set_input(g.input_hwc);
    g.features = m.build_forward(g.ctx, g.input_hwc, &g.pooled, &g.stages);
    ggml_set_name(g.features, "features");
    ggml_set_name(g.pooled, "pooled");
    ggml_build_forward_expand(g.gf, g.features);
    ggml_build_forward_expand(g.gf, g.pooled);
    if (m.classifier_w) {  // logits = pooled . kernel + bias
        g.logits = dense(g.ctx, ggml_reshape_2d(g.ctx, g.pooled, g.pooled->ne[2], g.pooled->ne[3]), m.classifier_w, m.classifier_b);
        ggml_set_name(g.logits, "logits");
        ggml_build_forward_expand(g.gf, g.logits);
    }
    if (debug_stages)
        for (ggml_tensor * t : g.stages) ggml_build_forward_expand(g.gf, t);  // marks them as outputs -> host shadows
}

forward_graph & model::graph_for(int n, int h, int w, int slot) {
    auto key = std::make_tuple(n, h, w, slot);
    auto it  = graphs.find(key);
    if (it != graphs.end()) return it->second;
    forward_graph g;
    const int    C         = conv_1x1_exp.out_channels();
    const size_t in_bytes  = (size_t)n * h * w * 3 * sizeof(float);
    size_t out_bytes = (size_t)n * C * ((size_t)(h / 32) * (w / 32) + 1) * sizeof(float);
    if (classifier_w) out_bytes += (size_t)n * (size_t)classifier_w->ne[0] * sizeof(float);
    const char * dbg = getenv("MVIT_DEBUG_STAGES");
    const bool   debug_stages = dbg && atoi(dbg) > 0;
    if (debug_stages) out_bytes += (size_t)n * h * w * 16 * sizeof(float);  // all stage taps together are < 16 floats per input pixel
    // the compute arena only holds tensor records, the input staging area and the output shadows:
    // intermediates live in the device plan's arena (the reference needs 1 GiB per image, main.cpp:605)
    const size_t arena_bytes = in_bytes + out_bytes + (size_t)(24u << 20);
    g.pinned_arena          = ggml_b200_host_malloc(arena_bytes);  // NULL without a GPU: ggml_init then mallocs
    ggml_init_params params = {arena_bytes, g.pinned_arena, false};
    g.ctx                   = ggml_init(params);
    GGML_ASSERT(g.ctx != nullptr);
    const int S = debug_stages ? 1 : choose_lanes(n, h, w);
    if (S == 1) {
        build_graph_body(*this, g, n, h, w, debug_stages);
        return graphs.emplace(key, g).first->second;
    }
    // Split request: this record only owns the caller-visible host buffers (pinned): the input images and the outputs are plain
    // leaf tensors here; every lane is a complete forward graph for n/S images whose input leaf and output shadows ALIAS the
    // lane's slice of those buffers (the batch is the slowest dimension of every one of them), so uploads and downloads of the
    // lanes go straight from / to the caller's arrays.
    g.input_hwc = ggml_new_tensor_4d(g.ctx, GGML_TYPE_F32, 3, w, h, n);
    g.features  = ggml_new_tensor_4d(g.ctx, GGML_TYPE_F32, w / 32, h / 32, C, n);
    g.pooled    = ggml_new_tensor_4d(g.ctx, GGML_TYPE_F32, 1, 1, C, n);
    if (classifier_w) g.logits = ggml_new_tensor_2d(g.ctx, GGML_TYPE_F32, classifier_w->ne[0], n);
    const int nl = n / S;
    for (int l = 0; l < S; l++) {
        forward_graph       lane;
        ggml_init_params lp = {(size_t)(24u << 20), nullptr, true};  // tensor records only: no_alloc
        lane.ctx            = ggml_init(lp);
        GGML_ASSERT(lane.ctx != nullptr);
        build_graph_body(*this, lane, nl, h, w, false);
        auto slice = [&](ggml_tensor * whole, ggml_tensor * part) { part->data = (char *)whole->data + (size_t)l * ggml_nbytes(part); };
        slice(g.input_hwc, lane.input_hwc);
        slice(g.features, lane.features);
        slice(g.pooled, lane.pooled);
        if (g.logits) slice(g.logits, lane.logits);
        g.lanes.push_back(lane);
    }
    return graphs.emplace(key, g).first->second;
}

static void free_graph(forward_graph & g) {
    for (forward_graph & l : g.lanes) {
        ggml_graph_release_plan(l.gf);
        ggml_free(l.ctx);
    }
    if (g.gf) ggml_graph_release_plan(g.gf);
    ggml_free(g.ctx);
    ggml_b200_host_free(g.pinned_arena);
    ggml_b200_host_free(g.input_u8);
}

void model::release(int n, int h, int w) {
    for (auto it = graphs.begin(); it != graphs.end();) {
        if (std::get<0>(it->first) == n && std::get<1>(it->first) == h && std::get<2>(it->first) == w) {
            free_graph(it->second);
            it = graphs.erase(it);
        } else {
            ++it;
        }
    }
}

}  // namespace mvit

// ------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------
struct mvit_model {
    mvit::model m;
};

// H and W must be multiples of 64: the network downsamples by 32 and the last ViT block unfolds 2x2 patches
// (the reference asserts this at main.cpp:729-730 for its fixed 256x256 input)
static bool shape_ok(int n, int h, int w) { return n > 0 && h > 0 && w > 0 && h % 64 == 0 && w % 64 == 0; }

extern "C" mvit_model * mvit_load(const char * path) {
    mvit_model * h = new mvit_model();
    if (!h->m.load(path)) {
        delete h;
        return nullptr;
    }
    return h;
}
extern "C" void    mvit_free(mvit_model * m) { delete m; }
extern "C" int     mvit_num_tensors(const mvit_model * m) { return (int)m->m.tensors.size(); }
extern "C" int64_t mvit_num_weights(const mvit_model * m) { return m->m.total_weights; }
extern "C" int     mvit_out_channels(const mvit_model * m) { return m->m.conv_1x1_exp.out_channels(); }
extern "C" int     mvit_num_classes(const mvit_model * m) { return m->m.classifier_w ? (int)m->m.classifier_w->ne[0] : 0; }

// ---- running one request: a single graph, or its concurrent lanes ------------------------------------------------------------
static void prepare_graph(mvit::forward_graph & g, bool own_stream) {
    if (g.lanes.empty()) {
        if (!g.gf->plan) {
            ggml_b200_graph_prepare(g.ctx, g.gf);
            if (own_stream) ggml_b200_graph_use_private_stream(g.gf);
        }
        return;
    }
    for (mvit::forward_graph & l : g.lanes)
        if (!l.gf->plan) {
            ggml_b200_graph_prepare(l.ctx, l.gf);
            ggml_b200_graph_use_private_stream(l.gf);
        }
}
// u8_* != 0: the images come as raw u8 [n][src_h][src_w][3] in g.input_u8 and are preprocessed on the device.
// owner_is_current_stream: synchronous entry points order the work after / before the library's current stream; pipelined slots
// run on streams of their own (the slot graph's private stream, or lane 0's).
static int run_graph(mvit::forward_graph & g, bool upload, bool download, bool wait, bool owner_is_current_stream, int u8_n = 0, int src_h = 0, int src_w = 0) {
    prepare_graph(g, !owner_is_current_stream);
    if (g.lanes.empty()) {
        if (u8_n) {
            if (ggml_b200_graph_upload_u8_images_fused(g.gf, g.input_hwc, g.input_u8, u8_n, src_h, src_w)) return 1;
            upload = false;
        }
        ggml_b200_graph_set_transfers(g.gf, upload, download);
        if (wait) ggml_graph_compute_with_ctx(g.ctx, g.gf, 1);
        else ggml_b200_graph_compute_async(g.ctx, g.gf);
        return 0;
    }
    std::vector<ggml_cgraph *> gfs;
    for (mvit::forward_graph & l : g.lanes) gfs.push_back(l.gf);
    const int S = (int)gfs.size();
    ggml_b200_graph_group_begin(gfs.data(), S, owner_is_current_stream ? 1 : 0);
    for (int l = 0; l < S; l++) {
        mvit::forward_graph & L = g.lanes[(size_t)l];
        bool up = upload;
        if (u8_n) {
            const int nl = u8_n / S;
            if (ggml_b200_graph_upload_u8_images_fused(L.gf, L.input_hwc, g.input_u8 + (size_t)l * nl * src_h * src_w * 3, nl, src_h, src_w)) return 1;
            up = false;
        }
        ggml_b200_graph_set_transfers(L.gf, up, download);
        ggml_b200_graph_compute_async(L.ctx, L.gf);
    }
    ggml_b200_graph_group_end(gfs.data(), S, owner_is_current_stream ? 1 : 0);
    if (wait) {
        if (owner_is_current_stream) ggml_b200_synchronize();
        else ggml_b200_graph_wait(gfs[0]);
    }
    return 0;
}
static void wait_graph(mvit::forward_graph & g) {
    if (g.lanes.empty()) ggml_b200_graph_wait(g.gf);
    else if (g.lanes[0].gf->plan) ggml_b200_graph_wait(g.lanes[0].gf);  // lane 0's stream is the owner: it waited for its siblings
}

extern "C" int mvit_compute(mvit_model * m, int n, int h, int w);

extern "C" int mvit_extract_features(mvit_model * m, const float * images_hwc, int n, int h, int w, float * features,
                                     float * pooled) {
    if (!m || !images_hwc || !shape_ok(n, h, w)) return 1;
    mvit::forward_graph & g = m->m.graph_for(n, h, w);
    memcpy(ggml_get_data(g.input_hwc), images_hwc, ggml_nbytes(g.input_hwc));
    mvit_compute(m, n, h, w);
    if (features) memcpy(features, ggml_get_data(g.features), ggml_nbytes(g.features));
    if (pooled) memcpy(pooled, ggml_get_data(g.pooled), ggml_nbytes(g.pooled));
    return 0;
}

extern "C" float * mvit_host_input(mvit_model * m, int n, int h, int w) {
    if (!m || !shape_ok(n, h, w)) return nullptr;
    return (float *)ggml_get_data(m->m.graph_for(n, h, w).input_hwc);
}
extern "C" int mvit_compute(mvit_model * m, int n, int h, int w) {
    if (!m || !shape_ok(n, h, w)) return 1;
    return run_graph(m->m.graph_for(n, h, w), true, true, true, true);
}
extern "C" const float * mvit_host_features(mvit_model * m, int n, int h, int w) {
    if (!m || !shape_ok(n, h, w)) return nullptr;
    return (const float *)ggml_get_data(m->m.graph_for(n, h, w).features);
}
extern "C" const float * mvit_host_pooled(mvit_model * m, int n, int h, int w) {
    if (!m || !shape_ok(n, h, w)) return nullptr;
    return (const float *)ggml_get_data(m->m.graph_for(n, h, w).pooled);
}
extern "C" const float * mvit_host_logits(mvit_model * m, int n, int h, int w) {
    if (!m || !shape_ok(n, h, w) || !m->m.classifier_w) return nullptr;
    return (const float *)ggml_get_data(m->m.graph_for(n, h, w).logits);
}
// classify: images -> class logits [n][classes] and/or top-1 class ids; 2 when the weight file has no classifier
extern "C" int mvit_classify(mvit_model * m, const float * images_hwc, int n, int h, int w, float * logits, int32_t * top1) {
    if (!m || !images_hwc || !shape_ok(n, h, w)) return 1;
    if (!m->m.classifier_w) return 2;
    mvit::forward_graph & g = m->m.graph_for(n, h, w);
    memcpy(ggml_get_data(g.input_hwc), images_hwc, ggml_nbytes(g.input_hwc));
    mvit_compute(m, n, h, w);
    const float * lg = (const float *)ggml_get_data(g.logits);
    const int     nc = (int)g.logits->ne[0];
    if (logits) memcpy(logits, lg, ggml_nbytes(g.logits));
    if (top1)
        for (int i = 0; i < n; i++) top1[i] = (int32_t)(std::max_element(lg + (size_t)i * nc, lg + (size_t)(i + 1) * nc) - (lg + (size_t)i * nc));
    return 0;
}
// per-launch profile of the plan (of lane 0 for a split request: n / lanes images)
extern "C" int mvit_profile_json(mvit_model * m, int n, int h, int w, int reps, char * buf, size_t cap) {
    if (mvit_prepare(m, n, h, w)) return -1;
    mvit::forward_graph & g = m->m.graph_for(n, h, w);
    return ggml_b200_graph_profile_json(g.lanes.empty() ? g.gf : g.lanes[0].gf, reps, buf, cap);
}

// debug: copy stage tap `idx` (0 stem, 1..5 layers, 6 exp) as [N][C][H][W] floats; needs MVIT_DEBUG_STAGES=1 and a compute
extern "C" int64_t mvit_debug_stage(mvit_model * m, int n, int h, int w, int idx, float * out, int64_t cap_floats, int64_t * ne4) {
    if (!m || !shape_ok(n, h, w)) return -1;
    mvit::forward_graph & g = m->m.graph_for(n, h, w);
    if (idx < 0 || idx >= (int)g.stages.size()) return -1;
    ggml_tensor * t = g.stages[idx];
    for (int i = 0; i < 4; i++) ne4[i] = t->ne[i];
    const int64_t cnt = ggml_nelements(t);
    if (!t->data || cnt > cap_floats) return -2;
    memcpy(out, t->data, (size_t)cnt * sizeof(float));
    return cnt;
}

// ---- pipelined slots: slot s has its own input buffer, output shadows, device arena and stream ----
static void wait_graph_if_planned(mvit::forward_graph & g) {
    if (g.lanes.empty()) { if (g.gf->plan) ggml_b200_graph_wait(g.gf); }
    else wait_graph(g);
}
static mvit::forward_graph * slot_graph(mvit_model * m, int n, int h, int w, int slot) {
    if (!m || !shape_ok(n, h, w) || slot < 0 || slot > 7) return nullptr;
    mvit::forward_graph & g = m->m.graph_for(n, h, w, 1 + slot);  // slot graphs are distinct from the synchronous one (0)
    prepare_graph(g, true);
    return &g;
}
extern "C" float * mvit_slot_input(mvit_model * m, int n, int h, int w, int slot) {
    mvit::forward_graph * g = slot_graph(m, n, h, w, slot);
    return g ? (float *)ggml_get_data(g->input_hwc) : nullptr;
}
extern "C" int mvit_slot_set_transfers(mvit_model * m, int n, int h, int w, int slot, int upload_inputs, int download_outputs) {
    mvit::forward_graph * g = slot_graph(m, n, h, w, slot);
    if (!g) return 1;
    g->slot_upload   = upload_inputs != 0;
    g->slot_download = download_outputs != 0;
    return 0;
}
extern "C" int mvit_slot_submit(mvit_model * m, int n, int h, int w, int slot) {
    mvit::forward_graph * g = slot_graph(m, n, h, w, slot);
    if (!g) return 1;
    return run_graph(*g, g->slot_upload, g->slot_download, false, false);
}
extern "C" int mvit_slot_wait(mvit_model * m, int n, int h, int w, int slot) {
    mvit::forward_graph * g = slot_graph(m, n, h, w, slot);
    if (!g) return 1;
    wait_graph(*g);
    return 0;
}
extern "C" const float * mvit_slot_features(mvit_model * m, int n, int h, int w, int slot) {
    mvit::forward_graph * g = slot_graph(m, n, h, w, slot);
    return g ? (const float *)ggml_get_data(g->features) : nullptr;
}
extern "C" const float * mvit_slot_logits(mvit_model * m, int n, int h, int w, int slot) {
    mvit::forward_graph * g = slot_graph(m, n, h, w, slot);
    return g && g->logits ? (const float *)ggml_get_data(g->logits) : nullptr;
}
extern "C" const float * mvit_slot_pooled(mvit_model * m, int n, int h, int w, int slot) {
    mvit::forward_graph * g = slot_graph(m, n, h, w, slot);
    return g ? (const float *)ggml_get_data(g->pooled) : nullptr;
}

// ---- u8 images, preprocessing on the device ----
static uint8_t * u8_staging(mvit::forward_graph & g, int n, int src_h, int src_w) {
    const size_t bytes = (size_t)n * src_h * src_w * 3;
    if (g.input_u8_bytes < bytes) {
        wait_graph_if_planned(g);  // a submitted copy may still read the old buffer
        ggml_b200_host_free(g.input_u8);
        g.input_u8       = (uint8_t *)ggml_b200_host_malloc(bytes);
        g.input_u8_bytes = g.input_u8 ? bytes : 0;
    }
    return g.input_u8;
}
static int submit_u8(mvit::forward_graph & g, int n, int src_h, int src_w, bool wait, bool owner_is_current_stream) {
    if (!g.input_u8 || g.input_u8_bytes < (size_t)n * src_h * src_w * 3) return 1;
    return run_graph(g, false, true, wait, owner_is_current_stream, n, src_h, src_w);
}
extern "C" uint8_t * mvit_host_input_u8(mvit_model * m, int n, int h, int w, int src_h, int src_w) {
    if (!m || !shape_ok(n, h, w) || src_h <= 0 || src_w <= 0) return nullptr;
    return u8_staging(m->m.graph_for(n, h, w), n, src_h, src_w);
}
extern "C" int mvit_compute_u8(mvit_model * m, int n, int h, int w, int src_h, int src_w) {
    if (!m || !shape_ok(n, h, w) || src_h <= 0 || src_w <= 0) return 1;
    return submit_u8(m->m.graph_for(n, h, w), n, src_h, src_w, true, true);
}
extern "C" uint8_t * mvit_slot_input_u8(mvit_model * m, int n, int h, int w, int slot, int src_h, int src_w) {
    mvit::forward_graph * g = slot_graph(m, n, h, w, slot);
    return g && src_h > 0 && src_w > 0 ? u8_staging(*g, n, src_h, src_w) : nullptr;
}
extern "C" int mvit_slot_submit_u8(mvit_model * m, int n, int h, int w, int slot, int src_h, int src_w) {
    mvit::forward_graph * g = slot_graph(m, n, h, w, slot);
    return g && src_h > 0 && src_w > 0 ? submit_u8(*g, n, src_h, src_w, false, false) : 1;
}
extern "C" int mvit_preprocess_u8(mvit_model * m, const uint8_t * images, int n, int src_h, int src_w, int h, int w, float * out_hwc) {
    if (!m || !images || !out_hwc || !shape_ok(n, h, w) || src_h <= 0 || src_w <= 0) return 1;
    mvit::forward_graph & g = m->m.graph_for(n, h, w);
    uint8_t * st = u8_staging(g, n, src_h, src_w);
    if (!st) return 1;
    memcpy(st, images, (size_t)n * src_h * src_w * 3);
    prepare_graph(g, false);
    if (g.lanes.empty()) {
        if (ggml_b200_graph_upload_u8_images(g.gf, g.input_hwc, st, n, src_h, src_w)) return 1;
        return ggml_b200_tensor_download(g.gf, g.input_hwc, out_hwc);
    }
    const int S = (int)g.lanes.size(), nl = n / S;
    for (int l = 0; l < S; l++) {
        mvit::forward_graph & L = g.lanes[(size_t)l];
        if (ggml_b200_graph_upload_u8_images(L.gf, L.input_hwc, st + (size_t)l * nl * src_h * src_w * 3, nl, src_h, src_w)) return 1;
        if (ggml_b200_tensor_download(L.gf, L.input_hwc, out_hwc + (size_t)l * nl * h * w * 3)) return 1;
    }
    return 0;
}

extern "C" int mvit_prepare(mvit_model * m, int n, int h, int w) {
    if (!m || !shape_ok(n, h, w)) return 1;
    prepare_graph(m->m.graph_for(n, h, w), false);
    return 0;
}
// device pointers of the synchronous graph's buffers; NULL for a request that is split into lanes (there is no single buffer)
extern "C" void * mvit_device_input(mvit_model * m, int n, int h, int w) {
    if (mvit_prepare(m, n, h, w)) return nullptr;
    mvit::forward_graph & g = m->m.graph_for(n, h, w);
    return g.lanes.empty() ? ggml_b200_tensor_get_device_data(g.gf, g.input_hwc) : nullptr;
}
extern "C" void * mvit_device_features(mvit_model * m, int n, int h, int w) {
    if (mvit_prepare(m, n, h, w)) return nullptr;
    mvit::forward_graph & g = m->m.graph_for(n, h, w);
    return g.lanes.empty() ? ggml_b200_tensor_get_device_data(g.gf, g.features) : nullptr;
}
extern "C" void * mvit_device_pooled(mvit_model * m, int n, int h, int w) {
    if (mvit_prepare(m, n, h, w)) return nullptr;
    mvit::forward_graph & g = m->m.graph_for(n, h, w);
    return g.lanes.empty() ? ggml_b200_tensor_get_device_data(g.gf, g.pooled) : nullptr;
}
extern "C" int mvit_forward_device(mvit_model * m, int n, int h, int w) {
    if (!m || !shape_ok(n, h, w)) return 1;
    mvit::forward_graph & g = m->m.graph_for(n, h, w);
    return run_graph(g, false, false, g.lanes.empty(), true);  // lanes: asynchronous on the current stream (the caller synchronises it)
}
extern "C" void mvit_release(mvit_model * m, int n, int h, int w) {
    if (m) m->m.release(n, h, w);
}
extern "C" int mvit_plan_info(mvit_model * m, int n, int h, int w, struct mvit_plan_info * out) {
    if (mvit_prepare(m, n, h, w)) return 1;
    mvit::forward_graph & g = m->m.graph_for(n, h, w);
    memset(out, 0, sizeof(*out));
    out->cuda_graph = 1;
    const size_t S = g.lanes.empty() ? 1 : g.lanes.size();
    for (size_t l = 0; l < S; l++) {
        ggml_b200_plan_stats s;
        ggml_b200_graph_plan_stats(g.lanes.empty() ? g.gf : g.lanes[l].gf, &s);
        out->mode         = s.mode;
        out->graph_nodes  = s.n_graph_nodes;
        out->launches    += s.n_launches;
        out->arena_bytes += s.arena_bytes;
        out->naive_bytes += s.naive_bytes;
        out->weight_bytes = s.weight_bytes;
        out->cuda_graph   = out->cuda_graph && s.used_cuda_graph;
    }
    out->lanes = (int)S;
    return 0;
}
