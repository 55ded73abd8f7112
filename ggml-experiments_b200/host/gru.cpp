// gru.cpp -- batched GRU text generator on the ggml boundary (include/gru_b200.h).
// Mirrors /root/reference/rnn_text_gen/rnn_text_generation.cpp: same weight file, same cell arithmetic (Keras GRU with
// reset_after=True, gate order z, r, h; sigmoid written as silu(x)/x like rnn.cpp:51-55), but for B streams at once and with the
// generation loop on the device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "ggml/ggml.h"
#include "gru_b200.h"

extern "C" void * ggml_b200_get_stream(void);

namespace {

struct cell_graph {
    ggml_context * ctx = nullptr;
    ggml_cgraph *  gf  = nullptr;
    ggml_tensor *  ids = nullptr, * states = nullptr, * next_ids = nullptr, * new_states = nullptr, * logits = nullptr;
    void *         pinned = nullptr;  // page-locked backing store of ctx: the step-0 upload of ids + state is a plain DMA
};

}  // namespace

struct gru_model {
    ggml_context * ctx_w = nullptr;
    // leafs, ggml ne order.  Kernels are stored transposed on the host once (rnn.cpp:129-158 does it with a graph per run)
    ggml_tensor * embeddings = nullptr;   // (E, V)
    ggml_tensor * kernel_t = nullptr;     // (E, 3U)   : mul_mat contracts over E
    ggml_tensor * recurrent_t = nullptr;  // (U, 3U)
    ggml_tensor * bias_x = nullptr, * bias_h = nullptr;  // (3U, 1) each: rows 0 / 1 of the (2, 3U) Keras bias
    ggml_tensor * dense_t = nullptr;      // (U, V)
    ggml_tensor * dense_bias = nullptr;   // (V, 1)
    int E = 0, U = 0, V = 0;
    std::map<int, cell_graph> graphs;     // per batch size
};

static bool read_block(FILE * f, int n_header_ints, std::vector<float> & dst, size_t count) {
    int32_t dummy;
    for (int i = 0; i < n_header_ints; i++)
        if (fread(&dummy, 4, 1, f) != 1) return false;  // rnn.cpp:123 skips n_dims + dims
    dst.resize(count);
    return fread(dst.data(), sizeof(float), count, f) == count;
}

// file (rows, cols) row-major  ->  ggml (rows, cols) tensor holding the transpose: element (i0 = r, i1 = c) = file[r][c]
static ggml_tensor * transposed_leaf(ggml_context * ctx, const std::vector<float> & src, int rows, int cols) {
    ggml_tensor * t = ggml_new_tensor_2d(ctx, GGML_TYPE_F32, rows, cols);
    float *       d = (float *)t->data;
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < cols; c++) d[(size_t)c * rows + r] = src[(size_t)r * cols + c];
    return t;
}

extern "C" gru_model * gru_load(const char * path) {
    FILE * f = fopen(path, "rb");
    if (!f) return nullptr;
    const int V = 66, E = 256, U = 1024;  // rnn.cpp:108-115
    std::vector<float> emb, W, Um, b, D, c;
    bool ok = read_block(f, 3, emb, (size_t)V * E) && read_block(f, 3, W, (size_t)E * 3 * U) && read_block(f, 3, Um, (size_t)U * 3 * U) &&
              read_block(f, 3, b, (size_t)2 * 3 * U) && read_block(f, 3, D, (size_t)U * V) && read_block(f, 2, c, (size_t)V);
    fclose(f);
    if (!ok) return nullptr;
    gru_model * m = new gru_model();
    m->E = E; m->U = U; m->V = V;
    ggml_init_params p = {(size_t)(4 * (emb.size() + W.size() + Um.size() + b.size() + D.size() + c.size())) + (8u << 20), nullptr, false};
    m->ctx_w = ggml_init(p);
    m->embeddings = ggml_new_tensor_2d(m->ctx_w, GGML_TYPE_F32, E, V);  // file (V, E): ggml ne = (E, V), same bytes
    memcpy(m->embeddings->data, emb.data(), emb.size() * 4);
    m->kernel_t    = transposed_leaf(m->ctx_w, W, E, 3 * U);   // file (E, 3U) [in][out] -> ne (E, 3U): K = E fastest
    m->recurrent_t = transposed_leaf(m->ctx_w, Um, U, 3 * U);
    m->dense_t     = transposed_leaf(m->ctx_w, D, U, V);
    m->bias_x = ggml_new_tensor_2d(m->ctx_w, GGML_TYPE_F32, 3 * U, 1);
    m->bias_h = ggml_new_tensor_2d(m->ctx_w, GGML_TYPE_F32, 3 * U, 1);
    memcpy(m->bias_x->data, b.data(), (size_t)3 * U * 4);            // rnn.cpp:208 slice_2d(cell_bias, 0, 1)
    memcpy(m->bias_h->data, b.data() + 3 * U, (size_t)3 * U * 4);    // rnn.cpp:223 slice_2d(cell_bias, 1, 2)
    m->dense_bias = ggml_new_tensor_2d(m->ctx_w, GGML_TYPE_F32, V, 1);
    memcpy(m->dense_bias->data, c.data(), (size_t)V * 4);
    return m;
}

extern "C" void gru_free(gru_model * m) {
    if (!m) return;
    for (auto & kv : m->graphs) {
        ggml_graph_release_plan(kv.second.gf);
        ggml_free(kv.second.ctx);
        ggml_b200_host_free(kv.second.pinned);
    }
    ggml_free(m->ctx_w);
    delete m;
}
extern "C" int gru_vocab(const gru_model * m) { return m->V; }
extern "C" int gru_units(const gru_model * m) { return m->U; }

static ggml_tensor * sigmoid(ggml_context * ctx, ggml_tensor * x) { return ggml_div(ctx, ggml_silu(ctx, x), x); }  // rnn.cpp:51-55

// gru_forward (rnn.cpp:186-263) for B streams: every (n, 1) vector of the reference becomes an (n, B) matrix
static cell_graph & graph_for(gru_model * m, int B) {
    auto it = m->graphs.find(B);
    if (it != m->graphs.end()) return it->second;
    cell_graph g;
    const int U = m->U;
    const size_t arena = (size_t)B * (U + m->V + 8) * 4 * 2 + (8u << 20);
    g.pinned           = ggml_b200_host_malloc(arena);  // NULL without a device: ggml_init then mallocs
    ggml_init_params p = {arena, g.pinned, false};
    g.ctx = ggml_init(p);
    ggml_context * c = g.ctx;
    g.gf     = ggml_new_graph(c);
    g.ids    = ggml_new_tensor_1d(c, GGML_TYPE_I32, B);
    g.states = ggml_new_tensor_2d(c, GGML_TYPE_F32, U, B);
    ggml_set_name(g.ids, "input_id");
    ggml_set_name(g.states, "states");
    ggml_set_input(g.ids);
    ggml_set_input(g.states);
    ggml_tensor * x  = ggml_get_rows(c, m->embeddings, g.ids);                                   // (E, B)      :200
    ggml_tensor * mx = ggml_add(c, ggml_mul_mat(c, m->kernel_t, x), m->bias_x);                  // (3U, B)     :203-209
    ggml_tensor * mh = ggml_add(c, ggml_mul_mat(c, m->recurrent_t, g.states), m->bias_h);        // (3U, B)     :217-225
    auto gate = [&](ggml_tensor * t, int k) { return ggml_view_2d(c, t, U, B, t->nb[1], (size_t)k * U * sizeof(float)); };  // :213-215,227-229
    ggml_tensor * z  = sigmoid(c, ggml_add(c, gate(mx, 0), gate(mh, 0)));                        // :231
    ggml_tensor * r  = sigmoid(c, ggml_add(c, gate(mx, 1), gate(mh, 1)));                        // :232
    ggml_tensor * hh = ggml_tanh(c, ggml_add(c, gate(mx, 2), ggml_mul(c, r, gate(mh, 2))));      // :235-236
    ggml_tensor * one_minus_z = ggml_sub(c, ggml_repeat(c, ggml_new_f32(c, 1.0f), z), z);        // :243-246
    g.new_states = ggml_add(c, ggml_mul(c, z, g.states), ggml_mul(c, one_minus_z, hh));          // :239-250
    g.logits     = ggml_add(c, ggml_mul_mat(c, m->dense_t, g.new_states), m->dense_bias);        // (V, B)      :252-258
    g.next_ids   = ggml_argmax(c, g.logits);                                                     // argmax_1d   :74-77,312
    ggml_set_name(g.new_states, "new_states");
    ggml_set_name(g.next_ids, "next_ids");
    ggml_build_forward_expand(g.gf, g.next_ids);  // the only per-step output (B token ids); the state stays on the device
    // the loop of rnn.cpp:293-313 stays on the device: next token and state feed the next step directly
    ggml_b200_graph_add_feedback(g.gf, g.next_ids, g.ids);
    ggml_b200_graph_add_feedback(g.gf, g.new_states, g.states);
    return m->graphs.emplace(B, g).first->second;
}

extern "C" float gru_generate(gru_model * m, const int32_t * first_tokens, int B, int steps, int32_t * out_tokens, float * final_state) {
    if (!m || !first_tokens || B <= 0 || steps <= 0 || !out_tokens) return -1.f;
    cell_graph & g = graph_for(m, B);
    memcpy(g.ids->data, first_tokens, (size_t)B * 4);
    memset(g.states->data, 0, (size_t)B * m->U * 4);
    // step 0 uploads ids/state from the host; later steps take them from the feedback copies.  The whole loop is enqueued
    // at once (no host round trip per token); the chosen ids of every step come back in one copy at the end.
    ggml_b200_graph_prepare(g.ctx, g.gf);
    ggml_b200_graph_set_transfers(g.gf, true, false);
    const int64_t t0 = ggml_time_us();
    if (getenv("GRU_B200_HOST_LOOP")) {  // the old per-step loop (one D2H + sync per token), kept for comparison
        ggml_b200_graph_set_transfers(g.gf, true, true);
        for (int t = 0; t < steps; t++) {
            ggml_graph_compute_with_ctx(g.ctx, g.gf, 1);
            memcpy(out_tokens + (size_t)t * B, g.next_ids->data, (size_t)B * 4);
            ggml_b200_graph_set_transfers(g.gf, false, true);
        }
    } else if (ggml_b200_graph_compute_steps(g.ctx, g.gf, steps, g.next_ids, out_tokens) != 0) {
        return -1.f;
    }
    const int64_t t1 = ggml_time_us();
    if (getenv("GRU_B200_PROFILE")) {  // per-launch device times of one cell step, to stderr
        std::vector<char> buf(1 << 16);
        if (ggml_b200_graph_profile_json(g.gf, 20, buf.data(), buf.size()) == 0) fprintf(stderr, "%s\n", buf.data());
    }
    if (final_state && ggml_b200_tensor_download(g.gf, g.new_states, final_state) != 0) return -1.f;
    return (float)(t1 - t0) / 1000.f;
}
