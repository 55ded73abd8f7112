/*
 * ggml_cpu_ref.c -- CPU ORACLE, part 2.  TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A small, self-contained CPU implementation of the ggml C API subset declared in include/ggml/ggml.h, with the
 * numerics of upstream ggml's CPU kernels restated per op ([ggml] = github.com/ggerganov/ggml, ~Feb-May 2024 by API
 * dating; upstream is not vendored by the reference and not fetchable here).  Its purpose: the
 * reference's OWN programs -- /root/reference/mobilevit/main.cpp and /root/reference/rnn_text_gen/
 * rnn_text_generation.cpp, compiled UNMODIFIED from where they lie (oracle/Makefile -> oracle/_ref/) -- run on the
 * CPU here, so that
 *   (1) the monolithic oracle (mobilevit_oracle.c) is pinned against the reference's own graph builder, and
 *   (2) the GPU library can be compared with the same program node by node (GGML_CPU_REF_DUMP / GGML_B200_DUMP_NODES).
 *
 * Semantics restated from upstream, with the reference call sites they serve:
 *   tensors are allocated eagerly in the context arena (no_alloc=false), views share their source's bytes;
 *   ggml_build_forward_expand = post-order DFS over src[0..] ([ggml] ggml_visit_parents);
 *   add/sub/mul/div broadcast src1 over src0 (main.cpp:810-846);  norm / soft_max accumulate in double ([ggml] ggml_float);
 *   mul_mat(a, b): a F16 -> b rounded to F16, f32 accumulation; a F32 -> plain f32 (main.cpp:1022-1151);
 *   conv_2d / conv_depthwise_2d = im2col (activations -> F16, zero padding, column order (ic, kh, kw)) + f16 dot (main.cpp:788,798).
 * GGML_CPU_REF_LEGACY=1 selects the legacy f16 lookup tables for silu and the softmax exponential (the ggml the author
 * ran: mobilevit/README.md:39-45 prints f16-representable values).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <immintrin.h>

#include "ggml/ggml.h"

struct ggml_context {
    size_t mem_size;
    char * mem;
    int    owned;
    size_t used;
};

static int g_legacy = -1;
static int legacy_mode(void) {
    if (g_legacy < 0) {
        const char * e = getenv("GGML_CPU_REF_LEGACY");
        g_legacy = e && atoi(e) > 0;
    }
    return g_legacy;
}

/* ---- fp16: [ggml] GGML_FP32_TO_FP16 on x86 with F16C = _cvtss_sh(x, 0), round to nearest even ---- */
float       ggml_fp16_to_fp32(ggml_fp16_t x) { return _cvtsh_ss(x); }
ggml_fp16_t ggml_fp32_to_fp16(float x) { return _cvtss_sh(x, 0); }
void ggml_fp16_to_fp32_row(const ggml_fp16_t * x, float * y, int n) { for (int i = 0; i < n; i++) y[i] = ggml_fp16_to_fp32(x[i]); }
void ggml_fp32_to_fp16_row(const float * x, ggml_fp16_t * y, int n) { for (int i = 0; i < n; i++) y[i] = ggml_fp32_to_fp16(x[i]); }

/* ---- timers (main.cpp:639-641,651) ---- */
void    ggml_time_init(void) {}
int64_t ggml_time_us(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (int64_t)ts.tv_sec * 1000000 + ts.tv_nsec / 1000; }
int64_t ggml_time_ms(void) { return ggml_time_us() / 1000; }

/* ---- context ---- */
struct ggml_context * ggml_init(struct ggml_init_params params) {
    struct ggml_context * ctx = (struct ggml_context *)calloc(1, sizeof(*ctx));
    ctx->mem_size = params.mem_size;
    ctx->mem      = params.mem_buffer ? (char *)params.mem_buffer : (char *)malloc(params.mem_size);
    ctx->owned    = params.mem_buffer == NULL;
    if (!ctx->mem) { free(ctx); return NULL; }
    return ctx;
}
void ggml_free(struct ggml_context * ctx) {
    if (!ctx) return;
    if (ctx->owned) free(ctx->mem);
    free(ctx);
}
size_t ggml_used_mem(const struct ggml_context * ctx) { return ctx->used; }
static void * arena(struct ggml_context * ctx, size_t bytes, size_t align) {
    size_t off = (ctx->used + align - 1) / align * align;
    if (off + bytes > ctx->mem_size) {
        fprintf(stderr, "ggml_cpu_ref: not enough space in the context's memory pool (needed %zu, available %zu)\n", off + bytes, ctx->mem_size);
        abort();
    }
    ctx->used = off + bytes;
    return ctx->mem + off;
}

size_t ggml_type_size(enum ggml_type t) { return t == GGML_TYPE_F16 ? 2 : 4; }
int64_t ggml_nelements(const struct ggml_tensor * t) { return t->ne[0] * t->ne[1] * t->ne[2] * t->ne[3]; }
size_t ggml_nbytes(const struct ggml_tensor * t) { return (size_t)ggml_nelements(t) * ggml_type_size(t->type); }
int ggml_n_dims(const struct ggml_tensor * t) {
    for (int i = GGML_MAX_DIMS - 1; i >= 1; i--)
        if (t->ne[i] > 1) return i + 1;
    return 1;
}
bool ggml_is_contiguous(const struct ggml_tensor * t) {
    return t->nb[0] == ggml_type_size(t->type) && t->nb[1] == t->nb[0] * (size_t)t->ne[0] && t->nb[2] == t->nb[1] * (size_t)t->ne[1] &&
           t->nb[3] == t->nb[2] * (size_t)t->ne[2];
}

static struct ggml_tensor * new_tensor(struct ggml_context * ctx, enum ggml_type type, int n_dims, const int64_t * ne, struct ggml_tensor * view_src,
                                       size_t view_offs) {
    struct ggml_tensor * t = (struct ggml_tensor *)arena(ctx, sizeof(*t), 16);
    memset(t, 0, sizeof(*t));
    t->type = type;
    t->n_dims = n_dims;
    for (int i = 0; i < 4; i++) t->ne[i] = i < n_dims ? ne[i] : 1;
    t->nb[0] = ggml_type_size(type);
    for (int i = 1; i < 4; i++) t->nb[i] = t->nb[i - 1] * (size_t)t->ne[i - 1];
    t->ctx = ctx;
    if (view_src) {
        t->view_src  = view_src->view_src ? view_src->view_src : view_src;
        t->view_offs = view_offs + (view_src->view_src ? view_src->view_offs : 0);
        t->data      = (char *)view_src->data + view_offs;
    } else {
        t->data = arena(ctx, ggml_nbytes(t) ? ggml_nbytes(t) : 4, 64);  /* [ggml] no_alloc = false: data lives in the arena */
    }
    return t;
}
struct ggml_tensor * ggml_new_tensor_1d(struct ggml_context * c, enum ggml_type t, int64_t a) { int64_t ne[4] = {a, 1, 1, 1}; return new_tensor(c, t, 1, ne, NULL, 0); }
struct ggml_tensor * ggml_new_tensor_2d(struct ggml_context * c, enum ggml_type t, int64_t a, int64_t b) { int64_t ne[4] = {a, b, 1, 1}; return new_tensor(c, t, 2, ne, NULL, 0); }
struct ggml_tensor * ggml_new_tensor_3d(struct ggml_context * c, enum ggml_type t, int64_t a, int64_t b, int64_t d) { int64_t ne[4] = {a, b, d, 1}; return new_tensor(c, t, 3, ne, NULL, 0); }
struct ggml_tensor * ggml_new_tensor_4d(struct ggml_context * c, enum ggml_type t, int64_t a, int64_t b, int64_t d, int64_t e) { int64_t ne[4] = {a, b, d, e}; return new_tensor(c, t, 4, ne, NULL, 0); }
struct ggml_tensor * ggml_new_f32(struct ggml_context * c, float v) {
    struct ggml_tensor * t = ggml_new_tensor_1d(c, GGML_TYPE_F32, 1);
    *(float *)t->data = v;
    return t;
}
struct ggml_tensor * ggml_set_name(struct ggml_tensor * t, const char * name) { snprintf(t->name, sizeof t->name, "%s", name); return t; }
void ggml_set_input(struct ggml_tensor * t) { t->flags |= GGML_TENSOR_FLAG_INPUT; }
void ggml_set_output(struct ggml_tensor * t) { t->flags |= GGML_TENSOR_FLAG_OUTPUT; }
void ggml_set_param(struct ggml_context * c, struct ggml_tensor * t) { (void)c; t->flags |= GGML_TENSOR_FLAG_PARAM; }

void *  ggml_get_data(const struct ggml_tensor * t) { return t->data; }
float * ggml_get_data_f32(const struct ggml_tensor * t) { GGML_ASSERT(t->type == GGML_TYPE_F32); return (float *)t->data; }
void    ggml_set_i32_1d(const struct ggml_tensor * t, int i, int32_t v) { ((int32_t *)t->data)[i] = v; }
int32_t ggml_get_i32_1d(const struct ggml_tensor * t, int i) { return ((int32_t *)t->data)[i]; }
void    ggml_set_f32_1d(const struct ggml_tensor * t, int i, float v) { ((float *)t->data)[i] = v; }
float   ggml_get_f32_1d(const struct ggml_tensor * t, int i) {
    if (t->type == GGML_TYPE_F16) return ggml_fp16_to_fp32(((ggml_fp16_t *)t->data)[i]);
    if (t->type == GGML_TYPE_I32) return (float)((int32_t *)t->data)[i];
    return ((float *)t->data)[i];
}

/* ---- op constructors ---- */
static struct ggml_tensor * op_result(struct ggml_context * ctx, enum ggml_op op, enum ggml_type type, const int64_t * ne, struct ggml_tensor * a,
                                      struct ggml_tensor * b) {
    struct ggml_tensor * t = new_tensor(ctx, type, 4, ne, NULL, 0);
    t->n_dims = ggml_n_dims(t);
    t->op = op; t->src[0] = a; t->src[1] = b;
    return t;
}
static int can_repeat(const struct ggml_tensor * b, const struct ggml_tensor * a) {  /* [ggml] ggml_can_repeat(b, a) */
    for (int i = 0; i < 4; i++) if (a->ne[i] % b->ne[i]) return 0;
    return 1;
}
static struct ggml_tensor * binary(struct ggml_context * ctx, enum ggml_op op, struct ggml_tensor * a, struct ggml_tensor * b) {
    GGML_ASSERT(can_repeat(b, a));
    return op_result(ctx, op, GGML_TYPE_F32, a->ne, a, b);
}
struct ggml_tensor * ggml_add(struct ggml_context * c, struct ggml_tensor * a, struct ggml_tensor * b) { return binary(c, GGML_OP_ADD, a, b); }
struct ggml_tensor * ggml_sub(struct ggml_context * c, struct ggml_tensor * a, struct ggml_tensor * b) { return binary(c, GGML_OP_SUB, a, b); }
struct ggml_tensor * ggml_mul(struct ggml_context * c, struct ggml_tensor * a, struct ggml_tensor * b) { return binary(c, GGML_OP_MUL, a, b); }
struct ggml_tensor * ggml_div(struct ggml_context * c, struct ggml_tensor * a, struct ggml_tensor * b) { return binary(c, GGML_OP_DIV, a, b); }
static struct ggml_tensor * unary(struct ggml_context * c, enum ggml_op op, struct ggml_tensor * a) { return op_result(c, op, GGML_TYPE_F32, a->ne, a, NULL); }
struct ggml_tensor * ggml_sqrt(struct ggml_context * c, struct ggml_tensor * a) { return unary(c, GGML_OP_SQRT, a); }
struct ggml_tensor * ggml_silu(struct ggml_context * c, struct ggml_tensor * a) { return unary(c, GGML_OP_SILU, a); }
struct ggml_tensor * ggml_tanh(struct ggml_context * c, struct ggml_tensor * a) { return unary(c, GGML_OP_TANH, a); }
struct ggml_tensor * ggml_soft_max(struct ggml_context * c, struct ggml_tensor * a) { return unary(c, GGML_OP_SOFT_MAX, a); }
struct ggml_tensor * ggml_norm(struct ggml_context * c, struct ggml_tensor * a, float eps) {
    struct ggml_tensor * t = unary(c, GGML_OP_NORM, a);
    memcpy(t->op_params, &eps, sizeof eps);
    return t;
}
struct ggml_tensor * ggml_mul_mat(struct ggml_context * c, struct ggml_tensor * a, struct ggml_tensor * b) {
    GGML_ASSERT(a->ne[0] == b->ne[0] && b->ne[2] % a->ne[2] == 0 && b->ne[3] % a->ne[3] == 0);  /* [ggml] ggml_can_mul_mat */
    const int64_t ne[4] = {a->ne[1], b->ne[1], b->ne[2], b->ne[3]};
    return op_result(c, GGML_OP_MUL_MAT, GGML_TYPE_F32, ne, a, b);
}
struct ggml_tensor * ggml_repeat(struct ggml_context * c, struct ggml_tensor * a, struct ggml_tensor * b) {
    GGML_ASSERT(can_repeat(a, b));
    return op_result(c, GGML_OP_REPEAT, a->type, b->ne, a, NULL);
}
struct ggml_tensor * ggml_concat(struct ggml_context * c, struct ggml_tensor * a, struct ggml_tensor * b) {  /* [ggml] 2-arg form: dim 2 */
    GGML_ASSERT(a->ne[0] == b->ne[0] && a->ne[1] == b->ne[1] && a->ne[3] == b->ne[3]);
    const int64_t ne[4] = {a->ne[0], a->ne[1], a->ne[2] + b->ne[2], a->ne[3]};
    return op_result(c, GGML_OP_CONCAT, GGML_TYPE_F32, ne, a, b);
}
struct ggml_tensor * ggml_get_rows(struct ggml_context * c, struct ggml_tensor * a, struct ggml_tensor * b) {
    GGML_ASSERT(b->type == GGML_TYPE_I32);
    const int64_t ne[4] = {a->ne[0], b->ne[0], b->ne[1], b->ne[2]};
    return op_result(c, GGML_OP_GET_ROWS, GGML_TYPE_F32, ne, a, b);
}
struct ggml_tensor * ggml_cont(struct ggml_context * c, struct ggml_tensor * a) { return op_result(c, GGML_OP_CONT, a->type, a->ne, a, NULL); }
struct ggml_tensor * ggml_cont_4d(struct ggml_context * c, struct ggml_tensor * a, int64_t n0, int64_t n1, int64_t n2, int64_t n3) {
    GGML_ASSERT(ggml_nelements(a) == n0 * n1 * n2 * n3);
    const int64_t ne[4] = {n0, n1, n2, n3};
    return op_result(c, GGML_OP_CONT, a->type, ne, a, NULL);
}
static struct ggml_tensor * view_of(struct ggml_context * c, enum ggml_op op, struct ggml_tensor * a, const int64_t * ne, const size_t * nb, size_t offs) {
    struct ggml_tensor * t = new_tensor(c, a->type, 4, ne, a, offs);
    if (nb) for (int i = 0; i < 4; i++) t->nb[i] = nb[i];
    t->n_dims = ggml_n_dims(t);
    t->op = op; t->src[0] = a;
    return t;
}
struct ggml_tensor * ggml_reshape_2d(struct ggml_context * c, struct ggml_tensor * a, int64_t n0, int64_t n1) { return ggml_reshape_4d(c, a, n0, n1, 1, 1); }
struct ggml_tensor * ggml_reshape_3d(struct ggml_context * c, struct ggml_tensor * a, int64_t n0, int64_t n1, int64_t n2) { return ggml_reshape_4d(c, a, n0, n1, n2, 1); }
struct ggml_tensor * ggml_reshape_4d(struct ggml_context * c, struct ggml_tensor * a, int64_t n0, int64_t n1, int64_t n2, int64_t n3) {
    GGML_ASSERT(ggml_is_contiguous(a) && ggml_nelements(a) == n0 * n1 * n2 * n3);
    const int64_t ne[4] = {n0, n1, n2, n3};
    return view_of(c, GGML_OP_RESHAPE, a, ne, NULL, 0);
}
struct ggml_tensor * ggml_permute(struct ggml_context * c, struct ggml_tensor * a, int ax0, int ax1, int ax2, int ax3) {
    const int ax[4] = {ax0, ax1, ax2, ax3};
    int64_t ne[4]; size_t nb[4];
    for (int i = 0; i < 4; i++) { ne[ax[i]] = a->ne[i]; nb[ax[i]] = a->nb[i]; }  /* [ggml] result.ne[axis_i] = a.ne[i] */
    struct ggml_tensor * t = view_of(c, GGML_OP_PERMUTE, a, ne, nb, 0);
    for (int i = 0; i < 4; i++) t->op_params[i] = ax[i];
    return t;
}
struct ggml_tensor * ggml_transpose(struct ggml_context * c, struct ggml_tensor * a) {
    const int64_t ne[4] = {a->ne[1], a->ne[0], a->ne[2], a->ne[3]};
    const size_t  nb[4] = {a->nb[1], a->nb[0], a->nb[2], a->nb[3]};
    return view_of(c, GGML_OP_TRANSPOSE, a, ne, nb, 0);
}
struct ggml_tensor * ggml_view_2d(struct ggml_context * c, struct ggml_tensor * a, int64_t n0, int64_t n1, size_t nb1, size_t offset) {
    const int64_t ne[4] = {n0, n1, 1, 1};
    const size_t  nb[4] = {a->nb[0], nb1, nb1 * (size_t)n1, nb1 * (size_t)n1};
    return view_of(c, GGML_OP_VIEW, a, ne, nb, offset);
}
struct ggml_tensor * ggml_argmax(struct ggml_context * c, struct ggml_tensor * a) {
    const int64_t ne[4] = {a->ne[1], 1, 1, 1};
    return op_result(c, GGML_OP_ARGMAX, GGML_TYPE_I32, ne, a, NULL);
}
static struct ggml_tensor * conv_result(struct ggml_context * c, enum ggml_op op, struct ggml_tensor * k, struct ggml_tensor * x, int s0, int s1, int p0, int p1,
                                        int d0, int d1, int64_t oc) {
    const int64_t ow = (x->ne[0] + 2 * p0 - d0 * (k->ne[0] - 1) - 1) / s0 + 1;  /* [ggml] ggml_calc_conv_output_size */
    const int64_t oh = (x->ne[1] + 2 * p1 - d1 * (k->ne[1] - 1) - 1) / s1 + 1;
    const int64_t ne[4] = {ow, oh, oc, x->ne[3]};
    struct ggml_tensor * t = op_result(c, op, GGML_TYPE_F32, ne, k, x);
    const int32_t prm[6] = {s0, s1, p0, p1, d0, d1};
    memcpy(t->op_params, prm, sizeof prm);
    return t;
}
struct ggml_tensor * ggml_conv_2d(struct ggml_context * c, struct ggml_tensor * k, struct ggml_tensor * x, int s0, int s1, int p0, int p1, int d0, int d1) {
    GGML_ASSERT(k->ne[2] == x->ne[2] && k->type == GGML_TYPE_F16);
    return conv_result(c, GGML_OP_CONV_2D, k, x, s0, s1, p0, p1, d0, d1, k->ne[3]);
}
struct ggml_tensor * ggml_conv_depthwise_2d(struct ggml_context * c, struct ggml_tensor * k, struct ggml_tensor * x, int s0, int s1, int p0, int p1, int d0,
                                            int d1) {
    GGML_ASSERT(k->ne[2] == 1 && k->ne[3] == x->ne[2] && k->type == GGML_TYPE_F16);
    return conv_result(c, GGML_OP_CONV_DEPTHWISE_2D, k, x, s0, s1, p0, p1, d0, d1, x->ne[2]);
}
struct ggml_tensor * ggml_b200_pool_mean_hw(struct ggml_context * c, struct ggml_tensor * a) {
    const int64_t ne[4] = {1, 1, a->ne[2], a->ne[3]};
    return op_result(c, GGML_OP_POOL_MEAN_HW, GGML_TYPE_F32, ne, a, NULL);
}

/* ---- graph: [ggml] ggml_visit_parents = post-order DFS over src[0..], leafs and nodes kept apart ---- */
static void fix_graph(struct ggml_cgraph * g) {
    if (g->size <= GGML_B200_STATIC_GRAPH_NODES) { g->size = GGML_B200_STATIC_GRAPH_NODES; g->nodes = g->static_nodes; g->leafs = g->static_leafs; }
}
struct ggml_cgraph * ggml_new_graph(struct ggml_context * ctx) {
    struct ggml_cgraph * g = (struct ggml_cgraph *)arena(ctx, sizeof(*g), 16);
    memset(g, 0, sizeof(*g));
    g->size  = GGML_DEFAULT_GRAPH_SIZE * 4;
    g->nodes = (struct ggml_tensor **)arena(ctx, sizeof(void *) * (size_t)g->size, 16);
    g->leafs = (struct ggml_tensor **)arena(ctx, sizeof(void *) * (size_t)g->size, 16);
    return g;
}
static int in_graph(const struct ggml_cgraph * g, const struct ggml_tensor * t) {
    for (int i = 0; i < g->n_nodes; i++) if (g->nodes[i] == t) return 1;
    for (int i = 0; i < g->n_leafs; i++) if (g->leafs[i] == t) return 1;
    return 0;
}
static void visit(struct ggml_cgraph * g, struct ggml_tensor * t) {
    if (in_graph(g, t)) return;
    for (int s = 0; s < GGML_MAX_SRC; s++) if (t->src[s]) visit(g, t->src[s]);
    if (t->op == GGML_OP_NONE) { GGML_ASSERT(g->n_leafs < g->size); g->leafs[g->n_leafs++] = t; }
    else { GGML_ASSERT(g->n_nodes < g->size); g->nodes[g->n_nodes++] = t; }
}
void ggml_build_forward_expand(struct ggml_cgraph * g, struct ggml_tensor * t) { fix_graph(g); visit(g, t); }
struct ggml_cgraph ggml_build_forward(struct ggml_tensor * t) {
    struct ggml_cgraph g;
    memset(&g, 0, sizeof g);
    ggml_build_forward_expand(&g, t);
    return g;
}
void ggml_graph_release_plan(struct ggml_cgraph * g) { (void)g; }

/* ---- kernels ---- */
#define AT(t, i0, i1, i2, i3) ((char *)(t)->data + (size_t)(i0) * (t)->nb[0] + (size_t)(i1) * (t)->nb[1] + (size_t)(i2) * (t)->nb[2] + (size_t)(i3) * (t)->nb[3])
static float ld(const struct ggml_tensor * t, int64_t i0, int64_t i1, int64_t i2, int64_t i3) {
    const char * p = AT(t, i0, i1, i2, i3);
    return t->type == GGML_TYPE_F16 ? ggml_fp16_to_fp32(*(const ggml_fp16_t *)p) : *(const float *)p;
}
static float silu_exact(float x) { return x / (1.0f + expf(-x)); }  /* [ggml] ggml_silu_f32 */
static ggml_fp16_t * g_tab_silu = NULL, * g_tab_exp = NULL;       /* [ggml, legacy] ggml_table_silu_f16 / ggml_table_exp_f16 */
static void build_tables(void) {
    if (g_tab_silu) return;
    g_tab_silu = (ggml_fp16_t *)malloc(65536 * 2);
    g_tab_exp  = (ggml_fp16_t *)malloc(65536 * 2);
    for (int i = 0; i < 65536; i++) {
        const float f = ggml_fp16_to_fp32((ggml_fp16_t)i);
        g_tab_silu[i] = ggml_fp32_to_fp16(silu_exact(f));
        g_tab_exp[i]  = ggml_fp32_to_fp16(expf(f));
    }
}

static void k_binary(struct ggml_tensor * d) {
    const struct ggml_tensor * a = d->src[0], * b = d->src[1];
    for (int64_t i3 = 0; i3 < d->ne[3]; i3++) for (int64_t i2 = 0; i2 < d->ne[2]; i2++) for (int64_t i1 = 0; i1 < d->ne[1]; i1++) {
        float * o = (float *)AT(d, 0, i1, i2, i3);
        for (int64_t i0 = 0; i0 < d->ne[0]; i0++) {
            const float x = ld(a, i0, i1, i2, i3), y = ld(b, i0 % b->ne[0], i1 % b->ne[1], i2 % b->ne[2], i3 % b->ne[3]);
            o[i0] = d->op == GGML_OP_ADD ? x + y : d->op == GGML_OP_SUB ? x - y : d->op == GGML_OP_MUL ? x * y : x / y;
        }
    }
}
static void k_unary(struct ggml_tensor * d) {
    const struct ggml_tensor * a = d->src[0];
    const int legacy = legacy_mode();
    if (legacy) build_tables();
    for (int64_t i3 = 0; i3 < d->ne[3]; i3++) for (int64_t i2 = 0; i2 < d->ne[2]; i2++) for (int64_t i1 = 0; i1 < d->ne[1]; i1++) {
        float * o = (float *)AT(d, 0, i1, i2, i3);
        for (int64_t i0 = 0; i0 < d->ne[0]; i0++) {
            const float x = ld(a, i0, i1, i2, i3);
            if (d->op == GGML_OP_SQRT) o[i0] = sqrtf(x);
            else if (d->op == GGML_OP_TANH) o[i0] = tanhf(x);
            else o[i0] = legacy ? ggml_fp16_to_fp32(g_tab_silu[ggml_fp32_to_fp16(x)]) : silu_exact(x);
        }
    }
}
static void k_norm(struct ggml_tensor * d) {  /* [ggml] ggml_compute_forward_norm_f32 */
    const struct ggml_tensor * a = d->src[0];
    float eps; memcpy(&eps, d->op_params, sizeof eps);
    const int64_t n = d->ne[0];
    for (int64_t i3 = 0; i3 < d->ne[3]; i3++) for (int64_t i2 = 0; i2 < d->ne[2]; i2++) for (int64_t i1 = 0; i1 < d->ne[1]; i1++) {
        float * o = (float *)AT(d, 0, i1, i2, i3);
        double sum = 0.0;
        for (int64_t i = 0; i < n; i++) sum += (double)ld(a, i, i1, i2, i3);
        const float mean = (float)(sum / (double)n);
        double sum2 = 0.0;
        for (int64_t i = 0; i < n; i++) { const float v = ld(a, i, i1, i2, i3) - mean; o[i] = v; sum2 += (double)(v * v); }
        const float var = (float)(sum2 / (double)n);
        const float scale = 1.0f / sqrtf(var + eps);
        for (int64_t i = 0; i < n; i++) o[i] *= scale;
    }
}
static void k_soft_max(struct ggml_tensor * d) {  /* [ggml] ggml_compute_forward_soft_max_f32 (no mask, scale 1) */
    const struct ggml_tensor * a = d->src[0];
    const int legacy = legacy_mode();
    if (legacy) build_tables();
    const int64_t n = d->ne[0];
    for (int64_t i3 = 0; i3 < d->ne[3]; i3++) for (int64_t i2 = 0; i2 < d->ne[2]; i2++) for (int64_t i1 = 0; i1 < d->ne[1]; i1++) {
        float * o = (float *)AT(d, 0, i1, i2, i3);
        float mx = -INFINITY;
        for (int64_t i = 0; i < n; i++) mx = fmaxf(mx, ld(a, i, i1, i2, i3));
        double sum = 0.0;
        for (int64_t i = 0; i < n; i++) {
            const float x = ld(a, i, i1, i2, i3) - mx;
            const float v = legacy ? ggml_fp16_to_fp32(g_tab_exp[ggml_fp32_to_fp16(x)]) : expf(x);
            o[i] = v;
            sum += (double)v;
        }
        const float inv = (float)(1.0 / sum);
        for (int64_t i = 0; i < n; i++) o[i] *= inv;
    }
}
static void k_mul_mat(struct ggml_tensor * d) {  /* [ggml] ggml_compute_forward_mul_mat: dst[m, n] = sum_k a[k, m] * b[k, n] */
    const struct ggml_tensor * a = d->src[0], * b = d->src[1];
    const int64_t K = a->ne[0], M = a->ne[1], N = b->ne[1];
    const int64_t r2 = b->ne[2] / a->ne[2], r3 = b->ne[3] / a->ne[3];
    const int f16 = a->type == GGML_TYPE_F16;  /* vec_dot_type F16: src1 is converted (rounded) to F16 first */
    float * brow = (float *)malloc(sizeof(float) * (size_t)K), * arow = (float *)malloc(sizeof(float) * (size_t)K * (size_t)M);
    for (int64_t i3 = 0; i3 < b->ne[3]; i3++) for (int64_t i2 = 0; i2 < b->ne[2]; i2++) {
        for (int64_t m = 0; m < M; m++) for (int64_t k = 0; k < K; k++) arow[m * K + k] = ld(a, k, m, i2 / r2, i3 / r3);
        for (int64_t n = 0; n < N; n++) {
            for (int64_t k = 0; k < K; k++) {
                float v = ld(b, k, n, i2, i3);
                brow[k] = f16 ? ggml_fp16_to_fp32(ggml_fp32_to_fp16(v)) : v;
            }
            for (int64_t m = 0; m < M; m++) {
                const float * ar = arow + m * K;
                float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};  /* f32 accumulation in SIMD-like lanes, as ggml_vec_dot_f32 / _f16 do */
                int64_t k = 0;
                for (; k + 8 <= K; k += 8) for (int j = 0; j < 8; j++) acc[j] += ar[k + j] * brow[k + j];
                float s = ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
                for (; k < K; k++) s += ar[k] * brow[k];
                *(float *)AT(d, m, n, i2, i3) = s;
            }
        }
    }
    free(brow); free(arow);
}
static void k_repeat(struct ggml_tensor * d) {
    const struct ggml_tensor * a = d->src[0];
    for (int64_t i3 = 0; i3 < d->ne[3]; i3++) for (int64_t i2 = 0; i2 < d->ne[2]; i2++) for (int64_t i1 = 0; i1 < d->ne[1]; i1++) for (int64_t i0 = 0; i0 < d->ne[0]; i0++)
        memcpy(AT(d, i0, i1, i2, i3), AT(a, i0 % a->ne[0], i1 % a->ne[1], i2 % a->ne[2], i3 % a->ne[3]), ggml_type_size(d->type));
}
static void k_concat(struct ggml_tensor * d) {
    const struct ggml_tensor * a = d->src[0], * b = d->src[1];
    for (int64_t i3 = 0; i3 < d->ne[3]; i3++) for (int64_t i2 = 0; i2 < d->ne[2]; i2++) for (int64_t i1 = 0; i1 < d->ne[1]; i1++) for (int64_t i0 = 0; i0 < d->ne[0]; i0++)
        *(float *)AT(d, i0, i1, i2, i3) = i2 < a->ne[2] ? ld(a, i0, i1, i2, i3) : ld(b, i0, i1, i2 - a->ne[2], i3);
}
static void k_get_rows(struct ggml_tensor * d) {  /* [ggml] ggml_compute_forward_get_rows_f32: rows are copied as nc contiguous floats */
    const struct ggml_tensor * a = d->src[0], * idx = d->src[1];
    const int64_t nc = a->ne[0];
    for (int64_t i12 = 0; i12 < idx->ne[2]; i12++) for (int64_t i11 = 0; i11 < idx->ne[1]; i11++) for (int64_t i10 = 0; i10 < idx->ne[0]; i10++) {
        const int32_t r = *(const int32_t *)AT(idx, i10, i11, i12, 0);
        const char * src = (const char *)a->data + (size_t)r * a->nb[1] + (size_t)i11 * a->nb[2] + (size_t)i12 * a->nb[3];
        float * dst = (float *)AT(d, 0, i10, i11, i12);
        if (a->type == GGML_TYPE_F16) for (int64_t i = 0; i < nc; i++) dst[i] = ggml_fp16_to_fp32(((const ggml_fp16_t *)src)[i]);
        else memcpy(dst, src, (size_t)nc * 4);
    }
}
static void k_cont(struct ggml_tensor * d) {  /* element order of the SOURCE view, written contiguously (cont_4d only relabels the shape) */
    const struct ggml_tensor * a = d->src[0];
    const size_t es = ggml_type_size(d->type);
    char * o = (char *)d->data;
    for (int64_t i3 = 0; i3 < a->ne[3]; i3++) for (int64_t i2 = 0; i2 < a->ne[2]; i2++) for (int64_t i1 = 0; i1 < a->ne[1]; i1++) for (int64_t i0 = 0; i0 < a->ne[0]; i0++) {
        memcpy(o, AT(a, i0, i1, i2, i3), es);
        o += es;
    }
}
static void k_argmax(struct ggml_tensor * d) {
    const struct ggml_tensor * a = d->src[0];
    for (int64_t r = 0; r < a->ne[1]; r++) {
        int best = 0; float bv = -INFINITY;
        for (int64_t i = 0; i < a->ne[0]; i++) { const float v = ld(a, i, r, 0, 0); if (v > bv) { bv = v; best = (int)i; } }
        ((int32_t *)d->data)[r] = best;
    }
}
/* [ggml] ggml_conv_2d = im2col(F16) + mul_mat(f16 x f16 -> f32).  Column index k = ic*KH*KW + kh*KW + kw; out-of-image taps are 0.
 * The batch dimension is handled per image (upstream's result reshape of that era is only right for N = 1; the reference is N = 1). */
static void k_conv(struct ggml_tensor * d, int depthwise) {
    const struct ggml_tensor * w = d->src[0], * x = d->src[1];
    int32_t prm[6]; memcpy(prm, d->op_params, sizeof prm);
    const int s0 = prm[0], s1 = prm[1], p0 = prm[2], p1 = prm[3], d0 = prm[4], d1 = prm[5];
    const int64_t KW = w->ne[0], KH = w->ne[1], IC = depthwise ? 1 : w->ne[2], OC = d->ne[2];
    const int64_t W = x->ne[0], H = x->ne[1], OW = d->ne[0], OH = d->ne[1], N = d->ne[3];
    const int64_t K = IC * KH * KW;
    float * col = (float *)malloc(sizeof(float) * (size_t)K), * wf = (float *)malloc(sizeof(float) * (size_t)K * (size_t)OC);
    for (int64_t oc = 0; oc < OC; oc++)
        for (int64_t ic = 0; ic < IC; ic++) for (int64_t kh = 0; kh < KH; kh++) for (int64_t kw = 0; kw < KW; kw++)
            wf[oc * K + (ic * KH + kh) * KW + kw] = depthwise ? ld(w, kw, kh, 0, oc) : ld(w, kw, kh, ic, oc);
    for (int64_t n = 0; n < N; n++) for (int64_t oh = 0; oh < OH; oh++) for (int64_t ow = 0; ow < OW; ow++) {
        if (!depthwise) {
            for (int64_t ic = 0; ic < IC; ic++) for (int64_t kh = 0; kh < KH; kh++) for (int64_t kw = 0; kw < KW; kw++) {
                const int64_t ih = oh * s1 + kh * d1 - p1, iw = ow * s0 + kw * d0 - p0;
                const float v = (ih < 0 || ih >= H || iw < 0 || iw >= W) ? 0.f : ld(x, iw, ih, ic, n);
                col[(ic * KH + kh) * KW + kw] = ggml_fp16_to_fp32(ggml_fp32_to_fp16(v));  /* im2col writes F16 */
            }
        }
        for (int64_t oc = 0; oc < OC; oc++) {
            if (depthwise) {
                for (int64_t kh = 0; kh < KH; kh++) for (int64_t kw = 0; kw < KW; kw++) {
                    const int64_t ih = oh * s1 + kh * d1 - p1, iw = ow * s0 + kw * d0 - p0;
                    const float v = (ih < 0 || ih >= H || iw < 0 || iw >= W) ? 0.f : ld(x, iw, ih, oc, n);
                    col[kh * KW + kw] = ggml_fp16_to_fp32(ggml_fp32_to_fp16(v));
                }
            }
            const float * wr = wf + oc * K;
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            int64_t k = 0;
            for (; k + 8 <= K; k += 8) for (int j = 0; j < 8; j++) acc[j] += col[k + j] * wr[k + j];
            float s = ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
            for (; k < K; k++) s += col[k] * wr[k];
            *(float *)AT(d, ow, oh, oc, n) = s;
        }
    }
    free(col); free(wf);
}
static void k_pool(struct ggml_tensor * d) {
    const struct ggml_tensor * a = d->src[0];
    for (int64_t n = 0; n < a->ne[3]; n++) for (int64_t c = 0; c < a->ne[2]; c++) {
        double s = 0.0;
        for (int64_t y = 0; y < a->ne[1]; y++) for (int64_t x = 0; x < a->ne[0]; x++) s += (double)ld(a, x, y, c, n);
        *(float *)AT(d, 0, 0, c, n) = (float)(s / (double)(a->ne[0] * a->ne[1]));
    }
}

static const char * op_name(enum ggml_op op) {
    static const char * names[GGML_OP_COUNT] = {"NONE", "ADD", "SUB", "MUL", "DIV", "SQRT", "SILU", "TANH", "NORM", "SOFT_MAX", "MUL_MAT", "REPEAT", "CONCAT",
                                                "GET_ROWS", "CONT", "RESHAPE", "VIEW", "PERMUTE", "TRANSPOSE", "CONV_2D", "CONV_DEPTHWISE_2D", "POOL_MEAN_HW",
                                                "ARGMAX"};
    return op < GGML_OP_COUNT ? names[op] : "?";
}

/* one line per node: index, op, shape, sum and sum of |x| in double -- the same format libggml_b200 writes with GGML_B200_DUMP_NODES */
static void dump_nodes(const struct ggml_cgraph * g, const char * path) {
    FILE * f = fopen(path, "w");
    if (!f) return;
    for (int i = 0; i < g->n_nodes; i++) {
        const struct ggml_tensor * t = g->nodes[i];
        double s = 0.0, sa = 0.0;
        const int is_view = t->op == GGML_OP_RESHAPE || t->op == GGML_OP_VIEW || t->op == GGML_OP_PERMUTE || t->op == GGML_OP_TRANSPOSE;
        if (!is_view && (t->type == GGML_TYPE_F32 || t->type == GGML_TYPE_F16))
            for (int64_t i3 = 0; i3 < t->ne[3]; i3++) for (int64_t i2 = 0; i2 < t->ne[2]; i2++) for (int64_t i1 = 0; i1 < t->ne[1]; i1++) for (int64_t i0 = 0; i0 < t->ne[0]; i0++) {
                const double v = (double)ld(t, i0, i1, i2, i3);
                s += v; sa += fabs(v);
            }
        fprintf(f, "%d %s %lld %lld %lld %lld %.9e %.9e\n", i, op_name(t->op), (long long)t->ne[0], (long long)t->ne[1], (long long)t->ne[2], (long long)t->ne[3], s, sa);
    }
    fclose(f);
}

void ggml_graph_compute_with_ctx(struct ggml_context * ctx, struct ggml_cgraph * g, int n_threads) {
    (void)ctx; (void)n_threads;
    fix_graph(g);
    for (int i = 0; i < g->n_nodes; i++) {
        struct ggml_tensor * t = g->nodes[i];
        switch (t->op) {
            case GGML_OP_ADD: case GGML_OP_SUB: case GGML_OP_MUL: case GGML_OP_DIV: k_binary(t); break;
            case GGML_OP_SQRT: case GGML_OP_SILU: case GGML_OP_TANH: k_unary(t); break;
            case GGML_OP_NORM: k_norm(t); break;
            case GGML_OP_SOFT_MAX: k_soft_max(t); break;
            case GGML_OP_MUL_MAT: k_mul_mat(t); break;
            case GGML_OP_REPEAT: k_repeat(t); break;
            case GGML_OP_CONCAT: k_concat(t); break;
            case GGML_OP_GET_ROWS: k_get_rows(t); break;
            case GGML_OP_CONT: k_cont(t); break;
            case GGML_OP_RESHAPE: case GGML_OP_VIEW: case GGML_OP_PERMUTE: case GGML_OP_TRANSPOSE: break;  /* views */
            case GGML_OP_CONV_2D: k_conv(t, 0); break;
            case GGML_OP_CONV_DEPTHWISE_2D: k_conv(t, 1); break;
            case GGML_OP_POOL_MEAN_HW: k_pool(t); break;
            case GGML_OP_ARGMAX: k_argmax(t); break;
            default: fprintf(stderr, "ggml_cpu_ref: op %d not implemented\n", (int)t->op); abort();
        }
    }
    const char * dump = getenv("GGML_CPU_REF_DUMP");
    if (dump && g->n_nodes > 256) dump_nodes(g, dump);  /* the MobileViT graph, not the tiny weight-transpose graphs */
}
