// ggml_b200.cpp -- host side of the drop-in boundary (include/ggml/ggml.h).
//
// Replaces upstream ggml's L1 tensor runtime for the two reference programs
// (/root/reference/mobilevit/main.cpp, /root/reference/rnn_text_gen/rnn_text_generation.cpp):
// the arena context, tensor records, op constructors and graph recording.  Nothing here computes:
// op constructors only record nodes; ggml_graph_compute_with_ctx() hands the recorded graph to the
// device plan builder (plan.cpp -> fuse.cpp / exec_exact.cu).
#include <time.h>

#include <cmath>
#include <cstring>
#include <unordered_set>

#include "internal.h"

using namespace b200;

// ------------------------------------------------------------------------------------------------------
// context / arena  (main.cpp:605-607,656-658,699; rnn.cpp:98-103,271-276)
// ------------------------------------------------------------------------------------------------------
extern "C" struct ggml_context * ggml_init(struct ggml_init_params params) {
    ggml_context * ctx = (ggml_context *)calloc(1, sizeof(ggml_context));
    if (!ctx) return nullptr;
    ctx->mem_size = params.mem_size < 4096 ? 4096 : params.mem_size;
    ctx->owned    = params.mem_buffer == nullptr;
    ctx->no_alloc = params.no_alloc;
    if (ctx->owned) {
        // Untouched pages are never committed, so a 1 GiB arena (main.cpp:605) costs only what is used:
        // in this library intermediates live on the device and take no host arena space.
        if (posix_memalign((void **)&ctx->mem_buffer, 4096, ctx->mem_size) != 0) {
            free(ctx);
            return nullptr;  // main.cpp:659 checks for NULL
        }
    } else {
        ctx->mem_buffer = (char *)params.mem_buffer;
    }
    return ctx;
}

extern "C" void ggml_free(struct ggml_context * ctx) {
    if (!ctx) return;
    destroy_plans_of(ctx);
    if (ctx->owned) free(ctx->mem_buffer);
    free(ctx);
}

extern "C" size_t ggml_used_mem(const struct ggml_context * ctx) { return ctx->used; }

void * b200::arena_alloc(ggml_context * ctx, size_t bytes, size_t align) {
    size_t off = (ctx->used + align - 1) / align * align;
    if (off + bytes > ctx->mem_size) {
        // upstream: "ggml_new_object: not enough space in the context's memory pool" + assert
        fprintf(stderr, "ggml_b200: not enough space in the context's memory pool (needed %zu, available %zu)\n",
                off + bytes, ctx->mem_size);
        abort();
    }
    ctx->used = off + bytes;
    ctx->n_objects++;
    return ctx->mem_buffer + off;
}

// ------------------------------------------------------------------------------------------------------
// timers (main.cpp:639-641,651,689-698)
// ------------------------------------------------------------------------------------------------------
static int64_t g_time_origin_us = 0;
static int64_t now_us() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (int64_t)ts.tv_sec * 1000000 + ts.tv_nsec / 1000;
}
extern "C" void    ggml_time_init(void) { g_time_origin_us = now_us(); }
extern "C" int64_t ggml_time_us(void) { return now_us() - g_time_origin_us; }
extern "C" int64_t ggml_time_ms(void) { return ggml_time_us() / 1000; }

// ------------------------------------------------------------------------------------------------------
// fp16 (main.cpp:929-930): IEEE binary16, round-to-nearest-even, like upstream's F16C path.
// ------------------------------------------------------------------------------------------------------
extern "C" ggml_fp16_t ggml_fp32_to_fp16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t absx = x & 0x7fffffffu;
    if (absx >= 0x7f800000u) return (ggml_fp16_t)(sign | 0x7c00u | (absx > 0x7f800000u ? 0x0200u : 0u));  // inf / nan
    if (absx >= 0x477ff000u) return (ggml_fp16_t)(sign | 0x7c00u);                                       // overflow -> inf
    if (absx < 0x33000001u) return (ggml_fp16_t)sign;                                                    // underflow -> 0
    int32_t  e = (int32_t)(absx >> 23) - 127;
    uint32_t m = (absx & 0x7fffffu) | 0x800000u;
    int      shift;
    uint32_t he;
    if (e < -14) {  // subnormal half
        shift = 13 + (-14 - e);
        he    = 0;
    } else {
        shift = 13;
        he    = (uint32_t)(e + 15);
    }
    uint32_t hm   = m >> shift;
    uint32_t rem  = m & ((1u << shift) - 1);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (hm & 1))) hm++;
    uint32_t h = (e < -14) ? hm : ((he << 10) + (hm - 0x400u));  // mantissa carry propagates into the exponent
    return (ggml_fp16_t)(sign | h);
}

extern "C" float ggml_fp16_to_fp32(ggml_fp16_t h) {
    const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t       e    = (h >> 10) & 0x1f;
    uint32_t       m    = h & 0x3ff;
    uint32_t       x;
    if (e == 0) {
        if (m == 0) {
            x = sign;
        } else {
            int s = 0;
            while (!(m & 0x400)) { m <<= 1; s++; }
            m &= 0x3ff;
            x = sign | ((uint32_t)(127 - 15 - s + 1) << 23) | (m << 13);
        }
    } else if (e == 31) {
        x = sign | 0x7f800000u | (m << 13);
    } else {
        x = sign | ((e + 112) << 23) | (m << 13);
    }
    float f;
    memcpy(&f, &x, 4);
    return f;
}
extern "C" void ggml_fp16_to_fp32_row(const ggml_fp16_t * x, float * y, int n) {
    for (int i = 0; i < n; i++) y[i] = ggml_fp16_to_fp32(x[i]);
}
extern "C" void ggml_fp32_to_fp16_row(const float * x, ggml_fp16_t * y, int n) {
    for (int i = 0; i < n; i++) y[i] = ggml_fp32_to_fp16(x[i]);
}

// ------------------------------------------------------------------------------------------------------
// tensors
// ------------------------------------------------------------------------------------------------------
extern "C" size_t ggml_type_size(enum ggml_type type) {
    switch (type) {
        case GGML_TYPE_F32: return 4;
        case GGML_TYPE_F16: return 2;
        case GGML_TYPE_I32: return 4;
        default: B200_ABORT("unsupported ggml_type %d", (int)type);
    }
    return 0;
}
extern "C" int64_t ggml_nelements(const struct ggml_tensor * t) { return t->ne[0] * t->ne[1] * t->ne[2] * t->ne[3]; }
extern "C" size_t  ggml_nbytes(const struct ggml_tensor * t) {
    // upstream: extent of the (possibly strided) tensor
    size_t nbytes = ggml_type_size(t->type);
    for (int i = 0; i < GGML_MAX_DIMS; i++) nbytes += (size_t)(t->ne[i] - 1) * t->nb[i];
    return nbytes;
}
extern "C" int ggml_n_dims(const struct ggml_tensor * t) {
    for (int i = GGML_MAX_DIMS - 1; i >= 1; --i)
        if (t->ne[i] > 1) return i + 1;
    return 1;
}
extern "C" bool ggml_is_contiguous(const struct ggml_tensor * t) {
    size_t expect = ggml_type_size(t->type);
    for (int i = 0; i < GGML_MAX_DIMS; i++) {
        if (t->ne[i] != 1 && t->nb[i] != expect) return false;
        expect *= (size_t)t->ne[i];
    }
    return true;
}
static bool same_shape(const ggml_tensor * a, const ggml_tensor * b) {
    return a->ne[0] == b->ne[0] && a->ne[1] == b->ne[1] && a->ne[2] == b->ne[2] && a->ne[3] == b->ne[3];
}
static bool can_repeat(const ggml_tensor * a, const ggml_tensor * b) {  // a tiles into b
    for (int i = 0; i < 4; i++)
        if (a->ne[i] <= 0 || b->ne[i] % a->ne[i] != 0) return false;
    return true;
}

static ggml_tensor * new_tensor_impl(ggml_context * ctx, enum ggml_type type, int n_dims, const int64_t * ne,
                                     ggml_tensor * view_src, size_t view_offs, bool alloc_data) {
    GGML_ASSERT(n_dims >= 1 && n_dims <= GGML_MAX_DIMS);
    ggml_tensor * t = (ggml_tensor *)arena_alloc(ctx, sizeof(ggml_tensor), 16);
    memset(t, 0, sizeof(*t));
    t->type   = type;
    t->n_dims = n_dims;
    for (int i = 0; i < 4; i++) t->ne[i] = i < n_dims ? ne[i] : 1;
    t->nb[0] = ggml_type_size(type);
    for (int i = 1; i < 4; i++) t->nb[i] = t->nb[i - 1] * (size_t)t->ne[i - 1];
    t->op        = GGML_OP_NONE;
    t->ctx       = ctx;
    t->view_src  = view_src;
    t->view_offs = view_offs;
    if (view_src) {
        t->data = view_src->data ? (char *)view_src->data + view_offs : nullptr;
    } else if (alloc_data && !ctx->no_alloc) {
        t->data = arena_alloc(ctx, (size_t)ggml_nelements(t) * ggml_type_size(type), 64);
    }
    return t;
}

extern "C" struct ggml_tensor * ggml_new_tensor_1d(struct ggml_context * ctx, enum ggml_type type, int64_t ne0) {
    return new_tensor_impl(ctx, type, 1, &ne0, nullptr, 0, true);
}
extern "C" struct ggml_tensor * ggml_new_tensor_2d(struct ggml_context * ctx, enum ggml_type type, int64_t ne0,
                                                   int64_t ne1) {
    const int64_t ne[2] = {ne0, ne1};
    return new_tensor_impl(ctx, type, 2, ne, nullptr, 0, true);
}
extern "C" struct ggml_tensor * ggml_new_tensor_3d(struct ggml_context * ctx, enum ggml_type type, int64_t ne0,
                                                   int64_t ne1, int64_t ne2) {
    const int64_t ne[3] = {ne0, ne1, ne2};
    return new_tensor_impl(ctx, type, 3, ne, nullptr, 0, true);
}
extern "C" struct ggml_tensor * ggml_new_tensor_4d(struct ggml_context * ctx, enum ggml_type type, int64_t ne0,
                                                   int64_t ne1, int64_t ne2, int64_t ne3) {
    const int64_t ne[4] = {ne0, ne1, ne2, ne3};
    return new_tensor_impl(ctx, type, 4, ne, nullptr, 0, true);
}
extern "C" struct ggml_tensor * ggml_new_f32(struct ggml_context * ctx, float value) {
    // main.cpp:833,1076; rnn.cpp:244 -- a 1-element F32 leaf.  It always owns host data, even in a
    // no_alloc context, because its value is part of the graph definition.
    const int64_t ne0 = 1;
    ggml_tensor * t   = new_tensor_impl(ctx, GGML_TYPE_F32, 1, &ne0, nullptr, 0, false);
    t->data           = arena_alloc(ctx, sizeof(float), 16);
    *(float *)t->data = value;
    t->flags |= 0x100;  // scalar-constant marker (treated as a plan constant)
    return t;
}
extern "C" struct ggml_tensor * ggml_set_name(struct ggml_tensor * t, const char * name) {
    strncpy(t->name, name, sizeof(t->name) - 1);
    t->name[sizeof(t->name) - 1] = 0;
    return t;
}
extern "C" void ggml_set_input(struct ggml_tensor * t) { t->flags |= GGML_TENSOR_FLAG_INPUT; }
extern "C" void ggml_set_output(struct ggml_tensor * t) { t->flags |= GGML_TENSOR_FLAG_OUTPUT; }
extern "C" void ggml_set_param(struct ggml_context *, struct ggml_tensor * t) {
    // rnn.cpp:150-152,286-287 mark leafs as "params" (autograd roots upstream).  Inference only here; the
    // flag is still recorded: a param leaf may be rewritten between computes, so it is re-uploaded.
    t->flags |= GGML_TENSOR_FLAG_PARAM;
}

extern "C" void *  ggml_get_data(const struct ggml_tensor * t) { return t->data; }
extern "C" float * ggml_get_data_f32(const struct ggml_tensor * t) {
    GGML_ASSERT(t->type == GGML_TYPE_F32);
    return (float *)t->data;
}
extern "C" void ggml_set_i32_1d(const struct ggml_tensor * t, int i, int32_t v) {
    GGML_ASSERT(t->data != nullptr);
    switch (t->type) {
        case GGML_TYPE_I32: ((int32_t *)t->data)[i] = v; break;
        case GGML_TYPE_F32: ((float *)t->data)[i] = (float)v; break;
        case GGML_TYPE_F16: ((ggml_fp16_t *)t->data)[i] = ggml_fp32_to_fp16((float)v); break;
        default: GGML_ASSERT(false);
    }
}
extern "C" int32_t ggml_get_i32_1d(const struct ggml_tensor * t, int i) {
    GGML_ASSERT(t->data != nullptr);
    switch (t->type) {
        case GGML_TYPE_I32: return ((int32_t *)t->data)[i];
        case GGML_TYPE_F32: return (int32_t)((float *)t->data)[i];
        case GGML_TYPE_F16: return (int32_t)ggml_fp16_to_fp32(((ggml_fp16_t *)t->data)[i]);
        default: GGML_ASSERT(false);
    }
    return 0;
}
extern "C" void ggml_set_f32_1d(const struct ggml_tensor * t, int i, float v) {
    GGML_ASSERT(t->data != nullptr);
    switch (t->type) {
        case GGML_TYPE_I32: ((int32_t *)t->data)[i] = (int32_t)v; break;
        case GGML_TYPE_F32: ((float *)t->data)[i] = v; break;
        case GGML_TYPE_F16: ((ggml_fp16_t *)t->data)[i] = ggml_fp32_to_fp16(v); break;
        default: GGML_ASSERT(false);
    }
}
extern "C" float ggml_get_f32_1d(const struct ggml_tensor * t, int i) {
    if (t->data == nullptr) B200_ABORT("ggml_get_f32_1d: tensor '%s' has no host data (not a leaf or a computed graph output)", t->name);
    switch (t->type) {
        case GGML_TYPE_I32: return (float)((int32_t *)t->data)[i];
        case GGML_TYPE_F32: return ((float *)t->data)[i];
        case GGML_TYPE_F16: return ggml_fp16_to_fp32(((ggml_fp16_t *)t->data)[i]);
        default: GGML_ASSERT(false);
    }
    return 0.f;
}

// ------------------------------------------------------------------------------------------------------
// op constructors: record only.
// ------------------------------------------------------------------------------------------------------
static ggml_tensor * new_result(ggml_context * ctx, enum ggml_type type, const int64_t * ne, enum ggml_op op,
                                ggml_tensor * a, ggml_tensor * b) {
    int n_dims = 4;
    while (n_dims > 1 && ne[n_dims - 1] == 1) n_dims--;
    int64_t ne4[4] = {ne[0], ne[1], ne[2], ne[3]};
    ggml_tensor * r = new_tensor_impl(ctx, type, 4, ne4, nullptr, 0, false);
    r->n_dims = n_dims;
    r->op     = op;
    r->src[0] = a;
    r->src[1] = b;
    return r;
}

static ggml_tensor * binary_op(ggml_context * ctx, enum ggml_op op, ggml_tensor * a, ggml_tensor * b) {
    GGML_ASSERT(can_repeat(b, a));  // upstream: b is broadcast onto a
    GGML_ASSERT(a->type == GGML_TYPE_F32 && b->type == GGML_TYPE_F32);
    return new_result(ctx, GGML_TYPE_F32, a->ne, op, a, b);
}
extern "C" struct ggml_tensor * ggml_add(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b) {
    return binary_op(ctx, GGML_OP_ADD, a, b);
}
extern "C" struct ggml_tensor * ggml_sub(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b) {
    return binary_op(ctx, GGML_OP_SUB, a, b);
}
extern "C" struct ggml_tensor * ggml_mul(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b) {
    return binary_op(ctx, GGML_OP_MUL, a, b);
}
extern "C" struct ggml_tensor * ggml_div(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b) {
    return binary_op(ctx, GGML_OP_DIV, a, b);
}

static ggml_tensor * unary_op(ggml_context * ctx, enum ggml_op op, ggml_tensor * a) {
    GGML_ASSERT(a->type == GGML_TYPE_F32);
    return new_result(ctx, GGML_TYPE_F32, a->ne, op, a, nullptr);
}
extern "C" struct ggml_tensor * ggml_sqrt(struct ggml_context * ctx, struct ggml_tensor * a) { return unary_op(ctx, GGML_OP_SQRT, a); }
extern "C" struct ggml_tensor * ggml_silu(struct ggml_context * ctx, struct ggml_tensor * a) { return unary_op(ctx, GGML_OP_SILU, a); }
extern "C" struct ggml_tensor * ggml_tanh(struct ggml_context * ctx, struct ggml_tensor * a) { return unary_op(ctx, GGML_OP_TANH, a); }
extern "C" struct ggml_tensor * ggml_soft_max(struct ggml_context * ctx, struct ggml_tensor * a) { return unary_op(ctx, GGML_OP_SOFT_MAX, a); }
extern "C" struct ggml_tensor * ggml_norm(struct ggml_context * ctx, struct ggml_tensor * a, float eps) {
    ggml_tensor * r = unary_op(ctx, GGML_OP_NORM, a);
    memcpy(r->op_params, &eps, sizeof(float));
    return r;
}

extern "C" struct ggml_tensor * ggml_mul_mat(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b) {
    // upstream ggml_can_mul_mat: same K, a broadcastable over dims 2,3
    GGML_ASSERT(a->ne[0] == b->ne[0]);
    GGML_ASSERT(b->ne[2] % a->ne[2] == 0 && b->ne[3] % a->ne[3] == 0);
    GGML_ASSERT(b->type == GGML_TYPE_F32);
    const int64_t ne[4] = {a->ne[1], b->ne[1], b->ne[2], b->ne[3]};
    return new_result(ctx, GGML_TYPE_F32, ne, GGML_OP_MUL_MAT, a, b);
}

extern "C" struct ggml_tensor * ggml_repeat(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b) {
    GGML_ASSERT(can_repeat(a, b));
    return new_result(ctx, a->type, b->ne, GGML_OP_REPEAT, a, b);
}

extern "C" struct ggml_tensor * ggml_concat(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b) {
    // main.cpp:1219 -- the 2-argument form of early-2024 upstream: concatenation along dim 2
    GGML_ASSERT(a->ne[0] == b->ne[0] && a->ne[1] == b->ne[1] && a->ne[3] == b->ne[3]);
    GGML_ASSERT(a->type == GGML_TYPE_F32 && b->type == GGML_TYPE_F32);
    const int64_t ne[4] = {a->ne[0], a->ne[1], a->ne[2] + b->ne[2], a->ne[3]};
    return new_result(ctx, GGML_TYPE_F32, ne, GGML_OP_CONCAT, a, b);
}

extern "C" struct ggml_tensor * ggml_get_rows(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b) {
    // rnn.cpp:47,200 -- rows of a selected by the I32 vector b; result F32 [a.ne0, b.ne0]
    GGML_ASSERT(b->type == GGML_TYPE_I32);
    GGML_ASSERT(a->ne[2] == 1 && a->ne[3] == 1 && b->ne[1] == 1 && b->ne[2] == 1 && b->ne[3] == 1);
    const int64_t ne[4] = {a->ne[0], b->ne[0], 1, 1};
    return new_result(ctx, GGML_TYPE_F32, ne, GGML_OP_GET_ROWS, a, b);
}

extern "C" struct ggml_tensor * ggml_cont(struct ggml_context * ctx, struct ggml_tensor * a) {
    return new_result(ctx, a->type, a->ne, GGML_OP_CONT, a, nullptr);
}
extern "C" struct ggml_tensor * ggml_cont_4d(struct ggml_context * ctx, struct ggml_tensor * a, int64_t ne0, int64_t ne1,
                                             int64_t ne2, int64_t ne3) {
    GGML_ASSERT(ggml_nelements(a) == ne0 * ne1 * ne2 * ne3);
    const int64_t ne[4] = {ne0, ne1, ne2, ne3};
    return new_result(ctx, a->type, ne, GGML_OP_CONT, a, nullptr);
}

static ggml_tensor * new_view(ggml_context * ctx, ggml_tensor * a, enum ggml_op op, const int64_t * ne, const size_t * nb) {
    ggml_tensor * base = a->view_src ? a->view_src : a;
    ggml_tensor * r    = new_tensor_impl(ctx, a->type, 4, ne, base, a->view_offs, false);
    int n_dims = 4;
    while (n_dims > 1 && ne[n_dims - 1] == 1) n_dims--;
    r->n_dims = n_dims;
    if (nb)
        for (int i = 0; i < 4; i++) r->nb[i] = nb[i];
    r->op     = op;
    r->src[0] = a;
    return r;
}

static ggml_tensor * reshape_impl(ggml_context * ctx, ggml_tensor * a, int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3) {
    GGML_ASSERT(ggml_is_contiguous(a));
    GGML_ASSERT(ggml_nelements(a) == ne0 * ne1 * ne2 * ne3);
    const int64_t ne[4] = {ne0, ne1, ne2, ne3};
    return new_view(ctx, a, GGML_OP_RESHAPE, ne, nullptr);
}
extern "C" struct ggml_tensor * ggml_reshape_2d(struct ggml_context * ctx, struct ggml_tensor * a, int64_t ne0, int64_t ne1) {
    return reshape_impl(ctx, a, ne0, ne1, 1, 1);
}
extern "C" struct ggml_tensor * ggml_reshape_3d(struct ggml_context * ctx, struct ggml_tensor * a, int64_t ne0, int64_t ne1,
                                                int64_t ne2) {
    return reshape_impl(ctx, a, ne0, ne1, ne2, 1);
}
extern "C" struct ggml_tensor * ggml_reshape_4d(struct ggml_context * ctx, struct ggml_tensor * a, int64_t ne0, int64_t ne1,
                                                int64_t ne2, int64_t ne3) {
    return reshape_impl(ctx, a, ne0, ne1, ne2, ne3);
}

extern "C" struct ggml_tensor * ggml_permute(struct ggml_context * ctx, struct ggml_tensor * a, int axis0, int axis1,
                                             int axis2, int axis3) {
    const int ax[4] = {axis0, axis1, axis2, axis3};
    for (int i = 0; i < 4; i++) GGML_ASSERT(ax[i] >= 0 && ax[i] < 4);
    GGML_ASSERT(axis0 != axis1 && axis0 != axis2 && axis0 != axis3 && axis1 != axis2 && axis1 != axis3 && axis2 != axis3);
    int64_t ne[4];
    size_t  nb[4];
    for (int i = 0; i < 4; i++) {  // upstream: result.ne[axis_i] = a.ne[i]
        ne[ax[i]] = a->ne[i];
        nb[ax[i]] = a->nb[i];
    }
    ggml_tensor * r = new_view(ctx, a, GGML_OP_PERMUTE, ne, nb);
    for (int i = 0; i < 4; i++) r->op_params[i] = ax[i];
    return r;
}
extern "C" struct ggml_tensor * ggml_transpose(struct ggml_context * ctx, struct ggml_tensor * a) {
    const int64_t ne[4] = {a->ne[1], a->ne[0], a->ne[2], a->ne[3]};
    const size_t  nb[4] = {a->nb[1], a->nb[0], a->nb[2], a->nb[3]};
    return new_view(ctx, a, GGML_OP_TRANSPOSE, ne, nb);
}

extern "C" struct ggml_tensor * ggml_view_2d(struct ggml_context * ctx, struct ggml_tensor * a, int64_t ne0, int64_t ne1, size_t nb1,
                                             size_t offset) {
    const int64_t ne[4] = {ne0, ne1, 1, 1};
    const size_t  nb[4] = {a->nb[0], nb1, nb1 * (size_t)ne1, nb1 * (size_t)ne1};
    ggml_tensor * base  = a->view_src ? a->view_src : a;
    ggml_tensor * r     = new_tensor_impl(ctx, a->type, 2, ne, base, a->view_offs + offset, false);
    for (int i = 0; i < 4; i++) r->nb[i] = nb[i];
    r->op     = GGML_OP_VIEW;
    r->src[0] = a;
    return r;
}

extern "C" struct ggml_tensor * ggml_argmax(struct ggml_context * ctx, struct ggml_tensor * a) {
    GGML_ASSERT(a->type == GGML_TYPE_F32 && a->ne[2] == 1 && a->ne[3] == 1);
    const int64_t ne[4] = {a->ne[1], 1, 1, 1};
    return new_result(ctx, GGML_TYPE_I32, ne, GGML_OP_ARGMAX, a, nullptr);
}

static int64_t conv_out(int64_t in, int64_t k, int s, int p, int d) { return (in + 2 * p - d * (k - 1) - 1) / s + 1; }

static ggml_tensor * conv_impl(ggml_context * ctx, enum ggml_op op, ggml_tensor * a, ggml_tensor * b, int s0, int s1,
                               int p0, int p1, int d0, int d1) {
    GGML_ASSERT(b->type == GGML_TYPE_F32);
    GGML_ASSERT(a->type == GGML_TYPE_F16 || a->type == GGML_TYPE_F32);
    int64_t oc;
    if (op == GGML_OP_CONV_2D) {
        GGML_ASSERT(a->ne[2] == b->ne[2]);
        oc = a->ne[3];
    } else {
        GGML_ASSERT(a->ne[2] == 1 && a->ne[3] == b->ne[2]);
        oc = b->ne[2];
    }
    const int64_t ne[4] = {conv_out(b->ne[0], a->ne[0], s0, p0, d0), conv_out(b->ne[1], a->ne[1], s1, p1, d1), oc, b->ne[3]};
    ggml_tensor * r = new_result(ctx, GGML_TYPE_F32, ne, op, a, b);
    const int32_t prm[6] = {s0, s1, p0, p1, d0, d1};
    memcpy(r->op_params, prm, sizeof(prm));
    return r;
}
extern "C" struct ggml_tensor * ggml_conv_2d(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b, int s0,
                                             int s1, int p0, int p1, int d0, int d1) {
    return conv_impl(ctx, GGML_OP_CONV_2D, a, b, s0, s1, p0, p1, d0, d1);
}
extern "C" struct ggml_tensor * ggml_conv_depthwise_2d(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b,
                                                       int s0, int s1, int p0, int p1, int d0, int d1) {
    return conv_impl(ctx, GGML_OP_CONV_DEPTHWISE_2D, a, b, s0, s1, p0, p1, d0, d1);
}

extern "C" struct ggml_tensor * ggml_b200_pool_mean_hw(struct ggml_context * ctx, struct ggml_tensor * a) {
    GGML_ASSERT(a->type == GGML_TYPE_F32);
    const int64_t ne[4] = {1, 1, a->ne[2], a->ne[3]};
    return new_result(ctx, GGML_TYPE_F32, ne, GGML_OP_POOL_MEAN_HW, a, nullptr);
}

// ------------------------------------------------------------------------------------------------------
// graph (main.cpp:608,636; rnn.cpp:149,154-156,289)
// ------------------------------------------------------------------------------------------------------
void b200::fix_graph_pointers(ggml_cgraph * gf) {
    // A by-value ggml_cgraph (`ggml_cgraph gf = {}` / returned by ggml_build_forward) carries its node
    // arrays inline; re-point after any struct copy.
    if (gf->size <= GGML_B200_STATIC_GRAPH_NODES) {
        gf->size  = GGML_B200_STATIC_GRAPH_NODES;
        gf->nodes = gf->static_nodes;
        gf->leafs = gf->static_leafs;
    }
}

extern "C" struct ggml_cgraph * ggml_new_graph(struct ggml_context * ctx) {
    ggml_cgraph * gf = (ggml_cgraph *)arena_alloc(ctx, sizeof(ggml_cgraph), 16);
    memset(gf, 0, sizeof(*gf));
    const int size = 4 * GGML_DEFAULT_GRAPH_SIZE;  // headroom over upstream's 2048 for batched builders
    gf->size  = size;
    gf->nodes = (ggml_tensor **)arena_alloc(ctx, sizeof(ggml_tensor *) * size, 16);
    gf->leafs = (ggml_tensor **)arena_alloc(ctx, sizeof(ggml_tensor *) * size, 16);
    return gf;
}

static void visit(ggml_cgraph * gf, ggml_tensor * t, std::unordered_set<const ggml_tensor *> & seen) {
    if (!t || seen.count(t)) return;
    seen.insert(t);
    for (int i = 0; i < GGML_MAX_SRC; i++) visit(gf, t->src[i], seen);
    if (t->view_src) visit(gf, t->view_src, seen);
    if (t->op == GGML_OP_NONE) {
        GGML_ASSERT(gf->n_leafs < gf->size);
        gf->leafs[gf->n_leafs++] = t;
    } else {
        GGML_ASSERT(gf->n_nodes < gf->size);
        gf->nodes[gf->n_nodes++] = t;
    }
}

extern "C" void ggml_build_forward_expand(struct ggml_cgraph * gf, struct ggml_tensor * tensor) {
    fix_graph_pointers(gf);
    std::unordered_set<const ggml_tensor *> seen;
    for (int i = 0; i < gf->n_nodes; i++) seen.insert(gf->nodes[i]);
    for (int i = 0; i < gf->n_leafs; i++) seen.insert(gf->leafs[i]);
    visit(gf, tensor, seen);
    // the tensors handed to build_forward_expand are the graph's outputs: they get host shadows
    tensor->flags |= GGML_TENSOR_FLAG_OUTPUT;
}

extern "C" struct ggml_cgraph ggml_build_forward(struct ggml_tensor * tensor) {
    ggml_cgraph gf;
    memset(&gf, 0, sizeof(gf));
    ggml_build_forward_expand(&gf, tensor);
    return gf;  // pointers are re-fixed on next use
}

// GGML_B200_DUMP_NODES=<file> (per-node plans only, graphs of more than 256 nodes): after the compute, one line per node --
// index, op, shape, sum and sum of |x| in double -- in the format the CPU run of the same program writes (test infrastructure, GGML_CPU_REF_DUMP), so the
// unmodified reference program can be compared node by node between its CPU run and its run on this library.
static void dump_nodes(Plan * plan, ggml_cgraph * gf, const char * path) {
    static const char * names[GGML_OP_COUNT] = {"NONE", "ADD", "SUB", "MUL", "DIV", "SQRT", "SILU", "TANH", "NORM", "SOFT_MAX", "MUL_MAT", "REPEAT", "CONCAT",
                                                "GET_ROWS", "CONT", "RESHAPE", "VIEW", "PERMUTE", "TRANSPOSE", "CONV_2D", "CONV_DEPTHWISE_2D", "POOL_MEAN_HW",
                                                "ARGMAX"};
    FILE * f = fopen(path, "w");
    if (!f) return;
    B200_CHECK(cudaDeviceSynchronize());
    std::vector<char> host;
    for (int i = 0; i < gf->n_nodes; i++) {
        const ggml_tensor * t = gf->nodes[i];
        double s = 0.0, sa = 0.0;
        auto it = plan->slots.find(t);
        if (!is_view_op(t->op) && it != plan->slots.end() && it->second.dptr && (t->type == GGML_TYPE_F32 || t->type == GGML_TYPE_F16)) {
            const size_t n = (size_t)ggml_nelements(t);
            host.resize(n * ggml_type_size(t->type));
            B200_CHECK(cudaMemcpy(host.data(), it->second.dptr, host.size(), cudaMemcpyDeviceToHost));
            for (size_t k = 0; k < n; k++) {
                const double v = t->type == GGML_TYPE_F32 ? (double)((const float *)host.data())[k] : (double)ggml_fp16_to_fp32(((const ggml_fp16_t *)host.data())[k]);
                s += v;
                sa += fabs(v);
            }
        }
        fprintf(f, "%d %s %lld %lld %lld %lld %.9e %.9e\n", i, names[t->op], (long long)t->ne[0], (long long)t->ne[1], (long long)t->ne[2], (long long)t->ne[3], s, sa);
    }
    fclose(f);
}

extern "C" void ggml_graph_compute_with_ctx(struct ggml_context * ctx, struct ggml_cgraph * gf, int n_threads) {
    (void)n_threads;  // main.cpp:640 passes 1; device execution ignores it (SURVEY 8b "Threading")
    fix_graph_pointers(gf);
    Plan * plan = get_or_build_plan(ctx, gf);
    run_plan(plan);
    if (const char * dump = getenv("GGML_B200_DUMP_NODES"))
        if (plan->mode != GGML_B200_MODE_FAST && gf->n_nodes > 256) dump_nodes(plan, gf, dump);
}

extern "C" void ggml_b200_graph_prepare(struct ggml_context * ctx, struct ggml_cgraph * gf) {
    fix_graph_pointers(gf);
    (void)get_or_build_plan(ctx, gf);
}

extern "C" const char * ggml_b200_version(void) { return "ggml_b200 0.1 (sm_100a)"; }
