// fuse.cpp -- FAST mode planner: pattern-matches the recorded ggml graph of the MobileViT forward pass
// (the op sequences of /root/reference/mobilevit/main.cpp:721-870,963-1223, for any batch) into a short plan of
// fused kernels over NHWC f16 activations.
//
//   ggml nodes (~1.3 k)                                      ->  fused unit (one kernel launch)
//   conv_2d 1x1 + sub/div/sqrt/add/mul/add (BN) + silu [+add] ->  tcgen05 GEMM, BN/SiLU/residual epilogue      (K1)
//   conv_2d 3x3 (+concat) + BN + silu                        ->  tcgen05 implicit GEMM, 1 or 2 TMA sources     (K1)
//   conv_2d 3x3 s2 on the 3-channel image + BN + silu        ->  stem kernel                                    (K2)
//   conv_depthwise_2d + BN + silu                            ->  depthwise kernel                               (K3)
//   cont/reshape/permute chains of unfolding / folding       ->  nothing: tokens stay in NHWC pixel order
//   norm + mul + add + cont                                  ->  LayerNorm kernel (f32 in, f16 out)             (K5)
//   3x (mul_mat + add bias) for q, k, v                      ->  ONE tcgen05 GEMM against [3C, C] weights       (K6)
//   permute/mul_mat/div/soft_max/mul_mat/permute/cont        ->  attention kernel                               (K7)
//   mul_mat + add bias [+ silu] [+ add residual]             ->  tcgen05 GEMM, bias/SiLU/residual epilogue      (K8)
//   cont(permute(kernel)), BN parameter chains, transposes   ->  constant-folded on the host at plan time
//
// Every matcher verifies shapes, axes and constants; anything it does not recognise makes build_fast_plan return
// false and the caller runs the per-node EXACT plan instead (still on the GPU).
#include <cmath>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <set>

#include "fast_kernels.h"
#include "gemm_tcgen05.h"
#include "internal.h"
#include "vit_stage.cuh"

namespace b200 {
namespace {

struct FNode;
struct FVal {  // an activation in pixel-major NHWC order: [N*H*W, C]
    int     N = 0, H = 0, W = 0, C = 0;
    FNode * prod = nullptr;
    bool    need16 = false, need32 = false;
    bool    need_stats = false;  // producer GEMM also writes per-row (sum, sum of squares): a LayerNorm of this value was folded into its consumers
    float * pstats = nullptr;
    int64_t off_stats = -1;
    __half * p16 = nullptr;
    float *  p32 = nullptr;
    int64_t  off16 = -1, off32 = -1;
    int      def = -1, last = -1;
    bool     is_input = false;
    std::vector<FNode *> users;
    int64_t rows() const { return (int64_t)N * H * W; }
};

enum FKind { FK_INPUT, FK_STEM, FK_CONV1, FK_CONV3, FK_DW, FK_LN, FK_LINEAR, FK_QKV, FK_ATTN, FK_ADD, FK_POOL, FK_IR, FK_VIT };

struct FNode;
// one transformer layer absorbed into an FK_VIT node (K8): the (dead) nodes whose parameters the fused stage kernel needs
struct VitLayerNodes { FNode *ln1, *qkv, *proj, *ln2, *up, *down; };

struct BNParams {
    const ggml_tensor *mean = nullptr, *var = nullptr, *gamma = nullptr, *beta = nullptr;
    float eps = 0.f;
};

struct FNode {
    FKind kind;
    FVal * out = nullptr;
    std::vector<FVal *> in;
    FVal * res = nullptr;  // residual fused into the epilogue
    const ggml_tensor * w = nullptr, * bias = nullptr;
    bool     has_bn = false;
    BNParams bn;
    int      act = 0, stride = 1;
    const ggml_tensor *wq = nullptr, *wk = nullptr, *wv = nullptr, *bq = nullptr, *bk = nullptr, *bv = nullptr;
    int   heads = 0;
    float eps   = 0.f;
    const ggml_tensor *g = nullptr, *b = nullptr;
    const ggml_tensor * leaf = nullptr;
    bool  chw  = false;
    bool  dead = false;
    bool  fused_into_reduce = false;  // depthwise node executed inside the following reduce conv's kernel (K4a)
    // K4: a reduce conv of kind FK_IR runs the whole inverted residual; ir_expand / ir_dw are the (dead) nodes it absorbed
    FNode *ir_expand = nullptr, *ir_dw = nullptr;
    bool   fused_ir  = false;         // this (dead) node's constants are still needed: it runs inside an FK_IR kernel
    std::vector<VitLayerNodes> vit;   // K8: the layers of a fused transformer stage (kind FK_VIT)
    int    vit_F = 0;
    int   order = -1;
    // folded constants (offsets into the plan's constant pool)
    int64_t c_w = -1, c_scale = -1, c_shift = -1, c_c1 = -1, c_w2 = -1;  // c_w2: 3x3 weights in the halo layout (conv3x3_pack_halo)
    // LayerNorm folded into this GEMM (consumer side): gamma / beta of the LN that preceded it
    const ggml_tensor *ln_g = nullptr, *ln_b = nullptr;
    float ln_eps = 0.f;
    std::string name;
};

struct Fail {};  // thrown by matchers; caught in build_fast_plan
#define REQUIRE(cond) do { if (!(cond)) { if (runtime().verbose) fprintf(stderr, "libggml_b200: fused planner: '%s' failed at %s:%d\n", #cond, __FILE__, __LINE__); throw Fail(); } } while (0)

// ---- small predicates on ggml tensors ---------------------------------------------------------------------
bool is_op(const ggml_tensor * t, enum ggml_op op) { return t && t->op == op; }
// A leaf may be constant-folded only if the caller cannot rewrite it between computes: leafs flagged input / param are
// re-uploaded on every compute by the exact plan (place_leafs), so the fused plan must not freeze them either.
// The same holds for leafs of the compute context itself (place_leafs treats them as per-call inputs), except the
// ggml_new_f32 scalars (flag 0x100), which are constants by construction.
thread_local const ggml_context * t_compute_ctx = nullptr;
bool is_weight_leaf(const ggml_tensor * t) {
    if (!t || t->op != GGML_OP_NONE || t->view_src != nullptr || t->data == nullptr) return false;
    if (t->flags & 0x100) return true;
    if (t->flags & (GGML_TENSOR_FLAG_INPUT | GGML_TENSOR_FLAG_PARAM)) return false;
    return t->ctx != t_compute_ctx;
}
bool same_ne(const ggml_tensor * a, const ggml_tensor * b) {
    return a->ne[0] == b->ne[0] && a->ne[1] == b->ne[1] && a->ne[2] == b->ne[2] && a->ne[3] == b->ne[3];
}
bool permute_is(const ggml_tensor * t, int a0, int a1, int a2, int a3) {
    return is_op(t, GGML_OP_PERMUTE) && t->op_params[0] == a0 && t->op_params[1] == a1 && t->op_params[2] == a2 && t->op_params[3] == a3;
}
bool ne_is(const ggml_tensor * t, int64_t n0, int64_t n1, int64_t n2, int64_t n3) {
    return t->ne[0] == n0 && t->ne[1] == n1 && t->ne[2] == n2 && t->ne[3] == n3;
}
// REPEAT(CONT|RESHAPE(p)) with p a 1-d f32 weight leaf of `c` elements broadcast along dim `dim` of the target
const ggml_tensor * channel_param(const ggml_tensor * t, int dim, int64_t c) {
    if (!is_op(t, GGML_OP_REPEAT)) return nullptr;
    const ggml_tensor * v = t->src[0];
    if (!(is_op(v, GGML_OP_CONT) || is_op(v, GGML_OP_RESHAPE))) return nullptr;
    const ggml_tensor * p = v->src[0];
    if (!is_weight_leaf(p) || p->type != GGML_TYPE_F32 || ggml_nelements(p) != c || p->ne[0] != c) return nullptr;
    for (int i = 0; i < 4; i++)
        if (v->ne[i] != (i == dim ? c : 1)) return nullptr;
    return p;
}
bool scalar_const(const ggml_tensor * t, float * v) {
    if (!is_weight_leaf(t) || t->type != GGML_TYPE_F32 || ggml_nelements(t) != 1) return false;
    *v = *(const float *)t->data;
    return true;
}

struct ConstPool {
    std::vector<uint8_t> host;
    char *               dev = nullptr;
    int64_t add(const void * p, size_t bytes) {
        size_t off = (host.size() + 255) & ~size_t(255);
        host.resize(off + bytes);
        memcpy(host.data() + off, p, bytes);
        return (int64_t)off;
    }
    template <typename T> T * ptr(int64_t off) const { return off < 0 ? nullptr : reinterpret_cast<T *>(dev + off); }
};

struct Planner {
    Plan *        plan;
    ggml_cgraph * gf;
    std::map<const ggml_tensor *, FVal *> memo;
    std::vector<std::unique_ptr<FVal>>    vals;
    std::vector<std::unique_ptr<FNode>>   nodes;
    ConstPool                             pool;

    FVal * new_val(int N, int H, int W, int C) {
        vals.emplace_back(new FVal());
        FVal * v = vals.back().get();
        v->N = N; v->H = H; v->W = W; v->C = C;
        return v;
    }
    FNode * new_node(FKind k, FVal * out, std::string name = "") {
        nodes.emplace_back(new FNode());
        FNode * n = nodes.back().get();
        n->kind = k;
        n->out  = out;
        n->name = std::move(name);
        if (out) out->prod = n;
        return n;
    }

    // ---- matchers ------------------------------------------------------------------------------------------
    // BatchNorm chain of conv_layer::forward (main.cpp:809-846); returns the tensor it is applied to
    const ggml_tensor * match_bn(const ggml_tensor * t, BNParams & bn) {
        if (!is_op(t, GGML_OP_ADD)) return nullptr;
        const int64_t C = t->ne[2];
        bn.beta = channel_param(t->src[1], 2, C);
        const ggml_tensor * m = t->src[0];
        if (!bn.beta || !is_op(m, GGML_OP_MUL)) return nullptr;
        bn.gamma = channel_param(m->src[1], 2, C);
        const ggml_tensor * d = m->src[0];
        if (!bn.gamma || !is_op(d, GGML_OP_DIV)) return nullptr;
        const ggml_tensor * q = d->src[1];
        if (!is_op(q, GGML_OP_SQRT) || !is_op(q->src[0], GGML_OP_ADD)) return nullptr;
        bn.var = channel_param(q->src[0]->src[0], 2, C);
        if (!bn.var || !scalar_const(q->src[0]->src[1], &bn.eps)) return nullptr;
        const ggml_tensor * s = d->src[0];
        if (!is_op(s, GGML_OP_SUB)) return nullptr;
        bn.mean = channel_param(s->src[1], 2, C);
        if (!bn.mean) return nullptr;
        return s->src[0];
    }

    // dense layer: ADD(MUL_MAT(CONT(PERMUTE(w,1,0,2,3) | TRANSPOSE(w)), x), REPEAT(RESHAPE(b)))  (main.cpp:1022-1035)
    bool match_dense(const ggml_tensor * t, const ggml_tensor ** x, const ggml_tensor ** w, const ggml_tensor ** b) {
        if (!is_op(t, GGML_OP_ADD)) return false;
        const ggml_tensor * mm = t->src[0];
        if (!is_op(mm, GGML_OP_MUL_MAT)) return false;
        *b = channel_param(t->src[1], 0, t->ne[0]);
        if (!*b) return false;
        const ggml_tensor * wt = mm->src[0];
        if (!is_op(wt, GGML_OP_CONT)) return false;
        const ggml_tensor * p = wt->src[0];
        if (!(permute_is(p, 1, 0, 2, 3) || is_op(p, GGML_OP_TRANSPOSE))) return false;
        const ggml_tensor * leaf = p->src[0];
        if (!is_weight_leaf(leaf) || leaf->type != GGML_TYPE_F32 || leaf->ne[2] != 1 || leaf->ne[3] != 1) return false;
        // leaf ne = (out, in); mul_mat contracts over `in`
        if (leaf->ne[1] != mm->src[1]->ne[0] || leaf->ne[0] != t->ne[0]) return false;
        *w = leaf;
        *x = mm->src[1];
        return true;
    }

    // CONT(ADD(MUL(NORM(x), REPEAT(RESHAPE(g))), REPEAT(RESHAPE(b))))  (main.cpp:1002-1019)
    bool match_ln(const ggml_tensor * t, const ggml_tensor ** x, const ggml_tensor ** g, const ggml_tensor ** b, float * eps) {
        if (!is_op(t, GGML_OP_CONT) || !is_op(t->src[0], GGML_OP_ADD)) return false;
        const ggml_tensor * a = t->src[0];
        const ggml_tensor * m = a->src[0];
        if (!is_op(m, GGML_OP_MUL) || !is_op(m->src[0], GGML_OP_NORM)) return false;
        *g = channel_param(m->src[1], 0, t->ne[0]);
        *b = channel_param(a->src[1], 0, t->ne[0]);
        if (!*g || !*b) return false;
        memcpy(eps, m->src[0]->op_params, sizeof(float));
        *x = m->src[0]->src[0];
        return true;
    }

    // unfolding (main.cpp:721-747 generalised to N images) [+ reshape to (C, L, 4N)]: returns the (W,H,C,N) source
    const ggml_tensor * match_unfold(const ggml_tensor * t) {
        const ggml_tensor * c2 = t;
        if (is_op(t, GGML_OP_RESHAPE) && is_op(t->src[0], GGML_OP_CONT)) c2 = t->src[0];
        if (!is_op(c2, GGML_OP_CONT)) return nullptr;
        const ggml_tensor * p2 = c2->src[0];
        if (!permute_is(p2, 2, 1, 0, 3)) return nullptr;
        const ggml_tensor * r2 = p2->src[0];
        if (!is_op(r2, GGML_OP_RESHAPE) || !is_op(r2->src[0], GGML_OP_CONT)) return nullptr;
        const ggml_tensor * p1 = r2->src[0]->src[0];
        if (!permute_is(p1, 0, 2, 1, 3)) return nullptr;
        const ggml_tensor * r1 = p1->src[0];
        if (!is_op(r1, GGML_OP_RESHAPE) || !is_op(r1->src[0], GGML_OP_CONT)) return nullptr;
        const ggml_tensor * x = r1->src[0]->src[0];
        const int64_t W = x->ne[0], H = x->ne[1], C = x->ne[2], N = x->ne[3];
        if (W % 2 || H % 2) return nullptr;
        const int64_t npw = W / 2, nph = H / 2;
        if (!ne_is(r1, 2, npw, 2, N * C * nph) || !ne_is(r2, 4, npw * nph, C, N)) return nullptr;
        if (c2 == t) { if (!ne_is(t, C, npw * nph, 4, N)) return nullptr; }
        else if (!ne_is(t, C, npw * nph, 4 * N, 1)) return nullptr;
        return x;
    }

    // folding (main.cpp:750-768 generalised): t = (W,H,C,N); returns the (C, L, 4N) token tensor
    const ggml_tensor * match_fold(const ggml_tensor * t) {
        if (!is_op(t, GGML_OP_RESHAPE) || !is_op(t->src[0], GGML_OP_CONT)) return nullptr;
        const ggml_tensor * p2 = t->src[0]->src[0];
        if (!permute_is(p2, 0, 2, 1, 3)) return nullptr;
        const ggml_tensor * f2 = p2->src[0];
        if (!is_op(f2, GGML_OP_RESHAPE) || !is_op(f2->src[0], GGML_OP_CONT)) return nullptr;
        const ggml_tensor * p1 = f2->src[0]->src[0];
        if (!permute_is(p1, 2, 1, 0, 3)) return nullptr;
        const ggml_tensor * f1 = p1->src[0];
        if (!is_op(f1, GGML_OP_RESHAPE)) return nullptr;
        const ggml_tensor * x = f1->src[0];
        const int64_t W = t->ne[0], H = t->ne[1], C = t->ne[2], N = t->ne[3];
        if (W % 2 || H % 2) return nullptr;
        const int64_t npw = W / 2, nph = H / 2, L = npw * nph;
        if (!ne_is(f1, C, L, 4, N) || !ne_is(f2, 2, 2, npw, nph * C * N)) return nullptr;
        if (!(x->ne[0] == C && x->ne[1] == L && x->ne[2] * x->ne[3] == 4 * N)) return nullptr;
        return x;
    }

    // transpose_for_score (main.cpp:975-986): PERMUTE(RESHAPE(dense, [d, heads, L, B]); 0,2,1,3)
    const ggml_tensor * match_heads(const ggml_tensor * t, int * heads) {
        if (!permute_is(t, 0, 2, 1, 3) || !is_op(t->src[0], GGML_OP_RESHAPE)) return nullptr;
        const ggml_tensor * r = t->src[0];
        *heads = (int)r->ne[1];
        const ggml_tensor * x = r->src[0];
        if (x->ne[0] != r->ne[0] * r->ne[1] || x->ne[1] != r->ne[2] || x->ne[2] * x->ne[3] != r->ne[3]) return nullptr;
        return x;
    }

    struct AttnMatch {
        const ggml_tensor *x, *wq, *wk, *wv, *bq, *bk, *bv;
        int heads;
    };
    // self-attention core of transformer_layer::forward (main.cpp:1022-1093); t = (C, L, B)
    bool match_attention(const ggml_tensor * t, AttnMatch & m) {
        if (!is_op(t, GGML_OP_RESHAPE) || !is_op(t->src[0], GGML_OP_CONT)) return false;
        const ggml_tensor * pa = t->src[0]->src[0];
        if (!permute_is(pa, 0, 2, 1, 3)) return false;
        const ggml_tensor * att = pa->src[0];
        if (!is_op(att, GGML_OP_MUL_MAT)) return false;
        const ggml_tensor * vt = att->src[0], * sm = att->src[1];
        if (!is_op(vt, GGML_OP_CONT) || !permute_is(vt->src[0], 1, 0, 2, 3) || !is_op(sm, GGML_OP_SOFT_MAX)) return false;
        const ggml_tensor * sc = sm->src[0];
        float scale;
        if (!is_op(sc, GGML_OP_DIV) || !scalar_const(sc->src[1], &scale) || !is_op(sc->src[0], GGML_OP_MUL_MAT)) return false;
        int hq, hk, hv;
        const ggml_tensor * dk = match_heads(sc->src[0]->src[0], &hk);
        const ggml_tensor * dq = match_heads(sc->src[0]->src[1], &hq);
        const ggml_tensor * dv = match_heads(vt->src[0]->src[0], &hv);
        if (!dk || !dq || !dv || hq != hk || hq != hv) return false;
        const ggml_tensor *xq, *xk, *xv;
        if (!match_dense(dq, &xq, &m.wq, &m.bq) || !match_dense(dk, &xk, &m.wk, &m.bk) || !match_dense(dv, &xv, &m.wv, &m.bv)) return false;
        if (xq != xk || xq != xv) return false;
        const int64_t C = t->ne[0];
        if (m.wq->ne[0] != C || m.wk->ne[0] != C || m.wv->ne[0] != C || C % hq) return false;
        if (fabsf(scale - sqrtf((float)(C / hq))) > 1e-4f * scale) return false;  // main.cpp:999,1076
        m.x     = xq;
        m.heads = hq;
        return true;
    }

    // ---- lowering ------------------------------------------------------------------------------------------
    FVal * lower_conv(const ggml_tensor * c, const BNParams * bn, int act) {
        REQUIRE(is_op(c, GGML_OP_CONV_2D) || is_op(c, GGML_OP_CONV_DEPTHWISE_2D));
        const ggml_tensor * wc = c->src[0];
        REQUIRE(is_op(wc, GGML_OP_CONT) && permute_is(wc->src[0], 3, 2, 0, 1));
        const ggml_tensor * kernel = wc->src[0]->src[0];  // leaf, ne = (OC, IC, KW, KH)
        REQUIRE(is_weight_leaf(kernel) && kernel->type == GGML_TYPE_F16);
        const int64_t OC = kernel->ne[0], IC = kernel->ne[1], KW = kernel->ne[2], KH = kernel->ne[3];
        REQUIRE(ne_is(wc, KW, KH, IC, OC));
        const int s0 = c->op_params[0], s1 = c->op_params[1], p0 = c->op_params[2], p1 = c->op_params[3];
        REQUIRE(s0 == s1 && p0 == p1 && c->op_params[4] == 1 && c->op_params[5] == 1 && KW == KH && p0 == (KW - 1) / 2);
        const ggml_tensor * xin = c->src[1];
        FVal * out = new_val((int)c->ne[3], (int)c->ne[1], (int)c->ne[0], (int)c->ne[2]);
        FNode * n;
        if (is_op(c, GGML_OP_CONV_DEPTHWISE_2D)) {
            REQUIRE(KW == 3 && IC == 1 && (s0 == 1 || s0 == 2) && OC % 8 == 0);
            REQUIRE(xin->ne[0] % s0 == 0 && xin->ne[1] % s0 == 0);
            n = new_node(FK_DW, out, "dwconv3x3");
            n->in.push_back(lower(xin));
        } else if (KW == 1) {
            REQUIRE(s0 == 1 && OC % 8 == 0 && IC % 8 == 0);
            n = new_node(FK_CONV1, out, "conv1x1");
            n->in.push_back(lower(xin));
        } else if (KW == 3 && s0 == 1) {
            REQUIRE(OC % 8 == 0 && IC % 8 == 0);
            n = new_node(FK_CONV3, out, "conv3x3");
            if (is_op(xin, GGML_OP_CONCAT)) {  // main.cpp:1219: fused as a second TMA source
                n->in.push_back(lower(xin->src[0]));
                n->in.push_back(lower(xin->src[1]));
                REQUIRE(n->in[0]->C % 8 == 0 && n->in[1]->C % 8 == 0);
            } else {
                n->in.push_back(lower(xin));
            }
        } else if (KW == 3 && s0 == 2 && IC == 3) {
            REQUIRE(OC % 8 == 0 && OC <= 32 && xin->ne[0] % 2 == 0 && xin->ne[1] % 2 == 0);
            n = new_node(FK_STEM, out, "stem");
            FVal * iv = lower(xin);
            REQUIRE(iv->is_input);
            n->in.push_back(iv);
        } else {
            REQUIRE(!"unsupported convolution shape");
            return nullptr;
        }
        n->w      = kernel;
        n->stride = s0;
        n->act    = act;
        if (bn) { n->has_bn = true; n->bn = *bn; }
        return out;
    }

    FVal * lower(const ggml_tensor * t) {
        auto it = memo.find(t);
        if (it != memo.end()) return it->second;
        FVal * v = lower_uncached(t);
        memo[t]  = v;
        return v;
    }

    FVal * lower_uncached(const ggml_tensor * t) {
        BNParams bn;
        const ggml_tensor *x, *w, *b, *g;
        float eps;
        switch (t->op) {
            case GGML_OP_NONE: {  // CHW image leaf of an unmodified main.cpp (main.cpp:612)
                REQUIRE(t->type == GGML_TYPE_F32 && t->ne[2] == 3 && t->data != nullptr && ggml_is_contiguous(t));
                FVal * v    = new_val((int)t->ne[3], (int)t->ne[1], (int)t->ne[0], 3);
                v->is_input = true;
                FNode * n   = new_node(FK_INPUT, v, "input_chw");
                n->leaf     = t;
                n->chw      = true;
                return v;
            }
            case GGML_OP_POOL_MEAN_HW: {
                FVal * in = lower(t->src[0]);
                FVal * v  = new_val(in->N, 1, 1, in->C);
                FNode * n = new_node(FK_POOL, v, "pool");
                n->in.push_back(in);
                return v;
            }
            case GGML_OP_SILU: {
                const ggml_tensor * u = t->src[0];
                if (const ggml_tensor * c = match_bn(u, bn)) return lower_conv(c, &bn, 1);
                if (match_dense(u, &x, &w, &b)) return lower_linear(u, x, w, b, 1);
                REQUIRE(!"silu over an unrecognised producer");
            }
            case GGML_OP_ADD: {
                if (const ggml_tensor * c = match_bn(t, bn)) return lower_conv(c, &bn, 0);
                if (match_dense(t, &x, &w, &b)) return lower_linear(t, x, w, b, 0);
                REQUIRE(same_ne(t->src[0], t->src[1]) && same_ne(t, t->src[0]));  // residual add (main.cpp:867,1111,1165)
                FVal * a = lower(t->src[0]);
                FVal * c = lower(t->src[1]);
                REQUIRE(a->rows() == c->rows() && a->C == c->C);
                FVal * v  = new_val(a->N, a->H, a->W, a->C);
                FNode * n = new_node(FK_ADD, v, "add");
                n->in     = {a, c};
                return v;
            }
            case GGML_OP_CONV_2D: return lower_conv(t, nullptr, 0);  // conv_1x1 of the ViT block: no BN, no act (main.cpp:1183)
            case GGML_OP_CONT: {
                if (match_ln(t, &x, &g, &b, &eps)) {
                    FVal * in = lower(x);
                    REQUIRE(in->C == t->ne[0] && in->rows() * in->C == ggml_nelements(t));
                    FVal * v  = new_val(in->N, in->H, in->W, in->C);
                    FNode * n = new_node(FK_LN, v, "layernorm");
                    n->in.push_back(in);
                    n->g = g; n->b = b; n->eps = eps;
                    return v;
                }
                if (const ggml_tensor * src = match_unfold(t)) return lower(src);  // batch-1 form without the 3-d reshape
                // HWC image input of this build: CONT(PERMUTE(leaf(3,W,H,N); 2,0,1,3))
                const ggml_tensor * p = t->src[0];
                REQUIRE(permute_is(p, 2, 0, 1, 3));
                const ggml_tensor * leaf = p->src[0];
                REQUIRE(leaf->op == GGML_OP_NONE && leaf->type == GGML_TYPE_F32 && leaf->ne[0] == 3 && ggml_is_contiguous(leaf));
                FVal * v    = new_val((int)leaf->ne[3], (int)leaf->ne[2], (int)leaf->ne[1], 3);
                v->is_input = true;
                FNode * n   = new_node(FK_INPUT, v, "input_hwc");
                n->leaf     = leaf;
                n->chw      = false;
                return v;
            }
            case GGML_OP_RESHAPE: {
                AttnMatch am;
                if (match_attention(t, am)) {
                    FVal * in = lower(am.x);
                    const int C = (int)t->ne[0];
                    REQUIRE(in->C == C && in->H % 2 == 0 && in->W % 2 == 0 && (C / am.heads) <= 64);
                    REQUIRE(t->ne[1] == (in->H / 2) * (in->W / 2) && t->ne[2] == 4 * in->N);
                    FVal * qkv = new_val(in->N, in->H, in->W, 3 * am.heads * attention_padded_head_dim(C / am.heads));
                    FNode * nq = new_node(FK_QKV, qkv, "qkv");
                    nq->in.push_back(in);
                    nq->wq = am.wq; nq->wk = am.wk; nq->wv = am.wv; nq->bq = am.bq; nq->bk = am.bk; nq->bv = am.bv;
                    nq->heads = am.heads;
                    FVal * ctx = new_val(in->N, in->H, in->W, C);
                    FNode * na = new_node(FK_ATTN, ctx, "attention");
                    na->in.push_back(qkv);
                    na->heads = am.heads;
                    return ctx;
                }
                if (const ggml_tensor * src = match_unfold(t)) return lower(src);
                if (const ggml_tensor * src = match_fold(t)) {
                    FVal * in = lower(src);
                    REQUIRE(in->W == t->ne[0] && in->H == t->ne[1] && in->C == t->ne[2] && in->N == t->ne[3]);
                    return in;
                }
                REQUIRE(!"unrecognised reshape");
            }
            default: REQUIRE(!"unsupported op in fused planner");
        }
        return nullptr;
    }

    FVal * lower_linear(const ggml_tensor * t, const ggml_tensor * x, const ggml_tensor * w, const ggml_tensor * b, int act) {
        FVal * in = lower(x);
        REQUIRE(in->C == w->ne[1] && w->ne[0] % 8 == 0 && w->ne[1] % 8 == 0);
        REQUIRE(in->rows() * w->ne[0] == ggml_nelements(t));
        FVal * v  = new_val(in->N, in->H, in->W, (int)w->ne[0]);
        FNode * n = new_node(FK_LINEAR, v, "linear");
        n->in.push_back(in);
        n->w = w; n->bias = b; n->act = act;
        return v;
    }

    // ---- constant folding (host) ---------------------------------------------------------------------------
    void fold_bn(FNode * n, int OC, const std::vector<float> * pre = nullptr) {
        std::vector<float> scale(OC, 1.f), shift(OC, 0.f);
        if (!n->has_bn && pre) {  // folded LayerNorm without BN: shift = sum_k beta_k W[n][k]
            n->c_shift = pool.add(pre->data(), OC * 4);
            return;
        }
        if (n->has_bn) {
            const float *mean = (const float *)n->bn.mean->data, *var = (const float *)n->bn.var->data;
            const float *gamma = (const float *)n->bn.gamma->data, *beta = (const float *)n->bn.beta->data;
            for (int i = 0; i < OC; i++) {
                // ((x - mean) / sqrt(var + eps)) * gamma + beta  ==  x * scale + shift   (main.cpp:809-846)
                const double sd = sqrt((double)(var[i] + n->bn.eps));
                const double sc = (double)gamma[i] / sd;
                scale[i]        = (float)sc;
                shift[i]        = (float)((double)beta[i] - (double)mean[i] * sc);
                if (pre) shift[i] = (float)((double)shift[i] + sc * (double)(*pre)[i]);  // BN applied after the folded LayerNorm's beta term
            }
            n->c_scale = pool.add(scale.data(), OC * 4);
            n->c_shift = pool.add(shift.data(), OC * 4);
            plan->n_folded += 14;  // the BN parameter chain of one conv_layer::forward
        }
    }

    void fold_constants() {
        for (auto & up : nodes) {
            FNode * n = up.get();
            if (n->dead && !n->fused_ir) continue;
            switch (n->kind) {
                case FK_STEM: case FK_CONV1: case FK_CONV3: case FK_DW: case FK_IR: {
                    const ggml_tensor * k = n->w;  // f16, ne=(OC,IC,KW,KH): memory index ((kh*KW+kw)*IC+ic)*OC+oc
                    const int OC = (int)k->ne[0], IC = (int)k->ne[1], KW = (int)k->ne[2], KH = (int)k->ne[3];
                    const uint16_t * src = (const uint16_t *)k->data;
                    std::vector<uint16_t> wt((size_t)OC * IC * KW * KH);
                    if (n->kind == FK_DW) {
                        memcpy(wt.data(), src, wt.size() * 2);  // already [kh][kw][c]
                    } else {  // [oc][kh][kw][ic]: K-major rows for the GEMM B operand (replaces cont(permute(kernel)), main.cpp:790-805)
                        for (int kh = 0; kh < KH; kh++)
                            for (int kw = 0; kw < KW; kw++)
                                for (int ic = 0; ic < IC; ic++)
                                    for (int oc = 0; oc < OC; oc++)
                                        wt[(((size_t)oc * KH + kh) * KW + kw) * IC + ic] = src[(((size_t)kh * KW + kw) * IC + ic) * OC + oc];
                    }
                    std::vector<float> pre;
                    if (n->ln_g) {  // 1x1 conv over LN(x): W' = f16(W * gamma), c1 = row sums of W', pre = sum_k beta_k W
                        const float *gm = (const float *)n->ln_g->data, *bt = (const float *)n->ln_b->data;
                        std::vector<float> c1(OC, 0.f);
                        pre.assign(OC, 0.f);
                        for (int oc = 0; oc < OC; oc++) {
                            double a1 = 0.0, a2 = 0.0;
                            for (int ic = 0; ic < IC; ic++) {
                                const float    w  = ggml_fp16_to_fp32(wt[(size_t)oc * IC + ic]);
                                const uint16_t wg = ggml_fp32_to_fp16(w * gm[ic]);
                                wt[(size_t)oc * IC + ic] = wg;
                                a1 += (double)ggml_fp16_to_fp32(wg);
                                a2 += (double)bt[ic] * (double)w;
                            }
                            c1[oc]  = (float)a1;
                            pre[oc] = (float)a2;
                        }
                        n->c_c1 = pool.add(c1.data(), OC * 4);
                    }
                    n->c_w = pool.add(wt.data(), wt.size() * 2);
                    if (n->kind == FK_CONV3 && KW == 3) {  // the same weights pre-tiled for the halo-mode kernel (gemm_tcgen05.cu, conv == 2)
                        const int c0 = n->in[0]->C, c1 = n->in.size() > 1 ? n->in[1]->C : 0;
                        if (c0 + c1 == IC) {
                            std::vector<uint8_t> halo;
                            conv3x3_pack_halo(wt.data(), OC, c0, c1, halo);
                            n->c_w2 = pool.add(halo.data(), halo.size());
                        }
                    }
                    plan->n_folded += 2;
                    fold_bn(n, OC, n->ln_g ? &pre : nullptr);
                } break;
                case FK_LINEAR: {
                    const int OUT = (int)n->w->ne[0], IN = (int)n->w->ne[1];
                    const float * src = (const float *)n->w->data;  // file (in,out): index in*OUT + out
                    std::vector<uint16_t> wt((size_t)OUT * IN);
                    std::vector<float> shift((const float *)n->bias->data, (const float *)n->bias->data + OUT);
                    if (n->ln_g) {
                        const float *gm = (const float *)n->ln_g->data, *bt = (const float *)n->ln_b->data;
                        std::vector<float> c1(OUT, 0.f);
                        for (int o = 0; o < OUT; o++) {
                            double a1 = 0.0, a2 = 0.0;
                            for (int i = 0; i < IN; i++) {
                                const float    w  = src[(size_t)i * OUT + o];
                                const uint16_t wg = ggml_fp32_to_fp16(w * gm[i]);
                                wt[(size_t)o * IN + i] = wg;
                                a1 += (double)ggml_fp16_to_fp32(wg);
                                a2 += (double)bt[i] * (double)w;
                            }
                            c1[o] = (float)a1;
                            shift[o] += (float)a2;
                        }
                        n->c_c1 = pool.add(c1.data(), OUT * 4);
                    } else {
                        for (int o = 0; o < OUT; o++)
                            for (int i = 0; i < IN; i++) wt[(size_t)o * IN + i] = ggml_fp32_to_fp16(src[(size_t)i * OUT + o]);
                    }
                    n->c_w     = pool.add(wt.data(), wt.size() * 2);
                    n->c_shift = pool.add(shift.data(), OUT * 4);
                    plan->n_folded += 4;
                } break;
                case FK_QKV: {
                    // rows ordered [q|k|v][head][DP]: each head padded to DP = 16-multiple with zero rows / zero bias, so the
                    // GEMM itself writes the 16-byte aligned, zero-padded per-head layout the attention kernel consumes
                    const int C = (int)n->wq->ne[0], IN = (int)n->wq->ne[1], heads = n->heads;
                    const int d = C / heads, dp = attention_padded_head_dim(d);
                    std::vector<uint16_t> wt((size_t)3 * heads * dp * IN, 0);
                    std::vector<float>    bias((size_t)3 * heads * dp, 0.f), c1q((size_t)3 * heads * dp, 0.f);
                    const ggml_tensor * ws[3] = {n->wq, n->wk, n->wv};
                    const ggml_tensor * bs[3] = {n->bq, n->bk, n->bv};
                    for (int s = 0; s < 3; s++) {
                        const float * src = (const float *)ws[s]->data;
                        for (int o = 0; o < C; o++) {
                            const size_t row = ((size_t)s * heads + o / d) * dp + o % d;
                            bias[row] = ((const float *)bs[s]->data)[o];
                            if (n->ln_g) {
                                const float *gm = (const float *)n->ln_g->data, *bt = (const float *)n->ln_b->data;
                                double a1 = 0.0, a2 = 0.0;
                                for (int i = 0; i < IN; i++) {
                                    const float    w  = src[(size_t)i * C + o];
                                    const uint16_t wg = ggml_fp32_to_fp16(w * gm[i]);
                                    wt[row * IN + i]  = wg;
                                    a1 += (double)ggml_fp16_to_fp32(wg);
                                    a2 += (double)bt[i] * (double)w;
                                }
                                c1q[row] = (float)a1;
                                bias[row] += (float)a2;
                            } else {
                                for (int i = 0; i < IN; i++) wt[row * IN + i] = ggml_fp32_to_fp16(src[(size_t)i * C + o]);
                            }
                        }
                    }
                    if (n->ln_g) n->c_c1 = pool.add(c1q.data(), c1q.size() * 4);
                    n->c_w     = pool.add(wt.data(), wt.size() * 2);
                    n->c_shift = pool.add(bias.data(), bias.size() * 4);
                    plan->n_folded += 12;
                } break;
                case FK_VIT: {
                    // K8: every layer's weights tiled into the streaming order of the fused stage kernel (vit_stage.cu)
                    std::vector<VitLayerHost> hl;
                    for (const VitLayerNodes & L : n->vit) {
                        VitLayerHost h = {};
                        if (L.qkv) {  // (MLP-only nodes carry just ln2 / up / down)
                            h.ln1_g = (const float *)L.ln1->g->data;  h.ln1_b = (const float *)L.ln1->b->data;
                            h.wq = (const float *)L.qkv->wq->data;    h.bq = (const float *)L.qkv->bq->data;
                            h.wk = (const float *)L.qkv->wk->data;    h.bk = (const float *)L.qkv->bk->data;
                            h.wv = (const float *)L.qkv->wv->data;    h.bv = (const float *)L.qkv->bv->data;
                            h.wo = (const float *)L.proj->w->data;    h.bo = (const float *)L.proj->bias->data;
                        }
                        h.ln2_g = (const float *)L.ln2->g->data;  h.ln2_b = (const float *)L.ln2->b->data;
                        h.w1 = (const float *)L.up->w->data;      h.b1 = (const float *)L.up->bias->data;
                        h.w2 = (const float *)L.down->w->data;    h.b2 = (const float *)L.down->bias->data;
                        hl.push_back(h);
                    }
                    std::vector<uint8_t> blob;
                    std::vector<float>   vec;
                    vit_stage_pack(hl.data(), (int)hl.size(), n->out->C, n->heads, n->vit_F, blob, vec);
                    n->c_w     = pool.add(blob.data(), blob.size());
                    n->c_shift = pool.add(vec.data(), vec.size() * 4);
                    plan->n_folded += 16 * (int)hl.size();
                } break;
                default: break;
            }
        }
    }

    // ---- scheduling ------------------------------------------------------------------------------------------
    std::vector<FNode *> order;
    void topo(FNode * n, std::set<FNode *> & seen) {
        if (!n || seen.count(n)) return;
        seen.insert(n);
        for (FVal * v : n->in) topo(v->prod, seen);
        if (n->res) topo(n->res->prod, seen);
        order.push_back(n);
    }
};

}  // namespace

static bool build_fast_plan_impl(Plan * plan, ggml_cgraph * gf) {
    Planner P;
    P.plan = plan;
    P.gf   = gf;
    t_compute_ctx = plan->ctx;
    // graph outputs: every node flagged as output (features, pooled, debug taps)
    std::vector<ggml_tensor *> outs;
    for (int i = 0; i < gf->n_nodes; i++)
        if (gf->nodes[i]->flags & GGML_TENSOR_FLAG_OUTPUT) outs.push_back(gf->nodes[i]);
    if (outs.empty()) return false;
    std::vector<FVal *> out_vals;
    // classifier head: ADD(MUL_MAT(CONT(TRANSPOSE(kernel)), RESHAPE(pooled)), bias) with `pooled` itself an output
    struct Head { const ggml_tensor *pooled, *w, *b; };
    std::map<const ggml_tensor *, Head> heads;
    try {
        for (ggml_tensor * t : outs) {
            const ggml_tensor *hx, *hw, *hb;
            if (t->op == GGML_OP_ADD && P.match_dense(t, &hx, &hw, &hb) && hx->op == GGML_OP_RESHAPE && hx->src[0]->op == GGML_OP_POOL_MEAN_HW) {
                const ggml_tensor * pooled = hx->src[0];
                if (!(pooled->flags & GGML_TENSOR_FLAG_OUTPUT) || hx->ne[0] != pooled->ne[2] || hx->ne[1] != pooled->ne[3]) throw Fail();
                heads[t] = Head{pooled, hw, hb};
                out_vals.push_back(P.lower(pooled));
                continue;
            }
            FVal * v = P.lower(t);
            if (t->op == GGML_OP_POOL_MEAN_HW) {
                if (!(t->ne[2] == v->C && t->ne[3] == v->N)) throw Fail();
            } else if (!(t->ne[0] == v->W && t->ne[1] == v->H && t->ne[2] == v->C && t->ne[3] == v->N && t->type == GGML_TYPE_F32)) {
                throw Fail();  // only (W,H,C,N) activations can be graph outputs of the fused plan
            }
            out_vals.push_back(v);
        }
    } catch (const Fail &) {
        plan->n_folded = 0;
        return false;
    }

    // ---- users, residual fusion ----
    auto rebuild_users = [&]() {
        for (auto & v : P.vals) v->users.clear();
        for (auto & n : P.nodes) {
            if (n->dead) continue;
            for (FVal * v : n->in) v->users.push_back(n.get());
            if (n->res) n->res->users.push_back(n.get());
        }
    };
    rebuild_users();
    std::set<FVal *> out_set(out_vals.begin(), out_vals.end());
    for (auto & up : P.nodes) {
        FNode * n = up.get();
        if (n->kind != FK_ADD || n->dead) continue;
        for (int side = 0; side < 2; side++) {
            FVal *  a  = n->in[side], * other = n->in[1 - side];
            FNode * pr = a->prod;
            if (!pr || pr->res || pr->dead) continue;
            if (!(pr->kind == FK_CONV1 || pr->kind == FK_LINEAR)) continue;
            if (a->users.size() != 1 || out_set.count(a)) continue;
            // residual add happens after BN/bias (+ activation): exactly the reference order (main.cpp:864-868,1108-1111)
            pr->res      = other;
            pr->out      = n->out;
            n->out->prod = pr;
            n->dead      = true;
            break;
        }
    }
    rebuild_users();

    // ---- K8: the transformer layers of one MobileViT block (main.cpp:1196-1204) as ONE kernel: x -> LN -> qkv -> attention -> projection
    // (+x) -> LN -> up (+SiLU) -> down (+res), repeated; tiles of whole sequences never meet, so the residual stream stays on chip for the
    // whole stage (vit_stage.cu).  GGML_B200_VIT_FUSE=0 keeps the five launches per layer; =1 fuses every stage the kernel covers;
    // default: stages of at most kVitAutoTiles 128-token tiles, where the separate launches are latency-bound. ----
    {
        const char * e    = getenv("GGML_B200_VIT_FUSE");
        const int    mode = e ? atoi(e) : 2;
        auto sole_user = [&](FVal * v, FKind k) -> FNode * {
            if (!v || out_set.count(v) || v->users.size() != 1 || v->users[0]->kind != k || v->users[0]->dead) return nullptr;
            return v->users[0];
        };
        // one layer starting at the residual stream value x; returns the layer's output value or nullptr
        auto match_layer = [&](FVal * x, VitLayerNodes & L, int & heads, int & F) -> FVal * {
            if (!x || out_set.count(x) || x->users.size() != 2) return nullptr;
            FNode *ln1 = nullptr, *proj = nullptr;
            for (FNode * u : x->users) {
                if (u->kind == FK_LN && u->in[0] == x && !u->dead) ln1 = u;
                else if (u->kind == FK_LINEAR && u->res == x && !u->dead) proj = u;
            }
            if (!ln1 || !proj || proj->act || proj->ln_g) return nullptr;
            FNode * qkv = sole_user(ln1->out, FK_QKV);
            if (!qkv || qkv->ln_g) return nullptr;
            FNode * attn = sole_user(qkv->out, FK_ATTN);
            if (!attn || sole_user(attn->out, FK_LINEAR) != proj || proj->in[0] != attn->out) return nullptr;
            FVal * x1 = proj->out;
            if (out_set.count(x1) || x1->users.size() != 2) return nullptr;
            FNode *ln2 = nullptr, *down = nullptr;
            for (FNode * u : x1->users) {
                if (u->kind == FK_LN && u->in[0] == x1 && !u->dead) ln2 = u;
                else if (u->kind == FK_LINEAR && u->res == x1 && !u->dead) down = u;
            }
            if (!ln2 || !down || down->act || down->ln_g) return nullptr;
            FNode * up = sole_user(ln2->out, FK_LINEAR);
            if (!up || !up->act || up->res || up->ln_g || sole_user(up->out, FK_LINEAR) != down || down->in[0] != up->out) return nullptr;
            if (ln1->eps != ln2->eps || x1->C != x->C || down->out->C != x->C) return nullptr;
            heads = qkv->heads;
            F     = up->out->C;
            L     = VitLayerNodes{ln1, qkv, proj, ln2, up, down};
            return down->out;
        };
        constexpr int kVitAutoTiles = 4 * 148;
        for (size_t vi = 0; mode > 0 && vi < P.vals.size(); vi++) {
            FVal * x0 = P.vals[vi].get();
            if (x0->prod && x0->prod->kind == FK_VIT) continue;
            std::vector<VitLayerNodes> layers;
            int    heads = 0, F = 0;
            FVal * x = x0;
            for (;;) {
                VitLayerNodes L;
                int           h2 = 0, f2 = 0;
                FVal *        nx = match_layer(x, L, h2, f2);
                if (!nx || (!layers.empty() && (h2 != heads || f2 != F || L.ln1->eps != layers[0].ln1->eps))) break;
                heads = h2; F = f2;
                layers.push_back(L);
                x = nx;
            }
            if (layers.empty() || !vit_stage_supported(x0->N, x0->H, x0->W, x0->C, heads, F)) continue;
            const int64_t tiles = (x0->rows() + 127) / 128;
            if (mode == 2 && tiles > kVitAutoTiles) continue;
            FNode * v = layers.back().down;  // becomes the stage node: its output value is the stage's output
            for (const VitLayerNodes & L : layers) {
                FNode * attn = L.qkv->out->users[0];
                for (FNode * d : {L.ln1, L.qkv, attn, L.proj, L.ln2, L.up, L.down})
                    if (d != v) d->dead = true;
            }
            v->kind  = FK_VIT;
            v->name  = "vit_stage";
            v->in    = {x0};
            v->res   = nullptr;
            v->heads = heads;
            v->vit_F = F;
            v->eps   = layers[0].ln1->eps;
            v->vit   = layers;
            plan->n_folded += 4 * (int)layers.size();
            rebuild_users();
        }
        // MLP-only fusion for the layers left over (sequences longer than a tile: L = 256 / 1024): x1 -> LN -> up (+SiLU) -> down (+x1) as
        // one launch of the same kernel (heads = 0); the hidden activation never reaches HBM.  Opt-in (GGML_B200_MLP_FUSE=1): parity-green,
        // but measured equal to the two GEMMs it replaces (batch 256: 237 us vs 91 + 109 us, the producing projection gets 15 us faster
        // because it no longer writes the f16 copy; whole step 5.09 vs 5.03 ms) -- one tile in flight per SM is a 16 us serial chain.
        const char * em = getenv("GGML_B200_MLP_FUSE");
        for (size_t vi = 0; mode > 0 && em && atoi(em) > 0 && vi < P.vals.size(); vi++) {
            FVal * x1 = P.vals[vi].get();
            if (out_set.count(x1) || x1->users.size() != 2) continue;
            FNode *ln2 = nullptr, *down = nullptr;
            for (FNode * u : x1->users) {
                if (u->kind == FK_LN && u->in[0] == x1 && !u->dead) ln2 = u;
                else if (u->kind == FK_LINEAR && u->res == x1 && !u->dead) down = u;
            }
            if (!ln2 || !down || down->act || down->ln_g) continue;
            FNode * up = sole_user(ln2->out, FK_LINEAR);
            if (!up || !up->act || up->res || up->ln_g || sole_user(up->out, FK_LINEAR) != down || down->in[0] != up->out) continue;
            if (down->out->C != x1->C || !vit_stage_supported(x1->N, x1->H, x1->W, x1->C, 0, up->out->C)) continue;
            ln2->dead = up->dead = true;
            down->kind  = FK_VIT;
            down->name  = "vit_mlp";
            down->in    = {x1};
            down->res   = nullptr;
            down->heads = 0;
            down->vit_F = up->out->C;
            down->eps   = ln2->eps;
            down->vit   = {VitLayerNodes{nullptr, nullptr, nullptr, ln2, up, down}};
            plan->n_folded += 2;
            rebuild_users();
        }
    }

    // ---- LayerNorm folding: LN(x) feeding only GEMMs disappears.  The GEMM that produces x (tile spans the row) also writes
    // per-row (sum, sum of squares) and an f16 copy of x; each consumer multiplies raw x with gamma-scaled weights and applies
    // r * (acc - mu * c1[n]) in its epilogue (beta folded into the shift).  Saves the LN kernels' 4+2 bytes per element. ----
    if (!getenv("GGML_B200_NO_LN_FOLD")) {
        for (auto & up : P.nodes) {
            FNode * ln = up.get();
            if (ln->kind != FK_LN || ln->dead) continue;
            FVal *x = ln->in[0], *y = ln->out;
            FNode * pr = x->prod;
            if (!pr || pr->dead || !(pr->kind == FK_CONV1 || pr->kind == FK_LINEAR || pr->kind == FK_VIT) || x->C > 256 || x->C % 8) continue;
            if (out_set.count(y) || y->users.empty()) continue;
            bool ok = true;
            for (FNode * u : y->users)
                if (!(u->kind == FK_QKV || u->kind == FK_LINEAR || u->kind == FK_CONV1) || u->in.size() != 1 || u->in[0] != y || u->res == y || u->ln_g) ok = false;
            if (!ok) continue;
            for (FNode * u : y->users) {
                u->in[0]  = x;
                u->ln_g   = ln->g;
                u->ln_b   = ln->b;
                u->ln_eps = ln->eps;
            }
            x->need_stats = true;
            ln->dead      = true;
            plan->n_folded += 6;
        }
        rebuild_users();
    }

    // ---- K4: expand 1x1 (+BN+SiLU) -> depthwise 3x3 (+BN+SiLU) -> reduce 1x1 (+BN) [+ residual] as ONE kernel: the expanded
    // activation and the depthwise output never reach HBM (inverted_residual_layer::forward, main.cpp:854-870) ----
    // Opt-in (GGML_B200_IR_FUSE=1): parity-green, but on B200 the fused block is ALU-bound -- depthwise + two SiLU epilogues on one
    // SM at 18 warps -- and measured 5-20 % SLOWER than the three HBM-bound kernels it replaces (profiles/README.md, round 2).
    if (getenv("GGML_B200_IR_FUSE") != nullptr && atoi(getenv("GGML_B200_IR_FUSE")) > 0) {
        for (auto & up : P.nodes) {
            FNode * c = up.get();
            if (c->dead || c->kind != FK_CONV1 || c->ln_g || c->act || !c->has_bn || c->out->need_stats) continue;
            FVal *  dv = c->in[0];
            FNode * b  = dv->prod;
            if (!b || b->dead || b->kind != FK_DW || !b->has_bn || !b->act || dv->users.size() != 1 || out_set.count(dv)) continue;
            FVal *  ev = b->in[0];
            FNode * a  = ev->prod;
            if (!a || a->dead || a->kind != FK_CONV1 || a->res || a->ln_g || !a->act || !a->has_bn || ev->users.size() != 1 || out_set.count(ev) || ev->need_stats) continue;
            FVal * xv = a->in[0];
            IrLaunch probe;
            if (!ir_fused_prepare(probe, nullptr, xv->N, xv->H, xv->W, xv->C, ev->C, c->out->C, b->stride, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                  nullptr, nullptr, nullptr, nullptr, nullptr))
                continue;
            c->kind      = FK_IR;
            c->name      = "inverted_residual";
            c->ir_expand = a;
            c->ir_dw     = b;
            c->in[0]     = xv;
            a->dead = b->dead = true;
            a->fused_ir = b->fused_ir = true;
            plan->n_folded += 2;
        }
        rebuild_users();
    }

    // ---- topological order ----
    std::set<FNode *> seen;
    for (FVal * v : out_vals) P.topo(v->prod, seen);
    for (size_t i = 0; i < P.order.size(); i++) P.order[i]->order = (int)i;
    const int n_steps = (int)P.order.size();

    // ---- representation needs ----
    for (FNode * n : P.order) {
        switch (n->kind) {
            case FK_CONV1: case FK_CONV3: case FK_DW: case FK_LINEAR: case FK_QKV: case FK_ATTN: case FK_IR:
                for (FVal * v : n->in) v->need16 = true;
                break;
            case FK_LN: case FK_ADD: case FK_VIT:
                for (FVal * v : n->in) v->need32 = true;
                break;
            case FK_POOL: n->in[0]->need32 = true; break;
            default: break;
        }
        if (n->res) n->res->need32 = true;
    }
    for (FVal * v : out_vals)
        if (!v->need16 && !v->need32) v->need32 = true;
    for (FNode * n : P.order) {
        FVal * o = n->out;
        if (n->kind == FK_INPUT) continue;
        if ((n->kind == FK_DW || n->kind == FK_ATTN || n->kind == FK_QKV) && o->need32) {
            if (runtime().verbose) fprintf(stderr, "libggml_b200: fused planner: f32 copy requested from an f16-only kernel (%s)\n", n->name.c_str());
            plan->n_folded = 0;
            return false;
        }
        if (n->kind == FK_POOL) { o->need32 = true; o->need16 = false; }
    }
    // shapes the tensor-core kernels cannot tile -> exact plan
    for (FNode * n : P.order) {
        if (n->kind == FK_CONV3) {
            const FVal * v = n->in[0];
            const bool ok = v->W <= 128;  // a tile is a whole number of image rows (conv3x3_prepare)
            if (!ok) { plan->n_folded = 0; return false; }
        }
    }

    // ---- liveness + arena ----
    for (FNode * n : P.order) {
        n->out->def = n->order;
        for (FVal * v : n->in) v->last = std::max(v->last, n->order);
        if (n->res) n->res->last = std::max(n->res->last, n->order);
    }
    for (FVal * v : out_vals) {
        v->last = n_steps;  // converted to the ggml layout after the last kernel
        if (v->prod && v->prod->kind == FK_POOL) v->prod->in[0]->last = n_steps;
    }
    ArenaPlanner ap;
    std::vector<std::vector<FVal *>> dies(n_steps + 1);
    for (FNode * n : P.order) {
        FVal * o = n->out;
        if (n->kind != FK_INPUT) {
            if (o->need16) { o->off16 = ap.alloc(o->rows() * o->C * 2); plan->naive_bytes += ArenaPlanner::align_up(o->rows() * o->C * 2); }
            if (o->need32) { o->off32 = ap.alloc(o->rows() * o->C * 4); plan->naive_bytes += ArenaPlanner::align_up(o->rows() * o->C * 4); }
            if (o->need_stats) o->off_stats = ap.alloc(o->rows() * 8);
            int last = std::max(o->last, n->order);
            if (last < n_steps) dies[last].push_back(o);
        }
        for (FVal * d : dies[n->order]) {
            if (d->off16 >= 0) ap.release(d->off16, d->rows() * d->C * 2);
            if (d->off32 >= 0) ap.release(d->off32, d->rows() * d->C * 4);
            if (d->off_stats >= 0) ap.release(d->off_stats, d->rows() * 8);
        }
    }
    // graph outputs in ggml layout
    std::vector<int64_t> out_off(outs.size());
    for (size_t i = 0; i < outs.size(); i++) out_off[i] = ap.alloc((int64_t)ggml_nelements(outs[i]) * 4);
    std::map<const ggml_tensor *, size_t> head_w, head_b;
    for (auto & kv : heads) {
        head_w[kv.first] = P.pool.add(kv.second.w->data, ggml_nbytes(kv.second.w));
        head_b[kv.first] = P.pool.add(kv.second.b->data, (size_t)kv.first->ne[0] * 4);
    }
    plan->arena_bytes = ap.extent;
    B200_CHECK(cudaMalloc((void **)&plan->arena, plan->arena_bytes > 0 ? plan->arena_bytes : 256));
    plan->owned_device.push_back(plan->arena);
    for (auto & v : P.vals) {
        if (v->off16 >= 0) v->p16 = (__half *)(plan->arena + v->off16);
        if (v->off32 >= 0) v->p32 = (float *)(plan->arena + v->off32);
        if (v->off_stats >= 0) v->pstats = (float *)(plan->arena + v->off_stats);
    }

    // ---- constants ----
    P.fold_constants();
    if (!P.pool.host.empty()) {
        B200_CHECK(cudaMalloc((void **)&P.pool.dev, P.pool.host.size()));
        plan->owned_device.push_back(P.pool.dev);
        B200_CHECK(cudaMemcpy(P.pool.dev, P.pool.host.data(), P.pool.host.size(), cudaMemcpyHostToDevice));
        plan->weight_bytes += (int64_t)P.pool.host.size();
    }

    // ---- emit launches ----
    for (FNode * n : P.order) {
        FVal * o = n->out;
        std::string what = n->name;
        {
            char shp[160];
            const FVal * i0 = n->in.empty() ? nullptr : n->in[0];
            snprintf(shp, sizeof shp, " in[%dx%dx%dx%d] out[%dx%dx%dx%d]%s%s%s%s", i0 ? i0->N : 0, i0 ? i0->H : 0, i0 ? i0->W : 0, i0 ? i0->C : 0,
                     o->N, o->H, o->W, o->C, n->act ? " silu" : "", n->res ? " +res" : "", o->need16 ? " f16" : "", o->need32 ? " f32" : "");
            what += shp;
        }
        switch (n->kind) {
            case FK_INPUT: {
                o->p32 = (float *)device_ptr_of(plan, n->leaf);
            } break;
            case FK_STEM: {
                FVal * in = n->in[0];
                const FNode * src = in->prod;
                const int64_t H = in->H, W = in->W;
                int64_t sn, sy, sx, sc;
                if (src->chw) { sn = 3 * H * W; sy = W; sx = 1; sc = H * W; }
                else { sn = H * W * 3; sy = W * 3; sx = 3; sc = 1; }
                const float * x = in->p32;
                const __half * wt = P.pool.ptr<__half>(n->c_w);
                const float *scale = P.pool.ptr<float>(n->c_scale), *shift = P.pool.ptr<float>(n->c_shift);
                const int N = in->N, OC = o->C, act = n->act;
                __half * o16 = o->p16; float * o32 = o->p32;
                // u8 route: the same launch stages its patch from the quantised u8 images when the plan's flag word says so
                const uint8_t * x8 = nullptr;
                const int *     x8_flag = nullptr;
                if (src->kind == FK_INPUT && src->leaf && stem_takes_u8(OC, (int)W, !src->chw, o16 != nullptr, o32 != nullptr) && !plan->u8_input) {
                    void * buf = nullptr, * flag = nullptr;
                    B200_CHECK(cudaMalloc(&buf, (size_t)N * H * W * 3));
                    B200_CHECK(cudaMalloc(&flag, 256));
                    B200_CHECK(cudaMemset(flag, 0, 256));
                    plan->owned_device.push_back(buf);
                    plan->owned_device.push_back(flag);
                    plan->u8_input = (uint8_t *)buf;
                    plan->u8_flag  = (int *)flag;
                    plan->u8_leaf  = src->leaf;
                    x8 = plan->u8_input;
                    x8_flag = plan->u8_flag;
                }
                add_launch(plan, "stem_conv3x3s2_bn_silu", [=](cudaStream_t st) { launch_stem(x, sn, sy, sx, sc, N, (int)H, (int)W, wt, OC, scale, shift, act, o16, o32, st, x8, x8_flag); },
                           2.0 * o->rows() * OC * 27, (double)in->rows() * 12 + (double)o->rows() * OC * (o16 ? 2 : 0) + (double)o->rows() * OC * (o32 ? 4 : 0), what,
                           (double)in->rows() * 6 + (double)o->rows() * OC * 2);
            } break;
            case FK_DW: {
                FVal * in = n->in[0];
                // K4a (experimental, GGML_B200_DWREDUCE=1): depthwise whose only consumer is a 1x1 conv (the reduce of an
                // inverted residual) runs inside that conv's kernel; its output never goes to HBM.  Parity-green, but in
                // round 1 still slower than the two separate kernels (one CTA of 8 compute warps per SM), so off by default.
                if (getenv("GGML_B200_DWREDUCE") != nullptr && o->users.size() == 1 && o->users[0]->kind == FK_CONV1 &&
                    o->users[0]->in[0] == o && !out_set.count(o) && o->users[0]->out->C <= 256) {
                    n->fused_into_reduce = true;
                    break;
                }
                const __half * x = in->p16;
                const __half * wt = P.pool.ptr<__half>(n->c_w);
                const float *scale = P.pool.ptr<float>(n->c_scale), *shift = P.pool.ptr<float>(n->c_shift);
                const int N = in->N, H = in->H, W = in->W, C = in->C, stride = n->stride, act = n->act;
                __half * o16 = o->p16;
                auto DL = std::make_shared<DwLaunch>();
                if (getenv("GGML_B200_DW_V1") == nullptr && dw_prepare(*DL, x, N, H, W, C, stride, wt, scale, shift, act, o16)) {
                    add_launch(plan, "dwconv3x3_tma_bn_silu", [DL](cudaStream_t st) { dw_launch(*DL, st); }, 2.0 * o->rows() * C * 9,
                               ((double)in->rows() + (double)o->rows()) * C * 2, what);
                } else {
                    add_launch(plan, "dwconv3x3_bn_silu", [=](cudaStream_t st) { launch_dwconv(x, N, H, W, C, stride, wt, scale, shift, act, o16, st); },
                               2.0 * o->rows() * C * 9, ((double)in->rows() + (double)o->rows()) * C * 2, what);
                }
            } break;
            case FK_CONV1: case FK_LINEAR: case FK_QKV: {
                FVal * in = n->in[0];
                if (n->kind == FK_CONV1 && in->prod && in->prod->kind == FK_DW && in->prod->fused_into_reduce) {
                    FNode * dw = in->prod;
                    FVal *  xe = dw->in[0];  // the expanded activation
                    auto DL = std::make_shared<DwRedLaunch>();
                    if (dwreduce_prepare(*DL, xe->p16, xe->N, xe->H, xe->W, xe->C, dw->stride, P.pool.ptr<__half>(dw->c_w), P.pool.ptr<float>(dw->c_scale),
                                         P.pool.ptr<float>(dw->c_shift), dw->act, P.pool.ptr<__half>(n->c_w), o->C, P.pool.ptr<float>(n->c_scale),
                                         P.pool.ptr<float>(n->c_shift), n->act, n->res ? n->res->p32 : nullptr, o->p16, o->p32)) {
                        const double bytes = (double)xe->rows() * xe->C * 2 + (double)o->rows() * o->C * ((o->p16 ? 2 : 0) + (o->p32 ? 4 : 0) + (n->res ? 4 : 0));
                        add_launch(plan, "dwconv3x3_reduce1x1_fused", [DL](cudaStream_t st) { dwreduce_launch(*DL, st); },
                                   2.0 * in->rows() * in->C * 9 + 2.0 * in->rows() * o->C * in->C, bytes, "dw+reduce " + what);
                        break;
                    }
                    // shape not supported by the fused kernel: run the depthwise on its own after all
                    const __half * x = xe->p16;
                    const __half * wt = P.pool.ptr<__half>(dw->c_w);
                    const float *scale = P.pool.ptr<float>(dw->c_scale), *shift = P.pool.ptr<float>(dw->c_shift);
                    const int N = xe->N, H = xe->H, W = xe->W, C = xe->C, stride = dw->stride, act = dw->act;
                    __half * o16 = in->p16;
                    add_launch(plan, "dwconv3x3_bn_silu", [=](cudaStream_t st) { launch_dwconv(x, N, H, W, C, stride, wt, scale, shift, act, o16, st); },
                               2.0 * in->rows() * C * 9, ((double)xe->rows() + (double)in->rows()) * C * 2, "dw (unfused fallback)");
                }
                GemmEpilogue ep;
                ep.scale = P.pool.ptr<float>(n->c_scale);
                ep.shift = P.pool.ptr<float>(n->c_shift);
                ep.act   = n->act;
                if (n->res) { ep.res32 = n->res->p32; ep.ldr32 = n->res->C; }
                ep.out16 = o->p16; ep.ld16 = o->C;
                ep.out32 = o->p32; ep.ld32 = o->C;
                if (o->need_stats) ep.stats_out = o->pstats;
                if (n->ln_g) {
                    if (!in->pstats) return false;
                    ep.ln_stats = in->pstats;
                    ep.ln_c1    = P.pool.ptr<float>(n->c_c1);
                    ep.ln_inv_c = 1.0f / (float)in->C;
                    ep.ln_eps   = n->ln_eps;
                    what += " ln-folded";
                }
                if (o->need_stats) what += " +rowstats";
                auto L = std::make_shared<GemmLaunch>();
                if (!gemm_prepare(*L, in->p16, in->C, P.pool.ptr<__half>(n->c_w), in->C, (int)in->rows(), o->C, in->C, ep)) return false;
                const char * kname = n->kind == FK_CONV1 ? "gemm_tcgen05_conv1x1" : (n->kind == FK_QKV ? "gemm_tcgen05_qkv" : "gemm_tcgen05_linear");
                const double bytes = (double)in->rows() * in->C * 2 + (double)o->C * in->C * 2 + (double)o->rows() * o->C * ((o->p16 ? 2 : 0) + (o->p32 ? 4 : 0)) +
                                     (n->res ? (double)o->rows() * o->C * 4 : 0.0);
                add_launch(plan, kname, [L](cudaStream_t st) { gemm_launch(*L, st); }, 2.0 * in->rows() * o->C * in->C, bytes, what,
                           (double)in->rows() * in->C * 2 + (double)o->C * in->C * 2 + (double)o->rows() * o->C * 2);
            } break;
            case FK_IR: {
                FVal *  in = n->in[0];
                FNode * a = n->ir_expand, * b = n->ir_dw;
                const int E = a->out->C;
                auto IL = std::make_shared<IrLaunch>();
                if (!ir_fused_prepare(*IL, in->p16, in->N, in->H, in->W, in->C, E, o->C, b->stride, P.pool.ptr<__half>(a->c_w), P.pool.ptr<float>(a->c_scale),
                                      P.pool.ptr<float>(a->c_shift), P.pool.ptr<__half>(b->c_w), P.pool.ptr<float>(b->c_scale), P.pool.ptr<float>(b->c_shift),
                                      P.pool.ptr<__half>(n->c_w), P.pool.ptr<float>(n->c_scale), P.pool.ptr<float>(n->c_shift), n->res ? n->res->p32 : nullptr, o->p16,
                                      o->p32))
                    return false;
                char cfg[96];
                snprintf(cfg, sizeof cfg, " expand %d s%d tile %dx%d", E, b->stride, IL->p.TH, IL->p.TW);
                const double flops = 2.0 * in->rows() * E * in->C + 2.0 * o->rows() * E * 9 + 2.0 * o->rows() * o->C * E;
                const double bytes = (double)in->rows() * in->C * 2 + (double)o->rows() * o->C * ((o->p16 ? 2 : 0) + (o->p32 ? 4 : 0) + (n->res ? 4 : 0)) +
                                     2.0 * ((double)E * in->C + 9.0 * E + (double)o->C * E);
                add_launch(plan, "ir_fused_expand_dw_reduce", [IL](cudaStream_t st) { ir_fused_launch(*IL, st); }, flops, bytes, what + cfg,
                           (double)in->rows() * in->C * 2 + (double)o->rows() * o->C * 2 + 2.0 * ((double)E * in->C + 9.0 * E + (double)o->C * E));
            } break;
            case FK_VIT: {
                FVal * in = n->in[0];
                auto VL = std::make_shared<VitStageLaunch>();
                const int nl = (int)n->vit.size(), F = n->vit_F, C = in->C;
                if (!vit_stage_prepare(*VL, in->p32, in->N, in->H, in->W, C, n->heads, F, nl, n->eps, P.pool.ptr<uint8_t>(n->c_w), P.pool.ptr<float>(n->c_shift),
                                       o->p32, o->p16, o->need_stats ? o->pstats : nullptr))
                    return false;
                const double rows = (double)in->rows(), L = (double)(in->H / 2) * (in->W / 2);
                const bool   mlp  = n->heads == 0;
                const double flops = mlp ? 4.0 * rows * C * F : nl * (2.0 * rows * C * (4.0 * C + 2.0 * F) + 4.0 * in->N * 4 * L * L * C);
                const double wbytes = mlp ? 4.0 * C * F : nl * 2.0 * C * (4.0 * C + 2.0 * F);
                char cfg[96];
                if (mlp) snprintf(cfg, sizeof cfg, " ffn %d, %d tiles", F, VL->p.tiles);
                else snprintf(cfg, sizeof cfg, " %d layers, %d heads, ffn %d, L=%d, %d tiles", nl, n->heads, F, (int)L, VL->p.tiles);
                add_launch(plan, mlp ? "vit_mlp_fused" : "vit_stage_fused", [VL](cudaStream_t st) { vit_stage_launch(*VL, st); }, flops,
                           rows * C * (4 + (o->p16 ? 2 : 0) + (o->p32 ? 4 : 0)) + wbytes, what + cfg, rows * C * 4 + wbytes);
            } break;
            case FK_CONV3: {
                FVal * a = n->in[0];
                FVal * b = n->in.size() > 1 ? n->in[1] : nullptr;
                GemmEpilogue ep;
                ep.scale = P.pool.ptr<float>(n->c_scale);
                ep.shift = P.pool.ptr<float>(n->c_shift);
                ep.act   = n->act;
                ep.out16 = o->p16; ep.ld16 = o->C;
                ep.out32 = o->p32; ep.ld32 = o->C;
                auto L = std::make_shared<GemmLaunch>();
                if (!conv3x3_prepare(*L, a->p16, a->C, b ? b->p16 : nullptr, b ? b->C : 0, a->N, a->H, a->W, P.pool.ptr<__half>(n->c_w), o->C, ep,
                                     P.pool.ptr<uint8_t>(n->c_w2)))
                    return false;
                if (L->p.conv == 2) what += " halo";
                const int ict = a->C + (b ? b->C : 0);
                const double bytes = (double)a->rows() * ict * 2 + (double)o->C * 9 * ict * 2 + (double)o->rows() * o->C * ((o->p16 ? 2 : 0) + (o->p32 ? 4 : 0));
                add_launch(plan, "conv3x3_tcgen05_implicit_gemm", [L](cudaStream_t st) { gemm_launch(*L, st); }, 2.0 * a->rows() * o->C * 9 * ict, bytes, what,
                           (double)a->rows() * ict * 2 + (double)o->C * 9 * ict * 2 + (double)o->rows() * o->C * 2);
            } break;
            case FK_LN: {
                FVal * in = n->in[0];
                const float * x = in->p32;
                const float * g = (const float *)device_ptr_of(plan, n->g);
                const float * b = (const float *)device_ptr_of(plan, n->b);
                const int64_t rows = in->rows();
                const int C = in->C;
                const float eps = n->eps;
                __half * o16 = o->p16; float * o32 = o->p32;
                add_launch(plan, "layernorm_f32_to_f16", [=](cudaStream_t st) { launch_layernorm(x, rows, C, g, b, eps, o16, o32, st); }, 8.0 * rows * C,
                           (double)rows * C * (4 + (o16 ? 2 : 0) + (o32 ? 4 : 0)), what);
            } break;
            case FK_ATTN: {
                FVal * qkv = n->in[0];
                const __half * q = qkv->p16;
                const int N = qkv->N, H = qkv->H, W = qkv->W, C = o->C, heads = n->heads;
                __half * o16 = o->p16;
                const double L = (double)(H / 2) * (W / 2);
                add_launch(plan, "attention_mma_flash", [=](cudaStream_t st) { launch_attention(q, N, H, W, C, heads, o16, st); }, 4.0 * N * 4 * L * L * C,
                           (double)qkv->rows() * (qkv->C + C) * 2, what);
            } break;
            case FK_ADD: {
                const float *a = n->in[0]->p32, *b = n->in[1]->p32;
                const int64_t cnt = o->rows() * o->C;
                float * o32 = o->p32; __half * o16 = o->p16;
                add_launch(plan, "residual_add", [=](cudaStream_t st) { launch_add(a, b, cnt, o32, o16, st); }, (double)cnt, (double)cnt * (8 + (o32 ? 4 : 0) + (o16 ? 2 : 0)), what);
            } break;
            case FK_POOL: {
                // the pooled vector IS the graph output layout ([1,1,C,N] == [N][C]); written straight to the output slot below
            } break;
        }
    }
    // ---- graph outputs: convert to ggml's (W,H,C,N) f32 ----
    for (size_t i = 0; i < outs.size(); i++) {
        ggml_tensor * t = outs[i];
        FVal * v = out_vals[i];
        float * dst = (float *)(plan->arena + out_off[i]);
        Slot s;
        s.kind = SLOT_ARENA;
        s.dptr = dst;
        s.bytes = (int64_t)ggml_nelements(t) * 4;
        plan->slots[t] = s;
        if (heads.count(t)) {
            const Head & h = heads[t];
            const float * pooled = (const float *)plan->slots.at(h.pooled).dptr;  // node order: the pooled output was emitted before its consumer
            const int N = (int)t->ne[1], C = (int)h.pooled->ne[2], OUT = (int)t->ne[0];
            const float * W  = P.pool.ptr<float>(head_w[t]);
            const float * bb = P.pool.ptr<float>(head_b[t]);
            add_launch(plan, "classifier_head_f32", [=](cudaStream_t st) { launch_head_linear(pooled, W, bb, N, C, OUT, dst, st); }, 2.0 * N * C * OUT,
                       4.0 * ((double)C * OUT + (double)N * (C + OUT)), t->name);
        } else if (t->op == GGML_OP_POOL_MEAN_HW) {
            FVal * in = v->prod->in[0];
            const __half * x16 = in->p32 ? nullptr : in->p16;
            const float * x32 = in->p32;
            const int N = in->N, HW = in->H * in->W, C = in->C;
            add_launch(plan, "pool_mean", [=](cudaStream_t st) { launch_pool_mean(x16, x32, N, HW, C, dst, st); }, (double)N * HW * C, (double)N * HW * C * 4, "pooled");
        } else {
            const __half * x16 = v->p32 ? nullptr : v->p16;
            const float * x32 = v->p32;
            const int N = v->N, H = v->H, W = v->W, C = v->C;
            add_launch(plan, "nhwc_to_ggml_layout", [=](cudaStream_t st) { launch_nhwc_to_nchw(x16, x32, N, H, W, C, dst, st); }, 0.0,
                       (double)v->rows() * C * (4 + (x32 ? 4 : 2)), t->name);
        }
    }
    return true;
}

bool build_fast_plan(Plan * plan, ggml_cgraph * gf) {
    const size_t n_owned = plan->owned_device.size();
    if (build_fast_plan_impl(plan, gf)) return true;
    // leave the plan exactly as place_leafs() left it, so the exact builder can take over
    plan->launches.clear();
    plan->meta.clear();
    for (size_t i = n_owned; i < plan->owned_device.size(); i++) cudaFree(plan->owned_device[i]);
    plan->owned_device.resize(n_owned);
    for (int i = 0; i < gf->n_nodes; i++) plan->slots.erase(gf->nodes[i]);
    plan->arena = nullptr;
    plan->arena_bytes = plan->naive_bytes = 0;
    plan->n_folded = 0;
    return false;
}

}  // namespace b200
