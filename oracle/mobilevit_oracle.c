/*
 * mobilevit_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the MobileViT forward pass that /root/reference/mobilevit/main.cpp builds as
 * a ggml graph, with the numerics of the (un-vendored, un-pinned) upstream ggml CPU kernels it runs on.
 * Nothing in the product (ggml-experiments_b200/, include/) includes, links or calls this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * PARITY UNPINNED: the reference has no tests and its single known answer (mobilevit/README.md:39-45)
 * needs pretrained weights that cannot be fetched here; upstream ggml (github.com/ggerganov/ggml,
 * HEAD of ~Feb-May 2024) is not in the container.  What pins this oracle instead is a
 * cross-check against the Hugging Face torch MobileViTModel (the model convert-tf-to-ggml.py exports)
 * in pure-f32 mode (tests/golden/, tests/gen_golden.py).
 *
 * Layout convention: ggml ne=(n0,n1,n2,n3), n0 fastest == C array [n3][n2][n1][n0].  An activation
 * ne=(W,H,C,1) is therefore a CHW float array.  The oracle, like the reference graph, is batch-1;
 * batches are a loop over images (main.cpp:612 hard-codes N=1).
 *
 * Each function cites the reference lines it follows.  [ggml] marks restated upstream-ggml semantics.
 */
#include <immintrin.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#include <unistd.h>

#define MVO_MAX_TENSORS 512
#define MVO_FLAG_PURE_F32 1 /* skip every f16 rounding point: == HF torch f32 semantics */
#define MVO_FLAG_LEGACY_F16_TABLES 4 /* SiLU and the softmax exponential through f16 lookup tables, y = f16(f(f16(x))) -- the ggml the
                                      author ran (GGML_SILU_FP16 / ggml_table_exp_f16; mobilevit/README.md:39-45 prints f16-representable values) */
#define MVO_FLAG_NO_ACT_ROUND 2 /* weights stay f16-rounded (as loaded, main.cpp:928-932) but activations are not rounded:
                                  a smooth function, used to validate the GPU EXACT_F32 mode to max-abs 1e-3 */

/* ------------------------------------------------------------------------------------------------
 * fp16 rounding: [ggml] GGML_FP32_TO_FP16 on x86 with F16C = _cvtss_sh(x, 0) (round-to-nearest-even)
 * ------------------------------------------------------------------------------------------------ */
static inline float round_f16(float x) { return _cvtsh_ss(_cvtss_sh(x, 0)); }

/* legacy f16 lookup tables ([ggml] ggml_table_silu_f16 / ggml_table_exp_f16), built on first use; the flag is process-wide state set by
 * mvo_forward / mvo_forward_batch before any worker thread starts */
static int g_legacy_tables = 0;
static uint16_t *g_tab_silu = NULL, *g_tab_exp = NULL;
static void build_legacy_tables(void) {
    if (g_tab_silu) return;
    uint16_t *ts = (uint16_t *)malloc(65536 * 2), *te = (uint16_t *)malloc(65536 * 2);
    for (int i = 0; i < 65536; i++) {
        const float f = _cvtsh_ss((uint16_t)i);
        ts[i] = _cvtss_sh(f / (1.0f + expf(-f)), 0);
        te[i] = _cvtss_sh(expf(f), 0);
    }
    g_tab_exp = te;
    g_tab_silu = ts;
}
static inline float silu_op(float x) { return g_legacy_tables ? _cvtsh_ss(g_tab_silu[_cvtss_sh(x, 0)]) : x / (1.0f + expf(-x)); }
static inline float exp_op(float x) { return g_legacy_tables ? _cvtsh_ss(g_tab_exp[_cvtss_sh(x, 0)]) : expf(x); }

static void round_f16_array(const float *src, float *dst, size_t n, int pure_f32) {
    if (pure_f32) {
        if (dst != src) memcpy(dst, src, n * sizeof(float));
        return;
    }
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        __m128i h = _mm256_cvtps_ph(_mm256_loadu_ps(src + i), 0);
        _mm256_storeu_ps(dst + i, _mm256_cvtph_ps(h));
    }
    for (; i < n; i++) dst[i] = round_f16(src[i]);
}

/* ------------------------------------------------------------------------------------------------
 * weight file: main.cpp:872-942 / convert-tf-to-ggml.py:16-33
 *   repeat { int32 name_len; char name[name_len]; int32 n_dims; int32 dims[n_dims] (TF order);
 *            float32 data[prod(dims)] }
 * ggml ne = dims reversed, same bytes (main.cpp:905-917).  Unlike main.cpp:939 we stop at clean EOF.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    char name[160];
    int n_dims;
    int dims[4]; /* TF order, as stored */
    float *data;
    size_t n;
} mvo_tensor;

typedef struct {
    const mvo_tensor *kernel, *gamma, *beta, *mean, *var; /* main.cpp:56-63 */
} mvo_conv;

typedef struct {
    mvo_conv expand, dw, reduce; /* main.cpp:75-87 */
    int stride;
} mvo_ir;

typedef struct {
    const mvo_tensor *qk, *qb, *kk, *kb, *vk, *vb, *ok, *ob, *ik, *ib, *dk, *db; /* main.cpp:108-137 */
    const mvo_tensor *lb_g, *lb_b, *la_g, *la_b;
} mvo_tlayer;

typedef struct {
    mvo_ir down;
    mvo_conv kxk, c1x1, proj, fusion; /* main.cpp:152-177 */
    const mvo_tensor *ln_g, *ln_b;
    int n_tlayers;
    mvo_tlayer tl[8];
} mvo_vit;

typedef struct {
    int n_tensors;
    mvo_tensor t[MVO_MAX_TENSORS];
    mvo_conv stem, exp1x1;
    int n_l1, n_l2;
    mvo_ir l1[4], l2[4];
    mvo_vit vit[3];
    int num_heads;
} mvo_model;

static const mvo_tensor *find_tensor(const mvo_model *m, const char *name) {
    for (int i = 0; i < m->n_tensors; i++)
        if (strcmp(m->t[i].name, name) == 0) return &m->t[i];
    return NULL;
}

static const mvo_tensor *need_tensor(const mvo_model *m, const char *name) {
    const mvo_tensor *t = find_tensor(m, name);
    if (!t) { /* main.cpp:225 tensors.at() throws */
        fprintf(stderr, "mvo: missing tensor %s\n", name);
        abort();
    }
    return t;
}

static void bind_conv(const mvo_model *m, mvo_conv *c, const char *path, int use_norm) {
    char buf[512]; /* main.cpp:218-234 */
    snprintf(buf, sizeof buf, "%s/convolution/kernel:0", path);
    c->kernel = need_tensor(m, buf);
    c->gamma = c->beta = c->mean = c->var = NULL;
    if (use_norm) {
        snprintf(buf, sizeof buf, "%s/normalization/gamma:0", path);           c->gamma = need_tensor(m, buf);
        snprintf(buf, sizeof buf, "%s/normalization/beta:0", path);            c->beta = need_tensor(m, buf);
        snprintf(buf, sizeof buf, "%s/normalization/moving_mean:0", path);     c->mean = need_tensor(m, buf);
        snprintf(buf, sizeof buf, "%s/normalization/moving_variance:0", path); c->var = need_tensor(m, buf);
    }
}

static void bind_ir(const mvo_model *m, mvo_ir *ir, const char *path, int stride) {
    char buf[512]; /* main.cpp:346-356 */
    snprintf(buf, sizeof buf, "%s/expand_1x1", path); bind_conv(m, &ir->expand, buf, 1);
    snprintf(buf, sizeof buf, "%s/conv_3x3", path);   bind_conv(m, &ir->dw, buf, 1);
    snprintf(buf, sizeof buf, "%s/reduce_1x1", path); bind_conv(m, &ir->reduce, buf, 1);
    ir->stride = stride;
}

#define MVO_P "tf_mobile_vi_t_model/mobilevit"

void mvo_free(void *model) {
    mvo_model *m = (mvo_model *)model;
    if (!m) return;
    for (int i = 0; i < m->n_tensors; i++) free(m->t[i].data);
    free(m);
}

void *mvo_load(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    mvo_model *m = (mvo_model *)calloc(1, sizeof(mvo_model));
    for (;;) {
        int32_t name_len, n_dims;
        if (fread(&name_len, 4, 1, f) != 1) break; /* clean EOF */
        if (m->n_tensors >= MVO_MAX_TENSORS || name_len <= 0 || name_len >= 160) { fclose(f); mvo_free(m); return NULL; }
        mvo_tensor *t = &m->t[m->n_tensors];
        if (fread(t->name, 1, name_len, f) != (size_t)name_len) { fclose(f); mvo_free(m); return NULL; }
        t->name[name_len] = 0;
        if (fread(&n_dims, 4, 1, f) != 1) { fclose(f); mvo_free(m); return NULL; }
        /* record-format extension (convert-tf-to-ggml.py:13-14 TODOs): bit 16 = f16 payload, bit 17 = dense kernel stored (out, in) */
        const int disk_f16 = (n_dims >> 16) & 1, disk_t = (n_dims >> 17) & 1;
        n_dims &= 0xFFFF;
        if (n_dims < 1 || n_dims > 4 || (disk_t && n_dims != 2)) { fclose(f); mvo_free(m); return NULL; }
        t->n_dims = n_dims;
        t->n = 1;
        for (int i = 0; i < 4; i++) t->dims[i] = 1;
        for (int i = 0; i < n_dims; i++) {
            int32_t d;
            if (fread(&d, 4, 1, f) != 1) { fclose(f); mvo_free(m); return NULL; }
            t->dims[i] = d;
            t->n *= (size_t)d;
        }
        t->data = (float *)malloc(t->n * sizeof(float));
        if (disk_f16) {
            uint16_t *hbuf = (uint16_t *)malloc(t->n * 2);
            if (fread(hbuf, 2, t->n, f) != t->n) { fclose(f); free(hbuf); free(t->data); mvo_free(m); return NULL; }
            for (size_t i = 0; i < t->n; i++) t->data[i] = _cvtsh_ss(hbuf[i]);
            free(hbuf);
        } else if (fread(t->data, sizeof(float), t->n, f) != t->n) { fclose(f); free(t->data); mvo_free(m); return NULL; }
        if (disk_t) { /* back to the canonical (in, out) */
            const size_t n_out = (size_t)t->dims[0], n_in = (size_t)t->dims[1];
            float *tr = (float *)malloc(t->n * sizeof(float));
            for (size_t oo = 0; oo < n_out; oo++)
                for (size_t ii = 0; ii < n_in; ii++) tr[ii * n_out + oo] = t->data[oo * n_in + ii];
            free(t->data);
            t->data = tr;
            t->dims[0] = (int)n_in; t->dims[1] = (int)n_out;
        }
        m->n_tensors++;
    }
    fclose(f);

    /* structure: main.cpp:314-515, but stage counts come from the names present */
    char buf[512];
    bind_conv(m, &m->stem, MVO_P "/conv_stem", 1);
    bind_conv(m, &m->exp1x1, MVO_P "/conv_1x1_exp", 1);
    for (m->n_l1 = 0; m->n_l1 < 4; m->n_l1++) {
        snprintf(buf, sizeof buf, MVO_P "/encoder/layer.0/layer.%d/expand_1x1/convolution/kernel:0", m->n_l1);
        if (!find_tensor(m, buf)) break;
        snprintf(buf, sizeof buf, MVO_P "/encoder/layer.0/layer.%d", m->n_l1);
        bind_ir(m, &m->l1[m->n_l1], buf, 1); /* main.cpp:337,352 */
    }
    for (m->n_l2 = 0; m->n_l2 < 4; m->n_l2++) {
        snprintf(buf, sizeof buf, MVO_P "/encoder/layer.1/layer.%d/expand_1x1/convolution/kernel:0", m->n_l2);
        if (!find_tensor(m, buf)) break;
        snprintf(buf, sizeof buf, MVO_P "/encoder/layer.1/layer.%d", m->n_l2);
        bind_ir(m, &m->l2[m->n_l2], buf, m->n_l2 == 0 ? 2 : 1); /* main.cpp:368,383 */
    }
    for (int v = 0; v < 3; v++) { /* main.cpp:280-312,393-503 */
        mvo_vit *L = &m->vit[v];
        char base[128];
        snprintf(base, sizeof base, MVO_P "/encoder/layer.%d", v + 2);
        snprintf(buf, sizeof buf, "%s/downsampling_layer", base); bind_ir(m, &L->down, buf, 2);
        snprintf(buf, sizeof buf, "%s/conv_kxk", base);           bind_conv(m, &L->kxk, buf, 1);
        snprintf(buf, sizeof buf, "%s/conv_1x1", base);           bind_conv(m, &L->c1x1, buf, 0); /* main.cpp:293 */
        snprintf(buf, sizeof buf, "%s/conv_projection", base);    bind_conv(m, &L->proj, buf, 1);
        snprintf(buf, sizeof buf, "%s/fusion", base);             bind_conv(m, &L->fusion, buf, 1);
        snprintf(buf, sizeof buf, "%s/layernorm/gamma:0", base);  L->ln_g = need_tensor(m, buf);
        snprintf(buf, sizeof buf, "%s/layernorm/beta:0", base);   L->ln_b = need_tensor(m, buf);
        for (L->n_tlayers = 0; L->n_tlayers < 8; L->n_tlayers++) {
            char tb[256];
            snprintf(tb, sizeof tb, "%s/transformer/layer.%d", base, L->n_tlayers);
            snprintf(buf, sizeof buf, "%s/attention/attention/query/kernel:0", tb);
            if (!find_tensor(m, buf)) break;
            mvo_tlayer *T = &L->tl[L->n_tlayers]; /* main.cpp:238-277 */
#define BIND(field, suffix) snprintf(buf, sizeof buf, "%s/" suffix, tb); T->field = need_tensor(m, buf)
            BIND(qk, "attention/attention/query/kernel:0"); BIND(qb, "attention/attention/query/bias:0");
            BIND(kk, "attention/attention/key/kernel:0");   BIND(kb, "attention/attention/key/bias:0");
            BIND(vk, "attention/attention/value/kernel:0"); BIND(vb, "attention/attention/value/bias:0");
            BIND(ok, "attention/output/dense/kernel:0");    BIND(ob, "attention/output/dense/bias:0");
            BIND(ik, "intermediate/dense/kernel:0");        BIND(ib, "intermediate/dense/bias:0");
            BIND(dk, "output/dense/kernel:0");              BIND(db, "output/dense/bias:0");
            BIND(lb_g, "layernorm_before/gamma:0");         BIND(lb_b, "layernorm_before/beta:0");
            BIND(la_g, "layernorm_after/gamma:0");          BIND(la_b, "layernorm_after/beta:0");
#undef BIND
        }
    }
    m->num_heads = 4; /* main.cpp:41 */
    return m;
}

int mvo_num_tensors(void *model) { return ((mvo_model *)model)->n_tensors; }
int mvo_out_channels(void *model) { return ((mvo_model *)model)->exp1x1.kernel->dims[3]; }
long mvo_num_weights(void *model) {
    mvo_model *m = (mvo_model *)model;
    long n = 0;
    for (int i = 0; i < m->n_tensors; i++) n += (long)m->t[i].n;
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * C[M][N] = A[M][K] * B[K][N], f32, per-element accumulation sequential in k.
 * [ggml] ggml_mul_mat accumulates each dot product in f32 (vec_dot_f32 / vec_dot_f16 with f32 lanes).
 * ------------------------------------------------------------------------------------------------ */
#define GB_M 4
#define GB_N 16
void mvo_gemm(int M, int N, int K, const float *A, int lda, const float *B, int ldb, float *C, int ldc) {
    for (int i0 = 0; i0 < M; i0 += GB_M) {
        int mb = M - i0 < GB_M ? M - i0 : GB_M;
        int j0 = 0;
        for (; j0 + GB_N <= N; j0 += GB_N) {
            float acc[GB_M][GB_N];
            for (int i = 0; i < GB_M; i++)
                for (int j = 0; j < GB_N; j++) acc[i][j] = 0.f;
            if (mb == GB_M) {
                for (int k = 0; k < K; k++) {
                    const float *b = B + (size_t)k * ldb + j0;
                    for (int i = 0; i < GB_M; i++) {
                        float a = A[(size_t)(i0 + i) * lda + k];
                        for (int j = 0; j < GB_N; j++) acc[i][j] += a * b[j];
                    }
                }
            } else {
                for (int k = 0; k < K; k++) {
                    const float *b = B + (size_t)k * ldb + j0;
                    for (int i = 0; i < mb; i++) {
                        float a = A[(size_t)(i0 + i) * lda + k];
                        for (int j = 0; j < GB_N; j++) acc[i][j] += a * b[j];
                    }
                }
            }
            for (int i = 0; i < mb; i++)
                for (int j = 0; j < GB_N; j++) C[(size_t)(i0 + i) * ldc + j0 + j] = acc[i][j];
        }
        for (; j0 < N; j0++) {
            for (int i = 0; i < mb; i++) {
                float s = 0.f;
                for (int k = 0; k < K; k++) s += A[(size_t)(i0 + i) * lda + k] * B[(size_t)k * ldb + j0];
                C[(size_t)(i0 + i) * ldc + j0] = s;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * ggml_conv_2d (main.cpp:798): [ggml] im2col (activations -> F16, zero padding, column order (ic,kh,kw)
 * with kw fastest) + mul_mat(F16 kernel, F16 columns) with f32 accumulation, result F32.
 *   x   : [C][H][W] f32          k : TF layout (KH,KW,IC,OC) as stored in the file (App. B)
 *   out : [OC][OH][OW] f32       OH = (H + 2p - (KH-1) - 1)/s + 1,  p = (KW-1)/2 (main.cpp:784)
 * The kernel is rounded to F16 at load time in the reference (main.cpp:887,928-932); here on the fly.
 * ------------------------------------------------------------------------------------------------ */
void mvo_conv2d(const float *x, int C, int H, int W, const float *k_tf, int KH, int KW, int OC, int stride,
                float *out, int flags) {
    const int pure = flags & MVO_FLAG_PURE_F32;
    const int pure_act = flags & (MVO_FLAG_PURE_F32 | MVO_FLAG_NO_ACT_ROUND);
    const int p = (KW - 1) / 2;
    const int OH = (H + 2 * p - (KH - 1) - 1) / stride + 1;
    const int OW = (W + 2 * p - (KW - 1) - 1) / stride + 1;
    const int K = C * KH * KW;
    const size_t P = (size_t)OH * OW;
    /* kernel: file (KH,KW,IC,OC) -> ggml ne=(OC,IC,KW,KH) -> permute(3,2,0,1)+cont -> ne=(KW,KH,IC,OC)
     * == C array [OC][IC][KH][KW] (main.cpp:790-805). */
    float *wk = (float *)malloc((size_t)OC * K * sizeof(float));
    for (int oc = 0; oc < OC; oc++)
        for (int ic = 0; ic < C; ic++)
            for (int kh = 0; kh < KH; kh++)
                for (int kw = 0; kw < KW; kw++) {
                    float v = k_tf[(((size_t)kh * KW + kw) * C + ic) * OC + oc];
                    wk[(size_t)oc * K + ((size_t)ic * KH + kh) * KW + kw] = pure ? v : round_f16(v);
                }
    /* im2col, transposed: col[k][pixel] */
    float *col = (float *)malloc((size_t)K * P * sizeof(float));
    for (int ic = 0; ic < C; ic++)
        for (int kh = 0; kh < KH; kh++)
            for (int kw = 0; kw < KW; kw++) {
                float *dst = col + ((size_t)(ic * KH + kh) * KW + kw) * P;
                for (int oy = 0; oy < OH; oy++) {
                    int iy = oy * stride + kh - p;
                    for (int ox = 0; ox < OW; ox++) {
                        int ix = ox * stride + kw - p;
                        float v = 0.f;
                        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[((size_t)ic * H + iy) * W + ix];
                        dst[(size_t)oy * OW + ox] = v;
                    }
                }
            }
    round_f16_array(col, col, (size_t)K * P, pure_act);
    mvo_gemm(OC, (int)P, K, wk, K, col, (int)P, out, (int)P);
    free(col);
    free(wk);
}

/* ggml_conv_depthwise_2d (main.cpp:788): [ggml] per-channel im2col(F16) + dot, f32 accumulation.
 * kernel file layout (KH,KW,1,C) (App. B). */
void mvo_dwconv2d(const float *x, int C, int H, int W, const float *k_tf, int KH, int KW, int stride, float *out,
                  int flags) {
    const int pure = flags & MVO_FLAG_PURE_F32;
    const int p = (KW - 1) / 2;
    const int OH = (H + 2 * p - (KH - 1) - 1) / stride + 1;
    const int OW = (W + 2 * p - (KW - 1) - 1) / stride + 1;
    float *xr = (float *)malloc((size_t)C * H * W * sizeof(float));
    round_f16_array(x, xr, (size_t)C * H * W, flags & (MVO_FLAG_PURE_F32 | MVO_FLAG_NO_ACT_ROUND));
    for (int c = 0; c < C; c++) {
        float wk[49];
        for (int kh = 0; kh < KH; kh++)
            for (int kw = 0; kw < KW; kw++) {
                float v = k_tf[((size_t)kh * KW + kw) * C + c];
                wk[kh * KW + kw] = pure ? v : round_f16(v);
            }
        const float *xc = xr + (size_t)c * H * W;
        float *oc = out + (size_t)c * OH * OW;
        for (int oy = 0; oy < OH; oy++)
            for (int ox = 0; ox < OW; ox++) {
                float s = 0.f;
                for (int kh = 0; kh < KH; kh++) {
                    int iy = oy * stride + kh - p;
                    if (iy < 0 || iy >= H) continue;
                    for (int kw = 0; kw < KW; kw++) {
                        int ix = ox * stride + kw - p;
                        if (ix < 0 || ix >= W) continue;
                        s += wk[kh * KW + kw] * xc[(size_t)iy * W + ix];
                    }
                }
                oc[(size_t)oy * OW + ox] = s;
            }
    }
    free(xr);
}

/* BatchNorm chain, main.cpp:809-846: ((x - mean) / sqrt(var + 1e-5)) * gamma + beta, all f32, in that
 * order (sub, add-eps, sqrt, div, mul, add).  SiLU main.cpp:848-850: [ggml] x / (1 + expf(-x)). */
void mvo_bn_silu(float *x, int C, size_t P, const float *mean, const float *var, const float *gamma,
                 const float *beta, int use_norm, int use_act) {
    for (int c = 0; c < C; c++) {
        float *xc = x + (size_t)c * P;
        if (use_norm) {
            const float mu = mean[c], g = gamma[c], b = beta[c];
            const float sd = sqrtf(var[c] + 1.0e-5f);
            for (size_t i = 0; i < P; i++) xc[i] = ((xc[i] - mu) / sd) * g + b;
        }
        if (use_act)
            for (size_t i = 0; i < P; i++) xc[i] = silu_op(xc[i]);
    }
}

/* mobilevit_conv_layer::forward, main.cpp:771-852.  Returns malloc'd [OC][OH][OW]. */
static float *conv_layer(const mvo_conv *c, const float *x, int C, int H, int W, int stride, int use_norm,
                         int use_act, int depthwise, int flags, int *oC, int *oH, int *oW) {
    const mvo_tensor *k = c->kernel; /* file dims (KH,KW,IC,OC) */
    const int KH = k->dims[0], KW = k->dims[1], IC = k->dims[2], OC = k->dims[3];
    const int p = (KW - 1) / 2;
    const int OH = (H + 2 * p - (KH - 1) - 1) / stride + 1;
    const int OW = (W + 2 * p - (KW - 1) - 1) / stride + 1;
    float *out = (float *)malloc((size_t)OC * OH * OW * sizeof(float));
    if (depthwise) {
        if (IC != 1 || OC != C) { fprintf(stderr, "mvo: depthwise shape mismatch\n"); abort(); }
        mvo_dwconv2d(x, C, H, W, k->data, KH, KW, stride, out, flags);
    } else {
        if (IC != C) { fprintf(stderr, "mvo: conv IC %d != C %d\n", IC, C); abort(); }
        mvo_conv2d(x, C, H, W, k->data, KH, KW, OC, stride, out, flags);
    }
    mvo_bn_silu(out, OC, (size_t)OH * OW, use_norm ? c->mean->data : NULL, use_norm ? c->var->data : NULL,
                use_norm ? c->gamma->data : NULL, use_norm ? c->beta->data : NULL, use_norm, use_act);
    *oC = OC; *oH = OH; *oW = OW;
    return out;
}

/* inverted_residual_layer::forward, main.cpp:854-870.  The residual condition uses the actual channel
 * counts (kernel shapes) instead of the hard-coded MobileViT-S hparams. */
static float *inverted_residual(const mvo_ir *ir, const float *x, int C, int H, int W, int flags, int *oC, int *oH,
                                int *oW) {
    int c1, h1, w1, c2, h2, w2, c3, h3, w3;
    float *a = conv_layer(&ir->expand, x, C, H, W, 1, 1, 1, 0, flags, &c1, &h1, &w1);
    float *b = conv_layer(&ir->dw, a, c1, h1, w1, ir->stride, 1, 1, 1, flags, &c2, &h2, &w2);
    free(a);
    float *c = conv_layer(&ir->reduce, b, c2, h2, w2, 1, 1, 0, 0, flags, &c3, &h3, &w3);
    free(b);
    if (ir->stride == 1 && c3 == C) { /* main.cpp:866-868 */
        size_t n = (size_t)c3 * h3 * w3;
        for (size_t i = 0; i < n; i++) c[i] = c[i] + x[i];
    }
    *oC = c3; *oH = h3; *oW = w3;
    return c;
}

/* ------------------------------------------------------------------------------------------------
 * ggml_permute + ggml_cont on a 4-d tensor: [ggml] result.ne[axis[i]] = a.ne[i] (strides likewise);
 * cont copies to the contiguous order of the result dims.  src is contiguous with dims ne[].
 * ------------------------------------------------------------------------------------------------ */
void mvo_permute_cont(const float *src, const int64_t ne[4], const int axis[4], float *dst, int64_t out_ne[4]) {
    int64_t nb[4] = {1, ne[0], ne[0] * ne[1], ne[0] * ne[1] * ne[2]}; /* element strides of src */
    int64_t rne[4], rnb[4];
    for (int i = 0; i < 4; i++) { rne[axis[i]] = ne[i]; rnb[axis[i]] = nb[i]; }
    for (int i = 0; i < 4; i++) out_ne[i] = rne[i];
    size_t o = 0;
    for (int64_t i3 = 0; i3 < rne[3]; i3++)
        for (int64_t i2 = 0; i2 < rne[2]; i2++)
            for (int64_t i1 = 0; i1 < rne[1]; i1++)
                for (int64_t i0 = 0; i0 < rne[0]; i0++)
                    dst[o++] = src[i0 * rnb[0] + i1 * rnb[1] + i2 * rnb[2] + i3 * rnb[3]];
}

/* mobile_vit_layer::unfolding, main.cpp:721-747.  in: ne=(W,H,C,1)  out: ne=(C, n_patches, ps*ps, 1) */
void mvo_unfold(const float *x, int C, int H, int W, int ps, float *out) {
    const int nph = H / ps, npw = W / ps;
    size_t n = (size_t)C * H * W;
    float *t1 = (float *)malloc(n * sizeof(float));
    float *t2 = (float *)malloc(n * sizeof(float));
    int64_t one[4];
    /* step 2-3: reshape (ps, npw, ps, C*nph) ; permute(0,2,1,3) ; cont */
    int64_t ne_a[4] = {ps, npw, ps, (int64_t)C * nph};
    int ax_a[4] = {0, 2, 1, 3};
    mvo_permute_cont(x, ne_a, ax_a, t1, one);
    /* step 4-5: reshape (ps*ps, npw*nph, C, 1) ; permute(2,1,0,3) ; cont */
    int64_t ne_b[4] = {(int64_t)ps * ps, (int64_t)npw * nph, C, 1};
    int ax_b[4] = {2, 1, 0, 3};
    mvo_permute_cont(t1, ne_b, ax_b, t2, one);
    memcpy(out, t2, n * sizeof(float));
    free(t1);
    free(t2);
}

/* mobile_vit_layer::folding, main.cpp:750-768.  in: ne=(C, n_patches, ps*ps)  out: ne=(W,H,C,1), square */
void mvo_fold(const float *x, int C, int n_patches, int ps, float *out) {
    const int np = (int)(sqrt(1.0 * n_patches)); /* main.cpp:754 */
    const int pa = ps * ps;
    size_t n = (size_t)C * n_patches * pa;
    float *t1 = (float *)malloc(n * sizeof(float));
    int64_t one[4];
    int64_t ne_a[4] = {C, n_patches, pa, 1};
    int ax_a[4] = {2, 1, 0, 3};
    mvo_permute_cont(x, ne_a, ax_a, t1, one); /* -> (pa, n_patches, C, 1) */
    int64_t ne_b[4] = {ps, ps, np, (int64_t)np * C};
    int ax_b[4] = {0, 2, 1, 3};
    mvo_permute_cont(t1, ne_b, ax_b, out, one); /* -> (ps, np, ps, np*C) == (W,H,C) */
    free(t1);
}

/* ggml_norm (main.cpp:1006,1118,1196) + gamma/beta: [ggml] ggml_compute_forward_norm_f32 -- sums in
 * double, mean/variance cast to float, y = (x-mean) * (1/sqrtf(var+eps)); then mul gamma, add beta. */
void mvo_layernorm(const float *x, int C, size_t rows, const float *g, const float *b, float eps, float *out) {
    for (size_t r = 0; r < rows; r++) {
        const float *xr = x + r * C;
        float *yr = out + r * C;
        double sum = 0.0;
        for (int i = 0; i < C; i++) sum += (double)xr[i];
        float mean = (float)(sum / C);
        double sum2 = 0.0;
        for (int i = 0; i < C; i++) {
            float v = xr[i] - mean;
            yr[i] = v;
            sum2 += (double)(v * v);
        }
        float variance = (float)(sum2 / C);
        const float scale = 1.0f / sqrtf(variance + eps);
        for (int i = 0; i < C; i++) yr[i] = (yr[i] * scale) * g[i] + b[i];
    }
}

/* dense: y[t][o] = sum_c x[t][c] * kernel[c][o] + bias[o]; kernel file layout (in,out) (App. B);
 * main.cpp:1022-1035 etc. -- F32 x F32 mul_mat then broadcast add. */
static void linear(const float *x, size_t T, int Cin, const mvo_tensor *kernel, const mvo_tensor *bias, float *y) {
    const int Cout = kernel->dims[1];
    if (kernel->dims[0] != Cin) { fprintf(stderr, "mvo: linear in %d != %d\n", kernel->dims[0], Cin); abort(); }
    mvo_gemm((int)T, Cout, Cin, x, Cin, kernel->data, Cout, y, Cout);
    for (size_t t = 0; t < T; t++)
        for (int o = 0; o < Cout; o++) y[t * Cout + o] = y[t * Cout + o] + bias->data[o];
}

/* ---- classifier head (not in main.cpp, which stops at the feature map :645): TFMobileViTForImageClassification's
 * `classifier/{kernel,bias}:0`, kernel (in, out) like every dense kernel of the file; logits = pooled . kernel + bias. ---- */
static const mvo_tensor *find_suffix(const mvo_model *m, const char *suffix) {
    const size_t ls = strlen(suffix);
    for (int i = 0; i < m->n_tensors; i++) {
        const size_t ln = strlen(m->t[i].name);
        if (ln >= ls && strcmp(m->t[i].name + ln - ls, suffix) == 0) return &m->t[i];
    }
    return NULL;
}
int mvo_num_classes(void *model) {
    const mvo_tensor *k = find_suffix((mvo_model *)model, "classifier/kernel:0");
    return k ? k->dims[1] : 0;
}
int mvo_classify(void *model, const float *pooled, int N, float *logits) {
    const mvo_model *m = (mvo_model *)model;
    const mvo_tensor *k = find_suffix(m, "classifier/kernel:0"), *b = find_suffix(m, "classifier/bias:0");
    if (!k || !b) return 2;
    linear(pooled, (size_t)N, k->dims[0], k, b, logits);
    return 0;
}

/* ---- sam_image_preprocess (main.cpp:538-601) for an H x W target: longer side fills the target, bilinear on the u8 source,
 * rounded back to u8, then (v - 0) / 255.  Deliberate deviation: rows are written with stride W, not nx3,
 * and the area outside the resized image is zero.  fp-contract is off so that no mul+add pair becomes an fma: the reference
 * Makefile builds without -mfma, and the CUDA kernel uses explicit _rn intrinsics -> bit-identical results. ---- */
__attribute__((optimize("fp-contract=off")))
void mvo_preprocess_u8(const uint8_t *img, int ny, int nx, int H, int W, float *out) {
    const float scale = H == W ? (float)(nx > ny ? nx : ny) * 1.0f / (float)W                      /* main.cpp:550 */
                               : fmaxf((float)nx / (float)W, (float)ny / (float)H);
    int nx3 = (int)(nx / scale + 0.5f), ny3 = (int)(ny / scale + 0.5f);                              /* :554-555 */
    if (nx3 > W) nx3 = W;
    if (ny3 > H) ny3 = H;
    memset(out, 0, (size_t)H * W * 3 * sizeof(float));
    for (int y = 0; y < ny3; y++)
        for (int x = 0; x < nx3; x++)
            for (int c = 0; c < 3; c++) {
                const float sx = (x + 0.5f) * scale - 0.5f, sy = (y + 0.5f) * scale - 0.5f;          /* :566-567 */
                int x0 = (int)floorf(sx), y0 = (int)floorf(sy);
                if (x0 < 0) x0 = 0;
                if (y0 < 0) y0 = 0;
                if (x0 > nx - 1) x0 = nx - 1;
                if (y0 > ny - 1) y0 = ny - 1;
                const int x1 = x0 + 1 < nx - 1 ? x0 + 1 : nx - 1, y1 = y0 + 1 < ny - 1 ? y0 + 1 : ny - 1; /* :572-573 */
                const float dx = sx - x0, dy = sy - y0;
                const float v00 = img[3 * (y0 * nx + x0) + c], v01 = img[3 * (y0 * nx + x1) + c];
                const float v10 = img[3 * (y1 * nx + x0) + c], v11 = img[3 * (y1 * nx + x1) + c];
                const float v0 = v00 * (1.0f - dx) + v01 * dx, v1 = v10 * (1.0f - dx) + v11 * dx;    /* :588-589 */
                const float v  = v0 * (1.0f - dy) + v1 * dy;
                const uint8_t v2 = (uint8_t)fminf(fmaxf(roundf(v), 0.0f), 255.0f);                   /* :593 */
                out[3 * ((size_t)y * W + x) + c] = ((float)v2 - 0.0f) / 255.0f;                       /* :596, stride W */
            }
}

/* softmax over ne0: [ggml] ggml_compute_forward_soft_max_f32 -- max, expf, sum in double, scale 1/sum */
void mvo_softmax_rows(float *x, int n, size_t rows) {
    for (size_t r = 0; r < rows; r++) {
        float *p = x + r * n;
        float mx = -INFINITY;
        for (int i = 0; i < n; i++) mx = p[i] > mx ? p[i] : mx;
        double sum = 0.0;
        for (int i = 0; i < n; i++) {
            float v = exp_op(p[i] - mx);
            p[i] = v;
            sum += (double)v;
        }
        float inv = (float)(1.0 / sum);
        for (int i = 0; i < n; i++) p[i] *= inv;
    }
}

/* mobilevit_transformer_layer::forward, main.cpp:988-1172.  x: [B][L][C] (ggml ne=(C,L,B)), in place. */
static void transformer_layer(const mvo_tlayer *T, float *x, int C, int L, int B, int heads, float eps) {
    const size_t rows = (size_t)B * L;
    const int d = C / heads;
    const float scale = sqrtf(1.0f * d); /* main.cpp:999 -- used as a divisor (:1076) */
    float *ln = (float *)malloc(rows * C * sizeof(float));
    float *q = (float *)malloc(rows * C * sizeof(float));
    float *k = (float *)malloc(rows * C * sizeof(float));
    float *v = (float *)malloc(rows * C * sizeof(float));
    float *ctx = (float *)malloc(rows * C * sizeof(float));
    float *sc = (float *)malloc((size_t)L * L * sizeof(float));
    mvo_layernorm(x, C, rows, T->lb_g->data, T->lb_b->data, eps, ln); /* :1002-1019 */
    linear(ln, rows, C, T->kk, T->kb, k);                             /* :1022-1036 */
    linear(ln, rows, C, T->vk, T->vb, v);                             /* :1039-1053 */
    linear(ln, rows, C, T->qk, T->qb, q);                             /* :1056-1070 */
    for (int b = 0; b < B; b++)
        for (int h = 0; h < heads; h++) {
            const float *qb = q + (size_t)b * L * C + h * d;
            const float *kb = k + (size_t)b * L * C + h * d;
            const float *vb = v + (size_t)b * L * C + h * d;
            /* scores[q][k] = (Q.K)/scale  (:1073-1077, ne0 = key index) */
            for (int i = 0; i < L; i++)
                for (int j = 0; j < L; j++) {
                    float s = 0.f;
                    for (int e = 0; e < d; e++) s += kb[(size_t)j * C + e] * qb[(size_t)i * C + e];
                    sc[(size_t)i * L + j] = s / scale;
                }
            mvo_softmax_rows(sc, L, (size_t)L); /* :1079 */
            /* context[q][e] = sum_k P[q][k] V[k][e] (:1082-1086), stored back head-interleaved (:1088-1093) */
            for (int i = 0; i < L; i++)
                for (int e = 0; e < d; e++) {
                    float s = 0.f;
                    for (int j = 0; j < L; j++) s += vb[(size_t)j * C + e] * sc[(size_t)i * L + j];
                    ctx[((size_t)b * L + i) * C + h * d + e] = s;
                }
        }
    linear(ctx, rows, C, T->ok, T->ob, q);                       /* :1095-1108 (q reused as scratch) */
    for (size_t i = 0; i < rows * C; i++) x[i] = x[i] + q[i];     /* :1111 hidden + attention_output */
    mvo_layernorm(x, C, rows, T->la_g->data, T->la_b->data, eps, ln); /* :1114-1131 */
    const int F = T->ik->dims[1];
    float *mid = (float *)malloc(rows * F * sizeof(float));
    linear(ln, rows, C, T->ik, T->ib, mid);                      /* :1134-1147 */
    for (size_t i = 0; i < rows * F; i++) mid[i] = silu_op(mid[i]); /* :1148 */
    linear(mid, rows, F, T->dk, T->db, q);                       /* :1151-1163 */
    for (size_t i = 0; i < rows * C; i++) x[i] = q[i] + x[i];     /* :1165-1169 */
    free(mid); free(sc); free(ctx); free(v); free(k); free(q); free(ln);
}

/* mobile_vit_layer::forward, main.cpp:1174-1223 */
static float *vit_layer(const mvo_model *m, const mvo_vit *L, const float *x, int C, int H, int W, int flags, int *oC,
                        int *oH, int *oW) {
    const int ps = 2;       /* main.cpp:38 */
    const float eps = 1e-5f; /* main.cpp:51 */
    int c0, h0, w0, c1, h1, w1, c2, h2, w2;
    float *res = inverted_residual(&L->down, x, C, H, W, flags, &c0, &h0, &w0); /* :1177 */
    float *f = conv_layer(&L->kxk, res, c0, h0, w0, 1, 1, 1, 0, flags, &c1, &h1, &w1); /* :1182 */
    float *g = conv_layer(&L->c1x1, f, c1, h1, w1, 1, 0, 0, 0, flags, &c2, &h2, &w2);  /* :1183 */
    free(f);
    if (h2 % ps || w2 % ps) { fprintf(stderr, "mvo: feature map not divisible by patch size\n"); abort(); } /* :729 */
    const int Lp = (h2 / ps) * (w2 / ps), PA = ps * ps;
    size_t n = (size_t)c2 * h2 * w2;
    float *tok = (float *)malloc(n * sizeof(float));
    mvo_unfold(g, c2, h2, w2, ps, tok); /* :1186 -> [PA][Lp][C] */
    for (int i = 0; i < L->n_tlayers; i++) transformer_layer(&L->tl[i], tok, c2, Lp, PA, m->num_heads, eps); /* :1189 */
    mvo_layernorm(tok, c2, (size_t)PA * Lp, L->ln_g->data, L->ln_b->data, eps, g); /* :1192-1209 (g reused) */
    mvo_fold(g, c2, Lp, ps, tok);                                                   /* :1212 -> [C][H][W] */
    free(g);
    int c3, h3, w3;
    float *pr = conv_layer(&L->proj, tok, c2, h2, w2, 1, 1, 1, 0, flags, &c3, &h3, &w3); /* :1215 */
    free(tok);
    /* ggml_concat(residual, features) along dim 2 = channels (:1219) */
    float *cat = (float *)malloc((size_t)(c0 + c3) * h3 * w3 * sizeof(float));
    memcpy(cat, res, (size_t)c0 * h0 * w0 * sizeof(float));
    memcpy(cat + (size_t)c0 * h0 * w0, pr, (size_t)c3 * h3 * w3 * sizeof(float));
    free(res);
    free(pr);
    float *out = conv_layer(&L->fusion, cat, c0 + c3, h3, w3, 1, 1, 1, 0, flags, oC, oH, oW); /* :1217-1221 */
    free(cat);
    return out;
}

/* mobilevit_model::extract_features, main.cpp:604-646, for one image.
 *   img_hwc : H*W*3 floats in [0,1], HWC (sam_image_f32, main.cpp:22-27); transposed to CHW (:627-634)
 *   feat    : [OC][H/32][W/32] == ggml ne=(W/32,H/32,OC,1)     (may be NULL)
 *   pooled  : [OC] mean over the spatial map (build addition)   (may be NULL)
 *   stage_out : optional 7 malloc'd-by-caller buffers receiving stem, layer1..5, exp outputs (CHW) */
int mvo_forward(void *model, const float *img_hwc, int H, int W, float *feat, float *pooled, int flags,
                float **stage_out) {
    const mvo_model *m = (const mvo_model *)model;
    if ((flags & MVO_FLAG_LEGACY_F16_TABLES) && !g_tab_silu) build_legacy_tables();
    g_legacy_tables = (flags & MVO_FLAG_LEGACY_F16_TABLES) != 0; /* workers of one batch all write the same value */
    float *chw = (float *)malloc((size_t)3 * H * W * sizeof(float));
    for (int k = 0; k < 3; k++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) chw[((size_t)k * H + y) * W + x] = img_hwc[((size_t)y * W + x) * 3 + k];
    int C, h, w, c2, h2, w2;
    int stage = 0;
#define TAP(ptr, cc, hh, ww) do { if (stage_out && stage_out[stage]) memcpy(stage_out[stage], ptr, (size_t)(cc) * (hh) * (ww) * sizeof(float)); stage++; } while (0)
    float *cur = conv_layer(&m->stem, chw, 3, H, W, 2, 1, 1, 0, flags, &C, &h, &w); /* :618 */
    free(chw);
    TAP(cur, C, h, w);
    for (int i = 0; i < m->n_l1; i++) { /* main.cpp:97-105 */
        float *nx = inverted_residual(&m->l1[i], cur, C, h, w, flags, &c2, &h2, &w2);
        free(cur); cur = nx; C = c2; h = h2; w = w2;
    }
    TAP(cur, C, h, w);
    for (int i = 0; i < m->n_l2; i++) {
        float *nx = inverted_residual(&m->l2[i], cur, C, h, w, flags, &c2, &h2, &w2);
        free(cur); cur = nx; C = c2; h = h2; w = w2;
    }
    TAP(cur, C, h, w);
    for (int v = 0; v < 3; v++) { /* main.cpp:187-199 */
        float *nx = vit_layer(m, &m->vit[v], cur, C, h, w, flags, &c2, &h2, &w2);
        free(cur); cur = nx; C = c2; h = h2; w = w2;
        TAP(cur, C, h, w);
    }
    float *out = conv_layer(&m->exp1x1, cur, C, h, w, 1, 1, 1, 0, flags, &c2, &h2, &w2); /* :624 */
    free(cur);
    TAP(out, c2, h2, w2);
#undef TAP
    if (feat) memcpy(feat, out, (size_t)c2 * h2 * w2 * sizeof(float));
    if (pooled)
        for (int c = 0; c < c2; c++) {
            float s = 0.f;
            for (int i = 0; i < h2 * w2; i++) s += out[(size_t)c * h2 * w2 + i];
            pooled[c] = s / (float)(h2 * w2);
        }
    free(out);
    return c2;
}

/* Batch = loop of batch-1 forwards spread over pthreads (images are independent);
 * n_threads <= 0 -> all online cores.  Returns wall seconds. */
typedef struct {
    void *model; const float *imgs; int N, H, W; float *feat, *pooled; int flags; int next; size_t fstride; int OC;
} mvo_job;

static void *mvo_worker(void *arg) {
    mvo_job *j = (mvo_job *)arg;
    for (;;) {
        int i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->N) break;
        mvo_forward(j->model, j->imgs + (size_t)i * j->H * j->W * 3, j->H, j->W,
                    j->feat ? j->feat + i * j->fstride : NULL, j->pooled ? j->pooled + (size_t)i * j->OC : NULL,
                    j->flags, NULL);
    }
    return NULL;
}

int mvo_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

double mvo_forward_batch(void *model, const float *imgs_hwc, int N, int H, int W, float *feat, float *pooled,
                         int flags, int n_threads) {
    mvo_job j = {model, imgs_hwc, N, H, W, feat, pooled, flags, 0, 0, 0};
    if (flags & MVO_FLAG_LEGACY_F16_TABLES) build_legacy_tables();
    g_legacy_tables = (flags & MVO_FLAG_LEGACY_F16_TABLES) != 0;  /* before the workers start */
    j.OC = mvo_out_channels(model);
    j.fstride = (size_t)j.OC * (H / 32) * (W / 32);
    if (n_threads <= 0) n_threads = mvo_max_threads();
    if (n_threads > N) n_threads = N;
    if (n_threads > 256) n_threads = 256;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_t th[256];
    for (int t = 1; t < n_threads; t++) pthread_create(&th[t], NULL, mvo_worker, &j);
    mvo_worker(&j);
    for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
