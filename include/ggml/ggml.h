/*
 * ggml.h -- the drop-in boundary of libggml_b200.
 *
 * This header declares the subset of the ggml public C API that the reference programs call
 * (/root/reference/mobilevit/main.cpp includes "ggml/ggml.h" at :1,
 *  /root/reference/rnn_text_gen/rnn_text_generation.cpp includes <ggml.h> at :1).
 * Upstream ggml is NOT vendored by the reference (mobilevit/README.md:10-14 tells the user to
 * `git clone` it), so the signatures below restate upstream ggml's early-2024 public API as the two
 * programs use it.  Every symbol cites the reference call site(s) it serves.
 *
 * Behind this surface there is no CPU interpreter: graph construction only records nodes;
 * ggml_graph_compute_with_ctx() lowers the recorded graph to hand-written sm_100a CUDA kernels
 * (see DESIGN.md).  If no CUDA device is usable the compute call aborts -- there is no CPU fallback.
 *
 * Plain C ABI: extern "C", pointers and sizes only.
 */
#ifndef GGML_B200_GGML_H
#define GGML_B200_GGML_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GGML_MAX_DIMS       4
#define GGML_MAX_SRC        4
#define GGML_MAX_NAME       64
#define GGML_MAX_OP_PARAMS  64
#define GGML_DEFAULT_GRAPH_SIZE 2048   /* upstream default; the reference graph is ~1.4k nodes */
#define GGML_B200_STATIC_GRAPH_NODES 512 /* capacity of a by-value ggml_cgraph (rnn.cpp:149,289) */

/* main.cpp:729,730,981 -- abort-on-failure convention (prints file:line, abort()). */
#define GGML_ASSERT(x)                                                                   \
    do {                                                                                 \
        if (!(x)) {                                                                      \
            fprintf(stderr, "GGML_ASSERT: %s:%d: %s\n", __FILE__, __LINE__, #x);         \
            abort();                                                                     \
        }                                                                                \
    } while (0)

typedef uint16_t ggml_fp16_t; /* main.cpp:929 */

enum ggml_type {            /* main.cpp:612,907-916; rnn.cpp:42,284 */
    GGML_TYPE_F32 = 0,
    GGML_TYPE_F16 = 1,
    GGML_TYPE_I32 = 26,
    GGML_TYPE_COUNT
};

enum ggml_op {
    GGML_OP_NONE = 0,
    GGML_OP_ADD, GGML_OP_SUB, GGML_OP_MUL, GGML_OP_DIV,
    GGML_OP_SQRT, GGML_OP_SILU, GGML_OP_TANH,
    GGML_OP_NORM, GGML_OP_SOFT_MAX,
    GGML_OP_MUL_MAT,
    GGML_OP_REPEAT, GGML_OP_CONCAT, GGML_OP_GET_ROWS,
    GGML_OP_CONT,
    GGML_OP_RESHAPE, GGML_OP_VIEW, GGML_OP_PERMUTE, GGML_OP_TRANSPOSE,
    GGML_OP_CONV_2D, GGML_OP_CONV_DEPTHWISE_2D,
    GGML_OP_POOL_MEAN_HW,   /* build addition: global average pool */
    GGML_OP_ARGMAX,         /* upstream ggml_argmax: index of the row maximum, I32 (batched greedy decoding, rnn.cpp:74-77,312) */
    GGML_OP_COUNT
};

enum ggml_tensor_flag { GGML_TENSOR_FLAG_INPUT = 1, GGML_TENSOR_FLAG_OUTPUT = 2, GGML_TENSOR_FLAG_PARAM = 4 };

struct ggml_context;        /* opaque arena (main.cpp:607,658) */

/* Public tensor record.  The reference reads type/ne (main.cpp:724-726,780-783,977-979), writes nb
 * (rnn.cpp:212,226), reads/writes data (main.cpp:931,934; rnn.cpp:124-147,306-307) and reads n_dims
 * (rnn.cpp:35), so those fields keep upstream's names and meaning.  `data` is a HOST pointer: it is
 * allocated eagerly for leaf tensors; for op results it is NULL until a compute call has produced the
 * tensor and it was a graph output (or small enough to be mirrored, see DESIGN.md "host shadows"). */
struct ggml_tensor {
    enum ggml_type type;
    int            n_dims;
    int64_t        ne[GGML_MAX_DIMS];
    size_t         nb[GGML_MAX_DIMS];
    enum ggml_op   op;
    int32_t        op_params[GGML_MAX_OP_PARAMS / sizeof(int32_t)];
    int32_t        flags;
    struct ggml_tensor * src[GGML_MAX_SRC];
    struct ggml_tensor * view_src;
    size_t         view_offs;
    void *         data;
    char           name[GGML_MAX_NAME];
    void *         extra;   /* libggml_b200 private */
    struct ggml_context * ctx; /* owning context (libggml_b200 private) */
};

/* main.cpp:608,636; rnn.cpp:149 (`ggml_cgraph gf = {}` on the stack) and :289 (returned by value):
 * the struct must be complete and value-copyable. */
struct ggml_cgraph {
    int size;
    int n_nodes;
    int n_leafs;
    struct ggml_tensor ** nodes;
    struct ggml_tensor ** leafs;
    void * plan;            /* libggml_b200 private: compiled device plan, built lazily */
    struct ggml_tensor * static_nodes[GGML_B200_STATIC_GRAPH_NODES];
    struct ggml_tensor * static_leafs[GGML_B200_STATIC_GRAPH_NODES];
};

struct ggml_init_params {   /* main.cpp:605,656; rnn.cpp:98-102,271-275 */
    size_t mem_size;
    void * mem_buffer;
    bool   no_alloc;
};

/* ---- context (main.cpp:607,658,699) ---- */
struct ggml_context * ggml_init(struct ggml_init_params params);
void                  ggml_free(struct ggml_context * ctx);
size_t                ggml_used_mem(const struct ggml_context * ctx);

/* ---- timers (main.cpp:639-641,651,689-698) ---- */
void    ggml_time_init(void);
int64_t ggml_time_ms(void);
int64_t ggml_time_us(void);

/* ---- tensor creation (main.cpp:612,833,907-916,1076; rnn.cpp:42,108-115,244,283-284) ---- */
struct ggml_tensor * ggml_new_tensor_1d(struct ggml_context * ctx, enum ggml_type type, int64_t ne0);
struct ggml_tensor * ggml_new_tensor_2d(struct ggml_context * ctx, enum ggml_type type, int64_t ne0, int64_t ne1);
struct ggml_tensor * ggml_new_tensor_3d(struct ggml_context * ctx, enum ggml_type type, int64_t ne0, int64_t ne1, int64_t ne2);
struct ggml_tensor * ggml_new_tensor_4d(struct ggml_context * ctx, enum ggml_type type, int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3);
struct ggml_tensor * ggml_new_f32(struct ggml_context * ctx, float value);
struct ggml_tensor * ggml_set_name(struct ggml_tensor * tensor, const char * name);   /* main.cpp:614 */
void                 ggml_set_input(struct ggml_tensor * tensor);                    /* main.cpp:615 */
void                 ggml_set_output(struct ggml_tensor * tensor);
void                 ggml_set_param(struct ggml_context * ctx, struct ggml_tensor * tensor); /* rnn.cpp:150-152,286-287 */

/* ---- accessors (main.cpp:627,931-934,952,1230; rnn.cpp:27,44,75,303) ---- */
void *  ggml_get_data(const struct ggml_tensor * tensor);
float * ggml_get_data_f32(const struct ggml_tensor * tensor);
size_t  ggml_nbytes(const struct ggml_tensor * tensor);
int64_t ggml_nelements(const struct ggml_tensor * tensor);
int     ggml_n_dims(const struct ggml_tensor * tensor);
size_t  ggml_type_size(enum ggml_type type);
bool    ggml_is_contiguous(const struct ggml_tensor * tensor);
void    ggml_set_i32_1d(const struct ggml_tensor * tensor, int i, int32_t value);
int32_t ggml_get_i32_1d(const struct ggml_tensor * tensor, int i);
void    ggml_set_f32_1d(const struct ggml_tensor * tensor, int i, float value);
float   ggml_get_f32_1d(const struct ggml_tensor * tensor, int i);

/* ---- fp16 (main.cpp:930) ---- */
float       ggml_fp16_to_fp32(ggml_fp16_t x);
ggml_fp16_t ggml_fp32_to_fp16(float x);
void        ggml_fp16_to_fp32_row(const ggml_fp16_t * x, float * y, int n);
void        ggml_fp32_to_fp16_row(const float * x, ggml_fp16_t * y, int n);

/* ---- binary ops with broadcast of b (main.cpp:810,821,826,837,842,867,1002-1018,1073,1111,1165) ---- */
struct ggml_tensor * ggml_add(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b);
struct ggml_tensor * ggml_sub(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b);
struct ggml_tensor * ggml_mul(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b);
struct ggml_tensor * ggml_div(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b);

/* ---- unary ops (main.cpp:824,849,1079,1148; rnn.cpp:53,236) ---- */
struct ggml_tensor * ggml_sqrt(struct ggml_context * ctx, struct ggml_tensor * a);
struct ggml_tensor * ggml_silu(struct ggml_context * ctx, struct ggml_tensor * a);
struct ggml_tensor * ggml_tanh(struct ggml_context * ctx, struct ggml_tensor * a);
struct ggml_tensor * ggml_soft_max(struct ggml_context * ctx, struct ggml_tensor * a);       /* over ne0 */
struct ggml_tensor * ggml_norm(struct ggml_context * ctx, struct ggml_tensor * a, float eps); /* over ne0, no affine (main.cpp:1006,1118,1196) */

/* ---- matmul (main.cpp:1022,1039,1056,1075,1082,1095,1134,1151; rnn.cpp:204,219,253) ----
 * a: [K, M, b2, b3]   b: [K, N, c2, c3]   result F32 [M, N, c2, c3]; a is broadcast over dims 2,3.
 * If a is F16, b is rounded to F16 first and products accumulate in f32 (upstream vec_dot_f16). */
struct ggml_tensor * ggml_mul_mat(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b);

/* ---- data movement / views (main.cpp:721-768,790-805,975-986,1009-1093,1136,1152,1219; rnn.cpp:47,129-143,200-250) ---- */
struct ggml_tensor * ggml_repeat(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b);
struct ggml_tensor * ggml_concat(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b); /* 2-arg form: along dim 2 */
struct ggml_tensor * ggml_get_rows(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b);
struct ggml_tensor * ggml_cont(struct ggml_context * ctx, struct ggml_tensor * a);
struct ggml_tensor * ggml_cont_4d(struct ggml_context * ctx, struct ggml_tensor * a, int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3);
struct ggml_tensor * ggml_reshape_2d(struct ggml_context * ctx, struct ggml_tensor * a, int64_t ne0, int64_t ne1);
struct ggml_tensor * ggml_reshape_3d(struct ggml_context * ctx, struct ggml_tensor * a, int64_t ne0, int64_t ne1, int64_t ne2);
struct ggml_tensor * ggml_reshape_4d(struct ggml_context * ctx, struct ggml_tensor * a, int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3);
struct ggml_tensor * ggml_permute(struct ggml_context * ctx, struct ggml_tensor * a, int axis0, int axis1, int axis2, int axis3); /* result.ne[axis_i] = a.ne[i] */
struct ggml_tensor * ggml_transpose(struct ggml_context * ctx, struct ggml_tensor * a);
/* upstream ggml_view_2d: ne0 x ne1 window into a, row stride nb1 bytes, starting `offset` bytes into a (the batched GRU
 * program slices the z / r / h gate blocks with it instead of rnn.cpp's get_rows-on-a-fake-transpose trick, :41-49,212) */
struct ggml_tensor * ggml_view_2d(struct ggml_context * ctx, struct ggml_tensor * a, int64_t ne0, int64_t ne1, size_t nb1, size_t offset);
/* upstream ggml_argmax: a [n, rows] F32 -> I32 [rows] (device-side replacement of rnn.cpp:74-77 argmax_1d) */
struct ggml_tensor * ggml_argmax(struct ggml_context * ctx, struct ggml_tensor * a);

/* ---- convolution (main.cpp:788,798) ----
 * kernel a: [KW, KH, IC, OC] F16 (depthwise: [KW, KH, 1, C]); input b: [W, H, C, N] F32; result F32
 * [OW, OH, OC, N].  Numerics follow upstream im2col(F16)+mul_mat: activations are rounded to F16,
 * products accumulate in f32. */
struct ggml_tensor * ggml_conv_2d(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b,
                                  int s0, int s1, int p0, int p1, int d0, int d1);
struct ggml_tensor * ggml_conv_depthwise_2d(struct ggml_context * ctx, struct ggml_tensor * a, struct ggml_tensor * b,
                                            int s0, int s1, int p0, int p1, int d0, int d1);

/* ---- graph (main.cpp:608,636,640; rnn.cpp:149,154-158,289,311) ---- */
struct ggml_cgraph * ggml_new_graph(struct ggml_context * ctx);
void                 ggml_build_forward_expand(struct ggml_cgraph * cgraph, struct ggml_tensor * tensor);
struct ggml_cgraph   ggml_build_forward(struct ggml_tensor * tensor);    /* old by-value API, rnn.cpp:289 */
void                 ggml_graph_compute_with_ctx(struct ggml_context * ctx, struct ggml_cgraph * cgraph, int n_threads);
void                 ggml_graph_release_plan(struct ggml_cgraph * cgraph); /* frees the cached device plan */

/* =====================================================================================================
 * libggml_b200 extensions (not in upstream ggml).  All optional: an unmodified caller never needs them.
 * ===================================================================================================== */

/* build addition: mean over ne0 x ne1 -> [1,1,C,N] (the "pooled logits"). */
struct ggml_tensor * ggml_b200_pool_mean_hw(struct ggml_context * ctx, struct ggml_tensor * a);

/* Execution mode of ggml_graph_compute_with_ctx:
 *   GGML_B200_MODE_FAST  (default) pattern-matched fused plan (NHWC f16 activations, tcgen05 GEMMs);
 *                        falls back (with a one-line notice) to EXACT for graphs it cannot match.
 *   GGML_B200_MODE_EXACT one simple f32-accurate CUDA kernel per ggml node, ggml layouts and rounding
 *                        points -- the "TF32/f32 validation mode" of BASELINE.json's north_star.
 *   GGML_B200_MODE_EXACT_F32  EXACT without the f16 rounding of conv ACTIVATIONS (weights keep their loaded f16 values):
 *                        a smooth function on both sides, so it can be compared with the oracle's matching mode to
 *                        max-abs 1e-3 -- the f16 rounding flips of the ggml semantics otherwise put a ~1e-3 relative
 *                        noise floor under any two implementations (DESIGN.md "the f16 noise floor").
 * Also settable with the environment variable GGML_B200_MODE=fast|exact|exact_f32. */
enum ggml_b200_mode { GGML_B200_MODE_FAST = 0, GGML_B200_MODE_EXACT = 1, GGML_B200_MODE_EXACT_F32 = 2 };
void ggml_b200_set_mode(enum ggml_b200_mode mode);
int  ggml_b200_get_mode(void);

/* Device selection and stream.  `stream` is a cudaStream_t passed as void* (NULL = the library's own). */
int    ggml_b200_device_count(void);
void   ggml_b200_set_device(int device);
void   ggml_b200_set_stream(void * stream);
void * ggml_b200_get_stream(void);
void   ggml_b200_synchronize(void);

/* Keep a leaf / an output on the device: bind a leaf to caller-owned device memory (skips the H2D copy
 * in compute), or ask where a graph output lives after compute (skips nothing; the D2H copy is skipped
 * with ggml_b200_graph_set_download(false)). */
void   ggml_b200_tensor_set_device_data(struct ggml_tensor * leaf, void * device_ptr);
void * ggml_b200_tensor_get_device_data(struct ggml_cgraph * cgraph, struct ggml_tensor * tensor);
void   ggml_b200_graph_set_transfers(struct ggml_cgraph * cgraph, bool upload_inputs, bool download_outputs);

/* Page-locked host memory to be passed as ggml_init_params.mem_buffer (returns NULL without a device). */
void * ggml_b200_host_malloc(size_t bytes);
void   ggml_b200_host_free(void * p);

/* Pipelined submission (throughput serving): give a graph a private stream, submit without waiting, wait later.
 * With two copies of a forward graph (own input leaf and output shadows each) the H2D copy of batch i+1 and the D2H
 * copy of batch i-1 overlap the kernels of batch i.  ggml_graph_compute_with_ctx itself stays synchronous. */
void ggml_b200_graph_use_private_stream(struct ggml_cgraph * cgraph);
void ggml_b200_graph_compute_async(struct ggml_context * ctx, struct ggml_cgraph * cgraph);
void ggml_b200_graph_wait(struct ggml_cgraph * cgraph);
/* Concurrent lanes: n prepared graphs (sub-batches of one request) run at the same time on private streams.  begin: every lane's
 * stream waits for the owner stream (on_current_stream != 0: the library's current stream; 0: lane 0's own stream, for pipelined
 * slots); the caller then enqueues per-lane work (ggml_b200_graph_upload_u8_images, ggml_b200_graph_compute_async); end: the owner
 * stream waits for every lane. */
void ggml_b200_graph_group_begin(struct ggml_cgraph ** cgraphs, int n, int on_current_stream);
void ggml_b200_graph_group_end(struct ggml_cgraph ** cgraphs, int n, int on_current_stream);

/* Device-side feedback for autoregressive loops: after every compute of `cgraph`, copy node `src` into
 * the device buffer of leaf `dst` (same byte size).  Replaces rnn.cpp:303-310 (host writes the next token id and memcpy's
 * the new state back into the input leaf).  Register before the first compute. */
void ggml_b200_graph_add_feedback(struct ggml_cgraph * cgraph, struct ggml_tensor * src, struct ggml_tensor * dst);

/* Device-resident autoregressive loop (the host loop of rnn.cpp:293-313): run the graph `steps` times back to back
 * without a host round trip (the registered feedback copies carry ids/state between steps; inputs are uploaded for step 0
 * only), keep `record` of every step in a device history and copy it to host_dst (steps * nbytes(record)) once at the end.
 * Returns 0 on success. */
int ggml_b200_graph_compute_steps(struct ggml_context * ctx, struct ggml_cgraph * cgraph, int steps, struct ggml_tensor * record, void * host_dst);

/* Image preprocessing on the device (sam_image_preprocess + the HWC copy, main.cpp:538-601,627-634): enqueue on
 * the graph's stream the H2D copy of `n` raw u8 images [src_h][src_w][3] and the bilinear resize / u8 rounding / 1/255 kernel
 * that fills the device copy of the f32 input leaf `input` (ne = (3, W, H, n)).  Call ggml_b200_graph_prepare first and run
 * the compute with uploads disabled (ggml_b200_graph_set_transfers(g, false, ...)).  Returns 0 on success. */
int ggml_b200_graph_upload_u8_images(struct ggml_cgraph * cgraph, struct ggml_tensor * input, const uint8_t * host_u8, int n, int src_h,
                                     int src_w);

/* The same request when only the next compute will read the images: a FAST plan with the tensor-core stem keeps the quantised u8
 * image (3 bytes per pixel; images already H x W are copied straight in) and the stem stages its patch from it -- same bits as the
 * f32 route.  The f32 input leaf is not written.  Falls back to ggml_b200_graph_upload_u8_images for other plans. */
int ggml_b200_graph_upload_u8_images_fused(struct ggml_cgraph * cgraph, struct ggml_tensor * input, const uint8_t * host_u8, int n,
                                           int src_h, int src_w);

/* Synchronous device->host copy of any (contiguous) tensor of the graph's plan, e.g. an intermediate that is not a
 * declared output.  Returns 0 on success. */
int ggml_b200_tensor_download(struct ggml_cgraph * cgraph, struct ggml_tensor * tensor, void * host_dst);

/* Introspection of the compiled plan (bench.py's gpu_launches, DESIGN.md's memory-planner numbers). */
struct ggml_b200_plan_stats {
    int     mode;              /* enum ggml_b200_mode actually used */
    int     n_graph_nodes;     /* ggml nodes in the graph */
    int     n_launches;        /* kernel launches per compute */
    int     n_folded;          /* nodes constant-folded at plan time */
    int64_t arena_bytes;       /* device arena after liveness planning */
    int64_t naive_bytes;       /* sum of all intermediate tensors (what ggml's arena would need) */
    int64_t weight_bytes;      /* device-resident constants */
    int     used_cuda_graph;
};
void ggml_b200_graph_plan_stats(struct ggml_cgraph * cgraph, struct ggml_b200_plan_stats * out);
/* Per-launch device times of the plan as a JSON array [{kernel, what, ms, flops, bytes}...] (direct launches
 * timed with CUDA events).  Returns 0, -1 if there is no plan, or the needed capacity if `cap` is too small. */
int ggml_b200_graph_profile_json(struct ggml_cgraph * cgraph, int reps, char * buf, size_t cap);
/* Build (or fetch) the plan without running it. */
void ggml_b200_graph_prepare(struct ggml_context * ctx, struct ggml_cgraph * cgraph);

const char * ggml_b200_version(void);

/* ---- single-kernel test entry points (host buffers in/out; used by tests/ only) ----------------------
 * f16 arrays are passed as uint16_t bit patterns.  Return 0 on success, 1 if the shape is not supported. */
/* out[M,N] = epilogue(A[M,K] * B[N,K]^T): tcgen05 GEMM (K1).  scale/shift/res32/out32/out16 may be NULL. */
int ggml_b200_debug_gemm(const uint16_t * A, const uint16_t * B, int M, int N, int K, const float * scale,
                         const float * shift, int act, const float * res32, float * out32, uint16_t * out16);
/* 3x3 stride-1 pad-1 conv over NHWC f16 (x0: C0 ch, optional x1: C1 ch = fused concat), Wt [OC][3][3][C0+C1] */
int ggml_b200_debug_conv3x3(const uint16_t * x0, int C0, const uint16_t * x1, int C1, int n_img, int H, int W,
                            const uint16_t * Wt, int OC, const float * scale, const float * shift, int act,
                            float * out32);
/* K3 depthwise 3x3 (stride 1|2, pad 1) + scale/shift + SiLU (main.cpp:788,809-850); x [N,H,W,C], Wt [3][3][C];
 * variant 0 = TMA kernel of the fused plan, 1 = register-window fallback */
int ggml_b200_debug_dwconv(const uint16_t * x, int N, int H, int W, int C, int stride, const uint16_t * Wt, const float * scale,
                           const float * shift, int act, int variant, uint16_t * out16);
/* K2 stem 3x3/s2 over f32 images (chw = 0: [N,H,W,3], chw = 1: [N,3,H,W] as main.cpp:627-634 writes it), Wt [OC][3][3][3] */
int ggml_b200_debug_stem(const float * x, int chw, int N, int H, int W, const uint16_t * Wt, int OC, const float * scale, const float * shift,
                         int act, uint16_t * out16, float * out32);
/* K7 attention core (main.cpp:1073-1086) on a head-padded qkv buffer [N*H*W][3][heads][DP], DP = ..._attention_dp(C/heads) */
int ggml_b200_debug_attention_dp(int d);
int ggml_b200_debug_attention(const uint16_t * qkv, int N, int H, int W, int C, int heads, uint16_t * out16);
float ggml_b200_debug_attention_time(const uint16_t * qkv, int N, int H, int W, int C, int heads, int reps);
/* LayerNorm folded around two GEMMs (main.cpp:1002-1019 + the following dense): producer x = A.B^T + shift0 with row statistics,
 * consumer y = act(LN(x).W^T + bias); Wf f32 [N][C]; x32 (optional) returns the producer's f32 output */
int ggml_b200_debug_gemm_ln(const uint16_t * A, const uint16_t * B, int M, int C, int K, const float * shift0, const float * gamma,
                            const float * beta, float eps, const float * Wf, const float * bias, int N, int act, float * x32, float * y32);
/* K4 fused inverted residual (inverted_residual_layer::forward, main.cpp:854-870): expand 1x1 (+BN+SiLU) -> depthwise 3x3 stride 1|2
 * (+BN+SiLU) -> reduce 1x1 (+BN) [+ f32 residual].  x [N,H,W,Cin], We [E][Cin], Wd [3][3][E], Wr [Cout][E] f16; s* / h* = folded BatchNorm
 * scale / shift.  Returns 2 when the shape is outside the fused kernel's envelope (the plan then runs the three separate kernels). */
int ggml_b200_debug_ir_fused(const uint16_t * x, int N, int H, int W, int Cin, int E, int Cout, int stride, const uint16_t * We,
                             const float * se, const float * he, const uint16_t * Wd, const float * sd, const float * hd, const uint16_t * Wr,
                             const float * sr, const float * hr, const float * res32, uint16_t * out16, float * out32);
/* timing probe: mean ms per launch of the fused block and (optionally) of the three separate kernels it replaces */
float ggml_b200_debug_ir_time(int N, int H, int W, int Cin, int E, int Cout, int stride, int with_res, int reps, float * unfused_ms);
/* K8 fused transformer stage (transformer_layer::forward, main.cpp:988-1172, x n_layers as mobile_vit_layer::forward loops it,
 * main.cpp:1196-1204) over the pixel-ordered f32 residual stream x [N,H,W,C]; sequences = pixels sharing a patch position (main.cpp:721-747).
 * params: n_layers x 16 host pointers (ln1_g ln1_b wq bq wk bk wv bv wo bo ln2_g ln2_b w1 b1 w2 b2), dense kernels f32 [in][out].
 * Returns 2 when the shape is outside the fused kernel's envelope.  reps > 0: *ms = mean ms per launch. */
int ggml_b200_debug_vit_stage(const float * x, int N, int H, int W, int C, int heads, int F, int n_layers, float eps,
                              const float * const * params, float * out32, uint16_t * out16, float * stats, int reps, float * ms);

#ifdef __cplusplus
}
#endif
#endif /* GGML_B200_GGML_H */
