// internal.h -- private declarations shared by the translation units of libggml_b200.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

#include "ggml/ggml.h"

#define B200_CHECK(call)                                                                              \
    do {                                                                                              \
        cudaError_t err__ = (call);                                                                   \
        if (err__ != cudaSuccess) {                                                                   \
            fprintf(stderr, "libggml_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(err__),    \
                    __FILE__, __LINE__, cudaGetErrorString(err__));                                   \
            abort();                                                                                  \
        }                                                                                             \
    } while (0)

#define B200_ABORT(...)                                    \
    do {                                                   \
        fprintf(stderr, "libggml_b200: " __VA_ARGS__);     \
        fprintf(stderr, " (%s:%d)\n", __FILE__, __LINE__); \
        abort();                                           \
    } while (0)

struct ggml_context {
    size_t mem_size;
    char * mem_buffer;
    bool   owned;
    bool   no_alloc;
    size_t used;
    int    n_objects;
};

namespace b200 {

// ---- process-wide runtime state ---------------------------------------------------------------------
struct Runtime {
    int          device          = 0;
    bool         initialised     = false;
    cudaStream_t own_stream      = nullptr;
    cudaStream_t user_stream     = nullptr;
    bool         use_user_stream = false;
    int          mode            = GGML_B200_MODE_FAST;
    int          sm_count        = 148;
    bool         verbose         = false;
    bool         use_cuda_graph  = true;
    cudaEvent_t  compute_chain   = nullptr;  // end of the most recent private-stream forward (pipelined slots run their kernels FIFO)
};
Runtime &    runtime();
void         ensure_device();  // aborts if no CUDA device is usable (there is no CPU fallback)
cudaStream_t current_stream();
void *       arena_alloc(ggml_context * ctx, size_t bytes, size_t align);  // host bump allocator

// ---- strided tensor view handed to kernels -------------------------------------------------------------
struct TView {
    void *  p;
    int64_t ne[4];
    int64_t nb[4];  // bytes, as in ggml
    int     type;   // ggml_type
};

// ---- liveness-based device arena planner (replaces ggml's "everything stays alive" CPU arena) ----------
struct ArenaPlanner {
    struct Block { int64_t off, size; };
    std::vector<Block> free_list;
    int64_t            extent = 0;
    static int64_t align_up(int64_t x) { return (x + 511) & ~int64_t(511); }
    int64_t alloc(int64_t bytes);
    void    release(int64_t off, int64_t bytes);
};

// Device residency of one ggml tensor inside a plan.
enum SlotKind { SLOT_NONE = 0, SLOT_ARENA, SLOT_CONST, SLOT_INPUT, SLOT_ALIAS, SLOT_EXTERNAL };
struct Slot {
    SlotKind kind   = SLOT_NONE;
    int64_t  offset = 0;  // arena offset (SLOT_ARENA)
    int64_t  bytes  = 0;
    void *   dptr   = nullptr;
    int      first_use = -1, last_use = -1;
};

struct Transfer {
    ggml_tensor * t;
    void *        dptr;
    size_t        bytes;
};

using Launch = std::function<void(cudaStream_t)>;

// What one launch does, for bench.py's roofline: algorithmic FLOPs and bytes (each operand read once, the
// result written once), and the ggml node / fused unit it implements.
struct LaunchMeta {
    const char * kernel = "";
    std::string  what;
    double       flops = 0, bytes = 0;  // bytes: what THIS plan moves (f32 side copies of the residual stream, residual reads, weights)
    double       bytes_min = 0;         // SURVEY 8(d) layer-wise minimum: input f16 once + output f16 once (+ weights); <= bytes
};

struct Plan {
    int                   mode = GGML_B200_MODE_EXACT;
    ggml_context *        ctx  = nullptr;
    int                   n_nodes_at_build = 0;
    uint64_t              graph_sig = 0;
    std::vector<Launch>   launches;
    std::vector<LaunchMeta> meta;
    std::unordered_map<const ggml_tensor *, Slot> slots;
    std::vector<void *>   owned_device;  // cudaMalloc'd blocks freed with the plan (const pool, arena, ...)
    char *                arena = nullptr;
    int64_t               arena_bytes = 0, naive_bytes = 0, weight_bytes = 0;
    int                   n_folded = 0;
    std::vector<Transfer> uploads, downloads;
    std::vector<void *>   pinned;        // host ranges registered with cudaHostRegister
    void *                host_mirror = nullptr;  // pinned host copies of small intermediates (tiny graphs only, see plan.cpp)
    std::vector<ggml_tensor *> mirrored;          // tensors whose ->data points into host_mirror (reset when the plan dies)
    bool                  upload_inputs = true, download_outputs = true;
    cudaGraphExec_t       graph_exec = nullptr;
    bool                  graph_failed = false;
    void *                history        = nullptr;  // per-step record of ggml_b200_graph_compute_steps (grow-only)
    size_t                history_bytes  = 0;
    void *                u8_stage       = nullptr;  // device staging of raw u8 images (ggml_b200_graph_upload_u8_images)
    size_t                u8_stage_bytes = 0;
    // u8 images handed straight to the stem (FAST plan, tensor-core stem): the quantised, resized images [N][H][W][3] and a device
    // word the stem reads -- non-zero: stage from u8_input instead of the f32 input leaf.  Armed by
    // ggml_b200_graph_upload_u8_images_fused, disarmed by run_plan after the launch that consumed it.
    const ggml_tensor *   u8_leaf        = nullptr;
    uint8_t *             u8_input       = nullptr;
    int *                 u8_flag        = nullptr;
    bool                  u8_armed       = false;
    cudaEvent_t           compute_done   = nullptr;  // recorded after this plan's kernels on its private stream
    cudaStream_t          private_stream = nullptr;  // set by ggml_b200_graph_use_private_stream (pipelined submission)
    bool                  concurrent     = false;    // lane of a group (ggml_b200_graph_group_begin): overlaps its siblings, no FIFO chain
    cudaEvent_t           group_ev = nullptr, fork_ev = nullptr;
    ~Plan();
};

// plan.cpp
Plan * get_or_build_plan(ggml_context * ctx, ggml_cgraph * gf);
void   run_plan(Plan * plan, bool wait_for_results = true);
void   add_launch(Plan * plan, const char * kernel, Launch l, double flops = 0, double bytes = 0, std::string what = "", double bytes_min = -1);
void   destroy_plans_of(ggml_context * ctx);
void   fix_graph_pointers(ggml_cgraph * gf);

// exec_exact.cu : one simple f32-accurate kernel per ggml node
void build_exact_plan(Plan * plan, ggml_cgraph * gf);
// fast path (fuse.cpp + kernels): returns false if the graph is not recognised
bool build_fast_plan(Plan * plan, ggml_cgraph * gf);

// helpers
bool   is_view_op(enum ggml_op op);
TView  make_view(const ggml_tensor * t, void * dptr);
void * device_ptr_of(Plan * plan, const ggml_tensor * t);

}  // namespace b200
