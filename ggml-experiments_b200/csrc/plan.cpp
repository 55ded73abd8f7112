// plan.cpp -- device plans: lifecycle, liveness-based memory planner, transfers, CUDA-graph replay.
//
// Replaces upstream ggml's ggml_graph_compute (CPU thread pool + work buffer) for the reference's
// ggml_graph_compute_with_ctx calls (main.cpp:640, rnn.cpp:158,311).
#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>

#include "fast_kernels.h"
#include "internal.h"

namespace b200 {

Runtime & runtime() {
    static Runtime rt;
    return rt;
}

void ensure_device() {
    Runtime & rt = runtime();
    if (rt.initialised) return;
    int         n   = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess || n == 0) {
        fprintf(stderr,
                "libggml_b200: no usable CUDA device (%s). This library has no CPU fallback: "
                "ggml_graph_compute_with_ctx requires an sm_100a GPU.\n",
                err == cudaSuccess ? "device count is 0" : cudaGetErrorString(err));
        abort();
    }
    B200_CHECK(cudaSetDevice(rt.device));
    cudaDeviceProp prop;
    B200_CHECK(cudaGetDeviceProperties(&prop, rt.device));
    if (prop.major != 10) {
        fprintf(stderr, "libggml_b200: device %d is sm_%d%d; kernels are built for sm_100a only.\n", rt.device, prop.major,
                prop.minor);
        abort();
    }
    rt.sm_count = prop.multiProcessorCount;
    B200_CHECK(cudaStreamCreateWithFlags(&rt.own_stream, cudaStreamNonBlocking));
    const char * m = getenv("GGML_B200_MODE");
    if (m) {
        if (!strcmp(m, "exact")) rt.mode = GGML_B200_MODE_EXACT;
        if (!strcmp(m, "exact_f32")) rt.mode = GGML_B200_MODE_EXACT_F32;
        if (!strcmp(m, "fast")) rt.mode = GGML_B200_MODE_FAST;
    }
    const char * v = getenv("GGML_B200_VERBOSE");
    rt.verbose     = v && atoi(v) > 0;
    const char * g = getenv("GGML_B200_CUDA_GRAPH");
    if (g) rt.use_cuda_graph = atoi(g) != 0;
    rt.initialised = true;
}

cudaStream_t current_stream() {
    Runtime & rt = runtime();
    return rt.use_user_stream ? rt.user_stream : rt.own_stream;
}

// ---- arena planner: best-fit free list with coalescing -------------------------------------------------
int64_t ArenaPlanner::alloc(int64_t bytes) {
    bytes    = align_up(bytes > 0 ? bytes : 1);
    int best = -1;
    for (int i = 0; i < (int)free_list.size(); i++)
        if (free_list[i].size >= bytes && (best < 0 || free_list[i].size < free_list[best].size)) best = i;
    if (best >= 0) {
        int64_t off = free_list[best].off;
        free_list[best].off += bytes;
        free_list[best].size -= bytes;
        if (free_list[best].size == 0) free_list.erase(free_list.begin() + best);
        return off;
    }
    // grow: if the last free block touches the end of the arena, extend it instead of leaving a hole
    for (int i = 0; i < (int)free_list.size(); i++)
        if (free_list[i].off + free_list[i].size == extent) {
            int64_t off = free_list[i].off;
            extent      = off + bytes;
            free_list.erase(free_list.begin() + i);
            return off;
        }
    int64_t off = extent;
    extent += bytes;
    return off;
}

void ArenaPlanner::release(int64_t off, int64_t bytes) {
    bytes = align_up(bytes > 0 ? bytes : 1);
    Block b{off, bytes};
    // insert sorted, then coalesce with neighbours
    size_t pos = 0;
    while (pos < free_list.size() && free_list[pos].off < off) pos++;
    free_list.insert(free_list.begin() + pos, b);
    if (pos + 1 < free_list.size() && free_list[pos].off + free_list[pos].size == free_list[pos + 1].off) {
        free_list[pos].size += free_list[pos + 1].size;
        free_list.erase(free_list.begin() + pos + 1);
    }
    if (pos > 0 && free_list[pos - 1].off + free_list[pos - 1].size == free_list[pos].off) {
        free_list[pos - 1].size += free_list[pos].size;
        free_list.erase(free_list.begin() + pos);
    }
}

void add_launch(Plan * plan, const char * kernel, Launch l, double flops, double bytes, std::string what, double bytes_min) {
    plan->launches.push_back(std::move(l));
    LaunchMeta m;
    m.kernel = kernel;
    m.what   = std::move(what);
    m.flops  = flops;
    m.bytes  = bytes;
    m.bytes_min = bytes_min >= 0 ? bytes_min : bytes;
    plan->meta.push_back(std::move(m));
}

bool is_view_op(enum ggml_op op) {
    return op == GGML_OP_RESHAPE || op == GGML_OP_VIEW || op == GGML_OP_PERMUTE || op == GGML_OP_TRANSPOSE;
}

TView make_view(const ggml_tensor * t, void * dptr) {
    TView v;
    v.p = dptr;
    for (int i = 0; i < 4; i++) {
        v.ne[i] = t->ne[i];
        v.nb[i] = (int64_t)t->nb[i];
    }
    v.type = (int)t->type;
    return v;
}

void * device_ptr_of(Plan * plan, const ggml_tensor * t) {
    auto it = plan->slots.find(t);
    if (it == plan->slots.end() || it->second.dptr == nullptr) B200_ABORT("tensor '%s' (op %d) has no device buffer in this plan", t->name, (int)t->op);
    return it->second.dptr;
}

// ---- plan registry ------------------------------------------------------------------------------------
static std::mutex                               g_plans_mu;
static std::multimap<ggml_context *, Plan *>    g_plans;
static std::mutex g_reg_mu;  // guards g_external / g_feedback (one host thread per device may register concurrently)
static std::unordered_map<const ggml_tensor *, void *> g_external;  // leaf -> caller-owned device memory
static std::unordered_map<const ggml_cgraph *, std::vector<std::pair<ggml_tensor *, ggml_tensor *>>> g_feedback;

Plan::~Plan() {
    if (private_stream) { cudaStreamSynchronize(private_stream); cudaStreamDestroy(private_stream); }
    if (compute_done) {
        if (runtime().compute_chain == compute_done) runtime().compute_chain = nullptr;
        cudaEventDestroy(compute_done);
    }
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    if (group_ev) cudaEventDestroy(group_ev);
    if (fork_ev) cudaEventDestroy(fork_ev);
    if (u8_stage) cudaFree(u8_stage);
    if (history) cudaFree(history);
    for (void * p : owned_device) cudaFree(p);
    for (void * p : pinned) cudaHostUnregister(p);
    // tensors (and views of them) whose ->data points into the mirror must not keep a dangling pointer: a rebuilt plan would
    // skip them ("already has host data") and the host would read -- or a download would write -- freed pinned memory
    for (ggml_tensor * t : mirrored) t->data = nullptr;
    if (host_mirror) cudaFreeHost(host_mirror);
}

static uint64_t graph_signature(const ggml_cgraph * gf) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { h ^= v; h *= 1099511628211ull; };
    mix((uint64_t)gf->n_nodes);
    mix((uint64_t)gf->n_leafs);
    for (int i = 0; i < gf->n_nodes; i++) {
        const ggml_tensor * t = gf->nodes[i];
        mix((uint64_t)(uintptr_t)t);
        mix((uint64_t)t->op);
        for (int d = 0; d < 4; d++) { mix((uint64_t)t->ne[d]); mix((uint64_t)t->nb[d]); }
        // flags (ggml_set_output on an interior node needs a new host shadow), sources and op parameters (eps, strides edited in
        // place) all change what the plan must do: a stale plan must never be replayed for them
        mix((uint64_t)(uint32_t)t->flags);
        for (int s = 0; s < GGML_MAX_SRC; s++) mix((uint64_t)(uintptr_t)t->src[s]);
        for (size_t k = 0; k < sizeof(t->op_params) / sizeof(t->op_params[0]); k++) mix((uint64_t)(uint32_t)t->op_params[k]);
    }
    for (int i = 0; i < gf->n_leafs; i++) mix((uint64_t)(uint32_t)gf->leafs[i]->flags);
    mix((uint64_t)runtime().mode);
    return h;
}

void destroy_plans_of(ggml_context * ctx) {
    std::lock_guard<std::mutex> lk(g_plans_mu);
    auto range = g_plans.equal_range(ctx);
    for (auto it = range.first; it != range.second; ++it) delete it->second;
    g_plans.erase(range.first, range.second);
    // registrations keyed by objects of this arena die with it: a later context may get the same addresses (a pinned
    // mem_buffer usually does), and a stale feedback pair would then point at unrelated tensors
    const char * lo = ctx->mem_buffer, * hi = ctx->mem_buffer + ctx->mem_size;
    auto inside = [&](const void * p) { return (const char *)p >= lo && (const char *)p < hi; };
    for (auto & kv : g_plans) {  // plans of OTHER contexts that mirrored tensors of this arena forget them
        auto & m = kv.second->mirrored;
        m.erase(std::remove_if(m.begin(), m.end(), [&](ggml_tensor * t) { return inside(t); }), m.end());
    }
    std::lock_guard<std::mutex> rk(g_reg_mu);
    for (auto it = g_feedback.begin(); it != g_feedback.end();) it = inside(it->first) ? g_feedback.erase(it) : std::next(it);
    for (auto it = g_external.begin(); it != g_external.end();) it = inside(it->first) ? g_external.erase(it) : std::next(it);
}

// Assign device memory to leafs; shared by both plan builders.
//   * leafs of another context than the compute context, and ggml_new_f32 scalars, are CONSTANTS (weights):
//     uploaded once when the plan is built.
//   * leafs of the compute context, or flagged input/param, are INPUTS: uploaded on every compute
//     (main.cpp:627-634 writes the image through ggml_get_data; rnn.cpp:303-310 rewrites id and state).
static void place_leafs(Plan * plan, ggml_cgraph * gf) {
    std::unordered_map<const ggml_tensor *, void *> g_external_snapshot;
    {
        std::lock_guard<std::mutex> rk(g_reg_mu);
        g_external_snapshot = g_external;
    }
    const auto & g_external = g_external_snapshot;  // shadows the global inside this function
    int64_t const_bytes = 0, input_bytes = 0;
    auto    is_input    = [&](const ggml_tensor * t) {
        if (t->flags & 0x100) return false;
        if (t->flags & (GGML_TENSOR_FLAG_INPUT | GGML_TENSOR_FLAG_PARAM)) return true;
        return t->ctx == plan->ctx;
    };
    std::vector<ggml_tensor *> roots;
    for (int i = 0; i < gf->n_leafs; i++) {
        ggml_tensor * t = gf->leafs[i];
        if (t->view_src) continue;  // views of leafs alias their base
        roots.push_back(t);
        int64_t b = ArenaPlanner::align_up((int64_t)ggml_nelements(t) * (int64_t)ggml_type_size(t->type));
        if (g_external.count(t)) continue;
        if (is_input(t)) input_bytes += b; else const_bytes += b;
    }
    char * cpool = nullptr, * ipool = nullptr;
    if (const_bytes) { B200_CHECK(cudaMalloc((void **)&cpool, const_bytes)); plan->owned_device.push_back(cpool); }
    if (input_bytes) { B200_CHECK(cudaMalloc((void **)&ipool, input_bytes)); plan->owned_device.push_back(ipool); }
    plan->weight_bytes = const_bytes;
    int64_t coff = 0, ioff = 0;
    cudaStream_t st = current_stream();
    for (ggml_tensor * t : roots) {
        Slot    s;
        int64_t raw = (int64_t)ggml_nelements(t) * (int64_t)ggml_type_size(t->type);
        int64_t b   = ArenaPlanner::align_up(raw);
        s.bytes     = raw;
        auto ext    = g_external.find(t);
        if (ext != g_external.end()) {
            s.kind = SLOT_EXTERNAL;
            s.dptr = ext->second;
        } else if (is_input(t)) {
            s.kind = SLOT_INPUT;
            s.dptr = ipool + ioff;
            ioff += b;
            if (!t->data) B200_ABORT("input leaf '%s' has no host data", t->name);
            plan->uploads.push_back({t, s.dptr, (size_t)raw});
        } else {
            s.kind = SLOT_CONST;
            s.dptr = cpool + coff;
            coff += b;
            if (!t->data) B200_ABORT("constant leaf '%s' has no host data", t->name);
            B200_CHECK(cudaMemcpyAsync(s.dptr, t->data, (size_t)raw, cudaMemcpyHostToDevice, st));
        }
        plan->slots[t] = s;
    }
    for (int i = 0; i < gf->n_leafs; i++) {
        ggml_tensor * t = gf->leafs[i];
        if (!t->view_src) continue;
        Slot s;
        s.kind = SLOT_ALIAS;
        s.dptr = (char *)device_ptr_of(plan, t->view_src) + t->view_offs;
        plan->slots[t] = s;
    }
    B200_CHECK(cudaStreamSynchronize(st));
}

Plan * get_or_build_plan(ggml_context * ctx, ggml_cgraph * gf) {
    ensure_device();
    uint64_t sig = graph_signature(gf);
    bool had_private_stream = false;
    if (gf->plan) {
        Plan * p = (Plan *)gf->plan;
        if (p->graph_sig == sig) return p;
        // the graph was extended or the mode changed: rebuild (a pipelined slot keeps running on a stream of its own)
        had_private_stream = p->private_stream != nullptr;
        ggml_graph_release_plan(gf);
    }
    Plan * plan      = new Plan();
    if (had_private_stream) B200_CHECK(cudaStreamCreateWithFlags(&plan->private_stream, cudaStreamNonBlocking));
    plan->ctx        = ctx;
    plan->graph_sig  = sig;
    plan->n_nodes_at_build = gf->n_nodes;
    place_leafs(plan, gf);
    bool fast_ok = false;
    if (runtime().mode == GGML_B200_MODE_FAST) {
        fast_ok = build_fast_plan(plan, gf);
        // a MobileViT-sized graph that misses the fused plan runs ~100x slower: say so once per plan (small graphs -- the GRU cell, unit
        // tests -- are expected to take the per-node plan and stay quiet unless GGML_B200_VERBOSE is set)
        if (!fast_ok && (runtime().verbose || gf->n_nodes > 256))
            fprintf(stderr, "libggml_b200: graph (%d nodes) not recognised by the fused planner; using the per-node exact plan (still on the device, much slower). "
                            "GGML_B200_VERBOSE=1 names the pattern that failed.\n", gf->n_nodes);
    }
    if (!fast_ok) build_exact_plan(plan, gf);
    plan->mode = fast_ok ? GGML_B200_MODE_FAST : (runtime().mode == GGML_B200_MODE_EXACT_F32 ? GGML_B200_MODE_EXACT_F32 : GGML_B200_MODE_EXACT);
    // host shadows for graph outputs
    for (int i = 0; i < gf->n_nodes; i++) {
        ggml_tensor * t = gf->nodes[i];
        if (!(t->flags & GGML_TENSOR_FLAG_OUTPUT)) continue;
        auto it = plan->slots.find(t);
        if (it == plan->slots.end() || !it->second.dptr) continue;
        if (!ggml_is_contiguous(t)) {
            // a strided view as an output: mirror its base instead (rnn.cpp:239 `states` is a transpose)
            ggml_tensor * base = t->view_src;
            if (base && plan->slots.count(base) && !base->data) {
                size_t bytes = (size_t)ggml_nelements(base) * ggml_type_size(base->type);
                base->data   = arena_alloc(base->ctx, bytes, 64);
                plan->downloads.push_back({base, plan->slots[base].dptr, bytes});
            }
            if (base && base->data) t->data = (char *)base->data + t->view_offs;
            continue;
        }
        size_t bytes = (size_t)ggml_nelements(t) * ggml_type_size(t->type);
        if (!t->data) t->data = arena_alloc(t->ctx, bytes, 4096);
        plan->downloads.push_back({t, it->second.dptr, bytes});
    }
    // device-side feedback copies (autoregressive loops): node -> input leaf, after the last kernel of each compute
    {
        std::vector<std::pair<ggml_tensor *, ggml_tensor *>> feedback;
        {
            std::lock_guard<std::mutex> rk(g_reg_mu);
            auto fb = g_feedback.find(gf);
            if (fb != g_feedback.end()) feedback = fb->second;
        }
        if (!feedback.empty())
            for (auto & pr : feedback) {
                void *       src   = device_ptr_of(plan, pr.first);
                void *       dst   = device_ptr_of(plan, pr.second);
                const size_t bytes = (size_t)ggml_nelements(pr.second) * ggml_type_size(pr.second->type);
                if (!ggml_is_contiguous(pr.first) || (size_t)ggml_nelements(pr.first) * ggml_type_size(pr.first->type) != bytes)
                    B200_ABORT("ggml_b200_graph_add_feedback: '%s' -> '%s' must be contiguous and of equal size", pr.first->name, pr.second->name);
                add_launch(plan, "feedback_copy", [=](cudaStream_t st) { B200_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st)); },
                           0.0, 2.0 * bytes, pr.second->name);
            }
    }
    // Tiny graphs (rnn_text_generation.cpp: ~60 nodes) keep upstream's "every tensor has host data after compute"
    // semantics: the program reads an intermediate (`output.states->data`, rnn.cpp:307) that was never declared an
    // output.  Every intermediate of at most 64 KiB gets a host mirror owned by the plan; big graphs never pay for this.
    if (plan->mode != GGML_B200_MODE_FAST && gf->n_nodes <= 256) {
        size_t total = 0;
        for (int i = 0; i < gf->n_nodes; i++) {
            ggml_tensor * t = gf->nodes[i];
            if (is_view_op(t->op) || t->data || !plan->slots.count(t)) continue;
            const size_t bytes = (size_t)ggml_nelements(t) * ggml_type_size(t->type);
            if (bytes <= (64u << 10)) total += (bytes + 255) & ~size_t(255);
        }
        if (total) {
            B200_CHECK(cudaMallocHost(&plan->host_mirror, total));
            size_t off = 0;
            for (int i = 0; i < gf->n_nodes; i++) {
                ggml_tensor * t = gf->nodes[i];
                if (is_view_op(t->op) || t->data || !plan->slots.count(t)) continue;
                const size_t bytes = (size_t)ggml_nelements(t) * ggml_type_size(t->type);
                if (bytes > (64u << 10)) continue;
                t->data = (char *)plan->host_mirror + off;
                plan->mirrored.push_back(t);
                off += (bytes + 255) & ~size_t(255);
                plan->downloads.push_back({t, plan->slots[t].dptr, bytes});
            }
        }
        for (int i = 0; i < gf->n_nodes; i++) {  // views of mirrored tensors read through their base
            ggml_tensor * t = gf->nodes[i];
            if (is_view_op(t->op) && !t->data && t->view_src && t->view_src->data) {
                t->data = (char *)t->view_src->data + t->view_offs;
                const char * hm = (const char *)plan->host_mirror;
                if (hm && (const char *)t->data >= hm && (const char *)t->data < hm + total) plan->mirrored.push_back(t);
            }
        }
    }
    gf->plan = plan;
    {
        std::lock_guard<std::mutex> lk(g_plans_mu);
        g_plans.insert({ctx, plan});
    }
    if (runtime().verbose)
        fprintf(stderr, "libggml_b200: plan mode=%s nodes=%d launches=%zu arena=%.1f MiB (naive %.1f MiB) weights=%.1f MiB\n",
                plan->mode == GGML_B200_MODE_FAST ? "fast" : "exact", gf->n_nodes, plan->launches.size(),
                plan->arena_bytes / 1048576.0, plan->naive_bytes / 1048576.0, plan->weight_bytes / 1048576.0);
    return plan;
}

void run_plan(Plan * plan, bool wait_for_results) {
    cudaStream_t st = plan->private_stream ? plan->private_stream : current_stream();
    if (plan->upload_inputs)
        for (const Transfer & u : plan->uploads) B200_CHECK(cudaMemcpyAsync(u.dptr, u.t->data, u.bytes, cudaMemcpyHostToDevice, st));
    Runtime & rt = runtime();
    if (rt.use_cuda_graph && !plan->graph_failed && plan->launches.size() > 1) {
        if (!plan->graph_exec) {
            // capture the launch sequence once; replay afterwards (launch-bound inner loop -> one submit)
            cudaStream_t cs;
            B200_CHECK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
            cudaGraph_t g   = nullptr;
            cudaError_t err = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
            if (err == cudaSuccess) {
                for (auto & l : plan->launches) l(cs);
                err = cudaStreamEndCapture(cs, &g);
            }
            if (err == cudaSuccess && g) err = cudaGraphInstantiate(&plan->graph_exec, g, 0);
            if (g) cudaGraphDestroy(g);
            cudaStreamDestroy(cs);
            if (err != cudaSuccess || !plan->graph_exec) {
                (void)cudaGetLastError();
                plan->graph_failed = true;
                plan->graph_exec   = nullptr;
                static bool warned = false;
                if (rt.verbose || !warned) fprintf(stderr, "libggml_b200: CUDA graph capture failed (%s); launching the %zu kernels directly\n", cudaGetErrorString(err), plan->launches.size());
                warned = true;
            }
        }
    }
    // Pipelined slots: the copies of one slot overlap the kernels of another, but the forwards themselves run FIFO --
    // two forwards sharing the SMs only thrash L2 and fight for CTA slots (bench e2e 9.4 -> see profiles/README.md).
    static const bool chain = getenv("GGML_B200_SLOT_OVERLAP") == nullptr;
    const bool chained = plan->private_stream && chain && !plan->concurrent;
    if (chained && rt.compute_chain) B200_CHECK(cudaStreamWaitEvent(st, rt.compute_chain, 0));
    if (plan->graph_exec) {
        B200_CHECK(cudaGraphLaunch(plan->graph_exec, st));
    } else {
        for (auto & l : plan->launches) l(st);
    }
    B200_CHECK(cudaGetLastError());
    if (plan->u8_armed) {  // the u8 images were for this launch only: the next compute reads the f32 input leaf again
        B200_CHECK(cudaMemsetAsync(plan->u8_flag, 0, sizeof(int), st));
        plan->u8_armed = false;
    }
    if (chained) {
        if (!plan->compute_done) B200_CHECK(cudaEventCreateWithFlags(&plan->compute_done, cudaEventDisableTiming));
        B200_CHECK(cudaEventRecord(plan->compute_done, st));
        rt.compute_chain = plan->compute_done;
    }
    if (plan->download_outputs) {
        for (const Transfer & d : plan->downloads) B200_CHECK(cudaMemcpyAsync(d.t->data, d.dptr, d.bytes, cudaMemcpyDeviceToHost, st));
        if (wait_for_results) B200_CHECK(cudaStreamSynchronize(st));  // ggml semantics: results are readable when compute returns
    }
}

}  // namespace b200

using namespace b200;

// ---- extension API -----------------------------------------------------------------------------------
extern "C" void ggml_graph_release_plan(struct ggml_cgraph * gf) {
    if (!gf->plan) return;
    Plan * p = (Plan *)gf->plan;
    {
        std::lock_guard<std::mutex> lk(g_plans_mu);
        for (auto it = g_plans.begin(); it != g_plans.end(); ++it)
            if (it->second == p) { g_plans.erase(it); break; }
    }
    cudaDeviceSynchronize();
    delete p;
    gf->plan = nullptr;
}

// ---- pipelined submission: several graphs (e.g. two copies of the same forward graph with their own input/output
// buffers) each on a private stream, so the H2D copy of batch i+1 and the D2H of batch i-1 overlap the kernels of batch i.
extern "C" void ggml_b200_graph_use_private_stream(struct ggml_cgraph * gf) {
    if (!gf->plan) B200_ABORT("ggml_b200_graph_use_private_stream: call ggml_b200_graph_prepare first");
    Plan * p = (Plan *)gf->plan;
    if (!p->private_stream) B200_CHECK(cudaStreamCreateWithFlags(&p->private_stream, cudaStreamNonBlocking));
}
extern "C" void ggml_b200_graph_compute_async(struct ggml_context * ctx, struct ggml_cgraph * gf) {
    fix_graph_pointers(gf);
    Plan * plan = get_or_build_plan(ctx, gf);
    run_plan(plan, /*wait_for_results=*/false);
}
// Device-resident autoregressive loop: run the plan `steps` times back to back on its stream without any host
// round trip -- the feedback copies registered with ggml_b200_graph_add_feedback carry the state from step to step -- and keep
// `record` (e.g. the token ids chosen at each step) in a device history that is copied to the host once, at the end.
// Inputs are uploaded for step 0 only.  host_dst receives steps * nbytes(record).  Replaces the host loop of rnn.cpp:293-313.
extern "C" int ggml_b200_graph_compute_steps(struct ggml_context * ctx, struct ggml_cgraph * gf, int steps, struct ggml_tensor * record, void * host_dst) {
    if (steps <= 0 || !record || !host_dst) return 1;
    fix_graph_pointers(gf);
    Plan * plan = get_or_build_plan(ctx, gf);
    auto it     = plan->slots.find(record);
    if (it == plan->slots.end() || !it->second.dptr || !ggml_is_contiguous(record)) return 1;
    const size_t bytes = (size_t)ggml_nelements(record) * ggml_type_size(record->type);
    cudaStream_t st    = plan->private_stream ? plan->private_stream : current_stream();
    if (plan->history_bytes < bytes * (size_t)steps) {  // grow-only and plan-owned: cudaMalloc/cudaFree per call cost up to 100 ms
        B200_CHECK(cudaStreamSynchronize(st));
        if (plan->history) B200_CHECK(cudaFree(plan->history));
        const size_t cap = bytes * (size_t)((steps + 255) / 256 * 256);  // in units of 256 steps: a short warm-up call sizes it for the real one
        B200_CHECK(cudaMalloc(&plan->history, cap));
        plan->history_bytes = cap;
    }
    char * hist = (char *)plan->history;
    const bool up = plan->upload_inputs, down = plan->download_outputs;
    plan->download_outputs = false;
    // the host runs at most kAhead steps ahead of the device (keeps the launch queue short; the device never waits: a step
    // is ~10x longer than its submission)
    constexpr int kAhead = 8;
    cudaEvent_t   ring[kAhead];
    for (auto & e : ring) B200_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    const bool trace = getenv("GGML_B200_STEP_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    if (trace) {
        tev.resize((size_t)steps + 1);
        for (auto & e : tev) B200_CHECK(cudaEventCreate(&e));
        B200_CHECK(cudaEventRecord(tev[0], st));
    }
    for (int t = 0; t < steps; t++) {
        if (t >= kAhead) B200_CHECK(cudaEventSynchronize(ring[t % kAhead]));
        plan->upload_inputs = up && t == 0;
        run_plan(plan, /*wait_for_results=*/false);
        if (bytes % 4 == 0) launch_copy_words(it->second.dptr, hist + (size_t)t * bytes, (int64_t)(bytes / 4), st);
        else B200_CHECK(cudaMemcpyAsync(hist + (size_t)t * bytes, it->second.dptr, bytes, cudaMemcpyDeviceToDevice, st));
        B200_CHECK(cudaEventRecord(ring[t % kAhead], st));
        if (trace) B200_CHECK(cudaEventRecord(tev[(size_t)t + 1], st));
    }
    for (auto & e : ring) cudaEventDestroy(e);
    if (trace) {
        B200_CHECK(cudaStreamSynchronize(st));
        std::vector<float> ms((size_t)steps);
        for (int t = 0; t < steps; t++) B200_CHECK(cudaEventElapsedTime(&ms[(size_t)t], tev[(size_t)t], tev[(size_t)t + 1]));
        std::vector<float> sorted = ms;
        std::sort(sorted.begin(), sorted.end());
        fprintf(stderr, "step trace: median %.1f us, max %.1f us, slow steps:", sorted[(size_t)steps / 2] * 1e3, sorted.back() * 1e3);
        for (int t = 0; t < steps; t++)
            if (ms[(size_t)t] > 3 * sorted[(size_t)steps / 2]) fprintf(stderr, " %d:%.0f", t, ms[(size_t)t] * 1e3);
        fprintf(stderr, "\n");
        for (auto & e : tev) cudaEventDestroy(e);
    }
    B200_CHECK(cudaMemcpyAsync(host_dst, hist, bytes * (size_t)steps, cudaMemcpyDeviceToHost, st));
    B200_CHECK(cudaStreamSynchronize(st));
    plan->upload_inputs    = up;
    plan->download_outputs = down;
    return 0;
}
// ---- concurrent lanes: several graphs (sub-batches of ONE request) on private streams at the same time ----------------------
// A forward of a few images is a chain of ~80 short kernels, each filling a fraction of the 148 SMs and each paying its own
// ramp-up / drain; run as S independent chains on S streams, the chains fill each other's gaps (strong scaling at 32 images per
// GPU: DESIGN.md).  begin: every lane's stream waits for the owner stream (the library's current stream, or lane 0's own stream
// for pipelined slots); the caller then enqueues uploads / ggml_b200_graph_compute_async per lane; end: the owner waits for all.
static cudaStream_t group_owner(Plan ** pl, int on_current_stream) { return on_current_stream ? current_stream() : pl[0]->private_stream; }
extern "C" void ggml_b200_graph_group_begin(struct ggml_cgraph ** gfs, int n, int on_current_stream) {
    std::vector<Plan *> pl((size_t)n);
    for (int i = 0; i < n; i++) {
        if (!gfs[i]->plan) B200_ABORT("ggml_b200_graph_group_begin: call ggml_b200_graph_prepare first");
        pl[(size_t)i] = (Plan *)gfs[i]->plan;
        if (!pl[(size_t)i]->private_stream) B200_CHECK(cudaStreamCreateWithFlags(&pl[(size_t)i]->private_stream, cudaStreamNonBlocking));
        pl[(size_t)i]->concurrent = true;  // lanes of one request overlap by design: no FIFO chain between them
        if (!pl[(size_t)i]->group_ev) B200_CHECK(cudaEventCreateWithFlags(&pl[(size_t)i]->group_ev, cudaEventDisableTiming));
    }
    cudaStream_t owner = group_owner(pl.data(), on_current_stream);
    if (!pl[0]->fork_ev) B200_CHECK(cudaEventCreateWithFlags(&pl[0]->fork_ev, cudaEventDisableTiming));
    B200_CHECK(cudaEventRecord(pl[0]->fork_ev, owner));
    for (int i = 0; i < n; i++)
        if (pl[(size_t)i]->private_stream != owner) B200_CHECK(cudaStreamWaitEvent(pl[(size_t)i]->private_stream, pl[0]->fork_ev, 0));
}
extern "C" void ggml_b200_graph_group_end(struct ggml_cgraph ** gfs, int n, int on_current_stream) {
    std::vector<Plan *> pl((size_t)n);
    for (int i = 0; i < n; i++) pl[(size_t)i] = (Plan *)gfs[i]->plan;
    cudaStream_t owner = group_owner(pl.data(), on_current_stream);
    for (int i = 0; i < n; i++) {
        if (pl[(size_t)i]->private_stream == owner) continue;
        B200_CHECK(cudaEventRecord(pl[(size_t)i]->group_ev, pl[(size_t)i]->private_stream));
        B200_CHECK(cudaStreamWaitEvent(owner, pl[(size_t)i]->group_ev, 0));
    }
}

extern "C" void ggml_b200_graph_wait(struct ggml_cgraph * gf) {
    if (!gf->plan) return;
    Plan * p = (Plan *)gf->plan;
    B200_CHECK(cudaStreamSynchronize(p->private_stream ? p->private_stream : current_stream()));
}

// Pinned host memory for a caller-provided arena (ggml_init_params.mem_buffer): with the input leaf and the output shadows
// in page-locked memory the per-compute copies are true asynchronous DMA.  (Registering ranges of a malloc'ed arena after
// the fact is not safe: two tensors sharing a page end up half pinned and cudaMemcpyAsync rejects them.)
extern "C" void * ggml_b200_host_malloc(size_t bytes) {
    void * p = nullptr;
    int    n = 0;
    if (cudaGetDeviceCount(&n) == cudaSuccess && n > 0 && cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess) return p;
    (void)cudaGetLastError();
    return nullptr;  // no device / no pinned memory: the caller falls back to an ordinary arena
}
extern "C" void ggml_b200_host_free(void * p) {
    if (p) cudaFreeHost(p);
}

extern "C" int ggml_b200_tensor_download(struct ggml_cgraph * gf, struct ggml_tensor * t, void * host_dst) {
    if (!gf->plan) return 1;
    Plan * p = (Plan *)gf->plan;
    auto it  = p->slots.find(t);
    if (it == p->slots.end() || !it->second.dptr || !ggml_is_contiguous(t)) return 1;
    cudaStream_t st = p->private_stream ? p->private_stream : current_stream();
    B200_CHECK(cudaStreamSynchronize(st));
    B200_CHECK(cudaMemcpy(host_dst, it->second.dptr, (size_t)ggml_nelements(t) * ggml_type_size(t->type), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" void ggml_b200_graph_add_feedback(struct ggml_cgraph * gf, struct ggml_tensor * src, struct ggml_tensor * dst) {
    GGML_ASSERT(dst->op == GGML_OP_NONE);
    std::lock_guard<std::mutex> rk(g_reg_mu);
    g_feedback[gf].push_back({src, dst});
}

extern "C" void ggml_b200_set_mode(enum ggml_b200_mode mode) { runtime().mode = (int)mode; }
extern "C" int  ggml_b200_get_mode(void) { return runtime().mode; }

extern "C" int ggml_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}
extern "C" void ggml_b200_set_device(int device) {
    Runtime & rt = runtime();
    if (rt.initialised && rt.device != device) B200_ABORT("ggml_b200_set_device must be called before the first compute (one process per GPU)");
    rt.device = device;
}
extern "C" void ggml_b200_set_stream(void * stream) {
    Runtime & rt       = runtime();
    rt.user_stream     = (cudaStream_t)stream;
    rt.use_user_stream = true;
}
extern "C" void * ggml_b200_get_stream(void) { ensure_device(); return (void *)current_stream(); }
extern "C" void   ggml_b200_synchronize(void) { ensure_device(); B200_CHECK(cudaStreamSynchronize(current_stream())); }

extern "C" void ggml_b200_tensor_set_device_data(struct ggml_tensor * leaf, void * device_ptr) {
    GGML_ASSERT(leaf->op == GGML_OP_NONE && leaf->view_src == nullptr);
    std::lock_guard<std::mutex> rk(g_reg_mu);
    if (device_ptr) g_external[leaf] = device_ptr; else g_external.erase(leaf);
}
extern "C" void * ggml_b200_tensor_get_device_data(struct ggml_cgraph * gf, struct ggml_tensor * t) {
    if (!gf->plan) return nullptr;
    Plan * p = (Plan *)gf->plan;
    auto it  = p->slots.find(t);
    return it == p->slots.end() ? nullptr : it->second.dptr;
}
// Device-side replacement of sam_image_preprocess + the HWC copy (main.cpp:538-601,627-634): enqueue, on the graph's stream,
// the H2D copy of n raw u8 images and the resize/normalise kernel that writes the device copy of the f32 input leaf.
extern "C" int ggml_b200_graph_upload_u8_images(struct ggml_cgraph * gf, struct ggml_tensor * input, const uint8_t * host_u8, int n, int src_h,
                                                int src_w) {
    if (!gf->plan) B200_ABORT("ggml_b200_graph_upload_u8_images: call ggml_b200_graph_prepare first");
    Plan * p = (Plan *)gf->plan;
    auto it  = p->slots.find(input);
    if (it == p->slots.end() || !it->second.dptr || input->type != GGML_TYPE_F32 || input->ne[0] != 3 || input->ne[3] != n || !host_u8 || src_h <= 0 ||
        src_w <= 0)
        return 1;
    const size_t bytes = (size_t)n * src_h * src_w * 3;
    cudaStream_t st    = p->private_stream ? p->private_stream : current_stream();
    if (p->u8_stage_bytes < bytes) {
        B200_CHECK(cudaStreamSynchronize(st));
        if (p->u8_stage) B200_CHECK(cudaFree(p->u8_stage));
        B200_CHECK(cudaMalloc(&p->u8_stage, bytes));
        p->u8_stage_bytes = bytes;
    }
    B200_CHECK(cudaMemcpyAsync(p->u8_stage, host_u8, bytes, cudaMemcpyHostToDevice, st));
    launch_preprocess_u8((const uint8_t *)p->u8_stage, n, src_h, src_w, (float *)it->second.dptr, (int)input->ne[2], (int)input->ne[1], st);
    B200_CHECK(cudaGetLastError());
    return 0;
}
// The same request when the next compute is all that will read the images: in a FAST plan whose stem is the tensor-core kernel, the
// quantised u8 image (sam_image_preprocess before its /255, main.cpp:592-597) is all that is written -- 3 bytes per pixel instead of
// 12 -- and the stem stages its patch from it (f16(v / 255) through a table: bit-identical to the f32 route).  Images that already have
// the target size are copied straight into that buffer: the resize is the identity there (scale 1: dx = dy = 0).  The f32 input
// leaf is NOT written; other plans fall back to ggml_b200_graph_upload_u8_images.  GGML_B200_NO_U8_STEM=1 forces the fallback.
extern "C" int ggml_b200_graph_upload_u8_images_fused(struct ggml_cgraph * gf, struct ggml_tensor * input, const uint8_t * host_u8, int n, int src_h,
                                                      int src_w) {
    if (!gf->plan) B200_ABORT("ggml_b200_graph_upload_u8_images_fused: call ggml_b200_graph_prepare first");
    Plan * p = (Plan *)gf->plan;
    static const bool off = getenv("GGML_B200_NO_U8_STEM") != nullptr;
    if (off || !p->u8_input || p->u8_leaf != input) return ggml_b200_graph_upload_u8_images(gf, input, host_u8, n, src_h, src_w);
    if (input->ne[0] != 3 || input->ne[3] != n || !host_u8 || src_h <= 0 || src_w <= 0) return 1;
    const int    H = (int)input->ne[2], W = (int)input->ne[1];
    const size_t bytes = (size_t)n * src_h * src_w * 3;
    cudaStream_t st    = p->private_stream ? p->private_stream : current_stream();
    if (src_h == H && src_w == W) {
        B200_CHECK(cudaMemcpyAsync(p->u8_input, host_u8, bytes, cudaMemcpyHostToDevice, st));
    } else {
        if (p->u8_stage_bytes < bytes) {
            B200_CHECK(cudaStreamSynchronize(st));
            if (p->u8_stage) B200_CHECK(cudaFree(p->u8_stage));
            B200_CHECK(cudaMalloc(&p->u8_stage, bytes));
            p->u8_stage_bytes = bytes;
        }
        B200_CHECK(cudaMemcpyAsync(p->u8_stage, host_u8, bytes, cudaMemcpyHostToDevice, st));
        launch_preprocess_u8((const uint8_t *)p->u8_stage, n, src_h, src_w, nullptr, H, W, st, p->u8_input);
    }
    B200_CHECK(cudaMemsetAsync(p->u8_flag, 1, sizeof(int), st));
    B200_CHECK(cudaGetLastError());
    p->u8_armed = true;
    return 0;
}
extern "C" void ggml_b200_graph_set_transfers(struct ggml_cgraph * gf, bool upload_inputs, bool download_outputs) {
    if (!gf->plan) B200_ABORT("ggml_b200_graph_set_transfers: call ggml_b200_graph_prepare first");
    Plan * p            = (Plan *)gf->plan;
    p->upload_inputs    = upload_inputs;
    p->download_outputs = download_outputs;
}
// Per-launch device times (CUDA events on the launching stream, direct launches -- not the CUDA-graph replay),
// averaged over `reps` passes after one warm-up pass, written as a JSON array into buf.
extern "C" int ggml_b200_graph_profile_json(struct ggml_cgraph * gf, int reps, char * buf, size_t cap) {
    if (!gf->plan || reps < 1) return -1;
    Plan *       p  = (Plan *)gf->plan;
    cudaStream_t st = current_stream();
    const size_t n  = p->launches.size();
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto & e : ev) B200_CHECK(cudaEventCreate(&e));
    std::vector<double> ms(n, 0.0);
    for (int r = -1; r < reps; r++) {
        B200_CHECK(cudaEventRecord(ev[0], st));
        for (size_t i = 0; i < n; i++) {
            p->launches[i](st);
            B200_CHECK(cudaEventRecord(ev[i + 1], st));
        }
        B200_CHECK(cudaStreamSynchronize(st));
        if (r < 0) continue;
        for (size_t i = 0; i < n; i++) {
            float t = 0;
            B200_CHECK(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
            ms[i] += t;
        }
    }
    for (auto & e : ev) cudaEventDestroy(e);
    std::string out = "[";
    char        tmp[768];
    for (size_t i = 0; i < n; i++) {
        snprintf(tmp, sizeof tmp, "%s{\"kernel\":\"%s\",\"what\":\"%.120s\",\"ms\":%.6f,\"flops\":%.0f,\"bytes\":%.0f,\"bytes_min\":%.0f}", i ? "," : "",
                 p->meta[i].kernel, p->meta[i].what.c_str(), ms[i] / reps, p->meta[i].flops, p->meta[i].bytes, p->meta[i].bytes_min);
        out += tmp;
    }
    out += "]";
    if (out.size() + 1 > cap) return (int)out.size() + 1;
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}

extern "C" void ggml_b200_graph_plan_stats(struct ggml_cgraph * gf, struct ggml_b200_plan_stats * out) {
    memset(out, 0, sizeof(*out));
    if (!gf->plan) return;
    Plan * p            = (Plan *)gf->plan;
    out->mode           = p->mode;
    out->n_graph_nodes  = p->n_nodes_at_build;
    out->n_launches     = (int)p->launches.size();
    out->n_folded       = p->n_folded;
    out->arena_bytes    = p->arena_bytes;
    out->naive_bytes    = p->naive_bytes;
    out->weight_bytes   = p->weight_bytes;
    out->used_cuda_graph = p->graph_exec != nullptr;
}
