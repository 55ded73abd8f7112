// vit_stage.cu -- K8: all transformer layers of one MobileViT block in ONE kernel (transformer_layer::forward, main.cpp:988-1172,
// looped by mobile_vit_layer::forward, main.cpp:1196-1204).
//
// A transformer layer only mixes tokens of one sequence (the (H/2)*(W/2) pixels that share a patch position, main.cpp:721-747), so a
// tile of 128 tokens made of WHOLE sequences can be carried through every layer of the stage without ever meeting another tile:
//
//   X (128 tokens x C, f32)  lives in TMEM columns [0, 256) for the whole stage: it is loaded once, and the attention output
//                            projection and the MLP down-projection accumulate straight into it (D += A.B is the residual add)
//   per layer:  LN1(X) -> f16 A tile                                   (4 warps, thread = token row, tcgen05.ld / st.shared)
//               per head h: [q|k|v]_h = A . Wqkv_h^T   -> TMEM         (tcgen05.mma, weights streamed from L2 through a 3-slot ring)
//                           + bias -> f16 Q_h, K_h, V_h in smem
//                           S = Q_h . K_h^T (128 x 128) -> TMEM        (all sequences of the tile at once; a row only reads its own
//                           softmax over the row's own L keys           sequence's L columns, the rest of P is written as zeros)
//                           O_h = P . V_h -> TMEM, / rowsum -> f16 smem
//                           X += O_h . Wo_h^T                          (per-head slice of the output projection)
//               LN2(X + bo) -> f16 A tile
//               per 128-wide chunk j of the hidden layer: U_j = A . W1_j^T -> TMEM (double buffered), + b1, SiLU -> f16 smem,
//                           X += H_j . W2_j^T
//   after the last layer: X + b2 -> out32 / out16 / row statistics (what the last down-projection GEMM of the unfused plan wrote)
//
// Weights never use tensor maps: the plan-time packer (vit_stage_pack) lays every B tile out exactly as the UMMA descriptor reads it
// (128-byte rows, 128B swizzle) in the order the kernel consumes them, so the producer warp streams one flat blob with 1-D bulk copies.
// The rounding points are those of the unfused FAST plan: f16 operands (LN output, q/k/v, P, attention output, hidden layer), f32
// accumulation, f32 residual stream.  LayerNorm statistics are computed directly (two passes over the row in TMEM); gamma is folded
// into the consuming weights and beta into their bias at plan time (W' = f16(W * gamma), b' = b + sum_k beta_k W_k), so the row warps
// only write f16((x - mean) * rstd).  The per-layer bias vectors travel with the weight stream into a small shared-memory block.
//
// Why: at small batches (strong scaling: 32 images per GPU; batch-1 latency) the 35 launches of the L = 64 and L = 16 stages are pure
// launch / pipeline-fill latency (12-25 us each for microseconds of work); here a stage is one launch and q/k/v, P, the hidden layer
// and the residual stream between layers never touch HBM.
#include "vit_stage.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "fast_kernels.h"
#include "internal.h"
#include "pdl.cuh"
#include "ptx_sm100.cuh"

namespace b200 {

using namespace ptx;

namespace {

typedef VitStageLaunch::Params VP;

constexpr int kRB      = 65536;  // attention / hidden-layer scratch
constexpr int kSlots   = 3;      // weight ring
constexpr int kThreads = 192;    // warps 0-3: token rows, warp 4: weight stream, warp 5: MMA issue

enum {
    B_FULL = 0, B_EMPTY = 3, B_A_READY = 6, B_QKV_DONE, B_QKV_DRAINED, B_S_DONE, B_S_LOADED, B_P_READY, B_O_DONE, B_OS_READY, B_PROJ_DONE,
    B_U_DONE, B_H_READY = B_U_DONE + 2, B_DOWN_DONE = B_H_READY + 2, B_VEC_FULL = B_DOWN_DONE + 2, B_VEC_FREE, B_XLOAD, B_COUNT
};

__device__ __forceinline__ uint32_t vs_idesc(int n, int b_mn_major) {
    return (1u << 4) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pk2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ void bulk_load_ws(uint32_t dst, const void * src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
// 64 consecutive columns of this thread's row in two loads and one wait (the wait names every destination register as read-write,
// so no use can be scheduled above it)
__device__ __forceinline__ void tld64(uint32_t taddr, float (&v)[64]) {
    uint32_t * r = reinterpret_cast<uint32_t *>(v);
#pragma unroll
    for (int hf = 0; hf < 2; hf++) {
        uint32_t * q = r + hf * 32;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9]), "=r"(q[10]),
              "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]), "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]),
              "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]), "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]),
              "=r"(q[31])
            : "r"(taddr + (uint32_t)(hf * 32))
            : "memory");
    }
#define W8(o) "+r"(r[o]), "+r"(r[o + 1]), "+r"(r[o + 2]), "+r"(r[o + 3]), "+r"(r[o + 4]), "+r"(r[o + 5]), "+r"(r[o + 6]), "+r"(r[o + 7])
    asm volatile("tcgen05.wait::ld.sync.aligned;" : W8(0), W8(8), W8(16), W8(24) : : "memory");
    asm volatile("" : W8(32), W8(40), W8(48), W8(56) : : "memory");
#undef W8
}
__device__ __forceinline__ void tst32(uint32_t taddr, const float * v) {
    const uint32_t * r = reinterpret_cast<const uint32_t *>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
          "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
          "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
// per-thread 1-D bulk copies (one token row each): global -> shared with mbarrier completion, shared -> global as a bulk group
__device__ __forceinline__ void bulk_load_row(uint32_t dst, const void * src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store_row(void * dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tst_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int DP>
__global__ void __launch_bounds__(kThreads, 1) k_vit_stage(const __grid_constant__ VP p) {
    extern __shared__ uint8_t vs_smem_raw[];
    uint8_t * smem = vs_smem_raw + ((1024u - (smem_u32(vs_smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bars[B_COUNT];
    __shared__ uint32_t tmem_slot;
    const uint32_t sRA = smem_u32(smem), sRB = sRA + (uint32_t)p.ra_bytes, sRing = sRB + (uint32_t)kRB;
    const uint32_t sVec = sRing + (uint32_t)(kSlots * p.slot_bytes);  // this layer's bias vectors: [pend 256 | bqkv | bo 256 | bf1]
    const uint32_t sQ = sRB, sK = sRB + 16384u, sV = sRB + 32768u, sO = sRB + 49152u, sP = sRB;  // P (2 key blocks) replaces Q and K once S is done
    const uint32_t bar0 = smem_u32(&bars[0]);
    auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;  // shuffle: provably warp-uniform -> role loops on the uniform datapath
    if (tid == 0) {
        for (int i = 0; i < B_COUNT; i++) {
            if (i == B_XLOAD) { mbar_init(bar(i), 128u); continue; }  // one arrive (+ its row's bytes) per token thread
            const bool from_rows = i == B_A_READY || i == B_QKV_DRAINED || i == B_S_LOADED || i == B_P_READY || i == B_OS_READY || i == B_H_READY || i == B_H_READY + 1 ||
                                   i == B_VEC_FREE;
            mbar_init(bar(i), from_rows ? 4u : 1u);  // one arrive per token warp / one TMA transaction or tcgen05.commit
        }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(smem_u32(&tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_slot;
    const uint32_t tX = tbase, tW0 = tbase + 256u, tO = tbase + 448u;  // X | [q|k|v]_h, then S | O_h ; the MLP uses [256,384) and [384,512)
    pdl_trigger();

    const int heads = p.heads, num_kb = p.num_kb, NP = p.NP, nch = p.nch, C = p.C;

    if (warp == 4) {
        // ===================== weight stream: the packed blob of every layer, block after block, through the ring =====================
        // (constants: no need to wait for the previous kernel)
        uint32_t it = 0, nvec = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            for (int l = 0; l <= p.n_layers; l++) {
                // bias vectors of layer l (block n_layers: only the bias still pending on X when the stage ends) -> sVec, once the row warps
                // are done with the previous block
                if (nvec > 0) mbar_wait(bar(B_VEC_FREE), (nvec - 1) & 1u);
                nvec++;
                mbar_expect_tx_ws(bar(B_VEC_FULL), (uint32_t)p.vec_stride * 4u);
                bulk_load_ws(sVec, p.vec + (size_t)l * (size_t)p.vec_stride, (uint32_t)p.vec_stride * 4u, bar(B_VEC_FULL));
                if (l == p.n_layers) break;
                const uint8_t * src = p.blob + (size_t)l * (size_t)p.layer_blob_bytes;
                for (int b = 0; b < p.n_blk; b++, it++) {
                    const uint32_t s = it % kSlots, ph = (it / kSlots) & 1u;
                    mbar_wait(bar(B_EMPTY + (int)s), ph ^ 1u);
                    const uint32_t bytes = (uint32_t)p.blk_rows[b] * 128u;
                    mbar_expect_tx_ws(bar(B_FULL + (int)s), bytes);
                    bulk_load_ws(sRing + s * (uint32_t)p.slot_bytes, src, bytes, bar(B_FULL + (int)s));
                    src += bytes;
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ===================== MMA issue (whole warp, elected lane issues) =====================
        uint32_t it = 0, phase = 0;
#ifdef GGML_B200_VIT_PROFILE
        long long t_ring = 0, t_rows = 0, t_all = clock64();
#define VT_M(acc, stmt) do { const long long t0_ = clock64(); stmt; acc += clock64() - t0_; } while (0)
#else
#define VT_M(acc, stmt) do { stmt; } while (0)
#endif
        auto wait = [&](int i) {
            VT_M(t_rows, mbar_wait(bar(i), (phase >> i) & 1u));
            phase ^= 1u << i;
            tc_fence_after();
        };
        auto slot_wait = [&]() -> uint32_t {  // next weight block
            const uint32_t s = it % kSlots, ph = (it / kSlots) & 1u;
            VT_M(t_ring, mbar_wait(bar(B_FULL + (int)s), ph));
            tc_fence_after();
            return sRing + s * (uint32_t)p.slot_bytes;
        };
        auto slot_free = [&]() {
            umma_commit_ws(bar(B_EMPTY + (int)(it % kSlots)));
            it++;
        };
        // D[128 x n] (+)= A (k-blocks of the LN output in R_A) . W^T (k-blocks from the ring)
        auto gemm_from_a = [&](uint32_t tmem_d, int n) {
            const uint32_t idesc = vs_idesc(n, 0);
            for (int kb = 0; kb < num_kb; kb++) {
                const uint32_t sb  = slot_wait();
                const int      rem = C - kb * 64;
                const int      ks  = rem >= 64 ? 4 : (rem + 15) / 16;
                const uint64_t ad = make_smem_desc(sRA + (uint32_t)kb * 16384u, 128), bd = make_smem_desc(sb, 128);
                for (int k = 0; k < ks; k++) umma_f16_ws(tmem_d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                slot_free();
            }
        };
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            for (int l = 0; l < p.n_layers; l++) {
                if (heads > 0) {  // (heads == 0: MLP-only launch, see vit_stage_prepare)
                    wait(B_A_READY);
                    gemm_from_a(tW0, 3 * DP);
                    umma_commit_ws(bar(B_QKV_DONE));
                }
                for (int h = 0; h < heads; h++) {
                    wait(B_QKV_DRAINED);
                    {   // S = Q_h . K_h^T
                        const uint64_t qd = make_smem_desc(sQ, 128), kd = make_smem_desc(sK, 128);
                        const uint32_t idesc = vs_idesc(128, 0);
#pragma unroll
                        for (int k = 0; k < DP / 16; k++) umma_f16_ws(tW0, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc, k != 0);
                        umma_commit_ws(bar(B_S_DONE));
                    }
                    wait(B_S_LOADED);
                    if (h + 1 < heads) {
                        gemm_from_a(tW0, 3 * DP);
                        umma_commit_ws(bar(B_QKV_DONE));
                    }
                    wait(B_P_READY);
                    {   // O_h = P . V_h : A = P (two blocks of 64 keys), B = V_h token rows as they lie (MN-major)
                        const uint64_t vd = make_smem_desc(sV, 128);
                        const uint32_t idesc = vs_idesc(DP, 1);
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const uint64_t pd = make_smem_desc(sP + (uint32_t)(k >> 2) * 16384u, 128);
                            umma_f16_ws(tO, pd + (uint64_t)(2 * (k & 3)), vd + (uint64_t)(128 * k), idesc, k != 0);
                        }
                        umma_commit_ws(bar(B_O_DONE));
                    }
                    wait(B_OS_READY);
                    {   // X += O_h . Wo_h^T
                        const uint32_t sb = slot_wait();
                        const uint64_t od = make_smem_desc(sO, 128), bd = make_smem_desc(sb, 128);
                        const uint32_t idesc = vs_idesc(NP, 0);
#pragma unroll
                        for (int k = 0; k < DP / 16; k++) umma_f16_ws(tX, od + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, 1u);
                        slot_free();
                        umma_commit_ws(bar(B_PROJ_DONE));
                    }
                }
                // ---- MLP ----
                wait(B_A_READY);
                auto up = [&](int j) {
                    const int nj = min(128, p.F - 128 * j);
                    gemm_from_a(tbase + 256u + 128u * (uint32_t)(j & 1), nj);
                    umma_commit_ws(bar(B_U_DONE + (j & 1)));
                };
                up(0);
                if (nch > 1) up(1);
                for (int j = 0; j < nch; j++) {
                    wait(B_H_READY + (j & 1));
                    const int      nj    = min(128, p.F - 128 * j);
                    const uint32_t idesc = vs_idesc(NP, 0);
                    for (int kb = 0; kb * 64 < nj; kb++) {
                        const uint32_t sb  = slot_wait();
                        const int      rem = nj - kb * 64;
                        const int      ks  = rem >= 64 ? 4 : rem / 16;
                        const uint64_t ad = make_smem_desc(sRB + (uint32_t)(j & 1) * 32768u + (uint32_t)kb * 16384u, 128), bd = make_smem_desc(sb, 128);
                        for (int k = 0; k < ks; k++) umma_f16_ws(tX, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, 1u);
                        slot_free();
                    }
                    umma_commit_ws(bar(B_DOWN_DONE + (j & 1)));
                    if (j + 2 < nch) up(j + 2);
                }
            }
        }
#ifdef GGML_B200_VIT_PROFILE
        if (lane == 0 && blockIdx.x == 0)
            printf("vit_stage MMA warp (cycles, CTA 0): total %lld | waiting for weight blocks %lld | waiting for the row warps %lld\n", clock64() - t_all, t_ring, t_rows);
#endif
        __syncwarp();
    } else {
        // ===================== token rows: thread = one token (TMEM lane) =====================
        const int      row  = tid;
        const uint32_t lsel = (uint32_t)(warp * 32) << 16;
        const uint32_t swz  = (uint32_t)row & 7u;
        const uint32_t rowb = (uint32_t)row * 128u;
        const float    inv_c  = 1.0f / (float)C;
        // this layer's bias vectors in shared memory (delivered by the weight stream): [pend 256 | bqkv heads*3*DP | bo 256 | bf1 nch*128]
        const uint32_t vPend = sVec, vQkv = sVec + 1024u, vBo = vQkv + (uint32_t)(heads * 3 * DP) * 4u, vBf1 = vBo + 1024u;
        uint32_t phase = 0;
#ifdef GGML_B200_VIT_PROFILE
        long long tw[B_COUNT], tsec[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64(), t_all = tprev;
        for (int i = 0; i < B_COUNT; i++) tw[i] = 0;
#define VT_SEC(i) do { const long long tn_ = clock64(); tsec[i] += tn_ - tprev; tprev = tn_; } while (0)
#else
#define VT_SEC(i) do { } while (0)
#endif
        auto wait = [&](int i) {
#ifdef GGML_B200_VIT_PROFILE
            const long long t0_ = clock64();
#endif
            mbar_wait(bar(i), (phase >> i) & 1u);
#ifdef GGML_B200_VIT_PROFILE
            tw[i] += clock64() - t0_;
            tprev = clock64();
#endif
            phase ^= 1u << i;
            tc_fence_after();
        };
        auto arrive = [&](int i, bool wrote_smem) {
            if (wrote_smem) fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(i));
        };
        // LayerNorm statistics of the row in TMEM (after folding `pend`, the bias of the GEMM that last accumulated into X) and the
        // normalised row (x - mean) * rstd as the f16 A tile in R_A; gamma / beta live in the consumer's weights / bias (vit_stage_pack)
        auto layer_norm = [&](uint32_t pend) {
            float s1 = 0.f, s2 = 0.f, pivot = 0.f;
            for (int kb = 0; kb < num_kb; kb++) {
                float v[64];
                tld64(tX + lsel + (uint32_t)(kb * 64), v);
                if (pend) {
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        const float4 bb = ld_shared_f4(pend + (uint32_t)(kb * 64 + q * 4) * 4u);
                        v[4 * q] += bb.x; v[4 * q + 1] += bb.y; v[4 * q + 2] += bb.z; v[4 * q + 3] += bb.w;
                    }
                    tst32(tX + lsel + (uint32_t)(kb * 64), v);
                    tst32(tX + lsel + (uint32_t)(kb * 64 + 32), v + 32);
                }
                if (kb == 0) pivot = v[0];
                // sums around a pivot (the row's first element): E[(x-p)^2] - E[x-p]^2 does not cancel when |mean| >> std
#pragma unroll
                for (int j = 0; j < 64; j++) {
                    if (kb * 64 + (j & ~7) < C) {  // C is a multiple of 8
                        const float dl = v[j] - pivot;
                        s1 += dl;
                        s2 = fmaf(dl, dl, s2);
                    }
                }
            }
            if (pend) tst_wait();
            const float mr = s1 * inv_c, var = fmaxf(fmaf(s2, inv_c, -mr * mr), 0.f);
            const float r = rsqrtf(var + p.eps), nmr = -(pivot + mr) * r;
            for (int kb = 0; kb < num_kb; kb++) {
                float v[64];
                tld64(tX + lsel + (uint32_t)(kb * 64), v);
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    uint32_t w[4] = {0u, 0u, 0u, 0u};
                    if (kb * 64 + q * 8 < C) {
#pragma unroll
                        for (int j = 0; j < 4; j++) w[j] = pk2(fmaf(v[q * 8 + 2 * j], r, nmr), fmaf(v[q * 8 + 2 * j + 1], r, nmr));
                    }
                    st_shared_v4(sRA + (uint32_t)kb * 16384u + rowb + (((uint32_t)q ^ swz) << 4), w[0], w[1], w[2], w[3]);
                }
            }
        };

        pdl_wait();  // x32 is the previous kernel's output; the outputs may still be read by it
        const int seq_per_tile = 128 / p.L;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            // token of this row: sequence g = (image, patch position), index j inside it -> pixel (main.cpp:721-747 as index arithmetic)
            const int  g     = tile * seq_per_tile + row / p.L, j = row % p.L;
            bool       valid = g < p.n_seq;
            size_t     pix   = 0;
            if (heads == 0) {  // MLP only: tokens are independent, a tile is 128 consecutive pixels
                pix   = (size_t)tile * 128 + (size_t)row;
                valid = pix < (size_t)p.total_rows;
                if (!valid) pix = 0;
            } else if (valid) {
                const int n = g >> 2, pos = g & 3, ty = j / p.w2, tx = j % p.w2;
                pix = ((size_t)n * p.H + (size_t)(2 * ty + (pos >> 1))) * p.W + (size_t)(2 * tx + (pos & 1));
            }
            // Staging rows in the (idle) A-tile / attention scratch: every thread moves ITS token row with one bulk copy, so all 128 rows
            // of the tile are in flight at once (per-thread 16-byte loads cost a DRAM round trip per 64 columns: 7.7 k cycles per tile).
            // Row pitch C*4 + 16 bytes: one-row-per-thread LDS.128 / STS.128 are conflict-free.
            const uint32_t srow = sRA + (uint32_t)row * ((uint32_t)C * 4u + 16u);
            {   // X <- x32 row
                const float * xr = p.x32 + pix * (size_t)C;
                if (p.stage_ok) {
                    fence_proxy_async();  // the area was last written through the generic proxy (A tiles / staged outputs)
                    if (valid) {
                        mbar_expect_tx(bar(B_XLOAD), (uint32_t)C * 4u);
                        bulk_load_row(srow, xr, (uint32_t)C * 4u, bar(B_XLOAD));
                    } else {
                        mbar_arrive(bar(B_XLOAD));
                    }
                    wait(B_XLOAD);
                }
                for (int kb = 0; kb < num_kb; kb++) {
                    float v[64];
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (valid && kb * 64 + q * 4 < C) {
                            if (p.stage_ok) t = ld_shared_f4(srow + (uint32_t)(kb * 64 + q * 4) * 4u);
                            else t = __ldg(reinterpret_cast<const float4 *>(xr + kb * 64) + q);
                        }
                        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                    }
                    tst32(tX + lsel + (uint32_t)(kb * 64), v);
                    tst32(tX + lsel + (uint32_t)(kb * 64 + 32), v + 32);
                }
                tst_wait();
                if (p.stage_ok) named_bar_sync(1, 128);  // every row has left the staging area: the A tile may be written over it
            }
            VT_SEC(7);
            for (int l = 0; l < p.n_layers; l++) {
                wait(B_VEC_FULL);
                if (heads > 0) {
                    layer_norm(l > 0 ? vPend : 0u);
                    arrive(B_A_READY, true);
                }
                VT_SEC(0);
                for (int h = 0; h < heads; h++) {
                    // ---- [q|k|v]_h + bias -> f16 rows of Q, K, V ----
                    wait(B_QKV_DONE);
#pragma unroll
                    for (int which = 0; which < 3; which++) {
                        float v[64];
                        tld64(tW0 + lsel + (uint32_t)(which * DP), v);  // DP columns are used (the load may run into the next operand)
                        const uint32_t bq  = vQkv + (uint32_t)((h * 3 + which) * DP) * 4u;
                        const uint32_t dst = sRB + (uint32_t)which * 16384u + rowb;
#pragma unroll
                        for (int c0 = 0; c0 < DP; c0 += 8) {
                            const float4 b0 = ld_shared_f4(bq + (uint32_t)c0 * 4u), b1 = ld_shared_f4(bq + (uint32_t)c0 * 4u + 16u);
                            st_shared_v4(dst + (((uint32_t)(c0 >> 3) ^ swz) << 4), pk2(v[c0] + b0.x, v[c0 + 1] + b0.y), pk2(v[c0 + 2] + b0.z, v[c0 + 3] + b0.w),
                                         pk2(v[c0 + 4] + b1.x, v[c0 + 5] + b1.y), pk2(v[c0 + 6] + b1.z, v[c0 + 7] + b1.w));
                        }
                    }
                    arrive(B_QKV_DRAINED, true);
                    VT_SEC(1);
                    // ---- softmax over the row's own sequence ----
                    wait(B_S_DONE);
                    float l_sum;
                    if (p.L == 64) {
                        float v[64];
                        const int kb_own = row >> 6;
                        tld64(tW0 + lsel + (uint32_t)(kb_own * 64), v);
                        arrive(B_S_LOADED, false);
                        float mx4[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
                        for (int c = 4; c < 60; c += 8)
#pragma unroll
                            for (int i = 0; i < 4; i++) mx4[i] = fmax3(mx4[i], v[c + i], v[c + 4 + i]);
#pragma unroll
                        for (int i = 0; i < 4; i++) mx4[i] = fmaxf(mx4[i], v[60 + i]);
                        const float m = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;
                        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
                        const uint32_t own = sP + (uint32_t)kb_own * 16384u + rowb, oth = sP + (uint32_t)(kb_own ^ 1) * 16384u + rowb;
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            float e[8];
#pragma unroll
                            for (int c = 0; c < 8; c++) {
                                e[c] = ex2f(fmaf(v[q * 8 + c], p.scale_log2, -m));
                                sum4[c & 3] += e[c];
                            }
                            st_shared_v4(own + (((uint32_t)q ^ swz) << 4), pk2(e[0], e[1]), pk2(e[2], e[3]), pk2(e[4], e[5]), pk2(e[6], e[7]));
                            st_shared_v4(oth + ((uint32_t)q << 4), 0u, 0u, 0u, 0u);
                        }
                        l_sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
                    } else {
                        // L <= 32: the warp's 32 rows hold 32 / L sequences; each row keeps the L columns of its own one
                        float v[32];
                        tmem_ld_32x32(tW0 + lsel + (uint32_t)(warp * 32), v);
                        arrive(B_S_LOADED, false);
                        const int lo = ((row & 31) / p.L) * p.L, hi = lo + p.L;
                        float m = -INFINITY;
#pragma unroll
                        for (int c = 0; c < 32; c++)
                            if (c >= lo && c < hi) m = fmaxf(m, v[c]);
                        m *= p.scale_log2;
                        float sum = 0.f;
                        float e[32];
#pragma unroll
                        for (int c = 0; c < 32; c++) {
                            e[c] = (c >= lo && c < hi) ? ex2f(fmaf(v[c], p.scale_log2, -m)) : 0.f;
                            sum += e[c];
                        }
                        l_sum = sum;
                        const int cb = warp * 4;  // first of the four 8-key chunks this warp's columns occupy
#pragma unroll
                        for (int q = 0; q < 16; q++) {
                            const uint32_t dst = sP + (uint32_t)(q >> 3) * 16384u + rowb + ((((uint32_t)q & 7u) ^ swz) << 4);
                            const int      r8  = q - cb;
                            if (r8 >= 0 && r8 < 4) {
                                // (r8 is warp-uniform; the four cases are unrolled so that e[] stays in registers)
                                uint32_t w0, w1, w2, w3;
                                switch (r8) {
                                    case 0: w0 = pk2(e[0], e[1]); w1 = pk2(e[2], e[3]); w2 = pk2(e[4], e[5]); w3 = pk2(e[6], e[7]); break;
                                    case 1: w0 = pk2(e[8], e[9]); w1 = pk2(e[10], e[11]); w2 = pk2(e[12], e[13]); w3 = pk2(e[14], e[15]); break;
                                    case 2: w0 = pk2(e[16], e[17]); w1 = pk2(e[18], e[19]); w2 = pk2(e[20], e[21]); w3 = pk2(e[22], e[23]); break;
                                    default: w0 = pk2(e[24], e[25]); w1 = pk2(e[26], e[27]); w2 = pk2(e[28], e[29]); w3 = pk2(e[30], e[31]); break;
                                }
                                st_shared_v4(dst, w0, w1, w2, w3);
                            } else {
                                st_shared_v4(dst, 0u, 0u, 0u, 0u);
                            }
                        }
                    }
                    arrive(B_P_READY, true);
                    VT_SEC(2);
                    // ---- O_h / rowsum -> f16 rows (A operand of this head's slice of the output projection) ----
                    wait(B_O_DONE);
                    if (h > 0) wait(B_PROJ_DONE);  // the previous head's projection has read its O rows
                    {
                        const float inv = 1.0f / l_sum;
                        float v[64];
                        tld64(tO + lsel, v);  // DP columns are used
#pragma unroll
                        for (int c0 = 0; c0 < DP; c0 += 8)
                            st_shared_v4(sO + rowb + (((uint32_t)(c0 >> 3) ^ swz) << 4), pk2(v[c0] * inv, v[c0 + 1] * inv), pk2(v[c0 + 2] * inv, v[c0 + 3] * inv),
                                         pk2(v[c0 + 4] * inv, v[c0 + 5] * inv), pk2(v[c0 + 6] * inv, v[c0 + 7] * inv));
                    }
                    arrive(B_OS_READY, true);
                    VT_SEC(3);
                }
                if (heads > 0) {
                    wait(B_PROJ_DONE);  // every head's projection has accumulated into X; all earlier MMAs (they read R_A and R_B) are complete
                    layer_norm(vBo);
                } else {
                    layer_norm(l > 0 ? vPend : 0u);
                }
                arrive(B_A_READY, true);
                VT_SEC(4);
                // ---- hidden layer: chunk j of U + b1 -> SiLU -> f16 rows ----
                for (int jc = 0; jc < nch; jc++) {
                    const int b = jc & 1, nj = min(128, p.F - 128 * jc);
                    wait(B_U_DONE + b);
                    if (jc >= 2) wait(B_DOWN_DONE + b);  // the down-projection of chunk jc - 2 has read this buffer
                    const uint32_t b1 = vBf1 + (uint32_t)(128 * jc) * 4u;
                    const uint32_t hs = sRB + (uint32_t)b * 32768u + rowb;
                    for (int c0 = 0; c0 < nj; c0 += 64) {
                        float v[64];
                        tld64(tbase + 256u + 128u * (uint32_t)b + lsel + (uint32_t)c0, v);
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            const int c = c0 + q * 8;
                            if (c < nj) {
                                const float4 b0 = ld_shared_f4(b1 + (uint32_t)c * 4u), b4 = ld_shared_f4(b1 + (uint32_t)c * 4u + 16u);
                                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b4.x, b4.y, b4.z, b4.w};
                                float y[8];
#pragma unroll
                                for (int k = 0; k < 8; k++) y[k] = silu_h(v[q * 8 + k] + bb[k]);  // weights and bias carry the 1/2 (vit_stage_pack)
                                st_shared_v4(hs + (uint32_t)(c >> 6) * 16384u + (((uint32_t)q ^ swz) << 4), pk2(y[0], y[1]), pk2(y[2], y[3]), pk2(y[4], y[5]),
                                             pk2(y[6], y[7]));
                            }
                        }
                    }
                    arrive(B_H_READY + b, true);
                    if (jc == nch - 1) arrive(B_VEC_FREE, false);  // this layer's bias vectors are no longer needed: the next block may land
                    VT_SEC(5);
                }
                if (nch >= 2) wait(B_DOWN_DONE + (nch & 1));  // chunk nch - 2
                wait(B_DOWN_DONE + ((nch - 1) & 1));          // chunk nch - 1: X holds the layer's output (minus b2)
            }
            // ---- X + b2 of the last layer -> global ----
            {
                wait(B_VEC_FULL);  // block n_layers: the last layer's b2 in the `pend` slot
                float st_sum = 0.f, st_sq = 0.f;
                for (int kb = 0; kb < num_kb; kb++) {
                    float v[64];
                    tld64(tX + lsel + (uint32_t)(kb * 64), v);
                    if (!valid) continue;
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const int c = kb * 64 + q * 8;
                        if (c >= C) continue;
                        const float4 b0 = ld_shared_f4(vPend + (uint32_t)c * 4u), b4 = ld_shared_f4(vPend + (uint32_t)c * 4u + 16u);
                        float y[8] = {v[q * 8] + b0.x, v[q * 8 + 1] + b0.y, v[q * 8 + 2] + b0.z, v[q * 8 + 3] + b0.w,
                                      v[q * 8 + 4] + b4.x, v[q * 8 + 5] + b4.y, v[q * 8 + 6] + b4.z, v[q * 8 + 7] + b4.w};
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            st_sum += y[k];
                            st_sq = fmaf(y[k], y[k], st_sq);
                        }
                        // one of the outputs (f32 if there is one, else f16) leaves through the row's staging slot and ONE bulk store
                        if (p.out32) {
                            if (p.stage_ok) {
                                st_shared_v4(srow + (uint32_t)c * 4u, __float_as_uint(y[0]), __float_as_uint(y[1]), __float_as_uint(y[2]), __float_as_uint(y[3]));
                                st_shared_v4(srow + (uint32_t)c * 4u + 16u, __float_as_uint(y[4]), __float_as_uint(y[5]), __float_as_uint(y[6]), __float_as_uint(y[7]));
                            } else {
                                float4 * o = reinterpret_cast<float4 *>(p.out32 + pix * (size_t)C + c);
                                o[0] = make_float4(y[0], y[1], y[2], y[3]);
                                o[1] = make_float4(y[4], y[5], y[6], y[7]);
                            }
                        }
                        if (p.out16) {
                            uint4 o;
                            o.x = pk2(y[0], y[1]); o.y = pk2(y[2], y[3]); o.z = pk2(y[4], y[5]); o.w = pk2(y[6], y[7]);
                            if (p.stage_ok && !p.out32) st_shared_v4(srow + (uint32_t)c * 2u, o.x, o.y, o.z, o.w);
                            else *reinterpret_cast<uint4 *>(p.out16 + pix * (size_t)C + c) = o;
                        }
                    }
                }
                if (p.stage_ok) {
                    fence_proxy_async();  // staged row (generic proxy) -> visible to the bulk-copy engine
                    if (valid) {
                        if (p.out32) bulk_store_row(p.out32 + pix * (size_t)C, srow, (uint32_t)C * 4u);
                        else if (p.out16) bulk_store_row(p.out16 + pix * (size_t)C, srow, (uint32_t)C * 2u);
                    }
                    tma_store_commit();
                    tma_store_wait_read();  // the slot is free again (the next tile loads into it); the global write completes by grid end
                }
                if (valid && p.stats) *reinterpret_cast<float2 *>(p.stats + 2 * pix) = make_float2(st_sum, st_sq);
                arrive(B_VEC_FREE, false);
                VT_SEC(6);
            }
        }
#ifdef GGML_B200_VIT_PROFILE
        if (tid == 0 && blockIdx.x == 0) {
            printf("vit_stage row warps (cycles, CTA 0, thread 0): total %lld\n  work: LN1 %lld | qkv drain %lld | softmax %lld | O drain %lld | LN2 %lld | hidden drain %lld | store %lld | load x %lld\n",
                   clock64() - t_all, tsec[0], tsec[1], tsec[2], tsec[3], tsec[4], tsec[5], tsec[6], tsec[7]);
            printf("  waits: vec %lld | qkv_done %lld | s_done %lld | o_done %lld | proj_done %lld | u_done %lld %lld | down_done %lld %lld\n", tw[B_VEC_FULL], tw[B_QKV_DONE],
                   tw[B_S_DONE], tw[B_O_DONE], tw[B_PROJ_DONE], tw[B_U_DONE], tw[B_U_DONE + 1], tw[B_DOWN_DONE], tw[B_DOWN_DONE + 1]);
        }
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tbase, 512);
    }
}

int ceil_div(int a, int b) { return (a + b - 1) / b; }

struct Geo {
    int d, dp, num_kb, NP, nch, L, ck;
    std::vector<int> rows;  // block heights of one layer in streaming order
};
Geo geometry(int H, int W, int C, int heads, int F) {
    Geo g;
    g.d      = heads ? C / heads : 0;
    g.dp     = heads ? attention_padded_head_dim(g.d) : 16;
    g.num_kb = ceil_div(C, 64);
    g.NP     = ceil_div(C, 16) * 16;
    g.nch    = ceil_div(F, 128);
    g.L      = (H / 2) * (W / 2);
    g.ck     = g.num_kb * 64;
    auto nj = [&](int j) { return std::min(128, F - 128 * j); };
    for (int kb = 0; heads > 0 && kb < g.num_kb; kb++) g.rows.push_back(3 * g.dp);
    for (int h = 0; h < heads; h++) {
        if (h + 1 < heads)
            for (int kb = 0; kb < g.num_kb; kb++) g.rows.push_back(3 * g.dp);
        g.rows.push_back(g.NP);
    }
    auto up = [&](int j) { for (int kb = 0; kb < g.num_kb; kb++) g.rows.push_back(nj(j)); };
    up(0);
    if (g.nch > 1) up(1);
    for (int j = 0; j < g.nch; j++) {
        for (int kb = 0; kb * 64 < nj(j); kb++) g.rows.push_back(g.NP);
        if (j + 2 < g.nch) up(j + 2);
    }
    return g;
}

// dynamic shared memory of a launch: alignment slack + A tile + attention / hidden scratch + weight ring + bias block
size_t smem_need(const Geo & g, int heads) {
    int max_rows = 128;
    for (int r : g.rows) max_rows = std::max(max_rows, r);
    return 1024 + (size_t)g.num_kb * 16384 + kRB + (size_t)kSlots * max_rows * 128 + (size_t)(512 + heads * 3 * g.dp + g.nch * 128) * 4;
}

}  // namespace

bool vit_stage_supported(int N, int H, int W, int C, int heads, int F) {
    if (heads == 0) {  // MLP-only launch (LN -> up + SiLU -> down + residual): per-token work, any map
        if (N <= 0 || H <= 0 || W <= 0 || C % 8 || C > 256 || F % 16 || F <= 0) return false;
        const Geo g0 = geometry(H, W, C, 0, F);
        return (int)g0.rows.size() <= 64 && smem_need(g0, 0) + 512 <= 227 * 1024;
    }
    if (N <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || heads <= 0 || heads > 8 || C % heads || C % 8 || C > 256 || F % 16 || F <= 0) return false;
    const int L = (H / 2) * (W / 2), d = C / heads;
    if (L > 64 || (L & (L - 1))) return false;  // whole sequences per 128-token tile; the softmax paths cover L = 64 and L <= 32
    if (d > 64 || (d & 1)) return false;
    const Geo g = geometry(H, W, C, heads, F);
    if ((int)g.rows.size() > 64) return false;
    return smem_need(g, heads) + 512 <= 227 * 1024;  // + the kernel's static shared memory (barriers)
}

void vit_stage_pack(const VitLayerHost * layers, int n_layers, int C, int heads, int F, std::vector<uint8_t> & blob, std::vector<float> & vec) {
    const Geo g = geometry(2, 2, C, heads, F);  // the block list does not depend on the map size
    const int d = g.d, DP = g.dp, NP = g.NP;
    size_t layer_bytes = 0;
    for (int r : g.rows) layer_bytes += (size_t)r * 128;
    // bias block l (l = 0 .. n_layers): [pend 256 | bqkv heads*3*DP | bo 256 | bf1 nch*128]; pend = b2 of layer l - 1 (the bias that is
    // still missing from X when layer l starts; block n_layers carries only that, for the final store)
    const int o_pend = 0, o_bqkv = 256, o_bo = o_bqkv + heads * 3 * DP, o_bf1 = o_bo + 256, stride = o_bf1 + g.nch * 128;
    blob.assign(layer_bytes * n_layers, 0);
    vec.assign((size_t)stride * (n_layers + 1), 0.f);
    for (int l = 0; l < n_layers; l++) {
        const VitLayerHost & w = layers[l];
        uint16_t * dst = reinterpret_cast<uint16_t *>(blob.data() + layer_bytes * l);
        // one block: `rows` x 64 halves, 128-byte rows, 16-byte chunk index XOR (row % 8)
        auto block = [&](int rows, auto && get) {
            for (int r = 0; r < rows; r++)
                for (int c = 0; c < 8; c++)
                    for (int e = 0; e < 8; e++) dst[(size_t)r * 64 + (size_t)((c ^ (r & 7)) * 8 + e)] = ggml_fp32_to_fp16(get(r, c * 8 + e));
            dst += (size_t)rows * 64;
        };
        // LayerNorm folded into its consumer (main.cpp:1002-1019 then :1022 / :1134): LN(x).W + b = xhat.(gamma * W) + (b + beta.W)
        auto qkv = [&](int h) {
            const float * ws[3] = {w.wq, w.wk, w.wv};
            for (int kb = 0; kb < g.num_kb; kb++)
                block(3 * DP, [&](int r, int k) -> float {
                    const int which = r / DP, i = r % DP, kk = kb * 64 + k;
                    return (i < d && kk < C) ? ws[which][(size_t)kk * C + (h * d + i)] * w.ln1_g[kk] : 0.f;
                });
        };
        auto up = [&](int j) {
            const int nj = std::min(128, F - 128 * j);
            for (int kb = 0; kb < g.num_kb; kb++)
                block(nj, [&](int r, int k) -> float {
                    const int kk = kb * 64 + k;
                    return kk < C ? 0.5f * w.w1[(size_t)kk * F + (128 * j + r)] * w.ln2_g[kk] : 0.f;  // 1/2: the SiLU drain works on h = y/2
                });
        };
        if (heads > 0) qkv(0);
        for (int h = 0; h < heads; h++) {
            if (h + 1 < heads) qkv(h + 1);
            block(NP, [&](int r, int k) -> float { return (r < C && k < d) ? w.wo[(size_t)(h * d + k) * C + r] : 0.f; });
        }
        up(0);
        if (g.nch > 1) up(1);
        for (int j = 0; j < g.nch; j++) {
            const int nj = std::min(128, F - 128 * j);
            for (int kb = 0; kb * 64 < nj; kb++)
                block(NP, [&](int r, int k) -> float {
                    const int kk = kb * 64 + k;
                    return (r < C && kk < nj) ? w.w2[(size_t)(128 * j + kk) * C + r] : 0.f;
                });
            if (j + 2 < g.nch) up(j + 2);
        }
        float * v = vec.data() + (size_t)stride * l;
        for (int c = 0; c < C; c++) {
            if (heads > 0) v[o_bo + c] = w.bo[c];
            v[stride + o_pend + c] = w.b2[c];  // pending bias of the NEXT block
        }
        for (int f = 0; f < F; f++) {
            double acc = w.b1[f];
            for (int k = 0; k < C; k++) acc += (double)w.ln2_b[k] * (double)w.w1[(size_t)k * F + f];
            v[o_bf1 + f] = (float)(0.5 * acc);
        }
        const float * ws[3] = {w.wq, w.wk, w.wv};
        const float * bs[3] = {w.bq, w.bk, w.bv};
        for (int h = 0; h < heads; h++)
            for (int which = 0; which < 3; which++)
                for (int i = 0; i < d; i++) {
                    double acc = bs[which][h * d + i];
                    for (int k = 0; k < C; k++) acc += (double)w.ln1_b[k] * (double)ws[which][(size_t)k * C + (h * d + i)];
                    v[o_bqkv + (h * 3 + which) * DP + i] = (float)acc;
                }
    }
}

bool vit_stage_prepare(VitStageLaunch & L, const float * x32, int N, int H, int W, int C, int heads, int F, int n_layers, float eps,
                       const uint8_t * blob_dev, const float * vec_dev, float * out32, __half * out16, float * stats) {
    if (!vit_stage_supported(N, H, W, C, heads, F) || n_layers <= 0) return false;
    const Geo g = geometry(H, W, C, heads, F);
    L = VitStageLaunch();
    VP & p = L.p;
    p.N = N; p.H = H; p.W = W; p.C = C; p.heads = heads; p.d = g.d; p.F = F; p.L = g.L; p.n_layers = n_layers;
    p.n_seq  = 4 * N;
    p.tiles  = ceil_div(p.n_seq * g.L, 128);
    p.total_rows = N * H * W;
    p.stage_ok   = (128 * (C * 4 + 16) <= g.num_kb * 16384 + kRB && getenv("GGML_B200_VIT_NO_STAGE") == nullptr) ? 1 : 0;
    if (heads == 0) {  // MLP only: 128 consecutive pixels per tile
        p.L     = 128;
        p.tiles = ceil_div(p.total_rows, 128);
    }
    p.num_kb = g.num_kb; p.NP = g.NP; p.nch = g.nch; p.w2 = W / 2;
    p.n_blk  = (int)g.rows.size();
    size_t layer_bytes = 0;
    int    max_rows    = 128;
    for (int i = 0; i < p.n_blk; i++) {
        p.blk_rows[i] = (uint16_t)g.rows[i];
        layer_bytes += (size_t)g.rows[i] * 128;
        max_rows = std::max(max_rows, g.rows[i]);
    }
    p.slot_bytes = max_rows * 128;
    p.ra_bytes   = g.num_kb * 16384;
    p.eps        = eps;
    p.scale_log2 = 1.4426950408889634f / sqrtf((float)g.d);
    p.x32 = x32; p.out32 = out32; p.out16 = out16; p.stats = stats;
    p.blob = blob_dev; p.layer_blob_bytes = (long long)layer_bytes;
    p.vec  = vec_dev;
    p.vec_stride = 512 + heads * 3 * g.dp + g.nch * 128;
    L.dp         = g.dp;
    L.grid       = std::min(p.tiles, runtime().sm_count);
    L.smem_bytes = 1024 + (size_t)p.ra_bytes + kRB + (size_t)kSlots * p.slot_bytes + (size_t)p.vec_stride * 4;
    return L.smem_bytes + 512 <= 227 * 1024;  // + the kernel's static shared memory (barriers)
}

template <int DP>
static void vit_launch_dp(const VitStageLaunch & L, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        B200_CHECK(cudaFuncSetAttribute(k_vit_stage<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 512));
        attr = true;
    }
    launch_pdl(k_vit_stage<DP>, dim3((unsigned)L.grid), dim3(kThreads), L.smem_bytes, st, L.p);
}

void vit_stage_launch(const VitStageLaunch & L, cudaStream_t st) {
    switch (L.dp) {
        case 16: vit_launch_dp<16>(L, st); break;
        case 32: vit_launch_dp<32>(L, st); break;
        case 48: vit_launch_dp<48>(L, st); break;
        default: vit_launch_dp<64>(L, st); break;
    }
}

}  // namespace b200
