// attention_tc.cu -- K7 on the 5th-gen tensor cores: softmax(Q K^T / sqrt(d)) V per (image, patch position, head)
// (main.cpp:1073-1086: mul_mat(K,Q) / sqrt(d) -> soft_max -> mul_mat(V^T, P), plus the permutes / conts around them).
//
// One CTA = one (sequence, head, block of 128 queries).  Both contractions run on tcgen05 with f32 accumulators in TMEM:
//
//   TMA   : Q block [128 tokens x DP] and, per 64-key block j, K_j and V_j [64 tokens x DP] -- 5-D boxes over the head-padded qkv
//           buffer {DP, which*heads+h, x/2, y/2, image}, one tensor map per patch position (y%2, x%2): the reference's unfold
//           (main.cpp:721-747) is a stride in the tensor map, never a copy.  Rows are 128 B (64 halves; columns >= DP are TMA zero fill).
//   MMA 1 : S (128 x 64, TMEM) = Q . K_j^T                  both operands K-major (the head dimension is contiguous)
//   warps : thread = one query row: tcgen05.ld S (64 scores into registers, S is released at once) -> reference max (online softmax
//           with a lazy rescale, see kTau) -> p = exp2(s * log2e / sqrt(d) - m) -> f16 P_j into shared memory in the K-major
//           128B-swizzled UMMA layout
//   MMA 2 : O (128 x DP, TMEM) += P_j . V_j                 A = P_j K-major, B = V_j MN-major (token rows of V as they lie in memory:
//           no transpose; instruction-descriptor bit 16); O never leaves TMEM during the key loop
//   warps : after the last block tcgen05.ld O -> O / l -> f16 -> out[token][h*d ..]
//
// A fifth warp issues every TMA and MMA.  K/V blocks are double-buffered; S_{j+1} is issued as soon as the softmax warps hold S_j
// in registers, so it runs under their exponentials.  Three CTAs fit an SM (64 KB of shared memory, 128 TMEM columns each), which
// is what overlaps the softmax of one query block with the MMAs / loads of the others.
//
// Used for sequences of a multiple of 128 tokens whose map width divides into 64-token row blocks (L = 256 / 1024 of the first ViT
// stage at 256^2 / 512^2 inputs); shorter sequences (L = 64, 16: 7 of the 9 layers, 15 % of the attention FLOPs) stay on the
// mma.sync kernel of fast_kernels.cu, whose 16-query warps fit them better than a 128-row MMA tile.
#include <cmath>
#include <mutex>
#include <vector>

#include "fast_kernels.h"
#include "gemm_tcgen05.h"
#include "internal.h"
#include "pdl.cuh"
#include "ptx_sm100.cuh"

namespace b200 {

using namespace ptx;

namespace {

constexpr int kQB = 128;  // queries per CTA (UMMA M)
constexpr int kKB = 64;   // keys per block (UMMA N of S, K of the second MMA)

struct AttnTcParams {
    int H, W, C, heads, d, L, nqb;  // map size, channels, heads, head dim, tokens per sequence, query blocks per sequence
    int items;                      // N * 4 * heads * nqb
    int rows_q, rows_k;             // map rows (of width W/2) per 64-token box
    int w2_shift;                   // log2(W/2): the envelope makes W/2 a power of two (it divides 64)
    float scale_log2;               // log2(e) / sqrt(d)
    __half * out;                   // [N*H*W][C]
};

__device__ __forceinline__ uint32_t attn_idesc(int n, int b_mn_major) {
    return (1u << 4) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap * map, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// 32 lanes x 16 consecutive f32 columns: TMEM <-> registers (used to rescale / read the output accumulator)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
          "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]),
          "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 64 score columns of this thread's row in two loads, one wait
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
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
    // the wait names every destination register as read-write, so no use of them can be scheduled above it
#define W8(o) "+r"(r[o]), "+r"(r[o + 1]), "+r"(r[o + 2]), "+r"(r[o + 3]), "+r"(r[o + 4]), "+r"(r[o + 5]), "+r"(r[o + 6]), "+r"(r[o + 7])
    asm volatile("tcgen05.wait::ld.sync.aligned;" : W8(0), W8(8), W8(16), W8(24) : : "memory");
    asm volatile("" : W8(32), W8(40), W8(48), W8(56) : : "memory");
#undef W8
}

// The output accumulator O (128 x DP, f32) stays in TMEM for the whole key loop: P_j . V_j accumulates into it.  The softmax keeps a
// per-row reference maximum m and only moves it -- rescaling l and the row of O by 2^(m_old - m_new) -- when the block maximum exceeds
// it by more than kTau (in log2 units): p = 2^(s - m) then stays below 2^kTau, well inside f16, and after the first block or two the
// rescale (a TMEM read-modify-write, done warp-wide when any lane needs it) practically never happens.
constexpr float kTau = 8.0f;

template <int DP>
__global__ void __launch_bounds__(160, 3) k_attention_tc(const __grid_constant__ CUtensorMap map00, const __grid_constant__ CUtensorMap map01,
                                                         const __grid_constant__ CUtensorMap map10, const __grid_constant__ CUtensorMap map11,
                                                         const AttnTcParams p) {
    // Persistent CTA: its items (sequence, head, query block) form one flat sequence of key blocks g = 0, 1, ...; every barrier counts
    // in g, so the loads and the first S of the next item are issued while the softmax warps still work on the current one.
    extern __shared__ uint8_t attn_smem_raw[];
    uint8_t * smem = attn_smem_raw + ((1024u - (smem_u32(attn_smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bars[16];  // q_full, q_free, s_full, s_free, p_full, pv_full, k_full[3], k_free[3], v_full[2], v_free[2]
    __shared__ uint32_t tmem_slot;
    const uint32_t sq = smem_u32(smem);               // Q   : 128 x 128 B
    const uint32_t sk = sq + 16384;                   // K[3]: 64 x 128 B each (needed one block ahead of the softmax: three deep)
    const uint32_t sv = sk + 3 * 8192;                // V[2]
    const uint32_t sp = sv + 2 * 8192;                // P   : 128 x 128 B (64 keys)
    const uint32_t q_full = smem_u32(&bars[0]), q_free = smem_u32(&bars[1]), s_full = smem_u32(&bars[2]), s_free = smem_u32(&bars[3]);
    const uint32_t p_full = smem_u32(&bars[4]), pv_full = smem_u32(&bars[5]);
    const uint32_t k_full = smem_u32(&bars[6]), v_full = smem_u32(&bars[12]), v_free = smem_u32(&bars[14]);  // (bars[9..11], once "K buffer free", are unused)

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;  // shuffle: provably warp-uniform -> role loops on the uniform datapath
    const int nkb   = p.L / kKB;
    const int nit   = (int)blockIdx.x < p.items ? (p.items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;  // items of this CTA
    const int nblk  = nit * nkb;                                                                            // key blocks of this CTA
    // item = (image n, patch position pos, head h, query block qb)
    auto decode = [&](int k, int & n, int & pos, int & h, int & qb) {
        int it = (int)blockIdx.x + k * (int)gridDim.x;
        qb  = it % p.nqb; it /= p.nqb;
        h   = it % p.heads; it /= p.heads;
        pos = it & 3;
        n   = it >> 2;
    };

    if (tid == 0) {
        for (int i = 0; i < 16; i++) mbar_init(smem_u32(&bars[i]), (i == 3 || i == 4) ? 4u : 1u);  // s_free / p_full: one arrive per softmax warp
        fence_barrier_init();
        tma_prefetch_desc(&map00);
        tma_prefetch_desc(&map01);
        tma_prefetch_desc(&map10);
        tma_prefetch_desc(&map11);
    }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_s = tmem_slot, tmem_o = tmem_slot + 64u;
    pdl_wait();  // qkv is the previous kernel's output; `out` may still be read by it
    pdl_trigger();

    if (warp == 4) {
        // ===================== control warp: TMA + MMA issue (whole warp, elected lane issues: ptx_sm100.cuh "_ws") =====================
        if (nblk > 0) {
            auto map_of = [&](int pos) { return pos == 0 ? &map00 : (pos == 1 ? &map01 : (pos == 2 ? &map10 : &map11)); };
            // Every stream of this warp (K loads, V loads, S issue, P.V issue) walks the CTA's key blocks g = 0, 1, 2, ... in order: a cursor
            // keeps (item, block inside the item) and decodes the item once -- the run-time divisions of g / nkb, g % nkb and decode() per
            // call were a third of the warp's instructions (they run on the uniform datapath, which has no divider)
            struct Cursor { int k = -1, j = 0, n = 0, pos = 0, h = 0, qb = 0; };
            auto step = [&](Cursor & c) {  // position the cursor on the next block
                if (c.k < 0 || ++c.j == nkb) {
                    c.j = 0;
                    decode(++c.k, c.n, c.pos, c.h, c.qb);
                }
            };
            Cursor ck, cv, cs;
            int    jpv = 0;  // P.V: block inside the item
            auto load_k = [&](int g) {  // keys of block g of this CTA -> K buffer g % 3
                step(ck);
                const uint32_t b = (uint32_t)(g % 3), bar = k_full + 8u * b;
                mbar_expect_tx_ws(bar, 8192u);
                tma_load_5d_ws(sk + b * 8192u, map_of(ck.pos), 0, 1 * p.heads + ck.h, 0, ck.j * p.rows_k, ck.n, bar);
            };
            auto load_v = [&](int g) {  // values of block g -> V buffer g & 1
                step(cv);
                const uint32_t b = (uint32_t)(g & 1), bar = v_full + 8u * b;
                mbar_expect_tx_ws(bar, 8192u);
                tma_load_5d_ws(sv + b * 8192u, map_of(cv.pos), 0, 2 * p.heads + cv.h, 0, cv.j * p.rows_k, cv.n, bar);
            };
            auto load_q = [&](int k) {
                int n, pos, h, qb;
                decode(k, n, pos, h, qb);
                mbar_expect_tx_ws(q_full, 16384u);
                tma_load_5d_ws(sq, map_of(pos), 0, h, 0, (qb * 2) * p.rows_q, n, q_full);
                tma_load_5d_ws(sq + 8192u, map_of(pos), 0, h, 0, (qb * 2 + 1) * p.rows_q, n, q_full);
            };
            const uint32_t idesc_s = attn_idesc(kKB, 0), idesc_pv = attn_idesc(DP, 1);
            const uint32_t q_lo = smem_desc_lo(sq), p_lo = smem_desc_lo(sp);  // descriptors as 32-bit halves (ptx_sm100.cuh)
            auto issue_s = [&](int g) {  // S of key block g; the first block of an item waits for the item's Q, the last one releases it
                step(cs);
                const int k = cs.k, j = cs.j;
                if (j == 0) mbar_wait(q_full, (uint32_t)(k & 1));
                mbar_wait(k_full + 8u * (uint32_t)(g % 3), (uint32_t)((g / 3) & 1));
                tc_fence_after();
                const uint32_t k_lo = smem_desc_lo(sk + (uint32_t)(g % 3) * 8192u);
#pragma unroll
                for (int ks = 0; ks < DP / 16; ks++) umma_f16_ws_split(tmem_s, q_lo + (uint32_t)(2 * ks), k_lo + (uint32_t)(2 * ks), smem_desc_hi(128), idesc_s, ks != 0);
                umma_commit_ws(s_full);  // (no separate "K buffer free" commit: see the K reload below)
                if (j == nkb - 1) {
                    umma_commit_ws(q_free);  // every S of this item has been issued: Q may be replaced once they complete
                    if (k + 1 < nit) {
                        mbar_wait(q_free, (uint32_t)(k & 1));
                        load_q(k + 1);
                    }
                }
            };
            load_q(0);
            for (int g = 0; g < 3 && g < nblk; g++) load_k(g);
            load_v(0);
            issue_s(0);
            for (int g = 0; g < nblk; g++) {
                if (g + 1 < nblk) {
                    // the softmax warps hold S_g in registers: S_{g+1} (possibly of the next item) is computed under their exponentials
                    mbar_wait(s_free, (uint32_t)(g & 1));
                    issue_s(g + 1);
                    // V_{g+1} -> the buffer P.V_{g-1} has read (it was issued a whole softmax ago); needed only after the NEXT softmax
                    if (g >= 1) mbar_wait(v_free + 8u * (uint32_t)((g - 1) & 1), (uint32_t)(((g - 1) >> 1) & 1));
                    load_v(g + 1);
                }
                // K_{g+3} -> the buffer S_g has read.  g + 3 < nblk implies the s_free wait above: the softmax warps have loaded S_g, so the
                // MMA that produced it -- the only reader of that K buffer -- is complete; no barrier of its own (a tcgen05.commit costs
                // ~100 cycles of issue time on this warp, which is on the critical path together with the softmax warps)
                if (g + 3 < nblk) load_k(g + 3);
                mbar_wait(p_full, (uint32_t)(g & 1));  // P_g is in shared memory (and any rescale of O has been stored)
                mbar_wait(v_full + 8u * (uint32_t)(g & 1), (uint32_t)((g >> 1) & 1));
                tc_fence_after();
                // V_g as the MN-major B operand: 64 token rows of 128 B, 8-row groups 1 KiB apart; 16 keys per k-step = +2 KiB
                const uint32_t v_lo = smem_desc_lo(sv + (uint32_t)(g & 1) * 8192u);
                const int j = jpv;
                if (++jpv == nkb) jpv = 0;
#pragma unroll
                for (int ks = 0; ks < kKB / 16; ks++) umma_f16_ws_split(tmem_o, p_lo + (uint32_t)(2 * ks), v_lo + (uint32_t)(128 * ks), smem_desc_hi(128), idesc_pv, (j | ks) != 0);
                umma_commit_ws(pv_full);
                umma_commit_ws(v_free + 8u * (uint32_t)(g & 1));
            }
        }
        __syncwarp();
    } else {
        // ===================== softmax warps: thread = query row =====================
        const int      row  = warp * 32 + lane;                          // TMEM lane
        const uint32_t lsel = (uint32_t)(warp * 32) << 16;
        const uint32_t prow = sp + (uint32_t)row * 128u, psw = (uint32_t)row & 7u;
        int g = 0;
#ifdef GGML_B200_ATTN_PROFILE
        long long tacc[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#define AT_TICK(i) do { if (tid == 0) { const long long tn = clock64(); tacc[i] += tn - tprev; tprev = tn; } } while (0)
#else
#define AT_TICK(i) do { } while (0)
#endif
        for (int k = 0; k < nit; k++) {
            float m = 0.f, l = 0.f;  // reference maximum (scaled log2 units) and running sum
            for (int j = 0; j < nkb; j++, g++) {
                AT_TICK(6);
                mbar_wait(s_full, (uint32_t)(g & 1));
                tc_fence_after();
                AT_TICK(j == 0 ? 0 : 7);
                float v[64];
                tmem_ld64(tmem_s + lsel, v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(s_free);  // S is in registers: the tensor core may overwrite it with the next S
                AT_TICK(1);
                float mx4[4] = {v[0], v[1], v[2], v[3]};  // four independent chains of three-input maxima (FMNMX3) instead of 63 dependent FMNMX
#pragma unroll
                for (int c = 4; c < 60; c += 8)
#pragma unroll
                    for (int i = 0; i < 4; i++) mx4[i] = fmax3(mx4[i], v[c + i], v[c + 4 + i]);
#pragma unroll
                for (int i = 0; i < 4; i++) mx4[i] = fmaxf(mx4[i], v[60 + i]);
                const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;
                // the previous P.V has finished: P may be overwritten and O may be rescaled
                AT_TICK(2);
                if (g > 0) {
                    mbar_wait(pv_full, (uint32_t)((g - 1) & 1));
                    tc_fence_after();
                }
                AT_TICK(3);
                if (j == 0) {
                    m = mx;
                } else {
                    const bool need = mx > m + kTau;
                    if (__any_sync(0xffffffffu, need)) {
                        const float m_new = need ? mx : m;
                        const float alpha = ex2(m - m_new);  // 1 for the lanes that keep their reference
                        l *= alpha;
                        m = m_new;
#pragma unroll
                        for (int c0 = 0; c0 < DP; c0 += 16) {
                            float o[16];
                            tmem_ld16(tmem_o + lsel + (uint32_t)c0, o);
#pragma unroll
                            for (int c = 0; c < 16; c++) o[c] *= alpha;
                            tmem_st16(tmem_o + lsel + (uint32_t)c0, o);
                        }
                        tmem_st_wait();
                    }
                }
                // p = exp2(s * scale - m), row sum, f16 P into the swizzled A tile
                float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int gq = 0; gq < 8; gq++) {
                    float e[8];
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        e[c] = ex2(fmaf(v[gq * 8 + c], p.scale_log2, -m));
                        sum4[c & 3] += e[c];
                    }
                    st_shared_v4(prow + ((((uint32_t)gq) ^ psw) << 4), pack2(e[0], e[1]), pack2(e[2], e[3]), pack2(e[4], e[5]), pack2(e[6], e[7]));
                }
                l += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
                AT_TICK(4);
                tc_fence_before();
                fence_proxy_async();  // P (generic proxy) -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full);
                AT_TICK(5);
            }
            // all of this item's P.V has accumulated: normalise and store
            mbar_wait(pv_full, (uint32_t)((g - 1) & 1));
            tc_fence_after();
            AT_TICK(3);
            const float inv = 1.0f / l;
            int n, pos, h, qb;
            decode(k, n, pos, h, qb);
            // O/l -> f16, staged in this warp's 32 rows of the (now idle) P tile so that the global stores run along the
            // head's channels: one thread per row would write 4 B to 32 different pixels per instruction.
            // Row pitch DP/2 + 2 words: 8-byte aligned rows (read back two words per lane below), two-way conflicts at most on the way in;
            // DP = 64 keeps dense rows of 32 words, rotated by twice the row index.
            constexpr int kPitch = DP / 2 + 2;
            static_assert(32 * (DP / 2 + 2) * 4 <= 4096 || DP == 64, "staging rows of one warp must fit its 4 KiB quarter of the P tile");
            const uint32_t stage = sp + (uint32_t)warp * 4096u;
            const uint32_t srow  = stage + (uint32_t)lane * (uint32_t)((DP == 64 ? 32 : kPitch) * 4);
#pragma unroll
            for (int c0 = 0; c0 < DP; c0 += 16) {
                float v[16];
                tmem_ld16(tmem_o + lsel + (uint32_t)c0, v);
#pragma unroll
                for (int c = 0; c < 16; c += 2) {
                    const int w = (c0 + c) >> 1;
                    st_shared_u32(srow + (uint32_t)((DP == 64 ? ((w + 2 * lane) & 31) : w) * 4), pack2(v[c] * inv, v[c + 1] * inv));
                }
            }
            tc_fence_before();
            AT_TICK(8);
            __syncwarp();
            AT_TICK(9);
            // Global stores: a lane carries 4 channels (8 bytes) of a pixel's head slice, d/4 lanes per row, 32 / (d/4) rows per instruction
            // (3 rows for d = 36 instead of 1 with 4-byte lanes: the store phase was 20 % of the kernel, profiles/README.md); head dims that
            // are not a multiple of 4 keep 4-byte lanes
            const int wpl = (p.d & 3) ? 1 : 2, lp = p.d / (2 * wpl), rstep = 32 / lp, sub = lane / lp, w = lane - sub * lp;
            const int t0 = qb * kQB + warp * 32;
            if (sub < rstep) {
                __half * obase = p.out + h * p.d + 2 * wpl * w;
#pragma unroll 2
                for (int r = sub; r < 32; r += rstep) {
                    // token of this row -> pixel: (ty, tx) in the half-resolution lattice of patch position pos (W/2 is a power of two)
                    const int t  = t0 + r;
                    const int ty = t >> p.w2_shift, tx = t & ((1 << p.w2_shift) - 1);
                    const size_t pix = ((size_t)n * p.H + (size_t)(2 * ty + (pos >> 1))) * p.W + (size_t)(2 * tx + (pos & 1));
                    auto word = [&](int i) {  // word i (channels 2i, 2i+1) of staged row r
                        return ld_shared_u32(stage + (uint32_t)((DP == 64 ? r * 32 + ((i + 2 * r) & 31) : r * kPitch + i) * 4));
                    };
                    if (wpl == 2) *reinterpret_cast<uint2 *>(obase + pix * p.C) = make_uint2(word(2 * w), word(2 * w + 1));
                    else *reinterpret_cast<uint32_t *>(obase + pix * p.C) = word(w);
                }
            }
            AT_TICK(10);
            __syncwarp();  // the next block's P overwrites the staging rows
            tc_fence_before();
        }
#ifdef GGML_B200_ATTN_PROFILE
        if (tid == 0 && blockIdx.x == 0)
            printf("attention_tc phases (cycles, CTA 0, %d items x %d blocks): wait S (first block of item) %lld | ld S %lld | max %lld | wait PV (incl. item end) %lld | exp+P %lld | fence+arrive %lld | store %lld | wait S (other blocks) %lld | O->smem %lld | syncwarp %lld | smem->global %lld\n",
                   nit, nkb, tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5], tacc[6], tacc[7], tacc[8], tacc[9], tacc[10]);
#endif
#undef AT_TICK
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_slot, 128);
    }
}

template <int DP>
void launch_tc(const CUtensorMap * maps, const AttnTcParams & p, int items, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        B200_CHECK(cudaFuncSetAttribute(k_attention_tc<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
        attr = true;
    }
    const int cap = 3 * runtime().sm_count;  // persistent: three CTAs per SM
    launch_pdl(k_attention_tc<DP>, dim3((unsigned)(items < cap ? items : cap)), dim3(160), (size_t)(1024 + 16384 + 5 * 8192 + 16384), st, maps[0], maps[1], maps[2],
               maps[3], p);
}

}  // namespace

// Returns false when the shape is outside this kernel's envelope (the caller then uses the mma.sync kernel).
bool launch_attention_tc(const __half * qkv, int N, int H, int W, int C, int heads, __half * out16, cudaStream_t st) {
    if (getenv("GGML_B200_ATTN_NO_TC") != nullptr) return false;
    const int d = C / heads, dp = attention_padded_head_dim(d);
    const int w2 = W / 2, h2 = H / 2, L = w2 * h2;
    if (dp > 64 || (d & 1) || L < 128 || L % kQB || w2 > kKB || kKB % w2 || (h2 % (kKB / w2))) return false;
    // tensor maps are cached per (buffer, shape): plans call this on every replay (and under CUDA-graph capture)
    struct Key { const void * q; int N, H, W, C, heads; };
    struct Entry { Key k; CUtensorMap maps[4]; };
    static std::vector<Entry> cache;
    static std::mutex mu;
    const CUtensorMap * maps = nullptr;
    {
        std::lock_guard<std::mutex> lk(mu);
        for (const Entry & e : cache)
            if (e.k.q == qkv && e.k.N == N && e.k.H == H && e.k.W == W && e.k.C == C && e.k.heads == heads) { maps = e.maps; break; }
        if (!maps) {
            Entry e;
            e.k = Key{qkv, N, H, W, C, heads};
            const uint64_t P = (uint64_t)3 * heads * dp * 2;  // bytes per pixel of the qkv buffer
            for (int pos = 0; pos < 4; pos++) {
                const uint64_t dims[5] = {(uint64_t)dp, (uint64_t)(3 * heads), (uint64_t)w2, (uint64_t)h2, (uint64_t)N};
                const uint64_t str[4]  = {(uint64_t)dp * 2, 2 * P, 2 * (uint64_t)W * P, (uint64_t)H * W * P};
                const uint32_t box[5]  = {64, 1, (uint32_t)w2, (uint32_t)(kKB / w2), 1};
                const char * base = (const char *)qkv + ((size_t)(pos >> 1) * W + (size_t)(pos & 1)) * P;
                tma_encode(&e.maps[pos], base, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
            }
            if (cache.size() > 256) cache.clear();
            cache.push_back(e);
            maps = cache.back().maps;
        }
    }
    AttnTcParams p;
    p.H = H; p.W = W; p.C = C; p.heads = heads; p.d = d; p.L = L; p.nqb = L / kQB;
    p.rows_k = kKB / w2; p.rows_q = kKB / w2;
    p.w2_shift = 0;
    while ((1 << p.w2_shift) < w2) p.w2_shift++;
    p.scale_log2 = 1.4426950408889634f / sqrtf((float)d);
    p.out = out16;
    const int items = N * 4 * heads * p.nqb;
    p.items = items;
    CUtensorMap local[4];
    {
        std::lock_guard<std::mutex> lk(mu);
        for (int i = 0; i < 4; i++) local[i] = maps[i];  // the cache vector may reallocate after the lock is released
    }
    switch (dp) {
        case 16: launch_tc<16>(local, p, items, st); break;
        case 32: launch_tc<32>(local, p, items, st); break;
        case 48: launch_tc<48>(local, p, items, st); break;
        default: launch_tc<64>(local, p, items, st); break;
    }
    return true;
}

}  // namespace b200
