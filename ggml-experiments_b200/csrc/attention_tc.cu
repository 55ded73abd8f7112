// attention_tc.cu -- K7 on the 5th-gen tensor cores: softmax(Q K^T / sqrt(d)) V per (image, patch position, head)
// (main.cpp:1073-1086: mul_mat(K,Q) / sqrt(d) -> soft_max -> mul_mat(V^T, P), plus the permutes / conts around them).
//
// One CTA = one (sequence, head, block of 128 queries).  Both contractions run on tcgen05 with f32 accumulators in TMEM:
//
//   TMA   : Q block [128 tokens x DP] and, per 64-key block j, K_j and V_j [64 tokens x DP] -- 5-D boxes over the head-padded qkv
//           buffer {DP, which*heads+h, x/2, y/2, image}, one tensor map per patch position (y%2, x%2): the reference's unfold
//           (main.cpp:721-747) is a stride in the tensor map, never a copy.  Rows are 128 B (64 halves; columns >= DP are TMA zero fill).
//   MMA 1 : S (128 x 64, TMEM) = Q . K_j^T                  both operands K-major (the head dimension is contiguous)
//   warps : thread = one query row: tcgen05.ld S -> running max (online softmax) -> p = exp2((s - m) * log2e / sqrt(d)) -> f16 P_j
//           into shared memory in the K-major 128B-swizzled UMMA layout; the row's output accumulator lives in REGISTERS
//   MMA 2 : PV (128 x DP, TMEM) = P_j . V_j                 A = P_j K-major, B = V_j MN-major (token rows of V as they lie in memory:
//           no transpose; instruction-descriptor bit 16)
//   warps : tcgen05.ld PV -> O = O * exp2(m_old - m_new) + PV ; after the last block O / l -> f16 -> out[token][h*d ..]
//
// A fifth warp issues every TMA and MMA.  K/V blocks are double-buffered; S_{j+1} is issued as soon as the softmax warps have read
// S_j, so it overlaps PV_j and the accumulator update.  Three CTAs fit an SM (64 KB of shared memory, 128 TMEM columns each), which
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
    int rows_q, rows_k;             // map rows (of width W/2) per 64-token box
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

template <int DP>
__global__ void __launch_bounds__(160, 3) k_attention_tc(const __grid_constant__ CUtensorMap map00, const __grid_constant__ CUtensorMap map01,
                                                         const __grid_constant__ CUtensorMap map10, const __grid_constant__ CUtensorMap map11,
                                                         const AttnTcParams p) {
    extern __shared__ uint8_t attn_smem_raw[];
    uint8_t * smem = attn_smem_raw + ((1024u - (smem_u32(attn_smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bars[8];  // q_full, kv_full[2], kv_free[2], s_full, p_full, pv_full
    __shared__ uint32_t tmem_slot;
    const uint32_t sq = smem_u32(smem);               // Q   : 128 x 128 B
    const uint32_t sk = sq + 16384;                   // K[2]: 64 x 128 B each
    const uint32_t sv = sk + 2 * 8192;                // V[2]
    const uint32_t sp = sv + 2 * 8192;                // P   : 128 x 128 B (64 keys)
    const uint32_t q_full = smem_u32(&bars[0]), kv_full = smem_u32(&bars[1]), kv_free = smem_u32(&bars[3]);
    const uint32_t s_full = smem_u32(&bars[5]), p_full = smem_u32(&bars[6]), pv_full = smem_u32(&bars[7]);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // item = (image n, patch position pos, head h, query block qb)
    int it = blockIdx.x;
    const int qb  = it % p.nqb; it /= p.nqb;
    const int h   = it % p.heads; it /= p.heads;
    const int pos = it & 3;
    const int n   = it >> 2;
    const int nkb = p.L / kKB;
    const CUtensorMap * map = pos == 0 ? &map00 : (pos == 1 ? &map01 : (pos == 2 ? &map10 : &map11));

    if (tid == 0) {
        for (int i = 0; i < 8; i++) mbar_init(smem_u32(&bars[i]), i == 6 ? 4u : 1u);  // p_full: one arrive per softmax warp
        fence_barrier_init();
        tma_prefetch_desc(map);
    }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_s = tmem_slot, tmem_pv = tmem_slot + 64u;
    pdl_wait();  // qkv is the previous kernel's output; `out` may still be read by it
    pdl_trigger();

    if (warp == 4) {
        // ===================== control warp: TMA + MMA issue =====================
        if (lane == 0) {
            auto load_kv = [&](int j) {
                const uint32_t b = (uint32_t)(j & 1), bar = kv_full + 8u * b;
                mbar_expect_tx(bar, 2u * 8192u);
                tma_load_5d(sk + b * 8192u, map, 0, 1 * p.heads + h, 0, j * p.rows_k, n, bar);
                tma_load_5d(sv + b * 8192u, map, 0, 2 * p.heads + h, 0, j * p.rows_k, n, bar);
            };
            mbar_expect_tx(q_full, 16384u);
            tma_load_5d(sq, map, 0, h, 0, (qb * 2) * p.rows_q, n, q_full);
            tma_load_5d(sq + 8192u, map, 0, h, 0, (qb * 2 + 1) * p.rows_q, n, q_full);
            load_kv(0);
            if (nkb > 1) load_kv(1);
            const uint32_t idesc_s = attn_idesc(kKB, 0), idesc_pv = attn_idesc(DP, 1);
            const uint64_t qdesc = make_smem_desc(sq, 128), pdesc = make_smem_desc(sp, 128);
            auto issue_s = [&](int j) {
                mbar_wait(kv_full + 8u * (uint32_t)(j & 1), (uint32_t)((j >> 1) & 1));
                tc_fence_after();
                const uint64_t kdesc = make_smem_desc(sk + (uint32_t)(j & 1) * 8192u, 128);
#pragma unroll
                for (int ks = 0; ks < DP / 16; ks++) umma_f16(tmem_s, qdesc + (uint64_t)(2 * ks), kdesc + (uint64_t)(2 * ks), idesc_s, ks != 0);
                umma_commit(s_full);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < nkb; j++) {
                mbar_wait(p_full, (uint32_t)(j & 1));  // P_j is in shared memory; every softmax warp has finished reading S_j
                tc_fence_after();
                // V_j as the MN-major B operand: 64 token rows of 128 B, 8-row groups 1 KiB apart; 16 keys per k-step = +2 KiB
                const uint64_t vdesc = make_smem_desc(sv + (uint32_t)(j & 1) * 8192u, 128);
#pragma unroll
                for (int ks = 0; ks < kKB / 16; ks++) umma_f16(tmem_pv, pdesc + (uint64_t)(2 * ks), vdesc + (uint64_t)(128 * ks), idesc_pv, ks != 0);
                umma_commit(pv_full);
                umma_commit(kv_free + 8u * (uint32_t)(j & 1));
                if (j + 1 < nkb) issue_s(j + 1);  // runs behind PV_j on the tensor pipe, under the accumulator update of block j
                if (j + 2 < nkb) {
                    mbar_wait(kv_free + 8u * (uint32_t)(j & 1), (uint32_t)((j >> 1) & 1));  // S_j and PV_j have read K_j / V_j
                    load_kv(j + 2);
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== softmax warps: thread = query row =====================
        const int      row  = warp * 32 + lane;                          // TMEM lane
        const uint32_t lsel = (uint32_t)(warp * 32) << 16;
        float O[DP];
#pragma unroll
        for (int c = 0; c < DP; c++) O[c] = 0.f;
        float m = -INFINITY, l = 0.f;  // running max (in scaled log2 units) and running sum
        const uint32_t prow = sp + (uint32_t)row * 128u, psw = (uint32_t)row & 7u;
        for (int j = 0; j < nkb; j++) {
            mbar_wait(s_full, (uint32_t)(j & 1));
            tc_fence_after();
            // pass 1: row maximum of the 64 scores
            float mx = -INFINITY;
#pragma unroll
            for (int hf = 0; hf < 2; hf++) {
                float v[32];
                tmem_ld_32x32(tmem_s + lsel + (uint32_t)(hf * 32), v);
#pragma unroll
                for (int c = 0; c < 32; c++) mx = fmaxf(mx, v[c]);
            }
            const float m_new = fmaxf(m, mx * p.scale_log2);
            const float alpha = ex2(m - m_new);  // first block: exp2(-inf) = 0
            // the previous block's P (and its PV accumulator) must have been consumed before P is overwritten / PV is read again
            if (j > 0) {
                mbar_wait(pv_full, (uint32_t)((j - 1) & 1));
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < DP; c0 += 16) {
                    float v[16];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
                          "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
                        : "r"(tmem_pv + lsel + (uint32_t)c0)
                        : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 16; c++) O[c0 + c] += v[c];
                }
            }
            // O currently holds sum_{i<j} (already rescaled); rescale to the new maximum
#pragma unroll
            for (int c = 0; c < DP; c++) O[c] *= alpha;
            // pass 2: p = exp2(s * scale - m_new), row sum, f16 P into the swizzled A tile
            float sum = 0.f;
#pragma unroll
            for (int hf = 0; hf < 2; hf++) {
                float v[32];
                tmem_ld_32x32(tmem_s + lsel + (uint32_t)(hf * 32), v);
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    float e[8];
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        e[c] = ex2(fmaf(v[g * 8 + c], p.scale_log2, -m_new));
                        sum += e[c];
                    }
                    st_shared_v4(prow + ((((uint32_t)(hf * 4 + g)) ^ psw) << 4), pack2(e[0], e[1]), pack2(e[2], e[3]), pack2(e[4], e[5]), pack2(e[6], e[7]));
                }
            }
            l = fmaf(l, alpha, sum);
            m = m_new;
            tc_fence_before();
            fence_proxy_async();  // P (generic proxy) -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // last block's PV, then normalise and store
        mbar_wait(pv_full, (uint32_t)((nkb - 1) & 1));
        tc_fence_after();
        const float inv = 1.0f / l;
        // token of this row -> pixel: t = qb*128 + row, (ty, tx) in the half-resolution lattice of patch position pos
        const int t  = qb * kQB + row;
        const int w2 = p.W >> 1;
        const int ty = t / w2, tx = t - ty * w2;
        const size_t pix = ((size_t)n * p.H + (size_t)(2 * ty + (pos >> 1))) * p.W + (size_t)(2 * tx + (pos & 1));
        __half * o = p.out + pix * p.C + h * p.d;
#pragma unroll
        for (int c0 = 0; c0 < DP; c0 += 16) {
            float v[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
                  "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
                : "r"(tmem_pv + lsel + (uint32_t)c0)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
                if (c0 + c < p.d) {  // d is even (C and heads are multiples of 4 / 8)
                    const uint32_t hv = pack2((O[c0 + c] + v[c]) * inv, (O[c0 + c + 1] + v[c + 1]) * inv);
                    *reinterpret_cast<uint32_t *>(o + c0 + c) = hv;
                }
            }
        }
        tc_fence_before();
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
        B200_CHECK(cudaFuncSetAttribute(k_attention_tc<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        attr = true;
    }
    launch_pdl(k_attention_tc<DP>, dim3((unsigned)items), dim3(160), (size_t)(1024 + 16384 + 4 * 8192 + 16384), st, maps[0], maps[1], maps[2], maps[3], p);
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
    p.scale_log2 = 1.4426950408889634f / sqrtf((float)d);
    p.out = out16;
    const int items = N * 4 * heads * p.nqb;
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
