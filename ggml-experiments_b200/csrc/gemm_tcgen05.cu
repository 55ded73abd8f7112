// gemm_tcgen05.cu -- K1: GEMM / implicit-GEMM 3x3 convolution on the 5th-gen tensor cores (sm_100a).
//
// Replaces ggml_conv_2d's im2col + ggml_mul_mat (main.cpp:798, [ggml-upstream] im2col(F16) + vec_dot_f16)
// and the dense ggml_mul_mat linears (main.cpp:1022,1039,1056,1095,1134,1151) of the reference.
//
//   D[M,N] (f32, TMEM) = A[M,K] (f16, K-major, smem via TMA) x B[N,K]^T (f16, K-major, smem via TMA)
//
// * operands are f16 and accumulation is f32, exactly the rounding points of ggml's f16 conv path;
// * persistent CTAs (1-2 per SM) loop over 128 x block_n output tiles (UMMA M=128, N=block_n<=256, K=16 per
//   instruction) with two accumulator stages in TMEM, so the epilogue of tile i overlaps the TMA+MMA of tile i+1;
// * warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2..5 = epilogue
//   (tcgen05.ld -> BN scale/shift or bias -> SiLU -> residual -> vectorised stores);
// * smem ring of `stages` x (A 128x64 + B block_n x 64) f16 tiles in the 128-byte swizzled K-major layout
//   shared by TMA (CU_TENSOR_MAP_SWIZZLE_128B) and the UMMA shared-memory descriptors;
// * the 3x3 convolution walks K as (source, 64-channel block, kw, kh): each step is one 4-D TMA box
//   {64 ch, W, rows, images} shifted by the tap offset, with the hardware zero-filling the halo, so no
//   im2col buffer ever exists; a second source tensor map implements ggml_concat (main.cpp:1219) for free.
//
#include "gemm_tcgen05.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "internal.h"
#include "pdl.cuh"
#include "ptx_sm100.cuh"

namespace b200 {

static constexpr int kBlockM   = 128;
static constexpr int kBlockK   = 64;   // 64 f16 = 128 bytes = one swizzle-128B row
static constexpr int kThreads  = 352;  // 11 warps: TMA, MMA, 2 x 4 epilogue, second MMA issuer (conv3x3 pair mode)
static constexpr int kMaxStage = 8;  // deep rings only for small grids (choose_tiling): a lone CTA per SM hides the TMA latency with loads in flight, not with neighbours
static constexpr int kCtrlBytes = 4096;
enum { kEpiAct = 1, kEpiRes32 = 2, kEpiOut16 = 4, kEpiOut32 = 8, kEpiLn = 16, kEpiStats = 32, kEpiRes16 = 64 };  // barriers + TMEM slot + per-column scale/shift, padded to keep 1 KiB alignment

using namespace ptx;  // mbarrier / TMA / tcgen05 wrappers: ptx_sm100.cuh (one copy for every kernel file)

// Instruction descriptor (cute InstrDescriptor): c=f32, a=b=f16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int block_n) {
    return (1u << 4) | ((uint32_t)(block_n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------
// EPI: compile-time epilogue variant (bit mask, kEpi* below) so the per-element loops carry no run-time feature tests;
// EPI = -1 is the generic kernel that reads the flags from the parameters.
// Position in a ring of n slots -- slot index and parity of the current pass -- advanced without the integer divisions that
// `it % n`, `(it / n) & 1` cost on the uniform datapath (a run-time divisor is a ~30-instruction sequence, once per k-block)
struct RingPos {
    uint32_t s = 0, ph = 0;
    __device__ __forceinline__ void advance(uint32_t n) {
        if (++s == n) { s = 0; ph ^= 1u; }
    }
};

#ifdef GGML_B200_GEMM_PROFILE
// clock64 phase profile of the halo-mode conv (build with GGML_B200_GEMM_PROFILE=1): cycles each role spends waiting on each barrier
__device__ unsigned long long g_gemm_prof[64];
__device__ unsigned long long g_gemm_prof1[16];  // per-tap conv at W == 8: [0] producer waits on empty, [1] producer total, [2] CTAs, [3] MMA waits on full, [4] MMA waits tmem, [5] MMA total, [6] epilogue wait on tmem_full (warp 2), [7] epilogue total
__device__ int g_gemm_noload;  // probe: the halo producer arrives on the full barriers without loading anything (pure MMA-loop timing)
#define PROF_WAIT(BAR, PAR, ACC)                 \
    do {                                         \
        const long long _t0 = clock64();         \
        mbar_wait((BAR), (PAR));                 \
        (ACC) += clock64() - _t0;                \
    } while (0)
#else
#define PROF_WAIT(BAR, PAR, ACC) mbar_wait((BAR), (PAR))
#endif

template <int EPI>
__global__ void __launch_bounds__(kThreads) k_gemm_tcgen05(const __grid_constant__ CUtensorMap map_a0,
                                                           const __grid_constant__ CUtensorMap map_a1,
                                                           const __grid_constant__ CUtensorMap map_b,
                                                           const __grid_constant__ CUtensorMap map_o16,
                                                           const __grid_constant__ CUtensorMap map_o32,
                                                           const __grid_constant__ CUtensorMap map_r32,
                                                           const GemmLaunch::Params p) {
    // Persistent CTA: blockIdx.y fixes the N tile, blockIdx.x strides over the M tiles.  Three decoupled pipelines:
    //   smem ring   full/empty[stages]   TMA producer  <-> MMA issuer      (runs continuously across tiles)
    //   TMEM ring   tmem_full/empty[2]   MMA issuer    <-> epilogue warps  (tile i+1 accumulates while tile i drains)
    // Two epilogue warp groups (warps 2-5 / 6-9) own one TMEM accumulator stage each and alternate tiles, so the
    // SiLU/convert/store work of consecutive tiles overlaps and each SM sub-partition always has epilogue warps to issue.
    extern __shared__ uint8_t smem_raw[];
    // swizzle-128B atoms need 1 KiB alignment; offset arithmetic (not a uintptr_t round trip) keeps the pointer in the
    // shared address space for the compiler (LDS/STS instead of generic LD/ST)
    uint8_t * smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int       a_bytes     = kBlockM * p.kb_elems * 2;
    const int       b_bytes     = p.block_n * p.kb_elems * 2;
    // Plain GEMMs keep the whole weight tile (all K blocks) resident: it is the same for every M tile of a persistent CTA, and
    // re-loading it per tile made the output-heavy transformer GEMMs L2->SM bound (qkv: 754 MB of operand traffic per launch,
    // 60 % of it weights).  The ring then holds A only.  The implicit-GEMM conv (9 taps x C of weights) keeps B in the ring.
    const int       stage_bytes = a_bytes + (p.b_resident ? 0 : b_bytes);
    uint8_t *       smem_bres   = smem;
    uint8_t *       ring        = smem + (p.b_resident ? (size_t)p.num_kb * b_bytes : 0);
    uint64_t *      bars        = (uint64_t *)(ring + (size_t)p.ring_bytes);
    uint64_t *      full_bar    = bars;
    uint64_t *      empty_bar   = bars + kMaxStage;
    uint64_t *      tmem_full   = bars + 2 * kMaxStage;      // [4]
    uint64_t *      tmem_empty  = bars + 2 * kMaxStage + 4;  // [4]
    uint64_t *      res_full    = bars + 2 * kMaxStage + 8;  // [2] residual slab landed (per epilogue group)
    uint32_t *      tmem_slot   = (uint32_t *)(bars + 2 * kMaxStage + 10);
    uint64_t *      bres_full   = bars + 2 * kMaxStage + 11;  // resident weight tile landed
    float *         s_scale     = (float *)(bars + 2 * kMaxStage + 16);  // 16-byte aligned: read back with LDS.128
    float *         s_shift     = s_scale + 256;
    float *         s_c1        = s_shift + 256;  // LayerNorm folding: per-column sum of the gamma-scaled weights
    uint64_t *      bfull_bar   = (uint64_t *)(s_c1 + 256);  // conv == 2: weight ring [kMaxStage] full / [kMaxStage] empty
    uint64_t *      bempty_bar  = bfull_bar + kMaxStage;
    // epilogue staging (per epilogue warp group): 128 rows x 128 B tiles in the TMA 128B-swizzle layout
    uint8_t *       stage_base  = ring + (size_t)p.ring_bytes + kCtrlBytes;
    const int       stg16_bytes = p.ep.out16 ? kBlockM * 128 : 0;      // 64 f16 columns per row
    const int       stg32_bytes = (p.ep.out32 || p.ep.res32) ? 2 * kBlockM * 128 : 0;  // 2 x 32 f32 columns per row

    // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the role loops' barrier addresses, smem / TMEM
    // addresses and UMMA descriptors in uniform registers (no per-MMA R2UR / VOTEU chains in front of every tcgen05.mma)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * p.block_n;
    const int tile_m      = p.tile_m;  // 128 except for conv tiles that are a whole number of image rows
    const int num_m_tiles = (p.M + tile_m - 1) / tile_m;

    // conv3x3 pair mode: the two M tiles of a box are issued by TWO warps (1 and 10, different SM sub-partitions), one accumulator each.
    // The issue loop -- barrier polls, commits and the uniform-datapath descriptor arithmetic -- costs ~120 cycles per MMA against 56 of
    // math at N = 96 (tests/conv_prof.py with loads and epilogue off), so a second issuer nearly doubles the tensor pipe's duty cycle.
    const uint32_t n_issuers = (p.conv == 2 && p.pair) ? 2u : 1u;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a0);
        tma_prefetch_desc(&map_b);
        if (p.cblk1 > 0) tma_prefetch_desc(&map_a1);
        if (p.ep.out16) tma_prefetch_desc(&map_o16);
        if (p.ep.out32) tma_prefetch_desc(&map_o32);
        if (p.ep.res32) tma_prefetch_desc(&map_r32);
        for (int s = 0; s < p.stages; s++) {
            mbar_init(smem_u32(&full_bar[s]), p.a_cp_async ? 33 : 1);  // TMA thread (+ 32 cp.async lanes)
            mbar_init(smem_u32(&empty_bar[s]), n_issuers);
        }
        for (int a = 0; a < 4; a++) {
            mbar_init(smem_u32(&tmem_full[a]), 1);
            mbar_init(smem_u32(&tmem_empty[a]), 4);  // one arrive per epilogue warp
        }
        for (int a = 0; a < 2; a++) mbar_init(smem_u32(&res_full[a]), 1);
        mbar_init(smem_u32(bres_full), 1);
        for (int s = 0; s < kMaxStage; s++) {
            mbar_init(smem_u32(&bfull_bar[s]), 1);
            mbar_init(smem_u32(&bempty_bar[s]), n_issuers);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
    // SiLU epilogues work on h = y/2 (silu(y) = h + h tanh h): the 1/2 is folded into scale and shift here, exactly (a power of two)
    const float ah = p.ep.act ? 0.5f : 1.0f;
    for (int i = threadIdx.x; i < p.block_n; i += blockDim.x) {
        const int n = n0 + i;
        s_scale[i]  = ((p.ep.scale && n < p.N) ? p.ep.scale[n] : 1.0f) * ah;
        s_shift[i]  = ((p.ep.shift && n < p.N) ? p.ep.shift[n] : 0.0f) * ah;
        s_c1[i]     = (p.ep.ln_c1 && n < p.N) ? p.ep.ln_c1[n] : 0.0f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // The resident weight tile is a constant: its TMA loads are issued before the PDL wait so that they, too, overlap the
    // tail of the previous kernel (both producer flavours: one k-block per load, num_kb == 1 in the cp.async mode)
    // (every TMA / MMA below is issued warp-uniformly: all lanes of the warp run the loop, one elected lane issues -- ptx_sm100.cuh "_ws")
    if (warp == 0 && p.b_resident && (int)blockIdx.x < num_m_tiles) {
        mbar_expect_tx_ws(smem_u32(bres_full), (uint32_t)(p.num_kb * b_bytes));
        for (int kb = 0; kb < p.num_kb; kb++) tma_load_2d_ws(smem_u32(smem_bres + (size_t)kb * b_bytes), &map_b, kb * p.kb_elems, n0, smem_u32(bres_full));
    }
    // PDL: everything above (barriers, TMEM, per-column constants, weights) overlapped the tail of the previous kernel;
    // activations, residuals and row statistics are only touched, and outputs only written, once that kernel has completed
    pdl_wait();
    pdl_trigger();

    const int cblk_tot = p.cblk0 + p.cblk1;

    if (warp == 0 && p.a_cp_async) {
        // ===================== producer, small-K mode: A by cp.async (whole warp), B by TMA =====================
        // With K = 16 / 32 a TMA box row is only 32 / 64 bytes and the TMA's per-row cost dominates (measured: 2.2 TB/s
        // at K=16).  The A tile is a contiguous 128 x K block, so the warp copies it with 16-byte cp.async into the
        // 32B / 64B-swizzled K-major layout the UMMA descriptor expects; completion is signalled on the same mbarrier.
        const int      cpr   = p.kb_elems >> 3;              // 16-byte chunks per row (2 or 4)
        const int      chunks = kBlockM * cpr;
        const uint32_t row_bytes = (uint32_t)p.kb_elems * 2;
        RingPos rp;
        for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x, rp.advance((uint32_t)p.stages)) {
            const int      m0 = tile * tile_m;
            const int      s  = (int)rp.s;
            const uint32_t ph = rp.ph;
            mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
            const uint32_t fb = smem_u32(&full_bar[s]);
            const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
            if (p.b_resident) {
                mbar_arrive_ws(fb);  // the weight tile was requested before the PDL wait
            } else {
                mbar_expect_tx_ws(fb, (uint32_t)b_bytes);
                tma_load_2d_ws(sa + a_bytes, &map_b, 0, n0, fb);
            }
            const int cshift = cpr == 2 ? 1 : 2;
            for (int i = lane; i < chunks; i += 32) {
                const int      row = i >> cshift, ch = i & (cpr - 1);
                const uint32_t sw  = cpr == 2 ? ((uint32_t)row >> 2) & 1u : ((uint32_t)row >> 1) & 3u;
                const uint32_t dst = sa + (uint32_t)row * row_bytes + (((uint32_t)ch ^ sw) << 4);
                const int      m   = m0 + row;
                const __half * src = p.a_ptr + (size_t)(m < p.M ? m : 0) * p.lda + ch * 8;
                const uint32_t nbytes = m < p.M ? 16u : 0u;  // rows past M are zero-filled
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(fb) : "memory");
        }
    } else if (warp == 0 && p.conv == 2) {
        // ===================== producer, halo mode: per (64-channel block, kw) ONE activation box with a one-row halo above and
        // below -- the three kh taps are start-address offsets of W rows into it -- and per tap one pre-tiled weight block by a 1-D
        // bulk copy.  A third of the activation boxes (and TMA box rows) of the per-tap scheme, no tensor map for the weights. =====
        RingPos ra, rb;
        const uint32_t n_sa = (uint32_t)p.stages, n_sb = (uint32_t)p.b_stages;
        long long w_aempty = 0, w_bempty = 0;
        const long long t_role0 = clock64();
        const uint8_t * ringB = ring + (size_t)p.stages * p.a_slot_bytes;
        // pair mode: the box holds TWO vertically adjacent M tiles (2 * tile_m + 2W rows): every weight block and every activation
        // row is fetched once per 256 pixels -- the kernel is bound by the L2 -> SM rate, so the bytes per pixel are what counts
        const int      tpb         = p.pair ? 2 : 1;  // M tiles per box
        const uint32_t a_box_bytes = (uint32_t)(tpb * p.tile_m + 2 * p.W) * 128u, wb_bytes = (uint32_t)p.block_n * 128u;
        for (int tile = blockIdx.x * tpb; tile < num_m_tiles; tile += gridDim.x * tpb) {
            const int m0 = tile * tile_m, hw = p.H * p.W;
            const int img = m0 / hw, y0 = p.rows_per_tile ? (m0 % hw) / p.W : 0;
            for (int cb = 0; cb < cblk_tot; cb++) {
                const int src = cb >= p.cblk0, cbl = src ? cb - p.cblk0 : cb;
                for (int kw = 0; kw < 3; kw++, ra.advance(n_sa)) {
                    const uint32_t sa = ra.s;
                    PROF_WAIT(smem_u32(&empty_bar[sa]), ra.ph ^ 1u, w_aempty);
                    const uint32_t dst = smem_u32(ring + (size_t)sa * p.a_slot_bytes);
#ifdef GGML_B200_GEMM_PROFILE
                    if (g_gemm_noload & 1) { if (lane == 0) mbar_arrive(smem_u32(&full_bar[sa])); } else
#endif
                    {
                    mbar_expect_tx_ws(smem_u32(&full_bar[sa]), a_box_bytes);
                    if (src) tma_load_4d_ws(dst, &map_a1, cbl * kBlockK, kw - 1, y0 - 1, img, smem_u32(&full_bar[sa]));
                    else tma_load_4d_ws(dst, &map_a0, cbl * kBlockK, kw - 1, y0 - 1, img, smem_u32(&full_bar[sa]));
                    }
                    for (int kh = 0; kh < 3; kh++, rb.advance(n_sb)) {
                        const uint32_t sb = rb.s;
                        PROF_WAIT(smem_u32(&bempty_bar[sb]), rb.ph ^ 1u, w_bempty);
                        const uint8_t * wsrc = p.w_halo + ((size_t)((cb * 3 + kw) * 3 + kh) * p.n_pad + (size_t)n0) * 128u;
#ifdef GGML_B200_GEMM_PROFILE
                        if (g_gemm_noload & 2) { if (lane == 0) mbar_arrive(smem_u32(&bfull_bar[sb])); } else
#endif
                        {
                        mbar_expect_tx_ws(smem_u32(&bfull_bar[sb]), wb_bytes);
                        bulk_load_1d_ws(smem_u32(ringB + (size_t)sb * wb_bytes), wsrc, wb_bytes, smem_u32(&bfull_bar[sb]));
                        }
                    }
                }
            }
        }
#ifdef GGML_B200_GEMM_PROFILE
        if (lane == 0) {
            unsigned long long * g = g_gemm_prof + 16 * ((p.W == 32 ? 0 : 1) * 2 + (p.C1 > 0));
            atomicAdd(g + 0, (unsigned long long)w_aempty); atomicAdd(g + 1, (unsigned long long)w_bempty);
            atomicAdd(g + 2, (unsigned long long)(clock64() - t_role0)); atomicAdd(g + 3, 1ull);
        }
#else
        (void)w_aempty; (void)w_bempty; (void)t_role0;
#endif
        __syncwarp();
    } else if (warp == 0) {
        // ===================== TMA producer (whole warp, elected lane issues) =====================
        {
            RingPos rp;  // k-block position across all tiles of this CTA
            const uint32_t n_st = (uint32_t)p.stages;
            long long w_empty = 0;
            const long long t_role0 = clock64();
            for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x) {
                const int m0 = tile * tile_m;
                int img = 0, y0 = 0;
                if (p.conv) {
                    const int hw = p.H * p.W;
                    img          = m0 / hw;
                    y0           = p.rows_per_tile ? (m0 % hw) / p.W : 0;
                }
                int r = 0, kw = 0, kh = 0;  // conv: k-block = (channel block r, kw, kh), kh fastest
                for (int kb = 0; kb < p.num_kb; kb++, rp.advance(n_st)) {
                    const int      s  = (int)rp.s;
                    const uint32_t ph = rp.ph;
                    PROF_WAIT(smem_u32(&empty_bar[s]), ph ^ 1u, w_empty);
                    const uint32_t fb = smem_u32(&full_bar[s]);
                    const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
                    const uint32_t sb = sa + a_bytes;
                    if (!p.conv) {
                        mbar_expect_tx_ws(fb, (uint32_t)stage_bytes);
                        tma_load_2d_ws(sa, &map_a0, kb * p.kb_elems, m0, fb);
                        if (!p.b_resident) tma_load_2d_ws(sb, &map_b, kb * p.kb_elems, n0, fb);
                    } else {
                        mbar_expect_tx_ws(fb, (uint32_t)(p.a_tx_bytes + b_bytes));  // the activation box may be shorter than 128 rows
                        // K order = (channel block, kw, kh): the same accumulation order as the halo scheme, so that a tile gives the same
                        // bits whichever of the two schemes (chosen by grid size) computes it
                        const int src = r >= p.cblk0;
                        const int cb  = src ? r - p.cblk0 : r;
                        const int tap = kh * 3 + kw;
                        if (src) tma_load_4d_ws(sa, &map_a1, cb * kBlockK, kw - 1, y0 + kh - 1, img, fb);
                        else tma_load_4d_ws(sa, &map_a0, cb * kBlockK, kw - 1, y0 + kh - 1, img, fb);
                        tma_load_3d_ws(sb, &map_b, (src ? p.C0 : 0) + cb * kBlockK, tap, n0, fb);
                        if (++kh == 3) {
                            kh = 0;
                            if (++kw == 3) { kw = 0; r++; }
                        }
                    }
                }
            }
#ifdef GGML_B200_GEMM_PROFILE
            if (lane == 0 && p.conv == 1 && p.W == 8) {
                atomicAdd(g_gemm_prof1 + 0, (unsigned long long)w_empty); atomicAdd(g_gemm_prof1 + 1, (unsigned long long)(clock64() - t_role0)); atomicAdd(g_gemm_prof1 + 2, 1ull);
            }
#else
            (void)w_empty; (void)t_role0;
#endif
        }
        __syncwarp();
    } else if ((warp == 1 || (warp == 10 && n_issuers == 2)) && p.conv == 2) {
        // ===================== MMA issuer, halo mode =====================
        const uint32_t idesc = make_idesc(p.block_n);
        const uint8_t * ringB = ring + (size_t)p.stages * p.a_slot_bytes;
        const uint32_t wb_bytes = (uint32_t)p.block_n * 128u;
        RingPos ra, rb;
        const uint32_t n_sa = (uint32_t)p.stages, n_sb = (uint32_t)p.b_stages;
        const uint32_t acc_mask = (uint32_t)p.acc_stages - 1u, acc_shift = p.acc_stages == 4 ? 2u : 1u;  // 2 or 4 accumulators
        uint32_t t = 0;
        long long w_tmem = 0, w_afull = 0, w_bfull = 0;
        const long long t_role0 = clock64();
        const int tpb = p.pair ? 2 : 1;  // M tiles per activation box: each weight block feeds tpb accumulators
        const int h   = warp == 1 ? 0 : 1;  // the half of the box this issuer owns
        for (int tile = blockIdx.x * tpb; tile < num_m_tiles; tile += gridDim.x * tpb, t += (uint32_t)tpb) {
            const uint32_t acc = (t + (uint32_t)h) & acc_mask, aph = ((t + (uint32_t)h) >> acc_shift) & 1u;
            PROF_WAIT(smem_u32(&tmem_empty[acc]), aph ^ 1u, w_tmem);
            const uint32_t tmem_d = tmem_base + acc * (uint32_t)p.block_n;
            tc_fence_after();
            uint32_t first = 1;
            for (int cb = 0; cb < cblk_tot; cb++) {
                const int src = cb >= p.cblk0;
                const int rem = src ? p.C1 - (cb - p.cblk0) * kBlockK : p.C0 - cb * kBlockK;
                const int ksteps = rem >= kBlockK ? 4 : (rem + 15) / 16;
                for (int kw = 0; kw < 3; kw++, ra.advance(n_sa)) {
                    const uint32_t sa = ra.s;
                    PROF_WAIT(smem_u32(&full_bar[sa]), ra.ph, w_afull);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(ring + (size_t)sa * p.a_slot_bytes);
                    for (int kh = 0; kh < 3; kh++, rb.advance(n_sb)) {
                        const uint32_t sb = rb.s;
                        PROF_WAIT(smem_u32(&bfull_bar[sb]), rb.ph, w_bfull);
                        tc_fence_after();
                        const uint32_t b_lo = smem_desc_lo(smem_u32(ringB + (size_t)sb * wb_bytes));
                        // tap (kh, kw) of M tile h: the box rows shifted down by kh image rows (+ one tile) = a multiple of the 1 KiB swizzle atom
                        const uint32_t a_lo = smem_desc_lo(a0 + (uint32_t)(kh * p.W + h * tile_m) * 128u);
                        for (int k = 0; k < ksteps; k++)
                            umma_f16_ws_split(tmem_d, a_lo + (uint32_t)(2 * k), b_lo + (uint32_t)(2 * k), smem_desc_hi(128), idesc, (first && k == 0) ? 0u : 1u);
                        first = 0;
                        umma_commit_ws(smem_u32(&bempty_bar[sb]));
                    }
                    umma_commit_ws(smem_u32(&empty_bar[sa]));
                }
            }
            umma_commit_ws(smem_u32(&tmem_full[acc]));
        }
#ifdef GGML_B200_GEMM_PROFILE
        if (lane == 0) {
            unsigned long long * g = g_gemm_prof + 16 * ((p.W == 32 ? 0 : 1) * 2 + (p.C1 > 0));
            atomicAdd(g + 4, (unsigned long long)w_tmem); atomicAdd(g + 5, (unsigned long long)w_afull); atomicAdd(g + 6, (unsigned long long)w_bfull);
            atomicAdd(g + 7, (unsigned long long)(clock64() - t_role0));
        }
#else
        (void)w_tmem; (void)w_afull; (void)w_bfull; (void)t_role0;
#endif
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp, elected lane issues) =====================
        {
            const uint32_t idesc = make_idesc(p.block_n);
            const uint32_t desc_hi = p.kb_elems == 64 ? smem_desc_hi(128) : (p.kb_elems == 32 ? smem_desc_hi(64) : smem_desc_hi(32));
            RingPos rp;
            const uint32_t n_st = (uint32_t)p.stages;
            const uint32_t acc_mask = (uint32_t)p.acc_stages - 1u, acc_shift = p.acc_stages == 4 ? 2u : 1u;  // 2 or 4 accumulators
            uint32_t t = 0;
            long long w_full = 0, w_tmem = 0;
            const long long t_role0 = clock64();
            if (p.b_resident && (int)blockIdx.x < num_m_tiles) mbar_wait(smem_u32(bres_full), 0);
            for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x, t++) {
                const uint32_t acc = t & acc_mask, aph = (t >> acc_shift) & 1u;
                PROF_WAIT(smem_u32(&tmem_empty[acc]), aph ^ 1u, w_tmem);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * (uint32_t)p.block_n;
                int r = 0, t9 = 0;  // conv: channel block of this k-block (9 taps each)
                for (int kb = 0; kb < p.num_kb; kb++, rp.advance(n_st)) {
                    const int      s  = (int)rp.s;
                    const uint32_t ph = rp.ph;
                    PROF_WAIT(smem_u32(&full_bar[s]), ph, w_full);
                    if (p.a_cp_async) fence_proxy_async();  // cp.async wrote through the generic proxy
                    tc_fence_after();
                    const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
                    const uint32_t sb = p.b_resident ? smem_u32(smem_bres + (size_t)kb * b_bytes) : sa + a_bytes;
                    int rem;
                    if (!p.conv) {
                        rem = p.K - kb * p.kb_elems;
                        if (rem > p.kb_elems) rem = p.kb_elems;
                    } else {
                        const int src = r >= p.cblk0;
                        rem           = (src ? p.C1 - (r - p.cblk0) * kBlockK : p.C0 - r * kBlockK);
                        if (++t9 == 9) { t9 = 0; r++; }
                    }
                    const int      ksteps = rem >= kBlockK ? 4 : (rem + 15) / 16;
                    const uint32_t a_lo = smem_desc_lo(sa), b_lo = smem_desc_lo(sb);
                    for (int k = 0; k < ksteps; k++) {
                        // advancing 16 f16 (32 bytes) along K inside the swizzle atom = +2 in the 16-byte address field
                        umma_f16_ws_split(tmem_d, a_lo + (uint32_t)(2 * k), b_lo + (uint32_t)(2 * k), desc_hi, idesc, (kb | k) != 0);
                    }
                    umma_commit_ws(smem_u32(&empty_bar[s]));  // frees the smem stage once these MMAs have read it
                }
                umma_commit_ws(smem_u32(&tmem_full[acc]));  // accumulator complete
            }
#ifdef GGML_B200_GEMM_PROFILE
            if (lane == 0 && p.conv == 1 && p.W == 8) {
                atomicAdd(g_gemm_prof1 + 3, (unsigned long long)w_full); atomicAdd(g_gemm_prof1 + 4, (unsigned long long)w_tmem); atomicAdd(g_gemm_prof1 + 5, (unsigned long long)(clock64() - t_role0));
            }
#else
            (void)w_full; (void)w_tmem; (void)t_role0;
#endif
        }
        __syncwarp();
    } else if (warp < 10) {
        // ===================== epilogue: TMEM -> registers -> swizzled smem -> TMA store =====================
        // A thread owns one output row (TMEM lane).  Writing rows straight to global costs one 16-byte wavefront per
        // lane (32 per instruction) and made output-heavy layers LSU-bound; instead each group stages a
        // 128 x 64-column slab in shared memory (128-byte swizzle: conflict-free) and one thread hands it to the TMA.
        const int      q      = warp & 3;          // TMEM lane quadrant this warp may access
        const uint32_t group  = (warp - 2) >> 2;   // 0: even local tiles / accumulator 0, 1: odd tiles / accumulator 1
        const int      row    = q * 32 + lane;
        const bool     leader = (warp - 2) % 4 == 0 && lane == 0;
        const GemmEpilogue & ep = p.ep;
        constexpr bool GEN  = EPI < 0;
        const bool f_act    = GEN ? ep.act != 0 : (EPI & kEpiAct) != 0;
        const bool f_res32  = GEN ? ep.res32 != nullptr : (EPI & kEpiRes32) != 0;
        const bool f_res16  = GEN ? ep.res16 != nullptr : (EPI & kEpiRes16) != 0;
        const bool f_out16  = GEN ? ep.out16 != nullptr : (EPI & kEpiOut16) != 0;
        const bool f_out32  = GEN ? ep.out32 != nullptr : (EPI & kEpiOut32) != 0;
        const bool f_ln     = GEN ? ep.ln_stats != nullptr : (EPI & kEpiLn) != 0;
        const bool f_stats  = GEN ? ep.stats_out != nullptr : (EPI & kEpiStats) != 0;
        const bool f_warp   = GEN ? p.ep_warp != 0 : !(EPI & kEpiRes32);
        const uint32_t stg16 = smem_u32(stage_base + (size_t)group * (stg16_bytes + stg32_bytes));
        const uint32_t stg32 = stg16 + (uint32_t)stg16_bytes;
        const uint32_t swz   = (uint32_t)(row & 7);
        const uint32_t rbar  = smem_u32(&res_full[group]);
        uint32_t t = 0, ri = 0;
        const bool paired = p.conv == 2 && p.pair;  // local tiles 2i, 2i+1 are the two halves of the CTA's i-th activation box
        for (int tile = paired ? 2 * (int)blockIdx.x : (int)blockIdx.x; tile < num_m_tiles;
             t++, tile = paired ? 2 * (int)(blockIdx.x + (t >> 1) * gridDim.x) + (int)(t & 1u) : (int)(blockIdx.x + t * gridDim.x)) {
            if ((t & 1u) != group) continue;
            const int      m0  = tile * tile_m;
            const int      m   = m0 + row;
            const uint32_t acc = t & ((uint32_t)p.acc_stages - 1u), aph = (t >> (p.acc_stages == 4 ? 2 : 1)) & 1u;  // 2 or 4 accumulators; (t & 1) == group
            // folded LayerNorm: this row's mean and 1/std from the producer's (sum, sum of squares)
            float ln_r = 1.f, ln_mr = 0.f, st_sum = 0.f, st_sq = 0.f;
            if (f_ln && m < p.M) {
                const float2 st  = *reinterpret_cast<const float2 *>(ep.ln_stats + 2 * (size_t)m);
                const float  mu  = st.x * ep.ln_inv_c;
                const float  var = fmaxf(fmaf(st.y, ep.ln_inv_c, -mu * mu), 0.f);
                ln_r             = rsqrtf(var + ep.ln_eps);
                ln_mr            = -mu * ln_r;
            }
            mbar_wait(smem_u32(&tmem_full[acc]), aph);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * (uint32_t)p.block_n + ((uint32_t)(q * 32) << 16);
#ifdef GGML_B200_GEMM_PROFILE
            if (g_gemm_noload & 4) {  // probe: the epilogue hands the accumulator straight back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
                continue;
            }
#endif
            for (int cc = 0; cc < p.block_n; cc += 64) {
                if (n0 + cc >= p.N) break;  // group-uniform
                // the previous slab must have been read out by the TMA before it is overwritten.  Without a residual slab
                // every warp owns its 32 rows end to end (stage, fence, store): no group barrier, four independent store streams
                if (f_warp) {
                    if (lane == 0) tma_store_wait_read();
                    __syncwarp();
                } else {
                    if (leader) tma_store_wait_read();
                    named_bar_sync(1 + group, 128);
                }
                if (f_res32) {
                    // the f32 residual slab(s) of this chunk land in the staging buffer and are updated in place
                    if (leader) {
                        const bool two = cc + 32 < p.block_n && n0 + cc + 32 < p.N;
                        mbar_expect_tx(rbar, (uint32_t)(kBlockM * 128 * (two ? 2 : 1)));
                        tma_load_2d(stg32, &map_r32, n0 + cc, m0, rbar);
                        if (two) tma_load_2d(stg32 + kBlockM * 128, &map_r32, n0 + cc + 32, m0, rbar);
                    }
                }
                bool res_ready = !f_res32;
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    const int c0 = cc + half * 32;
                    if (c0 >= p.block_n || n0 + c0 >= p.N) break;  // group-uniform
                    float v[32];
                    tmem_ld_32x32(tmem_d + (uint32_t)c0, v);
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const int n = n0 + c0 + g * 8;
                        float y[8];
                        {   // per-column scale/shift: 4 broadcast LDS.128 per 8 columns (scalar loads were 2 LDS per element)
                            const uint32_t sa = smem_u32(s_scale + c0 + g * 8), sb = smem_u32(s_shift + c0 + g * 8);
                            const float4 s0 = ld_shared_f4(sa), s1 = ld_shared_f4(sa + 16), h0 = ld_shared_f4(sb), h1 = ld_shared_f4(sb + 16);
                            const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                            const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                            if (f_ln) {  // y = r * (acc - mu * c1[n]), then the usual scale / shift
                                const uint32_t ca = smem_u32(s_c1 + c0 + g * 8);
                                const float4   c0v = ld_shared_f4(ca), c1v = ld_shared_f4(ca + 16);
                                const float    c1[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
#pragma unroll
                                for (int j = 0; j < 8; j++) v[g * 8 + j] = fmaf(c1[j], ln_mr, v[g * 8 + j] * ln_r);
                            }
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                float tt = fmaf(v[g * 8 + j], sc[j], sh[j]);
                                y[j]     = f_act ? silu_h(tt) : tt;  // (scale / shift were halved in the prologue)
                            }
                        }
                        if (f_res32) {
                            if (!res_ready) {
                                mbar_wait(rbar, ri & 1u);
                                res_ready = true;
                            }
                            const uint32_t base = stg32 + (uint32_t)half * (kBlockM * 128) + (uint32_t)row * 128;
                            const float4 r0 = ld_shared_f4(base + (((uint32_t)(2 * g) ^ swz) << 4));
                            const float4 r1 = ld_shared_f4(base + (((uint32_t)(2 * g + 1) ^ swz) << 4));
                            y[0] += r0.x; y[1] += r0.y; y[2] += r0.z; y[3] += r0.w;
                            y[4] += r1.x; y[5] += r1.y; y[6] += r1.z; y[7] += r1.w;
                        }
                        if (m < p.M && n + 8 <= p.N) {
                            if (f_res16) {
                                const uint4    rr = *reinterpret_cast<const uint4 *>(ep.res16 + (size_t)m * ep.ldr16 + n);
                                const __half2 * h = reinterpret_cast<const __half2 *>(&rr);
#pragma unroll
                                for (int j = 0; j < 4; j++) {
                                    const float2 f = __half22float2(h[j]);
                                    y[2 * j] += f.x;
                                    y[2 * j + 1] += f.y;
                                }
                            }
                        }
                        if (f_stats) {  // row statistics of the FINAL values (columns >= N contribute exact zeros)
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                st_sum += y[j];
                                st_sq = fmaf(y[j], y[j], st_sq);
                            }
                        }
                        if (f_out32) {  // slab `half`: row-major 32 floats = 8 x 16 B chunks, chunk index XOR (row % 8)
                            const uint32_t base = stg32 + (uint32_t)half * (kBlockM * 128) + (uint32_t)row * 128;
                            st_shared_v4(base + (((uint32_t)(2 * g) ^ swz) << 4), __float_as_uint(y[0]), __float_as_uint(y[1]),
                                         __float_as_uint(y[2]), __float_as_uint(y[3]));
                            st_shared_v4(base + (((uint32_t)(2 * g + 1) ^ swz) << 4), __float_as_uint(y[4]), __float_as_uint(y[5]),
                                         __float_as_uint(y[6]), __float_as_uint(y[7]));
                        }
                        if (f_out16) {  // 64 halves per row = 8 chunks; this 8-column group is chunk half*4+g
                            uint32_t h[4];
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                __half2 hh = __floats2half2_rn(y[2 * j], y[2 * j + 1]);
                                h[j]       = *reinterpret_cast<uint32_t *>(&hh);
                            }
                            st_shared_v4(stg16 + (uint32_t)row * 128 + (((uint32_t)(half * 4 + g) ^ swz) << 4), h[0], h[1], h[2], h[3]);
                        }
                    }
                }
                ri += f_res32 ? 1u : 0u;
                fence_proxy_async();             // generic-proxy smem writes -> visible to the TMA (async proxy)
                if (f_warp) {
                    __syncwarp();
                    if (lane == 0) {
                        const uint32_t woff = (uint32_t)q * 32u * 128u;
                        if (p.conv) {
                            // conv output maps are 3-D {N, tile_m, tiles}: rows >= tile_m of the 128-row MMA tile are clipped by the TMA
                            if (f_out16) tma_store_3d(&map_o16, stg16 + woff, n0 + cc, q * 32, tile);
                            if (f_out32) {
                                tma_store_3d(&map_o32, stg32 + woff, n0 + cc, q * 32, tile);
                                if (cc + 32 < p.block_n && n0 + cc + 32 < p.N) tma_store_3d(&map_o32, stg32 + kBlockM * 128 + woff, n0 + cc + 32, q * 32, tile);
                            }
                        } else {
                            if (f_out16) tma_store_2d(&map_o16, stg16 + woff, n0 + cc, m0 + q * 32);
                            if (f_out32) {
                                tma_store_2d(&map_o32, stg32 + woff, n0 + cc, m0 + q * 32);
                                if (cc + 32 < p.block_n && n0 + cc + 32 < p.N) tma_store_2d(&map_o32, stg32 + kBlockM * 128 + woff, n0 + cc + 32, m0 + q * 32);
                            }
                        }
                        tma_store_commit();
                    }
                    continue;
                }
                named_bar_sync(1 + group, 128);
                if (leader) {
                    if (f_out16) tma_store_2d(&map_o16, stg16, n0 + cc, m0);
                    if (f_out32) {
                        tma_store_2d(&map_o32, stg32, n0 + cc, m0);
                        if (cc + 32 < p.block_n && n0 + cc + 32 < p.N) tma_store_2d(&map_o32, stg32 + kBlockM * 128, n0 + cc + 32, m0);
                    }
                    tma_store_commit();
                }
            }
            if (f_stats && m < p.M) *reinterpret_cast<float2 *>(ep.stats_out + 2 * (size_t)m) = make_float2(st_sum, st_sq);
            // this warp's TMEM reads are complete (tcgen05.wait::ld inside tmem_ld_32x32): hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
        }
        // the staging slabs must have been READ by the TMA before the CTA (and its shared memory) goes away; the global writes
        // themselves complete by grid end (same contract as CUTLASS's tma_store_wait), no need to wait for them here
        if (f_warp ? lane == 0 : leader) tma_store_wait_read();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
    static encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *                           p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        B200_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
        if (!p || qr != cudaDriverEntryPointSuccess) B200_ABORT("cuTensorMapEncodeTiled is not available in this driver");
        fn = (encode_tiled_fn)p;
    }
    return fn;
}

void tma_encode(CUtensorMap * map, const void * base, CUtensorMapDataType dtype, int rank, const uint64_t * dims,
                const uint64_t * strides_bytes, const uint32_t * box, CUtensorMapSwizzle swizzle) {
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; i++) {
        gdim[i] = dims[i];
        bx[i]   = box[i];
        es[i]   = 1;
    }
    for (int i = 0; i + 1 < rank; i++) gstr[i] = strides_bytes[i];
    CUresult r = get_encode()(map, dtype, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) B200_ABORT("cuTensorMapEncodeTiled failed (%d), rank %d", (int)r, rank);
}

// rank-`rank` f16 tensor map, dims/strides innermost first (strides in bytes for dims 1..rank-1), 128B swizzle
static void make_map(CUtensorMap * map, const void * base, int rank, const uint64_t * dims, const uint64_t * strides_bytes,
                     const uint32_t * box, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT16) {
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; i++) {
        gdim[i] = dims[i];
        bx[i]   = box[i];
        es[i]   = 1;
    }
    for (int i = 0; i + 1 < rank; i++) gstr[i] = strides_bytes[i];
    CUresult r = get_encode()(map, dtype, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, bx, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "libggml_b200: cuTensorMapEncodeTiled failed (%d): rank %d dims", (int)r, rank);
        for (int i = 0; i < rank; i++) fprintf(stderr, " %llu", (unsigned long long)dims[i]);
        fprintf(stderr, " box");
        for (int i = 0; i < rank; i++) fprintf(stderr, " %u", box[i]);
        fprintf(stderr, "\n");
        abort();
    }
}

static void choose_tiling(GemmLaunch & L, int N) {
    GemmLaunch::Params & p = L.p;
    // Output-heavy shapes (K small against N) are epilogue-bound: 128-column tiles let two CTAs (16 epilogue warps)
    // share an SM, and re-reading the small A operand from L2 is cheap.  Otherwise one tile spans N (<= 256) so A is
    // read exactly once.  With several N tiles the tile width is a multiple of the 64-column store slab.
    // Compute-heavy shapes (K >= 512, the batched GRU's recurrent matmul) take 256-wide tiles again: with 128-wide tiles the
    // operand re-reads from L2 (403 MB at 4096 x 3072 x 1024) bound the kernel at ~570 TFLOP/s, 256-wide reaches ~900.
    int max_bn             = (N > 128 && 2 * p.K <= N && p.K < 512) ? 128 : 256;
    if (max_bn == 128 && N % 128 != 0 && N % 192 == 0) max_bn = 192;  // e.g. qkv N = 576: 3 x 192 instead of 4 x 128 + 64 (111 -> 96 us)
    if (p.ep.stats_out) max_bn = 256;  // row statistics need the whole row in one tile (N <= 256 is checked by the caller)
    if (const char * e = getenv("GGML_B200_GEMM_BN")) max_bn = atoi(e);  // tuning probe
    p.n_tiles              = (N + max_bn - 1) / max_bn;
    int per                = (N + p.n_tiles - 1) / p.n_tiles;
    p.block_n              = p.n_tiles > 1 ? (per + 63) / 64 * 64 : (per + 31) / 32 * 32;
    // Small problems (few M tiles: small batches, the last ViT block) are latency chains of K blocks on a handful of SMs:
    // spread them over more SMs by splitting N into 64-column tiles while (M tiles x N tiles) stays below half the SMs.
    if (!p.ep.stats_out && getenv("GGML_B200_GEMM_BN") == nullptr && getenv("GGML_B200_GEMM_NO_NSPLIT") == nullptr) {
        const int mt = (p.M + p.tile_m - 1) / p.tile_m;
        while (p.block_n > 64 && 2 * mt * p.n_tiles <= runtime().sm_count) {
            const int nb = ((p.block_n / 2) + 63) / 64 * 64;
            if (nb >= p.block_n) break;
            p.block_n = nb;
            p.n_tiles = (N + nb - 1) / nb;
        }
    }
    // accumulator stages in TMEM (tile i+1.. accumulate while tile i drains); power-of-two column count >= 32.  Narrow tiles
    // take 4 stages: with one 64-column chunk per tile the MMA turnaround after a release would otherwise be exposed
    p.acc_stages           = (p.block_n <= 64 && getenv("GGML_B200_GEMM_ACC2") == nullptr) ? 4 : 2;
    const int need         = p.acc_stages * p.block_n;
    p.tmem_cols            = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
    const int a_bytes = kBlockM * p.kb_elems * 2, b_bytes = p.block_n * p.kb_elems * 2;
    const int b_total = p.num_kb * b_bytes;
    const int staging = 2 * ((p.ep.out16 ? kBlockM * 128 : 0) + ((p.ep.out32 || p.ep.res32) ? 2 * kBlockM * 128 : 0));
    // TMEM (512 columns) and shared memory (~220 KiB usable) decide how many persistent CTAs share an SM.  The weight tile
    // stays resident (ring = A only) when that does not cost a CTA per SM: one CTA with resident weights was measured slower
    // than two CTAs re-loading them (ffn up-projection 87 -> 99 us), while with equal occupancy it is faster (qkv 96 -> 91 us,
    // expand 16->64 214 -> 129 us).
    // Small grids (fewer CTAs than SMs: small batches, the last stages) are latency chains of k-blocks on a lone CTA per SM: a second
    // CTA slot buys nothing there, so the shared memory goes into a deeper ring instead (conv3x3 at 8x8: 2 -> 6 stages in flight).
    const int  mt_all     = (p.M + p.tile_m - 1) / p.tile_m;
    const bool small_grid = mt_all * p.n_tiles <= runtime().sm_count && getenv("GGML_B200_GEMM_NO_DEEP") == nullptr;
    auto fit = [&](int stage_bytes, int fixed, int & ctas) {
        int stages = 0;
        for (ctas = (p.block_n <= 128 && !small_grid) ? 2 : 1; ctas >= 1; ctas--) {
            const int budget = (216 * 1024) / ctas - 1024 - kCtrlBytes - staging - fixed;
            stages           = budget > 0 ? budget / stage_bytes : 0;
            if (stages >= 2 || ctas == 1) break;
        }
        return stages;
    };
    int       ctas_ring = 1, ctas_res = 1;
    const int st_ring = fit(a_bytes + b_bytes, 0, ctas_ring);
    const int st_res  = fit(a_bytes, b_total, ctas_res);
    p.b_resident      = (!p.conv && st_res >= 2 && ctas_res >= ctas_ring && getenv("GGML_B200_GEMM_NO_BRES") == nullptr) ? 1 : 0;
    const int stage_bytes = a_bytes + (p.b_resident ? 0 : b_bytes);
    const int fixed       = p.b_resident ? b_total : 0;
    int       stages      = p.b_resident ? st_res : st_ring;
    L.ctas_per_sm         = p.b_resident ? ctas_res : ctas_ring;
    if (stages > (small_grid ? kMaxStage : 4)) stages = small_grid ? kMaxStage : 4;
    if (small_grid && stages > p.num_kb && !p.a_cp_async) stages = p.num_kb < 2 ? 2 : p.num_kb;  // one tile per CTA: no more slots than its k-blocks
    if (stages < 1) stages = 1;
    if (const char * e = getenv("GGML_B200_GEMM_STAGES")) stages = atoi(e);      // tuning probes
    if (const char * e = getenv("GGML_B200_GEMM_CTAS")) L.ctas_per_sm = atoi(e);
    p.stages     = stages;
    p.ring_bytes = stages * stage_bytes;
    L.smem_bytes = 1024 + (size_t)fixed + (size_t)stages * stage_bytes + kCtrlBytes + staging;
}

// output tensor maps for the TMA-store epilogue: 128-row x 128-byte boxes, 128B swizzle
static void make_output_maps(GemmLaunch & L) {
    const GemmEpilogue & ep = L.p.ep;
    L.map_o16 = L.map_a0;
    L.map_o32 = L.map_a0;
    L.map_r32 = L.map_a0;
    if (ep.res32) {
        const uint64_t dims[2] = {(uint64_t)L.p.N, (uint64_t)L.p.M};
        const uint64_t str[1]  = {(uint64_t)ep.ldr32 * 4};
        const uint32_t box[2]  = {32, (uint32_t)kBlockM};
        make_map(&L.map_r32, ep.res32, 2, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    }
    L.p.ep_warp = (!ep.res32 && getenv("GGML_B200_GEMM_GROUP_EPILOGUE") == nullptr) ? 1 : 0;
    const uint32_t box_rows = L.p.ep_warp ? 32u : (uint32_t)kBlockM;
    if (L.p.conv) {
        // {N, tile_m, tiles}: M is a multiple of tile_m by construction, and a 32-row warp box that reaches past tile_m is clipped
        const uint64_t tm = (uint64_t)L.p.tile_m, nt = (uint64_t)(L.p.M / L.p.tile_m);
        if (ep.out16) {
            const uint64_t dims[3] = {(uint64_t)L.p.N, tm, nt};
            const uint64_t str[2]  = {(uint64_t)ep.ld16 * 2, tm * (uint64_t)ep.ld16 * 2};
            const uint32_t box[3]  = {64, 32, 1};
            make_map(&L.map_o16, ep.out16, 3, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
        }
        if (ep.out32) {
            const uint64_t dims[3] = {(uint64_t)L.p.N, tm, nt};
            const uint64_t str[2]  = {(uint64_t)ep.ld32 * 4, tm * (uint64_t)ep.ld32 * 4};
            const uint32_t box[3]  = {32, 32, 1};
            make_map(&L.map_o32, ep.out32, 3, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
        }
        return;
    }
    if (ep.out16) {
        const uint64_t dims[2] = {(uint64_t)L.p.N, (uint64_t)L.p.M};
        const uint64_t str[1]  = {(uint64_t)ep.ld16 * 2};
        const uint32_t box[2]  = {64, box_rows};
        make_map(&L.map_o16, ep.out16, 2, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    }
    if (ep.out32) {
        const uint64_t dims[2] = {(uint64_t)L.p.N, (uint64_t)L.p.M};
        const uint64_t str[1]  = {(uint64_t)ep.ld32 * 4};
        const uint32_t box[2]  = {32, box_rows};
        make_map(&L.map_o32, ep.out32, 2, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    }
}

static void choose_grid(GemmLaunch & L) {
    const int num_m_tiles = (L.p.M + L.p.tile_m - 1) / L.p.tile_m / (L.p.conv == 2 && L.p.pair ? 2 : 1);  // pair mode: boxes of two tiles
    int       per_n       = (L.ctas_per_sm * runtime().sm_count) / L.p.n_tiles;
    if (per_n < 1) per_n = 1;
    L.grid = dim3((unsigned)(num_m_tiles < per_n ? num_m_tiles : per_n), (unsigned)L.p.n_tiles, 1);
}

bool gemm_prepare(GemmLaunch & L, const __half * A, int lda, const __half * B, int ldb, int M, int N, int K,
                  const GemmEpilogue & ep) {
    if (M <= 0 || N <= 0 || K <= 0 || (N % 8) || (lda % 8) || (ldb % 8) || (K % 8)) return false;
    if (((uintptr_t)A | (uintptr_t)B) & 15) return false;
    if (ep.stats_out && N > 256) return false;  // row statistics need one tile per row
    L = GemmLaunch();
    GemmLaunch::Params & p = L.p;
    p.M = M; p.N = N; p.K = K;
    p.conv   = 0;
    p.tile_m = kBlockM;
    // TMA cost is per box row: with K = 16 / 32 the rows are 32 / 64 bytes (or mostly out-of-bounds fill in a 64-wide
    // box) and the load runs at 2.2 TB/s (tests/gemm_probe.py).  Those shapes use one exact K block in the 32B / 64B
    // swizzled layout, filled by cp.async from the producer warp.
    // Larger K keeps 64-wide blocks (a partly out-of-bounds last box costs less than nine 32-byte-row boxes).
    // (measured, N=128 K=32: 64-wide box with fill 345 us, exact 64B-swizzle box 367 us, cp.async 395 us -> keep TMA;
    //            N=64  K=16: 310 us / 243 us / 212 us -> cp.async)
    p.kb_elems   = K == 16 ? 16 : 64;
    p.a_cp_async = K == 16 && getenv("GGML_B200_GEMM_NO_CPASYNC") == nullptr;
    p.a_ptr      = A;
    p.lda        = lda;
    p.num_kb = (K + p.kb_elems - 1) / p.kb_elems;
    p.ep     = ep;
    choose_tiling(L, N);
    const CUtensorMapSwizzle swz = p.kb_elems == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.kb_elems == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
        const uint64_t str[1]  = {(uint64_t)lda * 2};
        const uint32_t box[2]  = {(uint32_t)p.kb_elems, (uint32_t)kBlockM};
        tma_encode(&L.map_a0, A, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dims, str, box, swz);
        L.map_a1 = L.map_a0;
    }
    {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
        const uint64_t str[1]  = {(uint64_t)ldb * 2};
        const uint32_t box[2]  = {(uint32_t)p.kb_elems, (uint32_t)p.block_n};
        tma_encode(&L.map_b, B, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dims, str, box, swz);
    }
    make_output_maps(L);
    choose_grid(L);
    return true;
}

size_t conv3x3_pack_halo(const uint16_t * Wt, int OC, int C0, int C1, std::vector<uint8_t> & out) {
    // [channel block (source 0 blocks, then source 1)][kw][kh][n_pad rows][64 channels], rows of 128 B with the 16-byte chunk index
    // XOR (row % 8): exactly what the UMMA descriptor of a 128B-swizzled K-major B tile reads, for ANY run of rows starting at a
    // multiple of 8 -- so every N tiling of the kernel finds its block as one contiguous piece
    const int ict = C0 + C1, cb0 = (C0 + kBlockK - 1) / kBlockK, cb1 = C1 > 0 ? (C1 + kBlockK - 1) / kBlockK : 0;
    const int n_pad = (OC + 63) / 64 * 64;
    out.assign((size_t)(cb0 + cb1) * 9 * n_pad * 128, 0);
    uint16_t * dst = reinterpret_cast<uint16_t *>(out.data());
    for (int cb = 0; cb < cb0 + cb1; cb++) {
        const int src = cb >= cb0, cbl = src ? cb - cb0 : cb, cs = src ? C1 : C0, coff = src ? C0 : 0;
        for (int kw = 0; kw < 3; kw++)
            for (int kh = 0; kh < 3; kh++) {
                uint16_t * blk = dst + (size_t)((cb * 3 + kw) * 3 + kh) * n_pad * 64;
                for (int r = 0; r < OC; r++)
                    for (int c = 0; c < 8; c++)
                        for (int e = 0; e < 8; e++) {
                            const int k = cbl * kBlockK + c * 8 + e;
                            if (k < cs) blk[(size_t)r * 64 + (size_t)((c ^ (r & 7)) * 8 + e)] = Wt[(((size_t)r * 3 + kh) * 3 + kw) * ict + coff + k];
                        }
            }
    }
    return out.size();
}

bool conv3x3_prepare(GemmLaunch & L, const __half * x0, int C0, const __half * x1, int C1, int Nimg, int H, int W,
                     const __half * Wt, int OC, const GemmEpilogue & ep, const uint8_t * Wt_halo) {
    if (Nimg <= 0 || H <= 0 || W <= 0 || OC % 8 || C0 % 8 || C1 % 8 || C0 <= 0) return false;
    if (W > 128 || ep.res32 || ep.res16 || ep.stats_out || ep.ln_stats) return false;
    // An M tile is a whole number of image rows (or of whole images when an image has fewer than 128 pixels) that divides the
    // image (the batch), at most 128 pixels: 128 % W == 0 gives full tiles, other widths leave the last MMA rows unused.
    int box_h, box_n, rows_per_tile;
    if (H * W > 128) {
        box_h = 128 / W;
        while (box_h > 1 && H % box_h) box_h--;
        box_n         = 1;
        rows_per_tile = box_h;
    } else {
        box_h = H;
        box_n = 128 / (H * W);
        while (box_n > 1 && Nimg % box_n) box_n--;
        rows_per_tile = 0;
    }
    // halo mode: one image (or part of one) per tile, map rows of a multiple of 8 pixels (a kh shift must be a whole swizzle atom)
    const bool halo = Wt_halo != nullptr && W % 8 == 0 && W <= 64 && H * W >= 64 && getenv("GGML_B200_CONV_NO_HALO") == nullptr;
    if (halo && H * W <= 128) box_n = 1;
    const int tile_m = box_n * box_h * W;
    L = GemmLaunch();
    GemmLaunch::Params & p = L.p;
    p.M = Nimg * H * W; p.N = OC; p.K = 9 * (C0 + C1);
    p.conv = 1; p.H = H; p.W = W; p.rows_per_tile = rows_per_tile;
    p.tile_m = tile_m; p.a_tx_bytes = tile_m * 128;
    p.C0 = C0; p.C1 = C1;
    p.cblk0  = (C0 + kBlockK - 1) / kBlockK;
    p.cblk1  = C1 > 0 ? (C1 + kBlockK - 1) / kBlockK : 0;
    p.num_kb = 9 * (p.cblk0 + p.cblk1);
    p.kb_elems = kBlockK;
    p.ep     = ep;
    choose_tiling(L, OC);
    const int n_pad = (OC + 63) / 64 * 64;
    // Measured (profiles/README.md, round 2): the halo scheme moves 30 % fewer bytes but runs one CTA per SM, and at large batches the
    // two-CTA per-tap scheme keeps more loads in flight (conv 96->96 at 32x32, batch 256: 95 us per-tap vs 113 us halo).  On small
    // grids (fewer tiles than SMs: the 16x16 / 8x8 stages at batch <= 32, batch-1 latency) the chain of 27-45 dependent k-blocks is what
    // costs, and the halo scheme is 15-25 % faster -- so it is used there.  GGML_B200_CONV_HALO=1 forces it everywhere.
    // Pair mode (round 2): the kernel is bound by the L2 -> SM rate (~42 B/clk/SM chip-wide), i.e. by operand bytes per pixel: 387 KB per
    // 128 pixels at C = N = 96 in the per-tap scheme (9 activation boxes + 9 weight blocks), 276 KB with the halo box, 202 KB when two
    // vertically adjacent tiles share the box (rows 2 * tile_m + 2W) and every weight block (4 accumulators of N <= 128 columns in TMEM).
    // Full 128-pixel tiles of whole image rows, an even number of them per image.
    const bool pair_ok = halo && rows_per_tile > 0 && tile_m == kBlockM && (H / box_h) % 2 == 0 && p.block_n <= 128 && p.n_tiles == 1 &&
                         (2 * tile_m + 2 * W) / W <= 256 && getenv("GGML_B200_CONV_NO_PAIR") == nullptr;
    const bool big_grid = (p.M / tile_m) * p.n_tiles > runtime().sm_count;  // up to 148 tiles the plain halo scheme gives every tile its own SM
    const bool halo_ok = (p.M / tile_m) * p.n_tiles <= runtime().sm_count || (pair_ok && big_grid) || getenv("GGML_B200_CONV_HALO") != nullptr;
    if (halo && halo_ok && p.n_tiles * p.block_n <= n_pad) {
        p.pair = pair_ok && (big_grid || getenv("GGML_B200_CONV_PAIR") != nullptr) ? 1 : 0;
        if (p.pair) {
            p.acc_stages = 4;
            p.tmem_cols  = 4 * p.block_n <= 256 ? 256 : 512;
        }
        // one CTA per SM; shared memory = activation ring (slots of 2W + 128 rows: the MMA of tap kh reads 128 rows from row kh * W)
        // + weight ring + control + epilogue staging
        p.conv         = 2;
        p.w_halo       = Wt_halo;
        p.n_pad        = n_pad;
        p.a_slot_bytes = (2 * W + (p.pair ? 2 : 1) * 128) * 128;
        const int wb_bytes = p.block_n * 128;
        const int staging  = 2 * ((ep.out16 ? kBlockM * 128 : 0) + (ep.out32 ? 2 * kBlockM * 128 : 0));
        const int budget   = 216 * 1024 - 1024 - kCtrlBytes - staging;
        int sa = 3, sb = (budget - sa * p.a_slot_bytes) / wb_bytes;  // (2 activation slots + a deeper weight ring: measured equal)
        if (sb > kMaxStage) sb = kMaxStage;
        if (sb > 6) {  // room to spare: a fourth activation slot
            const int sb4 = (budget - 4 * p.a_slot_bytes) / wb_bytes;
            if (sb4 >= 6) { sa = 4; sb = sb4 > kMaxStage ? kMaxStage : sb4; }
        }
        if (sb < 3 && p.pair) {  // two (larger) activation slots leave room for the weight ring
            sa = 2;
            sb = (budget - sa * p.a_slot_bytes) / wb_bytes;
            if (sb > kMaxStage) sb = kMaxStage;
        }
        if (sb >= 3) {
            p.stages      = sa;
            p.b_stages    = sb;
            p.ring_bytes  = sa * p.a_slot_bytes + sb * wb_bytes;
            L.ctas_per_sm = 1;
            L.smem_bytes  = 1024 + (size_t)p.ring_bytes + kCtrlBytes + staging;
        } else {
            p.conv = 1;
            p.pair = 0;
            choose_tiling(L, OC);  // restores the accumulator stages of the per-tap scheme
        }
    }
    const int box_rows = p.conv == 2 ? (p.pair ? 2 : 1) * box_h + 2 : box_h;
    auto act_map = [&](CUtensorMap * map, const __half * x, int C) {
        const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)Nimg};
        const uint64_t str[3]  = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
        const uint32_t box[4]  = {(uint32_t)kBlockK, (uint32_t)W, (uint32_t)box_rows, (uint32_t)box_n};
        make_map(map, x, 4, dims, str, box);
    };
    act_map(&L.map_a0, x0, C0);
    if (C1 > 0) act_map(&L.map_a1, x1, C1); else L.map_a1 = L.map_a0;
    {
        const int      ict     = C0 + C1;
        const uint64_t dims[3] = {(uint64_t)ict, 9, (uint64_t)OC};
        const uint64_t str[2]  = {(uint64_t)ict * 2, (uint64_t)9 * ict * 2};
        const uint32_t box[3]  = {(uint32_t)kBlockK, 1, (uint32_t)p.block_n};
        make_map(&L.map_b, Wt, 3, dims, str, box);
    }
    make_output_maps(L);
    choose_grid(L);
    return true;
}

template <int EPI>
static void gemm_launch_variant(const GemmLaunch & L, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        B200_CHECK(cudaFuncSetAttribute(k_gemm_tcgen05<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    launch_pdl(k_gemm_tcgen05<EPI>, L.grid, dim3(L.p.conv == 2 && L.p.pair ? kThreads : kThreads - 32), L.smem_bytes, st,  // the 11th warp only exists in pair mode
               L.map_a0, L.map_a1, L.map_b, L.map_o16, L.map_o32, L.map_r32, L.p);
}

#ifdef GGML_B200_GEMM_PROFILE
extern "C" void ggml_b200_debug_gemm_prof(unsigned long long * out64, int reset) {
    if (out64) B200_CHECK(cudaMemcpyFromSymbol(out64, g_gemm_prof, sizeof(unsigned long long) * 64));
    if (out64 && reset == 0) {  // second block: the per-tap counters
        unsigned long long z[16];
        B200_CHECK(cudaMemcpyFromSymbol(z, g_gemm_prof1, sizeof z));
        fprintf(stderr, "per-tap conv at 8x8: CTAs*launches %llu: producer total %.0f clk (waits empty %.0f) | MMA total %.0f (waits full %.0f, tmem %.0f)\n", z[2],
                (double)z[1] / z[2], (double)z[0] / z[2], (double)z[5] / z[2], (double)z[3] / z[2], (double)z[4] / z[2]);
    }
    if (reset >= 16) {
        const int v = reset - 16;
        B200_CHECK(cudaMemcpyToSymbol(g_gemm_noload, &v, sizeof v));
    }
    if (reset) {
        unsigned long long z[64] = {0};
        B200_CHECK(cudaMemcpyToSymbol(g_gemm_prof, z, sizeof z));
        B200_CHECK(cudaMemcpyToSymbol(g_gemm_prof1, z, sizeof(unsigned long long) * 16));
    }
}
#endif

void gemm_launch(const GemmLaunch & L, cudaStream_t st) {
    const GemmEpilogue & ep = L.p.ep;
    const int mask = (ep.act ? kEpiAct : 0) | (ep.res32 ? kEpiRes32 : 0) | (ep.out16 ? kEpiOut16 : 0) | (ep.out32 ? kEpiOut32 : 0) |
                     (ep.ln_stats ? kEpiLn : 0) | (ep.stats_out ? kEpiStats : 0) | (ep.res16 ? kEpiRes16 : 0);
    static const bool generic_only = getenv("GGML_B200_GEMM_GENERIC") != nullptr;
    const bool warp_ok = (L.p.ep_warp != 0) == (ep.res32 == nullptr);  // the specialised kernels derive the store mode from the mask
    if (!generic_only && warp_ok) {
        switch (mask) {
#define EPI_CASE(M) case (M): gemm_launch_variant<(M)>(L, st); return;
            EPI_CASE(kEpiAct | kEpiOut16)                                  // expand 1x1 / 3x3 conv + SiLU, ffn up-projection
            EPI_CASE(kEpiOut16)                                            // reduce 1x1, qkv
            EPI_CASE(kEpiOut16 | kEpiOut32)                                // reduce 1x1 feeding a residual
            EPI_CASE(kEpiOut32)                                            // 1x1 into the transformer (f32 stream)
            EPI_CASE(kEpiRes32 | kEpiOut16)                                // reduce 1x1 + residual
            EPI_CASE(kEpiRes32 | kEpiOut16 | kEpiOut32)
            EPI_CASE(kEpiRes32 | kEpiOut32)                                // attention output / ffn down-projection
            EPI_CASE(kEpiAct | kEpiOut32)                                  // last 1x1 expansion (features)
            EPI_CASE(kEpiLn | kEpiOut16)                                   // qkv over a folded LayerNorm
            EPI_CASE(kEpiLn | kEpiAct | kEpiOut16)                         // ffn up-projection / conv_projection over a folded LayerNorm
            EPI_CASE(kEpiStats | kEpiOut16 | kEpiOut32)                    // producers of a folded LayerNorm's input
            EPI_CASE(kEpiStats | kEpiRes32 | kEpiOut16 | kEpiOut32)
#undef EPI_CASE
            default: break;
        }
    }
    gemm_launch_variant<-1>(L, st);
}

}  // namespace b200
