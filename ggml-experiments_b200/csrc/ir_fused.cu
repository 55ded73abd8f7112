// ir_fused.cu -- K4: the whole inverted-residual block in ONE kernel (inverted_residual_layer::forward, main.cpp:854-870):
//
//     expand 1x1 (+BN+SiLU)  ->  depthwise 3x3, stride 1|2 (+BN+SiLU)  ->  reduce 1x1 (+BN)  [-> + inp]
//
// Unfused, the 4x-expanded activation is written to HBM by the expand GEMM, read and re-written by the depthwise kernel and
// read again by the reduce GEMM: ~60 % of all activation traffic of MobileViT-S and 2.7 of 5.5 ms per
// batch-256 step (profiles/layers_r1.txt).  Here only the block input and the block output touch HBM:
//
//   tile = TH x TW output pixels of one image; halo = IH x IW input pixels (IH = (TH-1)*s+3), P_in = IH*IW rows
//   item = (tile, 64-channel chunk c of the expanded width E); a CTA walks its items as ONE flat sequence
//   TMA        : 4-D box {Cin block, IW, IH, 1} of the block input x (f16 NHWC), halo zero-filled      -> xs[k&1] (UMMA A operand)
//   TMA        : expand weights We[c*64.., Cin] and reduce weights Wr[Cout, c*64..]                     -> we[i&1], wr[i&1]
//   E(i)  tcgen05.mma: exp[mb] (128 halo pixels x 64 ch, f32 in TMEM) = xs[mb] . we^T       for every 128-row block mb
//   epi(i)  compute warps: tcgen05.ld -> BN + SiLU -> 0 outside the image (the depthwise pads the EXPANDED map, main.cpp:784)
//                  -> f16 (rounding point #1 = the depthwise conv's im2col)                            -> es  (pitch 144 B)
//   dw(i)   compute warps: depthwise 3x3 from es (FHFMA, sliding 3-row register window) + BN + SiLU -> f16 (rounding point #2
//                  = the reduce conv's im2col), written in the 128B-swizzled K-major UMMA layout         -> as
//   R(i)  tcgen05.mma: red[mo] (128 output pixels x Cout, f32 in TMEM) += as[mo] . wr^T
//   fin(k)  compute warps, after the last chunk of tile k: tcgen05.ld -> BN (+ f32 residual) -> f16 / f32 global stores
//
// One extra warp (the control warp) issues every TMA and MMA and never computes; the compute warps never issue.  They meet
// only through mbarriers, so in the steady state E(i+1) runs under dw(i) (also across tiles: the first E of the next tile runs
// under the last dw and fin of this one), R(i) under epi(i+1), weight chunks are requested one item ahead and the next tile's halo one tile ahead (two x buffers when shared memory allows):
//
//   control:  ... E(i) | we(i+1) | x(k+1) | wait as_full(i-1) -> R(i-1) | wr(i) | wait es_done(i) -> E(i+1) ...
//   compute:  ... wait exp_full(i) -> epi(i) -> es_done | bar | wait red_done(i-1) -> dw(i) -> as_full | bar | [fin(k) after the last chunk] ...
//
// Rounding points are those of the unfused plan and of ggml (f16 operands at both convolution inputs, f32 accumulation,
// f32 BatchNorm applied to the accumulator), so results agree with the three separate kernels to f32 summation order.
// SiLU(t) = h + h*tanh(h) with h = t/2; the 1/2 is folded into the BatchNorm scale/shift (exact: a power of two).
#include "gemm_tcgen05.h"
#include "internal.h"
#include "pdl.cuh"
#include "ptx_sm100.cuh"

namespace b200 {

using namespace ptx;

namespace {

constexpr int kEPitch = 144;  // bytes per halo pixel in the expanded-chunk buffer: 128 B of data + 16 B so that both the row-per-thread
                              // epilogue stores and the 8-lanes-per-pixel depthwise loads are bank-conflict free without a swizzle
constexpr int kMaxMB  = 5;    // at most 5 x 128 halo rows per tile

__device__ __forceinline__ uint32_t ir_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

// 32 lanes x 8 consecutive f32 columns, no wait (the caller batches several loads before one tcgen05.wait::ld)
__device__ __forceinline__ void tmem_ld_32x8_nowait(uint32_t taddr, float * v) {
    uint32_t * r = reinterpret_cast<uint32_t *>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// CW (32 or 16) consecutive columns, no wait: the epilogue keeps the load of the next 128-row block in flight while it works on
// the current one.  tmem_ld_wait_regs names the destination registers as read-write operands of the wait, so that the compiler
// cannot move a use of them above it.
template <int CW>
__device__ __forceinline__ void tmem_ld_cw_nowait(uint32_t taddr, float (&v)[CW]) {
    uint32_t * r = reinterpret_cast<uint32_t *>(v);
    if (CW == 32) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16 % CW]),
              "=r"(r[17 % CW]), "=r"(r[18 % CW]), "=r"(r[19 % CW]), "=r"(r[20 % CW]), "=r"(r[21 % CW]), "=r"(r[22 % CW]), "=r"(r[23 % CW]), "=r"(r[24 % CW]),
              "=r"(r[25 % CW]), "=r"(r[26 % CW]), "=r"(r[27 % CW]), "=r"(r[28 % CW]), "=r"(r[29 % CW]), "=r"(r[30 % CW]), "=r"(r[31 % CW])
            : "r"(taddr)
            : "memory");
    } else {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr)
            : "memory");
    }
}
template <int CW>
__device__ __forceinline__ void tmem_ld_wait_regs(float (&v)[CW]) {
    uint32_t * r = reinterpret_cast<uint32_t *>(v);
    if (CW == 32) {
        asm volatile("tcgen05.wait::ld.sync.aligned;"
                     : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                       "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16 % CW]), "+r"(r[17 % CW]), "+r"(r[18 % CW]),
                       "+r"(r[19 % CW]), "+r"(r[20 % CW]), "+r"(r[21 % CW]), "+r"(r[22 % CW]), "+r"(r[23 % CW]), "+r"(r[24 % CW]), "+r"(r[25 % CW]), "+r"(r[26 % CW]),
                       "+r"(r[27 % CW]), "+r"(r[28 % CW]), "+r"(r[29 % CW]), "+r"(r[30 % CW]), "+r"(r[31 % CW])
                     :
                     : "memory");
    } else {
        asm volatile("tcgen05.wait::ld.sync.aligned;"
                     : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                       "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                     :
                     : "memory");
    }
}

// acc[0..7] (+)= x[0..7] * w[0..7] for 8 packed halves each: 8 FHFMA (fma.rn.f32.f16: exact f16 x f16 product, one f32 rounding)
template <bool FIRST>
__device__ __forceinline__ void fhfma8(float (&acc)[8], const uint4 & x, const uint4 & w) {
    if (FIRST) {  // first tap: the addend is the constant 0, no accumulator initialisation instruction
#define IR_FHFMA2Z(A0, A1, X, W)                                                                                                        \
    asm("{\n\t.reg .b16 xl, xh, wl, wh;\n\tmov.b32 {xl, xh}, %2;\n\tmov.b32 {wl, wh}, %3;\n\tfma.rn.f32.f16 %0, xl, wl, 0f00000000;\n\t" \
        "fma.rn.f32.f16 %1, xh, wh, 0f00000000;\n\t}"                                                                                 \
        : "=f"(A0), "=f"(A1)                                                                                                          \
        : "r"(X), "r"(W))
        IR_FHFMA2Z(acc[0], acc[1], x.x, w.x);
        IR_FHFMA2Z(acc[2], acc[3], x.y, w.y);
        IR_FHFMA2Z(acc[4], acc[5], x.z, w.z);
        IR_FHFMA2Z(acc[6], acc[7], x.w, w.w);
#undef IR_FHFMA2Z
    } else {
#define IR_FHFMA2(A0, A1, X, W)                                                                                                         \
    asm("{\n\t.reg .b16 xl, xh, wl, wh;\n\tmov.b32 {xl, xh}, %2;\n\tmov.b32 {wl, wh}, %3;\n\tfma.rn.f32.f16 %0, xl, wl, %0;\n\t"     \
        "fma.rn.f32.f16 %1, xh, wh, %1;\n\t}"                                                                                       \
        : "+f"(A0), "+f"(A1)                                                                                                          \
        : "r"(X), "r"(W))
        IR_FHFMA2(acc[0], acc[1], x.x, w.x);
        IR_FHFMA2(acc[2], acc[3], x.y, w.y);
        IR_FHFMA2(acc[4], acc[5], x.z, w.z);
        IR_FHFMA2(acc[6], acc[7], x.w, w.w);
#undef IR_FHFMA2
    }
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
// silu(2h) = h + h * tanh(h): the caller passes h = t/2 (the 1/2 lives in the folded BatchNorm constants)
__device__ __forceinline__ float silu_half(float h) {
#ifdef GGML_B200_SILU_EXACT
    return __fdividef(2.0f * h, 1.0f + __expf(-2.0f * h));
#else
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
#endif
}

// NTC compute threads (256: two CTAs per SM, 512: one) + one control warp; STRIDE of the depthwise convolution.
template <int STRIDE, int NTC>
__global__ void __launch_bounds__(NTC + 32, NTC == 256 ? 2 : 1)
k_ir_fused(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_we, const __grid_constant__ CUtensorMap map_wr,
           const __grid_constant__ CUtensorMap map_o32, const __grid_constant__ CUtensorMap map_o16, const __grid_constant__ CUtensorMap map_r32,
           const IrLaunch::Params p) {
    constexpr int NCW = NTC / 32;  // compute warps
    constexpr int WQ  = NTC / 128; // compute warps per TMEM lane quadrant
    constexpr int CW  = 64 / WQ;   // expanded columns per warp in the expand epilogue (32 or 16)
    extern __shared__ uint8_t ir_smem_raw[];
    uint8_t * smem = ir_smem_raw + ((1024u - (smem_u32(ir_smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bars[12];  // x_full[2], we_full[2], wr_full[2], exp_full, red_done, es_done, as_full
    __shared__ uint32_t tmem_slot;
    const uint32_t sbase  = smem_u32(smem);
    const uint32_t xs     = sbase + p.off_x;
    const uint32_t es     = sbase + p.off_e;
    const uint32_t as     = sbase + p.off_a;
    const uint32_t wes    = sbase + p.off_we;
    const uint32_t wrs    = sbase + p.off_wr;
    // parameter block: se, he, sd, hd [Epad] f32 (already halved, see silu_half), sr, hr [Cout_pad] f32, depthwise weights [9][Epad] f16
    const int      Epad   = p.NC * 64;
    float *        s_se   = reinterpret_cast<float *>(smem + p.off_par);
    float *        s_he   = s_se + Epad;
    float *        s_sd   = s_he + Epad;
    float *        s_hd   = s_sd + Epad;
    float *        s_sr   = s_hd + Epad;
    float *        s_hr   = s_sr + p.Cout_pad;
    __half *       s_dww  = reinterpret_cast<__half *>(s_hr + p.Cout_pad);
    const uint32_t x_full = smem_u32(&bars[0]), we_full = smem_u32(&bars[2]), wr_full = smem_u32(&bars[4]);
    const uint32_t exp_full = smem_u32(&bars[6]), red_done = smem_u32(&bars[7]), es_done = smem_u32(&bars[8]), as_full = smem_u32(&bars[9]);
    const uint32_t res_bar = smem_u32(&bars[10]);  // residual slab landed (reduce epilogue)

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;  // shuffle: provably warp-uniform -> role loops on the uniform datapath

    if (tid == 0) {
        for (int i = 0; i < 8; i++) mbar_init(smem_u32(&bars[i]), 1);
        mbar_init(es_done, NCW);
        mbar_init(as_full, NCW);
        mbar_init(res_bar, 1);
        if (p.out32) tma_prefetch_desc(&map_o32);
        if (p.out16) tma_prefetch_desc(&map_o16);
        if (p.res32) tma_prefetch_desc(&map_r32);
        fence_barrier_init();
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_we);
        tma_prefetch_desc(&map_wr);
    }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)p.tmem_cols);
    // constants -> shared (padded channels get scale 1 / shift 0 / weight 0, so they produce exact zeros end to end)
    for (int i = tid; i < Epad; i += NTC + 32) {
        const bool ok = i < p.E;
        s_se[i] = 0.5f * (ok && p.se ? p.se[i] : 1.f);
        s_he[i] = 0.5f * (ok && p.he ? p.he[i] : 0.f);
        s_sd[i] = 0.5f * (ok && p.sd ? p.sd[i] : 1.f);
        s_hd[i] = 0.5f * (ok && p.hd ? p.hd[i] : 0.f);
    }
    for (int i = tid; i < p.Cout_pad; i += NTC + 32) {
        const bool ok = i < p.Cout;
        s_sr[i] = ok && p.sr ? p.sr[i] : 1.f;
        s_hr[i] = ok && p.hr ? p.hr[i] : 0.f;
    }
    for (int i = tid; i < 9 * Epad; i += NTC + 32) {
        const int k = i / Epad, c = i - k * Epad;
        s_dww[i] = c < p.E ? p.dwW[(size_t)k * p.E + c] : __float2half(0.f);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t tmem_red  = tmem_base + (uint32_t)p.MBI * 64u;

    // tiles of this CTA: blockIdx.x + k * gridDim.x, k < ntk; items i = k * NC + c
    const int ntk    = (int)blockIdx.x < p.ntiles ? (p.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int nitems = ntk * p.NC;
    // optional phase profile (debug probe, p.timing != null): compute thread 0 and the control thread accumulate clock64 deltas
#ifdef GGML_B200_IR_PROFILE
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = 0;
    const bool timing = p.timing != nullptr && (tid == 0 || tid == NTC);
    if (timing) tprev = clock64();
#define IR_TICK(i) do { if (timing) { const long long tn = clock64(); tacc[i] += tn - tprev; tprev = tn; } } while (0)
#else
#define IR_TICK(i) do { } while (0)
#endif

    if (warp == NCW) {
        // ===================== control warp: every TMA and MMA of the CTA (whole warp, elected lane issues: ptx_sm100.cuh "_ws") =====================
        if (nitems > 0) {
            auto issue_we = [&](uint32_t i, int c) {  // expand weights of chunk c -> buffer i & 1
                const uint32_t bar = we_full + 8u * (i & 1u);
                mbar_expect_tx_ws(bar, p.we_bytes);
                for (int kb = 0; kb < p.num_kb; kb++) tma_load_2d_ws(wes + (i & 1u) * p.we_bytes + (uint32_t)kb * p.we_kb_stride, &map_we, kb * 64, c * 64, bar);
            };
            auto issue_wr = [&](uint32_t i, int c) {  // reduce weights of chunk c -> buffer i & 1
                const uint32_t bar = wr_full + 8u * (i & 1u);
                mbar_expect_tx_ws(bar, p.wr_bytes);
                tma_load_2d_ws(wrs + (i & 1u) * p.wr_bytes, &map_wr, c * 64, 0, bar);
            };
            auto issue_x = [&](int k) {  // input halo of this CTA's k-th tile -> buffer k % nxb
                int t = (int)blockIdx.x + k * (int)gridDim.x;
                const int tx = t % p.tiles_x; t /= p.tiles_x;
                const int ty = t % p.tiles_y;
                const int n  = t / p.tiles_y;
                const uint32_t xb  = (uint32_t)(k % p.nxb);
                const uint32_t bar = x_full + 8u * xb;
                mbar_expect_tx_ws(bar, p.x_tx_bytes);
                for (int kb = 0; kb < p.num_kb; kb++)
                    tma_load_4d_ws(xs + xb * p.x_buf_stride + (uint32_t)kb * p.x_kb_stride, &map_x, kb * 64, tx * p.TW * STRIDE - 1, ty * p.TH * STRIDE - 1, n, bar);
            };
            const uint32_t idesc_e = ir_idesc(64), idesc_r = ir_idesc(p.Cout_pad);
            auto issue_expand = [&](uint32_t i, int k) {
                const uint32_t wb = wes + (i & 1u) * p.we_bytes;
                const uint32_t xb = xs + (uint32_t)(k % p.nxb) * p.x_buf_stride;
                for (int mb = 0; mb < p.MBI; mb++) {
                    for (int kb = 0; kb < p.num_kb; kb++) {
                        const uint64_t adesc = make_smem_desc(xb + (uint32_t)kb * p.x_kb_stride + (uint32_t)mb * 128u * (uint32_t)p.row_bytes, (uint32_t)p.row_bytes);
                        const uint64_t bdesc = make_smem_desc(wb + (uint32_t)kb * p.we_kb_stride, (uint32_t)p.row_bytes);
                        for (int ks = 0; ks < p.ksteps; ks++)
                            umma_f16_ws(tmem_base + (uint32_t)mb * 64u, adesc + (uint64_t)(2 * ks), bdesc + (uint64_t)(2 * ks), idesc_e, (kb | ks) != 0);
                    }
                }
                umma_commit_ws(exp_full);
            };
            auto issue_reduce = [&](uint32_t i, int c) {
                const uint64_t bdesc = make_smem_desc(wrs + (i & 1u) * p.wr_bytes, 128);
                for (int mo = 0; mo < p.MBO; mo++) {
                    const uint64_t adesc = make_smem_desc(as + (uint32_t)mo * 16384u, 128);
#pragma unroll
                    for (int ks = 0; ks < 4; ks++)
                        umma_f16_ws(tmem_red + (uint32_t)mo * (uint32_t)p.red_stride, adesc + (uint64_t)(2 * ks), bdesc + (uint64_t)(2 * ks), idesc_r, (c | ks) != 0);
                }
                umma_commit_ws(red_done);
            };
            // weights are constants: requested before the PDL wait, so they overlap the tail of the previous kernel
            issue_we(0, 0);
            issue_wr(0, 0);
            pdl_wait();
            issue_x(0);
            uint32_t i = 0;
            for (int k = 0; k < ntk; k++) {
                for (int c = 0; c < p.NC; c++, i++) {
                    const bool first = c == 0, has_next = (int)i + 1 < nitems;
                    IR_TICK(0);  // control: issue work
                    if (i > 0) {
                        mbar_wait(es_done, (i - 1) & 1u);  // epi(i-1) has drained the expand accumulators; E(i-1) is complete
                        tc_fence_after();
                    }
                    IR_TICK(1);  // control: wait es_done
                    if (first && p.nxb == 1 && k > 0) {
                        // single x buffer: it was free only now.  Its load takes about as long as dw(i-1), so R(i-1) goes first.
                        issue_x(k);
                        mbar_wait(as_full, (i - 1) & 1u);
                        mbar_wait(wr_full + 8u * ((i - 1) & 1u), ((i - 1) >> 1) & 1u);
                        tc_fence_after();
                        issue_reduce(i - 1, p.NC - 1);
                        issue_wr(i, 0);
                    }
                    IR_TICK(0);
                    if (first) mbar_wait(x_full + 8u * (uint32_t)(k % p.nxb), (uint32_t)((k / p.nxb) & 1));
                    IR_TICK(2);  // control: wait x
                    mbar_wait(we_full + 8u * (i & 1u), (i >> 1) & 1u);
                    IR_TICK(3);  // control: wait we
                    tc_fence_after();
                    issue_expand(i, k);
                    IR_TICK(6);  // control: expand MMA issue
                    if (has_next) issue_we(i + 1, c + 1 < p.NC ? c + 1 : 0);  // buffer (i+1)&1: last read by E(i-1), complete
                    if (first && p.nxb == 2 && k + 1 < ntk) issue_x(k + 1);  // buffer (k+1)&1: last read by the E's of tile k-1, complete
                    IR_TICK(7);  // control: TMA issue
                    if (i > 0 && !(first && p.nxb == 1 && k > 0)) {
                        IR_TICK(0);
                        mbar_wait(as_full, (i - 1) & 1u);  // dw(i-1) has written `as`
                        IR_TICK(4);  // control: wait as_full
                        mbar_wait(wr_full + 8u * ((i - 1) & 1u), ((i - 1) >> 1) & 1u);
                        IR_TICK(5);  // control: wait wr
                        tc_fence_after();
                        issue_reduce(i - 1, c == 0 ? p.NC - 1 : c - 1);
                        IR_TICK(6);
                        issue_wr(i, c);  // buffer i&1: last read by R(i-2), complete (the compute warps waited for it before dw(i-1))
                        IR_TICK(7);
                    }
                }
            }
            mbar_wait(as_full, (i - 1) & 1u);
            mbar_wait(wr_full + 8u * ((i - 1) & 1u), ((i - 1) >> 1) & 1u);
            tc_fence_after();
            issue_reduce(i - 1, p.NC - 1);
        } else {
            pdl_wait();
        }
        __syncwarp();
        pdl_trigger();
    } else {
        // ===================== compute warps =====================
        pdl_wait();  // the reduce epilogue writes global memory the previous kernel may still read
        pdl_trigger();
        const int quad = warp & 3, wsub = warp >> 2;
        // depthwise: thread = (8-channel group cg, output column xl, row split rs)
        const int cg   = tid & 7;
        const int slot = tid >> 3;
        const int xl   = slot % p.TW;
        const int rs   = slot / p.TW;
        const int RS   = (NTC / 8) / p.TW;
        const int rows_per = p.TH / RS;
        const uint32_t epitch = (uint32_t)p.IW * kEPitch;
        // number of 128-row blocks in which this warp's 32 lanes hold halo rows (warp-uniform)
        int nblk = 0;
        for (int mb = 0; mb < p.MBI; mb++) nblk += (mb * 128 + quad * 32 < p.P_in) ? 1 : 0;

        // ---- reduce epilogue of one finished tile: TMEM -> BN (+ residual) -> 32-channel slabs in shared memory -> TMA stores.
        // A thread owns one accumulator ROW; storing from registers would write 32 scattered 16-byte pieces per instruction (measured:
        // 15 k cycles per tile, LSU-bound).  Instead the rows go into swizzled staging slabs (the es buffer is free between the last
        // depthwise of a tile and the next expand epilogue): [pixel][32 f32] for the f32 copy, [pixel][32 f16] for the f16 copy, and one
        // thread hands each slab to the TMA as a {32 ch, TW, TH} box -- full lines, asynchronous, clipped at the image border.  The f32
        // residual slab is fetched by the TMA into the same staging slab first and updated in place. ----
        const uint32_t st32 = es, st16 = es + p.st16_off;
        uint32_t res_cnt = 0;
        auto final_epilogue = [&](int tile, uint32_t red_parity) {
            int t = tile;
            const int tx = t % p.tiles_x; t /= p.tiles_x;
            const int ty = t % p.tiles_y;
            const int n  = t / p.tiles_y;
            const int nslab = (p.Cout + 31) / 32;
            mbar_wait(red_done, red_parity);
            tc_fence_after();
            IR_TICK(4);
            for (int sl = 0; sl < nslab; sl++) {
                if (p.res32) {
                    if (tid == 0) {
                        mbar_expect_tx(res_bar, (uint32_t)p.P_out * 128u);
                        tma_load_4d(st32, &map_r32, sl * 32, tx * p.TW, ty * p.TH, n, res_bar);
                    }
                    mbar_wait(res_bar, res_cnt & 1u);
                    res_cnt++;
                }
                // 16-column units of this slab: (128-row block mo, half hf); warp wsub takes every WQ-th unit of its lane quadrant
                for (int u = wsub; u < p.MBO * 2; u += WQ) {
                    const int mo = u >> 1, hf = u & 1;
                    const int qi = mo * 128 + quad * 32 + lane;
                    float v[16];
                    __syncwarp();
                    tmem_ld_cw_nowait<16>(tmem_red + (uint32_t)mo * (uint32_t)p.red_stride + ((uint32_t)(quad * 32) << 16) + (uint32_t)(sl * 32 + hf * 16), v);
                    tmem_ld_wait_regs<16>(v);
                    if (qi < p.P_out) {
                        const uint32_t row32 = st32 + (uint32_t)qi * 128u, row16 = st16 + (uint32_t)qi * 64u;
                        const uint32_t sw32 = (uint32_t)qi & 7u, sw16 = ((uint32_t)qi >> 1) & 3u;
                        uint32_t h[8];
#pragma unroll
                        for (int g4 = 0; g4 < 4; g4++) {
                            const int nn = sl * 32 + hf * 16 + g4 * 4;  // s_sr / s_hr hold 1 / 0 up to Cout_pad; columns beyond Cout are clipped by the TMA
                            const float4 sc = *reinterpret_cast<const float4 *>(s_sr + (nn < p.Cout_pad ? nn : 0));
                            const float4 sh = *reinterpret_cast<const float4 *>(s_hr + (nn < p.Cout_pad ? nn : 0));
                            float4 y = make_float4(fmaf(v[g4 * 4], sc.x, sh.x), fmaf(v[g4 * 4 + 1], sc.y, sh.y), fmaf(v[g4 * 4 + 2], sc.z, sh.z), fmaf(v[g4 * 4 + 3], sc.w, sh.w));
                            const uint32_t a32 = row32 + ((((uint32_t)(hf * 4 + g4)) ^ sw32) << 4);
                            if (p.res32) {
                                const float4 rr = ld_shared_f4(a32);
                                y.x += rr.x; y.y += rr.y; y.z += rr.z; y.w += rr.w;
                            }
                            if (p.out32) st_shared_v4(a32, __float_as_uint(y.x), __float_as_uint(y.y), __float_as_uint(y.z), __float_as_uint(y.w));
                            h[2 * g4]     = pack_h2(y.x, y.y);
                            h[2 * g4 + 1] = pack_h2(y.z, y.w);
                        }
                        if (p.out16) {
                            st_shared_v4(row16 + ((((uint32_t)(hf * 2))     ^ sw16) << 4), h[0], h[1], h[2], h[3]);
                            st_shared_v4(row16 + ((((uint32_t)(hf * 2 + 1)) ^ sw16) << 4), h[4], h[5], h[6], h[7]);
                        }
                    }
                }
                if (sl + 1 == nslab) tc_fence_before();
                fence_proxy_async();     // generic-proxy writes of the slabs -> visible to the TMA
                named_bar_sync(1, NTC);
                if (tid == 0) {
                    if (p.out32) tma_store_4d(&map_o32, st32, sl * 32, tx * p.TW, ty * p.TH, n);
                    if (p.out16) tma_store_4d(&map_o16, st16, sl * 32, tx * p.TW, ty * p.TH, n);
                    tma_store_commit();
                    tma_store_wait_read();  // the slabs may be overwritten (next slab, next expand epilogue); the global writes complete by grid end
                }
                named_bar_sync(1, NTC);
            }
        };

        uint32_t i = 0;
        for (int k = 0; k < ntk; k++) {
            const int tile = (int)blockIdx.x + k * (int)gridDim.x;
            int t = tile;
            const int tx = t % p.tiles_x; t /= p.tiles_x;
            const int ty = t % p.tiles_y;
            const int y0 = ty * p.TH * STRIDE - 1, x0 = tx * p.TW * STRIDE - 1;  // image coordinates of halo pixel (0,0)
            // per 128-row block: where this lane's halo pixel goes in `es`, whether it exists and whether it lies inside the image
            const uint32_t dst0 = es + (uint32_t)(quad * 32 + lane) * kEPitch + (uint32_t)(wsub * CW) * 2u;  // block mb: + mb * 128 rows
            uint32_t inimg = 0, valid = 0;
#pragma unroll
            for (int mb = 0; mb < kMaxMB; mb++) {
                const int row = mb * 128 + quad * 32 + lane;
                const int yy = row / p.IW, xx = row - yy * p.IW;
                const int gy = y0 + yy, gx = x0 + xx;
                if (row < p.P_in) valid |= 1u << mb;
                if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) inimg |= 1u << mb;
            }
            for (int c = 0; c < p.NC; c++, i++) {
                // ---- expand epilogue: TMEM -> BN + SiLU -> mask -> f16 -> es ----
                IR_TICK(7);  // compute: rest (barriers, bookkeeping)
                mbar_wait(exp_full, i & 1u);
                tc_fence_after();
                IR_TICK(0);  // compute: wait exp_full
                {
                    // rounds of 16 columns x 32 rows (one tcgen05.ld.x16 each); the load of round r+1 is in flight while round r is
                    // processed (two 16-register buffers)
                    constexpr int HPB = CW / 16;  // 16-column halves per 128-row block for this warp
                    const uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(wsub * CW);
                    const int      col0 = c * 64 + wsub * CW;
                    const int      nround = nblk * HPB;
                    auto process = [&](float (&v)[16], int rd) {
                        const int mb = rd / HPB, hf = rd % HPB;
                        if (!((valid >> mb) & 1u)) return;
                        const bool in = (inimg >> mb) & 1u;
                        const uint32_t d = dst0 + (uint32_t)mb * (128u * kEPitch) + (uint32_t)hf * 32u;
#pragma unroll
                        for (int g8 = 0; g8 < 2; g8++) {
                            uint32_t o[4] = {0u, 0u, 0u, 0u};
                            if (in) {
                                const uint32_t sa = smem_u32(s_se + col0 + hf * 16 + g8 * 8), sb = smem_u32(s_he + col0 + hf * 16 + g8 * 8);
                                const float4 s0 = ld_shared_f4(sa), s1 = ld_shared_f4(sa + 16), h0 = ld_shared_f4(sb), h1 = ld_shared_f4(sb + 16);
                                o[0] = pack_h2(silu_half(fmaf(v[g8 * 8 + 0], s0.x, h0.x)), silu_half(fmaf(v[g8 * 8 + 1], s0.y, h0.y)));
                                o[1] = pack_h2(silu_half(fmaf(v[g8 * 8 + 2], s0.z, h0.z)), silu_half(fmaf(v[g8 * 8 + 3], s0.w, h0.w)));
                                o[2] = pack_h2(silu_half(fmaf(v[g8 * 8 + 4], s1.x, h1.x)), silu_half(fmaf(v[g8 * 8 + 5], s1.y, h1.y)));
                                o[3] = pack_h2(silu_half(fmaf(v[g8 * 8 + 6], s1.z, h1.z)), silu_half(fmaf(v[g8 * 8 + 7], s1.w, h1.w)));
                            }
                            st_shared_v4(d + (uint32_t)g8 * 16u, o[0], o[1], o[2], o[3]);
                        }
                    };
                    auto taddr = [&](int rd) { return tcol + (uint32_t)(rd / HPB) * 64u + (uint32_t)(rd % HPB) * 16u; };
                    float va[16], vb[16];
                    if (nround > 0) tmem_ld_cw_nowait<16>(taddr(0), va);
                    for (int rd = 0; rd < nround; rd += 2) {  // nround is warp-uniform
                        tmem_ld_wait_regs<16>(va);
                        if (rd + 1 < nround) tmem_ld_cw_nowait<16>(taddr(rd + 1), vb);
                        process(va, rd);
                        if (rd + 1 < nround) {
                            tmem_ld_wait_regs<16>(vb);
                            if (rd + 2 < nround) tmem_ld_cw_nowait<16>(taddr(rd + 2), va);
                            process(vb, rd + 1);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(es_done);  // this warp no longer reads the expand accumulators: E(i+1) may overwrite them
                IR_TICK(1);  // compute: expand epilogue
                named_bar_sync(1, NTC);               // es complete
                IR_TICK(2);  // compute: barrier after the epilogue


                // ---- depthwise 3x3 + BN + SiLU: es -> as (UMMA A operand, 128B swizzle) ----
                if (i > 0) mbar_wait(red_done, (i - 1) & 1u);  // R(i-1) has read `as`
                IR_TICK(4);  // compute: wait red_done
                {
                    // Register budget: two CTAs of 9 warps per SM leave 96 registers per thread.  The nine packed weight vectors (36) and the
                    // BatchNorm constants (16) stay in registers; a full 3-row sliding window (36 more) made the kernel spill.
                    const int ch = c * 64 + cg * 8;
                    uint4 w[9];
#pragma unroll
                    for (int kk = 0; kk < 9; kk++) w[kk] = *reinterpret_cast<const uint4 *>(s_dww + (size_t)kk * Epad + ch);
                    float sc[8], sh[8];
                    {
                        const float4 a0 = *reinterpret_cast<const float4 *>(s_sd + ch), a1 = *reinterpret_cast<const float4 *>(s_sd + ch + 4);
                        const float4 b0 = *reinterpret_cast<const float4 *>(s_hd + ch), b1 = *reinterpret_cast<const float4 *>(s_hd + ch + 4);
                        sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
                        sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
                    }
                    const int oyl0 = rs * rows_per;
                    uint32_t  rp   = es + (uint32_t)cg * 16u + (uint32_t)(xl * STRIDE) * kEPitch + (uint32_t)(oyl0 * STRIDE) * epitch;
                    uint32_t  ap   = as + (uint32_t)(oyl0 * p.TW + xl) * 128u;  // A-operand row of the first output; the swizzle is applied per store
                    const uint32_t astep = (uint32_t)p.TW * 128u;
                    // every output gathers its nine taps straight from shared memory (9 LDS.128).  Measured alternatives: a 3-row sliding
                    // register window (3 LDS.128 per output, +36 registers) spills at 96 registers; sharing rows between two vertically
                    // adjacent outputs (6 LDS.128 per output) is 4 % slower than this plain gather.
                    for (int r = 0; r < rows_per; r++) {
                        float acc[8];
#pragma unroll
                        for (int kh = 0; kh < 3; kh++) {
                            const uint32_t rr = rp + (uint32_t)kh * epitch;
                            const uint4 x0 = lds128(rr), x1 = lds128(rr + kEPitch), x2 = lds128(rr + 2 * kEPitch);
                            if (kh == 0) fhfma8<true>(acc, x0, w[0]); else fhfma8<false>(acc, x0, w[3 * kh]);
                            fhfma8<false>(acc, x1, w[3 * kh + 1]);
                            fhfma8<false>(acc, x2, w[3 * kh + 2]);
                        }
                        uint32_t o[4];
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            o[j] = pack_h2(silu_half(fmaf(acc[2 * j], sc[2 * j], sh[2 * j])), silu_half(fmaf(acc[2 * j + 1], sc[2 * j + 1], sh[2 * j + 1])));
                        st_shared_v4(ap + (((uint32_t)cg ^ ((ap >> 7) & 7u)) << 4), o[0], o[1], o[2], o[3]);
                        rp += (uint32_t)STRIDE * epitch;
                        ap += astep;
                    }
                }
                fence_proxy_async();  // generic-proxy writes of `as` -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(as_full);  // R(i) may be issued once every compute warp has arrived
                IR_TICK(5);  // compute: depthwise
                named_bar_sync(1, NTC);               // every warp is done reading es: the next epilogue may overwrite it
                IR_TICK(6);  // compute: barrier after the depthwise
            }
            IR_TICK(7);
            final_epilogue(tile, (i - 1) & 1u);  // waits for R of the tile's last chunk
            IR_TICK(3);  // compute: reduce epilogue
        }
    }
#ifdef GGML_B200_IR_PROFILE
    if (timing)
        for (int j = 0; j < 8; j++) p.timing[((size_t)blockIdx.x * 2 + (tid == 0 ? 0 : 1)) * 8 + j] = tacc[j];
#endif
#undef IR_TICK
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

int env_int(const char * name, int dflt) {
    const char * e = getenv(name);
    return e ? atoi(e) : dflt;
}

}  // namespace

// one tile configuration (TH x TW outputs, NT compute threads); *per_sm_out = CTAs per SM it allows
static bool ir_prepare_cfg(IrLaunch & L, const __half * x, int N, int H, int W, int Cin, int E, int Cout, int stride, const __half * We,
                           const float * se, const float * he, const __half * dwW, const float * sd, const float * hd, const __half * Wr,
                           const float * sr, const float * hr, const float * res32, __half * out16, float * out32, int TH, int TW, int NT, int * per_sm_out) {
    if (Cin % 8 || E % 8 || Cout % 8 || Cin < 8 || Cin > 128 || E > 1024 || Cout > 256 || (stride != 1 && stride != 2) || H % stride || W % stride) return false;
    L = IrLaunch();
    IrLaunch::Params & p = L.p;
    p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.E = E; p.Cout = Cout; p.stride = stride;
    p.Cout_pad = (Cout + 15) / 16 * 16;
    p.OH = H / stride; p.OW = W / stride;
    p.NC = (E + 63) / 64;
    // K blocking of the expand GEMM: exact 32 B / 64 B rows for Cin = 16 / 32 (swizzle-32B / 64B), otherwise 64-channel blocks
    p.kb_elems  = Cin <= 16 ? 16 : (Cin <= 32 ? 32 : 64);
    if (env_int("GGML_B200_IR_KB64", 0)) p.kb_elems = 64;  // probe: 128-byte rows with TMA zero fill for every Cin
    p.row_bytes = p.kb_elems * 2;
    p.num_kb    = (Cin + 63) / 64;
    p.ksteps    = p.kb_elems / 16;
    while (TW > 1 && TW / 2 >= p.OW) TW /= 2;
    if (NT != 256 && NT != 512) return false;
    if ((NT / 8) % TW) return false;
    const int RS = (NT / 8) / TW;
    if (TH % RS) TH = (TH + RS - 1) / RS * RS;
    p.TH = TH; p.TW = TW; p.nthreads = NT;
    p.IH = (TH - 1) * stride + 3; p.IW = (TW - 1) * stride + 3;
    p.P_in = p.IH * p.IW; p.P_out = TH * TW;
    p.MBI = (p.P_in + 127) / 128; p.MBO = (p.P_out + 127) / 128;
    if (p.MBI > kMaxMB) return false;
    if (TW > 256 || TH > 256) return false;                                    // TMA box of the reduce epilogue
    p.tiles_x = (p.OW + TW - 1) / TW; p.tiles_y = (p.OH + TH - 1) / TH;
    p.ntiles  = N * p.tiles_x * p.tiles_y;
    p.red_stride = (p.Cout_pad + 31) / 32 * 32;
    const int need = p.MBI * 64 + p.MBO * p.red_stride;
    if (need > 512 || p.IW > 256 || p.IH > 256) return false;
    p.tmem_cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
    // shared memory layout (offsets from a 1 KiB aligned base); tried with two x buffers first, then with one
    auto up = [](uint32_t v, uint32_t a) { return (v + a - 1) / a * a; };
    const uint32_t Epad = (uint32_t)p.NC * 64u;
    auto layout = [&](int nxb) {
        p.nxb = nxb;
        const uint32_t rows32 = up((uint32_t)p.P_in, 32);
        p.x_kb_stride  = up(rows32 * (uint32_t)p.row_bytes, 1024);
        p.x_buf_stride = (uint32_t)p.num_kb * p.x_kb_stride;
        p.x_tx_bytes   = (uint32_t)p.num_kb * (uint32_t)p.P_in * (uint32_t)p.row_bytes;
        uint32_t off   = 0;
        p.off_x = off; off += (uint32_t)nxb * p.x_buf_stride;
        // the last 128-row MMA block reads past P_in: those rows must only be addressable (their results are never used), so the
        // buffers that follow double as that slack
        const uint32_t x_span = (uint32_t)(nxb - 1) * p.x_buf_stride + (uint32_t)(p.num_kb - 1) * p.x_kb_stride + (uint32_t)p.MBI * 128u * (uint32_t)p.row_bytes;
        // `as`: P_out rows are written, the MMA reads whole 128-row blocks (the tail only has to be addressable)
        p.off_a = up(off, 1024); off = p.off_a + up((uint32_t)p.P_out * 128u, 1024);
        const uint32_t a_span = p.off_a + (uint32_t)p.MBO * 16384u;
        p.we_kb_stride = up(64u * (uint32_t)p.row_bytes, 1024);
        p.we_bytes     = (uint32_t)p.num_kb * p.we_kb_stride;
        p.off_we = up(off, 1024); off = p.off_we + 2 * p.we_bytes;
        p.wr_bytes = (uint32_t)p.Cout_pad * 128u;  // Cout_pad is a multiple of 16: the buffers stay 1 KiB aligned (2 KiB granules)
        p.off_wr = up(off, 1024); off = p.off_wr + 2 * p.wr_bytes;
        // es doubles as the staging of the reduce epilogue: [P_out][32 f32] + [P_out][32 f16], both 1 KiB aligned (TMA swizzle atoms)
        p.st16_off = up((uint32_t)p.P_out * 128u, 1024);
        const uint32_t e_bytes = (uint32_t)p.P_in * kEPitch, st_bytes = p.st16_off + (uint32_t)p.P_out * 64u;
        p.off_e = up(off, 1024); off = p.off_e + (e_bytes > st_bytes ? e_bytes : st_bytes);
        p.off_par = up(off, 16);
        off = p.off_par + (4 * Epad + 2 * (uint32_t)p.Cout_pad) * 4 + 9 * Epad * 2;
        if (off < p.off_x + x_span) off = p.off_x + x_span;
        if (off < a_span) off = a_span;
        return (size_t)1024 + off;
    };
    const size_t two = layout(2);
    // two x buffers unless they cost the second CTA per SM (or do not fit at all)
    const bool two_ok = two <= 225 * 1024 && (NT != 256 || p.tmem_cols > 256 || 2 * (two + 1024) <= 227 * 1024 || 2 * (layout(1) + 1024) > 227 * 1024);
    L.smem_bytes = two_ok ? layout(2) : layout(1);
    if (env_int("GGML_B200_IR_NXB", 0)) L.smem_bytes = layout(env_int("GGML_B200_IR_NXB", 0));
    if (L.smem_bytes > 225 * 1024) return false;
    int per_sm = (NT == 256 && 2 * (L.smem_bytes + 1024) <= 227 * 1024 && p.tmem_cols <= 256) ? 2 : 1;
    if (per_sm == 2 && L.smem_bytes < 78 * 1024) L.smem_bytes = 78 * 1024;  // never three CTAs on an SM: 3 x 256 TMEM columns do not exist
    if (per_sm == 1 && L.smem_bytes < 116 * 1024) L.smem_bytes = 116 * 1024;
    const int cap = per_sm * runtime().sm_count;
    L.grid = p.ntiles < cap ? p.ntiles : cap;
    if (per_sm_out) *per_sm_out = per_sm;
    p.dwW = dwW; p.se = se; p.he = he; p.sd = sd; p.hd = hd; p.sr = sr; p.hr = hr; p.res32 = res32; p.out16 = out16; p.out32 = out32;
    if (!x) return true;  // shape query only (fuse.cpp asks before it commits to the fusion)
    {
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        const uint64_t str[3]  = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
        const uint32_t box[4]  = {(uint32_t)p.kb_elems, (uint32_t)p.IW, (uint32_t)p.IH, 1};
        const CUtensorMapSwizzle swz = p.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        tma_encode(&L.map_x, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dims, str, box, swz);
        const uint64_t wdims[2] = {(uint64_t)Cin, (uint64_t)E};
        const uint64_t wstr[1]  = {(uint64_t)Cin * 2};
        const uint32_t wbox[2]  = {(uint32_t)p.kb_elems, 64};
        tma_encode(&L.map_we, We, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, wdims, wstr, wbox, swz);
    }
    {
        const uint64_t dims[2] = {(uint64_t)E, (uint64_t)Cout};
        const uint64_t str[1]  = {(uint64_t)E * 2};
        const uint32_t box[2]  = {64, (uint32_t)p.Cout_pad};
        tma_encode(&L.map_wr, Wr, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    }
    {   // reduce epilogue: {32 channels, TW, TH, 1} boxes of the outputs (and of the f32 residual), clipped at the image border
        const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)p.OW, (uint64_t)p.OH, (uint64_t)N};
        const uint32_t box[4]  = {32, (uint32_t)p.TW, (uint32_t)p.TH, 1};
        const uint64_t s32[3]  = {(uint64_t)Cout * 4, (uint64_t)p.OW * Cout * 4, (uint64_t)p.OH * p.OW * Cout * 4};
        const uint64_t s16[3]  = {(uint64_t)Cout * 2, (uint64_t)p.OW * Cout * 2, (uint64_t)p.OH * p.OW * Cout * 2};
        L.map_o32 = L.map_x; L.map_o16 = L.map_x; L.map_r32 = L.map_x;
        if (out32) tma_encode(&L.map_o32, out32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dims, s32, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (res32) tma_encode(&L.map_r32, res32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dims, s32, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (out16) tma_encode(&L.map_o16, out16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dims, s16, box, CU_TENSOR_MAP_SWIZZLE_64B);
    }
    return true;
}

// Tile choice (measured on B200, tests/ir_probe.py): the expand epilogue is balanced over the four TMEM lane quadrants only when
// the halo fills its 128-row blocks, and two CTAs per SM hide each other's barriers.  Stride 1: 16 x 16 outputs (324 halo rows)
// when two CTAs still fit, else 8 x 16 (180 rows).  Stride 2: 4 x 16 with two CTAs per SM, else 8 x 16 with 16 compute warps.
bool ir_fused_prepare(IrLaunch & L, const __half * x, int N, int H, int W, int Cin, int E, int Cout, int stride, const __half * We,
                      const float * se, const float * he, const __half * dwW, const float * sd, const float * hd, const __half * Wr,
                      const float * sr, const float * hr, const float * res32, __half * out16, float * out32) {
    auto cfg = [&](int TH, int TW, int NT, int * per_sm) {
        return ir_prepare_cfg(L, x, N, H, W, Cin, E, Cout, stride, We, se, he, dwW, sd, hd, Wr, sr, hr, res32, out16, out32, TH, TW, NT, per_sm);
    };
    if (getenv("GGML_B200_IR_TH") || getenv("GGML_B200_IR_TW") || getenv("GGML_B200_IR_NT"))  // tuning probe
        return cfg(env_int("GGML_B200_IR_TH", stride == 1 ? 8 : 4), env_int("GGML_B200_IR_TW", 16), env_int("GGML_B200_IR_NT", 256), nullptr);
    static const int cand[2][3][3] = {{{16, 16, 256}, {8, 16, 256}, {8, 16, 512}}, {{4, 16, 256}, {8, 16, 512}, {4, 8, 256}}};
    int per_sm = 0;
    for (int i = 0; i < 3; i++) {
        const int * c = cand[stride == 2 ? 1 : 0][i];
        if (cfg(c[0], c[1], c[2], &per_sm) && (per_sm == 2 || c[2] == 512 || i == 2)) return true;
    }
    return false;
}

template <int STRIDE, int NTC>
static void ir_launch_variant(const IrLaunch & L, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        B200_CHECK(cudaFuncSetAttribute(k_ir_fused<STRIDE, NTC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));  // + the static barriers stays under the 227 KiB limit
        attr = true;
    }
    launch_pdl(k_ir_fused<STRIDE, NTC>, dim3(L.grid), dim3(NTC + 32), L.smem_bytes, st, L.map_x, L.map_we, L.map_wr, L.map_o32, L.map_o16, L.map_r32, L.p);
}

void ir_fused_launch(const IrLaunch & L, cudaStream_t st) {
    if (L.p.stride == 1) { if (L.p.nthreads == 256) ir_launch_variant<1, 256>(L, st); else ir_launch_variant<1, 512>(L, st); }
    else { if (L.p.nthreads == 256) ir_launch_variant<2, 256>(L, st); else ir_launch_variant<2, 512>(L, st); }
}

}  // namespace b200
