// ir_fused.cu -- K4: the whole inverted-residual block in ONE kernel (inverted_residual_layer::forward, main.cpp:854-870):
//
//     expand 1x1 (+BN+SiLU)  ->  depthwise 3x3, stride 1|2 (+BN+SiLU)  ->  reduce 1x1 (+BN)  [-> + inp]
//
// Unfused, the 4x-expanded activation is written to HBM by the expand GEMM, read and re-written by the depthwise kernel and
// read again by the reduce GEMM: ~60 % of all activation traffic of MobileViT-S and 2.7 of 5.5 ms per
// batch-256 step (profiles/layers_r1.txt).  Here only the block input and the block output touch HBM:
//
//   tile = TH x TW output pixels of one image; halo = IH x IW input pixels (IH = (TH-1)*s+3), P_in = IH*IW rows
//   TMA        : 4-D box {Cin block, IW, IH, 1} of the block input x (f16 NHWC), halo zero-filled      -> xs   (UMMA A operand)
//   per 64-channel chunk c of the expanded width E  (E/64 chunks):
//     TMA        : expand weights We[c*64.., Cin] and reduce weights Wr[Cout, c*64..]                   -> we[b], wr[b]
//     tcgen05.mma: exp[mb] (128 halo pixels x 64 ch, f32 in TMEM) = xs[mb] . we^T          for every 128-row block mb
//     all warps  : tcgen05.ld -> BN + SiLU -> 0 outside the image (the depthwise pads the EXPANDED map, main.cpp:784)
//                  -> f16 (rounding point #1 = the depthwise conv's im2col)                            -> es  (pitch 144 B)
//     all warps  : depthwise 3x3 from es (FHFMA, sliding 3-row register window) + BN + SiLU -> f16 (rounding point #2 =
//                  the reduce conv's im2col), written in the 128B-swizzled K-major UMMA layout           -> as
//     tcgen05.mma: red[mo] (128 output pixels x Cout, f32 in TMEM) += as[mo] . wr^T
//   after the last chunk: tcgen05.ld -> BN (+ f32 residual) -> f16 / f32 global stores
//
// The expand MMA of chunk c+1 runs while the warps compute the depthwise of chunk c, the reduce MMA of chunk c while they
// run the expand epilogue of chunk c+1; weight chunks are double-buffered and requested one chunk ahead; the next tile's
// input halo is requested as soon as the last expand MMA of the current tile has completed.  One thread (thread 0) issues
// every TMA and MMA; phases are separated by __syncthreads, asynchronous completions by mbarriers.
//
// Rounding points are those of the unfused plan and of ggml (f16 operands at both convolution inputs, f32 accumulation,
// f32 BatchNorm applied to the accumulator), so results agree with the three separate kernels to f32 summation order.
#include "gemm_tcgen05.h"
#include "internal.h"
#include "pdl.cuh"
#include "ptx_sm100.cuh"

namespace b200 {

using namespace ptx;

namespace {

constexpr int kEPitch = 144;  // bytes per halo pixel in the expanded-chunk buffer: 128 B of data + 16 B so that both the row-per-thread
                              // epilogue stores and the 8-lanes-per-pixel depthwise loads are bank-conflict free without a swizzle

__device__ __forceinline__ uint32_t ir_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float * v) {
    uint32_t * r = reinterpret_cast<uint32_t *>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void fhfma8(float (&acc)[8], const uint4 & x, const uint4 & w) {
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
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// NT threads (256: two CTAs per SM, 512: one); STRIDE of the depthwise convolution.
template <int STRIDE, int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 2 : 1)
k_ir_fused(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_we, const __grid_constant__ CUtensorMap map_wr,
           const IrLaunch::Params p) {
    constexpr int WQ = NT / 128;   // warps per TMEM lane quadrant
    constexpr int CW = 64 / WQ;    // expanded columns per warp in the expand epilogue (32 or 16)
    extern __shared__ uint8_t ir_smem_raw[];
    uint8_t * smem = ir_smem_raw + ((1024u - (smem_u32(ir_smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bars[8];  // x_full, we_full[2], wr_full[2], exp_full, red_done
    __shared__ uint32_t tmem_slot;
    const uint32_t sbase  = smem_u32(smem);
    const uint32_t xs     = sbase + p.off_x;
    const uint32_t es     = sbase + p.off_e;
    const uint32_t as     = sbase + p.off_a;
    const uint32_t wes    = sbase + p.off_we;
    const uint32_t wrs    = sbase + p.off_wr;
    // parameter block: se, he, sd, hd [Epad] f32, sr, hr [Cout_pad] f32, depthwise weights [9][Epad] f16
    const int      Epad   = p.NC * 64;
    float *        s_se   = reinterpret_cast<float *>(smem + p.off_par);
    float *        s_he   = s_se + Epad;
    float *        s_sd   = s_he + Epad;
    float *        s_hd   = s_sd + Epad;
    float *        s_sr   = s_hd + Epad;
    float *        s_hr   = s_sr + p.Cout_pad;
    __half *       s_dww  = reinterpret_cast<__half *>(s_hr + p.Cout_pad);
    const uint32_t x_full = smem_u32(&bars[0]), we_full = smem_u32(&bars[1]), wr_full = smem_u32(&bars[3]);
    const uint32_t exp_full = smem_u32(&bars[5]), red_done = smem_u32(&bars[6]);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, wsub = warp >> 2;

    if (tid == 0) {
        for (int i = 0; i < 7; i++) mbar_init(smem_u32(&bars[i]), 1);
        fence_barrier_init();
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_we);
        tma_prefetch_desc(&map_wr);
    }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)p.tmem_cols);
    // constants -> shared (padded channels get scale 1 / shift 0 / weight 0, so they produce exact zeros end to end)
    for (int i = tid; i < Epad; i += NT) {
        const bool ok = i < p.E;
        s_se[i] = ok && p.se ? p.se[i] : 1.f;
        s_he[i] = ok && p.he ? p.he[i] : 0.f;
        s_sd[i] = ok && p.sd ? p.sd[i] : 1.f;
        s_hd[i] = ok && p.hd ? p.hd[i] : 0.f;
    }
    for (int i = tid; i < p.Cout_pad; i += NT) {
        const bool ok = i < p.Cout;
        s_sr[i] = ok && p.sr ? p.sr[i] : 1.f;
        s_hr[i] = ok && p.hr ? p.hr[i] : 0.f;
    }
    for (int i = tid; i < 9 * Epad; i += NT) {
        const int k = i / Epad, c = i - k * Epad;
        s_dww[i] = c < p.E ? p.dwW[(size_t)k * p.E + c] : __float2half(0.f);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t tmem_red  = tmem_base + (uint32_t)p.MBI * 64u;

    const bool has_work = (int)blockIdx.x < p.ntiles;
    // ---- thread-0 helpers: every TMA and MMA of the CTA is issued here -------------------------------------------------
    auto issue_we = [&](uint32_t g, int c) {  // expand weights of chunk c -> buffer g & 1
        const uint32_t bar = we_full + 8u * (g & 1u);
        mbar_expect_tx(bar, p.we_bytes);
        for (int kb = 0; kb < p.num_kb; kb++) tma_load_2d(wes + (g & 1u) * p.we_bytes + (uint32_t)kb * p.we_kb_stride, &map_we, kb * 64, c * 64, bar);
    };
    auto issue_wr = [&](uint32_t g, int c) {  // reduce weights of chunk c -> buffer g & 1
        const uint32_t bar = wr_full + 8u * (g & 1u);
        mbar_expect_tx(bar, p.wr_bytes);
        tma_load_2d(wrs + (g & 1u) * p.wr_bytes, &map_wr, c * 64, 0, bar);
    };
    auto issue_x = [&](int tile) {
        int t = tile;
        const int tx = t % p.tiles_x; t /= p.tiles_x;
        const int ty = t % p.tiles_y;
        const int n  = t / p.tiles_y;
        mbar_expect_tx(x_full, p.x_tx_bytes);
        for (int kb = 0; kb < p.num_kb; kb++)
            tma_load_4d(xs + (uint32_t)kb * p.x_kb_stride, &map_x, kb * 64, tx * p.TW * STRIDE - 1, ty * p.TH * STRIDE - 1, n, x_full);
    };
    const uint32_t idesc_e = ir_idesc(64), idesc_r = ir_idesc(p.Cout_pad);
    auto issue_expand = [&](uint32_t g) {
        mbar_wait(we_full + 8u * (g & 1u), (g >> 1) & 1u);
        tc_fence_after();
        const uint32_t wb = wes + (g & 1u) * p.we_bytes;
        for (int mb = 0; mb < p.MBI; mb++) {
            for (int kb = 0; kb < p.num_kb; kb++) {
                const uint64_t adesc = make_smem_desc(xs + (uint32_t)kb * p.x_kb_stride + (uint32_t)mb * 128u * (uint32_t)p.row_bytes, (uint32_t)p.row_bytes);
                const uint64_t bdesc = make_smem_desc(wb + (uint32_t)kb * p.we_kb_stride, (uint32_t)p.row_bytes);
                for (int k = 0; k < p.ksteps; k++) umma_f16(tmem_base + (uint32_t)mb * 64u, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_e, (kb | k) != 0);
            }
        }
        umma_commit(exp_full);
    };
    auto issue_reduce = [&](uint32_t g, int c) {
        mbar_wait(wr_full + 8u * (g & 1u), (g >> 1) & 1u);
        tc_fence_after();
        const uint64_t bdesc = make_smem_desc(wrs + (g & 1u) * p.wr_bytes, 128);
        for (int mo = 0; mo < p.MBO; mo++) {
            const uint64_t adesc = make_smem_desc(as + (uint32_t)mo * 16384u, 128);
#pragma unroll
            for (int k = 0; k < 4; k++) umma_f16(tmem_red + (uint32_t)mo * (uint32_t)p.red_stride, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_r, (c | k) != 0);
        }
        umma_commit(red_done);
    };

    // weights are constants: requested before the PDL wait, so they overlap the tail of the previous kernel
    if (tid == 0 && has_work) {
        issue_we(0, 0);
        issue_wr(0, 0);
    }
    pdl_wait();
    pdl_trigger();
    if (tid == 0 && has_work) issue_x(blockIdx.x);

    // ---- per-thread roles --------------------------------------------------------------------------------------------
    // depthwise: thread = (8-channel group cg, output column xl, row split rs)
    const int cg   = tid & 7;
    const int slot = tid >> 3;
    const int xl   = slot % p.TW;
    const int rs   = slot / p.TW;
    const int RS   = (NT / 8) / p.TW;
    const int rows_per = p.TH / RS;
    const uint32_t epitch = (uint32_t)p.IW * kEPitch;

    uint32_t g = 0, tcount = 0;  // chunk counter / tile counter of this CTA
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, tcount++) {
        int t = tile;
        const int tx = t % p.tiles_x; t /= p.tiles_x;
        const int ty = t % p.tiles_y;
        const int n  = t / p.tiles_y;
        const int y0 = ty * p.TH * STRIDE - 1, x0 = tx * p.TW * STRIDE - 1;  // image coordinates of halo pixel (0,0)
        const bool next_tile = tile + (int)gridDim.x < p.ntiles;

        if (tid == 0) {
            mbar_wait(x_full, tcount & 1u);
            issue_expand(g);
            if (p.NC > 1 || next_tile) issue_we(g + 1, p.NC > 1 ? 1 : 0);
        }
        for (int c = 0; c < p.NC; c++, g++) {
            const bool last_chunk = c + 1 == p.NC;
            // ---- expand accumulators of chunk c are complete ----
            mbar_wait(exp_full, g & 1u);
            tc_fence_after();
            if (tid == 0 && last_chunk && next_tile) issue_x(tile + gridDim.x);  // every expand MMA of this tile has read xs
            __syncwarp();  // warp 0 reconverges before the warp-aligned tcgen05.ld below

            // ---- expand epilogue: TMEM -> BN + SiLU -> mask -> f16 -> es ----
            for (int mb = 0; mb < p.MBI; mb++) {
                const int row0 = mb * 128 + quad * 32;
                if (row0 >= p.P_in) break;  // warp-uniform
                const int row = row0 + lane;
                float v[CW];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mb * 64 + wsub * CW);
                if (CW == 32) tmem_ld_32x32(taddr, v); else tmem_ld_32x16(taddr, v);
                const int yy = row / p.IW, xx = row - yy * p.IW;
                const int gy = y0 + yy, gx = x0 + xx;
                const bool inimg = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
                const uint32_t dst = es + (uint32_t)row * kEPitch + (uint32_t)(wsub * CW) * 2u;
#pragma unroll
                for (int g8 = 0; g8 < CW / 8; g8++) {
                    const int col = c * 64 + wsub * CW + g8 * 8;
                    const uint32_t sa = smem_u32(s_se + col), sb = smem_u32(s_he + col);
                    const float4 s0 = ld_shared_f4(sa), s1 = ld_shared_f4(sa + 16), h0 = ld_shared_f4(sb), h1 = ld_shared_f4(sb + 16);
                    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                    const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                    float y[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float tt = fmaf(v[g8 * 8 + j], sc[j], sh[j]);
                        y[j]           = inimg ? silu_f(tt) : 0.f;
                    }
                    if (row < p.P_in) st_shared_v4(dst + (uint32_t)g8 * 16u, pack_h2(y[0], y[1]), pack_h2(y[2], y[3]), pack_h2(y[4], y[5]), pack_h2(y[6], y[7]));
                }
            }
            tc_fence_before();
            __syncthreads();  // es complete; the expand accumulators are free again
            if (tid == 0 && !last_chunk) {
                // expand MMA of the next chunk runs under the depthwise phase below; its successor's weights are requested now
                issue_expand(g + 1);
                if (c + 2 < p.NC || next_tile) issue_we(g + 2, c + 2 < p.NC ? c + 2 : 0);
            }

            // ---- depthwise 3x3 + BN + SiLU: es -> as (UMMA A operand, 128B swizzle) ----
            if (g > 0) mbar_wait(red_done, (g - 1) & 1u);  // the reduce MMA of the previous chunk has read `as` (and its weight buffer)
            if (tid == 0 && (!last_chunk || next_tile)) issue_wr(g + 1, last_chunk ? 0 : c + 1);
            if (rs < RS) {
                const int ch = c * 64 + cg * 8;
                uint4 w[9];
#pragma unroll
                for (int k = 0; k < 9; k++) w[k] = *reinterpret_cast<const uint4 *>(s_dww + (size_t)k * Epad + ch);
                float sc[8], sh[8];
                {
                    const float4 a0 = *reinterpret_cast<const float4 *>(s_sd + ch), a1 = *reinterpret_cast<const float4 *>(s_sd + ch + 4);
                    const float4 b0 = *reinterpret_cast<const float4 *>(s_hd + ch), b1 = *reinterpret_cast<const float4 *>(s_hd + ch + 4);
                    sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
                    sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
                }
                const int oyl0 = rs * rows_per;
                uint32_t  rp   = es + (uint32_t)cg * 16u + (uint32_t)(xl * STRIDE) * kEPitch + (uint32_t)(oyl0 * STRIDE) * epitch;
                int       q    = oyl0 * p.TW + xl;  // row of the A operand = output pixel index inside the tile
                auto load_row = [&](uint4 (&dst)[3]) {
#pragma unroll
                    for (int kw = 0; kw < 3; kw++) dst[kw] = lds128(rp + (uint32_t)kw * kEPitch);
                    rp += epitch;
                };
                auto emit = [&](const uint4 (&r0)[3], const uint4 (&r1)[3], const uint4 (&r2)[3]) {
                    float acc[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[j] = 0.f;
#pragma unroll
                    for (int kw = 0; kw < 3; kw++) fhfma8(acc, r0[kw], w[kw]);
#pragma unroll
                    for (int kw = 0; kw < 3; kw++) fhfma8(acc, r1[kw], w[3 + kw]);
#pragma unroll
                    for (int kw = 0; kw < 3; kw++) fhfma8(acc, r2[kw], w[6 + kw]);
                    uint32_t o[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) o[j] = pack_h2(silu_f(fmaf(acc[2 * j], sc[2 * j], sh[2 * j])), silu_f(fmaf(acc[2 * j + 1], sc[2 * j + 1], sh[2 * j + 1])));
                    st_shared_v4(as + (uint32_t)q * 128u + (((uint32_t)cg ^ ((uint32_t)q & 7u)) << 4), o[0], o[1], o[2], o[3]);
                    q += p.TW;
                };
                const int rows = rows_per;
                uint4 ra[3], rb[3], rc[3];
                if (STRIDE == 1) {
                    load_row(ra);
                    load_row(rb);
                    for (int r = 0; r < rows; r += 3) {
                        load_row(rc);
                        emit(ra, rb, rc);
                        if (r + 1 < rows) { load_row(ra); emit(rb, rc, ra); }
                        if (r + 2 < rows) { load_row(rb); emit(rc, ra, rb); }
                    }
                } else {
                    load_row(ra);
                    for (int r = 0; r < rows; r += 2) {
                        load_row(rb);
                        load_row(rc);
                        emit(ra, rb, rc);
                        if (r + 1 < rows) {
                            load_row(ra);
                            load_row(rb);
                            emit(rc, ra, rb);
#pragma unroll
                            for (int kw = 0; kw < 3; kw++) ra[kw] = rb[kw];
                        }
                    }
                }
            }
            fence_proxy_async();  // generic-proxy writes of `as` -> visible to the tensor core (async proxy)
            __syncthreads();
            if (tid == 0) issue_reduce(g, c);
        }

        // ---- reduce epilogue: TMEM -> BN (+ residual) -> global ----
        mbar_wait(red_done, (g - 1) & 1u);
        tc_fence_after();
        __syncwarp();
        for (int mo = 0; mo < p.MBO; mo++) {
            const int row = quad * 32 + lane;
            const int qi  = mo * 128 + row;
            const int oyl = qi / p.TW, xr = qi - oyl * p.TW;
            const int oy  = ty * p.TH + oyl, ox = tx * p.TW + xr;
            const bool ok = qi < p.P_out && oy < p.OH && ox < p.OW;
            const size_t pix = ((size_t)n * p.OH + oy) * p.OW + ox;
            for (int cc = wsub; cc * 32 < p.Cout_pad; cc += WQ) {
                float v[32];
                __syncwarp();  // tcgen05.ld is warp-aligned: lanes whose pixel lies outside the image skip the stores, not the load
                tmem_ld_32x32(tmem_red + (uint32_t)mo * (uint32_t)p.red_stride + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cc * 32), v);
#pragma unroll
                for (int g8 = 0; g8 < 4; g8++) {
                    const int nn = cc * 32 + g8 * 8;
                    if (!ok || nn + 8 > p.Cout) continue;
                    float y[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) y[j] = fmaf(v[g8 * 8 + j], s_sr[nn + j], s_hr[nn + j]);
                    if (p.res32) {
                        const float4 r0 = *reinterpret_cast<const float4 *>(p.res32 + pix * p.Cout + nn);
                        const float4 r1 = *reinterpret_cast<const float4 *>(p.res32 + pix * p.Cout + nn + 4);
                        y[0] += r0.x; y[1] += r0.y; y[2] += r0.z; y[3] += r0.w;
                        y[4] += r1.x; y[5] += r1.y; y[6] += r1.z; y[7] += r1.w;
                    }
                    if (p.out32) {
                        float4 * o = reinterpret_cast<float4 *>(p.out32 + pix * p.Cout + nn);
                        o[0] = make_float4(y[0], y[1], y[2], y[3]);
                        o[1] = make_float4(y[4], y[5], y[6], y[7]);
                    }
                    if (p.out16) {
                        uint4 o;
                        o.x = pack_h2(y[0], y[1]); o.y = pack_h2(y[2], y[3]); o.z = pack_h2(y[4], y[5]); o.w = pack_h2(y[6], y[7]);
                        *reinterpret_cast<uint4 *>(p.out16 + pix * p.Cout + nn) = o;
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();  // the reduce accumulators are free for the next tile
    }
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

bool ir_fused_prepare(IrLaunch & L, const __half * x, int N, int H, int W, int Cin, int E, int Cout, int stride, const __half * We,
                      const float * se, const float * he, const __half * dwW, const float * sd, const float * hd, const __half * Wr,
                      const float * sr, const float * hr, const float * res32, __half * out16, float * out32) {
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
    // tile: 8 x 16 outputs for stride 1 (halo 10 x 18 = 180 rows, two 128-row MMA blocks); 4 x 8 for stride 2 (halo 9 x 17 = 153)
    int TW = stride == 1 ? 16 : 8, TH = stride == 1 ? 8 : 4, NT = 256;
    TW = env_int("GGML_B200_IR_TW", TW); TH = env_int("GGML_B200_IR_TH", TH); NT = env_int("GGML_B200_IR_NT", NT);
    while (TW > 1 && TW / 2 >= p.OW) TW /= 2;
    if (NT != 256 && NT != 512) return false;
    if ((NT / 8) % TW) return false;
    const int RS = (NT / 8) / TW;
    if (TH % RS) TH = (TH + RS - 1) / RS * RS;
    p.TH = TH; p.TW = TW; p.nthreads = NT;
    p.IH = (TH - 1) * stride + 3; p.IW = (TW - 1) * stride + 3;
    p.P_in = p.IH * p.IW; p.P_out = TH * TW;
    p.MBI = (p.P_in + 127) / 128; p.MBO = (p.P_out + 127) / 128;
    p.tiles_x = (p.OW + TW - 1) / TW; p.tiles_y = (p.OH + TH - 1) / TH;
    p.ntiles  = N * p.tiles_x * p.tiles_y;
    p.red_stride = (p.Cout_pad + 31) / 32 * 32;
    const int need = p.MBI * 64 + p.MBO * p.red_stride;
    if (need > 512 || p.IW > 256 || p.IH > 256) return false;
    p.tmem_cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
    // shared memory layout (offsets from a 1 KiB aligned base)
    auto up = [](uint32_t v, uint32_t a) { return (v + a - 1) / a * a; };
    const uint32_t rows32 = up((uint32_t)p.P_in, 32);
    p.x_kb_stride = up(rows32 * (uint32_t)p.row_bytes, 1024);
    p.x_tx_bytes  = (uint32_t)p.num_kb * (uint32_t)p.P_in * (uint32_t)p.row_bytes;
    uint32_t off  = 0;
    p.off_x = off; off += (uint32_t)p.num_kb * p.x_kb_stride;
    // the last 128-row MMA block reads past P_in: those rows must only be addressable (their results are never used), so the
    // buffers that follow double as that slack
    const uint32_t x_span = (uint32_t)(p.num_kb - 1) * p.x_kb_stride + (uint32_t)p.MBI * 128u * (uint32_t)p.row_bytes;
    p.off_a = up(off, 1024); off = p.off_a + (uint32_t)p.MBO * 16384u;
    p.we_kb_stride = up(64u * (uint32_t)p.row_bytes, 1024);
    p.we_bytes     = (uint32_t)p.num_kb * p.we_kb_stride;
    p.off_we = up(off, 1024); off = p.off_we + 2 * p.we_bytes;
    p.wr_bytes = (uint32_t)p.Cout_pad * 128u;  // Cout_pad is a multiple of 16: the buffers stay 1 KiB aligned (2 KiB granules)
    p.off_wr = up(off, 1024); off = p.off_wr + 2 * p.wr_bytes;
    p.off_e = up(off, 16); off = p.off_e + (uint32_t)p.P_in * kEPitch;
    p.off_par = up(off, 16);
    const uint32_t Epad = (uint32_t)p.NC * 64u;
    off = p.off_par + (4 * Epad + 2 * (uint32_t)p.Cout_pad) * 4 + 9 * Epad * 2;
    if (off < p.off_x + x_span) off = p.off_x + x_span;
    L.smem_bytes = 1024 + off;
    if (L.smem_bytes > 226 * 1024) return false;
    int per_sm = (NT == 256 && 2 * L.smem_bytes <= 226 * 1024 && p.tmem_cols <= 256) ? 2 : 1;
    if (per_sm == 2 && L.smem_bytes < 78 * 1024) L.smem_bytes = 78 * 1024;  // never three CTAs on an SM: 3 x 256 TMEM columns do not exist
    if (per_sm == 1 && L.smem_bytes < 116 * 1024) L.smem_bytes = 116 * 1024;
    const int cap = per_sm * runtime().sm_count;
    L.grid = p.ntiles < cap ? p.ntiles : cap;
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
    return true;
}

template <int STRIDE, int NT>
static void ir_launch_variant(const IrLaunch & L, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        B200_CHECK(cudaFuncSetAttribute(k_ir_fused<STRIDE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr = true;
    }
    launch_pdl(k_ir_fused<STRIDE, NT>, dim3(L.grid), dim3(NT), L.smem_bytes, st, L.map_x, L.map_we, L.map_wr, L.p);
}

void ir_fused_launch(const IrLaunch & L, cudaStream_t st) {
    if (L.p.stride == 1) { if (L.p.nthreads == 256) ir_launch_variant<1, 256>(L, st); else ir_launch_variant<1, 512>(L, st); }
    else { if (L.p.nthreads == 256) ir_launch_variant<2, 256>(L, st); else ir_launch_variant<2, 512>(L, st); }
}

}  // namespace b200
