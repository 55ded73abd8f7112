// dwconv_tma.cu -- K3: depthwise 3x3 + BatchNorm + SiLU over NHWC f16 (replaces ggml_conv_depthwise_2d + BN chain +
// silu, main.cpp:788,809-850; [ggml-upstream] im2col(F16) + per-channel dot with f32 accumulation).
//
// HBM-bound (AI ~3.4 flop/B).  The register-window version was latency-bound (one 8-warp CTA per SM, ~12 KB of loads
// in flight); here the memory-level parallelism comes from the TMA instead of from threads:
//   * persistent CTAs; a tile = TH x TW output pixels x 64 channels of one image;
//   * one thread issues a 4-D TMA box {64 ch, TW*s+2, TH*s+2, 1} for the NEXT tile into the other half of a
//     double buffer while all threads compute the current one; the image border (pad 1) is the TMA's zero fill,
//     so the inner loop has no bounds checks;
//   * thread = (8-channel group, column[, row split]); 128-bit conflict-free LDS, f32 accumulate, fused
//     scale/shift + SiLU, 128-bit coalesced stores (8 lanes = one pixel's 128 bytes).
#include "gemm_tcgen05.h"
#include "internal.h"
#include "pdl.cuh"
#include "ptx_sm100.cuh"

namespace b200 {

namespace {

using namespace ptx;

// acc[0..7] += x[0..7] * w[0..7] for 8 packed halves each: 8 FHFMA, operands taken from the packed registers (.H0 / .H1)
__device__ __forceinline__ void dw_fhfma8(float (&acc)[8], const uint4 & x, const uint4 & w) {
#define DW_FHFMA2(A0, A1, X, W)                                                                                                         \
    asm("{\n\t.reg .b16 xl, xh, wl, wh;\n\tmov.b32 {xl, xh}, %2;\n\tmov.b32 {wl, wh}, %3;\n\tfma.rn.f32.f16 %0, xl, wl, %0;\n\t"     \
        "fma.rn.f32.f16 %1, xh, wh, %1;\n\t}"                                                                                       \
        : "+f"(A0), "+f"(A1)                                                                                                          \
        : "r"(X), "r"(W))
    DW_FHFMA2(acc[0], acc[1], x.x, w.x);
    DW_FHFMA2(acc[2], acc[3], x.y, w.y);
    DW_FHFMA2(acc[4], acc[5], x.z, w.z);
    DW_FHFMA2(acc[6], acc[7], x.w, w.w);
#undef DW_FHFMA2
}

struct alignas(16) H8 {
    __half2 h[4];
};

template <int STRIDE>
__global__ void __launch_bounds__(256, 2) k_dwconv_tma(const __grid_constant__ CUtensorMap map_x, const DwLaunch::Params p) {
    extern __shared__ __align__(128) unsigned char dw_smem[];
    __shared__ __align__(8) uint64_t full_bar[2];
    const uint32_t box_bytes = (uint32_t)p.box_w * p.box_h * 128u;
    const uint32_t sbase     = smem_u32(dw_smem);

    const int cg = threadIdx.x & 7;
    const int xl = (threadIdx.x >> 3) % p.TW;
    const int rs = (threadIdx.x >> 3) / p.TW;
    const int rows_per = p.TH / p.RS;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&full_bar[0]), 1);
        mbar_init(smem_u32(&full_bar[1]), 1);
        fence_barrier_init();
    }
    __syncthreads();
    pdl_wait();  // PDL: the input tile loads and the output stores below must see the previous kernel complete
    pdl_trigger();

    // tile index -> (column tile, row tile, channel block, image).  Decoded with divisions ONCE (the CTA's first tile and its stride
    // gridDim.x); afterwards the position advances by the stride with carries -- three run-time divisions per tile and thread were 8 %
    // (stride 1) to 14 % (stride 2) of this issue-bound kernel's instructions.
    struct TilePos { int tx, ty, cb, n; };
    auto decode = [&](int t) {
        TilePos v;
        v.tx = t % p.tiles_x; t /= p.tiles_x;
        v.ty = t % p.tiles_y; t /= p.tiles_y;
        v.cb = t % p.cblocks;
        v.n  = t / p.cblocks;
        return v;
    };
    const TilePos step = decode((int)gridDim.x);
    auto advance = [&](TilePos & v) {
        v.tx += step.tx;
        int c = v.tx >= p.tiles_x;
        v.tx -= c ? p.tiles_x : 0;
        v.ty += step.ty + c;
        c = v.ty >= p.tiles_y;
        v.ty -= c ? p.tiles_y : 0;
        v.cb += step.cb + c;
        c = v.cb >= p.cblocks;
        v.cb -= c ? p.cblocks : 0;
        v.n += step.n + c;
    };
    auto issue = [&](const TilePos & v, int buf) {
        const uint32_t bar = smem_u32(&full_bar[buf]);
        mbar_expect_tx(bar, box_bytes);
        tma_load_4d(sbase + (uint32_t)buf * box_bytes, &map_x, v.cb * 64, v.tx * p.TW * STRIDE - 1, v.ty * p.TH * STRIDE - 1, v.n, bar);
    };

    uint32_t it = 0;
    uint4    w[9];
    float    sc[8], sh[8];
    int      cb_loaded = -1;
    TilePos nxt = decode((int)blockIdx.x);
    if (threadIdx.x == 0 && (int)blockIdx.x < p.ntiles) issue(nxt, 0);
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, it++) {
        const int buf  = it & 1;
        const int next = tile + gridDim.x;
        const TilePos cur = nxt;
        advance(nxt);
        if (threadIdx.x == 0 && next < p.ntiles) issue(nxt, buf ^ 1);  // buffer buf^1 was released by the barrier below
        const int tx = cur.tx, ty = cur.ty, cb = cur.cb, n = cur.n;
        const int c0 = cb * 64 + cg * 8;
        const int ox = tx * p.TW + xl;
        const bool lane_ok = c0 < p.C && ox < p.OW;

        // Weights stay packed f16 (36 registers instead of 72 f32) and every tap is one FHFMA (fma.rn.f32.f16: f16 x f16
        // product, exact in f32, added to the f32 accumulator with one rounding) -- bit-identical to converting both
        // operands and using FFMA, at half the instructions.
        // (declared outside the tile loop: they only change with the channel block, i.e. every tiles_x * tiles_y tiles)
        if (lane_ok && cb != cb_loaded) {
            cb_loaded = cb;
#pragma unroll
            for (int k = 0; k < 9; k++) w[k] = *reinterpret_cast<const uint4 *>(p.Wt + (size_t)k * p.C + c0);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float ah = p.act ? 0.5f : 1.f;  // SiLU works on y/2: folded into scale / shift, exactly (ptx_sm100.cuh silu_h)
                sc[j] = (p.scale ? p.scale[c0 + j] : 1.f) * ah;
                sh[j] = (p.shift ? p.shift[c0 + j] : 0.f) * ah;
            }
        }
        mbar_wait(smem_u32(&full_bar[buf]), (it >> 1) & 1u);
        if (lane_ok) {
            // running 32-bit shared addresses and one running output pointer: the row loops carry no index arithmetic
            const uint32_t pitch = (uint32_t)p.box_w * 128u;
            const int      oyl0  = rs * rows_per;
            uint32_t       rp    = sbase + (uint32_t)buf * box_bytes + (uint32_t)cg * 16u + (uint32_t)(xl * STRIDE) * 128u + (uint32_t)(oyl0 * STRIDE) * pitch;
            __half *       op    = p.out + (((size_t)n * p.OH + (ty * p.TH + oyl0)) * p.OW + ox) * p.C + c0;
            const size_t   opitch = (size_t)p.OW * p.C;
            auto load_row = [&](uint4 (&dst)[3]) {  // the 3 taps of the row at rp, then advance one input row
#pragma unroll
                for (int kw = 0; kw < 3; kw++)
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(dst[kw].x), "=r"(dst[kw].y), "=r"(dst[kw].z), "=r"(dst[kw].w) : "r"(rp + (uint32_t)kw * 128u));
                rp += pitch;
            };
            auto emit = [&](const uint4 (&r0)[3], const uint4 (&r1)[3], const uint4 (&r2)[3]) {
                float acc[8];
#pragma unroll
                for (int j = 0; j < 8; j++) acc[j] = 0.f;
#pragma unroll
                for (int kw = 0; kw < 3; kw++) dw_fhfma8(acc, r0[kw], w[kw]);
#pragma unroll
                for (int kw = 0; kw < 3; kw++) dw_fhfma8(acc, r1[kw], w[3 + kw]);
#pragma unroll
                for (int kw = 0; kw < 3; kw++) dw_fhfma8(acc, r2[kw], w[6 + kw]);
                H8 o;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    float y0 = fmaf(acc[2 * j], sc[2 * j], sh[2 * j]);
                    float y1 = fmaf(acc[2 * j + 1], sc[2 * j + 1], sh[2 * j + 1]);
                    if (p.act) { y0 = silu_h(y0); y1 = silu_h(y1); }
                    o.h[j] = __floats2half2_rn(y0, y1);
                }
                *reinterpret_cast<H8 *>(op) = o;
                op += opitch;
            };
            int rows = p.OH - (ty * p.TH + oyl0);  // output rows of this thread that exist
            if (rows > rows_per) rows = rows_per;
            if (STRIDE == 1) {
                // sliding 3-row window in registers: each output row loads only its new bottom row (3 LDS.128 instead of 9)
                uint4 ra[3], rb[3], rc[3];
                if (rows > 0) {
                    load_row(ra);
                    load_row(rb);
                }
                for (int r = 0; r < rows; r += 3) {
                    load_row(rc);
                    emit(ra, rb, rc);
                    if (r + 1 < rows) {
                        load_row(ra);
                        emit(rb, rc, ra);
                    }
                    if (r + 2 < rows) {
                        load_row(rb);
                        emit(rc, ra, rb);
                    }
                }
            } else {
                // stride 2: consecutive outputs share one input row (2r+2 is the next output's top row)
                uint4 ra[3], rb[3], rc[3];
                if (rows > 0) load_row(ra);
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
        __syncthreads();  // every thread is done with buffer `buf`: it may be refilled by the issue of the next iteration
    }
}

}  // namespace

bool dw_prepare(DwLaunch & L, const __half * x, int N, int H, int W, int C, int stride, const __half * Wt, const float * scale,
                const float * shift, int act, __half * out) {
    if (C % 8 || H % stride || W % stride || (stride != 1 && stride != 2)) return false;
    L = DwLaunch();
    DwLaunch::Params & p = L.p;
    p.N = N; p.H = H; p.W = W; p.C = C; p.stride = stride;
    p.OH = H / stride; p.OW = W / stride;
    p.TW = p.OW >= 32 && stride == 1 ? 32 : (p.OW >= 16 ? 16 : (p.OW >= 8 ? 8 : (p.OW >= 4 ? 4 : (p.OW >= 2 ? 2 : 1))));
    p.RS = 32 / p.TW;  // 256 threads = 8 channel groups x TW columns x RS row splits
    if (p.RS < 1) p.RS = 1;
    int th = stride == 1 ? 8 : 4;
    if (th < p.RS) th = p.RS;
    while (th > p.OH && th > p.RS) th /= 2;
    if (th % p.RS) th = p.RS;
    p.TH = th;
    p.tiles_x = (p.OW + p.TW - 1) / p.TW;
    p.tiles_y = (p.OH + p.TH - 1) / p.TH;
    p.cblocks = (C + 63) / 64;
    p.box_w   = p.TW * stride + 2;
    p.box_h   = p.TH * stride + 2;
    p.ntiles  = N * p.cblocks * p.tiles_y * p.tiles_x;
    p.act = act; p.Wt = Wt; p.scale = scale; p.shift = shift; p.out = out;
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t str[3]  = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    const uint32_t box[4]  = {64, (uint32_t)p.box_w, (uint32_t)p.box_h, 1};
    if (p.box_w > 256 || p.box_h > 256) return false;
    tma_encode(&L.map_x, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    L.smem_bytes = (size_t)2 * p.box_w * p.box_h * 128;
    const int per_sm = L.smem_bytes <= 100 * 1024 ? 2 : 1;
    const int cap    = per_sm * runtime().sm_count;
    L.grid           = p.ntiles < cap ? p.ntiles : cap;
    return true;
}

void dw_launch(const DwLaunch & L, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        B200_CHECK(cudaFuncSetAttribute(k_dwconv_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        B200_CHECK(cudaFuncSetAttribute(k_dwconv_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    const int threads = 8 * L.p.TW * L.p.RS;
    if (L.p.stride == 1) launch_pdl(k_dwconv_tma<1>, dim3(L.grid), dim3(threads), L.smem_bytes, st, L.map_x, L.p);
    else launch_pdl(k_dwconv_tma<2>, dim3(L.grid), dim3(threads), L.smem_bytes, st, L.map_x, L.p);
}

}  // namespace b200
