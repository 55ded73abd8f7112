// dwreduce.cu -- K4a: depthwise 3x3 + BN + SiLU fused with the following 1x1 "reduce" convolution + BN (+ residual).
//
// Second half of inverted_residual_layer::forward (main.cpp:863-868): conv_3x3 (depthwise, BN, SiLU) -> reduce_1x1 (BN,
// no activation) -> optional `+ inp`.  Unfused, the depthwise output (1.04 G elements per batch-256 step of MobileViT-S)
// is written to HBM and read back by the reduce GEMM; B200's write-only bandwidth (3.9 TB/s measured) makes that the
// most expensive kind of traffic.  Here the depthwise result never leaves the SM:
//
//   per CTA (persistent) : tile = TH x TW output pixels (128, or 64 for stride 2), loop over 64-channel blocks cb
//     TMA   : 4-D box {64 ch, TW*s+2, TH*s+2, 1} of the expanded activation (zero-filled halo)  -> x[buf]
//             2-D box {64 k, Cout} of the reduce weights [Cout][E] (128B swizzle)                -> wr[buf]
//     CUDA cores : depthwise 3x3 (f16 in, f32 accumulate) + BN + SiLU -> f16 -> ds[buf] written in the 128B-swizzled
//                  K-major layout of a UMMA A operand (row = pixel, K = 64 channels)
//     tcgen05.mma: acc[128 px, Cout] (+)= ds[buf] x wr[buf]^T   (f32 in TMEM), tcgen05.commit frees ds/wr[buf]
//   after the last cb : tcgen05.ld -> reduce BN (+ f32 residual) -> f16 / f32 global stores
//
// Rounding points are unchanged: the depthwise output is rounded to f16 exactly where the unfused pipeline (and ggml's
// im2col of the reduce conv) rounds it.
#include "gemm_tcgen05.h"
#include "internal.h"
#include "ptx_sm100.cuh"

namespace b200 {

using namespace ptx;

namespace {

__device__ __forceinline__ uint32_t dwr_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

struct alignas(16) H8 {
    __half2 h[4];
};

// 288 threads: warps 0-7 compute (depthwise + epilogue), warp 8 = control (TMA issue + single-thread MMA issue).
// There is no block-wide barrier in the steady state: compute warps and the control thread meet only through mbarriers
//   x_full/w_full[2]  TMA landed            ds_full[2]  8 compute warps wrote ds[buf]      ds_free[2]  MMA read ds/wr[buf]
//   acc_full          accumulator complete  acc_free    8 compute warps drained the accumulator
template <int STRIDE>
__global__ void __launch_bounds__(288, 2) k_dwreduce(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                                                     const DwRedLaunch::Params p) {
    extern __shared__ uint8_t dwr_smem_raw[];
    uint8_t * smem = dwr_smem_raw + ((1024u - (smem_u32(dwr_smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bars[10];  // x_full[2], w_full[2], ds_full[2], ds_free[2], acc_full, acc_free
    __shared__ uint32_t tmem_slot;
    const uint32_t ds_base  = smem_u32(smem);                       // 2 x 16 KiB
    const uint32_t wr_bytes = (uint32_t)p.Cout_pad * 128u;
    const uint32_t wr_base  = ds_base + 2u * 16384u;                // 2 x Cout_pad x 128 B
    const uint32_t box_bytes = (uint32_t)p.box_w * p.box_h * 128u;
    const uint32_t x_base   = wr_base + 2u * wr_bytes;              // 2 x box
    uint8_t *      x_ptr    = smem + 2 * 16384 + 2 * wr_bytes;
    const uint32_t x_full = smem_u32(&bars[0]), w_full = smem_u32(&bars[2]), ds_full = smem_u32(&bars[4]), ds_free = smem_u32(&bars[6]);
    const uint32_t acc_full = smem_u32(&bars[8]), acc_free = smem_u32(&bars[9]);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < 4; i++) mbar_init(smem_u32(&bars[i]), 1);
        mbar_init(ds_full, 8); mbar_init(ds_full + 8, 8);
        mbar_init(ds_free, 1); mbar_init(ds_free + 8, 1);
        mbar_init(acc_full, 1);
        mbar_init(acc_free, 8);
        fence_barrier_init();
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w);
    }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_slot;

    if (warp == 8) {
        // ===================== control thread: TMA producer + MMA issuer =====================
        if (lane == 0) {
            auto issue = [&](int tile, int cb, uint32_t buf) {
                int t = tile;
                const int tx = t % p.tiles_x; t /= p.tiles_x;
                const int ty = t % p.tiles_y;
                const int n  = t / p.tiles_y;
                mbar_expect_tx(x_full + 8u * buf, box_bytes);
                tma_load_4d(x_base + buf * box_bytes, &map_x, cb * 64, tx * p.TW * STRIDE - 1, ty * p.TH * STRIDE - 1, n, x_full + 8u * buf);
                mbar_expect_tx(w_full + 8u * buf, wr_bytes);
                tma_load_2d(wr_base + buf * wr_bytes, &map_w, cb * 64, 0, w_full + 8u * buf);
            };
            const uint32_t idesc = dwr_idesc(p.Cout_pad);
            uint32_t it = 0, tcount = 0;
            if ((int)blockIdx.x < p.ntiles) issue(blockIdx.x, 0, 0);
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, tcount++) {
                for (int cb = 0; cb < p.cblocks; cb++, it++) {
                    const uint32_t buf = it & 1u;
                    {   // prefetch item it+1 into the other buffers: free once the MMA of item it-1 (same buffers) has completed
                        int ncb = cb + 1, ntile = tile;
                        if (ncb == p.cblocks) { ncb = 0; ntile = tile + gridDim.x; }
                        if (ntile < p.ntiles) {
                            const uint32_t j = it + 1;
                            mbar_wait(ds_free + 8u * (j & 1u), ((j >> 1) & 1u) ^ 1u);
                            issue(ntile, ncb, j & 1u);
                        }
                    }
                    mbar_wait(ds_full + 8u * buf, (it >> 1) & 1u);
                    mbar_wait(w_full + 8u * buf, (it >> 1) & 1u);
                    if (cb == 0) mbar_wait(acc_free, (tcount & 1u) ^ 1u);  // previous tile's epilogue has drained the accumulator
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc(ds_base + buf * 16384u, 128);
                    const uint64_t bdesc = make_smem_desc(wr_base + buf * wr_bytes, 128);
#pragma unroll
                    for (int k = 0; k < 4; k++) umma_f16(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (cb | k) != 0);
                    umma_commit(ds_free + 8u * buf);
                    if (cb == p.cblocks - 1) umma_commit(acc_full);
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== compute warps =====================
        const int cg = tid & 7;
        const int xl = (tid >> 3) % p.TW;
        const int rs = (tid >> 3) / p.TW;
        const int RS = 32 / p.TW;
        const int rows_per = p.TH / RS;
        uint32_t it = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, tcount++) {
            int t = tile;
            const int tx = t % p.tiles_x; t /= p.tiles_x;
            const int ty = t % p.tiles_y;
            const int n  = t / p.tiles_y;
            for (int cb = 0; cb < p.cblocks; cb++, it++) {
                const uint32_t buf = it & 1u;
                const int  c0      = cb * 64 + cg * 8;
                const bool lane_ok = c0 < p.E;
                mbar_wait(x_full + 8u * buf, (it >> 1) & 1u);
                const uint8_t * xt = x_ptr + (size_t)buf * box_bytes + cg * 16;
                // tap-outer loop: one tap's 8 weights live at a time, `rows_per` (<= 4) output rows accumulate in registers
                // (keeps the kernel under ~96 registers so that more warps are resident: the loop is latency-bound)
                constexpr int MAXR = 4;
                float acc[MAXR][8];
#pragma unroll
                for (int r = 0; r < MAXR; r++)
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[r][j] = 0.f;
                if (lane_ok) {
#pragma unroll
                    for (int k = 0; k < 9; k++) {
                        const int kh = k / 3, kw = k % 3;
                        float wf[8];
                        {
                            const H8 wv = *reinterpret_cast<const H8 *>(p.dwW + (size_t)k * p.E + c0);
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const float2 f = __half22float2(wv.h[j]);
                                wf[2 * j] = f.x; wf[2 * j + 1] = f.y;
                            }
                        }
#pragma unroll
                        for (int r = 0; r < MAXR; r++) {
                            if (r < rows_per) {
                                const int oyl = rs * rows_per + r;
                                const H8 v = *reinterpret_cast<const H8 *>(xt + ((size_t)(oyl * STRIDE + kh) * p.box_w + (xl * STRIDE + kw)) * 128);
#pragma unroll
                                for (int j = 0; j < 4; j++) {
                                    const float2 f = __half22float2(v.h[j]);
                                    acc[r][2 * j]     = fmaf(f.x, wf[2 * j], acc[r][2 * j]);
                                    acc[r][2 * j + 1] = fmaf(f.y, wf[2 * j + 1], acc[r][2 * j + 1]);
                                }
                            }
                        }
                    }
                }
                float sc[8], sh[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    sc[j] = (lane_ok && p.dw_scale) ? p.dw_scale[c0 + j] : 1.f;
                    sh[j] = (lane_ok && p.dw_shift) ? p.dw_shift[c0 + j] : 0.f;
                }
                mbar_wait(ds_free + 8u * buf, ((it >> 1) & 1u) ^ 1u);  // the MMA that read ds[buf] two items ago has completed
#pragma unroll
                for (int r = 0; r < MAXR; r++) {
                    if (r < rows_per) {
                        const int oyl = rs * rows_per + r;
                        const int q   = oyl * p.TW + xl;  // row of the A operand
                        H8 o;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            float y0 = fmaf(acc[r][2 * j], sc[2 * j], sh[2 * j]);
                            float y1 = fmaf(acc[r][2 * j + 1], sc[2 * j + 1], sh[2 * j + 1]);
                            if (p.dw_act) { y0 = silu_f(y0); y1 = silu_f(y1); }
                            if (!lane_ok) { y0 = 0.f; y1 = 0.f; }
                            o.h[j] = __floats2half2_rn(y0, y1);
                        }
                        const uint32_t * ow = reinterpret_cast<const uint32_t *>(&o);
                        st_shared_v4(ds_base + buf * 16384u + (uint32_t)q * 128u + (((uint32_t)cg ^ ((uint32_t)q & 7u)) << 4), ow[0], ow[1], ow[2], ow[3]);
                    }
                }
                fence_proxy_async();  // generic-proxy writes of ds[buf] -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(ds_full + 8u * buf);
            }
            // ---- epilogue: reduce BN (+ residual); lane quadrant = warp & 3, 32-column chunks alternate between warp halves
            mbar_wait(acc_full, tcount & 1u);
            tc_fence_after();
            {
                const int qd  = warp & 3, half = warp >> 2;
                const int row = qd * 32 + lane;
                const int oyl = row / p.TW, xr = row % p.TW;
                const int oy  = ty * p.TH + oyl, oxx = tx * p.TW + xr;
                const bool ok = row < p.TH * p.TW && oy < p.OH && oxx < p.OW;
                const size_t pix = ((size_t)n * p.OH + oy) * p.OW + oxx;
                for (int c = half; c * 32 < p.Cout_pad; c += 2) {
                    float v[32];
                    tmem_ld_32x32(tmem_acc + ((uint32_t)(qd * 32) << 16) + (uint32_t)(c * 32), v);
                    if (!ok) continue;
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const int nn = c * 32 + g * 8;
                        if (nn + 8 > p.Cout) continue;
                        float y[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const float s_ = p.r_scale ? __ldg(p.r_scale + nn + j) : 1.f;
                            const float h_ = p.r_shift ? __ldg(p.r_shift + nn + j) : 0.f;
                            y[j] = fmaf(v[g * 8 + j], s_, h_);
                            if (p.r_act) y[j] = silu_f(y[j]);
                        }
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
                            H8 o;
#pragma unroll
                            for (int j = 0; j < 4; j++) o.h[j] = __floats2half2_rn(y[2 * j], y[2 * j + 1]);
                            *reinterpret_cast<H8 *>(p.out16 + pix * p.Cout + nn) = o;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_free);  // this warp no longer reads the accumulator
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_acc, (uint32_t)p.tmem_cols);
    }
}

}  // namespace

bool dwreduce_prepare(DwRedLaunch & L, const __half * x, int N, int H, int W, int E, int stride, const __half * dwW,
                      const float * dw_scale, const float * dw_shift, int dw_act, const __half * Wr, int Cout, const float * r_scale,
                      const float * r_shift, int r_act, const float * res32, __half * out16, float * out32) {
    if (E % 8 || Cout % 8 || Cout > 256 || H % stride || W % stride || (stride != 1 && stride != 2)) return false;
    L = DwRedLaunch();
    DwRedLaunch::Params & p = L.p;
    p.N = N; p.H = H; p.W = W; p.E = E; p.stride = stride; p.Cout = Cout;
    p.Cout_pad = (Cout + 15) / 16 * 16;
    p.OH = H / stride; p.OW = W / stride;
    // tile: 8 x 16 output pixels (UMMA M = 128); stride 2 halves it so the input box stays ~44 KiB
    p.TW = p.OW >= 16 ? 16 : (p.OW >= 8 ? 8 : (p.OW >= 4 ? 4 : (p.OW >= 2 ? 2 : 1)));
    const int RS = 32 / p.TW;
    int th = (stride == 1 ? 128 : 64) / p.TW;  // rows so that TH*TW = 128 (or 64)
    if (th % RS) return false;
    p.TH = th;
    if (p.TH * p.TW > 128 || p.TH / RS > 4) return false;
    p.tiles_x = (p.OW + p.TW - 1) / p.TW;
    p.tiles_y = (p.OH + p.TH - 1) / p.TH;
    p.cblocks = (E + 63) / 64;
    p.box_w   = p.TW * stride + 2;
    p.box_h   = p.TH * stride + 2;
    if (p.box_w > 256 || p.box_h > 256) return false;
    p.ntiles = N * p.tiles_y * p.tiles_x;
    p.dwW = dwW; p.dw_scale = dw_scale; p.dw_shift = dw_shift; p.dw_act = dw_act;
    p.r_scale = r_scale; p.r_shift = r_shift; p.r_act = r_act; p.res32 = res32; p.out16 = out16; p.out32 = out32;
    p.tmem_cols = p.Cout_pad <= 32 ? 32 : p.Cout_pad <= 64 ? 64 : p.Cout_pad <= 128 ? 128 : 256;
    {
        const uint64_t dims[4] = {(uint64_t)E, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        const uint64_t str[3]  = {(uint64_t)E * 2, (uint64_t)W * E * 2, (uint64_t)H * W * E * 2};
        const uint32_t box[4]  = {64, (uint32_t)p.box_w, (uint32_t)p.box_h, 1};
        tma_encode(&L.map_x, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    }
    {
        const uint64_t dims[2] = {(uint64_t)E, (uint64_t)Cout};
        const uint64_t str[1]  = {(uint64_t)E * 2};
        const uint32_t box[2]  = {64, (uint32_t)p.Cout_pad};
        tma_encode(&L.map_w, Wr, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    }
    L.smem_bytes = 1024 + 2 * 16384 + 2 * (size_t)p.Cout_pad * 128 + 2 * (size_t)p.box_w * p.box_h * 128;
    if (L.smem_bytes > 220 * 1024) return false;
    const int per_sm = (L.smem_bytes <= 112 * 1024 && p.tmem_cols <= 256) ? 2 : 1;
    const int cap    = per_sm * runtime().sm_count;
    L.grid           = p.ntiles < cap ? p.ntiles : cap;
    return true;
}

void dwreduce_launch(const DwRedLaunch & L, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        B200_CHECK(cudaFuncSetAttribute(k_dwreduce<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
        B200_CHECK(cudaFuncSetAttribute(k_dwreduce<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
        attr = true;
    }
    if (L.p.stride == 1) k_dwreduce<1><<<L.grid, 288, L.smem_bytes, st>>>(L.map_x, L.map_w, L.p);
    else k_dwreduce<2><<<L.grid, 288, L.smem_bytes, st>>>(L.map_x, L.map_w, L.p);
}

}  // namespace b200
