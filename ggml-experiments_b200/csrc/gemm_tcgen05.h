// gemm_tcgen05.h -- host interface of the tcgen05/TMA GEMM + implicit-GEMM 3x3 convolution kernel (K1).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

namespace b200 {

// Fused epilogue applied to the f32 accumulator tile read back from TMEM:
//   y = acc * scale[n] + shift[n]      (folded BatchNorm, or bias with scale == nullptr)
//   y = silu(y)                        (act == 1)
//   y += residual[m, n]                (f32 or f16 residual, optional)
//   store f16 and/or f32, row-major with leading dimension ld*
struct GemmEpilogue {
    const float *  scale = nullptr;
    const float *  shift = nullptr;
    int            act   = 0;
    const float *  res32 = nullptr;
    int            ldr32 = 0;
    const __half * res16 = nullptr;
    int            ldr16 = 0;
    __half *       out16 = nullptr;
    int            ld16  = 0;
    float *        out32 = nullptr;
    int            ld32  = 0;
    // LayerNorm folded around the GEMM (no separate LN kernel): a producer whose tile spans the whole row also writes
    // (sum, sum of squares) of its final values per row; a consumer multiplies RAW activations with gamma-scaled weights and
    // applies  y = r * (acc - mu * c1[n])  before scale/shift, with mu, r from the producer's row statistics and
    // c1[n] = sum_k W'[n][k] (beta and the bias are folded into `shift` on the host).
    float *        stats_out = nullptr;  // [M][2], requires N <= 256 (one tile per row)
    const float *  ln_stats  = nullptr;  // [M][2]
    const float *  ln_c1     = nullptr;  // [N]
    float          ln_inv_c  = 0.f;      // 1 / (channels of the normalised row)
    float          ln_eps    = 0.f;
};

// A prepared launch: tensor maps are encoded once at plan time, the launch is then replayable / capturable.
struct GemmLaunch {
    CUtensorMap map_a0, map_a1, map_b, map_o16, map_o32, map_r32;
    struct Params {
        int M, N, K;
        int block_n, n_tiles, num_kb, stages, tmem_cols;
        int b_resident;             // 1: all K blocks of this CTA's weight tile stay in shared memory (loaded once), the ring holds A only
        int acc_stages;             // TMEM accumulator stages: 2, or 4 for tiles of at most 64 columns (the MMA issuer runs further ahead)
        int kb_elems;               // K elements per smem stage: 64 (128B swizzle), 32 (64B swizzle) or 16 (32B swizzle)
        int a_cp_async;             // 1: A tile loaded by the producer warp with cp.async (K = 16 / 32: TMA rows would be 32-64 B)
        const __half * a_ptr;       // cp.async mode: A base pointer and leading dimension
        int lda;
        int ep_warp;                // 1: every epilogue warp stages and TMA-stores its own 32 rows (no residual slab to share)
        int conv;                   // 0: plain GEMM, 1: 3x3 stride-1 pad-1 implicit GEMM over NHWC (one TMA box per tap), 2: the same with halo
                                    //    boxes (one box per kw, the three kh taps are descriptor offsets into it) and pre-tiled weights
        int ring_bytes;             // bytes of the operand ring(s) in shared memory (the barriers follow)
        int a_slot_bytes, b_stages; // conv == 2: activation slot size, depth of the weight ring (`stages` = depth of the activation ring)
        int pair;                   // conv == 2: 1 = two vertically adjacent 128-pixel M tiles per activation box and weight block (4 accumulators)
        int n_pad;                  // conv == 2: rows of one (channel block, tap) weight block in the halo layout (OC padded to 64)
        const uint8_t * w_halo;     // conv == 2: weights as [channel block][kw][kh][n_pad rows x 64 ch], 128B-swizzled (conv3x3_pack_halo)
        int tile_m;                 // output rows (pixels) per M tile: 128, or less when a conv tile is a whole number of image rows
        int a_tx_bytes;             // conv: bytes one activation box delivers (tile_m * 128)
        int H, W, rows_per_tile;    // conv: image rows covered by one 128-pixel tile (0 if a tile spans whole images)
        int cblk0, cblk1, C0, C1;   // conv: 64-channel blocks / channels of source 0 and source 1 (concat fusion)
        GemmEpilogue ep;
    } p;
    size_t smem_bytes;
    int    ctas_per_sm;
    dim3   grid;
};

// C[M,N] = A[M,K] * B[N,K]^T.  A, B f16, K contiguous (lda/ldb in elements, multiples of 8).
bool gemm_prepare(GemmLaunch & L, const __half * A, int lda, const __half * B, int ldb, int M, int N, int K,
                  const GemmEpilogue & ep);

// 3x3, stride 1, pad 1 convolution over NHWC f16 activations as an implicit GEMM:
//   out[(n,y,x), oc] = sum_{kh,kw,ic} in[n, y+kh-1, x+kw-1, ic] * Wt[oc, kh, kw, ic]
// The input channels may come from two tensors (x0: C0 channels, x1: C1 channels) = a fused ggml_concat.
// Wt is [OC][3][3][C0+C1] f16.  Returns false if the shape cannot be tiled (W must divide 128 or 128 | W*k).
// Wt_halo (optional): the same weights in the halo layout (conv3x3_pack_halo); when given and the map qualifies (W % 8 == 0, W <= 64)
// the kernel runs in halo mode: 3 activation boxes per 64-channel block instead of 9, weights by 1-D bulk copies.
bool conv3x3_prepare(GemmLaunch & L, const __half * x0, int C0, const __half * x1, int C1, int Nimg, int H, int W,
                     const __half * Wt, int OC, const GemmEpilogue & ep, const uint8_t * Wt_halo = nullptr);
// host: Wt [OC][3][3][C0+C1] f16 bits -> halo layout; returns the bytes written to `out` (resized)
size_t conv3x3_pack_halo(const uint16_t * Wt, int OC, int C0, int C1, std::vector<uint8_t> & out);

void gemm_launch(const GemmLaunch & L, cudaStream_t st);

// generic cuTensorMapEncodeTiled wrapper (dims/strides innermost first; strides in bytes for dims 1..rank-1)
void tma_encode(CUtensorMap * map, const void * base, CUtensorMapDataType dtype, int rank, const uint64_t * dims,
                const uint64_t * strides_bytes, const uint32_t * box, CUtensorMapSwizzle swizzle);

// K3 depthwise 3x3 (stride 1|2, pad 1) + BN scale/shift + SiLU over NHWC f16, input tiles staged by TMA
struct DwLaunch {
    CUtensorMap map_x;
    struct Params {
        int N, H, W, C, OH, OW, stride;
        int TW, TH, RS, tiles_x, tiles_y, cblocks, box_w, box_h, ntiles, act;
        const __half * Wt;     // [3][3][C]
        const float *  scale;  // [C] or null
        const float *  shift;
        __half *       out;    // [N, OH, OW, C]
    } p;
    size_t smem_bytes;
    int    grid;
};
bool dw_prepare(DwLaunch & L, const __half * x, int N, int H, int W, int C, int stride, const __half * Wt, const float * scale,
                const float * shift, int act, __half * out);
void dw_launch(const DwLaunch & L, cudaStream_t st);

// K4a: depthwise 3x3 (+BN+SiLU) fused with the following 1x1 reduce convolution (+BN, + f32 residual); see dwreduce.cu
struct DwRedLaunch {
    CUtensorMap map_x, map_w;
    struct Params {
        int N, H, W, E, OH, OW, stride, Cout, Cout_pad;
        int TW, TH, tiles_x, tiles_y, cblocks, box_w, box_h, ntiles, tmem_cols;
        const __half * dwW;       // [3][3][E]
        const float *  dw_scale;  // [E]
        const float *  dw_shift;
        int            dw_act;
        const float *  r_scale;   // [Cout]
        const float *  r_shift;
        int            r_act;
        const float *  res32;     // [N*OH*OW, Cout] or null
        __half *       out16;     // [N*OH*OW, Cout] or null
        float *        out32;
    } p;
    size_t smem_bytes;
    int    grid;
};
bool dwreduce_prepare(DwRedLaunch & L, const __half * x, int N, int H, int W, int E, int stride, const __half * dwW,
                      const float * dw_scale, const float * dw_shift, int dw_act, const __half * Wr, int Cout, const float * r_scale,
                      const float * r_shift, int r_act, const float * res32, __half * out16, float * out32);
void dwreduce_launch(const DwRedLaunch & L, cudaStream_t st);

// K4: the whole inverted-residual block (expand 1x1 -> depthwise 3x3 -> reduce 1x1 [+ residual]) in one kernel; see ir_fused.cu
struct IrLaunch {
    CUtensorMap map_x, map_we, map_wr, map_o32, map_o16, map_r32;
    struct Params {
        int N, H, W, Cin, E, Cout, Cout_pad, stride, OH, OW;
        int TH, TW, IH, IW, P_in, P_out, MBI, MBO, tiles_x, tiles_y, ntiles, nthreads;
        int kb_elems, row_bytes, num_kb, ksteps;  // K blocking of the expand GEMM (x and We tiles)
        int NC;                                   // 64-channel chunks of the expanded width
        int tmem_cols, red_stride;
        int nxb;                                  // input-halo buffers in shared memory: 2 (next tile prefetched a tile ahead) or 1
        uint32_t off_x, x_kb_stride, x_buf_stride, x_tx_bytes, off_e, st16_off, off_a, off_we, we_bytes, we_kb_stride, off_wr, wr_bytes, off_par;
        const __half * dwW;                       // [3][3][E]
        const float *se, *he, *sd, *hd, *sr, *hr; // folded BatchNorm of expand / depthwise / reduce
        const float * res32;                      // [N*OH*OW, Cout] or null
        __half *      out16;
        float *       out32;
        long long *   timing;                     // debug probe: [grid][2][8] clock64 sums (compute thread 0, control thread), or null
    } p;
    size_t smem_bytes;
    int    grid;
};
// x == nullptr: shape query only (returns whether the fused kernel covers the shape, fills the tiling)
bool ir_fused_prepare(IrLaunch & L, const __half * x, int N, int H, int W, int Cin, int E, int Cout, int stride, const __half * We,
                      const float * se, const float * he, const __half * dwW, const float * sd, const float * hd, const __half * Wr,
                      const float * sr, const float * hr, const float * res32, __half * out16, float * out32);
void ir_fused_launch(const IrLaunch & L, cudaStream_t st);

}  // namespace b200
