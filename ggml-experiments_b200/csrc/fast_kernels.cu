// fast_kernels.cu -- memory-bound kernels of the FAST plan: vectorised, coalesced NHWC f16 (SURVEY.md K2,K3,K5,K7,K9).
// Rounding points follow ggml: conv inputs are f16, accumulation and BatchNorm/SiLU/LayerNorm/softmax are f32.
#include "fast_kernels.h"

#include "internal.h"
#include "pdl.cuh"
#include "ptx_sm100.cuh"

namespace b200 {

// x*sigmoid(x) = h + h*tanh(h), h = x/2: one MUFU op instead of two (see gemm_tcgen05.cu::silu_f)
__device__ __forceinline__ float silu_fast(float x) {
#ifdef GGML_B200_SILU_EXACT
    return __fdividef(x, 1.0f + __expf(-x));
#else
    const float h = 0.5f * x;
    float       t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
#endif
}

struct alignas(16) Half8 {
    __half2 h[4];
};
__device__ __forceinline__ void unpack8(const Half8 & v, float * f) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float2 t     = __half22float2(v.h[i]);
        f[2 * i]     = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ Half8 pack8(const float * f) {
    Half8 v;
#pragma unroll
    for (int i = 0; i < 4; i++) v.h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    return v;
}

// ---------------------------------------------------------------------------------------------------------
// K2 stem: replaces ggml_conv_2d (im2col f16 + mul_mat) + BN chain + silu of conv_stem (main.cpp:618,771-852)
// One thread = one output pixel, all OC channels (OC <= 32).  HBM-bound: 12 B/pixel in (x9 from L1/L2), 2*OC B out.
// ---------------------------------------------------------------------------------------------------------
// Block = 16 x 32 output pixels of one image (thread = 2 pixels, rows ty and ty+8).  The 33 x 65 x 3 input patch is
// staged once in shared memory (coalesced, rounded to f16 like ggml's im2col), weights are broadcast LDS.128 reads
// shared by both pixels of the thread.
template <int OC>
__global__ void __launch_bounds__(256) k_stem(const float * __restrict__ x, int64_t sn, int64_t sy, int64_t sx, int64_t sc, int N,
                                              int H, int W, const __half * __restrict__ Wt, const float * __restrict__ scale,
                                              const float * __restrict__ shift, int act, __half * __restrict__ out16,
                                              float * __restrict__ out32, int tiles_x, int tiles_y) {
    constexpr int TH = 16, TW = 32, IH = 2 * TH + 1, IW = 2 * TW + 1, IWS = IW * 3 + 1;
    __shared__ __align__(16) float sw[27 * OC];  // [tap*3+ic][oc]
    __shared__ float ss[OC], sh[OC];
    __shared__ float sx_[IH * IWS];
    for (int i = threadIdx.x; i < 27 * OC; i += blockDim.x) {
        int k = i / OC, oc = i % OC;  // k = (kh*3+kw)*3+ic ; Wt is [oc][kh][kw][ic]
        sw[i] = __half2float(Wt[oc * 27 + k]);
    }
    for (int i = threadIdx.x; i < OC; i += blockDim.x) {
        ss[i] = scale ? scale[i] : 1.f;
        sh[i] = shift ? shift[i] : 0.f;
    }
    const int OH = H / 2, OW = W / 2;
    int       b  = blockIdx.x;
    const int tx0 = (b % tiles_x) * TW; b /= tiles_x;
    const int ty0 = (b % tiles_y) * TH;
    const int n   = b / tiles_y;
    const int iy0 = 2 * ty0 - 1, ix0 = 2 * tx0 - 1;
    const float * xn = x + n * sn;
    for (int i = threadIdx.x; i < IH * IW * 3; i += blockDim.x) {
        const int c = i % 3, xx = (i / 3) % IW, yy = i / (3 * IW);
        const int iy = iy0 + yy, ix = ix0 + xx;
        float v = 0.f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __half2float(__float2half_rn(xn[iy * sy + ix * sx + c * sc]));  // ggml im2col rounds to f16
        sx_[yy * IWS + xx * 3 + c] = v;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    float acc[2][OC];
#pragma unroll
    for (int o = 0; o < OC; o++) acc[0][o] = acc[1][o] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; kh++)
#pragma unroll
        for (int kw = 0; kw < 3; kw++)
#pragma unroll
            for (int ic = 0; ic < 3; ic++) {
                const float v0 = sx_[(2 * ty + kh) * IWS + (2 * tx + kw) * 3 + ic];
                const float v1 = sx_[(2 * (ty + 8) + kh) * IWS + (2 * tx + kw) * 3 + ic];
                const float4 * w4 = reinterpret_cast<const float4 *>(&sw[((kh * 3 + kw) * 3 + ic) * OC]);
#pragma unroll
                for (int o4 = 0; o4 < OC / 4; o4++) {
                    const float4 w = w4[o4];
                    acc[0][4 * o4 + 0] = fmaf(v0, w.x, acc[0][4 * o4 + 0]); acc[1][4 * o4 + 0] = fmaf(v1, w.x, acc[1][4 * o4 + 0]);
                    acc[0][4 * o4 + 1] = fmaf(v0, w.y, acc[0][4 * o4 + 1]); acc[1][4 * o4 + 1] = fmaf(v1, w.y, acc[1][4 * o4 + 1]);
                    acc[0][4 * o4 + 2] = fmaf(v0, w.z, acc[0][4 * o4 + 2]); acc[1][4 * o4 + 2] = fmaf(v1, w.z, acc[1][4 * o4 + 2]);
                    acc[0][4 * o4 + 3] = fmaf(v0, w.w, acc[0][4 * o4 + 3]); acc[1][4 * o4 + 3] = fmaf(v1, w.w, acc[1][4 * o4 + 3]);
                }
            }
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int oy = ty0 + ty + 8 * r, ox = tx0 + tx;
        if (oy >= OH || ox >= OW) continue;
#pragma unroll
        for (int o = 0; o < OC; o++) {
            float y   = fmaf(acc[r][o], ss[o], sh[o]);
            acc[r][o] = act ? silu_fast(y) : y;
        }
        const int64_t p = ((int64_t)n * OH + oy) * OW + ox;
        if (out16) {
            Half8 * o = reinterpret_cast<Half8 *>(out16 + p * OC);
#pragma unroll
            for (int g = 0; g < OC / 8; g++) o[g] = pack8(acc[r] + g * 8);
        }
        if (out32) {
            float4 * o = reinterpret_cast<float4 *>(out32 + p * OC);
#pragma unroll
            for (int g = 0; g < OC / 4; g++) o[g] = make_float4(acc[r][4 * g], acc[r][4 * g + 1], acc[r][4 * g + 2], acc[r][4 * g + 3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// K2 on tensor cores.  The stem is a GEMM [pixels x 27] . [27 x OC]; on CUDA cores it needs 27*OC FMAs per pixel and ran at
// 1.06 TB/s (issue-bound).  Here K is laid out as k' = kh*10 + (kw*3+ic) (9 taps + 1 zero-weight pad per input row, 30 -> 32),
// so every even/odd K pair is one aligned 32-bit word of the f16-staged input patch: an A fragment of mma.m16n8k16 is 8 LDS.32,
// no im2col buffer, no shuffles.  Block = 16 x 32 output pixels, warp = 4 m-tiles of 16 consecutive pixels.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int OC>
__global__ void __launch_bounds__(256) k_stem_mma(const float * __restrict__ x, int64_t sn, int64_t sy, int64_t sx, int64_t sc, int N, int H, int W,
                                                  const __half * __restrict__ Wt, const float * __restrict__ scale, const float * __restrict__ shift,
                                                  int act, __half * __restrict__ out16, int tiles_x, int tiles_y,
                                                  const uint8_t * __restrict__ x8, const int * __restrict__ use_x8) {
    constexpr int TH = 16, TW = 32, IH = 2 * TH + 1, IW = 2 * TW + 1, RS = 200, NT = OC / 8;  // RS: halves per staged row (195 used)
    __shared__ __half s_lut[256];
    __shared__ __align__(16) __half s_in[(IH + 1) * RS + 8];  // +1 row: the zero-weight K pad (k' = 30, 31) reads input row 2*py + 3
    __shared__ __align__(16) __half s_out[8][16 * OC];
    __shared__ float ss[OC], sh[OC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int i = threadIdx.x; i < OC; i += blockDim.x) {  // SiLU works on y/2: folded into scale / shift, exactly
        ss[i] = (scale ? scale[i] : 1.f) * (act ? 0.5f : 1.f);
        sh[i] = (shift ? shift[i] : 0.f) * (act ? 0.5f : 1.f);
    }
    // B fragments (weights), constant per thread: bfrag[j][s][h] holds W[oc = 8j+g][k' = 16s + 2t + 8h, +1]
    uint32_t bfrag[NT][2][2];
    int      koff[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int kp = 2 * t + 8 * q;
        koff[q]      = (kp / 10) * RS + (kp % 10);
    }
#pragma unroll
    for (int j = 0; j < NT; j++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            __half2 w2;
            __half  wv[2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int kp = 2 * t + 8 * q + e, kh = kp / 10, jj = kp % 10;
                wv[e]        = (kh < 3 && jj < 9) ? Wt[(8 * j + g) * 27 + kh * 9 + jj] : __float2half_rn(0.f);  // Wt is [oc][kh][kw][ic]
            }
            w2                     = __halves2half2(wv[0], wv[1]);
            bfrag[j][q >> 1][q & 1] = *reinterpret_cast<uint32_t *>(&w2);
        }
    const int OH = H / 2, OW = W / 2;
    int       b  = blockIdx.x;
    const int tx0 = (b % tiles_x) * TW; b /= tiles_x;
    const int ty0 = (b % tiles_y) * TH;
    const int n   = b / tiles_y;
    const int iy0 = 2 * ty0 - 1, ix0 = 2 * tx0 - 1;
    const float * xn = x + n * sn;
    pdl_wait();  // PDL: weights / scale / shift above are constants; the image and the output are not
    pdl_trigger();
    if (use_x8 != nullptr && __ldg(use_x8) != 0) {
        // u8 source (the quantised image of sam_image_preprocess, main.cpp:592-597, straight into the stem): f16(v / 255) comes
        // from a 256-entry table built with the reference's own expression, so the staged patch is bit-identical to the f32 route
        // at a quarter of the bytes.  A row of the patch is 195 bytes starting 3 bytes before a 4-byte boundary (W % 4 == 0, tx0 % 32
        // == 0): lanes load aligned 32-bit words (51 per row, all of a warp's rows in flight), words are wholly inside or outside
        // the image row; words 49 and 50 (bytes 195..) are the zero tail of the staged row.
        s_lut[threadIdx.x] = __float2half_rn(__fdiv_rn((float)threadIdx.x, 255.0f));
        __syncthreads();
        const uint8_t * xn8 = x8 + (int64_t)n * H * W * 3;
        uint32_t wv[5][2];
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const int  yy = warp + 8 * r, iy = iy0 + yy;
            const bool row_ok = yy < IH && iy >= 0 && iy < H;
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int w = lane + 32 * u, cb = 6 * tx0 - 4 + 4 * w;  // byte offset of the word inside the image row
                wv[r][u]    = (row_ok && w < 49 && cb >= 0 && cb < 3 * W) ? __ldg(reinterpret_cast<const uint32_t *>(xn8 + (int64_t)iy * W * 3 + cb)) : 0u;
            }
        }
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const int yy = warp + 8 * r;
            if (yy >= IH + 1) continue;
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int w = lane + 32 * u;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int i = 4 * w + j - 1;
                    if (w < 51 && i >= 0 && i < RS) s_in[yy * RS + i] = s_lut[(wv[r][u] >> (8 * j)) & 255u];
                }
            }
        }
    } else if (sc == 1 && sx == 3 && sy == 3 * (int64_t)W && (W & 3) == 0 && (sn & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        // packed HWC f32 images: the same 51 aligned words per row, 16 bytes each (the scalar loop below spent ~9 instructions per
        // element on index arithmetic, a convert and a 2-byte store: the kernel was issue-bound at 2.6 TB/s)
#pragma unroll
        for (int r0 = 0; r0 < 5; r0 += 3) {  // rows of this warp in two batches (3 + 2): every load of a batch is in flight before its first convert
            float4 fv[3][2];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const int  yy = warp + 8 * (r0 + r), iy = iy0 + yy;
                const bool row_ok = r0 + r < 5 && yy < IH && iy >= 0 && iy < H;
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int w = lane + 32 * u, cf = 6 * tx0 - 4 + 4 * w;  // float offset of the word inside the image row
                    fv[r][u] = (row_ok && w < 49 && cf >= 0 && cf < 3 * W) ? __ldg(reinterpret_cast<const float4 *>(xn + (int64_t)iy * sy + cf)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const int yy = warp + 8 * (r0 + r);
                if (r0 + r >= 5 || yy >= IH + 1) continue;
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int w = lane + 32 * u;
                    if (w >= 51) continue;
                    __half * d = s_in + yy * RS + 4 * w;  // elements 4w-1 .. 4w+2
                    if (w > 0) d[-1] = __float2half_rn(fv[r][u].x);
                    if (w < 50) {
                        *reinterpret_cast<__half2 *>(d) = __floats2half2_rn(fv[r][u].y, fv[r][u].z);
                        d[2] = __float2half_rn(fv[r][u].w);
                    }
                }
            }
        }
    } else {
        // stage the (2*TH+2) x (2*TW+1) x 3 patch as f16 (ggml's im2col rounding point); warp = row, lanes = (x, c) pairs
        for (int yy = warp; yy < IH + 1; yy += 8) {
            const int  iy     = iy0 + yy;
            const bool row_ok = iy >= 0 && iy < H && yy < IH;
            float      v[7];  // all loads of a row are in flight before the first convert/store (the rolled loop exposed one global latency per element)
    #pragma unroll
            for (int u = 0; u < 7; u++) {
                const int i = lane + 32 * u, xx = i / 3, c = i - 3 * xx, ix = ix0 + xx;
                v[u]        = (row_ok && xx < IW && ix >= 0 && ix < W) ? __ldg(xn + iy * sy + ix * sx + c * sc) : 0.f;
            }
    #pragma unroll
            for (int u = 0; u < 7; u++) {
                const int i = lane + 32 * u;
                if (i < RS) s_in[yy * RS + i] = __float2half_rn(v[u]);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int mt = 0; mt < 4; mt++) {
        const int py = 2 * warp + (mt >> 1), x0 = (mt & 1) * 16;
        const __half * base = s_in + (2 * py) * RS + 6 * (x0 + g);
        float acc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; j++) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
        for (int s2 = 0; s2 < 2; s2++) {
            uint32_t a[4];
            a[0] = *reinterpret_cast<const uint32_t *>(base + koff[2 * s2]);
            a[1] = *reinterpret_cast<const uint32_t *>(base + 48 + koff[2 * s2]);
            a[2] = *reinterpret_cast<const uint32_t *>(base + koff[2 * s2 + 1]);
            a[3] = *reinterpret_cast<const uint32_t *>(base + 48 + koff[2 * s2 + 1]);
#pragma unroll
            for (int j = 0; j < NT; j++) mma_16816(acc[j], a, bfrag[j][s2]);
        }
        // BN + SiLU, pack, stage the warp's 16 x OC tile so that it leaves as 16-byte coalesced stores
        __half * so = s_out[warp];
#pragma unroll
        for (int j = 0; j < NT; j++) {
            const int   oc = 8 * j + 2 * t;
            const float s0 = ss[oc], s1 = ss[oc + 1], h0 = sh[oc], h1 = sh[oc + 1];
            float y0 = fmaf(acc[j][0], s0, h0), y1 = fmaf(acc[j][1], s1, h1), y2 = fmaf(acc[j][2], s0, h0), y3 = fmaf(acc[j][3], s1, h1);
            if (act) { y0 = ptx::silu_h(y0); y1 = ptx::silu_h(y1); y2 = ptx::silu_h(y2); y3 = ptx::silu_h(y3); }
            *reinterpret_cast<__half2 *>(so + g * OC + oc)       = __floats2half2_rn(y0, y1);
            *reinterpret_cast<__half2 *>(so + (g + 8) * OC + oc) = __floats2half2_rn(y2, y3);
        }
        __syncwarp();
        const int oy = ty0 + py;
        for (int ch = lane; ch < 2 * OC; ch += 32) {  // 16 * OC halves = 2 * OC chunks of 8
            const int pix = (ch * 8) / OC, ox = tx0 + x0 + pix;
            if (oy < OH && ox < OW)
                *reinterpret_cast<uint4 *>(out16 + (((int64_t)n * OH + oy) * OW + ox) * OC + (ch * 8) % OC) = *reinterpret_cast<const uint4 *>(so + ch * 8);
        }
        __syncwarp();
    }
}

bool stem_takes_u8(int OC, int W, bool hwc, bool out16, bool out32) {
    static const bool v1 = getenv("GGML_B200_STEM_V1") != nullptr;
    return !v1 && out16 && !out32 && (OC == 8 || OC == 16 || OC == 24 || OC == 32) && hwc && W % 4 == 0;
}
void launch_stem(const float * x, int64_t sn, int64_t sy, int64_t sx, int64_t sc, int N, int H, int W, const __half * Wt, int OC,
                 const float * scale, const float * shift, int act, __half * out16, float * out32, cudaStream_t st, const uint8_t * x8, const int * use_x8) {
    const int tiles_x = (W / 2 + 31) / 32, tiles_y = (H / 2 + 15) / 16;
    const int grid    = N * tiles_x * tiles_y;
    if (stem_takes_u8(OC, 4, true, out16 != nullptr, out32 != nullptr)) {  // W / layout only matter for the u8 source
        switch (OC) {
            case 8: launch_pdl(k_stem_mma<8>, dim3(grid), dim3(256), 0, st, x, sn, sy, sx, sc, N, H, W, Wt, scale, shift, act, out16, tiles_x, tiles_y, x8, use_x8); return;
            case 16: launch_pdl(k_stem_mma<16>, dim3(grid), dim3(256), 0, st, x, sn, sy, sx, sc, N, H, W, Wt, scale, shift, act, out16, tiles_x, tiles_y, x8, use_x8); return;
            case 24: launch_pdl(k_stem_mma<24>, dim3(grid), dim3(256), 0, st, x, sn, sy, sx, sc, N, H, W, Wt, scale, shift, act, out16, tiles_x, tiles_y, x8, use_x8); return;
            case 32: launch_pdl(k_stem_mma<32>, dim3(grid), dim3(256), 0, st, x, sn, sy, sx, sc, N, H, W, Wt, scale, shift, act, out16, tiles_x, tiles_y, x8, use_x8); return;
            default: break;
        }
    }
    switch (OC) {
        case 8: k_stem<8><<<grid, 256, 0, st>>>(x, sn, sy, sx, sc, N, H, W, Wt, scale, shift, act, out16, out32, tiles_x, tiles_y); break;
        case 16: k_stem<16><<<grid, 256, 0, st>>>(x, sn, sy, sx, sc, N, H, W, Wt, scale, shift, act, out16, out32, tiles_x, tiles_y); break;
        case 24: k_stem<24><<<grid, 256, 0, st>>>(x, sn, sy, sx, sc, N, H, W, Wt, scale, shift, act, out16, out32, tiles_x, tiles_y); break;
        case 32: k_stem<32><<<grid, 256, 0, st>>>(x, sn, sy, sx, sc, N, H, W, Wt, scale, shift, act, out16, out32, tiles_x, tiles_y); break;
        default: B200_ABORT("stem: unsupported OC %d", OC);
    }
}

// ---------------------------------------------------------------------------------------------------------
// K3 depthwise 3x3 + BN + SiLU: replaces ggml_conv_depthwise_2d + BN chain + silu (main.cpp:788,809-850).
// thread = (8-channel group, output pixel); 128-bit loads/stores; adjacent threads -> adjacent channel groups of the
// same pixel -> fully coalesced rows of C*2 bytes.  Per-thread weights/scale/shift live in registers across pixels.
// ---------------------------------------------------------------------------------------------------------
// A block owns a strip of TW output columns x TH output rows of one image; a thread owns one (column, channel group)
// and walks down the strip with the 3x3 input window in registers, so each output row costs 3 (stride 1) or 6
// (stride 2) new 128-bit loads instead of 9; the x-neighbour overlap between threads is served by L1.
template <int STRIDE>
__global__ void __launch_bounds__(256) k_dwconv(const __half * __restrict__ x, int H, int W, int C, const __half * __restrict__ Wt,
                                                const float * __restrict__ scale, const float * __restrict__ shift, int act,
                                                __half * __restrict__ out, int TW, int TH, int tiles_x, int tiles_y) {
    const int cgs = C >> 3;                  // 8-channel groups
    const int cg  = threadIdx.x % cgs;       // blockDim.x == cgs * TW
    const int xl  = threadIdx.x / cgs;
    const int OH = H / STRIDE, OW = W / STRIDE;  // pad 1, k 3
    int       b  = blockIdx.x;
    const int tx = b % tiles_x;
    b /= tiles_x;
    const int ty = b % tiles_y;
    const int n  = b / tiles_y;
    const int ox = tx * TW + xl;
    if (ox >= OW) return;
    const int oy0 = ty * TH, oy1 = min(OH, oy0 + TH);

    float w[9][8], sc[8], sh[8];
#pragma unroll
    for (int t = 0; t < 9; t++) unpack8(*reinterpret_cast<const Half8 *>(Wt + t * C + cg * 8), w[t]);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        sc[j] = scale ? scale[cg * 8 + j] : 1.f;
        sh[j] = shift ? shift[cg * 8 + j] : 0.f;
    }
    const __half * xin  = x + (int64_t)n * H * W * C + cg * 8;
    __half *       outp = out + (int64_t)n * OH * OW * C + cg * 8;
    const int  ix1 = ox * STRIDE, ix0 = ix1 - 1, ix2 = ix1 + 1;
    const bool v0 = ix0 >= 0, v2 = ix2 < W;
    Half8 zero;
#pragma unroll
    for (int i = 0; i < 4; i++) zero.h[i] = __floats2half2_rn(0.f, 0.f);
    auto load_row = [&](int iy, Half8 * r) {
        if (iy < 0 || iy >= H) {
            r[0] = zero; r[1] = zero; r[2] = zero;
        } else {
            const __half * row = xin + (int64_t)iy * W * C;
            r[0] = v0 ? *reinterpret_cast<const Half8 *>(row + ix0 * C) : zero;
            r[1] = *reinterpret_cast<const Half8 *>(row + ix1 * C);
            r[2] = v2 ? *reinterpret_cast<const Half8 *>(row + ix2 * C) : zero;
        }
    };
    Half8 win[3][3];
    load_row(oy0 * STRIDE - 1, win[0]);
    if (STRIDE == 1) load_row(oy0, win[1]);
    for (int oy = oy0; oy < oy1; oy++) {
        if (STRIDE == 1) {
            load_row(oy + 1, win[2]);
        } else {
            load_row(oy * 2, win[1]);
            load_row(oy * 2 + 1, win[2]);
        }
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; kh++)
#pragma unroll
            for (int kw = 0; kw < 3; kw++) {
                float v[8];
                unpack8(win[kh][kw], v);
#pragma unroll
                for (int j = 0; j < 8; j++) acc[j] = fmaf(v[j], w[kh * 3 + kw][j], acc[j]);
            }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            float y = fmaf(acc[j], sc[j], sh[j]);
            acc[j]  = act ? silu_fast(y) : y;
        }
        *reinterpret_cast<Half8 *>(outp + ((int64_t)oy * OW + ox) * C) = pack8(acc);
#pragma unroll
        for (int kw = 0; kw < 3; kw++) {
            if (STRIDE == 1) {
                win[0][kw] = win[1][kw];
                win[1][kw] = win[2][kw];
            } else {
                win[0][kw] = win[2][kw];
            }
        }
    }
}

void launch_dwconv(const __half * x, int N, int H, int W, int C, int stride, const __half * Wt, const float * scale,
                   const float * shift, int act, __half * out16, cudaStream_t st) {
    if (C % 8 || H % stride || W % stride) B200_ABORT("dwconv: unsupported shape C=%d H=%d W=%d stride=%d", C, H, W, stride);
    const int cgs = C / 8;
    const int OH = H / stride, OW = W / stride;
    int TW = 256 / cgs;
    if (TW < 1) TW = 1;
    if (TW > OW) TW = OW;
    if (cgs * TW > 256) B200_ABORT("dwconv: C=%d too wide", C);
    const int TH      = OH < 16 ? OH : 16;
    const int tiles_x = (OW + TW - 1) / TW, tiles_y = (OH + TH - 1) / TH;
    const int grid    = N * tiles_x * tiles_y;
    if (stride == 1) k_dwconv<1><<<grid, cgs * TW, 0, st>>>(x, H, W, C, Wt, scale, shift, act, out16, TW, TH, tiles_x, tiles_y);
    else k_dwconv<2><<<grid, cgs * TW, 0, st>>>(x, H, W, C, Wt, scale, shift, act, out16, TW, TH, tiles_x, tiles_y);
}

// ---------------------------------------------------------------------------------------------------------
// K5 LayerNorm: replaces ggml_norm + mul(gamma) + add(beta) + cont (main.cpp:1002-1019,1114-1131,1192-1209).
// One warp per row, the row lives in registers (C <= 512), two-pass mean / variance with warp-shuffle reductions.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 8 lanes per row (4 rows per warp): C/4 float4 per row split over 8 lanes keeps every lane busy for C = 144..256
// (one full warp per row left 28 of 32 lanes idle on the second pass of a 36-float4 row and made the kernel issue-bound).
template <int MAXV>  // float4 per lane = ceil(C / 32), compile-time so no predicated-off iterations are issued
__global__ void __launch_bounds__(256) k_layernorm(const float * __restrict__ x, int64_t rows, int C, const float * __restrict__ gamma,
                                                   const float * __restrict__ beta, float eps, __half * __restrict__ out16,
                                                   float * __restrict__ out32) {
    constexpr int LPR  = 8;   // lanes per row
    const int     sub  = threadIdx.x & (LPR - 1);
    const int64_t row  = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPR;
    const bool    live = row < rows;
    const float * xr   = x + (live ? row : 0) * C;
    const int     nv   = C >> 2;
    float4        v[MAXV];
    float         s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
        const int idx = sub + i * LPR;
        if (idx < nv) {
            v[i] = live ? reinterpret_cast<const float4 *>(xr)[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float       s2   = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
        const int idx = sub + i * LPR;
        if (idx < nv) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            s2 += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    const float rstd = rsqrtf(s2 / (float)C + eps);
    if (!live) return;
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
        const int idx = sub + i * LPR;
        if (idx < nv) {
            const float4 g = reinterpret_cast<const float4 *>(gamma)[idx];
            const float4 b = reinterpret_cast<const float4 *>(beta)[idx];
            float4       y;
            y.x = fmaf(v[i].x * rstd, g.x, b.x);
            y.y = fmaf(v[i].y * rstd, g.y, b.y);
            y.z = fmaf(v[i].z * rstd, g.z, b.z);
            y.w = fmaf(v[i].w * rstd, g.w, b.w);
            if (out32) reinterpret_cast<float4 *>(out32 + row * C)[idx] = y;
            if (out16) {
                uint2 o;
                __half2 h0 = __floats2half2_rn(y.x, y.y), h1 = __floats2half2_rn(y.z, y.w);
                o.x = *reinterpret_cast<uint32_t *>(&h0);
                o.y = *reinterpret_cast<uint32_t *>(&h1);
                reinterpret_cast<uint2 *>(out16 + row * C)[idx] = o;
            }
        }
    }
}

void launch_layernorm(const float * x, int64_t rows, int C, const float * gamma, const float * beta, float eps, __half * out16,
                      float * out32, cudaStream_t st) {
    if (C % 4 || C > 512) B200_ABORT("layernorm: unsupported C %d", C);
    const int grid = (int)((rows + 31) / 32);  // 256 threads = 32 rows
    const int nvl  = (C / 4 + 7) / 8;
#define LN_CASE(V) case V: k_layernorm<V><<<grid, 256, 0, st>>>(x, rows, C, gamma, beta, eps, out16, out32); break;
    switch (nvl) {
        LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
        LN_CASE(9) LN_CASE(10) LN_CASE(11) LN_CASE(12) LN_CASE(13) LN_CASE(14) LN_CASE(15) LN_CASE(16)
        default: B200_ABORT("layernorm: unsupported C %d", C);
    }
#undef LN_CASE
}

// ---------------------------------------------------------------------------------------------------------
// K7 attention: replaces mul_mat(K,Q) / sqrt(d) -> soft_max -> mul_mat(V^T, P) and the surrounding
// permutes/conts (main.cpp:975-986,1073-1093).
//
// qkv layout (written by the QKV GEMM against zero-padded weights): [token][3][heads][DP] f16, DP = head dim
// rounded up to 16, padding lanes are exact zeros, so every per-head row is a 16-byte aligned MMA operand.
// A sequence = the (H/2)*(W/2) pixels sharing (y%2, x%2) of one image: the unfold is only index arithmetic.
//
// k_attention_mma: flash-style, one warp = 16 query rows, S = Q K^T and O += P V on the tensor cores through
// mma.sync.m16n8k16 (f16 in, f32 accumulate), softmax in registers in the accumulator layout with quad shuffles.
// (mma.sync, not tcgen05: L <= 1024, d <= 64 tiles are far below a UMMA 128xN tile; see DESIGN.md.)
// k_attention_v1 (CUDA cores) remains for sequence lengths that are not a multiple of 16.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4(uint32_t * r, const void * p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t * r, const void * p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_16816(float * c, const uint32_t * a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void * dst, const void * src) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ float ex2_approx(float x) {  // one MUFU; exp2f() adds a range check and two fix-up multiplies per call
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// One block of up to 64 keys for a warp's 16 queries: S = Q K^T, online softmax, O += P V.  NTC = 8 compiles the full
// block without predicates (every sequence length of the model except L = 16); NTC = 0 takes the tile count at run time.
template <int DP, int NTC>
__device__ __forceinline__ void attn_block(const uint32_t (&qf)[DP / 16][4], float (&o)[DP / 8][4], float (&m)[2], float (&l)[2], uint32_t kaddr,
                                           uint32_t vaddr, float sl2, int ntile_rt) {
    constexpr int LDS = DP + 8, KS = DP / 16, DT = DP / 8;
    const int ntile = NTC ? NTC : ntile_rt;
    float     s[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
#pragma unroll
        for (int np = 0; np < 4; np++) {
            if (np * 2 < ntile) {
                uint32_t b[4];
                ldsm_x4(b, kaddr + (uint32_t)((np * 16 * LDS + ks * 16) * 2));
                mma_16816(s[np * 2], qf[ks], b[0], b[1]);
                mma_16816(s[np * 2 + 1], qf[ks], b[2], b[3]);
            }
        }
    }
    // ---- online softmax (rows g and g+8 of this warp's 16 queries) ----
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < 8; i++)
        if (i < ntile) {
            mx[0] = fmaxf(mx[0], fmaxf(s[i][0], s[i][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[i][2], s[i][3]));
        }
    float mb[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        const float mn   = fmaxf(m[r], mx[r]);
        const float corr = ex2_approx((m[r] - mn) * sl2);
        m[r]             = mn;
        mb[r]            = mn * sl2;
        l[r] *= corr;
#pragma unroll
        for (int i = 0; i < DT; i++) {
            o[i][2 * r] *= corr;
            o[i][2 * r + 1] *= corr;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++)
        if (i < ntile) {
            s[i][0] = ex2_approx(fmaf(s[i][0], sl2, -mb[0]));
            s[i][1] = ex2_approx(fmaf(s[i][1], sl2, -mb[0]));
            s[i][2] = ex2_approx(fmaf(s[i][2], sl2, -mb[1]));
            s[i][3] = ex2_approx(fmaf(s[i][3], sl2, -mb[1]));
            l[0] += s[i][0] + s[i][1];
            l[1] += s[i][2] + s[i][3];
        }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
        if (kk * 2 < ntile) {
            uint32_t a[4];
            a[0] = pack_half2(s[2 * kk][0], s[2 * kk][1]);
            a[1] = pack_half2(s[2 * kk][2], s[2 * kk][3]);
            a[2] = pack_half2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
            a[3] = pack_half2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
            for (int dp = 0; dp < DT / 2; dp++) {
                uint32_t b[4];
                ldsm_x4_trans(b, vaddr + (uint32_t)((kk * 16 * LDS + dp * 16) * 2));
                mma_16816(o[dp * 2], a, b[0], b[1]);
                mma_16816(o[dp * 2 + 1], a, b[2], b[3]);
            }
        }
    }
}

constexpr int kAttnQB = 128;  // queries per CTA (8 warps x 16)
constexpr int kAttnLK = 256;  // keys staged in shared memory at a time

template <int DP>
__global__ void __launch_bounds__(256) k_attention_mma(const __half * __restrict__ qkv, int H, int W, int C, int heads, int d,
                                                       __half * __restrict__ out, float sl2, int lk, int qblocks) {
    constexpr int LDS = DP + 8;  // padded smem row (halves): conflict-free ldmatrix
    constexpr int KS  = DP / 16; // k-steps of Q K^T
    constexpr int DT  = DP / 8;  // n-tiles of P V
    extern __shared__ __align__(16) unsigned char smem_attn[];
    // qblocks > 1 (whole sequence staged once, L <= kAttnLK): the CTA walks all 128-query blocks of its (sequence, head) against
    // one copy of K/V -- half the K/V loads and half the load-phase stalls per unit of work at L = 256
    const int qrows = kAttnQB * qblocks;
    __half * sQ = reinterpret_cast<__half *>(smem_attn);
    __half * sK = sQ + qrows * LDS;
    __half * sV = sK + lk * LDS;  // lk = min(L, kAttnLK) keys staged at a time

    const int npw = W >> 1, nph = H >> 1, L = npw * nph;
    const int head = blockIdx.x % heads;
    const int pp   = (blockIdx.x / heads) & 3;
    const int n    = blockIdx.x / (heads * 4);
    const int ph = pp >> 1, pw = pp & 1;
    const int64_t ld = 3 * (int64_t)heads * DP;
    auto tok = [&](int l) -> int64_t {
        const int iph = l / npw, ipw = l - iph * npw;
        return ((int64_t)n * H + (iph * 2 + ph)) * W + (ipw * 2 + pw);
    };
    const int qbase = blockIdx.y * qrows;
    const int nqall = min(qrows, L - qbase);
    pdl_wait();  // PDL: q/k/v come from the previous kernel
    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;

    // Loads: thread = (row slot = tid / 8, 16-byte chunk = tid % 8; chunks >= DP/8 idle), blockDim/8 rows per pass.  The token index of
    // a row (an integer division by the patch-grid width) is computed once and then advanced incrementally: the old
    // i / (DP/8), i % (DP/8), l / npw per 16-byte copy was a third of all instructions executed by this kernel.
    const int  lrow = threadIdx.x >> 3, lchunk = threadIdx.x & 7;
    const bool lact = lchunk < DP / 8;
    const int  rpp    = blockDim.x >> 3;                        // rows per pass (the block has 32 threads per 16 queries)
    const int  step_h = rpp / npw, step_w = rpp - step_h * npw;  // rpp rows further = step_h grid rows + step_w columns
    auto row_token = [&](int iph, int ipw) -> int64_t { return ((int64_t)n * H + (iph * 2 + ph)) * W + (ipw * 2 + pw); };
    {
        int l0 = qbase + lrow, iph = l0 / npw, ipw = l0 - iph * npw;
        for (int r = lrow; r < nqall; r += rpp) {
            if (lact) cp_async16(sQ + r * LDS + lchunk * 8, qkv + row_token(iph, ipw) * ld + head * DP + lchunk * 8);
            iph += step_h; ipw += step_w;
            if (ipw >= npw) { ipw -= npw; iph++; }
        }
    }

    uint32_t qf[KS][4];
    float    o[DT][4];
    float    m[2], l[2];
    auto reset = [&]() {
        m[0] = m[1] = -INFINITY;
        l[0] = l[1] = 0.f;
#pragma unroll
        for (int i = 0; i < DT; i++) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    };
    auto finish = [&](int q0) {  // normalise and store this warp's 16 query rows
#pragma unroll
        for (int r = 0; r < 2; r++) {
            l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
            l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
            l[r] = 1.f / l[r];
        }
        const int qa = q0 + warp * 16 + g, qb = qa + 8;
        __half *  oa = out + tok(qa) * (int64_t)C + head * d;
        __half *  ob = out + tok(qb) * (int64_t)C + head * d;
#pragma unroll
        for (int i = 0; i < DT; i++) {
            const int col = i * 8 + 2 * t;
            if (col < d) {
                *reinterpret_cast<__half2 *>(oa + col) = __floats2half2_rn(o[i][0] * l[0], o[i][1] * l[0]);
                *reinterpret_cast<__half2 *>(ob + col) = __floats2half2_rn(o[i][2] * l[1], o[i][3] * l[1]);
            }
        }
    };
    reset();

    for (int kc0 = 0; kc0 < L; kc0 += kAttnLK) {
        const int nk = min(kAttnLK, L - kc0);
        __syncthreads();  // previous chunk fully consumed
        {
            int l0 = kc0 + lrow, iph = l0 / npw, ipw = l0 - iph * npw;
            for (int r = lrow; r < nk; r += rpp) {
                if (lact) {
                    const __half * src = qkv + row_token(iph, ipw) * ld + (heads + head) * DP + lchunk * 8;
                    cp_async16(sK + r * LDS + lchunk * 8, src);
                    cp_async16(sV + r * LDS + lchunk * 8, src + heads * DP);
                }
                iph += step_h; ipw += step_w;
                if (ipw >= npw) { ipw -= npw; iph++; }
            }
        }
        cp_async_wait_all();
        __syncthreads();
        // per-lane ldmatrix base addresses in the shared window (one cvta per chunk instead of one per ldmatrix)
        const uint32_t kaddr0 = (uint32_t)__cvta_generic_to_shared(sK + (((lane >> 4) * 8 + (lane & 7)) * LDS + ((lane >> 3) & 1) * 8));
        const uint32_t vaddr0 = (uint32_t)__cvta_generic_to_shared(sV + ((((lane >> 3) & 1) * 8 + (lane & 7)) * LDS + (lane >> 4) * 8));
        for (int qb = 0; qb < qblocks; qb++) {
            const int qoff = qb * kAttnQB;
            if (qoff + warp * 16 >= nqall) break;  // this warp has no query rows in this block (nor in later ones)
            if (kc0 == 0 || qblocks > 1) {
#pragma unroll
                for (int ks = 0; ks < KS; ks++) ldmatrix_x4(qf[ks], sQ + (qoff + warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);
            }
            for (int kb = 0; kb < nk; kb += 64) {
                const uint32_t kaddr = kaddr0 + (uint32_t)(kb * LDS * 2), vaddr = vaddr0 + (uint32_t)(kb * LDS * 2);
                if (nk - kb >= 64) attn_block<DP, 8>(qf, o, m, l, kaddr, vaddr, sl2, 8);
                else attn_block<DP, 0>(qf, o, m, l, kaddr, vaddr, sl2, (nk - kb) >> 3);
            }
            if (qblocks > 1) {  // the whole sequence was in this chunk: this block is complete
                finish(qbase + qoff);
                reset();
            }
        }
    }
    if (qblocks == 1 && warp * 16 < nqall) finish(qbase);
}

// CUDA-core fallback for sequence lengths that are not a multiple of 16 (tiny feature maps); same qkv layout.
template <int DMAX>
__global__ void __launch_bounds__(128) k_attention_v1(const __half * __restrict__ qkv, int H, int W, int C, int heads, int d, int dp,
                                                      __half * __restrict__ out, int kch) {
    extern __shared__ float smem_f[];
    const int npw = W / 2, nph = H / 2, L = npw * nph;
    const int head = blockIdx.x % heads;
    const int pp   = (blockIdx.x / heads) % 4;
    const int n    = blockIdx.x / (heads * 4);
    const int ph = pp >> 1, pw = pp & 1;
    float * sK = smem_f;
    float * sV = smem_f + (size_t)kch * d;
    const float   scale = rsqrtf((float)d);
    const int64_t ld    = 3 * (int64_t)heads * dp;
    auto tok = [&](int l) -> int64_t {
        const int iph = l / npw, ipw = l % npw;
        return ((int64_t)n * H + (iph * 2 + ph)) * W + (ipw * 2 + pw);
    };
    for (int q0 = 0; q0 < L; q0 += blockDim.x) {
        const int  qi     = q0 + threadIdx.x;
        const bool active = qi < L;
        float q[DMAX], o[DMAX];
        float m = -INFINITY, l = 0.f;
        if (active) {
            const __half * qp = qkv + tok(qi) * ld + head * dp;
#pragma unroll
            for (int e = 0; e < DMAX; e++) {
                q[e] = e < d ? __half2float(qp[e]) * scale : 0.f;
                o[e] = 0.f;
            }
        }
        for (int k0 = 0; k0 < L; k0 += kch) {
            const int kn = min(kch, L - k0);
            __syncthreads();
            for (int i = threadIdx.x; i < kn * d; i += blockDim.x) {
                const int      j = i / d, e = i % d;
                const __half * kp = qkv + tok(k0 + j) * ld + (heads + head) * dp;
                sK[i] = __half2float(kp[e]);
                sV[i] = __half2float(kp[heads * dp + e]);
            }
            __syncthreads();
            if (active) {
                for (int j = 0; j < kn; j++) {
                    const float * kr = sK + j * d;
                    float         s  = 0.f;
#pragma unroll
                    for (int e = 0; e < DMAX; e++)
                        if (e < d) s = fmaf(q[e], kr[e], s);
                    const float mn = fmaxf(m, s);
                    const float c  = __expf(m - mn);
                    const float p  = __expf(s - mn);
                    l              = l * c + p;
                    m              = mn;
                    const float * vr = sV + j * d;
#pragma unroll
                    for (int e = 0; e < DMAX; e++)
                        if (e < d) o[e] = fmaf(o[e], c, p * vr[e]);
                }
            }
        }
        if (active) {
            const float inv = 1.f / l;
            __half *    op  = out + tok(qi) * (int64_t)C + head * d;
#pragma unroll
            for (int e = 0; e < DMAX; e++)
                if (e < d) op[e] = __float2half_rn(o[e] * inv);
        }
    }
}

int attention_padded_head_dim(int d) { return (d + 15) / 16 * 16; }

void launch_attention(const __half * qkv, int N, int H, int W, int C, int heads, __half * out16, cudaStream_t st) {
    const int d  = C / heads;
    const int dp = attention_padded_head_dim(d);
    const int L  = (H / 2) * (W / 2);
    if (dp > 64) B200_ABORT("attention: head dim %d > 64", d);
    if (launch_attention_tc(qkv, N, H, W, C, heads, out16, st)) return;  // tcgen05 kernel (attention_tc.cu) for sequences of k * 128 tokens
    static bool attr = false;
    if (!attr) {
        B200_CHECK(cudaFuncSetAttribute(k_attention_v1<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        B200_CHECK(cudaFuncSetAttribute(k_attention_v1<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        B200_CHECK(cudaFuncSetAttribute(k_attention_mma<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        B200_CHECK(cudaFuncSetAttribute(k_attention_mma<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        B200_CHECK(cudaFuncSetAttribute(k_attention_mma<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        B200_CHECK(cudaFuncSetAttribute(k_attention_mma<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr = true;
    }
    if (L % 16 == 0 && getenv("GGML_B200_ATTN_V1") == nullptr) {
        const int    warps = L >= kAttnQB ? 8 : L / 16;
        const int    lk    = L < kAttnLK ? L : kAttnLK;
        const int    nqb   = (L + kAttnQB - 1) / kAttnQB;
        int          qblocks = (L <= kAttnLK && nqb > 1 && getenv("GGML_B200_ATTN_QB1") == nullptr) ? nqb : 1;
        // all query blocks of a sequence share one staged K/V only while the CTA still fits the 100 KiB it may use (two per SM):
        // head dim 64 at L = 256 would need 108 KiB
        if ((size_t)(kAttnQB * qblocks + 2 * lk) * (dp + 8) * sizeof(__half) > 100 * 1024) qblocks = 1;
        const size_t smem  = (size_t)(kAttnQB * qblocks + 2 * lk) * (dp + 8) * sizeof(__half);
        dim3         grid(N * 4 * heads, nqb / qblocks);
        const float  sl2 = 1.4426950408889634f / sqrtf((float)d);  // log2(e) / sqrt(d)
        switch (dp) {
            case 16: launch_pdl(k_attention_mma<16>, grid, dim3(warps * 32), smem, st, qkv, H, W, C, heads, d, out16, sl2, lk, qblocks); break;
            case 32: launch_pdl(k_attention_mma<32>, grid, dim3(warps * 32), smem, st, qkv, H, W, C, heads, d, out16, sl2, lk, qblocks); break;
            case 48: launch_pdl(k_attention_mma<48>, grid, dim3(warps * 32), smem, st, qkv, H, W, C, heads, d, out16, sl2, lk, qblocks); break;
            default: launch_pdl(k_attention_mma<64>, grid, dim3(warps * 32), smem, st, qkv, H, W, C, heads, d, out16, sl2, lk, qblocks); break;
        }
        return;
    }
    const int    kch  = L < 256 ? L : 256;
    const size_t smem = (size_t)2 * kch * d * sizeof(float);
    const int    grid = N * 4 * heads;
    if (d <= 32) k_attention_v1<32><<<grid, 128, smem, st>>>(qkv, H, W, C, heads, d, dp, out16, kch);
    else k_attention_v1<64><<<grid, 128, smem, st>>>(qkv, H, W, C, heads, d, dp, out16, kch);
}

// ---------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------
__global__ void k_add(const float4 * __restrict__ a, const float4 * __restrict__ b, int64_t n4, float4 * __restrict__ o32,
                      __half2 * __restrict__ o16) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 x = a[i], y = b[i];
        const float4 z = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
        if (o32) o32[i] = z;
        if (o16) {
            o16[2 * i]     = __floats2half2_rn(z.x, z.y);
            o16[2 * i + 1] = __floats2half2_rn(z.z, z.w);
        }
    }
}
void launch_add(const float * a, const float * b, int64_t n, float * out32, __half * out16, cudaStream_t st) {
    if (n % 4) B200_ABORT("add: n %% 4 != 0");
    const int64_t n4 = n / 4;
    int64_t       nb = (n4 + 255) / 256;
    const int64_t cap = (int64_t)runtime().sm_count * 16;
    k_add<<<(int)(nb > cap ? cap : nb), 256, 0, st>>>((const float4 *)a, (const float4 *)b, n4, (float4 *)out32, (__half2 *)out16);
}

// NHWC -> [W,H,C,N] f32 via a 32x32 smem transpose of (pixel, channel) per image
__global__ void k_nhwc_to_nchw(const __half * __restrict__ x16, const float * __restrict__ x32, int HW, int C, float * __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    __shared__ float tile[32][33];
    const int     n  = blockIdx.z;
    const int     p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int64_t base = (int64_t)n * HW * C;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int p = p0 + r, c = c0 + threadIdx.x;
        float     v = 0.f;
        if (p < HW && c < C) v = x32 ? x32[base + (int64_t)p * C + c] : __half2float(x16[base + (int64_t)p * C + c]);
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, p = p0 + threadIdx.x;
        if (p < HW && c < C) out[base + (int64_t)c * HW + p] = tile[threadIdx.x][r];
    }
}
void launch_nhwc_to_nchw(const __half * x16, const float * x32, int N, int H, int W, int C, float * out, cudaStream_t st) {
    const int HW = H * W;
    dim3      grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
    launch_pdl(k_nhwc_to_nchw, grid, block, 0, st, x16, x32, HW, C, out);
}

__global__ void k_pool_mean(const __half * __restrict__ x16, const float * __restrict__ x32, int HW, int C, float * __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    const int n = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int64_t base = (int64_t)n * HW * C + c;
    float         s    = 0.f;
    for (int p = 0; p < HW; p++) s += x32 ? x32[base + (int64_t)p * C] : __half2float(x16[base + (int64_t)p * C]);
    out[(int64_t)n * C + c] = s / (float)HW;
}
void launch_pool_mean(const __half * x16, const float * x32, int N, int HW, int C, float * out, cudaStream_t st) {
    dim3 grid((C + 127) / 128, N);
    launch_pdl(k_pool_mean, grid, dim3(128), 0, st, x16, x32, HW, C, out);
}

// small device-to-device copy as a kernel: between two CUDA-graph launches a copy-engine memcpy costs two engine hand-overs
// (~30 us per GRU step), a kernel does not
__global__ void k_copy_words(const uint32_t * __restrict__ src, uint32_t * __restrict__ dst, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
void launch_copy_words(const void * src, void * dst, int64_t n_words, cudaStream_t st) {
    int64_t blocks = (n_words + 255) / 256;
    if (blocks > 1184) blocks = 1184;
    k_copy_words<<<(unsigned)blocks, 256, 0, st>>>((const uint32_t *)src, (uint32_t *)dst, n_words);
}

// ---- u8 image preprocessing -------------------------------------------------------------------------------
// One thread per output pixel, all 3 channels.  The arithmetic is the reference's, operation for operation, with explicit
// round-to-nearest intrinsics so that nvcc cannot contract mul+add into fma: the result is bit-identical to the CPU code.
__global__ void k_preprocess_u8(const uint8_t * __restrict__ src, int sh, int sw, float * __restrict__ dst, int H, int W, float scale, int nx3, int ny3,
                                int64_t total, uint8_t * __restrict__ dst8) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int     x = (int)(i % W);
    const int     y = (int)((i / W) % H);
    const int64_t n = i / ((int64_t)W * H);
    float *   o  = dst + i * 3;
    uint8_t * o8 = dst8 + i * 3;
    if (x >= nx3 || y >= ny3) {
        if (dst8) o8[0] = o8[1] = o8[2] = 0;
        else o[0] = o[1] = o[2] = 0.f;
        return;
    }
    const float sx = __fsub_rn(__fmul_rn((float)x + 0.5f, scale), 0.5f);
    const float sy = __fsub_rn(__fmul_rn((float)y + 0.5f, scale), 0.5f);
    int x0 = max(0, (int)floorf(sx)), y0 = max(0, (int)floorf(sy));
    x0 = min(x0, sw - 1);
    y0 = min(y0, sh - 1);
    const int   x1 = min(x0 + 1, sw - 1), y1 = min(y0 + 1, sh - 1);
    const float dx = __fsub_rn(sx, (float)x0), dy = __fsub_rn(sy, (float)y0);
    const float wx0 = __fsub_rn(1.0f, dx), wy0 = __fsub_rn(1.0f, dy);
    const uint8_t * img = src + n * (int64_t)sh * sw * 3;
    const uint8_t *p00 = img + 3 * ((int64_t)y0 * sw + x0), *p01 = img + 3 * ((int64_t)y0 * sw + x1);
    const uint8_t *p10 = img + 3 * ((int64_t)y1 * sw + x0), *p11 = img + 3 * ((int64_t)y1 * sw + x1);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float v0 = __fadd_rn(__fmul_rn((float)p00[c], wx0), __fmul_rn((float)p01[c], dx));
        const float v1 = __fadd_rn(__fmul_rn((float)p10[c], wx0), __fmul_rn((float)p11[c], dx));
        const float v  = __fadd_rn(__fmul_rn(v0, wy0), __fmul_rn(v1, dy));
        const float q  = fminf(fmaxf(roundf(v), 0.0f), 255.0f);
        if (dst8) o8[c] = (uint8_t)q;
        else o[c] = __fdiv_rn((float)(uint8_t)q, 255.0f);
    }
}
void launch_preprocess_u8(const uint8_t * src, int n, int sh, int sw, float * dst, int H, int W, cudaStream_t st, uint8_t * dst8) {
    // scale as main.cpp:550 for the square target; the general form keeps the whole image inside an H x W target
    const float scale = H == W ? (float)(sw > sh ? sw : sh) * 1.0f / (float)W : fmaxf((float)sw / (float)W, (float)sh / (float)H);
    int nx3 = (int)((float)sw / scale + 0.5f), ny3 = (int)((float)sh / scale + 0.5f);
    if (nx3 > W) nx3 = W;
    if (ny3 > H) ny3 = H;
    const int64_t total = (int64_t)n * H * W;
    k_preprocess_u8<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, sh, sw, dst, H, W, scale, nx3, ny3, total, dst8);
}

// Classifier head: logits[n][o] = sum_c pooled[n][c] * W[c][o] + bias[o], f32 FFMA.  W is the file's
// (in, out) kernel as loaded (ggml ne = (out, in): `out` fastest), so consecutive threads read consecutive weights;
// 8 pooled rows are staged in shared memory and share each weight load.
static constexpr int kHeadRows = 8;
__global__ void k_head_linear(const float * __restrict__ pooled, const float * __restrict__ W, const float * __restrict__ bias, float * __restrict__ out,
                              int N, int C, int OUT) {
    extern __shared__ float s_rows[];  // [kHeadRows][C]
    const int n0 = blockIdx.y * kHeadRows;
    for (int i = threadIdx.x; i < kHeadRows * C; i += blockDim.x) {
        const int r = i / C;
        s_rows[i]   = n0 + r < N ? pooled[(int64_t)(n0 + r) * C + (i - r * C)] : 0.f;
    }
    __syncthreads();
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= OUT) return;
    float acc[kHeadRows];
#pragma unroll
    for (int r = 0; r < kHeadRows; r++) acc[r] = 0.f;
    for (int c = 0; c < C; c++) {
        const float w = W[(int64_t)c * OUT + o];
#pragma unroll
        for (int r = 0; r < kHeadRows; r++) acc[r] = fmaf(s_rows[r * C + c], w, acc[r]);
    }
    const float b = bias[o];
#pragma unroll
    for (int r = 0; r < kHeadRows; r++)
        if (n0 + r < N) out[(int64_t)(n0 + r) * OUT + o] = acc[r] + b;
}
void launch_head_linear(const float * pooled, const float * W, const float * bias, int N, int C, int OUT, float * out, cudaStream_t st) {
    dim3 grid((OUT + 127) / 128, (N + kHeadRows - 1) / kHeadRows);
    k_head_linear<<<grid, 128, (size_t)kHeadRows * C * sizeof(float), st>>>(pooled, W, bias, out, N, C, OUT);
}

}  // namespace b200
