// fast_kernels.h -- the memory-bound kernels of the FAST plan (NHWC f16 activations).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace b200 {

// K2: stem convolution 3x3 / stride 2 / pad 1, 3 -> OC (OC <= 32, multiple of 8) channels, + BN scale/shift + SiLU.
// Input f32 (rounded to f16 like ggml's im2col), addressed as x[n*sn + y*sy + x*sx + c*sc] (elements), so both the
// HWC batch input of this build and the CHW input of an unmodified main.cpp work.  Output f16 NHWC.
// Wt: [OC][3][3][3] f16 (oc, kh, kw, ic).
void launch_stem(const float * x, int64_t sn, int64_t sy, int64_t sx, int64_t sc, int N, int H, int W, const __half * Wt, int OC,
                 const float * scale, const float * shift, int act, __half * out16, float * out32, cudaStream_t st,
                 const uint8_t * x8 = nullptr, const int * use_x8 = nullptr);
// true if launch_stem would run the tensor-core stem, which can also stage its patch from u8 images [N][H][W][3] (x8, when *use_x8 != 0)
bool stem_takes_u8(int OC, int W, bool hwc, bool out16, bool out32);

// K3: depthwise 3x3 (stride 1 or 2, pad 1) + BN scale/shift + SiLU over NHWC f16.  Wt: [3][3][C] f16.
void launch_dwconv(const __half * x, int N, int H, int W, int C, int stride, const __half * Wt, const float * scale,
                   const float * shift, int act, __half * out16, cudaStream_t st);

// K5: LayerNorm over C (eps, gamma, beta): f32 rows in, f16 rows out (operand of the next GEMM) and/or f32 out.
void launch_layernorm(const float * x, int64_t rows, int C, const float * gamma, const float * beta, float eps, __half * out16,
                      float * out32, cudaStream_t st);

// K7: softmax(Q K^T / sqrt(d)) V for every (image, patch position, head).
// qkv: f16 [N*H*W][3][heads][DP] in pixel order, DP = attention_padded_head_dim(C/heads), zero padded; a sequence is
// the (H/2)*(W/2) pixels that share the same (y%2, x%2) -- the reference's unfold (main.cpp:721-747) is never
// materialised.  out: f16 [N*H*W, C] (unpadded).
int attention_padded_head_dim(int d);
void launch_attention(const __half * qkv, int N, int H, int W, int C, int heads, __half * out16, cudaStream_t st);
// the same on tcgen05 (attention_tc.cu): S and P.V accumulate in TMEM; returns false for shapes it does not cover (L % 128, map width)
bool launch_attention_tc(const __half * qkv, int N, int H, int W, int C, int heads, __half * out16, cudaStream_t st);

// elementwise residual add: out = a + b (f32), optional f16 copy
void launch_add(const float * a, const float * b, int64_t n, float * out32, __half * out16, cudaStream_t st);

// NHWC (f16 or f32) -> ggml layout [W,H,C,N] f32 (graph outputs, debug taps)
void launch_nhwc_to_nchw(const __half * x16, const float * x32, int N, int H, int W, int C, float * out, cudaStream_t st);

// global average pool over H*W of NHWC f32/f16 -> [N][C] f32
void launch_pool_mean(const __half * x16, const float * x32, int N, int HW, int C, float * out, cudaStream_t st);
void launch_copy_words(const void * src, void * dst, int64_t n_words, cudaStream_t st);  // n_words 4-byte words
// sam_image_preprocess on the device (main.cpp:538-601): n u8 images [sh][sw][3] -> f32 [n][H][W][3], longer side fills the target,
// bilinear with the reference's arithmetic, rounded to u8, /255, zero padding, row stride W.
// dst8 != nullptr: write the quantised u8 image (the value before the /255) instead of the f32 one
void launch_preprocess_u8(const uint8_t * src, int n, int sh, int sw, float * dst, int H, int W, cudaStream_t st, uint8_t * dst8 = nullptr);
void launch_head_linear(const float * pooled, const float * W, const float * bias, int N, int C, int OUT, float * out, cudaStream_t st);

}  // namespace b200
