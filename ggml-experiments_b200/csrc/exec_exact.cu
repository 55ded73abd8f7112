// exec_exact.cu -- EXACT mode: one simple, f32-accurate CUDA kernel per ggml node.
//
// This is the "TF32/f32 validation mode" of BASELINE.json's north_star and the fall-back plan for graphs
// the fused planner does not recognise.  Layouts and rounding points are ggml's own:
//   * tensors keep ggml's ne/nb (W fastest, i.e. NCHW for activations), f32 storage;
//   * ggml_conv_2d / ggml_conv_depthwise_2d round BOTH operands to f16 and accumulate in f32
//     ([ggml-upstream] im2col(F16) + mul_mat with vec_dot_f16);
//   * ggml_mul_mat with an F32 `a` is plain f32 (FFMA), with an F16 `a` rounds `b` to f16 first;
//   * ggml_norm / ggml_soft_max accumulate their sums in double, like ggml's ggml_float.
// Kernels here are deliberately simple (CUDA cores, 64x64 smem tiles); the fast path lives in fuse.cpp.
#include <cstring>
#include <memory>

#include "gemm_tcgen05.h"
#include "internal.h"

namespace b200 {

struct V4 {  // POD copy of TView for kernel arguments
    char *  p;
    int64_t ne[4];
    int64_t nb[4];
    int     type;
};
static V4 v4(const TView & t) {
    V4 v;
    v.p = (char *)t.p;
    for (int i = 0; i < 4; i++) { v.ne[i] = t.ne[i]; v.nb[i] = t.nb[i]; }
    v.type = t.type;
    return v;
}

__device__ __forceinline__ float load_as_f32(const char * p, int type) {
    if (type == GGML_TYPE_F32) return *(const float *)p;
    if (type == GGML_TYPE_F16) return __half2float(*(const __half *)p);
    return (float)*(const int32_t *)p;
}
__device__ __forceinline__ float round_f16(float x) { return __half2float(__float2half_rn(x)); }

// ---- elementwise -------------------------------------------------------------------------------------
enum { BIN_ADD = 0, BIN_SUB, BIN_MUL, BIN_DIV };
template <int OP>
__global__ void k_binary(V4 a, V4 b, float * __restrict__ dst, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t i0 = i % a.ne[0], r = i / a.ne[0];
        int64_t i1 = r % a.ne[1];
        r /= a.ne[1];
        int64_t i2 = r % a.ne[2], i3 = r / a.ne[2];
        float x = *(const float *)(a.p + i0 * a.nb[0] + i1 * a.nb[1] + i2 * a.nb[2] + i3 * a.nb[3]);
        float y = *(const float *)(b.p + (i0 % b.ne[0]) * b.nb[0] + (i1 % b.ne[1]) * b.nb[1] + (i2 % b.ne[2]) * b.nb[2] +
                                   (i3 % b.ne[3]) * b.nb[3]);
        float z;
        if (OP == BIN_ADD) z = x + y;
        else if (OP == BIN_SUB) z = x - y;
        else if (OP == BIN_MUL) z = x * y;
        else z = x / y;
        dst[i] = z;
    }
}

enum { UN_SQRT = 0, UN_SILU, UN_TANH };
template <int OP>
__global__ void k_unary(V4 a, float * __restrict__ dst, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t i0 = i % a.ne[0], r = i / a.ne[0];
        int64_t i1 = r % a.ne[1];
        r /= a.ne[1];
        int64_t i2 = r % a.ne[2], i3 = r / a.ne[2];
        float x = *(const float *)(a.p + i0 * a.nb[0] + i1 * a.nb[1] + i2 * a.nb[2] + i3 * a.nb[3]);
        float z;
        if (OP == UN_SQRT) z = sqrtf(x);
        else if (OP == UN_SILU) z = x / (1.0f + expf(-x));  // [ggml] ggml_silu_f32
        else z = tanhf(x);
        dst[i] = z;
    }
}

// strided gather copy: dst contiguous in the logical order of `a` (ggml_cont / ggml_cont_4d), any 2/4-byte type
template <typename T>
__global__ void k_cont(V4 a, T * __restrict__ dst, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t i0 = i % a.ne[0], r = i / a.ne[0];
        int64_t i1 = r % a.ne[1];
        r /= a.ne[1];
        int64_t i2 = r % a.ne[2], i3 = r / a.ne[2];
        dst[i] = *(const T *)(a.p + i0 * a.nb[0] + i1 * a.nb[1] + i2 * a.nb[2] + i3 * a.nb[3]);
    }
}

// ggml_repeat: dst has shape d_ne, a is tiled
template <typename T>
__global__ void k_repeat(V4 a, T * __restrict__ dst, int64_t d0, int64_t d1, int64_t d2, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t i0 = i % d0, r = i / d0;
        int64_t i1 = r % d1;
        r /= d1;
        int64_t i2 = r % d2, i3 = r / d2;
        dst[i] = *(const T *)(a.p + (i0 % a.ne[0]) * a.nb[0] + (i1 % a.ne[1]) * a.nb[1] + (i2 % a.ne[2]) * a.nb[2] +
                              (i3 % a.ne[3]) * a.nb[3]);
    }
}

// ggml_concat (2-arg form): along dim 2
__global__ void k_concat2(V4 a, V4 b, float * __restrict__ dst, int64_t n) {
    const int64_t d2 = a.ne[2] + b.ne[2];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t i0 = i % a.ne[0], r = i / a.ne[0];
        int64_t i1 = r % a.ne[1];
        r /= a.ne[1];
        int64_t i2 = r % d2, i3 = r / d2;
        float v;
        if (i2 < a.ne[2]) v = *(const float *)(a.p + i0 * a.nb[0] + i1 * a.nb[1] + i2 * a.nb[2] + i3 * a.nb[3]);
        else v = *(const float *)(b.p + i0 * b.nb[0] + i1 * b.nb[1] + (i2 - a.ne[2]) * b.nb[2] + i3 * b.nb[3]);
        dst[i] = v;
    }
}

// ggml_get_rows: dst[i0, r] = a[i0, idx[r]] (as f32)
__global__ void k_get_rows(V4 a, V4 idx, float * __restrict__ dst, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t i0 = i % a.ne[0], r = i / a.ne[0];
        int32_t row = *(const int32_t *)(idx.p + r * idx.nb[0]);
        dst[i] = load_as_f32(a.p + i0 * a.nb[0] + (int64_t)row * a.nb[1], a.type);
    }
}

// mean over ne0 x ne1 -> [1,1,C,N]
__global__ void k_pool_mean_hw(V4 a, float * __restrict__ dst) {
    const int64_t c = blockIdx.x, n = blockIdx.y;
    const int64_t P = a.ne[0] * a.ne[1];
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < P; i += blockDim.x) {
        int64_t i0 = i % a.ne[0], i1 = i / a.ne[0];
        s += *(const float *)(a.p + i0 * a.nb[0] + i1 * a.nb[1] + c * a.nb[2] + n * a.nb[3]);
    }
    __shared__ float red[32];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (blockDim.x + 31) / 32; w++) t += red[w];
        dst[c + a.ne[2] * n] = t / (float)P;
    }
}

// ---- row reductions: one warp per row over ne0 ------------------------------------------------------------
__device__ __forceinline__ double warp_sum_d(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ const char * row_ptr(const V4 & a, int64_t row) {
    int64_t i1 = row % a.ne[1], r = row / a.ne[1];
    int64_t i2 = r % a.ne[2], i3 = r / a.ne[2];
    return a.p + i1 * a.nb[1] + i2 * a.nb[2] + i3 * a.nb[3];
}

// [ggml] ggml_compute_forward_norm_f32: sum (double) -> mean (float); sum of squares of (x-mean) (double)
// -> variance (float); y = (x-mean) * (1/sqrtf(var+eps))
__global__ void k_norm(V4 a, float * __restrict__ dst, float eps, int64_t rows) {
    const int lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const char * src = row_ptr(a, row);
    const int    n   = (int)a.ne[0];
    double       s   = 0.0;
    for (int i = lane; i < n; i += 32) s += (double)*(const float *)(src + (int64_t)i * a.nb[0]);
    const float mean = (float)(warp_sum_d(s) / n);
    double      s2   = 0.0;
    for (int i = lane; i < n; i += 32) {
        float v = *(const float *)(src + (int64_t)i * a.nb[0]) - mean;
        s2 += (double)(v * v);
    }
    const float variance = (float)(warp_sum_d(s2) / n);
    const float scale    = 1.0f / sqrtf(variance + eps);
    for (int i = lane; i < n; i += 32) dst[row * n + i] = (*(const float *)(src + (int64_t)i * a.nb[0]) - mean) * scale;
}

// [ggml] ggml_compute_forward_soft_max_f32: max, expf(x-max), sum in double, scale by 1/sum
__global__ void k_soft_max(V4 a, float * __restrict__ dst, int64_t rows) {
    const int lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const char * src = row_ptr(a, row);
    const int    n   = (int)a.ne[0];
    float        mx  = -INFINITY;
    for (int i = lane; i < n; i += 32) mx = fmaxf(mx, *(const float *)(src + (int64_t)i * a.nb[0]));
    mx       = warp_max_f(mx);
    double s = 0.0;
    for (int i = lane; i < n; i += 32) {
        float e = expf(*(const float *)(src + (int64_t)i * a.nb[0]) - mx);
        dst[row * n + i] = e;
        s += (double)e;
    }
    const float inv = (float)(1.0 / warp_sum_d(s));
    for (int i = lane; i < n; i += 32) dst[row * n + i] *= inv;
}

// f32 -> f16 (contiguous), operand preparation for tensor-core lowered mul_mat
__global__ void k_cast_f16(const float4 * __restrict__ src, uint2 * __restrict__ dst, int64_t n4) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = src[i];
        __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t *>(&a);
        o.y = *reinterpret_cast<uint32_t *>(&b);
        dst[i] = o;
    }
}

// Fused GRU gates (rnn.cpp:231-250, Keras reset_after=True): from mx = W^T x + b0 and mh = U^T h + b1 (both [3U, B], gate
// order z | r | h) and the previous state h [U, B]:  z = s(mx_z + mh_z), r = s(mx_r + mh_r), hh = tanh(mx_h + r * mh_h),
// h' = z * h + (1 - z) * hh.  s(x) is a true sigmoid here: the reference writes silu(x) / x (rnn.cpp:51-55), which is the same
// value to ~1 ulp but NaN at x == 0 -- with 4096 streams x 200 steps that 0/0 does occur.  The per-node
// EXACT plan keeps the literal silu/div nodes.
__global__ void k_gru_gates(const float * __restrict__ mx, const float * __restrict__ mh, const float * __restrict__ h,
                            float * __restrict__ out, int U, int64_t total) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / U;
        const int     u = (int)(i - b * U);
        const float * px = mx + b * 3 * U + u;
        const float * ph = mh + b * 3 * U + u;
        const float az = px[0] + ph[0], ar = px[U] + ph[U];
        const float z  = 1.0f / (1.0f + expf(-az));
        const float r  = 1.0f / (1.0f + expf(-ar));
        const float hh = tanhf(px[2 * U] + r * ph[2 * U]);
        out[i]         = z * h[i] + (1.0f - z) * hh;
    }
}

// dst[r, :] = table[ids[r], :] in 16-byte pieces (folded embedding projection)
__global__ void k_gather_rows_f4(const float4 * __restrict__ table, const int32_t * __restrict__ ids, float4 * __restrict__ dst, int m4, int64_t total) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / m4;
        dst[i]          = table[(int64_t)ids[r] * m4 + (i - r * m4)];
    }
}

// [rows, ldp] padded GEMM result -> contiguous [rows, m]
__global__ void k_unpad_rows(const float * __restrict__ src, float * __restrict__ dst, int m, int ldp, int64_t total) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / m;
        dst[i]          = src[r * ldp + (i - r * m)];
    }
}

// ggml_argmax: one warp per row; ties resolve to the lowest index (std::max_element, rnn.cpp:76)
__global__ void k_argmax(V4 a, int32_t * __restrict__ dst, int64_t rows) {
    const int lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const char * src = a.p + row * a.nb[1];
    float best = -INFINITY;
    int   bi   = 0x7fffffff;
    for (int i = lane; i < (int)a.ne[0]; i += 32) {
        const float v = *(const float *)(src + (int64_t)i * a.nb[0]);
        if (v > best) { best = v; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int   oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) dst[row] = bi == 0x7fffffff ? 0 : bi;  // nothing compared greater (all NaN / -inf): index 0, as upstream's vec_argmax
}

// ---- generic 64x64 tiled "GEMM with functor loaders" on CUDA cores ----------------------------------------
// C(m, n, batch) = sum_k A(m, k, batch) * B(k, n, batch), f32 accumulate, k ascending.
struct MulMatArgs {
    V4      a, b;
    float * dst;
    int64_t M, N, K;
    int64_t ne2;      // dst.ne[2] (batch = i2 + ne2*i3)
    int     round_b;  // a is F16 -> round b to f16 ([ggml] vec_dot_type)
};
struct MulMatLoader {
    const MulMatArgs & g;
    const char *       pa;
    const char *       pb;
    float *            pc;
    __device__ MulMatLoader(const MulMatArgs & g_, int64_t batch) : g(g_) {
        int64_t i2 = batch % g.ne2, i3 = batch / g.ne2;
        pa = g.a.p + (i2 % g.a.ne[2]) * g.a.nb[2] + (i3 % g.a.ne[3]) * g.a.nb[3];
        pb = g.b.p + i2 * g.b.nb[2] + i3 * g.b.nb[3];
        pc = g.dst + batch * g.M * g.N;
    }
    __device__ float A(int64_t m, int64_t k) const { return load_as_f32(pa + k * g.a.nb[0] + m * g.a.nb[1], g.a.type); }
    __device__ float B(int64_t k, int64_t n) const {
        float v = *(const float *)(pb + k * g.b.nb[0] + n * g.b.nb[1]);
        return g.round_b ? round_f16(v) : v;
    }
    __device__ void store(int64_t m, int64_t n, float v) const { pc[m + g.M * n] = v; }
};

struct ConvArgs {
    V4      w, x;  // w: [KW,KH,IC,OC] (f16 or f32, rounded to f16), x: [W,H,C,N] f32
    float * dst;   // [OW,OH,OC,N]
    int64_t M, N, K;  // M = OC, N = OW*OH, K = IC*KH*KW
    int     OW, OH, KW, KH, s0, s1, p0, p1, d0, d1;
    int     round_act;  // 0 in EXACT_F32 mode
};
struct ConvLoader {
    const ConvArgs & g;
    const char *     px;
    float *          pc;
    __device__ ConvLoader(const ConvArgs & g_, int64_t batch) : g(g_) {
        px = g.x.p + batch * g.x.nb[3];
        pc = g.dst + batch * g.M * g.N;
    }
    __device__ float A(int64_t m, int64_t k) const {  // kernel[kw,kh,ic,oc], k = (ic*KH + kh)*KW + kw
        int64_t kw = k % g.KW, r = k / g.KW;
        int64_t kh = r % g.KH, ic = r / g.KH;
        return round_f16(load_as_f32(g.w.p + kw * g.w.nb[0] + kh * g.w.nb[1] + ic * g.w.nb[2] + m * g.w.nb[3], g.w.type));
    }
    __device__ float B(int64_t k, int64_t n) const {  // im2col on the fly, zero padding, rounded to f16
        int64_t kw = k % g.KW, r = k / g.KW;
        int64_t kh = r % g.KH, ic = r / g.KH;
        int64_t ox = n % g.OW, oy = n / g.OW;
        int64_t ix = ox * g.s0 + kw * g.d0 - g.p0, iy = oy * g.s1 + kh * g.d1 - g.p1;
        if (ix < 0 || ix >= g.x.ne[0] || iy < 0 || iy >= g.x.ne[1]) return 0.f;
        const float v = *(const float *)(px + ix * g.x.nb[0] + iy * g.x.nb[1] + ic * g.x.nb[2]);
        return g.round_act ? round_f16(v) : v;
    }
    __device__ void store(int64_t m, int64_t n, float v) const { pc[n + g.N * m] = v; }
};

template <typename Args, typename Loader>
__global__ void __launch_bounds__(256) k_tiled_gemm(Args g) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const Loader  L(g, blockIdx.z);
    const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
    const int     tx = threadIdx.x % 16, ty = threadIdx.x / 16;  // 16x16 threads, 4x4 outputs each
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (int64_t k0 = 0; k0 < g.K; k0 += BK) {
        for (int e = threadIdx.x; e < BK * BM; e += 256) {
            int     kk = e % BK, mm = e / BK;  // k fastest: A rows are K-contiguous in both users
            int64_t m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < g.M && k < g.K) ? L.A(m, k) : 0.f;
        }
        for (int e = threadIdx.x; e < BK * BN; e += 256) {
            int     nn = e % BN, kk = e / BN;
            int64_t n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < g.N && k < g.K) ? L.B(k, n) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int64_t m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < g.M && n < g.N) L.store(m, n, acc[i][j]);
        }
}

// In the mul_mat the B loader walks n fastest but b rows are K-contiguous; swap the staging order for it.
template <>
__global__ void __launch_bounds__(256) k_tiled_gemm<MulMatArgs, MulMatLoader>(MulMatArgs g) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const MulMatLoader L(g, blockIdx.z);
    const int64_t      m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
    const int          tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (int64_t k0 = 0; k0 < g.K; k0 += BK) {
        for (int e = threadIdx.x; e < BK * BM; e += 256) {
            int     kk = e % BK, mm = e / BK;
            int64_t m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < g.M && k < g.K) ? L.A(m, k) : 0.f;
        }
        for (int e = threadIdx.x; e < BK * BN; e += 256) {
            int     kk = e % BK, nn = e / BK;
            int64_t n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < g.N && k < g.K) ? L.B(k, n) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[kk][tx * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[kk][ty * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int64_t m = m0 + tx * 4 + i, n = n0 + ty * 4 + j;  // m fastest across threads: coalesced dst rows
            if (m < g.M && n < g.N) L.store(m, n, acc[i][j]);
        }
}

// depthwise: one thread per output element; kernel [KW,KH,1,C]
struct DwArgs {
    V4      w, x;
    float * dst;
    int     OW, OH, KW, KH, s0, s1, p0, p1, d0, d1;
    int     round_act;
    int64_t C, Nb, total;
};
__global__ void k_dwconv(DwArgs g) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < g.total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t ox = i % g.OW, r = i / g.OW;
        int64_t oy = r % g.OH;
        r /= g.OH;
        int64_t c = r % g.C, n = r / g.C;
        float   s = 0.f;
        for (int kh = 0; kh < g.KH; kh++) {
            int64_t iy = oy * g.s1 + kh * g.d1 - g.p1;
            if (iy < 0 || iy >= g.x.ne[1]) continue;
            for (int kw = 0; kw < g.KW; kw++) {
                int64_t ix = ox * g.s0 + kw * g.d0 - g.p0;
                if (ix < 0 || ix >= g.x.ne[0]) continue;
                float wv = round_f16(load_as_f32(g.w.p + kw * g.w.nb[0] + kh * g.w.nb[1] + c * g.w.nb[3], g.w.type));
                float xv = *(const float *)(g.x.p + ix * g.x.nb[0] + iy * g.x.nb[1] + c * g.x.nb[2] + n * g.x.nb[3]);
                if (g.round_act) xv = round_f16(xv);
                s        = fmaf(wv, xv, s);
            }
        }
        g.dst[i] = s;
    }
}

// ---- plan builder -----------------------------------------------------------------------------------
static int grid_for(int64_t n, int threads = 256) {
    int64_t b = (n + threads - 1) / threads;
    int64_t cap = (int64_t)runtime().sm_count * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// mul_mat of a constant 2-d f32 weight leaf with a contiguous 2-d f32 activation, when the caller asked for FAST mode
static bool tensor_core_mul_mat(Plan * plan, const ggml_tensor * t) {
    if (t->op != GGML_OP_MUL_MAT || runtime().mode != GGML_B200_MODE_FAST || getenv("GGML_B200_NO_TC_MULMAT")) return false;
    const ggml_tensor * w = t->src[0];
    const ggml_tensor * x = t->src[1];
    if (w->op != GGML_OP_NONE || w->view_src || w->type != GGML_TYPE_F32 || !ggml_is_contiguous(w) || w->ne[2] != 1 || w->ne[3] != 1) return false;
    auto it = plan->slots.find(w);
    if (it == plan->slots.end() || it->second.kind != SLOT_CONST) return false;
    if (x->type != GGML_TYPE_F32 || !ggml_is_contiguous(x) || x->ne[2] != 1 || x->ne[3] != 1) return false;
    const int64_t K = w->ne[0], M = w->ne[1], N = x->ne[1];
    (void)M;  // M that is not a multiple of 8 (the GRU's 66-way dense layer) goes through a zero-padded weight and a padded scratch
    return K % 8 == 0 && N >= 64 && (K * N) % 4 == 0;
}

// ---- FAST-mode peephole fusions for graphs outside the MobileViT matcher (the batched GRU cell) ----------------------
struct GruGates { const ggml_tensor *mx, *mh, *h; int U; };

static const ggml_tensor * gate_view(const ggml_tensor * v, int k, int U, const ggml_tensor ** base) {
    // VIEW(base [3U,B]; ne0 = U, ne1 = B, nb1 = base->nb[1], offset k*U floats)
    if (!v || v->op != GGML_OP_VIEW || v->ne[0] != U || v->nb[1] != v->src[0]->nb[1]) return nullptr;
    const ggml_tensor * b = v->src[0];
    if (b->type != GGML_TYPE_F32 || !ggml_is_contiguous(b) || b->ne[0] != 3 * U || b->ne[1] != v->ne[1]) return nullptr;
    if (v->view_offs != b->view_offs + (size_t)k * U * sizeof(float)) return nullptr;
    if (*base && *base != b) return nullptr;
    *base = b;
    return b;
}
// s(a) = DIV(SILU(a), a) with a = ADD(view(mx,k), view(mh,k))
static bool match_gate_sigmoid(const ggml_tensor * t, int k, int U, const ggml_tensor ** mx, const ggml_tensor ** mh, std::vector<const ggml_tensor *> & inner) {
    if (t->op != GGML_OP_DIV || t->src[0]->op != GGML_OP_SILU || t->src[0]->src[0] != t->src[1]) return false;
    const ggml_tensor * a = t->src[1];
    if (a->op != GGML_OP_ADD || !gate_view(a->src[0], k, U, mx) || !gate_view(a->src[1], k, U, mh)) return false;
    inner.push_back(t->src[0]);
    inner.push_back(a);
    return true;
}
static bool match_gru_gates(const ggml_tensor * t, GruGates & g, std::vector<const ggml_tensor *> & inner) {
    // t = ADD(MUL(z, h), MUL(SUB(REPEAT(1), z), hh))
    if (t->op != GGML_OP_ADD || t->type != GGML_TYPE_F32 || t->src[0]->op != GGML_OP_MUL || t->src[1]->op != GGML_OP_MUL) return false;
    const ggml_tensor * zh = t->src[0], * rest = t->src[1];
    const ggml_tensor * z = zh->src[0], * h = zh->src[1];
    const int U = (int)t->ne[0];
    if (!ggml_is_contiguous(h) || h->type != GGML_TYPE_F32 || h->ne[0] != U || h->ne[1] != t->ne[1]) return false;
    const ggml_tensor * omz = rest->src[0], * hh = rest->src[1];
    if (omz->op != GGML_OP_SUB || omz->src[1] != z || omz->src[0]->op != GGML_OP_REPEAT) return false;
    const ggml_tensor * one = omz->src[0]->src[0];
    if (one->op != GGML_OP_NONE || ggml_nelements(one) != 1 || !one->data || *(const float *)one->data != 1.0f) return false;
    const ggml_tensor *mx = nullptr, *mh = nullptr;
    if (!match_gate_sigmoid(z, 0, U, &mx, &mh, inner)) return false;
    if (hh->op != GGML_OP_TANH || hh->src[0]->op != GGML_OP_ADD) return false;
    const ggml_tensor * ah = hh->src[0];
    const ggml_tensor * rm = ah->src[1];
    if (!gate_view(ah->src[0], 2, U, &mx) || rm->op != GGML_OP_MUL || !gate_view(rm->src[1], 2, U, &mh)) return false;
    if (!match_gate_sigmoid(rm->src[0], 1, U, &mx, &mh, inner)) return false;
    const ggml_tensor * chain[] = {zh, rest, omz, omz->src[0], hh, ah, rm, z, rm->src[0]};
    for (const ggml_tensor * x : chain) inner.push_back(x);
    g.mx = mx; g.mh = mh; g.h = h; g.U = U;
    return true;
}

void build_exact_plan(Plan * plan, ggml_cgraph * gf) {
    const int n = gf->n_nodes;
    // FAST-mode peepholes: (1) GRU gate chain -> one kernel, (2) mul_mat + bias add -> GEMM epilogue
    std::unordered_map<const ggml_tensor *, GruGates> gru_roots;
    std::unordered_map<const ggml_tensor *, int> skip;                           // nodes computed inside a fused kernel
    std::unordered_map<const ggml_tensor *, const ggml_tensor *> bias_of;         // mul_mat node -> bias leaf folded into its GEMM
    std::unordered_map<const ggml_tensor *, const ggml_tensor *> alias_of;        // bias-add node -> the mul_mat whose buffer it shares
    std::unordered_map<const ggml_tensor *, const ggml_tensor *> table_of;        // mul_mat node -> the get_rows it absorbed (folded embedding projection)
    if (runtime().mode == GGML_B200_MODE_FAST && !getenv("GGML_B200_NO_PEEPHOLE")) {
        std::unordered_map<const ggml_tensor *, int> uses;
        for (int i = 0; i < n; i++)
            for (int s = 0; s < GGML_MAX_SRC; s++)
                if (gf->nodes[i]->src[s]) uses[gf->nodes[i]->src[s]]++;
        for (int i = 0; i < n; i++) {
            const ggml_tensor * t = gf->nodes[i];
            GruGates g;
            std::vector<const ggml_tensor *> inner;
            if (match_gru_gates(t, g, inner)) {
                bool ok = true;
                for (const ggml_tensor * x : inner)
                    if (x->flags & GGML_TENSOR_FLAG_OUTPUT) ok = false;  // an inner value is observable: keep the chain
                if (ok) {
                    gru_roots[t] = g;
                    for (const ggml_tensor * x : inner) skip[x] = 1;
                }
            }
        }
        for (int i = 0; i < n; i++) {
            const ggml_tensor * t = gf->nodes[i];  // ADD(MUL_MAT(w, x), bias [M,1]) with the product used only here
            if (t->op != GGML_OP_ADD || t->src[0]->op != GGML_OP_MUL_MAT || uses[t->src[0]] != 1) continue;
            const ggml_tensor * mm = t->src[0], * b = t->src[1];
            if (b->op != GGML_OP_NONE || b->view_src || b->type != GGML_TYPE_F32 || b->ne[0] != mm->ne[0] || ggml_nelements(b) != mm->ne[0]) continue;
            if (!plan->slots.count(b) || plan->slots[b].kind != SLOT_CONST || (mm->flags & GGML_TENSOR_FLAG_OUTPUT)) continue;
            if (!tensor_core_mul_mat(plan, mm)) continue;
            bias_of[mm]  = b;
            alias_of[t]  = mm;
            // W^T E[ids] + b with a constant embedding table E (rnn.cpp:200-209): fold W^T E + b into one [M, V] table at plan
            // time (f32 on the host) and gather its rows -- the per-step GEMM disappears
            const ggml_tensor * gr = mm->src[1];
            if (gr->op == GGML_OP_GET_ROWS && uses[gr] == 1 && !(gr->flags & GGML_TENSOR_FLAG_OUTPUT) && mm->ne[0] % 4 == 0) {
                const ggml_tensor * E = gr->src[0];
                if (E->op == GGML_OP_NONE && !E->view_src && E->type == GGML_TYPE_F32 && ggml_is_contiguous(E) && E->data && E->ne[2] == 1 && E->ne[3] == 1 &&
                    E->ne[1] <= 65536 && plan->slots.count(E) && plan->slots[E].kind == SLOT_CONST && gr->src[1]->type == GGML_TYPE_I32 &&
                    ggml_is_contiguous(gr->src[1])) {
                    table_of[mm] = gr;
                    skip[gr]     = 1;
                }
            }
        }
    }
    // ---- liveness: last node index that reads each buffer-owning tensor (through any chain of views) ----
    auto base_of = [](const ggml_tensor * t) -> const ggml_tensor * {
        while (t && is_view_op(t->op)) t = t->src[0];
        return t;
    };
    std::unordered_map<const ggml_tensor *, int> last_use;
    for (int i = 0; i < n; i++) {
        ggml_tensor * t = gf->nodes[i];
        for (int s = 0; s < GGML_MAX_SRC; s++)
            if (t->src[s]) last_use[base_of(t->src[s])] = i;
    }
    for (int i = 0; i < n; i++)
        if (gf->nodes[i]->flags & GGML_TENSOR_FLAG_OUTPUT) last_use[base_of(gf->nodes[i])] = n;  // outputs live forever
    if (getenv("GGML_B200_DUMP_NODES") != nullptr)  // node-by-node comparison with the CPU reference run: nothing is ever overwritten
        for (int i = 0; i < n; i++) last_use[base_of(gf->nodes[i])] = n;
    // fused kernels read their operands at the ROOT node and a folded bias-add lives in its mul_mat's buffer
    for (int i = 0; i < n; i++) {
        const ggml_tensor * t = gf->nodes[i];
        auto extend = [&](const ggml_tensor * x, int until) {
            const ggml_tensor * b = base_of(x);
            if (!last_use.count(b) || last_use[b] < until) last_use[b] = until;
        };
        if (gru_roots.count(t)) {
            extend(gru_roots[t].mx, i);
            extend(gru_roots[t].mh, i);
            extend(gru_roots[t].h, i);
        }
        if (alias_of.count(t)) extend(alias_of[t], last_use.count(t) ? last_use[t] : i);
        if (table_of.count(t)) extend(table_of[t]->src[1], i);
    }
    // ---- assign arena offsets in execution order ----
    ArenaPlanner ap;
    std::unordered_map<const ggml_tensor *, int64_t> scratch_off;
    std::vector<std::vector<const ggml_tensor *>> dies_at(n + 1);
    for (int i = 0; i < n; i++) {
        ggml_tensor * t = gf->nodes[i];
        if (!is_view_op(t->op)) {
            Slot s;
            s.kind   = SLOT_ARENA;
            s.bytes  = (int64_t)ggml_nelements(t) * (int64_t)ggml_type_size(t->type);
            s.offset = ap.alloc(s.bytes);
            plan->naive_bytes += ArenaPlanner::align_up(s.bytes);
            int lu = last_use.count(t) ? last_use[t] : i;
            s.first_use = i;
            s.last_use  = lu;
            plan->slots[t] = s;
            if (lu < n) dies_at[lu < i ? i : lu].push_back(t);  // a never-read result is freed right after its node
            if (tensor_core_mul_mat(plan, t)) {  // scratch for the f16 copy of the activation operand, live for this node only
                int64_t sb = (int64_t)ggml_nelements(t->src[1]) * 2;
                if (t->ne[0] % 8) sb = ArenaPlanner::align_up(sb) + ((t->ne[0] + 7) / 8 * 8) * t->ne[1] * 4;  // + padded f32 result
                const int64_t so = ap.alloc(sb);
                scratch_off[t] = so;
                ap.release(so, sb);
            }
        }
        // buffers whose last reader is node i are released only after node i's own output was placed,
        // so an op never writes over one of its inputs
        for (const ggml_tensor * d : dies_at[i]) {
            const Slot & ds = plan->slots[d];
            ap.release(ds.offset, ds.bytes);
        }
    }
    plan->arena_bytes = ap.extent;
    if (plan->arena_bytes > 0) {
        B200_CHECK(cudaMalloc((void **)&plan->arena, plan->arena_bytes));
        plan->owned_device.push_back(plan->arena);
    }
    for (int i = 0; i < n; i++) {
        ggml_tensor * t = gf->nodes[i];
        if (is_view_op(t->op)) {
            Slot s;
            s.kind = SLOT_ALIAS;
            // views alias their source's storage; ggml keeps view_offs relative to view_src (the base)
            const ggml_tensor * base = t->view_src ? t->view_src : t->src[0];
            s.dptr = (char *)device_ptr_of(plan, base) + t->view_offs;
            plan->slots[t] = s;
        } else {
            plan->slots[t].dptr = plan->arena + plan->slots[t].offset;
        }
        if (alias_of.count(t)) plan->slots[t].dptr = plan->slots[alias_of[t]].dptr;  // the GEMM epilogue already added the bias in place
    }
    // NOTE: a view created over a *view* records view_src = the root tensor and view_offs from the root, so the
    // alias above is correct for chains (reshape(permute(x)) etc.).

    // ---- one launch per non-view node ----
    for (int i = 0; i < n; i++) {
        ggml_tensor * t = gf->nodes[i];
        if (is_view_op(t->op)) continue;
        void *        d  = plan->slots[t].dptr;
        const int64_t ne = ggml_nelements(t);
        if (skip.count(t) || alias_of.count(t)) { plan->n_folded++; continue; }
        if (gru_roots.count(t)) {
            const GruGates & g = gru_roots[t];
            const float *mx = (const float *)device_ptr_of(plan, g.mx), *mh = (const float *)device_ptr_of(plan, g.mh), *hp = (const float *)device_ptr_of(plan, g.h);
            const int U = g.U;
            const int gr = grid_for(ne);
            add_launch(plan, "gru_gates_fused", [=](cudaStream_t st) { k_gru_gates<<<gr, 256, 0, st>>>(mx, mh, hp, (float *)d, U, ne); }, 12.0 * ne,
                       4.0 * ne * 8, t->name);
            continue;
        }
        auto src_view = [&](int s) { return v4(make_view(t->src[s], device_ptr_of(plan, t->src[s]))); };
        switch (t->op) {
            case GGML_OP_ADD: case GGML_OP_SUB: case GGML_OP_MUL: case GGML_OP_DIV: {
                V4 a = src_view(0), b = src_view(1);
                int g = grid_for(ne);
                enum ggml_op op = t->op;
                add_launch(plan, "exact_binary", [=](cudaStream_t st) {
                    if (op == GGML_OP_ADD) k_binary<BIN_ADD><<<g, 256, 0, st>>>(a, b, (float *)d, ne);
                    else if (op == GGML_OP_SUB) k_binary<BIN_SUB><<<g, 256, 0, st>>>(a, b, (float *)d, ne);
                    else if (op == GGML_OP_MUL) k_binary<BIN_MUL><<<g, 256, 0, st>>>(a, b, (float *)d, ne);
                    else k_binary<BIN_DIV><<<g, 256, 0, st>>>(a, b, (float *)d, ne);
                });
            } break;
            case GGML_OP_SQRT: case GGML_OP_SILU: case GGML_OP_TANH: {
                V4 a = src_view(0);
                int g = grid_for(ne);
                enum ggml_op op = t->op;
                add_launch(plan, "exact_unary", [=](cudaStream_t st) {
                    if (op == GGML_OP_SQRT) k_unary<UN_SQRT><<<g, 256, 0, st>>>(a, (float *)d, ne);
                    else if (op == GGML_OP_SILU) k_unary<UN_SILU><<<g, 256, 0, st>>>(a, (float *)d, ne);
                    else k_unary<UN_TANH><<<g, 256, 0, st>>>(a, (float *)d, ne);
                });
            } break;
            case GGML_OP_NORM: case GGML_OP_SOFT_MAX: {
                V4 a = src_view(0);
                GGML_ASSERT(t->src[0]->type == GGML_TYPE_F32);
                const int64_t rows = ne / t->ne[0];
                float eps;
                memcpy(&eps, t->op_params, sizeof(float));
                int g = (int)((rows + 7) / 8);
                bool is_norm = t->op == GGML_OP_NORM;
                add_launch(plan, is_norm ? "exact_norm" : "exact_soft_max", [=](cudaStream_t st) {
                    if (is_norm) k_norm<<<g, 256, 0, st>>>(a, (float *)d, eps, rows);
                    else k_soft_max<<<g, 256, 0, st>>>(a, (float *)d, rows);
                });
            } break;
            case GGML_OP_MUL_MAT: {
                if (table_of.count(t)) {
                    const ggml_tensor * wt = t->src[0], * gr = table_of[t], * E = gr->src[0], * bt = bias_of[t];
                    const int K = (int)wt->ne[0], M = (int)wt->ne[1], V = (int)E->ne[1];
                    std::vector<float> table((size_t)V * M);
                    const float *Wd = (const float *)wt->data, *Ed = (const float *)E->data, *Bd = (const float *)bt->data;
                    for (int v = 0; v < V; v++)
                        for (int m = 0; m < M; m++) {
                            double acc = 0.0;
                            for (int k = 0; k < K; k++) acc += (double)Wd[(size_t)m * K + k] * (double)Ed[(size_t)v * K + k];
                            table[(size_t)v * M + m] = (float)acc + Bd[m];
                        }
                    void * dt = nullptr;
                    B200_CHECK(cudaMalloc(&dt, table.size() * 4));
                    plan->owned_device.push_back(dt);
                    B200_CHECK(cudaMemcpy(dt, table.data(), table.size() * 4, cudaMemcpyHostToDevice));
                    plan->weight_bytes += (int64_t)table.size() * 4;
                    const int32_t * ids = (const int32_t *)device_ptr_of(plan, gr->src[1]);
                    const int       m4  = M / 4;
                    const int64_t   tot = (int64_t)m4 * t->ne[1];
                    const int       gg  = grid_for(tot);
                    add_launch(plan, "gather_folded_projection", [=](cudaStream_t st) { k_gather_rows_f4<<<gg, 256, 0, st>>>((const float4 *)dt, ids, (float4 *)d, m4, tot); },
                               0.0, 32.0 * tot, t->name);
                    break;
                }
                if (scratch_off.count(t)) {
                    // FAST-mode lowering of a dense layer in a graph the fused planner does not know (e.g. the GRU cell):
                    // ggml [K,M] weights x [K,N] activations -> [M,N] is exactly the K-major A/B layout of the tcgen05 GEMM
                    // (C[N rows, M] = act[N,K] * W[M,K]^T); operands go to f16, accumulation stays f32.
                    const ggml_tensor * wt = t->src[0];
                    const ggml_tensor * xt = t->src[1];
                    const int K = (int)wt->ne[0], M = (int)wt->ne[1], N = (int)xt->ne[1];
                    const int Mp = (M + 7) / 8 * 8;
                    std::vector<uint16_t> w16((size_t)K * Mp, 0);
                    ggml_fp32_to_fp16_row((const float *)wt->data, w16.data(), K * M);
                    void * dw = nullptr;
                    B200_CHECK(cudaMalloc(&dw, w16.size() * 2));
                    plan->owned_device.push_back(dw);
                    B200_CHECK(cudaMemcpy(dw, w16.data(), w16.size() * 2, cudaMemcpyHostToDevice));
                    plan->weight_bytes += (int64_t)w16.size() * 2;
                    __half *      x16 = (__half *)(plan->arena + scratch_off[t]);
                    const float * x32 = (const float *)device_ptr_of(plan, xt);
                    const int64_t n4  = (int64_t)K * N / 4;
                    const int     cg  = grid_for(n4);
                    add_launch(plan, "cast_f32_to_f16", [=](cudaStream_t st) { k_cast_f16<<<cg, 256, 0, st>>>((const float4 *)x32, (uint2 *)x16, n4); }, 0.0,
                               6.0 * K * N, t->name);
                    GemmEpilogue ep;
                    float * padded = Mp == M ? nullptr : (float *)(plan->arena + scratch_off[t] + ArenaPlanner::align_up((int64_t)K * N * 2));
                    ep.out32 = padded ? padded : (float *)d;
                    ep.ld32  = Mp;
                    if (bias_of.count(t)) {  // bias add folded into the epilogue
                        ep.shift = (const float *)device_ptr_of(plan, bias_of[t]);
                        if (padded) {
                            std::vector<float> bp(Mp, 0.f);
                            memcpy(bp.data(), bias_of[t]->data, (size_t)M * 4);
                            void * db = nullptr;
                            B200_CHECK(cudaMalloc(&db, (size_t)Mp * 4));
                            plan->owned_device.push_back(db);
                            B200_CHECK(cudaMemcpy(db, bp.data(), (size_t)Mp * 4, cudaMemcpyHostToDevice));
                            ep.shift = (const float *)db;
                        }
                    }
                    auto L = std::make_shared<GemmLaunch>();
                    if (!gemm_prepare(*L, x16, K, (const __half *)dw, K, N, Mp, K, ep)) B200_ABORT("tensor-core mul_mat lowering failed for %dx%dx%d", N, M, K);
                    add_launch(plan, "gemm_tcgen05_mul_mat", [L](cudaStream_t st) { gemm_launch(*L, st); }, 2.0 * N * M * K,
                               2.0 * ((double)N * K + (double)M * K) + 4.0 * N * M, t->name);
                    if (padded) {
                        const int64_t tot = (int64_t)N * M;
                        const int     ug  = grid_for(tot);
                        float *       dst = (float *)d;
                        add_launch(plan, "unpad_rows", [=](cudaStream_t st) { k_unpad_rows<<<ug, 256, 0, st>>>(padded, dst, M, Mp, tot); }, 0.0, 8.0 * tot, t->name);
                    }
                    break;
                }
                MulMatArgs g;
                g.a = src_view(0);
                g.b = src_view(1);
                g.dst = (float *)d;
                g.M = t->ne[0]; g.N = t->ne[1]; g.K = t->src[0]->ne[0];
                g.ne2 = t->ne[2];
                g.round_b = t->src[0]->type == GGML_TYPE_F16 && runtime().mode != GGML_B200_MODE_EXACT_F32;
                dim3 grid((unsigned)((g.M + 63) / 64), (unsigned)((g.N + 63) / 64), (unsigned)(t->ne[2] * t->ne[3]));
                const double nb = (double)(t->ne[2] * t->ne[3]);
                add_launch(plan, "exact_mul_mat", [=](cudaStream_t st) { k_tiled_gemm<MulMatArgs, MulMatLoader><<<grid, 256, 0, st>>>(g); },
                           2.0 * g.M * g.N * g.K * nb, 4.0 * nb * (g.M * g.N + g.N * g.K) + (double)ggml_nbytes(t->src[0]), t->name);
            } break;
            case GGML_OP_CONV_2D: {
                ConvArgs g;
                g.w = src_view(0);
                g.x = src_view(1);
                g.dst = (float *)d;
                g.OW = (int)t->ne[0]; g.OH = (int)t->ne[1];
                g.KW = (int)t->src[0]->ne[0]; g.KH = (int)t->src[0]->ne[1];
                g.M = t->ne[2]; g.N = t->ne[0] * t->ne[1]; g.K = t->src[0]->ne[2] * g.KW * g.KH;
                g.s0 = t->op_params[0]; g.s1 = t->op_params[1]; g.p0 = t->op_params[2]; g.p1 = t->op_params[3];
                g.d0 = t->op_params[4]; g.d1 = t->op_params[5];
                g.round_act = runtime().mode != GGML_B200_MODE_EXACT_F32;
                dim3 grid((unsigned)((g.M + 63) / 64), (unsigned)((g.N + 63) / 64), (unsigned)t->ne[3]);
                add_launch(plan, "exact_conv_2d", [=](cudaStream_t st) { k_tiled_gemm<ConvArgs, ConvLoader><<<grid, 256, 0, st>>>(g); },
                           2.0 * g.M * g.N * g.K * (double)t->ne[3], (double)ggml_nbytes(t) + (double)ggml_nbytes(t->src[1]) + (double)ggml_nbytes(t->src[0]));
            } break;
            case GGML_OP_CONV_DEPTHWISE_2D: {
                DwArgs g;
                g.w = src_view(0);
                g.x = src_view(1);
                g.dst = (float *)d;
                g.OW = (int)t->ne[0]; g.OH = (int)t->ne[1]; g.C = t->ne[2]; g.Nb = t->ne[3]; g.total = ne;
                g.KW = (int)t->src[0]->ne[0]; g.KH = (int)t->src[0]->ne[1];
                g.s0 = t->op_params[0]; g.s1 = t->op_params[1]; g.p0 = t->op_params[2]; g.p1 = t->op_params[3];
                g.d0 = t->op_params[4]; g.d1 = t->op_params[5];
                g.round_act = runtime().mode != GGML_B200_MODE_EXACT_F32;
                int grid = grid_for(ne);
                add_launch(plan, "exact_conv_depthwise_2d", [=](cudaStream_t st) { k_dwconv<<<grid, 256, 0, st>>>(g); },
                           2.0 * (double)ne * g.KW * g.KH, (double)ggml_nbytes(t) + (double)ggml_nbytes(t->src[1]));
            } break;
            case GGML_OP_CONT: {
                V4 a = src_view(0);
                int g = grid_for(ne);
                size_t ts = ggml_type_size(t->type);
                add_launch(plan, "exact_cont", [=](cudaStream_t st) {
                    if (ts == 4) k_cont<uint32_t><<<g, 256, 0, st>>>(a, (uint32_t *)d, ne);
                    else k_cont<uint16_t><<<g, 256, 0, st>>>(a, (uint16_t *)d, ne);
                });
            } break;
            case GGML_OP_REPEAT: {
                V4 a = src_view(0);
                int g = grid_for(ne);
                size_t ts = ggml_type_size(t->type);
                int64_t d0 = t->ne[0], d1 = t->ne[1], d2 = t->ne[2];
                add_launch(plan, "exact_repeat", [=](cudaStream_t st) {
                    if (ts == 4) k_repeat<uint32_t><<<g, 256, 0, st>>>(a, (uint32_t *)d, d0, d1, d2, ne);
                    else k_repeat<uint16_t><<<g, 256, 0, st>>>(a, (uint16_t *)d, d0, d1, d2, ne);
                });
            } break;
            case GGML_OP_CONCAT: {
                V4 a = src_view(0), b = src_view(1);
                int g = grid_for(ne);
                add_launch(plan, "exact_concat", [=](cudaStream_t st) { k_concat2<<<g, 256, 0, st>>>(a, b, (float *)d, ne); });
            } break;
            case GGML_OP_GET_ROWS: {
                V4 a = src_view(0), b = src_view(1);
                int g = grid_for(ne);
                add_launch(plan, "exact_get_rows", [=](cudaStream_t st) { k_get_rows<<<g, 256, 0, st>>>(a, b, (float *)d, ne); });
            } break;
            case GGML_OP_ARGMAX: {
                V4 a = src_view(0);
                const int64_t rows = t->ne[0];
                int g = (int)((rows + 7) / 8);
                add_launch(plan, "exact_argmax", [=](cudaStream_t st) { k_argmax<<<g, 256, 0, st>>>(a, (int32_t *)d, rows); });
            } break;
            case GGML_OP_POOL_MEAN_HW: {
                V4 a = src_view(0);
                dim3 grid((unsigned)t->ne[2], (unsigned)t->ne[3]);
                add_launch(plan, "exact_pool_mean_hw", [=](cudaStream_t st) { k_pool_mean_hw<<<grid, 64, 0, st>>>(a, (float *)d); });
            } break;
            default: B200_ABORT("exact plan: unsupported op %d", (int)t->op);
        }
    }
}

}  // namespace b200
