// debug_api.cu -- host-buffer entry points that run ONE kernel of the fast path in isolation, so the parity
// tests can check each kernel against the oracle / numpy before it is trusted inside a fused plan.
#include <vector>

#include "gemm_tcgen05.h"
#include "internal.h"

using namespace b200;

namespace {
struct DevBuf {
    void * p = nullptr;
    DevBuf(const void * host, size_t bytes) {
        if (bytes == 0) return;
        B200_CHECK(cudaMalloc(&p, bytes));
        if (host) B200_CHECK(cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice));
        else B200_CHECK(cudaMemset(p, 0, bytes));
    }
    ~DevBuf() { if (p) cudaFree(p); }
};
}  // namespace

extern "C" int ggml_b200_debug_gemm(const uint16_t * A, const uint16_t * B, int M, int N, int K, const float * scale,
                                    const float * shift, int act, const float * res32, float * out32, uint16_t * out16) {
    ensure_device();
    DevBuf dA(A, (size_t)M * K * 2), dB(B, (size_t)N * K * 2);
    DevBuf dS(scale, scale ? (size_t)N * 4 : 0), dH(shift, shift ? (size_t)N * 4 : 0);
    DevBuf dR(res32, res32 ? (size_t)M * N * 4 : 0);
    DevBuf dO32(nullptr, out32 ? (size_t)M * N * 4 : 0), dO16(nullptr, out16 ? (size_t)M * N * 2 : 0);
    GemmEpilogue ep;
    ep.scale = (const float *)dS.p; ep.shift = (const float *)dH.p; ep.act = act;
    ep.res32 = (const float *)dR.p; ep.ldr32 = N;
    ep.out32 = (float *)dO32.p; ep.ld32 = N;
    ep.out16 = (__half *)dO16.p; ep.ld16 = N;
    GemmLaunch L;
    if (!gemm_prepare(L, (const __half *)dA.p, K, (const __half *)dB.p, K, M, N, K, ep)) return 1;
    cudaStream_t st = current_stream();
    gemm_launch(L, st);
    B200_CHECK(cudaGetLastError());
    B200_CHECK(cudaStreamSynchronize(st));
    if (out32) B200_CHECK(cudaMemcpy(out32, dO32.p, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
    if (out16) B200_CHECK(cudaMemcpy(out16, dO16.p, (size_t)M * N * 2, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int ggml_b200_debug_conv3x3(const uint16_t * x0, int C0, const uint16_t * x1, int C1, int Nimg, int H, int W,
                                       const uint16_t * Wt, int OC, const float * scale, const float * shift, int act,
                                       float * out32) {
    ensure_device();
    const size_t px = (size_t)Nimg * H * W;
    DevBuf dX0(x0, px * C0 * 2), dX1(x1, C1 ? px * C1 * 2 : 0), dW(Wt, (size_t)OC * 9 * (C0 + C1) * 2);
    DevBuf dS(scale, scale ? (size_t)OC * 4 : 0), dH(shift, shift ? (size_t)OC * 4 : 0);
    DevBuf dO(nullptr, px * OC * 4);
    GemmEpilogue ep;
    ep.scale = (const float *)dS.p; ep.shift = (const float *)dH.p; ep.act = act;
    ep.out32 = (float *)dO.p; ep.ld32 = OC;
    GemmLaunch L;
    if (!conv3x3_prepare(L, (const __half *)dX0.p, C0, (const __half *)dX1.p, C1, Nimg, H, W, (const __half *)dW.p, OC, ep)) return 1;
    cudaStream_t st = current_stream();
    gemm_launch(L, st);
    B200_CHECK(cudaGetLastError());
    B200_CHECK(cudaStreamSynchronize(st));
    B200_CHECK(cudaMemcpy(out32, dO.p, px * OC * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// Timing probe for kernel tuning: device buffers only (uninitialised A/B are fine for timing), returns mean ms per launch.
extern "C" float ggml_b200_debug_gemm_time(int M, int N, int K, int act, int want16, int want32, int want_res, int reps) {
    ensure_device();
    DevBuf dA(nullptr, (size_t)M * K * 2), dB(nullptr, (size_t)N * K * 2), dS(nullptr, (size_t)N * 4), dH(nullptr, (size_t)N * 4);
    DevBuf dR(nullptr, want_res ? (size_t)M * N * 4 : 0), dO32(nullptr, want32 ? (size_t)M * N * 4 : 0), dO16(nullptr, want16 ? (size_t)M * N * 2 : 0);
    GemmEpilogue ep;
    ep.scale = (const float *)dS.p; ep.shift = (const float *)dH.p; ep.act = act;
    ep.res32 = (const float *)dR.p; ep.ldr32 = N;
    ep.out32 = (float *)dO32.p; ep.ld32 = N;
    ep.out16 = (__half *)dO16.p; ep.ld16 = N;
    GemmLaunch L;
    if (!gemm_prepare(L, (const __half *)dA.p, K, (const __half *)dB.p, K, M, N, K, ep)) return -1.f;
    cudaStream_t st = current_stream();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    gemm_launch(L, st);
    cudaEventRecord(e0, st);
    for (int i = 0; i < reps; i++) gemm_launch(L, st);
    cudaEventRecord(e1, st);
    B200_CHECK(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    fprintf(stderr, "gemm %dx%dx%d act=%d: block_n=%d n_tiles=%d stages=%d ctas/sm=%d grid=%ux%u smem=%zu  %.1f us\n", M, N, K, act, L.p.block_n,
            L.p.n_tiles, L.p.stages, L.ctas_per_sm, L.grid.x, L.grid.y, L.smem_bytes, 1e3f * ms / reps);
    return ms / reps;
}

// ---- instruction-rate probe (tests/bw_probe.py): FFMA vs FHFMA (fma.rn.f32.f16) vs HFMA2 issue rate, 8 independent chains ----
template <int MODE>
__global__ void k_fma_rate(float * out, int iters, uint32_t xa, uint32_t xb) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = (float)(threadIdx.x + j);
    float    fa = __uint_as_float(xa), fb = __uint_as_float(xb);
    uint32_t ha = xa, hb = xb;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (MODE == 0) acc[j] = fmaf(acc[j], fa, fb);
            if (MODE == 1) asm volatile("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tfma.rn.f32.f16 %0, l, h, %0;\n\t}" : "+f"(acc[j]) : "r"(ha + (uint32_t)j));
            if (MODE == 2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(reinterpret_cast<uint32_t &>(acc[j])) : "r"(ha), "r"(hb));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) s += acc[j];
    if (s == 12345.678f) out[0] = s;
}
extern "C" float ggml_b200_debug_fma_rate(int mode, int iters) {
    b200::ensure_device();
    float * d = nullptr;
    cudaMalloc(&d, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 256;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (mode == 0) k_fma_rate<0><<<grid, block>>>(d, iters, 0x3f7fff00u, 0x3a000000u);
        if (mode == 1) k_fma_rate<1><<<grid, block>>>(d, iters, 0x3c003c00u, 0u);
        if (mode == 2) k_fma_rate<2><<<grid, block>>>(d, iters, 0x3bff3bffu, 0x10001000u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(d);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    // warp-instructions per clock per SM at 1.965 GHz
    const double instr = (double)grid * (block / 32) * (double)iters * 8.0;
    return (float)(instr / (ms * 1e-3) / 148.0 / 1.965e9);
}
