// debug_api.cu -- host-buffer entry points that run ONE kernel of the fast path in isolation, so the parity
// tests can check each kernel against the oracle / numpy before it is trusted inside a fused plan.
#include <cstring>
#include <vector>

#include "fast_kernels.h"
#include "gemm_tcgen05.h"
#include "internal.h"
#include "vit_stage.cuh"

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
    std::vector<uint8_t> halo;  // the halo-mode kernel is used where the map qualifies (GGML_B200_CONV_NO_HALO=1: per-tap boxes everywhere)
    conv3x3_pack_halo(Wt, OC, C0, C1, halo);
    DevBuf dWh(halo.data(), halo.size());
    GemmEpilogue ep;
    ep.scale = (const float *)dS.p; ep.shift = (const float *)dH.p; ep.act = act;
    ep.out32 = (float *)dO.p; ep.ld32 = OC;
    GemmLaunch L;
    if (!conv3x3_prepare(L, (const __half *)dX0.p, C0, (const __half *)dX1.p, C1, Nimg, H, W, (const __half *)dW.p, OC, ep, (const uint8_t *)dWh.p)) return 1;
    cudaStream_t st = current_stream();
    gemm_launch(L, st);
    B200_CHECK(cudaGetLastError());
    B200_CHECK(cudaStreamSynchronize(st));
    B200_CHECK(cudaMemcpy(out32, dO.p, px * OC * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// K3 in isolation: depthwise 3x3 (stride 1|2, pad 1) + scale/shift + SiLU.  x [N,H,W,C] f16, Wt [3][3][C] f16, out [N,H/s,W/s,C] f16.
// variant 0 = the TMA kernel of the fused plan (k_dwconv_tma), 1 = the register-window fallback (k_dwconv).
extern "C" int ggml_b200_debug_dwconv(const uint16_t * x, int N, int H, int W, int C, int stride, const uint16_t * Wt, const float * scale,
                                      const float * shift, int act, int variant, uint16_t * out16) {
    ensure_device();
    if (stride < 1 || H % stride || W % stride) return 1;
    const size_t in_e = (size_t)N * H * W * C, out_e = (size_t)N * (H / stride) * (W / stride) * C;
    DevBuf dX(x, in_e * 2), dW(Wt, (size_t)9 * C * 2), dS(scale, scale ? (size_t)C * 4 : 0), dH(shift, shift ? (size_t)C * 4 : 0), dO(nullptr, out_e * 2);
    cudaStream_t st = current_stream();
    if (variant == 0) {
        DwLaunch L;
        if (!dw_prepare(L, (const __half *)dX.p, N, H, W, C, stride, (const __half *)dW.p, (const float *)dS.p, (const float *)dH.p, act, (__half *)dO.p)) return 1;
        dw_launch(L, st);
    } else {
        launch_dwconv((const __half *)dX.p, N, H, W, C, stride, (const __half *)dW.p, (const float *)dS.p, (const float *)dH.p, act, (__half *)dO.p, st);
    }
    B200_CHECK(cudaGetLastError());
    B200_CHECK(cudaStreamSynchronize(st));
    B200_CHECK(cudaMemcpy(out16, dO.p, out_e * 2, cudaMemcpyDeviceToHost));
    return 0;
}

// K2 in isolation: stem 3x3 / stride 2 / pad 1 over f32 images (HWC: chw == 0, CHW: chw == 1), Wt [OC][3][3][3] f16.
extern "C" int ggml_b200_debug_stem(const float * x, int chw, int N, int H, int W, const uint16_t * Wt, int OC, const float * scale, const float * shift,
                                    int act, uint16_t * out16, float * out32) {
    ensure_device();
    if (H % 2 || W % 2) return 1;
    const size_t out_e = (size_t)N * (H / 2) * (W / 2) * OC;
    DevBuf dX(x, (size_t)N * H * W * 3 * 4), dW(Wt, (size_t)OC * 27 * 2), dS(scale, scale ? (size_t)OC * 4 : 0), dH(shift, shift ? (size_t)OC * 4 : 0);
    DevBuf dO16(nullptr, out16 ? out_e * 2 : 0), dO32(nullptr, out32 ? out_e * 4 : 0);
    int64_t sn, sy, sx, sc;
    if (chw) { sn = (int64_t)3 * H * W; sy = W; sx = 1; sc = (int64_t)H * W; }
    else { sn = (int64_t)H * W * 3; sy = (int64_t)W * 3; sx = 3; sc = 1; }
    cudaStream_t st = current_stream();
    launch_stem((const float *)dX.p, sn, sy, sx, sc, N, H, W, (const __half *)dW.p, OC, (const float *)dS.p, (const float *)dH.p, act, (__half *)dO16.p,
                (float *)dO32.p, st);
    B200_CHECK(cudaGetLastError());
    B200_CHECK(cudaStreamSynchronize(st));
    if (out16) B200_CHECK(cudaMemcpy(out16, dO16.p, out_e * 2, cudaMemcpyDeviceToHost));
    if (out32) B200_CHECK(cudaMemcpy(out32, dO32.p, out_e * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// K7 in isolation.  qkv: f16 [N*H*W][3][heads][DP] (DP = ggml_b200_debug_attention_dp(C/heads), padding zero), pixel order;
// out: f16 [N*H*W][C].  A sequence = the (H/2)*(W/2) pixels of one image sharing (y%2, x%2)  (main.cpp:721-747, 1073-1086).
extern "C" int ggml_b200_debug_attention_dp(int d) { return attention_padded_head_dim(d); }
extern "C" int ggml_b200_debug_attention(const uint16_t * qkv, int N, int H, int W, int C, int heads, uint16_t * out16) {
    ensure_device();
    if (heads <= 0 || C % heads || H % 2 || W % 2) return 1;
    const int    dp = attention_padded_head_dim(C / heads);
    const size_t px = (size_t)N * H * W;
    DevBuf dQ(qkv, px * 3 * heads * dp * 2), dO(nullptr, px * C * 2);
    cudaStream_t st = current_stream();
    launch_attention((const __half *)dQ.p, N, H, W, C, heads, (__half *)dO.p, st);
    B200_CHECK(cudaGetLastError());
    B200_CHECK(cudaStreamSynchronize(st));
    B200_CHECK(cudaMemcpy(out16, dO.p, px * C * 2, cudaMemcpyDeviceToHost));
    return 0;
}

// Average device time (ms) of `reps` back-to-back attention launches on the data of ggml_b200_debug_attention (timing probe).
extern "C" float ggml_b200_debug_attention_time(const uint16_t * qkv, int N, int H, int W, int C, int heads, int reps) {
    ensure_device();
    if (heads <= 0 || C % heads || H % 2 || W % 2 || reps <= 0) return -1.f;
    const int    dp = attention_padded_head_dim(C / heads);
    const size_t px = (size_t)N * H * W;
    DevBuf dQ(qkv, px * 3 * heads * dp * 2), dO(nullptr, px * C * 2);
    cudaStream_t st = current_stream();
    cudaEvent_t  e0, e1;
    B200_CHECK(cudaEventCreate(&e0));
    B200_CHECK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) launch_attention((const __half *)dQ.p, N, H, W, C, heads, (__half *)dO.p, st);
    B200_CHECK(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; i++) launch_attention((const __half *)dQ.p, N, H, W, C, heads, (__half *)dO.p, st);
    B200_CHECK(cudaEventRecord(e1, st));
    B200_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    B200_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms / (float)reps;
}

// LayerNorm folded around two GEMMs, in isolation:
//   producer  x = A[M,K] . B[C,K]^T + shift0[C]      -> f32 x, f16 x, per-row (sum, sum of squares)
//   consumer  y = act( LN(x; gamma, beta, eps) . W[N,C]^T + bias[N] ), computed as r * (f16(x) . W'^T - mu * c1) + shift'
// W is f32 [N][C]; the host folding below is the one fuse.cpp applies (W' = f16(W * gamma), c1 = row sums of W',
// shift' = bias + sum_k beta_k W[n][k]).  x32 (optional) returns the producer's f32 output for the two-pass reference.
extern "C" int ggml_b200_debug_gemm_ln(const uint16_t * A, const uint16_t * B, int M, int C, int K, const float * shift0, const float * gamma,
                                       const float * beta, float eps, const float * Wf, const float * bias, int N, int act, float * x32, float * y32) {
    ensure_device();
    if (C > 256 || C % 8 || N % 8 || K % 8) return 1;
    std::vector<uint16_t> wt((size_t)N * C);
    std::vector<float>    c1(N), sh(N);
    for (int n = 0; n < N; n++) {
        double a1 = 0.0, a2 = 0.0;
        for (int k = 0; k < C; k++) {
            const float    w  = Wf[(size_t)n * C + k];
            const uint16_t wg = ggml_fp32_to_fp16(w * gamma[k]);
            wt[(size_t)n * C + k] = wg;
            a1 += (double)ggml_fp16_to_fp32(wg);
            a2 += (double)beta[k] * (double)w;
        }
        c1[n] = (float)a1;
        sh[n] = (bias ? bias[n] : 0.f) + (float)a2;
    }
    DevBuf dA(A, (size_t)M * K * 2), dB(B, (size_t)C * K * 2), dS0(shift0, shift0 ? (size_t)C * 4 : 0);
    DevBuf dX32(nullptr, (size_t)M * C * 4), dX16(nullptr, (size_t)M * C * 2), dSt(nullptr, (size_t)M * 8);
    DevBuf dW(wt.data(), wt.size() * 2), dC1(c1.data(), (size_t)N * 4), dSh(sh.data(), (size_t)N * 4), dY(nullptr, (size_t)M * N * 4);
    GemmEpilogue e0;
    e0.shift = (const float *)dS0.p;
    e0.out32 = (float *)dX32.p; e0.ld32 = C;
    e0.out16 = (__half *)dX16.p; e0.ld16 = C;
    e0.stats_out = (float *)dSt.p;
    GemmLaunch L0, L1;
    if (!gemm_prepare(L0, (const __half *)dA.p, K, (const __half *)dB.p, K, M, C, K, e0)) return 1;
    GemmEpilogue e1;
    e1.shift = (const float *)dSh.p;
    e1.act   = act;
    e1.out32 = (float *)dY.p; e1.ld32 = N;
    e1.ln_stats = (const float *)dSt.p;
    e1.ln_c1    = (const float *)dC1.p;
    e1.ln_inv_c = 1.0f / (float)C;
    e1.ln_eps   = eps;
    if (!gemm_prepare(L1, (const __half *)dX16.p, C, (const __half *)dW.p, C, M, N, C, e1)) return 1;
    cudaStream_t st = current_stream();
    gemm_launch(L0, st);
    gemm_launch(L1, st);
    B200_CHECK(cudaGetLastError());
    B200_CHECK(cudaStreamSynchronize(st));
    if (x32) B200_CHECK(cudaMemcpy(x32, dX32.p, (size_t)M * C * 4, cudaMemcpyDeviceToHost));
    B200_CHECK(cudaMemcpy(y32, dY.p, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// K4 in isolation: expand 1x1 (+BN+SiLU) -> depthwise 3x3 (+BN+SiLU) -> reduce 1x1 (+BN) [+ f32 residual].
// x [N,H,W,Cin], We [E][Cin], Wd [3][3][E], Wr [Cout][E] (all f16); returns 2 if the shape is outside the fused kernel's envelope.
extern "C" int ggml_b200_debug_ir_fused(const uint16_t * x, int N, int H, int W, int Cin, int E, int Cout, int stride, const uint16_t * We,
                                        const float * se, const float * he, const uint16_t * Wd, const float * sd, const float * hd, const uint16_t * Wr,
                                        const float * sr, const float * hr, const float * res32, uint16_t * out16, float * out32) {
    ensure_device();
    if (stride < 1 || H % stride || W % stride) return 2;
    const size_t opx = (size_t)N * (H / stride) * (W / stride);
    DevBuf dX(x, (size_t)N * H * W * Cin * 2), dWe(We, (size_t)E * Cin * 2), dWd(Wd, (size_t)9 * E * 2), dWr(Wr, (size_t)Cout * E * 2);
    DevBuf dse(se, (size_t)E * 4), dhe(he, (size_t)E * 4), dsd(sd, (size_t)E * 4), dhd(hd, (size_t)E * 4), dsr(sr, (size_t)Cout * 4), dhr(hr, (size_t)Cout * 4);
    DevBuf dR(res32, res32 ? opx * Cout * 4 : 0), dO16(nullptr, out16 ? opx * Cout * 2 : 0), dO32(nullptr, out32 ? opx * Cout * 4 : 0);
    IrLaunch L;
    if (!ir_fused_prepare(L, (const __half *)dX.p, N, H, W, Cin, E, Cout, stride, (const __half *)dWe.p, (const float *)dse.p, (const float *)dhe.p,
                          (const __half *)dWd.p, (const float *)dsd.p, (const float *)dhd.p, (const __half *)dWr.p, (const float *)dsr.p,
                          (const float *)dhr.p, (const float *)dR.p, (__half *)dO16.p, (float *)dO32.p))
        return 2;
    cudaStream_t st = current_stream();
    ir_fused_launch(L, st);
    B200_CHECK(cudaGetLastError());
    B200_CHECK(cudaStreamSynchronize(st));
    if (out16) B200_CHECK(cudaMemcpy(out16, dO16.p, opx * Cout * 2, cudaMemcpyDeviceToHost));
    if (out32) B200_CHECK(cudaMemcpy(out32, dO32.p, opx * Cout * 4, cudaMemcpyDeviceToHost));
    return 0;
}
// timing probe (uninitialised device buffers): mean ms per launch of the fused block, and of the three separate kernels it replaces
extern "C" float ggml_b200_debug_ir_time(int N, int H, int W, int Cin, int E, int Cout, int stride, int with_res, int reps, float * unfused_ms) {
    ensure_device();
    const size_t ipx = (size_t)N * H * W, opx = (size_t)N * (H / stride) * (W / stride);
    DevBuf dX(nullptr, ipx * Cin * 2), dWe(nullptr, (size_t)E * Cin * 2), dWd(nullptr, (size_t)9 * E * 2), dWr(nullptr, (size_t)Cout * E * 2);
    DevBuf dse(nullptr, (size_t)E * 4), dhe(nullptr, (size_t)E * 4), dsr(nullptr, (size_t)Cout * 4);
    DevBuf dR(nullptr, with_res ? opx * Cout * 4 : 0), dO16(nullptr, opx * Cout * 2), dO32(nullptr, with_res ? opx * Cout * 4 : 0);
    DevBuf dE(nullptr, ipx * E * 2), dD(nullptr, opx * E * 2);
    cudaStream_t st = current_stream();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = -1.f;
    IrLaunch L;
    if (ir_fused_prepare(L, (const __half *)dX.p, N, H, W, Cin, E, Cout, stride, (const __half *)dWe.p, (const float *)dse.p, (const float *)dhe.p,
                         (const __half *)dWd.p, (const float *)dse.p, (const float *)dhe.p, (const __half *)dWr.p, (const float *)dsr.p,
                         (const float *)dsr.p, (const float *)dR.p, (__half *)dO16.p, (float *)dO32.p)) {
        ir_fused_launch(L, st);
        cudaEventRecord(e0, st);
        for (int i = 0; i < reps; i++) ir_fused_launch(L, st);
        cudaEventRecord(e1, st);
        B200_CHECK(cudaStreamSynchronize(st));
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= reps;
        if (getenv("GGML_B200_IR_PHASES")) {  // clock64 profile of compute thread 0 and of the control thread, averaged over the CTAs
            DevBuf dT(nullptr, (size_t)L.grid * 16 * sizeof(long long));
            L.p.timing = (long long *)dT.p;
            ir_fused_launch(L, st);
            B200_CHECK(cudaStreamSynchronize(st));
            std::vector<long long> hT((size_t)L.grid * 16);
            B200_CHECK(cudaMemcpy(hT.data(), dT.p, hT.size() * sizeof(long long), cudaMemcpyDeviceToHost));
            static const char * nm[2][8] = {{"wait exp_full", "expand epilogue", "barrier", "reduce epilogue", "wait red_done", "depthwise", "barrier", "other"},
                                            {"other", "wait es_done", "wait x", "wait we", "wait as_full", "wait wr", "MMA issue", "TMA issue"}};
            for (int who = 0; who < 2; who++) {
                double sum[8] = {0}, tot = 0;
                for (int b = 0; b < L.grid; b++) for (int j = 0; j < 8; j++) sum[j] += (double)hT[((size_t)b * 2 + who) * 8 + j];
                for (int j = 0; j < 8; j++) tot += sum[j];
                for (int j = 0; j < 8; j++)
                    if (sum[j] > 0) fprintf(stderr, "   %s %-26s %6.1f %%  (%.0f cycles per CTA)\n", who ? "control" : "compute", nm[who][j], 100.0 * sum[j] / tot, sum[j] / L.grid);
            }
            L.p.timing = nullptr;
        }
        fprintf(stderr, "ir_fused %dx%dx%d %d->%d->%d s%d: tile %dx%d halo %d rows (%d blocks) NT=%d nxb=%d smem=%zu tmem=%d grid=%d  %.1f us\n", N, H, W, Cin, E, Cout,
                stride, L.p.TH, L.p.TW, L.p.P_in, L.p.MBI, L.p.nthreads, L.p.nxb, L.smem_bytes, L.p.tmem_cols, L.grid, 1e3f * ms);
    }
    if (unfused_ms) {
        GemmEpilogue e1x; e1x.scale = (const float *)dse.p; e1x.shift = (const float *)dhe.p; e1x.act = 1; e1x.out16 = (__half *)dE.p; e1x.ld16 = E;
        GemmEpilogue e3x; e3x.scale = (const float *)dsr.p; e3x.shift = (const float *)dsr.p; e3x.out16 = (__half *)dO16.p; e3x.ld16 = Cout;
        if (with_res) { e3x.res32 = (const float *)dR.p; e3x.ldr32 = Cout; e3x.out32 = (float *)dO32.p; e3x.ld32 = Cout; }
        GemmLaunch G1, G3;
        DwLaunch D;
        if (gemm_prepare(G1, (const __half *)dX.p, Cin, (const __half *)dWe.p, Cin, (int)ipx, E, Cin, e1x) &&
            dw_prepare(D, (const __half *)dE.p, N, H, W, E, stride, (const __half *)dWd.p, (const float *)dse.p, (const float *)dhe.p, 1, (__half *)dD.p) &&
            gemm_prepare(G3, (const __half *)dD.p, E, (const __half *)dWr.p, E, (int)opx, Cout, E, e3x)) {
            gemm_launch(G1, st); dw_launch(D, st); gemm_launch(G3, st);
            cudaEventRecord(e0, st);
            for (int i = 0; i < reps; i++) { gemm_launch(G1, st); dw_launch(D, st); gemm_launch(G3, st); }
            cudaEventRecord(e1, st);
            B200_CHECK(cudaStreamSynchronize(st));
            float u = 0;
            cudaEventElapsedTime(&u, e0, e1);
            *unfused_ms = u / reps;
        } else {
            *unfused_ms = -1.f;
        }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms;
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

// ---- tcgen05.mma issue-rate probe (tests/mma_rate_probe.py): one thread per CTA issues `iters` back-to-back MMAs of shape 128 x N x 16 on
// the same shared-memory operands and waits for the final commit; returns cycles per MMA.  ctas_per_sm CTAs share every SM. ----
#include "ptx_sm100.cuh"
// warp-uniform issue: every lane of the warp executes the instruction stream with identical operands, one elected lane issues
__device__ __forceinline__ void umma_f16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__global__ void k_mma_rate(int N, int iters, int commit_every, long long * out, int uniform) {
    using namespace b200::ptx;
    extern __shared__ uint8_t probe_smem_raw[];
    uint8_t * smem = probe_smem_raw + ((1024u - (smem_u32(probe_smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&slot), 256);
    for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;  // ones
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (uniform && threadIdx.x < 32) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t adesc = make_smem_desc(smem_u32(smem), 128), bdesc = make_smem_desc(smem_u32(smem) + 16384, 128);
        const uint32_t tm = __shfl_sync(0xffffffffu, slot, 0);
        const long long t0 = clock64();
        if (commit_every < -1) {
            // the conv loop's shape: groups of -commit_every MMAs alternating between two accumulators, an asynchronous commit (to a
            // barrier nobody waits for) after every group
            __shared__ __align__(8) uint64_t sink;
            if (threadIdx.x == 0) mbar_init(smem_u32(&sink), 1 << 20);
            __syncwarp();
            const int g = -commit_every;
            for (int i = 0; i < iters; i += g) {
                for (int j = 0; j < g; j++)
                    umma_f16_elect(tm + (uint32_t)((j >> 2) & 1) * (uint32_t)N, adesc + (uint64_t)(2 * (j & 3)), bdesc + (uint64_t)(2 * (j & 3)), idesc, 1);
                umma_commit_ws(smem_u32(&sink));
            }
        } else
        for (int i = 0; i < iters; i++) umma_f16_elect(tm, adesc + (uint64_t)(2 * (i & 3)), bdesc + (uint64_t)(2 * (i & 3)), idesc, 1);
        const long long t1 = clock64();
        if (threadIdx.x == 0) umma_commit(smem_u32(&bar));
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        if (threadIdx.x == 0) {
            out[blockIdx.x * 2]     = t1 - t0;
            out[blockIdx.x * 2 + 1] = t2 - t0;
        }
    } else if (!uniform && threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t adesc = make_smem_desc(smem_u32(smem), 128), bdesc = make_smem_desc(smem_u32(smem) + 16384, 128);
        uint32_t phase = 0;
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
            umma_f16(slot, adesc + (uint64_t)(2 * (i & 3)), bdesc + (uint64_t)(2 * (i & 3)), idesc, 1);
            if (commit_every > 0 && (i + 1) % commit_every == 0) {
                umma_commit(smem_u32(&bar));
                mbar_wait(smem_u32(&bar), phase);
                phase ^= 1u;
            }
        }
        const long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), phase);
        const long long t2 = clock64();
        out[blockIdx.x * 2]     = t1 - t0;
        out[blockIdx.x * 2 + 1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(slot, 256); }
}
// returns cycles per MMA: issue-only in issue_cycles, issue + completion in the return value
extern "C" float ggml_b200_debug_mma_rate(int N, int iters, int ctas_per_sm, int commit_every, float * issue_cycles) {
    b200::ensure_device();
    const int grid = 148 * ctas_per_sm;
    long long * d = nullptr;
    cudaMalloc(&d, sizeof(long long) * 2 * grid);
    const size_t smem = 1024 + (128 + 256) * 128;
    cudaFuncSetAttribute(k_mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int uniform = commit_every < 0 ? 1 : 0;  // commit_every < 0: warp-uniform issue with an elected lane
    for (int rep = 0; rep < 2; rep++) k_mma_rate<<<grid, 128, smem>>>(N, iters, commit_every, d, uniform);
    cudaDeviceSynchronize();
    std::vector<long long> h((size_t)2 * grid);
    cudaMemcpy(h.data(), d, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
    cudaFree(d);
    double a = 0, b = 0;
    for (int i = 0; i < grid; i++) { a += (double)h[(size_t)2 * i]; b += (double)h[(size_t)2 * i + 1]; }
    if (issue_cycles) *issue_cycles = (float)(a / grid / iters);
    return (float)(b / grid / iters);
}

// K8 in isolation: n_layers transformer layers (main.cpp:988-1172) over the pixel-ordered f32 residual stream x [N,H,W,C].
// params: n_layers x 16 host pointers in VitLayerHost order (ln1_g ln1_b wq bq wk bk wv bv wo bo ln2_g ln2_b w1 b1 w2 b2), dense kernels
// f32 [in][out] as the weight file holds them.  reps > 0: returns the mean ms per launch in *ms (outputs then hold the last run).
extern "C" int ggml_b200_debug_vit_stage(const float * x, int N, int H, int W, int C, int heads, int F, int n_layers, float eps,
                                         const float * const * params, float * out32, uint16_t * out16, float * stats, int reps, float * ms) {
    ensure_device();
    if (!vit_stage_supported(N, H, W, C, heads, F)) return 2;
    std::vector<VitLayerHost> layers((size_t)n_layers);
    for (int l = 0; l < n_layers; l++) memcpy(&layers[(size_t)l], params + 16 * l, sizeof(VitLayerHost));
    std::vector<uint8_t> blob;
    std::vector<float>   vec;
    vit_stage_pack(layers.data(), n_layers, C, heads, F, blob, vec);
    const size_t px = (size_t)N * H * W;
    DevBuf dX(x, px * C * 4), dB(blob.data(), blob.size()), dV(vec.data(), vec.size() * 4);
    DevBuf dO32(nullptr, out32 ? px * C * 4 : 0), dO16(nullptr, out16 ? px * C * 2 : 0), dS(nullptr, stats ? px * 8 : 0);
    VitStageLaunch L;
    if (!vit_stage_prepare(L, (const float *)dX.p, N, H, W, C, heads, F, n_layers, eps, (const uint8_t *)dB.p, (const float *)dV.p, (float *)dO32.p,
                           (__half *)dO16.p, (float *)dS.p))
        return 2;
    cudaStream_t st = current_stream();
    vit_stage_launch(L, st);
    B200_CHECK(cudaGetLastError());
    B200_CHECK(cudaStreamSynchronize(st));
    if (reps > 0 && ms) {
        cudaEvent_t e0, e1;
        B200_CHECK(cudaEventCreate(&e0));
        B200_CHECK(cudaEventCreate(&e1));
        B200_CHECK(cudaEventRecord(e0, st));
        for (int i = 0; i < reps; i++) vit_stage_launch(L, st);
        B200_CHECK(cudaEventRecord(e1, st));
        B200_CHECK(cudaEventSynchronize(e1));
        B200_CHECK(cudaEventElapsedTime(ms, e0, e1));
        *ms /= (float)reps;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    if (out32) B200_CHECK(cudaMemcpy(out32, dO32.p, px * C * 4, cudaMemcpyDeviceToHost));
    if (out16) B200_CHECK(cudaMemcpy(out16, dO16.p, px * C * 2, cudaMemcpyDeviceToHost));
    if (stats) B200_CHECK(cudaMemcpy(stats, dS.p, px * 8, cudaMemcpyDeviceToHost));
    return 0;
}
