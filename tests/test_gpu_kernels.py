"""Every hand-written kernel of the FAST plan in isolation, against numpy float64 on the same seeded inputs
(`-m gpu`, through the C ABI: include/ggml/ggml.h `ggml_b200_debug_*`).  The whole-model parity tests
(test_gpu_parity.py) only see these kernels through a 1e-2 gate; here each one is pinned at its own precision:

  K3 depthwise (k_dwconv_tma / k_dwconv)   main.cpp:788,809-850      f16 products are exact in f32 -> f16 output rounding only
  K2 stem (k_stem_mma / k_stem)            main.cpp:798 (3->16, s2)  f16 operands, f32 accumulate
  K7 attention (k_attention_mma)           main.cpp:1073-1086        f16 Q/K/V/P operands, f32 online softmax
  K5 LayerNorm folded around two GEMMs     main.cpp:1002-1019        vs a two-pass float64 LayerNorm
  K4 fused inverted residual (k_ir_fused)  main.cpp:854-870          vs expand -> round -> depthwise -> round -> reduce in float64
"""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

u16p = ctypes.POINTER(ctypes.c_uint16)
f32p = ctypes.POINTER(ctypes.c_float)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _h(a):
    return _p(a.view(np.uint16), u16p)


def _silu(x):
    return x / (1.0 + np.exp(-x))


@pytest.fixture(scope="module")
def L():
    import ggml_experiments_b200 as G
    lib = G.lib_ggml()
    assert lib.ggml_b200_device_count() > 0, "no CUDA device: GPU tests need the B200 box"
    c_int, c_float = ctypes.c_int, ctypes.c_float
    lib.ggml_b200_debug_dwconv.argtypes = [u16p, c_int, c_int, c_int, c_int, c_int, u16p, f32p, f32p, c_int, c_int, u16p]
    lib.ggml_b200_debug_stem.argtypes = [f32p, c_int, c_int, c_int, c_int, u16p, c_int, f32p, f32p, c_int, u16p, f32p]
    lib.ggml_b200_debug_attention.argtypes = [u16p, c_int, c_int, c_int, c_int, c_int, u16p]
    lib.ggml_b200_debug_attention_dp.argtypes = [c_int]
    lib.ggml_b200_debug_gemm_ln.argtypes = [u16p, u16p, c_int, c_int, c_int, f32p, f32p, f32p, c_float, f32p, f32p, c_int, c_int, f32p, f32p]
    lib.ggml_b200_debug_ir_fused.argtypes = [u16p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, u16p, f32p, f32p, u16p, f32p, f32p,
                                             u16p, f32p, f32p, f32p, u16p, f32p]
    return lib


# ---- K3 depthwise -----------------------------------------------------------------------------------------
def _dw_ref(x, wt, stride):
    n, h, w, c = x.shape
    xp = np.pad(x.astype(np.float64), ((0, 0), (1, 1), (1, 1), (0, 0)))
    oh, ow = h // stride, w // stride
    acc = np.zeros((n, oh, ow, c))
    w64 = wt.astype(np.float64)
    for kh in range(3):
        for kw in range(3):
            acc += xp[:, kh:kh + stride * oh:stride, kw:kw + stride * ow:stride, :] * w64[kh, kw]
    return acc


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n,h,w,c,stride", [(2, 32, 32, 64, 1), (2, 32, 32, 64, 2), (1, 16, 16, 384, 2), (3, 16, 16, 512, 1),
                                            (2, 64, 64, 128, 2), (1, 128, 128, 64, 1), (2, 8, 8, 256, 1), (2, 12, 20, 192, 1),
                                            (3, 6, 6, 48, 1), (1, 4, 4, 32, 2), (5, 2, 2, 512, 2), (2, 24, 24, 96, 2),
                                            (1, 40, 56, 24, 1), (2, 10, 14, 72, 2)])
def test_depthwise_matches_numpy(L, variant, n, h, w, c, stride):
    rng = np.random.default_rng(n * 1000 + h * 7 + c + stride)
    x = rng.normal(size=(n, h, w, c)).astype(np.float16)
    wt = (rng.normal(size=(3, 3, c)) / 3.0).astype(np.float16)
    scale = rng.uniform(0.5, 1.5, c).astype(np.float32)
    shift = rng.normal(size=c).astype(np.float32) * 0.3
    acc = _dw_ref(x, wt, stride)
    for act in (0, 1):
        out = np.zeros((n, h // stride, w // stride, c), np.uint16)
        rc = L.ggml_b200_debug_dwconv(_h(x), n, h, w, c, stride, _h(wt), _p(scale, f32p), _p(shift, f32p), act, variant, _p(out, u16p))
        assert rc == 0
        ref = acc * scale + shift
        if act:
            ref = _silu(ref)
        got = out.view(np.float16).astype(np.float64)
        # f16 output rounding (2^-11 relative) + tanh.approx SiLU (2^-11): element gate, not a max-norm gate
        tol = (1.2e-3 if act else 6e-4) * np.abs(ref) + 2e-4
        assert (np.abs(got - ref) <= tol).all(), (variant, act, float(np.abs(got - ref).max()))


# ---- K2 stem ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,h,w,oc,chw", [(2, 64, 64, 16, 0), (1, 256, 256, 16, 0), (3, 32, 96, 16, 1), (2, 128, 64, 16, 0),
                                          (1, 64, 64, 24, 0), (2, 192, 192, 16, 1), (1, 34, 70, 16, 0), (2, 64, 64, 32, 0)])
def test_stem_matches_numpy(L, n, h, w, oc, chw):
    rng = np.random.default_rng(h * 3 + w + oc)
    img = rng.uniform(0, 1, size=(n, h, w, 3)).astype(np.float32)
    wt = (rng.normal(size=(oc, 3, 3, 3)) / np.sqrt(27)).astype(np.float16)
    scale = rng.uniform(0.5, 1.5, oc).astype(np.float32)
    shift = rng.normal(size=oc).astype(np.float32) * 0.2
    x16 = img.astype(np.float16).astype(np.float64)  # ggml's im2col rounds the activations to f16
    xp = np.pad(x16, ((0, 0), (1, 1), (1, 1), (0, 0)))
    oh, ow = h // 2, w // 2
    acc = np.zeros((n, oh, ow, oc))
    for kh in range(3):
        for kw in range(3):
            acc += np.einsum("nhwc,oc->nhwo", xp[:, kh:kh + 2 * oh:2, kw:kw + 2 * ow:2, :], wt[:, kh, kw, :].astype(np.float64))
    ref = _silu(acc * scale + shift)
    src = np.ascontiguousarray(img.transpose(0, 3, 1, 2)) if chw else img
    out16 = np.zeros((n, oh, ow, oc), np.uint16)
    out32 = np.zeros((n, oh, ow, oc), np.float32)
    assert L.ggml_b200_debug_stem(_p(src, f32p), chw, n, h, w, _h(wt), oc, _p(scale, f32p), _p(shift, f32p), 1, _p(out16, u16p), _p(out32, f32p)) == 0
    assert (np.abs(out32 - ref) <= 1e-3 * np.abs(ref) + 1e-4).all(), float(np.abs(out32 - ref).max())
    got16 = out16.view(np.float16).astype(np.float64)
    assert (np.abs(got16 - ref) <= 1.6e-3 * np.abs(ref) + 2e-4).all()


# ---- K7 attention ---------------------------------------------------------------------------------------------
def _attention_ref(qkv, n, h, w, heads, d, dp):
    """qkv: float64 [n, h, w, 3, heads, dp]; sequences = pixels sharing (y%2, x%2) of one image (unfold, main.cpp:721-747)."""
    out = np.zeros((n, h, w, heads * d))
    for py in range(2):
        for px in range(2):
            s = qkv[:, py::2, px::2]                      # [n, h/2, w/2, 3, heads, dp]
            s = s.reshape(n, -1, 3, heads, dp)[..., :d]   # [n, L, 3, heads, d]
            q, k, v = s[:, :, 0], s[:, :, 1], s[:, :, 2]
            sc = np.einsum("nqhd,nkhd->nhqk", q, k) / np.sqrt(d)
            sc -= sc.max(-1, keepdims=True)
            p = np.exp(sc)
            p /= p.sum(-1, keepdims=True)
            ctx = np.einsum("nhqk,nkhd->nqhd", p, v).reshape(n, h // 2, w // 2, heads * d)
            out[:, py::2, px::2] = ctx
    return out


@pytest.mark.parametrize("d", [16, 20, 24, 30, 36, 48, 60])
@pytest.mark.parametrize("n,h,w", [(2, 8, 8), (3, 16, 16), (1, 32, 32), (1, 48, 48), (1, 64, 64), (2, 16, 32)])
def test_attention_matches_numpy(L, d, n, h, w):
    """L = (h/2)*(w/2) in {16, 64, 256, 576, 1024, 128}: one-block, multi-query-block and multi-chunk K/V paths."""
    heads = 4
    c = heads * d
    dp = L.ggml_b200_debug_attention_dp(d)
    rng = np.random.default_rng(d * 100 + h)
    qkv = np.zeros((n, h, w, 3, heads, dp), np.float16)
    qkv[..., :d] = rng.normal(size=(n, h, w, 3, heads, d)).astype(np.float16)
    out = np.zeros((n, h, w, c), np.uint16)
    assert L.ggml_b200_debug_attention(_h(qkv), n, h, w, c, heads, _p(out, u16p)) == 0
    ref = _attention_ref(qkv.astype(np.float64), n, h, w, heads, d, dp)
    got = out.view(np.float16).astype(np.float64)
    # P is rounded to f16 for the P.V product and the result is stored as f16: ~1e-3 of the value scale
    err = np.abs(got - ref)
    assert err.max() < 4e-3 * max(1.0, np.abs(ref).max()), (d, h, w, float(err.max()))
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 2e-3


# ---- K5: LayerNorm folded around the GEMMs vs a two-pass LayerNorm ---------------------------------------------
@pytest.mark.parametrize("m,c,k,n,act", [(512, 144, 96, 432, 0), (300, 192, 128, 384, 1), (1024, 240, 160, 480, 1), (130, 64, 48, 64, 0),
                                         (256, 96, 64, 96, 0), (4096, 144, 288, 144, 0)])
@pytest.mark.parametrize("row_mean", [0.0, 8.0])
def test_layernorm_fold_matches_two_pass(L, m, c, k, n, act, row_mean):
    """row_mean = 8: every token has a mean of 8 standard deviations (the f16 copy of x that the consumer
    multiplies is then rounded at 8x the scale of the normalised value, and E[x^2]-mu^2 cancels 2 digits)."""
    rng = np.random.default_rng(m + c + n)
    A = rng.normal(size=(m, k)).astype(np.float16)
    B = (rng.normal(size=(c, k)) / np.sqrt(k)).astype(np.float16)
    shift0 = (rng.normal(size=c) * 0.1 + row_mean).astype(np.float32)
    gamma = rng.uniform(0.75, 1.25, c).astype(np.float32)
    beta = (rng.normal(size=c) * 0.1).astype(np.float32)
    Wf = (rng.normal(size=(n, c)) / np.sqrt(c)).astype(np.float32)
    bias = (rng.normal(size=n) * 0.02).astype(np.float32)
    eps = 1e-5
    x32 = np.zeros((m, c), np.float32)
    y32 = np.zeros((m, n), np.float32)
    assert L.ggml_b200_debug_gemm_ln(_h(A), _h(B), m, c, k, _p(shift0, f32p), _p(gamma, f32p), _p(beta, f32p), eps, _p(Wf, f32p),
                                     _p(bias, f32p), n, act, _p(x32, f32p), _p(y32, f32p)) == 0
    x = A.astype(np.float64) @ B.astype(np.float64).T + shift0
    assert np.abs(x32 - x).max() < 2e-3 * max(1.0, np.abs(x).max())
    # two-pass LayerNorm of the values the GPU actually produced (isolates the fold from the producer's own rounding)
    xg = x32.astype(np.float64)
    mu = xg.mean(1, keepdims=True)
    var = ((xg - mu) ** 2).mean(1, keepdims=True)
    ln = (xg - mu) / np.sqrt(var + eps) * gamma + beta
    ref = ln @ Wf.astype(np.float64).T + bias
    if act:
        ref = _silu(ref)
    err = np.abs(y32 - ref).max()
    scale = max(1.0, np.abs(ref).max())
    amp = 1.0 + row_mean  # rounding of the f16 copy is relative to |x|, the signal is relative to std(x) = 1
    print(f"ln-fold m={m} c={c} n={n} row_mean={row_mean}: max err {err:.2e} (scale {scale:.2f})")
    assert err < 1e-3 * amp * scale, (err, scale)


# ---- K4: fused inverted residual ----------------------------------------------------------------------------------
def _ir_ref(x, we, se, he, wd, sd, hd, wr, sr, hr, stride, res):
    """expand 1x1 (+BN+SiLU) -> f16 -> depthwise 3x3 (+BN+SiLU) -> f16 -> reduce 1x1 (+BN) [+ residual], float64 between
    the rounding points (inverted_residual_layer::forward, main.cpp:854-870; rounding points = ggml's im2col)."""
    e = _silu(x.astype(np.float64) @ we.astype(np.float64).T * se + he)
    e16 = e.astype(np.float16)
    dacc = _dw_ref(e16, wd, stride)
    dd = _silu(dacc * sd + hd).astype(np.float16)
    y = dd.astype(np.float64) @ wr.astype(np.float64).T * sr + hr
    if res is not None:
        y = y + res
    return y


@pytest.mark.parametrize("n,h,w,cin,e,cout,stride,res", [
    (2, 32, 32, 16, 64, 32, 1, False),    # S layer 1 (at 64x64 inputs)
    (2, 32, 32, 32, 128, 64, 2, False),   # S layer 2 downsample
    (3, 16, 16, 64, 256, 64, 1, True),    # S layer 2 residual blocks
    (2, 16, 16, 64, 256, 96, 2, False),   # ViT downsample
    (1, 128, 128, 16, 64, 32, 1, False),  # full-size layer 1 map
    (2, 64, 64, 64, 256, 64, 1, True),
    (1, 8, 8, 96, 384, 128, 2, False),
    (2, 4, 4, 128, 512, 160, 2, False),
    (3, 24, 24, 48, 192, 48, 1, True),    # XS: 48 / 192 channels, map width that does not divide the tile
    (2, 20, 12, 32, 128, 48, 2, False),
    (1, 16, 16, 16, 32, 16, 1, True),     # XXS layer 1: expand factor 2, residual
    (5, 6, 10, 24, 48, 24, 1, True),      # XXS: 24 channels
])
def test_fused_inverted_residual_matches_numpy(L, n, h, w, cin, e, cout, stride, res):
    rng = np.random.default_rng(h * 13 + cin + e + cout)
    x = rng.normal(size=(n, h, w, cin)).astype(np.float16)
    we = (rng.normal(size=(e, cin)) * np.sqrt(2.0 / cin)).astype(np.float16)
    wd = (rng.normal(size=(3, 3, e)) * np.sqrt(2.0 / 9)).astype(np.float16)
    wr = (rng.normal(size=(cout, e)) * np.sqrt(2.0 / e)).astype(np.float16)
    se, sd, sr = (rng.uniform(0.75, 1.25, k).astype(np.float32) for k in (e, e, cout))
    he, hd, hr = ((rng.normal(size=k) * 0.1).astype(np.float32) for k in (e, e, cout))
    oh, ow = h // stride, w // stride
    r32 = rng.normal(size=(n, oh, ow, cout)).astype(np.float32) if res else None
    out16 = np.zeros((n, oh, ow, cout), np.uint16)
    out32 = np.zeros((n, oh, ow, cout), np.float32)
    rc = L.ggml_b200_debug_ir_fused(_h(x), n, h, w, cin, e, cout, stride, _h(we), _p(se, f32p), _p(he, f32p), _h(wd), _p(sd, f32p),
                                    _p(hd, f32p), _h(wr), _p(sr, f32p), _p(hr, f32p), _p(r32, f32p), _p(out16, u16p), _p(out32, f32p))
    if rc == 2:
        pytest.skip("shape outside the fused kernel's envelope (the plan runs the three separate kernels)")
    assert rc == 0
    ref = _ir_ref(x, we, se, he, wd, sd, hd, wr, sr, hr, stride, r32)
    # two f16 rounding points + tanh.approx SiLU sit between input and output: errors of single roundings (2^-11 of a value ~1)
    # flow through the reduce GEMM; gate at a few f16 ulps of the output scale
    err = np.abs(out32 - ref)
    scale = max(1.0, np.abs(ref).max())
    rel = np.linalg.norm(out32 - ref) / np.linalg.norm(ref)
    print(f"ir_fused {n}x{h}x{w} {cin}->{e}->{cout} s{stride}: max err {err.max():.2e}, rel-L2 {rel:.2e}")
    assert err.max() < 4e-3 * scale and rel < 1.5e-3
    got16 = out16.view(np.float16).astype(np.float64)
    assert np.abs(got16 - out32).max() <= 1e-3 * scale


# ---- K8 fused transformer stage (vit_stage.cu) ---------------------------------------------------------------
def _vit_ref(x, layers, heads, eps):
    """n transformer layers (main.cpp:988-1172) in float64 over sequences = pixels sharing a patch position (main.cpp:721-747)."""
    n, h, w, c = x.shape
    seq_len, d = (h // 2) * (w // 2), c // heads
    t = x.astype(np.float64).reshape(n, h // 2, 2, w // 2, 2, c).transpose(0, 2, 4, 1, 3, 5).reshape(n * 4, seq_len, c)

    def ln(v, g, b):
        mu = v.mean(-1, keepdims=True)
        var = ((v - mu) ** 2).mean(-1, keepdims=True)
        return (v - mu) / np.sqrt(var + eps) * g + b

    for p in layers:
        (g1, b1, wq, bq, wk, bk, wv, bv, wo, bo, g2, b2, w1, bf1, w2, bf2) = [a.astype(np.float64) for a in p]
        y = ln(t, g1, b1)
        q, k, v = y @ wq + bq, y @ wk + bk, y @ wv + bv
        split = lambda a: a.reshape(n * 4, seq_len, heads, d).transpose(0, 2, 1, 3)
        s = split(q) @ split(k).transpose(0, 1, 3, 2) / np.sqrt(d)
        s = np.exp(s - s.max(-1, keepdims=True))
        s /= s.sum(-1, keepdims=True)
        a = (s @ split(v)).transpose(0, 2, 1, 3).reshape(n * 4, seq_len, c)
        t = t + a @ wo + bo
        y = ln(t, g2, b2)
        t = t + _silu(y @ w1 + bf1) @ w2 + bf2
    return t.reshape(n, 2, 2, h // 2, w // 2, c).transpose(0, 3, 1, 4, 2, 5).reshape(n, h, w, c)


def _vit_params(rng, c, f, n_layers):
    layers = []
    for _ in range(n_layers):
        dense = lambda i, o: (rng.standard_normal((i, o)) / np.sqrt(i)).astype(np.float32)
        vecr = lambda m, s=0.1: (rng.standard_normal(m) * s).astype(np.float32)
        layers.append([(1.0 + vecr(c)).astype(np.float32), vecr(c), dense(c, c), vecr(c), dense(c, c), vecr(c), dense(c, c), vecr(c),
                       (dense(c, c) * 0.5).astype(np.float32), vecr(c), (1.0 + vecr(c)).astype(np.float32), vecr(c), dense(c, f), vecr(f),
                       (dense(f, c) * 0.5).astype(np.float32), vecr(c)])
    return layers


@pytest.mark.parametrize("n,h,w,c,heads,f,nl", [
    (3, 16, 16, 192, 4, 384, 2),    # MobileViT-S stage 4 (L = 64)
    (5, 8, 8, 240, 4, 480, 3),      # S stage 5 (L = 16, head dim 60 padded to 64, hidden 480 = 3.75 chunks)
    (2, 16, 16, 120, 4, 240, 2),    # XS stage 4 (head dim 30 -> 32, C % 32 = 24, N padded to 128)
    (2, 8, 8, 144, 4, 288, 1),      # XS stage 5 (head dim 36 -> 48)
    (2, 16, 16, 80, 4, 160, 2),     # XXS stage 4 (head dim 20 -> 32)
    (7, 8, 8, 96, 4, 192, 2),       # XXS stage 5; 28 sequences of 16 tokens: the last tile is half empty
    (3, 4, 4, 64, 4, 128, 2),       # 128^2 input: L = 4 (eight sequences per warp)
    (1, 2, 2, 64, 4, 128, 1),       # 64^2 input: L = 1
    (1, 8, 16, 192, 4, 384, 1),     # non-square map, L = 32
    (100, 16, 16, 192, 4, 384, 1),  # 200 tiles: CTAs walk more than one tile
])
def test_vit_stage_fused_vs_float64(L, n, h, w, c, heads, f, nl):
    """k_vit_stage: LN -> qkv -> attention -> projection + residual -> LN -> MLP + residual, nl layers in one launch, vs float64.
    Operands are f16 (as in the unfused FAST plan), accumulation and the residual stream f32."""
    vpp = ctypes.POINTER(f32p)
    L.ggml_b200_debug_vit_stage.argtypes = [f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_float, vpp, f32p, u16p, f32p, ctypes.c_int, f32p]
    rng = np.random.default_rng(n * 1000 + c)
    x = rng.standard_normal((n, h, w, c)).astype(np.float32)
    x += rng.standard_normal((n, h, w, 1)).astype(np.float32) * 2.0  # per-token mean well away from zero (ADVICE: LayerNorm cancellation)
    layers = _vit_params(rng, c, f, nl)
    flat = [a for p in layers for a in p]
    arr = (f32p * len(flat))(*[_p(a, f32p) for a in flat])
    out32 = np.zeros((n, h, w, c), np.float32)
    out16 = np.zeros((n, h, w, c), np.float16)
    stats = np.zeros((n, h, w, 2), np.float32)
    rc = L.ggml_b200_debug_vit_stage(_p(x, f32p), n, h, w, c, heads, f, nl, 1e-5, arr, _p(out32, f32p), _h(out16), _p(stats, f32p), 0, None)
    assert rc == 0
    ref = _vit_ref(x, layers, heads, 1e-5)
    rms = float(np.sqrt((ref ** 2).mean()))
    err = np.abs(out32 - ref)
    rel = float(np.sqrt(((out32 - ref) ** 2).sum() / (ref ** 2).sum()))
    print(f"vit_stage n={n} {h}x{w} C={c} F={f} layers={nl}: rel-L2 {rel:.2e}, max-abs {err.max():.2e} (rms {rms:.2f})")
    assert np.isfinite(out32).all()
    assert rel < 3e-3 and err.max() < 3e-2 * rms
    assert np.abs(out16.astype(np.float32) - out32).max() <= np.abs(out32).max() * 2.0 ** -10
    assert np.allclose(stats[..., 0], out32.sum(-1), rtol=1e-4, atol=1e-3 * rms * c ** 0.5)
    assert np.allclose(stats[..., 1], (out32.astype(np.float64) ** 2).sum(-1), rtol=1e-4)


@pytest.mark.parametrize("n,h,w,c,f", [(3, 32, 32, 144, 288), (1, 32, 32, 96, 192), (2, 64, 64, 144, 288), (1, 6, 10, 64, 128), (40, 32, 32, 144, 288),
                                       (2, 16, 16, 240, 480)])
def test_vit_mlp_only_fused_vs_float64(L, n, h, w, c, f):
    """k_vit_stage with heads = 0: x + W2.silu(W1.LN(x) + b1) + b2 per token (the MLP half of transformer_layer::forward, main.cpp:1113-1165)
    for stages whose sequences do not fit a 128-token tile; any map size, tiles are 128 consecutive pixels."""
    vpp = ctypes.POINTER(f32p)
    L.ggml_b200_debug_vit_stage.argtypes = [f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_float, vpp, f32p, u16p, f32p, ctypes.c_int, f32p]
    rng = np.random.default_rng(n * 77 + c)
    x = rng.standard_normal((n, h, w, c)).astype(np.float32) + rng.standard_normal((n, h, w, 1)).astype(np.float32) * 3.0
    p = _vit_params(rng, c, f, 1)[0]
    arr = (f32p * 16)(*[_p(a, f32p) for a in p])
    out32 = np.zeros((n, h, w, c), np.float32)
    stats = np.zeros((n, h, w, 2), np.float32)
    rc = L.ggml_b200_debug_vit_stage(_p(x, f32p), n, h, w, c, 0, f, 1, 1e-5, arr, _p(out32, f32p), None, _p(stats, f32p), 0, None)
    assert rc == 0
    g2, b2, w1, bf1, w2, bf2 = [a.astype(np.float64) for a in p[10:16]]
    t = x.astype(np.float64)
    mu = t.mean(-1, keepdims=True)
    y = (t - mu) / np.sqrt(((t - mu) ** 2).mean(-1, keepdims=True) + 1e-5) * g2 + b2
    ref = t + _silu(y @ w1 + bf1) @ w2 + bf2
    rel = float(np.sqrt(((out32 - ref) ** 2).sum() / (ref ** 2).sum()))
    print(f"vit_mlp n={n} {h}x{w} C={c} F={f}: rel-L2 {rel:.2e}, max-abs {np.abs(out32 - ref).max():.2e}")
    assert np.isfinite(out32).all() and rel < 2e-3
    assert np.allclose(stats[..., 0], out32.sum(-1), rtol=1e-4, atol=1e-2)
