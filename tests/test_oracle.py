"""CPU tests: the oracle against the HF-torch golden fixtures (tests/golden, made by tests/gen_golden.py),
op-level checks of the oracle against numpy, and weight-file round trips."""
import os

import numpy as np
import pytest

from ggml_experiments_b200 import weights as W
from tests.util import parity_report

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("variant,hw", [("xxs", 256), ("xs", 256), ("s", 256), ("xxs", 128)])
def test_oracle_matches_hf_golden(oracle, weight_files, variant, hw):
    g = np.load(os.path.join(GOLD, f"hf_{variant}_{hw}.npz"))
    m = oracle.OracleModel(weight_files[variant])
    assert m.num_tensors == 313
    imgs = W.synthetic_images(int(g["n_img"]), hw, hw, seed=7)
    st = int(g["feat_stride"])
    # pure-f32 mode == HF torch semantics: only accumulation-order noise is allowed
    f32, p32 = m.forward(imgs, oracle.PURE_F32)
    r = parity_report(f32[:, ::st], g["feat"], rtol=1e-4, atol_rms=1e-4)
    assert r["violations"] == 0 and r["rel_l2"] < 2e-5, r
    assert np.abs(p32 - g["pooled"]).max() < 1e-4
    # ggml-faithful mode differs from HF only by the f16 rounding points of the conv path
    f16, p16 = m.forward(imgs, 0)
    r = parity_report(f16[:, ::st], g["feat"], rtol=5e-2, atol_rms=5e-2)
    assert r["rel_l2"] < 6e-3, r
    assert (p16.argmax(1) == g["pooled"].argmax(1)).all()


def test_weight_counts_match_survey(weight_files, oracle):
    # 313 tensors; 4 949 888 / 1 941 296 / 955 136 floats
    expect = {"s": 4949888, "xs": 1941296, "xxs": 955136}
    for v, n in expect.items():
        m = oracle.OracleModel(weight_files[v])
        assert (m.num_tensors, m.num_weights) == (313, n)


def test_weight_file_roundtrip(tmp_path):
    t = W.make_synthetic_weights("xxs", seed=3)
    p = str(tmp_path / "w.ggml")
    W.write_weight_file(p, t)
    back = W.read_weight_file(p)
    assert list(back.keys()) == list(t.keys())
    for k in t:
        assert back[k].shape == t[k].shape and np.array_equal(back[k], t[k])


def test_reference_test_image_pattern():
    # main.cpp:680-688: img[y*768 + x*3 + c] = ((y*768 + x*3 + c) % 256) / 255
    img = W.synthetic_images(1)[0]
    y, x, c = 17, 201, 2
    assert img[y, x, c] == np.float32(((y * 768 + x * 3 + c) % 256) / 255.0)


def test_oracle_unfold_fold_roundtrip_and_layout(oracle):
    import ctypes
    L = oracle.lib()
    C, H, Wd, ps = 5, 8, 12, 2
    x = np.arange(C * H * Wd, dtype=np.float32).reshape(C, H, Wd)
    tok = np.empty(C * H * Wd, dtype=np.float32)
    f32p = ctypes.POINTER(ctypes.c_float)
    L.mvo_unfold(x.ctypes.data_as(f32p), C, H, Wd, ps, tok.ctypes.data_as(f32p))
    tok = tok.reshape(ps * ps, (H // ps) * (Wd // ps), C)  # [P][L][C]
    # HF MobileViTLayer.unfolding: patch p = (ph, pw), token l = (iph, ipw)
    for p in range(4):
        ph, pw = divmod(p, 2)
        for l in (0, 5, 23):
            iph, ipw = divmod(l, Wd // ps)
            assert np.array_equal(tok[p, l], x[:, iph * ps + ph, ipw * ps + pw])
    # fold is square-only in the reference (main.cpp:754)
    C, H = 3, 8
    x = np.random.default_rng(0).normal(size=(C, H, H)).astype(np.float32)
    tok = np.empty_like(x).ravel()
    back = np.empty_like(x)
    L.mvo_unfold(x.ctypes.data_as(f32p), C, H, H, 2, tok.ctypes.data_as(f32p))
    L.mvo_fold(tok.ctypes.data_as(f32p), C, (H // 2) ** 2, 2, back.ctypes.data_as(f32p))
    assert np.array_equal(back, x)


def test_oracle_conv_matches_numpy_f16_semantics(oracle):
    import ctypes
    L = oracle.lib()
    f32p = ctypes.POINTER(ctypes.c_float)
    rng = np.random.default_rng(1)
    C, H, Wd, OC = 6, 9, 7, 5
    x = rng.normal(size=(C, H, Wd)).astype(np.float32)
    k = rng.normal(size=(3, 3, C, OC)).astype(np.float32)  # TF layout
    for stride in (1, 2):
        OH, OW = (H + 2 - 3) // stride + 1, (Wd + 2 - 3) // stride + 1
        out = np.empty((OC, OH, OW), dtype=np.float32)
        L.mvo_conv2d(x.ctypes.data_as(f32p), C, H, Wd, k.ctypes.data_as(f32p), 3, 3, OC, stride, out.ctypes.data_as(f32p), 0)
        xr = x.astype(np.float16).astype(np.float64)
        kr = k.astype(np.float16).astype(np.float64)
        xp = np.pad(xr, ((0, 0), (1, 1), (1, 1)))
        ref = np.zeros((OC, OH, OW))
        for oy in range(OH):
            for ox in range(OW):
                patch = xp[:, oy * stride:oy * stride + 3, ox * stride:ox * stride + 3]  # [C,3,3]
                ref[:, oy, ox] = np.einsum("chw,hwco->o", patch, kr)
        assert np.abs(out - ref).max() < 1e-4


def test_oracle_layernorm_softmax(oracle):
    import ctypes
    L = oracle.lib()
    f32p = ctypes.POINTER(ctypes.c_float)
    rng = np.random.default_rng(2)
    x = rng.normal(size=(7, 24)).astype(np.float32)
    g = rng.normal(size=24).astype(np.float32)
    b = rng.normal(size=24).astype(np.float32)
    y = np.empty_like(x)
    L.mvo_layernorm(x.ctypes.data_as(f32p), 24, 7, g.ctypes.data_as(f32p), b.ctypes.data_as(f32p), 1e-5, y.ctypes.data_as(f32p))
    xd = x.astype(np.float64)
    ref = (xd - xd.mean(1, keepdims=True)) / np.sqrt(xd.var(1, keepdims=True) + 1e-5) * g + b
    assert np.abs(y - ref).max() < 1e-5
    s = x.copy()
    L.mvo_softmax_rows(s.ctypes.data_as(f32p), 24, 7)
    e = np.exp(xd - xd.max(1, keepdims=True))
    assert np.abs(s - e / e.sum(1, keepdims=True)).max() < 1e-6


# ---- classifier head ---------------------------------------------------------------------------------------
def test_oracle_classifier_matches_hf_golden(oracle, weight_files):
    """HF MobileViTForImageClassification (backbone + pool + Linear) pins the head: kernel (in,out) layout and bias."""
    g = np.load(os.path.join(GOLD, "hf_cls_xxs_256.npz"))
    m = oracle.OracleModel(weight_files["xxs_cls"])
    assert (m.num_tensors, m.num_classes) == (315, 1000)
    imgs = W.synthetic_images(2, 256, 256, seed=7)
    _, pooled = m.forward(imgs, oracle.PURE_F32)
    logits = m.classify(pooled)
    assert np.abs(logits - g["logits"]).max() < 2e-4 * np.abs(g["logits"]).max()
    assert (logits.argmax(1) == g["logits"].argmax(1)).all()
    # a file without the head has no classes
    assert oracle.OracleModel(weight_files["xxs"]).num_classes == 0


# ---- sam_image_preprocess ----------------------------------------------------------------------------------
def preprocess_numpy(img, H, W_):
    """Independent numpy-f32 restatement of main.cpp:538-601 (stride-W variant): every op is a separate f32 operation."""
    f = np.float32
    ny, nx, _ = img.shape
    scale = f(max(nx, ny)) * f(1.0) / f(W_) if H == W_ else max(f(nx) / f(W_), f(ny) / f(H))
    nx3, ny3 = min(int(f(nx) / scale + f(0.5)), W_), min(int(f(ny) / scale + f(0.5)), H)
    sx = (np.arange(nx3, dtype=f) + f(0.5)) * scale - f(0.5)
    sy = (np.arange(ny3, dtype=f) + f(0.5)) * scale - f(0.5)
    x0 = np.clip(np.floor(sx).astype(np.int64), 0, nx - 1)
    y0 = np.clip(np.floor(sy).astype(np.int64), 0, ny - 1)
    x1, y1 = np.minimum(x0 + 1, nx - 1), np.minimum(y0 + 1, ny - 1)
    dx = (sx - x0.astype(f))[None, :, None]
    dy = (sy - y0.astype(f))[:, None, None]
    src = img.astype(f)
    v0 = src[y0][:, x0] * (f(1.0) - dx) + src[y0][:, x1] * dx
    v1 = src[y1][:, x0] * (f(1.0) - dx) + src[y1][:, x1] * dx
    v = v0 * (f(1.0) - dy) + v1 * dy
    q = np.clip(np.sign(v) * np.floor(np.abs(v) + f(0.5)), 0, 255).astype(np.uint8)  # std::round: half away from zero
    out = np.zeros((H, W_, 3), f)
    out[:ny3, :nx3] = q.astype(f) / f(255.0)
    return out


@pytest.mark.parametrize("sh,sw,H,W_", [(256, 256, 256, 256), (300, 400, 256, 256), (400, 300, 256, 256), (100, 80, 256, 256),
                                        (1080, 1920, 256, 256), (7, 5, 64, 64), (512, 512, 256, 256), (300, 500, 128, 256),
                                        (128, 256, 128, 256), (64, 64, 64, 64)])  # same size as the target: the identity the u8 stem route copies straight in
def test_oracle_preprocess_u8_is_the_reference_arithmetic(oracle, sh, sw, H, W_):
    img = np.random.default_rng(sh * 7 + sw).integers(0, 256, (2, sh, sw, 3), dtype=np.uint8)
    got = oracle.preprocess_u8(img, H, W_)
    for i in range(2):
        np.testing.assert_array_equal(got[i], preprocess_numpy(img[i], H, W_))
    if (sh, sw) == (H, W_):  # scale 1: the identity resize, exactly u8 / 255
        np.testing.assert_array_equal(got, img.astype(np.float32) / np.float32(255))
    if sh > sw and H == W_:  # portrait: the right part of the target stays zero, rows keep stride W (App. C #3)
        nx3 = int(np.float32(sw) / (np.float32(sh) / np.float32(W_)) + np.float32(0.5))
        assert (got[:, :, nx3:] == 0).all() and got[:, :, :nx3].max() > 0


# ---- torch-HF -> weight.ggml exporter ----------------------------------------------------------------------
def test_hf_exporter_roundtrip():
    """file tensors -> HF state dict -> file tensors is the identity (names, shapes, values) for backbone and head, and the
    exporter accepts MobileViTForImageClassification's `mobilevit.` prefix."""
    t = W.make_synthetic_weights("xxs", seed=3, num_classes=10)
    sd = W.to_hf_state_dict(t)
    back = W.from_hf_state_dict(sd)
    assert set(back) == set(t) and len(back) == 315
    for k in t:
        np.testing.assert_array_equal(back[k], t[k])
    prefixed = {(k if k.startswith("classifier.") else "mobilevit." + k): v for k, v in sd.items()}
    prefixed["mobilevit.conv_stem.normalization.num_batches_tracked"] = np.zeros((), np.int64)
    again = W.from_hf_state_dict(prefixed)
    assert set(again) == set(t)


def test_hf_exporter_on_a_real_hf_module(tmp_path, oracle):
    """Export a randomly initialised HF MobileViTForImageClassification and run the file through the oracle: logits == HF."""
    torch = pytest.importorskip("torch")
    tr = pytest.importorskip("transformers")
    torch.manual_seed(0)
    cfg = tr.MobileViTConfig(image_size=64, num_labels=7, **W.hf_config_kwargs("xxs"))
    model = tr.MobileViTForImageClassification(cfg).eval()
    with torch.no_grad():  # default init makes activations vanish: widen it so the comparison means something
        for n_, p_ in model.named_parameters():
            if p_.ndim > 1:
                p_.mul_(30.0)
    path = str(tmp_path / "hf.ggml")
    W.write_weight_file(path, W.from_hf_state_dict(model.state_dict()))
    imgs = W.synthetic_images(2, 64, 64, seed=7)
    with torch.no_grad():
        ref = model(torch.from_numpy(imgs).permute(0, 3, 1, 2).contiguous()).logits.numpy()
    m = oracle.OracleModel(path)
    assert (m.num_tensors, m.num_classes) == (315, 7)
    _, pooled = m.forward(imgs, oracle.PURE_F32)
    logits = m.classify(pooled)
    assert np.abs(logits - ref).max() < 1e-4 * max(1.0, np.abs(ref).max()), (np.abs(logits - ref).max(), np.abs(ref).max())


# ---- f16-on-disk and pre-transposed dense kernels (convert-tf-to-ggml.py:13-14 TODOs) ----------------------------
def test_f16_on_disk_and_pretransposed_files_load_to_the_same_model(oracle, tmp_path):
    import os
    t = W.make_synthetic_weights("xxs", seed=1234)
    paths = {}
    for tag, kw in {"ref": {}, "f16conv": {"f16": "conv"}, "pre": {"pretransposed": True}, "both": {"f16": "conv", "pretransposed": True},
                    "f16all": {"f16": "all"}}.items():
        paths[tag] = str(tmp_path / f"w_{tag}.ggml")
        W.write_weight_file(paths[tag], t, **kw)
    assert os.path.getsize(paths["f16conv"]) < 0.8 * os.path.getsize(paths["ref"])  # XXS: 47 % of the floats are convolution kernels
    assert os.path.getsize(paths["f16all"]) < 0.52 * os.path.getsize(paths["ref"])
    # python reader: canonical shapes come back; f16 conv payloads equal the f16 rounding the loader applies anyway
    back = W.read_weight_file(paths["both"])
    for k in t:
        assert back[k].shape == t[k].shape
        if "convolution" in k:
            assert np.array_equal(back[k], t[k].astype(np.float16).astype(np.float32))
        else:
            assert np.array_equal(back[k], t[k])
    # oracle loader: the lossless variants give BIT-identical features
    imgs = W.synthetic_images(2, 64, 64, seed=7)
    ref_f, ref_p = oracle.OracleModel(paths["ref"]).forward(imgs)
    for tag in ("f16conv", "pre", "both"):
        m = oracle.OracleModel(paths[tag])
        assert (m.num_tensors, m.num_weights) == (313, 955136)
        f, p = m.forward(imgs)
        assert np.array_equal(f, ref_f) and np.array_equal(p, ref_p), tag
    f, p = oracle.OracleModel(paths["f16all"]).forward(imgs)  # lossy for dense / norm tensors: close, not identical
    assert np.linalg.norm(f - ref_f) / np.linalg.norm(ref_f) < 5e-3


def test_host_loader_reads_the_extended_records(tmp_path):
    """include/mobilevit_b200.h mvit_load on the CPU (no compute): tensor and weight counts for every file flavour."""
    import ggml_experiments_b200 as G
    t = W.make_synthetic_weights("xs", seed=1234)
    for kw in ({}, {"f16": "conv"}, {"pretransposed": True}, {"f16": "all", "pretransposed": True}):
        p = str(tmp_path / "w.ggml")
        W.write_weight_file(p, t, **kw)
        m = G.MobileViT(p)
        try:
            assert (m.num_tensors, m.num_weights) == (313, 1941296), kw
        finally:
            m.close()
