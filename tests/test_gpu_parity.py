"""GPU parity tests (run with `-m gpu` on the B200 box).  Everything goes through the C ABI
(include/ggml/ggml.h, include/mobilevit_b200.h) and is compared with the CPU oracle on the same seeded inputs."""
import ctypes

import numpy as np
import pytest

from ggml_experiments_b200 import weights as W
from tests.util import parity_report, top1_report

pytestmark = pytest.mark.gpu

u16p = ctypes.POINTER(ctypes.c_uint16)
f32p = ctypes.POINTER(ctypes.c_float)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


@pytest.fixture(scope="module")
def G():
    import ggml_experiments_b200 as G
    assert G.lib_ggml().ggml_b200_device_count() > 0, "no CUDA device: GPU tests need the B200 box"
    L = G.lib_ggml()
    L.ggml_b200_debug_gemm.argtypes = [u16p, u16p, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p, f32p, ctypes.c_int, f32p,
                                       f32p, u16p]
    L.ggml_b200_debug_conv3x3.argtypes = [u16p, ctypes.c_int, u16p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          u16p, ctypes.c_int, f32p, f32p, ctypes.c_int, f32p]
    return G


def _silu(x):
    return x / (1.0 + np.exp(-x))


# ---- K1: tcgen05 GEMM in isolation ----------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(128, 32, 64), (128, 64, 16), (256, 32, 128), (1000, 96, 144), (4096, 640, 160),
                                   (333, 144, 96), (2048, 24, 48), (64, 256, 512), (512, 120, 120), (4096, 288, 144),
                                   (130, 8, 24)])
def test_gemm_tcgen05_matches_numpy(G, M, N, K):
    L = G.lib_ggml()
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.normal(size=(M, K)).astype(np.float16)
    B = (rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float16)
    scale = rng.uniform(0.5, 1.5, N).astype(np.float32)
    shift = rng.normal(size=N).astype(np.float32)
    res = rng.normal(size=(M, N)).astype(np.float32)
    ref_acc = A.astype(np.float64) @ B.astype(np.float64).T
    # plain
    out = np.zeros((M, N), np.float32)
    assert L.ggml_b200_debug_gemm(_p(A.view(np.uint16), u16p), _p(B.view(np.uint16), u16p), M, N, K, None, None, 0, None,
                                  _p(out, f32p), None) == 0
    assert np.abs(out - ref_acc).max() < 2e-3 * max(1.0, np.abs(ref_acc).max()), (M, N, K)
    # full epilogue: scale/shift + SiLU + residual, f32 and f16 outputs
    out32 = np.zeros((M, N), np.float32)
    out16 = np.zeros((M, N), np.uint16)
    assert L.ggml_b200_debug_gemm(_p(A.view(np.uint16), u16p), _p(B.view(np.uint16), u16p), M, N, K, _p(scale, f32p),
                                  _p(shift, f32p), 1, _p(res, f32p), _p(out32, f32p), _p(out16, u16p)) == 0
    ref = _silu(ref_acc * scale + shift) + res
    assert np.abs(out32 - ref).max() < 3e-3 * max(1.0, np.abs(ref).max())
    assert np.abs(out16.view(np.float16).astype(np.float64) - ref).max() < 6e-3 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("n_img,H,Wd,C0,C1,OC", [(2, 32, 32, 96, 0, 96), (3, 16, 16, 128, 128, 128), (5, 8, 8, 160, 160, 160),
                                                  (1, 8, 8, 64, 0, 80), (2, 64, 64, 48, 0, 48), (3, 4, 4, 80, 80, 80),
                                                  (1, 16, 16, 24, 24, 24),
                                                  # widths that do not divide 128: a tile is a whole number of image rows / images
                                                  (2, 24, 24, 96, 0, 96), (3, 40, 40, 96, 96, 96), (2, 20, 20, 128, 0, 128),
                                                  (4, 10, 10, 160, 160, 160), (5, 12, 12, 64, 0, 64), (3, 6, 6, 80, 0, 80),
                                                  (1, 28, 56, 48, 0, 48), (7, 5, 5, 32, 0, 40),
                                                  # halo mode with several tiles per CTA / a single image (N split into 64-column tiles)
                                                  (40, 32, 32, 96, 0, 96), (1, 8, 8, 160, 160, 160), (2, 16, 24, 64, 0, 72),
                                                  # pair mode (two M tiles per activation box and weight block): grids above 2 x 148 tiles
                                                  (80, 32, 32, 96, 0, 96), (40, 32, 32, 96, 96, 96), (160, 16, 16, 128, 128, 128),
                                                  (150, 16, 16, 128, 0, 128), (75, 32, 32, 64, 0, 72)])
def test_conv3x3_tcgen05_matches_numpy(G, n_img, H, Wd, C0, C1, OC):
    L = G.lib_ggml()
    rng = np.random.default_rng(H * 31 + C0 + OC)
    x0 = rng.normal(size=(n_img, H, Wd, C0)).astype(np.float16)
    x1 = rng.normal(size=(n_img, H, Wd, C1)).astype(np.float16) if C1 else None
    Wt = (rng.normal(size=(OC, 3, 3, C0 + C1)) / np.sqrt(9 * (C0 + C1))).astype(np.float16)
    out = np.zeros((n_img, H, Wd, OC), np.float32)
    rc = L.ggml_b200_debug_conv3x3(_p(x0.view(np.uint16), u16p), C0, _p(x1.view(np.uint16), u16p) if C1 else None, C1, n_img,
                                   H, Wd, _p(Wt.view(np.uint16), u16p), OC, None, None, 0, _p(out, f32p))
    assert rc == 0
    x = x0 if x1 is None else np.concatenate([x0, x1], axis=-1)
    xp = np.pad(x.astype(np.float64), ((0, 0), (1, 1), (1, 1), (0, 0)))
    ref = np.zeros((n_img, H, Wd, OC))
    w64 = Wt.astype(np.float64)
    for kh in range(3):
        for kw in range(3):
            ref += np.einsum("nhwc,oc->nhwo", xp[:, kh:kh + H, kw:kw + Wd, :], w64[:, kh, kw, :])
    assert np.abs(out - ref).max() < 3e-3 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("n_img,H,Wd,C0,C1,OC", [(3, 32, 32, 96, 0, 96), (1, 16, 16, 128, 128, 128), (5, 32, 32, 96, 96, 96), (2, 64, 32, 32, 0, 48)])
def test_conv3x3_pair_mode_gives_the_bits_of_the_other_schemes(G, n_img, H, Wd, C0, C1, OC, monkeypatch):
    """Per-tap boxes, the halo box and the halo box shared by two M tiles accumulate in the same K order: identical bits."""
    L = G.lib_ggml()
    rng = np.random.default_rng(H + C0 + OC)
    x0 = rng.normal(size=(n_img, H, Wd, C0)).astype(np.float16)
    x1 = rng.normal(size=(n_img, H, Wd, C1)).astype(np.float16) if C1 else None
    Wt = (rng.normal(size=(OC, 3, 3, C0 + C1)) / np.sqrt(9 * (C0 + C1))).astype(np.float16)
    outs = []
    for env in ({"GGML_B200_CONV_NO_HALO": "1"}, {"GGML_B200_CONV_NO_PAIR": "1"}, {"GGML_B200_CONV_PAIR": "1"}):
        for k in ("GGML_B200_CONV_NO_HALO", "GGML_B200_CONV_NO_PAIR", "GGML_B200_CONV_PAIR"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        out = np.zeros((n_img, H, Wd, OC), np.float32)
        assert L.ggml_b200_debug_conv3x3(_p(x0.view(np.uint16), u16p), C0, _p(x1.view(np.uint16), u16p) if C1 else None, C1, n_img,
                                         H, Wd, _p(Wt.view(np.uint16), u16p), OC, None, None, 0, _p(out, f32p)) == 0
        outs.append(out)
    np.testing.assert_array_equal(outs[0], outs[1])
    np.testing.assert_array_equal(outs[0], outs[2])


# ---- whole model through the ggml boundary ------------------------------------------------------------
def _run_model(G, path, imgs, mode):
    from ggml_experiments_b200 import mobilevit as MV
    MV.set_mode(mode)
    m = G.MobileViT(path)
    try:
        feat, pooled = m.extract_features(imgs)
        info = m.plan_info(*imgs.shape[:3])
    finally:
        m.close()
    return feat, pooled, info


@pytest.mark.parametrize("variant,n,hw", [("xxs", 1, 256), ("xxs", 3, 128), ("xs", 2, 256), ("s", 2, 256)])
def test_exact_mode_matches_oracle(G, oracle, weight_files, variant, n, hw):
    """EXACT mode: one f32-accurate kernel per ggml node with ggml's rounding points.  Early stages agree with the
    oracle to 1e-7; the f16 rounding points then amplify 1-ulp differences up to the ~1e-3 noise floor of the ggml
    semantics (DESIGN.md "the f16 noise floor"), so the gate here is the north_star element gate plus rel-L2."""
    from ggml_experiments_b200 import mobilevit as MV
    imgs = W.synthetic_images(n, hw, hw, seed=7)
    ref_f, ref_p = oracle.OracleModel(weight_files[variant]).forward(imgs)
    feat, pooled, info = _run_model(G, weight_files[variant], imgs, MV.EXACT)
    assert info["mode"] == MV.EXACT and info["launches"] > 100
    r = parity_report(feat, ref_f, rtol=1e-2, atol_rms=1e-2)
    print(variant, n, hw, r, info)
    assert r["violations"] == 0, r
    assert r["rel_l2"] < 2.5e-3, r
    assert np.abs(pooled - ref_p).max() < 1e-2
    assert top1_report(pooled, ref_p)["agree"] == 1.0
    # the liveness planner must beat "everything stays alive" (the reference's 1 GiB-per-image arena)
    assert info["arena_bytes"] < info["naive_bytes"] / 4


@pytest.mark.parametrize("variant,n,hw", [("xxs", 2, 256), ("s", 2, 256), ("xs", 1, 128)])
def test_exact_f32_validation_mode_max_abs_1e3(G, oracle, weight_files, variant, n, hw):
    """north_star's "max-abs 1e-3 in a TF32/f32 validation mode": with the activation rounding switched off on BOTH
    sides (weights keep their f16 values) the forward pass is a smooth function and the GPU must match the oracle to
    f32 accumulation noise.  This pins every kernel of the EXACT plan and the graph builder at 1e-3 absolute."""
    from ggml_experiments_b200 import mobilevit as MV
    imgs = W.synthetic_images(n, hw, hw, seed=7)
    ref_f, ref_p = oracle.OracleModel(weight_files[variant]).forward(imgs, oracle.NO_ACT_ROUND)
    feat, pooled, info = _run_model(G, weight_files[variant], imgs, MV.EXACT_F32)
    assert info["mode"] == MV.EXACT_F32
    r = parity_report(feat, ref_f, rtol=1e-3, atol_rms=1e-3)
    print(variant, n, hw, r)
    assert r["max_abs"] < 1e-3, r
    assert r["rel_l2"] < 2e-5, r
    assert np.abs(pooled - ref_p).max() < 1e-3


def test_stage_by_stage_exact_vs_oracle(G, oracle, weight_files):
    """Per-stage taps (stem, layer 1..5, exp): the first stages must agree with the oracle to accumulation noise."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "stage_debug.py"), "xxs", "1", "64", "exact"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("img 0")][:7]
    rel = [float(l.split("relL2")[1].split()[0]) for l in lines]
    print(rel)
    assert rel[0] < 1e-6 and rel[1] < 1e-6 and max(rel) < 3e-3


def test_batch_independence_exact(G, weight_files):
    """Row b of a batch-B run == the batch-1 run of image b."""
    from ggml_experiments_b200 import mobilevit as MV
    imgs = W.synthetic_images(4, 128, 128, seed=11)
    fb, pb, _ = _run_model(G, weight_files["xxs"], imgs, MV.EXACT)
    for b in (0, 3):
        f1, p1, _ = _run_model(G, weight_files["xxs"], imgs[b:b + 1], MV.EXACT)
        assert np.array_equal(f1[0], fb[b]) and np.array_equal(p1[0], pb[b])


@pytest.mark.parametrize("variant,n,hw", [("xxs", 2, 256), ("xs", 2, 256), ("s", 4, 256), ("s", 1, 512), ("xxs", 5, 128)])
def test_fast_mode_matches_oracle(G, oracle, weight_files, variant, n, hw):
    """FAST mode (fused tcgen05 plan): |d| <= 1e-2*|ref| + 1e-2*rms(ref), 100% top-1 (north_star tolerance)."""
    from ggml_experiments_b200 import mobilevit as MV
    imgs = W.synthetic_images(n, hw, hw, seed=7)
    ref_f, ref_p = oracle.OracleModel(weight_files[variant]).forward(imgs)
    feat, pooled, info = _run_model(G, weight_files[variant], imgs, MV.FAST)
    r = parity_report(feat, ref_f, rtol=1e-2, atol_rms=1e-2)
    t = top1_report(pooled, ref_p)
    print(variant, n, hw, r, t, info)
    if info["mode"] != MV.FAST:
        pytest.xfail("fused planner not yet covering this graph; exact plan was used")
    assert r["violations"] == 0, r  # north_star: EVERY element inside 1e-2*|ref| + 1e-2*rms
    assert r["rel_l2"] < 5e-3, r
    assert t["agree"] == 1.0, t


def test_unmodified_reference_main_runs_on_libggml_b200(G, oracle, weight_files, tmp_path):
    """The reference's own main.cpp (compiled untouched against our headers, linked with libggml_b200.so) classifies
    its built-in test image (main.cpp:680-688); the 10 values it prints (main.cpp:703) must match the oracle."""
    import os
    import re
    import shutil
    import subprocess
    exe = os.path.join(os.path.dirname(G.native_paths()["ggml"]), "ref_main_b200")
    if not os.path.exists(exe):
        pytest.skip("ref_main_b200 not built (needs /root/reference at build time)")
    shutil.copy(weight_files["s"], tmp_path / "weight.ggml")  # main.cpp:665 loads "weight.ggml" from the cwd
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=300, env=dict(os.environ, GGML_B200_VERBOSE="1"))
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert "output feature shape: : Dims: (8, 8, 640)" in r.stdout
    line = r.stdout.strip().splitlines()[-1]
    vals = [float(v) for v in re.findall(r"-?\d+\.?\d*(?:e-?\d+)?", line)]
    assert len(vals) == 10, line
    ref_f, _ = oracle.OracleModel(weight_files["s"]).forward(W.synthetic_images(1, 256, 256))
    ref = np.concatenate([ref_f[0, :5, 0, 0], ref_f[0, -5:, 0, 0]])
    print("reference main.cpp printed:", vals, "oracle:", ref.tolist(), r.stderr[-300:])
    # same element gate as the whole-model tests (the program prints 6 significant digits: +1e-5 relative)
    rms = float(np.sqrt((ref_f.astype(np.float64) ** 2).mean()))
    assert (np.abs(np.array(vals) - ref) <= 1e-2 * rms + 1e-2 * np.abs(ref) + 1e-5 * np.abs(ref)).all(), (vals, ref.tolist(), rms)


def test_unmodified_reference_rnn_runs_on_libggml_b200(G, tmp_path):
    """the reference's GRU text generator (rnn_text_generation.cpp, compiled untouched, linked with
    libggml_b200.so) runs its 200-step greedy loop on the GPU (EXACT plan); the generated token sequence must equal the
    numpy restatement of gru_forward wherever the greedy decision is not a numerical coin flip."""
    import os
    import subprocess
    from oracle import gru_oracle as GO
    exe = os.path.join(os.path.dirname(G.native_paths()["ggml"]), "ref_rnn_b200")
    if not os.path.exists(exe):
        pytest.skip("ref_rnn_b200 not built (needs /root/reference at build time)")
    w = GO.make_synthetic_gru(seed=5)
    os.makedirs(tmp_path / "rnn_text_gen")
    GO.write_gru_bin(str(tmp_path / "rnn_text_gen" / "gru.bin"), w)  # rnn.cpp:117 opens this relative path
    prompt = "ROMEO: what light"
    r = subprocess.run([exe], cwd=tmp_path, input=prompt + "\n", capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-800:], r.stderr[-800:])
    # print_text (rnn.cpp:92-96) prints the growing sequence followed by a separator line; take the last sequence
    blocks = r.stdout.split("\n--------\n")
    assert len(blocks) > 100, r.stdout[-500:]
    last = blocks[-2] if blocks[-1].strip() == "" else blocks[-1]
    ids, margins = GO.generate(w, prompt, steps=200)
    ref_text = "".join(GO.VOCAB[i] for i in ids)
    got = last[-len(ref_text):]
    # compare up to the first step whose greedy margin is within f32 noise
    n_cmp = len(ref_text)
    for i, mg in enumerate(margins[:len(ref_text)]):
        if mg < 1e-3:
            n_cmp = min(n_cmp, i)
            break
    assert n_cmp > len(prompt) + 20, margins[:40]
    assert got[:n_cmp] == ref_text[:n_cmp], (got[:80], ref_text[:80])
    print("GRU: %d of %d generated tokens compared equal; min margin %.3g" % (n_cmp, len(ref_text), min(margins)))


@pytest.mark.parametrize("hw", [64, 192, 320, 384, 448])
def test_fast_mode_at_resolutions_whose_maps_do_not_divide_128(G, oracle, weight_files, hw):
    """192 / 320 / 384 / 448 give 24 / 40 / 48 / 56-wide maps in the first ViT block: the fused plan must still cover them
    (384 and 448 also take the multi-chunk K/V path of the attention kernel: 576 / 784 keys); 64 is the smallest valid input."""
    from ggml_experiments_b200 import mobilevit as MV
    imgs = W.synthetic_images(2, hw, hw, seed=7)
    ref_f, ref_p = oracle.OracleModel(weight_files["xs"]).forward(imgs)
    feat, pooled, info = _run_model(G, weight_files["xs"], imgs, MV.FAST)
    assert info["mode"] == MV.FAST, info
    r = parity_report(feat, ref_f, rtol=1e-2, atol_rms=1e-2)
    assert r["violations"] == 0 and r["rel_l2"] < 5e-3, r
    assert top1_report(pooled, ref_p)["agree"] == 1.0


def test_non_square_and_odd_batch_fast_equals_exact(G, weight_files):
    """The reference is square-only (main.cpp:754); the batched builder is not.  For a non-square input and an odd batch the
    fused plan (tokens kept in NHWC pixel order, unfold/fold elided) must agree with the per-node EXACT plan, which executes
    the generalised unfold/fold permutations literally."""
    from ggml_experiments_b200 import mobilevit as MV
    imgs = W.synthetic_images(3, 128, 256, seed=3)  # widths must tile 128 for the fused 3x3 conv (else the EXACT plan runs)
    fe, pe, ie = _run_model(G, weight_files["xxs"], imgs, MV.EXACT)
    ff, pf, i_f = _run_model(G, weight_files["xxs"], imgs, MV.FAST)
    assert ie["mode"] == MV.EXACT and i_f["mode"] == MV.FAST
    r = parity_report(ff, fe, rtol=1e-2, atol_rms=1e-2)
    print(r)
    assert r["violations"] == 0 and r["rel_l2"] < 5e-3, r
    assert (pf.argmax(1) == pe.argmax(1)).all()


def test_batched_gru_matches_numpy_oracle(G, tmp_path):
    """BASELINE.json config 5 (row R1, batched): B independent GRU streams through the ggml boundary, loop on the device
    (argmax + state feedback).  Every greedy token must equal the numpy restatement up to the first decision whose logit
    margin is within f32 accumulation noise; the final state of never-diverged streams must match to 1e-4."""
    from ggml_experiments_b200.gru import GRU
    from ggml_experiments_b200 import mobilevit as MV
    from oracle import gru_oracle as GO
    w = GO.make_synthetic_gru(seed=5)
    path = str(tmp_path / "gru.bin")
    GO.write_gru_bin(path, w)
    B, steps = 48, 40
    first = (np.arange(B) * 7 % 66).astype(np.int32)
    MV.set_mode(MV.EXACT)  # f32 matmuls, like the reference's F32 x F32 ggml_mul_mat
    m = GRU(path)
    toks, state, ms = m.generate(first, steps)
    m.close()
    MV.set_mode(MV.FAST)
    ref, margins, ref_state = GO.generate_batch(w, first, steps)
    n_equal, n_compared, clean = 0, 0, []
    for b in range(B):
        risky = np.nonzero(margins[:, b] < 1e-3)[0]
        upto = int(risky[0]) if len(risky) else steps
        n_compared += upto
        n_equal += int((toks[:upto, b] == ref[:upto, b]).sum())
        if upto == steps:
            clean.append(b)
    print(f"batched GRU: {n_equal}/{n_compared} tokens equal, {len(clean)}/{B} streams without a coin-flip step, {ms:.2f} ms for {steps} steps")
    assert n_compared > 0.8 * B * steps and n_equal == n_compared
    assert len(clean) > B // 2
    assert np.abs(state[clean] - ref_state[clean]).max() < 1e-4


def test_batched_gru_fast_mode_tensor_cores(G, tmp_path):
    """FAST mode lowers the three dense layers of the cell to the tcgen05 GEMM (f16 operands, f32 accumulate).  One step from
    identical inputs must match the f32 oracle to f16-operand accuracy, and greedy tokens agree wherever the margin is clear."""
    from ggml_experiments_b200.gru import GRU
    from ggml_experiments_b200 import mobilevit as MV
    from oracle import gru_oracle as GO
    w = GO.make_synthetic_gru(seed=5)
    path = str(tmp_path / "gru.bin")
    GO.write_gru_bin(path, w)
    B, steps = 128, 12
    first = (np.arange(B) * 5 % 66).astype(np.int32)
    MV.set_mode(MV.FAST)
    m = GRU(path)
    toks, state, ms = m.generate(first, steps)
    m.close()
    ref, margins, ref_state = GO.generate_batch(w, first, steps)
    # step 0 starts from identical inputs: compare tokens where the f32 margin is clear of f16-operand noise
    clear0 = margins[0] > 2e-2
    assert clear0.sum() > B // 2 and (toks[0][clear0] == ref[0][clear0]).all()
    same = (toks == ref).all(axis=0)
    print(f"GRU fast: {same.sum()}/{B} streams identical to the f32 oracle over {steps} steps")
    assert same.sum() > B // 2
    err = np.abs(state[same] - ref_state[same]).max()
    assert err < 2e-2, err


# ---- classifier head --------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode_name", ["fast", "exact"])
def test_classifier_head_matches_oracle_and_hf(G, oracle, weight_files, mode_name):
    import os
    from ggml_experiments_b200 import mobilevit as MV
    mode = MV.FAST if mode_name == "fast" else MV.EXACT
    imgs = W.synthetic_images(2, 256, 256, seed=7)
    om = oracle.OracleModel(weight_files["xxs_cls"])
    _, ref_p = om.forward(imgs)
    ref_logits = om.classify(ref_p)
    MV.set_mode(mode)
    m = G.MobileViT(weight_files["xxs_cls"])
    try:
        logits, top1 = m.classify(imgs)
        info = m.plan_info(2, 256, 256)
        feat, pooled = m.extract_features(imgs)
        # the head itself is f32-exact: applied by the oracle to OUR pooled features it must reproduce our logits
        own = om.classify(pooled)
        assert np.abs(own - logits).max() < 1e-4 * np.abs(own).max()
    finally:
        m.close()
        MV.set_mode(MV.FAST)
    assert info["mode"] == mode, info  # the fused planner must cover the graph with the head attached
    assert np.abs(logits - ref_logits).max() < 1e-2 * np.abs(ref_logits).max()
    assert (top1 == ref_logits.argmax(1)).all()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hf_cls_xxs_256.npz"))
    assert (top1 == g["logits"].argmax(1)).all()


# ---- u8 images, preprocessing on the device ---------------------------------------------------------------
@pytest.mark.parametrize("sh,sw,hw", [(256, 256, 256), (300, 400, 256), (480, 270, 256), (100, 80, 256), (720, 1280, 256), (333, 333, 128)])
def test_device_preprocess_u8_is_bit_identical_to_the_oracle(G, oracle, weight_files, sh, sw, hw):
    img = np.random.default_rng(sh + 3 * sw).integers(0, 256, (3, sh, sw, 3), dtype=np.uint8)
    m = G.MobileViT(weight_files["xxs"])
    try:
        got = m.preprocess_u8(img, hw, hw)
    finally:
        m.close()
    np.testing.assert_array_equal(got, oracle.preprocess_u8(img, hw, hw))


def test_compute_u8_equals_f32_path_on_preprocessed_images(G, oracle, weight_files):
    """H2D(u8) + device preprocess + forward == forward on the oracle-preprocessed f32 images, bit for bit; also pipelined."""
    n, hw, sh, sw = 4, 256, 360, 480
    img = np.random.default_rng(11).integers(0, 256, (n, sh, sw, 3), dtype=np.uint8)
    pre = oracle.preprocess_u8(img, hw, hw)
    m = G.MobileViT(weight_files["xs"])
    try:
        f_ref, p_ref = m.extract_features(pre)
        m.host_input_u8(n, hw, hw, sh, sw)[:] = img
        f, p = m.compute_u8(n, hw, hw, sh, sw)
        np.testing.assert_array_equal(p, p_ref)
        np.testing.assert_array_equal(f, f_ref)
        for s in range(2):
            m.slot_input_u8(n, hw, hw, s, sh, sw)[:] = img
            m.slot_submit_u8(n, hw, hw, s, sh, sw)
        for s in range(2):
            fs, ps = m.slot_wait(n, hw, hw, s)
            np.testing.assert_array_equal(ps, p_ref)
    finally:
        m.close()


@pytest.mark.parametrize("variant,n,hw,sh,sw", [("xxs", 3, 256, 256, 256), ("s", 2, 128, 128, 128), ("xs", 2, 192, 192, 192), ("xxs", 1, 320, 500, 320),
                                               ("xxs", 5, 64, 64, 64), ("s", 2, 448, 448, 448)])
def test_u8_images_staged_by_the_stem_equal_the_f32_route(G, oracle, weight_files, variant, n, hw, sh, sw):
    """FAST plan: the stem stages its patch from the quantised u8 image (no f32 image in HBM; same-size images skip the resize kernel).
    Same bits as the forward on the oracle-preprocessed f32 images, for ragged stem tiles, every batch position, and the plan must
    read the f32 leaf again on the next f32 call (the u8 flag is armed for one compute only)."""
    rng = np.random.default_rng(n * 1000 + hw)
    img = rng.integers(0, 256, (n, sh, sw, 3), dtype=np.uint8)
    pre = oracle.preprocess_u8(img, hw, hw)
    other = rng.random((n, hw, hw, 3), dtype=np.float32)
    m = G.MobileViT(weight_files[variant])
    try:
        f_ref, p_ref = (a.copy() for a in m.extract_features(pre))
        f_oth, p_oth = (a.copy() for a in m.extract_features(other))
        assert not np.array_equal(p_ref, p_oth)
        m.host_input_u8(n, hw, hw, sh, sw)[:] = img
        f, p = (a.copy() for a in m.compute_u8(n, hw, hw, sh, sw))
        np.testing.assert_array_equal(p, p_ref)
        np.testing.assert_array_equal(f, f_ref)
        f2, p2 = m.extract_features(other)  # f32 route on the same plan right after a u8 compute
        np.testing.assert_array_equal(p2, p_oth)
        np.testing.assert_array_equal(f2, f_oth)
    finally:
        m.close()


# ---- BASELINE.json's full size (MobileViT-S, batch 256, 256x256) through size-independent properties --------------------
def test_full_size_batch_256_is_batch_independent_and_matches_the_oracle_sample(G, oracle, weight_files):
    """At the bench's own shape the oracle would need minutes, so: (1) every copy of an image inside the batch of 256 gives
    bit-identical features (tiles, CTAs and batch position must not leak into the result), (2) the same 8 images run as a batch
    of 8 give the same bits as inside the batch of 256, (3) those 8 are checked against the oracle with the north_star gate."""
    from ggml_experiments_b200 import mobilevit as MV
    base = W.synthetic_images(8, 256, 256, seed=7)
    imgs = np.tile(base, (32, 1, 1, 1))
    MV.set_mode(MV.FAST)
    m = G.MobileViT(weight_files["s"])
    try:
        feat, pooled = m.extract_features(imgs)
        info = m.plan_info(256, 256, 256)
        feat8, pooled8 = m.extract_features(base)
    finally:
        m.close()
    assert info["mode"] == MV.FAST
    f = feat.reshape(32, 8, *feat.shape[1:])
    assert np.array_equal(f, np.broadcast_to(f[0], f.shape)), "copies of the same image differ inside the batch"
    np.testing.assert_array_equal(feat8, f[0])
    np.testing.assert_array_equal(pooled8, pooled[:8])
    ref_f, ref_p = oracle.OracleModel(weight_files["s"]).forward(base)
    r = parity_report(feat8, ref_f, rtol=1e-2, atol_rms=1e-2)
    t = top1_report(pooled8, ref_p)
    assert r["violations"] == 0 and r["rel_l2"] < 5e-3, r
    assert t["agree"] == 1.0, t


@pytest.mark.parametrize("variant,hw", [("s", 256), ("xxs", 128), ("xs", 192)])
def test_every_batch_size_gives_the_same_bits_per_image(G, weight_files, variant, hw):
    """Tail handling of every kernel (partial M tiles, partial depthwise tiles, odd image counts per conv tile, attention grids):
    for batch sizes that are not multiples of anything, image i must come out bit-identical to the same image in a batch of 8."""
    from ggml_experiments_b200 import mobilevit as MV
    base = W.synthetic_images(8, hw, hw, seed=7)
    MV.set_mode(MV.FAST)
    m = G.MobileViT(weight_files[variant])
    try:
        feat8, pooled8 = m.extract_features(base)
        for b in (1, 2, 3, 5, 7, 9, 33, 100, 255, 257):
            imgs = base[np.arange(b) % 8]
            feat, pooled = m.extract_features(imgs)
            assert m.plan_info(b, hw, hw)["mode"] == MV.FAST
            np.testing.assert_array_equal(feat, feat8[np.arange(b) % 8], err_msg=f"batch {b}")
            np.testing.assert_array_equal(pooled, pooled8[np.arange(b) % 8], err_msg=f"batch {b}")
            m.release(b, hw, hw)
    finally:
        m.close()


# ---- K4: fused inverted residual inside the whole model (opt-in: GGML_B200_IR_FUSE=1) ---------------------------------
@pytest.mark.parametrize("variant,n,hw", [("s", 3, 256), ("xs", 2, 256), ("xxs", 5, 128), ("s", 1, 512)])
def test_fast_mode_with_fused_inverted_residuals_matches_oracle(G, oracle, weight_files, variant, n, hw, monkeypatch):
    """The seven inverted-residual blocks (main.cpp:854-870) run as ONE kernel each (expand -> depthwise -> reduce on chip, ir_fused.cu):
    same north_star gate as the default plan, and fewer launches."""
    from ggml_experiments_b200 import mobilevit as MV
    imgs = W.synthetic_images(n, hw, hw, seed=7)
    ref_f, ref_p = oracle.OracleModel(weight_files[variant]).forward(imgs)
    _, _, info0 = _run_model(G, weight_files[variant], imgs, MV.FAST)
    monkeypatch.setenv("GGML_B200_IR_FUSE", "1")
    feat, pooled, info = _run_model(G, weight_files[variant], imgs, MV.FAST)
    monkeypatch.delenv("GGML_B200_IR_FUSE")
    assert info["mode"] == MV.FAST and info["launches"] <= info0["launches"] - 10, (info0, info)
    r = parity_report(feat, ref_f, rtol=1e-2, atol_rms=1e-2)
    print(variant, n, hw, r, info)
    assert r["violations"] == 0 and r["rel_l2"] < 5e-3, r
    assert top1_report(pooled, ref_p)["agree"] == 1.0


def test_dwreduce_fusion_matches_oracle(G, oracle, weight_files, monkeypatch):
    """K4a (depthwise + reduce in one kernel, dwreduce.cu; opt-in GGML_B200_DWREDUCE=1) keeps the same gate."""
    from ggml_experiments_b200 import mobilevit as MV
    imgs = W.synthetic_images(2, 256, 256, seed=7)
    ref_f, ref_p = oracle.OracleModel(weight_files["s"]).forward(imgs)
    monkeypatch.setenv("GGML_B200_DWREDUCE", "1")
    feat, pooled, info = _run_model(G, weight_files["s"], imgs, MV.FAST)
    monkeypatch.delenv("GGML_B200_DWREDUCE")
    assert info["mode"] == MV.FAST
    r = parity_report(feat, ref_f, rtol=1e-2, atol_rms=1e-2)
    assert r["violations"] == 0 and r["rel_l2"] < 5e-3, r
    assert top1_report(pooled, ref_p)["agree"] == 1.0


def test_f16_on_disk_and_pretransposed_weight_files_give_identical_features(G, weight_files, tmp_path):
    """the lossless file flavours (f16 convolution kernels, dense kernels stored (out, in)) load to the same model bit for bit."""
    from ggml_experiments_b200 import mobilevit as MV
    imgs = W.synthetic_images(3, 128, 128, seed=7)
    MV.set_mode(MV.FAST)
    f0, p0, _ = _run_model(G, weight_files["xs"], imgs, MV.FAST)
    t = W.read_weight_file(weight_files["xs"])
    for kw in ({"f16": "conv"}, {"pretransposed": True}, {"f16": "conv", "pretransposed": True}):
        p = str(tmp_path / "w.ggml")
        W.write_weight_file(p, t, **kw)
        f, pp, info = _run_model(G, p, imgs, MV.FAST)
        assert info["mode"] == MV.FAST
        np.testing.assert_array_equal(f, f0, err_msg=str(kw))
        np.testing.assert_array_equal(pp, p0, err_msg=str(kw))


def test_concurrent_lanes_give_the_same_bits(G, weight_files, monkeypatch):
    """MVIT_LANES=S: one request runs as S sub-batches on S streams (host/mobilevit.cpp graph_for); inputs and outputs of the lanes alias
    slices of the caller's buffers.  Synchronous, u8 and pipelined-slot entry points must return exactly the unsplit result."""
    base = W.synthetic_images(12, 128, 128, seed=7)
    u8 = np.clip(np.rint(base * 255.0), 0, 255).astype(np.uint8)
    m = G.MobileViT(weight_files["xs"])
    try:
        f0, p0 = m.extract_features(base)
        m.host_input_u8(12, 128, 128, 128, 128)[:] = u8
        fu0, pu0 = (a.copy() for a in m.compute_u8(12, 128, 128, 128, 128))
        assert m.plan_info(12, 128, 128)["lanes"] == 1
    finally:
        m.close()
    for lanes in (2, 3, 4):
        monkeypatch.setenv("MVIT_LANES", str(lanes))
        m = G.MobileViT(weight_files["xs"])
        try:
            f, p = m.extract_features(base)
            info = m.plan_info(12, 128, 128)
            assert info["lanes"] == lanes and info["mode"] == 0 and info["cuda_graph"] == 1
            np.testing.assert_array_equal(f, f0)
            np.testing.assert_array_equal(p, p0)
            m.host_input_u8(12, 128, 128, 128, 128)[:] = u8
            fu, pu = m.compute_u8(12, 128, 128, 128, 128)
            np.testing.assert_array_equal(fu, fu0)
            np.testing.assert_array_equal(pu, pu0)
            for s in range(2):
                m.slot_input(12, 128, 128, s)[:] = base
                m.slot_submit(12, 128, 128, s)
            for s in range(2):
                fs, ps = m.slot_wait(12, 128, 128, s)
                np.testing.assert_array_equal(fs, f0)
                np.testing.assert_array_equal(ps, p0)
        finally:
            m.close()
        monkeypatch.delenv("MVIT_LANES")
