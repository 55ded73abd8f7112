"""CPU tests of the drop-in boundary: the native libraries load and export every symbol include/*.h declares,
the loader parses the reference's weight-file layout, and compute fails loudly without a GPU (no fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header: str):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"#define GGML_ASSERT.*?while \(0\)", "", src, flags=re.S)
    names = re.findall(r"\b((?:ggml|mvit)_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_libraries_export_every_declared_symbol():
    import ggml_experiments_b200 as G
    g, m = G.lib_ggml(), G.lib_mobilevit()
    fg = _declared_functions("ggml/ggml.h")
    fm = _declared_functions("mobilevit_b200.h")
    assert len(fg) > 60 and len(fm) >= 12
    for name in fg:
        assert hasattr(g, name), f"libggml_b200.so lacks {name}"
    for name in fm:
        assert hasattr(m, name), f"libmobilevit_b200.so lacks {name}"


def test_fp16_conversion_is_ieee_rne():
    import ggml_experiments_b200 as G
    L = G.lib_ggml()
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.normal(0, 1, 4000), rng.normal(0, 1e-5, 1000), rng.normal(0, 3e4, 1000),
                         [0.0, -0.0, 65504.0, 65519.9, 65520.0, 6e-8, 2.98e-8, 2.99e-8, 1e-8]]).astype(np.float32)
    with np.errstate(over="ignore"):
        ref = xs.astype(np.float16).view(np.uint16)
    for v, r in zip(xs, ref):
        assert L.ggml_fp32_to_fp16(float(v)) == int(r), v
    for h in list(range(0, 0x7c00, 97)) + [0x8001, 0x83ff, 0xfbff, 0x7c00]:
        back = L.ggml_fp16_to_fp32(h)
        assert np.float32(back) == np.uint16(h).view(np.float16).astype(np.float32)


def test_loader_reads_reference_layout(weight_files):
    import ggml_experiments_b200 as G
    expect = {"s": (4949888, 640), "xs": (1941296, 384), "xxs": (955136, 320)}
    for v, (n, oc) in expect.items():
        m = G.MobileViT(weight_files[v])
        assert (m.num_tensors, m.num_weights, m.out_channels) == (313, n, oc)
        m.close()
    with pytest.raises(FileNotFoundError):
        G.MobileViT("/nonexistent/weight.ggml")


def test_loader_reads_optional_classifier(weight_files):
    """a file with `classifier/{kernel,bias}:0` gets a head; the reference's 313-tensor file does not."""
    import ggml_experiments_b200 as G
    m = G.MobileViT(weight_files["xxs_cls"])
    assert (m.num_tensors, m.num_classes, m.out_channels) == (315, 1000, 320)
    m.close()
    m = G.MobileViT(weight_files["xxs"])
    assert m.num_classes == 0
    with pytest.raises(ValueError, match="no classifier"):
        m.classify(np.zeros((1, 64, 64, 3), np.float32))
    m.close()


def test_extract_features_rejects_bad_shapes(weight_files):
    import ggml_experiments_b200 as G
    m = G.MobileViT(weight_files["xxs"])
    with pytest.raises(ValueError):
        m.extract_features(np.zeros((1, 96, 96, 3), np.float32))  # not a multiple of 64


def test_compute_without_gpu_fails_loudly(weight_files):
    """The product path must not silently fall back to a CPU implementation."""
    import ggml_experiments_b200 as G
    if G.lib_ggml().ggml_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    code = ("import sys; sys.path.insert(0, %r); import numpy as np; import ggml_experiments_b200 as G; "
            "m = G.MobileViT(%r); m.extract_features(np.zeros((1, 64, 64, 3), np.float32))" % (ROOT, weight_files["xxs"]))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no usable CUDA device" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.skipif(not os.path.exists("/root/reference/mobilevit/main.cpp"), reason="reference sources not present")
def test_unmodified_reference_programs_compile_and_link_against_the_boundary(tmp_path):
    """mobilevit/main.cpp and rnn_text_gen/rnn_text_generation.cpp, untouched, build against include/ + libggml_b200.so."""
    import ggml_experiments_b200 as G
    lib_dir = os.path.dirname(G.native_paths()["ggml"])
    inc = os.path.join(ROOT, "include")
    for src, extra in (("/root/reference/mobilevit/main.cpp", ["-I/root/reference"]),
                       ("/root/reference/rnn_text_gen/rnn_text_generation.cpp", [])):
        exe = str(tmp_path / (os.path.basename(src) + ".bin"))
        r = subprocess.run(["g++", "-O0", "-std=c++17", "-w", "-I" + inc, *extra, src, "-o", exe, "-L" + lib_dir, "-lggml_b200"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]


def test_product_never_imports_or_links_the_oracle():
    """oracle/ is test infrastructure: nothing under the package or include/ may import, call or link it."""
    pkg = os.path.join(ROOT, "ggml-experiments_b200")
    banned = ("import oracle", "from oracle", "mvo_", "libmvit_oracle", "oracle/")
    for top in (pkg, os.path.join(ROOT, "include")):
        for base, _, files in os.walk(top):
            if "_build" in base or "__pycache__" in base:
                continue
            for fn in files:
                if fn.endswith((".py", ".cpp", ".cu", ".h", ".cuh")):
                    text = open(os.path.join(base, fn)).read()
                    for tok in banned:
                        assert tok not in text, f"{fn} references the oracle ({tok})"
    out = subprocess.run(["ldd", os.path.join(pkg, "_build", "libggml_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
