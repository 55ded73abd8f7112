"""Weight-file tooling for the `weight.ggml` layout of the reference.

Layout (reference: mobilevit/convert-tf-to-ggml.py:16-33, read back by mobilevit/main.cpp:872-942)::

    repeat { int32 name_len; char name[name_len]; int32 n_dims; int32 dims[n_dims];  # TF order
             float32 data[prod(dims)] }

Tensor names are the TF variable paths of `TFMobileViTModel`.  The reference script
needs TensorFlow and a download of apple/mobilevit-small; neither exists here, so this module writes
random-init weights of the same architecture and names (numpy only).  No compute happens here.
"""
from __future__ import annotations

import struct
from collections import OrderedDict

import numpy as np

P = "tf_mobile_vi_t_model/mobilevit"

# SURVEY.md 8(d) / App. A.  `stages` = transformer layers per ViT block, MobileNet stage counts are 1 and 3.
VARIANTS = {
    "s": dict(hidden=(144, 192, 240), neck=(16, 32, 64, 96, 128, 160, 640), expand=4.0),
    "xs": dict(hidden=(96, 120, 144), neck=(16, 32, 48, 64, 80, 96, 384), expand=4.0),
    "xxs": dict(hidden=(64, 80, 96), neck=(16, 16, 24, 48, 64, 80, 320), expand=2.0),
}
VIT_STAGES = (2, 4, 3)
MLP_RATIO = 2.0
NUM_HEADS = 4


def make_divisible(value: float, divisor: int = 8) -> int:
    new_value = max(divisor, int(value + divisor / 2) // divisor * divisor)
    if new_value < 0.9 * value:
        new_value += divisor
    return int(new_value)


def _conv(rng, out, path, kh, kw, ic, oc, norm=True, depthwise=False):
    # conv N(0, 2/fan_in); BN gamma U(.75,1.25), beta N(0,.1^2), mean N(0,.1^2), var U(.75,1.25)
    fan_in = kh * kw * (1 if depthwise else ic)
    shape = (kh, kw, 1, oc) if depthwise else (kh, kw, ic, oc)
    out[f"{path}/convolution/kernel:0"] = rng.normal(0.0, np.sqrt(2.0 / fan_in), shape).astype(np.float32)
    if norm:
        out[f"{path}/normalization/gamma:0"] = rng.uniform(0.75, 1.25, (oc,)).astype(np.float32)
        out[f"{path}/normalization/beta:0"] = rng.normal(0.0, 0.1, (oc,)).astype(np.float32)
        out[f"{path}/normalization/moving_mean:0"] = rng.normal(0.0, 0.1, (oc,)).astype(np.float32)
        out[f"{path}/normalization/moving_variance:0"] = rng.uniform(0.75, 1.25, (oc,)).astype(np.float32)


def _inverted_residual(rng, out, path, cin, cout, expand):
    e = make_divisible(int(round(cin * expand)), 8)
    _conv(rng, out, f"{path}/expand_1x1", 1, 1, cin, e)
    _conv(rng, out, f"{path}/conv_3x3", 3, 3, e, e, depthwise=True)
    _conv(rng, out, f"{path}/reduce_1x1", 1, 1, e, cout)


def _dense(rng, out, path, cin, cout):
    out[f"{path}/kernel:0"] = rng.normal(0.0, np.sqrt(1.0 / cin), (cin, cout)).astype(np.float32)
    out[f"{path}/bias:0"] = rng.normal(0.0, 0.02, (cout,)).astype(np.float32)


def _ln(rng, out, path, c):
    out[f"{path}/gamma:0"] = rng.uniform(0.75, 1.25, (c,)).astype(np.float32)
    out[f"{path}/beta:0"] = rng.normal(0.0, 0.1, (c,)).astype(np.float32)


CLASSIFIER = "tf_mobile_vi_t_for_image_classification/classifier"  # TFMobileViTForImageClassification's head variables


def make_synthetic_weights(variant: str = "s", seed: int = 1234, num_classes: int = 0) -> "OrderedDict[str, np.ndarray]":
    """Random-init MobileViT weights under the reference's tensor names (313 tensors for every variant).
    num_classes > 0 appends the classification head (`classifier/kernel:0` (C, classes), `classifier/bias:0`),
    drawn after everything else so the 313 backbone tensors do not depend on it."""
    cfg = VARIANTS[variant]
    neck, hidden, expand = cfg["neck"], cfg["hidden"], cfg["expand"]
    rng = np.random.default_rng(seed)
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    _conv(rng, out, f"{P}/conv_stem", 3, 3, 3, neck[0])
    # layer.0: one inverted residual, stride 1; layer.1: three, first stride 2 (main.cpp:334-391)
    _inverted_residual(rng, out, f"{P}/encoder/layer.0/layer.0", neck[0], neck[1], expand)
    cin = neck[1]
    for i in range(3):
        _inverted_residual(rng, out, f"{P}/encoder/layer.1/layer.{i}", cin, neck[2], expand)
        cin = neck[2]
    for v in range(3):  # main.cpp:393-503
        base = f"{P}/encoder/layer.{v + 2}"
        cin, cout, d = neck[2 + v], neck[3 + v], hidden[v]
        _inverted_residual(rng, out, f"{base}/downsampling_layer", cin, cout, expand)
        _conv(rng, out, f"{base}/conv_kxk", 3, 3, cout, cout)
        _conv(rng, out, f"{base}/conv_1x1", 1, 1, cout, d, norm=False)
        f = int(d * MLP_RATIO)
        for j in range(VIT_STAGES[v]):
            tb = f"{base}/transformer/layer.{j}"
            _dense(rng, out, f"{tb}/attention/attention/query", d, d)
            _dense(rng, out, f"{tb}/attention/attention/key", d, d)
            _dense(rng, out, f"{tb}/attention/attention/value", d, d)
            _dense(rng, out, f"{tb}/attention/output/dense", d, d)
            _dense(rng, out, f"{tb}/intermediate/dense", d, f)
            _dense(rng, out, f"{tb}/output/dense", f, d)
            _ln(rng, out, f"{tb}/layernorm_before", d)
            _ln(rng, out, f"{tb}/layernorm_after", d)
        _ln(rng, out, f"{base}/layernorm", d)
        _conv(rng, out, f"{base}/conv_projection", 1, 1, d, cout)
        _conv(rng, out, f"{base}/fusion", 3, 3, 2 * cout, cout)
    _conv(rng, out, f"{P}/conv_1x1_exp", 1, 1, neck[5], neck[6])
    if num_classes > 0:
        _dense(rng, out, CLASSIFIER, neck[6], num_classes)
    return out


# Record-format extension (the two TODOs of convert-tf-to-ggml.py:13-14).  The upper half of the n_dims word carries flags; a file
# without flags is exactly the reference layout and is what the unmodified main.cpp reads.
FLAG_F16 = 1 << 16            # payload is IEEE f16 instead of f32 ("f16 on disk")
FLAG_PRETRANSPOSED = 1 << 17  # 2-D dense kernel stored as (out, in): the in-graph cont(permute(w)) of main.cpp:1024 is already applied


def write_weight_file(path: str, tensors: "OrderedDict[str, np.ndarray]", f16: str = "none", pretransposed: bool = False) -> int:
    """Write tensors in the convert-tf-to-ggml.py record layout.  Returns the number of elements written.

    f16 = "none": the reference layout, f32 payloads (default).
    f16 = "conv": convolution kernels -- the tensors main.cpp:887-889,928-932 rounds to f16 at load time anyway -- are stored as f16:
                  the loaded model is bit-identical and the file is 23-40 % smaller (XXS ... S).
    f16 = "all":  every tensor as f16 (dense kernels, biases and norms lose precision: opt-in).
    pretransposed: 2-D dense kernels are stored (out, in), the layout the matmul consumes (main.cpp:1022-1035 transposes them in-graph)."""
    assert f16 in ("none", "conv", "all")
    total = 0
    with open(path, "wb") as f:
        for name, arr in tensors.items():
            arr = np.ascontiguousarray(arr, dtype=np.float32)
            flags = 0
            if f16 == "all" or (f16 == "conv" and "convolution" in name):
                flags |= FLAG_F16
            if pretransposed and arr.ndim == 2 and name.endswith("/kernel:0"):
                flags |= FLAG_PRETRANSPOSED
                arr = np.ascontiguousarray(arr.T)
            nb = name.encode("utf-8")
            f.write(struct.pack("i", len(nb)))
            f.write(nb)
            f.write(struct.pack("i", arr.ndim | flags))
            for d in arr.shape:
                f.write(struct.pack("i", int(d)))
            f.write(arr.astype(np.float16).tobytes() if flags & FLAG_F16 else arr.tobytes())
            total += arr.size
    return total


def read_weight_file(path: str) -> "OrderedDict[str, np.ndarray]":
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    with open(path, "rb") as f:
        while True:
            head = f.read(4)
            if len(head) < 4:
                break
            (n,) = struct.unpack("i", head)
            name = f.read(n).decode("utf-8")
            (nd,) = struct.unpack("i", f.read(4))
            flags, nd = nd & ~0xFFFF, nd & 0xFFFF
            dims = struct.unpack("i" * nd, f.read(4 * nd))
            cnt = int(np.prod(dims))
            if flags & FLAG_F16:
                arr = np.frombuffer(f.read(2 * cnt), dtype=np.float16).astype(np.float32).reshape(dims)
            else:
                arr = np.frombuffer(f.read(4 * cnt), dtype=np.float32).reshape(dims).copy()
            if flags & FLAG_PRETRANSPOSED:
                arr = np.ascontiguousarray(arr.T)  # back to the canonical (in, out)
            out[name] = arr
    return out


# ---- Hugging Face torch MobileViTModel <-> file names, used by the golden generator ----

def hf_config_kwargs(variant: str) -> dict:
    cfg = VARIANTS[variant]
    return dict(hidden_sizes=list(cfg["hidden"]), neck_hidden_sizes=list(cfg["neck"]), expand_ratio=cfg["expand"],
                num_attention_heads=NUM_HEADS, mlp_ratio=MLP_RATIO, hidden_dropout_prob=0.0,
                attention_probs_dropout_prob=0.0, classifier_dropout_prob=0.0)


def to_hf_state_dict(tensors: "OrderedDict[str, np.ndarray]") -> dict:
    """Map file tensors to torch MobileViTModel state-dict entries (numpy arrays)."""
    sd = {}
    for name, arr in tensors.items():
        if name.startswith(CLASSIFIER + "/"):  # MobileViTForImageClassification.classifier (kept outside the `mobilevit.` prefix)
            sd["classifier." + ("weight" if name.endswith("kernel:0") else "bias")] = np.ascontiguousarray(arr.T) if arr.ndim == 2 else arr
            continue
        assert name.startswith(P + "/") and name.endswith(":0")
        parts = name[len(P) + 1:-2].split("/")
        leaf = parts[-1]
        if parts[-2] == "convolution":
            key = ".".join(parts[:-2]) + ".convolution.weight"
            sd[key] = np.ascontiguousarray(arr.transpose(3, 2, 0, 1))  # (KH,KW,IC,OC) -> (OC,IC,KH,KW)
        elif parts[-2] == "normalization":
            m = {"gamma": "weight", "beta": "bias", "moving_mean": "running_mean", "moving_variance": "running_var"}
            sd[".".join(parts[:-2]) + ".normalization." + m[leaf]] = arr
        elif leaf == "kernel":
            sd[".".join(parts[:-1]) + ".weight"] = np.ascontiguousarray(arr.T)
        elif leaf == "bias":
            sd[".".join(parts[:-1]) + ".bias"] = arr
        elif leaf == "gamma":
            sd[".".join(parts[:-1]) + ".weight"] = arr
        elif leaf == "beta":
            sd[".".join(parts[:-1]) + ".bias"] = arr
        else:
            raise KeyError(name)
    return sd


def from_hf_state_dict(sd: dict) -> "OrderedDict[str, np.ndarray]":
    """torch-HF MobileViT state dict -> file tensors: the exporter that replaces the TensorFlow-only
    convert-tf-to-ggml.py (convert.py:7-33 needs TF and a hub download).  Accepts MobileViTModel keys and
    MobileViTForImageClassification keys (`mobilevit.` prefix + `classifier.*`); values are numpy arrays or torch tensors.
    Inverse of to_hf_state_dict: conv (OC,IC,KH,KW) -> (KH,KW,IC,OC); Linear (out,in) -> (in,out); BN / LN names as App. B."""
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    bn = {"weight": "gamma", "bias": "beta", "running_mean": "moving_mean", "running_var": "moving_variance"}
    for key, val in sd.items():
        arr = np.asarray(val.detach().cpu().numpy() if hasattr(val, "detach") else val, dtype=np.float32)
        if key.endswith("num_batches_tracked"):
            continue
        if key.startswith("mobilevit."):
            key = key[len("mobilevit."):]
        parts = key.split(".")
        leaf = parts[-1]
        if parts[0] == "classifier":
            out[f"{CLASSIFIER}/{'kernel' if leaf == 'weight' else 'bias'}:0"] = np.ascontiguousarray(arr.T) if arr.ndim == 2 else arr
            continue
        # HF module path "encoder.layer.2.transformer.layer.0" -> file path "encoder/layer.2/transformer/layer.0"
        path, i = [], 0
        mods = parts[:-1]
        while i < len(mods):
            if mods[i] == "layer" and i + 1 < len(mods) and mods[i + 1].isdigit():
                path.append(f"layer.{mods[i + 1]}")
                i += 2
            else:
                path.append(mods[i])
                i += 1
        base = P + "/" + "/".join(path)
        if path[-1] == "convolution":
            out[f"{base}/kernel:0"] = np.ascontiguousarray(arr.transpose(2, 3, 1, 0))
        elif path[-1] == "normalization":
            out[f"{base}/{bn[leaf]}:0"] = arr
        elif path[-1].startswith("layernorm"):
            out[f"{base}/{'gamma' if leaf == 'weight' else 'beta'}:0"] = arr
        elif leaf == "weight":
            out[f"{base}/kernel:0"] = np.ascontiguousarray(arr.T)
        elif leaf == "bias":
            out[f"{base}/bias:0"] = arr
        else:
            raise KeyError(key)
    return out


def export_hf_checkpoint(model_dir: str, out_path: str, f16: str = "none", pretransposed: bool = False) -> int:
    """Write `weight.ggml` from a local Hugging Face MobileViT checkpoint directory (no network).  Returns the float count."""
    from transformers import AutoModelForImageClassification, MobileViTModel  # imported lazily: tooling, not the product path
    try:
        model = AutoModelForImageClassification.from_pretrained(model_dir, local_files_only=True)
    except Exception:
        model = MobileViTModel.from_pretrained(model_dir, local_files_only=True)
    return write_weight_file(out_path, from_hf_state_dict(model.state_dict()), f16=f16, pretransposed=pretransposed)


def synthetic_images(n: int, h: int = 256, w: int = 256, seed: int = 7) -> np.ndarray:
    """image 0 = the reference's own test pattern (main.cpp:680-688); the rest are
    structured (per-channel offset + sinusoid + 0.15 U(0,1) noise, clamped to [0,1]).  Returns [n,h,w,3] f32."""
    rng = np.random.default_rng(seed)
    imgs = np.empty((n, h, w, 3), dtype=np.float32)
    idx = np.arange(h * w * 3, dtype=np.int64)
    imgs[0] = ((idx % 256) / 255.0).astype(np.float32).reshape(h, w, 3)
    yy, xx = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    for i in range(1, n):
        for c in range(3):
            off = rng.uniform(0.2, 0.8)
            fx, fy = rng.uniform(0.5, 6.0, 2) * 2 * np.pi / np.array([w, h])
            ph = rng.uniform(0, 2 * np.pi)
            amp = rng.uniform(0.1, 0.4)
            img = off + amp * np.sin(fx * xx + fy * yy + ph) + 0.15 * rng.uniform(0, 1, (h, w))
            imgs[i, :, :, c] = np.clip(img, 0.0, 1.0)
    return imgs


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="weight.ggml tooling (layout of mobilevit/convert-tf-to-ggml.py:16-33)")
    ap.add_argument("--hf", help="local Hugging Face MobileViT checkpoint directory to export")
    ap.add_argument("--synthetic", choices=sorted(VARIANTS), help="write random-init weights of this variant instead")
    ap.add_argument("--classes", type=int, default=0, help="with --synthetic: add a classifier head with this many classes")
    ap.add_argument("--out", default="weight.ggml")
    ap.add_argument("--f16", default="none", choices=["none", "conv", "all"], help="store convolution kernels (lossless: the loader rounds them "
                    "to f16 anyway) or every tensor as f16 (convert-tf-to-ggml.py:13 TODO); the unmodified main.cpp reads only 'none'")
    ap.add_argument("--pretransposed", action="store_true", help="store dense kernels (out, in) (convert-tf-to-ggml.py:14 TODO)")
    a = ap.parse_args()
    if a.hf:
        n = export_hf_checkpoint(a.hf, a.out, a.f16, a.pretransposed)
    elif a.synthetic:
        n = write_weight_file(a.out, make_synthetic_weights(a.synthetic, num_classes=a.classes), f16=a.f16, pretransposed=a.pretransposed)
    else:
        ap.error("give --hf DIR or --synthetic VARIANT")
    print(f"{a.out}: {n} floats")
