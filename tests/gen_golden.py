"""Generates tests/golden/*.npz: Hugging Face torch MobileViTModel (f32, CPU) outputs on synthetic weights
and images.  HF MobileViT is the model the reference's convert-tf-to-ggml.py exports (convert.py:7-9), so
it pins the GRAPH SEMANTICS of main.cpp (unfold/fold, attention, kernel layouts, BN, residual rules);
ggml's f16 rounding points are not in HF, so the oracle is compared in PURE_F32 mode (tight) and in
ggml-faithful mode (loose).  Run in the dev container (needs transformers + torch CPU):

    python tests/gen_golden.py

The fixtures are small (pooled vectors + sub-sampled feature maps + stage checksums)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ggml_experiments_b200 import weights as W  # noqa: E402


def hf_forward(variant, tensors, imgs_hwc):
    from transformers import MobileViTConfig, MobileViTModel
    cfg = MobileViTConfig(image_size=imgs_hwc.shape[1], **W.hf_config_kwargs(variant))
    model = MobileViTModel(cfg, expand_output=True).eval()
    sd = {k: torch.from_numpy(v) for k, v in W.to_hf_state_dict(tensors).items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [m for m in missing if "num_batches_tracked" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    x = torch.from_numpy(imgs_hwc).permute(0, 3, 1, 2).contiguous()
    with torch.no_grad():
        out = model(x, output_hidden_states=True)
    feat = out.last_hidden_state.numpy()          # [N, C, H/32, W/32]
    pooled = out.pooler_output.numpy()            # [N, C]
    hidden = [h.numpy() for h in out.hidden_states]
    return feat, pooled, hidden


def hf_classify(variant, tensors, imgs_hwc, num_classes):
    """HF MobileViTForImageClassification: backbone + global pool + Linear head."""
    from transformers import MobileViTConfig, MobileViTForImageClassification
    cfg = MobileViTConfig(image_size=imgs_hwc.shape[1], num_labels=num_classes, **W.hf_config_kwargs(variant))
    model = MobileViTForImageClassification(cfg).eval()
    sd = {}
    for k, v in W.to_hf_state_dict(tensors).items():
        sd[k if k.startswith("classifier.") else "mobilevit." + k] = torch.from_numpy(v)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [m for m in missing if "num_batches_tracked" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    x = torch.from_numpy(imgs_hwc).permute(0, 3, 1, 2).contiguous()
    with torch.no_grad():
        return model(x).logits.numpy()


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    for variant, n_img, hw in (("xxs", 2, 256), ("xs", 1, 256), ("s", 2, 256), ("xxs", 1, 128)):
        tensors = W.make_synthetic_weights(variant, seed=1234)
        imgs = W.synthetic_images(n_img, hw, hw, seed=7)
        feat, pooled, hidden = hf_forward(variant, tensors, imgs)
        name = f"hf_{variant}_{hw}.npz"
        np.savez_compressed(
            os.path.join(ROOT, "tests", "golden", name),
            variant=variant, seed=1234, img_seed=7, n_img=n_img, hw=hw,
            pooled=pooled.astype(np.float32),
            feat=feat.astype(np.float32) if feat.size <= 100_000 else feat[:, ::8].astype(np.float32),
            feat_stride=1 if feat.size <= 100_000 else 8,
            feat_l2=np.sqrt((feat.astype(np.float64) ** 2).sum()),
            stage_l2=np.array([np.sqrt((h.astype(np.float64) ** 2).sum()) for h in hidden]),
            stage_shapes=np.array([list(h.shape) for h in hidden]),
        )
        print(name, feat.shape, "pooled range", pooled.min(), pooled.max(), "n_hidden", len(hidden))
    # classification head: XXS + 1000 classes
    tensors = W.make_synthetic_weights("xxs", seed=1234, num_classes=1000)
    imgs = W.synthetic_images(2, 256, 256, seed=7)
    logits = hf_classify("xxs", tensors, imgs, 1000)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hf_cls_xxs_256.npz"), variant="xxs", seed=1234, img_seed=7, n_img=2, hw=256,
                        num_classes=1000, logits=logits.astype(np.float32))
    print("hf_cls_xxs_256.npz", logits.shape, "logit range", logits.min(), logits.max(), "top1", logits.argmax(-1))


if __name__ == "__main__":
    main()
