import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def weight_files(tmp_path_factory):
    """Random-init weight files in the convert-tf-to-ggml.py layout, one per variant (seed 1234)."""
    from ggml_experiments_b200 import weights as W
    d = tmp_path_factory.mktemp("weights")
    out = {}
    for v in ("xxs", "xs", "s"):
        p = str(d / f"weight_{v}.ggml")
        W.write_weight_file(p, W.make_synthetic_weights(v, seed=1234))
        out[v] = p
    # XXS with the optional 1000-class head; the 313 backbone tensors are identical to out["xxs"]
    p = str(d / "weight_xxs_cls.ggml")
    W.write_weight_file(p, W.make_synthetic_weights("xxs", seed=1234, num_classes=1000))
    out["xxs_cls"] = p
    return out


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.build()
    return binding
