"""Shared parity metrics: the gate is |d| <= atol + rtol*|ref| with atol = atol_rms * rms(ref)."""
import numpy as np


def parity_report(got: np.ndarray, ref: np.ndarray, rtol: float, atol_rms: float) -> dict:
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    d = np.abs(got - ref)
    rms = float(np.sqrt((ref ** 2).mean()))
    viol = d > (atol_rms * rms + rtol * np.abs(ref))
    return {
        "rel_l2": float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)),
        "max_abs": float(d.max()),
        "rms_ref": rms,
        "violations": int(viol.sum()),
        "n": int(ref.size),
    }


def top1_report(pooled_got: np.ndarray, pooled_ref: np.ndarray) -> dict:
    a, b = pooled_got.argmax(1), pooled_ref.argmax(1)
    srt = np.sort(pooled_ref, axis=1)
    return {"agree": float((a == b).mean()), "distinct": int(len(set(b.tolist()))),
            "min_margin": float((srt[:, -1] - srt[:, -2]).min())}
