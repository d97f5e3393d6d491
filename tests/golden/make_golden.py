"""Regenerates tests/golden/radon_golden.npz from the CPU oracle.

The reference itself cannot run here (TensorFlow / tfa / tfp are not installable), so
these vectors pin the ORACLE's outputs, not TensorFlow's: they guard the oracle and the
CUDA path against regressions.  The analytic known answers (toy sinograms, theta=0
column sums) live in tests/test_oracle.py.   Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import radon_oracle as orc  # noqa: E402


def main():
    rng = np.random.default_rng(20261018)
    B, X, Y, A = 3, 24, 24, 10
    img = orc.synthetic_foam(B, X, seed=7) * rng.random((B, X, Y), dtype=np.float32)
    theta = np.linspace(0, np.pi, A, endpoint=False)
    W = orc.frame_of(X, Y, True)[1]
    cot = rng.random((B, A, W), dtype=np.float32)
    out = {"img": img.astype(np.float32), "theta": theta, "cot": cot}
    for name, iid in (("nearest", 0), ("bilinear", 1)):
        out[f"sino_{name}"] = orc.forward(img, theta, True, iid)
        out[f"grad_exact_{name}"] = orc.adjoint_exact(cot, theta, X, Y, True, iid)
        out[f"grad_tf_compat_{name}"] = orc.adjoint_tf(cot, theta, X, Y, True, iid)
    filt = orc.get_fourier_filter(W, "ramp")
    fbp_sino = out["sino_bilinear"].astype(np.float64)
    out["fbp_sino"] = fbp_sino
    out["fbp_filter"] = filt
    out["fbp_recon"] = orc.iradon(fbp_sino, theta, X, Y, filt)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "radon_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
