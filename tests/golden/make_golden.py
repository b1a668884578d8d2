"""Generate tests/golden/cyclegan_golden.npz from the float64 CPU oracle.

The reference (TensorFlow 1.x graph) cannot run here, so these vectors pin the ORACLE (and through it the CUDA
path) against regressions; they are not outputs of the reference itself -- "parity unpinned", see DESIGN.md.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cyclegan_oracle as O  # noqa: E402

SEED_W, SEED_X = 1234, 77


def main():
    P = O.init_params(seed=SEED_W, dtype=torch.float64, perturb_affine=True)
    A, B = O.synthetic_batch(seed=SEED_X, batch=1, frames=128, dtype=torch.float64)
    out = {"seed_w": SEED_W, "seed_x": SEED_X}
    with torch.no_grad():
        taps = {}
        y = O.generator_forward(A, P, "generator_A2B", taps)
        out["gen_A2B_out"] = y.numpy()
        for k in ("h1_glu", "d2", "r6", "u2"):
            out["gen_tap_" + k + "_first64"] = taps[k].numpy().reshape(-1)[:64]
            out["gen_tap_" + k + "_norm"] = np.float64(taps[k].norm())
        out["disc_A_out"] = O.discriminator_forward(A, P, "discriminator_A").numpy()
        A516, _ = O.synthetic_batch(seed=SEED_X + 1, batch=1, frames=516, dtype=torch.float64)
        out["gen_B2A_out_T516"] = O.generator_forward(A516, P, "generator_B2A").numpy()
    L, G, gA, gB = O.gradients(A, B, P, 10.0, 5.0)
    out["losses"] = np.array([float(L[k]) for k in O.LOSS_NAMES])
    out["generation_A"] = gA.numpy(); out["generation_B"] = gB.numpy()
    names = list(G.keys())
    out["grad_norms"] = np.array([float(G[k].norm()) for k in names])
    for k in ("generator_A2B/h1_conv/kernel", "generator_B2A/residual1d_block3_h2_conv/kernel", "discriminator_A/downsample2d_block3_h1_gates/kernel",
              "generator_A2B/InstanceNorm_7/gamma", "discriminator_B/dense/kernel"):
        out["grad_first64/" + k] = G[k].numpy().reshape(-1)[:64]
    m = O.OracleCycleGAN(dtype=torch.float64, params={k: v.clone() for k, v in P.items()})
    g1, d1 = m.train(A.numpy(), B.numpy(), 10.0, 5.0, 2e-4, 1e-4)
    g2, d2 = m.train(A.numpy(), B.numpy(), 10.0, 5.0, 2e-4, 1e-4)
    out["train_losses"] = np.array([g1, d1, g2, d2], dtype=np.float64)
    out["post2_first64/generator_A2B/o1_conv/kernel"] = m.P["generator_A2B/o1_conv/kernel"].numpy().reshape(-1)[:64]
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cyclegan_golden.npz"), **out)
    print("wrote golden: losses", out["losses"], "train", out["train_losses"])


if __name__ == "__main__":
    main()
