"""Generate tests/golden/loss_curve.npz: the loss trajectory of a short training run in the CPU oracle.

The north star asks for "loss curves matching reference within tolerance".  TensorFlow 1.x cannot run here (parity
unpinned), so the curve is the ORACLE's: `STEPS` consecutive `train()` calls at batch `BATCH`
from hash-RNG weights, a fresh synthetic minibatch every step, lambda_identity switched off for the last quarter
(train.py:98-99 does that after 10k iterations), in float64 (the mathematical trajectory) and in float32 (what
fp32 TF arithmetic would follow; the f32-vs-f64 gap is the noise floor any fp32 implementation has).
The GPU test replays the same run on the engine and compares (tests/test_gpu_model.py::test_loss_curve_tracks_oracle).

Run from the repo root (CPU, ~10 min):  python tests/golden/make_loss_curve.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cyclegan_oracle as O  # noqa: E402

SEED_W, SEED_X, STEPS, BATCH = 4321, 500, 40, 2
LAMBDA_CYCLE, LR_G, LR_D = 10.0, 2e-4, 1e-4


def lambda_identity(step):
    return 5.0 if step < (3 * STEPS) // 4 else 0.0


def batch_for(step, dtype):
    return O.synthetic_batch(seed=SEED_X + step, batch=BATCH, frames=128, dtype=dtype)


def run(dtype):
    P = O.init_params(seed=SEED_W, dtype=dtype, perturb_affine=True)
    m = O.OracleCycleGAN(dtype=dtype, params=P)
    rows = []
    for t in range(STEPS):
        A, B = batch_for(t, dtype)
        m.train(A.numpy(), B.numpy(), LAMBDA_CYCLE, lambda_identity(t), LR_G, LR_D)
        rows.append([m.last_losses[k] for k in O.LOSS_NAMES])
        print(dtype, t, "G %.6f D %.6f" % (rows[-1][4], rows[-1][7]), flush=True)
    return np.array(rows, dtype=np.float64)


def main():
    c64 = run(torch.float64)
    c32 = run(torch.float32)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "loss_curve.npz"),
                        f64=c64, f32=c32, seed_w=SEED_W, seed_x=SEED_X, steps=STEPS, batch=BATCH, names=np.array(O.LOSS_NAMES))
    dev = np.abs(c32 - c64) / np.abs(c64)
    print("float32 oracle vs float64 oracle: worst relative loss deviation %.2e (step %d)" % (dev.max(), int(dev.max(axis=1).argmax())))


if __name__ == "__main__":
    main()
