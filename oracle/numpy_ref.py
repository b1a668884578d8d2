"""Second, independent CPU restatement (float64 numpy, no autograd)  --  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see oracle/cyclegan_oracle.py header): TensorFlow 1.x is not available, so the two
restatements pin each other and the structural known-answers, not TF itself.

This file states the forward AND the hand-derived backward of every primitive on the hot path in the
same "gather-GEMM" form the CUDA engine uses, so that

  * the torch/autograd oracle is checked against an implementation that shares no code with it, and
  * the formulas the kernels implement (conv dgrad/wgrad, instance-norm backward, GLU backward,
    loss gradients) are verified on the CPU before any GPU time is spent.

Everything is channels-last 2-D: x [B,H,W,C]; 1-D tensors use H = 1 and kh = 1.
References: module.py:3-213, utils.py:6-12 of /root/reference.
"""
from __future__ import annotations

import numpy as np

IN_EPS = 1e-6


def same_pad(n_in, k, s):
    out = -(-n_in // s)
    total = max((out - 1) * s + k - n_in, 0)
    return total // 2, total - total // 2, out


# ---------------------------------------------------------------- convolution (A.2)

def conv_fwd(x, w, b, stride):
    """x [B,H,W,Cin]; w [kh,kw,Cin,Cout]; stride (sh,sw); TF SAME.  y[b,ho,wo,:] = b + sum_taps x[...]@w[i,j]."""
    B, H, W, Cin = x.shape
    kh, kw, _, Cout = w.shape
    ph0, _, Ho = same_pad(H, kh, stride[0])
    pw0, _, Wo = same_pad(W, kw, stride[1])
    y = np.zeros((B, Ho, Wo, Cout), dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            for ho in range(Ho):
                h = ho * stride[0] + i - ph0
                if h < 0 or h >= H:
                    continue
                # valid wo range for this tap
                wos = [wo for wo in range(Wo) if 0 <= wo * stride[1] + j - pw0 < W]
                if not wos:
                    continue
                ws = [wo * stride[1] + j - pw0 for wo in wos]
                y[:, ho, wos, :] += x[:, h, ws, :] @ w[i, j]
    if b is not None:
        y += b
    return y


def conv_bwd(x, w, dy, stride):
    """Returns (dx, dw, db) for conv_fwd."""
    B, H, W, Cin = x.shape
    kh, kw, _, Cout = w.shape
    ph0, _, Ho = same_pad(H, kh, stride[0])
    pw0, _, Wo = same_pad(W, kw, stride[1])
    dx = np.zeros_like(x)
    dw = np.zeros_like(w)
    for i in range(kh):
        for j in range(kw):
            for ho in range(Ho):
                h = ho * stride[0] + i - ph0
                if h < 0 or h >= H:
                    continue
                wos = [wo for wo in range(Wo) if 0 <= wo * stride[1] + j - pw0 < W]
                if not wos:
                    continue
                ws = [wo * stride[1] + j - pw0 for wo in wos]
                g = dy[:, ho, wos, :]                                   # [B, n, Cout]
                dx[:, h, ws, :] += g @ w[i, j].T
                dw[i, j] += np.einsum("bnc,bnd->cd", x[:, h, ws, :], g)
    db = dy.sum(axis=(0, 1, 2))
    return dx, dw, db


# ---------------------------------------------------------------- instance norm (A.4, A.7)

def in_fwd(x, beta, gamma):
    axes = (1, 2)
    mean = x.mean(axis=axes, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=axes, keepdims=True)
    rstd = 1.0 / np.sqrt(var + IN_EPS)
    xhat = (x - mean) * rstd
    return xhat * gamma + beta, (xhat, rstd)


def in_bwd(dy, cache, gamma):
    xhat, rstd = cache
    R = xhat.shape[1] * xhat.shape[2]
    dxhat = dy * gamma
    s1 = dxhat.sum(axis=(1, 2), keepdims=True)
    s2 = (dxhat * xhat).sum(axis=(1, 2), keepdims=True)
    dx = (rstd / R) * (R * dxhat - s1 - xhat * s2)
    dgamma = (dy * xhat).sum(axis=(0, 1, 2))
    dbeta = dy.sum(axis=(0, 1, 2))
    return dx, dgamma, dbeta


# ---------------------------------------------------------------- GLU / sigmoid (A.7)

def sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def glu_fwd(a, g):
    s = sigmoid(g)
    return a * s, (a, s)


def glu_bwd(dy, cache):
    a, s = cache
    return dy * s, dy * a * s * (1.0 - s)


# ---------------------------------------------------------------- pixel shuffle (A.8)

def shuffle_fwd(x):           # [B,1,W,C] -> [B,1,2W,C/2] raw reshape
    B, H, W, C = x.shape
    return x.reshape(B, H, W * 2, C // 2)


def shuffle_bwd(dy):
    B, H, W2, C2 = dy.shape
    return dy.reshape(B, H, W2 // 2, C2 * 2)


# ---------------------------------------------------------------- networks (forward only; P is name -> ndarray)

def _in(i):
    return "InstanceNorm" if i == 0 else "InstanceNorm_%d" % i


def _k2d(k):
    """conv1d kernel [k,Cin,Cout] -> [1,k,Cin,Cout]."""
    return k[None] if k.ndim == 3 else k


def generator_forward(x, P, scope):
    p = lambda n: np.asarray(P[scope + "/" + n], dtype=np.float64)
    h = np.transpose(x, (0, 2, 1))[:, None]                 # [B,1,T,24]
    a = conv_fwd(h, _k2d(p("h1_conv/kernel")), p("h1_conv/bias"), (1, 1))
    g = conv_fwd(h, _k2d(p("h1_conv_gates/kernel")), p("h1_conv_gates/bias"), (1, 1))
    h = a * sigmoid(g)
    idx = 0

    def gated(h, pre, idx, stride, shuffle=False):
        a = conv_fwd(h, _k2d(p(pre + "h1_conv/kernel")), p(pre + "h1_conv/bias"), (1, stride))
        g = conv_fwd(h, _k2d(p(pre + "h1_gates/kernel")), p(pre + "h1_gates/bias"), (1, stride))
        if shuffle:
            a, g = shuffle_fwd(a), shuffle_fwd(g)
        a, _ = in_fwd(a, p(_in(idx) + "/beta"), p(_in(idx) + "/gamma"))
        g, _ = in_fwd(g, p(_in(idx + 1) + "/beta"), p(_in(idx + 1) + "/gamma"))
        return a * sigmoid(g)

    for i in (1, 2):
        h = gated(h, "downsample1d_block%d_" % i, idx, 2); idx += 2
    for i in range(1, 7):
        pre = "residual1d_block%d_" % i
        h1 = gated(h, pre, idx, 1)
        h2 = conv_fwd(h1, _k2d(p(pre + "h2_conv/kernel")), p(pre + "h2_conv/bias"), (1, 1))
        h2, _ = in_fwd(h2, p(_in(idx + 2) + "/beta"), p(_in(idx + 2) + "/gamma"))
        h = h + h2; idx += 3
    for i in (1, 2):
        h = gated(h, "upsample1d_block%d_" % i, idx, 1, shuffle=True); idx += 2
    o = conv_fwd(h, _k2d(p("o1_conv/kernel")), p("o1_conv/bias"), (1, 1))
    return np.transpose(o[:, 0], (0, 2, 1))


def discriminator_forward(x, P, scope):
    p = lambda n: np.asarray(P[scope + "/" + n], dtype=np.float64)
    h = x[..., None]
    a = conv_fwd(h, p("h1_conv/kernel"), p("h1_conv/bias"), (1, 2))
    g = conv_fwd(h, p("h1_conv_gates/kernel"), p("h1_conv_gates/bias"), (1, 2))
    h = a * sigmoid(g)
    idx = 0
    for i, st in ((1, (2, 2)), (2, (2, 2)), (3, (1, 2))):
        pre = "downsample2d_block%d_" % i
        a = conv_fwd(h, p(pre + "h1_conv/kernel"), p(pre + "h1_conv/bias"), st)
        g = conv_fwd(h, p(pre + "h1_gates/kernel"), p(pre + "h1_gates/bias"), st)
        a, _ = in_fwd(a, p(_in(idx) + "/beta"), p(_in(idx) + "/gamma"))
        g, _ = in_fwd(g, p(_in(idx + 1) + "/beta"), p(_in(idx + 1) + "/gamma"))
        h = a * sigmoid(g); idx += 2
    z = h @ p("dense/kernel") + p("dense/bias")
    return sigmoid(z)
