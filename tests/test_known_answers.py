"""Analytic known-answer tests: hand-computed tensors for the TF-1.x semantics that are easiest to get silently wrong.

The reference holds no golden vectors and TensorFlow cannot run here, so the oracle is a restatement.
These tests do NOT go through the restatement to get their expected values: every expected tensor below is a literal
worked out by hand from the TensorFlow definition quoted next to it (or a closed form in plain Python scalars), and is
then demanded of BOTH restatements (oracle/cyclegan_oracle.py, oracle/numpy_ref.py) here and of the CUDA kernels through
the C ABI in the `gpu`-marked half of this file.

  * tf.layers.conv1d / conv2d, padding='same' (module.py:22-64): out = ceil(n / s); pad_total = max((out-1)*s + k - n, 0);
    pad_before = pad_total // 2 (the odd element goes AFTER); cross-correlation (no kernel flip):
        y[i] = sum_k w[k] * x[s*i + k - pad_before]
  * pixel_shuffler (module.py:135-146): a raw row-major tf.reshape [n, w, c] -> [n, 2w, c/2]
  * tf.contrib.layers.instance_norm (module.py:9-20): per (sample, channel) mean and BIASED variance over the spatial axes,
    y = (x - mean) / sqrt(var + 1e-6) * gamma + beta
  * tf.train.AdamOptimizer (model.py:107-108): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t * m / (sqrt(v) + eps), eps OUTSIDE
    the bias correction
"""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import cyclegan_oracle as O
from oracle import numpy_ref as NR

# ------------------------------------------------------------------------------------------------ literals
# conv1d, k = 5, stride 2, T = 8, w = [1,2,3,4,5]: pad_total = 3 -> (1, 2);  y[i] = sum_k w[k] x[2i + k - 1].
# Row p = response to a unit impulse at position p.  (The "extra pad in front" convention would give row 0 = [3,1,0,0].)
K5S2_T8 = [[2, 0, 0, 0],
           [3, 1, 0, 0],
           [4, 2, 0, 0],
           [5, 3, 1, 0],
           [0, 4, 2, 0],
           [0, 5, 3, 1],
           [0, 0, 4, 2],
           [0, 0, 5, 3]]
# conv1d, k = 3, stride 1, T = 4, w = [1,2,3]: pad (1,1); y[i] = sum_k w[k] x[i + k - 1]   (a flipped kernel would give row 0 = [2,3,0,0])
K3S1_T4 = [[2, 1, 0, 0],
           [3, 2, 1, 0],
           [0, 3, 2, 1],
           [0, 0, 3, 2]]
# conv1d, k = 5, stride 2, odd T = 5: out = 3, pad_total = (3-1)*2 + 5 - 5 = 4 -> (2, 2); y[i] = sum_k w[k] x[2i + k - 2]
K5S2_T5 = [[3, 1, 0],
           [4, 2, 0],
           [5, 3, 1],
           [0, 4, 2],
           [0, 5, 3]]


def w63(a, b):
    return 10 * (a + 1) + (b + 1)


# conv2d 6x3, strides (1,2), H = 6, W = 4 (the discriminator's last block, module.py:208): H pad_total = 5 -> (2, 3),
# W pad_total = (2-1)*2 + 3 - 4 = 1 -> (0, 1);  y[i,j] = sum_ab w[a,b] x[i + a - 2, 2j + b],  w[a,b] = 10(a+1) + (b+1)
def k63_expected(h, wc):
    y = np.zeros((6, 2))
    if (h, wc) == (0, 0):
        y[0, 0], y[1, 0], y[2, 0] = 31, 21, 11
    elif (h, wc) == (5, 3):
        y[2, 1], y[3, 1], y[4, 1], y[5, 1] = 62, 52, 42, 32
    elif (h, wc) == (3, 2):
        y[:, 0] = [63, 53, 43, 33, 23, 13]
        y[:, 1] = [61, 51, 41, 31, 21, 11]
    else:
        raise KeyError
    return y


# conv2d 3x3, strides (2,2), H = W = 4 (module.py:206-207): pad_total = 1 -> (0, 1) both ways; y[i,j] = sum w[a,b] x[2i + a, 2j + b]
K33S22 = {(0, 0): {(0, 0): 11}, (3, 3): {(1, 1): 22}, (2, 1): {(0, 0): 32, (1, 0): 12}}
# conv2d 3x3, strides (1,2), H = 3, W = 4 (discriminator h1, module.py:196-197): H pad (1,1), W pad (0,1); y[i,j] = sum w[a,b] x[i + a - 1, 2j + b]
K33S12 = {(0, 0): {(0, 0): 21, (1, 0): 11}, (2, 3): {(1, 1): 32, (2, 1): 22}}

IN_UNIT = 0.999999500000375        # 1 / sqrt(1 + 1e-6)
IN_SMALL = 0.7071067811865476      # 1e-3 / sqrt(1e-6 + 1e-6): an epsilon of 1e-5 would give 0.3015
IN_VAR4 = 0.9999998750000235       # 2 / sqrt(4 + 1e-6)

# TF Adam, constant scalar gradient g = 1e-7 (comparable to eps / sqrt(1 - b2^t), so the epsilon placement matters:
# torch-style Adam would move 1.82e-4 on the first step), lr = 2e-4, b1 = 0.5, b2 = 0.999, eps = 1e-8, p0 = 1
ADAM_UPDATES = [4.805061467040843e-05, 6.179272044048543e-05, 7.075500626505954e-05]


def _impulse_conv(conv, T, k, s, w, diag_channels=1):
    """rows p: response of `conv` (x [1,1,T,C] -> y [1,1,To,C], kernel [1,k,C,C] diagonal = w) to a unit impulse at p, channel-scaled."""
    Cn = diag_channels
    kern = np.zeros((1, k, Cn, Cn))
    for c in range(Cn):
        kern[0, :, c, c] = w
    rows = []
    for p in range(T):
        x = np.zeros((1, 1, T, Cn)); x[0, 0, p, :] = np.arange(1, Cn + 1)
        rows.append(conv(x, kern, (1, s)))
    return rows


def _torch_conv(x, kern, strides):
    y = O.conv2d_same(torch.tensor(x), torch.tensor(kern), torch.zeros(kern.shape[-1], dtype=torch.float64), strides)
    return y.numpy()


def _numpy_conv(x, kern, strides):
    return NR.conv_fwd(x, kern, np.zeros(kern.shape[-1]), strides)


CONVS = {"torch_oracle": _torch_conv, "numpy_ref": _numpy_conv}


@pytest.mark.parametrize("impl", sorted(CONVS))
def test_conv1d_same_tap_positions(impl):
    conv = CONVS[impl]
    for T, k, s, w, table in ((8, 5, 2, [1, 2, 3, 4, 5], K5S2_T8), (4, 3, 1, [1, 2, 3], K3S1_T4), (5, 5, 2, [1, 2, 3, 4, 5], K5S2_T5)):
        rows = _impulse_conv(conv, T, k, s, w)
        got = np.array([r[0, 0, :, 0] for r in rows])
        assert np.array_equal(got, np.array(table, dtype=np.float64)), (impl, T, k, s, got)
    # the 1-D wrapper of the torch oracle (module.py:22-42) places the taps the same way
    if impl == "torch_oracle":
        for p in range(8):
            x = torch.zeros(1, 8, 1, dtype=torch.float64); x[0, p, 0] = 1
            y = O.conv1d_same(x, torch.tensor([1., 2, 3, 4, 5], dtype=torch.float64).reshape(5, 1, 1), None, 2)
            assert y.reshape(-1).tolist() == K5S2_T8[p]


@pytest.mark.parametrize("impl", sorted(CONVS))
def test_conv2d_same_tap_positions(impl):
    conv = CONVS[impl]
    kern = np.zeros((6, 3, 1, 1))
    for a in range(6):
        for b in range(3):
            kern[a, b, 0, 0] = w63(a, b)
    for (h, wc) in ((0, 0), (5, 3), (3, 2)):
        x = np.zeros((1, 6, 4, 1)); x[0, h, wc, 0] = 1
        y = conv(x, kern, (1, 2))
        assert y.shape == (1, 6, 2, 1)
        assert np.array_equal(y[0, :, :, 0], k63_expected(h, wc)), (impl, h, wc, y[0, :, :, 0])
    k33 = kern[:3]
    for table, H, W, strides in ((K33S22, 4, 4, (2, 2)), (K33S12, 3, 4, (1, 2))):
        for (h, wc), want in table.items():
            x = np.zeros((1, H, W, 1)); x[0, h, wc, 0] = 1
            y = conv(x, k33, strides)[0, :, :, 0]
            exp = np.zeros_like(y)
            for (i, j), v in want.items():
                exp[i, j] = v
            assert np.array_equal(y, exp), (impl, strides, h, wc, y)


def test_pixel_shuffle_literal():
    x = torch.arange(8, dtype=torch.float64).reshape(1, 2, 4)          # [[0,1,2,3],[4,5,6,7]]
    assert O.pixel_shuffle_reshape(x).tolist() == [[[0, 1], [2, 3], [4, 5], [6, 7]]]
    assert NR.shuffle_fwd(x.numpy().reshape(1, 1, 2, 4)).reshape(1, 4, 2).tolist() == [[[0, 1], [2, 3], [4, 5], [6, 7]]]


def test_instance_norm_literals():
    # 1-D: x [1, 4, 1] = [3,1,3,1]: mean 2, biased var 1
    x = torch.tensor([3., 1, 3, 1], dtype=torch.float64).reshape(1, 4, 1)
    y = O.instance_norm(x, torch.tensor([0.5], dtype=torch.float64), torch.tensor([2.0], dtype=torch.float64))
    assert np.allclose(y.reshape(-1).numpy(), [0.5 + 2 * IN_UNIT, 0.5 - 2 * IN_UNIT] * 2, rtol=0, atol=1e-15)
    # epsilon = 1e-6 inside the square root
    xs = 1e-3 * torch.tensor([1., -1, 1, -1], dtype=torch.float64).reshape(1, 4, 1)
    ys = O.instance_norm(xs, torch.zeros(1, dtype=torch.float64), torch.ones(1, dtype=torch.float64))
    assert np.allclose(ys.reshape(-1).numpy(), [IN_SMALL, -IN_SMALL] * 2, rtol=1e-14)
    # 2-D, statistics over H x W per (sample, channel); two samples must not mix
    x2 = torch.zeros(2, 2, 2, 2, dtype=torch.float64)
    x2[0, :, :, 0] = torch.tensor([[3., 1], [3, 1]]); x2[0, :, :, 1] = torch.tensor([[0., 0], [4, 4]])
    x2[1, :, :, 0] = 100 + torch.tensor([[3., 1], [3, 1]]); x2[1, :, :, 1] = -7
    y2 = O.instance_norm(x2, torch.zeros(2, dtype=torch.float64), torch.ones(2, dtype=torch.float64)).numpy()
    assert np.allclose(y2[0, :, :, 0], [[IN_UNIT, -IN_UNIT]] * 2, atol=1e-15)
    assert np.allclose(y2[0, :, :, 1], [[-IN_VAR4] * 2, [IN_VAR4] * 2], atol=1e-15)
    assert np.allclose(y2[1, :, :, 0], [[IN_UNIT, -IN_UNIT]] * 2, atol=1e-12) and np.all(y2[1, :, :, 1] == 0)
    yn, _ = NR.in_fwd(x2.numpy(), np.zeros(2), np.ones(2))
    assert np.allclose(yn, y2, atol=1e-15)


def test_glu_and_losses_literals():
    a = torch.tensor([2.0, -4.0], dtype=torch.float64); g = torch.tensor([0.0, math.log(3.0)], dtype=torch.float64)
    assert np.allclose(O.glu(a, g).numpy(), [1.0, -3.0], atol=1e-15)              # a * sigmoid(g): sigmoid(0) = 1/2, sigmoid(ln 3) = 3/4
    y = torch.tensor([1.0, 2.0, 3.0, 4.0]); yh = torch.tensor([2.0, 2.0, 1.0, 8.0])
    assert float(O.l1_loss(y, yh)) == 1.75 and float(O.l2_loss(y, yh)) == 5.25     # utils.py:6-12: means, not sums


def test_tf_adam_closed_form_three_steps():
    """Constant gradient g: m_t = (1-b1^t) g, v_t = (1-b2^t) g^2, so the TF update is lr * g * s / (|g| s + eps), s = sqrt(1-b2^t)."""
    P = {"generator_p": torch.tensor([1.0], dtype=torch.float64)}
    opt = O.TFAdam(P)
    p = 1.0
    for t in (1, 2, 3):
        opt.apply(P, {"generator_p": torch.tensor([1e-7], dtype=torch.float64)}, 2e-4, 1e-4)
        s = math.sqrt(1.0 - 0.999 ** t)
        u = 2e-4 * 1e-7 * s / (1e-7 * s + 1e-8)
        assert abs(u - ADAM_UPDATES[t - 1]) < 1e-18
        p -= u
        assert abs(float(P["generator_p"][0]) - p) < 1e-15, t


def test_generator_edge_transposes_and_shapes():
    """module.py:152,183: the generator works on [B,T,C] inside; an input whose frames differ gives outputs aligned per frame
    (a wrong transpose would mix MCEP index and time).  With all-zero conv kernels the output is the o1 bias per MCEP coefficient."""
    P = O.init_params(seed=0, dtype=torch.float64)
    for k in P:
        if k.startswith("generator_A2B") and k.endswith("/kernel"):
            P[k] = torch.zeros_like(P[k])
    P["generator_A2B/o1_conv/bias"] = torch.arange(24, dtype=torch.float64)
    x = torch.randn(2, 24, 16, dtype=torch.float64)
    y = O.generator_forward(x, P, "generator_A2B")
    assert tuple(y.shape) == (2, 24, 16)
    assert torch.equal(y, torch.arange(24, dtype=torch.float64).reshape(1, 24, 1).expand(2, 24, 16))


# =================================================================================================== GPU: the CUDA kernels
def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


@pytest.fixture(scope="module")
def eng():
    import cgvc  # noqa: F401
    from cgvc import native as N
    lib = N.load()
    cfg = N.Config(24, 1, 128, N.PREC_FP32_SIMT, 0, 0)
    h = C.c_void_p(0)
    assert lib.cgvc_create(C.byref(cfg), C.byref(h)) == 0, lib.cgvc_last_error(None)
    yield lib, h, N
    lib.cgvc_destroy(h)


def _gpu_conv(eng, prec):
    lib, h, N = eng

    def conv(x, kern, strides):
        xd = torch.tensor(x, dtype=torch.float32).cuda(); wd = torch.tensor(kern, dtype=torch.float32).cuda()
        B, H, W, Cin = x.shape
        kh, kw, _, Cout = kern.shape
        Ho, Wo = -(-H // strides[0]), -(-W // strides[1])
        y = torch.full((B, Ho, Wo, Cout), float("nan"), dtype=torch.float32, device="cuda")
        code = lib.cgvc_conv_forward(h, prec, _p(xd), _p(wd), None, _p(y), B, H, W, Cin, kh, kw, Cout, strides[0], strides[1], None)
        N.check(h, code)
        torch.cuda.synchronize()
        return y.cpu().numpy().astype(np.float64)
    return conv


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
def test_gpu_conv_tap_positions(eng, prec):
    """The same hand-computed impulse responses through cgvc_conv_forward (small integers: exact in bf16 hi/lo too).
    4 diagonal channels so that the tensor-core path accepts the shape and a channel mix-up shows."""
    from cgvc import native as N
    conv = _gpu_conv(eng, {"fp32": N.PREC_FP32_SIMT, "bf16x3": N.PREC_BF16X3}[prec])
    for T, k, s, w, table in ((8, 5, 2, [1, 2, 3, 4, 5], K5S2_T8), (4, 3, 1, [1, 2, 3], K3S1_T4), (5, 5, 2, [1, 2, 3, 4, 5], K5S2_T5)):
        rows = _impulse_conv(conv, T, k, s, w, diag_channels=4)
        for c in range(4):
            got = np.array([r[0, 0, :, c] for r in rows])
            assert np.array_equal(got, (c + 1) * np.array(table, dtype=np.float64)), (prec, T, k, s, c, got)
    kern = np.zeros((6, 3, 4, 4))
    for a in range(6):
        for b in range(3):
            for c in range(4):
                kern[a, b, c, c] = w63(a, b)
    for (hh, wc) in ((0, 0), (5, 3), (3, 2)):
        x = np.zeros((1, 6, 4, 4)); x[0, hh, wc, :] = [1, 2, 3, 4]
        y = conv(x, kern, (1, 2))
        for c in range(4):
            assert np.array_equal(y[0, :, :, c], (c + 1) * k63_expected(hh, wc)), (prec, hh, wc, c)
    for table, H, W, strides in ((K33S22, 4, 4, (2, 2)), (K33S12, 3, 4, (1, 2))):
        for (hh, wc), want in table.items():
            x = np.zeros((1, H, W, 4)); x[0, hh, wc, :] = [1, 2, 3, 4]
            y = conv(x, kern[:3], strides)
            exp = np.zeros(y.shape[1:3])
            for (i, j), v in want.items():
                exp[i, j] = v
            for c in range(4):
                assert np.array_equal(y[0, :, :, c], (c + 1) * exp), (prec, strides, hh, wc, c)


@pytest.mark.gpu
def test_gpu_instance_norm_glu_shuffle_literals(eng):
    """cgvc_in_glu_forward: instance norm (eps 1e-6, biased variance, per sample), sigmoid gate and the raw-reshape shuffle."""
    lib, h, N = eng
    Cn = 32
    dev = "cuda"
    ones, zeros = torch.ones(Cn, device=dev), torch.zeros(Cn, device=dev)
    # (a) two samples x 4 positions x 32 channels, a = [3,1,3,1] (+100 for sample 1), gate branch constant with gamma_g = 0 and
    #     beta_g = ln 3  ->  y = 0.75 * (0.5 +- 2 / sqrt(1 + 1e-6))
    P = torch.zeros(2, 4, 2 * Cn, device=dev)
    P[0, :, :Cn] = torch.tensor([3., 1, 3, 1], device=dev).reshape(4, 1)
    P[1, :, :Cn] = 100 + torch.tensor([3., 1, 3, 1], device=dev).reshape(4, 1)
    P[:, :, Cn:] = 5.0
    y = torch.empty(2, 4, Cn, device=dev); stats = torch.empty(2, 4, Cn, device=dev)
    beta_a, gamma_a, beta_g = 0.5 * ones, 2 * ones, math.log(3.0) * ones        # named: the kernel reads them after _p() returns
    N.check(h, lib.cgvc_in_glu_forward(h, _p(P), _p(beta_a), _p(gamma_a), _p(beta_g), _p(zeros), _p(y), _p(stats), 2, 4, Cn, 1, None))
    torch.cuda.synchronize()
    want = 0.75 * np.array([0.5 + 2 * IN_UNIT, 0.5 - 2 * IN_UNIT] * 2)
    for b in range(2):
        assert np.allclose(y[b].cpu().numpy(), want.reshape(4, 1), rtol=0, atol=(2e-5 if b else 2e-6)), b
    assert np.allclose(stats[0, 0].cpu().numpy(), 2.0) and np.allclose(stats[0, 1].cpu().numpy(), IN_UNIT, atol=1e-6)
    # (b) epsilon: a = +-1e-3 -> 0.7071 (1e-5 would give 0.30)
    P = torch.zeros(1, 4, 2 * Cn, device=dev)
    P[0, :, :Cn] = 1e-3 * torch.tensor([1., -1, 1, -1], device=dev).reshape(4, 1)
    y = torch.empty(1, 4, Cn, device=dev); stats = torch.empty(1, 4, Cn, device=dev)
    N.check(h, lib.cgvc_in_glu_forward(h, _p(P), _p(zeros), _p(ones), _p(zeros), _p(zeros), _p(y), _p(stats), 1, 4, Cn, 1, None))
    torch.cuda.synchronize()
    assert np.allclose(y[0].cpu().numpy(), 0.5 * IN_SMALL * np.array([1, -1, 1, -1]).reshape(4, 1), rtol=1e-5)
    # (c) shuffle = 2: conv rows w = 0..3 with 2*32 'a' columns; out[2w + s, o] = a[w, s*32 + o] (raw reshape).  The sign pattern
    #     (-1)^(2w + s + o) and gamma_a[o] = o + 1 identify position and channel; IN of a +-1 signal is +-1/sqrt(1+1e-6).
    R, Wc = 8, 4
    a = torch.zeros(Wc, 2 * Cn)
    for w in range(Wc):
        for s in range(2):
            for o in range(Cn):
                a[w, s * Cn + o] = 1.0 if (2 * w + s + o) % 2 == 0 else -1.0
    P = torch.zeros(1, Wc, 4 * Cn, device=dev)
    P[0, :, :2 * Cn] = a.to(dev)
    gam = torch.arange(1, Cn + 1, dtype=torch.float32, device=dev)
    y = torch.empty(1, R, Cn, device=dev); stats = torch.empty(1, 4, Cn, device=dev)
    N.check(h, lib.cgvc_in_glu_forward(h, _p(P), _p(zeros), _p(gam), _p(zeros), _p(zeros), _p(y), _p(stats), 1, R, Cn, 2, None))
    torch.cuda.synchronize()
    want = np.zeros((R, Cn))
    for r in range(R):
        for o in range(Cn):
            want[r, o] = 0.5 * (o + 1) * IN_UNIT * (1.0 if (r + o) % 2 == 0 else -1.0)
    assert np.allclose(y[0].cpu().numpy(), want, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("t0,lr_g,lr_d,gscale,pscale", [(0, 2e-4, 1e-4, 1.0, 1.0), (999, 2e-4, 1e-4, 0.125, 1.0),
                                                        (0, 1e-6, 3e-7, 0.125, 1e-3), (999, 1e-6, 3e-7, 1.0, 1e-3)])
def test_gpu_adam_step_matches_tf_formula(t0, lr_g, lr_d, gscale, pscale):
    """cgvc_adam_step on injected p / g / m / v arenas against the TF formula (and O.TFAdam), elementwise:
    p, m, v to <= 1e-6, and the update p1 - p0 itself to fp32 rounding.  Both learning-rate ranges (with parameters scaled so
    that the step is resolvable in fp32), step 1 and step 1000, the data-parallel grad_scale = 1/8, gradient magnitudes spanning
    1e-10 .. 1e-1 so that the epsilon placement (outside the bias correction) matters for a large part of the elements, and the
    generator / discriminator ranges each taking their own learning rate."""
    import cgvc
    from cgvc import native as N
    m = cgvc.CycleGAN(num_features=24, mode='train', max_batch=1, max_frames=128, precision="fp32", log_dir='/tmp/cgvc_log')
    n = max(off + int(np.prod(shp)) for off, shp in m._table.values())      # the arena's used length (its 256-byte rounding tail is never touched)
    gen = torch.Generator(device="cuda").manual_seed(5 + t0)
    p0 = pscale * (2 * torch.rand(n, device="cuda", generator=gen) - 1)
    mag = 10.0 ** (torch.rand(n, device="cuda", generator=gen) * 9 - 10)                 # 1e-10 .. 1e-1
    g = mag * torch.sign(torch.randn(n, device="cuda", generator=gen))
    m0 = 0.3 * gscale * g * torch.rand(n, device="cuda", generator=gen) if t0 else torch.zeros(n, device="cuda")
    v0 = (gscale * g) ** 2 * torch.rand(n, device="cuda", generator=gen) if t0 else torch.zeros(n, device="cuda")
    m._arenas[N.ARENA_PARAM][:n].copy_(p0); m._arenas[N.ARENA_GRAD][:n].copy_(g)
    m._arenas[N.ARENA_ADAM_M][:n].copy_(m0); m._arenas[N.ARENA_ADAM_V][:n].copy_(v0)
    assert m._lib.cgvc_set_adam_step(m._handle, t0) == 0
    N.check(m._handle, m._lib.cgvc_adam_step(m._handle, C.c_float(lr_g), C.c_float(lr_d), C.c_float(gscale), m._stream()))
    torch.cuda.synchronize()
    step = C.c_longlong(0); m._lib.cgvc_get_adam_step(m._handle, C.byref(step))
    t = t0 + 1
    assert step.value == t
    ge = m._generator_end
    # expected, in float64, straight from the TF definition (not via the oracle class)
    gd, pd, md, vd = ((g * gscale).double(), p0.double(), m0.double(), v0.double())
    # beta2 and epsilon are float32 attributes of the TF op (0.999f = 0.99900001287..., so 1 - beta2 = 9.99987e-4, not 1e-3)
    b2, eps = float(np.float32(0.999)), float(np.float32(1e-8))
    m1 = 0.5 * md + 0.5 * gd
    v1 = b2 * vd + (1.0 - b2) * gd * gd
    corr = math.sqrt(1.0 - b2 ** t) / (1.0 - 0.5 ** t)
    lr = torch.full((n,), lr_d, dtype=torch.float64, device="cuda"); lr[:ge] = lr_g
    p1 = pd - lr * corr * m1 / (v1.sqrt() + eps)
    got_p, got_m, got_v = (m._arenas[k][:n].double() for k in (N.ARENA_PARAM, N.ARENA_ADAM_M, N.ARENA_ADAM_V))
    assert float((got_p - p1).abs().max()) <= 1e-6 * pscale
    upd, upd_ref = got_p - pd, p1 - pd
    assert float(((upd - upd_ref).abs() - 2.0 ** -23 * pscale - 5e-6 * upd_ref.abs()).max()) <= 0.0
    assert float(((got_m - m1).abs() / (m1.abs() + 1e-30)).max()) < 1e-6
    assert float(((got_v - v1).abs() / (v1.abs() + 1e-30)).max()) < 1e-6
    # the test can tell the two Adams apart: torch-style Adam (eps inside the bias correction) is off by a large part of a step
    # wherever |g| ~ eps / sqrt(1 - b2^t), far outside the tolerances above
    p_torch = pd - lr / (1.0 - 0.5 ** t) * m1 / ((v1 / (1.0 - b2 ** t)).sqrt() + eps)
    assert float((p_torch - p1).abs().max()) > 100 * 2.0 ** -23 * pscale
    # and the oracle's optimizer class states the same formula
    names = ["generator_x", "discriminator_y"]
    idx = [slice(0, 4096), slice(ge, ge + 4096)]
    Pd = {k: pd[s].cpu().clone() for k, s in zip(names, idx)}
    opt = O.TFAdam(Pd)
    opt.t = t0
    for k, s in zip(names, idx):
        opt.m[k] = md[s].cpu().clone(); opt.v[k] = vd[s].cpu().clone()
    opt.apply(Pd, {k: gd[s].cpu() for k, s in zip(names, idx)}, lr_g, lr_d)
    for k, s in zip(names, idx):
        # (the oracle keeps beta2 = 0.999 in double: 1.3e-5 relative in v and in 1 - beta2^t, i.e. ~1e-5 of a step)
        assert float((Pd[k] - p1[s].cpu()).abs().max()) < 1e-4 * max(lr_g, lr_d)
