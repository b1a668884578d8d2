"""CPU oracle for the CycleGAN-VC hot path  --  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference's arithmetic lives in TensorFlow 1.x (README.md:21,
Dockerfile:1), which is neither vendored under /root/reference nor installable here,
and the reference ships no tests / golden vectors for this path.
This file is therefore a *restatement* of the reference graph, written against the
TF-1.x semantics summarised in SURVEY.md Appendix A.  It is cross-checked against a
second, independent numpy restatement (oracle/numpy_ref.py) and against structural
known-answers derived from the reference text (parameter counts, shapes, losses at init).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product (voice-converter-cyclegan_b200/) never does.

What is restated (all citations into /root/reference):
  module.py:3-7     gated_linear_layer      -> glu()
  module.py:9-20    instance_norm_layer     -> instance_norm()   (tf.contrib.layers.instance_norm, eps=1e-6)
  module.py:22-42   conv1d_layer            -> conv1d_same()     (tf.layers.conv1d, padding='same')
  module.py:44-64   conv2d_layer            -> conv2d_same()
  module.py:66-83   residual1d_block
  module.py:85-98   downsample1d_block
  module.py:100-113 downsample2d_block
  module.py:115-133 upsample1d_block
  module.py:135-146 pixel_shuffler          (a raw row-major reshape)
  module.py:148-185 generator_gatedcnn      -> generator_forward()
  module.py:188-213 discriminator           -> discriminator_forward()
  utils.py:6-12     l1_loss / l2_loss
  model.py:44-90    graph wiring + loss algebra -> losses()
  model.py:103-125  two Adam optimizers, G step then D step from the same pre-update weights -> train_step()

Layout conventions: network inputs/outputs are [B, 24, T] like the reference; inside the
networks tensors are channels-last ([B, T, C] and [B, H, W, C]) like TF.  Parameters are
kept in TF variable layout: conv1d kernel [k, Cin, Cout], conv2d kernel [kh, kw, Cin, Cout],
dense kernel [1024, 1].
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

IN_EPS = 1e-6            # module.py:11
ADAM_BETA1 = 0.5         # model.py:107-108
ADAM_BETA2 = 0.999       # tf.train.AdamOptimizer default
ADAM_EPS = 1e-8          # tf.train.AdamOptimizer default
NUM_FEATURES = 24        # train.py:22

NETS = ("generator_A2B", "generator_B2A", "discriminator_A", "discriminator_B")

# --------------------------------------------------------------------------------------
# Parameter table (TF variable names, TF creation order).  module.py:148-213, Appendix A.4/A.5
# --------------------------------------------------------------------------------------

def _conv_entries(name, kshape, fan_in, fan_out):
    return [(name + "/kernel", tuple(kshape), ("glorot", fan_in, fan_out)),
            (name + "/bias", (kshape[-1],), ("zeros",))]


def _in_entries(idx, c):
    n = "InstanceNorm" if idx == 0 else "InstanceNorm_%d" % idx
    return [(n + "/beta", (c,), ("zeros",)), (n + "/gamma", (c,), ("ones",))]


def _conv1d(name, k, cin, cout):
    return _conv_entries(name, (k, cin, cout), k * cin, k * cout)


def _conv2d(name, kh, kw, cin, cout):
    return _conv_entries(name, (kh, kw, cin, cout), kh * kw * cin, kh * kw * cout)


def generator_param_specs(num_features=NUM_FEATURES):
    """110 tensors, 38,055,704 parameters."""
    s = []
    s += _conv1d("h1_conv", 15, num_features, 128)
    s += _conv1d("h1_conv_gates", 15, num_features, 128)
    inorm = 0
    cin = 128
    for i, cout in ((1, 256), (2, 512)):
        p = "downsample1d_block%d_" % i
        s += _conv1d(p + "h1_conv", 5, cin, cout); s += _in_entries(inorm, cout); inorm += 1
        s += _conv1d(p + "h1_gates", 5, cin, cout); s += _in_entries(inorm, cout); inorm += 1
        cin = cout
    for i in range(1, 7):
        p = "residual1d_block%d_" % i
        s += _conv1d(p + "h1_conv", 3, 512, 1024); s += _in_entries(inorm, 1024); inorm += 1
        s += _conv1d(p + "h1_gates", 3, 512, 1024); s += _in_entries(inorm, 1024); inorm += 1
        s += _conv1d(p + "h2_conv", 3, 1024, 512); s += _in_entries(inorm, 512); inorm += 1
    cin = 512
    for i, cout in ((1, 1024), (2, 512)):
        p = "upsample1d_block%d_" % i
        # IN is applied after the shuffle, so its channel count is cout // 2 (module.py:124-125)
        s += _conv1d(p + "h1_conv", 5, cin, cout); s += _in_entries(inorm, cout // 2); inorm += 1
        s += _conv1d(p + "h1_gates", 5, cin, cout); s += _in_entries(inorm, cout // 2); inorm += 1
        cin = cout // 2
    s += _conv1d("o1_conv", 15, 256, num_features)
    return s


def discriminator_param_specs():
    """30 tensors, 21,837,825 parameters."""
    s = []
    s += _conv2d("h1_conv", 3, 3, 1, 128)
    s += _conv2d("h1_conv_gates", 3, 3, 1, 128)
    inorm = 0
    cin = 128
    for i, (kh, kw, cout) in ((1, (3, 3, 256)), (2, (3, 3, 512)), (3, (6, 3, 1024))):
        p = "downsample2d_block%d_" % i
        s += _conv2d(p + "h1_conv", kh, kw, cin, cout); s += _in_entries(inorm, cout); inorm += 1
        s += _conv2d(p + "h1_gates", kh, kw, cin, cout); s += _in_entries(inorm, cout); inorm += 1
        cin = cout
    s += [("dense/kernel", (1024, 1), ("glorot", 1024, 1)), ("dense/bias", (1,), ("zeros",))]
    return s


def param_specs(num_features=NUM_FEATURES):
    """Full table in flat-arena order [G_A2B | G_B2A | D_A | D_B]: list of (name, shape, init)."""
    out = []
    for net in NETS:
        specs = generator_param_specs(num_features) if net.startswith("generator") else discriminator_param_specs()
        out += [(net + "/" + n, shp, init) for (n, shp, init) in specs]
    return out


# --------------------------------------------------------------------------------------
# Deterministic, library-independent initialisation (glorot-uniform, Appendix A.3).
# A counter-based hash RNG (splitmix64 finaliser) so that any implementation -- numpy here,
# C in the engine tests -- regenerates bit-identical weights from (seed, tensor index, element).
# --------------------------------------------------------------------------------------

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return x ^ (x >> np.uint64(31))


def hash_uniform(seed: int, stream: int, n: int) -> np.ndarray:
    """n float64 uniforms in [0,1): 53 high bits of splitmix64(seed, stream, i)."""
    with np.errstate(over="ignore"):
        base = _splitmix64(np.array([(seed * 0x9E3779B1 + stream * 0x85EBCA77 + 1) & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0]
        idx = np.arange(n, dtype=np.uint64)
        h = _splitmix64(idx + base)
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def hash_normal(seed: int, stream: int, n: int) -> np.ndarray:
    """n float64 standard normals (Box-Muller on hash_uniform)."""
    u1 = hash_uniform(seed, 2 * stream + 1000003, n)
    u2 = hash_uniform(seed, 2 * stream + 1000004, n)
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)


def init_params(seed=0, dtype=torch.float32, num_features=NUM_FEATURES, perturb_affine=False):
    """OrderedDict name -> tensor.  glorot-uniform kernels, zero biases/beta, unit gamma.

    perturb_affine=True additionally randomises biases / beta / gamma (small) so that parity
    tests exercise those code paths (at TF init they are 0 / 0 / 1 and several bugs would hide).
    """
    P = OrderedDict()
    for ti, (name, shape, init) in enumerate(param_specs(num_features)):
        n = int(np.prod(shape))
        if init[0] == "glorot":
            lim = math.sqrt(6.0 / (init[1] + init[2]))
            a = (2.0 * hash_uniform(seed, ti, n) - 1.0) * lim
        elif init[0] == "zeros":
            a = np.zeros(n)
            if perturb_affine:
                a = 0.1 * (2.0 * hash_uniform(seed, ti, n) - 1.0)
        else:
            a = np.ones(n)
            if perturb_affine:
                a = 1.0 + 0.2 * (2.0 * hash_uniform(seed, ti, n) - 1.0)
        # weights are defined as the fp32 rounding of the float64 stream, whatever dtype we compute in
        a32 = a.astype(np.float32)
        P[name] = torch.from_numpy(a32.astype(np.float64 if dtype == torch.float64 else np.float32).reshape(shape)).clone()
    return P


def synthetic_batch(seed, batch, frames=128, num_features=NUM_FEATURES, dtype=torch.float32):
    """A, B ~ N(0,1) as [batch, 24, frames] (z-normalised MCEPs are zero-mean/unit-var, preprocess.py:106-116)."""
    n = batch * num_features * frames
    a = hash_normal(seed, 1, n).astype(np.float32).reshape(batch, num_features, frames)
    b = hash_normal(seed, 2, n).astype(np.float32).reshape(batch, num_features, frames)
    cast = np.float64 if dtype == torch.float64 else np.float32
    return torch.from_numpy(a.astype(cast)), torch.from_numpy(b.astype(cast))


# --------------------------------------------------------------------------------------
# Primitives
# --------------------------------------------------------------------------------------

def same_pad(n_in, k, s):
    """TF 'SAME' padding (Appendix A.1): extra pad goes after."""
    out = -(-n_in // s)
    total = max((out - 1) * s + k - n_in, 0)
    return total // 2, total - total // 2


def conv1d_same(x, kernel, bias, stride=1):
    """x [N,W,Cin], kernel [k,Cin,Cout] (TF layout), cross-correlation.  module.py:22-42."""
    k = kernel.shape[0]
    pl, pr = same_pad(x.shape[1], k, stride)
    xt = F.pad(x.transpose(1, 2), (pl, pr))
    y = F.conv1d(xt, kernel.permute(2, 1, 0), bias, stride=stride)
    return y.transpose(1, 2)


def conv2d_same(x, kernel, bias, strides):
    """x [N,H,W,Cin], kernel [kh,kw,Cin,Cout].  module.py:44-64."""
    kh, kw = kernel.shape[0], kernel.shape[1]
    pt, pb = same_pad(x.shape[1], kh, strides[0])
    pl, pr = same_pad(x.shape[2], kw, strides[1])
    xt = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(xt, kernel.permute(3, 2, 0, 1), bias, stride=tuple(strides))
    return y.permute(0, 2, 3, 1)


def instance_norm(x, beta, gamma, eps=IN_EPS):
    """tf.contrib.layers.instance_norm on channels-last input (Appendix A.4): biased two-pass variance
    over all axes but batch and channel."""
    axes = tuple(range(1, x.dim() - 1))
    mean = x.mean(dim=axes, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=axes, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * gamma + beta


def glu(a, g):
    """module.py:3-7."""
    return a * torch.sigmoid(g)


def pixel_shuffle_reshape(x, r=2):
    """module.py:135-146: raw reshape [n,w,c] -> [n,w*r,c//r]."""
    n, w, c = x.shape
    return x.reshape(n, w * r, c // r)


def _inname(i):
    return "InstanceNorm" if i == 0 else "InstanceNorm_%d" % i


# --------------------------------------------------------------------------------------
# Networks
# --------------------------------------------------------------------------------------

def generator_forward(x, P, scope, taps=None):
    """module.py:148-185.  x [B,24,T] -> [B,24,T].  `taps` (dict) collects layer-boundary activations."""
    def p(n):
        return P[scope + "/" + n]

    def tap(n, v):
        if taps is not None:
            taps[n] = v
        return v

    def gated_in(h, prefix, idx, k, stride, shuffle=False):
        a = conv1d_same(h, p(prefix + "h1_conv/kernel"), p(prefix + "h1_conv/bias"), stride)
        g = conv1d_same(h, p(prefix + "h1_gates/kernel"), p(prefix + "h1_gates/bias"), stride)
        if shuffle:
            a, g = pixel_shuffle_reshape(a), pixel_shuffle_reshape(g)
        a = instance_norm(a, p(_inname(idx) + "/beta"), p(_inname(idx) + "/gamma"))
        g = instance_norm(g, p(_inname(idx + 1) + "/beta"), p(_inname(idx + 1) + "/gamma"))
        return glu(a, g)

    h = x.transpose(1, 2)                                                          # module.py:152
    a = conv1d_same(h, p("h1_conv/kernel"), p("h1_conv/bias"))
    g = conv1d_same(h, p("h1_conv_gates/kernel"), p("h1_conv_gates/bias"))
    h = tap("h1_glu", glu(a, g))                                                   # :161-163
    idx = 0
    for i in (1, 2):                                                               # :166-167
        h = tap("d%d" % i, gated_in(h, "downsample1d_block%d_" % i, idx, 5, 2)); idx += 2
    for i in range(1, 7):                                                          # :170-175
        pre = "residual1d_block%d_" % i
        h1 = gated_in(h, pre, idx, 3, 1)
        h2 = conv1d_same(h1, p(pre + "h2_conv/kernel"), p(pre + "h2_conv/bias"))
        h2 = instance_norm(h2, p(_inname(idx + 2) + "/beta"), p(_inname(idx + 2) + "/gamma"))
        h = tap("r%d" % i, h + h2); idx += 3
    for i in (1, 2):                                                               # :178-179
        h = tap("u%d" % i, gated_in(h, "upsample1d_block%d_" % i, idx, 5, 1, shuffle=True)); idx += 2
    o = conv1d_same(h, p("o1_conv/kernel"), p("o1_conv/bias"))                      # :182
    return tap("out", o.transpose(1, 2))                                           # :183


def discriminator_forward(x, P, scope, taps=None):
    """module.py:188-213.  x [B,24,T] -> [B, 24/4, T/16, 1] sigmoid probabilities."""
    def p(n):
        return P[scope + "/" + n]

    def tap(n, v):
        if taps is not None:
            taps[n] = v
        return v

    h = x.unsqueeze(-1)                                                            # :192
    a = conv2d_same(h, p("h1_conv/kernel"), p("h1_conv/bias"), (1, 2))
    g = conv2d_same(h, p("h1_conv_gates/kernel"), p("h1_conv_gates/bias"), (1, 2))
    h = tap("h1_glu", glu(a, g))                                                   # :201-203
    idx = 0
    for i, st in ((1, (2, 2)), (2, (2, 2)), (3, (1, 2))):                          # :206-208
        pre = "downsample2d_block%d_" % i
        a = conv2d_same(h, p(pre + "h1_conv/kernel"), p(pre + "h1_conv/bias"), st)
        g = conv2d_same(h, p(pre + "h1_gates/kernel"), p(pre + "h1_gates/bias"), st)
        a = instance_norm(a, p(_inname(idx) + "/beta"), p(_inname(idx) + "/gamma"))
        g = instance_norm(g, p(_inname(idx + 1) + "/beta"), p(_inname(idx + 1) + "/gamma"))
        h = tap("d%d" % i, glu(a, g)); idx += 2
    z = h @ p("dense/kernel") + p("dense/bias")                                    # :211
    return tap("out", torch.sigmoid(z))


# --------------------------------------------------------------------------------------
# Losses + training step
# --------------------------------------------------------------------------------------

def l1_loss(y, y_hat):          # utils.py:6-8
    return (y - y_hat).abs().mean()


def l2_loss(y, y_hat):          # utils.py:10-12
    return ((y - y_hat) ** 2).mean()


LOSS_NAMES = ("cycle_loss", "identity_loss", "generator_loss_A2B", "generator_loss_B2A", "generator_loss",
              "discriminator_loss_A", "discriminator_loss_B", "discriminator_loss")     # model.py:153-169


def losses(A, B, P, lambda_cycle, lambda_identity, taps=None):
    """model.py:44-90.  Returns (dict of the 8 logged scalars, generation_A, generation_B).
    The discriminator-loss branch sees the fakes detached (they are fed back through placeholders,
    model.py:118-119)."""
    gen_B = generator_forward(A, P, "generator_A2B")
    cycle_A = generator_forward(gen_B, P, "generator_B2A")
    gen_A = generator_forward(B, P, "generator_B2A")
    cycle_B = generator_forward(gen_A, P, "generator_A2B")
    id_A = generator_forward(A, P, "generator_B2A")
    id_B = generator_forward(B, P, "generator_A2B")
    dA_fake = discriminator_forward(gen_A, P, "discriminator_A")
    dB_fake = discriminator_forward(gen_B, P, "discriminator_B")
    L = {}
    L["cycle_loss"] = l1_loss(A, cycle_A) + l1_loss(B, cycle_B)
    L["identity_loss"] = l1_loss(A, id_A) + l1_loss(B, id_B)
    L["generator_loss_A2B"] = l2_loss(torch.ones_like(dB_fake), dB_fake)
    L["generator_loss_B2A"] = l2_loss(torch.ones_like(dA_fake), dA_fake)
    L["generator_loss"] = (L["generator_loss_A2B"] + L["generator_loss_B2A"]
                           + lambda_cycle * L["cycle_loss"] + lambda_identity * L["identity_loss"])
    fA, fB = gen_A.detach(), gen_B.detach()
    dA_real = discriminator_forward(A, P, "discriminator_A")
    dB_real = discriminator_forward(B, P, "discriminator_B")
    dA_f = discriminator_forward(fA, P, "discriminator_A")
    dB_f = discriminator_forward(fB, P, "discriminator_B")
    L["discriminator_loss_A"] = (l2_loss(torch.ones_like(dA_real), dA_real) + l2_loss(torch.zeros_like(dA_f), dA_f)) / 2
    L["discriminator_loss_B"] = (l2_loss(torch.ones_like(dB_real), dB_real) + l2_loss(torch.zeros_like(dB_f), dB_f)) / 2
    L["discriminator_loss"] = L["discriminator_loss_A"] + L["discriminator_loss_B"]
    if taps is not None:
        taps.update(cycle_A=cycle_A, cycle_B=cycle_B, id_A=id_A, id_B=id_B, dA_fake=dA_fake, dB_fake=dB_fake,
                    dA_real=dA_real, dB_real=dB_real)
    return L, gen_A, gen_B


def gradients(A, B, P, lambda_cycle, lambda_identity):
    """Gradients exactly as the two `minimize` calls see them (model.py:107-108): generator variables
    w.r.t. generator_loss, discriminator variables w.r.t. discriminator_loss, both at the same weights."""
    Pg = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in P.items())
    L, gen_A, gen_B = losses(A, B, Pg, lambda_cycle, lambda_identity)
    gnames = [k for k in Pg if "generator" in k]          # model.py:94-95
    dnames = [k for k in Pg if "discriminator" in k]
    gg = torch.autograd.grad(L["generator_loss"], [Pg[k] for k in gnames], retain_graph=True)
    dg = torch.autograd.grad(L["discriminator_loss"], [Pg[k] for k in dnames])
    G = OrderedDict()
    for k, g in zip(gnames, gg):
        G[k] = g
    for k, g in zip(dnames, dg):
        G[k] = g
    G = OrderedDict((k, G[k]) for k in P)      # arena order
    return {k: v.detach() for k, v in L.items()}, G, gen_A.detach(), gen_B.detach()


class TFAdam:
    """tf.train.AdamOptimizer (Appendix A.6): eps is added to sqrt(v) un-bias-corrected."""

    def __init__(self, P):
        self.t = 0
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())

    def apply(self, P, G, lr_g, lr_d):
        self.t += 1
        for k in P:
            lr = lr_g if "generator" in k else lr_d
            lr_t = lr * math.sqrt(1.0 - ADAM_BETA2 ** self.t) / (1.0 - ADAM_BETA1 ** self.t)
            self.m[k].mul_(ADAM_BETA1).add_(G[k], alpha=1.0 - ADAM_BETA1)
            self.v[k].mul_(ADAM_BETA2).addcmul_(G[k], G[k], value=1.0 - ADAM_BETA2)
            P[k] = P[k] - lr_t * self.m[k] / (self.v[k].sqrt() + ADAM_EPS)


class OracleCycleGAN:
    """CPU restatement of model.py:7-150 with the reference's method signatures."""

    def __init__(self, num_features=NUM_FEATURES, seed=0, dtype=torch.float32, params=None):
        self.num_features = num_features
        self.dtype = dtype
        self.P = params if params is not None else init_params(seed, dtype, num_features)
        self.opt = TFAdam(self.P)
        self.train_step = 0
        self.last_losses = None

    def train(self, input_A, input_B, lambda_cycle, lambda_identity, generator_learning_rate, discriminator_learning_rate):
        A = torch.as_tensor(np.asarray(input_A), dtype=self.dtype)
        B = torch.as_tensor(np.asarray(input_B), dtype=self.dtype)
        L, G, _, _ = gradients(A, B, self.P, lambda_cycle, lambda_identity)
        self.opt.apply(self.P, G, generator_learning_rate, discriminator_learning_rate)
        self.train_step += 1
        self.last_losses = {k: float(v) for k, v in L.items()}
        return np.float32(L["generator_loss"]), np.float32(L["discriminator_loss"])

    def test(self, inputs, direction):
        x = torch.as_tensor(np.asarray(inputs), dtype=self.dtype)
        with torch.no_grad():
            if direction == "A2B":
                y = generator_forward(x, self.P, "generator_A2B")
            elif direction == "B2A":
                y = generator_forward(x, self.P, "generator_B2A")
            else:
                raise Exception("Conversion direction must be specified.")      # model.py:135
        return y.to(torch.float32).numpy()
