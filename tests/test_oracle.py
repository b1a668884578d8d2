"""CPU tests of the oracle (test infrastructure): two independent restatements agree, hand-derived backward
formulas match autograd, structural known-answers derived from the reference text hold, and the
committed golden vectors reproduce.  No GPU needed."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import cyclegan_oracle as O
from oracle import numpy_ref as NR
from parity_util import rel_l2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cyclegan_golden.npz")


@pytest.fixture(scope="module")
def P64():
    return O.init_params(seed=1234, dtype=torch.float64, perturb_affine=True)


def test_parameter_counts_and_names():
    g = O.generator_param_specs(); d = O.discriminator_param_specs()
    assert len(g) == 110 and sum(int(np.prod(s)) for _, s, _ in g) == 38055704
    assert len(d) == 30 and sum(int(np.prod(s)) for _, s, _ in d) == 21837825
    names = [n for n, _, _ in O.param_specs()]
    assert len(names) == 280 and len(set(names)) == 280
    assert "generator_A2B/residual1d_block3_h1_conv/kernel" in names                  # Appendix A.5
    assert "generator_B2A/InstanceNorm_25/gamma" in names and "discriminator_A/InstanceNorm_5/beta" in names
    assert "discriminator_B/dense/kernel" in names
    shapes = {n: s for n, s, _ in O.param_specs()}
    assert shapes["generator_A2B/upsample1d_block1_h1_conv/kernel"] == (5, 512, 1024)
    assert shapes["generator_A2B/InstanceNorm_22/gamma"] == (512,)                     # IN after the shuffle (module.py:124-125)
    assert shapes["discriminator_A/downsample2d_block3_h1_conv/kernel"] == (6, 3, 512, 1024)


def test_same_padding_table():
    # Appendix A.1
    assert O.same_pad(128, 15, 1) == (7, 7) and O.same_pad(128, 5, 2) == (1, 2) and O.same_pad(64, 5, 2) == (1, 2)
    assert O.same_pad(32, 3, 1) == (1, 1) and O.same_pad(32, 5, 1) == (2, 2)
    assert O.same_pad(24, 3, 1) == (1, 1) and O.same_pad(128, 3, 2) == (0, 1) and O.same_pad(24, 3, 2) == (0, 1)
    assert O.same_pad(6, 6, 1) == (2, 3)


def test_shapes_and_init_losses():
    P = O.init_params(seed=0, dtype=torch.float32)
    A, B = O.synthetic_batch(0, 2, 128)
    with torch.no_grad():
        y = O.generator_forward(A, P, "generator_A2B")
        d = O.discriminator_forward(A, P, "discriminator_A")
        y2 = O.generator_forward(A[:, :, :36], P, "generator_B2A")
        L, _, _ = O.losses(A, B, P, 10.0, 5.0)
    assert tuple(y.shape) == (2, 24, 128) and tuple(d.shape) == (2, 6, 8, 1) and tuple(y2.shape) == (2, 24, 36)
    # at glorot init the discriminators output ~0.5
    assert abs(float(L["discriminator_loss"]) - 0.5) < 0.1
    assert abs(float(L["generator_loss_A2B"]) - 0.25) < 0.06 and abs(float(L["generator_loss_B2A"]) - 0.25) < 0.06
    assert 25.0 < float(L["generator_loss"]) < 35.0


def test_numpy_and_torch_restatements_agree(P64):
    A, _ = O.synthetic_batch(3, 2, 64, dtype=torch.float64)
    Pn = {k: v.numpy() for k, v in P64.items()}
    with torch.no_grad():
        yt = O.generator_forward(A, P64, "generator_B2A").numpy()
        A2, _ = O.synthetic_batch(4, 1, 128, dtype=torch.float64)
        dt = O.discriminator_forward(A2, P64, "discriminator_B").numpy()
    assert rel_l2(NR.generator_forward(A.numpy(), Pn, "generator_B2A"), yt) < 1e-12
    assert rel_l2(NR.discriminator_forward(A2.numpy(), Pn, "discriminator_B"), dt) < 1e-12


@pytest.mark.parametrize("geom", [(1, 16, 8, 4, 1, 5, 6, 1, 2), (2, 6, 8, 3, 6, 3, 5, 1, 2), (2, 8, 8, 3, 3, 3, 4, 2, 2), (1, 1, 9, 4, 1, 3, 4, 1, 1)])
def test_conv_backward_formulas(geom):
    """numpy_ref.conv_bwd (the gather-GEMM statement the kernels implement) == autograd of the torch oracle."""
    B, H, W, Cin, kh, kw, Cout, sh, sw = geom
    rs = np.random.RandomState(0)
    x = rs.randn(B, H, W, Cin); w = rs.randn(kh, kw, Cin, Cout); b = rs.randn(Cout)
    xt, wt, bt = (torch.tensor(a, requires_grad=True) for a in (x, w, b))
    y = O.conv2d_same(xt, wt, bt, (sh, sw))
    dy = rs.randn(*y.shape)
    y.backward(torch.tensor(dy))
    assert rel_l2(NR.conv_fwd(x, w, b, (sh, sw)), y.detach().numpy()) < 1e-13
    dx, dw, db = NR.conv_bwd(x, w, dy, (sh, sw))
    assert rel_l2(dx, xt.grad.numpy()) < 1e-13 and rel_l2(dw, wt.grad.numpy()) < 1e-13 and rel_l2(db, bt.grad.numpy()) < 1e-13


def test_instance_norm_glu_backward_formulas():
    rs = np.random.RandomState(1)
    a = rs.randn(2, 1, 12, 6) * 2 + 0.5; g = rs.randn(2, 1, 12, 6)
    ba, ga, bg, gg = rs.randn(6), rs.rand(6) + 0.5, rs.randn(6), rs.rand(6) + 0.5
    ts = [torch.tensor(v, requires_grad=True) for v in (a, g, ba, ga, bg, gg)]
    y = O.glu(O.instance_norm(ts[0], ts[2], ts[3]), O.instance_norm(ts[1], ts[4], ts[5]))
    dy = rs.randn(*y.shape)
    y.backward(torch.tensor(dy))
    na, ca = NR.in_fwd(a, ba, ga); ng, cg = NR.in_fwd(g, bg, gg)
    yn, cglu = NR.glu_fwd(na, ng)
    assert rel_l2(yn, y.detach().numpy()) < 1e-13
    dna, dng = NR.glu_bwd(dy, cglu)
    da, dga, dba = NR.in_bwd(dna, ca, ga); dg, dgg, dbg = NR.in_bwd(dng, cg, gg)
    for got, ref in ((da, ts[0].grad), (dg, ts[1].grad), (dba, ts[2].grad), (dga, ts[3].grad), (dbg, ts[4].grad), (dgg, ts[5].grad)):
        assert rel_l2(got, ref.numpy()) < 1e-12


def test_pixel_shuffle_is_a_raw_reshape():
    x = torch.arange(2 * 3 * 8, dtype=torch.float64).reshape(2, 3, 8)
    y = O.pixel_shuffle_reshape(x)
    # out[n, 2w+s, o] = in[n, w, s*(c/2)+o]   (Appendix A.8)
    for w in range(3):
        for s in range(2):
            assert torch.equal(y[:, 2 * w + s, :], x[:, w, s * 4:(s + 1) * 4])


def test_tf_adam_formula():
    """Appendix A.6: eps is added to sqrt(v) (not bias-corrected), lr_t = lr*sqrt(1-b2^t)/(1-b1^t)."""
    P = {"generator_x": torch.tensor([1.0, -2.0], dtype=torch.float64), "discriminator_y": torch.tensor([0.5], dtype=torch.float64)}
    opt = O.TFAdam(P)
    G = {"generator_x": torch.tensor([0.1, -0.3], dtype=torch.float64), "discriminator_y": torch.tensor([2.0], dtype=torch.float64)}
    opt.apply(P, G, 2e-4, 1e-4)
    lr_t = 2e-4 * math.sqrt(1 - 0.999) / (1 - 0.5)
    m, v = 0.5 * 0.1, 0.001 * 0.01
    assert abs(float(P["generator_x"][0]) - (1.0 - lr_t * m / (math.sqrt(v) + 1e-8))) < 1e-15
    lr_d = 1e-4 * math.sqrt(1 - 0.999) / (1 - 0.5)
    assert abs(float(P["discriminator_y"][0]) - (0.5 - lr_d * (0.5 * 2.0) / (math.sqrt(0.001 * 4.0) + 1e-8))) < 1e-15


def test_hash_rng_is_stable():
    u = O.hash_uniform(7, 3, 5)
    assert np.all((u >= 0) & (u < 1))
    assert np.allclose(u, O.hash_uniform(7, 3, 5)) and not np.allclose(u, O.hash_uniform(7, 4, 5))
    z = O.hash_normal(0, 1, 200000)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01


def test_golden_vectors_reproduce(P64):
    """fp32 oracle vs the committed float64 golden vectors (tests/golden/make_golden.py)."""
    Z = np.load(GOLD)
    P32 = {k: v.float() for k, v in P64.items()}
    A, B = O.synthetic_batch(int(Z["seed_x"]), 1, 128)
    with torch.no_grad():
        y = O.generator_forward(A, P32, "generator_A2B").numpy()
        d = O.discriminator_forward(A, P32, "discriminator_A").numpy()
    assert rel_l2(y, Z["gen_A2B_out"]) < 2e-5 and rel_l2(d, Z["disc_A_out"]) < 2e-5
    L, G, gA, gB = O.gradients(A, B, P32, 10.0, 5.0)
    got = np.array([float(L[k]) for k in O.LOSS_NAMES])
    assert np.allclose(got, Z["losses"], rtol=2e-5)
    assert rel_l2(gA.numpy(), Z["generation_A"]) < 2e-5
    norms = np.array([float(G[k].norm()) for k in G])
    big = Z["grad_norms"] > 1e-6
    assert np.allclose(norms[big], Z["grad_norms"][big], rtol=2e-3)
    k = "generator_B2A/residual1d_block3_h2_conv/kernel"
    assert rel_l2(G[k].numpy().reshape(-1)[:64], Z["grad_first64/" + k]) < 2e-3


def test_cycle_direction_and_reference_api():
    m = O.OracleCycleGAN(seed=0)
    with pytest.raises(Exception, match="Conversion direction must be specified."):
        m.test(np.zeros((1, 24, 16)), "sideways")
