"""CPU checks of the algebra behind `edge_lower` (engine.cu edge_on, simt_kernels.cu im2col_taps_kernel / col2im_taps_kernel,
tc_gemm.cu w_src / tn_dst): the generator's two 15-tap layers with 24 channels on one side (module.py:85-86 h1, module.py:148 o1)
are computed on the GPU as dense 1 x 1 GEMMs over an im2col of the 24-channel tensor.  The index conventions the kernels implement
(dir = +1 / -1, pad_left = 7, zero outside the sample, TF kernel [15,24,128] read as a [360,128] matrix, o1's taps folded into
output columns t * 24 + n) are restated here in numpy and held against the oracle's TF-'SAME' convolution and its autograd
gradients, so a sign or offset error in the lowering would show without a GPU.  (The CUDA kernels themselves are compared with the
oracle and with the 15-tap gather-GEMMs in tests/test_gpu_model.py.)"""
import numpy as np
import torch

from oracle import cyclegan_oracle as O

KW, F_, PL = 15, 24, 7          # taps, narrow channel count, TF SAME pad_left at stride 1 = (kw - 1) // 2


def im2col_taps(x, direction):
    """x [n, T, C] -> [n, T, KW * C]: out[m, t*C + c] = x[m + direction * (t - PL), c], zero outside the sample (im2col_taps_kernel)."""
    n, T, C = x.shape
    out = np.zeros((n, T, KW * C), dtype=x.dtype)
    for t in range(KW):
        s = direction * (t - PL)
        lo, hi = max(0, -s), min(T, T - s)
        out[:, lo:hi, t * C:(t + 1) * C] = x[:, lo + s:hi + s, :]
    return out


def col2im_taps(z, C, direction, bias=None):
    """z [n, T, KW * C] -> [n, T, C]: y[m, c] = bias[c] + sum_t z[m + direction * (t - PL), t*C + c] over rows of the sample (col2im_taps_kernel)."""
    n, T, _ = z.shape
    y = np.zeros((n, T, C), dtype=z.dtype)
    for t in range(KW):
        s = direction * (t - PL)
        lo, hi = max(0, -s), min(T, T - s)
        y[:, lo:hi, :] += z[:, lo + s:hi + s, t * C:(t + 1) * C]
    return y if bias is None else y + bias


def fold_columns(w):
    """TF kernel [KW, Cin, Cout] -> [Cin, KW * Cout] with column t * Cout + n (TcLayer::fold / w_src)."""
    return np.concatenate([w[t] for t in range(KW)], axis=1)


def test_pad_left_matches_tf_same():
    for T in (36, 128, 516):
        assert O.same_pad(T, KW, 1) == (PL, KW - 1 - PL)


def test_h1_is_a_dense_gemm_over_the_im2col_of_the_input():
    rs = np.random.RandomState(0)
    n, T, Cout = 3, 36, 16
    x = rs.randn(n, T, F_); w = rs.randn(KW, F_, Cout); b = rs.randn(Cout); dy = rs.randn(n, T, Cout)
    xt = torch.tensor(x, requires_grad=True); wt = torch.tensor(w, requires_grad=True)
    y_ref = O.conv1d_same(xt, wt, torch.tensor(b))
    y_ref.backward(torch.tensor(dy))
    xcol = im2col_taps(x, +1)
    w2 = w.reshape(KW * F_, Cout)                       # TF's [15,24,Cout] kernel as it lies in memory
    # forward
    assert np.allclose(xcol @ w2 + b, y_ref.detach().numpy(), atol=1e-10)
    # weight gradient lands in the same [15,24,Cout] memory
    dw = np.einsum('ntk,ntc->kc', xcol, dy).reshape(KW, F_, Cout)
    assert np.allclose(dw, wt.grad.numpy(), atol=1e-10)
    # data gradient: dense dP . W^T, then the tap-shifted sum with the opposite direction
    dx = col2im_taps(dy @ w2.T, F_, -1)
    assert np.allclose(dx, xt.grad.numpy(), atol=1e-10)


def test_o1_is_a_dense_gemm_with_folded_taps_plus_a_tap_shifted_sum():
    rs = np.random.RandomState(1)
    n, T, Cin = 2, 64, 20
    u = rs.randn(n, T, Cin); w = rs.randn(KW, Cin, F_); b = rs.randn(F_); dout = rs.randn(n, T, F_)
    ut = torch.tensor(u, requires_grad=True); wt = torch.tensor(w, requires_grad=True)
    y_ref = O.conv1d_same(ut, wt, torch.tensor(b))
    y_ref.backward(torch.tensor(dout))
    wf = fold_columns(w)                                # [Cin, 15 * 24]
    # forward: Z = U . W', out[m, c] = b[c] + sum_t Z[m + t - 7, (t, c)]
    z = u @ wf
    assert np.allclose(col2im_taps(z, F_, +1, b), y_ref.detach().numpy(), atol=1e-10)
    # backward: dZ = im2col(d_out) with the opposite direction; dense data and weight gradients
    dz = im2col_taps(dout, -1)
    assert np.allclose(dz @ wf.T, ut.grad.numpy(), atol=1e-10)
    dwf = np.einsum('ntc,ntk->ck', u, dz)               # [Cin, (t, n)]; tn_dst scatters column t * 24 + n to kernel element [t][c][n]
    dw = np.stack([dwf[:, t * F_:(t + 1) * F_] for t in range(KW)], axis=0)
    assert np.allclose(dw, wt.grad.numpy(), atol=1e-10)


def test_fold_index_functions():
    """w_src / tn_dst of tc_gemm.cu restated: column co of the folded layer <-> element (t, ci, n) of the [KW, Cin, 24] kernel."""
    Cin, fold_n = 8, F_
    w = np.arange(KW * Cin * fold_n, dtype=np.int64).reshape(KW, Cin, fold_n)
    flat = w.ravel()
    for co in (0, 5, 23, 24, 100, KW * fold_n - 1):
        for ci in (0, 3, Cin - 1):
            t = co // fold_n
            assert flat[(t * Cin + ci) * fold_n + (co - t * fold_n)] == w[t, ci, co % fold_n]
