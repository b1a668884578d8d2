"""The callers either side of the hot path: schedule and sampler of train.py / preprocess.py."""
import importlib

import numpy as np
import pytest


def _drv():
    import cgvc  # noqa: F401
    return importlib.import_module("cgvc.train")


def test_schedule_replays_reference_loop():
    """Replay train.py:96-102 literally and compare with the closed form."""
    T = _drv()
    lam, lr_g, lr_d = 5, 0.0002, 0.0001
    dec_g, dec_d = lr_g / 200000, lr_d / 200000
    checkpoints = {0, 1, 9999, 10000, 10001, 200000, 200001, 250000, 399999, 400000, 400050}
    for n in range(0, 400051):
        if n > 10000:
            lam = 0
        if n > 200000:
            lr_g = max(0, lr_g - dec_g); lr_d = max(0, lr_d - dec_d)
        if n in checkpoints:
            a, b, c = T.schedule(n)
            assert a == lam and abs(b - lr_g) < 1e-12 and abs(c - lr_d) < 1e-12, n
    assert T.schedule(400050) == (0, 0.0, 0.0)


def test_sampler_contract():
    """preprocess.py:207-238: independent shuffles, truncation to the shorter list, one 128-frame crop per utterance."""
    T = _drv()
    rs = np.random.RandomState(0)
    A = [np.tile(np.arange(n)[None, :] + 1000 * i, (24, 1)).astype(float) for i, n in enumerate([128, 200, 333, 150, 129])]
    B = [np.tile(np.arange(n)[None, :] + 1000 * i, (24, 1)).astype(float) for i, n in enumerate([140, 128, 500])]
    a, b = T.sample_train_data(A, B, 128, rng=rs)
    assert a.shape == (3, 24, 128) and b.shape == (3, 24, 128)
    for arr, src in ((a, A), (b, B)):
        seen = set()
        for s in arr:
            utt = int(s[0, 0]) // 1000; start = int(s[0, 0]) % 1000
            assert utt not in seen; seen.add(utt)
            assert np.array_equal(s[0], np.arange(start, start + 128) + 1000 * utt)       # a contiguous crop
            assert start + 128 <= src[utt].shape[1]
    with pytest.raises(AssertionError):
        T.sample_train_data([np.zeros((24, 100))], B, 128, rng=rs)                       # utterance shorter than the crop


def test_normalization_fit():
    T = _drv()
    rs = np.random.RandomState(1)
    sps = [rs.randn(24, n) * 3 + 7 for n in (130, 260)]
    norm, mean, std = T.fit_normalization(sps)
    cat = np.concatenate(norm, axis=1)
    assert mean.shape == (24, 1) and np.allclose(cat.mean(axis=1), 0, atol=1e-12) and np.allclose(cat.std(axis=1), 1, atol=1e-12)


@pytest.mark.gpu
def test_train_driver_runs(tmp_path):
    T = _drv()
    model, g, d = T.train(None, None, str(tmp_path / "m"), "x.ckpt", 0, num_epochs=2, mini_batch_size=2, synthetic=5, log_every=1)
    assert model.train_step == 4 and np.isfinite(g) and np.isfinite(d)          # 5 utterances // batch 2 = 2 iterations per epoch
    z = np.load(str(tmp_path / "m" / "mcep_normalization.npz"))
    assert set(z.files) == {"mean_A", "std_A", "mean_B", "std_B"}               # train.py:57 / convert.py:18-22
    assert (tmp_path / "m" / "x.ckpt.npz").exists()
