"""The callers either side of the hot path: schedule and sampler of train.py / preprocess.py."""
import importlib
import os

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


def test_counter_sampler_contract():
    """The counter-based twin (host side of the device sampler) obeys the same contract as preprocess.py:207-238: independent
    shuffles of both lists, truncation to the shorter one, one in-bounds crop per utterance; keyed by (seed, epoch)."""
    import cgvc  # noqa: F401
    P = importlib.import_module("cgvc.preprocess")
    rs = np.random.RandomState(5)
    lens_A = list(rs.randint(128, 700, size=37)); lens_B = list(rs.randint(128, 700, size=23)); lens_A[3] = 128
    plans = {}
    for seed in (0, 11):
        for epoch in (0, 1, 2):
            ua, sa, ub, sb = P.counter_sample_plan(lens_A, lens_B, seed, epoch)
            assert len(ua) == len(ub) == len(sa) == len(sb) == 23
            assert len(set(ua.tolist())) == 23 and len(set(ub.tolist())) == 23 and sorted(ub.tolist()) == list(range(23))
            assert all(0 <= s <= lens_A[u] - 128 for u, s in zip(ua, sa)) and all(0 <= s <= lens_B[u] - 128 for u, s in zip(ub, sb))
            plans[(seed, epoch)] = (ua.tolist(), sa.tolist(), ub.tolist(), sb.tolist())
            assert plans[(seed, epoch)] == tuple(x.tolist() for x in P.counter_sample_plan(lens_A, lens_B, seed, epoch))   # reproducible
    assert len({str(v) for v in plans.values()}) == 6                                  # every (seed, epoch) draws its own epoch
    # over many epochs every utterance of the longer list is used, first positions and crop starts are spread uniformly
    first = np.zeros(37); starts = []
    for epoch in range(400):
        ua, sa, _, _ = P.counter_sample_plan(lens_A, lens_B, 3, epoch)
        first[ua[0]] += 1; starts += [s / max(lens_A[u] - 128, 1) for u, s in zip(ua, sa) if lens_A[u] > 256]
    assert first.min() > 0 and first.max() < 30 and abs(np.mean(starts) - 0.5) < 0.02
    A = [np.tile(np.arange(n)[None, :] + 1000 * i, (24, 1)).astype(float) for i, n in enumerate(lens_A)]
    B = [np.tile(np.arange(n)[None, :] + 1000 * i, (24, 1)).astype(float) for i, n in enumerate(lens_B)]
    a, b = P.sample_train_data_counter(A, B, seed=11, epoch=1)
    ua, sa, ub, sb = P.counter_sample_plan(lens_A, lens_B, 11, 1)
    assert a.shape == (23, 24, 128) and np.array_equal(a[:, 0, 0], 1000 * ua + sa) and np.array_equal(b[:, 5, 127], 1000 * ub + sb + 127)
    with pytest.raises(AssertionError):
        P.counter_sample_plan([100], [300], 0, 0)                                      # utterance shorter than the crop


def test_normalization_fit():
    T = _drv()
    rs = np.random.RandomState(1)
    sps = [rs.randn(24, n) * 3 + 7 for n in (130, 260)]
    norm, mean, std = T.fit_normalization(sps)
    cat = np.concatenate(norm, axis=1)
    assert mean.shape == (24, 1) and np.allclose(cat.mean(axis=1), 0, atol=1e-12) and np.allclose(cat.std(axis=1), 1, atol=1e-12)


@pytest.mark.gpu
def test_device_sampler_matches_host_twin_index_for_index():
    """cgvc_sample_plan / cgvc_gather_minibatch (device-resident corpus) against preprocess.counter_sample_plan /
    sample_train_data_counter: the same utterance order, the same crop starts, bit-identical minibatches."""
    import cgvc
    T = _drv()
    P = importlib.import_module("cgvc.preprocess")
    m = cgvc.CycleGAN(num_features=24, mode='train', max_batch=8, max_frames=128, precision="bf16x3", log_dir='/tmp/cgvc_log')
    rs = np.random.RandomState(9)
    for nA, nB, batch in ((13, 9, 4), (5, 40, 2), (700, 650, 8)):
        A = [rs.randn(24, n).astype(np.float32).astype(np.float64) for n in rs.randint(128, 520, size=nA)]
        B = [rs.randn(24, n).astype(np.float32).astype(np.float64) for n in rs.randint(128, 520, size=nB)]
        A[0] = A[0][:, :128]                                                            # exactly one crop position
        ds = T.DeviceDataset(m, A, B, batch, 128, seed=1234 + nA)
        assert ds.num_pairs == min(nA, nB) and ds.iterations_per_epoch() == min(nA, nB) // batch
        for epoch in (0, 1, 7):
            ds.plan(epoch)
            want = P.counter_sample_plan([a.shape[1] for a in A], [b.shape[1] for b in B], 1234 + nA, epoch)
            got = ds.plan_host()
            for g, w in zip(got, want):
                assert np.array_equal(g, w), (nA, nB, epoch)
            ha, hb = P.sample_train_data_counter(A, B, 1234 + nA, epoch)
            for i in (0, ds.iterations_per_epoch() - 1):
                a_dev, b_dev = ds.minibatch(i)
                assert np.array_equal(a_dev.cpu().numpy(), ha[i * batch:(i + 1) * batch].astype(np.float32))
                assert np.array_equal(b_dev.cpu().numpy(), hb[i * batch:(i + 1) * batch].astype(np.float32))
    short = T.DeviceDataset(m, [np.zeros((24, 127))] + A[:3], B[:4], 1, 128, seed=0)
    with pytest.raises(AssertionError, match="shorter than the 128-frame crop"):
        short.plan(0)
    with pytest.raises(Exception, match="outside the epoch"):
        ds.minibatch(ds.num_pairs)


@pytest.mark.gpu
def test_device_resident_loop_equals_host_fed_loop(tmp_path):
    """Training from the device-resident corpus is the same computation as feeding the same minibatches from the host: identical
    losses step by step (train_async on gathered device tensors vs train() on the host twin's crops)."""
    import cgvc
    T = _drv()
    P = importlib.import_module("cgvc.preprocess")
    A, B = T.synthetic_speaker(6, 1), T.synthetic_speaker(5, 2)
    A, _, _ = T.fit_normalization(A); B, _, _ = T.fit_normalization(B)
    ms = [cgvc.CycleGAN(num_features=24, mode='train', max_batch=2, max_frames=128, precision="fp32", seed=3, log_dir='/tmp/cgvc_log') for _ in range(2)]
    ds = T.DeviceDataset(ms[0], A, B, 2, 128, seed=3)
    for epoch in range(2):
        ds.plan(epoch)
        ha, hb = P.sample_train_data_counter(A, B, 3, epoch)
        for i in range(ds.iterations_per_epoch()):
            a_dev, b_dev = ds.minibatch(i)
            ms[0].train_async(a_dev, b_dev, 10, 5, 2e-4, 1e-4)
            g0, d0 = ms[0].fetch_losses()
            g1, d1 = ms[1].train(ha[2 * i:2 * i + 2], hb[2 * i:2 * i + 2], 10, 5, 2e-4, 1e-4)
            # same inputs, same kernels; two runs differ only through the order of their gradient atomics
            assert abs(g0 - g1) <= 1e-3 * abs(g1) and abs(d0 - d1) <= 1e-3 * abs(d1), (epoch, i, g0, g1, d0, d1)


@pytest.mark.gpu
def test_train_driver_runs(tmp_path):
    T = _drv()
    model, g, d = T.train(None, None, str(tmp_path / "m"), "x.ckpt", 0, num_epochs=2, mini_batch_size=2, synthetic=5, log_every=1)
    assert model.train_step == 4 and np.isfinite(g) and np.isfinite(d)          # 5 utterances // batch 2 = 2 iterations per epoch
    z = np.load(str(tmp_path / "m" / "mcep_normalization.npz"))
    assert set(z.files) == {"mean_A", "std_A", "mean_B", "std_B"}               # train.py:57 / convert.py:18-22
    assert (tmp_path / "m" / "x.ckpt.npz").exists()
    # the host-fed loop (the reference's feed_dict path) still works
    model2, g2, d2 = T.train(None, None, str(tmp_path / "m2"), "x.ckpt", 0, num_epochs=1, mini_batch_size=2, synthetic=5, log_every=1, device_data=False)
    assert model2.train_step == 2 and np.isfinite(g2) and np.isfinite(d2)


def test_train_loop_with_validation_on_a_stub_model(tmp_path, monkeypatch, capsys):
    """The whole driver on the CPU with a stand-in for the engine: schedule, host sampler, per-epoch save, side files, and the
    validation conversions of train.py:119-155 at epoch 0 (epoch % 50 == 0) through a forward-only model that gets the weights."""
    import cgvc  # noqa: F401
    T = _drv()
    M = importlib.import_module("cgvc.model")
    made = []

    class Stub:
        def __init__(self, num_features, mode='train', **kw):
            self.mode = mode; self.kw = kw; self.train_step = 0; self.calls = []; self.saved = []; self.params = {"w": np.zeros(1)}
            made.append(self)

        def train(self, input_A, input_B, lambda_cycle, lambda_identity, generator_learning_rate, discriminator_learning_rate):
            assert input_A.shape == input_B.shape == (2, 24, 128) and lambda_cycle == 10
            self.train_step += 1; self.params = {"w": np.full(1, float(self.train_step))}
            self.calls.append((lambda_identity, generator_learning_rate, discriminator_learning_rate))
            return np.float32(1.0 / self.train_step), np.float32(0.5)

        def save(self, directory, filename):
            os.makedirs(directory, exist_ok=True); self.saved.append(os.path.join(directory, filename)); return self.saved[-1]

        def get_params(self):
            return dict(self.params)

        def set_params(self, p):
            self.params = dict(p)

        def test(self, inputs, direction):
            return (inputs + self.params["w"][0]).astype(np.float32)

    monkeypatch.setattr(M, "CycleGAN", Stub)
    val = tmp_path / "val_A"; val.mkdir()
    np.savez(str(val / "u.npz"), f0=np.zeros(140), coded_sp=np.zeros((140, 24)))
    model, g, d = T.train(None, None, str(tmp_path / "m"), "x.ckpt", 0, num_epochs=2, mini_batch_size=2, synthetic=5, log_every=1,
                          device_data=False, validation_A_dir=str(val), validation_B_dir=None, output_dir=str(tmp_path / "out"),
                          tensorboard_log_dir=str(tmp_path / "tb"))
    assert model is made[0] and model.mode == 'train' and model.train_step == 4 and model.kw["log_dir"] == str(tmp_path / "tb")
    assert model.calls[0] == (5, 0.0002, 0.0001) and len(model.saved) == 2
    assert len(made) == 2 and made[1].mode == 'test'                       # the forward-only model of the validation conversions
    assert made[1].params["w"][0] == 2.0                                    # the weights after epoch 0 (2 iterations), not later ones
    z = np.load(str(tmp_path / "out" / "converted_A" / "u.npz"))
    st = np.load(str(tmp_path / "m" / "mcep_normalization.npz"))
    want = (((0.0 - st["mean_A"]) / st["std_A"] + 2.0).astype(np.float32).astype(np.float64) * st["std_B"] + st["mean_B"]).T
    assert z["coded_sp"].shape == (140, 24) and np.allclose(z["coded_sp"], np.broadcast_to(want, (140, 24)), atol=1e-6)
    assert "Generating Validation Data B from A..." in capsys.readouterr().out
