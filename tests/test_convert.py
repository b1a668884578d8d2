"""The callers after the hot path: feature helpers of preprocess.py and the convert.py driver."""
import importlib
import os

import numpy as np
import pytest


def _mod(name):
    import cgvc  # noqa: F401
    return importlib.import_module("cgvc." + name)


# ----------------------------------------------------------------------------------------------- preprocess helpers (CPU)
def test_coded_sp_padding_matches_reference_rule():
    """preprocess.py:135-146: pad to a multiple of 4, floor(diff/2) zeros in front, the rest behind."""
    P = _mod("preprocess")
    for frames in (1, 4, 5, 6, 7, 127, 128, 129, 130, 131):
        x = np.arange(24 * frames, dtype=float).reshape(24, frames) + 1
        y = P.coded_sp_padding(x, multiple=4)
        want = int(np.ceil(frames / 4)) * 4
        diff = want - frames
        assert y.shape == (24, want)
        assert np.array_equal(y[:, diff // 2: diff // 2 + frames], x)
        assert not y[:, :diff // 2].any() and not y[:, diff // 2 + frames:].any()


def test_wav_padding_gives_frame_multiple():
    """preprocess.py:148-158: after padding, WORLD's frame count floor(n / hop) + 1 is a multiple of 4."""
    P = _mod("preprocess")
    sr, fp = 16000, 5.0
    hop = sr * fp / 1000
    for n in (1, 79, 80, 81, 16000, 16001, 40123, 55555):
        w = P.wav_padding(np.ones(n), sr, fp, multiple=4)
        assert (int(np.floor(len(w) / hop)) + 1) % 4 == 0 and len(w) >= n
        left = (len(w) - n) // 2
        assert w[left:left + n].all() and not w[:left].any() and not w[left + n:].any()


def test_pitch_conversion_and_logf0_statistics():
    """preprocess.py:161-175: masked log statistics over voiced frames; log-Gaussian mapping; unvoiced stays 0."""
    P = _mod("preprocess")
    rs = np.random.RandomState(0)
    f0s = [np.where(rs.rand(n) < 0.3, 0.0, np.exp(rs.randn(n) * 0.2 + 5.0)) for n in (300, 500)]
    m, s = P.logf0_statistics(f0s)
    voiced = np.concatenate(f0s); voiced = voiced[voiced > 0]
    assert abs(m - np.log(voiced).mean()) < 1e-12 and abs(s - np.log(voiced).std()) < 1e-12
    f0 = f0s[0]
    out = P.pitch_conversion(f0, m, s, 4.5, 0.3)
    assert np.all(out[f0 == 0] == 0)
    v = f0 > 0
    assert np.allclose(np.log(out[v]), (np.log(f0[v]) - m) / s * 0.3 + 4.5, atol=1e-12)
    # identity mapping when source and target statistics agree
    assert np.allclose(P.pitch_conversion(f0, m, s, m, s)[v], f0[v], rtol=1e-12)


def test_normalization_roundtrip_and_names():
    P = _mod("preprocess")
    rs = np.random.RandomState(1)
    sps = [rs.randn(24, n) * 2 + 3 for n in (130, 170, 260)]
    norm, mean, std = P.coded_sps_normalization_fit_transoform(sps)
    cat = np.concatenate(norm, axis=1)
    assert np.allclose(cat.mean(axis=1), 0, atol=1e-12) and np.allclose(cat.std(axis=1), 1, atol=1e-12)
    again = P.coded_sps_normalization_transoform(sps, mean, std)
    back = P.coded_sps_normalization_inverse_transoform(norm, mean, std)
    for a, b, c, d in zip(norm, again, back, sps):
        assert np.allclose(a, b) and np.allclose(c, d)
    assert [a.shape for a in P.transpose_in_list(sps)] == [(130, 24), (170, 24), (260, 24)]


def test_sample_train_data_global_rng_contract():
    """preprocess.py:207-238 draws from the global numpy RNG in a fixed order: shuffle A, shuffle B, then per pair the
    A crop start before the B crop start.  Replay that order by hand and compare."""
    P = _mod("preprocess")
    A = [np.tile(np.arange(n)[None, :] + 1000 * i, (24, 1)).astype(float) for i, n in enumerate([128, 200, 333, 150])]
    B = [np.tile(np.arange(n)[None, :] + 1000 * i, (24, 1)).astype(float) for i, n in enumerate([140, 128, 500])]
    np.random.seed(7)
    a, b = P.sample_train_data(A, B, 128)
    np.random.seed(7)
    ia = np.arange(4); ib = np.arange(3)
    np.random.shuffle(ia); np.random.shuffle(ib)
    for k in range(3):
        sa = np.random.randint(A[ia[k]].shape[1] - 128 + 1)
        sb = np.random.randint(B[ib[k]].shape[1] - 128 + 1)
        assert np.array_equal(a[k], A[ia[k]][:, sa:sa + 128]) and np.array_equal(b[k], B[ib[k]][:, sb:sb + 128])
    assert a.shape == (3, 24, 128) and b.shape == (3, 24, 128)


def test_world_wrappers_fail_loudly_without_pyworld():
    P = _mod("preprocess")
    try:
        import pyworld  # noqa: F401
        pytest.skip("pyworld is installed")
    except ImportError:
        pass
    with pytest.raises(ImportError, match="pyworld"):
        P.world_decompose(np.zeros(16000), 16000)


# ----------------------------------------------------------------------------------------------- convert.py host logic (CPU)
class _AffineModel:
    """stand-in for CycleGAN.test: y = 2x + 1, records the batches it was called with"""

    def __init__(self):
        self.calls = []

    def test(self, inputs, direction):
        assert inputs.ndim == 3 and inputs.shape[1] == 24 and inputs.shape[2] % 4 == 0
        self.calls.append((inputs.shape, direction))
        return (2.0 * inputs + 1.0).astype(np.float32)


def _stats(rs):
    return {"mean_A": rs.randn(24, 1), "std_A": rs.rand(24, 1) + 0.5, "mean_B": rs.randn(24, 1), "std_B": rs.rand(24, 1) + 0.5}


def test_convert_features_batches_by_length_and_denormalises():
    Cv = _mod("convert")
    rs = np.random.RandomState(2)
    st = _stats(rs)
    utts = [rs.randn(n, 24) for n in (128, 130, 131, 128, 259)]          # padded lengths 128, 132, 132, 128, 260
    m = _AffineModel()
    out = Cv.convert_features(m, utts, "A2B", st)
    assert sorted(c[0] for c in m.calls) == [(1, 24, 260), (2, 24, 128), (2, 24, 132)]
    assert m.calls[0][0] == (1, 24, 260)                                  # longest group first
    for u, o in zip(utts, out):
        T = u.shape[0]; Tp = -(-T // 4) * 4; left = (Tp - T) // 2
        assert o.shape == (T, 24) and o.flags["C_CONTIGUOUS"]             # cropped back: lines up with f0 / ap
        x = np.pad(u.T, ((0, 0), (left, Tp - T - left)), mode="edge")    # edge frames replicated, smaller half in front
        want = ((2.0 * ((x - st["mean_A"]) / st["std_A"]) + 1.0).astype(np.float32).astype(np.float64) * st["std_B"] + st["mean_B"]).T
        assert np.allclose(o, want[left:left + T], rtol=1e-6, atol=1e-6)
    # batches are bounded by a frame budget as well as by a count
    assert [(f, len(p)) for f, p in Cv.plan_groups([1400] * 5 + [400] * 7, max_group=4, frame_budget=2800)] == \
        [(1400, 2), (1400, 2), (1400, 1), (400, 4), (400, 3)]
    # the other direction swaps the statistics
    out2 = Cv.convert_features(_AffineModel(), utts[:1], "B2A", st)
    x = utts[0].T
    want = ((2.0 * ((x - st["mean_B"]) / st["std_B"]) + 1.0).astype(np.float32).astype(np.float64) * st["std_A"] + st["mean_A"]).T
    assert np.allclose(out2[0], want, rtol=1e-6, atol=1e-6)
    with pytest.raises(Exception, match="Conversion direction must be specified."):
        Cv.convert_features(m, utts, "A2A", st)


def test_validation_conversions_follow_the_reference_loop(tmp_path, capsys):
    """train.py:119-155: every 50th epoch, with the current weights, A -> B into output_dir/converted_A and B -> A into
    output_dir/converted_B; the whole-utterance forwards go through a forward-only model that receives a copy of the weights."""
    Tr, Cv = _mod("train"), _mod("convert")
    rs = np.random.RandomState(4)
    st = _stats(rs)
    lf = {"mean_A": 5.0, "std_A": 0.2, "mean_B": 4.6, "std_B": 0.3}
    dA, dB = tmp_path / "val_A", tmp_path / "val_B"
    dA.mkdir(); dB.mkdir()
    utt = {}
    for d, names in ((dA, ("a1", "a2")), (dB, ("b1",))):
        for n in names:
            T = int(rs.randint(130, 300))
            utt[n] = dict(f0=np.abs(rs.randn(T)) * 100 * (rs.rand(T) > 0.3), coded_sp=rs.randn(T, 24), ap=rs.rand(T, 5))
            np.savez(str(d / (n + ".npz")), **utt[n])
        (d / "notes.txt").write_text("ignored")

    class Trainer:
        def get_params(self):
            return {"w": np.array([3.0])}

    class Tester(_AffineModel):
        def __init__(self):
            super().__init__(); self.params = None

        def set_params(self, p):
            self.params = p

    tr, te = Trainer(), Tester()
    out = tmp_path / "validation_output"
    assert Tr.validation_conversions(tr, 7, str(dA), str(dB), str(out), st, lf, test_model=te) == []        # not a 50th epoch
    assert te.params is None and not out.exists()
    written = Tr.validation_conversions(tr, 100, str(dA), str(dB), str(out), st, lf, test_model=te)
    assert te.params == {"w": np.array([3.0])}                                      # current weights copied before converting
    assert sorted(os.path.relpath(w, str(out)) for w in written) == ["converted_A/a1.npz", "converted_A/a2.npz", "converted_B/b1.npz"]
    assert [c[1] for c in te.calls].count("A2B") >= 1 and [c[1] for c in te.calls][-1] == "B2A"
    txt = capsys.readouterr().out
    assert "Generating Validation Data B from A..." in txt and "Generating Validation Data A from B..." in txt      # train.py:121,139
    z = np.load(str(out / "converted_A" / "a1.npz"))
    want = Cv.convert_features(_AffineModel(), [utt["a1"]["coded_sp"]], "A2B", st)[0]
    assert np.allclose(z["coded_sp"], want) and z["coded_sp"].shape == utt["a1"]["coded_sp"].shape
    assert np.allclose(z["f0"], Cv.convert_f0(utt["a1"]["f0"], "A2B", lf)) and np.array_equal(z["ap"], utt["a1"]["ap"])
    zb = np.load(str(out / "converted_B" / "b1.npz"))
    assert np.allclose(zb["f0"], Cv.convert_f0(utt["b1"]["f0"], "B2A", lf))
    # one side only, and epoch 0 is a 50th epoch as in the reference (epoch % 50 == 0)
    only_a = Tr.validation_conversions(tr, 0, str(dA), None, str(tmp_path / "o2"), st, None, test_model=te)
    assert len(only_a) == 2 and not (tmp_path / "o2" / "converted_B").exists()
    assert np.array_equal(np.load(only_a[0])["f0"], utt["a1"]["f0"])                # no log-f0 statistics: f0 passes through


def test_convert_f0_direction():
    Cv = _mod("convert")
    st = {"mean_A": 5.0, "std_A": 0.2, "mean_B": 4.6, "std_B": 0.3}
    f0 = np.array([0.0, 150.0, 200.0])
    ab = Cv.convert_f0(f0, "A2B", st); ba = Cv.convert_f0(ab, "B2A", st)
    assert ab[0] == 0 and np.allclose(ba[1:], f0[1:])


# ----------------------------------------------------------------------------------------------- end to end on the engine
@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["bf16x3", "f16f8"])
def test_conversion_driver_on_feature_files(tmp_path, precision):
    """conversion() on .npz feature files == per-utterance model.test + the reference's (de)normalisation, and batching
    utterances of equal padded length does not change any utterance's result."""
    import cgvc
    Cv = _mod("convert")
    rs = np.random.RandomState(3)
    mdir, ddir, odir = tmp_path / "model", tmp_path / "feat", tmp_path / "out"
    os.makedirs(mdir); os.makedirs(ddir)
    m = cgvc.CycleGAN(num_features=24, mode="test", max_batch=2, max_frames=160, seed=11, precision=precision)
    m.save(str(mdir), "x.ckpt")
    st = _stats(rs)
    np.savez(str(mdir / "mcep_normalization.npz"), **st)
    np.savez(str(mdir / "logf0s_normalization.npz"), mean_A=5.0, std_A=0.2, mean_B=4.6, std_B=0.3)
    lens = {"u0.npz": 128, "u1.npz": 141, "u2.npz": 144, "u3.npz": 126}
    feats = {}
    for name, T in lens.items():
        f0 = np.where(rs.rand(T) < 0.3, 0.0, np.exp(rs.randn(T) * 0.2 + 5.0))
        sp = rs.randn(T, 24) * 2 + 1
        feats[name] = (f0, sp)
        np.savez(str(ddir / name), f0=f0, coded_sp=sp, ap=rs.rand(T, 513))
    written = Cv.conversion(str(mdir), "x.ckpt", str(ddir), "A2B", str(odir), precision=precision)
    assert sorted(os.path.basename(w) for w in written) == sorted(lens)
    P = _mod("preprocess")
    for name, (f0, sp) in feats.items():
        z = np.load(str(odir / name))
        T = sp.shape[0]; Tp = -(-T // 4) * 4; left = (Tp - T) // 2
        x = (np.pad(sp.T, ((0, 0), (left, Tp - T - left)), mode="edge") - st["mean_A"]) / st["std_A"]
        y = m.test(np.array([x]), "A2B")[0]
        want = (y.astype(np.float64) * st["std_B"] + st["mean_B"]).T[left:left + T]
        err = np.linalg.norm(z["coded_sp"] - want) / np.linalg.norm(want)
        assert err < 1e-5, (name, err)                                     # same kernels; only the batch composition differs
        assert np.allclose(z["f0"], P.pitch_conversion(f0, 5.0, 0.2, 4.6, 0.3))
        assert z["ap"].shape == (lens[name], 513) and z["coded_sp"].shape == (lens[name], 24) and z["f0"].shape == (lens[name],)


@pytest.mark.gpu
def test_train_then_convert_end_to_end(tmp_path):
    """The whole caller chain of SURVEY.md 8f on feature files: `cgvc.train` (loop, schedule, per-epoch checkpoint, MCEP and log-f0
    normalisation side files, train.py:47-57,78-118) followed by `cgvc.convert.conversion` on the model directory it wrote."""
    T = _mod("train"); Cv = _mod("convert")
    rs = np.random.RandomState(4)
    dirs = {}
    for spk, base in (("SF1", 5.3), ("TM1", 4.7)):
        d = tmp_path / "feat" / spk; os.makedirs(d); dirs[spk] = str(d)
        for u in range(3):
            n = int(rs.randint(130, 200))
            f0 = np.where(rs.rand(n) < 0.25, 0.0, np.exp(rs.randn(n) * 0.15 + base))
            np.savez(str(d / ("u%d.npz" % u)), f0=f0, coded_sp=np.cumsum(rs.randn(n, 24), axis=0) * 0.1 + rs.randn(1, 24), ap=rs.rand(n, 4))
    mdir = str(tmp_path / "model")
    model, g, d_loss = T.train(dirs["SF1"], dirs["TM1"], mdir, "sf1_tm1.ckpt", 0, num_epochs=1, mini_batch_size=1, log_every=1)
    assert model.train_step == 3 and np.isfinite(g) and np.isfinite(d_loss)
    z = np.load(os.path.join(mdir, "logf0s_normalization.npz"))
    assert set(z.files) == {"mean_A", "std_A", "mean_B", "std_B"} and abs(float(z["mean_A"]) - 5.3) < 0.1 and abs(float(z["mean_B"]) - 4.7) < 0.1
    del model
    out = Cv.conversion(mdir, "sf1_tm1.ckpt", dirs["SF1"], "A2B", str(tmp_path / "converted"))
    assert len(out) == 3
    for path in out:
        r = np.load(path)
        src = np.load(os.path.join(dirs["SF1"], os.path.basename(path)))
        n = src["coded_sp"].shape[0]
        assert r["coded_sp"].shape == (n, 24) and r["f0"].shape == (n,) and np.isfinite(r["coded_sp"]).all()
        v = src["f0"] > 0
        assert (r["f0"][~v] == 0).all() and np.isfinite(r["f0"]).all()
        # converted pitch is centred on speaker B's log-f0 statistics
        assert abs(np.log(r["f0"][v]).mean() - float(z["mean_B"])) < 0.15
