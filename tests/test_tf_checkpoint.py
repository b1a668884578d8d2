"""TensorFlow V2 checkpoint bundles (model.py:140-150 saves with tf.train.Saver): the pure-Python reader CycleGAN.load() uses.

No TensorFlow-written file is available offline, so the reader is exercised against bundles produced by the minimal writer in the
same module (which follows the published table / BundleEntryProto layout); the test pins the byte-level pieces both rely on with
known answers: the crc32c check value, its TensorFlow masking, varint and footer encodings."""
import importlib
import os
import struct

import numpy as np
import pytest


def _mod():
    import cgvc  # noqa: F401
    return importlib.import_module("cgvc.tf_checkpoint")


def test_crc32c_known_answers():
    T = _mod()
    assert T.crc32c(b"123456789") == 0xE3069283                      # the standard CRC-32C check value
    assert T.crc32c(b"") == 0
    assert T.crc32c(bytes(32)) == 0x8A9136AA                         # RFC 3720 B.4: 32 bytes of zeros
    assert T.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43                # RFC 3720 B.4: 32 bytes of ones
    c = T.crc32c(b"foo")
    assert T.masked_crc(b"foo") == ((((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF)


def test_varint_and_footer_layout(tmp_path):
    T = _mod()
    for v in (0, 1, 127, 128, 300, 2 ** 31, 2 ** 40 + 5):
        enc = T._enc_varint(v)
        assert T._varint(enc, 0) == (v, len(enc)) and all(b & 0x80 for b in enc[:-1]) and not enc[-1] & 0x80
    prefix = str(tmp_path / "m.ckpt")
    T.write_checkpoint(prefix, {"a": np.arange(6, dtype=np.float32).reshape(2, 3)})
    raw = open(prefix + ".index", "rb").read()
    assert raw[-8:] == struct.pack("<Q", 0xDB4775248B80FB57)          # leveldb table magic, little-endian
    assert len(raw) >= 48
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    assert data == np.arange(6, dtype="<f4").tobytes()


def test_bundle_roundtrip_many_tensors(tmp_path):
    """multi-block index (prefix-compressed keys with restarts), several dtypes, scalars and empty shapes"""
    T = _mod()
    rs = np.random.RandomState(0)
    tensors = {}
    for net in ("generator_A2B", "generator_B2A", "discriminator_A"):
        for i in range(40):
            tensors["%s/residual1d_block%d_h1_conv/kernel" % (net, i)] = rs.randn(3, 4, 5).astype(np.float32)
            tensors["%s/residual1d_block%d_h1_conv/kernel/Adam" % (net, i)] = rs.randn(3, 4, 5).astype(np.float32)
            tensors["%s/InstanceNorm_%d/gamma" % (net, i)] = rs.randn(7).astype(np.float32)
    tensors["beta1_power"] = np.float32(0.5 ** 12)
    tensors["beta2_power"] = np.float32(0.999 ** 12)
    tensors["global_step"] = np.int64(12)
    tensors["some/double"] = rs.randn(2, 2)
    tensors["some/int32"] = np.arange(5, dtype=np.int32)
    prefix = str(tmp_path / "model.ckpt")
    T.write_checkpoint(prefix, tensors, block_entries=16)
    assert T.is_bundle(prefix) and not T.is_bundle(prefix + "x")
    back = T.read_checkpoint(prefix)
    assert list(back) == sorted(tensors, key=lambda s: s.encode())
    for k, v in tensors.items():
        assert back[k].dtype == np.asarray(v).dtype and back[k].shape == np.asarray(v).shape and np.array_equal(back[k], v), k
    some = T.read_checkpoint(prefix, names={"beta2_power", "generator_B2A/InstanceNorm_3/gamma"})
    assert set(some) == {"beta2_power", "generator_B2A/InstanceNorm_3/gamma"}


def test_reader_rejects_garbage(tmp_path):
    T = _mod()
    p = str(tmp_path / "bad")
    open(p + ".index", "wb").write(b"\x00" * 100)
    with pytest.raises(ValueError, match="magic"):
        T.read_checkpoint(p)
    open(p + ".index", "wb").write(b"\x00" * 10)
    with pytest.raises(ValueError, match="too short"):
        T.read_checkpoint(p)
    # a snappy-compressed block is refused, not misread
    prefix = str(tmp_path / "c.ckpt")
    T.write_checkpoint(prefix, {"a": np.zeros(3, np.float32)})
    raw = bytearray(open(prefix + ".index", "rb").read())
    entries_end = raw.index(b"\x00\x00\x00\x00\x01\x00\x00\x00")      # first block's restart array [0], count 1
    raw[entries_end + 8] = 1                                          # its trailer's compression-type byte
    open(prefix + ".index", "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="compressed"):
        T.read_checkpoint(prefix)


@pytest.mark.gpu
def test_model_loads_tf_bundle(tmp_path):
    """CycleGAN.load() on a TensorFlow bundle: weights, Adam slots and the shared Adam step land where the .npz path puts them."""
    import torch
    import cgvc
    from oracle import cyclegan_oracle as O
    T = _mod()
    P = O.init_params(seed=21, dtype=torch.float32, perturb_affine=True)
    tensors = {k: v.numpy() for k, v in P.items()}
    k0 = "generator_A2B/h1_conv/kernel"
    tensors[k0 + "/Adam"] = np.full(tensors[k0].shape, 0.25, np.float32)
    tensors[k0 + "/Adam_1"] = np.full(tensors[k0].shape, 0.5, np.float32)
    tensors["beta1_power"] = np.float32(0.5 ** 7); tensors["beta2_power"] = np.float32(0.999 ** 7)
    prefix = str(tmp_path / "sf1_tm1.ckpt")
    # 120 M weights through the pure-Python crc would take minutes: write the data shard directly, the index with dummy crcs
    real_masked = T.masked_crc
    T.masked_crc = lambda b: 0 if len(b) > 1 << 16 else real_masked(b)
    try:
        T.write_checkpoint(prefix, tensors)
    finally:
        T.masked_crc = real_masked
    m = cgvc.CycleGAN(num_features=24, mode='train', max_batch=1, max_frames=128, log_dir=str(tmp_path / "log"))
    m.load(prefix)
    got = m.get_params()
    for k in ("generator_A2B/h1_conv/kernel", "generator_B2A/residual1d_block3_h2_conv/kernel", "discriminator_B/dense/kernel", "discriminator_A/InstanceNorm_4/gamma"):
        assert np.array_equal(got[k], tensors[k]), k
    import ctypes as C
    step = C.c_longlong(0); m._lib.cgvc_get_adam_step(m._handle, C.byref(step))
    assert step.value == 7
    assert float(m._view(cgvc.native.ARENA_ADAM_M, k0).flatten()[0]) == 0.25 and float(m._view(cgvc.native.ARENA_ADAM_V, k0).flatten()[0]) == 0.5
    x = O.synthetic_batch(seed=3, batch=1, frames=128)[0]
    with torch.no_grad():
        ref = O.generator_forward(x, P, "generator_A2B").numpy()
    y = m.test(x.numpy(), 'A2B')
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 1e-3
