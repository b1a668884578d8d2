"""Reader for TensorFlow V2 checkpoint bundles (`<prefix>.index` + `<prefix>.data-00000-of-00001`), in pure Python / numpy.

Why: the reference saves its models with `tf.train.Saver` (`model.py:21,140-150`) and its README publishes a trained SF1-TM1 checkpoint;
`CycleGAN.load()` accepts such a bundle so that `convert.py` can run on it without TensorFlow.  The variable names in
the bundle are the reference's graph names, which are also this engine's parameter-table names (`generator_A2B/h1_conv/kernel`, ...,
`.../kernel/Adam`, `.../kernel/Adam_1`, `beta1_power`, `beta2_power`).

Formats restated here (TensorFlow sources are NOT under /root/reference; no TensorFlow-written file was available offline, so this reader is
verified only against bundles written by the minimal writer below, which follows the same format description -- see tests/test_tf_checkpoint.py):

  * `.index` is a LevelDB-style sorted string table (tensorflow/core/lib/io/table*, same layout as leveldb `table_format.md`):
    blocks of prefix-compressed entries `varint32 shared | varint32 unshared | varint32 value_len | key suffix | value`, a restart array
    (`fixed32 * n`, `fixed32 n`), a 5-byte block trailer (`compression type`, masked crc32c), an index block mapping last-key -> BlockHandle
    (`varint64 offset | varint64 size`), and a 48-byte footer (metaindex handle, index handle, padding, magic 0xdb4775248b80fb57).
    TensorFlow writes the bundle index uncompressed; a snappy block raises.
  * the entry with the empty key holds `BundleHeaderProto` (num_shards = 1, endianness = 2, version = 3); every other key is a tensor name
    whose value is a `BundleEntryProto`: dtype = 1, shape = 2 (TensorShapeProto: repeated dim = 2 {size = 1}), shard_id = 3, offset = 4,
    size = 5, crc32c = 6 (fixed32), slices = 7 (partitioned variables: not supported here).
  * the data shard holds the raw little-endian tensor bytes at [offset, offset + size).
"""
from __future__ import annotations

import os
import struct
from collections import OrderedDict

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_, 4: np.uint8, 6: np.int8, 5: np.int16}
DTYPE_CODES = {np.dtype(v): k for k, v in DTYPES.items()}


# ------------------------------------------------------------------------------------------------ varints / protobuf wire format
def _varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]; pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _enc_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F; v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b); return bytes(out)


def _pb_fields(buf):
    """iterate (field number, wire type, value) over a serialized protobuf message; value = int or bytes"""
    pos, n = 0, len(buf)
    while pos < n:
        tag, pos = _varint(buf, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]; pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos); v = bytes(buf[pos:pos + ln]); pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]; pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield f, wt, v


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_shape(buf):
    dims = []
    for f, wt, v in _pb_fields(buf):
        if f == 2 and wt == 2:                                  # TensorShapeProto.Dim
            size = 0
            for f2, wt2, v2 in _pb_fields(v):
                if f2 == 1 and wt2 == 0:
                    size = _signed64(v2)
            dims.append(size)
        elif f == 3 and wt == 0 and v:
            raise ValueError("tensor of unknown rank in checkpoint")
    return tuple(dims)


def _parse_entry(buf):
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
    for f, wt, v in _pb_fields(buf):
        if f == 1: e["dtype"] = v
        elif f == 2: e["shape"] = _parse_shape(v)
        elif f == 3: e["shard_id"] = v
        elif f == 4: e["offset"] = _signed64(v)
        elif f == 5: e["size"] = _signed64(v)
        elif f == 6: e["crc32c"] = v
        elif f == 7: e["sliced"] = True
    return e


# ------------------------------------------------------------------------------------------------ sorted string table
def _block_entries(block):
    """(key, value) pairs of one table block (contents without the 5-byte trailer)"""
    if len(block) < 4:
        raise ValueError("table block too short")
    nrestarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * nrestarts
    if limit < 0:
        raise ValueError("corrupt restart array")
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _varint(block, pos)
        unshared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        if shared > len(key):
            raise ValueError("corrupt prefix compression")
        key = key[:shared] + bytes(block[pos:pos + unshared]); pos += unshared
        yield key, bytes(block[pos:pos + vlen]); pos += vlen


def _read_block(data, offset, size):
    if offset + size + 5 > len(data):
        raise ValueError("block handle outside the file")
    ctype = data[offset + size]
    if ctype != 0:
        raise ValueError("compressed table block (type %d): only uncompressed bundle indexes are supported" % ctype)
    return data[offset:offset + size]


def read_table(path):
    """All (key, value) pairs of a LevelDB-format table file, in key order."""
    data = open(path, "rb").read()
    if len(data) < 48:
        raise ValueError("%s: too short for a table footer" % path)
    footer = data[-48:]
    lo, hi = struct.unpack_from("<II", footer, 40)
    if (hi << 32 | lo) != TABLE_MAGIC:
        raise ValueError("%s: bad table magic (not a TensorFlow checkpoint index?)" % path)
    pos = 0
    _, pos = _varint(footer, pos); _, pos = _varint(footer, pos)            # metaindex handle (unused)
    ioff, pos = _varint(footer, pos); isize, pos = _varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isize)):
        boff, p = _varint(handle, 0); bsize, p = _varint(handle, p)
        out.extend(_block_entries(_read_block(data, boff, bsize)))
    return out


# ------------------------------------------------------------------------------------------------ bundle
def is_bundle(prefix):
    return os.path.exists(prefix + ".index")


def read_checkpoint(prefix, names=None):
    """name -> numpy array for every (or the requested) tensor of the V2 bundle `<prefix>.index` / `.data-*`."""
    entries = read_table(prefix + ".index")
    num_shards = 1
    tensors = OrderedDict()
    metas = []
    for key, value in entries:
        if key == b"":
            for f, wt, v in _pb_fields(value):
                if f == 1 and wt == 0:
                    num_shards = v
                elif f == 2 and wt == 0 and v == 1:
                    raise ValueError("big-endian checkpoint")
            continue
        name = key.decode("utf-8")
        if names is not None and name not in names:
            continue
        metas.append((name, _parse_entry(value)))
    shards = {}
    for name, e in metas:
        if e["sliced"]:
            raise ValueError("%s is a partitioned variable (tensor slices are not supported)" % name)
        if e["dtype"] not in DTYPES:
            raise ValueError("%s: unsupported dtype code %d" % (name, e["dtype"]))
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap("%s.data-%05d-of-%05d" % (prefix, sid, num_shards), dtype=np.uint8, mode="r")
        dt = np.dtype(DTYPES[e["dtype"]])
        count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if count * dt.itemsize != e["size"]:
            raise ValueError("%s: size %d does not match shape %r" % (name, e["size"], e["shape"]))
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        tensors[name] = np.frombuffer(raw.tobytes(), dtype=dt.newbyteorder("<")).reshape(e["shape"]).astype(dt)
    return tensors


# ------------------------------------------------------------------------------------------------ minimal writer (tests; small tensors)
_CRC_TABLE = None


def crc32c(data, crc=0):
    """Castagnoli CRC (pure Python, byte at a time: for index blocks and small test tensors only)."""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _CRC_TABLE = t
    c = crc ^ 0xFFFFFFFF
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _pb_varint_field(f, v):
    return _enc_varint(f << 3) + _enc_varint(v)


def _pb_bytes_field(f, b):
    return _enc_varint(f << 3 | 2) + _enc_varint(len(b)) + b


def _build_block(pairs, restart_interval=16):
    out = bytearray(); restarts = []; prev = b""
    for i, (k, v) in enumerate(pairs):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _enc_varint(shared) + _enc_varint(len(k) - shared) + _enc_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_checkpoint(prefix, tensors, block_entries=64):
    """Write a single-shard V2 bundle (uncompressed index).  Meant for tests and small exports: the per-tensor crc32c is computed in
    pure Python."""
    names = sorted(tensors.keys(), key=lambda s: s.encode("utf-8"))
    data = bytearray(); pairs = []
    header = _pb_varint_field(1, 1) + _pb_varint_field(2, 0) + _pb_bytes_field(3, _pb_varint_field(1, 1))
    pairs.append((b"", header))
    for n in names:
        a = np.asarray(tensors[n])
        code = DTYPE_CODES[a.dtype]
        raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
        shape = b"".join(_pb_bytes_field(2, _pb_varint_field(1, int(d))) for d in a.shape)
        entry = _pb_varint_field(1, code) + _pb_bytes_field(2, shape) + _pb_varint_field(4, len(data)) + _pb_varint_field(5, len(raw)) \
            + _enc_varint(6 << 3 | 5) + struct.pack("<I", masked_crc(raw))
        pairs.append((n.encode("utf-8"), entry))
        data += raw
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
    out = bytearray(); index_pairs = []
    for s in range(0, len(pairs), block_entries):
        blk = _build_block(pairs[s:s + block_entries])
        handle = _enc_varint(len(out)) + _enc_varint(len(blk))
        out += blk + b"\x00" + struct.pack("<I", masked_crc(blk + b"\x00"))
        index_pairs.append((pairs[min(s + block_entries, len(pairs)) - 1][0], handle))
    meta = _build_block([])
    meta_handle = _enc_varint(len(out)) + _enc_varint(len(meta))
    out += meta + b"\x00" + struct.pack("<I", masked_crc(meta + b"\x00"))
    idx = _build_block(index_pairs, restart_interval=1)
    idx_handle = _enc_varint(len(out)) + _enc_varint(len(idx))
    out += idx + b"\x00" + struct.pack("<I", masked_crc(idx + b"\x00"))
    footer = meta_handle + idx_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<II", TABLE_MAGIC & 0xFFFFFFFF, TABLE_MAGIC >> 32)
    out += footer
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
    return prefix
