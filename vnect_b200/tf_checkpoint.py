"""Reader (and writer) for TensorFlow-1.x checkpoints, without TensorFlow.

The reference restores ``./models/tf_model/vnect_tf`` (src/estimator.py:55-60), a TF "tensor bundle" that
``init_weights.py:22-25`` saved from the pickled Caffe weights:

    vnect_tf.index                   an SSTable (LevelDB table format): key -> serialized proto
    vnect_tf.data-00000-of-00001     the raw tensor bytes
    checkpoint                       text: model_checkpoint_path: "vnect_tf"  (what tf.train.latest_checkpoint reads)

Index format (public TensorFlow sources, tensorflow/core/util/tensor_bundle + core/lib/io/table): a 48-byte footer
(metaindex handle, index handle, magic 0xdb4775248b80fb57); blocks of prefix-compressed entries
(varint shared, varint non_shared, varint value_len, key suffix, value) followed by a restart array, each block
trailed by 1 byte of compression type (0 = none, the bundle writer's setting) and a masked CRC32C.  Key "" maps to a
BundleHeaderProto {num_shards=1, endianness=2, version=3}; every other key is a variable name mapping to a
BundleEntryProto {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6 (masked, fixed32)}.

PARITY UNPINNED against real TensorFlow output: TensorFlow is not installable in this environment and the reference
ships no checkpoint (the trained weights are a separate download), so this module is pinned only against itself
(write -> read round trip of all 109 variables, tests/test_weight_formats.py) and the format description above.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
DT_FLOAT = 1
_MASK_DELTA = 0xa282ead8


# ------------------------------------------------------------------------------------------------ CRC32C (Castagnoli)
def _make_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_TABLE = _make_table()


def crc32c(data):
    """CRC32C of a bytes-like object.  Uses the native helper of libvnect_b200.so when the library is built (58 MB of
    weights take milliseconds instead of a minute), the table-driven Python loop otherwise."""
    try:
        from . import _capi
        import ctypes as C
        lib = _capi.load_library()
        buf = bytes(data)
        return int(lib.vnect_crc32c(C.cast(C.c_char_p(buf), C.c_void_p), len(buf)))
    except Exception:
        c = 0xFFFFFFFF
        for b in bytes(data):
            c = _TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
        return c ^ 0xFFFFFFFF


def mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ varints / protobuf
def _varint(buf, pos):
    shift, out = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """Minimal protobuf walk: {field: [values]} with varints as int, length-delimited as bytes, fixed32/64 as int."""
    pos, out = 0, {}
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.setdefault(field, []).append(v)
    return out


def _field(field, wt, payload):
    return _put_varint((field << 3) | wt) + payload


def _shape_proto(shape):
    return b"".join(_field(2, 2, _put_varint(len(d)) + d) for d in (_field(1, 0, _put_varint(int(s))) for s in shape))


def _parse_shape(buf):
    dims = []
    for d in _parse_proto(buf).get(2, []):
        dims.append(_parse_proto(d).get(1, [0])[0])
    return tuple(int(x) for x in dims)


# ------------------------------------------------------------------------------------------------ SSTable
def _read_block(data, offset, size, verify=True):
    body = data[offset:offset + size]
    ctype = data[offset + size]
    if ctype != 0:
        raise ValueError("compressed index block (type %d): only uncompressed tensor-bundle indexes are supported" % ctype)
    if verify:
        stored = struct.unpack_from("<I", data, offset + size + 1)[0]
        if mask_crc(crc32c(data[offset:offset + size + 1])) != stored:
            raise ValueError("index block checksum mismatch")
    n_restarts = struct.unpack_from("<I", body, len(body) - 4)[0]
    end = len(body) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _varint(body, pos)
        non_shared, pos = _varint(body, pos)
        vlen, pos = _varint(body, pos)
        key = key[:shared] + bytes(body[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(body[pos:pos + vlen])))
        pos += vlen
    return out


def _handle(buf, pos=0):
    off, pos = _varint(buf, pos)
    size, pos = _varint(buf, pos)
    return off, size, pos


def read_index(path, verify=True):
    """-> (header dict, {name: entry dict}) of a .index file."""
    data = open(path, "rb").read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != TABLE_MAGIC:
        raise ValueError("%s is not a TensorFlow checkpoint index (bad table magic)" % path)
    footer = data[-48:]
    _, _, p = _handle(footer)
    ioff, isize, _ = _handle(footer, p)
    entries, header = {}, None
    for _, handle in _read_block(data, ioff, isize, verify):
        boff, bsize, _ = _handle(handle)
        for key, val in _read_block(data, boff, bsize, verify):
            msg = _parse_proto(val)
            if key == b"":
                header = dict(num_shards=msg.get(1, [1])[0], endianness=msg.get(2, [0])[0])
            else:
                entries[key.decode()] = dict(dtype=msg.get(1, [0])[0], shape=_parse_shape(msg.get(2, [b""])[0]),
                                             shard=msg.get(3, [0])[0], offset=msg.get(4, [0])[0],
                                             size=msg.get(5, [0])[0], crc=msg.get(6, [None])[0])
    if header is None:
        raise ValueError("checkpoint index has no bundle header")
    if header["endianness"] != 0:
        raise ValueError("big-endian checkpoints are not supported")
    return header, entries


def read_checkpoint(prefix, verify=True):
    """{variable name: float32 ndarray} of the checkpoint ``prefix`` (e.g. './models/tf_model/vnect_tf').  Entries
    that are not float32 tensors (the saver's bookkeeping) are skipped.  ``verify`` checks every block and tensor CRC."""
    header, entries = read_index(prefix + ".index", verify)
    shards = {}
    out = {}
    for name, e in entries.items():
        if e["dtype"] != DT_FLOAT:
            continue
        if e["shard"] not in shards:
            shards[e["shard"]] = open("%s.data-%05d-of-%05d" % (prefix, e["shard"], header["num_shards"]), "rb").read()
        raw = shards[e["shard"]][e["offset"]:e["offset"] + e["size"]]
        n = int(np.prod(e["shape"])) if e["shape"] else 1
        if len(raw) != 4 * n:
            raise ValueError("variable %s: %d bytes for shape %s" % (name, len(raw), e["shape"]))
        if verify and e["crc"] is not None and mask_crc(crc32c(raw)) != e["crc"]:
            raise ValueError("variable %s: tensor checksum mismatch" % name)
        out[name] = np.frombuffer(raw, dtype="<f4").reshape(e["shape"]).copy()
    return out


def latest_checkpoint(directory):
    """tf.train.latest_checkpoint: the prefix named by ``directory/checkpoint`` (None when absent)."""
    state = os.path.join(directory, "checkpoint")
    if not os.path.isfile(state):
        return None
    for line in open(state):
        if line.startswith("model_checkpoint_path:"):
            name = line.split(":", 1)[1].strip().strip('"')
            return name if os.path.isabs(name) else os.path.join(directory, name)
    return None


# ------------------------------------------------------------------------------------------------ writer
def _build_block(items, restart_interval=16):
    body, restarts, prev = bytearray(), [], b""
    for i, (key, val) in enumerate(items):
        if i % restart_interval == 0:
            restarts.append(len(body))
            shared = 0
        else:
            shared = 0
            while shared < min(len(prev), len(key)) and prev[shared] == key[shared]:
                shared += 1
        body += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(val)) + key[shared:] + val
        prev = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        body += struct.pack("<I", r)
    body += struct.pack("<I", len(restarts))
    return bytes(body)


def write_checkpoint(prefix, variables, block_entries=32):
    """Write {name: ndarray} as a single-shard float32 tensor bundle + the ``checkpoint`` state file."""
    names = sorted(variables)
    data = bytearray()
    items = [(b"", _field(1, 0, _put_varint(1)) + _field(3, 2, (lambda v: _put_varint(len(v)) + v)(_field(1, 0, _put_varint(1)))))]
    for name in names:
        arr = np.asarray(variables[name], dtype="<f4")  # (ascontiguousarray would turn a scalar into shape (1,))
        raw = arr.tobytes()
        entry = _field(1, 0, _put_varint(DT_FLOAT))
        shp = _shape_proto(arr.shape)
        entry += _field(2, 2, _put_varint(len(shp)) + shp)
        if len(data):
            entry += _field(4, 0, _put_varint(len(data)))
        entry += _field(5, 0, _put_varint(len(raw)))
        entry += _field(6, 5, struct.pack("<I", mask_crc(crc32c(raw))))
        items.append((name.encode(), entry))
        data += raw
    out = bytearray()
    index_items = []

    def emit(block):
        off = len(out)
        out.extend(block)
        out.append(0)
        out.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))
    for i in range(0, len(items), block_entries):
        chunk = items[i:i + block_entries]
        handle = emit(_build_block(chunk))
        index_items.append((chunk[-1][0] + b"\x00" if i + block_entries < len(items) else chunk[-1][0], handle))
    meta = emit(_build_block([]))
    idx = emit(_build_block(index_items, restart_interval=1))
    footer = meta + idx
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
    with open(os.path.join(os.path.dirname(prefix) or ".", "checkpoint"), "w") as f:
        base = os.path.basename(prefix)
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))
