"""TensorFlow tensor-bundle ("checkpoint V2") reader / writer without TensorFlow.

The reference keeps its weights in TensorFlow files only: the pretrained VGG-16 and every saved FCN-8s are SavedModel
directories whose variables live in `variables/variables.index` + `variables/variables.data-00000-of-00001`
(`fcn8s_tensorflow.py:74,134` load them, `:922-925` writes them), and `saver='train_saver'` / `load_variables()` use the
same two files under the prefix `<dir>/variables` (`:926-934, 938-944`).  This module reads and writes that pair of files
so that weights move between the reference and this engine by TF variable name, TF HWIO layout
untouched.  `saved_model.pb` (the TF1 graph) is neither read nor written: the graph is this engine.

Format (restated from TensorFlow's public sources, tensorflow/core/util/tensor_bundle/tensor_bundle.{h,cc},
tensorflow/core/lib/io/{table_builder,format,block}.cc and tensorflow/core/protobuf/tensor_bundle.proto; the table is
LevelDB's):
  * `<prefix>.data-00000-of-00001`: the raw little-endian bytes of every tensor, back to back, in key order.
  * `<prefix>.index`: an immutable sorted string table.  Key "" -> BundleHeaderProto {num_shards=1, endianness=LITTLE,
    version{producer=1}}; key <tensor name> -> BundleEntryProto {dtype, shape, shard_id=0, offset, size,
    crc32c = masked CRC-32C of the tensor's bytes}.
  * table = data blocks + (empty) metaindex block + index block + 48-byte footer.  A block is a run of entries
    `varint shared | varint non_shared | varint value_len | key suffix | value`, then the uint32 restart offsets and
    their count; every block is followed by a 1-byte compression type (0 = none, the only one written and read here)
    and the masked CRC-32C of block + type.  Index-block values are block handles (varint offset, varint size); the
    footer holds the metaindex and index handles padded to 40 bytes and the magic 0xdb4775248b80fb57.
  * masked crc = rotr(crc, 15) + 0xa282ead8 (mod 2^32).

Parity note: TensorFlow cannot be installed here, so there is no TF-written golden file.  What
pins the format instead (tests/test_tf_bundle.py): the RFC 3720 CRC-32C vectors; TensorFlow's own CRC-32C / mask
code and the protobuf classes generated from TensorFlow's .proto files as shipped in the `tensorboard` wheel (checksum
and mask equal on random data, DataType enum values, TensorShapeProto / VersionDef sub-messages byte-identical); a
hand-assembled known-answer table for the LevelDB container; and round trips.
"""
import os
import struct
from collections import OrderedDict

import numpy as np

_MAGIC = 0xdb4775248b80fb57
_MASK_DELTA = 0xa282ead8
_BLOCK_BYTES = 64 * 1024        # flush a data block of the index table at this size
_RESTART_INTERVAL = 16

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


# --------------------------------------------------------------------------------------------------------- crc32c
_PY_TABLE = None


def _crc32c_py(data, crc=0):
    """Bytewise table CRC-32C: the fallback when the native library is not built (slow: small inputs only)."""
    global _PY_TABLE
    if _PY_TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _PY_TABLE = t
    c = crc ^ 0xFFFFFFFF
    for b in bytes(data):
        c = (c >> 8) ^ _PY_TABLE[(c ^ b) & 0xFF]
    return c ^ 0xFFFFFFFF


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli) of a bytes-like object or a C-contiguous ndarray; native (`fcn8_crc32c`) when the library is
    built, pure Python otherwise."""
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else \
        np.ascontiguousarray(data).reshape(-1).view(np.uint8)
    if a.size == 0:
        return crc
    try:
        from . import _capi
        lib = _capi.load()
    except Exception:                                   # library not built: tiny files still work
        return _crc32c_py(a.tobytes(), crc)
    return int(lib.fcn8_crc32c(a.ctypes.data, a.size, crc))


def mask_crc(crc):
    return (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(masked):
    rot = (masked - _MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# --------------------------------------------------------------------------------------------------------- varints / protobuf
def _varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _read_varint(buf, pos):
    v = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if b < 0x80:
            return v, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _pb_fields(buf):
    """Yield (field_number, wire_type, value) of one protobuf message; value is an int (wire types 0, 1, 5) or bytes."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield field, wt, v


def _pb_varint_field(field, v):
    return _varint(field << 3) + _varint(v)


def _pb_bytes_field(field, payload):
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _encode_header():
    # BundleHeaderProto: num_shards = 1 (field 1), endianness LITTLE = 0 (default, omitted), version (field 3) =
    # VersionDef{producer = 1}
    return _pb_varint_field(1, 1) + _pb_bytes_field(3, _pb_varint_field(1, 1))


def _encode_entry(dtype_id, shape, offset, size, masked_crc, shard_id=0):
    # proto3 serialisation: fields holding their default value (0) are omitted -- byte-identical to what TensorFlow's
    # generated classes write (checked against tensorboard's copies of them in tests/test_tf_bundle.py)
    dims = b"".join(_pb_bytes_field(2, _pb_varint_field(1, int(d)) if int(d) else b"") for d in shape)  # dim[].size
    out = _pb_varint_field(1, dtype_id) + _pb_bytes_field(2, dims)
    if shard_id:
        out += _pb_varint_field(3, shard_id)
    if offset:
        out += _pb_varint_field(4, offset)
    if size:
        out += _pb_varint_field(5, size)
    if masked_crc:
        out += _varint((6 << 3) | 5) + struct.pack("<I", masked_crc)                  # fixed32 crc32c
    return out


def _decode_entry(buf):
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for field, _, v in _pb_fields(buf):
        if field == 1:
            e["dtype"] = v
        elif field == 2:
            for f2, _, v2 in _pb_fields(v):
                if f2 == 2:   # dim
                    size = 0
                    for f3, _, v3 in _pb_fields(v2):
                        if f3 == 1:
                            size = v3 if v3 < (1 << 63) else v3 - (1 << 64)
                    e["shape"].append(size)
        elif field == 3:
            e["shard_id"] = v
        elif field == 4:
            e["offset"] = v
        elif field == 5:
            e["size"] = v
        elif field == 6:
            e["crc32c"] = v
        elif field == 7:
            e["slices"] += 1
    return e


# --------------------------------------------------------------------------------------------------------- table
class _BlockBuilder:
    def __init__(self, restart_interval=_RESTART_INTERVAL):
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last_key = b""
        self.interval = restart_interval

    def add(self, key, value):
        shared = 0
        if self.count and self.count % self.interval == 0:
            self.restarts.append(len(self.buf))
        elif self.count:
            m = min(len(key), len(self.last_key))
            while shared < m and key[shared] == self.last_key[shared]:
                shared += 1
        self.buf += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
        self.last_key = key
        self.count += 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + \
            struct.pack("<I", len(self.restarts))

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4


def _write_block(f, contents):
    """Append block + trailer (type 0 = uncompressed, masked crc32c of contents + type); return its handle."""
    off = f.tell()
    trailer_type = b"\x00"
    f.write(contents)
    f.write(trailer_type)
    f.write(struct.pack("<I", mask_crc(crc32c(contents + trailer_type))))
    return off, len(contents)


def write_table(path, items):
    """Write a LevelDB-format sorted string table. `items`: iterable of (key bytes, value bytes) in ascending key order."""
    with open(path, "wb") as f:
        index = _BlockBuilder(restart_interval=1)
        block = _BlockBuilder()
        prev = None
        for key, value in items:
            if prev is not None and key <= prev:
                raise ValueError("table keys must be strictly ascending")
            prev = key
            block.add(key, value)
            if block.size() >= _BLOCK_BYTES:
                off, size = _write_block(f, block.finish())
                index.add(block.last_key, _varint(off) + _varint(size))   # separator = the block's last key
                block = _BlockBuilder()
        if block.count:
            off, size = _write_block(f, block.finish())
            index.add(block.last_key, _varint(off) + _varint(size))
        meta_handle = _write_block(f, _BlockBuilder().finish())
        index_handle = _write_block(f, index.finish())
        footer = _varint(meta_handle[0]) + _varint(meta_handle[1]) + _varint(index_handle[0]) + _varint(index_handle[1])
        footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
        f.write(footer)


def _read_block(buf, off, size, verify=True):
    contents = bytes(buf[off:off + size])
    ctype = buf[off + size]
    if verify:
        stored = struct.unpack_from("<I", buf, off + size + 1)[0]
        if unmask_crc(stored) != crc32c(contents + bytes([ctype])):
            raise ValueError("table block at offset %d: checksum mismatch" % off)
    if ctype != 0:
        raise NotImplementedError("table block at offset %d is compressed (type %d); only uncompressed tensor-bundle "
                                  "index files (what TensorFlow's BundleWriter produces) are supported" % (off, ctype))
    return contents


def _block_entries(contents):
    n_restarts = struct.unpack_from("<I", contents, len(contents) - 4)[0]
    end = len(contents) - 4 - 4 * n_restarts
    pos = 0
    key = b""
    while pos < end:
        shared, pos = _read_varint(contents, pos)
        non_shared, pos = _read_varint(contents, pos)
        vlen, pos = _read_varint(contents, pos)
        key = key[:shared] + contents[pos:pos + non_shared]
        pos += non_shared
        yield key, contents[pos:pos + vlen]
        pos += vlen


def read_table(path, verify=True):
    """All (key, value) pairs of a LevelDB-format table, in file order."""
    buf = memoryview(open(path, "rb").read())
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != _MAGIC:
        raise ValueError("%s is not a TensorFlow / LevelDB table (bad magic)" % path)
    footer = bytes(buf[len(buf) - 48:len(buf) - 8])
    pos = 0
    _, pos = _read_varint(footer, pos)
    _, pos = _read_varint(footer, pos)
    ioff, pos = _read_varint(footer, pos)
    isize, pos = _read_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        boff, p2 = _read_varint(handle, 0)
        bsize, _ = _read_varint(handle, p2)
        out.extend(_block_entries(_read_block(buf, boff, bsize, verify)))
    return out


# --------------------------------------------------------------------------------------------------------- bundle
def data_path(prefix, shard=0, num_shards=1):
    return "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)


def write_bundle(prefix, tensors):
    """Write `tensors` (name -> array-like) as `<prefix>.index` + `<prefix>.data-00000-of-00001`."""
    names = sorted(tensors, key=lambda n: n.encode("utf-8"))
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = [(b"", _encode_header())]
    offset = 0
    with open(data_path(prefix), "wb") as f:
        for name in names:
            if not name:
                raise ValueError("tensor names must be non-empty")
            a = np.asarray(tensors[name])
            if a.dtype.byteorder == ">":
                a = a.astype(a.dtype.newbyteorder("<"))
            if a.dtype not in _DTYPE_IDS:
                raise TypeError("tensor %r: dtype %s has no TensorFlow DataType here" % (name, a.dtype))
            shape = a.shape                      # (np.ascontiguousarray would turn a scalar into shape (1,))
            a = np.ascontiguousarray(a).reshape(-1)
            f.write(memoryview(a.view(np.uint8)) if a.size else b"")
            items.append((name.encode("utf-8"),
                          _encode_entry(_DTYPE_IDS[a.dtype], shape, offset, a.nbytes, mask_crc(crc32c(a)))))
            offset += a.nbytes
    write_table(prefix + ".index", items)


def list_bundle(prefix):
    """OrderedDict name -> entry dict (dtype id, shape, shard_id, offset, size, crc32c) + the header's shard count."""
    entries = OrderedDict()
    num_shards = 1
    for key, value in read_table(prefix + ".index"):
        if key == b"":
            for field, _, v in _pb_fields(value):
                if field == 1:
                    num_shards = v
                elif field == 2 and v != 0:
                    raise NotImplementedError("big-endian tensor bundles are not supported")
            continue
        entries[key.decode("utf-8")] = _decode_entry(value)
    return entries, num_shards


def read_bundle(prefix, names=None, verify=True):
    """Read a tensor bundle into an OrderedDict name -> ndarray (copies, native byte order).  `names` restricts the
    read; `verify` checks every tensor's CRC-32C the way TensorFlow's BundleReader does."""
    if not os.path.exists(prefix + ".index"):
        raise FileNotFoundError(prefix + ".index")
    entries, num_shards = list_bundle(prefix)
    maps = {}
    out = OrderedDict()
    for name, e in entries.items():
        if names is not None and name not in names:
            continue
        if e["slices"]:
            raise NotImplementedError("tensor %r is stored as slices of a partitioned variable" % name)
        if e["dtype"] not in _DTYPES:
            raise TypeError("tensor %r: TensorFlow DataType %d is not supported" % (name, e["dtype"]))
        sid = e["shard_id"]
        if sid not in maps:
            maps[sid] = np.memmap(data_path(prefix, sid, num_shards), dtype=np.uint8, mode="r")
        raw = maps[sid][e["offset"]:e["offset"] + e["size"]]
        dt = np.dtype(_DTYPES[e["dtype"]])
        count = int(np.prod(e["shape"])) if e["shape"] else 1
        if raw.size != e["size"] or count * dt.itemsize != e["size"]:
            raise ValueError("tensor %r: %d bytes on disk, shape %s of %s needs %d"
                             % (name, raw.size, e["shape"], dt, count * dt.itemsize))
        a = np.array(raw).view(dt).reshape(e["shape"])   # np.array: copy out of the map
        if verify and e["crc32c"] is not None and unmask_crc(e["crc32c"]) != crc32c(a):
            raise ValueError("tensor %r: checksum mismatch" % name)
        out[name] = a
    return out


def find_bundle(path):
    """Prefix of the tensor bundle behind `path`: a SavedModel directory (variables/variables), a train_saver directory
    (variables), or a prefix itself.  None if there is none."""
    cands = [os.path.join(path, "variables", "variables"), os.path.join(path, "variables"), path] \
        if os.path.isdir(path) else [path]
    for c in cands:
        if os.path.exists(c + ".index"):
            return c
    return None
