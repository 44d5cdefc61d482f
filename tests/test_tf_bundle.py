"""TensorFlow tensor-bundle files without TensorFlow (fcn8s_tensorflow_b200/tf_bundle.py): the reference stores every
weight in them (fcn8s_tensorflow.py:74,134 load a SavedModel's variables/variables.*; :922-934 write SavedModel /
tf.train.Saver files; :938-944 restore from a prefix).  TensorFlow is not installable here, so the pins are: the RFC 3720
CRC-32C vectors, a container assembled by hand in this file from the published format, and round trips."""
import os
import struct

import numpy as np
import pytest

from fcn8s_tensorflow_b200 import tf_bundle as tb


def test_crc32c_known_answers_native_and_fallback():
    # RFC 3720 B.4 + the classic check value
    vectors = [(b"123456789", 0xE3069283), (bytes(32), 0x8A9136AA), (b"\xff" * 32, 0x62A8AB43),
               (bytes(range(32)), 0x46DD794E), (bytes(range(31, -1, -1)), 0x113FDB5C)]
    for data, want in vectors:
        assert tb.crc32c(data) == want
        assert tb._crc32c_py(data) == want
    rng = np.random.default_rng(0)
    blob = rng.integers(0, 256, 100003, dtype=np.uint8)          # odd length, unaligned tail
    assert tb.crc32c(blob) == tb._crc32c_py(blob.tobytes())
    assert tb.crc32c(blob[1:]) == tb._crc32c_py(blob[1:].tobytes())      # unaligned start
    a, b = blob[:777], blob[777:]
    assert tb.crc32c(b, tb.crc32c(a)) == tb.crc32c(blob)                 # running value
    assert tb.crc32c(b"") == 0


def test_crc_mask_is_leveldbs():
    # rotr(crc, 15) + 0xa282ead8; unmask inverts it
    for crc in (0, 1, 0xE3069283, 0xFFFFFFFF, 0x80000000):
        m = tb.mask_crc(crc)
        assert m == ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF
        assert tb.unmask_crc(m) == crc
    assert tb.mask_crc(0) == 0xA282EAD8


def _block_with_trailer(contents):
    return contents + b"\x00" + struct.pack("<I", tb.mask_crc(tb.crc32c(contents + b"\x00")))


def test_index_file_matches_a_hand_assembled_table(tmp_path):
    """One float32 vector [1, 2] named "w": every byte of the .index file, put together here from the format."""
    prefix = str(tmp_path / "variables")
    w = np.array([1.0, 2.0], np.float32)
    tb.write_bundle(prefix, {"w": w})
    assert open(tb.data_path(prefix), "rb").read() == w.tobytes()

    header = b"\x08\x01" + b"\x1a\x02\x08\x01"                 # num_shards = 1; version { producer = 1 }
    shape = b"\x12\x02\x08\x02"                                # TensorShapeProto { dim { size: 2 } }
    entry = (b"\x08\x01"                                       # dtype = DT_FLOAT
             + b"\x12" + bytes([len(shape)]) + shape           # shape
             + b"\x28\x08"                                     # size = 8 (offset 0 and shard 0 are defaults)
             + b"\x35" + struct.pack("<I", tb.mask_crc(tb.crc32c(w.tobytes()))))
    # data block: entries (shared, non_shared, value_len, key suffix, value), restart array [0], count 1
    data = (b"\x00\x00" + bytes([len(header)]) + header
            + b"\x00\x01" + bytes([len(entry)]) + b"w" + entry
            + struct.pack("<I", 0) + struct.pack("<I", 1))
    meta = struct.pack("<I", 0) + struct.pack("<I", 1)         # empty block
    data_off, meta_off = 0, len(data) + 5
    index_off = meta_off + len(meta) + 5
    handle = bytes([data_off, len(data)])
    index = b"\x00\x01" + bytes([len(handle)]) + b"w" + handle + struct.pack("<I", 0) + struct.pack("<I", 1)
    footer = bytes([meta_off, len(meta), index_off, len(index)])
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    want = _block_with_trailer(data) + _block_with_trailer(meta) + _block_with_trailer(index) + footer
    assert open(prefix + ".index", "rb").read() == want
    # and the reader accepts the hand-made file as well as keys with shared prefixes written by someone else
    got = tb.read_bundle(prefix)
    assert list(got) == ["w"] and np.array_equal(got["w"], w)


def test_reader_follows_prefix_compression_and_restarts(tmp_path):
    """A table written with shared key prefixes inside a restart run (what TensorFlow's TableBuilder emits)."""
    entries = [(b"", b"H"), (b"conv1_1/biases", b"A"), (b"conv1_1/filter", b"B"), (b"conv1_2/biases", b"C")]
    blk = b"\x00\x00\x01H" + b"\x00\x0e\x01conv1_1/biasesA" + b"\x08\x06\x01filterB" + b"\x06\x08\x012/biasesC"
    blk += struct.pack("<I", 0) + struct.pack("<I", 1)
    meta = struct.pack("<I", 0) + struct.pack("<I", 1)
    handle = bytes([0, len(blk)])
    index = b"\x00\x01" + bytes([len(handle)]) + b"d" + handle + struct.pack("<I", 0) + struct.pack("<I", 1)
    meta_off = len(blk) + 5
    index_off = meta_off + len(meta) + 5
    footer = bytes([meta_off, len(meta), index_off, len(index)])
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    path = str(tmp_path / "t.index")
    open(path, "wb").write(_block_with_trailer(blk) + _block_with_trailer(meta) + _block_with_trailer(index) + footer)
    assert tb.read_table(path) == entries


def test_round_trip_many_tensors_dtypes_scalars_and_empties(tmp_path):
    rng = np.random.default_rng(1)
    tensors = {
        "conv1_1/filter": rng.standard_normal((3, 3, 3, 64)).astype(np.float32),
        "conv1_1/filter/Adam": rng.standard_normal((3, 3, 3, 64)).astype(np.float32),
        "conv1_1/biases": rng.standard_normal(64).astype(np.float32),
        "optimizer/global_step": np.asarray(12345, np.int32),
        "optimizer/beta1_power": np.asarray(0.9 ** 3, np.float32),
        "misc/int64": np.arange(7, dtype=np.int64) - 3,
        "misc/bool": np.array([[True, False], [False, True]]),
        "misc/f64": rng.standard_normal((2, 5)),
        "misc/empty": np.zeros((0, 4), np.float32),
        "misc/strided": rng.standard_normal((6, 6)).astype(np.float32)[::2, ::3],
    }
    for i in range(4000):   # > one 64 KB table block, long shared prefixes
        tensors["fc7_pool4_pool3_conv2d_trans/slot_%06d" % i] = np.full((2,), i, np.float32)
    prefix = str(tmp_path / "variables" / "variables")
    tb.write_bundle(prefix, tensors)
    assert sorted(os.listdir(tmp_path / "variables")) == ["variables.data-00000-of-00001", "variables.index"]
    got = tb.read_bundle(prefix)
    assert list(got) == sorted(tensors, key=lambda n: n.encode())
    for name, a in tensors.items():
        a = np.asarray(a)
        assert got[name].dtype == a.dtype and got[name].shape == a.shape, name
        assert np.array_equal(got[name], a), name
    some = tb.read_bundle(prefix, names={"conv1_1/biases", "optimizer/global_step"})
    assert set(some) == {"conv1_1/biases", "optimizer/global_step"} and int(some["optimizer/global_step"]) == 12345
    entries, shards = tb.list_bundle(prefix)
    assert shards == 1 and entries["conv1_1/filter"]["shape"] == [3, 3, 3, 64] and entries["conv1_1/filter"]["dtype"] == 1
    # offsets are contiguous in key order
    off = 0
    for name in got:
        assert entries[name]["offset"] == off
        off += entries[name]["size"]
    assert off == os.path.getsize(tb.data_path(prefix))


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "v")
    tb.write_bundle(prefix, {"a": np.arange(64, dtype=np.float32), "b": np.ones(3, np.float32)})
    blob = bytearray(open(tb.data_path(prefix), "rb").read())
    blob[17] ^= 0x40
    open(tb.data_path(prefix), "wb").write(blob)
    with pytest.raises(ValueError, match="tensor 'a': checksum mismatch"):
        tb.read_bundle(prefix)
    assert np.array_equal(tb.read_bundle(prefix, names={"b"})["b"], np.ones(3, np.float32))
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[10] ^= 0x01
    open(prefix + ".index", "wb").write(idx)
    with pytest.raises(ValueError, match="checksum mismatch"):
        tb.read_bundle(prefix)
    open(prefix + ".index", "wb").write(idx[:-3])
    with pytest.raises(ValueError, match="bad magic"):
        tb.read_bundle(prefix)
    with pytest.raises(FileNotFoundError):
        tb.read_bundle(str(tmp_path / "nothing"))


def test_weight_loader_finds_the_references_three_layouts(tmp_path):
    """SavedModel directory (variables/variables.*), train_saver directory (variables.*), bare prefix, legacy .npz."""
    from fcn8s_tensorflow_b200.fcn8s import _load_npz_weights
    w = {"fc7_1x1/bias": np.arange(5, dtype=np.float32), "fc7_1x1/kernel": np.ones((1, 1, 8, 5), np.float32)}
    sm = tmp_path / "saved_model_x"
    tb.write_bundle(str(sm / "variables" / "variables"), w)
    (sm / "saved_model.pb").write_bytes(b"")           # never parsed: the graph is the engine
    ts = tmp_path / "train_saver_x"
    tb.write_bundle(str(ts / "variables"), w)
    for path in (str(sm), str(ts), str(ts / "variables")):
        got = _load_npz_weights(path)
        assert set(got) == set(w) and all(np.array_equal(got[k], w[k]) for k in w), path
    np.savez(str(tmp_path / "variables.npz"), **w)
    legacy = tmp_path / "legacy"
    legacy.mkdir()
    os.rename(str(tmp_path / "variables.npz"), str(legacy / "variables.npz"))
    got = _load_npz_weights(str(legacy))
    assert all(np.array_equal(got[k], w[k]) for k in w)
    with pytest.raises(FileNotFoundError):
        _load_npz_weights(str(tmp_path))


def test_reader_handles_sharded_bundles(tmp_path):
    """tf.train.Saver(sharded=True) / large SavedModels spread the tensors over data-0000i-of-0000n files; the index
    names the shard of every tensor."""
    prefix = str(tmp_path / "variables")
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    b = np.arange(5, dtype=np.int64)
    open(tb.data_path(prefix, 0, 2), "wb").write(a.tobytes())
    open(tb.data_path(prefix, 1, 2), "wb").write(b"\x00" * 16 + b.tobytes())      # b starts at offset 16 of shard 1
    header = tb._pb_varint_field(1, 2) + tb._pb_bytes_field(3, tb._pb_varint_field(1, 1))    # num_shards = 2
    items = [(b"", header),
             (b"a", tb._encode_entry(1, a.shape, 0, a.nbytes, tb.mask_crc(tb.crc32c(a)))),
             (b"b", tb._encode_entry(9, b.shape, 16, b.nbytes, tb.mask_crc(tb.crc32c(b)), shard_id=1))]
    tb.write_table(prefix + ".index", items)
    entries, shards = tb.list_bundle(prefix)
    assert shards == 2 and entries["b"]["shard_id"] == 1 and entries["b"]["offset"] == 16
    got = tb.read_bundle(prefix)
    assert np.array_equal(got["a"], a) and np.array_equal(got["b"], b) and got["b"].dtype == np.int64


# ------------------------------------------------------------------------------------------------------------------
# Pins against TensorFlow-team code that IS available offline: the `tensorboard` wheel ships TensorFlow's own
# CRC-32C / mask implementation (tensorboard/compat/tensorflow_stub/pywrap_tensorflow.py, a transcription of
# tensorflow/core/lib/hash/crc32c.h) and protobuf modules generated from TensorFlow's .proto files
# (tensorboard/compat/proto).  tensor_bundle.proto itself is not among them, but every message it embeds is.
def _tb_stub():
    return pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")


def test_crc_and_mask_equal_tensorflows_own_implementation():
    stub = _tb_stub()
    rng = np.random.default_rng(3)
    for n in (0, 1, 7, 64, 1000, 65537):
        blob = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert tb.crc32c(blob) == stub.crc32c(blob)
        assert tb._crc32c_py(blob) == stub.crc32c(blob)
        assert tb.mask_crc(tb.crc32c(blob)) == stub.masked_crc32c(blob)


def test_dtype_ids_are_tensorflows_datatype_enum():
    types_pb2 = pytest.importorskip("tensorboard.compat.proto.types_pb2")
    names = {np.float32: "DT_FLOAT", np.float64: "DT_DOUBLE", np.int32: "DT_INT32", np.uint8: "DT_UINT8",
             np.int16: "DT_INT16", np.int8: "DT_INT8", np.int64: "DT_INT64", np.bool_: "DT_BOOL", np.uint16: "DT_UINT16",
             np.float16: "DT_HALF", np.uint32: "DT_UINT32", np.uint64: "DT_UINT64"}
    for np_t, name in names.items():
        assert tb._DTYPE_IDS[np.dtype(np_t)] == types_pb2.DataType.Value(name), name


def test_embedded_messages_parse_with_tensorflows_generated_protos():
    """BundleEntryProto.shape is a TensorShapeProto, BundleHeaderProto.version a VersionDef: the bytes this module
    writes for them are parsed by the classes generated from TensorFlow's own .proto files, and the bytes those classes
    serialise are what this module writes (field order included)."""
    shape_pb2 = pytest.importorskip("tensorboard.compat.proto.tensor_shape_pb2")
    versions_pb2 = pytest.importorskip("tensorboard.compat.proto.versions_pb2")
    for shape in ((), (7,), (3, 3, 512, 4096), (0, 5), (1 << 33, 2)):
        entry = tb._encode_entry(1, shape, 128, 4 * int(np.prod(shape)) if shape else 4, 0x12345678)
        fields = {f: v for f, _, v in tb._pb_fields(entry)}
        msg = shape_pb2.TensorShapeProto()
        msg.ParseFromString(bytes(fields[2]))
        assert [d.size for d in msg.dim] == list(shape) and not msg.unknown_rank
        ref = shape_pb2.TensorShapeProto(dim=[shape_pb2.TensorShapeProto.Dim(size=d) for d in shape])
        assert bytes(fields[2]) == ref.SerializeToString()
        assert tb._decode_entry(entry)["shape"] == list(shape)
    header = {f: v for f, _, v in tb._pb_fields(tb._encode_header())}
    ver = versions_pb2.VersionDef()
    ver.ParseFromString(bytes(header[3]))
    assert ver.producer == 1 and ver.min_consumer == 0 and not ver.bad_consumers
    assert bytes(header[3]) == versions_pb2.VersionDef(producer=1).SerializeToString()
    assert header[1] == 1            # num_shards
