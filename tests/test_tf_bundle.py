"""tf_bundle.py: reader / writer of TensorFlow's V2 checkpoint format (host code, no GPU, no TensorFlow).

The primitives are pinned by published known-answer vectors (CRC-32C: RFC 3720 B.4; protobuf varints; Snappy's raw
format); the container itself by write -> read round trips and by decoding the written bytes by hand.  No
TensorFlow-written file exists in this environment -- see the module's header."""
import os
import struct

import numpy as np
import pytest

from gcn_fmri_decoding_b200 import tf_bundle as tb


def test_crc32c_known_answers_and_mask():
    assert tb.crc32c(b"123456789") == 0xE3069283
    assert tb.crc32c(bytes(32)) == 0x8A9136AA                       # RFC 3720 B.4
    assert tb.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tb.crc32c(bytes(range(32))) == 0x46DD794E
    assert tb.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert tb.crc32c(b"world", tb.crc32c(b"hello ")) == tb.crc32c(b"hello world")   # extends like crc32c::Extend
    c = tb.crc32c(b"foo")
    assert tb.mask_crc(c) != c and tb.mask_crc(tb.mask_crc(c)) != c
    assert tb.unmask_crc(tb.mask_crc(c)) == c and tb.unmask_crc(tb.unmask_crc(tb.mask_crc(tb.mask_crc(c)))) == c


def test_varints_and_protobuf_fields():
    assert tb.put_varint(0) == b"\x00" and tb.put_varint(127) == b"\x7f" and tb.put_varint(300) == b"\xac\x02"
    assert tb.put_varint(-1) == b"\xff" * 9 + b"\x01"
    for v in (0, 1, 127, 128, 16383, 16384, 2 ** 32, 2 ** 63 - 1):
        assert tb.get_varint(tb.put_varint(v) + b"junk", 0) == (v, len(tb.put_varint(v)))
    with pytest.raises(tb.BundleError):
        tb.get_varint(b"\x80\x80", 0)
    # BundleEntryProto of a float32 [75, 32] tensor at offset 128: decoded field by field
    raw = tb._entry_bytes(1, (75, 32), 0, 128, 9600, 0xDEADBEEF)
    assert raw == (b"\x08\x01" b"\x12\x08" b"\x12\x02\x08\x4b" b"\x12\x02\x08\x20" b"\x20\x80\x01" b"\x28\x80\x4b"
                   b"\x35\xef\xbe\xad\xde")
    e = tb._parse_entry(raw)
    assert e == dict(dtype=1, shape=(75, 32), shard_id=0, offset=128, size=9600, crc32c=0xDEADBEEF, sliced=False)


def test_snappy_raw_format():
    assert tb.snappy_uncompress(b"\x0c\x08abc\x15\x03") == b"abc" * 4            # literal + 1-byte-offset copy (overlapping)
    lit = bytes(range(70))
    assert tb.snappy_uncompress(bytes([70, 60 << 2, 69]) + lit) == lit             # literal with a one-byte length
    two = tb.snappy_uncompress(tb.put_varint(140) + bytes([60 << 2, 69]) + lit + bytes([(63 << 2) | 2, 70, 0, (5 << 2) | 2, 70, 0]))
    assert two == lit + lit                                                        # 2-byte-offset copies of 64 + 6 bytes
    with pytest.raises(tb.BundleError):
        tb.snappy_uncompress(b"\x05\x08abc")                                       # declared length does not match
    with pytest.raises(tb.BundleError):
        tb.snappy_uncompress(b"\x08\x00a\x15\x09")                                 # copy from before the start


def _tensors():
    rng = np.random.RandomState(0)
    t = {}
    for scope, shape in (("conv1", (75, 32)), ("conv2", (160, 32)), ("fc1", (25, 512)), ("fc2", (512, 256)), ("logits", (256, 22))):
        for suffix in ("", "/Adam", "/Adam_1"):   # what a reference checkpoint holds: variables + optimiser slots
            t[scope + "/weights" + suffix] = rng.randn(*shape).astype(np.float32)
            t[scope + "/bias" + suffix] = rng.randn(shape[1]).astype(np.float32)
    t["beta1_power"] = np.asarray(0.9 ** 7, np.float32)
    t["beta2_power"] = np.asarray(0.999 ** 7, np.float32)
    t["global_step"] = np.asarray(7, np.int32)
    t["counts"] = np.arange(-3, 9, dtype=np.int64).reshape(3, 4)
    t["flags"] = np.array([True, False, True])
    t["empty"] = np.zeros((0, 4), np.float64)
    return t


def test_round_trip_and_on_disk_layout(tmp_path):
    t = _tensors()
    prefix = tb.write_tf_checkpoint(str(tmp_path / "ckpt" / "model-7"), t)
    assert sorted(os.listdir(tmp_path / "ckpt")) == ["model-7.data-00000-of-00001", "model-7.index"]
    got = tb.read_tf_checkpoint(prefix)
    assert sorted(got) == sorted(t)
    for k, v in t.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    only = tb.read_tf_checkpoint(prefix, names=["fc1/bias", "global_step"])
    assert list(only) == ["fc1/bias", "global_step"] and int(only["global_step"]) == 7
    header, entries = tb.read_index(prefix)
    assert header == dict(num_shards=1, endianness=0) and len(entries) == len(t)
    # the data file is the tensors back to back in key order
    order = sorted(t, key=lambda s: s.encode())
    off = 0
    for k in order:
        assert entries[k]["offset"] == off and entries[k]["size"] == t[k].nbytes
        off += t[k].nbytes
    assert os.path.getsize(prefix + ".data-00000-of-00001") == off
    # the index file decoded by hand: footer, index block with one data block, first entries of the data block
    buf = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", buf[-8:])[0] == 0xDB4775248B80FB57 and len(buf[-48:]) == 48
    moff, p = tb.get_varint(buf[-48:], 0)
    msize, p = tb.get_varint(buf[-48:], p)
    ioff, p = tb.get_varint(buf[-48:], p)
    isize, p = tb.get_varint(buf[-48:], p)
    assert buf[moff:moff + msize] == struct.pack("<II", 0, 1)            # empty meta-index block: one restart at 0
    assert ioff == moff + msize + 5 and ioff + isize + 5 == len(buf) - 48  # blocks carry a 5-byte trailer
    assert buf[0:3] == b"\x00\x00" + bytes([len(b"\x08\x01\x1a\x02\x08\x01")]) and buf[3:9] == b"\x08\x01\x1a\x02\x08\x01"
    # second entry: key "beta1_power" shares nothing with ""; third "beta2_power" shares "beta" (prefix compression)
    pos = 9
    assert buf[pos:pos + 2] == bytes([0, 11]) and buf[pos + 3:pos + 14] == b"beta1_power"
    vlen = buf[pos + 2]
    pos += 3 + 11 + vlen
    assert buf[pos:pos + 2] == bytes([4, 7]) and buf[pos + 3:pos + 10] == b"2_power"


def test_corruption_is_detected(tmp_path):
    t = _tensors()
    prefix = tb.write_tf_checkpoint(str(tmp_path / "m"), t)
    with pytest.raises(tb.BundleError, match="not in"):
        tb.read_tf_checkpoint(prefix, names=["conv9/weights"])
    data = prefix + ".data-00000-of-00001"
    raw = bytearray(open(data, "rb").read())
    raw[100] ^= 0x40
    open(data, "wb").write(bytes(raw))
    with pytest.raises(tb.BundleError, match="data checksum"):
        tb.read_tf_checkpoint(prefix)
    assert len(tb.read_tf_checkpoint(prefix, verify_checksums=False)) == len(t)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[20] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(tb.BundleError, match="checksum"):
        tb.read_index(prefix)
    open(prefix + ".index", "wb").write(bytes(idx[:-1]) + b"\x00")
    with pytest.raises(tb.BundleError, match="magic"):
        tb.read_index(prefix)


def test_many_variables_span_restart_points_and_blocks(tmp_path, monkeypatch):
    monkeypatch.setattr(tb, "BLOCK_SIZE", 600)  # force several data blocks
    t = {"layer%03d/kernel/part_%d" % (i, i % 3): np.full((i % 5 + 1,), i, np.float32) for i in range(150)}
    prefix = tb.write_tf_checkpoint(str(tmp_path / "many"), t)
    got = tb.read_tf_checkpoint(prefix)
    assert sorted(got) == sorted(t) and all(np.array_equal(got[k], v) for k, v in t.items())


def test_model_checkpoint_in_tf_format(graph_l4, tmp_path):
    """A cgcnn written as a TF checkpoint and loaded into a differently seeded one; optimiser slots in
    the file are ignored by the loader."""
    from gcn_fmri_decoding_b200.models import cgcnn

    def make(seed):
        return cgcnn(L=graph_l4["L"], F=[32, 32], K=[5, 5], p=[4, 4], M=[512, 256, 22], channel=15, device="cpu", seed=seed)

    a, b = make(1), make(2)
    prefix = tb.save_tf_checkpoint(a, str(tmp_path / "best.ckpt"), step=1200)
    assert prefix.endswith("best.ckpt-1200") and os.path.exists(prefix + ".index")
    extra = dict(tb.read_tf_checkpoint(prefix))
    extra["conv1/weights/Adam"] = np.zeros((75, 32), np.float32)
    tb.write_tf_checkpoint(prefix, extra)
    tb.load_tf_checkpoint(b, prefix)
    for k, v in a.state_dict_tf().items():
        assert np.array_equal(v, b.state_dict_tf()[k]), k
    from gcn_fmri_decoding_b200 import checkpoints

    c = checkpoints.load_checkpoint(make(3), prefix)   # the generic loader recognises a TF prefix
    tb.save_tf_checkpoint(a, str(tmp_path / "best.ckpt"), step=300)
    assert checkpoints.latest_checkpoint(str(tmp_path), prefix="best.ckpt") == prefix   # highest step, TF files included
    assert all(np.array_equal(v, c.state_dict_tf()[k]) for k, v in a.state_dict_tf().items())
    assert int(tb.read_tf_checkpoint(prefix, names=["global_step"])["global_step"]) == 1200
