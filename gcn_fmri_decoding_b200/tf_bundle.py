"""TensorFlow "tensor bundle" (checkpoint format V2) reader and writer in plain Python -- no TensorFlow needed.

The reference saves its models with ``tf.train.Saver`` (``/root/reference/lib_new/models_gcn.py:220``, written through
``lib_new/checkmat.py:43-84``); from TF 1.0 on that is the V2 format: ``<prefix>.index`` + ``<prefix>.data-SSSSS-of-NNNNN``.
With this module a reference-trained checkpoint loads straight into ``cgcnn`` (``load_tf_checkpoint``), and a model
trained here can be written in a form ``tf.train.load_checkpoint`` / ``Saver(var_list).restore`` accept
(``write_tf_checkpoint``).  Host code; nothing here is on the timed path.

PROVENANCE / VALIDATION: written from the published on-disk format (TensorFlow ``core/util/tensor_bundle``,
``core/lib/io/{table_builder,block_builder,format}.cc`` -- the LevelDB table layout -- and ``core/protobuf/
tensor_bundle.proto``).  TensorFlow cannot be installed in the build container and the reference ships no checkpoint,
so the code is validated by known-answer tests of its primitives (CRC-32C, varints, Snappy, prefix-compressed blocks)
and by write -> read round trips (``tests/test_tf_bundle.py``), NOT against a TensorFlow-written file.

Format, as implemented:

* ``.index`` is an immutable sorted table: data blocks, an (empty) meta-index block, an index block, a 48-byte footer
  (two block handles, padding, magic ``0xdb4775248b80fb57``).  A block = entries ``varint shared | varint non_shared |
  varint value_len | key suffix | value`` (keys prefix-compressed against the previous key, restart points every 16
  entries), then the restart offsets (uint32 each) and their count; on disk every block is followed by a 1-byte
  compression type (0 none, 1 Snappy) and the masked CRC-32C of block + type.
* key ``""`` -> ``BundleHeaderProto {num_shards = 1; endianness = 2; version = 3}``; every other key is a variable name ->
  ``BundleEntryProto {dtype = 1; shape = 2; shard_id = 3; offset = 4; size = 5; crc32c = 6 (fixed32, masked); slices = 7}``.
* ``.data-*`` files hold the raw little-endian bytes of the tensors at ``offset``.
"""
from __future__ import annotations

import os
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
RESTART_INTERVAL = 16
BLOCK_SIZE = 262144  # table::Options::block_size of the bundle writer

# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
DTYPE_CODES = {np.dtype(v): k for k, v in DTYPES.items()}


class BundleError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------ CRC-32C (Castagnoli)
def _crc_table():
    table = []
    for n in range(256):
        c = n
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        table.append(c)
    return table


_CRC_LIST = _crc_table()


def crc32c(data, crc=0):
    """CRC-32C of ``data`` (bytes-like); ``crc32c(b"123456789") == 0xE3069283``."""
    c = crc ^ 0xFFFFFFFF
    tab = _CRC_LIST
    for b in bytes(data):
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc):
    """``crc32c::Mask``: rotate right by 15 bits and add a constant (CRCs of data that embeds CRCs)."""
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def unmask_crc(masked):
    rot = (masked - 0xA282EAD8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ varints / protobuf
def put_varint(n):
    n = int(n)
    if n < 0:
        n += 1 << 64  # two's complement, as protobuf encodes negative int64
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def get_varint(buf, pos):
    shift = result = 0
    while True:
        if pos >= len(buf):
            raise BundleError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise BundleError("varint longer than 64 bits")


def _pb_fields(buf):
    """Yield ``(field number, wire type, value)`` of a serialized protobuf message (values: int or bytes)."""
    pos = 0
    while pos < len(buf):
        tag, pos = get_varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            val, pos = get_varint(buf, pos)
        elif wire == 1:
            val, pos = struct.unpack_from("<Q", buf, pos)[0], pos + 8
        elif wire == 2:
            n, pos = get_varint(buf, pos)
            if pos + n > len(buf):
                raise BundleError("truncated length-delimited field")
            val, pos = bytes(buf[pos:pos + n]), pos + n
        elif wire == 5:
            val, pos = struct.unpack_from("<I", buf, pos)[0], pos + 4
        else:
            raise BundleError("unsupported protobuf wire type %d" % wire)
        yield field, wire, val


def _pb_field(field, wire, payload):
    head = put_varint((field << 3) | wire)
    if wire == 0:
        return head + put_varint(payload)
    if wire == 2:
        return head + put_varint(len(payload)) + payload
    if wire == 5:
        return head + struct.pack("<I", payload)
    raise ValueError(wire)


def _signed64(v):
    return v - (1 << 64) if v >= 1 << 63 else v


def _parse_entry(value):
    """BundleEntryProto -> dict(dtype, shape, shard_id, offset, size, crc32c, sliced)."""
    e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for field, _, val in _pb_fields(value):
        if field == 1:
            e["dtype"] = val
        elif field == 2:
            dims = []
            for f2, _, v2 in _pb_fields(val):
                if f2 == 2:  # Dim
                    size = 0
                    for f3, _, v3 in _pb_fields(v2):
                        if f3 == 1:
                            size = _signed64(v3)
                    dims.append(size)
            e["shape"] = tuple(dims)
        elif field == 3:
            e["shard_id"] = val
        elif field == 4:
            e["offset"] = val
        elif field == 5:
            e["size"] = val
        elif field == 6:
            e["crc32c"] = val
        elif field == 7:
            e["sliced"] = True
    return e


def _entry_bytes(dtype_code, shape, shard_id, offset, size, masked_crc):
    dims = b"".join(_pb_field(2, 2, _pb_field(1, 0, d) if d else b"") for d in shape)
    out = _pb_field(1, 0, dtype_code) + _pb_field(2, 2, dims)
    if shard_id:
        out += _pb_field(3, 0, shard_id)
    if offset:
        out += _pb_field(4, 0, offset)
    out += _pb_field(5, 0, size) + _pb_field(6, 5, masked_crc)
    return out


# ------------------------------------------------------------------------------------------------ Snappy (raw format)
def snappy_uncompress(buf):
    """Raw Snappy block decoder (the index may be Snappy-compressed by writers other than the bundle writer)."""
    n, pos = get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:  # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            if pos + ln > len(buf):
                raise BundleError("snappy: truncated literal")
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = 4 + ((tag >> 2) & 7)
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise BundleError("snappy: bad copy offset")
        for _ in range(ln):  # byte by byte: copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise BundleError("snappy: length mismatch (%d != %d)" % (len(out), n))
    return bytes(out)


# ------------------------------------------------------------------------------------------------ table: reading
def _read_block(buf, offset, size, verify):
    if offset + size + 5 > len(buf):
        raise BundleError("block handle past the end of the index file")
    body, kind = buf[offset:offset + size], buf[offset + size]
    if verify:
        stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
        if unmask_crc(stored) != crc32c(buf[offset:offset + size + 1]):
            raise BundleError("index block checksum mismatch at offset %d" % offset)
    if kind == 0:
        return bytes(body)
    if kind == 1:
        return snappy_uncompress(bytes(body))
    raise BundleError("unknown block compression type %d" % kind)


def _block_entries(block):
    """Key/value pairs of a decoded block, undoing the prefix compression."""
    if len(block) < 4:
        raise BundleError("block too short")
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    if end < 0:
        raise BundleError("bad restart array")
    pos, key = 0, b""
    while pos < end:
        shared, pos = get_varint(block, pos)
        non_shared, pos = get_varint(block, pos)
        vlen, pos = get_varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > end:
            raise BundleError("corrupt block entry")
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_index(prefix, verify_checksums=True):
    """``(header dict, {name: entry dict})`` of ``<prefix>.index``."""
    path = prefix + ".index"
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise BundleError("%s is not a TensorFlow V2 checkpoint index (bad table magic)" % path)
    footer = buf[len(buf) - 48:]
    _, pos = get_varint(footer, 0)          # meta-index handle (unused)
    _, pos = get_varint(footer, pos)
    ioff, pos = get_varint(footer, pos)     # index handle
    isize, pos = get_varint(footer, pos)
    header, entries = None, {}
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify_checksums)):
        off, p = get_varint(handle, 0)
        size, p = get_varint(handle, p)
        for key, value in _block_entries(_read_block(buf, off, size, verify_checksums)):
            if key == b"":
                header = dict(num_shards=0, endianness=0)
                for field, _, val in _pb_fields(value):
                    if field == 1:
                        header["num_shards"] = val
                    elif field == 2:
                        header["endianness"] = val
            else:
                entries[key.decode("utf-8")] = _parse_entry(value)
    if header is None:
        raise BundleError("%s has no bundle header entry" % path)
    if header["endianness"] != 0:
        raise BundleError("big-endian bundles are not supported")
    return header, entries


def read_tf_checkpoint(prefix, names=None, verify_checksums=True):
    """Tensors of the checkpoint ``<prefix>.index`` + ``<prefix>.data-*`` as ``{variable name: ndarray}``.

    ``names``: only these variables (default: all of numeric dtype; string tensors and sliced (partitioned) variables
    are skipped unless asked for by name, in which case they raise)."""
    header, entries = read_index(prefix, verify_checksums)
    out, files = {}, {}
    try:
        for name in (sorted(entries) if names is None else names):
            if name not in entries:
                raise BundleError("variable %r is not in %s.index (has: %s)" % (name, prefix, ", ".join(sorted(entries))))
            e = entries[name]
            if e["sliced"] or e["dtype"] not in DTYPES:
                if names is None:
                    continue
                raise BundleError("variable %r: sliced or non-numeric tensors (dtype %d) are not supported" % (name, e["dtype"]))
            dtype = np.dtype(DTYPES[e["dtype"]])
            count = int(np.prod(e["shape"], dtype=np.int64)) if len(e["shape"]) else 1
            if count * dtype.itemsize != e["size"]:
                raise BundleError("variable %r: %d bytes stored for shape %s of %s" % (name, e["size"], e["shape"], dtype))
            fn = "%s.data-%05d-of-%05d" % (prefix, e["shard_id"], header["num_shards"])
            if fn not in files:
                files[fn] = open(fn, "rb")
            files[fn].seek(e["offset"])
            raw = files[fn].read(e["size"])
            if len(raw) != e["size"]:
                raise BundleError("variable %r: %s is truncated" % (name, fn))
            if verify_checksums and e["crc32c"] is not None and unmask_crc(e["crc32c"]) != crc32c(raw):
                raise BundleError("variable %r: data checksum mismatch" % name)
            out[name] = np.frombuffer(raw, dtype.newbyteorder("<")).astype(dtype).reshape(e["shape"])
    finally:
        for f in files.values():
            f.close()
    return out


# ------------------------------------------------------------------------------------------------ table: writing
def _build_block(items, restart_interval):
    out, restarts, last, count = bytearray(), [], b"", 0
    for key, value in items:
        if count % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            for a, b in zip(last, key):
                if a != b:
                    break
                shared += 1
        out += put_varint(shared) + put_varint(len(key) - shared) + put_varint(len(value)) + key[shared:] + value
        last = key
        count += 1
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _emit_block(fileobj, block):
    offset = fileobj.tell()
    fileobj.write(block)
    fileobj.write(b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))  # no compression
    return put_varint(offset) + put_varint(len(block))


def write_tf_checkpoint(prefix, tensors):
    """Write ``{variable name: array}`` as a one-shard V2 checkpoint ``<prefix>.index`` + ``<prefix>.data-00000-of-00001``."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = [(b"", None)]
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as data:
        for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
            if not name:
                raise ValueError("variable names must not be empty")
            a = np.asarray(tensors[name])  # (ascontiguousarray would turn a scalar into shape (1,); tobytes is C order)
            if a.dtype not in DTYPE_CODES:
                raise ValueError("variable %r: dtype %s is not supported" % (name, a.dtype))
            raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
            data.write(raw)
            items.append((name.encode("utf-8"), _entry_bytes(DTYPE_CODES[a.dtype], a.shape, 0, offset, len(raw),
                                                             mask_crc(crc32c(raw)))))
            offset += len(raw)
    version = _pb_field(1, 0, 1)  # VersionDef.producer = kTensorBundleVersion
    items[0] = (b"", _pb_field(1, 0, 1) + _pb_field(3, 2, version))  # num_shards = 1, little endian (0) is the default
    with open(prefix + ".index", "wb") as f:
        index, chunk, size = [], [], 0
        for key, value in items:
            chunk.append((key, value))
            size += len(key) + len(value) + 3
            if size >= BLOCK_SIZE:
                index.append((chunk[-1][0], _emit_block(f, _build_block(chunk, RESTART_INTERVAL))))
                chunk, size = [], 0
        if chunk:
            index.append((chunk[-1][0], _emit_block(f, _build_block(chunk, RESTART_INTERVAL))))
        meta = _emit_block(f, _build_block([], RESTART_INTERVAL))
        idx = _emit_block(f, _build_block(index, 1))
        footer = meta + idx
        f.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
    return prefix


# ------------------------------------------------------------------------------------------------ model <-> checkpoint
def load_tf_checkpoint(model, prefix, verify_checksums=True):
    """Load a reference-trained TF checkpoint into a ``cgcnn``: the variables ``conv{i}/weights|bias``, ``fc{i}/...``,
    ``logits/...`` (models_gcn.py:662,343,351,675,680); optimiser slots (``.../Adam``, ``beta1_power``) and
    ``global_step`` in the file are ignored."""
    names = list(model.state_dict_tf())
    model.load_state_dict_tf(read_tf_checkpoint(prefix, names=names, verify_checksums=verify_checksums))
    return model


def save_tf_checkpoint(model, prefix, step=None):
    """Write the model's parameters (and ``global_step`` when given) as a TF V2 checkpoint; returns the prefix."""
    tensors = dict(model.state_dict_tf())
    if step is not None:
        prefix = "%s-%d" % (prefix, int(step))
        tensors["global_step"] = np.asarray(int(step), np.int32)
    return write_tf_checkpoint(prefix, tensors)
