"""Reader (and a minimal writer) for TensorFlow V2 checkpoint bundles without TensorFlow.

The reference restores ``bn_17.sn_360.epoch_N.ckpt`` with ``tf.train.Saver().restore``
(call_modifications.py:210-211); a V2 bundle is
  <prefix>.index                  a LevelDB-format SSTable: key = variable name, value = BundleEntryProto
                                  (key "" holds the BundleHeaderProto)
  <prefix>.data-00000-of-00001    the raw little-endian tensor bytes, addressed by (offset, size)
Format restated from the published TensorFlow sources (tensorflow/core/util/tensor_bundle/, core/lib/io/
table format = LevelDB's): data blocks of prefix-compressed entries + restart array, each followed by a
1-byte compression type and a masked crc32c; a 48-byte footer with the metaindex / index block handles and the
magic 0xdb4775248b80fb57.  Blocks may be stored raw (type 0) or Snappy-compressed (type 1: TF's table builder
compresses a block when that saves >= 12.5 %, which an index full of repetitive variable names usually does), so the
reader carries a pure-Python decoder of the Snappy raw format (varint length, then literal / copy elements) and
verifies every block's masked crc32c.  TensorFlow is not installable here and no real checkpoint is reachable, so this
module is verified by write->read round trips (raw and compressed, prefix-shared keys, multi-block indexes), by
hand-assembled Snappy streams from the format description and by a second, bitwise crc32c implementation in the
tests: PARITY UNPINNED against a TF-written file.
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = 0xDB4775248B80FB57
DT_FLOAT, DT_INT32, DT_INT64 = 1, 3, 9
_DTYPES = {DT_FLOAT: np.dtype("<f4"), DT_INT32: np.dtype("<i4"), DT_INT64: np.dtype("<i8"), 2: np.dtype("<f8")}


def _varint(buf, pos):
    out = shift = 0
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
    """Flat protobuf walk -> list of (field, wire_type, value)."""
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type {}".format(wt))
        out.append((f, wt, v))
    return out


def _parse_entry(val):
    """BundleEntryProto: 1 dtype, 2 shape{2 dim{1 size}}, 3 shard_id, 4 offset, 5 size, 6 crc32c, 7 slices."""
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, sliced=False)
    for f, wt, v in _parse_proto(val):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            for f2, _, v2 in _parse_proto(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _parse_proto(v2):
                        if f3 == 1:
                            size = v3
                    e["shape"].append(size)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 7:
            e["sliced"] = True
    return e


def _crc32c_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TAB = _crc32c_table()


def crc32c(data, crc=0):
    crc ^= 0xFFFFFFFF
    for b in data:
        crc = _CRC_TAB[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _mask(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def snappy_decompress(buf):
    """Snappy raw format (the block format, not the framing format): varint uncompressed length, then elements whose
    tag byte's low two bits select literal (00), copy with 1-byte offset (01), 2-byte offset (10), 4-byte offset (11)."""
    n, pos = _varint(buf, 0)
    out = bytearray()
    end = len(buf)
    while pos < end:
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            if pos + ln > end:
                raise ValueError("snappy: literal runs past the end of the block")
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = 4 + ((tag >> 2) & 7)
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = 1 + (tag >> 2)
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:
            ln = 1 + (tag >> 2)
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("snappy: copy offset {} outside the {} bytes produced so far".format(off, len(out)))
        start = len(out) - off
        if off >= ln:
            out += out[start:start + ln]
        else:                               # overlapping copy: the pattern repeats
            for i in range(ln):
                out.append(out[start + i])
    if len(out) != n:
        raise ValueError("snappy: produced {} bytes, header says {}".format(len(out), n))
    return bytes(out)


def _read_block(data, offset, size, verify=True):
    if offset + size + 5 > len(data):
        raise ValueError("SSTable block handle ({}, {}) points outside the file".format(offset, size))
    block = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        stored = struct.unpack_from("<I", data, offset + size + 1)[0]
        if stored != _mask(crc32c(data[offset:offset + size + 1])):
            raise ValueError("SSTable block at {} fails its crc32c check".format(offset))
    if ctype == 1:
        block = snappy_decompress(block)
    elif ctype != 0:
        raise ValueError("SSTable block compression type {} is not supported (0 = raw, 1 = snappy)".format(ctype))
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < limit:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(block[pos:pos + vlen])))
        pos += vlen
    return out


def read_index(prefix):
    with open(prefix + ".index", "rb") as fh:
        data = fh.read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != MAGIC:
        raise ValueError("{}.index is not a TensorFlow bundle index (bad magic)".format(prefix))
    footer = data[-48:]
    pos = 0
    _, pos = _varint(footer, pos)          # metaindex handle
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    entries = {}
    for _, handle in _read_block(data, idx_off, idx_size):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        for key, val in _read_block(data, boff, bsize):
            if key == b"":
                continue                    # BundleHeaderProto
            entries[key.decode()] = _parse_entry(val)
    return entries


def read_bundle(prefix, names=None):
    """{variable name: ndarray} for every (or the named) non-sliced tensor of a single-shard bundle."""
    entries = read_index(prefix)
    shards = sorted({e["shard_id"] for e in entries.values()})
    n_shards = (max(shards) + 1) if shards else 1
    files = {}
    out = {}
    for name, e in entries.items():
        if names is not None and name not in names:
            continue
        if e["sliced"] or e["dtype"] not in _DTYPES:
            continue
        sid = e["shard_id"]
        if sid not in files:
            files[sid] = np.memmap("{}.data-{:05d}-of-{:05d}".format(prefix, sid, n_shards), dtype=np.uint8, mode="r")
        raw = files[sid][e["offset"]:e["offset"] + e["size"]]
        out[name] = np.frombuffer(bytes(raw), dtype=_DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    return out


# ---- writer (tests / converting an .npz to a bundle) -------------------------------------------------------
def _block(entries, restart_interval=16):
    """LevelDB block: entries with keys prefix-compressed against the previous key, a full key every
    ``restart_interval`` entries (TF / LevelDB default 16), then the restart offsets and their count."""
    out = bytearray()
    restarts = []
    prev = b""
    for i, (k, v) in enumerate(entries):
        if i % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts.append(0)
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def snappy_compress(data):
    """Small greedy Snappy encoder (4-byte hash matches, copies with 2-byte offsets, literals otherwise) used to write
    compressed test bundles; any conforming decoder reads its output."""
    out = bytearray(_put_varint(len(data)))
    table = {}
    pos = lit = 0
    n = len(data)

    def emit_literal(lo, hi):
        while lo < hi:
            ln = min(hi - lo, 1 << 16)
            if ln <= 60:
                out.append((ln - 1) << 2)
            elif ln <= 256:
                out.extend(bytes([60 << 2, ln - 1]))
            else:
                out.extend(bytes([61 << 2]) + struct.pack("<H", ln - 1))
            out.extend(data[lo:lo + ln])
            lo += ln

    while pos + 4 <= n:
        key = data[pos:pos + 4]
        cand = table.get(key)
        table[key] = pos
        if cand is not None and pos - cand < 65536:
            ln = 4
            while pos + ln < n and ln < 64 and data[cand + ln] == data[pos + ln]:
                ln += 1
            emit_literal(lit, pos)
            out.extend(bytes([((ln - 1) << 2) | 2]) + struct.pack("<H", pos - cand))
            pos += ln
            lit = pos
        else:
            pos += 1
    emit_literal(lit, n)
    return bytes(out)


def _entry_proto(dtype, shape, offset, size, crc):
    dims = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(s) for s in shape))
    msg = b"\x08" + _put_varint(dtype)
    msg += b"\x12" + _put_varint(len(dims)) + dims
    if offset:
        msg += b"\x20" + _put_varint(offset)
    msg += b"\x28" + _put_varint(size)
    msg += b"\x35" + struct.pack("<I", crc)
    return msg


def write_bundle(prefix, tensors: dict, compress=False, block_entries=0):
    """Single-shard float32/int32 bundle with the TF file names; enough for round trips and for handing a
    synthetic checkpoint to tools that expect the ``.ckpt`` prefix.  ``compress`` stores the blocks Snappy-compressed
    (type 1) like TF's table builder does when it pays; ``block_entries`` > 0 splits the entries over several data blocks."""
    data = bytearray()
    kv = []
    for name in sorted(tensors):
        a = np.asarray(tensors[name], order="C")
        dt = {np.dtype("float32"): DT_FLOAT, np.dtype("int32"): DT_INT32, np.dtype("int64"): DT_INT64}[a.dtype]
        raw = a.tobytes()
        kv.append((name.encode(), _entry_proto(dt, a.shape, len(data), len(raw), _mask(crc32c(raw[:4096])) if len(raw) <= 4096 else 0)))
        data += raw
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"          # num_shards = 1, version{producer = 1}
    blocks = bytearray()

    def add_block(entries):
        body = _block(entries)
        ctype = b"\x00"
        if compress:
            body, ctype = snappy_compress(body), b"\x01"
        off = len(blocks)
        blocks.extend(body)
        blocks.extend(ctype + struct.pack("<I", _mask(crc32c(body + ctype))))
        return off, len(body)

    all_entries = [(b"", header)] + kv
    per = block_entries if block_entries > 0 else len(all_entries)
    index_entries = []
    for lo in range(0, len(all_entries), per):
        part = all_entries[lo:lo + per]
        d_off, d_size = add_block(part)
        index_entries.append((part[-1][0], _put_varint(d_off) + _put_varint(d_size)))
    m_off, m_size = add_block([])
    i_off, i_size = add_block(index_entries)
    footer = _put_varint(m_off) + _put_varint(m_size) + _put_varint(i_off) + _put_varint(i_size)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
    with open(prefix + ".index", "wb") as fh:
        fh.write(bytes(blocks) + footer)
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        fh.write(bytes(data))
