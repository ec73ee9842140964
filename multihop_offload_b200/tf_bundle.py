"""TensorFlow "tensor bundle" checkpoints, read and written without TensorFlow.

Keeps ``ACOAgent.load`` / ``ACOAgent.save`` (src/gnn_offloading_agent.py:125-132) drop-in: the
reference calls ``tf.train.latest_checkpoint(dir)`` + ``model.load_weights`` and
``model.save_weights(dir/cp-XXXX.ckpt)``.  Layout:

* ``<prefix>.index``  - LevelDB-style SSTable: one data block of prefix-compressed
  (key -> BundleEntryProto) records, restart array, 5-byte block trailers (type 0 + masked
  crc32c), an empty metaindex block, an index block, 48-byte footer with magic db4775248b80fb57.
* ``<prefix>.data-00000-of-00001`` - raw little-endian tensors in variable-creation order
  (kernel, bias per layer), then the serialized ``TrackableObjectGraph`` string tensor.
* ``checkpoint`` - text file naming the latest prefix.

The object-graph blob carries no shapes and is byte-identical in both shipped checkpoints, so the
writer re-emits it verbatim (``assets/object_graph.bin``); it is valid for any Chebyshev order K.
"""
from __future__ import annotations

import os
import struct

import numpy as np

MAGIC = bytes.fromhex("57fb808b247547db")
DT_FLOAT, DT_DOUBLE, DT_STRING = 1, 2, 7
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
_ASSET = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "object_graph.bin")
OBJECT_GRAPH_CRC = 0x42E62450  # masked crc TF stored for that blob


# ---- crc32c (Castagnoli), masked as in tensorflow/core/lib/hash/crc32c.h ------------------------
def _make_table():
    tbl = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tbl.append(c)
    return tbl


_TBL = _make_table()


def crc32c(data, crc=0):
    c = crc ^ 0xFFFFFFFF
    for b in bytes(data):
        c = _TBL[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(c):
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- varints / tiny protobuf helpers -------------------------------------------------------------
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


def _get_varint(buf, pos):
    out = shift = 0
    while True:
        c = buf[pos]; pos += 1
        out |= (c & 0x7F) << shift
        if c < 0x80:
            return out, pos
        shift += 7


def _fields(buf):
    pos, out = 0, []
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln]); pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]; pos += 4
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]; pos += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.append((f, wt, v))
    return out


def _entry_proto(dtype, shape, offset, size, crc):
    shp = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(s) for s in shape))
    out = b"\x08" + _put_varint(dtype) + b"\x12" + _put_varint(len(shp)) + shp
    if offset:
        out += b"\x20" + _put_varint(offset)
    if size:
        out += b"\x28" + _put_varint(size)
    out += b"\x35" + struct.pack("<I", crc)
    return out


# ---- SSTable ---------------------------------------------------------------------------------------
def _block(entries, restart_interval=16):
    """entries: sorted [(key bytes, value bytes)] -> block contents (without trailer)."""
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            m = min(len(prev), len(k))
            while shared < m and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _with_trailer(block):
    return block + b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00")))


def _short_successor(key):
    """leveldb BytewiseComparator::FindShortSuccessor."""
    for i, b in enumerate(key):
        if b != 0xFF:
            return key[:i] + bytes([b + 1])
    return key


def _read_block(buf, off, size):
    blk = buf[off:off + size]
    nrest = struct.unpack_from("<I", blk, len(blk) - 4)[0]
    end = len(blk) - 4 - 4 * nrest
    pos, key = 0, b""
    while pos < end:
        sh, pos = _get_varint(blk, pos); ns, pos = _get_varint(blk, pos); vl, pos = _get_varint(blk, pos)
        key = key[:sh] + blk[pos:pos + ns]; pos += ns
        yield key, blk[pos:pos + vl]
        pos += vl


# ---- public API ------------------------------------------------------------------------------------
def latest_checkpoint(ckpt_dir):
    """tf.train.latest_checkpoint: follow the text file 'checkpoint'.  None when absent."""
    state = os.path.join(ckpt_dir, "checkpoint")
    if not os.path.isfile(state):
        return None
    for line in open(state):
        if line.startswith("model_checkpoint_path:"):
            name = line.split('"')[1]
            prefix = name if os.path.isabs(name) else os.path.join(ckpt_dir, name)
            return prefix if os.path.isfile(prefix + ".index") else None
    return None


def read_bundle(prefix, verify_crc=True):
    """-> {key: float64 ndarray} for numeric tensors (string tensors are skipped)."""
    idx = open(prefix + ".index", "rb").read()
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    if idx[-8:] != MAGIC:
        raise ValueError("%s.index: bad SSTable magic" % prefix)
    footer = idx[-48:]
    _, p = _get_varint(footer, 0); _, p = _get_varint(footer, p)
    ioff, p = _get_varint(footer, p); isz, _ = _get_varint(footer, p)
    out = {}
    for _, handle in _read_block(idx, ioff, isz):
        boff, p = _get_varint(handle, 0); bsz, _ = _get_varint(handle, p)
        if verify_crc:
            want = struct.unpack_from("<I", idx, boff + bsz + 1)[0]
            if mask_crc(crc32c(idx[boff:boff + bsz + 1])) != want:
                raise ValueError("%s.index: block checksum mismatch" % prefix)
        for key, val in _read_block(idx, boff, bsz):
            if key == b"":
                continue
            dtype, shape, offset, size, crc = 0, [], 0, 0, None
            for f, _, v in _fields(val):
                if f == 1: dtype = v
                elif f == 2:
                    for f2, _, v2 in _fields(v):
                        if f2 == 2:
                            dim = 0
                            for f3, _, v3 in _fields(v2):
                                if f3 == 1: dim = v3
                            shape.append(dim)
                elif f == 4: offset = v
                elif f == 5: size = v
                elif f == 6: crc = v
            if dtype not in (DT_FLOAT, DT_DOUBLE):
                continue
            raw = data[offset:offset + size]
            if verify_crc and crc is not None and mask_crc(crc32c(raw)) != crc:
                raise ValueError("%s: tensor %s checksum mismatch" % (prefix, key.decode()))
            arr = np.frombuffer(raw, "<f8" if dtype == DT_DOUBLE else "<f4").reshape(shape)
            out[key.decode()] = arr.astype(np.float64)
    return out


def _var_key(layer, kind):
    return "layer_with_weights-%d/%s/.ATTRIBUTES/VARIABLE_VALUE" % (layer, kind)


def load_weights(prefix):
    """-> [(kernel[K,f_in,f_out], bias[f_out]), ...] in layer order."""
    t = read_bundle(prefix)
    ws, li = [], 0
    while _var_key(li, "kernel") in t:
        ws.append((t[_var_key(li, "kernel")], t[_var_key(li, "bias")]))
        li += 1
    if not ws:
        raise ValueError("%s: no layer_with_weights-*/kernel entries" % prefix)
    return ws


def save_weights(prefix, weights, update_state=True):
    """Write <prefix>.index / .data-00000-of-00001 (fp64, like the reference) and the 'checkpoint' file."""
    blob = open(_ASSET, "rb").read()
    data, entries = bytearray(), []
    for li, (W, b) in enumerate(weights):
        for kind, arr in (("kernel", W), ("bias", b)):
            arr = np.ascontiguousarray(np.asarray(arr, dtype="<f8"))
            raw = arr.tobytes()
            entries.append((_var_key(li, kind).encode(),
                            _entry_proto(DT_DOUBLE, arr.shape, len(data), len(raw), mask_crc(crc32c(raw)))))
            data += raw
    entries.append((OBJECT_GRAPH_KEY.encode(), _entry_proto(DT_STRING, (), len(data), len(blob), OBJECT_GRAPH_CRC)))
    data += blob
    header = b"\x08\x01\x1a\x02\x08\x01"  # BundleHeaderProto{num_shards=1, version{producer=1}}
    entries = [(b"", header)] + sorted(entries)
    dblock = _block(entries)
    out = bytearray(_with_trailer(dblock))
    meta_off = len(out)
    mblock = _block([])
    out += _with_trailer(mblock)
    idx_off = len(out)
    handle = _put_varint(0) + _put_varint(len(dblock))
    iblock = _block([(_short_successor(entries[-1][0]), handle)])
    out += _with_trailer(iblock)
    footer = _put_varint(meta_off) + _put_varint(len(mblock)) + _put_varint(idx_off) + _put_varint(len(iblock))
    out += footer + b"\x00" * (40 - len(footer)) + MAGIC
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
    if update_state:
        name = os.path.basename(prefix)
        with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), "checkpoint"), "w") as f:
            f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (name, name))
    return prefix
