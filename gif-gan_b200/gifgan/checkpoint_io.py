"""Checkpoint interchange: reading and writing TensorFlow "V2" checkpoints (the TensorBundle format
`tf.train.Saver` writes since 0.12: `<prefix>.index` + `<prefix>.data-00000-of-00001`, next to a `checkpoint` state
file) without TensorFlow, so that weights trained with the reference (`DCGAN.save`, models/recurrent_z/model.py:428-439;
`VID_DCGAN` saver, z_model_lib.py:204,256-259) load into the variable store here under the same names, and weights
trained here can be handed back.  Also a plain `.npz` form keyed by the same names.

UNPINNED: no TensorFlow checkpoint ships with the reference and TensorFlow cannot run in this image, so the reader is
checked against the writer in this file, against TensorBoard's independent crc32c and protobuf encoders, and against the
published format -- not against a file TensorFlow wrote:

  * `.index` is a LevelDB-format table (tensorflow/core/lib/io/table*.cc): data blocks of prefix-compressed entries
    (varint32 shared, non_shared, value_len; key suffix; value) with a restart array, each block followed by a 1-byte
    compression tag (0 = none, what the bundle writer emits; 1 = snappy, decoded here too) and a masked crc32c; a meta-index block; an index
    block (separator key -> BlockHandle{offset, size} as varint64s); a 48-byte footer (two handles padded to 40 bytes +
    magic 0xdb4775248b80fb57).
  * key "" -> BundleHeaderProto{num_shards=1, endianness=2, version=3}; every other key is a variable name ->
    BundleEntryProto{dtype=1, shape=2 (TensorShapeProto), shard_id=3, offset=4, size=5, crc32c=6 (fixed32, masked)}.
  * `.data-SSSSS-of-NNNNN`: the tensors' little-endian row-major bytes at [offset, offset + size).

Adam state, WHEN a checkpoint holds it: slots `<var>/Adam` (m), `<var>/Adam_1` (v) and the scalars `beta1_power`,
`beta2_power` (= beta^(t+1) after t updates), `<scope>/beta1_power` for the second optimiser.  The reference's image DCGAN
builds its `tf.train.Saver()` BEFORE the optimisers exist (model.py:141 vs :146-149), so its checkpoints hold the model
variables only; the video GAN's Saver (z_model_lib.py:204) is built after them and holds the slots.  Both load here: slots
are taken when both of a variable's are present and left at zero otherwise.
"""
from __future__ import annotations

import os
import struct

import numpy as np

MAGIC = 0xdb4775248b80fb57
DT = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 4: np.uint8, 6: np.int8, 5: np.int16, 10: np.bool_}
DT_OF = {np.dtype(v): k for k, v in DT.items()}

# ---- crc32c (Castagnoli), masked as in tensorflow/core/lib/hash/crc32c.h -----------------------------------------
_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _TABLE.append(_c)


_TABLE16 = None


def crc32c(data: bytes, crc: int = 0) -> int:
    """Two bytes per table step for the bulk (tensor data runs to tens of MB), one byte per step for the tail."""
    global _TABLE16
    t, c = _TABLE, crc ^ 0xFFFFFFFF
    n2 = len(data) & ~1
    if n2 >= 4096:
        import sys
        if _TABLE16 is None:
            _TABLE16 = [t[(t[x & 0xFF] ^ (x >> 8)) & 0xFF] ^ (t[x & 0xFF] >> 8) for x in range(65536)]
        t16 = _TABLE16
        words = memoryview(bytes(data[:n2])).cast("H")
        if sys.byteorder != "little":
            words = [((w & 0xFF) << 8) | (w >> 8) for w in words]
        for w in words:
            c = t16[(c ^ w) & 0xFFFF] ^ (c >> 16)
        data = data[n2:]
    for b in data:
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(c: int) -> int:
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- protobuf wire helpers ---------------------------------------------------------------------------
def _varint(n: int) -> bytes:
    out = bytearray()
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, pos):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _fields(buf):
    """Yields (field number, wire type, value) of one serialized message; value is int or bytes."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v, pos = struct.unpack_from("<Q", buf, pos)[0], pos + 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            v, pos = bytes(buf[pos:pos + ln]), pos + ln
        elif wt == 5:
            v, pos = struct.unpack_from("<I", buf, pos)[0], pos + 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield fno, wt, v


def _shape_proto(shape) -> bytes:
    out = b""
    for d in shape:
        dim = b"\x08" + _varint(int(d))
        out += b"\x12" + _varint(len(dim)) + dim
    return out


def _parse_shape(buf):
    dims = []
    for fno, _, v in _fields(buf):
        if fno == 2:
            size = 0
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    size = v2
            dims.append(size)
    return tuple(dims)


def _entry_proto(dtype, shape, offset, size, crc) -> bytes:
    sp = _shape_proto(shape)
    return (b"\x08" + _varint(dtype) + b"\x12" + _varint(len(sp)) + sp + b"\x20" + _varint(offset) + b"\x28" + _varint(size)
            + b"\x35" + struct.pack("<I", crc))                      # shard_id 0 is the proto default: omitted


# ---- LevelDB-format table ----------------------------------------------------------------------------
def _block(entries, restart_interval=16) -> bytes:
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            m = min(len(k), len(last))
            while shared < m and k[shared] == last[shared]:
                shared += 1
        out += _varint(shared) + _varint(len(k) - shared) + _varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _parse_block(buf):
    n_restarts = struct.unpack_from("<I", buf, len(buf) - 4)[0]
    end = len(buf) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _read_varint(buf, pos)
        non_shared, pos = _read_varint(buf, pos)
        vlen, pos = _read_varint(buf, pos)
        key = key[:shared] + bytes(buf[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(buf[pos:pos + vlen])))
        pos += vlen
    return out


def snappy_decompress(buf: bytes) -> bytes:
    """Raw snappy block format (the table format's compression tag 1): varint length, then literals (tag & 3 == 0) and
    copies with 1-, 2- or 4-byte offsets (tags 1, 2, 3).  The bundle writer does not compress; tables written with the
    LevelDB default do."""
    n, pos = _read_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                extra = ln - 59
                ln = int.from_bytes(buf[pos:pos + extra], "little")
                pos += extra
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln, off = ((tag >> 2) & 7) + 4, ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln, off = (tag >> 2) + 1, int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln, off = (tag >> 2) + 1, int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy block")
        for _ in range(ln):                  # byte-wise: a copy may overlap its own output (run-length encoding)
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy block decompressed to %d bytes, header says %d" % (len(out), n))
    return bytes(out)


def _handle(offset, size) -> bytes:
    return _varint(offset) + _varint(size)


def _read_block(buf, offset, size, verify):
    body, tag = buf[offset:offset + size], buf[offset + size]
    if verify:
        want = struct.unpack_from("<I", buf, offset + size + 1)[0]
        if mask_crc(crc32c(bytes(buf[offset:offset + size + 1]))) != want:
            raise ValueError("table block checksum mismatch at offset %d" % offset)
    if tag == 1:
        body = snappy_decompress(bytes(body))
    elif tag != 0:
        raise ValueError("table block with unknown compression tag %d" % tag)
    return _parse_block(body)


def read_table(path, verify=True):
    """-> list of (key bytes, value bytes) of a LevelDB-format table file, in key order."""
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != MAGIC:
        raise ValueError("%s is not a table file (bad magic)" % path)
    foot = buf[len(buf) - 48:]
    _, p = _read_varint(foot, 0)
    _, p = _read_varint(foot, p)
    ioff, p = _read_varint(foot, p)
    isz, p = _read_varint(foot, p)
    out = []
    for _, hv in _read_block(buf, ioff, isz, verify):
        off, q = _read_varint(hv, 0)
        sz, _ = _read_varint(hv, q)
        out.extend(_read_block(buf, off, sz, verify))
    return out


def write_table(path, entries, block_bytes=4096):
    """entries: (key bytes, value bytes) sorted by key."""
    out, index, cur, cur_size = bytearray(), [], [], 0

    def emit(block_entries):
        body = _block(block_entries)
        off = len(out)
        out.extend(body + b"\x00" + struct.pack("<I", mask_crc(crc32c(body + b"\x00"))))
        return off, len(body)

    def flush():
        nonlocal cur, cur_size
        if cur:
            off, sz = emit(cur)
            index.append((cur[-1][0], _handle(off, sz)))        # the block's last key is a valid separator
            cur, cur_size = [], 0

    for k, v in entries:
        cur.append((k, v))
        cur_size += len(k) + len(v) + 3
        if cur_size >= block_bytes:
            flush()
    flush()
    moff, msz = emit([])                                         # empty meta-index block
    ioff, isz = emit(index)
    foot = _handle(moff, msz) + _handle(ioff, isz)
    out.extend(foot + b"\x00" * (40 - len(foot)) + struct.pack("<Q", MAGIC))
    with open(path, "wb") as f:
        f.write(bytes(out))


# ---- tensor bundles -------------------------------------------------------------------------------------
def read_tf_bundle(prefix, verify=False):
    """`prefix` as TensorFlow names it (e.g. .../DCGAN.model-1502) -> {variable name: numpy array}."""
    entries = read_table(prefix + ".index", verify=True)
    if not entries or entries[0][0] != b"":
        raise ValueError("%s.index has no bundle header" % prefix)
    num_shards = 1
    for fno, _, v in _fields(entries[0][1]):
        if fno == 1:
            num_shards = v
        if fno == 2 and v != 0:
            raise ValueError("big-endian bundle")
    shards, out = {}, {}
    for key, val in entries[1:]:
        dtype, shape, shard, offset, size, crc, sliced = 0, (), 0, 0, 0, None, False
        for fno, _, v in _fields(val):
            if fno == 1:
                dtype = v
            elif fno == 2:
                shape = _parse_shape(v)
            elif fno == 3:
                shard = v
            elif fno == 4:
                offset = v
            elif fno == 5:
                size = v
            elif fno == 6:
                crc = v
            elif fno == 7:
                sliced = True
        if sliced:
            raise ValueError("variable %s is stored as slices (partitioned variable): not supported" % key.decode())
        if dtype not in DT:
            raise ValueError("variable %s: unsupported dtype enum %d" % (key.decode(), dtype))
        if shard not in shards:
            shards[shard] = np.memmap("%s.data-%05d-of-%05d" % (prefix, shard, num_shards), dtype=np.uint8, mode="r")
        raw = bytes(shards[shard][offset:offset + size])
        if verify and crc is not None and mask_crc(crc32c(raw)) != crc:
            raise ValueError("variable %s: data checksum mismatch" % key.decode())
        arr = np.frombuffer(raw, dtype=np.dtype(DT[dtype]).newbyteorder("<")).reshape(shape)
        out[key.decode()] = arr.astype(DT[dtype])
    return out


def write_tf_bundle(prefix, tensors):
    """{name: array} -> `<prefix>.index` + `<prefix>.data-00000-of-00001` (one shard, keys in byte order)."""
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"                  # num_shards = 1, (endianness = little: default), version{producer = 1}
    entries, offset = [(b"", header)], 0
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + ".data-00000-of-00001", "wb") as data:
        for name in sorted(tensors, key=lambda s: s.encode()):
            a = np.asarray(tensors[name], order="C")
            if a.dtype not in DT_OF:
                raise ValueError("%s: dtype %s has no TensorFlow enum here" % (name, a.dtype))
            raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
            data.write(raw)
            entries.append((name.encode(), _entry_proto(DT_OF[a.dtype], a.shape, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    write_table(prefix + ".index", entries)


# ---- V1 checkpoints (tf.train.Saver(write_version=1): what utils/downgrade_tf_checkpoint.py of the reference produces) ----
# One table file (blocks usually snappy-compressed).  Key "" -> SavedTensorSlices{meta = 1: SavedTensorSliceMeta{tensor = 1:
# SavedSliceMeta{name = 1, shape = 2, type = 3, slice = 4}}}; every other key (an order-preserving encoding of name + slice
# that is not needed to read the file) -> SavedTensorSlices{data = 2: SavedSlice{name = 1, slice = 2, data = 3: TensorProto}}
# with the values in the TensorProto's typed repeated field (float_val = 5, double_val = 6, int_val = 7, int64_val = 10,
# bool_val = 11; packed) or in tensor_content = 4.
_V1_VAL_FIELD = {1: (5, "<f4"), 2: (6, "<f8"), 3: (7, None), 9: (10, None), 10: (11, None)}


def _tensor_proto_values(buf, dtype):
    field, fixed = _V1_VAL_FIELD[dtype]
    vals, content = [], None
    for fno, wt, v in _fields(buf):
        if fno == 4 and wt == 2:
            content = v
        elif fno == field:
            if wt == 2:                                             # packed
                if fixed:
                    vals.append(np.frombuffer(v, dtype=fixed))
                else:
                    pos, out = 0, []
                    while pos < len(v):
                        x, pos = _read_varint(v, pos)
                        out.append(x - (1 << 64) if x >> 63 else x)
                    vals.append(np.array(out, dtype=np.int64))
            elif wt == 5:
                vals.append(np.frombuffer(struct.pack("<I", v), dtype="<f4"))
            elif wt == 1:
                vals.append(np.frombuffer(struct.pack("<Q", v), dtype="<f8"))
            else:
                vals.append(np.array([v - (1 << 64) if v >> 63 else v], dtype=np.int64))
    if content is not None:
        return np.frombuffer(content, dtype=np.dtype(DT[dtype]).newbyteorder("<"))
    return np.concatenate(vals) if vals else np.zeros(0, dtype=DT[dtype])


def read_tf_v1_checkpoint(path):
    """A V1 checkpoint file (e.g. .../DCGAN.model-1502 itself, no .index) -> {variable name: numpy array}.
    Variables saved as several slices (partitioned variables) are not supported."""
    entries = read_table(path, verify=True)
    if not entries or entries[0][0] != b"":
        raise ValueError("%s: no SavedTensorSliceMeta entry" % path)
    shapes = {}
    for fno, _, v in _fields(entries[0][1]):
        if fno != 1:
            continue
        for f2, _, t in _fields(v):
            if f2 != 1:
                continue
            name, shape, dtype, nslices = None, (), 0, 0
            for f3, _, x in _fields(t):
                if f3 == 1:
                    name = x.decode()
                elif f3 == 2:
                    shape = _parse_shape(x)
                elif f3 == 3:
                    dtype = x
                elif f3 == 4:
                    nslices += 1
            if nslices > 1:
                raise ValueError("variable %s is stored as %d slices: not supported" % (name, nslices))
            shapes[name] = (shape, dtype)
    out = {}
    for key, val in entries[1:]:
        for fno, _, v in _fields(val):
            if fno != 2:
                continue
            name, data = None, b""
            for f2, _, x in _fields(v):
                if f2 == 1:
                    name = x.decode()
                elif f2 == 3:
                    data = x
            shape, dtype = shapes[name]
            if dtype not in _V1_VAL_FIELD:
                raise ValueError("variable %s: unsupported dtype enum %d" % (name, dtype))
            vals = _tensor_proto_values(data, dtype)
            n = int(np.prod(shape)) if shape else 1
            if vals.size != n:
                raise ValueError("variable %s: %d values for shape %s" % (name, vals.size, shape))
            out[name] = vals.astype(DT[dtype]).reshape(shape)
    return out


def tf_format(prefix):
    """'v2' when `<prefix>.index` exists, 'v1' when `prefix` itself is a table file, None otherwise (e.g. a torch payload)."""
    if os.path.exists(prefix + ".index"):
        return "v2"
    if os.path.isfile(prefix) and os.path.getsize(prefix) >= 48:
        with open(prefix, "rb") as f:
            f.seek(-8, os.SEEK_END)
            if struct.unpack("<Q", f.read(8))[0] == MAGIC:
                return "v1"
    return None


def latest_checkpoint(directory):
    """tf.train.get_checkpoint_state(directory).model_checkpoint_path, re-rooted at `directory` (the stored path may be
    absolute on the machine that wrote it; z_space_finder.py:63-67 does the same)."""
    index = os.path.join(directory, "checkpoint")
    if not os.path.exists(index):
        return None
    for line in open(index):
        if line.startswith("model_checkpoint_path:"):
            return os.path.join(directory, os.path.basename(line.split('"')[1]))
    return None


# ---- variable store <-> named arrays ---------------------------------------------------------------------
def export_named(store, optimisers=()):
    """Variables (and, for each optimiser, its Adam slots and beta powers) under the reference's TensorFlow names
    (slot and beta-power names as main.py's two top-level optimisers get them: `beta1_power`, `beta1_power_1`)."""
    out = {k: v.data.detach().cpu().numpy() for k, v in store.vars.items()}
    for i, opt in enumerate(optimisers):
        for v in opt.var_list or []:
            n = v.numel()
            out[v.name + "/Adam"] = store.flat["m"][v.offset:v.offset + n].reshape(v.shape).cpu().numpy()
            out[v.name + "/Adam_1"] = store.flat["v"][v.offset:v.offset + n].reshape(v.shape).cpu().numpy()
        sfx = "" if i == 0 else "_%d" % i
        out["beta1_power" + sfx] = np.float32(opt.b1 ** (opt.t + 1))
        out["beta2_power" + sfx] = np.float32(opt.b2 ** (opt.t + 1))
    return out


def import_named(store, named, optimisers=(), prefix="", strict=True):
    """Load {TensorFlow name: array} into the store: variables by name (a trailing ':0' is ignored, `prefix` is put in
    front of the file's names -- load_image_gan's scope stripping in reverse, z_model_lib.py:117-134), Adam slots and
    step counts when the file has them.  Returns the names of the store's variables the file did not contain."""
    import torch
    named = {(k[:-2] if k.endswith(":0") else k): v for k, v in named.items()}
    missing = []
    with torch.no_grad():
        for k, v in store.vars.items():
            if not k.startswith(prefix):
                continue
            key = k[len(prefix):]
            if key in named:
                a = np.asarray(named[key])
                if tuple(a.shape) != tuple(v.shape):
                    raise ValueError("variable %s: checkpoint shape %s, model shape %s" % (key, a.shape, tuple(v.shape)))
                v.data.copy_(torch.as_tensor(a.astype(np.float32)))
                v.version += 1
                n = v.numel()
                if store.flat is not None and key + "/Adam" in named and key + "/Adam_1" in named and v.offset + n <= store.flat["m"].numel():
                    store.flat["m"][v.offset:v.offset + n].copy_(torch.as_tensor(np.asarray(named[key + "/Adam"], dtype=np.float32)).reshape(-1))
                    store.flat["v"][v.offset:v.offset + n].copy_(torch.as_tensor(np.asarray(named[key + "/Adam_1"], dtype=np.float32)).reshape(-1))
            else:
                missing.append(key)
    for i, opt in enumerate(optimisers):
        bp = named.get("beta1_power" + ("" if i == 0 else "_%d" % i))
        if bp is not None and 0.0 < float(bp) < 1.0:
            opt.t = max(0, int(round(np.log(float(bp)) / np.log(opt.b1))) - 1)
            opt.state[0] = opt.t
    if strict and missing:
        raise KeyError("variables missing from the checkpoint: %s%s" % (missing[:5], "..." if len(missing) > 5 else ""))
    return missing


def save_npz(path, store, optimisers=()):
    np.savez(path, **export_named(store, optimisers))


def load_npz(path, store, optimisers=(), prefix="", strict=True):
    with np.load(path) as z:
        return import_named(store, {k: z[k] for k in z.files}, optimisers, prefix, strict)


def save_tf_checkpoint(directory, name, store, optimisers=()):
    """Write `<directory>/<name>.index|.data-00000-of-00001` and the `checkpoint` state file tf.train.Saver keeps."""
    os.makedirs(directory, exist_ok=True)
    write_tf_bundle(os.path.join(directory, name), export_named(store, optimisers))
    with open(os.path.join(directory, "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (name, name))
    return os.path.join(directory, name)


def load_tf_checkpoint(directory_or_prefix, store, optimisers=(), prefix="", strict=True, verify=False):
    """Restore from a TensorFlow V2 checkpoint: a directory holding a `checkpoint` state file, or a bundle prefix."""
    p = directory_or_prefix
    if os.path.isdir(p):
        p = latest_checkpoint(p)
        if p is None:
            raise IOError("no checkpoint state file in %s" % directory_or_prefix)
    if os.path.exists(p + ".index"):
        named = read_tf_bundle(p, verify=verify)
    elif os.path.isfile(p):
        named = read_tf_v1_checkpoint(p)                  # a single table file: the V1 format
    else:
        raise IOError("neither %s.index (V2) nor %s (V1) exists" % (p, p))
    return import_named(store, named, optimisers, prefix, strict)
