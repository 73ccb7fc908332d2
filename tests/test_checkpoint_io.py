"""CPU tests of the TensorFlow-V2 checkpoint reader / writer (gifgan/checkpoint_io.py).  No TensorFlow-written file
exists to test against (UNPINNED, see the module header): crc32c and the shape message are pinned against TensorBoard's
independent implementations, the table and bundle layers by round trips and corruption checks, the store mapping on a
small DCGAN / VID_DCGAN."""
import os
import struct

import numpy as np
import pytest
import torch


def test_crc32c_and_mask_against_tensorboard():
    from gifgan import checkpoint_io as C
    from tensorboard.compat.tensorflow_stub.pywrap_tensorflow import crc32c as tb_crc, masked_crc32c as tb_masked
    assert C.crc32c(b"123456789") == 0xE3069283                                     # the standard CRC-32C check value
    rs = np.random.RandomState(0)
    for n in (0, 1, 7, 64, 1000, 4096, 10001):                 # >= 4096: the two-bytes-per-step path
        data = rs.randint(0, 256, n).astype(np.uint8).tobytes()
        assert C.crc32c(data) == tb_crc(data) and C.mask_crc(C.crc32c(data)) == tb_masked(data), n


def test_shape_message_against_tensorboard_protobuf():
    from gifgan import checkpoint_io as C
    from tensorboard.compat.proto import tensor_shape_pb2, types_pb2
    for shape in [(), (7,), (5, 5, 3, 64), (8192, 1), (300000, 2)]:
        msg = tensor_shape_pb2.TensorShapeProto()
        for d in shape:
            msg.dim.add().size = d
        assert C._shape_proto(shape) == msg.SerializeToString() and C._parse_shape(msg.SerializeToString()) == shape
    assert C.DT[types_pb2.DT_FLOAT] is np.float32 and C.DT[types_pb2.DT_INT32] is np.int32 and C.DT[types_pb2.DT_INT64] is np.int64
    assert C.DT[types_pb2.DT_DOUBLE] is np.float64


def test_table_round_trip_and_corruption(tmp_path):
    from gifgan import checkpoint_io as C
    rs = np.random.RandomState(1)
    keys = sorted({("scope_%d/layer_%d/w" % (rs.randint(5), rs.randint(400))).encode() for _ in range(600)})
    entries = [(b"", b"header")] + [(k, rs.randint(0, 256, rs.randint(1, 90)).astype(np.uint8).tobytes()) for k in keys]
    path = str(tmp_path / "t.index")
    C.write_table(path, entries, block_bytes=512)
    assert C.read_table(path) == entries
    raw = bytearray(open(path, "rb").read())
    assert struct.unpack("<Q", raw[-8:])[0] == 0xdb4775248b80fb57 and len(raw) > 20 * 512           # many blocks
    raw[100] ^= 0x40
    open(path, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        C.read_table(path)
    assert len(C.read_table(path, verify=False)) == len(entries)                     # same structure, damaged payload
    raw[-1] ^= 1
    open(path, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="magic"):
        C.read_table(path)


def test_bundle_round_trip(tmp_path):
    from gifgan import checkpoint_io as C
    rs = np.random.RandomState(2)
    tensors = {"d_h0_conv/w": rs.randn(5, 5, 3, 64).astype(np.float32), "d_h0_conv/biases": np.zeros(64, np.float32),
               "beta1_power": np.float32(0.25), "global_step": np.int64(1502), "counts": np.arange(6, dtype=np.int32).reshape(2, 3),
               "g_h0_lin/Matrix": rs.randn(100, 512).astype(np.float32), "d_h0_conv/w/Adam": rs.randn(5, 5, 3, 64).astype(np.float32)}
    prefix = str(tmp_path / "ck" / "DCGAN.model-1502")
    C.write_tf_bundle(prefix, tensors)
    assert sorted(os.listdir(tmp_path / "ck")) == ["DCGAN.model-1502.data-00000-of-00001", "DCGAN.model-1502.index"]
    got = C.read_tf_bundle(prefix, verify=True)
    assert sorted(got) == sorted(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == np.asarray(v).dtype and got[k].shape == np.asarray(v).shape and np.array_equal(got[k], v), k
    # the header entry: num_shards = 1, version.producer = 1; entries sorted by name as the bundle writer does
    table = C.read_table(prefix + ".index")
    assert table[0] == (b"", b"\x08\x01\x1a\x02\x08\x01") and [k for k, _ in table] == sorted(k for k, _ in table)
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    assert len(data) == sum(np.asarray(v).nbytes for v in tensors.values())
    data[10] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ValueError, match="data checksum"):
        C.read_tf_bundle(prefix, verify=True)


def test_dcgan_weights_and_adam_state_through_a_tf_checkpoint(tmp_path):
    from gifgan import checkpoint_io as C, ops
    from gifgan.model import DCGAN
    ops.set_precision("fp32")
    ops.reset_default_store(device="cpu", seed=1)
    m = DCGAN(None, batch_size=4, output_size=16, gf_dim=8, df_dim=8, c_dim=3)
    m.store.flat["m"].uniform_(-1, 1); m.store.flat["v"].uniform_(0, 1)
    m.d_optim.t, m.g_optim.t = 7, 14
    path = C.save_tf_checkpoint(str(tmp_path), "DCGAN.model-7", m.store, (m.d_optim, m.g_optim))
    assert C.latest_checkpoint(str(tmp_path)) == path
    named = C.read_tf_bundle(path)
    assert named["d_h1_conv/w"].shape == (5, 5, 8, 16) and "g_h1/w/Adam_1" in named and abs(named["beta1_power_1"] - 0.5 ** 15) < 1e-9
    ops.reset_default_store(device="cpu", seed=99)
    m2 = DCGAN(None, batch_size=4, output_size=16, gf_dim=8, df_dim=8, c_dim=3)
    assert not torch.equal(m2.store.vars["g_h1/w"].data, m.store.vars["g_h1/w"].data)
    v0 = m2.store.vars["g_h1/w"].version
    assert C.load_tf_checkpoint(str(tmp_path), m2.store, (m2.d_optim, m2.g_optim), verify=True) == []
    for k, v in m.store.vars.items():
        assert torch.equal(m2.store.vars[k].data, v.data), k
    for v in m.d_vars + m.g_vars:                                                       # (the flat buffers also hold alignment padding)
        a, b = v.offset, v.offset + v.numel()
        assert torch.equal(m2.store.flat["m"][a:b], m.store.flat["m"][a:b]) and torch.equal(m2.store.flat["v"][a:b], m.store.flat["v"][a:b]), v.name
    assert (m2.d_optim.t, m2.g_optim.t) == (7, 14) and int(m2.g_optim.state[0]) == 14
    assert m2.store.vars["g_h1/w"].version > v0                                       # stale bf16 copies get rebuilt
    # npz form, ':0' suffixes, strictness
    C.save_npz(str(tmp_path / "w.npz"), m.store)
    ops.reset_default_store(device="cpu", seed=5)
    m3 = DCGAN(None, batch_size=4, output_size=16, gf_dim=8, df_dim=8, c_dim=3)
    C.load_npz(str(tmp_path / "w.npz"), m3.store)
    assert torch.equal(m3.store.vars["d_h3_lin/Matrix"].data, m.store.vars["d_h3_lin/Matrix"].data)
    partial = {k + ":0": v for k, v in C.export_named(m.store).items() if not k.startswith("g_bn0")}
    with pytest.raises(KeyError, match="g_bn0"):
        C.import_named(m3.store, partial)
    assert sorted(C.import_named(m3.store, partial, strict=False)) == ["g_bn0/beta", "g_bn0/gamma", "g_bn0/moving_mean", "g_bn0/moving_variance"]
    with pytest.raises(ValueError, match="shape"):
        C.import_named(m3.store, {"d_h0_conv/w": np.zeros((5, 5, 3, 9), np.float32)}, strict=False)


def test_model_load_paths_accept_tf_checkpoints(tmp_path):
    """DCGAN.load (model.py:441-452) and VID_DCGAN.load_image_gan (z_model_lib.py:117-134) pick the TensorFlow bundle
    when the newest checkpoint of the directory is one, the torch payload otherwise."""
    from gifgan import checkpoint_io as C, ops
    from gifgan.model import DCGAN
    from gifgan.z_model_lib import VID_DCGAN
    ops.set_precision("fp32")
    ops.reset_default_store(device="cpu", seed=3)
    a = DCGAN(None, batch_size=4, output_size=64, c_dim=3, dataset_name="faces")
    ck = tmp_path / "faces_4_64"                                   # the directory DCGAN.save / load derive (model.py:430-431)
    C.save_tf_checkpoint(str(ck), "DCGAN.model-3", a.store)
    # a state file written elsewhere holds an absolute path: only its basename counts
    open(ck / "checkpoint", "w").write('model_checkpoint_path: "/somewhere/else/DCGAN.model-3"\n')
    ops.reset_default_store(device="cpu", seed=4)
    b = DCGAN(None, batch_size=4, output_size=64, c_dim=3, dataset_name="faces")
    assert b.load(str(tmp_path))
    assert torch.equal(b.store.vars["g_h2/w"].data, a.store.vars["g_h2/w"].data)
    ops.reset_default_store(device="cpu", seed=5)
    with ops.variable_scope("video_gan"):
        vid = VID_DCGAN(None, batch_size=2, z_input_size=120, z_output_size=100, vid_length=16, input_image_size=64,
                        output_image_size=64, c_dim=3, sample_cols=2)
    assert vid.load_image_gan(None, str(ck))
    for k, v in a.store.vars.items():                              # every image-GAN variable, under the scope prefix
        assert torch.equal(vid.store.vars[vid.image_gan_scope_name + k].data, v.data), k
    # torch payloads still load
    a.save(str(tmp_path / "torch"), 9)
    assert b.load(str(tmp_path / "torch"))


def test_snappy_blocks_in_a_table(tmp_path):
    """Tables written with the LevelDB default compress their blocks (tag 1); the raw snappy format: literals, copies
    with 1- / 2-byte offsets, overlapping copies (run-length), long literals."""
    from gifgan import checkpoint_io as C
    # hand-assembled streams: "abcabcabcabc" = literal "abc" + copy(offset 3, length 9, 2-byte-offset form)
    assert C.snappy_decompress(bytes([12, (3 - 1) << 2]) + b"abc" + bytes([((9 - 1) << 2) | 2, 3, 0])) == b"abcabcabcabc"
    # 1-byte-offset copy: length 4..11 in bits 2-4, offset high bits in 5-7:  "xyzw" + copy(offset 4, length 7)
    assert C.snappy_decompress(bytes([11, (4 - 1) << 2]) + b"xyzw" + bytes([((7 - 4) << 2) | 1 | (0 << 5), 4])) == b"xyzwxyzwxyz"
    # literal of 70 bytes: length-1 = 69 >= 60 -> tag 60<<2, one extra length byte
    lit = bytes(range(70))
    assert C.snappy_decompress(bytes([70, 60 << 2, 69]) + lit) == lit
    with pytest.raises(ValueError):
        C.snappy_decompress(bytes([5, (3 - 1) << 2]) + b"abc")                     # declared 5 bytes, holds 3

    def literal_only(raw):              # a valid (if pointless) snappy stream: everything as literals of <= 60 bytes
        out = bytearray(C._varint(len(raw)))
        for i in range(0, len(raw), 60):
            chunk = raw[i:i + 60]
            out += bytes([(len(chunk) - 1) << 2]) + chunk
        return bytes(out)

    entries = [(b"", b"hdr")] + [(("var_%03d/w" % i).encode(), bytes([i]) * (i % 40 + 1)) for i in range(80)]
    body = literal_only(C._block(entries))
    index_body = C._block([(entries[-1][0], C._handle(0, len(body)))])
    meta = C._block([])
    out = bytearray()
    for blk, tag in ((body, 1), (meta, 0), (index_body, 0)):
        out += blk + bytes([tag]) + struct.pack("<I", C.mask_crc(C.crc32c(blk + bytes([tag]))))
    moff, ioff = len(body) + 5, len(body) + 5 + len(meta) + 5
    foot = C._handle(moff, len(meta)) + C._handle(ioff, len(index_body))
    out += foot + b"\x00" * (40 - len(foot)) + struct.pack("<Q", C.MAGIC)
    path = str(tmp_path / "snappy.index")
    open(path, "wb").write(bytes(out))
    assert C.read_table(path) == entries


def test_v1_checkpoint_reader(tmp_path):
    """tf.train.Saver(write_version=1) files (what the reference's utils/downgrade_tf_checkpoint.py writes): one table;
    key "" holds the SavedTensorSliceMeta (name, shape, dtype per variable), every other entry a SavedSlice whose
    TensorProto carries the values.  The TensorProto side is built with TensorBoard's protobuf classes (independent of the
    parser); the SavedTensorSlices wrappers are assembled by hand from the published field numbers."""
    from gifgan import checkpoint_io as C, ops
    from gifgan.model import DCGAN
    from tensorboard.compat.proto import tensor_pb2, tensor_shape_pb2, types_pb2

    def ld(field, payload):                                     # length-delimited field
        return bytes([(field << 3) | 2]) + C._varint(len(payload)) + payload

    rs = np.random.RandomState(4)
    tensors = {"d_h0_conv/w": rs.randn(5, 5, 3, 8).astype(np.float32), "d_h0_conv/biases": rs.randn(8).astype(np.float32),
               "beta1_power": np.float32(0.125), "global_step": np.int64(77), "flags": np.array([3, -4], np.int32)}
    enum = {np.dtype(np.float32): types_pb2.DT_FLOAT, np.dtype(np.int64): types_pb2.DT_INT64, np.dtype(np.int32): types_pb2.DT_INT32}
    metas, entries = b"", []
    for name in sorted(tensors):
        a = np.asarray(tensors[name])
        shape = tensor_shape_pb2.TensorShapeProto()
        for d in a.shape:
            shape.dim.add().size = d
        full_slice = b"".join(ld(1, b"") for _ in a.shape)       # TensorSliceProto: one empty Extent per dimension = everything
        metas += ld(1, ld(1, name.encode()) + ld(2, shape.SerializeToString()) + bytes([3 << 3, enum[a.dtype]]) + ld(4, full_slice))
        tp = tensor_pb2.TensorProto(dtype=enum[a.dtype])
        if a.dtype == np.float32:
            tp.float_val.extend(a.reshape(-1).tolist())
        elif a.dtype == np.int64:
            tp.int64_val.extend(a.reshape(-1).tolist())
        else:
            tp.int_val.extend(a.reshape(-1).tolist())
        saved_slice = ld(1, name.encode()) + ld(2, full_slice) + ld(3, tp.SerializeToString())
        entries.append((b"\x00" + name.encode(), ld(2, saved_slice)))      # (real keys are an ordered encoding of name + slice)
    table = [(b"", ld(1, metas))] + entries
    path = str(tmp_path / "DCGAN.model-77")
    C.write_table(path, table, block_bytes=1024)
    got = C.read_tf_v1_checkpoint(path)
    assert sorted(got) == sorted(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == np.asarray(v).dtype and got[k].shape == np.asarray(v).shape and np.array_equal(got[k], v), k
    # load_tf_checkpoint picks the format from what is on disk
    open(tmp_path / "checkpoint", "w").write('model_checkpoint_path: "DCGAN.model-77"\n')
    ops.set_precision("fp32")
    ops.reset_default_store(device="cpu", seed=2)
    m = DCGAN(None, batch_size=2, output_size=16, gf_dim=8, df_dim=8, c_dim=3)
    missing = C.load_tf_checkpoint(str(tmp_path), m.store, (m.d_optim, m.g_optim), strict=False)
    assert "d_h0_conv/w" not in missing and "g_h1/w" in missing
    assert np.array_equal(m.store.vars["d_h0_conv/w"].data.numpy(), tensors["d_h0_conv/w"]) and m.d_optim.t == 2   # 0.5^(t+1) = 0.125
    with pytest.raises(IOError):
        C.load_tf_checkpoint(str(tmp_path / "nothing"), m.store)
    assert C.tf_format(path) == "v1" and C.tf_format(str(tmp_path / "nothing")) is None
    m.save(str(tmp_path / "torch"), 5)
    assert C.tf_format(os.path.join(str(tmp_path / "torch"), "default_2_16", "DCGAN.model-5")) is None      # a torch payload
