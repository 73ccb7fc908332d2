"""CPU tests of the host input pipeline (gifgan/input_pipeline.py; the decode work of model.py:212-219 and
z_model_lib.py:332-351 moved ahead of the step): order and content against the synchronous loaders, bounded look-ahead,
buffer recycling, error delivery, shutdown."""
import os
import threading
import time

import numpy as np
import pytest
import torch


def test_batches_arrive_in_order_with_bounded_lookahead():
    from gifgan.input_pipeline import Prefetcher, chunks
    items = list(range(50))
    started = []
    lock = threading.Lock()

    def load(i):
        with lock:
            started.append(i)
        time.sleep(0.002 * (i % 3))                              # out-of-order completion inside a batch
        return np.full((2, 3), i, dtype=np.float64)

    batches = chunks(items, 4)
    assert len(batches) == 12 and batches[-1] == [44, 45, 46, 47] and len(chunks(items, 4, drop_last=False)) == 13
    with Prefetcher(batches, load, (2, 3), depth=2, workers=3, pin=False) as pf:
        assert len(pf) == 12
        time.sleep(0.3)                                          # the consumer is slow: the producer must stop `depth` ahead
        assert max(started) < 4 * (pf.depth + 1), max(started)   # depth ready + one being filled
        seen, held = [], []
        for k, b in enumerate(pf):
            assert isinstance(b, torch.Tensor) and b.dtype == torch.float32 and tuple(b.shape) == (4, 2, 3)
            assert torch.equal(b, torch.tensor(batches[k], dtype=torch.float32).reshape(4, 1, 1).expand(4, 2, 3))
            held.append((b, b.clone()))
            if len(held) >= 2:                                   # the previous batch is still intact while this one is in use
                assert torch.equal(*held[-2])
            seen.append(k)
        assert seen == list(range(12)) and pf.max_ahead <= pf.depth + 1
    assert not pf._producer.is_alive()


def test_loader_error_is_raised_at_its_batch_and_close_stops_the_workers():
    from gifgan.input_pipeline import Prefetcher

    def load(i):
        if i == 9:
            raise IOError("cannot read item 9")
        return np.zeros(4)

    pf = Prefetcher([[0, 1], [2, 3], [8, 9], [10, 11]], load, (4,), depth=1, workers=2, pin=False)
    assert tuple(next(pf).shape) == (2, 4) and tuple(next(pf).shape) == (2, 4)
    with pytest.raises(IOError, match="item 9"):
        next(pf)
    pf._producer.join(timeout=2.0)
    assert not pf._producer.is_alive()
    # closing early must not hang even though batches remain
    pf2 = Prefetcher([[i] for i in range(100)], lambda i: np.zeros(4), (4,), depth=2, workers=2, pin=False)
    next(pf2)
    pf2.close()
    pf2._producer.join(timeout=2.0)
    assert not pf2._producer.is_alive()


def test_dcgan_file_batches_equal_the_synchronous_loader(tmp_path):
    """DCGAN.file_batches against the loop of model.py:212-219 on generated image files (incl. centre crop + resize)."""
    import cv2
    from gifgan import ops
    from gifgan.model import DCGAN
    from gifgan.utils import get_image
    rs = np.random.RandomState(0)
    files = []
    for i in range(7):
        f = str(tmp_path / ("img%02d.png" % i))
        cv2.imwrite(f, rs.randint(0, 256, (40, 48, 3)).astype(np.uint8))
        files.append(f)
    ops.set_precision("fp32")
    ops.reset_default_store(device="cpu", seed=1)
    m = DCGAN(None, image_size=32, is_crop=True, batch_size=3, output_size=16, gf_dim=8, df_dim=8, c_dim=3)
    got = list(m.file_batches(files, 3, depth=2, workers=2))
    assert len(got) == 2                                          # 7 // 3, the ragged tail is dropped like the reference's loop
    for k, b in enumerate(got):
        want = np.array([get_image(f, 32, is_crop=True, resize_w=16) for f in files[3 * k:3 * k + 3]]).astype(np.float32)
        assert tuple(b.shape) == (3, 16, 16, 3) and np.array_equal(b.numpy(), want)
        assert -1.0 <= float(b.min()) and float(b.max()) <= 1.0


def test_vid_dcgan_video_batches_equal_load_videos(tmp_path):
    """VID_DCGAN.video_batches against get_videos (z_model_lib.py:332-351) on generated clips."""
    import cv2
    from gifgan import ops
    from gifgan.z_model_lib import VID_DCGAN
    files = []
    for i in range(5):
        f = str(tmp_path / ("clip%d.avi" % i))
        wr = cv2.VideoWriter(f, cv2.VideoWriter_fourcc(*"MJPG"), 25.0, (24, 24))
        if not wr.isOpened():
            pytest.skip("no MJPG writer in this OpenCV build")
        for t in range(4):
            wr.write(np.full((24, 24, 3), 40 * i + 10 * t, np.uint8))
        wr.release()
        files.append(f)
    ops.set_precision("fp32")
    ops.reset_default_store(device="cpu", seed=1)
    with ops.variable_scope("video_gan"):
        m = VID_DCGAN(None, batch_size=2, z_input_size=120, z_output_size=100, vid_length=4, input_image_size=64,
                      output_image_size=64, c_dim=3, sample_cols=2)
    got = list(m.video_batches(files, 2, workers=2))
    assert len(got) == 2
    for k, b in enumerate(got):
        want = m.load_videos(files[2 * k:2 * k + 2]).astype(np.float32)
        assert tuple(b.shape) == (2, 4, 64, 64, 3)
        assert np.array_equal(b.reshape(-1, 64, 64, 3).numpy(), want)


def test_vid_dcgan_raw_batches_plus_decode_tail_equal_load_videos(tmp_path):
    """The split loader -- raw uint8 BGR frames from the workers (raw_video_batches), resize + BGR->RGB + /127.5-1 afterwards --
    gives exactly load_videos' batch (z_model_lib.py:332-351).  The tail is evaluated here with the oracle restatement (CPU);
    tests/test_gpu_frames.py checks the CUDA kernel against the same oracle bit for bit."""
    import cv2
    from gifgan import ops
    from gifgan.z_model_lib import VID_DCGAN
    from oracle import image_ops as I
    rs = np.random.RandomState(4)
    files = []
    for i in range(4):
        f = str(tmp_path / ("clip%d.avi" % i))
        wr = cv2.VideoWriter(f, cv2.VideoWriter_fourcc(*"MJPG"), 25.0, (40, 24))       # (width, height): non-square, enlarged AND reduced
        if not wr.isOpened():
            pytest.skip("no MJPG writer in this OpenCV build")
        for t in range(4):
            wr.write(cv2.resize(rs.randint(0, 256, (6, 10, 3)).astype(np.uint8), (40, 24), interpolation=cv2.INTER_CUBIC))
        wr.release()
        files.append(f)
    ops.set_precision("fp32")
    ops.reset_default_store(device="cpu", seed=1)
    with ops.variable_scope("video_gan"):
        m = VID_DCGAN(None, batch_size=2, z_input_size=120, z_output_size=100, vid_length=4, input_image_size=32,
                      output_image_size=32, c_dim=3, sample_cols=2)
    got = list(m.raw_video_batches(files, 2, (24, 40), workers=2))
    assert len(got) == 2
    for k, b in enumerate(got):
        assert b.dtype == torch.uint8 and tuple(b.shape) == (2, 4, 24, 40, 3)
        want = m.load_videos(files[2 * k:2 * k + 2]).astype(np.float32)
        tail = I.frames_to_input(b.reshape(-1, 24, 40, 3).numpy(), 32)
        assert np.array_equal(tail, want)
    with pytest.raises(ValueError):
        m.load_videos_raw(files[:1], (24, 24))                                           # wrong declared resolution


def test_dropping_the_iterator_mid_epoch_stops_the_threads():
    """The producer and the decode workers hold only the shared state, never the Prefetcher: `del` + gc must stop them."""
    import gc
    import threading
    import time
    from gifgan.input_pipeline import Prefetcher
    before = {t.ident for t in threading.enumerate()}
    pf = Prefetcher([[i] * 2 for i in range(50)], lambda i: np.full((4,), float(i), dtype=np.float32), (4,), depth=2, workers=2, pin=False)
    assert float(next(pf)[0, 0]) == 0.0
    prod = pf._producer
    del pf
    gc.collect()
    deadline = time.time() + 5.0
    while prod.is_alive() and time.time() < deadline:
        time.sleep(0.05)
    assert not prod.is_alive()
    time.sleep(0.2)
    leftover = [t.name for t in threading.enumerate() if t.ident not in before and t.name.startswith("gifgan-prefetch")]
    assert not leftover, leftover
