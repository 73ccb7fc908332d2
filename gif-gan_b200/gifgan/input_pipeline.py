"""Host input pipeline.

The reference decodes its batch inside the step loop, on the thread that then calls sess.run: `get_image` per file at
models/recurrent_z/model.py:212-219, `get_videos` (OpenCV mp4 decode + resize) at z_model_lib.py:332-351 -- with a
1.6 ms train step the decode of 64 JPEGs (or 32 x 16 video frames) is what the loop would wait for.  `Prefetcher` moves
that work off the step's thread without changing what the loop sees:

  * batches are produced IN ORDER by `load_item(item)` calls fanned out over a small thread pool (OpenCV releases the
    GIL while it decodes and resizes), `depth` batches ahead of the consumer and never further;
  * every batch is assembled in one of `depth + 3` reusable float32 buffers -- page-locked when a CUDA device is
    present, so `train_step`'s `copy_(non_blocking=True)` is a true asynchronous H2D copy; a buffer is handed out again
    only after the consumer has come back for the batch after the next one, i.e. after the step that read it has
    synchronised;
  * an exception in a loader surfaces in the consumer at the batch it belongs to; `close()` (or leaving the `with`
    block, or exhausting the iterator) stops the workers.

It yields `torch.Tensor` views `[len(items), *item_shape]` of those buffers.
"""
from __future__ import annotations

import queue
import threading
import weakref
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch


class _Shared(object):
    """Everything the producer thread and the decode workers touch.  They hold THIS object, never the Prefetcher, so that
    dropping the iterator mid-epoch lets it be collected -- its finalizer then stops the threads and releases the ring."""

    def __init__(self, batches, load_item, item_shape, depth, workers, pin, dtype=torch.float32):
        self.batches, self.load_item, self.item_shape, self.depth = batches, load_item, item_shape, depth
        self.dtype = dtype
        self.np_dtype = np.uint8 if dtype == torch.uint8 else np.float32
        rows = max((len(b) for b in batches), default=0)
        self.ring = [torch.empty((rows,) + item_shape, dtype=dtype) for _ in range(depth + 3)]
        if pin:
            self.ring = [t.pin_memory() for t in self.ring]
        self.free = queue.Queue()
        for i in range(len(self.ring)):
            self.free.put(i)
        self.ready = queue.Queue(maxsize=depth)
        self.stop = threading.Event()
        self.pool = ThreadPoolExecutor(max_workers=max(1, int(workers)), thread_name_prefix="gifgan-decode")
        self.max_ahead = 0                    # diagnostics: how far the producer ever was ahead of the consumer
        self.consumed = 0

    def fill(self, buf, row, item):
        a = self.load_item(item)
        if self.np_dtype == np.uint8 and np.asarray(a).dtype != np.uint8:
            raise TypeError("a uint8 Prefetcher needs uint8 items (raw decoded bytes)")
        buf[row].copy_(torch.from_numpy(np.ascontiguousarray(a, dtype=self.np_dtype).reshape(self.item_shape)))

    def take_free(self):
        while not self.stop.is_set():
            try:
                return self.free.get(timeout=0.05)
            except queue.Empty:
                continue
        return None

    def shutdown(self):
        self.stop.set()
        self.pool.shutdown(wait=False, cancel_futures=True)


def _produce(sh: _Shared):
    for k, items in enumerate(sh.batches):
        slot = sh.take_free()
        if slot is None:
            return
        buf = sh.ring[slot]
        err = None
        try:
            futures = [sh.pool.submit(sh.fill, buf, r, it) for r, it in enumerate(items)]
            for f in futures:
                f.result()
        except BaseException as e:        # delivered to the consumer in batch order
            err = e
        sh.max_ahead = max(sh.max_ahead, k + 1 - sh.consumed)
        while not sh.stop.is_set():
            try:
                sh.ready.put((slot, len(items), err), timeout=0.05)
                break
            except queue.Full:
                continue
        if err is not None or sh.stop.is_set():
            return


class Prefetcher(object):
    def __init__(self, batches, load_item, item_shape, depth=2, workers=4, pin=None, dtype=torch.float32):
        """`batches`: a sequence of batches, each a sequence of items (file names); `load_item(item)` -> array of
        `item_shape` (any float / int dtype; stored as float32).  dtype=torch.uint8: the buffers hold raw decoded bytes
        (items must be uint8) for a GPU-side decode tail (ops.frames_to_input): a quarter of the host-to-device bytes."""
        self.batches = list(batches)
        self.load_item, self.item_shape = load_item, tuple(item_shape)
        self.depth = max(1, int(depth))
        pin = torch.cuda.is_available() if pin is None else pin
        self._sh = sh = _Shared(self.batches, load_item, self.item_shape, self.depth, workers, pin, dtype)
        self._ring = sh.ring
        self._held = []                       # buffers the consumer may still be reading (newest last)
        self._producer = threading.Thread(target=_produce, args=(sh,), name="gifgan-prefetch", daemon=True)
        self._producer.start()
        self._finalizer = weakref.finalize(self, _Shared.shutdown, sh)     # runs when the Prefetcher is collected (or at exit)

    @property
    def max_ahead(self):
        return self._sh.max_ahead

    # ---- consumer ---------------------------------------------------------------------------------
    def __iter__(self):
        return self

    def __next__(self):
        sh = self._sh
        if sh.consumed >= len(self.batches):
            self.close()
            raise StopIteration
        # the batch handed out two calls ago is no longer in use: its step has synchronised before this call
        while len(self._held) >= 2:
            sh.free.put(self._held.pop(0))
        slot, n, err = sh.ready.get()
        sh.consumed += 1
        if err is not None:
            self.close()
            raise err
        self._held.append(slot)
        return sh.ring[slot][:n]

    def __len__(self):
        return len(self.batches)

    def close(self):
        self._finalizer()                    # idempotent: stop event + pool shutdown

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        if self._producer.is_alive():
            self._producer.join(timeout=2.0)


def chunks(items, batch_size, drop_last=True):
    """Consecutive batches of `batch_size` items (the reference's `data[idx*B:(idx+1)*B]`, model.py:212)."""
    n = len(items) // batch_size if drop_last else -(-len(items) // batch_size)
    return [items[i * batch_size:(i + 1) * batch_size] for i in range(n)]
