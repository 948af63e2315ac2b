"""Teacher-dump writer and reader.

The product of the reference's "teacher inference" is one ``.npz`` per image
(``teacher_inference.py:83-90``):

    np.savez_compressed(out_path, pred_heatmaps=preds[:17], embeddings=preds[17:],
                        heatmaps_refined=refined, heatmaps_order=HEATMAPS_ORDER)

with ``out_path = <out_dir>/<basename(img_path)>_w48_predictions`` (``:67-68``), read back by
``rtpe/dataloaders.py:140-165`` (``heatmaps_refined`` + ``embeddings``, optionally resized with
bilinear ``align_corners=True``).  Once the forward runs at ~2000 images/s per GPU the deflate of
10.4 MB of float32 per image is what bounds the dump, so the writer here

  * takes whole batches from the device with ONE device->host copy per output into pinned
    buffers (double buffered, on a side stream, so the copy of batch i overlaps the forward of
    batch i+1), and
  * compresses / writes the files on a pool of host threads (zlib releases the GIL), with a
    bounded queue so that a slow disk applies back-pressure instead of exhausting memory.

The file format is the reference's, byte-for-byte at the ``np.load`` level: same four keys, same
dtypes (float32 maps, the ``<U7`` string array of joint names), same shapes.
"""
import os
import queue
import threading

import numpy as np
import torch

from . import _lib as L
from . import inference

# teacher_inference.py:38-40
HEATMAPS_ORDER = ["nose", "leye", "reye", "lear", "rear", "lshould", "rshould",
                  "lelbow", "relbow", "lwrist", "rwrist", "lhip", "rhip",
                  "lknee", "rknee", "lankle", "rankle"]


def dump_path(out_dir, img_path):
    """teacher_inference.py:67-68 (np.savez_compressed appends '.npz')."""
    return os.path.join(out_dir, os.path.basename(img_path)) + "_w48_predictions"


def write_teacher_npz(out_path, preds, refined, num_joints=17):
    """One image: preds (2J, H/4, W/4) float32, refined (J, H/2, W/2) float32 (numpy)."""
    np.savez_compressed(out_path,
                        pred_heatmaps=preds[:num_joints],
                        embeddings=preds[num_joints:],
                        heatmaps_refined=refined,
                        heatmaps_order=HEATMAPS_ORDER)


class TeacherDumpWriter:
    """Bounded pool of compression threads writing reference-format ``.npz`` files.

    ``submit`` copies nothing: the caller hands over numpy arrays it will not touch again (the
    batch helpers below hand over slices of a pinned buffer and recycle the buffer through
    ``on_done``).  Errors raised in a worker are re-raised by ``submit`` / ``close``."""

    def __init__(self, out_dir, workers=None, max_pending=None, num_joints=17):
        self.out_dir = out_dir
        os.makedirs(out_dir, exist_ok=True)
        self.num_joints = int(num_joints)
        self.workers = int(workers) if workers else max(1, (os.cpu_count() or 2) - 1)
        self._q = queue.Queue(maxsize=max_pending or 4 * self.workers)
        self._err = None
        self._written = 0
        self._lock = threading.Lock()
        self._threads = [threading.Thread(target=self._run, daemon=True) for _ in range(self.workers)]
        for t in self._threads:
            t.start()

    def _run(self):
        while True:
            job = self._q.get()
            if job is None:
                self._q.task_done()
                return
            path, preds, refined, on_done = job
            try:
                write_teacher_npz(path, preds, refined, self.num_joints)
                with self._lock:
                    self._written += 1
            except BaseException as e:  # noqa: BLE001 -- handed to the submitting thread
                with self._lock:
                    if self._err is None:
                        self._err = e
            finally:
                if on_done is not None:
                    on_done()
                self._q.task_done()

    def _check(self):
        with self._lock:
            if self._err is not None:
                e, self._err = self._err, None
                raise e

    def submit(self, img_path, preds, refined, on_done=None):
        self._check()
        if preds.shape[0] != 2 * self.num_joints or refined.shape[0] != self.num_joints:
            raise ValueError("expected preds (%d,h,w) and refined (%d,H,W), got %s and %s" % (
                2 * self.num_joints, self.num_joints, tuple(preds.shape), tuple(refined.shape)))
        self._q.put((dump_path(self.out_dir, img_path), preds, refined, on_done))

    def flush(self):
        self._q.join()
        self._check()

    @property
    def written(self):
        with self._lock:
            return self._written

    def close(self):
        self._q.join()
        for _ in self._threads:
            self._q.put(None)
        for t in self._threads:
            t.join()
        self._check()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class _PinnedSlot:
    """Pinned host copies of one batch of network outputs + the event that says they landed."""

    def __init__(self, y0, y1):
        self.y0 = torch.empty(y0.shape, dtype=torch.float32, pin_memory=True)
        self.y1 = torch.empty(y1.shape, dtype=torch.float32, pin_memory=True)
        self.event = torch.cuda.Event()
        self.pending = 0
        self.cv = threading.Condition()

    def release_one(self):
        with self.cv:
            self.pending -= 1
            if self.pending == 0:
                self.cv.notify_all()

    def wait_free(self):
        with self.cv:
            while self.pending:
                self.cv.wait()


class TeacherDumper:
    """``teacher_inference.py:65-90`` for batches of equally sized, already pre-processed images:
    forward on the GPU, one D2H copy per output per batch into pinned double buffers on a side
    stream, ``.npz`` files written by a ``TeacherDumpWriter``.

    ``model``: this package's network (``get_hrnet_w48_teacher``), returning ``[y0, y1]``."""

    def __init__(self, model, writer):
        if not torch.cuda.is_available():
            raise L.BrtpeError("TeacherDumper needs a CUDA device (there is no CPU fallback)")
        self.model = model
        self.writer = writer
        self._slots = {}
        self._turn = 0
        self._copy_stream = torch.cuda.Stream()
        self._inflight = None                      # (slot, paths) whose D2H copy is still running

    def _finish_inflight(self):
        if self._inflight is None:
            return
        slot, paths = self._inflight
        self._inflight = None
        slot.event.synchronize()                   # the copy has landed in the pinned buffers
        p0, p1 = slot.y0.numpy(), slot.y1.numpy()
        submitted = 0
        try:
            for i, path in enumerate(paths):
                self.writer.submit(path, p0[i], p1[i], on_done=slot.release_one)
                submitted += 1
        except BaseException:
            # submit() re-raises worker errors (disk full, ...): the files that were never handed over
            # will not call release_one, so take them off the slot's count -- otherwise the next
            # dump_batch / close that re-uses the slot waits for ever instead of seeing the exception
            with slot.cv:
                slot.pending -= len(paths) - submitted
                if slot.pending <= 0:
                    slot.pending = 0
                    slot.cv.notify_all()
            raise

    @torch.no_grad()
    def dump_batch(self, x, img_paths):
        """x (N,3,H,W) CUDA tensor of pre-processed images; one file per entry of ``img_paths``.
        The forward and the D2H copy of this batch are only ENQUEUED here; the files of the previous
        batch are handed to the writer while they run (``close`` flushes the last batch)."""
        if x.shape[0] != len(img_paths):
            raise ValueError("%d images but %d paths" % (x.shape[0], len(img_paths)))
        y0, y1 = self.model(x)
        y0 = y0.float()
        y1 = y1.float()
        key = (tuple(y0.shape), tuple(y1.shape), self._turn & 1)
        self._turn += 1
        slot = self._slots.get(key)
        if slot is None:
            slot = self._slots[key] = _PinnedSlot(y0, y1)
        if self._inflight is not None and self._inflight[0] is slot:
            self._finish_inflight()                       # same buffers (shape changed back): drain first
        slot.wait_free()                                  # its previous files are on disk
        done = torch.cuda.Event()
        done.record()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(done)
            slot.y0.copy_(y0, non_blocking=True)
            slot.y1.copy_(y1, non_blocking=True)
            slot.event.record()
        y0.record_stream(self._copy_stream)
        y1.record_stream(self._copy_stream)
        slot.pending = len(img_paths)
        self._finish_inflight()                           # previous batch: overlaps this batch's GPU work
        self._inflight = (slot, list(img_paths))

    def close(self):
        self._finish_inflight()
        self.writer.flush()


def load_teacher_data(teacher_dir, img_id, out_hw=None, device="cuda"):
    """``COCODataset._get_teacher_data`` (rtpe/dataloaders.py:140-165): ``(t_hms, t_ae)`` =
    ``heatmaps_refined`` and ``embeddings`` of ``<img_id>.jpg_w48_predictions.npz`` as float
    tensors, resized to ``out_hw`` with bilinear ``align_corners=True`` when given (on the GPU
    through ``brtpe_bilinear_resize``; the result stays on ``device``).  ``teacher_dir=None``
    returns two empty tensors like the reference."""
    if teacher_dir is None:
        return torch.zeros(0), torch.zeros(0)
    with np.load(os.path.join(teacher_dir, img_id + ".jpg_w48_predictions.npz")) as npz:
        t_hms = torch.from_numpy(np.ascontiguousarray(npz["heatmaps_refined"], dtype=np.float32))
        t_ae = torch.from_numpy(np.ascontiguousarray(npz["embeddings"], dtype=np.float32))
    if out_hw is None:
        return t_hms, t_ae
    dev = torch.device(device)
    if dev.type != "cuda":
        raise L.BrtpeError("load_teacher_data(out_hw=...) resizes on the GPU; device must be CUDA")
    t_hms = inference.bilinear_resize(t_hms.to(dev).unsqueeze(0), out_hw, True)[0]
    t_ae = inference.bilinear_resize(t_ae.to(dev).unsqueeze(0), out_hw, True)[0]
    return t_hms, t_ae
