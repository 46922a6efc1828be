"""Video / frame-folder feeder: decode threads -> pinned host ring -> the GPU path.

Replaces the reference's frame extraction + per-frame ``Image.open`` loop
(scripts/cv2_frame_extraction.py:1-14, scripts/decord_frame_extraction.py:28-55 write every frame as a
JPEG; src/stable-diffusion/get_percep_embeddings.py:54-56,94-97 reads them back one at a time).  Here the
frames never touch the disk: ``n_decoders`` threads each decode a contiguous sub-range of this rank's frame
range with OpenCV (``cv2.VideoCapture.read`` releases the GIL) straight into slots of a ring of PINNED host
buffers, and the consumer uploads a whole slot with one async H2D copy.  Frames are independent on this path
(SURVEY F10), so slots are consumed in completion order, not frame order; every slot carries the index of its
first frame.

Sources
  VideoSource     a video file (cv2 / FFmpeg); frame k = the k-th decoded frame, RGB
  FrameDirSource  a folder of extracted frames (the reference's IMAGE_FOLDER): sorted basenames are the keys
  ArraySource     uint8 [N,H,W,3] already in memory (fixtures, synthetic frames)
"""
from __future__ import annotations

import os
import queue
import threading
import time

import numpy as np
import torch


class ArraySource:
    def __init__(self, frames, keys=None):
        self.frames = frames.numpy() if isinstance(frames, torch.Tensor) else np.asarray(frames)
        if self.frames.dtype != np.uint8 or self.frames.ndim != 4 or self.frames.shape[-1] != 3:
            raise ValueError(f"expected uint8 [N,H,W,3], got {self.frames.dtype} {self.frames.shape}")
        self.keys = keys

    def __len__(self):
        return self.frames.shape[0]

    @property
    def frame_hw(self):
        return int(self.frames.shape[1]), int(self.frames.shape[2])

    def key(self, index: int) -> str:
        return self.keys[index] if self.keys is not None else f"{index:010d}.jpg"

    def reader(self, lo: int, hi: int):
        def read_into(dst, start, n):
            dst[:n] = self.frames[start:start + n]
            return n
        return read_into


class FrameDirSource:
    """The reference's IMAGE_FOLDER (get_percep_embeddings.py:85): every file of the folder is a frame, its
    basename is the embedding key (:106).  Sorted, so the order is reproducible (glob order is not)."""

    EXT = (".jpg", ".jpeg", ".png", ".bmp")

    def __init__(self, folder: str):
        self.folder = folder
        self.names = sorted(n for n in os.listdir(folder) if n.lower().endswith(self.EXT))
        if not self.names:
            raise FileNotFoundError(f"No images found in the folder: {folder}")
        import cv2
        first = cv2.imread(os.path.join(folder, self.names[0]), cv2.IMREAD_COLOR)
        if first is None:
            raise ValueError(f"cannot decode {self.names[0]}")
        self._hw = (int(first.shape[0]), int(first.shape[1]))

    def __len__(self):
        return len(self.names)

    @property
    def frame_hw(self):
        return self._hw

    def key(self, index: int) -> str:
        return self.names[index]

    def reader(self, lo: int, hi: int):
        import cv2

        def read_into(dst, start, n):
            for i in range(n):
                img = cv2.imread(os.path.join(self.folder, self.names[start + i]), cv2.IMREAD_COLOR)
                if img is None or img.shape[:2] != self._hw:
                    raise ValueError(f"{self.names[start + i]}: undecodable or not {self._hw[1]}x{self._hw[0]}")
                cv2.cvtColor(img, cv2.COLOR_BGR2RGB, dst=dst[i])
            return n
        return read_into


class VideoSource:
    """A video file decoded with OpenCV.  Frame k is the k-th frame ``VideoCapture.read`` yields (what
    scripts/cv2_frame_extraction.py numbers ``%010d.jpg``); RGB order."""

    def __init__(self, path: str, key_ext: str = ".jpg"):
        import cv2
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path, self.key_ext = path, key_ext
        cap = cv2.VideoCapture(path)
        if not cap.isOpened():
            raise ValueError(f"cannot open video {path}")
        self._n = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        self._hw = (int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)))
        cap.release()
        if self._n <= 0:
            raise ValueError(f"{path}: container reports no frames")

    def __len__(self):
        return self._n

    @property
    def frame_hw(self):
        return self._hw

    def key(self, index: int) -> str:
        return f"{index:010d}{self.key_ext}"

    def reader(self, lo: int, hi: int):
        """A reader positioned at frame `lo` (its own capture: one per decode thread).  Seeking lands on the
        previous keyframe and decodes forward; if the backend cannot position exactly, fall back to decoding from
        the start and discarding."""
        import cv2
        cap = cv2.VideoCapture(self.path)
        pos = 0
        if lo > 0:
            cap.set(cv2.CAP_PROP_POS_FRAMES, lo)
            pos = int(round(cap.get(cv2.CAP_PROP_POS_FRAMES)))
            if pos != lo:
                cap.release()
                cap = cv2.VideoCapture(self.path)
                pos = 0
                while pos < lo and cap.grab():
                    pos += 1
        state = dict(pos=pos)

        def read_into(dst, start, n):
            if start != state["pos"]:
                raise RuntimeError(f"video reader is sequential: at frame {state['pos']}, asked for {start}")
            got = 0
            for i in range(n):
                ok, bgr = cap.read()
                if not ok:
                    break
                cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB, dst=dst[i])
                got += 1
            state["pos"] += got
            return got
        return read_into


class Slot:
    __slots__ = ("buf", "first", "n", "id")

    def __init__(self, buf, idx):
        self.buf, self.first, self.n, self.id = buf, 0, 0, idx


class PinnedRing:
    """`slots` page-locked uint8 buffers [batch,H,W,3].  Decode threads take a free slot, fill it, post it;
    the consumer returns it after its H2D copy has been enqueued AND completed (the caller records an event)."""

    def __init__(self, slots: int, batch: int, H: int, W: int, pin: bool | None = None):
        pin = torch.cuda.is_available() if pin is None else pin
        self.slots = [Slot(torch.empty((batch, H, W, 3), dtype=torch.uint8, pin_memory=pin), i) for i in range(slots)]
        self.free: queue.Queue = queue.Queue()
        self.ready: queue.Queue = queue.Queue()
        for s in self.slots:
            self.free.put(s)
        self.pinned = pin


class Feeder:
    """Decode [lo, hi) of `source` into a PinnedRing with `n_decoders` threads.

    Iterate to receive filled slots (completion order); call ``release(slot)`` when the slot's contents have been
    consumed.  ``stats()`` reports the decode rate: frames decoded / time the decode threads were busy, i.e. the
    decode-bound ceiling of the job."""

    def __init__(self, source, lo: int = 0, hi: int | None = None, batch: int = 32, slots: int = 4, n_decoders: int = 2,
                 pin: bool | None = None):
        hi = len(source) if hi is None else hi
        if not (0 <= lo <= hi <= len(source)):
            raise ValueError(f"bad frame range [{lo},{hi}) for {len(source)} frames")
        self.source, self.lo, self.hi, self.batch = source, lo, hi, batch
        H, W = source.frame_hw
        self.ring = PinnedRing(max(2, slots), batch, H, W, pin)
        n = hi - lo
        n_decoders = max(1, min(n_decoders, (n + batch - 1) // batch)) if n else 1
        # contiguous sub-ranges, aligned to whole batches so only the last slot of the range is short
        nb = (n + batch - 1) // batch
        per = [(nb * t) // n_decoders for t in range(n_decoders + 1)]
        self.ranges = [(lo + per[t] * batch, min(hi, lo + per[t + 1] * batch)) for t in range(n_decoders)]
        self._threads = []
        self._err = None
        self._busy = [0.0] * n_decoders
        self._decoded = [0] * n_decoders
        self._remaining = n_decoders
        self._lock = threading.Lock()
        self._stop = False
        for t, (a, b) in enumerate(self.ranges):
            th = threading.Thread(target=self._run, args=(t, a, b), daemon=True)
            self._threads.append(th)
        self._t0 = time.perf_counter()
        for th in self._threads:
            th.start()

    def _run(self, t, a, b):
        try:
            read_into = self.source.reader(a, b) if b > a else None
            pos = a
            while pos < b and not self._stop:
                slot = self.ring.free.get()
                if slot is None:
                    break
                t0 = time.perf_counter()
                want = min(self.batch, b - pos)
                got = read_into(slot.buf.numpy(), pos, want)
                self._busy[t] += time.perf_counter() - t0
                if got != want:
                    raise RuntimeError(f"{getattr(self.source, 'path', 'source')}: frame {pos + got} could not be decoded "
                                       f"(range [{a},{b}))")
                slot.first, slot.n = pos, got
                self._decoded[t] += got
                pos += got
                self.ring.ready.put(slot)
        except BaseException as e:          # surfaced to the consumer
            self._err = e
        finally:
            with self._lock:
                self._remaining -= 1
                if self._remaining == 0:
                    self.ring.ready.put(None)

    def __iter__(self):
        while True:
            slot = self.ring.ready.get()
            if self._err is not None:
                self.close()
                raise self._err
            if slot is None:
                return
            yield slot

    def release(self, slot):
        self.ring.free.put(slot)

    def close(self):
        self._stop = True
        for _ in self._threads:
            self.ring.free.put(None)

    def stats(self):
        busy = max(self._busy) if self._busy else 0.0
        n = sum(self._decoded)
        return dict(frames=n, decoders=len(self._threads), decode_busy_s=busy,
                    decode_fps=(n / busy) if busy > 0 else None, pinned=self.ring.pinned,
                    wall_s=time.perf_counter() - self._t0)
