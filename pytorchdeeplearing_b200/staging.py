"""Device-side input staging: the per-sample host work of the reference's 2-D dataset, done on the GPU.

Reference: ``datasetModelSegwithopencv.__getitem__`` (model/dataset.py:128-158) z-scores every 8-bit image on the host
in float64 (``(image - image.mean()) / image.std()``), converts it to fp32 (4 B/pixel) and the label to int64
(8 B/pixel); the trainer then binarises the label (``y[y != 0] = 1``, model/modelUnet.py:130) and copies both to the
device (``:132``).  Here the raw uint8 image and label batches (1 B/pixel each) are copied from pinned host memory and
normalised / widened by two small kernels (csrc/staging.cu), on a side stream, one batch ahead of the step that consumes
them.  File reading, ``cv2.imread`` and ``cv2.resize`` stay where they are (host I/O, out of scope); the 3-D dataset
(``datasetModelSegwithnpy``, dataset.py:82-117) stores already-normalised float volumes and needs no arithmetic.

    stager = InputStager(device, (8, 512, 512))
    stager.put(images_u8, labels_u8)           # host uint8 arrays/tensors [N, H, W]; returns immediately
    x, y = stager.get()                        # x [N, 1, H, W] fp32, y [N, H, W] int64, ordered after the current stream
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import runtime


class InputStager:
    """Double-buffered uint8 -> (z-scored float image, int64 label) staging for a fixed batch shape [N, *spatial]."""

    def __init__(self, device, shape: Sequence[int], dtype: torch.dtype = torch.float32, binarize_labels: bool = True,
                 slots: int = 2, backend=None):
        self.device = torch.device(device)
        self.shape = tuple(int(v) for v in shape)
        self.dtype = dtype
        self.binarize = bool(binarize_labels)
        self.be = backend if backend is not None else runtime.get_backend(torch.empty(0, device=self.device))
        cuda = self.device.type == "cuda"
        self._stream = torch.cuda.Stream(self.device) if cuda else None
        self._slots = []
        for _ in range(max(1, int(slots))):
            h_img = torch.empty(self.shape, dtype=torch.uint8, pin_memory=cuda)
            h_lab = torch.empty(self.shape, dtype=torch.uint8, pin_memory=cuda)
            d_img = torch.empty(self.shape, dtype=torch.uint8, device=self.device)
            d_lab = torch.empty(self.shape, dtype=torch.uint8, device=self.device)
            x = torch.empty((self.shape[0], 1) + self.shape[1:], dtype=dtype, device=self.device)
            y = torch.empty(self.shape, dtype=torch.int64, device=self.device)
            ev = torch.cuda.Event() if cuda else None
            self._slots.append(dict(h_img=h_img, h_lab=h_lab, d_img=d_img, d_lab=d_lab, x=x, y=y, ready=ev, free=None))
        self._put = 0
        self._got = 0
        self.h2d_bytes_per_batch = 2 * int(torch.Size(self.shape).numel())

    def put(self, images, labels=None) -> None:
        """Queue one batch: uint8 images [N, *spatial] (numpy array or tensor) and, optionally, uint8 labels."""
        if self._put - self._got >= len(self._slots):
            raise RuntimeError("InputStager: every slot holds a batch that has not been taken (call get())")
        s = self._slots[self._put % len(self._slots)]
        img = torch.as_tensor(images)
        if img.dtype != torch.uint8 or tuple(img.shape) != self.shape:
            raise ValueError(f"InputStager: expected uint8 images of shape {self.shape}, got {img.dtype} {tuple(img.shape)}")
        lab = None
        if labels is not None:
            lab = torch.as_tensor(labels)
            if lab.dtype != torch.uint8 or tuple(lab.shape) != self.shape:
                raise ValueError(f"InputStager: expected uint8 labels of shape {self.shape}, got {lab.dtype} {tuple(lab.shape)}")
        if s["free"] is not None:
            s["free"].synchronize()            # the step that consumed this slot's last batch has finished with it
        s["h_img"].copy_(img)
        if lab is not None:
            s["h_lab"].copy_(lab)
        s["has_labels"] = lab is not None
        if self._stream is None:
            self._run(s)
        else:
            with torch.cuda.stream(self._stream):
                self._run(s)
                s["ready"].record(self._stream)
        self._put += 1

    def _run(self, s) -> None:
        s["d_img"].copy_(s["h_img"], non_blocking=True)
        self.be.stage_images_u8(s["d_img"], s["x"])
        if s["has_labels"]:
            s["d_lab"].copy_(s["h_lab"], non_blocking=True)
            self.be.stage_labels_u8(s["d_lab"], s["y"], self.binarize)

    def get(self) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """The oldest queued batch: (x [N, 1, *spatial], y [N, *spatial] int64 or None).  The tensors belong to the slot:
        they are valid until ``slots`` further batches have been queued."""
        if self._got >= self._put:
            raise RuntimeError("InputStager: no batch queued")
        s = self._slots[self._got % len(self._slots)]
        self._got += 1
        if self._stream is not None:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(s["ready"])
            ev = torch.cuda.Event()
            s["free"] = ev
            ev.record(cur)                     # refined by release(); until then: "whatever was queued before get()"
        return s["x"], (s["y"] if s["has_labels"] else None)

    def release(self) -> None:
        """Mark the batch taken by the last ``get()`` as consumed by everything queued on the current stream so far
        (call after the step's launches; ``put`` waits on it before overwriting that slot)."""
        s = self._slots[(self._got - 1) % len(self._slots)]
        if s.get("free") is not None:
            s["free"].record(torch.cuda.current_stream(self.device))


def stage_batch(images_u8: torch.Tensor, labels_u8: Optional[torch.Tensor] = None, dtype: torch.dtype = torch.float32,
                binarize_labels: bool = True, backend=None):
    """One-shot form on tensors that are already on the device: -> (x [N, 1, ...], y int64 or None)."""
    be = backend if backend is not None else runtime.get_backend(images_u8)
    x = torch.empty((images_u8.shape[0], 1) + tuple(images_u8.shape[1:]), dtype=dtype, device=images_u8.device)
    be.stage_images_u8(images_u8.contiguous(), x)
    y = None
    if labels_u8 is not None:
        y = torch.empty(labels_u8.shape, dtype=torch.int64, device=labels_u8.device)
        be.stage_labels_u8(labels_u8.contiguous(), y, binarize_labels)
    return x, y
