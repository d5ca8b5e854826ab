"""The data contract in front of the hot path: what the reference's datasets hand to the training loop, and a
pinned-memory feeder that moves it onto the GPU ahead of the step.

Reference: data/cocostuff_loader.py:222-380 (`__getitem__` -> image FloatTensor (3,H,W) in [-1,1], objs LongTensor (O,),
boxes float (O,4) xywh in [0,1]); images with fewer than `max_objects_per_image` objects are padded with the
`__image__` label 0 and the box [-0.6,-0.6,0.5,0.5] (:301-303); the loop casts with `label.long()`, `bbox.float()`
(train_context_app_v2.py:153).  Dataset decoding itself (COCO / VG annotation files) is out of scope.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Sequence, Tuple

import numpy as np
import torch

PAD_LABEL = 0                              # vocab['object_name_to_idx']['__image__']
PAD_BOX = (-0.6, -0.6, 0.5, 0.5)           # cocostuff_loader.py:301-303


def pack_layout(image, objs: Sequence[int], boxes, max_objects: int = 8) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """One dataset item in the loader's output format: (image (3,H,W) f32, objs (max_objects,) i64, boxes (max_objects,4) f32),
    padded at the END with (PAD_LABEL, PAD_BOX) exactly as cocostuff_loader.py:299-305 does.  More than max_objects raises
    (the reference datasets filter such images out at construction, :160-190)."""
    objs = list(int(o) for o in objs)
    boxes = [np.asarray(b, dtype=np.float64).reshape(4) for b in boxes]
    if len(objs) != len(boxes):
        raise ValueError(f"{len(objs)} labels for {len(boxes)} boxes")
    if len(objs) > max_objects:
        raise ValueError(f"{len(objs)} objects > max_objects_per_image = {max_objects}")
    for _ in range(len(objs), max_objects):
        objs.append(PAD_LABEL)
        boxes.append(np.array(PAD_BOX))
    image = torch.as_tensor(image, dtype=torch.float32)
    if image.dim() != 3 or image.shape[0] != 3:
        raise ValueError(f"image must be (3,H,W), got {tuple(image.shape)}")
    return image, torch.LongTensor(objs), torch.from_numpy(np.vstack(boxes)).float()


def collate(items) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """torch's default collate for (image, objs, boxes) items -> (b,3,H,W) f32, (b,O) i64, (b,O,4) f32."""
    imgs, objs, boxes = zip(*items)
    return torch.stack(imgs), torch.stack(objs), torch.stack([torch.as_tensor(b).float() for b in boxes])


class PinnedFeeder:
    """Wraps any iterable of (real_images, label, bbox) CPU batches (a DataLoader over the reference's datasets, or
    synth batches): every batch is staged in one of `depth` pinned host buffers and copied to the device on a side
    stream while the previous step is still running; iteration yields device tensors (real f32, label i64 (b,O), bbox
    f32) that the compute stream has been ordered after.  `bytes_per_batch` is what bench.py reports as H2D traffic."""

    def __init__(self, loader: Iterable, device, depth: int = 2):
        self.loader, self.device, self.depth = loader, torch.device(device), max(2, int(depth))
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [None] * self.depth
        self.bytes_per_batch = 0

    def _stage(self, slot: int, batch):
        real, label, bbox = batch[0], batch[1], batch[2]
        srcs = (real.float(), label.long().view(label.shape[0], -1), bbox.float())
        bufs = self.slots[slot]
        if bufs is None or any(b["host"].shape != s.shape for b, s in zip(bufs, srcs)):
            bufs = [{"host": torch.empty(s.shape, dtype=s.dtype).pin_memory(),
                     "dev": torch.empty(s.shape, dtype=s.dtype, device=self.device), "free": None} for s in srcs]
            self.slots[slot] = bufs
        for b, s in zip(bufs, srcs):
            if b["free"] is not None:
                b["free"].synchronize()               # the compute stream is done with this slot's device tensors
            b["host"].copy_(s)
        self.bytes_per_batch = sum(s.numel() * s.element_size() for s in srcs)
        with torch.cuda.stream(self.stream):
            for b in bufs:
                b["dev"].copy_(b["host"], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        return bufs, ready

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        it = iter(self.loader)
        pending, slot = [], 0
        try:
            for _ in range(self.depth - 1):
                pending.append(self._stage(slot, next(it)))
                slot = (slot + 1) % self.depth
        except StopIteration:
            pass
        while pending:
            bufs, ready = pending.pop(0)
            try:
                pending.append(self._stage(slot, next(it)))
                slot = (slot + 1) % self.depth
            except StopIteration:
                pass
            torch.cuda.current_stream(self.device).wait_event(ready)
            yield tuple(b["dev"] for b in bufs)
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))
            for b in bufs:
                b["free"] = done
