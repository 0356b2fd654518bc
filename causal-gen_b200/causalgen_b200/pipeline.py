"""Input pipeline of the training step: pinned, double-buffered host -> device staging of uint8 batches on a
copy stream, and the reference's train-time augmentation on the device.

Reference behaviour covered (citations relative to /root/reference):
  * DataLoader(pin_memory=True) batches of uint8 images + float parents (src/train_setup.py:28-39), then
    preprocess_batch (src/trainer.py:16-21) -- the normalisation itself is `cg_normalise_u8` inside Trainer;
  * RandomCrop(size, padding) [+ RandomHorizontalFlip(p)] of the UKBB / Morpho-MNIST / Colour-MNIST train splits
    (src/datasets.py:107-118,281-286,371) as ONE kernel over the staged uint8 batch (`cg_augment_u8`); the random
    draws (top, left, flip) come from a host generator, one triple per sample, like torchvision's per-sample draws.

`DeviceLoader` wraps any iterator of (x_uint8 (B,C,H,W), parents (B,ctx)) host batches: while the trainer computes on
batch i, batch i+1 is copied into the other pinned slot and sent H2D on the copy stream; the compute stream only waits
on the copy's event.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional, Tuple

import torch

from . import _lib as L


class Augment:
    """RandomCrop(size=res, padding=(pad_left, pad_top), fill=0) + RandomHorizontalFlip(p) on the device"""

    def __init__(self, res: int, pad_top: int, pad_left: Optional[int] = None, hflip: float = 0.0, seed: int = 0):
        self.res, self.pad_top = int(res), int(pad_top)
        self.pad_left = int(pad_top if pad_left is None else pad_left)
        self.hflip = float(hflip)
        self.gen = torch.Generator().manual_seed(seed)

    @classmethod
    def ukbb(cls, args, seed: int = 0):       # src/datasets.py:111-115: padding=[2 * pad, pad] = (left/right, top/bottom)
        return cls(args.input_res, args.pad, 2 * args.pad, getattr(args, "hflip", 0.0), seed)

    @classmethod
    def mnist(cls, args, seed: int = 0):      # src/datasets.py:284,371
        return cls(args.input_res, args.pad, args.pad, 0.0, seed)

    def draw(self, n: int, hi: int, wi: int) -> torch.Tensor:
        """(n, 3) int32 host tensor of (top, left, flip), torchvision's ranges (RandomCrop.get_params: randint(0, h - th + 1))"""
        top = torch.randint(0, hi + 2 * self.pad_top - self.res + 1, (n,), generator=self.gen)
        left = torch.randint(0, wi + 2 * self.pad_left - self.res + 1, (n,), generator=self.gen)
        flip = (torch.rand(n, generator=self.gen) < self.hflip).long()
        return torch.stack([top, left, flip], 1).to(torch.int32)

    def apply(self, x8: torch.Tensor, params: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x8 (N,C,Hi,Wi) uint8 on the device, params (N,3) int32 on the device -> (N,C,res,res) uint8"""
        n, c, hi, wi = x8.shape
        if out is None:
            out = torch.empty(n, c, self.res, self.res, dtype=torch.uint8, device=x8.device)
        L.check(L.load().cg_augment_u8(x8.data_ptr(), out.data_ptr(), params.data_ptr(), n, c, hi, wi, self.res,
                                       self.pad_top, self.pad_left, torch.cuda.current_stream().cuda_stream),
                "cg_augment_u8")
        return out

    @staticmethod
    def reference(x8: torch.Tensor, params: torch.Tensor, res: int, pad_top: int, pad_left: int) -> torch.Tensor:
        """the same transform written with torch ops (what torchvision's F.pad + F.crop + F.hflip compute); host-side
        checker for the tests"""
        import torch.nn.functional as F
        xp = F.pad(x8, [pad_left, pad_left, pad_top, pad_top])
        out = []
        for i in range(x8.shape[0]):
            t, l, f = (int(v) for v in params[i])
            crop = xp[i, :, t: t + res, l: l + res]
            out.append(crop.flip(-1) if f else crop)
        return torch.stack(out)


class DeviceLoader:
    """pinned staging + async H2D on a copy stream, `depth` slots, fed by a background thread (+ optional on-device
    augmentation on the consumer's stream).  While the consumer computes on batch i the thread copies batch i+1 into the
    next pinned slot and sends it to the device; the compute stream only waits on that copy's event.  A slot is reused
    after the consumer has come back for the NEXT batch (everything it enqueued on the old one is then ordered before
    the slot's `consumed` event) and that event has completed."""

    def __init__(self, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], device=None, augment: Optional[Augment] = None,
                 depth: int = 2):
        self.src = batches
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.augment = augment
        self.depth = max(2, int(depth))
        self.copy_stream = torch.cuda.Stream(device=self.device)

    def _new_slot(self, x: torch.Tensor, pa: torch.Tensor):
        s = dict(xh=torch.empty(x.shape, dtype=torch.uint8).pin_memory(), ph=torch.empty(pa.shape, dtype=torch.float32).pin_memory(),
                 xd=torch.empty(x.shape, dtype=torch.uint8, device=self.device),
                 pd=torch.empty(pa.shape, dtype=torch.float32, device=self.device),
                 ready=torch.cuda.Event(), consumed=torch.cuda.Event(), shape=(tuple(x.shape), tuple(pa.shape)))
        if self.augment is not None:
            s["prm_h"] = torch.empty(x.shape[0], 3, dtype=torch.int32).pin_memory()
            s["prm_d"] = torch.empty(x.shape[0], 3, dtype=torch.int32, device=self.device)
            s["xa"] = torch.empty(x.shape[0], x.shape[1], self.augment.res, self.augment.res, dtype=torch.uint8,
                                  device=self.device)
        return s

    def _stage(self, slot, x: torch.Tensor, pa: torch.Tensor):
        slot["consumed"].synchronize()          # the compute stream is done with this slot's buffers (no-op when fresh)
        slot["xh"].copy_(x)                     # host -> pinned (what a pin_memory DataLoader thread does)
        slot["ph"].copy_(pa.to(torch.float32))
        if self.augment is not None:
            slot["prm_h"].copy_(self.augment.draw(x.shape[0], x.shape[2], x.shape[3]))
        with torch.cuda.stream(self.copy_stream):
            slot["xd"].copy_(slot["xh"], non_blocking=True)
            slot["pd"].copy_(slot["ph"], non_blocking=True)
            if self.augment is not None:
                slot["prm_d"].copy_(slot["prm_h"], non_blocking=True)
            slot["ready"].record(self.copy_stream)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        import queue
        import threading
        free_q: "queue.Queue" = queue.Queue()
        ready_q: "queue.Queue" = queue.Queue()
        dev = self.device

        def feeder():
            made = 0
            try:
                with torch.cuda.device(dev):
                    for x, pa in self.src:
                        slot = None
                        if made < self.depth:
                            slot, made = self._new_slot(x, pa), made + 1
                        else:
                            slot = free_q.get()
                            if slot is None:
                                return
                            if slot["shape"] != (tuple(x.shape), tuple(pa.shape)):   # ragged last batch
                                slot = self._new_slot(x, pa)
                        self._stage(slot, x, pa)
                        ready_q.put(slot)
                ready_q.put(None)
            except BaseException as ex:  # surface loader errors in the consumer
                ready_q.put(ex)

        th = threading.Thread(target=feeder, daemon=True)
        th.start()
        prev = None
        try:
            while True:
                slot = ready_q.get()
                cur = torch.cuda.current_stream(dev)
                if prev is not None:            # the consumer is back: its work on the previous batch is enqueued
                    prev["consumed"].record(cur)
                    free_q.put(prev)
                    prev = None
                if slot is None:
                    break
                if isinstance(slot, BaseException):
                    raise slot
                cur.wait_event(slot["ready"])
                x = slot["xd"]
                if self.augment is not None:
                    x = self.augment.apply(slot["xd"], slot["prm_d"], slot["xa"])
                prev = slot
                yield x, slot["pd"]
        finally:
            free_q.put(None)
