"""Evaluation counters on the device (``da_model.evaluate``, framework/domain_adaptation/methods/adaptation_model.py:127-166).

The reference upsamples every prediction to full resolution (``self.interp``, bilinear, ``align_corners=True``),
softmaxes it, takes the per-image argmax, copies it to the host and bincounts there (``fast_hist``,
framework/utils/func.py:77-79).  ``ConfusionMeter.update`` does interpolation + argmax + counting in one kernel that never
materialises the upsampled tensor, and keeps the ``C x C`` counters on the device until they are read.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nat


class ConfusionMeter:
    """Accumulates ``fast_hist(label, prediction, num_classes)`` over batches; ``hist()`` / ``per_class_iu()`` read it."""

    def __init__(self, num_classes, device="cuda"):
        if not 0 < num_classes <= nat.MAX_CLASSES:
            raise ValueError(f"{num_classes} classes unsupported (max {nat.MAX_CLASSES})")
        self.num_classes = num_classes
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("onda_b200 runs on CUDA devices only (there is no CPU path)")
        self._lib = nat.load()
        self._hist = torch.zeros((num_classes * num_classes,), dtype=torch.int64, device=self.device)

    def reset(self):
        self._hist.zero_()

    def update(self, pred, labels, return_prediction=False):
        """``pred``: (B, C, h, w) float32 logits (or probabilities) at network resolution, on the device;
        ``labels``: (B, H, W) ground truth of any integer / float dtype, host or device (values outside
        ``[0, num_classes)`` -- 255, -1 -- are ignored).  Returns the (B, H, W) uint8 prediction if asked."""
        if not (isinstance(pred, torch.Tensor) and pred.is_cuda):
            raise RuntimeError("onda_b200 runs on CUDA devices only: pred must be a CUDA tensor (there is no CPU path)")
        if pred.dim() != 4 or pred.shape[1] != self.num_classes:
            raise ValueError(f"pred must be (B, {self.num_classes}, h, w), got {tuple(pred.shape)}")
        labels = torch.as_tensor(labels)
        if labels.dim() != 3 or labels.shape[0] != pred.shape[0]:
            raise ValueError(f"labels must be (B, H, W) with B={pred.shape[0]}, got {tuple(labels.shape)}")
        pred = pred.detach().to(torch.float32).contiguous()
        lab = labels.to(device=pred.device, dtype=torch.int64, non_blocking=True).contiguous()   # like a.astype(int)
        B, C, h, w = pred.shape
        _, H, W = lab.shape
        out = torch.empty((B, H, W), dtype=torch.uint8, device=pred.device) if return_prediction else None
        with torch.cuda.device(pred.device):          # the C ABI launches on the current device
            stream = nat.C.c_void_p(torch.cuda.current_stream(pred.device).cuda_stream)
            nat.check(self._lib.onda_confusion_update(nat.ptr(pred), B, C, h, w, nat.ptr(lab), H, W, nat.ptr(self._hist),
                                                      nat.ptr(out), stream), "onda_confusion_update")
        return out

    def hist(self):
        """The accumulated (C, C) int64 matrix, row = label, column = prediction (what ``counters[key]`` holds)."""
        return self._hist.view(self.num_classes, self.num_classes).cpu().numpy()

    def per_class_iu(self):
        """framework/utils/func.py:82-85."""
        hist = self.hist()
        return np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist) + np.finfo(float).eps)
