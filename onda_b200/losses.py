"""Loss-side consumers of the pseudo-labels on one fused kernel (``online_proDA.pseudolabel_loss``,
framework/domain_adaptation/methods/prototypes.py:313-336).

The reference computes, on the student logits ``out`` (B, C, h, w) and the hard pseudo-labels,
``loss_calc`` (``cross_entropy_2d``, framework/utils/loss.py:16-45: a boolean-mask gather of the logits followed by
``F.cross_entropy``), ``rce`` (loss.py:88-112: softmax, one-hot, clamp, log, masked sum) and ``regular_loss``
(prototypes.py:29-39), then lets autograd walk all of it backwards.  ``target_losses`` returns the same scalars from
``onda_target_loss_fused``: one pass forward that also writes the gradient of the weighted total, wrapped in a
``torch.autograd.Function`` so that ``total.backward()`` feeds the student network exactly as before.

Only the configuration every shipped YAML uses is covered: hard labels (``SOFT_LABELS`` unset; the reference's soft
branch takes the logarithm of raw logits, loss.py:12-13,33-35) and no JS term (``JS_D`` unset).
"""
from __future__ import annotations

import torch

from . import _native as nat

_work = {}


def _workspace(lib, device):
    key = str(device)
    w = _work.get(key)
    if w is None:
        w = _work[key] = torch.zeros(lib.onda_target_loss_workspace_bytes(), dtype=torch.uint8, device=device)
    return w


class _TargetLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, labels, n_valid, alpha, beta, reg_weight, regularizer):
        lib = nat.load()
        if not out.is_cuda:
            raise RuntimeError("onda_b200 runs on CUDA devices only: the student logits are on " + str(out.device))
        if out.dim() != 4:
            raise ValueError(f"out must be (B, C, h, w), got {tuple(out.shape)}")
        B, C, h, w = out.shape
        if C > nat.MAX_CLASSES:
            raise ValueError(f"{C} classes unsupported (max {nat.MAX_CLASSES})")
        logits = out.detach().to(torch.float32).contiguous()
        lab = labels.detach().to(device=out.device, dtype=torch.int64).reshape(-1).contiguous()
        if lab.numel() != B * h * w:
            raise ValueError(f"labels must have {B * h * w} entries, got {tuple(labels.shape)}")
        need_grad = ctx.needs_input_grad[0]
        grad = torch.empty_like(logits) if need_grad else None
        out6 = torch.empty(6, dtype=torch.float32, device=out.device)
        work = _workspace(lib, out.device)
        nv = None
        if n_valid is not None:
            nv = n_valid.detach().to(device=out.device, dtype=torch.float32).reshape(-1)[:1].contiguous()
        with torch.cuda.device(out.device):
            stream = nat.C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)
            nat.check(lib.onda_target_loss_fused(nat.ptr(logits), nat.ptr(lab), B, C, h * w, nat.ptr(nv), float(alpha),
                                                 float(beta), float(reg_weight), nat.REGULARIZER[regularizer], nat.ptr(grad),
                                                 nat.ptr(out6), nat.ptr(work), work.numel(), stream), "onda_target_loss_fused")
        ctx.save_for_backward(grad if need_grad else out6.new_empty(0))
        ctx.out_dtype = out.dtype
        ctx.mark_non_differentiable(out6)
        return out6[3].clone(), out6

    @staticmethod
    def backward(ctx, g_total, _g_parts):
        (grad,) = ctx.saved_tensors
        if grad.numel() == 0:
            return (None,) * 7
        return (grad * g_total).to(ctx.out_dtype), None, None, None, None, None, None


def target_losses(out, pseudolabels, rce_alpha=0.1, rce_beta=1.0, regularizer_weight=0.1, regularizer="MRKLD", n_valid=None):
    """The target losses of ``pseudolabel_loss`` (prototypes.py:313-336) for hard pseudo-labels.

    ``out``: student logits (B, C, h, w), may require grad; ``pseudolabels``: the (N, 1) / (B, h, w) int64 labels of
    ``prototype_predictions`` (255 = ignore); ``n_valid``: optional 0-dim / 1-element device tensor with the number of
    non-ignored labels (``handler.last_pixel_count()``), which saves a counting launch.  Returns a dict with the
    reference's keys: ``ce_loss``, ``rce_loss``, ``sym_loss``, ``regularization_loss``, ``Total target loss`` (the only
    differentiable entry), ``output & prototype agreement`` and ``pseudolabel_pixel_num``, all 0-dim device tensors.
    """
    total, parts = _TargetLoss.apply(out, pseudolabels, n_valid, rce_alpha, rce_beta, regularizer_weight, regularizer)
    return {
        "ce_loss": parts[0], "rce_loss": parts[1], "sym_loss": rce_alpha * parts[0] + rce_beta * parts[1],
        "regularization_loss": parts[2], "Total target loss": total,
        "output & prototype agreement": parts[4], "pseudolabel_pixel_num": parts[5],
    }
