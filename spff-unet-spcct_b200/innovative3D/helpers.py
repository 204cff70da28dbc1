"""Loss and step metrics of the hot path, on the B200 kernels.

Same call surface as the reference's `innovative3D/helpers.py`:
  ce_plus_macro_dice_loss(logits, labels, num_classes, ignore_index=255, smooth=1e-6) -> Tensor   (:797-803)
  macro_dice_loss(logits, labels, num_classes, ignore_index=255, smooth=1e-6) -> float            (:782-795)
  per_class_metrics_3d / _2d(preds, labels, num_classes, smooth=1e-6, ignore_index=None) -> 9-tuple (:668-779)

The reference evaluates these with ~140 boolean-mask reductions and `.item()` host syncs per step
(SURVEY.md §2.3 K9/K10). Here ONE kernel pass (`spff_ce_confusion`) over logits + labels produces the
sufficient statistics — sum of nll, number of valid voxels, the [label][argmax] tally — and
everything else is derived from those 171 integers; the CE gradient is `spff_ce_grad`. A tally is
cached per (logits, labels) pair so that the loss and the metrics of one step share one pass.
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import math
import weakref
from typing import List, Optional

import numpy as np
import torch

from spff_b200 import ops
from spff_b200.engine import LossTally

_NO_IGNORE = -1   # label value that never occurs: "no ignore_index"


def _prep(logits: torch.Tensor, labels: torch.Tensor):
    if not logits.is_cuda:
        raise RuntimeError("innovative3D.helpers (B200 build) needs CUDA tensors; there is no CPU fallback")
    if logits.dim() == 4:      # 2-D models: [B,K,H,W] is a depth-1 volume
        logits = logits.unsqueeze(2)
        labels = labels.unsqueeze(1) if labels.dim() == 3 else labels
    if logits.dim() != 5:
        raise ValueError(f"logits must be [B,K,D,H,W] or [B,K,H,W], got {tuple(logits.shape)}")
    lg = logits.detach()
    if lg.dtype != torch.float32 or not lg.is_contiguous():
        lg = lg.float().contiguous()
    lb = labels.to(logits.device)
    if lb.dtype not in (torch.uint8, torch.int64):
        lb = lb.long()
    lb = lb.contiguous()
    n, k, d, h, w = lg.shape
    if tuple(lb.shape) != (n, d, h, w):
        raise ValueError(f"labels {tuple(lb.shape)} do not match logits {tuple(lg.shape)}")
    return lg, lb


class _TallyCache:
    """Tally of the most recent (logits, labels) pair, so that the loss and the metrics of one step
    share one kernel pass. A hit needs the SAME live tensor objects at the same version — data_ptr
    alone is not an identity (the caching allocator hands a freed block to the next tensor)."""
    logits_ref = None
    labels_ref = None
    versions = None
    ign = None
    tally: Optional[LossTally] = None


def _tally(logits: torch.Tensor, labels: torch.Tensor, ignore_index: Optional[int]) -> LossTally:
    ign = _NO_IGNORE if ignore_index is None else int(ignore_index)
    c = _TallyCache
    if (c.tally is not None and c.logits_ref is not None and c.logits_ref() is logits and c.labels_ref() is labels
            and c.versions == (logits._version, labels._version) and c.ign == ign):
        return c.tally
    lg, lb = _prep(logits, labels)
    t = LossTally(lg.shape[1], lg.device)
    ops.ce_confusion(lg, lb, ign, t.nll, t.count, t.confusion)
    c.logits_ref, c.labels_ref = weakref.ref(logits), weakref.ref(labels)
    c.versions, c.ign, c.tally = (logits._version, labels._version), ign, t
    return t


class _CrossEntropyFn(torch.autograd.Function):
    """F.cross_entropy(logits, labels, ignore_index) (mean over valid voxels), helpers.py:798-801."""

    @staticmethod
    def forward(ctx, logits, labels, ignore_index, t):
        ctx.save_for_backward(logits, labels, t.count)
        ctx.ignore_index = ignore_index
        return (t.nll[0] / t.count[0].double()).to(torch.float32)   # 0/0 = nan when nothing is valid, as torch

    @staticmethod
    def backward(ctx, g):
        logits, labels, count = ctx.saved_tensors
        lg, lb = _prep(logits, labels)
        dl = torch.empty_like(lg)
        ops.ce_grad(lg, lb, int(ctx.ignore_index), count, g.detach().reshape(1).float().contiguous(), dl)
        return dl.view_as(logits).to(logits.dtype), None, None, None


def masked_ce_loss(logits, labels, ignore_index=255):
    """Sum of the valid voxels' nll over max(#valid, 1): `_weighted_softmax_ce` of the Cicek wrapper with
    class_weights None and no voxel weights (models.py:779-798). Shares the tally pass with the metrics."""
    ign = _NO_IGNORE if ignore_index is None else int(ignore_index)
    t = _tally(logits, labels, ign)
    clamped = LossTally.__new__(LossTally)           # same statistics, denominator clamped like .clamp_min(1.0)
    clamped.k, clamped.nll, clamped.count, clamped.confusion = t.k, t.nll, t.count.clamp(min=1), t.confusion
    return _CrossEntropyFn.apply(logits, labels, ign, clamped)


def _dice_from_tally(t: LossTally, smooth: float) -> torch.Tensor:
    cm = t.confusion.double()
    tp = cm.diagonal()[1:]
    fp = cm.sum(0)[1:] - tp
    fn = cm.sum(1)[1:] - tp
    if tp.numel() == 0:
        return cm.new_ones(())
    return ((2 * tp + smooth) / (2 * tp + fp + fn + smooth)).mean()


def macro_dice_loss(logits, labels, num_classes, ignore_index=255, smooth=1e-6) -> float:
    """1 - hard macro Dice over classes 1..K-1 on argmax(logits); a Python float like the reference's
    (helpers.py:782-795), hence one host sync. ce_plus_macro_dice_loss does not call this."""
    return float(1.0 - _dice_from_tally(_tally(logits, labels, ignore_index), smooth))


def ce_plus_macro_dice_loss(logits, labels, num_classes, ignore_index=255, smooth=1e-6):
    """CE(ignore_index) + 0.5 * (1 - hard macro Dice) (helpers.py:797-803). The Dice term carries no
    gradient in the reference (argmax + .item()); here it stays on the device, so the call does not
    synchronise the host."""
    ign = _NO_IGNORE if ignore_index is None else int(ignore_index)
    t = _tally(logits, labels, ign)        # one pass; per_class_metrics_3d on the same tensors reuses it
    ce = _CrossEntropyFn.apply(logits, labels, ign, t)
    dice = _dice_from_tally(t, smooth)
    return ce + (0.5 * (1.0 - dice)).to(torch.float32)


def metrics_from_confusion(cm: np.ndarray, total_voxels: int, smooth: float = 1e-6):
    """The 9-tuple of per_class_metrics_3d (helpers.py:668-725) from the [label][argmax] tally over
    valid voxels. `total_voxels` counts every voxel, ignored ones included: the reference's true
    negatives are `(~pred_c & ~label_c).sum()` with both masks cleared on ignored voxels (:684-690),
    so ignored voxels are true negatives of every class."""
    k = cm.shape[0]
    row, col, diag = cm.sum(1), cm.sum(0), np.diagonal(cm)
    dice_l: List[float] = []
    sens_l: List[float] = []
    spec_l: List[float] = []
    for c in range(k):
        tp = int(diag[c]); fp = int(col[c] - diag[c]); fn = int(row[c] - diag[c])
        tn = int(total_voxels) - tp - fp - fn
        if (tp + fn) == 0 and fp == 0:          # absent in GT and never predicted (:693-696)
            dice = sens = float("nan")
        else:
            dice = (2 * tp + smooth) / (2 * tp + fp + fn + smooth)
            sens = (tp + smooth) / (tp + fn + smooth) if (tp + fn) > 0 else float("nan")
        spec = (tn + smooth) / (tn + fp + smooth) if (tn + fp) > 0 else float("nan")
        dice_l.append(dice); sens_l.append(sens); spec_l.append(spec)

    def _nanmean_fg(v):
        if k <= 1:
            return float("nan")
        vals = [x for x in v[1:] if not math.isnan(x)]
        return float(np.mean(vals)) if vals else float("nan")

    tp_s = int(diag[1:].sum()); fp_s = int((col[1:] - diag[1:]).sum()); fn_s = int((row[1:] - diag[1:]).sum())
    tn_s = int(cm[0, 0])
    den = 2 * tp_s + fp_s + fn_s
    micro_dice = (2 * tp_s + smooth) / (den + smooth) if den > 0 else float("nan")
    micro_sens = (tp_s + smooth) / (tp_s + fn_s + smooth) if (tp_s + fn_s) > 0 else float("nan")
    micro_spec = (tn_s + smooth) / (tn_s + fp_s + smooth) if (tn_s + fp_s) > 0 else float("nan")
    return (dice_l, sens_l, spec_l, _nanmean_fg(dice_l), _nanmean_fg(sens_l), _nanmean_fg(spec_l),
            micro_dice, micro_sens, micro_spec)


def per_class_metrics_3d(preds, labels, num_classes, smooth=1e-6, ignore_index=None):
    """(dice_list, sens_list, spec_list, macro_dice, macro_sens, macro_spec, micro_dice, micro_sens,
    micro_spec) as helpers.py:668-725 — one small device->host copy instead of 102 `.item()` calls."""
    t = _tally(preds, labels, ignore_index)
    cm = t.confusion.cpu().numpy()
    return metrics_from_confusion(cm, labels.numel(), smooth)


def per_class_metrics_2d(preds, labels, num_classes, smooth=1e-6, ignore_index=None):
    """2-D twin (helpers.py:728-779): same counting rules on [B,K,H,W] logits."""
    return per_class_metrics_3d(preds, labels, num_classes, smooth, ignore_index)


LOSS_REGISTRY = {
    "ce_plus_macro_dice": lambda logits, labels, nc, ignore_index: ce_plus_macro_dice_loss(
        logits, labels, nc, ignore_index=ignore_index),
}


# Everything else of the reference's helpers.py (DICOM ingest, legacy augmentation, visualisers, unused losses:
# helpers.py:43-662, 811-943) is outside the hot path; those names resolve to the reference's module when a reference
# checkout is on sys.path, so `innovative3D/datasets.py:31-34` of the reference imports them from here unchanged.
from ._fallthrough import module_getattr as _module_getattr  # noqa: E402

__getattr__ = _module_getattr(__name__, "helpers", "only the loss / step-metric functions of the hot path are rebuilt here")
