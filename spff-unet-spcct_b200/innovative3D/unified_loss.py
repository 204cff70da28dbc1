"""`apply_unified_loss()` — patch every Lightning module of `innovative3D.models` to the unified
ce_plus_macro_dice loss + metrics step (reference innovative3D/unified_loss.py:29-144; dormant in the
reference's train.py:31-32,673-674). The loss and the metrics are the kernel-backed ones of
`innovative3D.helpers`."""
from __future__ import annotations

from . import models as models_mod
from .config import IGNORE_INDEX, NUM_CLASSES
from .helpers import ce_plus_macro_dice_loss, per_class_metrics_2d, per_class_metrics_3d
from .models import _canonicalize_targets_2d, _canonicalize_targets_3d, _pick_first_if_seq, pl


def _get_num_classes(self) -> int:
    return int(getattr(getattr(self, "hparams", object()), "num_classes", NUM_CLASSES))


def _unified_shared_step(self, batch, stage: str):
    """stage in {'train','val','test'}; 3-D [B,K,D,H,W] or 2-D [B,K,H,W] logits; a tuple/list of
    outputs (deep supervision) contributes its main head (unified_loss.py:49-50)."""
    if isinstance(batch, dict):
        imgs, lbls = batch.get("image"), batch.get("label")
    else:
        imgs, lbls = batch
    imgs, lbls = _pick_first_if_seq(imgs), _pick_first_if_seq(lbls)
    logits = self(imgs)
    if isinstance(logits, (list, tuple)):
        logits = logits[0]
    nc = _get_num_classes(self)
    ign = int(getattr(getattr(self, "hparams", object()), "ignore_index", IGNORE_INDEX))
    if logits.ndim == 5:
        tgt = _canonicalize_targets_3d(lbls).to(logits.device).long()
        metrics_fn = per_class_metrics_3d
    elif logits.ndim == 4:
        tgt = _canonicalize_targets_2d(lbls).to(logits.device).long()
        metrics_fn = per_class_metrics_2d
    else:
        raise RuntimeError(f"Unexpected logits ndim {logits.ndim}; expected 4D or 5D.")
    loss = ce_plus_macro_dice_loss(logits, tgt, nc, ignore_index=ign)
    (_, _, _, macro_dice, macro_sens, macro_spec, micro_dice, micro_sens, micro_spec) = metrics_fn(
        logits, tgt, nc, ignore_index=ign)
    kw = dict(on_step=False, on_epoch=True, sync_dist=True)
    self.log(f"{stage}_loss", loss, prog_bar=(stage == "train"), **kw)
    self.log(f"{stage}_macro_dice", macro_dice, prog_bar=(stage != "test"), **kw)
    for name, v in (("micro_dice", micro_dice), ("macro_sens", macro_sens), ("macro_spec", macro_spec),
                    ("micro_sens", micro_sens), ("micro_spec", micro_spec)):
        self.log(f"{stage}_{name}", v, prog_bar=True, **kw)
    return loss


def _training_step(self, batch, batch_idx):
    return _unified_shared_step(self, batch, "train")


def _validation_step(self, batch, batch_idx):
    return _unified_shared_step(self, batch, "val")


def _test_step(self, batch, batch_idx):
    return _unified_shared_step(self, batch, "test")


def apply_unified_loss():
    """Monkey-patch the train/val/test steps of every LightningModule class in innovative3D.models
    except BaseLitModel (unified_loss.py:114-144). Returns the patched class names."""
    patched = []
    for name in dir(models_mod):
        obj = getattr(models_mod, name)
        if not isinstance(obj, type) or not issubclass(obj, pl.LightningModule) or name == "BaseLitModel":
            continue
        obj.training_step, obj.validation_step, obj.test_step = _training_step, _validation_step, _test_step
        patched.append(name)
    return sorted(patched)
