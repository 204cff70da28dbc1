"""CPU oracle for the "3DUNet" control (Cicek 3D U-Net + depth adapter) — TEST INFRASTRUCTURE, not product code.

Plain PyTorch fp32 (CPU) restatement, functional over a `state_dict`-style dict, of
  Cicek3DUNet                                  reference innovative3D/models.py:718-751
  LitCicek3DUNet_DepthAdapter_Published        models.py:753-846 (forward :773-777, CE :779-798, SGD :844-846)
  _resize_depth_like / _resize_logits_depth_like   models.py:153-163
as configured by `make_cicek_depth_adapter_sgd_wce` (config.py:283-303): 13 classes, target_depth 16,
BatchNorm, plain CE (class_weights None, dice_weight 0), ignore_index 255, SGD(lr 1e-2, momentum 0.99).

Pinned against the reference itself: `oracle/make_golden_3dunet.py` imports /root/reference and writes
`tests/golden/cicek*.npz`; `tests/test_oracle_golden.py` holds this file to those vectors.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from .spff_oracle import IGNORE_INDEX, NUM_CLASSES, det_weights as _det_weights

Params = Dict[str, torch.Tensor]
TARGET_DEPTH = 16          # config.py:300
BN_EPS, BN_MOMENTUM = 1e-5, 0.1   # nn.BatchNorm3d defaults (models.py:721)
ENC = ("enc1", "enc2", "enc3", "enc4", "bott")
DEC = ("dec4", "dec3", "dec2", "dec1")


def block_channels(base: int = 32) -> Dict[str, Tuple[int, int]]:
    """(cin, cout) of every double-conv block (models.py:728-740)."""
    f = base
    return {"enc1": (1, f), "enc2": (f, 2 * f), "enc3": (2 * f, 4 * f), "enc4": (4 * f, 8 * f), "bott": (8 * f, 16 * f),
            "dec4": (16 * f, 8 * f), "dec3": (8 * f, 4 * f), "dec2": (4 * f, 2 * f), "dec1": (2 * f, f)}


def param_shapes(num_classes: int = NUM_CLASSES, base: int = 32) -> Dict[str, Tuple[int, ...]]:
    """state_dict keys / shapes of LitCicek3DUNet_DepthAdapter_Published (prefix `backbone.`): per block
    Sequential(conv, BN, ReLU, conv, BN, ReLU) -> indices 0,1,3,4 (models.py:722-726); buffers included."""
    s: Dict[str, Tuple[int, ...]] = {}
    ch = block_channels(base)

    def block(name):
        ci, co = ch[name]
        for conv, bn, cin in ((0, 1, ci), (3, 4, co)):
            s[f"backbone.{name}.{conv}.weight"] = (co, cin, 3, 3, 3)
            s[f"backbone.{name}.{bn}.weight"] = (co,)
            s[f"backbone.{name}.{bn}.bias"] = (co,)
            s[f"backbone.{name}.{bn}.running_mean"] = (co,)
            s[f"backbone.{name}.{bn}.running_var"] = (co,)
            s[f"backbone.{name}.{bn}.num_batches_tracked"] = ()

    for b in ENC:
        block(b)
    for up, dec in (("up4", "dec4"), ("up3", "dec3"), ("up2", "dec2"), ("up1", "dec1")):
        co = ch[dec][1]
        s[f"backbone.{up}.weight"] = (2 * co, co, 2, 2, 2)   # nn.ConvTranspose3d(2co, co, 2, stride=2)
        s[f"backbone.{up}.bias"] = (co,)
        block(dec)
    s["backbone.out.weight"] = (num_classes, base, 1, 1, 1)
    s["backbone.out.bias"] = (num_classes,)
    return s


def is_buffer(name: str) -> bool:
    return name.endswith(("running_mean", "running_var", "num_batches_tracked"))


def det_weights(seed: int = 42, num_classes: int = NUM_CLASSES) -> Params:
    """Name-seeded weights as for the SPCT family; BatchNorm buffers start at their constructor values
    perturbed deterministically (running_var kept positive) so that eval-mode parity means something."""
    shapes = param_shapes(num_classes)
    w = _det_weights({k: v for k, v in shapes.items() if not k.endswith("num_batches_tracked")}, seed)
    for k in shapes:
        if k.endswith("running_var"):
            w[k] = 0.5 + w[k].abs()          # positive
        elif k.endswith("running_mean"):
            w[k] = 0.1 * w[k] / max(1e-6, float(w[k].abs().max()))
        elif k.endswith("num_batches_tracked"):
            w[k] = torch.zeros((), dtype=torch.long)
    return w


def resize_depth(x: torch.Tensor, depth: int) -> torch.Tensor:
    """models.py:153-163: trilinear resize of D only (align_corners=False)."""
    if x.shape[2] == depth:
        return x
    return F.interpolate(x, size=(depth, x.shape[3], x.shape[4]), mode="trilinear", align_corners=False)


def depth_matrix(din: int, dout: int) -> torch.Tensor:
    """[dout, din] matrix of that resize (H, W unchanged -> linear interpolation along D only), obtained by
    pushing the identity through F.interpolate itself."""
    eye = torch.eye(din).reshape(din, 1, din, 1, 1)
    return resize_depth(eye, dout).reshape(din, dout).t().contiguous()


def backbone_forward(p: Params, x: torch.Tensor, training: bool = True, new_stats: Dict[str, torch.Tensor] | None = None):
    """Cicek3DUNet.forward (models.py:741-751). x [B,1,D,H,W] with D, H, W multiples of 16.
    training=True: batch statistics (and, if `new_stats` is given, the updated running buffers are written
    into it, as nn.BatchNorm3d does in place)."""

    def cna(pre_conv: str, pre_bn: str, t: torch.Tensor) -> torch.Tensor:
        t = F.conv3d(t, p[f"{pre_conv}.weight"], None, padding=1)
        rm, rv = p[f"{pre_bn}.running_mean"].clone(), p[f"{pre_bn}.running_var"].clone()
        t = F.batch_norm(t, rm, rv, p[f"{pre_bn}.weight"], p[f"{pre_bn}.bias"], training, BN_MOMENTUM, BN_EPS)
        if training and new_stats is not None:
            new_stats[f"{pre_bn}.running_mean"], new_stats[f"{pre_bn}.running_var"] = rm, rv
        return F.relu(t)

    def block(name: str, t: torch.Tensor) -> torch.Tensor:
        t = cna(f"backbone.{name}.0", f"backbone.{name}.1", t)
        return cna(f"backbone.{name}.3", f"backbone.{name}.4", t)

    def up(name: str, t: torch.Tensor) -> torch.Tensor:
        return F.conv_transpose3d(t, p[f"backbone.{name}.weight"], p[f"backbone.{name}.bias"], stride=2)

    e1 = block("enc1", x)
    e2 = block("enc2", F.max_pool3d(e1, 2))
    e3 = block("enc3", F.max_pool3d(e2, 2))
    e4 = block("enc4", F.max_pool3d(e3, 2))
    b = block("bott", F.max_pool3d(e4, 2))
    d4 = block("dec4", torch.cat([up("up4", b), e4], 1))
    d3 = block("dec3", torch.cat([up("up3", d4), e3], 1))
    d2 = block("dec2", torch.cat([up("up2", d3), e2], 1))
    d1 = block("dec1", torch.cat([up("up1", d2), e1], 1))
    return F.conv3d(d1, p["backbone.out.weight"], p["backbone.out.bias"])


def forward(p: Params, x: torch.Tensor, training: bool = True, new_stats=None, target_depth: int = TARGET_DEPTH):
    """LitCicek3DUNet_DepthAdapter_Published.forward (models.py:773-777)."""
    d0 = x.shape[2]
    return resize_depth(backbone_forward(p, resize_depth(x, target_depth), training, new_stats), d0)


def ce_loss(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = IGNORE_INDEX) -> torch.Tensor:
    """_weighted_softmax_ce with class_weights None, no voxel weights (models.py:779-798): sum of the valid
    voxels' nll over max(#valid, 1); `_loss_and_log` with ce_weight 1, dice_weight 0 (:817-821)."""
    ce = F.cross_entropy(logits, labels.long(), ignore_index=ignore_index, reduction="none")
    valid = (labels != ignore_index).float()
    return (ce * valid).sum() / valid.sum().clamp_min(1.0)


def loss_and_grads(p: Params, x: torch.Tensor, labels: torch.Tensor):
    """One training-mode fwd + CE + backward: (loss, logits, {name: grad}, {buffer: new value})."""
    q = {k: (v.detach().clone().requires_grad_(True) if not is_buffer(k) else v.detach().clone()) for k, v in p.items()}
    stats: Dict[str, torch.Tensor] = {}
    logits = forward(q, x, True, stats)
    loss = ce_loss(logits, labels)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in q.items() if not is_buffer(k)}
    return float(loss.detach()), logits.detach(), grads, stats


def sgd_step(p: Params, grads: Params, bufs: Params | None, lr: float = 1e-2, momentum: float = 0.99):
    """torch.optim.SGD(momentum, dampening 0) (models.py:844-846): returns (new params, new momentum buffers)."""
    new_p, new_b = dict(p), {}
    for k, g in grads.items():
        b = g.clone() if bufs is None else momentum * bufs[k] + g
        new_b[k] = b
        new_p[k] = p[k] - lr * b
    return new_p, new_b
