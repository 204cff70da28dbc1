"""CPU oracle for the SPFF-UNet hot path — TEST INFRASTRUCTURE, not product code.

A plain PyTorch fp32 (CPU) restatement of what the reference computes on the path
`BASELINE.json:north_star` names: the depth-preserving 3D encoder-decoder forward (+ autograd
backward) and the CE + hard-macro-Dice loss with its step metrics. It is written functionally over a
`state_dict`-style dict of tensors so that it shares no code with the product (`spff-unet-spcct_b200/`).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import it.

Pinning: the reference ships no tests or golden vectors, so this oracle is pinned
against outputs of the reference code itself: `oracle/make_golden.py` imports `/root/reference`
under a stub shim, runs it on seeded inputs/weights and writes `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks this file against those vectors.

Reference citations are `innovative3D/<file>:<line>` of NF-91/spff-unet-spcct.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

NUM_CLASSES = 13      # config.py:23
NUM_FRAMES = 5        # config.py:22
IGNORE_INDEX = 255    # config.py:26
BEST_LR = 1e-4        # config.py:25

Params = Dict[str, torch.Tensor]

# variant -> (block sub-module names, efilm, fgate, specse+chanse)
VARIANTS = {
    "SPFF-UNet": dict(names=("pre", "body"), efilm=True, fgate=True, se=True),        # models.py:1558-1564
    "E_SP_UNet": dict(names=("pre", "body"), efilm=True, fgate=False, se=True),       # models.py:1565-1573
    "FG_SP_UNet": dict(names=("pre", "body"), efilm=False, fgate=True, se=True),      # models.py:1575-1583
    "PlainCore_UNet": dict(names=("b1", "b2"), efilm=False, fgate=False, se=False),   # models.py:1594-1607
    # LitSPCT_SEspec: plain blocks + SpectralSE + ChannelSE, input replicate-padded to multiples of 16 in
    # (D,H,W) and the logits centre-cropped back (_LitSPCT_Base.forward, models.py:703-712, 1585-1592)
    "SP_UNet": dict(names=("b1", "b2"), efilm=False, fgate=False, se=True, pad=16),
}
BLOCKS = ("enc1", "enc2", "enc3", "bott", "dec3", "dec2", "dec1")


# ------------------------------------------------------------------------------------------------
# parameter surface
# ------------------------------------------------------------------------------------------------
def param_shapes(variant: str = "SPFF-UNet", num_classes: int = NUM_CLASSES, base: int = 32,
                 frames: int = NUM_FRAMES) -> Dict[str, Tuple[int, ...]]:
    """state_dict keys/shapes of the Lit module (`model.` prefix), as built by
    UNet3D_SpectralCore.__init__ (models.py:654-682) + upgrade_spct_with_novel_blocks (:1416-1446).
    `freq_mask` is the lazily registered FourierGate parameter (:1532-1535), rfft length frames//2+1."""
    v = VARIANTS[variant]
    n1, n2 = v["names"]
    f = base
    chans = {"enc1": (1, f), "enc2": (f, 2 * f), "enc3": (2 * f, 4 * f), "bott": (4 * f, 8 * f),
             "dec3": (8 * f, 4 * f), "dec2": (4 * f, 2 * f), "dec1": (2 * f, f)}
    s: Dict[str, Tuple[int, ...]] = {}
    for b in BLOCKS:
        ci, co = chans[b]
        s[f"model.{b}.{n1}.0.weight"] = (co, ci, 3, 3, 3)
        s[f"model.{b}.{n1}.1.weight"] = (co,)
        s[f"model.{b}.{n1}.1.bias"] = (co,)
        s[f"model.{b}.{n2}.0.weight"] = (co, co, 3, 3, 3)
        s[f"model.{b}.{n2}.1.weight"] = (co,)
        s[f"model.{b}.{n2}.1.bias"] = (co,)
        if v["efilm"]:  # EnergyFiLM3D(channels, hidden=32, pe_dims=16)  models.py:1484-1492
            s[f"model.{b}.efilm.mlp.0.weight"] = (32, 16, 1)
            s[f"model.{b}.efilm.mlp.0.bias"] = (32,)
            s[f"model.{b}.efilm.mlp.2.weight"] = (2 * co, 32, 1)
            s[f"model.{b}.efilm.mlp.2.bias"] = (2 * co,)
        if v["fgate"]:  # FourierGate3D  models.py:1521-1535
            s[f"model.{b}.fgate.mag_scale"] = (1,)
            s[f"model.{b}.fgate.freq_mask"] = (1, 1, frames // 2 + 1, 1, 1)
    for name, (ci, co) in (("up3", (8 * f, 4 * f)), ("up2", (4 * f, 2 * f)), ("up1", (2 * f, f))):
        s[f"model.{name}.weight"] = (ci, co, 1, 2, 2)   # nn.ConvTranspose3d  models.py:668-672
        s[f"model.{name}.bias"] = (co,)
    s["model.out.weight"] = (num_classes, f, 1, 1, 1)   # models.py:674
    s["model.out.bias"] = (num_classes,)
    if v["se"]:  # _SEChannelLite(c, r=16), hidden max(4, c//16)   models.py:600-609
        for i, c in enumerate((f, 2 * f, 4 * f, 8 * f)):
            h = max(4, c // 16)
            s[f"model.se.{i}.fc.0.weight"] = (h, c, 1, 1, 1)
            s[f"model.se.{i}.fc.0.bias"] = (h,)
            s[f"model.se.{i}.fc.2.weight"] = (c, h, 1, 1, 1)
            s[f"model.se.{i}.fc.2.bias"] = (c,)
    return s


def det_weights(shapes: Dict[str, Tuple[int, ...]], seed: int = 42) -> Params:
    """Deterministic, name-keyed weights (independent of module construction order), used by the
    golden generator and by every parity test so that weights never have to be stored."""
    out: Params = {}
    for name in sorted(shapes):
        shp = shapes[name]
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        r = torch.randn(shp, generator=g, dtype=torch.float32)
        if name.endswith("mag_scale") or name.endswith("freq_mask"):
            t = 1.0 + 0.2 * r
        elif len(shp) == 1 and name.endswith("weight"):      # norm gamma
            t = 1.0 + 0.1 * r
        elif name.endswith("bias"):
            t = 0.1 * r
        else:                                                 # conv / linear weights: fan-in scaling
            fan_in = int(np.prod(shp[1:])) if len(shp) > 1 else 1
            if ".up" in name:                                 # ConvTranspose3d [Cin,Cout,1,2,2]
                fan_in = shp[0]
            t = r * (1.5 / math.sqrt(fan_in))
        out[name] = t
    return out


# ------------------------------------------------------------------------------------------------
# forward
# ------------------------------------------------------------------------------------------------
def _conv_norm_act(p: Params, pre: str, x: torch.Tensor) -> torch.Tensor:
    """_conv3x3xk + InstanceNorm3d(affine, eps 1e-5) + LeakyReLU(0.01)  (models.py:616-618,168-181)."""
    x = F.conv3d(x, p[f"{pre}.0.weight"], None, padding=1)
    x = F.instance_norm(x, weight=p[f"{pre}.1.weight"], bias=p[f"{pre}.1.bias"], eps=1e-5)
    return F.leaky_relu(x, 0.01)


def sinusoidal_pe(frames: int, d: int) -> torch.Tensor:
    """EnergyFiLM3D._sinusoidal_pe (models.py:1494-1503): [1, d, F]."""
    pos = torch.arange(frames, dtype=torch.float32)[None, None, :]
    i = torch.arange(max(1, d // 2), dtype=torch.float32)[None, :, None]
    denom = torch.exp(i * (-math.log(10000.0) / max(1, d // 2)))
    pe = torch.cat([torch.sin(pos * denom), torch.cos(pos * denom)], dim=1)
    if pe.shape[1] < d:
        pe = torch.cat([pe, torch.zeros(1, 1, pe.shape[-1])], dim=1)
    return pe


def efilm_tables(p: Params, pre: str, c: int, frames: int):
    """(1 + tanh(gamma))[C,F] and beta[C,F] of EnergyFiLM3D.forward (models.py:1505-1512); they
    depend on the parameters only (the positional code is a constant)."""
    pe = sinusoidal_pe(frames, 16)
    h = F.relu(F.conv1d(pe, p[f"{pre}.mlp.0.weight"], p[f"{pre}.mlp.0.bias"]))
    gb = F.conv1d(h, p[f"{pre}.mlp.2.weight"], p[f"{pre}.mlp.2.bias"])[0]   # [2C, F]
    return 1.0 + torch.tanh(gb[:c]), gb[c:]


def _efilm(p: Params, pre: str, x: torch.Tensor) -> torch.Tensor:
    c, frames = x.shape[1], x.shape[2]
    g1, beta = efilm_tables(p, pre, c, frames)
    return x * g1[None, :, :, None, None] + beta[None, :, :, None, None]


def _fgate(p: Params, pre: str, x: torch.Tensor) -> torch.Tensor:
    """FourierGate3D.forward with learn_phase=False (models.py:1527-1544)."""
    frames = x.shape[2]
    s = x.mean(dim=(1, 3, 4), keepdim=True)
    sf = torch.fft.rfft(s, dim=2)
    sf = sf * (p[f"{pre}.freq_mask"] * p[f"{pre}.mag_scale"])
    w = torch.fft.irfft(sf, n=frames, dim=2)
    return x * torch.sigmoid(w)


def _spectral_se(x: torch.Tensor) -> torch.Tensor:
    """_SpectralSE (models.py:611-614)."""
    return x * torch.sigmoid(x.mean(dim=(1, 3, 4), keepdim=True))


def _channel_se(p: Params, pre: str, x: torch.Tensor) -> torch.Tensor:
    """_SEChannelLite (models.py:600-609): x * sigmoid(fc2(relu(fc1(avgpool(x)))))."""
    s = x.mean(dim=(2, 3, 4), keepdim=True)
    h = F.relu(F.conv3d(s, p[f"{pre}.fc.0.weight"], p[f"{pre}.fc.0.bias"]))
    return x * torch.sigmoid(F.conv3d(h, p[f"{pre}.fc.2.weight"], p[f"{pre}.fc.2.bias"]))


def pad_to_mult(x: torch.Tensor, m: int):
    """_pad_to_mult_3d (models.py:109-120): replicate-pad [B,C,D,H,W] so that D, H, W are multiples of m,
    the odd element of an uneven split on the right. Returns (x_pad, (D,H,W)) or (x, None)."""
    _, _, d, h, w = x.shape
    pd, ph, pw = (-d) % m, (-h) % m, (-w) % m
    if not (pd or ph or pw):
        return x, None
    x = F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2, pd // 2, pd - pd // 2), mode="replicate")
    return x, (d, h, w)


def center_crop(x: torch.Tensor, orig):
    """_center_crop_to_3d (models.py:122-127)."""
    if orig is None:
        return x
    d, h, w = orig
    sd, sh, sw = (x.shape[2] - d) // 2, (x.shape[3] - h) // 2, (x.shape[4] - w) // 2
    return x[:, :, sd:sd + d, sh:sh + h, sw:sw + w]


def unet_forward(p: Params, x: torch.Tensor, variant: str = "SPFF-UNet", taps: Dict[str, torch.Tensor] | None = None
                 ) -> torch.Tensor:
    """Variant forward: the core (below), wrapped in pad / crop for the variants that pad (SP_UNet)."""
    m = VARIANTS[variant].get("pad")
    if not m:
        return core_forward(p, x, variant, taps)
    xp, orig = pad_to_mult(x, m)
    return center_crop(core_forward(p, xp, variant, taps), orig)


def core_forward(p: Params, x: torch.Tensor, variant: str = "SPFF-UNet", taps: Dict[str, torch.Tensor] | None = None
                 ) -> torch.Tensor:
    """UNet3D_SpectralCore.forward (models.py:693-701) with the blocks of the given variant:
    _DoubleConvSpectral_Novel.forward (:1473-1478) or _DoubleConvSpectral (:620-625), `_post`
    = SpectralSE -> ChannelSE on the encoder/bottleneck outputs only (:684-685), MaxPool3d((1,2,2))
    (:658-665), ConvTranspose3d k=s=(1,2,2) (:668-672), cat([up, skip]) (:687-691), 1x1x1 head (:674).
    x: [B,1,F,H,W] fp32 -> logits [B,num_classes,F,H,W]. `taps` (optional) collects block outputs."""
    v = VARIANTS[variant]
    n1, n2 = v["names"]

    def block(name: str, t: torch.Tensor, stage: int | None) -> torch.Tensor:
        t = _conv_norm_act(p, f"model.{name}.{n1}", t)
        t = _conv_norm_act(p, f"model.{name}.{n2}", t)
        if v["efilm"]:
            t = _efilm(p, f"model.{name}.efilm", t)
        if v["fgate"]:
            t = _fgate(p, f"model.{name}.fgate", t)
        if stage is not None and v["se"]:
            t = _channel_se(p, f"model.se.{stage}", _spectral_se(t))
        if taps is not None:
            taps[name] = t
        return t

    def up(name: str, t: torch.Tensor) -> torch.Tensor:
        return F.conv_transpose3d(t, p[f"model.{name}.weight"], p[f"model.{name}.bias"], stride=(1, 2, 2))

    pool = lambda t: F.max_pool3d(t, (1, 2, 2))
    e1 = block("enc1", x, 0)
    e2 = block("enc2", pool(e1), 1)
    e3 = block("enc3", pool(e2), 2)
    b = block("bott", pool(e3), 3)
    d3 = block("dec3", torch.cat([up("up3", b), e3], 1), None)
    d2 = block("dec2", torch.cat([up("up2", d3), e2], 1), None)
    d1 = block("dec1", torch.cat([up("up1", d2), e1], 1), None)
    return F.conv3d(d1, p["model.out.weight"], p["model.out.bias"])


# ------------------------------------------------------------------------------------------------
# loss + metrics
# ------------------------------------------------------------------------------------------------
def confusion(pred: torch.Tensor, labels: torch.Tensor, num_classes: int, ignore_index: int | None) -> np.ndarray:
    """int64 [label][pred] tally over valid voxels — the sufficient statistic of every count the
    reference takes with `.sum().item()` (helpers.py:687-690, 716-719, 789-791)."""
    pred = pred.reshape(-1).long()
    lab = labels.reshape(-1).long()
    if ignore_index is not None:
        m = lab != ignore_index
        pred, lab = pred[m], lab[m]
    ok = (lab >= 0) & (lab < num_classes)
    idx = lab[ok] * num_classes + pred[ok]
    return torch.bincount(idx, minlength=num_classes * num_classes).reshape(num_classes, num_classes).numpy()


def macro_dice_from_confusion(cm: np.ndarray, smooth: float = 1e-6) -> float:
    """macro_dice_loss's dice (helpers.py:782-795): classes 1..K-1, absent class -> (0+s)/(0+s) = 1."""
    k = cm.shape[0]
    dl = []
    for c in range(1, k):
        tp = int(cm[c, c]); fp = int(cm[:, c].sum() - tp); fn = int(cm[c, :].sum() - tp)
        dl.append((2 * tp + smooth) / (2 * tp + fp + fn + smooth))
    return float(np.mean(dl)) if dl else 1.0


def ce_plus_macro_dice_loss(logits: torch.Tensor, labels: torch.Tensor, num_classes: int = NUM_CLASSES,
                            ignore_index: int = IGNORE_INDEX, smooth: float = 1e-6) -> torch.Tensor:
    """helpers.py:797-803: F.cross_entropy(ignore_index) + 0.5 * (1 - hard macro dice); the dice term
    is a Python float (no gradient)."""
    ce = F.cross_entropy(logits, labels.long(), ignore_index=ignore_index)
    cm = confusion(torch.argmax(logits, dim=1), labels, num_classes, ignore_index)
    return ce + 0.5 * (1.0 - macro_dice_from_confusion(cm, smooth))


def metrics_from_confusion(cm: np.ndarray, total_voxels: int | None = None, smooth: float = 1e-6):
    """per_class_metrics_3d (helpers.py:668-725) from the [label][pred] tally: the 9-tuple
    (dice_list, sens_list, spec_list, macro_dice, macro_sens, macro_spec, micro_dice, micro_sens,
    micro_spec) with the reference's NaN rules (:693-701) and nanmean over classes 1.. (:708-710).
    `total_voxels` = ALL voxels, ignored ones included: the reference's tn is
    `(~pred_c & ~label_c).sum()` where both masks are already cleared on ignored voxels (:684-690),
    so an ignored voxel is a true negative of every class (pinned by tests/golden case0/case4)."""
    import warnings

    k = cm.shape[0]
    total = int(cm.sum()) if total_voxels is None else int(total_voxels)
    dice_l: List[float] = []; sens_l: List[float] = []; spec_l: List[float] = []
    for c in range(k):
        tp = int(cm[c, c]); fp = int(cm[:, c].sum() - tp); fn = int(cm[c, :].sum() - tp)
        tn = total - tp - fp - fn
        gt_present = (tp + fn) > 0
        if (not gt_present) and fp == 0:
            dice = float("nan"); sens = float("nan")
        else:
            dice = (2 * tp + smooth) / (2 * tp + fp + fn + smooth)
            sens = (tp + smooth) / (tp + fn + smooth) if (tp + fn) > 0 else float("nan")
        spec = (tn + smooth) / (tn + fp + smooth) if (tn + fp) > 0 else float("nan")
        dice_l.append(dice); sens_l.append(sens); spec_l.append(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        macro = [float(np.nanmean(l[1:])) if k > 1 else float("nan") for l in (dice_l, sens_l, spec_l)]
    tp_s = sum(int(cm[c, c]) for c in range(1, k))
    fp_s = sum(int(cm[:, c].sum() - cm[c, c]) for c in range(1, k))
    fn_s = sum(int(cm[c, :].sum() - cm[c, c]) for c in range(1, k))
    tn_s = int(cm[0, 0])
    den = 2 * tp_s + fp_s + fn_s
    micro_dice = (2 * tp_s + smooth) / (den + smooth) if den > 0 else float("nan")
    micro_sens = (tp_s + smooth) / (tp_s + fn_s + smooth) if (tp_s + fn_s) > 0 else float("nan")
    micro_spec = (tn_s + smooth) / (tn_s + fp_s + smooth) if (tn_s + fp_s) > 0 else float("nan")
    return (dice_l, sens_l, spec_l, macro[0], macro[1], macro[2], micro_dice, micro_sens, micro_spec)


def per_class_metrics_3d(logits, labels, num_classes, smooth=1e-6, ignore_index=None):
    cm = confusion(torch.argmax(logits, dim=1), labels, num_classes, ignore_index)
    return metrics_from_confusion(cm, labels.numel(), smooth)


# ------------------------------------------------------------------------------------------------
# synthetic data and a short pre-training loop
# ------------------------------------------------------------------------------------------------
def phantom_batch(b: int, h: int, w: int, seed: int, frames: int = NUM_FRAMES, num_classes: int = NUM_CLASSES,
                  ignore_frac: float = 0.0):
    """Phantom-like slices: background + elliptical inserts of constant class with a per-class
    5-bin signature + N(0, 0.1^2) noise; labels from the same ellipses (mirrors the ROI painting
    of helpers.py:177-209). Returns x [b,1,F,h,w] fp32, labels [b,F,h,w] int64."""
    rng = np.random.RandomState(seed)
    sig = np.random.RandomState(1234).uniform(0.2, 1.5, size=(num_classes, frames)).astype(np.float32)
    sig[0] = 0.05
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    x = np.zeros((b, 1, frames, h, w), np.float32)
    lab = np.zeros((b, frames, h, w), np.int64)
    for i in range(b):
        lm = np.zeros((h, w), np.int64)
        for _ in range(8):
            c = rng.randint(1, num_classes)
            cy, cx = rng.uniform(0.15, 0.85) * h, rng.uniform(0.15, 0.85) * w
            ry, rx = rng.uniform(0.06, 0.16) * h, rng.uniform(0.06, 0.16) * w
            lm[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0] = c
        lab[i] = lm[None]
        x[i, 0] = sig[lm].transpose(2, 0, 1)
    x += rng.normal(0, 0.1, size=x.shape).astype(np.float32)
    if ignore_frac > 0:
        lab[rng.uniform(size=lab.shape) < ignore_frac] = IGNORE_INDEX
    return torch.from_numpy(x), torch.from_numpy(lab)


def loss_and_grads(p: Params, x: torch.Tensor, labels: torch.Tensor, variant: str = "SPFF-UNet"):
    """One fwd + loss + backward: returns (loss float, logits, {name: grad})."""
    q = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    logits = unet_forward(q, x, variant)
    loss = ce_plus_macro_dice_loss(logits, labels, logits.shape[1])
    loss.backward()
    return float(loss.detach()), logits.detach(), {k: (v.grad if v.grad is not None else torch.zeros_like(v))
                                                   for k, v in q.items()}


def pretrain(p: Params, variant: str, steps: int, b: int, h: int, w: int, lr: float = 1e-3, seed: int = 7) -> Params:
    """`steps` Adam steps on phantom batches so that logits have real margins (argmax agreement and
    Dice tolerances are meaningless on random-init logits)."""
    q = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    opt = torch.optim.Adam(list(q.values()), lr=lr)
    for i in range(steps):
        x, lab = phantom_batch(b, h, w, seed + i)
        opt.zero_grad(set_to_none=True)
        loss = ce_plus_macro_dice_loss(unet_forward(q, x, variant), lab)
        loss.backward()
        opt.step()
    return {k: v.detach() for k, v in q.items()}
