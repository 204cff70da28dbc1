"""Generate tests/golden/trained_*.npz: SPFF-UNet TRAINED BY THE REFERENCE, and the reference's own outputs on it.

TEST INFRASTRUCTURE. Run in the build container only (`python oracle/make_golden_trained.py`, ~3 min); needs
/root/reference. Everything numeric below is produced by the reference's code (innovative3D/models.py, helpers.py)
on CPU fp32 under the inert stubs of oracle/make_golden.py; nothing of its arithmetic is stubbed.

Protocol:
  1. `torch.manual_seed(42)`; `LitSPCT_EFiLM_FourierGate()` (reference); `configure_optimizers()` BEFORE the first
     forward, exactly as Lightning would (so the lazily registered FourierGate masks stay out of Adam,
     models.py:1532-1535 + :591-594); lr set to 1e-3.
  2. 60 steps of the reference's own `training_step` + `backward` + `Adam.step` on phantom batches
     `phantom_batch(8, 32, 32, seed=7+i)`. The loss of every step and a strided sample of the weights after 5 steps
     are kept: they pin the repo's fused step (forward + loss + backward + Adam) against the reference's.
  3. The trained weights are stored as an int8-quantised DELTA to the seed-42 initialisation (which the repo's
     constructors reproduce bit for bit, tests/golden/init_seed42.npz): W = W0 + scale * q. The reference then LOADS
     that W, so both sides of every parity test hold identical weights.
  4. On W and BASELINE.json configs[0] — `phantom_batch(128, 64, 64, seed=42, ignore_frac=0.01)` = x[128,1,5,64,64] —
     the reference's loss, metrics 9-tuple, arg-max map, strided samples + norms of the logits, of every block output
     ("per-layer activations"), of the gradient w.r.t. every block output, and of every parameter gradient.
  5. Yardstick for what bf16 storage costs the *reference itself*: the same quantities on the first 2 slices under
     `torch.autocast("cpu", dtype=torch.bfloat16)` vs fp32, as per-tensor rel-L2 (encoder block-output gradients also
     after summing each 2x2 pooling window: a max-pool arg-max that flips under rounding moves a gradient inside its
     window, which no bf16 implementation can avoid).
"""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import spff_oracle as O  # noqa: E402
from oracle.make_golden import import_reference  # noqa: E402

GOLD = ROOT / "tests" / "golden"
BLOCKS = ("enc1", "enc2", "enc3", "bott", "dec3", "dec2", "dec1")
TRAIN_STEPS, TRAIN_LR, EARLY = 60, 1e-3, 5
NSAMPLE = 4096
YARD_SLICES = 2


def strided(t: torch.Tensor, n: int = NSAMPLE) -> np.ndarray:
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].float().clone().numpy()     # a copy: the source may be a live parameter


def pub(name: str) -> str:
    return name.replace("fgate._mask", "fgate.freq_mask")


def win_sum(g: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.avg_pool3d(g, (1, 2, 2)) * 4.0


def run_with_taps(lit, x, lab, autocast: bool = False):
    """forward + loss + backward of the reference module with hooks on the 7 block outputs (after `_post` for the
    encoder stages = what the skip connection and the pool consume). Returns loss, logits, taps, tap grads, param grads."""
    core = lit.model
    taps = {}
    hooks = []

    def mk(name):
        def hook(_m, _inp, out):
            taps[name] = out
            out.retain_grad()
        return hook

    # encoder outputs are post-processed by core._post(i, .) (models.py:684-685): hook the last module of that chain
    post_last = {}
    for i, b in enumerate(("enc1", "enc2", "enc3", "bott")):
        post_last[b] = core.se[i]          # _post = se[i](sp[i](x)); sa is Identity for these variants
    for b in BLOCKS:
        m = post_last.get(b, getattr(core, b))
        hooks.append(m.register_forward_hook(mk(b)))
    lit.zero_grad(set_to_none=True)
    ctx = torch.autocast("cpu", dtype=torch.bfloat16) if autocast else torch.autocast("cpu", enabled=False)
    with ctx:
        logits = lit(x)
        loss = lit.compute_loss(logits.float(), lab)
    loss.backward()
    for h in hooks:
        h.remove()
    pg = {pub(k): p.grad.detach().float().clone() for k, p in lit.named_parameters() if p.grad is not None}
    return float(loss), logits.detach().float(), {k: v.detach().float() for k, v in taps.items()}, \
        {k: v.grad.detach().float() for k, v in taps.items()}, pg


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def main():
    M, H = import_reference()
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.time()

    # ---- 1/2: the reference trains itself ------------------------------------------------------------------------
    torch.manual_seed(42)
    lit = M.LitSPCT_EFiLM_FourierGate()
    w0 = {k: v.detach().clone() for k, v in lit.state_dict().items()}
    opt = lit.configure_optimizers()["optimizer"]
    for g in opt.param_groups:
        g["lr"] = TRAIN_LR
    losses, early = [], None
    lit.train()
    for i in range(TRAIN_STEPS):
        x, lab = O.phantom_batch(8, 32, 32, seed=7 + i)
        opt.zero_grad(set_to_none=True)
        loss = lit.training_step((x, lab), i)
        loss.backward()
        opt.step()
        losses.append(float(loss))
        if i + 1 == EARLY:
            early = {pub(k): strided(v) for k, v in lit.state_dict().items() if not k.endswith("fgate._mask")}
    print(f"reference training: loss {losses[0]:.4f} -> {losses[-1]:.4f} ({time.time() - t0:.0f}s)")

    # ---- 3: quantised delta, loaded back into the reference ------------------------------------------------------
    sd = lit.state_dict()
    out = dict(train_losses=np.array(losses, np.float64), train_steps=np.int64(TRAIN_STEPS), train_lr=np.float64(TRAIN_LR),
               early_steps=np.int64(EARLY))
    for k, v in early.items():
        out["early|" + k] = v
    wq = {}
    for k, v in sd.items():
        if k.endswith("fgate._mask"):
            continue
        k = pub(k)
        base = w0[k] if k in w0 else torch.ones_like(v)      # lazy masks are created as ones (models.py:1533)
        delta = (v.detach() - base).float()
        scale = float(delta.abs().max()) / 127.0
        q = torch.zeros_like(delta, dtype=torch.int8) if scale == 0.0 else torch.round(delta / scale).clamp(-127, 127).to(torch.int8)
        out["q|" + k] = q.numpy()
        out["s|" + k] = np.float32(scale)
        wq[k] = base + np.float32(scale) * q.float()         # fp32 multiply-add: what the tests recompute
    alias = {k.replace("freq_mask", "_mask"): v for k, v in wq.items() if k.endswith("fgate.freq_mask")}
    lit.load_state_dict({**wq, **alias}, strict=True)
    lit.eval()      # no dropout / batch statistics in this family; eval == train arithmetic

    # ---- 4: the reference on configs[0] ---------------------------------------------------------------------------
    x, lab = O.phantom_batch(128, 64, 64, seed=42, ignore_frac=0.01)
    loss, logits, taps, tgrads, pgrads = run_with_taps(lit, x, lab)
    metrics = H.per_class_metrics_3d(logits, lab, O.NUM_CLASSES, ignore_index=O.IGNORE_INDEX)
    print(f"configs[0]: loss {loss:.5f} macro dice {metrics[3]:.4f} ({time.time() - t0:.0f}s)")
    out.update(loss=np.float64(loss), argmax=logits.argmax(1).numpy().astype(np.uint8), logits_norm=np.float64(logits.double().norm()),
               logits_sample=strided(logits), dice_list=np.array(metrics[0], np.float64), sens_list=np.array(metrics[1], np.float64),
               spec_list=np.array(metrics[2], np.float64), scalars=np.array(metrics[3:], np.float64))
    srt = torch.sort(logits, dim=1, descending=True).values
    out["margin_q01"] = np.float64(torch.quantile((srt[:, 0] - srt[:, 1]).reshape(-1)[::37], 0.01))
    for b in BLOCKS:
        out["act|" + b] = strided(taps[b])
        out["act_norm|" + b] = np.float64(taps[b].double().norm())
        out["dact|" + b] = strided(tgrads[b])
        out["dact_norm|" + b] = np.float64(tgrads[b].double().norm())
        if b.startswith("enc"):
            ws = win_sum(tgrads[b])
            out["dactw|" + b] = strided(ws)
            out["dactw_norm|" + b] = np.float64(ws.double().norm())
    names = sorted(pgrads)
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array([float(pgrads[k].double().norm()) for k in names], np.float64)
    for k in names:
        g = pgrads[k]
        out["g|" + k] = g.numpy() if g.numel() <= NSAMPLE else strided(g)

    path = GOLD / "trained_SPFFUNet_cfg0.npz"
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({path.stat().st_size / 1e6:.2f} MB) in {time.time() - t0:.0f}s")

    # ---- 5: what bf16 costs the reference itself (CPU autocast; torch's CPU bf16 conv3d is slow: 2 slices) ----
    out = {}
    xs, ls = x[:YARD_SLICES], lab[:YARD_SLICES]
    _, lg32, t32, d32, p32 = run_with_taps(lit, xs, ls)
    _, lg16, t16, d16, p16 = run_with_taps(lit, xs, ls, autocast=True)
    yard = {"logits": rel(lg16, lg32)}
    for b in BLOCKS:
        yard["act|" + b] = rel(t16[b], t32[b])
        yard["dact|" + b] = rel(d16[b], d32[b])
        if b.startswith("enc"):
            yard["dactw|" + b] = rel(win_sum(d16[b]), win_sum(d32[b]))
    for k in names:
        if float(p32[k].norm()) > 0:
            yard["g|" + k] = rel(p16[k], p32[k])
    out["yard_names"] = np.array(sorted(yard))
    out["yard_vals"] = np.array([yard[k] for k in sorted(yard)], np.float64)
    out["yard_argmax_agree"] = np.float64((lg16.argmax(1) == lg32.argmax(1)).float().mean())
    print("autocast yardstick:", {k: round(v, 4) for k, v in yard.items() if not k.startswith("g|")},
          "argmax agreement", float(out["yard_argmax_agree"]))
    print("autocast param-grad yardstick (worst 8):", sorted(((round(v, 4), k) for k, v in yard.items() if k.startswith("g|")), reverse=True)[:8])

    out["yard_slices"] = np.int64(YARD_SLICES)
    path = GOLD / "trained_SPFFUNet_yardstick.npz"
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({path.stat().st_size / 1e6:.2f} MB) in {time.time() - t0:.0f}s")


if __name__ == "__main__":
    main()
