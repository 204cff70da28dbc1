"""Generate tests/golden/cicek*.npz by running the REFERENCE's "3DUNet" variant itself
(`make_cicek_depth_adapter_sgd_wce`, config.py:283-303 -> LitCicek3DUNet_DepthAdapter_Published,
models.py:753-846) on seeded inputs and name-seeded weights. TEST INFRASTRUCTURE; build container only
(`python oracle/make_golden_3dunet.py`) — see oracle/make_golden.py for the stub shim it shares.

Stored per case: logits, the CE loss, per_class_metrics_3d scalars, gradient norms + strided samples, the
BatchNorm running buffers after the training-mode forward, eval-mode logits, and the parameters'
norms after one SGD(lr 1e-2, momentum 0.99) step + a second step (momentum buffer in use).
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import cicek_oracle as CO  # noqa: E402
from oracle import spff_oracle as O  # noqa: E402
from oracle.make_golden import GOLD, compact_grads, import_reference  # noqa: E402

# (batch, H, W, ignore fraction)
CASES = [(2, 16, 16, 0.02), (1, 32, 32, 0.0), (3, 32, 16, 0.01)]


def main():
    M, H = import_reference()
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(42)
    lit = M.LitCicek3DUNet_DepthAdapter_Published(num_classes=O.NUM_CLASSES, lr=1e-2, momentum=0.99, nesterov=False,
                                                  weight_decay=0.0, ignore_index=255, class_weights=None,
                                                  voxel_weight_key=None, ce_weight=1.0, dice_weight=0.0, use_bn=True,
                                                  target_depth=16, include_bg_in_dice=False)
    # the stub LightningModule's argument-less save_hyperparameters() does not inspect the caller's frame as the
    # real one does: record what configure_optimizers reads (models.py:844-846)
    lit.hparams.__dict__.update(lr=1e-2, momentum=0.99, nesterov=False, weight_decay=0.0, num_classes=O.NUM_CLASSES)
    sd = lit.state_dict()
    np.savez_compressed(GOLD / "cicek_init_seed42.npz", **{
        k: np.array(list(v.shape) + [float(v.double().sum()), float(v.double().abs().sum())], dtype=np.float64)
        for k, v in sd.items()})
    for idx, (b, h, w, ign) in enumerate(CASES):
        seed = 300 + idx
        x, lab = O.phantom_batch(b, h, w, seed=seed, ignore_frac=ign)
        weights = CO.det_weights(seed=42)
        assert set(weights) == set(sd), sorted(set(weights) ^ set(sd))
        lit.load_state_dict(weights, strict=True)
        lit.train()
        lit.zero_grad()
        logits = lit(x)
        loss = lit._weighted_softmax_ce(logits, lab, None) * lit.ce_weight
        loss.backward()
        metrics = H.per_class_metrics_3d(logits.detach(), lab, O.NUM_CLASSES, ignore_index=255)
        grads = {k: p.grad for k, p in lit.named_parameters()}
        norms, samples = compact_grads(grads)
        names = sorted(norms)
        after = lit.state_dict()
        buf_names = sorted(k for k in after if k.endswith(("running_mean", "running_var")))
        out = dict(
            logits=logits.detach().numpy().astype(np.float32), loss=np.float64(float(loss)),
            scalars=np.array(metrics[3:], np.float64), grad_names=np.array(names),
            grad_norms=np.array([norms[k] for k in names], np.float64),
            buf_names=np.array(buf_names), case=np.array([str(b), str(h), str(w), str(ign), str(seed)]),
            nbt=np.int64(int(after["backbone.enc1.1.num_batches_tracked"])),
        )
        for k in names:
            out["g|" + k] = samples[k]
        for k in buf_names:
            out["b|" + k] = after[k].detach().numpy().astype(np.float32)
        # two SGD steps as configure_optimizers builds them (models.py:844-846)
        opt = lit.configure_optimizers()
        opt.step()
        out["p1_norms"] = np.array([float(dict(lit.named_parameters())[k].detach().double().norm()) for k in names])
        lit.zero_grad()
        loss2 = lit._weighted_softmax_ce(lit(x), lab, None)
        loss2.backward()
        opt.step()
        out["loss2"] = np.float64(float(loss2))
        out["p2_norms"] = np.array([float(dict(lit.named_parameters())[k].detach().double().norm()) for k in names])
        # eval-mode forward (running statistics) on the ORIGINAL weights
        lit.load_state_dict(weights, strict=True)
        lit.eval()
        with torch.no_grad():
            out["logits_eval"] = lit(x).numpy().astype(np.float32)
        np.savez_compressed(GOLD / f"cicek{idx}_3DUNet_{b}x{h}x{w}.npz", **out)
        print(f"cicek case {idx}: b={b} {h}x{w} loss={float(loss):.6f} loss2={float(loss2):.6f} macro_dice={metrics[3]:.4f}")


if __name__ == "__main__":
    main()
