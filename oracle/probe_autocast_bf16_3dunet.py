"""Yardstick (test infrastructure): error of PyTorch's own CPU bf16 autocast of the 3DUNet oracle vs fp32, per
parameter gradient, on the name-seeded fixture weights. Measured here at [2,1,5,32,32]: logits 3.0 %, gradients
1.5 % (dec1) rising to 30-48 % in the encoder — ReLU masks that flip under bf16 rounding compound through the
backward pass of an untrained network. Run: python oracle/probe_autocast_bf16_3dunet.py (seconds)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cicek_oracle as CO, spff_oracle as O  # noqa: E402

rel = lambda a, b: float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))
w = CO.det_weights(seed=42)
x, lab = O.phantom_batch(2, 32, 32, seed=5, ignore_frac=0.02)
l0, lg0, g0, _ = CO.loss_and_grads(w, x, lab)
q = {k: (v.detach().clone().requires_grad_(True) if not CO.is_buffer(k) else v.clone()) for k, v in w.items()}
with torch.autocast("cpu", dtype=torch.bfloat16):
    lg = CO.forward(q, x, True, None)
loss = CO.ce_loss(lg.float(), lab)
loss.backward()
print("logits rel", rel(lg.detach(), lg0), "loss", float(loss.detach()), "fp32", l0)
for k in g0:
    if k.endswith(".0.weight") or k.endswith(".3.weight"):
        print(f"{k:32s} {rel(q[k].grad, g0[k]):.4f}")
