"""Yardstick (test infrastructure): error of PyTorch's own CPU bf16 autocast of the oracle vs fp32 on briefly
trained weights, per parameter gradient. Measured here: logits 0.5 %, conv-weight gradients 1.7 % median,
2.5-3.0 % on the 16x16 bottleneck layers. Run: python oracle/probe_autocast_bf16.py (about 3 min on 8 cores)."""
import sys, time, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import spff_oracle as O
torch.manual_seed(0)
torch.set_num_threads(8)
p = O.det_weights(O.param_shapes("SPFF-UNet"), seed=42)
# reference-style init is better conditioned; use small random init scaled like torch default
t0=time.time()
p = O.pretrain(p, "SPFF-UNet", 60, 8, 32, 32, lr=1e-3)
print("pretrain", time.time()-t0)
x, lab = O.phantom_batch(2, 128, 128, seed=999, ignore_frac=0.01)
l32, logits32, g32 = O.loss_and_grads(p, x, lab)
q = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
with torch.autocast("cpu", dtype=torch.bfloat16):
    logits = O.unet_forward(q, x, "SPFF-UNet")
loss = O.ce_plus_macro_dice_loss(logits.float(), lab)
loss.backward()
rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))
print("logits rel", rel(logits.float().detach(), logits32), "argmax agree", float((logits.argmax(1) == logits32.argmax(1)).float().mean()))
errs = {k: rel(q[k].grad, g32[k]) for k in g32 if float(g32[k].norm()) > 1e-6}
for k, v in sorted(errs.items(), key=lambda kv: -kv[1])[:12]:
    print(f"{k:36s} {v:.4f}")
convs = {k: v for k, v in errs.items() if k.endswith(".0.weight")}
print("conv weight grads: max", max(convs.values()), "median", sorted(convs.values())[len(convs)//2])
