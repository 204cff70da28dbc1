"""Diagnostic: per-parameter gradient error of the 3DUNet fused step and autograd path against the CPU oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "spff-unet-spcct_b200")]
import torch
from innovative3D import config as C
from oracle import cicek_oracle as CO, spff_oracle as O

def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))

b, h, w = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
weights = CO.det_weights(seed=42)
x, lab = O.phantom_batch(b, h, w, seed=5, ignore_frac=0.02)
ref_loss, ref_logits, ref_grads, _ = CO.loss_and_grads(weights, x, lab)
mk = lambda: dict((v[0], v[1]) for v in C.VARIANTS)["3DUNet"]().cuda()
A, Bm = mk(), mk()
for m in (A, Bm):
    m.load_state_dict(weights, strict=True); m.backbone.materialize(); m.train()
out = A.fit_step((x.cuda(), lab.cuda()), optimize=False)
loss = Bm.training_step((x.cuda(), lab.cuda()), 0)
loss.backward()
print("loss fused", float(out["loss"]), "autograd", float(loss), "oracle", ref_loss)
G = A.fused_grads()
pb = dict(Bm.backbone.named_parameters())
for n in G:
    r = ref_grads["backbone." + n]
    print(f"{n:24s} |ref| {float(r.norm()):.3e}  fused {rel(G[n], r):.3e}  autograd {rel(pb[n].grad, r):.3e}  fused-vs-auto {rel(G[n], pb[n].grad):.3e}")
