"""Diagnostic: per-parameter gradient / per-block activation error of the GPU path vs the oracle.
usage: diag_grads.py [variant] [b] [h] [w]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "spff-unet-spcct_b200")]
from innovative3D import config as C
from oracle import spff_oracle as O

variant = sys.argv[1] if len(sys.argv) > 1 else "SPFF-UNet"
b, h, w = [int(a) for a in (sys.argv[2:5] if len(sys.argv) > 4 else (2, 16, 16))]
lit = dict((v[0], v[1]) for v in C.VARIANTS)[variant]().cuda()
lit.model.materialize(5)
wts = O.det_weights(O.param_shapes(variant), seed=42)
alias = {k.replace("freq_mask", "_mask"): v for k, v in wts.items() if k.endswith("freq_mask")}
lit.load_state_dict({**wts, **alias})
x, lab = O.phantom_batch(b, h, w, seed=100, ignore_frac=0.02)
ref_loss, ref_logits, ref_grads = O.loss_and_grads(wts, x, lab, variant)
rel = lambda a, r: float((a.detach().float().cpu() - r).norm() / (r.norm() + 1e-30))
out = lit.fit_step((x.cuda(), lab.cuda()), optimize=False)
print("loss", float(out["loss"]), ref_loss)
taps = {}
O.unet_forward(wts, x, variant, taps)
B = lit.model.engine.buffers(b, 5, h, w, torch.device("cuda", 0), train=True)
for name, t in taps.items():
    print(f"act {name:6s} {rel(B.out[name].permute(0, 4, 1, 2, 3), t):.4f}")
print(f"logits {rel(B.logits, ref_logits):.4f}")
for k, g in lit.fused_grads().items():
    r = ref_grads["model." + k]
    print(f"{k:32s} |ref| {float(r.norm()):.3e}  rel {rel(g, r):.4f}")
