"""Diagnostic: run-to-run determinism of the forward and agreement of the fused argmax head."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "spff-unet-spcct_b200")]
from innovative3D import config as C
from oracle import spff_oracle as O
variant = sys.argv[1] if len(sys.argv) > 1 else "SPFF-UNet"
b, h, w = [int(a) for a in (sys.argv[2:5] if len(sys.argv) > 4 else (2, 128, 128))]
torch.manual_seed(0)
lit = dict((v[0], v[1]) for v in C.VARIANTS)[variant]().cuda()
x, lab = O.phantom_batch(b, h, w, seed=3)
xg = x.cuda()
with torch.no_grad():
    l1 = lit(xg).clone()
    junk = torch.full((64 << 20,), float("nan"), device="cuda"); del junk
    l2 = lit(xg).clone()
    lab8 = lit.model.predict_labels(xg)
l3 = lit(xg)
print("infer vs infer bitwise:", torch.equal(l1, l2), float((l1 - l2).abs().max()))
print("infer vs saved bitwise:", torch.equal(l1, l3.detach()), float((l1 - l3.detach()).abs().max()))
am = l1.argmax(1)
print("argmax head agreement:", float((lab8.long() == am).float().mean()), int((lab8.long() != am).sum()))
bad = (lab8.long() != am).nonzero()[:5]
for i in bad:
    n, d, hh, ww = [int(v) for v in i]
    v = l1[n, :, d, hh, ww]
    top = v.topk(2)
    print(i.tolist(), top.values.tolist(), top.indices.tolist(), int(lab8[n, d, hh, ww]))
