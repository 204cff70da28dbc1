"""Diagnostic: compare every intermediate buffer of two forward passes (cached vs fresh/poisoned buffers)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "spff-unet-spcct_b200")]
from innovative3D import config as C
from oracle import spff_oracle as O
from spff_b200.engine import GateTables, BLOCKS
variant = sys.argv[1] if len(sys.argv) > 1 else "SPFF-UNet"
b, h, w = [int(a) for a in (sys.argv[2:5] if len(sys.argv) > 4 else (2, 128, 128))]
torch.manual_seed(0)
lit = dict((v[0], v[1]) for v in C.VARIANTS)[variant]().cuda()
x, lab = O.phantom_batch(b, h, w, seed=3)
xg = x.cuda()
core = lit.model
core.materialize(5)
eng = core.engine
eng.refresh_weights()
T = GateTables(eng.cfg, eng.params(), 5, need_grad=False)
dev = xg.device
B1 = eng.buffers(b, 5, h, w, dev, train=False, fresh=True)
eng.forward_group(B1, T, xg)
os.environ["SPFF_POISON"] = "1"
B2 = eng.buffers(b, 5, h, w, dev, train=True, fresh=True)
eng.forward_group(B2, T, xg)
torch.cuda.synchronize()
def cmp(name, a, c):
    a, c = a.float(), c.float()
    nan = int(torch.isnan(c).sum())
    d = float((a - c).abs().nan_to_num(1e9).max())
    print(f"{name:12s} maxdiff {d:.4g} nan {nan}")
for blk in BLOCKS:
    for nm in ("x1", "a1", "x2", "out"):
        cmp(f"{blk}.{nm}", getattr(B1, nm)[blk], getattr(B2, nm)[blk])
    if blk in ("enc1", "enc2", "enc3"):
        l = int(blk[-1]); cmp(f"pool{l}", B1.pool[l], B2.pool[l])
    if blk.startswith("dec"):
        l = int(blk[-1]); cmp(f"cat{l}", B1.cat[l], B2.cat[l])
cmp("logits", B1.logits, B2.logits)
