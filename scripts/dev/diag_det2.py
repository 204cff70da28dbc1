import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "spff-unet-spcct_b200")]
from innovative3D import config as C
from oracle import spff_oracle as O
from spff_b200.engine import GateTables, BLOCKS
b, h, w = [int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (2, 128, 128))]
torch.manual_seed(0)
lit = dict((v[0], v[1]) for v in C.VARIANTS)["SPFF-UNet"]().cuda()
x, lab = O.phantom_batch(b, h, w, seed=3)
xg = x.cuda()
core = lit.model; core.materialize(5); eng = core.engine; eng.refresh_weights()
dev = xg.device
def tabs(need):
    T = GateTables(eng.cfg, eng.params(), 5, need_grad=need)
    return T
T0 = tabs(False); T1 = tabs(True); T2 = tabs(False)
for blk in BLOCKS:
    for nm in ("g1", "bt", "kfg"):
        a, c, e = getattr(T0, nm)[blk], getattr(T1, nm)[blk], getattr(T2, nm)[blk]
        print(blk, nm, "nograd-vs-grad", float((a - c.detach()).abs().max()), "nograd-vs-nograd", float((a - e).abs().max()))
B = eng.buffers(b, 5, h, w, dev, train=False)
ref = None
for it in range(4):
    eng.forward_group(B, T0, xg)
    torch.cuda.synchronize()
    snap = {f"{blk}.{nm}": getattr(B, nm)[blk].clone() for blk in BLOCKS for nm in ("x1", "a1", "x2", "out")}
    snap["logits"] = B.logits.clone()
    if ref is None:
        ref = snap
    else:
        bad = [(k, float((snap[k].float() - ref[k].float()).abs().max())) for k in snap if not torch.equal(snap[k], ref[k])]
        print("iter", it, "mismatching buffers:", bad[:6])
lab8 = torch.empty(b, 5, h, w, dtype=torch.uint8, device=dev)
eng.forward_group(B, T0, xg, head="argmax", labels_out=lab8)
torch.cuda.synchronize()
am = ref["logits"].argmax(1)
print("argmax mismatches", int((lab8.long() != am).sum()), "dec1.out equal", torch.equal(B.out["dec1"], ref["dec1.out"]))
