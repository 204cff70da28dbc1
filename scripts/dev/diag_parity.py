"""Developer diagnostic (not a test): per-layer parity numbers of the GPU path against the CPU oracle on trained weights,
for several batch shapes — parameter gradients, per-block activations and per-block activation gradients."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "spff-unet-spcct_b200")]
from oracle import spff_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    from innovative3D import config as C
    torch.manual_seed(42)
    variant = sys.argv[1] if len(sys.argv) > 1 else "SPFF-UNet"
    lit = dict((v[0], v[1]) for v in C.VARIANTS)[variant]().cuda()
    lit.hparams["lr"] = 1e-3
    for i in range(80):
        x, lab = O.phantom_batch(8, 32, 32, seed=7 + i)
        out = lit.fit_step((x.cuda(), lab.cuda()))
    print("final train loss", float(out["loss"]))
    sd = {k: v.detach().cpu().clone() for k, v in lit.state_dict().items() if not k.endswith("fgate._mask")}
    eng = lit.model.engine
    for (b, h, w) in ((2, 128, 128), (8, 128, 128), (128, 64, 64)):
        x, lab = O.phantom_batch(b, h, w, seed=999, ignore_frac=0.01)
        t0 = time.time()
        q = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
        taps = {}
        logits = O.unet_forward(q, x, variant, taps)
        for t in taps.values():
            t.retain_grad()
        loss = O.ce_plus_macro_dice_loss(logits, lab, logits.shape[1])
        loss.backward()
        print(f"--- batch {b}x{h}x{w}: oracle {time.time() - t0:.1f}s loss {float(loss):.5f}")
        out = lit.fit_step((x.cuda(), lab.cuda()), optimize=False, sample_group=b)
        G = lit.fused_grads()
        errs = {n: rel(g, q["model." + n].grad) for n, g in G.items() if float(q["model." + n].grad.norm()) > 1e-9}
        top = sorted(errs.items(), key=lambda kv: -kv[1])
        print("loss gpu", float(out["loss"]))
        print("worst param grads:", [(k, round(v, 4)) for k, v in top[:14]])
        conv = {k: v for k, v in errs.items() if k.endswith(".0.weight") and "efilm" not in k}
        print("conv weights:", [(k, round(v, 4)) for k, v in sorted(conv.items())])
        print("up/out:", [(k, round(v, 4)) for k, v in errs.items() if k.startswith("up") or k.startswith("out")])
        dev = torch.device("cuda", torch.cuda.current_device())
        B = eng.buffers(b, 5, h, w, dev, train=True)
        nchw = lambda t: t.permute(0, 4, 1, 2, 3).float()
        acts = {n: rel(nchw(B.out[n]), taps[n]) for n in taps}
        print("block activations:", {k: round(v, 4) for k, v in acts.items()})
        gr = {"dec1": B.gout[1], "dec2": B.gout[2], "dec3": B.gout[3], "bott": B.gout[4]}
        for l, e in ((1, "enc1"), (2, "enc2"), (3, "enc3")):
            gr[e] = B.dcat[l][..., B.C[l]:]
        print("block output gradients:", {k: round(rel(nchw(v), taps[k].grad), 4) for k, v in gr.items()})
        with torch.no_grad():
            lg = lit(x.cuda())
        print("logits rel", rel(lg, logits), "argmax agree", float((lg.argmax(1).cpu() == logits.argmax(1)).float().mean()))


if __name__ == "__main__":
    main()
