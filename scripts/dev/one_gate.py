"""Run the gate micro kernels a few times (ncu target). usage: one_gate.py n c h flags iters"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spff-unet-spcct_b200"))
from spff_b200 import ops
from spff_b200._lib import Shape
n, c, h, flags = [int(a) for a in sys.argv[1:5]]
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
d = 5; hid = max(4, c // 16)
dev = "cuda"
r = lambda *s: torch.randn(*s, device=dev)
S = r(n, d, c).abs() * h * h; R = r(n, d, c, 6); coef = r(n, c, 4).abs(); gamma = r(c)
g1 = 1 + 0.1 * r(c, d); bt = 0.1 * r(c, d); kfg = r(d)
se = (r(hid, c), r(hid), r(c, hid), r(c))
P = torch.empty(n, d, c, device=dev); Q = torch.empty(n, d, c, device=dev)
bcoef = torch.empty(n, c, 4, device=dev); dSa = torch.empty(n, d, c, device=dev); Pout = torch.empty(n, d, c, device=dev)
z = lambda *s: torch.zeros(*s, device=dev)
dgamma, dbeta, dg1, dbt, dk = z(c), z(c), z(c, d), z(c, d), z(d)
dse = (z(hid, c), z(hid), z(c, hid), z(c))
shape = Shape(n, d, h, h)
def fwd(): ops.gate_micro_fwd(S, g1 if flags & 1 else None, bt if flags & 1 else None, kfg if flags & 2 else None, se if flags & 8 else None, flags, c, shape, P, Q)
def bwd(): ops.gate_micro_bwd(R, S, coef, gamma, g1 if flags & 1 else None, bt if flags & 1 else None, kfg if flags & 2 else None,
                              se if flags & 8 else None, flags, c, shape, bcoef, dSa, Pout, dgamma, dbeta, dg1 if flags & 1 else None,
                              dbt if flags & 1 else None, dk if flags & 2 else None, dse if flags & 8 else None)
for name, fn in (("gate_fwd", fwd), ("gate_bwd", bwd)):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name} n={n} c={c} flags={flags}: {e0.elapsed_time(e1)/iters*1e3:.1f} us")
