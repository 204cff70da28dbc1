"""Run one conv3 wgrad layer a few times (ncu target). usage: one_wgrad.py n h cin cout [iters]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spff-unet-spcct_b200"))
from spff_b200 import ops
n, h, cin, cout = [int(a) for a in sys.argv[1:5]]
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
x = torch.randn(n, 5, h, h, cin, device="cuda").to(torch.bfloat16)
dy = torch.randn(n, 5, h, h, cout, device="cuda").to(torch.bfloat16)
dw = torch.zeros(cout, cin, 3, 3, 3, device="cuda")
ws = ops.conv3d_k3_wgrad_workspace(cin, cout, x)
for _ in range(iters):
    ops.conv3d_k3_wgrad(x, cin, dy, cout, dw, 0.0, ws)
torch.cuda.synchronize()
print("done")
