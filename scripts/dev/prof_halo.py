"""Developer tool: per-role wait cycles of the halo conv kernel (spff_debug_set key 4 = counter buffer)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "spff-unet-spcct_b200"))
from spff_b200 import ops, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for h, cin, cout in [(128, 32, 32), (128, 64, 32), (64, 32, 64), (64, 64, 64), (32, 128, 128), (16, 256, 256)]:
    x = torch.randn(n, 5, h, h, cin, device="cuda").to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.05
    wf, wd = ops.pack_conv3_weight(w)
    y = torch.empty(n, 5, h, h, cout, device="cuda", dtype=torch.bfloat16)
    ops.conv3d_k3_fwd(x, cin, wf, y, cout)
    buf = torch.zeros(148 * 3 * 4, dtype=torch.int64, device="cuda")
    _lib.lib.spff_debug_set(4, buf.data_ptr())
    ops.conv3d_k3_fwd(x, cin, wf, y, cout)
    torch.cuda.synchronize()
    _lib.lib.spff_debug_set(4, 0)
    b = buf.view(148, 3, 4).double().mean(0)
    names = ["epilogue: total, wait acc_full", "producer: total, wait aempty, wait bempty", "issuer:   total, wait afull, wait bfull, wait acc_empty"]
    print(f"{h}^2 {cin}->{cout}")
    for r in range(3):
        t = b[r]
        print(f"   {names[r]:58s} {t[0]:10.0f} " + " ".join(f"{float(v) / float(t[0]) * 100:5.1f}%" for v in t[1:]))
