import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spff-unet-spcct_b200"))
from spff_b200 import ops
from spff_b200._lib import Shape
n, d, h, w, cin, cout = [int(a) for a in sys.argv[1:7]]
torch.manual_seed(0)
x = torch.randn(n, d, h, w, cin, device="cuda").to(torch.bfloat16)
wt = torch.randn(cout, cin, 3, 3, 3, device="cuda") * (1.0 / (27 * cin) ** 0.5)
wf, _ = ops.pack_conv3_weight(wt)
slots = ops.conv3d_k3_stat_slots(Shape(n, d, h, w))
partial = torch.full((n, slots, 2, cout), float("nan"), device="cuda")
y = torch.empty(n, d, h, w, cout, dtype=torch.bfloat16, device="cuda")
ops.conv3d_k3_fwd_stats(x, cin, wf, y, cout, partial)
torch.cuda.synchronize()
got = partial.sum(1)            # [n,2,cout]
ref_sum = y.float().sum(dim=(1, 2, 3)); ref_sq = (y.float() ** 2).sum(dim=(1, 2, 3))
print("slots", slots)
print("sum  got", got[0, 0, :8].tolist(), "\n     ref", ref_sum[0, :8].tolist())
print("sum  got[32:40]", got[0, 0, 32:40].tolist(), "\n     ref", ref_sum[0, 32:40].tolist())
print("sq   got", got[0, 1, :4].tolist(), "ref", ref_sq[0, :4].tolist())
print("ratio", (got[0, 0] / ref_sum[0])[:8].tolist())
