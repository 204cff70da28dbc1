"""Run the bandwidth-bound kernels once each at one level's shape (ncu target / micro-benchmark).
usage: one_ew.py [n] [h] [c] [iters]   -> prints achieved GB/s per kernel (algorithmic bytes / event time)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spff-unet-spcct_b200"))
from spff_b200 import ops
from spff_b200._lib import Shape
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
h = int(sys.argv[2]) if len(sys.argv) > 2 else 128
c = int(sys.argv[3]) if len(sys.argv) > 3 else 32
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
d = 5
dev = "cuda"
from spff_b200 import _lib
for key, env in ((3, "EW_GRID4"), (4, "EW_GRID8")):
    if os.getenv(env):
        _lib.lib.spff_debug_set(key, int(os.environ[env]))
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
x = bf(n, d, h, h, c); dout = bf(n, d, h, h, c); y = torch.empty_like(x); dx = torch.empty_like(x)
cat = bf(n, d, h, h, 2 * c)
yp = torch.empty(n, d, h // 2, h // 2, c, dtype=torch.bfloat16, device=dev); dpool = bf(n, d, h // 2, h // 2, c)
gamma = torch.ones(c, device=dev); beta = torch.zeros(c, device=dev)
stats = torch.zeros(n, c, 2, dtype=torch.float64, device=dev); coef = torch.empty(n, c, 4, device=dev)
ops.in_stats(x, c, stats); ops.in_coeffs(stats, gamma, beta, 1e-5, n, c, d * h * h, coef)
S = torch.zeros(n, d, c, device=dev); P = torch.rand(n, d, c, device=dev); Q = torch.rand(n, d, c, device=dev)
R = torch.zeros(n, d, c, 6, device=dev); bcoef = torch.rand(n, c, 4, device=dev)
vox = n * d * h * h
el = vox * c * 2  # bytes of one bf16 tensor
xin = torch.randn(n, 1, d, h, h, device=dev); wst = torch.randn(32, 1, 3, 3, 3, device=dev); dwst = torch.zeros_like(wst)
y32 = torch.empty(n, d, h, h, 32, dtype=torch.bfloat16, device=dev); dy32 = bf(n, d, h, h, 32)
wh = torch.randn(13, 32, device=dev); bh = torch.zeros(13, device=dev); lab = torch.randint(0, 13, (n, d, h, h), device=dev)
acc = torch.zeros(1, dtype=torch.float64, device=dev); cnt = torch.zeros(1, dtype=torch.int64, device=dev)
conf = torch.zeros(13, 13, dtype=torch.int64, device=dev); nv = torch.tensor([vox], device=dev)
dwh = torch.zeros(13, 32, device=dev); dbh = torch.zeros(13, device=dev)
cases = [
    ("in_stats", el, lambda: ops.in_stats(x, c, stats)),
    ("norm_act_apply", 2 * el, lambda: ops.norm_act_apply(x, coef, y, c, 0.01)),
    ("norm_act_reduce", el, lambda: ops.norm_act_reduce(x, coef, S, c, 0.01)),
    ("affine_apply+pool", 2.25 * el, lambda: ops.norm_act_affine_apply(x, coef, P, Q, cat[..., c:], yp, c, 0.01)),
    ("bwd_reduce", 2 * el, lambda: ops.norm_act_bwd_reduce(dout, x, coef, R, c, 0.01)),
    ("bwd_reduce_plain", 2 * el, lambda: ops.norm_act_bwd_reduce(dout, x, coef, R, c, 0.01, plain=True)),
    ("bwd_reduce_fixed_lean", 2 * el, lambda: ops.norm_act_bwd_reduce(dout, x, coef, R, c, 0.01, fixed_order=True, S=S)),
    ("bwd_reduce_fixed_old", 2 * el, lambda: ops.norm_act_bwd_reduce(dout, x, coef, R, c, 0.01, fixed_order=True)),
    ("bwd_reduce_plain_fixed_lean", 2 * el, lambda: ops.norm_act_bwd_reduce(dout, x, coef, R, c, 0.01, plain=True, fixed_order=True)),
    ("norm_act_reduce_fixed", el, lambda: ops.norm_act_reduce(x, coef, S, c, 0.01, fixed_order=True)),
    ("bwd_apply", 3 * el, lambda: ops.norm_act_bwd_apply(dout, x, coef, bcoef, P, Q, dx, c, 0.01)),
    ("maxpool_bwd_add", 3.25 * el, lambda: ops.maxpool_bwd_add(dpool, cat[..., c:], cat[..., :c], c, True)),
]
if c == 32:
    cases += [
        ("stem_fwd", vox * (4 + 64), lambda: ops.conv3d_stem_fwd(xin, wst, y32, 32)),
        ("stem_wgrad", vox * (4 + 64), lambda: ops.conv3d_stem_wgrad(xin, dy32, 32, dwst, 0.0)),
        ("head_loss_fused", vox * (64 + 64 + 8), lambda: ops.head_loss_fused(x, wh, bh, lab, 255, nv, None, acc, cnt, conf, dx, dwh, dbh, 0.0)),
    ]
for name, nbytes, fn in cases:
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / iters * 1e-3
    print(f"{name:20s} {t*1e6:9.1f} us  {nbytes/t/1e9:8.1f} GB/s  ({nbytes/1e6:.0f} MB algorithmic)", flush=True)
