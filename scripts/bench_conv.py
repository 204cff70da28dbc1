"""Micro-benchmark of the conv3 kernels at the layer shapes of config 2 (per-sample-group).
Prints achieved TFLOP/s (algorithmic 2*27*cin*cout per position). Not the repo's headline bench."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spff-unet-spcct_b200"))
from spff_b200 import ops  # noqa: E402

LAYERS = [  # (h, cin, cout)
    (128, 32, 32), (128, 64, 32), (64, 32, 64), (64, 64, 64), (64, 128, 64),
    (32, 64, 128), (32, 128, 128), (32, 256, 128), (16, 128, 256), (16, 256, 256),
]


def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    which = sys.argv[2] if len(sys.argv) > 2 else "fwd,dgrad,wgrad"
    if len(sys.argv) > 3 and sys.argv[3] == "rows":      # force the flattened-row kernel of conv3_fprop.cu (debug key 7)
        from spff_b200 import _lib
        _lib.lib.spff_debug_set(7, 1)
    if len(sys.argv) > 3 and sys.argv[3].startswith("key"):      # keyK=V: spff_debug_set(K, V)
        from spff_b200 import _lib
        k, v = sys.argv[3][3:].split("=")
        _lib.lib.spff_debug_set(int(k), int(v))
    if len(sys.argv) > 3 and sys.argv[3].startswith("dbg"):
        from spff_b200 import _lib
        _lib.lib.spff_debug_set(5, int(sys.argv[3][3:]))
    for h, cin, cout in LAYERS:
        x = torch.randn(n, 5, h, h, cin, device="cuda").to(torch.bfloat16)
        dy = torch.randn(n, 5, h, h, cout, device="cuda").to(torch.bfloat16)
        w = torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.05
        wf, wd = ops.pack_conv3_weight(w)
        y = torch.empty(n, 5, h, h, cout, device="cuda", dtype=torch.bfloat16)
        dx = torch.empty(n, 5, h, h, cin, device="cuda", dtype=torch.bfloat16)
        flop = 2.0 * 27 * cin * cout * n * 5 * h * h
        line = f"{h:4d}^2 {cin:3d}->{cout:3d}  M={n*5*h*h:9d}"
        if "fwd" in which:
            t = timeit(lambda: ops.conv3d_k3_fwd(x, cin, wf, y, cout))
            line += f"  fwd {t*1e3:8.3f} ms {flop/t/1e12:7.1f} TF/s"
        if "dgrad" in which:
            t = timeit(lambda: ops.conv3d_k3_dgrad(dy, cout, wd, dx, cin))
            line += f"  dgrad {t*1e3:8.3f} ms {flop/t/1e12:7.1f} TF/s"
        if "wgrad" in which and hasattr(ops, "conv3d_k3_wgrad"):
            dw = torch.zeros(cout, cin, 3, 3, 3, device="cuda")
            ws = ops.conv3d_k3_wgrad_workspace(cin, cout, x)
            t = timeit(lambda: ops.conv3d_k3_wgrad(x, cin, dy, cout, dw, 0.0, ws))
            line += f"  wgrad {t*1e3:8.3f} ms {flop/t/1e12:7.1f} TF/s"
        print(line, flush=True)
        del x, dy, y, dx


if __name__ == "__main__":
    main()
