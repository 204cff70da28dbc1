"""Micro-benchmark of the data-path kernels at the bench batch: TrainGridAug over x[1024,1,5,128,128] fp32 + int64 labels,
and the ROI rasteriser at 5 x 512 x 512. Prints algorithmic GB/s (bytes read + written / CUDA-event time)."""
import os, random, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "spff-unet-spcct_b200")]
from innovative3D.datapath_gpu import TrainGridAug, rasterize_roi_labels
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.randn(n, 1, 5, 128, 128, device="cuda")
y = torch.randint(0, 13, (n, 5, 128, 128), device="cuda")
for name, aug, lab in (("aug (no noise), int64 labels", TrainGridAug(noise_p=0.0), y), ("aug + noise, int64 labels", TrainGridAug(noise_p=1.0), y),
                       ("aug (no noise), uint8 labels", TrainGridAug(noise_p=0.0), y.to(torch.uint8))):
    random.seed(0)
    aug(x, lab, 4); torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        aug(x, lab, 4)
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 5
    nbytes = 2 * (x.numel() * 4 + lab.numel() * lab.element_size())
    print(f"{name:32s} {wall*1e3:7.2f} ms wall per batch (host draws + tables + kernel), {nbytes/1e9:.2f} GB moved")
rois = [(60 + 40 * i, 80 + 35 * i, 60, 55, i + 1) for i in range(9)]
rasterize_roi_labels(rois, 5, 512, 512); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    rasterize_roi_labels(rois, 5, 512, 512)
e1.record(); torch.cuda.synchronize()
print(f"roi labels 5x512x512, 9 rois: {e0.elapsed_time(e1)/20*1e3:.1f} us per volume")
# kernel-only time of the gather (tables already on the device)
from spff_b200 import ops
import numpy as np
aug = TrainGridAug(noise_p=0.0)
random.seed(0)
tabs = [aug._draw(128, 128, 4) for _ in range(n)]
dev = x.device
amap = torch.from_numpy(np.stack([t[0].a for t in tabs])).to(dev); bmap = torch.from_numpy(np.stack([t[0].b for t in tabs])).to(dev)
tr = torch.tensor([t[0].t for t in tabs], dtype=torch.int32, device=dev); sc = torch.tensor([t[1] for t in tabs], device=dev)
sh = torch.tensor([t[2] for t in tabs], device=dev); nz = torch.zeros(n, device=dev); st = torch.tensor([t[4] for t in tabs], dtype=torch.int32, device=dev)
seeds = torch.zeros(n, dtype=torch.int64, device=dev)
xin = x.view(n, 5, 128, 128); xo = torch.empty_like(xin); yo = torch.empty_like(y)
for lab, lo in ((y, yo), (y.to(torch.uint8), yo.to(torch.uint8))):
    ops.grid_aug(xin, lab, xo, lo, amap, bmap, tr, sc, sh, nz, seeds, st, False, True); torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        ops.grid_aug(xin, lab, xo, lo, amap, bmap, tr, sc, sh, nz, seeds, st, False, True)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10 * 1e-3
    nbytes = 2 * (x.numel() * 4 + lab.numel() * lab.element_size())
    print(f"grid_aug kernels only, {lab.dtype}: {t*1e6:.0f} us, {nbytes/t/1e9:.0f} GB/s")
