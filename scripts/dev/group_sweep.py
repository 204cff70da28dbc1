"""Step time of the fused SPFF-UNet training step against the sample-group size (x[1024,1,5,128,128])."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "spff-unet-spcct_b200"))
from innovative3D import config as C  # noqa: E402


def main():
    torch.manual_seed(0)
    lit = dict((n, b) for n, b, *_ in C.VARIANTS)["SPFF-UNet"]().cuda()
    x = torch.randn(1024, 1, 5, 128, 128, device="cuda")
    lab = torch.randint(0, 13, (1024, 5, 128, 128), device="cuda")
    for g in [int(a) for a in sys.argv[1:]] or [256, 128, 64, 32]:
        lit.model.engine.release_buffers()
        torch.cuda.empty_cache()
        for _ in range(2):
            lit.fit_step((x, lab), sample_group=g)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(3):
            lit.fit_step((x, lab), sample_group=g)
        e1.record()
        t_host = (time.time() - t0) / 3 * 1e3
        torch.cuda.synchronize()
        print(f"group {g:4d}: {e0.elapsed_time(e1) / 3:8.2f} ms/step (host enqueue {t_host:7.2f} ms)", flush=True)


if __name__ == "__main__":
    main()
