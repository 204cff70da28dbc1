"""Generate tests/golden/datapath.npz from the REFERENCE's own data-path functions (TEST INFRASTRUCTURE; build
container only): `innovative3D.datasets.TrainGridAug` (datasets.py:134-206) on seeded samples with a seeded `random`,
and `innovative3D.helpers.is_pixel_in_ellipse` (helpers.py:125-129) driven by the ROI loop of helpers.py:197-206
(the loop is restated here because the function around it reads DICOM files). noise_p = 0: the noise draws from
torch's generator and has no bit-level counterpart on the device."""
from __future__ import annotations

import random
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import datapath_oracle as DO  # noqa: E402
from oracle.make_golden import GOLD, import_reference  # noqa: E402

# (seed, frames, H, W, gs)
AUG_CASES = [(1, 5, 64, 64, 3), (2, 5, 64, 64, None), (3, 5, 48, 48, 5), (4, 5, 40, 64, 2), (5, 5, 64, 64, 4), (6, 3, 96, 96, 5),
             (7, 5, 64, 64, 1), (8, 5, 32, 32, 3)]
ROI_CASES = [  # (height, width, rois)
    (64, 64, [(5, 6, 20, 14, 3), (18, 10, 16, 30, 7), (40, 40, 9, 9, 1), (0, 50, 12, 10, 12), (30, 2, 1, 1, 4), (-4, 20, 10, 8, 5)]),
    (48, 80, [(10, 5, 50, 40, 2), (20, 15, 25, 12, 9), (60, 30, 20, 18, 6), (33, 3, 2, 44, 11)]),
]


def main():
    M, H = import_reference()
    import innovative3D.datasets as D
    out = {}
    aug = D.TrainGridAug(noise_p=0.0)
    for i, (seed, f, h, w, gs) in enumerate(AUG_CASES):
        x, y = DO.aug_input(seed, f, h, w)
        random.seed(1000 + seed)
        if h != w:      # odd rotations change the shape of non-square samples: keep the case rotation free
            aug_i = D.TrainGridAug(noise_p=0.0, rot90_p=0.0)
            xo, yo = aug_i(x.clone(), y.clone(), gs)
        else:
            xo, yo = aug(x.clone(), y.clone(), gs)
        out[f"aug{i}_x"] = xo.numpy().astype(np.float32)
        out[f"aug{i}_y"] = yo.numpy().astype(np.int64)
        out[f"aug{i}_next"] = np.float64(random.random())     # the position of the `random` stream after the call
        out[f"aug{i}_case"] = np.array([seed, f, h, w, -1 if gs is None else gs])
    for i, (h, w, rois) in enumerate(ROI_CASES):
        lb = np.zeros((2, h, w), dtype=np.int64)
        for f in range(2):
            for (x0, y0, w0, h0, lab) in rois:       # helpers.py:202-206
                for px in range(x0, x0 + w0):
                    for py in range(y0, y0 + h0):
                        if H.is_pixel_in_ellipse(px, py, (x0, y0, w0, h0)):
                            lb[f, py, px] = lab
        out[f"roi{i}_labels"] = lb
        out[f"roi{i}_rois"] = np.array(rois, dtype=np.int64)
    np.savez_compressed(GOLD / "datapath.npz", **out)
    print("wrote", GOLD / "datapath.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
