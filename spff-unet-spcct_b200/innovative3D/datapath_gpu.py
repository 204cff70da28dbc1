"""Device side of the reference's data path for the hot path's inputs.

The reference builds phantom labels with a per-pixel Python loop (`helpers.py:197-206`) and augments every sample on
the CPU in DataLoader workers (`datasets.py:56-206`, 16 processes). Here both run on the B200 over whole batches:

  rasterize_roi_labels(...)   labels of the elliptical ROIs, one kernel (spff_roi_labels)
  TrainGridAug                same constructor and the same random decisions, drawn from Python's `random` in the
                              reference's order (so a seeded run takes the same flips / rotations / jitter / stripe
                              permutations), applied by ONE gather kernel per batch (spff_grid_aug): the flips, the
                              rot90 and the separable stripe shuffle compose into two index tables per sample.

DICOM ingest (pydicom, `DicomDataset3D`, the Lightning data modules) stays the reference's: it is file IO, outside the
hot path. CUDA tensors only; there is no CPU fallback.
"""
from __future__ import annotations

import random
from collections import defaultdict
from typing import Optional, Sequence

import numpy as np
import torch

from spff_b200 import ops

from .config import IMAGE_HEIGHT, IMAGE_WIDTH, global_label_names


def scaled_rois(cfg: dict, image_height: int = IMAGE_HEIGHT, image_width: int = IMAGE_WIDTH):
    """[(x0, y0, w0, h0, label)] of a dataset config: the integer scaling of helpers.py:186-195 (1300-pixel frame)."""
    scale_x, scale_y = image_width / 1300.0, image_height / 1300.0
    ox, oy = cfg["offset"]
    rois = []
    for (x, y, w, h, lab_str) in cfg["original_rois"]:
        lab_idx = next((i for i, n in global_label_names.items() if n == lab_str), 0)
        rois.append((int((x + ox) * scale_x), int((y + oy) * scale_y), int(w * scale_x), int(h * scale_y), lab_idx))
    return rois


def rasterize_roi_labels(rois, frames: int, height: int = IMAGE_HEIGHT, width: int = IMAGE_WIDTH, device="cuda") -> torch.Tensor:
    """int64 [frames, height, width]: `lb_arr` of create_image_and_labels_for_dataset (helpers.py:197-206)."""
    return ops.roi_labels(rois, frames, height, width, torch.device(device))


def _grid_boundaries(n: int, g: int):
    """datasets.py:56-58."""
    return [(i * n) // g for i in range(g)] + [n]


def _stripe_map(n: int, g: int) -> np.ndarray:
    """source index per target index of one axis of _shuffle_stripes (datasets.py:71-112): stripes of equal size are
    permuted among themselves, `random.shuffle` once per size group in first-seen order."""
    bounds = _grid_boundaries(n, max(1, int(g)))
    groups = defaultdict(list)
    for i in range(len(bounds) - 1):
        groups[bounds[i + 1] - bounds[i]].append((bounds[i], bounds[i + 1]))
    m = np.arange(n, dtype=np.int32)
    for _, lst in groups.items():
        perm = lst[:]
        random.shuffle(perm)
        for (t0, t1), (s0, s1) in zip(lst, perm):
            m[t0:t1] = np.arange(s0, s1, dtype=np.int32)
    return m


class _IndexMap:
    """out[h][w] = in[A[u]][B[v]], (u, v) = (h, w) or (w, h) when transposed."""

    def __init__(self, h: int, w: int):
        self.a, self.b, self.t, self.h, self.w = np.arange(h, dtype=np.int32), np.arange(w, dtype=np.int32), 0, h, w

    def along_h(self, m):       # out'[h][w] = out[m[h]][w]
        if self.t:
            self.b = self.b[m]
        else:
            self.a = self.a[m]

    def along_w(self, m):       # out'[h][w] = out[h][m[w]]
        if self.t:
            self.a = self.a[m]
        else:
            self.b = self.b[m]

    def flip_h(self):
        self.along_h(np.arange(self.h - 1, -1, -1))

    def flip_w(self):
        self.along_w(np.arange(self.w - 1, -1, -1))

    def transpose(self):
        if self.h != self.w:
            raise NotImplementedError("rot90 by an odd multiple needs square slices (the reference's are 512 x 512)")
        self.t ^= 1

    def rot90(self, k: int):    # torch.rot90(x, k, dims=(-2, -1))
        k %= 4
        if k == 1:
            self.flip_w(); self.transpose()
        elif k == 2:
            self.flip_h(); self.flip_w()
        elif k == 3:
            self.flip_h(); self.transpose()


class TrainGridAug:
    """datasets.py:134-206 on the device, over a batch: random H/V flips, random 90-degree rotations, intensity
    jitter, Gaussian noise, separable grid shuffle with the per-sample grid size `gs`, the bright 32 x 32 stamp.
    `__call__(x, y, gs)`: x [B,1,F,H,W] (or one sample [1,F,H,W]) CUDA fp32, y [B,F,H,W] ([F,H,W]) uint8 / int64 or None,
    gs: None, an int, or one int per sample. Samples are processed in order, each drawing from `random` exactly as
    one reference call would; the noise values come from the kernel's own Philox stream (seeded from torch's CPU
    generator), everything else is bit-identical to the reference for the same `random` state."""

    def __init__(self, gs_choices=(2, 3, 4, 5), p_grid=1.0, flip_p=0.5, rot90_p=0.5, jitter_p=0.3, noise_p=0.3,
                 noise_std=0.01, stamp_top_left=True):
        self.gs_choices = tuple(int(g) for g in gs_choices)
        self.p_grid, self.flip_p, self.rot90_p = float(p_grid), float(flip_p), float(rot90_p)
        self.jitter_p, self.noise_p, self.noise_std = float(jitter_p), float(noise_p), float(noise_std)
        self.stamp = bool(stamp_top_left)

    def _draw(self, h: int, w: int, gs: Optional[int]):
        m = _IndexMap(h, w)
        if random.random() < self.flip_p:
            m.flip_w()
        if random.random() < self.flip_p:
            m.flip_h()
        if random.random() < self.rot90_p:
            m.rot90(random.randint(1, 3))
        scale, shift = 1.0, 0.0
        if random.random() < self.jitter_p:
            scale = 1.0 + 0.1 * (2 * random.random() - 1)
            shift = 0.05 * (2 * random.random() - 1)
        noise = self.noise_std if random.random() < self.noise_p else 0.0
        run_grid = random.random() < self.p_grid
        use_gs = int(gs) if gs is not None else None
        if use_gs is None or use_gs < 1:
            use_gs = random.choice(self.gs_choices) if self.gs_choices else 1
        stamp = 0
        if run_grid and use_gs > 1:
            m.along_h(_stripe_map(m.h, use_gs))      # rows first, then columns (datasets.py:85-112)
            m.along_w(_stripe_map(m.w, use_gs))
            stamp = int(self.stamp)
        return m, scale, shift, noise, stamp

    def __call__(self, x: torch.Tensor, y: Optional[torch.Tensor], gs=None):
        if x.ndim == 3:
            raise NotImplementedError("the 2-D (C,H,W) form is not on the B200 path; pass 3-D samples (1,F,H,W)")
        single = x.ndim == 4
        if single:
            x = x.unsqueeze(0)
            y = y.unsqueeze(0) if y is not None else None
        if x.ndim != 5 or x.shape[1] != 1:
            raise ValueError(f"expected x [B,1,F,H,W], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("TrainGridAug (B200 build) needs CUDA tensors; there is no CPU fallback")
        bsz, _, f, h, w = x.shape
        gss: Sequence = [gs] * bsz if (gs is None or isinstance(gs, int)) else list(gs)
        amap, bmap = np.empty((bsz, h), np.int32), np.empty((bsz, w), np.int32)
        tr, st = np.zeros(bsz, np.int32), np.zeros(bsz, np.int32)
        sc, sh, nz = np.ones(bsz, np.float32), np.zeros(bsz, np.float32), np.zeros(bsz, np.float32)
        for i in range(bsz):
            m, scale, shift, noise, stamp = self._draw(h, w, gss[i])
            amap[i], bmap[i], tr[i], st[i] = m.a, m.b, m.t, stamp
            sc[i], sh[i], nz[i] = scale, shift, noise
        any_noise, any_stamp = bool((nz > 0).any()), bool(st.any())
        seeds = torch.randint(0, 2 ** 62, (bsz,), dtype=torch.int64) if any_noise else torch.zeros(bsz, dtype=torch.int64)
        dev = x.device
        to = lambda a: torch.from_numpy(a).to(dev, non_blocking=True)
        xin = x.float().contiguous().view(bsz, f, h, w)
        xo = torch.empty_like(xin)
        yin = yo = None
        if y is not None:
            yin = y.to(dev)
            if yin.dtype not in (torch.uint8, torch.int64):
                yin = yin.long()
            yin = yin.contiguous()
            yo = torch.empty_like(yin)
        ops.grid_aug(xin, yin, xo, yo, to(amap), to(bmap), to(tr), to(sc), to(sh), to(nz), seeds.to(dev), to(st), any_noise,
                     any_stamp)
        xo = xo.view(bsz, 1, f, h, w)
        if single:
            return xo[0], (yo[0] if yo is not None else None)
        return xo, yo
