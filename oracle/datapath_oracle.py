"""CPU oracle for the data path — TEST INFRASTRUCTURE, not product code.

Restates, in plain Python / numpy / torch-CPU:
  roi_labels        the per-pixel ROI loop of create_image_and_labels_for_dataset   reference innovative3D/helpers.py:197-206
                    with is_pixel_in_ellipse                                          helpers.py:125-129
  train_grid_aug    TrainGridAug.__call__ for 3-D samples                            datasets.py:134-206
                    with _shuffle_stripes / _grid_boundaries                         datasets.py:56-121
Pinned against the reference's own functions by oracle/make_golden_datapath.py -> tests/golden/datapath.npz.
"""
from __future__ import annotations

import random
from collections import defaultdict

import numpy as np
import torch


def is_pixel_in_ellipse(x, y, roi):
    """helpers.py:125-129 (Python floats = doubles)."""
    cx, cy = roi[0] + roi[2] / 2, roi[1] + roi[3] / 2
    a, b = roi[2] / 2, roi[3] / 2
    return ((x - cx) ** 2) / (a * a) + ((y - cy) ** 2) / (b * b) <= 1


def roi_labels(rois, frames: int, height: int, width: int) -> np.ndarray:
    """lb_arr of helpers.py:197-206: int64 [frames, height, width]; later rois overwrite earlier ones."""
    lb = np.zeros((frames, height, width), dtype=np.int64)
    for f in range(frames):
        for (x0, y0, w0, h0, lab) in rois:
            for px in range(x0, x0 + w0):
                for py in range(y0, y0 + h0):
                    if is_pixel_in_ellipse(px, py, (x0, y0, w0, h0)):
                        lb[f, py, px] = lab
    return lb


def grid_boundaries(n: int, g: int):
    return [(i * n) // g for i in range(g)] + [n]


def shuffle_stripes(x: torch.Tensor, y, g_rows: int, g_cols: int):
    """datasets.py:60-121."""
    if g_rows <= 1 and g_cols <= 1:
        return x, y
    H, W = x.shape[-2], x.shape[-1]
    hs, ws = grid_boundaries(H, max(1, int(g_rows))), grid_boundaries(W, max(1, int(g_cols)))
    row_groups, col_groups = defaultdict(list), defaultdict(list)
    for i in range(len(hs) - 1):
        row_groups[hs[i + 1] - hs[i]].append((hs[i], hs[i + 1]))
    for j in range(len(ws) - 1):
        col_groups[ws[j + 1] - ws[j]].append((ws[j], ws[j + 1]))
    xr, yr = x.clone(), (y.clone() if y is not None else None)
    for _, lst in row_groups.items():
        perm = lst[:]
        random.shuffle(perm)
        for (t0, t1), (s0, s1) in zip(lst, perm):
            xr[..., t0:t1, :] = x[..., s0:s1, :]
            if yr is not None:
                yr[..., t0:t1, :] = y[..., s0:s1, :]
    xc, yc = xr.clone(), (yr.clone() if yr is not None else None)
    for _, lst in col_groups.items():
        perm = lst[:]
        random.shuffle(perm)
        for (t0, t1), (s0, s1) in zip(lst, perm):
            xc[..., :, t0:t1] = xr[..., :, s0:s1]
            if yc is not None:
                yc[..., :, t0:t1] = yr[..., :, s0:s1]
    return xc, yc


def train_grid_aug(x: torch.Tensor, y, gs, gs_choices=(2, 3, 4, 5), p_grid=1.0, flip_p=0.5, rot90_p=0.5, jitter_p=0.3,
                   noise_p=0.3, noise_std=0.01, stamp=True):
    """TrainGridAug.__call__ (datasets.py:159-203) for x (1,F,H,W), y (F,H,W)."""
    if random.random() < flip_p:
        x = torch.flip(x, dims=(-1,))
        y = torch.flip(y, dims=(-1,)) if y is not None else None
    if random.random() < flip_p:
        x = torch.flip(x, dims=(-2,))
        y = torch.flip(y, dims=(-2,)) if y is not None else None
    if random.random() < rot90_p:
        k = random.randint(1, 3)
        x = torch.rot90(x, k, dims=(-2, -1))
        y = torch.rot90(y, k, dims=(-2, -1)) if y is not None else None
    if random.random() < jitter_p:
        scale = 1.0 + 0.1 * (2 * random.random() - 1)
        shift = 0.05 * (2 * random.random() - 1)
        x = x * scale + shift
    if random.random() < noise_p:
        v = x.detach().std().item()
        if v > 0:
            x = x + torch.randn_like(x) * min(noise_std, 0.25 * v)
    run_grid = random.random() < p_grid
    use_gs = int(gs) if gs is not None else None
    if use_gs is None or use_gs < 1:
        use_gs = random.choice(tuple(gs_choices)) if gs_choices else 1
    if run_grid and use_gs > 1:
        x, y = shuffle_stripes(x, y, use_gs, use_gs)
        if stamp:
            x[0, 0, :32, :32] = x[0, 0, :32, :32].max() + x.abs().max().clamp(min=1.0) * 0.25
    return x, y


def aug_input(seed: int, frames: int, h: int, w: int):
    """Seeded sample: x (1,F,H,W) fp32 with structure, y (F,H,W) int64 in 0..12 plus some 255."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, frames, h, w, generator=g) * 0.4 + torch.linspace(0, 1, w)[None, None, None, :]
    y = torch.randint(0, 13, (frames, h, w), generator=g)
    y[torch.rand(frames, h, w, generator=g) < 0.02] = 255
    return x, y
