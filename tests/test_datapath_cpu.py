"""CPU: pin oracle/datapath_oracle.py to the reference's own TrainGridAug / is_pixel_in_ellipse outputs
(tests/golden/datapath.npz) and check the host logic of this tree's TrainGridAug — the composition of flips, rot90 and
the stripe shuffle into two index tables, drawn in the reference's order — without any device call."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import datapath_oracle as DO

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "datapath.npz"))
AUG = sorted({k.split("_")[0] for k in GOLD.files if k.startswith("aug")}, key=lambda s: int(s[3:]))
ROI = sorted({k.split("_")[0] for k in GOLD.files if k.startswith("roi")})


def _case(name):
    seed, f, h, w, gs = [int(v) for v in GOLD[name + "_case"]]
    return seed, f, h, w, (None if gs < 0 else gs)


@pytest.mark.parametrize("name", AUG)
def test_oracle_aug_matches_reference(name):
    seed, f, h, w, gs = _case(name)
    x, y = DO.aug_input(seed, f, h, w)
    random.seed(1000 + seed)
    xo, yo = DO.train_grid_aug(x.clone(), y.clone(), gs, noise_p=0.0, rot90_p=0.5 if h == w else 0.0)
    assert np.array_equal(xo.numpy(), GOLD[name + "_x"]) and np.array_equal(yo.numpy(), GOLD[name + "_y"])
    assert random.random() == float(GOLD[name + "_next"])


@pytest.mark.parametrize("name", ROI)
def test_oracle_roi_labels_match_reference(name):
    rois = [tuple(int(v) for v in r) for r in GOLD[name + "_rois"]]
    want = GOLD[name + "_labels"]
    assert np.array_equal(DO.roi_labels(rois, want.shape[0], want.shape[1], want.shape[2]), want)


@pytest.mark.parametrize("name", AUG)
def test_host_index_tables_reproduce_the_reference(name):
    """TrainGridAug._draw (this tree): same `random` draws, and gathering with its tables == the reference's output."""
    from innovative3D.datapath_gpu import TrainGridAug
    seed, f, h, w, gs = _case(name)
    x, y = DO.aug_input(seed, f, h, w)
    aug = TrainGridAug(noise_p=0.0, rot90_p=0.5 if h == w else 0.0)
    random.seed(1000 + seed)
    m, scale, shift, noise, stamp = aug._draw(h, w, gs)
    assert random.random() == float(GOLD[name + "_next"]) and noise == 0.0
    hh, ww = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    u, v = (ww, hh) if m.t else (hh, ww)
    xs = x.numpy()[0][:, m.a[u], m.b[v]]
    if not (scale == 1.0 and shift == 0.0):
        xs = (xs * np.float32(scale)).astype(np.float32) + np.float32(shift)
    ys = y.numpy()[:, m.a[u], m.b[v]]
    if stamp:
        xs[0, :32, :32] = xs[0, :32, :32].max() + np.float32(max(np.abs(xs).max(), np.float32(1.0))) * np.float32(0.25)
    assert np.array_equal(xs.astype(np.float32), GOLD[name + "_x"][0]) and np.array_equal(ys, GOLD[name + "_y"])


def test_grid_boundaries_ragged():
    from innovative3D.datapath_gpu import _grid_boundaries
    assert _grid_boundaries(512, 5) == [0, 102, 204, 307, 409, 512]     # the example in datasets.py:57
    assert _grid_boundaries(7, 3) == DO.grid_boundaries(7, 3)


@pytest.mark.parametrize("seed", range(40))
def test_host_index_tables_random_draws_match_the_oracle(seed):
    """Many random decision sequences (flips, rot90, jitter, stripe shuffle with random grid sizes, ragged stripes):
    gathering with this tree's index tables == the oracle's tensor-by-tensor TrainGridAug on the same `random` stream."""
    from innovative3D.datapath_gpu import TrainGridAug
    rng = random.Random(seed)
    square = seed % 3 != 0
    h = rng.choice([24, 40, 56, 64])
    w = h if square else rng.choice([16, 48, 72])
    gs = rng.choice([None, 1, 2, 3, 4, 5, 7])
    x, y = DO.aug_input(500 + seed, 3, h, w)
    kw = dict(noise_p=0.0, rot90_p=0.5 if square else 0.0)
    random.seed(9000 + seed)
    xo, yo = DO.train_grid_aug(x.clone(), y.clone(), gs, **kw)
    nxt = random.random()
    random.seed(9000 + seed)
    m, scale, shift, noise, stamp = TrainGridAug(**kw)._draw(h, w, gs)
    assert random.random() == nxt
    hh, ww = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    u, v = (ww, hh) if m.t else (hh, ww)
    xs = x.numpy()[0][:, m.a[u], m.b[v]]
    if not (scale == 1.0 and shift == 0.0):
        xs = (xs * np.float32(scale)).astype(np.float32) + np.float32(shift)
    if stamp:
        xs[0, :32, :32] = xs[0, :32, :32].max() + np.float32(max(np.abs(xs).max(), np.float32(1.0))) * np.float32(0.25)
    assert np.array_equal(xs.astype(np.float32), xo.numpy()[0]) and np.array_equal(y.numpy()[:, m.a[u], m.b[v]], yo.numpy())
