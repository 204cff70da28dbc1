"""GPU: the data-path kernels (spff_roi_labels, spff_grid_aug) through innovative3D.datapath_gpu of this tree against the
reference fixtures (tests/golden/datapath.npz) and the CPU oracle — bit-exact for labels, indices, jitter and stamp;
the Gaussian noise (a different generator) statistically."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "datapath.npz"))
AUG = sorted({k.split("_")[0] for k in GOLD.files if k.startswith("aug")}, key=lambda s: int(s[3:]))
ROI = sorted({k.split("_")[0] for k in GOLD.files if k.startswith("roi")})


@pytest.mark.parametrize("name", AUG)
def test_aug_matches_reference_fixture(name):
    from innovative3D.datapath_gpu import TrainGridAug
    from oracle import datapath_oracle as DO
    seed, f, h, w, gs = [int(v) for v in GOLD[name + "_case"]]
    gs = None if gs < 0 else gs
    x, y = DO.aug_input(seed, f, h, w)
    aug = TrainGridAug(noise_p=0.0, rot90_p=0.5 if h == w else 0.0)
    random.seed(1000 + seed)
    xo, yo = aug(x.cuda(), y.cuda(), gs)          # single-sample form, as the reference calls it
    assert random.random() == float(GOLD[name + "_next"])
    assert xo.shape == (1, f, h, w) and np.array_equal(xo.cpu().numpy(), GOLD[name + "_x"])
    assert np.array_equal(yo.cpu().numpy(), GOLD[name + "_y"])


@pytest.mark.parametrize("labels_dtype", [torch.int64, torch.uint8])
def test_aug_batch_equals_oracle_sample_by_sample(labels_dtype):
    """A batch of native-size slices: sample i of the batched device call == the i-th oracle call on the same `random` stream."""
    from innovative3D.datapath_gpu import TrainGridAug
    from oracle import datapath_oracle as DO
    bsz, f, h, w = 6, 5, 512, 512
    xs, ys = zip(*[DO.aug_input(50 + i, f, h, w) for i in range(bsz)])
    x, y = torch.stack(xs), torch.stack(ys)
    gss = [2, 3, 4, 5, None, 5]
    random.seed(77)
    want = [DO.train_grid_aug(x[i].clone(), y[i].clone(), gss[i], noise_p=0.0) for i in range(bsz)]
    after = random.random()
    random.seed(77)
    xo, yo = TrainGridAug(noise_p=0.0)(x.cuda(), y.to(labels_dtype).cuda(), gss)
    assert random.random() == after
    for i in range(bsz):
        assert torch.equal(xo[i].cpu(), want[i][0]), i
        assert torch.equal(yo[i].cpu().long(), want[i][1]), i


def test_aug_noise_statistics():
    from innovative3D.datapath_gpu import TrainGridAug
    torch.manual_seed(0)
    x = torch.randn(4, 1, 5, 128, 128, device="cuda") * 2.0
    aug = TrainGridAug(p_grid=0.0, flip_p=0.0, rot90_p=0.0, jitter_p=0.0, noise_p=1.0, noise_std=0.01)
    random.seed(1)
    xo, _ = aug(x, None)
    d = (xo - x).flatten(1)
    assert torch.allclose(d.mean(1), torch.zeros(4, device="cuda"), atol=2e-4)
    assert torch.allclose(d.std(1), torch.full((4,), 0.01, device="cuda"), rtol=2e-2)      # min(0.01, 0.25 * 2.0)
    assert abs(float(torch.corrcoef(torch.stack([d[0], d[1]]))[0, 1])) < 0.02             # per-sample streams differ
    # amplitude capped by a quarter of the sample's own std
    small = x * 0.01
    xo2, _ = aug(small, None)
    assert torch.allclose((xo2 - small).flatten(1).std(1), 0.25 * small.flatten(1).std(1), rtol=3e-2)
    kurt = float(((d[0] / d[0].std()) ** 4).mean())
    assert 2.9 < kurt < 3.1                                                                 # Gaussian


@pytest.mark.parametrize("name", ROI)
def test_roi_labels_match_reference_fixture(name):
    from innovative3D.datapath_gpu import rasterize_roi_labels
    rois = [tuple(int(v) for v in r) for r in GOLD[name + "_rois"]]
    want = GOLD[name + "_labels"]
    got = rasterize_roi_labels(rois, want.shape[0], want.shape[1], want.shape[2])
    assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), want)


def test_roi_labels_native_size_equal_oracle_and_errors():
    from innovative3D.datapath_gpu import rasterize_roi_labels, scaled_rois
    from oracle import datapath_oracle as DO
    cfg = {"offset": (10, -5), "original_rois": [(200, 220, 160, 150, "c1"), (500, 300, 170, 165, "c2"), (820, 760, 140, 160, "c3"),
                                                 (420, 640, 150, 150, "c4"), (640, 900, 155, 145, "c5"), (480, 310, 80, 300, "c6")]}
    rois = scaled_rois(cfg)
    rois = [(x, y, w, h, 1 + i) for i, (x, y, w, h, _) in enumerate(rois)]     # distinct labels whatever the name table holds
    got = rasterize_roi_labels(rois, 5, 512, 512)
    assert np.array_equal(got.cpu().numpy(), DO.roi_labels(rois, 5, 512, 512))
    assert int((got > 0).sum()) > 5 * 5000
    with pytest.raises(RuntimeError):
        rasterize_roi_labels([(500, 10, 40, 40, 1)], 1, 64, 64)               # IndexError in the reference
