"""CPU: pin oracle/cicek_oracle.py (the "3DUNet" control: Cicek 3D U-Net + depth adapter, BatchNorm, CE, SGD)
against outputs of the reference itself (tests/golden/cicek*.npz, written by oracle/make_golden_3dunet.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import cicek_oracle as CO
from oracle import spff_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(glob.glob(os.path.join(GOLD, "cicek[0-9]*.npz")))


def _load(path):
    z = np.load(path, allow_pickle=False)
    b, h, w, ign, seed = [str(v) for v in z["case"]]
    return z, int(b), int(h), int(w), float(ign), int(seed)


def test_fixtures_exist():
    assert len(CASES) >= 3


def test_parameter_surface_matches_reference():
    z = np.load(os.path.join(GOLD, "cicek_init_seed42.npz"))
    ref = {k: tuple(int(v) for v in z[k][:-2]) for k in z.files}
    assert CO.param_shapes() == ref
    assert sum(int(np.prod(s)) for k, s in ref.items() if not CO.is_buffer(k)) == 22_578_669


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p) for p in CASES])
def test_oracle_matches_reference(path):
    z, b, h, w, ign, seed = _load(path)
    x, lab = O.phantom_batch(b, h, w, seed=seed, ignore_frac=ign)
    p = CO.det_weights(seed=42)
    loss, logits, grads, stats = CO.loss_and_grads(p, x, lab)
    ref = torch.from_numpy(z["logits"])
    assert logits.shape == ref.shape
    assert float((logits - ref).norm() / ref.norm()) < 1e-5
    assert abs(loss - float(z["loss"])) < 1e-5
    names = [str(n) for n in z["grad_names"]]
    assert sorted(names) == sorted(grads)
    for name, gn in zip(names, z["grad_norms"]):
        g = grads[name].double().reshape(-1)
        assert abs(float(g.norm()) - gn) <= 2e-3 * gn + 1e-7, name
        step = max(1, g.numel() // 512)
        np.testing.assert_allclose(g[::step][:512].float().numpy(), z["g|" + name], rtol=5e-3,
                                   atol=5e-3 * (gn / max(1.0, g.numel() ** 0.5)) + 1e-7, err_msg=name)
    for name in [str(n) for n in z["buf_names"]]:
        np.testing.assert_allclose(stats[name].numpy(), z["b|" + name], rtol=1e-4, atol=1e-6, err_msg=name)
    # two SGD steps (momentum buffer initialised by the first)
    p1, bufs = CO.sgd_step({k: v for k, v in p.items() if not CO.is_buffer(k)}, grads, None)
    np.testing.assert_allclose([float(p1[k].double().norm()) for k in names], z["p1_norms"], rtol=1e-5)
    loss2, _, grads2, _ = CO.loss_and_grads({**p, **p1, **stats}, x, lab)
    assert abs(loss2 - float(z["loss2"])) < 2e-4
    p2, _ = CO.sgd_step(p1, grads2, bufs)
    np.testing.assert_allclose([float(p2[k].double().norm()) for k in names], z["p2_norms"], rtol=1e-4)
    # eval mode: running statistics
    with torch.no_grad():
        le = CO.forward(p, x, training=False)
    ref_e = torch.from_numpy(z["logits_eval"])
    assert float((le - ref_e).norm() / ref_e.norm()) < 1e-5


def test_depth_matrix_is_the_trilinear_resize():
    """With H, W unchanged F.interpolate(trilinear, align_corners=False) is a [Dout x Din] matrix over the planes."""
    torch.manual_seed(0)
    x = torch.randn(2, 3, 5, 4, 6)
    for din, dout in ((5, 16), (16, 5)):
        t = torch.randn(2, 3, din, 4, 6)
        m = CO.depth_matrix(din, dout)
        assert torch.allclose(m.sum(1), torch.ones(dout), atol=1e-6)
        assert torch.allclose(torch.einsum("od,ncdhw->ncohw", m, t), CO.resize_depth(t, dout), atol=1e-6)
    # head (1x1x1 conv + bias) commutes with the resize: the engine resamples the 32-channel activations instead of
    # the 13-channel logits
    wgt, bias = torch.randn(13, 3, 1, 1, 1), torch.randn(13)
    t = torch.randn(2, 3, 16, 4, 6)
    a = CO.resize_depth(torch.nn.functional.conv3d(t, wgt, bias), 5)
    b = torch.nn.functional.conv3d(CO.resize_depth(t, 5), wgt, bias)
    assert torch.allclose(a, b, atol=1e-5)
