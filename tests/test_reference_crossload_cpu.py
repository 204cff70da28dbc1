"""CPU, build container only (skipped where /root/reference is absent): the drop-in contract against the LIVE reference —
state_dicts cross-load strictly in both directions for every variant, and the reference's own `train.py`-style builder
loop (`_build_lit`, train.py:1251-1271: call the zero-arg builder, else try (num_classes, lr) / (num_classes) / ()) works
on this tree's VARIANTS. No compute runs (there is no CPU path); this is registry + parameter-surface parity."""
import os
import sys

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "innovative3D")), reason="reference tree not mounted")

REF_CLASS = {"SPFF-UNet": "LitSPCT_EFiLM_FourierGate", "E_SP_UNet": "LitSPCT_EnergyFiLM", "FG_SP_UNet": "LitSPCT_FourierGate",
             "SP_UNet": "LitSPCT_SEspec", "PlainCore_UNet": "LitSPCT_ControlUNet"}


@pytest.fixture(scope="module")
def reference_models():
    """The reference's innovative3D.models, imported under the stub shim in a way that leaves this tree's
    `innovative3D` package importable afterwards (the two share a package name)."""
    mine = {k: v for k, v in sys.modules.items() if k == "innovative3D" or k.startswith("innovative3D.")}
    for k in mine:
        del sys.modules[k]
    path0 = list(sys.path)
    from oracle.make_golden import import_reference
    M, _ = import_reference()
    ref_mods = {k: v for k, v in sys.modules.items() if k == "innovative3D" or k.startswith("innovative3D.")}
    for k in ref_mods:
        del sys.modules[k]
    sys.path[:] = path0
    sys.modules.update(mine)
    yield M
    for name in ("pytorch_lightning", "pytorch_lightning.callbacks", "pytorch_lightning.loggers", "torchmetrics", "matplotlib",
                 "matplotlib.pyplot", "matplotlib.patches", "seaborn", "pydicom"):
        sys.modules.pop(name, None)


def _mine():
    from innovative3D import config as C
    return dict((v[0], v[1]) for v in C.VARIANTS)


@pytest.mark.parametrize("variant", list(REF_CLASS) + ["3DUNet"])
def test_state_dicts_cross_load(reference_models, variant):
    M = reference_models
    torch.manual_seed(1)
    if variant == "3DUNet":
        ref = M.LitCicek3DUNet_DepthAdapter_Published(num_classes=13, target_depth=16)
    else:
        ref = getattr(M, REF_CLASS[variant])()
    torch.manual_seed(2)
    mine = _mine()[variant]()
    sd_ref, sd_mine = ref.state_dict(), mine.state_dict()
    assert list(sd_ref) == list(sd_mine)                       # same keys in the same order
    assert all(sd_ref[k].shape == sd_mine[k].shape and sd_ref[k].dtype == sd_mine[k].dtype for k in sd_ref)
    mine.load_state_dict(sd_ref, strict=True)
    assert all(torch.equal(mine.state_dict()[k], sd_ref[k]) for k in sd_ref)
    torch.manual_seed(3)
    other = _mine()[variant]()
    ref.load_state_dict(other.state_dict(), strict=True)
    assert all(torch.equal(ref.state_dict()[k], other.state_dict()[k]) for k in sd_ref)
    assert sum(p.numel() for p in ref.parameters()) == sum(p.numel() for p in mine.parameters())


def test_train_py_builder_loop_on_this_trees_variants():
    """train.py:1251-1271 `_build_lit`, restated: a VARIANTS entry is a zero-arg factory or a class tried with
    (num_classes, lr), (num_classes), ()."""
    from innovative3D import config as C

    def build_lit(builder):
        try:
            return builder()
        except TypeError:
            for args in ((C.NUM_CLASSES, C.BEST_LR), (C.NUM_CLASSES,), ()):
                try:
                    return builder(*args)
                except TypeError:
                    continue
            raise

    for name, builder, dm, ckpt in C.VARIANTS[:len(C.B200_VARIANT_NAMES)]:     # the B200 variants; the rest are the reference's own
        lit = build_lit(builder)
        assert lit.hparams.num_classes == C.NUM_CLASSES and hasattr(lit, "training_step") and hasattr(lit, "configure_optimizers")
        assert hasattr(lit, "model") and isinstance(lit.model, torch.nn.Module)
        assert str(ckpt).endswith(name)
