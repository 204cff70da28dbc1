"""The plugin boundary as the reference's own callers see it, no GPU needed: with this repo's
`innovative3D` package on the path next to a reference checkout, `train.py` imports, finds every registry entry, builds
the data module the way `train.py:1410` does and builds every B200 Lightning module the way `train.py:1251-1271` does.
Covers the name collision of round 1 (`innovative3D.datasets` of this tree shadowing the reference's data layer)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _ref_env  # noqa: E402

B200 = ["SPFF-UNet", "E_SP_UNet", "FG_SP_UNet", "SP_UNet", "PlainCore_UNet", "3DUNet"]


@pytest.fixture(scope="module")
def train_mod(tmp_path_factory):
    ref = _ref_env.prepare(tmp_path_factory.mktemp("boundary"))
    return _ref_env.import_script(ref, "train")


def test_package_shadows_hot_path_and_falls_through_for_the_rest(train_mod):
    import innovative3D
    import innovative3D.config as C
    import innovative3D.helpers as H
    import innovative3D.models as M
    here = str(_ref_env.PKG)
    assert innovative3D.__file__.startswith(here) and C.__file__.startswith(here) and M.__file__.startswith(here)
    # hot-path names are this tree's
    assert M.LitSPCT_EFiLM_FourierGate.__module__ == "innovative3D.models"
    assert H.ce_plus_macro_dice_loss.__module__ == "innovative3D.helpers"
    # the data layer is the reference's: its module, its tables, its ingest helpers
    import innovative3D.datasets as D
    assert not D.__file__.startswith(here) and hasattr(D, "MultiDicomDataModule3D") and hasattr(D, "DicomDataset3D")
    assert len(C.trainval_sets) == 4 and len(C.test_set) == 1 and "original_rois" in C.trainval_sets[0]   # config.py:117-124
    assert callable(H.create_image_and_labels_for_dataset) and callable(H.generate_cumulative_grid_sizes)
    assert H.create_image_and_labels_for_dataset.__module__ == "innovative3D._reference_helpers"
    with pytest.raises(AttributeError, match="B200 hot-path build"):
        C.no_such_name
    # this tree's GPU data path lives under its own name
    from innovative3D.datapath_gpu import TrainGridAug   # noqa: F401


def test_train_py_sees_every_variant_and_builds_datamodule_and_module(train_mod):
    names = [v[0] for v in train_mod.VARIANTS]
    assert names[:6] == B200
    # the four other model families stay registered, with the reference's own builders (config.py:316-391)
    assert {"UNETR", "ResUNet++"} <= set(names) and len(names) == 10, names
    for name, builder, DataMod, base in train_mod.VARIANTS:
        dm = DataMod(train_mod.trainval_sets, batch_size=train_mod.BATCH_SIZE, num_frames=train_mod.NUM_FRAMES)   # train.py:1410
        assert type(dm).__name__ == "MultiDicomDataModule3D" and dm.batch_size == 1 and len(dm.configs) == 4
        if name not in B200:
            continue
        lit = train_mod._build_lit(builder)                                                                      # train.py:1413
        assert type(lit).__module__ == "innovative3D.models" and hasattr(lit, "model")
        assert int(lit.hparams.num_classes) == 13
        for meth in ("training_step", "validation_step", "test_step", "configure_optimizers", "forward"):
            assert callable(getattr(lit, meth))
        opt = lit.configure_optimizers()
        assert hasattr(opt, "param_groups") or isinstance(opt, (list, tuple)) or "optimizer" in opt
        keys = lit.state_dict().keys()
        assert len(keys) > 10 and (name == "3DUNet" or any(k.startswith("model.") for k in keys))


def test_other_families_resolve_to_the_reference_code(train_mod):
    import innovative3D.models as M
    cls = M.R2UNet3D if hasattr(M, "R2UNet3D") else None
    ref_models = sys.modules.get("innovative3D._reference_models")
    assert ref_models is not None, "models.py fall-through did not load the reference module"
    assert cls is None or cls.__module__ == "innovative3D._reference_models"
    assert M.LitSPCT_ControlUNet.__module__ == "innovative3D.models"      # never the reference's class of the same name


def test_cpu_forward_is_loud_not_a_fallback(train_mod):
    import torch
    lit = train_mod._build_lit(dict((v[0], v[1]) for v in train_mod.VARIANTS)["SPFF-UNet"])
    with pytest.raises(RuntimeError, match="no CPU"):
        lit(torch.zeros(1, 1, 5, 16, 16))
