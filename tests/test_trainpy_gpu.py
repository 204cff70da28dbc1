"""GPU: the reference's unmodified `train.py` drives this repo's `innovative3D` package, and the
Lightning step surface of the SPCT classes (`_shared_step`, `training_step`, `validation_step`, `apply_unified_loss`:
reference models.py:479-588, unified_loss.py:114-144) is checked against the oracle.

`train.train_and_log` (train.py:1398-1583) is run as is — seed, `DataMod(trainval_sets, ...)`, `_build_lit`, the compute
read-out, `TrainValCSVLogger`, both `ModelCheckpoint`s, `EarlyStopping`, `Trainer(...).fit`, the custom test pass
`write_test_metrics_csv_from_pass` (full sklearn path), `write_summary_csv` — with a synthetic
DataModule in place of the DICOM one and FINAL_EPOCHS cut to 2; then the checkpoint is reloaded the way
`test.py:626-644` does and a second call resumes from `last.ckpt`. Lightning itself is the repo's stand-in when the real
package is absent (tests/_ref_env.py)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _ref_env  # noqa: E402


@pytest.fixture(scope="module")
def env(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("trainpy")
    ref = _ref_env.prepare(tmp)
    return ref, tmp, _ref_env.import_script(ref, "train")


def _synthetic_datamodule():
    from torch.utils.data import DataLoader, Dataset

    from oracle import spff_oracle as O

    class Phantoms(Dataset):
        """(img [1,F,H,W] fp32, lbl [F,H,W] int64) like DicomDataset3D.__getitem__ (datasets.py:227-238)."""

        def __init__(self, n, seed):
            x, lab = O.phantom_batch(n, 64, 64, seed=seed, ignore_frac=0.01)
            self.x, self.lab = x, lab

        def __len__(self):
            return len(self.x)

        def __getitem__(self, i):
            return self.x[i], self.lab[i]

    class SyntheticDataModule:
        """Constructor and loader methods of MultiDicomDataModule3D (datasets.py:280-340)."""

        def __init__(self, configs, batch_size=1, num_frames=5):
            self.configs, self.batch_size, self.num_frames = configs, batch_size, num_frames
            self.stages = []

        def setup(self, stage=None):
            self.stages.append(stage)
            self.train_set, self.val_set, self.test_set = Phantoms(6, 1), Phantoms(2, 2), Phantoms(2, 3)

        def train_dataloader(self):
            return DataLoader(self.train_set, batch_size=self.batch_size, shuffle=True)

        def val_dataloader(self):
            return DataLoader(self.val_set, batch_size=self.batch_size, shuffle=False)

        def test_dataloader(self):
            return DataLoader(self.test_set, batch_size=self.batch_size, shuffle=False)

    return SyntheticDataModule


def test_train_py_runs_end_to_end_and_checkpoint_reloads(env, monkeypatch):
    import pandas as pd
    ref, tmp, train = env
    test_py = _ref_env.import_script(ref, "test")
    DM = _synthetic_datamodule()
    variants = dict((v[0], v) for v in train.VARIANTS)
    name, builder, _, base = variants["SPFF-UNet"]
    monkeypatch.setattr(train, "FINAL_EPOCHS", 2)
    monkeypatch.setattr(train, "FAST_SKIP_VIZ", True)
    monkeypatch.setattr(train, "IMAGE_HEIGHT", 64)     # the compute read-out runs the core on CPU at this size: keep it small
    monkeypatch.setattr(train, "IMAGE_WIDTH", 64)
    out = train.train_and_log(name, builder, DM, base, 42)
    folder = train.CHECKPOINT_DIR / name / "seed42"
    assert (folder / "last.ckpt").is_file()
    best = sorted(folder.glob("best-*.ckpt"))
    assert len(best) == 1, best
    # TrainValCSVLogger rows: one train and one val row per epoch, val_macro_dice monitored (train.py:1438-1458)
    log = pd.read_csv(folder / "logs" / "metrics.csv")
    assert sorted(log["phase"].tolist()) == ["train", "train", "val", "val"]
    assert log[log.phase == "val"]["val_macro_dice"].notna().all() and log[log.phase == "train"]["train_loss"].notna().all()
    assert float(log[log.phase == "train"]["train_loss"].iloc[-1]) < float(log[log.phase == "train"]["train_loss"].iloc[0])
    # the custom test pass of train.py wrote its files from model(x) of this package
    tm = pd.read_csv(folder / "logs" / "test_metrics.csv")
    assert 0.0 <= float(tm["test_macro_dice"].iloc[-1]) <= 1.0 and "test_pr_auc_macro" in tm.columns
    # (train.py:262-331 `write_test_details_3d` compares device predictions with host labels and raises on any CUDA
    # model, the reference's own included; train.py:1549-1556 catches it and goes on - so no test_details.csv here)
    assert out == "DONE" or 0.0 <= float(out) <= 1.0

    # ---- reload like test.py:615-644: prototype from the registry, PL restore, manual fallback ------------------
    ckpt_path = str(folder / "last.ckpt")
    proto = builder()
    ModelCls = proto.__class__
    dev = torch.device("cuda")
    try:      # "PL restore attempt" (test.py:626-630). For this class it raises with real Lightning and the reference's
        # own module too: the saved hyper-parameter `is_3d` travels through **kw into UNet3D_SpectralCore.__init__
        # (models.py:1558-1564 -> :1547-1555 -> :654) - which is why test.py carries the manual fallback below
        restored = ModelCls.load_from_checkpoint(ckpt_path, map_location=dev)
    except TypeError:
        restored = None
    assert restored is None
    ck = torch.load(ckpt_path, map_location=dev, weights_only=False)
    assert {"state_dict", "optimizer_states", "epoch", "global_step", "hyper_parameters"} <= set(ck)
    assert ck["epoch"] == 1 and ck["global_step"] == 12
    sd = test_py._align_state_dict_keys(ck.get("state_dict", ck), proto.state_dict())
    proto.load_state_dict(sd, strict=False)
    manual = proto.to(dev).eval()
    x = _synthetic_datamodule()([], 1, 5)
    x.setup("test")
    xb, _ = next(iter(x.test_dataloader()))
    # a class without **kw restores through the PL path: SP_UNet (models.py:1585-1592)
    sp = dict((v[0], v[1]) for v in train.VARIANTS)["SP_UNet"]()
    sp_path = str(tmp / "sp_unet.ckpt")
    torch.save({"state_dict": sp.state_dict(), "hyper_parameters": dict(sp.hparams)}, sp_path)
    sp2 = type(sp).load_from_checkpoint(sp_path, map_location=dev)
    assert all(torch.equal(a.cpu(), b.cpu()) for a, b in zip(sp.state_dict().values(), sp2.state_dict().values()))
    with torch.no_grad():
        a, b = manual(xb.to(dev)), manual(xb.to(dev))
    assert a.shape == (1, 13, 5, 64, 64) and torch.equal(a, b)
    for k, v in ck["state_dict"].items():                       # the trained weights, not the seed's
        assert torch.equal(manual.state_dict()[k].cpu(), v.cpu()), k
    fresh = builder().to(dev).eval()
    with torch.no_grad():
        assert not torch.equal(fresh(xb.to(dev)), a)
    logits = test_py._extract_logits_from_output(a, prefer_classes=13)
    assert logits is a

    # ---- resume: a second call finds last.ckpt (train.py:504-509, 1509-1516) and runs only the missing epoch ----
    monkeypatch.setattr(train, "FINAL_EPOCHS", 3)
    train.train_and_log(name, builder, DM, base, 42)
    log2 = pd.read_csv(folder / "logs" / "metrics.csv")
    assert sorted(log2["epoch"].astype(int).tolist()) == [0, 0, 1, 1, 2, 2]
    ck2 = torch.load(folder / "last.ckpt", map_location="cpu", weights_only=False)
    assert ck2["epoch"] == 2 and ck2["global_step"] == 18


@pytest.mark.parametrize("variant", ["SPFF-UNet", "PlainCore_UNet", "SP_UNet"])
def test_lightning_step_surface_matches_the_oracle(env, variant):
    """training_step returns the differentiable ce_plus_macro_dice loss and logs the reference's metric names;
    validation_step returns {'val_loss': ...} and logs val_macro_dice (the monitored quantity)."""
    from innovative3D import config as C
    from oracle import spff_oracle as O
    lit = dict((v[0], v[1]) for v in C.VARIANTS)[variant]().cuda()
    x, lab = O.phantom_batch(2, 32, 32, seed=77, ignore_frac=0.02)
    batch = (x.cuda(), lab.cuda())
    loss = lit.training_step(batch, 0)
    assert loss.requires_grad and loss.dim() == 0
    with torch.no_grad():
        logits = lit(batch[0]).cpu()
    want = O.ce_plus_macro_dice_loss(logits, lab)
    assert abs(float(loss) - float(want)) < 1e-5
    mo = O.per_class_metrics_3d(logits, lab, 13, ignore_index=255)
    names = {"train_loss", "train_macro_dice", "train_micro_dice", "train_macro_sens", "train_macro_spec", "train_micro_sens",
             "train_micro_spec"} | {f"train_{m}_class_{i}" for m in ("dice", "sens", "spec") for i in range(13)}
    assert names <= set(lit.logged), sorted(names - set(lit.logged))
    assert abs(float(lit.logged["train_macro_dice"]) - mo[3]) < 1e-12
    np.testing.assert_allclose([float(lit.logged[f"train_dice_class_{i}"]) for i in range(13)], mo[0], rtol=1e-12, equal_nan=True)
    loss.backward()
    g = [p.grad for p in lit.parameters() if p.grad is not None]
    assert len(g) >= 40 and all(torch.isfinite(t).all() for t in g)
    # the same through a dict batch and list-wrapped tensors (models.py:480-481)
    loss2 = lit.training_step({"image": [batch[0]], "label": [batch[1]]}, 0)
    assert abs(float(loss2) - float(loss)) < 1e-6
    lit.eval()
    with torch.no_grad():
        out = lit.validation_step(batch, 0)
    assert set(out) == {"val_loss"} and abs(float(out["val_loss"]) - float(want)) < 1e-5
    assert abs(float(lit.logged["val_macro_dice"]) - mo[3]) < 1e-12
    with torch.no_grad():
        t = lit.test_step(batch, 0)
    assert abs(float(t) - float(want)) < 1e-5 and "test_macro_dice" in lit.logged


def test_apply_unified_loss_patches_every_lit_class(env):
    """unified_loss.apply_unified_loss (unified_loss.py:114-144): every Lightning class of innovative3D.models except
    BaseLitModel gets the unified train/val/test steps; they compute the same loss and log the macro/micro metrics."""
    from innovative3D import models as M
    from innovative3D import unified_loss as U
    from oracle import spff_oracle as O
    saved = {n: (c.training_step, c.validation_step, c.test_step) for n, c in vars(M).items()
             if isinstance(c, type) and issubclass(c, M.pl.LightningModule)}
    try:
        patched = U.apply_unified_loss()
        assert "BaseLitModel" not in patched
        assert {"LitSPCT_EFiLM_FourierGate", "LitSPCT_ControlUNet", "LitSPCT_SEspec", "LitCicek3DUNet_DepthAdapter_Published"} <= set(patched)
        lit = M.LitSPCT_EFiLM_FourierGate().cuda()
        assert lit.training_step.__func__ is U._training_step
        x, lab = O.phantom_batch(2, 32, 32, seed=78, ignore_frac=0.02)
        # labels as [B,1,F,H,W] exercise the canonicalisation (unified_loss.py:56-75, models.py:68-83)
        loss = lit.training_step((x.cuda(), lab.cuda().unsqueeze(1)), 0)
        with torch.no_grad():
            logits = lit(x.cuda()).cpu()
        assert abs(float(loss) - float(O.ce_plus_macro_dice_loss(logits, lab))) < 1e-5 and loss.requires_grad
        mo = O.per_class_metrics_3d(logits, lab, 13, ignore_index=255)
        assert abs(float(lit.logged["train_macro_dice"]) - mo[3]) < 1e-12
        assert abs(float(lit.logged["train_micro_dice"]) - mo[6]) < 1e-12
        with torch.no_grad():
            v = lit.validation_step({"image": x.cuda(), "label": lab.cuda()}, 0)
        assert abs(float(v) - float(loss)) < 1e-5 and "val_macro_dice" in lit.logged
    finally:
        for n, (a, b, c) in saved.items():
            cls = getattr(M, n)
            cls.training_step, cls.validation_step, cls.test_step = a, b, c
