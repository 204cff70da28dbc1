"""Constants and the `VARIANTS` plugin registry of the hot path.

Mirrors the part of the reference's `innovative3D/config.py` that the callers of the hot path
import (`train.py:64-78,89`, `test.py:62-72`): the label space, the training constants
(config.py:21-33) and `VARIANTS` — a list of `(name, zero-arg builder, DataModule, ckpt_dir)`
(config.py:271-280) whose builders construct the Lightning modules of `innovative3D.models`
(config.py:410-476). Dataset geometry (ROI tables, DICOM roots, `dataset_configs`, `trainval_sets`,
`test_set`, the VMI switches: config.py:36-124, 232-248) is the data layer's and is NOT restated here:
any name this module does not define is handed through to the reference's `config.py` when a
reference checkout is on `sys.path` (`_fallthrough.py`), so `train.py:64-78`, `test.py:62-72` and the
reference's `datasets.py:24-35` import what they need from `innovative3D.config` unchanged. Without a
reference checkout the three dataset lists read as empty lists and other data-layer names raise
AttributeError. Unlike the reference, importing this module creates no directories under a hard-coded
home path (config.py:15-19) — only CHECKPOINT_DIR / LOG_DIR are used, both overridable through the same
environment variables (config.py:252-253).
"""
import inspect
import os
from importlib import import_module
from pathlib import Path

IMAGE_HEIGHT, IMAGE_WIDTH = 512, 512   # config.py:21
NUM_FRAMES = 5                         # config.py:22
NUM_CLASSES = 13                       # config.py:23
FINAL_EPOCHS = 200
BEST_LR = 1e-4                         # config.py:25
IGNORE_INDEX = 255                     # config.py:26
BATCH_SIZE = 1                         # config.py:27
NUM_WORKERS = 16
num_workers = NUM_WORKERS
grid_size = 10
SEEDS = [42, 123, 999]                 # config.py:33

global_label_names = {
    0: "BG", 1: "HA800", 2: "HA400", 3: "HA200", 4: "HA100", 5: "Lung", 6: "Liver", 7: "Adipose",
    8: "Water", 9: "I15", 10: "I10", 11: "I5", 12: "HA50",
}
label_colors = {
    0: (0, 0, 0), 1: (255, 0, 0), 2: (255, 127, 0), 3: (255, 255, 0), 4: (0, 255, 0), 5: (0, 255, 255),
    6: (0, 0, 255), 7: (139, 69, 19), 8: (255, 255, 255), 9: (255, 0, 255), 10: (128, 0, 128),
    11: (0, 128, 128), 12: (128, 128, 0),
}

LOSS_NAME = "ce_plus_macro_dice"
FOCAL_ALPHA, FOCAL_GAMMA, GRAD_WEIGHT = 0.25, 2.0, 1.0
USE_VMI = False

_PROJECT_ROOT = Path(__file__).resolve().parents[1]
CHECKPOINT_DIR = Path(os.getenv("CHECKPOINT_DIR", str(_PROJECT_ROOT / "runs" / "checkpoints"))).resolve()
LOG_DIR = Path(os.getenv("LOG_DIR", str(_PROJECT_ROOT / "runs"))).resolve()
CKPT_DIR = CHECKPOINT_DIR


def _data_module(name):
    """The reference's DICOM data modules (datasets.py:280-364, 367-422) are the caller's side of the boundary:
    `innovative3D.datasets` is not a module of this tree, it resolves to the reference checkout's file through the
    extended package path (`__init__.py`), exactly like config.py:130-142 resolves it lazily."""
    def factory(*args, **kwargs):
        try:
            mod = import_module("innovative3D.datasets")
        except ImportError as e:
            raise ImportError("innovative3D.datasets (the reference's CPU data pipeline: DICOM ingest, DataModules) is "
                              "outside the B200 hot path and not part of this tree; run with the reference checkout on "
                              "sys.path (e.g. from its directory) so that its datasets.py is found") from e
        return getattr(mod, name)(*args, **kwargs)
    factory.__name__ = factory.__qualname__ = name
    return factory


MultiDicomDataModule3D = _data_module("MultiDicomDataModule3D")
MultiDicomDataModule2D = _data_module("MultiDicomDataModule2D")


def build_class(class_name: str, **ctor_kwargs):
    """Zero-arg factory for `innovative3D.models.<class_name>`, passing only the kwargs its
    constructor accepts (same contract as config.py:159-182)."""
    def _factory():
        cls = getattr(import_module("innovative3D.models"), class_name, None)
        if cls is None:
            raise ImportError(f"[config] {class_name} not found in innovative3D.models")
        params = inspect.signature(cls.__init__).parameters
        if any(p.kind == inspect.Parameter.VAR_KEYWORD for p in params.values()):
            return cls(**ctor_kwargs)
        return cls(**{k: v for k, v in ctor_kwargs.items() if k in params})
    return _factory


VARIANTS = []


def _add_variant(name, builder_or_class, dm_cls, ckpt_dir):
    VARIANTS.append((name, builder_or_class, dm_cls, Path(ckpt_dir)))


# shared constructor arguments of the SPCT family (config.py:410-419)
_SPCT_COMMON = dict(num_classes=NUM_CLASSES, lr=BEST_LR, base=32, ksd=3, use_se=True, use_specse=True,
                    use_spatial=False, use_skip_gate=False)

_add_variant("SPFF-UNet", build_class("LitSPCT_EFiLM_FourierGate", **_SPCT_COMMON), MultiDicomDataModule3D,
             CHECKPOINT_DIR / "SPFF-UNet")                                         # config.py:423-428
_add_variant("E_SP_UNet", build_class("LitSPCT_EnergyFiLM", **_SPCT_COMMON), MultiDicomDataModule3D,
             CHECKPOINT_DIR / "E_SP_UNet")                                         # config.py:433-438
_add_variant("FG_SP_UNet", build_class("LitSPCT_FourierGate", **_SPCT_COMMON), MultiDicomDataModule3D,
             CHECKPOINT_DIR / "FG_SP_UNet")                                        # config.py:443-448
_add_variant("SP_UNet", build_class("LitSPCT_SEspec", num_classes=NUM_CLASSES, lr=BEST_LR), MultiDicomDataModule3D,
             CHECKPOINT_DIR / "SP_UNet")                                           # config.py:451-456
_add_variant("PlainCore_UNet",
             build_class("LitSPCT_ControlUNet", **{**_SPCT_COMMON, "use_se": False, "use_specse": False}),
             MultiDicomDataModule3D, CHECKPOINT_DIR / "PlainCore_UNet")            # config.py:460-476



def make_cicek_depth_adapter_sgd_wce():
    """config.py:283-303: Cicek 3D U-Net + depth adapter, SGD like the paper, plain softmax CE."""
    from innovative3D.models import LitCicek3DUNet_DepthAdapter_Published
    return LitCicek3DUNet_DepthAdapter_Published(
        num_classes=NUM_CLASSES, lr=1e-2, momentum=0.99, nesterov=False, weight_decay=0.0, ignore_index=255,
        class_weights=None, voxel_weight_key=None, ce_weight=1.0, dice_weight=0.0, use_bn=True, target_depth=16,
        include_bg_in_dice=False)


_add_variant("3DUNet", make_cicek_depth_adapter_sgd_wce, MultiDicomDataModule3D, CHECKPOINT_DIR / "3DUNet")   # config.py:306-311

B200_VARIANT_NAMES = [v[0] for v in VARIANTS]     # the variants that run on the B200 kernels


def _append_reference_variants():
    """The reference registers four more model families (UNETR, SwinUNETR, R2U-Net, ResUNet++: config.py:316-391). They are
    outside the hot path; when a reference checkout is on sys.path their registry entries are kept, unchanged, after
    the B200 ones (their builders resolve `innovative3D.models.<Class>` through this tree's models.py to the reference's
    own PyTorch code), so `train.py:1615-1618` still loops over every variant."""
    from ._fallthrough import reference_module
    try:
        ref = reference_module("config")
    except Exception as e:   # a broken / partial checkout must not take the hot path down with it
        import warnings
        warnings.warn(f"innovative3D.config: reference config.py found but not importable ({e!r}); "
                      "only the B200 variants are registered")
        return
    if ref is not None:
        for v in getattr(ref, "VARIANTS", []):
            if v[0] not in B200_VARIANT_NAMES:
                VARIANTS.append(v)


_append_reference_variants()
VARIANT_NAMES = [v[0] for v in VARIANTS]
SELECTED_VARIANT = os.getenv("INNOVATIVE3D_VARIANT")


from ._fallthrough import module_getattr as _module_getattr  # noqa: E402

_ref_getattr = _module_getattr(__name__, "config", "dataset tables and data-layer switches live in the reference's config.py")


def __getattr__(name):
    try:
        return _ref_getattr(name)
    except AttributeError:
        if name in ("dataset_configs", "trainval_sets", "test_set"):   # no data layer on the path: nothing to train on
            return []
        raise
