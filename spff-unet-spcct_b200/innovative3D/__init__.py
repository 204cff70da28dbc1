"""Drop-in `innovative3D` package for the SPFF-UNet hot path on B200.

Same module / class / registry surface as the reference's `innovative3D` (config.VARIANTS,
models.LitSPCT_*, helpers.ce_plus_macro_dice_loss / per_class_metrics_3d, unified_loss), with the
compute behind `model(x)`, the loss and the step metrics running in libspff_b200.so (sm_100a CUDA).

Deployment: put `spff-unet-spcct_b200/` on `PYTHONPATH` and run the reference's `train.py` / `test.py`
from the reference checkout. This is a regular package, so it wins over the reference's
`innovative3D/` directory (which has no `__init__.py` — its `_init_.py` is mis-named — and is
therefore a namespace portion) wherever the two sit on `sys.path`. The reference's directory is then
appended to this package's `__path__`: modules this tree does NOT provide — the CPU data pipeline
`innovative3D.datasets`, `stats_and_plots`, `ablation_tools`, … — resolve to the reference's files,
and `config` / `helpers` / `models` of this tree hand every name outside the hot path (dataset tables,
DICOM ingest, the other model families) through to the reference's module of the same name
(`_fallthrough.py`). Without a reference checkout on the path the package is self-contained for the
hot path; names of the data layer then raise a clear ImportError / AttributeError.
"""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
