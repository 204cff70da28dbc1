"""Drop-in `innovative3D` package for the SPFF-UNet hot path on B200.

Same module / class / registry surface as the reference's `innovative3D` (config.VARIANTS,
models.LitSPCT_*, helpers.ce_plus_macro_dice_loss / per_class_metrics_3d, unified_loss), with the
compute behind `model(x)`, the loss and the step metrics running in libspff_b200.so (sm_100a CUDA).
Put `spff-unet-spcct_b200/` ahead of the reference on `sys.path` and `train.py` / `test.py` pick it up.
"""
