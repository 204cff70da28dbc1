"""spff_b200 — host-side binding of the B200 SPFF-UNet kernels (libspff_b200.so).

`_lib` is the ctypes binding of include/spff_b200.h, `ops` the tensor-level wrappers the model code
(innovative3D.models in this tree) calls. Importing this package loads the shared object and fails
if it has not been built; there is no CPU or PyTorch fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
