"""Tensor-level wrappers over the C ABI (include/spff_b200.h).

Activations are bf16 "position-major" views [N, D, H, W, C] whose last dimension is contiguous and
whose position pitch `ld` (= stride of W, in elements) may exceed C — a channel slice of a wider
buffer is passed as the sliced view itself. These functions only validate, extract pointers and
enqueue on the current CUDA stream.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import Shape, call, ptr, stream_ptr


def _view(t: torch.Tensor, c: int):
    """(shape, ld) of a position-major bf16 view holding >= c channels."""
    if t.dtype != torch.bfloat16 or t.dim() != 5 or not t.is_cuda:
        raise ValueError(f"expected a CUDA bf16 [N,D,H,W,C] view, got {t.dtype} {tuple(t.shape)} on {t.device}")
    n, d, h, w, cc = t.shape
    if cc < c:
        raise ValueError(f"view has {cc} channels, need {c}")
    ld = t.stride(3)
    if t.stride(4) != 1 or t.stride(2) != w * ld or t.stride(1) != h * w * ld or t.stride(0) != d * h * w * ld:
        raise ValueError(f"view is not position-major with a uniform pitch: strides {t.stride()}")
    return Shape(n, d, h, w), ld


def conv3_kc(k_channels: int) -> int:
    return 64 if k_channels % 64 == 0 else 32


def pack_conv3_weight(w: torch.Tensor, fwd: bool = True, dgrad: bool = True):
    """nn.Conv3d weight [Cout,Cin,3,3,3] fp32 -> (w_fwd, w_dgrad) bf16 GEMM operands."""
    cout, cin = w.shape[0], w.shape[1]
    w = w.detach().contiguous().float()
    wf = torch.empty(27 * cin * cout, dtype=torch.bfloat16, device=w.device) if fwd else None
    wd = torch.empty(27 * cin * cout, dtype=torch.bfloat16, device=w.device) if dgrad else None
    call("spff_pack_conv3_weight", ptr(w), ptr(wf), ptr(wd), cout, cin, stream_ptr())
    return wf, wd


def conv3d_k3_fwd(x, cin, w_fwd, y, cout):
    s, ldx = _view(x, cin)
    s2, ldy = _view(y, cout)
    assert (s.n, s.d, s.h, s.w) == (s2.n, s2.d, s2.h, s2.w)
    call("spff_conv3d_k3_fwd", ptr(x), ldx, cin, ptr(w_fwd), ptr(y), ldy, cout, s, stream_ptr())


def conv3d_k3_dgrad(dy, cout, w_dgrad, dx, cin):
    s, lddy = _view(dy, cout)
    s2, lddx = _view(dx, cin)
    assert (s.n, s.d, s.h, s.w) == (s2.n, s2.d, s2.h, s2.w)
    call("spff_conv3d_k3_dgrad", ptr(dy), lddy, cout, ptr(w_dgrad), ptr(dx), lddx, cin, s, stream_ptr())


_WS = {}


def workspace(nbytes: int, device) -> torch.Tensor:
    """Grow-only scratch buffer per device (the library itself allocates nothing)."""
    key = str(device)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


def conv3d_k3_wgrad_workspace(cin, cout, x) -> torch.Tensor:
    s, _ = _view(x, cin)
    return workspace(int(_lib.lib.spff_conv3d_k3_wgrad_workspace(cin, cout, s)), x.device)


def conv3d_k3_wgrad(x, cin, dy, cout, dw, beta: float = 0.0, ws: torch.Tensor | None = None):
    """dw [Cout,Cin,3,3,3] fp32 = beta*dw + grad_weight."""
    s, ldx = _view(x, cin)
    s2, lddy = _view(dy, cout)
    assert (s.n, s.d, s.h, s.w) == (s2.n, s2.d, s2.h, s2.w)
    assert dw.dtype == torch.float32 and dw.is_contiguous() and tuple(dw.shape) == (cout, cin, 3, 3, 3)
    if ws is None:
        ws = conv3d_k3_wgrad_workspace(cin, cout, x)
    call("spff_conv3d_k3_wgrad", ptr(x), ldx, cin, ptr(dy), lddy, cout, s, ptr(dw), float(beta), ptr(ws),
         ws.numel(), stream_ptr())
