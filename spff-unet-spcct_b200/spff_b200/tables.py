"""Parameter-only tables of the SPFF gates (tiny, differentiable torch ops on the parameters' device).

EnergyFiLM's (gamma, beta) come from an MLP over a constant sinusoidal code of the bin index
(reference innovative3D/models.py:1494-1512) and FourierGate's spectral mask acts as a circular
convolution with kernel irfft(freq_mask * mag_scale) (models.py:1537-1542): both depend on parameters
only, never on the activations. The CUDA gate kernels take the resulting tables as inputs and
return gradients w.r.t. them; autograd through these few ops carries them to the parameters.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def sinusoidal_pe(frames: int, d: int, device) -> torch.Tensor:
    """[1, d, F] positional code (EnergyFiLM3D._sinusoidal_pe, models.py:1494-1503)."""
    pos = torch.arange(frames, dtype=torch.float32, device=device)[None, None, :]
    i = torch.arange(max(1, d // 2), dtype=torch.float32, device=device)[None, :, None]
    denom = torch.exp(i * (-math.log(10000.0) / max(1, d // 2)))
    pe = torch.cat([torch.sin(pos * denom), torch.cos(pos * denom)], dim=1)
    if pe.shape[1] < d:
        pe = torch.cat([pe, torch.zeros(1, 1, pe.shape[-1], device=device)], dim=1)
    return pe


def efilm_tables(w0, b0, w2, b2, channels: int, frames: int):
    """(g1, bt), each [C, F] fp32 contiguous: g1 = 1 + tanh(gamma), bt = beta (models.py:1505-1512)."""
    pe = sinusoidal_pe(frames, w0.shape[1], w0.device)
    h = F.relu(F.conv1d(pe, w0, b0))
    gb = F.conv1d(h, w2, b2)[0]
    return (1.0 + torch.tanh(gb[:channels])).contiguous(), gb[channels:].contiguous()


def fourier_kernel(freq_mask, mag_scale, frames: int):
    """kfg [F]: irfft(rfft(s) * M, n=F) == circular_conv(s, kfg) with kfg = irfft(M, n=F)."""
    m = (freq_mask * mag_scale).reshape(-1)
    return torch.fft.irfft(m.to(torch.complex64), n=frames).contiguous()
