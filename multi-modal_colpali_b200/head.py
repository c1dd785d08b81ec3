"""Fused retrieval head for ingestion (K3): ``mask * normalize(Linear(hidden))`` in one kernel.

Replaces the tail of ``model(**batch)`` on the ingestion path (functions.py:795, 839; query side
functions.py:888, 05_experiment02.py:211); body restated from HF ``modeling_colpali.py:148-155``.
The encoder itself stays the reference PyTorch model.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _native as N
from .scoring import _DTYPES, _stream


_ROUND = {"f32": N.ROUND_F32, "reference": N.ROUND_REFERENCE}


def mask_for_kernel(attention_mask: torch.Tensor, dev: torch.device) -> torch.Tensor:
    """Flat attention mask in a form K3 reads as it is (1-, 4- or 8-byte integers)."""
    m = attention_mask.reshape(-1).to(dev).contiguous()
    if m.dtype == torch.bool:
        m = m.view(torch.uint8)
    if m.element_size() not in (1, 4, 8) or m.is_floating_point():
        m = (m != 0).view(torch.uint8)
    return m


def project_normalize(hidden: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
                      attention_mask: Optional[torch.Tensor] = None, *, round_mode: str = "reference") -> torch.Tensor:
    """``hidden [..., n_tok, H]`` (CUDA, bf16/fp16), ``weight [128, H]`` (``nn.Linear.weight``),
    ``bias [128]`` or None, ``attention_mask [..., n_tok]`` (any integer/bool dtype) or None
    -> ``[..., n_tok, 128]`` unit-norm rows, zeroed where the mask is 0, in the input dtype.

    ``round_mode="reference"`` (default) rounds where the reference's 16-bit model rounds (Linear output, norm,
    quotient: HF modeling_colpali.py:149-152), so the stored embedding is the one the reference would store;
    ``"f32"`` normalises the fp32 accumulators and rounds once (closer to the exact unit vector)."""
    lib = N.load()
    if round_mode not in _ROUND:
        raise ValueError(f"round_mode must be one of {sorted(_ROUND)}")
    if not hidden.is_cuda:
        raise RuntimeError("project_normalize runs on an sm_100 GPU only; there is no CPU fallback")
    dt = hidden.dtype
    if dt not in _DTYPES:
        raise NotImplementedError(f"hidden dtype {dt}: bfloat16 or float16")
    lead = hidden.shape[:-1]
    hdim = hidden.shape[-1]
    if weight.shape != (N.DIM, hdim):
        raise ValueError(f"weight must be [{N.DIM}, {hdim}], got {tuple(weight.shape)}")
    dev = hidden.device
    h2 = hidden.reshape(-1, hdim).contiguous()
    w = weight.to(device=dev, dtype=dt).contiguous()
    b = None if bias is None else bias.to(device=dev, dtype=dt).contiguous()
    m = None
    if attention_mask is not None:
        if attention_mask.shape != lead:
            raise ValueError("attention_mask must match hidden's leading dimensions")
        m = mask_for_kernel(attention_mask, dev)
    out = torch.empty((h2.shape[0], N.DIM), dtype=dt, device=dev)
    with torch.cuda.device(dev):
        N.check(lib.lis_project_normalize(h2.data_ptr(), h2.shape[0], hdim, w.data_ptr(),
                                          None if b is None else b.data_ptr(), None if m is None else m.data_ptr(),
                                          0 if m is None else m.element_size(), _DTYPES[dt], _ROUND[round_mode],
                                          None, out.data_ptr(), _stream(dev)))
    return out.reshape(*lead, N.DIM)
