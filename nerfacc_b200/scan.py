"""Segmented inclusive / exclusive sum and product.

Mirrors /root/reference/nerfacc/scan.py (public functions :14-282, autograd
Functions :285-468): three addressing modes -- batched (torch.cumsum/cumprod on
the last dim, works on CPU), ``packed_info`` (one warp per chunk, nfa_scan_packed)
and ``indices`` (runs of equal keys, nfa_scan_by_key; the reference uses CUB there).
"""
from typing import Optional

import torch
from torch import Tensor

from . import _lib


def _native_scan(inputs: Tensor, packed_info: Optional[Tensor], indices: Optional[Tensor], prod: bool,
                 inclusive: bool, reverse: bool, normalize: bool = False) -> Tensor:
    _lib.require_cuda(inputs, "scan")
    inputs = inputs.contiguous()
    if inputs.dtype != torch.float32:
        raise RuntimeError("nerfacc_b200 scans support float32 inputs only.")
    out = torch.empty_like(inputs)
    if inputs.numel() == 0:
        return out
    device = inputs.device
    if packed_info is not None:
        pi = packed_info.contiguous()
        if pi.dtype != torch.int64:
            pi = pi.to(torch.int64)
        _lib.call("nfa_scan_packed", device, pi.shape[0], _lib.ptr(pi), _lib.ptr(inputs), _lib.ptr(out),
                  int(prod), int(inclusive), int(reverse), int(normalize))
    else:
        keys = indices.contiguous()
        if keys.dtype != torch.int64:
            keys = keys.to(torch.int64)
        lib = _lib.load()
        ws = torch.empty(lib.nfa_scan_by_key_workspace_bytes(inputs.numel()), dtype=torch.uint8, device=device)
        _lib.call("nfa_scan_by_key", device, inputs.numel(), _lib.ptr(keys), _lib.ptr(inputs), _lib.ptr(out),
                  int(prod), int(inclusive), int(reverse), _lib.ptr(ws))
    return out


class _SegScan(torch.autograd.Function):
    """One Function for all eight native scan variants (reference scan.py:285-468)."""

    @staticmethod
    def forward(ctx, inputs, packed_info, indices, prod: bool, inclusive: bool, normalize: bool):
        outputs = _native_scan(inputs, packed_info, indices, prod, inclusive, False, normalize)
        if ctx.needs_input_grad[0]:
            ctx.prod, ctx.inclusive, ctx.normalize = prod, inclusive, normalize
            ctx.packed_info, ctx.indices = packed_info, indices
            if prod:
                ctx.save_for_backward(inputs, outputs)
        return outputs

    @staticmethod
    def backward(ctx, grad_outputs):
        assert not ctx.normalize, "Only support backward for normalize==False."
        grad_outputs = grad_outputs.contiguous()
        if not ctx.prod:
            # reference scan.py:307-309,335-337,403,423: same scan over the reversed chunk
            g = _native_scan(grad_outputs, ctx.packed_info, ctx.indices, False, ctx.inclusive, True)
        else:
            # reference scan.cu:199-210,289-300 / scan_cub.cu:205-211,274-280
            inputs, outputs = ctx.saved_tensors
            g = _native_scan(grad_outputs * outputs, ctx.packed_info, ctx.indices, False, ctx.inclusive, True)
            g = g / inputs.clamp_min(1e-10)
        return g, None, None, None, None, None


def _check_modes(inputs: Tensor, packed_info: Optional[Tensor], indices: Optional[Tensor]) -> None:
    if indices is not None and packed_info is not None:
        raise ValueError("Only one of `indices` and `packed_info` can be specified.")
    if indices is not None:
        assert indices.dim() == 1 and indices.shape == inputs.shape, \
            "indices must be 1-D with the same shape as inputs."
    if packed_info is not None:
        assert inputs.dim() == 1, "inputs must be flattened."
        assert packed_info.dim() == 2 and packed_info.shape[-1] == 2, "packed_info must be 2-D with shape (B, 2)."


def inclusive_sum(inputs: Tensor, packed_info: Optional[Tensor] = None, indices: Optional[Tensor] = None) -> Tensor:
    """Inclusive sum along the last dim, or per chunk of a flattened tensor (reference scan.py:14-77)."""
    _check_modes(inputs, packed_info, indices)
    if indices is None and packed_info is None:
        return torch.cumsum(inputs, dim=-1)
    return _SegScan.apply(inputs, packed_info, indices, False, True, False)


def exclusive_sum(inputs: Tensor, packed_info: Optional[Tensor] = None, indices: Optional[Tensor] = None) -> Tensor:
    """Exclusive sum (reference scan.py:80-145)."""
    _check_modes(inputs, packed_info, indices)
    if indices is None and packed_info is None:
        shifted = torch.cat([torch.zeros_like(inputs[..., :1]), inputs[..., :-1]], dim=-1)
        return torch.cumsum(shifted, dim=-1)
    return _SegScan.apply(inputs, packed_info, indices, False, False, False)


def inclusive_prod(inputs: Tensor, packed_info: Optional[Tensor] = None, indices: Optional[Tensor] = None) -> Tensor:
    """Inclusive product (reference scan.py:148-211)."""
    _check_modes(inputs, packed_info, indices)
    if indices is None and packed_info is None:
        return torch.cumprod(inputs, dim=-1)
    return _SegScan.apply(inputs, packed_info, indices, True, True, False)


def exclusive_prod(inputs: Tensor, packed_info: Optional[Tensor] = None, indices: Optional[Tensor] = None) -> Tensor:
    """Exclusive product (reference scan.py:214-282)."""
    _check_modes(inputs, packed_info, indices)
    if indices is None and packed_info is None:
        shifted = torch.cat([torch.ones_like(inputs[..., :1]), inputs[..., :-1]], dim=-1)
        return torch.cumprod(shifted, dim=-1)
    return _SegScan.apply(inputs, packed_info, indices, True, False, False)
