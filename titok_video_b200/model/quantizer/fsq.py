"""Finite scalar quantizer with the interface of the reference's model/quantizer/fsq.py (class FSQ).

`forward` and `indices_to_codes` run the CUDA kernels of csrc/fsq.cu (one fused launch instead of ~15
elementwise kernels, fsq.py:123-135). The small tensor helpers (`bound`, `quantize`, `codes_to_indices`, ...)
keep the reference's names and semantics for callers that poke at them; they are host-side conveniences used
to build the `implicit_codebook` buffer at construction and are not on the hot path.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

_DT = {torch.bfloat16: 0, torch.float32: 1}


class _FsqFn(torch.autograd.Function):
    """codes, indices = FSQ(z) with the straight-through gradient of round_ste (fsq.py:48-51)."""

    @staticmethod
    def forward(ctx, z2: torch.Tensor, fsq: "FSQ"):
        from ... import _lib, engine

        n, d = z2.shape
        codes = torch.empty_like(z2)
        idx = torch.empty((n,), dtype=torch.int32, device=z2.device)
        c = fsq._consts(z2.device)
        _lib.call("ttk_fsq_fwd", engine._ptr(z2), engine._ptr(codes), engine._ptr(idx), n, _DT[z2.dtype], d, *c,
                  engine._stream())
        ctx.save_for_backward(z2)
        ctx.fsq = fsq
        ctx.mark_non_differentiable(idx)
        return codes, idx

    @staticmethod
    def backward(ctx, dcodes, _didx):
        from ... import _lib, engine

        (z2,) = ctx.saved_tensors
        dcodes = dcodes.to(z2.dtype).contiguous()
        dz = torch.empty_like(z2)
        c = ctx.fsq._consts(z2.device)
        _lib.call("ttk_fsq_bwd", engine._ptr(z2), engine._ptr(dcodes), engine._ptr(dz), z2.shape[0], _DT[z2.dtype],
                  z2.shape[1], *c, engine._stream())
        return dz, None


class FSQ(nn.Module):
    def __init__(self, levels: List[int], dim: Optional[int] = None):
        super().__init__()
        levels = [int(v) for v in levels]
        if not 1 <= len(levels) <= 8:
            raise ValueError("FSQ supports 1..8 levels")
        lv = torch.tensor(levels, dtype=torch.int32)
        self.register_buffer("_levels", lv, persistent=False)
        basis = torch.cumprod(torch.tensor([1] + levels[:-1]), dim=0, dtype=torch.int32)
        self.register_buffer("_basis", basis, persistent=False)
        self.codebook_dim = len(levels)
        self.dim = dim if dim is not None else len(levels)
        self.codebook_size = int(self._levels.prod().item())
        self.register_buffer("implicit_codebook", self._indices_to_codes(torch.arange(self.codebook_size)),
                             persistent=False)
        self._const_cache = {}

    # ---- constants handed to the kernels: evaluated with torch, op for op as fsq.py:78-90 does ----
    def _consts(self, device):
        key = str(device)
        c = self._const_cache.get(key)
        if c is None:
            from ... import _lib

            lv = self._levels.to(device)
            eps = 1e-3
            half_l = (lv - 1) * (1 + eps) / 2
            offset = torch.where(lv % 2 == 0, 0.5, 0.0)
            shift = (offset / half_l).atanh()
            half_width = (lv // 2).to(torch.float32)
            c = (_lib.float_array(half_l.cpu().tolist()), _lib.float_array(offset.cpu().tolist()),
                 _lib.float_array(shift.cpu().tolist()), _lib.float_array(half_width.cpu().tolist()),
                 _lib.int_array(self._basis.cpu().tolist()), _lib.int_array(self._levels.cpu().tolist()))
            self._const_cache[key] = c
        return c

    # ---- reference-named helpers (fsq.py:78-121) ----
    def bound(self, z: torch.Tensor, eps: float = 1e-3) -> torch.Tensor:
        half_l = (self._levels - 1) * (1 + eps) / 2
        offset = torch.where(self._levels % 2 == 0, 0.5, 0.0)
        shift = (offset / half_l).atanh()
        return (z + shift).tanh() * half_l - offset

    def quantize(self, z: torch.Tensor) -> torch.Tensor:
        b = self.bound(z)
        q = b + (b.round() - b).detach()
        return q / (self._levels // 2)

    def _scale_and_shift(self, zhat_normalized):
        hw = self._levels // 2
        return zhat_normalized * hw + hw

    def _scale_and_shift_inverse(self, zhat):
        hw = self._levels // 2
        return (zhat - hw) / hw

    def indices_to_level_indices(self, indices: torch.Tensor) -> torch.Tensor:
        return (indices.unsqueeze(-1) // self._basis) % self._levels

    def _indices_to_codes(self, indices: torch.Tensor) -> torch.Tensor:
        return self._scale_and_shift_inverse(self.indices_to_level_indices(indices))

    def codes_to_indices(self, zhat: torch.Tensor) -> torch.Tensor:
        return (self._scale_and_shift(zhat) * self._basis).sum(dim=-1).to(torch.int32)

    def indices_to_codes(self, indices: torch.Tensor, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """Inverse of codes_to_indices (fsq.py:111-121) on the fused CUDA kernel. CPU tensors are rejected: the only
        host-side evaluation is the one-off construction of the `implicit_codebook` buffer in __init__."""
        assert indices is not None
        from ... import _lib, engine

        engine.require_cuda(indices.device)

        if indices.dtype not in (torch.int32, torch.int64):
            indices = indices.to(torch.int64)
        flat = indices.reshape(-1).contiguous()
        kdt = out_dtype if out_dtype in _DT else torch.float32
        codes = torch.empty((flat.numel(), self.codebook_dim), dtype=kdt, device=flat.device)
        c = self._consts(flat.device)
        _lib.call("ttk_fsq_indices_to_codes", engine._ptr(flat), 0 if flat.dtype == torch.int32 else 1,
                  engine._ptr(codes), _DT[kdt], flat.numel(), self.codebook_dim, c[3], c[4], c[5], engine._stream())
        return codes.view(*indices.shape, self.codebook_dim).to(out_dtype)

    @torch.compiler.disable()
    def forward(self, z: torch.Tensor):
        """(codes in z's dtype, {'indices': int32}) -- the reference's return contract (fsq.py:123-135)."""
        from ... import engine

        engine.require_cuda(z.device)
        if z.shape[-1] != self.codebook_dim:
            raise ValueError(f"expected last dim {self.codebook_dim}, got {tuple(z.shape)}")
        orig_dtype = z.dtype
        zk = z if z.dtype in _DT else z.float()  # the reference computes in fp32 and casts codes back
        z2 = zk.reshape(-1, self.codebook_dim).contiguous()
        if torch.is_grad_enabled() and z2.requires_grad:
            codes, idx = _FsqFn.apply(z2, self)
        else:
            codes, idx = _FsqFn.forward(_NoCtx(), z2.detach(), self)
        codes = codes.view(z.shape).to(orig_dtype)
        return codes, {"indices": idx.view(z.shape[:-1])}


class _NoCtx:
    def save_for_backward(self, *a):
        pass

    def mark_non_differentiable(self, *a):
        pass
