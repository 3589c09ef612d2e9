"""Model size table, weight init and the patch feature layout (reference model/base/utils.py)."""
from __future__ import annotations

import torch
import torch.nn as nn

HEAD_DIM = 64

# model_size -> (layers, [q heads, kv heads])   (reference utils.py:8-23)
_SIZES = {
    "tiny": (4, [4, 2]),
    "small": (8, [8, 2]),
    "base": (12, [12, 4]),
    "large": (24, [16, 4]),
}


def get_model_dims(model_size: str = "tiny", head_dim: int = HEAD_DIM, mlp_ratio: float = 4.0):
    """(width, layers, [q_heads, kv_heads], mlp_ratio) -- same table and return order as utils.py:8-23."""
    layers, heads = _SIZES[model_size]
    return int(head_dim * heads[0]), layers, list(heads), mlp_ratio


def geglu_inner_dim(dim: int, mult: float = 4.0, mult_of: int = 32) -> int:
    """GEGLU hidden size: int(mult * 2/3 * dim) rounded up to a multiple of 32 (transformer.py:38-40)."""
    inner = int(mult * (2 / 3) * dim)
    return mult_of * ((inner + mult_of - 1) // mult_of)


class RMSNorm(nn.Module):
    """Parameter container with the state-dict layout of flash_attn.ops.triton.layer_norm.RMSNorm
    (weight [dim], no bias, eps 1e-5). The arithmetic lives in the CUDA kernels (csrc/rowops.cu, csrc/gemm.cu)."""

    def __init__(self, hidden_size: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.register_parameter("bias", None)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from ... import engine, _lib

        engine.require_cuda(x.device)
        x2 = x.reshape(-1, x.shape[-1]).to(torch.bfloat16).contiguous()
        y = torch.empty_like(x2)
        w = self.weight.detach().float().contiguous()
        _lib.call("ttk_rmsnorm_fwd", engine._ptr(x2), x2.stride(0), engine._ptr(w), engine._ptr(y), y.stride(0),
                  x2.shape[0], x2.shape[1], engine._stream())
        return y.view(x.shape).to(x.dtype)


def init_weights(module: nn.Module) -> None:
    """Same initial distribution as utils.py:54-66: Linear ~ trunc_normal(std 0.02), zero bias; norm weights 1."""
    if isinstance(module, nn.Linear):
        nn.init.trunc_normal_(module.weight.data, mean=0.0, std=0.02)
        if module.bias is not None:
            nn.init.zeros_(module.bias)
    elif isinstance(module, (nn.LayerNorm, RMSNorm)):
        if getattr(module, "bias", None) is not None:
            nn.init.zeros_(module.bias)
        if getattr(module, "weight", None) is not None:
            nn.init.ones_(module.weight)
    elif isinstance(module, (nn.Conv3d, nn.Conv2d)):
        nn.init.xavier_uniform_(module.weight)
        nn.init.zeros_(module.bias)
