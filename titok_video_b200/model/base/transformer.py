"""Parameter containers of the transformer stack (reference model/base/transformer.py).

The modules keep the reference's attribute names and parameter shapes so state dicts interchange
(`attn_layer.{i}.{pre_ln,to_qkv,out_proj}`, `ffd_layer.{i}.{norm,w12,w3}`, `attn_post_ln`, `ffd_post_ln`);
the arithmetic of Attn.forward / GEGLU.forward / ResidualAttentionBlock.forward runs in engine._layers.
"""
from __future__ import annotations

import torch.nn as nn

from .utils import RMSNorm, geglu_inner_dim


class GEGLU(nn.Module):
    def __init__(self, dim: int, mult: float = 4, mult_of: int = 32, dropout: float = 0.0):
        super().__init__()
        inner_dim = geglu_inner_dim(dim, mult, mult_of)
        self.inner_dim = inner_dim
        self.norm = RMSNorm(dim)
        self.w12 = nn.Linear(dim, inner_dim * 2, bias=False)  # rows [0,inner) = value, [inner,2*inner) = gate
        self.drop1 = nn.Dropout(dropout)
        self.w3 = nn.Linear(inner_dim, dim, bias=False)


class Attn(nn.Module):
    def __init__(self, dim: int, heads):
        super().__init__()
        self.dim = dim
        self.q_heads, self.kv_heads = heads
        self.head_dim = dim // self.q_heads
        self.gqa_dim = self.head_dim * self.kv_heads
        self.pre_ln = RMSNorm(dim)
        # output columns: [q (dim) | gate (dim) | k (gqa) | v (gqa)]   (transformer.py:78,87)
        self.to_qkv = nn.Linear(dim, (self.gqa_dim * 2) + (dim * 2), bias=False)
        self.out_proj = nn.Linear(dim, dim, bias=False)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, embed_dim: int = 512, heads=(8, 2), mlp_ratio: float = 4, num_layer: int = 2):
        super().__init__()
        self.num_layer = num_layer
        self.alpha = num_layer * 2  # KEEL residual scale (transformer.py:117)
        self.attn_layer = nn.ModuleList([Attn(embed_dim, heads) for _ in range(num_layer)])
        self.ffd_layer = nn.ModuleList([GEGLU(embed_dim, mult=mlp_ratio) for _ in range(num_layer)])
        self.attn_post_ln = nn.ModuleList([RMSNorm(embed_dim) for _ in range(num_layer - 1)])
        self.ffd_post_ln = nn.ModuleList([RMSNorm(embed_dim) for _ in range(num_layer - 1)])
