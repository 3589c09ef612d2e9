"""TiTokEncoder / TiTokDecoder with the constructor, forward signature and state-dict layout of the reference's
model/base/blocks.py, executing on the sm_100a kernels (titok_video_b200.engine).

Batches are *packed*: `videos` is a list of B clips [C, T_i, H_i, W_i] (bf16, T%p0 = H%p1 = W%p2 = 0) and
`token_counts[i]` latent tokens are produced per clip. Per clip the packed rows are latent rows first, then the
patch rows (blocks.py:85-86). All packing metadata is derived on the host (plan.py) -- pass `token_counts`
(and `grids`) as lists or CPU tensors to keep the forward free of device synchronisation.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from ... import engine
from .transformer import ResidualAttentionBlock
from .utils import RMSNorm, geglu_inner_dim, get_model_dims

def _float_dtype(t: torch.Tensor) -> torch.dtype:
    return t.dtype if t.is_floating_point() else torch.bfloat16


def _wants_grad(module: nn.Module, *inputs) -> bool:
    """True when autograd has to see this forward: grad mode on and a parameter (or a tensor input) requires grad.
    The training path (titok_video_b200/backward.py) then records the activations the backward kernels need."""
    if not torch.is_grad_enabled():
        return False
    if any(isinstance(t, torch.Tensor) and t.requires_grad for t in inputs):
        return True
    from ... import backward

    return any(p.requires_grad for p in backward.stack_params(module))


class _Stack(nn.Module):
    """Shared construction of the two stacks (blocks.py:31-69,108-146)."""

    def _build(self, model_size, patch_size, in_features, out_features):
        self.patch_size = torch.tensor(list(patch_size), dtype=torch.int32)  # plain attribute, as in the reference
        self.patch_size_tuple = tuple(int(p) for p in patch_size)
        self.width, self.num_layers, self.heads, mlp_ratio = get_model_dims(model_size)
        self.inner_dim = geglu_inner_dim(self.width, mlp_ratio)
        scale = self.width ** -0.5
        self.proj_in = nn.Linear(in_features, self.width, bias=True)
        self.mask_token = nn.Parameter(scale * torch.randn(1, 1))
        self.ln_pre_t = RMSNorm(self.width)
        self.ln_pre_p = RMSNorm(self.width)
        self.model_layers = ResidualAttentionBlock(embed_dim=self.width, heads=self.heads, mlp_ratio=mlp_ratio,
                                                   num_layer=self.num_layers)
        self.ln_post = RMSNorm(self.width)
        self.proj_out = nn.Linear(self.width, out_features, bias=True)

    def _plan(self, grids_px, token_counts, device) -> engine.DevicePlan:
        return engine.get_device_plan(grids_px, token_counts, self.patch_size_tuple, self.patch_channels, device)


class TiTokEncoder(_Stack):
    def __init__(self, model_size="tiny", patch_size=(4, 8, 8), in_channels=3, out_channels=5):
        super().__init__()
        self.token_size = out_channels
        self.in_channels = in_channels
        self.patch_channels = in_channels
        if not 1 <= out_channels <= 8:
            raise ValueError("token size must be in 1..8")
        self._build(model_size, patch_size, in_channels * math.prod(patch_size), out_channels)

    @torch.compiler.disable()
    def forward_impl(self, videos: Sequence[torch.Tensor], token_counts, grids=None, fsq=None):
        """Returns (z, codes, indices, device_plan). z: [sum(token_counts), token_size] pre-quantisation tokens."""
        device = videos[0].device
        engine.require_cuda(device)
        tcs = engine.to_host_ints(token_counts)
        if len(tcs) != len(videos):
            raise ValueError("len(token_counts) must equal the number of clips")
        if grids is None:
            gpx = [tuple(v.shape[1:]) for v in videos]
        else:
            gpx = [tuple(g) for g in engine.to_host_ints(grids)]
            for v, g in zip(videos, gpx):
                if tuple(v.shape[1:]) != tuple(g):
                    raise ValueError(f"grid {g} does not match clip shape {tuple(v.shape)}")
        for v in videos:
            if v.shape[0] != self.in_channels:
                raise ValueError(f"expected {self.in_channels} channels, got {tuple(v.shape)}")
        dp = self._plan(gpx, tcs, device)
        consts = (fsq if fsq is not None else _dummy_fsq(self.token_size))._consts(device)
        if _wants_grad(self, *videos):
            # training path: recorded by autograd through the registered operator titok_b200::encoder_stack (ops.py);
            # pixels get a gradient if they ask for one (the discriminator's gradient penalties differentiate w.r.t.
            # the input, loss_module.py:149-152)
            from ... import ops

            if any(v.requires_grad for v in videos):
                flat = torch.cat([v.reshape(-1).to(torch.bfloat16) for v in videos])
            else:
                with torch.no_grad():
                    flat = engine.flatten_clips(videos, dp)
            z, codes, idx = ops.encoder_call(self, dp, consts, flat, True)
            return z, codes, idx, dp
        with torch.no_grad():
            flat = engine.flatten_clips(videos, dp, keep_u8=True)
            z, codes, idx = engine.encoder_launch(self, dp, flat, consts)
        return z, codes, idx, dp

    @torch.compiler.disable()
    def forward(self, videos, token_counts, grids=None):
        z, _, _, _ = self.forward_impl(videos, token_counts, grids)
        if z.requires_grad:
            return z.to(_float_dtype(videos[0]))
        return z.clone().to(_float_dtype(videos[0]))


class TiTokDecoder(_Stack):
    def __init__(self, model_size="tiny", patch_size=(4, 8, 8), in_channels=5, out_channels=3):
        super().__init__()
        self.token_size = in_channels
        self.out_channels = out_channels
        self.patch_channels = out_channels
        if not 1 <= in_channels <= 8:
            raise ValueError("token size must be in 1..8")
        self._build(model_size, patch_size, in_channels, out_channels * math.prod(patch_size))

    @torch.compiler.disable()
    def forward_impl(self, tokens: torch.Tensor, token_counts, grids) -> (torch.Tensor, engine.DevicePlan):
        device = tokens.device
        engine.require_cuda(device)
        tcs = engine.to_host_ints(token_counts)
        gpx = [tuple(g) for g in engine.to_host_ints(grids)]
        if tokens.shape[0] != sum(tcs) or tokens.shape[-1] != self.token_size:
            raise ValueError(f"tokens {tuple(tokens.shape)} do not match token_counts (sum {sum(tcs)})")
        dp = self._plan(gpx, tcs, device)
        if _wants_grad(self, tokens):
            from ... import ops

            return ops.decoder_call(self, dp, tokens, True), dp
        with torch.no_grad():
            codes = tokens.detach().to(torch.bfloat16).contiguous()
            out = dp.buf("clips_out", (dp.plan.total_numel,))
            engine.decoder_launch(self, dp, codes, out)
        return out, dp

    @torch.compiler.disable()
    def forward(self, tokens, token_counts, grids) -> List[torch.Tensor]:
        out, dp = self.forward_impl(tokens, token_counts, grids)
        if out.requires_grad:
            from ... import backward

            return backward.split_clips_autograd(out.to(tokens.dtype if tokens.is_floating_point() else torch.bfloat16), dp.plan)
        return engine.split_clips(out.clone().to(tokens.dtype if tokens.is_floating_point() else torch.bfloat16), dp.plan)


_DUMMY = {}


def _dummy_fsq(token_size: int):
    """FSQ constants for an encoder used without a quantizer (e.g. as the discriminator, loss_module.py:43-48):
    the fused head still needs well-formed constants; its codes / indices outputs are simply ignored."""
    if token_size not in _DUMMY:
        from ..quantizer.fsq import FSQ

        _DUMMY[token_size] = FSQ([3] * token_size)
    return _DUMMY[token_size]
