"""`torch.library` registration of the two stack operators (SURVEY 8b "Registration"; train.py:38-39 may wrap the model
in torch.compile, Lightning runs it under autocast).

    titok_b200::encoder_stack(flat, params, handle, train) -> (z, codes, indices)
    titok_b200::decoder_stack(codes, params, handle, train) -> flat reconstruction

Each is ONE dispatcher-visible operator for a whole stack (patchify -> proj_in -> L transformer layers -> head (+ FSQ), or
its mirror): the ~25 / ~110 kernel launches behind it go through the C ABI (`_lib`). Registered with
  * a fake (meta) implementation -- output shapes come from the packing plan, no kernel runs -- so FakeTensor tracing
    (torch.compile, torch.export) can see through a call,
  * an autograd formula (`register_autograd`): the training forward records its tape, the backward launches the CUDA
    backward kernels (backward.encoder_backward / decoder_backward) and returns fp32 parameter gradients in the
    reference's parameter layout, the clip / code gradients in bf16,
and therefore composes with autocast (the op takes the fp32 master parameters as graph edges and reads the prepared
bf16 copies itself) and with `torch.library.opcheck` (tests/test_gpu_ops.py).

`handle` names the host-side state a call needs but a tensor cannot carry: the module (prepared weights), the cached
device plan of this batch composition, the FSQ constants. Handles live in a bounded registry; plans are cached per
(shapes, token_counts), so a compiled graph that replays the same composition finds its handle again.

The module facade (`TiTok.forward`, ...) still starts with `@torch.compiler.disable()`: it turns `token_counts` /
shapes into host integers to build the packing plan, which is data-dependent Python that Dynamo cannot trace (the
reference's own forward breaks the graph at the same place, blocks.py:82-86).
"""
from __future__ import annotations

import collections
import itertools
import weakref
from typing import List, Sequence, Tuple

import torch

from . import backward, engine

bf16 = torch.bfloat16

_MAX_HANDLES = 512
_calls: "collections.OrderedDict[int, _Call]" = collections.OrderedDict()
_by_key = {}
_ids = itertools.count(1)


class _Call:
    """(module, device plan, FSQ constants) behind a handle. The module is held weakly: the registry must not keep models
    (and their prepared weights) alive."""
    __slots__ = ("_m", "dp", "consts", "tape", "key")

    def __init__(self, m, dp, consts, key):
        self._m, self.dp, self.consts, self.tape, self.key = weakref.ref(m), dp, consts, None, key

    @property
    def m(self):
        m = self._m()
        if m is None:
            raise RuntimeError("titok_b200: the module behind this stack handle has been freed")
        return m


def handle_for(module, dp, consts=None) -> int:
    """Stable integer for (module, device plan): the same batch composition maps to the same handle."""
    key = (id(module), id(dp))
    h = _by_key.get(key)
    if h is not None and h in _calls and _calls[h]._m() is module and _calls[h].dp is dp:
        _calls.move_to_end(h)
        if consts is not None:
            _calls[h].consts = consts
        return h
    h = next(_ids)
    _calls[h] = _Call(module, dp, consts, key)
    _by_key[key] = h
    while len(_calls) > _MAX_HANDLES:
        _, old = _calls.popitem(last=False)
        _by_key.pop(old.key, None)
    return h


def _call(handle: int) -> _Call:
    c = _calls.get(int(handle))
    if c is None:
        raise RuntimeError(f"titok_b200: stale stack handle {handle} (the plan it named was evicted; call the module again)")
    return c


# ------------------------------------------------------------------------------------------------
# encoder
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("titok_b200::encoder_stack", mutates_args=())
def encoder_stack(flat: torch.Tensor, params: Sequence[torch.Tensor], handle: int,
                  train: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """flat: all clips of the batch back to back (bf16). Returns z, codes [T, token_size] bf16 and indices int32 [T]
    (blocks.py:71-104 + fsq.py:123-135)."""
    c = _call(handle)
    engine.require_cuda(flat.device)
    if train:
        z, codes, idx, tape = backward.encoder_forward_train(c.m, c.dp, flat, c.consts)
        c.tape = tape
        return z, codes, idx
    z, codes, idx = engine.encoder_launch(c.m, c.dp, flat, c.consts)
    return z.clone(), codes.clone(), idx.clone()  # (the launch writes into the plan's workspace)


@encoder_stack.register_fake
def _(flat, params, handle, train):
    c = _call(handle)
    t, ts = c.dp.plan.T, c.m.token_size
    return (flat.new_empty((t, ts), dtype=bf16), flat.new_empty((t, ts), dtype=bf16),
            flat.new_empty((t,), dtype=torch.int32))


def _enc_setup(ctx, inputs, output):
    flat, params, handle, train = inputs
    c = _call(handle)
    if not train or c.tape is None:
        raise RuntimeError("titok_b200::encoder_stack was recorded by autograd without its training forward (train=False)")
    ctx.call, ctx.tape, c.tape = c, c.tape, None
    ctx.need_input = flat.requires_grad
    ctx.meta = backward._meta(c.m)
    ctx.mark_non_differentiable(output[1], output[2])


def _enc_backward(ctx, dz, _dcodes, _didx):
    c = ctx.call
    grads, dflat = backward.encoder_backward(c.m, c.dp, ctx.tape, dz.to(bf16).contiguous(), ctx.need_input)
    return dflat, list(backward._ordered(c.m, "enc", grads, ctx.meta)), None, None


encoder_stack.register_autograd(_enc_backward, setup_context=_enc_setup)


# ------------------------------------------------------------------------------------------------
# decoder
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("titok_b200::decoder_stack", mutates_args=())
def decoder_stack(codes: torch.Tensor, params: Sequence[torch.Tensor], handle: int, train: bool) -> torch.Tensor:
    """codes [T, token_size] -> the reconstructed clips back to back, bf16 [sum 3*T*H*W] (blocks.py:148-177)."""
    c = _call(handle)
    engine.require_cuda(codes.device)
    # always a private copy of the codes: they may be a view of the device-wide 'codes' arena (see backward.DecoderFn)
    codes_b = codes.detach().to(bf16).clone(memory_format=torch.contiguous_format)
    if train:
        out, tape = backward.decoder_forward_train(c.m, c.dp, codes_b)
        c.tape = tape
        return out
    out = torch.empty((c.dp.plan.total_numel,), dtype=bf16, device=codes.device)
    engine.decoder_launch(c.m, c.dp, codes_b, out)
    return out


@decoder_stack.register_fake
def _(codes, params, handle, train):
    c = _call(handle)
    return codes.new_empty((c.dp.plan.total_numel,), dtype=bf16)


def _dec_setup(ctx, inputs, output):
    codes, params, handle, train = inputs
    c = _call(handle)
    if not train or c.tape is None:
        raise RuntimeError("titok_b200::decoder_stack was recorded by autograd without its training forward (train=False)")
    ctx.call, ctx.tape, c.tape = c, c.tape, None
    ctx.codes_dtype, ctx.need_codes = codes.dtype, codes.requires_grad
    ctx.meta = backward._meta(c.m)


def _dec_backward(ctx, dout):
    c = ctx.call
    grads, dcodes = backward.decoder_backward(c.m, c.dp, ctx.tape, dout.to(bf16).contiguous())
    dc = dcodes.to(ctx.codes_dtype) if ctx.need_codes else None
    return dc, list(backward._ordered(c.m, "dec", grads, ctx.meta)), None, None


decoder_stack.register_autograd(_dec_backward, setup_context=_dec_setup)


def encoder_call(module, dp, consts, flat: torch.Tensor, train: bool):
    return torch.ops.titok_b200.encoder_stack(flat, backward.stack_params(module), handle_for(module, dp, consts), train)


def decoder_call(module, dp, codes: torch.Tensor, train: bool) -> torch.Tensor:
    return torch.ops.titok_b200.decoder_stack(codes, backward.stack_params(module), handle_for(module, dp), train)
