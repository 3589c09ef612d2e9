"""Training path: launch sequences that record what the backward needs, the backward launch sequences, and the
torch.autograd.Function wrappers that plug them into autograd (so `train.py:68-80` -- forward under bf16 autocast,
`manual_backward(loss)`, `grad_norm(self.model)`, AdamW on `model.parameters()` -- works on the drop-in modules).

What the reference gets from torch.autograd for model/base/blocks.py:71-104,148-177 and
model/base/transformer.py:47-56,85-104,126-146 is computed here by the sm_100a kernels of csrc/attn_bwd.cu (attention),
csrc/wgrad.cu (weight gradients), csrc/gemm.cu (input gradients: `ttk_gemm_bf16(..., w_is_kn=1)`) and
csrc/bwd_rows.cu (RMSNorm / GEGLU / gathers / small projections). Activation gradients are bf16 (as under the
reference's autocast), parameter gradients fp32 in the reference's parameter layout.

Saved activations live in ordinary torch tensors (not in the shared workspace arenas), so several forwards may be
in flight before their backwards (generator + discriminator passes, gradient accumulation).
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .engine import NATIVE_SEQ, DevicePlan, cached_named_params, PreparedStack, _ptr, _stream, _vp, key_norms, layers_desc, patch_feature_perm_on, prepared

bf16 = torch.bfloat16

# Training forward + backward of the ENCODER's last layer on the latent rows only (the head reads nothing else,
# blocks.py:101; see engine._layer_latent for the inference side). The layers before it run through the native sequencers
# as before (n_layers - 1 of them); the last layer is enqueued kernel by kernel from Python (about 45 launches, forward +
# backward), which costs host time: "auto" (default) takes this path when the packed batch has at least
# TRAIN_TAIL_MIN_ROWS rows -- where the step is bound by the GPU (16 clips A per GPU: 8.38 -> 7.87 ms per step) -- and
# runs every row through the layer, as the reference does, for small host-bound batches. True / False (env 1 / 0) force it.
import os as _os

_env_tail = _os.environ.get("TTK_TRAIN_LATENT_TAIL", "auto")
TRAIN_LATENT_TAIL = {"0": False, "1": True}.get(_env_tail, "auto")
TRAIN_TAIL_MIN_ROWS = int(_os.environ.get("TTK_TRAIN_TAIL_MIN_ROWS", "24000"))


def latent_tail_active(packed_rows: int) -> bool:
    """Whether a training forward over `packed_rows` rows carries only the latent rows through the encoder's last layer."""
    if TRAIN_LATENT_TAIL == "auto":
        return packed_rows >= TRAIN_TAIL_MIN_ROWS
    return bool(TRAIN_LATENT_TAIL)


_IDENT: Dict[str, torch.Tensor] = {}


def _ident(n: int, device) -> torch.Tensor:
    """int32 [n] = 0..n-1 on `device` (row map of compact buffers), cached."""
    t = _IDENT.get(str(device))
    if t is None or t.numel() < n:
        if torch.cuda.is_current_stream_capturing():  # (memory allocated under capture belongs to the graph: not cached)
            return torch.arange(n, dtype=torch.int32, device=device)
        t = torch.arange(max(n, 4096), dtype=torch.int32, device=device)
        _IDENT[str(device)] = t
    return t[:n]


def _new(shape, device, dtype=bf16) -> torch.Tensor:
    return torch.empty(shape, dtype=dtype, device=device)


class Tape:
    """Activations of one stack forward, kept for its backward."""

    def __init__(self):
        self.lt = None  # LayerTape of the transformer layers
        self.tail = None  # encoder with TRAIN_LATENT_TAIL: tensors of the last layer (_last_layer_train)
        self.latent_tail = False
        self.t: Dict[str, torch.Tensor] = {}
        self.weights_sig = None  # (parameter signature, optimizer-step counter) the forward ran with


def _weights_sig(m):
    """What identifies the parameter values a stack's prepared bf16 copies were made from: parameter objects / versions /
    storages (PreparedStack._signature) and the process-wide optimizer-step counter (fused optimizers do not bump
    versions)."""
    from . import engine

    ps = m.__dict__.get("_ttk_prepared")
    return (ps._signature() if ps is not None else None, engine._OPT_STEPS[0])


def _check_tape_weights(m, tape: "Tape") -> None:
    """The backward kernels read the stack's CURRENT prepared weights: if the parameters changed between a forward and its
    backward (an optimizer step in between, retain_graph reuse after a step), old activations would be combined with new
    weights without any error. PyTorch raises a version-counter error in that situation; so do we."""
    if tape.weights_sig is not None and tape.weights_sig != _weights_sig(m):
        raise RuntimeError("titok_video_b200: the parameters of this stack were modified (optimizer step / in-place edit) "
                           "between the forward that recorded this graph and its backward")


# --------------------------------------------------------------------------------------------------
# forward (unfused where the backward needs the intermediate)
# --------------------------------------------------------------------------------------------------
def _gemm(st, a: torch.Tensor, wmat: torch.Tensor, out: torch.Tensor, N: int, K: int, bias=None, kn: int = 0) -> None:
    """`st`: the stream handle, looked up ONCE per launch sequence (torch.cuda.current_stream() costs ~14 us a call)."""
    _lib.call("ttk_gemm_bf16", _ptr(a), a.stride(0), _ptr(wmat), wmat.stride(0), a.shape[0], N, K, _ptr(bias), _ptr(out),
              out.stride(0), _vp(0), kn, st)


def _wgrad(st, dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor) -> None:
    """dw [n_out, k_in] fp32 += dy^T x."""
    _lib.call("ttk_gemm_wgrad", _ptr(dy), dy.stride(0), _ptr(x), x.stride(0), dy.shape[0], dw.shape[0], dw.shape[1],
              _ptr(dw), dw.stride(0), st)


def _rmsnorm_bwd(st, x, w, dy, dx, dw, *, y=None, alpha=1.0, add=None, add_scale=1.0, sel=None, w2=None, dw2=None) -> None:
    M, width = x.shape
    _lib.call("ttk_rmsnorm_bwd", _ptr(x), _ptr(y), float(alpha), _ptr(w), _ptr(w2), _ptr(sel), _ptr(dy), _ptr(add),
              float(add_scale), _ptr(dx), _ptr(dw), _ptr(dw2), M, width, x.stride(0), st)


# block order inside a layer's slab (ttk_layers_fwd_train) and inside the backward work buffer (ttk_layers_bwd)
F_QKV, F_ATT, F_O, F_YA, F_XF, F_XNF, F_H12, F_H, F_YF, F_XN, F_XNN = range(11)
B_DUF, B_DH, B_DH12, B_DXN, B_GF, B_DUA, B_DATT, B_DQKV, B_DO, B_G0, B_G1 = range(11)
_LAYOUTS: Dict[tuple, tuple] = {}


def _layouts(M: int, w: int, gqa: int, inner: int):
    """(forward cols, forward offsets, per-layer elements, backward cols, backward offsets, work elements): element
    offsets of the 2-D bf16 blocks, every block 16-byte aligned. Cached per shape."""
    key = (M, w, gqa, inner)
    hit = _LAYOUTS.get(key)
    if hit is None:
        ldq = 2 * w + 2 * gqa

        def lay(cols):
            offs, tot = [], 0
            for c in cols:
                offs.append(tot)
                tot += (M * c + 7) // 8 * 8
            return np.asarray(offs, dtype=np.int64), tot

        fcols = [ldq, w, w, w, w, w, 2 * inner, inner, w, w, w]
        bcols = [w, inner, 2 * inner, w, w, w, w, ldq, w, w, w]
        foffs, ftot = lay(fcols)
        boffs, btot = lay(bcols)
        if len(_LAYOUTS) > 256:
            _LAYOUTS.clear()
        hit = (fcols, foffs, ftot, bcols, boffs, btot)
        _LAYOUTS[key] = hit
    return hit


class LayerTape:
    """Activations of all layers of one stack forward: one bf16 slab [n_layers][per_layer] + the log-sum-exps."""

    def __init__(self, slab, per_layer, cols, offs, lse, x0, xn0, M):
        self.slab, self.per_layer, self.cols, self.offs, self.lse, self.x0, self.xn0, self.M = slab, per_layer, cols, offs, lse, x0, xn0, M

    def view(self, layer: int, col: int) -> torch.Tensor:
        o = layer * self.per_layer + int(self.offs[col])
        return self.slab[o:o + self.M * self.cols[col]].view(self.M, self.cols[col])


def _layers_train(m, W: PreparedStack, dp: DevicePlan, x: torch.Tensor, xn: torch.Tensor, tape: Tape,
                  n_run: Optional[int] = None):
    """ResidualAttentionBlock.forward (transformer.py:126-146), recording per-layer activations. n_run: only the first
    n_run layers (the encoder's last layer then follows in _last_layer_train)."""
    M, w = x.shape
    dev = x.device
    hq, hkv = m.heads
    gqa = hkv * 64
    inner = m.inner_dim
    L = m.num_layers
    n_run = L if n_run is None else n_run
    alpha = float(2 * L)
    st = _stream()
    fcols, foffs, per_layer, _, _, _ = _layouts(M, w, gqa, inner)
    slab = torch.empty(max(n_run, 1) * per_layer, dtype=bf16, device=dev)
    lse_all = torch.empty((L, hq, M), dtype=torch.float32, device=dev)
    lt = LayerTape(slab, per_layer, fcols, foffs, lse_all, x, xn, M)
    tape.lt = lt
    if n_run <= 0:
        return x, xn
    if NATIVE_SEQ and not _lib.profiling():
        d = layers_desc(m, W, dp, M)
        d.n_layers = n_run
        _lib.call("ttk_layers_fwd_train", ctypes.byref(d), _ptr(x), _ptr(xn), _ptr(slab), per_layer, _vp(foffs.ctypes.data),
                  _ptr(lse_all), st, launches=8 * n_run)
        return lt.view(n_run - 1, F_XN), lt.view(n_run - 1, F_XNN)
    work = dp.attn_work(hq, hkv)
    knorm = key_norms(dp, M, hkv)
    scale = 1.0 / math.sqrt(64.0)
    T = W.t
    for i in range(n_run):
        mode = 0 if i == 0 else 1
        qkv, att, o, lse = lt.view(i, F_QKV), lt.view(i, F_ATT), lt.view(i, F_O), lse_all[i]
        _lib.call("ttk_gemm_qkv_rope", _ptr(xn), xn.stride(0), _ptr(T[f"to_qkv{i}"]), w, M, w, w, gqa, _ptr(dp.rope),
                  _ptr(qkv), qkv.stride(0), _ptr(knorm), st)
        _lib.call("ttk_attn_varlen_fwd_train", _ptr(qkv), qkv.stride(0), M, w, gqa, _ptr(work), work.shape[0], scale,
                  _ptr(att), att.stride(0), _ptr(o), _ptr(lse), _ptr(knorm), st)
        y_a = lt.view(i, F_YA)
        _gemm(st, att, T[f"out_proj{i}"], y_a, w, w)
        x_f, xn_f = lt.view(i, F_XF), lt.view(i, F_XNF)
        _lib.call("ttk_resid_norm", _ptr(x), _ptr(y_a), _ptr(x_f), _ptr(xn_f), _ptr(T.get(f"attn_post_ln{i}")),
                  _ptr(T[f"ffn_norm{i}"]), alpha, mode, M, w, w, st)
        h12 = lt.view(i, F_H12)
        _gemm(st, xn_f, T[f"w12_{i}"], h12, 2 * inner, w)
        h = lt.view(i, F_H)
        _lib.call("ttk_geglu_fwd", _ptr(h12), h12.stride(0), inner, _ptr(h), h.stride(0), M, st)
        y_f = lt.view(i, F_YF)
        _gemm(st, h, T[f"w3_{i}"], y_f, w, inner)
        x_n, xn_n = lt.view(i, F_XN), lt.view(i, F_XNN)
        w_next = T[f"pre_ln{i + 1}"] if i + 1 < L else T["ln_post"]
        _lib.call("ttk_resid_norm", _ptr(x_f), _ptr(y_f), _ptr(x_n), _ptr(xn_n), _ptr(T.get(f"ffd_post_ln{i}")),
                  _ptr(w_next), alpha, mode, M, w, w, st)
        x, xn = x_n, xn_n
    return x, xn


def _use_latent_tail(m, dp: DevicePlan) -> bool:
    return latent_tail_active(dp.plan.M) and dp.plan.T > 0 and m.num_layers > 1


def _last_layer_train(m, W: PreparedStack, dp: DevicePlan, x_a: torch.Tensor, xn_a: torch.Tensor, tape: Tape):
    """The encoder's last layer, recorded for the backward, on what the head reads: qkv projection on all M rows (every row
    is a key), attention for the query tiles that hold latent rows, everything behind it on the T latent rows gathered
    into compact [T, .] tensors. Returns (x_out, RMSNorm(x_out) * ln_post.weight) of the latent rows, in token order."""
    M, w = x_a.shape
    dev = x_a.device
    hq, hkv = m.heads
    gqa = hkv * 64
    inner = m.inner_dim
    L = m.num_layers
    i = L - 1
    mode = 0 if i == 0 else 1
    alpha = float(2 * L)
    Tn = dp.plan.T
    st = _stream()
    T = W.t
    ldq = 2 * w + 2 * gqa
    qkv = _new((M, ldq), dev)
    att = _new((M, w), dev)                                   # only its latent rows are written and read
    o = torch.zeros((M, w), dtype=bf16, device=dev)           # the backward's row pass reads every row: zeros elsewhere
    lse = tape.lt.lse[i]
    lse.zero_()
    work = dp.attn_work_latent(hq, hkv)
    knorm = key_norms(dp, M, hkv)
    _lib.call("ttk_gemm_qkv_rope", _ptr(xn_a), xn_a.stride(0), _ptr(T[f"to_qkv{i}"]), w, M, w, w, gqa, _ptr(dp.rope),
              _ptr(qkv), ldq, _ptr(knorm), st)
    _lib.call("ttk_attn_varlen_fwd_train", _ptr(qkv), ldq, M, w, gqa, _ptr(work), work.shape[0], 1.0 / math.sqrt(64.0),
              _ptr(att), w, _ptr(o), _ptr(lse), _ptr(knorm), st)
    att_c, x_c = _new((Tn, w), dev), _new((Tn, w), dev)
    _lib.call("ttk_gather_rows", _ptr(att), w, _ptr(dp.latent_row), _ptr(att_c), w, Tn, w, st)
    _lib.call("ttk_gather_rows", _ptr(x_a), x_a.stride(0), _ptr(dp.latent_row), _ptr(x_c), w, Tn, w, st)
    y_a = _new((Tn, w), dev)
    _gemm(st, att_c, T[f"out_proj{i}"], y_a, w, w)
    x_f, xn_f = _new((Tn, w), dev), _new((Tn, w), dev)
    _lib.call("ttk_resid_norm", _ptr(x_c), _ptr(y_a), _ptr(x_f), _ptr(xn_f), _ptr(T.get(f"attn_post_ln{i}")),
              _ptr(T[f"ffn_norm{i}"]), alpha, mode, Tn, w, w, st)
    h12 = _new((Tn, 2 * inner), dev)
    _gemm(st, xn_f, T[f"w12_{i}"], h12, 2 * inner, w)
    h = _new((Tn, inner), dev)
    _lib.call("ttk_geglu_fwd", _ptr(h12), 2 * inner, inner, _ptr(h), inner, Tn, st)
    y_f = _new((Tn, w), dev)
    _gemm(st, h, T[f"w3_{i}"], y_f, w, inner)
    x_n, xn_n = _new((Tn, w), dev), _new((Tn, w), dev)
    _lib.call("ttk_resid_norm", _ptr(x_f), _ptr(y_f), _ptr(x_n), _ptr(xn_n), _ptr(T.get(f"ffd_post_ln{i}")),
              _ptr(T["ln_post"]), alpha, mode, Tn, w, w, st)
    tape.tail = dict(qkv=qkv, o=o, att_c=att_c, x_a=x_a, xn_a=xn_a, x_c=x_c, y_a=y_a, x_f=x_f, xn_f=xn_f, h12=h12, h=h,
                     y_f=y_f)
    return x_n, xn_n


def _last_layer_backward(m, W: PreparedStack, dp: DevicePlan, tape: Tape, g: torch.Tensor, grads) -> torch.Tensor:
    """g = dL/dx_out of the latent rows [T, w] -> dL/dx at the input of the last layer for all M packed rows. The part
    behind the attention runs on T rows (same kernels and order as _layers_backward); its two results -- the gradient of
    the attention output and of the residual branch -- go back to their packed rows (zeros elsewhere), the attention
    backward follows work lists restricted to the latent query rows (dK / dV of every key row, dQ of the latent tiles), and
    the qkv projection's backward runs on all rows again."""
    Tn, w = g.shape
    dev = g.device
    hq, hkv = m.heads
    gqa = hkv * 64
    inner = m.inner_dim
    L = m.num_layers
    i = L - 1
    mode = 0 if i == 0 else 1
    alpha = float(2 * L)
    c = alpha if mode == 1 else 1.0
    st = _stream()
    T = W.t
    t = tape.tail
    x_a, xn_a, qkv = t["x_a"], t["xn_a"], t["qkv"]
    M = x_a.shape[0]
    ldq = 2 * w + 2 * gqa
    # ---- GEGLU block
    if mode == 1:
        du = _new((Tn, w), dev)
        _rmsnorm_bwd(st, t["x_f"], T[f"ffd_post_ln{i}"], g, du, grads[f"ffd_post_ln{i}"], y=t["y_f"], alpha=alpha)
    else:
        du = g
    dh = _new((Tn, inner), dev)
    _gemm(st, du, T[f"w3_{i}"], dh, inner, w, kn=1)
    _wgrad(st, du, t["h"], grads[f"w3_{i}"])
    dh12 = _new((Tn, 2 * inner), dev)
    _lib.call("ttk_geglu_bwd", _ptr(t["h12"]), 2 * inner, inner, _ptr(dh), inner, _ptr(dh12), 2 * inner, Tn, st)
    dxn = _new((Tn, w), dev)
    _gemm(st, dh12, T[f"w12_{i}"], dxn, w, 2 * inner, kn=1)
    _wgrad(st, dh12, t["xn_f"], grads[f"w12_{i}"])
    g_f = _new((Tn, w), dev)
    _rmsnorm_bwd(st, t["x_f"], T[f"ffn_norm{i}"], dxn, g_f, grads[f"ffn_norm{i}"], add=du, add_scale=c)
    # ---- attention block
    if mode == 1:
        du_a = _new((Tn, w), dev)
        _rmsnorm_bwd(st, t["x_c"], T[f"attn_post_ln{i}"], g_f, du_a, grads[f"attn_post_ln{i}"], y=t["y_a"], alpha=alpha)
    else:
        du_a = g_f
    d_att_c = _new((Tn, w), dev)
    _gemm(st, du_a, T[f"out_proj{i}"], d_att_c, w, w, kn=1)
    _wgrad(st, du_a, t["att_c"], grads[f"out_proj{i}"])
    # ---- back to the packed rows: zeros wherever no gradient arrives
    d_att = torch.zeros((M, w), dtype=bf16, device=dev)
    du_full = torch.zeros((M, w), dtype=bf16, device=dev)
    _lib.call("ttk_scatter_rows", _ptr(d_att_c), w, _ptr(dp.latent_row), _ptr(d_att), w, Tn, w, st)
    _lib.call("ttk_scatter_rows", _ptr(du_a), w, _ptr(dp.latent_row), _ptr(du_full), w, Tn, w, st)
    dqkv = torch.zeros((M, ldq), dtype=bf16, device=dev)  # dQ of the other query tiles, dK / dV of clips without tokens
    dO = _new((M, w), dev)
    delta = _new((hq, M), dev, torch.float32)
    lse = tape.lt.lse[i]
    _lib.call("ttk_attn_bwd_prep", _ptr(d_att), w, _ptr(t["o"]), w, _ptr(qkv), ldq, M, w, _ptr(dO), w, _ptr(dqkv), ldq,
              _ptr(delta), st)
    wk_dkv, wk_dq = dp.attn_bwd_work_latent(hq, hkv)
    for name, wk in (("ttk_attn_bwd_dkv", wk_dkv), ("ttk_attn_bwd_dq", wk_dq)):
        _lib.call(name, _ptr(qkv), ldq, _ptr(dO), w, M, w, gqa, _ptr(wk), wk.shape[0], _ptr(lse), _ptr(delta),
                  _ptr(dp.rope), 1.0 / math.sqrt(64.0), _ptr(dqkv), ldq, st)
    dxn_full = _new((M, w), dev)
    _gemm(st, dqkv, T[f"to_qkv{i}"], dxn_full, w, ldq, kn=1)
    _wgrad(st, dqkv, xn_a, grads[f"to_qkv{i}"])
    g_next = _new((M, w), dev)
    _rmsnorm_bwd(st, x_a, T[f"pre_ln{i}"], dxn_full, g_next, grads[f"pre_ln{i}"], add=du_full, add_scale=c)
    return g_next


def encoder_forward_train(m, dp: DevicePlan, clips_flat: torch.Tensor, fsq_consts):
    """TiTokEncoder.forward recording a tape. Returns (z, codes, idx, tape); z/codes bf16 [T, ts]."""
    W = prepared(m, "enc", force=True)  # a training step always re-reads the parameters (see PreparedStack.refresh)
    pl = dp.plan
    M, G, Tn, w = pl.M, pl.G, pl.T, m.width
    P0, P1, P2 = pl.patch_size
    feat = pl.channels * P0 * P1 * P2
    dev = clips_flat.device
    st = _stream()
    tape = Tape()
    T = W.t
    patches = _new((G, feat), dev)
    proj = _new((G, w), dev)
    x, xn, e0 = _new((M, w), dev), _new((M, w), dev), _new((M, w), dev)
    _lib.call("ttk_patchify", _ptr(clips_flat), _ptr(dp.geom), pl.channels, P0, P1, P2, _ptr(patches), feat, G, st)
    _gemm(st, patches, T["proj_in_w"], proj, w, feat, bias=T["proj_in_b"])
    _lib.call("ttk_enc_embed_train", _ptr(proj), w, _ptr(dp.enc_src_row), _ptr(T["mask_token"]), _ptr(T["ln_pre_t"]),
              _ptr(T["ln_pre_p"]), _ptr(T["pre_ln0"]), _ptr(x), _ptr(xn), _ptr(e0), M, w, w, st)
    tail = _use_latent_tail(m, dp)
    if tail:
        x_a, xn_a = _layers_train(m, W, dp, x, xn, tape, n_run=m.num_layers - 1)
        x_fin, xn_fin = _last_layer_train(m, W, dp, x_a, xn_a, tape)  # latent rows only, in token order
    else:
        x_fin, xn_fin = _layers_train(m, W, dp, x, xn, tape)
    ts = m.token_size
    z = _new((max(Tn, 1), ts), dev)
    codes = _new((max(Tn, 1), ts), dev)
    idx = _new((max(Tn, 1),), dev, torch.int32)
    half_l, offset, shift, half_width, basis, levels = fsq_consts
    _lib.call("ttk_enc_head_fsq", _ptr(xn_fin), w, _vp(0) if tail else _ptr(dp.latent_row), _ptr(T["ln_post"]), 1, _ptr(T["proj_out_w"]),
              _ptr(T["proj_out_b"]), ts, _ptr(z), _ptr(codes), _ptr(idx), Tn, w, half_l, offset, shift, half_width,
              basis, levels, st)
    tape.t.update(patches=patches, e0=e0, x_fin=x_fin, xn_fin=xn_fin)
    tape.latent_tail = tail
    tape.weights_sig = _weights_sig(m)
    return z[:Tn], codes[:Tn], idx[:Tn], tape


def decoder_forward_train(m, dp: DevicePlan, codes: torch.Tensor):
    """TiTokDecoder.forward recording a tape. Returns (out_flat bf16 [sum 3*T*H*W], tape)."""
    W = prepared(m, "dec", force=True)
    pl = dp.plan
    M, G, w = pl.M, pl.G, m.width
    P0, P1, P2 = pl.patch_size
    feat = pl.channels * P0 * P1 * P2
    dev = codes.device
    st = _stream()
    tape = Tape()
    T = W.t
    x, xn, e0 = _new((M, w), dev), _new((M, w), dev), _new((M, w), dev)
    _lib.call("ttk_dec_embed_train", _ptr(codes), m.token_size, _ptr(dp.dec_src_row), _ptr(T["proj_in_w"]),
              _ptr(T["proj_in_b"]), _ptr(T["mask_token"]), _ptr(T["ln_pre_t"]), _ptr(T["ln_pre_p"]), _ptr(T["pre_ln0"]),
              _ptr(x), _ptr(xn), _ptr(e0), M, w, w, st)
    x_fin, xn_fin = _layers_train(m, W, dp, x, xn, tape)
    rows = _new((M, feat), dev)
    _gemm(st, xn_fin, T["proj_out_w"], rows, feat, w, bias=T["proj_out_b"])
    out = _new((pl.total_numel,), dev)
    _lib.call("ttk_unpatchify", _ptr(rows), feat, _ptr(dp.patch_row), _ptr(dp.geom), pl.channels, P0, P1, P2, _ptr(out),
              G, st)
    tape.t.update(codes=codes, e0=e0, x_fin=x_fin, xn_fin=xn_fin)
    tape.weights_sig = _weights_sig(m)
    return out, tape


# --------------------------------------------------------------------------------------------------
# backward
# --------------------------------------------------------------------------------------------------
_GRAD_KEYS = ("ffd_post_ln{i}", "w3_{i}", "w12_{i}", "ffn_norm{i}", "attn_post_ln{i}", "out_proj{i}", "to_qkv{i}", "pre_ln{i}")


def _zero_grads(W: PreparedStack) -> Dict[str, torch.Tensor]:
    """fp32 zero buffers with the shapes of the prepared (kernel-layout) parameters: views of ONE flat buffer (one
    memset per stack; every view starts on a 16-byte boundary for the vector reductions of ttk_gemm_wgrad)."""
    lay = W.__dict__.get("_grad_layout")
    if lay is None or lay[0] != tuple((k, tuple(v.shape)) for k, v in W.t.items()):
        offs, total = {}, 0
        for k, v in W.t.items():
            offs[k] = total
            total += (v.numel() + 3) // 4 * 4
        L = W.module.num_layers
        tab = np.full((L, len(_GRAD_KEYS)), -1, dtype=np.int64)  # element offsets of the layers' gradients, -1 = absent
        for i in range(L):
            for j, pat in enumerate(_GRAD_KEYS):
                k = pat.format(i=i)
                if k in offs:
                    tab[i, j] = offs[k]
        lay = (tuple((k, tuple(v.shape)) for k, v in W.t.items()), offs, total, tab)
        W._grad_layout = lay
    _, offs, total, _ = lay
    dev = next(iter(W.t.values())).device
    flat = torch.zeros(total, dtype=torch.float32, device=dev)
    grads = {k: flat[offs[k]:offs[k] + v.numel()].view(v.shape) for k, v in W.t.items()}
    grads["__flat__"] = flat
    return grads


def _layers_backward(m, W: PreparedStack, dp: DevicePlan, tape: Tape, g: torch.Tensor, grads,
                     n_run: Optional[int] = None) -> torch.Tensor:
    """g = dL/dx at the output of the last layer -> dL/dx at the input of layer 0. n_run: g belongs to the output of layer
    n_run - 1 (the forward ran its last layer through _last_layer_train)."""
    M, w = g.shape
    dev = g.device
    hq, hkv = m.heads
    gqa = hkv * 64
    inner = m.inner_dim
    L = m.num_layers
    n_run = L if n_run is None else n_run
    if n_run <= 0:
        return g
    alpha = float(2 * L)
    st = _stream()
    lt: LayerTape = tape.lt
    _, _, _, bcols, boffs, btot = _layouts(M, w, gqa, inner)
    work = torch.empty(btot, dtype=bf16, device=dev)  # temporaries are reused by every layer (stream order)
    delta = _new((hq, M), dev, torch.float32)

    def wv(col: int) -> torch.Tensor:
        o = int(boffs[col])
        return work[o:o + M * bcols[col]].view(M, bcols[col])

    if NATIVE_SEQ and not _lib.profiling():
        d = layers_desc(m, W, dp, M, backward=True)
        d.n_layers = n_run
        flat = grads["__flat__"]
        tab = W._grad_layout[3]
        gptr = np.where(tab >= 0, flat.data_ptr() + 4 * tab, 0).astype(np.int64)
        g_out = ctypes.c_void_p(0)
        n_launch = sum(14 + (2 if i > 0 else 0) for i in range(n_run))
        _lib.call("ttk_layers_bwd", ctypes.byref(d), _ptr(lt.x0), _ptr(lt.xn0), _ptr(lt.slab), lt.per_layer,
                  _vp(lt.offs.ctypes.data), _ptr(lt.lse), _ptr(g), _ptr(work), _vp(boffs.ctypes.data), _ptr(delta),
                  _vp(gptr.ctypes.data), ctypes.byref(g_out), st, launches=n_launch)
        res = wv(B_G0) if g_out.value == wv(B_G0).data_ptr() else wv(B_G1)
        assert res.data_ptr() == g_out.value
        return res
    scale = 1.0 / math.sqrt(64.0)
    T = W.t
    wk_dkv, wk_dq = dp.attn_bwd_work(hq, hkv)
    for i in reversed(range(n_run)):
        mode = 0 if i == 0 else 1
        c = alpha if mode == 1 else 1.0
        x_a = lt.x0 if i == 0 else lt.view(i - 1, F_XN)
        xn_a = lt.xn0 if i == 0 else lt.view(i - 1, F_XNN)
        x_f = lt.view(i, F_XF)
        # ---- GEGLU block: x_out = x_f + y_f | RMSNorm(alpha x_f + y_f)
        if mode == 1:
            du = wv(B_DUF)
            _rmsnorm_bwd(st, x_f, T[f"ffd_post_ln{i}"], g, du, grads[f"ffd_post_ln{i}"], y=lt.view(i, F_YF), alpha=alpha)
        else:
            du = g
        dh = wv(B_DH)
        _gemm(st, du, T[f"w3_{i}"], dh, inner, w, kn=1)
        _wgrad(st, du, lt.view(i, F_H), grads[f"w3_{i}"])
        dh12 = wv(B_DH12)
        _lib.call("ttk_geglu_bwd", _ptr(lt.view(i, F_H12)), 2 * inner, inner, _ptr(dh), inner, _ptr(dh12), 2 * inner, M, st)
        dxn = wv(B_DXN)
        _gemm(st, dh12, T[f"w12_{i}"], dxn, w, 2 * inner, kn=1)
        _wgrad(st, dh12, lt.view(i, F_XNF), grads[f"w12_{i}"])
        g_f = wv(B_GF)
        _rmsnorm_bwd(st, x_f, T[f"ffn_norm{i}"], dxn, g_f, grads[f"ffn_norm{i}"], add=du, add_scale=c)
        # ---- attention block: x_f = x_a + y_a | RMSNorm(alpha x_a + y_a)
        if mode == 1:
            du = wv(B_DUA)
            _rmsnorm_bwd(st, x_a, T[f"attn_post_ln{i}"], g_f, du, grads[f"attn_post_ln{i}"], y=lt.view(i, F_YA), alpha=alpha)
        else:
            du = g_f
        d_att = wv(B_DATT)
        _gemm(st, du, T[f"out_proj{i}"], d_att, w, w, kn=1)
        _wgrad(st, du, lt.view(i, F_ATT), grads[f"out_proj{i}"])
        qkv = lt.view(i, F_QKV)
        dqkv, dO = wv(B_DQKV), wv(B_DO)
        _lib.call("ttk_attn_bwd_prep", _ptr(d_att), w, _ptr(lt.view(i, F_O)), w, _ptr(qkv), qkv.stride(0), M, w, _ptr(dO), w,
                  _ptr(dqkv), dqkv.stride(0), _ptr(delta), st)
        for name, wk in (("ttk_attn_bwd_dkv", wk_dkv), ("ttk_attn_bwd_dq", wk_dq)):
            _lib.call(name, _ptr(qkv), qkv.stride(0), _ptr(dO), w, M, w, gqa, _ptr(wk), wk.shape[0], _ptr(lt.lse[i]),
                      _ptr(delta), _ptr(dp.rope), scale, _ptr(dqkv), dqkv.stride(0), st)
        _gemm(st, dqkv, T[f"to_qkv{i}"], dxn, w, 2 * w + 2 * gqa, kn=1)
        _wgrad(st, dqkv, xn_a, grads[f"to_qkv{i}"])
        g_next = wv(B_G1 if ((L - 1 - i) & 1) else B_G0)
        _rmsnorm_bwd(st, x_a, T[f"pre_ln{i}"], dxn, g_next, grads[f"pre_ln{i}"], add=du, add_scale=c)
        g = g_next
    return g


def encoder_backward(m, dp: DevicePlan, tape: Tape, dz: torch.Tensor, need_input_grad: bool = False):
    """dz bf16 [T, ts] -> (grads in kernel layout, d clips_flat or None)."""
    _check_tape_weights(m, tape)
    W = prepared(m, "enc")
    pl = dp.plan
    M, G, Tn, w = pl.M, pl.G, pl.T, m.width
    P0, P1, P2 = pl.patch_size
    feat = pl.channels * P0 * P1 * P2
    dev = dz.device
    st = _stream()
    T = W.t
    grads = _zero_grads(W)
    if getattr(tape, "latent_tail", False):
        # the forward carried only the latent rows through the last layer: x_fin / xn_fin are [T, w] in token order
        dxn = torch.zeros((Tn, w), dtype=bf16, device=dev)
        _lib.call("ttk_head_bwd", _ptr(dz), m.token_size, _ptr(tape.t["xn_fin"]), w, _ptr(_ident(Tn, dev)), _ptr(T["proj_out_w"]),
                  _ptr(dxn), _ptr(grads["proj_out_w"]), _ptr(grads["proj_out_b"]), Tn, w, st)
        g_c = _new((Tn, w), dev)
        _rmsnorm_bwd(st, tape.t["x_fin"], T["ln_post"], dxn, g_c, grads["ln_post"])
        g = _last_layer_backward(m, W, dp, tape, g_c, grads)
        g0 = _layers_backward(m, W, dp, tape, g, grads, n_run=m.num_layers - 1)
    else:
        dxn = torch.zeros((M, w), dtype=bf16, device=dev)
        _lib.call("ttk_head_bwd", _ptr(dz), m.token_size, _ptr(tape.t["xn_fin"]), w, _ptr(dp.latent_row), _ptr(T["proj_out_w"]),
                  _ptr(dxn), _ptr(grads["proj_out_w"]), _ptr(grads["proj_out_b"]), Tn, w, st)
        g = _new((M, w), dev)
        _rmsnorm_bwd(st, tape.t["x_fin"], T["ln_post"], dxn, g, grads["ln_post"])
        g0 = _layers_backward(m, W, dp, tape, g, grads)
    # embed: latent rows (enc_src_row < 0) went through ln_pre_t, patch rows through ln_pre_p (blocks.py:95-97)
    d_e0 = _new((M, w), dev)
    _rmsnorm_bwd(st, tape.t["e0"], T["ln_pre_p"], g0, d_e0, grads["ln_pre_p"], sel=dp.enc_src_row, w2=T["ln_pre_t"],
                 dw2=grads["ln_pre_t"])
    _lib.call("ttk_colsum", _ptr(d_e0), w, M, w, _vp(0), _ptr(grads["mask_token"]), st)
    dproj = _new((G, w), dev)
    _lib.call("ttk_gather_rows", _ptr(d_e0), w, _ptr(dp.patch_row), _ptr(dproj), w, G, w, st)
    _lib.call("ttk_colsum", _ptr(dproj), w, G, w, _ptr(grads["proj_in_b"]), _vp(0), st)
    _wgrad(st, dproj, tape.t["patches"], grads["proj_in_w"])
    dflat = None
    if need_input_grad:
        dpatches = _new((G, feat), dev)
        _gemm(st, dproj, T["proj_in_w"], dpatches, feat, w, kn=1)
        ident = torch.arange(G, dtype=torch.int32, device=dev)
        dflat = _new((pl.total_numel,), dev)
        _lib.call("ttk_unpatchify", _ptr(dpatches), feat, _ptr(ident), _ptr(dp.geom), pl.channels, P0, P1, P2,
                  _ptr(dflat), G, st)
    return grads, dflat


def decoder_backward(m, dp: DevicePlan, tape: Tape, dout: torch.Tensor):
    """dout bf16 flat [sum 3*T*H*W] -> (grads in kernel layout, dcodes fp32 [T, ts])."""
    _check_tape_weights(m, tape)
    W = prepared(m, "dec")
    pl = dp.plan
    M, G, Tn, w = pl.M, pl.G, pl.T, m.width
    P0, P1, P2 = pl.patch_size
    feat = pl.channels * P0 * P1 * P2
    dev = dout.device
    st = _stream()
    T = W.t
    grads = _zero_grads(W)
    d_rows = _new((G, feat), dev)
    _lib.call("ttk_patchify", _ptr(dout), _ptr(dp.geom), pl.channels, P0, P1, P2, _ptr(d_rows), feat, G, st)
    xn_patch = _new((G, w), dev)
    _lib.call("ttk_gather_rows", _ptr(tape.t["xn_fin"]), w, _ptr(dp.patch_row), _ptr(xn_patch), w, G, w, st)
    _wgrad(st, d_rows, xn_patch, grads["proj_out_w"])
    _lib.call("ttk_colsum", _ptr(d_rows), feat, G, feat, _ptr(grads["proj_out_b"]), _vp(0), st)
    dxn_patch = _new((G, w), dev)
    _gemm(st, d_rows, T["proj_out_w"], dxn_patch, w, feat, kn=1)
    dxn = torch.zeros((M, w), dtype=bf16, device=dev)
    _lib.call("ttk_scatter_rows", _ptr(dxn_patch), w, _ptr(dp.patch_row), _ptr(dxn), w, G, w, st)
    g = _new((M, w), dev)
    _rmsnorm_bwd(st, tape.t["x_fin"], T["ln_post"], dxn, g, grads["ln_post"])
    g0 = _layers_backward(m, W, dp, tape, g, grads)
    # embed: latent rows (dec_src_row >= 0) went through ln_pre_t, patch rows through ln_pre_p (blocks.py:164-167)
    d_e0 = _new((M, w), dev)
    _rmsnorm_bwd(st, tape.t["e0"], T["ln_pre_t"], g0, d_e0, grads["ln_pre_t"], sel=dp.dec_src_row, w2=T["ln_pre_p"],
                 dw2=grads["ln_pre_p"])
    _lib.call("ttk_colsum", _ptr(d_e0), w, M, w, _vp(0), _ptr(grads["mask_token"]), st)
    dcodes = torch.zeros((max(Tn, 1), m.token_size), dtype=torch.float32, device=dev)
    _lib.call("ttk_dec_in_bwd", _ptr(d_e0), w, _ptr(dp.latent_row), _ptr(tape.t["codes"]), m.token_size,
              _ptr(T["proj_in_w"]), _ptr(dcodes), _ptr(grads["proj_in_w"]), _ptr(grads["proj_in_b"]), Tn, w, st)
    return grads, dcodes[:Tn]


# --------------------------------------------------------------------------------------------------
# kernel-layout gradients -> the module's parameters (reference layout, blocks.py state-dict names)
# --------------------------------------------------------------------------------------------------
def _param_grads(m, kind: str, grads: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    perm = patch_feature_perm_on(m.patch_size_tuple, m.patch_channels, grads["mask_token"].device)
    out["mask_token"] = grads["mask_token"].view(1, 1)
    out["ln_pre_t.weight"] = grads["ln_pre_t"]
    out["ln_pre_p.weight"] = grads["ln_pre_p"]
    out["ln_post.weight"] = grads["ln_post"]
    # the two permuted tensors are un-permuted back INTO their slots of the stack's flat gradient buffer, so that every
    # parameter gradient of a stack is a view of ONE buffer (dist.GradientAllReducer all-reduces that buffer in place)
    if kind == "enc":
        gw = torch.empty_like(grads["proj_in_w"])
        gw[:, perm] = grads["proj_in_w"]  # kernel layout is weight[:, perm]
        grads["proj_in_w"].copy_(gw)
        out["proj_in.weight"] = grads["proj_in_w"]
        out["proj_in.bias"] = grads["proj_in_b"]
        out["proj_out.weight"] = grads["proj_out_w"]
        out["proj_out.bias"] = grads["proj_out_b"]
    else:
        out["proj_in.weight"] = grads["proj_in_w"]
        out["proj_in.bias"] = grads["proj_in_b"]
        gw = torch.empty_like(grads["proj_out_w"])
        gw[perm, :] = grads["proj_out_w"]
        gb = torch.empty_like(grads["proj_out_b"])
        gb[perm] = grads["proj_out_b"]
        grads["proj_out_w"].copy_(gw)
        grads["proj_out_b"].copy_(gb)
        out["proj_out.weight"] = grads["proj_out_w"]
        out["proj_out.bias"] = grads["proj_out_b"]
    for i in range(m.num_layers):
        out[f"model_layers.attn_layer.{i}.pre_ln.weight"] = grads[f"pre_ln{i}"]
        out[f"model_layers.attn_layer.{i}.to_qkv.weight"] = grads[f"to_qkv{i}"]
        out[f"model_layers.attn_layer.{i}.out_proj.weight"] = grads[f"out_proj{i}"]
        out[f"model_layers.ffd_layer.{i}.norm.weight"] = grads[f"ffn_norm{i}"]
        out[f"model_layers.ffd_layer.{i}.w12.weight"] = grads[f"w12_{i}"]
        out[f"model_layers.ffd_layer.{i}.w3.weight"] = grads[f"w3_{i}"]
        if i > 0:
            out[f"model_layers.attn_post_ln.{i - 1}.weight"] = grads[f"attn_post_ln{i}"]
            out[f"model_layers.ffd_post_ln.{i - 1}.weight"] = grads[f"ffd_post_ln{i}"]
    return out


def _ordered(m, kind: str, grads, params_meta) -> Tuple[Optional[torch.Tensor], ...]:
    pg = _param_grads(m, kind, grads)
    res = []
    for name, shape, dtype, req in params_meta:
        if not req:
            res.append(None)
            continue
        g = pg[name].reshape(shape)
        res.append(g if g.dtype == dtype else g.to(dtype))
    return tuple(res)


def _named_params(m):
    return cached_named_params(m)


def stack_params(m):
    return [p for _, p in _named_params(m)]


def _meta(m):
    return [(n, p.shape, p.dtype, p.requires_grad) for n, p in _named_params(m)]


class EncoderFn(torch.autograd.Function):
    """z = TiTokEncoder(clips) recorded for autograd. Inputs after `flat` are the module's parameters (graph edges only;
    the kernels read the prepared bf16 copies)."""

    @staticmethod
    def forward(ctx, m, dp, fsq_consts, flat, *params):
        z, codes, idx, tape = encoder_forward_train(m, dp, flat, fsq_consts)
        ctx.m, ctx.dp, ctx.tape, ctx.meta = m, dp, tape, _meta(m)
        ctx.need_input = flat.requires_grad
        ctx.mark_non_differentiable(idx, codes)
        return z, codes, idx

    @staticmethod
    def backward(ctx, dz, _dcodes, _didx):
        dz = dz.to(bf16).contiguous()
        grads, dflat = encoder_backward(ctx.m, ctx.dp, ctx.tape, dz, ctx.need_input)  # (the tape dies with the graph node)
        return (None, None, None, dflat) + _ordered(ctx.m, "enc", grads, ctx.meta)


class DecoderFn(torch.autograd.Function):
    """flat reconstruction = TiTokDecoder(codes) recorded for autograd."""

    @staticmethod
    def forward(ctx, m, dp, codes, *params):
        # always a private copy: `codes` may be a view of the device-wide 'codes' arena (engine.encoder_launch under
        # no_grad, the frozen-encoder case), which any later encoder launch overwrites before this node's backward runs
        codes_b = codes.detach().to(bf16).clone(memory_format=torch.contiguous_format)
        out, tape = decoder_forward_train(m, dp, codes_b)
        ctx.m, ctx.dp, ctx.tape, ctx.meta = m, dp, tape, _meta(m)
        ctx.codes_dtype = codes.dtype
        ctx.need_codes = codes.requires_grad
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = dout.to(bf16).contiguous()
        grads, dcodes = decoder_backward(ctx.m, ctx.dp, ctx.tape, dout)
        dc = dcodes.to(ctx.codes_dtype) if ctx.need_codes else None
        return (None, None, dc) + _ordered(ctx.m, "dec", grads, ctx.meta)


class SplitClips(torch.autograd.Function):
    """flat reconstruction buffer -> per-clip [C, T, H, W] views (engine.split_clips), with a backward that writes every
    clip's gradient straight into ONE flat buffer. (autograd's own backward of B slices materialises B full-size zero
    tensors and adds them up: O(B^2) traffic for a batch of B clips.)"""

    @staticmethod
    def forward(ctx, flat, plan):
        ctx.plan = plan
        ctx.flat_meta = (flat.shape, flat.dtype, flat.device)
        outs = []
        for off, n, g in zip(plan.clip_offset, plan.clip_numel, plan.grids_px):
            outs.append(flat[off:off + n].view(plan.channels, *g))
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        shape, dtype, device = ctx.flat_meta
        plan = ctx.plan
        if all(g is not None for g in grads) and len(grads) > 0:
            return torch.cat([g.reshape(-1).to(dtype) for g in grads]), None
        out = torch.zeros(shape, dtype=dtype, device=device)
        for g, off, n in zip(grads, plan.clip_offset, plan.clip_numel):
            if g is not None:
                out[off:off + n].copy_(g.reshape(-1))
        return out, None


def split_clips_autograd(flat: torch.Tensor, plan) -> List[torch.Tensor]:
    return list(SplitClips.apply(flat, plan))
