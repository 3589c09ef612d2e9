"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference from /root/reference on CPU.

The reference cannot be imported as-is without a GPU: `flash_attn.ops.triton.layer_norm` queries the CUDA driver
at import time and `flash_attn_varlen_func` is CUDA/half only; `xformers` (imported by model/base/transformer.py:6,
used only by dead code) is not installed. This module injects three stand-ins into sys.modules *before* importing
the reference's own files, which then run unchanged:

  * flash_attn.ops.triton.layer_norm.RMSNorm -> pure-torch module: fp32 math, eps 1e-5, weight only, output in
    x.dtype (the arithmetic of flash-attn 2.8.3's rms_norm_ref / _layer_norm_fwd_1pass_kernel)
  * flash_attn.flash_attn_varlen_func        -> per-segment F.scaled_dot_product_attention, kv heads repeated
    for GQA (q head h uses kv head h // (Hq/Hkv)), softmax scale d^-0.5, non-causal
  * xformers.ops.SwiGLU                      -> placeholder class (never instantiated)

It exists to (a) generate the golden fixtures under tests/golden/ (make_golden.py), (b) pin the oracle
restatement (oracle/titok_oracle.py) against the real reference, (c) run the reference itself as the baseline arm of
bench.py (`--impl reference` on the host cores, `config.gpu_reference` on the GPU) and as the live GPU oracle of
tests/test_gpu_reference.py.

Where the reference comes from: $TITOK_REFERENCE_ROOT, else /root/reference (the build container), else the git-ignored
copy under <repo>/baseline/_ref/ that scripts/vendor_reference.sh makes at build time and that travels to the GPU box
with the gpurun snapshot.

Two modes (one per process, sys.modules is global):
  install("cpu")  the three stand-ins above (no CUDA needed)
  install("gpu")  ONLY the xformers placeholder: the reference then runs its REAL third-party kernels -- flash-attn
                  2.8.3 `flash_attn_varlen_func` (transformer.py:100), flash-attn's Triton RMSNorm (blocks.py:27) and
                  cuBLAS through nn.Linear -- i.e. the same-box bar of SURVEY 2.1 K1-K3.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VENDORED_ROOT = os.path.join(_REPO, "baseline", "_ref")


def _find_root() -> str:
    for cand in (os.environ.get("TITOK_REFERENCE_ROOT"), "/root/reference", VENDORED_ROOT):
        if cand and os.path.isdir(os.path.join(cand, "model")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model"))


class _RMSNorm(nn.Module):
    def __init__(self, hidden_size, eps=1e-5, dropout_p=0.0, zero_centered_weight=False, device=None, dtype=None):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(hidden_size, device=device, dtype=dtype))
        self.register_parameter("bias", None)

    def forward(self, x, residual=None, prenorm=False, residual_in_fp32=False):
        xf = x.float()
        rstd = 1.0 / torch.sqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)
        return (xf * rstd * self.weight.float()).to(x.dtype)


def _flash_attn_varlen_func(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, dropout_p=0.0,
                            softmax_scale=None, causal=False, **kw):
    assert not causal and dropout_p == 0.0
    hq, hk = q.shape[1], k.shape[1]
    out = torch.empty_like(q)
    cu = cu_seqlens_q.tolist()
    for a, b in zip(cu[:-1], cu[1:]):
        qs = q[a:b].transpose(0, 1)  # [H, L, D]
        ks = k[a:b].transpose(0, 1).repeat_interleave(hq // hk, dim=0)
        vs = v[a:b].transpose(0, 1).repeat_interleave(hq // hk, dim=0)
        o = F.scaled_dot_product_attention(qs.unsqueeze(0), ks.unsqueeze(0), vs.unsqueeze(0), scale=softmax_scale)
        out[a:b] = o.squeeze(0).transpose(0, 1)
    return out


_installed = None  # "cpu" | "gpu"


def install(mode: str = "cpu"):
    global _installed
    if _installed == mode:
        return
    if _installed is not None:
        raise RuntimeError(f"reference already imported in {_installed!r} mode in this process")
    if mode not in ("cpu", "gpu"):
        raise ValueError(mode)
    if not available():
        raise RuntimeError(f"reference not found (looked at $TITOK_REFERENCE_ROOT, /root/reference, {VENDORED_ROOT})")
    mods = []
    if mode == "cpu":
        fa = types.ModuleType("flash_attn")
        fa.flash_attn_varlen_func = _flash_attn_varlen_func
        fa_ops = types.ModuleType("flash_attn.ops")
        fa_tr = types.ModuleType("flash_attn.ops.triton")
        fa_ln = types.ModuleType("flash_attn.ops.triton.layer_norm")
        fa_ln.RMSNorm = _RMSNorm
        mods += [("flash_attn", fa), ("flash_attn.ops", fa_ops), ("flash_attn.ops.triton", fa_tr),
                 ("flash_attn.ops.triton.layer_norm", fa_ln)]
    try:
        import xformers.ops  # noqa: F401  (not installed in this image; if it ever is, use it)
        if not hasattr(sys.modules["xformers.ops"], "SwiGLU"):
            raise ImportError
    except Exception:
        xf = types.ModuleType("xformers")
        xf_ops = types.ModuleType("xformers.ops")
        xf_ops.SwiGLU = type("SwiGLU", (), {})
        xf.ops = xf_ops
        mods += [("xformers", xf), ("xformers.ops", xf_ops)]
    for name, mod in mods:
        sys.modules[name] = mod
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = mode


def reference_modules(mode: str = "cpu"):
    """(model.titok, model.base.blocks, model.quantizer.fsq, train_utils.codebook_logging) of the reference."""
    install(mode)
    return (importlib.import_module("model.titok"), importlib.import_module("model.base.blocks"),
            importlib.import_module("model.quantizer.fsq"), importlib.import_module("train_utils.codebook_logging"))


class AttrDict(dict):
    __getattr__ = dict.__getitem__

    @staticmethod
    def wrap(o):
        if isinstance(o, dict):
            return AttrDict({k: AttrDict.wrap(v) for k, v in o.items()})
        return o


def build_reference_titok(fsq_levels=(7, 5, 5, 5, 5), patch_size=(4, 8, 8), encoder_size="tiny", decoder_size="tiny",
                          seed=42, mode="cpu"):
    titok_mod, _, _, _ = reference_modules(mode)
    cfg = AttrDict.wrap({"tokenizer": {"model": {"patch_size": list(patch_size), "fsq_levels": list(fsq_levels),
                                                  "encoder_size": encoder_size, "decoder_size": decoder_size}}})
    torch.manual_seed(seed)
    return titok_mod.TiTok(cfg)
