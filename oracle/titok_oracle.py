"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's tokenizer hot path (encode -> FSQ -> decode).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs may import this file;
the product (titok_video_b200) never does.

This is an independent, functional re-derivation of what the reference computes (no reference code is
imported or copied): plain torch on CPU, fp32 arithmetic, with a bf16 rounding at every point where the
reference's bf16 execution materialises a tensor. Each function cites the reference lines it restates.

Parity status: PINNED. tests/test_oracle_golden.py checks this oracle against fixtures under tests/golden/ that
were produced by running the UNMODIFIED reference (oracle/ref_shim.py, tests/golden/make_golden.py) in this
container, and -- when /root/reference is mounted -- against the live reference. The third-party arithmetic
the reference delegates (flash-attn RMSNorm / varlen attention, einops) is restated from those packages'
documented semantics at the versions this image ships (flash_attn 2.8.3, einops 0.8.2); the reference pins no
versions and has no tests of its own.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch

BF16 = torch.bfloat16
RMS_EPS = 1e-5

_SIZES = {"tiny": (4, (4, 2)), "small": (8, (8, 2)), "base": (12, (12, 4)), "large": (24, (16, 4))}


def model_dims(size: str) -> Tuple[int, int, Tuple[int, int]]:
    """(width, layers, (q_heads, kv_heads)); head_dim 64. [model/base/utils.py:8-23]"""
    layers, heads = _SIZES[size]
    return 64 * heads[0], layers, heads


def r(x: torch.Tensor) -> torch.Tensor:
    """Round to bf16 and return as fp32: the storage rounding of a bf16 tensor."""
    return x.to(BF16).to(torch.float32)


# ------------------------------------------------------------------------------------------------
# primitives
# ------------------------------------------------------------------------------------------------
def rmsnorm(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """flash-attn RMSNorm: fp32, rstd = 1/sqrt(mean(x^2)+1e-5), y = x*rstd*w, stored in the activation dtype.
    [flash_attn/ops/triton/layer_norm.py:1093-1126; used at blocks.py:51-52,66 and transformer.py:42,77,122-123]"""
    rstd = 1.0 / torch.sqrt(x.pow(2).mean(-1, keepdim=True) + RMS_EPS)
    return r(x * rstd * w.float())


def linear(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor = None) -> torch.Tensor:
    """nn.Linear in bf16: bf16 operands, fp32 accumulation, bias added before the output rounding."""
    y = x @ r(w.float()).t()
    if b is not None:
        y = y + r(b.float())
    return r(y)


def patchify(clip: torch.Tensor, patch: Sequence[int]) -> torch.Tensor:
    """'c (d0 p0) (d1 p1) (d2 p2) -> (d0 d1 d2) (p0 p1 p2 c)'  [model/base/utils.py:26-34]"""
    c, t, h, w = clip.shape
    p0, p1, p2 = patch
    x = clip.reshape(c, t // p0, p0, h // p1, p1, w // p2, p2)
    x = x.permute(1, 3, 5, 2, 4, 6, 0)
    return x.reshape((t // p0) * (h // p1) * (w // p2), p0 * p1 * p2 * c)


def unpatchify(rows: torch.Tensor, grid: Sequence[int], patch: Sequence[int], c: int) -> torch.Tensor:
    """'(d0 d1 d2) (p0 p1 p2 c) -> c (d0 p0) (d1 p1) (d2 p2)'  [model/base/utils.py:37-51]"""
    d0, d1, d2 = grid
    p0, p1, p2 = patch
    x = rows.reshape(d0, d1, d2, p0, p1, p2, c).permute(6, 0, 3, 1, 4, 2, 5)
    return x.reshape(c, d0 * p0, d1 * p1, d2 * p2)


def rope_cos_sin(grid: Sequence[int], token_count: int, theta: float = 10000.0, head_dim: int = 64):
    """fp32 cos/sin [s, 30] for one clip. ids: latent j -> (j,j,j); patch (a,b,c) -> (a,b,c)+token_count, last axis
    fastest; angle[l, f*3+axis] = inv_freq[f] * id[l, axis] in fp64; inv_freq = theta^linspace(0,1,10) * pi/2.
    [model/base/rope.py:40-45 (inv_freqs), :48-54 (_get_freqs_cis), :57-71 (forward)]"""
    nd = len(grid)
    nf = head_dim // (nd * 2)
    inv = torch.pow(theta, torch.linspace(0.0, 1.0, nf, dtype=torch.float64)) * torch.pi / 2.0
    tok = torch.arange(token_count, dtype=torch.float64).unsqueeze(-1).expand(-1, nd)
    coords = torch.cartesian_prod(*[torch.arange(g, dtype=torch.float64) for g in grid]).reshape(-1, nd)
    ids = torch.cat([tok, coords + float(token_count)], dim=0)
    ang = (inv.view(1, -1, 1) * ids.unsqueeze(-2)).reshape(ids.shape[0], -1)
    return torch.cos(ang).float(), torch.sin(ang).float()


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x [s, H, 64] (bf16-valued fp32): complex lanes (2i, 2i+1), i < 30 rotated in fp32, lanes 30,31 untouched,
    result rounded to bf16.  [model/base/rope.py:19-27]"""
    n = cos.shape[-1]
    xe, xo = x[..., 0:2 * n:2], x[..., 1:2 * n:2]
    c, s = cos.unsqueeze(1), sin.unsqueeze(1)
    out = x.clone()
    out[..., 0:2 * n:2] = xe * c - xo * s
    out[..., 1:2 * n:2] = xe * s + xo * c
    return r(out)


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, q_chunk: int = 2048) -> torch.Tensor:
    """One clip: q [s,Hq,64], k/v [s,Hkv,64]; exact softmax(q k^T / 8) v per head, q head h reads kv head
    h // (Hq/Hkv), output stored in bf16.  [transformer.py:100 -> flash_attn_varlen_func, non-causal, default scale]
    Query rows are independent, so long sequences are processed `q_chunk` rows at a time (bounds the [H, q, s] score
    matrix: 12 heads x 8448^2 fp32 would be 3.4 GB)."""
    hq, hk = q.shape[1], k.shape[1]
    kk = k.repeat_interleave(hq // hk, dim=1)
    vv = v.repeat_interleave(hq // hk, dim=1)
    out = []
    for a in range(0, q.shape[0], q_chunk):
        s = torch.einsum("qhd,khd->hqk", q[a:a + q_chunk], kk) * (1.0 / math.sqrt(q.shape[-1]))
        p = torch.softmax(s, dim=-1)
        out.append(torch.einsum("hqk,khd->qhd", p, vv))
    return r(torch.cat(out, dim=0))


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


# ------------------------------------------------------------------------------------------------
# transformer stack
# ------------------------------------------------------------------------------------------------
def attn_block(x, sd, pre, i, heads, cos, sin):
    """Attn.forward  [transformer.py:85-104]: pre_ln -> to_qkv -> split [q|gate|k|v] -> RoPE(q,k) -> attention ->
    * sigmoid(gate) -> out_proj."""
    w = x.shape[-1]
    hq, hk = heads
    g = hk * 64
    xn = rmsnorm(x, sd[f"{pre}.attn_layer.{i}.pre_ln.weight"])
    qkv = linear(xn, sd[f"{pre}.attn_layer.{i}.to_qkv.weight"])
    q, gate, k, v = qkv.split([w, w, g, g], dim=-1)
    q = apply_rope(q.reshape(-1, hq, 64), cos, sin)
    k = apply_rope(k.reshape(-1, hk, 64), cos, sin)
    o = attention(q, k, v.reshape(-1, hk, 64)).reshape(-1, w)
    o = r(o * r(torch.sigmoid(gate)))
    return linear(o, sd[f"{pre}.attn_layer.{i}.out_proj.weight"])


def geglu_block(x, sd, pre, i):
    """GEGLU.forward  [transformer.py:47-56]: norm -> w12 -> (value, gate) = chunk(2) -> gelu(gate)*value -> w3."""
    xn = rmsnorm(x, sd[f"{pre}.ffd_layer.{i}.norm.weight"])
    h = linear(xn, sd[f"{pre}.ffd_layer.{i}.w12.weight"])
    val, gate = h.chunk(2, dim=-1)
    h = r(r(gelu_erf(gate)) * val)
    return linear(h, sd[f"{pre}.ffd_layer.{i}.w3.weight"])


def layers(x, sd, pre, n_layers, heads, cos, sin, taps=None):
    """ResidualAttentionBlock.forward  [transformer.py:126-146]: layer 0 pre-LN residual; layers >= 1
    x = RMSNorm(alpha*x + f(x)) with alpha = 2*num_layer (KEEL)."""
    alpha = float(2 * n_layers)
    for i in range(n_layers):
        if i == 0:
            x = r(x + attn_block(x, sd, pre, i, heads, cos, sin))
            x = r(x + geglu_block(x, sd, pre, i))
        else:
            x = rmsnorm(r(r(alpha * x) + attn_block(x, sd, pre, i, heads, cos, sin)),
                        sd[f"{pre}.attn_post_ln.{i - 1}.weight"])
            x = rmsnorm(r(r(alpha * x) + geglu_block(x, sd, pre, i)), sd[f"{pre}.ffd_post_ln.{i - 1}.weight"])
        if taps is not None:
            taps.append(x.clone())
    return x


# ------------------------------------------------------------------------------------------------
# encoder / quantizer / decoder
# ------------------------------------------------------------------------------------------------
def encoder_forward(sd: Dict[str, torch.Tensor], size: str, patch: Sequence[int], clips: List[torch.Tensor],
                    token_counts: Sequence[int], prefix: str = "encoder", taps=None) -> torch.Tensor:
    """TiTokEncoder.forward  [blocks.py:71-104] -> z [sum(token_counts), token_size] (bf16-valued fp32).
    Clips never interact (attention is block-diagonal over cu_seqlens), so each is processed on its own."""
    width, n_layers, heads = model_dims(size)
    mt = r(sd[f"{prefix}.mask_token"].float().reshape(()))
    out = []
    for clip, tc in zip(clips, token_counts):
        tc = int(tc)
        grid = [s // p for s, p in zip(clip.shape[1:], patch)]
        cos, sin = rope_cos_sin(grid, tc)
        patches = linear(r(patchify(clip.float(), patch)), sd[f"{prefix}.proj_in.weight"], sd[f"{prefix}.proj_in.bias"])
        prow = rmsnorm(r(patches + mt), sd[f"{prefix}.ln_pre_p.weight"])
        lrow = rmsnorm(mt.expand(1, width).clone(), sd[f"{prefix}.ln_pre_t.weight"]).expand(tc, -1)
        x = torch.cat([lrow, prow], dim=0)
        x = layers(x, sd, f"{prefix}.model_layers", n_layers, heads, cos, sin, taps)
        tok = rmsnorm(x[:tc], sd[f"{prefix}.ln_post.weight"])
        out.append(linear(tok, sd[f"{prefix}.proj_out.weight"], sd[f"{prefix}.proj_out.bias"]))
    return torch.cat(out, dim=0)


def fsq_constants(levels: Sequence[int]):
    """[fsq.py:63-67,78-90]: basis = cumprod([1]+levels[:-1]); half_l = (L-1)*(1+1e-3)/2; offset = 0.5 if L even;
    shift = atanh(offset/half_l); half_width = L // 2.  (fp32, evaluated with torch like the reference)"""
    lv = torch.tensor(list(levels), dtype=torch.int32)
    basis = torch.cumprod(torch.tensor([1] + list(levels[:-1])), dim=0, dtype=torch.int32)
    half_l = (lv - 1) * (1 + 1e-3) / 2
    offset = torch.where(lv % 2 == 0, 0.5, 0.0)
    shift = (offset / half_l).atanh()
    return lv, basis, half_l, offset, shift, lv // 2


def fsq_forward(z: torch.Tensor, levels: Sequence[int]):
    """FSQ.forward  [fsq.py:123-135]: fp32; bound -> round-half-even -> /half_width -> mixed-radix index.
    Returns (codes fp32 (not yet cast to z's dtype), indices int32, bounded)."""
    lv, basis, half_l, offset, shift, hw = fsq_constants(levels)
    z = z.float()
    bounded = (z + shift).tanh() * half_l - offset
    q = bounded.round()
    codes = q / hw
    idx = ((codes * hw + hw) * basis).sum(dim=-1).to(torch.int32)
    return codes, idx, bounded


def fsq_indices_to_codes(idx: torch.Tensor, levels: Sequence[int]) -> torch.Tensor:
    """FSQ.indices_to_codes  [fsq.py:96-103,111-121]."""
    lv, basis, _, _, _, hw = fsq_constants(levels)
    lvl = (idx.unsqueeze(-1) // basis) % lv
    return (lvl - hw) / hw


def fsq_boundary_gap(bounded: torch.Tensor) -> torch.Tensor:
    """min over dims of | |b - rint(b)| - 0.5 |: distance of a token to the nearest rounding boundary."""
    return ((bounded - bounded.round()).abs() - 0.5).abs().min(dim=-1).values


def decoder_forward(sd: Dict[str, torch.Tensor], size: str, patch: Sequence[int], codes: torch.Tensor,
                    token_counts: Sequence[int], grids_px: Sequence[Sequence[int]], prefix: str = "decoder",
                    out_channels: int = 3, taps=None) -> List[torch.Tensor]:
    """TiTokDecoder.forward  [blocks.py:148-177] -> list of clips [3,T,H,W] (bf16-valued fp32)."""
    width, n_layers, heads = model_dims(size)
    mt = r(sd[f"{prefix}.mask_token"].float().reshape(()))
    out, t0 = [], 0
    for tc, gpx in zip(token_counts, grids_px):
        tc = int(tc)
        grid = [s // p for s, p in zip(gpx, patch)]
        g = math.prod(grid)
        cos, sin = rope_cos_sin(grid, tc)
        tok = linear(r(codes[t0:t0 + tc].float()), sd[f"{prefix}.proj_in.weight"], sd[f"{prefix}.proj_in.bias"])
        lrow = rmsnorm(r(tok + mt), sd[f"{prefix}.ln_pre_t.weight"])
        prow = rmsnorm(mt.expand(1, width).clone(), sd[f"{prefix}.ln_pre_p.weight"]).expand(g, -1)
        x = torch.cat([lrow, prow], dim=0)
        x = layers(x, sd, f"{prefix}.model_layers", n_layers, heads, cos, sin, taps)
        rows = rmsnorm(x[tc:], sd[f"{prefix}.ln_post.weight"])
        rows = linear(rows, sd[f"{prefix}.proj_out.weight"], sd[f"{prefix}.proj_out.bias"])
        out.append(unpatchify(rows, grid, patch, out_channels))
        t0 += tc
    return out


def titok_forward(sd: Dict[str, torch.Tensor], levels: Sequence[int], patch: Sequence[int], clips: List[torch.Tensor],
                  token_counts: Sequence[int], enc_size: str = "tiny", dec_size: str = "tiny"):
    """TiTok.forward  [model/titok.py:68-74]. Returns dict(z, bounded, codes, indices, recon)."""
    z = encoder_forward(sd, enc_size, patch, clips, token_counts)
    codes, idx, bounded = fsq_forward(z, levels)
    codes = r(codes)  # codes.to(orig_dtype), fsq.py:133
    recon = decoder_forward(sd, dec_size, patch, codes, token_counts, [c.shape[1:] for c in clips])
    return {"z": z, "bounded": bounded, "codes": codes, "indices": idx, "recon": recon}


def fsq_forward_ste(z: torch.Tensor, levels: Sequence[int]) -> torch.Tensor:
    """FSQ.quantize with the straight-through rounding of round_ste  [fsq.py:48-51,85-90]: differentiable codes."""
    lv, basis, half_l, offset, shift, hw = fsq_constants(levels)
    bounded = (z.float() + shift).tanh() * half_l - offset
    q = bounded + (bounded.round() - bounded).detach()
    return q / hw


def recon_l1_loss(clips: List[torch.Tensor], recon: List[torch.Tensor]) -> torch.Tensor:
    """Reconstruction term of ReconstructionLoss: mean over clips of the per-clip mean |x - recon|
    [model/losses/loss_module.py:118,160]."""
    return torch.stack([(r_.float() - c.float()).abs().mean() for c, r_ in zip(clips, recon)]).mean()


def titok_train_grads(sd: Dict[str, torch.Tensor], levels: Sequence[int], patch: Sequence[int],
                      clips: List[torch.Tensor], token_counts: Sequence[int], enc_size: str = "tiny",
                      dec_size: str = "tiny"):
    """One generator training step without the optimizer  [train.py:68-80: forward, L1 reconstruction loss, backward]:
    returns (loss, {state-dict name: fp32 gradient}). torch.autograd differentiates the restated forward; every r()
    rounds the gradient to bf16 on the way back exactly where bf16 autocast would hold a bf16 tensor."""
    leaves = {k: v.detach().clone().float().requires_grad_(True) for k, v in sd.items()}
    z = encoder_forward(leaves, enc_size, patch, clips, token_counts)
    codes = r(fsq_forward_ste(z, levels))
    recon = decoder_forward(leaves, dec_size, patch, codes, token_counts, [c.shape[1:] for c in clips])
    loss = recon_l1_loss(clips, recon)
    loss.backward()
    return float(loss), {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}


# ------------------------------------------------------------------------------------------------
# generic VQ oracle (north_star): torch.cdist(z, C).argmin(-1), chunked
# ------------------------------------------------------------------------------------------------
def vq_argmin(z: torch.Tensor, codebook: torch.Tensor, chunk: int = 8192):
    """indices = argmin_k ||z - c_k|| in fp32 (first minimum). Also returns the top-2 squared-distance gap used to
    classify near-ties."""
    zf, cf = z.float(), codebook.float()
    idx, gap = [], []
    for a in range(0, zf.shape[0], chunk):
        d = torch.cdist(zf[a:a + chunk], cf)
        idx.append(d.argmin(-1).to(torch.int32))
        if cf.shape[0] > 1:
            two = (d * d).topk(2, dim=-1, largest=False).values
            gap.append(two[:, 1] - two[:, 0])
        else:
            gap.append(torch.full((d.shape[0],), float("inf")))
    return torch.cat(idx), torch.cat(gap)


# ------------------------------------------------------------------------------------------------
# deterministic weights / inputs shared by tests, golden generation and the bench
# ------------------------------------------------------------------------------------------------
def codebook_scores(samples: Sequence[torch.Tensor], codebook_size: int) -> Tuple[float, float]:
    """(usage %, entropy in nats) of the reference CodebookLogger.get_scores
    (train_utils/codebook_logging.py:13-35): FIFO window of the last `codebook_size` SAMPLES (:13-17), sum of
    per-sample bincounts, nonzero fraction, scipy-style entropy of the normalised frequencies (sum p ln(1/p), p > 0)."""
    freq = torch.zeros(codebook_size, dtype=torch.float64)
    for s_ in list(samples)[-codebook_size:]:
        freq += torch.bincount(s_.reshape(-1).to(torch.int64), minlength=codebook_size)[:codebook_size].double()
    usage = float((freq > 0).sum()) / codebook_size * 100.0
    p = freq[freq > 0] / freq.sum()
    return usage, float(-(p * p.log()).sum())


def stress_init_(sd: Dict[str, torch.Tensor], seed: int = 1) -> Dict[str, torch.Tensor]:
    """the reference init (std 0.02) maps every latent to one code; the stress init draws every 2-D
    weight from N(0, (2/sqrt(fan_in))^2) so that tokens spread over the codebook. In place, deterministic."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd.keys()):
        v = sd[k]
        if v.dim() == 2 and v.shape[0] > 1 and v.shape[1] > 1:
            v.copy_(torch.randn(v.shape, generator=g) * (2.0 / math.sqrt(v.shape[1])))
        elif k.endswith(".bias"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
    return sd


def random_state_dict(seed: int = 42, enc: str = "tiny", dec: str = "tiny", patch: Sequence[int] = (4, 8, 8),
                      token_size: int = 5) -> Dict[str, torch.Tensor]:
    """A state dict with the reference's 76-key layout and initialiser STATISTICS (Linear ~ N(0, 0.02^2), bias 0, norm
    weights 1, mask_token ~ width^-0.5 N(0,1); utils.py:54-66, blocks.py:50,126) -- NOT the reference's RNG stream. Only for
    timing the port when no reference copy is at hand (bench.py fallback); parity tests draw weights through the modules."""
    g = torch.Generator().manual_seed(seed)
    pf = 3 * patch[0] * patch[1] * patch[2]
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, out_f, in_f, bias):
        sd[name + ".weight"] = torch.randn((out_f, in_f), generator=g) * 0.02
        if bias:
            sd[name + ".bias"] = torch.zeros(out_f)

    for prefix, size, fin, fout in (("encoder", enc, pf, token_size), ("decoder", dec, token_size, pf)):
        w, n_layers, (hq, hkv) = model_dims(size)
        inner = 32 * ((int(4.0 * (2 / 3) * w) + 31) // 32)  # transformer.py:39-40
        sd[f"{prefix}.mask_token"] = torch.randn((1, 1), generator=g) * w ** -0.5
        lin(f"{prefix}.proj_in", w, fin, True)
        sd[f"{prefix}.ln_pre_t.weight"] = torch.ones(w)
        sd[f"{prefix}.ln_pre_p.weight"] = torch.ones(w)
        for i in range(n_layers):
            ml = f"{prefix}.model_layers"
            sd[f"{ml}.attn_layer.{i}.pre_ln.weight"] = torch.ones(w)
            lin(f"{ml}.attn_layer.{i}.to_qkv", 2 * w + 2 * hkv * 64, w, False)
            lin(f"{ml}.attn_layer.{i}.out_proj", w, w, False)
            sd[f"{ml}.ffd_layer.{i}.norm.weight"] = torch.ones(w)
            lin(f"{ml}.ffd_layer.{i}.w12", 2 * inner, w, False)
            lin(f"{ml}.ffd_layer.{i}.w3", w, inner, False)
            if i < n_layers - 1:
                sd[f"{ml}.attn_post_ln.{i}.weight"] = torch.ones(w)
                sd[f"{ml}.ffd_post_ln.{i}.weight"] = torch.ones(w)
        sd[f"{prefix}.ln_post.weight"] = torch.ones(w)
        lin(f"{prefix}.proj_out", fout, w, True)
    return sd


def make_clips(shapes: Sequence[Sequence[int]], seed: int = 0) -> List[torch.Tensor]:
    """uniform [-1, 1) clips in bf16 (torch.manual_seed(0); rand*2-1)."""
    g = torch.Generator().manual_seed(seed)
    return [(torch.rand((3, *s), generator=g) * 2 - 1).to(BF16) for s in shapes]
