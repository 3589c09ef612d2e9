"""Per-kernel parity on the B200: every entry point of libtitok_b200.so is called through the C ABI and compared
with the CPU oracle (oracle/titok_oracle.py) on the same seeded inputs.

Tolerances. Integer / index outputs: bit-exact (FSQ, histogram, gathers), or exact modulo stated near-ties (VQ).
bf16 GEMM-class outputs: the products are exact in fp32, only the accumulation order differs from the oracle's,
so results agree to ~1 bf16 ulp: |d| <= 2^-7 * |ref| + small absolute term.
"""
import ctypes
import math

import numpy as np
import pytest
import torch

from conftest import from_bits, load_golden
from oracle import titok_oracle as O

pytestmark = pytest.mark.gpu

BF = torch.bfloat16
DEV = "cuda"


def lib():
    from titok_video_b200 import _lib

    return _lib


def P(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def ST():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_KEEP = []


def G(t):
    """Pointer to a CUDA copy of `t` that stays alive until the end of the test. (A temporary created inline in the
    argument list would be freed -- and its block reused by the next temporary -- before the asynchronous kernel
    reads it.)"""
    if t is None:
        return ctypes.c_void_p(0)
    d = t.to(DEV)
    _KEEP.append(d)
    return ctypes.c_void_p(d.data_ptr())


@pytest.fixture(autouse=True)
def _release_keepalive():
    yield
    torch.cuda.synchronize()
    _KEEP.clear()


def randn(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(BF)


def close_bf16(out, ref, ulps=2.0, atol=None, what=""):
    """out (cuda bf16) vs ref (cpu fp32, bf16-valued)."""
    o = out.float().cpu()
    scale = ref.abs().max().item() + 1e-20
    atol = (2.0 ** -8) * scale * 0.05 if atol is None else atol
    d = (o - ref).abs()
    # one bf16 ulp of ref is 2^(floor(log2|ref|) - 7): between 2^-8 |ref| (top of a binade) and 2^-7 |ref| (bottom)
    ulp = torch.exp2(torch.floor(torch.log2(ref.abs().clamp_min(1e-30))) - 7.0)
    tol = ulps * ulp + atol
    bad = (d > tol)
    assert torch.isfinite(o).all(), f"{what}: non-finite output"
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.numel()} outside {ulps} bf16 ulp; max|d|={d.max().item():.4g} "
                           f"scale={scale:.4g} first bad={torch.nonzero(bad)[:4].tolist()}")


# --------------------------------------------------------------------------------------------------
# GEMMs
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,bias", [(128, 128, 64, False), (300, 256, 256, True), (1000, 768, 256, True),
                                        (517, 256, 704, False), (256, 768, 768, True), (77, 40, 128, True),
                                        (5676, 256, 768, True),
                                        # >= 2 x 148 tiles with K <= 256: the weight-stationary kernel (ragged M and N)
                                        (12800 + 77, 768, 256, True), (20000, 600, 192, False)])
def test_gemm_store(M, N, K, bias):
    A, W = randn(M, K, seed=1), randn(N, K, seed=2, scale=0.1)
    b = randn(N, seed=3) if bias else None
    ref = O.linear(A.float(), W, b)
    out = torch.full((M, N), float("nan"), dtype=BF, device=DEV)
    Ad, Wd = A.to(DEV), W.to(DEV)
    bd = b.to(DEV) if bias else None
    lib().call("ttk_gemm_bf16", P(Ad), K, P(Wd), K, M, N, K, P(bd), P(out), N, P(None), 0, ST())
    torch.cuda.synchronize()
    close_bf16(out, ref, what=f"gemm {M}x{N}x{K}")


def test_gemm_store_row_map_and_ld():
    M, N, K = 200, 256, 128
    A, W = randn(M, K, seed=4), randn(N, K, seed=5, scale=0.1)
    ref = O.linear(A.float(), W)
    g = torch.Generator().manual_seed(0)
    rmap = torch.randperm(M + 50, generator=g)[:M].to(torch.int32)
    rmap[7] = -1
    out = torch.zeros((M + 50, N + 64), dtype=BF, device=DEV)
    lib().call("ttk_gemm_bf16", G(A), K, G(W), K, M, N, K, P(None), P(out), N + 64, G(rmap), 0,
               ST())
    torch.cuda.synchronize()
    o = out.cpu().float()
    keep = rmap >= 0
    close_bf16(out[rmap[keep].long().to(DEV), :N], ref[keep], what="row-mapped gemm")
    assert o[:, N:].abs().max() == 0
    untouched = torch.ones(M + 50, dtype=torch.bool)
    untouched[rmap[keep].long()] = False
    assert o[untouched].abs().max() == 0


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (300, 256, 256), (129, 64, 192)])
def test_gemm_weight_given_as_kn(M, N, K):
    """B operand with N contiguous (MN-major UMMA descriptor) -- the layout the attention kernel uses for V."""
    A, Wkn = randn(M, K, seed=6), randn(K, N, seed=7, scale=0.1)
    ref = O.r(A.float() @ Wkn.float())
    out = torch.empty((M, N), dtype=BF, device=DEV)
    lib().call("ttk_gemm_bf16", G(A), K, G(Wkn), N, M, N, K, P(None), P(out), N, P(None), 1, ST())
    torch.cuda.synchronize()
    close_bf16(out, ref, what="gemm kn")


@pytest.mark.parametrize("M,width,gqa", [(300, 256, 128), (1892, 256, 128), (200, 768, 256),
                                         (13000 + 5, 256, 128)])  # last: many tiles per CTA
def test_gemm_qkv_rope(M, width, gqa):
    K = width
    A, W = randn(M, K, seed=8), randn(2 * width + 2 * gqa, K, seed=9, scale=0.08)
    cos, sin = O.rope_cos_sin([3, 10, 10], M - 300 if M > 300 else 0)
    cos, sin = cos[:M], sin[:M]
    if cos.shape[0] < M:  # tile the table for the large case
        rep = (M + cos.shape[0] - 1) // cos.shape[0]
        cos, sin = cos.repeat(rep, 1)[:M], sin.repeat(rep, 1)[:M]
    rope = torch.stack([cos, sin], dim=-1).reshape(M, 60).contiguous()
    qkv = O.linear(A.float(), W)
    q, gate, k, v = qkv.split([width, width, gqa, gqa], dim=-1)
    q = O.apply_rope(q.reshape(M, -1, 64), cos, sin).reshape(M, -1)
    k = O.apply_rope(k.reshape(M, -1, 64), cos, sin).reshape(M, -1)
    ref = torch.cat([q, gate, k, v], dim=-1)
    out = torch.empty((M, 2 * width + 2 * gqa), dtype=BF, device=DEV)
    knorm = torch.full((gqa // 64, M), float("nan"), dtype=torch.float32, device=DEV)
    lib().call("ttk_gemm_qkv_rope", G(A), K, G(W), K, M, K, width, gqa, G(rope), P(out),
               out.stride(0), P(knorm), ST())
    torch.cuda.synchronize()
    # the rotation is done on bf16-rounded GEMM outputs: a 1-ulp difference of an input moves the output by ~1 ulp of it
    close_bf16(out, ref, ulps=3.0, atol=2.0 ** -8 * ref.abs().max().item() * 0.5, what="qkv+rope")
    # by-product: |k|^2 per row and kv head of the (pre-rotation, bf16) keys -- the rotation preserves the norm up to its
    # own rounding, which is why the attention kernel's bound carries a 2 % margin
    k_dev = out.float().cpu()[:, 2 * width:2 * width + gqa].reshape(M, gqa // 64, 64)
    n2 = (k_dev ** 2).sum(-1).t()
    got = knorm.cpu()
    assert torch.isfinite(got).all()
    assert ((got - n2).abs() <= 1.5e-2 * n2 + 1e-6).all(), "key norms"
    # without the by-product the call is unchanged
    out2 = torch.empty_like(out)
    lib().call("ttk_gemm_qkv_rope", G(A), K, G(W), K, M, K, width, gqa, G(rope), P(out2), out2.stride(0), P(None), ST())
    torch.cuda.synchronize()
    assert torch.equal(out2, out)


@pytest.mark.parametrize("M,inner,K", [(300, 704, 256), (1000, 1376, 512), (64, 704, 256),
                                       (6400 + 33, 704, 256)])  # last: weight-stationary kernel
def test_gemm_geglu(M, inner, K):
    A, W = randn(M, K, seed=10), randn(2 * inner, K, seed=11, scale=0.1)
    h = O.linear(A.float(), W)
    val, gate = h.chunk(2, dim=-1)
    ref = O.r(O.r(O.gelu_erf(gate)) * val)
    out = torch.empty((M, inner), dtype=BF, device=DEV)
    lib().call("ttk_gemm_geglu", G(A), K, G(W), K, M, inner, K, P(out), inner, ST())
    torch.cuda.synchronize()
    close_bf16(out, ref, ulps=4.0, atol=2.0 ** -8 * ref.abs().max().item() * 0.5, what="geglu")


def _resid_ref(x, y, mode, alpha, w_post, w_next):
    if mode == 0:
        xn = O.r(x + y)
    else:
        xn = O.rmsnorm(O.r(O.r(alpha * x) + y), w_post)
    return xn, O.rmsnorm(xn, w_next)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("M,K", [(300, 256), (1000, 704)])
def test_gemm_resid_norm256(mode, M, K):
    A, W = randn(M, K, seed=12), randn(256, K, seed=13, scale=0.1)
    x = randn(M, 256, seed=14)
    g = torch.Generator().manual_seed(15)
    w_post = 1 + 0.1 * torch.randn(256, generator=g)
    w_next = 1 + 0.1 * torch.randn(256, generator=g)
    y = O.linear(A.float(), W)
    ref_x, ref_xn = _resid_ref(x.float(), y, mode, 8.0, w_post, w_next)
    xo = torch.empty((M, 256), dtype=BF, device=DEV)
    xno = torch.empty((M, 256), dtype=BF, device=DEV)
    lib().call("ttk_gemm_resid_norm256", G(A), K, G(W), K, M, K, G(x), 256, mode, 8.0,
               G(w_post), G(w_next), P(xo), P(xno), 256, ST())
    torch.cuda.synchronize()
    close_bf16(xo, ref_x, ulps=3.0, atol=2.0 ** -8 * ref_x.abs().max().item() * 0.5, what="resid x")
    close_bf16(xno, ref_xn, ulps=4.0, atol=2.0 ** -8 * ref_xn.abs().max().item() * 0.5, what="resid xn")


# --------------------------------------------------------------------------------------------------
# attention
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seq_lens,hq,hkv", [([128], 4, 2), ([200, 64, 513], 4, 2), ([1892, 576], 4, 2),
                                               ([300, 129], 12, 4), ([257], 8, 2)])
@pytest.mark.parametrize("qk_scale", [1.0, 2.0, 3.5, "mixed"])
def test_attn_varlen(seq_lens, hq, hkv, qk_scale):
    """qk_scale 1.0: every tile takes the kernel's bounded-score loop (|q| max|k| / 8 * log2 e ~ 15 << 90); 2.0: the same
    loop near its limit (bound ~ 60..85: exponentials up to 2^80, peaked rows); 3.5: the bound
    is ~170, every tile takes the running-maximum loop (peaked rows: lazy rescales fire); "mixed": clips alternate, so
    both loops run in one launch."""
    from titok_video_b200.plan import attn_work_list

    width, gqa = hq * 64, hkv * 64
    M = sum(seq_lens)
    qkv = randn(M, 2 * width + 2 * gqa, seed=20, scale=1.0)
    starts = np.concatenate([[0], np.cumsum(seq_lens)[:-1]]).tolist()
    qk_cols = torch.ones(2 * width + 2 * gqa)
    qk_cols[:width] = 0
    qk_cols[2 * width:2 * width + gqa] = 0  # 0 where the column is q or k
    for ci, (s0, sl) in enumerate(zip(starts, seq_lens)):
        f = qk_scale if qk_scale != "mixed" else (3.5 if ci % 2 == 0 else 1.0)
        qkv[s0:s0 + sl] = (qkv[s0:s0 + sl].float() * (qk_cols + (1 - qk_cols) * f)).to(BF)
    work = torch.from_numpy(attn_work_list(starts, seq_lens, hq, hkv))
    q, gate, k, v = qkv.float().split([width, width, gqa, gqa], dim=-1)
    ref = torch.empty(M, width)
    for s0, sl in zip(starts, seq_lens):
        o = O.attention(q[s0:s0 + sl].reshape(sl, hq, 64), k[s0:s0 + sl].reshape(sl, hkv, 64),
                        v[s0:s0 + sl].reshape(sl, hkv, 64)).reshape(sl, width)
        ref[s0:s0 + sl] = O.r(o * O.r(torch.sigmoid(gate[s0:s0 + sl])))
    out = torch.full((M, width), float("nan"), dtype=BF, device=DEV)
    qd = qkv.to(DEV)
    # with the key norms that ttk_gemm_qkv_rope leaves behind (one launch) ...
    knorm = (k ** 2).reshape(M, hkv, 64).sum(-1).t().contiguous().to(DEV)
    lib().call("ttk_attn_varlen_fwd", P(qd), qd.stride(0), M, width, gqa, G(work), work.shape[0], 0.125, P(out),
               width, P(knorm), ST())
    # ... and without them (the library derives the bound from K with its own pre-kernel): same loop, same reference
    # up to the rounding of s c - ref
    out_nk = torch.full((M, width), float("nan"), dtype=BF, device=DEV)
    lib().call("ttk_attn_varlen_fwd", P(qd), qd.stride(0), M, width, gqa, G(work), work.shape[0], 0.125, P(out_nk),
               width, P(None), ST())
    torch.cuda.synchronize()
    o = out.float().cpu()
    assert (out_nk.float().cpu() - o).abs().max().item() <= 2.0 ** -7 * o.abs().max().item()
    assert torch.isfinite(o).all()
    d = (o - ref).abs()
    # P is rounded to bf16 before P.V (as in flash-attention): errors ~2^-9 relative to the value scale
    tol = 2e-2 * ref.abs() + 6e-3 * ref.abs().max()
    assert (d <= tol).all(), f"attention max|d|={d.max().item():.4g} scale={ref.abs().max().item():.4g} bad={int((d > tol).sum())}"
    assert d.mean().item() < 2e-3 * ref.abs().max().item()


# --------------------------------------------------------------------------------------------------
# row kernels
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("width", [256, 512, 768, 1024])
def test_rmsnorm_and_resid_norm(width):
    M = 333
    x, y = randn(M, width, seed=30), randn(M, width, seed=31)
    g = torch.Generator().manual_seed(32)
    w1 = 1 + 0.1 * torch.randn(width, generator=g)
    w2 = 1 + 0.1 * torch.randn(width, generator=g)
    out = torch.empty((M, width), dtype=BF, device=DEV)
    lib().call("ttk_rmsnorm_fwd", G(x), width, G(w1), P(out), width, M, width, ST())
    close_bf16(out, O.rmsnorm(x.float(), w1), ulps=1.0, what="rmsnorm")
    for mode in (0, 1):
        rx, rxn = _resid_ref(x.float(), y.float(), mode, 8.0, w1, w2)
        xo = torch.empty((M, width), dtype=BF, device=DEV)
        xno = torch.empty((M, width), dtype=BF, device=DEV)
        lib().call("ttk_resid_norm", G(x), G(y), P(xo), P(xno), G(w1), G(w2), 8.0, mode,
                   M, width, width, ST())
        close_bf16(xo, rx, ulps=1.5, what=f"resid_norm x mode {mode}")  # rstd may differ in its last bit
        close_bf16(xno, rxn, ulps=2.5, what=f"resid_norm xn mode {mode}")


def test_patchify_unpatchify_roundtrip_and_layout():
    from titok_video_b200.engine import patch_feature_perm
    from titok_video_b200.plan import make_plan

    shapes, tcs, patch = [(8, 32, 48), (4, 16, 24), (16, 168, 168)], [3, 0, 7], (4, 8, 8)
    clips = O.make_clips(shapes, 0)
    pl = make_plan(shapes, tcs, patch)
    flat = torch.cat([c.reshape(-1) for c in clips]).to(DEV)
    patches = torch.empty((pl.G, 768), dtype=BF, device=DEV)
    geom = torch.from_numpy(pl.geom).to(DEV)
    lib().call("ttk_patchify", P(flat), P(geom), 3, 4, 8, 8, P(patches), 768, pl.G, ST())
    ref = torch.cat([O.patchify(c.float(), patch) for c in clips], dim=0)  # reference feature order
    perm = patch_feature_perm(patch, 3)
    assert torch.equal(patches.float().cpu(), ref[:, perm]), "patchify is a pure permutation: must be bit-exact"
    # scatter back through a row map (rows = packed rows)
    rows = torch.zeros((pl.M, 768), dtype=BF, device=DEV)
    prow = torch.from_numpy(pl.patch_row).to(DEV)
    rows[prow.long()] = patches
    out = torch.zeros_like(flat)
    lib().call("ttk_unpatchify", P(rows), 768, P(prow), P(geom), 3, 4, 8, 8, P(out), pl.G, ST())
    assert torch.equal(out, flat)


def _fsq_consts(levels, device="cpu"):
    import titok_video_b200 as T

    return T.FSQ(list(levels))._consts(torch.device(device))


def test_enc_dec_embed_and_head():
    from titok_video_b200.plan import make_plan

    width, ts = 256, 5
    shapes, tcs = [(8, 32, 32), (4, 16, 24)], [5, 2]
    pl = make_plan(shapes, tcs, (4, 8, 8))
    g = torch.Generator().manual_seed(40)
    proj = randn(pl.G, width, seed=41)
    mt = torch.tensor([0.37])
    w_t, w_p, w_n = [1 + 0.1 * torch.randn(width, generator=g) for _ in range(3)]
    mtb = O.r(mt)
    # encoder embed
    ref = torch.empty(pl.M, width)
    lat = O.rmsnorm(mtb.expand(1, width).clone(), w_t)
    ref[torch.from_numpy(pl.latent_row).long()] = lat
    ref[torch.from_numpy(pl.patch_row).long()] = O.rmsnorm(O.r(proj.float() + mtb), w_p)
    xo = torch.empty((pl.M, width), dtype=BF, device=DEV)
    xno = torch.empty((pl.M, width), dtype=BF, device=DEV)
    lib().call("ttk_enc_embed", G(proj), width, G(torch.from_numpy(pl.enc_src_row)), G(mt),
               G(w_t), G(w_p), G(w_n), P(xo), P(xno), pl.M, width, width, ST())
    close_bf16(xo, ref, ulps=1.0, what="enc_embed x")
    close_bf16(xno, O.rmsnorm(ref, w_n), ulps=1.5, what="enc_embed xn")
    # decoder embed
    codes = (torch.randint(-2, 3, (pl.T, ts), generator=g).float() / 2).to(BF)
    w_in, b_in = randn(width, ts, seed=42, scale=0.3), randn(width, seed=43, scale=0.1)
    ref = torch.empty(pl.M, width)
    ref[torch.from_numpy(pl.latent_row).long()] = O.rmsnorm(O.r(O.linear(codes.float(), w_in, b_in) + mtb), w_t)
    ref[torch.from_numpy(pl.patch_row).long()] = O.rmsnorm(mtb.expand(1, width).clone(), w_p)
    lib().call("ttk_dec_embed", G(codes), ts, G(torch.from_numpy(pl.dec_src_row)), G(w_in),
               G(b_in), G(mt), G(w_t), G(w_p), G(w_n), P(xo), P(xno), pl.M,
               width, width, ST())
    close_bf16(xo, ref, ulps=1.5, what="dec_embed x")
    # encoder head + FSQ
    levels = [7, 5, 5, 5, 5]
    x = randn(pl.M, width, seed=44, scale=2.0)
    w_out, b_out = randn(ts, width, seed=45, scale=0.15), randn(ts, seed=46, scale=0.1)
    tok = O.rmsnorm(x.float()[torch.from_numpy(pl.latent_row).long()], w_n)
    z_ref = O.linear(tok, w_out, b_out)
    z = torch.empty((pl.T, ts), dtype=BF, device=DEV)
    cd = torch.empty((pl.T, ts), dtype=BF, device=DEV)
    idx = torch.empty((pl.T,), dtype=torch.int32, device=DEV)
    lib().call("ttk_enc_head_fsq", G(x), width, G(torch.from_numpy(pl.latent_row)), G(w_n), 0,
               G(w_out), G(b_out), ts, P(z), P(cd), P(idx), pl.T, width, *_fsq_consts(levels), ST())
    close_bf16(z, z_ref, ulps=2.0, what="enc head z")
    codes_o, idx_o, _ = O.fsq_forward(z.float().cpu(), levels)  # FSQ of the kernel's own z must be exact
    assert torch.equal(idx.cpu(), idx_o)
    assert torch.equal(cd.float().cpu(), O.r(codes_o))


# --------------------------------------------------------------------------------------------------
# quantizer
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_fsq_against_reference_vectors(tag):
    """Bit-exact against vectors produced by the reference's FSQ (tests/golden/fsq_kat.npz)."""
    import titok_video_b200 as T

    kat = load_golden("fsq_kat")
    levels = kat[f"{tag}_levels"].tolist()
    q = T.FSQ(levels).to(DEV)
    assert q.codebook_size == int(kat[f"{tag}_codebook_size"])
    assert q._basis.cpu().tolist() == kat[f"{tag}_basis"].tolist()
    assert torch.equal(q.implicit_codebook.cpu(), torch.from_numpy(kat[f"{tag}_implicit_codebook"]))
    z = torch.from_numpy(kat[f"{tag}_z"])
    codes, d = q(z.to(DEV))
    ref_idx = torch.from_numpy(kat[f"{tag}_indices"])
    ref_codes = torch.from_numpy(kat[f"{tag}_codes"])
    # stated epsilon: an index may differ only where bound(z) is within eps_b of a rounding boundary
    # (tanh on the GPU vs the CPU that generated the vectors can differ by an ulp)
    _, _, bounded = O.fsq_forward(z, levels)
    gap = O.fsq_boundary_gap(bounded)
    hl = max(levels) * 0.5
    eps_b = 4 * 2.0 ** -24 * hl
    neq = d["indices"].cpu() != ref_idx
    assert not (neq & (gap > eps_b)).any(), f"{int(neq.sum())} index mismatches away from boundaries"
    assert torch.equal(codes.cpu()[~neq], ref_codes[~neq])
    assert int(neq.sum()) <= 2
    # bf16 input path
    zb = z.to(BF)
    codes_b, db = q(zb.to(DEV))
    neq_b = db["indices"].cpu() != torch.from_numpy(kat[f"{tag}_indices_bf16"])
    _, _, bounded_b = O.fsq_forward(zb.float(), levels)
    assert not (neq_b & (O.fsq_boundary_gap(bounded_b) > eps_b)).any()
    assert torch.equal(codes_b.cpu()[~neq_b], from_bits(kat[f"{tag}_codes_bf16_bits"])[~neq_b])
    # indices -> codes, int32 and int64
    i2c = torch.from_numpy(kat[f"{tag}_i2c"])
    assert torch.equal(q.indices_to_codes(ref_idx.to(DEV)).cpu(), i2c)
    assert torch.equal(q.indices_to_codes(ref_idx.to(DEV).long()).cpu(), i2c)
    # round trip: indices_to_codes(forward(z).indices) == forward(z).codes
    assert torch.equal(q.indices_to_codes(d["indices"]).cpu(), codes.cpu())


def test_fsq_large_and_ragged_sizes():
    import titok_video_b200 as T

    q = T.FSQ([7, 5, 5, 5, 5]).to(DEV)
    for n in (0, 1, 1023, 1025, 2 ** 20 + 3):
        g = torch.Generator().manual_seed(n)
        z = (torch.randn((n, 5), generator=g) * 2).to(BF)
        codes, d = q(z.to(DEV))
        assert codes.shape == (n, 5) and d["indices"].shape == (n,)
        if n:
            c_ref, i_ref, bounded = O.fsq_forward(z.float(), [7, 5, 5, 5, 5])
            neq = d["indices"].cpu() != i_ref
            assert not (neq & (O.fsq_boundary_gap(bounded) > 1e-6)).any()
            assert int(neq.sum()) <= max(2, n // 100000)
            # properties at full size: indices_to_codes inverts the index map exactly (fsq.py:111-121), and the
            # codes lie on the level grid (code * half_width is an integer in [-half_width, half_width]).
            # (FSQ.forward is NOT idempotent on its own codes: tanh(1) * 3.003 rounds to 2, not 3, for a 7-level dim.)
            assert torch.equal(q.indices_to_codes(d["indices"]).to(codes.dtype), codes)
            hw = torch.tensor([3, 2, 2, 2, 2], dtype=torch.float32, device=DEV)
            lv = codes.float() * hw
            # codes are bf16 here (2/3 * 3 = 2.0039): on the grid up to one bf16 rounding of level / half_width
            assert bool(((lv - lv.round()).abs() <= 2.0 ** -7).all()) and bool((lv.round().abs() <= hw).all())


def test_fsq_backward_ste():
    import titok_video_b200 as T

    levels = [8, 8, 8, 6, 5]
    q = T.FSQ(levels).to(DEV)
    g = torch.Generator().manual_seed(1)
    z = (torch.randn((4096, 5), generator=g) * 1.5)
    zc = z.to(DEV).requires_grad_(True)
    codes, _ = q(zc)
    w = torch.randn((4096, 5), generator=g)
    (codes * w.to(DEV)).sum().backward()
    zr = z.clone().requires_grad_(True)
    lv, basis, half_l, offset, shift, hw = O.fsq_constants(levels)
    b = (zr + shift).tanh() * half_l - offset
    cr = (b + (b.round() - b).detach()) / hw
    (cr * w).sum().backward()
    assert torch.allclose(zc.grad.cpu(), zr.grad, rtol=1e-5, atol=1e-6)


def test_histogram_and_stats():
    from scipy.stats import entropy

    for K, n in [(4375, 300000), (16, 1000), (65536, 200000), (15360, 7)]:
        g = torch.Generator().manual_seed(K)
        idx = torch.randint(0, K, (n,), generator=g, dtype=torch.int32)
        idx[::3] = idx[0]
        cnt = torch.zeros(K, dtype=torch.int32, device=DEV)
        lib().call("ttk_hist_u32", G(idx), n, K, P(cnt), ST())
        ref = torch.bincount(idx.long(), minlength=K)
        assert torch.equal(cnt.cpu().long(), ref)
        out = torch.empty(3, dtype=torch.float64, device=DEV)
        lib().call("ttk_codebook_stats", P(cnt), K, P(out), ST())
        nz, ent, tot = out.cpu().tolist()
        f = ref.double().numpy()
        assert nz == float((ref > 0).sum()) and tot == float(n)
        assert abs(ent - entropy(f / f.sum())) < 1e-9


def test_codebook_logger_matches_reference_vector():
    import titok_video_b200 as T

    kat = load_golden("fsq_kat")
    lens = kat["logger_lens"].tolist()
    flat = torch.from_numpy(kat["logger_samples"])
    samples = list(torch.split(flat, lens))
    for dev in ("cpu", DEV):
        lg = T.CodebookLogger(16)
        lg([s.to(dev) for s in samples])
        sc = lg.get_scores()
        assert abs(float(sc["codebook/usage_percent"]) - float(kat["logger_usage"])) < 1e-4
        assert abs(float(sc["codebook/entropy"]) - float(kat["logger_entropy"])) < 1e-5
        assert lg.codebook_indices == []


# --------------------------------------------------------------------------------------------------
# generic VQ: tensor-core distance + fused argmin
# --------------------------------------------------------------------------------------------------
def _vq_run(z, cb):
    N, D = z.shape
    K = cb.shape[0]
    DA = lib().fn("ttk_vq_aug_dim")(D)
    D8 = (D + 7) // 8 * 8
    zp = torch.zeros((N, D8), dtype=BF, device=DEV)
    zp[:, :D] = z.to(DEV)
    cbd = cb.to(DEV).contiguous()
    aug = torch.empty((lib().fn("ttk_vq_aug_rows")(K, D), DA), dtype=BF, device=DEV)
    lib().call("ttk_vq_prepare_codebook", P(cbd), D, K, D, P(aug), DA, ST())
    idx = torch.full((N,), -1, dtype=torch.int32, device=DEV)
    best = torch.empty((N,), dtype=torch.float32, device=DEV)
    lib().call("ttk_vq_argmin", P(zp), D8, P(aug), DA, N, K, D, P(idx), P(best), ST())
    torch.cuda.synchronize()
    return idx.cpu(), best.cpu()


@pytest.mark.parametrize("N,K,D", [(1000, 256, 16), (4096, 1024, 64), (3000, 4096, 128), (2048, 1000, 5),
                                   (5000, 4375, 5), (1500, 777, 256), (300, 16384, 32),
                                   # > 148 row tiles and D <= 189: the two-tiles-per-CTA kernel (odd tile count, ragged K)
                                   (19072 + 77, 1000, 64), (40000, 4375, 5), (25000, 777, 128), (20000, 300, 189)])
def test_vq_argmin(N, K, D):
    g = torch.Generator().manual_seed(N + K + D)
    z = (torch.randn((N, D), generator=g) * 2).to(BF)
    if (K, D) == (4375, 5):
        import titok_video_b200 as T

        cb = T.FSQ([7, 5, 5, 5, 5]).implicit_codebook.to(BF)
    else:
        cb = torch.randn((K, D), generator=g).to(BF)
    idx, best = _vq_run(z, cb)
    ref_idx, gap = O.vq_argmin(z, cb)
    assert (idx >= 0).all() and (idx < K).all()
    neq = idx != ref_idx
    # stated near-tie criterion: top-2 squared-distance gap below 2^-18 * (|z|^2 + max|c|^2)
    scale = z.float().pow(2).sum(-1) + cb.float().pow(2).sum(-1).max()
    assert not (neq & (gap > 2.0 ** -18 * scale)).any(), f"{int(neq.sum())} mismatches, some with a clear gap"
    assert neq.float().mean().item() < 1e-3
    # the reported score is |c|^2 - 2 z.c of the chosen code
    c = cb.float()[idx.long()]
    want = c.pow(2).sum(-1) - 2 * (z.float() * c).sum(-1)
    assert torch.allclose(best, want, rtol=1e-4, atol=1e-3)


def test_vq_matches_fsq_indices():
    """cdist/argmin over FSQ.implicit_codebook reproduces FSQ's indices (product code). Checked against the
    ORACLE on both sides: O.vq_argmin over the implicit codebook == O.fsq_forward's indices == the CUDA argmin."""
    import titok_video_b200 as T

    levels = [7, 5, 5, 5, 5]
    q = T.FSQ(levels).to(DEV)
    g = torch.Generator().manual_seed(11)
    z = (torch.randn((20000, 5), generator=g) * 2).to(BF)
    codes_o, idx_o, _ = O.fsq_forward(z.float(), levels)
    cb = q.implicit_codebook.cpu().to(BF)
    codes_b = O.r(codes_o).to(BF)  # the codes as the bf16 path materialises them
    ref_idx, gap = O.vq_argmin(codes_b, cb)
    # codes are exact grid points up to bf16 rounding of k/3, k/2: the nearest codeword is the code itself
    assert (ref_idx == idx_o).float().mean().item() > 0.999
    idx, _ = _vq_run(codes_b, cb)
    neq = idx != ref_idx
    scale = codes_b.float().pow(2).sum(-1) + cb.float().pow(2).sum(-1).max()
    assert not (neq & (gap > 2.0 ** -18 * scale)).any()
    # and the drop-in FSQ kernel agrees with both
    _, d = q(z.to(DEV))
    assert torch.equal(d["indices"].cpu(), idx_o)
    assert (idx == d["indices"].cpu()).float().mean().item() > 0.999


@pytest.mark.parametrize("K,D", [(1024, 64), (4096, 128), (16384, 128), (65536, 128), (4096, 256)])
def test_vq_argmin_quantizer_microbench_shapes(K, D):
    """BASELINE configs[1] / SURVEY C2 at full size: N = 2^20 latent vectors, codebooks 1K..64K. The chunked fp32
    cdist/argmin oracle runs on a strided sample of 4096 rows (the full 2^20 x 64K fp32 distance matrix is 262 GB); the
    size-independent property over ALL rows is that the reported score of the winner equals |c|^2 - 2 z.c of that code
    and is <= the score of a random competitor."""
    N = 1 << 20
    g = torch.Generator().manual_seed(K + D)
    z = (torch.randn((N, D), generator=g) * 2).to(BF)
    cb = torch.randn((K, D), generator=g).to(BF)
    idx, best = _vq_run(z, cb)
    assert (idx >= 0).all() and (idx < K).all()
    sel = torch.arange(0, N, N // 4096)[:4096]
    ref_idx, gap = O.vq_argmin(z[sel], cb, chunk=512)
    neq = idx[sel] != ref_idx
    scale = z[sel].float().pow(2).sum(-1) + cb.float().pow(2).sum(-1).max()
    assert not (neq & (gap > 2.0 ** -18 * scale)).any(), f"{int(neq.sum())} mismatches, some with a clear gap"
    assert neq.float().mean().item() < 2e-3
    # all rows: score of the winner, and no random competitor beats it (computed on the GPU in chunks)
    zd, cd, id_, bd = z.to(DEV).float(), cb.to(DEV).float(), idx.to(DEV).long(), best.to(DEV)
    c2 = cd.pow(2).sum(-1)
    rnd = torch.randint(0, K, (N,), generator=g).to(DEV)
    for a in range(0, N, 1 << 17):
        sl = slice(a, a + (1 << 17))
        want = c2[id_[sl]] - 2 * (zd[sl] * cd[id_[sl]]).sum(-1)
        assert torch.allclose(bd[sl], want, rtol=1e-4, atol=2e-3 * (1 + D / 64))
        other = c2[rnd[sl]] - 2 * (zd[sl] * cd[rnd[sl]]).sum(-1)
        assert (want <= other + 2.0 ** -16 * (zd[sl].pow(2).sum(-1) + c2.max())).all()


def test_vq_fp32_latents_near_tie_criterion():
    """north_star: indices bit-exact against the fp32 cdist/argmin path except where the top-2 distance gap is below a
    stated fp32 epsilon. fp32 latents / codebooks are rounded to bf16 for the tensor-core GEMM (what autocast does to a
    Linear), which moves every squared distance by at most 2^-8 (|z|^2 + 2|c|^2); two candidates can therefore swap
    only when their fp32 gap is below  eps = 2^-7 * (|z|^2 + 2 max|c|^2)."""
    from titok_video_b200.model.quantizer.vq import VectorQuantizer

    N, K, D = 30000, 2048, 96
    g = torch.Generator().manual_seed(3)
    z = torch.randn((N, D), generator=g) * 1.5           # fp32, NOT bf16-representable
    q = VectorQuantizer(K, D).to(DEV)
    with torch.no_grad():
        q.codebook.weight.copy_(torch.randn((K, D), generator=g).to(DEV))
    cb = q.codebook.weight.detach().cpu()
    with torch.no_grad():
        zq, d = q(z.to(DEV))
    idx = d["indices"].cpu()
    ref_idx, gap = O.vq_argmin(z, cb)                     # fp32 oracle on the fp32 inputs
    neq = idx != ref_idx
    eps = 2.0 ** -7 * (z.pow(2).sum(-1) + 2 * cb.pow(2).sum(-1).max())
    assert not (neq & (gap > eps)).any(), f"{int((neq & (gap > eps)).sum())} flips with a clear fp32 gap"
    assert 0 < neq.float().mean().item() < 0.05           # the criterion is exercised, and rarely
    # against the oracle on the bf16-rounded operands the tight criterion of test_vq_argmin holds
    ref_b, gap_b = O.vq_argmin(z.to(BF), cb.to(BF))
    nb = idx != ref_b
    scale = z.pow(2).sum(-1) + cb.pow(2).sum(-1).max()
    assert not (nb & (gap_b > 2.0 ** -18 * scale)).any()
    # quantized latents are the bf16 codebook rows of the chosen indices, returned in the input dtype
    assert zq.dtype == torch.float32 and torch.equal(zq.cpu(), cb.to(BF)[idx.long()].float())


def test_vector_quantizer_module_losses_and_straight_through_backward():
    """VectorQuantizer (north_star's quantizer contract): indices, quantized latents, commitment / codebook losses and
    the straight-through backward, against the standard VQ-VAE formulation in plain torch fp32 autograd on the same
    bf16-rounded operands. Tolerances: losses rel 1e-4; dz within bf16 rounding (2^-8 rel + 1e-6); dC rel 1e-4 (fp32
    atomics: summation order)."""
    from titok_video_b200.model.quantizer.vq import VectorQuantizer

    N, K, D = 4000, 300, 40
    g = torch.Generator().manual_seed(8)
    q = VectorQuantizer(K, D, commitment_weight=0.25, codebook_weight=1.0).to(DEV)
    with torch.no_grad():
        q.codebook.weight.copy_(torch.randn((K, D), generator=g).to(BF).float().to(DEV))
    z0 = torch.randn((N, D), generator=g).to(BF)
    w_out = torch.randn((N, D), generator=g).to(BF).float()
    # ours
    z = z0.to(DEV).requires_grad_(True)
    zq, d = q(z.view(40, 100, D))
    assert zq.shape == (40, 100, D) and d["indices"].shape == (40, 100) and d["indices"].dtype == torch.int32
    total = (zq.float().view(N, D) * w_out.to(DEV)).sum() + 3.0 * d["loss"] + 0.5 * d["commitment_loss"]
    total.backward()
    torch.cuda.synchronize()
    # oracle
    zr = z0.float().requires_grad_(True)
    C = q.codebook.weight.detach().cpu().clone().requires_grad_(True)
    idx_ref, gap = O.vq_argmin(z0, C.detach().to(BF))
    idx = d["indices"].view(-1).cpu()
    assert (idx == idx_ref).float().mean() > 0.999
    c = C[idx.long()]
    commit = ((zr - c.detach()) ** 2).mean()
    code = ((zr.detach() - c) ** 2).mean()
    zq_ref = zr + (c - zr).detach()
    total_ref = (zq_ref * w_out).sum() + 3.0 * (0.25 * commit + 1.0 * code) + 0.5 * commit
    total_ref.backward()
    assert torch.equal(zq.detach().float().view(N, D).cpu(), c.detach().to(BF).float())
    assert abs(float(d["commitment_loss"]) - float(commit)) < 1e-4 * float(commit)
    assert abs(float(d["codebook_loss"]) - float(code)) < 1e-4 * float(code)
    assert abs(float(d["loss"]) - float(0.25 * commit + code)) < 1e-4 * float(code)
    dz, dz_ref = z.grad.float().cpu(), zr.grad
    assert (dz - dz_ref).abs().max() <= 2.0 ** -8 * dz_ref.abs().max() + 1e-6
    dC, dC_ref = q.codebook.weight.grad.cpu(), C.grad
    assert (dC - dC_ref).norm() <= 1e-4 * dC_ref.norm()
    # unused codes get exactly zero gradient
    unused = torch.ones(K, dtype=torch.bool)
    unused[idx.long()] = False
    assert unused.any() or K <= N
    assert (dC[unused] == 0).all()
    # no_grad path: same values, no graph; indices_to_codes inverts the lookup
    with torch.no_grad():
        zq2, d2 = q(z0.to(DEV))
    assert torch.equal(zq2, zq.detach().view(N, D)) and torch.equal(d2["indices"], d["indices"].view(-1))
    assert torch.equal(q.indices_to_codes(d2["indices"]), zq2)
    # an optimizer step changes the codebook: the prepared operand is rebuilt
    with torch.no_grad():
        q.codebook.weight.add_(1.0)
        _, d3 = q(z0.to(DEV))
    ref3, _ = O.vq_argmin(z0, q.codebook.weight.detach().cpu().to(BF))
    assert (d3["indices"].cpu() == ref3).float().mean() > 0.999


def test_vq_gather_and_loss():
    N, K, D = 5000, 512, 64
    g = torch.Generator().manual_seed(5)
    z = torch.randn((N, D), generator=g).to(BF)
    cb = torch.randn((K, D), generator=g).to(BF)
    idx = torch.randint(0, K, (N,), generator=g, dtype=torch.int32)
    zq = torch.empty((N, D), dtype=BF, device=DEV)
    loss = torch.zeros(1, dtype=torch.float32, device=DEV)
    lib().call("ttk_vq_gather_loss", G(z), D, G(cb), D, G(idx), N, D, P(zq), D, P(loss), ST())
    ref = cb[idx.long()]
    assert torch.equal(zq.cpu(), ref)
    want = (ref.float() - z.float()).pow(2).sum().item()
    assert abs(loss.item() - want) / want < 1e-4


def test_bad_arguments_return_errors_not_crashes():
    L = lib()
    x = torch.zeros((8, 256), dtype=BF, device=DEV)
    w = torch.ones(256, device=DEV)
    assert L.fn("ttk_rmsnorm_fwd")(P(None), 256, P(w), P(x), 256, 8, 256, ST()) == -1
    assert L.fn("ttk_rmsnorm_fwd")(P(x), 256, P(w), P(x), 256, 8, 200, ST()) == -2
    assert L.fn("ttk_gemm_bf16")(P(x), 255, P(x), 256, 8, 8, 256, P(None), P(x), 256, P(None), 0, ST()) == -3
    with pytest.raises(L.TitokB200Error):
        L.call("ttk_patchify", P(x), P(x), 3, 4, 8, 4, P(x), 768, 1, ST())


def test_device_built_plan_equals_host_planner():
    """engine.DevicePlan expands the packing metadata on the device (ttk_build_plan) from per-clip descriptors and gathers
    the [M,60] RoPE table (ttk_rope_table_gather); every array must equal the host planner's (plan.make_plan, which
    tests/test_host_logic.py and tests/test_oracle_golden.py pin to the reference's packing and RoPE.forward)."""
    from titok_video_b200 import engine
    from titok_video_b200.plan import make_plan

    shapes, tcs = [(16, 168, 168), (8, 128, 136), (4, 16, 24), (12, 160, 152)], [128, 1, 0, 77]
    engine.clear_caches()
    dp = engine.get_device_plan(shapes, tcs, (4, 8, 8), 3, torch.device(DEV))
    torch.cuda.synchronize()
    hp = make_plan(shapes, tcs, (4, 8, 8), 3)
    assert (dp.plan.M, dp.plan.T, dp.plan.G) == (hp.M, hp.T, hp.G)
    for name in ("enc_src_row", "dec_src_row", "latent_row", "patch_row", "geom", "rope_pos"):
        want = torch.from_numpy(getattr(hp, name))
        got = getattr(dp, name).cpu()[:want.shape[0]]
        assert torch.equal(got, want), name
    assert torch.equal(dp.rope.cpu()[:hp.M], torch.from_numpy(hp.rope))
    assert torch.equal(dp.clip_offset.cpu(), torch.tensor(hp.clip_offset)) and torch.equal(dp.clip_numel.cpu(), torch.tensor(hp.clip_numel))


def test_normalize_u8_is_bit_identical_to_the_reference_expression():
    """dataset/video_dataset.py:118-119 on bf16 tensors: `chunk.to(bf16) / 255` then `chunk * 2 - 1`."""
    g = torch.Generator().manual_seed(0)
    for n in (256, 16 * 1000 + 7, 3 * 8 * 64 * 48):
        src = torch.randint(0, 256, (n,), generator=g, dtype=torch.uint8)
        src[:256] = torch.arange(256, dtype=torch.uint8)  # every possible value
        ref = (src.to(BF) / 255) * 2 - 1
        sd = src.to(DEV)
        out = torch.empty(n, dtype=BF, device=DEV)
        lib().call("ttk_normalize_u8", P(sd), P(out), n, ST())
        torch.cuda.synchronize()
        assert torch.equal(out.cpu().view(torch.int16), ref.view(torch.int16))
