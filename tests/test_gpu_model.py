"""Model-level parity on the B200: titok_video_b200.TiTok (CUDA kernels through the C ABI) against
(a) the golden fixtures produced by the unmodified reference and (b) the CPU oracle, on identical inputs and
weights, at the default init and at the stress init.

Stated tolerances (bf16 execution, 8 transformer layers):
  default init : elementwise |d| <= 2e-2 + 2e-2*|ref|                     (observed ~1e-3)
  stress init  : relative Frobenius error <= 3e-2 and max|d| <= 4e-2*max|ref|  -- the reference's own CPU bf16
                 run differs from the fp32-accumulating oracle by 1.5e-2 / 2.0e-2 on these inputs
  indices      : bit-exact wherever the oracle's bound(z) is farther from a rounding boundary than
                 half_l * |dz| (dz = observed |z_cuda - z_ref| for that token, floor 1e-3); FSQ on identical z is
                 bit-exact (tests/test_gpu_kernels.py)
"""
import numpy as np
import pytest
import torch

from conftest import build_model, checksum_matches, from_bits, param_checksum
from oracle import titok_oracle as O

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
LEVELS = [7, 5, 5, 5, 5]
PATCH = [4, 8, 8]


def _case(golden):
    shapes = [tuple(s) for s in golden["shapes"].tolist()]
    tcs = golden["token_counts"].tolist()
    clips = O.make_clips(shapes, int(golden["clip_seed"]))
    return shapes, tcs, clips


def _rel_fro(a, b):
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def _check(a, b, stress, what):
    a, b = a.float().cpu(), b.float().cpu()
    assert torch.isfinite(a).all(), what
    d = (a - b).abs()
    if stress:
        assert _rel_fro(a, b) <= 3e-2, f"{what}: rel fro {_rel_fro(a, b):.4g}"
        assert d.max() <= 4e-2 * b.abs().max(), f"{what}: max|d| {d.max():.4g} vs scale {b.abs().max():.4g}"
    else:
        bad = d > 2e-2 + 2e-2 * b.abs()
        assert not bad.any(), f"{what}: {int(bad.sum())} elements outside tolerance, max|d| {d.max():.4g}"


def _check_indices(idx, ref_idx, z, z_ref, what):
    idx, ref_idx = idx.cpu(), ref_idx.cpu()
    _, _, bounded = O.fsq_forward(z_ref.float().cpu(), LEVELS)
    gap = O.fsq_boundary_gap(bounded)
    dz = (z.float().cpu() - z_ref.float().cpu()).abs().max(dim=-1).values
    allowed = gap <= 3.5 * torch.clamp(dz, min=1e-3)  # half_l <= 3.003; d bound/dz <= half_l
    neq = idx != ref_idx
    assert not (neq & ~allowed).any(), f"{what}: index flips away from rounding boundaries: {torch.nonzero(neq & ~allowed).flatten().tolist()}"
    return int(neq.sum())


@pytest.mark.parametrize("stress", [False, True])
def test_forward_matches_reference_fixture_and_oracle(stress, golden_default, golden_stress):
    golden = golden_stress if stress else golden_default
    shapes, tcs, clips = _case(golden)
    model = build_model(stress)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    assert checksum_matches(sd, golden["weight_checksum"]), \
        "weights regenerated from the seed differ from the ones the reference produced"
    model = model.cuda().eval()
    cl = [c.cuda() for c in clips]
    with torch.no_grad():
        z = model.encoder(cl, tcs)
        x_q, d = model.encode(cl, tcs)
        recon, d2 = model(cl, tcs)
    z_ref = from_bits(golden["z_bits"])
    idx_ref = torch.from_numpy(golden["indices"])
    assert z.shape == z_ref.shape and z.dtype == BF and d["indices"].dtype == torch.int32
    assert torch.equal(d["indices"], d2["indices"])
    # (a) encoder output vs the reference's own output
    _check(z, z_ref, stress, "z vs reference fixture")
    flips = _check_indices(d["indices"], idx_ref, z, z_ref, "indices vs reference fixture")
    # (b) vs the oracle
    res = O.titok_forward(sd, LEVELS, PATCH, clips, tcs)
    _check(z, res["z"], stress, "z vs oracle")
    _check_indices(d["indices"], res["indices"], z, res["z"], "indices vs oracle")
    # codes are the FSQ of our own z, exactly
    c_own, i_own, _ = O.fsq_forward(z.float().cpu(), LEVELS)
    assert torch.equal(d["indices"].cpu(), i_own) and torch.equal(x_q.float().cpu(), O.r(c_own))
    # (c) decoder on the REFERENCE's codes (decouples index flips from decoder error)
    codes_ref = from_bits(golden["codes_bits"])
    with torch.no_grad():
        rec = model.decode(codes_ref.cuda(), tcs, shapes)
    for i, rr in enumerate(rec):
        assert rr.shape == (3, *shapes[i])
        _check(rr, from_bits(golden[f"recon{i}_bits"]).view(3, *shapes[i]), stress, f"recon{i} vs reference fixture")
    rec_o = O.decoder_forward(sd, "tiny", PATCH, codes_ref.float(), tcs, shapes)
    for i, rr in enumerate(rec):
        _check(rr, rec_o[i], stress, f"recon{i} vs oracle")
    # (d) end to end: clips whose tokens all match the reference reconstruct within tolerance
    if flips == 0:
        for i, rr in enumerate(recon):
            _check(rr, from_bits(golden[f"recon{i}_bits"]).view(3, *shapes[i]), stress, f"e2e recon{i}")


def test_decode_indices_equals_forward_and_accepts_both_forms(golden_stress):
    shapes, tcs, clips = _case(golden_stress)
    model = build_model(True).cuda().eval()
    cl = [c.cuda() for c in clips]
    with torch.no_grad():
        recon, d = model(cl, torch.tensor(tcs, dtype=torch.int32))
        a = model.decode_indices(d["indices"], shapes, tcs)
        b = model.decode_indices(list(torch.split(d["indices"], tcs)), torch.tensor(shapes, dtype=torch.int32))
        _, ds = model.encode(cl, tcs, split_indices=True)
    for r0, r1, r2 in zip(recon, a, b):
        assert torch.equal(r0, r1) and torch.equal(r0, r2)
    assert isinstance(ds["indices"], tuple) and [t.shape[0] for t in ds["indices"]] == tcs


def test_batch_composition_does_not_change_results(golden_stress):
    """Clips never interact (block-diagonal attention): a clip tokenises identically alone or packed."""
    shapes, tcs, clips = _case(golden_stress)
    model = build_model(True).cuda().eval()
    cl = [c.cuda() for c in clips]
    with torch.no_grad():
        rec_all, d_all = model(cl, tcs)
        off = 0
        for i in range(len(cl)):
            rec_i, d_i = model([cl[i]], [tcs[i]])
            assert torch.equal(d_i["indices"], d_all["indices"][off:off + tcs[i]])
            assert torch.equal(rec_i[0], rec_all[i])
            off += tcs[i]
        # order permutation
        rec_p, d_p = model(cl[::-1], tcs[::-1])
        assert torch.equal(rec_p[0], rec_all[-1]) and torch.equal(rec_p[-1], rec_all[0])


def test_canonical_clip_A_against_oracle():
    """C1 of SURVEY 8d: clip A = 3x16x168x168 with 128 latent tokens plus clip B = 3x8x128x128 with 64."""
    shapes, tcs = [(16, 168, 168), (8, 128, 128)], [128, 64]
    clips = O.make_clips(shapes, 0)
    for stress in (False, True):
        model = build_model(stress)
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        model = model.cuda().eval()
        with torch.no_grad():
            z = model.encoder([c.cuda() for c in clips], tcs)
            x_q, d = model.encode([c.cuda() for c in clips], tcs)
        z_o = O.encoder_forward(sd, "tiny", PATCH, clips, tcs)
        _check(z, z_o, stress, f"clip A/B z (stress={stress})")
        _, idx_o, _ = O.fsq_forward(z_o, LEVELS)
        _check_indices(d["indices"], idx_o, z, z_o, "clip A/B indices")
        with torch.no_grad():
            rec = model.decode(x_q, tcs, shapes)
        rec_o = O.decoder_forward(sd, "tiny", PATCH, x_q.float().cpu(), tcs, shapes)
        for i in range(2):
            _check(rec[i], rec_o[i], stress, f"clip {'AB'[i]} recon (stress={stress})")


def test_fused_and_unfused_residual_paths_agree(golden_stress, monkeypatch):
    from titok_video_b200 import engine

    shapes, tcs, clips = _case(golden_stress)
    model = build_model(True).cuda().eval()
    cl = [c.cuda() for c in clips]
    with torch.no_grad():
        monkeypatch.setattr(engine, "FUSE_RESID_256", True)
        z1 = model.encoder(cl, tcs).clone()
        monkeypatch.setattr(engine, "FUSE_RESID_256", False)
        engine.clear_caches()
        z2 = model.encoder(cl, tcs).clone()
    _check(z1, z2, True, "fused vs unfused residual epilogue")


@pytest.mark.parametrize("size,native", [("tiny", True), ("tiny", False), ("small", True), ("base", False)])
def test_latent_tail_of_the_last_encoder_layer_is_bit_identical(size, native, monkeypatch):
    """The encoder's head reads the latent rows only (blocks.py:101), so the last layer carries nothing else past its
    attention (engine._layer_latent / ttk_layer_fwd_latent: attention for the query tiles that hold latent rows, the rest
    of the layer on the gathered latent rows). Every step is row-wise, so z, the codes and the indices must equal the
    all-rows path (TTK_LATENT_TAIL=0, what the reference computes) BIT FOR BIT -- fused (width 256) and unfused residual
    kernels, even and odd query-head groups, through the native sequencer and kernel by kernel; clips with 0 tokens, with
    latent rows spilling into a second query tile, and with exactly one tile of them."""
    from titok_video_b200 import engine

    shapes, tcs = [(8, 64, 64), (4, 32, 48), (8, 96, 64), (4, 16, 16)], [200, 0, 37, 128]
    clips = [c.cuda() for c in O.make_clips(shapes, 17)]
    model = build_model(True, enc=size, dec=size).cuda().eval()
    monkeypatch.setattr(engine, "NATIVE_SEQ", native)
    out = {}
    with torch.no_grad():
        for tail in (False, True):
            monkeypatch.setattr(engine, "LATENT_TAIL", tail)
            engine.clear_caches()
            _poison_workspace()  # (the all-rows run must not leave the right values behind for the tail run to find)
            z = model.encoder(clips, tcs).clone()
            _poison_workspace()
            rec, d = model(clips, tcs)
            out[tail] = (z, d["indices"].clone(), [r.clone() for r in rec])
    z0, i0, r0 = out[False]
    z1, i1, r1 = out[True]
    assert z0.shape == (sum(tcs), 5) and torch.isfinite(z0.float()).all()
    assert torch.equal(z0, z1), f"z differs: max|d| {(z0.float() - z1.float()).abs().max().item():.4g}"
    assert torch.equal(i0, i1), f"{int((i0 != i1).sum())} indices differ"
    assert all(torch.equal(a, b) for a, b in zip(r0, r1))
    assert len(torch.unique(i0)) > 8  # (the stress initialiser spreads the tokens: the comparison is not vacuous)


def _poison_workspace():
    """Every byte of the device-wide workspace arenas becomes 0xFF: NaN as bf16 / fp32, -1 as int32. A launch sequence that
    reads anything it did not write itself (stale results of an earlier call used to mask exactly that) shows up as NaN."""
    from titok_video_b200 import engine

    torch.cuda.synchronize()
    for t in engine._ARENA.values():
        t.fill_(0xFF)
    torch.cuda.synchronize()


@pytest.mark.parametrize("tail", [True, False])
def test_results_do_not_depend_on_what_the_workspace_held(tail, monkeypatch):
    """The launch sequences share device-wide workspace arenas, so a call always finds the leftovers of earlier calls there
    -- often the right values of the very same batch, which would hide a kernel that fails to write (or reads past) its
    rows. With the arenas poisoned (NaN everywhere) before each call, the eager launches, the per-composition graph replay
    and the bucketed graph replay must still return the same tokens and reconstructions, bit for bit. Covers the latent
    tail of the last encoder layer (rows of `att` it leaves unwritten are never read) and the padded rows of a bucket (the
    attention kernel's last key box of the last clip reads into them: they must hold finite values)."""
    from titok_video_b200 import engine

    monkeypatch.setattr(engine, "LATENT_TAIL", tail)
    engine.clear_caches()
    model = build_model(True).cuda().eval()
    comps = [([(16, 168, 168), (8, 96, 64), (16, 168, 168)], [128, 37, 64]),   # multi-tile clips, s = 1892 / 229 / 1828
             ([(8, 64, 64), (4, 16, 16), (8, 64, 48)], [200, 128, 16])]
    for shapes, tcs in comps:
        clips = [c.cuda() for c in O.make_clips(shapes, 23)]
        with torch.no_grad():
            rec, d = model.tokenize_reconstruct_(clips, tcs, use_graph=False)
            ref_idx, ref_rec = d["indices"].clone(), [r.clone() for r in rec]
            assert len(torch.unique(ref_idx)) > 8 and all(torch.isfinite(r.float()).all() for r in ref_rec)
            for what, run in [("eager", lambda: model.tokenize_reconstruct_(clips, tcs, use_graph=False)),
                              ("graph", lambda: model.tokenize_reconstruct_(clips, tcs, use_graph=True)),
                              ("graph replay", lambda: model.tokenize_reconstruct_(clips, tcs, use_graph=True)),
                              ("bucketed", lambda: model.tokenize_reconstruct_bucketed_(clips, tcs)),
                              ("bucketed replay", lambda: model.tokenize_reconstruct_bucketed_(clips, tcs))]:
                _poison_workspace()
                rec, d = run()
                bad = int((d["indices"] != ref_idx).sum())
                assert bad == 0, f"{what}: {bad} of {ref_idx.numel()} indices differ after poisoning the workspace"
                for a, b in zip(rec, ref_rec):
                    assert torch.equal(a, b), f"{what}: reconstruction differs after poisoning the workspace"


def test_other_model_sizes_run_and_match_oracle():
    """'small' (width 512, 8 layers, heads 8/2) exercises the generic-width path (unfused residual kernels)."""
    shapes, tcs = [(4, 32, 32), (8, 16, 24)], [4, 9]
    clips = O.make_clips(shapes, 3)
    model = build_model(False, enc="small", dec="small")
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda().eval()
    with torch.no_grad():
        z = model.encoder([c.cuda() for c in clips], tcs)
        x_q, _ = model.encode([c.cuda() for c in clips], tcs)
        rec = model.decode(x_q, tcs, shapes)
    _check(z, O.encoder_forward(sd, "small", PATCH, clips, tcs), False, "small z")
    rec_o = O.decoder_forward(sd, "small", PATCH, x_q.float().cpu(), tcs, shapes)
    for a, b in zip(rec, rec_o):
        _check(a, b, False, "small recon")


def test_discriminator_style_encoder():
    """loss_module.py:43-48 builds TiTokEncoder(out_channels=1) with 4 register tokens per clip."""
    import titok_video_b200 as T

    torch.manual_seed(0)
    enc = T.TiTokEncoder("tiny", (4, 8, 8), in_channels=3, out_channels=1).cuda()
    clips = [c.cuda() for c in O.make_clips([(8, 32, 32), (4, 16, 16)], 1)]
    with torch.no_grad():
        out = enc(clips, torch.tensor([4, 4], dtype=torch.int32))
    assert out.shape == (8, 1) and torch.isfinite(out.float()).all()
    assert out.view(2, -1).shape == (2, 4)


def test_weight_update_is_picked_up():
    model = build_model(True).cuda().eval()
    clips = [c.cuda() for c in O.make_clips([(4, 16, 16)], 2)]
    with torch.no_grad():
        z0 = model.encoder(clips, [4]).clone()
        model.encoder.proj_out.bias.add_(1.0)
        z1 = model.encoder(clips, [4]).clone()
    assert torch.allclose((z1 - z0).float(), torch.ones_like(z0.float()), atol=0.1)


def test_state_dict_roundtrip_with_reference_layout(golden_default):
    model = build_model(False)
    sd = model.state_dict()
    assert len(sd) == 76 and sum(v.numel() for v in sd.values()) == 6828295
    m2 = build_model(True)
    m2.load_state_dict(sd, strict=True)
    assert checksum_matches(m2.state_dict(), golden_default["weight_checksum"])


def test_cuda_graph_replay_equals_eager_launches(golden_stress):
    """The throughput path replays one captured CUDA graph per shape signature; results must be bit-identical to the
    kernel-by-kernel launches, for new inputs and after an in-place weight update (the graph must not go stale)."""
    shapes, tcs, clips = _case(golden_stress)
    model = build_model(True).cuda().eval()
    cl = [c.cuda() for c in clips]
    with torch.no_grad():
        ref, dref = model(cl, tcs)
        ref = [r.clone() for r in ref]
        for _ in range(2):  # first call captures, second replays
            rec, d = model.tokenize_reconstruct_(cl, tcs, use_graph=True)
            assert torch.equal(d["indices"], dref["indices"])
            assert all(torch.equal(a, b) for a, b in zip(rec, ref))
        cl2 = [torch.flip(c, dims=(-1,)).contiguous() for c in cl]
        ref2, dref2 = model(cl2, tcs)
        ref2 = [r.clone() for r in ref2]
        rec2, d2 = model.tokenize_reconstruct_(cl2, tcs, use_graph=True)
        assert torch.equal(d2["indices"], dref2["indices"]) and all(torch.equal(a, b) for a, b in zip(rec2, ref2))
        model.decoder.proj_out.bias.add_(0.25)
        ref3, _ = model(cl2, tcs)
        ref3 = [r.clone() for r in ref3]
        rec3, _ = model.tokenize_reconstruct_(cl2, tcs, use_graph=True)
        assert all(torch.equal(a, b) for a, b in zip(rec3, ref3)) and not torch.equal(ref3[0], ref2[0])


def test_ragged_batch_from_config_ranges_and_long_clip():
    """Packed ragged batch drawn from the sampling ranges of configs/tiny.yaml (tiny.yaml:56-66: T 8..16, H/W 128..168,
    1..128 latent tokens, multiples of the patch size) plus one long clip (8x256x256, 256 tokens, s = 2304) of the
    scaled configuration (BASELINE configs[4]); includes token_count == 1 and sequence lengths that are not multiples
    of the 128-row attention tile."""
    import random

    rnd = random.Random(0)
    shapes = [(rnd.choice([8, 12, 16]), rnd.choice([128, 136, 152, 168]), rnd.choice([128, 144, 160, 168])) for _ in range(4)]
    tcs = [1, rnd.randint(2, 128), rnd.randint(2, 128), 128]
    shapes.append((8, 256, 256))
    tcs.append(256)
    clips = O.make_clips(shapes, 5)
    model = build_model(True)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda().eval()
    with torch.no_grad():
        z = model.encoder([c.cuda() for c in clips], tcs)
        x_q, d = model.encode([c.cuda() for c in clips], tcs)
        rec = model.decode(x_q, tcs, shapes)
    z_o = O.encoder_forward(sd, "tiny", PATCH, clips, tcs)
    _check(z, z_o, True, "ragged z")
    _, idx_o, _ = O.fsq_forward(z_o, LEVELS)
    _check_indices(d["indices"], idx_o, z, z_o, "ragged indices")
    rec_o = O.decoder_forward(sd, "tiny", PATCH, x_q.float().cpu(), tcs, shapes)
    for i, (a, b) in enumerate(zip(rec, rec_o)):
        assert tuple(a.shape) == (3, *shapes[i])
        _check(a, b, True, f"ragged recon clip {i}")


def test_base_size_odd_gqa_ratio():
    """'base' (width 768, 12 layers, heads 12/4: three query heads per kv head) takes the other attention work-list
    branch (two consecutive row tiles of one head share a K/V stream, last tile of a head possibly alone)."""
    shapes, tcs = [(8, 48, 40), (4, 16, 16)], [7, 2]  # s = 67 and 6 -> one 128-row tile each; odd tile counts
    clips = O.make_clips(shapes, 9)
    model = build_model(False, enc="base", dec="base")
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda().eval()
    with torch.no_grad():
        z = model.encoder([c.cuda() for c in clips], tcs)
    _check(z, O.encoder_forward(sd, "base", PATCH, clips, tcs), False, "base z")


def test_clip_error_kernel_matches_torch(golden_stress):
    """ttk_clip_error: per-clip sum |x - recon| and sum (x - recon)^2 in fp64 (loss_module.py:118 numerator, PSNR)."""
    shapes, tcs, clips = _case(golden_stress)
    model = build_model(True).cuda().eval()
    cl = [c.cuda() for c in clips]
    for use_graph in (False, True, True):
        with torch.no_grad():
            rec, d = model.tokenize_reconstruct_(cl, tcs, use_graph=use_graph, with_error=True)
        err = d["clip_error"].cpu()
        for i, (a, b) in enumerate(zip(cl, rec)):
            diff = a.double().cpu() - b.double().cpu()
            assert abs(err[i, 0].item() - diff.abs().sum().item()) <= 1e-4 * diff.abs().sum().item() + 1e-6
            assert abs(err[i, 1].item() - (diff * diff).sum().item()) <= 1e-4 * (diff * diff).sum().item() + 1e-6


def test_torch_compile_wrapper_is_transparent():
    """train.py:38-39 / loss_module.py:50-51 may wrap the modules in torch.compile: the drop-in modules opt out of Dynamo
    tracing (their kernels are opaque ctypes calls) and produce the same tokens and gradients as the unwrapped module."""
    model = build_model(True).cuda()
    clips = [c.cuda() for c in O.make_clips([(8, 32, 32), (4, 16, 24)], 0)]
    tcs = torch.tensor([8, 3], dtype=torch.int32)
    with torch.no_grad():
        rec, d = model(clips, tcs)
    cm = torch.compile(model)
    with torch.no_grad():
        rec_c, d_c = cm(clips, tcs)
    assert torch.equal(d["indices"], d_c["indices"])
    for a, b in zip(rec, rec_c):
        assert torch.equal(a, b)
    rec_t, _ = cm(clips, tcs)
    torch.stack([r.float().abs().mean() for r in rec_t]).mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


def test_scaled_config_long_sequences():
    """BASELINE configs[4] / SURVEY C5: a 32x256x256 clip with 256 latent tokens (8448 packed rows: 66 attention tiles per
    head). Size-independent properties at full size: results do not depend on the batch composition (bit-exact),
    decode_indices(forward().indices) reproduces forward()'s reconstruction (bit-exact), and one training step on it
    yields finite gradients for every parameter."""
    model = build_model(True).cuda().eval()
    big, small = O.make_clips([(32, 256, 256), (4, 16, 24)], 11)
    big, small = big.cuda(), small.cuda()
    with torch.no_grad():
        rec_a, d_a = model([big], [256])
        rec_b, d_b = model([small, big], [3, 256])
        rec_c = model.decode_indices(d_a["indices"], [(32, 256, 256)], [256])
    assert d_a["indices"].shape == (256,)
    assert torch.equal(d_a["indices"], d_b["indices"][3:])
    assert torch.equal(rec_a[0], rec_b[1])
    assert torch.equal(rec_a[0], rec_c[0])
    assert torch.isfinite(rec_a[0].float()).all()
    model.train()
    rec_t, _ = model([big], [256])
    (rec_t[0].float() - big.float()).abs().mean().backward()
    for k, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
    assert sum(float(p.grad.abs().sum()) for p in model.parameters()) > 0


@pytest.mark.parametrize("size,stress,shape", [
    ("tiny", False, (32, 256, 256)), ("tiny", True, (32, 256, 256)),
    # base: ~8 TFLOP of CPU oracle for the full clip (about 25 s on the GPU box's host cores)
    ("base", False, (32, 256, 256)),
])
def test_scaled_config_matches_oracle(size, stress, shape):
    """BASELINE configs[4] / SURVEY C5 against the ORACLE at full size: one 32x256x256 clip with 256 latent tokens (8448
    packed rows, 132 kv sub-tiles per attention row) through `tiny` and `base` stacks (base: width 768, 12 layers, 12/4
    heads -- the unfused residual path and 3-head kv groups). Same tolerances as the C1 cases."""
    t = 256
    model = build_model(stress, enc=size, dec=size)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda().eval()
    clip = O.make_clips([shape], 21)
    with torch.no_grad():
        z = model.encoder([clip[0].cuda()], [t])
        x_q, d = model.encode([clip[0].cuda()], [t])
        rec = model.decode(x_q, [t], [shape])
    z_o = O.encoder_forward(sd, size, PATCH, clip, [t])
    _check(z, z_o, stress, f"C5 {size} z vs oracle")
    res_idx = O.fsq_forward(z_o, LEVELS)[1]
    _check_indices(d["indices"], res_idx, z, z_o, f"C5 {size} indices vs oracle")
    rec_o = O.decoder_forward(sd, size, PATCH, x_q.float().cpu(), [t], [shape])  # decoder on OUR codes: no flip coupling
    _check(rec[0], rec_o[0], stress, f"C5 {size} recon vs oracle")


def test_uint8_frames_equal_host_normalised_clips(golden_stress):
    """Decoded uint8 frames handed straight to the model give bit-identical tokens and reconstructions to clips that
    were normalised on the host the way the reference's dataset does (video_dataset.py:118-119)."""
    model = build_model(True).cuda().eval()
    g = torch.Generator().manual_seed(4)
    shapes, tcs = [(8, 64, 48), (4, 16, 24)], [16, 3]
    raw = [torch.randint(0, 256, (3, *s), generator=g, dtype=torch.uint8) for s in shapes]
    host = [((r.to(torch.bfloat16) / 255) * 2 - 1) for r in raw]
    with torch.no_grad():
        rec_a, d_a = model([h.cuda() for h in host], tcs)
        rec_b, d_b = model([r.cuda() for r in raw], tcs)
        rec_c, d_c = model.tokenize_reconstruct_([r.cuda() for r in raw], tcs)
        rec_c = [r.clone() for r in rec_c]
        rec_e, d_e = model.tokenize_reconstruct_([r.cuda() for r in raw], tcs, with_error=True)  # separate normalise pass
        rec_e = [r.clone() for r in rec_e]
    assert torch.equal(d_a["indices"], d_b["indices"]) and torch.equal(d_a["indices"], d_c["indices"])
    assert torch.equal(d_a["indices"], d_e["indices"])
    for a, b, c, e, h in zip(rec_a, rec_b, rec_c, rec_e, host):
        assert b.dtype == torch.bfloat16
        assert torch.equal(a, b) and torch.equal(a, c) and torch.equal(a, e)
    l1 = torch.stack([(h.cuda().float() - a.float()).abs().sum() for h, a in zip(host, rec_a)]).double()
    assert torch.allclose(d_e["clip_error"][:, 0], l1, rtol=1e-5)
    # kernel level: the fused gather (ttk_patchify_u8) == ttk_normalize_u8 followed by ttk_patchify, bit for bit
    from titok_video_b200 import _lib, engine
    from titok_video_b200.plan import make_plan

    pl = make_plan(shapes, tcs, (4, 8, 8))
    flat8 = torch.cat([r.reshape(-1) for r in raw]).cuda()
    flatb = torch.empty(flat8.numel(), dtype=torch.bfloat16, device="cuda")
    geom = torch.from_numpy(pl.geom).cuda()
    pa = torch.empty((pl.G, 768), dtype=torch.bfloat16, device="cuda")
    pb = torch.empty_like(pa)
    st = engine._stream()
    _lib.call("ttk_normalize_u8", engine._ptr(flat8), engine._ptr(flatb), flat8.numel(), st)
    _lib.call("ttk_patchify", engine._ptr(flatb), engine._ptr(geom), 3, 4, 8, 8, engine._ptr(pa), 768, pl.G, st)
    _lib.call("ttk_patchify_u8", engine._ptr(flat8), engine._ptr(geom), 3, 4, 8, 8, engine._ptr(pb), 768, pl.G, st)
    torch.cuda.synchronize()
    assert torch.equal(pa, pb)


def test_tokens_written_to_a_container_decode_to_the_same_clips(tmp_path, golden_stress):
    """tokenise -> write_tokens -> read_tokens -> decode_indices reproduces forward()'s reconstruction bit for bit."""
    from titok_video_b200.data import read_tokens, write_tokens

    model = build_model(True).cuda().eval()
    shapes, tcs = [(8, 64, 48), (4, 16, 24)], [16, 3]
    clips = [c.cuda() for c in O.make_clips(shapes, 0)]
    with torch.no_grad():
        rec, d = model(clips, tcs)
        per_clip = torch.split(d["indices"], tcs)
        p = str(tmp_path / "t.ttkv")
        write_tokens(p, per_clip, shapes, model.quantize.codebook_size)
        idx, grids, K = read_tokens(p)
        assert K == 4375 and grids == shapes
        rec2 = model.decode_indices([i.cuda() for i in idx], grids)
    for a, b in zip(rec, rec2):
        assert torch.equal(a, b)


def test_bucketed_graph_replay_equals_the_per_composition_path():
    """ragged batch compositions that fall into ONE shape bucket are served by ONE captured launch sequence
    (TiTok.tokenize_reconstruct_bucketed_); tokens, reconstructions and per-clip errors are bit-identical to the eager
    per-composition path for every composition, including uint8 frames and a batch that opens a second bucket."""
    from titok_video_b200 import engine

    model = build_model(True).cuda().eval()
    g = torch.Generator().manual_seed(12)
    comps = [
        ([(8, 64, 48), (4, 16, 24), (8, 32, 32)], [16, 3, 8]),
        ([(4, 16, 24), (8, 64, 48)], [1, 40]),
        ([(12, 40, 24), (8, 32, 32), (4, 16, 24), (4, 24, 16), (8, 16, 16)], [7, 7, 2, 30, 5]),
        ([(16, 168, 168), (16, 168, 168)], [128, 64]),               # a different (larger) bucket: 3528 patches
        ([(8, 64, 48), (4, 16, 24), (8, 32, 32)], [16, 3, 8]),        # the first composition again
    ]
    engine._BUCKET_CACHE.clear()
    graphs_before = None
    for n, (shapes, tcs) in enumerate(comps):
        clips = [(torch.rand((3, *s), generator=g) * 2 - 1).to(torch.bfloat16).cuda() for s in shapes]
        with torch.no_grad():
            rec_e, d_e = model.tokenize_reconstruct_(clips, tcs, use_graph=False, with_error=True)
            rec_e = [r.clone() for r in rec_e]
            idx_e, err_e = d_e["indices"].clone(), d_e["clip_error"].clone()
            rec_b, d_b = model.tokenize_reconstruct_bucketed_(clips, tcs, with_error=True)
        assert torch.equal(d_b["indices"], idx_e), n
        # (the per-clip sums are reduced over a grid sized by the bucket's upper bound: same values, other summation order)
        assert torch.allclose(d_b["clip_error"], err_e, rtol=1e-6, atol=0), n
        for a, b in zip(rec_b, rec_e):
            assert a.shape == b.shape and torch.equal(a, b), n
        if n == 2:
            graphs_before = sum(len(bp.graphs) for bp in engine._BUCKET_CACHE.values())
            assert len(engine._BUCKET_CACHE) == 1 and graphs_before == 1  # three compositions, one bucket, one graph
    assert len(engine._BUCKET_CACHE) == 2
    # uint8 frames through the same bucket (own graph: the patch gather differs)
    shapes, tcs = comps[1]
    raw = [torch.randint(0, 256, (3, *s), generator=g, dtype=torch.uint8).cuda() for s in shapes]
    with torch.no_grad():
        rec_e, d_e = model.tokenize_reconstruct_(raw, tcs, use_graph=False)
        rec_e, idx_e = [r.clone() for r in rec_e], d_e["indices"].clone()
        rec_b, d_b = model.tokenize_reconstruct_bucketed_(raw, tcs)
    assert torch.equal(d_b["indices"], idx_e)
    for a, b in zip(rec_b, rec_e):
        assert torch.equal(a, b)
