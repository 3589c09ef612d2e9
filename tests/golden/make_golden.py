"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU through
oracle/ref_shim.py. Run in the build container:  python tests/golden/make_golden.py

The fixtures pin (a) the oracle restatement and (b) the CUDA path to the reference's own outputs. Weights are
not stored: they are regenerated from seeds (torch.manual_seed(seed); TiTok(cfg) draws the same stream in the
reference and in titok_video_b200 because module construction order and initialisers coincide); a checksum of
every parameter is stored so a mismatch in that assumption is detected rather than silently accepted.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim, titok_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
LEVELS = [7, 5, 5, 5, 5]
PATCH = [4, 8, 8]
SHAPES = [(8, 32, 32), (4, 16, 24), (8, 64, 48)]
TCS = [8, 3, 16]


def bits(t: torch.Tensor) -> np.ndarray:
    return t.detach().to(torch.bfloat16).contiguous().view(torch.int16).numpy().copy()


def checksum(sd) -> np.ndarray:
    return np.array([float(v.double().sum()) for _, v in sorted(sd.items())] +
                    [float(v.double().abs().sum()) for _, v in sorted(sd.items())])


def titok_case(name: str, stress: bool):
    m = ref_shim.build_reference_titok(fsq_levels=LEVELS, patch_size=PATCH, seed=42)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    if stress:
        O.stress_init_(sd, 1)
        m.load_state_dict(sd)
    mb = m.to(torch.bfloat16).eval()
    clips = O.make_clips(SHAPES, 0)
    tc = torch.tensor(TCS, dtype=torch.int32)
    grids = torch.tensor([c.shape[1:] for c in clips], dtype=torch.int32)
    with torch.no_grad():
        z = mb.encoder(clips, tc)
        codes, d = mb.quantize(z)
        recon = mb.decoder(codes, tc, grids)
        recon2, d2 = mb(clips, tc)
        recon3 = mb.decode_indices(d["indices"], grids, tc)
    assert torch.equal(d["indices"], d2["indices"])
    for a, b, c in zip(recon, recon2, recon3):
        assert torch.equal(a, b) and torch.equal(a, c)
    arrays = {
        "levels": np.array(LEVELS), "patch": np.array(PATCH), "shapes": np.array(SHAPES), "token_counts": np.array(TCS),
        "weight_seed": np.array(42), "stress": np.array(int(stress)), "clip_seed": np.array(0),
        "weight_checksum": checksum(sd),
        "z_bits": bits(z), "codes_bits": bits(codes), "indices": d["indices"].numpy().astype(np.int32),
    }
    for i, rr in enumerate(recon):
        arrays[f"recon{i}_bits"] = bits(rr)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **arrays)
    print(name, "indices", d["indices"].tolist())


def fsq_case():
    _, _, fsq_mod, cb_mod = ref_shim.reference_modules()
    arrays = {}
    for tag, levels in [("a", [7, 5, 5, 5, 5]), ("b", [8, 8, 8, 6, 5]), ("c", [4, 3])]:
        q = fsq_mod.FSQ(levels)
        g = torch.Generator().manual_seed(7)
        z = torch.randn((4096, len(levels)), generator=g) * 2.0
        # include exact boundary / saturation probes
        z[:16] = torch.linspace(-6, 6, 16).unsqueeze(-1)
        codes, d = q(z)
        zb = z.to(torch.bfloat16)
        codes_b, db = q(zb)
        arrays[f"{tag}_levels"] = np.array(levels)
        arrays[f"{tag}_z"] = z.numpy()
        arrays[f"{tag}_codes"] = codes.numpy()
        arrays[f"{tag}_indices"] = d["indices"].numpy()
        arrays[f"{tag}_codes_bf16_bits"] = bits(codes_b)
        arrays[f"{tag}_indices_bf16"] = db["indices"].numpy()
        arrays[f"{tag}_basis"] = q._basis.numpy()
        arrays[f"{tag}_codebook_size"] = np.array(q.codebook_size)
        arrays[f"{tag}_implicit_codebook"] = q.implicit_codebook.numpy()
        arrays[f"{tag}_i2c"] = q.indices_to_codes(d["indices"]).numpy()
        assert torch.equal(q.indices_to_codes(d["indices"]), codes)
        assert torch.equal(torch.cdist(codes, q.implicit_codebook).argmin(-1).to(torch.int32), d["indices"])
    # CodebookLogger known answer
    lg = cb_mod.CodebookLogger(16)
    g = torch.Generator().manual_seed(3)
    samples = [torch.randint(0, 16, (int(n),), generator=g, dtype=torch.int32) for n in torch.randint(1, 9, (20,), generator=g)]
    lg(samples)
    sc = lg.get_scores()
    arrays["logger_samples"] = np.concatenate([s.numpy() for s in samples])
    arrays["logger_lens"] = np.array([len(s) for s in samples])
    arrays["logger_usage"] = np.array(float(sc["codebook/usage_percent"]))
    arrays["logger_entropy"] = np.array(float(sc["codebook/entropy"]))
    np.savez_compressed(os.path.join(OUT, "fsq_kat.npz"), **arrays)


def rope_case():
    install = ref_shim.install
    install()
    import importlib

    rope_mod = importlib.import_module("model.base.rope")
    rp = rope_mod.RoPE(head_dim=64, grid_dims=3)
    grids = torch.tensor([[2, 3, 4], [1, 2, 2]], dtype=torch.int32)
    tcs = torch.tensor([3, 5], dtype=torch.int32)
    f = rp(grids, tcs, torch.device("cpu"))
    g = torch.Generator().manual_seed(5)
    x = torch.randn((f.shape[0], 4, 64), generator=g).to(torch.bfloat16)
    y = rope_mod.apply_rotary_emb(x, f)
    np.savez_compressed(os.path.join(OUT, "rope_kat.npz"), grids=grids.numpy(), token_counts=tcs.numpy(),
                        cos=f.real.numpy(), sin=f.imag.numpy(), x_bits=bits(x), y_bits=bits(y))


GRAD_SHAPES = [(8, 32, 32), (4, 16, 24)]
GRAD_TCS = [8, 3]
GRAD_SAMPLES = 2048


def grad_sample_index(numel: int) -> np.ndarray:
    """Indices of the (at most GRAD_SAMPLES) gradient entries a fixture keeps per parameter."""
    stride = max(1, -(-numel // GRAD_SAMPLES))
    return np.arange(0, numel, stride)


def grads_case(name: str, stress: bool):
    """Parameter gradients of one generator step (L1 reconstruction loss, train.py:68-80; loss_module.py:118) from the
    unmodified reference in bf16 on CPU. Per parameter: the gradient's norm and a strided sample of its entries."""
    m = ref_shim.build_reference_titok(fsq_levels=LEVELS, patch_size=PATCH, seed=42)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    if stress:
        O.stress_init_(sd, 1)
        m.load_state_dict(sd)
    mb = m.to(torch.bfloat16).train()
    clips = O.make_clips(GRAD_SHAPES, 0)
    tc = torch.tensor(GRAD_TCS, dtype=torch.int32)
    recon, d = mb(clips, tc)
    loss = torch.stack([(r_.float() - c.float()).abs().mean() for c, r_ in zip(clips, recon)]).mean()
    loss.backward()
    arrays = {"shapes": np.array(GRAD_SHAPES), "token_counts": np.array(GRAD_TCS), "stress": np.array(int(stress)),
              "loss": np.array(float(loss)), "indices": d["indices"].numpy().astype(np.int32),
              "weight_checksum": checksum(sd)}
    for k, p in mb.named_parameters():
        g = (p.grad if p.grad is not None else torch.zeros_like(p)).float().reshape(-1)
        arrays[f"norm/{k}"] = np.array(float(g.double().norm()))
        arrays[f"sample/{k}"] = g[torch.from_numpy(grad_sample_index(g.numel()))].numpy()
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **arrays)
    print(name, "loss", float(loss))


if __name__ == "__main__":
    if "--grads-only" in sys.argv:
        grads_case("titok_grads_default", stress=False)
        grads_case("titok_grads_stress", stress=True)
        sys.exit(0)
    grads_case("titok_grads_default", stress=False)
    grads_case("titok_grads_stress", stress=True)
    titok_case("titok_default", stress=False)
    titok_case("titok_stress", stress=True)
    fsq_case()
    rope_case()
    print("written to", OUT)
