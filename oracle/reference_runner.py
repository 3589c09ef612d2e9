#!/usr/bin/env python
"""TEST / BASELINE INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference (through oracle/ref_shim.py) as a separate
process, so that nothing of the reference's third-party stack (flash-attn, Triton) is ever loaded into a process that
runs the product, and nothing of the product (titok_video_b200, libtitok_b200.so) into the reference's.

    python oracle/reference_runner.py parity  --mode gpu --shapes '[[16,168,168],[8,128,128]]' --tcs '[128,64]' \
                                              --stress 1 --out /tmp/ref.npz
    python oracle/reference_runner.py kernels --out /tmp/k.npz          # flash-attn varlen + Triton RMSNorm on seeded inputs
    python oracle/reference_runner.py bench   --mode gpu|cpu --batch 64 --steps 10 --warmup 3      # one JSON line

mode gpu = the reference's real path: flash_attn_varlen_func (flash-attn 2.8.3 CUDA extension), flash-attn's Triton
RMSNorm, cuBLAS via nn.Linear, ATen elementwise; only `xformers.ops.SwiGLU` (dead code in
transformer.py:59-66) is a placeholder. mode cpu = the three stand-ins of ref_shim (the reference has no CPU path).

Weights always come from the reference's own `TiTok(cfg)` after `torch.manual_seed(42)` (tiny.yaml:76); the stress
variant re-draws every 2-D parameter like tests/conftest.build_model does (oracle.stress_init_).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

LEVELS = [7, 5, 5, 5, 5]
PATCH = [4, 8, 8]
CLIP_A = (16, 168, 168)
TOKENS_A = 128


def _bits(t: torch.Tensor) -> np.ndarray:
    return t.detach().to(torch.bfloat16).contiguous().view(torch.int16).cpu().numpy().copy()


def _checksum(sd) -> np.ndarray:
    return np.array([float(v.double().sum()) for _, v in sorted(sd.items())] +
                    [float(v.double().abs().sum()) for _, v in sorted(sd.items())])


def _model(mode: str, stress: bool, enc="tiny", dec="tiny"):
    from oracle import ref_shim, titok_oracle as O

    m = ref_shim.build_reference_titok(fsq_levels=LEVELS, patch_size=PATCH, encoder_size=enc, decoder_size=dec, seed=42,
                                       mode=mode)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    if stress:
        O.stress_init_(sd, 1)
        m.load_state_dict(sd)
    return m, sd


def cmd_parity(a):
    """z / codes / indices / reconstructions of the live reference on seeded clips (oracle.make_clips)."""
    from oracle import titok_oracle as O

    shapes = [tuple(s) for s in json.loads(a.shapes)]
    tcs = json.loads(a.tcs)
    m, sd = _model(a.mode, bool(a.stress), a.size, a.size)
    dev = torch.device("cuda" if a.mode == "gpu" else "cpu")
    mb = m.to(torch.bfloat16).to(dev).eval()
    clips = [c.to(dev) for c in O.make_clips(shapes, a.clip_seed)]
    tc = torch.tensor(tcs, dtype=torch.int32, device=dev)
    grids = torch.tensor([c.shape[1:] for c in clips], dtype=torch.int32, device=dev)
    with torch.no_grad():
        z = mb.encoder(clips, tc, grids.clone())
        codes, d = mb.quantize(z)
        recon = mb.decoder(codes, tc, grids.clone())
        recon2, d2 = mb(clips, tc)
    assert torch.equal(d["indices"], d2["indices"])
    arrays = {"shapes": np.array(shapes), "token_counts": np.array(tcs), "stress": np.array(int(a.stress)),
              "weight_checksum": _checksum(sd), "z_bits": _bits(z), "codes_bits": _bits(codes),
              "indices": d["indices"].cpu().numpy().astype(np.int32), "mode": np.array(a.mode)}
    for i, rr in enumerate(recon):
        arrays[f"recon{i}_bits"] = _bits(rr)
    np.savez_compressed(a.out, **arrays)
    print(json.dumps({"ok": True, "mode": a.mode, "tokens": int(z.shape[0])}))


def cmd_kernels(a):
    """The two third-party kernels the reference's arithmetic lives in, on seeded inputs (GPU only):
    flash_attn_varlen_func (transformer.py:100) and flash-attn's Triton RMSNorm (blocks.py:27)."""
    from flash_attn import flash_attn_varlen_func
    from flash_attn.ops.triton.layer_norm import RMSNorm

    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(a.seed)
    lens = json.loads(a.lens)
    total = sum(lens)
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device=dev)
    arrays = {"lens": np.array(lens)}
    for tag, scale in (("unit", 1.0), ("wide", 4.0)):
        q = (torch.randn((total, 4, 64), generator=g) * scale).to(torch.bfloat16).to(dev)
        k = (torch.randn((total, 2, 64), generator=g) * scale).to(torch.bfloat16).to(dev)
        v = torch.randn((total, 2, 64), generator=g).to(torch.bfloat16).to(dev)
        o = flash_attn_varlen_func(q, k, v, cu, cu, max(lens), max(lens))
        arrays.update({f"{tag}_q": _bits(q), f"{tag}_k": _bits(k), f"{tag}_v": _bits(v), f"{tag}_o": _bits(o)})
    for w in (256, 512, 768, 1024):
        x = (torch.randn((777, w), generator=g) * 3.0).to(torch.bfloat16).to(dev)
        n = RMSNorm(w).to(dev)
        with torch.no_grad():
            n.weight.copy_((torch.randn((w,), generator=g) * 0.5 + 1.0).to(dev))
            y = n(x)
        arrays.update({f"rms{w}_x": _bits(x), f"rms{w}_w": n.weight.detach().float().cpu().numpy(), f"rms{w}_y": _bits(y)})
    np.savez_compressed(a.out, **arrays)
    print(json.dumps({"ok": True}))


def cmd_disc(a):
    """The reference's ReconstructionLoss (model/losses/loss_module.py, unmodified) on seeded clips: generator-side GAN
    loss and the discriminator step's loss dict + parameter gradients. LPIPS / gram are switched off in the config
    (their weights need a download); the discriminator is the reference's TiTokEncoder(out_channels=1)."""
    import importlib

    from oracle import ref_shim, titok_oracle as O

    ref_shim.install(a.mode)
    lm = importlib.import_module("model.losses.loss_module")
    A = ref_shim.AttrDict.wrap
    cfg = A({"tokenizer": {"losses": {"disc_weight": 0.4, "perceptual_weight": 0.0, "gram_weight": 0.0,
                                      "perceptual_samples_per_step": 24, "perceptual_sampling_size": 128}},
             "discriminator": {"model": {"patch_size": PATCH, "model_size": "tiny"},
                               "losses": {"gp_weight": 0.1, "gp_noise": 0.1, "centering_weight": 0.01}},
             "training": {"main": {"torch_compile": False, "max_steps": 1000}}})
    torch.manual_seed(5)
    loss = lm.ReconstructionLoss(cfg)
    dev = torch.device("cuda" if a.mode == "gpu" else "cpu")
    sd = {k: v.detach().clone() for k, v in loss.disc_model.state_dict().items()}
    if a.stress:
        O.stress_init_(sd, 2)
        loss.disc_model.load_state_dict(sd)
    loss = loss.to(torch.bfloat16).to(dev)
    shapes = [tuple(s) for s in json.loads(a.shapes)]
    target = [c.to(dev) for c in O.make_clips(shapes, 31)]
    recon = [c.to(dev) for c in O.make_clips(shapes, 32)]
    g = torch.Generator().manual_seed(33)
    unit = [torch.randn(c.shape, generator=g).to(torch.bfloat16).to(dev) for c in target]
    noise = [u * 0.1 for u in unit]  # what `torch.randn_like(x) * self.gp_noise` (loss_module.py:186) evaluates to
    # the reference draws its noise with randn_like; make that call return ours so both sides see the same numbers
    it = iter(unit)
    real_randn_like = torch.randn_like
    torch.randn_like = lambda x, *aa, **kw: next(it)
    try:
        total, logs = loss(target, recon, disc_forward=True)
    finally:
        torch.randn_like = real_randn_like
    total.backward()
    arrays = {"shapes": np.array(shapes), "total": np.array(float(total)), "weight_checksum": _checksum(sd)}
    for k, v in logs.items():
        arrays["log/" + k] = np.array(float(v))
    for k, p_ in loss.disc_model.named_parameters():
        gg = (p_.grad if p_.grad is not None else torch.zeros_like(p_)).float().reshape(-1).cpu()
        arrays["norm/" + k] = np.array(float(gg.double().norm()))
        stride = max(1, -(-gg.numel() // 2048))
        arrays["sample/" + k] = gg[::stride].numpy()
    for i, n in enumerate(noise):
        arrays[f"noise{i}_bits"] = _bits(n)
    # generator side: frozen discriminator, gradient w.r.t. the fake pixels
    rec_leaf = [r.detach().clone().requires_grad_(True) for r in recon]
    for p_ in loss.disc_model.parameters():
        p_.requires_grad = False
    lr, lf = loss.disc_wrapper([t.detach() for t in target]), loss.disc_wrapper(rec_leaf)
    g_loss = torch.nn.functional.softplus(-(lf - lr))
    g_loss.mean().backward()
    arrays["g_loss"] = g_loss.detach().float().cpu().numpy()
    for i, r_ in enumerate(rec_leaf):
        arrays[f"g_pixgrad{i}_norm"] = np.array(float(r_.grad.float().double().norm()))
    np.savez_compressed(a.out, **arrays)
    print(json.dumps({"ok": True, "total": float(total)}))


def _bench_inputs(batch, dev, dtype, shape=CLIP_A, tokens=TOKENS_A, sets=2):
    g = torch.Generator().manual_seed(1000)
    out = []
    for _ in range(sets):
        out.append([(torch.rand((3, *shape), generator=g) * 2 - 1).to(dtype).to(dev) for _ in range(batch)])
    return out, torch.tensor([tokens] * batch, dtype=torch.int32, device=dev)


def gpu_bench(batch, steps, warmup, shape=CLIP_A, tokens=TOKENS_A):
    """The reference's own forward (titok.py:68-74) on the GPU with its real kernels, bf16 weights and clips, no_grad."""
    m, _ = _model("gpu", False)
    dev = torch.device("cuda")
    mb = m.to(torch.bfloat16).to(dev).eval()
    sets, tc = _bench_inputs(batch, dev, torch.bfloat16, shape, tokens)
    with torch.no_grad():
        for i in range(max(warmup, 3)):
            mb(sets[i % 2], tc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            mb(sets[i % 2], tc)
        e1.record()
        torch.cuda.synchronize()
        w1 = time.perf_counter()
    ms = e0.elapsed_time(e1) / steps
    out = {"impl": "reference-gpu", "clips_per_s": batch / (ms * 1e-3), "ms_per_step": ms,
           "wall_ms_per_step": (w1 - w0) * 1e3 / steps, "clips_per_step": batch, "steps": steps,
           "what": "unmodified reference TiTok.forward (bf16 weights + clips, no_grad): flash-attn 2.8.3 varlen kernels, "
                   "flash-attn Triton RMSNorm, cuBLAS nn.Linear, ATen elementwise; xformers.SwiGLU placeholder only"}
    # kernel census of one forward: launches, attention kernel time, host syncs (cudaStreamSynchronize / memcpy DtoH)
    try:
        from torch.profiler import ProfilerActivity, profile

        with torch.no_grad(), profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            mb(sets[0], tc)
            torch.cuda.synchronize()
        kern, attn_us, attn_n, syncs = 0, 0.0, 0, 0
        for ev in prof.events():
            dt = getattr(ev, "device_type", None)
            if str(dt).endswith("CUDA"):
                name = ev.name
                if name.startswith("Memcpy") or name.startswith("Memset"):
                    if "DtoH" in name:
                        syncs += 1
                    continue
                kern += 1
                if "flash" in name.lower():
                    attn_us += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
                    attn_n += 1
            elif ev.name in ("cudaStreamSynchronize", "cudaDeviceSynchronize", "cudaEventSynchronize"):
                syncs += 1
        out.update({"launches": kern, "attn_launches": attn_n, "attn_us_per_launch": (attn_us / attn_n) if attn_n else None,
                    "syncs": syncs})
    except Exception as e:  # profiler (CUPTI) unavailable: the timing above stands
        out.update({"launches": None, "attn_us_per_launch": None, "syncs": None, "census_error": repr(e)[:200]})
    # attention alone at this shape (CUDA events)
    try:
        from flash_attn import flash_attn_varlen_func

        g = 1
        for s_, p_ in zip(shape, PATCH):
            g *= s_ // p_
        s = g + tokens
        cu = torch.arange(0, (batch + 1) * s, s, dtype=torch.int32, device=dev)
        q = torch.randn((batch * s, 4, 64), device=dev, dtype=torch.bfloat16)
        k = torch.randn((batch * s, 2, 64), device=dev, dtype=torch.bfloat16)
        v = torch.randn((batch * s, 2, 64), device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            flash_attn_varlen_func(q, k, v, cu, cu, s, s)
        torch.cuda.synchronize()
        n = 10
        e0.record()
        for _ in range(n):
            flash_attn_varlen_func(q, k, v, cu, cu, s, s)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        out["attn_alone_us_per_launch"] = us
        out["attn_alone_tflops"] = batch * 4.0 * s * s * 256 / (us * 1e-6) / 1e12
    except Exception as e:
        out["attn_alone_error"] = repr(e)[:200]
    return out


def cpu_bench(sample_clips, steps, warmup):
    """The reference's own files on the host cores under the three CPU stand-ins (the reference has no CPU path of its
    own). bf16 (the reference's data dtype) and fp32 are both tried once; the faster one is timed."""
    torch.set_num_threads(os.cpu_count() or 1)
    m, _ = _model("cpu", False)
    dev = torch.device("cpu")
    best = None
    for dt in (torch.float32, torch.bfloat16):
        mm = m.to(dt).eval()
        sets, tc = _bench_inputs(sample_clips, dev, dt, sets=1)
        with torch.no_grad():
            mm(sets[0], tc)
            t0 = time.perf_counter()
            mm(sets[0], tc)
            dtm = time.perf_counter() - t0
        if best is None or dtm < best[0]:
            best = (dtm, dt)
    dt = best[1]
    mm = m.to(dt).eval()
    sets, tc = _bench_inputs(sample_clips, dev, dt, sets=2)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            mm(sets[i % 2], tc)
            d = time.perf_counter() - t0
            if i >= warmup:
                times.append(d)
    total = sum(times)
    return {"clips_per_s": sample_clips * len(times) / total, "ms_per_step": 1e3 * total / len(times),
            "cores": torch.get_num_threads(), "dtype": str(dt).replace("torch.", ""),
            "sample": f"{sample_clips} clip(s) 3x16x168x168 / 128 tokens per step, {len(times)} timed steps; unmodified reference "
                      f"files under CPU stand-ins for flash-attn varlen attention / Triton RMSNorm / xformers, "
                      f"{str(dt).replace('torch.', '')}, no_grad"}


def cmd_bench(a):
    if a.mode == "gpu":
        print(json.dumps(gpu_bench(a.batch, a.steps, a.warmup)), flush=True)
    else:
        print(json.dumps(cpu_bench(a.batch, a.steps, a.warmup)), flush=True)


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("parity")
    p.add_argument("--mode", default="gpu", choices=["gpu", "cpu"])
    p.add_argument("--shapes", required=True)
    p.add_argument("--tcs", required=True)
    p.add_argument("--stress", type=int, default=0)
    p.add_argument("--size", default="tiny")
    p.add_argument("--clip-seed", type=int, default=0)
    p.add_argument("--out", required=True)
    p.set_defaults(fn=cmd_parity)
    p = sub.add_parser("kernels")
    p.add_argument("--lens", default="[1892, 576, 130, 1]")
    p.add_argument("--seed", type=int, default=11)
    p.add_argument("--out", required=True)
    p.set_defaults(fn=cmd_kernels)
    p = sub.add_parser("disc")
    p.add_argument("--mode", default="cpu", choices=["gpu", "cpu"])
    p.add_argument("--shapes", default="[[8, 32, 32], [4, 16, 24], [8, 64, 48]]")
    p.add_argument("--stress", type=int, default=1)
    p.add_argument("--out", required=True)
    p.set_defaults(fn=cmd_disc)
    p = sub.add_parser("bench")
    p.add_argument("--mode", default="gpu", choices=["gpu", "cpu"])
    p.add_argument("--batch", type=int, default=64)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.set_defaults(fn=cmd_bench)
    a = ap.parse_args()
    a.fn(a)


if __name__ == "__main__":
    main()
