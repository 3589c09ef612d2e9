#!/usr/bin/env python
"""Benchmark of the tokenizer hot path (encode -> quantize -> decode) on B200.

    python bench.py --gpus N --steps K --warmup W            # one process per GPU (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2]): configs/tiny.yaml, batch inference tokenisation+reconstruction,
per GPU per step a packed batch of `--batch` canonical clips A = 3x16x168x168 (bf16) with 128 latent tokens each,
synthetic uniform[-1,1) pixels, random-init weights (seed 42, the reference's own initialiser). Clips are sharded
over ranks (weak scaling: fixed per-GPU batch); there is no data-path collective, NCCL is used once per run for
the codebook-usage histogram reduction and for the max-over-ranks timing.

One JSON line on rank 0:
  value    clips/s with the inputs resident in HBM (device-timed, CUDA events, max over ranks)
  e2e      clips/s through the public API from pinned HOST clips to HOST results, H2D and D2H inside the timed region
  roofline the dominant kernel (by device time inside the timed region) against the measured tensor peak
  cpu_baseline / --impl reference: the CPU oracle port of the reference path on the host cores, bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

try:
    ORIG_AFFINITY = os.sched_getaffinity(0)  # the reference / CPU-baseline child processes get all host cores back
except Exception:
    ORIG_AFFINITY = None

CLIP_A = (16, 168, 168)
TOKENS_A = 128
PATCH = (4, 8, 8)
LEVELS = (7, 5, 5, 5, 5)
WIDTH, LAYERS, HQ, HKV, INNER = 256, 4, 4, 2, 704
INPUT_SETS = 4


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(kernel, batch):
    """dram bytes (read + write) per launch of `kernel` from the committed ncu --set full summary of this same bench
    command (profiles/r1_traffic.json, written by scripts/summarize_profiles.py); None when no capture matches."""
    p = next((q for q in (os.path.join(ROOT, "profiles", f"r{r}_traffic.json") for r in (2, 1)) if os.path.exists(q)), None)
    if p is None:
        return None
    try:
        d = json.load(open(p))
        e = d.get(kernel)
        if e and int(e.get("batch", -1)) == int(batch):
            return float(e["dram_bytes_per_launch"])
    except Exception:
        return None
    return None


def clip_flops(shape=CLIP_A, t=TOKENS_A):
    """Forward FLOPs of one clip through encoder + decoder."""
    g = (shape[0] // PATCH[0]) * (shape[1] // PATCH[1]) * (shape[2] // PATCH[2])
    s = g + t
    lin = LAYERS * s * 2 * (WIDTH * (2 * WIDTH + 2 * HKV * 64) + WIDTH * WIDTH + 3 * WIDTH * INNER)
    attn = LAYERS * 4 * s * s * WIDTH
    io = 2 * 768 * WIDTH * g * 2 + 2 * 5 * WIDTH * t * 2
    return 2 * (lin + attn) + io, s, g


def latent_tail_savings(s, t, w, inner, enabled=None):
    """FLOPs of the reference's algorithm that this implementation does NOT execute: the encoder's last layer carries only
    the t latent rows past its attention (engine._layer_latent; the head reads nothing else, blocks.py:101) -- the other
    s - t rows skip attention as queries, out_proj, w12 and w3. 0 when the engine runs all rows (TTK_LATENT_TAIL=0)."""
    if enabled is None:
        from titok_video_b200 import engine

        enabled = engine.LATENT_TAIL
    if not enabled:
        return 0.0
    return float((s - t) * 2 * (w * w + 3 * w * inner) + 4 * s * (s - t) * w)


# ----------------------------------------------------------------------------------------------------
# per-kernel device timing (CUDA events on the launching stream, inside the timed region)
# ----------------------------------------------------------------------------------------------------
class KernelTimer:
    def __init__(self):
        self.events = []

    def begin(self, name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def end(self, name, e0):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.events.append((name, e0, e1))

    def summary(self):
        out = {}
        for name, e0, e1 in self.events:
            ms = e0.elapsed_time(e1)
            a = out.setdefault(name, [0.0, 0])
            a[0] += ms
            a[1] += 1
        return out


def kernel_flops_per_launch(name, B, s, g, t):
    """Algorithmic FLOPs of one launch of a GEMM-class kernel at batch B of clips A."""
    M = B * s
    if name == "ttk_attn_varlen_fwd":
        return B * 4.0 * s * s * WIDTH
    if name == "ttk_gemm_qkv_rope":
        return 2.0 * M * WIDTH * (2 * WIDTH + 2 * HKV * 64)
    if name == "ttk_gemm_geglu":
        return 2.0 * M * WIDTH * 2 * INNER
    if name == "ttk_gemm_resid_norm256":
        return None  # two shapes (K=256 and K=704) share this entry point: reported as time share only
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.gpu), "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        note = None
        if not sm and self.lines:
            # the timed region was shorter than the sampling period: use the sample closest to it and say so
            ts, line = min(self.lines, key=lambda tl: min(abs(tl[0] - t0), abs(tl[0] - t1)))
            f = [x.strip() for x in line.split(",")]
            try:
                sm, mx = [float(f[1])], float(f[2])
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
                note = "no sample fell inside the timed region; nearest sample (%.0f ms away) reported" % (
                    1e3 * min(abs(ts - t0), abs(ts - t1)))
            except (ValueError, IndexError):
                sm = []
        sm.sort()
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


# ----------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path on the host cores
# ----------------------------------------------------------------------------------------------------
def _runner(args, timeout=900):
    """oracle/reference_runner.py in a SEPARATE process (the reference's stack never shares a process with the product);
    returns its JSON line or {"unavailable": why}."""
    env = dict(os.environ)
    env.setdefault("TRITON_CACHE_DIR", "/tmp/triton_cache_ref")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID"):
        env.pop(k, None)
    try:
        def _all_cores():
            if ORIG_AFFINITY:
                os.sched_setaffinity(0, ORIG_AFFINITY)

        r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "reference_runner.py"), *args], capture_output=True,
                           text=True, timeout=timeout, env=env, cwd=ROOT, preexec_fn=_all_cores)
    except subprocess.TimeoutExpired:
        return {"unavailable": "reference runner timed out"}
    if r.returncode != 0:
        return {"unavailable": (r.stderr or "").strip().splitlines()[-1][:300] if r.stderr else f"rc {r.returncode}"}
    for line in reversed(r.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    return {"unavailable": "no JSON line from the reference runner"}


def reference_available():
    return any(os.path.isdir(os.path.join(c, "model")) for c in
               (os.environ.get("TITOK_REFERENCE_ROOT") or "/nonexistent", "/root/reference", os.path.join(ROOT, "baseline", "_ref")))


def cpu_reference_run(steps, warmup, sample_clips=1, in_process=False):
    """CPU arm. kind "reference+shims": the UNMODIFIED reference files (baseline/_ref, vendored by
    scripts/vendor_reference.sh) on the host cores with the three CPU stand-ins for its CUDA-only third-party kernels
    (oracle/ref_shim.py); weights drawn by the reference's own TiTok(cfg). kind "port" (only when no reference copy
    travelled): the oracle restatement. Never imports titok_video_b200."""
    if reference_available():
        if in_process:
            from oracle import reference_runner as RR

            r = RR.cpu_bench(sample_clips, steps, warmup)
        else:
            r = _runner(["bench", "--mode", "cpu", "--batch", str(sample_clips), "--steps", str(steps), "--warmup", str(warmup)])
        if "unavailable" not in r:
            r["kind"] = "reference"  # the reference's own files; its three CUDA-only third-party symbols are stand-ins
            r["shims"] = ["flash_attn.flash_attn_varlen_func -> per-segment SDPA", "flash_attn Triton RMSNorm -> torch fp32",
                          "xformers.ops.SwiGLU -> placeholder (dead code)"]
            return r
    from oracle import titok_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.random_state_dict(42)
    clips = O.make_clips([CLIP_A] * sample_clips, 0)
    tcs = [TOKENS_A] * sample_clips
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.titok_forward(sd, list(LEVELS), list(PATCH), clips, tcs)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    total = sum(times)
    return {"clips_per_s": sample_clips * len(times) / total, "ms_per_step": 1e3 * total / len(times),
            "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample_clips} clip(s) 3x16x168x168 / 128 tokens per step, {len(times)} timed steps, torch CPU oracle port"}


class numa_local:
    """Context manager for pinned host allocations: while it is active the calling thread runs on the CPUs NVML reports
    as local to the rank's GPU, so the pinned pages are first touched on the GPU's own NUMA node; the previous affinity
    is restored afterwards (the launch path itself is NOT pinned to a subset of the cores: on a shared host that costs
    more than it gives). Only used for multi-GPU runs; `desc` goes into the JSON line."""
    desc = "not bound (single GPU)"

    def __init__(self, gpu_index, enabled):
        self.gpu, self.enabled, self.prev = gpu_index, enabled, None

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
            prev = os.sched_getaffinity(0)
            allowed = sorted(set(cpus) & set(prev))
            if allowed:
                os.sched_setaffinity(0, allowed)
                self.prev = prev
                numa_local.desc = f"pinned buffers allocated on the {len(allowed)} CPUs local to the rank's GPU (NVML affinity)"
            else:
                numa_local.desc = "NVML affinity empty or outside the cgroup: unchanged"
        except Exception as e:  # no NVML / not permitted: leave the affinity alone
            numa_local.desc = f"unchanged ({type(e).__name__})"
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            os.sched_setaffinity(0, self.prev)
        return False


def workload_string(B):
    _, s, _ = clip_flops()
    return (f"configs/tiny.yaml batch tokenise+reconstruct (C3): per GPU {B} clips 3x16x168x168 bf16, "
            f"128 latent tokens each, packed rows/step {B * s}")


# ----------------------------------------------------------------------------------------------------
# training step (BASELINE configs[3]): forward + L1 loss + backward + gradient all-reduce + AdamW
# ----------------------------------------------------------------------------------------------------
def train_flops(B):
    """Algorithmic FLOPs of one training step on B clips A: forward + backward (2x forward: dgrad + wgrad; the attention
    backward's recomputation of S and dP is NOT counted)."""
    f, s_rows, _ = clip_flops()
    from titok_video_b200 import backward

    # (with the encoder's last layer carried on the latent rows only, those FLOPs are not executed, forward or backward)
    f -= latent_tail_savings(s_rows, TOKENS_A, WIDTH, INNER, enabled=backward.latent_tail_active(B * s_rows))
    return 3.0 * B * f


def train_leg(T, _lib, dev, world, rank, dist, batch, steps, warmup, graphed=False):
    """One generator training step per iteration through the public module API (train.py:68-80): bf16 autocast forward of
    TiTok (fp32 master parameters), L1 reconstruction loss (loss_module.py:118, plain torch: the loss module is not on
    the path), backward through the CUDA backward kernels, gradient all-reduce over NCCL (N > 1, decoder bucket
    overlapped with the encoder backward), fused AdamW with tiny.yaml's betas / weight decay, codebook histogram of the
    step's indices. Device-timed with CUDA events, max over ranks."""
    from titok_video_b200.config import tiny_config
    from titok_video_b200.dist import GradientAllReducer

    torch.manual_seed(42)
    model = T.TiTok(tiny_config(LEVELS, PATCH)).to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.5, 0.96), weight_decay=1e-4, fused=True)
    red = GradientAllReducer(model)
    hist = torch.zeros(model.quantize.codebook_size, dtype=torch.int32, device=dev)
    gen = torch.Generator().manual_seed(2000 + rank)
    sets = [[(torch.rand((3, *CLIP_A), generator=gen) * 2 - 1).to(torch.bfloat16).to(dev) for _ in range(batch)]
            for _ in range(2)]
    tcs = [TOKENS_A] * batch

    gstep = None
    if graphed:
        from titok_video_b200.train_utils.graphed_step import GraphedTrainStep

        gstep = GraphedTrainStep(model, warmup=1)

    def step(clips):
        opt.zero_grad(set_to_none=True)
        if gstep is not None:
            # forward + loss + backward replayed as ONE CUDA graph once the composition has been seen (the bench's fixed
            # batch repeats every step); the all-reduce and the optimizer stay eager
            with red.no_sync():  # (autograd hooks must not fire collectives inside a capture)
                loss, d = gstep(clips, tcs)
            red.reduce_now()
        else:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                recon, d = model(clips, tcs)
            loss = torch.stack([(r_.float() - c.float()).abs().mean() for c, r_ in zip(clips, recon)]).mean()
            loss.backward()
        red.finish()
        opt.step()
        idx = d["indices"]
        _lib.call("ttk_hist_u32", T.engine._ptr(idx), idx.numel(), hist.numel(), T.engine._ptr(hist), T.engine._stream())
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(warmup, 3)):
        step(sets[i % 2])
    barrier()
    l0 = _lib.LAUNCHES
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    barrier()
    w0 = time.perf_counter()
    evs[0].record()
    for i in range(steps):
        loss = step(sets[i % 2])
        evs[i + 1].record()
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3 / steps
    launches = _lib.LAUNCHES - l0
    e0, e1 = evs[0], evs[-1]
    per_step = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(steps))
    median_ms = per_step[len(per_step) // 2]
    # per-kernel breakdown: separate pass (the per-launch CUDA events cost host time, so they stay out of the timed region)
    timer = KernelTimer()
    prof_steps = 2
    if gstep is None:
        _lib.set_profiler(timer)
        for i in range(prof_steps):
            step(sets[i % 2])
        barrier()
        _lib.set_profiler(None)
    # a host-bound step on a shared host: the MEDIAN step (max over ranks) is the figure, the mean is given beside it
    ms = torch.tensor([median_ms, e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(hist)
    ms_step, mean_ms = float(ms[0].item()), float(ms[1].item())
    ks = timer.summary()
    tot = sum(v[0] for v in ks.values()) or 1.0
    kernels = {k: {"ms_per_step": v[0] / prof_steps, "launches_per_step": v[1] / prof_steps, "share_of_kernel_time": v[0] / tot}
               for k, v in sorted(ks.items(), key=lambda kv: -kv[1][0])}
    pk = peaks()
    tf = train_flops(batch) / (ms_step * 1e-3) / 1e12
    red.remove()
    return {"clips_per_s": world * batch / (ms_step * 1e-3), "ms_per_step": ms_step, "clips_per_gpu_per_step": batch,
            "packed_rows_per_gpu": batch * clip_flops()[1], "loss": float(loss.detach()), "gpu_launches_per_step": launches / steps,
            "tflops_algorithmic": tf, "frac_of_tensor_peak": tf / pk["bf16_tflops_sustained"],
            "latent_tail": bool(__import__("titok_video_b200.backward", fromlist=["x"]).latent_tail_active(batch * clip_flops()[1])),
            "ms_per_step_mean": mean_ms, "timing": "median of the per-step CUDA-event times, max over ranks",
            "wall_ms_per_step": wall_ms, "kernel_ms_per_step": sum(v[0] for v in ks.values()) / prof_steps, "kernels": kernels,
            "graph_replays": (gstep.replays if gstep is not None else 0),
            "what": ("forward + L1 loss + backward REPLAYED AS ONE CUDA GRAPH (train_utils.GraphedTrainStep; the fixed bench "
                     "batch repeats every step -- the reference's ragged batches would not) + " if graphed else "")
                    + "TiTok.forward under bf16 autocast + L1 loss + backward (CUDA backward kernels) + "
                    + ("NCCL gradient all-reduce (per-stack buckets, overlapped) + " if world > 1 else "")
                    + "fused AdamW + codebook histogram; synthetic clips 3x16x168x168 / 128 tokens, random-init weights"}


def gan_leg(T, _lib, dev, batch, steps, warmup):
    """The full training step of train.py:48-115 around the path: generator step (TiTok forward, L1, the
    GAN term through the discriminator = a TiTokEncoder(out_channels=1), backward, AdamW) and discriminator step
    (relativistic loss + noise R1/R2 + centering, loss_module.py:166-214, backward, AdamW). LPIPS is not on the path (a
    VGG whose weights need a download). Measured twice: the discriminator's forwards of each step PACKED into one launch
    sequence (train_utils.PackedDiscriminator) and issued one by one like the reference does (6 encoder calls)."""
    import torch.nn.functional as F

    from titok_video_b200.config import tiny_config
    from titok_video_b200.model.base.utils import init_weights
    from titok_video_b200.train_utils.disc_step import PackedDiscriminator

    torch.manual_seed(42)
    model = T.TiTok(tiny_config(LEVELS, PATCH)).to(dev).train()
    disc = T.TiTokEncoder("tiny", PATCH, 3, 1).apply(init_weights).to(dev).train()
    opt_g = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.5, 0.96), weight_decay=1e-4, fused=True)
    opt_d = torch.optim.AdamW(disc.parameters(), lr=1.5e-5, betas=(0.5, 0.96), weight_decay=1e-4, fused=True)
    pd = PackedDiscriminator(disc, 4)
    gen = torch.Generator().manual_seed(4000)
    clips = [(torch.rand((3, *CLIP_A), generator=gen) * 2 - 1).to(torch.bfloat16).to(dev) for _ in range(batch)]
    tcs = [TOKENS_A] * batch

    def sep_logits(groups):
        return [pd.logits([g])[0] for g in groups]

    def step(packed):
        logits = pd.logits if packed else sep_logits
        # ---- generator
        pd.set_requires_grad(False)
        opt_g.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            recon, d = model(clips, tcs)
            recon_loss = torch.stack([(r_.float() - c.float()).abs().mean() for c, r_ in zip(clips, recon)])
            lr_, lf_ = logits([[c.detach() for c in clips], list(recon)])
            g_loss = F.softplus(-(lf_ - lr_))
            total = (recon_loss + 0.4 * g_loss.float()).mean()
        total.backward()
        opt_g.step()
        # ---- discriminator
        opt_d.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if packed:
                total_d, _ = pd.discriminator_loss(clips, [r_.detach() for r_ in recon], 0.1, 0.1, 0.01)
            else:
                pd.set_requires_grad(True)
                rec = [r_.detach() for r_ in recon]
                noise = [torch.randn_like(x) * 0.1 for x in clips]
                a, b, c2, d2 = sep_logits([clips, rec, [x + n for x, n in zip(clips, noise)], [x + n for x, n in zip(rec, noise)]])
                total_d = (F.softplus(-(a - b)) + 0.1 / 0.01 * ((a - c2) ** 2 + (b - d2) ** 2) + 0.01 * ((a + b) ** 2) / 2).mean()
        total_d.backward()
        opt_d.step()
        return total.detach()

    out = {}
    for name, packed in (("packed", True), ("separate", False)):
        for _ in range(max(warmup, 3)):
            step(packed)
        torch.cuda.synchronize()
        l0 = _lib.LAUNCHES
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            step(packed)
        e1.record()
        torch.cuda.synchronize()
        out[name] = {"ms_per_step": e0.elapsed_time(e1) / steps, "wall_ms_per_step": (time.perf_counter() - w0) * 1e3 / steps,
                     "gpu_launches_per_step": (_lib.LAUNCHES - l0) / steps}
    out["clips_per_gpu_per_step"] = batch
    out["clips_per_s_packed"] = batch / (out["packed"]["ms_per_step"] * 1e-3)
    out["what"] = ("generator step (TiTok fwd + L1 + GAN term via the frozen discriminator + backward + AdamW) and discriminator "
                   "step (relativistic + noise R1/R2 + centering, backward, AdamW); discriminator = TiTokEncoder(out_channels=1), "
                   "4 register tokens; `packed` = each step's discriminator forwards as one launch sequence, `separate` = the "
                   "reference's 6 encoder calls")
    return out


def train_ragged_leg(T, _lib, dev, rank, steps):
    """The regime train.py really runs in: every step a NEW batch composition from the dataloader's token-budget batching
    (titok_video_b200.data.dynamic_batches = video_dataset.py:130-172: shapes in [min_grid, max_grid], 1..128 tokens,
    6144 packed rows per batch), so every step pays the host planner, the metadata upload and the work-list construction
    on top of forward + loss + backward + AdamW. Wall clock around a synchronised loop, rank 0 only."""
    import random

    from titok_video_b200 import engine as _eng
    from titok_video_b200.config import tiny_config
    from titok_video_b200.data import dynamic_batches

    torch.manual_seed(42)
    model = T.TiTok(tiny_config(LEVELS, PATCH)).to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.5, 0.96), weight_decay=1e-4, fused=True)
    rnd = random.Random(1)

    def samples():
        while True:
            shp = (rnd.choice([8, 12, 16]), rnd.choice([128, 136, 144, 152, 160, 168]), rnd.choice([128, 136, 144, 152, 160, 168]))
            yield {"video": (torch.rand((3, *shp), device=dev) * 2 - 1).to(torch.bfloat16)}

    it = dynamic_batches(samples(), list(PATCH), [1, 128], [16, 168, 168], 6144, randrange=rnd.randrange)
    batches = [next(it) for _ in range(steps + 3)]

    def step(b):
        clips, tcs = b["video"], b["token_counts"]
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            recon, d = model(clips, tcs)
        loss = torch.stack([(r_.float() - c.float()).abs().mean() for c, r_ in zip(clips, recon)]).mean()
        loss.backward()
        opt.step()
        return loss

    _eng.clear_caches()
    for b in batches[:3]:
        step(b)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for b in batches[3:]:
        step(b)
    torch.cuda.synchronize()
    w1 = time.perf_counter()
    n_clips = sum(len(b["video"]) for b in batches[3:])
    rows = sum(int(b["token_counts"].sum()) + sum((v.shape[1] // 4) * (v.shape[2] // 8) * (v.shape[3] // 8) for v in b["video"])
               for b in batches[3:])
    return {"clips_per_s": n_clips / (w1 - w0), "ms_per_step_wall": 1e3 * (w1 - w0) / steps, "steps": steps,
            "clips_per_step": n_clips / steps, "packed_rows_per_step": rows / steps,
            "note": "new shapes / token counts every step (dynamic_batches, 6144-row budget): host planning + metadata upload + "
                    "forward + L1 loss + backward + fused AdamW; wall clock"}


def model_flops(size, shape, t):
    """Forward FLOPs of one clip through encoder + decoder stacks of a given size (SURVEY 8d formula)."""
    w, layers, hq, hkv = {"tiny": (256, 4, 4, 2), "small": (512, 8, 8, 2), "base": (768, 12, 12, 4), "large": (1024, 24, 16, 4)}[size]
    inner = 32 * ((int(4.0 * (2 / 3) * w) + 31) // 32)
    g = (shape[0] // PATCH[0]) * (shape[1] // PATCH[1]) * (shape[2] // PATCH[2])
    s = g + t
    lin = layers * s * 2 * (w * (2 * w + 2 * hkv * 64) + w * w + 3 * w * inner)
    attn = layers * 4 * s * s * w
    io = 2 * 768 * w * g * 2 + 2 * 5 * w * t * 2
    return 2 * (lin + attn) + io - latent_tail_savings(s, t, w, inner), s


def scaled_leg(T, dev, world, rank, dist, steps, warmup, size="tiny", clips_per_gpu=4):
    """BASELINE configs[4]: the scaled-up variant that stresses the attention sequence length -- clips of
    32x256x256 with 256 latent tokens (8192 patches, 8448 packed rows per clip; attention is 84 % of the FLOPs at tiny),
    tiny / base / large stacks, forward tokenise + reconstruct through TiTok.tokenize_reconstruct_ (CUDA-graph replay),
    device-timed."""
    from titok_video_b200.config import tiny_config

    shape, t = (32, 256, 256), 256
    torch.manual_seed(42)
    model = T.TiTok(tiny_config(LEVELS, PATCH, size, size)).to(dev).eval()
    gen = torch.Generator().manual_seed(3000 + rank)
    sets = [[(torch.rand((3, *shape), generator=gen) * 2 - 1).to(torch.bfloat16).to(dev) for _ in range(clips_per_gpu)]
            for _ in range(2)]
    tcs = [t] * clips_per_gpu
    fl, s_rows = model_flops(size, shape, t)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(max(warmup, 3)):
            model.tokenize_reconstruct_(sets[i % 2], tcs)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            model.tokenize_reconstruct_(sets[i % 2], tcs)
        e1.record()
        barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / steps
    cps = world * clips_per_gpu / (ms_step * 1e-3)
    tf = cps / world * fl / 1e12
    del model
    torch.cuda.empty_cache()
    return {"workload": f"per GPU {clips_per_gpu} clips 3x32x256x256 bf16, 256 latent tokens each ({s_rows} packed rows per clip), "
                        f"{size} encoder / decoder, forward tokenise + reconstruct",
            "clips_per_s": cps, "latent_tokens_per_s": cps * t, "ms_per_step": ms_step, "gflop_per_clip": fl / 1e9,
            "tflops_per_gpu": tf, "frac_of_tensor_peak": tf / peaks()["bf16_tflops_sustained"]}


def run_reference(args, rank):
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, max(args.warmup, 1), in_process=True)
    line = {"impl": "reference", "metric": "clips/sec encode+decode", "value": r["clips_per_s"], "unit": "clips/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": {"bfloat16": "bf16", "float32": "f32"}.get(r.get("dtype"), "bf16"), "data": "synthetic",
            "config": {"workload": workload_string(args.batch), "clips_per_gpu_per_step": args.batch,
                       "sample": "each timed step is a bounded sample of that workload: " + r["sample"],
                       "note": "CPU arm: the reference has no CPU path of its own (flash-attn / Triton RMSNorm are CUDA-only); "
                               "its unmodified files run on all host threads with stand-ins for those third-party kernels. "
                               "The same-box GPU run of the reference (real flash-attn 2.8.3 / Triton / cuBLAS) is reported "
                               "by the product arm as config.gpu_reference"},
            "cpu_baseline": {"value": r["clips_per_s"], "unit": "clips/s", "cores": r["cores"], "kind": r["kind"],
                             "sample": r["sample"], "shims": r.get("shims")},
            "e2e": {"value": r["clips_per_s"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vq", action="store_true", help="skip the quantizer microbench (BASELINE configs[1]) leg")
    ap.add_argument("--ragged-stream", type=int, default=8,
                    help="steps of the ragged-stream leg (new batch composition every step; 0 = skip)")
    ap.add_argument("--train-batch", default="3,16",
                    help="clips per GPU per step of the training-step leg(s) (BASELINE configs[3]; 3 clips A = tiny.yaml's "
                         "6144-token budget); empty = skip")
    ap.add_argument("--no-scaled", action="store_true", help="skip the scaled-up (32x256x256) leg (BASELINE configs[4])")
    ap.add_argument("--e2e-tokens-only", action="store_true",
                    help="skip the full-result e2e leg (default: `e2e` brings the decoded clips back to the host too; "
                         "`e2e.tokens_only` is the same leg with token indices + per-clip error only)")
    ap.add_argument("--no-gpu-reference", action="store_true",
                    help="skip the same-box GPU run of the unmodified reference (config.gpu_reference)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist

    import titok_video_b200 as T
    from titok_video_b200 import _lib
    from titok_video_b200.config import tiny_config

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pin_here = lambda: numa_local(local_rank, world > 1)  # around every pinned allocation (first touch decides the node)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.batch
    global INPUT_SETS
    if B * 3 * CLIP_A[0] * CLIP_A[1] * CLIP_A[2] * 2 > 126e6:
        INPUT_SETS = 2  # one step's clips already exceed the 126 MB L2
    torch.manual_seed(42)
    model = T.TiTok(tiny_config(LEVELS, PATCH)).to(dev).eval()
    tcs = [TOKENS_A] * B
    flops_clip, s, g = clip_flops()

    # synthetic clips: INPUT_SETS distinct batches so that a step never finds its input in L2
    gen = torch.Generator().manual_seed(1000 + rank)
    # every set is one flat pinned buffer (one H2D copy per step); the clips handed to the model are views into it
    clip_numel = 3 * CLIP_A[0] * CLIP_A[1] * CLIP_A[2]

    def views(flat):
        return [flat[i * clip_numel:(i + 1) * clip_numel].view(3, *CLIP_A) for i in range(B)]

    with pin_here():
        host_flat = [(torch.rand((B * clip_numel,), generator=gen) * 2 - 1).to(torch.bfloat16).pin_memory()
                     for _ in range(INPUT_SETS)]
    dev_sets = [views(hf.to(dev)) for hf in host_flat]
    clip_bytes = 3 * CLIP_A[0] * CLIP_A[1] * CLIP_A[2] * 2

    hist = torch.zeros(model.quantize.codebook_size, dtype=torch.int32, device=dev)

    def step(clips, use_graph=False, with_error=False):
        with torch.no_grad():
            recon, d = model.tokenize_reconstruct_(clips, tcs, use_graph=use_graph, with_error=with_error)
            _lib.call("ttk_hist_u32", T.engine._ptr(d["indices"]), d["indices"].numel(), hist.numel(),
                      T.engine._ptr(hist), T.engine._stream())
        return recon, d

    # ---------------- value: inputs resident in HBM ----------------
    # the objects that exist now (torch, the model, the input sets) move to the collector's permanent generation: a full
    # collection of Python's cyclic GC over them takes 20-40 ms and would otherwise land at random inside the host-bound
    # legs (ragged stream, 3-clip training step); what a serving / training process does once after start-up
    import gc

    gc.collect()
    gc.freeze()
    sampler = ClockSampler(local_rank)  # started ahead of the warm-up: nvidia-smi needs ~0.3 s to deliver its first sample
    sampler.start()
    for i in range(args.warmup):
        step(dev_sets[i % INPUT_SETS])
    barrier()
    timer = KernelTimer()
    _lib.set_profiler(timer)
    launches0 = _lib.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    for i in range(args.steps):
        step(dev_sets[i % INPUT_SETS])
    e1.record()
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    launches = _lib.LAUNCHES - launches0
    _lib.set_profiler(None)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    value = world * B * args.steps / (total_ms * 1e-3)
    ksum = timer.summary()

    # ---------------- the same step with wide ("stress") weights: tokens spread over the codebook, attention rows are peaked and
    # the lazy rescale of the attention kernel (rare at the reference initialiser, whose attention is near-uniform) fires
    stress = None
    if rank == 0:
        import math

        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        gs = torch.Generator().manual_seed(1)
        for k in sorted(sd):
            v = sd[k]
            if v.dim() == 2 and v.shape[0] > 1 and v.shape[1] > 1:
                v.copy_((torch.randn(v.shape, generator=gs) * (2.0 / math.sqrt(v.shape[1]))).to(v.device))
            elif k.endswith(".bias"):
                v.copy_((torch.randn(v.shape, generator=gs) * 0.1).to(v.device))
        torch.manual_seed(42)
        model_s = T.TiTok(tiny_config(LEVELS, PATCH)).to(dev).eval()
        model_s.load_state_dict(sd)
        hist_s = torch.zeros_like(hist)

        def step_s(clips):
            with torch.no_grad():
                _, d = model_s.tokenize_reconstruct_(clips, tcs, use_graph=False)
                _lib.call("ttk_hist_u32", T.engine._ptr(d["indices"]), d["indices"].numel(), hist_s.numel(),
                          T.engine._ptr(hist_s), T.engine._stream())

        for i in range(3):
            step_s(dev_sets[i % INPUT_SETS])
        torch.cuda.synchronize()
        timer_s = KernelTimer()
        _lib.set_profiler(timer_s)
        es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        es0.record()
        n_s = max(4, args.steps // 2)
        for i in range(n_s):
            step_s(dev_sets[i % INPUT_SETS])
        es1.record()
        torch.cuda.synchronize()
        _lib.set_profiler(None)
        ks = timer_s.summary().get("ttk_attn_varlen_fwd")
        stress = {"ms_per_step": es0.elapsed_time(es1) / n_s, "clips_per_s": B * n_s / (es0.elapsed_time(es1) * 1e-3),
                  "attn_avg_launch_ms": (ks[0] / ks[1]) if ks else None,
                  "codebook_usage_percent": float((hist_s > 0).sum().item()) / hist_s.numel() * 100.0,
                  "weights": "every 2-D parameter ~ N(0, (2/sqrt(fan_in))^2), seed 1 (the parity tests' stress initialiser)"}
        del model_s
    if world > 1:
        dist.barrier()

    # ---------------- e2e: pinned host clips -> H2D -> public API -> D2H of indices + reconstructions ----------------
    h2d_stream, d2h_stream = torch.cuda.Stream(), torch.cuda.Stream()
    SLOTS = 2
    main_stream = torch.cuda.current_stream()

    def _flat_of(recon):
        # the reconstructed clips are views of one flat workspace buffer (engine.split_clips)
        base = recon[0]
        n = B * 3 * CLIP_A[0] * CLIP_A[1] * CLIP_A[2]
        return base.reshape(-1).as_strided((n,), (1,), base.storage_offset())

    def e2e_leg(host_bufs, in_dtype, full):
        """Timed region per step: H2D of the step's clips from pinned host memory (side stream), the public API call,
        D2H of the results into pinned host memory (side stream). Returns (device ms, wall ms) per step, max over ranks."""
        in_flat = [torch.empty((B * clip_numel,), dtype=in_dtype, device=dev) for _ in range(SLOTS)]
        in_slots = [views(f) for f in in_flat]
        out_slots = [torch.empty((B * clip_numel,) if full else (1,), dtype=torch.bfloat16, device=dev) for _ in range(SLOTS)]
        idx_slots = [torch.empty((B * TOKENS_A,), dtype=torch.int32, device=dev) for _ in range(SLOTS)]
        err_slots = [torch.empty((B, 2), dtype=torch.float64, device=dev) for _ in range(SLOTS)]
        with pin_here():
            host_out = [torch.empty_like(out_slots[0], device="cpu").pin_memory() for _ in range(SLOTS)]
            host_idx = [torch.empty_like(idx_slots[0], device="cpu").pin_memory() for _ in range(SLOTS)]
            host_err = [torch.empty_like(err_slots[0], device="cpu").pin_memory() for _ in range(SLOTS)]
        ev_in = [torch.cuda.Event() for _ in range(SLOTS)]
        ev_compute = [torch.cuda.Event() for _ in range(SLOTS)]
        ev_out = [torch.cuda.Event() for _ in range(SLOTS)]
        ev_consumed = [torch.cuda.Event() for _ in range(SLOTS)]

        def e2e_step(i):
            sl = i % SLOTS
            with torch.cuda.stream(h2d_stream):
                h2d_stream.wait_event(ev_consumed[sl])  # the slot's previous contents were consumed by compute
                in_flat[sl].copy_(host_bufs[i % len(host_bufs)], non_blocking=True)
                ev_in[sl].record()
            main_stream.wait_event(ev_in[sl])
            main_stream.wait_event(ev_out[sl])  # the slot's previous results have left the device
            # the public API's default: one CUDA-graph replay per step; results = token indices + per-clip error
            recon, d = step(in_slots[sl], use_graph=True, with_error=True)
            ev_consumed[sl].record()
            if full:
                out_slots[sl].copy_(_flat_of(recon), non_blocking=True)
            idx_slots[sl].copy_(d["indices"], non_blocking=True)
            err_slots[sl].copy_(d["clip_error"], non_blocking=True)
            ev_compute[sl].record()
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(ev_compute[sl])
                if full:
                    host_out[sl].copy_(out_slots[sl], non_blocking=True)
                host_idx[sl].copy_(idx_slots[sl], non_blocking=True)
                host_err[sl].copy_(err_slots[sl], non_blocking=True)
                ev_out[sl].record()

        for i in range(3):
            e2e_step(i)
        barrier()
        d2h_stream.synchronize()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.perf_counter()
        ea.record()
        for i in range(args.steps):
            e2e_step(i)
        main_stream.wait_stream(d2h_stream)
        eb.record()
        barrier()
        d2h_stream.synchronize()
        w1 = time.perf_counter()
        t = torch.tensor([max(ea.elapsed_time(eb), 0.0), (w1 - w0) * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()), float(t[1].item())

    full = not args.e2e_tokens_only
    e2e_ms, e2e_wall_ms = e2e_leg(host_flat, torch.bfloat16, full)
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    tok_ms, tok_wall_ms = e2e_leg(host_flat, torch.bfloat16, False) if full else (e2e_ms, e2e_wall_ms)

    # plain pinned-copy ceiling of this box at N ranks: the same 173 MB H2D (and D2H) per step with NO compute, all ranks
    # at once -- what `e2e` cannot beat (shows whether the e2e scaling curve is the platform's host<->device path)
    def copy_probe():
        dbuf = torch.empty((B * clip_numel,), dtype=torch.bfloat16, device=dev)
        with pin_here():
            hout = torch.empty((B * clip_numel,), dtype=torch.bfloat16).pin_memory()
        res = {}
        for name, fn in (("h2d", lambda i: dbuf.copy_(host_flat[i % INPUT_SETS], non_blocking=True)),
                         ("d2h", lambda i: hout.copy_(dbuf, non_blocking=True))):
            for i in range(2):
                fn(i)
            barrier()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            n = 8
            for i in range(n):
                fn(i)
            eb.record()
            barrier()
            t = torch.tensor([ea.elapsed_time(eb) / n], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[name + "_ms_per_step"] = float(t.item())
            res[name + "_gbs_per_gpu"] = B * clip_bytes / (float(t.item()) * 1e-3) / 1e9
        res["note"] = (f"{B * clip_bytes / 1e6:.0f} MB pinned<->device copies, all {world} rank(s) at once, no compute: "
                       "the floor of an e2e step is max(compute, h2d, d2h) since the three overlap")
        return res

    probe = copy_probe()
    # same leg fed with decoded uint8 frames (what the reference's dataset holds before `/255*2-1`, video_dataset.py:111-119):
    # half the PCIe bytes, normalised on the device by ttk_normalize_u8 (bit-identical to the host expression)
    with pin_here():
        host_u8 = [torch.randint(0, 256, (B * clip_numel,), generator=gen, dtype=torch.uint8).pin_memory() for _ in range(INPUT_SETS)]
    u8_ms, u8_wall_ms = e2e_leg(host_u8, torch.uint8, full)
    e2e_u8 = {"value": world * B * args.steps / (u8_ms * 1e-3), "unit": "clips/s", "h2d_bytes_per_step": B * clip_numel,
              "d2h_bytes_per_step": (B * clip_bytes if full else 0) + B * TOKENS_A * 4 + B * 16, "ms_per_step": u8_ms / args.steps,
              "wall_ms_per_step": u8_wall_ms / args.steps,
              "input": "uint8 frames [3,T,H,W] from pinned host memory (the reference's dataset output before normalisation), "
                       "normalised on the device; everything else as `e2e`"}
    del host_u8

    # ---------------- ragged stream: a NEW batch composition every step (what train.py / a tokenisation job feeds) ------
    # shapes and token counts drawn from the sampling ranges of configs/tiny.yaml (tiny.yaml:56-66); every step pays the
    # host planner, the metadata upload and kernel-by-kernel launches (no plan cache hit, no graph replay).
    ragged = None
    if args.ragged_stream > 0 and rank == 0:
        import random

        from titok_video_b200 import engine as _eng

        rnd = random.Random(0)
        n_clips = min(B, 16)
        batches = []
        RW = 8  # untimed compositions first: one-time costs (lazily loaded kernel variants of new shape classes, the
        #         allocator's first blocks, workspace growth) belong to the warm-up, as in every other leg
        for _ in range(args.ragged_stream + RW):
            shp = [(rnd.choice([8, 12, 16]), rnd.choice([128, 136, 144, 152, 160, 168]), rnd.choice([128, 136, 144, 152, 160, 168]))
                   for _ in range(n_clips)]
            tc = [rnd.randint(1, 128) for _ in range(n_clips)]
            batches.append(([(torch.rand((3, *sh), device=dev) * 2 - 1).to(torch.bfloat16) for sh in shp], tc))
        _eng.clear_caches()
        for clips_r, tc_r in batches[:RW]:
            with torch.no_grad():
                model.tokenize_reconstruct_(clips_r, tc_r, use_graph=False)
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for clips_r, tc_r in batches[RW:]:
            with torch.no_grad():
                model.tokenize_reconstruct_(clips_r, tc_r, use_graph=False)
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        # the same batches replayed as CUDA graphs (each composition captured once): their pure kernel time, the floor
        # that host planning + eager launching is measured against
        with torch.no_grad():
            for clips_r, tc_r in batches[RW:]:
                model.tokenize_reconstruct_(clips_r, tc_r, use_graph=True)
            torch.cuda.synchronize()
            eg0, eg1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eg0.record()
            for clips_r, tc_r in batches[RW:]:
                model.tokenize_reconstruct_(clips_r, tc_r, use_graph=True)
            eg1.record()
            torch.cuda.synchronize()
        kernel_ms = eg0.elapsed_time(eg1) / args.ragged_stream
        # shape-bucketed graph replay (TiTok.tokenize_reconstruct_bucketed_): 40 warm-up batches populate the buckets, then 40
        # batches that were NEVER seen are timed (wall clock; a bucket met for the first time pays its capture inside)
        bucketed = None
        try:
            nb = 40
            more = []
            for _ in range(2 * nb):
                shp = [(rnd.choice([8, 12, 16]), rnd.choice([128, 136, 144, 152, 160, 168]), rnd.choice([128, 136, 144, 152, 160, 168]))
                       for _ in range(n_clips)]
                more.append((shp, [rnd.randint(1, 128) for _ in range(n_clips)]))
            pool = [(torch.rand((3, 16, 168, 168), device=dev) * 2 - 1).to(torch.bfloat16) for _ in range(n_clips)]

            def batch_of(shp):
                return [pool[i][:, :s[0], :s[1], :s[2]].contiguous() for i, s in enumerate(shp)]

            _eng._BUCKET_CACHE.clear()
            with torch.no_grad():
                for shp, tc_r in more[:nb]:
                    model.tokenize_reconstruct_bucketed_(batch_of(shp), tc_r)
                torch.cuda.synchronize()
                graphs0 = sum(len(bp.graphs) for bp in _eng._BUCKET_CACHE.values())
                timed = [(batch_of(shp), tc_r) for shp, tc_r in more[nb:]]
                torch.cuda.synchronize()
                wb0 = time.perf_counter()
                for clips_r, tc_r in timed:
                    model.tokenize_reconstruct_bucketed_(clips_r, tc_r)
                torch.cuda.synchronize()
                wb1 = time.perf_counter()
                _eng._PLAN_CACHE.clear()  # (evicting plans that own captured graphs from the leg above would be timed otherwise)
                torch.cuda.synchronize()
                we0 = time.perf_counter()
                for clips_r, tc_r in timed:
                    model.tokenize_reconstruct_(clips_r, tc_r, use_graph=False)
                torch.cuda.synchronize()
                we1 = time.perf_counter()
            graphs1 = sum(len(bp.graphs) for bp in _eng._BUCKET_CACHE.values())
            bucketed = {"ms_per_step_wall": 1e3 * (wb1 - wb0) / nb, "eager_ms_per_step_wall_same_batches": 1e3 * (we1 - we0) / nb,
                        "steps": nb, "clips_per_step": n_clips, "clips_per_s": n_clips * nb / (wb1 - wb0),
                        "graphs_after_warmup": graphs0, "graphs_captured_during_timing": graphs1 - graphs0,
                        "bucket_steps": [_eng.BucketPlan.G_STEP, _eng.BucketPlan.T_STEP, _eng.BucketPlan.B_STEP],
                        "note": "40 never-seen batch compositions through graphs captured per shape bucket (patches / tokens / "
                                "clips rounded up); per step: O(B) host planning, one H2D of the descriptors, one cat, one graph launch"}
        except Exception as e:  # keep the bench line alive
            bucketed = {"error": repr(e)[:300]}
        ragged = {"clips_per_s": n_clips * args.ragged_stream / (w1 - w0), "clips_per_step": n_clips, "steps": args.ragged_stream,
                  "bucketed": bucketed,
                  "ms_per_step_wall": 1e3 * (w1 - w0) / args.ragged_stream, "kernel_ms_per_step": kernel_ms,
                  "wall_over_kernel": 1e3 * (w1 - w0) / args.ragged_stream / kernel_ms,
                  "note": "new shapes / token counts every step (tiny.yaml sampling ranges): includes host planning, metadata "
                          "upload and eager launches; wall clock around a synchronised loop"}

    # ---------------- training step (BASELINE configs[3]) ----------------
    train = None
    if args.train_batch:
        train = {}
        for tb in [int(v) for v in args.train_batch.split(",") if v]:
            train[f"batch{tb}"] = train_leg(T, _lib, dev, world, rank, dist, tb, max(4, args.steps // 2), args.warmup)
            train[f"batch{tb}_graph"] = train_leg(T, _lib, dev, world, rank, dist, tb, max(4, args.steps // 2), args.warmup,
                                                  graphed=True)
        if rank == 0:
            train["ragged"] = train_ragged_leg(T, _lib, dev, rank, 8)
        if rank == 0 and world == 1:
            train["gan"] = gan_leg(T, _lib, dev, 3, max(4, args.steps // 2), args.warmup)

    # ---------------- scaled-up variant (BASELINE configs[4]) ----------------
    scaled = None
    if not args.no_scaled:
        scaled = scaled_leg(T, dev, world, rank, dist, max(4, args.steps // 2), args.warmup)
        # the wider stacks of the size table (utils.py:8-23): widths 768 / 1024 take the unfused residual path
        scaled["base"] = scaled_leg(T, dev, world, rank, dist, 4, 3, size="base", clips_per_gpu=2)
        scaled["large"] = scaled_leg(T, dev, world, rank, dist, 3, 3, size="large", clips_per_gpu=1)

    # ---------------- codebook usage over the whole job (the only data-path collective) ----------------
    if world > 1:
        dist.all_reduce(hist)
    usage = float((hist > 0).sum().item()) / hist.numel() * 100.0

    # ---------------- roofline of the dominant kernel ----------------
    pk = peaks()
    tot_k = sum(v[0] for v in ksum.values()) or 1.0
    top = max(ksum.items(), key=lambda kv: kv[1][0])[0] if ksum else None
    kernels = {}
    for name, (kms, n) in sorted(ksum.items(), key=lambda kv: -kv[1][0]):
        fl = kernel_flops_per_launch(name, B, s, g, TOKENS_A)
        kernels[name] = {"ms_per_step": kms / args.steps, "launches_per_step": n / args.steps, "share": kms / tot_k,
                         "tflops": (fl / (kms / n * 1e-3) / 1e12) if fl else None}
    roofline = None
    if top is not None:
        fl = kernel_flops_per_launch(top, B, s, g, TOKENS_A)
        avg_ms = ksum[top][0] / ksum[top][1]
        # the timed region is ~0.1 s at full clocks (see `clocks`), not a seconds-long power-capped loop: the burst figure
        # is the like-for-like denominator; the fraction of the sustained figure is given beside it
        peak = pk["bf16_tflops"]
        ach = fl / (avg_ms * 1e-3) / 1e12 if fl else None
        roofline = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                    "frac": (ach / peak) if ach else None, "traffic": ncu_traffic(top, B), "peak_source": pk["source"] + " (burst bf16)",
                    "frac_of_sustained": (ach / pk["bf16_tflops_sustained"]) if ach else None,
                    "avg_launch_ms": avg_ms, "share_of_step": ksum[top][0] / tot_k,
                    "mufu_bound_tflops": 16 * 148 * 1.965e9 * 256 / 1e12,
                    "note": "head dim 64: 256 FLOP per exponential; the XU pipe's 16 ex2/clk/SM caps the kernel at 1191 TFLOP/s "
                            "(75 % of the burst tensor peak; 1361 with the 12.5 % of the exponentials this kernel evaluates on the "
                            "FMA pipes) before any other limit. One launch per call: the score bound of the bounded-score "
                            "softmax comes from key norms that the qkv GEMM epilogue leaves behind"}
    # FLOPs the GPU executes per clip: the reference's algorithm minus the rows the last encoder layer skips
    exec_clip = flops_clip - latent_tail_savings(s, TOKENS_A, WIDTH, INNER)
    whole = {"tflops": value * exec_clip / 1e12, "frac_of_tensor_peak": value * exec_clip / 1e12 / pk["bf16_tflops_sustained"],
             "gflop_per_clip_executed": exec_clip / 1e9, "gflop_per_clip_reference_algorithm": flops_clip / 1e9,
             "note": "executed FLOPs: the encoder's last layer runs attention queries, out_proj and the GEGLU block on the latent "
                     "rows only (the head reads nothing else; results bit-identical to running every row)"}

    vq = None
    if rank == 0 and not args.no_vq:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import vq_bench

        vq = vq_bench.main(quick=True, quiet=True)
    if world > 1:
        dist.barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(steps=6, warmup=1)
        cpu = {"value": r["clips_per_s"], "unit": "clips/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
               "shims": r.get("shims")}

    # same-box GPU run of the UNMODIFIED reference (flash-attn 2.8.3 varlen + Triton RMSNorm + cuBLAS), separate process
    gpu_ref = None
    if rank == 0 and world == 1 and not args.no_gpu_reference:
        if reference_available():
            gpu_ref = _runner(["bench", "--mode", "gpu", "--batch", str(B), "--steps", str(max(4, args.steps // 2)), "--warmup", "3"])
            if "unavailable" not in gpu_ref:
                gpu_ref["ours_over_reference"] = value / gpu_ref["clips_per_s"]
                top_attn = ksum.get("ttk_attn_varlen_fwd")
                if top_attn and gpu_ref.get("attn_alone_us_per_launch"):
                    gpu_ref["ours_attn_us_per_launch"] = 1e3 * top_attn[0] / top_attn[1]
        else:
            gpu_ref = {"unavailable": "no reference copy under baseline/_ref (scripts/vendor_reference.sh not run)"}

    if rank == 0:
        def vq_line(r):
            return {"K": r["K"], "D": r["D"], "N": r["N"], "ms": r["ms"], "tflops": r["tflops_algorithmic"],
                    "frac_burst": r["frac_of_tensor_peak"], "mma_frac_burst": r["mma_frac_of_tensor_peak"]}

        if roofline is not None and vq:
            roofline["vq"] = [vq_line(r) for r in vq if r["kernel"] == "ttk_vq_argmin"]
            roofline["fsq"] = [{"dtype": r["dtype"], "N": r["N"], "ms": r["ms"], "gbs": r["gbs"], "frac_hbm": r["frac_of_hbm_peak"]}
                               for r in vq if r["kernel"] == "ttk_fsq_fwd"]
            roofline["vq_peak"] = {"tflops": pk["bf16_tflops"], "source": pk["source"] + " (burst bf16: kernel timed alone)"}
        if roofline is not None and stress is not None:
            roofline["stress_init"] = stress
        cfg = {"workload": workload_string(B), "clips_per_gpu_per_step": B,
               "global_clips_per_step": world * B, "latent_tokens_per_s": value * TOKENS_A,
               "parallelism": f"clip-sharded x{world}, no data-path collective",
               "l2": f"inputs rotate over {INPUT_SETS} sets ({INPUT_SETS * B * clip_bytes / 1e6:.0f} MB) and the per-step "
                     f"activation working set exceeds the 126 MB L2",
               "weights": "random init, seed 42 (reference initialiser)", "codebook_usage_percent": usage,
               "host_affinity": numa_local.desc,
               "whole_step_tflops": whole["tflops"], "whole_step_frac_of_tensor_peak": whole["frac_of_tensor_peak"],
               "gpu_reference": gpu_ref}
        if train:
            for k, v in train.items():
                cfg["train_step_" + k] = {kk: vv for kk, vv in v.items() if kk not in ("kernels", "what", "note")}
        if scaled:
            cfg["scaled_config"] = scaled
        if ragged:
            cfg["ragged_stream"] = {k: v for k, v in ragged.items() if k != "note"}
        d2h_small = B * TOKENS_A * 4 + B * 16
        line = {
            "metric": "clips/sec encode+decode", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": cfg,
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": B * clip_bytes,
                    "d2h_bytes_per_step": (B * clip_bytes if full else 0) + d2h_small,
                    "result": ("token indices + per-clip (L1, squared) reconstruction error" +
                               (" + the decoded clips (full reconstructions)" if full else
                                "; reconstructions stay on the device (--e2e-tokens-only)")),
                    "ms_per_step": e2e_ms / args.steps,
                    "wall_ms_per_step": e2e_wall_ms / args.steps,
                    "tokens_only": {"value": world * B * args.steps / (tok_ms * 1e-3), "ms_per_step": tok_ms / args.steps,
                                    "d2h_bytes_per_step": d2h_small,
                                    "result": "token indices + per-clip error only (a tokenisation job; the decoded clips stay on the device)"},
                    "copy_probe": probe,
                    "u8": e2e_u8,
                    "api": "TiTok.tokenize_reconstruct_(clips, token_counts) from pinned host clips, 2-slot pipeline, CUDA-graph replay "
                           "(the value leg launches the same kernels one by one so that each can be timed with CUDA events)"},
            "gpu_launches": launches, "roofline": roofline, "whole_step": whole, "kernels": kernels, "clocks": clocks,
            "cpu_baseline": cpu, "quantizer_microbench": vq, "ragged_stream": ragged, "train_step": train, "scaled_config": scaled, "e2e_u8": e2e_u8,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
