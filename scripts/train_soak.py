#!/usr/bin/env python
"""Soak: N ragged training steps (new batch composition every step); prints allocated / reserved device memory over time."""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import bench, titok_video_b200 as T
from titok_video_b200.config import tiny_config
from titok_video_b200.data import dynamic_batches

dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
torch.manual_seed(42)
model = T.TiTok(tiny_config(bench.LEVELS, bench.PATCH)).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.5, 0.96), weight_decay=1e-4, fused=True)
rnd = random.Random(1)
def samples():
    while True:
        shp = (rnd.choice([8, 12, 16]), rnd.choice([128, 136, 144, 152, 160, 168]), rnd.choice([128, 136, 144, 152, 160, 168]))
        yield {"video": (torch.rand((3, *shp), device=dev) * 2 - 1).to(torch.bfloat16)}
it = dynamic_batches(samples(), list(bench.PATCH), [1, 128], [16, 168, 168], 6144, randrange=rnd.randrange)
losses = []
for i in range(N):
    b = next(it)
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        recon, d = model(b["video"], b["token_counts"])
    loss = torch.stack([(r_.float() - c.float()).abs().mean() for c, r_ in zip(b["video"], recon)]).mean()
    loss.backward(); opt.step()
    losses.append(loss.detach())
    if i % 40 == 39 or i == N - 1:
        torch.cuda.synchronize()
        print(f"step {i + 1}: loss {float(torch.stack(losses[-40:]).mean()):.4f} allocated {torch.cuda.memory_allocated() / 2**20:.0f} MiB "
              f"reserved {torch.cuda.memory_reserved() / 2**20:.0f} MiB", flush=True)
assert all(torch.isfinite(l) for l in losses)
