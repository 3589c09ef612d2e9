#!/usr/bin/env python
"""Host-side profile (cProfile) of the training step: where the Python / ctypes / autograd time goes."""
import cProfile, os, pstats, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import bench, titok_video_b200 as T
from titok_video_b200 import _lib
from titok_video_b200.config import tiny_config

dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
torch.manual_seed(42)
model = T.TiTok(tiny_config(bench.LEVELS, bench.PATCH)).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.5, 0.96), weight_decay=1e-4, fused=True)
clips = [(torch.rand((3, *bench.CLIP_A)) * 2 - 1).to(torch.bfloat16).to(dev) for _ in range(B)]
tcs = [bench.TOKENS_A] * B
def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        recon, d = model(clips, tcs)
    loss = torch.stack([(r_.float() - c.float()).abs().mean() for c, r_ in zip(clips, recon)]).mean()
    loss.backward(); opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(10): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host issue time {1e3*(t1-t0)/10:.2f} ms/step, with final sync {1e3*(t2-t0)/10:.2f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:5000])
