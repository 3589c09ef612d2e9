#!/usr/bin/env python
"""Host-side profile of the ragged inference stream (new batch composition every step)."""
import cProfile, io, os, pstats, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import bench, titok_video_b200 as T
from titok_video_b200 import engine
from titok_video_b200.config import tiny_config
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
torch.manual_seed(42)
model = T.TiTok(tiny_config(bench.LEVELS, bench.PATCH)).to(dev).eval()
rnd = random.Random(0)
batches = []
for _ in range(30):
    shp = [(rnd.choice([8, 12, 16]), rnd.choice([128, 136, 144, 152, 160, 168]), rnd.choice([128, 136, 144, 152, 160, 168])) for _ in range(16)]
    tc = [rnd.randint(1, 128) for _ in range(16)]
    batches.append(([(torch.rand((3, *sh), device=dev) * 2 - 1).to(torch.bfloat16) for sh in shp], tc))
def run(bs):
    with torch.no_grad():
        for c, t in bs:
            model.tokenize_reconstruct_(c, t, use_graph=False)
run(batches[:4]); torch.cuda.synchronize()
t0 = time.perf_counter(); run(batches[4:14]); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host issue {1e3*(t1-t0)/10:.2f} ms/step, with sync {1e3*(t2-t0)/10:.2f} ms/step")
pr = cProfile.Profile(); pr.enable(); run(batches[14:24]); pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
