#!/usr/bin/env python
"""Host-side profile of the eager ragged path after the bucketed leg (the order bench.py runs them in)."""
import cProfile, io, os, pstats, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import torch
import bench, titok_video_b200 as T
from titok_video_b200 import engine
from titok_video_b200.config import tiny_config
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
torch.manual_seed(42)
model = T.TiTok(tiny_config(bench.LEVELS, bench.PATCH)).to(dev).eval()
rnd = random.Random(0)
n_clips, nb = 16, 40
more = []
for _ in range(2 * nb):
    shp = [(rnd.choice([8, 12, 16]), rnd.choice([128, 136, 144, 152, 160, 168]), rnd.choice([128, 136, 144, 152, 160, 168])) for _ in range(n_clips)]
    more.append((shp, [rnd.randint(1, 128) for _ in range(n_clips)]))
pool = [(torch.rand((3, 16, 168, 168), device=dev) * 2 - 1).to(torch.bfloat16) for _ in range(n_clips)]
batch_of = lambda shp: [pool[i][:, :s[0], :s[1], :s[2]].contiguous() for i, s in enumerate(shp)]
tag = f"tail={int(engine.LATENT_TAIL)}"
with torch.no_grad():
    for shp, tc in more[:nb]:
        model.tokenize_reconstruct_bucketed_(batch_of(shp), tc)
    torch.cuda.synchronize()
    timed = [(batch_of(shp), tc) for shp, tc in more[nb:]]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for c, tc in timed:
        model.tokenize_reconstruct_bucketed_(c, tc)
    torch.cuda.synchronize()
    print(f"[{tag}] bucketed {1e3 * (time.perf_counter() - t0) / nb:.3f} ms/step", flush=True)
    engine._PLAN_CACHE.clear()
    torch.cuda.synchronize()
    g0 = engine.arena_generation()
    per = []
    for c, tc in timed[:20]:
        t0 = time.perf_counter()
        model.tokenize_reconstruct_(c, tc, use_graph=False)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        per.append((1e3 * (t1 - t0), 1e3 * (time.perf_counter() - t0)))
    print(f"[{tag}] eager per step (issue, with sync) ms: " + " ".join(f"{a:.2f}/{b:.2f}" for a, b in per), flush=True)
    print(f"[{tag}] arena generations during the eager steps: {engine.arena_generation() - g0}", flush=True)
    pr = cProfile.Profile(); pr.enable()
    for c, tc in timed[20:]:
        model.tokenize_reconstruct_(c, tc, use_graph=False)
    pr.disable(); torch.cuda.synchronize()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(12); print(s.getvalue()[:2600])
