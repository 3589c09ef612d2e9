#!/usr/bin/env python
"""A few generator training steps (bench.py's training-step leg on its own): the command the training ncu captures run.
    python scripts/train_step.py [--batch 16] [--steps 3]"""
import argparse, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import titok_video_b200 as T  # noqa: E402
from titok_video_b200 import _lib  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--graph", action="store_true", help="also run the leg with forward + loss + backward replayed as one CUDA graph")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    r = bench.train_leg(T, _lib, dev, 1, 0, None, a.batch, a.steps, 3)
    r.pop("what")
    print(json.dumps(r))
    if a.graph:
        r = bench.train_leg(T, _lib, dev, 1, 0, None, a.batch, a.steps, 3, graphed=True)
        r.pop("what")
        print(json.dumps(r))
