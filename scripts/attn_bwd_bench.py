#!/usr/bin/env python
"""Micro-benchmark of the attention backward kernels at the bench shape (clips A, tiny heads): python scripts/attn_bwd_bench.py [clips]"""
import ctypes, math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from titok_video_b200 import _lib
from titok_video_b200.plan import attn_bwd_work_lists, attn_work_list

def P(t): return ctypes.c_void_p(t.data_ptr())
def main(B=16, s=1892, hq=4, hkv=2, reps=10):
    dev = "cuda"; w, g = hq * 64, hkv * 64; M = B * s; ld = 2 * w + 2 * g
    torch.manual_seed(0)
    qkv = torch.randn(M, ld, device=dev).bfloat16(); dOut = torch.randn(M, w, device=dev).bfloat16()
    starts = [i * s for i in range(B)]; lens = [s] * B
    work = torch.from_numpy(attn_work_list(starts, lens, hq, hkv)).to(dev)
    wk = [torch.from_numpy(a).to(dev) for a in attn_bwd_work_lists(starts, lens, hq, hkv)]
    out = torch.empty(M, w, device=dev, dtype=torch.bfloat16); o = torch.empty_like(out); dO = torch.empty_like(out)
    lse = torch.empty(hq, M, device=dev); delta = torch.empty(hq, M, device=dev); dqkv = torch.empty_like(qkv)
    rope = torch.zeros(M, 60, device=dev); rope[:, 0::2] = 1
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    def fwd(): _lib.call("ttk_attn_varlen_fwd_train", P(qkv), ld, M, w, g, P(work), work.shape[0], 0.125, P(out), w, P(o), P(lse), ctypes.c_void_p(0), st)
    def prep(): _lib.call("ttk_attn_bwd_prep", P(dOut), w, P(o), w, P(qkv), ld, M, w, P(dO), w, P(dqkv), ld, P(delta), st)
    def bw(name, k): _lib.call(name, P(qkv), ld, P(dO), w, M, w, g, P(k), k.shape[0], P(lse), P(delta), P(rope), 0.125, P(dqkv), ld, st)
    fl = B * 4.0 * s * s * w  # forward FLOPs
    res = {}
    for name, fn, mult in [("fwd", fwd, 1.0), ("prep", prep, 0), ("dkv", lambda: bw("ttk_attn_bwd_dkv", wk[0]), 2.0),
                           ("dq", lambda: bw("ttk_attn_bwd_dq", wk[1]), 1.5)]:
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[name] = ms
        print(f"{name:5s} {ms*1e3:8.1f} us   {mult*fl/ms/1e9:7.1f} TFLOP/s (useful MMA flops: fwd 4s^2w, dkv 2x, dq 1.5x)")
    return res
if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 16)
