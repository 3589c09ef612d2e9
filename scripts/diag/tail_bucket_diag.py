"""Diagnostic (GPU box): bucketed / eager paths with the workspace arenas poisoned (0xFF bytes = NaN in bf16 / fp32), so that
any read of memory the launch sequence did not write itself shows up. Usage: python scripts/diag/tail_bucket_diag.py <scenario>"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
import torch

from conftest import build_model
from titok_video_b200 import engine

scen = sys.argv[1]
model = build_model(True).cuda().eval()
g = torch.Generator().manual_seed(12)
comps = {
    "small": ([(8, 64, 48), (4, 16, 24), (8, 32, 32)], [16, 3, 8]),
    "big": ([(16, 168, 168), (16, 168, 168)], [128, 64]),
    "mid": ([(8, 64, 64), (8, 96, 64), (4, 16, 16)], [200, 37, 128]),
}


def poison():
    torch.cuda.synchronize()
    for t in engine._ARENA.values():
        t.fill_(0xFF)
    torch.cuda.synchronize()


def report(tag, idx, ref, tcs):
    idx, ref = idx.cpu(), ref.cpu()
    off = 0
    per = []
    for t in tcs:
        per.append(int((idx[off:off + t] != ref[off:off + t]).sum()))
        off += t
    print(f"[{scen}] {tag}: mismatching indices per clip {per} of {tcs}; zeros {int((idx == 0).sum())}/{idx.numel()}", flush=True)


for name in sys.argv[2:]:
    shapes, tcs = comps[name]
    clips = [(torch.rand((3, *s), generator=g) * 2 - 1).to(torch.bfloat16).cuda() for s in shapes]
    with torch.no_grad():
        if scen.startswith("eager"):
            _, d = model.tokenize_reconstruct_(clips, tcs, use_graph=False)
            ref = d["indices"].clone()
            z_ref = model.encoder(clips, tcs).clone()
            poison()
            _, d = model.tokenize_reconstruct_(clips, tcs, use_graph=False)
            report(f"{name} eager after poison", d["indices"], ref, tcs)
            poison()
            z = model.encoder(clips, tcs)
            print(f"[{scen}] {name} encoder z equal after poison: {torch.equal(z, z_ref)}; nan {int(torch.isnan(z.float()).sum())}", flush=True)
            poison()
            _, d = model.tokenize_reconstruct_(clips, tcs, use_graph=True)
            poison()
            _, d = model.tokenize_reconstruct_(clips, tcs, use_graph=True)
            report(f"{name} graph replay after poison", d["indices"], ref, tcs)
        else:
            _, d = model.tokenize_reconstruct_(clips, tcs, use_graph=False)
            ref = d["indices"].clone()
            _, d = model.tokenize_reconstruct_bucketed_(clips, tcs)
            report(f"{name} bucketed first call", d["indices"], ref, tcs)
            _, d = model.tokenize_reconstruct_bucketed_(clips, tcs)
            report(f"{name} bucketed second call", d["indices"], ref, tcs)
            poison()
            _, d = model.tokenize_reconstruct_bucketed_(clips, tcs)
            report(f"{name} bucketed after poison", d["indices"], ref, tcs)
            rec, d = model.tokenize_reconstruct_bucketed_(clips, tcs)
            bad = [int(torch.isnan(r.float()).sum()) for r in rec]
            report(f"{name} bucketed again", d["indices"], ref, tcs)
            print(f"[{scen}] {name} recon NaNs per clip {bad}", flush=True)
