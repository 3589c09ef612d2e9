"""NCCL legs of the multi-GPU path on real GPUs (world size 2, one process per GPU; skipped on a 1-GPU box):
GradientAllReducer on the drop-in TiTok -- the CUDA backward's flat per-stack gradient buffers are all-reduced in place
over NVLink -- against single-process gradients of the same clips, and the fused small all-reduce of the step
statistics. The host-side logic of the same classes is covered on CPU with gloo (tests/test_dist_gloo.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu
needs_2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")

SHAPES = [[(8, 64, 48), (4, 16, 24)], [(8, 32, 32), (12, 40, 24)]]
TCS = [[16, 3], [8, 5]]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _step(model, clips, tcs):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        recon, d = model(clips, tcs)
    loss = torch.stack([(r_.float() - c.float()).abs().mean() for c, r_ in zip(clips, recon)]).mean()
    loss.backward()
    return loss.detach(), d["indices"]


def _worker(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from conftest import build_model
    from oracle import titok_oracle as O
    from titok_video_b200 import dist as D

    model = build_model(True).cuda().train()
    red = D.GradientAllReducer(model)
    clips = [c.cuda() for c in O.make_clips(SHAPES[rank], 3 + rank)]
    paths = []
    for it in range(2):  # second iteration: hooks re-armed, buffers re-pointed
        model.zero_grad(set_to_none=True)
        loss, idx = _step(model, clips, TCS[rank])
        red.finish()
        paths.append(red.last_path)
    hist = torch.bincount(idx.long(), minlength=4375).to(torch.int32)
    h, means = D.allreduce_step_stats(hist, {"l1": loss * len(clips)}, {"l1": len(clips)})
    torch.cuda.synchronize()
    torch.save({"grads": {k: p.grad.detach().float().cpu() for k, p in model.named_parameters()}, "paths": paths,
                "hist_total": int(h.sum()), "l1": float(means["l1"]), "loss": float(loss)},
               os.path.join(out_dir, f"n{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@needs_2
def test_nccl_gradient_allreduce_matches_single_process_mean(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(tmp_path / f"n{r}.pt") for r in range(world)]
    from conftest import build_model
    from oracle import titok_oracle as O

    model = build_model(True).cuda().train()
    want, losses = None, []
    for r in range(world):
        model.zero_grad(set_to_none=True)
        clips = [c.cuda() for c in O.make_clips(SHAPES[r], 3 + r)]
        loss, _ = _step(model, clips, TCS[r])
        losses.append(float(loss))
        g = {k: p.grad.detach().float().cpu().clone() for k, p in model.named_parameters()}
        want = g if want is None else {k: want[k] + g[k] for k in g}
    for r in range(world):
        assert got[r]["paths"] == ["inplace", "inplace"], got[r]["paths"]  # the flat buffers were reduced directly
        assert got[r]["hist_total"] == sum(sum(t) for t in TCS)
        assert abs(got[r]["l1"] - sum(l * len(s) for l, s in zip(losses, SHAPES)) / sum(len(s) for s in SHAPES)) < 1e-5
        for k in want:
            w = want[k] / world
            d = (got[r]["grads"][k] - w).norm() / (w.norm() + 1e-30)
            # split-K weight gradients use float atomics: equal up to summation order
            assert d < 1e-4, f"rank {r} {k}: rel {float(d):.3g}"
    for k in want:
        assert torch.equal(got[0]["grads"][k], got[1]["grads"][k]), k  # both ranks hold the same averaged gradient
