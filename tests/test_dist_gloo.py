"""world_size-2 `gloo` tests of the multi-GPU plumbing on CPU: clip sharding is a deterministic partition, the
histogram all-reduce equals the single-process histogram, gathered indices come back in clip order. The per-rank
"tokeniser" here is the CPU oracle's FSQ (the CUDA forward needs a GPU); the property under test -- a sharded job
produces exactly the single-process result on the concatenated batch -- does not depend on which it is."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

LEVELS = [7, 5, 5, 5, 5]
SHAPES = [(16, 168, 168), (8, 128, 128), (8, 168, 128), (16, 128, 128), (12, 136, 152)]
TCS = [128, 64, 17, 1, 90]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _z_for_clip(i):
    g = torch.Generator().manual_seed(100 + i)
    return (torch.randn((TCS[i], 5), generator=g) * 2).to(torch.bfloat16)


def _worker(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import titok_oracle as O
    from titok_video_b200 import dist as D

    owned = D.shard_batch(SHAPES, TCS)
    local_idx = [O.fsq_forward(_z_for_clip(i).float(), LEVELS)[1] for i in owned]
    counts = torch.zeros(4375, dtype=torch.int64)
    for t in local_idx:
        counts += torch.bincount(t.long(), minlength=4375)
    D.allreduce_counts(counts)
    gathered = D.gather_indices(local_idx, owned, len(SHAPES), dst=0)
    torch.save({"owned": owned, "counts": counts, "gathered": gathered}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_job_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    # partition
    assert sorted(res[0]["owned"] + res[1]["owned"]) == list(range(len(SHAPES)))
    assert not set(res[0]["owned"]) & set(res[1]["owned"])
    # single-process result on the concatenated batch
    from oracle import titok_oracle as O

    ref = [O.fsq_forward(_z_for_clip(i).float(), LEVELS)[1] for i in range(len(SHAPES))]
    ref_counts = torch.bincount(torch.cat(ref).long(), minlength=4375)
    for r in range(world):
        assert torch.equal(res[r]["counts"], ref_counts)
    assert res[1]["gathered"] is None
    for a, b in zip(res[0]["gathered"], ref):
        assert torch.equal(a, b)


def test_sharding_is_balanced_and_rank_independent():
    from titok_video_b200.plan import clip_cost, shard_clips

    costs = [clip_cost(s, t, (4, 8, 8), 256, 4) for s, t in zip(SHAPES * 4, TCS * 4)]
    parts = shard_clips(costs, 8)
    assert sorted(i for p in parts for i in p) == list(range(20))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) <= 1.5 * (sum(costs) / 8)


def test_single_process_helpers_are_noops():
    from titok_video_b200 import dist as D

    assert D.shard_batch(SHAPES, TCS) == list(range(len(SHAPES)))
    c = torch.arange(5)
    assert D.allreduce_counts(c) is c
    out = D.gather_indices([torch.tensor([1, 2]), torch.tensor([3])], [1, 0], 2)
    assert out[0].tolist() == [3] and out[1].tolist() == [1, 2]


def _grad_worker(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from titok_video_b200 import dist as D

    torch.manual_seed(0)
    net = torch.nn.ModuleDict({"encoder": torch.nn.Linear(6, 4), "decoder": torch.nn.Linear(4, 3)})
    red = D.GradientAllReducer(net)
    for step in range(2):  # hooks re-arm after finish()
        net.zero_grad(set_to_none=True)
        x = torch.full((5, 6), float(rank + 1 + step))
        net["decoder"](net["encoder"](x)).square().sum().backward()
        red.finish()
    torch.save({k: p.grad.clone() for k, p in net.named_parameters()}, os.path.join(out_dir, f"g{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreducer_averages_per_stack_buckets(tmp_path):
    """world_size 2, gloo: after finish() every rank holds the mean of the per-rank gradients (DDP semantics)."""
    world = 2
    mp.spawn(_grad_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(tmp_path / f"g{r}.pt") for r in range(world)]
    torch.manual_seed(0)
    net = torch.nn.ModuleDict({"encoder": torch.nn.Linear(6, 4), "decoder": torch.nn.Linear(4, 3)})
    want = None
    for r in range(world):
        net.zero_grad(set_to_none=True)
        x = torch.full((5, 6), float(r + 1 + 1))
        net["decoder"](net["encoder"](x)).square().sum().backward()
        g = {k: p.grad.clone() for k, p in net.named_parameters()}
        want = g if want is None else {k: want[k] + g[k] for k in g}
    for k in want:
        for r in range(world):
            assert torch.allclose(got[r][k], want[k] / world, rtol=1e-5, atol=1e-6), k


def _accum_worker(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from titok_video_b200 import dist as D

    torch.manual_seed(0)
    net = torch.nn.ModuleDict({"encoder": torch.nn.Linear(6, 4), "decoder": torch.nn.Linear(4, 3)})
    red = D.GradientAllReducer(net)
    net.zero_grad(set_to_none=True)
    with red.no_sync():  # first micro-batch: accumulate only
        net["decoder"](net["encoder"](torch.full((5, 6), float(rank + 1)))).square().sum().backward()
    net["decoder"](net["encoder"](torch.full((5, 6), float(rank + 3)))).square().sum().backward()
    red.finish()
    torch.save({k: p.grad.clone() for k, p in net.named_parameters()}, os.path.join(out_dir, f"a{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreducer_no_sync_accumulates_before_reducing(tmp_path):
    """Two micro-batches per rank, the first under no_sync(): every rank ends with the mean over ranks of the SUM of its
    two micro-batch gradients."""
    world = 2
    mp.spawn(_accum_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(tmp_path / f"a{r}.pt") for r in range(world)]
    torch.manual_seed(0)
    net = torch.nn.ModuleDict({"encoder": torch.nn.Linear(6, 4), "decoder": torch.nn.Linear(4, 3)})
    want = None
    for r in range(world):
        net.zero_grad(set_to_none=True)
        for v in (r + 1, r + 3):
            net["decoder"](net["encoder"](torch.full((5, 6), float(v)))).square().sum().backward()
        g = {k: p.grad.clone() for k, p in net.named_parameters()}
        want = g if want is None else {k: want[k] + g[k] for k in g}
    for k in want:
        for r in range(world):
            assert torch.allclose(got[r][k], want[k] / world, rtol=1e-5, atol=1e-6), k


def _stats_worker(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from titok_video_b200 import dist as D

    hist = torch.tensor([1, 0, 2 + rank, 0, 5 * rank], dtype=torch.int32)
    h, means = D.allreduce_step_stats(hist, {"l1": torch.tensor(3.0 * (rank + 1)), "psnr": 10.0 + rank},
                                      {"l1": 2 + rank, "psnr": torch.tensor(1)})
    torch.save({"hist": h, "l1": means["l1"], "psnr": means["psnr"]}, os.path.join(out_dir, f"s{rank}.pt"))
    # a second backward before finish() is an error, not a silent overwrite
    net = torch.nn.ModuleDict({"encoder": torch.nn.Linear(3, 2)})
    red = D.GradientAllReducer(net)
    net["encoder"](torch.ones(1, 3)).sum().backward()
    try:
        net["encoder"](torch.ones(1, 3)).sum().backward()
        raised = False
    except RuntimeError:
        raised = True
    red.finish()
    torch.save({"raised": raised, "path": red.last_path}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_fused_small_allreduce_of_histogram_losses_and_counts(tmp_path):
    """histogram + loss numerators + counts travel in ONE all-reduce; means are global sum / global count."""
    world = 2
    mp.spawn(_stats_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        got = torch.load(tmp_path / f"s{r}.pt")
        assert got["hist"].dtype == torch.int32 and got["hist"].tolist() == [2, 0, 5, 0, 5]
        assert abs(float(got["l1"]) - (3.0 + 6.0) / (2 + 3)) < 1e-12
        assert abs(float(got["psnr"]) - (10.0 + 11.0) / 2) < 1e-12
        rr = torch.load(tmp_path / f"r{r}.pt")
        assert rr["raised"] and rr["path"] == "cat"
    # single process: same arithmetic, no collective
    from titok_video_b200 import dist as D

    h, m = D.allreduce_step_stats(torch.tensor([1, 2], dtype=torch.int64), {"a": 6.0}, {"a": 4})
    assert h.tolist() == [1, 2] and float(m["a"]) == 1.5
    h, m = D.allreduce_step_stats(None, {"a": 6.0}, {"a": 0})
    assert h is None and float(m["a"]) == 6.0
