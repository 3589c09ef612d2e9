#!/usr/bin/env python
"""Tokenise-to-disk batch job: clips are sharded over the ranks of one box (no data-path collective), every rank runs the
encoder + FSQ on its shard, the per-clip token indices are gathered on rank 0 and written to one token container
(titok_video_b200.data.write_tokens) that `TiTok.decode_indices` can consume later.

    python scripts/tokenize_job.py --clips 64 --out /tmp/tokens.ttkv                       # 1 GPU, synthetic clips
    torchrun --nproc-per-node 8 scripts/tokenize_job.py --clips 512 --out /tmp/tokens.ttkv   # 8 GPUs

Synthetic uint8 frames with the shapes / token counts of configs/tiny.yaml's sampling ranges stand in for a dataset (there
is no network in this environment); `--verify` decodes the container on rank 0 and checks the reconstruction of the first
clips against a direct forward pass (bit-exact)."""
import argparse, os, random, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import titok_video_b200 as T  # noqa: E402
from titok_video_b200 import dist as D  # noqa: E402
from titok_video_b200.data import canonical_order, read_tokens, write_tokens  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=64)
    ap.add_argument("--out", default="/tmp/tokens.ttkv")
    ap.add_argument("--budget", type=int, default=30000, help="packed rows per launch sequence")
    ap.add_argument("--verify", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(42)
    model = T.TiTok(T.load_config(os.path.join(ROOT, "configs", "tiny.yaml"))).to(dev).eval()
    rnd = random.Random(0)  # every rank draws the same job description
    shapes = [(rnd.choice([8, 12, 16]), rnd.choice([128, 144, 168]), rnd.choice([128, 144, 168])) for _ in range(a.clips)]
    tcs = [rnd.randint(1, 128) for _ in range(a.clips)]
    owned = D.shard_batch(shapes, tcs)  # deterministic LPT partition by forward cost, identical on every rank

    def frames(i):  # decoded uint8 frames of clip i (synthetic)
        g = torch.Generator().manual_seed(1000 + i)
        return torch.randint(0, 256, (3, *shapes[i]), generator=g, dtype=torch.uint8)

    t0 = time.perf_counter()
    local_idx = {}
    batch, rows = [], 0
    def flush():
        nonlocal batch, rows
        if not batch:
            return
        perm, _ = canonical_order([shapes[i] for i in batch], [tcs[i] for i in batch])  # same multiset -> same plan
        ids = [batch[j] for j in perm]
        with torch.no_grad():
            _, d = model.encode([frames(i).to(dev, non_blocking=True) for i in ids], [tcs[i] for i in ids], split_indices=True)
        for i, idx in zip(ids, d["indices"]):
            local_idx[i] = idx
        batch, rows = [], 0
    for i in owned:
        s = (shapes[i][0] // 4) * (shapes[i][1] // 8) * (shapes[i][2] // 8) + tcs[i]
        if rows + s > a.budget:
            flush()
        batch.append(i)
        rows += s
    flush()
    torch.cuda.synchronize()
    gathered = D.gather_indices([local_idx[i] for i in owned], owned, a.clips, dst=0)
    dt = time.perf_counter() - t0
    if rank == 0:
        n = write_tokens(a.out, gathered, shapes, model.quantize.codebook_size)
        print(f"{a.clips} clips, {sum(tcs)} tokens on {world} GPU(s) in {dt * 1e3:.1f} ms -> {a.out} ({n} bytes)")
        if a.verify:
            idx, grids, K = read_tokens(a.out)
            assert K == model.quantize.codebook_size and grids == shapes and [len(t) for t in idx] == tcs
            k = min(3, a.clips)
            with torch.no_grad():
                rec_a = model.decode_indices([t.to(dev) for t in idx[:k]], grids[:k])
                rec_b, d = model([frames(i).to(dev) for i in range(k)], tcs[:k])
            assert torch.equal(torch.cat([t.to(dev) for t in idx[:k]]), d["indices"])
            assert all(torch.equal(x, y) for x, y in zip(rec_a, rec_b))
            print("verify ok: container -> decode_indices == forward (bit-exact)")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
