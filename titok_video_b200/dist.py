"""Multi-GPU plumbing: one process per GPU, clips sharded across ranks, no data-path collective.

Clips never interact (attention is block-diagonal over cu_seqlens, norms are per row, FSQ per vector), so the
tokenizer shards by clip. The only collectives are bookkeeping:
  * codebook-usage histogram: one all-reduce(SUM) of [K] counts (the reference keeps per-rank statistics)
  * optional gather of the per-clip index tensors to one rank (tokenise-to-disk jobs)
`torch.distributed` (NCCL on the GPU box, gloo in the CPU tests) is used as plumbing only.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .plan import clip_cost, shard_clips


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_batch(shapes: Sequence[Sequence[int]], token_counts: Sequence[int], patch_size=(4, 8, 8), width: int = 256,
                layers: int = 4, group=None) -> List[int]:
    """Indices of the clips this rank owns. Deterministic on every rank (pure function of the shapes): greedy
    longest-processing-time balancing by forward FLOPs (linear + quadratic attention term)."""
    rank, world = _world(group)
    costs = [clip_cost(s, t, patch_size, width, layers) for s, t in zip(shapes, token_counts)]
    return shard_clips(costs, world)[rank]


def allreduce_counts(counts: torch.Tensor, group=None) -> torch.Tensor:
    """Sum a [K] histogram over ranks in place (no-op for a single process)."""
    _, world = _world(group)
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def gather_indices(local: Sequence[torch.Tensor], owned: Sequence[int], n_clips: int, dst: int = 0,
                   group=None) -> Optional[List[torch.Tensor]]:
    """Collect per-clip index tensors on rank `dst` in the original clip order. `local[i]` belongs to clip
    `owned[i]`. Returns the list on `dst`, None elsewhere. One padded all_gather (indices are a few KB per clip)."""
    rank, world = _world(group)
    if world == 1:
        out = [None] * n_clips
        for t, i in zip(local, owned):
            out[i] = t
        return out
    dev = local[0].device if len(local) else torch.device("cpu")
    lens = torch.tensor([t.numel() for t in local] + [0] * (n_clips - len(local)), dtype=torch.int64, device=dev)
    ids = torch.tensor(list(owned) + [-1] * (n_clips - len(owned)), dtype=torch.int64, device=dev)
    meta = torch.stack([ids, lens])  # [2, n_clips]
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    max_len = int(max(int(m[1].sum()) for m in metas))
    flat = torch.zeros(max(max_len, 1), dtype=torch.int32, device=dev)
    if len(local):
        cat = torch.cat([t.reshape(-1).to(torch.int32) for t in local])
        flat[:cat.numel()] = cat
    flats = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(flats, flat, group=group)
    if rank != dst:
        return None
    out: List[Optional[torch.Tensor]] = [None] * n_clips
    for m, f in zip(metas, flats):
        off = 0
        for cid, ln in zip(m[0].tolist(), m[1].tolist()):
            if cid >= 0:
                out[cid] = f[off:off + ln].clone()
                off += ln
    return out


class GradientAllReducer:
    """DDP-equivalent gradient averaging for the drop-in TiTok (train.py runs Lightning DDP, train.py:172-181).

    The backward of a stack (backward.EncoderFn / DecoderFn) produces ALL of that stack's parameter gradients at once,
    decoder first, then (through the FSQ straight-through estimator) the encoder. Parameters are therefore bucketed per
    top-level child module (`decoder`, `encoder`, ...): a post-accumulate-grad hook counts the bucket's parameters
    down, and when the last one has its gradient the bucket is flattened and all-reduced asynchronously -- the
    decoder bucket (13.7 MB fp32) crosses NVLink while the encoder backward is still computing. `finish()` waits,
    divides by the world size and scatters the averages back into `.grad`. One NCCL call per bucket; no per-parameter
    collectives. Single process: every method is a no-op.
    """

    def __init__(self, module: torch.nn.Module, group=None):
        self.group = group
        self.rank, self.world = _world(group)
        self.buckets: List[List[torch.nn.Parameter]] = []
        self._pending: List[int] = []
        self._work: List[Optional[tuple]] = []
        self._handles = []
        self._sync = True
        if self.world == 1:
            return
        by_child = {}
        for name, p in module.named_parameters():
            if p.requires_grad:
                by_child.setdefault(name.split(".")[0], []).append(p)
        for b, params in enumerate(by_child.values()):
            self.buckets.append(params)
            self._pending.append(len(params))
            self._work.append(None)
            for p in params:
                self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(b)))

    def no_sync(self):
        """Context manager for gradient accumulation (same contract as DistributedDataParallel.no_sync): backward passes
        inside it only accumulate into `.grad`; the first backward outside it all-reduces the accumulated gradients."""
        reducer = self

        class _NoSync:
            def __enter__(self_inner):
                reducer._sync = False

            def __exit__(self_inner, *exc):
                reducer._sync = True
                return False

        return _NoSync()

    def _make_hook(self, b: int):
        def hook(_param):
            if not self._sync:
                return
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch(b)
        return hook

    def _launch(self, b: int) -> None:
        flat = torch.cat([p.grad.reshape(-1).float() for p in self.buckets[b]])
        # NCCL averages inside the collective; gloo (CPU tests) has no AVG: sum now, divide in finish()
        avg = dist.get_backend(self.group) == "nccl"
        work = dist.all_reduce(flat, op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._work[b] = (work, flat, avg)

    def finish(self) -> None:
        """Call after loss.backward(): completes the outstanding all-reduces and writes the averaged gradients."""
        if self.world == 1:
            return
        for b, params in enumerate(self.buckets):
            if self._work[b] is None:
                if all(p.grad is not None for p in params):
                    self._launch(b)  # a bucket whose hooks did not all fire (gradient accumulation without zero_grad)
                else:
                    self._pending[b] = len(params)
                    continue
            work, flat, averaged = self._work[b]
            work.wait()
            if not averaged:
                flat.div_(self.world)
            off = 0
            for p in params:
                n = p.numel()
                g = flat[off:off + n].view_as(p)
                # re-point .grad at the averaged bucket (no copy kernels) when the dtypes agree
                if p.grad.dtype == g.dtype:
                    p.grad = g
                else:
                    p.grad.copy_(g)
                off += n
            self._work[b] = None
            self._pending[b] = len(params)

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []
