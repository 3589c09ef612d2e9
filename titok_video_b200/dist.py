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
    # a rank that owns no clip must still take part with tensors on the backend's device (NCCL: the current GPU)
    if len(local):
        dev = local[0].device
    elif dist.get_backend(group) == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
    else:
        dev = torch.device("cpu")
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


def allreduce_step_stats(hist: Optional[torch.Tensor], sums: dict, counts: dict, group=None):
    """ONE small all-reduce per step for everything train.py logs (train.py:111-115): the [K] codebook
    histogram, loss numerators and their denominators. The reference takes unweighted per-rank means; the global values
    are sum / count, so the step ships sums and counts and divides afterwards:

        hist, means = allreduce_step_stats(hist_i32, {"l1": l1_sum}, {"l1": n_clips})

    `sums` values are 0-d tensors (or floats), `counts` values ints / 0-d tensors. Everything travels as float64 (exact
    for counts below 2^53) in a single buffer [K + 2 n]. Returns (hist summed over ranks in its own dtype, or None;
    {name: global sum / global count}). Single process: no collective, same arithmetic."""
    _, world = _world(group)
    names = sorted(sums)
    if sorted(counts) != names:
        raise ValueError("sums and counts need the same keys")
    dev = hist.device if hist is not None else next((v.device for v in sums.values() if torch.is_tensor(v)), torch.device("cpu"))
    k = hist.numel() if hist is not None else 0
    buf = torch.empty(k + 2 * len(names), dtype=torch.float64, device=dev)
    if k:
        buf[:k] = hist.reshape(-1).to(torch.float64)
    for i, n in enumerate(names):
        buf[k + 2 * i] = sums[n].detach().to(torch.float64) if torch.is_tensor(sums[n]) else float(sums[n])
        buf[k + 2 * i + 1] = counts[n].to(torch.float64) if torch.is_tensor(counts[n]) else float(counts[n])
    if world > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    out_hist = buf[:k].round().to(hist.dtype).view(hist.shape) if hist is not None else None
    means = {n: buf[k + 2 * i] / buf[k + 2 * i + 1].clamp(min=1.0) for i, n in enumerate(names)}
    return out_hist, means


class GradientAllReducer:
    """DDP-equivalent gradient averaging for the drop-in TiTok (train.py runs Lightning DDP, train.py:172-181).

    The backward of a stack (backward.EncoderFn / DecoderFn) produces ALL of that stack's parameter gradients at once,
    decoder first, then (through the FSQ straight-through estimator) the encoder. Parameters are therefore bucketed per
    top-level child module (`decoder`, `encoder`, ...): a post-accumulate-grad hook counts the bucket's parameters
    down, and when the last one has its gradient the bucket is all-reduced asynchronously -- the decoder bucket
    (13.7 MB fp32) crosses NVLink while the encoder backward is still computing. The CUDA backward leaves every
    gradient of a stack as a view of ONE flat fp32 buffer, and that buffer is what is reduced, in place: no flatten, no
    write-back (`last_path == "inplace"`); gradients that live in separate tensors (plain torch modules, the CPU tests)
    are concatenated first (`"cat"`). `finish()` waits (and divides by the world size where the backend cannot
    average). One NCCL call per bucket; no per-parameter collectives. Single process: every method is a no-op.
    """

    def __init__(self, module: torch.nn.Module, group=None):
        self.group = group
        self.rank, self.world = _world(group)
        self.buckets: List[List[torch.nn.Parameter]] = []
        self._pending: List[int] = []
        self._work: List[Optional[tuple]] = []
        self._handles = []
        self._sync = True
        self.last_path = None  # "inplace" (the stack's flat gradient buffer was all-reduced directly) or "cat"
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
            elif self._pending[b] < 0:
                raise RuntimeError("GradientAllReducer: a second backward ran before finish() (use no_sync() for "
                                   "gradient accumulation)")
        return hook

    @staticmethod
    def _shared_flat(params) -> Optional[torch.Tensor]:
        """The contiguous fp32 range that holds EVERY gradient of the bucket when they are all views of one buffer (the
        CUDA backward writes a stack's gradients into one flat buffer, backward._zero_grads), else None."""
        g0 = params[0].grad
        if g0 is None or g0.dtype != torch.float32:
            return None
        base = g0.untyped_storage().data_ptr()
        lo, hi = None, None
        for p in params:
            g = p.grad
            if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.untyped_storage().data_ptr() != base:
                return None
            a, b_ = g.storage_offset(), g.storage_offset() + g.numel()
            lo = a if lo is None else min(lo, a)
            hi = b_ if hi is None else max(hi, b_)
        if hi - lo > 2 * sum(p.numel() for p in params) + 4096:  # a buffer shared with something much larger: do not
            return None
        return torch.empty(0, dtype=torch.float32, device=g0.device).set_(g0.untyped_storage(), lo, (hi - lo,), (1,))

    def _launch(self, b: int) -> None:
        if self._work[b] is not None:
            raise RuntimeError("GradientAllReducer: a second backward reached this bucket before finish() was called "
                               "(use no_sync() for gradient accumulation)")
        # NCCL averages inside the collective; gloo (CPU tests) has no AVG: sum now, divide in finish()
        avg = dist.get_backend(self.group) == "nccl"
        flat = self._shared_flat(self.buckets[b])
        inplace = flat is not None
        if not inplace:
            flat = torch.cat([p.grad.reshape(-1).float() for p in self.buckets[b]])
        self.last_path = "inplace" if inplace else "cat"
        work = dist.all_reduce(flat, op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._work[b] = (work, flat, avg, inplace)

    def reduce_now(self) -> None:
        """Launch every bucket's all-reduce on the gradients as they are (no autograd hooks involved): for steps whose
        backward did not run through autograd in this process -- a CUDA-graph replay (train_utils.GraphedTrainStep)."""
        if self.world == 1:
            return
        for b, params in enumerate(self.buckets):
            if self._work[b] is None and all(p.grad is not None for p in params):
                self._launch(b)

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
            work, flat, averaged, inplace = self._work[b]
            work.wait()
            if not averaged:
                flat.div_(self.world)
            if inplace:  # the gradients ARE the buffer that was reduced: nothing to write back
                self._work[b] = None
                self._pending[b] = len(params)
                continue
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
