"""CUDA-graph replay of a training step for batch compositions that repeat.

At the reference's batch size (3 clips A under tiny.yaml's 6144-token budget) a training step is ~215 library launches of
~18 us each plus the loss in torch: the step is bound by the host's launch rate, and in DDP every rank waits for the
slowest host. `GraphedTrainStep` records forward + loss + backward of one batch composition (clip shapes and token
counts, in order) into a CUDA graph the first time it sees it and replays the graph afterwards: inputs are copied into
static buffers, gradients land in the same `.grad` tensors every replay, the optimizer step and the gradient all-reduce
stay outside the graph (eager, after the replay).

The reference's dataloader draws ragged batches (dataset/video_dataset.py:130-172), so compositions repeat only when the
data is bucketed (`data.canonical_order` sorts a batch so that equal multisets of shapes share one plan); a composition
that has not been seen simply runs eagerly the first time. This is the training-side counterpart of
`TiTok.tokenize_reconstruct_`'s graph replay.

    step = GraphedTrainStep(model, loss_fn)          # loss_fn(clips, recon, out_dict) -> scalar loss
    with reducer.no_sync():                          # DDP (dist.GradientAllReducer): its autograd hooks must not fire
        loss, out = step(clips, token_counts)        #   collectives inside a capture; forward + loss + backward
    reducer.reduce_now(); reducer.finish()           # all-reduce the gradients as they are, then optimizer.step()
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence, Tuple

import torch


def l1_loss(clips, recon, _out) -> torch.Tensor:
    """mean over clips of the per-clip mean absolute error (loss_module.py:118,160)."""
    return torch.stack([(r.float() - c.float()).abs().mean() for c, r in zip(clips, recon)]).mean()


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, loss_fn: Callable = l1_loss, autocast_dtype=torch.bfloat16,
                 max_graphs: int = 16, warmup: int = 2):
        self.model, self.loss_fn, self.dtype = model, loss_fn, autocast_dtype
        self.max_graphs, self.warmup = max_graphs, warmup
        self._entries: Dict[Tuple, dict] = {}
        self._seen: Dict[Tuple, int] = {}
        self.replays = 0
        self.eager_steps = 0

    def _eager(self, clips, tcs):
        with torch.autocast("cuda", dtype=self.dtype):
            recon, out = self.model(clips, tcs)
        loss = self.loss_fn(clips, recon, out)
        loss.backward()
        return loss.detach(), out

    def __call__(self, clips: Sequence[torch.Tensor], token_counts) -> Tuple[torch.Tensor, dict]:
        """Gradients are ACCUMULATED into existing `.grad`s in eager mode and OVERWRITTEN by a replay: call
        `optimizer.zero_grad(set_to_none=True)` (or not at all) between steps, like train.py does."""
        from .. import engine

        tcs = engine.to_host_ints(token_counts)
        key = (tuple(tuple(c.shape) for c in clips), tuple(tcs), clips[0].dtype)
        ent = self._entries.get(key)
        if ent is not None and ent["gen"] != engine.arena_generation():
            self._entries.pop(key)
            ent = None  # a workspace arena moved since the capture
        if ent is None:
            n = self._seen.get(key, 0) + 1
            self._seen[key] = n
            if n <= self.warmup or len(self._entries) >= self.max_graphs:
                self.eager_steps += 1
                return self._eager(clips, tcs)
            ent = self._capture(clips, tcs)
            self._entries[key] = ent
        for s, c in zip(ent["static_clips"], clips):
            s.copy_(c, non_blocking=True)
        ent["graph"].replay()
        self.replays += 1
        # hand the captured gradient tensors to the parameters (an optimizer's zero_grad(set_to_none=True) dropped them)
        for p, g in ent["grads"]:
            p.grad = g
        return ent["loss"], ent["out"]

    def _capture(self, clips, tcs) -> dict:
        from .. import engine

        static_clips = [torch.empty_like(c) for c in clips]
        for s, c in zip(static_clips, clips):
            s.copy_(c)
        params = [p for p in self.model.parameters() if p.requires_grad]
        for p in params:
            p.grad = None
        torch.cuda.current_stream().synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            loss, out = self._eager(static_clips, tcs)
        grads = [(p, p.grad) for p in params if p.grad is not None]
        return {"graph": g, "static_clips": static_clips, "loss": loss, "out": {k: v for k, v in out.items()},
                "grads": grads, "gen": engine.arena_generation()}
