"""Codebook-usage statistics with the interface of the reference's train_utils/codebook_logging.py.

Same sliding window (the last `codebook_size` samples), same two scores; CUDA index tensors are histogrammed
on the device with the shared-memory-atomics kernel (ttk_hist_u32) and reduced to usage / entropy by
ttk_codebook_stats, so nothing forces a per-step `.cpu()` (train.py:115). CPU index tensors (what the
reference's train.py hands over) are uploaded and take the same kernels; there is no CPU arithmetic path.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn


class CodebookLogger(nn.Module):
    def __init__(self, codebook_size: int):
        super().__init__()
        self.codebook_size = int(codebook_size)
        self.code_frequencies = torch.zeros(self.codebook_size)
        self.codebook_counter = 0
        self.codebook_indices = []

    def forward(self, codes):
        for sample in codes:
            if len(self.codebook_indices) == self.codebook_size:
                self.codebook_indices.pop(0)
            self.codebook_indices.append(sample)

    def is_score_ready(self) -> bool:
        return len(self.codebook_indices) == self.codebook_size

    def _device(self) -> torch.device:
        for s in self.codebook_indices:
            if s.device.type == "cuda":
                return s.device
        from .. import engine

        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        engine.require_cuda(dev)  # no CUDA device: fail loudly, there is no CPU histogram
        return dev

    def counts(self, sync_dist: bool = False, process_group=None) -> torch.Tensor:
        """int32 [codebook_size] histogram of the buffered samples, computed by ttk_hist_u32 on the device. Samples
        that arrive as CPU tensors (what the reference's train.py:115 hands over) are uploaded first.
        sync_dist=True sums the histogram over the ranks of `process_group` (one small all-reduce; the reference
        keeps per-rank statistics)."""
        from .. import _lib, engine

        K = self.codebook_size
        dev = self._device()
        flat = torch.cat([s.reshape(-1).to(torch.int32) for s in self.codebook_indices]) if self.codebook_indices \
            else torch.zeros(0, dtype=torch.int32)
        flat = flat.to(dev).contiguous()
        cnt = torch.zeros(K, dtype=torch.int32, device=dev)
        if flat.numel():
            _lib.call("ttk_hist_u32", engine._ptr(flat), flat.numel(), K, engine._ptr(cnt), engine._stream())
        if sync_dist and torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(cnt, group=process_group)
        return cnt

    @staticmethod
    def scores_from_counts(counts: torch.Tensor) -> dict:
        """usage % and entropy (nats) of a [K] count vector on the device (ttk_codebook_stats)."""
        from .. import _lib, engine

        engine.require_cuda(counts.device)
        K = counts.numel()
        c32 = counts.to(torch.int32).contiguous()
        out = torch.empty(3, dtype=torch.float64, device=counts.device)
        _lib.call("ttk_codebook_stats", engine._ptr(c32), K, engine._ptr(out), engine._stream())
        nz, ent, _ = out.cpu().tolist()
        return {"codebook/usage_percent": torch.tensor(nz / K * 100.0), "codebook/entropy": ent}

    def get_scores(self) -> Optional[dict]:
        if not self.is_score_ready():
            return None
        d = self.scores_from_counts(self.counts())
        self.codebook_indices = []
        return d
