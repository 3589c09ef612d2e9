"""Codebook-usage statistics with the interface of the reference's train_utils/codebook_logging.py.

Same sliding window (the last `codebook_size` samples), same two scores; CUDA index tensors are histogrammed
on the device with the shared-memory-atomics kernel (ttk_hist_u32) and reduced to usage / entropy by
ttk_codebook_stats, so nothing forces a per-step `.cpu()` (train.py:115). CPU index tensors (what the
reference's train.py hands over) take the torch.bincount route of the reference.
"""
from __future__ import annotations

import math
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

    def counts(self, sync_dist: bool = False, process_group=None) -> torch.Tensor:
        """int64 [codebook_size] histogram of the buffered samples. sync_dist=True sums it over the ranks of
        `process_group` (one small all-reduce; the reference keeps per-rank statistics)."""
        K = self.codebook_size
        cuda = [s for s in self.codebook_indices if s.device.type == "cuda"]
        cpu = [s for s in self.codebook_indices if s.device.type != "cuda"]
        total = torch.zeros(K, dtype=torch.int64)
        if cuda:
            from .. import _lib, engine

            dev = cuda[0].device
            flat = torch.cat([s.reshape(-1).to(torch.int32) for s in cuda]).contiguous()
            cnt = torch.zeros(K, dtype=torch.int32, device=dev)
            _lib.call("ttk_hist_u32", engine._ptr(flat), flat.numel(), K, engine._ptr(cnt), engine._stream())
            total = cnt.to(torch.int64)
        if cpu:
            c = torch.bincount(torch.cat([s.reshape(-1) for s in cpu]).to(torch.int64), minlength=K)[:K]
            total = total + c.to(total.device)
        if sync_dist and torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(total, group=process_group)
        return total

    @staticmethod
    def scores_from_counts(counts: torch.Tensor) -> dict:
        K = counts.numel()
        if counts.device.type == "cuda":
            from .. import _lib, engine

            c32 = counts.to(torch.int32).contiguous()
            out = torch.empty(3, dtype=torch.float64, device=counts.device)
            _lib.call("ttk_codebook_stats", engine._ptr(c32), K, engine._ptr(out), engine._stream())
            nz, ent, _ = out.cpu().tolist()
        else:
            f = counts.to(torch.float64)
            tot = float(f.sum())
            nz = float((f > 0).sum())
            p = f[f > 0] / tot if tot > 0 else f[:0]
            ent = float(-(p * p.log()).sum())
        return {"codebook/usage_percent": torch.tensor(nz / K * 100.0), "codebook/entropy": ent}

    def get_scores(self) -> Optional[dict]:
        if not self.is_score_ready():
            return None
        d = self.scores_from_counts(self.counts())
        self.codebook_indices = []
        return d
