"""Dynamic token-budget batching with the semantics of the reference's `_dynamic_batching`
(dataset/video_dataset.py:130-172), plus a canonical clip order that lets batches with the same multiset of
(shape, token count) share one packing plan / CUDA graph.

`dynamic_batches` is host logic only (Python ints, no tensors are touched): it groups a stream of samples so that the
packed sequence length  sum_i (patches_i + tokens_i)  never exceeds the budget -- the quantity that sizes every
buffer of the path (M rows of the packed activation matrix).
"""
from __future__ import annotations

import math
import random
from typing import Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch


def _grid_size(shape_thw: Sequence[int], patch_size: Sequence[int]) -> int:
    return math.prod(int(x) // int(y) for x, y in zip(shape_thw, patch_size))


def dynamic_batches(data: Iterable[Dict], patch_size: Sequence[int], token_range: Sequence[int], max_grid: Sequence[int],
                    max_seq_len: int, eval: bool = False, max_samples: Optional[int] = None,
                    randrange: Callable[[int, int], int] = random.randrange) -> Iterator[Dict]:
    """Yields dicts {'video': [clips], ..., 'token_counts': int32 tensor} like the reference's dataloader stage.

    Semantics kept from video_dataset.py:130-172 (including its edge behaviour):
      * every sample draws `token_count = randrange(token_range[0], token_range[1] + 1)` (:148);
      * a batch is emitted when the NEXT sample would push  sum(grid_size + token_count)  over `max_seq_len` (:157-168);
        that sample opens the next batch; the last, partially filled batch of a finite stream is never emitted;
      * eval: stops after `max_samples + 1` samples have been seen (:150-155);
      * the budget must fit the largest possible sample (:143).
    `randrange` is injectable so that a seeded generator reproduces the reference's draw sequence.
    """
    assert _grid_size(max_grid, patch_size) + token_range[1] <= max_seq_len, \
        "max seq_len (max_grid/patch_size + token_range[1]) must be less than trg_seq_len"
    chunks: List[Dict] = []
    token_counts: List[int] = []
    curr, seen = 0, 0
    for sample in data:
        grid_size = _grid_size(sample["video"].shape[1:], patch_size)
        token_count = randrange(token_range[0], token_range[1] + 1)
        if eval:
            if max_samples is not None and seen > max_samples:
                break
            seen += 1
        if curr + grid_size + token_count > max_seq_len:
            out = {k: [c[k] for c in chunks] for k in chunks[0].keys()}
            out["token_counts"] = torch.tensor(token_counts, dtype=torch.int32)
            yield out
            chunks, token_counts, curr = [], [], 0
        curr += grid_size + token_count
        chunks.append(sample)
        token_counts.append(token_count)


def canonical_order(shapes: Sequence[Sequence[int]], token_counts: Sequence[int]) -> Tuple[List[int], List[int]]:
    """(perm, inverse): clips sorted by (shape, token count). Results of the path do not depend on the batch composition
    or order (attention is block-diagonal, tests/test_gpu_model.py::test_batch_composition_does_not_change_results), so
    feeding `[clips[i] for i in perm]` and un-permuting the outputs with `inverse` is exact -- and every batch with the same
    multiset of (shape, token count) then maps to ONE cached packing plan and ONE captured CUDA graph."""
    perm = sorted(range(len(token_counts)), key=lambda i: (tuple(int(v) for v in shapes[i]), int(token_counts[i])))
    inverse = [0] * len(perm)
    for pos, i in enumerate(perm):
        inverse[i] = pos
    return perm, inverse
