"""The step after the path: a compact on-disk / on-wire container for tokenised clips -- exactly what
`TiTok.decode_indices(indices, grids)` (model/titok.py:54-62) needs to reconstruct them: per clip the pixel grid
(T, H, W), the token count and the FSQ indices.

Layout (little endian):
    magic  b"TTKV1\\0"        6 bytes
    u16    index width in bytes (2 when codebook_size <= 65536, else 4)
    u32    codebook_size
    u32    n_clips
    n_clips x { u16 T, u16 H, u16 W, u16 token_count }
    indices of all clips back to back (u16 or u32), clip order = record order
The reference has no token container (its training loop never stores tokens); this one is deliberately trivial: no
compression, O(1) seek to a clip through the prefix sum of the token counts.
"""
from __future__ import annotations

import struct
from typing import BinaryIO, List, Sequence, Tuple, Union

import numpy as np
import torch

MAGIC = b"TTKV1\0"


def write_tokens(f: Union[str, BinaryIO], indices: Sequence[torch.Tensor], grids: Sequence[Sequence[int]],
                 codebook_size: int) -> int:
    """Writes one container; returns the number of bytes written. `indices[i]`: 1-D integer tensor of clip i (any device)."""
    if len(indices) != len(grids):
        raise ValueError("one grid per clip")
    width = 2 if codebook_size <= 65536 else 4
    dt = np.uint16 if width == 2 else np.uint32
    head = [MAGIC, struct.pack("<HII", width, int(codebook_size), len(indices))]
    body = []
    for idx, g in zip(indices, grids):
        a = idx.detach().reshape(-1).cpu().numpy()
        if a.size and (a.min() < 0 or a.max() >= codebook_size):
            raise ValueError("index outside the codebook")
        if a.size > 65535 or max(int(v) for v in g) > 65535:
            raise ValueError("clip too large for the record format")
        head.append(struct.pack("<HHHH", int(g[0]), int(g[1]), int(g[2]), int(a.size)))
        body.append(a.astype(dt).tobytes())
    blob = b"".join(head + body)
    if isinstance(f, str):
        with open(f, "wb") as fh:
            fh.write(blob)
    else:
        f.write(blob)
    return len(blob)


def read_tokens(f: Union[str, BinaryIO, bytes]) -> Tuple[List[torch.Tensor], List[Tuple[int, int, int]], int]:
    """-> (per-clip int32 index tensors (CPU), per-clip (T, H, W), codebook_size): the arguments of decode_indices."""
    if isinstance(f, str):
        with open(f, "rb") as fh:
            blob = fh.read()
    elif isinstance(f, (bytes, bytearray)):
        blob = bytes(f)
    else:
        blob = f.read()
    if blob[:6] != MAGIC:
        raise ValueError("not a TTKV1 token container")
    width, codebook_size, n = struct.unpack_from("<HII", blob, 6)
    if width not in (2, 4):
        raise ValueError("corrupt header")
    off = 6 + 10
    recs = np.frombuffer(blob, dtype="<u2", count=4 * n, offset=off).reshape(n, 4)
    off += 8 * n
    counts = recs[:, 3].astype(np.int64)
    total = int(counts.sum())
    if len(blob) != off + total * width:
        raise ValueError("truncated or oversized token container")
    flat = np.frombuffer(blob, dtype="<u2" if width == 2 else "<u4", count=total, offset=off).astype(np.int32)
    if total and int(flat.max()) >= codebook_size:
        raise ValueError("index outside the codebook")
    starts = np.concatenate([[0], np.cumsum(counts)])
    indices = [torch.from_numpy(flat[starts[i]:starts[i + 1]].copy()) for i in range(n)]
    grids = [tuple(int(v) for v in recs[i, :3]) for i in range(n)]
    return indices, grids, int(codebook_size)
