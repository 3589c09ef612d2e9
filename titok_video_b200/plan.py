"""Host-side metadata planner for a packed (variable-length) batch of clips.

Everything TiTokEncoder.forward / TiTokDecoder.forward derive on the device with synchronising ops
(reference model/base/blocks.py:72-89,154-162 and RoPE.forward, model/base/rope.py:57-71) is derived here from
Python ints with numpy, once per (shapes, token_counts) signature, with zero device syncs:

  * packed row layout: per clip, `token_count` latent rows followed by the patch rows (blocks.py:85-86)
  * cu_seqlens (blocks.py:81-83)
  * gather / scatter maps between packed rows, latent tokens and patches
  * patch geometry for patchify / unpatchify (model/base/utils.py:26-51)
  * the RoPE table: fp64 angles -> fp32 (cos, sin), 30 complex lanes per token (rope.py:40-54)
  * the attention work list (pairs of 128-row query tiles sharing one K/V stream)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

ATTN_TILE = 128
ROPE_LANES_PER_AXIS = 10  # head_dim 64 // (3 axes * 2)


@dataclass
class PackedPlan:
    # batch description (host ints)
    grids_px: Tuple[Tuple[int, ...], ...]  # per clip (T, H, W) in pixels
    grids: Tuple[Tuple[int, ...], ...]  # per clip (T/p0, H/p1, W/p2) in patches
    token_counts: Tuple[int, ...]
    patch_size: Tuple[int, ...]
    channels: int
    # derived sizes
    M: int = 0  # packed rows
    T: int = 0  # latent tokens
    G: int = 0  # patches
    seq_lens: Tuple[int, ...] = ()
    clip_numel: Tuple[int, ...] = ()
    clip_offset: Tuple[int, ...] = ()
    total_numel: int = 0
    # numpy metadata (uploaded by the engine)
    cu_seqlens: np.ndarray = None  # int32 [B+1]
    enc_src_row: np.ndarray = None  # int32 [M]: patch index or -1 (latent)
    dec_src_row: np.ndarray = None  # int32 [M]: latent token index or -1 (patch)
    latent_row: np.ndarray = None  # int32 [T]
    patch_row: np.ndarray = None  # int32 [G]
    geom: np.ndarray = None  # int64 [G,4]: offset, W, H*W, T*H*W
    clip_desc: np.ndarray = None  # int64 [B,12]: per-clip descriptor for the device-side expansion (ttk_build_plan)
    max_pos: int = 0
    rope_pos: np.ndarray = None  # int32 [M,3]: integer (T,H,W) position ids of every packed row (rope.py:57-71)
    _rope: np.ndarray = None  # float32 [M,60], built on demand (tests / oracle); the engine builds it on the device
    attn_work: Dict[Tuple[int, int], np.ndarray] = field(default_factory=dict)  # (hq,hkv) -> int32 [n,12]

    @property
    def rope(self) -> np.ndarray:
        """float32 [M,60] (cos, sin) table on the host (the device copy is gathered by ttk_rope_table_gather from
        `rope_pos` and the per-id cos / sin table, see engine.DevicePlan)."""
        if self._rope is None:
            self._rope = rope_table_from_int_ids(self.rope_pos, rope_inv_freqs())
        return self._rope

    @property
    def key(self):
        return (self.grids_px, self.token_counts, self.patch_size, self.channels)


def rope_inv_freqs(theta: float = 10000.0, n: int = ROPE_LANES_PER_AXIS) -> np.ndarray:
    """theta ** linspace(0, 1, n) * pi / 2 in float64 (rope.py:42-45). Evaluated with torch on the host so the
    ten frequencies are bit-identical to the reference's buffer (angles reach ~2.4e6 rad)."""
    import torch

    f = torch.pow(theta, torch.linspace(0.0, 1.0, n, dtype=torch.float64)) * torch.pi / 2.0
    return f.numpy().copy()


def rope_ids(grid: Sequence[int], token_count: int) -> np.ndarray:
    """Position ids [token_count + prod(grid), len(grid)]: latent j -> (j,..,j); patch (a,b,c) -> (a,b,c)+token_count
    with the last axis fastest (torch.cartesian_prod order, rope.py:61-66)."""
    nd = len(grid)
    tok = np.repeat(np.arange(token_count, dtype=np.float64)[:, None], nd, axis=1)
    axes = np.meshgrid(*[np.arange(g, dtype=np.float64) for g in grid], indexing="ij")
    pat = np.stack([a.reshape(-1) for a in axes], axis=-1) + float(token_count)
    return np.concatenate([tok, pat], axis=0)


def rope_table(ids: np.ndarray, inv_freqs: np.ndarray) -> np.ndarray:
    """(cos, sin) pairs, float32 [L, n_freq*n_axes*2]; complex lane index = freq * n_axes + axis (interleaved
    layout of rope.py:49-53); angles are formed in float64 and only the final cos/sin are rounded to float32
    (apply_rotary_emb casts the complex128 table to complex64, rope.py:24)."""
    ang = inv_freqs[None, :, None] * ids[:, None, :]  # [L, F, A]
    ang = ang.reshape(ids.shape[0], -1)
    out = np.empty((ids.shape[0], ang.shape[1], 2), dtype=np.float32)
    out[..., 0] = np.cos(ang)
    out[..., 1] = np.sin(ang)
    return out.reshape(ids.shape[0], -1)


_CS_CACHE = {"n": 0, "table": None}


def _cos_sin_by_id(n_ids: int, inv_freqs: np.ndarray) -> np.ndarray:
    """float32 [n_ids, F, 2]: (cos, sin) of inv_freqs[f] * id for the integer position ids 0..n_ids-1. Position ids are
    small integers (latent index, or patch coordinate + token count), so every table entry of rope_table is one of
    these values: same float64 product, same float64 cos / sin, same rounding. Grown on demand, computed once."""
    if _CS_CACHE["n"] < n_ids:
        n = max(1024, 2 * n_ids)
        ang = np.arange(n, dtype=np.float64)[:, None] * inv_freqs[None, :]
        t = np.empty((n, inv_freqs.shape[0], 2), dtype=np.float32)
        t[..., 0] = np.cos(ang)
        t[..., 1] = np.sin(ang)
        _CS_CACHE["n"], _CS_CACHE["table"] = n, t
    return _CS_CACHE["table"]


def rope_table_from_int_ids(ids: np.ndarray, inv_freqs: np.ndarray) -> np.ndarray:
    """rope_table for integer-valued ids [L, A] through the per-id cache: a gather instead of L*F*A cos/sin pairs
    (bit-identical to rope_table, checked in tests/test_host_logic.py)."""
    ii = ids.astype(np.int64)
    t = _cos_sin_by_id(int(ii.max()) + 1 if ii.size else 1, inv_freqs)  # [n, F, 2]
    out = t[ii]  # [L, A, F, 2]
    return np.ascontiguousarray(out.transpose(0, 2, 1, 3)).reshape(ids.shape[0], -1)


def attn_work_list(seq_starts: Sequence[int], seq_lens: Sequence[int], hq: int, hkv: int,
                   q_lens: Sequence[int] = None) -> np.ndarray:
    """int32 [n, 12] records {q_row0[2], q_valid[2], q_head[2], kv_head, kv_row0, kv_len, kmax2, leader, kmax2b}
    (csrc/attn.cu). Two query tiles per record share one kv head: two heads of the same group when the group size is
    even, otherwise two consecutive row tiles of one head. Longest sequences first (LPT) to shorten the tail.
    `leader` = index of the first record of the same (clip, kv head): the kernel library keeps that pair's score bound
    in the leader's `kmax2` / `kmax2b` fields (device-side scratch, zero here).
    `q_lens` (optional, per clip): only the row tiles that hold a clip's first `q_lens[i]` rows get a record -- the
    encoder's last layer needs attention output for the latent rows alone (blocks.py:101: `x[latent_mask]`), which lead
    every clip. The tiles that remain are exactly the records of the full list (same rows, same valid counts)."""
    ratio = hq // hkv
    if q_lens is None:
        q_lens = seq_lens
    if ratio % 2 == 0 and len(seq_starts):
        # vectorised: one record per (clip, row tile, pair of query heads of one kv group)
        st = np.asarray(seq_starts, dtype=np.int64)
        sl = np.asarray(seq_lens, dtype=np.int64)
        nt = (np.minimum(np.asarray(q_lens, dtype=np.int64), sl) + ATTN_TILE - 1) // ATTN_TILE
        clip = np.repeat(np.arange(len(sl)), nt)
        ti = np.arange(int(nt.sum()), dtype=np.int64) - np.repeat(np.concatenate([[0], np.cumsum(nt)[:-1]]), nt)
        r0 = st[clip] + ti * ATTN_TILE
        valid = np.minimum(ATTN_TILE, sl[clip] - ti * ATTN_TILE)
        hp = hq // 2
        rec = np.zeros((len(ti) * hp, 12), dtype=np.int32)
        h = np.tile(np.arange(0, hq, 2), len(ti))
        rep = lambda v: np.repeat(v, hp)
        rec[:, 0] = rec[:, 1] = rep(r0)
        rec[:, 2] = rec[:, 3] = rep(valid)
        rec[:, 4], rec[:, 5], rec[:, 6] = h, h + 1, h // ratio
        rec[:, 7], rec[:, 8] = rep(st[clip]), rep(sl[clip])
        order = np.argsort(-rec[:, 8], kind="stable")  # longest sequences first (LPT), ties in generation order
        return _with_leaders(np.ascontiguousarray(rec[order]))
    recs: List[List[int]] = []
    for start, slen, qlen in zip(seq_starts, seq_lens, q_lens):
        n_tiles = (min(slen, qlen) + ATTN_TILE - 1) // ATTN_TILE
        if ratio % 2 == 0:
            for ti in range(n_tiles):
                r0 = start + ti * ATTN_TILE
                valid = min(ATTN_TILE, slen - ti * ATTN_TILE)
                for h in range(0, hq, 2):
                    recs.append([r0, r0, valid, valid, h, h + 1, h // ratio, start, slen, 0, 0, 0])
        else:
            for h in range(hq):
                for ti in range(0, n_tiles, 2):
                    r0 = start + ti * ATTN_TILE
                    v0 = min(ATTN_TILE, slen - ti * ATTN_TILE)
                    if ti + 1 < n_tiles:
                        r1 = r0 + ATTN_TILE
                        v1 = min(ATTN_TILE, slen - (ti + 1) * ATTN_TILE)
                    else:
                        r1, v1 = r0, 0
                    recs.append([r0, r1, v0, v1, h, h, h // ratio, start, slen, 0, 0, 0])
    recs.sort(key=lambda r: -r[8])
    return _with_leaders(np.asarray(recs, dtype=np.int32).reshape(-1, 12))


def _with_leaders(rec: np.ndarray) -> np.ndarray:
    """Fills column 10 (`leader`): the index of the first record with the same (kv_row0, kv_head)."""
    if rec.shape[0]:
        key = rec[:, 7].astype(np.int64) * 4096 + rec[:, 6]
        _, first, inv = np.unique(key, return_index=True, return_inverse=True)
        rec[:, 10] = first[inv.reshape(-1)]
    return rec


def make_plan(grids_px: Sequence[Sequence[int]], token_counts: Sequence[int], patch_size: Sequence[int],
              channels: int = 3, arrays: bool = True) -> PackedPlan:
    """arrays=False computes only the O(B) part (sizes, prefix sums, `clip_desc`); the engine then expands the per-row
    metadata on the device (ttk_build_plan). arrays=True also fills the numpy arrays on the host (tests, tools)."""
    patch_size = tuple(int(p) for p in patch_size)
    grids_px = tuple(tuple(int(v) for v in g) for g in grids_px)
    token_counts = tuple(int(t) for t in token_counts)
    if len(grids_px) != len(token_counts):
        raise ValueError("one token count per clip is required")
    if len(patch_size) != 3:
        raise ValueError("the CUDA path supports 3-D (T,H,W) patching only")
    grids = []
    for g in grids_px:
        if len(g) != 3 or any(v <= 0 or v % p for v, p in zip(g, patch_size)):
            raise ValueError(f"clip shape {g} is not a positive multiple of patch size {patch_size}")
        grids.append(tuple(v // p for v, p in zip(g, patch_size)))
    if any(t < 0 for t in token_counts):
        raise ValueError("token counts must be non-negative")
    plan = PackedPlan(grids_px=grids_px, grids=tuple(grids), token_counts=token_counts, patch_size=patch_size,
                      channels=channels)

    gsz = [g[0] * g[1] * g[2] for g in grids]
    seq = [g + t for g, t in zip(gsz, token_counts)]
    plan.seq_lens = tuple(seq)
    plan.cu_seqlens = np.concatenate([[0], np.cumsum(seq)]).astype(np.int32)
    plan.M, plan.T, plan.G = int(sum(seq)), int(sum(token_counts)), int(sum(gsz))
    numel = [channels * t * h * w for (t, h, w) in grids_px]
    plan.clip_numel = tuple(numel)
    plan.clip_offset = tuple(int(v) for v in np.concatenate([[0], np.cumsum(numel)[:-1]]))
    plan.total_numel = int(sum(numel))

    B = len(grids)
    g_arr = np.asarray(grids, dtype=np.int64).reshape(B, 3)
    px = np.asarray(grids_px, dtype=np.int64).reshape(B, 3)
    tc_arr = np.asarray(token_counts, dtype=np.int64)
    ng_arr = np.asarray(gsz, dtype=np.int64)
    row_start = np.concatenate([[0], np.cumsum(tc_arr + ng_arr)[:-1]]).astype(np.int64)  # first packed row of a clip
    tok_start = np.concatenate([[0], np.cumsum(tc_arr)[:-1]]).astype(np.int64)
    pat_start = np.concatenate([[0], np.cumsum(ng_arr)[:-1]]).astype(np.int64)
    clip_off = np.asarray(plan.clip_offset, dtype=np.int64)
    # per-clip descriptor consumed by ttk_build_plan (csrc/rowops.cu): int64 [B, 12]
    desc = np.zeros((B, 12), dtype=np.int64)
    desc[:, 0], desc[:, 1], desc[:, 2], desc[:, 3] = row_start, tok_start, pat_start, tc_arr
    desc[:, 4], desc[:, 5], desc[:, 6] = ng_arr, g_arr[:, 1], g_arr[:, 2]
    desc[:, 7], desc[:, 8], desc[:, 9], desc[:, 10] = clip_off, px[:, 2], px[:, 1] * px[:, 2], px[:, 0] * px[:, 1] * px[:, 2]
    plan.clip_desc = desc
    plan.max_pos = int((tc_arr + g_arr.max(axis=1)).max()) if B else 0  # upper bound of every position id (+1)
    if not arrays:
        return plan

    # Everything below is vectorised over the whole batch (no per-clip numpy calls): a stream of ragged batches pays
    # this planner on every step (train.py / tokenisation jobs never repeat a batch composition).
    # np.repeat of per-clip values is much cheaper than fancy indexing; 32-bit arithmetic wherever the range allows
    rep_t = lambda v: np.repeat(v, tc_arr)
    rep_p = lambda v: np.repeat(v, ng_arr)
    i32 = np.int32
    # latent tokens: index within the clip and packed row
    t_loc = np.arange(plan.T, dtype=i32) - rep_t(tok_start.astype(i32))
    latent_row = rep_t(row_start.astype(i32)) + t_loc
    # patches: index within the clip, (d0, d1, d2) with the last axis fastest (torch.cartesian_prod order)
    p_loc = np.arange(plan.G, dtype=i32) - rep_p(pat_start.astype(i32))
    g2 = rep_p(g_arr[:, 2].astype(i32))
    g12 = rep_p((g_arr[:, 1] * g_arr[:, 2]).astype(i32))
    d0 = p_loc // g12
    rem = p_loc - d0 * g12
    d1 = rem // g2
    d2 = rem - d1 * g2
    tcp = rep_p(tc_arr.astype(i32))
    patch_row = rep_p(row_start.astype(i32)) + tcp + p_loc

    enc_src = np.full(plan.M, -1, dtype=np.int32)
    dec_src = np.full(plan.M, -1, dtype=np.int32)
    enc_src[patch_row] = np.arange(plan.G, dtype=np.int32)
    dec_src[latent_row] = np.arange(plan.T, dtype=np.int32)

    p0, p1, p2 = patch_size
    W_ = rep_p(px[:, 2])
    HW_ = rep_p(px[:, 1] * px[:, 2])
    geom = np.empty((plan.G, 4), dtype=np.int64)
    geom[:, 0] = rep_p(clip_off) + (d0 * p0) * HW_ + (d1 * p1) * W_ + d2 * p2
    geom[:, 1] = W_
    geom[:, 2] = HW_
    geom[:, 3] = rep_p(px[:, 0] * px[:, 1] * px[:, 2])

    rope_pos = np.empty((plan.M, 3), dtype=np.int32)
    rope_pos[latent_row] = t_loc[:, None]                 # latent j -> (j, j, j)
    rope_pos[patch_row, 0] = d0 + tcp                      # patch (a, b, c) -> (a, b, c) + token_count
    rope_pos[patch_row, 1] = d1 + tcp
    rope_pos[patch_row, 2] = d2 + tcp
    plan.enc_src_row, plan.dec_src_row = enc_src, dec_src
    plan.latent_row, plan.patch_row, plan.geom, plan.rope_pos = latent_row, patch_row, geom, rope_pos
    return plan


def cos_sin_id_table(n_ids: int) -> np.ndarray:
    """float32 [n, F, 2] (cos, sin) of inv_freq[f] * id for id < n (n >= n_ids): the only trigonometry of a plan."""
    return _cos_sin_by_id(n_ids, rope_inv_freqs())


def get_attn_work(plan: PackedPlan, hq: int, hkv: int) -> np.ndarray:
    k = (hq, hkv)
    if k not in plan.attn_work:
        plan.attn_work[k] = attn_work_list(plan.cu_seqlens[:-1].tolist(), plan.seq_lens, hq, hkv)
    return plan.attn_work[k]


def get_attn_work_latent(plan: PackedPlan, hq: int, hkv: int) -> np.ndarray:
    """Work list of the encoder's LAST layer: only the query tiles that hold latent rows (they lead every clip)."""
    k = ("latent", hq, hkv)
    if k not in plan.attn_work:
        if (hq // hkv) % 2 == 0 and len(plan.token_counts):
            # one record per (row tile, pair of heads): the latent list is a row filter of the full list (same order)
            full = get_attn_work(plan, hq, hkv)
            tok = np.asarray(plan.token_counts, dtype=np.int64)[np.searchsorted(plan.cu_seqlens[:-1], full[:, 7])]
            plan.attn_work[k] = _with_leaders(np.ascontiguousarray(full[(full[:, 0] - full[:, 7]) < tok]))
        else:
            plan.attn_work[k] = attn_work_list(plan.cu_seqlens[:-1].tolist(), plan.seq_lens, hq, hkv, q_lens=plan.token_counts)
    return plan.attn_work[k]


def attn_bwd_work_lists(seq_starts: Sequence[int], seq_lens: Sequence[int], hq: int, hkv: int,
                        q_lens: Sequence[int] = None):
    """Work lists of the attention backward kernels (csrc/attn_bwd.cu), int32 [n, 8] records
    {st_row0, st_valid, st_head, o_head0, n_heads, clip_row0, clip_len, 0}, longest work first:
      dkv: one record per (128-key tile, kv head); streams the clip's query tiles of the `hq // hkv` grouped query heads
      dq : one record per (128-row query tile, query head); streams the clip's key tiles of kv head `h // (hq // hkv)`.
    `q_lens` (optional, per clip): only a clip's first `q_lens[i]` rows carry an output gradient (the encoder's last layer:
    the head reads the latent rows alone, blocks.py:101). dkv then streams just those query rows (its records' `clip_len`
    is the streamed range) and clips without such rows get no record (their dK / dV are zero); dq keeps only the query
    tiles that hold such rows (they stream all keys of the clip)."""
    grp = hq // hkv
    st = np.asarray(seq_starts, dtype=np.int64).reshape(-1)
    sl = np.asarray(seq_lens, dtype=np.int64).reshape(-1)
    if len(sl) == 0:
        return np.zeros((0, 8), dtype=np.int32), np.zeros((0, 8), dtype=np.int32)
    ql = sl if q_lens is None else np.minimum(np.asarray(q_lens, dtype=np.int64).reshape(-1), sl)
    # vectorised over all 128-row tiles of the batch (a ragged stream builds these lists every step)
    nt = (sl + ATTN_TILE - 1) // ATTN_TILE
    clip = np.repeat(np.arange(len(sl)), nt)
    ti = np.arange(int(nt.sum()), dtype=np.int64) - np.repeat(np.cumsum(nt) - nt, nt)
    row0 = st[clip] + ti * ATTN_TILE
    valid = np.minimum(ATTN_TILE, sl[clip] - ti * ATTN_TILE)

    def records(sel, n_heads_out, other_of, n_heads, stream_len):
        """one record per (selected tile, head < n_heads_out); other_of(head) -> o_head0; stream_len: per clip"""
        n_t = int(sel.sum())
        rec = np.zeros((n_t * n_heads_out, 8), dtype=np.int32)
        rep = lambda v: np.repeat(v[sel], n_heads_out)
        head = np.tile(np.arange(n_heads_out), n_t)
        rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3], rec[:, 4] = rep(row0), rep(valid), head, other_of(head), n_heads
        rec[:, 5], rec[:, 6] = rep(st[clip]), rep(stream_len[clip])
        return rec

    a = records(ql[clip] > 0, hkv, lambda h: h * grp, grp, ql)
    b = records(ti * ATTN_TILE < ql[clip], hq, lambda h: h // grp, 1, sl)
    a = a[np.argsort(-(a[:, 4].astype(np.int64) * ((a[:, 6] + 127) // 128)), kind="stable")]
    b = b[np.argsort(-((b[:, 6] + 127) // 128), kind="stable")]
    return np.ascontiguousarray(a), np.ascontiguousarray(b)


def get_attn_bwd_work(plan: PackedPlan, hq: int, hkv: int):
    k = ("bwd", hq, hkv)
    if k not in plan.attn_work:
        plan.attn_work[k] = attn_bwd_work_lists(plan.cu_seqlens[:-1].tolist(), plan.seq_lens, hq, hkv)
    return plan.attn_work[k]


def get_attn_bwd_work_latent(plan: PackedPlan, hq: int, hkv: int):
    """(dkv, dq) lists of the encoder's LAST layer in training: only the latent rows carry an output gradient."""
    k = ("bwd_latent", hq, hkv)
    if k not in plan.attn_work:
        plan.attn_work[k] = attn_bwd_work_lists(plan.cu_seqlens[:-1].tolist(), plan.seq_lens, hq, hkv, q_lens=plan.token_counts)
    return plan.attn_work[k]


def shard_clips(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Greedy longest-processing-time partition of clip indices over ranks (balance by cost).
    Returns, per rank, the sorted list of clip indices it owns."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    loads = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda j: (loads[j], j))
        out[r].append(i)
        loads[r] += costs[i]
    return [sorted(o) for o in out]


def clip_cost(grid_px: Sequence[int], token_count: int, patch_size: Sequence[int], width: int, layers: int) -> float:
    """Forward FLOPs of one clip through one stack: linear part + quadratic attention part."""
    g = math.prod(v // p for v, p in zip(grid_px, patch_size))
    s = g + token_count
    return layers * (s * 1605632.0 * (width / 256.0) ** 2 + 4.0 * s * s * width)
