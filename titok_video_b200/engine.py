"""Launch sequences of the sm_100a kernels for the TiTok encoder / decoder stacks.

This is the only place that talks to libtitok_b200.so. PyTorch is used for device memory, the current CUDA
stream and (optionally) CUDA-graph capture; all arithmetic happens in the library's kernels.

Per forward the sequence is (reference call stack: model/titok.py:68-74 -> model/base/blocks.py:71-104,148-177
-> model/base/transformer.py:126-146):

  encoder: patchify -> proj_in GEMM -> embed rows (+pre-norm) -> L x [qkv GEMM+RoPE, attention*sigmoid(gate),
           out_proj GEMM (+residual/KEEL+next norm), w12 GEMM+GEGLU, w3 GEMM (+residual/KEEL+next norm)]
           -> latent head + FSQ
  decoder: embed rows from codes -> L x [...] -> proj_out GEMM -> unpatchify
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .plan import PackedPlan, get_attn_bwd_work, get_attn_bwd_work_latent, get_attn_work, get_attn_work_latent, make_plan

_vp = ctypes.c_void_p


def _ptr(t: Optional[torch.Tensor]):
    return _vp(t.data_ptr()) if t is not None else _vp(0)


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def require_cuda(device: torch.device) -> None:
    if device.type != "cuda":
        raise _lib.TitokB200Error(
            f"titok_video_b200 runs on CUDA sm_100a devices only (got tensors on '{device}'); there is no CPU path"
        )


# --------------------------------------------------------------------------------------------------
# device-resident plan (metadata uploaded once) + workspace
# --------------------------------------------------------------------------------------------------
class DevicePlan:
    """Device copy of a PackedPlan. All integer metadata travels in ONE pinned staging buffer and one asynchronous
    H2D copy (a ragged stream builds a new plan every step); the fp32 RoPE table is gathered on the device."""

    def __init__(self, plan: PackedPlan, device: torch.device):
        self.plan = plan
        self.device = device
        pieces = {
            "clip_desc": plan.clip_desc,
            "clip_offset": np.asarray(plan.clip_offset, dtype=np.int64),
            "clip_numel": np.asarray(plan.clip_numel, dtype=np.int64),
        }
        offs, total = {}, 0
        for k, a in pieces.items():
            offs[k] = total
            total += (a.nbytes + 255) // 256 * 256
        total = max(total, 256)
        host, done = _staging(total)
        hv = host.numpy()
        for k, a in pieces.items():
            hv[offs[k]:offs[k] + a.nbytes] = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        self._meta = torch.empty(total, dtype=torch.uint8, device=device)
        self._meta.copy_(host[:total], non_blocking=True)
        done.record()
        for k, a in pieces.items():
            setattr(self, k, self._meta[offs[k]:offs[k] + a.nbytes].view(torch.int64).view(a.shape))
        # per-row metadata: expanded on the device from the per-clip descriptors (the host only did O(B) work)
        M, T, G = plan.M, plan.T, plan.G
        i32 = dict(dtype=torch.int32, device=device)
        self.enc_src_row = torch.empty(max(M, 1), **i32)
        self.dec_src_row = torch.empty(max(M, 1), **i32)
        self.latent_row = torch.empty(max(T, 1), **i32)
        self.patch_row = torch.empty(max(G, 1), **i32)
        self.geom = torch.empty((max(G, 1), 4), dtype=torch.int64, device=device)
        self.rope_pos = torch.empty((max(M, 1), 3), **i32)
        P0, P1, P2 = plan.patch_size
        _lib.call("ttk_build_plan", _ptr(self.clip_desc), len(plan.token_counts), M, P0, P1, P2, _ptr(self.enc_src_row),
                  _ptr(self.dec_src_row), _ptr(self.latent_row), _ptr(self.patch_row), _ptr(self.geom), _ptr(self.rope_pos),
                  _stream())
        # RoPE table: gathered on the device from the integer position ids and the shared per-id cos / sin table
        cs, n_ids = _cs_table(device, plan.max_pos + 1)
        self.rope = torch.empty((max(M, 1), 60), dtype=torch.float32, device=device)
        _lib.call("ttk_rope_table_gather", _ptr(self.rope_pos), _ptr(cs), n_ids, _ptr(self.rope), M, _stream())
        self._attn: Dict[Tuple[int, int], torch.Tensor] = {}
        self.ws: Dict[str, torch.Tensor] = {}
        self.graphs: Dict[tuple, "torch.cuda.CUDAGraph"] = {}

    def attn_work(self, hq: int, hkv: int) -> torch.Tensor:
        return self._attn_lists(hq, hkv)[0]

    def attn_work_latent(self, hq: int, hkv: int) -> torch.Tensor:
        """Work list of the encoder's last layer: only the query tiles that hold latent rows (plan.get_attn_work_latent)."""
        return self._attn_lists(hq, hkv)[1]

    def _attn_lists(self, hq: int, hkv: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(all query tiles, query tiles that hold latent rows): both lists travel in one staging buffer / one copy."""
        k = (hq, hkv)
        if k not in self._attn:
            full, lat = get_attn_work(self.plan, hq, hkv), get_attn_work_latent(self.plan, hq, hkv)
            both = self._upload_i32(np.concatenate([full, lat], axis=0))
            self._attn[k] = (both[:full.shape[0]], both[full.shape[0]:])
        return self._attn[k]

    def _upload_i32(self, w: np.ndarray) -> torch.Tensor:
        w = np.ascontiguousarray(w)
        host, done = _staging(max(w.nbytes, 256))
        host.numpy()[:w.nbytes] = w.view(np.uint8).reshape(-1)
        dev = torch.empty(max(w.nbytes, 256), dtype=torch.uint8, device=self.device)
        dev.copy_(host[:dev.numel()], non_blocking=True)
        done.record()
        return dev[:w.nbytes].view(torch.int32).view(w.shape)

    def attn_bwd_work(self, hq: int, hkv: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(dkv, dq) work lists of the attention backward kernels (plan.attn_bwd_work_lists)."""
        k = ("bwd", hq, hkv)
        if k not in self._attn:
            a, b = get_attn_bwd_work(self.plan, hq, hkv)
            self._attn[k] = (self._upload_i32(a), self._upload_i32(b))
        return self._attn[k]

    def attn_bwd_work_latent(self, hq: int, hkv: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(dkv, dq) work lists of the encoder's last layer when only the latent rows carry a gradient."""
        k = ("bwd_latent", hq, hkv)
        if k not in self._attn:
            a, b = get_attn_bwd_work_latent(self.plan, hq, hkv)
            self._attn[k] = (self._upload_i32(a), self._upload_i32(b))
        return self._attn[k]

    def buf(self, name: str, shape: Sequence[int], dtype=torch.bfloat16) -> torch.Tensor:
        """Named workspace tensor: a typed view of the device-wide arena of that name (see _arena). Plans of different
        batch compositions share the arenas, so a ragged stream does not allocate device memory per step; results that
        live in workspace buffers are therefore only valid until the next launch sequence on the device."""
        shape = tuple(int(v) for v in shape)
        key = (name, shape, dtype)
        hit = self.ws.get(key)
        if hit is not None and hit[0] == _ARENA_GEN:
            return hit[1]
        nbytes = max(1, math.prod(shape)) * _ELEM_SIZE[dtype]
        base = _arena(self.device, name, nbytes)
        t = base[:nbytes].view(dtype)[:math.prod(shape)].view(shape)
        self.ws[key] = (_ARENA_GEN, t)
        return t


_ELEM_SIZE = {torch.bfloat16: 2, torch.float16: 2, torch.float32: 4, torch.float64: 8, torch.int32: 4, torch.int64: 8,
              torch.uint8: 1, torch.int8: 1, torch.int16: 2, torch.bool: 1}
_ARENA: Dict[Tuple[str, str], torch.Tensor] = {}
_ARENA_GEN = 0  # bumped whenever an arena is re-allocated: cached views and captured CUDA graphs of older generations are stale


def _arena(device: torch.device, name: str, nbytes: int) -> torch.Tensor:
    global _ARENA_GEN
    k = (str(device), name)
    t = _ARENA.get(k)
    if t is None or t.numel() < nbytes:
        cap = max(nbytes, int(1.5 * t.numel()) if t is not None else 0)
        cap = (cap + (1 << 20) - 1) >> 20 << 20
        t = torch.empty(cap, dtype=torch.uint8, device=device)
        _ARENA[k] = t
        _ARENA_GEN += 1
    return t


def arena_generation() -> int:
    return _ARENA_GEN


# Ring of pinned staging buffers for plan metadata (pinning memory costs milliseconds, so it is done once): a slot is
# reused only after the copy that last read it has finished.
_STAGING: List[list] = []
_STAGING_NEXT = 0
_STAGING_SLOTS = 8


def _staging(nbytes: int):
    global _STAGING_NEXT
    if not _STAGING:
        for _ in range(_STAGING_SLOTS):
            _STAGING.append([torch.empty(1 << 22, dtype=torch.uint8).pin_memory(), torch.cuda.Event()])
    slot = _STAGING[_STAGING_NEXT]
    _STAGING_NEXT = (_STAGING_NEXT + 1) % _STAGING_SLOTS
    slot[1].synchronize()
    if slot[0].numel() < nbytes:
        slot[0] = torch.empty(1 << max(22, (nbytes - 1).bit_length()), dtype=torch.uint8).pin_memory()
    return slot[0], slot[1]


_CS_TABLES: Dict[str, Tuple[torch.Tensor, int]] = {}


def _cs_table(device: torch.device, n_ids: int) -> Tuple[torch.Tensor, int]:
    """Device copy of plan.cos_sin_id_table (fp32 [n,10,2]), shared by every plan on that device."""
    from .plan import cos_sin_id_table

    cur = _CS_TABLES.get(str(device))
    if cur is None or cur[1] < n_ids:
        t = cos_sin_id_table(n_ids)
        cur = (torch.from_numpy(np.ascontiguousarray(t)).to(device), int(t.shape[0]))
        _CS_TABLES[str(device)] = cur
    return cur


_PLAN_CACHE: Dict[tuple, DevicePlan] = {}
_PLAN_CACHE_MAX = 64


def get_device_plan(grids_px, token_counts, patch_size, channels, device) -> DevicePlan:
    key = (tuple(tuple(int(v) for v in g) for g in grids_px), tuple(int(t) for t in token_counts),
           tuple(int(p) for p in patch_size), int(channels), str(device))
    dp = _PLAN_CACHE.get(key)
    if dp is None:
        if len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
        dp = DevicePlan(make_plan(key[0], key[1], key[2], key[3], arrays=False), device)
        _PLAN_CACHE[key] = dp
    return dp


def clear_caches() -> None:
    _PLAN_CACHE.clear()  # (the workspace arenas stay: they are sized by the largest batch seen, not per plan)
    _BUCKET_CACHE.clear()


# --------------------------------------------------------------------------------------------------
# shape buckets: ONE captured launch sequence for every batch composition that fits a bucket
# --------------------------------------------------------------------------------------------------
class _MaxPlan:
    """The extents a bucket's launch sequence runs with (what DevicePlan.plan is to the launch functions)."""

    def __init__(self, G: int, T: int, B: int, patch_size, channels: int, feat: int):
        self.G, self.T, self.M, self.B = G, T, G + T, B
        self.patch_size, self.channels = tuple(patch_size), channels
        self.pixels = G * feat                # every patch is `feat` pixels-times-channels: sum 3*T*H*W == G * feat
        self.total_numel = self.pixels + feat  # + the scratch patch of the padded patches
        self.clip_numel = (self.pixels,) * B  # B entries (their maximum sizes the grid of ttk_clip_error)


class BucketPlan:
    """Device metadata + workspace of a BUCKET of batch compositions: any ragged batch with at most G patches, T latent
    tokens and B clips (the reference's dataloader draws a new composition every step, video_dataset.py:130-172).

    Extents are fixed per bucket, so the whole encoder -> FSQ -> decoder launch sequence is captured into ONE CUDA graph
    per (model, bucket) and replayed for every composition that falls into the bucket: per step the host does the O(B)
    planning, packs [sizes | per-clip descriptors | attention work list | clip offsets] into one pinned buffer, issues ONE
    H2D copy and one graph launch. The per-row metadata is expanded on the device INSIDE the graph
    (ttk_build_plan_bucket: real sizes from device memory, harmless defaults for the padding), the RoPE table is gathered
    from it. Padded rows never touch real rows (every kernel is row-wise; attention follows its work list, whose padded
    records are empty), so results are bit-identical to the per-composition path."""

    G_STEP, T_STEP, B_STEP = 2048, 256, 8

    def __init__(self, G_max: int, T_max: int, B_max: int, patch_size, channels: int, device: torch.device, heads):
        P0, P1, P2 = patch_size
        feat = channels * P0 * P1 * P2
        self.plan = _MaxPlan(G_max, T_max, B_max, patch_size, channels, feat)
        self.device = device
        self.heads = tuple(heads)
        M = self.plan.M
        hq, hkv = self.heads
        # upper bound of the attention work list: one record per (row tile, pair of query heads) -- or per (head, pair of
        # row tiles) when the kv group is odd -- and at most M/128 + B row tiles in any composition
        tiles = M // 128 + B_max + 1
        self.W_max = tiles * (hq // 2 if (hq // hkv) % 2 == 0 else hq)
        i32 = dict(dtype=torch.int32, device=device)
        self.enc_src_row = torch.empty(M, **i32)
        self.dec_src_row = torch.empty(M, **i32)
        self.latent_row = torch.zeros(max(T_max, 1), **i32)
        self.patch_row = torch.zeros(max(G_max, 1), **i32)
        self.geom = torch.zeros((max(G_max, 1), 4), dtype=torch.int64, device=device)
        self.rope_pos = torch.zeros((M, 3), **i32)
        self.rope = torch.empty((M, 60), dtype=torch.float32, device=device)
        # the last encoder layer's list (query tiles that hold latent rows): at most T/128 + B row tiles
        self.W_lat_max = (T_max // 128 + B_max + 1) * (hq // 2 if (hq // hkv) % 2 == 0 else hq)
        # [hdr 8 | desc B*12 | clip_offset B | clip_numel B] int64, then the work lists int32 [W_max, 12], [W_lat_max, 12]
        self.n_i64 = 8 + 14 * B_max
        self.meta_bytes = (self.n_i64 * 8 + (self.W_max + self.W_lat_max) * 48 + 255) // 256 * 256
        self._meta = torch.zeros(self.meta_bytes, dtype=torch.uint8, device=device)
        m64 = self._meta[:self.n_i64 * 8].view(torch.int64)
        self.hdr = m64[:8]
        self.clip_desc = m64[8:8 + 12 * B_max].view(B_max, 12)
        self.clip_offset = m64[8 + 12 * B_max:8 + 13 * B_max]
        self.clip_numel = m64[8 + 13 * B_max:8 + 14 * B_max]
        self._work = self._meta[self.n_i64 * 8:self.n_i64 * 8 + self.W_max * 48].view(torch.int32).view(self.W_max, 12)
        o_lat = self.n_i64 * 8 + self.W_max * 48
        self._work_lat = self._meta[o_lat:o_lat + self.W_lat_max * 48].view(torch.int32).view(self.W_lat_max, 12)
        self.ws: Dict[str, torch.Tensor] = {}
        self.graphs: Dict[tuple, tuple] = {}
        self.cs, self.n_ids = _cs_table(device, 1024)  # position ids are < token_count + max grid side << 1024

    buf = DevicePlan.buf

    def attn_work(self, hq: int, hkv: int) -> torch.Tensor:
        assert (hq, hkv) == self.heads
        return self._work

    def attn_work_latent(self, hq: int, hkv: int) -> torch.Tensor:
        assert (hq, hkv) == self.heads
        return self._work_lat

    def build_launch(self) -> None:
        """The metadata expansion + RoPE gather: the first two launches of the captured sequence."""
        pl = self.plan
        P0, P1, P2 = pl.patch_size
        st = _stream()
        _lib.call("ttk_build_plan_bucket", _ptr(self.clip_desc), _ptr(self.hdr), pl.M, pl.T, pl.G, P0, P1, P2,
                  _ptr(self.enc_src_row), _ptr(self.dec_src_row), _ptr(self.latent_row), _ptr(self.patch_row), _ptr(self.geom),
                  _ptr(self.rope_pos), st)
        _lib.call("ttk_rope_table_gather", _ptr(self.rope_pos), _ptr(self.cs), self.n_ids, _ptr(self.rope), pl.M, st)

    def upload(self, plan: PackedPlan) -> None:
        """This step's composition: one pinned buffer, one asynchronous copy."""
        mp = self.plan
        B = len(plan.token_counts)
        if plan.G > mp.G or plan.T > mp.T or B > mp.B:
            raise _lib.TitokB200Error("batch does not fit this bucket")
        if plan.max_pos + 1 > self.n_ids:
            raise _lib.TitokB200Error("position ids exceed the bucket's RoPE table")
        work = get_attn_work(plan, *self.heads)
        work_lat = get_attn_work_latent(plan, *self.heads)
        if work.shape[0] > self.W_max or work_lat.shape[0] > self.W_lat_max:
            raise _lib.TitokB200Error("attention work list exceeds the bucket's bound")
        host, done = _staging(self.meta_bytes)
        hv = host.numpy()
        hv[:self.meta_bytes] = 0
        h64 = hv[:self.n_i64 * 8].view(np.int64)
        h64[0], h64[1], h64[2], h64[3], h64[4] = B, plan.M, plan.T, plan.G, mp.pixels
        h64[8:8 + 12 * B] = plan.clip_desc.reshape(-1)
        h64[8 + 12 * mp.B:8 + 12 * mp.B + B] = plan.clip_offset
        h64[8 + 13 * mp.B:8 + 13 * mp.B + B] = plan.clip_numel
        hw = hv[self.n_i64 * 8:self.n_i64 * 8 + work.size * 4].view(np.int32)
        hw[:] = work.reshape(-1)
        o_lat = self.n_i64 * 8 + self.W_max * 48
        hv[o_lat:o_lat + work_lat.size * 4].view(np.int32)[:] = work_lat.reshape(-1)
        self._meta.copy_(host[:self.meta_bytes], non_blocking=True)
        done.record()


_BUCKET_CACHE: Dict[tuple, BucketPlan] = {}


def _round_up(v: int, step: int) -> int:
    return max(step, (v + step - 1) // step * step)


def get_bucket_plan(plan: PackedPlan, device, heads) -> BucketPlan:
    key = (_round_up(plan.G, BucketPlan.G_STEP), _round_up(plan.T, BucketPlan.T_STEP),
           _round_up(len(plan.token_counts), BucketPlan.B_STEP), tuple(plan.patch_size), plan.channels, str(device), tuple(heads))
    bp = _BUCKET_CACHE.get(key)
    if bp is None:
        bp = BucketPlan(key[0], key[1], key[2], plan.patch_size, plan.channels, device, heads)
        _BUCKET_CACHE[key] = bp
    return bp


# --------------------------------------------------------------------------------------------------
# prepared weights: bf16 GEMM operands, fp32 norm weights; refreshed in place when parameters change
# --------------------------------------------------------------------------------------------------
_PERM_CACHE: Dict[tuple, torch.Tensor] = {}


def patch_feature_perm_on(patch_size: Sequence[int], channels: int, device) -> torch.Tensor:
    """patch_feature_perm as a tensor on `device`, cached (an H2D copy per call would stall the launch stream)."""
    key = (tuple(int(p) for p in patch_size), int(channels), str(device))
    t = _PERM_CACHE.get(key)
    if t is None:
        t = patch_feature_perm(patch_size, channels).to(device)
        _PERM_CACHE[key] = t
    return t


def patch_feature_perm(patch_size: Sequence[int], channels: int) -> torch.Tensor:
    """perm[j_new] = j_ref with j_ref = ((p0*P1+p1)*P2+p2)*C + c (einops '(p0 p1 p2 c)', utils.py:26-34) and
    j_new = ((c*P0+p0)*P1+p1)*P2+p2 (what patchify / unpatchify move as 16-byte runs)."""
    P0, P1, P2 = patch_size
    j = torch.arange(channels * P0 * P1 * P2)
    p2 = j % P2
    p1 = (j // P2) % P1
    p0 = (j // (P2 * P1)) % P0
    c = j // (P2 * P1 * P0)
    return ((p0 * P1 + p1) * P2 + p2) * channels + c


def cached_named_params(module) -> list:
    """[(qualified name, parameter)] of a module, cached on it: walking the module tree costs ~50 us per call and the launch
    sequences need the list several times per step. The cache is validated against the owners' `_parameters` dicts (76
    identity checks), so replacing a parameter object (not just its data) is noticed."""
    c = module.__dict__.get("_ttk_param_cache")
    if c is not None and all(owner._parameters.get(leaf) is p for owner, leaf, _, p in c):
        return [(name, p) for _, _, name, p in c]
    c = []
    for mod_name, sub in module.named_modules():
        for leaf, p in sub._parameters.items():
            if p is not None:
                c.append((sub, leaf, (mod_name + "." if mod_name else "") + leaf, p))
    module.__dict__["_ttk_param_cache"] = c
    return [(name, p) for _, _, name, p in c]


class PreparedStack:
    """bf16 / fp32 device copies of one TiTokEncoder / TiTokDecoder's parameters in kernel layout."""

    def __init__(self, module, kind: str):
        self.kind = kind  # 'enc' | 'dec'
        self.module = module
        self.sig = None
        self.t: Dict[str, torch.Tensor] = {}
        self._pending_dst: List[torch.Tensor] = []
        self._pending_src: List[torch.Tensor] = []
        self._all_dst: List[torch.Tensor] = []
        self._all_src: List[torch.Tensor] = []

    def _signature(self):
        return tuple((p.data_ptr(), p._version, p.dtype) for _, p in cached_named_params(self.module))

    def _set(self, name: str, value: torch.Tensor, dtype) -> None:
        cur = self.t.get(name)
        if cur is not None and cur.shape == value.shape and cur.dtype == dtype and cur.device == value.device:
            # keep the address stable (captured CUDA graphs stay valid); the copies of one refresh are batched into a few
            # multi-tensor kernels (a refresh follows every optimizer step: 38 parameters per stack)
            self._pending_dst.append(cur)
            self._pending_src.append(value.detach())
        else:
            self.t[name] = value.detach().to(dtype, copy=True).contiguous()
            self._table = None  # an address changed: the native sequencers' pointer table is stale
        self._all_dst.append(self.t[name])
        self._all_src.append(value.detach())

    @staticmethod
    def _cast_table(dst: List[torch.Tensor], src: List[torch.Tensor]):
        """(device int64 [n,4] table, n, max numel) for ttk_multi_cast, or None when a dtype / layout is outside what the
        kernel handles (then the refresh uses torch's multi-tensor copy)."""
        rows = []
        for d, s_ in zip(dst, src):
            if s_.dtype not in (torch.float32, torch.bfloat16) or d.dtype not in (torch.float32, torch.bfloat16):
                return None
            if not (s_.is_contiguous() and d.is_contiguous()) or s_.numel() != d.numel() or not s_.is_cuda:
                return None
            rows.append([s_.data_ptr(), d.data_ptr(), d.numel(), 2 * int(s_.dtype == torch.bfloat16) + int(d.dtype == torch.float32)])
        if not rows:
            return None
        tab = torch.from_numpy(np.asarray(rows, dtype=np.int64)).to(dst[0].device)
        return tab, len(rows), max(r[2] for r in rows)

    def layer_table(self) -> np.ndarray:
        """HOST int64 [n_layers, 9] device pointers for the native sequencers (include/titok_b200.h: ttk_layers_desc)."""
        if self.__dict__.get("_table") is None:
            L = self.module.num_layers
            tab = np.zeros((L, 9), dtype=np.int64)
            for i in range(L):
                nxt = self.t[f"pre_ln{i + 1}"] if i + 1 < L else self.t["ln_post"]
                row = [self.t[f"to_qkv{i}"], self.t[f"out_proj{i}"], self.t[f"w12_{i}"], self.t[f"w3_{i}"], self.t[f"pre_ln{i}"],
                       self.t.get(f"attn_post_ln{i}"), self.t[f"ffn_norm{i}"], self.t.get(f"ffd_post_ln{i}"), nxt]
                tab[i] = [0 if t is None else t.data_ptr() for t in row]
            self._table = tab
        return self._table

    @torch.no_grad()
    def refresh(self, force: bool = False) -> "PreparedStack":
        """Bring the kernel-layout copies up to date with the module's parameters.

        Staleness cannot be detected from the parameters alone: fused optimizer kernels (`AdamW(fused=True)`) and in-place
        edits through `p.data` do not bump `p._version`. So (a) every optimizer step in the process bumps a counter that
        is part of the signature (`_OPT_STEPS`, a global optimizer post-step hook), and (b) callers that are about to
        record a training step pass `force=True`. A refresh with unchanged parameter objects is cheap: two or three
        index_selects for the permuted patch projections and ONE multi-tensor copy for all 38 parameters of the stack."""
        sig = (self._signature(), _OPT_STEPS[0])
        if not force and sig == self.sig:
            return self
        fast = self.__dict__.get("_fast")
        key = tuple((id(p), p.data_ptr(), p.dtype) for _, p in cached_named_params(self.module))
        if fast is not None and fast[0] == key:
            _, gathers, dst, src, table = fast
            for w, dim, perm, tmp in gathers:
                torch.index_select(w, dim, perm, out=tmp)
            if table is not None:
                _lib.call("ttk_multi_cast", _ptr(table[0]), table[1], table[2], _stream())
            else:
                torch._foreach_copy_(dst, src)
            self.sig = sig
            return self
        m = self.module
        bf, f32 = torch.bfloat16, torch.float32
        perm = patch_feature_perm_on(m.patch_size_tuple, m.patch_channels, m.mask_token.device)
        self._pending_dst, self._pending_src, self._all_dst, self._all_src = [], [], [], []
        gathers = []  # (source parameter, dim, perm, persistent fp32 temporary) of the permuted patch projections

        def permuted(w: torch.Tensor, dim: int) -> torch.Tensor:
            tmp = torch.index_select(w.detach(), dim, perm)
            gathers.append((w.detach(), dim, perm, tmp))
            return tmp

        self._set("mask_token", m.mask_token.reshape(1), f32)
        self._set("ln_pre_t", m.ln_pre_t.weight, f32)
        self._set("ln_pre_p", m.ln_pre_p.weight, f32)
        self._set("ln_post", m.ln_post.weight, f32)
        if self.kind == "enc":
            self._set("proj_in_w", permuted(m.proj_in.weight, 1), bf)
            self._set("proj_in_b", m.proj_in.bias, bf)
            self._set("proj_out_w", m.proj_out.weight, bf)
            self._set("proj_out_b", m.proj_out.bias, bf)
        else:
            self._set("proj_in_w", m.proj_in.weight, bf)
            self._set("proj_in_b", m.proj_in.bias, bf)
            self._set("proj_out_w", permuted(m.proj_out.weight, 0), bf)
            self._set("proj_out_b", permuted(m.proj_out.bias, 0), bf)
        ml = m.model_layers
        for i in range(m.num_layers):
            a, f = ml.attn_layer[i], ml.ffd_layer[i]
            self._set(f"pre_ln{i}", a.pre_ln.weight, f32)
            self._set(f"to_qkv{i}", a.to_qkv.weight, bf)
            self._set(f"out_proj{i}", a.out_proj.weight, bf)
            self._set(f"ffn_norm{i}", f.norm.weight, f32)
            self._set(f"w12_{i}", f.w12.weight, bf)
            self._set(f"w3_{i}", f.w3.weight, bf)
            if i > 0:
                self._set(f"attn_post_ln{i}", ml.attn_post_ln[i - 1].weight, f32)
                self._set(f"ffd_post_ln{i}", ml.ffd_post_ln[i - 1].weight, f32)
        if self._pending_dst:
            torch._foreach_copy_(self._pending_dst, self._pending_src)
        # the (destination, source) pairs of this refresh are the recipe of every later one, as long as the parameter
        # objects and their storages stay the same (sources are aliases of the parameters / the persistent temporaries)
        if len(self._all_dst) == len(self.t):
            self._fast = (key, gathers, list(self._all_dst), list(self._all_src), self._cast_table(self._all_dst, self._all_src))
        else:
            self._fast = None
        self._pending_dst, self._pending_src, self._all_dst, self._all_src = [], [], [], []
        self.sig = sig
        return self


def prepared(module, kind: str, force: bool = False) -> PreparedStack:
    ps = module.__dict__.get("_ttk_prepared")
    if ps is None or ps.kind != kind:
        ps = PreparedStack(module, kind)
        module.__dict__["_ttk_prepared"] = ps
    return ps.refresh(force)


def invalidate(module: torch.nn.Module) -> None:
    """Force the next launch sequence of every stack under `module` to re-read its parameters. Needed only after editing
    parameters in a way PyTorch does not record (e.g. `p.data.add_(...)`) outside of a training step."""
    for sub in module.modules():
        ps = sub.__dict__.get("_ttk_prepared")
        if ps is not None:
            ps.sig = None


# every optimizer step in the process (fused kernels do not bump parameter versions) invalidates the prepared weights
_OPT_STEPS = [0]


def _count_optimizer_step(*_args, **_kwargs) -> None:
    _OPT_STEPS[0] += 1


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_opt_hook

    _reg_opt_hook(_count_optimizer_step)
except Exception:  # pragma: no cover  (very old torch: training forwards still refresh unconditionally)
    pass


# --------------------------------------------------------------------------------------------------
# launch sequences
# --------------------------------------------------------------------------------------------------
FUSE_RESID_256 = os.environ.get("TTK_FUSE_RESID", "1") != "0"
NATIVE_SEQ = os.environ.get("TTK_NATIVE_SEQ", "1") != "0"  # 0: enqueue the layers kernel by kernel from Python
# The encoder's last layer carries only the latent rows past its attention (blocks.py:101 reads `x[latent_mask]` alone;
# every other step of the layer is row-wise): 0 runs all rows through it, as the reference does -- same results bit for bit.
LATENT_TAIL = os.environ.get("TTK_LATENT_TAIL", "1") != "0"
# (test hook) 0: a bucket's launch sequence does not clear the attention output of its padded rows -- see _clear_padding
BUCKET_CLEAR_ATT = os.environ.get("TTK_BUCKET_CLEAR_ATT", "1") != "0"


def key_norms(dp: DevicePlan, M: int, hkv: int) -> torch.Tensor:
    """fp32 [kv heads, M] scratch: |k|^2 of every packed row, written by the qkv GEMM epilogue and read by the attention
    kernel of the same layer (its score bound)."""
    return dp.buf("knorm", (hkv, M), torch.float32)


def layers_desc(m, W: PreparedStack, dp: DevicePlan, M: int, backward: bool = False) -> "_lib.LayersDesc":
    """ttk_layers_desc for one stack and one packed batch (the weight table and the work lists stay referenced by W / dp)."""
    hq, hkv = m.heads
    work = dp.attn_work(hq, hkv)
    d = _lib.LayersDesc()
    d.M, d.width, d.gqa, d.inner, d.n_layers = M, m.width, hkv * 64, m.inner_dim, m.num_layers
    d.n_attn_work = work.shape[0]
    d.alpha, d.softmax_scale = float(2 * m.num_layers), 1.0 / math.sqrt(64.0)
    d.rope, d.attn_work = dp.rope.data_ptr(), work.data_ptr()
    d.k_norm2 = key_norms(dp, M, hkv).data_ptr()
    if backward:
        wk_dkv, wk_dq = dp.attn_bwd_work(hq, hkv)
        d.n_dkv_work, d.n_dq_work = wk_dkv.shape[0], wk_dq.shape[0]
        d.dkv_work, d.dq_work = wk_dkv.data_ptr(), wk_dq.data_ptr()
    d.weights = W.layer_table().ctypes.data
    return d


def _clear_padding(dp, M: int, w: int) -> None:
    """A bucket's launch sequence runs with the bucket's extents: packed rows past the step's real M are padding. No
    attention record covers them, so their rows of `att` are never written, and whatever the arena held there would reach
    `x` / `qkv` of the padded rows through out_proj. Padded rows never write real rows -- but the attention kernel READS
    them: the last 64-key box of the last real clip extends into the rows behind it, and although those keys are masked
    (P = 0), a NaN in their value rows survives the product (0 x NaN). In the per-composition path the rows behind the last
    clip are past the tensor map's extent (TMA fills zeros); here they are padded rows, so `att` is cleared once per launch
    sequence (one memset node in the captured graph), which keeps every padded row finite through both stacks."""
    if isinstance(dp, BucketPlan) and BUCKET_CLEAR_ATT:
        dp.buf("att", (M, w)).zero_()


def _layers(m, W: PreparedStack, dp: DevicePlan, x: torch.Tensor, xn: torch.Tensor, n_run: Optional[int] = None) -> None:
    """ResidualAttentionBlock.forward (transformer.py:126-146). x, xn are updated in place; on return xn holds
    RMSNorm(x) * ln_post.weight for every packed row. n_run: run only the first n_run layers (xn then holds the pre-norm
    of layer n_run)."""
    M, w = x.shape
    hq, hkv = m.heads
    gqa = hkv * 64
    inner = m.inner_dim
    L = m.num_layers
    alpha = float(2 * L)
    n_run = L if n_run is None else n_run
    if n_run <= 0:
        return
    st = _stream()
    qkv = dp.buf("qkv", (M, 2 * w + 2 * gqa))
    att = dp.buf("att", (M, w))
    h = dp.buf("h", (M, inner))
    y = None if (FUSE_RESID_256 and w == 256) else dp.buf("y", (M, w))
    work = dp.attn_work(hq, hkv)
    knorm = key_norms(dp, M, hkv)
    scale = 1.0 / math.sqrt(64.0)
    T = W.t
    if NATIVE_SEQ and not _lib.profiling():
        d = layers_desc(m, W, dp, M)
        d.n_layers = n_run
        _lib.call("ttk_layers_fwd", ctypes.byref(d), _ptr(x), _ptr(xn), _ptr(qkv), _ptr(att), _ptr(h), _ptr(y), st,
                  launches=n_run * (5 if y is None else 7))
        return

    def out_update(a: torch.Tensor, wmat: torch.Tensor, K: int, mode: int, w_post, w_next) -> None:
        if y is None:
            _lib.call("ttk_gemm_resid_norm256", _ptr(a), a.stride(0), _ptr(wmat), wmat.stride(0), M, K, _ptr(x),
                      x.stride(0), mode, alpha, _ptr(w_post), _ptr(w_next), _ptr(x), _ptr(xn), x.stride(0), st)
        else:
            _lib.call("ttk_gemm_bf16", _ptr(a), a.stride(0), _ptr(wmat), wmat.stride(0), M, w, K, _vp(0), _ptr(y),
                      y.stride(0), _vp(0), 0, st)
            _lib.call("ttk_resid_norm", _ptr(x), _ptr(y), _ptr(x), _ptr(xn), _ptr(w_post), _ptr(w_next), alpha, mode,
                      M, w, x.stride(0), st)

    for i in range(n_run):
        mode = 0 if i == 0 else 1
        _lib.call("ttk_gemm_qkv_rope", _ptr(xn), xn.stride(0), _ptr(T[f"to_qkv{i}"]), w, M, w, w, gqa, _ptr(dp.rope),
                  _ptr(qkv), qkv.stride(0), _ptr(knorm), st)
        _lib.call("ttk_attn_varlen_fwd", _ptr(qkv), qkv.stride(0), M, w, gqa, _ptr(work), work.shape[0], scale,
                  _ptr(att), att.stride(0), _ptr(knorm), st)
        out_update(att, T[f"out_proj{i}"], w, mode, T.get(f"attn_post_ln{i}"), T[f"ffn_norm{i}"])
        _lib.call("ttk_gemm_geglu", _ptr(xn), xn.stride(0), _ptr(T[f"w12_{i}"]), w, M, inner, w, _ptr(h), h.stride(0),
                  st)
        w_next = T[f"pre_ln{i + 1}"] if i + 1 < L else T["ln_post"]
        out_update(h, T[f"w3_{i}"], inner, mode, T.get(f"ffd_post_ln{i}"), w_next)


def _layer_latent(m, W: PreparedStack, dp: DevicePlan, x: torch.Tensor, xn: torch.Tensor, Tn: int) -> torch.Tensor:
    """The encoder's last layer on what the head reads. The qkv projection runs on all rows (every row is a key), attention
    only for the query tiles that hold latent rows, and everything behind it on the Tn latent rows gathered into compact
    [Tn, w] buffers. Returns xnc = RMSNorm(x_out[latent rows]) * ln_post.weight, bit-identical to those rows of _layers."""
    M, w = x.shape
    hq, hkv = m.heads
    gqa = hkv * 64
    inner = m.inner_dim
    L = m.num_layers
    i = L - 1
    alpha = float(2 * L)
    st = _stream()
    qkv = dp.buf("qkv", (M, 2 * w + 2 * gqa))
    att = dp.buf("att", (M, w))
    xc = dp.buf("x_lat", (Tn, w))
    xnc = dp.buf("xn_lat", (Tn, w))
    attc = dp.buf("att_lat", (Tn, w))
    hc = dp.buf("h_lat", (Tn, inner))
    yc = None if (FUSE_RESID_256 and w == 256) else dp.buf("y_lat", (Tn, w))
    work = dp.attn_work_latent(hq, hkv)
    knorm = key_norms(dp, M, hkv)
    scale = 1.0 / math.sqrt(64.0)
    T = W.t
    if NATIVE_SEQ and not _lib.profiling():
        d = layers_desc(m, W, dp, M)
        _lib.call("ttk_layer_fwd_latent", ctypes.byref(d), i, _ptr(x), _ptr(xn), _ptr(qkv), _ptr(att), _ptr(work),
                  work.shape[0], _ptr(dp.latent_row), Tn, _ptr(xc), _ptr(xnc), _ptr(attc), _ptr(hc), _ptr(yc), st,
                  launches=4 + (3 if yc is None else 5))
        return xnc
    mode = 0 if i == 0 else 1
    tag = "[latent rows]"  # (the per-kernel profiler keeps these launches apart from the full-size ones)

    def out_update(a: torch.Tensor, wmat: torch.Tensor, K: int, w_post, w_next) -> None:
        if yc is None:
            _lib.call("ttk_gemm_resid_norm256", _ptr(a), a.stride(0), _ptr(wmat), wmat.stride(0), Tn, K, _ptr(xc), w, mode,
                      alpha, _ptr(w_post), _ptr(w_next), _ptr(xc), _ptr(xnc), w, st, label="ttk_gemm_resid_norm256" + tag)
        else:
            _lib.call("ttk_gemm_bf16", _ptr(a), a.stride(0), _ptr(wmat), wmat.stride(0), Tn, w, K, _vp(0), _ptr(yc), w,
                      _vp(0), 0, st, label="ttk_gemm_bf16" + tag)
            _lib.call("ttk_resid_norm", _ptr(xc), _ptr(yc), _ptr(xc), _ptr(xnc), _ptr(w_post), _ptr(w_next), alpha, mode,
                      Tn, w, w, st, label="ttk_resid_norm" + tag)

    _lib.call("ttk_gemm_qkv_rope", _ptr(xn), xn.stride(0), _ptr(T[f"to_qkv{i}"]), w, M, w, w, gqa, _ptr(dp.rope),
              _ptr(qkv), qkv.stride(0), _ptr(knorm), st)
    _lib.call("ttk_attn_varlen_fwd", _ptr(qkv), qkv.stride(0), M, w, gqa, _ptr(work), work.shape[0], scale, _ptr(att),
              att.stride(0), _ptr(knorm), st, label="ttk_attn_varlen_fwd" + tag)
    _lib.call("ttk_gather_rows", _ptr(att), w, _ptr(dp.latent_row), _ptr(attc), w, Tn, w, st)
    _lib.call("ttk_gather_rows", _ptr(x), w, _ptr(dp.latent_row), _ptr(xc), w, Tn, w, st)
    out_update(attc, T[f"out_proj{i}"], w, T.get(f"attn_post_ln{i}"), T[f"ffn_norm{i}"])
    _lib.call("ttk_gemm_geglu", _ptr(xnc), w, _ptr(T[f"w12_{i}"]), w, Tn, inner, w, _ptr(hc), inner, st,
              label="ttk_gemm_geglu" + tag)
    out_update(hc, T[f"w3_{i}"], inner, T.get(f"ffd_post_ln{i}"), T["ln_post"])
    return xnc


def encoder_launch(m, dp: DevicePlan, clips_flat: torch.Tensor, fsq_consts) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """TiTokEncoder.forward (+ FSQ.forward) on a flat bf16 clip buffer. Returns (z, codes, indices) workspace
    tensors: z/codes bf16 [T, token_size], indices int32 [T]."""
    W = prepared(m, "enc")
    pl = dp.plan
    M, G, Tn, w = pl.M, pl.G, pl.T, m.width
    P0, P1, P2 = pl.patch_size
    feat = pl.channels * P0 * P1 * P2
    st = _stream()
    patches = dp.buf("patches", (G, feat))
    proj = dp.buf("proj", (G, w))
    x = dp.buf("x", (M, w))
    xn = dp.buf("xn", (M, w))
    ts = m.token_size
    z = dp.buf("z", (max(Tn, 1), ts))
    codes = dp.buf("codes", (max(Tn, 1), ts))
    idx = dp.buf("idx", (max(Tn, 1),), torch.int32)
    T = W.t
    # decoded uint8 frames are normalised inside the patch gather (no separate pass over the pixels)
    _lib.call("ttk_patchify_u8" if clips_flat.dtype == torch.uint8 else "ttk_patchify", _ptr(clips_flat), _ptr(dp.geom),
              pl.channels, P0, P1, P2, _ptr(patches), feat, G, st)
    _lib.call("ttk_gemm_bf16", _ptr(patches), feat, _ptr(T["proj_in_w"]), feat, G, w, feat, _ptr(T["proj_in_b"]),
              _ptr(proj), w, _vp(0), 0, st)
    _lib.call("ttk_enc_embed", _ptr(proj), w, _ptr(dp.enc_src_row), _ptr(T["mask_token"]), _ptr(T["ln_pre_t"]),
              _ptr(T["ln_pre_p"]), _ptr(T["pre_ln0"]), _ptr(x), _ptr(xn), M, w, w, st)
    _clear_padding(dp, M, w)
    half_l, offset, shift, half_width, basis, levels = fsq_consts
    if LATENT_TAIL and Tn > 0:
        # the head reads the latent rows only: the last layer carries nothing else past its attention
        _layers(m, W, dp, x, xn, n_run=m.num_layers - 1)
        head_in, head_map = _layer_latent(m, W, dp, x, xn, Tn), None
    else:
        _layers(m, W, dp, x, xn)
        head_in, head_map = xn, dp.latent_row
    _lib.call("ttk_enc_head_fsq", _ptr(head_in), w, _ptr(head_map), _ptr(T["ln_post"]), 1, _ptr(T["proj_out_w"]),
              _ptr(T["proj_out_b"]), ts, _ptr(z), _ptr(codes), _ptr(idx), Tn, w, half_l, offset, shift, half_width,
              basis, levels, st)
    return z[:Tn], codes[:Tn], idx[:Tn]


def decoder_launch(m, dp: DevicePlan, codes: torch.Tensor, out_flat: torch.Tensor) -> torch.Tensor:
    """TiTokDecoder.forward: codes bf16 [T, token_size] -> out_flat bf16 [sum 3*T*H*W]."""
    W = prepared(m, "dec")
    pl = dp.plan
    M, G, w = pl.M, pl.G, m.width
    P0, P1, P2 = pl.patch_size
    feat = pl.channels * P0 * P1 * P2
    st = _stream()
    x = dp.buf("x", (M, w))
    xn = dp.buf("xn", (M, w))
    rows = dp.buf("rows", (M, feat))
    T = W.t
    _lib.call("ttk_dec_embed", _ptr(codes), m.token_size, _ptr(dp.dec_src_row), _ptr(T["proj_in_w"]),
              _ptr(T["proj_in_b"]), _ptr(T["mask_token"]), _ptr(T["ln_pre_t"]), _ptr(T["ln_pre_p"]), _ptr(T["pre_ln0"]),
              _ptr(x), _ptr(xn), M, w, w, st)
    _clear_padding(dp, M, w)
    _layers(m, W, dp, x, xn)
    _lib.call("ttk_gemm_bf16", _ptr(xn), w, _ptr(T["proj_out_w"]), w, M, feat, w, _ptr(T["proj_out_b"]), _ptr(rows),
              feat, _vp(0), 0, st)
    _lib.call("ttk_unpatchify", _ptr(rows), feat, _ptr(dp.patch_row), _ptr(dp.geom), pl.channels, P0, P1, P2,
              _ptr(out_flat), G, st)
    return out_flat


def clip_error_launch(dp: DevicePlan, a_flat: torch.Tensor, b_flat: torch.Tensor) -> torch.Tensor:
    """fp64 [B, 2] per-clip (sum |a-b|, sum (a-b)^2) of two flat clip buffers laid out like the plan's clips: the
    numerators of the L1 reconstruction loss (loss_module.py:118) and of PSNR (eval_metrics.py)."""
    B = len(dp.plan.clip_numel)
    out = dp.buf("clip_err", (B, 2), torch.float64)
    out.zero_()
    _lib.call("ttk_clip_error", _ptr(a_flat), _ptr(b_flat), _ptr(dp.clip_offset), _ptr(dp.clip_numel), B,
              max(dp.plan.clip_numel), _ptr(out), _stream())
    return out


# --------------------------------------------------------------------------------------------------
# helpers for the module layer
# --------------------------------------------------------------------------------------------------
def to_host_ints(v) -> List[int]:
    """token_counts / grids as Python ints. A CUDA tensor costs one device sync (the reference syncs many
    times per forward here, blocks.py:85-86); pass CPU tensors or lists to avoid it."""
    if isinstance(v, torch.Tensor):
        return v.detach().cpu().tolist()
    return [(to_host_ints(e) if isinstance(e, (list, tuple, torch.Tensor)) else int(e)) for e in v]


def flatten_clips(videos: Sequence[torch.Tensor], dp: DevicePlan, name: str = "clips_in", keep_u8: bool = False) -> torch.Tensor:
    """Concatenate the clips into the plan's static flat bf16 input buffer (one cat kernel). uint8 clips (decoded
    frames, values 0..255) are normalised to [-1, 1] by ttk_normalize_u8 on the way -- or, with keep_u8, returned as the
    flat uint8 buffer for the inference encoder launch, whose patch gather normalises them itself (ttk_patchify_u8)."""
    if videos[0].dtype == torch.uint8:
        # decoded frames: one cat into the uint8 staging buffer, normalised on the device (dataset/video_dataset.py:118-119)
        raw = dp.buf(name + "_u8", (dp.plan.total_numel,), torch.uint8)
        if len(videos) == 1 and videos[0].is_contiguous():
            raw.copy_(videos[0].reshape(-1))
        else:
            torch.cat([v.reshape(-1) for v in videos], out=raw)
        if keep_u8:
            return raw
        buf = dp.buf(name, (dp.plan.total_numel,))
        _lib.call("ttk_normalize_u8", _ptr(raw), _ptr(buf), dp.plan.total_numel, _stream())
        return buf
    buf = dp.buf(name, (dp.plan.total_numel,))
    if len(videos) == 1 and videos[0].dtype == torch.bfloat16 and videos[0].is_contiguous():
        buf.copy_(videos[0].reshape(-1))
    else:
        torch.cat([v.reshape(-1).to(torch.bfloat16) for v in videos], out=buf)
    return buf


def split_clips(flat: torch.Tensor, plan: PackedPlan) -> List[torch.Tensor]:
    out = []
    for off, n, g in zip(plan.clip_offset, plan.clip_numel, plan.grids_px):
        out.append(flat[off:off + n].view(plan.channels, *g))
    return out
