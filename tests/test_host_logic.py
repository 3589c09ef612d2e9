"""Host-side logic (no GPU): the packing planner, attention work list, weight layout permutation, config loader,
module / state-dict layout and clip sharding."""
import math
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, build_model
from oracle import titok_oracle as O
from titok_video_b200 import plan as P
from titok_video_b200.config import AttrDict, load_config, tiny_config


def test_plan_layout_matches_reference_packing():
    shapes, tcs = [(8, 32, 48), (4, 16, 24), (16, 168, 168)], [3, 0, 128]
    pl = P.make_plan(shapes, tcs, (4, 8, 8))
    gs = [2 * 4 * 6, 1 * 2 * 3, 4 * 21 * 21]
    assert (pl.G, pl.T, pl.M) == (sum(gs), 131, sum(gs) + 131)
    # cu_seqlens = cumsum(grid + tokens)  (blocks.py:81-83)
    assert pl.cu_seqlens.tolist() == [0, gs[0] + 3, gs[0] + 3 + gs[1], pl.M]
    # the reference's mask: per clip token_count ones then grid_size zeros (blocks.py:85-86)
    mask = np.concatenate([np.r_[np.ones(t, bool), np.zeros(g, bool)] for t, g in zip(tcs, gs)])
    assert np.array_equal(pl.enc_src_row < 0, mask) and np.array_equal(pl.dec_src_row >= 0, mask)
    assert np.array_equal(np.nonzero(mask)[0], pl.latent_row) and np.array_equal(np.nonzero(~mask)[0], pl.patch_row)
    assert np.array_equal(pl.enc_src_row[~mask], np.arange(pl.G)) and np.array_equal(pl.dec_src_row[mask], np.arange(pl.T))
    assert pl.clip_offset == (0, 3 * 8 * 32 * 48, 3 * 8 * 32 * 48 + 3 * 4 * 16 * 24)


def test_plan_geometry_reproduces_patchify():
    """Gathering 16-byte runs at geom offsets == the oracle's einops-equivalent patchify, up to the feature
    permutation (c p0 p1 p2) <- (p0 p1 p2 c)."""
    from titok_video_b200.engine import patch_feature_perm

    shapes, tcs, patch = [(8, 16, 24), (4, 32, 8)], [2, 5], (4, 8, 8)
    clips = O.make_clips(shapes, 0)
    pl = P.make_plan(shapes, tcs, patch)
    flat = torch.cat([c.reshape(-1) for c in clips]).float()
    rows = torch.empty(pl.G, 768)
    for g in range(pl.G):
        off, W, HW, THW = pl.geom[g].tolist()
        for c in range(3):
            for p0 in range(4):
                for p1 in range(8):
                    src = off + c * THW + p0 * HW + p1 * W
                    rows[g, c * 256 + p0 * 64 + p1 * 8:c * 256 + p0 * 64 + p1 * 8 + 8] = flat[src:src + 8]
    ref = torch.cat([O.patchify(c.float(), patch) for c in clips])
    perm = patch_feature_perm(patch, 3)
    assert torch.equal(rows, ref[:, perm])
    assert sorted(perm.tolist()) == list(range(768))


def test_rope_ids_and_table():
    ids = P.rope_ids((2, 3, 4), 3)
    assert ids.shape == (3 + 24, 3)
    assert ids[:3].tolist() == [[0, 0, 0], [1, 1, 1], [2, 2, 2]]  # latent j -> (j,j,j)
    assert ids[3].tolist() == [3, 3, 3] and ids[4].tolist() == [3, 3, 4] and ids[3 + 4].tolist() == [3, 4, 3]
    assert ids[-1].tolist() == [1 + 3, 2 + 3, 3 + 3]
    inv = P.rope_inv_freqs()
    assert inv.shape == (10,) and abs(inv[0] - math.pi / 2) < 1e-15 and abs(inv[-1] - 10000 * math.pi / 2) < 1e-9
    tab = P.rope_table(ids, inv)
    assert tab.shape == (27, 60) and tab.dtype == np.float32
    cos, sin = O.rope_cos_sin((2, 3, 4), 3)
    assert np.allclose(tab.reshape(27, 30, 2)[..., 0], cos.numpy(), atol=1e-6)
    assert np.allclose(tab.reshape(27, 30, 2)[..., 1], sin.numpy(), atol=1e-6)
    # lane index = freq*3 + axis: lane 1 uses axis 1 with the lowest frequency
    assert abs(tab[4, 2 * 2] - math.cos(inv[0] * 4)) < 1e-6  # row 4 = patch (0,0,1)+3 -> axis2 id 4, lane 2


@pytest.mark.parametrize("hq,hkv", [(4, 2), (8, 2), (12, 4), (16, 4)])
def test_attention_work_list_covers_every_query_tile_once(hq, hkv):
    seq = [1892, 576, 130, 128, 1]
    starts = np.concatenate([[0], np.cumsum(seq)[:-1]]).tolist()
    w = P.attn_work_list(starts, seq, hq, hkv)
    assert w.dtype == np.int32 and w.shape[1] == 12
    seen = {}
    ratio = hq // hkv
    for r in w.tolist():
        q_row0, q_valid, q_head, kv_head, kv_row0, kv_len = r[0:2], r[2:4], r[4:6], r[6], r[7], r[8]
        assert kv_row0 in starts and kv_len == seq[starts.index(kv_row0)]
        for t in range(2):
            if q_valid[t] == 0:
                continue
            assert q_head[t] // ratio == kv_head  # GQA mapping: q head h reads kv head h // ratio
            assert kv_row0 <= q_row0[t] < kv_row0 + kv_len and (q_row0[t] - kv_row0) % 128 == 0
            assert q_valid[t] == min(128, kv_row0 + kv_len - q_row0[t])
            key = (q_row0[t], q_head[t])
            assert key not in seen
            seen[key] = 1
    want = sum(((s + 127) // 128) * hq for s in seq)
    assert len(seen) == want
    assert w[:, 8].tolist() == sorted(w[:, 8].tolist(), reverse=True)  # longest sequences first
    # leader records: every record names the FIRST record of its (clip, kv head) -- that record is its own leader and is
    # where the kernel library keeps the pair's score bound (kmax2 / kmax2b: scratch, zero in the plan)
    rows = w.tolist()
    first = {}
    for i, r in enumerate(rows):
        first.setdefault((r[7], r[6]), i)
    for i, r in enumerate(rows):
        assert r[10] == first[(r[7], r[6])] and rows[r[10]][10] == r[10]
        assert r[9] == 0 and r[11] == 0
    assert len(first) == len(seq) * hkv


def test_plan_rejects_bad_shapes():
    with pytest.raises(ValueError):
        P.make_plan([(8, 30, 32)], [4], (4, 8, 8))
    with pytest.raises(ValueError):
        P.make_plan([(8, 32, 32)], [4, 5], (4, 8, 8))
    with pytest.raises(ValueError):
        P.make_plan([(32, 32)], [4], (8, 8))
    pl = P.make_plan([(4, 8, 8)], [0], (4, 8, 8))  # zero latent tokens is a legal (if useless) request
    assert (pl.M, pl.T, pl.G) == (1, 0, 1)


def test_config_schema_accepts_reference_yaml():
    cfg = load_config(os.path.join(ROOT, "configs", "tiny.yaml"))
    m = cfg.tokenizer.model
    assert m.patch_size == [4, 8, 8] and m.fsq_levels == [7, 5, 5, 5, 5] and m.encoder_size == "tiny"
    assert cfg.training.sampling.train_seq_len == 6144
    assert isinstance(cfg, AttrDict) and tiny_config().tokenizer.model.decoder_size == "tiny"
    ref_yaml = "/root/reference/configs/tiny.yaml"
    if os.path.exists(ref_yaml):  # the reference's own file, verbatim
        c2 = load_config(ref_yaml)
        assert c2.tokenizer.model.fsq_levels == m.fsq_levels and c2.tokenizer.model.patch_size == m.patch_size


def test_module_layout_and_model_dims():
    from titok_video_b200.model.base.utils import geglu_inner_dim, get_model_dims

    assert get_model_dims("tiny") == (256, 4, [4, 2], 4.0)
    assert get_model_dims("small") == (512, 8, [8, 2], 4.0)
    assert get_model_dims("base") == (768, 12, [12, 4], 4.0)
    assert get_model_dims("large") == (1024, 24, [16, 4], 4.0)
    assert [geglu_inner_dim(w) for w in (256, 512, 768, 1024)] == [704, 1376, 2048, 2752]
    m = build_model(False)
    sd = m.state_dict()
    assert len(sd) == 76 and sum(v.numel() for v in sd.values()) == 6828295
    assert sd["encoder.proj_in.weight"].shape == (256, 768) and sd["encoder.proj_out.weight"].shape == (5, 256)
    assert sd["decoder.proj_in.weight"].shape == (256, 5) and sd["decoder.proj_out.weight"].shape == (768, 256)
    assert sd["encoder.model_layers.attn_layer.3.to_qkv.weight"].shape == (768, 256)
    assert sd["encoder.model_layers.ffd_layer.0.w12.weight"].shape == (1408, 256)
    assert sd["encoder.model_layers.ffd_layer.0.w3.weight"].shape == (256, 704)
    assert "encoder.model_layers.attn_post_ln.2.weight" in sd and "encoder.model_layers.attn_post_ln.3.weight" not in sd
    assert sd["encoder.mask_token"].shape == (1, 1)
    assert not any(k.startswith("quantize") for k in sd)  # FSQ buffers are non-persistent (fsq.py:64,67,76)
    assert m.quantize.codebook_size == 4375 and m.quantize._basis.tolist() == [1, 7, 35, 175, 875]
    assert m.encoder.model_layers.alpha == 8


def test_shard_clips_balances_cost():
    costs = [P.clip_cost(s, t, (4, 8, 8), 256, 4) for s, t in
             [((16, 168, 168), 128), ((8, 128, 128), 64), ((16, 128, 168), 32), ((8, 168, 168), 1)] * 4]
    parts = P.shard_clips(costs, 4)
    assert sorted(i for p in parts for i in p) == list(range(16))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) / min(loads) < 1.15
    assert P.shard_clips([1.0], 2) == [[0], []]


def test_fsq_division_free_arithmetic_is_exact():
    """csrc/fsq.cuh replaces q / half_width by q*RN(1/hw) plus one FMA correction, and rintf by the 1.5*2^23 magic
    constant; both must reproduce fp32 division / round-half-even bit for bit on every value FSQ can produce."""
    import numpy as np

    f = np.float32
    for hw in range(1, 17):
        r = f(1) / f(hw)
        for q in range(-hw - 2, hw + 3):
            r0 = f(f(q) * r)
            e = f(np.float64(f(q)) - np.float64(f(hw)) * np.float64(r0))      # fma(-hw, r0, q)
            r1 = f(np.float64(r0) + np.float64(e) * np.float64(r))            # fma(e, rcp, r0)
            assert r1 == f(q) / f(hw), (q, hw)
    x = np.concatenate([(np.random.default_rng(0).standard_normal(200000) * 4).astype(f), np.arange(-9, 9.5, 0.5, dtype=f)])
    m = f(12582912.0)
    assert np.array_equal((x + m) - m, np.rint(x))


def test_rope_table_gather_is_bit_identical_to_direct_evaluation():
    from titok_video_b200.plan import rope_ids, rope_inv_freqs, rope_table, rope_table_from_int_ids

    inv = rope_inv_freqs()
    for grid, tc in (((4, 21, 21), 128), ((2, 16, 16), 1), ((8, 32, 32), 256), ((3, 17, 20), 0)):
        ids = rope_ids(grid, tc)
        assert np.array_equal(rope_table(ids, inv), rope_table_from_int_ids(ids, inv))


@pytest.mark.parametrize("hq,hkv", [(4, 2), (12, 4)])
def test_attention_backward_work_lists_cover_every_tile_once(hq, hkv):
    """dkv: every (128-key tile, kv head) once, streaming all grouped query heads; dq: every (query tile, query head)."""
    from titok_video_b200.plan import attn_bwd_work_lists

    seq_lens = [513, 128, 1892, 77]
    starts = np.concatenate([[0], np.cumsum(seq_lens)[:-1]]).tolist()
    dkv, dq = attn_bwd_work_lists(starts, seq_lens, hq, hkv)
    assert dkv.dtype == np.int32 and dkv.shape[1] == 8 and dq.shape[1] == 8
    n_tiles = sum((s + 127) // 128 for s in seq_lens)
    assert len(dkv) == n_tiles * hkv and len(dq) == n_tiles * hq
    grp = hq // hkv
    seen = set()
    for st_row0, st_valid, st_head, o_head0, n_heads, clip_row0, clip_len, _ in dkv.tolist():
        assert (st_row0, st_head) not in seen
        seen.add((st_row0, st_head))
        assert o_head0 == st_head * grp and n_heads == grp
        assert clip_row0 in starts and clip_len == seq_lens[starts.index(clip_row0)]
        assert 0 < st_valid <= 128 and st_row0 + st_valid <= clip_row0 + clip_len and (st_row0 - clip_row0) % 128 == 0
    rows = np.zeros((sum(seq_lens), hq), dtype=np.int32)
    for st_row0, st_valid, st_head, o_head0, n_heads, clip_row0, clip_len, _ in dq.tolist():
        rows[st_row0:st_row0 + st_valid, st_head] += 1
        assert o_head0 == st_head // grp and n_heads == 1
    assert (rows == 1).all()
    # longest work first
    cost = dkv[:, 4].astype(np.int64) * ((dkv[:, 6] + 127) // 128)
    assert (np.diff(cost) <= 0).all()


def _reference_dynamic_batching(data, patch_size, token_range, max_grid, max_seq_len, eval, max_samples, rnd):
    """Restatement of dataset/video_dataset.py:130-172 for the test (the reference module needs webdataset / decord, which
    this image does not have): same control flow, written independently from titok_video_b200.data.batching."""
    assert math.prod(x // y for x, y in zip(max_grid, patch_size)) + token_range[1] <= max_seq_len
    batches, chunks, tcs, cur, seen = [], [], [], 0, 0
    for sample in data:
        g = math.prod(x // y for x, y in zip(sample["video"].shape[1:], patch_size))
        t = rnd.randrange(token_range[0], token_range[1] + 1)
        if eval:
            if seen > max_samples:
                break
            seen += 1
        if cur + g + t > max_seq_len:
            batches.append(([c["key"] for c in chunks], list(tcs)))
            chunks, tcs, cur = [], [], 0
        cur += g + t
        chunks.append(sample)
        tcs.append(t)
    return batches


@pytest.mark.parametrize("eval_mode", [False, True])
def test_dynamic_batches_follow_the_reference_semantics(eval_mode):
    import random

    from titok_video_b200.data import canonical_order, dynamic_batches

    rnd = random.Random(3)
    samples = []
    for i in range(60):
        shp = (3, rnd.choice([8, 12, 16]), rnd.choice([128, 136, 152, 168]), rnd.choice([128, 144, 168]))
        samples.append({"video": torch.empty(shp, dtype=torch.bfloat16, device="meta"), "key": i})
    kw = dict(patch_size=[4, 8, 8], token_range=[1, 128], max_grid=[16, 168, 168], max_seq_len=6144)
    want = _reference_dynamic_batching(samples, eval=eval_mode, max_samples=20, rnd=random.Random(11), **kw)
    got = list(dynamic_batches(samples, eval=eval_mode, max_samples=20, randrange=random.Random(11).randrange, **kw))
    assert [(b["key"], b["token_counts"].tolist()) for b in got] == want
    assert len(got) > 3
    for b in got:
        assert b["token_counts"].dtype == torch.int32
        s = sum(math.prod(x // y for x, y in zip(v.shape[1:], [4, 8, 8])) + int(t) for v, t in zip(b["video"], b["token_counts"]))
        assert s <= 6144
    with pytest.raises(AssertionError):
        list(dynamic_batches(samples, [4, 8, 8], [1, 128], [16, 168, 168], 1800))
    # canonical order: a permutation with its inverse; batches with the same multiset of (shape, tokens) get the same key
    shapes, tcs = [(8, 128, 128), (16, 168, 168), (8, 128, 128)], [5, 7, 3]
    perm, inv = canonical_order(shapes, tcs)
    assert sorted(perm) == [0, 1, 2] and [perm[inv[i]] for i in range(3)] == [0, 1, 2]
    key = lambda sh, tc, p: tuple((sh[i], tc[i]) for i in p)
    perm2, _ = canonical_order(shapes[::-1], tcs[::-1])
    assert key(shapes, tcs, perm) == key(shapes[::-1], tcs[::-1], perm2)


def test_token_container_roundtrip_and_validation(tmp_path):
    import io

    from titok_video_b200.data import read_tokens, write_tokens

    g = torch.Generator().manual_seed(0)
    idx = [torch.randint(0, 4375, (n,), generator=g, dtype=torch.int32) for n in (128, 1, 0, 77)]
    grids = [(16, 168, 168), (8, 128, 128), (4, 8, 8), (12, 136, 152)]
    p = str(tmp_path / "clips.ttkv")
    n = write_tokens(p, idx, grids, 4375)
    assert n == 6 + 10 + 8 * 4 + 2 * (128 + 1 + 0 + 77)
    got, gg, K = read_tokens(p)
    assert K == 4375 and gg == grids
    for a, b in zip(got, idx):
        assert a.dtype == torch.int32 and torch.equal(a, b)
    # wide codebooks use 32-bit indices; in-memory streams work too
    buf = io.BytesIO()
    wide = [torch.tensor([0, 70000, 99999])]
    write_tokens(buf, wide, [(4, 8, 8)], 100000)
    got, _, K = read_tokens(buf.getvalue())
    assert K == 100000 and got[0].tolist() == [0, 70000, 99999]
    with pytest.raises(ValueError):
        write_tokens(io.BytesIO(), [torch.tensor([4375])], [(4, 8, 8)], 4375)
    with pytest.raises(ValueError):
        read_tokens(open(p, "rb").read()[:-1])
    with pytest.raises(ValueError):
        read_tokens(b"nope" + bytes(40))


def test_cached_parameter_list_matches_named_parameters_and_notices_replacement():
    from titok_video_b200 import engine

    m = build_model(False)
    for stack in (m.encoder, m.decoder):
        a, b = engine.cached_named_params(stack), list(stack.named_parameters())
        assert [n for n, _ in a] == [n for n, _ in b] and all(x is y for (_, x), (_, y) in zip(a, b))
    old = m.encoder.proj_in.weight
    m.encoder.proj_in.weight = torch.nn.Parameter(torch.zeros_like(old))
    c = dict(engine.cached_named_params(m.encoder))
    assert c["proj_in.weight"] is m.encoder.proj_in.weight and c["proj_in.weight"] is not old


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) needs no GPU: one JSON line with the keys of
    the bench contract, `impl: reference`, a cpu_baseline block and an e2e block that repeats the line's value."""
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "clips/s" and line["value"] > 0
    # the unmodified reference files when a copy is at hand (/root/reference or the vendored baseline/_ref), else the port
    from oracle import ref_shim

    assert line["cpu_baseline"]["kind"] == ("reference" if ref_shim.available() else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("configs/tiny.yaml batch tokenise+reconstruct (C3)")
    # the arm is independent of the product: it must not have loaded the CUDA library
    probe = subprocess.run([sys.executable, "-c",
                            "import sys, runpy; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','1'];"
                            "runpy.run_path(%r, run_name='__main__');"
                            "print('LOADED' if any('titok_video_b200' in m for m in sys.modules) else 'CLEAN')"
                            % os.path.join(ROOT, "bench.py")], capture_output=True, text=True, timeout=600)
    assert probe.stdout.strip().endswith("CLEAN"), probe.stdout[-500:] + probe.stderr[-500:]

    # the port fallback (no reference copy anywhere) still prints the contract line
    env = dict(os.environ, TITOK_REFERENCE_ROOT="/nonexistent")
    code = ("import sys; sys.path.insert(0, %r); import bench; bench.reference_available = lambda: False;"
            "r = bench.cpu_reference_run(1, 1); print(r['kind'], r['clips_per_s'] > 0)" % ROOT)
    fb = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert fb.stdout.strip() == "port True", fb.stdout[-300:] + fb.stderr[-800:]


def test_latent_work_list_is_the_leading_tiles_of_the_full_list():
    """The encoder's last layer needs attention output for latent rows only (blocks.py:101); they lead every clip. Its
    work list must consist of records of the full list restricted to the tiles that hold latent rows, cover every latent
    row, hold nothing for clips without tokens, and carry valid leader indices."""
    from titok_video_b200.plan import ATTN_TILE, get_attn_work, get_attn_work_latent, make_plan

    rng = np.random.default_rng(3)
    for _ in range(6):
        B = int(rng.integers(1, 7))
        shapes = [(4 * int(rng.integers(1, 5)), 8 * int(rng.integers(1, 12)), 8 * int(rng.integers(1, 12))) for _ in range(B)]
        tcs = [int(rng.choice([0, 1, 37, 128, 129, 256, 300])) for _ in range(B)]
        pl = make_plan(shapes, tcs, (4, 8, 8), arrays=True)
        starts = pl.cu_seqlens[:-1].tolist()
        for hq, hkv in [(4, 2), (8, 2), (12, 4), (16, 16)]:
            full, lat = get_attn_work(pl, hq, hkv), get_attn_work_latent(pl, hq, hkv)
            # (even head groups take a row filter of the full list: it must equal the generic construction)
            from titok_video_b200.plan import attn_work_list
            assert np.array_equal(lat, attn_work_list(starts, pl.seq_lens, hq, hkv, q_lens=pl.token_counts))
            # (head, first row, valid rows, kv head, clip start, clip length) of every used query tile
            tiles = lambda w: {(int(r[4 + j]), int(r[j]), int(r[2 + j]), int(r[6]), int(r[7]), int(r[8]))
                               for r in w for j in (0, 1) if r[2 + j] > 0}
            tf, tl = tiles(full), tiles(lat)
            assert tl <= tf
            want = {t for t in tf if t[1] - t[4] < tcs[starts.index(t[4])]}
            assert tl == want
            covered = np.zeros(pl.M, dtype=np.int64)
            for (h, r0, v, _, _, _) in tl:
                covered[r0:r0 + v] += 1
            assert (covered[pl.latent_row] == hq).all()
            assert lat.shape[0] <= (pl.T // ATTN_TILE + B + 1) * (hq // 2 if (hq // hkv) % 2 == 0 else hq)  # BucketPlan.W_lat_max
            for i, r in enumerate(lat):
                lead = lat[r[10]]
                assert r[10] <= i and lead[6] == r[6] and lead[7] == r[7] and lat[r[10]][10] == r[10]


def test_latent_backward_work_lists():
    """Attention backward of the encoder's last layer in training: only the latent rows carry an output gradient. dq keeps
    the full list's records of the query tiles that hold latent rows; dkv keeps every key tile of the clips that have
    tokens and streams just their latent query rows (`clip_len` = token count); nothing for clips without tokens."""
    from titok_video_b200.plan import ATTN_TILE, get_attn_bwd_work, get_attn_bwd_work_latent, make_plan

    rng = np.random.default_rng(5)
    for _ in range(6):
        B = int(rng.integers(1, 7))
        shapes = [(4 * int(rng.integers(1, 5)), 8 * int(rng.integers(1, 12)), 8 * int(rng.integers(1, 12))) for _ in range(B)]
        tcs = [int(rng.choice([0, 1, 37, 128, 129, 256, 300])) for _ in range(B)]
        pl = make_plan(shapes, tcs, (4, 8, 8), arrays=True)
        starts = pl.cu_seqlens[:-1].tolist()
        tok = {st: t for st, t in zip(starts, tcs)}
        for hq, hkv in [(4, 2), (12, 4), (16, 16)]:
            (a, b), (al, bl) = get_attn_bwd_work(pl, hq, hkv), get_attn_bwd_work_latent(pl, hq, hkv)
            rows = lambda w: {tuple(int(v) for v in r) for r in w}
            assert rows(bl) == {r for r in rows(b) if r[0] - r[5] < tok[r[5]]}
            want = {r[:6] + (tok[r[5]], 0) for r in rows(a) if tok[r[5]] > 0}
            assert rows(al) == want
            assert all(r[6] <= pl.seq_lens[starts.index(r[5])] for r in rows(al))


def test_training_latent_tail_is_chosen_by_batch_size(monkeypatch):
    """backward.latent_tail_active: "auto" takes the latent-tail training path for packed batches of at least
    TRAIN_TAIL_MIN_ROWS rows (GPU-bound steps: its last layer is enqueued from Python), True / False force it."""
    from titok_video_b200 import backward

    monkeypatch.setattr(backward, "TRAIN_LATENT_TAIL", "auto")
    monkeypatch.setattr(backward, "TRAIN_TAIL_MIN_ROWS", 24000)
    assert backward.latent_tail_active(16 * 1892) and not backward.latent_tail_active(3 * 1892)
    monkeypatch.setattr(backward, "TRAIN_LATENT_TAIL", True)
    assert backward.latent_tail_active(1)
    monkeypatch.setattr(backward, "TRAIN_LATENT_TAIL", False)
    assert not backward.latent_tail_active(10 ** 6)
