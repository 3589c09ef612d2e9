"""TiTok model facade with the API of the reference's model/titok.py:
`TiTok(config)`, `.encoder/.quantize/.decoder`, `encode`, `decode`, `decode_indices`, `forward`.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from .. import engine
from .base.blocks import TiTokDecoder, TiTokEncoder
from .base.utils import init_weights
from .quantizer.fsq import FSQ


def _float_dtype(t: torch.Tensor) -> torch.dtype:
    """dtype of results derived from a clip: uint8 clips (decoded frames) produce bf16 results."""
    return t.dtype if t.is_floating_point() else torch.bfloat16


class TiTok(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        conf = config.tokenizer.model
        token_size = len(conf.fsq_levels)
        self.encoder = TiTokEncoder(model_size=conf.encoder_size, patch_size=conf.patch_size, in_channels=3,
                                    out_channels=token_size)
        self.quantize = FSQ(levels=list(conf.fsq_levels))
        self.decoder = TiTokDecoder(model_size=conf.decoder_size, patch_size=conf.patch_size, in_channels=token_size,
                                    out_channels=3)
        self.apply(init_weights)

    # ---- titok.py:47-52 -------------------------------------------------------------------------
    @torch.compiler.disable()  # train.py:38-39 may wrap the model in torch.compile: the kernels are opaque to Dynamo
    def encode(self, x: Sequence[torch.Tensor], token_counts, grids=None, split_indices: bool = False):
        """x_q [sum(token_counts), token_size] in the clips' dtype, {'indices': int32 [sum(token_counts)]}.
        split_indices=True returns a tuple of per-clip index tensors (the reference's intent; its own
        torch.split call fails for tensor token_counts)."""
        z, codes, idx, _ = self.encoder.forward_impl(x, token_counts, grids, fsq=self.quantize)
        if z.requires_grad:  # training: FSQ with its straight-through gradient (fsq.py:48-51)
            x_q, d = self.quantize(z)
            x_q, indices = x_q.to(_float_dtype(x[0])), d["indices"]
        else:
            x_q = codes.clone().to(_float_dtype(x[0]))
            indices = idx.clone()
        if split_indices:
            indices = torch.split(indices, engine.to_host_ints(token_counts), dim=0)
        return x_q, {"indices": indices}

    # ---- titok.py:54-62 -------------------------------------------------------------------------
    @torch.compiler.disable()  # train.py:38-39 may wrap the model in torch.compile: the kernels are opaque to Dynamo
    def decode_indices(self, indices, grids, token_counts=None):
        if token_counts is None:
            assert type(indices) in [list, tuple]
            token_counts = [int(t.shape[0]) for t in indices]
            indices = torch.cat(list(indices), dim=0)
        x_q = self.quantize.indices_to_codes(indices, out_dtype=torch.bfloat16)
        return self.decoder(x_q, token_counts, grids)

    # ---- titok.py:64-66 -------------------------------------------------------------------------
    @torch.compiler.disable()  # train.py:38-39 may wrap the model in torch.compile: the kernels are opaque to Dynamo
    def decode(self, x, token_counts, grids):
        return self.decoder(x, token_counts, grids)

    # ---- titok.py:68-74 -------------------------------------------------------------------------
    @torch.compiler.disable()  # train.py:38-39 may wrap the model in torch.compile: the kernels are opaque to Dynamo
    def forward(self, x: Sequence[torch.Tensor], token_counts):
        """list of reconstructed clips [3,T,H,W] (input dtype) and {'indices': int32}."""
        grids = [tuple(v.shape[1:]) for v in x]
        tcs = engine.to_host_ints(token_counts)
        z, codes, idx, dp = self.encoder.forward_impl(x, tcs, grids, fsq=self.quantize)
        if z.requires_grad:  # training: encoder -> FSQ (straight-through) -> decoder, all recorded by autograd
            codes, d = self.quantize(z)
            out, _ = self.decoder.forward_impl(codes, tcs, grids)
            from .. import backward

            return backward.split_clips_autograd(out.to(_float_dtype(x[0])), dp.plan), {"indices": d["indices"]}
        out, _ = self.decoder.forward_impl(codes, tcs, grids)
        if out.requires_grad:  # frozen encoder, trainable decoder
            from .. import backward

            return backward.split_clips_autograd(out.to(_float_dtype(x[0])), dp.plan), {"indices": idx.clone()}
        recon = engine.split_clips(out.clone().to(_float_dtype(x[0])), dp.plan)
        return recon, {"indices": idx.clone()}

    # ---- throughput path: no defensive copies, results live in the plan's workspace ---------------
    @torch.no_grad()
    def tokenize_reconstruct_(self, x: Sequence[torch.Tensor], token_counts, use_graph: Optional[bool] = None,
                              with_error: bool = False):
        """forward() without the output clones: the returned clips / indices alias workspace buffers that the next
        call with the same shapes overwrites. Used by bench.py and batch jobs that consume results immediately.

        The 49 kernel launches of encoder + FSQ + decoder are captured once per (shapes, token_counts) signature into
        a CUDA graph and replayed (`use_graph=False` or TTK_CUDA_GRAPH=0 launches them one by one). Inputs are copied
        into the plan's static clip buffer first; weights are refreshed in place, so a captured graph stays valid
        across optimizer steps.

        with_error=True also returns 'clip_error': fp64 [B, 2] = per-clip (sum |x - recon|, sum (x - recon)^2), the
        numerators of the reference's L1 reconstruction loss (loss_module.py:118) and of PSNR, computed on the device
        by ttk_clip_error so that a tokenisation job only has to bring indices and two scalars per clip to the host."""
        dev = x[0].device
        engine.require_cuda(dev)
        grids = [tuple(v.shape[1:]) for v in x]
        tcs = engine.to_host_ints(token_counts)
        if len(tcs) != len(x):
            raise ValueError("len(token_counts) must equal the number of clips")
        enc, dec = self.encoder, self.decoder
        dp = enc._plan(grids, tcs, dev)
        # (the per-clip error compares against the normalised clips, so it needs them materialised)
        flat = engine.flatten_clips(x, dp, keep_u8=not with_error)
        consts = self.quantize._consts(dev)
        engine.prepared(enc, "enc")
        engine.prepared(dec, "dec")
        out = dp.buf("clips_out", (dp.plan.total_numel,))

        def launch():
            _, codes, idx = engine.encoder_launch(enc, dp, flat, consts)
            engine.decoder_launch(dec, dp, codes, out)
            err = engine.clip_error_launch(dp, flat, out) if with_error else None
            return idx, err

        if use_graph is None:
            use_graph = os.environ.get("TTK_CUDA_GRAPH", "1") != "0"
        if not use_graph:
            idx, err = launch()
        else:
            key = ("tokenize_reconstruct", id(self), flat.data_ptr(), bool(with_error))
            entry = dp.graphs.get(key)
            if entry is not None and entry[3] != engine.arena_generation():
                entry = None  # a workspace arena was re-allocated since the capture: the graph's pointers are stale
            if entry is None:
                launch()  # eager warm-up: workspace allocation, one-time function attributes
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    idx, err = launch()
                entry = (g, idx, err, engine.arena_generation())
                dp.graphs[key] = entry
            entry[0].replay()
            idx, err = entry[1], entry[2]
        res = {"indices": idx}
        if with_error:
            res["clip_error"] = err
        return engine.split_clips(out, dp.plan), res

    @torch.no_grad()
    def tokenize_reconstruct_bucketed_(self, x: Sequence[torch.Tensor], token_counts, with_error: bool = False):
        """`tokenize_reconstruct_` for RAGGED streams (the reference's dataloader draws a new batch composition every
        step, dataset/video_dataset.py:130-172): the launch sequence is captured ONCE per shape bucket -- (patches, latent
        tokens, clips) rounded up to engine.BucketPlan.{G,T,B}_STEP -- and replayed for every composition that falls into
        the bucket. Per step the host plans in O(B), uploads one small buffer and launches one graph; the per-row metadata
        is expanded on the device inside the graph. Results are bit-identical to `tokenize_reconstruct_` / `forward`
        (tests/test_gpu_model.py); like there, they alias workspace buffers until the next call."""
        from ..plan import make_plan

        dev = x[0].device
        engine.require_cuda(dev)
        enc, dec = self.encoder, self.decoder
        if tuple(enc.heads) != tuple(dec.heads) or enc.patch_size_tuple != dec.patch_size_tuple:
            return self.tokenize_reconstruct_(x, token_counts, with_error=with_error)  # (buckets assume one head layout)
        grids = [tuple(v.shape[1:]) for v in x]
        tcs = engine.to_host_ints(token_counts)
        if len(tcs) != len(x):
            raise ValueError("len(token_counts) must equal the number of clips")
        plan = make_plan(grids, tcs, enc.patch_size_tuple, enc.patch_channels, arrays=False)
        bp = engine.get_bucket_plan(plan, dev, enc.heads)
        bp.upload(plan)
        u8 = x[0].dtype == torch.uint8
        if u8 and with_error:
            raise ValueError("with_error needs float clips (the error is taken against the normalised input)")
        consts = self.quantize._consts(dev)
        engine.prepared(enc, "enc")
        engine.prepared(dec, "dec")

        def stage():
            """(re)fetch the static input / output buffers of the bucket and copy this step's clips in (one cat kernel)"""
            f = bp.buf("clips_in_u8" if u8 else "clips_in", (bp.plan.total_numel,), torch.uint8 if u8 else torch.bfloat16)
            if u8:
                torch.cat([v.reshape(-1) for v in x], out=f[:plan.total_numel])
            else:
                torch.cat([v.reshape(-1).to(torch.bfloat16) for v in x], out=f[:plan.total_numel])
            return f, bp.buf("clips_out", (bp.plan.total_numel,))

        flat, out = stage()

        def launch():
            bp.build_launch()
            _, codes, idx = engine.encoder_launch(enc, bp, flat, consts)
            engine.decoder_launch(dec, bp, codes, out)
            err = engine.clip_error_launch(bp, flat, out) if with_error else None
            return idx, err

        key = ("bucketed", id(self), u8, bool(with_error))
        entry = bp.graphs.get(key)
        if entry is not None and entry[3] != engine.arena_generation():
            entry = None  # a workspace arena was re-allocated since the capture: the graph's pointers are stale
        if entry is None:
            gen = engine.arena_generation()
            launch()  # eager warm-up: workspace allocation, one-time function attributes
            torch.cuda.current_stream().synchronize()
            if engine.arena_generation() != gen:
                flat, out = stage()  # the warm-up grew an arena: the buffers moved
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                idx, err = launch()
            entry = (g, idx, err, engine.arena_generation(), out)
            bp.graphs[key] = entry
        out = entry[4]
        entry[0].replay()
        idx, err = entry[1][:plan.T], entry[2]
        res = {"indices": idx}
        if with_error:
            res["clip_error"] = err[:len(tcs)]
        return engine.split_clips(out[:plan.total_numel], plan), res

