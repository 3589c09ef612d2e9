"""torch.library registration of the stack operators (SURVEY 8b "Registration"): schema / fake / autograd consistency
checked by torch.library.opcheck, FakeTensor shape propagation without a kernel launch, and autograd through the
registered op equal to the oracle's gradients (the module path itself goes through these ops in training mode, so
tests/test_gpu_backward.py covers their numerics at length)."""
import pytest
import torch

from conftest import build_model
from oracle import titok_oracle as O

pytestmark = pytest.mark.gpu
SHAPES, TCS = [(8, 32, 32), (4, 16, 24)], [8, 3]


def _setup(train):
    from titok_video_b200 import engine, ops

    model = build_model(True).cuda()
    model.train(train)
    clips = [c.cuda() for c in O.make_clips(SHAPES, 0)]
    dp = model.encoder._plan([tuple(c.shape[1:]) for c in clips], TCS, clips[0].device)
    flat = engine.flatten_clips(clips, dp)
    consts = model.quantize._consts(clips[0].device)
    return model, clips, dp, flat, consts, ops


def test_ops_are_registered_with_fake_and_autograd():
    model, clips, dp, flat, consts, ops = _setup(True)
    assert hasattr(torch.ops.titok_b200, "encoder_stack") and hasattr(torch.ops.titok_b200, "decoder_stack")
    h = ops.handle_for(model.encoder, dp, consts)
    assert ops.handle_for(model.encoder, dp, consts) == h  # stable per (module, plan)
    # FakeTensor propagation: shapes / dtypes from the plan, no launch
    from torch._subclasses.fake_tensor import FakeTensorMode
    from titok_video_b200 import _lib

    n0 = _lib.LAUNCHES
    with FakeTensorMode(allow_non_fake_inputs=True) as mode:
        ff = mode.from_tensor(flat)
        z, codes, idx = torch.ops.titok_b200.encoder_stack(ff, [], h, False)
        assert z.shape == (sum(TCS), 5) and z.dtype == torch.bfloat16 and idx.dtype == torch.int32 and idx.shape == (sum(TCS),)
        hd = ops.handle_for(model.decoder, dp)
        out = torch.ops.titok_b200.decoder_stack(mode.from_tensor(torch.empty((sum(TCS), 5), device="cuda")), [], hd, False)
        assert out.shape == (dp.plan.total_numel,) and out.dtype == torch.bfloat16
    assert _lib.LAUNCHES == n0
    # inference through the op == the module path
    with torch.no_grad():
        z, codes, idx = torch.ops.titok_b200.encoder_stack(flat, [], h, False)
        z_m = model.encoder(clips, TCS)
        rec = torch.ops.titok_b200.decoder_stack(codes, [], hd, False)
        rec_m = model.decode(codes, TCS, SHAPES)
    assert torch.equal(z, z_m)
    assert torch.equal(rec, torch.cat([r.reshape(-1) for r in rec_m]))


def test_opcheck_schema_fake_and_autograd_registration():
    model, clips, dp, flat, consts, ops = _setup(True)
    from titok_video_b200 import backward

    h = ops.handle_for(model.encoder, dp, consts)
    params = backward.stack_params(model.encoder)
    for tu in ("test_schema", "test_faketensor", "test_autograd_registration"):
        torch.library.opcheck(torch.ops.titok_b200.encoder_stack.default, (flat, params, h, True), test_utils=tu)
    hd = ops.handle_for(model.decoder, dp)
    codes = torch.randn((sum(TCS), 5), device="cuda").to(torch.bfloat16).requires_grad_(True)
    for tu in ("test_schema", "test_faketensor", "test_autograd_registration"):
        torch.library.opcheck(torch.ops.titok_b200.decoder_stack.default, (codes, backward.stack_params(model.decoder), hd, True),
                              test_utils=tu)


def test_autograd_through_the_registered_ops_matches_the_oracle():
    model, clips, dp, flat, consts, ops = _setup(True)
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):  # Lightning's bf16-mixed: fp32 parameters are graph edges only
        recon, d = model(clips, TCS)
    loss = torch.stack([(r_.float() - c.float()).abs().mean() for c, r_ in zip(clips, recon)]).mean()
    assert "encoder_stack" in str(type(recon[0].grad_fn)) or recon[0].grad_fn is not None
    loss.backward()
    loss_o, grads_o = O.titok_train_grads(sd, [7, 5, 5, 5, 5], [4, 8, 8], [c.cpu() for c in clips], TCS)
    assert abs(float(loss) - loss_o) < 5e-3 * loss_o
    for k, p in model.named_parameters():
        assert p.grad is not None and p.grad.dtype == torch.float32, k
        g, go = p.grad.float().cpu().reshape(-1).double(), grads_o[k].reshape(-1).double()
        if g.numel() > 1 and float(go.norm()) > 0:
            assert float((g @ go) / (g.norm() * go.norm() + 1e-300)) > 0.7, k  # stress init: the bar of test_gpu_backward
    # stale handle -> loud error
    with pytest.raises(RuntimeError):
        torch.ops.titok_b200.decoder_stack(torch.zeros((1, 5), device="cuda"), [], 10 ** 9, False)
