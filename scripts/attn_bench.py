"""Development aid: times ttk_attn_varlen_fwd alone at the bench shape (B clips A, 4 q heads / 2 kv heads, d = 64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from titok_video_b200 import _lib
from titok_video_b200.engine import _ptr, _stream
from titok_video_b200.plan import attn_work_list

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
s, w, gqa = int(os.environ.get("ATTN_S", "1892")), 256, 128
dev = torch.device("cuda:0")
M = B * s
qkv = (torch.randn(M, 2 * w + 2 * gqa, device=dev) * 1.0).to(torch.bfloat16)
out = torch.empty(M, w, device=dev, dtype=torch.bfloat16)
work = torch.from_numpy(np.ascontiguousarray(attn_work_list([i * s for i in range(B)], [s] * B, 4, 2))).to(dev)
st = _stream()
from titok_video_b200.engine import _vp
kn = (qkv[:, 2 * w:2 * w + gqa].float() ** 2).reshape(M, gqa // 64, 64).sum(-1).t().contiguous()  # what ttk_gemm_qkv_rope leaves behind
kn_ptr = _vp(0) if "--no-knorm" in sys.argv else _ptr(kn)
run = lambda: _lib.call("ttk_attn_varlen_fwd", _ptr(qkv), qkv.stride(0), M, w, gqa, _ptr(work), work.shape[0], 0.125, _ptr(out), w, kn_ptr, st)
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{os.environ.get('TTK_LIB_PATH', 'default')}: B={B} {ms * 1e3:.1f} us/launch  {B * 4.0 * s * s * w / ms / 1e9:.1f} TFLOP/s  checksum {out.float().abs().mean().item():.5f}")

if "--trace" in sys.argv:
    from titok_video_b200.engine import _vp
    n_cta = 2 * work.shape[0]  # one CTA per query tile (two per work record)
    tr = torch.zeros(n_cta, 64, dtype=torch.int64, device=dev)
    _lib.check(_lib.fn("ttk_debug_set_trace")(_ptr(tr)))
    run()
    torch.cuda.synchronize()
    _lib.check(_lib.fn("ttk_debug_set_trace")(_vp(0)))
    t = tr.cpu()
    names = ["S ready", "S in regs", "max+xchg", "exps done", "pv_done ok", "P stored", "-", "-", "MMA S(j+1) issue", "-", "MMA PV issue", "-"]
    for cta in (0, 1, 1001, 2500):
        r = t[cta]
        t0 = int(r[0])
        print(f"cta {cta}: cycles relative to 'S ready' of kv iteration 3")
        for jj in range(4):
            ev = [int(r[jj * 12 + k]) - t0 for k in range(12)]
            print(f"  j={jj + 3}: " + "  ".join(f"{n}={v}" for n, v in zip(names, ev) if n != "-"))
    # iteration period statistics over all CTAs: S ready(j=6) - S ready(j=3)
    d = (t[:, 36] - t[:, 0]).float() / 3
    ok = (t[:, 36] > 0) & (t[:, 0] > 0)
    print(f"mean kv-iteration period {d[ok].mean().item():.0f} cycles (median {d[ok].median().item():.0f}) over {int(ok.sum())} CTAs")
    for k in range(1, 6):
        seg = (t[:, k] - t[:, k - 1]).float()
        print(f"  {names[k - 1]} -> {names[k]}: mean {seg[ok].mean().item():.0f} median {seg[ok].median().item():.0f}")
    seg = (t[:, 12] - t[:, 5]).float()
    print(f"  P stored(j) -> S ready(j+1): mean {seg[ok].mean().item():.0f} median {seg[ok].median().item():.0f}")
    seg = (t[:, 8] - t[:, 1]).float()
    print(f"  S in regs(j) -> MMA S(j+1) issue: mean {seg[ok].mean().item():.0f} median {seg[ok].median().item():.0f}")
    seg = (t[:, 10] - t[:, 5]).float()
    print(f"  P stored(j) -> MMA PV(j) issue: mean {seg[ok].mean().item():.0f} median {seg[ok].median().item():.0f}")
    seg = (t[:, 12] - t[:, 8]).float()
    print(f"  MMA S(j+1) issue -> S ready(j+1) seen by softmax: mean {seg[ok].mean().item():.0f} median {seg[ok].median().item():.0f}")
    # CTA-level phases (slots 48..51: entry, loop entered, loop left, epilogue stores issued)
    okc = (t[:, 48] > 0) & (t[:, 51] > 0)
    for a, b, nm in ((48, 49, "entry -> loop entered (prologue)"), (49, 50, "loop"), (50, 51, "epilogue"), (48, 51, "CTA lifetime")):
        seg = (t[:, b] - t[:, a]).float()
        print(f"  {nm}: mean {seg[okc].mean().item():.0f} median {seg[okc].median().item():.0f} cycles")
