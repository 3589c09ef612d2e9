"""Development aid: times ttk_attn_varlen_fwd alone at the bench shape (B clips A, 4 q heads / 2 kv heads, d = 64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from titok_video_b200 import _lib
from titok_video_b200.engine import _ptr, _stream
from titok_video_b200.plan import attn_work_list

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
s, w, gqa = 1892, 256, 128
dev = torch.device("cuda:0")
M = B * s
qkv = (torch.randn(M, 2 * w + 2 * gqa, device=dev) * 1.0).to(torch.bfloat16)
out = torch.empty(M, w, device=dev, dtype=torch.bfloat16)
work = torch.from_numpy(np.ascontiguousarray(attn_work_list([i * s for i in range(B)], [s] * B, 4, 2))).to(dev)
st = _stream()
run = lambda: _lib.call("ttk_attn_varlen_fwd", _ptr(qkv), qkv.stride(0), M, w, gqa, _ptr(work), work.shape[0], 0.125, _ptr(out), w, st)
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
    tr = torch.zeros(2048, 64, dtype=torch.int64, device=dev)
    _lib.check(_lib.fn("ttk_debug_set_trace")(_ptr(tr)))
    run()
    torch.cuda.synchronize()
    _lib.check(_lib.fn("ttk_debug_set_trace")(_vp(0)))
    t = tr.cpu()
    names = ["S ready", "S in regs", "max+xchg", "exps done", "pv_done ok", "P stored", "t1 S ready", "t1 P stored", "MMA S0 issue", "MMA S1 issue", "MMA PV0 issue", "MMA PV1 issue"]
    # whole-CTA phases and back-to-back CTAs on one SM (globaltimer-free: clock64 is per SM, so compare CTAs on the same SM)
    by_sm = {}
    for cta in range(work.shape[0]):
        r = t[cta]
        if int(r[63]):
            by_sm.setdefault(int(r[59]), []).append((int(r[60]), int(r[58]), int(r[61]), int(r[62]), int(r[63]), cta))
    for sm in (0, 77):
        seq = sorted(by_sm.get(sm, []))
        print(f"SM {sm}: per CTA [setup->firstS | kv loop | epilogue | teardown] and gap to the next CTA's setup stamp")
        for a, b in zip(seq[:6], seq[1:7]):
            print(f"   cta {a[5]}: {a[1]-a[0]} | {a[2]-a[1]} | {a[3]-a[2]} | {a[4]-a[3]} | gap {b[0]-a[4]}")
    for cta in (0, 500):
        r = t[cta]
        t0 = int(r[0])
        print(f"cta {cta}: cycles relative to 'S ready' of kv iteration 3 (tile 0)")
        for jj in range(4):
            ev = [int(r[jj * 12 + k]) - t0 for k in range(12)]
            print(f"  j={jj + 3}: " + "  ".join(f"{n}={v}" for n, v in zip(names, ev)))
