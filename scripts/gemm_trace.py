"""Development aid: per-CTA pipeline timelines of the GEMM kernels at the bench shapes (runs on the GPU box).
Prints, for a few CTAs, clock64 offsets (cycles from kernel start) of: producer first-stage issue, MMA start/first
operand/last k-block, epilogue wait begin/end and tile end."""
import ctypes, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from titok_video_b200 import _lib
from titok_video_b200.engine import _ptr, _stream, _vp

dev = torch.device("cuda:0")
M, w, inner, gqa = 30272, 256, 704, 128
bf = torch.bfloat16
A = torch.randn(M, w, device=dev).to(bf)
H = torch.randn(M, inner, device=dev).to(bf)
X = torch.randn(M, w, device=dev).to(bf)
XN = torch.empty_like(X)
Wqkv = (torch.randn(2 * w + 2 * gqa, w, device=dev) * 0.05).to(bf)
W12 = (torch.randn(2 * inner, w, device=dev) * 0.05).to(bf)
W3 = (torch.randn(w, inner, device=dev) * 0.05).to(bf)
Wo = (torch.randn(w, w, device=dev) * 0.05).to(bf)
Wp = (torch.randn(768, w, device=dev) * 0.05).to(bf)
bp = torch.zeros(768, device=dev).to(bf)
rope = torch.rand(M, 60, device=dev)
qkv = torch.empty(M, 2 * w + 2 * gqa, device=dev, dtype=bf)
hout = torch.empty(M, inner, device=dev, dtype=bf)
rows = torch.empty(M, 768, device=dev, dtype=bf)
wn = torch.ones(w, device=dev)
st = _stream()

def run(name):
    if name == "qkv":
        _lib.call("ttk_gemm_qkv_rope", _ptr(A), w, _ptr(Wqkv), w, M, w, w, gqa, _ptr(rope), _ptr(qkv), qkv.stride(0), _vp(0), st)
    elif name == "geglu":
        _lib.call("ttk_gemm_geglu", _ptr(A), w, _ptr(W12), w, M, inner, w, _ptr(hout), inner, st)
    elif name == "resid256":
        _lib.call("ttk_gemm_resid_norm256", _ptr(A), w, _ptr(Wo), w, M, w, _ptr(X), w, 1, 8.0, _ptr(wn), _ptr(wn), _ptr(X), _ptr(XN), w, st)
    elif name == "resid704":
        _lib.call("ttk_gemm_resid_norm256", _ptr(H), inner, _ptr(W3), inner, M, inner, _ptr(X), w, 1, 8.0, _ptr(wn), _ptr(wn), _ptr(X), _ptr(XN), w, st)
    elif name == "store768":
        _lib.call("ttk_gemm_bf16", _ptr(A), w, _ptr(Wp), w, M, 768, w, _ptr(bp), _ptr(rows), 768, _vp(0), 0, st)

names = ["store768", "qkv", "geglu", "resid256", "resid704"]
for name in names:
    for _ in range(3):
        run(name)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run(name)
    e1.record()
    torch.cuda.synchronize()
    print(f"== {name}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us/launch")
    tr = torch.zeros(148, 64, dtype=torch.int64, device=dev)
    _lib.check(_lib.fn("ttk_debug_set_trace")(_ptr(tr)))
    run(name)
    torch.cuda.synchronize()
    _lib.check(_lib.fn("ttk_debug_set_trace")(_vp(0)))
    t = tr.cpu()
    for cta in (0, 73, 147):
        r = t[cta]
        t0 = int(r[0])
        line = []
        for it in range(7):
            ev = [int(r[1 + 8 * it + k]) for k in range(7)]
            if ev[5] == 0:
                break
            line.append("t%d[P %d | M %d %d %d | E %d %d %d]" % ((it,) + tuple(e - t0 if e else -1 for e in ev)))
        print(f" cta {cta}: end {int(r[63]) - t0}  " + "  ".join(line))
        if int(r[48]):
            print("    resid epilogue tile0 phases (after accum+resid | after norm1 | x' store issued | x' read done | before final wait):",
                  [int(r[k]) - t0 for k in range(48, 53)])
