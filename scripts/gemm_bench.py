"""Development aid: times the GEMM-family entry points alone at the bench shape (B clips A, tiny width)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from titok_video_b200 import _lib
from titok_video_b200.engine import _ptr, _stream, _vp

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
M, w, inner, gqa = B * 1892, 256, 704, 128
bf = torch.bfloat16
A = torch.randn(M, w, device=dev).to(bf)
H = torch.randn(M, inner, device=dev).to(bf)
X = torch.randn(M, w, device=dev).to(bf)
XO, XN = torch.empty_like(X), torch.empty_like(X)
Wqkv = (torch.randn(2 * w + 2 * gqa, w, device=dev) * 0.05).to(bf)
W12 = (torch.randn(2 * inner, w, device=dev) * 0.05).to(bf)
W3 = (torch.randn(w, inner, device=dev) * 0.05).to(bf)
Wo = (torch.randn(w, w, device=dev) * 0.05).to(bf)
rope = torch.rand(M, 60, device=dev)
qkv = torch.empty(M, 2 * w + 2 * gqa, device=dev, dtype=bf)
hout = torch.empty(M, inner, device=dev, dtype=bf)
knorm = torch.empty(gqa // 64, M, device=dev, dtype=torch.float32)
wn = torch.ones(w, device=dev)
st = _stream()
runs = {
    "qkv": (lambda: _lib.call("ttk_gemm_qkv_rope", _ptr(A), w, _ptr(Wqkv), w, M, w, w, gqa, _ptr(rope), _ptr(qkv), qkv.stride(0), _ptr(knorm), st),
            2.0 * M * w * (2 * w + 2 * gqa), M * (w + 2 * w + 2 * gqa) * 2),
    "geglu": (lambda: _lib.call("ttk_gemm_geglu", _ptr(A), w, _ptr(W12), w, M, inner, w, _ptr(hout), inner, st),
              2.0 * M * w * 2 * inner, M * (w + inner) * 2),
    "resid256": (lambda: _lib.call("ttk_gemm_resid_norm256", _ptr(A), w, _ptr(Wo), w, M, w, _ptr(X), w, 1, 8.0, _ptr(wn), _ptr(wn), _ptr(XO), _ptr(XN), w, st),
                 2.0 * M * w * w, M * 4 * w * 2),
    "resid704": (lambda: _lib.call("ttk_gemm_resid_norm256", _ptr(H), inner, _ptr(W3), inner, M, inner, _ptr(X), w, 1, 8.0, _ptr(wn), _ptr(wn), _ptr(XO), _ptr(XN), w, st),
                 2.0 * M * inner * w, M * (inner + 3 * w) * 2),
}
for name, (fn, flops, bytes_) in runs.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    print(f"{os.environ.get('TTK_LIB_PATH', 'default')[-24:]:24s} {name:9s} {us:7.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  {bytes_ / us / 1e3:7.1f} GB/s")
