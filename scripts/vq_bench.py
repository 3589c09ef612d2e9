"""Quantizer microbench (BASELINE.json configs[1]): N = 2^20 latent vectors, codebooks 1K..64K.
Prints one JSON line per (K, D): time of ttk_vq_argmin (distance GEMM + fused argmin), algorithmic TFLOP/s =
2*N*K*D / t (un-padded D) and the MMA TFLOP/s actually executed (padded augmented D), both against the measured
bf16 peak; plus the FSQ closed-form kernel in GB/s against the measured HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from titok_video_b200 import _lib
from titok_video_b200.engine import _ptr, _stream, _vp
import titok_video_b200 as T


def peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def time_ms(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def vq_case(N, K, D, dev, cb=None):
    g = torch.Generator().manual_seed(K + D)
    D8 = (D + 7) // 8 * 8
    z = torch.zeros((N, D8), dtype=torch.bfloat16, device=dev)
    z[:, :D] = (torch.randn((N, D), generator=g) * 2).to(torch.bfloat16).to(dev)
    if cb is None:
        cb = torch.randn((K, D), generator=g)
    cbd = cb.to(torch.bfloat16).to(dev).contiguous()
    DA = _lib.fn("ttk_vq_aug_dim")(D)
    aug = torch.empty((_lib.fn("ttk_vq_aug_rows")(K, D), DA), dtype=torch.bfloat16, device=dev)
    st = _stream()
    _lib.call("ttk_vq_prepare_codebook", _ptr(cbd), D, K, D, _ptr(aug), DA, st)
    idx = torch.empty((N,), dtype=torch.int32, device=dev)
    ms = time_ms(lambda: _lib.call("ttk_vq_argmin", _ptr(z), D8, _ptr(aug), DA, N, K, D, _ptr(idx), _vp(0), st))
    return ms, DA


def main(quick=False, quiet=False):
    dev = torch.device("cuda:0")
    hbm, tf, tf_sus, src = peaks()
    N = 1 << 20
    out = []
    fsq_cb = T.FSQ([7, 5, 5, 5, 5]).implicit_codebook
    cases = [(4375, 5, fsq_cb)] + [(K, D, None) for D in ((64, 128, 256) if not quick else (128,))
                                   for K in ((1024, 4096, 16384, 65536) if not quick else (4096, 65536))]
    for K, D, cb in cases:
        ms, DA = vq_case(N, K, D, dev, cb)
        kpad = D if D % 64 == 0 else (DA + 15) // 16 * 16  # D % 64 == 0: the norms are added in the epilogue, no extra k block
        rec = {"kernel": "ttk_vq_argmin", "N": N, "K": K, "D": D, "ms": ms,
               "tflops_algorithmic": 2.0 * N * K * D / (ms * 1e-3) / 1e12,
               "tflops_mma_issued": 2.0 * N * ((K + 255) // 256 * 256) * kpad / (ms * 1e-3) / 1e12}
        rec["frac_of_tensor_peak"] = rec["tflops_algorithmic"] / tf
        rec["mma_frac_of_tensor_peak"] = rec["tflops_mma_issued"] / tf
        rec["peak"] = tf
        rec["peak_source"] = src + " (burst bf16: kernel timed alone)"
        rec["frac_of_sustained_tensor_peak"] = rec["tflops_algorithmic"] / tf_sus  # the 10-launch loop of a >= 10 ms kernel runs under the power cap
        out.append(rec)
        if not quiet:
            print(json.dumps(rec), flush=True)
    # FSQ closed form (what the reference's quantizer actually computes)
    # kernel-only timing through the C ABI on preallocated buffers, 2^24 vectors so the launch is not latency-bound
    q = T.FSQ([7, 5, 5, 5, 5]).to(dev)
    consts = q._consts(dev)
    NF = 1 << 24
    st = _stream()
    for dt, code, bpv in ((torch.bfloat16, 0, 24), (torch.float32, 1, 44)):
        z = (torch.randn((NF, 5), device=dev) * 2).to(dt)
        codes = torch.empty_like(z)
        idx = torch.empty((NF,), dtype=torch.int32, device=dev)
        ms = time_ms(lambda: _lib.call("ttk_fsq_fwd", _ptr(z), _ptr(codes), _ptr(idx), NF, code, 5, *consts, st))
        rec = {"kernel": "ttk_fsq_fwd", "N": NF, "dtype": str(dt).split(".")[-1], "ms": ms, "gbs": NF * bpv / (ms * 1e-3) / 1e9,
               "frac_of_hbm_peak": NF * bpv / (ms * 1e-3) / 1e9 / hbm, "peak": hbm, "peak_source": src}
        out.append(rec)
        if not quiet:
            print(json.dumps(rec), flush=True)
    return out


if __name__ == "__main__":
    main("--quick" in sys.argv)
