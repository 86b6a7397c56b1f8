"""Per-shape timing of the dense projections of the bench workload: 3xFP16 persistent kernel (sdetr_gemm_f16x3_pre) vs
3xTF32 persistent kernel (sdetr_gemm_3xtf32_pre) vs cuBLAS (3xTF32 split + TF32 GEMM; fp32 SGEMM), with the max-abs
error of each against fp64.  L2 flushed between timed launches, CUDA events on the launching stream."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import salience_detr_b200 as pkg
dev = "cuda:0"
shapes = [  # (name, M, K, N, relu)
    ("enc_output", 44646, 256, 256, 0), ("value_proj x6", 44646, 256, 1536, 0), ("value_proj x1", 44646, 256, 256, 0),
    ("mask_pred l0", 33600, 256, 256, 0),
    ("proj L0", 22726, 256, 384, 0), ("out_proj L0", 22726, 256, 256, 0), ("class L0", 22726, 256, 91, 0),
    ("ffn1 L0", 22726, 256, 2048, 0), ("ffn2 L0", 22726, 2048, 256, 1),
    ("proj L5", 4544, 256, 384, 0), ("ffn1 L5", 4544, 256, 2048, 0), ("ffn2 L5", 4544, 2048, 256, 1),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=10):
    for _ in range(2): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1000)
    return statistics.median(ts)
print(f"{'gemm':14s} {'M':>6s} {'K':>5s} {'N':>5s} | f16x3 us (err) | tf32x3 us (err) | cublas3x us | fp32 us (err) | f16x3 logical TF/s | speedup vs tf32x3")
tot = {}
for name, M, K, N, relu in shapes:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    ref = torch.nn.functional.linear((x.relu() if relu else x).double(), w.double(), b.double())
    r, e = {}, {}
    for mode, own in (("auto", "f16x3"), ("auto", "f16x3-noas"), ("auto", "tf32x3"), ("3xtf32", ""), ("fp32", "")):
        pkg.gemm.MODE, pkg.gemm.OWN_KERNEL = mode, (own or "f16x3").split("-")[0]
        pkg.cabi.lib().sdetr_gemm_f16x3_set_as(0)
        pkg.cabi.lib().sdetr_gemm_f16x3_set_epilogue_warps(8 if own.endswith("noas") else 4)
        key = own or mode
        fn = lambda: pkg.gemm.linear(x, w, b, relu_input=bool(relu))
        e[key] = (fn().double() - ref).abs().max().item()
        r[key] = timeit(fn)
        tot[key] = tot.get(key, 0) + r[key]
    print(f"{name:14s} {M:6d} {K:5d} {N:5d} | {r['f16x3']:7.1f} ({e['f16x3']:.1e}) [8 epilogue warps {r['f16x3-noas']:7.1f}] | {r['tf32x3']:7.1f} ({e['tf32x3']:.1e}) | {r['3xtf32']:8.1f} | "
          f"{r['fp32']:7.1f} ({e['fp32']:.1e}) | {2*M*N*K/r['f16x3']/1e6:7.1f} | {r['tf32x3']/r['f16x3']:.2f}x")
print("sum us", {k: round(v, 1) for k, v in tot.items()})
