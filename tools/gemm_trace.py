"""Pipeline trace of the tcgen05 GEMM (CTA (0,0)): clock64() at producer / converter / MMA events per k-block."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import salience_detr_b200 as pkg
dev = "cuda:0"
lib = pkg.cabi.lib()
M, K, N = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (16896, 256, 256)))
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 0
lib.sdetr_gemm_set_variant(variant)
if variant >= 2:  # raw-weight kernels (TS2 / persistent)
    pkg.cabi.gemm_3xtf32 = lambda x, hi, lo: pkg.cabi.gemm_3xtf32_raw(x, w)
x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / 16
hi, lo = pkg.cabi.split_tf32_pair(w)
for _ in range(3): pkg.cabi.gemm_3xtf32(x, hi, lo)
buf = torch.zeros(6 * 128, dtype=torch.int64, device=dev)
lib.sdetr_gemm_set_trace(buf.data_ptr())
pkg.cabi.gemm_3xtf32(x, hi, lo); torch.cuda.synchronize()
lib.sdetr_gemm_set_trace(None)
t = buf.cpu().view(6, 128); t0 = int(t[5, 2]); nk = K // 32
print(f"M={M} K={K} N={N}; cycles relative to setup-done")
print("kb  tma_issue  tma_landed(mma)  conv_start  conv_done  mma_issue")
for kb in range(nk):
    print(f"{kb:2d} {int(t[0,kb])-t0:9d} {int(t[1,kb])-t0:14d} {int(t[3,kb])-t0:11d} {int(t[4,kb])-t0:10d} {int(t[2,kb])-t0:10d}")
print("epilogue start/end", int(t[5,0])-t0, int(t[5,1])-t0)
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20): pkg.cabi.gemm_3xtf32(x, hi, lo)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 20
ref = (x.double() @ w.double().t())
print("max abs err vs fp64", (pkg.cabi.gemm_3xtf32(x, hi, lo).double() - ref).abs().max().item())
print(f"{ms*1000:.1f} us/launch, {2*3*M*N*K/ms/1e9:.1f} TF32-TFLOP/s ({2*M*N*K/ms/1e9:.1f} fp32-equivalent)")
