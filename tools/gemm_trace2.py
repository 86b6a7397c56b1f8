"""clock64() trace of CTA 0 of the 3xFP16 streaming GEMM: per k-block, when the producer issued the TMA, when the converters saw
it land / finished, when the MMA warp started issuing; per tile, when the accumulator completed and when the epilogue was done.
Usage: python tools/gemm_trace2.py M K N"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import salience_detr_b200 as pkg
M, K, N = (int(v) for v in sys.argv[1:4])
dev = "cuda:0"
x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
hi, lo, sc = pkg.cabi.split_f16_pair(w)
for _ in range(3):
    pkg.cabi.gemm_f16x3_pre(x, hi, lo, sc, b)
buf = torch.zeros(10 * 256, dtype=torch.int64, device=dev)
pkg.cabi.lib().sdetr_gemm_f16x3_set_trace(buf.data_ptr())
pkg.cabi.gemm_f16x3_pre(x, hi, lo, sc, b)
torch.cuda.synchronize()
pkg.cabi.lib().sdetr_gemm_f16x3_set_trace(None)
t = buf.view(10, 256).cpu()
nk = K // 64
t0 = int(t[0, 0])
names = ["tma_issue", "landed", "conv_done", "mma_issue"]
print(f"M={M} K={K} N={N}: k-blocks per tile {nk}; times in clk relative to the first TMA issue")
print("kb   " + " ".join(f"{n:>10s}" for n in names) + "   d(issue->landed) d(landed->conv) d(conv->mma) d(mma->next mma)")
n = min(int((t[3] > 0).sum()), 48)
for i in range(n):
    r = [int(t[e, i]) - t0 for e in range(4)]
    nxt = int(t[3, i + 1]) - t0 if i + 1 < n else r[3]
    print(f"{i:3d}  " + " ".join(f"{v:10d}" for v in r) + f"   {r[1]-r[0]:8d} {r[2]-r[1]:8d} {r[3]-r[2]:8d} {nxt-r[3]:8d}" + ("   <- tile boundary" if (i + 1) % nk == 0 else ""))
nt = min(int((t[5] > 0).sum()), 12)
print("tile  acc_complete  epilogue_done  (duration)")
for i in range(nt):
    a, e = int(t[5, i]) - t0, int(t[6, i]) - t0
    print(f"{i:3d}  {a:10d} {e:10d}  {e - a:8d}")
print("epilogue detail per 32-column block: after TMEM load | after box-free wait + barrier | after box write + barrier   (clk since accumulator complete)")
for i in range(min(nt, 8)):
    base = int(t[5, i])
    print(f"tile {i}: " + "  ".join(f"[{int(t[7, 4*i+c]) - base:6d} {int(t[8, 4*i+c]) - base:6d} {int(t[9, 4*i+c]) - base:6d}]" for c in range(4)))
