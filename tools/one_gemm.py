"""One dense projection through gemm.linear (default mode), for an ncu capture:  python tools/one_gemm.py M K N [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import salience_detr_b200 as pkg
M, K, N = (int(v) for v in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
x = torch.randn(M, K, device="cuda:0"); w = torch.randn(N, K, device="cuda:0") / K ** 0.5; b = torch.randn(N, device="cuda:0")
for _ in range(reps):
    y = pkg.gemm.linear(x, w, b)
torch.cuda.synchronize()
print("ok", tuple(y.shape), pkg.gemm.MODE)
