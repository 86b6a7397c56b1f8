"""Fused FFN against the two-GEMM + add_layernorm path on the six layer row counts of config 2 (cold L2 between launches)."""
import importlib
import sys

import torch

sys.path.insert(0, ".")
pkg = importlib.import_module("salience-detr_b200")
cabi = pkg.cabi
dev = "cuda"
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=10):
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(400000)  # the host queues the launches while the device spins: no launch gaps inside the events
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts[2:])[len(ts[2:]) // 2]


hidden = 2048
w1 = torch.randn(hidden, 256, device=dev) / 16; b1 = torch.randn(hidden, device=dev)
w2 = torch.randn(256, hidden, device=dev) / 45; b2 = torch.randn(256, device=dev)
gamma, beta = torch.ones(256, device=dev), torch.zeros(256, device=dev)
s1, s2 = cabi.split_f16_pair(w1), cabi.split_f16_pair(w2)
rows_list = [int(a) for a in sys.argv[1:]] or [22726, 18181, 13636, 13636, 9090, 4545]
print("rows | two GEMMs + add_layernorm us | fused: balanced ranges us | whole panels us | 3-pass TFLOP/s (balanced)")
tot_a = tot_b = tot_c = 0.0
lib = cabi.lib()
for rows in rows_list:
    x = torch.randn(rows, 256, device=dev)

    def base():
        h = cabi.gemm_f16x3_pre(x, *s1, b1, 0)
        f = cabi.gemm_f16x3_pre(h, *s2, b2, 1)
        return cabi.add_layernorm(x, f, gamma, beta, 1e-5)

    t0 = timed(base)
    res = {}
    for balance in (1, 0):
        lib.sdetr_ffn_fused_set_balance(balance)
        res[balance] = timed(lambda: cabi.ffn_fused_layernorm(x, s1, b1, s2, b2, gamma, beta, 1e-5))
    lib.sdetr_ffn_fused_set_balance(1)
    tot_a += t0; tot_b += res[1]; tot_c += res[0]
    print(f"{rows:6d} | {t0:7.1f} | {res[1]:7.1f} | {res[0]:7.1f} | {rows * 256 * hidden * 4 * 3 / res[1] / 1e6:6.0f}")
print(f"sum: two-GEMM path {tot_a:.1f} us, fused balanced {tot_b:.1f} us, fused whole panels {tot_c:.1f} us")
