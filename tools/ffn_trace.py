"""clock64() trace of CTA 0 of the fused FFN kernel.  Usage: python tools/ffn_trace.py ROWS [HIDDEN [SPLIT]]"""
import importlib, sys, torch
sys.path.insert(0, ".")
pkg = importlib.import_module("salience-detr_b200")
cabi = pkg.cabi
rows = int(sys.argv[1]); hidden = int(sys.argv[2]) if len(sys.argv) > 2 else 2048; balance = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cabi.lib().sdetr_ffn_fused_set_balance(balance)
x = torch.randn(rows, 256, device="cuda")
w1 = torch.randn(hidden, 256, device="cuda") / 16; b1 = torch.randn(hidden, device="cuda")
w2 = torch.randn(256, hidden, device="cuda") / hidden ** 0.5; b2 = torch.randn(256, device="cuda")
s1, s2 = cabi.split_f16_pair(w1), cabi.split_f16_pair(w2)
for _ in range(3):
    cabi.ffn_fused_layernorm(x, s1, b1, s2, b2)
buf = torch.zeros(8 * 256, dtype=torch.int64, device="cuda")
cabi.lib().sdetr_ffn_fused_set_trace(buf.data_ptr())
cabi.ffn_fused_layernorm(x, s1, b1, s2, b2)
torch.cuda.synchronize()
cabi.lib().sdetr_ffn_fused_set_trace(None)
t = torch.cat([buf.view(8, 256).cpu(), torch.zeros(8, 64, dtype=torch.int64)], 1)
t0 = int(t[6, 0])
chunks = hidden // 128
panels = (rows + 127) // 128
G = min(148, (panels * chunks + 1) // 2 if balance else panels)
hi0 = panels * chunks // G if balance else (panels // G) * chunks
units, pos = [], 0
while pos < hi0:
    n = min(hi0, (pos // chunks + 1) * chunks) - pos
    units.append(n); pos += n
print(f"rows={rows} hidden={hidden} balance={balance}: CTA 0 has {hi0} chunks in units of {units}; clk relative to the first TMA issue")
print("ring slot uses: it | kind | TMA issued | consumer (MMA) found it landed, issuing | delta to previous MMA issue")
prev = None
it = 0
rowsout = []
for unit, nc in enumerate(units):
    for j in range(nc + 1):
        if j < nc:
            for kb in range(4):
                rowsout.append((it, f"u{unit} W1 c{j} kb{kb}", int(t[6, it]) - t0, int(t[0, it]) - t0)); it += 1
        if j > 0:
            for k in range(4):
                rowsout.append((it, f"u{unit} W2 c{j-1} t{k}", int(t[6, it]) - t0, int(t[1, it]) - t0)); it += 1
        if it >= 250:
            break
    if it >= 250:
        break
for i, kind, a, b in rowsout:
    if i >= 256 or a < 0:
        break
    d = "" if b is None or prev is None else f"{b - prev:7d}"
    print(f"{i:4d} {kind:14s} {a:9d} " + (f"{b:9d} {d}" if b is not None else ""))
    if b is not None:
        prev = b
print("chunk | converters: accumulator-1 complete | hidden slot free (after split math) | hidden written   (d = since acc complete)")
for gi in range(min(hi0, 40)):
    a, b, c = (int(t[e, gi]) - t0 for e in (2, 3, 4))
    if a < 0: break
    print(f"{gi:4d}  {a:9d}  {b:9d} (+{b - a:5d})  {c:9d} (+{c - a:5d})")
print("unit: output accumulator complete:", [int(v) - t0 for v in t[5, :4] if int(v) > 0])
print("x panel per unit (TMA issued, first k-block landed, panel split):", [tuple(int(t[7, 4 * uu + e]) - t0 for e in range(3)) for uu in range(3)])
