import importlib, sys, torch
sys.path.insert(0, ".")
pkg = importlib.import_module("salience-detr_b200")
cabi = pkg.cabi
rows, hidden, balance = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cabi.lib().sdetr_ffn_fused_set_balance(balance)
if len(sys.argv) > 4:
    cabi.lib().sdetr_ffn_fused_set_max_ctas(int(sys.argv[4]))
torch.manual_seed(0)
x = torch.randn(rows, 256, device="cuda")
w1 = torch.randn(hidden, 256, device="cuda") / 16; b1 = torch.randn(hidden, device="cuda")
w2 = torch.randn(256, hidden, device="cuda") / hidden ** 0.5; b2 = torch.randn(256, device="cuda")
s1, s2 = cabi.split_f16_pair(w1), cabi.split_f16_pair(w2)
F = torch.nn.functional
ref = F.linear(F.relu(F.linear(x.double(), w1.double(), b1.double())), w2.double(), b2.double())
y = cabi.ffn_fused_layernorm(x, s1, b1, s2, b2)
torch.cuda.synchronize()
print("max err", (y.double() - ref).abs().max().item(), "ref max", ref.abs().max().item())
