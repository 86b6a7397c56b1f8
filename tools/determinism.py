"""Debug aid: run-to-run bitwise determinism of the GEMM kernels and of the whole encoder half."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import salience_detr_b200 as pkg
from salience_detr_b200 import cabi, gemm
from salience_detr_b200.synthetic import build_model, make_inputs

dev = torch.device("cuda:0")
config = sys.argv[1] if len(sys.argv) > 1 else "resnet50_5scale_bs2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
torch.manual_seed(0)
for (m, n, k, act) in [(178500, 256, 256, 0), (90660, 256, 2048, 1), (90660, 384, 256, 0), (35700, 256, 256, 0), (600, 768, 256, 0)]:
    a = torch.randn(m, k, device=dev)
    w = torch.randn(n, k, device=dev) * 0.05
    bias = torch.randn(n, device=dev)
    first = None
    bad = 0
    for r in range(reps):
        junk = torch.empty(m, n, device=dev).fill_(float("nan"))
        out = cabi.gemm_3xtf32_raw(a, w, bias, act=act)
        if first is None:
            first = out.clone()
            ref = torch.nn.functional.linear((a.relu() if act == 1 else a).double(), w.double(), bias.double())
            print("gemm", m, n, k, "err vs fp64 %.2e" % (out.double() - ref).abs().max().item(), flush=True)
        elif not torch.equal(out, first):
            d = (out != first).nonzero()
            bad += 1
            print("  rep", r, "differs at", d.shape[0], "elements; rows", d[:, 0].min().item(), "..", d[:, 0].max().item(),
                  "cols", d[:, 1].min().item(), "..", d[:, 1].max().item(), "max diff %.3e" % (out - first).abs().max().item(), flush=True)
        del junk
    print("gemm", m, n, k, "nondeterministic reps:", bad, "/", reps, flush=True)

strides = (4, 8, 16, 32) if "5scale" in config else (8, 16, 32)
model = build_model(strides=strides).to(dev)
feats, masks, pos = make_inputs(config, seed=1, device=dev)
for mode in ("auto", "fp32", "3xtf32"):
    gemm.MODE = mode
    first = None
    bad = 0
    for r in range(reps):
        with torch.no_grad():
            mem, aux = model.forward_encoder(feats, masks, pos)
        cur = {"mem": mem.clone(), "raw": aux["raw_score"].clone(), "inds": aux["selected_inds"].clone(), "fg": aux["foreground_score"].clone()}
        if first is None:
            first = cur
            continue
        for key in cur:
            if not torch.equal(cur[key], first[key]):
                d = (cur[key] != first[key]).nonzero()
                bad += 1
                print("  mode", mode, "rep", r, key, "differs at", d.shape[0], "elements, first", d[0].tolist(), "last", d[-1].tolist(),
                      "max diff %.3e" % (cur[key].double() - first[key].double()).abs().max().item(), flush=True)
    print("encoder", config, mode, "nondeterministic outputs:", bad, flush=True)
