"""Step time of the encoder half (CUDA-graph replay) under implementation switches, to pick defaults."""
import itertools, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import salience_detr_b200 as pkg
import salience_detr_b200.salience_transformer as st
from salience_detr_b200.runner import EncoderRunner
from salience_detr_b200.synthetic import build_model, make_inputs
dev = torch.device("cuda:0")
model = build_model().to(dev)
feats, masks, pos = make_inputs("resnet50_800_1333_bs2", seed=0, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def step_ms(runner, n=20):
    ts = []
    for _ in range(5): runner.step()
    for _ in range(n):
        with torch.cuda.stream(runner.stream):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(runner.stream); runner.step(); b.record(runner.stream)
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)
base = dict(FUSED_QUERY_SUM=True, FUSED_PRE_ATTENTION=True, FUSED_GELU_MEAN=True, PRESPLIT_PERSISTENT=True, SMALL_M=2304, OVERLAP_VALUE_PROJ=True)
variants = [("defaults", {}), ("separate q+pos add", dict(FUSED_QUERY_SUM=False)), ("defaults (2)", {}), ("separate q+pos add (2)", dict(FUSED_QUERY_SUM=False)),
            ("library pre-attention", dict(FUSED_PRE_ATTENTION=False)), ("torch gelu+mean", dict(FUSED_GELU_MEAN=False)),
            ("in-kernel weight split", dict(PRESPLIT_PERSISTENT=False)), ("no small-M fp32 route", dict(SMALL_M=0)),
            ("no value-proj overlap", dict(OVERLAP_VALUE_PROJ=False)), ("defaults again", {})]
print("variant | ms/step | images/s")
for name, over in variants:
    cfg = dict(base, **over)
    st.FUSED_PRE_ATTENTION, st.FUSED_GELU_MEAN, st.OVERLAP_VALUE_PROJ = cfg["FUSED_PRE_ATTENTION"], cfg["FUSED_GELU_MEAN"], cfg["OVERLAP_VALUE_PROJ"]
    st.FUSED_QUERY_SUM = cfg["FUSED_QUERY_SUM"]
    pkg.gemm.PRESPLIT_PERSISTENT, pkg.gemm.SMALL_M = cfg["PRESPLIT_PERSISTENT"], cfg["SMALL_M"]
    r = EncoderRunner(model, feats, masks, pos)
    ms = step_ms(r)
    print(f"{name:26s} | {ms:.3f} | {2000.0 / ms:.1f}")
    del r
