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
print("small_attn mha_tc gemm_mode | ms/step")
for sa, tc, mode in itertools.product((False, True), (False, True), ("auto", "3xtf32")):
    st.SMALL_ATTENTION, st.MHA_GEMM_TENSOR_CORE, pkg.gemm.MODE = sa, tc, mode
    r = EncoderRunner(model, feats, masks, pos)
    print(f"{int(sa)} {int(tc)} {mode:7s} | {step_ms(r):.3f}")
    del r
