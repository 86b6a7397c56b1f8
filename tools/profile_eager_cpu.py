"""Host-side cost of the EAGER encoder half (the path a batch with never-seen padding masks takes): cProfile over N forwards."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import salience_detr_b200 as pkg
from salience_detr_b200.synthetic import build_model, make_inputs

dev = torch.device("cuda:0")
model = build_model().to(dev)
feats, masks, pos = make_inputs("resnet50_800_1333_bs2", seed=0, device=dev)
N = 30
with torch.no_grad():
    plan = model.make_plan(masks)
    for _ in range(3):
        model.forward_encoder(feats, masks, pos, plan=plan)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(N):
        model.forward_encoder(feats, masks, pos, plan=plan)
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f"eager forward: host issue {1e3 * t_issue / N:.3f} ms per step, wall {1e3 * t_all / N:.3f} ms per step, launches per step {pkg.cabi.launch_count() // (N + 3)}")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(N):
        model.forward_encoder(feats, masks, pos, plan=plan)
    pr.disable()
    torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
