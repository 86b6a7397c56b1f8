"""Where the training step's device time goes: torch.profiler over a few steps of bench.py's --mode train step (1 GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import salience_detr_b200 as pkg
from salience_detr_b200.synthetic import build_model, make_inputs

dev = torch.device("cuda:0")
model = build_model().to(dev).train()
feats, masks, pos = make_inputs("resnet50_800_1333_bs2", seed=0, device=dev)
params = [p for p in model.parameters() if p.requires_grad]
opt = torch.optim.SGD(params, lr=1e-5, foreach=True)


def step():
    opt.zero_grad(set_to_none=False)
    mem, _ = model.forward_encoder(feats, masks, pos)
    loss = mem.square().mean()
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
