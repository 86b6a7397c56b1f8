"""Debug aid: poison the caching allocator's memory with a byte pattern, then run the 5-scale (or default) encoder half.
Finds reads of memory the path never wrote (they turn into NaNs / wild indices instead of stale-but-valid data).
    CUDA_LAUNCH_BLOCKING=1 python tools/poison_run.py [pattern_hex] [config]"""
import sys
import os
import traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import salience_detr_b200 as pkg
from salience_detr_b200.synthetic import build_model, make_inputs

pattern = int(sys.argv[1], 16) if len(sys.argv) > 1 else 0xFF
config = sys.argv[2] if len(sys.argv) > 2 else "resnet50_5scale_bs2"
dev = torch.device("cuda:0")
strides = (4, 8, 16, 32) if "5scale" in config else (8, 16, 32)
model = build_model(strides=strides).to(dev)
feats, masks, pos = make_inputs(config, seed=1, device=dev)
torch.cuda.synchronize()
free, _ = torch.cuda.mem_get_info()
big = [torch.empty(int(free * 0.45), dtype=torch.uint8, device=dev).fill_(pattern) for _ in range(2)]
small = [torch.empty(s, dtype=torch.uint8, device=dev).fill_(pattern)
         for s in (512, 4096, 65536, 524288, 1 << 20) for _ in range(400)]
torch.cuda.synchronize()
del big, small
print("poisoned with 0x%02x, reserved %.1f GB" % (pattern, torch.cuda.memory_reserved() / 1e9), flush=True)
for it in range(3):
    try:
        with torch.no_grad():
            mem, aux = model.forward_encoder(feats, masks, pos)
        torch.cuda.synchronize()
        print("iter", it, "ok finite", bool(torch.isfinite(mem).all()), "K", aux["plan"].num_selected,
              "checksum %.6f" % mem.double().sum().item(), flush=True)
    except Exception:
        traceback.print_exc()
        break
