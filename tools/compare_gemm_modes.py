"""Encoder-half output of the bench workload under the three projection modes (fp32 SGEMM / 3xTF32 / TF32):
max-abs differences of scores and memory and overlap of the selected indices, fp32 mode as the reference."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import salience_detr_b200 as pkg  # noqa: E402
from salience_detr_b200.synthetic import build_model, make_inputs  # noqa: E402

dev = torch.device("cuda:0")
model = build_model().to(dev)
feats, masks, pos = make_inputs("resnet50_800_1333_bs2", seed=0, device=dev)
res = {}
with torch.no_grad():
    plan = model.make_plan(masks)
    for mode in ("fp32", "auto", "tcgen05", "3xtf32", "tf32"):
        pkg.gemm.MODE = mode
        mem, aux = model.forward_encoder(feats, masks, pos, plan=plan)
        res[mode] = (mem.clone(), aux["raw_score"].clone(), aux["selected_inds"].clone())
        # same selection injected: isolates the numeric error of the layers from index flips
        inds = res["fp32"][2]
        feat = pkg.flatten_levels(feats)
        lpos = pkg.flatten_levels([p + e.view(1, -1, 1, 1) for p, e in zip(pos, model.level_embeds)])
        fg = torch.where(plan.mask_flat, res["fp32"][1].min(), res["fp32"][1])
        mem_inj = model.encoder(query=feat, query_pos=lpos, query_key_padding_mask=plan.mask_flat,
                                spatial_shapes=plan.spatial_shapes, level_start_index=plan.level_start_index,
                                valid_ratios=plan.valid_ratios, foreground_score=fg, focus_token_nums=plan.focus_token_nums,
                                foreground_inds=[inds[:, :n] for n in plan.layer_num_query], multi_level_masks=masks)
        res[mode] += (mem_inj.clone(),)
ref = res["fp32"]
print("mode     | score max-abs | same index at same rank | same selected set | memory max-abs | memory max-abs (same indices)")
for mode in ("fp32", "auto", "tcgen05", "3xtf32", "tf32"):
    mem, raw, inds, mem_inj = res[mode]
    same_rank = (inds == ref[2]).float().mean().item()
    same_set = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(inds, ref[2])) / inds.numel()
    print(f"{mode:8s} | {(raw - ref[1]).abs().max().item():.3e} | {same_rank:.4f} | {same_set:.4f} | "
          f"{(mem - ref[0]).abs().max().item():.3e} | {(mem_inj - ref[3]).abs().max().item():.3e}")
