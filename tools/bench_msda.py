"""Sweep of the fused MSDA forward kernel variants on the bench workload (layer inputs captured from a real
encoder forward): resident CTAs/SM, schedule, processing order, chunk, cell size.  Prints a table; used to pick
the defaults (results summarised under profiles/)."""
import itertools
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import salience_detr_b200 as pkg  # noqa: E402
import salience_detr_b200.salience_transformer as st  # noqa: E402
from salience_detr_b200.synthetic import build_model, make_inputs  # noqa: E402

dev = torch.device("cuda:0")
workload = sys.argv[1] if len(sys.argv) > 1 else "resnet50_800_1333_bs2"
model = build_model().to(dev)
feats, masks, pos = make_inputs(workload, seed=0, device=dev)
cabi = pkg.cabi
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def capture(cell_px, use_order=True):
    st.TILE_CELL_PX = cell_px
    calls = []
    orig = cabi.msda_fused_forward

    def spy(*a, **k):
        calls.append((a, k))
        return orig(*a, **k)

    cabi.msda_fused_forward = spy
    try:
        with torch.no_grad():
            plan = model.make_plan(masks)
            model.forward_encoder(feats, masks, pos, plan=plan, use_order=use_order)
    finally:
        cabi.msda_fused_forward = orig
    torch.cuda.synchronize()
    return calls, plan


def time_calls(calls, schedule, use_order, reps=15):
    per = [[] for _ in calls]
    for _ in range(reps):
        flush.zero_()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(calls) + 1)]
        evs[0].record()
        for i, (a, k) in enumerate(calls):
            a = list(a)
            a[13] = a[13] if use_order else None   # query_order
            a[14] = schedule
            cabi.msda_fused_forward(*a, **k)
            evs[i + 1].record()
        torch.cuda.synchronize()
        for i in range(len(calls)):
            per[i].append(evs[i].elapsed_time(evs[i + 1]) * 1000)
    return [statistics.median(x) for x in per]


def alg_bytes(plan):
    b, nv = plan.mask_flat.shape
    return [b * (4 * nv * 256 + nq * (12 * 128 + 4 * 256)) for nq in plan.layer_num_query]


ref_out = None
print(f"workload {workload}")
print("cell  order sched minb chunk | per-layer us | total us | GB/s")
for cell in (64,):
    calls, plan = capture(cell)
    byts = alg_bytes(plan)
    for use_order, schedule, minb in itertools.product((True,), (1,), (41, 31)):
        if not use_order and cell != 64:
            continue
        for chunk in ((64,) if schedule == 0 else (64, 128, 256)):
            cabi.set_option("msda_smem_broadcast", 1 if minb in (41, 31) else 0)  # 41 / 31 = minb 4 / 3 + shared-memory broadcast
            cabi.set_option("msda_min_blocks", 4 if minb in (41, 31) else minb)
            cabi.set_option("msda_chunk", chunk)
            t = time_calls(calls, schedule, use_order)
            # correctness of every variant against the first one
            a = list(calls[0][0]); a[13] = a[13] if use_order else None; a[14] = schedule
            out = cabi.msda_fused_forward(*a, **calls[0][1])
            if ref_out is None:
                ref_out = out.clone()
            err = (out - ref_out).abs().max().item()
            print(f"{cell:4d}  {int(use_order)}     {schedule}     {minb}    {chunk:4d} | " +
                  " ".join(f"{x:6.1f}" for x in t) + f" | {sum(t):7.1f} | {sum(byts) / sum(t) / 1e3:7.1f}  err {err:.1e}")
