"""world_size-2 `gloo` tests (CPU) of the N>1 host logic: max-over-ranks timing, per-rank batches, gradient
all-reduce.  The forward itself has no collective (independent replicas)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from salience_detr_b200 import dist as sd
    from salience_detr_b200.synthetic import make_inputs
    assert sd.init_from_env("gloo")
    assert sd.env_rank_world() == (rank, world, rank)
    # slowest rank decides
    t = sd.max_over_ranks(10.0 + rank)
    # every rank draws a different batch of the same shape
    feats, masks, _ = make_inputs("cpu_512", embed_dim=8, seed=sd.shard_batch_seed(0, rank))
    chk = torch.tensor([float(feats[0].sum())], dtype=torch.float64)
    gathered = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(gathered, chk)
    # gradient mean all-reduce in buckets
    p1, p2 = torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(3, 2))
    p1.grad, p2.grad = torch.full((5,), float(rank + 1)), torch.full((3, 2), float(10 * (rank + 1)))
    sd.allreduce_gradients_([p1, p2], bucket_bytes=16)
    # overlapped bucketed all-reduce: gradients are views into the communication buffers, buckets launch from hooks
    torch.manual_seed(0)  # same weights on both ranks, different data
    net = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.ReLU(), torch.nn.Linear(8, 4), torch.nn.Linear(4, 2))
    unused = torch.nn.Parameter(torch.ones(3))                    # never reached by the loss
    params = [unused] + list(net.parameters())                    # lands in the last bucket
    gb = sd.GradientBuckets(params, bucket_bytes=200)             # several small buckets
    x = torch.randn(5, 6, generator=torch.Generator().manual_seed(100 + rank))
    local = []
    for step in range(2):                                         # buffers are reused across steps
        gb.zero_()
        net(x * (step + 1)).square().mean().backward()
        local.append([p.grad.clone() for p in net.parameters()])  # hooks already fired: this is pre-finish() local+maybe-reduced
        gb.finish()
    # reference: plain autograd per rank + mean over ranks via all_gather
    net2 = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.ReLU(), torch.nn.Linear(8, 4), torch.nn.Linear(4, 2))
    net2.load_state_dict(net.state_dict())
    net2(x * 2).square().mean().backward()
    ok = True
    for p, q in zip(net.parameters(), net2.parameters()):
        parts = [torch.zeros_like(q.grad) for _ in range(world)]
        dist.all_gather(parts, q.grad)
        ok = ok and torch.allclose(p.grad, sum(parts) / world, atol=1e-6)
    ok = ok and unused.grad is not None and float(unused.grad.abs().max()) == 0.0
    ok = ok and len(gb.buckets) >= 2 and sorted(gb.launch_order) == list(range(len(gb.buckets)))
    ok = ok and gb.launch_order[0] == 0   # the last layer's bucket (reverse registration order) is launched first
    out[rank] = (t, [float(g) for g in gathered], p1.grad.tolist(), p2.grad.flatten().tolist(),
                 sd.aggregate_throughput(2, 20, world, 1000.0 * t), bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0][0] == out[1][0] == 11.0                       # max over ranks
    assert out[0][1] == out[1][1] and out[0][1][0] != out[0][1][1]  # different batches, consistently gathered
    assert out[0][2] == [1.5] * 5 and out[1][3] == [15.0] * 6   # mean of (1,2) and (10,20)
    assert abs(out[0][4] - 2 * 20 * 2 / 11.0) < 1e-9
    assert out[0][5] and out[1][5]                              # GradientBuckets == plain autograd + mean over ranks
