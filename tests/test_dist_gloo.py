"""CPU, world_size 2, gloo: the N>1 host logic (rank sharding, per-rank seeds, bucketed gradient
all-reduce, max-over-ranks timing)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from fami_pose_b200 import parallel
    r, w, l = parallel.init_distributed("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    frozen = torch.nn.Linear(3, 3)
    for p in frozen.parameters():
        p.requires_grad = False
    buckets = parallel.GradBuckets(list(lin.parameters()) + list(frozen.parameters()), bucket_bytes=32)
    assert len(buckets.params) == 2 and len(buckets.buckets) == 2      # 60 B weight, 12 B bias
    buckets.zero()
    lin.weight.grad += float(rank + 1)
    lin.bias.grad += float(10 * (rank + 1))
    buckets.allreduce_mean()
    ok = bool(torch.allclose(lin.weight.grad, torch.full_like(lin.weight, 1.5)) and
              torch.allclose(lin.bias.grad, torch.full_like(lin.bias, 15.0)))
    # overlapped form: each bucket's all-reduce is issued from an autograd hook as soon as its last gradient lands;
    # buckets are laid out in reverse parameter order (= the order backward produces gradients)
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Linear(4, 6), torch.nn.Tanh(), torch.nn.Linear(6, 2))
    ob = parallel.GradBuckets(list(net.parameters()), bucket_bytes=64, overlap=True)
    assert [len(b) for b in ob.buckets] == [2, 1, 1] and ob.buckets[0][0] is net[2].bias     # last layer first
    ob.zero()
    xin = torch.full((3, 4), float(rank + 1))
    net(xin).sum().backward()
    first_issued = list(ob.launch_order)
    ob.finish()
    ref = torch.nn.Sequential(torch.nn.Linear(4, 6), torch.nn.Tanh(), torch.nn.Linear(6, 2))
    ref.load_state_dict(net.state_dict())
    tot = [torch.zeros_like(p) for p in ref.parameters()]
    for r in range(world):
        ref.zero_grad()
        ref(torch.full((3, 4), float(r + 1))).sum().backward()
        tot = [t + p.grad / world for t, p in zip(tot, ref.parameters())]
    ok = ok and all(torch.allclose(p.grad, t, atol=1e-6) for p, t in zip(net.parameters(), tot))
    ok = ok and first_issued[0] == 0 and sorted(first_issued) == [0, 1, 2]                       # head bucket went out first
    mx = parallel.max_over_ranks(float(rank) * 2.0, torch.device("cpu"))
    q.put((rank, ok, mx, parallel.shard_range(65, rank, world), parallel.rank_seed(100, rank)))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gradient_allreduce_and_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1]
    assert res[0][2] == 2.0 and res[1][2] == 2.0
    assert res[0][3] == (0, 33) and res[1][3] == (33, 65)
    assert (res[0][4], res[1][4]) == (100, 101)


def test_single_process_helpers():
    from fami_pose_b200 import parallel
    assert parallel.shard_range(256, 3, 8) == (96, 128)
    assert parallel.max_over_ranks(3.0, torch.device("cpu")) == 3.0
