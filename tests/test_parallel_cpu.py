"""world_size-2 gloo tests (CPU) of the data-parallel host logic: bucket coalescing over the flat gradient buffer,
parameter broadcast, unit sharding for the samplers."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _worker(rank, world, port, q):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200.parallel import DataParallel
    from models.U_Net import U_Net
    torch.manual_seed(100 + rank)                      # different init per rank: broadcast must fix it
    net = U_Net(num_resnet_blocks=1, num_layers=2, attn_layers=[1], min_channel=32, max_channel=64, time_dim=32)
    dp = DataParallel(net, bucket_bytes=256 << 10, device=torch.device("cpu"))
    lay = dp.layout
    # 1. replicas identical after construction
    ref = [torch.zeros_like(lay.params_flat) for _ in range(world)]
    dist.all_gather(ref, lay.params_flat)
    same = all(torch.equal(ref[0], r) for r in ref)
    # 2. simulate backward: fill the gradient buffer with rank-dependent values, report ranges in completion order
    lay.flat.copy_(torch.arange(lay.total, dtype=torch.float32) % 97 + rank)
    expect = sum((torch.arange(lay.total, dtype=torch.float32) % 97 + r) for r in range(world))
    eng = net.engine()
    order = [net.out_layers] + list(reversed(net.up_layers)) + [net.middle_layer] + list(reversed(net.down_layers)) + [net.in_layer]
    covered = 0
    for m in order:
        lo, hi = lay.module_range(m)
        covered += hi - lo
        eng.on_grads_ready(lay, lo, hi)
    eng.on_grads_ready(lay, 0, lay.front_end)
    lo, hi = lay.module_range(net.cond_emb)
    covered += (hi - lo) + lay.front_end
    eng.on_grads_ready(lay, lo, hi)
    eng.post_backward(lay)
    ok_sum = torch.equal(lay.flat, expect)
    n_dead = sum(1 for n, p in net.named_parameters() if id(p) not in lay.offsets)
    q.put((rank, same, ok_sum, covered == lay.total, len(dp.launched), n_dead, dp.grad_scale))
    dist.destroy_process_group()


def test_bucketed_allreduce_and_broadcast_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, same, ok_sum, covered, n_launched, n_dead, scale in res:
        assert same, "parameter broadcast failed"
        assert ok_sum, "bucketed all-reduce did not produce the sum"
        assert covered, "reported ranges must tile the whole gradient buffer exactly once"
        assert 2 <= n_launched <= 12          # coalesced into a handful of buckets, not one call per tensor
        assert n_dead == 20                   # y_shift.* / attention norm.* never enter the buffer
        assert scale == 0.5


def test_shard_range_partitions_units():
    from b200.parallel import shard_range
    for total in (0, 1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
