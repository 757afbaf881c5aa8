"""World-size-2 `gloo` tests (CPU) of the N>1 host logic: scene sharding, gradient all-reduce, result gathering."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from room_instance_segmentation_mask3d_b200 import parallel


def test_shard_scenes_balanced_and_complete():
    sizes = [889057, 216152, 869473, 14633, 223648, 55328, 2000000, 794]
    for w in (1, 2, 4, 8):
        owned = parallel.shard_scenes(sizes, w)
        assert sorted(i for o in owned for i in o) == list(range(len(sizes)))  # every scene exactly once
        loads = [sum(sizes[i] for i in o) for o in owned]
        assert max(loads) <= max(max(sizes), 4 / 3 * sum(sizes) / w + 1)  # LPT bound
    assert parallel.shard_scenes([], 4) == [[], [], [], []]
    assert parallel.shard_scenes([5, 5, 5], 2) == [[0, 2], [1]]  # ties -> lowest rank, deterministic


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 2))
        # rank-dependent data; the last layer ("backbone.final"-like) is unused on rank 1 -> no grad there
        x = torch.full((3, 8), float(rank + 1))
        y = model[:3](x)
        if rank == 0:
            y = model[3](y)
        y.sum().backward()
        local = [None if p.grad is None else p.grad.clone() for p in model.parameters()]
        n_coll = parallel.allreduce_gradients(model.parameters(), bucket_bytes=256)  # several small buckets
        grads = [p.grad.clone() for p in model.parameters()]
        s = parallel.allreduce_scalar(float(rank + 1), "cpu")
        owned = parallel.shard_scenes([10, 30, 20], world)[rank]
        res = parallel.gather_scene_results([f"r{rank}s{i}" for i in owned], owned, 3)
        q.put((rank, [None if g is None else g.numpy() for g in local], [g.numpy() for g in grads], n_coll, s, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gradient_allreduce_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(world):
        r = q.get(timeout=100)
        out[r[0]] = r[1:]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    local0, g0, n0, s0, res0 = out[0]
    local1, g1, n1, s1, res1 = out[1]
    assert n0 == n1 and n0 >= 2
    for a, b, l0, l1 in zip(g0, g1, local0, local1):
        np.testing.assert_array_equal(a, b)  # identical on both ranks
        z = np.zeros_like(a)
        want = ((l0 if l0 is not None else z) + (l1 if l1 is not None else z)) / 2
        np.testing.assert_allclose(a, want, rtol=1e-6, atol=1e-7)
    assert local1[-1] is None  # the unused layer had no gradient on rank 1 and still received the averaged one
    assert s0 == s1 == 3.0
    assert res0 == ["r1s0", "r0s1", "r1s2"] and res1 is None


def _bucket_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 2),
                                    torch.nn.Linear(2, 2))  # [4] is never called on any rank ("backbone.final")
        used = [p for n, p in model.named_parameters() if not n.startswith("4.")]
        gb = parallel.GradientBuckets(used, bucket_bytes=256)  # several small buckets
        opt = torch.optim.SGD(used, lr=0.1)
        hist = []
        for step in range(3):
            gb.zero()
            x = torch.full((3, 8), float(rank + 1 + step))
            y = model[:3](x)
            if rank == 0:  # layer [3] only produces a gradient on rank 0
                y = model[3](y)
            y.sum().backward()
            n_coll = gb.finish()
            hist.append(([p.grad.clone().numpy() for p in used], n_coll, [p.detach().clone().numpy() for p in used]))
            opt.step()
        unused_grad_is_none = all(p.grad is None for p in model[4].parameters())
        views = all(p.grad.data_ptr() >= gb.buckets[gb._bucket_of[id(p)]].data_ptr() for p in used)
        # one-shot form: a parameter without gradient on every rank keeps grad None
        lin = torch.nn.Linear(2, 2)
        other = torch.nn.Linear(2, 2)
        lin(torch.ones(1, 2) * (rank + 1)).sum().backward()
        parallel.allreduce_gradients(list(lin.parameters()) + list(other.parameters()))
        q.put((rank, hist, unused_grad_is_none, views, len(gb.buckets), all(p.grad is None for p in other.parameters()),
               lin.weight.grad.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gradient_buckets_overlapped_world2_gloo():
    """GradientBuckets: grads are views into persistent buckets, hooks launch the collective per bucket, ranks end up with the
    same averaged gradients and identical parameters after the optimizer step; never-used parameters keep grad None."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(world):
        r = q.get(timeout=100)
        out[r[0]] = r[1:]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    h0, none0, views0, nb0, other0, lg0 = out[0]
    h1, none1, views1, nb1, other1, lg1 = out[1]
    assert none0 and none1 and views0 and views1 and nb0 == nb1 and nb0 >= 2 and other0 and other1
    np.testing.assert_allclose(lg0, lg1)
    np.testing.assert_allclose(lg0, np.full((2, 2), 1.5), rtol=1e-6)  # mean of x = 1 and x = 2
    for (g0, n0, p0), (g1, n1, p1) in zip(h0, h1):
        assert n0 == n1 == nb0  # every bucket reduced exactly once per step
        for a, b in zip(g0, g1):
            np.testing.assert_array_equal(a, b)
        for a, b in zip(p0, p1):
            np.testing.assert_array_equal(a, b)  # replicas stay in lock step
    # the layer used on rank 0 only: averaged gradient = half of rank 0's (bias gradient of y.sum() is the row count 3 -> 1.5)
    np.testing.assert_allclose(h0[0][0][-1], np.full(2, 1.5), rtol=1e-6)


def test_gradient_buckets_world1_no_process_group():
    model = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.Linear(4, 1))
    gb = parallel.GradientBuckets(model.parameters(), bucket_bytes=32)
    ref = [torch.zeros_like(p) for p in model.parameters()]
    for step in range(2):
        gb.zero()
        x = torch.randn(5, 4)
        model(x).sum().backward()
        assert gb.finish() == 0
        want = torch.autograd.grad(model(x).sum(), list(model.parameters()))
        for p, w in zip(model.parameters(), want):
            torch.testing.assert_close(p.grad, w)
    gb.remove()
