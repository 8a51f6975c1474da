"""World-size-2 `gloo` tests (CPU) of the data-parallel host logic in video_vae_b200/ddp.py: flat parameter /
gradient re-homing, reverse-order bucketing, per-bucket all-reduce triggered as backward Functions report their
parameters, and the "replicas stay identical" property the reference checks in
claude_distributed/test_distributed.py:159-163.  No libvvae compute runs here (there is no CPU path); the gradients
are synthetic."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(5)
        self.a = torch.nn.Parameter(torch.randn(7, 5, generator=g))
        self.b = torch.nn.Parameter(torch.randn(13, generator=g))
        self.c = torch.nn.Parameter(torch.randn(3, 3, 3, generator=g))
        self.d = torch.nn.Parameter(torch.randn(64, generator=g))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from video_vae_b200 import functional as F_
        from video_vae_b200.ddp import FlatParams, GradAllReducer
        model = _Toy()
        ref = {n: p.detach().clone() for n, p in model.named_parameters()}
        flat = FlatParams(model)
        # re-homing keeps values, makes params/grads views of the flat buffers, 8-element aligned offsets
        for n, p in model.named_parameters():
            assert torch.equal(p.detach(), ref[n])
            assert p.grad is not None and p.grad.shape == p.shape
        assert all(o % 8 == 0 for o in flat.offsets) and flat.total % 8 == 0
        # start-up replication: rank 1 perturbs its copy, the broadcast restores rank 0's values everywhere
        if rank == 1:
            with torch.no_grad():
                flat.flat.add_(1.0)
        flat.broadcast(src=0)
        for n, p in model.named_parameters():
            assert torch.equal(p.detach(), ref[n]), n
        red = GradAllReducer(flat, bucket_bytes=128)           # tiny buckets -> several of them
        covered = sorted((s, e) for s, e, _ in red.buckets)
        assert covered[0][0] == 0 and covered[-1][1] == flat.total
        assert all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))
        assert sum(n for _, _, n in red.buckets) == len(flat.params)
        assert red.buckets[0][1] == flat.total                 # bucket 0 holds the LAST parameters (backward order)

        params = list(model.parameters())
        for step in range(3):
            flat.zero_grad()
            red.start_step()
            for idx in range(len(params) - 1, -1, -1):         # "backward": last layer first
                p = params[idx]
                p.grad.add_(torch.full_like(p, float((rank + 1) * (idx + 1) * (step + 1))))
                F_._notify([p])
            red.finish_step()
            for idx, p in enumerate(params):
                want = float(sum(r + 1 for r in range(world)) * (idx + 1) * (step + 1))
                assert torch.allclose(p.grad, torch.full_like(p, want)), (rank, idx, step)
            assert red.launch_order == sorted(red.launch_order), "buckets must go out in backward order"
            assert len(red.launch_order) == len(red.buckets)
            with torch.no_grad():                              # SGD on the mean gradient: replicas must stay identical
                flat.flat.add_(flat.grad, alpha=-0.01 / world)
        gathered = [torch.empty_like(flat.flat) for _ in range(world)]
        dist.all_gather(gathered, flat.flat)
        assert all(torch.equal(gathered[0], g) for g in gathered[1:]), "replicas diverged"
        # data sharding: per-rank seeds give different clips (bench.py uses seed 1234 + rank)
        x = torch.rand(4, generator=torch.Generator().manual_seed(1234 + rank))
        xs = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(xs, x)
        assert not torch.equal(xs[0], xs[1])
        # batch sharding (test_training_loop.py:221-233, test_distributed.py): row block d on rank d, shards are
        # disjoint, and together they are the global batch (global batch = local x processes)
        from video_vae_b200.ddp import shard_batch
        glob = torch.arange(6 * 3, dtype=torch.float32).reshape(6, 3)
        mine = shard_batch(glob)
        assert mine.shape == (6 // world, 3) and torch.equal(mine, glob[rank * 3:(rank + 1) * 3])
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine.contiguous())
        assert torch.equal(torch.cat(parts), glob)
        try:
            shard_batch(glob[:5])
            raise AssertionError("an indivisible global batch must be rejected")
        except ValueError:
            pass
        red.close()
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_flat_params_and_bucketed_allreduce_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_flat_params_route_collectives_through_a_comm_object():
    """``FlatParams.all_reduce_grads(comm=)`` / ``.broadcast(comm=)`` (the NativeComm route: vvae_comm_* instead of
    torch.distributed) hand the flat buffers to the communicator, and do nothing at world size 1."""
    sys.path.insert(0, ROOT)
    from video_vae_b200.ddp import FlatParams

    class Comm:
        def __init__(self, world):
            self.world, self.calls = world, []

        def all_reduce(self, t, average=False):
            self.calls.append(("all_reduce", t.data_ptr(), t.numel(), average))
            t.mul_(self.world)                      # what a sum over identical ranks does

        def broadcast(self, t, src=0):
            self.calls.append(("broadcast", t.data_ptr(), t.numel(), src))

    m = _Toy()
    flat = FlatParams(m)
    flat.grad.fill_(1.5)
    one = Comm(1)
    flat.all_reduce_grads(comm=one)
    flat.broadcast(comm=one)
    assert one.calls == [] and bool((flat.grad == 1.5).all())
    two = Comm(2)
    flat.all_reduce_grads(comm=two)
    flat.broadcast(src=1, comm=two)
    assert two.calls == [("all_reduce", flat.grad.data_ptr(), flat.total, False), ("broadcast", flat.flat.data_ptr(), flat.total, 1)]
    assert bool((flat.grad == 3.0).all())
    assert all(p.grad.data_ptr() >= flat.grad.data_ptr() for p in m.parameters())      # the views still alias the flat buffer
