"""N > 1 host logic on CPU: pair sharding + the SE(3) all-gather over gloo at world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_pairs, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from importlib import import_module
    sharding = import_module("dreg-nerf_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard_pairs(n_pairs, rank, world)
    # a pose that encodes the pair id, standing in for NeRFRegTr's output
    local = torch.stack([torch.full((3, 4), float(i)) for i in mine]) if mine else torch.zeros((0, 3, 4))
    allp = sharding.gather_poses(local, n_pairs)
    torch.save(allp, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [5, 8])
def test_pair_sharding_and_pose_allgather(tmp_path, n_pairs):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_pairs, str(tmp_path)), nprocs=world, join=True)
    want = torch.stack([torch.full((3, 4), float(i)) for i in range(n_pairs)])
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r))
        assert torch.equal(got, want)


def test_shard_partition_properties():
    from importlib import import_module
    sharding = import_module("dreg-nerf_b200.sharding")
    for n, w in [(0, 4), (1, 8), (7, 2), (256, 8)]:
        parts = [sharding.shard_pairs(n, r, w) for r in range(w)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        sharding.shard_pairs(4, 2, 2)


def _grad_worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from importlib import import_module
    sharding = import_module("dreg-nerf_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2))]
    params[0].grad = torch.full((5, 3), float(rank + 1))
    params[1].grad = torch.arange(7.0) * (rank + 1)
    sharding.allreduce_gradients(params)                 # params[2] has no gradient: skipped
    torch.save([p.grad for p in params], os.path.join(out_dir, "g%d.pt" % rank))
    dist.destroy_process_group()


def test_gradient_allreduce_averages_over_ranks(tmp_path):
    world = 2
    mp.spawn(_grad_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        g = torch.load(os.path.join(str(tmp_path), "g%d.pt" % r))
        assert torch.equal(g[0], torch.full((5, 3), 1.5)) and torch.equal(g[1], torch.arange(7.0) * 1.5) and g[2] is None
