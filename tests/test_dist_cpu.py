"""CPU: the N>1 path (sharding + the final gather) with gloo, world_size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, sizes_by_rank):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from keypoint_diffusion_b200 import dist as kdist
    sizes = sizes_by_rank[rank]
    n = sum(sizes)
    x = torch.full((n, 3), float(rank + 1))
    h = torch.arange(n * 10, dtype=torch.float32).reshape(n, 10) + 1000 * rank
    xs, hs, all_sizes = kdist.gather_ligands(x, h, sizes)
    assert all_sizes == sizes_by_rank
    for r in range(world):
        nr = sum(sizes_by_rank[r])
        assert xs[r].shape == (nr, 3) and hs[r].shape == (nr, 10)
        assert float(xs[r].min()) == float(xs[r].max()) == float(r + 1)
        assert float(hs[r][0, 0]) == 1000.0 * r
    dist.destroy_process_group()


def test_gather_ligands_world2():
    sizes = [[20, 8, 35], [13, 27]]
    mp.spawn(_worker, args=(2, _free_port(), sizes), nprocs=2, join=True)


def test_shard_complexes_balanced():
    from keypoint_diffusion_b200.dist import shard_complexes
    n_lig = [20] * 50 + [8] * 25 + [35] * 25 + [60] * 4
    n_kp = [20] * len(n_lig)
    for w in (1, 2, 4, 8):
        shards = shard_complexes(n_lig, n_kp, w)
        assert sorted(i for s in shards for i in s) == list(range(len(n_lig)))
        counts = [len(s) for s in shards]
        assert max(counts) - min(counts) <= 1
        cost = [sum(n_lig[i] * (n_lig[i] - 1) for i in s) for s in shards]
        assert max(cost) <= 1.25 * (sum(cost) / w) + 60 * 59
