"""Multi-GPU sharding of sampling jobs.

Every (pocket, ligand) complex is an independent sample, so the path shards with no per-step
communication: complexes are dealt to ranks by estimated cost, each rank runs its own captured
loop, and the only collective is one final gather of coordinates + atom features
(torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests).  The reference's
equivalent is a slurm array of independent processes (gen_test_commands.py:36-40).
"""
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_complexes(n_lig_atoms: Sequence[int], n_kp: Sequence[int], world_size: int) -> List[List[int]]:
    """Indices of the complexes each rank samples: sort by estimated edge count (descending) and deal
    round-robin in snake order, so every rank gets the same number of complexes (+-1) and a similar
    number of edges."""
    cost = [nl * (nl - 1) + 2 * nk * min(nl, 7) for nl, nk in zip(n_lig_atoms, n_kp)]
    order = sorted(range(len(cost)), key=lambda i: (-cost[i], i))
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for pos, idx in enumerate(order):
        rnd, r = divmod(pos, world_size)
        shards[r if rnd % 2 == 0 else world_size - 1 - r].append(idx)
    return [sorted(s) for s in shards]


def gather_ligands(x_lig: torch.Tensor, h_lig: torch.Tensor, sizes: Sequence[int], group=None
                   ) -> Tuple[List[torch.Tensor], List[torch.Tensor], List[List[int]]]:
    """All-gather every rank's sampled ligands.  x_lig [n,3], h_lig [n,F] and the per-complex atom counts
    of this rank -> per-rank lists (positions, features, sizes) on every rank.  One padded all_gather."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [x_lig], [h_lig], [list(sizes)]
    ws = dist.get_world_size(group)
    all_sizes: List[List[int]] = [None] * ws
    dist.all_gather_object(all_sizes, list(sizes), group=group)
    n_max = max(sum(s) for s in all_sizes)
    F = h_lig.shape[1]
    buf = torch.zeros(n_max, 3 + F, dtype=torch.float32, device=x_lig.device)
    buf[: x_lig.shape[0], :3] = x_lig
    buf[: x_lig.shape[0], 3:] = h_lig
    out = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(out, buf, group=group)
    xs = [o[: sum(s), :3].contiguous() for o, s in zip(out, all_sizes)]
    hs = [o[: sum(s), 3:].contiguous() for o, s in zip(out, all_sizes)]
    return xs, hs, all_sizes
