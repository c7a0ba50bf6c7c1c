"""Synthetic already-encoded pockets (no dataset or checkpoint is available offline).

Shapes follow SURVEY.md section 8(d): seed = 1234 + pocket_id; keypoint models: n_k
positions uniform in a 4-10 A shell, kp.h_0 ~ N(0,1) of the encoder's output width, GVP
kp.v_0 ~ 0.1 N(0,1) [n_k,V,3], kk = radius graph (graph_cutoffs.kk) on the keypoints;
all-atom: ~500 points from a jittered 1.5 A lattice inside a 14 A ball minus a 4 A cavity,
one-hot element features, kk = radius graph r = graph_cutoffs.rr (what
FixedReceptorEncoder copies from rr, reference models/receptor_encoder_fixed.py:41-44);
C-alpha: 42 points >= 3.8 A apart.

Pure setup code on the CPU (runs once per pocket, outside the timed hot path).
"""
import math
from dataclasses import dataclass
from typing import List, Optional

import torch


@dataclass
class EncodedPocket:
    kp_x: torch.Tensor                 # [n_k,3] fp32
    kp_h: torch.Tensor                 # [n_k,C] fp32
    kk_src: torch.Tensor               # int64 [E_kk] (local indices)
    kk_dst: torch.Tensor
    kp_v: Optional[torch.Tensor] = None  # [n_k,V,3] fp32 (GVP)

    @property
    def n_kp(self):
        return int(self.kp_x.shape[0])


def _radius_graph_local(x, r):
    """Setup-time brute-force radius graph (src=neighbour, dst=centre, no self loops),
    grouped by dst with ascending src -- the order torch_cluster.radius_graph emits."""
    d = x[None, :, :] - x[:, None, :]
    d2 = (d * d).sum(-1)
    hit = d2 < float(r) * float(r)
    hit.fill_diagonal_(False)
    dst, src = torch.nonzero(hit, as_tuple=True)
    return src.contiguous(), dst.contiguous()


def keypoint_pocket(pocket_id: int, n_kp: int = 20, feat_dim: int = 128, vector_size: int = 0,
                    kk_cutoff: float = 8.0) -> EncodedPocket:
    g = torch.Generator().manual_seed(1234 + pocket_id)
    dirs = torch.randn(n_kp, 3, generator=g)
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    u = torch.rand(n_kp, 1, generator=g)
    rad = (4.0 ** 3 + u * (10.0 ** 3 - 4.0 ** 3)) ** (1.0 / 3.0)
    x = (dirs * rad).float()
    h = torch.randn(n_kp, feat_dim, generator=g)
    v = 0.1 * torch.randn(n_kp, vector_size, 3, generator=g) if vector_size > 0 else None
    s, d = _radius_graph_local(x, kk_cutoff)
    return EncodedPocket(x, h, s, d, v)


def all_atom_pocket(pocket_id: int, n_atoms: int = 500, n_elements: int = 10, vector_size: int = 0,
                    rr_cutoff: float = 3.5) -> EncodedPocket:
    g = torch.Generator().manual_seed(1234 + pocket_id)
    ax = torch.arange(-14.0, 14.01, 1.5)
    grid = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
    grid = grid + 0.3 * torch.randn(grid.shape, generator=g)
    rn = grid.norm(dim=1)
    grid = grid[(rn < 14.0) & (rn > 4.0)]
    # keep the n_atoms closest to the cavity (a pocket is the shell around the ligand site)
    order = torch.argsort(grid.norm(dim=1))
    x = grid[order[:n_atoms]].float().contiguous()
    n = x.shape[0]
    probs = torch.zeros(n_elements)
    probs[:4] = torch.tensor([0.63, 0.17, 0.19, 0.01])
    el = torch.multinomial(probs, n, replacement=True, generator=g)
    h = torch.nn.functional.one_hot(el, n_elements).float()
    v = torch.zeros(n, vector_size, 3) if vector_size > 0 else None  # fixed encoder: v_0 = 0
    s, d = _radius_graph_local(x, rr_cutoff)
    return EncodedPocket(x, h, s, d, v)


def ca_pocket(pocket_id: int, n_res: int = 42, feat_dim: int = 10, vector_size: int = 0,
              rr_cutoff: float = 3.5) -> EncodedPocket:
    g = torch.Generator().manual_seed(1234 + pocket_id)
    pts: List[torch.Tensor] = []
    while len(pts) < n_res:
        dirs = torch.randn(3, generator=g)
        dirs = dirs / dirs.norm()
        rad = (5.0 ** 3 + torch.rand(1, generator=g) * (12.0 ** 3 - 5.0 ** 3)) ** (1.0 / 3.0)
        p = dirs * rad
        if all((p - q).norm() >= 3.8 for q in pts):
            pts.append(p)
    x = torch.stack(pts).float()
    el = torch.randint(0, feat_dim, (n_res,), generator=g)
    h = torch.nn.functional.one_hot(el, feat_dim).float()
    v = torch.zeros(n_res, vector_size, 3) if vector_size > 0 else None
    s, d = _radius_graph_local(x, rr_cutoff)
    return EncodedPocket(x, h, s, d, v)


def ligand_noise_state(n_lig_atoms: List[int], atom_nf: int, seed: int):
    """Seeded stand-in for a mid-trajectory ligand state (teacher-forced parity inputs)."""
    g = torch.Generator().manual_seed(seed)
    N = int(sum(n_lig_atoms))
    x = 2.0 * torch.randn(N, 3, generator=g)
    h = torch.randn(N, atom_nf, generator=g)
    return x, h


def raw_pocket(pocket_id: int, n_atoms: int = 336, n_elements: int = 10, atoms_per_residue: int = 8):
    """An un-encoded synthetic pocket for the learned receptor encoders: (positions [n,3], one-hot elements [n,F],
    residue index [n]); the all-atom lattice above with consecutive atoms grouped into residues."""
    pk = all_atom_pocket(pocket_id, n_atoms, n_elements)
    return pk.kp_x, pk.kp_h, torch.arange(pk.n_kp) // atoms_per_residue
