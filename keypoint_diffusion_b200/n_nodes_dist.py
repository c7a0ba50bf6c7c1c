"""Ligand-size prior (reference models/n_nodes_dist.py:8-60): joint histogram over
(receptor nodes, ligand atoms) -> multinomial ligand sizes conditioned on the pocket size."""
import pickle
from pathlib import Path

import torch


class LigandSizeDistribution:
    def __init__(self, processed_dataset_dir: Path):
        joint_dist_file = Path(processed_dataset_dir) / "train_n_node_joint_dist.pkl"
        if not joint_dist_file.exists():
            raise ValueError(f"Joint distribution file {joint_dist_file} does not exist")
        with open(joint_dist_file, "rb") as f:
            joint_histogram, rec_bounds, lig_bounds = pickle.load(f)
        self.joint_histogram = torch.from_numpy(joint_histogram)
        self.rec_bounds = (int(rec_bounds[0]), int(rec_bounds[1]))
        self.lig_bounds = (int(lig_bounds[0]), int(lig_bounds[1]))
        self.rec_idx_to_size = torch.arange(self.rec_bounds[0], self.rec_bounds[1] + 1)
        self.lig_idx_to_size = torch.arange(self.lig_bounds[0], self.lig_bounds[1] + 1)
        self.rec_size_to_idx = {int(s): i for i, s in enumerate(self.rec_idx_to_size)}

    def sample(self, n_nodes_rec: torch.Tensor, n_replicates: int) -> torch.Tensor:
        """[len(n_nodes_rec), n_replicates] ligand sizes; receptor sizes outside the training
        range are clamped to it (reference :44-55 prints a warning and does the same)."""
        sizes = n_nodes_rec.clone().long().clamp(self.rec_bounds[0], self.rec_bounds[1])
        rec_idxs = torch.tensor([self.rec_size_to_idx[int(s)] for s in sizes])
        lig_idxs = torch.multinomial(self.joint_histogram[rec_idxs], n_replicates, replacement=True)
        return self.lig_idx_to_size[lig_idxs]
