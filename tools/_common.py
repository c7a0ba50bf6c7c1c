"""Shared set-up of the profiling tools: the bench workload as (model, device graph, exact-layout sampler)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def setup(workload="gvp_20kp", precision="bf16x3", ligands=None, dense_ll=True, use_cuda_graph=False):
    from keypoint_diffusion_b200 import HeteroBatch, synthetic
    wl = dict(bench.WORKLOADS[workload])
    B = int(ligands) if ligands else wl["ligands"]
    dev = torch.device("cuda:0")
    cfg = bench.load_config(wl["cfg"], dense_ll=dense_ll)
    arch = cfg["diffusion"].get("architecture", "egnn")
    model = bench.build_model(cfg, dev)
    if precision != "fp32":
        model.dynamics.set_precision(precision)
    pocket = bench.make_pocket(synthetic, wl, 0, cfg, arch)
    g = HeteroBatch.from_pockets([pocket], [wl["atoms"]] * B, 10).to(dev)
    sampler = model._sampler(g, 50, use_cuda_graph)
    kp = g.nodes["kp"].data
    run = lambda n_steps, seed=1: sampler.run(kp["x_0"], kp["h_0"], kp.get("v_0"), torch.zeros(B, 3, device=dev), seed=seed,
                                              n_steps=n_steps)
    return model, g, sampler, run, arch
