"""A short, non-graph run of the hot path for ncu: a few reverse steps of the bench workload.
    python tools/profile_step.py [workload] [n_steps] [ligands]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from keypoint_diffusion_b200 import HeteroBatch  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "gvp_20kp"
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg_name, kind, n_kp, B, n_atoms = bench.WORKLOADS[wl]
if len(sys.argv) > 3:
    B = int(sys.argv[3])
dev = torch.device("cuda:0")
cfg = bench.load_config(cfg_name)
arch = cfg["diffusion"].get("architecture", "egnn")
model = bench.build_model(cfg, dev)
if len(sys.argv) > 4 and sys.argv[4] != "fp32":
    model.dynamics.set_precision(sys.argv[4])
pocket = bench.make_pocket(kind, 0, cfg, arch)
g = HeteroBatch.from_pockets([pocket], [n_atoms] * B, 10).to(dev)
sampler = model._sampler(g, 50, False)
kp = g.nodes["kp"].data
x, h, _ = sampler.run(kp["x_0"], kp["h_0"], kp.get("v_0"), torch.zeros(B, 3, device=dev), seed=1, n_steps=n_steps)
torch.cuda.synchronize()
print("ok", wl, n_steps, "steps; launches/step", sampler.launches_per_step, "finite", bool(torch.isfinite(x).all()))
