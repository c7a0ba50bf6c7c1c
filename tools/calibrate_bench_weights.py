"""Find a scale for the coordinate-output layer of the seeded random model such that the sampled
ligands stay compact over the 1000 steps (an untrained denoiser does not cancel the 1/alpha growth
of the posterior mean, so ligands inflate to hundreds of Angstrom and the ll graph empties, which
would under-load the benchmark).  Prints mean edges per step and the final ligand radius."""
import ctypes as C
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from keypoint_diffusion_b200 import HeteroBatch, _lib  # noqa: E402

dev = torch.device("cuda:0")
for wl, scales in (("gvp_20kp", [1, -1, 30, -30, 100, -100, 300, -300]), ("egnn_20kp", [1, -1, 30, -30, 100, -100, 1000, -1000])):
    cfg_name, kind, n_kp, B, n_atoms = bench.WORKLOADS[wl]
    B = 32
    cfg = bench.load_config(cfg_name)
    arch = cfg["diffusion"].get("architecture", "egnn")
    pocket = bench.make_pocket(kind, 0, cfg, arch)
    for sc in scales:
        model = bench.build_model(cfg, dev)
        with torch.no_grad():
            for name, p in model.named_parameters():
                if arch == "gvp" and name.endswith(f"noise_predictor.gvps.{cfg['dynamics_gvp']['n_noise_gvps'] - 1}.Wu"):
                    p.mul_(sc)
                if arch == "egnn" and ".coord_mlp." in name and name.endswith(".4.weight"):
                    p.mul_(sc)
        g = HeteroBatch.from_pockets([pocket], [n_atoms] * B, 10).to(dev)
        x, h = model.sample_from_encoded_receptors(g, init_lig_pos=torch.zeros(B, 3, device=dev), seed=7,
                                                   return_device_tensors=True, sub_batches=1)
        torch.cuda.synchronize()
        s = next(iter(model._samplers.values()))
        st = (C.c_double * 4)()
        _lib.check(_lib.lib.kpd_sampler_edge_stats(s.handle, st))
        xr = x.view(B, n_atoms, 3)
        rad = (xr - xr.mean(1, keepdim=True)).norm(dim=-1)
        print(f"{wl} scale {sc:>6}: mean ll/complex {st[0] / B:7.1f}  kl/complex {st[1] / B:6.1f}  final radius "
              f"median {float(rad.median()):9.2f} max {float(rad.max()):10.2f}  finite={bool(torch.isfinite(x).all())}",
              flush=True)
