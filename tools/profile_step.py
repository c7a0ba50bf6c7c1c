"""A short, non-graph run of the hot path for ncu: a few reverse steps of the bench workload.
    python tools/profile_step.py [workload] [n_steps] [ligands]"""
import sys

import torch

from _common import setup

wl = sys.argv[1] if len(sys.argv) > 1 else "gvp_20kp"
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
model, g, sampler, run, arch = setup(wl, sys.argv[4] if len(sys.argv) > 4 else "fp32", sys.argv[3] if len(sys.argv) > 3 else None)
x, h, _ = run(n_steps)
torch.cuda.synchronize()
print("ok", wl, n_steps, "steps; launches/step", sampler.launches_per_step, "finite", bool(torch.isfinite(x).all()))
