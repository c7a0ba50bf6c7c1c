"""Phase timers of the warp-specialised EGNN edge kernel over a few reverse steps:  python tools/eg_phase_times.py"""
import ctypes as C
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
from keypoint_diffusion_b200 import HeteroBatch, _lib

dev = torch.device("cuda:0")
cfg_name, kind, n_kp, B, n_atoms = bench.WORKLOADS["egnn_20kp"]
cfg = bench.load_config(cfg_name)
model = bench.build_model(cfg, dev)
model.dynamics.set_precision("bf16x3")
pocket = bench.make_pocket(kind, 0, cfg, "egnn")
g = HeteroBatch.from_pockets([pocket], [n_atoms] * B, 10).to(dev)
sampler = model._sampler(g, 50, False)
kp = g.nodes["kp"].data
fn = _lib.lib.kpd_debug_eg_times
fn.restype = C.c_int
buf = (C.c_ulonglong * 16)()
sampler.run(kp["x_0"], kp["h_0"], kp.get("v_0"), torch.zeros(B, 3, device=dev), seed=1, n_steps=100)
fn(buf)
sampler.run(kp["x_0"], kp["h_0"], kp.get("v_0"), torch.zeros(B, 3, device=dev), seed=1, n_steps=200)
fn(buf)
t = list(buf)
names = ["set-up", "indices + geometry", "build A (edge)", "build A (coord)", "wait + epilogue (edge)", "reduce (edge)",
         "wait + epilogue (coord)", "reduce (coord)"]
ct = max(t[8], 1)
tot = max(sum(t[:8]), 1)
print("egnn edge CTAs", ct)
for n, v in zip(names, t[:8]):
    print(f"  {n:26s} {v / ct:9.0f} cycles/CTA  {100 * v / tot:5.1f}%")
print(f"  total                      {tot / ct:9.0f} cycles/CTA = {tot / ct / 1.9e3:.1f} us")
