"""Print the in-kernel phase timers of the tensor-core GVP kernels for a few reverse steps."""
import ctypes as C
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
from keypoint_diffusion_b200 import HeteroBatch, _lib

dev = torch.device("cuda:0")
cfg_name, kind, n_kp, B, n_atoms = bench.WORKLOADS["gvp_20kp"]
cfg = bench.load_config(cfg_name)
model = bench.build_model(cfg, dev)
model.dynamics.set_precision(sys.argv[1] if len(sys.argv) > 1 else "bf16")
pocket = bench.make_pocket(kind, 0, cfg, "gvp")
g = HeteroBatch.from_pockets([pocket], [n_atoms] * B, 10).to(dev)
sampler = model._sampler(g, 50, False)
kp = g.nodes["kp"].data
fn = _lib.lib.kpd_debug_tc_times
fn.restype = C.c_int
buf = (C.c_ulonglong * 16)()
sampler.run(kp["x_0"], kp["h_0"], kp.get("v_0"), torch.zeros(B, 3, device=dev), seed=1, n_steps=100)
fn(buf)   # reset after warm-up (includes dense early steps)
sampler.run(kp["x_0"], kp["h_0"], kp.get("v_0"), torch.zeros(B, 3, device=dev), seed=1, n_steps=300)
fn(buf)
t = list(buf)
names = ["stage_w", "Vh", "Vu+fence", "featsGEMM", "epi1", "gatesGEMM", "epi2"]
calls = max(t[7], 1)
print("gvp_tile_tc calls", calls, " (all tc kernels)")
tot = max(sum(t[:7]), 1)
for n, v in zip(names, t[:7]):
    print(f"  {n:10s} {v / calls:9.0f} cycles/call  {100 * v / tot:5.1f}%")
print(f"  total      {tot / calls:9.0f} cycles/call = {tot / calls / 1.9e3:.1f} us")
ct = max(t[13], 1)
en = ["setup", "geom+gather", "gvp chain", "seg-reduce", "teardown"]
et = max(sum(t[8:13]), 1)
print("edge CTAs", ct)
for n, v in zip(en, t[8:13]):
    print(f"  {n:12s} {v / ct:9.0f} cycles/CTA  {100 * v / et:5.1f}%")
print(f"  total        {et / ct:9.0f} cycles/CTA = {et / ct / 1.9e3:.1f} us")

fn2 = _lib.lib.kpd_debug_ws_times
fn2.restype = C.c_int
fn2(buf)
t = list(buf)
names = ["Vh+|Vh|", "Vu", "wait acc", "epi1", "wait gates", "epi2"]
calls = max(t[6], 1)
tot = sum(t[:6])
print("ws gvp_simt calls", calls)
for n, v in zip(names, t[:6]):
    print(f"  {n:10s} {v / calls:9.0f} cycles/call  {100 * v / max(tot, 1):5.1f}%")
print(f"  total      {tot / calls:9.0f} cycles/call")
ct = max(t[13], 1)
en = ["setup", "gather", "gvp chain", "seg-reduce"]
et = sum(t[8:12])
print("ws edge CTAs", ct)
for n, v in zip(en, t[8:12]):
    print(f"  {n:12s} {v / ct:9.0f} cycles/CTA  {100 * v / max(et, 1):5.1f}%")
print(f"  total        {et / ct:9.0f} cycles/CTA = {et / ct / 1.9e3:.1f} us")
