"""Print the in-kernel phase timers of the warp-specialised tensor-core GVP kernels for a few reverse steps.
    python tools/tc_phase_times.py [bf16|bf16x3]"""
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
fn = _lib.lib.kpd_debug_ws_times
fn.restype = C.c_int
buf = (C.c_ulonglong * 64)()
sampler.run(kp["x_0"], kp["h_0"], kp.get("v_0"), torch.zeros(B, 3, device=dev), seed=1, n_steps=100)
fn(buf)   # reset after warm-up (includes dense early steps)
sampler.run(kp["x_0"], kp["h_0"], kp.get("v_0"), torch.zeros(B, 3, device=dev), seed=1, n_steps=300)
fn(buf)
t = list(buf)
names = ["Vh+|Vh|", "Vu", "wait acc", "epi1", "wait gates", "epi2"]
for kname, b in (("edge", 0), ("node", 16), ("head", 32)):
    calls = max(t[b + 6], 1)
    tot = max(sum(t[b:b + 6]), 1)
    print(f"{kname} kernel: gvp_simt calls {calls}")
    for n, v in zip(names, t[b:b + 6]):
        print(f"  {n:10s} {v / calls:9.0f} cycles/call  {100 * v / tot:5.1f}%")
    print(f"  total      {tot / calls:9.0f} cycles/call")
for kname, b, en in (("edge", 8, ["setup", "gather", "gvp chain", "seg-reduce"]),
                     ("node", 24, ["setup", "phase 1a (scalars)", "phase 1b (vectors)", "gvp chain", "phase 3"])):
    ct = max(t[b + 5], 1)
    et = max(sum(t[b:b + len(en)]), 1)
    print(f"{kname} CTAs {ct}")
    for n, v in zip(en, t[b:b + len(en)]):
        print(f"  {n:20s} {v / ct:9.0f} cycles/CTA  {100 * v / et:5.1f}%")
    print(f"  total                {et / ct:9.0f} cycles/CTA = {et / ct / 1.9e3:.1f} us")
