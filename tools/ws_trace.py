"""Timeline of one edge-kernel CTA (build the library with `make TRACE=1`):  python tools/ws_trace.py [bf16|bf16x3]
Runs ONE denoiser step (first edge-kernel launch fills the trace) and prints events sorted by time."""
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
model.dynamics.set_precision(sys.argv[1] if len(sys.argv) > 1 else "bf16x3")
pocket = bench.make_pocket(kind, 0, cfg, "gvp")
g = HeteroBatch.from_pockets([pocket], [n_atoms] * B, 10).to(dev)
sampler = model._sampler(g, 50, False)
kp = g.nodes["kp"].data
sampler.run(kp["x_0"], kp["h_0"], kp.get("v_0"), torch.zeros(B, 3, device=dev), seed=1, n_steps=1)
fn = _lib.lib.kpd_debug_ws_trace
fn.restype = C.c_int
buf = (C.c_ulonglong * (3 * 2048))()
n = C.c_int32()
fn(buf, 2048, C.byref(n))
ev = sorted(((buf[3 * i + 2], buf[3 * i + 1], buf[3 * i]) for i in range(n.value)))
names = {1: "MMA: feats_ready(0) seen", 2: "MMA: main k-steps issued", 3: "MMA: tail_ready seen", 4: "MMA: acc_done committed",
         5: "MMA: feats_ready(g+1) seen", 6: "MMA: gates committed", 10: "SIMT: GVP start", 11: "SIMT: |Vh| published",
         12: "SIMT: Vu done, wait acc", 13: "SIMT: acc_done seen", 14: "SIMT: epi1 published", 15: "SIMT: gates_done seen",
         16: "SIMT: GVP end"}
t0 = ev[0][0] if ev else 0
for t, w, tag in ev:
    nm = names.get(tag, f"MMA: slab {tag - 200} arrived, k-step issued" if tag >= 200 else
                   (f"PROD: slab {tag - 100} issued" if tag >= 100 else str(tag)))
    print(f"{t - t0:8d}  warp {w:2d}  {nm}")
