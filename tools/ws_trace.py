"""Timeline of one edge-kernel CTA (build the library with `make TRACE=1`):  python tools/ws_trace.py [bf16|bf16x3]
Runs ONE denoiser step (first edge-kernel launch fills the trace) and prints events sorted by time."""
import ctypes as C
import sys
from _common import setup
from keypoint_diffusion_b200 import _lib

model, g, sampler, run, arch = setup("gvp_20kp", sys.argv[1] if len(sys.argv) > 1 else "bf16x3")
run(1)
fn = _lib.lib.kpd_debug_ws_trace
fn.restype = C.c_int
buf = (C.c_ulonglong * (3 * 2048))()
n = C.c_int32()
fn(buf, 2048, C.byref(n))
ev = sorted(((buf[3 * i + 2], buf[3 * i + 1], buf[3 * i]) for i in range(n.value) if buf[3 * i + 2] > 0))
# keep the first launch only (later launches -- other edge types, node kernels -- start much later)
cut = next((k for k in range(1, len(ev)) if ev[k][0] - ev[k - 1][0] > 200000), len(ev))
ev = ev[:cut]
names = {1: "MMA: feats_ready(0) seen", 2: "MMA: main k-steps issued", 3: "MMA: tail_ready seen", 4: "MMA: acc_done committed",
         5: "MMA: feats_ready(g+1) seen", 6: "MMA: gates issued (first halves)", 7: "MMA: gates issued (second halves)", 8: "MMA: gates issued (all)", 10: "SIMT: GVP start", 11: "SIMT: |Vh| published",
         12: "SIMT: Vu done, wait acc", 13: "SIMT: acc_done seen", 14: "SIMT: epi1 published", 15: "SIMT: gates_done seen",
         16: "SIMT: GVP end"}
t0 = ev[0][0] if ev else 0
for t, w, tag in ev:
    nm = names.get(tag, f"MMA: slab {tag - 200} arrived, k-step issued" if tag >= 200 else
                   (f"PROD: slab {tag - 100} issued" if tag >= 100 else str(tag)))
    print(f"{t - t0:8d}  warp {w:2d}  {nm}")
