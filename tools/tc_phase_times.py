"""Print the in-kernel phase timers of the warp-specialised tensor-core GVP kernels for a few reverse steps.
    python tools/tc_phase_times.py [bf16|bf16x3]"""
import ctypes as C
import sys
from _common import setup
from keypoint_diffusion_b200 import _lib

model, g, sampler, run, arch = setup("gvp_20kp", sys.argv[1] if len(sys.argv) > 1 else "bf16")
fn = _lib.lib.kpd_debug_ws_times
fn.restype = C.c_int
buf = (C.c_ulonglong * 64)()
run(100)
fn(buf)   # reset after warm-up
run(300)
fn(buf)
t = list(buf)
names = ["Vh+|Vh|", "Vu", "wait acc", "epi1", "wait gates", "epi2"]
for kname, b in (("edge", 0), ("node", 16), ("head", 32)):
    calls = max(t[b + 6], 1)
    tot = max(sum(t[b:b + 6]), 1)
    print(f"{kname} kernel: gvp_simt calls {calls}")
    for n, v in zip(names, t[b:b + 6]):
        print(f"  {n:10s} {v / calls:9.0f} cycles/call  {100 * v / tot:5.1f}%")
    print(f"  total      {tot / calls:9.0f} cycles/call   (of Vh+|Vh|: {t[b + 7] / calls:.0f} cycles waiting for the small weights)")
ct_e = max(t[13], 1)
print(f"edge gather breakdown: index staging {t[12] / ct_e:.0f} | gathers issued, geometry + vectors loaded, segments built "
      f"{t[14] / ct_e:.0f} | rbf + control join {t[15] / ct_e:.0f} | waiting for the gathered rows {(t[9] - t[12] - t[14] - t[15]) / ct_e:.0f} cycles/CTA")
print(f"   of the second: cp.async issue {t[40] / ct_e:.0f} | geometry + vector loads {t[41] / ct_e:.0f} | zero fill {t[42] / ct_e:.0f} | "
      f"segment table {t[43] / ct_e:.0f}")
for kname, b, en in (("edge", 8, ["setup", "gather", "gvp chain", "seg-reduce"]),
                     ("node", 24, ["setup", "phase 1a (scalars)", "phase 1b (vectors)", "gvp chain", "phase 3"])):
    ct = max(t[b + 5], 1)
    et = max(sum(t[b:b + len(en)]), 1)
    print(f"{kname} CTAs {ct}")
    for n, v in zip(en, t[b:b + len(en)]):
        print(f"  {n:20s} {v / ct:9.0f} cycles/CTA  {100 * v / et:5.1f}%")
    print(f"  total                {et / ct:9.0f} cycles/CTA = {et / ct / 1.9e3:.1f} us")
