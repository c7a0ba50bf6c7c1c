"""Phase timers of the warp-specialised EGNN edge kernel over a few reverse steps:  python tools/eg_phase_times.py"""
import ctypes as C
from _common import setup
from keypoint_diffusion_b200 import _lib

model, g, sampler, run, arch = setup("egnn_20kp", "bf16x3")
fn = _lib.lib.kpd_debug_eg_times
fn.restype = C.c_int
buf = (C.c_ulonglong * 16)()
run(100)
fn(buf)
run(200)
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
