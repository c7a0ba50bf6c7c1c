"""Launch timeline of the GVP convs inside ONE replayed CUDA-graph reverse step (needs a library built with
-DKPD_TIMELINE:  make -C keypoint_diffusion_b200/csrc clean; make -C keypoint_diffusion_b200/csrc EXTRA=-DKPD_TIMELINE
OUT=/tmp/libkpd_tl.so;  KPD_LIB=/tmp/libkpd_tl.so python tools/conv_timeline.py [bf16x3|bf16] [step]).
Prints per launch: first CTA start, last CTA end (us, relative to the step's first launch) and working CTAs."""
import ctypes as C
import sys
import torch
from _common import setup
from keypoint_diffusion_b200 import _lib

n_warm = int(sys.argv[2]) if len(sys.argv) > 2 else 200
model, g, _, _, arch = setup("gvp_20kp", sys.argv[1] if len(sys.argv) > 1 else "bf16x3")
sampler = model._sampler(g, 1, True)          # one reverse step per CUDA graph
kp = g.nodes["kp"].data
B = g.batch_size
dev = g.device
buf = (C.c_ulonglong * (3 * 256))()
zeros = torch.zeros(B, 3, device=dev)
# slots keep the start of CTA 0 of the LAST launch and the latest end: run n_warm steps and read the last one
n = _lib.lib.kpd_debug_timeline(buf, 1)
assert n > 0, "library was not built with -DKPD_TIMELINE"
sampler.run(kp["x_0"], kp["h_0"], kp.get("v_0"), zeros, seed=1, n_steps=n_warm)
torch.cuda.synchronize()
_lib.lib.kpd_debug_timeline(buf, 1)
t = list(buf)
rows = []
names = ["E ll", "E kl", "E lk", "E kk", "N lig", "N kp"]
for slot in range(256):
    s0, s1, c = t[3 * slot], t[3 * slot + 1], t[3 * slot + 2]
    if c:
        rows.append((s0, s1, c / n_warm, slot // 8, names[slot % 8] if slot % 8 < 6 else "?"))
t0 = min(r[0] for r in rows)
print(f"# {sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'}: reverse step {n_warm} of a trajectory, launches/step {sampler.launches_per_step}")
print("# conv kernel   start_us   end_us   dur_us   mean working CTAs per launch")
for s0, s1, c, l, nm in sorted(rows):
    print(f"  {l}   {nm:6s} {(s0 - t0) / 1e3:9.1f} {(s1 - t0) / 1e3:8.1f} {(s1 - s0) / 1e3:8.1f} {c:8.1f}")
print(f"# span {(max(r[1] for r in rows) - t0) / 1e3:.1f} us")
