"""Stand-alone timing of the tcgen05 node GEMM (csrc/tc_gemm.cu) at the EGNN's shapes (egnn_20kp bench workload:
2000 ligand atoms / 2000 keypoints per launch problem):  python tools/tc_linear_bench.py [M]
CUDA events over 100 graph-captured launches per shape; weights and activations L2-resident as in the captured loop."""
import sys

import torch

from _common import ROOT  # noqa: F401  (puts the repo on sys.path)
import ctypes as C

from keypoint_diffusion_b200 import _lib, ops, pack

M = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
shapes = [("first-layer products", 257, 2080, 0, False), ("node_mlp.0", 517, 257, 1, False), ("node_mlp.2 + residual", 257, 257, 0, True)]
for name, K, N, act, res in shapes:
    x = torch.randn(M, (K + 3) // 4 * 4, generator=g).to(dev)[:, :K]
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g).to(dev)
    wp = pack.pack_tc_weight(w, True).to(dev)
    r = torch.randn(M, N, generator=g).to(dev) if res else None
    y = ops.tc_linear(x, wp, N, b, r, act, 2)
    ref = x.double().cpu() @ w.double().t() + b.double().cpu()
    if act:
        ref = ref * torch.sigmoid(ref)
    if res:
        ref = ref + r.double().cpu()
    err = (y.double().cpu() - ref).abs().max().item() / ref.abs().max().item()
    # 20 launches captured into one CUDA graph (the Python call costs more than the small shapes run)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ops.tc_linear(x, wp, N, b, r, act, 2)
    torch.cuda.current_stream().wait_stream(side)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(20):
            ops.tc_linear(x, wp, N, b, r, act, 2)
    gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 100
    fl = 2.0 * M * K * N
    tfn = getattr(_lib.lib, "kpd_debug_tcg_times", None) if hasattr(_lib.lib, "kpd_debug_tcg_times") else None
    print(f"{name:24s} M={M} K={K} N={N}: {us:7.1f} us per launch {fl / us / 1e6:6.1f} TFLOP/s algorithmic, rel err {err:.1e}")
    if tfn is not None:      # library built with -DKPD_TCG_TIMERS: where a CTA spends its cycles
        buf = (C.c_ulonglong * 16)()
        tfn(buf)
        gr.replay()
        tfn(buf)
        t = list(buf)
        ct = max(t[5], 1)
        print(f"    cycles per CTA: set-up {t[0] / ct:.0f} | stage A {t[1] / ct:.0f} | wait first accumulator {t[2] / ct:.0f} | "
              f"epilogues {t[3] / ct:.0f}   MMA warp: wait A {t[8] / ct:.0f} | first slab {t[9] / ct:.0f} | issue loop {t[10] / ct:.0f}   ({ct} CTAs with work)")
