"""CPU study (oracle only): end-to-end error of the denoisers when every Linear runs as a split low-precision
GEMM with fp32 accumulation, against the plain fp32 oracle.  Decides the operand format of the tensor-core
parity mode.    python tools/split_precision_study.py"""
import sys
from pathlib import Path

import torch
import yaml

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle import flat  # noqa: E402
from helpers import GOLDEN, flat_batch, oracle_cfg, oracle_forward, rel_err  # noqa: E402
from test_gpu_parity import _full_size_case  # noqa: E402


def split(x, dt, n):
    parts, r = [], x
    for _ in range(n):
        p = r.to(dt).float()
        parts.append(p)
        r = r - p
    return parts


def make_lin(dt, na, nw, terms):
    def _lin(sd, name, x):
        w = sd[name + ".weight"].float()
        b = sd.get(name + ".bias")
        xs, ws = split(x.float(), dt, na), split(w, dt, nw)
        y = 0
        for (i, j) in terms:
            if i < na and j < nw:
                y = y + (xs[i].double() @ ws[j].double().t())
        y = y.float()
        return y if b is None else y + b
    return _lin


cfgs = yaml.safe_load(open(GOLDEN / "shipped_configs.yml"))
orig = flat._lin
modes = {
    "bf16 x1": (torch.bfloat16, 1, 1, [(0, 0)]),
    "bf16 x3 (hh,lh,hl)": (torch.bfloat16, 2, 2, [(0, 0), (1, 0), (0, 1)]),
    "bf16 x4": (torch.bfloat16, 2, 2, [(0, 0), (1, 0), (0, 1), (1, 1)]),
    "bf16 x6 (3-way split)": (torch.bfloat16, 3, 3, [(0, 0), (1, 0), (0, 1), (1, 1), (2, 0), (0, 2)]),
    "fp16 x1": (torch.float16, 1, 1, [(0, 0)]),
    "fp16 x3 (hh,lh,hl)": (torch.float16, 2, 2, [(0, 0), (1, 0), (0, 1)]),
    "fp16 A-split only x2": (torch.float16, 2, 1, [(0, 0), (1, 0)]),
}
for arch in ("gvp", "egnn"):
    sd, kw, rec_nf, inputs = _full_size_case(arch, cfgs)
    cfg = oracle_cfg(arch, kw, 10, rec_nf)
    for tval in (0.5,):
        fb = flat_batch(inputs)
        t = torch.full((fb.B,), tval)
        flat._lin = orig
        ref_h, ref_x = oracle_forward(arch, sd, cfg, fb, t)
        fb64 = flat_batch(inputs, torch.float64)
        sd64 = {k: v.double() for k, v in sd.items()}
        h64, x64 = oracle_forward(arch, sd64, cfg, fb64, t.double())
        print(f"{arch} t={tval}: fp32 oracle vs fp64 oracle: {rel_err(ref_h, h64):.2e} {rel_err(ref_x, x64):.2e}")
        for name, (dt, na, nw, terms) in modes.items():
            flat._lin = make_lin(dt, na, nw, terms)
            h, x = oracle_forward(arch, sd, cfg, flat_batch(inputs), t)
            print(f"  {name:24s} eps_h {rel_err(h, ref_h):.2e}  eps_x {rel_err(x, ref_x):.2e}")
flat._lin = orig
