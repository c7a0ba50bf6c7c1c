"""Shared test helpers: golden-fixture loading, oracle configs, error metrics."""
from pathlib import Path

import torch

from oracle import flat

GOLDEN = Path(__file__).resolve().parent / "golden"

DENOISER_FIXTURES = ["egnn_small_kp", "egnn_small_nokp", "gvp_small_sum", "gvp_small_mean", "gvp_small_zero"]


def load_golden(name):
    return torch.load(GOLDEN / f"{name}.pt", weights_only=False)


def oracle_cfg(kind, kwargs, atom_nf, rec_nf):
    kw = dict(kwargs)
    kw.pop("n_keypoints", None)
    kw.pop("dropout", None)
    kw.pop("no_cg", None)
    if kind == "egnn":
        return flat.EGNNConfig(atom_nf=atom_nf, rec_nf=rec_nf, **kw)
    return flat.GVPConfig(n_lig_scalars=atom_nf, n_kp_scalars=rec_nf, **kw)


def flat_batch(inputs, dtype=torch.float32):
    def f(k):
        v = inputs.get(k)
        return None if v is None else v.to(dtype)
    return flat.FlatBatch(lig_n=inputs["lig_n"], kp_n=inputs["kp_n"], kp_x=f("kp_x"), kp_h=f("kp_h"),
                          kk_src=inputs["kk_src"], kk_dst=inputs["kk_dst"], kp_v=f("kp_v"),
                          lig_x=f("lig_x"), lig_h=f("lig_h"))


def oracle_forward(kind, sd, cfg, batch, t, **kw):
    fn = flat.egnn_forward if kind == "egnn" else flat.gvp_forward
    return fn(sd, cfg, batch, t, **kw)


def rel_err(a, b):
    """max |a-b| / max |b|  -- the 'relative' of the north_star's 1e-4 bar, taken against the
    tensor's own scale (element-wise relative error is meaningless near zero crossings)."""
    a = a.double()
    b = b.double()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def edge_set(ei):
    """A sorted list of (src, dst) pairs for exact set comparison."""
    return sorted(zip(ei[0].tolist(), ei[1].tolist()))


def err_report(a, b):
    """Three views of the error of `a` against the reference `b` ([N, C] tensors):
      norm  max|a-b| / max|b|                       -- the headline bar (north_star's 1e-4 "relative")
      rms   rms(a-b) / rms(b)                       -- the typical element
      chan  max over channels c of max_i|a-b|[:,c] / max_i|b|[:,c]
                                                    -- a small-magnitude channel cannot hide behind a large one
    """
    a = a.double().reshape(b.shape[0], -1)
    b = b.double().reshape(b.shape[0], -1)
    d = (a - b).abs()
    norm = float(d.max() / b.abs().max().clamp(min=1e-30))
    rms = float(d.pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp(min=1e-30))
    chan = float((d.max(0).values / b.abs().max(0).values.clamp(min=1e-30)).max())
    return {"norm": norm, "rms": rms, "chan": chan}


def fmt_err(r):
    return f"norm {r['norm']:.2e} rms {r['rms']:.2e} per-channel {r['chan']:.2e}"
