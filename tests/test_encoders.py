"""Learned receptor encoders against fixtures produced by the reference's own
models/receptor_encoder.py / receptor_encoder_gvp.py (tests/golden/make_golden_encoders.py): keypoint positions and
features within 1e-4 relative, rk / kk edge sets exact.  Same checks on CPU and (gpu marker) on cuda:0."""
from pathlib import Path

import pytest
import torch

from helpers import rel_err

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = ["enc_egnn_knn", "enc_egnn_rad", "enc_gvp_knn", "enc_gvp_mean_rad", "enc_gvp_zero"]
TOL = 1e-4


def _raw_batch(fx, device):
    from keypoint_diffusion_b200 import hetero
    cut = fx["kwargs"]["graph_cutoffs"]
    gs = [hetero.build_initial_complex_graph(p["x"], p["h"], p["res"], fx["kwargs"]["n_keypoints"], cut) for p in fx["pockets"]]
    return hetero.batch(gs).to(device)


def _edge_set(e):
    return sorted(zip(e[0].tolist(), e[1].tolist()))


def _check(name, device):
    from keypoint_diffusion_b200.receptor_encoder import ReceptorEncoder, ReceptorEncoderGVP
    from keypoint_diffusion_b200.utils import get_batch_idxs
    fx = torch.load(GOLDEN / f"{name}.pt")
    enc = (ReceptorEncoder if fx["kind"] == "egnn" else ReceptorEncoderGVP)(**fx["kwargs"]).eval()
    enc.load_state_dict(fx["state_dict"], strict=True)
    enc = enc.to(device)
    g = _raw_batch(fx, device)
    # the raw graph builder reproduces the reference's rr graph (pdbbind_processing.py:246-250)
    assert torch.equal(torch.stack(g.edges(form="uv", etype="rr")).cpu(), fx["rr"])
    assert torch.equal(g.batch_num_edges("rr").cpu(), fx["rr_n"])
    with torch.no_grad():
        out = enc(g, get_batch_idxs(g))
    kp = out.nodes["kp"].data
    assert rel_err(kp["x_0"].cpu(), fx["kp_x"]) < TOL
    assert rel_err(kp["h_0"].cpu(), fx["kp_h"]) < TOL
    if fx["kp_v"] is not None:
        assert rel_err(kp["v_0"].cpu(), fx["kp_v"]) < TOL
    for et in ("kk", "rk"):
        s, d = out.edges(form="uv", etype=et)
        assert _edge_set((s.cpu(), d.cpu())) == _edge_set(fx[et]), et
        assert torch.equal(out.batch_num_edges(et).cpu(), fx[et + "_n"]), et
    # kk edges arrive grouped by destination in the order the reference adds them
    assert torch.equal(torch.stack(out.edges(form="uv", etype="kk")).cpu(), fx["kk"])
    return out


@pytest.mark.parametrize("name", CASES)
def test_encoder_matches_reference_cpu(name):
    _check(name, "cpu")


def test_encoded_pockets_feed_the_sampler_layout():
    """encode_receptors output has what sample_from_encoded_receptors reads: kp data, kk edges, per-complex counts."""
    out = _check("enc_gvp_knn", "cpu")
    n_kp = out.batch_num_nodes("kp")
    assert out.nodes["kp"].data["v_0"].shape == (int(n_kp.sum()), 4, 3)
    assert int(out.batch_num_edges("kk").sum()) == out.num_edges("kk")
    s, d = out.edges(form="uv", etype="kk")
    off = torch.cumsum(n_kp, 0) - n_kp
    b = torch.bucketize(d, torch.cumsum(n_kp, 0), right=True)
    assert bool(((s >= off[b]) & (s < off[b] + n_kp[b])).all())          # no edge crosses complexes


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_encoder_matches_reference_gpu(name):
    _check(name, "cuda:0")
