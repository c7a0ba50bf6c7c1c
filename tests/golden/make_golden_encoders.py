"""Golden fixtures for the learned receptor encoders: tests/golden/enc_*.pt.

Runs the REFERENCE's own models/receptor_encoder.py and models/receptor_encoder_gvp.py (imported read-only from
/root/reference over the DGL / torch_cluster / torch_scatter stand-ins of oracle/ref_shim) on small seeded raw pocket
graphs and stores inputs, weights and outputs.  Build container only; the fixtures travel to the GPU box.

    python tests/golden/make_golden_encoders.py
"""
import importlib
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle.ref_shim import loader  # noqa: E402
from oracle import graph as G  # noqa: E402
from keypoint_diffusion_b200 import synthetic  # noqa: E402

OUT = Path(__file__).resolve().parent
CUT = {"kk": 2.5, "kl": 8, "ll": 5, "rk": 100, "rr": 3.5}

CASES = {
    # name: (kind, ctor kwargs, atoms per pocket)
    "enc_egnn_knn": ("egnn", dict(n_convs=3, n_keypoints=6, graph_cutoffs=CUT, in_n_node_feat=10, use_sameres_feat=True,
                                  hidden_n_node_feat=32, out_n_node_feat=24, use_tanh=True, coords_range=10, message_norm=0.0,
                                  kp_rad=0.0, k_closest=4, norm=True, fix_pos=False), [40, 57, 33]),
    "enc_egnn_rad": ("egnn", dict(n_convs=2, n_keypoints=5, graph_cutoffs=CUT, in_n_node_feat=10, use_sameres_feat=False,
                                  hidden_n_node_feat=24, out_n_node_feat=24, use_tanh=False, message_norm=3.0,
                                  kp_rad=4.6, k_closest=0, norm=False, fix_pos=True), [35, 48]),
    "enc_gvp_knn": ("gvp", dict(in_scalar_size=10, out_scalar_size=32, n_message_gvps=3, n_update_gvps=2, vector_size=4,
                                n_rr_convs=2, n_rk_convs=2, message_norm=10.0, k_closest=4, kp_rad=0, dropout=0.1,
                                n_keypoints=6, graph_cutoffs=CUT), [40, 57, 33]),
    "enc_gvp_mean_rad": ("gvp", dict(in_scalar_size=10, out_scalar_size=24, n_message_gvps=2, n_update_gvps=1, vector_size=4,
                                     n_rr_convs=2, n_rk_convs=2, message_norm="mean", k_closest=0, kp_rad=6.0, dropout=0.0,
                                     n_keypoints=5, graph_cutoffs=CUT), [35, 48]),
    "enc_gvp_zero": ("gvp", dict(in_scalar_size=10, out_scalar_size=24, n_message_gvps=2, n_update_gvps=1, vector_size=4,
                                 n_rr_convs=1, n_rk_convs=2, message_norm=0, k_closest=3, kp_rad=0, dropout=0.0,
                                 n_keypoints=5, graph_cutoffs=CUT), [35, 48]),
}


def raw_graph(dgl, pocket_id, n_atoms, n_kp):
    """The reference's build_initial_complex_graph (data_processing/pdbbind_processing.py:221-274) on a synthetic pocket,
    over the stand-ins."""
    x, h, res = synthetic.raw_pocket(pocket_id, n_atoms)
    n = x.shape[0]
    e = G.radius_graph(x, CUT["rr"], torch.zeros(n, dtype=torch.long), False, 100)
    no = (torch.zeros(0, dtype=torch.long), torch.zeros(0, dtype=torch.long))
    g = dgl.heterograph({("rec", "rr", "rec"): (e[0], e[1]),
                         ("rec", "rk", "kp"): (torch.arange(n).repeat(n_kp), torch.arange(n_kp).repeat_interleave(n)),
                         ("kp", "kk", "kp"): no, ("kp", "kl", "lig"): no, ("lig", "ll", "lig"): no, ("lig", "lk", "kp"): no},
                        num_nodes_dict={"rec": n, "kp": n_kp, "lig": 0})
    g.nodes["rec"].data["x_0"] = x
    g.nodes["rec"].data["h_0"] = h
    g.edges["rr"].data["same_res"] = (res[e[0]] == res[e[1]]).view(-1, 1)
    return g, (x, h, res)


def main():
    ref = loader.import_reference()
    dgl = ref.dgl
    enc_mod = importlib.import_module("models.receptor_encoder")
    gvp_mod = importlib.import_module("models.receptor_encoder_gvp")
    for name, (kind, kw, sizes) in CASES.items():
        torch.manual_seed(21)
        enc = (enc_mod.ReceptorEncoder if kind == "egnn" else gvp_mod.ReceptorEncoderGVP)(**kw).eval()
        with torch.no_grad():
            for pname, p in enc.named_parameters():
                if ".coord_mlp." in pname and pname.endswith(".2.weight"):
                    p.mul_(300.0)          # the gain=0.001 init would leave the coordinate path untested
                if pname.endswith(".Wh") or pname.endswith(".Wu"):
                    p.mul_(2.5)
                # larger query/key projections spread the keypoints over the pocket, so that the kNN / radius / kk edge
                # sets are non-trivial
                if pname.endswith("fc_src.weight") or pname.endswith("src_net.weight"):
                    p.mul_(1.5)
        parts = [raw_graph(dgl, 100 + i, n, kw["n_keypoints"]) for i, n in enumerate(sizes)]
        g = dgl.batch([p[0] for p in parts])
        with torch.no_grad():
            out = enc(g, ref.utils.get_batch_idxs(g))
        kp = out.nodes["kp"].data
        fx = {"kind": kind, "kwargs": kw,
              "pockets": [{"x": x, "h": h, "res": res} for _, (x, h, res) in parts],
              "rr": torch.stack(g.edges(form="uv", etype="rr")), "rr_n": g.batch_num_edges("rr").clone(),
              "state_dict": {k: v.clone() for k, v in enc.state_dict().items()},
              "kp_x": kp["x_0"].clone(), "kp_h": kp["h_0"].clone(), "kp_v": kp["v_0"].clone() if "v_0" in kp else None,
              "kk": torch.stack(out.edges(form="uv", etype="kk")), "kk_n": out.batch_num_edges("kk").clone(),
              "rk": torch.stack(out.edges(form="uv", etype="rk")), "rk_n": out.batch_num_edges("rk").clone()}
        torch.save(fx, OUT / f"{name}.pt")
        print(name, "kp_x", float(fx["kp_x"].abs().max()), "kp_h", float(fx["kp_h"].abs().max()),
              "kp_v", None if fx["kp_v"] is None else float(fx["kp_v"].abs().max()), "kk", fx["kk"].shape[1], "rk", fx["rk"].shape[1])


if __name__ == "__main__":
    main()
