"""Golden fixtures for the batch driver KeypointDiffusion._sample: tests/golden/sample_*.pt.

Runs the REFERENCE's own models/ligand_diffuser.py:_sample (imported read-only from /root/reference over the DGL /
torch_cluster / torch_scatter stand-ins of oracle/ref_shim) end to end on raw pocket graphs: receptor encoder ->
utils.copy_graph with the requested ligand sizes -> diffusion batches of diff_batch_size (which straddle receptors) ->
sample_from_encoded_receptors per batch -> regrouping per receptor.  Stores inputs, weights, the seed of the global
generator and the outputs.  Build container only; the fixtures travel to the GPU box.

    python tests/golden/make_golden_sample.py
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))

from oracle.ref_shim import loader  # noqa: E402
from make_golden import rescale_coord_layers  # noqa: E402
from make_golden_encoders import CUT, raw_graph  # noqa: E402

OUT = Path(__file__).resolve().parent
T = 8
N_LIG_ATOMS = [[5, 3, 7], [4, 6]]        # 2 receptors; diff_batch_size 2 -> batches [r0,r0] [r0,r1] [r1]
POCKET_ATOMS = [35, 48]
REF_LIG_ATOMS = [6, 4]                    # the reference ligand each raw graph carries (zero-filled by copy_graph: N10)

CASES = {
    "sample_egnn": dict(
        arch="egnn", rec_nf=24, use_ref_lig_com=False,
        graph=dict(n_keypoints=6, graph_cutoffs=CUT),
        dynamics=dict(n_layers=2, hidden_nf=32, use_tanh=True, message_norm=0.0, update_kp_feat=True, norm=True, ll_k=0, kl_k=3),
        rec_encoder=dict(n_convs=2, in_n_node_feat=10, use_sameres_feat=True, hidden_n_node_feat=32, out_n_node_feat=24,
                         use_tanh=True, coords_range=10, message_norm=0.0, kp_rad=0.0, k_closest=4, norm=True, fix_pos=False)),
    "sample_gvp": dict(
        arch="gvp", rec_nf=32, use_ref_lig_com=True,
        graph=dict(n_keypoints=6, graph_cutoffs=CUT),
        dynamics=dict(vector_size=4, n_convs=2, n_hidden_scalars=32, message_norm=10.0, update_kp=True, ll_k=0, kl_k=3,
                      n_message_gvps=2, n_update_gvps=1, n_noise_gvps=3, dropout=0.1),
        rec_encoder=dict(in_scalar_size=10, out_scalar_size=32, n_message_gvps=2, n_update_gvps=1, vector_size=4,
                         n_rr_convs=1, n_rk_convs=2, message_norm=10.0, k_closest=4, kp_rad=0, dropout=0.1)),
}


def main():
    ref = loader.import_reference()
    dgl = ref.dgl
    for name, c in CASES.items():
        torch.manual_seed(33)
        model = ref.ligand_diffuser.KeypointDiffusion(
            10, c["rec_nf"], processed_dataset_dir=Path(loader.REFERENCE_ROOT) / "data/bindingmoad_processed", n_timesteps=T,
            architecture=c["arch"], rec_encoder_type="learned", graph_config=c["graph"], dynamics_config=c["dynamics"],
            rec_encoder_config=c["rec_encoder"], rec_encoder_loss_config={"loss_type": "none"}, precision=1e-5,
            lig_feat_norm_constant=1).eval()
        rescale_coord_layers(model)
        with torch.no_grad():
            for pname, p in model.named_parameters():
                if pname.startswith("rec_encoder") and ".coord_mlp." in pname and pname.endswith(".2.weight"):
                    p.mul_(300.0)
                if pname.endswith("fc_src.weight") or pname.endswith("src_net.weight"):
                    p.mul_(1.5)
        graphs, pockets = [], []
        gen = torch.Generator().manual_seed(5)
        for i, (n_atoms, n_ref) in enumerate(zip(POCKET_ATOMS, REF_LIG_ATOMS)):
            g, (x, h, res) = raw_graph(dgl, 200 + i, n_atoms, c["graph"]["n_keypoints"])
            g.add_nodes(n_ref, ntype="lig")
            lx, lh = 3.0 * torch.randn(n_ref, 3, generator=gen) + 1.0, torch.randn(n_ref, 10, generator=gen)
            g.nodes["lig"].data["x_0"] = lx
            g.nodes["lig"].data["h_0"] = lh
            graphs.append(g)
            pockets.append({"x": x, "h": h, "res": res, "lig_x": lx, "lig_h": lh})
        torch.manual_seed(4321)
        with torch.no_grad():
            samples = model._sample(graphs, N_LIG_ATOMS, rec_enc_batch_size=1, diff_batch_size=2,
                                    use_ref_lig_com=c["use_ref_lig_com"])
        fx = {"kind": c["arch"], "rec_nf": c["rec_nf"], "T": T, "graph": c["graph"], "dynamics": c["dynamics"],
              "rec_encoder": c["rec_encoder"], "use_ref_lig_com": c["use_ref_lig_com"], "pockets": pockets,
              "n_lig_atoms": N_LIG_ATOMS, "diff_batch_size": 2, "noise_seed": 4321,
              "state_dict": {k: v.clone() for k, v in model.state_dict().items()},
              "samples": [{"positions": [p.clone() for p in s["positions"]], "features": [f.clone() for f in s["features"]]}
                          for s in samples]}
        torch.save(fx, OUT / f"{name}.pt")
        print(name, [[tuple(p.shape) for p in s["positions"]] for s in samples], float(samples[0]["positions"][0].abs().max()),
              float(samples[1]["features"][1].abs().max()))


if __name__ == "__main__":
    main()
