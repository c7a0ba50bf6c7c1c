"""Generate the golden fixtures under tests/golden/ (run in the build container only).

Executes the REFERENCE's own code -- models/dynamics.py, models/dynamics_gvp.py,
models/gvp.py, models/ligand_diffuser.py imported read-only from /root/reference -- over
the DGL / torch_cluster stand-ins in oracle/ref_shim, on small seeded models and batches,
and stores inputs, weights, built edge lists and outputs.  The fixtures travel to the GPU
box (which has no /root/reference); tests compare oracle/flat.py and the CUDA path with
them.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.pt, *.json

Nothing here is copied from the reference: only its *outputs* are stored.
"""
import json
import os
import sys
from pathlib import Path

import torch
import yaml

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle.ref_shim import loader  # noqa: E402
from oracle import params as P  # noqa: E402
from keypoint_diffusion_b200 import synthetic  # noqa: E402

OUT = Path(__file__).resolve().parent


def make_graph(dgl, pockets, n_lig, atom_nf, seed, with_v):
    """A batched shim heterograph of len(n_lig) complexes (complex i uses pockets[i % len])."""
    graphs = []
    x_l, h_l = synthetic.ligand_noise_state(n_lig, atom_nf, seed)
    off = 0
    for i, n in enumerate(n_lig):
        pk = pockets[i % len(pockets)]
        no = (torch.zeros(0, dtype=torch.long), torch.zeros(0, dtype=torch.long))
        data = {
            ("rec", "rr", "rec"): no, ("rec", "rk", "kp"): no,
            ("kp", "kk", "kp"): (pk.kk_src, pk.kk_dst),
            ("kp", "kl", "lig"): no, ("lig", "ll", "lig"): no, ("lig", "lk", "kp"): no,
        }
        g = dgl.heterograph(data, num_nodes_dict={"rec": 0, "kp": pk.n_kp, "lig": n})
        g.nodes["kp"].data["x_0"] = pk.kp_x.clone()
        g.nodes["kp"].data["h_0"] = pk.kp_h.clone()
        if with_v:
            g.nodes["kp"].data["v_0"] = pk.kp_v.clone()
        g.nodes["lig"].data["x_0"] = x_l[off:off + n].clone()
        g.nodes["lig"].data["h_0"] = h_l[off:off + n].clone()
        g.nodes["rec"].data["x_0"] = torch.zeros(0, 3)
        g.nodes["rec"].data["h_0"] = torch.zeros(0, 1)
        off += n
        graphs.append(g)
    return dgl.batch(graphs)


def graph_tensors(g, with_v):
    d = {
        "lig_n": g.batch_num_nodes("lig").clone(), "kp_n": g.batch_num_nodes("kp").clone(),
        "lig_x": g.nodes["lig"].data["x_0"].clone(), "lig_h": g.nodes["lig"].data["h_0"].clone(),
        "kp_x": g.nodes["kp"].data["x_0"].clone(), "kp_h": g.nodes["kp"].data["h_0"].clone(),
        "kk_src": g.edges(etype="kk")[0].clone(), "kk_dst": g.edges(etype="kk")[1].clone(),
    }
    if with_v:
        d["kp_v"] = g.nodes["kp"].data["v_0"].clone()
    return d


def capture_edges(dyn):
    """Record the edge lists the reference builds in add_lig_edges before it removes them."""
    cap = {}
    orig = dyn.remove_lig_edges

    def wrapped(g):
        for et in ("ll", "kl", "lk"):
            s, d = g.edges(form="uv", etype=et)
            cap[et] = torch.stack([s.clone(), d.clone()])
        return orig(g)

    dyn.remove_lig_edges = wrapped
    return cap


def rescale_coord_layers(module):
    # gain=0.001 init would leave the coordinate path untested on random weights
    with torch.no_grad():
        for name, p in module.named_parameters():
            if ".coord_mlp." in name and name.endswith(".4.weight"):
                p.mul_(300.0)
            # GVP: U(+-1/sqrt(v)) Wh/Wu products shrink the vector channel to ~1e-6 on random
            # weights; scale them so eps_x is O(0.1) and the vector path is really tested
            if name.endswith(".Wh") or name.endswith(".Wu"):
                p.mul_(2.5)


EGNN_CASES = {
    # name: (ctor kwargs, n_kp, kp feat, n_lig list)
    "egnn_small_kp": (dict(n_layers=2, hidden_nf=32, use_tanh=True, message_norm=0.0, update_kp_feat=True,
                           norm=True, ll_k=0, kl_k=3, n_keypoints=6,
                           graph_cutoffs={"ll": 3.0, "kl": 8, "kk": 8, "rk": 100, "rr": 3.5}),
                      6, 12, [5, 2, 9]),
    "egnn_small_nokp": (dict(n_layers=2, hidden_nf=31, use_tanh=False, message_norm=2.0, update_kp_feat=False,
                             norm=False, ll_k=3, kl_k=0, n_keypoints=5,
                             graph_cutoffs={"ll": 3.0, "kl": 9.0, "kk": 8, "rk": 100, "rr": 3.5}),
                        5, 31, [6, 4]),
}

GVP_CASES = {
    "gvp_small_sum": (dict(vector_size=4, n_convs=3, n_hidden_scalars=32, message_norm=10.0, update_kp=True,
                           ll_k=0, kl_k=3, n_message_gvps=3, n_update_gvps=2, n_noise_gvps=4, dropout=0.1,
                           n_keypoints=6, graph_cutoffs={"ll": 3.5, "kl": 8, "kk": 8, "rk": 100, "rr": 3.5}),
                      6, 12, [5, 2, 9]),
    "gvp_small_mean": (dict(vector_size=4, n_convs=2, n_hidden_scalars=32, message_norm="mean", update_kp=True,
                            ll_k=0, kl_k=2, n_message_gvps=2, n_update_gvps=1, n_noise_gvps=3, dropout=0.0,
                            n_keypoints=5, graph_cutoffs={"ll": 3.5, "kl": 8, "kk": 8, "rk": 100, "rr": 3.5}),
                       5, 10, [4, 7]),
    "gvp_small_zero": (dict(vector_size=4, n_convs=2, n_hidden_scalars=32, message_norm=0, update_kp=True,
                            ll_k=0, kl_k=2, n_message_gvps=2, n_update_gvps=1, n_noise_gvps=3, dropout=0.0,
                            n_keypoints=5, graph_cutoffs={"ll": 3.5, "kl": 8, "kk": 8, "rk": 100, "rr": 3.5}),
                       5, 10, [4, 7]),
}


def main():
    ref = loader.import_reference()
    dgl = ref.dgl
    atom_nf = 10
    t_vals = [0.001, 0.5, 1.0]

    # ---- state_dict key/shape inventory of the eight shipped configs
    inv = {}
    cwd = os.getcwd()
    os.chdir(loader.REFERENCE_ROOT)  # dataset.location in the YAMLs is relative
    try:
        for cfg_path in sorted(Path("trained_models").glob("*/config.yml")):
            cfg = yaml.safe_load(open(cfg_path))
            model = ref.model_setup.model_from_config(cfg)
            inv[cfg_path.parent.name] = {k: list(v.shape) for k, v in model.state_dict().items()}
    finally:
        os.chdir(cwd)
    json.dump(inv, open(OUT / "state_dict_shapes.json", "w"), indent=0, sort_keys=True)
    # the hyper-parameters of the shipped configs that the hot path reads (a reduced extract, so
    # that GPU-box tests can build the real model shapes without /root/reference)
    keep = {"dataset": ["lig_elements", "rec_elements", "location", "max_fake_atom_frac"],
            "diffusion": None, "dynamics": None, "dynamics_gvp": None, "graph": None,
            "rec_encoder": None, "rec_encoder_gvp": None, "rec_encoder_loss": None, "sampling_config": None}
    shipped = {}
    for cfg_path in sorted((Path(loader.REFERENCE_ROOT) / "trained_models").glob("*/config.yml")):
        cfg = yaml.safe_load(open(cfg_path))
        shipped[cfg_path.parent.name] = {
            sec: ({k: cfg[sec][k] for k in ks if k in cfg[sec]} if ks else cfg[sec])
            for sec, ks in keep.items() if sec in cfg}
    yaml.safe_dump(shipped, open(OUT / "shipped_configs.yml", "w"), sort_keys=True)
    print("state_dict inventory:", {k: len(v) for k, v in inv.items()})

    # ---- EGNN denoiser
    for name, (kw, n_kp, c, n_lig) in EGNN_CASES.items():
        torch.manual_seed(7)
        dyn = ref.dynamics.LigRecDynamics(atom_nf, c, **kw).eval()
        rescale_coord_layers(dyn)
        pockets = [synthetic.keypoint_pocket(i, n_kp, c, 0, kw["graph_cutoffs"]["kk"]) for i in range(2)]
        g = make_graph(dgl, pockets, n_lig, atom_nf, seed=11, with_v=False)
        inputs = graph_tensors(g, False)
        cap = capture_edges(dyn)
        outs = {}
        for t in t_vals:
            tt = torch.full((len(n_lig),), t)
            with torch.no_grad():
                eps_h, eps_x = dyn(g, tt, ref.utils.get_batch_idxs(g))
            outs[f"{t}"] = {"eps_h": eps_h.clone(), "eps_x": eps_x.clone()}
        # dynamics.forward must leave g unchanged (local_scope + remove_lig_edges)
        assert torch.equal(g.nodes["lig"].data["h_0"], inputs["lig_h"])
        assert g.num_edges("ll") == 0 and g.num_edges("kl") == 0
        sd = {"dynamics." + k: v.clone() for k, v in dyn.state_dict().items()}
        torch.save({"kind": "egnn", "kwargs": kw, "atom_nf": atom_nf, "rec_nf": c, "inputs": inputs,
                    "state_dict": sd, "edges": {k: v for k, v in cap.items()}, "outputs": outs},
                   OUT / f"{name}.pt")
        print(name, "eps_h", float(eps_h.abs().max()), "eps_x", float(eps_x.abs().max()),
              {k: v.shape[1] for k, v in cap.items()})

    # ---- GVP denoiser
    for name, (kw, n_kp, c, n_lig) in GVP_CASES.items():
        torch.manual_seed(9)
        dyn = ref.dynamics_gvp.LigRecDynamicsGVP(atom_nf, c, **kw).eval()
        rescale_coord_layers(dyn)
        pockets = [synthetic.keypoint_pocket(i, n_kp, c, kw["vector_size"], kw["graph_cutoffs"]["kk"])
                   for i in range(2)]
        g = make_graph(dgl, pockets, n_lig, atom_nf, seed=13, with_v=True)
        inputs = graph_tensors(g, True)
        cap = capture_edges(dyn)
        outs = {}
        for t in t_vals:
            tt = torch.full((len(n_lig),), t)
            with torch.no_grad():
                eps_h, eps_x = dyn(g, tt, ref.utils.get_batch_idxs(g))
            outs[f"{t}"] = {"eps_h": eps_h.clone(), "eps_x": eps_x.clone()}
        sd = {"dynamics." + k: v.clone() for k, v in dyn.state_dict().items()}
        torch.save({"kind": "gvp", "kwargs": kw, "atom_nf": atom_nf, "rec_nf": c, "inputs": inputs,
                    "state_dict": sd, "edges": {k: v for k, v in cap.items()}, "outputs": outs},
                   OUT / f"{name}.pt")
        print(name, "eps_h", float(eps_h.abs().max()), "eps_x", float(eps_x.abs().max()),
              {k: v.shape[1] for k, v in cap.items()})

    # ---- the whole reverse-diffusion loop through KeypointDiffusion (short schedule)
    for arch in ("egnn", "gvp"):
        T = 12
        if arch == "egnn":
            kw, n_kp, c, n_lig = EGNN_CASES["egnn_small_kp"]
        else:
            kw, n_kp, c, n_lig = GVP_CASES["gvp_small_sum"]
        kw = dict(kw)
        gc = kw.pop("graph_cutoffs")
        nk = kw.pop("n_keypoints")
        rec_cfg = (dict(in_n_node_feat=10, hidden_n_node_feat=16, out_n_node_feat=c, n_convs=1, k_closest=3)
                   if arch == "egnn" else
                   dict(in_scalar_size=10, out_scalar_size=c, vector_size=kw["vector_size"], n_rr_convs=1,
                        n_rk_convs=1, k_closest=3))
        torch.manual_seed(21)
        model = ref.ligand_diffuser.KeypointDiffusion(
            atom_nf, c, processed_dataset_dir=Path(loader.REFERENCE_ROOT) / "data/bindingmoad_processed",
            n_timesteps=T, architecture=arch, rec_encoder_type="learned",
            graph_config={"graph_cutoffs": gc, "n_keypoints": nk}, dynamics_config=kw,
            rec_encoder_config=rec_cfg, rec_encoder_loss_config={"loss_type": "none"},
            precision=1e-5, lig_feat_norm_constant=1).eval()
        rescale_coord_layers(model)
        pockets = [synthetic.keypoint_pocket(i, n_kp, c, kw.get("vector_size", 0), gc["kk"]) for i in range(2)]
        g = make_graph(dgl, pockets, n_lig, atom_nf, seed=17, with_v=(arch == "gvp"))
        inputs = graph_tensors(g, arch == "gvp")
        init_lig_pos = torch.tensor([[0.5, -0.25, 0.125]] * len(n_lig))
        torch.manual_seed(1234)
        with torch.no_grad():
            pos, feat = model.sample_from_encoded_receptors(g, init_lig_pos=init_lig_pos.clone())
        sd = {k: v.clone() for k, v in model.state_dict().items() if k.startswith(("dynamics.", "gamma."))}
        full_kw = dict(kw)
        full_kw["graph_cutoffs"] = gc
        full_kw["n_keypoints"] = nk
        torch.save({"kind": arch, "kwargs": full_kw, "atom_nf": atom_nf, "rec_nf": c, "T": T,
                    "precision": 1e-5, "inputs": inputs, "init_lig_pos": init_lig_pos,
                    "noise_seed": 1234, "state_dict": sd,
                    "positions": [p.clone() for p in pos], "features": [f.clone() for f in feat]},
                   OUT / f"loop_{arch}.pt")
        print("loop", arch, "pos[0][0]", pos[0][0].tolist())

    # ---- noise-schedule known answers straight from the reference class
    sched = ref.ligand_diffuser.PredefinedNoiseSchedule("polynomial_2", timesteps=1000, precision=1e-5)
    gamma = sched.gamma.detach().clone()
    kd = ref.ligand_diffuser.KeypointDiffusion
    rows = {}
    for s_int in (0, 1, 10, 500, 998, 999):
        s = torch.full((1,), s_int) / 1000
        t = (torch.full((1,), s_int) + 1) / 1000
        gs, gt = sched(s), sched(t)
        s2, s1, a = kd.sigma_and_alpha_t_given_s(None, gt, gs)
        sig_s, sig_t = kd.sigma(None, gs), kd.sigma(None, gt)
        rows[str(s_int)] = [float(a), float(s2 / a / sig_t), float(s1 * sig_s / sig_t)]
    torch.save({"gamma": gamma, "coef": rows}, OUT / "schedule.pt")
    print("schedule gamma[0,500,1000]:", float(gamma[0]), float(gamma[500]), float(gamma[1000]))


if __name__ == "__main__":
    main()
