"""GPU: the batch drivers above the loop against the reference's own code, and the capacity-bucketed samplers.

* KeypointDiffusion._sample against fixtures produced by the reference's _sample (tests/golden/make_golden_sample.py:
  receptor encoder -> copies with requested ligand sizes -> diffusion batches that straddle receptors -> regrouping),
  with the reference's global-generator draws injected; bar 1e-3 relative over the whole (short) trajectory.
* capacity samplers: a padded, bucketed sampler gives the ligands of a sampler captured for the exact layout (<= 1e-5),
  and a second batch with other sizes / another pocket re-uses the captured graphs (no new capture).
* output decode on the device against utils.decode_ligands / torch.argmax.
"""
import pytest
import torch

from helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu


def _model_from_fixture(fx, dev):
    from pathlib import Path
    from keypoint_diffusion_b200 import KeypointDiffusion
    root = Path(__file__).resolve().parents[1]
    model = KeypointDiffusion(10, fx["rec_nf"], processed_dataset_dir=root / "data/bindingmoad_processed", n_timesteps=fx["T"],
                              architecture=fx["kind"], rec_encoder_type="learned", graph_config=fx["graph"],
                              dynamics_config=fx["dynamics"], rec_encoder_config=fx["rec_encoder"],
                              rec_encoder_loss_config={"loss_type": "none"}, precision=1e-5, lig_feat_norm_constant=1).eval()
    model.load_state_dict(fx["state_dict"], strict=True)
    return model.to(dev)


@pytest.mark.parametrize("mode", ["fp32", "bf16x3"])
@pytest.mark.parametrize("name", ["sample_egnn", "sample_gvp"])
def test_sample_driver_matches_reference(name, mode):
    from keypoint_diffusion_b200 import hetero
    dev = torch.device("cuda:0")
    fx = load_golden(name)
    model = _model_from_fixture(fx, dev)
    model.dynamics.set_precision(mode)
    cut, n_kp = fx["graph"]["graph_cutoffs"], fx["graph"]["n_keypoints"]
    graphs = [hetero.build_initial_complex_graph(p["x"], p["h"], p["res"], n_kp, cut, p["lig_x"], p["lig_h"]) for p in fx["pockets"]]
    # the reference draws x then h from the global generator at the start of every diffusion batch and at every step
    sizes = [n for rec in fx["n_lig_atoms"] for n in rec]
    dbs, T, F = fx["diff_batch_size"], fx["T"], 10
    torch.manual_seed(fx["noise_seed"])
    noise = []
    for b in range(0, len(sizes), dbs):
        N_l = sum(sizes[b:b + dbs])
        slots = []
        for _ in range(T + 1):
            nx, nh = torch.randn(N_l, 3), torch.randn(N_l, F)
            slots.append(torch.cat([nx.reshape(-1), nh.reshape(-1)]))
        noise.append(torch.stack(slots).to(dev).contiguous())
    out = model._sample(graphs, fx["n_lig_atoms"], rec_enc_batch_size=1, diff_batch_size=dbs,
                        use_ref_lig_com=fx["use_ref_lig_com"], noise=noise, steps_per_graph=4)
    assert len(out) == len(fx["samples"])
    worst = 0.0
    for got, ref in zip(out, fx["samples"]):
        assert len(got["positions"]) == len(ref["positions"])
        for gp_, rp, gf, rf in zip(got["positions"], ref["positions"], got["features"], ref["features"]):
            assert gp_.shape == rp.shape and gf.shape == rf.shape and gp_.device.type == "cpu"
            worst = max(worst, rel_err(gp_, rp), rel_err(gf, rf))
    print(f"_sample {name} [{mode}]: worst rel_err over all ligands {worst:.2e}")
    assert worst < 1e-3


def _bench_like_model(arch, dev, T=40):
    import os
    from pathlib import Path
    from shipped_cases import shipped_configs
    from keypoint_diffusion_b200 import model_from_config
    os.chdir(Path(__file__).resolve().parents[1])
    cfg = shipped_configs()["gvp_20kp" if arch == "gvp" else "egnn_20kp"]
    cfg["diffusion"]["n_timesteps"] = T
    torch.manual_seed(0)
    return model_from_config(cfg).to(dev).eval(), cfg


@pytest.mark.parametrize("arch", ["gvp", "egnn"])
def test_capacity_sampler_equals_exact_layout_and_reuses_graphs(arch):
    from keypoint_diffusion_b200 import HeteroBatch, synthetic
    dev = torch.device("cuda:0")
    model, cfg = _bench_like_model(arch, dev)
    vs = cfg["dynamics_gvp"]["vector_size"] if arch == "gvp" else 0
    width = cfg["rec_encoder_gvp"]["out_scalar_size"] if arch == "gvp" else cfg["rec_encoder"]["out_n_node_feat"]
    pockets = [synthetic.keypoint_pocket(i, 20, width, vs, cfg["graph"]["graph_cutoffs"]["kk"]) for i in range(3)]
    sizes_a = [20, 8, 35, 20, 13, 27, 22, 19, 31, 16, 25, 9, 20, 20, 18, 23, 30, 12, 21, 17, 26, 14, 20, 28, 11, 24, 20, 19, 33, 15, 22, 18]
    sizes_b = [s + (3 if i % 2 else -3) for i, s in enumerate(sizes_a)]          # other sizes, same totals -> same buckets
    sizes_b[0] += sum(sizes_a) - sum(sizes_b)
    results = {}
    for tag, sizes, pk in (("a", sizes_a, pockets[:2]), ("b", sizes_b, pockets[1:])):
        g = HeteroBatch.from_pockets(pk, sizes, 10).to(dev)
        init = torch.zeros(len(sizes), 3, device=dev)
        for capacity in (True, False):
            before = model.cold_captures
            x, h = model.sample_from_encoded_receptors(g, init_lig_pos=init, seed=99, steps_per_graph=10, sub_batches=2,
                                                       return_device_tensors=True, capacity=capacity)
            torch.cuda.synchronize()
            results[(tag, capacity)] = (x.clone(), h.clone(), model.cold_captures - before)
    for tag in ("a", "b"):
        xc, hc, _ = results[(tag, True)]
        xe, he, _ = results[(tag, False)]
        assert torch.isfinite(xc).all() and torch.isfinite(hc).all()
        # real complexes come first in the padded layout: same atom indices -> same noise; the only difference is the
        # association of fp32 partial sums in the edge tile that straddles the real / filler boundary (~1 ulp)
        ex, eh = rel_err(xc.cpu(), xe.cpu()), rel_err(hc.cpu(), he.cpu())
        print(f"capacity vs exact layout [{arch} {tag}]: rel_err pos {ex:.1e} feat {eh:.1e}")
        assert ex < 1e-5 and eh < 1e-5, tag
    assert results[("a", True)][2] == 2            # two concurrent groups -> two captured loops ...
    assert results[("b", True)][2] == 0, "another batch in the same capacity bucket must not capture again"


def test_decode_on_device_matches_host_decode():
    from keypoint_diffusion_b200 import ops, utils
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    h = torch.randn(977, 10, generator=g)
    h[5] = 0.25                                       # an all-ties row: lowest index wins
    h[6, 3] = h[6, 7] = 9.0
    got = ops.decode_atom_types(h.to(dev)).cpu()
    assert got.dtype == torch.int32 and torch.equal(got.long(), torch.argmax(h, dim=1))
    assert int(got[5]) == 0 and int(got[6]) == 3
    elements = ["C", "N", "O", "S", "P", "F", "Cl", "Br", "I", "B"]
    sizes = [20, 957]
    ref = utils.decode_ligands(list(torch.split(torch.zeros(977, 3), sizes)), list(torch.split(h, sizes)), elements)
    dec = utils.decode_ligands(list(torch.split(torch.zeros(977, 3), sizes)), None, elements,
                               atom_types=list(torch.split(got, sizes)))
    assert [e for _, e in ref] == [e for _, e in dec]


def test_sampler_returns_decoded_atom_types():
    from keypoint_diffusion_b200 import HeteroBatch, synthetic
    dev = torch.device("cuda:0")
    model, cfg = _bench_like_model("gvp", dev, T=20)
    pk = synthetic.keypoint_pocket(0, 20, cfg["rec_encoder_gvp"]["out_scalar_size"], 16, 8.0)
    g = HeteroBatch.from_pockets([pk], [20, 7, 31], 10).to(dev)
    pos, feat, types = model.sample_from_encoded_receptors(g, init_lig_pos=torch.zeros(3, 3, device=dev), seed=5,
                                                           steps_per_graph=10, decode=True)
    assert [t.shape[0] for t in types] == [20, 7, 31]
    for f, t in zip(feat, types):
        assert torch.equal(torch.argmax(f, dim=1), t.long())
