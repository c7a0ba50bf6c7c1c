"""Debug helper (not a test): bisect the GVP CUDA path against oracle intermediates."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import flat, params as P  # noqa: E402
from helpers import flat_batch, rel_err  # noqa: E402
from keypoint_diffusion_b200 import ops, synthetic  # noqa: E402
import torch.nn.functional as F  # noqa: E402

dev = torch.device("cuda:0")


def case(n_convs, n_msg, n_upd, n_noise, S=32, V=4, mn=10.0):
    rec_nf = 12
    shapes = P.gvp_dynamics_shapes(10, rec_nf, V, n_convs, S, True, n_msg, n_upd, n_noise)
    sd = P.init_state_dict(shapes, seed=1)
    for k in sd:
        if k.endswith(".Wh") or k.endswith(".Wu"):
            sd[k] = sd[k] * 2.0
    cut = {"ll": 3.5, "kl": 8, "kk": 8}
    n_lig = [5, 2, 9]
    pockets = [synthetic.keypoint_pocket(i, 6, rec_nf, V, 8.0) for i in range(2)]
    x_l, h_l = synthetic.ligand_noise_state(n_lig, 10, seed=13)
    kp_x, kp_h, kp_v, ks, kd, off = [], [], [], [], [], 0
    for i in range(len(n_lig)):
        pk = pockets[i % 2]
        kp_x.append(pk.kp_x); kp_h.append(pk.kp_h); kp_v.append(pk.kp_v)
        ks.append(pk.kk_src + off); kd.append(pk.kk_dst + off); off += pk.n_kp
    inputs = {"lig_n": torch.tensor(n_lig), "kp_n": torch.tensor([6] * 3), "lig_x": x_l, "lig_h": h_l,
              "kp_x": torch.cat(kp_x), "kp_h": torch.cat(kp_h), "kp_v": torch.cat(kp_v),
              "kk_src": torch.cat(ks), "kk_dst": torch.cat(kd)}
    cfg = flat.GVPConfig(10, rec_nf, V, n_convs, S, mn, True, 0, 3, n_msg, n_upd, n_noise, cut)
    fb = flat_batch(inputs)
    t = torch.full((3,), 0.5)
    ref_h, ref_x, edges, counts = flat.gvp_forward(sd, cfg, fb, t, return_edges=True)

    sdd = {k[len("dynamics."):]: v for k, v in sd.items()}
    model = ops.GvpModel(sdd, n_lig_scalars=10, n_kp_scalars=rec_nf, vector_size=V, n_convs=n_convs, n_hidden_scalars=S,
                         update_kp=True, n_message_gvps=n_msg, n_update_gvps=n_upd, n_noise_gvps=n_noise,
                         message_norm=mn, device=dev)
    batch = ops.DeviceBatch(n_lig, [6] * 3, dev)
    kk = ops.Csr.from_edges(inputs["kk_src"], inputs["kk_dst"], batch.n_kp, dev)
    gp = ops.GraphParams(ll_k=0, ll_r=3.5, kl_k=3, kl_r=8)
    graphs = ops.LigandGraphs(batch, gp, True).build(x_l.to(dev), inputs["kp_x"].to(dev))
    eps_h, eps_x = model.forward(batch, graphs, kk, h_l.to(dev), x_l.to(dev), inputs["kp_h"].to(dev), inputs["kp_x"].to(dev),
                                 inputs["kp_v"].to(dev), torch.full((1,), 0.5, device=dev))
    torch.cuda.synchronize()
    print(f"convs={n_convs} msg={n_msg} upd={n_upd} noise={n_noise}: eps_h {rel_err(eps_h.cpu(), ref_h):.2e} "
          f"eps_x {rel_err(eps_x.cpu(), ref_x):.2e}")

    if n_convs == 1:
        # workspace layout (gvp_carve): s0 v0 s1 v1 | sm0 vm0 part0 | sm1 vm1 part1 ...
        ws = next(iter(model._ws.values())).view(torch.float32)
        N_l, N_k = batch.n_lig, batch.n_kp

        def al(n):  # Carver aligns every take to 256 bytes = 64 floats
            return (n + 63) // 64 * 64
        o = 0
        s0 = ws[o:o + N_l * S].view(N_l, S).cpu(); o = al(o + N_l * S)
        v0 = ws[o:o + N_l * V * 3].view(N_l, V, 3).cpu(); o = al(o + N_l * V * 3)
        o = al(o + N_k * S); o = al(o + N_k * V * 3)
        sm0 = ws[o:o + N_l * S].view(N_l, S).cpu(); o = al(o + N_l * S)
        vm0 = ws[o:o + N_l * V * 3].view(N_l, V, 3).cpu()
        # oracle intermediates for etype ll of layer 0
        lig_b, kp_b = fb.batch_idx()
        tt = t

        def enc(name, x):
            y = F.silu(flat._lin(sd, name + ".0", x))
            return F.layer_norm(y, (y.shape[1],), sd[name + ".2.weight"], sd[name + ".2.bias"], 1e-5)
        lig_s = enc("dynamics.lig_encoder", torch.cat([fb.lig_h, tt[lig_b].view(-1, 1)], 1))
        es, ed = edges["ll"]
        x_diff = fb.lig_x[es] - fb.lig_x[ed]
        dij = flat._norm_no_nan(x_diff, keepdims=True) + 1e-8
        x_diff = x_diff / dij
        rbf = flat._rbf(dij.squeeze(1), D_max=15.0, D_count=16)
        vec = torch.cat([x_diff.unsqueeze(1), torch.zeros(es.numel(), V, 3)], 1)
        sca = torch.cat([lig_s[es], rbf], 1)
        for i in range(n_msg):
            sca, vec = flat.gvp_apply(sd, f"dynamics.noise_predictor.conv_layers.0.edge_message_fns.lig_ll_lig.{i}", sca, vec)
        agg_s = torch.zeros(N_l, S).index_add_(0, ed, sca)
        agg_v = torch.zeros(N_l, V, 3).index_add_(0, ed, vec)
        has = torch.bincount(ed, minlength=N_l) > 0
        print("   sm[ll] err", rel_err(sm0[has], agg_s[has]), " vm[ll] err", rel_err(vm0[has], agg_v[has]),
              " |agg_s|", float(agg_s.abs().max()), "|agg_v|", float(agg_v.abs().max()))
        print("   sm0 row0[:6]", sm0[has][0, :6].tolist(), "\n   ref row0[:6]", agg_s[has][0, :6].tolist())


for args in [(1, 1, 1, 1), (1, 2, 1, 1), (1, 3, 2, 4), (2, 1, 1, 1), (3, 3, 2, 4)]:
    case(*args)
