"""Flat-tensor CPU restatement of the sampling hot path (oracle).

Follows, on flat tensors + explicit edge lists (no DGL):
  models/ligand_diffuser.py  :185-203 (remove_com), :342-370, :403-410, :437-447,
                             :462-469 (sample_from_encoded_receptors), :497-538
                             (sample_p_zs_given_zt)
  models/dynamics.py         :89-122 (message), :124-217 (LigRecConv.forward),
                             :266-294 (LigRecEGNN.forward), :342-420 (LigRecDynamics)
  models/dynamics_gvp.py     :38-44, :93-101, :149-234
  models/gvp.py              :12-41, :89-116, :152-166, :459-550
  utils.py                   :92-98, :158-170

Works in fp32 (parity checker, CPU baseline) or fp64 (to tell fp32 reassociation
noise from real bugs).  Aggregations use index_add_ (sequential, deterministic on CPU).

Test infrastructure only (see oracle/__init__.py): never imported by the product.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import graph as G
from . import schedule as S
from .params import gvp_layer_etypes


# --------------------------------------------------------------------------- batch

@dataclass
class FlatBatch:
    """A batch of already-encoded complexes; nodes of one complex are contiguous."""
    lig_n: torch.Tensor            # int64 [B] ligand atoms per complex
    kp_n: torch.Tensor             # int64 [B] keypoints per complex
    kp_x: torch.Tensor             # [N_k, 3]
    kp_h: torch.Tensor             # [N_k, C]
    kk_src: torch.Tensor           # int64 [E_kk] (global kp indices)
    kk_dst: torch.Tensor
    kp_v: Optional[torch.Tensor] = None   # [N_k, V, 3] (GVP)
    lig_x: Optional[torch.Tensor] = None  # [N_l, 3]
    lig_h: Optional[torch.Tensor] = None  # [N_l, F]

    @property
    def B(self):
        return int(self.lig_n.numel())

    def batch_idx(self):
        """utils.py:158-170"""
        ar = torch.arange(self.B)
        return ar.repeat_interleave(self.lig_n), ar.repeat_interleave(self.kp_n)

    def kk_per_batch(self):
        _, kb = self.batch_idx()
        return G.edges_per_batch(self.kk_dst, self.B, kb)


def _lin(sd, name, x):
    w = sd[name + ".weight"].to(x.dtype)
    b = sd.get(name + ".bias")
    y = x @ w.t()
    return y + b.to(x.dtype) if b is not None else y


def _segment_mean(x, batch, B):
    out = torch.zeros(B, x.shape[1], dtype=x.dtype)
    out.index_add_(0, batch, x)
    cnt = torch.bincount(batch, minlength=B).to(x.dtype)
    return out / cnt[:, None]


# --------------------------------------------------------------------------- graph build

def build_lig_edges(lig_x, kp_x, lig_b, kp_b, B, cutoffs, ll_k, kl_k, update_kp):
    """dynamics.py:387-420 / dynamics_gvp.py:201-234 (add_lig_edges).  Returns dict etype ->
    (src, dst) and dict etype -> per-complex counts."""
    if ll_k > 0:
        ll = G.knn_graph(lig_x, k=ll_k, batch=lig_b)
    else:
        ll = G.radius_graph(lig_x, r=cutoffs["ll"], batch=lig_b, max_num_neighbors=200)
    if kl_k > 0:
        kl = G.knn(x=lig_x, y=kp_x, k=kl_k, batch_x=lig_b, batch_y=kp_b)
    else:
        kl = G.radius(x=lig_x, y=kp_x, r=cutoffs["kl"], batch_x=lig_b, batch_y=kp_b,
                      max_num_neighbors=100)
    edges = {"ll": (ll[0], ll[1]), "kl": (kl[0], kl[1])}
    counts = {"ll": G.edges_per_batch(ll[0], B, lig_b), "kl": G.edges_per_batch(kl[0], B, kp_b)}
    if update_kp:
        edges["lk"] = (kl[1], kl[0])
        counts["lk"] = counts["kl"]
    return edges, counts


# --------------------------------------------------------------------------- EGNN

@dataclass
class EGNNConfig:
    atom_nf: int
    rec_nf: int
    n_layers: int = 4
    hidden_nf: int = 255
    use_tanh: bool = False
    message_norm: float = 1
    update_kp_feat: bool = False
    norm: bool = False
    ll_k: int = 0
    kl_k: int = 0
    graph_cutoffs: Dict[str, float] = field(default_factory=dict)
    coords_range: float = 10.0
    # dynamics.py:188-192 divides h_neigh / x_neigh by z through ``graph.ndata[key][ntype] = ...``.
    # On a multi-node-type DGL graph ``graph.ndata[key]`` returns a fresh dict per access
    # (dgl/view.py HeteroNodeDataView.__getitem__), so the assignment never reaches the graph
    # and :195 re-reads the UN-normalised sums.  That is the behaviour the shipped weights were
    # trained under, hence the default.  True = the normalisation the code comments intend.
    z_effective: bool = False


def _mlp2_silu(sd, name, x):
    # Sequential(Linear, SiLU, Linear, SiLU)
    return F.silu(_lin(sd, name + ".2", F.silu(_lin(sd, name + ".0", x))))


def egnn_forward(sd, cfg: EGNNConfig, batch: FlatBatch, t, prefix="dynamics.", edges=None,
                 return_edges=False):
    """LigRecDynamics.forward (dynamics.py:342-385).  t: [B]."""
    p = prefix
    dt = batch.lig_x.dtype
    lig_b, kp_b = batch.batch_idx()
    B = batch.B
    t = t.to(dt)

    lig_feat = _mlp2_silu(sd, p + "lig_encoder", batch.lig_h)                        # :355
    if (p + "rec_encoder.0.weight") in sd:
        kp_feat = _mlp2_silu(sd, p + "rec_encoder", batch.kp_h)                      # :356
    else:
        kp_feat = batch.kp_h
    lig_feat = torch.cat([lig_feat, t[lig_b].view(-1, 1)], dim=1)                    # :359-363
    kp_feat = torch.cat([kp_feat, t[kp_b].view(-1, 1)], dim=1)

    if edges is None:
        edges, counts = build_lig_edges(batch.lig_x, batch.kp_x, lig_b, kp_b, B, cfg.graph_cutoffs,
                                        cfg.ll_k, cfg.kl_k, cfg.update_kp_feat)      # :370
    else:
        edges, counts = edges
    edges = dict(edges)
    counts = dict(counts)
    edges["kk"] = (batch.kk_src, batch.kk_dst)
    counts["kk"] = batch.kk_per_batch()

    etypes = ["ll", "kl", "lk", "kk"] if cfg.update_kp_feat else ["ll", "kl"]
    upd = ["lig", "kp"] if cfg.update_kp_feat else ["lig"]
    src_nt = {"ll": "lig", "kl": "kp", "lk": "lig", "kk": "kp"}
    dst_nt = {"ll": "lig", "kl": "lig", "lk": "kp", "kk": "kp"}
    bidx = {"lig": lig_b, "kp": kp_b}
    nnodes = {"lig": batch.lig_n, "kp": batch.kp_n}

    # z (dynamics.py:277-285)
    z = {}
    for nt in upd:
        if cfg.message_norm == 0:
            tot = torch.stack([counts[et] for et in etypes if et[-1] == nt[0]], dim=0).sum(dim=0)
            zz = (tot / nnodes[nt]).to(torch.float32).to(dt)   # int64/int64 -> fp32 in the reference
            z[nt] = zz[bidx[nt]].view(-1, 1) + 1
        else:
            z[nt] = cfg.message_norm

    h = {"lig": lig_feat, "kp": kp_feat}
    x = {"lig": batch.lig_x, "kp": batch.kp_x}
    for l in range(cfg.n_layers):
        q = f"{p}egnn.conv_layers.{l}."
        h_neigh = {nt: torch.zeros_like(h[nt]) for nt in upd}
        x_neigh = {nt: torch.zeros_like(x[nt]) for nt in upd}
        for et in etypes:
            s, d = edges[et]
            xs, xd = x[src_nt[et]], x[dst_nt[et]]
            x_diff = xs[s] - xd[d]                                                   # :160
            dij = torch.linalg.vector_norm(x_diff, dim=1).unsqueeze(-1)              # :211
            x_diff = x_diff / (dij + 1)                                              # :169
            f = torch.cat([h[src_nt[et]][s], h[dst_nt[et]][d], dij], dim=-1)         # :103-105
            msg_h = _mlp2_silu(sd, f"{q}edge_mlp.{et}", f)                           # :111
            msg_h = msg_h * torch.sigmoid(_lin(sd, f"{q}soft_attention.{et}.0", msg_h))  # :112
            c = _lin(sd, f"{q}coord_mlp.{et}.4", _mlp2_silu(sd, f"{q}coord_mlp.{et}", f))
            if cfg.use_tanh:                                                         # :117-120
                msg_x = torch.tanh(c) * x_diff * cfg.coords_range
            else:
                msg_x = c * x_diff
            h_neigh[dst_nt[et]].index_add_(0, d, msg_h)                              # :177-185
            x_neigh[dst_nt[et]].index_add_(0, d, msg_x)
        h_new, x_new = {}, {}
        for nt in upd:
            if cfg.z_effective:
                hn = h_neigh[nt] / z[nt]                                             # :188-192 (as intended)
                xn = x_neigh[nt] / z[nt]
            else:
                hn, xn = h_neigh[nt], x_neigh[nt]                                    # :188-195 as executed
            inp = torch.cat([h[nt], hn], dim=1)                                      # :202
            y = _lin(sd, f"{q}node_mlp.{nt}.2", F.silu(_lin(sd, f"{q}node_mlp.{nt}.0", inp)))
            y = h[nt] + y                                                            # :203
            if cfg.norm:
                y = F.layer_norm(y, (y.shape[1],), sd[f"{q}layer_norm.{nt}.weight"].to(dt),
                                 sd[f"{q}layer_norm.{nt}.bias"].to(dt), 1e-5)        # :204
            h_new[nt] = y
            x_new[nt] = x[nt] + xn                                                   # :206
        if "kp" not in h_new:                                                        # :289-291
            h_new["kp"], x_new["kp"] = kp_feat, batch.kp_x
        h, x = h_new, x_new

    hl = h["lig"][:, :-1]                                                            # :376
    eps_h = _lin(sd, p + "lig_decoder.2", F.silu(_lin(sd, p + "lig_decoder.0", hl)))  # :380
    eps_x = x["lig"] - batch.lig_x                                                   # :381
    if return_edges:
        return eps_h, eps_x, edges, counts
    return eps_h, eps_x


# --------------------------------------------------------------------------- GVP

@dataclass
class GVPConfig:
    n_lig_scalars: int
    n_kp_scalars: int
    vector_size: int = 16
    n_convs: int = 4
    n_hidden_scalars: int = 128
    message_norm: object = 1       # float | 'mean' | 0
    update_kp: bool = False
    ll_k: int = 0
    kl_k: int = 0
    n_message_gvps: int = 3
    n_update_gvps: int = 2
    n_noise_gvps: int = 3
    graph_cutoffs: Dict[str, float] = field(default_factory=dict)
    rbf_dmax: float = 15.0
    rbf_dim: int = 16


def _norm_no_nan(x, axis=-1, keepdims=False, eps=1e-8, sqrt=True):
    out = torch.clamp(torch.sum(torch.square(x), axis, keepdims), min=eps)           # gvp.py:12-19
    return torch.sqrt(out) if sqrt else out


def _rbf(D, D_min=0.0, D_max=20.0, D_count=16):
    D_mu = torch.linspace(D_min, D_max, D_count).to(D.dtype).view(1, -1)             # gvp.py:26-41
    D_sigma = (D_max - D_min) / D_count
    return torch.exp(-((D.unsqueeze(-1) - D_mu) / D_sigma) ** 2)


def gvp_apply(sd, name, feats, vectors, vec_act="sigmoid"):
    """GVP.forward (gvp.py:89-116) with vector gating."""
    dt = feats.dtype
    Wh, Wu = sd[name + ".Wh"].to(dt), sd[name + ".Wu"].to(dt)
    Vh = torch.einsum("bvc,vh->bhc", vectors, Wh)
    Vu = torch.einsum("bhc,hu->buc", Vh, Wu)
    sh = _norm_no_nan(Vh)
    s = torch.cat((feats, sh), dim=1)
    feats_out = F.silu(_lin(sd, name + ".to_feats_out.0", s))
    gating = _lin(sd, name + ".scalar_to_vector_gates", feats_out).unsqueeze(-1)
    if vec_act == "sigmoid":
        gating = torch.sigmoid(gating)
    return feats_out, gating * Vu


def gvp_layernorm(sd, name, feats, vectors, eps=1e-5):
    """GVPLayerNorm.forward (gvp.py:159-166)."""
    dt = feats.dtype
    nf = F.layer_norm(feats, (feats.shape[1],), sd[name + ".feat_norm.weight"].to(dt),
                      sd[name + ".feat_norm.bias"].to(dt), 1e-5)
    vn = _norm_no_nan(vectors, axis=-1, keepdims=True, sqrt=False)
    vn = torch.sqrt(torch.mean(vn, dim=-2, keepdim=True) + eps) + eps
    return nf, vectors / vn


def gvp_forward(sd, cfg: GVPConfig, batch: FlatBatch, t, prefix="dynamics.", edges=None,
                return_edges=False):
    """LigRecDynamicsGVP.forward (dynamics_gvp.py:149-199)."""
    p = prefix
    dt = batch.lig_x.dtype
    lig_b, kp_b = batch.batch_idx()
    B = batch.B
    t = t.to(dt)
    V = cfg.vector_size

    def enc(name, x):  # Sequential(Linear, SiLU, LayerNorm)  dynamics_gvp.py:124-134
        y = F.silu(_lin(sd, name + ".0", x))
        return F.layer_norm(y, (y.shape[1],), sd[name + ".2.weight"].to(dt), sd[name + ".2.bias"].to(dt), 1e-5)

    lig_s = enc(p + "lig_encoder", torch.cat([batch.lig_h, t[lig_b].view(-1, 1)], dim=1))  # :161-169
    kp_s = enc(p + "kp_encoder", torch.cat([batch.kp_h, t[kp_b].view(-1, 1)], dim=1))
    lig_v = torch.zeros(lig_s.shape[0], V, 3, dtype=dt)                              # :179-184
    kp_v = batch.kp_v.to(dt)

    if edges is None:
        edges, counts = build_lig_edges(batch.lig_x, batch.kp_x, lig_b, kp_b, B, cfg.graph_cutoffs,
                                        cfg.ll_k, cfg.kl_k, cfg.update_kp)           # :192
    else:
        edges, counts = edges
    edges = dict(edges)
    counts = dict(counts)
    edges["kk"] = (batch.kk_src, batch.kk_dst)
    counts["kk"] = batch.kk_per_batch()

    bidx = {"lig": lig_b, "kp": kp_b}
    nnodes = {"lig": batch.lig_n, "kp": batch.kp_n}
    s = {"lig": lig_s, "kp": kp_s}
    v = {"lig": lig_v, "kp": kp_v}
    x = {"lig": batch.lig_x, "kp": batch.kp_x}

    for l in range(cfg.n_convs):
        etypes = gvp_layer_etypes(l, cfg.n_convs, cfg.update_kp)
        dst_ntypes = sorted(set(e[2] for e in etypes))
        q = f"{p}noise_predictor.conv_layers.{l}."
        s_msg = {nt: torch.zeros_like(s[nt]) for nt in dst_ntypes}
        v_msg = {nt: torch.zeros_like(v[nt]) for nt in dst_ntypes}
        for et in etypes:
            sn, name, dn = et
            es, ed = edges[name]
            x_diff = x[sn][es] - x[dn][ed]                                           # gvp.py:474
            dij = _norm_no_nan(x_diff, keepdims=True) + 1e-8                         # :478
            x_diff = x_diff / dij                                                    # :479
            d_rbf = _rbf(dij.squeeze(1), D_max=cfg.rbf_dmax, D_count=cfg.rbf_dim)    # :480
            vec = torch.cat([x_diff.unsqueeze(1), v[sn][es]], dim=1)                 # :545
            sca = torch.cat([s[sn][es], d_rbf], dim=1)                               # :547
            key = "_".join(et)
            for i in range(cfg.n_message_gvps):
                sca, vec = gvp_apply(sd, f"{q}edge_message_fns.{key}.{i}", sca, vec)  # :549
            agg_s = torch.zeros_like(s[dn])
            agg_v = torch.zeros_like(v[dn])
            agg_s.index_add_(0, ed, sca)
            agg_v.index_add_(0, ed, vec)
            if cfg.message_norm == "mean":                                           # :386-389 fn.mean
                deg = torch.bincount(ed, minlength=s[dn].shape[0]).to(dt).clamp(min=1)
                agg_s = agg_s / deg[:, None]
                agg_v = agg_v / deg[:, None, None]
            s_msg[dn] = s_msg[dn] + agg_s                                            # cross_reducer='sum'
            v_msg[dn] = v_msg[dn] + agg_v
        out_s, out_v = {}, {}
        for nt in dst_ntypes:
            nv = 1.0 if cfg.message_norm == "mean" else cfg.message_norm             # :373-383
            if nv == 0:                                                              # :504-507
                tot = torch.stack([counts[e[1]] for e in etypes if e[-1] == nt], dim=0).sum(dim=0)
                nvt = ((tot / nnodes[nt]).to(torch.float32) + 1).to(dt)
                nvt = nvt[bidx[nt]].unsqueeze(1)
                sm = s_msg[nt] / nvt
                vm = v_msg[nt] / nvt.unsqueeze(-1)
            else:
                sm = s_msg[nt] / nv
                vm = v_msg[nt] / nv
            sf = s[nt] + sm                                                          # :519-521
            vf = v[nt] + vm
            sf, vf = gvp_layernorm(sd, f"{q}message_layer_norms.{nt}", sf, vf)
            rs, rv = sf, vf
            for i in range(cfg.n_update_gvps):                                       # :524
                rs, rv = gvp_apply(sd, f"{q}node_update_fns.{nt}.{i}", rs, rv)
            sf = sf + rs                                                             # :530-532
            vf = vf + rv
            sf, vf = gvp_layernorm(sd, f"{q}update_layer_norms.{nt}", sf, vf)
            out_s[nt], out_v[nt] = sf, vf
        # a conv returns features only for its dst node types (gvp.py:501,536); the next conv
        # re-sets only what it is given (:465-469).  With the shipped configs the only case is the
        # last (lig-only) conv, after which kp is no longer read.
        for nt in dst_ntypes:
            s[nt], v[nt] = out_s[nt], out_v[nt]

    # NoisePredictionBlock (dynamics_gvp.py:38-44)
    q = f"{p}noise_predictor.noise_predictor."
    ns, nv_ = s["lig"], v["lig"]
    for i in range(cfg.n_noise_gvps):
        last = i == cfg.n_noise_gvps - 1
        ns, nv_ = gvp_apply(sd, f"{q}gvps.{i}", ns, nv_, vec_act="identity" if last else "sigmoid")
    eps_h = _lin(sd, q + "to_scalar_output", ns)
    eps_x = nv_.squeeze(1)
    if return_edges:
        return eps_h, eps_x, edges, counts
    return eps_h, eps_x


# --------------------------------------------------------------------------- diffusion loop

def remove_com(batch: FlatBatch, lig_b, kp_b, com):
    """ligand_diffuser.py:185-203"""
    if com == "ligand":
        c = _segment_mean(batch.lig_x, lig_b, batch.B)
    elif com == "receptor":
        c = _segment_mean(batch.kp_x, kp_b, batch.B)
    else:
        raise ValueError(com)
    batch.lig_x = batch.lig_x - c[lig_b]
    batch.kp_x = batch.kp_x - c[kp_b]
    return batch


def sample_p_zs_given_zt(dyn_fn, gamma, T, s_int, batch: FlatBatch, pos_noise, feat_noise):
    """ligand_diffuser.py:497-538 with injected noise (pos first, then feat: :530-531)."""
    lig_b, kp_b = batch.batch_idx()
    B = batch.B
    dt = batch.lig_x.dtype
    s = torch.full((B,), s_int) / T
    t = (torch.full((B,), s_int) + 1) / T
    gamma_s = S.gamma_lookup(gamma, s, T)
    gamma_t = S.gamma_lookup(gamma, t, T)
    sigma2_ts, sigma_ts, alpha_ts = S.sigma_and_alpha_t_given_s(gamma_t, gamma_s)
    sigma_s, sigma_t = S.sigma(gamma_s), S.sigma(gamma_t)
    eps_h, eps_x = dyn_fn(batch, t)                                                  # :513
    var_terms = sigma2_ts / alpha_ts / sigma_t                                       # :515
    a = alpha_ts[lig_b].view(-1, 1).to(dt)
    vt = var_terms[lig_b].view(-1, 1).to(dt)
    mu_pos = batch.lig_x / a - vt * eps_x                                            # :522-523
    mu_feat = batch.lig_h / a - vt * eps_h
    sig = (sigma_ts * sigma_s / sigma_t)[lig_b].view(-1, 1).to(dt)                   # :526-527
    batch.lig_x = mu_pos + sig * pos_noise
    batch.lig_h = mu_feat + sig * feat_noise
    return remove_com(batch, lig_b, kp_b, "ligand")                                  # :536


def sample_from_encoded_receptors(dyn_fn, gamma, T, batch: FlatBatch, init_lig_pos, noise_fn,
                                  lig_feat_norm_constant=1.0, n_steps=None, atom_nf=None):
    """ligand_diffuser.py:342-469 (visualize=False, use_fake_atoms=False).

    noise_fn(step, kind, shape) -> tensor supplies the Gaussian draws: step == -1 for the
    initial x_0 / h_0 (:366-367, kind 'x' then 'h'), step == s for the reverse step s
    (:530-531, 'x' then 'h').  ``n_steps`` < T runs only the first n_steps reverse steps
    (used to bound CPU-baseline time); the frame restore is applied regardless.
    """
    lig_b, kp_b = batch.batch_idx()
    B = batch.B
    dt = batch.kp_x.dtype
    init_kp_com = _segment_mean(batch.kp_x, kp_b, B)                                 # :348
    assert init_lig_pos.shape == (B, 3)                                              # :357
    batch.kp_x = batch.kp_x - init_lig_pos.to(dt)[kp_b]                              # :363
    N_l = int(batch.lig_n.sum())
    batch.lig_x = noise_fn(-1, "x", (N_l, 3)).to(dt)                                 # :366-367
    batch.lig_h = noise_fn(-1, "h", (N_l, atom_nf)).to(dt)
    batch = remove_com(batch, lig_b, kp_b, "ligand")                                 # :370
    steps = list(reversed(range(0, T)))
    if n_steps is not None:
        steps = steps[:n_steps]
    for s_int in steps:                                                              # :404-410
        pn = noise_fn(s_int, "x", (N_l, 3)).to(dt)
        fnz = noise_fn(s_int, "h", (N_l, atom_nf)).to(dt)
        batch = sample_p_zs_given_zt(dyn_fn, gamma, T, s_int, batch, pn, fnz)
    batch = remove_com(batch, lig_b, kp_b, "receptor")                               # :438
    batch.lig_x = batch.lig_x + init_kp_com[lig_b]                                   # :443-444
    batch.kp_x = batch.kp_x + init_kp_com[kp_b]
    batch.lig_h = batch.lig_h * lig_feat_norm_constant                               # :447
    ptr = torch.zeros(B + 1, dtype=torch.long)
    ptr[1:] = torch.cumsum(batch.lig_n, 0)
    lig_pos = [batch.lig_x[ptr[i]:ptr[i + 1]].clone() for i in range(B)]             # :462-469
    lig_feat = [batch.lig_h[ptr[i]:ptr[i + 1]].clone() for i in range(B)]
    return lig_pos, lig_feat
