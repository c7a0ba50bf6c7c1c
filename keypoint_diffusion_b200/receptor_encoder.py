"""Learned receptor encoders without DGL / torch-cluster / torch-scatter.

They turn a raw pocket graph (`rec` atoms with x_0 / h_0, `rr` edges [+ same_res flag], `kp` placeholders) into the
encoded pocket the sampling hot path starts from: `kp.x_0`, `kp.h_0` [, `kp.v_0`] and the `kk` edge list.  They run once
per pocket, outside the 1000-step loop, so this is flat-tensor PyTorch (gather / index_add over edge lists, dense
per-complex attention) on whatever device the graph lives on; the denoiser kernels are not involved.  Module and
parameter names mirror the reference so that the `rec_encoder.*` entries of a shipped checkpoint load unchanged:

  ReceptorEncoder       reference models/receptor_encoder.py:381-555 (ReceptorConv :14-153, RecKeyConv :156-300)
  ReceptorEncoderGVP    reference models/receptor_encoder_gvp.py:97-322 (KeypointInitializer :15-93),
                        GVPEdgeConv models/gvp.py:170-341, GVP :43-116, GVPLayerNorm :152-166

Graph searches restate torch_cluster's semantics: fp32 unfused squared distances, strict `<` radius
test with the neighbour cap applied in ascending source order, kNN with the lower index winning ties.
"""
import math
from typing import Dict, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import hetero

# ------------------------------------------------------------------ per-complex graph searches (dense, padded)


def _offsets(counts: torch.Tensor) -> torch.Tensor:
    off = torch.zeros(counts.shape[0] + 1, dtype=torch.long, device=counts.device)
    off[1:] = torch.cumsum(counts, 0)
    return off


def _padded(x: torch.Tensor, counts: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """[N, 3] rows grouped by complex -> ([B, nmax, 3], valid mask [B, nmax])."""
    B, nmax = counts.shape[0], int(counts.max()) if counts.numel() else 0
    slot = torch.arange(nmax, device=x.device)[None, :]
    valid = slot < counts[:, None]
    out = torch.zeros(B, nmax, 3, dtype=x.dtype, device=x.device)
    out[valid] = x
    return out, valid


def _pair_d2(y: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """[B, ny, nx] squared distances, each operation rounded to fp32 in the order ((dx*dx)+(dy*dy))+(dz*dz)."""
    dx = x[:, None, :, 0] - y[:, :, None, 0]
    dy = x[:, None, :, 1] - y[:, :, None, 1]
    dz = x[:, None, :, 2] - y[:, :, None, 2]
    return ((dx * dx) + (dy * dy)) + (dz * dz)


def knn_to(x: torch.Tensor, y: torch.Tensor, k: int, nx: torch.Tensor, ny: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """For every y (ascending) its k nearest x of the same complex, ascending distance, lower index first on ties;
    complexes with fewer than k points give fewer.  Returns (y_idx, x_idx) as global node indices."""
    xp, xv = _padded(x, nx)
    yp, yv = _padded(y, ny)
    d2 = _pair_d2(yp, xp).masked_fill(~xv[:, None, :], float("inf"))
    kk = min(k, d2.shape[-1])
    order = torch.sort(d2, dim=-1, stable=True).indices[..., :kk]                  # [B, ny, kk]
    keep = yv[:, :, None] & (torch.arange(kk, device=x.device)[None, None, :] < nx[:, None, None])
    xo, yo = _offsets(nx)[:-1], _offsets(ny)[:-1]
    yi = (torch.arange(d2.shape[1], device=x.device)[None, :, None] + yo[:, None, None]).expand_as(order)
    xi = order + xo[:, None, None]
    return yi[keep], xi[keep]


def radius_to(x: torch.Tensor, y: torch.Tensor, r: float, nx: torch.Tensor, ny: torch.Tensor, max_num_neighbors: int,
              drop_self: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """For every y (ascending) the x of the same complex with d^2 < r^2, ascending x, at most max_num_neighbors of them
    (drop_self: radius_graph semantics -- the cap is counted over max_num_neighbors + 1 hits including the point
    itself, which is then removed).  Returns (y_idx, x_idx) as global node indices."""
    xp, xv = _padded(x, nx)
    yp, yv = _padded(y, ny)
    r2 = torch.tensor(float(r) * float(r), dtype=torch.float64).to(torch.float32).to(x.device)
    hit = (_pair_d2(yp, xp) < r2) & xv[:, None, :] & yv[:, :, None]
    hit &= torch.cumsum(hit.long(), dim=-1) <= (max_num_neighbors + 1 if drop_self else max_num_neighbors)
    if drop_self:
        eye = torch.eye(hit.shape[1], hit.shape[2], dtype=torch.bool, device=x.device)
        hit &= ~eye[None]
    b, yi, xi = torch.nonzero(hit, as_tuple=True)                                    # sorted by (complex, y, x)
    return yi + _offsets(ny)[:-1][b], xi + _offsets(nx)[:-1][b]


def radius_graph(x: torch.Tensor, r: float, n: torch.Tensor, max_num_neighbors: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(src, dst) of the radius graph on one node set: grouped by ascending dst, sources ascending."""
    dst, src = radius_to(x, x, r, n, n, max_num_neighbors, drop_self=True)
    return src, dst


def _edges_per_complex(dst: torch.Tensor, dst_batch: torch.Tensor, B: int) -> torch.Tensor:
    """reference utils.py:92-98 for edges grouped by complex."""
    if dst.numel() == 0:
        return torch.zeros(B, dtype=torch.long, device=dst_batch.device)
    return torch.bincount(dst_batch[dst], minlength=B)


def _scatter_sum(msg: torch.Tensor, dst: torch.Tensor, n: int) -> torch.Tensor:
    out = torch.zeros((n,) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
    return out.index_add_(0, dst, msg)


def _scatter_mean(msg: torch.Tensor, dst: torch.Tensor, n: int) -> torch.Tensor:
    deg = torch.bincount(dst, minlength=n).clamp(min=1).to(msg.dtype)
    return _scatter_sum(msg, dst, n) / deg.view(-1, *([1] * (msg.dim() - 1)))


def _dense_attention_positions(q_src, q_dst, pos, src, dst, n_dst, scale):
    """Keypoint positions from the rec->kp "attention" exactly as the reference evaluates it
    (receptor_encoder.py:194-221, receptor_encoder_gvp.py:66-87): the per-keypoint normaliser is sum_i exp(<q_i, q_k> /
    scale), but the weight multiplied onto the positions is the RAW dot product <q_i, q_k> -- the exponentiated,
    scaled score is only a local variable there and the edge field 'a' that v_mul_e reads still holds the dot product.
    So kp = sum_i x_i <q_i, q_k> / sum_i exp(<q_i, q_k> / scale): not a convex combination, but it is what the shipped
    weights were trained under."""
    dot = (q_src[src] * q_dst[dst]).sum(-1, keepdim=True)
    denom = _scatter_sum(torch.exp(dot / scale), dst, n_dst)
    return _scatter_sum(pos[src] * (dot / denom[dst]), dst, n_dst)


def _encoded(g, kp: Dict[str, torch.Tensor], rk, kk, kp_batch: torch.Tensor, extra_rec: Dict[str, torch.Tensor] = None):
    """The encoded batch: same node sets as g, new keypoint data, rk replaced, kk added."""
    B = g.batch_size
    nd = {nt: dict(g.nodes[nt].data) for nt in g.ntypes}
    nd["kp"].update(kp)
    if extra_rec:
        nd["rec"].update(extra_rec)
    rr = g.edges(form="uv", etype="rr")
    edges = {("rec", "rr", "rec"): rr, ("rec", "rk", "kp"): rk, ("kp", "kk", "kp"): kk}
    bne = {("rec", "rr", "rec"): g.batch_num_edges("rr"),
           ("rec", "rk", "kp"): _edges_per_complex(rk[1], kp_batch, B),
           ("kp", "kk", "kp"): _edges_per_complex(kk[0], kp_batch, B)}
    out = hetero.HeteroBatch({nt: g.batch_num_nodes(nt) for nt in g.ntypes}, nd, edges, bne)
    for k, v in g.edges["rr"].data.items():
        out.edges["rr"].data[k] = v
    return out


def _batch_index(counts: torch.Tensor) -> torch.Tensor:
    return torch.arange(counts.shape[0], device=counts.device).repeat_interleave(counts)


# ------------------------------------------------------------------ EGNN-type encoder


class ReceptorConv(nn.Module):
    """One E(n)-equivariant layer over the rr edges (reference receptor_encoder.py:14-153)."""

    def __init__(self, in_size, hidden_size, out_size, edge_feat_size=0, use_tanh=True, coords_range=10, message_norm=1,
                 fix_pos: bool = False, norm: bool = False):
        super().__init__()
        self.edge_feat_size, self.use_tanh, self.coords_range, self.fix_pos = edge_feat_size, use_tanh, coords_range, fix_pos
        act = nn.SiLU()
        f_in = in_size * 2 + edge_feat_size + 1
        self.edge_mlp = nn.Sequential(nn.Linear(f_in, hidden_size), act, nn.Linear(hidden_size, hidden_size), act)
        self.node_mlp = nn.Sequential(nn.Linear(in_size + hidden_size, hidden_size), act, nn.Linear(hidden_size, out_size))
        self.soft_attention = nn.Sequential(nn.Linear(hidden_size, 1), nn.Sigmoid())
        self.layer_norm = nn.LayerNorm(out_size) if norm else nn.Identity()
        if not fix_pos:
            last = nn.Linear(hidden_size, 1, bias=False)
            nn.init.xavier_uniform_(last.weight, gain=0.001)
            self.coord_mlp = nn.Sequential(nn.Linear(f_in, hidden_size), act, last)

    def forward(self, src, dst, h, x, z, edge_feat=None):
        x_diff = x[src] - x[dst]
        radial = torch.norm(x_diff, dim=1).unsqueeze(-1)             # the distance itself
        x_diff = x_diff / (radial + 1)
        f = [h[src], h[dst], radial]
        if self.edge_feat_size > 0:
            assert edge_feat is not None, "Edge features must be provided."
            f.append(edge_feat.to(h.dtype))
        f = torch.cat(f, dim=-1)
        msg_h = self.edge_mlp(f)
        msg_h = msg_h * self.soft_attention(msg_h)
        n = h.shape[0]
        h_neigh = _scatter_sum(msg_h, dst, n) / z
        h_out = self.layer_norm(self.node_mlp(torch.cat([h, h_neigh], dim=-1)))
        if self.fix_pos:
            return h_out, x
        w = self.coord_mlp(f)
        msg_x = torch.tanh(w) * x_diff * self.coords_range if self.use_tanh else w * x_diff
        return h_out, x + _scatter_sum(msg_x, dst, n) / z


class RecKeyConv(nn.Module):
    """Attention placement of the keypoints + their features (reference receptor_encoder.py:156-300)."""

    def __init__(self, in_feats: int, out_feats: int, n_keypoints: int, num_heads: int = 1, k_closest: int = 0,
                 kp_rad: float = 0, fix_pos: bool = False, norm: bool = False):
        super().__init__()
        assert num_heads == 1
        self.out_feats, self.n_keypoints, self.k_closest, self.kp_rad, self.fix_pos = out_feats, n_keypoints, k_closest, kp_rad, fix_pos
        self.fc_src = nn.Linear(in_feats, out_feats * num_heads, bias=False)
        self.fc_dst = nn.Linear(in_feats, out_feats * num_heads, bias=False)     # never applied (see forward)
        self.kp_feature_mlp = nn.Sequential(nn.Linear(out_feats + k_closest, out_feats), nn.SiLU())
        self.layer_norm = nn.LayerNorm(out_feats) if norm else nn.Identity()

    def forward(self, g, h_rec, x_rec, h_kp):
        n_rec, n_kp = g.batch_num_nodes("rec"), g.batch_num_nodes("kp")
        N_kp = h_kp.shape[0]
        src, dst = g.edges(form="uv", etype="rk")
        # the reference projects BOTH sides with fc_src (receptor_encoder.py:190-191); fc_dst only holds parameters
        kp_pos = _dense_attention_positions(self.fc_src(h_rec), self.fc_src(h_kp),
                                            g.nodes["rec"].data["x_0"] if self.fix_pos else x_rec, src, dst, N_kp,
                                            self.out_feats ** 0.5)
        x0 = g.nodes["rec"].data["x_0"]
        if self.k_closest != 0:
            if int(n_rec.min()) < self.k_closest:
                raise ValueError("every pocket needs at least k_closest receptor atoms")
            kp_i, rec_i = knn_to(x0, kp_pos, self.k_closest, n_rec, n_kp)
            h_m = _scatter_mean(h_rec[rec_i], kp_i, N_kp)
            d = torch.norm(x0[rec_i] - kp_pos[kp_i] + 1e-30, dim=1)
            kp_feat = torch.cat([h_m, d.view(N_kp, self.k_closest)], dim=1)          # ascending distance per keypoint
        elif self.kp_rad != 0:
            kp_i, rec_i = radius_to(x0, kp_pos, self.kp_rad, n_rec, n_kp, 100)
            z = _edges_per_complex(kp_i, _batch_index(n_kp), n_kp.shape[0]) / n_kp
            kp_feat = _scatter_sum(h_rec[rec_i], kp_i, N_kp) / (z[_batch_index(n_kp)].view(-1, 1) + 1)
        else:
            raise NotImplementedError
        return kp_pos, self.layer_norm(self.kp_feature_mlp(kp_feat)), (rec_i, kp_i)


class KeyKeyConv(nn.Module):
    """Parameter holder: the reference's forward raises NotImplementedError (receptor_encoder.py:337)."""

    def __init__(self, in_feats: int, out_feats: int, num_heads: int = 1, pre_norm=False, post_norm=True):
        super().__init__()
        hs = in_feats // num_heads
        self.fc_src = nn.Linear(in_feats, hs * num_heads, bias=False)
        self.fc_dst = nn.Linear(in_feats, hs * num_heads, bias=False)
        self.val_fn = nn.Linear(in_feats, hs * num_heads, bias=False)
        self.merge_heads = nn.Linear(hs * num_heads, out_feats, bias=False)
        self.pre_norm = nn.LayerNorm(in_feats) if pre_norm else nn.Identity()
        self.post_norm = nn.LayerNorm(out_feats) if post_norm else nn.Identity()
        self.dense = nn.Sequential(nn.Linear(out_feats, out_feats * 2), nn.SiLU(), nn.Linear(out_feats * 2, out_feats), nn.SiLU())

    def forward(self, g):
        raise NotImplementedError


class ReceptorEncoder(nn.Module):

    def __init__(self, n_convs: int = 6, n_keypoints: int = 10, graph_cutoffs: dict = {}, in_n_node_feat: int = 13,
                 use_sameres_feat: bool = False, hidden_n_node_feat: int = 256, out_n_node_feat: int = 256, use_tanh=True,
                 coords_range=10, kp_feat_scale=1, message_norm=1, kp_rad: float = 0, k_closest: int = 0, norm: bool = False,
                 no_cg=False, fix_pos=False, n_kk_convs: int = 0, n_kk_heads: int = 4):
        super().__init__()
        if kp_rad != 0 and k_closest != 0:
            raise ValueError('one of kp_rad and kp_closest can be zero but not both')
        elif kp_rad == 0 and k_closest == 0:
            raise ValueError('one of kp_rad and kp_closest must be non-zero')
        if no_cg:
            raise NotImplementedError
        self.n_keypoints, self.out_n_node_feat, self.message_norm = n_keypoints, out_n_node_feat, message_norm
        self.use_sameres_feat, self.graph_cutoffs, self.n_kk_convs = use_sameres_feat, graph_cutoffs, n_kk_convs
        convs = []
        for i in range(n_convs):
            in_size = in_n_node_feat if i == 0 else hidden_n_node_feat
            out_size = out_n_node_feat if i == n_convs - 1 else hidden_n_node_feat
            convs.append(ReceptorConv(in_size, hidden_n_node_feat, out_size, edge_feat_size=int(use_sameres_feat),
                                      use_tanh=use_tanh, coords_range=coords_range, message_norm=message_norm, norm=norm,
                                      fix_pos=fix_pos))
        self.rec_convs = nn.ModuleList(convs)
        self.keypoint_embedding = nn.Sequential(nn.Linear(out_n_node_feat, out_n_node_feat * n_keypoints), nn.SiLU())
        self.rec_kp_conv = RecKeyConv(out_n_node_feat, out_n_node_feat, n_keypoints, fix_pos=fix_pos, num_heads=1,
                                      k_closest=k_closest, kp_rad=kp_rad, norm=norm)
        if n_kk_convs > 0:
            self.kk_convs = nn.ModuleList([KeyKeyConv(out_size, out_size, num_heads=n_kk_heads, pre_norm=i > 0)
                                           for i in range(n_kk_convs)])

    def forward(self, g, batch_idxs: Optional[Dict[str, torch.Tensor]] = None):
        if self.n_kk_convs > 0:
            raise NotImplementedError       # as the reference's KeyKeyConv.forward does
        x, h = g.nodes['rec'].data['x_0'], g.nodes['rec'].data['h_0']
        n_rec, n_kp = g.batch_num_nodes('rec'), g.batch_num_nodes('kp')
        rec_batch = batch_idxs['rec'] if batch_idxs is not None else _batch_index(n_rec)
        kp_batch = batch_idxs['kp'] if batch_idxs is not None else _batch_index(n_kp)
        src, dst = g.edges(form='uv', etype='rr')
        edge_feat = g.edges['rr'].data['same_res'] if self.use_sameres_feat else None
        if self.message_norm == 0:
            z = (g.batch_num_edges('rr') / n_rec)[rec_batch].view(-1, 1)
        else:
            z = self.message_norm
        for conv in self.rec_convs:
            h, x = conv(src, dst, h, x, z, edge_feat)
        mean_h = _scatter_mean(h, rec_batch, n_rec.shape[0])
        h_kp = self.keypoint_embedding(mean_h).view(-1, self.out_n_node_feat)        # 'b (k d) -> (b k) d'
        kp_pos, kp_feat, (rk_s, rk_d) = self.rec_kp_conv(g, h, x, h_kp)
        kk = radius_graph(kp_pos, self.graph_cutoffs['kk'], n_kp, 100)
        return _encoded(g, {'x_0': kp_pos, 'h_0': kp_feat}, (rk_s, rk_d), kk, kp_batch, {'x': x, 'h': h})


# ------------------------------------------------------------------ GVP-type encoder


def _norm_no_nan(x, axis=-1, keepdims=False, eps=1e-8, sqrt=True):
    out = torch.clamp(torch.sum(torch.square(x), axis, keepdims), min=eps)
    return torch.sqrt(out) if sqrt else out


def _rbf(d, d_max, count=16):
    mu = torch.linspace(0., d_max, count, device=d.device).view(1, -1)
    sigma = d_max / count
    return torch.exp(-((d.unsqueeze(-1) - mu) / sigma) ** 2)


class GVP(nn.Module):
    """Geometric vector perceptron with vector gating (reference gvp.py:43-116)."""

    def __init__(self, dim_vectors_in, dim_vectors_out, dim_feats_in, dim_feats_out, feats_activation=None,
                 vectors_activation=None):
        super().__init__()
        dim_h = max(dim_vectors_in, dim_vectors_out)
        kh, ku = 1 / math.sqrt(dim_vectors_in), 1 / math.sqrt(dim_h)
        self.Wh = nn.Parameter(torch.zeros(dim_vectors_in, dim_h).uniform_(-kh, kh))
        self.Wu = nn.Parameter(torch.zeros(dim_h, dim_vectors_out).uniform_(-ku, ku))
        self.vectors_activation = vectors_activation if vectors_activation is not None else nn.Sigmoid()
        self.to_feats_out = nn.Sequential(nn.Linear(dim_h + dim_feats_in, dim_feats_out),
                                          feats_activation if feats_activation is not None else nn.SiLU())
        self.scalar_to_vector_gates = nn.Linear(dim_feats_out, dim_vectors_out)

    def forward(self, data):
        feats, vectors = data
        Vh = torch.einsum('bvc,vh->bhc', vectors, self.Wh)
        Vu = torch.einsum('bhc,hu->buc', Vh, self.Wu)
        feats_out = self.to_feats_out(torch.cat((feats, _norm_no_nan(Vh)), dim=1))
        gate = self.scalar_to_vector_gates(feats_out).unsqueeze(-1)          # from the activated scalars
        return feats_out, self.vectors_activation(gate) * Vu


class _VDropout(nn.Module):
    def __init__(self, drop_rate):
        super().__init__()
        self.drop_rate = drop_rate
        self.dummy_param = nn.Parameter(torch.empty(0))          # zero-element parameter of the checkpoints


class GVPDropout(nn.Module):
    """Holds the reference's (eval-mode no-op) dropout sub-modules so the state_dict keys exist."""

    def __init__(self, rate):
        super().__init__()
        self.vector_dropout = _VDropout(rate)
        self.feat_dropout = nn.Dropout(rate)


class GVPLayerNorm(nn.Module):
    def __init__(self, feats_h_size, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.feat_norm = nn.LayerNorm(feats_h_size)

    def forward(self, feats, vectors):
        vn = _norm_no_nan(vectors, axis=-1, keepdims=True, sqrt=False)
        vn = torch.sqrt(torch.mean(vn, dim=-2, keepdim=True) + self.eps) + self.eps
        return self.feat_norm(feats), vectors / vn


class GVPEdgeConv(nn.Module):
    """GVP message passing over one edge type (reference gvp.py:170-341)."""

    def __init__(self, edge_type, scalar_size=128, vector_size=16, n_message_gvps=1, n_update_gvps=1, use_dst_feats=False,
                 rbf_dmax=15, rbf_dim=16, edge_feat_size=0, message_norm: Union[float, str] = 10, dropout=0.0):
        super().__init__()
        self.edge_type, self.use_dst_feats, self.rbf_dmax, self.rbf_dim = edge_type, use_dst_feats, rbf_dmax, rbf_dim
        self.edge_feat_size, self.message_norm = edge_feat_size, message_norm
        msg = []
        for i in range(n_message_gvps):
            vin, fin = vector_size, scalar_size
            if i == 0:
                vin, fin = vin + 1, fin + rbf_dim
                if use_dst_feats:
                    vin, fin = vin + vector_size, fin + scalar_size
            msg.append(GVP(vin, vector_size, fin, scalar_size))
        self.edge_message = nn.Sequential(*msg)
        self.node_update = nn.Sequential(*[GVP(vector_size, vector_size, scalar_size, scalar_size) for _ in range(n_update_gvps)])
        self.dropout = GVPDropout(dropout)
        self.message_layer_norm = GVPLayerNorm(scalar_size)
        self.update_layer_norm = GVPLayerNorm(scalar_size)

    def forward(self, src, dst, src_feats, dst_feats=None, edge_feats=None, z=1):
        if self.training:
            raise NotImplementedError("sampling only: dropout is an eval-mode no-op here")
        s_src, x_src, v_src = src_feats
        s_dst, x_dst, v_dst = src_feats if dst_feats is None else dst_feats
        x_diff = x_src[src] - x_dst[dst]
        dij = _norm_no_nan(x_diff, keepdims=True) + 1e-8
        x_diff = x_diff / dij
        vec = [x_diff.unsqueeze(1), v_src[src]]
        sca = [s_src[src], _rbf(dij.squeeze(1), self.rbf_dmax, self.rbf_dim)]
        if self.edge_feat_size > 0:
            assert edge_feats is not None, "Edge features must be provided."
            sca.append(edge_feats.to(s_src.dtype))
        if self.use_dst_feats:
            vec.append(v_dst[dst])
            sca.append(s_dst[dst])
        s_msg, v_msg = self.edge_message((torch.cat(sca, dim=1), torch.cat(vec, dim=1)))
        agg = _scatter_mean if self.message_norm == 'mean' else _scatter_sum
        n = s_dst.shape[0]
        s_msg = agg(s_msg, dst, n) / z
        v_msg = agg(v_msg, dst, n) / (z.unsqueeze(-1) if isinstance(z, torch.Tensor) else z)
        s, v = self.message_layer_norm(s_dst + s_msg, v_dst + v_msg)
        s_res, v_res = self.node_update((s, v))
        return self.update_layer_norm(s + s_res, v + v_res)


class KeypointInitializer(nn.Module):
    """Initial keypoint positions by attention over the pocket atoms (reference receptor_encoder_gvp.py:15-93)."""

    def __init__(self, n_keypoints: int, scalar_size: int, vector_size: int):
        super().__init__()
        self.scalar_size, self.vector_size, self.n_keypoints = scalar_size, vector_size, n_keypoints
        self.src_net = nn.Linear(scalar_size, scalar_size, bias=False)
        self.dst_net = nn.Linear(scalar_size, scalar_size, bias=False)
        self.keypoint_embedding = nn.Sequential(nn.Linear(scalar_size, scalar_size * n_keypoints), nn.SiLU(),
                                                nn.LayerNorm(scalar_size * n_keypoints))
        self.norm = nn.LayerNorm(scalar_size)            # unused by the reference's forward as well

    def forward(self, g, rec_scalars, rec_batch):
        B = g.batch_size
        emb = self.keypoint_embedding(_scatter_mean(rec_scalars, rec_batch, B)).view(-1, self.scalar_size)
        src, dst = g.edges(form='uv', etype='rk')
        kp_pos = _dense_attention_positions(self.src_net(rec_scalars), self.dst_net(emb), g.nodes['rec'].data['x_0'], src, dst,
                                            emb.shape[0], self.scalar_size ** 0.5)
        dev = rec_scalars.device
        return (kp_pos, torch.zeros(emb.shape[0], self.scalar_size, device=dev),
                torch.zeros(emb.shape[0], self.vector_size, 3, device=dev))


class ReceptorEncoderGVP(nn.Module):

    def __init__(self, in_scalar_size: int, out_scalar_size: int = 128, n_message_gvps: int = 1, n_update_gvps: int = 1,
                 vector_size: int = 16, n_rr_convs: int = 3, n_rk_convs: int = 2, message_norm: Union[float, str] = 10,
                 use_sameres_feat: bool = False, kp_rad: float = 0, k_closest: int = 0, dropout: float = 0.0,
                 n_keypoints: int = 20, no_cg: bool = False, graph_cutoffs: dict = {}):
        super().__init__()
        if no_cg:
            raise NotImplementedError('no_cg is not implemented yet')
        if kp_rad != 0 and k_closest != 0:
            raise ValueError('one of kp_rad and kp_closest can be zero but not both')
        elif kp_rad == 0 and k_closest == 0:
            raise ValueError('one of kp_rad and kp_closest must be non-zero')
        if (isinstance(message_norm, str) and message_norm != 'mean') or not isinstance(message_norm, (str, float, int)):
            raise ValueError(f'message norm must be either a float, int, or "mean". Got {message_norm}')
        self.vector_size, self.message_norm, self.use_sameres_feat = vector_size, message_norm, use_sameres_feat
        self.kp_rad, self.k_closest, self.graph_cutoffs = kp_rad, k_closest, graph_cutoffs
        self.scalar_embed = nn.Sequential(nn.Linear(in_scalar_size, out_scalar_size), nn.SiLU(),
                                          nn.Linear(out_scalar_size, out_scalar_size), nn.SiLU())
        self.scalar_norm = nn.LayerNorm(out_scalar_size)
        common = dict(scalar_size=out_scalar_size, vector_size=vector_size, n_message_gvps=n_message_gvps,
                      n_update_gvps=n_update_gvps, edge_feat_size=int(use_sameres_feat), dropout=dropout,
                      message_norm=message_norm)
        self.rr_conv_layers = nn.ModuleList([GVPEdgeConv(('rec', 'rr', 'rec'), rbf_dmax=graph_cutoffs['rr'], **common)
                                             for _ in range(n_rr_convs)])
        self.keypoint_initializer = KeypointInitializer(n_keypoints, out_scalar_size, vector_size)
        self.rk_conv_layers = nn.ModuleList([GVPEdgeConv(('rec', 'rk', 'kp'), use_dst_feats=i != 0,
                                                         rbf_dmax=graph_cutoffs['rk'], **common) for i in range(n_rk_convs)])

    def forward(self, g, batch_idxs: Optional[Dict[str, torch.Tensor]] = None):
        n_rec, n_kp = g.batch_num_nodes('rec'), g.batch_num_nodes('kp')
        rec_batch = batch_idxs['rec'] if batch_idxs is not None else _batch_index(n_rec)
        kp_batch = batch_idxs['kp'] if batch_idxs is not None else _batch_index(n_kp)
        x = g.nodes['rec'].data['x_0']
        s = self.scalar_norm(self.scalar_embed(g.nodes['rec'].data['h_0']))
        v = torch.zeros((x.shape[0], self.vector_size, 3), device=x.device)
        edge_feat = g.edges['rr'].data['a'] if self.use_sameres_feat else None
        if self.message_norm == 'mean':
            z = 1
        elif self.message_norm == 0:
            z = (g.batch_num_edges('rr') / n_rec)[rec_batch].view(-1, 1)
        else:
            z = self.message_norm
        src, dst = g.edges(form='uv', etype='rr')
        for conv in self.rr_conv_layers:
            s, v = conv(src, dst, (s, x, v), edge_feats=edge_feat, z=z)
        kp_pos, kp_s, kp_v = self.keypoint_initializer(g, s, rec_batch)
        if self.k_closest > 0:
            kp_i, rec_i = knn_to(x, kp_pos, self.k_closest, n_rec, n_kp)
        else:
            kp_i, rec_i = radius_to(x, kp_pos, self.kp_rad, n_rec, n_kp, 10)
        if self.message_norm == 0:
            z = (_edges_per_complex(kp_i, kp_batch, n_kp.shape[0]) / n_kp)[kp_batch].view(-1, 1)
        for conv in self.rk_conv_layers:
            kp_s, kp_v = conv(rec_i, kp_i, (s, x, v), dst_feats=(kp_s, kp_pos, kp_v), z=z)
        kk = radius_graph(kp_pos, self.graph_cutoffs['kk'], n_kp, 100)
        return _encoded(g, {'x_0': kp_pos, 'h_0': kp_s, 'v_0': kp_v}, (rec_i, kp_i), kk, kp_batch)
