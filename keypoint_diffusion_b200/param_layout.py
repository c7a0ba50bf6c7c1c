"""Parameter names / shapes / initialisers of the reference modules, and a container that
registers them under the same dotted names, so that ``state_dict()`` keys and
``load_state_dict(strict=True)`` match the reference's ``trained_models/*/model.pt`` layout
(SURVEY.md section 8b "Checkpoint").

The drop-in modules hold parameters only; their ``forward`` runs in the CUDA library.
Shapes restate what the reference constructors register:
  models/dynamics.py:15-87, :223-264, :300-339      (EGNN denoiser)
  models/dynamics_gvp.py:12-36, :58-91, :106-147    (GVP denoiser)
  models/gvp.py:44-87, :170-247, :347-437           (GVP, GVPEdgeConv, GVPMultiEdgeConv)
  models/receptor_encoder.py:17-66, :158-180, :383-482, :303-335   (EGNN receptor encoder)
  models/receptor_encoder_gvp.py:19-37, :99-205                    (GVP receptor encoder)
tests/test_modules_cpu.py checks every shipped config against the inventory the reference's own
model_from_config produces (tests/golden/state_dict_shapes.json).
"""
import math
from collections import OrderedDict
from typing import Dict, Tuple

import torch
import torch.nn as nn

Shapes = "OrderedDict[str, Tuple[Tuple[int, ...], str]]"   # name -> (shape, init kind)


def _lin(d, name, out_f, in_f, bias=True, init="linear"):
    d[name + ".weight"] = ((out_f, in_f), init)
    if bias:
        d[name + ".bias"] = ((out_f,), f"bias:{in_f}")


def _ln(d, name, n):
    d[name + ".weight"] = ((n,), "ones")
    d[name + ".bias"] = ((n,), "zeros")


def _gvp(d, name, vin, vout, fin, fout):
    h = max(vin, vout)
    d[name + ".Wh"] = ((vin, h), "gvp_w")
    d[name + ".Wu"] = ((h, vout), "gvp_w")
    _lin(d, name + ".to_feats_out.0", fout, h + fin)
    _lin(d, name + ".scalar_to_vector_gates", vout, fout)


def egnn_dynamics_shapes(atom_nf, rec_nf, n_layers, hidden_nf, update_kp_feat, norm):
    d = OrderedDict()
    _lin(d, "lig_encoder.0", 64, atom_nf)
    _lin(d, "lig_encoder.2", hidden_nf, 64)
    _lin(d, "lig_decoder.0", 2 * atom_nf, hidden_nf)
    _lin(d, "lig_decoder.2", atom_nf, 2 * atom_nf)
    if rec_nf != hidden_nf:
        _lin(d, "rec_encoder.0", 2 * rec_nf, rec_nf)
        _lin(d, "rec_encoder.2", hidden_nf, 2 * rec_nf)
    H = hidden_nf + 1
    etypes = ["ll", "kl", "lk", "kk"] if update_kp_feat else ["ll", "kl"]
    ntypes = ["lig", "kp"] if update_kp_feat else ["lig"]
    for l in range(n_layers):
        q = f"egnn.conv_layers.{l}."
        for et in etypes:
            _lin(d, f"{q}edge_mlp.{et}.0", H, 2 * H + 1)
            _lin(d, f"{q}edge_mlp.{et}.2", H, H)
        for et in etypes:
            _lin(d, f"{q}soft_attention.{et}.0", 1, H)
        for nt in ntypes:
            _lin(d, f"{q}node_mlp.{nt}.0", H, 2 * H)
            _lin(d, f"{q}node_mlp.{nt}.2", H, H)
        for et in etypes:
            _lin(d, f"{q}coord_mlp.{et}.0", H, 2 * H + 1)
            _lin(d, f"{q}coord_mlp.{et}.2", H, H)
            _lin(d, f"{q}coord_mlp.{et}.4", 1, H, bias=False, init="xavier_small")
        if norm:
            for nt in ntypes:
                _ln(d, f"{q}layer_norm.{nt}", H)
    return d


def gvp_layer_etypes(l, n_convs, update_kp):
    base = [("lig", "ll", "lig"), ("kp", "kl", "lig")]
    if (not update_kp) or l == n_convs - 1:
        return base
    return base + [("lig", "lk", "kp"), ("kp", "kk", "kp")]


def gvp_dynamics_shapes(n_lig_scalars, n_kp_scalars, vector_size, n_convs, n_hidden_scalars, update_kp,
                        n_message_gvps, n_update_gvps, n_noise_gvps, rbf_dim=16):
    d = OrderedDict()
    S, V = n_hidden_scalars, vector_size
    _lin(d, "lig_encoder.0", S, n_lig_scalars + 1)
    _ln(d, "lig_encoder.2", S)
    _lin(d, "kp_encoder.0", S, n_kp_scalars + 1)
    _ln(d, "kp_encoder.2", S)
    for l in range(n_convs):
        etypes = gvp_layer_etypes(l, n_convs, update_kp)
        dst = sorted(set(e[2] for e in etypes))
        q = f"noise_predictor.conv_layers.{l}."
        for et in etypes:
            for i in range(n_message_gvps):
                _gvp(d, f"{q}edge_message_fns.{'_'.join(et)}.{i}", V + 1 if i == 0 else V, V,
                     S + rbf_dim if i == 0 else S, S)
        for nt in dst:
            for i in range(n_update_gvps):
                _gvp(d, f"{q}node_update_fns.{nt}.{i}", V, V, S, S)
        for nt in dst:
            _ln(d, f"{q}update_layer_norms.{nt}.feat_norm", S)
        for nt in dst:
            _ln(d, f"{q}message_layer_norms.{nt}.feat_norm", S)
        d[f"{q}dropout.vector_dropout.dummy_param"] = ((0,), "empty")
    q = "noise_predictor.noise_predictor."
    for i in range(n_noise_gvps):
        last = i == n_noise_gvps - 1
        _gvp(d, f"{q}gvps.{i}", V, 1 if last else V, S, 64 if last else S)
    _lin(d, q + "to_scalar_output", n_lig_scalars, 64)
    return d


def egnn_rec_encoder_shapes(n_convs=6, n_keypoints=10, in_n_node_feat=13, use_sameres_feat=False,
                            hidden_n_node_feat=256, out_n_node_feat=256, k_closest=0, norm=False, fix_pos=False,
                            n_kk_convs=0, n_kk_heads=4, **_unused):
    d = OrderedDict()
    ef = 1 if use_sameres_feat else 0
    out_size = out_n_node_feat
    for i in range(n_convs):
        in_size = in_n_node_feat if i == 0 else hidden_n_node_feat
        out_size = out_n_node_feat if i == n_convs - 1 else hidden_n_node_feat
        hid = hidden_n_node_feat
        q = f"rec_convs.{i}."
        _lin(d, q + "edge_mlp.0", hid, 2 * in_size + ef + 1)
        _lin(d, q + "edge_mlp.2", hid, hid)
        _lin(d, q + "node_mlp.0", hid, in_size + hid)
        _lin(d, q + "node_mlp.2", out_size, hid)
        _lin(d, q + "soft_attention.0", 1, hid)
        if norm:
            _ln(d, q + "layer_norm", out_size)
        if not fix_pos:
            _lin(d, q + "coord_mlp.0", hid, 2 * in_size + ef + 1)
            _lin(d, q + "coord_mlp.2", 1, hid, bias=False, init="xavier_small")
    o = out_n_node_feat
    _lin(d, "keypoint_embedding.0", o * n_keypoints, o)
    _lin(d, "rec_kp_conv.fc_src", o, o, bias=False)
    _lin(d, "rec_kp_conv.fc_dst", o, o, bias=False)
    _lin(d, "rec_kp_conv.kp_feature_mlp.0", o, o + k_closest)
    if norm:
        _ln(d, "rec_kp_conv.layer_norm", o)
    for c in range(n_kk_convs):
        q = f"kk_convs.{c}."
        hs = out_size // n_kk_heads * n_kk_heads
        for nm in ("fc_src", "fc_dst", "val_fn"):
            _lin(d, q + nm, hs, out_size, bias=False)
        _lin(d, q + "merge_heads", out_size, hs, bias=False)
        if c > 0:
            _ln(d, q + "pre_norm", out_size)
        _ln(d, q + "post_norm", out_size)
        _lin(d, q + "dense.0", 2 * out_size, out_size)
        _lin(d, q + "dense.2", out_size, 2 * out_size)
    return d


def _gvp_edge_conv(d, q, S, V, n_msg, n_upd, use_dst_feats, edge_feat_size, rbf_dim=16):
    for i in range(n_msg):
        vin, fin = V, S
        if i == 0:
            vin += 1
            fin += rbf_dim
            if use_dst_feats:
                vin += V
                fin += S
        # NB the reference does not widen the first message GVP for edge features
        # (models/gvp.py:205-213), so edge_feat_size does not enter the shapes
        _gvp(d, f"{q}edge_message.{i}", vin, V, fin, S)
    for i in range(n_upd):
        _gvp(d, f"{q}node_update.{i}", V, V, S, S)
    d[q + "dropout.vector_dropout.dummy_param"] = ((0,), "empty")
    _ln(d, q + "message_layer_norm.feat_norm", S)
    _ln(d, q + "update_layer_norm.feat_norm", S)


def gvp_rec_encoder_shapes(in_scalar_size, out_scalar_size=128, n_message_gvps=1, n_update_gvps=1, vector_size=16,
                           n_rr_convs=3, n_rk_convs=2, use_sameres_feat=False, n_keypoints=20, **_unused):
    d = OrderedDict()
    S, V = out_scalar_size, vector_size
    _lin(d, "scalar_embed.0", S, in_scalar_size)
    _lin(d, "scalar_embed.2", S, S)
    _ln(d, "scalar_norm", S)
    ef = 1 if use_sameres_feat else 0
    for i in range(n_rr_convs):
        _gvp_edge_conv(d, f"rr_conv_layers.{i}.", S, V, n_message_gvps, n_update_gvps, False, ef)
    _lin(d, "keypoint_initializer.src_net", S, S, bias=False)
    _lin(d, "keypoint_initializer.dst_net", S, S, bias=False)
    _lin(d, "keypoint_initializer.keypoint_embedding.0", S * n_keypoints, S)
    _ln(d, "keypoint_initializer.keypoint_embedding.2", S * n_keypoints)
    _ln(d, "keypoint_initializer.norm", S)
    for i in range(n_rk_convs):
        _gvp_edge_conv(d, f"rk_conv_layers.{i}.", S, V, n_message_gvps, n_update_gvps, i != 0, ef)
    return d


def _init(shape, kind, gen):
    if kind == "empty" or (len(shape) == 1 and shape[0] == 0):
        return torch.empty(shape)
    if kind == "ones":
        return torch.ones(shape)
    if kind == "zeros":
        return torch.zeros(shape)
    if kind == "gvp_w":                       # models/gvp.py:64-67
        k = 1.0 / math.sqrt(shape[0])
        return (torch.rand(shape, generator=gen) * 2 - 1) * k
    if kind == "xavier_small":                # xavier_uniform_(gain=0.001), models/dynamics.py:70
        bound = 0.001 * math.sqrt(6.0 / (shape[0] + shape[1]))
        return (torch.rand(shape, generator=gen) * 2 - 1) * bound
    if kind == "linear":                      # nn.Linear default: kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in))
        k = 1.0 / math.sqrt(shape[1])
        return (torch.rand(shape, generator=gen) * 2 - 1) * k
    if kind.startswith("bias:"):
        k = 1.0 / math.sqrt(int(kind.split(":")[1]))
        return (torch.rand(shape, generator=gen) * 2 - 1) * k
    raise ValueError(kind)


class ParamTree(nn.Module):
    """Registers parameters under dotted names as nested sub-modules, so that the owning
    module's state_dict has exactly the reference's keys."""

    def __init__(self, shapes=None, generator=None):
        super().__init__()
        if shapes:
            for name, (shape, kind) in shapes.items():
                self.add(name, _init(tuple(shape), kind, generator))

    def add(self, dotted: str, value: torch.Tensor):
        head, _, rest = dotted.partition(".")
        if not rest:
            self.register_parameter(head, nn.Parameter(value, requires_grad=False))
            return
        if head not in self._modules:
            self.add_module(head, ParamTree())
        self._modules[head].add(rest, value)

    def flat(self) -> Dict[str, torch.Tensor]:
        return {k: v for k, v in self.state_dict().items()}
