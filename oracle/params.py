"""Parameter layout of the reference's dynamics modules + seeded init (oracle).

The key names / shapes below restate what the reference's constructors register
(models/dynamics.py:15-87, :223-264, :300-339; models/dynamics_gvp.py:12-36,
:58-91, :106-147; models/gvp.py:44-87, :347-437).  tests/test_oracle_vs_reference.py
checks them against the reference's own ``model_from_config(...).state_dict()``
for all eight shipped configs when /root/reference is present, and
tests/golden/state_dict_shapes.json holds the same comparison as a fixture.

The checkpoints themselves are missing from the reference tree
(.MISSING_LARGE_BLOBS), so parity runs on seeded random weights in this layout.

Test infrastructure only (see oracle/__init__.py).
"""
from collections import OrderedDict
import math
import torch

EGNN_ETYPES_KP = ["ll", "kl", "lk", "kk"]
EGNN_ETYPES_NOKP = ["ll", "kl"]
GVP_ETYPES_NOKP = [("lig", "ll", "lig"), ("kp", "kl", "lig")]
GVP_ETYPES_KP = GVP_ETYPES_NOKP + [("lig", "lk", "kp"), ("kp", "kk", "kp")]


def _lin(d, name, out_f, in_f, bias=True):
    d[name + ".weight"] = (out_f, in_f)
    if bias:
        d[name + ".bias"] = (out_f,)


def egnn_dynamics_shapes(atom_nf, rec_nf, n_layers=4, hidden_nf=255, update_kp_feat=False,
                         norm=False, prefix="dynamics."):
    """models/dynamics.py:300-339 (LigRecDynamics), :223-264 (LigRecEGNN), :15-87 (LigRecConv)."""
    d = OrderedDict()
    p = prefix
    _lin(d, p + "lig_encoder.0", 64, atom_nf)
    _lin(d, p + "lig_encoder.2", hidden_nf, 64)
    _lin(d, p + "lig_decoder.0", 2 * atom_nf, hidden_nf)
    _lin(d, p + "lig_decoder.2", atom_nf, 2 * atom_nf)
    if rec_nf != hidden_nf:
        _lin(d, p + "rec_encoder.0", 2 * rec_nf, rec_nf)
        _lin(d, p + "rec_encoder.2", hidden_nf, 2 * rec_nf)
    H = hidden_nf + 1
    etypes = EGNN_ETYPES_KP if update_kp_feat else EGNN_ETYPES_NOKP
    ntypes = ["lig", "kp"] if update_kp_feat else ["lig"]
    for l in range(n_layers):
        q = f"{p}egnn.conv_layers.{l}."
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
            _lin(d, f"{q}coord_mlp.{et}.4", 1, H, bias=False)
        if norm:
            for nt in ntypes:
                d[f"{q}layer_norm.{nt}.weight"] = (H,)
                d[f"{q}layer_norm.{nt}.bias"] = (H,)
    return d


def _gvp(d, name, vin, vout, fin, fout):
    """models/gvp.py:44-87"""
    h = max(vin, vout)
    d[name + ".Wh"] = (vin, h)
    d[name + ".Wu"] = (h, vout)
    _lin(d, name + ".to_feats_out.0", fout, h + fin)
    _lin(d, name + ".scalar_to_vector_gates", vout, fout)


def gvp_dynamics_shapes(n_lig_scalars, n_kp_scalars, vector_size=16, n_convs=4, n_hidden_scalars=128,
                        update_kp=False, n_message_gvps=3, n_update_gvps=2, n_noise_gvps=3,
                        rbf_dim=16, prefix="dynamics."):
    """models/dynamics_gvp.py:106-147, :58-91, :12-36; models/gvp.py:347-437."""
    d = OrderedDict()
    p = prefix
    S, V = n_hidden_scalars, vector_size
    _lin(d, p + "lig_encoder.0", S, n_lig_scalars + 1)
    d[p + "lig_encoder.2.weight"] = (S,)
    d[p + "lig_encoder.2.bias"] = (S,)
    _lin(d, p + "kp_encoder.0", S, n_kp_scalars + 1)
    d[p + "kp_encoder.2.weight"] = (S,)
    d[p + "kp_encoder.2.bias"] = (S,)
    for l in range(n_convs):
        etypes = gvp_layer_etypes(l, n_convs, update_kp)
        dst_ntypes = sorted(set(e[2] for e in etypes))
        q = f"{p}noise_predictor.conv_layers.{l}."
        for et in etypes:
            key = "_".join(et)
            for i in range(n_message_gvps):
                vin = V + 1 if i == 0 else V
                fin = S + rbf_dim if i == 0 else S
                _gvp(d, f"{q}edge_message_fns.{key}.{i}", vin, V, fin, S)
        for nt in dst_ntypes:
            for i in range(n_update_gvps):
                _gvp(d, f"{q}node_update_fns.{nt}.{i}", V, V, S, S)
        for nt in dst_ntypes:
            d[f"{q}update_layer_norms.{nt}.feat_norm.weight"] = (S,)
            d[f"{q}update_layer_norms.{nt}.feat_norm.bias"] = (S,)
        for nt in dst_ntypes:
            d[f"{q}message_layer_norms.{nt}.feat_norm.weight"] = (S,)
            d[f"{q}message_layer_norms.{nt}.feat_norm.bias"] = (S,)
        d[f"{q}dropout.vector_dropout.dummy_param"] = (0,)
    q = f"{p}noise_predictor.noise_predictor."
    for i in range(n_noise_gvps):
        last = i == n_noise_gvps - 1
        _gvp(d, f"{q}gvps.{i}", V, 1 if last else V, S, 64 if last else S)
    _lin(d, q + "to_scalar_output", n_lig_scalars, 64)
    return d


def gvp_layer_etypes(l, n_convs, update_kp):
    """models/dynamics_gvp.py:65-74: the last conv of an update_kp model is lig-only."""
    if (not update_kp) or l == n_convs - 1:
        return list(GVP_ETYPES_NOKP)
    return list(GVP_ETYPES_KP)


def init_state_dict(shapes, seed=0, coord_gain=1.0, dtype=torch.float32):
    """Seeded random weights in the reference layout.

    Linear-like tensors: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (PyTorch's Linear default
    range); LayerNorm: weight 1 + 0.1 N(0,1), bias 0.1 N(0,1) so that the affine part is
    exercised; GVP Wh/Wu: U(+-1/sqrt(rows)) (models/gvp.py:64-69).
    The reference initialises the last coord layer with xavier_uniform_(gain=0.001)
    (models/dynamics.py:69-70), which would make eps_x ~ 0 on random weights and leave
    the coordinate path untested (SURVEY note N5); ``coord_gain`` scales that layer
    relative to the ordinary Linear range instead.
    """
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, shape in shapes.items():
        if len(shape) == 1 and shape[0] == 0:
            sd[name] = torch.empty(0, dtype=dtype)
            continue
        if len(shape) == 1 and _is_layernorm(name, shapes):
            if name.endswith("weight"):
                sd[name] = (1.0 + 0.1 * torch.randn(shape, generator=g)).to(dtype)
            else:
                sd[name] = (0.1 * torch.randn(shape, generator=g)).to(dtype)
            continue
        if name.endswith(".Wh") or name.endswith(".Wu"):
            k = 1.0 / math.sqrt(shape[0])
        elif len(shape) == 2:
            k = 1.0 / math.sqrt(shape[1])
        else:  # bias: fan_in unknown here; use the matching weight's fan_in if present
            w = sd.get(name[:-4] + "weight")
            k = 1.0 / math.sqrt(w.shape[1]) if w is not None and w.dim() == 2 else 0.05
        t = (torch.rand(shape, generator=g) * 2 - 1) * k
        if ".coord_mlp." in name and name.endswith(".4.weight"):
            t = t * coord_gain
        sd[name] = t.to(dtype)
    return sd


def _is_layernorm(name, shapes):
    """A 1-D tensor belongs to a LayerNorm iff its sibling ``.weight`` is 1-D too
    (Linear weights are 2-D)."""
    stem = name.rsplit(".", 1)[0]
    w = shapes.get(stem + ".weight")
    return w is not None and len(w) == 1
