"""HeteroBatch: a DGL-free container exposing the small DGL surface the sampling hot path
touches: batched heterograph with node types rec / kp / lig, per-type
node data, the static kk edge list, batch_num_nodes / batch_num_edges, local_scope, to().

It is plumbing around torch tensors: no message passing lives here (that is the CUDA library).
Anything that duck-types the same accessors -- including a real dgl.DGLHeteroGraph -- can be
passed to the drop-in modules instead.

Reference schema: data_processing/pdbbind_processing.py:236-243 (node / edge types),
utils.py:81-170 (batch helpers), models/ligand_diffuser.py:462-469 (unbatch at the end).
"""
import contextlib
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .synthetic import EncodedPocket

import weakref

NTYPES = ["kp", "lig", "rec"]
# device copies of (immutable) edge-index tensors, keyed by the identity of the host tensor: repeated
# uploads of the same static kk graph reuse one device tensor, so downstream caches (CSR, captured
# sampler) hit.  (id-keyed dict + weakref finalizer: tensors cannot be WeakKeyDictionary keys because
# their == is element-wise.)
_EDGE_DEVICE_CACHE = {}


def _edge_to(t, device):
    if t.device == torch.device(device):
        return t
    key = (id(t), str(device))
    hit = _EDGE_DEVICE_CACHE.get(key)
    if hit is not None and hit[0]() is t:
        return hit[1]
    d = t.to(device)
    _EDGE_DEVICE_CACHE[key] = (weakref.ref(t, lambda _r, k=key: _EDGE_DEVICE_CACHE.pop(k, None)), d)
    return d


CANONICAL_ETYPES = [("kp", "kk", "kp"), ("kp", "kl", "lig"), ("lig", "lk", "kp"), ("lig", "ll", "lig"),
                    ("rec", "rk", "kp"), ("rec", "rr", "rec")]


class _Space:
    def __init__(self, data):
        self.data = data


class _Nodes:
    def __init__(self, g):
        self._g = g

    def __getitem__(self, nt):
        return _Space(self._g._ndata[nt])

    def __call__(self, ntype=None):
        return torch.arange(self._g.num_nodes(ntype), device=self._g.device)


class _Edges:
    def __init__(self, g):
        self._g = g

    def __getitem__(self, et):
        return _Space(self._g._edata.setdefault(self._g.to_canonical_etype(et), {}))

    def __call__(self, form="uv", etype=None):
        s, d = self._g._edges[self._g.to_canonical_etype(etype)]
        if form == "uv":
            return s, d
        if form == "eid":
            return torch.arange(s.shape[0], device=s.device)
        raise ValueError(form)


class HeteroBatch:
    def __init__(self, batch_num_nodes: Dict[str, torch.Tensor], ndata: Dict[str, Dict[str, torch.Tensor]],
                 edges: Optional[Dict[Tuple[str, str, str], Tuple[torch.Tensor, torch.Tensor]]] = None,
                 batch_num_edges: Optional[Dict[Tuple[str, str, str], torch.Tensor]] = None):
        self._bnn = {nt: batch_num_nodes.get(nt, torch.zeros_like(next(iter(batch_num_nodes.values()))))
                     for nt in NTYPES}
        self._ndata = {nt: dict(ndata.get(nt, {})) for nt in NTYPES}
        dev = self.device
        empty = torch.zeros(0, dtype=torch.long, device=dev)
        self._edges = {et: (empty, empty) for et in CANONICAL_ETYPES}
        if edges:
            for et, (s, d) in edges.items():
                self._edges[self.to_canonical_etype(et)] = (s.long(), d.long())
        B = self.batch_size
        self._bne = {et: torch.zeros(B, dtype=torch.long, device=dev) for et in CANONICAL_ETYPES}
        if batch_num_edges:
            for et, v in batch_num_edges.items():
                self._bne[self.to_canonical_etype(et)] = v
        self._edata = {}

    # ---- schema
    ntypes = NTYPES
    canonical_etypes = CANONICAL_ETYPES

    @staticmethod
    def to_canonical_etype(et):
        if isinstance(et, tuple):
            return et
        for c in CANONICAL_ETYPES:
            if c[1] == et:
                return c
        raise KeyError(et)

    @property
    def device(self):
        for st in self._ndata.values():
            for v in st.values():
                return v.device
        return next(iter(self._bnn.values())).device

    @property
    def batch_size(self):
        return int(self._bnn["kp"].shape[0])

    def batch_num_nodes(self, ntype=None):
        return self._bnn[ntype]

    def batch_num_edges(self, etype=None):
        return self._bne[self.to_canonical_etype(etype)]

    def set_batch_num_nodes(self, val):
        self._bnn.update(val)

    def set_batch_num_edges(self, val):
        for et, v in val.items():
            self._bne[self.to_canonical_etype(et)] = v

    def num_nodes(self, ntype=None):
        return int(self._bnn[ntype].sum())

    def num_edges(self, etype=None):
        return int(self._edges[self.to_canonical_etype(etype)][0].shape[0])

    @property
    def nodes(self):
        return _Nodes(self)

    @property
    def edges(self):
        return _Edges(self)

    @contextlib.contextmanager
    def local_scope(self):
        nd = {nt: dict(st) for nt, st in self._ndata.items()}
        try:
            yield
        finally:
            self._ndata = nd

    def to(self, device):
        g = HeteroBatch({k: v.to(device) for k, v in self._bnn.items()},
                        {nt: {k: v.to(device, non_blocking=True) for k, v in st.items()} for nt, st in self._ndata.items()},
                        {et: (_edge_to(s, device), _edge_to(d, device)) for et, (s, d) in self._edges.items()},
                        {et: v.to(device) for et, v in self._bne.items()})
        g._edata = {et: {k: v.to(device) for k, v in st.items()} for et, st in self._edata.items()}
        return g

    # ---- construction helpers
    @staticmethod
    def from_pockets(pockets: Sequence[EncodedPocket], n_lig_atoms: Sequence[int], atom_nf: int,
                     device="cpu", pin=False) -> "HeteroBatch":
        """One complex per entry of n_lig_atoms; complex i uses pockets[i % len(pockets)].  Ligand data
        is zero-filled, as utils.copy_graph does (reference utils.py:142-144)."""
        B = len(n_lig_atoms)
        ks, kd, kx, kh, kv, off = [], [], [], [], [], 0
        kp_n, kk_n = [], []
        for i in range(B):
            pk = pockets[i % len(pockets)]
            kx.append(pk.kp_x); kh.append(pk.kp_h)
            if pk.kp_v is not None:
                kv.append(pk.kp_v)
            ks.append(pk.kk_src + off); kd.append(pk.kk_dst + off)
            off += pk.n_kp
            kp_n.append(pk.n_kp); kk_n.append(int(pk.kk_src.numel()))
        N_l = int(sum(n_lig_atoms))
        nd = {"kp": {"x_0": torch.cat(kx), "h_0": torch.cat(kh)},
              "lig": {"x_0": torch.zeros(N_l, 3), "h_0": torch.zeros(N_l, atom_nf)},
              "rec": {"x_0": torch.zeros(0, 3), "h_0": torch.zeros(0, 1)}}
        if kv:
            nd["kp"]["v_0"] = torch.cat(kv)
        if pin and torch.cuda.is_available():
            nd = {nt: {k: v.pin_memory() for k, v in st.items()} for nt, st in nd.items()}
        g = HeteroBatch({"kp": torch.tensor(kp_n), "lig": torch.tensor(list(n_lig_atoms)), "rec": torch.zeros(B, dtype=torch.long)},
                        nd, {("kp", "kk", "kp"): (torch.cat(ks), torch.cat(kd))},
                        {("kp", "kk", "kp"): torch.tensor(kk_n)})
        return g.to(device) if str(device) != "cpu" else g


def batch(graphs: List[HeteroBatch]) -> HeteroBatch:
    bnn = {nt: torch.cat([g.batch_num_nodes(nt) for g in graphs]) for nt in NTYPES}
    nd = {}
    for nt in NTYPES:
        keys = graphs[0]._ndata[nt].keys()
        nd[nt] = {k: torch.cat([g._ndata[nt][k] for g in graphs]) for k in keys}
    edges, bne = {}, {}
    for et in CANONICAL_ETYPES:
        so = do = 0
        ss, dd = [], []
        for g in graphs:
            s, d = g._edges[et]
            ss.append(s + so); dd.append(d + do)
            so += g.num_nodes(et[0]); do += g.num_nodes(et[2])
        edges[et] = (torch.cat(ss), torch.cat(dd))
        bne[et] = torch.cat([g.batch_num_edges(et) for g in graphs])
    out = HeteroBatch(bnn, nd, edges, bne)
    for et, st in graphs[0]._edata.items():
        out._edata[et] = {k: torch.cat([g._edata[et][k] for g in graphs]) for k in st}
    return out


def expand_complexes(enc: HeteroBatch, pocket_idx: torch.Tensor, n_lig_atoms: torch.Tensor, atom_nf: int = None) -> HeteroBatch:
    """One batched graph with a complex per entry of pocket_idx: complex c = the encoded pocket enc[pocket_idx[c]]
    (its keypoint data and kk edges) with n_lig_atoms[c] zero-filled ligand atoms.

    The device-side replacement of the reference's per-complex host loop -- utils.copy_graph (utils.py:103-156) called
    once per receptor and dgl.batch per diffusion batch (ligand_diffuser.py:292-313): there every copy deep-clones
    every tensor of its pocket on the host (seconds at thousands of complexes); here the whole batch is a handful of
    index computations (cumsum / repeat_interleave / gather) on the device the encoded pockets live on, with no Python
    loop over complexes.  'rec' nodes are not carried over (nothing behind the encoder reads them; callers pass
    init_lig_pos)."""
    dev = enc.device
    pocket_idx = pocket_idx.to(dev, torch.long)
    n_lig = n_lig_atoms.to(dev, torch.long)
    C = int(pocket_idx.numel())
    kp_n_p = enc.batch_num_nodes("kp").to(dev, torch.long)
    kp_off_p = torch.cumsum(kp_n_p, 0) - kp_n_p
    kp_n_c = kp_n_p[pocket_idx]
    kp_off_c = torch.cumsum(kp_n_c, 0) - kp_n_c
    K = int(kp_n_c.sum())
    # row gather: complex c takes rows kp_off_p[p_c] + [0, kp_n_c[c])
    shift = torch.repeat_interleave(kp_off_p[pocket_idx] - kp_off_c, kp_n_c, output_size=K)
    rows = shift + torch.arange(K, device=dev)
    kp_data = {k: v.index_select(0, rows) for k, v in enc.nodes["kp"].data.items()}
    # kk edges: complex c takes the edges of pocket p_c, renumbered from the pocket's rows to the complex's rows
    ks, kd = enc.edges(form="uv", etype="kk")
    e_n_p = enc.batch_num_edges("kk").to(dev, torch.long)
    e_off_p = torch.cumsum(e_n_p, 0) - e_n_p
    e_n_c = e_n_p[pocket_idx]
    e_off_c = torch.cumsum(e_n_c, 0) - e_n_c
    E = int(e_n_c.sum())
    eshift = torch.repeat_interleave(e_off_p[pocket_idx] - e_off_c, e_n_c, output_size=E)
    eidx = eshift + torch.arange(E, device=dev)
    renum = torch.repeat_interleave(kp_off_c - kp_off_p[pocket_idx], e_n_c, output_size=E)
    new_s, new_d = ks.index_select(0, eidx) + renum, kd.index_select(0, eidx) + renum
    N_l = int(n_lig.sum())
    lig_ref = enc.nodes["lig"].data
    F = atom_nf if atom_nf is not None else (lig_ref["h_0"].shape[1] if "h_0" in lig_ref else 0)
    nd = {"kp": kp_data,
          "lig": {"x_0": torch.zeros(N_l, 3, device=dev), "h_0": torch.zeros(N_l, F, device=dev)},     # utils.py:142-144
          "rec": {"x_0": torch.zeros(0, 3, device=dev), "h_0": torch.zeros(0, 1, device=dev)}}
    return HeteroBatch({"kp": kp_n_c, "lig": n_lig, "rec": torch.zeros(C, dtype=torch.long, device=dev)}, nd,
                       {("kp", "kk", "kp"): (new_s, new_d)}, {("kp", "kk", "kp"): e_n_c})


def unbatch(g: HeteroBatch) -> List[HeteroBatch]:
    B = g.batch_size
    out = []
    noff = {nt: 0 for nt in NTYPES}
    eoff = {et: 0 for et in CANONICAL_ETYPES}
    for b in range(B):
        nn_ = {nt: int(g.batch_num_nodes(nt)[b]) for nt in NTYPES}
        nd = {nt: {k: v[noff[nt]:noff[nt] + nn_[nt]] for k, v in g._ndata[nt].items()} for nt in NTYPES}
        edges, bne = {}, {}
        for et in CANONICAL_ETYPES:
            ne = int(g.batch_num_edges(et)[b])
            s, d = g._edges[et]
            edges[et] = (s[eoff[et]:eoff[et] + ne] - noff[et[0]], d[eoff[et]:eoff[et] + ne] - noff[et[2]])
            bne[et] = torch.tensor([ne], device=s.device)
        one = HeteroBatch({nt: torch.tensor([nn_[nt]], device=g.device) for nt in NTYPES}, nd, edges, bne)
        for et, st in g._edata.items():
            ne = int(g.batch_num_edges(et)[b])
            one._edata[et] = {k: v[eoff[et]:eoff[et] + ne] for k, v in st.items()}
        for et in CANONICAL_ETYPES:
            eoff[et] += int(g.batch_num_edges(et)[b])
        out.append(one)
        for nt in NTYPES:
            noff[nt] += nn_[nt]
    return out


def readout_nodes(g, feat, op="mean", ntype=None):
    """dgl.readout_nodes for op in {sum, mean}; small torch plumbing used only outside the loop."""
    x = g.nodes[ntype].data[feat]
    counts = g.batch_num_nodes(ntype).to(x.device)
    B = counts.shape[0]
    idx = torch.arange(B, device=x.device).repeat_interleave(counts)
    out = torch.zeros((B,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    out.index_add_(0, idx, x)
    if op == "mean":
        out = out / counts.to(x.dtype).view(-1, *([1] * (x.dim() - 1)))
    return out


def build_initial_complex_graph(rec_atom_positions: torch.Tensor, rec_atom_features: torch.Tensor, pocket_res_idx: torch.Tensor,
                                n_keypoints: int, cutoffs: dict, lig_atom_positions: torch.Tensor = None,
                                lig_atom_features: torch.Tensor = None) -> HeteroBatch:
    """The raw (un-encoded) graph of one complex, as the reference's data pipeline builds it
    (data_processing/pdbbind_processing.py:221-274): rr = radius graph (cutoffs['rr'], at most 100 neighbours) with the
    `same_res` edge flag, rk = every pocket atom -> every keypoint placeholder, the other edge types empty."""
    from .receptor_encoder import radius_graph
    if (lig_atom_positions is not None) ^ (lig_atom_features is not None):
        raise ValueError('ligand position and features must be either be both supplied or both left as None')
    dev = rec_atom_positions.device
    n_rec = rec_atom_positions.shape[0]
    n_lig = 0 if lig_atom_positions is None else lig_atom_positions.shape[0]
    cnt = lambda n: torch.tensor([n], dtype=torch.long, device=dev)
    rr_s, rr_d = radius_graph(rec_atom_positions.float(), cutoffs['rr'], cnt(n_rec), 100)
    rk_s = torch.arange(n_rec, device=dev).repeat(n_keypoints)
    rk_d = torch.arange(n_keypoints, device=dev).repeat_interleave(n_rec)
    nd = {"rec": {"x_0": rec_atom_positions, "h_0": rec_atom_features}, "kp": {}, "lig": {}}
    if lig_atom_positions is not None:
        nd["lig"] = {"x_0": lig_atom_positions, "h_0": lig_atom_features}
    g = HeteroBatch({"rec": cnt(n_rec), "kp": cnt(n_keypoints), "lig": cnt(n_lig)}, nd,
                    {("rec", "rr", "rec"): (rr_s, rr_d), ("rec", "rk", "kp"): (rk_s, rk_d)},
                    {("rec", "rr", "rec"): cnt(rr_s.shape[0]), ("rec", "rk", "kp"): cnt(rk_s.shape[0])})
    g.edges["rr"].data["same_res"] = (pocket_res_idx[rr_s] == pocket_res_idx[rr_d]).view(-1, 1)
    return g
