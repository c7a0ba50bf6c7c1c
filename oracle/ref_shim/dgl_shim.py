"""A small pure-PyTorch stand-in for the DGL subset the reference's hot path calls.

Purpose: let the reference's own ``nn.Module.forward`` code (imported read-only from
/root/reference) execute unchanged on CPU in the build container, where DGL is not
installed, so that oracle/flat.py can be pinned against it and golden vectors generated
(tests/golden/make_golden.py).  It is *not* DGL: semantics are restated from DGL's
documented behaviour, including one that matters for parity --

  ``g.ndata[key]`` on a graph with more than one node type returns a NEW dict
  {ntype: tensor} on every access (dgl/view.py HeteroNodeDataView.__getitem__), so
  ``g.ndata[key][ntype] = value`` does not write to the graph.

Surface covered (SURVEY section 8b "DGL surface the hot path touches"): heterograph,
batch/unbatch, nodes[nt].data / edges[et].data, ndata/srcdata/dstdata, local_scope,
apply_edges (builtin + UDF), update_all / multi_update_all (copy_e|copy_u|u_mul_e x
sum|mean, cross_reducer='sum'), add_edges/remove_edges, add_nodes/remove_nodes,
batch_num_nodes/edges (+set_*), readout_nodes, num_nodes/num_edges, edges(form=), to().

Test infrastructure only; never shipped to the GPU box code path.
"""
import contextlib
import types

import torch

ALL = "__ALL__"


# ------------------------------------------------------------------ dgl.function

class _MsgBuiltin:
    def __init__(self, kind, a, b, out):
        self.kind, self.a, self.b, self.out = kind, a, b, out

    def __call__(self, edges):
        k = self.kind
        if k == "u_sub_v":
            return {self.out: edges.src[self.a] - edges.dst[self.b]}
        if k == "copy_e":
            return {self.out: edges.data[self.a]}
        if k == "copy_u":
            return {self.out: edges.src[self.a]}
        if k == "u_mul_e":
            return {self.out: _bcast_mul(edges.src[self.a], edges.data[self.b])}
        if k == "v_mul_e":
            return {self.out: _bcast_mul(edges.dst[self.a], edges.data[self.b])}
        if k == "u_dot_v":
            return {self.out: (edges.src[self.a] * edges.dst[self.b]).sum(-1, keepdim=True)}
        raise NotImplementedError(k)


def _bcast_mul(a, b):
    # DGL broadcasts the FEATURE shapes numpy-style (right-aligned): (3,) x (1, 1) -> (1, 3)
    while b.dim() < a.dim():
        b = b.unsqueeze(1)
    while a.dim() < b.dim():
        a = a.unsqueeze(1)
    return a * b


class _Reducer:
    def __init__(self, op, msg, out):
        self.op, self.msg, self.out = op, msg, out


function = types.ModuleType("dgl.function")
function.u_sub_v = lambda a, b, out: _MsgBuiltin("u_sub_v", a, b, out)
function.copy_e = lambda a, out: _MsgBuiltin("copy_e", a, None, out)
function.copy_u = lambda a, out: _MsgBuiltin("copy_u", a, None, out)
function.u_mul_e = lambda a, b, out: _MsgBuiltin("u_mul_e", a, b, out)
function.v_mul_e = lambda a, b, out: _MsgBuiltin("v_mul_e", a, b, out)
function.u_dot_v = lambda a, b, out: _MsgBuiltin("u_dot_v", a, b, out)
function.sum = lambda msg, out: _Reducer("sum", msg, out)
function.mean = lambda msg, out: _Reducer("mean", msg, out)


# ------------------------------------------------------------------ views

class _EdgeBatch:
    def __init__(self, g, cet):
        s, d = g._edges[cet]
        self.canonical_etype = cet
        self._g, self._s, self._d = g, s, d
        self.src = _Gather(g._ndata[cet[0]], s)
        self.dst = _Gather(g._ndata[cet[2]], d)
        self.data = g._edata[cet]


class _NodeBatch:
    def __init__(self, mailbox):
        self.mailbox = mailbox


class _Gather:
    def __init__(self, store, idx):
        self._store, self._idx = store, idx

    def __getitem__(self, key):
        return self._store[key][self._idx]


class _Space:
    def __init__(self, data):
        self.data = data


class _NodesAccessor:
    def __init__(self, g):
        self._g = g

    def __getitem__(self, ntype):
        return _Space(self._g._ndata[ntype])

    def __call__(self, ntype=None):
        return torch.arange(self._g._num_nodes[ntype], device=self._g.device)


class _EdgesAccessor:
    def __init__(self, g):
        self._g = g

    def __getitem__(self, etype):
        return _Space(self._g._edata[self._g.to_canonical_etype(etype)])

    def __call__(self, form="uv", etype=None, order="eid"):
        cet = self._g.to_canonical_etype(etype)
        s, d = self._g._edges[cet]
        if form == "uv":
            return s, d
        if form == "eid":
            return torch.arange(s.shape[0], device=self._g.device)
        if form == "all":
            return s, d, torch.arange(s.shape[0], device=self._g.device)
        raise ValueError(form)


class _MultiNodeDataView:
    """g.ndata / g.srcdata / g.dstdata on a multi-ntype graph: fresh dict per access."""

    def __init__(self, g):
        self._g = g

    def __getitem__(self, key):
        return {nt: st[key] for nt, st in self._g._ndata.items() if key in st}

    def __setitem__(self, key, val):
        assert isinstance(val, dict), "multi-type graph: value must be {ntype: tensor}"
        for nt, v in val.items():
            self._g._ndata[nt][key] = v

    def __contains__(self, key):
        return any(key in st for st in self._g._ndata.values())


# ------------------------------------------------------------------ graph

class DGLHeteroGraph:
    def __init__(self, edges, num_nodes, device="cpu"):
        self._edges = {k: (torch.as_tensor(s, dtype=torch.long), torch.as_tensor(d, dtype=torch.long))
                       for k, (s, d) in edges.items()}
        self._num_nodes = dict(num_nodes)
        self._ndata = {nt: {} for nt in self._num_nodes}
        self._edata = {et: {} for et in self._edges}
        self._bnn = None
        self._bne = None
        self.device = torch.device(device)

    # --- schema
    @property
    def ntypes(self):
        return sorted(self._num_nodes.keys())

    @property
    def canonical_etypes(self):
        return sorted(self._edges.keys(), key=lambda e: e[1])

    @property
    def etypes(self):
        return [e[1] for e in self.canonical_etypes]

    def to_canonical_etype(self, etype):
        if isinstance(etype, tuple):
            return etype
        for cet in self._edges:
            if cet[1] == etype:
                return cet
        raise KeyError(etype)

    def num_nodes(self, ntype=None):
        return self._num_nodes[ntype]

    def num_edges(self, etype=None):
        return int(self._edges[self.to_canonical_etype(etype)][0].shape[0])

    # --- views
    @property
    def nodes(self):
        return _NodesAccessor(self)

    @property
    def edges(self):
        return _EdgesAccessor(self)

    @property
    def ndata(self):
        return _MultiNodeDataView(self)

    srcdata = ndata
    dstdata = ndata

    # --- batching info
    @property
    def batch_size(self):
        if self._bnn is None:
            return 1
        return int(next(iter(self._bnn.values())).shape[0])

    def batch_num_nodes(self, ntype=None):
        if self._bnn is None:
            return torch.tensor([self._num_nodes[ntype]], dtype=torch.long, device=self.device)
        return self._bnn[ntype]

    def batch_num_edges(self, etype=None):
        cet = self.to_canonical_etype(etype)
        if self._bne is None:
            return torch.tensor([self.num_edges(cet)], dtype=torch.long, device=self.device)
        return self._bne[cet]

    def set_batch_num_nodes(self, val):
        self._bnn = {k: v for k, v in val.items()}

    def set_batch_num_edges(self, val):
        self._bne = {self.to_canonical_etype(k): v for k, v in val.items()}

    # --- scope
    @contextlib.contextmanager
    def local_scope(self):
        nd = {nt: dict(st) for nt, st in self._ndata.items()}
        ed = {et: dict(st) for et, st in self._edata.items()}
        try:
            yield
        finally:
            self._ndata = nd
            # edge sets may have been mutated inside the scope; keep only features whose
            # length still matches (DGL drops features of removed edges likewise)
            for et, st in ed.items():
                n = self._edges[et][0].shape[0]
                ed[et] = {k: v for k, v in st.items() if v.shape[0] == n}
            self._edata = ed

    # --- mutation
    def add_edges(self, u, v, data=None, etype=None):
        cet = self.to_canonical_etype(etype)
        s, d = self._edges[cet]
        u = torch.as_tensor(u, dtype=torch.long)
        v = torch.as_tensor(v, dtype=torch.long)
        old = s.shape[0]
        self._edges[cet] = (torch.cat([s, u]), torch.cat([d, v]))
        for k, t in list(self._edata[cet].items()):
            pad = torch.zeros((u.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype)
            self._edata[cet][k] = torch.cat([t[:old], pad])
        if data:
            for k, t in data.items():
                self._edata[cet][k] = t
        self._bnn = self._bnn  # DGL resets batch info on mutation; the reference re-sets it itself

    def remove_edges(self, eids, etype=None):
        cet = self.to_canonical_etype(etype)
        s, d = self._edges[cet]
        keep = torch.ones(s.shape[0], dtype=torch.bool)
        keep[eids] = False
        self._edges[cet] = (s[keep], d[keep])
        for k, t in list(self._edata[cet].items()):
            self._edata[cet][k] = t[keep]

    def add_nodes(self, num, data=None, ntype=None):
        old = self._num_nodes[ntype]
        self._num_nodes[ntype] = old + int(num)
        for k, t in list(self._ndata[ntype].items()):
            pad = torch.zeros((int(num),) + tuple(t.shape[1:]), dtype=t.dtype)
            self._ndata[ntype][k] = torch.cat([t, pad])
        if data:
            for k, t in data.items():
                if k in self._ndata[ntype]:
                    self._ndata[ntype][k] = torch.cat([self._ndata[ntype][k][:old], t])
                else:
                    assert old == 0, "new feature on non-empty node set not supported by the shim"
                    self._ndata[ntype][k] = t

    def remove_nodes(self, nids, ntype=None):
        n = self._num_nodes[ntype]
        keep = torch.ones(n, dtype=torch.bool)
        keep[nids] = False
        remap = torch.cumsum(keep.long(), 0) - 1
        self._num_nodes[ntype] = int(keep.sum())
        for k, t in list(self._ndata[ntype].items()):
            self._ndata[ntype][k] = t[keep]
        for cet, (s, d) in list(self._edges.items()):
            ek = torch.ones(s.shape[0], dtype=torch.bool)
            if cet[0] == ntype:
                ek &= keep[s]
            if cet[2] == ntype:
                ek &= keep[d]
            s2, d2 = s[ek], d[ek]
            if cet[0] == ntype:
                s2 = remap[s2]
            if cet[2] == ntype:
                d2 = remap[d2]
            self._edges[cet] = (s2, d2)
            for k, t in list(self._edata[cet].items()):
                self._edata[cet][k] = t[ek]

    # --- compute
    def apply_edges(self, func, etype=None):
        cet = self.to_canonical_etype(etype)
        out = func(_EdgeBatch(self, cet))
        for k, v in out.items():
            self._edata[cet][k] = v

    def _reduce(self, cet, mfunc, rfunc):
        msgs = mfunc(_EdgeBatch(self, cet))
        m = msgs[rfunc.msg]
        d = self._edges[cet][1]
        n = self._num_nodes[cet[2]]
        out = torch.zeros((n,) + tuple(m.shape[1:]), dtype=m.dtype)
        out.index_add_(0, d, m)
        if rfunc.op == "mean":
            deg = torch.bincount(d, minlength=n).clamp(min=1).to(m.dtype)
            out = out / deg.view(-1, *([1] * (m.dim() - 1)))
        return out

    def update_all(self, mfunc, rfunc, etype=None):
        cet = self.to_canonical_etype(etype)
        if not isinstance(rfunc, _Reducer):
            # user-defined reduce function over nodes.mailbox (reference receptor_encoder.py:294,299-300).  DGL buckets
            # nodes by in-degree and orders each node's mailbox by edge id; the reference only uses this with one
            # common in-degree (k nearest neighbours per keypoint), which is the case restated here.
            msgs = mfunc(_EdgeBatch(self, cet))
            d = self._edges[cet][1]
            n = self._num_nodes[cet[2]]
            deg = torch.bincount(d, minlength=n)
            assert n > 0 and bool((deg == deg[0]).all()) and int(deg[0]) > 0, "mailbox stand-in needs one common in-degree"
            order = torch.sort(d, stable=True).indices
            mailbox = {k: v[order].reshape((n, int(deg[0])) + tuple(v.shape[1:])) for k, v in msgs.items()}
            for k, v in rfunc(_NodeBatch(mailbox)).items():
                self._ndata[cet[2]][k] = v
            return
        self._ndata[cet[2]][rfunc.out] = self._reduce(cet, mfunc, rfunc)

    def multi_update_all(self, etype_dict, cross_reducer="sum"):
        assert cross_reducer == "sum"
        acc = {}
        for etype, (mfunc, rfunc) in etype_dict.items():
            cet = self.to_canonical_etype(etype)
            r = self._reduce(cet, mfunc, rfunc)
            key = (cet[2], rfunc.out)
            acc[key] = r if key not in acc else acc[key] + r
        for (nt, out), v in acc.items():
            self._ndata[nt][out] = v

    def in_degrees(self, v, etype=None):
        cet = self.to_canonical_etype(etype)
        return torch.bincount(self._edges[cet][1], minlength=self._num_nodes[cet[2]])[v]

    def out_degrees(self, u, etype=None):
        cet = self.to_canonical_etype(etype)
        return torch.bincount(self._edges[cet][0], minlength=self._num_nodes[cet[0]])[u]

    def to(self, device):
        return self  # CPU only


DGLGraph = DGLHeteroGraph


def heterograph(data_dict, num_nodes_dict=None, device="cpu"):
    edges = {}
    for cet, (s, d) in data_dict.items():
        edges[cet] = (torch.as_tensor(s, dtype=torch.long).reshape(-1), torch.as_tensor(d, dtype=torch.long).reshape(-1))
    return DGLHeteroGraph(edges, num_nodes_dict)


def batch(graphs):
    g0 = graphs[0]
    edges, num_nodes = {}, {}
    off = {nt: 0 for nt in g0._num_nodes}
    parts = {cet: ([], []) for cet in g0._edges}
    for g in graphs:
        for cet, (s, d) in g._edges.items():
            parts[cet][0].append(s + off[cet[0]])
            parts[cet][1].append(d + off[cet[2]])
        for nt in off:
            off[nt] += g._num_nodes[nt]
    for cet, (ss, dd) in parts.items():
        edges[cet] = (torch.cat(ss), torch.cat(dd))
    out = DGLHeteroGraph(edges, off)
    for nt in g0._num_nodes:
        for k in g0._ndata[nt]:
            out._ndata[nt][k] = torch.cat([g._ndata[nt][k] for g in graphs])
    for cet in g0._edges:
        for k in g0._edata[cet]:
            out._edata[cet][k] = torch.cat([g._edata[cet][k] for g in graphs])
    out._bnn = {nt: torch.cat([g.batch_num_nodes(nt) for g in graphs]) for nt in g0._num_nodes}
    out._bne = {cet: torch.cat([g.batch_num_edges(cet) for g in graphs]) for cet in g0._edges}
    return out


def unbatch(g):
    B = g.batch_size
    outs = []
    noff = {nt: 0 for nt in g._num_nodes}
    eoff = {cet: 0 for cet in g._edges}
    for b in range(B):
        nn = {nt: int(g.batch_num_nodes(nt)[b]) for nt in g._num_nodes}
        edges = {}
        ne = {}
        for cet, (s, d) in g._edges.items():
            n = int(g.batch_num_edges(cet)[b])
            ne[cet] = n
            sl = slice(eoff[cet], eoff[cet] + n)
            edges[cet] = (s[sl] - noff[cet[0]], d[sl] - noff[cet[2]])
        gi = DGLHeteroGraph(edges, nn)
        for nt in nn:
            for k, t in g._ndata[nt].items():
                gi._ndata[nt][k] = t[noff[nt]:noff[nt] + nn[nt]]
        for cet in edges:
            for k, t in g._edata[cet].items():
                gi._edata[cet][k] = t[eoff[cet]:eoff[cet] + ne[cet]]
        for nt in nn:
            noff[nt] += nn[nt]
        for cet in ne:
            eoff[cet] += ne[cet]
        outs.append(gi)
    return outs


def readout_nodes(g, feat, weight=None, op="sum", ntype=None):
    x = g._ndata[ntype][feat]
    counts = g.batch_num_nodes(ntype)
    B = counts.shape[0]
    idx = torch.arange(B).repeat_interleave(counts)
    out = torch.zeros((B,) + tuple(x.shape[1:]), dtype=x.dtype)
    out.index_add_(0, idx, x)
    if op == "mean":
        out = out / counts.to(x.dtype).view(-1, *([1] * (x.dim() - 1)))
    elif op != "sum":
        raise NotImplementedError(op)
    return out


def make_module():
    m = types.ModuleType("dgl")
    m.function = function
    m.DGLHeteroGraph = DGLHeteroGraph
    m.DGLGraph = DGLGraph
    m.heterograph = heterograph
    m.batch = batch
    m.unbatch = unbatch
    m.readout_nodes = readout_nodes
    return m
