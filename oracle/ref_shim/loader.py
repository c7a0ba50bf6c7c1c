"""Import the reference's own modules (read-only, from /root/reference) in an environment
where dgl / torch_cluster / torch_scatter / ot / openbabel are not installed, by
registering stand-ins in ``sys.modules`` first:

  dgl, dgl.function, dgl.nn.functional   -> oracle/ref_shim/dgl_shim.py
  torch_cluster                          -> oracle/graph.py (restated radius/knn semantics)
  torch_scatter                          -> small segment_csr / segment_coo below
  ot, openbabel                          -> empty modules (training loss / file I/O only)

Only usable in the build container (the GPU box has no /root/reference); used by
tests/golden/make_golden.py and tests/test_oracle_vs_reference.py.

Test infrastructure only.
"""
import importlib
import os
import sys
import types

import torch

from .. import graph as G
from . import dgl_shim

REFERENCE_ROOT = os.environ.get("KPD_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "ligand_diffuser.py"))


def _segment_csr(src, indptr, out=None, reduce="sum"):
    n = indptr.shape[0] - 1
    counts = (indptr[1:] - indptr[:-1]).long()
    idx = torch.arange(n).repeat_interleave(counts)
    res = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype)
    res.index_add_(0, idx, src[int(indptr[0]):int(indptr[-1])])
    if reduce == "mean":
        res = res / counts.clamp(min=1).to(src.dtype).view(-1, *([1] * (src.dim() - 1)))
    elif reduce != "sum":
        raise NotImplementedError(reduce)
    return res


def _segment_coo(src, index, out=None, dim_size=None, reduce="sum"):
    n = int(index.max()) + 1 if dim_size is None else dim_size
    res = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype)
    res.index_add_(0, index, src)
    return res


def _edge_softmax(graph, logits, eids=None, norm_by="dst"):
    raise NotImplementedError("edge_softmax is only used by the dead KeyKeyConv path")


_installed = False


def install_stubs():
    global _installed
    if _installed:
        return
    dgl = dgl_shim.make_module()
    sys.modules["dgl"] = dgl
    sys.modules["dgl.function"] = dgl_shim.function
    nn_mod = types.ModuleType("dgl.nn")
    nnf = types.ModuleType("dgl.nn.functional")
    nnf.edge_softmax = _edge_softmax
    nn_mod.functional = nnf
    dgl.nn = nn_mod
    sys.modules["dgl.nn"] = nn_mod
    sys.modules["dgl.nn.functional"] = nnf

    tc = types.ModuleType("torch_cluster")
    tc.radius = lambda x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32: G.radius(
        x, y, r, _b(batch_x, x), _b(batch_y, y), max_num_neighbors)
    tc.radius_graph = lambda x, r, batch=None, loop=False, max_num_neighbors=32: G.radius_graph(
        x, r, _b(batch, x), loop, max_num_neighbors)
    tc.knn = lambda x, y, k, batch_x=None, batch_y=None: G.knn(x, y, k, _b(batch_x, x), _b(batch_y, y))
    tc.knn_graph = lambda x, k, batch=None, loop=False: G.knn_graph(x, k, _b(batch, x), loop)
    sys.modules["torch_cluster"] = tc

    ts = types.ModuleType("torch_scatter")
    ts.segment_csr = _segment_csr
    ts.segment_coo = _segment_coo
    sys.modules["torch_scatter"] = ts

    for name in ("ot", "openbabel"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    _installed = True


def _b(batch, x):
    return torch.zeros(x.shape[0], dtype=torch.long) if batch is None else batch


def import_reference():
    """Returns a namespace with the reference's modules: .model_setup, .ligand_diffuser,
    .dynamics, .dynamics_gvp, .gvp, .utils."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    ns.utils = importlib.import_module("utils")
    ns.gvp = importlib.import_module("models.gvp")
    ns.dynamics = importlib.import_module("models.dynamics")
    ns.dynamics_gvp = importlib.import_module("models.dynamics_gvp")
    ns.ligand_diffuser = importlib.import_module("models.ligand_diffuser")
    ns.model_setup = importlib.import_module("model_setup")
    ns.dgl = sys.modules["dgl"]
    return ns
