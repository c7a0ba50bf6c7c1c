"""CPU oracle for the keypoint-diffusion sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as
the thing timed *as the CPU baseline* -- never as a fallback for the CUDA path.

Parity pinning status (see DESIGN.md section "Oracle"):
  * the reference (Dunni3/keypoint-diffusion) ships no tests, golden vectors
    or checkpoints, and its arithmetic lives in un-vendored third-party
    libraries (DGL, torch_cluster, torch_scatter; versions unpinned,
    readme.md:13-21), so there is nothing reference-owned to pin against;
  * what we pin instead: ``oracle/flat.py`` (this restatement) is checked
    against the reference's *own* ``nn.Module.forward`` code, imported
    read-only from /root/reference and executed over a small pure-PyTorch
    stand-in for the DGL / torch_cluster subset it calls
    (``oracle/ref_shim``).  Inputs, weights and outputs of those runs are
    committed under ``tests/golden/`` together with the generating script;
  * the third-party semantics themselves (torch_cluster radius/knn,
    DGL reducers) are restated from their documented behaviour:
    "parity unpinned" at that boundary.
"""
