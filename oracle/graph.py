"""Brute-force radius / kNN graph construction (oracle, CPU).

The reference calls torch_cluster (third-party, un-vendored, unpinned --
readme.md:15) at models/dynamics.py:393-404 and models/dynamics_gvp.py:207-218.
torch_cluster is absent from /root/reference and from this image, so its
semantics are restated here from its documented behaviour (torch-cluster 1.6.x):

  radius(x, y, r, batch_x, batch_y, max_num_neighbors)
      for every y (ascending), scan the x of the same example in ascending
      index, keep those with squared distance < r*r (strict), stop after
      max_num_neighbors hits.  Returns [y_idx; x_idx] grouped by y.
  radius_graph(x, r, batch, loop=False, max_num_neighbors, flow='source_to_target')
      = radius(x, x, r, batch, batch, max_num_neighbors + 1), rows swapped to
      [neighbour; centre], self pairs dropped.
  knn(x, y, k, batch_x, batch_y)
      for every y the k nearest x of the same example, ascending squared
      distance, insertion with strict '<' (lower index wins ties), fewer than k
      when the example has fewer x.  Returns [y_idx; x_idx].
  knn_graph(x, k, batch, loop=False) = knn(x, x, k + 1, ...) swapped, self dropped.

Distance arithmetic (the part that decides bit-exact edge sets): fp32,
unfused, sequential over the three coordinates --
    d2 = ((dx*dx) + (dy*dy)) + (dz*dz),   dx = x[0] - y[0]  (each op rounded)
which is what torch_cluster's CPU path (nanoflann L2_Simple_Adaptor) and a
non-contracted build of its CUDA loop compute.  The CUDA kernel in
keypoint_diffusion_b200/csrc/graph_build.cu uses __fmul_rn/__fadd_rn to match.

"parity unpinned" at this boundary: no torch_cluster build is available to
confirm tie/cap ordering; edge *sets* are what the tests compare.

Test infrastructure only (see oracle/__init__.py).
"""
import torch


def _ptr_from_batch(batch, n_examples):
    counts = torch.bincount(batch, minlength=n_examples)
    ptr = torch.zeros(n_examples + 1, dtype=torch.long)
    ptr[1:] = torch.cumsum(counts, 0)
    return ptr


def pairwise_d2(y, x):
    """[n_y, n_x] fp32 squared distances, unfused sequential (see module doc)."""
    y = y.float()
    x = x.float()
    dx = x[None, :, 0] - y[:, None, 0]
    dy = x[None, :, 1] - y[:, None, 1]
    dz = x[None, :, 2] - y[:, None, 2]
    return ((dx * dx) + (dy * dy)) + (dz * dz)


def radius(x, y, r, batch_x, batch_y, max_num_neighbors=32):
    n_ex = int(max(batch_x.max().item() if batch_x.numel() else -1,
                   batch_y.max().item() if batch_y.numel() else -1)) + 1
    ptr_x = _ptr_from_batch(batch_x, n_ex)
    ptr_y = _ptr_from_batch(batch_y, n_ex)
    r2 = torch.tensor(float(r) * float(r), dtype=torch.float64).float()  # (float)(r*r)
    rows, cols = [], []
    for b in range(n_ex):
        xs, xe = int(ptr_x[b]), int(ptr_x[b + 1])
        ys, ye = int(ptr_y[b]), int(ptr_y[b + 1])
        if xe == xs or ye == ys:
            continue
        d2 = pairwise_d2(y[ys:ye], x[xs:xe])
        hit = d2 < r2
        keep = hit & (torch.cumsum(hit.long(), dim=1) <= max_num_neighbors)
        yi, xi = torch.nonzero(keep, as_tuple=True)  # row-major: grouped by y, x ascending
        rows.append(yi + ys)
        cols.append(xi + xs)
    if not rows:
        return torch.zeros(2, 0, dtype=torch.long)
    return torch.stack([torch.cat(rows), torch.cat(cols)])


def radius_graph(x, r, batch, loop=False, max_num_neighbors=32):
    ei = radius(x, x, r, batch, batch, max_num_neighbors if loop else max_num_neighbors + 1)
    row, col = ei[1], ei[0]  # flow = source_to_target: [neighbour; centre]
    if not loop:
        m = row != col
        row, col = row[m], col[m]
    return torch.stack([row, col])


def knn(x, y, k, batch_x, batch_y):
    n_ex = int(max(batch_x.max().item() if batch_x.numel() else -1,
                   batch_y.max().item() if batch_y.numel() else -1)) + 1
    ptr_x = _ptr_from_batch(batch_x, n_ex)
    ptr_y = _ptr_from_batch(batch_y, n_ex)
    rows, cols = [], []
    for b in range(n_ex):
        xs, xe = int(ptr_x[b]), int(ptr_x[b + 1])
        ys, ye = int(ptr_y[b]), int(ptr_y[b + 1])
        if xe == xs or ye == ys:
            continue
        d2 = pairwise_d2(y[ys:ye], x[xs:xe])
        kk = min(k, xe - xs)
        order = torch.sort(d2, dim=1, stable=True).indices[:, :kk]
        yi = torch.arange(ye - ys)[:, None].expand(-1, kk)
        rows.append(yi.reshape(-1) + ys)
        cols.append(order.reshape(-1) + xs)
    if not rows:
        return torch.zeros(2, 0, dtype=torch.long)
    return torch.stack([torch.cat(rows), torch.cat(cols)])


def knn_graph(x, k, batch, loop=False):
    ei = knn(x, x, k if loop else k + 1, batch, batch)
    row, col = ei[1], ei[0]
    if not loop:
        m = row != col
        row, col = row[m], col[m]
    return torch.stack([row, col])


def edges_per_batch(edge_node_idxs, batch_size, node_batch_idxs):
    """/root/reference/utils.py:92-98 (get_edges_per_batch), via bincount (same result
    for edges grouped by complex, which is what the reference assumes)."""
    if edge_node_idxs.numel() == 0:
        return torch.zeros(batch_size, dtype=torch.long)
    return torch.bincount(node_batch_idxs[edge_node_idxs], minlength=batch_size)
