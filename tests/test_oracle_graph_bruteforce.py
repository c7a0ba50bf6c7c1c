"""CPU: the vectorised graph oracle (oracle/graph.py) against literal, loop-by-loop restatements of torch_cluster's
documented algorithms (radius: ascending scan with a strict '<' test and a neighbour cap; knn: insertion into a sorted
candidate list with strict '<', i.e. the lower index wins ties), on small ragged batches whose coordinates sit on a
coarse lattice so that equal distances and cap overflows are frequent.  numpy float32 scalar arithmetic in the same
unfused order as the oracle / the CUDA kernel."""
import numpy as np
import pytest
import torch

from oracle import graph as G


def _d2(a, b):
    dx = np.float32(a[0] - b[0]); dy = np.float32(a[1] - b[1]); dz = np.float32(a[2] - b[2])
    return np.float32(np.float32(np.float32(dx * dx) + np.float32(dy * dy)) + np.float32(dz * dz))


def _radius_loops(x, y, r, bx, by, cap):
    r2 = np.float32(float(r) * float(r))
    out = []
    for j in range(len(y)):
        cnt = 0
        for i in range(len(x)):
            if bx[i] != by[j]:
                continue
            if _d2(x[i], y[j]) < r2:
                out.append((j, i))
                cnt += 1
                if cnt >= cap:
                    break
    return out


def _knn_loops(x, y, k, bx, by):
    out = []
    for j in range(len(y)):
        best = []                                   # ascending (d2, i); a new candidate goes BEHIND equal distances
        for i in range(len(x)):
            if bx[i] != by[j]:
                continue
            d = _d2(x[i], y[j])
            pos = len(best)
            while pos > 0 and d < best[pos - 1][0]:
                pos -= 1
            best.insert(pos, (d, i))
            del best[k:]
        out += [(j, i) for _, i in best]
    return out


def _case(seed, sizes_x, sizes_y, quantum):
    rng = np.random.default_rng(seed)
    xs, ys, bx, by = [], [], [], []
    for b, (nx, ny) in enumerate(zip(sizes_x, sizes_y)):
        c = rng.normal(size=3) * 10
        xs.append(np.round((c + rng.normal(size=(nx, 3)) * 1.5) / quantum) * quantum)
        ys.append(np.round((c + rng.normal(size=(ny, 3)) * 2.0) / quantum) * quantum)
        bx += [b] * nx
        by += [b] * ny
    return np.concatenate(xs).astype(np.float32), np.concatenate(ys).astype(np.float32), np.array(bx), np.array(by)


def _pairs(ei):
    return list(zip(ei[0].tolist(), ei[1].tolist()))


@pytest.mark.parametrize("seed", range(4))
def test_radius_and_knn_match_the_literal_loops(seed):
    x, y, bx, by = _case(seed, [1, 7, 20, 3, 12], [5, 9, 1, 14, 6], quantum=0.5)
    tx, ty, tbx, tby = torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(bx), torch.from_numpy(by)
    for r, cap in ((2.0, 100), (3.0, 4), (1.0, 1)):
        got = _pairs(G.radius(tx, ty, r, tbx, tby, max_num_neighbors=cap))
        assert got == _radius_loops(x, y, r, bx, by, cap), (r, cap)          # same pairs in the same order
    for k in (1, 3, 5, 25):
        got = _pairs(G.knn(tx, ty, k, tbx, tby))
        assert got == _knn_loops(x, y, k, bx, by), k


@pytest.mark.parametrize("seed", range(3))
def test_graph_variants_match_the_literal_loops(seed):
    """radius_graph = radius(x, x, cap + 1) with the self pair dropped, rows swapped to [neighbour; centre];
    knn_graph = knn(x, x, k + 1) likewise."""
    x, _, bx, _ = _case(10 + seed, [1, 9, 25, 4], [1, 1, 1, 1], quantum=0.5)
    tx, tbx = torch.from_numpy(x), torch.from_numpy(bx)
    for r, cap in ((2.5, 200), (2.5, 3)):
        ref = [(i, j) for j, i in _radius_loops(x, x, r, bx, bx, cap + 1) if i != j]
        assert _pairs(G.radius_graph(tx, r, tbx, max_num_neighbors=cap)) == ref
    for k in (2, 4):
        ref = [(i, j) for j, i in _knn_loops(x, x, k + 1, bx, bx) if i != j]
        assert _pairs(G.knn_graph(tx, k, tbx)) == ref
