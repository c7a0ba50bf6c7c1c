// (a) Per-step graph construction: ligand-ligand radius/kNN graph and keypoint<->ligand
// kNN/radius graph, emitted as dst-sorted CSR (+COO) without leaving the device.
//
// Replaces torch_cluster.radius_graph / knn_graph / knn / radius + DGL add_edges/remove_edges +
// utils.get_edges_per_batch as called from models/dynamics.py:387-442 and
// models/dynamics_gvp.py:201-255 of the reference.
//
// One CTA per complex.  Coordinates of the complex are staged in shared memory; every thread
// owns one query node and scans the candidates of its complex.  Distances use unfused fp32
// (__fsub_rn/__fmul_rn/__fadd_rn, sequential over x,y,z) so that the `< r*r` and kNN
// comparisons are bit-identical to the oracle (oracle/graph.py) -- edge sets are compared
// exactly.  Adjacency is recorded as bit rows, then
//   pass 1 (graph_mark):  bit rows + per-node degrees + per-complex totals
//   pass 2 (graph_scan):  exclusive scan of the per-complex totals (one CTA)
//   pass 3 (graph_fill):  rowptr + src/dst by enumerating the bit rows in ascending order
// Work is HBM/L2-trivial (12 B per node in, 8 B per edge out); the kernels are latency bound.
#include "common.cuh"

namespace kpd {

struct GraphWs {
    uint32_t* bits_ll;   // [n_lig][wl]   row j: ligand sources i of dst j
    uint32_t* bits_kl;   // [n_lig][wk]   row j: keypoint sources y of dst ligand j
    uint32_t* bits_lk;   // [n_kp][wl]    row y: ligand sources of dst keypoint y
    int* deg_ll;         // [n_lig]
    int* deg_kl;         // [n_lig]
    int* deg_lk;         // [n_kp]
    int* tot_ll;         // [B]
    int* tot_kl;         // [B]
    int* off_ll;         // [B]
    int* off_kl;         // [B]
    int wl, wk;
};

static GraphWs carve_graph_ws(const kpd_batch* b, void* ws, int64_t* bytes) {
    GraphWs g;
    g.wl = cdiv(b->max_lig, 32);
    g.wk = cdiv(b->max_kp, 32);
    Carver c(ws);
    g.bits_ll = c.take<uint32_t>((int64_t)b->n_lig * g.wl);
    g.bits_kl = c.take<uint32_t>((int64_t)b->n_lig * g.wk);
    g.bits_lk = c.take<uint32_t>((int64_t)b->n_kp * g.wl);
    g.deg_ll = c.take<int>(b->n_lig);
    g.deg_kl = c.take<int>(b->n_lig);
    g.deg_lk = c.take<int>(b->n_kp);
    g.tot_ll = c.take<int>(b->B);
    g.tot_kl = c.take<int>(b->B);
    g.off_ll = c.take<int>(b->B);
    g.off_kl = c.take<int>(b->B);
    if (bytes) *bytes = c.bytes();
    return g;
}

__device__ __forceinline__ float dist2(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// k nearest candidates of (qx,qy,qz) among xs[0..n): ascending distance, strict '<' insertion
// (earlier index wins ties) -- torch_cluster knn semantics.  Returns the count (<= k).
__device__ int knn_select(float qx, float qy, float qz, const float* xs, int n, int k,
                          float* best_d, int* best_i) {
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        const float d = dist2(xs[3 * i], xs[3 * i + 1], xs[3 * i + 2], qx, qy, qz);
        int pos = cnt;
        while (pos > 0 && best_d[pos - 1] > d) --pos;   // strict: equal distances stay ahead
        if (pos >= k) continue;
        const int last = min(cnt, k - 1);
        for (int m = last; m > pos; --m) {
            best_d[m] = best_d[m - 1];
            best_i[m] = best_i[m - 1];
        }
        best_d[pos] = d;
        best_i[pos] = i;
        if (cnt < k) ++cnt;
    }
    return cnt;
}

__global__ void __launch_bounds__(128)
graph_mark_kernel(kpd_batch b, const float* __restrict__ x_lig, const float* __restrict__ x_kp,
                  kpd_graph_params p, GraphWs g) {
    extern __shared__ float smem[];
    const int c = blockIdx.x;
    const int l0 = b.lig_ptr[c], nl = b.lig_ptr[c + 1] - l0;
    const int k0 = b.kp_ptr[c], nk = b.kp_ptr[c + 1] - k0;
    float* xl = smem;                                         // [max_lig*3]
    uint32_t* klb = reinterpret_cast<uint32_t*>(smem + 3 * b.max_lig);  // [max_lig][wk]
    __shared__ int s_tot_ll, s_tot_kl;

    for (int i = threadIdx.x; i < 3 * nl; i += blockDim.x) xl[i] = x_lig[3 * l0 + i];
    for (int i = threadIdx.x; i < nl * g.wk; i += blockDim.x) klb[i] = 0u;
    if (threadIdx.x == 0) { s_tot_ll = 0; s_tot_kl = 0; }
    __syncthreads();

    float best_d[KPD_MAX_KNN + 1];
    int best_i[KPD_MAX_KNN + 1];

    // ---- ll: one thread per destination ligand atom j
    for (int j = threadIdx.x; j < nl; j += blockDim.x) {
        uint32_t* row = g.bits_ll + (size_t)(l0 + j) * g.wl;
        for (int w = 0; w < g.wl; ++w) row[w] = 0u;
        const float qx = xl[3 * j], qy = xl[3 * j + 1], qz = xl[3 * j + 2];
        int deg = 0;
        if (p.ll_k > 0) {
            // knn_graph: knn(x, x, k+1) then drop the self pair
            const int cnt = knn_select(qx, qy, qz, xl, nl, p.ll_k + 1, best_d, best_i);
            for (int m = 0; m < cnt; ++m) {
                if (best_i[m] == j) continue;
                row[best_i[m] >> 5] |= 1u << (best_i[m] & 31);
                ++deg;
            }
        } else {
            // radius_graph: radius(x, x, r, max_num_neighbors + 1) then drop the self pair
            const float r2 = (float)(p.ll_r * p.ll_r);
            int hits = 0;
            for (int i = 0; i < nl && hits < p.ll_cap + 1; ++i) {
                const float d = dist2(xl[3 * i], xl[3 * i + 1], xl[3 * i + 2], qx, qy, qz);
                if (d < r2) {
                    ++hits;
                    if (i != j) { row[i >> 5] |= 1u << (i & 31); ++deg; }
                }
            }
        }
        g.deg_ll[l0 + j] = deg;
        atomicAdd(&s_tot_ll, deg);
    }

    // ---- kl / lk: one thread per keypoint y (the query of knn / radius)
    for (int y = threadIdx.x; y < nk; y += blockDim.x) {
        uint32_t* row = g.bits_lk + (size_t)(k0 + y) * g.wl;
        for (int w = 0; w < g.wl; ++w) row[w] = 0u;
        const float qx = x_kp[3 * (k0 + y)], qy = x_kp[3 * (k0 + y) + 1], qz = x_kp[3 * (k0 + y) + 2];
        int deg = 0;
        if (p.kl_k > 0) {
            const int cnt = knn_select(qx, qy, qz, xl, nl, p.kl_k, best_d, best_i);
            for (int m = 0; m < cnt; ++m) {
                const int i = best_i[m];
                row[i >> 5] |= 1u << (i & 31);
                atomicOr(&klb[i * g.wk + (y >> 5)], 1u << (y & 31));
            }
            deg = cnt;
        } else {
            const float r2 = (float)(p.kl_r * p.kl_r);
            for (int i = 0; i < nl && deg < p.kl_cap; ++i) {
                const float d = dist2(xl[3 * i], xl[3 * i + 1], xl[3 * i + 2], qx, qy, qz);
                if (d < r2) {
                    row[i >> 5] |= 1u << (i & 31);
                    atomicOr(&klb[i * g.wk + (y >> 5)], 1u << (y & 31));
                    ++deg;
                }
            }
        }
        g.deg_lk[k0 + y] = deg;
        atomicAdd(&s_tot_kl, deg);
    }
    __syncthreads();

    for (int j = threadIdx.x; j < nl; j += blockDim.x) {
        int deg = 0;
        for (int w = 0; w < g.wk; ++w) {
            const uint32_t v = klb[j * g.wk + w];
            g.bits_kl[(size_t)(l0 + j) * g.wk + w] = v;
            deg += __popc(v);
        }
        g.deg_kl[l0 + j] = deg;
    }
    if (threadIdx.x == 0) { g.tot_ll[c] = s_tot_ll; g.tot_kl[c] = s_tot_kl; }
}

// exclusive scan of the per-complex totals; also publishes the edge totals into rowptr[n_dst]
__global__ void __launch_bounds__(1024)
graph_scan_kernel(int B, GraphWs g, int* rowptr_ll, int n_lig, int* rowptr_kl, int* rowptr_lk, int n_kp,
                  int* counts_ll, int* counts_kl, long long* edge_accum) {
    __shared__ int s_a[1024], s_b[1024];
    __shared__ int carry_a, carry_b;
    if (threadIdx.x == 0) { carry_a = 0; carry_b = 0; }
    __syncthreads();
    for (int base = 0; base < B; base += 1024) {
        const int i = base + threadIdx.x;
        const int va = i < B ? g.tot_ll[i] : 0, vb = i < B ? g.tot_kl[i] : 0;
        s_a[threadIdx.x] = va;
        s_b[threadIdx.x] = vb;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {   // Hillis-Steele inclusive scan
            int ta = 0, tb = 0;
            if (threadIdx.x >= off) { ta = s_a[threadIdx.x - off]; tb = s_b[threadIdx.x - off]; }
            __syncthreads();
            s_a[threadIdx.x] += ta;
            s_b[threadIdx.x] += tb;
            __syncthreads();
        }
        if (i < B) {
            g.off_ll[i] = carry_a + s_a[threadIdx.x] - va;
            g.off_kl[i] = carry_b + s_b[threadIdx.x] - vb;
            if (counts_ll) counts_ll[i] = va;
            if (counts_kl) counts_kl[i] = vb;
        }
        __syncthreads();
        if (threadIdx.x == 1023) { carry_a += s_a[1023]; carry_b += s_b[1023]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        rowptr_ll[n_lig] = carry_a;
        rowptr_kl[n_lig] = carry_b;
        if (rowptr_lk) rowptr_lk[n_kp] = carry_b;
        if (edge_accum) { edge_accum[0] += carry_a; edge_accum[1] += carry_b; edge_accum[2] += 1; }
    }
}

// block-wide exclusive scan of vals[0..n) in shared memory (n may exceed blockDim)
__device__ void block_excl_scan(int* vals, int n, int* scratch /*[blockDim]*/) {
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < n ? vals[i] : 0;
        scratch[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < blockDim.x; off <<= 1) {
            int t = 0;
            if (threadIdx.x >= off) t = scratch[threadIdx.x - off];
            __syncthreads();
            scratch[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n) vals[i] = carry + scratch[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry += scratch[blockDim.x - 1];
        __syncthreads();
    }
}

__device__ __forceinline__ void emit_row(const uint32_t* bits, int words, int src_base, int dst_node,
                                         int pos, int* src, int* dst, int cap) {
    for (int w = 0; w < words; ++w) {
        uint32_t v = bits[w];
        while (v) {
            const int bit = __ffs(v) - 1;
            v &= v - 1;
            if (pos < cap) { src[pos] = src_base + (w << 5) + bit; dst[pos] = dst_node; }
            ++pos;
        }
    }
}

__global__ void __launch_bounds__(128)
graph_fill_kernel(kpd_batch b, GraphWs g, kpd_csr ll, kpd_csr kl, kpd_csr lk, int have_lk) {
    extern __shared__ int ismem[];
    int* pre = ismem;                 // [max(max_lig, max_kp)]
    int* scratch = ismem + max(b.max_lig, b.max_kp);   // [blockDim]
    const int c = blockIdx.x;
    const int l0 = b.lig_ptr[c], nl = b.lig_ptr[c + 1] - l0;
    const int k0 = b.kp_ptr[c], nk = b.kp_ptr[c + 1] - k0;

    // ll
    for (int j = threadIdx.x; j < nl; j += blockDim.x) pre[j] = g.deg_ll[l0 + j];
    __syncthreads();
    block_excl_scan(pre, nl, scratch);
    for (int j = threadIdx.x; j < nl; j += blockDim.x) {
        const int pos = g.off_ll[c] + pre[j];
        ll.rowptr[l0 + j] = pos;
        emit_row(g.bits_ll + (size_t)(l0 + j) * g.wl, g.wl, l0, l0 + j, pos, ll.src, ll.dst, ll.cap);
    }
    __syncthreads();
    // kl (dst = ligand, src = keypoint)
    for (int j = threadIdx.x; j < nl; j += blockDim.x) pre[j] = g.deg_kl[l0 + j];
    __syncthreads();
    block_excl_scan(pre, nl, scratch);
    for (int j = threadIdx.x; j < nl; j += blockDim.x) {
        const int pos = g.off_kl[c] + pre[j];
        kl.rowptr[l0 + j] = pos;
        emit_row(g.bits_kl + (size_t)(l0 + j) * g.wk, g.wk, k0, l0 + j, pos, kl.src, kl.dst, kl.cap);
    }
    __syncthreads();
    // lk (dst = keypoint, src = ligand)
    if (have_lk) {
        for (int y = threadIdx.x; y < nk; y += blockDim.x) pre[y] = g.deg_lk[k0 + y];
        __syncthreads();
        block_excl_scan(pre, nk, scratch);
        for (int y = threadIdx.x; y < nk; y += blockDim.x) {
            const int pos = g.off_kl[c] + pre[y];
            lk.rowptr[k0 + y] = pos;
            emit_row(g.bits_lk + (size_t)(k0 + y) * g.wl, g.wl, l0, k0 + y, pos, lk.src, lk.dst, lk.cap);
        }
    }
}

}  // namespace kpd

using namespace kpd;

extern "C" int64_t kpd_graph_workspace_bytes(const kpd_batch* batch) {
    int64_t bytes = 0;
    carve_graph_ws(batch, nullptr, &bytes);
    return bytes;
}

extern "C" int kpd_build_graph(const kpd_batch* batch, const float* x_lig, const float* x_kp,
                               const kpd_graph_params* p, kpd_csr* ll, kpd_csr* kl, kpd_csr* lk,
                               int32_t* counts_ll, int32_t* counts_kl, void* workspace, void* stream) {
    return build_graph_impl(batch, x_lig, x_kp, p, ll, kl, lk, counts_ll, counts_kl, workspace, nullptr,
                            static_cast<cudaStream_t>(stream));
}

int kpd::build_graph_impl(const kpd_batch* batch, const float* x_lig, const float* x_kp, const kpd_graph_params* p,
                          kpd_csr* ll, kpd_csr* kl, kpd_csr* lk, int32_t* counts_ll, int32_t* counts_kl,
                          void* workspace, long long* edge_accum, cudaStream_t st) {
    KPD_REQUIRE(batch && p && ll && kl, "kpd_build_graph: null argument");
    KPD_REQUIRE(batch->B > 0, "kpd_build_graph: empty batch");
    KPD_REQUIRE(p->ll_k <= KPD_MAX_KNN - 1 && p->kl_k <= KPD_MAX_KNN, "kpd_build_graph: k > %d unsupported", KPD_MAX_KNN);
    KPD_REQUIRE(ll->n_dst == batch->n_lig && kl->n_dst == batch->n_lig, "kpd_build_graph: n_dst mismatch");
    KPD_REQUIRE(!lk || lk->n_dst == batch->n_kp, "kpd_build_graph: lk n_dst mismatch");
    GraphWs g = carve_graph_ws(batch, workspace, nullptr);
    prof_begin(PROF_GRAPH, st);
    const size_t smem1 = (size_t)3 * batch->max_lig * sizeof(float) + (size_t)batch->max_lig * g.wk * sizeof(uint32_t);
    KPD_REQUIRE(smem1 <= 200 * 1024, "kpd_build_graph: complex too large for shared memory (%zu B)", smem1);
    if (smem1 > 48 * 1024)
        cudaFuncSetAttribute(graph_mark_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    graph_mark_kernel<<<batch->B, 128, smem1, st>>>(*batch, x_lig, x_kp, *p, g);
    KPD_TRY(check_launch("graph_mark_kernel"));
    graph_scan_kernel<<<1, 1024, 0, st>>>(batch->B, g, ll->rowptr, batch->n_lig, kl->rowptr,
                                          lk ? lk->rowptr : nullptr, batch->n_kp, counts_ll, counts_kl, edge_accum);
    KPD_TRY(check_launch("graph_scan_kernel"));
    const size_t smem3 = ((size_t)max(batch->max_lig, batch->max_kp) + 128) * sizeof(int);
    kpd_csr lk_v = lk ? *lk : kpd_csr{0, 0, nullptr, nullptr, nullptr};
    graph_fill_kernel<<<batch->B, 128, smem3, st>>>(*batch, g, *ll, *kl, lk_v, lk ? 1 : 0);
    KPD_TRY(check_launch("graph_fill_kernel"));
    prof_end(PROF_GRAPH, st);
    return 0;
}
