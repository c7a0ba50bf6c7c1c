// bf16 tensor-core variants of the GVP kernels (included inside namespace kpd by gvp.cu).
//
// Same algorithm as gvp_edge/node/head_kernel, but a tile is 128 rows and the scalar features of the
// tile live in shared memory ONLY as bf16 in the UMMA K-major canonical layout (tc.cuh), i.e. directly
// as the A operand of tcgen05.mma:
//   feats_out = SiLU(Linear(cat(feats, sh)))   -> tcgen05.mma  [128 x K] x [K x 256], fp32 accumulators in
//                                                 TMEM columns [0,256); weights arrive as packed k-step
//                                                 slabs through a cp.async.bulk + mbarrier ring
//   gating    = Linear(feats_out)               -> tcgen05.mma  [128 x 256] x [256 x 16] into TMEM columns
//                                                 [256,272) (the A operand is what epilogue 1 just wrote)
// Epilogues read TMEM with tcgen05.ld (one accumulator row per thread).  The small vector einsums,
// LayerNorms, gathers and the segmented reduction stay on the SIMT pipes in fp32.
// This is the "bf16 GEMM mode" the north star asks to state separately: operands are rounded to bf16,
// accumulation is fp32; outputs differ from the fp32 mode at the 1e-3..1e-2 level.

constexpr int TCR = 128;              // rows per tile
constexpr int TC_NODE_ROWS = 32;      // valid rows per node/head tile (MMA still M = 128 over zero rows):
                                      // node counts are small, so more, lighter CTAs fill the GPU
constexpr int NT_TC = 512;            // 16 warps: the SIMT phases of these kernels are latency-bound
constexpr int TC_WSM = 2 * VMAX * VMAX + 256 + 32;   // floats: Wh, Wu, feats bias, gate bias staged per GVP

// bf16 mode does not need fp32-faithful transcendentals
// one MUFU op per element: silu(x) = x * sigmoid(x) = h + h * tanh(h), sigmoid(x) = 0.5 + 0.5 * tanh(h), h = x / 2
// (the exp + reciprocal form costs two MUFU ops and made the SiLU epilogue SFU-bound)
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float silu_fast(float x) { const float h = 0.5f * x; return fmaf(h, tanh_fast(h), h); }
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
constexpr int TC_KCS = (TCR / 8) * 128;   // bytes between k-chunks of the A tile
constexpr int TC_STAGES = 8;
constexpr int TC_SLAB_MAX = 2 * (256 / 8) * 128;   // one k-step of a 256-row weight
constexpr int TC_WG_MAX = 16 * 512;                // gates weight: 16 k-steps x (2 x 2 x 128 B)
constexpr uint32_t TC_TMEM_COLS = 512;
constexpr uint32_t TC_GATE_COL = 256;

// Debug phase timers (cycles, summed over CTAs by thread 0): read with kpd_debug_tc_times().
// [0..7] gvp_tile_tc phases: stage weights, Vh, Vu+fence, feats GEMM, epilogue 1, gates GEMM, epilogue 2, calls
// [8..13] edge kernel: setup, geometry+gather, GVP chain, segmented reduce, teardown, CTAs
__device__ unsigned long long g_tc_times[16];
#define TC_T(var) const long long var = clock64()
#define TC_ACC(slot, a, b) do { if (threadIdx.x == 0) atomicAdd(&g_tc_times[slot], (unsigned long long)((b) - (a))); } while (0)

struct TcSm {
    unsigned char* A;
    float* V;
    float* Vh;
    unsigned char* ring;
    unsigned char* Wg;
    uint64_t* full;
    uint64_t* empty;
    uint64_t* done;
    uint64_t* gdone;
    uint64_t* wgbar;
    uint32_t* tmem_slot;
    int* src_s;
    int* dst_s;
    float* wsm;       // [TC_WSM] per-GVP small weights: Wh | Wu | bf | bg
    int* seg;         // [TCR + 4] segment starts of the dst-sorted tile (+ count at seg[TCR + 1])
    int* rp;          // [2 * TCR] rowptr[dst], rowptr[dst + 1] per row
    int kch;          // k-chunks (of 8 bf16) the A tile holds
    int rows;         // valid rows of this tile (multiple of 32, <= TCR); the rest of A stays zero
};

struct TcCtx {
    uint32_t tmem;
    uint32_t it;       // slabs consumed so far (ring position)
    uint32_t pre;      // slabs of the NEXT feats GEMM already in flight (issued by thread 0)
    uint32_t ph_done, ph_g, ph_wg;
};

static size_t gvp_tc_smem_bytes(int kch) {
    return (size_t)kch * TC_KCS + 2 * sizeof(float) * TCR * VMAX * 3 + (size_t)TC_STAGES * TC_SLAB_MAX + TC_WG_MAX +
           (2 * TC_STAGES + 3) * sizeof(uint64_t) + 16 + 2 * sizeof(int) * TCR + sizeof(float) * TC_WSM +
           sizeof(int) * (3 * TCR + 8) + 128;
}

__device__ __forceinline__ TcSm gvp_tc_carve(unsigned char* smem, int kch, int rows = TCR) {
    TcSm m;
    m.kch = kch;
    m.rows = rows;
    m.A = smem;
    m.V = reinterpret_cast<float*>(m.A + (size_t)kch * TC_KCS);
    m.Vh = m.V + TCR * VMAX * 3;
    m.ring = reinterpret_cast<unsigned char*>(m.Vh + TCR * VMAX * 3);
    m.Wg = m.ring + TC_STAGES * TC_SLAB_MAX;
    m.full = reinterpret_cast<uint64_t*>(m.Wg + TC_WG_MAX);
    m.empty = m.full + TC_STAGES;
    m.done = m.empty + TC_STAGES;
    m.gdone = m.done + 1;
    m.wgbar = m.gdone + 1;
    m.tmem_slot = reinterpret_cast<uint32_t*>(m.wgbar + 1);
    m.src_s = reinterpret_cast<int*>(m.tmem_slot + 4);
    m.dst_s = m.src_s + TCR;
    m.wsm = reinterpret_cast<float*>(m.dst_s + TCR);
    m.seg = reinterpret_cast<int*>(m.wsm + TC_WSM);
    m.rp = m.seg + TCR + 8;
    return m;
}

// barriers + TMEM + a zeroed A tile; ends with a __syncthreads()
__device__ __forceinline__ TcCtx gvp_tc_setup(TcSm& m) {
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < TC_STAGES; ++i) { tc::mbar_init(&m.full[i], 1); tc::mbar_init(&m.empty[i], 1); }
        tc::mbar_init(m.done, 1);
        tc::mbar_init(m.gdone, 1);
        tc::mbar_init(m.wgbar, 1);
        tc::fence_barrier_init();
    }
    if ((tid >> 5) == 0) { tc::tmem_alloc(m.tmem_slot, TC_TMEM_COLS); tc::tmem_relinquish(); }
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < m.kch * (TC_KCS / 16); i += blockDim.x) reinterpret_cast<uint4*>(m.A)[i] = z;
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    TcCtx cx;
    cx.tmem = *m.tmem_slot;
    cx.it = 0; cx.pre = 0; cx.ph_done = 0; cx.ph_g = 0; cx.ph_wg = 0;
    return cx;
}

__device__ __forceinline__ void gvp_tc_teardown(TcSm& m, TcCtx& cx) {
    tc::fence_before_sync();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc(cx.tmem, TC_TMEM_COLS);
}

__device__ __forceinline__ float bf16_at(const unsigned char* A, int r, int k) {
    return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(A + tc::canon_off(r, k, TC_KCS)));
}

constexpr int TC_PRODUCER = 32;   // thread that feeds the weight ring (lane 0 of warp 1); thread 0 issues the MMAs

// the few fields of the NEXT GVP the producer needs for prefetching (passed by value: taking the address of
// a kernel-parameter element would force the whole parameter struct into local memory)
struct TcNext { const uint4* WfP; int ksf; int NBf; };
__device__ __forceinline__ TcNext tc_next(const GvpW& g) {
    TcNext n; n.WfP = g.WfP; n.ksf = (g.fin + g.hd + 15) >> 4; n.NBf = (g.fout + 15) & ~15; return n;
}
__device__ __forceinline__ TcNext tc_no_next() { TcNext n; n.WfP = nullptr; n.ksf = 0; n.NBf = 0; return n; }

// producer thread: put slabs [from, to) of a feats weight in flight; ring positions start at it0
__device__ __forceinline__ void gvp_tc_produce(const uint4* WfP, int NBf, TcSm& m, uint32_t it0, int from, int to) {
    const uint32_t slab = 2 * (NBf / 8) * 128;
    for (int i = from; i < to; ++i) {
        const uint32_t L = it0 + i, st = L % TC_STAGES;
        if (L >= TC_STAGES) tc::mbar_wait(&m.empty[st], ((L / TC_STAGES) - 1) & 1);   // previous occupant consumed
        tc::mbar_arrive_expect_tx(&m.full[st], slab);
        tc::bulk_g2s(m.ring + (size_t)st * TC_SLAB_MAX, WfP + (size_t)i * (slab / 16), slab, &m.full[st]);
    }
}

// producer thread: first slabs of a GVP (at ring position it0) so that their L2 latency overlaps SIMT work
__device__ __forceinline__ uint32_t gvp_tc_prefetch(const TcNext& nx, TcSm& m, uint32_t it0) {
    const int n = nx.ksf < TC_STAGES ? nx.ksf : TC_STAGES;
    gvp_tc_produce(nx.WfP, nx.NBf, m, it0, 0, n);
    return (uint32_t)n;
}

// one 32-column chunk of epilogue 1: bias + SiLU -> bf16 -> canonical A tile (static indexing only)
__device__ __forceinline__ void epi1_chunk(const uint32_t (&v)[32], int c0, int fout, int NBf, const float* bf_s,
                                           unsigned char* A, int row) {
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int col = c0 + 8 * kc + e;
            f[e] = col < fout ? silu_fast(__uint_as_float(v[8 * kc + e]) + bf_s[col]) : 0.0f;
        }
        if (c0 + 8 * kc < NBf) {
            uint4 pk;
            pk.x = tc::pack_bf16x2(f[0], f[1]); pk.y = tc::pack_bf16x2(f[2], f[3]);
            pk.z = tc::pack_bf16x2(f[4], f[5]); pk.w = tc::pack_bf16x2(f[6], f[7]);
            *reinterpret_cast<uint4*>(A + (size_t)((c0 >> 3) + kc) * TC_KCS + (row >> 3) * 128 + (row & 7) * 16) = pk;
        }
    }
}

// GVP.forward (models/gvp.py:89-116) on a 128-row tile; scalars in A (bf16 canonical), vectors in V (fp32).
// `next`: the GVP that will run after this one on the same tile (its first weight slabs are prefetched as soon
// as this GVP's MMAs have been issued), or nullptr.
__device__ __forceinline__ void gvp_tile_tc(const GvpW& g, TcSm& m, TcCtx& cx, const TcNext next) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NBf = (g.fout + 15) & ~15;
    const int ksf = (g.fin + g.hd + 15) >> 4;
    const int ksg = NBf >> 4;
    TC_T(t0);
    // 0. gates weight -> smem (one bulk copy, overlaps the vector work below); small weights -> smem
    if (tid == 0) {
        tc::mbar_arrive_expect_tx(m.wgbar, ksg * 512);
        tc::bulk_g2s(m.Wg, g.WgP, ksg * 512, m.wgbar);
    }
    float* Wh_s = m.wsm;
    float* Wu_s = Wh_s + VMAX * VMAX;
    float* bf_s = Wu_s + VMAX * VMAX;
    float* bg_s = bf_s + 256;
    for (int i = tid; i < g.vin * g.hd; i += blockDim.x) Wh_s[i] = g.Wh[i];
    for (int i = tid; i < g.hd * g.vout; i += blockDim.x) Wu_s[i] = g.Wu[i];
    for (int i = tid; i < g.fout; i += blockDim.x) bf_s[i] = g.bf[i];
    if (tid < g.vout) bg_s[tid] = g.bg[tid];
    __syncthreads();
    TC_T(t1);
    // a. Vh = V^T Wh ; sh = |Vh| -> A[:, fin + h]   (gvp.py:96, :99)
    for (int idx = tid; idx < m.rows * g.hd; idx += blockDim.x) {
        const int r = idx / g.hd, hh = idx - r * g.hd;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        const float* v = m.V + r * (VMAX * 3);
        for (int k = 0; k < g.vin; ++k) {
            const float w = Wh_s[k * g.hd + hh];
            a0 = fmaf(v[3 * k], w, a0); a1 = fmaf(v[3 * k + 1], w, a1); a2 = fmaf(v[3 * k + 2], w, a2);
        }
        float* o = m.Vh + r * (VMAX * 3) + 3 * hh;
        o[0] = a0; o[1] = a1; o[2] = a2;
        *reinterpret_cast<__nv_bfloat16*>(m.A + tc::canon_off(r, g.fin + hh, TC_KCS)) =
            __float2bfloat16(sqrtf(fmaxf(a0 * a0 + a1 * a1 + a2 * a2, 1e-8f)));
    }
    __syncthreads();
    TC_T(t2);
    // b. Vu = Vh^T Wu -> V   (gvp.py:97)
    for (int idx = tid; idx < m.rows * g.vout; idx += blockDim.x) {
        const int r = idx / g.vout, u = idx - r * g.vout;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        const float* vh = m.Vh + r * (VMAX * 3);
        for (int k = 0; k < g.hd; ++k) {
            const float w = Wu_s[k * g.vout + u];
            a0 = fmaf(vh[3 * k], w, a0); a1 = fmaf(vh[3 * k + 1], w, a1); a2 = fmaf(vh[3 * k + 2], w, a2);
        }
        float* o = m.V + r * (VMAX * 3) + 3 * u;
        o[0] = a0; o[1] = a1; o[2] = a2;
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    TC_T(t3);
    // c. feats GEMM: D[128 x NBf] = A[128 x 16*ksf] * Wf^T, weights through the slab ring
    if (tid == TC_PRODUCER) {
        // remaining slabs of this GEMM, then the first slabs of the next GVP (behind this GEMM's MMAs)
        gvp_tc_produce(g.WfP, NBf, m, cx.it, (int)cx.pre, ksf);
        cx.pre = next.WfP ? gvp_tc_prefetch(next, m, cx.it + ksf) : 0u;
    }
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(TCR, NBf);
        const int b_kstride = (NBf / 8) * 128;
        for (int j = 0; j < ksf; ++j) {
            const uint32_t Mi = cx.it + j, st = Mi % TC_STAGES;
            tc::mbar_wait(&m.full[st], (Mi / TC_STAGES) & 1);
            tc::fence_after_sync();
            const uint64_t ad = tc::make_smem_desc(tc::smem_u32(m.A + (size_t)2 * j * TC_KCS), TC_KCS, 128);
            const uint64_t bd = tc::make_smem_desc(tc::smem_u32(m.ring + (size_t)st * TC_SLAB_MAX), b_kstride, 128);
            tc::mma_bf16_ss(cx.tmem, ad, bd, idesc, j > 0 ? 1u : 0u);
            tc::mma_commit(&m.empty[st]);
        }
        tc::mma_commit(m.done);
        tc::mbar_wait(m.done, cx.ph_done);      // only the issuing thread polls the mbarrier ...
    }
    cx.it += ksf;
    cx.ph_done ^= 1;
    __syncthreads();                            // ... everyone else sleeps at the block barrier
    tc::fence_after_sync();
    TC_T(t4);
    // d. epilogue 1: feats_out = SiLU(acc + b) -> bf16 -> A[:, 0:fout]   (gvp.py:103)
    {
        const int q = warp & 3, cq = warp >> 2, nq = blockDim.x >> 7, row = 32 * q + lane;
        const int cpw = 256 / nq;                       // columns per warp group
        const uint32_t taddr = cx.tmem + ((uint32_t)(32 * q) << 16);
        const int cend = 32 * q < m.rows ? min(NBf, cq * cpw + cpw) : 0;   // warps of empty row quarters skip
        // two 32-column TMEM loads in flight per wait (cpw is 64 with 16 warps)
        for (int cb = cq * cpw; cb < cend; cb += 64) {
            uint32_t v0[32], v1[32];
            const bool two = cb + 32 < cend;
            tc::tmem_ld_x32(taddr + cb, v0);
            if (two) tc::tmem_ld_x32(taddr + cb + 32, v1);
            tc::tmem_ld_wait();
            epi1_chunk(v0, cb, g.fout, NBf, bf_s, m.A, row);
            if (two) epi1_chunk(v1, cb + 32, g.fout, NBf, bf_s, m.A, row);
        }
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    TC_T(t5);
    // e. gates GEMM: G[128 x 16] = feats_out[128 x NBf] * Wg^T -> TMEM columns [256, 272)
    if (tid == 0) {
        tc::mbar_wait(m.wgbar, cx.ph_wg);
        tc::fence_after_sync();
        const uint32_t idesc = tc::make_idesc_bf16(TCR, 16);
        for (int j = 0; j < ksg; ++j) {
            const uint64_t ad = tc::make_smem_desc(tc::smem_u32(m.A + (size_t)2 * j * TC_KCS), TC_KCS, 128);
            const uint64_t bd = tc::make_smem_desc(tc::smem_u32(m.Wg + (size_t)j * 512), 256, 128);
            tc::mma_bf16_ss(cx.tmem + TC_GATE_COL, ad, bd, idesc, j > 0 ? 1u : 0u);
        }
        tc::mma_commit(m.gdone);
        tc::mbar_wait(m.gdone, cx.ph_g);
    }
    cx.ph_wg ^= 1;
    cx.ph_g ^= 1;
    __syncthreads();
    tc::fence_after_sync();
    TC_T(t6);
    // f. epilogue 2: vectors_out = act(gating) * Vu   (gvp.py:105-111)
    if (warp < 4 && 32 * warp < m.rows) {
        const int row = 32 * warp + lane;
        uint32_t v[16];
        tc::tmem_ld_x16(cx.tmem + ((uint32_t)(32 * warp) << 16) + TC_GATE_COL, v);
        tc::tmem_ld_wait();
        float* o = m.V + row * (VMAX * 3);
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (u < g.vout) {
                float a = __uint_as_float(v[u]) + bg_s[u];
                if (g.sigmoid_gate) a = sigmoid_fast(a);
                o[3 * u] *= a; o[3 * u + 1] *= a; o[3 * u + 2] *= a;
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    TC_T(t7);
    TC_ACC(0, t0, t1); TC_ACC(1, t1, t2); TC_ACC(2, t2, t3); TC_ACC(3, t3, t4); TC_ACC(4, t4, t5); TC_ACC(5, t5, t6);
    TC_ACC(6, t6, t7); TC_ACC(7, 0, 1);
}

// Segment table of a dst-sorted tile: seg[0..nseg) = first row of every run of equal dst, seg[nseg] = n;
// the count goes to seg[TCR + 1].  Threads 0..TCR-1 take part (4 warps); ends with a __syncthreads().
__device__ __forceinline__ void build_segments(TcSm& m, int n) {
    __shared__ int warp_cnt[TCR / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bool start = false;
    if (tid < TCR) {
        start = tid < n && (tid == 0 || m.dst_s[tid] != m.dst_s[tid - 1]);
        const unsigned bal = __ballot_sync(0xffffffffu, start);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
    }
    __syncthreads();
    if (tid < TCR) {
        const unsigned bal = __ballot_sync(0xffffffffu, start);
        int base = 0;
        for (int w = 0; w < warp; ++w) base += warp_cnt[w];
        if (start) m.seg[base + __popc(bal & ((1u << lane) - 1u))] = tid;
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < TCR / 32; ++w) tot += warp_cnt[w];
            m.seg[tot] = n;
            m.seg[TCR + 1] = tot;
        }
    }
    __syncthreads();
}

// where the sum of segment [a,b) of this tile goes (see seg_reduce_column in common.cuh)
__device__ __forceinline__ float* seg_target(const TcSm& m, int a, int b, int n, int tile_begin, const SegOut& o, int out_col) {
    const bool from_prev = (a == 0) && (m.rp[2 * a] < tile_begin);
    const bool into_next = (b == n) && (m.rp[2 * a + 1] > tile_begin + n);
    if (from_prev) return o.part0 + out_col;
    if (into_next) return o.part1 + out_col;
    return o.out + (size_t)m.dst_s[a] * o.ld_out + out_col;
}

__global__ void __launch_bounds__(NT_TC, 1) gvp_edge_tc_kernel(const GvpEdgeLaunch L) {
    const GvpEtypeArgs& a = L.e[blockIdx.y];
    const int E = a.rowptr[a.n_dst];
    const int tile_begin = blockIdx.x * TCR;
    if (tile_begin >= E) return;
    const int n = min(TCR, E - tile_begin);
    extern __shared__ __align__(128) unsigned char smem_tc[];
    TC_T(e0);
    TcSm m = gvp_tc_carve(smem_tc, L.kch);
    TcCtx cx = gvp_tc_setup(m);
    const int tid = threadIdx.x;
    const int Sd = L.Sdim, Vd = L.Vdim;
    TC_T(e1);

    if (tid < TCR) {
        const int e = tile_begin + min(tid, n - 1);
        const int s = a.src[e], d = a.dst[e];
        m.src_s[tid] = s;
        m.dst_s[tid] = d;
        m.rp[2 * tid] = a.rowptr[d];
        m.rp[2 * tid + 1] = a.rowptr[d + 1];
        const float dx = a.xs[3 * s] - a.xd[3 * d], dy = a.xs[3 * s + 1] - a.xd[3 * d + 1],
                    dz = a.xs[3 * s + 2] - a.xd[3 * d + 2];
        const float dij = sqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1e-8f)) + 1e-8f;
        float* v0 = m.V + tid * (VMAX * 3);
        v0[0] = dx / dij; v0[1] = dy / dij; v0[2] = dz / dij;
        for (int k = 0; k < L.rbf_dim; ++k) {
            const float z = (dij - (float)k * L.rbf_step) / L.rbf_sigma;
            *reinterpret_cast<__nv_bfloat16*>(m.A + tc::canon_off(tid, Sd + k, TC_KCS)) = __float2bfloat16(expf(-(z * z)));
        }
    }
    if (tid == TC_PRODUCER) cx.pre = gvp_tc_prefetch(tc_next(a.msg[0]), m, cx.it);   // first GVP's weights start streaming now
    __syncthreads();
    build_segments(m, n);
    // gather s_src -> bf16 canonical (consecutive lanes = consecutive rows: conflict-free 16 B stores)
    {
        const int items = TCR * (Sd >> 3);
        for (int base = 0; base < items; base += 2 * blockDim.x) {
            float4 u[2], w[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {                 // all loads of the batch first (4 x 16 B in flight)
                const int idx = base + q * blockDim.x + tid;
                if (idx < items) {
                    const int r = idx & (TCR - 1), kc = idx >> 7;
                    const float4* sp = reinterpret_cast<const float4*>(a.s_src + (size_t)m.src_s[r] * Sd + 8 * kc);
                    u[q] = sp[0]; w[q] = sp[1];
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int idx = base + q * blockDim.x + tid;
                if (idx < items) {
                    const int r = idx & (TCR - 1), kc = idx >> 7;
                    uint4 pk;
                    pk.x = tc::pack_bf16x2(u[q].x, u[q].y); pk.y = tc::pack_bf16x2(u[q].z, u[q].w);
                    pk.z = tc::pack_bf16x2(w[q].x, w[q].y); pk.w = tc::pack_bf16x2(w[q].z, w[q].w);
                    *reinterpret_cast<uint4*>(m.A + (size_t)kc * TC_KCS + (r >> 3) * 128 + (r & 7) * 16) = pk;
                }
            }
        }
    }
    for (int idx = tid; idx < TCR * Vd * 3; idx += blockDim.x) {
        const int r = idx / (Vd * 3), k = idx - r * (Vd * 3);
        m.V[r * (VMAX * 3) + 3 + k] = a.v_src[(size_t)m.src_s[r] * (Vd * 3) + k];
    }
    __syncthreads();
    TC_T(e2);
    for (int i = 0; i < L.n_msg; ++i) gvp_tile_tc(a.msg[i], m, cx, i + 1 < L.n_msg ? tc_next(a.msg[i + 1]) : tc_no_next());
    TC_T(e3);

    SegOut o;
    o.part0 = a.part + ((size_t)blockIdx.x * 2 + 0) * L.pw;
    o.part1 = a.part + ((size_t)blockIdx.x * 2 + 1) * L.pw;
    o.out = a.sm; o.ld_out = Sd;
    const int nseg = m.seg[TCR + 1];
    for (int col = tid; col < Sd + Vd * 3; col += blockDim.x) {
        SegOut oc = o;
        int oc_col = col;
        if (col >= Sd) { oc.out = a.vm; oc.ld_out = Vd * 3; oc.part0 += Sd; oc.part1 += Sd; oc_col = col - Sd; }
        for (int sg = 0; sg < nseg; ++sg) {
            const int ra = m.seg[sg], rb = m.seg[sg + 1];
            float acc = 0.0f;
            if (col < Sd) {
                const unsigned char* p = m.A + (size_t)(col >> 3) * TC_KCS + (col & 7) * 2;
                for (int j = ra; j < rb; ++j)
                    acc += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(p + (j >> 3) * 128 + (j & 7) * 16));
            } else {
                for (int j = ra; j < rb; ++j) acc += m.V[j * (VMAX * 3) + oc_col];
            }
            *seg_target(m, ra, rb, n, tile_begin, oc, oc_col) = acc;
        }
    }
    __syncthreads();
    TC_T(e4);
    gvp_tc_teardown(m, cx);
    TC_T(e5);
    TC_ACC(8, e0, e1); TC_ACC(9, e1, e2); TC_ACC(10, e2, e3); TC_ACC(11, e3, e4); TC_ACC(12, e4, e5); TC_ACC(13, 0, 1);
}

// warp-level LayerNorm helpers for one row held as `per` values per lane (feature f = lane + 32*i)
template <int PER>
__device__ __forceinline__ void warp_layernorm(float (&x)[PER], int Sdim, const float* w, const float* b, int lane) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) if (lane + 32 * i < Sdim) s += x[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)Sdim;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) if (lane + 32 * i < Sdim) { const float d = x[i] - mean; v += d * d; }
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = 1.0f / sqrtf(v / (float)Sdim + 1e-5f);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int f = lane + 32 * i;
        if (f < Sdim) x[i] = (x[i] - mean) * rstd * w[f] + b[f];
    }
}

// vector part of GVPLayerNorm (gvp.py:163-165) for one row held in shared memory (pre-norm);
// returns vn = sqrt(mean_v clamp(|v|^2, 1e-8) + eps) + eps.  nv <= 32.
__device__ __forceinline__ float warp_vec_norm(const float* vrow, int nv, int lane) {
    float q = 0.f;
    if (lane < nv)
        q = fmaxf(vrow[3 * lane] * vrow[3 * lane] + vrow[3 * lane + 1] * vrow[3 * lane + 1] +
                  vrow[3 * lane + 2] * vrow[3 * lane + 2], 1e-8f);
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    return sqrtf(q / (float)nv + 1e-5f) + 1e-5f;
}

__global__ void __launch_bounds__(NT_TC, 1) gvp_node_tc_kernel(const GvpNodeLaunch L) {
    const GvpNodeArgs& a = L.nt[blockIdx.y];
    const int n0 = blockIdx.x * TC_NODE_ROWS;
    if (n0 >= a.n) return;
    const int n = min(TC_NODE_ROWS, a.n - n0);
    extern __shared__ __align__(128) unsigned char smem_tc[];
    TcSm m = gvp_tc_carve(smem_tc, a.kch, TC_NODE_ROWS);
    TcCtx cx = gvp_tc_setup(m);
    if (threadIdx.x == TC_PRODUCER) cx.pre = gvp_tc_prefetch(tc_next(a.upd[0]), m, cx.it);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Sd = a.Sdim, Vd = a.Vdim;
    constexpr int PER = 8;   // Sd <= 256

    // ---- phase 1: features + messages / norm, GVPLayerNorm; residual stash in global; bf16 tile
    for (int r = warp; r < TC_NODE_ROWS; r += NT_TC / 32) {
        const int nd = n0 + min(r, n - 1);
        int r0[2], r1[2];
        for (int e = 0; e < a.n_et; ++e) { r0[e] = a.rowptr[e][nd]; r1[e] = a.rowptr[e][nd + 1]; }
        float nv = a.norm_const;
        if (a.norm_mode == 1) nv = 1.0f;
        else if (a.norm_mode == 2) {
            const int b = a.node_batch[nd];
            const int p0 = a.ptr[b], p1 = a.ptr[b + 1];
            int tot = 0;
            for (int e = 0; e < a.n_et; ++e) tot += a.rowptr[e][p1] - a.rowptr[e][p0];
            nv = (float)tot / (float)(p1 - p0) + 1.0f;
        }
        float x[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int f = lane + 32 * i;
            x[i] = 0.f;
            if (f < Sd) {
                float msg = 0.f;
                for (int e = 0; e < a.n_et; ++e) {
                    float gsum = seg_gather(a.sm[e], Sd, a.part[e], a.pw, r0[e], r1[e], nd, f, a.edge_tile);
                    if (a.norm_mode == 1) gsum = gsum / (float)max(r1[e] - r0[e], 1);
                    msg += gsum;
                }
                x[i] = a.s[(size_t)nd * Sd + f] + msg / nv;
            }
        }
        warp_layernorm<PER>(x, Sd, a.mln_w, a.mln_b, lane);
        float* vrow = m.V + r * (VMAX * 3);
        for (int k = lane; k < 3 * Vd; k += 32) {
            float msg = 0.f;
            for (int e = 0; e < a.n_et; ++e) {
                float gsum = seg_gather(a.vm[e], 3 * Vd, a.part[e] + Sd, a.pw, r0[e], r1[e], nd, k, a.edge_tile);
                if (a.norm_mode == 1) gsum = gsum / (float)max(r1[e] - r0[e], 1);
                msg += gsum;
            }
            vrow[k] = a.v[(size_t)nd * (3 * Vd) + k] + msg / nv;
        }
        __syncwarp();
        const float vn = warp_vec_norm(vrow, Vd, lane);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int f = lane + 32 * i;
            if (f < Sd) {
                if (r < n) a.s[(size_t)(n0 + r) * Sd + f] = x[i];
                *reinterpret_cast<__nv_bfloat16*>(m.A + tc::canon_off(r, f, TC_KCS)) = __float2bfloat16(x[i]);
            }
        }
        for (int k = lane; k < 3 * Vd; k += 32) {
            const float val = vrow[k] / vn;
            vrow[k] = val;
            if (r < n) a.v[(size_t)(n0 + r) * (3 * Vd) + k] = val;
        }
    }
    __syncthreads();
    // ---- phase 2: update GVPs on the tensor cores
    for (int i = 0; i < a.n_upd; ++i) gvp_tile_tc(a.upd[i], m, cx, i + 1 < a.n_upd ? tc_next(a.upd[i + 1]) : tc_no_next());
    // ---- phase 3: residual + GVPLayerNorm -> global
    for (int r = warp; r < n; r += NT_TC / 32) {
        const int nd = n0 + r;
        float x[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int f = lane + 32 * i;
            x[i] = f < Sd ? bf16_at(m.A, r, f) + a.s[(size_t)nd * Sd + f] : 0.f;
        }
        warp_layernorm<PER>(x, Sd, a.uln_w, a.uln_b, lane);
        float* vrow = m.V + r * (VMAX * 3);
        for (int k = lane; k < 3 * Vd; k += 32) vrow[k] += a.v[(size_t)nd * (3 * Vd) + k];
        __syncwarp();
        const float vn = warp_vec_norm(vrow, Vd, lane);
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int f = lane + 32 * i;
            if (f < Sd) a.s[(size_t)nd * Sd + f] = x[i];
        }
        for (int k = lane; k < 3 * Vd; k += 32) a.v[(size_t)nd * (3 * Vd) + k] = vrow[k] / vn;
    }
    gvp_tc_teardown(m, cx);
}

__global__ void __launch_bounds__(NT_TC, 1) gvp_head_tc_kernel(const GvpHeadArgs a) {
    const int n0 = blockIdx.x * TC_NODE_ROWS;
    const int n = min(TC_NODE_ROWS, a.n - n0);
    extern __shared__ __align__(128) unsigned char smem_tc[];
    TcSm m = gvp_tc_carve(smem_tc, a.kch, TC_NODE_ROWS);
    TcCtx cx = gvp_tc_setup(m);
    if (threadIdx.x == TC_PRODUCER) cx.pre = gvp_tc_prefetch(tc_next(a.g[0]), m, cx.it);
    const int tid = threadIdx.x;
    const int Sd = a.Sdim, Vd = a.Vdim;
    for (int idx = tid; idx < TC_NODE_ROWS * (Sd >> 3); idx += blockDim.x) {
        const int r = idx % TC_NODE_ROWS, kc = idx / TC_NODE_ROWS;
        const int nd = n0 + min(r, n - 1);
        const float4* sp = reinterpret_cast<const float4*>(a.s + (size_t)nd * Sd + 8 * kc);
        const float4 u = sp[0], w = sp[1];
        uint4 pk;
        pk.x = tc::pack_bf16x2(u.x, u.y); pk.y = tc::pack_bf16x2(u.z, u.w);
        pk.z = tc::pack_bf16x2(w.x, w.y); pk.w = tc::pack_bf16x2(w.z, w.w);
        *reinterpret_cast<uint4*>(m.A + (size_t)kc * TC_KCS + (r >> 3) * 128 + (r & 7) * 16) = pk;
    }
    for (int idx = tid; idx < TC_NODE_ROWS * Vd * 3; idx += blockDim.x) {
        const int r = idx / (Vd * 3), k = idx - r * (Vd * 3);
        m.V[r * (VMAX * 3) + k] = a.v[(size_t)(n0 + min(r, n - 1)) * (Vd * 3) + k];
    }
    __syncthreads();
    for (int i = 0; i < a.n_gvps; ++i) gvp_tile_tc(a.g[i], m, cx, i + 1 < a.n_gvps ? tc_next(a.g[i + 1]) : tc_no_next());
    for (int idx = tid; idx < n * (a.F + 3); idx += blockDim.x) {
        const int r = idx / (a.F + 3), c = idx - r * (a.F + 3);
        if (c < a.F) {
            float s = a.bo[c];
            for (int k = 0; k < a.hid_out; ++k) s = fmaf(bf16_at(m.A, r, k), a.WoT[k * a.Fp + c], s);
            a.eps_h[(size_t)(n0 + r) * a.F + c] = s;
        } else {
            a.eps_x[(size_t)(n0 + r) * 3 + (c - a.F)] = m.V[r * (VMAX * 3) + (c - a.F)];
        }
    }
    gvp_tc_teardown(m, cx);
}
