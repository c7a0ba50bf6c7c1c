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
constexpr int TC_KCS = (TCR / 8) * 128;   // bytes between k-chunks of the A tile
constexpr int TC_STAGES = 6;
constexpr int TC_SLAB_MAX = 2 * (256 / 8) * 128;   // one k-step of a 256-row weight
constexpr int TC_WG_MAX = 16 * 512;                // gates weight: 16 k-steps x (2 x 2 x 128 B)
constexpr uint32_t TC_TMEM_COLS = 512;
constexpr uint32_t TC_GATE_COL = 256;

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
    int kch;          // k-chunks (of 8 bf16) the A tile holds
};

struct TcCtx {
    uint32_t tmem;
    uint32_t it;       // slabs consumed so far (ring position)
    uint32_t ph_done, ph_g, ph_wg;
};

static size_t gvp_tc_smem_bytes(int kch) {
    return (size_t)kch * TC_KCS + 2 * sizeof(float) * TCR * VMAX * 3 + (size_t)TC_STAGES * TC_SLAB_MAX + TC_WG_MAX +
           (2 * TC_STAGES + 3) * sizeof(uint64_t) + 16 + 2 * sizeof(int) * TCR + 128;
}

__device__ __forceinline__ TcSm gvp_tc_carve(unsigned char* smem, int kch) {
    TcSm m;
    m.kch = kch;
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
    cx.it = 0; cx.ph_done = 0; cx.ph_g = 0; cx.ph_wg = 0;
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

// GVP.forward (models/gvp.py:89-116) on a 128-row tile; scalars in A (bf16 canonical), vectors in V (fp32).
__device__ void gvp_tile_tc(const GvpW& g, TcSm& m, TcCtx& cx) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NBf = (g.fout + 15) & ~15;
    const int ksf = (g.fin + g.hd + 15) >> 4;
    const int ksg = NBf >> 4;
    // 0. gates weight -> smem (one bulk copy, overlaps the vector work below)
    if (tid == 0) {
        tc::mbar_arrive_expect_tx(m.wgbar, ksg * 512);
        tc::bulk_g2s(m.Wg, g.WgP, ksg * 512, m.wgbar);
    }
    // a. Vh = V^T Wh ; sh = |Vh| -> A[:, fin + h]   (gvp.py:96, :99)
    for (int idx = tid; idx < TCR * g.hd; idx += blockDim.x) {
        const int r = idx / g.hd, hh = idx - r * g.hd;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        const float* v = m.V + r * (VMAX * 3);
        for (int k = 0; k < g.vin; ++k) {
            const float w = g.Wh[k * g.hd + hh];
            a0 = fmaf(v[3 * k], w, a0); a1 = fmaf(v[3 * k + 1], w, a1); a2 = fmaf(v[3 * k + 2], w, a2);
        }
        float* o = m.Vh + r * (VMAX * 3) + 3 * hh;
        o[0] = a0; o[1] = a1; o[2] = a2;
        *reinterpret_cast<__nv_bfloat16*>(m.A + tc::canon_off(r, g.fin + hh, TC_KCS)) =
            __float2bfloat16(sqrtf(fmaxf(a0 * a0 + a1 * a1 + a2 * a2, 1e-8f)));
    }
    __syncthreads();
    // b. Vu = Vh^T Wu -> V   (gvp.py:97)
    for (int idx = tid; idx < TCR * g.vout; idx += blockDim.x) {
        const int r = idx / g.vout, u = idx - r * g.vout;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        const float* vh = m.Vh + r * (VMAX * 3);
        for (int k = 0; k < g.hd; ++k) {
            const float w = g.Wu[k * g.vout + u];
            a0 = fmaf(vh[3 * k], w, a0); a1 = fmaf(vh[3 * k + 1], w, a1); a2 = fmaf(vh[3 * k + 2], w, a2);
        }
        float* o = m.V + r * (VMAX * 3) + 3 * u;
        o[0] = a0; o[1] = a1; o[2] = a2;
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    // c. feats GEMM: D[128 x NBf] = A[128 x 16*ksf] * Wf^T, weights through the slab ring
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(TCR, NBf);
        const int b_kstride = (NBf / 8) * 128;
        const uint32_t slab = 2 * b_kstride;
        int issued = 0;
        for (int j = 0; j < ksf; ++j) {
            while (issued < ksf && issued - j < TC_STAGES) {
                const uint32_t L = cx.it + issued, st = L % TC_STAGES;
                if (L >= TC_STAGES) tc::mbar_wait(&m.empty[st], ((L / TC_STAGES) - 1) & 1);
                tc::mbar_arrive_expect_tx(&m.full[st], slab);
                tc::bulk_g2s(m.ring + (size_t)st * TC_SLAB_MAX, g.WfP + (size_t)issued * (slab / 16), slab, &m.full[st]);
                ++issued;
            }
            const uint32_t Mi = cx.it + j, st = Mi % TC_STAGES;
            tc::mbar_wait(&m.full[st], (Mi / TC_STAGES) & 1);
            tc::fence_after_sync();
            const uint64_t ad = tc::make_smem_desc(tc::smem_u32(m.A + (size_t)2 * j * TC_KCS), TC_KCS, 128);
            const uint64_t bd = tc::make_smem_desc(tc::smem_u32(m.ring + (size_t)st * TC_SLAB_MAX), b_kstride, 128);
            tc::mma_bf16_ss(cx.tmem, ad, bd, idesc, j > 0 ? 1u : 0u);
            tc::mma_commit(&m.empty[st]);
        }
        tc::mma_commit(m.done);
    }
    cx.it += ksf;
    __syncwarp();
    tc::mbar_wait(m.done, cx.ph_done);
    cx.ph_done ^= 1;
    tc::fence_after_sync();
    // d. epilogue 1: feats_out = SiLU(acc + b) -> bf16 -> A[:, 0:fout]   (gvp.py:103)
    {
        const int q = warp & 3, hf = warp >> 2, row = 32 * q + lane;
        const uint32_t taddr = cx.tmem + ((uint32_t)(32 * q) << 16);
        const int cend = min(NBf, hf * 128 + 128);
        for (int c0 = hf * 128; c0 < cend; c0 += 32) {
            uint32_t v[32];
            tc::tmem_ld_x32(taddr + c0, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int col = c0 + 8 * kc + e;
                    f[e] = col < g.fout ? silu_f(__uint_as_float(v[8 * kc + e]) + g.bf[col]) : 0.0f;
                }
                if (c0 + 8 * kc < NBf) {
                    uint4 pk;
                    pk.x = tc::pack_bf16x2(f[0], f[1]); pk.y = tc::pack_bf16x2(f[2], f[3]);
                    pk.z = tc::pack_bf16x2(f[4], f[5]); pk.w = tc::pack_bf16x2(f[6], f[7]);
                    *reinterpret_cast<uint4*>(m.A + (size_t)((c0 >> 3) + kc) * TC_KCS + (row >> 3) * 128 + (row & 7) * 16) = pk;
                }
            }
        }
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
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
    }
    cx.ph_wg ^= 1;
    __syncwarp();
    tc::mbar_wait(m.gdone, cx.ph_g);
    cx.ph_g ^= 1;
    tc::fence_after_sync();
    // f. epilogue 2: vectors_out = act(gating) * Vu   (gvp.py:105-111)
    if (warp < 4) {
        const int row = 32 * warp + lane;
        uint32_t v[16];
        tc::tmem_ld_x16(cx.tmem + ((uint32_t)(32 * warp) << 16) + TC_GATE_COL, v);
        tc::tmem_ld_wait();
        float* o = m.V + row * (VMAX * 3);
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (u < g.vout) {
                float a = __uint_as_float(v[u]) + g.bg[u];
                if (g.sigmoid_gate) a = sigmoid_f(a);
                o[3 * u] *= a; o[3 * u + 1] *= a; o[3 * u + 2] *= a;
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
}

// segmented reduction of column `col` of the bf16 canonical tile (see seg_reduce_column)
__device__ __forceinline__ void seg_reduce_column_canon(const unsigned char* A, int col, int n, const int* dst_s,
                                                        const int* rowptr, int tile_begin, const SegOut& o, int out_col) {
    int i = 0;
    while (i < n) {
        const int d = dst_s[i];
        float s = 0.0f;
        int j = i;
        while (j < n && dst_s[j] == d) { s += bf16_at(A, j, col); ++j; }
        const bool from_prev = (i == 0) && (rowptr[d] < tile_begin);
        const bool into_next = (j == n) && (rowptr[d + 1] > tile_begin + n);
        if (from_prev) o.part0[out_col] = s;
        else if (into_next) o.part1[out_col] = s;
        else o.out[(size_t)d * o.ld_out + out_col] = s;
        i = j;
    }
}

__global__ void __launch_bounds__(NT, 1) gvp_edge_tc_kernel(const GvpEdgeLaunch L) {
    const GvpEtypeArgs& a = L.e[blockIdx.y];
    const int E = a.rowptr[a.n_dst];
    const int tile_begin = blockIdx.x * TCR;
    if (tile_begin >= E) return;
    const int n = min(TCR, E - tile_begin);
    extern __shared__ __align__(128) unsigned char smem_tc[];
    TcSm m = gvp_tc_carve(smem_tc, L.kch);
    TcCtx cx = gvp_tc_setup(m);
    const int tid = threadIdx.x;
    const int Sd = L.Sdim, Vd = L.Vdim;

    if (tid < TCR) {
        const int e = tile_begin + min(tid, n - 1);
        const int s = a.src[e], d = a.dst[e];
        m.src_s[tid] = s;
        m.dst_s[tid] = d;
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
    __syncthreads();
    // gather s_src -> bf16 canonical (consecutive lanes = consecutive rows: conflict-free 16 B stores)
    for (int idx = tid; idx < TCR * (Sd >> 3); idx += blockDim.x) {
        const int r = idx & (TCR - 1), kc = idx >> 7;
        const float4* sp = reinterpret_cast<const float4*>(a.s_src + (size_t)m.src_s[r] * Sd + 8 * kc);
        const float4 u = sp[0], w = sp[1];
        uint4 pk;
        pk.x = tc::pack_bf16x2(u.x, u.y); pk.y = tc::pack_bf16x2(u.z, u.w);
        pk.z = tc::pack_bf16x2(w.x, w.y); pk.w = tc::pack_bf16x2(w.z, w.w);
        *reinterpret_cast<uint4*>(m.A + (size_t)kc * TC_KCS + (r >> 3) * 128 + (r & 7) * 16) = pk;
    }
    for (int idx = tid; idx < TCR * Vd * 3; idx += blockDim.x) {
        const int r = idx / (Vd * 3), k = idx - r * (Vd * 3);
        m.V[r * (VMAX * 3) + 3 + k] = a.v_src[(size_t)m.src_s[r] * (Vd * 3) + k];
    }
    __syncthreads();
    for (int i = 0; i < L.n_msg; ++i) gvp_tile_tc(a.msg[i], m, cx);

    SegOut o;
    o.part0 = a.part + ((size_t)blockIdx.x * 2 + 0) * L.pw;
    o.part1 = a.part + ((size_t)blockIdx.x * 2 + 1) * L.pw;
    o.out = a.sm; o.ld_out = Sd;
    for (int col = tid; col < Sd; col += blockDim.x)
        seg_reduce_column_canon(m.A, col, n, m.dst_s, a.rowptr, tile_begin, o, col);
    if (tid < Vd * 3) {
        SegOut ov = o;
        ov.out = a.vm; ov.ld_out = Vd * 3;
        ov.part0 += Sd; ov.part1 += Sd;
        seg_reduce_column(m.V, VMAX * 3, tid, n, m.dst_s, a.rowptr, tile_begin, ov, tid);
    }
    gvp_tc_teardown(m, cx);
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

__global__ void __launch_bounds__(NT, 1) gvp_node_tc_kernel(const GvpNodeLaunch L) {
    const GvpNodeArgs& a = L.nt[blockIdx.y];
    const int n0 = blockIdx.x * TCR;
    if (n0 >= a.n) return;
    const int n = min(TCR, a.n - n0);
    extern __shared__ __align__(128) unsigned char smem_tc[];
    TcSm m = gvp_tc_carve(smem_tc, a.kch);
    TcCtx cx = gvp_tc_setup(m);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Sd = a.Sdim, Vd = a.Vdim;
    constexpr int PER = 8;   // Sd <= 256

    // ---- phase 1: features + messages / norm, GVPLayerNorm; residual stash in global; bf16 tile
    for (int r = warp; r < TCR; r += NT / 32) {
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
    for (int i = 0; i < a.n_upd; ++i) gvp_tile_tc(a.upd[i], m, cx);
    // ---- phase 3: residual + GVPLayerNorm -> global
    for (int r = warp; r < n; r += NT / 32) {
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

__global__ void __launch_bounds__(NT, 1) gvp_head_tc_kernel(const GvpHeadArgs a) {
    const int n0 = blockIdx.x * TCR;
    const int n = min(TCR, a.n - n0);
    extern __shared__ __align__(128) unsigned char smem_tc[];
    TcSm m = gvp_tc_carve(smem_tc, a.kch);
    TcCtx cx = gvp_tc_setup(m);
    const int tid = threadIdx.x;
    const int Sd = a.Sdim, Vd = a.Vdim;
    for (int idx = tid; idx < TCR * (Sd >> 3); idx += blockDim.x) {
        const int r = idx & (TCR - 1), kc = idx >> 7;
        const int nd = n0 + min(r, n - 1);
        const float4* sp = reinterpret_cast<const float4*>(a.s + (size_t)nd * Sd + 8 * kc);
        const float4 u = sp[0], w = sp[1];
        uint4 pk;
        pk.x = tc::pack_bf16x2(u.x, u.y); pk.y = tc::pack_bf16x2(u.z, u.w);
        pk.z = tc::pack_bf16x2(w.x, w.y); pk.w = tc::pack_bf16x2(w.z, w.w);
        *reinterpret_cast<uint4*>(m.A + (size_t)kc * TC_KCS + (r >> 3) * 128 + (r & 7) * 16) = pk;
    }
    for (int idx = tid; idx < TCR * Vd * 3; idx += blockDim.x) {
        const int r = idx / (Vd * 3), k = idx - r * (Vd * 3);
        m.V[r * (VMAX * 3) + k] = a.v[(size_t)(n0 + min(r, n - 1)) * (Vd * 3) + k];
    }
    __syncthreads();
    for (int i = 0; i < a.n_gvps; ++i) gvp_tile_tc(a.g[i], m, cx);
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
