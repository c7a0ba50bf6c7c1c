// Warp-specialised tensor-core EGNN edge kernel (included inside namespace kpd by egnn.cu).
//
// LigRecConv.message + aggregation (models/dynamics.py:89-122, :159-185) for all edge types in one launch, on
// tiles of 64 dst-sorted edges.  Per branch (edge_mlp, coord_mlp):
//   SIMT:   A = SiLU(P_src[src] + P_dst[dst] + w1c * dij)  -> bf16 (hi, lo) rows stacked into ONE 128-row UMMA
//           operand (ws_common.cuh); the few output columns beyond a multiple of 8 (H = 257 -> column 256) as
//           fp32 dot products on the way.  The kernel is bound by the latency of the P gathers (L2: 225 KB of shared
//           memory leave no L1) and by SIMT instruction count, so every warp fetches the (src, dst) of its four rows
//           itself at the top of the kernel, issues the gathers of branch 0 BEFORE the tile set-up, keeps them in flight
//           as row pairs in registers and fetches branch 1 behind the rows of branch 0 as those are consumed;
//   MMA:    acc[64 x 256] (+)= A W2^T, two tcgen05.mma per k-step (W_hi, W_lo) = all four hi/lo products, fp32 in
//           TMEM; weights streamed as packed k-step slabs through a cp.async.bulk ring;
//   SIMT:   m2 = SiLU(acc + b2) straight out of TMEM; the row dot products with the attention / coordinate vector;
//           edge branch: m2 -> shared memory, att = sigmoid(.), messages m2 * att summed per destination;
//           coord branch: cw = tanh(.) * range, x messages cw * x_diff / (dij + 1) summed per destination.
// Both branches have their own A buffer and TMEM columns, so the tensor core works on one while the SIMT warps
// build / drain the other.  Deterministic segmented reduction and output layout as in egnn_edge_kernel.
// phase timers (cycles, SIMT thread 0, summed over CTAs): set-up, indices + geometry, build A (edge), build A (coord),
// wait + epilogue (edge), reduce (edge), wait + epilogue (coord), reduce (coord), [8] = CTAs
__device__ unsigned long long g_eg_times[16];
#define EG_ACC(slot, a, b) do { if (threadIdx.x == 0) atomicAdd(&g_eg_times[slot], (unsigned long long)((b) - (a))); } while (0)

namespace egws {

using C = ws::Cfg<64, 2, 1, 8>;           // 64-edge tiles, (hi, lo) stacked: M = 128; 16 SIMT warps
constexpr int NCG = C::NCG;               // column groups of the epilogue (4 TMEM lane quarters x NCG warps)
constexpr int CPW = 256 / NCG;            // accumulator columns per epilogue warp
constexpr int R = C::R, NW = C::NW, NT_SIMT = C::NT_SIMT, NT = C::NT;
constexpr int VEC_LD = 264;                   // per-etype vectors staged in shared memory (floats; even, >= Hp)
static_assert(VEC_LD >= KPD_MAX_HIDDEN + 3 && VEC_LD % 2 == 0, "VEC_LD");

struct Sm {
    unsigned char* A[2];
    unsigned char* ring;
    float *w1c[2], *b2[2], *wv[2], *w2lo[2];     // [VEC_LD] each; w2lo: [3][VEC_LD]
    float *dij, *xsc, *lo, *dotp, *att, *xm;     // [R], [R][3], [2][R][4], [NCG][R], [R], [R][3]
    int *src_s, *dst_s, *seg, *rp, *warp_cnt;
    uint64_t *full, *empty, *a_ready, *acc_done;
    uint32_t* tmem_slot;
};

__host__ __device__ inline size_t a_bytes(int kch) { return ((size_t)kch * C::KCS + 127) & ~(size_t)127; }

static size_t smem_bytes(int kch) {
    return 2 * a_bytes(kch) + (size_t)C::STAGES * C::SLAB + sizeof(float) * (2 * 3 * VEC_LD + 2 * 3 * VEC_LD) +
           sizeof(float) * (R + 3 * R + 8 * R + NCG * R + R + 3 * R) + sizeof(int) * (5 * R + 16) +
           sizeof(uint64_t) * (2 * C::STAGES + 4) + 16 + 128;
}

__device__ __forceinline__ Sm carve(unsigned char* smem, int kch) {
    Sm m;
    m.A[0] = smem;
    m.A[1] = smem + a_bytes(kch);
    m.ring = smem + 2 * a_bytes(kch);
    float* f = reinterpret_cast<float*>(m.ring + C::STAGES * C::SLAB);
    for (int br = 0; br < 2; ++br) { m.w1c[br] = f; f += VEC_LD; m.b2[br] = f; f += VEC_LD; m.wv[br] = f; f += VEC_LD; }
    for (int br = 0; br < 2; ++br) { m.w2lo[br] = f; f += 3 * VEC_LD; }
    m.dij = f; f += R;
    m.xsc = f; f += 3 * R;
    m.lo = f; f += 8 * R;
    m.dotp = f; f += NCG * R;
    m.att = f; f += R;
    m.xm = f; f += 3 * R;
    m.src_s = reinterpret_cast<int*>(f);
    m.dst_s = m.src_s + R;
    m.seg = m.dst_s + R;
    m.rp = m.seg + R + 8;
    m.warp_cnt = m.rp + 2 * R;
    m.full = reinterpret_cast<uint64_t*>(m.warp_cnt + 8);
    m.empty = m.full + C::STAGES;
    m.a_ready = m.empty + C::STAGES;
    m.acc_done = m.a_ready + 2;
    m.tmem_slot = reinterpret_cast<uint32_t*>(m.acc_done + 2);
    return m;
}

__device__ __forceinline__ void simt_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NT_SIMT) : "memory"); }

__device__ __forceinline__ void publish(uint64_t* bar) {
    tc::fence_proxy_async();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) tc::mbar_arrive(bar);
}

}  // namespace egws

struct EgnnWsEtype {
    const uint4* W2P[2];        // packed (hi, lo) k-step slabs of W2[:nmain, :H] per branch (pack.pack_egnn_tc)
};
struct EgnnWsLaunch {
    EgnnEdgeLaunch L;
    EgnnWsEtype t[4];
    int kch;
    int tile_off[5];            // 1-D grid: CTAs [tile_off[e], tile_off[e+1]) are the tiles of edge type e (at capacity)
};

namespace egws {
// A pair of tile rows of first-layer products in flight (registers): the P_src chunks of both rows and the P_dst chunks --
// the second row's only when its destination differs from the first's (the tile is dst-sorted: 19 ll edges per atom)
struct RowPair { float4 u[2][2], w[2][2]; float ut[2], wt[2]; };
// issue the loads of rows 2 P and 2 P + 1 of this warp for branch br (lane = 8-feature chunk; ntail leftover features)
template <int P>
__device__ __forceinline__ void load_pair(const EgnnEtypeArgs& a, int Hp, int br, const int (&rs)[R / NW], const int (&rd)[R / NW],
                                          int lane, int nfull, int ntail, RowPair& q) {
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
        const int j = 2 * P + jj;
        const float* ps = a.Ps + (size_t)rs[j] * a.ldps + (a.slot_s + br) * Hp;
        const float* pd = a.Pd + (size_t)rd[j] * a.ldpd + (a.slot_d + br) * Hp;
        const bool need_d = jj == 0 || rd[j] != rd[jj == 0 ? j : j - 1];       // (warp-uniform)
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        q.u[jj][0] = q.u[jj][1] = q.w[jj][0] = q.w[jj][1] = z;
        if (lane < nfull) {
            q.u[jj][0] = __ldg(reinterpret_cast<const float4*>(ps + 8 * lane));
            q.u[jj][1] = __ldg(reinterpret_cast<const float4*>(ps + 8 * lane + 4));
            if (need_d) {
                q.w[jj][0] = __ldg(reinterpret_cast<const float4*>(pd + 8 * lane));
                q.w[jj][1] = __ldg(reinterpret_cast<const float4*>(pd + 8 * lane + 4));
            }
        }
        q.ut[jj] = lane < ntail ? __ldg(ps + 8 * nfull + lane) : 0.f;
        q.wt[jj] = (need_d && lane < ntail) ? __ldg(pd + 8 * nfull + lane) : 0.f;
    }
}
}  // namespace egws

// HS: the hidden width H = hidden_nf + 1 as a compile-time constant (257 for every shipped model: 32 full chunks, one
// leftover feature, one leftover output column), or 0 for run-time widths.  With run-time widths every row of build A
// carried the dot products, shuffle reductions and predicates of three possible leftover columns: 260 instructions per
// row and lane for 8 features.
template <int HS>
__global__ void __launch_bounds__(egws::NT, 1) egnn_edge_ws_kernel(const __grid_constant__ EgnnWsLaunch W) {
    using namespace egws;
    const EgnnEdgeLaunch& L = W.L;
    // each edge type gets exactly the tiles its capacity needs (a 2-D grid sized by the largest type would launch
    // ~2x as many CTAs that only find out they are empty)
    int et = 0;
    while (et < 3 && (int)blockIdx.x >= W.tile_off[et + 1]) ++et;
    const int bx = (int)blockIdx.x - W.tile_off[et];
    const EgnnEtypeArgs& a = L.e[et];
    const int tile_begin = bx * R;
    // this thread's edge, fetched together with the edge count (arrays are sized at capacity: the speculative read is
    // in bounds; rows past the end are re-read from the tile's last edge below)
    int my_s = 0, my_d = 0;
    if (threadIdx.x < R) {
        const int e = min(tile_begin + (int)threadIdx.x, max(a.cap, 1) - 1);
        my_s = __ldg(a.src + e); my_d = __ldg(a.dst + e);
    }
    // ... and the (source, destination) of the RPW rows every SIMT warp builds: warp-uniform addresses, one broadcast
    // transaction each, so that the gathers of the first-layer products can be issued before anything else
    constexpr int RPW = R / NW;
    int rs[RPW], rd[RPW];
    if ((int)(threadIdx.x >> 5) < NW) {
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const int e = min(tile_begin + (int)(threadIdx.x >> 5) * RPW + j, max(a.cap, 1) - 1);
            rs[j] = __ldg(a.src + e); rd[j] = __ldg(a.dst + e);
        }
    }
    const int E = a.rowptr[a.n_dst];
    if (tile_begin >= E) return;
    const int n = min(R, E - tile_begin);
    extern __shared__ __align__(128) unsigned char smem_eg[];
    TC_T(g0);
    const int kch = HS ? 2 * ((HS + 15) / 16) : W.kch;
    Sm m = carve(smem_eg, kch);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = HS ? HS : L.H, Hp = HS ? ((HS + 3) & ~3) : L.Hp, nmain = HS ? (HS < 256 ? HS : 256) & ~3 : L.nmain;
    const int nlo = HS ? HS - ((HS < 256 ? HS : 256) & ~3) : L.nlo;
    const int ksteps = kch >> 1;
    const int NB = (nmain + 15) & ~15;

    // asynchronous set-up: the control warps initialise the barriers and TMEM and start streaming weights at once; the
    // SIMT warps meet them at named barrier 3 only when they first need a barrier (after building the first A tile)
    uint32_t tmem = 0;
    if (warp >= NW) {
        if (warp == NW) {
            if (lane == 0) {
                for (int i = 0; i < C::STAGES; ++i) { tc::mbar_init(&m.full[i], 1); tc::mbar_init(&m.empty[i], 1); }
                for (int i = 0; i < 2; ++i) { tc::mbar_init(&m.a_ready[i], NW); tc::mbar_init(&m.acc_done[i], 1); }
                tc::fence_barrier_init();
            }
            __syncwarp();
            tc::tmem_alloc(m.tmem_slot, 512);
            tc::tmem_relinquish();
            tc::fence_before_sync();
        }
        asm volatile("bar.sync 2, 64;" ::: "memory");
        asm volatile("bar.arrive 3, %0;" ::"n"(egws::NT) : "memory");
        tc::fence_after_sync();
        tmem = *m.tmem_slot;
    }

    if (warp == NW + 1) {
        // ---- producer: W2 slabs of the edge branch, then of the coord branch
        if (lane == 0) {
            const uint32_t slab = C::NS * 2 * (NB / 8) * 128;
            uint32_t it = 0;
#pragma unroll
            for (int br = 0; br < 2; ++br)
                for (int j = 0; j < ksteps; ++j, ++it) {
                    const uint32_t st = it % C::STAGES;
                    if (it >= (uint32_t)C::STAGES) tc::mbar_wait(&m.empty[st], ((it / C::STAGES) - 1) & 1);
                    tc::mbar_arrive_expect_tx(&m.full[st], slab);
                    tc::bulk_g2s(m.ring + (size_t)st * C::SLAB, W.t[et].W2P[br] + (size_t)j * (slab / 16), slab, &m.full[st]);
                }
        }
    } else if (warp == NW) {
        // ---- MMA issuer: the whole (converged) warp runs the loop, one elected lane executes the tcgen05 instructions
        //      (tc::elect_one(): descriptors stay in uniform registers, no waterfall loop around every UTCHMMA)
        {
            const uint32_t idesc = tc::make_idesc_bf16(C::MMA_M, NB);
            const uint32_t b_k = (NB / 8) * 128, slab1 = 2 * b_k;
            uint32_t it = 0;
#pragma unroll
            for (int br = 0; br < 2; ++br) {
                tc::mbar_wait(&m.a_ready[br], 0);
                tc::fence_after_sync();
                for (int j = 0; j < ksteps; ++j, ++it) {
                    const uint32_t st = it % C::STAGES;
                    tc::mbar_wait(&m.full[st], (it / C::STAGES) & 1);
                    tc::fence_after_sync();
                    const uint32_t bs = tc::smem_u32(m.ring + (size_t)st * C::SLAB);
                    const uint64_t ad = tc::make_smem_desc(tc::smem_u32(m.A[br] + (size_t)2 * j * C::KCS), C::KCS, 128);
                    const uint64_t b0 = tc::make_smem_desc(bs, b_k, 128), b1 = tc::make_smem_desc(bs + slab1, b_k, 128);
                    if (tc::elect_one()) {
                        tc::mma_bf16_ss(tmem + 256 * br, ad, b0, idesc, j > 0 ? 1u : 0u);
                        tc::mma_bf16_ss(tmem + 256 * br, ad, b1, idesc, 1u);
                        tc::mma_commit(&m.empty[st]);
                    }
                }
                if (tc::elect_one()) tc::mma_commit(&m.acc_done[br]);
            }
        }
    } else {
        // ---- SIMT warps
        TC_T(g1);
        long long gt[8];
        // Row pairs of first-layer products in flight (registers): P_src chunks of both rows, P_dst chunks (the second row's
        // only when its destination differs: the tile is dst-sorted, 19 ll edges per ligand atom).  The kernel is bound by
        // the latency of these L2 gathers (225 KB of shared memory leave no L1), so they are issued as early as their
        // addresses exist -- branch 0 before the tile set-up below, branch 1 behind the rows of branch 0 as those are
        // consumed -- and added only where they are used.
        const int nfull = H >> 3, ntail = H & 7;       // full 8-feature chunks (one per lane: H <= 263) and leftover features
        if (n < R) {        // last tile of the edge type: rows >= n replicate the last valid edge
            const int sl = __ldg(a.src + tile_begin + n - 1), dl = __ldg(a.dst + tile_begin + n - 1);
#pragma unroll
            for (int j = 0; j < RPW; ++j)
                if (warp * RPW + j >= n) { rs[j] = sl; rd[j] = dl; }
            if (tid >= n && tid < R) { my_s = sl; my_d = dl; }
        }
        // Set-up loads first (they only need this thread's own edge, fetched at the top): the per-etype vectors and the
        // geometry (models/dynamics.py:160, :211, :169) -- then the gathers, whose row indices have arrived meanwhile -- and
        // only then the stores that wait for the set-up loads: one exposed L2 round trip instead of three.
        // (2 Hp <= 528 items over 512 threads: both rounds' loads are issued together.)
        static_assert(2 * VEC_LD <= 2 * NT_SIMT, "vector staging: two items per thread");
        float sv[2][6];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = tid + u * NT_SIMT;
            if (i < 2 * Hp) {
                const int br = i / Hp, k = i - br * Hp;
                sv[u][0] = a.w1c[br][k];
                sv[u][1] = a.b2[br][k];
                sv[u][2] = br == 0 ? a.watt[k] : a.w3c[k];
#pragma unroll
                for (int c = 0; c < 3; ++c) sv[u][3 + c] = c < nlo ? a.W2lo[br][c * Hp + k] : 0.f;
            }
        }
        int rp0 = 0, rp1 = 0;
        float gx[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (tid < R) {              // (rows >= n replicate the last valid edge, see above)
            rp0 = a.rowptr[my_d]; rp1 = a.rowptr[my_d + 1];
#pragma unroll
            for (int c = 0; c < 3; ++c) { gx[c] = a.xs[3 * my_s + c]; gx[3 + c] = a.xd[3 * my_d + c]; }
        }
        egws::RowPair q0, q1;
        egws::load_pair<0>(a, Hp, 0, rs, rd, lane, nfull, ntail, q0);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = tid + u * NT_SIMT;
            if (i < 2 * Hp) {
                const int br = i / Hp, k = i - br * Hp;
                // (carve(): w1c | b2 | wv of branch 0, then of branch 1, VEC_LD floats each; w2lo[br] = [3][VEC_LD].  No dynamic
                // index into the Sm arrays: that would put the whole struct into local memory, an L2 round trip per use)
                float* vb = m.w1c[0] + br * 3 * VEC_LD + k;
                vb[0] = sv[u][0];
                vb[VEC_LD] = sv[u][1];
                vb[2 * VEC_LD] = sv[u][2];
#pragma unroll
                for (int c = 0; c < 3; ++c) m.w2lo[0][(br * 3 + c) * VEC_LD + k] = sv[u][3 + c];
            }
        }
        if (tid < R) {
            m.src_s[tid] = my_s;
            m.dst_s[tid] = my_d;
            m.rp[2 * tid] = rp0;
            m.rp[2 * tid + 1] = rp1;
            const float dx = gx[0] - gx[3], dy = gx[1] - gx[4], dz = gx[2] - gx[5];
            const float dij = sqrtf(dx * dx + dy * dy + dz * dz);
            m.dij[tid] = dij;
            const float inv = 1.0f / (dij + 1.0f);
            m.xsc[3 * tid] = dx * inv; m.xsc[3 * tid + 1] = dy * inv; m.xsc[3 * tid + 2] = dz * inv;
        }
        egws::load_pair<1>(a, Hp, 0, rs, rd, lane, nfull, ntail, q1);       // (behind the stores: their registers are free now)
        simt_bar();
        // segment table of the dst-sorted tile (as ws::build_segments)
        {
            bool start = false;
            unsigned bal = 0;
            if (tid < R) {
                start = tid < n && (tid == 0 || m.dst_s[tid] != m.dst_s[tid - 1]);
                bal = __ballot_sync(0xffffffffu, start);
                if (lane == 0) m.warp_cnt[warp] = __popc(bal);
            }
            simt_bar();
            if (tid < R) {
                int base = 0;
                for (int w = 0; w < warp; ++w) base += m.warp_cnt[w];
                if (start) m.seg[base + __popc(bal & ((1u << lane) - 1u))] = tid;
                if (tid == 0) {
                    int tot = 0;
                    for (int w = 0; w < R / 32; ++w) tot += m.warp_cnt[w];
                    m.seg[tot] = n;
                    m.seg[R + 1] = tot;
                }
            }
        }
        // ---- 1. both branches: first Linear (factorised) + SiLU -> A[br]; leftover output columns as fp32 dots
        TC_T(g2);
        static_assert(RPW == 4, "build A: two row pairs per warp");
        // one row: v = P_src + P_dst chunk of this lane (8 features); dot = this lane's share of the leftover output columns.
        // The row's leftover FEATURE and the cross-lane sums of the leftover columns wait for finish_rows(): four rows at a
        // time, so that their serial chains (one lane's SiLU, five dependent shuffles) overlap instead of adding up
        auto build_row = [&](int br, int r, float (&v)[8], float (&dot)[3]) __attribute__((always_inline)) {
            const float* w1c = m.w1c[br];
            const float* w2lo = m.w2lo[br];
            unsigned char* Ab = m.A[br];
            const float d = m.dij[r];
            const uint32_t rof = ws::row_off<C>(r);
            dot[0] = dot[1] = dot[2] = 0.f;
            if (lane < nfull) {
                const float4 wa = *reinterpret_cast<const float4*>(w1c + 8 * lane), wb = *reinterpret_cast<const float4*>(w1c + 8 * lane + 4);
                float f[8];
                f[0] = ws::silu_acc(fmaf(wa.x, d, v[0])); f[1] = ws::silu_acc(fmaf(wa.y, d, v[1]));
                f[2] = ws::silu_acc(fmaf(wa.z, d, v[2])); f[3] = ws::silu_acc(fmaf(wa.w, d, v[3]));
                f[4] = ws::silu_acc(fmaf(wb.x, d, v[4])); f[5] = ws::silu_acc(fmaf(wb.y, d, v[5]));
                f[6] = ws::silu_acc(fmaf(wb.z, d, v[6])); f[7] = ws::silu_acc(fmaf(wb.w, d, v[7]));
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    if (cc < nlo) {
                        const float4 la = *reinterpret_cast<const float4*>(w2lo + cc * VEC_LD + 8 * lane);
                        const float4 lb = *reinterpret_cast<const float4*>(w2lo + cc * VEC_LD + 8 * lane + 4);
                        const float t = f[0] * la.x + f[1] * la.y + f[2] * la.z + f[3] * la.w + f[4] * lb.x + f[5] * lb.y + f[6] * lb.z + f[7] * lb.w;
                        dot[cc] = t;
                    }
                }
                uint4 hi, lo;
                ws::split8(f, hi, lo);
                const uint32_t off = (uint32_t)(lane * C::KCS) + rof;
                *reinterpret_cast<uint4*>(Ab + off) = hi;
                *reinterpret_cast<uint4*>(Ab + off + 256) = lo;
            }
            // leftover features (H = 257: feature 256) and the K padding: zero the chunks, then 2-byte stores
            if (lane < kch - nfull) {
                const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                const uint32_t off = (uint32_t)((nfull + lane) * C::KCS) + rof;
                *reinterpret_cast<uint4*>(Ab + off) = z;
                *reinterpret_cast<uint4*>(Ab + off + 256) = z;
            }
        };
        auto finish_rows = [&](int br, float (&dot)[RPW][3], const float (&vt)[RPW]) __attribute__((always_inline)) {
            const float* w1c = m.w1c[br];
            const float* w2lo = m.w2lo[br];
            unsigned char* Ab = m.A[br];
            __syncwarp();           // (the zero fill of the leftover chunk, by other lanes, precedes the 2-byte stores)
            if (lane < ntail) {
                const int k = 8 * nfull + lane;
#pragma unroll
                for (int j = 0; j < RPW; ++j) {
                    const int r = warp * RPW + j;
                    const float f = ws::silu_acc(fmaf(w1c[k], m.dij[r], vt[j]));
                    const uint32_t off = (uint32_t)((k >> 3) * C::KCS + (k & 7) * 2) + ws::row_off<C>(r);
                    const __nv_bfloat16 hi = __float2bfloat16(f);
                    *reinterpret_cast<__nv_bfloat16*>(Ab + off) = hi;
                    *reinterpret_cast<__nv_bfloat16*>(Ab + off + 256) = __float2bfloat16(f - __bfloat162float(hi));
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc)
                        if (cc < nlo) dot[j][cc] = fmaf(f, w2lo[cc * VEC_LD + k], dot[j][cc]);
                }
            }
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                if (cc < nlo) {
#pragma unroll
                    for (int o = 16; o; o >>= 1)
#pragma unroll
                        for (int j = 0; j < RPW; ++j) dot[j][cc] += __shfl_xor_sync(0xffffffffu, dot[j][cc], o);
                    // every lane holds the four sums: lane j finishes row j
                    float sj = dot[0][cc];
#pragma unroll
                    for (int j = 1; j < RPW; ++j) sj = lane == j ? dot[j][cc] : sj;
                    if (lane < RPW) m.lo[(br * R + warp * RPW + lane) * 4 + cc] = ws::silu_acc(sj + m.b2[br][nmain + cc]);
                }
            }
        };
        float dots[RPW][3], vts[RPW];
        auto build_pair = [&](int br, int j0, const egws::RowPair& q) __attribute__((always_inline)) {
            // (j0 = first row of the pair inside the warp's rows: a literal at every call)
            {
                float v[8];
                v[0] = q.u[0][0].x + q.w[0][0].x; v[1] = q.u[0][0].y + q.w[0][0].y; v[2] = q.u[0][0].z + q.w[0][0].z; v[3] = q.u[0][0].w + q.w[0][0].w;
                v[4] = q.u[0][1].x + q.w[0][1].x; v[5] = q.u[0][1].y + q.w[0][1].y; v[6] = q.u[0][1].z + q.w[0][1].z; v[7] = q.u[0][1].w + q.w[0][1].w;
                vts[j0] = q.ut[0] + q.wt[0];
                build_row(br, warp * RPW + j0, v, dots[j0]);
            }
            {
                const bool own_d = rd[j0 + 1] != rd[j0];
                const float4 w0 = own_d ? q.w[1][0] : q.w[0][0], w1 = own_d ? q.w[1][1] : q.w[0][1];
                const float wt = own_d ? q.wt[1] : q.wt[0];
                float v[8];
                v[0] = q.u[1][0].x + w0.x; v[1] = q.u[1][0].y + w0.y; v[2] = q.u[1][0].z + w0.z; v[3] = q.u[1][0].w + w0.w;
                v[4] = q.u[1][1].x + w1.x; v[5] = q.u[1][1].y + w1.y; v[6] = q.u[1][1].z + w1.z; v[7] = q.u[1][1].w + w1.w;
                vts[j0 + 1] = q.ut[1] + wt;
                build_row(br, warp * RPW + j0 + 1, v, dots[j0 + 1]);
            }
        };
        // branch 0, the rows of branch 1 fetched behind it
        build_pair(0, 0, q0);
        egws::load_pair<0>(a, Hp, 1, rs, rd, lane, nfull, ntail, q0);
        build_pair(0, 2, q1);
        egws::load_pair<1>(a, Hp, 1, rs, rd, lane, nfull, ntail, q1);
        finish_rows(0, dots, vts);
        // first use of an mbarrier / of TMEM: join the control warps' set-up
        asm volatile("bar.sync 3, %0;" ::"n"(egws::NT) : "memory");
        tc::fence_after_sync();
        tmem = *m.tmem_slot;
        publish(&m.a_ready[0]);
        gt[0] = clock64();
        build_pair(1, 0, q0);
        build_pair(1, 2, q1);
        finish_rows(1, dots, vts);
        publish(&m.a_ready[1]);
        gt[1] = clock64();
        // ---- 2. per branch: epilogue straight out of TMEM, then the deterministic segmented reduction
        const int q4 = warp & 3, cg = warp >> 2;
        const int ra = 16 * q4 + (lane >> 2), cp = 2 * (lane & 3);
        const uint32_t ro = ws::row_off<C>(ra) + cp * 2;
#pragma unroll
        for (int br = 0; br < 2; ++br) {
            tc::mbar_wait(&m.acc_done[br], 0);
            tc::fence_after_sync();
            const uint32_t taddr = tmem + ((uint32_t)(32 * q4) << 16) + 256 * br;
            float dota = 0.f, dotb = 0.f;
            const int cend = min(nmain, cg * CPW + CPW);
            for (int cb = cg * CPW; cb < cend; cb += 32) {
                uint32_t v0[16], v1[16];
                tc::tmem_ld_16x256b_x4(taddr + cb, v0);
                tc::tmem_ld_16x256b_x4(taddr + (16u << 16) + cb, v1);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int col = cb + 8 * i + cp;
                    if (cb + 8 * i < nmain) {
                        const float2 b = *reinterpret_cast<const float2*>(m.b2[br] + col);
                        const float2 wv = *reinterpret_cast<const float2*>(m.wv[br] + col);
                        const float fa0 = ws::silu_acc(__uint_as_float(v0[4 * i + 0]) + __uint_as_float(v1[4 * i + 0]) + b.x);
                        const float fa1 = ws::silu_acc(__uint_as_float(v0[4 * i + 1]) + __uint_as_float(v1[4 * i + 1]) + b.y);
                        const float fb0 = ws::silu_acc(__uint_as_float(v0[4 * i + 2]) + __uint_as_float(v1[4 * i + 2]) + b.x);
                        const float fb1 = ws::silu_acc(__uint_as_float(v0[4 * i + 3]) + __uint_as_float(v1[4 * i + 3]) + b.y);
                        dota = fmaf(fa0, wv.x, fmaf(fa1, wv.y, dota));
                        dotb = fmaf(fb0, wv.x, fmaf(fb1, wv.y, dotb));
                        if (br == 0) {          // m2 of the edge branch -> A[0] (its MMAs have completed), as (hi, lo) bf16
                            const uint32_t off = (uint32_t)((col >> 3) * C::KCS) + ro;
                            const uint32_t ha = tc::pack_bf16x2(fa0, fa1), hb = tc::pack_bf16x2(fb0, fb1);
                            *reinterpret_cast<uint32_t*>(m.A[0] + off) = ha;
                            *reinterpret_cast<uint32_t*>(m.A[0] + off + 128) = hb;
                            *reinterpret_cast<uint32_t*>(m.A[0] + off + 256) =
                                tc::pack_bf16x2(fa0 - __uint_as_float(ha << 16), fa1 - __uint_as_float(ha & 0xffff0000u));
                            *reinterpret_cast<uint32_t*>(m.A[0] + off + 384) =
                                tc::pack_bf16x2(fb0 - __uint_as_float(hb << 16), fb1 - __uint_as_float(hb & 0xffff0000u));
                        }
                    }
                }
            }
            dota += __shfl_xor_sync(0xffffffffu, dota, 1); dota += __shfl_xor_sync(0xffffffffu, dota, 2);
            dotb += __shfl_xor_sync(0xffffffffu, dotb, 1); dotb += __shfl_xor_sync(0xffffffffu, dotb, 2);
            if ((lane & 3) == 0) { m.dotp[cg * R + ra] = dota; m.dotp[cg * R + ra + 8] = dotb; }
            simt_bar();
            if (tid < R) {
                float dot = 0.f;
#pragma unroll
                for (int c2 = 0; c2 < NCG; ++c2) dot += m.dotp[c2 * R + tid];
                for (int cc = 0; cc < nlo; ++cc) dot = fmaf(m.lo[(br * R + tid) * 4 + cc], m.wv[br][nmain + cc], dot);
                if (br == 0) {
                    m.att[tid] = ws::sigmoid_acc(dot + a.batt[0]);          // msg_h = m2 * sigmoid(Linear(m2))  (:111-112)
                } else {
                    // msg_x = tanh(coord_mlp(f)) * x_diff * coords_range | coord_mlp(f) * x_diff  (:117-120)
                    const float cw = L.use_tanh ? tanhf(dot) * L.coords_range : dot;
                    m.xm[3 * tid] = cw * m.xsc[3 * tid]; m.xm[3 * tid + 1] = cw * m.xsc[3 * tid + 1]; m.xm[3 * tid + 2] = cw * m.xsc[3 * tid + 2];
                }
            }
            simt_bar();
            gt[2 + 2 * br] = clock64();
            // segmented reduction by destination (copy_e + sum, :177-185)
            const int nseg = m.seg[R + 1];
            float* part0 = a.part + ((size_t)bx * 2 + 0) * L.pw;
            float* part1 = a.part + ((size_t)bx * 2 + 1) * L.pw;
            if (br == 0) {
                // one warp per group of segments, one lane per 8-column chunk: 16-byte reads of the hi and lo planes
                // (bank-conflict free thanks to the +16 B chunk stride), 32-byte coalesced stores
                const int nchunk = nmain >> 3;
                if (lane < nchunk) {
                    const unsigned char* p0 = m.A[0] + (size_t)lane * C::KCS;
                    for (int sg = warp; sg < nseg; sg += NW) {
                        const int r0 = m.seg[sg], r1 = m.seg[sg + 1];
                        float acc[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) acc[q] = 0.f;
                        for (int j = r0; j < r1; ++j) {
                            const uint32_t o2 = ws::row_off<C>(j);
                            const uint4 wh = *reinterpret_cast<const uint4*>(p0 + o2), wl = *reinterpret_cast<const uint4*>(p0 + o2 + 256);
                            const float at = m.att[j];
                            acc[0] = fmaf(__uint_as_float(wh.x << 16) + __uint_as_float(wl.x << 16), at, acc[0]);
                            acc[1] = fmaf(__uint_as_float(wh.x & 0xffff0000u) + __uint_as_float(wl.x & 0xffff0000u), at, acc[1]);
                            acc[2] = fmaf(__uint_as_float(wh.y << 16) + __uint_as_float(wl.y << 16), at, acc[2]);
                            acc[3] = fmaf(__uint_as_float(wh.y & 0xffff0000u) + __uint_as_float(wl.y & 0xffff0000u), at, acc[3]);
                            acc[4] = fmaf(__uint_as_float(wh.z << 16) + __uint_as_float(wl.z << 16), at, acc[4]);
                            acc[5] = fmaf(__uint_as_float(wh.z & 0xffff0000u) + __uint_as_float(wl.z & 0xffff0000u), at, acc[5]);
                            acc[6] = fmaf(__uint_as_float(wh.w << 16) + __uint_as_float(wl.w << 16), at, acc[6]);
                            acc[7] = fmaf(__uint_as_float(wh.w & 0xffff0000u) + __uint_as_float(wl.w & 0xffff0000u), at, acc[7]);
                        }
                        const bool from_prev = (r0 == 0) && (m.rp[2 * r0] < tile_begin);
                        const bool into_next = (r1 == n) && (m.rp[2 * r0 + 1] > tile_begin + n);
                        float* t = (from_prev ? part0 : into_next ? part1 : a.hn + (size_t)m.dst_s[r0] * Hp) + 8 * lane;
                        *reinterpret_cast<float4*>(t) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                        *reinterpret_cast<float4*>(t + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
                    }
                }
                // the leftover feature columns: one thread per (segment, column)
                for (int i = tid; i < nseg * nlo; i += NT_SIMT) {
                    const int sg = i / nlo, cc = i - sg * nlo;
                    const int r0 = m.seg[sg], r1 = m.seg[sg + 1];
                    float s = 0.f;
                    for (int j = r0; j < r1; ++j) s = fmaf(m.lo[j * 4 + cc], m.att[j], s);
                    const bool from_prev = (r0 == 0) && (m.rp[2 * r0] < tile_begin);
                    const bool into_next = (r1 == n) && (m.rp[2 * r0 + 1] > tile_begin + n);
                    float* t = from_prev ? part0 : into_next ? part1 : a.hn + (size_t)m.dst_s[r0] * Hp;
                    t[nmain + cc] = s;
                }
            } else {
                // x messages: one thread per (segment, component); their partial slots follow the Hp feature columns
                for (int i = tid; i < nseg * 3; i += NT_SIMT) {
                    const int sg = i / 3, cc = i - sg * 3;
                    const int r0 = m.seg[sg], r1 = m.seg[sg + 1];
                    float s = 0.f;
                    for (int j = r0; j < r1; ++j) s += m.xm[3 * j + cc];
                    const bool from_prev = (r0 == 0) && (m.rp[2 * r0] < tile_begin);
                    const bool into_next = (r1 == n) && (m.rp[2 * r0 + 1] > tile_begin + n);
                    float* t = from_prev ? part0 + Hp + cc : into_next ? part1 + Hp + cc : a.xn + (size_t)m.dst_s[r0] * 4 + cc;
                    t[0] = s;
                }
            }
            // (the next branch's epilogue touches neither A[0] nor att/xm before its own barriers)
            gt[3 + 2 * br] = clock64();
        }
        EG_ACC(0, g0, g1); EG_ACC(1, g1, g2); EG_ACC(2, g2, gt[0]); EG_ACC(3, gt[0], gt[1]); EG_ACC(4, gt[1], gt[2]);
        EG_ACC(5, gt[2], gt[3]); EG_ACC(6, gt[3], gt[4]); EG_ACC(7, gt[4], gt[5]); EG_ACC(8, 0, 1);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == egws::NW) tc::tmem_dealloc(tmem, 512);
}
