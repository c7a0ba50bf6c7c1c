// tcgen05 GEMM for node-level rows, and the unit under test for the tensor-core primitives of tc.cuh:
//     Y[m, 0:N] = act( X[m, 0:K] @ W^T + bias ) (+R)
//   X fp32 row-major, staged into shared memory as the bf16 UMMA A operand; W pre-packed bf16 "k-step slabs"
//   (keypoint_diffusion_b200/pack.py: pack_tc_weight), streamed through a cp.async.bulk ring; fp32 accumulation in
//   TMEM, fp32 output.  NS = 1: plain bf16 operands, 128 rows per CTA.  NS = 2 ("bf16x3"): 64 rows per CTA, the
//   (hi, lo) bf16 rows stacked into one 128-row operand and two MMAs per k-step (W_hi, W_lo), which together give
//   all four hi/lo products -- fp32-grade results on the tensor cores (ws_common.cuh).
// blockIdx.y selects a 256-column block of the output (UMMA N <= 256).
#include "common.cuh"
#include "ws_common.cuh"

namespace kpd {

constexpr int TCG_STAGES = 4;

template <int NS>
__global__ void __launch_bounds__(128, 1)
tc_linear_kernel(const float* __restrict__ X, int ldx, const uint4* __restrict__ Wp, const float* __restrict__ bias,
                 const float* __restrict__ R, int ldr, float* __restrict__ Y, int ldy, int M, int K, int N, int NBmax,
                 int act) {
    using C = ws::Cfg<128 / NS, NS, 1>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int NB = N - 256 * (int)blockIdx.y;                // rows of this output block, padded to 16
    NB = NB > 256 ? 256 : (NB + 15) & ~15;
    const int ksteps = (K + 15) / 16;
    unsigned char* a_s = smem_raw;                     // [2*ksteps] k-chunks of C::KCS bytes
    unsigned char* b_s = a_s + (((size_t)2 * ksteps * C::KCS + 127) & ~(size_t)127);   // [STAGES][NS][2][NBmax/8][128 B]
    const int b_kstride = (NB / 8) * 128;
    const int slab1 = 2 * b_kstride;                   // one k-step of this block (hi or lo)
    const int slab_bytes = NS * slab1;
    const int slab_stride = NS * 2 * (NBmax / 8) * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + (size_t)TCG_STAGES * slab_stride);   // full[S], empty[S], done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TCG_STAGES + 1);
    uint64_t* full = bars;
    uint64_t* empty = bars + TCG_STAGES;
    uint64_t* done = bars + 2 * TCG_STAGES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * C::R;
    const int nblk = blockIdx.y;                       // 256-column block of the output
    // all blocks before this one are full (256 rows): their slabs are NS * 2 * 32 * 128 B per k-step
    const uint4* Wblk = Wp + (size_t)nblk * ksteps * (NS * 2 * 32 * 128 / 16);
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < NB) tmem_cols <<= 1;

    if (tid == 0) {
        for (int i = 0; i < TCG_STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) { tc::tmem_alloc(tmem_slot, tmem_cols); tc::tmem_relinquish(); }
    __syncthreads();
    // ---- the producer starts streaming weights while everybody stages A
    if (tid == 32) {
        for (int j = 0; j < ksteps && j < TCG_STAGES; ++j) {
            tc::mbar_arrive_expect_tx(&full[j], slab_bytes);
            tc::bulk_g2s(b_s + (size_t)j * slab_stride, Wblk + (size_t)j * (slab_bytes / 16), slab_bytes, &full[j]);
        }
    }
    // ---- stage the A tile: one warp per row, a lane per 8-element k-chunk (coalesced 32-byte reads)
    {
        const int nch = 2 * ksteps;
        for (int r = warp; r < C::R; r += 4) {
            const int gm = m0 + r;
            const float* xr = X + (size_t)min(gm, M - 1) * ldx;
            for (int c = lane; c < nch; c += 32) {
                float v[8];
                if (gm < M && 8 * c + 8 <= K) {
                    const float4 u0 = *reinterpret_cast<const float4*>(xr + 8 * c), u1 = *reinterpret_cast<const float4*>(xr + 8 * c + 4);
                    v[0] = u0.x; v[1] = u0.y; v[2] = u0.z; v[3] = u0.w; v[4] = u1.x; v[5] = u1.y; v[6] = u1.z; v[7] = u1.w;
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] = (gm < M && 8 * c + q < K) ? xr[8 * c + q] : 0.0f;
                }
                uint4 hi, lo;
                ws::split8(v, hi, lo);
                const uint32_t off = (uint32_t)(c * C::KCS) + ws::row_off<C>(r);
                *reinterpret_cast<uint4*>(a_s + off) = hi;
                if (NS == 2) *reinterpret_cast<uint4*>(a_s + off + 256) = lo;
            }
        }
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // ---- warp 1 lane 0 keeps feeding the weight ring (cp.async.bulk), thread 0 issues the MMAs into TMEM
    if (tid == 32) {
        for (int j = TCG_STAGES; j < ksteps; ++j) {
            const int st = j % TCG_STAGES;
            tc::mbar_wait(&empty[st], ((j / TCG_STAGES) - 1) & 1);
            tc::mbar_arrive_expect_tx(&full[st], slab_bytes);
            tc::bulk_g2s(b_s + (size_t)st * slab_stride, Wblk + (size_t)j * (slab_bytes / 16), slab_bytes, &full[st]);
        }
    }
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(128, NB);
        for (int j = 0; j < ksteps; ++j) {
            const int st = j % TCG_STAGES;
            tc::mbar_wait(&full[st], (j / TCG_STAGES) & 1);
            tc::fence_after_sync();
            const uint64_t adesc = tc::make_smem_desc(tc::smem_u32(a_s + (size_t)2 * j * C::KCS), C::KCS, 128);
            const uint32_t bs = tc::smem_u32(b_s + (size_t)st * slab_stride);
            tc::mma_bf16_ss(tmem_base, adesc, tc::make_smem_desc(bs, b_kstride, 128), idesc, j > 0 ? 1u : 0u);
            if (NS == 2) tc::mma_bf16_ss(tmem_base, adesc, tc::make_smem_desc(bs + slab1, b_kstride, 128), idesc, 1u);
            tc::mma_commit(&empty[st]);
        }
        tc::mma_commit(done);
    }
    __syncwarp();
    tc::mbar_wait(done, 0);
    tc::fence_after_sync();

    // ---- epilogue
    if (NS == 1) {          // thread r reads its accumulator row from TMEM
        const int r = tid, gm = m0 + r;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < NB; c0 += 32) {
            uint32_t v[32];
            tc::tmem_ld_x32(lane_addr + c0, v);
            tc::tmem_ld_wait();
            if (gm < M) {
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const int gn = nblk * 256 + c0 + q;
                    if (c0 + q < NB && gn < N) {
                        float f = __uint_as_float(v[q]) + (bias ? bias[gn] : 0.0f);
                        if (act == 1) f = silu_f(f);
                        if (R) f += R[(size_t)gm * ldr + gn];
                        Y[(size_t)gm * ldy + gn] = f;
                    }
                }
            }
        }
    } else {                // lanes [32w, 32w+16) = hi rows, [32w+16, 32w+32) = lo rows of tile rows [16w, 16w+16)
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int ra = 16 * warp + (lane >> 2), cp = 2 * (lane & 3);
        for (int c0 = 0; c0 < NB; c0 += 64) {
            uint32_t v0[32], v1[32];
            tc::tmem_ld_16x256b_x8(lane_addr + c0, v0);
            tc::tmem_ld_16x256b_x8(lane_addr + (16u << 16) + c0, v1);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int gn = nblk * 256 + c0 + 8 * i + cp;
                if (c0 + 8 * i < NB) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int gm = m0 + ra + 8 * h;
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            if (gm < M && gn + e < N) {
                                float f = __uint_as_float(v0[4 * i + 2 * h + e]) + __uint_as_float(v1[4 * i + 2 * h + e]) +
                                          (bias ? bias[gn + e] : 0.0f);
                                if (act == 1) f = silu_f(f);
                                if (R) f += R[(size_t)gm * ldr + gn + e];
                                Y[(size_t)gm * ldy + gn + e] = f;
                            }
                        }
                    }
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, tmem_cols);
}

template <int NS>
static size_t tc_linear_smem(int K, int NB) {
    using C = ws::Cfg<128 / NS, NS, 1>;
    const int ksteps = (K + 15) / 16;
    return (((size_t)2 * ksteps * C::KCS + 127) & ~(size_t)127) + (size_t)TCG_STAGES * NS * 2 * (NB / 8) * 128 + (2 * TCG_STAGES + 1) * 8 + 16;
}

// Wp: packed by pack_tc_weight(W[N,K], split = (nsplit == 2)) -> for each 256-row block nb: [ksteps][hi(, lo)][2][NB/8][8][8]
// bf16, NB = rows of the block rounded up to 16 (all blocks but the last have NB = 256).
// X rows must be 16-byte aligned (ldx % 4 == 0).
template <int NS>
static int launch_tc_linear_ns(const float* X, int ldx, const void* Wp, const float* bias, const float* R, int ldr, float* Y,
                               int ldy, int M, int K, int N, int act, cudaStream_t st) {
    const int NB = N >= 256 ? 256 : (N + 15) / 16 * 16;
    const int nblocks = cdiv(N, 256);
    const size_t smem = tc_linear_smem<NS>(K, NB);
    KPD_REQUIRE(smem <= 227 * 1024, "tc_linear: K=%d needs %zu B of shared memory", K, smem);
    KPD_REQUIRE(ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0, "tc_linear: X rows must be 16-byte aligned (ldx=%d)", ldx);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(tc_linear_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        KPD_REQUIRE(e == cudaSuccess, "tc_linear: cannot set %zu B shared memory: %s", smem, cudaGetErrorString(e));
        configured = smem;
    }
    tc_linear_kernel<NS><<<dim3(cdiv(M, 128 / NS), nblocks), 128, smem, st>>>(X, ldx, static_cast<const uint4*>(Wp), bias, R, ldr,
                                                                             Y, ldy, M, K, N, NB, act);
    return check_launch("tc_linear_kernel");
}

int launch_tc_linear(const float* X, int ldx, const void* Wp, const float* bias, const float* R, int ldr, float* Y,
                     int ldy, int M, int K, int N, int act, int nsplit, cudaStream_t st) {
    if (M <= 0 || N <= 0) return 0;
    KPD_REQUIRE(nsplit == 1 || nsplit == 2, "tc_linear: nsplit must be 1 (bf16) or 2 (bf16x3)");
    return nsplit == 1 ? launch_tc_linear_ns<1>(X, ldx, Wp, bias, R, ldr, Y, ldy, M, K, N, act, st)
                       : launch_tc_linear_ns<2>(X, ldx, Wp, bias, R, ldr, Y, ldy, M, K, N, act, st);
}

}  // namespace kpd

extern "C" int kpd_tc_linear(const float* X, int32_t ldx, const void* Wp, const float* bias, const float* R, int32_t ldr,
                             float* Y, int32_t ldy, int32_t M, int32_t K, int32_t N, int32_t act, int32_t nsplit, void* stream) {
    return kpd::launch_tc_linear(X, ldx, Wp, bias, R, ldr, Y, ldy, M, K, N, act, nsplit, static_cast<cudaStream_t>(stream));
}
