// tcgen05 GEMM for node-level rows in bf16 mode, and the unit under test for the tensor-core
// primitives of tc.cuh:   Y[m, 0:N] = act( X[m, 0:K] @ W^T + bias ) (+R)
//   X fp32 row-major (converted to bf16 while staging), W pre-packed bf16 "k-step slabs"
//   (keypoint_diffusion_b200/pack.py: pack_tc_weight), fp32 accumulation in TMEM, fp32 output.
// One CTA per 128 rows; N <= 256 per CTA column block (blockIdx.y selects a 256-column block).
#include "common.cuh"
#include "tc.cuh"

namespace kpd {

constexpr int TCG_STAGES = 4;

// W is packed per 256-row output block (pack.pack_tc_weight); slab j of a block with NB rows (multiple of 16,
// <= 256): 2 k-chunks x (NB/8) groups x 128 B
__global__ void __launch_bounds__(128, 1)
tc_linear_kernel(const float* __restrict__ X, int ldx, const uint4* __restrict__ Wp, const float* __restrict__ bias,
                 const float* __restrict__ R, int ldr, float* __restrict__ Y, int ldy, int M, int K, int N, int NBmax,
                 int act) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int NB = N - 256 * (int)blockIdx.y;                // rows of this output block, padded to 16
    NB = NB > 256 ? 256 : (NB + 15) & ~15;
    const int ksteps = (K + 15) / 16;
    const int a_kstride = 16 * 128;                    // 128 rows -> 16 groups x 128 B per k-chunk
    unsigned char* a_s = smem_raw;                     // [2*ksteps][16][128 B]
    unsigned char* b_s = a_s + (size_t)2 * ksteps * a_kstride;   // [STAGES][2][NBmax/8][128 B]
    const int b_kstride = (NB / 8) * 128;
    const int slab_bytes = 2 * b_kstride;
    const int slab_stride = 2 * (NBmax / 8) * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + (size_t)TCG_STAGES * slab_stride);   // full[S], empty[S], done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TCG_STAGES + 1);
    uint64_t* full = bars;
    uint64_t* empty = bars + TCG_STAGES;
    uint64_t* done = bars + 2 * TCG_STAGES;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.x * 128;
    const int nblk = blockIdx.y;                       // 256-column block of the output
    // all blocks before this one are full (256 rows): their slabs are 2 * 32 * 128 B per k-step
    const uint4* Wblk = Wp + (size_t)nblk * ksteps * (2 * 32 * 128 / 16);
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < NB) tmem_cols <<= 1;

    if (tid == 0) {
        for (int i = 0; i < TCG_STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) { tc::tmem_alloc(tmem_slot, tmem_cols); tc::tmem_relinquish(); }

    // ---- stage the A tile: thread r owns row r; converts fp32 -> bf16, 8 elements (16 B) per k-chunk
    {
        const int r = tid, gm = m0 + r;
        for (int c = 0; c < 2 * ksteps; ++c) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int k = 8 * c + q;
                v[q] = (gm < M && k < K) ? X[(size_t)gm * ldx + k] : 0.0f;
            }
            uint4 pk;
            pk.x = tc::pack_bf16x2(v[0], v[1]); pk.y = tc::pack_bf16x2(v[2], v[3]);
            pk.z = tc::pack_bf16x2(v[4], v[5]); pk.w = tc::pack_bf16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(a_s + (size_t)c * a_kstride + (r >> 3) * 128 + (r & 7) * 16) = pk;
        }
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // ---- warp 1 lane 0 feeds the weight ring (cp.async.bulk), thread 0 issues the MMAs into TMEM
    if (tid == 32) {
        for (int j = 0; j < ksteps; ++j) {
            const int st = j % TCG_STAGES;
            if (j >= TCG_STAGES) tc::mbar_wait(&empty[st], ((j / TCG_STAGES) - 1) & 1);
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
            const uint64_t adesc = tc::make_smem_desc(tc::smem_u32(a_s + (size_t)2 * j * a_kstride), a_kstride, 128);
            const uint64_t bdesc = tc::make_smem_desc(tc::smem_u32(b_s + (size_t)st * slab_stride), b_kstride, 128);
            tc::mma_bf16_ss(tmem_base, adesc, bdesc, idesc, j > 0 ? 1u : 0u);
            tc::mma_commit(&empty[st]);
        }
        tc::mma_commit(done);
    }
    __syncwarp();
    tc::mbar_wait(done, 0);
    tc::fence_after_sync();

    // ---- epilogue: thread r reads its accumulator row from TMEM
    {
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
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, tmem_cols);
}

static size_t tc_linear_smem(int K, int NB) {
    const int ksteps = (K + 15) / 16;
    return (size_t)2 * ksteps * 16 * 128 + (size_t)TCG_STAGES * 2 * (NB / 8) * 128 + (2 * TCG_STAGES + 1) * 8 + 16;
}

// Wp: packed by pack_tc_weight(W[N,K]) -> for each 256-row block nb: [ksteps][2][NB/8][8][8] bf16, NB = rows of the block
// rounded up to 16 (all blocks but the last have NB = 256).  Only N <= 256 or N % 256 == 0 plus a tail is supported.
int launch_tc_linear(const float* X, int ldx, const void* Wp, const float* bias, const float* R, int ldr, float* Y,
                     int ldy, int M, int K, int N, int act, cudaStream_t st) {
    if (M <= 0 || N <= 0) return 0;
    const int NB = N >= 256 ? 256 : (N + 15) / 16 * 16;
    const int nblocks = cdiv(N, 256);
    const size_t smem = tc_linear_smem(K, NB);
    KPD_REQUIRE(smem <= 227 * 1024, "tc_linear: K=%d needs %zu B of shared memory", K, smem);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(tc_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        KPD_REQUIRE(e == cudaSuccess, "tc_linear: cannot set %zu B shared memory: %s", smem, cudaGetErrorString(e));
        configured = smem;
    }
    tc_linear_kernel<<<dim3(cdiv(M, 128), nblocks), 128, smem, st>>>(X, ldx, static_cast<const uint4*>(Wp), bias, R, ldr, Y, ldy,
                                                                M, K, N, NB, act);
    return check_launch("tc_linear_kernel");
}

}  // namespace kpd

extern "C" int kpd_tc_linear(const float* X, int32_t ldx, const void* Wp, const float* bias, const float* R, int32_t ldr,
                             float* Y, int32_t ldy, int32_t M, int32_t K, int32_t N, int32_t act, void* stream) {
    return kpd::launch_tc_linear(X, ldx, Wp, bias, R, ldr, Y, ldy, M, K, N, act, static_cast<cudaStream_t>(stream));
}
