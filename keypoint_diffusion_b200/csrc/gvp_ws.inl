// Warp-specialised tensor-core GVP kernels (included inside namespace kpd by gvp.cu).
//
// A tile is R rows (edges or nodes).  Its scalar features live in shared memory ONLY as bf16 in the UMMA K-major
// canonical layout (tc.cuh), i.e. directly as the A operand of tcgen05.mma.  NS = 1: plain bf16, R = 128 (M = 128)
// or R = 64 (M = 64).  NS = 2 ("bf16x3"): every row is kept as hi = bf16(x) and lo = bf16(x - hi) and the hi / lo
// rows of 64 tile rows are STACKED into one 128-row operand (ws_common.cuh: row_off); two MMAs per k-step (W_hi,
// W_lo) then give all four hi/lo products, the epilogue adds the hi-row and lo-row accumulators: ~16 mantissa bits
// per operand, inside the 1e-4 fp32 parity bar (tools/split_precision_study.py).
// Vector channels never touch shared memory during the chain: they live in registers as mma.sync fragments (VF), so
// Vh = V^T Wh and Vu = Vh^T Wu are warp-level tensor-core GEMMs (TF32, 3xTF32 for NS = 2) chained without shuffles.
//
// Warp roles (one CTA): R/8 SIMT warps | 1 MMA-issuing warp (one lane) | 1 weight-producer warp (one lane).
//   producer: bulk-copies the shared-memory images of the chain's small fp32 weights (pack.pack_gvp_small), streams
//             every GVP's packed to_feats_out weight as k-step slabs through a cp.async.bulk ring (full/empty
//             mbarriers) and the small gates weight behind them;
//   MMA warp: issues the k-steps over the feats columns as soon as those are complete (feats_ready) -- i.e.
//             while the SIMT warps still compute Vh -- then the k-steps over the |Vh| columns (tail_ready),
//             commits acc_done; the gates GEMM follows progressively behind the two halves of epilogue 1
//             (half_ready, feats_ready) and, right behind it, the NEXT GVP's main k-steps;
//   SIMT:     Vh, |Vh| -> A, Vu, epilogue 1 (TMEM -> bias + SiLU -> bf16 planes of A), epilogue 2 (gates ->
//             sigmoid -> V), gathers (16-byte cp.async straight into A), LayerNorms and the deterministic
//             segmented reduction.
// SIMT-only synchronisation uses named barrier 1; cross-role synchronisation uses mbarriers only.  Optional
// thread-block clusters (Cfg::CL > 1) let neighbouring tiles share ONE multicast weight stream.
// phase timers of the warp-specialised kernels (cycles, SIMT thread 0, summed over CTAs):
// [0..6] gvp_simt: Vh+|Vh|, Vu, wait acc, epilogue 1, wait gates, epilogue 2, calls
__device__ unsigned long long g_ws_times[64];   // edge kernel: base 0, node kernel: base 16, head kernel: base 32 (+8..13: kernel phases)
#define WS_ACC(slot, a, b) do { if (threadIdx.x == 0) atomicAdd(&g_ws_times[slot], (unsigned long long)((b) - (a))); } while (0)

// event trace of ONE CTA (block (1, 1) of the edge kernel) for timeline analysis: (tag, warp, clock) triples
#ifdef KPD_WS_TRACE
#ifndef KPD_TRACE_BX
#define KPD_TRACE_BX 2          // (even: the leader of a CTA pair)
#endif
__device__ unsigned long long g_ws_trace[3 * 2048];
__device__ int g_ws_trace_n;
#define WS_TRACE(tag) do { if (blockIdx.x == KPD_TRACE_BX && blockIdx.y == 1 && (threadIdx.x & 31) == 0 && ((threadIdx.x >> 5) == 0 || (threadIdx.x >> 5) >= 16)) { \
        const int _i = atomicAdd(&g_ws_trace_n, 1); \
        if (_i < 2048) { g_ws_trace[3 * _i] = (tag); g_ws_trace[3 * _i + 1] = threadIdx.x >> 5; g_ws_trace[3 * _i + 2] = clock64(); } } } while (0)
#else
#define WS_TRACE(tag) do { } while (0)
#endif

namespace ws {

constexpr int WH_LD_KS = 28, WU_LD_KS = 20;          // fp32 [24][LD] images of Wh / Wu for vec_fma
constexpr int WSM_KS_W = 24 * WH_LD_KS + 24 * WU_LD_KS;     // (== Cfg::WSM_W of the KS configuration)
constexpr int VS_LD = 52;                                   // fp32 vector staging row (48 used)
constexpr int GATE_LD = 20;                                 // gates staged as [R][20]
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t GATE_COL = 256;
// KS edge kernel: accumulator [0, 256) | gates accumulator [256, 272) | hi plane of feats_out as the next GVP's A operand
// (bf16 pairs: column 320 + k / 2) [320, 448)
constexpr uint32_t KS_A_COL = 320;


struct Sm {
    unsigned char* A[2];
    unsigned char* ring;
    unsigned char* Wg[2];
    float* wsm;       // Wh | Wu of the current GVP (one buffer, refilled after every Vu GEMM)
    float* bias;      // bf | bg, two buffers (GVP parity)
    float* gate;      // NS = 1: gates [R][GATE_LD]; NS = 2: per-column-group partial gates [NCG][R][16]
    int *src_s, *dst_s, *seg, *rp;
    uint64_t *full, *empty, *wg_full, *wg_empty, *feats_ready, *tail_ready, *acc_done, *gates_done, *wsm_full, *half_ready, *wsm_empty;
    uint32_t* tmem_slot;
    int* warp_cnt;
    // CTA pairs (Cfg::CL == 2): rank in the pair; solo = the peer's tile is empty (it only lends its half of B)
    uint32_t rank;
    bool solo;
};

// bytes of one A operand (all MMA_M rows); KS keeps two of them (hi plane, lo plane)
template <class C>
__host__ __device__ inline size_t plane_bytes(int kch) { return ((size_t)kch * C::KCS + 127) & ~(size_t)127; }
template <class C>
__host__ __device__ inline size_t planes_bytes(int kch) { return (C::KS ? 2 : 1) * plane_bytes<C>(kch); }
// per-row staging of the gates between epilogue 1 / the gates GEMM and epilogue 2 (floats per tile row): STACK keeps
// the per-column-group partial sums of the mma.sync gates, the bf16 mode the finished gates; KS reads its gates from
// TMEM directly in the vector-fragment layout (no staging)
template <class C>
constexpr int GATE_FLOATS = C::KS ? 0 : (C::STACK ? C::NCG * 16 : GATE_LD);


// Stages of the weight ring.  The ring is latency-bound (a slab is re-requested when its MMA has completed and lands
// ~1350 cycles later), so the k-step rate is (MMA completion + copy latency) / stages: as deep as shared memory allows.
template <class C>
constexpr int GST = C::NS == 1 ? 10 : C::KS ? 4 : 6;
// bytes of a ring slot: one k-step of a 256-row weight (hi [, lo]).  (KS with 8 KB slots -- the hi and the lo slab of a
// k-step as separate copies, 6-7 slots -- was measured SLOWER, ~1000 instead of ~600 cycles per k-step: the cost of a
// bulk copy is dominated by a per-copy term, not by its bytes.)
template <class C>
constexpr int SLOT = C::SLAB;

template <class C>
static size_t smem_bytes(int kch) {
    return planes_bytes<C>(kch) + (size_t)GST<C> * SLOT<C> + C::WGB * C::WG_BYTES + sizeof(float) * (C::WSM_W + 2 * C::WSM_B) +
           sizeof(float) * GATE_FLOATS<C> * C::R + sizeof(int) * (5 * C::R + 8 + 8) + sizeof(uint64_t) * (2 * GST<C> + 12) + 16 + 128;
}

template <class C>
__device__ __forceinline__ Sm carve(unsigned char* smem, int kch) {
    Sm m;
    m.A[0] = smem;
    // lo rows: STACK two row groups after their hi rows (same operand); KS a plane of their own
    m.A[1] = C::KS ? smem + plane_bytes<C>(kch) : smem + (C::NS - 1) * 256;
    m.ring = smem + planes_bytes<C>(kch);
    m.Wg[0] = m.ring + GST<C> * SLOT<C>;
    m.Wg[1] = m.Wg[0] + C::WG_BYTES;
    m.wsm = reinterpret_cast<float*>(m.Wg[0] + C::WGB * C::WG_BYTES);
    m.bias = m.wsm + C::WSM_W;
    m.gate = m.bias + 2 * C::WSM_B;
    m.src_s = reinterpret_cast<int*>(m.gate + GATE_FLOATS<C> * C::R);
    m.dst_s = m.src_s + C::R;
    m.seg = m.dst_s + C::R;            // [R + 8]
    m.rp = m.seg + C::R + 8;           // [2R]
    m.warp_cnt = m.rp + 2 * C::R;      // [8]
    m.full = reinterpret_cast<uint64_t*>(m.warp_cnt + 8);
    m.empty = m.full + GST<C>;
    m.wg_full = m.empty + GST<C>;      // [2]
    m.wg_empty = m.wg_full + 2;        // [2]
    m.feats_ready = m.wg_empty + 2;
    m.tail_ready = m.feats_ready + 1;
    m.acc_done = m.tail_ready + 1;
    m.gates_done = m.acc_done + 1;
    m.wsm_full = m.gates_done + 1;
    m.half_ready = m.wsm_full + 1;
    m.wsm_empty = m.half_ready + 1;
    m.tmem_slot = reinterpret_cast<uint32_t*>(m.wsm_empty + 1);
    m.rank = 0;
    m.solo = false;
    return m;
}

template <class C>
__device__ __forceinline__ void simt_bar() { asm volatile("bar.sync 1, %0;" ::"n"(C::NT_SIMT) : "memory"); }

template <class C>
__device__ __forceinline__ void init_barriers(Sm& m) {      // one thread
    // CTA pairs: the leader's MMA thread waits for BOTH CTAs' operands, so the leader's full / *_ready barriers count
    // one more arrival: the peer's forwarding threads (relay(), forward_ready()); everything else stays per CTA
    const uint32_t extra = (C::CL == 2 && m.rank == 0 && !m.solo) ? 1u : 0u;
    for (int i = 0; i < GST<C>; ++i) { tc::mbar_init(&m.full[i], (C::CL == 2 && m.rank == 0) ? 2 : 1); tc::mbar_init(&m.empty[i], 1); }
    // wg_empty: NS = 1 the tcgen05 gates GEMM commits it; STACK every SIMT warp arrives after its mma.sync gates
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&m.wg_full[i], 1); tc::mbar_init(&m.wg_empty[i], C::STACK ? C::NW : 1); }
    tc::mbar_init(m.feats_ready, C::NW + extra);
    tc::mbar_init(m.tail_ready, C::NWV + extra);
    tc::mbar_init(m.acc_done, 1);
    tc::mbar_init(m.gates_done, 1);
    tc::mbar_init(m.wsm_full, 1);
    tc::mbar_init(m.half_ready, C::NW + extra);
    tc::mbar_init(m.wsm_empty, C::NWV);
    tc::fence_barrier_init();
}

// barriers + TMEM + zeroed A planes; ends with a __syncthreads() (the small fp32 weights arrive as one bulk copy
// per GVP issued by the producer, see produce())
template <class C>
__device__ __forceinline__ uint32_t setup(Sm& m, int kch) {
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) init_barriers<C>(m);
    if (warp == C::NW) {
        if (C::CL == 2) { tc::tmem_alloc_pair(m.tmem_slot, TMEM_COLS); tc::tmem_relinquish_pair(); }
        else { tc::tmem_alloc(m.tmem_slot, TMEM_COLS); tc::tmem_relinquish(); }
    }
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const int nz = (int)(planes_bytes<C>(kch) / 16);
    for (int i = tid; i < nz; i += C::NT) reinterpret_cast<uint4*>(m.A[0])[i] = z;
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    if (C::CL > 1) tc::cluster_sync();      // both CTAs' barriers and TMEM exist before the pair's first MMA / remote arrive
    tc::fence_after_sync();
    return *m.tmem_slot;
}

// Asynchronous set-up (edge kernel, no clusters): the control warps initialise the barriers and TMEM and start
// streaming weights at once; the SIMT warps meet them at named barrier 3 only after their gathers are in flight.
template <class C>
__device__ __forceinline__ void control_setup(Sm& m) {      // the two control warps
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == C::NW) {
        if (lane == 0) init_barriers<C>(m);
        __syncwarp();
        tc::tmem_alloc(m.tmem_slot, TMEM_COLS);
        tc::tmem_relinquish();
        tc::fence_before_sync();
    }
    asm volatile("bar.sync 2, 64;" ::: "memory");
    asm volatile("bar.arrive 3, %0;" ::"n"(C::NT) : "memory");
    tc::fence_after_sync();
}
template <class C>
__device__ __forceinline__ uint32_t simt_join(Sm& m) {       // the SIMT warps, before their first mbarrier / TMEM use
    asm volatile("bar.sync 3, %0;" ::"n"(C::NT) : "memory");
    tc::fence_after_sync();
    return *m.tmem_slot;
}

template <class C>
__device__ __forceinline__ void teardown(uint32_t tmem) {
    tc::fence_before_sync();
    __syncthreads();
    if (C::CL > 1) tc::cluster_sync();      // no CTA frees TMEM or leaves while the pair's MMAs / arrives may still touch it
    if ((threadIdx.x >> 5) == C::NW) {
        if (C::CL == 2) tc::tmem_dealloc_pair(tmem, TMEM_COLS);
        else tc::tmem_dealloc(tmem, TMEM_COLS);
    }
}

// Order of the feats GEMM's k-steps.  bf16x3 (NS = 2), GVPs after the first of a chain: epilogue 1 of the previous GVP
// writes the A columns in two halves per column group, and the k-steps over the first halves are issued (into the
// other TMEM accumulator) while the second halves are still being produced -- so sequence position i maps to k-step
// j = first halves of all column groups, then second halves, then the |Vh| tail.  Otherwise j = i.
template <class C>
__device__ __forceinline__ int kstep_at(int i, int ksm, bool chained) {
    if (!C::STACK || !chained || i >= ksm) return i;
    constexpr int kpg = (256 / C::NCG) / 16, kph = kpg / 2;      // k-steps per column group / per half
    int n0 = (ksm / kpg) * kph + min(ksm % kpg, kph);            // k-steps that lie in first halves
    const bool second = i >= n0;
    const int r = second ? i - n0 : i;
    // r-th k-step of its half-set: full groups contribute kph each (the last group may be partial)
    if (!second) return (r / kph) * kpg + (r % kph);
    return (r / kph) * kpg + kph + (r % kph);
}
template <class C>
__device__ __forceinline__ int first_half_ksteps(int ksm) {
    constexpr int kpg = (256 / C::NCG) / 16, kph = kpg / 2;
    return (ksm / kpg) * kph + min(ksm % kpg, kph);
}

// ------------------------------------------------------------------ producer (one thread)
template <class C>
__device__ __forceinline__ void produce(const GvpW* gv, int n_gvps, Sm& m, bool dead = false) {
    uint32_t it = 0;
#ifdef KPD_WS_TRACE
    unsigned long long tp[64];
#endif
    for (int g = 0; g < n_gvps; ++g) {
        const GvpW& w = gv[g];
        const int NBf = (w.fout + 15) & ~15, ksf = (w.fin + w.hd + 15) >> 4, ksg = NBf >> 4;
        const uint32_t slab = C::NS * 2 * (NBf / 8) * 128;
        [[maybe_unused]] const int b = g % (C::WGB > 0 ? C::WGB : 1);
        const uint4* WfP = C::NS == 2 ? w.WfP2 : w.WfP;
        if (!dead) {
            // shared-memory image of this GVP's small fp32 weights: Wh | Wu into the single buffer once the vector warps
            // are through the previous GVP's Vu GEMM, bf | bg into the buffer of this GVP's parity
            if (g > 0) tc::mbar_wait(m.wsm_empty, (g - 1) & 1);
            if constexpr (C::NS == 2) {
                // bf16x3: the plain fp32 image for the FP32-pipe vector GEMMs (vec_fma); it follows the fragment image
                const float* img = w.wsmP2 + C::WSM;
                tc::mbar_arrive_expect_tx(m.wsm_full, (uint32_t)((WSM_KS_W + C::WSM_B) * sizeof(float)));
                tc::bulk_g2s(m.wsm, img, WSM_KS_W * sizeof(float), m.wsm_full);
                tc::bulk_g2s(m.bias + (g & 1) * C::WSM_B, img + WSM_KS_W, C::WSM_B * sizeof(float), m.wsm_full);
            } else {
                const float* img = w.wsmP;
                tc::mbar_arrive_expect_tx(m.wsm_full, (uint32_t)(C::WSM * sizeof(float)));
                tc::bulk_g2s(m.wsm, img, C::WSM_W * sizeof(float), m.wsm_full);
                tc::bulk_g2s(m.bias + (g & 1) * C::WSM_B, img + C::WSM_W, C::WSM_B * sizeof(float), m.wsm_full);
            }
        }
        for (int i = 0; i < ksf; ++i, ++it) {
            const int j = kstep_at<C>(i, w.fin >> 4, g > 0);
            const uint32_t st = it % GST<C>;
            if (it >= (uint32_t)GST<C>) tc::mbar_wait(&m.empty[st], ((it / GST<C>) - 1) & 1);
            if (C::CL == 1) {
                tc::mbar_arrive_expect_tx(&m.full[st], slab);
                tc::bulk_g2s(m.ring + (size_t)st * SLOT<C>, WfP + (size_t)j * (slab / 16), slab, &m.full[st]);
            } else {
                // CTA pair: this CTA holds rows [rank * N/2, (rank + 1) * N/2) of the weight; pack.pack_tc_weight_pair
                // stores the slab as [half][hi, lo][2 k-chunks][N/16 row groups x 128 B]: one contiguous copy per CTA
                // (four 2 KB pieces of the single-CTA packing instead were 2.5x slower: measured)
                const uint32_t half = C::NS * 2 * (NBf / 16) * 128;
                tc::mbar_arrive_expect_tx(&m.full[st], half);
                tc::bulk_g2s(m.ring + (size_t)st * SLOT<C>, reinterpret_cast<const unsigned char*>(WfP) + (size_t)j * slab + m.rank * half,
                             half, &m.full[st]);
            }
#ifdef KPD_WS_TRACE
            if (it < 64) tp[it] = clock64();
#endif
        }
        // the gates weight is needed only after this GVP's feats GEMM: queue it behind the slabs
        if (!dead) {
            if (g >= C::WGB) tc::mbar_wait(&m.wg_empty[b], ((g / C::WGB) - 1) & 1);   // gates MMA g-WGB has consumed the buffer
            tc::mbar_arrive_expect_tx(&m.wg_full[b], C::NS * ksg * 512);
            tc::bulk_g2s(m.Wg[0] + b * C::WG_BYTES, C::NS == 2 ? w.WgP2 : w.WgP, C::NS * ksg * 512, &m.wg_full[b]);
        }
    }
#ifdef KPD_WS_TRACE
    if (blockIdx.x == KPD_TRACE_BX && blockIdx.y == 1) {
        for (uint32_t i = 0; i < it && i < 64; ++i) {
            const int k = atomicAdd(&g_ws_trace_n, 1);
            if (k < 2048) { g_ws_trace[3 * k] = 100 + i; g_ws_trace[3 * k + 1] = threadIdx.x >> 5; g_ws_trace[3 * k + 2] = tp[i]; }
        }
    }
#endif
}

// ------------------------------------------------------------------ MMA issuer (one thread)
template <class C>
__device__ __forceinline__ void wait_ready(uint64_t* bar, uint32_t parity) {     // barriers the peer CTA also arrives on
#ifdef KPD_PAIR_CLUSTER_ACQUIRE
    if (C::CL == 2) { tc::mbar_wait_cluster(bar, parity); return; }
#endif
    tc::mbar_wait(bar, parity);
}
// Called by the WHOLE issuing warp when C::CL == 1 (one elected lane executes the tcgen05 instructions, tc::elect_one())
// and by lane 0 only in the CTA-pair configuration.
template <class C>
__device__ __forceinline__ bool issuing_lane() { if constexpr (C::CL == 1) return tc::elect_one(); else return true; }
template <class C>
__device__ __forceinline__ void issue(const GvpW* gv, int n_gvps, Sm& m, uint32_t tmem) {
    uint32_t it = 0;
#ifdef KPD_WS_TRACE
    unsigned long long tk[64];
#endif
    wait_ready<C>(m.feats_ready, 0);
    tc::fence_after_sync();
    WS_TRACE(1);
    for (int g = 0; g < n_gvps; ++g) {
        const GvpW& w = gv[g];
        const int NBf = (w.fout + 15) & ~15, ksf = (w.fin + w.hd + 15) >> 4, ksm = w.fin >> 4, ksg = NBf >> 4;
        constexpr bool PAIR = C::CL == 2;
        const uint32_t idesc = tc::make_idesc_bf16(PAIR ? 2 * C::MMA_M : C::MMA_M, NBf);
        const uint32_t b_k = PAIR ? (NBf / 16) * 128 : (NBf / 8) * 128, slab1 = 2 * b_k;    // (pair: N / 2 rows per CTA)
        // bf16x3: two accumulators, so that the next GVP's k-steps can start while epilogue 1 still reads this one
        const uint32_t acc = tmem + ((C::STACK && (g & 1)) ? 256u : 0u);
        const bool chained = C::STACK && g > 0;
        const int n_first = first_half_ksteps<C>(ksm);
        if (chained) { wait_ready<C>(m.half_ready, (g - 1) & 1); tc::fence_after_sync(); }
        for (int i = 0; i < ksf; ++i, ++it) {
            const int j = kstep_at<C>(i, ksm, g > 0);
            if (chained && i == n_first) {
                // epilogue 1 of the previous GVP done: all of its feats_out is in A
                wait_ready<C>(m.feats_ready, g & 1);
                tc::fence_after_sync();
                WS_TRACE(5);
            }
            if (i == ksm) {
                if (chained && n_first >= ksm) { wait_ready<C>(m.feats_ready, g & 1); tc::fence_after_sync(); }
                WS_TRACE(2); wait_ready<C>(m.tail_ready, g & 1); tc::fence_after_sync(); WS_TRACE(3);
            }
            const uint32_t st = it % GST<C>;
            wait_ready<C>(&m.full[st], (it / GST<C>) & 1);
            tc::fence_after_sync();
#ifdef KPD_WS_TRACE
            if (it < 64) tk[it] = clock64();
#endif
            const uint32_t bs = tc::smem_u32(m.ring + (size_t)st * SLOT<C>);
            const uint64_t a0 = tc::make_smem_desc(tc::smem_u32(m.A[0] + (size_t)2 * j * C::KCS), C::KCS, 128);
            const uint64_t b0 = tc::make_smem_desc(bs, b_k, 128);
            const uint64_t b1 = tc::make_smem_desc(bs + slab1, b_k, 128);
            if (issuing_lane<C>()) {
                if (PAIR) tc::mma_bf16_ss_pair(acc, a0, b0, idesc, i > 0 ? 1u : 0u);
                else tc::mma_bf16_ss(acc, a0, b0, idesc, i > 0 ? 1u : 0u);
                if (C::STACK) {     // [A_hi; A_lo] x W_lo: with the MMA above all four hi/lo products in two instructions
                    if (PAIR) tc::mma_bf16_ss_pair(acc, a0, b1, idesc, 1u);
                    else tc::mma_bf16_ss(acc, a0, b1, idesc, 1u);
                }
                if (PAIR) tc::mma_commit_pair(&m.empty[st], 3);       // frees the slot in both CTAs
                else tc::mma_commit(&m.empty[st]);
            }
        }
        if (issuing_lane<C>()) {
            if (PAIR) tc::mma_commit_pair(m.acc_done, 3);
            else tc::mma_commit(m.acc_done);
        }
        WS_TRACE(4);
        if constexpr (C::STACK) {
            // stacked bf16x3: the gates GEMM runs on the warp-level tensor cores straight from the epilogue registers
            // (gates_mma); the next GVP's k-steps wait for epilogue 1 half by half (above)
            continue;
        }
        // gates GEMM, issued progressively: every epilogue-1 warp writes its feats_out columns in two halves, so the
        // k-steps over the first halves run on the tensor core while the second halves are still being produced
        constexpr int cpw = 256 / C::NCG, hb = cpw / 2;
        const uint32_t idg = tc::make_idesc_bf16(C::MMA_M, 16);
        const uint32_t wg = tc::smem_u32(m.Wg[0] + (g % (C::WGB > 0 ? C::WGB : 1)) * C::WG_BYTES);
        uint32_t gacc = 0u;
        for (int half = 0; half < 2; ++half) {
            if (half == 0) {
                tc::mbar_wait(m.half_ready, g & 1);
                tc::fence_after_sync();
                tc::mbar_wait(&m.wg_full[g % (C::WGB > 0 ? C::WGB : 1)], (g / (C::WGB > 0 ? C::WGB : 1)) & 1);
            } else {
                // epilogue 1 of this GVP done: feats_out is in A, the accumulator columns are free again
                tc::mbar_wait(m.feats_ready, (g + 1) & 1);
                WS_TRACE(5);
            }
            tc::fence_after_sync();
            for (int j = 0; j < ksg; ++j) {
                if ((((16 * j) % cpw) >= hb) != (half == 1)) continue;
                const uint64_t a0 = tc::make_smem_desc(tc::smem_u32(m.A[0] + (size_t)2 * j * C::KCS), C::KCS, 128);
                const uint64_t b0 = tc::make_smem_desc(wg + j * (C::NS * 512), 256, 128);
                if (issuing_lane<C>()) tc::mma_bf16_ss(tmem + GATE_COL, a0, b0, idg, gacc);
                gacc = 1u;
            }
        }
        if (issuing_lane<C>()) {
            tc::mma_commit(m.gates_done);
            tc::mma_commit(&m.wg_empty[g % (C::WGB > 0 ? C::WGB : 1)]);
        }
        WS_TRACE(6);
    }
#ifdef KPD_WS_TRACE
    if (blockIdx.x == KPD_TRACE_BX && blockIdx.y == 1) {
        for (uint32_t i = 0; i < it && i < 64; ++i) {
            const int k = atomicAdd(&g_ws_trace_n, 1);
            if (k < 2048) { g_ws_trace[3 * k] = 200 + i; g_ws_trace[3 * k + 1] = threadIdx.x >> 5; g_ws_trace[3 * k + 2] = tk[i]; }
        }
    }
#endif
}


// ------------------------------------------------------------------ KS (hi / lo planes, 128-row tiles): producer + issuer
// Ring order == consumption order (one FIFO of 8 KB slots):
//   GVP 0:      [W_hi, W_lo] per k-step
//   GVP g > 0:  gates weight of GVP g-1 | [W_hi | W_lo] per k-step
//   at the end: the last GVP's gates weight
// (A chained order -- the next GVP's k-steps over the first halves of the column groups issued behind half_ready into a
//  second accumulator -- was measured SLOWER, 21.4k vs 19.4k cycles per GVP: the tensor pipe executes in issue order, so
//  the gates GEMM, which the SIMT warps wait for, queues behind the early k-steps.)
// The gates weight image is packed with its first-half k-steps first (pack.pack_gates_ks); a gates accumulator lives
// in columns [0, 16) of the accumulator its GVP has just drained.
template <class C>
__device__ __forceinline__ bool gates_first_half(int j) { return ((16 * j) % (256 / C::NCG)) < (256 / C::NCG) / 2; }

template <class C>
__device__ __forceinline__ void produce_ks(const GvpW* gv, int n_gvps, Sm& m) {
    uint32_t it = 0;
    auto push = [&](const void* src, uint32_t bytes) {
        const uint32_t st = it % GST<C>;
        if (it >= (uint32_t)GST<C>) tc::mbar_wait(&m.empty[st], ((it / GST<C>) - 1) & 1);
        tc::mbar_arrive_expect_tx(&m.full[st], bytes);
        tc::bulk_g2s(m.ring + (size_t)st * SLOT<C>, src, bytes, &m.full[st]);
        WS_TRACE(100 + it);
        ++it;
    };
    auto gates_bytes = [&](const GvpW& w, int half) {      // k-steps of that half x (hi 512 B | lo 512 B)
        const int ksg = ((w.fout + 15) & ~15) >> 4;
        int n0 = 0;
        for (int j = 0; j < ksg; ++j) n0 += gates_first_half<C>(j) ? 1 : 0;
        return (uint32_t)((half == 0 ? n0 : ksg - n0) * 1024);
    };
    for (int g = 0; g <= n_gvps; ++g) {
        if (g < n_gvps) {
            // shared-memory image of this GVP's small fp32 weights (see produce())
            const GvpW& w = gv[g];
            if (g > 0) tc::mbar_wait(m.wsm_empty, (g - 1) & 1);
            // (the plain fp32 image for vec_fma follows the fragment image of the other kernels: pack.pack_gvp_small_ks)
            const float* img = w.wsmP2 + C::WSM;
            tc::mbar_arrive_expect_tx(m.wsm_full, (uint32_t)((WSM_KS_W + C::WSM_B) * sizeof(float)));
            tc::bulk_g2s(m.wsm, img, WSM_KS_W * sizeof(float), m.wsm_full);
            tc::bulk_g2s(m.bias + (g & 1) * C::WSM_B, img + WSM_KS_W, C::WSM_B * sizeof(float), m.wsm_full);
        }
        const int ksf = g < n_gvps ? (gv[g].fin + gv[g].hd + 15) >> 4 : 0;
        const int ksm = g < n_gvps ? gv[g].fin >> 4 : 0;
        const uint32_t plane = g < n_gvps ? (uint32_t)(2 * (((gv[g].fout + 15) & ~15) / 8) * 128) : 0u;   // bytes of one plane of a k-step
        // the whole gates weight of GVP g-1 (<= 16 KB) is one ring slot AHEAD of this GVP's k-steps: the issuer holds it
        // until epilogue 1 of GVP g-1 completes and runs the gates GEMM at once then, before the queued k-steps
        if (g > 0) push(gv[g - 1].WgP2c, gates_bytes(gv[g - 1], 0) + gates_bytes(gv[g - 1], 1));
        for (int i = 0; i < ksf; ++i) {
            const int j = i;
            const unsigned char* src = reinterpret_cast<const unsigned char*>(gv[g].WfP2) + (size_t)j * 2 * plane;
            push(src, 2 * plane);
        }
    }
}

template <class C>
__device__ __forceinline__ void issue_ks(const GvpW* gv, int n_gvps, Sm& m, uint32_t tmem) {
    uint32_t it = 0;
    auto take = [&]() -> uint32_t {             // next ring slot, once its copy has landed
        const uint32_t st = it % GST<C>;
        tc::mbar_wait(&m.full[st], (it / GST<C>) & 1);
        tc::fence_after_sync();
        return st;
    };
    // The hi plane of every GVP's feats_out but the last's lives in TENSOR MEMORY (epilogue 1 writes it there with
    // tcgen05.st): an SS-mode MMA fetches its operands from shared memory at ~64 B/clk, so a 128 x 256 x 16 MMA spends
    // ~190 cycles on 12 KB of operands for 128 cycles of math (measured: ~215 cycles per MMA whatever the weight ring
    // does); with A from TMEM only the 8 KB of B remain.  Two of the three MMAs of a k-step, and one of the two of a
    // gates k-step, read A_hi.
    // gates GEMM of GVP gp over the slot `st` that holds its weight, the k-steps of one half of every column group
    // (which = 0 / 1): feats_out (hi, lo planes) x Wg (hi, lo) into TMEM columns [GATE_COL, GATE_COL + 16)
    auto gates = [&](int gp, uint32_t st, int which) {
        const GvpW& w = gv[gp];
        const int ksg = ((w.fout + 15) & ~15) >> 4;
        const bool a_tmem = gp < n_gvps - 1;
        const uint32_t wg = tc::smem_u32(m.ring + (size_t)st * SLOT<C>);
        const uint32_t idg = tc::make_idesc_bf16(C::MMA_M, 16);
        const uint32_t gcol = tmem + GATE_COL;
        int k = 0;
        for (int half = 0; half < 2; ++half)
            for (int j = 0; j < ksg; ++j) {
                if (gates_first_half<C>(j) != (half == 0)) continue;
                if (which != half) { ++k; continue; }
                const uint64_t a1 = tc::make_smem_desc(tc::smem_u32(m.A[1] + (size_t)2 * j * C::KCS), C::KCS, 128);
                // three small MMAs per k-step: A_hi Wg_hi + A_hi Wg_lo (A_hi from tensor memory when another GVP follows) +
                // A_lo Wg_hi.  (Stacking [Wg_hi ; Wg_lo] along N -- two MMAs per k-step, the epilogue summing two column
                // groups -- left the gates GEMM no faster and made the epilogue's TMEM load of the gates ~1.7k cycles
                // slower: measured, reverted.)
                const uint64_t b0 = tc::make_smem_desc(wg + k * 1024, 256, 128), b1 = tc::make_smem_desc(wg + k * 1024 + 512, 256, 128);
                const uint64_t a0 = tc::make_smem_desc(tc::smem_u32(m.A[0] + (size_t)2 * j * C::KCS), C::KCS, 128);
                if (tc::elect_one()) {
                    if (a_tmem) {
                        tc::mma_bf16_ts(gcol, tmem + KS_A_COL + 8 * j, b0, idg, k == 0 ? 0u : 1u);
                        tc::mma_bf16_ts(gcol, tmem + KS_A_COL + 8 * j, b1, idg, 1u);    // (dropping the W_lo term misses the 1e-4 bar: measured)
                    } else {
                        tc::mma_bf16_ss(gcol, a0, b0, idg, k == 0 ? 0u : 1u);
                        tc::mma_bf16_ss(gcol, a0, b1, idg, 1u);
                    }
                    tc::mma_bf16_ss(gcol, a1, b0, idg, 1u);
                }
                ++k;
            }
        WS_TRACE(6 + which);
        if (which == 0) return;
        if (tc::elect_one()) {
            tc::mma_commit(m.gates_done);
            tc::mma_commit(&m.empty[st]);
        }
    };
    tc::mbar_wait(m.feats_ready, 0);
    tc::fence_after_sync();
    for (int g = 0; g < n_gvps; ++g) {
        const GvpW& w = gv[g];
        const int NBf = (w.fout + 15) & ~15, ksf = (w.fin + w.hd + 15) >> 4, ksm = w.fin >> 4;
        const uint32_t idesc = tc::make_idesc_bf16(C::MMA_M, NBf);
        const uint32_t b_k = (NBf / 8) * 128;
        if (g > 0) {
            // the gates GEMM of GVP g-1 progressively behind the two halves of its epilogue 1, then this GVP (by then the
            // accumulator has been drained and all of feats_out is in place)
            const uint32_t gst = take();
            ++it;
            tc::mbar_wait(m.half_ready, (g - 1) & 1);
            tc::fence_after_sync();
            gates(g - 1, gst, 0);
            tc::mbar_wait(m.feats_ready, g & 1);
            tc::fence_after_sync();
            gates(g - 1, gst, 1);
        }
        for (int i = 0; i < ksf; ++i) {
            if (i == ksm) {
                WS_TRACE(2);
                tc::mbar_wait(m.tail_ready, g & 1);
                tc::fence_after_sync();
                WS_TRACE(3);
            }
            const bool a_tmem = g > 0 && i < ksm;       // (the |Vh| tail columns are written to shared memory)
            const uint64_t a0 = tc::make_smem_desc(tc::smem_u32(m.A[0] + (size_t)2 * i * C::KCS), C::KCS, 128);
            const uint64_t a1 = tc::make_smem_desc(tc::smem_u32(m.A[1] + (size_t)2 * i * C::KCS), C::KCS, 128);
            const uint32_t st = take();         // [W_hi | W_lo] of this k-step
            WS_TRACE(200 + it);
            const uint32_t bs = tc::smem_u32(m.ring + (size_t)st * SLOT<C>);
            const uint64_t b0 = tc::make_smem_desc(bs, b_k, 128), b1 = tc::make_smem_desc(bs + 2 * b_k, b_k, 128);
            if (tc::elect_one()) {
                if (a_tmem) {
                    const uint32_t at = tmem + KS_A_COL + 8 * i;
                    tc::mma_bf16_ts(tmem, at, b0, idesc, i > 0 ? 1u : 0u);  // A_hi W_hi
                    tc::mma_bf16_ss(tmem, a1, b0, idesc, 1u);               // + A_lo W_hi
                    tc::mma_bf16_ts(tmem, at, b1, idesc, 1u);               // + A_hi W_lo (lo x lo is below the fp32 rounding of the sum)
                } else {
                    tc::mma_bf16_ss(tmem, a0, b0, idesc, i > 0 ? 1u : 0u);
                    tc::mma_bf16_ss(tmem, a1, b0, idesc, 1u);
                    tc::mma_bf16_ss(tmem, a0, b1, idesc, 1u);
                }
                tc::mma_commit(&m.empty[st]);
            }
            ++it;
        }
        if (tc::elect_one()) tc::mma_commit(m.acc_done);
        WS_TRACE(4);
    }
    // the last GVP's gates
    const uint32_t gst = take();
    ++it;
    tc::mbar_wait(m.half_ready, (n_gvps - 1) & 1);
    tc::fence_after_sync();
    gates(n_gvps - 1, gst, 0);
    tc::mbar_wait(m.feats_ready, n_gvps & 1);
    tc::fence_after_sync();
    gates(n_gvps - 1, gst, 1);
}

// CTA pairs: the peer's control thread tells the leader's MMA thread when the peer's half of a weight slab has landed
// (remote arrive on the leader's full barrier, which counts its own producer + this relay).
template <class C>
__device__ __forceinline__ void relay(const GvpW* gv, int n_gvps, Sm& m) {
    uint32_t it = 0;
    for (int g = 0; g < n_gvps; ++g) {
        const int ksf = (gv[g].fin + gv[g].hd + 15) >> 4;
        for (int j = 0; j < ksf; ++j, ++it) {
            const uint32_t st = it % GST<C>;
            tc::mbar_wait(&m.full[st], (it / GST<C>) & 1);
            tc::mbar_arrive_remote_relaxed(tc::map_to_cta(&m.full[st], 0));
        }
    }
}

// CTA pairs: one thread of the peer forwards the phases of its CTA's tail_ready / half_ready / feats_ready barriers to
// the leader's (in the order the SIMT warps complete them; a phase cannot complete twice before the leader has
// consumed the forward, so the parities cannot alias).
template <class C>
__device__ __forceinline__ void forward_ready(int n_gvps, Sm& m) {
    const uint32_t feats = tc::map_to_cta(m.feats_ready, 0), tail = tc::map_to_cta(m.tail_ready, 0), half = tc::map_to_cta(m.half_ready, 0);
    tc::mbar_wait(m.feats_ready, 0);
    tc::mbar_arrive_remote_relaxed(feats);
    for (int g = 0; g < n_gvps; ++g) {
        tc::mbar_wait(m.tail_ready, g & 1);
        tc::mbar_arrive_remote_relaxed(tail);
        tc::mbar_wait(m.half_ready, g & 1);
        tc::mbar_arrive_remote_relaxed(half);
        tc::mbar_wait(m.feats_ready, (g + 1) & 1);
        tc::mbar_arrive_remote_relaxed(feats);
    }
}

// ------------------------------------------------------------------ SIMT helpers
// store one value into the bf16 plane(s) at (row, col)
template <class C>
__device__ __forceinline__ void put_scalar(const Sm& m, int row, int col, float x) {
    const uint32_t off = (uint32_t)((col >> 3) * C::KCS + (col & 7) * 2) + row_off<C>(row);
    const __nv_bfloat16 hi = __float2bfloat16(x);
    *reinterpret_cast<__nv_bfloat16*>(m.A[0] + off) = hi;
    if (C::NS == 2) *reinterpret_cast<__nv_bfloat16*>(m.A[1] + off) = __float2bfloat16(x - __bfloat162float(hi));
}

// 8 consecutive columns (one k-chunk) of one row
template <class C>
__device__ __forceinline__ void put_chunk(const Sm& m, int row, int kc, const float (&f)[8]) {
    const uint32_t off = (uint32_t)(kc * C::KCS) + row_off<C>(row);
    uint4 hi;
    hi.x = tc::pack_bf16x2(f[0], f[1]); hi.y = tc::pack_bf16x2(f[2], f[3]);
    hi.z = tc::pack_bf16x2(f[4], f[5]); hi.w = tc::pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(m.A[0] + off) = hi;
    if (C::NS == 2) {
        uint4 lo;
        lo.x = tc::pack_bf16x2(f[0] - __uint_as_float(hi.x << 16), f[1] - __uint_as_float(hi.x & 0xffff0000u));
        lo.y = tc::pack_bf16x2(f[2] - __uint_as_float(hi.y << 16), f[3] - __uint_as_float(hi.y & 0xffff0000u));
        lo.z = tc::pack_bf16x2(f[4] - __uint_as_float(hi.z << 16), f[5] - __uint_as_float(hi.z & 0xffff0000u));
        lo.w = tc::pack_bf16x2(f[6] - __uint_as_float(hi.w << 16), f[7] - __uint_as_float(hi.w & 0xffff0000u));
        *reinterpret_cast<uint4*>(m.A[1] + off) = lo;
    }
}

// value of (row, col) read back from the plane(s)
template <class C>
__device__ __forceinline__ float get_scalar(const Sm& m, int row, int col) {
    const uint32_t off = (uint32_t)((col >> 3) * C::KCS + (col & 7) * 2) + row_off<C>(row);
    float x = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(m.A[0] + off));
    if (C::NS == 2) x += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(m.A[1] + off));
    return x;
}

// all SIMT lanes made their shared-memory writes -> publish to the async proxy and arrive (one per warp)
__device__ __forceinline__ void publish(uint64_t* bar) {
    tc::fence_proxy_async();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) tc::mbar_arrive(bar);
}

// same, on a barrier the MMA thread waits on.  In a CTA pair that thread lives in the leader CTA; the peer's warps still
// arrive on their own CTA's barrier and ONE peer thread forwards each completed phase (forward_ready): a remote
// release-arrive per warp would wait for the SM's in-flight bulk copies every time (measured: +15 % epilogue time).
template <class C>
__device__ __forceinline__ void publish_mma(const Sm& m, uint64_t* bar) { publish(bar); }

// Register-resident vectors.  The 8 tile rows of a SIMT warp x 3 components form the 24 (of 32) rows of two
// m16n8k8 MMA tiles, ordered component-major: tile 0 rows 0-7 = x, rows 8-15 = y; tile 1 rows 0-7 = z.  Lane
// (g = lane / 4, t = lane % 4) therefore holds, for tile row 8 * warp + g and ALL THREE components, the vector
// channels 8 s + 2 t + e (s = k-step / n-tile, e = 0, 1): exactly the accumulator fragment of one GEMM and, with the
// K index permuted the same way in the staged weights, the A fragment of the next (no shuffles, no shared memory).
struct VF { float x[3][3][2]; };     // [component][s][e]
struct Lane {
    int row;           // tile row of this lane
    int t;             // lane % 4
    uint32_t rowoff;   // byte offset of the row inside a k-chunk of A
};
template <class C>
__device__ __forceinline__ Lane lane_geometry() {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Lane L;
    // KS: warp w owns the 8 rows [32 (w % 4) + 8 (w / 4), + 8) -- inside the TMEM lane quarter the warp may read, so that
    // epilogue 2 can take its gates straight from TMEM in this fragment layout (gates_from_tmem)
    L.row = C::KS ? 32 * (warp & 3) + 8 * (warp >> 2) + (lane >> 2) : 8 * warp + (lane >> 2);
    L.t = lane & 3;
    L.rowoff = row_off<C>(L.row);
    return L;
}
__device__ __forceinline__ void vf_zero(VF& v) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int s = 0; s < 3; ++s) v.x[c][s][0] = v.x[c][s][1] = 0.f;
}
// channels of this lane from / to a [nv][3] row in global or shared memory (channels >= nv read as 0)
__device__ __forceinline__ void vf_load(VF& v, const float* __restrict__ p, int nv, int t) {
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int u = 8 * s + 2 * t + e;
#pragma unroll
            for (int c = 0; c < 3; ++c) v.x[c][s][e] = u < nv ? p[3 * u + c] : 0.f;
        }
#pragma unroll
    for (int c = 0; c < 3; ++c) v.x[c][2][0] = v.x[c][2][1] = 0.f;
}
__device__ __forceinline__ void vf_store(const VF& v, float* p, int nv, int t) {
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int u = 8 * s + 2 * t + e;
            if (u < nv) {
#pragma unroll
                for (int c = 0; c < 3; ++c) p[3 * u + c] = v.x[c][s][e];
            }
        }
}
// D[c][j][e] += sum_k A[c][k-step][.] W[k][8 j + 2 t + e]: one GEMM of the vector path on the warp-level tensor cores.
// W: staged B fragments (hi at W, lo at W + lo_off for the 3xTF32 mode); nks k-steps, NT n-tiles (the last only if nt3).
template <int NS, int LD, int NT>
__device__ __forceinline__ void vec_gemm(const float (&A)[3][3][2], float (&D)[3][NT][2], const float* __restrict__ W, int lo_off,
                                         int nks, bool nt_last, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int j = 0; j < NT; ++j) D[c][j][0] = D[c][j][1] = 0.f;
    float junk0 = 0.f, junk1 = 0.f;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        if (s < nks) {
            float ah[3][2], al[3][2];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (NS == 2) { ah[c][e] = tf32_rna(A[c][s][e]); al[c][e] = tf32_rna(A[c][s][e] - ah[c][e]); }
                    else ah[c][e] = tf32_rna(A[c][s][e]);   // round (the tensor core would truncate the low mantissa bits)
                }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                if (j + 1 < NT || nt_last) {
                    const float2 b = *reinterpret_cast<const float2*>(W + ((4 * s + t) * LD + 8 * j + g) * 2);
                    // tile 0: rows 0-7 = component 0, rows 8-15 = component 1; tile 1: rows 0-7 = component 2
                    mma_tf32(D[0][j][0], D[0][j][1], D[1][j][0], D[1][j][1], ah[0][0], ah[1][0], ah[0][1], ah[1][1], b.x, b.y);
                    mma_tf32(D[2][j][0], D[2][j][1], junk0, junk1, ah[2][0], 0.f, ah[2][1], 0.f, b.x, b.y);
                    if (NS == 2) {
                        const float2 bl = *reinterpret_cast<const float2*>(W + lo_off + ((4 * s + t) * LD + 8 * j + g) * 2);
                        mma_tf32(D[0][j][0], D[0][j][1], D[1][j][0], D[1][j][1], al[0][0], al[1][0], al[0][1], al[1][1], b.x, b.y);
                        mma_tf32(D[2][j][0], D[2][j][1], junk0, junk1, al[2][0], 0.f, al[2][1], 0.f, b.x, b.y);
                        mma_tf32(D[0][j][0], D[0][j][1], D[1][j][0], D[1][j][1], ah[0][0], ah[1][0], ah[0][1], ah[1][1], bl.x, bl.y);
                        mma_tf32(D[2][j][0], D[2][j][1], junk0, junk1, ah[2][0], 0.f, ah[2][1], 0.f, bl.x, bl.y);
                    }
                }
            }
        }
    }
}

// The same GEMM on the FP32 pipe (KS edge kernel).  The legacy warp-level tensor-core path shares the tensor cores with
// tcgen05.mma and only gets the gaps of a busy UTCMMA stream: with the main GEMM of the chain running underneath, the
// mma.sync vector GEMMs took 2-3x their stand-alone time (measured: Vh + Vu 10.8k instead of 4.4k cycles per GVP).  Here
// every lane multiplies its own input channels (8 s + 2 t + e) into ALL eight outputs of a block, and a reduce-scatter
// over the four lanes of a row (two shuffle rounds) leaves each lane with the outputs it owns (8 ob + 2 t + e) -- exact
// fp32, no split operands.  W: plain fp32 [24][LD] in shared memory (pack.pack_gvp_small_ks; LD = 28 / 20 keeps the
// four distinct rows a warp instruction reads on distinct banks).
template <int LD, int NT>
__device__ __forceinline__ void vec_fma(const float (&A)[3][3][2], float (&D)[3][NT][2], const float* __restrict__ W, int nks,
                                        bool nt_last, int lane) {
    const int t = lane & 3;
#pragma unroll
    for (int ob = 0; ob < NT; ++ob) {
        float acc[3][8];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[c][k] = 0.f;
        if (ob + 1 < NT || nt_last) {
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                if (s < nks) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float* wr = W + (8 * s + 2 * t + e) * LD + 8 * ob;
                        const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + 4);
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float a = A[c][s][e];
                            acc[c][0] = fmaf(a, w0.x, acc[c][0]); acc[c][1] = fmaf(a, w0.y, acc[c][1]);
                            acc[c][2] = fmaf(a, w0.z, acc[c][2]); acc[c][3] = fmaf(a, w0.w, acc[c][3]);
                            acc[c][4] = fmaf(a, w1.x, acc[c][4]); acc[c][5] = fmaf(a, w1.y, acc[c][5]);
                            acc[c][6] = fmaf(a, w1.z, acc[c][6]); acc[c][7] = fmaf(a, w1.w, acc[c][7]);
                        }
                    }
                }
            }
            // reduce-scatter over the four lanes of the row: lanes t = 0, 1 end up with outputs 0-3, lanes 2, 3 with 4-7 ...
            const bool up = (t & 2) != 0, odd = (t & 1) != 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float keep[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float give = up ? acc[c][k] : acc[c][k + 4];
                    keep[k] = (up ? acc[c][k + 4] : acc[c][k]) + __shfl_xor_sync(0xffffffffu, give, 2);
                }
                // ... then even lanes with the first pair of their four, odd lanes with the second
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float give = odd ? keep[k] : keep[k + 2];
                    D[c][ob][k] = (odd ? keep[k + 2] : keep[k]) + __shfl_xor_sync(0xffffffffu, give, 1);
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) D[c][ob][0] = D[c][ob][1] = 0.f;
        }
    }
}

// one 32-column chunk of epilogue 1: bias + SiLU -> bf16 plane(s) of A (static register indexing only)
template <class C>
__device__ __forceinline__ void epi1_chunk(const uint32_t (&v)[32], int c0, int fout, int NBf, const float* bf_s,
                                           const Sm& m, int row) {
    // fout is a multiple of 8 (host-checked): a chunk is either all feats_out or all padding
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
        if (c0 + 8 * kc < NBf) {
            float f[8];
            if (c0 + 8 * kc < fout) {
                const float4 b0 = *reinterpret_cast<const float4*>(bf_s + c0 + 8 * kc);
                const float4 b1 = *reinterpret_cast<const float4*>(bf_s + c0 + 8 * kc + 4);
                f[0] = act_silu<C::NS>(__uint_as_float(v[8 * kc + 0]) + b0.x); f[1] = act_silu<C::NS>(__uint_as_float(v[8 * kc + 1]) + b0.y);
                f[2] = act_silu<C::NS>(__uint_as_float(v[8 * kc + 2]) + b0.z); f[3] = act_silu<C::NS>(__uint_as_float(v[8 * kc + 3]) + b0.w);
                f[4] = act_silu<C::NS>(__uint_as_float(v[8 * kc + 4]) + b1.x); f[5] = act_silu<C::NS>(__uint_as_float(v[8 * kc + 5]) + b1.y);
                f[6] = act_silu<C::NS>(__uint_as_float(v[8 * kc + 6]) + b1.z); f[7] = act_silu<C::NS>(__uint_as_float(v[8 * kc + 7]) + b1.w);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = 0.f;
            }
            put_chunk<C>(m, row, (c0 >> 3) + kc, f);
        }
    }
}

// same for the KS edge kernel when another GVP follows: the hi plane goes to TENSOR memory -- hw[] receives the 16 packed
// bf16 pairs of the 32 columns, for one tcgen05.st by the caller --, the lo plane to shared memory
template <class C>
__device__ __forceinline__ void epi1_chunk_ks(const uint32_t (&v)[32], int c0, int fout, int NBf, const float* bf_s,
                                              const Sm& m, int row, uint32_t (&hw)[16]) {
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
        uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = make_uint4(0u, 0u, 0u, 0u);
        if (c0 + 8 * kc < fout) {
            const float4 b0 = *reinterpret_cast<const float4*>(bf_s + c0 + 8 * kc);
            const float4 b1 = *reinterpret_cast<const float4*>(bf_s + c0 + 8 * kc + 4);
            float f[8];
            f[0] = act_silu<C::NS>(__uint_as_float(v[8 * kc + 0]) + b0.x); f[1] = act_silu<C::NS>(__uint_as_float(v[8 * kc + 1]) + b0.y);
            f[2] = act_silu<C::NS>(__uint_as_float(v[8 * kc + 2]) + b0.z); f[3] = act_silu<C::NS>(__uint_as_float(v[8 * kc + 3]) + b0.w);
            f[4] = act_silu<C::NS>(__uint_as_float(v[8 * kc + 4]) + b1.x); f[5] = act_silu<C::NS>(__uint_as_float(v[8 * kc + 5]) + b1.y);
            f[6] = act_silu<C::NS>(__uint_as_float(v[8 * kc + 6]) + b1.z); f[7] = act_silu<C::NS>(__uint_as_float(v[8 * kc + 7]) + b1.w);
            split8(f, hi, lo);
        }
        hw[4 * kc + 0] = hi.x; hw[4 * kc + 1] = hi.y; hw[4 * kc + 2] = hi.z; hw[4 * kc + 3] = hi.w;
        if (c0 + 8 * kc < NBf) *reinterpret_cast<uint4*>(m.A[1] + (uint32_t)(((c0 >> 3) + kc) * C::KCS) + row_off<C>(row)) = lo;
    }
}

// M = 64 accumulators: 64 columns loaded with the 16x256b shape (all 32 lanes hold data: rows t/4 and t/4 + 8 of the
// warp's 16-lane TMEM quarter, column pairs 2(t%4) + 8i).  bias + SiLU -> bf16x2 -> A plane(s), 4-byte stores.
// fr (NS = 2 only): the packed bf16 pairs just written, [hi row a, hi row b, lo row a, lo row b][block] -- as they are,
// the mma.sync A fragments of the gates GEMM (gates_mma)
template <class C, int NBLK>
__device__ __forceinline__ void epi1_frag64(const uint32_t (&v)[4 * NBLK], int c0, int fout, int NBf, const float* bf_s,
                                            const Sm& m, int row_a, int lane, uint32_t (&fr)[4][NBLK]) {
    const int cp = 2 * (lane & 3);
    const uint32_t ro = row_off<C>(row_a) + cp * 2;     // row_a + 8 is the next 8-row group: + 128 bytes
#pragma unroll
    for (int i = 0; i < NBLK; ++i) {
        const int col = c0 + 8 * i;
        if (col < NBf) {
            float fa0 = 0.f, fa1 = 0.f, fb0 = 0.f, fb1 = 0.f;
            if (col < fout) {
                const float2 b = *reinterpret_cast<const float2*>(bf_s + col + cp);
                fa0 = act_silu<C::NS>(__uint_as_float(v[4 * i + 0]) + b.x); fa1 = act_silu<C::NS>(__uint_as_float(v[4 * i + 1]) + b.y);
                fb0 = act_silu<C::NS>(__uint_as_float(v[4 * i + 2]) + b.x); fb1 = act_silu<C::NS>(__uint_as_float(v[4 * i + 3]) + b.y);
            }
            const uint32_t off = (uint32_t)((col >> 3) * C::KCS) + ro;
            const uint32_t ha = tc::pack_bf16x2(fa0, fa1), hb = tc::pack_bf16x2(fb0, fb1);
            *reinterpret_cast<uint32_t*>(m.A[0] + off) = ha;
            *reinterpret_cast<uint32_t*>(m.A[0] + off + 128) = hb;       // row + 8: next 8-row group
            fr[0][i] = ha; fr[1][i] = hb; fr[2][i] = 0u; fr[3][i] = 0u;
            if (C::NS == 2) {
                const uint32_t la = tc::pack_bf16x2(fa0 - __uint_as_float(ha << 16), fa1 - __uint_as_float(ha & 0xffff0000u));
                const uint32_t lb = tc::pack_bf16x2(fb0 - __uint_as_float(hb << 16), fb1 - __uint_as_float(hb & 0xffff0000u));
                *reinterpret_cast<uint32_t*>(m.A[1] + off) = la;
                *reinterpret_cast<uint32_t*>(m.A[1] + off + 128) = lb;
                fr[2][i] = la; fr[3][i] = lb;
            }
        } else {
            fr[0][i] = fr[1][i] = fr[2][i] = fr[3][i] = 0u;
        }
    }
}

// warp-level bf16 tensor-core MMA, D(16x8) += A(16x16) B(16x8), fp32 accumulation
__device__ __forceinline__ void mma_bf16_k16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                             uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Gates GEMM of the bf16x3 mode on the warp-level tensor cores, straight from the epilogue registers: the 16x256b
// TMEM fragment of epilogue 1 (rows g, g+8; column pairs 2t + 8i) IS the m16n8k16 A fragment layout, so feats_out never
// has to be re-read.  gD[j]: this warp's partial gates (16 rows x n-tile j) over its own feats_out columns.
// Wg: staged fragments (pack.pack_gates_frag): [hi | lo][k16 step][t][(n + 4t) & 15][2 words].
template <int NBLK>
__device__ __forceinline__ void gates_mma(const uint32_t (&fr)[4][NBLK], int c0, int NBf, const uint32_t* __restrict__ Wg,
                                          int lo_words, int lane, float (&gD)[2][4]) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int p = 0; p < NBLK / 2; ++p) {
        const int col = c0 + 16 * p;
        if (col < NBf) {
            const int S = col >> 4;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int wi = ((S * 4 + t) * 16 + ((8 * j + g + 4 * t) & 15)) * 2;
                const uint2 bh = *reinterpret_cast<const uint2*>(Wg + wi);
                const uint2 bl = *reinterpret_cast<const uint2*>(Wg + lo_words + wi);
                mma_bf16_k16(gD[j], fr[0][2 * p], fr[1][2 * p], fr[0][2 * p + 1], fr[1][2 * p + 1], bh.x, bh.y);
                mma_bf16_k16(gD[j], fr[2][2 * p], fr[3][2 * p], fr[2][2 * p + 1], fr[3][2 * p + 1], bh.x, bh.y);
                mma_bf16_k16(gD[j], fr[0][2 * p], fr[1][2 * p], fr[0][2 * p + 1], fr[1][2 * p + 1], bl.x, bl.y);
            }
        }
    }
}

// GVP.forward (models/gvp.py:89-116), SIMT side, for GVP number gi of the chain.
// In: v[0..vin) (registers), feats in A[:, 0:fin) already published.  Out: v[0..vout), feats_out in A[:, 0:fout)
// published through feats_ready.
template <class C>
__device__ __forceinline__ void gvp_simt(const GvpW& g, int gi, const Sm& m, uint32_t tmem, VF& v, const Lane& L,
                                         int rows_valid, int tb, bool last = true) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* Wh_s = m.wsm;
    const float* Wu_s = Wh_s + (C::NS == 2 ? 24 * WH_LD_KS : C::NS * WH_SZ);
    const float* bf_s = m.bias + (gi & 1) * C::WSM_B;
    const float* bg_s = bf_s + 256;
    const int NBf = (g.fout + 15) & ~15;
    // a. Vh = V^T Wh (gvp.py:96) on the warp-level tensor cores; sh = sqrt(clamp(|Vh|^2)) -> A[:, fin + h] (gvp.py:99)
    TC_T(t0);
    WS_TRACE(10);
    tc::mbar_wait(m.wsm_full, gi & 1);
    TC_T(t0b);
    WS_ACC(tb + 7, t0, t0b);                 // (of the Vh phase: waiting for this GVP's small weights)
    float vh[3][3][2];
    float vu[3][2][2];
    [[maybe_unused]] float gD[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};     // bf16x3: this warp's partial gates
    const bool vecw = warp < C::NWV;          // (warp-uniform) this warp owns vector rows
    if (vecw) {
    if constexpr (C::NS == 2) vec_fma<WH_LD_KS, 3>(v.x, vh, Wh_s, g.vin > 16 ? 3 : 2, g.hd > 16, lane);
    else vec_gemm<C::NS, WH_LD, 3>(v.x, vh, Wh_s, WH_SZ, g.vin > 16 ? 3 : 2, g.hd > 16, lane);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int h = 8 * j + 2 * L.t;
        if (h < g.hd) {       // both columns of the pair are written; a column >= hd meets zero weight rows
            const float s0 = vh[0][j][0] * vh[0][j][0] + vh[1][j][0] * vh[1][j][0] + vh[2][j][0] * vh[2][j][0];
            const float s1 = vh[0][j][1] * vh[0][j][1] + vh[1][j][1] * vh[1][j][1] + vh[2][j][1] * vh[2][j][1];
            const float x0 = sqrt_fast(fmaxf(s0, 1e-8f)), x1 = sqrt_fast(fmaxf(s1, 1e-8f));
            const int col = g.fin + h;
            const uint32_t off = (uint32_t)((col >> 3) * C::KCS + (col & 7) * 2) + L.rowoff;
            const uint32_t hi = tc::pack_bf16x2(x0, x1);
            *reinterpret_cast<uint32_t*>(m.A[0] + off) = hi;
            if (C::NS == 2)
                *reinterpret_cast<uint32_t*>(m.A[1] + off) =
                    tc::pack_bf16x2(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xffff0000u));
        }
    }
    publish_mma<C>(m, m.tail_ready);
    }
    TC_T(t1);
    WS_TRACE(11);
    // b. Vu = Vh^T Wu (gvp.py:97), while the tensor core finishes the feats GEMM
    if (vecw) {
        if constexpr (C::NS == 2) vec_fma<WU_LD_KS, 2>(vh, vu, Wu_s, g.hd > 16 ? 3 : 2, true, lane);
        else vec_gemm<C::NS, WU_LD, 2>(vh, vu, Wu_s, WU_SZ, g.hd > 16 ? 3 : 2, true, lane);
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(m.wsm_empty);     // Wh | Wu may be overwritten with the next GVP's
    }
    // c. epilogue 1: feats_out = SiLU(acc + b) -> A[:, 0:fout)   (gvp.py:101-103)
    const int q = warp & 3, cg = warp >> 2;
    const int row_e = C::R == 128 ? 32 * q + lane : 16 * q + (lane & 15);   // M = 64: lanes 0-15 of each TMEM quarter
    const bool valid_e = C::R == 128 ? true : lane < 16;
    const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + ((C::STACK && (gi & 1)) ? 256u : 0u);
    TC_T(t2);
    WS_TRACE(12);
    tc::mbar_wait(m.acc_done, gi & 1);
    tc::fence_after_sync();
    TC_T(t3);
    WS_TRACE(13);
    {
        constexpr int cpw = 256 / C::NCG, hb = cpw / 2;      // the warp's columns, written in two halves (see issue())
        const int row0 = C::R == 128 ? 32 * q : 16 * q;
        [[maybe_unused]] uint32_t fr8[4][8];
        [[maybe_unused]] uint32_t fr4[4][4];
        if constexpr (C::STACK) {                            // the gates weight fragments of this GVP have landed
            tc::mbar_wait(&m.wg_full[gi % C::WGB], (gi / C::WGB) & 1);
        }
        [[maybe_unused]] const uint32_t* wgf = reinterpret_cast<const uint32_t*>(m.Wg[0] + (gi % (C::WGB > 0 ? C::WGB : 1)) * C::WG_BYTES);
        const int wg_lo = ((NBf >> 4) * 512) / 4;            // words between the hi and the lo fragment image
        const int cend = row0 < rows_valid ? min(NBf, cg * cpw + cpw) : 0;   // warps of empty row quarters skip
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int cb = cg * cpw + half * hb;
            if (cb < cend) {
                if constexpr (C::R == 128) {                 // hb = 32 columns: thread = row
                    uint32_t v0[32];
                    tc::tmem_ld_x32(taddr + cb, v0);
                    tc::tmem_ld_wait();
                    if (C::KS && !last) {
                        // hi plane -> tensor memory (the next GVP's A operand, issue_ks()), lo plane -> shared memory
                        uint32_t hw[16];
                        epi1_chunk_ks<C>(v0, cb, g.fout, NBf, bf_s, m, row_e, hw);
                        tc::tmem_st_x16(tmem + ((uint32_t)(32 * q) << 16) + KS_A_COL + (cb >> 1), hw);
                    } else {
                        epi1_chunk<C>(v0, cb, g.fout, NBf, bf_s, m, row_e);
                    }
                } else if constexpr (C::NS == 1 && hb == 64) {           // hb = 64 columns, M = 64 accumulator
                    uint32_t v0[32];
                    tc::tmem_ld_16x256b_x8(taddr + cb, v0);
                    tc::tmem_ld_wait();
                    epi1_frag64<C, 8>(v0, cb, g.fout, NBf, bf_s, m, 16 * q + (lane >> 2), lane, fr8);
                } else if constexpr (C::NS == 1) {                        // 16 SIMT warps: 32 columns per half
                    uint32_t v0[16];
                    tc::tmem_ld_16x256b_x4(taddr + cb, v0);
                    tc::tmem_ld_wait();
                    epi1_frag64<C, 4>(v0, cb, g.fout, NBf, bf_s, m, 16 * q + (lane >> 2), lane, fr4);
                } else if constexpr (hb == 64) {
                    // stacked operand: lanes [32q, 32q+16) = A_hi (W_hi + W_lo), lanes [32q+16, 32q+32) = A_lo (W_hi + W_lo)
                    uint32_t v0[32], v1[32];
                    tc::tmem_ld_16x256b_x8(taddr + cb, v0);
                    tc::tmem_ld_16x256b_x8(taddr + (16u << 16) + cb, v1);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) v0[i] = __float_as_uint(__uint_as_float(v0[i]) + __uint_as_float(v1[i]));
                    epi1_frag64<C, 8>(v0, cb, g.fout, NBf, bf_s, m, 16 * q + (lane >> 2), lane, fr8);
                    gates_mma<8>(fr8, cb, NBf, wgf, wg_lo, lane, gD);
                } else {                                       // 16 SIMT warps: 32 columns per half
                    uint32_t v0[16], v1[16];
                    tc::tmem_ld_16x256b_x4(taddr + cb, v0);
                    tc::tmem_ld_16x256b_x4(taddr + (16u << 16) + cb, v1);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) v0[i] = __float_as_uint(__uint_as_float(v0[i]) + __uint_as_float(v1[i]));
                    epi1_frag64<C, 4>(v0, cb, g.fout, NBf, bf_s, m, 16 * q + (lane >> 2), lane, fr4);
                    gates_mma<4>(fr4, cb, NBf, wgf, wg_lo, lane, gD);
                }
            }
            if (C::KS && !last) { tc::tmem_st_wait(); tc::fence_before_sync(); }
            if (half == 0) publish_mma<C>(m, m.half_ready);
        }
    }
    if constexpr (C::STACK) {
        // partial gates of this warp (its 16 rows x its feats_out columns) -> shared memory; the buffer of the gates
        // weight is free again
        const int ra = 16 * q + (lane >> 2);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float* gp = m.gate + ((size_t)cg * C::R + ra) * 16 + 8 * j + 2 * (lane & 3);
            *reinterpret_cast<float2*>(gp) = make_float2(gD[j][0], gD[j][1]);
            *reinterpret_cast<float2*>(gp + 8 * 16) = make_float2(gD[j][2], gD[j][3]);
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&m.wg_empty[gi % C::WGB]);
    }
    tc::fence_before_sync();
    publish_mma<C>(m, m.feats_ready);
    TC_T(t4);
    WS_TRACE(14);
    // d. epilogue 2: vectors_out = act(gating) * Vu   (gvp.py:105-111)
    if constexpr (C::KS) {
        // the gates accumulator (TMEM columns GATE_COL .. GATE_COL + 16) read with the 16x256b shape: lane (g, t) gets
        // rows g and g + 8 of a 16-lane window, columns 8 j + 2 t + e -- exactly the channels its vector fragment holds;
        // no staging through shared memory, no SIMT-wide barrier
        tc::mbar_wait(m.gates_done, gi & 1);
        tc::fence_after_sync();
        TC_T(t5);
        {
            const int sub = warp >> 2;                       // which 8 rows of the warp's TMEM lane quarter
            uint32_t gv[8];
            tc::tmem_ld_16x256b_x2(tmem + ((uint32_t)(32 * q + 16 * (sub >> 1)) << 16) + GATE_COL, gv);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int u = 8 * j + 2 * L.t;
                const int o = (sub & 1) ? 2 : 0;
                float a0 = __uint_as_float(gv[4 * j + o]) + bg_s[u];
                float a1 = __uint_as_float(gv[4 * j + o + 1]) + bg_s[u + 1];
                if (g.sigmoid_gate) { a0 = sigmoid_acc(a0); a1 = sigmoid_acc(a1); }
#pragma unroll
                for (int c = 0; c < 3; ++c) { v.x[c][j][0] = vu[c][j][0] * a0; v.x[c][j][1] = vu[c][j][1] * a1; }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) v.x[c][2][0] = v.x[c][2][1] = 0.f;
        }
        tc::fence_before_sync();            // the next gates GEMM overwrites these TMEM columns (ordered through half_ready)
        TC_T(t6);
        WS_ACC(tb + 0, t0, t1); WS_ACC(tb + 1, t1, t2); WS_ACC(tb + 2, t2, t3); WS_ACC(tb + 3, t3, t4); WS_ACC(tb + 4, t4, t5);
        WS_ACC(tb + 5, t5, t6); WS_ACC(tb + 6, 0, 1);
        return;
    }
    if constexpr (C::STACK) {
        TC_T(t5);
        simt_bar<C>();
        if (vecw) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int u = 8 * j + 2 * L.t;
                float a0 = bg_s[u], a1 = bg_s[u + 1];
#pragma unroll
                for (int c2 = 0; c2 < C::NCG; ++c2) {
                    const float2 pp = *reinterpret_cast<const float2*>(m.gate + ((size_t)c2 * C::R + L.row) * 16 + u);
                    a0 += pp.x; a1 += pp.y;
                }
                if (g.sigmoid_gate) { a0 = sigmoid_acc(a0); a1 = sigmoid_acc(a1); }
#pragma unroll
                for (int c = 0; c < 3; ++c) { v.x[c][j][0] = vu[c][j][0] * a0; v.x[c][j][1] = vu[c][j][1] * a1; }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) v.x[c][2][0] = v.x[c][2][1] = 0.f;
        }
        TC_T(t6);
        WS_ACC(tb + 0, t0, t1); WS_ACC(tb + 1, t1, t2); WS_ACC(tb + 2, t2, t3); WS_ACC(tb + 3, t3, t4); WS_ACC(tb + 4, t4, t5);
        WS_ACC(tb + 5, t5, t6); WS_ACC(tb + 6, 0, 1);
        return;
    }
    tc::mbar_wait(m.gates_done, gi & 1);
    tc::fence_after_sync();
    TC_T(t5);
    WS_TRACE(15);
    {
        constexpr int CG = 16 / C::NCG;          // gate columns per warp
        uint32_t gv[CG];
        if constexpr (CG == 4) tc::tmem_ld_x4(taddr + GATE_COL + cg * CG, gv);
        else tc::tmem_ld_x8(taddr + GATE_COL + cg * CG, gv);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < CG; ++i) {
            const int u = cg * CG + i;
            float a = __uint_as_float(gv[i]);
            if (C::STACK) a += __shfl_xor_sync(0xffffffffu, a, 16);     // hi rows (lanes 0-15) + lo rows (lanes 16-31)
            a += bg_s[u];
            if (g.sigmoid_gate) a = act_sigmoid<C::NS>(a);
            if (valid_e) m.gate[row_e * GATE_LD + u] = a;
        }
    }
    simt_bar<C>();
    if (vecw) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float2 gt = *reinterpret_cast<const float2*>(m.gate + L.row * GATE_LD + 8 * j + 2 * L.t);
#pragma unroll
            for (int c = 0; c < 3; ++c) { v.x[c][j][0] = vu[c][j][0] * gt.x; v.x[c][j][1] = vu[c][j][1] * gt.y; }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) v.x[c][2][0] = v.x[c][2][1] = 0.f;
    }
    TC_T(t6);
    WS_TRACE(16);
    WS_ACC(tb + 0, t0, t1); WS_ACC(tb + 1, t1, t2); WS_ACC(tb + 2, t2, t3); WS_ACC(tb + 3, t3, t4); WS_ACC(tb + 4, t4, t5);
    WS_ACC(tb + 5, t5, t6); WS_ACC(tb + 6, 0, 1);
}

// Segment table of a dst-sorted tile: seg[0..nseg) = first row of every run of equal dst, seg[nseg] = n;
// the count goes to seg[R + 1].  All SIMT threads call it; ends with a SIMT barrier.
template <class C>
__device__ __forceinline__ void build_segments(const Sm& m, int n) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int W = C::R / 32;
    bool start = false;
    unsigned bal = 0;
    if (tid < C::R) {
        start = tid < n && (tid == 0 || m.dst_s[tid] != m.dst_s[tid - 1]);
        bal = __ballot_sync(0xffffffffu, start);
        if (lane == 0) m.warp_cnt[warp] = __popc(bal);
    }
    simt_bar<C>();
    if (tid < C::R) {
        int base = 0;
        for (int w = 0; w < warp; ++w) base += m.warp_cnt[w];
        if (start) m.seg[base + __popc(bal & ((1u << lane) - 1u))] = tid;
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < W; ++w) tot += m.warp_cnt[w];
            m.seg[tot] = n;
            m.seg[C::R + 1] = tot;
        }
    }
    simt_bar<C>();
}

}  // namespace ws

#ifndef KPD_WS_CLUSTER
#define KPD_WS_CLUSTER 1
#endif
#ifndef KPD_WS_SPLIT_XWARPS
#define KPD_WS_SPLIT_XWARPS 8
#endif
using WsBf16 = ws::Cfg<128, 1, KPD_WS_CLUSTER>;    // bf16 operands, 128-row tiles (M = 128)
// bf16 (hi, lo) rows stacked into one M = 128 operand, 64-row tiles; 8 vector warps + 8 extra epilogue warps
using WsSplit = ws::Cfg<64, 2, KPD_WS_CLUSTER, KPD_WS_SPLIT_XWARPS>;
// bf16x3 edge kernel as CTA pairs (cta_group::2): two neighbouring 64-row tiles share every weight slab, each SM
// holding (and reading) half of it
using WsSplitPair = ws::Cfg<64, 2, 2, KPD_WS_SPLIT_XWARPS>;
// bf16x3 edge kernel, default: (hi, lo) PLANES of a 128-row tile, three MMAs per k-step into one accumulator
// (ws_common.cuh: Cfg::KS); all 16 SIMT warps own vector rows
using WsKS = ws::Cfg<128, 2, 1, 0, true>;
using WsBf16N = ws::Cfg<64, 1, KPD_WS_CLUSTER, KPD_WS_SPLIT_XWARPS>;    // bf16 operands, 64-row MMA tiles (M = 64): node / head kernels

// ------------------------------------------------------------------ edge kernel
// GVPMultiEdgeConv.message + aggregation (models/gvp.py:472-497, :540-550) for all edge types in one launch
template <class C>
__global__ void __launch_bounds__(C::NT, 1) gvp_edge_ws_kernel(const __grid_constant__ GvpEdgeLaunch L) {
    const GvpEtypeArgs& a = L.e[blockIdx.y];
    const int tile_begin = blockIdx.x * C::R;
    // this thread's edge, fetched together with the edge count (the arrays are sized at capacity, so the speculative
    // read is in bounds; rows past the end are mapped to the tile's last edge below)
    int my_s = 0, my_d = 0;
    if (threadIdx.x < C::R) {
        const int e = min(tile_begin + (int)threadIdx.x, a.cap - 1);
        my_s = __ldg(a.src + e); my_d = __ldg(a.dst + e);
    }
    const int E = a.rowptr[a.n_dst];
    if ((int)(blockIdx.x / C::CL) * C::CL * C::R >= E) return;      // the whole cluster is past the last edge
    const bool dead = tile_begin >= E;                              // (CTA pairs) this CTA only lends its half of the weights
    const int n = min(C::R, E - tile_begin);
    extern __shared__ __align__(128) unsigned char smem_ws[];
    TL_BEGIN(L.tl_slot);
    TC_T(e0);
    ws::Sm m = ws::carve<C>(smem_ws, L.kch);
    if (C::CL == 2) {
        m.rank = blockIdx.x & 1;                                    // (1-D clusters along x: rank in the pair)
        m.solo = (int)(blockIdx.x | 1) * C::R >= E;                 // the odd CTA's tile is empty
    }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Sd = L.Sdim, Vd = L.Vdim;
    constexpr bool ASYNC = C::CL == 1;        // pairs need both CTAs' barriers and TMEM before the first MMA: common set-up
    uint32_t tmem = 0;
    if (!ASYNC) tmem = ws::setup<C>(m, L.kch);
    if (warp >= C::NW) {
        if (ASYNC) { ws::control_setup<C>(m); tmem = *m.tmem_slot; }
        if (warp == C::NW) {
            if constexpr (C::KS) {
                ws::issue_ks<C>(a.msg, L.n_msg, m, tmem);       // the whole (converged) warp: see tc::elect_one()
            } else if (C::CL == 1) {
                ws::issue<C>(a.msg, L.n_msg, m, tmem);          // whole warp
            } else if (lane == 0) {
                if (C::CL == 2 && m.rank != 0) ws::relay<C>(a.msg, L.n_msg, m);
                else ws::issue<C>(a.msg, L.n_msg, m, tmem);
            } else if (lane == 1 && C::CL == 2 && m.rank != 0 && !dead) {
                ws::forward_ready<C>(L.n_msg, m);
            }
        } else {
            if (lane == 0) {
                if constexpr (C::KS) ws::produce_ks<C>(a.msg, L.n_msg, m);
                else ws::produce<C>(a.msg, L.n_msg, m, dead);
            }
        }
    } else if (!dead) {
        TC_T(e1);
        if (tid < C::R) { m.src_s[tid] = my_s; m.dst_s[tid] = my_d; }
        ws::simt_bar<C>();
        if (tid >= n && tid < C::R) { my_s = m.src_s[n - 1]; my_d = m.dst_s[n - 1]; }     // rows past the end mirror the last edge
        ws::simt_bar<C>();
        if (tid >= n && tid < C::R) { m.src_s[tid] = my_s; m.dst_s[tid] = my_d; }
        if (tid < C::R) {           // needed only by the segmented reduction at the end
            m.rp[2 * tid] = __ldg(a.rowptr + my_d);
            m.rp[2 * tid + 1] = __ldg(a.rowptr + my_d + 1);
        }
        ws::simt_bar<C>();
        TC_T(ga);
        // s_src: 16-byte cp.async straight into the canonical bf16 plane(s); consecutive lanes = consecutive rows
        {
            // 16-byte loads into registers + 16-byte shared-memory stores, a batch of GB rows-chunks (x hi, lo) in flight per
            // thread.  (cp.async straight into the planes was bound by the ISSUE rate of LDGSTS: ~17 cycles per warp
            // instruction, 8.8k of the 13k cycles a 128-row tile spent before its first MMA: measured.)  Consecutive lanes =
            // consecutive chunks of ONE source row: whole sectors per request; the shared-memory side stays conflict-free
            // through the +16-byte rotation of the k-chunk stride.
            const int cpr = Sd >> 3;                 // chunks per row
            const int items = C::R * cpr;
            constexpr int GB = 4;
            for (int base = 0; base < items; base += GB * C::NT_SIMT) {
                uint4 vh[GB];
                [[maybe_unused]] uint4 vl[GB];
                uint32_t off[GB];
#pragma unroll
                for (int k = 0; k < GB; ++k) {
                    const int idx = base + k * C::NT_SIMT + tid;
                    off[k] = 0xffffffffu;
                    if (idx < items) {
                        const int r = idx / cpr, kc = idx - r * cpr;
                        off[k] = (uint32_t)(kc * C::KCS) + ws::row_off<C>(r);
                        const size_t g = (size_t)m.src_s[r] * Sd + 8 * kc;
                        vh[k] = __ldg(reinterpret_cast<const uint4*>(a.s_hi + g));
                        if (C::NS == 2) vl[k] = __ldg(reinterpret_cast<const uint4*>(a.s_lo + g));
                    }
                }
#pragma unroll
                for (int k = 0; k < GB; ++k) {
                    if (off[k] != 0xffffffffu) {
                        *reinterpret_cast<uint4*>(m.A[0] + off[k]) = vh[k];
                        if (C::NS == 2) *reinterpret_cast<uint4*>(m.A[1] + off[k]) = vl[k];
                    }
                }
            }
        }
        TC_T(ga1);
        // geometry + v_src -> registers (gvp.py:474-480), in flight together with the gather
        const ws::Lane Ln = ws::lane_geometry<C>();
        const bool vecw = warp < C::NWV;              // this warp owns 8 tile rows of vectors
        ws::VF v;
        float dx = 0.f, dy = 0.f, dz = 0.f;
        if (vecw) {
            const int sI = m.src_s[Ln.row], dI = m.dst_s[Ln.row];
            dx = a.xs[3 * sI] - a.xd[3 * dI]; dy = a.xs[3 * sI + 1] - a.xd[3 * dI + 1]; dz = a.xs[3 * sI + 2] - a.xd[3 * dI + 2];
            ws::vf_load(v, a.v_src + (size_t)sI * (Vd * 3), Vd, Ln.t);
        }
        TC_T(ga2);
        if (ASYNC) {        // only the k-chunks behind the gathered scalars need zeros (rbf / |Vh| columns and K padding)
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            const int c0 = Sd >> 3, per = C::KCS / 16;
            for (int i = tid; i < (L.kch - c0) * per; i += C::NT_SIMT) {
                reinterpret_cast<uint4*>(m.A[0] + (size_t)c0 * C::KCS)[i] = z;
                if (C::KS) reinterpret_cast<uint4*>(m.A[1] + (size_t)c0 * C::KCS)[i] = z;
            }
        }
        TC_T(ga3);
        ws::build_segments<C>(m, n);             // (its barriers also order the zero fill before the rbf stores)
        TC_T(gb);
        WS_ACC(40, ga, ga1); WS_ACC(41, ga1, ga2); WS_ACC(42, ga2, ga3); WS_ACC(43, ga3, gb);
        if (vecw) {
            const float dij = sqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1e-8f)) + 1e-8f;
            // the unit x_diff is the LAST input channel here (channel Vd; Wh is staged with its rows permuted to match)
            const int sx = Vd >> 3, tx = (Vd & 7) >> 1, ex = Vd & 1;
            if (Ln.t == tx) {
#pragma unroll
                for (int ss = 0; ss < 3; ++ss)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        if (ss == sx && e == ex) { v.x[0][ss][e] = dx / dij; v.x[1][ss][e] = dy / dij; v.x[2][ss][e] = dz / dij; }
            }
            for (int k = Ln.t; k < L.rbf_dim; k += 4) {
                const float z = (dij - (float)k * L.rbf_step) / L.rbf_sigma;
                ws::put_scalar<C>(m, Ln.row, Sd + k, expf(-(z * z)));
            }
        }
        if (ASYNC) tmem = ws::simt_join<C>(m);
        TC_T(gc);
        ws::publish_mma<C>(m, m.feats_ready);
        TC_T(e2);
        WS_ACC(12, e1, ga); WS_ACC(14, ga, gb); WS_ACC(15, gb, gc);
        for (int i = 0; i < L.n_msg; ++i) ws::gvp_simt<C>(a.msg[i], i, m, tmem, v, Ln, C::R, 0, i == L.n_msg - 1);
        TC_T(e3);
        // ---- deterministic segmented reduction by destination (all MMAs and bulk copies have completed)
        float* VS = reinterpret_cast<float*>(m.ring);
        if (vecw) ws::vf_store(v, VS + Ln.row * ws::VS_LD, Vd, Ln.t);
        ws::simt_bar<C>();
        {
            const int nseg = m.seg[C::R + 1];
            float* part0 = a.part + ((size_t)blockIdx.x * 2 + 0) * L.pw;
            float* part1 = a.part + ((size_t)blockIdx.x * 2 + 1) * L.pw;
            // scalars: one warp per group of segments, one lane per 8-column chunk: 16-byte reads of the plane(s)
            // (bank-conflict free thanks to the +16 B chunk stride), 32-byte coalesced stores
            if (lane < (Sd >> 3)) {
                const unsigned char* p0 = m.A[0] + (size_t)lane * C::KCS;
                [[maybe_unused]] const unsigned char* p1 = m.A[1] + (size_t)lane * C::KCS;
                for (int sg = warp; sg < nseg; sg += C::NW) {
                    const int ra = m.seg[sg], rb = m.seg[sg + 1];
                    float acc[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
                    for (int j = ra; j < rb; ++j) {
                        const uint32_t ro = ws::row_off<C>(j);
                        const uint4 wh = *reinterpret_cast<const uint4*>(p0 + ro);
                        float x[8] = {__uint_as_float(wh.x << 16), __uint_as_float(wh.x & 0xffff0000u), __uint_as_float(wh.y << 16),
                                      __uint_as_float(wh.y & 0xffff0000u), __uint_as_float(wh.z << 16), __uint_as_float(wh.z & 0xffff0000u),
                                      __uint_as_float(wh.w << 16), __uint_as_float(wh.w & 0xffff0000u)};
                        if (C::NS == 2) {
                            const uint4 wl = *reinterpret_cast<const uint4*>(p1 + ro);
                            x[0] += __uint_as_float(wl.x << 16); x[1] += __uint_as_float(wl.x & 0xffff0000u);
                            x[2] += __uint_as_float(wl.y << 16); x[3] += __uint_as_float(wl.y & 0xffff0000u);
                            x[4] += __uint_as_float(wl.z << 16); x[5] += __uint_as_float(wl.z & 0xffff0000u);
                            x[6] += __uint_as_float(wl.w << 16); x[7] += __uint_as_float(wl.w & 0xffff0000u);
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q) acc[q] += x[q];
                    }
                    const bool from_prev = (ra == 0) && (m.rp[2 * ra] < tile_begin);
                    const bool into_next = (rb == n) && (m.rp[2 * ra + 1] > tile_begin + n);
                    float* t = (from_prev ? part0 : into_next ? part1 : a.sm + (size_t)m.dst_s[ra] * Sd) + 8 * lane;
                    *reinterpret_cast<float4*>(t) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                    *reinterpret_cast<float4*>(t + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
                }
            }
            // vectors: one thread per (segment, vector column)
            const int nvc = Vd * 3;
            for (int i = tid; i < nseg * nvc; i += C::NT_SIMT) {
                const int sg = i / nvc, col = i - sg * nvc;
                const int ra = m.seg[sg], rb = m.seg[sg + 1];
                float s0 = 0.f;
                for (int j = ra; j < rb; ++j) s0 += VS[j * ws::VS_LD + col];
                const bool from_prev = (ra == 0) && (m.rp[2 * ra] < tile_begin);
                const bool into_next = (rb == n) && (m.rp[2 * ra + 1] > tile_begin + n);
                float* t = from_prev ? part0 + Sd + col : into_next ? part1 + Sd + col : a.vm + (size_t)m.dst_s[ra] * nvc + col;
                t[0] = s0;
            }
        }
        TC_T(e4);
        WS_ACC(8, e0, e1); WS_ACC(9, e1, e2); WS_ACC(10, e2, e3); WS_ACC(11, e3, e4); WS_ACC(13, 0, 1);
    }
    if (ASYNC && warp < C::NW && dead) tmem = ws::simt_join<C>(m);     // (unreachable without clusters; keeps barrier 3 balanced)
    ws::teardown<C>(tmem);
    TL_END(L.tl_slot);
}

// ------------------------------------------------------------------ node kernel
namespace ws {

// 8 consecutive columns [c0, c0+8) of the aggregated message of node d for one edge type (see seg_gather)
__device__ __forceinline__ void seg_gather8(const float* __restrict__ out, int ld_out, const float* __restrict__ part, int pw,
                                            int r0, int r1, int d, int c0, int tile, float (&acc)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if (r1 <= r0) return;
    const int t0 = r0 / tile, t1 = (r1 - 1) / tile;
    auto add = [&](const float* p) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w; acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
    };
    if (t0 == t1) { add(out + (size_t)d * ld_out + c0); return; }
    add(part + ((size_t)t0 * 2 + 1) * pw + c0);
    for (int t = t0 + 1; t <= t1; ++t) add(part + ((size_t)t * 2 + 0) * pw + c0);
}

// LayerNorm over a row held as 8 consecutive features per lane (features 8*lane .. 8*lane+7; lanes beyond Sdim idle);
// w8 / b8: this lane's slice of the affine parameters, loaded once per warp (ln_params8)
__device__ __forceinline__ void ln_params8(const float* __restrict__ w, const float* __restrict__ b, bool act, int lane,
                                           float (&w8)[8], float (&b8)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { w8[i] = 0.f; b8[i] = 0.f; }
    if (act) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + 8 * lane)), w1 = __ldg(reinterpret_cast<const float4*>(w + 8 * lane + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + 8 * lane)), b1 = __ldg(reinterpret_cast<const float4*>(b + 8 * lane + 4));
        w8[0] = w0.x; w8[1] = w0.y; w8[2] = w0.z; w8[3] = w0.w; w8[4] = w1.x; w8[5] = w1.y; w8[6] = w1.z; w8[7] = w1.w;
        b8[0] = b0.x; b8[1] = b0.y; b8[2] = b0.z; b8[3] = b0.w; b8[4] = b1.x; b8[5] = b1.y; b8[6] = b1.z; b8[7] = b1.w;
    }
}
__device__ __forceinline__ void warp_layernorm8(float (&x)[8], bool act, int Sdim, const float (&w8)[8], const float (&b8)[8]) {
    float s = 0.f;
    if (act) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s += x[i];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)Sdim;
    float v = 0.f;
    if (act) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = x[i] - mean; v += d * d; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = 1.0f / sqrtf(v / (float)Sdim + 1e-5f);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = (x[i] - mean) * rstd * w8[i] + b8[i];
}

// vector part of GVPLayerNorm (gvp.py:163-165) on the register-resident vectors of one row:
// vn = sqrt(mean_u clamp(|v_u|^2, 1e-8) + eps) + eps
__device__ __forceinline__ float vec_norm(const VF& v, int nv, const Lane& L) {
    float q = 0.f;
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const float n2 = v.x[0][s][e] * v.x[0][s][e] + v.x[1][s][e] * v.x[1][s][e] + v.x[2][s][e] * v.x[2][s][e];
            if (8 * s + 2 * L.t + e < nv) q += fmaxf(n2, 1e-8f);
        }
    q += __shfl_xor_sync(0xffffffffu, q, 1);      // the four lanes of a row hold disjoint channels
    q += __shfl_xor_sync(0xffffffffu, q, 2);
    return sqrtf(q / (float)nv + 1e-5f) + 1e-5f;
}

// 8 consecutive feats of one row read back from the bf16 plane(s)
template <class C>
__device__ __forceinline__ void get_chunk(const Sm& m, int row, int kc, float (&f)[8]) {
    const uint32_t off = (uint32_t)(kc * C::KCS) + row_off<C>(row);
    const uint4 hi = *reinterpret_cast<const uint4*>(m.A[0] + off);
    f[0] = __uint_as_float(hi.x << 16); f[1] = __uint_as_float(hi.x & 0xffff0000u);
    f[2] = __uint_as_float(hi.y << 16); f[3] = __uint_as_float(hi.y & 0xffff0000u);
    f[4] = __uint_as_float(hi.z << 16); f[5] = __uint_as_float(hi.z & 0xffff0000u);
    f[6] = __uint_as_float(hi.w << 16); f[7] = __uint_as_float(hi.w & 0xffff0000u);
    if (C::NS == 2) {
        const uint4 lo = *reinterpret_cast<const uint4*>(m.A[1] + off);
        f[0] += __uint_as_float(lo.x << 16); f[1] += __uint_as_float(lo.x & 0xffff0000u);
        f[2] += __uint_as_float(lo.y << 16); f[3] += __uint_as_float(lo.y & 0xffff0000u);
        f[4] += __uint_as_float(lo.z << 16); f[5] += __uint_as_float(lo.z & 0xffff0000u);
        f[6] += __uint_as_float(lo.w << 16); f[7] += __uint_as_float(lo.w & 0xffff0000u);
    }
}

// 8 fp32 values -> bf16 hi (and lo) planes in global memory (row-major, 16-byte stores)
__device__ __forceinline__ void store_planes8(__nv_bfloat16* hp, __nv_bfloat16* lp, size_t i, const float (&f)[8]) {
    uint4 hi, lo;
    hi.x = tc::pack_bf16x2(f[0], f[1]); hi.y = tc::pack_bf16x2(f[2], f[3]);
    hi.z = tc::pack_bf16x2(f[4], f[5]); hi.w = tc::pack_bf16x2(f[6], f[7]);
    lo.x = tc::pack_bf16x2(f[0] - __uint_as_float(hi.x << 16), f[1] - __uint_as_float(hi.x & 0xffff0000u));
    lo.y = tc::pack_bf16x2(f[2] - __uint_as_float(hi.y << 16), f[3] - __uint_as_float(hi.y & 0xffff0000u));
    lo.z = tc::pack_bf16x2(f[4] - __uint_as_float(hi.z << 16), f[5] - __uint_as_float(hi.z & 0xffff0000u));
    lo.w = tc::pack_bf16x2(f[6] - __uint_as_float(hi.w << 16), f[7] - __uint_as_float(hi.w & 0xffff0000u));
    *reinterpret_cast<uint4*>(hp + i) = hi;
    *reinterpret_cast<uint4*>(lp + i) = lo;
}

}  // namespace ws

// ------------------------------------------------------------------ encoders (tensor-core modes)
// Both node encoders of LigRecDynamicsGVP.forward in ONE launch (dynamics_gvp.py:161-184): time concatenated first,
// Linear + SiLU + LayerNorm -> s (fp32 + bf16 planes); ligand vectors start at zero, keypoint vectors are copied.
// One warp per ENC_ROWS rows (weights re-used across them), a lane per 8 output features.
constexpr int ENC_ROWS = 4;
struct GvpEncArgs {
    int n[2], K[2];                       // rows and input width (without the time channel) per node type
    const float* in[2];                   // [n][K]
    const float* WT[2]; const float* bias[2]; const float* lnw[2]; const float* lnb[2];   // WT: [K + 1][Sp] K-major
    const int* batch[2];
    float* s[2]; __nv_bfloat16* s_hi[2]; __nv_bfloat16* s_lo[2];
    float* v[2]; const float* v_kp;
    const float* t_ptr; int t_per_complex, S, Sp, V;
};
__global__ void __launch_bounds__(256) gvp_encode_kernel(const __grid_constant__ GvpEncArgs a) {
    const int nt = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = (blockIdx.x * 8 + warp) * ENC_ROWS;
    const int n = a.n[nt], K = a.K[nt], S = a.S;
    if (r0 >= n) return;
    const bool act = 8 * lane < S;
    const int col = act ? 8 * lane : 0;
    float acc[ENC_ROWS][8];
    {
        const float4 b0 = *reinterpret_cast<const float4*>(a.bias[nt] + col), b1 = *reinterpret_cast<const float4*>(a.bias[nt] + col + 4);
#pragma unroll
        for (int j = 0; j < ENC_ROWS; ++j) {
            acc[j][0] = b0.x; acc[j][1] = b0.y; acc[j][2] = b0.z; acc[j][3] = b0.w;
            acc[j][4] = b1.x; acc[j][5] = b1.y; acc[j][6] = b1.z; acc[j][7] = b1.w;
        }
    }
    const float* WT = a.WT[nt];
    for (int k0 = 0; k0 <= K; k0 += 32) {
        // lane l holds input feature k0 + l of each row (the time channel is feature K)
        float xin[ENC_ROWS];
#pragma unroll
        for (int j = 0; j < ENC_ROWS; ++j) {
            const int r = min(r0 + j, n - 1), k = k0 + lane;
            xin[j] = k < K ? a.in[nt][(size_t)r * K + k] : (k == K ? a.t_ptr[a.t_per_complex ? a.batch[nt][r] : 0] : 0.f);
        }
        const int kend = min(32, K + 1 - k0);
#pragma unroll 8
        for (int kk = 0; kk < kend; ++kk) {
            const float4 w0 = *reinterpret_cast<const float4*>(WT + (size_t)(k0 + kk) * a.Sp + col);
            const float4 w1 = *reinterpret_cast<const float4*>(WT + (size_t)(k0 + kk) * a.Sp + col + 4);
#pragma unroll
            for (int j = 0; j < ENC_ROWS; ++j) {
                const float x = __shfl_sync(0xffffffffu, xin[j], kk);
                acc[j][0] = fmaf(x, w0.x, acc[j][0]); acc[j][1] = fmaf(x, w0.y, acc[j][1]); acc[j][2] = fmaf(x, w0.z, acc[j][2]);
                acc[j][3] = fmaf(x, w0.w, acc[j][3]); acc[j][4] = fmaf(x, w1.x, acc[j][4]); acc[j][5] = fmaf(x, w1.y, acc[j][5]);
                acc[j][6] = fmaf(x, w1.z, acc[j][6]); acc[j][7] = fmaf(x, w1.w, acc[j][7]);
            }
        }
    }
    float lw[8], lb[8];
    ws::ln_params8(a.lnw[nt], a.lnb[nt], act, lane, lw, lb);
#pragma unroll
    for (int j = 0; j < ENC_ROWS; ++j) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = act ? silu_f(acc[j][i]) : 0.f;
        ws::warp_layernorm8(acc[j], act, S, lw, lb);
        const int r = r0 + j;
        if (r < n && act) {
            const size_t o = (size_t)r * S + col;
            *reinterpret_cast<float4*>(a.s[nt] + o) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
            *reinterpret_cast<float4*>(a.s[nt] + o + 4) = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
            ws::store_planes8(a.s_hi[nt], a.s_lo[nt], o, acc[j]);
        }
        if (r < n) {
            const int nv = 3 * a.V;
            for (int c = lane; c < nv; c += 32) a.v[nt][(size_t)r * nv + c] = nt == 0 ? 0.f : a.v_kp[(size_t)r * nv + c];
        }
    }
}

// node / head tiles hold NR valid rows of the C::R-row MMA tile; rows are dealt round-robin to the SIMT warps for the
// scalar phases.  NR = 32: more, lighter CTAs -- the shorter critical path wins while the GPU is not full (gvp_ca, 16
// ligands: 33.1 against 29.9 ligands/s).  NR = 64 (full tiles): half the CTAs for ~1.4x the cycles each -- less SM time,
// which wins once concurrent sub-batches keep every SM busy (headline 89.3 -> 91.9, gvp_ca 256 ligands 67.9 -> 71.0).
// kpd_gvp_forward picks per call by the number of edge tiles of the call (node_tile_rows()).
constexpr int NODE_ROWS = 32, NODE_ROWS_FULL = 64;

// GVPMultiEdgeConv.forward after the message pass (models/gvp.py:501-536): messages / norm, residual,
// GVPLayerNorm, update GVPs, residual, GVPLayerNorm; both node types in one launch.
template <class C, int NODE_ROWS>
__global__ void __launch_bounds__(C::NT, 1) gvp_node_ws_kernel(const __grid_constant__ GvpNodeLaunch L) {
    const GvpNodeArgs& a = L.nt[blockIdx.y];
    const int n0 = blockIdx.x * NODE_ROWS;
    if ((int)(blockIdx.x / C::CL) * C::CL * NODE_ROWS >= a.n) return;
    const bool dead = n0 >= a.n;
    const int n = min(NODE_ROWS, a.n - n0);
    constexpr int RPW = NODE_ROWS / C::NW;     // rows per warp in the scalar phases (row = warp + NW * j)
    extern __shared__ __align__(128) unsigned char smem_ws[];
    TL_BEGIN(L.tl_slot + (int)blockIdx.y);
    TC_T(n0t);
    ws::Sm m = ws::carve<C>(smem_ws, a.kch);
    const uint32_t tmem = ws::setup<C>(m, a.kch);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Sd = a.Sdim, Vd = a.Vdim;
    if (warp == C::NW) {
        if (C::CL == 1 || lane == 0) ws::issue<C>(a.upd, a.n_upd, m, tmem);      // (CL == 1: the whole warp, see issue())
    } else if (warp == C::NW + 1) {
        if (lane == 0) ws::produce<C>(a.upd, a.n_upd, m, dead);
    } else if (!dead) {
        const ws::Lane Ln = ws::lane_geometry<C>();
        const bool act = 8 * lane < Sd;
        // ---- phase 0: per-row metadata -> shared memory (one round trip for the whole tile)
        int* meta = m.src_s;                                   // [R][4]: r0, r1 of both edge types
        float* nvs = reinterpret_cast<float*>(m.src_s + 4 * C::R);   // [R]: the message normaliser
        if (tid < C::R) {
            const int nd = n0 + min(tid, n - 1);
            for (int e = 0; e < 2; ++e) {
                meta[4 * tid + 2 * e] = e < a.n_et ? a.rowptr[e][nd] : 0;
                meta[4 * tid + 2 * e + 1] = e < a.n_et ? a.rowptr[e][nd + 1] : 0;
            }
            float nv = a.norm_const;
            if (a.norm_mode == 1) nv = 1.0f;
            else if (a.norm_mode == 2) {
                const int b = a.node_batch[nd];
                const int p0 = a.ptr[b], p1 = a.ptr[b + 1];
                int tot = 0;
                for (int e = 0; e < a.n_et; ++e) tot += a.rowptr[e][p1] - a.rowptr[e][p0];
                nv = (float)tot / (float)(p1 - p0) + 1.0f;
            }
            nvs[tid] = nv;
        }
        ws::simt_bar<C>();
        TC_T(n1t);
        // ---- phase 1a: scalars, RPW rows per warp with every load of the batch in flight together:
        //      features + messages / norm, LayerNorm -> residual (global) + A
        {
            const int lg = a.edge_tile == 128 ? 7 : 6;
            float lw[8], lb[8];
            ws::ln_params8(a.mln_w, a.mln_b, act, lane, lw, lb);
            // (batches of RB = 2 rows: their 12 16-byte loads per lane are in flight together; all four rows of a 64-row tile
            // at once would need 128 registers for the loads alone -- 300 B of spills per thread, each an L2 round trip)
            constexpr int RB = RPW < 2 ? RPW : 2;
#pragma unroll 1
            for (int jb = 0; jb < RPW; jb += RB) {
            float x[RB][8];
            float4 ga[RB][2][2], sa[RB][2];
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const int r = warp + C::NW * (jb + j);
                const int nd = n0 + min(r, n - 1);
                const int col = act ? 8 * lane : 0;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int r0 = meta[4 * r + 2 * e], r1 = meta[4 * r + 2 * e + 1];
                    const int t0 = r0 >> lg, t1 = (max(r1, 1) - 1) >> lg;
                    // whole CSR row inside one tile: the direct sum; else the first tile's "continues" partial
                    const float* p = (r1 <= r0 || t0 == t1) ? a.sm[e] + (size_t)nd * Sd : a.part[e] + ((size_t)t0 * 2 + 1) * a.pw;
                    ga[j][e][0] = __ldg(reinterpret_cast<const float4*>(p + col));
                    ga[j][e][1] = __ldg(reinterpret_cast<const float4*>(p + col + 4));
                }
                sa[j][0] = __ldg(reinterpret_cast<const float4*>(a.s + (size_t)nd * Sd + col));
                sa[j][1] = __ldg(reinterpret_cast<const float4*>(a.s + (size_t)nd * Sd + col + 4));
            }
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const int r = warp + C::NW * (jb + j);
                float msg[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) msg[i] = 0.f;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int r0 = meta[4 * r + 2 * e], r1 = meta[4 * r + 2 * e + 1];
                    if (r1 > r0) {
                        float g8[8] = {ga[j][e][0].x, ga[j][e][0].y, ga[j][e][0].z, ga[j][e][0].w,
                                       ga[j][e][1].x, ga[j][e][1].y, ga[j][e][1].z, ga[j][e][1].w};
                        const int t0 = r0 >> lg, t1 = (r1 - 1) >> lg;
                        for (int t = t0 + 1; t <= t1; ++t) {      // rare: the row continues into further tiles
                            const float* q = a.part[e] + ((size_t)t * 2 + 0) * a.pw + (act ? 8 * lane : 0);
                            const float4 q0 = *reinterpret_cast<const float4*>(q), q1 = *reinterpret_cast<const float4*>(q + 4);
                            g8[0] += q0.x; g8[1] += q0.y; g8[2] += q0.z; g8[3] += q0.w;
                            g8[4] += q1.x; g8[5] += q1.y; g8[6] += q1.z; g8[7] += q1.w;
                        }
                        const float icnt = a.norm_mode == 1 ? 1.0f / (float)(r1 - r0) : 1.0f;      // fn.mean per edge type
#pragma unroll
                        for (int i = 0; i < 8; ++i) msg[i] = fmaf(g8[i], icnt, msg[i]);
                    }
                }
                const float inv = 1.0f / nvs[r];          // (the tensor-core modes multiply by reciprocals)
                x[j][0] = fmaf(msg[0], inv, sa[j][0].x); x[j][1] = fmaf(msg[1], inv, sa[j][0].y);
                x[j][2] = fmaf(msg[2], inv, sa[j][0].z); x[j][3] = fmaf(msg[3], inv, sa[j][0].w);
                x[j][4] = fmaf(msg[4], inv, sa[j][1].x); x[j][5] = fmaf(msg[5], inv, sa[j][1].y);
                x[j][6] = fmaf(msg[6], inv, sa[j][1].z); x[j][7] = fmaf(msg[7], inv, sa[j][1].w);
                ws::warp_layernorm8(x[j], act, Sd, lw, lb);
                if (act) {
                    if (r < n) {
                        float* sp = a.s_out + (size_t)(n0 + r) * Sd + 8 * lane;
                        *reinterpret_cast<float4*>(sp) = make_float4(x[j][0], x[j][1], x[j][2], x[j][3]);
                        *reinterpret_cast<float4*>(sp + 4) = make_float4(x[j][4], x[j][5], x[j][6], x[j][7]);
                    }
                    ws::put_chunk<C>(m, r, lane, x[j]);
                }
            }
            }
        }
        // ---- phase 1b: vectors of lane (row, c) -> registers; residual stash in global
        TC_T(n2t);
        ws::VF v;
        const bool vecw = warp < C::NWV;              // this warp owns 8 tile rows of vectors
        if (vecw) {
            const int nd = n0 + min(Ln.row, n - 1);
            const float nv = nvs[Ln.row];
            ws::VF msg;
            ws::vf_zero(msg);
            for (int e = 0; e < a.n_et; ++e) {
                const int r0 = meta[4 * Ln.row + 2 * e], r1 = meta[4 * Ln.row + 2 * e + 1];
                const float icnt = a.norm_mode == 1 ? 1.0f / (float)max(r1 - r0, 1) : 1.0f;
                if (r1 > r0) {
                    const int t0 = r0 / a.edge_tile, t1 = (r1 - 1) / a.edge_tile;
                    ws::VF gsum;
                    if (t0 == t1) {            // the common case: one tile holds the whole CSR row
                        ws::vf_load(gsum, a.vm[e] + (size_t)nd * (3 * Vd), Vd, Ln.t);
                    } else {
                        ws::vf_load(gsum, a.part[e] + ((size_t)t0 * 2 + 1) * a.pw + Sd, Vd, Ln.t);
                        for (int t = t0 + 1; t <= t1; ++t) {
                            ws::VF q;
                            ws::vf_load(q, a.part[e] + ((size_t)t * 2 + 0) * a.pw + Sd, Vd, Ln.t);
#pragma unroll
                            for (int c = 0; c < 3; ++c)
#pragma unroll
                                for (int ss = 0; ss < 2; ++ss) { gsum.x[c][ss][0] += q.x[c][ss][0]; gsum.x[c][ss][1] += q.x[c][ss][1]; }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 3; ++c)
#pragma unroll
                        for (int ss = 0; ss < 2; ++ss)
#pragma unroll
                            for (int ee = 0; ee < 2; ++ee)
                                msg.x[c][ss][ee] = fmaf(gsum.x[c][ss][ee], icnt, msg.x[c][ss][ee]);
                }
            }
            ws::vf_load(v, a.v + (size_t)nd * (3 * Vd), Vd, Ln.t);
            const float inv_nv = 1.0f / nv;
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int ss = 0; ss < 2; ++ss)
#pragma unroll
                    for (int ee = 0; ee < 2; ++ee) v.x[c][ss][ee] = fmaf(msg.x[c][ss][ee], inv_nv, v.x[c][ss][ee]);
            const float ivn = 1.0f / ws::vec_norm(v, Vd, Ln);
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int ss = 0; ss < 2; ++ss)
#pragma unroll
                    for (int ee = 0; ee < 2; ++ee) v.x[c][ss][ee] *= ivn;
            if (Ln.row < n) ws::vf_store(v, a.v_out + (size_t)nd * (3 * Vd), Vd, Ln.t);
        }
        ws::publish_mma<C>(m, m.feats_ready);
        TC_T(n3t);
        // ---- phase 2: update GVPs on the tensor cores
        for (int i = 0; i < a.n_upd; ++i) ws::gvp_simt<C>(a.upd[i], i, m, tmem, v, Ln, n, 16);
        // ---- phase 3: residual + GVPLayerNorm -> global (fp32 + the bf16 planes the next edge kernel gathers)
        TC_T(n4t);
        {
            float lw[8], lb[8];
            ws::ln_params8(a.uln_w, a.uln_b, act, lane, lw, lb);
            constexpr int RB = RPW < 2 ? RPW : 2;           // (row pairs, as in phase 1a: the vectors stay live underneath)
#pragma unroll 1
            for (int jb = 0; jb < RPW; jb += RB) {
            float x[RB][8];
            float4 sa[RB][2];
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const int r = warp + C::NW * (jb + j);
                const int nd = n0 + min(r, n - 1);
                const int col = act ? 8 * lane : 0;
                sa[j][0] = *reinterpret_cast<const float4*>(a.s_out + (size_t)nd * Sd + col);
                sa[j][1] = *reinterpret_cast<const float4*>(a.s_out + (size_t)nd * Sd + col + 4);
            }
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const int r = warp + C::NW * (jb + j);
#pragma unroll
                for (int i = 0; i < 8; ++i) x[j][i] = 0.f;
                if (act) {
                    ws::get_chunk<C>(m, r, lane, x[j]);
                    x[j][0] += sa[j][0].x; x[j][1] += sa[j][0].y; x[j][2] += sa[j][0].z; x[j][3] += sa[j][0].w;
                    x[j][4] += sa[j][1].x; x[j][5] += sa[j][1].y; x[j][6] += sa[j][1].z; x[j][7] += sa[j][1].w;
                }
                ws::warp_layernorm8(x[j], act, Sd, lw, lb);
                if (act && r < n) {
                    const size_t o = (size_t)(n0 + r) * Sd + 8 * lane;
                    *reinterpret_cast<float4*>(a.s_out + o) = make_float4(x[j][0], x[j][1], x[j][2], x[j][3]);
                    *reinterpret_cast<float4*>(a.s_out + o + 4) = make_float4(x[j][4], x[j][5], x[j][6], x[j][7]);
                    ws::store_planes8(a.s_hi_out, a.s_lo_out, o, x[j]);
                }
            }
            }
        }
        if (vecw) {
            const int nd = n0 + min(Ln.row, n - 1);
            ws::VF res;
            ws::vf_load(res, a.v_out + (size_t)nd * (3 * Vd), Vd, Ln.t);
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int ss = 0; ss < 2; ++ss)
#pragma unroll
                    for (int ee = 0; ee < 2; ++ee) v.x[c][ss][ee] += res.x[c][ss][ee];
            const float ivn = 1.0f / ws::vec_norm(v, Vd, Ln);
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int ss = 0; ss < 2; ++ss)
#pragma unroll
                    for (int ee = 0; ee < 2; ++ee) v.x[c][ss][ee] *= ivn;
            if (Ln.row < n) ws::vf_store(v, a.v_out + (size_t)nd * (3 * Vd), Vd, Ln.t);
        }
        TC_T(n5t);
        WS_ACC(24, n0t, n1t); WS_ACC(25, n1t, n2t); WS_ACC(26, n2t, n3t); WS_ACC(27, n3t, n4t); WS_ACC(28, n4t, n5t); WS_ACC(29, 0, 1);
    }
    ws::teardown<C>(tmem);
    TL_END(L.tl_slot + (int)blockIdx.y);
}

// NoisePredictionBlock (models/dynamics_gvp.py:38-44): noise GVPs + Linear(64 -> atom_nf); eps_x = vectors.squeeze(1)
template <class C, int NODE_ROWS>
__global__ void __launch_bounds__(C::NT, 1) gvp_head_ws_kernel(const __grid_constant__ GvpHeadArgs a) {
    const int n0 = blockIdx.x * NODE_ROWS;
    if ((int)(blockIdx.x / C::CL) * C::CL * NODE_ROWS >= a.n) return;
    const bool dead = n0 >= a.n;
    const int n = min(NODE_ROWS, a.n - n0);
    extern __shared__ __align__(128) unsigned char smem_ws[];
    ws::Sm m = ws::carve<C>(smem_ws, a.kch);
    const uint32_t tmem = ws::setup<C>(m, a.kch);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Sd = a.Sdim, Vd = a.Vdim;
    if (warp == C::NW) {
        if (C::CL == 1 || lane == 0) ws::issue<C>(a.g, a.n_gvps, m, tmem);
    } else if (warp == C::NW + 1) {
        if (lane == 0) ws::produce<C>(a.g, a.n_gvps, m, dead);
    } else if (!dead) {
        const ws::Lane Ln = ws::lane_geometry<C>();
        {
            const int cpr = Sd >> 3;                 // (lanes along a row: whole sectors per request, see the edge kernel)
            const int items = NODE_ROWS * cpr;
            for (int idx = tid; idx < items; idx += C::NT_SIMT) {
                const int r = idx / cpr, kc = idx - r * cpr;
                const uint32_t off = (uint32_t)(kc * C::KCS) + ws::row_off<C>(r);
                const size_t g = (size_t)(n0 + min(r, n - 1)) * Sd + 8 * kc;
                cp_async16(m.A[0] + off, a.s_hi + g);
                if (C::NS == 2) cp_async16(m.A[1] + off, a.s_lo + g);
            }
            cp_async_commit();
        }
        ws::VF v;
        const bool vecw = warp < C::NWV;
        if (vecw) ws::vf_load(v, a.v + (size_t)(n0 + min(Ln.row, n - 1)) * (3 * Vd), Vd, Ln.t);
        cp_async_wait<0>();
        ws::publish_mma<C>(m, m.feats_ready);
        for (int i = 0; i < a.n_gvps; ++i) ws::gvp_simt<C>(a.g[i], i, m, tmem, v, Ln, n, 32);
        // to_scalar_output + vectors.squeeze(1)  (dynamics_gvp.py:42-43)
        for (int idx = tid; idx < n * a.F; idx += C::NT_SIMT) {
            const int r = idx / a.F, c = idx - r * a.F;
            float s = a.bo[c];
            for (int k = 0; k < a.hid_out; ++k) s = fmaf(ws::get_scalar<C>(m, r, k), a.WoT[k * a.Fp + c], s);
            a.eps_h[(size_t)(n0 + r) * a.F + c] = s;
        }
        if (vecw && Ln.t == 0 && Ln.row < n) {
            float* ex = a.eps_x + (size_t)(n0 + Ln.row) * 3;
            ex[0] = v.x[0][0][0]; ex[1] = v.x[1][0][0]; ex[2] = v.x[2][0][0];
        }
    }
    ws::teardown<C>(tmem);
}
