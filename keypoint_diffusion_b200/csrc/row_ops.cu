// Node-level dense rows: Y = act(X @ WT + b) (+R), LayerNorm, and a few row utilities.
// Replaces the nn.Linear / SiLU / LayerNorm calls on node tensors in the reference
// (models/dynamics.py:355-356, :202-204, :380; models/dynamics_gvp.py:168-169).
//
// fp32 SIMT tiles: 64x64 output tile per CTA, 16-wide K slices in shared memory, 4x4
// micro-tile per thread.  These are the small-M GEMMs of the path (M = nodes in the batch);
// the per-edge contractions live in the fused tile kernels (egnn.cu / gvp.cu).
#include "common.cuh"
#include <stdarg.h>
#include <vector>

namespace kpd {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static long long g_launches = 0;
long long launch_count() { return g_launches; }

int check_launch(const char* what) {
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return -2;
    }
    return 0;
}

struct ProfRec { cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof;
static int g_prof_id = 0;
static size_t g_prof_used = 0;
static bool g_prof_open = false;

static bool prof_active(int id, cudaStream_t st) {
    if (g_prof_id != id || g_prof_used >= g_prof.size()) return false;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return false;
    return true;
}

bool prof_enabled() { return g_prof_id > 0; }

void prof_begin(int id, cudaStream_t st) {
    if (!prof_active(id, st)) return;
    cudaEventRecord(g_prof[g_prof_used].a, st);
    g_prof_open = true;
}

void prof_end(int id, cudaStream_t st) {
    if (!g_prof_open || g_prof_id != id) return;
    cudaEventRecord(g_prof[g_prof_used].b, st);
    ++g_prof_used;
    g_prof_open = false;
}

constexpr int LBM = 64, LBN = 64, LBK = 16;

__global__ void __launch_bounds__(256)
linear_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ WT, int ldw,
              const float* __restrict__ bias, const float* __restrict__ R, int ldr,
              float* __restrict__ Y, int ldy, int M, int K, int N, int act) {
    __shared__ float As[LBK][LBM + 4];
    __shared__ __align__(16) float Bs[LBK][LBN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * LBM, n0 = blockIdx.x * LBN;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += LBK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // X tile: 64 rows x 16 k, stored k-major
            const int idx = tid + 256 * i;
            const int r = idx >> 4, k = idx & 15;
            const int gm = m0 + r, gk = k0 + k;
            As[k][r] = (gm < M && gk < K) ? X[(size_t)gm * ldx + gk] : 0.0f;
        }
        {   // W tile: 16 k x 64 n, one float4 per thread
            const int k = tid >> 4, c4 = (tid & 15) * 4;
            const int gk = k0 + k, gn = n0 + c4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gk < K && gn + 3 < ldw) v = *reinterpret_cast<const float4*>(WT + (size_t)gk * ldw + gn);
            else if (gk < K) {
                if (gn + 0 < ldw) v.x = WT[(size_t)gk * ldw + gn + 0];
                if (gn + 1 < ldw) v.y = WT[(size_t)gk * ldw + gn + 1];
                if (gn + 2 < ldw) v.z = WT[(size_t)gk * ldw + gn + 2];
            }
            *reinterpret_cast<float4*>(&Bs[k][c4]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < LBK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j] + (bias ? bias[gn] : 0.0f);
            if (act == 1) v = silu_f(v);
            if (R) v += R[(size_t)gm * ldr + gn];
            Y[(size_t)gm * ldy + gn] = v;
        }
    }
}

int launch_linear(const float* X, int ldx, const float* WT, int ldw, const float* bias, const float* R,
                  int ldr, float* Y, int ldy, int M, int K, int N, int act, cudaStream_t st) {
    if (M <= 0 || N <= 0) return 0;
    KPD_REQUIRE(ldw % 4 == 0 && (reinterpret_cast<uintptr_t>(WT) & 15) == 0,
                "linear: WT must be 16-byte aligned with ldw %% 4 == 0 (ldw=%d)", ldw);
    dim3 grid(cdiv(N, LBN), cdiv(M, LBM));
    linear_kernel<<<grid, 256, 0, st>>>(X, ldx, WT, ldw, bias, R, ldr, Y, ldy, M, K, N, act);
    return check_launch("linear_kernel");
}

// LayerNorm over the first H columns of each row (eps inside the sqrt, biased variance --
// torch.nn.LayerNorm); one warp per row.  in/out may alias.
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ in, int ldi, float* __restrict__ out, int ldo, int M, int H,
                 const float* __restrict__ w, const float* __restrict__ b, float eps) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* x = in + (size_t)row * ldi;
    float s = 0.f;
    for (int c = lane; c < H; c += 32) s += x[c];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)H;
    float v = 0.f;
    for (int c = lane; c < H; c += 32) { const float d = x[c] - mean; v += d * d; }
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = 1.0f / sqrtf(v / (float)H + eps);
    float* y = out + (size_t)row * ldo;
    for (int c = lane; c < H; c += 32) y[c] = (x[c] - mean) * rstd * w[c] + b[c];
}

// two independent LayerNorm problems (the ligand and the keypoint rows of one layer) in ONE launch: these kernels are a few
// microseconds of pure latency each, so a launch saved is its whole duration saved.  Same per-row arithmetic as above.
struct LayerNormPair { const float* in[2]; float* out[2]; const float* w[2]; const float* b[2]; int M[2]; int ldi, ldo, H, blocks0; };
__global__ void __launch_bounds__(256) layernorm_pair_kernel(const LayerNormPair p, float eps) {
    const int k = (int)blockIdx.x >= p.blocks0 ? 1 : 0;
    const int row = ((int)blockIdx.x - (k ? p.blocks0 : 0)) * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= p.M[k]) return;
    const int H = p.H;
    const float* x = p.in[k] + (size_t)row * p.ldi;
    const float* w = p.w[k];
    const float* b = p.b[k];
    float s = 0.f;
    for (int c = lane; c < H; c += 32) s += x[c];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)H;
    float v = 0.f;
    for (int c = lane; c < H; c += 32) { const float d = x[c] - mean; v += d * d; }
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = 1.0f / sqrtf(v / (float)H + eps);
    float* y = p.out[k] + (size_t)row * p.ldo;
    for (int c = lane; c < H; c += 32) y[c] = (x[c] - mean) * rstd * w[c] + b[c];
}

int launch_layernorm_pair(const float* in0, float* out0, int M0, const float* w0, const float* b0, const float* in1, float* out1,
                          int M1, const float* w1, const float* b1, int ldi, int ldo, int H, cudaStream_t st) {
    LayerNormPair p;
    p.in[0] = in0; p.out[0] = out0; p.w[0] = w0; p.b[0] = b0; p.M[0] = M0 > 0 ? M0 : 0;
    p.in[1] = in1; p.out[1] = out1; p.w[1] = w1; p.b[1] = b1; p.M[1] = M1 > 0 ? M1 : 0;
    p.ldi = ldi; p.ldo = ldo; p.H = H; p.blocks0 = cdiv(p.M[0], 8);
    const int blocks = p.blocks0 + cdiv(p.M[1], 8);
    if (blocks <= 0) return 0;
    layernorm_pair_kernel<<<blocks, 256, 0, st>>>(p, 1e-5f);
    return check_launch("layernorm_pair_kernel");
}

int launch_layernorm(const float* in, int ldi, float* out, int ldo, int M, int H, const float* w,
                     const float* b, cudaStream_t st) {
    if (M <= 0) return 0;
    layernorm_kernel<<<cdiv(M, 8), 256, 0, st>>>(in, ldi, out, ldo, M, H, w, b, 1e-5f);
    return check_launch("layernorm_kernel");
}

// out[n, col] = t  where t = t_ptr[per_complex ? batch[n] : 0]
__global__ void set_time_col_kernel(float* out, int ld, int col, int n, const float* t_ptr,
                                    const int* batch, int per_complex) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[(size_t)i * ld + col] = t_ptr[per_complex ? batch[i] : 0];
}

int launch_set_time_col(float* out, int ld, int col, int n, const float* t_ptr, const int* batch,
                        int per_complex, cudaStream_t st) {
    if (n <= 0) return 0;
    set_time_col_kernel<<<cdiv(n, 256), 256, 0, st>>>(out, ld, col, n, t_ptr, batch, per_complex);
    return check_launch("set_time_col_kernel");
}

// out[n, 0:w] = in[n, 0:w]; out[n, w] = t      (GVP: time is concatenated BEFORE the encoders,
// models/dynamics_gvp.py:161-165)
__global__ void concat_time_kernel(const float* in, int w, float* out, int ldo, int n, const float* t_ptr,
                                   const int* batch, int per_complex) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * (w + 1)) return;
    const int r = i / (w + 1), c = i % (w + 1);
    out[(size_t)r * ldo + c] = c < w ? in[(size_t)r * w + c] : t_ptr[per_complex ? batch[r] : 0];
}

int launch_concat_time(const float* in, int w, float* out, int ldo, int n, const float* t_ptr,
                       const int* batch, int per_complex, cudaStream_t st) {
    if (n <= 0) return 0;
    concat_time_kernel<<<cdiv(n * (w + 1), 256), 256, 0, st>>>(in, w, out, ldo, n, t_ptr, batch, per_complex);
    return check_launch("concat_time_kernel");
}

__global__ void copy_rows_kernel(const float* in, int ldi, float* out, int ldo, int n, int w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * w) return;
    const int r = i / w, c = i % w;
    out[(size_t)r * ldo + c] = in[(size_t)r * ldi + c];
}

int launch_copy_rows(const float* in, int ldi, float* out, int ldo, int n, int w, cudaStream_t st) {
    if (n <= 0 || w <= 0) return 0;
    copy_rows_kernel<<<cdiv(n * w, 256), 256, 0, st>>>(in, ldi, out, ldo, n, w);
    return check_launch("copy_rows_kernel");
}

// out = a - b  (eps_x = x_final - x_0, models/dynamics.py:381)
__global__ void sub_kernel(const float* a, const float* b, float* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] - b[i];
}

int launch_sub(const float* a, const float* b, float* out, int n, cudaStream_t st) {
    if (n <= 0) return 0;
    sub_kernel<<<cdiv(n, 256), 256, 0, st>>>(a, b, out, n);
    return check_launch("sub_kernel");
}

}  // namespace kpd

extern "C" const char* kpd_last_error(void) { return kpd::g_err; }
extern "C" int64_t kpd_launch_count(void) { return kpd::launch_count(); }

extern "C" int kpd_profile_enable(int32_t kernel_id, int32_t max_records) {
    using namespace kpd;
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_used = 0;
    g_prof_open = false;
    g_prof_id = kernel_id;
    if (kernel_id <= 0) return 0;
    g_prof.resize(max_records > 0 ? max_records : 0);
    for (auto& r : g_prof) {
        if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) {
            set_error("kpd_profile_enable: cudaEventCreate failed");
            return -1;
        }
    }
    return 0;
}

extern "C" int kpd_profile_collect(double* total_ms, int32_t* count) {
    using namespace kpd;
    double tot = 0.0;
    for (size_t i = 0; i < g_prof_used; ++i) {
        cudaEventSynchronize(g_prof[i].b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_prof[i].a, g_prof[i].b) == cudaSuccess) tot += ms;
    }
    if (total_ms) *total_ms = tot;
    if (count) *count = (int32_t)g_prof_used;
    g_prof_used = 0;
    return 0;
}
extern "C" int kpd_version(void) { return 100; }

extern "C" int kpd_linear(const float* X, int32_t ldx, const float* WT, int32_t ldw, const float* bias,
                          const float* R, int32_t ldr, float* Y, int32_t ldy, int32_t M, int32_t K,
                          int32_t N, int32_t act, void* stream) {
    return kpd::launch_linear(X, ldx, WT, ldw, bias, R, ldr, Y, ldy, M, K, N, act,
                              static_cast<cudaStream_t>(stream));
}
