// Batched tcgen05 linear layers (tc_gemm.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kpd {

struct TcLinProblem {
    const float* X; const uint4* Wp; const float* bias; const float* R; float* Y;
    int ldx, ldr, ldy, M, K, N, act;
};
struct TcLinBatch {
    TcLinProblem p[2];      // blockIdx.z selects the problem (e.g. the ligand and the keypoint rows of one layer)
    int NBmax;              // widest column block of the launch (ring stage size)
    int kmax;               // largest K of the launch (A tile size)
    int bpc;                // 256-column blocks per CTA (blockIdx.y strides over groups of bpc blocks)
    int stages;             // weight-ring stages (set by the launcher: as many as fit beside the A tile)
};

TcLinProblem tc_problem(const float* X, int ldx, const void* Wp, const float* bias, const float* R, int ldr, float* Y, int ldy,
                        int M, int K, int N, int act);
// Y = act(X W^T + b) (+R) for 1 or 2 problems in one launch; nsplit 1 = bf16, 2 = bf16x3 (split operands)
int launch_tc_batch(TcLinBatch& B, int nprob, int nsplit, cudaStream_t st);
int launch_tc_linear(const float* X, int ldx, const void* Wp, const float* bias, const float* R, int ldr, float* Y,
                     int ldy, int M, int K, int N, int act, int nsplit, cudaStream_t st);

}  // namespace kpd
