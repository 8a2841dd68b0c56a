// Interface of the general bf16x3 tensor-core GEMM (gemm_tc.cu).
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

namespace rb {

struct GemmArgs {
    const float* A; int lda, a_mn;   // x_mn = 0: element (row, k) at X[row * ld + k];  1: at X[k * ld + row]
    const float* B; int ldb, b_mn;
    float* C; int ldc;               // C[M][N] row-major
    int M, N, K;
    const float* bias;               // [N] or NULL
    int act, accumulate;             // act: 0 none, 1 tanh;  accumulate: C += result
    const float* H; int ldh;         // optional: result *= 1 - H[m][n]^2   (tanh backward)
    int ksplit;                      // set by gemm_bf16x3
    float* partial;                  // set by gemm_bf16x3
};
int gemm_pick_split(int M, int N, int K, int sms);
int gemm_bf16x3(GemmArgs g, float* splitk_ws, size_t splitk_ws_floats, int sms, cudaStream_t st);

}  // namespace rb
