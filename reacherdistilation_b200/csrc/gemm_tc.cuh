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
    int batch;                       // >= 1: independent products, operand b of batch i at X + i * stride (elements); 0 is treated as 1
    long long sA, sB, sC, sBias, sH;
    int force_split;                 // > 0: use this many K slices (when a workspace is given)
    int ksplit, nsplit;              // set by gemm_bf16x3
    float* partial;                  // set by gemm_bf16x3
};
int gemm_bf16x3(GemmArgs g, float* splitk_ws, size_t splitk_ws_floats, int sms, cudaStream_t st);

// column sums (bias gradients) of `batch` row-major blocks X[b] (rows x n, leading dimension ld) in two fixed-order stages;
// colpart = COLPART_FLOATS floats of scratch
constexpr size_t COLPART_FLOATS = 64 * 1024;     // <= 64 row blocks x (n x batch <= 1024)
int colsum(const float* X, int ld, int64_t rows, int n, int batch, long long sX, float* out, long long sOut, float* colpart, cudaStream_t st);
// out[0] = x[0] + x[1] + ... in index order (one thread: the deterministic tail of the loss reductions)
int sum_serial(const float* x, int n, float* out, cudaStream_t st);

}  // namespace rb
