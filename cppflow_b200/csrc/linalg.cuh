// Small dense SPD helpers, fully unrolled so the matrices live in registers.
#pragma once
#include <cuda_runtime.h>

namespace cppflow {

// packed lower-triangular index, row-major: (i,j) with j <= i
__host__ __device__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// In-place Cholesky A = L L^T on the lower triangle of A[N][N]; the diagonal of L is stored as its RECIPROCAL
// in dinv[] (A[i][i] keeps L[i][i]).  Pivots are floored to keep a rounding-negative pivot from producing NaNs.
template <int N>
__device__ __forceinline__ void chol_lower(float (&A)[N][N], float (&dinv)[N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
        float s = A[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) s = fmaf(-A[j][k], A[j][k], s);
        s = fmaxf(s, 1e-30f);
        const float r = rsqrtf(s);
        dinv[j] = r;
        A[j][j] = s * r;
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            float v = A[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) v = fmaf(-A[i][k], A[j][k], v);
            A[i][j] = v * r;
        }
    }
}

// x = (L L^T)^-1 b
template <int N>
__device__ __forceinline__ void chol_solve(const float (&L)[N][N], const float (&dinv)[N], const float (&b)[N],
                                           float (&x)[N]) {
    float y[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        float v = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) v = fmaf(-L[i][k], y[k], v);
        y[i] = v * dinv[i];
    }
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
        float v = y[i];
#pragma unroll
        for (int k = i + 1; k < N; ++k) v = fmaf(-L[k][i], x[k], v);
        x[i] = v * dinv[i];
    }
}

// Given the Cholesky factor L (lower, reciprocal diagonal in dinv) overwrite the lower triangle of S with
// (L L^T)^-1 = W^T W, W = L^-1.
template <int N>
__device__ __forceinline__ void chol_inverse(const float (&L)[N][N], const float (&dinv)[N], float (&S)[N][N]) {
    float W[N][N];  // lower triangular inverse of L
#pragma unroll
    for (int j = 0; j < N; ++j) {
        W[j][j] = dinv[j];
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            float v = 0.f;
#pragma unroll
            for (int k = j; k < i; ++k) v = fmaf(-L[i][k], W[k][j], v);
            W[i][j] = v * dinv[i];
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            float v = 0.f;
#pragma unroll
            for (int k = i; k < N; ++k) v = fmaf(W[k][i], W[k][j], v);
            S[i][j] = v;
        }
}

}  // namespace cppflow
