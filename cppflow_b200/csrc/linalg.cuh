// Small dense SPD helpers, fully unrolled so the matrices live in registers.
#pragma once
#include <cuda_runtime.h>

namespace cppflow {

// packed lower-triangular index, row-major: (i,j) with j <= i
__host__ __device__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// In-place Cholesky A = L L^T on the lower triangle of A[N][N]; the diagonal of L is stored as its RECIPROCAL
// in dinv[] (A[i][i] keeps L[i][i]).  Pivots are floored at `floor`: the callers factor matrices of the form
// (PSD + lambda I), whose exact pivots (Schur complements) are >= lambda, so floor = lambda only ever replaces a pivot
// that rounding pushed below its mathematical lower bound - it cannot blow a near-singular waypoint up to 1e15 the way
// an absolute 1e-30 floor did.
template <int N>
__device__ __forceinline__ void chol_lower(float (&A)[N][N], float (&dinv)[N], float floor = 1e-30f) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
        float s = A[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) s = fmaf(-A[j][k], A[j][k], s);
        s = fmaxf(s, floor);
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

// Reciprocal with one Newton step on top of MUFU.RCP: <= 1 ulp, no slow path / branch.
__device__ __forceinline__ float rcp_nr(float p) {
    float d;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(p));
    const float e = fmaf(-p, d, 1.f);
    return fmaf(d, e, d);
}

// Symmetric sweep operator on the packed lower triangle a[] of an SPD matrix A, carrying one right-hand side:
// after sweeping every pivot, a = -A^-1 and y = A^-1 y.  Sweeping pivot k:
//     d = 1 / a_kk;  a_ij -= a_ik a_jk d (i, j != k);  a_ik = a_ik d;  a_kk = -d;  y_i -= a_ik d y_k;  y_k = y_k d.
// Every intermediate is symmetric, so only the N (N + 1) / 2 packed entries are touched: N (N + 1) / 2 - N + N = 28 + 7
// FMAs per pivot for N = 8, all independent of each other (the dependent chain is one reciprocal + one FMA per pivot) -
// about 0.6x the instructions of Cholesky + triangular inverse + L^-T L^-1, with a much shorter critical path (measured
// in the block solve: 1024 vs 1284 cycles per block, same error against the fp64 oracle).  The
// unswept part stays the (positive definite) Schur complement, so the pivots are positive (>= lambda for J^T J +
// lambda I); they are floored at `floor` like the Cholesky pivots.
template <int N>
__device__ __forceinline__ void sweep_neg_inverse(float (&a)[N * (N + 1) / 2], float (&y)[N], float floor = 1e-30f) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const float d = rcp_nr(fmaxf(a[tri(k, k)], floor));
        float c[N], cd[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            c[i] = i == k ? 0.f : (i > k ? a[tri(i, k)] : a[tri(k, i)]);
            cd[i] = c[i] * d;
        }
        const float yk = y[k];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (i == k) continue;
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                if (j == k) continue;
                a[tri(i, j)] = fmaf(-cd[i], c[j], a[tri(i, j)]);
            }
            y[i] = fmaf(-cd[i], yk, y[i]);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (i == k) continue;
            if (i > k) a[tri(i, k)] = cd[i]; else a[tri(k, i)] = cd[i];
        }
        a[tri(k, k)] = -d;
        y[k] = yk * d;
    }
}

}  // namespace cppflow
