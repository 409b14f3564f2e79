// Per-waypoint pose machinery shared by the pose kernels (k_pose.cu) and the fused pose-step + metrics kernel of the
// single-path LM loop (k_metrics.cu): FK sinks, q row load / store, joint-limit clamp, geometric Jacobian and the
// pose-only LM update (levenberg_marquardt_only_pose, optimization.py:61-92; clamp_to_joint_limits,
// optimization_utils.py:823-833).
#pragma once
#include "kinematics.cuh"
#include "linalg.cuh"

namespace cppflow {

template <class M>
struct JointSink {
    float a[M::NDOF][3];
    float o[M::NDOF][3];
    template <int D>
    __device__ __forceinline__ void joint(std::integral_constant<int, D>, const float* axis, const float* origin) {
#pragma unroll
        for (int r = 0; r < 3; ++r) { a[D][r] = axis[r]; o[D][r] = origin[r]; }
    }
    template <int F>
    __device__ __forceinline__ void frame(std::integral_constant<int, F>, const Frame&) {}
};

struct NullSink {
    template <int D>
    __device__ __forceinline__ void joint(std::integral_constant<int, D>, const float*, const float*) {}
    template <int F>
    __device__ __forceinline__ void frame(std::integral_constant<int, F>, const Frame&) {}
};

template <class M>
__device__ __forceinline__ void load_q(const float* __restrict__ q, int64_t i, float (&x)[M::NDOF]) {
    if constexpr (M::NDOF == 8) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(q + i * 8));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(q + i * 8) + 1);
        x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w;
        x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
    } else {
#pragma unroll
        for (int d = 0; d < M::NDOF; ++d) x[d] = __ldg(q + i * M::NDOF + d);
    }
}

template <class M>
__device__ __forceinline__ void store_q(float* __restrict__ q, int64_t i, const float (&x)[M::NDOF]) {
    if constexpr (M::NDOF == 8) {
        reinterpret_cast<float4*>(q + i * 8)[0] = make_float4(x[0], x[1], x[2], x[3]);
        reinterpret_cast<float4*>(q + i * 8)[1] = make_float4(x[4], x[5], x[6], x[7]);
    } else {
#pragma unroll
        for (int d = 0; d < M::NDOF; ++d) q[i * M::NDOF + d] = x[d];
    }
}

template <class M>
__device__ __forceinline__ void clamp_limits(float (&x)[M::NDOF]) {
    static_for<M::NDOF>([&](auto Dd) {
        constexpr int d = decltype(Dd)::value;
        x[d] = fminf(fmaxf(x[d], dof_lower<M>(d)), dof_upper<M>(d));
    });
}

// J[r][d], rows 0-2 angular, 3-5 linear (optimization.py:77-80)
template <class M>
__device__ __forceinline__ void geometric_jacobian(const JointSink<M>& js, const Frame& F, float (&J)[6][M::NDOF]) {
    static_for<M::NDOF>([&](auto Dd) {
        constexpr int d = decltype(Dd)::value;
        if constexpr (dof_is_prismatic<M>(d)) {
            J[0][d] = 0.f; J[1][d] = 0.f; J[2][d] = 0.f;
            J[3][d] = js.a[d][0]; J[4][d] = js.a[d][1]; J[5][d] = js.a[d][2];
        } else {
            const float r[3] = {F.p[0] - js.o[d][0], F.p[1] - js.o[d][1], F.p[2] - js.o[d][2]};
            float v[3];
            cross3(js.a[d], r, v);
            J[0][d] = js.a[d][0]; J[1][d] = js.a[d][1]; J[2][d] = js.a[d][2];
            J[3][d] = v[0]; J[4][d] = v[1]; J[5][d] = v[2];
        }
    });
}

// Pose-only LM step of ONE waypoint in registers.  (J^T J + lambda I) dx = J^T e is solved in its dual form
// dx = J^T (J J^T + lambda I)^-1 e: algebraically identical, but 6x6 instead of DxD and free of the lambda-only
// null-space directions that make the primal fp32 solve lose ~1e-2 rad on 7/8-dof arms.  One step of iterative
// refinement with the residual formed through J (not J J^T) recovers cond(J) instead of cond(J)^2 accuracy.
// x <- clamp(x + dx); J_out / e_out (alpha-scaled, row i) may be null.
template <class M>
__device__ __forceinline__ void pose_lm_update(float (&x)[M::NDOF], const float (&tg)[7], float alpha_pos, float alpha_rot,
                                               float lambda, int do_clamp, float* __restrict__ J_out,
                                               float* __restrict__ e_out, int64_t i) {
    constexpr int D = M::NDOF;
    JointSink<M> js;
    Frame F;
    fk_chain<M>(x, js, F);
    float e[6];
    pose_error(tg, F, e);
    float J[6][D];
    geometric_jacobian<M>(js, F, J);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        e[r] *= alpha_rot;
        e[r + 3] *= alpha_pos;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            J[r][d] *= alpha_rot;
            J[r + 3][d] *= alpha_pos;
        }
    }
    if (J_out) {
        float* o = J_out + i * 6 * D;
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int d = 0; d < D; ++d) o[r * D + d] = J[r][d];
    }
    if (e_out) {
#pragma unroll
        for (int r = 0; r < 6; ++r) e_out[i * 6 + r] = e[r];
    }
    float A[6][6], dinv[6];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) {
            float s = (r == c) ? lambda : 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) s = fmaf(J[r][d], J[c][d], s);
            A[r][c] = s;
        }
    chol_lower<6>(A, dinv, fmaxf(lambda, 1e-30f));  // exact pivots of J J^T + lambda I are >= lambda
    float z[6];
    chol_solve<6>(A, dinv, e, z);
    float dx[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 6; ++r) s = fmaf(J[r][d], z[r], s);
        dx[d] = s;
    }
    // refinement: rho = e - J dx - lambda z
    float rho[6], dz[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        float s = fmaf(-lambda, z[r], e[r]);
#pragma unroll
        for (int d = 0; d < D; ++d) s = fmaf(-J[r][d], dx[d], s);
        rho[r] = s;
    }
    chol_solve<6>(A, dinv, rho, dz);
    float step[D], chk = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        float s = dx[d];
#pragma unroll
        for (int r = 0; r < 6; ++r) s = fmaf(J[r][d], dz[r], s);
        step[d] = s;
        chk += s;
    }
    // a non-finite update (it cannot come from the floored factorisation, only from non-finite inputs) must not be
    // turned into "joint at its limit" by the clamp below: the waypoint is left where it was
    const bool ok = fabsf(chk) <= 3.0e38f;
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] += ok ? step[d] : 0.f;
    if (do_clamp) clamp_limits<M>(x);
}

}  // namespace cppflow
