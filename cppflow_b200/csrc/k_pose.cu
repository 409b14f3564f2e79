// K1: forward kinematics, geometric Jacobian, 6-d pose error and the pose-only Levenberg-Marquardt step.
// One thread per waypoint; everything (chain frames, 6 x D Jacobian, 6 x 6 normal equations) stays in registers.
//
// Reference being replaced:
//   jrl Robot.forward_kinematics / Robot.jacobian             (optimization_utils.py:811, optimization.py:74)
//   get_6d_pose_errors                                        (optimization_utils.py:802-820)
//   levenberg_marquardt_only_pose                             (optimization.py:61-92)
//   clamp_to_joint_limits                                     (optimization_utils.py:823-833)
#include "common.cuh"
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

template <class M>
__global__ void __launch_bounds__(128) fk_kernel(const float* __restrict__ q, int64_t n, float* __restrict__ poses) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x[M::NDOF];
    load_q<M>(q, i, x);
    NullSink sink;
    Frame F;
    fk_chain<M>(x, sink, F);
    float qt[4];
    rotmat_to_quat(F.R, qt);
    float* o = poses + i * 7;
    o[0] = F.p[0]; o[1] = F.p[1]; o[2] = F.p[2];
    o[3] = qt[0]; o[4] = qt[1]; o[5] = qt[2]; o[6] = qt[3];
}

template <class M>
__global__ void __launch_bounds__(128) jacobian_kernel(const float* __restrict__ q, int64_t n, float* __restrict__ Jout) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x[M::NDOF];
    load_q<M>(q, i, x);
    JointSink<M> js;
    Frame F;
    fk_chain<M>(x, js, F);
    float J[6][M::NDOF];
    geometric_jacobian<M>(js, F, J);
    float* o = Jout + i * 6 * M::NDOF;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int d = 0; d < M::NDOF; ++d) o[r * M::NDOF + d] = J[r][d];
}

template <class M>
__global__ void __launch_bounds__(128)
pose_error_kernel(const float* __restrict__ q, const float* __restrict__ target, int64_t n, int64_t n_targets,
                  float* __restrict__ err, float* __restrict__ cur) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x[M::NDOF];
    load_q<M>(q, i, x);
    NullSink sink;
    Frame F;
    fk_chain<M>(x, sink, F);
    float tg[7];
    const float* tp = target + (i % n_targets) * 7;
#pragma unroll
    for (int k = 0; k < 7; ++k) tg[k] = __ldg(tp + k);
    float e[6];
    pose_error(tg, F, e);
    if (err) {
#pragma unroll
        for (int k = 0; k < 6; ++k) err[i * 6 + k] = e[k];
    }
    if (cur) {
        float qt[4];
        rotmat_to_quat(F.R, qt);
        float* o = cur + i * 7;
        o[0] = F.p[0]; o[1] = F.p[1]; o[2] = F.p[2];
        o[3] = qt[0]; o[4] = qt[1]; o[5] = qt[2]; o[6] = qt[3];
    }
}

// Pose-only LM step.  (J^T J + lambda I) dx = J^T e is solved in its dual form dx = J^T (J J^T + lambda I)^-1 e:
// algebraically identical, but 6x6 instead of DxD and free of the lambda-only null-space directions that make the
// primal fp32 solve lose ~1e-2 rad on 7/8-dof arms.  One step of iterative refinement with the residual formed
// through J (not J J^T) recovers cond(J) instead of cond(J)^2 accuracy.
template <class M>
__global__ void __launch_bounds__(128)
lm_pose_step_kernel(const float* __restrict__ q, const float* __restrict__ target, int64_t n, int64_t n_targets,
                    float alpha_pos, float alpha_rot, float lambda, int do_clamp, float* __restrict__ x_out,
                    float* __restrict__ J_out, float* __restrict__ e_out) {
    constexpr int D = M::NDOF;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x[D];
    load_q<M>(q, i, x);
    JointSink<M> js;
    Frame F;
    fk_chain<M>(x, js, F);
    float tg[7];
    const float* tp = target + (i % n_targets) * 7;
#pragma unroll
    for (int k = 0; k < 7; ++k) tg[k] = __ldg(tp + k);
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
    chol_lower<6>(A, dinv);
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
#pragma unroll
    for (int d = 0; d < D; ++d) {
        float s = dx[d];
#pragma unroll
        for (int r = 0; r < 6; ++r) s = fmaf(J[r][d], dz[r], s);
        x[d] += s;
    }
    if (do_clamp) clamp_limits<M>(x);
    store_q<M>(x_out, i, x);
}

template <class M>
__global__ void __launch_bounds__(256) clamp_kernel(float* __restrict__ q, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x[M::NDOF];
    load_q<M>(q, i, x);
    clamp_limits<M>(x);
    store_q<M>(q, i, x);
}

}  // namespace cppflow

using namespace cppflow;

extern "C" int cppflow_forward_kinematics(int robot, const float* d_q, int64_t n, float* d_poses, void* stream) {
    CPPFLOW_CHECK_ARG(n >= 0, "n");
    if (n == 0) return CPPFLOW_OK;
    CPPFLOW_CHECK_ARG(d_q && d_poses, "null pointer");
    CPPFLOW_DISPATCH_ROBOT(robot, fk_kernel<M><<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(d_q, n, d_poses));
    CPPFLOW_CHECK_LAUNCH();
    return CPPFLOW_OK;
}

extern "C" int cppflow_jacobian(int robot, const float* d_q, int64_t n, float* d_J, void* stream) {
    CPPFLOW_CHECK_ARG(n >= 0, "n");
    if (n == 0) return CPPFLOW_OK;
    CPPFLOW_CHECK_ARG(d_q && d_J, "null pointer");
    CPPFLOW_DISPATCH_ROBOT(robot, jacobian_kernel<M><<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(d_q, n, d_J));
    CPPFLOW_CHECK_LAUNCH();
    return CPPFLOW_OK;
}

extern "C" int cppflow_pose_errors(int robot, const float* d_q, const float* d_target, int64_t n, int64_t n_targets,
                                   float* d_err, float* d_cur_poses, void* stream) {
    CPPFLOW_CHECK_ARG(n >= 0, "n");
    if (n == 0) return CPPFLOW_OK;
    CPPFLOW_CHECK_ARG(d_q && d_target, "null pointer");
    CPPFLOW_CHECK_ARG(n_targets > 0, "n_targets");
    CPPFLOW_DISPATCH_ROBOT(robot, pose_error_kernel<M><<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(
                                      d_q, d_target, n, n_targets, d_err, d_cur_poses));
    CPPFLOW_CHECK_LAUNCH();
    return CPPFLOW_OK;
}

static int robot_dof(int robot) {
    switch (robot) {
        case ROBOT_FETCH: return Fetch::NDOF;
        case ROBOT_FETCH_ARM: return FetchArm::NDOF;
        case ROBOT_PANDA: return Panda::NDOF;
        default: return 0;
    }
}

extern "C" int cppflow_lm_pose_step(int robot, const cppflow_lm_params* p, const float* d_q, const float* d_target,
                                    int64_t n, int64_t n_targets, int do_clamp, float* d_x_out, float* d_J_out,
                                    float* d_e_out, void* stream) {
    CPPFLOW_CHECK_ARG(p != nullptr, "params");
    CPPFLOW_CHECK_ARG(n >= 0, "n");
    if (n == 0) return CPPFLOW_OK;
    CPPFLOW_CHECK_ARG(d_q && d_target && d_x_out, "null pointer");
    CPPFLOW_CHECK_ARG(n_targets > 0, "n_targets");
    CPPFLOW_DISPATCH_ROBOT(robot, lm_pose_step_kernel<M><<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(
                                      d_q, d_target, n, n_targets, p->alpha_position, p->alpha_rotation, p->lm_lambda,
                                      do_clamp, d_x_out, d_J_out, d_e_out));
    CPPFLOW_CHECK_LAUNCH();
    return CPPFLOW_OK;
}

extern "C" int cppflow_lm_pose_steps(int robot, const cppflow_lm_params* p, const float* h_lambdas, int n_steps,
                                     float* d_q, float* d_tmp, const float* d_target, int64_t n, int64_t n_targets,
                                     int do_clamp, void* stream) {
    CPPFLOW_CHECK_ARG(p != nullptr && n >= 0 && n_steps >= 0, "params, n, n_steps");
    if (n == 0 || n_steps == 0) return CPPFLOW_OK;
    CPPFLOW_CHECK_ARG(d_q && d_tmp && d_target && h_lambdas && n_targets > 0, "null pointer");
    float* bufs[2] = {d_q, d_tmp};
    for (int i = 0; i < n_steps; ++i) {
        const float lam = h_lambdas[i];
        CPPFLOW_DISPATCH_ROBOT(robot, lm_pose_step_kernel<M><<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(
                                          bufs[i & 1], d_target, n, n_targets, p->alpha_position, p->alpha_rotation, lam,
                                          do_clamp, bufs[(i + 1) & 1], nullptr, nullptr));
    }
    CPPFLOW_CHECK_LAUNCH();
    if (n_steps & 1) {  // the result sits in d_tmp: bring it home
        cudaError_t e = cudaMemcpyAsync(d_q, d_tmp, (size_t)n * robot_dof(robot) * sizeof(float), cudaMemcpyDeviceToDevice,
                                        (cudaStream_t)stream);
        if (e != cudaSuccess) return fail(CPPFLOW_E_CUDA, "lm_pose_steps: %s", cudaGetErrorString(e));
    }
    return CPPFLOW_OK;
}

extern "C" int cppflow_clamp_to_joint_limits(int robot, float* d_q, int64_t n, void* stream) {
    CPPFLOW_CHECK_ARG(n >= 0, "n");
    if (n == 0) return CPPFLOW_OK;
    CPPFLOW_CHECK_ARG(d_q != nullptr, "null pointer");
    CPPFLOW_DISPATCH_ROBOT(robot, clamp_kernel<M><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(d_q, n));
    CPPFLOW_CHECK_LAUNCH();
    return CPPFLOW_OK;
}
