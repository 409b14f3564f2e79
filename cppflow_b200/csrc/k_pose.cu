// K1: forward kinematics, geometric Jacobian, 6-d pose error and the pose-only Levenberg-Marquardt step.
// One thread per waypoint; everything (chain frames, 6 x D Jacobian, 6 x 6 normal equations) stays in registers.
//
// Reference being replaced:
//   jrl Robot.forward_kinematics / Robot.jacobian             (optimization_utils.py:811, optimization.py:74)
//   get_6d_pose_errors                                        (optimization_utils.py:802-820)
//   levenberg_marquardt_only_pose                             (optimization.py:61-92)
//   clamp_to_joint_limits                                     (optimization_utils.py:823-833)
#include "common.cuh"
#include "pose_step.cuh"

namespace cppflow {

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

// Pose-only LM step, one thread per waypoint (pose_step.cuh: pose_lm_update).
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
    float tg[7];
    const float* tp = target + (i % n_targets) * 7;
#pragma unroll
    for (int k = 0; k < 7; ++k) tg[k] = __ldg(tp + k);
    pose_lm_update<M>(x, tg, alpha_pos, alpha_rot, lambda, do_clamp, J_out, e_out, i);
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
