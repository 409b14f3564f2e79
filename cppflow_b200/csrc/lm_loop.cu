// Host-side driver of the alternating LM loop for ONE path: the control flow of run_lm_alternating_loss
// (optimization.py:147-373) in C++, on top of the kernels' C entry points.
//
// The reference decides the type of the next step from the validity of the current iterate (pose-only until position
// and rotation errors are inside the constraints, joint-differencing otherwise, :219-260), watches the trajectory length
// after differencing steps for convergence (:268-283) and keeps the last valid iterate (:318-352).  Every iteration
// therefore needs a device -> host read of the path metrics (the reference's `.item()` calls, :175).  Here that read
// is 8 floats into page-locked memory and one cudaStreamSynchronize; the kernel launches, the decision logic and the
// bookkeeping of an iteration cost ~10 us of host time instead of ~55 us when driven from Python through ctypes
// (the per-plan latency of fetch__circle is dominated by these 20 round trips, not by the kernels).
#include <chrono>
#include <cmath>

#include "common.cuh"

using namespace cppflow;

namespace {

inline size_t align256(size_t n) { return (n + 255) / 256 * 256; }

struct LoopLayout {
    size_t x_bytes, off_xa, off_xb, off_valid, off_metrics, off_lm, total;
    LoopLayout(int robot, int ndof, int64_t T) {
        x_bytes = align256((size_t)T * ndof * sizeof(float));
        off_xa = 0;
        off_xb = off_xa + x_bytes;
        off_valid = off_xb + x_bytes;
        off_metrics = off_valid + x_bytes;
        off_lm = off_metrics + 256;
        total = off_lm + align256(cppflow_lm_full_workspace_bytes(robot, 1, T));
    }
};

int robot_ndof(int robot) {
    cppflow_robot_info info;
    if (cppflow_robot_info_get(robot, &info) != CPPFLOW_OK) return -1;
    return info.ndof;
}

}  // namespace

extern "C" size_t cppflow_lm_alternating_workspace_bytes(int robot, int64_t T) {
    const int ndof = robot_ndof(robot);
    if (ndof < 0 || T <= 0) return 0;
    return LoopLayout(robot, ndof, T).total;
}

extern "C" int cppflow_lm_alternating_loss(int robot, const cppflow_lm_params* params_diff,
                                           const cppflow_lm_params* params_pose, const cppflow_constraints* constraints,
                                           const float* d_x_seed, const float* d_target, int64_t T,
                                           const float* h_cuboids, const float* h_Tcuboids, int n_obstacles,
                                           int max_n_steps, double tmax_sec, int return_if_valid_after_n_steps,
                                           double convergence_threshold, void* d_workspace, size_t workspace_bytes,
                                           float* h_pinned_metrics, float* d_x_out, cppflow_lm_loop_result* result,
                                           void* stream) {
    CPPFLOW_CHECK_ARG(params_diff && params_pose && constraints && result, "null parameter struct");
    CPPFLOW_CHECK_ARG(d_x_seed && d_target && d_x_out && d_workspace && h_pinned_metrics, "null pointer");
    CPPFLOW_CHECK_ARG(T > 0 && max_n_steps >= 0, "T, max_n_steps");
    const int ndof = robot_ndof(robot);
    if (ndof < 0) return CPPFLOW_E_INVALID;
    const LoopLayout L(robot, ndof, T);
    if (workspace_bytes < L.total)
        return fail(CPPFLOW_E_WORKSPACE, "lm_alternating_loss: workspace too small (%zu < %zu)", workspace_bytes, L.total);
    CPPFLOW_CHECK_ARG(((uintptr_t)d_workspace & 255) == 0, "workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)d_workspace;
    float* x_cur = (float*)(ws + L.off_xa);
    float* x_new = (float*)(ws + L.off_xb);
    float* x_valid = (float*)(ws + L.off_valid);
    float* d_metrics = (float*)(ws + L.off_metrics);
    void* lm_ws = ws + L.off_lm;
    const size_t lm_ws_bytes = L.total - L.off_lm;
    const size_t row_bytes = (size_t)T * ndof * sizeof(float);

#define CUDA_OK(expr)                                                                                        \
    do {                                                                                                     \
        cudaError_t e_ = (expr);                                                                             \
        if (e_ != cudaSuccess) return fail(CPPFLOW_E_CUDA, "lm_alternating_loss: %s", cudaGetErrorString(e_)); \
    } while (0)

    CUDA_OK(cudaMemcpyAsync(x_cur, d_x_seed, row_bytes, cudaMemcpyDeviceToDevice, st));

    // the reference starts with (pose_pos_valid, pose_rot_valid) = (True, False): the first step is pose-only (:219-220)
    bool pose_pos_valid = true, pose_rot_valid = false, converged = false, has_valid = false;
    int last_valid_idx = -1, n_tls = 0, i = 0;
    double last_tl = 0.0;
    int n_sched = 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (i = 0; i < max_n_steps; ++i) {
        bool was_differencing;
        if (pose_pos_valid && pose_rot_valid) {
            // virtual configs = the current iterate (:253): their residual is identically zero -> d_xv = NULL
            if (int rc = cppflow_lm_full_step(robot, params_diff, x_cur, nullptr, d_target, 1, T, h_cuboids, h_Tcuboids,
                                              n_obstacles, CPPFLOW_LM_CLAMP, lm_ws, lm_ws_bytes, x_new, stream))
                return rc;
            was_differencing = true;
        } else {
            if (int rc = cppflow_lm_pose_step(robot, params_pose, x_cur, d_target, T, T, 1, x_new, nullptr, nullptr, stream))
                return rc;
            was_differencing = false;
        }
        if (n_sched < CPPFLOW_LM_SCHEDULE_MAX - 1) result->schedule[n_sched++] = was_differencing ? 'd' : 'p';
        float* tmp = x_cur; x_cur = x_new; x_new = tmp;  // clamp_to_joint_limits is fused into both steps (:259)

        if (int rc = cppflow_path_metrics(robot, x_cur, d_target, 1, T, h_cuboids, h_Tcuboids, n_obstacles, d_metrics, stream))
            return rc;
        CUDA_OK(cudaMemcpyAsync(h_pinned_metrics, d_metrics, 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));  // the one host round trip of the iteration
        const float* m = h_pinned_metrics;   // max_pos_cm, max_rot_deg, mjac_deg, mjac_cm, tl, min_self, min_env
        const double tl_new = m[4];
        if (was_differencing) {
            if (!converged && n_tls > 0) {
                if (std::fabs(tl_new - last_tl) < convergence_threshold) {
                    converged = true;
                    if (last_valid_idx == i - 1) break;
                }
            }
            last_tl = tl_new;
            ++n_tls;
        }
        // x_is_valid (:836-923): strict '<' thresholds (evaluation_utils.py:29-75), then the collision check
        pose_pos_valid = m[0] < constraints->max_allowed_position_error_cm;
        pose_rot_valid = m[1] < constraints->max_allowed_rotation_error_deg;
        const bool mjac_ok = m[2] < constraints->max_allowed_mjac_deg && m[3] < constraints->max_allowed_mjac_cm;
        const bool is_valid = pose_pos_valid && pose_rot_valid && mjac_ok && !(m[5] < 0.f) && !(m[6] < 0.f);
        if (is_valid) {
            last_valid_idx = i;
            has_valid = true;
            CUDA_OK(cudaMemcpyAsync(x_valid, x_cur, row_bytes, cudaMemcpyDeviceToDevice, st));
            if (converged) break;
        }
        const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (elapsed > tmax_sec) break;
        if (has_valid) {
            if (i > return_if_valid_after_n_steps) break;
            if (i > max_n_steps) break;
        }
    }
    if (i == max_n_steps && max_n_steps > 0) i = max_n_steps - 1;  // Python's loop variable after an exhausted range()
    CUDA_OK(cudaMemcpyAsync(d_x_out, has_valid ? x_valid : x_cur, row_bytes, cudaMemcpyDeviceToDevice, st));
    result->schedule[n_sched] = 0;
    result->n_steps_taken = i;
    result->is_valid = has_valid ? 1 : 0;
    for (int k = 0; k < 8; ++k) result->last_metrics[k] = max_n_steps > 0 ? h_pinned_metrics[k] : 0.f;
#undef CUDA_OK
    return CPPFLOW_OK;
}
