// K3: the full Levenberg-Marquardt step (pose + joint-difference + virtual-config + capsule collision residuals)
// for P independent paths of T waypoints.
//
// Reference being replaced (dense, one path at a time, (T*D)^2 floats per path):
//   LmResidualFns.get_r_and_J           optimization_utils.py:486-731  (terms :264-285, :335-349, :430-483, :643-725)
//   _lm_full_step                       optimization.py:95-113         (J^T J + lambda I, Cholesky, 2 triangular solves)
//   levenberg_marquardt_full            optimization.py:116-144
//   clamp_to_joint_limits               optimization_utils.py:823-833
//
// Structure exploited: every residual row touches the D joints of ONE waypoint, except the differencing rows
// r = alpha * wrap(x[t+1] - x[t]) which couple neighbours with +-alpha * I.  Hence
//     A = J^T J + lambda I  is symmetric block-tridiagonal with dense D x D diagonal blocks
//         A_tt = Jp_t^T Jp_t + sum_active w^2 g g^T + beta * n_t + gamma^2 [t in virtual set] + lambda I
//     and off-diagonal blocks -diag(beta),  beta_d = (alpha_diff * prismatic_scale_d)^2,
//     b_t = Jp_t^T e_t - sum_active w^2 d g + beta * (wrap(x[t+1]-x[t]) - wrap(x[t]-x[t-1])) - gamma^2 wrap(x_t - xv_t).
// Kernel A (assembly) evaluates FK, the Jacobian, all capsule distances and the active gradients of one waypoint per
// thread and writes the packed (A_tt, b_t) block (44 floats for D = 8) to the workspace.
// Kernel B (solve) runs a block Cholesky (block Thomas) sweep per path, one thread per path, and writes
// clamp(x + dx).  Nothing of size (T*D)^2 is ever formed.
#include "common.cuh"
#include "collision.cuh"
#include "linalg.cuh"

namespace cppflow {

constexpr int ABLOCK = 128;

template <int D>
struct BlockLayout {
    static constexpr int NT = D * (D + 1) / 2;             // packed lower triangle
    static constexpr int NW = ((NT + D) + 3) / 4 * 4;      // floats per waypoint block, float4 aligned
};

struct AssembleParams {
    float lambda;
    float a_pos, a_rot;
    float w2_self, w2_env;  // alpha^2
    float gamma2;           // (alpha_virtual * alpha_diff)^2
    float beta[CPPFLOW_MAX_DOF];
    int use_pose, use_diff, use_virtual, n_virtual, use_self, use_env;
};

template <class M>
__device__ __forceinline__ void rank1_update(float (&A)[BlockLayout<M::NDOF>::NT], float (&b)[M::NDOF],
                                             const float (&g)[M::NDOF], float w2, float d) {
#pragma unroll
    for (int i = 0; i < M::NDOF; ++i) {
        const float wg = w2 * g[i];
        b[i] = fmaf(-d, wg, b[i]);
#pragma unroll
        for (int j = 0; j <= i; ++j) A[tri(i, j)] = fmaf(wg, g[j], A[tri(i, j)]);
    }
}

template <class M>
__global__ void __launch_bounds__(ABLOCK)
lm_assemble_kernel(const float* __restrict__ q, const float* __restrict__ xv, const float* __restrict__ target,
                   int64_t P, int64_t T, const Obstacles ob, const AssembleParams prm, float* __restrict__ ws) {
    constexpr int D = M::NDOF;
    constexpr int NT = BlockLayout<D>::NT;
    constexpr int NW = BlockLayout<D>::NW;
    extern __shared__ float smem[];
    const int64_t i = (int64_t)blockIdx.x * ABLOCK + threadIdx.x;
    if (i >= P * T) return;
    const int64_t t = i % T;
    float* sm = smem + threadIdx.x;

    float x[D];
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = __ldg(q + i * D + d);
    CollisionSink<M, ABLOCK, true> sink{sm};
    Frame F;
    fk_chain<M>(x, sink, F);

    float A[NT], b[D];
#pragma unroll
    for (int k = 0; k < NT; ++k) A[k] = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) b[d] = 0.f;

    if (prm.use_pose) {
        float tg[7];
        const float* tp = target + t * 7;
#pragma unroll
        for (int k = 0; k < 7; ++k) tg[k] = __ldg(tp + k);
        float e[6];
        pose_error(tg, F, e);
        float J[6][D];
        static_for<D>([&](auto Dd) {
            constexpr int d = decltype(Dd)::value;
            float a[3], o[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                a[r] = sm[(SmemLayout<M>::JOINTS + d * 6 + r) * ABLOCK];
                o[r] = sm[(SmemLayout<M>::JOINTS + d * 6 + 3 + r) * ABLOCK];
            }
            if constexpr (dof_is_prismatic<M>(d)) {
                J[0][d] = 0.f; J[1][d] = 0.f; J[2][d] = 0.f;
                J[3][d] = a[0] * prm.a_pos; J[4][d] = a[1] * prm.a_pos; J[5][d] = a[2] * prm.a_pos;
            } else {
                const float rr[3] = {F.p[0] - o[0], F.p[1] - o[1], F.p[2] - o[2]};
                float v[3];
                cross3(a, rr, v);
                J[0][d] = a[0] * prm.a_rot; J[1][d] = a[1] * prm.a_rot; J[2][d] = a[2] * prm.a_rot;
                J[3][d] = v[0] * prm.a_pos; J[4][d] = v[1] * prm.a_pos; J[5][d] = v[2] * prm.a_pos;
            }
        });
#pragma unroll
        for (int r = 0; r < 3; ++r) { e[r] *= prm.a_rot; e[r + 3] *= prm.a_pos; }
#pragma unroll
        for (int c = 0; c < D; ++c) {
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < 6; ++r) s = fmaf(J[r][c], e[r], s);
            b[c] = s;
#pragma unroll
            for (int c2 = 0; c2 <= c; ++c2) {
                float v = 0.f;
#pragma unroll
                for (int r = 0; r < 6; ++r) v = fmaf(J[r][c], J[r][c2], v);
                A[tri(c, c2)] = v;
            }
        }
    }

    if (prm.use_self) {
        for (int p = 0; p < M::NPAIR; ++p) {
            float C2[3], nrm[3];
            const float d = self_pair_distance<M, ABLOCK>(sm, p, C2, nrm, 0.f);
            if (d < 0.f) {  // residual -alpha d > 0 (optimization_utils.py:653-660)
                float g[D];
                self_pair_gradient<M, ABLOCK>(sm, p, C2, nrm, g);
                rank1_update<M>(A, b, g, prm.w2_self, d);
            }
        }
    }
    if (prm.use_env) {
        for (int o = 0; o < ob.n; ++o)
            for (int c = 0; c < M::NCAP; ++c) {
                float Cw[3], nrm[3];
                const float d = env_capsule_distance<M, ABLOCK>(sm, c, ob, o, Cw, nrm, 0.f);
                if (d < 0.f) {
                    float g[D];
                    env_capsule_gradient<M, ABLOCK>(sm, c, Cw, nrm, g);
                    rank1_update<M>(A, b, g, prm.w2_env, d);
                }
            }
    }

    if (prm.use_diff) {
        const bool has_prev = t > 0, has_next = t < T - 1;
        const float nt = (has_prev ? 1.f : 0.f) + (has_next ? 1.f : 0.f);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float wn = has_next ? wrap_pi(__ldg(q + (i + 1) * D + d) - x[d]) : 0.f;
            const float wp = has_prev ? wrap_pi(x[d] - __ldg(q + (i - 1) * D + d)) : 0.f;
            b[d] = fmaf(prm.beta[d], wn - wp, b[d]);
            A[tri(d, d)] = fmaf(prm.beta[d], nt, A[tri(d, d)]);
        }
    }
    if (prm.use_virtual && (t < prm.n_virtual || t >= T - prm.n_virtual)) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            A[tri(d, d)] += prm.gamma2;
            if (xv) b[d] = fmaf(-prm.gamma2, wrap_pi(x[d] - __ldg(xv + i * D + d)), b[d]);
        }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) A[tri(d, d)] += prm.lambda;

    float4* out = reinterpret_cast<float4*>(ws + i * NW);
    float blk[NW];
#pragma unroll
    for (int k = 0; k < NT; ++k) blk[k] = A[k];
#pragma unroll
    for (int d = 0; d < D; ++d) blk[NT + d] = b[d];
#pragma unroll
    for (int k = NT + D; k < NW; ++k) blk[k] = 0.f;
#pragma unroll
    for (int k = 0; k < NW / 4; ++k) out[k] = make_float4(blk[4 * k], blk[4 * k + 1], blk[4 * k + 2], blk[4 * k + 3]);
}

struct SolveParams {
    float beta[CPPFLOW_MAX_DOF];
    int do_clamp;
};

template <int NW>
__device__ __forceinline__ void load_block(const float* __restrict__ p, float (&v)[NW]) {
    const float4* s = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int k = 0; k < NW / 4; ++k) {
        const float4 f = s[k];
        v[4 * k] = f.x; v[4 * k + 1] = f.y; v[4 * k + 2] = f.z; v[4 * k + 3] = f.w;
    }
}
template <int NW>
__device__ __forceinline__ void store_block(float* __restrict__ p, const float (&v)[NW]) {
    float4* s = reinterpret_cast<float4*>(p);
#pragma unroll
    for (int k = 0; k < NW / 4; ++k) s[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
template <int NW>
__device__ __forceinline__ void prefetch_block(const float* p) {
    prefetch_l1(p);
    prefetch_l1(p + 32);
    if (NW > 32) prefetch_l1(p + NW - 1);
}
constexpr int SOLVE_PF = 6;  // blocks prefetched ahead of the sweep (per thread, into L1)

// Block-Thomas sweep, one thread per path.  Forward: S_t = A_t - E S_{t-1}^-1 E, y_t = b_t - E u_{t-1} with
// E = -diag(beta), u_t = S_t^-1 y_t; the workspace block is overwritten by (S_t^-1, u_t).
// Backward: dx_t = u_t + S_t^-1 (beta . dx_{t+1}).
template <class M>
__global__ void __launch_bounds__(32)
lm_block_solve_kernel(const float* __restrict__ q, int64_t P, int64_t T, const SolveParams prm, float* __restrict__ ws,
                      float* __restrict__ x_out) {
    constexpr int D = M::NDOF;
    constexpr int NT = BlockLayout<D>::NT;
    constexpr int NW = BlockLayout<D>::NW;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float* w = ws + p * T * NW;
    float Sinv[D][D];
    float u[D];
    float cur[NW], nxt[NW];
    load_block<NW>(w, cur);
    for (int64_t t = 0; t < T; ++t) {
        if (t + SOLVE_PF < T) prefetch_block<NW>(w + (t + SOLVE_PF) * NW);
        if (t + 1 < T) load_block<NW>(w + (t + 1) * NW, nxt);
        float S[D][D], y[D], dinv[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            y[i] = cur[NT + i];
#pragma unroll
            for (int j = 0; j <= i; ++j) S[i][j] = cur[tri(i, j)];
        }
        if (t > 0) {
#pragma unroll
            for (int i = 0; i < D; ++i) {
                y[i] = fmaf(prm.beta[i], u[i], y[i]);
#pragma unroll
                for (int j = 0; j <= i; ++j) S[i][j] = fmaf(-prm.beta[i] * prm.beta[j], Sinv[i][j], S[i][j]);
            }
        }
        chol_lower<D>(S, dinv);
        chol_inverse<D>(S, dinv, Sinv);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < D; ++j) s = fmaf(j <= i ? Sinv[i][j] : Sinv[j][i], y[j], s);
            u[i] = s;
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {
            cur[NT + i] = u[i];
#pragma unroll
            for (int j = 0; j <= i; ++j) cur[tri(i, j)] = Sinv[i][j];
        }
        store_block<NW>(w + t * NW, cur);
#pragma unroll
        for (int k = 0; k < NW; ++k) cur[k] = nxt[k];
    }
    // backward substitution
    float dx[D];
#pragma unroll
    for (int i = 0; i < D; ++i) dx[i] = u[i];
    for (int64_t t = T - 1; t >= 0; --t) {
        if (t >= SOLVE_PF) {
            prefetch_block<NW>(w + (t - SOLVE_PF) * NW);
            prefetch_l1(q + (p * T + t - SOLVE_PF) * D);
        }
        if (t < T - 1) {
            float z[D];
#pragma unroll
            for (int i = 0; i < D; ++i) z[i] = prm.beta[i] * dx[i];
#pragma unroll
            for (int i = 0; i < D; ++i) {
                float s = cur[NT + i];
#pragma unroll
                for (int j = 0; j < D; ++j) s = fmaf(j <= i ? cur[tri(i, j)] : cur[tri(j, i)], z[j], s);
                dx[i] = s;
            }
        }
        if (t > 0) load_block<NW>(w + (t - 1) * NW, cur);  // issued before the stores below: overlaps with them
        float xn[D];
        const float* qp = q + (p * T + t) * D;
#pragma unroll
        for (int i = 0; i < D; ++i) xn[i] = __ldg(qp + i) + dx[i];
        if (prm.do_clamp) {
            static_for<D>([&](auto Dd) {
                constexpr int d = decltype(Dd)::value;
                xn[d] = fminf(fmaxf(xn[d], dof_lower<M>(d)), dof_upper<M>(d));
            });
        }
        float* xo = x_out + (p * T + t) * D;
#pragma unroll
        for (int i = 0; i < D; ++i) xo[i] = xn[i];
    }
}

template <class M>
static void make_params(const cppflow_lm_params* p, int n_obstacles, int do_clamp, AssembleParams& ap, SolveParams& sp) {
    ap = AssembleParams{};
    sp = SolveParams{};
    ap.lambda = p->lm_lambda;
    ap.a_pos = p->alpha_position;
    ap.a_rot = p->alpha_rotation;
    ap.w2_self = p->alpha_self_collision * p->alpha_self_collision;
    ap.w2_env = p->alpha_env_collision * p->alpha_env_collision;
    const float gam = p->alpha_virtual_configs * p->alpha_differencing;
    ap.gamma2 = gam * gam;
    for (int d = 0; d < M::NDOF; ++d) {
        float a = p->alpha_differencing;
        if (dof_is_prismatic<M>(d)) a *= p->alpha_differencing_prismatic_scaling;
        ap.beta[d] = p->use_differencing ? a * a : 0.f;
        sp.beta[d] = ap.beta[d];
    }
    ap.use_pose = p->use_pose;
    ap.use_diff = p->use_differencing;
    ap.use_virtual = p->use_virtual_configs;
    ap.n_virtual = p->n_virtual_configs;
    ap.use_self = p->use_self_collisions;
    ap.use_env = p->use_env_collisions && n_obstacles > 0;
    sp.do_clamp = do_clamp;
}

template <class M>
static int launch_assemble(const cppflow_lm_params* p, const float* q, const float* xv, const float* target, int64_t P,
                           int64_t T, const Obstacles& ob, float* ws, cudaStream_t st) {
    AssembleParams ap;
    SolveParams sp;
    make_params<M>(p, ob.n, 0, ap, sp);
    const size_t sh = sizeof(float) * ABLOCK * SmemLayout<M>::N_FULL;
    static bool attr_set = false;  // per template instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(lm_assemble_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
        if (e != cudaSuccess) return fail(CPPFLOW_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    lm_assemble_kernel<M><<<grid_for(P * T, ABLOCK), ABLOCK, sh, st>>>(q, xv, target, P, T, ob, ap, ws);
    return CPPFLOW_OK;
}

template <class M>
static int launch_solve(const cppflow_lm_params* p, const float* q, int64_t P, int64_t T, int do_clamp, float* ws,
                        float* x_out, cudaStream_t st) {
    AssembleParams ap;
    SolveParams sp;
    make_params<M>(p, 0, do_clamp, ap, sp);
    lm_block_solve_kernel<M><<<grid_for(P, 32), 32, 0, st>>>(q, P, T, sp, ws, x_out);
    return CPPFLOW_OK;
}

template <class M>
static size_t ws_bytes(int64_t P, int64_t T) {
    return (size_t)P * (size_t)T * BlockLayout<M::NDOF>::NW * sizeof(float);
}

}  // namespace cppflow

using namespace cppflow;

extern "C" size_t cppflow_lm_full_workspace_bytes(int robot, int64_t P, int64_t T) {
    if (P < 0 || T < 0) return 0;
    switch (robot) {
        case ROBOT_FETCH: return ws_bytes<Fetch>(P, T);
        case ROBOT_FETCH_ARM: return ws_bytes<FetchArm>(P, T);
        case ROBOT_PANDA: return ws_bytes<Panda>(P, T);
        default: return 0;
    }
}

static int check_common(int robot, const cppflow_lm_params* params, int64_t P, int64_t T, const void* d_workspace,
                        size_t workspace_bytes) {
    CPPFLOW_CHECK_ARG(params != nullptr, "params");
    CPPFLOW_CHECK_ARG(P >= 0 && T >= 0, "P, T");
    CPPFLOW_CHECK_ARG(d_workspace != nullptr, "workspace");
    CPPFLOW_CHECK_ARG(!params->use_virtual_configs || (params->n_virtual_configs > 0 && 2 * params->n_virtual_configs < T),
                      "2 * n_virtual_configs must be < T (optimization_utils.py:457-459)");
    CPPFLOW_CHECK_ARG(((uintptr_t)d_workspace & 15) == 0, "workspace must be 16-byte aligned");
    if (workspace_bytes < cppflow_lm_full_workspace_bytes(robot, P, T))
        return fail(CPPFLOW_E_WORKSPACE, "lm_full: workspace too small (%zu < %zu)", workspace_bytes,
                    cppflow_lm_full_workspace_bytes(robot, P, T));
    return CPPFLOW_OK;
}

extern "C" int cppflow_lm_full_assemble(int robot, const cppflow_lm_params* params, const float* d_q, const float* d_xv,
                                        const float* d_target, int64_t P, int64_t T, const float* h_cuboids,
                                        const float* h_Tcuboids, int n_obstacles, void* d_workspace,
                                        size_t workspace_bytes, void* stream) {
    if (P == 0 || T == 0) return CPPFLOW_OK;
    if (int rc = check_common(robot, params, P, T, d_workspace, workspace_bytes)) return rc;
    CPPFLOW_CHECK_ARG(d_q != nullptr, "null pointer");
    CPPFLOW_CHECK_ARG(!params->use_pose || d_target, "target path required when use_pose");
    Obstacles ob;
    if (int rc = make_obstacles(h_cuboids, h_Tcuboids, n_obstacles, ob)) return rc;
    int rc = CPPFLOW_OK;
    CPPFLOW_DISPATCH_ROBOT(robot, rc = launch_assemble<M>(params, d_q, d_xv, d_target, P, T, ob, (float*)d_workspace,
                                                          (cudaStream_t)stream));
    if (rc) return rc;
    CPPFLOW_CHECK_LAUNCH();
    return CPPFLOW_OK;
}

extern "C" int cppflow_lm_full_solve(int robot, const cppflow_lm_params* params, const float* d_q, int64_t P, int64_t T,
                                     int do_clamp, void* d_workspace, size_t workspace_bytes, float* d_x_out,
                                     void* stream) {
    if (P == 0 || T == 0) return CPPFLOW_OK;
    if (int rc = check_common(robot, params, P, T, d_workspace, workspace_bytes)) return rc;
    CPPFLOW_CHECK_ARG(d_q && d_x_out, "null pointer");
    int rc = CPPFLOW_OK;
    CPPFLOW_DISPATCH_ROBOT(robot, rc = launch_solve<M>(params, d_q, P, T, do_clamp, (float*)d_workspace, d_x_out,
                                                       (cudaStream_t)stream));
    if (rc) return rc;
    CPPFLOW_CHECK_LAUNCH();
    return CPPFLOW_OK;
}

extern "C" int cppflow_lm_full_step(int robot, const cppflow_lm_params* params, const float* d_q, const float* d_xv,
                                    const float* d_target, int64_t P, int64_t T, const float* h_cuboids,
                                    const float* h_Tcuboids, int n_obstacles, int do_clamp, void* d_workspace,
                                    size_t workspace_bytes, float* d_x_out, void* stream) {
    if (int rc = cppflow_lm_full_assemble(robot, params, d_q, d_xv, d_target, P, T, h_cuboids, h_Tcuboids, n_obstacles,
                                          d_workspace, workspace_bytes, stream))
        return rc;
    return cppflow_lm_full_solve(robot, params, d_q, P, T, do_clamp, d_workspace, workspace_bytes, d_x_out, stream);
}
