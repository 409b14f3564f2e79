// K4: dp_search - the bottleneck dynamic program over k candidate joint-space paths.
//
// Reference being replaced:
//   joint_limit_almost_violations_3d   search.py:25-52
//   _get_mjacs                         search.py:100-125   (materialises [k,k,T-1,D]; here [T-1,k,k] only)
//   dp_search                          search.py:128-173   (T-1 dependent k x k steps + backtrack)
// Bit-exactness contract: memo / costs / best_path equal the reference's fp32 results exactly.  Every
// arithmetic op is an explicit round-to-nearest fp32 op in the reference's order (no FMA contraction):
//   dq = q[i,t+1,d] - q[j,t,d];  dq *= 5 for prismatic d (before the wrap);  w = |remainder(dq + pi, 2 pi) - pi|
//   mjac = max_d w;  c = max(mjac, cost[j,t-1]) + ext[i,t];  cost[i,t], memo[i,t] = min_j / first argmin_j.
#include "common.cuh"
#include "kinematics.cuh"

namespace cppflow {

template <class M>
__device__ __forceinline__ bool near_joint_limit(const float* __restrict__ q, float eps_rev, float eps_pris) {
    bool bad = false;
    static_for<M::NDOF>([&](auto Dd) {
        constexpr int d = decltype(Dd)::value;
        const float eps = dof_is_prismatic<M>(d) ? eps_pris : eps_rev;
        const float lo = __fadd_rn(dof_lower<M>(d), eps);  // l_lim[idx] += eps in fp32 (search.py:48-49)
        const float hi = __fsub_rn(dof_upper<M>(d), eps);
        const float v = q[d];
        bad = bad || (v < lo) || (v > hi);
    });
    return bad;
}

template <class M>
__global__ void joint_limit_flags_kernel(const float* __restrict__ q, int64_t n, float eps_rev, float eps_pris,
                                         float* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = near_joint_limit<M>(q + i * M::NDOF, eps_rev, eps_pris) ? 1.f : 0.f;
}

// ext[t][i] = 100 * jlim + 1000 * env + 1000 * self   (search.py:146-150), stored t-major for the DP sweep
template <class M>
__global__ void dp_ext_kernel(const float* __restrict__ q, const uint8_t* __restrict__ self_f,
                              const uint8_t* __restrict__ env_f, int k, int T, float eps_rev, float eps_pris,
                              float* __restrict__ ext) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)k * T) return;
    const int i = (int)(idx / T), t = (int)(idx % T);
    const float jl = near_joint_limit<M>(q + idx * M::NDOF, eps_rev, eps_pris) ? 1.f : 0.f;
    float c = __fmul_rn(100.f, jl);
    c = __fadd_rn(c, env_f[idx] ? 1000.f : 0.f);
    c = __fadd_rn(c, self_f[idx] ? 1000.f : 0.f);
    ext[(int64_t)t * k + i] = c;
}

// mj[t][i][j] = max_d |wrap(s_d (q[i,t+1,d] - q[j,t,d]))|, t in [0, T-1)
constexpr int MJ_ROWS = 16;
template <class M>
__global__ void __launch_bounds__(256)
dp_mjac_kernel(const float* __restrict__ q, int k, int T, float* __restrict__ mj) {
    constexpr int D = M::NDOF;
    __shared__ float qto[MJ_ROWS][D];
    const int t = blockIdx.x;
    const int i0 = blockIdx.y * MJ_ROWS;
    const int rows = min(MJ_ROWS, k - i0);
    for (int e = threadIdx.x; e < rows * D; e += blockDim.x) {
        const int r = e / D, d = e % D;
        qto[r][d] = q[((int64_t)(i0 + r) * T + (t + 1)) * D + d];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        float qf[D];
#pragma unroll
        for (int d = 0; d < D; ++d) qf[d] = q[((int64_t)j * T + t) * D + d];
        for (int r = 0; r < rows; ++r) {
            float m = 0.f;
            static_for<D>([&](auto Dd) {
                constexpr int d = decltype(Dd)::value;
                float dq = __fsub_rn(qto[r][d], qf[d]);
                if constexpr (dof_is_prismatic<M>(d)) dq = __fmul_rn(dq, 5.0f);
                m = fmaxf(m, fabsf(wrap_pi(dq)));
            });
            mj[((int64_t)t * k + (i0 + r)) * k + j] = m;
        }
    }
}

// Sequential sweep over t.  One CTA; warp w owns rows i = w, w + nwarps, ...; lanes stride over j and keep the
// first minimum; the cross-lane reduction orders (value, index) lexicographically so ties pick the lowest j, as
// torch.min does.  memo is also kept in shared memory (uint16) when it fits so the backtrack never leaves the SM.
template <bool MEMO_SMEM>
__global__ void __launch_bounds__(1024)
dp_sweep_kernel(const float* __restrict__ q, const float* __restrict__ ext, const float* __restrict__ mj, int k, int T,
                int D, float* __restrict__ costs, int32_t* __restrict__ memo, int32_t* __restrict__ chosen,
                float* __restrict__ best_path) {
    extern __shared__ unsigned char dsm[];
    float* cost_a = reinterpret_cast<float*>(dsm);
    float* cost_b = cost_a + k;
    uint16_t* memo_s = reinterpret_cast<uint16_t*>(cost_b + k);
    __shared__ int s_best;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const float c = ext[i];
        cost_a[i] = c;
        costs[(int64_t)i * T] = c;
        memo[(int64_t)i * T] = 0;
        if (MEMO_SMEM) memo_s[i] = 0;
    }
    __syncthreads();
    float* prev = cost_a;
    float* next = cost_b;
    for (int t = 1; t < T; ++t) {
        const float* mjt = mj + (int64_t)(t - 1) * k * k;
        const float* ext_t = ext + (int64_t)t * k;
        for (int i = warp; i < k; i += nwarps) {
            const float* row = mjt + (int64_t)i * k;
            const float e = ext_t[i];
            float best = INFINITY;
            int bj = 0x7fffffff;
            for (int j = lane; j < k; j += 32) {
                const float v = __fadd_rn(fmaxf(row[j], prev[j]), e);
                if (v < best) { best = v; bj = j; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ov = __shfl_down_sync(0xffffffffu, best, off);
                const int oj = __shfl_down_sync(0xffffffffu, bj, off);
                if (ov < best || (ov == best && oj < bj)) { best = ov; bj = oj; }
            }
            if (lane == 0) {
                next[i] = best;
                costs[(int64_t)i * T + t] = best;
                memo[(int64_t)i * T + t] = bj;
                if (MEMO_SMEM) memo_s[(int64_t)t * k + i] = (uint16_t)bj;
            }
        }
        __syncthreads();
        float* tmp = prev; prev = next; next = tmp;
    }
    // final argmin over the last column (first index on ties), then backtrack (search.py:162-173)
    if (warp == 0) {
        float best = INFINITY;
        int bi = 0x7fffffff;
        for (int i = lane; i < k; i += 32) {
            const float v = prev[i];
            if (v < best) { best = v; bi = i; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_down_sync(0xffffffffu, best, off);
            const int oi = __shfl_down_sync(0xffffffffu, bi, off);
            if (ov < best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) {
            int i = bi;
            for (int t = T - 1; t >= 0; --t) {
                chosen[t] = i;
                i = MEMO_SMEM ? (int)memo_s[(int64_t)t * k + i] : memo[(int64_t)i * T + t];
            }
        }
    }
    __threadfence_block();
    __syncthreads();
    for (int e = threadIdx.x; e < T * D; e += blockDim.x) {
        const int t = e / D, d = e % D;
        best_path[e] = q[((int64_t)chosen[t] * T + t) * D + d];
    }
    (void)s_best;
}

inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace cppflow

using namespace cppflow;

extern "C" int cppflow_joint_limit_flags(int robot, const float* d_q, int64_t n, float eps_revolute,
                                         float eps_prismatic, float* d_flags, void* stream) {
    CPPFLOW_CHECK_ARG(n >= 0, "n");
    if (n == 0) return CPPFLOW_OK;
    CPPFLOW_CHECK_ARG(d_q && d_flags, "null pointer");
    CPPFLOW_DISPATCH_ROBOT(robot, joint_limit_flags_kernel<M><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
                                      d_q, n, eps_revolute, eps_prismatic, d_flags));
    CPPFLOW_CHECK_LAUNCH();
    return CPPFLOW_OK;
}

extern "C" size_t cppflow_dp_search_workspace_bytes(int64_t k, int64_t T) {
    if (k <= 0 || T <= 0) return 0;
    return align256((size_t)T * k * sizeof(float)) + align256((size_t)(T > 1 ? T - 1 : 0) * k * k * sizeof(float));
}

extern "C" int cppflow_dp_search(int robot, const float* d_q, const uint8_t* d_self_flags, const uint8_t* d_env_flags,
                                 int64_t k, int64_t T, void* d_workspace, size_t workspace_bytes, float* d_best_path,
                                 int32_t* d_memo, float* d_costs, int32_t* d_chosen, void* stream) {
    CPPFLOW_CHECK_ARG(k > 0 && T > 0, "k, T must be positive");
    CPPFLOW_CHECK_ARG(k <= 65535 && T <= 65535, "k, T must be <= 65535");
    CPPFLOW_CHECK_ARG(d_q && d_self_flags && d_env_flags && d_workspace && d_best_path && d_memo && d_costs && d_chosen,
                      "null pointer");
    CPPFLOW_CHECK_ARG(((uintptr_t)d_workspace & 255) == 0, "workspace must be 256-byte aligned");
    if (workspace_bytes < cppflow_dp_search_workspace_bytes(k, T))
        return fail(CPPFLOW_E_WORKSPACE, "cppflow_dp_search: workspace too small (%zu < %zu)", workspace_bytes,
                    cppflow_dp_search_workspace_bytes(k, T));
    cudaStream_t st = (cudaStream_t)stream;
    float* ext = (float*)d_workspace;
    float* mj = (float*)((char*)d_workspace + align256((size_t)T * k * sizeof(float)));
    // DEFAULT_JLIM_SAFETY_PADDING_REVOLUTE = deg2rad(1.5), _PRISMATIC = 3 cm (search.py:20-21)
    const float eps_rev = (float)(1.5 * 3.14159265358979323846 / 180.0);
    const float eps_pris = 0.03f;
    int D = 0;
    CPPFLOW_DISPATCH_ROBOT(robot, {
        D = M::NDOF;
        dp_ext_kernel<M><<<grid_for(k * T, 256), 256, 0, st>>>(d_q, d_self_flags, d_env_flags, (int)k, (int)T, eps_rev,
                                                              eps_pris, ext);
        if (T > 1) {
            dim3 grid((unsigned)(T - 1), (unsigned)((k + MJ_ROWS - 1) / MJ_ROWS));
            dp_mjac_kernel<M><<<grid, 256, 0, st>>>(d_q, (int)k, (int)T, mj);
        }
    });
    CPPFLOW_CHECK_LAUNCH();
    const size_t sh_cost = 2 * (size_t)k * sizeof(float);
    const size_t sh_memo = (size_t)T * k * sizeof(uint16_t);
    const bool memo_smem = sh_cost + sh_memo <= 200 * 1024;
    const size_t sh = sh_cost + (memo_smem ? sh_memo : 0);
    const int threads = k >= 512 ? 1024 : (k >= 128 ? 512 : 256);
    cudaError_t e;
    if (memo_smem) {
        e = cudaFuncSetAttribute(dp_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
        if (e != cudaSuccess) return fail(CPPFLOW_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        dp_sweep_kernel<true><<<1, threads, sh, st>>>(d_q, ext, mj, (int)k, (int)T, D, d_costs, d_memo, d_chosen,
                                                      d_best_path);
    } else {
        dp_sweep_kernel<false><<<1, threads, sh, st>>>(d_q, ext, mj, (int)k, (int)T, D, d_costs, d_memo, d_chosen,
                                                       d_best_path);
    }
    CPPFLOW_CHECK_LAUNCH();
    return CPPFLOW_OK;
}
