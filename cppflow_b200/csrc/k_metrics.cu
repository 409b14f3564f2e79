// K5: per-path validity metrics - the tensor part of x_is_valid and the cost the multi-GPU argmin ranks paths by.
//
// Reference being replaced:
//   calculate_pose_error_cm_deg       evaluation_utils.py:113-117 (+ positional_errors / rotational_errors :131-141)
//   angular_changes / prismatic_changes, errors_are_below_threshold   evaluation_utils.py:29-75, :97-98, :144-154
//   calc_TL                           optimization.py:173-175
//   capsule pre-filter for the klampt mesh checks of x_is_valid       optimization_utils.py:889-900
// One CTA per path, threads stride over its waypoints, block reduction at the end.
#include "common.cuh"
#include "collision.cuh"

namespace cppflow {

constexpr int MBLOCK = 128;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <class M>
__global__ void __launch_bounds__(MBLOCK)
path_metrics_kernel(const float* __restrict__ q, const float* __restrict__ target, int64_t T, const Obstacles ob,
                    float* __restrict__ out) {
    constexpr int D = M::NDOF;
    extern __shared__ float smem[];
    __shared__ float red[7][MBLOCK / 32];
    const int64_t p = blockIdx.x;
    float* sm = smem + threadIdx.x;
    float m_pos = 0.f, m_rot = 0.f, m_rev = 0.f, m_pri = 0.f, tl = 0.f, d_self = INFINITY, d_env = INFINITY;
    for (int64_t t = threadIdx.x; t < T; t += MBLOCK) {
        const int64_t i = p * T + t;
        float x[D];
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = __ldg(q + i * D + d);
        CollisionSink<M, MBLOCK, false> sink{sm};
        Frame F;
        fk_chain<M>(x, sink, F);
        // pose error: 100 * |dt| cm; geodesic quaternion distance in degrees (data_types.py:408-411 formula,
        // taken on |dot| so that q and -q are the same rotation)
        const float* tg = target + t * 7;
        const float dx = __ldg(tg) - F.p[0], dy = __ldg(tg + 1) - F.p[1], dz = __ldg(tg + 2) - F.p[2];
        m_pos = fmaxf(m_pos, 100.f * sqrtf(dx * dx + dy * dy + dz * dz));
        float qc[4];
        rotmat_to_quat(F.R, qc);
        float dot = fabsf(qc[0] * __ldg(tg + 3) + qc[1] * __ldg(tg + 4) + qc[2] * __ldg(tg + 5) + qc[3] * __ldg(tg + 6));
        dot = fminf(dot, 1.f - 1e-7f);
        m_rot = fmaxf(m_rot, 2.f * acosf(dot) * 57.29577951308232f);
        if (t > 0) {
            static_for<D>([&](auto Dd) {
                constexpr int d = decltype(Dd)::value;
                const float prev = __ldg(q + (i - 1) * D + d);
                if constexpr (dof_is_prismatic<M>(d)) {
                    m_pri = fmaxf(m_pri, 100.f * fabsf(x[d] - prev));
                } else {
                    const float w = fabsf(wrap_pi(x[d] - prev));
                    m_rev = fmaxf(m_rev, w * 57.29577951308232f);
                    tl += w;
                }
            });
        }
        for (int pr = 0; pr < M::NPAIR; ++pr) {
            float C2[3], nrm[3];
            d_self = fminf(d_self, self_pair_distance<M, MBLOCK>(sm, pr, C2, nrm, d_self));
        }
        for (int o = 0; o < ob.n; ++o)
            for (int c = 0; c < M::NCAP; ++c) {
                float Cw[3], nrm[3];
                d_env = fminf(d_env, env_capsule_distance<M, MBLOCK>(sm, c, ob, o, Cw, nrm, d_env));
            }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float v[7] = {warp_max(m_pos), warp_max(m_rot), warp_max(m_rev), warp_max(m_pri), warp_sum(tl), warp_min(d_self),
                  warp_min(d_env)};
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) red[k][warp] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float r[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) r[k] = red[k][0];
        for (int w = 1; w < MBLOCK / 32; ++w) {
            r[0] = fmaxf(r[0], red[0][w]); r[1] = fmaxf(r[1], red[1][w]);
            r[2] = fmaxf(r[2], red[2][w]); r[3] = fmaxf(r[3], red[3][w]);
            r[4] += red[4][w];
            r[5] = fminf(r[5], red[5][w]); r[6] = fminf(r[6], red[6][w]);
        }
        float* o = out + p * 8;
#pragma unroll
        for (int k = 0; k < 7; ++k) o[k] = r[k];
        o[7] = 0.f;
    }
}

}  // namespace cppflow

using namespace cppflow;

extern "C" int cppflow_path_metrics(int robot, const float* d_q, const float* d_target, int64_t P, int64_t T,
                                    const float* h_cuboids, const float* h_Tcuboids, int n_obstacles, float* d_out,
                                    void* stream) {
    CPPFLOW_CHECK_ARG(P >= 0 && T > 0, "P, T");
    if (P == 0) return CPPFLOW_OK;
    CPPFLOW_CHECK_ARG(d_q && d_target && d_out, "null pointer");
    Obstacles ob;
    if (int rc = make_obstacles(h_cuboids, h_Tcuboids, n_obstacles, ob)) return rc;
    CPPFLOW_DISPATCH_ROBOT(robot, {
        const size_t sh = sizeof(float) * MBLOCK * SmemLayout<M>::N_DIST;
        path_metrics_kernel<M><<<(unsigned)P, MBLOCK, sh, (cudaStream_t)stream>>>(d_q, d_target, T, ob, d_out);
    });
    CPPFLOW_CHECK_LAUNCH();
    return CPPFLOW_OK;
}
