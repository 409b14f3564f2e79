// K2: capsule-capsule self-collision and capsule-cuboid environment distances, their joint-space Jacobians, and
// the per-configuration collision flags.  One thread per configuration; all-link FK streams capsule endpoints into
// a per-thread shared-memory column, then the pair / capsule loops read them back.
//
// Reference being replaced:
//   jrl Robot.self_collision_distances[_jacobian], Robot.env_collision_distances[_jacobian]
//       (collision_detection.py:40,65; optimization_utils.py:652,670,690,710)
//   qpaths_batched_self_collisions / qpaths_batched_env_collisions   (collision_detection.py:27-69)
#include "common.cuh"
#include "collision.cuh"

namespace cppflow {

constexpr int CBLOCK = 128;

template <class M>
__device__ __forceinline__ void load_q_plain(const float* __restrict__ q, int64_t i, float (&x)[M::NDOF]) {
#pragma unroll
    for (int d = 0; d < M::NDOF; ++d) x[d] = __ldg(q + i * M::NDOF + d);
}

template <class M, bool WITH_J>
__global__ void __launch_bounds__(CBLOCK)
self_dist_kernel(const float* __restrict__ q, int64_t n, float* __restrict__ dist, float* __restrict__ Jout) {
    extern __shared__ float smem[];
    const int64_t i = (int64_t)blockIdx.x * CBLOCK + threadIdx.x;
    if (i >= n) return;
    float* sm = smem + threadIdx.x;
    float x[M::NDOF];
    load_q_plain<M>(q, i, x);
    CollisionSink<M, CBLOCK, WITH_J> sink{sm};
    Frame F;
    fk_chain<M>(x, sink, F);
    for (int p = 0; p < M::NPAIR; ++p) {
        float C2[3], nrm[3];
        const float d = self_pair_distance<M, CBLOCK>(sm, p, C2, nrm);
        dist[i * M::NPAIR + p] = d;
        if constexpr (WITH_J) {
            float g[M::NDOF];
            self_pair_gradient<M, CBLOCK>(sm, p, C2, nrm, g);
#pragma unroll
            for (int k = 0; k < M::NDOF; ++k) Jout[(i * M::NPAIR + p) * M::NDOF + k] = g[k];
        }
    }
}

template <class M, bool WITH_J>
__global__ void __launch_bounds__(CBLOCK)
env_dist_kernel(const float* __restrict__ q, int64_t n, const Obstacles ob, float* __restrict__ dist,
                float* __restrict__ Jout) {
    extern __shared__ float smem[];
    const int64_t i = (int64_t)blockIdx.x * CBLOCK + threadIdx.x;
    if (i >= n) return;
    float* sm = smem + threadIdx.x;
    float x[M::NDOF];
    load_q_plain<M>(q, i, x);
    CollisionSink<M, CBLOCK, WITH_J> sink{sm};
    Frame F;
    fk_chain<M>(x, sink, F);
    for (int c = 0; c < M::NCAP; ++c) {
        float Cw[3], nrm[3];
        const float d = env_capsule_distance<M, CBLOCK>(sm, c, ob, 0, Cw, nrm);
        dist[i * M::NCAP + c] = d;
        if constexpr (WITH_J) {
            float g[M::NDOF];
            env_capsule_gradient<M, CBLOCK>(sm, c, Cw, nrm, g);
#pragma unroll
            for (int k = 0; k < M::NDOF; ++k) Jout[(i * M::NCAP + c) * M::NDOF + k] = g[k];
        }
    }
}

template <class M>
__global__ void __launch_bounds__(CBLOCK)
collision_flags_kernel(const float* __restrict__ q, int64_t n, const Obstacles ob, uint8_t* __restrict__ self_flags,
                       uint8_t* __restrict__ env_flags) {
    extern __shared__ float smem[];
    __shared__ CollTables<M> tb;
    fill_coll_tables<M>(tb, ob, threadIdx.x, CBLOCK);
    const int64_t i_raw = (int64_t)blockIdx.x * CBLOCK + threadIdx.x;
    const bool live = i_raw < n;
    const int64_t i = live ? i_raw : n - 1;
    float* sm = smem + threadIdx.x;
    float x[M::NDOF];
    load_q_plain<M>(q, i, x);
    __syncthreads();
    MidSink<M, CBLOCK, false> sink;
    sink.sm = sm;
    Frame F;
    fk_chain<M>(x, sink, F);
    // only the sign of the minimum matters: bounding-sphere culls in registers, exact distance for the survivors
    if (self_flags) {
        unsigned mask = self_cull_mask<M>(sink.mid2);
        bool hit = false;
        while (mask) {
            const int p = __ffs(mask) - 1;
            mask &= mask - 1;
            float C2[3], nrm[3];
            if (self_pair_exact<M, CBLOCK>(sm, tb, p, C2, nrm) < 0.f) {
                hit = true;
                mask = 0u;
            }
        }
        if (live) self_flags[i] = hit ? 1 : 0;
    }
    if (env_flags) {
        bool hit = false;
        for (int o = 0; o < tb.ob.n; ++o) {
            unsigned mask = hit ? 0u : env_cull_mask<M>(sink.mid2, tb.ob, o);
            while (mask) {
                const int c = __ffs(mask) - 1;
                mask &= mask - 1;
                float Cw[3], nrm[3];
                if (env_capsule_exact<M, CBLOCK>(sm, tb, c, o, Cw, nrm) < 0.f) {
                    hit = true;
                    mask = 0u;
                }
            }
        }
        if (live) env_flags[i] = hit ? 1 : 0;
    }
}

template <class K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return fail(CPPFLOW_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    return CPPFLOW_OK;
}

template <class M>
static int launch_self(const float* q, int64_t n, float* dist, float* J, cudaStream_t st) {
    if (J) {
        const size_t sh = sizeof(float) * CBLOCK * SmemLayout<M>::N_FULL;
        if (int rc = set_smem(self_dist_kernel<M, true>, sh)) return rc;
        self_dist_kernel<M, true><<<grid_for(n, CBLOCK), CBLOCK, sh, st>>>(q, n, dist, J);
    } else {
        const size_t sh = sizeof(float) * CBLOCK * SmemLayout<M>::N_DIST;
        self_dist_kernel<M, false><<<grid_for(n, CBLOCK), CBLOCK, sh, st>>>(q, n, dist, nullptr);
    }
    return CPPFLOW_OK;
}

template <class M>
static int launch_env(const float* q, int64_t n, const Obstacles& ob, float* dist, float* J, cudaStream_t st) {
    if (J) {
        const size_t sh = sizeof(float) * CBLOCK * SmemLayout<M>::N_FULL;
        if (int rc = set_smem(env_dist_kernel<M, true>, sh)) return rc;
        env_dist_kernel<M, true><<<grid_for(n, CBLOCK), CBLOCK, sh, st>>>(q, n, ob, dist, J);
    } else {
        const size_t sh = sizeof(float) * CBLOCK * SmemLayout<M>::N_DIST;
        env_dist_kernel<M, false><<<grid_for(n, CBLOCK), CBLOCK, sh, st>>>(q, n, ob, dist, nullptr);
    }
    return CPPFLOW_OK;
}

}  // namespace cppflow

using namespace cppflow;

extern "C" int cppflow_self_collision_distances(int robot, const float* d_q, int64_t n, float* d_dist, float* d_J,
                                                void* stream) {
    CPPFLOW_CHECK_ARG(n >= 0, "n");
    if (n == 0) return CPPFLOW_OK;
    CPPFLOW_CHECK_ARG(d_q && d_dist, "null pointer");
    int rc = CPPFLOW_OK;
    CPPFLOW_DISPATCH_ROBOT(robot, rc = launch_self<M>(d_q, n, d_dist, d_J, (cudaStream_t)stream));
    if (rc) return rc;
    CPPFLOW_CHECK_LAUNCH();
    return CPPFLOW_OK;
}

extern "C" int cppflow_env_collision_distances(int robot, const float* d_q, int64_t n, const float* h_cuboid,
                                               const float* h_Tcuboid, float* d_dist, float* d_J, void* stream) {
    CPPFLOW_CHECK_ARG(n >= 0, "n");
    if (n == 0) return CPPFLOW_OK;
    CPPFLOW_CHECK_ARG(d_q && d_dist && h_cuboid && h_Tcuboid, "null pointer");
    Obstacles ob;
    if (int rc = make_obstacles(h_cuboid, h_Tcuboid, 1, ob)) return rc;
    int rc = CPPFLOW_OK;
    CPPFLOW_DISPATCH_ROBOT(robot, rc = launch_env<M>(d_q, n, ob, d_dist, d_J, (cudaStream_t)stream));
    if (rc) return rc;
    CPPFLOW_CHECK_LAUNCH();
    return CPPFLOW_OK;
}

extern "C" int cppflow_collision_flags(int robot, const float* d_q, int64_t n, const float* h_cuboids,
                                       const float* h_Tcuboids, int n_obstacles, uint8_t* d_self_flags,
                                       uint8_t* d_env_flags, void* stream) {
    CPPFLOW_CHECK_ARG(n >= 0, "n");
    if (n == 0) return CPPFLOW_OK;
    CPPFLOW_CHECK_ARG(d_q != nullptr, "null pointer");
    Obstacles ob;
    if (int rc = make_obstacles(h_cuboids, h_Tcuboids, n_obstacles, ob)) return rc;
    CPPFLOW_DISPATCH_ROBOT(robot, {
        const size_t sh = sizeof(float) * CBLOCK * SmemLayout<M>::N_DIST;
        collision_flags_kernel<M><<<grid_for(n, CBLOCK), CBLOCK, sh, (cudaStream_t)stream>>>(d_q, n, ob, d_self_flags,
                                                                                             d_env_flags);
    });
    CPPFLOW_CHECK_LAUNCH();
    return CPPFLOW_OK;
}
