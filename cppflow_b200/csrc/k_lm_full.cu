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

// one q row -> registers; 16-byte vector loads when the row is 16-byte aligned (D % 4 == 0)
template <int D>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&x)[D]) {
    if constexpr (D % 4 == 0) {
#pragma unroll
        for (int d = 0; d < D; d += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p + d));
            x[d] = v.x; x[d + 1] = v.y; x[d + 2] = v.z; x[d + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = __ldg(p + d);
    }
}

// wrap to [-pi, pi) for the LM residuals (evaluation_utils.py:144-154): one rounding step instead of fmodf.  Within
// 1 ulp of torch.remainder for |d| < 4 pi (joint differences are < 2 pi); the bit-exact wrap_pi stays in dp_search.
__device__ __forceinline__ float wrap_pi_lm(float d) {
    const float k = floorf(fmaf(d, 0.15915494309189535f, 0.5f));
    return fmaf(k, -6.28318548202514648f, d);  // float32(2 pi), the modulus torch uses
}

template <class M>
__global__ void __launch_bounds__(ABLOCK, 4)
lm_assemble_kernel(const float* __restrict__ q, const float* __restrict__ xv, const float* __restrict__ target,
                   int P, int T, const Obstacles ob, const AssembleParams prm, float* __restrict__ ws) {
    constexpr int D = M::NDOF;
    constexpr int NT = BlockLayout<D>::NT;
    constexpr int NW = BlockLayout<D>::NW;
    extern __shared__ float smem[];
    __shared__ CollTables<M> tb;
    fill_coll_tables<M>(tb, ob, threadIdx.x, ABLOCK);
    // grid = (path blocks, waypoints): consecutive threads take consecutive PATHS at the same waypoint t, so the
    // target pose is block-uniform and the block stores below are fully coalesced in the [t][k][path] workspace
    const int t = blockIdx.y;
    const int pth_raw = blockIdx.x * ABLOCK + threadIdx.x;
    const bool live = pth_raw < P;
    const int pth = live ? pth_raw : P - 1;  // idle lanes shadow the last path and skip the stores
    const int64_t i = (int64_t)pth * T + t;  // row of q / xv
    float* sm = smem + threadIdx.x;

    // all global loads up front (the neighbour rows are strided across paths, like the row itself); the wrapped
    // differences are formed right away so that only D of the 3 D neighbour values stay live through the FK
    float x[D], dwrap[D], vwrap[D];
    const bool has_prev = t > 0, has_next = t < T - 1;
    const bool in_virtual = prm.use_virtual && (t < prm.n_virtual || t >= T - prm.n_virtual);
    load_row<D>(q + i * D, x);
    float tg[7];
    if (prm.use_pose) {
#pragma unroll
        for (int k = 0; k < 7; ++k) tg[k] = __ldg(target + (int64_t)t * 7 + k);
    }
    if (prm.use_diff) {
        float xp[D], xn[D];
        load_row<D>(q + (has_prev ? i - 1 : i) * D, xp);
        load_row<D>(q + (has_next ? i + 1 : i) * D, xn);
#pragma unroll
        for (int d = 0; d < D; ++d) dwrap[d] = wrap_pi_lm(xn[d] - x[d]) - wrap_pi_lm(x[d] - xp[d]);  // 0 at the ends
    }
    if (in_virtual && xv) {
        float xvv[D];
        load_row<D>(xv + i * D, xvv);
#pragma unroll
        for (int d = 0; d < D; ++d) vwrap[d] = wrap_pi_lm(x[d] - xvv[d]);
    }
    __syncthreads();  // collision tables visible

    MidSink<M, ABLOCK, true> sink;
    sink.sm = sm;
    Frame F;
    fk_chain<M>(x, sink, F);

    float A[NT], b[D];
#pragma unroll
    for (int k = 0; k < NT; ++k) A[k] = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) b[d] = 0.f;

    // ---- capsule collisions: register culls -> survivor masks -> exact distance; d < 0 adds w^2 g g^T, -w^2 d g
    if (prm.use_self) {
        unsigned mask = self_cull_mask<M>(sink.mid2);
        while (mask) {
            const int p = __ffs(mask) - 1;
            mask &= mask - 1;
            float C2[3], nrm[3];
            const float d = self_pair_exact<M, ABLOCK>(sm, tb, p, C2, nrm);
            if (d < 0.f) {  // residual -alpha d > 0 (optimization_utils.py:653-660)
                float g[D];
                self_pair_gradient_rt<M, ABLOCK>(sm, tb, p, C2, nrm, g);
                rank1_update<M>(A, b, g, prm.w2_self, d);
            }
        }
    }
    if (prm.use_env) {
        // survivors of up to three obstacles share one 32-bit mask (bit = oo * NCAP + c) so that the lanes of a warp
        // walk lists of similar length
        constexpr int OGRP = 32 / M::NCAP;
        for (int o0 = 0; o0 < tb.ob.n; o0 += OGRP) {
            unsigned mask = 0u;
#pragma unroll
            for (int oo = 0; oo < OGRP; ++oo)
                if (o0 + oo < tb.ob.n) mask |= env_cull_mask<M>(sink.mid2, tb.ob, o0 + oo) << (oo * M::NCAP);
            while (mask) {
                const int bit = __ffs(mask) - 1;
                mask &= mask - 1;
                const int oo = bit / M::NCAP, c = bit - oo * M::NCAP;
                float Cw[3], nrm[3];
                const float d = env_capsule_exact<M, ABLOCK>(sm, tb, c, o0 + oo, Cw, nrm);
                if (d < 0.f) {
                    float g[D];
                    env_capsule_gradient_rt<M, ABLOCK>(sm, tb, c, Cw, nrm, g);
                    rank1_update<M>(A, b, g, prm.w2_env, d);
                }
            }
        }
    }

    // ---- pose term: Jp^T Jp, Jp^T e with the alpha-scaled geometric Jacobian (optimization.py:77-87)
    if (prm.use_pose) {
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
        static_for<D>([&](auto Cc) {
            constexpr int c = decltype(Cc)::value;
            constexpr int r0c = dof_is_prismatic<M>(c) ? 3 : 0;  // rows 0-2 of a prismatic column are zero
            float s = b[c];
#pragma unroll
            for (int r = r0c; r < 6; ++r) s = fmaf(J[r][c], e[r], s);
            b[c] = s;
            static_for<c + 1>([&](auto C2c) {
                constexpr int c2 = decltype(C2c)::value;
                constexpr int r0 = (dof_is_prismatic<M>(c) || dof_is_prismatic<M>(c2)) ? 3 : 0;
                float v = A[tri(c, c2)];
#pragma unroll
                for (int r = r0; r < 6; ++r) v = fmaf(J[r][c], J[r][c2], v);
                A[tri(c, c2)] = v;
            });
        });
    }

    if (prm.use_diff) {
        const float nt = (has_prev ? 1.f : 0.f) + (has_next ? 1.f : 0.f);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            b[d] = fmaf(prm.beta[d], dwrap[d], b[d]);
            A[tri(d, d)] = fmaf(prm.beta[d], nt, A[tri(d, d)]);
        }
    }
    if (in_virtual) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            A[tri(d, d)] += prm.gamma2;
            if (xv) b[d] = fmaf(-prm.gamma2, vwrap[d], b[d]);
        }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) A[tri(d, d)] += prm.lambda;

    if (!live) return;
    float4* out = reinterpret_cast<float4*>(ws) + ((int64_t)t * (NW / 4)) * P + pth;  // float4 k at out[k * P]
    float blk[NW];
#pragma unroll
    for (int k = 0; k < NT; ++k) blk[k] = A[k];
#pragma unroll
    for (int d = 0; d < D; ++d) blk[NT + d] = b[d];
#pragma unroll
    for (int k = NT + D; k < NW; ++k) blk[k] = 0.f;
#pragma unroll
    for (int k = 0; k < NW / 4; ++k) out[(int64_t)k * P] = make_float4(blk[4 * k], blk[4 * k + 1], blk[4 * k + 2], blk[4 * k + 3]);
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

// ----------------------------------------------------------------------------------------------------------------
// Block-tridiagonal solve.  Two lanes per path run a TWISTED (two-sided) block-Thomas factorisation: lane side 0
// eliminates t = 0, 1, ... upwards, lane side 1 eliminates t = T-1, T-2, ... downwards; they meet at the middle block
// m = T/2, whose Schur complement takes both neighbours' inverses.  This halves the sequential chain (the kernel is
// bound by the latency of T dependent 8x8 factorisations, not by FLOPs or bandwidth).
//   elimination (per side, "in" = the neighbour already eliminated):
//       S_t = A_t - E S_in^-1 E,  y_t = b_t - E u_in,  u_t = S_t^-1 y_t,  E = -diag(beta);  block t <- (S_t^-1, u_t)
//   middle:  S_m = A_m - E (S_{m-1}^-1 + S_{m+1}^-1) E,  y_m = b_m - E (u_{m-1} + u_{m+1}),  dx_m = S_m^-1 y_m
//   back-substitution (per side, outwards from m):  dx_t = u_t + S_t^-1 (beta . dx_{inner neighbour})
// Blocks are streamed through a per-thread ring in shared memory with cp.async (LDGSTS): the back-substitution step
// is only ~64 FMAs, so without a deep asynchronous ring every step would expose a full DRAM round trip.
constexpr int SOLVE_RING = 8;      // ring slots per thread
constexpr int SOLVE_RING_FWD = 3;  // slots in flight during the elimination sweep (each step is ~1k cycles)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int D>
struct SolveSmem {
    static constexpr int NV = BlockLayout<D>::NW / 4;                 // float4 per block
    static constexpr int SLOT_BYTES = (NV * 16 + ((D + 3) / 4) * 16) * 32;  // one ring slot for the 32 lanes of the warp
    static constexpr int BYTES = SLOT_BYTES * SOLVE_RING;
    // float4 k of the block in `slot` for `lane`; q value d
    __device__ static float4* blk(unsigned char* base, int slot, int k, int lane) {
        return reinterpret_cast<float4*>(base + (size_t)slot * SLOT_BYTES) + k * 32 + lane;
    }
    // q value d of `lane`: groups of 4 dofs are contiguous per lane so a 16-byte cp.async can fill them
    __device__ static float* qv(unsigned char* base, int slot, int d, int lane) {
        return reinterpret_cast<float*>(base + (size_t)slot * SLOT_BYTES + NV * 16 * 32) + ((d >> 2) * 32 + lane) * 4 + (d & 3);
    }
};

template <class M>
__global__ void __launch_bounds__(32)
lm_block_solve_kernel(const float* __restrict__ q, int64_t P, int64_t T, const SolveParams prm, float* __restrict__ ws,
                      float* __restrict__ x_out) {
    constexpr int D = M::NDOF;
    constexpr int NT = BlockLayout<D>::NT;
    constexpr int NW = BlockLayout<D>::NW;
    constexpr int NV = NW / 4;
    using SM = SolveSmem<D>;
    extern __shared__ __align__(16) unsigned char ring[];
    const int lane = threadIdx.x;
    const int side = lane & 1;
    const int64_t p_raw = (int64_t)blockIdx.x * 16 + (lane >> 1);
    const bool active = p_raw < P;
    const int64_t p = active ? p_raw : P - 1;  // idle lanes shadow the last path (no stores) so the warp stays converged
    float4* w4 = reinterpret_cast<float4*>(ws) + p;  // float4 k of block t at w4[(t * NV + k) * P]
    const float* qp = q + p * T * D;
    const int64_t m = T / 2;
    const int64_t n_side = side == 0 ? m : T - 1 - m;  // blocks this lane eliminates
    const int64_t n_iter = m > T - 1 - m ? m : T - 1 - m;

    auto t_of = [&](int64_t k) { return side == 0 ? k : T - 1 - k; };
    auto issue_block = [&](int slot, int64_t t, bool with_q) {
        const float4* src = w4 + (t * NV) * P;
#pragma unroll
        for (int k = 0; k < NV; ++k) cp_async16(SM::blk(ring, slot, k, lane), src + k * P);
        if (with_q) {
            if constexpr (D % 4 == 0) {
#pragma unroll
                for (int d = 0; d < D; d += 4) cp_async16(SM::qv(ring, slot, d, lane), qp + t * D + d);
            } else {
#pragma unroll
                for (int d = 0; d < D; ++d) cp_async4(SM::qv(ring, slot, d, lane), qp + t * D + d);
            }
        }
    };
    auto store_blk = [&](int64_t t, const float (&v)[NW]) {
        float4* dst = w4 + (t * NV) * P;
#pragma unroll
        for (int k = 0; k < NV; ++k) dst[k * P] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    };
    auto store_x = [&](int64_t t, const float (&xn)[D]) {
        float* xo = x_out + (p * T + t) * D;
        if constexpr (D % 4 == 0) {
#pragma unroll
            for (int d = 0; d < D; d += 4)
                *reinterpret_cast<float4*>(xo + d) = make_float4(xn[d], xn[d + 1], xn[d + 2], xn[d + 3]);
        } else {
#pragma unroll
            for (int d = 0; d < D; ++d) xo[d] = xn[d];
        }
    };
    auto read_block = [&](int slot, float (&v)[NW]) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const float4 f = *SM::blk(ring, slot, k, lane);
            v[4 * k] = f.x; v[4 * k + 1] = f.y; v[4 * k + 2] = f.z; v[4 * k + 3] = f.w;
        }
    };

    float Sinv[D][D];
    float u[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        u[i] = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) Sinv[i][j] = 0.f;
    }
    // factor S (lower triangle in, Cholesky), invert, u = S^-1 y
    auto factor = [&](float (&S)[D][D], const float (&y)[D]) {
        float dinv[D];
        chol_lower<D>(S, dinv);
        chol_inverse<D>(S, dinv, Sinv);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < D; ++j) acc = fmaf(j <= i ? Sinv[i][j] : Sinv[j][i], y[j], acc);
            u[i] = acc;
        }
    };

    // ---- elimination sweep
#pragma unroll
    for (int k = 0; k < SOLVE_RING_FWD; ++k) {
        if (k < n_side) issue_block(k, t_of(k), false);
        cp_async_commit();
    }
    for (int64_t k = 0; k < n_iter; ++k) {
        cp_async_wait<SOLVE_RING_FWD - 1>();
        if (k < n_side) {
            const int slot = (int)(k % SOLVE_RING_FWD);
            float blk[NW];
            read_block(slot, blk);
            float S[D][D], y[D];
#pragma unroll
            for (int i = 0; i < D; ++i) {
                y[i] = fmaf(prm.beta[i], u[i], blk[NT + i]);  // u = 0 on the first block
#pragma unroll
                for (int j = 0; j <= i; ++j) S[i][j] = fmaf(-prm.beta[i] * prm.beta[j], Sinv[i][j], blk[tri(i, j)]);
            }
            if (k + SOLVE_RING_FWD < n_side) issue_block(slot, t_of(k + SOLVE_RING_FWD), false);
            factor(S, y);
#pragma unroll
            for (int i = 0; i < D; ++i) {
                blk[NT + i] = u[i];
#pragma unroll
                for (int j = 0; j <= i; ++j) blk[tri(i, j)] = Sinv[i][j];
            }
            if (active) store_blk(t_of(k), blk);
        }
        cp_async_commit();
    }
    cp_async_wait<0>();

    // ---- middle block: side 0 owns it; side 1 hands over its last (S^-1, u)
    float dx[D];
    {
        float Sm[D][D], ym[D];
        float blk[NW];
#pragma unroll
        for (int k = 0; k < NV; ++k) {  // both lanes of the pair read the middle block (same address)
            const float4 f = w4[(m * NV + k) * P];
            blk[4 * k] = f.x; blk[4 * k + 1] = f.y; blk[4 * k + 2] = f.z; blk[4 * k + 3] = f.w;
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const float uo = __shfl_xor_sync(0xffffffffu, u[i], 1);
            ym[i] = fmaf(prm.beta[i], u[i] + uo, blk[NT + i]);
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const float so = __shfl_xor_sync(0xffffffffu, Sinv[i][j], 1);
                Sm[i][j] = fmaf(-prm.beta[i] * prm.beta[j], Sinv[i][j] + so, blk[tri(i, j)]);
            }
        }
        if (side == 0) {
            float dinv[D], Minv[D][D];
            chol_lower<D>(Sm, dinv);
            chol_inverse<D>(Sm, dinv, Minv);
#pragma unroll
            for (int i = 0; i < D; ++i) {
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < D; ++j) acc = fmaf(j <= i ? Minv[i][j] : Minv[j][i], ym[j], acc);
                dx[i] = acc;
            }
        }
#pragma unroll
        for (int i = 0; i < D; ++i) dx[i] = __shfl_sync(0xffffffffu, dx[i], lane & ~1);
        if (side == 0 && active) {
            float xn[D];
#pragma unroll
            for (int i = 0; i < D; ++i) xn[i] = __ldg(qp + m * D + i) + dx[i];
            if (prm.do_clamp) {
                static_for<D>([&](auto Dd) {
                    constexpr int d = decltype(Dd)::value;
                    xn[d] = fminf(fmaxf(xn[d], dof_lower<M>(d)), dof_upper<M>(d));
                });
            }
            store_x(m, xn);
        }
    }

    // ---- back-substitution outwards from the middle: this lane's blocks k = n_side-1 ... 0
#pragma unroll
    for (int r = 0; r < SOLVE_RING; ++r) {
        const int64_t k = n_side - 1 - r;
        if (k >= 0) issue_block(r, t_of(k), true);
        cp_async_commit();
    }
    for (int64_t r = 0; r < n_iter; ++r) {
        cp_async_wait<SOLVE_RING - 1>();
        const int64_t k = n_side - 1 - r;
        if (k >= 0) {
            const int slot = (int)(r % SOLVE_RING);
            float blk[NW], qv[D];
            read_block(slot, blk);
#pragma unroll
            for (int d = 0; d < D; ++d) qv[d] = *SM::qv(ring, slot, d, lane);
            float z[D];
#pragma unroll
            for (int i = 0; i < D; ++i) z[i] = prm.beta[i] * dx[i];
#pragma unroll
            for (int i = 0; i < D; ++i) {
                float acc = blk[NT + i];
#pragma unroll
                for (int j = 0; j < D; ++j) acc = fmaf(j <= i ? blk[tri(i, j)] : blk[tri(j, i)], z[j], acc);
                dx[i] = acc;
            }
            const int64_t kn = k - SOLVE_RING;
            if (kn >= 0) issue_block(slot, t_of(kn), true);
            float xn[D];
#pragma unroll
            for (int i = 0; i < D; ++i) xn[i] = qv[i] + dx[i];
            if (prm.do_clamp) {
                static_for<D>([&](auto Dd) {
                    constexpr int d = decltype(Dd)::value;
                    xn[d] = fminf(fmaxf(xn[d], dof_lower<M>(d)), dof_upper<M>(d));
                });
            }
            if (active) store_x(t_of(k), xn);
        }
        cp_async_commit();
    }
    cp_async_wait<0>();
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
    const dim3 grid(grid_for(P, ABLOCK), (unsigned)T);
    lm_assemble_kernel<M><<<grid, ABLOCK, sh, st>>>(q, xv, target, (int)P, (int)T, ob, ap, ws);
    return CPPFLOW_OK;
}

template <class M>
static int launch_solve(const cppflow_lm_params* p, const float* q, int64_t P, int64_t T, int do_clamp, float* ws,
                        float* x_out, cudaStream_t st) {
    AssembleParams ap;
    SolveParams sp;
    make_params<M>(p, 0, do_clamp, ap, sp);
    const size_t sh = SolveSmem<M::NDOF>::BYTES;
    static bool attr_set = false;  // per template instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(lm_block_solve_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
        if (e != cudaSuccess) return fail(CPPFLOW_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    lm_block_solve_kernel<M><<<grid_for(P, 16), 32, sh, st>>>(q, P, T, sp, ws, x_out);
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
    CPPFLOW_CHECK_ARG(T <= 65535 && P <= (int64_t)1 << 30, "T must be <= 65535 (grid.y) and P <= 2^30");
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
