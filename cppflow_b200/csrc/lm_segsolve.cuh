// Segmented (parallel-in-time) block-tridiagonal solve for FEW paths.
//
// The twisted block-Thomas solve of k_lm_full.cu is a dependent chain of T / 2 eliminations + T / 2 back-substitutions
// per path: with <= ~2000 paths the GPU is mostly empty and the chain (0.4 us per step) is all there is (0.125 ms for
// T = 300 whatever the path count).  Here a path's waypoints are cut into S segments by S - 1 SEPARATOR waypoints
// s_j = floor(j T / S); the segments between them do not see each other once the separators are known:
//
//   pass 1 (one lane pair per (path, segment), all in parallel; chain = segment length L ~ T / S):
//       lane side 0 eliminates the segment a..e upwards, side 1 downwards, both over the WHOLE segment, with the
//       segment's coupling to its separators left out.  That yields the corners of the segment's inverse
//       [M^-1]_ee = S_e(up)^-1, [M^-1 b]_e = u_e(up), [M^-1]_aa = S_a(down)^-1, [M^-1 b]_a = u_a(down), and the up lane
//       also carries Q = prod_t (S_t^-1 beta), so that [M^-1]_ae beta = Q.  The factors (-S_t^-1, u_t) of the lower
//       half (up lane) and of the upper half (down lane) go to the factor area for pass 3.
//   pass 2 (one CTA per 16 paths; chain = (S - 1) / 2): the separators' own block-tridiagonal system
//       (A_s + beta (nS_left + nS_right) beta) x_s + K_{j-1}^T x_{s_{j-1}} + K_j x_{s_{j+1}} = b_s + beta (u_left + u_right),
//       K_j = -beta Q(segment j + 1) (a full D x D block), built in shared memory by all warps and solved by a twisted
//       block Thomas (two lanes per path), the coupling products of a step one column per warp.
//   pass 3 (one lane pair per (path, segment); chain = L / 2 light steps + one sweep + L / 2 light steps):
//       with x at both separators known, the stored right-hand sides are corrected (du_t = S_t^-1 beta du_{t-1}, a
//       matrix-vector chain starting at the separator's x), the segment's middle block is solved exactly like the
//       twisted solve's, and the back-substitution runs outwards from it.  Default variant: the half-segment's factors
//       and q rows gathered into shared memory once with cp.async (lm_seg_substitute_staged_kernel).
//
// Same arithmetic kernels (sweep_neg_inverse, the beta-folded updates); the result differs from the twisted solve's
// by rounding only (another elimination order), and depends on S but NOT on the path count, the chunking or the
// launch geometry: S is part of the caller's flags (CPPFLOW_LM_SEGMENTS).
//
// Included by k_lm_full.cu after BlockLayout / SolveParams / BetaSel.
#pragma once
#include <type_traits>

namespace cppflow {

struct SegGeom {
    int T, S;
    __host__ __device__ int sep(int j) const { return (int)(((int64_t)j * T) / S); }  // separator j = 1 .. S-1
    __host__ __device__ int first(int s) const { return s == 0 ? 0 : sep(s) + 1; }
    __host__ __device__ int last(int s) const { return s == S - 1 ? T - 1 : sep(s + 1) - 1; }
};

template <int D>
struct SegLayout {
    static constexpr int NT = BlockLayout<D>::NT;
    static constexpr int NW = BlockLayout<D>::NW;
    static constexpr int NV = NW / 4;
    static constexpr int CORNER_F = (NW + D * D + 3) / 4 * 4;  // (nS, u) like a block, then Q row-major
    static constexpr int CORNER_V = CORNER_F / 4;
    static constexpr int DQ = (D + 3) / 4 * 4;
    static constexpr int XV = DQ / 4;
    static_assert(NT % 4 == 0, "the right-hand side of a block must start on a float4");
    // float4 counts per 16-path group
    static constexpr int64_t blk_f4() { return (int64_t)NV * 16; }
    static constexpr int64_t corner_f4() { return (int64_t)CORNER_V * 16; }  // per (segment, side)
    static constexpr int64_t x_f4() { return (int64_t)XV * 16; }             // per separator
};

// v[4k .. 4k+3] = src[k * 16] (one lane's float4 k of a 16-lane interleaved record)
template <int NF4, int NF>
__device__ __forceinline__ void ld_lane(const float4* src, float (&v)[NF]) {  // no __restrict__: pass 2 / 3 re-read their own stores
    static_assert(NF >= NF4 * 4, "destination too small");
#pragma unroll
    for (int k = 0; k < NF4; ++k) {
        const float4 f = src[k * 16];
        v[4 * k] = f.x; v[4 * k + 1] = f.y; v[4 * k + 2] = f.z; v[4 * k + 3] = f.w;
    }
}
template <int NF4, int NF>
__device__ __forceinline__ void st_lane(float4* dst, const float (&v)[NF]) {
    static_assert(NF >= NF4 * 4, "source too small");
#pragma unroll
    for (int k = 0; k < NF4; ++k) dst[k * 16] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}

template <class M>
__device__ __forceinline__ void seg_store_x(float* __restrict__ xo, float (&xn)[M::NDOF], int do_clamp) {
    constexpr int D = M::NDOF;
    if (do_clamp) {
        static_for<D>([&](auto Dd) {
            constexpr int d = decltype(Dd)::value;
            xn[d] = fminf(fmaxf(xn[d], dof_lower<M>(d)), dof_upper<M>(d));
        });
    }
#pragma unroll
    for (int d = 0; d < D; ++d) xo[d] = xn[d];
}

// (nS, u) <- the next Schur complement's negative inverse / eliminated right-hand side, from the block (A, b) in blk
template <class M>
__device__ __forceinline__ void seg_eliminate(const BetaSel<M>& bs, const float (&blk)[BlockLayout<M::NDOF>::NW],
                                              float (&nS)[BlockLayout<M::NDOF>::NT], float (&u)[M::NDOF], float floor) {
    constexpr int D = M::NDOF;
    constexpr int NT = BlockLayout<D>::NT;
    static_for<D>([&](auto Ii) {
        constexpr int i = decltype(Ii)::value;
        u[i] = fmaf(bs.template b<i>(), u[i], blk[NT + i]);
        static_for<i + 1>([&](auto Jj) {
            constexpr int j = decltype(Jj)::value;
            nS[tri(i, j)] = fmaf(bs.template bb<i, j>(), nS[tri(i, j)], blk[tri(i, j)]);
        });
    });
    sweep_neg_inverse<D>(nS, u, floor);
}

// y = -nS z  (nS packed lower triangle of a symmetric matrix), two accumulators per row
template <int D>
__device__ __forceinline__ void neg_symv(const float* __restrict__ nS, const float (&z)[D], float (&y)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j = 0; j < D; j += 2) {
            a0 = fmaf(-(j <= i ? nS[tri(i, j)] : nS[tri(j, i)]), z[j], a0);
            if (j + 1 < D) a1 = fmaf(-(j + 1 <= i ? nS[tri(i, j + 1)] : nS[tri(j + 1, i)]), z[j + 1], a1);
        }
        y[i] = a0 + a1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// pass 1: one warp = 16 paths x 2 directions of ONE segment
template <class M, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 1)
lm_seg_eliminate_kernel(int64_t P, SegGeom geo, const SolveParams prm, const float* __restrict__ ws,
                        float* __restrict__ fac, float* __restrict__ corners) {
    constexpr int D = M::NDOF;
    using LY = SegLayout<D>;
    constexpr int NT = LY::NT, NW = LY::NW, NV = LY::NV;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t n_groups = (P + 15) / 16;
    if (w >= n_groups * geo.S) return;
    const int64_t g = w / geo.S;
    const int seg = (int)(w - g * geo.S);
    const int side = lane & 1, l = lane >> 1;
    const int a = geo.first(seg), e = geo.last(seg);
    const int L = e - a + 1, mid = a + L / 2;
    const bool interior = seg > 0 && seg < geo.S - 1;  // only those couple two separators: Q needed
    const float4* wsg = reinterpret_cast<const float4*>(ws) + g * geo.T * LY::blk_f4() + l;
    float4* facg = reinterpret_cast<float4*>(fac) + g * geo.T * LY::blk_f4() + l;
    const BetaSel<M> bs(prm.b_rev, prm.b_pri);
    auto t_of = [&](int k) { return side == 0 ? a + k : e - k; };

    // Q = prod (S_t^-1 beta) of the UP lane's sweep; its rows are independent, so the pair splits them: the up lane
    // keeps rows 0 .. R0-1, the down lane rows R0 .. D-1 (it gets the up lane's -S_t^-1 by shuffle)
    constexpr int R0 = (D + 1) / 2;
    static_assert((R0 * D) % 4 == 0, "the down lane's rows of Q must start on a float4");
    float nS[NT], u[D], Q[R0 * D];
#pragma unroll
    for (int k = 0; k < NT; ++k) nS[k] = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) u[d] = 0.f;
#pragma unroll
    for (int i = 0; i < R0 * D; ++i) Q[i] = (side * R0 + i / D == i % D) ? 1.f : 0.f;

    // Q <- Q (S^-1 beta): row i of Q times the symmetric G = -nS of the up lane, column j scaled by beta_j
    float G[NT];
    auto q_update = [&]() {
#pragma unroll
        for (int i = 0; i < R0; ++i) {
            float r[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int c = 0; c < D; c += 2) {
                    a0 = fmaf(Q[i * D + c], (c <= j ? G[tri(j, c)] : G[tri(c, j)]), a0);
                    if (c + 1 < D) a1 = fmaf(Q[i * D + c + 1], (c + 1 <= j ? G[tri(j, c + 1)] : G[tri(c + 1, j)]), a1);
                }
                r[j] = a0 + a1;
            }
            static_for<D>([&](auto Jj) {
                constexpr int j = decltype(Jj)::value;
                Q[i * D + j] = -bs.template b<j>() * r[j];
            });
        }
    };
    float nxt[NW];
    ld_lane<NV>(wsg + t_of(0) * LY::blk_f4(), nxt);
    // one elimination step; WITH_Q: the Q update of the PREVIOUS step sits in the same basic block as this step's sweep
    // (it does not depend on it), so its 256 FMAs fill the issue slots the sweep's reciprocal chain leaves empty
    auto step = [&](int k, auto with_q, auto keep_g) {
        const int t = t_of(k);
        float blk[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) blk[i] = nxt[i];
        if (k + 1 < L) ld_lane<NV>(wsg + t_of(k + 1) * LY::blk_f4(), nxt);
        if constexpr (decltype(with_q)::value) q_update();
        seg_eliminate<M>(bs, blk, nS, u, prm.pivot_floor);
        if constexpr (decltype(keep_g)::value) {
#pragma unroll
            for (int i = 0; i < NT; ++i) G[i] = __shfl_sync(0xffffffffu, nS[i], lane & ~1);
        }
        if (side == 0 ? t < mid : t > mid) {
            float v[NW];
#pragma unroll
            for (int i = 0; i < NT; ++i) v[i] = nS[i];
#pragma unroll
            for (int d = 0; d < D; ++d) v[NT + d] = u[d];
#pragma unroll
            for (int i = NT + D; i < NW; ++i) v[i] = 0.f;
            st_lane<NV>(facg + t * LY::blk_f4(), v);
        }
    };
    if (interior) {
        step(0, std::false_type{}, std::true_type{});
        for (int k = 1; k < L; ++k) step(k, std::true_type{}, std::true_type{});
        q_update();
    } else {
        for (int k = 0; k < L; ++k) step(k, std::false_type{}, std::false_type{});
    }
    // corner record of this (segment, side): (nS, u) of the last block eliminated, then Q
    float4* cg = reinterpret_cast<float4*>(corners) + ((g * geo.S + seg) * 2 + side) * LY::corner_f4() + l;
    {
        float v[NW];
#pragma unroll
        for (int i = 0; i < NT; ++i) v[i] = nS[i];
#pragma unroll
        for (int d = 0; d < D; ++d) v[NT + d] = u[d];
#pragma unroll
        for (int i = NT + D; i < NW; ++i) v[i] = 0.f;
        st_lane<NV>(cg, v);
    }
    if (interior) {  // both lanes: their rows of Q, behind the UP lane's (nS, u)
        constexpr int QF = LY::CORNER_F - NW;
        constexpr int F0 = R0 * D / 4, F1 = QF / 4 - F0;  // float4 per lane
        static_assert(F1 * 4 >= (D - R0) * D && F1 <= F0, "Q rows of the down lane");
        float4* cq = reinterpret_cast<float4*>(corners) + ((g * geo.S + seg) * 2 + 0) * LY::corner_f4() + l + NV * 16;
        float v[F0 * 4];
#pragma unroll
        for (int i = 0; i < F0 * 4; ++i) v[i] = (side * R0 * D + i < D * D) ? Q[i] : 0.f;
        if (side == 0) st_lane<F0>(cq, v);
        else st_lane<F1>(cq + F0 * 16, v);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// pass 2: the separators' block-tridiagonal system (full D x D couplings K_j between separators j and j + 1).
// One CTA per 16-path group.  Phase A, all warps: thread (separator j, path) builds the separator's block
//   A^_j = A_s + beta (nS_left + nS_right) beta,  b^_j = b_s + beta (u_left + u_right),  K_j = -beta Q(segment j + 1)
// in shared memory (the loads of all separators are in flight together instead of one dependent batch per step of the
// chain).  Phase B, warp 0: twisted block Thomas over the n = S - 1 separators, two lanes per path, out of shared memory:
//   step into node j from the node `in` already eliminated, C = the block (in, j) of the reduced matrix:
//       S_j = A^_j - C^T S_in^-1 C,  y_j = b^_j - C^T u_in,  node j <- (-S_j^-1, u_j = S_j^-1 y_j)
//   the middle node takes both neighbours' terms (lane shuffles); back-substitution x_j = u_j - S_j^-1 C x_inner with
//   the C of the step from j into its inner neighbour.  Up lane: C = K_{j-1}; down lane: C = K_j^T.
constexpr int SEG_MAX_SEGMENTS = 32;  // shared memory of pass 2: (S - 1) x (NW + D D) x 16 floats
constexpr int SEG_REDUCED_WARPS = 8;

template <int D>
constexpr int seg_node_floats() { return (BlockLayout<D>::NW + D * D) * 16; }  // [NW + D*D][16 paths]
template <int D>
constexpr size_t seg_reduced_smem(int S) {  // the separators' nodes + one coupling term [NT + D][32 lanes]
    return ((size_t)(S - 1) * seg_node_floats<D>() + (size_t)(BlockLayout<D>::NT + D) * 32) * sizeof(float);
}

template <class M>
__global__ void __launch_bounds__(32 * SEG_REDUCED_WARPS, 1)
lm_seg_reduced_kernel(int64_t P, SegGeom geo, const SolveParams prm, const float* __restrict__ ws,
                      const float* __restrict__ corners, float* __restrict__ sepx) {
    constexpr int D = M::NDOF;
    using LY = SegLayout<D>;
    constexpr int NT = LY::NT, NW = LY::NW, NV = LY::NV;
    constexpr int NODE = seg_node_floats<D>();
    extern __shared__ __align__(16) float seg_sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t g = blockIdx.x;
    const int n = geo.S - 1;
    const BetaSel<M> bs(prm.b_rev, prm.b_pri);
    // node j, float f, path l  ->  seg_sm[j * NODE + f * 16 + l]
    {
        const int l = lane & 15;
        const float4* wsg = reinterpret_cast<const float4*>(ws) + g * geo.T * LY::blk_f4() + l;
        const float4* cgrp = reinterpret_cast<const float4*>(corners) + g * geo.S * 2 * LY::corner_f4() + l;
        for (int j = warp * 2 + (lane >> 4); j < n; j += 2 * SEG_REDUCED_WARPS) {
            float A[NW], cl[NW], cr[NW];
            ld_lane<NV>(wsg + (int64_t)geo.sep(j + 1) * LY::blk_f4(), A);
            ld_lane<NV>(cgrp + (j * 2 + 0) * LY::corner_f4(), cl);        // left segment, eliminated upwards: its last block
            ld_lane<NV>(cgrp + ((j + 1) * 2 + 1) * LY::corner_f4(), cr);  // right segment, eliminated downwards: its first
            float* node = seg_sm + j * NODE + l;
            static_for<D>([&](auto Ii) {
                constexpr int i = decltype(Ii)::value;
                node[(NT + i) * 16] = fmaf(bs.template b<i>(), cl[NT + i] + cr[NT + i], A[NT + i]);
                static_for<i + 1>([&](auto Jj) {
                    constexpr int jj = decltype(Jj)::value;
                    node[tri(i, jj) * 16] = fmaf(bs.template bb<i, jj>(), cl[tri(i, jj)] + cr[tri(i, jj)], A[tri(i, jj)]);
                });
            });
            if (j + 1 < n) {  // K_j through segment j + 1 (an interior segment)
                constexpr int QF = LY::CORNER_F - NW;
                float v[QF];
                ld_lane<QF / 4>(cgrp + ((j + 1) * 2 + 0) * LY::corner_f4() + NV * 16, v);
                static_for<D>([&](auto Ii) {
                    constexpr int i = decltype(Ii)::value;
#pragma unroll
                    for (int c = 0; c < D; ++c) node[(NW + i * D + c) * 16] = -bs.template b<i>() * v[i * D + c];
                });
            }
        }
    }
    __syncthreads();

    const int side = lane & 1, l = lane >> 1;
    const int m = n / 2;
    const int n_side = side == 0 ? m : n - 1 - m;
    const int n_steps = m > n - 1 - m ? m : n - 1 - m;  // uniform: every warp meets every barrier
    auto node_of = [&](int k) { return side == 0 ? k : n - 1 - k; };
    float* contrib = seg_sm + n * NODE;  // [NT + D][32 lanes]: the coupling term of the step being taken
    // The step into a node costs C^T (nS C) (two D x D x D products) before its sweep can start: the eight warps take one
    // COLUMN c of it each (same lane <-> (path, side) mapping in every warp), warp 0 adds the columns to the node's block
    // and sweeps.  Term of the eliminated node `from` for its inner neighbour, C = the block (from, inner):
    //     contrib[tri(i, c)] = (C^T (nS C))_ic  (i >= c),   contrib[NT + c] = -(C^T u)_c,   nS = -S_from^-1
    auto contribute = [&](int from, int c) {
        const float* nd = seg_sm + from * NODE + l;
        const float* K = seg_sm + (side == 0 ? from : from - 1) * NODE + NW * 16 + l;  // up lane K_from, down lane K_{from-1}^T
        auto Cel = [&](int i, int cc) { return K[(side == 0 ? i * D + cc : cc * D + i) * 16]; };
        float nSin[NT], Cc[D], Pc[D];
#pragma unroll
        for (int i = 0; i < NT; ++i) nSin[i] = nd[i * 16];
#pragma unroll
        for (int k = 0; k < D; ++k) Cc[k] = Cel(k, c);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int k = 0; k < D; k += 2) {
                a0 = fmaf((k <= i ? nSin[tri(i, k)] : nSin[tri(k, i)]), Cc[k], a0);
                if (k + 1 < D) a1 = fmaf((k + 1 <= i ? nSin[tri(i, k + 1)] : nSin[tri(k + 1, i)]), Cc[k + 1], a1);
            }
            Pc[i] = a0 + a1;
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int k = 0; k < D; k += 2) {
                a0 = fmaf(Cel(k, i), Pc[k], a0);
                if (k + 1 < D) a1 = fmaf(Cel(k + 1, i), Pc[k + 1], a1);
            }
            if (i >= c) contrib[(i * (i + 1) / 2 + c) * 32 + lane] = a0 + a1;
        }
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) acc = fmaf(-Cc[k], nd[(NT + k) * 16], acc);
        contrib[(NT + c) * 32 + lane] = acc;
    };
    // C of the step from node `from` into its inner neighbour, whole (back-substitution)
    auto load_C = [&](int from, float (&C)[D * D]) {
        const float* K = seg_sm + (side == 0 ? from : from - 1) * NODE + NW * 16 + l;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int c = 0; c < D; ++c) C[i * D + c] = K[(side == 0 ? i * D + c : c * D + i) * 16];
    };

    for (int k = 0; k < n_steps; ++k) {
        const bool valid = k < n_side;
        if (k > 0 && valid && warp < D) contribute(node_of(k - 1), warp);
        __syncthreads();
        if (warp == 0 && valid) {
            float* node = seg_sm + node_of(k) * NODE + l;
            float nS[NT], u[D];
#pragma unroll
            for (int i = 0; i < NT; ++i) nS[i] = node[i * 16] + (k > 0 ? contrib[i * 32 + lane] : 0.f);
#pragma unroll
            for (int d = 0; d < D; ++d) u[d] = node[(NT + d) * 16] + (k > 0 ? contrib[(NT + d) * 32 + lane] : 0.f);
            sweep_neg_inverse<D>(nS, u, prm.pivot_floor);
#pragma unroll
            for (int i = 0; i < NT; ++i) node[i * 16] = nS[i];
#pragma unroll
            for (int d = 0; d < D; ++d) node[(NT + d) * 16] = u[d];
        }
        __syncthreads();
    }
    if (n_side > 0 && warp < D) contribute(node_of(n_side - 1), warp);  // both sides' terms of the middle node
    __syncthreads();
    if (warp != 0) return;
    float x[D], C[D * D];
    {
        const float* node = seg_sm + m * NODE + l;
        float Sm[NT];
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            const float w = n_side > 0 ? contrib[i * 32 + lane] : 0.f;
            Sm[i] = node[i * 16] + (w + __shfl_xor_sync(0xffffffffu, w, 1));
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float w = n_side > 0 ? contrib[(NT + d) * 32 + lane] : 0.f;
            x[d] = node[(NT + d) * 16] + (w + __shfl_xor_sync(0xffffffffu, w, 1));
        }
        sweep_neg_inverse<D>(Sm, x, prm.pivot_floor);
    }
    // the separators' dx go to `sepx`; pass 3 adds them to q (a global load here would stall the chain at every node)
    auto emit = [&](int j, const float (&xj)[D]) {
        float xs[LY::DQ];
#pragma unroll
        for (int d = 0; d < LY::DQ; ++d) xs[d] = d < D ? xj[d] : 0.f;
        st_lane<LY::XV>(reinterpret_cast<float4*>(sepx) + (g * n + j) * LY::x_f4() + l, xs);
    };
    if (side == 0) emit(m, x);
    // back-substitution outwards: x_j = u_j + nS_j (C x_inner)
    for (int k = n_side - 1; k >= 0; --k) {
        const int j = node_of(k);
        const float* node = seg_sm + j * NODE + l;
        load_C(j, C);
        float z[D], y[D], f[NT];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < D; ++c) acc = fmaf(C[i * D + c], x[c], acc);
            z[i] = -acc;
        }
#pragma unroll
        for (int i = 0; i < NT; ++i) f[i] = node[i * 16];
        neg_symv<D>(f, z, y);  // -nS (-C x) = nS C x
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = node[(NT + d) * 16] + y[d];
        emit(j, x);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// pass 3: inside the segments, separators known
template <class M, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 1)
lm_seg_substitute_kernel(const float* __restrict__ q, int64_t P, SegGeom geo, const SolveParams prm,
                         const float* __restrict__ ws, float* __restrict__ fac, const float* __restrict__ sepx,
                         float* __restrict__ x_out) {
    constexpr int D = M::NDOF;
    using LY = SegLayout<D>;
    constexpr int NT = LY::NT, NW = LY::NW, NV = LY::NV;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t n_groups = (P + 15) / 16;
    if (w >= n_groups * geo.S) return;
    const int64_t g = w / geo.S;
    const int seg = (int)(w - g * geo.S);
    const int side = lane & 1, l = lane >> 1;
    const int64_t p_raw = g * 16 + l;
    const bool active = p_raw < P;
    const int64_t p = active ? p_raw : P - 1;
    const int a = geo.first(seg), e = geo.last(seg);
    const int L = e - a + 1, mid = a + L / 2;
    const int n_side = side == 0 ? mid - a : e - mid;
    const float4* wsg = reinterpret_cast<const float4*>(ws) + g * geo.T * LY::blk_f4() + l;
    float4* facg = reinterpret_cast<float4*>(fac) + g * geo.T * LY::blk_f4() + l;
    const float4* xg = reinterpret_cast<const float4*>(sepx) + g * (geo.S - 1) * LY::x_f4() + l;
    const float* qp = q + p * geo.T * D;
    float* xo = x_out + p * geo.T * D;
    const BetaSel<M> bs(prm.b_rev, prm.b_pri);
    auto t_of = [&](int k) { return side == 0 ? a + k : e - k; };

    auto load_qrow = [&](int t, float (&v)[D]) {
#pragma unroll
        for (int d = 0; d < D; ++d) v[d] = __ldg(qp + (int64_t)t * D + d);
    };
    // the separator this side faces (none at the ends of the path: coupling 0); the up lane also writes ITS x
    float du[D], xsep[D], qsep[D], qmid[D];
    const bool writes_sep = side == 0 && seg > 0;
    {
        float xs[LY::DQ];
#pragma unroll
        for (int d = 0; d < LY::DQ; ++d) xs[d] = 0.f;
        const int js = side == 0 ? seg - 1 : seg;  // separator index 0 .. S-2
        if (js >= 0 && js < geo.S - 1) ld_lane<LY::XV>(xg + js * LY::x_f4(), xs);
        load_qrow(writes_sep ? geo.sep(seg) : mid, qsep);  // both q rows are needed at the END of the chains only
        load_qrow(mid, qmid);
#pragma unroll
        for (int d = 0; d < D; ++d) du[d] = xsep[d] = xs[d];
    }
    // corrected right-hand sides: du_t = S_t^-1 (beta . du_{t-1}),  u_t += du_t
    float nS[NT], u[D];
#pragma unroll
    for (int k = 0; k < NT; ++k) nS[k] = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) u[d] = du[d];
    // the steps are light (one matrix-vector product): blocks are requested TWO steps ahead of their use
    float nxt[NW], nx2[NW];
    if (n_side > 0) ld_lane<NV>(facg + t_of(0) * LY::blk_f4(), nxt);
    if (n_side > 1) ld_lane<NV>(facg + t_of(1) * LY::blk_f4(), nx2);
    for (int k = 0; k < n_side; ++k) {
        const int t = t_of(k);
        float blk[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            blk[i] = nxt[i];
            nxt[i] = nx2[i];
        }
        if (k + 2 < n_side) ld_lane<NV>(facg + t_of(k + 2) * LY::blk_f4(), nx2);
        float z[D], y[D];
        static_for<D>([&](auto Ii) {
            constexpr int i = decltype(Ii)::value;
            z[i] = bs.template b<i>() * du[i];
        });
        neg_symv<D>(blk, z, y);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            du[d] = y[d];
            u[d] = blk[NT + d] + y[d];
        }
#pragma unroll
        for (int i = 0; i < NT; ++i) nS[i] = blk[i];
        float v[NW - NT];
#pragma unroll
        for (int i = 0; i < NW - NT; ++i) v[i] = i < D ? u[i] : 0.f;
        st_lane<(NW - NT) / 4>(facg + t * LY::blk_f4() + (NT / 4) * 16, v);
    }

    // first back-substitution block and q row on their way while the middle block is factorised
    float qn[D], qn2[D];
    if (n_side > 0) {
        ld_lane<NV>(facg + t_of(n_side - 1) * LY::blk_f4(), nxt);  // own stores: same-thread ordering
        load_qrow(t_of(n_side - 1), qn);
    }
    if (n_side > 1) {
        ld_lane<NV>(facg + t_of(n_side - 2) * LY::blk_f4(), nx2);
        load_qrow(t_of(n_side - 2), qn2);
    }

    __syncwarp();
    float dx[D];
    {
        float Sm[NT], blk[NW];
        ld_lane<NV>(wsg + (int64_t)mid * LY::blk_f4(), blk);
        static_for<D>([&](auto Ii) {
            constexpr int i = decltype(Ii)::value;
            const float uo = __shfl_xor_sync(0xffffffffu, u[i], 1);
            dx[i] = fmaf(bs.template b<i>(), u[i] + uo, blk[NT + i]);
            static_for<i + 1>([&](auto Jj) {
                constexpr int j = decltype(Jj)::value;
                const float so = __shfl_xor_sync(0xffffffffu, nS[tri(i, j)], 1);
                Sm[tri(i, j)] = fmaf(bs.template bb<i, j>(), nS[tri(i, j)] + so, blk[tri(i, j)]);
            });
        });
        sweep_neg_inverse<D>(Sm, dx, prm.pivot_floor);
        if (side == 0 && active) {
            float xn[D];
#pragma unroll
            for (int i = 0; i < D; ++i) xn[i] = qmid[i] + dx[i];
            seg_store_x<M>(xo + (int64_t)mid * D, xn, prm.do_clamp);
        }
        if (writes_sep && active) {
#pragma unroll
            for (int i = 0; i < D; ++i) qsep[i] += xsep[i];
            seg_store_x<M>(xo + (int64_t)geo.sep(seg) * D, qsep, prm.do_clamp);
        }
    }

    for (int k = n_side - 1; k >= 0; --k) {
        float blk[NW], xn[D];
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            blk[i] = nxt[i];
            nxt[i] = nx2[i];
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
            xn[d] = qn[d];
            qn[d] = qn2[d];
        }
        if (k > 1) {
            ld_lane<NV>(facg + t_of(k - 2) * LY::blk_f4(), nx2);
            load_qrow(t_of(k - 2), qn2);
        }
        float z[D], y[D];
        static_for<D>([&](auto Ii) {
            constexpr int i = decltype(Ii)::value;
            z[i] = bs.template b<i>() * dx[i];
        });
        neg_symv<D>(blk, z, y);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            dx[i] = blk[NT + i] + y[i];
            xn[i] += dx[i];
        }
        if (active) seg_store_x<M>(xo + (int64_t)t_of(k) * D, xn, prm.do_clamp);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// pass 3, staged: the steps of pass 3 are light (one matrix-vector product), so with registers only two blocks can be
// requested ahead and at >= 256 paths every step waits for its block (ncu: 70 % of the samples on the long scoreboard).
// Here a warp gathers the factors AND q rows of its whole half-segment into shared memory with cp.async at the start
// (every load of the pass in flight at once), runs the du chain, the middle block and the back-substitution out of
// shared memory (the corrected right-hand sides never go back to global memory), and reads the factor area once instead
// of twice.  One single-warp CTA per (16-path group, segment); shared memory = ceil(L / 2) x (NW + DQ) x 32 lanes.
template <int D>
constexpr int seg_stage_floats() { return (BlockLayout<D>::NW + (D + 3) / 4 * 4) * 32; }  // per step of a half-segment

template <class M>
__global__ void __launch_bounds__(32, 1)
lm_seg_substitute_staged_kernel(const float* __restrict__ q, int64_t P, SegGeom geo, const SolveParams prm,
                                const float* __restrict__ ws, const float* __restrict__ fac,
                                const float* __restrict__ sepx, float* __restrict__ x_out) {
    constexpr int D = M::NDOF;
    using LY = SegLayout<D>;
    constexpr int NT = LY::NT, NW = LY::NW, NV = LY::NV, DQ = LY::DQ;
    constexpr int STAGE = seg_stage_floats<D>();
    extern __shared__ __align__(16) float seg_sm[];
    const int lane = threadIdx.x;
    const int64_t w = blockIdx.x;
    const int64_t g = w / geo.S;
    const int seg = (int)(w - g * geo.S);
    const int side = lane & 1, l = lane >> 1;
    const int64_t p_raw = g * 16 + l;
    const bool active = p_raw < P;
    const int64_t p = active ? p_raw : P - 1;
    const int a = geo.first(seg), e = geo.last(seg);
    const int L = e - a + 1, mid = a + L / 2;
    const int n_side = side == 0 ? mid - a : e - mid;
    const float4* wsg = reinterpret_cast<const float4*>(ws) + g * geo.T * LY::blk_f4() + l;
    const float4* facg = reinterpret_cast<const float4*>(fac) + g * geo.T * LY::blk_f4() + l;
    const float4* xg = reinterpret_cast<const float4*>(sepx) + g * (geo.S - 1) * LY::x_f4() + l;
    const float* qp = q + p * geo.T * D;
    float* xo = x_out + p * geo.T * D;
    const BetaSel<M> bs(prm.b_rev, prm.b_pri);
    auto t_of = [&](int k) { return side == 0 ? a + k : e - k; };
    // step k of this lane: floats [f][lane] at seg_sm + k * STAGE; f < NW the factor block, then the q row
    auto gather = [&](int k) {
        const float4* src = facg + t_of(k) * LY::blk_f4();
#pragma unroll
        for (int f = 0; f < NV; ++f) cp_async16(reinterpret_cast<float4*>(seg_sm + k * STAGE) + f * 32 + lane, src + f * 16);
        const float* qr = qp + (int64_t)t_of(k) * D;
        if constexpr (D % 4 == 0) {
#pragma unroll
            for (int d = 0; d < D; d += 4)
                cp_async16(reinterpret_cast<float4*>(seg_sm + k * STAGE + NW * 32) + (d / 4) * 32 + lane, qr + d);
        } else {
#pragma unroll
            for (int d = 0; d < D; ++d) cp_async4(seg_sm + k * STAGE + NW * 32 + (d / 4) * 128 + lane * 4 + (d & 3), qr + d);
        }
    };
    // float f of this lane's block at step k (blocks are stored as float4 [f / 4][lane])
    auto at = [&](int k, int f) -> float& { return seg_sm[k * STAGE + (f >> 2) * 128 + lane * 4 + (f & 3)]; };

    // first two steps in one group (the du chain starts on them), the rest in a second
    for (int k = 0; k < n_side && k < 2; ++k) gather(k);
    cp_async_commit();
    for (int k = 2; k < n_side; ++k) gather(k);
    cp_async_commit();

    auto load_qrow = [&](int t, float (&v)[D]) {
#pragma unroll
        for (int d = 0; d < D; ++d) v[d] = __ldg(qp + (int64_t)t * D + d);
    };
    float du[D], xsep[D], qsep[D], qmid[D], bmid[NW];
    const bool writes_sep = side == 0 && seg > 0;
    {
        float xs[DQ];
#pragma unroll
        for (int d = 0; d < DQ; ++d) xs[d] = 0.f;
        const int js = side == 0 ? seg - 1 : seg;  // separator index 0 .. S-2
        if (js >= 0 && js < geo.S - 1) ld_lane<LY::XV>(xg + js * LY::x_f4(), xs);
        ld_lane<NV>(wsg + (int64_t)mid * LY::blk_f4(), bmid);  // (A, b) of the middle block: needed after the du chain
        load_qrow(writes_sep ? geo.sep(seg) : mid, qsep);
        load_qrow(mid, qmid);
#pragma unroll
        for (int d = 0; d < D; ++d) du[d] = xsep[d] = xs[d];
    }
    auto read_block = [&](int k, float (&v)[NW]) {
        const float4* src = reinterpret_cast<const float4*>(seg_sm + k * STAGE) + lane;
#pragma unroll
        for (int f = 0; f < NV; ++f) {
            const float4 x4 = src[f * 32];
            v[4 * f] = x4.x; v[4 * f + 1] = x4.y; v[4 * f + 2] = x4.z; v[4 * f + 3] = x4.w;
        }
    };

    float nS[NT], u[D];
#pragma unroll
    for (int k = 0; k < NT; ++k) nS[k] = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) u[d] = du[d];
    cp_async_wait<1>();
    __syncwarp();
    for (int k = 0; k < n_side; ++k) {
        if (k == 2) cp_async_wait<0>();  // a lane only reads what its own cp.async wrote: no warp barrier (the two
                                         // lanes of a pair may leave this loop one step apart)
        float blk[NW];
        read_block(k, blk);
        float z[D], y[D];
        static_for<D>([&](auto Ii) {
            constexpr int i = decltype(Ii)::value;
            z[i] = bs.template b<i>() * du[i];
        });
        neg_symv<D>(blk, z, y);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            du[d] = y[d];
            u[d] = blk[NT + d] + y[d];
            at(k, NT + d) = u[d];  // the corrected right-hand side, for the back-substitution
        }
#pragma unroll
        for (int i = 0; i < NT; ++i) nS[i] = blk[i];
    }
    cp_async_wait<0>();
    __syncwarp();

    float dx[D];
    {
        float Sm[NT];
        static_for<D>([&](auto Ii) {
            constexpr int i = decltype(Ii)::value;
            const float uo = __shfl_xor_sync(0xffffffffu, u[i], 1);
            dx[i] = fmaf(bs.template b<i>(), u[i] + uo, bmid[NT + i]);
            static_for<i + 1>([&](auto Jj) {
                constexpr int j = decltype(Jj)::value;
                const float so = __shfl_xor_sync(0xffffffffu, nS[tri(i, j)], 1);
                Sm[tri(i, j)] = fmaf(bs.template bb<i, j>(), nS[tri(i, j)] + so, bmid[tri(i, j)]);
            });
        });
        sweep_neg_inverse<D>(Sm, dx, prm.pivot_floor);
        if (side == 0 && active) {
            float xn[D];
#pragma unroll
            for (int i = 0; i < D; ++i) xn[i] = qmid[i] + dx[i];
            seg_store_x<M>(xo + (int64_t)mid * D, xn, prm.do_clamp);
        }
        if (writes_sep && active) {
#pragma unroll
            for (int i = 0; i < D; ++i) qsep[i] += xsep[i];
            seg_store_x<M>(xo + (int64_t)geo.sep(seg) * D, qsep, prm.do_clamp);
        }
    }

    for (int k = n_side - 1; k >= 0; --k) {
        float blk[NW], xn[D];
        read_block(k, blk);
#pragma unroll
        for (int d = 0; d < D; ++d) xn[d] = at(k, NW + d);
        float z[D], y[D];
        static_for<D>([&](auto Ii) {
            constexpr int i = decltype(Ii)::value;
            z[i] = bs.template b<i>() * dx[i];
        });
        neg_symv<D>(blk, z, y);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            dx[i] = blk[NT + i] + y[i];
            xn[i] += dx[i];
        }
        if (active) seg_store_x<M>(xo + (int64_t)t_of(k) * D, xn, prm.do_clamp);
    }
}

}  // namespace cppflow
