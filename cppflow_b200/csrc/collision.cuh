// Device-side capsule collision evaluation shared by the distance kernels (k_collision.cu) and the fused LM
// assembly (k_lm_full.cu).  World-space capsule endpoints and joint axes/origins of one configuration live in a
// per-thread column of shared memory (slot k of thread tid at sm[k * BLOCK + tid]: conflict-free).
#pragma once
#include "robots.cuh"
#include "kinematics.cuh"
#include "geom.cuh"

namespace cppflow {

// constexpr square root (Newton) for the compile-time capsule half-lengths
__host__ __device__ constexpr double csqrt(double x) {
    if (x <= 0.0) return 0.0;
    double r = x > 1.0 ? x : 1.0;
    for (int i = 0; i < 60; ++i) r = 0.5 * (r + x / r);
    return r;
}
template <class M>
__host__ __device__ constexpr float cap_half_length(int c) {
    const double dx = (double)M::cap(c, 3) - (double)M::cap(c, 0);
    const double dy = (double)M::cap(c, 4) - (double)M::cap(c, 1);
    const double dz = (double)M::cap(c, 5) - (double)M::cap(c, 2);
    return (float)(0.5 * csqrt(dx * dx + dy * dy + dz * dz));
}
// slack added to every culling bound so that fp32 rounding can never cull a pair whose distance is < threshold
#define CPPFLOW_CULL_MARGIN 1e-4f

template <class M>
struct PairTable {
    int a[M::NPAIR];
    int b[M::NPAIR];
    int fa[M::NPAIR];  // link frame of capsule a / b
    int fb[M::NPAIR];
    float rsum[M::NPAIR];
    float reach[M::NPAIR];  // half-lengths + radii + margin: |mid_a - mid_b| - reach is a lower bound of the distance
};
template <class M>
__host__ __device__ constexpr PairTable<M> make_pair_table() {
    PairTable<M> t{};
    for (int p = 0; p < M::NPAIR; ++p) {
        t.a[p] = pair_cap<M>(p, 0);
        t.b[p] = pair_cap<M>(p, 1);
        t.fa[p] = M::cap_frame(t.a[p]);
        t.fb[p] = M::cap_frame(t.b[p]);
        t.rsum[p] = M::cap(t.a[p], 6) + M::cap(t.b[p], 6);
        t.reach[p] = t.rsum[p] + cap_half_length<M>(t.a[p]) + cap_half_length<M>(t.b[p]) + CPPFLOW_CULL_MARGIN;
    }
    return t;
}
template <class M>
struct CapTable {
    int frame[M::NCAP];
    float radius[M::NCAP];
    float reach[M::NCAP];  // half-length + radius + margin
};
template <class M>
__host__ __device__ constexpr CapTable<M> make_cap_table() {
    CapTable<M> t{};
    for (int c = 0; c < M::NCAP; ++c) {
        t.frame[c] = M::cap_frame(c);
        t.radius[c] = M::cap(c, 6);
        t.reach[c] = M::cap(c, 6) + cap_half_length<M>(c) + CPPFLOW_CULL_MARGIN;
    }
    return t;
}
template <class M>
__constant__ PairTable<M> c_pair_table = make_pair_table<M>();
template <class M>
__constant__ CapTable<M> c_cap_table = make_cap_table<M>();

template <class M>
struct SmemLayout {
    static constexpr int CAPS = 0;                 // capsule c endpoint e coordinate r: CAPS + c*6 + e*3 + r
    static constexpr int JOINTS = M::NCAP * 6;     // joint d axis r: JOINTS + d*6 + r ; origin r: JOINTS + d*6 + 3 + r
    static constexpr int N_DIST = M::NCAP * 6;     // floats per thread when only distances are needed
    static constexpr int N_FULL = M::NCAP * 6 + M::NDOF * 6;
};

// FK sink writing world capsule endpoints (and optionally joint axes/origins) into the thread's smem column
template <class M, int BLOCK, bool WITH_JOINTS>
struct CollisionSink {
    float* sm;  // smem base + tid
    template <int D>
    __device__ __forceinline__ void joint(std::integral_constant<int, D>, const float* axis, const float* origin) {
        if constexpr (WITH_JOINTS) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                sm[(SmemLayout<M>::JOINTS + D * 6 + r) * BLOCK] = axis[r];
                sm[(SmemLayout<M>::JOINTS + D * 6 + 3 + r) * BLOCK] = origin[r];
            }
        }
    }
    template <int F>
    __device__ __forceinline__ void frame(std::integral_constant<int, F>, const Frame& fr) {
        static_for<M::NCAP>([&](auto Cc) {
            constexpr int c = decltype(Cc)::value;
            if constexpr (M::cap_frame(c) == F) {
                float P[3], Q[3];
                capsule_endpoint<M, c, 0>(fr, P);
                capsule_endpoint<M, c, 1>(fr, Q);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    sm[(c * 6 + r) * BLOCK] = P[r];
                    sm[(c * 6 + 3 + r) * BLOCK] = Q[r];
                }
            }
        });
    }
};

template <int BLOCK>
__device__ __forceinline__ void load_capsule(const float* sm, int c, float (&P)[3], float (&Q)[3]) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        P[r] = sm[(c * 6 + r) * BLOCK];
        Q[r] = sm[(c * 6 + 3 + r) * BLOCK];
    }
}

// Signed distance of self-collision pair p; C2 = closest point on the axis of capsule b, nrm = unit vector from it
// to the closest point on capsule a.  Pairs whose bounding spheres prove distance > cull_thr are skipped and
// +INFINITY is returned (pass cull_thr = INFINITY to always get the exact distance).
template <class M, int BLOCK>
__device__ __forceinline__ float self_pair_distance(const float* sm, int p, float (&C2)[3], float (&nrm)[3],
                                                    float cull_thr = INFINITY) {
    float P1[3], Q1[3], P2[3], Q2[3];
    load_capsule<BLOCK>(sm, c_pair_table<M>.a[p], P1, Q1);
    load_capsule<BLOCK>(sm, c_pair_table<M>.b[p], P2, Q2);
    {
        const float mx = (P1[0] + Q1[0]) - (P2[0] + Q2[0]);
        const float my = (P1[1] + Q1[1]) - (P2[1] + Q2[1]);
        const float mz = (P1[2] + Q1[2]) - (P2[2] + Q2[2]);
        const float lim = cull_thr + c_pair_table<M>.reach[p];  // > 0 whenever a cull is intended
        if (0.25f * (mx * mx + my * my + mz * mz) > lim * lim && lim > 0.f) return INFINITY;
    }
    float s, t;
    segseg_closest(P1, Q1, P2, Q2, s, t);
    float diff[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float c1 = fmaf(s, Q1[r] - P1[r], P1[r]);
        C2[r] = fmaf(t, Q2[r] - P2[r], P2[r]);
        diff[r] = c1 - C2[r];
    }
    const float d2 = dot3(diff, diff);
    const float dist = sqrtf(d2);
    const float inv = d2 > 1e-24f ? rsqrtf(d2) : 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) nrm[r] = diff[r] * inv;
    return dist - c_pair_table<M>.rsum[p];
}

// d(distance)/dq for pair p: only the joints between the two links contribute: g_d = -n . v_d(C2)
template <class M, int BLOCK>
__device__ __forceinline__ void self_pair_gradient(const float* sm, int p, const float (&C2)[3], const float (&nrm)[3],
                                                   float (&g)[M::NDOF]) {
    const int fa = c_pair_table<M>.fa[p], fb = c_pair_table<M>.fb[p];
    static_for<M::NDOF>([&](auto Dd) {
        constexpr int d = decltype(Dd)::value;
        constexpr int ci = chain_of_dof<M>(d);
        float gd = 0.f;
        if (ci >= fa && ci < fb) {
            float a[3], o[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                a[r] = sm[(SmemLayout<M>::JOINTS + d * 6 + r) * BLOCK];
                o[r] = sm[(SmemLayout<M>::JOINTS + d * 6 + 3 + r) * BLOCK];
            }
            if constexpr (dof_is_prismatic<M>(d)) {
                gd = -dot3(nrm, a);
            } else {
                const float rr[3] = {C2[0] - o[0], C2[1] - o[1], C2[2] - o[2]};
                float v[3];
                cross3(a, rr, v);
                gd = -dot3(nrm, v);
            }
        }
        g[d] = gd;
    });
}

// signed distance of capsule c to obstacle o; Cw = closest point on the capsule axis (world), nrm = world normal
template <class M, int BLOCK>
__device__ __forceinline__ float env_capsule_distance(const float* sm, int c, const Obstacles& ob, int o,
                                                      float (&Cw)[3], float (&nrm)[3], float cull_thr = INFINITY) {
    float P[3], Q[3], A[3], B[3];
    load_capsule<BLOCK>(sm, c, P, Q);
    to_box_frame(ob, o, P, A);
    to_box_frame(ob, o, Q, B);
    {
        // distance(segment midpoint, box) - half length - radius is a lower bound of the capsule-box distance
        float dd = 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float m = 0.5f * (A[r] + B[r]);
            const float e = m - fminf(fmaxf(m, ob.lo[o][r]), ob.hi[o][r]);
            dd = fmaf(e, e, dd);
        }
        const float lim = cull_thr + c_cap_table<M>.reach[c];
        if (dd > lim * lim && lim > 0.f) return INFINITY;
    }
    const float t = segbox_closest(A, B, ob.lo[o], ob.hi[o]);
    float diff[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float cb = fmaf(t, B[r] - A[r], A[r]);
        diff[r] = cb - fminf(fmaxf(cb, ob.lo[o][r]), ob.hi[o][r]);
        Cw[r] = fmaf(t, Q[r] - P[r], P[r]);
    }
    const float d2 = dot3(diff, diff);
    const float dist = sqrtf(d2);
    const float inv = d2 > 1e-24f ? rsqrtf(d2) : 0.f;
    float nb[3] = {diff[0] * inv, diff[1] * inv, diff[2] * inv};
    if (ob.has_rot[o]) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
            nrm[r] = fmaf(ob.R[o][3 * r + 2], nb[2], fmaf(ob.R[o][3 * r + 1], nb[1], ob.R[o][3 * r] * nb[0]));
    } else {
        nrm[0] = nb[0]; nrm[1] = nb[1]; nrm[2] = nb[2];
    }
    return dist - c_cap_table<M>.radius[c];
}

// d(distance)/dq for capsule c against an obstacle: g_d = n . v_d(Cw) for the joints that move the link
template <class M, int BLOCK>
__device__ __forceinline__ void env_capsule_gradient(const float* sm, int c, const float (&Cw)[3],
                                                     const float (&nrm)[3], float (&g)[M::NDOF]) {
    const int fc = c_cap_table<M>.frame[c];
    static_for<M::NDOF>([&](auto Dd) {
        constexpr int d = decltype(Dd)::value;
        constexpr int ci = chain_of_dof<M>(d);
        float gd = 0.f;
        if (ci < fc) {
            float a[3], o[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                a[r] = sm[(SmemLayout<M>::JOINTS + d * 6 + r) * BLOCK];
                o[r] = sm[(SmemLayout<M>::JOINTS + d * 6 + 3 + r) * BLOCK];
            }
            if constexpr (dof_is_prismatic<M>(d)) {
                gd = dot3(nrm, a);
            } else {
                const float rr[3] = {Cw[0] - o[0], Cw[1] - o[1], Cw[2] - o[2]};
                float v[3];
                cross3(a, rr, v);
                gd = dot3(nrm, v);
            }
        }
        g[d] = gd;
    });
}

}  // namespace cppflow
