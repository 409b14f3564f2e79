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
// A capsule axis that passes through (or touches) a cuboid has no outward normal: the oracle returns a zero gradient
// when the axis-cuboid distance is <= 1e-12 in fp64.  In fp32 the closest point lands on a face plane up to rounding
// (|diff| ~ 1e-8), so the same rule needs a threshold above that noise: 1 micrometre.
#define CPPFLOW_AXIS_INSIDE_D2 1e-12f

// ----------------------------------------------------------------------------------------------------------------
// The per-thread shared-memory column: world capsule endpoints, joint origins and joint axes of ONE configuration (slot k
// of thread tid at sm[k * BLOCK + tid]: conflict-free).  On a serial arm most capsule endpoints ARE joint origins (a
// capsule starts at its link frame's origin = the origin of the joint before it, and ends at the origin of the joint
// after it): such an endpoint gets no slot of its own, its capsule reads the joint's origin slot.  The two values are
// the same bits - the FK adds the next element's fixed translation with the same fmaf sequence capsule_endpoint uses.
// Fetch: 13 of 20 endpoints alias (102 -> 69 floats per thread with joints, 60 -> 45 without); Panda: none.
template <class M>
struct SmemLayout {
    // dof whose joint origin IS endpoint e of capsule c, or -1
    static __host__ __device__ constexpr int alias(int c, int e) {
        const int f = M::cap_frame(c);
        const float x = M::cap(c, 3 * e), y = M::cap(c, 3 * e + 1), z = M::cap(c, 3 * e + 2);
        if (f >= 1 && f <= M::NCHAIN && M::jtype(f - 1) == J_REVOLUTE && x == 0.f && y == 0.f && z == 0.f)
            return dof_of_chain<M>(f - 1);  // the link frame's origin: a revolute joint does not move it
        if (f < M::NCHAIN && M::jtype(f) != J_FIXED && x == M::origin(f, 0) && y == M::origin(f, 1) && z == M::origin(f, 2))
            return dof_of_chain<M>(f);      // the next joint's origin (taken before a prismatic joint's displacement)
        return -1;
    }
    static __host__ __device__ constexpr int n_own() {
        int n = 0;
        for (int c = 0; c < M::NCAP; ++c)
            for (int e = 0; e < 2; ++e) n += alias(c, e) < 0 ? 1 : 0;
        return n;
    }
    static constexpr bool ANY_ALIAS = n_own() < 2 * M::NCAP;
    static constexpr int ORIGINS = n_own() * 3;              // joint d origin r: ORIGINS + d*3 + r
    static constexpr int AXES = ORIGINS + M::NDOF * 3;       // joint d axis r:   AXES + d*3 + r
    static constexpr int N_DIST = ANY_ALIAS ? AXES : ORIGINS;  // floats per thread when only distances are needed
    static constexpr int N_FULL = AXES + M::NDOF * 3;
    // slot of the x coordinate of endpoint e of capsule c
    static __host__ __device__ constexpr int slot(int c, int e) {
        const int a = alias(c, e);
        if (a >= 0) return ORIGINS + a * 3;
        int n = 0;
        for (int cc = 0; cc < M::NCAP; ++cc)
            for (int ee = 0; ee < 2; ++ee) {
                if (cc == c && ee == e) return n * 3;
                n += alias(cc, ee) < 0 ? 1 : 0;
            }
        return -1;
    }
    static __host__ __device__ constexpr int slots(int c) { return slot(c, 0) | (slot(c, 1) << 16); }  // P | Q << 16
};
static_assert(SmemLayout<Fetch>::N_FULL == 69 && SmemLayout<Fetch>::N_DIST == 45, "Fetch: 13 of 20 endpoints are joint origins");
static_assert(!SmemLayout<Panda>::ANY_ALIAS && SmemLayout<Panda>::N_DIST == 54, "Panda: every endpoint has its own slot");

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
    int slots[M::NCAP];  // SmemLayout::slots
    float radius[M::NCAP];
    float reach[M::NCAP];  // half-length + radius + margin
};
template <class M>
__host__ __device__ constexpr CapTable<M> make_cap_table() {
    CapTable<M> t{};
    for (int c = 0; c < M::NCAP; ++c) {
        t.frame[c] = M::cap_frame(c);
        t.slots[c] = SmemLayout<M>::slots(c);
        t.radius[c] = M::cap(c, 6);
        t.reach[c] = M::cap(c, 6) + cap_half_length<M>(c) + CPPFLOW_CULL_MARGIN;
    }
    return t;
}
template <class M>
__constant__ PairTable<M> c_pair_table = make_pair_table<M>();
template <class M>
__constant__ CapTable<M> c_cap_table = make_cap_table<M>();

// shared by the FK sinks: joint origins go to the column when anything reads them (Jacobian columns, or capsule
// endpoints that alias them), axes only with WITH_JOINTS; a capsule endpoint is stored unless it aliases a joint origin
template <class M, int BLOCK, bool WITH_JOINTS, int D>
__device__ __forceinline__ void store_joint(float* sm, const float* axis, const float* origin) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        if constexpr (WITH_JOINTS) sm[(SmemLayout<M>::AXES + D * 3 + r) * BLOCK] = axis[r];
        if constexpr (WITH_JOINTS || SmemLayout<M>::ANY_ALIAS) sm[(SmemLayout<M>::ORIGINS + D * 3 + r) * BLOCK] = origin[r];
    }
}
template <class M, int BLOCK, int C>
__device__ __forceinline__ void store_capsule(float* sm, const float (&P)[3], const float (&Q)[3]) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        if constexpr (SmemLayout<M>::alias(C, 0) < 0) sm[(SmemLayout<M>::slot(C, 0) + r) * BLOCK] = P[r];
        if constexpr (SmemLayout<M>::alias(C, 1) < 0) sm[(SmemLayout<M>::slot(C, 1) + r) * BLOCK] = Q[r];
    }
}

// FK sink writing world capsule endpoints (and optionally joint axes/origins) into the thread's smem column
template <class M, int BLOCK, bool WITH_JOINTS>
struct CollisionSink {
    float* sm;  // smem base + tid
    template <int D>
    __device__ __forceinline__ void joint(std::integral_constant<int, D>, const float* axis, const float* origin) {
        store_joint<M, BLOCK, WITH_JOINTS, D>(sm, axis, origin);
    }
    template <int F>
    __device__ __forceinline__ void frame(std::integral_constant<int, F>, const Frame& fr) {
        static_for<M::NCAP>([&](auto Cc) {
            constexpr int c = decltype(Cc)::value;
            if constexpr (M::cap_frame(c) == F) {
                float P[3], Q[3];
                capsule_endpoint<M, c, 0>(fr, P);
                capsule_endpoint<M, c, 1>(fr, Q);
                store_capsule<M, BLOCK, c>(sm, P, Q);
            }
        });
    }
};

// endpoints of a capsule from its slot word (SmemLayout::slots: P | Q << 16)
template <int BLOCK>
__device__ __forceinline__ void load_capsule(const float* sm, int slots, float (&P)[3], float (&Q)[3]) {
    const int sp = slots & 0xffff, sq = slots >> 16;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        P[r] = sm[(sp + r) * BLOCK];
        Q[r] = sm[(sq + r) * BLOCK];
    }
}

// ---- shared cores: the arithmetic of one capsule pair / capsule-cuboid test with its table entries passed in, so that
// the API kernels (compile-time tables in constant memory) and the active-set kernels (tables in shared memory, runtime
// index per lane) run the same code.

// signed distance of capsules a and b; C2 = closest point on the axis of b, nrm = unit vector from it to the closest
// point on the axis of a
template <int BLOCK>
__device__ __forceinline__ float capsule_pair_core(const float* sm, int slots_a, int slots_b, float rsum, float (&C2)[3],
                                                   float (&nrm)[3]) {
    float P1[3], Q1[3], P2[3], Q2[3];
    load_capsule<BLOCK>(sm, slots_a, P1, Q1);
    load_capsule<BLOCK>(sm, slots_b, P2, Q2);
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
    return dist - rsum;
}

// d(distance)/dq of a capsule pair: only the joints between the two links (frames fa < fb) contribute: g_d = -n . v_d(C2)
template <class M, int BLOCK>
__device__ __forceinline__ void capsule_pair_gradient_core(const float* sm, int fa, int fb, const float (&C2)[3],
                                                           const float (&nrm)[3], float (&g)[M::NDOF]) {
    static_for<M::NDOF>([&](auto Dd) {
        constexpr int d = decltype(Dd)::value;
        constexpr int ci = chain_of_dof<M>(d);
        float gd = 0.f;
        if (ci >= fa && ci < fb) {
            float a[3], o[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                a[r] = sm[(SmemLayout<M>::AXES + d * 3 + r) * BLOCK];
                o[r] = sm[(SmemLayout<M>::ORIGINS + d * 3 + r) * BLOCK];
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

// signed distance of capsule c to obstacle o, endpoints A / B already in the cuboid's frame; Cw = closest point on the
// capsule axis (world), nrm = world normal
template <int BLOCK>
__device__ __forceinline__ float capsule_cuboid_core(const float (&P)[3], const float (&Q)[3], const float (&A)[3],
                                                     const float (&B)[3], const Obstacles& ob, int o, float radius,
                                                     float (&Cw)[3], float (&nrm)[3]) {
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
    const float inv = d2 > CPPFLOW_AXIS_INSIDE_D2 ? rsqrtf(d2) : 0.f;
    const float nb[3] = {diff[0] * inv, diff[1] * inv, diff[2] * inv};
    if (ob.has_rot[o]) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
            nrm[r] = fmaf(ob.R[o][3 * r + 2], nb[2], fmaf(ob.R[o][3 * r + 1], nb[1], ob.R[o][3 * r] * nb[0]));
    } else {
        nrm[0] = nb[0]; nrm[1] = nb[1]; nrm[2] = nb[2];
    }
    return dist - radius;
}

// d(distance)/dq of a capsule on link frame fc against a world-fixed obstacle: g_d = n . v_d(Cw) for the joints before fc
template <class M, int BLOCK>
__device__ __forceinline__ void capsule_cuboid_gradient_core(const float* sm, int fc, const float (&Cw)[3],
                                                             const float (&nrm)[3], float (&g)[M::NDOF]) {
    static_for<M::NDOF>([&](auto Dd) {
        constexpr int d = decltype(Dd)::value;
        constexpr int ci = chain_of_dof<M>(d);
        float gd = 0.f;
        if (ci < fc) {
            float a[3], o[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                a[r] = sm[(SmemLayout<M>::AXES + d * 3 + r) * BLOCK];
                o[r] = sm[(SmemLayout<M>::ORIGINS + d * 3 + r) * BLOCK];
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

// ---- API kernels (k_collision.cu distance / Jacobian entry points, single-path metrics): compile-time tables.
// Pairs whose bounding spheres prove distance > cull_thr are skipped and +INFINITY is returned (cull_thr = INFINITY:
// always the exact distance).
template <class M, int BLOCK>
__device__ __forceinline__ float self_pair_distance(const float* sm, int p, float (&C2)[3], float (&nrm)[3],
                                                    float cull_thr = INFINITY) {
    const int a = c_cap_table<M>.slots[c_pair_table<M>.a[p]], b = c_cap_table<M>.slots[c_pair_table<M>.b[p]];
    {
        float P1[3], Q1[3], P2[3], Q2[3];
        load_capsule<BLOCK>(sm, a, P1, Q1);
        load_capsule<BLOCK>(sm, b, P2, Q2);
        const float mx = (P1[0] + Q1[0]) - (P2[0] + Q2[0]);
        const float my = (P1[1] + Q1[1]) - (P2[1] + Q2[1]);
        const float mz = (P1[2] + Q1[2]) - (P2[2] + Q2[2]);
        const float lim = cull_thr + c_pair_table<M>.reach[p];  // > 0 whenever a cull is intended
        if (0.25f * (mx * mx + my * my + mz * mz) > lim * lim && lim > 0.f) return INFINITY;
    }
    return capsule_pair_core<BLOCK>(sm, a, b, c_pair_table<M>.rsum[p], C2, nrm);
}
template <class M, int BLOCK>
__device__ __forceinline__ void self_pair_gradient(const float* sm, int p, const float (&C2)[3], const float (&nrm)[3],
                                                   float (&g)[M::NDOF]) {
    capsule_pair_gradient_core<M, BLOCK>(sm, c_pair_table<M>.fa[p], c_pair_table<M>.fb[p], C2, nrm, g);
}
template <class M, int BLOCK>
__device__ __forceinline__ float env_capsule_distance(const float* sm, int c, const Obstacles& ob, int o,
                                                      float (&Cw)[3], float (&nrm)[3], float cull_thr = INFINITY) {
    float P[3], Q[3], A[3], B[3];
    load_capsule<BLOCK>(sm, c_cap_table<M>.slots[c], P, Q);
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
    return capsule_cuboid_core<BLOCK>(P, Q, A, B, ob, o, c_cap_table<M>.radius[c], Cw, nrm);
}
template <class M, int BLOCK>
__device__ __forceinline__ void env_capsule_gradient(const float* sm, int c, const float (&Cw)[3],
                                                     const float (&nrm)[3], float (&g)[M::NDOF]) {
    capsule_cuboid_gradient_core<M, BLOCK>(sm, c_cap_table<M>.frame[c], Cw, nrm, g);
}


// ----------------------------------------------------------------------------------------------------------------
// Two-phase active-set evaluation (used where only the SIGN of a distance matters: the collision flags, the LM
// assembly and the validity metrics).
//   phase 1: every pair is tested against its conservative bounding-sphere bound with the capsule midpoints held in
//            registers, fully unrolled (compile-time pair table, ~9 instructions per pair) -> bitmask of survivors;
//   phase 2: `while (mask)` over the survivors with a runtime pair index: exact closed-form distance from the
//            endpoints in the thread's shared-memory column.  Lanes of a warp walk their own survivor lists in
//            lock-step, so a warp pays max-over-lanes(#survivors) evaluations, not the union.
// Tables indexed by a per-lane runtime index live in shared memory (divergent constant-bank reads serialise).

template <class M>
struct CollTables {
    int pair_ab[M::NPAIR];      // a | b << 8 | frame(a) << 16 | frame(b) << 24
    float pair_rsum[M::NPAIR];
    int cap_frame[M::NCAP];
    int cap_slots[M::NCAP];     // SmemLayout::slots
    float cap_radius[M::NCAP];
    Obstacles ob;
};

template <class M>
__device__ __forceinline__ void fill_coll_tables(CollTables<M>& tb, const Obstacles& ob, int tid, int nthreads) {
    for (int p = tid; p < M::NPAIR; p += nthreads) {
        tb.pair_ab[p] = c_pair_table<M>.a[p] | (c_pair_table<M>.b[p] << 8) | (c_pair_table<M>.fa[p] << 16) |
                        (c_pair_table<M>.fb[p] << 24);
        tb.pair_rsum[p] = c_pair_table<M>.rsum[p];
    }
    for (int c = tid; c < M::NCAP; c += nthreads) {
        tb.cap_frame[c] = c_cap_table<M>.frame[c];
        tb.cap_slots[c] = c_cap_table<M>.slots[c];
        tb.cap_radius[c] = c_cap_table<M>.radius[c];
    }
    const int* src = reinterpret_cast<const int*>(&ob);
    int* dst = reinterpret_cast<int*>(&tb.ob);
    for (int k = tid; k < (int)(sizeof(Obstacles) / sizeof(int)); k += nthreads) dst[k] = src[k];
}

// FK sink that additionally keeps TWICE the capsule midpoints (P + Q) in registers for the unrolled culls
template <class M, int BLOCK, bool WITH_JOINTS>
struct MidSink {
    float* sm;
    float mid2[M::NCAP][3];
    template <int D>
    __device__ __forceinline__ void joint(std::integral_constant<int, D>, const float* axis, const float* origin) {
        store_joint<M, BLOCK, WITH_JOINTS, D>(sm, axis, origin);
    }
    template <int F>
    __device__ __forceinline__ void frame(std::integral_constant<int, F>, const Frame& fr) {
        static_for<M::NCAP>([&](auto Cc) {
            constexpr int c = decltype(Cc)::value;
            if constexpr (M::cap_frame(c) == F) {
                float P[3], Q[3];
                capsule_endpoint<M, c, 0>(fr, P);
                capsule_endpoint<M, c, 1>(fr, Q);
                store_capsule<M, BLOCK, c>(sm, P, Q);
#pragma unroll
                for (int r = 0; r < 3; ++r) mid2[c][r] = P[r] + Q[r];
            }
        });
    }
};

template <class M>
__host__ __device__ constexpr float pair_reach(int p) {
    const int a = pair_cap<M>(p, 0), b = pair_cap<M>(p, 1);
    return M::cap(a, 6) + M::cap(b, 6) + cap_half_length<M>(a) + cap_half_length<M>(b) + CPPFLOW_CULL_MARGIN;
}
template <class M>
__host__ __device__ constexpr float cap_reach(int c) {
    return M::cap(c, 6) + cap_half_length<M>(c) + CPPFLOW_CULL_MARGIN;
}

// bit p set <=> self-collision pair p may have distance <= 0 (not proven apart by the bounding spheres)
template <class M>
__device__ __forceinline__ unsigned self_cull_mask(const float (&mid2)[M::NCAP][3]) {
    static_assert(M::NPAIR <= 32, "pair mask is 32 bits");
    unsigned mask = 0u;
    static_for<M::NPAIR>([&](auto Pp) {
        constexpr int p = decltype(Pp)::value;
        constexpr int a = pair_cap<M>(p, 0), b = pair_cap<M>(p, 1);
        constexpr float lim = pair_reach<M>(p);
        constexpr float lim2x4 = 4.f * lim * lim;  // the midpoints are doubled
        const float mx = mid2[a][0] - mid2[b][0], my = mid2[a][1] - mid2[b][1], mz = mid2[a][2] - mid2[b][2];
        const float dd = fmaf(mz, mz, fmaf(my, my, mx * mx));
        mask |= (dd > lim2x4) ? 0u : (1u << p);
    });
    return mask;
}

// number of leading capsules attached to links that no actuated joint moves (capsules are ordered along the chain):
// their distance to a world-fixed obstacle does not depend on q, so they contribute nothing to the LM normal equations
template <class M>
__host__ __device__ constexpr int n_static_capsules() {
    int first_moving_frame = M::NCHAIN + 1;
    for (int i = M::NCHAIN - 1; i >= 0; --i)
        if (M::jtype(i) != J_FIXED) first_moving_frame = i + 1;  // joint i moves frames > i
    int n = 0;
    while (n < M::NCAP && M::cap_frame(n) < first_moving_frame) ++n;
    return n;
}

static_assert(n_static_capsules<Fetch>() == 1 && n_static_capsules<FetchArm>() == 2 && n_static_capsules<Panda>() == 1,
              "base link (and the fixed torso of fetch_arm) are the static capsules");

// bit c set <=> capsule c may touch obstacle o; capsules below FIRST are not tested.
// Per axis the (doubled) distance of the capsule midpoint to the cuboid is max(|m - centre| - half extent, 0), evaluated
// as ONE saturating add (FADD.SAT with the |.| source modifier): 3 FMA-pipe instructions per axis and none on the
// half-rate ALU pipe, against FMNMX + FMNMX + FADD + FFMA for m - clamp(m, lo, hi).  Saturation at 1 is harmless where
// the capsule's threshold is below 1 (an axis that saturates culls the pair, as its true value would); the few capsules
// with a larger reach (a robot's base) take max(., 0) instead.
template <bool SAT>
__device__ __forceinline__ float positive_part(float x) {
    if constexpr (SAT) return __saturatef(x);
    else return fmaxf(x, 0.f);
}
template <class M, int FIRST = 0>
__device__ __forceinline__ unsigned env_cull_mask(const float (&mid2)[M::NCAP][3], const Obstacles& ob, int o) {
    unsigned mask = 0u;
    const float t2[3] = {2.f * ob.t[o][0], 2.f * ob.t[o][1], 2.f * ob.t[o][2]};
    // doubled centre / half extent of the cuboid in its own frame
    const float cb[3] = {ob.lo[o][0] + ob.hi[o][0], ob.lo[o][1] + ob.hi[o][1], ob.lo[o][2] + ob.hi[o][2]};
    const float hb[3] = {ob.hi[o][0] - ob.lo[o][0], ob.hi[o][1] - ob.lo[o][1], ob.hi[o][2] - ob.lo[o][2]};
    if (ob.has_rot[o]) {
        static_for<M::NCAP - FIRST>([&](auto Cc) {
            constexpr int c = decltype(Cc)::value + FIRST;
            constexpr float lim = cap_reach<M>(c);
            constexpr float lim2x4 = 4.f * lim * lim;
            constexpr bool SAT = lim2x4 < 1.f;
            const float v[3] = {mid2[c][0] - t2[0], mid2[c][1] - t2[1], mid2[c][2] - t2[2]};
            float dd = 0.f;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float m = fmaf(ob.R[o][6 + r], v[2], fmaf(ob.R[o][3 + r], v[1], ob.R[o][r] * v[0]));
                const float e = positive_part<SAT>(fabsf(m - cb[r]) - hb[r]);
                dd = fmaf(e, e, dd);
            }
            mask |= (dd > lim2x4) ? 0u : (1u << c);
        });
    } else {
        const float cw[3] = {cb[0] + t2[0], cb[1] + t2[1], cb[2] + t2[2]};
        static_for<M::NCAP - FIRST>([&](auto Cc) {
            constexpr int c = decltype(Cc)::value + FIRST;
            constexpr float lim = cap_reach<M>(c);
            constexpr float lim2x4 = 4.f * lim * lim;
            constexpr bool SAT = lim2x4 < 1.f;
            float dd = 0.f;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float e = positive_part<SAT>(fabsf(mid2[c][r] - cw[r]) - hb[r]);
                dd = fmaf(e, e, dd);
            }
            mask |= (dd > lim2x4) ? 0u : (1u << c);
        });
    }
    return mask;
}

// ---- active-set kernels: runtime pair / capsule index per lane, tables from shared memory
template <class M, int BLOCK>
__device__ __forceinline__ float self_pair_exact(const float* sm, const CollTables<M>& tb, int p, float (&C2)[3],
                                                 float (&nrm)[3]) {
    const int ab = tb.pair_ab[p];
    return capsule_pair_core<BLOCK>(sm, tb.cap_slots[ab & 0xff], tb.cap_slots[(ab >> 8) & 0xff], tb.pair_rsum[p], C2, nrm);
}
template <class M, int BLOCK>
__device__ __forceinline__ void self_pair_gradient_rt(const float* sm, const CollTables<M>& tb, int p,
                                                      const float (&C2)[3], const float (&nrm)[3], float (&g)[M::NDOF]) {
    const int ab = tb.pair_ab[p];
    capsule_pair_gradient_core<M, BLOCK>(sm, (ab >> 16) & 0xff, (ab >> 24) & 0xff, C2, nrm, g);
}
template <class M, int BLOCK>
__device__ __forceinline__ float env_capsule_exact(const float* sm, const CollTables<M>& tb, int c, int o,
                                                   float (&Cw)[3], float (&nrm)[3]) {
    float P[3], Q[3], A[3], B[3];
    load_capsule<BLOCK>(sm, tb.cap_slots[c], P, Q);
    to_box_frame(tb.ob, o, P, A);
    to_box_frame(tb.ob, o, Q, B);
    return capsule_cuboid_core<BLOCK>(P, Q, A, B, tb.ob, o, tb.cap_radius[c], Cw, nrm);
}
template <class M, int BLOCK>
__device__ __forceinline__ void env_capsule_gradient_rt(const float* sm, const CollTables<M>& tb, int c,
                                                        const float (&Cw)[3], const float (&nrm)[3],
                                                        float (&g)[M::NDOF]) {
    capsule_cuboid_gradient_core<M, BLOCK>(sm, tb.cap_frame[c], Cw, nrm, g);
}

}  // namespace cppflow
