// Closed-form capsule geometry: segment-segment and segment-box closest points.
// Replaces the capsule distance routines behind jrl's Robot.self_collision_distances /
// Robot.env_collision_distances (reference call sites: collision_detection.py:40,65;
// optimization_utils.py:652,690).  Negative distance = overlap.
#pragma once
#include <cuda_runtime.h>

namespace cppflow {

#define CPPFLOW_GEOM_EPS 1e-12f

__device__ __forceinline__ float dot3(const float* a, const float* b) {
    return fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0]));
}
__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }

// Closest points of segments P1Q1 and P2Q2: parameters s,t in [0,1] (clamped 2-variable box QP, exact).
__device__ __forceinline__ void segseg_closest(const float* P1, const float* Q1, const float* P2, const float* Q2,
                                               float& s, float& t) {
    const float d1[3] = {Q1[0] - P1[0], Q1[1] - P1[1], Q1[2] - P1[2]};
    const float d2[3] = {Q2[0] - P2[0], Q2[1] - P2[1], Q2[2] - P2[2]};
    const float r[3] = {P1[0] - P2[0], P1[1] - P2[1], P1[2] - P2[2]};
    const float a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r);
    const float c = dot3(d1, r), b = dot3(d1, d2);
    const float denom = fmaf(a, e, -b * b);
    const float inv_a = __fdividef(1.f, fmaxf(a, CPPFLOW_GEOM_EPS));
    const float inv_e = __fdividef(1.f, fmaxf(e, CPPFLOW_GEOM_EPS));
    float s0 = denom > CPPFLOW_GEOM_EPS ? clamp01(__fdividef(fmaf(b, f, -c * e), fmaxf(denom, CPPFLOW_GEOM_EPS))) : 0.f;
    float t0 = fmaf(b, s0, f) * inv_e;
    const float s_lo = clamp01(-c * inv_a);
    const float s_hi = clamp01((b - c) * inv_a);
    s = t0 < 0.f ? s_lo : (t0 > 1.f ? s_hi : s0);
    t = clamp01(t0);
}

// Parameter t in [0,1] of the point of segment AB closest to the axis-aligned box [lo,hi] (exact).
// h(t) = d . (P(t) - clamp(P(t), lo, hi)), half the derivative of the squared segment-box distance, is monotone
// piecewise linear with kinks where P(t) crosses a face plane: its root is bracketed between the kinks and interpolated.
// Along axis k the segment is inside the slab [lo_k, hi_k] for t in [a_k, b_k] (the two kinks of that axis), so
//     h(t) = sum_k w_k (t - clamp(t, a_k, b_k)),   w_k = d_k^2,
// which costs 4 instructions per axis (instead of 5 through P(t)) and makes the own-axis term of a kink exactly zero:
// h at a kink of axis k is a sum over the OTHER two axes.  (~195 instead of ~285 instructions per call; this routine
// is a third of the LM assembly kernel's instructions.)
__device__ __forceinline__ float segbox_closest(const float* A, const float* B, const float* lo, const float* hi) {
    float a[3], b[3], w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float dk = B[k] - A[k];
        const bool ok = fabsf(dk) > CPPFLOW_GEOM_EPS;  // a segment parallel to the slab adds nothing to h
        const float inv = __fdividef(1.f, ok ? dk : 1.f);
        const float ta = (lo[k] - A[k]) * inv, tb = (hi[k] - A[k]) * inv;
        a[k] = fminf(ta, tb);
        b[k] = fmaxf(ta, tb);
        w[k] = ok ? dk * dk : 0.f;
    }
    auto term = [&](int k, float t) { return w[k] * (t - fminf(fmaxf(t, a[k]), b[k])); };
    const float h0 = term(0, 0.f) + term(1, 0.f) + term(2, 0.f);
    const float h1 = term(0, 1.f) + term(1, 1.f) + term(2, 1.f);
    float t_lo = 0.f, h_lo = h0, t_hi = 1.f, h_hi = h1;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const float tb = side == 0 ? a[k] : b[k];
            const bool inside = tb > 0.f && tb < 1.f && w[k] > 0.f;
            const float hb = term(k1, tb) + term(k2, tb);
            const bool up_lo = inside && hb <= 0.f && tb > t_lo;
            const bool up_hi = inside && hb > 0.f && tb < t_hi;
            t_lo = up_lo ? tb : t_lo;
            h_lo = up_lo ? hb : h_lo;
            t_hi = up_hi ? tb : t_hi;
            h_hi = up_hi ? hb : h_hi;
        }
    }
    const float dh = h_hi - h_lo;
    const float frac = dh > CPPFLOW_GEOM_EPS ? __fdividef(-h_lo, fmaxf(dh, CPPFLOW_GEOM_EPS)) : 0.f;
    const float t_mid = fmaf(t_hi - t_lo, frac, t_lo);
    const float t = h0 >= 0.f ? 0.f : (h1 <= 0.f ? 1.f : t_mid);
    return clamp01(t);
}

// Cuboid obstacles, passed by value as a kernel parameter.  cuboid = [lo(3), hi(3)] in the frame Tcuboid
// (data_type_utils.py:109-127); R is Tcuboid[:3,:3] row-major, t = Tcuboid[:3,3].
struct Obstacles {
    int n;
    int has_rot[8];
    float lo[8][3];
    float hi[8][3];
    float R[8][9];
    float t[8][3];
};

// world point -> obstacle frame: Rb^T (p - tb)
__device__ __forceinline__ void to_box_frame(const Obstacles& ob, int o, const float* p, float* out) {
    const float v[3] = {p[0] - ob.t[o][0], p[1] - ob.t[o][1], p[2] - ob.t[o][2]};
    if (ob.has_rot[o]) {
#pragma unroll
        for (int c = 0; c < 3; ++c) out[c] = fmaf(ob.R[o][6 + c], v[2], fmaf(ob.R[o][3 + c], v[1], ob.R[o][c] * v[0]));
    } else {
        out[0] = v[0]; out[1] = v[1]; out[2] = v[2];
    }
}

}  // namespace cppflow
