// Unrolled forward kinematics over a compile-time chain (robots.cuh).
// Replaces jrl's Robot.forward_kinematics / Robot.jacobian (reference call sites: optimization_utils.py:811,
// optimization.py:74, optimization_utils.py:281) and the all-link FK inside Robot.self_collision_distances /
// env_collision_distances (collision_detection.py:40,65).
#pragma once
#include "robots.cuh"

namespace cppflow {

// c is a compile-time coefficient: skip zeros, avoid multiplies by +-1
template <int DUMMY = 0>
__device__ __forceinline__ float cfma(const float c, const float x, const float acc, const bool acc_is_zero) {
    (void)DUMMY;
    if (c == 0.f) return acc;
    if (acc_is_zero) return c == 1.f ? x : (c == -1.f ? -x : c * x);
    return c == 1.f ? acc + x : (c == -1.f ? acc - x : fmaf(c, x, acc));
}

// sin and cos of a joint angle: Cody-Waite reduction by pi/2 (two constants: exact to ~1e-15 for |x| < 64, far beyond
// any joint range) + the cephes single-precision minimax polynomials on [-pi/4, pi/4]; <= 2 ulp, 22 instructions and
// no slow path (libdevice sincosf carries a Payne-Hanek branch: ~120 SASS instructions per call site, and the FK
// chain has 7 of them - 14 KB of code in every kernel that walks the chain).
__device__ __forceinline__ void sincos_joint(float x, float& s, float& c) {
    const float kf = rintf(x * 0.636619772367581343f);  // x / (pi/2)
    const int k = __float2int_rn(kf);
    float r = fmaf(kf, -1.5707963705062866f, x);          // float32(pi/2)
    r = fmaf(kf, 4.371139000186241e-8f, r);               // pi/2 - float32(pi/2) = -4.371139e-8
    const float r2 = r * r;
    float ps = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(ps, r2, -1.6666654611e-1f);
    ps = fmaf(ps * r2, r, r);                              // sin(r)
    float pc = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(pc, r2, 4.166664568298827e-2f);
    pc = fmaf(pc * r2, r2, fmaf(-0.5f, r2, 1.f));          // cos(r)
    const float ss = (k & 1) ? pc : ps;
    const float cc = (k & 1) ? ps : pc;
    s = (k & 2) ? -ss : ss;
    c = ((k + 1) & 2) ? -cc : cc;
}

struct Frame {
    float R[9];  // row-major world rotation: column k = world direction of the local k axis
    float p[3];
};

// Walk the chain.  The sink receives
//   sink.joint(std::integral_constant<int,dof>, axis_world[3], origin_world[3])  before joint `dof` is applied
//   sink.frame(std::integral_constant<int,f>, const Frame&)                       for f = 0 (base) .. NCHAIN
template <class M, class Sink>
__device__ __forceinline__ void fk_chain(const float (&q)[M::NDOF], Sink& sink, Frame& F) {
    F.R[0] = 1.f; F.R[1] = 0.f; F.R[2] = 0.f;
    F.R[3] = 0.f; F.R[4] = 1.f; F.R[5] = 0.f;
    F.R[6] = 0.f; F.R[7] = 0.f; F.R[8] = 1.f;
    F.p[0] = F.p[1] = F.p[2] = 0.f;
    sink.frame(std::integral_constant<int, 0>{}, F);
    static_for<M::NCHAIN>([&](auto I) {
        constexpr int i = decltype(I)::value;
        // p += R * t_fixed
        static_for<3>([&](auto Kk) {
            constexpr int k = decltype(Kk)::value;
            constexpr float t = M::origin(i, k);
            if constexpr (t != 0.f) {
#pragma unroll
                for (int r = 0; r < 3; ++r) F.p[r] = fmaf(F.R[3 * r + k], t, F.p[r]);
            }
        });
        // R = R * Rfix
        constexpr bool ident = M::rfix(i, 0, 0) == 1.f && M::rfix(i, 1, 1) == 1.f && M::rfix(i, 2, 2) == 1.f;
        if constexpr (!ident) {
            float Rn[9];
            static_for<3>([&](auto Cc) {
                constexpr int c = decltype(Cc)::value;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    float acc = 0.f;
                    acc = cfma(M::rfix(i, 0, c), F.R[3 * r + 0], acc, true);
                    acc = cfma(M::rfix(i, 1, c), F.R[3 * r + 1], acc, M::rfix(i, 0, c) == 0.f);
                    acc = cfma(M::rfix(i, 2, c), F.R[3 * r + 2], acc, M::rfix(i, 0, c) == 0.f && M::rfix(i, 1, c) == 0.f);
                    Rn[3 * r + c] = acc;
                }
            });
#pragma unroll
            for (int k = 0; k < 9; ++k) F.R[k] = Rn[k];
        }
        if constexpr (M::jtype(i) != J_FIXED) {
            constexpr int d = dof_of_chain<M>(i);
            constexpr int ax = M::axis(i, 0) != 0.f ? 0 : (M::axis(i, 1) != 0.f ? 1 : 2);
            static_assert(M::axis(i, ax) == 1.f, "joint axes must be +x, +y or +z of the joint frame");
            float a[3] = {F.R[ax], F.R[3 + ax], F.R[6 + ax]};
            sink.joint(std::integral_constant<int, d>{}, a, F.p);
            if constexpr (M::jtype(i) == J_REVOLUTE) {
                float s, c;
                sincos_joint(q[d], s, c);
                // columns (u, v) rotate in the plane orthogonal to the axis: u' = c u + s v, v' = -s u + c v
                constexpr int u = (ax + 1) % 3, v = (ax + 2) % 3;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float cu = F.R[3 * r + u], cv = F.R[3 * r + v];
                    F.R[3 * r + u] = fmaf(c, cu, s * cv);
                    F.R[3 * r + v] = fmaf(c, cv, -s * cu);
                }
            } else {
#pragma unroll
                for (int r = 0; r < 3; ++r) F.p[r] = fmaf(a[r], q[d], F.p[r]);
            }
        }
        sink.frame(std::integral_constant<int, i + 1>{}, F);
    });
}

// world position of a point given in a link frame (compile-time coordinates fold)
template <class M, int C, int END>
__device__ __forceinline__ void capsule_endpoint(const Frame& F, float (&out)[3]) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float acc = F.p[r];
        acc = cfma(M::cap(C, 3 * END + 0), F.R[3 * r + 0], acc, false);
        acc = cfma(M::cap(C, 3 * END + 1), F.R[3 * r + 1], acc, false);
        acc = cfma(M::cap(C, 3 * END + 2), F.R[3 * r + 2], acc, false);
        out[r] = acc;
    }
}

// rotation matrix (row-major) -> quaternion wxyz; four-candidate method, the largest component is made positive
__device__ __forceinline__ void rotmat_to_quat(const float (&R)[9], float (&q)[4]) {
    const float m00 = R[0], m01 = R[1], m02 = R[2], m10 = R[3], m11 = R[4], m12 = R[5], m20 = R[6], m21 = R[7], m22 = R[8];
    const float t0 = fmaxf(1.f + m00 + m11 + m22, 0.f);
    const float t1 = fmaxf(1.f + m00 - m11 - m22, 0.f);
    const float t2 = fmaxf(1.f - m00 + m11 - m22, 0.f);
    const float t3 = fmaxf(1.f - m00 - m11 + m22, 0.f);
    float c0, c1, c2, c3, tm;
    if (t0 >= t1 && t0 >= t2 && t0 >= t3) {
        tm = t0; c0 = t0; c1 = m21 - m12; c2 = m02 - m20; c3 = m10 - m01;
    } else if (t1 >= t2 && t1 >= t3) {
        tm = t1; c0 = m21 - m12; c1 = t1; c2 = m10 + m01; c3 = m02 + m20;
    } else if (t2 >= t3) {
        tm = t2; c0 = m02 - m20; c1 = m10 + m01; c2 = t2; c3 = m12 + m21;
    } else {
        tm = t3; c0 = m10 - m01; c1 = m20 + m02; c2 = m21 + m12; c3 = t3;
    }
    const float inv = 0.5f * rsqrtf(tm);  // tm >= 1 for the selected candidate
    q[0] = c0 * inv; q[1] = c1 * inv; q[2] = c2 * inv; q[3] = c3 * inv;
}

// 6-vector pose error of optimization_utils.py:802-820: e[0:3] = rpy(q_target * conj(q_current)), e[3:6] = dt
__device__ __forceinline__ void pose_error(const float* __restrict__ tgt /*[7] xyz wxyz*/, const Frame& F, float (&e)[6]) {
    float qc[4];
    rotmat_to_quat(F.R, qc);
    const float w1 = tgt[3], x1 = tgt[4], y1 = tgt[5], z1 = tgt[6];
    const float w2 = qc[0], x2 = -qc[1], y2 = -qc[2], z2 = -qc[3];
    const float w = w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2;
    const float x = w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2;
    const float y = w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2;
    const float z = w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2;
    e[0] = atan2f(2.f * (w * x + y * z), 1.f - 2.f * (x * x + y * y));
    e[1] = asinf(fminf(fmaxf(2.f * (w * y - z * x), -1.f), 1.f));
    e[2] = atan2f(2.f * (w * z + x * y), 1.f - 2.f * (y * y + z * z));
    e[3] = tgt[0] - F.p[0];
    e[4] = tgt[1] - F.p[1];
    e[5] = tgt[2] - F.p[2];
}

__device__ __forceinline__ void cross3(const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// torch.remainder(x + pi, 2 pi) - pi in fp32 (evaluation_utils.py:153; search.py:124)
__device__ __forceinline__ float wrap_pi(float d) {
    const float PI_F = 3.14159274101257324f;      // float32(torch.pi)
    const float TWO_PI_F = 6.28318548202514648f;  // float32(2 * torch.pi)
    float m = fmodf(__fadd_rn(d, PI_F), TWO_PI_F);
    if (m < 0.f) m = __fadd_rn(m, TWO_PI_F);
    return __fsub_rn(m, PI_F);
}

}  // namespace cppflow
