// Compile-time robot models for the three chains the reference plans for
// (planners.py:34-41: Panda, Fetch, FetchArm - all taken from jrl.robots, which is not in the reference tree).
// Chains restated from the public URDFs (fetch_description fetch.urdf, franka_description panda_arm_hand.urdf);
// joint limits for Fetch are pinned by the reference at tests/search_test.py:35-42.  The capsule tables
// ([x1,y1,z1,x2,y2,z2,r] in the link frame - the layout of robot._collision_capsules_by_link,
// collision_detection.py:137) are this repo's own approximation of the link geometry.
//
// Everything is constexpr so that the FK / Jacobian / capsule code unrolls per robot and folds zero and unit
// coefficients at compile time (identity origin rotations, coordinate-axis joints, capsule endpoints at the origin).
#pragma once
#include <cuda_runtime.h>
#include <utility>

namespace cppflow {

enum JType : int { J_FIXED = 0, J_REVOLUTE = 1, J_PRISMATIC = 2 };
enum RobotId : int { ROBOT_FETCH = 0, ROBOT_FETCH_ARM = 1, ROBOT_PANDA = 2, ROBOT_COUNT = 3 };

#define CPPFLOW_MAX_DOF 8
#define CPPFLOW_MAX_CHAIN 9
#define CPPFLOW_MAX_CAPS 10
#define CPPFLOW_MAX_PAIRS 28
#define CPPFLOW_MAX_OBSTACLES 8

#define HD __host__ __device__ constexpr

constexpr float kPi = 3.14159265358979323846f;
constexpr float kHalfSqrt2 = 0.70710678118654752440f;

// ----------------------------------------------------------------------------------------------------------
// Fetch (8 dof: prismatic torso + 7 revolute) and FetchArm (torso fixed at 0; data_type_utils.py:155-158)
template <bool TORSO_FIXED>
struct FetchT {
    static constexpr int ID = TORSO_FIXED ? ROBOT_FETCH_ARM : ROBOT_FETCH;
    static constexpr int NDOF = TORSO_FIXED ? 7 : 8;
    static constexpr int NCHAIN = 9;
    static constexpr int NCAP = 10;
    static constexpr int NPAIR = 28;
    static HD const char* name() { return TORSO_FIXED ? "fetch_arm" : "fetch"; }

    static HD int jtype(int i) {
        constexpr int t[NCHAIN] = {TORSO_FIXED ? J_FIXED : J_PRISMATIC, J_REVOLUTE, J_REVOLUTE, J_REVOLUTE, J_REVOLUTE,
                                   J_REVOLUTE, J_REVOLUTE, J_REVOLUTE, J_FIXED};
        return t[i];
    }
    // joint axis in the joint frame: torso z, pan z, lift y, roll x, flex y, roll x, flex y, roll x
    static HD float axis(int i, int k) {
        constexpr float a[NCHAIN][3] = {{0, 0, 1}, {0, 0, 1}, {0, 1, 0}, {1, 0, 0}, {0, 1, 0},
                                        {1, 0, 0}, {0, 1, 0}, {1, 0, 0}, {1, 0, 0}};
        return a[i][k];
    }
    static HD float origin(int i, int k) {
        constexpr float o[NCHAIN][3] = {{-0.086875f, 0.f, 0.37743f}, {0.119525f, 0.f, 0.34858f}, {0.117f, 0.f, 0.06f},
                                        {0.219f, 0.f, 0.f},          {0.133f, 0.f, 0.f},         {0.197f, 0.f, 0.f},
                                        {0.1245f, 0.f, 0.f},         {0.1385f, 0.f, 0.f},        {0.16645f, 0.f, 0.f}};
        return o[i][k];
    }
    // fixed origin rotation of chain element i (row r, col c): identity everywhere on the Fetch arm
    static HD float rfix(int, int r, int c) { return r == c ? 1.f : 0.f; }
    static HD float lower(int i) {
        constexpr float l[NCHAIN] = {0.f, -1.6056f, -1.221f, -kPi, -2.251f, -kPi, -2.16f, -kPi, 0.f};
        return l[i];
    }
    static HD float upper(int i) {
        constexpr float u[NCHAIN] = {0.38615f, 1.6056f, 1.518f, kPi, 2.251f, kPi, 2.16f, kPi, 0.f};
        return u[i];
    }
    // capsules: frame f = 0 for the base link, f = i+1 for the child link of chain element i
    static HD int cap_frame(int c) {
        constexpr int f[NCAP] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9};
        return f[c];
    }
    static HD int cap_sep(int c) { return cap_frame(c); }
    static HD float cap(int c, int k) {  // x1 y1 z1 x2 y2 z2 r
        constexpr float t[NCAP][7] = {
            {-0.02f, 0.f, 0.15f, -0.02f, 0.f, 0.22f, 0.27f},  // base_link
            {0.f, 0.f, 0.05f, 0.f, 0.f, 0.55f, 0.10f},        // torso_lift_link
            {0.f, 0.f, 0.f, 0.117f, 0.f, 0.06f, 0.07f},       // shoulder_pan_link
            {0.f, 0.f, 0.f, 0.219f, 0.f, 0.f, 0.065f},        // shoulder_lift_link
            {0.f, 0.f, 0.f, 0.133f, 0.f, 0.f, 0.06f},         // upperarm_roll_link
            {0.f, 0.f, 0.f, 0.197f, 0.f, 0.f, 0.06f},         // elbow_flex_link
            {0.f, 0.f, 0.f, 0.1245f, 0.f, 0.f, 0.055f},       // forearm_roll_link
            {0.f, 0.f, 0.f, 0.1385f, 0.f, 0.f, 0.055f},       // wrist_flex_link
            {0.f, 0.f, 0.f, 0.08f, 0.f, 0.f, 0.045f},         // wrist_roll_link
            {-0.05f, 0.f, 0.f, 0.03f, 0.f, 0.f, 0.07f},       // gripper_link
        };
        return t[c][k];
    }
};
using Fetch = FetchT<false>;
using FetchArm = FetchT<true>;

// ----------------------------------------------------------------------------------------------------------
// Panda: panda_link0 -> panda_hand (ros2/ros2_publisher.py:60-61), 7 revolute joints about local z,
// then the fixed flange (0,0,0.107) and the fixed hand joint rpy(0,0,-pi/4).
struct Panda {
    static constexpr int ID = ROBOT_PANDA;
    static constexpr int NDOF = 7;
    static constexpr int NCHAIN = 9;
    static constexpr int NCAP = 9;
    static constexpr int NPAIR = 21;
    static HD const char* name() { return "panda"; }

    static HD int jtype(int i) { return i < 7 ? J_REVOLUTE : J_FIXED; }
    static HD float axis(int, int k) { return k == 2 ? 1.f : 0.f; }
    static HD float origin(int i, int k) {
        constexpr float o[NCHAIN][3] = {{0.f, 0.f, 0.333f},      {0.f, 0.f, 0.f},   {0.f, -0.316f, 0.f},
                                        {0.0825f, 0.f, 0.f},     {-0.0825f, 0.384f, 0.f}, {0.f, 0.f, 0.f},
                                        {0.088f, 0.f, 0.f},      {0.f, 0.f, 0.107f}, {0.f, 0.f, 0.f}};
        return o[i][k];
    }
    // rpy(+pi/2,0,0) = [[1,0,0],[0,0,-1],[0,1,0]];  rpy(-pi/2,0,0) = [[1,0,0],[0,0,1],[0,-1,0]];
    // rpy(0,0,-pi/4) = [[c,s,0],[-s,c,0],[0,0,1]] with c = s = sqrt(2)/2
    static HD float rfix(int i, int r, int c) {
        constexpr float I3[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        constexpr float P[3][3] = {{1, 0, 0}, {0, 0, -1}, {0, 1, 0}};
        constexpr float N[3][3] = {{1, 0, 0}, {0, 0, 1}, {0, -1, 0}};
        constexpr float H[3][3] = {{kHalfSqrt2, kHalfSqrt2, 0}, {-kHalfSqrt2, kHalfSqrt2, 0}, {0, 0, 1}};
        return i == 0 ? I3[r][c]
             : i == 1 ? N[r][c]
             : i == 2 ? P[r][c]
             : i == 3 ? P[r][c]
             : i == 4 ? N[r][c]
             : i == 5 ? P[r][c]
             : i == 6 ? P[r][c]
             : i == 7 ? I3[r][c]
                      : H[r][c];
    }
    static HD float lower(int i) {
        constexpr float l[NCHAIN] = {-2.8973f, -1.7628f, -2.8973f, -3.0718f, -2.8973f, -0.0175f, -2.8973f, 0.f, 0.f};
        return l[i];
    }
    static HD float upper(int i) {
        constexpr float u[NCHAIN] = {2.8973f, 1.7628f, 2.8973f, -0.0698f, 2.8973f, 3.7525f, 2.8973f, 0.f, 0.f};
        return u[i];
    }
    static HD int cap_frame(int c) {
        constexpr int f[NCAP] = {0, 1, 2, 3, 4, 5, 6, 7, 9};
        return f[c];
    }
    // panda_hand is rigidly attached to panda_link7: it counts as position 8 for the pair rule
    static HD int cap_sep(int c) { return c == 8 ? 8 : cap_frame(c); }
    static HD float cap(int c, int k) {
        constexpr float t[NCAP][7] = {
            {-0.09f, 0.f, 0.06f, -0.06f, 0.f, 0.06f, 0.09f},  // panda_link0
            {0.f, 0.f, -0.30f, 0.f, 0.f, -0.05f, 0.07f},      // panda_link1
            {0.f, 0.f, -0.06f, 0.f, 0.f, 0.06f, 0.07f},       // panda_link2
            {0.f, 0.f, -0.22f, 0.f, 0.f, -0.07f, 0.07f},      // panda_link3
            {0.f, 0.f, -0.06f, 0.f, 0.f, 0.06f, 0.07f},       // panda_link4
            {0.f, 0.f, -0.30f, 0.f, 0.06f, -0.06f, 0.065f},   // panda_link5
            {0.f, 0.f, -0.07f, 0.f, 0.f, 0.01f, 0.06f},       // panda_link6
            {0.f, 0.f, -0.06f, 0.f, 0.f, 0.08f, 0.05f},       // panda_link7
            {0.f, -0.07f, 0.04f, 0.f, 0.07f, 0.04f, 0.05f},   // panda_hand
        };
        return t[c][k];
    }
};

// ----------------------------------------------------------------------------------------------------------
// derived compile-time tables

// chain index of actuated joint d
template <class M>
HD int chain_of_dof(int d) {
    int n = 0;
    for (int i = 0; i < M::NCHAIN; ++i) {
        if (M::jtype(i) != J_FIXED) {
            if (n == d) return i;
            ++n;
        }
    }
    return -1;
}
template <class M>
HD int dof_of_chain(int i) {
    int n = 0;
    for (int k = 0; k < i; ++k)
        if (M::jtype(k) != J_FIXED) ++n;
    return n;
}
template <class M>
HD bool dof_is_prismatic(int d) { return M::jtype(chain_of_dof<M>(d)) == J_PRISMATIC; }
template <class M>
HD float dof_lower(int d) { return M::lower(chain_of_dof<M>(d)); }
template <class M>
HD float dof_upper(int d) { return M::upper(chain_of_dof<M>(d)); }

// self-collision pair p -> (capsule a, capsule b): all pairs whose pair-rule positions are >= 3 apart
template <class M>
HD int pair_cap(int p, int which) {
    int n = 0;
    for (int a = 0; a < M::NCAP; ++a)
        for (int b = a + 1; b < M::NCAP; ++b)
            if (M::cap_sep(b) - M::cap_sep(a) >= 3) {
                if (n == p) return which == 0 ? a : b;
                ++n;
            }
    return -1;
}
template <class M>
HD int count_pairs() {
    int n = 0;
    for (int a = 0; a < M::NCAP; ++a)
        for (int b = a + 1; b < M::NCAP; ++b)
            if (M::cap_sep(b) - M::cap_sep(a) >= 3) ++n;
    return n;
}
static_assert(count_pairs<Fetch>() == Fetch::NPAIR, "Fetch pair count");
static_assert(count_pairs<FetchArm>() == FetchArm::NPAIR, "FetchArm pair count");
static_assert(count_pairs<Panda>() == Panda::NPAIR, "Panda pair count");

// compile-time loop: f(std::integral_constant<int, I>) for I in [0, N)
template <int... Is, class F>
__host__ __device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, F&& f) {
    (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
__host__ __device__ __forceinline__ void static_for(F&& f) {
    static_for_impl(std::make_integer_sequence<int, N>{}, static_cast<F&&>(f));
}

#undef HD
}  // namespace cppflow
