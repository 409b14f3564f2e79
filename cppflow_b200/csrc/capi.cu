// C-ABI plumbing: version, last-error text, robot tables.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace cppflow {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* msg) {
    std::strncpy(g_last_error, msg, sizeof(g_last_error) - 1);
    g_last_error[sizeof(g_last_error) - 1] = 0;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

template <class M>
static void fill_info(cppflow_robot_info* out) {
    std::memset(out, 0, sizeof(*out));
    out->ndof = M::NDOF;
    out->n_capsules = M::NCAP;
    out->n_pairs = M::NPAIR;
    out->n_chain = M::NCHAIN;
    for (int d = 0; d < M::NDOF; ++d) {
        out->lower[d] = dof_lower<M>(d);
        out->upper[d] = dof_upper<M>(d);
        out->is_prismatic[d] = dof_is_prismatic<M>(d) ? 1 : 0;
    }
    for (int c = 0; c < M::NCAP; ++c) {
        for (int k = 0; k < 7; ++k) out->capsules[c][k] = M::cap(c, k);
        out->capsule_frame[c] = M::cap_frame(c);
    }
    for (int p = 0; p < M::NPAIR; ++p) {
        out->pairs[p][0] = pair_cap<M>(p, 0);
        out->pairs[p][1] = pair_cap<M>(p, 1);
    }
    std::strncpy(out->name, M::name(), sizeof(out->name) - 1);
}

}  // namespace cppflow

using namespace cppflow;

extern "C" const char* cppflow_version(void) { return "cppflow_b200 0.2.0 (sm_100a)"; }

extern "C" int cppflow_abi_info(int64_t* out, int n) {
    const int64_t v[6] = {CPPFLOW_ABI_VERSION, (int64_t)sizeof(cppflow_lm_params), (int64_t)sizeof(cppflow_robot_info),
                          (int64_t)sizeof(cppflow_constraints), (int64_t)sizeof(cppflow_lm_loop_result),
                          (int64_t)sizeof(cppflow_lm_loop_job)};
    for (int k = 0; out && k < n && k < 6; ++k) out[k] = v[k];
    return 6;
}

extern "C" const char* cppflow_last_error(void) { return g_last_error; }

extern "C" int cppflow_robot_info_get(int robot, cppflow_robot_info* out) {
    CPPFLOW_CHECK_ARG(out != nullptr, "out");
    CPPFLOW_DISPATCH_ROBOT(robot, fill_info<M>(out));
    return CPPFLOW_OK;
}
