// Spatial partition of the GPU's SMs into two CUDA green contexts (driver API, CUDA 12.4+), with streams in each.
//
// Why: in the chunk-pipelined LM iteration (pipeline.ResidentPipeline) the block solve of one chunk runs while another
// chunk is in its assembly.  Sharing SMs (CPPFLOW_LM_OVERLAP) works, but a 4-warp solve CTA displaces an 8-warp
// assembly CTA for its whole life and is itself slowed down by the assembly warps it shares the schedulers with
// (0.345 ms instead of 0.245).  With a partition the solves own a few SMs and run at full speed while the assembly
// owns the rest.  The driver entry points are resolved at run time (cudaGetDriverEntryPoint): the library does not link
// against libcuda.  No reference counterpart.
#include <cuda.h>

#include "common.cuh"

using namespace cppflow;

namespace {

template <class F>
bool resolve(const char* name, F& fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return false;
    fn = reinterpret_cast<F>(p);
    return true;
}

}  // namespace

extern "C" int cppflow_sm_partition_create(int device, int min_sms_first, int n_streams_first, int n_streams_second,
                                           void** streams_first, void** streams_second, int* sms_first, int* sms_second) {
    CPPFLOW_CHECK_ARG(min_sms_first > 0 && n_streams_first >= 0 && n_streams_second >= 0, "sizes");
    CPPFLOW_CHECK_ARG((n_streams_first == 0 || streams_first) && (n_streams_second == 0 || streams_second), "stream arrays");
    CUresult (*p_cuDeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*p_cuDeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
    CUresult (*p_cuDevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int,
                                              unsigned int) = nullptr;
    CUresult (*p_cuDevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
    CUresult (*p_cuGreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    CUresult (*p_cuGreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
    if (!(resolve("cuDeviceGet", p_cuDeviceGet) && resolve("cuDeviceGetDevResource", p_cuDeviceGetDevResource) &&
          resolve("cuDevSmResourceSplitByCount", p_cuDevSmResourceSplitByCount) &&
          resolve("cuDevResourceGenerateDesc", p_cuDevResourceGenerateDesc) && resolve("cuGreenCtxCreate", p_cuGreenCtxCreate) &&
          resolve("cuGreenCtxStreamCreate", p_cuGreenCtxStreamCreate)))
        return fail(CPPFLOW_E_CUDA, "sm_partition: the driver does not provide the green-context entry points");
    if (cudaSetDevice(device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess)  // primary context initialised and current
        return fail(CPPFLOW_E_CUDA, "sm_partition: cannot initialise device %d", device);
#define DRV_OK(expr)                                                                             \
    do {                                                                                         \
        CUresult r_ = (expr);                                                                    \
        if (r_ != CUDA_SUCCESS) return fail(CPPFLOW_E_CUDA, "sm_partition: %s failed (CUresult %d)", #expr, (int)r_); \
    } while (0)
    CUdevice dev;
    DRV_OK(p_cuDeviceGet(&dev, device));
    CUdevResource all, first, rest;
    DRV_OK(p_cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
    unsigned int n_groups = 1;
    DRV_OK(p_cuDevSmResourceSplitByCount(&first, &n_groups, &all, &rest, 0, (unsigned)min_sms_first));
    if (n_groups < 1 || rest.type != CU_DEV_RESOURCE_TYPE_SM || rest.sm.smCount == 0)
        return fail(CPPFLOW_E_INVALID, "sm_partition: cannot split %u SMs into %d + rest", all.sm.smCount, min_sms_first);
    CUdevResourceDesc d_first, d_rest;
    DRV_OK(p_cuDevResourceGenerateDesc(&d_first, &first, 1));
    DRV_OK(p_cuDevResourceGenerateDesc(&d_rest, &rest, 1));
    CUgreenCtx g_first, g_rest;  // live until the process ends (a handful per process)
    DRV_OK(p_cuGreenCtxCreate(&g_first, d_first, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    DRV_OK(p_cuGreenCtxCreate(&g_rest, d_rest, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    for (int i = 0; i < n_streams_first; ++i) {
        CUstream s;
        DRV_OK(p_cuGreenCtxStreamCreate(&s, g_first, CU_STREAM_NON_BLOCKING, 0));
        streams_first[i] = (void*)s;
    }
    for (int i = 0; i < n_streams_second; ++i) {
        CUstream s;
        DRV_OK(p_cuGreenCtxStreamCreate(&s, g_rest, CU_STREAM_NON_BLOCKING, 0));
        streams_second[i] = (void*)s;
    }
    if (sms_first) *sms_first = (int)first.sm.smCount;
    if (sms_second) *sms_second = (int)rest.sm.smCount;
#undef DRV_OK
    return CPPFLOW_OK;
}
