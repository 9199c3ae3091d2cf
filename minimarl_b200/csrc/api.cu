// Status / error plumbing of the C ABI (include/minimarl_b200.h).
#include "common.cuh"

namespace {
thread_local const char* g_last_error = "no error";
}

namespace mm {
void set_last_error(cudaError_t e) { g_last_error = cudaGetErrorString(e); }
}  // namespace mm

extern "C" const char* mm_status_string(int status) {
    switch (status) {
        case MM_OK: return "ok";
        case MM_ERR_BAD_ARG: return "bad argument (null pointer, non-positive size or misaligned buffer)";
        case MM_ERR_LAUNCH: return "CUDA launch failed";
        case MM_ERR_UNSUPPORTED: return "shape not supported by this kernel";
        default: return "unknown status";
    }
}

extern "C" const char* mm_last_cuda_error(void) { return g_last_error; }

extern "C" int mm_version(void) { return 1; }
