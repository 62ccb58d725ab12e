// Library-level entry points: version, error strings, device info.
#include "gr_common.cuh"

namespace gr {
static thread_local cudaError_t t_last_cuda = cudaSuccess;
void set_last_cuda_error(cudaError_t e) { t_last_cuda = e; }
int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kSmCountFallback;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = kSmCountFallback;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}
}  // namespace gr

extern "C" const char *gr_version(void) { return "gr_b200 0.1.0 (sm_100a)"; }

extern "C" const char *gr_error_string(int code) {
    switch (code) {
        case GR_OK: return "ok";
        case GR_ERR_INVALID: return "invalid argument";
        case GR_ERR_UNSUPPORTED: return "unsupported configuration";
        case GR_ERR_CUDA: return "CUDA runtime error";
        case GR_ERR_WORKSPACE: return "workspace too small";
        case GR_ERR_OVERFLOW: return "size does not fit the 32-bit CSR";
        default: return "unknown error";
    }
}

extern "C" const char *gr_last_cuda_error(void) { return cudaGetErrorString(gr::t_last_cuda); }

extern "C" int gr_device_info(int *sm_count_host, int *cc_major_host, int *cc_minor_host) {
    int dev = 0;
    GR_CUDA_CHECK(cudaGetDevice(&dev));
    if (sm_count_host) GR_CUDA_CHECK(cudaDeviceGetAttribute(sm_count_host, cudaDevAttrMultiProcessorCount, dev));
    if (cc_major_host) GR_CUDA_CHECK(cudaDeviceGetAttribute(cc_major_host, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_minor_host) GR_CUDA_CHECK(cudaDeviceGetAttribute(cc_minor_host, cudaDevAttrComputeCapabilityMinor, dev));
    return GR_OK;
}
