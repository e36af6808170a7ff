// api.cu -- version / error strings / device probe of the C ABI (include/memflow_b200.h).
#include "common.cuh"

extern "C" int mf_abi_version(void) { return MF_ABI_VERSION; }

extern "C" const char* mf_error_string(int code) {
    if (code < 0) return cudaGetErrorString((cudaError_t)(-code));
    switch (code) {
    case MF_OK: return "ok";
    case MF_E_BADARG: return "bad argument (null pointer or negative size)";
    case MF_E_NFINAL: return "n_final outside [3, 8]";
    case MF_E_ALIGN: return "pointer not aligned for the vector width of this layout";
    case MF_E_SHAPE: return "bad shape (spline K/F, event_stride or order permutation)";
    case MF_E_NODEVICE: return "no CUDA device of compute capability 10.x";
    default: return "unknown memflow_b200 error";
    }
}

extern "C" int mf_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return MF_E_NODEVICE;
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return MF_E_NODEVICE;
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return MF_OK;
}
