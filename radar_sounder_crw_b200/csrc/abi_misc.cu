// abi_misc.cu -- version / error-string entry points of libcrw_b200.so.
#include "common.cuh"

extern "C" int crw_version(void) { return 100; }   // 0.1.0

extern "C" int crw_built_arch(void) { return 100; }  // sm_100a

extern "C" const char* crw_error_string(int code) {
    switch (code) {
        case CRW_OK: return "ok";
        case CRW_ERR_INVALID: return "invalid argument (shape, parameter or null pointer)";
        case CRW_ERR_UNSUPPORTED: return "unsupported configuration for this build";
        case CRW_ERR_ALIGN: return "pointer not 16-byte aligned or C not a multiple of 4";
        case CRW_ERR_WORKSPACE: return "workspace missing or too small";
        default: break;
    }
    if (code <= CRW_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(CRW_ERR_CUDA_BASE - code));
    return "unknown error";
}
