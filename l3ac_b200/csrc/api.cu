#include "common.cuh"

extern "C" int l3ac_abi_version(void) { return 1; }

extern "C" const char* l3ac_error_string(int code) {
    switch (code) {
        case L3AC_OK: return "ok";
        case L3AC_EINVAL: return "invalid argument";
        case L3AC_EUNSUPPORTED: return "unsupported shape or configuration";
        case L3AC_EDRIVER: return "CUDA driver entry point unavailable";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown l3ac error";
    }
}
