// cuda_utils.h — same error convention as the reference (src/cuda_utils.h:6-16): print
// "CUDA ERROR = <n> at file:line '<expr>'", cudaDeviceReset(), exit(99).  Also used for the
// non-zero return codes of the C ABI (include/ptcore.h), which carry their own message.
#pragma once

#include <cuda_runtime.h>

#include <cstdlib>
#include <iostream>

#define checkCudaErrors(val) check_cuda((val), #val, __FILE__, __LINE__)
inline void check_cuda(cudaError_t result, char const *const func, const char *const file, int const line) {
    if (result) {
        std::cerr << "CUDA ERROR = " << static_cast<unsigned int>(result) << " at " << file << ":" << line << " '" << func << "' \n";
        cudaDeviceReset();
        exit(99);
    }
}

#define checkPtcore(handle, val) check_ptcore((handle), (val), #val, __FILE__, __LINE__)
struct ptcore;
extern "C" const char *ptcore_last_error(const ptcore *h);
inline void check_ptcore(const ptcore *h, int result, char const *const func, const char *const file, int const line) {
    if (result) {
        std::cerr << "CUDA ERROR = " << static_cast<unsigned int>(result) << " at " << file << ":" << line << " '" << func << "' " << ptcore_last_error(h) << "\n";
        cudaDeviceReset();
        exit(99);
    }
}
