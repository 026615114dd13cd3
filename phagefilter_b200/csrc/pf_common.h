// pf_common.h -- error plumbing shared by the libpfgpu translation units.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/pfgpu.h"

namespace pf {

void set_error(const char *fmt, ...);  // thread-local message returned by pf_last_error()

#define PF_CUDA_OK(expr)                                                                      \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            pf::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                          cudaGetErrorString(_e));                                            \
            return PF_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

constexpr uint32_t NONE32 = 0xFFFFFFFFu;

}  // namespace pf
