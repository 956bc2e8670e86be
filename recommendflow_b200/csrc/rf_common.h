// rf_common.h -- error plumbing shared by the .cu files behind include/rf_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <string>

#include "../../include/rf_b200.h"

namespace rf {

inline std::string &last_error_ref() {
    static thread_local std::string msg;
    return msg;
}

inline int set_error(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}

#define RF_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t rf_e_ = (expr);                                                                \
        if (rf_e_ != cudaSuccess)                                                                  \
            return ::rf::set_error(RF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(rf_e_), \
                                   __FILE__, __LINE__);                                            \
    } while (0)

}  // namespace rf
