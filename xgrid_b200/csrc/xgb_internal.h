// xgb_internal.h -- what the translation units of libxgrid_b200.so share (not part of the C ABI).
#ifndef XGB_INTERNAL_H
#define XGB_INTERNAL_H
#include <cuda_runtime.h>

#include "../../include/xgrid_b200.h"

namespace xgb_internal {
int fail(const char *msg);              // sets the thread's xgb_last_error text, returns 1
int require_init();                     // 0 once xgb_init has run, else fail(...)
cudaStream_t stream_of(xgb_handle h);   // handle 0 = the backend's compute stream
void count_launch();                    // one more kernel launched through this library
}  // namespace xgb_internal
#endif
