// dataflow.cuh -- persistent single-launch dataflow kernel for one large factor.
// (placeholder: filled in after the panel path is validated on the GPU)
#pragma once

#include "large.cuh"

namespace chub {

inline size_t dataflow_flag_bytes(int64_t n) { (void)n; return 0; }

template <typename T>
inline bool dataflow_supported(int64_t n, int64_t ld, int layout, const T *L) {
  (void)n; (void)ld; (void)layout; (void)L;
  return false;
}

template <typename T, int LAYOUT, bool UPDATE>
int launch_dataflow(T *L, int64_t n, int64_t ld, T *v, int flags, const LargeWs &ws,
                    int32_t *status, cudaStream_t st, int64_t *launches) {
  (void)L; (void)n; (void)ld; (void)v; (void)flags; (void)ws; (void)status; (void)st;
  *launches = 0;
  return (int)cudaErrorNotSupported;
}

}  // namespace chub
