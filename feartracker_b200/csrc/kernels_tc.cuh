// tcgen05 / TMA kernels (placeholder until the tensor-core path lands).
#pragma once
#include <cuda_runtime.h>
namespace fear {
namespace tc {
inline int init() { return 0; }
inline bool available() { return false; }
inline bool pw_supported(int, int) { return false; }
inline int launch_pw(cudaStream_t, const float*, int, const float*, const float*, const float*, int, float*, int, int,
                     int, int, int) {
  return -1;
}
inline int launch_corr(cudaStream_t, const float*, int, float*, int) { return -1; }
}  // namespace tc
}  // namespace fear
