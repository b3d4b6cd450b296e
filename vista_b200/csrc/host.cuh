// Host-side helpers shared by the launchers: error handling and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace vb {

// last-error slot returned by b200v_last_error()
void set_error(const char* fmt, ...);
const char* last_error();
int device_sm_count();
// true the first time it is called with `flags` on the current device: function attributes
// (cudaFuncAttributeMaxDynamicSharedMemorySize ...) are per device, a process may drive several
inline bool first_use_on_device(bool (&flags)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}

#define VB_CHECK_CUDA(expr)                                                                        \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      vb::set_error("%s:%d CUDA error %d (%s) in %s", __FILE__, __LINE__, (int)_e, cudaGetErrorString(_e), #expr); \
      return 1;                                                                                    \
    }                                                                                              \
  } while (0)

#define VB_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      vb::set_error(__VA_ARGS__);    \
      return 2;                      \
    }                                \
  } while (0)

// Encodes a tiled fp16/bf16 tensor map with 128B swizzle and zero OOB fill.
// dims/box innermost first; strides_bytes has rank-1 entries (stride of dim 1..rank-1).
// Returns 0 on success.
int encode_tmap_16bit(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides, int bf16);

}  // namespace vb
