// sr_host.h -- host-side helpers shared by the C-ABI translation units: error reporting, the
// cuTensorMapEncodeTiled driver entry point (resolved at run time, no link-time libcuda dependency),
// SM count cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "sr_common.cuh"

namespace sr {

void set_last_error(const char* fmt, ...);
const char* get_last_error();

int sm_count();            // SMs of the current device (cached per device)
int max_smem_optin();      // max dynamic shared memory per block (opt-in)

// Encode a bf16 tiled tensor map.  rank 2 or 3; dims/strides innermost first; strides in BYTES for
// dims 1.. (dim 0 is contiguous).  box = box extents (elements).  swizzle_bytes in {0, 32, 64, 128}.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle_bytes);

bool pdl_enabled();  // SR_PDL env (default 1) / sr_set_pdl()

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE function attribute: set it once per (kernel instantiation, device).
// `done` is the caller's function-local static bitmask (one bit per device ordinal < 64).
template <typename K>
static inline cudaError_t ensure_dynamic_smem(K kern, int bytes, unsigned long long* done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
  if (bit != 0 && (*done & bit) != 0) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) *done |= bit;
  return e;
}

// Launch with the programmatic-dependent-launch attribute (kernels must call pdl_wait() before touching global memory).
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#define SR_CHECK_CUDA(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::sr::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return SR_ERR_CUDA;                                                              \
    }                                                                                        \
  } while (0)

#define SR_REQUIRE(cond, ...)                \
  do {                                       \
    if (!(cond)) {                           \
      ::sr::set_last_error(__VA_ARGS__);     \
      return SR_ERR_INVALID;           \
    }                                        \
  } while (0)

}  // namespace sr
