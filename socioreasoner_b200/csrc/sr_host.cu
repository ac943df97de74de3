// sr_host.cu -- host helpers (see sr_host.h) + library-level C-ABI entry points.
#include <cstring>
#include <mutex>
#include <unordered_map>
#include "sr_host.h"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/socioreasoner_b200.h"

namespace sr {

void set_pdl(int on);

static thread_local char g_err[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_last_error() { return g_err; }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int max_smem_optin() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) n = 0;
  return n;
}

static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("SR_PDL");
    g_pdl = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return g_pdl != 0;
}
void set_pdl(int on) { g_pdl = on ? 1 : 0; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
    set_last_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// A tensor map is a pure function of (address, shape, strides, box, swizzle): the ~300 GEMM / attention launches of a step use
// the same few dozen operands step after step, so encoded maps are kept in a small per-process cache instead of calling the
// driver's encoder twice or three times per launch.
namespace {
struct TmapKey {
  uint64_t w[16];
  bool operator==(const TmapKey& o) const { return std::memcmp(w, o.w, sizeof(w)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint64_t x : k.w) h = (h ^ x) * 0x100000001b3ull;
    return static_cast<size_t>(h ^ (h >> 29));
  }
};
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
constexpr size_t kTmapCacheMax = 8192;
}  // namespace

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle_bytes) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return SR_ERR_DRIVER;
  TmapKey key = {};
  key.w[0] = reinterpret_cast<uint64_t>(base);
  key.w[1] = (static_cast<uint64_t>(rank) << 32) | static_cast<uint32_t>(swizzle_bytes);
  for (int i = 0; i < rank && i < 5; ++i) {
    key.w[2 + i] = dims[i];
    key.w[7 + i] = (i > 0) ? strides_bytes[i - 1] : 0;
    key.w[12 + (i >> 1)] |= static_cast<uint64_t>(box[i]) << (32 * (i & 1));
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      *out = it->second;
      return SR_OK;
    }
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_last_error("tensor map base %p not 16-byte aligned", base);
    return SR_ERR_INVALID;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      if (gstr[i - 1] % 16 != 0) {
        set_last_error("tensor map stride[%d]=%llu bytes not a multiple of 16", i, (unsigned long long)gstr[i - 1]);
        return SR_ERR_INVALID;
      }
    }
  }
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  else if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed: CUresult=%d (rank=%d dims=%llu,%llu box=%u,%u)", (int)r, rank,
                   (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0), bx[0], rank > 1 ? bx[1] : 0);
    return SR_ERR_DRIVER;
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmap_cache.size() >= kTmapCacheMax) g_tmap_cache.clear();
    g_tmap_cache.emplace(key, *out);
  }
  return SR_OK;
}

}  // namespace sr

extern "C" {

const char* sr_last_error(void) { return sr::get_last_error(); }

int sr_abi_version(void) { return SR_ABI_VERSION; }

int sr_set_pdl(int enabled) {
  sr::set_pdl(enabled);
  return SR_OK;
}

int sr_device_check(void) {
  int dev = 0;
  SR_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  SR_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  SR_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    sr::set_last_error("device %d is sm_%d%d; this library contains sm_100a code only", dev, major, minor);
    return SR_ERR_UNSUPPORTED;
  }
  return SR_OK;
}

}  // extern "C"
