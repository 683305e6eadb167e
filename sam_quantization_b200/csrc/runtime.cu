// Host-side runtime of libsamq: thread-local error string, launch counter,
// device check, TMA descriptor construction (driver entry point resolved at run
// time so the library loads on a machine without libcuda, e.g. for the CPU-only
// symbol-export test).
#include "common.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace samq {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return SAMQ_ERR_LAUNCH;
  }
  return SAMQ_OK;
}

// ---------------------------------------------------------------------------
// TMA descriptors
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

struct MapKey {
  uint64_t base;
  uint64_t dims[5];
  uint64_t strides[4];
  uint32_t box[5];
  int rank, elem_bytes, swizzle;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* p = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) {
      h ^= p[i];
      h *= 1099511628211ull;
    }
    return static_cast<size_t>(h);
  }
};

static std::mutex g_map_mu;
static std::unordered_map<MapKey, CUtensorMap*, MapKeyHash> g_maps;
static std::vector<CUtensorMap*> g_retired;

const CUtensorMap* get_tensor_map_nd(const void* base, int rank, const uint64_t* dims,
                                     const uint64_t* strides_bytes, const uint32_t* box,
                                     int elem_bytes, int swizzle) {
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.base = reinterpret_cast<uint64_t>(base);
  key.rank = rank;
  key.elem_bytes = elem_bytes;
  key.swizzle = swizzle;
  for (int i = 0; i < rank; ++i) {
    key.dims[i] = dims[i];
    key.box[i] = box[i];
    if (i > 0) key.strides[i - 1] = strides_bytes[i - 1];
  }
  std::lock_guard<std::mutex> lock(g_map_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) return it->second;

  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available (no CUDA driver?)");
    return nullptr;
  }
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                           : (elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_INT32
                                                              : CU_TENSOR_MAP_DATA_TYPE_UINT8);
  CUtensorMapSwizzle sw = swizzle == 3   ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B
                                         : CU_TENSOR_MAP_SWIZZLE_NONE;
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstrides[i - 1] = strides_bytes[i - 1];
  }
  // 64-byte aligned storage that lives for the process lifetime
  void* mem = nullptr;
  if (posix_memalign(&mem, 64, sizeof(CUtensorMap)) != 0) {
    set_error("posix_memalign failed");
    return nullptr;
  }
  CUtensorMap* map = reinterpret_cast<CUtensorMap*>(mem);
  CUresult r = fn(map, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdims, gstrides,
                  gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu box %u,%u elem %d swz %d)",
              static_cast<int>(r), rank, (unsigned long long)dims[0],
              (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0,
              elem_bytes, swizzle);
    free(mem);
    return nullptr;
  }
  // bound the cache: descriptors are tiny, but activation pointers may churn.  Evicted
  // descriptors are freed one generation late: a caller on another thread may still be between
  // "got the pointer" and "passed *map by value to its launch".
  if (g_maps.size() > 16384) {
    for (CUtensorMap* m : g_retired) free(m);
    g_retired.clear();
    for (auto& kv : g_maps) g_retired.push_back(kv.second);
    g_maps.clear();
  }
  g_maps.emplace(key, map);
  return map;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count are per DEVICE: a process that
// drives several GPUs must not reuse what it set / read for the first one.
int ensure_dynamic_smem(const void* func, int bytes, const char* what) {
  int dev = 0;
  cudaGetDevice(&dev);
  static std::mutex mu;
  static std::unordered_map<uint64_t, int> done;     // (function, device) -> bytes set
  const uint64_t key = reinterpret_cast<uint64_t>(func) * 64 + static_cast<uint64_t>(dev & 63);
  std::lock_guard<std::mutex> lock(mu);
  auto it = done.find(key);
  if (it != done.end() && it->second >= bytes) return SAMQ_OK;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(%s, smem=%d): %s", what, bytes, cudaGetErrorString(e));
    return SAMQ_ERR_LAUNCH;
  }
  done[key] = bytes;
  return SAMQ_OK;
}

static Config g_config;
static std::once_flag g_config_once;

static void load_config() {
  auto is = [](const char* name, const char* val) {
    const char* v = getenv(name);
    return v && strcmp(v, val) == 0;
  };
  Config c = {};
  c.gemm = is("SAMQ_GEMM", "fused") ? 1 : is("SAMQ_GEMM", "dense") ? 2 : is("SAMQ_GEMM", "2cta") ? 3 : 0;
  c.dense_1cta = is("SAMQ_DENSE", "1cta");
  c.attn_exact_max = is("SAMQ_ATTN_MAX", "exact");
  c.attn_win = is("SAMQ_ATTN_WIN", "v1") ? 1 : is("SAMQ_ATTN_WIN", "v2") ? 2 : 0;
  c.attn_glob = is("SAMQ_ATTN_GLOB", "v1") ? 1 : is("SAMQ_ATTN_GLOB", "v2") ? 2 : 0;
  const char* pdl = getenv("SAMQ_PDL");
  c.pdl_mask = pdl ? atoi(pdl) : 0x3;
  g_config = c;
}

const Config& config() {
  std::call_once(g_config_once, load_config);
  return g_config;
}

bool pdl_enabled(int which) { return (config().pdl_mask >> which) & 1; }

int device_sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  static std::mutex mu;
  static int cached[64] = {0};
  std::lock_guard<std::mutex> lock(mu);
  int& n = cached[dev & 63];
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

const CUtensorMap* get_tensor_map_2d(const void* base, uint64_t rows, uint64_t cols,
                                     uint64_t row_stride_bytes, uint32_t box_rows,
                                     uint32_t box_cols, int elem_bytes, int swizzle) {
  uint64_t dims[2] = {cols, rows};
  uint64_t strides[1] = {row_stride_bytes};
  uint32_t box[2] = {box_cols, box_rows};
  return get_tensor_map_nd(base, 2, dims, strides, box, elem_bytes, swizzle);
}

}  // namespace samq

// ---------------------------------------------------------------------------
// C ABI: library / device
// ---------------------------------------------------------------------------
extern "C" {

int samq_abi_version(void) { return 4; }

void samq_config_reload(void) {
  (void)samq::config();
  samq::load_config();
}

int samq_has_ablations(void) {
#ifdef SAMQ_ABLATIONS
  return 1;
#else
  return 0;
#endif
}

const char* samq_last_error(void) { return samq::g_err; }

uint64_t samq_launch_count(void) { return samq::g_launches.load(std::memory_order_relaxed); }

int samq_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    samq::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return SAMQ_ERR_UNSUPPORTED_ARCH;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    samq::set_error("libsamq needs an sm_100a device (B200); found compute capability %d.%d",
                    major, minor);
    return SAMQ_ERR_UNSUPPORTED_ARCH;
  }
  return SAMQ_OK;
}

}  // extern "C"
