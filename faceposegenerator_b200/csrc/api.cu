// C-ABI plumbing: version, error reporting, device checks, tensor-map encoding.
#include <mutex>
#include <unordered_map>

#include "../../include/idb.h"
#include <cstdlib>

#include "idb_host.h"

namespace idb {

static thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

struct DeviceInfo {
  int major = -1, minor = -1, sms = 0;
};
static DeviceInfo g_dev[64];
static std::mutex g_dev_mu;

static const DeviceInfo* device_info() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(g_dev_mu);
  DeviceInfo& d = g_dev[dev];
  if (d.major < 0) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return nullptr;
    d.major = prop.major;
    d.minor = prop.minor;
    d.sms = prop.multiProcessorCount;
  }
  return &d;
}

int require_sm100() {
  const DeviceInfo* d = device_info();
  if (d == nullptr) return fail(IDB_E_CUDA, "no CUDA device available (libidb_b200 has no CPU fallback)");
  if (d->major != 10)
    return fail(IDB_E_ARCH, "device is sm_" + std::to_string(d->major) + std::to_string(d->minor) +
                                "; libidb_b200 is built for sm_100a only (no fallback path)");
  return IDB_OK;
}

int num_sms() {
  const DeviceInfo* d = device_info();
  return d ? d->sms : 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static std::atomic<unsigned long long> g_sk_launches{0};
void note_stream_k_launch() { g_sk_launches.fetch_add(1, std::memory_order_relaxed); }

bool pdl_enabled() {
  static const bool on = !(getenv("IDB_PDL") && atoi(getenv("IDB_PDL")) == 0);
  return on;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  return make_tmap(out, base, 2, 128, rank, dims, strides_bytes, box);
}

// Encoded tensor maps are cached per (pointer, element size, swizzle, shape, strides, box) -- SURVEY 8(b): an eager caller
// that re-issues the same layer on the same buffers (the allocator hands a loop the same blocks again) pays a hash lookup
// instead of up to nine cuTensorMapEncodeTiled calls per GEMM.  A map is a pure function of the key, so a stale entry can
// never be wrong, only unused; the table is dropped wholesale when it fills up.
struct TmapKey {
  uint64_t v[16];   // base, (elem | swizzle << 8 | rank << 16), dims[5], strides[4], box[5]
  bool operator==(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 0xcbf29ce484222325ull;
    for (int i = 0; i < 16; ++i) h = (h ^ k.v[i]) * 0x100000001b3ull;
    return static_cast<size_t>(h ^ (h >> 29));
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
static std::mutex g_tmap_mu;
static const bool g_tmap_cache_on = !(getenv("IDB_TMAP_CACHE") && atoi(getenv("IDB_TMAP_CACHE")) == 0);
constexpr size_t TMAP_CACHE_MAX = 1 << 15;

int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int swizzle_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail(IDB_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(IDB_E_BADARG, "TMA operand must be 16-byte aligned");
  if (rank < 1 || rank > 5) return fail(IDB_E_BADARG, "TMA rank must be 1..5");
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.v[0] = reinterpret_cast<uintptr_t>(base);
  key.v[1] = static_cast<uint64_t>(elem_bytes) | (static_cast<uint64_t>(swizzle_bytes) << 8) | (static_cast<uint64_t>(rank) << 16);
  for (int i = 0; i < rank; ++i) key.v[2 + i] = dims[i], key.v[11 + i] = box[i];
  for (int i = 0; i + 1 < rank; ++i) key.v[7 + i] = strides_bytes[i];
  if (g_tmap_cache_on) {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) {
      *out = it->second;
      return IDB_OK;
    }
  }
  cuuint64_t gdims[5];
  cuuint64_t gstr[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (box[i] == 0 || box[i] > 256) return fail(IDB_E_BADARG, "TMA box dimension out of range");
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    if (gstr[i] % 16) return fail(IDB_E_BADARG, "TMA stride must be a multiple of 16 bytes");
  }
  const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdims,
                  gstr, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::string d = "cuTensorMapEncodeTiled failed (CUresult " + std::to_string(static_cast<int>(r)) + ") rank " +
                    std::to_string(rank) + " dims";
    for (int i = 0; i < rank; ++i) d += " " + std::to_string(dims[i]);
    d += " box";
    for (int i = 0; i < rank; ++i) d += " " + std::to_string(box[i]);
    return fail(IDB_E_CUDA, d);
  }
  if (g_tmap_cache_on) {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmaps.size() >= TMAP_CACHE_MAX) g_tmaps.clear();
    g_tmaps.emplace(key, *out);
  }
  return IDB_OK;
}

}  // namespace idb

extern "C" int idb_version(void) { return IDB_VERSION; }

extern "C" int idb_last_error(char* buf, size_t n) {
  if (buf == nullptr || n == 0) return IDB_E_BADARG;
  const std::string& s = idb::g_last_error;
  size_t m = s.size() < n - 1 ? s.size() : n - 1;
  memcpy(buf, s.data(), m);
  buf[m] = 0;
  return IDB_OK;
}

extern "C" size_t idb_sizeof_args(int32_t which) {
  switch (which) {
    case 0: return sizeof(idb_gemm_conv_args);
    case 1: return sizeof(idb_attention_args);
    case 2: return sizeof(idb_groupnorm_args);
    case 3: return sizeof(idb_time_embed_args);
    case 4: return sizeof(idb_attention_bwd_args);
    case 5: return sizeof(idb_groupnorm_bwd_args);
    default: return 0;
  }
}

extern "C" uint64_t idb_launch_count(void) { return idb::g_launches.load(std::memory_order_relaxed); }
extern "C" uint64_t idb_stream_k_launch_count(void) { return idb::g_sk_launches.load(std::memory_order_relaxed); }

extern "C" int idb_device_check(void) { return idb::require_sm100(); }
extern "C" int idb_num_sms(void) { return idb::num_sms(); }
