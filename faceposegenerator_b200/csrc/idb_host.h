// Host-side helpers shared by the C-ABI translation units: thread-local error
// message, device checks, and TMA tensor-map encoding through the driver entry
// point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstring>
#include <string>

namespace idb {

int fail(int code, const std::string& msg);  // records msg for idb_last_error(), returns code
int require_sm100();                         // IDB_OK or IDB_E_ARCH
int num_sms();

// rank-D bf16 tensor map, SWIZZLE_128B, zero OOB fill.  dims/box innermost first;
// strides_bytes has rank-1 entries (stride of dims[1..]).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);
// general form: elem_bytes in {2 (bf16), 4 (fp32)}, swizzle_bytes in {0, 32, 64, 128}
int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int swizzle_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: one flag per (kernel instantiation, device),
// so a second GPU used by the same process gets its own opt-in.  Racing threads at worst set the attribute twice.
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0};
};
template <typename K>
inline cudaError_t ensure_dynamic_smem(K kern, int bytes, PerDeviceOnce& once) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (once.done.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) once.done.fetch_or(bit, std::memory_order_release);
  return e;
}

// every kernel launch of the library goes through one of two sites (launch_pdl below, launch_gemm_e in gemm_tc.cu); both
// count it, so a caller can report exactly how many kernels its calls enqueued (idb_launch_count())
void note_launch();
void note_stream_k_launch();   // idb_gemm_conv calls that took the stream-K schedule (idb_stream_k_launch_count())

// kernel<<<grid, block, smem, stream>>>(args...) with the programmatic-dependent-launch attribute (IDB_PDL=0 disables)
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  note_launch();
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace idb
