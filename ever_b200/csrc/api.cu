// Host-side plumbing shared by all kernels: driver entry point for cuTensorMapEncodeTiled (resolved at
// run time so the library does not link libcuda), tensor-map construction, library/version queries.
#include "common.cuh"

#include <mutex>

PFN_encodeTiled evb_get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  });
  return fn;
}

int evb_make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box) {
  PFN_encodeTiled enc = evb_get_encode_tiled();
  if (!enc) return EVB_ERR_DRIVER;
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? EVB_OK : EVB_ERR_DRIVER;
}

#include <cstdlib>
int g_evb_pdl = [] {
  const char* e = getenv("EVB_PDL");
  return (e && e[0] == '0') ? 0 : 1;
}();

// programmatic dependent launch of the tensor-core kernels (prologue overlaps the predecessor's tail): 1 on, 0 off
extern "C" int evb_set_pdl(int on) {
  g_evb_pdl = on ? 1 : 0;
  return EVB_OK;
}

int g_evb_pdl_small = [] {
  const char* e = getenv("EVB_PDL_SMALL");
  return (e && e[0] == '0') ? 0 : 1;
}();
// programmatic dependent launch of the BatchNorm finalize / apply kernels (their launch overlaps the producer's tail)
extern "C" int evb_set_pdl_small(int on) {
  g_evb_pdl_small = on ? 1 : 0;
  return EVB_OK;
}

extern "C" int evb_version() { return 101; }

// Last CUDA error string for diagnostics (does not clear sticky errors).
extern "C" const char* evb_last_cuda_error() { return cudaGetErrorString(cudaPeekAtLastError()); }

// Reads the device-side watchdog words (set when a pipeline wait timed out before the trap).
extern "C" int evb_device_sync_check() {
  cudaError_t e = cudaDeviceSynchronize();
  return e == cudaSuccess ? EVB_OK : EVB_ERR_CUDA;
}
