// m3l_b200 — host-side helpers: error reporting and TMA tensor-map construction.
#include "common.cuh"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

namespace m3l {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return M3L_OK;
  set_last_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return M3L_ERR_CUDA;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
  });
  return fn;
}

// cuTensorMapEncodeTiled is a DRIVER entry point: it needs a current context on the calling thread, and
// only runtime-API calls bind the primary context implicitly.  PyTorch runs backward() on its own
// threads, where our first call can be this one (CUDA_ERROR_INVALID_CONTEXT, 201, otherwise).
static void bind_primary_context() {
  static thread_local bool bound = false;
  if (!bound) {
    cudaFree(nullptr);
    bound = true;
  }
}

static int make_tmap_2d(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int esize,
                        uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, int row_bytes = 128) {
  bind_primary_context();
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return M3L_ERR_CUDA;
  }
  M3L_REQUIRE(((uintptr_t)base & 15) == 0, "tensor map base %p not 16-byte aligned", base);
  M3L_REQUIRE((ld * esize) % 16 == 0, "tensor map leading dimension %llu not a multiple of 16 bytes",
              (unsigned long long)ld);
  M3L_REQUIRE(box_rows >= 1 && box_rows <= 256, "tensor map box rows %u out of range", box_rows);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * esize};
  cuuint32_t box[2] = {(cuuint32_t)(row_bytes / esize), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(2d rows=%llu cols=%llu ld=%llu box_rows=%u esize=%d) failed: %d",
                   (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld,
                   box_rows, esize, (int)r);
    return M3L_ERR_CUDA;
  }
  return M3L_OK;
}

int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t ld, uint32_t box_rows) {
  return make_tmap_2d(map, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, rows, cols, ld, box_rows);
}

int make_tmap_2d_bf16_sw64(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                           uint64_t ld, uint32_t box_rows) {
  return make_tmap_2d(map, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, rows, cols, ld, box_rows, 64);
}

int make_tmap_2d_f32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                     uint64_t ld, uint32_t box_rows) {
  return make_tmap_2d(map, base, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, rows, cols, ld, box_rows);
}

int make_tmap_3d_bf16(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                      uint64_t s1, uint64_t s2, uint32_t box1) {
  bind_primary_context();
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return M3L_ERR_CUDA;
  }
  M3L_REQUIRE(((uintptr_t)base & 15) == 0, "tensor map base %p not 16-byte aligned", base);
  M3L_REQUIRE((s1 * 2) % 16 == 0 && (s2 * 2) % 16 == 0, "tensor map strides not 16-byte multiples");
  M3L_REQUIRE(box1 >= 1 && box1 <= 256, "tensor map box rows %u out of range", box1);
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstr[2] = {s1 * 2, s2 * 2};
  cuuint32_t box[3] = {64, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(3d %llu x %llu x %llu) failed: %d",
                   (unsigned long long)d2, (unsigned long long)d1, (unsigned long long)d0, (int)r);
    return M3L_ERR_CUDA;
  }
  return M3L_OK;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("M3L_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

int device_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
    n = p.multiProcessorCount;
  }
  return n;
}

}  // namespace m3l

extern "C" const char* m3l_last_error(void) { return m3l::g_last_error; }
