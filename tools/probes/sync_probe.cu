// Micro-probes of on-chip synchronisation latencies on sm_100a (measurement tool, not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I m3l_b200/csrc tools/probes/sync_probe.cu -o /tmp/sync_probe
#include "common.cuh"
#include <cstdio>
using namespace m3l;

M3L_DEVINL uint32_t mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}

struct Bars { uint64_t a, b, c[4]; uint32_t tmem; };

// mode 0: try_wait ping-pong between warp 0 and warp 1; mode 1: test_wait spin ping-pong
__global__ void pingpong(int mode, int iters, long long* out) {
  __shared__ Bars bars;
  if (threadIdx.x == 0) { mbar_init(&bars.a, 1); mbar_init(&bars.b, 1); fence_barrier_init(); }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane != 0) return;
  long long t0 = clock64();
  if (warp == 0) {
    for (int i = 0; i < iters; ++i) {
      mbar_arrive(&bars.a);
      if (mode == 0) { while (!mbar_try_wait(&bars.b, i & 1)) {} } else { while (!mbar_test_wait(&bars.b, i & 1)) {} }
    }
    out[0] = clock64() - t0;
  } else if (warp == 1) {
    for (int i = 0; i < iters; ++i) {
      if (mode == 0) { while (!mbar_try_wait(&bars.a, i & 1)) {} } else { while (!mbar_test_wait(&bars.a, i & 1)) {} }
      mbar_arrive(&bars.b);
    }
  }
}

// MMA issue -> commit -> wait, by one thread.  nmma MMAs (128 x N x 16) per iteration.
__global__ void mma_commit(int nmma, int n, int iters, int use_test_wait, long long* out, int a_mn = 0, int b_mn = 0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ Bars bars;
  if (threadIdx.x == 0) { mbar_init(&bars.a, 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x < 32) tmem_alloc(&bars.tmem, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = bars.tmem;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, n, a_mn, b_mn);
    const uint32_t a = smem_u32(smem), b = a + 16384;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      for (int k = 0; k < nmma; ++k) {
        const uint64_t ad = a_mn ? umma_smem_desc(a + (k & 3) * 2048, 8192, 1024) : umma_smem_desc(a + (k & 3) * 32, 16, 1024);
        const uint64_t bd = b_mn ? umma_smem_desc(b + (k & 3) * 2048, 8192, 1024) : umma_smem_desc(b + (k & 3) * 32, 16, 1024);
        umma_bf16(tm, ad, bd, idesc, k > 0);
      }
      umma_commit(&bars.a);
      if (use_test_wait) { while (!mbar_test_wait(&bars.a, i & 1)) {} } else { while (!mbar_try_wait(&bars.a, i & 1)) {} }
      tc_fence_after_sync();
    }
    out[0] = clock64() - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc(tm, 512); }
}

// TMEM drain: `warps` warps each load `cols` columns (x32 chunks) of their lane quadrant per iteration.
__global__ void tmem_drain(int cols, int iters, int wait_each, long long* out) {
  __shared__ Bars bars;
  if (threadIdx.x < 32) tmem_alloc(&bars.tmem, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = bars.tmem;
  const int warp = threadIdx.x >> 5;
  const uint32_t base = tm + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    for (int c = 0; c < cols; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(base + c, v);
      if (wait_each) tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= v[j];
    }
    if (!wait_each) tmem_ld_wait();
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345) out[1] = acc;
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 64); long long h[2];
  const int it = 2000;
  for (int mode = 0; mode < 2; ++mode) {
    pingpong<<<1, 64>>>(mode, it, d); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("pingpong %-9s : %.1f clk per round trip (2 hand-offs)\n", mode ? "test_wait" : "try_wait", (double)h[0] / it);
  }
  cudaFuncSetAttribute(mma_commit, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int tw = 0; tw < 2; ++tw)
    for (int nm : {0, 1, 4, 16}) {
      mma_commit<<<1, 128, 64 * 1024>>>(nm, 256, it, tw, d); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("mma x%-2d (128x256x16) + commit + %-9s : %.1f clk per iteration\n", nm, tw ? "test_wait" : "try_wait", (double)h[0] / it);
    }
  for (int amn = 0; amn < 2; ++amn)
    for (int bmn = 0; bmn < 2; ++bmn)
      for (int n : {64, 128, 256}) {
        mma_commit<<<1, 128, 64 * 1024>>>(16, n, it, 0, d, amn, bmn); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("mma x16 128x%dx16 A %s B %s : %.1f clk per MMA\n", n, amn ? "MN" : "K ", bmn ? "MN" : "K ", ((double)h[0] / it - 319) / 16);
      }
  for (int warps : {1, 4, 8})
    for (int we = 0; we < 2; ++we) {
      tmem_drain<<<1, warps * 32>>>(256, it, we, d); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("tmem drain %d warps x 256 cols (%s) : %.1f clk per iteration  -> %.1f B/clk/SM\n", warps,
             we ? "wait each x32" : "one wait", (double)h[0] / it, warps * 32.0 * 256 * 4 / ((double)h[0] / it));
    }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
