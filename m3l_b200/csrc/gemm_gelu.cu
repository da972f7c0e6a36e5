// m3l_b200 — feed-forward first Linear + exact-erf GELU, forward, training mode:
//
//   H[M,N] = GELU(A[M,K] W[N,K]^T + b)      G'[M,N] = GELU'(A W^T + b)      (both bf16; G' optional)
//
// i.e. vit_pytorch FeedForward.net[1..2] (nn.Linear -> nn.GELU()) with the derivative the
// backward pass multiplies by stored next to the activation.  Same tcgen05 / TMEM / TMA pipeline as the
// generic kernel in gemm.cu (weight-stationary: the CTA's 256 x K weight slice stays in shared memory,
// A tiles stream through a 4-stage ring, two 256-column TMEM accumulators), but with SIXTEEN epilogue warps.
//
// Why a dedicated kernel (measured, profiles/r01_ncu_full_gemm_v3.txt + tools/microbench.py): with K = 256 a
// tile is 2 k clk of tensor work, while GELU + GELU' is ~14 packed / 20 scalar instructions per element in
// dependent chains of ~100 clk.  With 8 epilogue warps (2 per scheduler, 168 registers each, 96 of them
// pinned by the 64-column round) the schedulers issued 0.5 instructions / clk and the kernel ran 72 us for
// 35 us of HBM traffic — latency-bound, not issue-bound (halving the instruction count with FFMA2 bought
// 4 %).  Here each warp owns a 32-row x 64-column slice processed as two 32-column units (32 accumulator
// registers live instead of 64), so four warps per scheduler hide each other's MUFU / FMA latency.
// Output tiles are 32 rows x 64 B (SWIZZLE_64B tensor maps), one 2 KB staging buffer per warp.
#include <stdlib.h>

#include "common.cuh"
#include "m3l_internal.h"

namespace m3l {

namespace {

constexpr int kBM = 128, kBN = 256, kBK = 64;
constexpr int kEpiWarps = 16;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kStages = 4, kMaxKb = 4;
constexpr int kABytes = kBM * kBK * 2, kBBytes = kBN * kBK * 2;
constexpr int kStgBytes = 32 * 64;                       // per warp: 32 rows x 32 bf16
constexpr int kSmemBytes = 1024 + kStages * kABytes + kMaxKb * kBBytes + kEpiWarps * kStgBytes + kBN * 4 + 256;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

struct alignas(8) Bars {
  uint64_t full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], b_full;
  uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 256, "barrier block");

// one 32 x 32 bf16 unit: registers (thread = row, 16 packed words) -> 64 B-swizzled staging -> TMA store
M3L_DEVINL void store_unit(const CUtensorMap* map, uint32_t stg, int lane, const uint32_t (&w)[16], int col, int row) {
  tma_wait_group_read<0>();               // (issuing lane) the previous store from this buffer has been read
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t addr = stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[4 * j]), "r"(w[4 * j + 1]),
                 "r"(w[4 * j + 2]), "r"(w[4 * j + 3])
                 : "memory");
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (elect_one()) {
    tma_store_2d(map, stg, col, row);
    tma_commit_group();
  }
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_gelu16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_aux,
                   const GemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* resident = smem + kStages * kABytes;           // [kb][256 x 64] weight slice
  uint8_t* staging = resident + kMaxKb * kBBytes;
  float* s_bias = reinterpret_cast<float*>(staging + kEpiWarps * kStgBytes);
  Bars* bars = reinterpret_cast<Bars*>(s_bias + kBN);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = p.N / kBN;
  const int tiles_m = (p.M + kBM - 1) / kBM;
  const int kb_total = (p.K + kBK - 1) / kBK;
  // weight-stationary schedule: this CTA owns n tile (blockIdx.x % tiles_n) and walks m tiles q, q + group, ...
  const int n_idx = blockIdx.x % tiles_n, q = blockIdx.x / tiles_n;
  const int group = ((int)gridDim.x - n_idx + tiles_n - 1) / tiles_n;
  const int count = q < tiles_m ? (tiles_m - q + group - 1) / group : 0;
  const int n0 = n_idx * kBN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_out);
    tma_prefetch_desc(&map_aux);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->tmem_full[b], 1);
      mbar_init(&bars->tmem_empty[b], kEpiWarps);
    }
    mbar_init(&bars->b_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, 2 * kBN);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_wait();      // prologue above overlapped the predecessor kernel; global memory from here on
  if (threadIdx.x < kBN) s_bias[threadIdx.x] = p.bias != nullptr ? p.bias[n0 + threadIdx.x] : 0.f;
  __syncthreads();
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------- TMA producer ---------------------------------------
    if (count > 0 && elect_one()) {
      mbar_arrive_expect_tx(&bars->b_full, kb_total * kBBytes);
      for (int kb = 0; kb < kb_total; ++kb) tma_load_2d(resident + kb * kBBytes, &map_b, &bars->b_full, kb * kBK, n0);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < count; ++i) {
        const int m0 = (q + i * group) * kBM;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&bars->full[stage], kABytes);
          tma_load_2d(smem + stage * kABytes, &map_a, &bars->full[stage], kb * kBK, m0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -----------------------------------------
    if (count > 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, kBN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      mbar_wait(&bars->b_full, 0);
      for (int it = 0; it < count; ++it) {
        const int buf = it & 1;
        mbar_wait(&bars->tmem_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + buf * kBN;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + stage * kABytes);
          const uint32_t b_addr = smem_u32(resident + kb * kBBytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_bf16(tmem_d, umma_smem_desc(a_addr + k * 32, 16, 1024), umma_smem_desc(b_addr + k * 32, 16, 1024),
                      idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&bars->empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&bars->tmem_full[buf]);
      }
    }
  } else {
    // ------------------------------- epilogue: 16 warps ---------------------------------
    const int ew = warp - 2;
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int part = ew >> 2;                  // 64-column slice of the tile
    const uint32_t stg = smem_u32(staging + ew * kStgBytes);
    for (int it = 0; it < count; ++it) {
      const int m0 = (q + it * group) * kBM;
      const int buf = it & 1;
      mbar_wait(&bars->tmem_full[buf], (it >> 1) & 1);
      tc_fence_after_sync();
      const int row0 = m0 + quad * 32;
      const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + buf * kBN + part * 64;
#pragma unroll 1
      for (int u = 0; u < 2; ++u) {
        const int c0 = part * 64 + u * 32;
        uint32_t v[32];
        tmem_ld_32x32(t_acc + u * 32, v);
        tmem_ld_wait();
        if (u == 1) {                          // the accumulator is in registers: hand it back to the tensor core
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->tmem_empty[buf]);
        }
        uint32_t hp[16], gp[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 b2 = *reinterpret_cast<const float2*>(s_bias + c0 + 2 * j);
          const f32x2 x = f2_add(f2_packu(v[2 * j], v[2 * j + 1]), f2_pack(b2.x, b2.y));
          uint32_t x0, x1;
          f2_unpacku(x, x0, x1);
          f32x2 hh, gg;
          gelu_pair(x0, x1, hh, gg);
          float a0, a1;
          f2_unpack(hh, a0, a1);
          hp[j] = pack_bf16x2(a0, a1);
          f2_unpack(gg, a0, a1);
          gp[j] = pack_bf16x2(a0, a1);
        }
        if (p.aux_out != nullptr) store_unit(&map_aux, stg, lane, gp, n0 + c0, row0);
        store_unit(&map_out, stg, lane, hp, n0 + c0, row0);
      }
    }
    tma_wait_group_read<0>();   // (issuing lane) the staging tile has been read; writes complete with the grid
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 2 * kBN);
  }
}

}  // namespace

int gemm_gelu16_run(const GemmPlan& plan, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    M3L_CUDA(cudaFuncSetAttribute(gemm_gelu16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  M3L_CUDA(launch_kernel(gemm_gelu16_kernel, dim3(plan.grid), dim3(kThreads), kSmemBytes, stream, plan.map_a, plan.map_b,
                         plan.map_out32, plan.map_aux32, plan.args));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

}  // namespace m3l
