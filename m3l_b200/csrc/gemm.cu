// m3l_b200 — persistent warp-specialised bf16 GEMM on tcgen05 / TMEM / TMA (sm_100a).
//
//   C[M,N] = epilogue( alpha * sum_k A[m,k] * B[n,k] )
//
// This one kernel serves every dense contraction of the VTMAE step that the reference hands to
// cuBLAS through nn.Linear (vit_pytorch Attention.to_qkv / to_out / FeedForward.net, the patch
// embedding Linear and the to_pixels / to_tactiles heads: /root/reference/models/
// pretrain_models.py:113-116,766-779,784) and their autograd dgrad / wgrad products:
//   * forward and dgrad: both operands K-major (activations [M,K]; weights [N,K], or the bf16
//     transposed shadow copy of the weight for dgrad);
//   * wgrad dW[n,k] = sum_m dY[m,n] X[m,k]: both operands MN-major (the contraction runs over
//     the token rows), split-K over the token dimension, fp32 vector red.add into the grad arena.
//
// Structure (one CTA per SM, persistent over a static round-robin tile schedule):
//   warp 0      TMA producer   cp.async.bulk.tensor -> 128B-swizzled smem ring (mbarrier tx)
//   warp 1      MMA issuer     one lane issues tcgen05.mma (128 x BN x 16), fp32 accum in TMEM,
//                              tcgen05.commit frees smem stages / publishes the accumulator
//   warps 2..9  epilogue       tcgen05.ld TMEM -> regs -> swizzled smem transpose -> coalesced
//                              16-byte global accesses; bias / GELU / GELU' / residual fused
// TMEM holds two accumulators (2 x BN columns) so the epilogue of tile i overlaps the MMAs of
// tile i+1 (K is only 256 for most of these GEMMs: SURVEY.md §7.3 hard part 2).
#include "common.cuh"
#include "m3l_internal.h"

namespace m3l {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kNumEpiWarps = 8;
constexpr int kNumThreads = 32 * (2 + kNumEpiWarps);
constexpr int kStagingBytesPerWarp = 32 * 128;

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes =
      1024 /*align slack*/ + kStages * kStageBytes + kNumEpiWarps * kStagingBytesPerWarp + 256;
};

struct alignas(8) GemmBarriers {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

M3L_DEVINL void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// Epilogue flavours (compile time, so the inner loops carry no mode branches):
enum : int {
  EPI_BF16 = 0,      // out bf16 = acc (+ bias) (+ residual)
  EPI_GELU_FWD = 1,  // out bf16 = GELU(acc + bias) ; aux_out bf16 = GELU'(acc + bias)
  EPI_GELU_BWD = 2,  // out bf16 = acc * aux_in   (aux_in = the stored GELU')
  EPI_F32 = 3,       // out fp32 = acc (+ bias)
  EPI_F32_RED = 4,   // out fp32 += acc  (red.global.add, split-K)
};

// ---- per-warp staging tile: 32 rows x 128 B, 16-byte chunks XOR-swizzled by (row & 7) ------------
// "row" mapping: thread `lane` owns row `lane` (this is how tcgen05.ld hands out the accumulator);
// "co" mapping : iteration i, row 4*i + (lane >> 3), chunk lane & 7 (8 lanes cover one 128 B row,
//                so global accesses are full-line coalesced).  Both mappings are bank-conflict free.
M3L_DEVINL void stg_store_row(uint32_t stg, int lane, const uint32_t* w) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t addr = stg + lane * 128 + ((j ^ (lane & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[4 * j]), "r"(w[4 * j + 1]),
                 "r"(w[4 * j + 2]), "r"(w[4 * j + 3])
                 : "memory");
  }
}
M3L_DEVINL void stg_load_row(uint32_t stg, int lane, uint32_t* w) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t addr = stg + lane * 128 + ((j ^ (lane & 7)) << 4);
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(w[4 * j]), "=r"(w[4 * j + 1]), "=r"(w[4 * j + 2]), "=r"(w[4 * j + 3])
                 : "r"(addr)
                 : "memory");
  }
}
M3L_DEVINL uint4 stg_load_co(uint32_t stg, int lane, int i) {
  const int r = 4 * i + (lane >> 3);
  const uint32_t addr = stg + r * 128 + (((lane & 7) ^ (r & 7)) << 4);
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}
M3L_DEVINL void stg_store_co(uint32_t stg, int lane, int i, uint4 v) {
  const int r = 4 * i + (lane >> 3);
  const uint32_t addr = stg + r * 128 + (((lane & 7) ^ (r & 7)) << 4);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const GemmArgs p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* staging = smem + Cfg::kStages * Cfg::kStageBytes;
  GemmBarriers* bars =
      reinterpret_cast<GemmBarriers*>(staging + kNumEpiWarps * kStagingBytesPerWarp);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_m = (p.M + BM - 1) / BM;
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_per_split = (kb_total + p.splits - 1) / p.splits;
  const int num_items = tiles_m * tiles_n * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->tmem_full[b], 1);
      mbar_init(&bars->tmem_empty[b], kNumEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, Cfg::kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------- TMA producer ---------------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int split = item % p.splits;
        const int t = item / p.splits;
        const int n0 = (t % tiles_n) * BN;
        const int m0 = (t / tiles_n) * BM;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_arrive_expect_tx(&bars->full[stage], Cfg::kStageBytes);
          if (!A_MN) {
            tma_load_2d(sa, &map_a, &bars->full[stage], kb * BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(sa + j * 8192, &map_a, &bars->full[stage], m0 + 64 * j, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(sb, &map_b, &bars->full[stage], kb * BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * 8192, &map_b, &bars->full[stage], n0 + 64 * j, kb * BK);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -----------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int split = item % p.splits;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, kb_total);
        const int buf = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&bars->tmem_empty[buf], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + buf * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = A_MN ? umma_smem_desc(a_addr + k * 2048, 8192, 1024)
                                        : umma_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? umma_smem_desc(b_addr + k * 2048, 8192, 1024)
                                        : umma_smem_desc(b_addr + k * 32, 16, 1024);
            umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&bars->empty[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&bars->tmem_full[buf]);
      }
    }
  } else {
    // ------------------------------- epilogue -------------------------------------------
    const int ew = warp - 2;
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int col_half = ew >> 2;
    // columns of the tile this warp drains: BN >= 128 -> its half; BN == 64 -> half 0 takes all
    constexpr int kSpan = BN >= 128 ? BN / 2 : 64;
    const bool active = BN >= 128 || col_half == 0;
    const int span0 = BN >= 128 ? col_half * kSpan : 0;
    constexpr bool kF32 = (EPI == EPI_F32 || EPI == EPI_F32_RED);
    constexpr int kRoundCols = kF32 ? 32 : 64;
    constexpr int kRounds = kSpan / kRoundCols;
    const uint32_t stg = smem_u32(staging + ew * kStagingBytesPerWarp);
    const bool has_bias = (EPI != EPI_GELU_BWD && EPI != EPI_F32_RED) && p.bias != nullptr;
    const bool has_res = (EPI == EPI_BF16) && p.residual != nullptr;
    const int co_row = lane >> 3, co_chunk = lane & 7;
    // fused column sums of the output (bias gradient): kept in registers while this CTA stays on
    // the same N tile (the host sizes the grid as a multiple of tiles_n so that it always does)
    const bool want_colsum = !kF32 && p.colsum_out != nullptr;
    float csum[kF32 ? 1 : kRounds][8];
    int csum_n0 = -1;
    auto flush_colsum = [&]() {
      if (csum_n0 < 0) return;
#pragma unroll
      for (int r = 0; r < (kF32 ? 1 : kRounds); ++r) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float v = csum[r][e];
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          const int gc = csum_n0 + span0 + r * kRoundCols + co_chunk * 8 + e;
          if (co_row == 0 && gc < p.N) atomicAdd(p.colsum_out + gc, v);
        }
      }
    };
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int t = item / p.splits;
      const int n0 = (t % tiles_n) * BN;
      const int m0 = (t / tiles_n) * BM;
      if (want_colsum && active && n0 != csum_n0) {
        flush_colsum();
        csum_n0 = n0;
#pragma unroll
        for (int r = 0; r < (kF32 ? 1 : kRounds); ++r)
#pragma unroll
          for (int e = 0; e < 8; ++e) csum[r][e] = 0.f;
      }
      const int buf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&bars->tmem_full[buf], acc_phase);
      tc_fence_after_sync();
      const int row_base = m0 + quad * 32;
      const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + buf * BN;
      if (!active) {
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tmem_empty[buf]);
        continue;
      }
#pragma unroll 1
      for (int r = 0; r < kRounds; ++r) {
        const int c0 = span0 + r * kRoundCols;        // first tile column of this round
        if constexpr (!kF32) {
          // ---------------- bf16 outputs: 64 columns per round ----------------
          const int gcol = n0 + c0 + co_chunk * 8;      // this lane's 8 columns in the co mapping
          const bool col_ok = gcol < p.N;
          uint4 side[8];
          if (EPI == EPI_GELU_BWD || has_res) {
            const bf16* sp = (EPI == EPI_GELU_BWD) ? p.aux_in : p.residual;
            const int ld = (EPI == EPI_GELU_BWD) ? p.ld_aux : p.ldr;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int grow = row_base + 4 * i + co_row;
              side[i] = (col_ok && grow < p.M)
                            ? *reinterpret_cast<const uint4*>(sp + (size_t)grow * ld + gcol)
                            : make_uint4(0, 0, 0, 0);
            }
          }
          uint32_t v[64];
          tmem_ld_32x32(t_acc + c0, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          tmem_ld_32x32(t_acc + c0 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          tmem_ld_wait();
          if (r == kRounds - 1) {
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[buf]);
          }
          if (has_bias) {
#pragma unroll
            for (int j = 0; j < 64; j += 4) {
              const int c = n0 + c0 + j;
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
              if (c < p.N) b = __ldg(reinterpret_cast<const float4*>(p.bias + c));
              v[j] = __float_as_uint(__uint_as_float(v[j]) + b.x);
              v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) + b.y);
              v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) + b.z);
              v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) + b.w);
            }
          }
          uint32_t w[32];
          if constexpr (EPI == EPI_GELU_FWD) {
            // out = GELU(v); aux_out = GELU'(v) (bf16) so that the backward pass is a plain multiply
            // and never evaluates erf again (the exp(-v^2/2) is shared between the two here)
            if (p.aux_out != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float x0 = __uint_as_float(v[2 * j]), x1 = __uint_as_float(v[2 * j + 1]);
                const GeluParts g0 = gelu_parts(x0), g1 = gelu_parts(x1);
                w[j] = pack_bf16x2(fmaf(x0, g0.pdf, g0.cdf), fmaf(x1, g1.pdf, g1.cdf));
                v[2 * j] = __float_as_uint(x0 * g0.cdf);
                v[2 * j + 1] = __float_as_uint(x1 * g1.cdf);
              }
              stg_store_row(stg, lane, w);
              __syncwarp();
              uint4 u[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) u[i] = stg_load_co(stg, lane, i);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int grow = row_base + 4 * i + co_row;
                if (col_ok && grow < p.M)
                  *reinterpret_cast<uint4*>(p.aux_out + (size_t)grow * p.ld_aux + gcol) = u[i];
              }
              __syncwarp();
            } else {
#pragma unroll
              for (int j = 0; j < 64; ++j) v[j] = __float_as_uint(gelu_erf(__uint_as_float(v[j])));
            }
          }
          if (EPI == EPI_GELU_BWD || has_res) {
#pragma unroll
            for (int i = 0; i < 8; ++i) stg_store_co(stg, lane, i, side[i]);
            __syncwarp();
            stg_load_row(stg, lane, w);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float2 s2 = unpack_bf16x2(w[j]);
              if constexpr (EPI == EPI_GELU_BWD) {   // aux_in holds GELU'(pre-activation)
                v[2 * j] = __float_as_uint(__uint_as_float(v[2 * j]) * s2.x);
                v[2 * j + 1] = __float_as_uint(__uint_as_float(v[2 * j + 1]) * s2.y);
              } else {
                v[2 * j] = __float_as_uint(__uint_as_float(v[2 * j]) + s2.x);
                v[2 * j + 1] = __float_as_uint(__uint_as_float(v[2 * j + 1]) + s2.y);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j)
            w[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          stg_store_row(stg, lane, w);
          __syncwarp();
          uint4 u[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) u[i] = stg_load_co(stg, lane, i);
          bf16* outp = reinterpret_cast<bf16*>(p.out);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int grow = row_base + 4 * i + co_row;
            if (col_ok && grow < p.M) *reinterpret_cast<uint4*>(outp + (size_t)grow * p.ldo + gcol) = u[i];
          }
          if (want_colsum) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (row_base + 4 * i + co_row < p.M) {
                const float2 a = unpack_bf16x2(u[i].x), b = unpack_bf16x2(u[i].y), c = unpack_bf16x2(u[i].z),
                             d = unpack_bf16x2(u[i].w);
#pragma unroll
                for (int rr = 0; rr < kRounds; ++rr)      // static register indexing
                  if (rr == r) {
                    csum[rr][0] += a.x; csum[rr][1] += a.y; csum[rr][2] += b.x; csum[rr][3] += b.y;
                    csum[rr][4] += c.x; csum[rr][5] += c.y; csum[rr][6] += d.x; csum[rr][7] += d.y;
                  }
              }
            }
          }
          __syncwarp();
        } else {
          // ---------------- fp32 outputs: 32 columns per round ----------------
          const int gcol = n0 + c0 + co_chunk * 4;
          const bool col_ok = gcol < p.N;
          uint32_t v[32];
          tmem_ld_32x32(t_acc + c0, v);
          tmem_ld_wait();
          if (r == kRounds - 1) {
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[buf]);
          }
          if (has_bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int c = n0 + c0 + j;
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
              if (c < p.N) b = __ldg(reinterpret_cast<const float4*>(p.bias + c));
              v[j] = __float_as_uint(__uint_as_float(v[j]) + b.x);
              v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) + b.y);
              v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) + b.z);
              v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) + b.w);
            }
          }
          stg_store_row(stg, lane, v);
          __syncwarp();
          uint4 u[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) u[i] = stg_load_co(stg, lane, i);
          float* outp = reinterpret_cast<float*>(p.out);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int grow = row_base + 4 * i + co_row;
            if (col_ok && grow < p.M) {
              float* dst = outp + (size_t)grow * p.ldo + gcol;
              if constexpr (EPI == EPI_F32_RED) {
                red_add_v4(dst, make_float4(__uint_as_float(u[i].x), __uint_as_float(u[i].y),
                                            __uint_as_float(u[i].z), __uint_as_float(u[i].w)));
              } else {
                *reinterpret_cast<uint4*>(dst) = u[i];
              }
            }
          }
          __syncwarp();
        }
      }
    }
    if (want_colsum && active) flush_colsum();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI>
int launch_variant(const GemmPlan& plan, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, EPI>;
  static bool configured = false;
  if (!configured) {
    M3L_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  kern<<<plan.grid, kNumThreads, Cfg::kSmemBytes, stream>>>(plan.map_a, plan.map_b, plan.args);
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

template <int BN>
int launch_bn(const GemmPlan& plan, cudaStream_t stream) {
  const GemmArgs& p = plan.args;
  if (p.a_mn_major) return launch_variant<BN, true, true, EPI_F32_RED>(plan, stream);
  if (p.out_mode == 1) return launch_variant<BN, false, false, EPI_F32>(plan, stream);
  if (p.act == 1) return launch_variant<BN, false, false, EPI_GELU_FWD>(plan, stream);
  if (p.act == 2) return launch_variant<BN, false, false, EPI_GELU_BWD>(plan, stream);
  return launch_variant<BN, false, false, EPI_BF16>(plan, stream);
}

}  // namespace

int gemm_pick_bn(int M, int N) {
  const int sms = device_sm_count();
  const int tiles_m = (M + BM - 1) / BM;
  if (N % 256 == 0 && tiles_m * (N / 256) >= sms) return 256;
  if (N >= 128 && tiles_m * ((N + 127) / 128) >= sms / 2) return 128;
  if (N % 128 == 0 && N >= 512) return 128;
  return N >= 128 && (N % 128 == 0) ? 128 : 64;
}

int gemm_make_plan(GemmPlan* plan, const GemmArgs& args, int bn) {
  GemmArgs p = args;
  M3L_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm: bad shape M=%d N=%d K=%d", p.M, p.N, p.K);
  M3L_REQUIRE(p.N % 8 == 0, "gemm: N=%d must be a multiple of 8", p.N);
  M3L_REQUIRE(p.a_mn_major == p.b_mn_major,
              "gemm: mixed operand majors are not instantiated (a=%d b=%d)", p.a_mn_major,
              p.b_mn_major);
  M3L_REQUIRE(p.out != nullptr && p.a != nullptr && p.b != nullptr, "gemm: null pointer");
  if (p.splits < 1) p.splits = 1;
  const int kb_total = (p.K + BK - 1) / BK;
  if (p.splits > kb_total) p.splits = kb_total;
  // every split must own at least one k-block
  while (p.splits > 1 && ((kb_total + p.splits - 1) / p.splits) * (p.splits - 1) >= kb_total)
    --p.splits;
  M3L_REQUIRE(p.splits == 1 || p.out_mode == 2, "gemm: split-K requires the atomic output mode");
  M3L_REQUIRE(p.alpha == 1.0f, "gemm: alpha != 1 is not supported");
  M3L_REQUIRE((p.out_mode == 2) == (p.a_mn_major != 0),
              "gemm: the accumulate output mode goes with MN-major operands (wgrad) and vice versa");
  M3L_REQUIRE(p.a_mn_major == 0 || (p.bias == nullptr && p.residual == nullptr && p.act == 0),
              "gemm: wgrad mode takes no bias / residual / activation");
  M3L_REQUIRE(p.out_mode != 1 || (p.act == 0 && p.residual == nullptr),
              "gemm: fp32 store mode supports bias only");
  M3L_REQUIRE(p.act == 0 || p.residual == nullptr, "gemm: activation and residual cannot be combined");
  M3L_REQUIRE(p.act != 2 || (p.aux_in != nullptr && p.bias == nullptr), "gemm: GELU' needs aux_in and no bias");
  M3L_REQUIRE(p.ldo % 8 == 0 && (p.residual == nullptr || p.ldr % 8 == 0) &&
                  ((p.aux_in == nullptr && p.aux_out == nullptr) || p.ld_aux % 8 == 0),
              "gemm: output / residual / aux leading dimensions must be multiples of 8");
  if (bn == 0) bn = gemm_pick_bn(p.M, p.N);
  M3L_REQUIRE(bn == 64 || bn == 128 || bn == 256, "gemm: BN=%d unsupported", bn);
  plan->bn = bn;
  plan->args = p;
  if (!p.a_mn_major) {
    int s = make_tmap_2d_bf16(&plan->map_a, p.a, p.M, p.K, p.lda, BM);
    if (s) return s;
    s = make_tmap_2d_bf16(&plan->map_b, p.b, p.N, p.K, p.ldb, bn);
    if (s) return s;
  } else {
    // operands stored [K rows, M or N columns]
    int s = make_tmap_2d_bf16(&plan->map_a, p.a, p.K, p.M, p.lda, BK);
    if (s) return s;
    s = make_tmap_2d_bf16(&plan->map_b, p.b, p.K, p.N, p.ldb, BK);
    if (s) return s;
  }
  const int tiles_n = (p.N + bn - 1) / bn;
  const int tiles = ((p.M + BM - 1) / BM) * tiles_n * p.splits;
  plan->grid = tiles < device_sm_count() ? tiles : device_sm_count();
  if (p.colsum_out != nullptr) {
    M3L_REQUIRE(p.out_mode == 0, "gemm: colsum_out needs the bf16 output mode");
    // keep every CTA on one N tile so the column sums stay in registers across its tiles
    if (plan->grid > tiles_n) plan->grid = (plan->grid / tiles_n) * tiles_n;
  }
  return M3L_OK;
}

int gemm_run(const GemmPlan& plan, cudaStream_t stream) {
  switch (plan.bn) {
    case 64: return launch_bn<64>(plan, stream);
    case 128: return launch_bn<128>(plan, stream);
    case 256: return launch_bn<256>(plan, stream);
  }
  set_last_error("gemm: BN=%d unsupported", plan.bn);
  return M3L_ERR_INVALID;
}

}  // namespace m3l

// ---------------------------------------------------------------------------------------
// C-ABI (include/m3l_b200.h)
// ---------------------------------------------------------------------------------------
extern "C" int m3l_gemm_bf16(const m3l_gemm_args* a, void* stream) {
  if (a == nullptr) return m3l::M3L_ERR_INVALID;
  m3l::GemmArgs g;
  g.a = a->a; g.b = a->b; g.lda = a->lda; g.ldb = a->ldb;
  g.a_mn_major = a->a_mn_major; g.b_mn_major = a->b_mn_major;
  g.M = a->m; g.N = a->n; g.K = a->k; g.splits = a->splits;
  g.out = a->out; g.ldo = a->ldo; g.out_mode = a->out_mode;
  g.bias = a->bias; g.residual = (const m3l::bf16*)a->residual; g.ldr = a->ldr;
  g.act = a->act; g.aux_out = (m3l::bf16*)a->aux_out; g.aux_in = (const m3l::bf16*)a->aux_in;
  g.ld_aux = a->ld_aux; g.alpha = a->alpha; g.colsum_out = a->colsum_out;
  m3l::GemmPlan plan;
  int s = m3l::gemm_make_plan(&plan, g, a->bn);
  if (s) return s;
  return m3l::gemm_run(plan, (cudaStream_t)stream);
}
