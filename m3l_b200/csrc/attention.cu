// m3l_b200 — fused multi-head attention for short sequences (n <= 256, dim_head = 64) on
// tcgen05 / TMEM / TMA (sm_100a), forward and backward.
//
// Replaces vit_pytorch.vit.Attention's  softmax((q k^T) * scale) v  and its autograd backward
// (the reference materialises the B x H x n x n score tensor through matmul -> Softmax -> matmul:
// /root/reference/models/pretrain_models.py:113,784 via vit-pytorch==1.6.4).
// The whole key sequence of one (sample, head) fits one CTA, so there is no online-softmax loop:
//
//   forward, CTA = (sample, head, 128-query tile):
//     TMA: Q tile, K, V (3-D tensor map over qkv[B, n, 3*inner]; rows >= n are zero-filled)
//     S = Q K^T            tcgen05.mma 128 x NK x 64 -> TMEM            (NK = n rounded up to 16)
//     softmax              4 warps, one TMEM lane (= query row) per thread; P (bf16) -> swizzled smem
//     O = P V              tcgen05.mma 128 x 64 x NK, V consumed as an MN-major operand
//     O / rowsum -> bf16 global; LSE saved for the backward
//
//   backward, CTA = (sample, head), loops over the 128-query tiles:
//     S = Q K^T -> P = exp(S*scale - LSE) -> smem ; dP = dO V^T -> dS = P (dP - delta) scale -> smem
//     dQ_t = dS K ; dV += P^T dO ; dK += dS^T Q   (P / dS slabs double as MN-major A operands)
#include <stdlib.h>

#include "common.cuh"
#include "m3l_internal.h"

// Cycle-counter probes (tools/attn_probe.py) are compiled in only with -DM3L_ATTN_PROFILE: the counters
// cost 18 registers in the backward kernel (spills) and a CS2R per phase.
#ifdef M3L_ATTN_PROFILE
#define M3L_CLK() clock64()
// event timeline of CTA 0 (backward): prof[64 + role * 512 + step * 8 + event] = clock64(), first 64 steps
#define M3L_EVT(role, step, ev)                                                          \
  do {                                                                                   \
    if (p.prof && blockIdx.x == 0 && (step) < 64) p.prof[64 + (role) * 512 + (step) * 8 + (ev)] = clock64(); \
  } while (0)
#else
#define M3L_CLK() 0LL
#define M3L_EVT(role, step, ev) do { } while (0)
#endif

namespace m3l {
namespace {

constexpr int kDh = 64;
// backward probes (profile build): the producer warps whose counters / events are recorded.  Producer warps are 2..9,
// group = (warp - 2) / 4, TMEM lane quadrant = warp % 4: warp 4 = group 0 / rows 0-31 (has work in EVERY step: the
// critical path), warp 8 = group 1 / rows 0-31.  (The first profiles recorded warps 2 and 7 - quadrants 2 and 3, idle in
// every step of the second query tile of n = 192 - and showed mostly waiting.)
#ifndef M3L_ATTN_PROBE_W0
#define M3L_ATTN_PROBE_W0 4
#define M3L_ATTN_PROBE_W1 8
#endif
constexpr int kProbeW0 = M3L_ATTN_PROBE_W0, kProbeW1 = M3L_ATTN_PROBE_W1;
constexpr float kLog2e = 1.4426950408889634f;

M3L_DEVINL uint32_t round_up_pow2_cols(int c) {
  uint32_t r = 32;
  while ((int)r < c) r <<= 1;
  return r;
}

// store 8 bf16 (one 16-byte chunk) of row `row`, logical chunk `chunk` (0..7) into a [rows x 128 B]
// 128B-swizzled K-major slab
M3L_DEVINL void st_swz_chunk(uint32_t slab_u32, int row, int chunk, uint4 v) {
  const uint32_t addr = slab_u32 + row * 128 + ((chunk ^ (row & 7)) << 4);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
M3L_DEVINL uint4 ld_swz_chunk(uint32_t slab_u32, int row, int chunk) {
  const uint32_t addr = slab_u32 + row * 128 + ((chunk ^ (row & 7)) << 4);
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}

// ------------------------------------------------------------------------------------------
// forward: persistent CTA, warp-specialised and software-pipelined over (sample, head) items
//   warp 0       TMA loader: K,V of item i (double-buffered when it fits) and the Q tiles
//   warp 1       MMA issuer: S(j) = Q_j K^T is issued BEFORE O(j-1) = P(j-1) V, so the tensor core
//                works on the next tile's scores while a softmax group is busy with the previous tile
//   warps 2..5   softmax group 0 (TMEM slot 0)     one TMEM lane (= query row) per thread;
//   warps 6..9   softmax group 1 (TMEM slot 1)     tiles alternate between the two groups
// Query rows of an item are split evenly over its tiles (n = 192 -> 2 tiles of 96 valid rows, both
// issued as 128-row MMAs) so the two softmax groups carry equal exp2 work (the MUFU-bound part).
// ------------------------------------------------------------------------------------------
struct AttnFwdBars {
  uint64_t kv_full[2], kv_empty[2], q_full[2], q_empty[2], s_full[2], p_full[2], o_full[2], slot_free[2];
  uint32_t tmem_base;
};

struct AttnFwdParams {
  bf16* out;
  float* lse;
  int n, heads, inner, q_tiles, tile_rows, num_items, kv_bufs, slot_cols;
  float scale;
  long long* prof;   // M3L_ATTN_PROF: cycle counters of CTA 0 (measurement only)
};

__global__ void __launch_bounds__(320, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int n = p.n;
  const int NK = (n + 15) & ~15;                 // keys padded to the MMA granularity
  const int kv_bytes = NK * 128;                 // [NK rows x 64 d] bf16
  const int kv_region = (kv_bytes + 1023) & ~1023;
  const int p_slabs = (NK + 63) / 64;
  uint8_t* sKV = smem;                                           // [kv_bufs][K | V]
  uint8_t* sQ = sKV + p.kv_bufs * 2 * kv_region;                 // [2][128 x 64]
  uint8_t* sP = sQ + 2 * 16384;                                  // [2][p_slabs][128 x 64]
  AttnFwdBars* bars = reinterpret_cast<AttnFwdBars*>(sP + 2 * p_slabs * 16384);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = 2 * p.slot_cols;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_kv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->kv_full[i], 1);
      mbar_init(&bars->kv_empty[i], 1);
      mbar_init(&bars->q_full[i], 1);
      mbar_init(&bars->q_empty[i], 1);
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->p_full[i], 128);
      mbar_init(&bars->o_full[i], 1);
      mbar_init(&bars->slot_free[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, tmem_cols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_wait();      // prologue above overlapped the predecessor kernel; global memory from here on
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------- TMA loader -----------------------------------------
    if (elect_one()) {      // single issuing thread; elect (not lane == 0) keeps TMA / MMA operands in uniform registers
      int it = 0, jt = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        const int h = item % p.heads, b = item / p.heads;
        const int kb = p.kv_bufs == 2 ? (it & 1) : 0;
        const uint32_t kph = (it / p.kv_bufs) & 1;
        mbar_wait(&bars->kv_empty[kb], kph ^ 1);
        uint8_t* k_dst = sKV + kb * 2 * kv_region;
        mbar_arrive_expect_tx(&bars->kv_full[kb], 2 * kv_bytes);
        tma_load_3d(k_dst, &map_kv, &bars->kv_full[kb], p.inner + h * kDh, 0, b);
        tma_load_3d(k_dst + kv_region, &map_kv, &bars->kv_full[kb], 2 * p.inner + h * kDh, 0, b);
        for (int t = 0; t < p.q_tiles; ++t, ++jt) {
          const int slot = jt & 1;
          const uint32_t ph = (jt >> 1) & 1;
          mbar_wait(&bars->q_empty[slot], ph ^ 1);
          mbar_arrive_expect_tx(&bars->q_full[slot], 16384);
          tma_load_3d(sQ + slot * 16384, &map_q, &bars->q_full[slot], h * kDh, t * p.tile_rows, b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -----------------------------------------
    if (elect_one()) {      // single issuing thread; elect (not lane == 0) keeps TMA / MMA operands in uniform registers
      const uint32_t idesc_s = umma_idesc_bf16(128, NK, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, kDh, 0, 1);
      int pv_slot = -1, pv_kb = 0, pv_last = 0;
      uint32_t pv_ph = 0;
      auto issue_pv = [&]() {
        mbar_wait(&bars->p_full[pv_slot], pv_ph);
        tc_fence_after_sync();
        const uint32_t p_addr = smem_u32(sP + pv_slot * p_slabs * 16384);
        const uint32_t v_addr = smem_u32(sKV + pv_kb * 2 * kv_region + kv_region);
        const uint32_t tmem_o = tmem_base + pv_slot * p.slot_cols;
        uint64_t d_v = umma_smem_desc(v_addr, 8192, 1024);       // advanced by a constant add per 16-key step
        for (int kp = 0; kp < NK / 64; ++kp) {
          const uint64_t d_p = umma_smem_desc(p_addr + kp * 16384, 16, 1024);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            umma_bf16(tmem_o, d_p + k4 * (32 >> 4), d_v, idesc_o, (kp > 0 || k4 > 0) ? 1u : 0u);
            d_v += 2048 >> 4;
          }
        }
        for (int kk = (NK / 64) * 4; kk < NK / 16; ++kk) {
          umma_bf16(tmem_o, umma_smem_desc(p_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024), d_v, idesc_o,
                    kk > 0 ? 1u : 0u);
          d_v += 2048 >> 4;
        }
        umma_commit(&bars->o_full[pv_slot]);
        if (pv_last) umma_commit(&bars->kv_empty[pv_kb]);
        pv_slot = -1;
      };
      int it = 0, jt = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        const int kb = p.kv_bufs == 2 ? (it & 1) : 0;
        const uint32_t kph = (it / p.kv_bufs) & 1;
        if (p.kv_bufs == 1 && pv_slot >= 0) issue_pv();   // single K/V buffer: drain before it is reloaded
        mbar_wait(&bars->kv_full[kb], kph);
        const uint32_t k_addr = smem_u32(sKV + kb * 2 * kv_region);
        for (int t = 0; t < p.q_tiles; ++t, ++jt) {
          const int slot = jt & 1;
          const uint32_t ph = (jt >> 1) & 1;
          mbar_wait(&bars->q_full[slot], ph);
          mbar_wait(&bars->slot_free[slot], ph ^ 1);
          tc_fence_after_sync();
          const uint32_t q_addr = smem_u32(sQ + slot * 16384);
          const uint32_t tmem_s = tmem_base + slot * p.slot_cols;
#pragma unroll
          for (int k = 0; k < kDh / 16; ++k)
            umma_bf16(tmem_s, umma_smem_desc(q_addr + k * 32, 16, 1024),
                      umma_smem_desc(k_addr + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&bars->s_full[slot]);
          umma_commit(&bars->q_empty[slot]);
          if (pv_slot >= 0) issue_pv();
          pv_slot = slot; pv_ph = ph; pv_kb = kb; pv_last = (t == p.q_tiles - 1);
        }
      }
      if (pv_slot >= 0) issue_pv();
    }
  } else {
    // ------------------------------- softmax groups -------------------------------------
    const int wg = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;              // TMEM lane == query row inside the tile
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + wg * p.slot_cols;
    const uint32_t p_u32 = smem_u32(sP + wg * p_slabs * 16384);
    const float sl2 = p.scale * kLog2e;
    const int nchunks = (NK + 31) / 32;
    int it = 0, jt = 0;
    long long pf_ws = 0, pf_p1 = 0, pf_p2 = 0, pf_wo = 0, pf_ep = 0, pf_n = 0;
    const long long pf_t0 = M3L_CLK();
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
      const int h = item % p.heads, b = item / p.heads;
      for (int t = 0; t < p.q_tiles; ++t, ++jt) {
        if ((jt & 1) != wg) continue;
        const uint32_t ph = (jt >> 1) & 1;
        const int grow = t * p.tile_rows + row;
        const bool warp_active = (quad * 32 < p.tile_rows) && (t * p.tile_rows + quad * 32 < n);
        const bool valid = row < p.tile_rows && grow < n;
        long long c0 = M3L_CLK();
        mbar_wait(&bars->s_full[wg], ph);
        tc_fence_after_sync();
        long long c1 = M3L_CLK(); pf_ws += c1 - c0;
        float mx = -INFINITY, sum = 0.f;
        if (warp_active) {
          // both passes read S from TMEM in 32-column chunks, software-pipelined over two register
          // buffers: the tcgen05.ld of chunk c+1 is in flight while chunk c is processed
          uint32_t va[32], vb[32];
          // four independent running maxima: a single fmaxf chain over 192 columns is ~1 k clk of pure latency
          float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          auto max_chunk = [&](const uint32_t (&v)[32], int c) {
            if (c * 32 + 32 <= n) {
#pragma unroll
              for (int j = 0; j < 32; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(v[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c * 32 + j < n) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(v[j]));
            }
          };
          tmem_ld_32x32(t_row, va);
          for (int c = 0; c < nchunks; c += 2) {
            tmem_ld_wait();
            if (c + 1 < nchunks) tmem_ld_32x32(t_row + (c + 1) * 32, vb);
            max_chunk(va, c);
            if (c + 1 < nchunks) {
              tmem_ld_wait();
              if (c + 2 < nchunks) tmem_ld_32x32(t_row + (c + 2) * 32, va);
              max_chunk(vb, c + 1);
            }
          }
          mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
          const float mxs = mx * sl2;
          { long long c2 = M3L_CLK(); pf_p1 += c2 - c1; c1 = c2; }
          float sum1 = 0.f;                          // two partial sums: shorter FADD dependency chains
          const f32x2 sl2_2 = f2_splat(sl2), nmx_2 = f2_splat(-mxs);
          auto exp_chunk = [&](const uint32_t (&v)[32], int c) {
            const int col0 = c * 32;
            const bool full = col0 + 32 <= n;          // no padded key column in this chunk: no per-element select
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (col0 + g * 8 < NK) {
                float pv[8];
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                  float a0, a1;                        // x * scale * log2(e) - max, two columns per FFMA2
                  f2_unpack(f2_fma(f2_packu(v[g * 8 + j], v[g * 8 + j + 1]), sl2_2, nmx_2), a0, a1);
                  pv[j] = exp2f(a0);
                  pv[j + 1] = exp2f(a1);
                }
                if (!full) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) pv[j] = (col0 + g * 8 + j < n) ? pv[j] : 0.f;
                }
                sum += (pv[0] + pv[1]) + (pv[2] + pv[3]);
                sum1 += (pv[4] + pv[5]) + (pv[6] + pv[7]);
                uint4 u;
                u.x = pack_bf16x2(pv[0], pv[1]); u.y = pack_bf16x2(pv[2], pv[3]);
                u.z = pack_bf16x2(pv[4], pv[5]); u.w = pack_bf16x2(pv[6], pv[7]);
                const int col = col0 + g * 8;
                st_swz_chunk(p_u32 + (col >> 6) * 16384, row, (col & 63) >> 3, u);
              }
            }
          };
          tmem_ld_32x32(t_row, va);
          for (int c = 0; c < nchunks; c += 2) {
            tmem_ld_wait();
            if (c + 1 < nchunks) tmem_ld_32x32(t_row + (c + 1) * 32, vb);
            exp_chunk(va, c);
            if (c + 1 < nchunks) {
              tmem_ld_wait();
              if (c + 2 < nchunks) tmem_ld_32x32(t_row + (c + 2) * 32, va);
              exp_chunk(vb, c + 1);
            }
          }
          sum += sum1;
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        mbar_arrive(&bars->p_full[wg]);
        { long long c2 = M3L_CLK(); pf_p2 += c2 - c1; c1 = c2; }
        // ---- epilogue
        mbar_wait(&bars->o_full[wg], ph);
        tc_fence_after_sync();
        { long long c2 = M3L_CLK(); pf_wo += c2 - c1; c1 = c2; }
        if (warp_active) {
          const float inv = 1.0f / sum;
          uint32_t o0[32], o1[32];
          tmem_ld_32x32(t_row, o0);
          tmem_ld_32x32(t_row + 32, o1);
          tmem_ld_wait();
          if (valid) {
            bf16* dst = p.out + ((size_t)b * n + grow) * p.inner + h * kDh;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(o0[g * 8 + 0]) * inv, __uint_as_float(o0[g * 8 + 1]) * inv);
              u.y = pack_bf16x2(__uint_as_float(o0[g * 8 + 2]) * inv, __uint_as_float(o0[g * 8 + 3]) * inv);
              u.z = pack_bf16x2(__uint_as_float(o0[g * 8 + 4]) * inv, __uint_as_float(o0[g * 8 + 5]) * inv);
              u.w = pack_bf16x2(__uint_as_float(o0[g * 8 + 6]) * inv, __uint_as_float(o0[g * 8 + 7]) * inv);
              *reinterpret_cast<uint4*>(dst + g * 8) = u;
              u.x = pack_bf16x2(__uint_as_float(o1[g * 8 + 0]) * inv, __uint_as_float(o1[g * 8 + 1]) * inv);
              u.y = pack_bf16x2(__uint_as_float(o1[g * 8 + 2]) * inv, __uint_as_float(o1[g * 8 + 3]) * inv);
              u.z = pack_bf16x2(__uint_as_float(o1[g * 8 + 4]) * inv, __uint_as_float(o1[g * 8 + 5]) * inv);
              u.w = pack_bf16x2(__uint_as_float(o1[g * 8 + 6]) * inv, __uint_as_float(o1[g * 8 + 7]) * inv);
              *reinterpret_cast<uint4*>(dst + 32 + g * 8) = u;
            }
            if (p.lse) p.lse[((size_t)b * p.heads + h) * n + grow] = mx * p.scale + __logf(sum);
          }
        }
        tc_fence_before_sync();
        mbar_arrive(&bars->slot_free[wg]);
        { long long c2 = M3L_CLK(); pf_ep += c2 - c1; pf_n += 1; }
      }
    }
    if (p.prof && blockIdx.x == 0 && warp == 2 && lane == 0) {
      p.prof[0] = pf_ws; p.prof[1] = pf_p1; p.prof[2] = pf_p2; p.prof[3] = pf_wo; p.prof[4] = pf_ep;
      p.prof[5] = pf_n; p.prof[6] = M3L_CLK() - pf_t0;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// backward: persistent CTA over (sample, head) items; key tiles (128) outer, query tiles (128) inner.
//   P_ij = exp(S_ij * scale - LSE_i) and dS_ij = P_ij (dP_ij - delta_i) scale are element-wise given
//   LSE / delta, so no row-wide pass is needed and S_ij, dP_ij sit in TMEM side by side:
//     TMEM columns: [0,128) S_ij  [128,256) dP_ij  [256,320) dQ_0  [320,384) dQ_1  [384,448) dV_j  [448,512) dK_j
//   warp 0      TMA loader, running ahead of the tensor core: Q / dO double-buffered across items, the
//               last key tile's K / V double-buffered, the earlier key tiles single-buffered but released
//               as soon as their last product is issued (half way through the item)
//   warp 1      MMA issuer, ONE flat software pipeline over all steps of all items of this CTA:
//                 [S, dP](step g+1)  is issued as soon as the producers hold step g's values in registers,
//                 [dV += P^T dO, dK += dS^T Q, dQ += dS K](step g)  once the P / dS slabs are written,
//               so neither item boundaries nor the producers' math leave the tensor core idle
//   warps 2..9  two groups of 128 threads; thread = (query row, 64-key half): both read their half of
//               S and dP from TMEM and write the P and dS slabs (bf16, swizzled) the MMAs consume as
//               K-major (dQ) and MN-major (dV, dK) operands.  Group 0 drains dV_j / dQ_0, group 1 dK_j / dQ_1.
//   Measured before this structure (cycle counters, tools/attn_probe.py): 9.3 k clk per step of which the
//   tensor core was busy 2.7 k; the rest were exposed Q/dO load latency at item starts (4.3 k clk per
//   item), the delta prologue (4.8 k per item) and hand-offs that serialised math and MMAs.
// ------------------------------------------------------------------------------------------
struct AttnBwdBars {
  uint64_t qdo_full[2], qdo_empty[2], kva_full, kva_empty, kvl_full[2], kvl_empty[2];
  uint64_t sdp_full, sdp_free, pds_full, pds_free, dkv_full, dkv_free, item_done;
  uint32_t tmem_base;
};

struct AttnBwdParams {
  const bf16* o;
  const bf16* dout;
  const float* lse;
  const float* delta;   // [batch*n, heads] precomputed rowsum(dO * O), or nullptr
  bf16* dqkv;
  int n, heads, inner, q_tiles, key_tiles, num_items;
  int qdo_bufs, kvl_bufs;       // 1 or 2
  int q_rows[2], q_off[2];      // per query tile: rows loaded (multiple of 64), byte offset in a Q / dO region
  int k_rows[2];                // per key tile: rows loaded (multiple of 64)
  int q_region;                 // bytes of one Q (= one dO) region
  int kva_region, kvl_region;   // bytes of K (= V) of the early key tiles / of the last key tile
  float scale;
  long long* prof;   // M3L_ATTN_PROF: cycle counters of CTA 0 (measurement only)
};

M3L_DEVINL void store_row64(bf16* dst, const uint32_t (&a0)[32], const uint32_t (&a1)[32]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 u;
    u.x = pack_bf16x2(__uint_as_float(a0[g * 8 + 0]), __uint_as_float(a0[g * 8 + 1]));
    u.y = pack_bf16x2(__uint_as_float(a0[g * 8 + 2]), __uint_as_float(a0[g * 8 + 3]));
    u.z = pack_bf16x2(__uint_as_float(a0[g * 8 + 4]), __uint_as_float(a0[g * 8 + 5]));
    u.w = pack_bf16x2(__uint_as_float(a0[g * 8 + 6]), __uint_as_float(a0[g * 8 + 7]));
    *reinterpret_cast<uint4*>(dst + g * 8) = u;
    u.x = pack_bf16x2(__uint_as_float(a1[g * 8 + 0]), __uint_as_float(a1[g * 8 + 1]));
    u.y = pack_bf16x2(__uint_as_float(a1[g * 8 + 2]), __uint_as_float(a1[g * 8 + 3]));
    u.z = pack_bf16x2(__uint_as_float(a1[g * 8 + 4]), __uint_as_float(a1[g * 8 + 5]));
    u.w = pack_bf16x2(__uint_as_float(a1[g * 8 + 6]), __uint_as_float(a1[g * 8 + 7]));
    *reinterpret_cast<uint4*>(dst + 32 + g * 8) = u;
  }
}

__global__ void __launch_bounds__(320, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                const __grid_constant__ CUtensorMap map_dqkv, const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int n = p.n;
  const int NK = (n + 15) & ~15;
  uint8_t* sQ = smem;                                     // [qdo_bufs][q_region]
  uint8_t* sDO = sQ + p.qdo_bufs * p.q_region;            // [qdo_bufs][q_region]
  uint8_t* sP = sDO + p.qdo_bufs * p.q_region;            // [2 slabs][128 q][64 keys]
  uint8_t* sDS = sP + 2 * 16384;                          // [2 slabs][128 q][64 keys]
  uint8_t* sKA = sDS + 2 * 16384;                         // early key tiles: [K | V], kva_region each
  uint8_t* sKL = sKA + 2 * p.kva_region;                  // last key tile: [kvl_bufs][K | V]
  AttnBwdBars* bars = reinterpret_cast<AttnBwdBars*>(sKL + p.kvl_bufs * 2 * p.kvl_region);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kColS = 0, kColDP = 128, kColDQ = 256, kColDV = 384, kColDK = 448;
  const int steps_per_item = p.key_tiles * p.q_tiles;
  const int my_items = (int)blockIdx.x < p.num_items ? (p.num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int jl = p.key_tiles - 1;                         // the last key tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_do);
    tma_prefetch_desc(&map_dqkv);
    // the operand regions are released by the MMA thread (its last product on them) AND by the two producer groups,
    // which stage dQ / dK / dV in the dead regions and TMA-store them from there: 1 + 2 arrivals
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->qdo_full[i], 1);
      mbar_init(&bars->qdo_empty[i], 3);
      mbar_init(&bars->kvl_full[i], 1);
      mbar_init(&bars->kvl_empty[i], 3);
    }
    mbar_init(&bars->kva_full, 1);
    mbar_init(&bars->kva_empty, 3);
    mbar_init(&bars->sdp_full, 1);
    mbar_init(&bars->sdp_free, 256);
    mbar_init(&bars->pds_full, 256);
    mbar_init(&bars->pds_free, 1);
    mbar_init(&bars->dkv_full, 1);
    mbar_init(&bars->dkv_free, 256);
    mbar_init(&bars->item_done, 256);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_wait();      // prologue above overlapped the predecessor kernel; global memory from here on
  pdl_trigger();

  // buffer selection helpers (identical in all roles)
  auto qdo_buf = [&](int it) { return p.qdo_bufs == 2 ? (it & 1) : 0; };
  auto qdo_par = [&](int it) { return (uint32_t)((it / p.qdo_bufs) & 1); };
  auto kvl_buf = [&](int it) { return p.kvl_bufs == 2 ? (it & 1) : 0; };
  auto kvl_par = [&](int it) { return (uint32_t)((it / p.kvl_bufs) & 1); };

  if (warp == 0) {
    // ------------------------------- TMA loader -----------------------------------------
    // Row statistics of ALL items of this CTA -> L2, by the 31 otherwise idle lanes of this warp.  The producers read
    // LSE / delta into runtime-indexed (local-memory) arrays, which blocks until the loads land: 1.3 k clk per step in
    // the r02 timeline, because LSE was written a whole forward pass earlier and came from HBM.  (Prefetching from the
    // producer warps themselves cost registers they do not have: 168 / 168 used, the address math spilled.)
    for (int it = 0; it < my_items; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int h = item % p.heads, b = item / p.heads;
      const char* l0 = reinterpret_cast<const char*>(p.lse + ((size_t)b * p.heads + h) * n);
      for (int off = lane * 128; off < n * 4; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(l0 + off));
      if (p.delta != nullptr) {                 // delta rows of a sample hold all heads (the block is shared by its 4 items)
        const char* d0 = reinterpret_cast<const char*>(p.delta + (size_t)b * n * p.heads);
        for (int off = lane * 128; off < n * p.heads * 4; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(d0 + off));
      }
    }
    if (elect_one()) {      // single issuing thread; elect (not lane == 0) keeps TMA / MMA operands in uniform registers
      for (int it = 0; it < my_items; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int h = item % p.heads, b = item / p.heads;
        // Q / dO of the item
        {
          const int s = qdo_buf(it);
          mbar_wait(&bars->qdo_empty[s], qdo_par(it) ^ 1);
          int bytes = 0;
          for (int i = 0; i < p.q_tiles; ++i) bytes += 2 * p.q_rows[i] * 128;
          mbar_arrive_expect_tx(&bars->qdo_full[s], bytes);
          for (int i = 0; i < p.q_tiles; ++i)
            for (int r = 0; r < p.q_rows[i]; r += 64) {
              tma_load_3d(sQ + s * p.q_region + p.q_off[i] + r * 128, &map_qkv, &bars->qdo_full[s], h * kDh, i * 128 + r, b);
              tma_load_3d(sDO + s * p.q_region + p.q_off[i] + r * 128, &map_do, &bars->qdo_full[s], h * kDh, i * 128 + r, b);
            }
        }
        // K / V of the last key tile
        {
          const int kb = kvl_buf(it);
          mbar_wait(&bars->kvl_empty[kb], kvl_par(it) ^ 1);
          mbar_arrive_expect_tx(&bars->kvl_full[kb], 2 * p.k_rows[jl] * 128);
          uint8_t* kd = sKL + kb * 2 * p.kvl_region;
          for (int r = 0; r < p.k_rows[jl]; r += 64) {
            tma_load_3d(kd + r * 128, &map_qkv, &bars->kvl_full[kb], p.inner + h * kDh, jl * 128 + r, b);
            tma_load_3d(kd + p.kvl_region + r * 128, &map_qkv, &bars->kvl_full[kb], 2 * p.inner + h * kDh, jl * 128 + r, b);
          }
        }
        // K / V of the early key tiles (single buffer, released half way through the previous item)
        if (p.key_tiles > 1) {
          mbar_wait(&bars->kva_empty, (uint32_t)(it & 1) ^ 1);
          mbar_arrive_expect_tx(&bars->kva_full, 2 * jl * 128 * 128);
          for (int r = 0; r < jl * 128; r += 64) {
            tma_load_3d(sKA + r * 128, &map_qkv, &bars->kva_full, p.inner + h * kDh, r, b);
            tma_load_3d(sKA + p.kva_region + r * 128, &map_qkv, &bars->kva_full, 2 * p.inner + h * kDh, r, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -----------------------------------------
    if (my_items > 0 && elect_one()) {
      const uint32_t idesc_dq = umma_idesc_bf16(128, kDh, 0, 1);     // A K-major (dS), B MN-major (K)
      const uint32_t idesc_dkv = umma_idesc_bf16(128, kDh, 1, 1);    // both MN-major
      const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sDS);
      auto q_addr = [&](int it, int i) { return smem_u32(sQ + qdo_buf(it) * p.q_region + p.q_off[i]); };
      auto do_addr = [&](int it, int i) { return smem_u32(sDO + qdo_buf(it) * p.q_region + p.q_off[i]); };
      auto k_addr = [&](int it, int j) {
        return j < jl ? smem_u32(sKA + j * 16384) : smem_u32(sKL + kvl_buf(it) * 2 * p.kvl_region);
      };
      auto v_addr = [&](int it, int j) {
        return j < jl ? smem_u32(sKA + p.kva_region + j * 16384)
                      : smem_u32(sKL + kvl_buf(it) * 2 * p.kvl_region + p.kvl_region);
      };
      auto wait_loads = [&](int it) {
        mbar_wait(&bars->qdo_full[qdo_buf(it)], qdo_par(it));
        mbar_wait(&bars->kvl_full[kvl_buf(it)], kvl_par(it));
        if (p.key_tiles > 1) mbar_wait(&bars->kva_full, (uint32_t)(it & 1));
      };
      // S / dP of step (it, j, i):  S_ij = Q_i K_j^T, dP_ij = dO_i V_j^T
      auto issue_sdp = [&](int it, int j, int i) {
        const int nkj = min(128, NK - j * 128);                      // valid (padded) keys of this tile
        const uint32_t idesc_s = umma_idesc_bf16(128, nkj, 0, 0);
        const uint32_t kj = k_addr(it, j), vj = v_addr(it, j), qi = q_addr(it, i), doi = do_addr(it, i);
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k)
          umma_bf16(tmem_base + kColS, umma_smem_desc(qi + k * 32, 16, 1024),
                    umma_smem_desc(kj + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k)
          umma_bf16(tmem_base + kColDP, umma_smem_desc(doi + k * 32, 16, 1024),
                    umma_smem_desc(vj + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&bars->sdp_full);
      };
      // S / dP of the next item's first step can only be issued ahead when that item's operands have
      // their own buffers
      const bool cross = p.qdo_bufs == 2 && p.kvl_bufs == 2;
      int g = 0, dk = 0;                                             // global step / key-tile counters
      long long m_sdp = 0, m_acc = 0, m_w1 = 0, m_w2 = 0;
      const bool serial = p.prof != nullptr && p.prof[32] != 0;      // measurement: wait for every batch
      for (int it = 0; it < my_items; ++it) {
        if (it == 0 || !cross) {
          wait_loads(it);
          if (g > 0) mbar_wait(&bars->sdp_free, (g - 1) & 1);
          tc_fence_after_sync();
          issue_sdp(it, 0, 0);
        }
        for (int j = 0; j < p.key_tiles; ++j) {
          const int nkj = min(128, NK - j * 128);
          const uint32_t kj = k_addr(it, j);
          for (int i = 0; i < p.q_tiles; ++i, ++g) {
            // ---- S / dP of the next step (possibly the next item's first) go to the tensor core first
            int it2 = it, j2 = j, i2 = i + 1;
            if (i2 == p.q_tiles) { i2 = 0; ++j2; }
            if (j2 == p.key_tiles) { j2 = 0; ++it2; }
            if (it2 < my_items && (it2 == it || cross)) {
              if (it2 != it) wait_loads(it2);
              long long c0 = M3L_CLK();
              mbar_wait(&bars->sdp_free, g & 1);       // the producers hold step g's S / dP in registers
              tc_fence_after_sync();
              long long c1 = M3L_CLK(); m_w1 += c1 - c0;
              M3L_EVT(0, g, 0);
              issue_sdp(it2, j2, i2);
              M3L_EVT(0, g, 1);
              if (serial) { mbar_wait(&bars->sdp_full, (g + 1) & 1); m_sdp += M3L_CLK() - c1; }
            }
            // ---- wait for P_ij / dS_ij, then the three accumulating products
            long long c2 = M3L_CLK();
            mbar_wait(&bars->pds_full, g & 1);
            long long c3 = M3L_CLK(); m_w2 += c3 - c2;
            M3L_EVT(0, g, 2);
            if (i == 0 && dk > 0) mbar_wait(&bars->dkv_free, (dk - 1) & 1);               // dV/dK accumulators drained
            if (j == 0 && i == 0 && it > 0) mbar_wait(&bars->item_done, (it - 1) & 1);   // dQ accumulators drained
            tc_fence_after_sync();
            const uint32_t qi = q_addr(it, i), doi = do_addr(it, i);
            const int qk = p.q_rows[i] / 16;                         // contraction over the loaded query rows
            {
              // descriptors advance by a constant per k-step: one 64-bit add each instead of rebuilding the bit
              // fields (the start-address field holds addr >> 4 and cannot carry out: smem addresses < 256 KB)
              uint64_t d_p = umma_smem_desc(p_addr, 16384, 1024), d_do = umma_smem_desc(doi, 8192, 1024);
              uint64_t d_ds = umma_smem_desc(ds_addr, 16384, 1024), d_q = umma_smem_desc(qi, 8192, 1024);
              for (int kk = 0; kk < qk; ++kk) {
                const uint32_t acc = (i > 0 || kk > 0) ? 1u : 0u;
                umma_bf16(tmem_base + kColDV, d_p, d_do, idesc_dkv, acc);
                umma_bf16(tmem_base + kColDK, d_ds, d_q, idesc_dkv, acc);
                d_p += 2048 >> 4; d_do += 2048 >> 4; d_ds += 2048 >> 4; d_q += 2048 >> 4;
              }
              uint64_t d_k = umma_smem_desc(kj, 8192, 1024);
              const uint32_t tq = tmem_base + kColDQ + i * 64;
              for (int kp = 0; kp < nkj / 64; ++kp) {                // contraction over this tile's keys, 64-key panels
                const uint64_t d_a = umma_smem_desc(ds_addr + kp * 16384, 16, 1024);
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                  umma_bf16(tq, d_a + k4 * (32 >> 4), d_k, idesc_dq, (j > 0 || kp > 0 || k4 > 0) ? 1u : 0u);
                  d_k += 2048 >> 4;
                }
              }
              for (int kk = (nkj / 64) * 4; kk < nkj / 16; ++kk) {   // remaining 16-key steps of a partial panel
                umma_bf16(tq, umma_smem_desc(ds_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024), d_k, idesc_dq,
                          (j > 0 || kk > 0) ? 1u : 0u);
                d_k += 2048 >> 4;
              }
            }
            umma_commit(&bars->pds_free);                            // P / dS slabs may be rewritten
            M3L_EVT(0, g, 3);
            if (i == p.q_tiles - 1) {
              umma_commit(&bars->dkv_full);
              ++dk;
              if (j == jl - 1) umma_commit(&bars->kva_empty);        // last product on the early key tiles
            }
            if (serial) { long long c4 = M3L_CLK(); mbar_wait(&bars->pds_free, g & 1); m_acc += M3L_CLK() - c4; }
          }
        }
        umma_commit(&bars->qdo_empty[qdo_buf(it)]);
        umma_commit(&bars->kvl_empty[kvl_buf(it)]);
      }
      if (p.prof && blockIdx.x == 0) { p.prof[26] = m_w1; p.prof[27] = m_sdp; p.prof[28] = m_w2; p.prof[29] = m_acc; }
    }
  } else {
    // ------------------------------- P / dS producers + epilogues ------------------------
    const int wg = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t p_slab = smem_u32(sP + wg * 16384), ds_slab = smem_u32(sDS + wg * 16384);
    const float sl2 = p.scale * kLog2e;
    int g = 0, dk = 0;
    const long long pf_t0 = M3L_CLK();
    long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long pc = pf_t0;
#define PF(k) { const long long c_ = M3L_CLK(); pf[k] += c_ - pc; pc = c_; \
                if (lane == 0 && (warp == kProbeW0 || warp == kProbeW1)) M3L_EVT(warp == kProbeW0 ? 1 : 2, g, k); }
    // delta_i = rowsum(dO * O) and LSE (log2 domain) of this thread's row in each query tile; with a
    // precomputed delta the next item's values are fetched one item ahead
    float delta[2] = {0.f, 0.f}, l2[2] = {0.f, 0.f}, delta_n[2] = {0.f, 0.f}, l2_n[2] = {0.f, 0.f};
    auto fetch_stats = [&](int it, float (&d)[2], float (&l)[2]) {
      const int item = blockIdx.x + it * gridDim.x;
      const int h = item % p.heads, b = item / p.heads;
      if (p.delta != nullptr) {
        // all (up to four) loads are issued before any result is stored to the local-memory arrays: in the loop form
        // below every store waited for its own load, four L2 round trips back to back (3 k clk per item in the r02
        // timeline, on the producers' critical path)
        const int g0 = row, g1 = 128 + row;
        const bool v0 = g0 < n, v1 = p.q_tiles > 1 && g1 < n;
        const float* pl = p.lse + ((size_t)b * p.heads + h) * n;
        const float* pdl = p.delta + (size_t)b * n * p.heads + h;
        const float la = v0 ? __ldg(pl + g0) : 0.f, lb = v1 ? __ldg(pl + g1) : 0.f;
        const float da = v0 ? __ldg(pdl + (size_t)g0 * p.heads) : 0.f, db = v1 ? __ldg(pdl + (size_t)g1 * p.heads) : 0.f;
        l[0] = la; l[1] = lb; d[0] = da; d[1] = db;
        return;
      }
      for (int i = 0; i < p.q_tiles; ++i) {
        const int grow = i * 128 + row;
        d[i] = 0.f; l[i] = 0.f;
        if (grow < n) {
          l[i] = p.lse[((size_t)b * p.heads + h) * n + grow];     // raw: scaled at use, so the load stays in flight
          if (p.delta != nullptr) {
            d[i] = p.delta[((size_t)b * n + grow) * p.heads + h];
          } else {
            const bf16* po = p.o + ((size_t)b * n + grow) * p.inner + h * kDh;
            const bf16* pd = p.dout + ((size_t)b * n + grow) * p.inner + h * kDh;
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint4 a = *reinterpret_cast<const uint4*>(po + q * 8);
              const uint4 dd = *reinterpret_cast<const uint4*>(pd + q * 8);
              const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
              const float2 d0 = unpack_bf16x2(dd.x), d1 = unpack_bf16x2(dd.y), d2 = unpack_bf16x2(dd.z), d3 = unpack_bf16x2(dd.w);
              acc += a0.x * d0.x + a0.y * d0.y + a1.x * d1.x + a1.y * d1.y + a2.x * d2.x + a2.y * d2.y +
                     a3.x * d3.x + a3.y * d3.y;
            }
            d[i] = acc;
          }
        }
      }
    };
    // ---- dQ / dK / dV epilogues.  A finished accumulator tile is read from tensor memory (which frees it for the tensor
    // core at once), staged as bf16 in the shared-memory region of the INPUT tile it is the gradient of (dQ_i over Q_i,
    // dK_j over K_j, dV_j over V_j: same shape, and dead once the accumulator is complete) and written with TMA.  The
    // r02 timeline showed why: the thread-per-row global stores (16 bytes per lane at a 1536-byte stride, 32 sectors per
    // instruction) cost the producer warps 2 - 3 k clk per tile, three tiles per item, on the critical path of the
    // kernel; eight conflict-free shared-memory stores per thread cost a few hundred.  All epilogues are deferred until
    // after the NEXT step's P / dS hand-over - across item boundaries as well, where math, hand-over, drains and the
    // statistics fetch used to run back to back while the tensor core waited (11.7 k of 29 k clk per item).
    // the thread of each group that issues its TMA stores: in the LAST lane quadrant, which has no rows in the steps of
    // a partial second query tile and can afford to wait for a store's shared-memory read when a region is needed back
    // at once (quadrant 0 is the critical path of every step)
    // The whole warp runs the bookkeeping (warp-uniform) and the stores are issued behind `elect.sync`, which keeps the
    // TMA operands in uniform registers (`lane == 0` made every UTMASTG a 20-instruction waterfall loop; elect.sync on a
    // full warp always names the same lane, so commit / wait groups stay with one thread).
    const bool issuer = quad == 3;
    const bool cross_p = p.qdo_bufs == 2 && p.kvl_bufs == 2;   // operands of the next item have their own buffers
    uint64_t* rel[4];                                        // (issuer) regions to release once the stores have read them
    int n_rel = 0;
    auto flush_releases = [&]() {
      if (issuer && n_rel > 0) {
        if (elect_one()) {
          tma_wait_group_read<0>();
          for (int k = 0; k < n_rel; ++k) mbar_arrive(rel[k]);
        }
        n_rel = 0;
      }
    };
    // accumulator columns [col, col + 64) of this thread's row -> tensor core may reuse them (free_bar) -> bf16 rows in
    // `region` (rows_loaded x 128 B, SWIZZLE_128B).  Reading + staging one tile at a time keeps 64 registers live, not
    // 128; the group barrier and the TMA stores come once, after BOTH tiles of an item boundary are staged (with a
    // barrier per tile the second accumulator was handed back 1.5 k clk after the first; r02 timeline).
    auto read_and_stage = [&](uint32_t col, uint64_t* free_bar, uint32_t region, int rows_loaded) {
      if (quad * 32 < rows_loaded) {
        uint32_t a0[32], a1[32];
        tmem_ld_32x32(t_row + col, a0);
        tmem_ld_32x32(t_row + col + 32, a1);
        tmem_ld_wait();
        tc_fence_before_sync();
        mbar_arrive(free_bar);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          st_swz_chunk(region, row, c4, make_uint4(pack_bf16x2(__uint_as_float(a0[8 * c4 + 0]), __uint_as_float(a0[8 * c4 + 1])),
                                                   pack_bf16x2(__uint_as_float(a0[8 * c4 + 2]), __uint_as_float(a0[8 * c4 + 3])),
                                                   pack_bf16x2(__uint_as_float(a0[8 * c4 + 4]), __uint_as_float(a0[8 * c4 + 5])),
                                                   pack_bf16x2(__uint_as_float(a0[8 * c4 + 6]), __uint_as_float(a0[8 * c4 + 7]))));
          st_swz_chunk(region, row, 4 + c4, make_uint4(pack_bf16x2(__uint_as_float(a1[8 * c4 + 0]), __uint_as_float(a1[8 * c4 + 1])),
                                                       pack_bf16x2(__uint_as_float(a1[8 * c4 + 2]), __uint_as_float(a1[8 * c4 + 3])),
                                                       pack_bf16x2(__uint_as_float(a1[8 * c4 + 4]), __uint_as_float(a1[8 * c4 + 5])),
                                                       pack_bf16x2(__uint_as_float(a1[8 * c4 + 6]), __uint_as_float(a1[8 * c4 + 7]))));
        }
      } else {
        tc_fence_before_sync();
        mbar_arrive(free_bar);
      }
    };
    auto issue_store = [&](uint32_t region, int rows_loaded, int gcol, int grow0, int gb, uint64_t* release) {
      // (issuer warp, after the group barrier that follows the staging)
      if (elect_one()) {
        for (int r = 0; r < rows_loaded; r += 64) tma_store_3d(&map_dqkv, region + r * 128, gcol, grow0 + r, gb);
        tma_commit_group();
      }
      rel[n_rel++] = release;
      // the early-key-tile region is single-buffered and the next item's K / V load is waiting for it (r02 timeline:
      // released one step later, the load landed 4 k clk after the tensor core wanted it)
      if (release == &bars->kva_empty) flush_releases();
    };
    int pend_j = -1, pend_b = 0, pend_h = 0, pend_it = 0;     // key tile whose dV (group 0) / dK (group 1) is complete
    bool pend_dq = false;                                     // dQ tiles of the previous item
    int dq_b = 0, dq_h = 0, dq_it = 0;
    // epilogues that are due: dV / dK of a finished key tile and, at an item boundary, the dQ tiles of the finished item
    auto drain_pending = [&]() {
      const bool do_kv = pend_j >= 0, do_q = pend_dq;
      if (!do_kv && !do_q) return;
      uint32_t reg_kv = 0, reg_q = 0;
      int rows_kv = 0, rows_q = 0;
      uint64_t *rel_kv = nullptr, *rel_q = nullptr;
      if (do_kv) {
        mbar_wait(&bars->dkv_full, dk & 1);
        tc_fence_after_sync();
        const bool last = pend_j == jl;
        const uint32_t base = last ? smem_u32(sKL + kvl_buf(pend_it) * 2 * p.kvl_region) : smem_u32(sKA + pend_j * 16384);
        reg_kv = base + (wg == 0 ? (last ? p.kvl_region : p.kva_region) : 0);     // V region : K region
        rows_kv = last ? p.k_rows[jl] : 128;
        rel_kv = last ? &bars->kvl_empty[kvl_buf(pend_it)] : &bars->kva_empty;
        read_and_stage(wg == 0 ? kColDV : kColDK, &bars->dkv_free, reg_kv, rows_kv);
        ++dk;
      }
      if (do_q) {                       // group w drains query tile w (the dkv_full wait above covered every product)
        rel_q = &bars->qdo_empty[qdo_buf(dq_it)];
        if (wg < p.q_tiles) {
          reg_q = smem_u32(sQ + qdo_buf(dq_it) * p.q_region + p.q_off[wg]);
          rows_q = p.q_rows[wg];
        }
        read_and_stage(kColDQ + wg * 64, &bars->item_done, reg_q, rows_q);
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + wg, 128);                             // every row of this group's tiles is staged
      if (issuer) {
        if (do_q) {
          if (rows_q > 0) issue_store(reg_q, rows_q, dq_h * kDh, wg * 128, dq_b, rel_q);
          else rel[n_rel++] = rel_q;                           // no tile for this group: only the hand-shake
        }
        if (do_kv) issue_store(reg_kv, rows_kv, (wg == 0 ? 2 : 1) * p.inner + pend_h * kDh, pend_j * 128, pend_b, rel_kv);
      }
      pend_j = -1;
      pend_dq = false;
    };
    if (my_items > 0) fetch_stats(0, delta_n, l2_n);
    for (int it = 0; it < my_items; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int h = item % p.heads, b = item / p.heads;
      delta[0] = delta_n[0]; delta[1] = delta_n[1]; l2[0] = l2_n[0]; l2[1] = l2_n[1];
      bool stats_due = it + 1 < my_items;          // next item's statistics: fetched after this item's first hand-over
      // With ONE step per item the regions released at this item's hand-over would be the ones the item after it is
      // waiting to load into (two buffers, the stores of item it-1 were issued during item it): release them now.
      if (steps_per_item == 1) flush_releases();
      PF(0)
      for (int j = 0; j < p.key_tiles; ++j) {
        for (int i = 0; i < p.q_tiles; ++i, ++g) {
          const int grow = i * 128 + row;
          const bool valid = grow < n;
          const bool warp_rows = (i * 128 + quad * 32) < n;            // any valid row in this warp
          const int key0 = j * 128 + wg * 64;                          // first key of this thread's half
          const bool cols_any = key0 < n;
          const bool interior = (i * 128 + quad * 32 + 32 <= n) && (key0 + 64 <= n);      // warp-uniform
          // this step's row statistics are read from their (runtime-indexed, hence local-memory) arrays BEFORE the
          // wait: behind it the LDL latency sat on the producers' critical path (r01_ncu_attention_hot_lines_v5.txt)
          const float dl = delta[i], lg = l2[i] * kLog2e;
          mbar_wait(&bars->sdp_full, g & 1);
          tc_fence_after_sync();
          PF(1)
          // P / dS of this thread's (row, 64-key half): TMEM -> registers first, so that the S / dP
          // columns can be handed back to the tensor core (next step's products) before the math
          uint32_t pw[32], dw[32];
          if (warp_rows && cols_any) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t sv[32], dv[32];
              tmem_ld_32x32(t_row + kColS + wg * 64 + c * 32, sv);
              tmem_ld_32x32(t_row + kColDP + wg * 64 + c * 32, dv);
              tmem_ld_wait();
              if (c == 1) {
                tc_fence_before_sync();
                mbar_arrive(&bars->sdp_free);
              }
              if (interior) {
                // every (row, key) of this warp's 32 x 64 block is inside the sequence (all blocks of n = 192 are either
                // this or skipped): no per-element masks, packed fp32 arithmetic (two elements per FFMA2 / FMUL2 / FADD2)
                const f32x2 sl2_2 = f2_splat(sl2), nlg2 = f2_splat(-lg), ndl2 = f2_splat(-dl), sc2 = f2_splat(p.scale);
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                  const f32x2 t = f2_fma(f2_packu(sv[e], sv[e + 1]), sl2_2, nlg2);
                  float t0, t1;
                  f2_unpack(t, t0, t1);
                  const float p0 = exp2f(t0), p1 = exp2f(t1);
                  const f32x2 pp = f2_pack(p0, p1);
                  const f32x2 ds = f2_mul(f2_mul(pp, f2_add(f2_packu(dv[e], dv[e + 1]), ndl2)), sc2);
                  float d0, d1;
                  f2_unpack(ds, d0, d1);
                  pw[c * 16 + (e >> 1)] = pack_bf16x2(p0, p1);
                  dw[c * 16 + (e >> 1)] = pack_bf16x2(d0, d1);
                }
              } else {
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                  const int key = key0 + c * 32 + e;
                  const float p0 = exp2f(fmaf(__uint_as_float(sv[e]), sl2, -lg));
                  const float p1 = exp2f(fmaf(__uint_as_float(sv[e + 1]), sl2, -lg));
                  const bool ok0 = valid && key < n, ok1 = valid && key + 1 < n;
                  const float q0 = ok0 ? p0 : 0.f, q1 = ok1 ? p1 : 0.f;
                  pw[c * 16 + (e >> 1)] = pack_bf16x2(q0, q1);
                  dw[c * 16 + (e >> 1)] = pack_bf16x2(q0 * (__uint_as_float(dv[e]) - dl) * p.scale,
                                                      q1 * (__uint_as_float(dv[e + 1]) - dl) * p.scale);
                }
              }
            }
          } else {
            tc_fence_before_sync();
            mbar_arrive(&bars->sdp_free);
#pragma unroll
            for (int e = 0; e < 32; ++e) pw[e] = dw[e] = 0u;
          }
          PF(2)
          // the slabs are still being read by the previous step's dV / dK / dQ products
          if (g > 0) mbar_wait(&bars->pds_free, (g - 1) & 1);
          PF(3)
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            st_swz_chunk(p_slab, row, ch, make_uint4(pw[4 * ch], pw[4 * ch + 1], pw[4 * ch + 2], pw[4 * ch + 3]));
            st_swz_chunk(ds_slab, row, ch, make_uint4(dw[4 * ch], dw[4 * ch + 1], dw[4 * ch + 2], dw[4 * ch + 3]));
          }
          fence_proxy_async_smem();
          tc_fence_before_sync();
          mbar_arrive(&bars->pds_full);
          PF(4)
          // ---- epilogues of accumulators completed by EARLIER steps, behind this step's hand-over
          flush_releases();                 // (stores issued a step or more ago have long read their staging regions)
          drain_pending();
          PF(5)
          PF(6)
          if (stats_due) { fetch_stats(it + 1, delta_n, l2_n); stats_due = false; }
          PF(7)
          if (i == p.q_tiles - 1) { pend_j = j; pend_b = b; pend_h = h; pend_it = it; }   // dV_j / dK_j: after the next hand-over
        }
      }
      pend_dq = true; dq_b = b; dq_h = h; dq_it = it;
      if (!cross_p || it + 1 == my_items) {
        // the next item's operands reuse this item's regions (single buffers), or there is no next item: finish now
        drain_pending();
        flush_releases();
      }
    }
#undef PF
    if (p.prof && blockIdx.x == 0 && (warp == kProbeW0 || warp == kProbeW1) && lane == 0) {
      const int o = warp == kProbeW0 ? 0 : 16;
      for (int k = 0; k < 8; ++k) p.prof[o + k] = pf[k];
      p.prof[o + 8] = g;
      p.prof[o + 9] = M3L_CLK() - pf_t0;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// short sequences (n <= 16, e.g. the 10 visible tokens the masked encoder sees): one warp per
// (sample, head) on mma.sync m16n8k16 (bf16 x bf16 -> fp32) with every tile padded to 16 rows.
// A 128-row tcgen05 tile would be > 90 % padding here and the persistent pipeline's barrier round
// trips dominate (measured 14.6 us forward / 32 us backward for 1024 items of n = 10); a first CUDA-core
// version (lane = query row, 10 of 32 lanes busy, ~8 k dependent instructions per item) took 10 / 25 us.
// Here an item is 24 (forward) / 56 (backward) tensor instructions:
//   forward : S = Q K^T, row softmax on the accumulator fragments (4 lanes share a row), the bf16 P fragments
//             ARE the A operand of O = P V (accumulator layout == A layout), V through ldmatrix.trans
//   backward: S, dP = dO V^T and their transposes S^T = K Q^T, dP^T = V dO^T (8 instructions each - cheaper
//             than transposing fragments across lanes); P / dS feed dQ = dS K, P^T / dS^T feed dV = P^T dO and
//             dK = dS^T Q; delta_i = rowsum(P * dP) is shuffled to the lanes that hold column i of the transposes
// ------------------------------------------------------------------------------------------
constexpr int kSmallMaxN = 16;
constexpr int kSmallPitch = kDh + 8;    // bf16 row pitch 144 B: fragment reads and ldmatrix rows are conflict free
constexpr int kSmallTile = kSmallMaxN * kSmallPitch;      // bf16 elements per staged matrix

M3L_DEVINL void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 16 rows x 64 bf16 of one (sample, head) slice -> staged tile; rows >= n are zero
M3L_DEVINL void small_stage(const bf16* g, int ld, int n, bf16* dst, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = lane + 32 * i, row = c >> 3, ch = c & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row < n) v = *reinterpret_cast<const uint4*>(g + (size_t)row * ld + ch * 8);
    *reinterpret_cast<uint4*>(dst + row * kSmallPitch + ch * 8) = v;
  }
}
// A fragment (16 x 16, rows of T, k = columns k0 .. k0 + 15)
M3L_DEVINL void small_frag_a(const bf16* T, int k0, int lane, uint32_t (&a)[4]) {
  const int g = lane >> 2, t = lane & 3;
  a[0] = *reinterpret_cast<const uint32_t*>(T + g * kSmallPitch + k0 + 2 * t);
  a[1] = *reinterpret_cast<const uint32_t*>(T + (g + 8) * kSmallPitch + k0 + 2 * t);
  a[2] = *reinterpret_cast<const uint32_t*>(T + g * kSmallPitch + k0 + 8 + 2 * t);
  a[3] = *reinterpret_cast<const uint32_t*>(T + (g + 8) * kSmallPitch + k0 + 8 + 2 * t);
}
// B fragment with B[k][n] = T[n0 + n][k0 + k] (T rows along n, contiguous in k)
M3L_DEVINL void small_frag_b(const bf16* T, int n0, int k0, int lane, uint32_t& b0, uint32_t& b1) {
  const int g = lane >> 2, t = lane & 3;
  b0 = *reinterpret_cast<const uint32_t*>(T + (n0 + g) * kSmallPitch + k0 + 2 * t);
  b1 = *reinterpret_cast<const uint32_t*>(T + (n0 + g) * kSmallPitch + k0 + 8 + 2 * t);
}
// two B fragments with B[k][n] = T[k][n0 + n] (T rows along k = 0 .. 15): n tiles n0 and n0 + 8
M3L_DEVINL void small_frag_bt(const bf16* T, int n0, int lane, uint32_t (&r)[4]) {
  const int row = (lane & 7) + ((lane >> 3) & 1) * 8, col = n0 + (lane >> 4) * 8;
  const uint32_t addr = smem_u32(T + row * kSmallPitch + col);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// C[16 x 16] = X Y^T over the 64 feature columns (X rows = output rows, Y rows = output columns)
M3L_DEVINL void small_xyt(const bf16* X, const bf16* Y, int lane, float (&c)[2][4]) {
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) c[nt][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < kDh / 16; ++ks) {
    uint32_t a[4];
    small_frag_a(X, ks * 16, lane, a);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      uint32_t b0, b1;
      small_frag_b(Y, nt * 8, ks * 16, lane, b0, b1);
      mma_bf16_16816(c[nt], a, b0, b1);
    }
  }
}
// C[16 x 64] = A(16 x 16 fragments) * T (16 rows k, 64 columns)
M3L_DEVINL void small_ax(const uint32_t (&a)[4], const bf16* T, int lane, float (&o)[8][4]) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[nt][e] = 0.f;
#pragma unroll
  for (int np = 0; np < 4; ++np) {
    uint32_t r[4];
    small_frag_bt(T, np * 16, lane, r);
    mma_bf16_16816(o[2 * np], a, r[0], r[1]);
    mma_bf16_16816(o[2 * np + 1], a, r[2], r[3]);
  }
}
// rows g / g + 8 of a 16 x 64 accumulator -> bf16 global rows (row stride ld), scaled per row
M3L_DEVINL void small_store_rows(bf16* dst, int ld, int n, int lane, const float (&o)[8][4], float s_lo, float s_hi) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (g < n)
      *reinterpret_cast<uint32_t*>(dst + (size_t)g * ld + nt * 8 + 2 * t) = pack_bf16x2(o[nt][0] * s_lo, o[nt][1] * s_lo);
    if (g + 8 < n)
      *reinterpret_cast<uint32_t*>(dst + (size_t)(g + 8) * ld + nt * 8 + 2 * t) = pack_bf16x2(o[nt][2] * s_hi, o[nt][3] * s_hi);
  }
}

__global__ void __launch_bounds__(128)
attn_small_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse, int n,
                      int heads, int inner, int num_items, float scale) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t sm_small_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  bf16* sQ = reinterpret_cast<bf16*>(sm_small_raw) + (size_t)warp * 3 * kSmallTile;
  bf16* sK = sQ + kSmallTile;
  bf16* sV = sK + kSmallTile;
  const int ld = 3 * inner;
  const int g = lane >> 2, t = lane & 3;
  for (int item = blockIdx.x * wpb + warp; item < num_items; item += gridDim.x * wpb) {
    const int h = item % heads, b = item / heads;
    const bf16* base = qkv + (size_t)b * n * ld + h * kDh;
    __syncwarp();
    small_stage(base, ld, n, sQ, lane);
    small_stage(base + inner, ld, n, sK, lane);
    small_stage(base + 2 * inner, ld, n, sV, lane);
    __syncwarp();
    float sc[2][4];
    small_xyt(sQ, sK, lane, sc);                    // sc[nt][0,1]: row g, cols nt*8 + 2t, +1; [2,3]: row g + 8
    float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = nt * 8 + 2 * t + e < n;
        sc[nt][e] = ok ? sc[nt][e] * scale : -INFINITY;
        sc[nt][2 + e] = ok ? sc[nt][2 + e] * scale : -INFINITY;
        m_lo = fmaxf(m_lo, sc[nt][e]);
        m_hi = fmaxf(m_hi, sc[nt][2 + e]);
      }
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1)); m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1)); m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
    float s_lo = 0.f, s_hi = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        sc[nt][e] = __expf(sc[nt][e] - m_lo);       // exp(-inf) = 0 for the padded key columns
        sc[nt][2 + e] = __expf(sc[nt][2 + e] - m_hi);
        s_lo += sc[nt][e];
        s_hi += sc[nt][2 + e];
      }
    s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 1); s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 2);
    s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 1); s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 2);
    const uint32_t pa[4] = {pack_bf16x2(sc[0][0], sc[0][1]), pack_bf16x2(sc[0][2], sc[0][3]),
                            pack_bf16x2(sc[1][0], sc[1][1]), pack_bf16x2(sc[1][2], sc[1][3])};
    float o[8][4];
    small_ax(pa, sV, lane, o);
    small_store_rows(out + (size_t)b * n * inner + h * kDh, inner, n, lane, o, 1.0f / s_lo, 1.0f / s_hi);
    if (lse && t == 0) {
      float* l = lse + ((size_t)b * heads + h) * n;
      if (g < n) l[g] = m_lo + __logf(s_lo);
      if (g + 8 < n) l[g + 8] = m_hi + __logf(s_hi);
    }
  }
}

__global__ void __launch_bounds__(128)
attn_small_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, const float* __restrict__ lse,
                      bf16* __restrict__ dqkv, int n, int heads, int inner, int num_items, float scale) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t sm_small_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  bf16* sQ = reinterpret_cast<bf16*>(sm_small_raw) + (size_t)warp * 4 * kSmallTile;
  bf16* sK = sQ + kSmallTile;
  bf16* sV = sK + kSmallTile;
  bf16* sDO = sV + kSmallTile;
  const int ld = 3 * inner;
  const int g = lane >> 2, t = lane & 3;
  for (int item = blockIdx.x * wpb + warp; item < num_items; item += gridDim.x * wpb) {
    const int h = item % heads, b = item / heads;
    const bf16* base = qkv + (size_t)b * n * ld + h * kDh;
    __syncwarp();
    small_stage(base, ld, n, sQ, lane);
    small_stage(base + inner, ld, n, sK, lane);
    small_stage(base + 2 * inner, ld, n, sV, lane);
    small_stage(dout + (size_t)b * n * inner + h * kDh, inner, n, sDO, lane);
    __syncwarp();
    const float* l = lse + ((size_t)b * heads + h) * n;
    bf16* dbase = dqkv + (size_t)b * n * ld + h * kDh;
    // ---- query-major pass: P, dP, delta, dS -> dQ
    float s[2][4], dp[2][4];
    small_xyt(sQ, sK, lane, s);
    small_xyt(sDO, sV, lane, dp);
    const float l_lo = g < n ? l[g] : 0.f, l_hi = g + 8 < n ? l[g + 8] : 0.f;
    float d_lo = 0.f, d_hi = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool col_ok = nt * 8 + 2 * t + e < n;
        s[nt][e] = (col_ok && g < n) ? __expf(s[nt][e] * scale - l_lo) : 0.f;
        s[nt][2 + e] = (col_ok && g + 8 < n) ? __expf(s[nt][2 + e] * scale - l_hi) : 0.f;
        d_lo = fmaf(s[nt][e], dp[nt][e], d_lo);
        d_hi = fmaf(s[nt][2 + e], dp[nt][2 + e], d_hi);
      }
    d_lo += __shfl_xor_sync(0xffffffffu, d_lo, 1); d_lo += __shfl_xor_sync(0xffffffffu, d_lo, 2);
    d_hi += __shfl_xor_sync(0xffffffffu, d_hi, 1); d_hi += __shfl_xor_sync(0xffffffffu, d_hi, 2);
    {
      uint32_t dsa[4];
      dsa[0] = pack_bf16x2(s[0][0] * (dp[0][0] - d_lo) * scale, s[0][1] * (dp[0][1] - d_lo) * scale);
      dsa[1] = pack_bf16x2(s[0][2] * (dp[0][2] - d_hi) * scale, s[0][3] * (dp[0][3] - d_hi) * scale);
      dsa[2] = pack_bf16x2(s[1][0] * (dp[1][0] - d_lo) * scale, s[1][1] * (dp[1][1] - d_lo) * scale);
      dsa[3] = pack_bf16x2(s[1][2] * (dp[1][2] - d_hi) * scale, s[1][3] * (dp[1][3] - d_hi) * scale);
      float dq[8][4];
      small_ax(dsa, sK, lane, dq);
      small_store_rows(dbase, ld, n, lane, dq, 1.f, 1.f);
    }
    // ---- key-major pass: P^T, dS^T (rows = keys j, columns = queries i) -> dV, dK
    // this lane's columns are queries 2t, 2t+1, 2t+8, 2t+9: their LSE / delta come from the lanes that own those rows
    const int src0 = (2 * t) * 4, src1 = (2 * t + 1) * 4;
    const float lc[4] = {__shfl_sync(0xffffffffu, l_lo, src0), __shfl_sync(0xffffffffu, l_lo, src1),
                         __shfl_sync(0xffffffffu, l_hi, src0), __shfl_sync(0xffffffffu, l_hi, src1)};
    const float dc[4] = {__shfl_sync(0xffffffffu, d_lo, src0), __shfl_sync(0xffffffffu, d_lo, src1),
                         __shfl_sync(0xffffffffu, d_hi, src0), __shfl_sync(0xffffffffu, d_hi, src1)};
    small_xyt(sK, sQ, lane, s);                      // S^T
    small_xyt(sV, sDO, lane, dp);                    // dP^T
    uint32_t pta[4], dsta[4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      float pv[4], dv_[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + 2 * t + (e & 1);    // query index
        const int row = g + (e >> 1) * 8;            // key index
        const float lcv = lc[nt * 2 + (e & 1)], dcv = dc[nt * 2 + (e & 1)];
        const float pp = (col < n && row < n) ? __expf(s[nt][e] * scale - lcv) : 0.f;
        pv[e] = pp;
        dv_[e] = pp * (dp[nt][e] - dcv) * scale;
      }
      pta[nt * 2] = pack_bf16x2(pv[0], pv[1]);
      pta[nt * 2 + 1] = pack_bf16x2(pv[2], pv[3]);
      dsta[nt * 2] = pack_bf16x2(dv_[0], dv_[1]);
      dsta[nt * 2 + 1] = pack_bf16x2(dv_[2], dv_[3]);
    }
    {
      float acc[8][4];
      small_ax(pta, sDO, lane, acc);                 // dV = P^T dO
      small_store_rows(dbase + 2 * inner, ld, n, lane, acc, 1.f, 1.f);
      small_ax(dsta, sQ, lane, acc);                 // dK = dS^T Q
      small_store_rows(dbase + inner, ld, n, lane, acc, 1.f, 1.f);
    }
  }
}

long long* attn_prof_buf() {
  static long long* buf = [] {
    long long* b = nullptr;
    if (getenv("M3L_ATTN_PROF")) {
      cudaMalloc(&b, 2048 * sizeof(long long));
      cudaMemset(b, 0, 2048 * sizeof(long long));
      const long long flag = getenv("M3L_ATTN_SERIAL") ? 1 : 0;     // [32]: wait for every MMA batch
      cudaMemcpy(b + 32, &flag, sizeof(flag), cudaMemcpyHostToDevice);
    }
    return b;
  }();
  return buf;
}

int attn_check(int n, int heads, int dim_head, int batch) {
  M3L_REQUIRE(dim_head == kDh, "attention: dim_head=%d unsupported (only 64)", dim_head);
  M3L_REQUIRE(n >= 1 && n <= 256, "attention: sequence length %d unsupported (1..256)", n);
  M3L_REQUIRE(heads >= 1 && batch >= 0, "attention: bad heads/batch");
  return M3L_OK;
}

}  // namespace
}  // namespace m3l

using namespace m3l;

extern "C" int m3l_attention_fwd(const void* qkv_bf16, int batch, int n, int heads, int dim_head, float scale,
                                 void* out_bf16, float* lse, void* stream) {
  M3L_REQUIRE(qkv_bf16 && out_bf16, "attention_fwd: null pointer");
  int s = attn_check(n, heads, dim_head, batch);
  if (s) return s;
  if (batch == 0) return M3L_OK;
  const int inner = heads * kDh;
  if (n <= kSmallMaxN) {
    const int items = batch * heads, wpb = 4;
    const size_t smem = (size_t)wpb * 3 * kSmallTile * sizeof(bf16);
    const int grid = std::min((items + wpb - 1) / wpb, device_sm_count() * 8);
    M3L_CUDA(launch_kernel(attn_small_fwd_kernel, dim3(grid), dim3(wpb * 32), smem, (cudaStream_t)stream, (const bf16*)qkv_bf16, (bf16*)out_bf16, lse, n,
                                                                         heads, inner, items, scale));
    M3L_CUDA(cudaGetLastError());
    return M3L_OK;
  }
  const int NK = (n + 15) & ~15;
  CUtensorMap map_q, map_kv;
  s = make_tmap_3d_bf16(&map_q, qkv_bf16, 3 * inner, n, batch, 3 * inner, (uint64_t)n * 3 * inner, 128);
  if (s) return s;
  s = make_tmap_3d_bf16(&map_kv, qkv_bf16, 3 * inner, n, batch, 3 * inner, (uint64_t)n * 3 * inner, NK);
  if (s) return s;
  AttnFwdParams p;
  p.out = (bf16*)out_bf16; p.lse = lse; p.n = n; p.heads = heads; p.inner = inner; p.scale = scale;
  p.prof = attn_prof_buf();
  p.q_tiles = (n + 127) / 128;
  const int per_tile = (n + p.q_tiles - 1) / p.q_tiles;
  p.tile_rows = std::min(128, (per_tile + 31) & ~31);
  p.num_items = batch * heads;
  p.slot_cols = 64;
  while (p.slot_cols < NK) p.slot_cols <<= 1;
  const int kv_region = (NK * 128 + 1023) & ~1023;
  const int p_slabs = (NK + 63) / 64;
  const int fixed = 1024 + 2 * 16384 + 2 * p_slabs * 16384 + 256;
  p.kv_bufs = (fixed + 2 * 2 * kv_region <= 227 * 1024) ? 2 : 1;
  const int smem = fixed + p.kv_bufs * 2 * kv_region;
  M3L_REQUIRE(smem <= 227 * 1024, "attention_fwd: shared memory budget exceeded (n=%d)", n);
  static bool configured = false;
  if (!configured) {
    M3L_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int grid = std::min(p.num_items, device_sm_count());
  M3L_CUDA(launch_kernel(attn_fwd_kernel, dim3(grid), dim3(320), smem, (cudaStream_t)stream, map_q, map_kv, p));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_attention_bwd(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16,
                                 const float* lse, const float* delta, int batch, int n, int heads, int dim_head,
                                 float scale, void* dqkv_bf16, void* stream) {
  M3L_REQUIRE(qkv_bf16 && out_bf16 && dout_bf16 && lse && dqkv_bf16, "attention_bwd: null pointer");
  int s = attn_check(n, heads, dim_head, batch);
  if (s) return s;
  if (batch == 0) return M3L_OK;
  const int inner = heads * kDh;
  if (n <= kSmallMaxN) {
    const int items = batch * heads, wpb = 4;
    const size_t smem = (size_t)wpb * 4 * kSmallTile * sizeof(bf16);
    const int grid = std::min((items + wpb - 1) / wpb, device_sm_count() * 8);
    M3L_CUDA(launch_kernel(attn_small_bwd_kernel, dim3(grid), dim3(wpb * 32), smem, (cudaStream_t)stream, (const bf16*)qkv_bf16, (const bf16*)dout_bf16, lse,
                                                                         (bf16*)dqkv_bf16, n, heads, inner, items, scale));
    M3L_CUDA(cudaGetLastError());
    return M3L_OK;
  }
  CUtensorMap map_qkv, map_do;
  s = make_tmap_3d_bf16(&map_qkv, qkv_bf16, 3 * inner, n, batch, 3 * inner, (uint64_t)n * 3 * inner, 64);
  if (s) return s;
  s = make_tmap_3d_bf16(&map_do, dout_bf16, inner, n, batch, inner, (uint64_t)n * inner, 64);
  if (s) return s;
  CUtensorMap map_dqkv;      // same geometry as map_qkv: gradient tiles leave through 64-row TMA stores
  s = make_tmap_3d_bf16(&map_dqkv, dqkv_bf16, 3 * inner, n, batch, 3 * inner, (uint64_t)n * 3 * inner, 64);
  if (s) return s;
  AttnBwdParams p;
  p.o = (const bf16*)out_bf16; p.dout = (const bf16*)dout_bf16; p.lse = lse; p.delta = delta; p.dqkv = (bf16*)dqkv_bf16;
  p.n = n; p.heads = heads; p.inner = inner; p.scale = scale;
  p.prof = attn_prof_buf();
  p.q_tiles = (n + 127) / 128;
  p.key_tiles = p.q_tiles;
  p.num_items = batch * heads;
  p.q_region = 0;
  for (int i = 0; i < 2; ++i) {
    const int rows = i < p.q_tiles ? std::min(128, ((n - 128 * i) + 63) & ~63) : 0;   // 64-row TMA boxes
    p.q_rows[i] = rows; p.k_rows[i] = rows;
    p.q_off[i] = p.q_region;
    p.q_region += rows * 128;
  }
  const int jl = p.key_tiles - 1;
  p.kva_region = jl * 16384;
  p.kvl_region = p.k_rows[jl] * 128;
  // a query tile of 64 loaded rows is still read as a 128-row MMA operand: the over-read must stay inside
  // the Q / dO regions (the P slabs follow them)
  const int fixed = 1024 + 4 * 16384 + 2 * p.kva_region + 512;
  int smem = 0;
  p.qdo_bufs = p.kvl_bufs = 0;
  const int tries[4][2] = {{2, 2}, {1, 2}, {2, 1}, {1, 1}};
  for (int t = 0; t < 4 && p.qdo_bufs == 0; ++t) {
    const int need = fixed + tries[t][0] * 2 * p.q_region + tries[t][1] * 2 * p.kvl_region;
    if (need <= 227 * 1024) { p.qdo_bufs = tries[t][0]; p.kvl_bufs = tries[t][1]; smem = need; }
  }
  M3L_REQUIRE(p.qdo_bufs != 0, "attention_bwd: shared memory budget exceeded (n=%d)", n);
  static bool configured = false;
  if (!configured) {
    M3L_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int grid = std::min(p.num_items, device_sm_count());
  M3L_CUDA(launch_kernel(attn_bwd_kernel, dim3(grid), dim3(320), smem, (cudaStream_t)stream, map_qkv, map_do, map_dqkv, p));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

// measurement only (M3L_ATTN_PROF=1): read and reset the in-kernel cycle counters of CTA 0
extern "C" int m3l_debug_attn_prof(long long* host_out, int n) {
  long long* b = m3l::attn_prof_buf();
  if (b == nullptr || n > 2048) return M3L_ERR_INVALID;
  cudaDeviceSynchronize();
  cudaMemcpy(host_out, b, n * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaMemset(b, 0, 32 * sizeof(long long));
  return M3L_OK;
}
