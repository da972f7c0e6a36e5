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
#include "common.cuh"
#include "m3l_internal.h"

namespace m3l {
namespace {

constexpr int kDh = 64;
constexpr float kLog2e = 1.4426950408889634f;

M3L_DEVINL uint32_t round_up_pow2_cols(int c) {
  uint32_t r = 32;
  while ((int)r < c) r <<= 1;
  return r;
}

struct AttnFwdBars {
  uint64_t qk, v, s, p, o;
  uint32_t tmem_base;
};

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
// forward
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                bf16* __restrict__ out, float* __restrict__ lse, int n, int heads, int inner,
                int q_tiles, float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int NK = (n + 15) & ~15;                 // keys padded to the MMA granularity
  const int kv_bytes = NK * 128;                 // [NK rows x 64 d] bf16
  const int kv_region = (kv_bytes + 1023) & ~1023;
  const int p_slabs = (NK + 63) / 64;
  const int regA = max(16384 + kv_region, p_slabs * 16384);   // (Q | K) aliased with P
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 16384;
  uint8_t* sP = smem;
  uint8_t* sV = smem + regA;
  AttnFwdBars* bars = reinterpret_cast<AttnFwdBars*>(sV + kv_region);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % q_tiles;
  const int bh = blockIdx.x / q_tiles;
  const int h = bh % heads, b = bh / heads;
  const uint32_t tmem_cols = round_up_pow2_cols(max(NK, 64));

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&map_q);
      tma_prefetch_desc(&map_kv);
      mbar_init(&bars->qk, 1);
      mbar_init(&bars->v, 1);
      mbar_init(&bars->s, 1);
      mbar_init(&bars->p, 128);
      mbar_init(&bars->o, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&bars->tmem_base, tmem_cols);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars->qk, 16384 + kv_bytes);
      tma_load_3d(sQ, &map_q, &bars->qk, h * kDh, qt * 128, b);
      tma_load_3d(sK, &map_kv, &bars->qk, inner + h * kDh, 0, b);
      mbar_arrive_expect_tx(&bars->v, kv_bytes);
      tma_load_3d(sV, &map_kv, &bars->v, 2 * inner + h * kDh, 0, b);
      // S = Q K^T
      mbar_wait(&bars->qk, 0);
      tc_fence_after_sync();
      const uint32_t idesc_s = umma_idesc_bf16(128, NK, 0, 0);
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK);
#pragma unroll
      for (int k = 0; k < kDh / 16; ++k)
        umma_bf16(tmem_base, umma_smem_desc(q_addr + k * 32, 16, 1024),
                  umma_smem_desc(k_addr + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
      umma_commit(&bars->s);
      // O = P V
      mbar_wait(&bars->p, 0);
      tc_fence_after_sync();
      mbar_wait(&bars->v, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, kDh, 0, 1);
      const uint32_t p_addr = smem_u32(sP), v_addr = smem_u32(sV);
      for (int kk = 0; kk < NK / 16; ++kk)
        umma_bf16(tmem_base, umma_smem_desc(p_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                  umma_smem_desc(v_addr + kk * 2048, 8192, 1024), idesc_o, kk > 0 ? 1u : 0u);
      umma_commit(&bars->o);
    }
  } else {
    const int row = warp * 32 + lane;            // TMEM lane == query row inside the tile
    const int grow = qt * 128 + row;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const float sl2 = scale * kLog2e;
    mbar_wait(&bars->s, 0);
    tc_fence_after_sync();
    const int nchunks = (NK + 31) / 32;
    float mx = -INFINITY;
    for (int c = 0; c < nchunks; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(t_row + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c * 32 + j < n) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    float sum = 0.f;
    const uint32_t p_u32 = smem_u32(sP);
    for (int c = 0; c < nchunks; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(t_row + c * 32, v);
      tmem_ld_wait();
      float pv[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float p = (c * 32 + j < n) ? exp2f((__uint_as_float(v[j]) - mx) * sl2) : 0.f;
        sum += p;
        pv[j] = p;
      }
      const int col0 = c * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (col0 + g * 8 < NK) {
          uint4 u;
          u.x = pack_bf16x2(pv[g * 8 + 0], pv[g * 8 + 1]);
          u.y = pack_bf16x2(pv[g * 8 + 2], pv[g * 8 + 3]);
          u.z = pack_bf16x2(pv[g * 8 + 4], pv[g * 8 + 5]);
          u.w = pack_bf16x2(pv[g * 8 + 6], pv[g * 8 + 7]);
          const int col = col0 + g * 8;
          st_swz_chunk(p_u32 + (col >> 6) * 16384, row, (col & 63) >> 3, u);
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    mbar_arrive(&bars->p);
    // epilogue
    mbar_wait(&bars->o, 0);
    tc_fence_after_sync();
    const float inv = 1.0f / sum;
    uint32_t o0[32], o1[32];
    tmem_ld_32x32(t_row, o0);
    tmem_ld_32x32(t_row + 32, o1);
    tmem_ld_wait();
    if (grow < n) {
      bf16* dst = out + ((size_t)b * n + grow) * inner + h * kDh;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(o0[g * 8 + 0]) * inv, __uint_as_float(o0[g * 8 + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(o0[g * 8 + 2]) * inv, __uint_as_float(o0[g * 8 + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(o0[g * 8 + 4]) * inv, __uint_as_float(o0[g * 8 + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(o0[g * 8 + 6]) * inv, __uint_as_float(o0[g * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(dst + g * 8) = u;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(o1[g * 8 + 0]) * inv, __uint_as_float(o1[g * 8 + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(o1[g * 8 + 2]) * inv, __uint_as_float(o1[g * 8 + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(o1[g * 8 + 4]) * inv, __uint_as_float(o1[g * 8 + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(o1[g * 8 + 6]) * inv, __uint_as_float(o1[g * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(dst + 32 + g * 8) = u;
      }
      if (lse) lse[((size_t)b * heads + h) * n + grow] = mx * scale + __logf(sum);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
struct AttnBwdBars {
  uint64_t kv;        // K and V landed
  uint64_t qdo;       // Q tile and dO tile landed (per q tile)
  uint64_t mma;       // MMA group finished (reused; phases tracked)
  uint64_t warps;     // softmax warps finished a stage (count 128)
  uint32_t tmem_base;
};

// TMEM column map (512 allocated): [0, 256) scratch (S_t, then dP_t, then dQ_t), [256, 320) dV keys 0..127,
// [320, 384) dV keys 128..255, [384, 448) dK keys 0..127, [448, 512) dK keys 128..255.
__global__ void __launch_bounds__(160)
attn_bwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                const __grid_constant__ CUtensorMap map_do, const bf16* __restrict__ o,
                const bf16* __restrict__ dout, const float* __restrict__ lse, bf16* __restrict__ dqkv,
                int n, int heads, int inner, float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int NK = (n + 15) & ~15;
  const int kv_bytes = NK * 128;
  const int kv_region = (kv_bytes + 1023) & ~1023;
  const int slabs = 2 * ((NK + 127) / 128);  // 64-key slabs, whole 128-key tiles (unused keys zeroed)
  uint8_t* sQ = smem;                       // 16 KB  [128 q][64 d]
  uint8_t* sDO = sQ + 16384;                // 16 KB  [128 q][64 d]
  uint8_t* sK = sDO + 16384;                // [NK][64]
  uint8_t* sV = sK + kv_region;             // [NK][64]
  uint8_t* sP = sV + kv_region;             // slabs x 16 KB  [128 q][64 keys]
  uint8_t* sDS = sP + slabs * 16384;        // slabs x 16 KB
  AttnBwdBars* bars = reinterpret_cast<AttnBwdBars*>(sDS + slabs * 16384);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % heads, b = blockIdx.x / heads;
  const int q_tiles = (n + 127) / 128;
  const int key_tiles = (NK + 127) / 128;
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kColDV = 256, kColDK = 384;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&map_q);
      tma_prefetch_desc(&map_kv);
      tma_prefetch_desc(&map_do);
      mbar_init(&bars->kv, 1);
      mbar_init(&bars->qdo, 1);
      mbar_init(&bars->mma, 1);
      mbar_init(&bars->warps, 128);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&bars->tmem_base, kTmemCols);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const float sl2 = scale * kLog2e;

  if (warp == 4) {
    if (lane == 0) {
      uint32_t ph_mma = 0, ph_warps = 0;
      mbar_arrive_expect_tx(&bars->kv, 2 * kv_bytes);
      tma_load_3d(sK, &map_kv, &bars->kv, inner + h * kDh, 0, b);
      tma_load_3d(sV, &map_kv, &bars->kv, 2 * inner + h * kDh, 0, b);
      const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sDO), k_addr = smem_u32(sK),
                     v_addr = smem_u32(sV), p_addr = smem_u32(sP), ds_addr = smem_u32(sDS);
      const uint32_t idesc_s = umma_idesc_bf16(128, NK, 0, 0);       // S / dP: K-major x K-major
      const uint32_t idesc_dq = umma_idesc_bf16(128, kDh, 0, 1);     // dQ: A K-major, B (=K) MN-major
      const uint32_t idesc_dkv = umma_idesc_bf16(128, kDh, 1, 1);    // dK/dV: both MN-major
      for (int t = 0; t < q_tiles; ++t) {
        if (t > 0) {
          // previous tile's dQ epilogue + operand reads must be finished before Q/dO are overwritten
          mbar_wait(&bars->warps, ph_warps); ph_warps ^= 1;
        }
        mbar_arrive_expect_tx(&bars->qdo, 2 * 16384);
        tma_load_3d(sQ, &map_q, &bars->qdo, h * kDh, t * 128, b);
        tma_load_3d(sDO, &map_do, &bars->qdo, h * kDh, t * 128, b);
        if (t == 0) mbar_wait(&bars->kv, 0);
        mbar_wait(&bars->qdo, t & 1);
        tc_fence_after_sync();
        // ---- S_t = Q_t K^T -> scratch
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k)
          umma_bf16(tmem_base, umma_smem_desc(q_addr + k * 32, 16, 1024),
                    umma_smem_desc(k_addr + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&bars->mma);
        // warps: read S, write P
        mbar_wait(&bars->warps, ph_warps); ph_warps ^= 1;
        tc_fence_after_sync();
        // ---- dP_t = dO_t V^T -> scratch
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k)
          umma_bf16(tmem_base, umma_smem_desc(do_addr + k * 32, 16, 1024),
                    umma_smem_desc(v_addr + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&bars->mma);
        // warps: read dP, write dS
        mbar_wait(&bars->warps, ph_warps); ph_warps ^= 1;
        tc_fence_after_sync();
        // ---- dQ_t = dS_t K   (A = dS K-major slabs, B = K as MN-major [keys][d])
        for (int kk = 0; kk < NK / 16; ++kk)
          umma_bf16(tmem_base, umma_smem_desc(ds_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                    umma_smem_desc(k_addr + kk * 2048, 8192, 1024), idesc_dq, kk > 0 ? 1u : 0u);
        // ---- dV += P_t^T dO_t ; dK += dS_t^T Q_t   (A = slabs as MN-major [q][keys], 128 keys per tile)
        for (int kt = 0; kt < key_tiles; ++kt) {
          for (int kk = 0; kk < 128 / 16; ++kk) {       // contraction over the 128 query rows
            const uint32_t acc = (t > 0 || kk > 0) ? 1u : 0u;
            umma_bf16(tmem_base + kColDV + kt * 64,
                      umma_smem_desc(p_addr + kt * 2 * 16384 + kk * 2048, 16384, 1024),
                      umma_smem_desc(do_addr + kk * 2048, 8192, 1024), idesc_dkv, acc);
            umma_bf16(tmem_base + kColDK + kt * 64,
                      umma_smem_desc(ds_addr + kt * 2 * 16384 + kk * 2048, 16384, 1024),
                      umma_smem_desc(q_addr + kk * 2048, 8192, 1024), idesc_dkv, acc);
          }
        }
        umma_commit(&bars->mma);
      }
      (void)ph_mma;
    }
  } else {
    const int row = warp * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t p_u32 = smem_u32(sP), ds_u32 = smem_u32(sDS);
    const int nchunks = (NK + 31) / 32;
    const int ncols_slab = slabs * 64;
    uint32_t ph_mma = 0;
    for (int t = 0; t < q_tiles; ++t) {
      const int grow = t * 128 + row;
      const bool valid = grow < n;
      // delta = rowsum(dO * O), LSE
      float delta = 0.f, l2 = 0.f;
      if (valid) {
        const bf16* po = o + ((size_t)b * n + grow) * inner + h * kDh;
        const bf16* pd = dout + ((size_t)b * n + grow) * inner + h * kDh;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint4 a = *reinterpret_cast<const uint4*>(po + g * 8);
          const uint4 d = *reinterpret_cast<const uint4*>(pd + g * 8);
          const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
          const float2 d0 = unpack_bf16x2(d.x), d1 = unpack_bf16x2(d.y), d2 = unpack_bf16x2(d.z), d3 = unpack_bf16x2(d.w);
          delta += a0.x * d0.x + a0.y * d0.y + a1.x * d1.x + a1.y * d1.y + a2.x * d2.x + a2.y * d2.y +
                   a3.x * d3.x + a3.y * d3.y;
        }
        l2 = lse[((size_t)b * heads + h) * n + grow] * kLog2e;
      }
      // ---- P_t
      mbar_wait(&bars->mma, ph_mma); ph_mma ^= 1;
      tc_fence_after_sync();
      for (int c = 0; c < nchunks; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = c * 32 + g * 8;
          if (col < ncols_slab) {
            float pv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              pv[j] = (valid && col + j < n) ? exp2f(__uint_as_float(v[g * 8 + j]) * sl2 - l2) : 0.f;
            uint4 u;
            u.x = pack_bf16x2(pv[0], pv[1]); u.y = pack_bf16x2(pv[2], pv[3]);
            u.z = pack_bf16x2(pv[4], pv[5]); u.w = pack_bf16x2(pv[6], pv[7]);
            st_swz_chunk(p_u32 + (col >> 6) * 16384, row, (col & 63) >> 3, u);
          }
        }
      }
      // zero the slab tail beyond the last 32-column chunk (keys in [nchunks*32, slabs*64))
      for (int col = nchunks * 32; col < ncols_slab; col += 8) {
        st_swz_chunk(p_u32 + (col >> 6) * 16384, row, (col & 63) >> 3, make_uint4(0, 0, 0, 0));
        st_swz_chunk(ds_u32 + (col >> 6) * 16384, row, (col & 63) >> 3, make_uint4(0, 0, 0, 0));
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(&bars->warps);
      // ---- dS_t
      mbar_wait(&bars->mma, ph_mma); ph_mma ^= 1;
      tc_fence_after_sync();
      for (int c = 0; c < nchunks; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = c * 32 + g * 8;
          if (col < ncols_slab) {
            const uint4 pk = ld_swz_chunk(p_u32 + (col >> 6) * 16384, row, (col & 63) >> 3);
            const float2 p0 = unpack_bf16x2(pk.x), p1 = unpack_bf16x2(pk.y), p2 = unpack_bf16x2(pk.z),
                         p3 = unpack_bf16x2(pk.w);
            const float pp[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
            float ds[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              ds[j] = (valid && col + j < n) ? pp[j] * (__uint_as_float(v[g * 8 + j]) - delta) * scale : 0.f;
            uint4 u;
            u.x = pack_bf16x2(ds[0], ds[1]); u.y = pack_bf16x2(ds[2], ds[3]);
            u.z = pack_bf16x2(ds[4], ds[5]); u.w = pack_bf16x2(ds[6], ds[7]);
            st_swz_chunk(ds_u32 + (col >> 6) * 16384, row, (col & 63) >> 3, u);
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(&bars->warps);
      // ---- dQ_t epilogue
      mbar_wait(&bars->mma, ph_mma); ph_mma ^= 1;
      tc_fence_after_sync();
      {
        uint32_t a0[32], a1[32];
        tmem_ld_32x32(t_row, a0);
        tmem_ld_32x32(t_row + 32, a1);
        tmem_ld_wait();
        if (valid) {
          bf16* dst = dqkv + ((size_t)b * n + grow) * (3 * inner) + h * kDh;
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
      }
      if (t + 1 < q_tiles) {
        tc_fence_before_sync();
        mbar_arrive(&bars->warps);   // Q / dO / P / dS / scratch may be reused
      }
    }
    // ---- dK / dV epilogue (the last commit covered every MMA)
    for (int kt = 0; kt < key_tiles; ++kt) {
      const int key = kt * 128 + row;
#pragma unroll
      for (int which = 0; which < 2; ++which) {     // 0: dV, 1: dK
        uint32_t a0[32], a1[32];
        const uint32_t col = (which == 0 ? kColDV : kColDK) + kt * 64;
        tmem_ld_32x32(t_row + col, a0);
        tmem_ld_32x32(t_row + col + 32, a1);
        tmem_ld_wait();
        if (key < n) {
          bf16* dst = dqkv + ((size_t)b * n + key) * (3 * inner) + (which == 0 ? 2 : 1) * inner + h * kDh;
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
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
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
  const int NK = (n + 15) & ~15;
  CUtensorMap map_q, map_kv;
  s = make_tmap_3d_bf16(&map_q, qkv_bf16, 3 * inner, n, batch, 3 * inner, (uint64_t)n * 3 * inner, 128);
  if (s) return s;
  s = make_tmap_3d_bf16(&map_kv, qkv_bf16, 3 * inner, n, batch, 3 * inner, (uint64_t)n * 3 * inner, NK);
  if (s) return s;
  const int kv_region = (NK * 128 + 1023) & ~1023;
  const int p_slabs = (NK + 63) / 64;
  const int regA = std::max(16384 + kv_region, p_slabs * 16384);
  const int smem = 1024 + regA + kv_region + 128;
  static int configured = 0;
  if (configured < smem) {
    M3L_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  const int q_tiles = (n + 127) / 128;
  attn_fwd_kernel<<<batch * heads * q_tiles, 160, smem, (cudaStream_t)stream>>>(
      map_q, map_kv, (bf16*)out_bf16, lse, n, heads, inner, q_tiles, scale);
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_attention_bwd(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16,
                                 const float* lse, int batch, int n, int heads, int dim_head, float scale,
                                 void* dqkv_bf16, void* stream) {
  M3L_REQUIRE(qkv_bf16 && out_bf16 && dout_bf16 && lse && dqkv_bf16, "attention_bwd: null pointer");
  int s = attn_check(n, heads, dim_head, batch);
  if (s) return s;
  if (batch == 0) return M3L_OK;
  const int inner = heads * kDh;
  const int NK = (n + 15) & ~15;
  CUtensorMap map_q, map_kv, map_do;
  s = make_tmap_3d_bf16(&map_q, qkv_bf16, 3 * inner, n, batch, 3 * inner, (uint64_t)n * 3 * inner, 128);
  if (s) return s;
  s = make_tmap_3d_bf16(&map_kv, qkv_bf16, 3 * inner, n, batch, 3 * inner, (uint64_t)n * 3 * inner, NK);
  if (s) return s;
  s = make_tmap_3d_bf16(&map_do, dout_bf16, inner, n, batch, inner, (uint64_t)n * inner, 128);
  if (s) return s;
  const int kv_region = (NK * 128 + 1023) & ~1023;
  const int slabs = 2 * ((NK + 127) / 128);
  const int smem = 1024 + 2 * 16384 + 2 * kv_region + 2 * slabs * 16384 + 128;
  static bool configured = false;
  if (!configured) {
    M3L_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  attn_bwd_kernel<<<batch * heads, 160, smem, (cudaStream_t)stream>>>(
      map_q, map_kv, map_do, (const bf16*)out_bf16, (const bf16*)dout_bf16, lse, (bf16*)dqkv_bf16, n, heads,
      inner, scale);
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}
