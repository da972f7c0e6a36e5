// m3l_b200 — fused pre-norm feed-forward block, forward (sm_100a: TMA / tcgen05 / TMEM):
//
//   out[M,256] = x + W2 · GELU(W1 · LayerNorm(x) + b1) + b2
//
// i.e. one whole `x = FeedForward(x) + x` step of vit_pytorch's Transformer (FeedForward.net =
// LayerNorm -> Linear -> GELU -> Linear), as the reference runs it for every encoder /
// decoder layer (/root/reference/models/pretrain_models.py:113,784 -> vit-pytorch 1.6.4).  The
// [M, hidden] activation never touches HBM in inference; in training the kernel can additionally
// write what the (unfused) backward consumes: LN statistics, LN(x), GELU(pre) and GELU'(pre).
//
// One persistent CTA per SM walks 128-row tiles.  Per tile:
//   TMA            x tile [128 x 256] bf16 -> smem, 4 K-major 128B-swizzled k-blocks (the A operand layout)
//   row warps (4)  LayerNorm IN PLACE on the smem tile (two rows in flight per warp), stats -> HBM
//   per 128-wide hidden chunk c (weights stream from L2 through a ring of 16 KB slots):
//     MMA          acc1[c&1] (TMEM, 128 cols) = LN(x) · W1[c]^T            16 x tcgen05.mma 128x128x16, SS
//     GELU warps   (16: 4 per TMEM lane quadrant x 4 column parts) acc1 -> +b1 -> GELU -> bf16, packed two per
//                  32-bit TMEM column and written back OVER the accumulator (tcgen05.st): the hidden chunk becomes
//                  the A operand of the second product without passing through shared memory
//     MMA          acc2 (TMEM, 256 cols) += h[c] · W2[:, c]^T             16 x tcgen05.mma 128x128x16, TS (A in TMEM)
//   row warps      acc2 + b2 + x (re-fetched by TMA from L2) -> bf16 -> TMA store
// TMEM: 2 x 128 (acc1 / h, double-buffered so GELU(c) overlaps the products of c+1) + 256 (acc2) = 512 columns.
// The tensor pipe executes tcgen05.mma in issue order, so "h[c] consumed" needs no barrier of its own: the
// product that overwrites acc1[c&1] (chunk c+2) is issued after the product that reads h[c].
//
// Budget per tile (hidden 1024): 2 x 8192 clk of tensor work, 1 MB of weights through L2 (the binding
// resource when all 148 CTAs stream at once: ~42 B/clk/SM), 128 x 1024 GELUs with ONE MUFU each.
#include <stdlib.h>

#include "common.cuh"
#include "m3l_internal.h"

namespace m3l {

namespace {

constexpr int kBM = 128, kD = 256, kCH = 128;
constexpr int kRowWarps = 4, kGeluWarps = 16;
constexpr int kWarps = 2 + kRowWarps + kGeluWarps;
constexpr int kThreads = 32 * kWarps;              // 704
constexpr int kKbBytes = 16384;                    // one [128 rows x 64] bf16 k-block
constexpr int kSlotBytes = 32768;                  // ring slot: two W1 k-blocks, or one [256 x 64] W2 k-block
constexpr int kABytes = 4 * kKbBytes;
constexpr int kStgUnit = 32 * 64;                  // staging unit: 32 rows x 64 B (32 bf16 columns, SWIZZLE_64B)
constexpr int kMaxHidden = 1024;
constexpr int kMaxSlots = 4;

template <bool SAVE>
struct Cfg {
  static constexpr int kSlots = SAVE ? 3 : 4;
  static constexpr int kGeluStg = SAVE ? kGeluWarps * 2048 : 0;     // one 32 x 32 bf16 unit per GELU warp
  static constexpr int kSmem = 1024 + kABytes + kSlots * kSlotBytes + kRowWarps * 2 * kStgUnit + kGeluStg +
                               kMaxHidden * 4 + 3 * kD * 4 + 512;
};

struct alignas(8) Bars {
  uint64_t full[kMaxSlots], empty[kMaxSlots];
  uint64_t a_loaded, a_ready, a_free, a_stored;
  uint64_t acc1_full[2], h_full[2], h_read[2];
  uint64_t acc2_full, acc2_empty;
  uint64_t side_full[kRowWarps][2];
  uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 512, "barrier block");

// Cycle-counter probes of CTA 0 (tools/rb_timeline.py), compiled in only with -DM3L_RB_PROFILE.
#ifdef M3L_RB_PROFILE
__device__ long long g_rb_prof[4 * 512];
#define RB_EV(role, idx) do { if (blockIdx.x == 0 && (idx) < 512) g_rb_prof[(role) * 512 + (idx)] = clock64(); } while (0)
#else
#define RB_EV(role, idx) do { } while (0)
#endif

struct Args {
  int M, hidden;
  const float* gamma;
  const float* beta;
  float eps;
  const float* b1;
  const float* b2;
  float* stats;          // [M, 2] or null
  int want_xn;           // store LN(x)
  int out_has_x;         // out already holds x (in place, or pre-copied): the block output is ADDED to it
  bf16* out;
};

// ---- TS-mode product (A operand in TMEM: lane = row, bf16 pairs packed per 32-bit column) and TMEM stores
M3L_DEVINL void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
M3L_DEVINL void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0],"
      " {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
M3L_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- exact-erf GELU with ONE MUFU and 8 packed FMA-pipe instructions per PAIR of elements ------------------
//   GELU(x) = x Phi(x) = max(x, 0) - |x| * (erfc(|x| / sqrt 2) / 2)
// erfc(z) / 2 = exp2(P(z)): P is the degree-6 weighted-minimax fit of log2(erfc(z) / 2) on [0, 4.3]
// (|GELU error| <= 6e-7 in fp32 against scipy, relative accuracy kept in the negative tail; the Abramowitz-Stegun
// form used by the unfused epilogues needs a reciprocal as well).  The polynomial is evaluated in the variable
// na = max(-|x|, -6.08) (1/sqrt 2 and the sign folded into the coefficients); beyond 6.08 Phi is 0 / 1 in fp32.
// The first version (Phi = 1/2 + copysign(1/2, x)(1 - erfc)) spent 13 packed FMA-pipe instructions per pair and the
// 16 GELU warps needed 2400 clk per 128-column chunk against 2048 clk of tensor work.
constexpr float kE1 = 1.1511168561f, kE2 = -4.5908273735e-01f, kE3 = 5.2926736750e-02f, kE4 = 7.7240424434e-03f,
                kE5 = 6.4775743001e-04f, kE6 = 1.7755137836e-05f;
constexpr float kClamp = -6.08f;

// GELU(x) of two pre-activations (packed fp32 arithmetic); if WANT_GRAD also GELU'(x) = Phi(x) + x phi(x)
template <bool WANT_GRAD>
M3L_DEVINL void gelu2(uint32_t x0u, uint32_t x1u, f32x2& gelu, f32x2& dgelu) {
  const float x0 = __uint_as_float(x0u), x1 = __uint_as_float(x1u);
  const f32x2 na = f2_pack(fmaxf(-fabsf(x0), kClamp), fmaxf(-fabsf(x1), kClamp));
  const f32x2 relu = f2_pack(fmaxf(x0, 0.f), fmaxf(x1, 0.f));
  f32x2 t = f2_fma(f2_splat(kE6), na, f2_splat(kE5));
  t = f2_fma(t, na, f2_splat(kE4));
  t = f2_fma(t, na, f2_splat(kE3));
  t = f2_fma(t, na, f2_splat(kE2));
  t = f2_fma(t, na, f2_splat(kE1));
  const f32x2 pw = f2_fma(t, na, f2_splat(-1.0f));
  float p0, p1;
  f2_unpack(pw, p0, p1);
  const f32x2 eh = f2_pack(exp2f(p0), exp2f(p1));                    // erfc(|x| / sqrt 2) / 2 = Phi(-|x|), MUFU.EX2 x 2
  gelu = f2_fma(na, eh, relu);
  if (WANT_GRAD) {
    // GELU'(x) = Phi(x) + x phi(x) is "odd about 1/2": GELU'(-a) = 1 - GELU'(a).  With q = GELU'(-|x|) = Phi(-|x|) -
    // |x| phi(x) = eh + na * phi (everything already in the -|x| variable, na^2 = x^2 up to the clamp, where phi = 0):
    //   GELU'(x) = 1/2 + sign(x) * (1/2 - q)
    // 3 packed FMA-pipe instructions + 2 sign-bit LOP3 after the second exponential, instead of building Phi(x) and
    // x phi(x) separately (the GELU warps are instruction-issue bound in the training form: r02 ncu, issue 49 %)
    const f32x2 xx = f2_mul(f2_mul(na, na), f2_splat(-0.5f * 1.4426950408889634f));
    float q0, q1;
    f2_unpack(xx, q0, q1);
    const f32x2 pdf = f2_pack(exp2f(q0), exp2f(q1));                 // exp(-x^2 / 2)
    const f32x2 q = f2_fma(f2_mul(na, f2_splat(0.39894228040143268f)), pdf, eh);
    uint32_t h0, h1;
    f2_unpacku(f2_fma(q, f2_splat(-1.0f), f2_splat(0.5f)), h0, h1);              // 1/2 - q
    dgelu = f2_add(f2_splat(0.5f), f2_packu(h0 ^ (x0u & 0x80000000u), h1 ^ (x1u & 0x80000000u)));
  }
}

// ---- staging helpers (same layouts as the GEMM epilogues) ---------------------------------------------
// one 32 x 32 bf16 unit: registers (thread = row, 16 packed words) -> 64 B-swizzled staging -> TMA store
M3L_DEVINL void rb_store_unit(const CUtensorMap* map, uint32_t stg, int lane, const uint32_t (&w)[16], int col, int row) {
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
M3L_DEVINL uint4 rb_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
M3L_DEVINL float4 rb_lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
M3L_DEVINL void rb_sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// The per-tile sequence of products, walked identically by the TMA producer (weight slots) and the MMA issuer:
//   G1(0) G1(1)  then for c = 0 .. NC-1:  G2(c)  [G1(c+2)]          (2 NC items)
// G1(c): 2 slots, each two [128 x 64] k-blocks of W1[c*128 .. +128, :];   G2(c): 2 slots, each one [256 x 64] k-block
// of W2[:, c*128 .. +128] (the second product runs N = 256 wide: half as many tcgen05.mma for the one issuing thread,
// which at ~80 clk of issue work per instruction could not keep up with 64-clk N = 128 products).
M3L_DEVINL void tile_item(int s, int nc, bool* is_g2, int* c) {
  if (s == 2 * nc - 1) { *is_g2 = true; *c = nc - 1; }
  else if (s < 2) { *is_g2 = false; *c = s; }
  else if ((s & 1) == 0) { *is_g2 = true; *c = (s >> 1) - 1; }
  else { *is_g2 = false; *c = (s + 1) >> 1; }
}

template <bool SAVE>
__global__ void __launch_bounds__(kThreads, 1)
ln_mlp_fwd_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                  const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_x32,
                  const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_xn,
                  const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_gp, const Args p) {
  using C = Cfg<SAVE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_tile = smem;                                         // 4 x [128 x 64] bf16, swizzled
  uint8_t* ring = a_tile + kABytes;
  uint8_t* row_stg = ring + C::kSlots * kSlotBytes;               // [row warp][2][2 KB]
  uint8_t* gelu_stg = row_stg + kRowWarps * 2 * kStgUnit;         // SAVE: [gelu warp][2 KB]
  float* s_b1 = reinterpret_cast<float*>(gelu_stg + C::kGeluStg);
  float* s_gamma = s_b1 + kMaxHidden;
  float* s_beta = s_gamma + kD;
  float* s_b2 = s_beta + kD;
  Bars* bars = reinterpret_cast<Bars*>(s_b2 + kD);
  // bias / LayerNorm parameters are READ through 32-bit shared-state-space addresses (ld.shared): through the float*
  // above the compiler emits generic LD.E, which goes down the global-memory path of the LSU ("lg" throttle stalls)
  const uint32_t s_b1_u32 = smem_u32(s_b1), s_gamma_u32 = smem_u32(s_gamma), s_beta_u32 = smem_u32(s_beta),
                 s_b2_u32 = smem_u32(s_b2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (p.M + kBM - 1) / kBM;
  const int n_tiles = (int)blockIdx.x < tiles_m ? (tiles_m - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int nc = p.hidden / kCH;
  // Every CTA walks the hidden chunks in its own rotation: all CTAs start together and stream the SAME weights, and
  // with one common order 148 SMs ask the same L2 lines at the same moment (measured: weight tiles took > 2000 clk to
  // arrive while the L2 slices were 11 % busy).  The sum over chunks does not care about the order.
  const int rot = (int)blockIdx.x % nc;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w1);
    tma_prefetch_desc(&map_w2);
    tma_prefetch_desc(&map_x32);
    tma_prefetch_desc(&map_out);
    for (int s = 0; s < C::kSlots; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->a_loaded, 1);
    mbar_init(&bars->a_ready, kGeluWarps);
    mbar_init(&bars->a_free, 1);
    mbar_init(&bars->a_stored, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->acc1_full[b], 1);
      mbar_init(&bars->h_full[b], kGeluWarps);
      mbar_init(&bars->h_read[b], kRowWarps);
    }
    mbar_init(&bars->acc2_full, 1);
    mbar_init(&bars->acc2_empty, kRowWarps);
    for (int w = 0; w < kRowWarps; ++w) {
      mbar_init(&bars->side_full[w][0], 1);
      mbar_init(&bars->side_full[w][1], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_wait();      // prologue above overlapped the predecessor kernel; global memory from here on
  for (int i = threadIdx.x; i < p.hidden; i += kThreads) s_b1[i] = p.b1[i];
  if (threadIdx.x < kD) {
    s_gamma[threadIdx.x] = p.gamma[threadIdx.x];
    s_beta[threadIdx.x] = p.beta[threadIdx.x];
    s_b2[threadIdx.x] = p.b2[threadIdx.x];
  }
  __syncthreads();
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------- TMA producer ---------------------------------------
    if (elect_one()) {
      int slot = 0;
      uint32_t phase = 0;
#pragma unroll 1
      for (int i = 0; i < n_tiles; ++i) {
        const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * kBM;
        if (i > 0) {
          mbar_wait(&bars->a_free, (i - 1) & 1);              // every product reading the previous tile is complete
          if (p.want_xn) mbar_wait(&bars->a_stored, (i - 1) & 1);   // ... and its LN(x) copy has left shared memory
        }
        mbar_arrive_expect_tx(&bars->a_loaded, kABytes);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(a_tile + kb * kKbBytes, &map_x, &bars->a_loaded, kb * 64, m0);
#pragma unroll 1
        for (int it = 0; it < 2 * nc; ++it) {
          bool is_g2;
          int c;
          tile_item(it, nc, &is_g2, &c);
          const int pc = (c + rot) % nc;      // physical hidden chunk
#pragma unroll 1
          for (int q = 0; q < 2; ++q) {
            mbar_wait(&bars->empty[slot], phase ^ 1);
            mbar_arrive_expect_tx(&bars->full[slot], kSlotBytes);
            uint8_t* dst = ring + slot * kSlotBytes;
            if (is_g2) {            // k-block q of the chunk, all 256 output rows of W2
              tma_load_2d(dst, &map_w2, &bars->full[slot], pc * kCH + q * 64, 0);
            } else {                // k-blocks 2q, 2q+1 of W1 rows [pc*128, +128)
              tma_load_2d(dst, &map_w1, &bars->full[slot], (2 * q) * 64, pc * kCH);
              tma_load_2d(dst + kKbBytes, &map_w1, &bars->full[slot], (2 * q + 1) * 64, pc * kCH);
            }
            if (++slot == C::kSlots) { slot = 0; phase ^= 1; }
          }
          RB_EV(0, i * 2 * nc + it);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -----------------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(kBM, 128, 0, 0);     // first product: N = 128 (one hidden chunk)
      constexpr uint32_t idesc2 = umma_idesc_bf16(kBM, 256, 0, 0);     // second product: N = 256 (all outputs)
      int slot = 0;
      uint32_t phase = 0;
      int gc = 0;                                     // chunks issued before this tile
      // descriptors: only the 14-bit start-address field changes, so each one is a 32-bit add on the low word
      const uint64_t a_desc0 = umma_smem_desc(smem_u32(a_tile), 16, 1024);
      const uint64_t r_desc0 = umma_smem_desc(smem_u32(ring), 16, 1024);
      const uint32_t d_hi = (uint32_t)(a_desc0 >> 32), a_lo0 = (uint32_t)a_desc0, r_lo0 = (uint32_t)r_desc0;
      auto mk = [&](uint32_t lo) {
        uint64_t d;
        asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(d_hi));
        return d;
      };
#pragma unroll 1
      for (int i = 0; i < n_tiles; ++i) {
        mbar_wait(&bars->a_ready, i & 1);
        tc_fence_after_sync();
#pragma unroll 1
        for (int it = 0; it < 2 * nc; ++it) {
          bool is_g2;
          int c;
          tile_item(it, nc, &is_g2, &c);
          const int k = gc + c, buf = k & 1;
          if (!is_g2) {
            // ---- acc1[buf] = LN(x) . W1[c]^T
            // (training: the row warps copy the hidden chunk that lived in this buffer, h(k - 2), out of tensor memory;
            // the product that consumed it was issued earlier, only their read has to be waited for)
            if (SAVE && k >= 2) {
              mbar_wait(&bars->h_read[buf], ((k - 2) >> 1) & 1);
              tc_fence_after_sync();
            }
            const uint32_t d = tmem_base + buf * 128;
            RB_EV(1, (i * 2 * nc + it) * 2);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              mbar_wait(&bars->full[slot], phase);
              tc_fence_after_sync();
              const uint32_t b_lo = r_lo0 + slot * (kSlotBytes >> 4);
#pragma unroll
              for (int j = 0; j < 8; ++j) {          // k-block 2q + (j >> 2), k-step j & 3
                const uint32_t a_off = ((2 * q + (j >> 2)) * kKbBytes + (j & 3) * 32) >> 4;
                const uint32_t b_off = ((j >> 2) * kKbBytes + (j & 3) * 32) >> 4;
                umma_bf16(d, mk(a_lo0 + a_off), mk(b_lo + b_off), idesc1, (q > 0 || j > 0) ? 1u : 0u);
              }
              umma_commit(&bars->empty[slot]);
              if (++slot == C::kSlots) { slot = 0; phase ^= 1; }
            }
            umma_commit(&bars->acc1_full[buf]);
            if (c == nc - 1) umma_commit(&bars->a_free);
            RB_EV(1, (i * 2 * nc + it) * 2 + 1);
          } else {
            // ---- acc2 += h[c] . W2[:, c]^T   (A = the bf16 hidden chunk the GELU warps left in acc1[buf])
            mbar_wait(&bars->h_full[buf], (k >> 1) & 1);
            if (c == 0 && i > 0) mbar_wait(&bars->acc2_empty, (i - 1) & 1);
            tc_fence_after_sync();
            RB_EV(1, (i * 2 * nc + it) * 2);
            const uint32_t d = tmem_base + 256;
            const uint32_t h_col = tmem_base + buf * 128;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              mbar_wait(&bars->full[slot], phase);
              tc_fence_after_sync();
              const uint32_t b_lo = r_lo0 + slot * (kSlotBytes >> 4);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int t = q * 4 + j;              // k-step of the chunk: k = 16 t .. 16 t + 15
                umma_bf16_ts(d, h_col + 32 * (t >> 1) + 8 * (t & 1), mk(b_lo + ((j * 32) >> 4)), idesc2,
                             (c > 0 || q > 0 || j > 0) ? 1u : 0u);
              }
              umma_commit(&bars->empty[slot]);
              if (++slot == C::kSlots) { slot = 0; phase ^= 1; }
            }
            if (c == nc - 1) umma_commit(&bars->acc2_full);
            RB_EV(1, (i * 2 * nc + it) * 2 + 1);
          }
        }
        gc += nc;
      }
    }
  } else if (warp < 2 + kRowWarps) {
    // ------------------------------- row warps: output epilogue --------------------------
    const int rw = warp - 2;
    const int quad = warp & 3;                       // TMEM lane quadrant (epilogue rows)
    const uint32_t stg = smem_u32(row_stg + rw * 2 * kStgUnit);
    int g_round = 0;                                  // global staging round counter of this warp
    // output rounds are 32 rows x 32 columns (64 B rows, SWIZZLE_64B): 32 accumulator registers live per thread
    auto issue_side = [&](int gi, int m0, int r) {    // elected lane: fetch the residual unit of round r
      const int b = gi & 1;
      mbar_arrive_expect_tx(&bars->side_full[rw][b], kStgUnit);
      tma_load_2d_u32(stg + b * kStgUnit, &map_x32, &bars->side_full[rw][b], r * 32, m0 + quad * 32);
    };
    int gc_r = 0;                                     // chunks before this tile (same counter as the MMA / GELU roles)
#pragma unroll 1
    for (int i = 0; i < n_tiles; ++i) {
      const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * kBM;
      const int row0 = m0 + quad * 32;
      const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + 256;
      if (SAVE) {
        // ---- training: these warps are idle during the chunk loop, so THEY write GELU(pre) to HBM.  The bf16 hidden
        // chunk already sits in tensor memory (packed, the A operand of the second product): read it back (4 x 16
        // columns = this row's 128 values), release the buffer, then stage + TMA-store two [32 x 64] tiles.  The GELU
        // warps are left with ONE store per chunk (GELU'): with both on them the kernel was store-wait bound
        // (102 us vs 67 us without the training outputs, r02 ncu).
#pragma unroll 1
        for (int c = 0; c < nc; ++c) {
          const int k = gc_r + c, buf = k & 1;
          const int pc = (c + rot) % nc;
          mbar_wait(&bars->h_full[buf], (k >> 1) & 1);
          tc_fence_after_sync();
          uint32_t hw[64];
          const uint32_t th = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + buf * 128;
#pragma unroll
          for (int part = 0; part < 4; ++part)
            tmem_ld_32x16(th + part * 32, *reinterpret_cast<uint32_t(*)[16]>(&hw[part * 16]));
          tmem_ld_wait();
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->h_read[buf]);
#pragma unroll
          for (int u = 0; u < 2; ++u) {               // hidden columns pc*128 + u*64 .. +64 of rows row0 .. +32
            tma_wait_group_read<0>();                 // (issuing lane) the previous tile has left the staging buffer
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {             // 128-byte rows, 16-byte chunks XOR-swizzled by (row & 7)
              const uint32_t addr = stg + lane * 128 + ((j ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(hw[u * 32 + 4 * j]),
                           "r"(hw[u * 32 + 4 * j + 1]), "r"(hw[u * 32 + 4 * j + 2]), "r"(hw[u * 32 + 4 * j + 3])
                           : "memory");
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (elect_one()) {
              tma_store_2d(&map_h, stg, pc * kCH + u * 64, row0);
              tma_commit_group();
            }
          }
        }
        gc_r += nc;
      }
      if (p.out_has_x) {
        // ---- out += acc2 + b2 with bf16 vector reductions: no residual fetch on the path that frees acc2
        mbar_wait(&bars->acc2_full, i & 1);
        tc_fence_after_sync();
        if (rw == 0 && lane == 0) RB_EV(2, i * 4 + 2);
#pragma unroll 1
        for (int r = 0; r < 8; ++r, ++g_round) {
          uint32_t v[32];
          tmem_ld_32x32(t_acc + r * 32, v);
          tmem_ld_wait();
          if (r == 7) {
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->acc2_empty);
          }
          uint32_t w[16];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = rb_lds_f4(s_b2_u32 + (r * 32 + j) * 4);
            w[j >> 1] = pack_bf16x2(__uint_as_float(v[j]) + b.x, __uint_as_float(v[j + 1]) + b.y);
            w[(j >> 1) + 1] = pack_bf16x2(__uint_as_float(v[j + 2]) + b.z, __uint_as_float(v[j + 3]) + b.w);
          }
          // 16-byte bf16 reductions from registers (REDG.ADD.BF16x8): thread = row, 64 contiguous bytes per round.
          // (Staging through shared memory + TMA reduce-add cost ~1000 clk per round here: the bulk group's
          // "read done" comes back slowly and only two 2 KB buffers fit beside the weight ring.)
          if (row0 + lane < p.M) {
            bf16* dst = p.out + (size_t)(row0 + lane) * kD + r * 32;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst + 8 * j), "r"(w[4 * j]),
                           "r"(w[4 * j + 1]), "r"(w[4 * j + 2]), "r"(w[4 * j + 3])
                           : "memory");
          }
        }
        if (rw == 0 && lane == 0) RB_EV(2, i * 4 + 3);
        continue;
      }
      // ---- out = acc2 + b2 + x with x re-fetched by TMA (out does not hold x)
      // buffer (g_round & 1) was last used two rounds ago and its store has been waited for below
      if (elect_one()) issue_side(g_round, m0, 0);
      __syncwarp();
      mbar_wait(&bars->acc2_full, i & 1);
      tc_fence_after_sync();
      if (rw == 0 && lane == 0) RB_EV(2, i * 4 + 2);
#pragma unroll 1
      for (int r = 0; r < 8; ++r, ++g_round) {
        uint32_t v[32];
        tmem_ld_32x32(t_acc + r * 32, v);
        if (r + 1 < 8) {
          __syncwarp();       // every lane is done with the other buffer (its row reads of round r-1)
          if (elect_one()) {
            tma_wait_group_read<0>();     // ... and so is the TMA store issued from it
            fence_proxy_async_smem();
            issue_side(g_round + 1, m0, r + 1);
          }
        }
        tmem_ld_wait();
        if (r == 7) {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->acc2_empty);
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b = rb_lds_f4(s_b2_u32 + (r * 32 + j) * 4);
          v[j] = __float_as_uint(__uint_as_float(v[j]) + b.x);
          v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) + b.y);
          v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) + b.z);
          v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) + b.w);
        }
        const uint32_t sb = stg + (g_round & 1) * kStgUnit;
        mbar_wait(&bars->side_full[rw][g_round & 1], (g_round >> 1) & 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {      // in place: every thread only touches its own 64-byte row
          const uint32_t addr = sb + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
          uint4 u = rb_lds128(addr);
          const float2 s0 = unpack_bf16x2(u.x), s1 = unpack_bf16x2(u.y), s2 = unpack_bf16x2(u.z), s3 = unpack_bf16x2(u.w);
          u.x = pack_bf16x2(__uint_as_float(v[8 * j]) + s0.x, __uint_as_float(v[8 * j + 1]) + s0.y);
          u.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]) + s1.x, __uint_as_float(v[8 * j + 3]) + s1.y);
          u.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]) + s2.x, __uint_as_float(v[8 * j + 5]) + s2.y);
          u.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]) + s3.x, __uint_as_float(v[8 * j + 7]) + s3.y);
          rb_sts128(addr, u);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          tma_store_2d(&map_out, sb, r * 32, row0);
          tma_commit_group();
        }
      }
      if (rw == 0 && lane == 0) RB_EV(2, i * 4 + 3);
      // the next tile's first side fetch reuses buffer (g_round & 1), last stored from two rounds ago
      __syncwarp();
      if (elect_one()) {
        tma_wait_group_read<0>();
        fence_proxy_async_smem();
      }
      __syncwarp();
    }
    tma_wait_group_read<0>();
  } else {
    // ------------------------------- GELU warps (they also run the LayerNorm prologue) ----
    const int gw = warp - 2 - kRowWarps;
    const int quad = warp & 3;
    const int part = gw >> 2;                         // 32-column part of the 128-column chunk
    const uint32_t stg = smem_u32(gelu_stg + gw * 2048);
    const uint32_t a_u32 = smem_u32(a_tile);
    // LayerNorm in place on the smem tile: warp gw owns rows gw*8 .. +8; a lane holds ONE 16-byte chunk of a row
    // (k-block lane >> 3, chunk lane & 7), so gamma / beta are 16 registers re-read once per tile, and four rows are
    // in flight per iteration (one pass: sum and sum of squares, 5 butterfly stages for the eight partials).
    // History: on the 4 row warps, one row at a time with 5 + 5 dependent shuffles, this took 10.5 k clk per tile; with
    // 8 lanes per row (3-stage butterflies) 7 k clk, because every lane then re-read 64 gamma / beta floats per
    // iteration: 4 k shared-memory wavefronts per tile, the LSU pipe was the limit.
    auto layer_norm_tile = [&](int i) {
      const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * kBM;
      const int kb = lane >> 3, jc = lane & 7;
      f32x2 gg[4], bb[4];
      {
        const float4 g0 = rb_lds_f4(s_gamma_u32 + (kb * 64 + jc * 8) * 4);
        const float4 g1 = rb_lds_f4(s_gamma_u32 + (kb * 64 + jc * 8 + 4) * 4);
        const float4 b0 = rb_lds_f4(s_beta_u32 + (kb * 64 + jc * 8) * 4);
        const float4 b1 = rb_lds_f4(s_beta_u32 + (kb * 64 + jc * 8 + 4) * 4);
        gg[0] = f2_pack(g0.x, g0.y); gg[1] = f2_pack(g0.z, g0.w); gg[2] = f2_pack(g1.x, g1.y); gg[3] = f2_pack(g1.z, g1.w);
        bb[0] = f2_pack(b0.x, b0.y); bb[1] = f2_pack(b0.z, b0.w); bb[2] = f2_pack(b1.x, b1.y); bb[3] = f2_pack(b1.z, b1.w);
      }
      mbar_wait(&bars->a_loaded, i & 1);
      if (gw == 0 && lane == 0) RB_EV(2, i * 4);
#pragma unroll 1
      for (int it = 0; it < 2; ++it) {
        const int r0 = gw * 8 + it * 4;
        uint32_t addr[4];
        f32x2 v[4][4];
        float sum[4], sq[4];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int row = r0 + rr;
          addr[rr] = a_u32 + kb * kKbBytes + row * 128 + ((jc ^ (row & 7)) << 4);
          const uint4 u = rb_lds128(addr[rr]);
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
          f32x2 s2 = f2_splat(0.f), q2 = f2_splat(0.f);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[rr][e] = f2_packu(w[e] << 16, w[e] & 0xffff0000u);      // bf16 pair -> fp32 pair (exact)
            s2 = f2_add(s2, v[rr][e]);
            q2 = f2_fma(v[rr][e], v[rr][e], q2);
          }
          float sa, sb, qa, qb;
          f2_unpack(s2, sa, sb);
          f2_unpack(q2, qa, qb);
          sum[rr] = sa + sb;
          sq[rr] = qa + qb;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            sum[rr] += __shfl_xor_sync(0xffffffffu, sum[rr], o);
            sq[rr] += __shfl_xor_sync(0xffffffffu, sq[rr], o);
          }
        }
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int row = r0 + rr;
          const float mean = sum[rr] * (1.0f / kD);
          // var = E[x^2] - mean^2 (bf16 inputs, fp32 sums over 256 values: the cancellation costs ~mean^2 / var ulps,
          // far below the bf16 rounding of the output)
          const float var = fmaxf(fmaf(-mean, mean, sq[rr] * (1.0f / kD)), 0.f);
          const float rstd = rsqrtf(var + p.eps);
          if (p.stats != nullptr && lane == 0 && m0 + row < p.M)
            *reinterpret_cast<float2*>(p.stats + 2 * (size_t)(m0 + row)) = make_float2(mean, rstd);
          const f32x2 r2 = f2_splat(rstd), nm2 = f2_splat(-mean * rstd);
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const f32x2 y = f2_fma(f2_fma(v[rr][e], r2, nm2), gg[e], bb[e]);
            float y0, y1;
            f2_unpack(y, y0, y1);
            o[e] = pack_bf16x2(y0, y1);
          }
          rb_sts128(addr[rr], make_uint4(o[0], o[1], o[2], o[3]));
        }
      }
      fence_proxy_async_smem();       // the normalised rows are read by tcgen05.mma / TMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->a_ready);
      if (gw == 0 && lane == 0) RB_EV(2, i * 4 + 1);
      if (p.want_xn) {
        // all 16 warps have normalised their rows -> one thread copies the whole tile out
        asm volatile("bar.sync 1, %0;" ::"n"(kGeluWarps * 32) : "memory");
        if (gw == 0 && elect_one()) {
          for (int k4 = 0; k4 < 4; ++k4) tma_store_2d(&map_xn, a_u32 + k4 * kKbBytes, k4 * 64, m0);
          tma_commit_group();
          tma_wait_group_read<0>();
          mbar_arrive(&bars->a_stored);
        }
        __syncwarp();
      }
    };
    int gc = 0;
    if (n_tiles > 0) layer_norm_tile(0);
    for (int i = 0; i < n_tiles; ++i) {
      const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * kBM;
      const int row0 = m0 + quad * 32;
#pragma unroll 1
      for (int c = 0; c < nc; ++c) {
        const int k = gc + c, buf = k & 1;
        const int pc = (c + rot) % nc;               // physical hidden chunk (bias, saved columns)
        mbar_wait(&bars->acc1_full[buf], (k >> 1) & 1);
        tc_fence_after_sync();
        if (gw == 0 && lane == 0) RB_EV(3, k * 4);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + buf * 128 + part * 32;
        uint32_t v[32];
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        if (gw == 0 && lane == 0) RB_EV(3, k * 4 + 1);
        const uint32_t bias = s_b1_u32 + (pc * kCH + part * 32) * 4;
        uint32_t hp[16], gp[16];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float4 b4 = rb_lds_f4(bias + 8 * j);      // all lanes read the same address: one broadcast wavefront
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
          const int j2 = j + jj;
          const f32x2 x = f2_add(f2_packu(v[2 * j2], v[2 * j2 + 1]), jj == 0 ? f2_pack(b4.x, b4.y) : f2_pack(b4.z, b4.w));
          uint32_t x0, x1;
          f2_unpacku(x, x0, x1);
          f32x2 hh, gg;
          gelu2<SAVE>(x0, x1, hh, gg);
          float a0, a1;
          f2_unpack(hh, a0, a1);
          hp[j2] = pack_bf16x2(a0, a1);
          if (SAVE) {
            f2_unpack(gg, a0, a1);
            gp[j2] = pack_bf16x2(a0, a1);
          }
          }
        }
        // hidden chunk -> TMEM, over the accumulator columns this warp has just read: column part*32 + j holds
        // k = part*32 + 2j (low half) and 2j + 1 (high half) of row (quad*32 + lane)
        if (gw == 0 && lane == 0) RB_EV(3, k * 4 + 2);
        tmem_st_32x16(taddr, hp);
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->h_full[buf]);
        if (gw == 0 && lane == 0) RB_EV(3, k * 4 + 3);
        if (SAVE) rb_store_unit(&map_gp, stg, lane, gp, pc * kCH + part * 32, row0);      // GELU(pre): row warps
      }
      gc += nc;
      if (i + 1 < n_tiles) layer_norm_tile(i + 1);
    }
    if (SAVE || p.want_xn) tma_wait_group_read<0>();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

template <bool SAVE>
int launch(const m3l_ln_mlp_args* a, cudaStream_t stream) {
  using C = Cfg<SAVE>;
  static_assert(C::kSmem <= 227 * 1024, "shared memory budget");
  CUtensorMap map_x, map_w1, map_w2, map_x32, map_out, map_xn, map_h, map_gp;
  int s;
  if ((s = make_tmap_2d_bf16(&map_x, a->x, a->rows, kD, kD, 128))) return s;
  if ((s = make_tmap_2d_bf16(&map_w1, a->w1, a->hidden, kD, kD, 128))) return s;
  if ((s = make_tmap_2d_bf16(&map_w2, a->w2, kD, a->hidden, a->hidden, 256))) return s;
  if ((s = make_tmap_2d_bf16_sw64(&map_x32, a->x, a->rows, kD, kD, 32))) return s;
  if ((s = make_tmap_2d_bf16_sw64(&map_out, a->out, a->rows, kD, kD, 32))) return s;
  map_xn = map_out;
  map_h = map_out;
  map_gp = map_out;
  if (a->xn_out != nullptr && (s = make_tmap_2d_bf16(&map_xn, a->xn_out, a->rows, kD, kD, 128))) return s;
  if (SAVE) {
    if ((s = make_tmap_2d_bf16(&map_h, a->h_out, a->rows, a->hidden, a->hidden, 32))) return s;
    if ((s = make_tmap_2d_bf16_sw64(&map_gp, a->gp_out, a->rows, a->hidden, a->hidden, 32))) return s;
  }
  Args p;
  p.M = a->rows; p.hidden = a->hidden;
  p.gamma = a->gamma; p.beta = a->beta; p.eps = a->eps;
  p.b1 = a->b1; p.b2 = a->b2;
  p.stats = a->stats;
  p.want_xn = a->xn_out != nullptr ? 1 : 0;
  p.out = (bf16*)a->out;
  p.out_has_x = (a->out == a->x || a->out_has_x) ? 1 : 0;
  auto kern = ln_mlp_fwd_kernel<SAVE>;
  static bool configured = false;
  if (!configured) {
    M3L_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    configured = true;
  }
  const int tiles = (a->rows + kBM - 1) / kBM;
  const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
  M3L_CUDA(launch_kernel(kern, dim3(grid), dim3(kThreads), C::kSmem, stream, map_x, map_w1, map_w2, map_x32, map_out,
                         map_xn, map_h, map_gp, p));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

}  // namespace

}  // namespace m3l

// ---------------------------------------------------------------------------------------
// C-ABI (include/m3l_b200.h)
// ---------------------------------------------------------------------------------------
#ifdef M3L_RB_PROFILE
// measurement only: read and reset the in-kernel cycle counters of CTA 0
extern "C" int m3l_debug_rb_prof(long long* host_out, int n) {
  if (n > 4 * 512) return m3l::M3L_ERR_INVALID;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host_out, m3l::g_rb_prof, n * sizeof(long long));
  static long long zeros[4 * 512];
  cudaMemcpyToSymbol(m3l::g_rb_prof, zeros, sizeof(zeros));
  return m3l::M3L_OK;
}
#endif

extern "C" int m3l_ln_mlp_fwd(const m3l_ln_mlp_args* a, void* stream) {
  using namespace m3l;
  if (a == nullptr) return M3L_ERR_INVALID;
  M3L_REQUIRE(a->dim == kD, "ln_mlp_fwd: dim=%d unsupported (the fused block is built for dim 256)", a->dim);
  M3L_REQUIRE(a->rows > 0 && a->hidden >= kCH && a->hidden <= kMaxHidden && a->hidden % kCH == 0,
              "ln_mlp_fwd: bad shape rows=%d hidden=%d (hidden must be a multiple of 128 up to 1024)", a->rows, a->hidden);
  M3L_REQUIRE(a->x && a->gamma && a->beta && a->w1 && a->b1 && a->w2 && a->b2 && a->out, "ln_mlp_fwd: null pointer");
  M3L_REQUIRE((a->h_out == nullptr) == (a->gp_out == nullptr), "ln_mlp_fwd: h_out and gp_out go together");
  M3L_REQUIRE(a->h_out == nullptr || a->out == a->x || a->out_has_x,
              "ln_mlp_fwd: the training outputs need the accumulate-into-out mode (out == x or out_has_x)");
  if (a->h_out != nullptr) return launch<true>(a, (cudaStream_t)stream);
  return launch<false>(a, (cudaStream_t)stream);
}
