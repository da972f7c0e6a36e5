// m3l_b200 — common device helpers for sm_100a (B200).
//
// Thin inline-PTX wrappers around the Blackwell primitives the kernels use:
//   mbarrier (init / arrive / expect_tx / parity wait), TMA bulk-tensor loads
//   (cp.async.bulk.tensor), tcgen05 (TMEM alloc, mma, commit, ld) and the
//   shared-memory / instruction descriptors tcgen05.mma consumes.
// Nothing here is a port of the reference (the reference ships no native code:
// SURVEY.md §2.1); these are the building blocks of the from-scratch kernels.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <utility>

namespace m3l {

typedef __nv_bfloat16 bf16;

#define M3L_DEVINL __device__ __forceinline__

// ---------------------------------------------------------------------------------------
// status codes of the C-ABI (include/m3l_b200.h)
// ---------------------------------------------------------------------------------------
enum : int {
  M3L_OK = 0,
  M3L_ERR_INVALID = 1,   // bad argument (shape / alignment / null pointer)
  M3L_ERR_CUDA = 2,      // a CUDA runtime / driver call failed (see m3l_last_error)
  M3L_ERR_UNSUPPORTED = 3,
};

void set_last_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define M3L_CUDA(call)                                  \
  do {                                                  \
    int _s = ::m3l::check_cuda((call), #call);          \
    if (_s != ::m3l::M3L_OK) return _s;                 \
  } while (0)

#define M3L_REQUIRE(cond, ...)                          \
  do {                                                  \
    if (!(cond)) {                                      \
      ::m3l::set_last_error(__VA_ARGS__);               \
      return ::m3l::M3L_ERR_INVALID;                    \
    }                                                   \
  } while (0)

// ---------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel of the step is launched with the
// programmatic-stream-serialisation attribute (launch_kernel below) and starts with
//     <prologue that touches no global memory: barrier init, TMEM alloc, descriptor prefetch>
//     pdl_wait();      // predecessor grid complete, its writes visible
//     pdl_trigger();   // successor grid may start ITS prologue as SM resources free up
// so launch latency and prologues of consecutive kernels overlap the predecessor's tail, also
// inside a captured CUDA graph (programmatic edges).  Every kernel launched this way must
// execute pdl_wait() before its first global-memory access (reads AND writes): completion of
// grid N then implies completion of grid N-1, which keeps the stream order transitive.
// M3L_PDL=0 in the environment launches everything fully serialised.
// ---------------------------------------------------------------------------------------
M3L_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
M3L_DEVINL void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(std::forward<Args>(args))...);
}

// ---------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------
M3L_DEVINL uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

M3L_DEVINL bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

M3L_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
M3L_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Exact-erf GELU (nn.GELU() default, as vit_pytorch's FeedForward uses).  erf through Abramowitz &
// Stegun 7.1.26 (|abs err| <= 1.5e-7, i.e. fp32-level): one MUFU.RCP + one MUFU.EX2 + 5 FMA; the
// same exp(-x^2/2) serves the Gaussian pdf of the derivative.
struct GeluParts {
  float cdf;   // Phi(x)
  float pdf;   // phi(x)
};
M3L_DEVINL GeluParts gelu_parts(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));   // MUFU.RCP
  const float e = __expf(-z * z);                       // exp(-x^2 / 2)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, e, 1.0f);       // erf(|x| / sqrt 2)
  GeluParts g;
  g.cdf = 0.5f + copysignf(0.5f * erf_abs, x);
  g.pdf = 0.39894228040143268f * e;
  return g;
}
M3L_DEVINL float gelu_erf(float x) { return x * gelu_parts(x).cdf; }
M3L_DEVINL float gelu_erf_grad(float x) {
  const GeluParts g = gelu_parts(x);
  return fmaf(x, g.pdf, g.cdf);
}

// ---- packed fp32 (Blackwell FFMA2 / FMUL2 / FADD2: two IEEE fp32 lanes per issue slot) -----------------
// The issue-bound epilogues (exact-erf GELU: ~20 scalar instructions per element) spend most of their slots
// on FFMA / FMUL / FADD; the packed forms halve that with bit-identical per-lane arithmetic.
typedef unsigned long long f32x2;
M3L_DEVINL f32x2 f2_pack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
M3L_DEVINL f32x2 f2_packu(uint32_t lo, uint32_t hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
M3L_DEVINL void f2_unpack(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
M3L_DEVINL void f2_unpacku(f32x2 v, uint32_t& lo, uint32_t& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
M3L_DEVINL f32x2 f2_splat(float c) { return f2_pack(c, c); }
M3L_DEVINL f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
M3L_DEVINL f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
M3L_DEVINL f32x2 f2_add(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// GELU(x) and GELU'(x) of two pre-activations at once; same Abramowitz & Stegun 7.1.26 evaluation as
// gelu_parts (one MUFU.RCP + one MUFU.EX2 per element), polynomial / products in packed fp32.
//   u = 0.5 poly(t) t e,  Phi(x) = x >= 0 ? 1 - u : u = base + sg * (-u),  sg = copysign(1, x), base = (sg + 1) / 2
M3L_DEVINL void gelu_pair(uint32_t x0u, uint32_t x1u, f32x2& gelu, f32x2& dgelu) {
  const f32x2 x = f2_packu(x0u, x1u);
  const f32x2 ax = f2_packu(x0u & 0x7fffffffu, x1u & 0x7fffffffu);
  const f32x2 sg = f2_packu((x0u & 0x80000000u) | 0x3f800000u, (x1u & 0x80000000u) | 0x3f800000u);
  const f32x2 den = f2_fma(ax, f2_splat(0.3275911f * 0.70710678118654752f), f2_splat(1.0f));
  float d0, d1;
  f2_unpack(den, d0, d1);
  const f32x2 t = f2_pack(__fdividef(1.0f, d0), __fdividef(1.0f, d1));           // MUFU.RCP x 2
  const f32x2 ea = f2_mul(f2_mul(x, x), f2_splat(-0.5f * 1.4426950408889634f));
  float a0, a1;
  f2_unpack(ea, a0, a1);
  const f32x2 e = f2_pack(exp2f(a0), exp2f(a1));                                  // exp(-x^2 / 2), MUFU.EX2 x 2
  // -0.5 * poly(t): coefficients pre-multiplied
  f32x2 q = f2_fma(f2_splat(-0.5f * 1.061405429f), t, f2_splat(0.5f * 1.453152027f));
  q = f2_fma(q, t, f2_splat(-0.5f * 1.421413741f));
  q = f2_fma(q, t, f2_splat(0.5f * 0.284496736f));
  q = f2_fma(q, t, f2_splat(-0.5f * 0.254829592f));
  const f32x2 nu = f2_mul(f2_mul(q, t), e);                                       // -u
  const f32x2 base = f2_fma(sg, f2_splat(0.5f), f2_splat(0.5f));
  const f32x2 cdf = f2_fma(sg, nu, base);
  gelu = f2_mul(x, cdf);
  dgelu = f2_fma(f2_mul(x, e), f2_splat(0.39894228040143268f), cdf);
}

M3L_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
M3L_DEVINL float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// ---------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------
M3L_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
M3L_DEVINL void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
M3L_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
M3L_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
M3L_DEVINL uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
M3L_DEVINL uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Parity wait with a watchdog: a pipeline bug must trap (-> CUDA error on the host) instead of
// hanging the GPU box.  The watchdog reads %globaltimer only once every 4096 failed polls: a
// timer read per poll (the first version) put a slow special-register access on the wake-up path
// of EVERY barrier hand-off and cost ~2 us per GEMM tile (measured with all memory traffic and
// 15/16 of the MMAs removed, tools/gemm_probe.py).
M3L_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t polls = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++polls & 4095u) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) {
        t0 = now;
      } else if (now - t0 > 4000000000ull) {  // 4 s
        printf("m3l: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x,
               (int)threadIdx.x);
        __trap();
      }
    }
  }
}

// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05.mma operand reads)
M3L_DEVINL void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) loads, completion on an mbarrier
// ---------------------------------------------------------------------------------------
M3L_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
M3L_DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1)
      : "memory");
}
M3L_DEVINL void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2)
      : "memory");
}

// TMA stores (smem -> global, bulk async-group completion).  The smem tile must have been made
// visible to the async proxy (fence_proxy_async_smem) before the issuing thread calls these.
M3L_DEVINL void tma_store_2d(const CUtensorMap* map, uint32_t smem_src_u32, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src_u32), "r"(c0), "r"(c1)
               : "memory");
}
M3L_DEVINL void tma_store_3d(const CUtensorMap* map, uint32_t smem_src_u32, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src_u32), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// named barrier over `nthreads` threads (ids 1..15; 0 is __syncthreads)
M3L_DEVINL void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// global[tile] += smem tile (element type taken from the tensor map; fp32 here)
M3L_DEVINL void tma_reduce_add_2d(const CUtensorMap* map, uint32_t smem_src_u32, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src_u32), "r"(c0), "r"(c1)
               : "memory");
}
M3L_DEVINL void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's bulk groups may still be READING their smem source
template <int N>
M3L_DEVINL void tma_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// at most N of this thread's bulk groups may still be incomplete (writes not yet performed)
template <int N>
M3L_DEVINL void tma_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
M3L_DEVINL void tma_load_2d_u32(uint32_t smem_dst_u32, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_dst_u32), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// ---------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, loads
// ---------------------------------------------------------------------------------------
// One full warp allocates `ncols` (power of two >= 32) TMEM columns; base address -> *smem_slot.
M3L_DEVINL void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
M3L_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
M3L_DEVINL void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
M3L_DEVINL void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; single thread issues on behalf of the CTA.
M3L_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All tcgen05.mma issued so far by this thread -> arrive(1) on `bar` when they complete.
// (implies tcgen05.fence::before_thread_sync)
// Same, issued by the lane(s) whose `pred` is non-zero while the WHOLE warp executes the surrounding code:
// operands computed by converged, warp-uniform code live in uniform registers, whereas an `if (lane == 0)` body
// makes the compiler move every descriptor into uniform registers through an ELECT / R2UR.BROADCAST waterfall
// loop per instruction (~22 SASS instructions, 100-150 clk per MMA; measured on the attention backward issue path).
M3L_DEVINL void umma_bf16_p(uint32_t pred, uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(pred)
      : "memory");
}
M3L_DEVINL void umma_commit_p(uint32_t pred, uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(pred)
      : "memory");
}
M3L_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
M3L_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 columns of 32-bit: thread t of the warp gets lane (base_lane + t), v[j] = column j.
M3L_DEVINL void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31},"
      " [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
M3L_DEVINL void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15},"
      " [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------------------------------
// tcgen05.mma descriptors (bit layouts: PTX ISA "tcgen05 matrix / instruction descriptor")
// ---------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, 128-byte swizzle.
//   bits [0,14)  start address >> 4          bits [16,30) leading-dim byte offset >> 4
//   bits [32,46) stride-dim byte offset >> 4 bits [46,48) version (1 on sm_100)
//   bits [61,64) layout type (2 = SWIZZLE_128B)
// K-major operand tile  = rows of 64 bf16 (128 B), 8-row swizzle atoms 1024 B apart (SBO).
// MN-major operand tile = k-rows of 64 bf16 along M/N (128 B), 8-k-row atoms 1024 B apart (SBO),
//                         further 64-wide M/N blocks `lbo_bytes` apart (LBO).
M3L_DEVINL uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32.
//   [4,6) D fmt (1 = f32)  [7,10) A fmt (1 = bf16)  [10,13) B fmt (1 = bf16)
//   [15] A major (0 = K, 1 = MN)  [16] B major  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, int a_mn_major,
                                                       int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------
// host: TMA tensor maps (driver entry point fetched at run time; no -lcuda link dependency)
// ---------------------------------------------------------------------------------------
// 2-D bf16 row-major matrix [rows, cols] with leading dimension `ld` elements; box = 64 columns
// (128 B, SWIZZLE_128B) x box_rows rows.  Out-of-bounds elements are zero-filled.
int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t ld, uint32_t box_rows);
// bf16 with a 32-column box (64 B rows, SWIZZLE_64B): the half-width epilogue tiles of gemm_gelu.cu
int make_tmap_2d_bf16_sw64(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                           uint64_t ld, uint32_t box_rows);
// same for fp32: box = 32 columns (128 B, SWIZZLE_128B) x box_rows rows
int make_tmap_2d_f32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                     uint64_t ld, uint32_t box_rows);
// 3-D bf16 tensor [d2, d1, d0 (contiguous)] with strides (elements) s2, s1; box = 64 x box1 x 1.
int make_tmap_3d_bf16(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                      uint64_t s1, uint64_t s2, uint32_t box1);

int device_sm_count();

}  // namespace m3l
