// m3l_b200 — flat-arena optimizer kernels: gradient sum of squares, fused clip + AdamW, bf16
// shadow-weight refresh (plain and transposed copies for the dgrad GEMMs).
//
// Reference semantics: /root/reference/models/pretrain_models.py:670-676,707-711 —
//   torch.nn.utils.clip_grad_norm_(params, 0.5): total = ||g||_2 over all grads,
//   g *= min(1, 0.5 / (total + 1e-6)); then torch.optim.AdamW(lr).step() with defaults
//   (betas (0.9, 0.999), eps 1e-8, weight_decay 0.01, decoupled decay, bias-corrected moments).
// All hyper-state lives on the device (step counter, sum of squares) so the whole step replays
// inside a CUDA graph without host round trips.
#include "common.cuh"
#include "m3l_internal.h"

namespace m3l {
namespace {

__global__ void sumsq_kernel(const float* __restrict__ g, size_t n, double* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  __shared__ double red[32];
  double acc = 0.0;
  const size_t n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    acc += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
  }
  for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc += (double)g[i] * g[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    double t = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) atomicAdd(out, t);
  }
}

// state[0] = step counter (as double), state[1] = sum of squares of all grads, state[2] = total norm (out),
// state[3] = 1 - beta1^step, state[4] = sqrt(1 - beta2^step): the bias corrections, in double like the Python
// scalars torch.optim.AdamW computes them in (an fp32 powf per thread is ~1e-4 off for small steps).
// hyper (optional, device): [lr, beta1, beta2, eps, weight_decay, max_norm] — read at run time so that a captured
// CUDA graph follows optimizer.param_groups without being re-captured.
__global__ void step_begin_kernel(double* state, float beta1, float beta2, const float* __restrict__ hyper) {
  pdl_wait();
  pdl_trigger();
  if (hyper != nullptr) { beta1 = hyper[1]; beta2 = hyper[2]; }
  state[0] += 1.0;
  state[2] = sqrt(state[1]);
  state[3] = 1.0 - pow((double)beta1, state[0]);
  state[4] = sqrt(1.0 - pow((double)beta2, state[0]));
}

__global__ void adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, size_t n, const double* __restrict__ state, float lr,
                             float beta1, float beta2, float eps, float weight_decay, float max_norm,
                             int write_clipped_grad, const float* __restrict__ hyper) {
  pdl_wait();
  pdl_trigger();
  if (hyper != nullptr) {
    lr = hyper[0]; beta1 = hyper[1]; beta2 = hyper[2]; eps = hyper[3]; weight_decay = hyper[4]; max_norm = hyper[5];
  }
  const float total = (float)state[2];
  float coef = 1.0f;
  if (max_norm > 0.f) coef = fminf(1.0f, max_norm / (total + 1e-6f));
  const float bc2_sqrt = (float)state[4];
  const float step_size = (float)((double)lr / state[3]);
  const float decay = 1.0f - lr * weight_decay;
  const size_t n4 = n >> 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 gv = reinterpret_cast<float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gg = gp[k] * coef;
      gp[k] = gg;
      pp[k] *= decay;
      mp[k] = mp[k] + (gg - mp[k]) * (1.0f - beta1);
      vp[k] = vp[k] * beta2 + (1.0f - beta2) * gg * gg;
      const float denom = sqrtf(vp[k]) / bc2_sqrt + eps;
      pp[k] -= step_size * (mp[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (write_clipped_grad) reinterpret_cast<float4*>(g)[i] = gv;
  }
  for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gg = g[i] * coef;
    float pp = p[i] * decay;
    const float mm = m[i] + (gg - m[i]) * (1.0f - beta1);
    const float vv = v[i] * beta2 + (1.0f - beta2) * gg * gg;
    pp -= step_size * (mm / (sqrtf(vv) / bc2_sqrt + eps));
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (write_clipped_grad) g[i] = gg;
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n) {
  pdl_wait();
  pdl_trigger();
  const size_t n8 = n >> 3;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(src)[2 * i];
    const float4 b = reinterpret_cast<const float4*>(src)[2 * i + 1];
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w);
    u.z = pack_bf16x2(b.x, b.y); u.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = u;
  }
  for (size_t i = (n8 << 3) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

// teacher[i] = teacher[i] * beta + (1 - beta) * student[i] over a flat fp32 arena (DINO momentum teacher:
// tactile_ssl/utils/ema.py update_moving_average, called from models/vtdino.py:159-173); same operation order as the
// reference (old * beta + (1.0 - beta) * new) so the fp32 results are bit-identical
__global__ void ema_kernel(float* __restrict__ teacher, const float* __restrict__ student, size_t n, float beta, float omb) {
  pdl_wait();
  pdl_trigger();
  const size_t n4 = n >> 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 t = reinterpret_cast<float4*>(teacher)[i];
    const float4 s = reinterpret_cast<const float4*>(student)[i];
    t.x = __fadd_rn(__fmul_rn(t.x, beta), __fmul_rn(omb, s.x));
    t.y = __fadd_rn(__fmul_rn(t.y, beta), __fmul_rn(omb, s.y));
    t.z = __fadd_rn(__fmul_rn(t.z, beta), __fmul_rn(omb, s.z));
    t.w = __fadd_rn(__fmul_rn(t.w, beta), __fmul_rn(omb, s.w));
    reinterpret_cast<float4*>(teacher)[i] = t;
  }
  for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    teacher[i] = __fadd_rn(__fmul_rn(teacher[i], beta), __fmul_rn(omb, student[i]));
}

// transposed bf16 copies of a table of fp32 matrices: dst[c, r] = src[r, c]
__global__ void transpose_cast_kernel(const float* __restrict__ src_base, bf16* __restrict__ dst_base,
                                      const m3l_matrix_desc* __restrict__ descs) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[32][33];
  const m3l_matrix_desc d = descs[blockIdx.y];
  const int tiles_c = (d.cols + 31) / 32, tiles_r = (d.rows + 31) / 32;
  const float* src = src_base + d.src_offset;
  bf16* dst = dst_base + d.dst_offset;
  for (int t = blockIdx.x; t < tiles_r * tiles_c; t += gridDim.x) {
    const int tr = t / tiles_c, tc = t - tr * tiles_c;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int r = tr * 32 + i, c = tc * 32 + threadIdx.x;
      tile[i][threadIdx.x] = (r < d.rows && c < d.cols) ? src[(size_t)r * d.cols + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int c = tc * 32 + i, r = tr * 32 + threadIdx.x;
      if (r < d.rows && c < d.cols) dst[(size_t)c * d.rows + r] = __float2bfloat16(tile[threadIdx.x][i]);
    }
    __syncthreads();
  }
}

int stream_grid(size_t n, int per_thread) {
  size_t blocks = (n / per_thread + 255) / 256;
  const size_t cap = (size_t)device_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace
}  // namespace m3l

using namespace m3l;

extern "C" int m3l_grad_sumsq(const float* grads, size_t count, double* state, void* stream) {
  M3L_REQUIRE(grads && state, "grad_sumsq: null pointer");
  M3L_REQUIRE(((uintptr_t)grads & 15) == 0, "grad_sumsq: grads not 16-byte aligned");
  if (count == 0) return M3L_OK;
  M3L_CUDA(launch_kernel(sumsq_kernel, dim3(stream_grid(count, 4)), dim3(256), 0, (cudaStream_t)stream, grads, count, state + 1));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_optimizer_step_begin(double* state, float beta1, float beta2, const float* hyper_dev, void* stream) {
  M3L_REQUIRE(state, "optimizer_step_begin: null pointer");
  M3L_CUDA(launch_kernel(step_begin_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, state, beta1, beta2, hyper_dev));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_clip_adamw(float* params, float* grads, float* exp_avg, float* exp_avg_sq, size_t count,
                              const double* state, float lr, float beta1, float beta2, float eps,
                              float weight_decay, float max_norm, int write_clipped_grad, const float* hyper_dev,
                              void* stream) {
  M3L_REQUIRE(params && grads && exp_avg && exp_avg_sq && state, "clip_adamw: null pointer");
  M3L_REQUIRE((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
              "clip_adamw: arenas must be 16-byte aligned");
  if (count == 0) return M3L_OK;
  M3L_CUDA(launch_kernel(adamw_kernel, dim3(stream_grid(count, 4)), dim3(256), 0, (cudaStream_t)stream, 
      params, grads, exp_avg, exp_avg_sq, count, state, lr, beta1, beta2, eps, weight_decay, max_norm,
      write_clipped_grad, hyper_dev));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_cast_bf16(const float* src, void* dst_bf16, size_t count, void* stream) {
  M3L_REQUIRE(src && dst_bf16, "cast_bf16: null pointer");
  M3L_REQUIRE((((uintptr_t)src | (uintptr_t)dst_bf16) & 15) == 0, "cast_bf16: pointers must be 16-byte aligned");
  if (count == 0) return M3L_OK;
  M3L_CUDA(launch_kernel(cast_bf16_kernel, dim3(stream_grid(count, 8)), dim3(256), 0, (cudaStream_t)stream, src, (bf16*)dst_bf16, count));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_transpose_cast_bf16(const float* src_base, void* dst_base_bf16, const m3l_matrix_desc* descs_dev,
                                       int count, void* stream) {
  M3L_REQUIRE(src_base && dst_base_bf16 && descs_dev, "transpose_cast_bf16: null pointer");
  if (count == 0) return M3L_OK;
  M3L_CUDA(launch_kernel(transpose_cast_kernel, dim3(dim3(64, count)), dim3(dim3(32, 8)), 0, (cudaStream_t)stream, src_base, (bf16*)dst_base_bf16,
                                                                                  descs_dev));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_ema_update(float* teacher, const float* student, size_t count, float beta, float one_minus_beta,
                              void* stream) {
  M3L_REQUIRE(teacher && student, "ema_update: null pointer");
  M3L_REQUIRE((((uintptr_t)teacher | (uintptr_t)student) & 15) == 0, "ema_update: pointers must be 16-byte aligned");
  if (count == 0) return M3L_OK;
  M3L_CUDA(launch_kernel(ema_kernel, dim3(stream_grid(count, 4)), dim3(256), 0, (cudaStream_t)stream, teacher, student, count, beta, one_minus_beta));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}
