// m3l_b200 — HBM-bound kernels of the VTMAE step: mask sampling, fused patchify + LayerNorm
// gather, token-embedding finish, LayerNorm forward/backward, decoder-sequence assembly and its
// backward, masked-patch MSE (+ its gradient), column sums.
//
// Reference semantics (paths relative to /root/reference):
//   mask sampling             models/pretrain_models.py:223-248
//   patchify + LN(P)          models/pretrain_models.py:768-769,775-776 (Rearrange + LayerNorm)
//   LN(D) + modality + pos    models/pretrain_models.py:771,778,202-219
//   decoder assembly          models/pretrain_models.py:270-307
//   masked MSE                models/pretrain_models.py:327-340 (and :311-322 for early_conv_masking)
#include <type_traits>

#include "common.cuh"
#include "m3l_internal.h"

namespace m3l {
namespace {

// ------------------------------------------------------------------------------------------
// 1. mask sampling: per (sample, segment) rank sort of the noise keys (stable, ascending)
// ------------------------------------------------------------------------------------------
__global__ void mask_indices_kernel(const float* __restrict__ noise, int n_total, m3l_mask_segments segs,
                                    int64_t* __restrict__ masked, int n_masked_total,
                                    int64_t* __restrict__ unmasked, int n_unmasked_total,
                                    int32_t* __restrict__ slot_of_token, int32_t* __restrict__ unmasked_i32,
                                    int32_t* __restrict__ masked_row_of_token, int n_masked_first, int batch) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float keys[];
  const int b = blockIdx.x;
  const int s = blockIdx.y;
  const int off = segs.offset[s], len = segs.length[s], nmask = segs.n_masked[s];
  int moff = 0, uoff = 0;
  for (int i = 0; i < s; ++i) {
    moff += segs.n_masked[i];
    uoff += segs.length[i] - segs.n_masked[i];
  }
  const float* row = noise + (size_t)b * n_total + off;
  for (int i = threadIdx.x; i < len; i += blockDim.x) keys[i] = row[i];
  __syncthreads();
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    const float ki = keys[i];
    int rank = 0;
    for (int j = 0; j < len; ++j) {
      const float kj = keys[j];
      rank += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
    }
    const int64_t tok = off + i;
    if (rank < nmask) {
      const int k = moff + rank;
      masked[(size_t)b * n_masked_total + k] = tok;
      if (slot_of_token) slot_of_token[(size_t)b * n_total + tok] = -(1 + k);
      if (masked_row_of_token)
        masked_row_of_token[(size_t)b * n_total + tok] =
            k < n_masked_first ? b * n_masked_first + k
                               : batch * n_masked_first + b * (n_masked_total - n_masked_first) + (k - n_masked_first);
    } else {
      const int k = uoff + rank - nmask;
      unmasked[(size_t)b * n_unmasked_total + k] = tok;
      if (unmasked_i32) unmasked_i32[(size_t)b * n_unmasked_total + k] = (int32_t)tok;
      if (slot_of_token) slot_of_token[(size_t)b * n_total + tok] = k;
      if (masked_row_of_token) masked_row_of_token[(size_t)b * n_total + tok] = -1;
    }
  }
}

// ------------------------------------------------------------------------------------------
// patch addressing shared by the gather kernels
// ------------------------------------------------------------------------------------------
struct PatchSrc {
  const float* src[4];   // per sensor [B, C, H, W] fp32 (image: one entry); layout 1: raw observation base pointers
  int C, H, W, ph, pw;   // channels, height, width, patch height / width
  int gw;                // patches per row (W / pw)
  int n_per_src;         // patches per source (gh * gw)
  int tok_base;          // global token index of this modality's first token
  int P;                 // ph * pw * C
  // layout 1: the maps are read straight from the RAW observation tensors (vt_load and the 5-D frame-stack reshape
  // fused into the patch loads: utils/pretrain_utils.py:7-57, models/pretrain_models.py:823-827):
  //   element (b, c, y, x), c = f * cg + ch, lives at  b*sb + f*sf + ch*sch + y*sy + x*sx  (elements) and is
  //   normalised as (raw - lo) / span  (uint8 sources: raw / 255 first)
  int layout, u8;
  long long sb;
  int cg, sf, sch, sy, sx;
  float lo, span;
};

// source pointer of sensor s WITHOUT indexing the kernel parameter dynamically (a runtime index makes the compiler copy
// the whole struct to local memory: 120-byte stack frame and an LDL on every row's critical path)
M3L_DEVINL const float* src_of(const PatchSrc& ps, int s) {
  return s == 0 ? ps.src[0] : (s == 1 ? ps.src[1] : (s == 2 ? ps.src[2] : ps.src[3]));
}

M3L_DEVINL const float* patch_origin(const PatchSrc& ps, int b, int tok, int* sensor) {
  const int t = tok - ps.tok_base;
  const int s = t / ps.n_per_src;
  const int tl = t - s * ps.n_per_src;
  const int hh = tl / ps.gw, ww = tl - hh * ps.gw;
  *sensor = s;
  return src_of(ps, s) + ((size_t)b * ps.C * ps.H + (size_t)hh * ps.ph) * ps.W + (size_t)ww * ps.pw;
}

// element e of the flattened patch, order (p1, p2, c) with c fastest
M3L_DEVINL float patch_elem(const PatchSrc& ps, const float* origin, int e) {
  const int c = e % ps.C;
  const int pp = e / ps.C;
  const int p1 = pp / ps.pw, p2 = pp - p1 * ps.pw;
  return origin[((size_t)c * ps.H + p1) * ps.W + p2];
}

// Loads one patch into smem in destination order (p1, p2, c): one thread per (c, p1) source row,
// which is pw contiguous floats along W.
M3L_DEVINL void load_patch_smem(const PatchSrc& ps, const float* origin, float* patch) {
  const int rows = ps.C * ps.ph;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    const int c = i / ps.ph, p1 = i - c * ps.ph;
    const float* src = origin + ((size_t)c * ps.H + p1) * ps.W;
    float* dst = patch + (p1 * ps.pw) * ps.C + c;
    if ((ps.pw & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
      for (int p2 = 0; p2 < ps.pw; p2 += 4) {
        const float4 v = *reinterpret_cast<const float4*>(src + p2);
        dst[(p2 + 0) * ps.C] = v.x; dst[(p2 + 1) * ps.C] = v.y;
        dst[(p2 + 2) * ps.C] = v.z; dst[(p2 + 3) * ps.C] = v.w;
      }
    } else {
      for (int p2 = 0; p2 < ps.pw; ++p2) dst[p2 * ps.C] = src[p2];
    }
  }
}

// ---- layout 1 (raw observations) ------------------------------------------------------------------------
// normalisation exactly as the reference computes it: (x - lo) / (hi - lo) in fp32 with IEEE division (this file is
// compiled without fast-math); uint8 frames are x / 255 first
M3L_DEVINL float raw_norm(const PatchSrc& ps, float v) { return (v - ps.lo) / ps.span; }
M3L_DEVINL float raw_norm_u8(const PatchSrc& ps, unsigned int v) { return ((float)v / 255.0f - ps.lo) / ps.span; }

// element offset of the patch origin (b, token) in the raw tensor of its sensor
M3L_DEVINL long long raw_origin(const PatchSrc& ps, int b, int tok, int* sensor) {
  const int t = tok - ps.tok_base;
  const int s = t / ps.n_per_src;
  const int tl = t - s * ps.n_per_src;
  const int hh = tl / ps.gw, ww = tl - hh * ps.gw;
  *sensor = s;
  return (long long)b * ps.sb + (long long)hh * ps.ph * ps.sy + (long long)ww * ps.pw * ps.sx;
}

// Gathers one patch from a raw observation tensor into `dst` in destination order (p1, p2, c), patch row p1 at
// dst + p1 * pitch; `nthr` cooperating threads, this one is `t`.  Work is split into runs that are CONTIGUOUS in the
// source so that each thread issues wide loads:
//   sx == 1            (channel planes, e.g. the tactile maps [B, F, 6, h, w]): run = (c, p1), pw elements along x
//   sch == 1, sx == cg (interleaved frames, e.g. images [B, F, H, W, 3]):       run = (f, p1), pw * cg elements (x, ch)
//   otherwise          (e.g. NHWC with all stacked channels contiguous [B, H, W, C]): run = (p1, p2), C elements
template <typename T>
M3L_DEVINL void gather_patch_raw(const PatchSrc& ps, const T* base, float* dst, int pitch, int t, int nthr) {
  const int C = ps.C, pw = ps.pw, cg = ps.cg;
  auto cvt = [&](T v) -> float {
    if constexpr (sizeof(T) == 1) return raw_norm_u8(ps, (unsigned int)v); else return raw_norm(ps, (float)v);
  };
  if (ps.sx == 1) {
    const int rows = C * ps.ph;
    for (int i = t; i < rows; i += nthr) {
      const int p1 = i / C, c = i - p1 * C;
      const int f = c / cg, ch = c - f * cg;
      const T* src = base + (long long)f * ps.sf + (long long)ch * ps.sch + (long long)p1 * ps.sy;
      float* d = dst + p1 * pitch + c;
      if constexpr (sizeof(T) == 4) {
        if ((pw & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
          for (int p2 = 0; p2 < pw; p2 += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src + p2));
            d[(p2 + 0) * C] = raw_norm(ps, v.x); d[(p2 + 1) * C] = raw_norm(ps, v.y);
            d[(p2 + 2) * C] = raw_norm(ps, v.z); d[(p2 + 3) * C] = raw_norm(ps, v.w);
          }
          continue;
        }
      }
      for (int p2 = 0; p2 < pw; ++p2) d[p2 * C] = cvt(src[p2]);
    }
  } else if (ps.sch == 1 && ps.sx == cg) {
    const int F = C / cg, rows = F * ps.ph, run = pw * cg;
    for (int i = t; i < rows; i += nthr) {
      const int p1 = i / F, f = i - p1 * F;
      const T* src = base + (long long)f * ps.sf + (long long)p1 * ps.sy;
      float* d = dst + p1 * pitch + f * cg;
      // destination of run element q = p2 * cg + ch is d[p2 * C + ch]
      if constexpr (sizeof(T) == 4) {
        if ((run & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
          int p2 = 0, ch = 0;
          for (int q = 0; q < run; q += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src + q));
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              d[p2 * C + ch] = raw_norm(ps, vv[e]);
              if (++ch == cg) { ch = 0; ++p2; }
            }
          }
          continue;
        }
      } else {
        if ((run & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0) {
          int p2 = 0, ch = 0;
          for (int q = 0; q < run; q += 4) {
            const unsigned int w = __ldg(reinterpret_cast<const unsigned int*>(src + q));
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              d[p2 * C + ch] = raw_norm_u8(ps, (w >> (8 * e)) & 0xffu);
              if (++ch == cg) { ch = 0; ++p2; }
            }
          }
          continue;
        }
      }
      int p2 = 0, ch = 0;
      for (int q = 0; q < run; ++q) {
        d[p2 * C + ch] = cvt(src[q]);
        if (++ch == cg) { ch = 0; ++p2; }
      }
    }
  } else {
    const int rows = ps.ph * pw;
    for (int i = t; i < rows; i += nthr) {
      const int p1 = i / pw, p2 = i - p1 * pw;
      const T* src = base + (long long)p1 * ps.sy + (long long)p2 * ps.sx;
      float* d = dst + p1 * pitch + p2 * C;
      for (int c = 0; c < C; ++c) {
        const int f = c / cg, ch = c - f * cg;
        d[c] = cvt(src[(long long)f * ps.sf + (long long)ch * ps.sch]);
      }
    }
  }
}

// layout-dispatching gather used by the patch kernels (t / nthr as above)
M3L_DEVINL void gather_patch(const PatchSrc& ps, int b, int tok, float* dst, int pitch, int t, int nthr) {
  int sensor;
  const long long off = raw_origin(ps, b, tok, &sensor);
  if (ps.u8) gather_patch_raw(ps, reinterpret_cast<const unsigned char*>(src_of(ps, sensor)) + off, dst, pitch, t, nthr);
  else gather_patch_raw(ps, src_of(ps, sensor) + off, dst, pitch, t, nthr);
}

// Two-stage column reduction without contended atomics (same-sector fp32 atomics serialise at
// roughly one per 10 ns on B200 and dominated these kernels): every block stores its partial
// vector to ws_part[block][nv]; the block that draws the last ticket sums all partials and adds
// them into out[] (plain adds: kernels on a stream are ordered).  ws layout: [0] ticket counter
// (zero on entry, reset to zero on exit), partials from byte 256.
struct ReduceWs {
  unsigned int* counter;
  float* partials;
};
M3L_DEVINL ReduceWs reduce_ws(void* ws) {
  ReduceWs r;
  r.counter = reinterpret_cast<unsigned int*>(ws);
  r.partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + 256);
  return r;
}
// Call from all threads of every block after the block's partial vector (nv floats) has been
// written to ws.partials[linear_block * nv ...] by this block.  `sink(col, total)` is invoked by
// the last block once per column.
template <typename Sink>
M3L_DEVINL void last_block_reduce(const ReduceWs& ws, int nv, int num_blocks, Sink sink) {
  __shared__ unsigned int s_ticket;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) s_ticket = atomicAdd(ws.counter, 1u);
  __syncthreads();
  if (s_ticket != (unsigned)num_blocks - 1) return;
  __threadfence();
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthreads = blockDim.x * blockDim.y;
  for (int col = tid; col < nv; col += nthreads) {
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    int b = 0;
    for (; b + 3 < num_blocks; b += 4) {
      t0 += __ldcg(ws.partials + (size_t)b * nv + col);
      t1 += __ldcg(ws.partials + (size_t)(b + 1) * nv + col);
      t2 += __ldcg(ws.partials + (size_t)(b + 2) * nv + col);
      t3 += __ldcg(ws.partials + (size_t)(b + 3) * nv + col);
    }
    for (; b < num_blocks; ++b) t0 += __ldcg(ws.partials + (size_t)b * nv + col);
    sink(col, (t0 + t1) + (t2 + t3));
  }
  if (tid == 0) *ws.counter = 0u;
}

M3L_DEVINL float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  const int nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

// ------------------------------------------------------------------------------------------
// 2. fused patchify + gather + LayerNorm(P): one block per output row
//    row r = b * ncols + jj ; token = tok_idx ? tok_idx[b, col0 + jj] : tok_base + jj
// ------------------------------------------------------------------------------------------
__global__ void patch_ln_kernel(PatchSrc ps, const int64_t* __restrict__ tok_idx, int idx_ld, int col0,
                                int ncols, const float* __restrict__ gamma, const float* __restrict__ beta,
                                bf16* __restrict__ out, bf16* __restrict__ xhat_out, float eps) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float patch[];   // P floats
  __shared__ float red[32];
  const int r = blockIdx.x;
  const int b = r / ncols, jj = r - b * ncols;
  const int tok = tok_idx ? (int)tok_idx[(size_t)b * idx_ld + col0 + jj] : ps.tok_base + jj;
  const int P = ps.P;
  if (ps.layout == 0) {
    int sensor;
    const float* origin = patch_origin(ps, b, tok, &sensor);
    load_patch_smem(ps, origin, patch);
  } else {
    gather_patch(ps, b, tok, patch, ps.pw * ps.C, threadIdx.x, blockDim.x);
  }
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < P; i += blockDim.x) s += patch[i];
  const float mean = block_sum(s, red) / P;
  float v = 0.f;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const float d = patch[i] - mean;
    v += d * d;
  }
  const float rstd = rsqrtf(block_sum(v, red) / P + eps);
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const float xh = (patch[i] - mean) * rstd;
    out[(size_t)r * P + i] = __float2bfloat16(xh * gamma[i] + beta[i]);
    if (xhat_out) xhat_out[(size_t)r * P + i] = __float2bfloat16(xh);
  }
}

// ------------------------------------------------------------------------------------------
// 2b. plain patchify (no LayerNorm): row r = (b, jj) <- patch of token tok_idx[b, col0 + jj] (or tok_base + jj) in
//     (p1, p2, c) order as bf16, row pitch `ld` >= P, columns [P, ld) zero (K padding for the GEMM that follows).
//     The patch embedding of a ViT whose first layer is Conv2d(kernel = stride = patch), e.g. the frozen DINOv2
//     image branch (train_dino_tac_mae.py:29; weights permuted once to this K order).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
patchify_kernel(PatchSrc ps, const int64_t* __restrict__ tok_idx, int idx_ld, int col0, int ncols,
                bf16* __restrict__ out, int ld) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float patch[];   // P floats
  const int r = blockIdx.x;
  const int b = r / ncols, jj = r - b * ncols;
  const int tok = tok_idx ? (int)tok_idx[(size_t)b * idx_ld + col0 + jj] : ps.tok_base + jj;
  const int P = ps.P;
  if (ps.layout == 0) {
    int sensor;
    const float* origin = patch_origin(ps, b, tok, &sensor);
    load_patch_smem(ps, origin, patch);
  } else {
    gather_patch(ps, b, tok, patch, ps.pw * ps.C, threadIdx.x, blockDim.x);
  }
  __syncthreads();
  bf16* orow = out + (size_t)r * ld;
  for (int i = threadIdx.x * 2; i < ld; i += blockDim.x * 2) {
    const float v0 = i < P ? patch[i] : 0.f, v1 = i + 1 < P ? patch[i + 1] : 0.f;
    *reinterpret_cast<uint32_t*>(orow + i) = pack_bf16x2(v0, v1);
  }
}

// ------------------------------------------------------------------------------------------
// generic warp-per-row LayerNorm helpers (D multiple of 8, D <= 1024); lane owns 8-element chunks
// ------------------------------------------------------------------------------------------
constexpr int kMaxChunks = 4;  // D <= 32 * 8 * 4 = 1024

template <typename T>
M3L_DEVINL void load8(const T* p, float (&v)[8]);
template <>
M3L_DEVINL void load8<bf16>(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
template <>
M3L_DEVINL void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
M3L_DEVINL void store8(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
M3L_DEVINL void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// raw (unconverted) 8-element chunk, so the next row can be prefetched into registers
template <typename T> struct Raw8;
template <> struct Raw8<bf16> { uint4 a; };
template <> struct Raw8<float> { float4 a, b; };
M3L_DEVINL void raw_load(const bf16* p, Raw8<bf16>& r) { r.a = *reinterpret_cast<const uint4*>(p); }
M3L_DEVINL void raw_load(const float* p, Raw8<float>& r) {
  r.a = *reinterpret_cast<const float4*>(p);
  r.b = *reinterpret_cast<const float4*>(p + 4);
}
M3L_DEVINL void raw_cvt(const Raw8<bf16>& r, float (&v)[8]) {
  float2 a = unpack_bf16x2(r.a.x), b = unpack_bf16x2(r.a.y), c = unpack_bf16x2(r.a.z), d = unpack_bf16x2(r.a.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
M3L_DEVINL void raw_cvt(const Raw8<float>& r, float (&v)[8]) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}
M3L_DEVINL void raw_zero(Raw8<bf16>& r) { r.a = make_uint4(0, 0, 0, 0); }

// 3./4. LayerNorm forward.  x: TIn [M, D]; y: bf16.  Optional additive terms (token embedding
// finish / nothing for plain LN): add0[row_class[r]] and add1[row_pos[r]] fp32 rows of D.
// Optional output row remap (dst_row[r] < 0 -> row skipped).  NCH = 8-element chunks per lane.
template <typename TIn, int NCH>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const TIn* __restrict__ x, int M, int D, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, bf16* __restrict__ y,
                     float* __restrict__ stats, const int32_t* __restrict__ dst_row,
                     const float* __restrict__ add0, const int32_t* __restrict__ add0_row,
                     const float* __restrict__ add1, const int32_t* __restrict__ add1_row) {
  pdl_wait();
  pdl_trigger();
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nchunk = D >> 3;
  const float inv_d = 1.0f / D;
  float g[NCH][8], bt[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nchunk) {
      load8<float>(gamma + ch * 8, g[c]);
      load8<float>(beta + ch * 8, bt[c]);
    }
  }
  const int rstride = gridDim.x * warps_per_block;
  int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  Raw8<TIn> nxt[NCH];
  if (r < M) {
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      if (lane + 32 * c < nchunk) raw_load(x + (size_t)r * D + (lane + 32 * c) * 8, nxt[c]);
  }
  for (; r < M; r += rstride) {
    float v[NCH][8];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunk) {
        raw_cvt(nxt[c], v[c]);
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[c][i];
      }
    }
    if (r + rstride < M) {
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        if (lane + 32 * c < nchunk) raw_load(x + (size_t)(r + rstride) * D + (lane + 32 * c) * 8, nxt[c]);
    }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunk) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[c][i] -= mean;
          q += v[c][i] * v[c][i];
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
    if (stats && lane == 0) *reinterpret_cast<float2*>(stats + 2 * r) = make_float2(mean, rstd);
    const int dr = dst_row ? dst_row[r] : r;
    if (dr < 0) continue;
    const float* a0 = add0 ? add0 + (size_t)(add0_row ? add0_row[r] : 0) * D : nullptr;
    const float* a1 = add1 ? add1 + (size_t)(add1_row ? add1_row[r] : 0) * D : nullptr;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunk) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(v[c][i] * rstd, g[c][i], bt[c][i]);
        if (a0) {
          float t[8];
          load8<float>(a0 + ch * 8, t);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] += t[i];
        }
        if (a1) {
          float t[8];
          load8<float>(a1 + ch * 8, t);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] += t[i];
        }
        store8(y + (size_t)dr * D + ch * 8, o);
      }
    }
  }
}

// LayerNorm backward.  dy: bf16 rows (optionally gathered: src_row[r] < 0 -> dy row is zero),
// x: TIn LN input, stats (mean, rstd).  dx = LN'(dy) (+ skip[r]) written as TOut;
// dgamma / dbeta (and optionally the column sums of dx: the bias gradient of the Linear whose
// output gradient dx is) reduced per warp in registers, per block through smem, then fp32 atomics.
template <typename TIn, typename TOut, int NCH>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const bf16* __restrict__ dy, const int32_t* __restrict__ src_row,
                     const TIn* __restrict__ x, const float* __restrict__ stats, int M, int D,
                     const float* __restrict__ gamma, const bf16* __restrict__ skip,
                     TOut* __restrict__ dx, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ dx_colsum) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float spart[];  // [warps][3][D]
  const int warps_per_block = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = D >> 3;
  const float inv_d = 1.0f / D;
  float gm[NCH][8], ag[NCH][8], ab[NCH][8], ac[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nchunk) load8<float>(gamma + ch * 8, gm[c]);
#pragma unroll
    for (int i = 0; i < 8; ++i) ag[c][i] = ab[c][i] = ac[c][i] = 0.f;
  }
  const int rstride = gridDim.x * warps_per_block;
  int r = blockIdx.x * warps_per_block + warp;
  Raw8<TIn> nx[NCH];
  Raw8<bf16> nd[NCH], nk[NCH];
  float2 nms = make_float2(0.f, 0.f);
  auto prefetch = [&](int row) {
    const int sr = src_row ? src_row[row] : row;
    nms = *reinterpret_cast<const float2*>(stats + 2 * row);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunk) {
        raw_load(x + (size_t)row * D + ch * 8, nx[c]);
        if (sr >= 0) raw_load(dy + (size_t)sr * D + ch * 8, nd[c]); else raw_zero(nd[c]);
        if (skip) raw_load(skip + (size_t)row * D + ch * 8, nk[c]);
      }
    }
  };
  if (r < M) prefetch(r);
  for (; r < M; r += rstride) {
    const float mean = nms.x, rstd = nms.y;
    float xh[NCH][8], g[NCH][8], sk[NCH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunk) {
        float xv[8], dv[8];
        raw_cvt(nx[c], xv);
        raw_cvt(nd[c], dv);
        if (skip) raw_cvt(nk[c], sk[c]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xh[c][i] = (xv[i] - mean) * rstd;
          g[c][i] = dv[i] * gm[c][i];
          s1 += g[c][i];
          s2 = fmaf(g[c][i], xh[c][i], s2);
          ag[c][i] = fmaf(dv[i], xh[c][i], ag[c][i]);
          ab[c][i] += dv[i];
        }
      }
    }
    if (r + rstride < M) prefetch(r + rstride);
    s1 = warp_sum(s1) * inv_d;
    s2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunk) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = rstd * (g[c][i] - s1 - xh[c][i] * s2);
        if (skip) {
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] += sk[c][i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) ac[c][i] += o[i];
        store8(dx + (size_t)r * D + ch * 8, o);
      }
    }
  }
  if (dgamma == nullptr && dx_colsum == nullptr) return;
  float* mine = spart + (size_t)warp * 3 * D;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nchunk) {
      store8(mine + ch * 8, ag[c]);
      store8(mine + D + ch * 8, ab[c]);
      store8(mine + 2 * D + ch * 8, ac[c]);
    }
  }
  __syncthreads();
  // grid is capped at two blocks per SM by the host, so each address sees <= ~300 atomics
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < warps_per_block; ++w) t += spart[(size_t)w * 3 * D + i];
    const int which = i / D, col = i - which * D;
    if (which == 0) { if (dgamma) atomicAdd(&dgamma[col], t); }
    else if (which == 1) { if (dbeta) atomicAdd(&dbeta[col], t); }
    else if (dx_colsum) atomicAdd(&dx_colsum[col], t);
  }
}

// ------------------------------------------------------------------------------------------
// bf16 fast paths: warp-private cp.async (LDGSTS) rings keep kStages rows per stream in flight per
// warp without holding them in registers (these kernels are HBM-latency bound otherwise: the ncu
// source page showed ~60 % long-scoreboard stalls on the row loads).
// ------------------------------------------------------------------------------------------
M3L_DEVINL void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
M3L_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
M3L_DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
M3L_DEVINL uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
M3L_DEVINL void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
M3L_DEVINL void cvt8(uint4 u, float (&v)[8]) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

template <int NCH, int kStages>
__global__ void __launch_bounds__(256)
ln_fwd_pipe_kernel(const bf16* __restrict__ x, int M, int D, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float eps, bf16* __restrict__ y, float* __restrict__ stats,
                   const int32_t* __restrict__ dst_row, const float* __restrict__ add0,
                   const int32_t* __restrict__ add0_row, const float* __restrict__ add1,
                   const int32_t* __restrict__ add1_row) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t ring_raw[];
  const int warps_per_block = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = D >> 3;
  const float inv_d = 1.0f / D;
  constexpr int kRowBytes = NCH * 512;
  const uint32_t ring = smem_u32(ring_raw) + warp * (kStages * kRowBytes);
  float g[NCH][8], bt[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nchunk) {
      load8<float>(gamma + ch * 8, g[c]);
      load8<float>(beta + ch * 8, bt[c]);
    }
  }
  const int rstride = gridDim.x * warps_per_block;
  const int r0 = blockIdx.x * warps_per_block + warp;
  auto issue = [&](int row, int stage) {
    if (row < M) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunk) cp_async16(ring + stage * kRowBytes + ch * 16, x + (size_t)row * D + ch * 8);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < kStages; ++s) issue(r0 + s * rstride, s);
  int stage = 0;
  for (int r = r0; r < M; r += rstride) {
    cp_async_wait<kStages - 1>();
    float v[NCH][8];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunk) {
        cvt8(lds128(ring + stage * kRowBytes + ch * 16), v[c]);
#pragma unroll
        for (int i = 0; i < 8; ++i) sum += v[c][i];
      }
    }
    issue(r + kStages * rstride, stage);
    stage = (stage + 1 == kStages) ? 0 : stage + 1;
    const float mean = warp_sum(sum) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if (lane + 32 * c < nchunk) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[c][i] -= mean;
          q += v[c][i] * v[c][i];
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
    if (stats && lane == 0) *reinterpret_cast<float2*>(stats + 2 * r) = make_float2(mean, rstd);
    const int dr = dst_row ? dst_row[r] : r;
    if (dr < 0) continue;
    const float* a0 = add0 ? add0 + (size_t)(add0_row ? add0_row[r] : 0) * D : nullptr;
    const float* a1 = add1 ? add1 + (size_t)(add1_row ? add1_row[r] : 0) * D : nullptr;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunk) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(v[c][i] * rstd, g[c][i], bt[c][i]);
        if (a0) {
          float t[8];
          load8<float>(a0 + ch * 8, t);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] += t[i];
        }
        if (a1) {
          float t[8];
          load8<float>(a1 + ch * 8, t);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] += t[i];
        }
        store8(y + (size_t)dr * D + ch * 8, o);
      }
    }
  }
  cp_async_wait<0>();
}

template <int NCH, int kStages>
__global__ void __launch_bounds__(256)
ln_bwd_pipe_kernel(const bf16* __restrict__ dy, const int32_t* __restrict__ src_row, const bf16* __restrict__ x,
                   const float* __restrict__ stats, int M, int D, const float* __restrict__ gamma,
                   const bf16* __restrict__ skip, bf16* __restrict__ dx, float* __restrict__ dgamma,
                   float* __restrict__ dbeta, float* __restrict__ dx_colsum) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t ring_raw[];
  const int warps_per_block = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = D >> 3;
  const float inv_d = 1.0f / D;
  constexpr int kRowBytes = NCH * 512;
  constexpr int kStageBytes = 3 * kRowBytes;           // x | dy | skip
  const uint32_t ring = smem_u32(ring_raw) + warp * (kStages * kStageBytes);
  float* spart = reinterpret_cast<float*>(ring_raw + (size_t)warps_per_block * kStages * kStageBytes);  // [warps][3][D]
  float gm[NCH][8], ag[NCH][8], ab[NCH][8], ac[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nchunk) load8<float>(gamma + ch * 8, gm[c]);
#pragma unroll
    for (int i = 0; i < 8; ++i) ag[c][i] = ab[c][i] = ac[c][i] = 0.f;
  }
  const int rstride = gridDim.x * warps_per_block;
  const int r0 = blockIdx.x * warps_per_block + warp;
  auto issue = [&](int row, int stage) {
    if (row < M) {
      const int sr = src_row ? src_row[row] : row;
      const uint32_t base = ring + stage * kStageBytes;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunk) {
          cp_async16(base + ch * 16, x + (size_t)row * D + ch * 8);
          if (sr >= 0) cp_async16(base + kRowBytes + ch * 16, dy + (size_t)sr * D + ch * 8);
          else sts128(base + kRowBytes + ch * 16, make_uint4(0, 0, 0, 0));
          if (skip) cp_async16(base + 2 * kRowBytes + ch * 16, skip + (size_t)row * D + ch * 8);
        }
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < kStages; ++s) issue(r0 + s * rstride, s);
  int stage = 0;
  // TWO rows per iteration: their dependent chains (smem reads -> sums -> two warp reductions -> output) interleave,
  // which hides the ~500 clk of per-row latency that four warps per scheduler could not (IPC 0.45 before)
  static_assert(kStages % 2 == 0 && kStages >= 4, "two rows in compute, at least two in flight");
  for (int r = r0; r < M; r += 2 * rstride) {
    const int rb = r + rstride;
    const bool has_b = rb < M;
    const float2 ms_a = *reinterpret_cast<const float2*>(stats + 2 * r);
    const float2 ms_b = has_b ? *reinterpret_cast<const float2*>(stats + 2 * rb) : make_float2(0.f, 0.f);
    cp_async_wait<kStages - 2>();
    const uint32_t base_a = ring + stage * kStageBytes, base_b = ring + (stage + 1) * kStageBytes;
    float xh_a[NCH][8], g_a[NCH][8], sk_a[NCH][8], xh_b[NCH][8], g_b[NCH][8], sk_b[NCH][8];
    float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
    auto phase1 = [&](uint32_t base, float mean, float rstd, bool on, float (&xh)[NCH][8], float (&g)[NCH][8],
                      float (&sk)[NCH][8], float& s1, float& s2) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunk) {
          float xv[8], dv[8];
          cvt8(lds128(base + ch * 16), xv);
          cvt8(lds128(base + kRowBytes + ch * 16), dv);
          if (skip) cvt8(lds128(base + 2 * kRowBytes + ch * 16), sk[c]);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (!on) dv[i] = 0.f;                      // row beyond M: contributes nothing to the parameter sums
            xh[c][i] = (xv[i] - mean) * rstd;
            g[c][i] = dv[i] * gm[c][i];
            s1 += g[c][i];
            s2 = fmaf(g[c][i], xh[c][i], s2);
            ag[c][i] = fmaf(dv[i], xh[c][i], ag[c][i]);
            ab[c][i] += dv[i];
          }
        }
      }
    };
    phase1(base_a, ms_a.x, ms_a.y, true, xh_a, g_a, sk_a, s1a, s2a);
    if (has_b) phase1(base_b, ms_b.x, ms_b.y, true, xh_b, g_b, sk_b, s1b, s2b);   // never touch an unloaded stage (0 * NaN)
    issue(r + kStages * rstride, stage);
    issue(rb + kStages * rstride, stage + 1);
    stage = (stage + 2 == kStages) ? 0 : stage + 2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                 // four interleaved butterflies
      s1a += __shfl_xor_sync(0xffffffffu, s1a, o); s2a += __shfl_xor_sync(0xffffffffu, s2a, o);
      s1b += __shfl_xor_sync(0xffffffffu, s1b, o); s2b += __shfl_xor_sync(0xffffffffu, s2b, o);
    }
    s1a *= inv_d; s2a *= inv_d; s1b *= inv_d; s2b *= inv_d;
    auto phase2 = [&](int row, float rstd, float s1, float s2, const float (&xh)[NCH][8], const float (&g)[NCH][8],
                      const float (&sk)[NCH][8]) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunk) {
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = rstd * (g[c][i] - s1 - xh[c][i] * s2);
          if (skip) {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] += sk[c][i];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) ac[c][i] += o[i];
          store8(dx + (size_t)row * D + ch * 8, o);
        }
      }
    };
    phase2(r, ms_a.y, s1a, s2a, xh_a, g_a, sk_a);
    if (has_b) phase2(rb, ms_b.y, s1b, s2b, xh_b, g_b, sk_b);
  }
  cp_async_wait<0>();
  if (dgamma == nullptr && dx_colsum == nullptr) return;
  float* mine = spart + (size_t)warp * 3 * D;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nchunk) {
      store8(mine + ch * 8, ag[c]);
      store8(mine + D + ch * 8, ab[c]);
      store8(mine + 2 * D + ch * 8, ac[c]);
    }
  }
  __syncthreads();
  // grid is capped at two blocks per SM by the host, so each address sees <= ~300 atomics
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < warps_per_block; ++w) t += spart[(size_t)w * 3 * D + i];
    const int which = i / D, col = i - which * D;
    if (which == 0) { if (dgamma) atomicAdd(&dgamma[col], t); }
    else if (which == 1) { if (dbeta) atomicAdd(&dbeta[col], t); }
    else if (dx_colsum) atomicAdd(&dx_colsum[col], t);
  }
}

// ------------------------------------------------------------------------------------------
// 5. decoder sequence assembly: z[b, t] = (slot >= 0 ? d[b*nv + slot] : mask_token) + add0[cls] + add1[t]
// ------------------------------------------------------------------------------------------
__global__ void assemble_fwd_kernel(const bf16* __restrict__ d, int nv, const float* __restrict__ mask_token,
                                    const int32_t* __restrict__ slot_of_token, int B, int n, int D,
                                    const float* __restrict__ add0, const int32_t* __restrict__ tok_class,
                                    const float* __restrict__ add1, bf16* __restrict__ z) {
  pdl_wait();
  pdl_trigger();
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nchunk = D >> 3;
  const int M = B * n;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < M; r += gridDim.x * warps_per_block) {
    const int b = r / n, t = r - b * n;
    const int slot = slot_of_token[r];
    for (int ch = lane; ch < nchunk; ch += 32) {
      float v[8];
      if (slot >= 0) load8<bf16>(d + ((size_t)b * nv + slot) * D + ch * 8, v);
      else load8<float>(mask_token + ch * 8, v);
      if (add0) {
        float a[8];
        load8<float>(add0 + (size_t)tok_class[t] * D + ch * 8, a);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += a[i];
      }
      if (add1) {
        float a[8];
        load8<float>(add1 + (size_t)t * D + ch * 8, a);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += a[i];
      }
      store8(z + (size_t)r * D + ch * 8, v);
    }
  }
}

// backward: dd[b*nv+slot] = dz[b,t] (visible); d mask_token += sum(masked rows);
// d add0[class] += sum rows of the class; d add1[t] += sum over b (only if dadd1 != null; that
// learned-position variant keeps atomics: its sums are per token, not contended).
// Warp per row (lane = 8 columns), accumulators [masked | class 0..3] in registers, block partials
// through smem, then one atomic per block and column (grid capped at two blocks per SM).
constexpr int kAsmClasses = 4;
template <int NCH>
__global__ void __launch_bounds__(256)
assemble_bwd_kernel(const bf16* __restrict__ dz, const int32_t* __restrict__ slot_of_token,
                    int B, int n, int D, int nv, bf16* __restrict__ dd,
                    float* __restrict__ dmask_token, float* __restrict__ dadd0,
                    const int32_t* __restrict__ tok_class, int n_classes, float* __restrict__ dadd1) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sred[];   // [warps][(1 + kAsmClasses)][D]
  const int warps_per_block = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = D >> 3;
  const int M = B * n;
  float acc[1 + kAsmClasses][NCH][8];
#pragma unroll
  for (int k = 0; k < 1 + kAsmClasses; ++k)
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[k][c][i] = 0.f;
  const int rstride = gridDim.x * warps_per_block;
  int r = blockIdx.x * warps_per_block + warp;
  Raw8<bf16> nxt[NCH];
  int nslot = 0;
  auto prefetch = [&](int row) {
    nslot = slot_of_token[row];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      if (lane + 32 * c < nchunk) raw_load(dz + (size_t)row * D + (lane + 32 * c) * 8, nxt[c]);
  };
  if (r < M) prefetch(r);
  for (; r < M; r += rstride) {
    const int b = r / n, t = r - b * n;
    const int slot = nslot;
    Raw8<bf16> cur[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) cur[c] = nxt[c];
    if (r + rstride < M) prefetch(r + rstride);
    const int cls = (dadd0 && tok_class) ? tok_class[t] : 0;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunk) {
        float v[8];
        raw_cvt(cur[c], v);
        if (slot >= 0) {
          *reinterpret_cast<uint4*>(dd + ((size_t)b * nv + slot) * D + ch * 8) = cur[c].a;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[0][c][i] += v[i];
        }
#pragma unroll
        for (int k = 0; k < kAsmClasses; ++k)
          if (cls == k) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[1 + k][c][i] += v[i];
          }
        if (dadd1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) atomicAdd(&dadd1[(size_t)t * D + ch * 8 + i], v[i]);
        }
      }
    }
  }
  const int nvec = (1 + kAsmClasses) * D;
  float* mine = sred + (size_t)warp * nvec;
#pragma unroll
  for (int k = 0; k < 1 + kAsmClasses; ++k)
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunk) store8(mine + k * D + ch * 8, acc[k][c]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    float tsum = 0.f;
    for (int w = 0; w < warps_per_block; ++w) tsum += sred[(size_t)w * nvec + i];
    const int k = i / D, col = i - k * D;
    if (k == 0) { if (dmask_token) atomicAdd(&dmask_token[col], tsum); }
    else if (dadd0 && (k - 1) < n_classes) atomicAdd(&dadd0[(size_t)(k - 1) * D + col], tsum);
  }
}

// token-embedding finish backward: class / position sums of dx0 rows (rows ordered [b, nv]).
__global__ void rowclass_sum_kernel(const bf16* __restrict__ dx, int B, int nv, int D,
                                    const int32_t* __restrict__ row_class, float* __restrict__ dclass,
                                    const int32_t* __restrict__ row_pos, float* __restrict__ dpos) {
  pdl_wait();
  pdl_trigger();
  // block (j, y): visible slot j (class is a function of j only; position varies per sample),
  // samples y, y + gridDim.y, ...
  const int j = blockIdx.x;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float s = 0.f;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
      const size_t r = (size_t)b * nv + j;
      const float g = __bfloat162float(dx[r * D + c]);
      s += g;
      if (dpos) atomicAdd(&dpos[(size_t)row_pos[r] * D + c], g);
    }
    if (dclass) atomicAdd(&dclass[(size_t)row_class[j] * D + c], s);
  }
}

// ------------------------------------------------------------------------------------------
// 6. masked-patch MSE: pred fp32 [rows, P]; target gathered from the raw maps.
//    loss_acc[slot] += weight * sum((pred - tgt)^2) ; dpred = 2 * weight * (pred - tgt)  (bf16)
//    dcolsum[p] += sum_rows dpred[row, p]  (optional: the bias gradient of the head Linear)
//    Warp per row.  The patch is transposed through smem from the source order (c, p1, p2) to the
//    destination order (p1, p2, c): lanes walk (p1, c) with c fastest and every p1 row of the patch
//    has pitch pw*C + pad floats with pitch % 32 == C % 32, which keeps the scattered stores
//    conflict free (the first version put all 8 p1 rows of a lane group on one bank).  The final
//    loss reduction over the block partials is done by ALL threads of the last block in a fixed
//    order (deterministic, and not a 1184-long chain of L2 round trips on one thread).
//    GATHER selects how the target patch reaches shared memory.  The generic loops (0) issue one wide load, scatter
//    its values, then issue the next: with run lengths only known at run time nothing is hoisted, and a row costs up to
//    six DEPENDENT DRAM round trips (r02 ncu source page: 34 % of all stall samples on the first scatter store, long
//    scoreboard; 50 us for 118 MB).  The specialised forms load every run of the lane into registers first and
//    scatter afterwards (compile-time run count SEGS and length V4):
//      1  channel planes, fp32, x contiguous (layout 0 maps, or raw tactile [B, F, 6, h, w]): SEGS runs of 4*V4 floats
//      2  interleaved fp32 frames (raw images [B, F, H, W, 3]): one run of 4*V4 floats per lane
//      3  interleaved uint8 frames: one run of 4*V4 bytes per lane
//    The token index of the warp's NEXT row is fetched one iteration ahead (it heads the chain of every row).
// ------------------------------------------------------------------------------------------
constexpr int kMseMaxT = 8;   // column sums kept in registers: P <= 128 * 4 * ... = 1024
template <int GATHER, int SEGS, int V4, int T, int MINB>      // T: float4 per lane and row (P <= 128 T); MINB: blocks per SM
__global__ void __launch_bounds__(256, MINB)
mse_loss_kernel(PatchSrc ps, const int64_t* __restrict__ tok_idx, int idx_ld, int col0, int ncols, int rows,
                const float* __restrict__ pred, float weight, bf16* __restrict__ dpred,
                float* __restrict__ loss_acc, float* __restrict__ dcolsum, int pitch, void* ws) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) float patches[];   // [warps][ph * pitch] floats, then [P] column sums
  __shared__ float red[8];
  __shared__ unsigned int s_ticket;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int P = ps.P;
  const int rowlen = ps.pw * ps.C;                   // floats per patch row p1 (destination order)
  float* patch = patches + (size_t)warp * ps.ph * pitch;
  float* s_cs = patches + (size_t)nwarps * ps.ph * pitch;
  const int src_rows = ps.C * ps.ph;
  const float w2 = 2.f * weight;
  float acc = 0.f;
  float cs[T][4];
#pragma unroll
  for (int t = 0; t < T; ++t) cs[t][0] = cs[t][1] = cs[t][2] = cs[t][3] = 0.f;
  if (dcolsum) {
    for (int i = threadIdx.x; i < P; i += blockDim.x) s_cs[i] = 0.f;
  }
  auto token_of = [&](int r) -> int {
    const int b = r / ncols, jj = r - b * ncols;
    return tok_idx ? (int)tok_idx[(size_t)b * idx_ld + col0 + jj] : ps.tok_base + jj;
  };
  const int stride = gridDim.x * nwarps;
  int r = blockIdx.x * nwarps + warp;
  int tok_next = r < rows ? token_of(r) : 0;
  for (; r < rows; r += stride) {
    const int b = r / ncols;
    const int tok = tok_next;
    if (r + stride < rows) tok_next = token_of(r + stride);      // consumed one iteration later
    int sensor = 0;
    const float* origin = (GATHER == 0 && ps.layout == 0) ? patch_origin(ps, b, tok, &sensor) : nullptr;
    // the predictions of this row are requested BEFORE the target gather so that both global round trips
    // overlap (they used to be exposed back to back, one row at a time per warp)
    const float* prow = pred + (size_t)r * P;
    float4 pvr[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int e = t * 128 + lane * 4;
      if (e < P) pvr[t] = __ldcs(reinterpret_cast<const float4*>(prow + e));
    }
    if constexpr (GATHER == 1) {
      // channel planes: run i = (p1, c), c fastest, 4 * V4 floats along x.  All loads first, then the scatter.
      const float* base;
      if (ps.layout == 0) {
        base = patch_origin(ps, b, tok, &sensor);
      } else {
        const long long off = raw_origin(ps, b, tok, &sensor);
        base = src_of(ps, sensor) + off;
      }
      float4 tg[SEGS][V4];
#pragma unroll
      for (int k = 0; k < SEGS; ++k) {
        const int i = lane + 32 * k;
        if (i < src_rows) {
          const int p1 = i / ps.C, c = i - p1 * ps.C;
          const float* src;
          if (ps.layout == 0) {
            src = base + ((size_t)c * ps.H + p1) * ps.W;
          } else {
            const int f = c / ps.cg, ch = c - f * ps.cg;
            src = base + (long long)f * ps.sf + (long long)ch * ps.sch + (long long)p1 * ps.sy;
          }
#pragma unroll
          for (int v = 0; v < V4; ++v) tg[k][v] = __ldg(reinterpret_cast<const float4*>(src) + v);
        }
      }
#pragma unroll
      for (int k = 0; k < SEGS; ++k) {
        const int i = lane + 32 * k;
        if (i < src_rows) {
          const int p1 = i / ps.C, c = i - p1 * ps.C;
          float* d = patch + p1 * pitch + c;
#pragma unroll
          for (int v = 0; v < V4; ++v) {
            const float4 t4 = tg[k][v];
            if (ps.layout == 0) {
              d[(4 * v + 0) * ps.C] = t4.x; d[(4 * v + 1) * ps.C] = t4.y;
              d[(4 * v + 2) * ps.C] = t4.z; d[(4 * v + 3) * ps.C] = t4.w;
            } else {
              d[(4 * v + 0) * ps.C] = raw_norm(ps, t4.x); d[(4 * v + 1) * ps.C] = raw_norm(ps, t4.y);
              d[(4 * v + 2) * ps.C] = raw_norm(ps, t4.z); d[(4 * v + 3) * ps.C] = raw_norm(ps, t4.w);
            }
          }
        }
      }
    } else if constexpr (GATHER == 2 || GATHER == 3) {
      // interleaved frames: run i = (p1, f), f fastest; run element q = p2 * cg + ch goes to (p1, p2, f * cg + ch)
      using RawT = typename std::conditional<GATHER == 2, float, unsigned char>::type;
      using VecT = typename std::conditional<GATHER == 2, float4, unsigned int>::type;
      const int F = ps.C / ps.cg;
      const long long off = raw_origin(ps, b, tok, &sensor);
      const RawT* base = reinterpret_cast<const RawT*>(src_of(ps, sensor)) + off;
      const int p1 = lane / F, f = lane - p1 * F;
      VecT tg[V4];
      if (lane < F * ps.ph) {
        const RawT* src = base + (long long)f * ps.sf + (long long)p1 * ps.sy;
#pragma unroll
        for (int v = 0; v < V4; ++v) tg[v] = __ldg(reinterpret_cast<const VecT*>(src) + v);
        float* d = patch + p1 * pitch + f * ps.cg;
        int p2 = 0, ch = 0;
#pragma unroll
        for (int v = 0; v < V4; ++v) {
          float vv[4];
          if constexpr (GATHER == 2) {
            const float4 t4 = *reinterpret_cast<const float4*>(&tg[v]);
            vv[0] = raw_norm(ps, t4.x); vv[1] = raw_norm(ps, t4.y); vv[2] = raw_norm(ps, t4.z); vv[3] = raw_norm(ps, t4.w);
          } else {
            const unsigned int w = *reinterpret_cast<const unsigned int*>(&tg[v]);
#pragma unroll
            for (int e = 0; e < 4; ++e) vv[e] = raw_norm_u8(ps, (w >> (8 * e)) & 0xffu);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            d[p2 * ps.C + ch] = vv[e];
            if (++ch == ps.cg) { ch = 0; ++p2; }
          }
        }
      }
    } else if (ps.layout != 0) {
      gather_patch(ps, b, tok, patch, pitch, lane, 32);     // raw observations (vt_load fused)
    }
    for (int i = lane; i < (origin != nullptr ? src_rows : 0); i += 32) {      // one (patch row, channel) per lane: pw contiguous floats
      const int p1 = i / ps.C, c = i - p1 * ps.C;
      const float* src = origin + ((size_t)c * ps.H + p1) * ps.W;
      float* dst = patch + p1 * pitch + c;
      if ((ps.pw & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        for (int p2 = 0; p2 < ps.pw; p2 += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src + p2));
          dst[(p2 + 0) * ps.C] = v.x; dst[(p2 + 1) * ps.C] = v.y;
          dst[(p2 + 2) * ps.C] = v.z; dst[(p2 + 3) * ps.C] = v.w;
        }
      } else {
        for (int p2 = 0; p2 < ps.pw; ++p2) dst[p2 * ps.C] = src[p2];
      }
    }
    __syncwarp();
    bf16* drow = dpred + (size_t)r * P;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int e = t * 128 + lane * 4;
      if (e < P) {
        const int p1 = e / rowlen;
        const float4 tv = *reinterpret_cast<const float4*>(patch + p1 * pitch + (e - p1 * rowlen));
        const float4 pv = pvr[t];
        const float d0 = pv.x - tv.x, d1 = pv.y - tv.y, d2 = pv.z - tv.z, d3 = pv.w - tv.w;
        acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        const float g0 = w2 * d0, g1 = w2 * d1, g2 = w2 * d2, g3 = w2 * d3;
        *reinterpret_cast<uint2*>(drow + e) = make_uint2(pack_bf16x2(g0, g1), pack_bf16x2(g2, g3));
        cs[t][0] += g0; cs[t][1] += g1; cs[t][2] += g2; cs[t][3] += g3;
      }
    }
    for (int e = T * 128 + lane * 4; e < P; e += 128) {      // P > 1024: no column sums
      const int p1 = e / rowlen;
      const float4 tv = *reinterpret_cast<const float4*>(patch + p1 * pitch + (e - p1 * rowlen));
      const float4 pv = *reinterpret_cast<const float4*>(prow + e);
      const float d0 = pv.x - tv.x, d1 = pv.y - tv.y, d2 = pv.z - tv.z, d3 = pv.w - tv.w;
      acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      *reinterpret_cast<uint2*>(drow + e) = make_uint2(pack_bf16x2(w2 * d0, w2 * d1), pack_bf16x2(w2 * d2, w2 * d3));
    }
    __syncwarp();
  }
  if (dcolsum) {
    __syncthreads();
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int e = t * 128 + lane * 4;
      if (e < P) {
        atomicAdd(&s_cs[e], cs[t][0]); atomicAdd(&s_cs[e + 1], cs[t][1]);
        atomicAdd(&s_cs[e + 2], cs[t][2]); atomicAdd(&s_cs[e + 3], cs[t][3]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P; i += blockDim.x) atomicAdd(&dcolsum[i], s_cs[i]);
  }
  acc = warp_sum(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  const ReduceWs rws = reduce_ws(ws);
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < nwarps; ++w) t += red[w];
    rws.partials[blockIdx.x] = weight * t;
    __threadfence();
    s_ticket = atomicAdd(rws.counter, 1u);
  }
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  float t = 0.f;                                        // fixed assignment of partials to threads
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(rws.partials + i);
  t = warp_sum(t);
  __syncthreads();
  if (lane == 0) red[warp] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < nwarps; ++w) tot += red[w];
    atomicAdd(loss_acc, tot);      // the two heads may run on parallel graph branches; a two-term sum is order free
    *rws.counter = 0u;
  }
}

// ------------------------------------------------------------------------------------------
// 7. column sums of a bf16 matrix: out[n] += sum_m x[m, n]   (bias gradients)
// ------------------------------------------------------------------------------------------
__global__ void colsum_kernel(const bf16* __restrict__ x, int M, int N, int ld, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  // block: 32 x 8 threads; each thread owns 8 consecutive columns, 4 rows in flight
  const int col = (blockIdx.x * 32 + threadIdx.x) * 8;
  __shared__ float part[8][32][8];
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (col < N) {
    const int rows_per_block = (M + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(r0 + rows_per_block, M);
    int r = r0 + threadIdx.y;
    for (; r + 56 < r1; r += 64) {
      Raw8<bf16> raw[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) raw_load(x + (size_t)(r + 8 * u) * ld + col, raw[u]);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float v[8];
        raw_cvt(raw[u], v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += v[i];
      }
    }
    for (; r < r1; r += 8) {
      float v[8];
      load8<bf16>(x + (size_t)r * ld + col, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) part[threadIdx.y][threadIdx.x][i] = acc[i];
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = 0.f;
      for (int y = 0; y < 8; ++y) s += part[y][threadIdx.x][i];
      atomicAdd(&out[col + i], s);
    }
  }
}

// LayerNorm(P) parameter gradients of the patch embedding: dgamma[p] += sum_r dA[r,p]*xhat[r,p]; dbeta[p] += sum_r dA[r,p]
// block 32 x 8 threads; thread owns 8 consecutive columns (16-byte loads), rows strided by 8
__global__ void ln_param_grad_kernel(const bf16* __restrict__ dA, const bf16* __restrict__ xhat, int M, int P,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_wait();
  pdl_trigger();
  const int col = (blockIdx.x * 32 + threadIdx.x) * 8;
  __shared__ float part[2][8][32][8];
  float g[8], bs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) g[i] = bs[i] = 0.f;
  if (col < P) {
    const int rows_per = (M + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per, r1 = min(r0 + rows_per, M);
#pragma unroll 4
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      float d[8], x[8];
      load8<bf16>(dA + (size_t)r * P + col, d);
      load8<bf16>(xhat + (size_t)r * P + col, x);
#pragma unroll
      for (int i = 0; i < 8; ++i) { g[i] = fmaf(d[i], x[i], g[i]); bs[i] += d[i]; }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { part[0][threadIdx.y][threadIdx.x][i] = g[i]; part[1][threadIdx.y][threadIdx.x][i] = bs[i]; }
  __syncthreads();
  if (threadIdx.y < 2 && col < P) {
    float* out = threadIdx.y == 0 ? dgamma : dbeta;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t = 0.f;
      for (int y = 0; y < 8; ++y) t += part[threadIdx.y][y][threadIdx.x][i];
      atomicAdd(&out[col + i], t);
    }
  }
}

// ------------------------------------------------------------------------------------------
// 8. token mean (MAEExtractor: torch.mean(tokens, dim=1), pretrain_models.py:837) and its backward
//    fwd: out[b, :] = mean_t x[b, t, :]   (bf16 in, fp32 out; one warp per (sample, 256-column slab))
//    bwd: dx[b, t, :] = dout[b, :] / n    (fp32 in, bf16 out)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
token_mean_fwd_kernel(const bf16* __restrict__ x, int B, int n, int D, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int slabs = (D + 255) / 256;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= B * slabs) return;
  const int b = w / slabs, col = (w - b * slabs) * 256 + lane * 8;
  if (col >= D) return;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const bf16* p = x + (size_t)b * n * D + col;
#pragma unroll 4
  for (int t = 0; t < n; ++t) {
    float v[8];
    load8<bf16>(p + (size_t)t * D, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += v[i];
  }
  const float inv = 1.0f / n;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] *= inv;
  store8(out + (size_t)b * D + col, acc);
}

__global__ void __launch_bounds__(256)
token_mean_bwd_kernel(const float* __restrict__ dout, int B, int n, int D, bf16* __restrict__ dx) {
  pdl_wait();
  pdl_trigger();
  const int chunks = D >> 3;
  const size_t total = (size_t)B * n * chunks;
  const float inv = 1.0f / n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chunks);
    const int b = (int)(i / ((size_t)n * chunks));
    float v[8];
    load8<float>(dout + (size_t)b * D + ch * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= inv;
    store8(dx + i * 8, v);
  }
}

// ------------------------------------------------------------------------------------------
// 9. EarlyCNN conv stem (early_conv_masking=True, pretrain_models.py:37-56,180-191): the convolutions run
//    as im2col + tcgen05 GEMM (+bias +ReLU epilogue); activations are NHWC bf16 ([B*H*W, C] matrices, which
//    is also the token layout flatten(2).transpose(1,2) produces).  K order of the im2col matrix is
//    (cin, ky, kx) = the flattening of nn.Conv2d.weight [cout, cin, kh, kw], so weights are used as stored.
// ------------------------------------------------------------------------------------------
template <bool NHWC_BF16>
__global__ void __launch_bounds__(256)
im2col_kernel(const void* __restrict__ xin, int B, int C, int H, int W, int k, int stride, int pad, int Ho, int Wo,
              bf16* __restrict__ col) {
  pdl_wait();
  pdl_trigger();
  const int K = C * k * k;
  const int K2 = K >> 1;                                   // two K entries per thread (K is even: k*k even or C even)
  const size_t total = (size_t)B * Ho * Wo * K2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i % K2) * 2;
    const size_t m = i / K2;
    const int ox = (int)(m % Wo), oy = (int)((m / Wo) % Ho), b = (int)(m / ((size_t)Wo * Ho));
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int kq = kk + e;
      const int ci = kq / (k * k), r = kq - ci * k * k, ky = r / k, kx = r - ky * k;
      const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
      float t = 0.f;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        if (NHWC_BF16) t = __bfloat162float(reinterpret_cast<const bf16*>(xin)[(((size_t)b * H + iy) * W + ix) * C + ci]);
        else t = reinterpret_cast<const float*>(xin)[(((size_t)b * C + ci) * H + iy) * W + ix];
      }
      v[e] = t;
    }
    *reinterpret_cast<uint32_t*>(col + m * K + kk) = pack_bf16x2(v[0], v[1]);
  }
}

// dgrad of the convolution as a gather (no atomics): dx[b,iy,ix,ci] = sum over the (ky,kx) taps that hit an
// output position; multiplied by the ReLU mask of the layer input (relu_out > 0) when given.
__global__ void __launch_bounds__(256)
col2im_relu_kernel(const bf16* __restrict__ dcol, int B, int C, int H, int W, int k, int stride, int pad, int Ho, int Wo,
                   const bf16* __restrict__ relu_out, bf16* __restrict__ dx) {
  pdl_wait();
  pdl_trigger();
  const int K = C * k * k;
  const size_t total = (size_t)B * H * W * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % C);
    const size_t pix = i / C;
    const int ix = (int)(pix % W), iy = (int)((pix / W) % H), b = (int)(pix / ((size_t)W * H));
    float acc = 0.f;
    if (relu_out == nullptr || __bfloat162float(relu_out[i]) > 0.f) {
      for (int ky = 0; ky < k; ++ky) {
        const int ty = iy + pad - ky;
        if (ty < 0 || ty % stride != 0) continue;
        const int oy = ty / stride;
        if (oy >= Ho) continue;
        for (int kx = 0; kx < k; ++kx) {
          const int tx = ix + pad - kx;
          if (tx < 0 || tx % stride != 0) continue;
          const int ox = tx / stride;
          if (ox >= Wo) continue;
          acc += __bfloat162float(dcol[(((size_t)b * Ho + oy) * Wo + ox) * K + ci * k * k + ky * k + kx]);
        }
      }
    }
    dx[i] = __float2bfloat16(acc);
  }
}

// token finish of the conv-stem tokens: out[dst_row[r]] = x[src(b, tok - tok_base)] + add0[tok_class[tok]] + add1[tok]
//   r = b*ncols + jj ; tok = tok_idx ? tok_idx[b*idx_ld + col0 + jj] : tok_base + jj   (pretrain_models.py:202-216,256)
//   x holds the modality's sources (sensors) stacked along the batch: src(b, tl) = ((tl / n_per)*B + b)*n_per + tl % n_per
__global__ void __launch_bounds__(256)
token_finish_kernel(const bf16* __restrict__ x, int B, int n_per, const int32_t* __restrict__ tok_idx, int idx_ld, int col0,
                    int ncols, int tok_base, const float* __restrict__ add0, const int32_t* __restrict__ tok_class,
                    const float* __restrict__ add1, const int32_t* __restrict__ dst_row, bf16* __restrict__ out, int D) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int nchunk = D >> 3;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < B * ncols; r += gridDim.x * wpb) {
    const int b = r / ncols, jj = r - b * ncols;
    const int tok = tok_idx ? tok_idx[(size_t)b * idx_ld + col0 + jj] : tok_base + jj;
    const int tl = tok - tok_base, sidx = tl / n_per;
    const bf16* src = x + (((size_t)sidx * B + b) * n_per + (tl - sidx * n_per)) * D;
    const int dr = dst_row ? dst_row[r] : r;
    for (int ch = lane; ch < nchunk; ch += 32) {
      float v[8];
      load8<bf16>(src + ch * 8, v);
      if (add0) {
        float a[8];
        load8<float>(add0 + (size_t)tok_class[tok] * D + ch * 8, a);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += a[i];
      }
      if (add1) {
        float a[8];
        load8<float>(add1 + (size_t)tok * D + ch * 8, a);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += a[i];
      }
      store8(out + (size_t)dr * D + ch * 8, v);
    }
  }
}

// dst[b*n_total + tok_idx[b, j]] += src[b*ncols + j]  (bf16 rows; the tokens of one sample are distinct, so no two
// source rows hit the same destination row): adds the gradient of a row gather (the masked encoder's visible tokens)
// into the gradient of the full token sequence it was taken from (joint MAE + policy-feature step)
__global__ void __launch_bounds__(256)
row_scatter_add_kernel(const bf16* __restrict__ src, int B, int ncols, const int32_t* __restrict__ tok_idx, int idx_ld,
                       int n_total, int D, bf16* __restrict__ dst) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int nchunk = D >> 3;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < B * ncols; r += gridDim.x * wpb) {
    const int b = r / ncols, jj = r - b * ncols;
    const int tok = tok_idx[(size_t)b * idx_ld + jj];
    bf16* d = dst + ((size_t)b * n_total + tok) * D;
    for (int ch = lane; ch < nchunk; ch += 32) {
      float v[8], a[8];
      load8<bf16>(d + ch * 8, v);
      load8<bf16>(src + (size_t)r * D + ch * 8, a);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += a[i];
      store8(d + ch * 8, v);
    }
  }
}

// backward of the token gather: dtok[src(b, tl)] = slot >= 0 ? dx0[b*rows_per_sample + slot] : 0,
//   slot = slot_of_token ? slot_of_token[b*n_total + tok_base + tl] : tok_base + tl   (src() as above)
__global__ void __launch_bounds__(256)
token_finish_bwd_kernel(const bf16* __restrict__ dx0, int B, int rows_per_sample, int n_total,
                        const int32_t* __restrict__ slot_of_token, int tok_base, int n_mod, int n_per, int D,
                        bf16* __restrict__ dtok) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int nchunk = D >> 3;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < B * n_mod; r += gridDim.x * wpb) {
    const int b = r / n_mod, tl = r - b * n_mod;
    const int slot = slot_of_token ? slot_of_token[(size_t)b * n_total + tok_base + tl] : tok_base + tl;
    const int sidx = tl / n_per;
    bf16* dst = dtok + (((size_t)sidx * B + b) * n_per + (tl - sidx * n_per)) * D;
    for (int ch = lane; ch < nchunk; ch += 32) {
      uint4 u = make_uint4(0, 0, 0, 0);
      if (slot >= 0) u = *reinterpret_cast<const uint4*>(dx0 + ((size_t)b * rows_per_sample + slot) * D + ch * 8);
      *reinterpret_cast<uint4*>(dst + ch * 8) = u;
    }
  }
}

PatchSrc make_patch_src(const m3l_patch_source* s) {
  PatchSrc ps;
  for (int i = 0; i < 4; ++i) ps.src[i] = static_cast<const float*>(s->src[i]);
  ps.C = s->channels; ps.H = s->height; ps.W = s->width; ps.ph = s->patch_h; ps.pw = s->patch_w;
  ps.gw = s->width / s->patch_w;
  ps.n_per_src = (s->height / s->patch_h) * ps.gw;
  ps.tok_base = s->token_base;
  ps.P = s->patch_h * s->patch_w * s->channels;
  ps.layout = s->layout; ps.u8 = s->dtype == 1 ? 1 : 0;
  ps.sb = s->stride_b; ps.cg = s->chan_group > 0 ? s->chan_group : 1;
  ps.sf = s->stride_f; ps.sch = s->stride_ch; ps.sy = s->stride_y; ps.sx = s->stride_x;
  ps.lo = s->norm_lo; ps.span = s->norm_span;
  return ps;
}

int check_patch_src(const m3l_patch_source* s, const char* what) {
  M3L_REQUIRE(s->layout == 0 || s->layout == 1, "%s: patch source layout %d unknown", what, s->layout);
  if (s->layout == 1) {
    M3L_REQUIRE(s->dtype == 0 || s->dtype == 1, "%s: raw observation dtype %d unknown (0 fp32, 1 uint8)", what, s->dtype);
    M3L_REQUIRE(s->chan_group >= 1 && s->channels % s->chan_group == 0, "%s: channels %d not a multiple of the channel group %d",
                what, s->channels, s->chan_group);
    M3L_REQUIRE(s->norm_span != 0.f, "%s: normalisation span must be non-zero", what);
  } else {
    M3L_REQUIRE(s->dtype == 0, "%s: layout 0 takes fp32 maps", what);
  }
  return M3L_OK;
}

// vt_load as a kernel of its own (callers that need the maps materialised: the conv stem, reconstruct()):
// out[b, c, y, x] fp32 NCHW contiguous <- raw observation (layout 1 addressing), one thread per output element
__global__ void __launch_bounds__(256)
vt_load_kernel(PatchSrc ps, int sensor, long long total, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const long long plane = (long long)ps.H * ps.W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long bc = i / plane;
    const int yx = (int)(i - bc * plane);
    const int y = yx / ps.W, x = yx - y * ps.W;
    const int b = (int)(bc / ps.C), c = (int)(bc - (long long)b * ps.C);
    const int f = c / ps.cg, ch = c - f * ps.cg;
    const long long off = (long long)b * ps.sb + (long long)f * ps.sf + (long long)ch * ps.sch + (long long)y * ps.sy +
                          (long long)x * ps.sx;
    out[i] = ps.u8 ? raw_norm_u8(ps, reinterpret_cast<const unsigned char*>(src_of(ps, sensor))[off])
                   : raw_norm(ps, src_of(ps, sensor)[off]);
  }
}

int ln_grid(int M, int warps_per_block) {
  const int blocks = (M + warps_per_block - 1) / warps_per_block;
  const int cap = device_sm_count() * 8;
  return blocks < cap ? blocks : cap;
}

}  // namespace
}  // namespace m3l

using namespace m3l;

extern "C" int m3l_mask_indices(const float* noise, int batch, int n_total, const m3l_mask_segments* segs,
                                int64_t* masked, int64_t* unmasked, int32_t* slot_of_token,
                                int32_t* unmasked_i32, int32_t* masked_row_of_token, int n_masked_first,
                                void* stream) {
  M3L_REQUIRE(noise && segs && masked && unmasked, "mask_indices: null pointer");
  M3L_REQUIRE(segs->count >= 1 && segs->count <= M3L_MAX_SEGMENTS, "mask_indices: bad segment count %d", segs->count);
  if (batch == 0) return M3L_OK;
  int nm = 0, nu = 0, maxlen = 0;
  for (int i = 0; i < segs->count; ++i) {
    M3L_REQUIRE(segs->n_masked[i] >= 0 && segs->n_masked[i] <= segs->length[i] &&
                    segs->offset[i] >= 0 && segs->offset[i] + segs->length[i] <= n_total,
                "mask_indices: bad segment %d", i);
    nm += segs->n_masked[i];
    nu += segs->length[i] - segs->n_masked[i];
    maxlen = segs->length[i] > maxlen ? segs->length[i] : maxlen;
  }
  M3L_REQUIRE(maxlen <= 8192, "mask_indices: segment longer than 8192");
  dim3 grid(batch, segs->count);
  const int threads = maxlen <= 64 ? 64 : (maxlen <= 128 ? 128 : 256);
  M3L_CUDA(launch_kernel(mask_indices_kernel, dim3(grid), dim3(threads), maxlen * sizeof(float), (cudaStream_t)stream, 
      noise, n_total, *segs, masked, nm, unmasked, nu, slot_of_token, unmasked_i32, masked_row_of_token,
      n_masked_first, batch));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_patch_layernorm(const m3l_patch_source* src, int batch, const int64_t* tok_idx, int idx_ld,
                                   int col0, int ncols, const float* gamma, const float* beta, float eps,
                                   void* out_bf16, void* xhat_bf16, void* stream) {
  M3L_REQUIRE(src && gamma && beta && out_bf16, "patch_layernorm: null pointer");
  if (batch * ncols == 0) return M3L_OK;
  { const int s_ = check_patch_src(src, "patch_layernorm"); if (s_) return s_; }
  PatchSrc ps = make_patch_src(src);
  M3L_REQUIRE(ps.P * sizeof(float) <= 48 * 1024, "patch_layernorm: patch dim %d too large", ps.P);
  M3L_CUDA(launch_kernel(patch_ln_kernel, dim3(batch * ncols), dim3(128), ps.P * sizeof(float), (cudaStream_t)stream, 
      ps, tok_idx, idx_ld, col0, ncols, gamma, beta, (bf16*)out_bf16, (bf16*)xhat_bf16, eps));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_layernorm_fwd(const void* x, int x_fp32, int rows, int dim, const float* gamma,
                                 const float* beta, float eps, void* y_bf16, float* stats,
                                 const int32_t* dst_row, const float* add0, const int32_t* add0_row,
                                 const float* add1, const int32_t* add1_row, void* stream) {
  M3L_REQUIRE(x && gamma && beta && y_bf16, "layernorm_fwd: null pointer");
  M3L_REQUIRE(dim % 8 == 0 && dim <= 1024, "layernorm_fwd: dim %d unsupported", dim);
  if (rows == 0) return M3L_OK;
  const int wpb = 8;
  int grid = (rows + 2 * wpb - 1) / (2 * wpb);
  const int cap = device_sm_count() * 8;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  const int nch = dim <= 256 ? 1 : (dim <= 512 ? 2 : 4);
  cudaStream_t st = (cudaStream_t)stream;
#define M3L_LN_FWD(T, N)                                                                                   \
  M3L_CUDA(launch_kernel(layernorm_fwd_kernel<T, N>, dim3(grid), dim3(wpb * 32), 0, st, (const T*)x, rows, dim, gamma, beta, eps, (bf16*)y_bf16, \
                                                        stats, dst_row, add0, add0_row, add1, add1_row))
  if (!x_fp32 && nch <= 2) {
    // bf16 fast path: 4-deep warp-private cp.async ring; one resident wave (64 registers: four blocks per SM)
    const size_t ring = (size_t)wpb * 4 * nch * 512;
    grid = std::min(grid, device_sm_count() * 4);
    if (nch == 1)
      M3L_CUDA(launch_kernel(ln_fwd_pipe_kernel<1, 4>, dim3(grid), dim3(wpb * 32), ring, st, (const bf16*)x, rows, dim, gamma, beta, eps, (bf16*)y_bf16,
                                                            stats, dst_row, add0, add0_row, add1, add1_row));
    else
      M3L_CUDA(launch_kernel(ln_fwd_pipe_kernel<2, 4>, dim3(grid), dim3(wpb * 32), ring, st, (const bf16*)x, rows, dim, gamma, beta, eps, (bf16*)y_bf16,
                                                            stats, dst_row, add0, add0_row, add1, add1_row));
  } else if (x_fp32) {
    if (nch == 1) M3L_LN_FWD(float, 1); else if (nch == 2) M3L_LN_FWD(float, 2); else M3L_LN_FWD(float, 4);
  } else {
    if (nch == 1) M3L_LN_FWD(bf16, 1); else if (nch == 2) M3L_LN_FWD(bf16, 2); else M3L_LN_FWD(bf16, 4);
  }
#undef M3L_LN_FWD
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

template <typename TIn, typename TOut>
static int launch_ln_bwd(int nch, int grid, size_t smem, cudaStream_t st, const bf16* dy, const int32_t* src_row,
                         const TIn* x, const float* stats, int rows, int dim, const float* gamma, const bf16* skip,
                         TOut* dx, float* dgamma, float* dbeta, float* dx_colsum) {
  static bool configured = false;
  if (!configured) {   // dim 1024: 8 warps x 3 x 1024 floats = 96 KB of reduction scratch
    M3L_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<TIn, TOut, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    M3L_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<TIn, TOut, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured = true;
  }
  if (nch == 1)
    M3L_CUDA(launch_kernel(layernorm_bwd_kernel<TIn, TOut, 1>, dim3(grid), dim3(256), smem, st, dy, src_row, x, stats, rows, dim, gamma, skip, dx, dgamma, dbeta, dx_colsum));
  else if (nch == 2)
    M3L_CUDA(launch_kernel(layernorm_bwd_kernel<TIn, TOut, 2>, dim3(grid), dim3(256), smem, st, dy, src_row, x, stats, rows, dim, gamma, skip, dx, dgamma, dbeta, dx_colsum));
  else
    M3L_CUDA(launch_kernel(layernorm_bwd_kernel<TIn, TOut, 4>, dim3(grid), dim3(256), smem, st, dy, src_row, x, stats, rows, dim, gamma, skip, dx, dgamma, dbeta, dx_colsum));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_layernorm_bwd(const void* dy_bf16, const int32_t* src_row, const void* x, int x_fp32,
                                 const float* stats, int rows, int dim, const float* gamma,
                                 const void* skip_bf16, void* dx, int dx_fp32, float* dgamma, float* dbeta,
                                 float* dx_colsum, void* stream) {
  M3L_REQUIRE(dy_bf16 && x && stats && gamma && dx, "layernorm_bwd: null pointer");
  M3L_REQUIRE(dim % 8 == 0 && dim <= 1024, "layernorm_bwd: dim %d unsupported", dim);
  M3L_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "layernorm_bwd: dgamma/dbeta must both be set");
  if (rows == 0) return M3L_OK;
  const int wpb = 8;
  int grid = (rows + 2 * wpb - 1) / (2 * wpb);   // small problems: favour parallelism (2 rows per warp)
  const int cap = device_sm_count() * 2;    // bounds the atomics per gradient address (contended fp32 atomics ~10 ns each)
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  const size_t smem = (size_t)wpb * 3 * dim * sizeof(float);
  const int nch = dim <= 256 ? 1 : (dim <= 512 ? 2 : 4);
  cudaStream_t st = (cudaStream_t)stream;
  const bf16* dy = (const bf16*)dy_bf16;
  const bf16* skip = (const bf16*)skip_bf16;
  if (!x_fp32 && !dx_fp32 && nch <= 2) {
    constexpr int kSt = 4;

    const size_t ring = (size_t)wpb * kSt * 3 * nch * 512;
    const size_t total = ring + smem;
    static bool configured = false;
    if (!configured) {
      M3L_CUDA(cudaFuncSetAttribute(ln_bwd_pipe_kernel<1, kSt>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
      M3L_CUDA(cudaFuncSetAttribute(ln_bwd_pipe_kernel<2, kSt>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured = true;
    }
    if (nch == 1)
      M3L_CUDA(launch_kernel(ln_bwd_pipe_kernel<1, kSt>, dim3(grid), dim3(256), total, st, dy, src_row, (const bf16*)x, stats, rows, dim, gamma, skip, (bf16*)dx, dgamma, dbeta, dx_colsum));
    else
      M3L_CUDA(launch_kernel(ln_bwd_pipe_kernel<2, kSt>, dim3(grid), dim3(256), total, st, dy, src_row, (const bf16*)x, stats, rows, dim, gamma, skip, (bf16*)dx, dgamma, dbeta, dx_colsum));
    M3L_CUDA(cudaGetLastError());
    return M3L_OK;
  }
  if (x_fp32 && !dx_fp32)
    return launch_ln_bwd<float, bf16>(nch, grid, smem, st, dy, src_row, (const float*)x, stats, rows, dim, gamma, skip, (bf16*)dx, dgamma, dbeta, dx_colsum);
  if (!x_fp32 && !dx_fp32)
    return launch_ln_bwd<bf16, bf16>(nch, grid, smem, st, dy, src_row, (const bf16*)x, stats, rows, dim, gamma, skip, (bf16*)dx, dgamma, dbeta, dx_colsum);
  if (x_fp32 && dx_fp32)
    return launch_ln_bwd<float, float>(nch, grid, smem, st, dy, src_row, (const float*)x, stats, rows, dim, gamma, skip, (float*)dx, dgamma, dbeta, dx_colsum);
  return launch_ln_bwd<bf16, float>(nch, grid, smem, st, dy, src_row, (const bf16*)x, stats, rows, dim, gamma, skip, (float*)dx, dgamma, dbeta, dx_colsum);
}

extern "C" int m3l_decoder_assemble_fwd(const void* d_bf16, int n_visible, const float* mask_token,
                                        const int32_t* slot_of_token, int batch, int n_tokens, int dim,
                                        const float* add0, const int32_t* tok_class, const float* add1,
                                        void* z_bf16, void* stream) {
  M3L_REQUIRE(d_bf16 && mask_token && slot_of_token && z_bf16, "decoder_assemble_fwd: null pointer");
  M3L_REQUIRE(dim % 8 == 0, "decoder_assemble_fwd: dim %d not a multiple of 8", dim);
  M3L_REQUIRE(add0 == nullptr || tok_class != nullptr, "decoder_assemble_fwd: add0 needs tok_class");
  if (batch * n_tokens == 0) return M3L_OK;
  const int wpb = 8;
  // one resident wave (40 registers: six blocks of 256 threads per SM): 18.9 -> 16.1 us
  M3L_CUDA(launch_kernel(assemble_fwd_kernel, dim3(std::min(ln_grid(batch * n_tokens, wpb), device_sm_count() * 6)), dim3(wpb * 32), 0, (cudaStream_t)stream, 
      (const bf16*)d_bf16, n_visible, mask_token, slot_of_token, batch, n_tokens, dim, add0, tok_class, add1,
      (bf16*)z_bf16));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_decoder_assemble_bwd(const void* dz_bf16, const int32_t* slot_of_token, int batch,
                                        int n_tokens, int dim, int n_visible, void* dd_bf16,
                                        float* dmask_token, float* dadd0, const int32_t* tok_class,
                                        int n_classes, float* dadd1, void* stream) {
  M3L_REQUIRE(dz_bf16 && slot_of_token && dd_bf16, "decoder_assemble_bwd: null pointer");
  M3L_REQUIRE(dadd0 == nullptr || tok_class != nullptr, "decoder_assemble_bwd: dadd0 needs tok_class");
  M3L_REQUIRE(dim % 8 == 0 && dim <= 1024, "decoder_assemble_bwd: dim %d unsupported", dim);
  M3L_REQUIRE(n_classes >= 0 && n_classes <= kAsmClasses, "decoder_assemble_bwd: at most %d token classes", kAsmClasses);
  if (batch * n_tokens == 0) return M3L_OK;
  const int rows = batch * n_tokens;
  const int wpb = 8;
  int grid = (rows + 4 * wpb - 1) / (4 * wpb);
  if (grid > device_sm_count() * 2) grid = device_sm_count() * 2;
  const int nvec = (1 + kAsmClasses) * dim;
  const size_t smem = (size_t)wpb * nvec * sizeof(float);
  const int nch = dim <= 256 ? 1 : (dim <= 512 ? 2 : 4);
  cudaStream_t st = (cudaStream_t)stream;
  static bool configured = false;
  if (!configured) {
    M3L_CUDA(cudaFuncSetAttribute(assemble_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024));
    M3L_CUDA(cudaFuncSetAttribute(assemble_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024));
    M3L_CUDA(cudaFuncSetAttribute(assemble_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024));
    configured = true;
  }
#define M3L_ASM_BWD(N)                                                                                          \
  M3L_CUDA(launch_kernel(assemble_bwd_kernel<N>, dim3(grid), dim3(wpb * 32), smem, st, (const bf16*)dz_bf16, slot_of_token, batch, n_tokens, dim, \
                                                       n_visible, (bf16*)dd_bf16, dmask_token, dadd0, tok_class, \
                                                       n_classes, dadd1))
  if (nch == 1) M3L_ASM_BWD(1); else if (nch == 2) M3L_ASM_BWD(2); else M3L_ASM_BWD(4);
#undef M3L_ASM_BWD
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_rowclass_sum(const void* dx_bf16, int batch, int n_visible, int dim,
                                const int32_t* slot_class, float* dclass, const int32_t* row_pos, float* dpos,
                                void* stream) {
  M3L_REQUIRE(dx_bf16, "rowclass_sum: null pointer");
  if (batch * n_visible == 0) return M3L_OK;
  M3L_CUDA(launch_kernel(rowclass_sum_kernel, dim3(n_visible, batch >= 64 ? 16 : 1), dim3(256), 0, (cudaStream_t)stream, (const bf16*)dx_bf16, batch, n_visible, dim,
                                                                   slot_class, dclass, row_pos, dpos));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_mse_loss(const m3l_patch_source* src, int batch, const int64_t* tok_idx, int idx_ld, int col0,
                            int ncols, const float* pred, float weight, void* dpred_bf16, float* loss_acc,
                            float* dpred_colsum, void* workspace, size_t workspace_bytes, void* stream) {
  M3L_REQUIRE(src && pred && dpred_bf16 && loss_acc, "mse_loss: null pointer");
  if (batch * ncols == 0) return M3L_OK;
  { const int s_ = check_patch_src(src, "mse_loss"); if (s_) return s_; }
  PatchSrc ps = make_patch_src(src);
  const int rowlen = ps.pw * ps.C;
  M3L_REQUIRE(ps.P % 4 == 0 && rowlen % 4 == 0, "mse_loss: patch dim %d / row %d must be multiples of 4", ps.P, rowlen);
  M3L_REQUIRE(dpred_colsum == nullptr || ps.P <= kMseMaxT * 128, "mse_loss: fused column sums need patch dim <= %d", kMseMaxT * 128);
  // pitch = rowlen + pad with pitch % 32 == round_up(C, 4) % 32 (rowlen % 4 == 0, so pitch % 4 == 0 for
  // the float4 reads): the lane groups of consecutive p1 rows land on disjoint banks
  const int want = ((ps.C + 3) & ~3) % 32;
  const int pad = ((want - rowlen) % 32 + 32) % 32;
  const int pitch = rowlen + pad;
  const size_t smem = ((size_t)8 * ps.ph * pitch + ps.P) * sizeof(float);
  M3L_REQUIRE(smem <= 160 * 1024, "mse_loss: patch dim %d unsupported", ps.P);
  const int rows = batch * ncols;
  int grid = (rows + 15) / 16;                       // two rows per warp: the per-row gather chain is latency bound
  // one resident wave (126 registers: two blocks of 256 threads per SM): measured 53.8 -> 43.0 us against 4-8 blocks
  // per SM running as 2-3 waves (each block pays its start-up again); also bounds the atomics per bias-gradient address
  const int cap = device_sm_count() * 2;
  if (grid > cap) grid = cap;
  M3L_REQUIRE(workspace != nullptr && workspace_bytes >= 256 + (size_t)grid * sizeof(float),
              "mse_loss: workspace too small (%zu bytes)", workspace_bytes);
  // which gather (see the kernel): the specialised forms need every run 16-byte (fp32) / 4-byte (uint8) aligned
  static const bool fast_on = [] { const char* e = getenv("M3L_MSE_FAST"); return !(e && e[0] == '0'); }();
  auto al = [](const void* q, uintptr_t a) { return (reinterpret_cast<uintptr_t>(q) & (a - 1)) == 0; };
  bool ptrs16 = true, ptrs4 = true;
  for (int i = 0; i < 4; ++i)
    if (ps.src[i] != nullptr) { ptrs16 = ptrs16 && al(ps.src[i], 16); ptrs4 = ptrs4 && al(ps.src[i], 4); }
  int mode = 0;
  const int segs = (ps.C * ps.ph + 31) / 32;
  if (fast_on && !ps.u8 && ps.pw % 4 == 0 && ptrs16 && segs <= 3 && (ps.pw == 8 || ps.pw == 4)) {
    if (ps.layout == 0 && ps.W % 4 == 0 && ((long long)ps.H * ps.W) % 4 == 0) mode = ps.pw == 8 ? 18 : 14;
    if (ps.layout == 1 && ps.sx == 1 && ps.sb % 4 == 0 && ps.sf % 4 == 0 && ps.sch % 4 == 0 && ps.sy % 4 == 0)
      mode = ps.pw == 8 ? 18 : 14;
  }
  if (fast_on && ps.layout == 1 && ps.sch == 1 && ps.sx == ps.cg && ps.pw * ps.cg == 24 && (ps.C / ps.cg) * ps.ph <= 32 &&
      ps.sb % 4 == 0 && ps.sf % 4 == 0 && ps.sy % 4 == 0)
    mode = ps.u8 ? (ptrs4 ? 3 : 0) : (ptrs16 ? 2 : 0);
  if (mode == 14 && ps.P <= 256 && segs <= 2) mode = 15;
  if (mode == 15) {
    // 64 registers: four blocks per SM, still ONE resident wave
    grid = (rows + 15) / 16;
    if (grid > device_sm_count() * 4) grid = device_sm_count() * 4;
    M3L_REQUIRE(workspace_bytes >= 256 + (size_t)grid * sizeof(float), "mse_loss: workspace too small (%zu bytes)",
                workspace_bytes);
  }
  static bool configured[6] = {false, false, false, false, false, false};     // per kernel variant
  auto go = [&](auto kern, int slot) -> int {
    if (!configured[slot]) {
      M3L_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      configured[slot] = true;
    }
    M3L_CUDA(launch_kernel(kern, dim3(grid), dim3(256), smem, (cudaStream_t)stream, ps, tok_idx, idx_ld, col0, ncols, rows,
                           pred, weight, (bf16*)dpred_bf16, loss_acc, dpred_colsum, pitch, workspace));
    M3L_CUDA(cudaGetLastError());
    return M3L_OK;
  };
  switch (mode) {
    case 18: return go(mse_loss_kernel<1, 3, 2, 8, 2>, 1);    // planes, runs of 8 floats (64 x 64 x 12 frames, patch 8)
    case 14: return go(mse_loss_kernel<1, 3, 1, 8, 2>, 2);    // planes, runs of 4 floats, any patch dim
    case 15: return go(mse_loss_kernel<1, 2, 1, 2, 4>, 5);    // ... patch dim <= 256 (32 x 32 x 12 tactile maps, patch 4)
    case 2: return go(mse_loss_kernel<2, 1, 6, 8, 2>, 3);     // raw fp32 frames [B, F, H, W, 3], patch 8
    case 3: return go(mse_loss_kernel<3, 1, 6, 8, 2>, 4);     // raw uint8 frames
    default: return go(mse_loss_kernel<0, 1, 1, 8, 2>, 0);
  }
}

extern "C" int m3l_colsum(const void* x_bf16, int rows, int cols, int ld, float* out, void* stream) {
  M3L_REQUIRE(x_bf16 && out, "colsum: null pointer");
  M3L_REQUIRE(cols % 8 == 0 && ld % 8 == 0, "colsum: cols/ld must be multiples of 8");
  if (rows == 0) return M3L_OK;
  const int gx = (cols / 8 + 31) / 32;
  int gy = (device_sm_count() * 2 + gx - 1) / gx;   // <= ~2 blocks per SM: few atomics per output address
  if (gy > (rows + 63) / 64) gy = (rows + 63) / 64;
  if (gy < 1) gy = 1;
  M3L_CUDA(launch_kernel(colsum_kernel, dim3(dim3(gx, gy)), dim3(dim3(32, 8)), 0, (cudaStream_t)stream, (const bf16*)x_bf16, rows, cols, ld, out));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_ln_param_grad(const void* da_bf16, const void* xhat_bf16, int rows, int dim, float* dgamma,
                                 float* dbeta, void* stream) {
  M3L_REQUIRE(da_bf16 && xhat_bf16 && dgamma && dbeta, "ln_param_grad: null pointer");
  M3L_REQUIRE(dim % 8 == 0, "ln_param_grad: dim %d must be a multiple of 8", dim);
  if (rows == 0) return M3L_OK;
  const int gx = (dim / 8 + 31) / 32;
  int gy = (device_sm_count() * 2 + gx - 1) / gx;
  if (gy > (rows + 31) / 32) gy = (rows + 31) / 32;
  if (gy < 1) gy = 1;
  M3L_CUDA(launch_kernel(ln_param_grad_kernel, dim3(gx, gy), dim3(32, 8), 0, (cudaStream_t)stream,
                         (const bf16*)da_bf16, (const bf16*)xhat_bf16, rows, dim, dgamma, dbeta));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_token_mean_fwd(const void* x_bf16, int batch, int n_tokens, int dim, float* out, void* stream) {
  M3L_REQUIRE(x_bf16 && out, "token_mean_fwd: null pointer");
  M3L_REQUIRE(dim % 8 == 0 && n_tokens > 0, "token_mean_fwd: dim %d must be a multiple of 8", dim);
  if (batch == 0) return M3L_OK;
  const int warps = batch * ((dim + 255) / 256);
  M3L_CUDA(launch_kernel(token_mean_fwd_kernel, dim3((warps + 7) / 8), dim3(256), 0, (cudaStream_t)stream,
                         (const bf16*)x_bf16, batch, n_tokens, dim, out));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_token_mean_bwd(const float* dout, int batch, int n_tokens, int dim, void* dx_bf16, void* stream) {
  M3L_REQUIRE(dout && dx_bf16, "token_mean_bwd: null pointer");
  M3L_REQUIRE(dim % 8 == 0 && n_tokens > 0, "token_mean_bwd: dim %d must be a multiple of 8", dim);
  if (batch == 0) return M3L_OK;
  const size_t total = (size_t)batch * n_tokens * (dim / 8);
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)device_sm_count() * 8) blocks = (size_t)device_sm_count() * 8;
  M3L_CUDA(launch_kernel(token_mean_bwd_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, dout, batch,
                         n_tokens, dim, (bf16*)dx_bf16));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_im2col(const void* x, int x_nhwc_bf16, int batch, int channels, int height, int width, int k,
                          int stride, int pad, void* col_bf16, void* stream) {
  M3L_REQUIRE(x && col_bf16, "im2col: null pointer");
  M3L_REQUIRE(k >= 1 && stride >= 1 && pad >= 0 && (channels * k * k) % 8 == 0,
              "im2col: channels*k*k = %d must be a multiple of 8", channels * k * k);
  const int Ho = (height + 2 * pad - k) / stride + 1, Wo = (width + 2 * pad - k) / stride + 1;
  if (batch == 0) return M3L_OK;
  const size_t total = (size_t)batch * Ho * Wo * (channels * k * k / 2);
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)device_sm_count() * 16) blocks = (size_t)device_sm_count() * 16;
  if (x_nhwc_bf16)
    M3L_CUDA(launch_kernel(im2col_kernel<true>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, x, batch, channels,
                           height, width, k, stride, pad, Ho, Wo, (bf16*)col_bf16));
  else
    M3L_CUDA(launch_kernel(im2col_kernel<false>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, x, batch, channels,
                           height, width, k, stride, pad, Ho, Wo, (bf16*)col_bf16));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_col2im_relu(const void* dcol_bf16, int batch, int channels, int height, int width, int k, int stride,
                               int pad, const void* relu_out_bf16, void* dx_bf16, void* stream) {
  M3L_REQUIRE(dcol_bf16 && dx_bf16, "col2im_relu: null pointer");
  M3L_REQUIRE(k >= 1 && stride >= 1 && pad >= 0, "col2im_relu: bad geometry");
  const int Ho = (height + 2 * pad - k) / stride + 1, Wo = (width + 2 * pad - k) / stride + 1;
  if (batch == 0) return M3L_OK;
  const size_t total = (size_t)batch * height * width * channels;
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)device_sm_count() * 16) blocks = (size_t)device_sm_count() * 16;
  M3L_CUDA(launch_kernel(col2im_relu_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const bf16*)dcol_bf16,
                         batch, channels, height, width, k, stride, pad, Ho, Wo, (const bf16*)relu_out_bf16, (bf16*)dx_bf16));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_token_finish(const void* x_bf16, int batch, int n_per, const int32_t* tok_idx, int idx_ld, int col0,
                                int ncols, int tok_base, const float* add0, const int32_t* tok_class, const float* add1,
                                const int32_t* dst_row, void* out_bf16, int dim, void* stream) {
  M3L_REQUIRE(x_bf16 && out_bf16, "token_finish: null pointer");
  M3L_REQUIRE(dim % 8 == 0, "token_finish: dim %d must be a multiple of 8", dim);
  M3L_REQUIRE(add0 == nullptr || tok_class != nullptr, "token_finish: add0 needs tok_class");
  M3L_REQUIRE(n_per > 0, "token_finish: n_per must be positive");
  if (batch * ncols == 0) return M3L_OK;
  M3L_CUDA(launch_kernel(token_finish_kernel, dim3(ln_grid(batch * ncols, 8)), dim3(256), 0, (cudaStream_t)stream,
                         (const bf16*)x_bf16, batch, n_per, tok_idx, idx_ld, col0, ncols, tok_base, add0, tok_class, add1,
                         dst_row, (bf16*)out_bf16, dim));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_token_finish_bwd(const void* dx0_bf16, int batch, int rows_per_sample, int n_total,
                                    const int32_t* slot_of_token, int tok_base, int n_mod, int n_per, int dim,
                                    void* dtok_bf16, void* stream) {
  M3L_REQUIRE(dx0_bf16 && dtok_bf16, "token_finish_bwd: null pointer");
  M3L_REQUIRE(dim % 8 == 0, "token_finish_bwd: dim %d must be a multiple of 8", dim);
  M3L_REQUIRE(n_per > 0 && n_mod % n_per == 0, "token_finish_bwd: n_mod %d must be a multiple of n_per %d", n_mod, n_per);
  if (batch * n_mod == 0) return M3L_OK;
  M3L_CUDA(launch_kernel(token_finish_bwd_kernel, dim3(ln_grid(batch * n_mod, 8)), dim3(256), 0, (cudaStream_t)stream,
                         (const bf16*)dx0_bf16, batch, rows_per_sample, n_total, slot_of_token, tok_base, n_mod, n_per, dim,
                         (bf16*)dtok_bf16));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_vt_load(const m3l_patch_source* src, int batch, int sensor, float* out_nchw, void* stream) {
  M3L_REQUIRE(src && out_nchw, "vt_load: null pointer");
  M3L_REQUIRE(src->layout == 1, "vt_load: the source must be a raw observation (layout 1)");
  M3L_REQUIRE(sensor >= 0 && sensor < 4 && src->src[sensor] != nullptr, "vt_load: bad sensor index %d", sensor);
  { const int s_ = check_patch_src(src, "vt_load"); if (s_) return s_; }
  if (batch == 0) return M3L_OK;
  PatchSrc ps = make_patch_src(src);
  const long long total = (long long)batch * ps.C * ps.H * ps.W;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  M3L_CUDA(launch_kernel(vt_load_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, ps, sensor, total, out_nchw));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_patchify(const m3l_patch_source* src, int batch, const int64_t* tok_idx, int idx_ld, int col0, int ncols,
                            void* out_bf16, int ld, void* stream) {
  M3L_REQUIRE(src && out_bf16, "patchify: null pointer");
  if (batch * ncols == 0) return M3L_OK;
  { const int s_ = check_patch_src(src, "patchify"); if (s_) return s_; }
  PatchSrc ps = make_patch_src(src);
  M3L_REQUIRE(ld >= ps.P && ld % 8 == 0, "patchify: row pitch %d must be a multiple of 8 and >= the patch dim %d", ld, ps.P);
  M3L_REQUIRE(ps.P * sizeof(float) <= 48 * 1024, "patchify: patch dim %d too large", ps.P);
  M3L_CUDA(launch_kernel(patchify_kernel, dim3(batch * ncols), dim3(128), ps.P * sizeof(float), (cudaStream_t)stream,
                         ps, tok_idx, idx_ld, col0, ncols, (bf16*)out_bf16, ld));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}

extern "C" int m3l_row_scatter_add(const void* src_bf16, int batch, int ncols, const int32_t* tok_idx, int idx_ld, int n_total,
                                   int dim, void* dst_bf16, void* stream) {
  M3L_REQUIRE(src_bf16 && tok_idx && dst_bf16, "row_scatter_add: null pointer");
  M3L_REQUIRE(dim % 8 == 0, "row_scatter_add: dim %d must be a multiple of 8", dim);
  if (batch * ncols == 0) return M3L_OK;
  M3L_CUDA(launch_kernel(row_scatter_add_kernel, dim3(ln_grid(batch * ncols, 8)), dim3(256), 0, (cudaStream_t)stream,
                         (const bf16*)src_bf16, batch, ncols, tok_idx, idx_ld, n_total, dim, (bf16*)dst_bf16));
  M3L_CUDA(cudaGetLastError());
  return M3L_OK;
}
