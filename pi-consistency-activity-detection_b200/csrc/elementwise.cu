// Bandwidth-bound kernels of the step: layout conversion, train-mode BatchNorm (+ReLU, + concat slot),
// same-padding max-pool, Dropout3d channel scale, activation backward + bias gradient, the 27-tap
// 'smooth' stencil.  Activations are channels-last bf16 views (ptr, row_stride, c_off); every thread
// moves 16-byte (8-channel) vectors; reductions go warp/registers -> shared -> one fp32 atomic per
// block and channel.
#include "common.cuh"
#include <stdlib.h>
#include "../../include/b200caps.h"

long long b2c_launches_add(long long n);

namespace {

constexpr int kBlock = 256;

struct RowMap {
  int cv;        // this thread's channel-vector index
  int rlane;     // this thread's row lane inside the block
  int rpb;       // rows per block iteration
  bool active;
};
// block threads are arranged as [rpb][CV]; threads beyond rpb*CV idle
__device__ __forceinline__ RowMap row_map(int CV) {
  RowMap m;
  m.rpb = kBlock / CV;
  if (m.rpb < 1) m.rpb = 1;
  m.cv = threadIdx.x % CV;
  m.rlane = threadIdx.x / CV;
  m.active = m.rlane < m.rpb;
  return m;
}

__device__ __forceinline__ uint4 ld16(const bf16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void st16(bf16* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

// Activations are stored as T = bf16 (default) or T = float (B2C precision mode 1: fp32 activations, tf32 GEMM operands).
// Every kernel moves 8-channel vectors: 16 bytes of bf16 or 32 bytes of fp32.  `round` (fp32 only): the value is the
// operand of a later tcgen05 kind::tf32 GEMM and is rounded to tf32 (nearest, ties away) here -- the tensor core would
// otherwise truncate the low 13 mantissa bits, a systematic -2.4e-4 relative bias per operand.
template <typename T>
__device__ __forceinline__ void ld8(const T* p, float* v);
template <>
__device__ __forceinline__ void ld8<bf16>(const bf16* p, float* v) { unpack8(ld16(p), v); }
template <>
__device__ __forceinline__ void ld8<float>(const float* p, float* v) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <typename T>
__device__ __forceinline__ void st8(T* p, const float* v, bool round);
template <>
__device__ __forceinline__ void st8<bf16>(bf16* p, const float* v, bool) { st16(p, pack8(v)); }
template <>
__device__ __forceinline__ void st8<float>(float* p, const float* v, bool round) {
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = round ? tf32_rna(v[j]) : v[j];
  reinterpret_cast<float4*>(p)[0] = make_float4(r[0], r[1], r[2], r[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(r[4], r[5], r[6], r[7]);
}

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void ncdhw_to_ndhwc_kernel(const float* __restrict__ in, T* __restrict__ out, int N, int C, long long THW,
                                      int Cpad) {
  const long long total = (long long)N * THW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / THW, pos = i - n * THW;
    const float* src = in + n * C * THW + pos;
    T* dst = out + i * Cpad;
    for (int c0 = 0; c0 < Cpad; c0 += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c0 + j < C) ? src[(long long)(c0 + j) * THW] : 0.f;
      st8(dst + c0, v, true);
    }
  }
}

// uint8 clips as the dataloader decodes them (ucf_dataloader.py:162-185: img / 255., aug = the clip mirrored in W) ->
// channels-last activations of BOTH forward passes: out[n] = u8 / 255, out[P + n] = the same clip mirrored in W.
// One thread per (n, t*h row, w): 3 byte loads (coalesced across w), two 8-channel vector stores.
template <typename T>
__global__ void u8_clip_to_cl_kernel(const uint8_t* __restrict__ in, T* __restrict__ out, int P, int C, long long rows, int W) {
  const long long total = (long long)P * rows * W;
  const long long plane = rows * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const long long r = i / W;                 // n * rows + row
    const long long n = r / rows, row = r - n * rows;
    const uint8_t* src = in + n * C * plane + row * W + w;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = j < C ? (float)src[(long long)j * plane] / 255.f : 0.f;
    st8(out + i * 8, v, true);
    st8(out + (((long long)(P + n) * rows + row) * W + (W - 1 - w)) * 8, v, true);
  }
}

__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, long long n, float scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (float)in[i] * scale;
}

template <typename T>
__global__ void ndhwc_to_ncdhw_kernel(const T* __restrict__ in, long long in_row_stride, int in_c_off, float* __restrict__ out,
                                      int N, int C, long long THW) {
  __shared__ float tile[32][33];
  const long long n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long p = p0 + j;
    const int c = c0 + threadIdx.x;
    float v = 0.f;
    if (p < THW && c < C) v = (float)in[(n * THW + p) * in_row_stride + in_c_off + c];
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j;
    const long long p = p0 + threadIdx.x;
    if (p < THW && c < C) out[(n * C + c) * THW + p] = tile[threadIdx.x][j];
  }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm statistics: grid (blocks, groups)
// finalize arguments of bn_stats_kernel (mean == nullptr: statistics only, a bn_finalize launch follows)
struct BnFinalize {
  float* mean;
  float* rstd;
  float* running_mean;
  float* running_var;
  float momentum, eps;
  int groups;
};
__device__ unsigned g_bn_blocks_done = 0;   // blocks of the running bn_stats launch that have published their sums

template <typename T>
__global__ void __launch_bounds__(kBlock) bn_stats_kernel(const T* __restrict__ x, long long rows_per_group, int C,
                                                          long long row_stride, int c_off, float* __restrict__ ws, BnFinalize F) {
  const int CV = C / 8;
  const RowMap m = row_map(CV);
  const int g = blockIdx.y;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  if (m.active) {
    const T* base = x + (long long)g * rows_per_group * row_stride + c_off + m.cv * 8;
#pragma unroll 4
    for (long long r = (long long)blockIdx.x * m.rpb + m.rlane; r < rows_per_group; r += (long long)gridDim.x * m.rpb) {
      float v[8];
      ld8(base + r * row_stride, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += v[j];
        s2[j] += v[j] * v[j];
      }
    }
  }
  // block reduce over row lanes through shared memory in a FIXED order (per-block result is deterministic)
  extern __shared__ float sh[];  // [rpb][2][C]
  if (m.active) {
    float* mine = sh + (size_t)m.rlane * 2 * C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mine[m.cv * 8 + j] = s1[j];
      mine[C + m.cv * 8 + j] = s2[j];
    }
  }
  __syncthreads();
  float* w = ws + (long long)g * 2 * C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float t = 0.f;
    for (int r = 0; r < m.rpb; ++r) t += sh[(size_t)r * 2 * C + i];
    atomicAdd(&w[i], t);
  }
  if (F.mean == nullptr) return;
  // The block that publishes last turns the sums into mean / rstd / running statistics (bn_finalize_kernel's arithmetic):
  // one launch less per layer.  These launches run one at a time on the step's stream, so one counter serves them all.
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(&g_bn_blocks_done, 1u);
    s_last = done == gridDim.x * gridDim.y - 1;
    if (s_last) g_bn_blocks_done = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const double M = (double)rows_per_group;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float rm = F.running_mean ? F.running_mean[c] : 0.f, rv = F.running_var ? F.running_var[c] : 0.f;
    for (int gg = 0; gg < F.groups; ++gg) {
      const double s1 = __ldcg(ws + (long long)gg * 2 * C + c), s2 = __ldcg(ws + (long long)gg * 2 * C + C + c);
      const double mu = s1 / M;
      double var = s2 / M - mu * mu;
      if (var < 0) var = 0;
      F.mean[gg * C + c] = (float)mu;
      F.rstd[gg * C + c] = (float)(1.0 / sqrt(var + (double)F.eps));
      const double unb = rows_per_group > 1 ? var * M / (M - 1.0) : var;
      rm = (1.f - F.momentum) * rm + F.momentum * (float)mu;
      rv = (1.f - F.momentum) * rv + F.momentum * (float)unb;
    }
    if (F.running_mean) F.running_mean[c] = rm;
    if (F.running_var) F.running_var[c] = rv;
  }
}

__global__ void bn_finalize_kernel(const float* __restrict__ ws_all, int ws_C, int c_off, long long rows_per_group, int C,
                                   int groups, float* __restrict__ mean, float* __restrict__ rstd, float* running_mean,
                                   float* running_var, float momentum, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float* ws = ws_all + c_off;
  const int Cw = ws_C;
  const double M = (double)rows_per_group;
  float rm = running_mean ? running_mean[c] : 0.f, rv = running_var ? running_var[c] : 0.f;
  for (int g = 0; g < groups; ++g) {
    const double s1 = ws[(long long)g * 2 * Cw + c], s2 = ws[(long long)g * 2 * Cw + Cw + c];
    const double mu = s1 / M;
    double var = s2 / M - mu * mu;
    if (var < 0) var = 0;
    mean[g * C + c] = (float)mu;
    rstd[g * C + c] = (float)(1.0 / sqrt(var + (double)eps));
    const double unb = rows_per_group > 1 ? var * M / (M - 1.0) : var;
    rm = (1.f - momentum) * rm + momentum * (float)mu;
    rv = (1.f - momentum) * rv + momentum * (float)unb;
  }
  if (running_mean) running_mean[c] = rm;
  if (running_var) running_var[c] = rv;
}

// y = fma(x, sc, sf) with sc = rstd * gamma, sf = fma(-mean, sc, beta): ONE definition, because the backward kernels
// recompute the ReLU mask (y > 0) from x with it instead of reading y back (a quarter / a third of their traffic)
__device__ __forceinline__ void bn_scale_shift(float mean, float rstd, float gamma, float beta, float& sc, float& sf) {
  sc = rstd * gamma;
  sf = fmaf(-mean, sc, beta);
}

template <typename T>
__global__ void __launch_bounds__(kBlock) bn_relu_apply_kernel(const T* __restrict__ x, long long rows_per_group, int C,
                                                               long long x_rs, int x_co, const float* __restrict__ mean,
                                                               const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, T* __restrict__ y, long long y_rs,
                                                               int y_co, int relu) {
  const int CV = C / 8;
  const RowMap m = row_map(CV);
  if (!m.active) return;
  const int g = blockIdx.y;
  float sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = m.cv * 8 + j;
    bn_scale_shift(mean[g * C + c], rstd[g * C + c], gamma[c], beta[c], sc[j], sf[j]);
  }
  const long long row0 = (long long)g * rows_per_group;
#pragma unroll 4
  for (long long r = (long long)blockIdx.x * m.rpb + m.rlane; r < rows_per_group; r += (long long)gridDim.x * m.rpb) {
    float v[8];
    ld8(x + (row0 + r) * x_rs + x_co + m.cv * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = fmaf(v[j], sc[j], sf[j]);
      if (relu) v[j] = fmaxf(v[j], 0.f);
    }
    st8(y + (row0 + r) * y_rs + y_co + m.cv * 8, v, true);          // next conv's operand
  }
}

template <typename T, bool kRemask>
__global__ void __launch_bounds__(kBlock, kRemask ? 3 : 4) bn_bwd_reduce_kernel(const T* __restrict__ dy, long long dy_rs, int dy_co,
                                                               const T* __restrict__ y, long long y_rs, int y_co,
                                                               const T* __restrict__ x, long long x_rs, int x_co,
                                                               long long rows_per_group, int C, const float* __restrict__ mean,
                                                               const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float* __restrict__ ws, int relu) {
  const int CV = C / 8;
  const RowMap m = row_map(CV);
  const int g = blockIdx.y;
  constexpr bool remask = kRemask;              // ReLU mask recomputed from x (see bn_scale_shift); relu is set then
  float s1[8], s2[8], mu[8], rs[8], sc[kRemask ? 8 : 1], sf[kRemask ? 8 : 1];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  if (m.active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mu[j] = mean[g * C + m.cv * 8 + j];
      rs[j] = rstd[g * C + m.cv * 8 + j];
      if constexpr (kRemask) bn_scale_shift(mu[j], rs[j], gamma[m.cv * 8 + j], beta[m.cv * 8 + j], sc[j], sf[j]);
    }
    const long long row0 = (long long)g * rows_per_group;
#pragma unroll 2
    for (long long r = (long long)blockIdx.x * m.rpb + m.rlane; r < rows_per_group; r += (long long)gridDim.x * m.rpb) {
      float d[8], yy[8], xx[8];
      ld8(dy + (row0 + r) * dy_rs + dy_co + m.cv * 8, d);
      ld8(x + (row0 + r) * x_rs + x_co + m.cv * 8, xx);
      if (relu && !remask) ld8(y + (row0 + r) * y_rs + y_co + m.cv * 8, yy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if constexpr (kRemask) yy[j] = fmaf(xx[j], sc[j], sf[j]);
        const float dr = (!relu || yy[j] > 0.f) ? d[j] : 0.f;
        s1[j] += dr;
        s2[j] += dr * (xx[j] - mu[j]) * rs[j];
      }
    }
  }
  extern __shared__ float sh[];  // [rpb][2][C], fixed-order reduction
  if (m.active) {
    float* mine = sh + (size_t)m.rlane * 2 * C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mine[m.cv * 8 + j] = s1[j];
      mine[C + m.cv * 8 + j] = s2[j];
    }
  }
  __syncthreads();
  float* w = ws + (long long)g * 2 * C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float t = 0.f;
    for (int r = 0; r < m.rpb; ++r) t += sh[(size_t)r * 2 * C + i];
    atomicAdd(&w[i], t);
  }
}

template <typename T, bool kRemask>
__global__ void __launch_bounds__(kBlock, 3) bn_bwd_apply_kernel(const T* __restrict__ dy, long long dy_rs, int dy_co,
                                                              const T* __restrict__ y, long long y_rs, int y_co,
                                                              const T* __restrict__ x, long long x_rs, int x_co,
                                                              long long rows_per_group, int C, int groups,
                                                              const float* __restrict__ mean, const float* __restrict__ rstd,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ ws, T* __restrict__ dx, long long dx_rs,
                                                              int dx_co, float* dgamma, float* dbeta, int relu) {
  const int CV = C / 8;
  const RowMap m = row_map(CV);
  const int g = blockIdx.y;
  if (blockIdx.x == 0 && blockIdx.y == 0 && dgamma) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float a = 0.f, b = 0.f;
      for (int gg = 0; gg < groups; ++gg) {
        b += ws[(long long)gg * 2 * C + c];
        a += ws[(long long)gg * 2 * C + C + c];
      }
      dgamma[c] += a;
      dbeta[c] += b;
    }
  }
  if (!m.active) return;
  const float invM = 1.f / (float)rows_per_group;
  constexpr bool remask = kRemask;
  float mu[8], rs[8], k0[8], a1[8], a2[8], sc[kRemask ? 8 : 1], sf[kRemask ? 8 : 1];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = m.cv * 8 + j;
    mu[j] = mean[g * C + c];
    rs[j] = rstd[g * C + c];
    if constexpr (kRemask) bn_scale_shift(mu[j], rs[j], gamma[c], beta[c], sc[j], sf[j]);
    k0[j] = gamma[c] * rs[j];
    a1[j] = ws[(long long)g * 2 * C + c] * invM;
    a2[j] = ws[(long long)g * 2 * C + C + c] * invM;
  }
  const long long row0 = (long long)g * rows_per_group;
  for (long long r = (long long)blockIdx.x * m.rpb + m.rlane; r < rows_per_group; r += (long long)gridDim.x * m.rpb) {
    float d[8], yy[8], xx[8], o[8];
    ld8(dy + (row0 + r) * dy_rs + dy_co + m.cv * 8, d);
    ld8(x + (row0 + r) * x_rs + x_co + m.cv * 8, xx);
    if (relu && !remask) ld8(y + (row0 + r) * y_rs + y_co + m.cv * 8, yy);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if constexpr (kRemask) yy[j] = fmaf(xx[j], sc[j], sf[j]);
      const float dr = (!relu || yy[j] > 0.f) ? d[j] : 0.f;
      o[j] = k0[j] * (dr - a1[j] - (xx[j] - mu[j]) * rs[j] * a2[j]);
    }
    st8(dx + (row0 + r) * dx_rs + dx_co + m.cv * 8, o, true);       // dgrad / wgrad operand
  }
}

// ---------------------------------------------------------------------------------------------
struct PoolGeom {
  int N, C, Ti, Hi, Wi, To, Ho, Wo, kt, kh, kw, st, sh, sw, pt, ph, pw;
};
// v / s and v % s for the (power-of-two in this model) pooling strides without the emulated integer division
__device__ __forceinline__ bool pool_div(int v, int s, int& q) {
  if ((s & (s - 1)) == 0) {
    q = v >> (31 - __clz(s));
    return (v & (s - 1)) == 0;
  }
  q = v / s;
  return v - q * s == 0;
}

__global__ void __launch_bounds__(kBlock) maxpool_fwd_kernel(const bf16* __restrict__ x, long long x_rs, int x_co,
                                                             bf16* __restrict__ y, long long y_rs, int y_co,
                                                             uint8_t* __restrict__ idx, PoolGeom G) {
  const int CV = G.C / 8;
  // (host guarantees the element count < 2^31: 32-bit index arithmetic instead of emulated 64-bit divisions)
  const unsigned total = (unsigned)G.N * G.To * G.Ho * G.Wo * CV;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int cv = (int)(i % (unsigned)CV);
    unsigned o = i / (unsigned)CV;
    const long long orow = o;
    const int ow = (int)(o % (unsigned)G.Wo); o /= (unsigned)G.Wo;
    const int oh = (int)(o % (unsigned)G.Ho); o /= (unsigned)G.Ho;
    const int ot = (int)(o % (unsigned)G.To); o /= (unsigned)G.To;
    const int n = (int)o;
    // packed bf16x2 running maximum + packed 16-bit tap indices: per tap and channel pair ONE compare-to-mask and two
    // bit selects (ncu r01b: the fp32 compare / select version was ALU-bound at 25 % of HBM bandwidth).  Bit selects keep
    // the reference semantics exactly: strict '>' (first occurrence wins, like ATen), NaN never wins, -0/+0 untouched.
    uint32_t best2[4], idx2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      best2[j] = 0xff80ff80u;   // (-inf, -inf)
      idx2[j] = 0x00ff00ffu;
    }
    int tap = 0;
    bool seen_pad = false;
    for (int a = 0; a < G.kt; ++a) {
      const int it = ot * G.st - G.pt + a;
      for (int b = 0; b < G.kh; ++b) {
        const int ih = oh * G.sh - G.ph + b;
        for (int c = 0; c < G.kw; ++c, ++tap) {
          const int iw = ow * G.sw - G.pw + c;
          const bool inb = (unsigned)it < (unsigned)G.Ti && (unsigned)ih < (unsigned)G.Hi && (unsigned)iw < (unsigned)G.Wi;
          // F.pad zeros are real candidates (pytorch_i3d.py:44) -- but only the FIRST padding tap can ever win: once it has
          // been considered every running maximum is >= 0 and '>' is strict, so later padding zeros are skipped (the 3x3x3
          // pools of Mixed_4b..4f see one frame: 18 of their 27 taps are padding)
          if (!inb) {
            if (seen_pad) continue;
            seen_pad = true;
          }
          uint4 raw = make_uint4(0, 0, 0, 0);
          if (inb) raw = ld16(x + ((((long long)n * G.Ti + it) * G.Hi + ih) * G.Wi + iw) * x_rs + x_co + cv * 8);
          const uint32_t t2 = inb ? ((uint32_t)tap | ((uint32_t)tap << 16)) : 0x00ff00ffu;
          const uint32_t rv[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&rv[j]),
                                           *reinterpret_cast<const __nv_bfloat162*>(&best2[j]));
            best2[j] = (rv[j] & m) | (best2[j] & ~m);
            idx2[j] = (t2 & m) | (idx2[j] & ~m);
          }
        }
      }
    }
    st16(y + orow * y_rs + y_co + cv * 8, make_uint4(best2[0], best2[1], best2[2], best2[3]));
    uint2 pk;
    pk.x = __byte_perm(idx2[0], idx2[1], 0x6420);
    pk.y = __byte_perm(idx2[2], idx2[3], 0x6420);
    *reinterpret_cast<uint2*>(idx + orow * G.C + cv * 8) = pk;
  }
}

// fp32 activations (precision mode 1): plain compare / select, same semantics (strict '>', first occurrence wins, zero
// padding candidates are real, NaN never wins).
__global__ void __launch_bounds__(kBlock) maxpool_fwd_f32_kernel(const float* __restrict__ x, long long x_rs, int x_co,
                                                                 float* __restrict__ y, long long y_rs, int y_co,
                                                                 uint8_t* __restrict__ idx, PoolGeom G) {
  const int CV = G.C / 8;
  const unsigned total = (unsigned)G.N * G.To * G.Ho * G.Wo * CV;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int cv = (int)(i % (unsigned)CV);
    unsigned o = i / (unsigned)CV;
    const long long orow = o;
    const int ow = (int)(o % (unsigned)G.Wo); o /= (unsigned)G.Wo;
    const int oh = (int)(o % (unsigned)G.Ho); o /= (unsigned)G.Ho;
    const int ot = (int)(o % (unsigned)G.To); o /= (unsigned)G.To;
    const int n = (int)o;
    float best[8];
    uint32_t bidx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = __int_as_float(0xff800000);
      bidx[j] = 0xffu;
    }
    int tap = 0;
    bool seen_pad = false;
    for (int a = 0; a < G.kt; ++a) {
      const int it = ot * G.st - G.pt + a;
      for (int b = 0; b < G.kh; ++b) {
        const int ih = oh * G.sh - G.ph + b;
        for (int c = 0; c < G.kw; ++c, ++tap) {
          const int iw = ow * G.sw - G.pw + c;
          const bool inb = (unsigned)it < (unsigned)G.Ti && (unsigned)ih < (unsigned)G.Hi && (unsigned)iw < (unsigned)G.Wi;
          if (!inb) {                   // only the first padding zero can win (see maxpool_fwd_kernel)
            if (seen_pad) continue;
            seen_pad = true;
          }
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = 0.f;
          if (inb) ld8(x + ((((long long)n * G.Ti + it) * G.Hi + ih) * G.Wi + iw) * x_rs + x_co + cv * 8, v);
          const uint32_t t = inb ? (uint32_t)tap : 0xffu;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (v[j] > best[j]) {
              best[j] = v[j];
              bidx[j] = t;
            }
        }
      }
    }
    st8(y + orow * y_rs + y_co + cv * 8, best, false);
    uint2 pk;
    pk.x = bidx[0] | (bidx[1] << 8) | (bidx[2] << 16) | (bidx[3] << 24);
    pk.y = bidx[4] | (bidx[5] << 8) | (bidx[6] << 16) | (bidx[7] << 24);
    *reinterpret_cast<uint2*>(idx + orow * G.C + cv * 8) = pk;
  }
}

template <typename T>
__global__ void __launch_bounds__(kBlock) maxpool_bwd_kernel(const T* __restrict__ dy, long long dy_rs, int dy_co,
                                                             const uint8_t* __restrict__ idx, T* __restrict__ dx,
                                                             long long dx_rs, int dx_co, PoolGeom G, int accumulate) {
  const int CV = G.C / 8;
  const unsigned total = (unsigned)G.N * G.Ti * G.Hi * G.Wi * CV;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int cv = (int)(i % (unsigned)CV);
    unsigned p = i / (unsigned)CV;
    const long long irow = p;
    const int iw = (int)(p % (unsigned)G.Wi); p /= (unsigned)G.Wi;
    const int ih = (int)(p % (unsigned)G.Hi); p /= (unsigned)G.Hi;
    const int it = (int)(p % (unsigned)G.Ti); p /= (unsigned)G.Ti;
    const int n = (int)p;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // walk the output windows that contain this input position (per dimension o in [ceil((i + p - k + 1) / s),
    // floor((i + p) / s)] clipped to the output), not all k^3 taps: a stride-2 3x3 pool has at most 2 x 2 of them, and the
    // tap-loop version spent its time rejecting the others (ncu r02b: 0.28 ms at 73 % issue-slot utilisation)
    const int ot_hi = min((it + G.pt) / G.st, G.To - 1), ot_lo = max((it + G.pt - G.kt + G.st) / G.st, 0);
    const int oh_hi = min((ih + G.ph) / G.sh, G.Ho - 1), oh_lo = max((ih + G.ph - G.kh + G.sh) / G.sh, 0);
    const int ow_hi = min((iw + G.pw) / G.sw, G.Wo - 1), ow_lo = max((iw + G.pw - G.kw + G.sw) / G.sw, 0);
    for (int ot = ot_lo; ot <= ot_hi; ++ot) {
      const int a = it + G.pt - ot * G.st;
      for (int oh = oh_lo; oh <= oh_hi; ++oh) {
        const int b = ih + G.ph - oh * G.sh;
        for (int ow = ow_lo; ow <= ow_hi; ++ow) {
          const int c = iw + G.pw - ow * G.sw;
          const int tap = (a * G.kh + b) * G.kw + c;
          const long long orow = (((long long)n * G.To + ot) * G.Ho + oh) * G.Wo + ow;
          const uint2 pk = *reinterpret_cast<const uint2*>(idx + orow * G.C + cv * 8);
          float d[8];
          ld8(dy + orow * dy_rs + dy_co + cv * 8, d);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t w = j < 4 ? pk.x : pk.y;
            const int bi = (w >> ((j & 3) * 8)) & 0xff;
            if (bi == tap) acc[j] += d[j];
          }
        }
      }
    }
    T* dst = dx + irow * dx_rs + dx_co + cv * 8;
    if (accumulate) {
      float e[8];
      ld8(dst, e);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += e[j];
    }
    st8(dst, acc, false);
  }
}

// ---------------------------------------------------------------------------------------------
// Specialised max-pool kernels (r02d).  The generic kernels above decode (n, t, h, w, cv) from a flat index and walk
// runtime window extents: ncu r02b showed them ISSUE bound (73 % issue slots, ~580 instructions per thread, most of them
// emulated integer divisions) at 1 TB/s.  Here the window (KT,KH,KW) and strides are template parameters, one CTA owns
// one output row (fwd) / input row (bwd) through a 3-D grid -- no division except one multiply-shift per thread -- and
// the tap loops unroll.  Same semantics bit for bit (strict '>', first occurrence in (t,h,w) tap order, the first
// padding zero is a candidate with index 0xff).
struct PoolRowGeom {
  int C, CV, Ti, Hi, Wi, To, Ho, Wo, pt, ph, pw;
  uint32_t cv_magic;   // floor(2^32 / CV) + 1: e / CV = umulhi(e, cv_magic) for e * CV < 2^32
};

template <typename T, int KT, int KH, int KW, int ST, int SH, int SW>
__global__ void __launch_bounds__(kBlock) maxpool_fwd_row_kernel(const T* __restrict__ x, long long x_rs, int x_co, T* __restrict__ y,
                                                                 long long y_rs, int y_co, uint8_t* __restrict__ idx, PoolRowGeom G) {
  const int oh = blockIdx.x, ot = blockIdx.y, n = blockIdx.z;
  const long long orow0 = (((long long)n * G.To + ot) * G.Ho + oh) * G.Wo;
  const int total = G.Wo * G.CV;
  for (int e = threadIdx.x; e < total; e += kBlock) {
    const int ow = G.CV == 1 ? e : (int)__umulhi((uint32_t)e, G.cv_magic);   // (2^32 / 1 + 1 does not fit the magic word)
    const int cv = e - ow * G.CV;
    bool seen_pad = false;
    if constexpr (sizeof(T) == 2) {
      uint32_t best2[4], idx2[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        best2[j] = 0xff80ff80u;   // (-inf, -inf)
        idx2[j] = 0x00ff00ffu;
      }
#pragma unroll
      for (int a = 0; a < KT; ++a) {
        const int it = ot * ST - G.pt + a;
#pragma unroll
        for (int b = 0; b < KH; ++b) {
          const int ih = oh * SH - G.ph + b;
          const bool row_ok = (unsigned)it < (unsigned)G.Ti && (unsigned)ih < (unsigned)G.Hi;
          if (!row_ok) {
            // CTA-uniform: the whole tap row is padding (18 of the 27 taps of the single-frame 3x3x3 pools).  Only the first
            // padding tap of the window can win, so the row is at most one zero candidate.
            if (!seen_pad) {
              seen_pad = true;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t zero = 0u;
                const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&zero),
                                               *reinterpret_cast<const __nv_bfloat162*>(&best2[j]));
                best2[j] = best2[j] & ~m;
                idx2[j] = (0x00ff00ffu & m) | (idx2[j] & ~m);
              }
            }
            continue;
          }
          const bf16* xrow = x + (((long long)n * G.Ti + it) * G.Hi + ih) * G.Wi * x_rs + x_co + cv * 8;
#pragma unroll
          for (int c = 0; c < KW; ++c) {
            const int tap = (a * KH + b) * KW + c;
            const int iw = ow * SW - G.pw + c;
            const bool inb = (unsigned)iw < (unsigned)G.Wi;
            if (!inb) {
              if (seen_pad) continue;
              seen_pad = true;
            }
            uint4 raw = make_uint4(0, 0, 0, 0);
            if (inb) raw = ld16(reinterpret_cast<const bf16*>(xrow) + (long long)iw * x_rs);
            const uint32_t t2 = inb ? ((uint32_t)tap | ((uint32_t)tap << 16)) : 0x00ff00ffu;
            const uint32_t rv[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&rv[j]),
                                             *reinterpret_cast<const __nv_bfloat162*>(&best2[j]));
              best2[j] = (rv[j] & m) | (best2[j] & ~m);
              idx2[j] = (t2 & m) | (idx2[j] & ~m);
            }
          }
        }
      }
      st16(reinterpret_cast<bf16*>(y) + (orow0 + ow) * y_rs + y_co + cv * 8, make_uint4(best2[0], best2[1], best2[2], best2[3]));
      uint2 pk;
      pk.x = __byte_perm(idx2[0], idx2[1], 0x6420);
      pk.y = __byte_perm(idx2[2], idx2[3], 0x6420);
      *reinterpret_cast<uint2*>(idx + (orow0 + ow) * G.C + cv * 8) = pk;
    } else {
      float best[8];
      uint32_t bidx[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        best[j] = __int_as_float(0xff800000);
        bidx[j] = 0xffu;
      }
#pragma unroll
      for (int a = 0; a < KT; ++a) {
        const int it = ot * ST - G.pt + a;
#pragma unroll
        for (int b = 0; b < KH; ++b) {
          const int ih = oh * SH - G.ph + b;
          const bool row_ok = (unsigned)it < (unsigned)G.Ti && (unsigned)ih < (unsigned)G.Hi;
          if (!row_ok) {          // CTA-uniform padding row: at most one zero candidate (see the bf16 branch)
            if (!seen_pad) {
              seen_pad = true;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (0.f > best[j]) {
                  best[j] = 0.f;
                  bidx[j] = 0xffu;
                }
            }
            continue;
          }
          const T* xrow = x + (((long long)n * G.Ti + it) * G.Hi + ih) * G.Wi * x_rs + x_co + cv * 8;
#pragma unroll
          for (int c = 0; c < KW; ++c) {
            const int tap = (a * KH + b) * KW + c;
            const int iw = ow * SW - G.pw + c;
            const bool inb = (unsigned)iw < (unsigned)G.Wi;
            if (!inb) {
              if (seen_pad) continue;
              seen_pad = true;
            }
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
            if (inb) ld8(xrow + (long long)iw * x_rs, v);
            const uint32_t t = inb ? (uint32_t)tap : 0xffu;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (v[j] > best[j]) {
                best[j] = v[j];
                bidx[j] = t;
              }
          }
        }
      }
      st8(y + (orow0 + ow) * y_rs + y_co + cv * 8, best, false);
      uint2 pk;
      pk.x = bidx[0] | (bidx[1] << 8) | (bidx[2] << 16) | (bidx[3] << 24);
      pk.y = bidx[4] | (bidx[5] << 8) | (bidx[6] << 16) | (bidx[7] << 24);
      *reinterpret_cast<uint2*>(idx + (orow0 + ow) * G.C + cv * 8) = pk;
    }
  }
}

// one CTA per input row; a thread visits the (at most ceil(K/S) per dimension) output windows that contain its position
template <typename T, int KT, int KH, int KW, int ST, int SH, int SW>
__global__ void __launch_bounds__(kBlock) maxpool_bwd_row_kernel(const T* __restrict__ dy, long long dy_rs, int dy_co,
                                                                 const uint8_t* __restrict__ idx, T* __restrict__ dx, long long dx_rs,
                                                                 int dx_co, PoolRowGeom G, int accumulate) {
  constexpr int MT = (KT + ST - 1) / ST, MH = (KH + SH - 1) / SH, MW = (KW + SW - 1) / SW;
  const int ih = blockIdx.x, it = blockIdx.y, n = blockIdx.z;
  const long long irow0 = (((long long)n * G.Ti + it) * G.Hi + ih) * G.Wi;
  const int ot_hi = min((it + G.pt) / ST, G.To - 1), ot_lo = max((it + G.pt - KT + ST) / ST, 0);
  const int oh_hi = min((ih + G.ph) / SH, G.Ho - 1), oh_lo = max((ih + G.ph - KH + SH) / SH, 0);
  const int total = G.Wi * G.CV;
  for (int e = threadIdx.x; e < total; e += kBlock) {
    const int iw = G.CV == 1 ? e : (int)__umulhi((uint32_t)e, G.cv_magic);
    const int cv = e - iw * G.CV;
    const int ow_hi = min((iw + G.pw) / SW, G.Wo - 1), ow_lo = max((iw + G.pw - KW + SW) / SW, 0);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int qt = 0; qt < MT; ++qt) {
      const int ot = ot_lo + qt;
      if (ot > ot_hi) break;
      const int a = it + G.pt - ot * ST;
#pragma unroll
      for (int qh = 0; qh < MH; ++qh) {
        const int oh = oh_lo + qh;
        if (oh > oh_hi) break;
        const int b = ih + G.ph - oh * SH;
        const long long orow0 = (((long long)n * G.To + ot) * G.Ho + oh) * G.Wo;
#pragma unroll
        for (int qw = 0; qw < MW; ++qw) {
          const int ow = ow_lo + qw;
          if (ow <= ow_hi) {
            const int c = iw + G.pw - ow * SW;
            const uint32_t tap = (uint32_t)((a * KH + b) * KW + c);
            const long long orow = orow0 + ow;
            const uint2 pk = *reinterpret_cast<const uint2*>(idx + orow * G.C + cv * 8);
            float d[8];
            ld8(dy + orow * dy_rs + dy_co + cv * 8, d);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t w = j < 4 ? pk.x : pk.y;
              const uint32_t bi = (w >> ((j & 3) * 8)) & 0xffu;
              if (bi == tap) acc[j] += d[j];
            }
          }
        }
      }
    }
    T* dst = dx + (irow0 + iw) * dx_rs + dx_co + cv * 8;
    if (accumulate) {
      float ev[8];
      ld8(dst, ev);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += ev[j];
    }
    st8(dst, acc, false);
  }
}

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kBlock) channel_scale_kernel(const T* __restrict__ x, long long x_rs, int x_co,
                                                               const float* __restrict__ scale, T* __restrict__ y, long long y_rs,
                                                               int y_co, int N, long long rows_per_n, int C) {
  const int CV = C / 8;
  const long long total = (long long)N * rows_per_n * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    const long long row = i / CV;
    const long long n = row / rows_per_n;
    float v[8];
    ld8(x + row * x_rs + x_co + cv * 8, v);
    const float* s = scale + n * C + cv * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= s[j];
    st8(y + row * y_rs + y_co + cv * 8, v, false);                   // x{0,2}: exact
  }
}

// grid (blocks, N): per-sample so the scale row is fixed per block
template <typename T>
__global__ void __launch_bounds__(kBlock) act_bwd_kernel(const T* __restrict__ dy, long long dy_rs, int dy_co,
                                                         const T* __restrict__ y, long long y_rs, int y_co,
                                                         const float* __restrict__ scale, T* __restrict__ dz, long long dz_rs,
                                                         int dz_co, float* __restrict__ dbias, long long rows_per_n, int C, int relu) {
  const int CV = C / 8;
  const RowMap m = row_map(CV);
  const int n = blockIdx.y;
  float s1[8], sc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = 0.f;
  if (m.active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sc[j] = scale ? scale[(long long)n * C + m.cv * 8 + j] : 1.f;
    const long long row0 = (long long)n * rows_per_n;
    // (not unrolled: 4 rows in flight per thread cost 24 registers and occupancy -- measured 366 -> 461 us in the step)
    for (long long r = (long long)blockIdx.x * m.rpb + m.rlane; r < rows_per_n; r += (long long)gridDim.x * m.rpb) {
      float d[8], yy[8];
      ld8(dy + (row0 + r) * dy_rs + dy_co + m.cv * 8, d);
      if (relu) ld8(y + (row0 + r) * y_rs + y_co + m.cv * 8, yy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = d[j] * sc[j];
        if (relu && !(yy[j] > 0.f)) v = 0.f;
        d[j] = v;
        s1[j] += v;
      }
      if (dz) st8(dz + (row0 + r) * dz_rs + dz_co + m.cv * 8, d, true);   // dgrad / wgrad operand
    }
  }
  if (!dbias) return;
  extern __shared__ float sh[];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  if (m.active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sh[m.cv * 8 + j], s1[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&dbias[i], sh[i]);
}

template <typename T>
__global__ void __launch_bounds__(kBlock) add_kernel(const T* __restrict__ a, long long a_rs, int a_co, const T* __restrict__ b,
                                                     long long b_rs, int b_co, T* __restrict__ o, long long o_rs, int o_co,
                                                     long long rows, int C) {
  const int CV = C / 8;
  const long long total = rows * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    const long long row = i / CV;
    float x[8], y[8];
    ld8(a + row * a_rs + a_co + cv * 8, x);
    ld8(b + row * b_rs + b_co + cv * 8, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    st8(o + row * o_rs + o_co + cv * 8, x, false);
  }
}

// fp32 view -> two compact bf16 tensors hi = bf16(x), lo = bf16(x - hi): x = hi + lo to 2^-17 relative.  The tf32
// precision mode's weight gradients run as three bf16 GEMMs (hi*hi + hi*lo + lo*hi, fp32 accumulate) on the
// position-major (MN-major) wgrad kernel: tcgen05 kind::tf32 reads MN-major 32-bit operands only through a dedicated
// shared-memory layout (128B swizzle, 32-byte base) that the im2col TMA path does not produce.
__global__ void __launch_bounds__(kBlock) split_bf16_kernel(const float* __restrict__ x, long long x_rs, int x_co, bf16* __restrict__ hi,
                                                            bf16* __restrict__ lo, long long rows, int C) {
  const int CV = C / 8;
  const long long total = rows * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    const long long row = i / CV;
    float v[8], h[8], l[8];
    ld8(x + row * x_rs + x_co + cv * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      h[j] = __bfloat162float(__float2bfloat16(v[j]));
      l[j] = v[j] - h[j];
    }
    st16(hi + row * C + cv * 8, pack8(h));
    st16(lo + row * C + cv * 8, pack8(l));
  }
}

// ---------------------------------------------------------------------------------------------
// Folded stem (Conv3d_1a_7x7, pytorch_i3d.py:224: 3 -> 64, 7x7x7, stride 2): the time axis is folded into the channel
// axis so that the few-channel 3-D convolution becomes a 2-D convolution over 64 "channels" on the TMA im2col path,
// without the (rows x 1088) im2col matrix (3.5 GB written and read twice per step at 16+16 clips).
//   xs[n][h][w][fp * 4 + c] = x[n][fp - pt][h][w][c]   (fp = padded frame index, zero outside the clip; c < 4)
// Output frame t reads padded frames st*t .. st*t + kt - 1: one weight set per output frame (an output class of the
// implicit GEMM) holds the kernel shifted to that window,
//   W2[t][co][kh*kw_+kw][fp * 4 + c] = w[co][c][fp - st*t][kh][kw]   (zero outside the window / for c >= Cin).
struct StemFold {
  int N, T, H, W, Cs, pt, Tp;     // Tp = padded frames held per pixel (Tp * 4 = folded channels, <= 16)
};
template <typename T>
__global__ void __launch_bounds__(kBlock) stem_fold_input_kernel(const T* __restrict__ x, T* __restrict__ xs, StemFold G, long long total) {
  // one thread per (pixel, pair of padded frames): 8 folded channels = one 16-byte (bf16) / 32-byte (fp32) store
  const int pairs = G.Tp / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % pairs);
    long long pix = i / pairs;                       // (n, h, w)
    const long long hw = (long long)G.H * G.W;
    const long long n = pix / hw, r = pix - n * hw;
    float v[8];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int f = 2 * q + j - G.pt;
      float t8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (f >= 0 && f < G.T) ld8(x + ((n * G.T + f) * hw + r) * G.Cs, t8);
#pragma unroll
      for (int c = 0; c < 4; ++c) v[j * 4 + c] = t8[c];
    }
    st8(xs + pix * (G.Tp * 4) + q * 8, v, false);      // values are already operand-rounded
  }
}

// Clips straight into the folded stem layout (no intermediate channels-last copy): in = fp32 (P,C,T,H,W) or uint8 (then
// / 255 and, with `mirror`, also the W-mirrored clip at batch offset P).  One thread per (n, h, pair of padded frames, w)
// with w fastest: every load is a coalesced run of one (channel, frame) plane.
template <typename T, typename S>
__global__ void __launch_bounds__(kBlock) clips_to_folded_kernel(const S* __restrict__ in, T* __restrict__ xs, int P, int C, int Tn, int H,
                                                                 int W, int pt, int Tp, int mirror, long long total) {
  const int pairs = Tp / 2;
  const float scale = sizeof(S) == 1 ? 1.f / 255.f : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // frame pair fastest: the 8 threads of a pixel write its whole 128-byte row (with w fastest every 16-byte store went to
    // a different line: in-graph 176 us for 283 MB, r02d_graph_busy.json); the 4-byte reads of 8 planes x 4 pixels per warp
    // still use full 32-byte sectors across neighbouring warps
    // (host guarantees total < 2^31: 32-bit divisions -- the emulated 64-bit ones were most of this kernel's instructions)
    unsigned r = (unsigned)i;
    const int q = (int)(r % (unsigned)pairs);
    r /= (unsigned)pairs;
    const int w = (int)(r % (unsigned)W);
    r /= (unsigned)W;
    const int h = (int)(r % (unsigned)H);
    const long long n = r / (unsigned)H;
    float v[8];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int f = 2 * q + j - pt;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float x = 0.f;
        if (c < C && f >= 0 && f < Tn) {
          const S raw = in[(((n * C + c) * Tn + f) * H + h) * W + w];
          x = sizeof(S) == 1 ? (float)raw / 255.f : (float)raw;
        }
        v[j * 4 + c] = x;
      }
    }
    (void)scale;
    st8(xs + ((n * H + h) * W + w) * (Tp * 4) + q * 8, v, true);
    if (mirror) st8(xs + (((P + n) * H + h) * W + (W - 1 - w)) * (Tp * 4) + q * 8, v, true);
  }
}

__global__ void stem_fold_weights_kernel(const float* __restrict__ w, float* __restrict__ w2, int Cout, int Cin, int kt, int khw, int st,
                                         int To, int Kf) {
  const long long total = (long long)To * Cout * khw * Kf;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % Kf);
    long long r = i / Kf;
    const int tap = (int)(r % khw);
    r /= khw;
    const int co = (int)(r % Cout), t = (int)(r / Cout);
    const int fp = e >> 2, c = e & 3, k = fp - st * t;
    float v = 0.f;
    if (c < Cin && k >= 0 && k < kt) v = w[(((long long)co * Cin + c) * kt + k) * khw + tap];
    w2[i] = v;
  }
}

// dw[co][c][k][kh][kw] += sum_t dW2[t][co][tap][(st*t + k) * 4 + c]
__global__ void stem_unfold_wgrad_kernel(const float* __restrict__ dw2, float* __restrict__ dw, int Cout, int Cin, int kt, int khw, int st,
                                         int To, int Kf) {
  const long long total = (long long)Cout * Cin * kt * khw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % khw);
    long long r = i / khw;
    const int k = (int)(r % kt);
    r /= kt;
    const int c = (int)(r % Cin), co = (int)(r / Cin);
    float acc = 0.f;
    for (int t = 0; t < To; ++t) {
      const int e = (st * t + k) * 4 + c;
      if (e < Kf) acc += dw2[(((long long)t * Cout + co) * khw + tap) * Kf + e];
    }
    dw[i] += acc;
  }
}

// ---------------------------------------------------------------------------------------------
// smooth = ConvTranspose3d(128->1, k3, p1): out[o] = bias + sum_k P_k[o + 1 - k], P planar fp32 [32][rows]
__global__ void __launch_bounds__(kBlock) stencil27_fwd_kernel(const float* __restrict__ P, float* __restrict__ out,
                                                               const float* __restrict__ bias_p, int N, int T, int H, int W) {
  const long long rows = (long long)N * T * H * W;
  const float bias = bias_p ? __ldg(bias_p) : 0.f;
  // (host guarantees rows < 2^31: 32-bit index arithmetic, the emulated 64-bit divisions dominated this kernel)
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)rows; i += gridDim.x * blockDim.x) {
    unsigned p = i;
    const int w = (int)(p % (unsigned)W); p /= (unsigned)W;
    const int h = (int)(p % (unsigned)H); p /= (unsigned)H;
    const int t = (int)(p % (unsigned)T);
    float acc = bias;
    int tap = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int tt = t + 1 - a;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int hh = h + 1 - b;
#pragma unroll
        for (int c = 0; c < 3; ++c, ++tap) {
          const int ww = w + 1 - c;
          if ((unsigned)tt < (unsigned)T && (unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W)
            acc += __ldg(P + (long long)tap * rows + (long long)i + (((1 - a) * H + (1 - b)) * W + (1 - c)));
        }
      }
    }
    out[i] = acc;
  }
}
// adjoint: dP[i][k] = dout[i - 1 + k] (bf16 rows of 32, taps 27..31 zero)
template <typename T_>
__global__ void __launch_bounds__(kBlock) stencil27_bwd_kernel(const float* __restrict__ dout, T_* __restrict__ dP,
                                                               float* __restrict__ dbias, int N, int T, int H, int W, int cpad) {
  const long long rows = (long long)N * T * H * W;
  float bsum = 0.f;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)rows; i += gridDim.x * blockDim.x) {
    unsigned p = i;
    const int w = (int)(p % (unsigned)W); p /= (unsigned)W;
    const int h = (int)(p % (unsigned)H); p /= (unsigned)H;
    const int t = (int)(p % (unsigned)T);
    float v[32];
    int tap = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int tt = t - 1 + a;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int hh = h - 1 + b;
#pragma unroll
        for (int c = 0; c < 3; ++c, ++tap) {
          const int ww = w - 1 + c;
          float g = 0.f;
          if ((unsigned)tt < (unsigned)T && (unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W)
            g = __ldg(dout + (long long)i + (((a - 1) * H + (b - 1)) * W + (c - 1)));
          v[tap] = g;
        }
      }
    }
#pragma unroll
    for (int k = 27; k < 32; ++k) v[k] = 0.f;
    bsum += v[13];  // centre tap == dout[i]
    T_* dst = dP + (long long)i * cpad;
    const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) st8(dst + q * 8, v + q * 8, true);
    for (int q = 4; q < cpad / 8; ++q) st8(dst + q * 8, zero8, false);
  }
  if (dbias) {
    __shared__ float sb[kBlock / 32];
    bsum = warp_sum(bsum);
    if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = bsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int q = 0; q < kBlock / 32; ++q) t += sb[q];
      atomicAdd(dbias, t);
    }
  }
}

// Explicit im2col for convolutions with very few input channels (the 7x7x7 stride-2 RGB stem): the implicit-GEMM
// gather would issue one 16-byte load per (row, tap) for 6 useful bytes.  out[(n,to,ho,wo)][tap*C + c] (bf16, zero
// padded to Kpad, a multiple of 64) is then consumed by the TMA GEMM path as a 1x1x1 convolution with Cin = Kpad.
struct Im2colGeom {
  int N, Cs, C, T, H, W, To, Ho, Wo, kt, kh, kw, st, sh, sw, pt, ph, pw, K, Kpad;
};
// One block per strip of WB consecutive output positions along W at a fixed (n, to, ho).  The strip's input footprint
// -- kt*kh input rows x ((WB-1)*sw + kw) pixels, real channels only -- is staged in shared memory (zero outside the
// volume), so that the (kw x C) block of an im2col row for a fixed (kt, kh) is one contiguous run of the staged row.
// The Kpad-wide rows then stream out as coalesced 16-byte stores, each assembled from 8 staged elements through a
// per-block column -> staged-offset table (padding columns point at a zero slot).
constexpr int kIm2colWB = 56;
template <typename E>   // E = the element's storage word: uint16_t (bf16) or uint32_t (fp32)
__global__ void __launch_bounds__(512) im2col_small_kernel(const E* __restrict__ x, E* __restrict__ out, Im2colGeom G,
                                                           int strips, int pitch) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int nthr = (int)blockDim.x;      // a multiple of the 16-byte column groups per row (host)
  const int pairs = G.kt * G.kh;
  const int run = G.kw * G.C;
  const int zero_slot = pairs * pitch;
  const int zero_len = (kIm2colWB - 1) * G.sw * G.C + 8;      // padding columns may be read at zero_slot + row base too
  unsigned short* tab = reinterpret_cast<unsigned short*>(s_raw);                       // Kpad entries
  E* sin = reinterpret_cast<E*>(s_raw + (((size_t)G.Kpad * 2 + 15) & ~(size_t)15));     // pairs * pitch + zero_len elements
  for (int k = threadIdx.x; k < G.Kpad; k += nthr) {
    int off = zero_slot;
    if (k < G.K) {
      const int p = k / run;
      off = p * pitch + (k - p * run);
    }
    tab[k] = (unsigned short)off;
  }
  for (int k = threadIdx.x; k < zero_len; k += nthr) sin[zero_slot + k] = (E)0;
  __syncthreads();
  const int wpix = (kIm2colWB - 1) * G.sw + G.kw;                                       // staged pixels per input row
  const int vec_per_row = G.Kpad / 8;
  // Each thread owns ONE 8-element column group of the row (its 8 staged offsets live in registers) and walks the rows
  // of the strip: per output vector 8 shared loads + one 16-byte (bf16) / two 16-byte (fp32) stores (ncu r01f: the
  // per-vector table fetch and zero-slot selects made this kernel issue/L1 bound at 2.4 TB/s of writes).
  const int lanes = (nthr / vec_per_row) * vec_per_row;                               // threads that own a column group
  const int rstep = nthr / vec_per_row;                                                 // rows written per sweep
  const int my_v = threadIdx.x % vec_per_row, my_r0 = threadIdx.x / vec_per_row;
  unsigned o[8];
  {
    const uint4 tv = *reinterpret_cast<const uint4*>(tab + my_v * 8);
    o[0] = tv.x & 0xffffu; o[1] = tv.x >> 16; o[2] = tv.y & 0xffffu; o[3] = tv.y >> 16;
    o[4] = tv.z & 0xffffu; o[5] = tv.z >> 16; o[6] = tv.w & 0xffffu; o[7] = tv.w >> 16;
  }
  const long long nblk = (long long)G.N * G.To * G.Ho * strips;
  for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    long long r = blk;
    const int sidx = (int)(r % strips); r /= strips;
    const int ho = (int)(r % G.Ho); r /= G.Ho;
    const int to = (int)(r % G.To); r /= G.To;
    const int n = (int)r;
    const int wo0 = sidx * kIm2colWB;
    const int nrow = min(kIm2colWB, G.Wo - wo0);
    const int t0 = to * G.st - G.pt, h0 = ho * G.sh - G.ph, w0 = wo0 * G.sw - G.pw;
    __syncthreads();                                                                    // previous strip fully written out
    // staging: one warp per (kt, kh) input row, lanes along W -- no per-pixel index divisions
    // (only FULL warps stage: the block size is a multiple of the row's column groups, not of 32)
    for (int p = ((int)threadIdx.x >> 5) < (nthr >> 5) ? (int)(threadIdx.x >> 5) : pairs; p < pairs; p += nthr >> 5) {
      const int a = p / G.kh, b = p - a * G.kh;
      const int t = t0 + a, h = h0 + b;
      const bool rowok = (unsigned)t < (unsigned)G.T && (unsigned)h < (unsigned)G.H;
      const E* src = x + (((long long)n * G.T + t) * G.H + h) * (long long)G.W * G.Cs;
      E* drow = sin + p * pitch;
      for (int wl = (int)(threadIdx.x & 31); wl < wpix; wl += 32) {
        const int w = w0 + wl;
        E pv[8];                                                                        // Cs == 8: one pixel = 8 elements
#pragma unroll
        for (int q = 0; q < (int)(sizeof(E) * 8 / 16); ++q) reinterpret_cast<uint4*>(pv)[q] = make_uint4(0, 0, 0, 0);
        if (rowok && (unsigned)w < (unsigned)G.W) {
#pragma unroll
          for (int q = 0; q < (int)(sizeof(E) * 8 / 16); ++q)
            reinterpret_cast<uint4*>(pv)[q] = reinterpret_cast<const uint4*>(src + (long long)w * G.Cs)[q];
        }
        E* dst = drow + wl * G.C;
        for (int ch = 0; ch < G.C; ++ch) dst[ch] = pv[ch];
      }
    }
    __syncthreads();
    if ((int)threadIdx.x < lanes) {
      E* orow = out + ((((long long)n * G.To + to) * G.Ho + ho) * (long long)G.Wo + wo0) * G.Kpad + my_v * 8;
      const int bstep = G.sw * G.C;
      for (int rl = my_r0; rl < nrow; rl += rstep) {
        const E* sb = sin + rl * bstep;
        E v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = sb[o[q]];
#pragma unroll
        for (int q = 0; q < (int)(sizeof(E) * 8 / 16); ++q)
          reinterpret_cast<uint4*>(orow + (long long)rl * G.Kpad)[q] = reinterpret_cast<const uint4*>(v)[q];
      }
    }
  }
}

__global__ void fill_f32_kernel(float* p, long long n, float v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

inline int grid_for(long long work_items, int per_block = kBlock, int waves = 8) {
  long long b = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)b2c_num_sms() * waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
// deterministic mode (tests): cross-block fp32 atomics of the BatchNorm reductions are avoided by using ONE block per
// statistic group, so two runs -- or two schedules of the same step -- produce bit-identical activations.
static int g_deterministic = 0;
static int g_pool_generic = -1;   // 1: generic max-pool kernels even where a specialisation exists (tests compare the two)
inline size_t bn_red_smem(int C) {
  int rpb = kBlock / (C / 8);
  if (rpb < 1) rpb = 1;
  return (size_t)rpb * 2 * C * sizeof(float);
}
// grid of the two BatchNorm REDUCTION kernels: every block ends with 2 C fp32 atomics on the same 2 C addresses, which the L2
// serialises (592 blocks on a 28 x 28 layer: ~6 us of atomics for ~2 us of loads).  At least 8 rows per thread, at most one
// block per SM and group pair.
inline int reduce_grid(long long rows, int C, int groups) {
  int rpb = kBlock / (C / 8);
  if (rpb < 1) rpb = 1;
  long long b = (rows + (long long)rpb * 8 - 1) / ((long long)rpb * 8);
  long long cap = (long long)b2c_num_sms() * 2 / (groups > 0 ? groups : 1);
  if (cap < 1) cap = 1;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
// grid of the two BatchNorm APPLY kernels: every thread first gathers its 8 channels' statistics / affine parameters (up to 48
// scalar loads), so give it at least 4 rows to stream; at most 4 blocks per SM over all groups.
inline int apply_grid(long long rows, int C, int groups) {
  int rpb = kBlock / (C / 8);
  if (rpb < 1) rpb = 1;
  long long b = (rows + (long long)rpb * 4 - 1) / ((long long)rpb * 4);
  long long cap = (long long)b2c_num_sms() * 4 / (groups > 0 ? groups : 1);
  if (cap < 1) cap = 1;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
inline int row_grid(long long rows, int C, int waves = 4) {
  int rpb = kBlock / (C / 8);
  if (rpb < 1) rpb = 1;
  return grid_for(rows, rpb, waves);
}

// launch KERNEL<T> with T = the activation storage type of the current precision mode; the argument list may use T
#define LAUNCH_T(KERNEL, GRID, BLOCK, SMEM, STREAM, ...)                                   \
  do {                                                                                     \
    if (b2c_precision()) {                                                                 \
      using T = float;                                                                     \
      KERNEL<T><<<GRID, BLOCK, SMEM, (cudaStream_t)(STREAM)>>>(__VA_ARGS__);               \
    } else {                                                                               \
      using T = bf16;                                                                      \
      KERNEL<T><<<GRID, BLOCK, SMEM, (cudaStream_t)(STREAM)>>>(__VA_ARGS__);               \
    }                                                                                      \
  } while (0)

// geometry of the specialised (row-per-CTA) pooling kernels; returns 0 when the configuration has no specialisation
// (B2C_POOL_GENERIC=1 forces the generic kernels), else the index of the instantiated (window, stride) pair
inline int pool_row_geom(PoolRowGeom& R, int N, int C, int Ti, int Hi, int Wi, int To, int Ho, int Wo, int kt, int kh, int kw, int st,
                         int sh, int sw, int pt, int ph, int pw, int Wthreads) {
  if (g_pool_generic < 0) {
    const char* e = getenv("B2C_POOL_GENERIC");
    g_pool_generic = (e && e[0] == '1') ? 1 : 0;
  }
  const int generic = g_pool_generic;
  const int CV = C / 8;
  R = PoolRowGeom{C, CV, Ti, Hi, Wi, To, Ho, Wo, pt, ph, pw, (uint32_t)((1ULL << 32) / (unsigned)CV + 1)};
  if (generic || N > 65535 || Ti > 65535 || To > 65535 || (long long)Wthreads * CV * CV >= (1LL << 31)) return 0;
  if (kt == 1 && kh == 3 && kw == 3 && st == 1 && sh == 2 && sw == 2) return 1;
  if (kt == 3 && kh == 3 && kw == 3 && st == 2 && sh == 1 && sw == 1) return 2;
  if (kt == 3 && kh == 3 && kw == 3 && st == 1 && sh == 1 && sw == 1) return 3;
  return 0;
}

#define CHECK_VIEW(name, C, rs, co)                                                                         \
  B2C_REQUIRE((C) > 0 && (C) % 8 == 0 && (rs) % 8 == 0 && (co) % 8 == 0 && (C) / 8 <= kBlock, name ": bad view C=%d rs=%lld co=%d", \
              (int)(C), (long long)(rs), (int)(co))

}  // namespace

B2C_API int b2c_ncdhw_to_ndhwc(const float* in, void* out, int32_t N, int32_t C, int64_t THW, int32_t Cpad, b2c_stream_t s) {
  B2C_REQUIRE(in && out && N > 0 && C > 0 && Cpad % 8 == 0 && Cpad >= C, "ncdhw_to_ndhwc: bad args");
  if (b2c_precision())
    ncdhw_to_ndhwc_kernel<float><<<grid_for((long long)N * THW), kBlock, 0, (cudaStream_t)s>>>(in, (float*)out, N, C, THW, Cpad);
  else
    ncdhw_to_ndhwc_kernel<bf16><<<grid_for((long long)N * THW), kBlock, 0, (cudaStream_t)s>>>(in, (bf16*)out, N, C, THW, Cpad);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("ncdhw_to_ndhwc");
  return 0;
}

B2C_API int b2c_u8_clip_to_cl(const uint8_t* in, void* out, int32_t P, int32_t C, int64_t rows, int32_t W, b2c_stream_t s) {
  B2C_REQUIRE(in && out && P > 0 && C > 0 && C <= 8 && rows > 0 && W > 0, "u8_clip_to_cl: bad args");
  const long long total = (long long)P * rows * W;
  if (b2c_precision())
    u8_clip_to_cl_kernel<float><<<grid_for(total), kBlock, 0, (cudaStream_t)s>>>(in, (float*)out, P, C, rows, W);
  else
    u8_clip_to_cl_kernel<bf16><<<grid_for(total), kBlock, 0, (cudaStream_t)s>>>(in, (bf16*)out, P, C, rows, W);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("u8_clip_to_cl");
  return 0;
}

B2C_API int b2c_u8_to_f32(const uint8_t* in, float* out, int64_t n, float scale, b2c_stream_t s) {
  B2C_REQUIRE(in && out && n > 0, "u8_to_f32: bad args");
  u8_to_f32_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)s>>>(in, out, n, scale);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("u8_to_f32");
  return 0;
}

B2C_API int b2c_ndhwc_to_ncdhw_f32(const void* in, int64_t in_row_stride, int32_t in_c_off, float* out, int32_t N, int32_t C,
                                   int64_t THW, b2c_stream_t s) {
  B2C_REQUIRE(in && out && N > 0 && C > 0, "ndhwc_to_ncdhw: bad args");
  dim3 grid((unsigned)((THW + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)N), block(32, 8);
  if (b2c_precision())
    ndhwc_to_ncdhw_kernel<float><<<grid, block, 0, (cudaStream_t)s>>>((const float*)in, in_row_stride, in_c_off, out, N, C, THW);
  else
    ndhwc_to_ncdhw_kernel<bf16><<<grid, block, 0, (cudaStream_t)s>>>((const bf16*)in, in_row_stride, in_c_off, out, N, C, THW);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("ndhwc_to_ncdhw");
  return 0;
}

static int bn_stats_launch(const void* x, int64_t rows, int32_t C, int64_t row_stride, int32_t c_off, int32_t groups, float* ws,
                           BnFinalize F, b2c_stream_t s) {
  const long long rpg = rows / groups;
  dim3 grid((unsigned)(g_deterministic ? 1 : reduce_grid(rpg, C, groups)), (unsigned)groups);
  if (b2c_precision())
    bn_stats_kernel<float><<<grid, kBlock, bn_red_smem(C), (cudaStream_t)s>>>((const float*)x, rpg, C, row_stride, c_off, ws, F);
  else
    bn_stats_kernel<bf16><<<grid, kBlock, bn_red_smem(C), (cudaStream_t)s>>>((const bf16*)x, rpg, C, row_stride, c_off, ws, F);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("bn_sums");
  return 0;
}

B2C_API int b2c_bn_sums(const void* x, int64_t rows, int32_t C, int64_t row_stride, int32_t c_off, int32_t groups, float* ws,
                        b2c_stream_t s) {
  B2C_REQUIRE(x && ws, "bn_sums: null pointer");
  CHECK_VIEW("bn_sums", C, row_stride, c_off);
  B2C_REQUIRE(groups >= 1 && rows % groups == 0, "bn_sums: rows=%lld not divisible by groups=%d", (long long)rows, groups);
  return bn_stats_launch(x, rows, C, row_stride, c_off, groups, ws, BnFinalize{nullptr, nullptr, nullptr, nullptr, 0.f, 0.f, groups}, s);
}

B2C_API int b2c_bn_sums_finalize(const void* x, int64_t rows, int32_t C, int64_t row_stride, int32_t c_off, int32_t groups, float* ws,
                                 float* mean, float* rstd, float* running_mean, float* running_var, float momentum, float eps,
                                 b2c_stream_t s) {
  B2C_REQUIRE(x && ws && mean && rstd, "bn_sums_finalize: null pointer");
  CHECK_VIEW("bn_sums_finalize", C, row_stride, c_off);
  B2C_REQUIRE(groups >= 1 && rows % groups == 0, "bn_sums_finalize: rows=%lld not divisible by groups=%d", (long long)rows, groups);
  return bn_stats_launch(x, rows, C, row_stride, c_off, groups, ws, BnFinalize{mean, rstd, running_mean, running_var, momentum, eps, groups},
                         s);
}

B2C_API int b2c_bn_finalize(const float* ws, int32_t ws_C, int32_t c_off, int32_t C, int32_t groups, int64_t rows_per_group,
                            float* mean, float* rstd, float* running_mean, float* running_var, float momentum, float eps,
                            b2c_stream_t s) {
  B2C_REQUIRE(ws && mean && rstd && C > 0 && c_off >= 0 && c_off + C <= ws_C && groups >= 1 && rows_per_group > 0,
              "bn_finalize: bad args");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)s>>>(ws, ws_C, c_off, rows_per_group, C, groups, mean, rstd,
                                                                  running_mean, running_var, momentum, eps);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("bn_finalize");
  return 0;
}

B2C_API int b2c_bn_relu_apply(const void* x, int64_t rows, int32_t C, int64_t x_rs, int32_t x_co, int32_t groups, const float* mean,
                              const float* rstd, const float* gamma, const float* beta, void* y, int64_t y_rs, int32_t y_co,
                              int32_t relu, b2c_stream_t s) {
  B2C_REQUIRE(x && y && mean && rstd && gamma && beta, "bn_relu_apply: null pointer");
  CHECK_VIEW("bn_relu_apply", C, x_rs, x_co);
  CHECK_VIEW("bn_relu_apply(y)", C, y_rs, y_co);
  B2C_REQUIRE(groups >= 1 && rows % groups == 0, "bn_relu_apply: rows not divisible by groups");
  const long long rpg = rows / groups;
  dim3 grid((unsigned)apply_grid(rpg, C, groups), (unsigned)groups);
  LAUNCH_T(bn_relu_apply_kernel, grid, kBlock, 0, s, (const T*)x, rpg, C, x_rs, x_co, mean, rstd, gamma, beta, (T*)y, y_rs, y_co, relu);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("bn_relu_apply");
  return 0;
}

B2C_API int b2c_bn_relu_bwd_reduce(const void* dy, int64_t dy_rs, int32_t dy_co, const void* y, int64_t y_rs, int32_t y_co,
                                   const void* x, int64_t x_rs, int32_t x_co, int64_t rows, int32_t C, int32_t groups,
                                   const float* mean, const float* rstd, const float* gamma, const float* beta, float* ws,
                                   int32_t relu, b2c_stream_t s) {
  B2C_REQUIRE(dy && x && mean && rstd && ws && (y || !relu || (gamma && beta)), "bn_bwd_reduce: null pointer");
  CHECK_VIEW("bn_bwd_reduce(dy)", C, dy_rs, dy_co);
  CHECK_VIEW("bn_bwd_reduce(x)", C, x_rs, x_co);
  B2C_REQUIRE(groups >= 1 && rows % groups == 0, "bn_bwd_reduce: rows not divisible by groups");
  const long long rpg = rows / groups;
  dim3 grid((unsigned)(g_deterministic ? 1 : reduce_grid(rpg, C, groups)), (unsigned)groups);
#define B2C_BN_BWD_REDUCE(T, RM)                                                                                              \
  bn_bwd_reduce_kernel<T, RM><<<grid, kBlock, bn_red_smem(C), (cudaStream_t)s>>>((const T*)dy, dy_rs, dy_co, (const T*)y, y_rs, y_co, \
                                                                               (const T*)x, x_rs, x_co, rpg, C, mean, rstd, gamma,   \
                                                                               beta, ws, relu)
  const bool remask = relu && y == nullptr;
  if (b2c_precision()) {
    if (remask) B2C_BN_BWD_REDUCE(float, true);
    else B2C_BN_BWD_REDUCE(float, false);
  } else {
    if (remask) B2C_BN_BWD_REDUCE(bf16, true);
    else B2C_BN_BWD_REDUCE(bf16, false);
  }
#undef B2C_BN_BWD_REDUCE
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("bn_bwd_reduce");
  return 0;
}

B2C_API int b2c_bn_relu_bwd_apply(const void* dy, int64_t dy_rs, int32_t dy_co, const void* y, int64_t y_rs, int32_t y_co,
                                  const void* x, int64_t x_rs, int32_t x_co, int64_t rows, int32_t C, int32_t groups,
                                  const float* mean, const float* rstd, const float* gamma, const float* beta, const float* ws,
                                  void* dx, int64_t dx_rs, int32_t dx_co, float* dgamma, float* dbeta, int32_t relu, b2c_stream_t s) {
  B2C_REQUIRE(dy && x && dx && mean && rstd && gamma && ws && (y || !relu || beta), "bn_bwd_apply: null pointer");
  CHECK_VIEW("bn_bwd_apply(dy)", C, dy_rs, dy_co);
  CHECK_VIEW("bn_bwd_apply(dx)", C, dx_rs, dx_co);
  B2C_REQUIRE(groups >= 1 && rows % groups == 0, "bn_bwd_apply: rows not divisible by groups");
  const long long rpg = rows / groups;
  dim3 grid((unsigned)apply_grid(rpg, C, groups), (unsigned)groups);
#define B2C_BN_BWD_APPLY(T, RM)                                                                                              \
  bn_bwd_apply_kernel<T, RM><<<grid, kBlock, 0, (cudaStream_t)s>>>((const T*)dy, dy_rs, dy_co, (const T*)y, y_rs, y_co, (const T*)x, \
                                                                 x_rs, x_co, rpg, C, groups, mean, rstd, gamma, beta, ws, (T*)dx,    \
                                                                 dx_rs, dx_co, dgamma, dbeta, relu)
  const bool remask = relu && y == nullptr;
  if (b2c_precision()) {
    if (remask) B2C_BN_BWD_APPLY(float, true);
    else B2C_BN_BWD_APPLY(float, false);
  } else {
    if (remask) B2C_BN_BWD_APPLY(bf16, true);
    else B2C_BN_BWD_APPLY(bf16, false);
  }
#undef B2C_BN_BWD_APPLY
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("bn_bwd_apply");
  return 0;
}

B2C_API int b2c_maxpool_fwd(const void* x, int64_t x_rs, int32_t x_co, void* y, int64_t y_rs, int32_t y_co, uint8_t* idx, int32_t N,
                            int32_t C, int32_t Ti, int32_t Hi, int32_t Wi, int32_t To, int32_t Ho, int32_t Wo, int32_t kt, int32_t kh,
                            int32_t kw, int32_t st, int32_t sh, int32_t sw, int32_t pt, int32_t ph, int32_t pw, b2c_stream_t s) {
  B2C_REQUIRE(x && y && idx, "maxpool_fwd: null pointer");
  CHECK_VIEW("maxpool_fwd(x)", C, x_rs, x_co);
  CHECK_VIEW("maxpool_fwd(y)", C, y_rs, y_co);
  B2C_REQUIRE(kt * kh * kw < 255, "maxpool_fwd: window too large");
  PoolGeom G{N, C, Ti, Hi, Wi, To, Ho, Wo, kt, kh, kw, st, sh, sw, pt, ph, pw};
  const long long total = (long long)N * To * Ho * Wo * (C / 8);
  B2C_REQUIRE(total < (1LL << 31) - (1LL << 22) && (long long)N * Ti * Hi * Wi * (C / 8) < (1LL << 31), "maxpool_fwd: tensor too large");
  PoolRowGeom R;
  const int spec = pool_row_geom(R, N, C, Ti, Hi, Wi, To, Ho, Wo, kt, kh, kw, st, sh, sw, pt, ph, pw, Wo);
  const dim3 rgrid((unsigned)Ho, (unsigned)To, (unsigned)N);
#define B2C_POOL_FWD(KT_, KH_, KW_, ST_, SH_, SW_)                                                                                  \
  do {                                                                                                                             \
    if (b2c_precision())                                                                                                           \
      maxpool_fwd_row_kernel<float, KT_, KH_, KW_, ST_, SH_, SW_><<<rgrid, kBlock, 0, (cudaStream_t)s>>>((const float*)x, x_rs, x_co, \
                                                                                                         (float*)y, y_rs, y_co, idx, R); \
    else                                                                                                                           \
      maxpool_fwd_row_kernel<bf16, KT_, KH_, KW_, ST_, SH_, SW_><<<rgrid, kBlock, 0, (cudaStream_t)s>>>((const bf16*)x, x_rs, x_co,   \
                                                                                                        (bf16*)y, y_rs, y_co, idx, R);  \
  } while (0)
  if (spec == 1) B2C_POOL_FWD(1, 3, 3, 1, 2, 2);
  else if (spec == 2) B2C_POOL_FWD(3, 3, 3, 2, 1, 1);
  else if (spec == 3) B2C_POOL_FWD(3, 3, 3, 1, 1, 1);
  else if (b2c_precision())
    maxpool_fwd_f32_kernel<<<grid_for(total), kBlock, 0, (cudaStream_t)s>>>((const float*)x, x_rs, x_co, (float*)y, y_rs, y_co, idx, G);
  else
    maxpool_fwd_kernel<<<grid_for(total), kBlock, 0, (cudaStream_t)s>>>((const bf16*)x, x_rs, x_co, (bf16*)y, y_rs, y_co, idx, G);
#undef B2C_POOL_FWD
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("maxpool_fwd");
  return 0;
}

B2C_API int b2c_maxpool_bwd(const void* dy, int64_t dy_rs, int32_t dy_co, const uint8_t* idx, void* dx, int64_t dx_rs, int32_t dx_co,
                            int32_t N, int32_t C, int32_t Ti, int32_t Hi, int32_t Wi, int32_t To, int32_t Ho, int32_t Wo, int32_t kt,
                            int32_t kh, int32_t kw, int32_t st, int32_t sh, int32_t sw, int32_t pt, int32_t ph, int32_t pw,
                            int32_t accumulate, b2c_stream_t s) {
  B2C_REQUIRE(dy && dx && idx, "maxpool_bwd: null pointer");
  CHECK_VIEW("maxpool_bwd(dy)", C, dy_rs, dy_co);
  CHECK_VIEW("maxpool_bwd(dx)", C, dx_rs, dx_co);
  PoolGeom G{N, C, Ti, Hi, Wi, To, Ho, Wo, kt, kh, kw, st, sh, sw, pt, ph, pw};
  const long long total = (long long)N * Ti * Hi * Wi * (C / 8);
  B2C_REQUIRE(total < (1LL << 31) - (1LL << 22), "maxpool_bwd: tensor too large");
  PoolRowGeom R;
  const int spec = pool_row_geom(R, N, C, Ti, Hi, Wi, To, Ho, Wo, kt, kh, kw, st, sh, sw, pt, ph, pw, Wi);
  const dim3 rgrid((unsigned)Hi, (unsigned)Ti, (unsigned)N);
#define B2C_POOL_BWD(KT_, KH_, KW_, ST_, SH_, SW_)                                                                       \
  do {                                                                                                                  \
    if (b2c_precision())                                                                                                \
      maxpool_bwd_row_kernel<float, KT_, KH_, KW_, ST_, SH_, SW_><<<rgrid, kBlock, 0, (cudaStream_t)s>>>(                 \
          (const float*)dy, dy_rs, dy_co, idx, (float*)dx, dx_rs, dx_co, R, accumulate);                                \
    else                                                                                                                \
      maxpool_bwd_row_kernel<bf16, KT_, KH_, KW_, ST_, SH_, SW_><<<rgrid, kBlock, 0, (cudaStream_t)s>>>(                  \
          (const bf16*)dy, dy_rs, dy_co, idx, (bf16*)dx, dx_rs, dx_co, R, accumulate);                                  \
  } while (0)
  if (spec == 1) B2C_POOL_BWD(1, 3, 3, 1, 2, 2);
  else if (spec == 2) B2C_POOL_BWD(3, 3, 3, 2, 1, 1);
  else if (spec == 3) B2C_POOL_BWD(3, 3, 3, 1, 1, 1);
  else
    LAUNCH_T(maxpool_bwd_kernel, grid_for(total), kBlock, 0, s, (const T*)dy, dy_rs, dy_co, idx, (T*)dx, dx_rs, dx_co, G, accumulate);
#undef B2C_POOL_BWD
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("maxpool_bwd");
  return 0;
}

B2C_API int b2c_channel_scale(const void* x, int64_t x_rs, int32_t x_co, const float* scale_nc, void* y, int64_t y_rs, int32_t y_co,
                              int32_t N, int64_t rows_per_n, int32_t C, b2c_stream_t s) {
  B2C_REQUIRE(x && y && scale_nc, "channel_scale: null pointer");
  CHECK_VIEW("channel_scale(x)", C, x_rs, x_co);
  CHECK_VIEW("channel_scale(y)", C, y_rs, y_co);
  LAUNCH_T(channel_scale_kernel, grid_for((long long)N * rows_per_n * (C / 8)), kBlock, 0, s,  (const T*)x, x_rs, x_co, scale_nc, (T*)y, y_rs, y_co, N, rows_per_n, C);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("channel_scale");
  return 0;
}

B2C_API int b2c_act_bwd(const void* dy, int64_t dy_rs, int32_t dy_co, const void* y, int64_t y_rs, int32_t y_co, const float* scale_nc,
                        void* dz, int64_t dz_rs, int32_t dz_co, float* dbias, int32_t N, int64_t rows_per_n, int32_t C, int32_t relu,
                        b2c_stream_t s) {
  B2C_REQUIRE(dy && (y || !relu) && (dz || dbias), "act_bwd: null pointer");
  CHECK_VIEW("act_bwd(dy)", C, dy_rs, dy_co);
  if (dz) CHECK_VIEW("act_bwd(dz)", C, dz_rs, dz_co);
  int gx = row_grid(rows_per_n, C, 2);
  if (gx * N > b2c_num_sms() * 8) gx = (b2c_num_sms() * 8 + N - 1) / N;
  dim3 grid((unsigned)gx, (unsigned)N);
  LAUNCH_T(act_bwd_kernel, grid, kBlock, C * sizeof(float), s, (const T*)dy, dy_rs, dy_co, (const T*)y, y_rs, y_co, scale_nc, (T*)dz, dz_rs, dz_co, dbias, rows_per_n, C, relu);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("act_bwd");
  return 0;
}

B2C_API int b2c_add(const void* a, int64_t a_rs, int32_t a_co, const void* b, int64_t b_rs, int32_t b_co, void* out, int64_t o_rs,
                    int32_t o_co, int64_t rows, int32_t C, b2c_stream_t s) {
  B2C_REQUIRE(a && b && out, "add: null pointer");
  CHECK_VIEW("add(a)", C, a_rs, a_co);
  CHECK_VIEW("add(b)", C, b_rs, b_co);
  CHECK_VIEW("add(out)", C, o_rs, o_co);
  LAUNCH_T(add_kernel, grid_for(rows * (C / 8)), kBlock, 0, s, (const T*)a, a_rs, a_co, (const T*)b, b_rs, b_co, (T*)out, o_rs, o_co, rows, C);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("add");
  return 0;
}

B2C_API int b2c_stem_fold_input(const void* x, void* xs, int32_t N, int32_t T, int32_t H, int32_t W, int32_t Cs, int32_t pt,
                                int32_t Tp, b2c_stream_t s) {
  B2C_REQUIRE(x && xs && N > 0 && T > 0 && Cs == 8 && Tp > 0 && Tp % 2 == 0 && Tp <= 16 && pt >= 0 && pt + T <= Tp, "stem_fold_input: bad args");
  StemFold G{N, T, H, W, Cs, pt, Tp};
  const long long total = (long long)N * H * W * (Tp / 2);
  if (b2c_precision())
    stem_fold_input_kernel<float><<<grid_for(total), kBlock, 0, (cudaStream_t)s>>>((const float*)x, (float*)xs, G, total);
  else
    stem_fold_input_kernel<bf16><<<grid_for(total), kBlock, 0, (cudaStream_t)s>>>((const bf16*)x, (bf16*)xs, G, total);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("stem_fold_input");
  return 0;
}

B2C_API int b2c_clips_to_folded(const void* in, int32_t in_u8, void* xs, int32_t P, int32_t C, int32_t T, int32_t H, int32_t W, int32_t pt,
                                int32_t Tp, int32_t mirror, b2c_stream_t s) {
  B2C_REQUIRE(in && xs && P > 0 && C > 0 && C <= 4 && T > 0 && Tp > 0 && Tp % 2 == 0 && Tp <= 16 && pt >= 0 && pt + T <= Tp,
              "clips_to_folded: bad args");
  const long long total = (long long)P * H * (Tp / 2) * W;
  B2C_REQUIRE(total < (1LL << 31), "clips_to_folded: too many elements for the 32-bit index split");
  const int tf = b2c_precision();
  if (in_u8) {
    if (tf) clips_to_folded_kernel<float, uint8_t><<<grid_for(total), kBlock, 0, (cudaStream_t)s>>>((const uint8_t*)in, (float*)xs, P, C, T, H, W, pt, Tp, mirror, total);
    else clips_to_folded_kernel<bf16, uint8_t><<<grid_for(total), kBlock, 0, (cudaStream_t)s>>>((const uint8_t*)in, (bf16*)xs, P, C, T, H, W, pt, Tp, mirror, total);
  } else {
    if (tf) clips_to_folded_kernel<float, float><<<grid_for(total), kBlock, 0, (cudaStream_t)s>>>((const float*)in, (float*)xs, P, C, T, H, W, pt, Tp, mirror, total);
    else clips_to_folded_kernel<bf16, float><<<grid_for(total), kBlock, 0, (cudaStream_t)s>>>((const float*)in, (bf16*)xs, P, C, T, H, W, pt, Tp, mirror, total);
  }
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("clips_to_folded");
  return 0;
}

B2C_API int b2c_stem_fold_weights(const float* w, float* w2, int32_t Cout, int32_t Cin, int32_t kt, int32_t khw, int32_t st, int32_t To,
                                  int32_t Kf, b2c_stream_t s) {
  B2C_REQUIRE(w && w2 && Cout > 0 && Cin > 0 && Cin <= 4 && kt > 0 && khw > 0 && st > 0 && To > 0 && Kf % 4 == 0, "stem_fold_weights: bad args");
  stem_fold_weights_kernel<<<grid_for((long long)To * Cout * khw * Kf), kBlock, 0, (cudaStream_t)s>>>(w, w2, Cout, Cin, kt, khw, st, To, Kf);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("stem_fold_weights");
  return 0;
}

B2C_API int b2c_stem_unfold_wgrad(const float* dw2, float* dw, int32_t Cout, int32_t Cin, int32_t kt, int32_t khw, int32_t st, int32_t To,
                                  int32_t Kf, b2c_stream_t s) {
  B2C_REQUIRE(dw2 && dw && Cout > 0 && Cin > 0 && Cin <= 4 && kt > 0 && khw > 0 && st > 0 && To > 0 && Kf % 4 == 0, "stem_unfold_wgrad: bad args");
  stem_unfold_wgrad_kernel<<<grid_for((long long)Cout * Cin * kt * khw), kBlock, 0, (cudaStream_t)s>>>(dw2, dw, Cout, Cin, kt, khw, st, To, Kf);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("stem_unfold_wgrad");
  return 0;
}

B2C_API int b2c_split_bf16(const float* x, int64_t x_rs, int32_t x_co, void* hi, void* lo, int64_t rows, int32_t C, b2c_stream_t s) {
  B2C_REQUIRE(x && hi && lo && rows > 0, "split_bf16: bad args");
  CHECK_VIEW("split_bf16", C, x_rs, x_co);
  split_bf16_kernel<<<grid_for(rows * (C / 8)), kBlock, 0, (cudaStream_t)s>>>(x, x_rs, x_co, (bf16*)hi, (bf16*)lo, rows, C);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("split_bf16");
  return 0;
}

B2C_API int b2c_stencil27_fwd(const float* P, float* out, const float* bias, int32_t N, int32_t T, int32_t H, int32_t W,
                              b2c_stream_t s) {
  B2C_REQUIRE(P && out && N > 0 && (long long)N * T * H * W < (1LL << 31) - (1LL << 22), "stencil27_fwd: bad args");
  stencil27_fwd_kernel<<<grid_for((long long)N * T * H * W), kBlock, 0, (cudaStream_t)s>>>(P, out, bias, N, T, H, W);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("stencil27_fwd");
  return 0;
}

B2C_API int b2c_stencil27_bwd(const float* dout, void* dP, float* dbias, int32_t N, int32_t T, int32_t H, int32_t W, int32_t cpad,
                              b2c_stream_t s) {
  B2C_REQUIRE(dout && dP && N > 0 && cpad >= 32 && cpad % 8 == 0 && (long long)N * T * H * W < (1LL << 31) - (1LL << 22),
              "stencil27_bwd: bad args");
  if (b2c_precision())
    stencil27_bwd_kernel<float><<<grid_for((long long)N * T * H * W), kBlock, 0, (cudaStream_t)s>>>(dout, (float*)dP, dbias, N, T, H, W, cpad);
  else
    stencil27_bwd_kernel<bf16><<<grid_for((long long)N * T * H * W), kBlock, 0, (cudaStream_t)s>>>(dout, (bf16*)dP, dbias, N, T, H, W, cpad);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("stencil27_bwd");
  return 0;
}

B2C_API int b2c_fill_f32(float* p, int64_t n, float v, b2c_stream_t s) {
  B2C_REQUIRE(p || n == 0, "fill_f32: null pointer");
  if (n == 0) return 0;
  fill_f32_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)s>>>(p, n, v);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("fill_f32");
  return 0;
}

B2C_API int b2c_im2col_small(const void* x, void* out, int32_t N, int32_t Cs, int32_t C, int32_t T, int32_t H, int32_t W, int32_t To,
                             int32_t Ho, int32_t Wo, int32_t kt, int32_t kh, int32_t kw, int32_t st, int32_t sh, int32_t sw,
                             int32_t pt, int32_t ph, int32_t pw, int32_t Kpad, b2c_stream_t s) {
  B2C_REQUIRE(x && out && N > 0 && C > 0 && C <= Cs && C < 256, "im2col_small: bad args");
  const int K = kt * kh * kw * C;
  B2C_REQUIRE(Kpad % 64 == 0 && Kpad >= K && Cs == 8, "im2col_small: bad K / Cs");
  Im2colGeom G{N, Cs, C, T, H, W, To, Ho, Wo, kt, kh, kw, st, sh, sw, pt, ph, pw, K, Kpad};
  const int pitch = ((kIm2colWB - 1) * sw + kw) * C;
  const size_t zero_len = (size_t)(kIm2colWB - 1) * sw * C + 8;
  const size_t esz = b2c_precision() ? 4 : 2;
  const size_t smem = (((size_t)Kpad * 2 + 15) & ~(size_t)15) + ((size_t)kt * kh * pitch + zero_len) * esz;
  B2C_REQUIRE(smem <= 200 * 1024 && (size_t)kt * kh * pitch + zero_len < 65536 && Kpad / 8 <= 512, "im2col_small: footprint too large");
  const int strips = (Wo + kIm2colWB - 1) / kIm2colWB;
  const long long nblk = (long long)N * To * Ho * strips;
  const int vpr = Kpad / 8;
  const int threads = vpr * (320 / vpr > 0 ? 320 / vpr : 1);      // whole rows per sweep: 272 threads for Kpad = 1088
  B2C_REQUIRE(threads <= 512, "im2col_small: Kpad too large");
  if (b2c_precision()) {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(im2col_small_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return b2c_cuda_check(e, "im2col_small: cudaFuncSetAttribute");
      configured = true;
    }
    im2col_small_kernel<uint32_t><<<grid_for(nblk, 1, 8), threads, smem, (cudaStream_t)s>>>((const uint32_t*)x, (uint32_t*)out, G, strips, pitch);
  } else {
    B2C_REQUIRE(smem <= 48 * 1024, "im2col_small: footprint too large");
    im2col_small_kernel<uint16_t><<<grid_for(nblk, 1, 24), threads, smem, (cudaStream_t)s>>>((const uint16_t*)x, (uint16_t*)out, G, strips, pitch);
  }
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("im2col_small");
  return 0;
}

B2C_API int b2c_set_pool_generic(int32_t on) {
  g_pool_generic = on ? 1 : 0;
  return 0;
}

B2C_API int b2c_set_deterministic(int32_t on) {
  g_deterministic = on ? 1 : 0;
  return 0;
}
