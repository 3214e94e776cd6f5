// Capsule head: fused EM routing (ConvCaps with K=(1,1), capsules_ucf101.py:290-309; m_step :108-156,
// e_step :158-182; 3 iterations) forward and hand-derived backward, plus the small glue kernels around
// it (class-activation mean, pose masking, their adjoints).
//
// Parallelisation: one CTA (8 warps) per spatial location, persistent over locations.
//   lane  = output capsule type j (C <= 32 active lanes)
//   warp w owns input capsule types i = w, w+8, w+16, w+24   (B = 32)
// so every (i,j) vote V_ij (4x4) lives in the registers of exactly one thread for the whole routing:
// the 32x24x16 votes are never written to memory (the reference materialises them 60 times).
// Sums over j are warp shuffles; sums over i are 8-way cross-warp reductions through shared memory.
// The transformation matrices W (32,C,4,4) sit in shared memory as [i][h][lane] (conflict-free).
//
// cost_h_stdv (capsules_ucf101.py:144) squares a SUM of deviations that is analytically zero; the
// reference's fp32 value is rounding noise.  We evaluate that 24-term sum in fp64 so the result sits
// on the fp64 reference (the parity yardstick); its gradient is analytically zero.
#include "common.cuh"
#include "../../include/b200caps.h"
#include <stdlib.h>

long long b2c_launches_add(long long n);

namespace {

constexpr int kB = 32;        // input capsule types
#ifndef B2C_ROUTING_WARPS
#define B2C_ROUTING_WARPS 8
#endif
constexpr int kNW = B2C_ROUTING_WARPS;   // warps per CTA (8 or 16)
constexpr int kIPT = kB / kNW;  // i's per thread
constexpr int kRT = kNW * 32;
constexpr float kEps = 1e-8f;
constexpr float kLambda = 1e-6f;
constexpr float kHalfLn2Pi16 = 14.703016531274763f;  // 16 * 0.5 * ln(2*pi)

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Cross-warp sum of N per-lane values in two steps: every warp stores its partials, warp w then reduces quantities
// w, w+8, w+16 over the 8 warps and publishes them, and all warps read the N results back.  (The one-step version --
// every thread summing all 8 partials of all N quantities -- was 35 % of the kernel's instructions, ncu r01d.)
// `red` holds [8][17][32] partials followed by [17][32] results; the two barriers also order buffer reuse across calls.
template <int N>
__device__ __forceinline__ void block_sum(float* vals, float* red, int& pp, int w, int lane) {
  (void)pp;
  float* buf = red;
  float* res = red + kNW * 17 * 32;
#pragma unroll
  for (int q = 0; q < N; ++q) buf[(w * 17 + q) * 32 + lane] = vals[q];
  __syncthreads();
#pragma unroll
  for (int qq = 0; qq < (N + kNW - 1) / kNW; ++qq) {
    const int q = w + qq * kNW;
    if (q < N) {
      float s = 0.f;
#pragma unroll
      for (int ww = 0; ww < kNW; ++ww) s += buf[(ww * 17 + q) * 32 + lane];
      res[q * 32 + lane] = s;
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < N; ++q) vals[q] = res[q * 32 + lane];
}

__device__ __forceinline__ void load_W_smem(const float* __restrict__ W, float* sW, int C) {
  for (int idx = threadIdx.x; idx < kB * 16 * 32; idx += kRT) sW[idx] = 0.f;
  __syncthreads();
  for (int idx = threadIdx.x; idx < kB * C * 16; idx += kRT) {
    const int h = idx & 15;
    const int j = (idx >> 4) % C;
    const int i = (idx >> 4) / C;
    sW[(i * 16 + h) * 32 + j] = W[idx];
  }
}

__device__ __forceinline__ void compute_votes(const float* s_caps, const float* sW, int w, int lane, float (*V)[16]) {
#pragma unroll
  for (int k = 0; k < kIPT; ++k) {
    const int i = w + kNW * k;
    float M[16], Wr[16];
#pragma unroll
    for (int h = 0; h < 16; ++h) {
      M[h] = s_caps[i * 16 + h];
      Wr[h] = sW[(i * 16 + h) * 32 + lane];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) acc = fmaf(M[r * 4 + kk], Wr[kk * 4 + c], acc);
        V[k][r * 4 + c] = acc;
      }
  }
}

struct MStepOut {
  float R, T, a, inv_s;
};

// one M step: in r_prev[k]; out rn[k], Z[k], mu[16], S[16] and scalars.  All warps end with identical
// per-j values.
__device__ __forceinline__ MStepOut m_step(const float (*V)[16], const float* r_prev, const float* s_ain, const float* bu,
                                           float ba, int C, bool active, int w, int lane, float* red, int& pp, float* rn,
                                           float* Z, float* mu, float* S) {
  float acc[17];
#pragma unroll
  for (int q = 0; q < 17; ++q) acc[q] = 0.f;
#pragma unroll
  for (int k = 0; k < kIPT; ++k) {
    const int i = w + kNW * k;
    const float rp = active ? r_prev[k] * s_ain[i] : 0.f;
    Z[k] = warp_sum(rp) + kEps;
    rn[k] = rp / Z[k];
    acc[16] += rn[k];
#pragma unroll
    for (int h = 0; h < 16; ++h) acc[h] = fmaf(rn[k], V[k][h], acc[h]);
  }
  block_sum<17>(acc, red, pp, w, lane);
  MStepOut o;
  o.R = acc[16];
  const float invR = 1.f / (o.R + kEps);
#pragma unroll
  for (int h = 0; h < 16; ++h) mu[h] = acc[h] * invR;
  float sacc[16];
#pragma unroll
  for (int h = 0; h < 16; ++h) sacc[h] = 0.f;
#pragma unroll
  for (int k = 0; k < kIPT; ++k) {
    const float c = rn[k] * invR;
#pragma unroll
    for (int h = 0; h < 16; ++h) {
      const float dv = V[k][h] - mu[h];
      sacc[h] = fmaf(c, dv * dv, sacc[h]);
    }
  }
  block_sum<16>(sacc, red, pp, w, lane);
  float T = 0.f;
#pragma unroll
  for (int h = 0; h < 16; ++h) {
    S[h] = sacc[h] + kEps;
    T += bu[h] + 0.5f * logf(S[h]);
  }
  o.T = T;
  const float cost = T * o.R;
  // mean and the (analytically zero) sum of deviations, in fp64 over the C active lanes
  const double cd = active ? (double)cost : 0.0;
  const double md = warp_sum_d(cd) / (double)C;
  const double dev = active ? (cd - md) : 0.0;
  const double sdev = warp_sum_d(dev);
  const double stdv = sqrt(sdev * sdev / (double)C + (double)kEps);
  o.inv_s = (float)(1.0 / (stdv + (double)kEps));
  const float u = kLambda * (ba - ((float)md - cost) * o.inv_s);
  o.a = 1.f / (1.f + expf(-u));
  return o;
}

__device__ __forceinline__ void e_step(const float (*V)[16], const float* mu, const float* S, float a, bool active, float* r_out) {
  float inv2S[16], lnS = 0.f;
#pragma unroll
  for (int h = 0; h < 16; ++h) {
    inv2S[h] = 0.5f / S[h];
    lnS += logf(S[h]);
  }
  const float base = -0.5f * lnS - kHalfLn2Pi16 + logf(kEps + a);
#pragma unroll
  for (int k = 0; k < kIPT; ++k) {
    float q = 0.f;
#pragma unroll
    for (int h = 0; h < 16; ++h) {
      const float dv = V[k][h] - mu[h];
      q = fmaf(dv * dv, inv2S[h], q);
    }
    const float z = active ? (base - q) : -INFINITY;
    const float mx = warp_max(z);
    const float e = active ? expf(z - mx) : 0.f;
    r_out[k] = e / warp_sum(e);
  }
}

// =====================================================================================
__global__ void __launch_bounds__(kRT, kNW == 8 ? 2 : 1) em_routing_fwd_kernel(const float* __restrict__ caps, const float* __restrict__ W,
                                                                const float* __restrict__ beta_u, const float* __restrict__ beta_a,
                                                                float* __restrict__ out, long long b, int C) {
  extern __shared__ float sm[];
  float* sW = sm;                          // [32][16][32]
  float* red = sW + kB * 16 * 32;          // [2][8][17][32]
  float* s_caps = red + (kNW + 1) * 17 * 32; // [544]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool active = lane < C;
  load_W_smem(W, sW, C);
  float bu[16], ba = active ? beta_a[lane] : 0.f;
#pragma unroll
  for (int h = 0; h < 16; ++h) bu[h] = active ? beta_u[lane * 16 + h] : 0.f;
  const int ocols = C * 17;
  int pp = 0;
  for (long long loc = blockIdx.x; loc < b; loc += gridDim.x) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 544; idx += kRT) s_caps[idx] = caps[loc * 544 + idx];
    __syncthreads();
    float V[kIPT][16];
    compute_votes(s_caps, sW, w, lane, V);
    float r[kIPT], rn[kIPT], Z[kIPT], mu[16], S[16];
#pragma unroll
    for (int k = 0; k < kIPT; ++k) r[k] = 1.f / (float)C;
    MStepOut o;
#pragma unroll 1
    for (int t = 0; t < 3; ++t) {
      o = m_step(V, r, s_caps + 512, bu, ba, C, active, w, lane, red, pp, rn, Z, mu, S);
      if (t < 2) e_step(V, mu, S, o.a, active, r);
    }
    if (w == 0 && active) {
      float* dst = out + loc * ocols;
#pragma unroll
      for (int h = 0; h < 16; ++h) dst[lane * 16 + h] = mu[h];   // rows are C*17 floats: not 16B aligned for odd C
      dst[C * 16 + lane] = o.a;
    }
  }
}

// =====================================================================================
// Warp-per-location forward (r02).  The CTA-per-location kernel above spends its time in cross-warp reductions (four
// __syncthreads per M step) at ~5 % of the fp32 pipe (ncu r02a: 1.31 ms for 12 800 locations).  Here ONE WARP owns a
// location: lane = output capsule j, the 32 input capsules i are walked sequentially, so every sum over i is a private
// register accumulation and only the sums over j (the per-i normalisers) are warp shuffles.  Votes are recomputed where
// they are needed (64 FMAs) instead of being held in 64 registers per thread; the normalised assignments rn_ij of an
// iteration sit in 4 KB of shared memory per warp.  No block-level barrier inside the location loop.
//   pass A (per i): V_ij, E step of the previous iteration -> r_ij, rn_ij = r_ij a_i / Z_i, accumulate R_j, sum rn V
//   pass B (per i): V_ij again, accumulate the variance around the new mean
// Same formulas and evaluation order per j as m_step / e_step above.
// =====================================================================================

// Routing state saved by the training forward for the backward kernels (floats per location, 32-lane rows):
//   r_t[i][32] (t = 1, 2: the E-step assignments; t = 0 is the constant 1/C), rn_t[i][32] (t = 0..2), Z_t[i] (t = 0..2),
//   per-j scalars [t][R, T, a, 1/(stdv + eps)][32], mu_t[h][32], S_t[h][32]
// plus the rows the first backward kernel adds for the second (see em_routing_bwd_coef_kernel): it overwrites r_t with
// gz_{t-1} and fills X'_t, G'_t (t = 0..2), U_t (t = 0, 1) and one row of per-j scalars.  Everything the second kernel
// reads is the contiguous prefix [0, kStS).
constexpr int kStR = 0;
constexpr int kStRN = kStR + 2 * kB * 32;
constexpr int kStZ = kStRN + 3 * kB * 32;
constexpr int kStSC = kStZ + 3 * kB;
constexpr int kStMU = kStSC + 3 * 4 * 32;
constexpr int kStX = kStMU + 3 * 16 * 32;
constexpr int kStG = kStX + 3 * 16 * 32;
constexpr int kStU = kStG + 3 * 16 * 32;
constexpr int kStK = kStU + 2 * 16 * 32;
constexpr int kStS = kStK + 32;
constexpr int kStFloats = kStS + 3 * 16 * 32;     // 12800 floats = 51.2 KB per location
static_assert(kStS == 11264 && kStFloats == 12800, "routing state layout");

// warp maximum in ONE instruction: floats mapped to order-preserving unsigned keys, redux.sync.max.u32
__device__ __forceinline__ float warp_max_redux(float v) {
  uint32_t u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  u = __reduce_max_sync(0xffffffffu, u);
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(u);
}

__device__ __forceinline__ void votes_of(const float* __restrict__ sc, const float* __restrict__ sW, int i, int lane, float* V, int wst) {
  float M[16], Wr[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 m4 = *reinterpret_cast<const float4*>(sc + i * 16 + q * 4);
    M[q * 4 + 0] = m4.x; M[q * 4 + 1] = m4.y; M[q * 4 + 2] = m4.z; M[q * 4 + 3] = m4.w;
  }
#pragma unroll
  for (int h = 0; h < 16; ++h) Wr[h] = sW[(i * 16 + h) * wst + lane];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float acc = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) acc = fmaf(M[r * 4 + kk], Wr[kk * 4 + c], acc);
      V[r * 4 + c] = acc;
    }
}

// shared-memory row pitch of the per-j tables: 24 floats when C <= 24 (two CTAs per SM fit), else 32.  Lanes >= C read
// their neighbours' (valid) words and are masked out of every result.
__host__ __device__ __forceinline__ int routing_pitch(int C) { return C <= 24 ? 24 : 32; }

// reciprocal without the IEEE-division slow path: MUFU.RCP + one Newton step (<= 1 ulp for the normal-range operands here)
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * fmaf(-x, r, 2.f);
}

// two independent warp reductions with their shuffles interleaved (the location loops below walk TWO input capsules per
// iteration: a single capsule's chain  votes -> exponent -> max -> exp -> sum -> normaliser sum  is ~500 cycles of
// dependent latency, and with 12..16 warps per SM that chain, not the issue rate, set the kernel time)
__device__ __forceinline__ void warp_sum2(float& a, float& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ta = __shfl_xor_sync(0xffffffffu, a, o), tb = __shfl_xor_sync(0xffffffffu, b, o);
    a += ta;
    b += tb;
  }
}

template <int kW, int kPitch>
__global__ void __launch_bounds__(kW * 32, 1) em_routing_fwd_warp_kernel(const float* __restrict__ caps, const float* __restrict__ W,
                                                                        const float* __restrict__ beta_u, const float* __restrict__ beta_a,
                                                                        float* __restrict__ out, float* __restrict__ state, long long b,
                                                                        int C) {
  extern __shared__ float sm[];
  constexpr int wst = kPitch;                       // compile-time row pitch: shared-memory offsets become immediates
  float* sW = sm;                                   // [32][16][wst] (+ 8 floats of slack for the masked lanes' reads)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* sc = sW + kB * 16 * wst + 8 + w * 544;     // this warp's capsules (poses | activations), 16-byte aligned
  float* srn = sW + kB * 16 * wst + 8 + kW * 544 + w * (kB * wst + 8);   // this warp's rn[i][lane]
  const bool active = lane < C;
  for (int idx = threadIdx.x; idx < kB * 16 * wst + 8; idx += blockDim.x) sW[idx] = 0.f;
  __syncthreads();
  for (int idx = threadIdx.x; idx < kB * C * 16; idx += blockDim.x) {
    const int h = idx & 15, j = (idx >> 4) % C, i = (idx >> 4) / C;
    sW[(i * 16 + h) * wst + j] = W[idx];
  }
  __syncthreads();
  float bu_sum = 0.f;
  const float ba = active ? beta_a[lane] : 0.f;
#pragma unroll
  for (int h = 0; h < 16; ++h) bu_sum += active ? beta_u[lane * 16 + h] : 0.f;
  const int ocols = C * 17;
  for (long long loc = (long long)blockIdx.x * kW + w; loc < b; loc += (long long)gridDim.x * kW) {
    __syncwarp();
    for (int idx = lane; idx < 544 / 4; idx += 32)
      reinterpret_cast<float4*>(sc)[idx] = reinterpret_cast<const float4*>(caps + loc * 544)[idx];
    __syncwarp();
    const float* s_ain = sc + 512;
    float* st = state ? state + loc * kStFloats : nullptr;
    float mu[16], inv2S[16], base = 0.f, a_out = 0.f;
#pragma unroll 1
    for (int t = 0; t < 3; ++t) {
      // ---- pass A: assignments of this iteration and the weighted vote sums ----
      float A1[16], R = 0.f;
#pragma unroll
      for (int h = 0; h < 16; ++h) A1[h] = 0.f;
#pragma unroll 1
      for (int i = 0; i < kB; i += 2) {
        float Va[16], Vb[16];
        votes_of(sc, sW, i, lane, Va, wst);
        votes_of(sc, sW, i + 1, lane, Vb, wst);
        float ra, rb;
        if (t == 0) {
          ra = rb = 1.f / (float)C;
        } else {
          float qa = 0.f, qb = 0.f;
#pragma unroll
          for (int h = 0; h < 16; ++h) {
            const float da = Va[h] - mu[h], db = Vb[h] - mu[h];
            qa = fmaf(da * da, inv2S[h], qa);
            qb = fmaf(db * db, inv2S[h], qb);
          }
          const float za = active ? (base - qa) : -INFINITY, zb = active ? (base - qb) : -INFINITY;
          const float mxa = warp_max_redux(za), mxb = warp_max_redux(zb);
          const float ea = active ? __expf(za - mxa) : 0.f, eb = active ? __expf(zb - mxb) : 0.f;
          float sa = ea, sb = eb;
          warp_sum2(sa, sb);
          ra = ea * fast_rcp(sa);
          rb = eb * fast_rcp(sb);
        }
        const float rpa = active ? ra * s_ain[i] : 0.f, rpb = active ? rb * s_ain[i + 1] : 0.f;
        float Za = rpa, Zb = rpb;
        warp_sum2(Za, Zb);
        Za += kEps;
        Zb += kEps;
        const float rna = rpa * fast_rcp(Za), rnb = rpb * fast_rcp(Zb);
        if (active) {
          srn[i * wst + lane] = rna;
          srn[(i + 1) * wst + lane] = rnb;
        }
        if (st) {
          if (t > 0) {
            st[kStR + ((t - 1) * kB + i) * 32 + lane] = ra;
            st[kStR + ((t - 1) * kB + i + 1) * 32 + lane] = rb;
          }
          st[kStRN + (t * kB + i) * 32 + lane] = rna;
          st[kStRN + (t * kB + i + 1) * 32 + lane] = rnb;
          if (lane < 2) st[kStZ + t * kB + i + lane] = lane ? Zb : Za;
        }
        R += rna;
        R += rnb;
#pragma unroll
        for (int h = 0; h < 16; ++h) {
          A1[h] = fmaf(rna, Va[h], A1[h]);
          A1[h] = fmaf(rnb, Vb[h], A1[h]);
        }
      }
      const float invR = 1.f / (R + kEps);
#pragma unroll
      for (int h = 0; h < 16; ++h) mu[h] = A1[h] * invR;
      // ---- pass B: variance around the new mean ----
      float S[16];
#pragma unroll
      for (int h = 0; h < 16; ++h) S[h] = 0.f;
#pragma unroll 1
      for (int i = 0; i < kB; i += 2) {
        float Va[16], Vb[16];
        votes_of(sc, sW, i, lane, Va, wst);
        votes_of(sc, sW, i + 1, lane, Vb, wst);
        const float ca = srn[i * wst + lane] * invR, cb = srn[(i + 1) * wst + lane] * invR;
#pragma unroll
        for (int h = 0; h < 16; ++h) {
          const float da = Va[h] - mu[h], db = Vb[h] - mu[h];
          S[h] = fmaf(ca, da * da, S[h]);
          S[h] = fmaf(cb, db * db, S[h]);
        }
      }
      float lnS = 0.f;
#pragma unroll
      for (int h = 0; h < 16; ++h) {
        S[h] += kEps;
        lnS += logf(S[h]);
        inv2S[h] = 0.5f * fast_rcp(S[h]);
      }
      const float T = bu_sum + 0.5f * lnS;
      const float cost = T * R;
      const double cd = active ? (double)cost : 0.0;
      const double md = warp_sum_d(cd) / (double)C;
      const double dev = active ? (cd - md) : 0.0;
      const double sdev = warp_sum_d(dev);
      const double stdv = sqrt(sdev * sdev / (double)C + (double)kEps);
      const float inv_s = (float)(1.0 / (stdv + (double)kEps));
      const float u = kLambda * (ba - ((float)md - cost) * inv_s);
      a_out = 1.f / (1.f + expf(-u));
      base = -0.5f * lnS - kHalfLn2Pi16 + logf(kEps + a_out);
      if (st) {
        float* sc4 = st + kStSC + t * 4 * 32 + lane;
        sc4[0] = R; sc4[32] = T; sc4[64] = a_out; sc4[96] = inv_s;
#pragma unroll
        for (int h = 0; h < 16; ++h) {
          st[kStMU + (t * 16 + h) * 32 + lane] = mu[h];
          st[kStS + (t * 16 + h) * 32 + lane] = S[h];
        }
      }
    }
    if (active) {
      float* dst = out + loc * ocols;
#pragma unroll
      for (int h = 0; h < 16; ++h) dst[lane * 16 + h] = mu[h];
      dst[C * 16 + lane] = a_out;
    }
  }
}

// =====================================================================================
// kSaved: the per-iteration routing state comes from the training forward (em_routing_fwd_warp_kernel, `state`) instead
// of being recomputed here -- the recomputation was a third of this kernel (six block-wide reductions per location).
template <bool kSaved>
__global__ void __launch_bounds__(kRT, 1) em_routing_bwd_kernel(const float* __restrict__ caps, const float* __restrict__ W,
                                                                const float* __restrict__ beta_u, const float* __restrict__ beta_a,
                                                                const float* __restrict__ dout, const float* __restrict__ state,
                                                                float* __restrict__ dcaps,
                                                                float* __restrict__ dW, float* __restrict__ dbeta_u,
                                                                float* __restrict__ dbeta_a, long long b, int C) {
  extern __shared__ float sm[];
  float* sW = sm;                            // [32][16][32]
  float* sgW = sW + kB * 16 * 32;            // [32][16][32] per-CTA dW accumulator (owner-thread RMW, no atomics)
  float* red = sgW + kB * 16 * 32;           // [2][8][17][32]
  float* s_mu = red + (kNW + 1) * 17 * 32;     // [3][16][32]
  float* s_S = s_mu + 3 * 16 * 32;           // [3][16][32]
  float* s_gbu = s_S + 3 * 16 * 32;          // [17][32]  dbeta_u (16) + dbeta_a (1) accumulators (warp 0)
  float* s_caps = s_gbu + 17 * 32;           // [544]
  float* s_dout = s_caps + 544;              // [32*17]
  // per-iteration routing state of every thread ([t][which][k][thread]: conflict-free) and the per-j scalars.  Keeping
  // them in shared memory lets the iteration loops stay ROLLED: fully unrolled the kernel was 16 K instructions
  // (260 KB) and spent 31 % of its issue slots waiting for instruction fetch (ncu r01d, stall_no_inst).
  float* s_st = s_dout + 32 * 17;            // [3][3][kIPT][kRT]   rp, rn, Z
  float* s_sc = s_st + 3 * 3 * kIPT * kRT;   // [3][4][32]          R, T, a, inv_s
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool active = lane < C;
  load_W_smem(W, sW, C);
  for (int idx = threadIdx.x; idx < kB * 16 * 32; idx += kRT) sgW[idx] = 0.f;
  for (int idx = threadIdx.x; idx < 17 * 32; idx += kRT) s_gbu[idx] = 0.f;
  float bu[16], ba = active ? beta_a[lane] : 0.f;
#pragma unroll
  for (int h = 0; h < 16; ++h) bu[h] = active ? beta_u[lane * 16 + h] : 0.f;
  const int ocols = C * 17;
  int pp = 0;
  for (long long loc = blockIdx.x; loc < b; loc += gridDim.x) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 544; idx += kRT) s_caps[idx] = caps[loc * 544 + idx];
    for (int idx = threadIdx.x; idx < ocols; idx += kRT) s_dout[idx] = dout[loc * ocols + idx];
    __syncthreads();
    const float* s_ain = s_caps + 512;
    float V[kIPT][16];
    compute_votes(s_caps, sW, w, lane, V);

    if (kSaved) {
      // ---- per-iteration state of this location, saved by the training forward ----
      const float* stg = state + loc * kStFloats;
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        float* st = s_st + (size_t)t * 3 * kIPT * kRT + threadIdx.x;
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
          const int i = w + kNW * k;
          st[(0 * kIPT + k) * kRT] = t == 0 ? 1.f / (float)C : stg[kStR + ((t - 1) * kB + i) * 32 + lane];
          st[(1 * kIPT + k) * kRT] = stg[kStRN + (t * kB + i) * 32 + lane];
          st[(2 * kIPT + k) * kRT] = stg[kStZ + t * kB + i];
        }
      }
      for (int idx = threadIdx.x; idx < 3 * 4 * 32; idx += kRT) s_sc[idx] = stg[kStSC + idx];
      for (int idx = threadIdx.x; idx < 3 * 16 * 32; idx += kRT) {
        s_mu[idx] = stg[kStMU + idx];
        s_S[idx] = stg[kStS + idx];
      }
    } else {
      // ---- forward with the per-iteration state kept in shared memory ----
      float r[kIPT], rn[kIPT], Z[kIPT], mu[16], S[16];
#pragma unroll
      for (int k = 0; k < kIPT; ++k) r[k] = 1.f / (float)C;
#pragma unroll 1
      for (int t = 0; t < 3; ++t) {
        float* st = s_st + (size_t)t * 3 * kIPT * kRT + threadIdx.x;
#pragma unroll
        for (int k = 0; k < kIPT; ++k) st[(0 * kIPT + k) * kRT] = r[k];
        const MStepOut o = m_step(V, r, s_ain, bu, ba, C, active, w, lane, red, pp, rn, Z, mu, S);
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
          st[(1 * kIPT + k) * kRT] = rn[k];
          st[(2 * kIPT + k) * kRT] = Z[k];
        }
        if (w == 0) {
          float* sc = s_sc + t * 4 * 32 + lane;
          sc[0] = o.R; sc[32] = o.T; sc[64] = o.a; sc[96] = o.inv_s;
#pragma unroll
          for (int h = 0; h < 16; ++h) {
            s_mu[(t * 16 + h) * 32 + lane] = mu[h];
            s_S[(t * 16 + h) * 32 + lane] = S[h];
          }
        }
        if (t < 2) e_step(V, mu, S, o.a, active, r);
      }
    }
    __syncthreads();  // s_mu / s_S / s_sc visible to all warps

    // ---- backward sweep ----
    float gV[kIPT][16];
#pragma unroll
    for (int k = 0; k < kIPT; ++k)
#pragma unroll
      for (int h = 0; h < 16; ++h) gV[k][h] = 0.f;
    float gmu[16], gS[16], ga, gr[kIPT], gain[kIPT];
#pragma unroll
    for (int h = 0; h < 16; ++h) {
      gmu[h] = active ? s_dout[lane * 16 + h] : 0.f;
      gS[h] = 0.f;
    }
    ga = active ? s_dout[C * 16 + lane] : 0.f;
#pragma unroll
    for (int k = 0; k < kIPT; ++k) gr[k] = gain[k] = 0.f;

#pragma unroll 1
    for (int t = 2; t >= 0; --t) {
      const float* mu_t = s_mu + t * 16 * 32 + lane;   // stride 32 per h
      const float* S_t = s_S + t * 16 * 32 + lane;
      const float* st = s_st + (size_t)t * 3 * kIPT * kRT + threadIdx.x;
      const float* sc = s_sc + t * 4 * 32 + lane;
      float rp_t[kIPT], rn_t[kIPT], Z_t[kIPT];
#pragma unroll
      for (int k = 0; k < kIPT; ++k) {
        rp_t[k] = st[(0 * kIPT + k) * kRT];
        rn_t[k] = st[(1 * kIPT + k) * kRT];
        Z_t[k] = st[(2 * kIPT + k) * kRT];
      }
      const float R_t = sc[0], T_t = sc[32], a_t = sc[64], is_t = sc[96];
      if (t < 2) {
        // E-step backward: r^t = softmax_j(ln p_ij + ln(eps + a_j)) feeds iteration t+1
        float gz[kIPT];
        float acc[17];
#pragma unroll
        for (int q = 0; q < 17; ++q) acc[q] = 0.f;
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
          const float r = st[(3 * kIPT + k) * kRT];   // rp of iteration t+1 (next [t] block, which = 0)
          const float dot = warp_sum(gr[k] * r);
          gz[k] = r * (gr[k] - dot);
          acc[16] += gz[k];
        }
#pragma unroll
        for (int h = 0; h < 16; ++h) {
          const float m = mu_t[h * 32], invS = 1.f / S_t[h * 32];
          float gm = 0.f;
#pragma unroll
          for (int k = 0; k < kIPT; ++k) {
            const float dv = (V[k][h] - m) * invS;
            gm = fmaf(gz[k], dv, gm);
            gV[k][h] = fmaf(-gz[k], dv, gV[k][h]);
          }
          acc[h] = gm;
        }
        block_sum<17>(acc, red, pp, w, lane);
#pragma unroll
        for (int h = 0; h < 16; ++h) gmu[h] = acc[h];
        ga = acc[16] / (kEps + a_t);
#pragma unroll
        for (int h = 0; h < 16; ++h) {
          const float m = mu_t[h * 32], invS = 1.f / S_t[h * 32];
          float g = 0.f;
#pragma unroll
          for (int k = 0; k < kIPT; ++k) {
            const float dv = V[k][h] - m;
            g = fmaf(gz[k], 0.5f * invS * (dv * dv * invS - 1.f), g);
          }
          acc[h] = g;
        }
        block_sum<16>(acc, red, pp, w, lane);
#pragma unroll
        for (int h = 0; h < 16; ++h) gS[h] = acc[h];
      }
      // M-step backward
      const float R = R_t, invR = 1.f / (R + kEps);
      const float a = a_t;
      const float gu = active ? ga * a * (1.f - a) : 0.f;
      const float gcost = kLambda * is_t * (gu - warp_sum(gu) / (float)C);
      if (w == 0 && active) {
        s_gbu[16 * 32 + lane] += kLambda * gu;
#pragma unroll
        for (int h = 0; h < 16; ++h) s_gbu[h * 32 + lane] += gcost * R;
      }
      const float gR = gcost * T_t;
      const float one_m_csum = 1.f - R * invR;
      float gc[kIPT];
#pragma unroll
      for (int k = 0; k < kIPT; ++k) gc[k] = 0.f;
#pragma unroll
      for (int h = 0; h < 16; ++h) {
        const float m = mu_t[h * 32], Sv = S_t[h * 32];
        const float gSh = gS[h] + gcost * R * 0.5f / Sv;
        const float gmh = gmu[h] - 2.f * gSh * m * one_m_csum;
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
          const float dv = V[k][h] - m;
          gc[k] = fmaf(gSh, dv * dv, gc[k]);
          gc[k] = fmaf(gmh, V[k][h], gc[k]);
          const float c = rn_t[k] * invR;
          gV[k][h] = fmaf(c, gmh + 2.f * gSh * dv, gV[k][h]);
        }
      }
      float D[1] = {0.f};
#pragma unroll
      for (int k = 0; k < kIPT; ++k) D[0] = fmaf(gc[k], rn_t[k] * invR, D[0]);
      block_sum<1>(D, red, pp, w, lane);
      const float gR_tot = gR - D[0] * invR;
#pragma unroll
      for (int k = 0; k < kIPT; ++k) {
        const int i = w + kNW * k;
        const float grn = active ? (gc[k] * invR + gR_tot) : 0.f;
        const float dot2 = warp_sum(grn * rn_t[k]);
        const float grp = active ? (grn - dot2) / Z_t[k] : 0.f;
        gain[k] += warp_sum(grp * rp_t[k]);
        gr[k] = grp * s_ain[i];
      }
    }

    // ---- votes -> poses / W ----
#pragma unroll
    for (int k = 0; k < kIPT; ++k) {
      const int i = w + kNW * k;
      float M[16], Wr[16];
#pragma unroll
      for (int h = 0; h < 16; ++h) {
        M[h] = s_caps[i * 16 + h];
        Wr[h] = sW[(i * 16 + h) * 32 + lane];
      }
      // dM_i[r][kk] = sum_j (gV_ij W_ij^T)[r][kk]: 16 sums over the 32 lanes by recursive halving (16 shuffles instead of
      // 16 x 5): after the five steps lanes 2m and 2m+1 both hold element m
      float P[16];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          float g = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) g = fmaf(gV[k][r * 4 + c], Wr[kk * 4 + c], g);
          P[r * 4 + kk] = active ? g : 0.f;
        }
#pragma unroll
      for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
        const bool hi = (lane & bit) != 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (q < half) {
            const float send = hi ? P[q] : P[q + half];
            const float keep = hi ? P[q + half] : P[q];
            P[q] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
          }
        }
      }
      P[0] += __shfl_xor_sync(0xffffffffu, P[0], 1);
      if ((lane & 1) == 0) dcaps[loc * 544 + i * 16 + (lane >> 1)] = P[0];
      if (lane == 0) dcaps[loc * 544 + 512 + i] = gain[k];
      if (active) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float g = 0.f;
#pragma unroll
            for (int r = 0; r < 4; ++r) g = fmaf(M[r * 4 + kk], gV[k][r * 4 + c], g);
            sgW[(i * 16 + kk * 4 + c) * 32 + lane] += g;
          }
      }
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < kB * C * 16; idx += kRT) {
    const int h = idx & 15;
    const int j = (idx >> 4) % C;
    const int i = (idx >> 4) / C;
    atomicAdd(dW + idx, sgW[(i * 16 + h) * 32 + j]);
  }
  if (w == 0 && active) {
#pragma unroll
    for (int h = 0; h < 16; ++h) atomicAdd(dbeta_u + lane * 16 + h, s_gbu[h * 32 + lane]);
    atomicAdd(dbeta_a + lane, s_gbu[16 * 32 + lane]);
  }
}

// =====================================================================================
// Two-kernel backward on the saved state (r02d).  The CTA-per-location kernel above is latency bound (8 warps per SM,
// seven block-wide reductions per location, every warp recomputing the per-j terms: 2.5 ms for 12 800 locations).  The
// chain rule through the three EM iterations only couples the (i,j) pairs through per-j sums, and the sum
// D_j = sum_i gc_ij c_ij has a closed form (sum_i c_ij (V_ij - mu_j)^2 = S_j - eps, sum_i c_ij V_ij = mu_j), so:
//
//  1. em_routing_bwd_coef_kernel -- ONE WARP per location (lane = output capsule j, the 32 input capsules walked
//     sequentially with the votes recomputed, like the forward).  Walks t = 2, 1 and produces only scalars per (i,j):
//         gz^{t-1}_ij = r^t_ij a_i (grp_ij - sum_j grp_ij r^t_ij),   grp_ij = (grn_ij - sum_j grn_ij rn^t_ij) / Z^t_i,
//         grn_ij = (sum_h gS'_h dv^2 + gmu'_h V_h) / (R+eps) + gR_tot_j
//     (stored over r^t in the state), the activation gradient of iterations 2 and 1, d beta, and per-j vectors
//         X'_t = 2 gS'_t / (R_t+eps),  G'_t = gmu'_t / (R_t+eps)  (t = 2, 1, 0),   U_t = 1 / S_t  (t = 1, 0)
//     where gS', gmu' are the M-step gradients of iteration t (tests/routing_manual.py::backward_split is this
//     formulation in tensor ops, checked against autograd on the CPU).
//  2. em_routing_bwd_final_kernel -- thread = one (i,j) pair for all locations of the CTA, so the weight gradient
//     accumulates in registers.  Assembles the vote gradient of all three iterations in one pass,
//         gV_ij = sum_t rn^t_ij (G'_t + X'_t (V_ij - mu_t)) - sum_{t<2} gz^t_ij (V_ij - mu_t) U_t,
//     then dM_i = sum_j gV_ij W_ij^T (recursive-halving warp reduction), dW_ij += M_i^T gV_ij, and iteration 0's
//     activation-gradient term.  The 45 KB state prefix + capsules of a location arrive by cp.async.bulk into a
//     double-buffered stage.
// =====================================================================================
constexpr int kCoefVecRows = 4 * 16;   // per-warp shared vectors of the coefficient kernel: gS', gmu', mu_t, mu_{t-1}

// shared-memory load the compiler may not hoist out of a loop (the 16-warp coefficient kernel has 128 registers: hoisting
// the 64 loop-invariant per-j vector elements makes it spill)
template <bool kPin>
__device__ __forceinline__ float lds_vec(const float* p) {
  if (!kPin) return *p;
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(p)));
  return v;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int kNWc, int kPitch>
__global__ void __launch_bounds__(kNWc * 32, 1) em_routing_bwd_coef_kernel(const float* __restrict__ caps, const float* __restrict__ W,
                                                                      const float* __restrict__ dout, float* __restrict__ state,
                                                                      float* __restrict__ dcaps, float* __restrict__ dbeta_u,
                                                                      float* __restrict__ dbeta_a, long long b, int C) {
  extern __shared__ __align__(16) float sm[];
  constexpr int wst = kPitch;
  constexpr int nw = kNWc;
  constexpr bool kPin = kNWc >= 12;
  float* sW = sm;                                               // [32][16][wst] (+8)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* sc = sW + kB * 16 * wst + 8 + w * 544;                 // this warp's capsules
  float* vec = sW + kB * 16 * wst + 8 + nw * 544 + w * (kCoefVecRows * wst + 8);
  float* vgS = vec;                // [16][wst]
  float* vgm = vec + 16 * wst;
  float* vmu = vec + 32 * wst;
  float* vmp = vec + 48 * wst;     // mu of iteration t-1
  float* sred = sW + kB * 16 * wst + 8 + nw * 544 + nw * (kCoefVecRows * wst + 8);   // [nw][2][32]
  const bool active = lane < C;
  for (int idx = threadIdx.x; idx < kB * 16 * wst + 8; idx += blockDim.x) sW[idx] = 0.f;
  __syncthreads();
  for (int idx = threadIdx.x; idx < kB * C * 16; idx += blockDim.x) {
    const int h = idx & 15, j = (idx >> 4) % C, i = (idx >> 4) / C;
    sW[(i * 16 + h) * wst + j] = W[idx];
  }
  __syncthreads();
  const int ocols = C * 17;
  const float invC = 1.f / (float)C;
  float dbu_acc = 0.f, dba_acc = 0.f;
  for (long long loc = (long long)blockIdx.x * nw + w; loc < b; loc += (long long)gridDim.x * nw) {
    __syncwarp();
    for (int idx = lane; idx < 544 / 4; idx += 32)
      reinterpret_cast<float4*>(sc)[idx] = reinterpret_cast<const float4*>(caps + loc * 544)[idx];
    __syncwarp();
    const float* s_ain = sc + 512;
    float* st = state + loc * kStFloats;
    {  // the next location of this warp: pull the rows both loops will read (r, rn of iterations 1 and 2, Z, scalars, mu, S)
      const long long nl = loc + (long long)gridDim.x * nw;
      if (nl < b) {
        const float* sn = state + nl * kStFloats;
#pragma unroll
        for (int q = 0; q < 2; ++q) prefetch_l2(sn + kStR + (q * 32 + lane) * 32);
#pragma unroll
        for (int q = 1; q < 3; ++q) prefetch_l2(sn + kStRN + (q * 32 + lane) * 32);
        if (lane < 15) prefetch_l2(sn + kStZ + lane * 32);
        prefetch_l2(sn + kStMU + lane * 32);
        if (lane < 16) prefetch_l2(sn + kStMU + (32 + lane) * 32);
        prefetch_l2(sn + kStS + lane * 32);
        if (lane < 16) prefetch_l2(sn + kStS + (32 + lane) * 32);
      }
    }
    float gmu[16], gS[16], ga, gain_acc = 0.f;
#pragma unroll
    for (int h = 0; h < 16; ++h) {
      gmu[h] = active ? dout[loc * ocols + lane * 16 + h] : 0.f;
      gS[h] = 0.f;
    }
    ga = active ? dout[loc * ocols + C * 16 + lane] : 0.f;
#pragma unroll 1
    for (int t = 2; t >= 0; --t) {
      // ---- per-j M-step gradient of iteration t ----
      __syncwarp();
      const float* sc4 = st + kStSC + t * 4 * 32 + lane;
      const float R = sc4[0], T = sc4[32], a = sc4[64], is = sc4[96];
      const float invR = fast_rcp(R + kEps);
      const float gu = active ? ga * a * (1.f - a) : 0.f;
      const float gcost = kLambda * is * (gu - warp_sum(gu) * invC);
      dba_acc += kLambda * gu;
      dbu_acc += active ? gcost * R : 0.f;
      const float one_m_csum = 1.f - R * invR;
      float D = 0.f, Kc = 0.f;
#pragma unroll
      for (int h = 0; h < 16; ++h) {
        const float m = st[kStMU + (t * 16 + h) * 32 + lane], Sv = st[kStS + (t * 16 + h) * 32 + lane];
        const float gSh = fmaf(gcost * R * 0.5f, fast_rcp(Sv), gS[h]);
        const float gmh = gmu[h] - 2.f * gSh * m * one_m_csum;
        D = fmaf(gSh, Sv - kEps, D);
        D = fmaf(gmh, m, D);
        Kc = fmaf(gmh, m, Kc);
        st[kStX + (t * 16 + h) * 32 + lane] = active ? 2.f * gSh * invR : 0.f;
        st[kStG + (t * 16 + h) * 32 + lane] = active ? gmh * invR : 0.f;
        if (active) {    // rows are wst (24) floats apart: a masked lane would land in the next row
          vgS[h * wst + lane] = gSh;
          vgm[h * wst + lane] = gmh;
          vmu[h * wst + lane] = m;
        }
      }
      const float gR_tot = gcost * T - D * invR;
      if (t == 0) {
        st[kStK + lane] = active ? gR_tot : 0.f;
        break;
      }
      // ---- E step of iteration t-1 feeds r^t: per-j vectors of that iteration ----
#pragma unroll
      for (int h = 0; h < 16; ++h) {
        const float iS = fast_rcp(st[kStS + ((t - 1) * 16 + h) * 32 + lane]);
        st[kStU + ((t - 1) * 16 + h) * 32 + lane] = active ? iS : 0.f;
        if (active) vmp[h * wst + lane] = st[kStMU + ((t - 1) * 16 + h) * 32 + lane];
      }
      __syncwarp();
      float A[16], Bq[16], gzsum = 0.f;
#pragma unroll
      for (int h = 0; h < 16; ++h) A[h] = Bq[h] = 0.f;
      const float* p_rn = st + kStRN + t * kB * 32 + lane;
      float* p_r = st + kStR + (t - 1) * kB * 32 + lane;
      const float* p_Z = st + kStZ + t * kB;
      // two input capsules per iteration (independent chains interleaved, see warp_sum2); their state rows are loaded one
      // iteration ahead
      float rn_a = p_rn[0], r_a = p_r[0], Z_a = p_Z[0], rn_b = p_rn[32], r_b = p_r[32], Z_b = p_Z[1];
#pragma unroll 1
      for (int i = 0; i < kB; i += 2) {
        const float rna = rn_a, ra = r_a, Za = Z_a, rnb = rn_b, rb = r_b, Zb = Z_b;
        if (i + 2 < kB) {
          rn_a = p_rn[(i + 2) * 32];
          r_a = p_r[(i + 2) * 32];
          Z_a = p_Z[i + 2];
          rn_b = p_rn[(i + 3) * 32];
          r_b = p_r[(i + 3) * 32];
          Z_b = p_Z[i + 3];
        }
        float Va[16], Vb[16];
        votes_of(sc, sW, i, lane, Va, wst);
        votes_of(sc, sW, i + 1, lane, Vb, wst);
        float gca = Kc, gcb = Kc;
#pragma unroll
        for (int h = 0; h < 16; ++h) {
          const float m = lds_vec<kPin>(vmu + h * wst + lane), gs = lds_vec<kPin>(vgS + h * wst + lane),
                      gm = lds_vec<kPin>(vgm + h * wst + lane);
          const float da = Va[h] - m, db = Vb[h] - m;
          gca = fmaf(fmaf(gs, da, gm), da, gca);
          gcb = fmaf(fmaf(gs, db, gm), db, gcb);
        }
        const float grna = active ? fmaf(gca, invR, gR_tot) : 0.f, grnb = active ? fmaf(gcb, invR, gR_tot) : 0.f;
        float d2a = grna * rna, d2b = grnb * rnb;
        warp_sum2(d2a, d2b);
        const float grpa = active ? (grna - d2a) * fast_rcp(Za) : 0.f, grpb = active ? (grnb - d2b) * fast_rcp(Zb) : 0.f;
        float gsa = grpa * ra, gsb = grpb * rb;          // d a_in_i of this iteration; sum_j gr r = a_i gsum
        warp_sum2(gsa, gsb);
        if (lane == i) gain_acc += gsa;
        if (lane == i + 1) gain_acc += gsb;
        const float gza = ra * s_ain[i] * (grpa - gsa), gzb = rb * s_ain[i + 1] * (grpb - gsb);
        p_r[i * 32] = gza;
        p_r[(i + 1) * 32] = gzb;
        gzsum += gza;
        gzsum += gzb;
#pragma unroll
        for (int h = 0; h < 16; ++h) {
          const float mp = lds_vec<kPin>(vmp + h * wst + lane);
          const float da = Va[h] - mp, db = Vb[h] - mp;
          const float ua = gza * da, ub = gzb * db;
          A[h] += ua;
          Bq[h] = fmaf(ua, da, Bq[h]);
          A[h] += ub;
          Bq[h] = fmaf(ub, db, Bq[h]);
        }
      }
      const float a_prev = st[kStSC + (t - 1) * 4 * 32 + 64 + lane];
      ga = gzsum * fast_rcp(kEps + a_prev);
#pragma unroll
      for (int h = 0; h < 16; ++h) {
        const float iS = st[kStU + ((t - 1) * 16 + h) * 32 + lane];   // written above by this lane (0 for the masked lanes)
        gmu[h] = iS * A[h];
        gS[h] = 0.5f * iS * (iS * Bq[h] - gzsum);
      }
    }
    dcaps[loc * 544 + 512 + lane] = gain_acc;
  }
  sred[(w * 2 + 0) * 32 + lane] = dbu_acc;
  sred[(w * 2 + 1) * 32 + lane] = dba_acc;
  __syncthreads();
  if (w == 0 && active) {
    float su = 0.f, sa = 0.f;
    for (int ww = 0; ww < nw; ++ww) {
      su += sred[(ww * 2 + 0) * 32 + lane];
      sa += sred[(ww * 2 + 1) * 32 + lane];
    }
#pragma unroll
    for (int h = 0; h < 16; ++h) atomicAdd(dbeta_u + lane * 16 + h, su);
    atomicAdd(dbeta_a + lane, sa);
  }
}

constexpr int kFinStage = kStS + 544;      // floats per stage: state prefix + capsules
constexpr int kFinWarps = 16;              // thread (w, lane) owns the pairs (i = w, j = lane) and (i = w + 16, j = lane)

// the final backward kernel's two (i, j) pairs of a thread.  final_gv: vote gradients of all three iterations for both
// pairs (the eleven per-j vector elements of every pose component are loaded once and used for both); final_tail:
// iteration 0's activation gradient, pose gradient (warp reduction over j) and the weight-gradient accumulation of ONE
// pair
__device__ __forceinline__ void final_gv(const float* __restrict__ sv, int ia, int ib, int lane, bool active, float (&Va)[16],
                                         float (&Vb)[16], float& gc0a, float& gc0b, float& rn0a, float& rn0b) {
  rn0a = sv[kStRN + (0 * kB + ia) * 32 + lane];
  rn0b = sv[kStRN + (0 * kB + ib) * 32 + lane];
  const float rn1a = sv[kStRN + (1 * kB + ia) * 32 + lane], rn2a = sv[kStRN + (2 * kB + ia) * 32 + lane];
  const float rn1b = sv[kStRN + (1 * kB + ib) * 32 + lane], rn2b = sv[kStRN + (2 * kB + ib) * 32 + lane];
  const float gz0a = sv[kStR + (0 * kB + ia) * 32 + lane], gz1a = sv[kStR + (1 * kB + ia) * 32 + lane];
  const float gz0b = sv[kStR + (0 * kB + ib) * 32 + lane], gz1b = sv[kStR + (1 * kB + ib) * 32 + lane];
  gc0a = gc0b = 0.f;
#pragma unroll
  for (int h = 0; h < 16; ++h) {
    const float mu0 = sv[kStMU + (0 * 16 + h) * 32 + lane], mu1 = sv[kStMU + (1 * 16 + h) * 32 + lane],
                mu2 = sv[kStMU + (2 * 16 + h) * 32 + lane];
    const float X0 = sv[kStX + (0 * 16 + h) * 32 + lane], X1 = sv[kStX + (1 * 16 + h) * 32 + lane],
                X2 = sv[kStX + (2 * 16 + h) * 32 + lane];
    const float G0 = sv[kStG + (0 * 16 + h) * 32 + lane], G1 = sv[kStG + (1 * 16 + h) * 32 + lane],
                G2 = sv[kStG + (2 * 16 + h) * 32 + lane];
    const float U0 = sv[kStU + (0 * 16 + h) * 32 + lane], U1 = sv[kStU + (1 * 16 + h) * 32 + lane];
    const float hX0 = 0.5f * X0, G0mu = G0 * mu0;
    {
      const float d0 = Va[h] - mu0, d1 = Va[h] - mu1, d2 = Va[h] - mu2;
      gc0a = fmaf(fmaf(hX0, d0, G0), d0, gc0a + G0mu);
      float g = rn0a * G0;
      g = fmaf(rn1a, G1, g);
      g = fmaf(rn2a, G2, g);
      g = fmaf(fmaf(rn0a, X0, -gz0a * U0), d0, g);
      g = fmaf(fmaf(rn1a, X1, -gz1a * U1), d1, g);
      g = fmaf(rn2a * X2, d2, g);
      Va[h] = active ? g : 0.f;                              // V now holds gV
    }
    {
      const float d0 = Vb[h] - mu0, d1 = Vb[h] - mu1, d2 = Vb[h] - mu2;
      gc0b = fmaf(fmaf(hX0, d0, G0), d0, gc0b + G0mu);
      float g = rn0b * G0;
      g = fmaf(rn1b, G1, g);
      g = fmaf(rn2b, G2, g);
      g = fmaf(fmaf(rn0b, X0, -gz0b * U0), d0, g);
      g = fmaf(fmaf(rn1b, X1, -gz1b * U1), d1, g);
      g = fmaf(rn2b * X2, d2, g);
      Vb[h] = active ? g : 0.f;
    }
  }
}

__device__ __forceinline__ void final_tail(const float* __restrict__ sv, const float* __restrict__ scap, const float* __restrict__ sW,
                                           int i, int lane, int wst, bool active, float invC, float gRtot0, float gc0, float rn0,
                                           const float (&V)[16], float* __restrict__ dcl, float* __restrict__ dWacc) {
    // iteration 0's activation gradient: r^0 = 1/C
    const float grn = active ? gc0 + gRtot0 : 0.f;
    const float dot2 = warp_sum(grn * rn0);
    const float grp = active ? (grn - dot2) * fast_rcp(sv[kStZ + i]) : 0.f;
    const float gain0 = warp_sum(grp) * invC;
    if (lane == 0) dcl[512 + i] += gain0;
    float M[16], Wr[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 m4 = *reinterpret_cast<const float4*>(scap + i * 16 + q * 4);
      M[q * 4 + 0] = m4.x; M[q * 4 + 1] = m4.y; M[q * 4 + 2] = m4.z; M[q * 4 + 3] = m4.w;
    }
#pragma unroll
    for (int h = 0; h < 16; ++h) Wr[h] = sW[(i * 16 + h) * wst + lane];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        // owner-thread accumulator in shared memory ([component][thread]: conflict-free); in registers the two pairs'
        // 32 accumulators pushed the kernel past 128 registers (spills in the location loop)
        float g = dWacc[(kk * 4 + c) * (kFinWarps * 32)];
#pragma unroll
        for (int r = 0; r < 4; ++r) g = fmaf(M[r * 4 + kk], V[r * 4 + c], g);
        dWacc[(kk * 4 + c) * (kFinWarps * 32)] = g;
      }
    // dM_i[r][kk] = sum_j (gV_ij W_ij^T)[r][kk]: 16 sums over the 32 lanes by recursive halving
    float P[16];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float g = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) g = fmaf(V[r * 4 + c], Wr[kk * 4 + c], g);
        P[r * 4 + kk] = active ? g : 0.f;
      }
#pragma unroll
    for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
      const bool hi = (lane & bit) != 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q < half) {
          const float send = hi ? P[q] : P[q + half];
          const float keep = hi ? P[q + half] : P[q];
          P[q] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
      }
    }
    P[0] += __shfl_xor_sync(0xffffffffu, P[0], 1);
    if ((lane & 1) == 0) dcl[i * 16 + (lane >> 1)] = P[0];
}

template <int kPitch>
__global__ void __launch_bounds__(kFinWarps * 32, 1) em_routing_bwd_final_kernel(const float* __restrict__ caps, const float* __restrict__ W,
                                                                                 const float* __restrict__ state, float* __restrict__ dcaps,
                                                                                 float* __restrict__ dW, long long b, int C) {
  extern __shared__ __align__(128) float sm[];
  constexpr int wst = kPitch;
  float* stage0 = sm;                                   // [2][kFinStage]
  float* sW = sm + 2 * kFinStage;                       // [32][16][wst] (+8)
  uint64_t* full = reinterpret_cast<uint64_t*>(sW + kB * 16 * wst + 8);
  float* sdW = sW + kB * 16 * wst + 8 + 4;              // [2][16][512] weight-gradient accumulators of the threads' two pairs
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool active = lane < C;
  for (int idx = threadIdx.x; idx < kB * 16 * wst + 8; idx += blockDim.x) sW[idx] = 0.f;
  for (int idx = threadIdx.x; idx < 2 * 16 * kFinWarps * 32; idx += blockDim.x) sdW[idx] = 0.f;
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < kB * C * 16; idx += blockDim.x) {
    const int h = idx & 15, j = (idx >> 4) % C, i = (idx >> 4) / C;
    sW[(i * 16 + h) * wst + j] = W[idx];
  }
  constexpr uint32_t kStBytes = kStS * 4, kCapBytes = 544 * 4;
  if (threadIdx.x == 0 && (long long)blockIdx.x < b) {
    mbar_arrive_expect_tx(&full[0], kStBytes + kCapBytes);
    bulk_g2s(smem_u32(stage0), state + (long long)blockIdx.x * kStFloats, kStBytes, &full[0]);
    bulk_g2s(smem_u32(stage0 + kStS), caps + (long long)blockIdx.x * 544, kCapBytes, &full[0]);
  }
  __syncthreads();
  float* dWacc0 = sdW + threadIdx.x;
  float* dWacc1 = sdW + 16 * kFinWarps * 32 + threadIdx.x;
  const float invC = 1.f / (float)C;
  int it = 0;
  for (long long loc = blockIdx.x; loc < b; loc += gridDim.x, ++it) {
    const int s = it & 1;
    const long long nxt = loc + gridDim.x;
    if (threadIdx.x == 0 && nxt < b) {   // stage s^1 was released by the barrier that ended the previous iteration
      float* dst = stage0 + (s ^ 1) * kFinStage;
      mbar_arrive_expect_tx(&full[s ^ 1], kStBytes + kCapBytes);
      bulk_g2s(smem_u32(dst), state + nxt * kStFloats, kStBytes, &full[s ^ 1]);
      bulk_g2s(smem_u32(dst + kStS), caps + nxt * 544, kCapBytes, &full[s ^ 1]);
    }
    mbar_wait(&full[s], (uint32_t)((it >> 1) & 1), 900 + s);
    const float* sv = stage0 + s * kFinStage;
    const float* scap = sv + kStS;
    const float gRtot0 = sv[kStK + lane];
    {
      float Va[16], Vb[16], gc0a, gc0b, rn0a, rn0b;
      votes_of(scap, sW, w, lane, Va, wst);
      votes_of(scap, sW, w + kFinWarps, lane, Vb, wst);
      final_gv(sv, w, w + kFinWarps, lane, active, Va, Vb, gc0a, gc0b, rn0a, rn0b);
      final_tail(sv, scap, sW, w, lane, wst, active, invC, gRtot0, gc0a, rn0a, Va, dcaps + loc * 544, dWacc0);
      asm volatile("" ::: "memory");   // keep the two tails sequential (register pressure)
      final_tail(sv, scap, sW, w + kFinWarps, lane, wst, active, invC, gRtot0, gc0b, rn0b, Vb, dcaps + loc * 544, dWacc1);
    }
    __syncthreads();   // every thread is done with stage s before it is refilled (two iterations ahead)
  }
  if (active) {
#pragma unroll
    for (int h = 0; h < 16; ++h) {
      atomicAdd(dW + ((long long)w * C + lane) * 16 + h, dWacc0[h * (kFinWarps * 32)]);
      atomicAdd(dW + ((long long)(w + kFinWarps) * C + lane) * 16 + h, dWacc1[h * (kFinWarps * 32)]);
    }
  }
}

// =====================================================================================
// glue: rout (N, L, C*16 + C) fp32
__global__ void __launch_bounds__(256) class_mean_kernel(const float* __restrict__ rout, float* __restrict__ act, int L, int C) {
  // one CTA per clip: warp w sums the locations w, w + 8, ... (lane = class), then the 8 partials meet in shared memory.
  // (One warp walking all 400 locations serially took 104 us in the captured step: 400 dependent-latency loads.)
  // reference: mean over h then mean over w (capsules_ucf101.py:450-451) == mean over L for a full grid
  const int n = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int ocols = C * 17;
  __shared__ float part[8][32];
  float s = 0.f;
  if (lane < C)
    for (int l = w; l < L; l += 8) s += rout[((long long)n * L + l) * ocols + C * 16 + lane];
  part[w][lane] = s;
  __syncthreads();
  if (w == 0 && lane < C) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += part[q][lane];
    act[n * C + lane] = t / (float)L;
  }
}

// store 8 activation values as the current precision mode's storage type (fp32: rounded to tf32 -- GEMM operand)
template <typename T>
__device__ __forceinline__ void store8_act(T* p, const float* v);
template <>
__device__ __forceinline__ void store8_act<bf16>(bf16* p, const float* v) { *reinterpret_cast<uint4*>(p) = pack8(v); }
template <>
__device__ __forceinline__ void store8_act<float>(float* p, const float* v) {
  reinterpret_cast<float4*>(p)[0] = make_float4(tf32_rna(v[0]), tf32_rna(v[1]), tf32_rna(v[2]), tf32_rna(v[3]));
  reinterpret_cast<float4*>(p)[1] = make_float4(tf32_rna(v[4]), tf32_rna(v[5]), tf32_rna(v[6]), tf32_rna(v[7]));
}

template <typename T>
__global__ void pose_mask_kernel(const float* __restrict__ rout, const float* __restrict__ mask, T* __restrict__ x, int L, int C,
                                 long long total) {
  const int ocols = C * 17, pc = C * 16;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % (pc / 8));
    const long long row = i / (pc / 8);
    const int n = (int)(row / L);
    const float* src = rout + row * ocols + e * 8;
    const float m = mask[n * C + (e * 8) / 16];
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = src[q] * m;
    store8_act(x + row * pc + e * 8, v);
  }
}

template <typename T>
__global__ void caps_head_bwd_kernel(const T* __restrict__ dx, const float* __restrict__ mask, const float* __restrict__ dact,
                                     const float* __restrict__ dfeat, float* __restrict__ drout, int L, int C, long long rows) {
  const int ocols = C * 17, pc = C * 16;
  const long long total = rows * ocols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % ocols);
    const long long row = i / ocols;
    const int n = (int)(row / L);
    float v;
    if (col < pc) {
      v = dx ? (float)dx[row * pc + col] * mask[n * C + col / 16] : 0.f;
    } else {
      const int j = col - pc;
      v = (dact ? dact[n * C + j] / (float)L : 0.f) + (dfeat ? dfeat[row * C + j] : 0.f);
    }
    drout[i] = v;
  }
}

// PrimaryCaps backward prologue: g fp32 (rows, 544) is the gradient w.r.t. [poses | sigmoid(act)];
// dz = g * (col >= 512 ? a (1 - a) : 1) as bf16 rows for the dgrad / wgrad GEMMs, dbias[col] += sum_rows dz.
template <typename T>
__global__ void __launch_bounds__(256) primarycaps_bwd_prep_kernel(const float* __restrict__ g, const float* __restrict__ out,
                                                                   T* __restrict__ dz, float* __restrict__ dbias, long long rows, int dz_pitch,
                                                                   T* __restrict__ dz_rows, int Nc, int Hq, int Wq) {
  // block = 256 threads: 4 row lanes x 68 column groups of 8 (544 = 68 * 8); threads >= 272 idle
  const int cg = threadIdx.x % 68, rl = threadIdx.x / 68;
  const bool act = rl < 3;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (act) {
    for (long long r = (long long)blockIdx.x * 3 + rl; r < rows; r += (long long)gridDim.x * 3) {
      const float* gp = g + r * 544 + cg * 8;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = gp[j];
      if (cg >= 64) {
        const float* op = out + r * 544 + cg * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = op[j];
          v[j] *= a * (1.f - a);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
      store8_act(dz + r * dz_pitch + cg * 8, v);
      if (dz_rows) {       // second copy with image rows outermost (rows-major dgrad)
        const int L = Hq * Wq;
        const int n = (int)(r / L), l = (int)(r - (long long)n * L);
        const int h = l / Wq, w = l - h * Wq;
        store8_act(dz_rows + ((long long)(h * Wq + w) * Nc + n) * dz_pitch + cg * 8, v);
      }
    }
  }
  __shared__ float sh[544];
  for (int i = threadIdx.x; i < 544; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  if (act) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sh[cg * 8 + j], acc[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 544; i += blockDim.x) atomicAdd(&dbias[i], sh[i]);
}

// PrimaryCaps forward epilogue after a K-split GEMM: sum of the K slices' partial outputs (frames of `part`) + bias, sigmoid on
// the activation columns (>= 512)
__global__ void __launch_bounds__(256) primarycaps_finish_kernel(const float* __restrict__ part, int nslice, const float* __restrict__ bias,
                                                                 float* __restrict__ out, int L, long long total4) {
  const long long per = (long long)L * 136;     // column quads per (clip, slice): 544 / 4 per location
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / per, r = i - n * per;
    const int c4 = (int)(r % 136);
    const float4* src = reinterpret_cast<const float4*>(part) + n * nslice * per + r;
    float4 v = src[0];
    for (int sl = 1; sl < nslice; ++sl) {
      const float4 u = src[(long long)sl * per];
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    const float4 b = reinterpret_cast<const float4*>(bias)[c4];
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    if (c4 >= 128) {
      v.x = sigmoidf_(v.x); v.y = sigmoidf_(v.y); v.z = sigmoidf_(v.z); v.w = sigmoidf_(v.w);
    }
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

}  // namespace

B2C_API int b2c_primarycaps_finish(const float* part, int32_t nslice, const float* bias, float* out, int32_t N, int32_t L,
                                   b2c_stream_t s) {
  B2C_REQUIRE(part && out && bias && nslice >= 1 && N > 0 && L > 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)part & 15) == 0 &&
                  ((uintptr_t)bias & 15) == 0,
              "primarycaps_finish: bad args");
  const long long total4 = (long long)N * L * 136;
  long long blocks = (total4 + 255) / 256;
  const long long cap = (long long)b2c_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  primarycaps_finish_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(part, nslice, bias, out, L, total4);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("primarycaps_finish");
  return 0;
}

static int routing_fwd_impl(const float* caps, const float* W, const float* beta_u, const float* beta_a, float* out, float* state,
                            int64_t b, int32_t C, b2c_stream_t s) {
  B2C_REQUIRE(caps && W && beta_u && beta_a && out, "em_routing_fwd: null pointer");
  B2C_REQUIRE(C >= 1 && C <= 32, "em_routing_fwd: C=%d must be in [1,32]", C);
  if (b <= 0) return 0;
  // default: warp-per-location kernel (r02); B2C_ROUTING=cta selects the CTA-per-location kernel of round 1
  static int use_warp = -1;
  if (use_warp < 0) {
    const char* e = getenv("B2C_ROUTING");
    use_warp = (e && e[0] == 'c') ? 0 : 1;
  }
  if (use_warp || state) {
    const int wst = routing_pitch(C);
    static int wenv = -1;
    if (wenv < 0) {
      const char* e = getenv("B2C_ROUTING_FWD_WARPS");
      wenv = e ? atoi(e) : 0;
    }
    // one CTA per SM; measured at 12 800 locations, C = 24: 8 warps 1.01 ms, 12 warps 0.885 ms, 16 warps 0.80 ms
    const int kw = C > 24 ? 8 : (wenv == 12 || wenv == 8 ? wenv : 16);
    const size_t smw = (size_t)(kB * 16 * wst + 8 + kw * 544 + kw * (kB * wst + 8)) * sizeof(float);
    long long gridw = b2c_num_sms();
    if (gridw * kw > b) gridw = (b + kw - 1) / kw;
#define B2C_RFWD(KW_, P_)                                                                                                          \
  do {                                                                                                                             \
    static bool cfg_ = false;                                                                                                      \
    if (!cfg_) {                                                                                                                   \
      cudaError_t e = cudaFuncSetAttribute(em_routing_fwd_warp_kernel<KW_, P_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      if (e != cudaSuccess) return b2c_cuda_check(e, "em_routing_fwd(warp) attr");                                                 \
      cfg_ = true;                                                                                                                 \
    }                                                                                                                              \
    em_routing_fwd_warp_kernel<KW_, P_><<<(unsigned)gridw, KW_ * 32, smw, (cudaStream_t)s>>>(caps, W, beta_u, beta_a, out, state, b, C); \
  } while (0)
    if (wst == 32) B2C_RFWD(8, 32);
    else if (kw == 16) B2C_RFWD(16, 24);
    else if (kw == 12) B2C_RFWD(12, 24);
    else B2C_RFWD(8, 24);
#undef B2C_RFWD
    b2c_launches_add(1);
    B2C_LAUNCH_CHECK("em_routing_fwd(warp)");
    return 0;
  }
  const size_t smem = (size_t)(kB * 16 * 32 + (kNW + 1) * 17 * 32 + 544) * sizeof(float);
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(em_routing_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return b2c_cuda_check(e, "em_routing_fwd attr");
    cfg = true;
  }
  long long grid = (kNW == 8 ? 2LL : 1LL) * b2c_num_sms();
  if (grid > b) grid = b;
  em_routing_fwd_kernel<<<(unsigned)grid, kRT, smem, (cudaStream_t)s>>>(caps, W, beta_u, beta_a, out, b, C);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("em_routing_fwd");
  return 0;
}

B2C_API int b2c_em_routing_fwd(const float* caps, const float* W, const float* beta_u, const float* beta_a, float* out, int64_t b,
                               int32_t C, b2c_stream_t s) {
  return routing_fwd_impl(caps, W, beta_u, beta_a, out, nullptr, b, C, s);
}

B2C_API int64_t b2c_em_routing_state_floats(void) { return kStFloats; }

B2C_API int b2c_em_routing_fwd_train(const float* caps, const float* W, const float* beta_u, const float* beta_a, float* out,
                                     float* state, int64_t b, int32_t C, b2c_stream_t s) {
  B2C_REQUIRE(state, "em_routing_fwd_train: null state buffer");
  return routing_fwd_impl(caps, W, beta_u, beta_a, out, state, b, C, s);
}

static int routing_bwd_impl(const float* caps, const float* W, const float* beta_u, const float* beta_a, const float* dout,
                            float* state, float* dcaps, float* dW, float* dbeta_u, float* dbeta_a, int64_t b, int32_t C,
                            b2c_stream_t s) {
  B2C_REQUIRE(caps && W && beta_u && beta_a && dout && dcaps && dW && dbeta_u && dbeta_a, "em_routing_bwd: null pointer");
  B2C_REQUIRE(C >= 1 && C <= 32, "em_routing_bwd: C=%d must be in [1,32]", C);
  if (b <= 0) return 0;
  // saved state: two-kernel backward (r02d); B2C_ROUTING_BWD=cta selects the CTA-per-location kernel on the same state
  static int split = -1;
  if (split < 0) {
    const char* e = getenv("B2C_ROUTING_BWD");
    split = (e && e[0] == 'c') ? 0 : 1;
  }
  if (state && split) {
    const int wst = routing_pitch(C);
    static int nw_env = -1;
    if (nw_env < 0) {
      const char* e = getenv("B2C_ROUTING_COEF_WARPS");
      nw_env = e ? atoi(e) : 0;
    }
    // measured (12 800 locations, C = 24): 16 warps / 128 registers 0.68 ms, 12 warps / 168 registers 0.51 ms
    const int nw = C > 24 ? 8 : (nw_env == 16 || nw_env == 8 ? nw_env : 12);
    const size_t sm1 = (size_t)(kB * 16 * wst + 8 + nw * 544 + nw * (kCoefVecRows * wst + 8) + nw * 64) * sizeof(float);
    const size_t sm2 = (size_t)(2 * kFinStage + kB * 16 * wst + 8 + 4 + 2 * 16 * kFinWarps * 32) * sizeof(float);
    long long g1 = b2c_num_sms();
    if (g1 * nw > b) g1 = (b + nw - 1) / nw;
    long long g2 = b2c_num_sms();
    if (g2 > b) g2 = b;
#define B2C_RBWD(NW_, P_)                                                                                                          \
  do {                                                                                                                             \
    static bool cfg_ = false;                                                                                                      \
    if (!cfg_) {                                                                                                                   \
      cudaError_t e = cudaFuncSetAttribute(em_routing_bwd_coef_kernel<NW_, P_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      if (e == cudaSuccess)                                                                                                        \
        e = cudaFuncSetAttribute(em_routing_bwd_final_kernel<P_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);        \
      if (e != cudaSuccess) return b2c_cuda_check(e, "em_routing_bwd(split) attr");                                                \
      cfg_ = true;                                                                                                                 \
    }                                                                                                                              \
    em_routing_bwd_coef_kernel<NW_, P_><<<(unsigned)g1, NW_ * 32, sm1, (cudaStream_t)s>>>(caps, W, dout, state, dcaps, dbeta_u,     \
                                                                                          dbeta_a, b, C);                        \
    B2C_LAUNCH_CHECK("em_routing_bwd(coef)");                                                                                      \
    em_routing_bwd_final_kernel<P_><<<(unsigned)g2, kFinWarps * 32, sm2, (cudaStream_t)s>>>(caps, W, state, dcaps, dW, b, C);       \
  } while (0)
    if (wst == 32) B2C_RBWD(8, 32);
    else if (nw == 16) B2C_RBWD(16, 24);
    else if (nw == 8) B2C_RBWD(8, 24);
    else B2C_RBWD(12, 24);
#undef B2C_RBWD
    b2c_launches_add(2);
    B2C_LAUNCH_CHECK("em_routing_bwd(final)");
    return 0;
  }
  const size_t smem = (size_t)(2 * kB * 16 * 32 + (kNW + 1) * 17 * 32 + 2 * 3 * 16 * 32 + 17 * 32 + 544 + 32 * 17 + 3 * 3 * kIPT * kRT +
                               3 * 4 * 32) * sizeof(float);
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(em_routing_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(em_routing_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return b2c_cuda_check(e, "em_routing_bwd attr");
    cfg = true;
  }
  long long grid = b2c_num_sms();
  if (grid > b) grid = b;
  if (state)
    em_routing_bwd_kernel<true><<<(unsigned)grid, kRT, smem, (cudaStream_t)s>>>(caps, W, beta_u, beta_a, dout, state, dcaps, dW, dbeta_u,
                                                                               dbeta_a, b, C);
  else
    em_routing_bwd_kernel<false><<<(unsigned)grid, kRT, smem, (cudaStream_t)s>>>(caps, W, beta_u, beta_a, dout, nullptr, dcaps, dW,
                                                                                dbeta_u, dbeta_a, b, C);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("em_routing_bwd");
  return 0;
}

B2C_API int b2c_em_routing_bwd(const float* caps, const float* W, const float* beta_u, const float* beta_a, const float* dout,
                               float* dcaps, float* dW, float* dbeta_u, float* dbeta_a, int64_t b, int32_t C, b2c_stream_t s) {
  return routing_bwd_impl(caps, W, beta_u, beta_a, dout, nullptr, dcaps, dW, dbeta_u, dbeta_a, b, C, s);
}

B2C_API int b2c_em_routing_bwd_state(const float* caps, const float* W, const float* beta_u, const float* beta_a, const float* dout,
                                     float* state, float* dcaps, float* dW, float* dbeta_u, float* dbeta_a, int64_t b, int32_t C,
                                     b2c_stream_t s) {
  B2C_REQUIRE(state, "em_routing_bwd_state: null state buffer");
  return routing_bwd_impl(caps, W, beta_u, beta_a, dout, state, dcaps, dW, dbeta_u, dbeta_a, b, C, s);
}

B2C_API int b2c_class_mean_fwd(const float* rout, float* act, int32_t N, int32_t L, int32_t C, b2c_stream_t s) {
  B2C_REQUIRE(rout && act && N > 0 && L > 0 && C > 0, "class_mean_fwd: bad args");
  B2C_REQUIRE(C <= 32, "class_mean_fwd: C=%d > 32", C);
  class_mean_kernel<<<N, 256, 0, (cudaStream_t)s>>>(rout, act, L, C);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("class_mean_fwd");
  return 0;
}

B2C_API int b2c_pose_mask_fwd(const float* rout, const float* mask, void* x, int32_t N, int32_t L, int32_t C, b2c_stream_t s) {
  B2C_REQUIRE(rout && mask && x && N > 0 && C > 0, "pose_mask_fwd: bad args");
  const long long total = (long long)N * L * (C * 16 / 8);
  int blocks = (int)((total + 255) / 256);
  if (b2c_precision()) pose_mask_kernel<float><<<blocks, 256, 0, (cudaStream_t)s>>>(rout, mask, (float*)x, L, C, total);
  else pose_mask_kernel<bf16><<<blocks, 256, 0, (cudaStream_t)s>>>(rout, mask, (bf16*)x, L, C, total);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("pose_mask_fwd");
  return 0;
}

B2C_API int b2c_caps_head_bwd(const void* dx, const float* mask, const float* dact, const float* dfeat, float* drout, int32_t N,
                              int32_t L, int32_t C, b2c_stream_t s) {
  B2C_REQUIRE(drout && mask && N > 0, "caps_head_bwd: bad args");
  const long long rows = (long long)N * L;
  const long long total = rows * C * 17;
  int blocks = (int)((total + 255) / 256);
  const int cap = b2c_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (b2c_precision())
    caps_head_bwd_kernel<float><<<blocks, 256, 0, (cudaStream_t)s>>>((const float*)dx, mask, dact, dfeat, drout, L, C, rows);
  else
    caps_head_bwd_kernel<bf16><<<blocks, 256, 0, (cudaStream_t)s>>>((const bf16*)dx, mask, dact, dfeat, drout, L, C, rows);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("caps_head_bwd");
  return 0;
}

B2C_API int b2c_primarycaps_bwd_prep(const float* g, const float* out, void* dz, float* dbias, int64_t rows, int32_t dz_pitch,
                                     b2c_stream_t s) {
  B2C_REQUIRE(g && out && dz && dbias && rows > 0 && dz_pitch >= 544 && dz_pitch % 8 == 0, "primarycaps_bwd_prep: bad args");
  long long blocks = (rows + 2) / 3;
  const long long cap = (long long)b2c_num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (b2c_precision())
    primarycaps_bwd_prep_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(g, out, (float*)dz, dbias, rows, dz_pitch, nullptr, 0, 0, 0);
  else
    primarycaps_bwd_prep_kernel<bf16><<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(g, out, (bf16*)dz, dbias, rows, dz_pitch, nullptr, 0, 0, 0);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("primarycaps_bwd_prep");
  return 0;
}

B2C_API int b2c_primarycaps_bwd_prep2(const float* g, const float* out, void* dz, void* dz_rows, float* dbias, int32_t N, int32_t Hq,
                                      int32_t Wq, int32_t dz_pitch, b2c_stream_t s) {
  B2C_REQUIRE(g && out && dz && dz_rows && dbias && N > 0 && Hq > 0 && Wq > 0 && dz_pitch >= 544 && dz_pitch % 8 == 0,
              "primarycaps_bwd_prep2: bad args");
  const long long rows = (long long)N * Hq * Wq;
  long long blocks = (rows + 2) / 3;
  const long long cap = (long long)b2c_num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (b2c_precision())
    primarycaps_bwd_prep_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(g, out, (float*)dz, dbias, rows, dz_pitch,
                                                                                      (float*)dz_rows, N, Hq, Wq);
  else
    primarycaps_bwd_prep_kernel<bf16><<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(g, out, (bf16*)dz, dbias, rows, dz_pitch,
                                                                                     (bf16*)dz_rows, N, Hq, Wq);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("primarycaps_bwd_prep2");
  return 0;
}

namespace {
// (H, W, N, C) -> (N, H, W, C) channel window of a wider tensor (kToRows: the opposite direction): one 16-byte vector per thread
template <bool kToRows>
__global__ void __launch_bounds__(256) rows_clips_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int N, int H, int W, int CV,
                                                         long long clip_rs16, int clip_co16, unsigned total) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned r = i / (unsigned)CV;
    const unsigned cv = i - r * (unsigned)CV;
    const unsigned w = r % (unsigned)W;
    r /= (unsigned)W;
    const unsigned h = r % (unsigned)H, n = r / (unsigned)H;      // clip-major order (n, h, w)
    const size_t rows_idx = (((size_t)h * W + w) * N + n) * CV + cv;                              // compact (H, W, N, C)
    const size_t clip_idx = (((size_t)n * H + h) * W + w) * (size_t)clip_rs16 + clip_co16 + cv;   // view of (N, H, W, Ctot)
    if (kToRows) out[rows_idx] = in[clip_idx];
    else out[clip_idx] = in[rows_idx];
  }
}

int rows_clips_launch(bool to_rows, const void* in, void* out, int32_t N, int32_t H, int32_t W, int32_t C, int64_t clip_rs, int32_t clip_co,
                      b2c_stream_t s, const char* what) {
  const int esz = b2c_precision() ? 4 : 2;
  B2C_REQUIRE(in && out && N > 0 && H > 0 && W > 0 && C > 0 && (C * esz) % 16 == 0 && (clip_rs * esz) % 16 == 0 && (clip_co * esz) % 16 == 0 &&
                  clip_rs >= clip_co + C,
              "%s: bad args", what);
  const int CV = C * esz / 16;
  const long long total = (long long)N * H * W * CV;
  B2C_REQUIRE(total < (1LL << 31), "%s: tensor too large", what);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)b2c_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  const long long rs16 = clip_rs * esz / 16;
  const int co16 = (int)(clip_co * esz / 16);
  if (to_rows)
    rows_clips_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>((const uint4*)in, (uint4*)out, N, H, W, CV, rs16, co16, (unsigned)total);
  else
    rows_clips_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>((const uint4*)in, (uint4*)out, N, H, W, CV, rs16, co16, (unsigned)total);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK(what);
  return 0;
}
}  // namespace

B2C_API int b2c_rows_to_clips(const void* in, void* out, int64_t out_row_stride, int32_t out_c_off, int32_t N, int32_t H, int32_t W,
                              int32_t C, b2c_stream_t s) {
  return rows_clips_launch(false, in, out, N, H, W, C, out_row_stride, out_c_off, s, "rows_to_clips");
}

B2C_API int b2c_clips_to_rows(const void* in, int64_t in_row_stride, int32_t in_c_off, void* out, int32_t N, int32_t H, int32_t W,
                              int32_t C, b2c_stream_t s) {
  return rows_clips_launch(true, in, out, N, H, W, C, in_row_stride, in_c_off, s, "clips_to_rows");
}
