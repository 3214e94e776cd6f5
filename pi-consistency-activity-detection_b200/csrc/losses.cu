// Losses of the step, forward + backward, fused and kept on the device (the reference round-trips
// the consistency masks through numpy on the host, utils/helpers.py:29,87):
//   * BCEWithLogits + Dice on the labeled clips          (main_ucf101.py:89-92, utils/losses.py:44-57)
//   * SpreadLoss                                          (utils/losses.py:14-37)
//   * temporal-variance attentive masks (--bv)            (utils/helpers.py:8-67)
//   * gradient-smoothness mask (--gv)                     (utils/helpers.py:70-95)
//   * flip-consistency weighted MSE incl. the gv (B,B,..) broadcast quirk (main_ucf101.py:100-148)
// All are HBM-bound streaming kernels: coalesced fp32 loads along W, warp-shuffle + one fp64 atomic
// per block for the sums, per-clip min/max through float CAS atomics.
#include "common.cuh"
#include "../../include/b200caps.h"

long long b2c_launches_add(long long n);

namespace {

constexpr int kBlock = 256;
constexpr int kT = 8;   // frames per clip (helpers.py hard-codes 8 x 224 x 224)

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-reduce N float partials and add them to double accumulators
template <int N>
__device__ __forceinline__ void block_accumulate(const float* part, double* acc) {
  __shared__ double sh[N][kBlock / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const double s = warp_sum_d((double)part[q]);
    if (lane == 0) sh[q][w] = s;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    double s = 0;
    for (int i = 0; i < kBlock / 32; ++i) s += sh[threadIdx.x][i];
    atomicAdd(&acc[threadIdx.x], s);
  }
}
__device__ __forceinline__ void atomic_min_f(float* addr, float v) {
  int* a = reinterpret_cast<int*>(addr);
  int old = *a;
  while (__int_as_float(old) > v) {
    const int prev = atomicCAS(a, old, __float_as_int(v));
    if (prev == old) break;
    old = prev;
  }
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
  int* a = reinterpret_cast<int*>(addr);
  int old = *a;
  while (__int_as_float(old) < v) {
    const int prev = atomicCAS(a, old, __float_as_int(v));
    if (prev == old) break;
    old = prev;
  }
}
__device__ __forceinline__ void block_minmax(float mn, float mx, float* dst_min, float* dst_max) {
  __shared__ float smn[kBlock / 32], smx[kBlock / 32];
  mn = warp_min(mn);
  mx = warp_max(mx);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    smn[w] = mn;
    smx[w] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < kBlock / 32; ++i) {
      mn = fminf(mn, smn[i]);
      mx = fmaxf(mx, smx[i]);
    }
    atomic_min_f(dst_min, mn);
    atomic_max_f(dst_max, mx);
  }
  __syncthreads();
}
__device__ __forceinline__ float sigmoid_precise(float x) { return 1.f / (1.f + expf(-x)); }

// ---------------------------------------------------------------------------------------------
// segmentation loss: grid (blocks, n_lab)
__global__ void __launch_bounds__(kBlock) seg_loss_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ targets,
                                                              const int32_t* __restrict__ lab_idx, long long V, double* sums) {
  const long long row = lab_idx[blockIdx.y];
  const float* x = logits + row * V;
  const float* t = targets + row * V;
  float part[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long i = ((long long)blockIdx.x * kBlock + threadIdx.x) * 4; i < V; i += (long long)gridDim.x * kBlock * 4) {
    const float4 xv = *reinterpret_cast<const float4*>(x + i);
    const float4 tv = *reinterpret_cast<const float4*>(t + i);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ts[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float p = sigmoid_precise(xs[q]);
      part[0] += fmaxf(xs[q], 0.f) - xs[q] * ts[q] + log1pf(expf(-fabsf(xs[q])));
      part[1] += p * ts[q];
      part[2] += p;
      part[3] += ts[q];
    }
  }
  block_accumulate<4>(part, sums);
}
__global__ void seg_loss_finish_kernel(const double* sums, int n_lab, long long V, float* loss) {
  const double cnt = (double)n_lab * (double)V;
  loss[0] = (float)(sums[0] / cnt);
  loss[1] = (float)(1.0 - (2.0 * sums[1] + 1.0) / (sums[2] + sums[3] + 1.0));
}
__global__ void __launch_bounds__(kBlock) seg_loss_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ targets,
                                                              const int32_t* __restrict__ lab_idx, int n_lab, long long V,
                                                              const double* __restrict__ sums, float w_bce, float w_dice,
                                                              float* __restrict__ dlogits) {
  const long long row = lab_idx[blockIdx.y];
  const float* x = logits + row * V;
  const float* t = targets + row * V;
  float* g = dlogits + row * V;
  const double den = sums[2] + sums[3] + 1.0;
  const float k_bce = w_bce / (float)((double)n_lab * (double)V);
  const float c1 = (float)(-2.0 / den) * w_dice;                       // * t
  const float c2 = (float)((2.0 * sums[1] + 1.0) / (den * den)) * w_dice;
  for (long long i = ((long long)blockIdx.x * kBlock + threadIdx.x) * 4; i < V; i += (long long)gridDim.x * kBlock * 4) {
    const float4 xv = *reinterpret_cast<const float4*>(x + i);
    const float4 tv = *reinterpret_cast<const float4*>(t + i);
    float4 gv = *reinterpret_cast<float4*>(g + i);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ts[4] = {tv.x, tv.y, tv.z, tv.w};
    float gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float p = sigmoid_precise(xs[q]);
      gs[q] += k_bce * (p - ts[q]) + (c1 * ts[q] + c2) * p * (1.f - p);
    }
    *reinterpret_cast<float4*>(g + i) = make_float4(gs[0], gs[1], gs[2], gs[3]);
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void spread_loss_kernel(const float* __restrict__ act, const float* __restrict__ target, const int32_t* __restrict__ lab_idx,
                                   int n_lab, int C, float m_min, float* loss, float w, float* dact) {
  __shared__ float s_loss[32], s_abs[32];
  const int lane = threadIdx.x;  // one warp
  float l = 0.f, la = 0.f;
  const float b = (float)n_lab;
  for (int i = 0; i < n_lab; ++i) {
    const int row = lab_idx[i];
    const int tg = (int)target[row];
    const float at = act[row * C + tg];
    float sum_pos = 0.f;
    for (int j = lane; j < C; j += 32) {
      const float x = act[row * C + j];
      const float e = fmaxf(m_min - (at - x), 0.f);
      const float ea = fmaxf(0.9f - (at - x), 0.f);
      l += e * e;
      la += ea * ea;
      if (dact && j != tg) {
        const float gj = 2.f * e / (b * b) * w;
        dact[row * C + j] += gj;
        sum_pos += gj;
      }
    }
    sum_pos = warp_sum(sum_pos);
    if (dact && lane == 0) dact[row * C + tg] -= sum_pos;
  }
  l = warp_sum(l);
  la = warp_sum(la);
  if (lane == 0 && loss) {
    loss[0] = (l / b - m_min * m_min) / b;
    loss[1] = la / b - 0.81f;
  }
  (void)s_loss;
  (void)s_abs;
}

// ---------------------------------------------------------------------------------------------
__global__ void init_minmax_kernel(float* mm, int n_pairs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pairs) {
    mm[2 * i] = INFINITY;
    mm[2 * i + 1] = -INFINITY;
  }
}

// cyclic windowed population variance, numpy order: mean = (x0+..)/n ; var = sum((x-mean)^2)/n  (fp32)
template <int NF>
__device__ __forceinline__ void cyc_var(const float* cyc, float* v) {
  constexpr int h = NF / 2;
#pragma unroll
  for (int p = 0; p < 14; ++p) {
    float s = 0.f;
#pragma unroll
    for (int q = -h; q <= h; ++q) s += cyc[(p + q + 14) % 14];
    const float mean = s / (float)NF;
    float a = 0.f;
#pragma unroll
    for (int q = -h; q <= h; ++q) {
      const float dv = cyc[(p + q + 14) % 14] - mean;
      a += dv * dv;
    }
    v[p] = a / (float)NF;
  }
}
__device__ __forceinline__ void fold14(const float* v, float* o) {
  o[0] = 2.f * v[0];
  o[7] = 2.f * v[7];
#pragma unroll
  for (int k = 1; k < 7; ++k) o[k] = v[k] + v[14 - k];
}

// grid (blocks over H*W, P).  cyc = pred[0..7] ++ flip_pred[1..6] (helpers.py:29) with optional index
// transforms so the fused step can feed the raw second-pass output: pred_tflip / fp_tflip reverse time,
// fp_wmirror mirrors W (main_ucf101.py:100,114-115).
template <int NF>
__global__ void __launch_bounds__(kBlock) bv_mask_kernel(const float* __restrict__ pred, const float* __restrict__ fpred,
                                                         float* __restrict__ m, float* __restrict__ mm, int H, int W, int use_sig,
                                                         int pred_tflip, int fp_tflip, int fp_wmirror) {
  const int p = blockIdx.y;
  const long long HW = (long long)H * W;
  const float* o = pred + (long long)p * kT * HW;
  const float* f = fpred + (long long)p * kT * HW;
  float mn0 = INFINITY, mx0 = -INFINITY;
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < HW; i += (long long)gridDim.x * kBlock) {
    const int w = (int)(i % W);
    const long long im = fp_wmirror ? (i - w + (W - 1 - w)) : i;
    float cyc[14], v[14], r[8];
#pragma unroll
    for (int t = 0; t < kT; ++t) {
      float a = o[(pred_tflip ? 7 - t : t) * HW + i];
      if (use_sig) a = sigmoid_precise(a);
      cyc[t] = a;
    }
#pragma unroll
    for (int k = 1; k < 7; ++k) {
      float a = f[(fp_tflip ? 7 - k : k) * HW + im];
      if (use_sig) a = sigmoid_precise(a);
      cyc[7 + k] = a;
    }
    cyc_var<NF>(cyc, v);
    fold14(v, r);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      m[((long long)p * kT + t) * HW + i] = r[t];
      mn0 = fminf(mn0, r[t]);
      mx0 = fmaxf(mx0, r[t]);
    }
  }
  block_minmax(mn0, mx0, mm + p * 2 + 0, mm + p * 2 + 1);
}
// per-clip min-max normalisation exactly as the reference: x -= min ; x /= (max(x) - min(x) + 1e-7)
__global__ void __launch_bounds__(kBlock) minmax_normalize_kernel(float* __restrict__ m, const float* __restrict__ mm, long long per_clip) {
  const int p = blockIdx.y;
  const float mn = mm[p * 2], mx = mm[p * 2 + 1];
  const float den = ((mx - mn) - 0.f) + 1e-7f;
  float* x = m + (long long)p * per_clip;
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < per_clip; i += (long long)gridDim.x * kBlock)
    x[i] = (x[i] - mn) / den;
}

__device__ __forceinline__ void grad_t8(const float* a, float* g) {   // np.gradient along an axis of length 8
  g[0] = a[1] - a[0];
  g[7] = a[7] - a[6];
#pragma unroll
  for (int t = 1; t < 7; ++t) g[t] = (a[t + 1] - a[t - 1]) / 2.f;
}
__global__ void __launch_bounds__(kBlock) gv_mask_kernel(const float* __restrict__ out, float* __restrict__ m, float* __restrict__ mm,
                                                         int H, int W, float lower, float upper, int use_lower, int use_upper) {
  const int p = blockIdx.y;
  const long long HW = (long long)H * W;
  const float* o = out + (long long)p * kT * HW;
  float mn = INFINITY, mx = -INFINITY;
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < HW; i += (long long)gridDim.x * kBlock) {
    float a[8], g1[8], g2[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      float s = sigmoid_precise(o[t * HW + i]);
      if (use_lower && s < lower) s = 0.f;
      if (use_upper && s > upper) s = 1.f;
      a[t] = s;
    }
    grad_t8(a, g1);
    grad_t8(g1, g2);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      m[((long long)p * kT + t) * HW + i] = g2[t];
      mn = fminf(mn, g2[t]);
      mx = fmaxf(mx, g2[t]);
    }
  }
  block_minmax(mn, mx, mm + p * 2, mm + p * 2 + 1);
}

// consistency reduce: thread per (t,h,w), loop over clips.  Weights are already normalised.
//   d = (mirror ? flipW(flp) : flp) - out
//   acc: double[4] = sum d^2, sum w1 d^2, sum w2' d^2 (w2' = time-flipped w2 if w2_tflip), sum_thw (sum_i d_i^2)(sum_j wg_j)
__global__ void __launch_bounds__(kBlock) cons_reduce_kernel(const float* __restrict__ out, const float* __restrict__ flp,
                                                             const float* __restrict__ w1, const float* __restrict__ w2,
                                                             const float* __restrict__ wg, double* acc, int P, int H, int W,
                                                             int mirror, int w2_tflip) {
  const long long HW = (long long)H * W, THW = kT * HW;
  float part[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < THW; i += (long long)gridDim.x * kBlock) {
    const int w = (int)(i % W);
    const int t = (int)(i / HW);
    const long long im = mirror ? (i - w + (W - 1 - w)) : i;
    const long long ia = w2_tflip ? (i + (long long)(7 - 2 * t) * HW) : i;
    float A = 0.f, B = 0.f;
    for (int p = 0; p < P; ++p) {
      const float d = flp[p * THW + im] - out[p * THW + i];
      const float d2 = d * d;
      part[0] += d2;
      if (w1) part[1] += w1[p * THW + i] * d2;
      if (w2) part[2] += w2[p * THW + ia] * d2;
      if (wg) {
        A += d2;
        B += wg[p * THW + i];
      }
    }
    part[3] += A * B;
  }
  block_accumulate<4>(part, acc);
}
__global__ void cons_finish_kernel(const double* acc, float* loss, int P, long long THW, int mode, float wt_ramp, float bv_wt,
                                   float gv_wt, const float* __restrict__ dev) {
  if (dev) {   // device-resident schedule scalars: one captured graph serves every epoch
    wt_ramp = dev[0];
    bv_wt = dev[1];
    gv_wt = dev[2];
  }
  const double cnt = (double)P * (double)THW;
  const double l2 = acc[0] / cnt;
  const double lv = (acc[1] + acc[2]) / cnt;
  const double lg = acc[3] / ((double)THW * (double)P * (double)P);
  const double c1 = wt_ramp * lv + (1.0 - wt_ramp) * l2;
  double cons;
  if ((mode & 3) == 3) cons = bv_wt * c1 + gv_wt * lg;
  else if (mode & 2) cons = lg;
  else if (mode & 1) cons = c1;
  else cons = l2;
  loss[0] = (float)cons;
  loss[1] = (float)l2;
  loss[2] = (float)lv;
  loss[3] = (float)lg;
}
// g_d = (a_l2 + a_lv (w1 + w2')) 2 d / (P THW) + a_lg 2 B d / (THW P^2) ; dout -= g_d ; dflp[mirrored] += g_d
__global__ void __launch_bounds__(kBlock) cons_grad_kernel(const float* __restrict__ out, const float* __restrict__ flp,
                                                           const float* __restrict__ w1, const float* __restrict__ w2,
                                                           const float* __restrict__ wg, float* __restrict__ dout,
                                                           float* __restrict__ dflp, int P, int H, int W, int mirror, int w2_tflip,
                                                           float a_l2, float a_lv, float a_lg, const float* __restrict__ dev) {
  if (dev) {
    a_l2 = dev[0];
    a_lv = dev[1];
    a_lg = dev[2];
  }
  const long long HW = (long long)H * W, THW = kT * HW;
  const float k1 = 2.f / (float)((double)P * (double)THW);
  const float k2 = 2.f / (float)((double)THW * (double)P * (double)P);
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < THW; i += (long long)gridDim.x * kBlock) {
    const int w = (int)(i % W);
    const int t = (int)(i / HW);
    const long long im = mirror ? (i - w + (W - 1 - w)) : i;
    const long long ia = w2_tflip ? (i + (long long)(7 - 2 * t) * HW) : i;
    float B = 0.f;
    if (wg)
      for (int p = 0; p < P; ++p) B += wg[p * THW + i];
    for (int p = 0; p < P; ++p) {
      const float d = flp[p * THW + im] - out[p * THW + i];
      float coef = a_l2 * k1;
      float ws = 0.f;
      if (w1) ws += w1[p * THW + i];
      if (w2) ws += w2[p * THW + ia];
      coef += a_lv * k1 * ws;
      if (wg) coef += a_lg * k2 * B;
      const float g = coef * d;
      if (dout) dout[p * THW + i] -= g;
      if (dflp) dflp[p * THW + im] += g;
    }
  }
}

inline int blocks_for(long long n, int per_thread = 1) {
  long long b = (n + (long long)kBlock * per_thread - 1) / ((long long)kBlock * per_thread);
  const long long cap = (long long)b2c_num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

B2C_API int b2c_seg_loss_fwd(const float* logits, const float* targets, const int32_t* lab_idx, int32_t n_lab, int64_t V,
                             double* sums, float* loss, b2c_stream_t s) {
  B2C_REQUIRE(logits && targets && lab_idx && sums && loss, "seg_loss_fwd: null pointer");
  B2C_REQUIRE(n_lab > 0 && V > 0 && V % 4 == 0, "seg_loss_fwd: n_lab=%d V=%lld (V must be a multiple of 4)", n_lab, (long long)V);
  cudaMemsetAsync(sums, 0, 4 * sizeof(double), (cudaStream_t)s);
  int bx = blocks_for(V, 4);
  if ((long long)bx * n_lab > (long long)b2c_num_sms() * 8) bx = (b2c_num_sms() * 8 + n_lab - 1) / n_lab;
  seg_loss_fwd_kernel<<<dim3(bx, n_lab), kBlock, 0, (cudaStream_t)s>>>(logits, targets, lab_idx, V, sums);
  seg_loss_finish_kernel<<<1, 1, 0, (cudaStream_t)s>>>(sums, n_lab, V, loss);
  b2c_launches_add(2);
  B2C_LAUNCH_CHECK("seg_loss_fwd");
  return 0;
}

B2C_API int b2c_seg_loss_bwd(const float* logits, const float* targets, const int32_t* lab_idx, int32_t n_lab, int64_t V,
                             const double* sums, float w_bce, float w_dice, float* dlogits, b2c_stream_t s) {
  B2C_REQUIRE(logits && targets && lab_idx && sums && dlogits, "seg_loss_bwd: null pointer");
  B2C_REQUIRE(n_lab > 0 && V > 0 && V % 4 == 0, "seg_loss_bwd: bad sizes");
  int bx = blocks_for(V, 4);
  if ((long long)bx * n_lab > (long long)b2c_num_sms() * 8) bx = (b2c_num_sms() * 8 + n_lab - 1) / n_lab;
  seg_loss_bwd_kernel<<<dim3(bx, n_lab), kBlock, 0, (cudaStream_t)s>>>(logits, targets, lab_idx, n_lab, V, sums, w_bce, w_dice,
                                                                       dlogits);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("seg_loss_bwd");
  return 0;
}

B2C_API int b2c_spread_loss(const float* act, const float* target, const int32_t* lab_idx, int32_t n_lab, int32_t C, float m_min,
                            float* loss, float w, float* dact, b2c_stream_t s) {
  B2C_REQUIRE(act && target && lab_idx && n_lab > 0 && C > 0, "spread_loss: bad args");
  spread_loss_kernel<<<1, 32, 0, (cudaStream_t)s>>>(act, target, lab_idx, n_lab, C, m_min, loss, w, dact);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("spread_loss");
  return 0;
}

B2C_API int b2c_bv_mask(const float* pred, const float* flip_pred, float* m, float* mm, int32_t P, int32_t H, int32_t W,
                        int32_t frames_cnt, int32_t use_sigmoid, int32_t pred_tflip, int32_t fp_tflip, int32_t fp_wmirror,
                        b2c_stream_t s) {
  B2C_REQUIRE(pred && flip_pred && m && mm && P > 0, "bv_mask: bad args");
  B2C_REQUIRE(frames_cnt == 3 || frames_cnt == 5, "bv_mask: frames_cnt=%d (the reference implements only 3 and 5)", frames_cnt);
  init_minmax_kernel<<<(P + 127) / 128, 128, 0, (cudaStream_t)s>>>(mm, P);
  int bx = blocks_for((long long)H * W);
  if ((long long)bx * P > (long long)b2c_num_sms() * 8) bx = (b2c_num_sms() * 8 + P - 1) / P;
  if (frames_cnt == 3)
    bv_mask_kernel<3><<<dim3(bx, P), kBlock, 0, (cudaStream_t)s>>>(pred, flip_pred, m, mm, H, W, use_sigmoid, pred_tflip, fp_tflip,
                                                                   fp_wmirror);
  else
    bv_mask_kernel<5><<<dim3(bx, P), kBlock, 0, (cudaStream_t)s>>>(pred, flip_pred, m, mm, H, W, use_sigmoid, pred_tflip, fp_tflip,
                                                                   fp_wmirror);
  minmax_normalize_kernel<<<dim3(bx, P), kBlock, 0, (cudaStream_t)s>>>(m, mm, (long long)kT * H * W);
  b2c_launches_add(3);
  B2C_LAUNCH_CHECK("bv_mask");
  return 0;
}

B2C_API int b2c_gv_mask(const float* out, float* m, float* mm, int32_t P, int32_t H, int32_t W, float lower, float upper,
                        int32_t use_lower, int32_t use_upper, b2c_stream_t s) {
  B2C_REQUIRE(out && m && mm && P > 0, "gv_mask: bad args");
  init_minmax_kernel<<<(P + 127) / 128, 128, 0, (cudaStream_t)s>>>(mm, P);
  int bx = blocks_for((long long)H * W);
  if ((long long)bx * P > (long long)b2c_num_sms() * 8) bx = (b2c_num_sms() * 8 + P - 1) / P;
  gv_mask_kernel<<<dim3(bx, P), kBlock, 0, (cudaStream_t)s>>>(out, m, mm, H, W, lower, upper, use_lower, use_upper);
  minmax_normalize_kernel<<<dim3(bx, P), kBlock, 0, (cudaStream_t)s>>>(m, mm, (long long)kT * H * W);
  b2c_launches_add(3);
  B2C_LAUNCH_CHECK("gv_mask");
  return 0;
}

B2C_API int b2c_cons_reduce(const float* out, const float* flp, const float* w1, const float* w2, const float* wg, double* acc,
                            int32_t P, int32_t H, int32_t W, int32_t mirror, int32_t w2_tflip, b2c_stream_t s) {
  B2C_REQUIRE(out && flp && acc && P > 0, "cons_reduce: bad args");
  cudaMemsetAsync(acc, 0, 4 * sizeof(double), (cudaStream_t)s);
  cons_reduce_kernel<<<blocks_for((long long)kT * H * W), kBlock, 0, (cudaStream_t)s>>>(out, flp, w1, w2, wg, acc, P, H, W, mirror,
                                                                                       w2_tflip);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("cons_reduce");
  return 0;
}

B2C_API int b2c_cons_finish(const double* acc, float* loss, int32_t P, int32_t H, int32_t W, int32_t mode, float wt_ramp, float bv_wt,
                            float gv_wt, const float* dev_scalars, b2c_stream_t s) {
  B2C_REQUIRE(acc && loss, "cons_finish: null pointer");
  cons_finish_kernel<<<1, 1, 0, (cudaStream_t)s>>>(acc, loss, P, (long long)kT * H * W, mode, wt_ramp, bv_wt, gv_wt, dev_scalars);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("cons_finish");
  return 0;
}

B2C_API int b2c_cons_grad(const float* out, const float* flp, const float* w1, const float* w2, const float* wg, float* dout,
                          float* dflp, int32_t P, int32_t H, int32_t W, int32_t mirror, int32_t w2_tflip, float a_l2, float a_lv,
                          float a_lg, const float* dev_scalars, b2c_stream_t s) {
  B2C_REQUIRE(out && flp && (dout || dflp) && P > 0, "cons_grad: bad args");
  cons_grad_kernel<<<blocks_for((long long)kT * H * W), kBlock, 0, (cudaStream_t)s>>>(out, flp, w1, w2, wg, dout, dflp, P, H, W,
                                                                                     mirror, w2_tflip, a_l2, a_lv, a_lg, dev_scalars);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("cons_grad");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam semantics, main_ucf101.py:416: betas (0.9, 0.999), eps 1e-6, no weight decay).
// The step counter lives in device memory so the launch is CUDA-graph replayable.
namespace {
__global__ void adam_tick_kernel(int* step) { step[0] += 1; }
__global__ void __launch_bounds__(kBlock) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                      float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                      const int* __restrict__ step_ptr, float gscale, const float* __restrict__ lr_dev) {
  if (lr_dev) lr = lr_dev[0];   // device-resident learning rate (ReduceLROnPlateau changes it between graph replays)
  const int step = step_ptr[0];
  const float bc1 = 1.f - powf(b1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, (float)step));
  const float step_size = lr / bc1;
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n4; i += (long long)gridDim.x * kBlock) {
    float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float ps[4] = {pv.x, pv.y, pv.z, pv.w}, ms[4] = {mv.x, mv.y, mv.z, mv.w}, vs[4] = {vv.x, vv.y, vv.z, vv.w};
    const float gs[4] = {gv.x * gscale, gv.y * gscale, gv.z * gscale, gv.w * gscale};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      ms[q] = b1 * ms[q] + (1.f - b1) * gs[q];
      vs[q] = b2 * vs[q] + (1.f - b2) * gs[q] * gs[q];
      const float denom = sqrtf(vs[q]) / bc2_sqrt + eps;
      ps[q] -= step_size * (ms[q] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(ps[0], ps[1], ps[2], ps[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(ms[0], ms[1], ms[2], ms[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(vs[0], vs[1], vs[2], vs[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long long i = n4 * 4; i < n; ++i) {
      const float gg = g[i] * gscale;
      m[i] = b1 * m[i] + (1.f - b1) * gg;
      v[i] = b2 * v[i] + (1.f - b2) * gg * gg;
      p[i] -= step_size * (m[i] / (sqrtf(v[i]) / bc2_sqrt + eps));
    }
  }
}
}  // namespace

B2C_API int b2c_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                          int32_t* step_dev, float grad_scale, const float* lr_dev, b2c_stream_t s) {
  B2C_REQUIRE(p && g && m && v && step_dev && n > 0, "adam_step: bad args");
  B2C_REQUIRE(((uintptr_t)p & 15) == 0 && ((uintptr_t)g & 15) == 0 && ((uintptr_t)m & 15) == 0 && ((uintptr_t)v & 15) == 0,
              "adam_step: buffers must be 16B aligned");
  adam_tick_kernel<<<1, 1, 0, (cudaStream_t)s>>>(step_dev);
  adam_kernel<<<blocks_for(n, 4), kBlock, 0, (cudaStream_t)s>>>(p, g, m, v, n, lr, beta1, beta2, eps, step_dev, grad_scale, lr_dev);
  b2c_launches_add(2);
  B2C_LAUNCH_CHECK("adam_step");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Evaluation consumer (evaluate_ucf101.py:127,151-168): per frame, pred = sigmoid(logit) >= 0.5 against the ground-truth
// mask -> intersection, union and ground-truth pixel counts.  One block per frame; the host only thresholds the ratios.
namespace {
__global__ void __launch_bounds__(256) frame_iou_counts_kernel(const float* __restrict__ logits, const float* __restrict__ gt,
                                                               int32_t* __restrict__ out, int HW) {
  const long long f = blockIdx.x;
  const float* l = logits + f * HW;
  const float* g = gt + f * HW;
  int inter = 0, uni = 0, ng = 0;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const bool p = 1.f / (1.f + expf(-l[i])) >= 0.5f;      // the reference thresholds the fp32 sigmoid, not the logit
    const bool t = g[i] > 0.f;
    inter += (p && t) ? 1 : 0;
    uni += (p || t) ? 1 : 0;
    ng += t ? 1 : 0;
  }
  __shared__ int sh[3][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    inter += __shfl_xor_sync(0xffffffffu, inter, o);
    uni += __shfl_xor_sync(0xffffffffu, uni, o);
    ng += __shfl_xor_sync(0xffffffffu, ng, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = inter;
    sh[1][threadIdx.x >> 5] = uni;
    sh[2][threadIdx.x >> 5] = ng;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
    out[f * 3 + threadIdx.x] = t;
  }
}
}  // namespace

B2C_API int b2c_frame_iou_counts(const float* logits, const float* gt, int32_t* out, int64_t frames, int32_t HW, b2c_stream_t s) {
  B2C_REQUIRE(logits && gt && out && frames > 0 && HW > 0, "frame_iou_counts: bad args");
  frame_iou_counts_kernel<<<(unsigned)frames, 256, 0, (cudaStream_t)s>>>(logits, gt, out, HW);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("frame_iou_counts");
  return 0;
}
