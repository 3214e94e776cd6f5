// Collapsed decoder tail (include/b200caps.h, "Collapsed decoder tail"): the small kernels around the three per-clip
// GEMMs -- composite-weight construction, the stride-2 gather that turns the GEMM output columns into logits, its
// transpose for the backward pass, and the chain rule back to upsample4 / smooth parameters.
//
// Index conventions (per dimension; the 3-D objects are products of these):
//   upsample4 (ConvTranspose3d k3 s2 p1 op1):  p = 2 i - 1 + k,  k in 0..2,  p in [0, 2I)   (i = 0, k = 0 falls off: p = -1)
//   smooth    (ConvTranspose3d k3 s1 p1)    :  o = p - 1 + m,    m in 0..2
//   composite                               :  o = 2 i - 2 + e,  e = k + m in 0..4
//   column c in 0..5: c < 5 -> e = c with all (k, m), k + m = c;  c = 5 -> e = 2 without (k = 0, m = 2), used when i = 0
//   column index of a GEMM row: ((ct * 6 + ch) * 6 + cw), 216 columns, stored kTailCols = 224 wide.
#include "common.cuh"
#include "../../include/b200caps.h"
#include <type_traits>

long long b2c_launches_add(long long n);

namespace {

constexpr int kC = 128;          // channels of upsample4's input and output
constexpr int kTailCols = 224;   // 216 composite columns padded to a multiple of 16 (UMMA N) / 32 (tf32 K-block)

// (k, m) pairs of a per-dimension column, encoded k * 3 + m; count in the last slot
__device__ __constant__ int8_t kPairs[6][4] = {
    {0, -1, -1, 1},        // e = 0: (0,0)
    {1, 3, -1, 2},         // e = 1: (0,1) (1,0)
    {2, 4, 6, 3},          // e = 2: (0,2) (1,1) (2,0)
    {5, 7, -1, 2},         // e = 3: (1,2) (2,1)
    {8, -1, -1, 1},        // e = 4: (2,2)
    {4, 6, -1, 2},         // e = 2': (1,1) (2,0)
};

// ------------------------------------------------------------------------------------------------------------------
// Weff[n][ci][col] = sum over the column's (k, m) triples of T[k][m],  T[k][m] = sum_c drop[n][c] W4[ci][c][k] Ws[c][m]
// One CTA per (ci, n): A[c][k] = drop * W4 and B[c][m] = Ws in shared memory, the 27 x 27 matrix T = A^T B, then the
// 216 column sums written into both packed operand images.  CTA (0, n) also builds the clip's bias field.
// ------------------------------------------------------------------------------------------------------------------
template <bool kTF32>
__global__ void __launch_bounds__(256) tail_weff_kernel(const float* __restrict__ w4, const float* __restrict__ b4,
                                                        const float* __restrict__ ws, const float* __restrict__ drop,
                                                        void* __restrict__ pf, long long f_stride, void* __restrict__ pd,
                                                        long long d_stride, int d_nkb, float* __restrict__ biasfield) {
  __shared__ float A[kC * 27];
  __shared__ float B[kC * 27];
  __shared__ float T[27 * 27];
  __shared__ float Bn[27];
  const int ci = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < kC * 27; i += 256) {
    const int c = i / 27;
    A[i] = w4[(size_t)ci * kC * 27 + i] * drop[n * kC + c];
    B[i] = ws[i];
  }
  __syncthreads();
  for (int idx = tid; idx < 729; idx += 256) {
    const int k = idx / 27, m = idx - k * 27;
    float acc = 0.f;
#pragma unroll 8
    for (int c = 0; c < kC; ++c) acc = fmaf(A[c * 27 + k], B[c * 27 + m], acc);
    T[idx] = acc;
  }
  if (ci == 0 && tid < 27) {
    float acc = 0.f;
    for (int c = 0; c < kC; ++c) acc = fmaf(drop[n * kC + c] * b4[c], B[c * 27 + tid], acc);
    Bn[tid] = acc;
  }
  __syncthreads();
  if (tid < 216) {
    const int ct = tid / 36, ch = (tid / 6) % 6, cw = tid % 6;
    float acc = 0.f;
    for (int a = 0; a < kPairs[ct][3]; ++a)
      for (int b = 0; b < kPairs[ch][3]; ++b)
        for (int c = 0; c < kPairs[cw][3]; ++c) {
          const int pt = kPairs[ct][a], ph = kPairs[ch][b], pw = kPairs[cw][c];
          const int k = ((pt / 3) * 3 + ph / 3) * 3 + pw / 3;
          const int m = ((pt % 3) * 3 + ph % 3) * 3 + pw % 3;
          acc += T[k * 27 + m];
        }
    using E = typename std::conditional<kTF32, float, bf16>::type;
    // fprop image: GEMM row = column, K = ci (one 224-row N tile, K = 128)
    pack_store<kTF32>(reinterpret_cast<E*>(pf) + (size_t)n * f_stride, 0, tid, ci, acc, kTailCols, kC / (kTF32 ? 32 : 64), 0);
    // dgrad image: GEMM row = ci, K = column (one 128-row N tile)
    pack_store<kTF32>(reinterpret_cast<E*>(pd) + (size_t)n * d_stride, 0, ci, tid, acc, kC, d_nkb, 0);
  }
  if (ci == 0 && tid < 27) {
    // bias field of border class (bt, bh, bw): taps m whose source position p = o + 1 - m lies inside upsample4's output
    const int bt = tid / 9, bh = (tid / 3) % 3, bw = tid % 3;
    float acc = 0.f;
    for (int m = 0; m < 27; ++m) {
      const int mt = m / 9, mh = (m / 3) % 3, mw = m % 3;
      const bool ok = !((bt == 0 && mt == 2) || (bt == 2 && mt == 0) || (bh == 0 && mh == 2) || (bh == 2 && mh == 0) ||
                        (bw == 0 && mw == 2) || (bw == 2 && mw == 0));
      if (ok) acc += Bn[m];
    }
    biasfield[n * 27 + tid] = acc;
  }
}

// per-dimension candidates of an output index: up to 3 (input index, column) pairs
struct Cand {
  int i[3], c[3], n;
};
__device__ __forceinline__ Cand cands(int o, int I) {
  Cand r;
  r.n = 0;
  const int e0 = o & 1;     // even outputs take e = 0, 2, 4; odd ones e = 1, 3
  for (int e = e0; e < 5; e += 2) {
    const int i = (o + 2 - e) >> 1;
    if (i >= 0 && i < I) {
      r.i[r.n] = i;
      r.c[r.n] = (e == 2 && i == 0) ? 5 : e;
      ++r.n;
    }
  }
  return r;
}
__device__ __forceinline__ int border_class(int o, int O) { return o == 0 ? 0 : (o == O - 1 ? 2 : 1); }

// ------------------------------------------------------------------------------------------------------------------
// logits[n][ot][oh][2q], [2q+1] from the planar GEMM output.  One thread per (n, ot, oh, q): all loads of a warp are
// consecutive floats of one plane (coalesced), every element of Y is read exactly once.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tail_gather_fwd_kernel(const float* __restrict__ y, const float* __restrict__ biasfield,
                                                              const float* __restrict__ bs, float* __restrict__ logits, int N,
                                                              int It, int Ih, int Iw) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Iw) return;
  const int oh = blockIdx.y % (2 * Ih), ot = blockIdx.y / (2 * Ih), n = blockIdx.z;
  const long long rows = (long long)N * It * Ih * Iw;
  const Cand ct = cands(ot, It), ch = cands(oh, Ih);
  float a0 = 0.f, a1 = 0.f;
  for (int a = 0; a < ct.n; ++a)
    for (int b = 0; b < ch.n; ++b) {
      const float* p = y + (long long)((ct.c[a] * 6 + ch.c[b]) * 6) * rows + (((long long)n * It + ct.i[a]) * Ih + ch.i[b]) * Iw + q;
      // even output 2q: (e=0, i=q+1) (e=2 or 2', i=q) (e=4, i=q-1); odd output 2q+1: (e=1, i=q+1) (e=3, i=q)
      a0 += p[(q == 0 ? 5 : 2) * rows];
      a1 += p[3 * rows];
      if (q + 1 < Iw) {
        a0 += p[1];
        a1 += p[rows + 1];
      }
      if (q > 0) a0 += p[4 * rows - 1];
    }
  const int cb = (border_class(ot, 2 * It) * 3 + border_class(oh, 2 * Ih)) * 3;
  const float b0 = bs[0];
  float* o = logits + ((((long long)n * 2 * It + ot) * 2 * Ih + oh) * 2 * Iw) + 2 * q;
  const float v0 = a0 + b0 + biasfield[n * 27 + cb + (q == 0 ? 0 : 1)];
  const float v1 = a1 + b0 + biasfield[n * 27 + cb + (q == Iw - 1 ? 2 : 1)];
  *reinterpret_cast<float2*>(o) = make_float2(v0, v1);
}

// ------------------------------------------------------------------------------------------------------------------
// Backward of the gather: dY[pos][col] = dlogits[n][2 i - 2 + e] where the column applies to this position, else 0.
// One thread per (position, 8-column group); the row is 224 wide in the activation precision.
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void store8_row(T* p, const float* v);
template <>
__device__ __forceinline__ void store8_row<bf16>(bf16* p, const float* v) { *reinterpret_cast<uint4*>(p) = pack8(v); }
template <>
__device__ __forceinline__ void store8_row<float>(float* p, const float* v) {
  reinterpret_cast<float4*>(p)[0] = make_float4(tf32_rna(v[0]), tf32_rna(v[1]), tf32_rna(v[2]), tf32_rna(v[3]));
  reinterpret_cast<float4*>(p)[1] = make_float4(tf32_rna(v[4]), tf32_rna(v[5]), tf32_rna(v[6]), tf32_rna(v[7]));
}

// One thread per (position, group of 4 (ct, ch) column pairs = 24 columns = three 16-byte bf16 stores): the 6 columns of a
// pair read 5 consecutive dlogits of one output row, so a thread issues 20 loads for 24 outputs and no per-column index
// arithmetic (the per-(position, 8 columns) version spent 1.46 ms on integer divisions; r02 bench).
template <typename T>
__global__ void __launch_bounds__(256) tail_gather_bwd_kernel(const float* __restrict__ g, T* __restrict__ dy, int N, int It, int Ih,
                                                              int Iw, long long total) {
  constexpr int kGroups = 9;   // 36 (ct, ch) pairs / 4
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    // (host guarantees total < 2^31: 32-bit divisions instead of emulated 64-bit ones)
    const unsigned u = (unsigned)idx;
    const int grp = (int)(u % (unsigned)kGroups);
    const unsigned upos = u / (unsigned)kGroups;
    const long long pos = upos;
    const int iw = (int)(upos % (unsigned)Iw);
    unsigned r = upos / (unsigned)Iw;
    const int ih = (int)(r % (unsigned)Ih);
    r /= (unsigned)Ih;
    const int it = (int)(r % (unsigned)It);
    const int n = (int)(r / (unsigned)It);
    const int Ot = 2 * It, Oh = 2 * Ih, Ow = 2 * Iw;
    float v[24];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pair = grp * 4 + j;
      const int ct = pair / 6, ch = pair - ct * 6;
      const int et = ct < 5 ? ct : 2, eh = ch < 5 ? ch : 2;
      const int ot = 2 * it - 2 + et, oh = 2 * ih - 2 + eh;
      // e = 2 belongs to i > 0, its primed twin (column 5) to i = 0
      const bool ok = ot >= 0 && ot < Ot && oh >= 0 && oh < Oh && (ct != 5 || it == 0) && (ct != 2 || it != 0) &&
                      (ch != 5 || ih == 0) && (ch != 2 || ih != 0);
      float w5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      if (ok) {
        const float* row = g + (((long long)n * Ot + ot) * Oh + oh) * Ow + (2 * iw - 2);
#pragma unroll
        for (int e = 0; e < 5; ++e) {
          const int ow = 2 * iw - 2 + e;
          if (ow >= 0 && ow < Ow) w5[e] = __ldg(row + e);
        }
      }
      v[j * 6 + 0] = w5[0];
      v[j * 6 + 1] = w5[1];
      v[j * 6 + 2] = iw != 0 ? w5[2] : 0.f;
      v[j * 6 + 3] = w5[3];
      v[j * 6 + 4] = w5[4];
      v[j * 6 + 5] = iw == 0 ? w5[2] : 0.f;
    }
    T* dst = dy + pos * kTailCols + grp * 24;
    store8_row(dst, v);
    store8_row(dst + 8, v + 8);
    store8_row(dst + 16, v + 16);
    if (grp == kGroups - 1) {
      const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      store8_row(dst + 24, z);     // columns 216..223: padding
    }
  }
}

// per-clip sums of dlogits over the 27 border classes (gradient of the bias field); one block per (n, ot) plane: every thread
// keeps the 9 (oh class, ow class) sums of the elements it walks, the block reduces them and issues 9 atomics (one block
// per ROW was 57 344 blocks of 128 threads for 224 floats each: 71 us in the captured step).
__global__ void __launch_bounds__(256) tail_class_sums_kernel(const float* __restrict__ g, float* __restrict__ sums, int Ot, int Oh,
                                                              int Ow) {
  const int ot = blockIdx.x % Ot, n = blockIdx.x / Ot;
  const float* plane = g + (long long)blockIdx.x * Oh * Ow;
  float acc[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) acc[q] = 0.f;
  // whole rows per warp: lane walks the row with stride 32, so the ow class is known per element and the oh class per row
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int oh = w; oh < Oh; oh += 8) {
    const int ch = border_class(oh, Oh);
    const float* row = plane + (long long)oh * Ow;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int x = lane; x < Ow; x += 32) {
      const float v = row[x];
      if (x == 0) a0 += v;
      else if (x == Ow - 1) a2 += v;
      else a1 += v;
    }
#pragma unroll
    for (int q = 0; q < 3; ++q)
      if (ch == q) {
        acc[q * 3 + 0] += a0;
        acc[q * 3 + 1] += a1;
        acc[q * 3 + 2] += a2;
      }
  }
  __shared__ float sh[8][9];
#pragma unroll
  for (int q = 0; q < 9; ++q) {
    const float t = warp_sum(acc[q]);
    if (lane == 0) sh[w][q] = t;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    float t = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) t += sh[ww][threadIdx.x];
    const int cb = n * 27 + border_class(ot, Ot) * 9;
    atomicAdd(sums + cb + threadIdx.x, t);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Chain rule.  dT[n][ci][k][m] = sum of dWeff[n][ci][col] over the columns that contain (k, m);
//   dW4[ci][c][k] += sum_n drop[n][c] sum_m Ws[c][m] dT[k][m]         dWs[c][m] += sum_n drop[n][c] sum_k W4[ci][c][k] dT[k][m]
// One CTA per ci walks all clips: threads 0..127 own dW4[ci][c][:] (no atomics needed across clips), threads 128..255
// own the CTA's partial dWs[c][:] (one atomicAdd per element per CTA).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tail_chain_kernel(const float* __restrict__ dweff, const float* __restrict__ w4,
                                                         const float* __restrict__ ws, const float* __restrict__ drop,
                                                         float* __restrict__ dw4, float* __restrict__ dws, int N) {
  // grid (ci, clip groups).  dT rows are padded to 28 floats and read as float4 (all threads read the same row: the
  // broadcast loads were this kernel's bound -- 729 per thread and clip, 164 us in the captured step).
  __shared__ float dW[216];
  __shared__ __align__(16) float dT[27 * 28];
  const int ci = blockIdx.x, tid = threadIdx.x;
  const int c = tid & 127;
  const bool second = tid >= 128;
  float coef[28], acc[27];
#pragma unroll
  for (int j = 0; j < 27; ++j) {
    coef[j] = second ? w4[((size_t)ci * kC + c) * 27 + j] : ws[c * 27 + j];
    acc[j] = 0.f;
  }
  coef[27] = 0.f;
  const int per = (N + gridDim.y - 1) / gridDim.y;
  const int n_lo = blockIdx.y * per, n_hi = min(N, n_lo + per);
  for (int n = n_lo; n < n_hi; ++n) {
    __syncthreads();
    if (tid < 216) dW[tid] = dweff[((size_t)n * kC + ci) * kTailCols + tid];
    __syncthreads();
    for (int idx = tid; idx < 27 * 28; idx += 256) {
      const int k = idx / 28, m = idx - k * 28;
      float s = 0.f;
      if (m < 27) {
        const int kd[3] = {k / 9, (k / 3) % 3, k % 3}, md[3] = {m / 9, (m / 3) % 3, m % 3};
        int cols[3][2], nc[3];
#pragma unroll
        for (int dd = 0; dd < 3; ++dd) {
          cols[dd][0] = kd[dd] + md[dd];
          nc[dd] = 1;
          if (kd[dd] + md[dd] == 2 && kd[dd] >= 1) cols[dd][nc[dd]++] = 5;
        }
        for (int a = 0; a < nc[0]; ++a)
          for (int b = 0; b < nc[1]; ++b)
            for (int e = 0; e < nc[2]; ++e) s += dW[(cols[0][a] * 6 + cols[1][b]) * 6 + cols[2][e]];
      }
      dT[idx] = s;
    }
    __syncthreads();
    const float dr = drop[n * kC + c];
    if (dr != 0.f) {
      float tmp[28];
#pragma unroll
      for (int m = 0; m < 28; ++m) tmp[m] = 0.f;
#pragma unroll
      for (int k = 0; k < 27; ++k) {
        float row[28];
#pragma unroll
        for (int q = 0; q < 7; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(dT + k * 28 + q * 4);
          row[q * 4 + 0] = v.x; row[q * 4 + 1] = v.y; row[q * 4 + 2] = v.z; row[q * 4 + 3] = v.w;
        }
        if (!second) {
          // dW4[ci][c][k] += dr * sum_m Ws[c][m] dT[k][m]
          float sk = 0.f;
#pragma unroll
          for (int m = 0; m < 27; ++m) sk = fmaf(coef[m], row[m], sk);
          acc[k] = fmaf(dr, sk, acc[k]);
        } else {
          // dWs[c][m] += dr * sum_k W4[ci][c][k] dT[k][m]
#pragma unroll
          for (int m = 0; m < 27; ++m) tmp[m] = fmaf(coef[k], row[m], tmp[m]);
        }
      }
      if (second) {
#pragma unroll
        for (int m = 0; m < 27; ++m) acc[m] = fmaf(dr, tmp[m], acc[m]);
      }
    }
  }
  if (!second) {
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      if (gridDim.y == 1) dw4[((size_t)ci * kC + c) * 27 + k] += acc[k];
      else atomicAdd(dw4 + ((size_t)ci * kC + c) * 27 + k, acc[k]);
    }
  } else {
#pragma unroll
    for (int m = 0; m < 27; ++m) atomicAdd(dws + c * 27 + m, acc[m]);
  }
}

// bias part: dBn[n][m] = sum over border classes that see tap m of class_sums[n][cls];
//   db4[c] += sum_n drop[n][c] sum_m Ws[c][m] dBn[n][m]    dWs[c][m] += sum_n drop[n][c] b4[c] dBn[n][m]    dbs += sum class_sums
__global__ void __launch_bounds__(128) tail_bias_chain_kernel(const float* __restrict__ sums, const float* __restrict__ b4,
                                                              const float* __restrict__ ws, const float* __restrict__ drop,
                                                              float* __restrict__ db4, float* __restrict__ dws, float* __restrict__ dbs,
                                                              int N) {
  extern __shared__ float dBn[];     // [N][27]
  const int tid = threadIdx.x;
  for (int i = tid; i < N * 27; i += blockDim.x) {
    const int n = i / 27, m = i - n * 27;
    const int mt = m / 9, mh = (m / 3) % 3, mw = m % 3;
    float acc = 0.f;
    for (int cl = 0; cl < 27; ++cl) {
      const int bt = cl / 9, bh = (cl / 3) % 3, bw = cl % 3;
      const bool ok = !((bt == 0 && mt == 2) || (bt == 2 && mt == 0) || (bh == 0 && mh == 2) || (bh == 2 && mh == 0) ||
                        (bw == 0 && mw == 2) || (bw == 2 && mw == 0));
      if (ok) acc += sums[n * 27 + cl];
    }
    dBn[i] = acc;
  }
  __syncthreads();
  const int c = tid;     // 128 threads = 128 channels
  float g4 = 0.f, gs[27];
#pragma unroll
  for (int m = 0; m < 27; ++m) gs[m] = 0.f;
  for (int n = 0; n < N; ++n) {
    const float dr = drop[n * kC + c];
    if (dr == 0.f) continue;
#pragma unroll
    for (int m = 0; m < 27; ++m) {
      g4 = fmaf(dr * ws[c * 27 + m], dBn[n * 27 + m], g4);
      gs[m] = fmaf(dr * b4[c], dBn[n * 27 + m], gs[m]);
    }
  }
  db4[c] += g4;
#pragma unroll
  for (int m = 0; m < 27; ++m) dws[c * 27 + m] += gs[m];     // runs after tail_chain_kernel on the same stream
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < N * 27; ++i) t += sums[i];
    dbs[0] += t;
  }
}

}  // namespace

B2C_API int b2c_tail_weff(const float* w4, const float* b4, const float* ws, const float* drop_nc, void* packed_fprop,
                          int64_t fprop_stride, void* packed_dgrad, int64_t dgrad_stride, int32_t dgrad_nkb, float* biasfield, int32_t N,
                          b2c_stream_t s) {
  B2C_REQUIRE(w4 && b4 && ws && drop_nc && packed_fprop && packed_dgrad && biasfield && N > 0 && dgrad_nkb > 0, "tail_weff: bad args");
  dim3 grid(kC, (unsigned)N);
  if (b2c_precision())
    tail_weff_kernel<true><<<grid, 256, 0, (cudaStream_t)s>>>(w4, b4, ws, drop_nc, packed_fprop, fprop_stride, packed_dgrad, dgrad_stride,
                                                             dgrad_nkb, biasfield);
  else
    tail_weff_kernel<false><<<grid, 256, 0, (cudaStream_t)s>>>(w4, b4, ws, drop_nc, packed_fprop, fprop_stride, packed_dgrad, dgrad_stride,
                                                              dgrad_nkb, biasfield);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("tail_weff");
  return 0;
}

B2C_API int b2c_tail_gather_fwd(const float* y_planar, const float* biasfield, const float* bs, float* logits, int32_t N, int32_t It,
                                int32_t Ih, int32_t Iw, b2c_stream_t s) {
  B2C_REQUIRE(y_planar && biasfield && bs && logits && N > 0 && It > 0 && Ih > 0 && Iw > 0, "tail_gather_fwd: bad args");
  B2C_REQUIRE((long long)N * It * Ih * Iw * kTailCols < (1LL << 40) && 4LL * It * Ih <= 65535 && N <= 65535, "tail_gather_fwd: too large");
  dim3 grid((unsigned)((Iw + 127) / 128), (unsigned)(4 * It * Ih), (unsigned)N);
  tail_gather_fwd_kernel<<<grid, 128, 0, (cudaStream_t)s>>>(y_planar, biasfield, bs, logits, N, It, Ih, Iw);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("tail_gather_fwd");
  return 0;
}

B2C_API int b2c_tail_gather_bwd(const float* dlogits, void* dy, float* class_sums, int32_t N, int32_t It, int32_t Ih, int32_t Iw,
                                b2c_stream_t s) {
  B2C_REQUIRE(dlogits && dy && class_sums && N > 0 && It > 0 && Ih > 0 && Iw > 0, "tail_gather_bwd: bad args");
  const long long total = (long long)N * It * Ih * Iw * 9;
  B2C_REQUIRE(total < (1LL << 31), "tail_gather_bwd: too many elements for the 32-bit index split");
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)b2c_num_sms() * 32;
  if (blocks > cap) blocks = cap;
  if (b2c_precision())
    tail_gather_bwd_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(dlogits, (float*)dy, N, It, Ih, Iw, total);
  else
    tail_gather_bwd_kernel<bf16><<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(dlogits, (bf16*)dy, N, It, Ih, Iw, total);
  B2C_LAUNCH_CHECK("tail_gather_bwd");
  tail_class_sums_kernel<<<(unsigned)(N * 2 * It), 256, 0, (cudaStream_t)s>>>(dlogits, class_sums, 2 * It, 2 * Ih, 2 * Iw);
  b2c_launches_add(2);
  B2C_LAUNCH_CHECK("tail_class_sums");
  return 0;
}

B2C_API int b2c_tail_chain_bwd(const float* dweff, const float* class_sums, const float* w4, const float* b4, const float* ws,
                               const float* drop_nc, float* dw4, float* db4, float* dws, float* dbs, int32_t N, b2c_stream_t s) {
  B2C_REQUIRE(dweff && class_sums && w4 && b4 && ws && drop_nc && dw4 && db4 && dws && dbs && N > 0 && N * 27 * 4 <= 40000,
              "tail_chain_bwd: bad args");
  const int groups = N >= 16 ? 4 : (N >= 4 ? 2 : 1);      // clip groups: 128 x 4 CTAs fill the 148 SMs three deep
  tail_chain_kernel<<<dim3(kC, groups), 256, 0, (cudaStream_t)s>>>(dweff, w4, ws, drop_nc, dw4, dws, N);
  B2C_LAUNCH_CHECK("tail_chain");
  tail_bias_chain_kernel<<<1, 128, (size_t)N * 27 * sizeof(float), (cudaStream_t)s>>>(class_sums, b4, ws, drop_nc, db4, dws, dbs, N);
  b2c_launches_add(2);
  B2C_LAUNCH_CHECK("tail_bias_chain");
  return 0;
}
