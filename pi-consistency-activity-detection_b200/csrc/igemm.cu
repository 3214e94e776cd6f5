// Generalised implicit-GEMM convolution for sm_100a: tcgen05.mma (bf16 x bf16 -> fp32 in TMEM), operands staged
// by TMA (im2col tensor maps for activations, bulk copies for pre-swizzled weight tiles) into 128B-swizzled UMMA
// tiles, mbarrier producer/consumer pipelines, TMA tensor stores for the results.
//
//   fprop-like kernel : D[M = output positions (128 or 2x128 per unit), N = Cout tile] = A_im2col[M,K] * W[N,K]^T
//                       A and W are K-major.  Covers conv fprop, dgrad, transposed-conv fprop
//                       (by output-parity class) and transposed-conv dgrad.  Persistent, 416 threads:
//                       warps 0-3 producers (TMA: one thread; cp.async gather: 128 threads), warp 4 TMEM allocator +
//                       single-thread UMMA issuer, warps 5-12 epilogue (two quartets alternating tiles).
//   wgrad kernel      : D[M = 1-3 (tap,ci) tiles of 128, N = Cp tile] = sum_positions Gcol[pos,M] * P[pos,N]
//                       both operands MN-major (positions are the GEMM-K dimension), split over
//                       position ranges, fp32 red.add into the weight gradient.  160 threads: warps 0-3 producers
//                       then epilogue (warp w owns TMEM lanes 32w..32w+31), warp 4 allocator + UMMA issuer.
#include "common.cuh"
#include "../../include/b200caps.h"
#include <string.h>
#include <stdlib.h>
#include <cuda.h>   // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint (no -lcuda)

static long long g_launches = 0;
long long b2c_launches_add(long long n) {
  g_launches += n;
  return g_launches;
}

namespace {

// Optional in-kernel phase timing (compile with -DB2C_PROF): CTA 0 prints, per warp role, total cycles and the
// cycles spent waiting on each mbarrier.
#ifdef B2C_PROF
#define B2C_PROF_DECL(x) long long x = 0
#define B2C_PROF_START(x) x = clock64()
#define B2C_PROF_ACC(a, t0) a += clock64() - t0
#define B2C_PROF_PRINT2(name, t0, a, b) \
  if (blockIdx.x == 0) printf("[prof] %-14s total %lld cyc  waitA %lld  B %lld  tiles/CTA %lld\n", name, clock64() - t0, (long long)(a), (long long)(b), (total_tiles + gridDim.x - 1) / gridDim.x)
#define B2C_PROF_PRINTW(name, t0, a, n) \
  if (blockIdx.x == 0) printf("[prof] %-14s total %lld cyc  wait %lld  k-blocks %d\n", name, clock64() - t0, (long long)(a), (int)(n))
#else
#define B2C_PROF_PRINTW(name, t0, a, n)
#define B2C_PROF_DECL(x)
#define B2C_PROF_START(x)
#define B2C_PROF_ACC(a, t0)
#define B2C_PROF_PRINT2(name, t0, a, b)
#endif

constexpr int kTileM = 128;      // UMMA M
constexpr int kBlockK = 64;      // bf16 elements per 128-byte swizzle row
constexpr int kATileBytes = kTileM * 128;
constexpr int kMaxStages = 8;
constexpr int kThreads = 160;

struct PipeSmem {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t accum;
  uint32_t tmem_base;
  uint32_t pad_;
};

__device__ __forceinline__ int tap_dt(int32_t t) { return (int)(int8_t)(t & 0xff); }
__device__ __forceinline__ int tap_dh(int32_t t) { return (int)(int8_t)((t >> 8) & 0xff); }
__device__ __forceinline__ int tap_dw(int32_t t) { return (int)(int8_t)((t >> 16) & 0xff); }

__device__ __forceinline__ uint32_t tmem_cols_for(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

// =====================================================================================
// fprop-like kernel: persistent, warp specialised
//   warps 0-3 : im2col gather producers (one GEMM row per thread, 16-byte cp.async with zero fill);
//               thread 0 also arms the bulk copy of the pre-swizzled weight tile of each K-block
//   warp  4   : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 5-12: epilogue (TMEM -> registers -> bias / scale / ReLU / sigmoid -> global), two warps per TMEM lane
//               quarter splitting the columns, overlapped with the next tile's main loop through two TMEM accumulators
// Each CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...; the smem ring and its mbarrier phases run
// continuously across tiles, so there is no pipeline drain / fill between tiles.
// =====================================================================================
constexpr int kMaxScaleSmem = 4096;   // floats of dropout scale kept in shared memory
constexpr int kFpropThreads = 416;   // 4 producer warps + 1 MMA warp + 8 epilogue warps

struct alignas(64) TmaMaps {
  CUtensorMap a[8];   // per class: im2col map of the gathered activation view
};
struct alignas(64) OutMaps {
  CUtensorMap o[8];   // per class: tiled map of the bf16 output view (TMA-store epilogue)
};
constexpr int kStageChunkCols = 64;                 // epilogue staging chunk: 64 bf16 columns = one 128-byte swizzle row
constexpr int kStagingBytes = 2 * kTileM * 128;     // one 128-row x 128-byte buffer per epilogue half

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c, int row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c), "r"(row)
               : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c, int w, int h, int t, int n) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c), "r"(w), "r"(h), "r"(t), "r"(n)
               : "memory");
}

// im2col TMA: 128 pixels x 64 channels of one filter tap, 128B-swizzled, zero fill outside the tensor
__device__ __forceinline__ void tma_im2col_5d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c, int w, int h, int t,
                                              int n, uint16_t ow, uint16_t oh, uint16_t ot) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], {%8, %9, %10};" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(t), "r"(n), "h"(ow), "h"(oh), "h"(ot)
      : "memory");
}

// Division by a runtime-constant divisor (Granlund-Montgomery): q = n / d for any 32-bit n.  The per-tile index
// arithmetic (tile -> class -> (clip, t, h, w)) sat on every warp's critical path as hardware-emulated divisions.
struct FastDiv {
  uint32_t m, sh1, sh2, d;
};
__device__ __forceinline__ FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  uint32_t l = 0;
  while ((1ull << l) < (unsigned long long)d) ++l;
  f.m = (uint32_t)((((1ull << l) - d) << 32) / d + 1ull);
  f.sh1 = l < 1 ? l : 1;
  f.sh2 = l > 0 ? l - 1 : 0;
  f.d = d;
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  const uint32_t t = __umulhi(f.m, n);
  return (t + ((n - t) >> f.sh1)) >> f.sh2;
}
// q -> (q / d, q % d)
__device__ __forceinline__ uint32_t fdivmod(uint32_t n, const FastDiv& f, uint32_t& rem) {
  const uint32_t q = fdiv(n, f);
  rem = n - q * f.d;
  return q;
}
struct FpropSmem {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint32_t tmem_base;
  uint32_t pad_;
  FastDiv fd[25];   // [3*c + 0/1/2] = Qw / Qh / Qt of class c; [24] = n_tiles
};

struct TileInfo {
  int cls, n_idx;
  long long m0;
};
struct TileSched {
  unsigned start[9];   // first tile index of each class (prefix sums), start[nclass] = total
  unsigned interleave; // != 0: all classes have the same tile count and are walked class-minor (t % nclass), so that
                       // the CTAs running at any time read the SAME input region for all output-parity classes
                       // (ncu r01f: class-major order re-read the 411 MB upsample4 input once per class, 3.6 GB)
};

__device__ __forceinline__ TileInfo decode_tile(const TileSched& ts, int nclass, long long t, int n_tiles, const FastDiv& fnt,
                                                int MT) {
  TileInfo ti;
  int c = 0;
  unsigned local;
  if (ts.interleave) {
    // nclass is 2, 4 or 8 here (parity classes of a strided transposed conv)
    c = (int)((unsigned)t & (unsigned)(nclass - 1));
    local = (unsigned)t >> (31 - __clz(nclass));
  } else {
#pragma unroll
    for (int i = 1; i < 8; ++i)
      if (i < nclass && (unsigned)t >= ts.start[i]) c = i;
    local = (unsigned)t - ts.start[c];
  }
  ti.cls = c;
  if (n_tiles == 1) {
    ti.n_idx = 0;
    ti.m0 = (long long)local * (kTileM * MT);
  } else {
    uint32_t rem;
    const uint32_t q = fdivmod(local, fnt, rem);
    ti.n_idx = (int)rem;
    ti.m0 = (long long)q * (kTileM * MT);
  }
  return ti;
}

// Taps a unit of rows [m_first, m_last] can skip: when the class marks its taps as blocks of `h_block` taps sharing one
// H offset (2-D layers: PrimaryCaps' 9x9 dgrad reads a 20x20 gradient from 28x28 positions, upsample1's transposed 9x9
// likewise) and the unit lies inside one (clip, t) plane, tap blocks whose source rows fall outside [0, Hi) for EVERY row
// of the unit contribute nothing (TMA would zero-fill them) -- the valid blocks are a contiguous range.  Producer and MMA
// issuer call this with the same arguments.
__device__ __forceinline__ void tap_range(const b2c_conv_desc& d, const b2c_conv_class& cc, const int32_t* taps, unsigned m_first,
                                          unsigned m_last, const FastDiv* fd3, int& lo, int& hi, int& jlo, int& jhi) {
  lo = 0;
  hi = cc.ntaps;
  jlo = 0;
  jhi = 0x7fffffff;                        // [jlo, jhi): taps kept INSIDE every block (h_block < 0 mode only)
  if (cc.h_block == 0) return;
  if (cc.h_block < 0) {
    // Blocks of |h_block| taps share one T offset, the taps inside a block walk the H offsets: a 2-D layer presented with
    // its image rows on the T axis, its image columns on the H axis and the clips on the W axis, so that a 128-position
    // tile holds a few columns of ONE image row of all clips (plans.ConvPlan rows-major).  Tap rows that are padding for
    // that image row are skipped as whole blocks, tap columns that are padding for all of the tile's columns inside them.
    const int hb = -cc.h_block;
    uint32_t r, h0, h1, t0, t1;
    uint32_t q0 = fdivmod(m_first, fd3[0], r);
    uint32_t q1 = fdivmod(m_last, fd3[0], r);
    q0 = fdivmod(q0, fd3[1], h0);
    q1 = fdivmod(q1, fd3[1], h1);
    q0 = fdivmod(q0, fd3[2], t0);
    q1 = fdivmod(q1, fd3[2], t1);
    if (q0 != q1) return;                  // the unit straddles samples
    const int nb = cc.ntaps / hb;
    int first = -1, last = -1;
    for (int b = 0; b < nb; ++b) {
      const int dt = tap_dt(taps[b * hb]);
      const bool ok = (int)t1 * d.si_t + dt >= 0 && (int)t0 * d.si_t + dt <= d.Ti - 1;
      if (ok) {
        if (first < 0) first = b;
        last = b;
      }
    }
    if (first < 0) first = last = 0;
    lo = first * hb;
    hi = (last + 1) * hb;
    if (t0 == t1) {                        // one image row: its columns h0..h1 bound the useful H offsets
      int f2 = -1, l2 = -1;
      for (int j = 0; j < hb; ++j) {
        const int dh = tap_dh(taps[j]);    // (every block walks the same H offsets: host-checked product order)
        const bool ok = (int)h1 * d.si_h + dh >= 0 && (int)h0 * d.si_h + dh <= d.Hi - 1;
        if (ok) {
          if (f2 < 0) f2 = j;
          l2 = j;
        }
      }
      if (f2 < 0) f2 = l2 = 0;
      jlo = f2;
      jhi = l2 + 1;
    }
    return;
  }
  uint32_t r0, r1, h0, h1;
  uint32_t q0 = fdivmod(m_first, fd3[0], r0);
  uint32_t q1 = fdivmod(m_last, fd3[0], r1);
  q0 = fdivmod(q0, fd3[1], h0);
  q1 = fdivmod(q1, fd3[1], h1);
  if (q0 != q1) return;                    // the unit straddles (clip, t) planes
  const int nb = cc.ntaps / cc.h_block;
  int first = -1, last = -1;
  for (int b = 0; b < nb; ++b) {
    const int dh = tap_dh(taps[b * cc.h_block]);
    // some row h in [h0, h1] with 0 <= h * si_h + dh < Hi
    const bool ok = (int)h1 * d.si_h + dh >= 0 && (int)h0 * d.si_h + dh <= d.Hi - 1;
    if (ok) {
      if (first < 0) first = b;
      last = b;
    }
  }
  if (first < 0) first = last = 0;         // nothing valid: keep one (all-zero) block so the accumulator is defined
  lo = first * cc.h_block;
  hi = (last + 1) * cc.h_block;
}

__device__ __forceinline__ float fmax_nan(float a, float b) {   // NaN-propagating max: relu(NaN) = NaN, max(x, -inf) = x
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
// Branch-free epilogue body for bf16 staged outputs: kCols (32 / 16) accumulator columns of one TMEM lane ->
// + bias [* dropout scale] -> max(., lo) -> bf16 -> 128B-swizzled staging row.  The bias / scale vectors are fetched
// from shared memory while the TMEM load is in flight.
template <bool kScale, int kCols>
__device__ __forceinline__ void epi_fast(uint32_t taddr, const float* __restrict__ bias, const float* __restrict__ scale, float lo,
                                         uint32_t row_addr, uint32_t swz, uint32_t u0) {
  uint32_t r[32];
  if (kCols == 32) tmem_ld32_issue(taddr, r);
  else tmem_ld16_issue(taddr, r);
  float4 bb[kCols / 4], ss[kCols / 4];
#pragma unroll
  for (int i = 0; i < kCols / 4; ++i) {
    bb[i] = *reinterpret_cast<const float4*>(bias + 4 * i);
    if (kScale) ss[i] = *reinterpret_cast<const float4*>(scale + 4 * i);
  }
  if (kCols == 32) tmem_ld_fence(r);
  else tmem_ld_fence16(r);
#pragma unroll
  for (int h = 0; h < kCols / 8; ++h) {
    float vv[8];
    const float4 b0 = bb[2 * h], b1 = bb[2 * h + 1];
    vv[0] = __uint_as_float(r[h * 8 + 0]) + b0.x; vv[1] = __uint_as_float(r[h * 8 + 1]) + b0.y;
    vv[2] = __uint_as_float(r[h * 8 + 2]) + b0.z; vv[3] = __uint_as_float(r[h * 8 + 3]) + b0.w;
    vv[4] = __uint_as_float(r[h * 8 + 4]) + b1.x; vv[5] = __uint_as_float(r[h * 8 + 5]) + b1.y;
    vv[6] = __uint_as_float(r[h * 8 + 6]) + b1.z; vv[7] = __uint_as_float(r[h * 8 + 7]) + b1.w;
    if (kScale) {
      const float4 s0 = ss[2 * h], s1 = ss[2 * h + 1];
      vv[0] *= s0.x; vv[1] *= s0.y; vv[2] *= s0.z; vv[3] *= s0.w;
      vv[4] *= s1.x; vv[5] *= s1.y; vv[6] *= s1.z; vv[7] *= s1.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) vv[i] = fmax_nan(vv[i], lo);
    st_shared16(row_addr + (((u0 + (uint32_t)h) ^ swz) << 4), pack8(vv));
  }
}

// kTF32: fp32 activations / weights read by the tensor core as tf32 (tcgen05.mma kind::tf32): a K-block is 32
// channels (still 128 bytes per row), everything else -- tile bytes, swizzle, descriptor stepping (32 bytes per MMA
// K step: 16 bf16 or 8 tf32) -- is identical, so the two precisions share this kernel.
template <bool kTF32>
__global__ void __launch_bounds__(kFpropThreads, 1) igemm_fprop_kernel(const __grid_constant__ b2c_conv_desc d,
                                                                       const __grid_constant__ TmaMaps maps,
                                                                       const __grid_constant__ OutMaps omaps,
                                                                       const __grid_constant__ TileSched ts, int use_tma,
                                                                       int stages, int lag, long long total_tiles, int n_tiles,
                                                                       int sum_taps, int store_mode, int piece, int MT) {
  // MT (1 or 2) consecutive 128-row M tiles form one scheduling unit and share each K-block's weight tile: half the
  // weight traffic from L2 per output tile (in-kernel timing r01: the short-N layers wait on operand delivery).
  const int acc_cols = (d.bn_tile + 15) & ~15;                     // TMEM columns per accumulator
  const int b_tile_bytes = ((acc_cols * 128) + 1023) & ~1023;      // packed weight tile (layer-wide bn_tile)
  // K elements per tap in the packed weights.  TMA path: a multiple of 64 >= Cin -- the im2col box of the last channel
  // block reaches past Cin, where TMA zero-fills and the packed weights hold zeros.
  const int k_pitch = d.tap_pitch > 0 ? d.tap_pitch : d.Cin;
  constexpr int kBK = kTF32 ? 32 : 64;                             // K elements (channels) per 128-byte K-block row
  const int a_bytes = MT * kATileBytes;
  const int stage_bytes = a_bytes + b_tile_bytes;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  // [pipeline stages][epilogue staging: 2 halves x (128 rows x 128 B, 128B-swizzled)][barriers][tap tables][bias][scale]
  const uint32_t stg_base = smem_base + (uint32_t)stages * stage_bytes;                 // 1024-aligned
  FpropSmem* ps = reinterpret_cast<FpropSmem*>(smem_al + (size_t)stages * stage_bytes + kStagingBytes);
  int32_t* s_taps = reinterpret_cast<int32_t*>(ps + 1);            // all classes back to back
  // bias (Cout floats, zeros when the layer has none) and, when it fits, the per-(clip, channel) dropout scale
  float* s_bias = reinterpret_cast<float*>(smem_al + ((kStagingBytes + (size_t)stages * stage_bytes + sizeof(FpropSmem) + (size_t)sum_taps * 4 + 15) & ~(size_t)15));
  float* s_scale = s_bias + ((d.Cout + 3) & ~3);
  const bool scale_smem = d.scale_nc != nullptr && d.N * d.Cout <= kMaxScaleSmem;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  int tap_off[8];
  {
    int o = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      tap_off[c] = o;
      if (c < d.nclass) o += d.cls[c].ntaps;
    }
    for (int c = 0; c < d.nclass; ++c)
      for (int i = tid; i < d.cls[c].ntaps; i += kFpropThreads) s_taps[tap_off[c] + i] = d.cls[c].taps[i];
  }
  const FastDiv* s_fd = ps->fd;
  if (tid < 25) {
    uint32_t dv = (uint32_t)n_tiles;
    if (tid < 24) {
      const int c = tid / 3, k = tid - 3 * c;
      dv = 1;
      if (c < d.nclass) dv = (uint32_t)(k == 0 ? d.cls[c].Qw : k == 1 ? d.cls[c].Qh : d.cls[c].Qt);
    }
    ps->fd[tid] = make_fastdiv(dv < 1 ? 1u : dv);
  }
  for (int i = tid; i < d.Cout; i += kFpropThreads) s_bias[i] = d.bias ? d.bias[i] : 0.f;
  if (scale_smem)
    for (int i = tid; i < d.N * d.Cout; i += kFpropThreads) s_scale[i] = d.scale_nc[i];
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      // gather path: 128 gather threads + the thread that arms the weight bulk copy; TMA path: one producer thread
      mbar_init(&ps->full[s], use_tma ? 1 : 128 + 1);
      mbar_init(&ps->empty[s], 1);
    }
    mbar_init(&ps->tfull[0], 1);
    mbar_init(&ps->tfull[1], 1);
    mbar_init(&ps->tempty[0], 4);      // one elected lane of each of the 4 epilogue warps that drain this accumulator
    mbar_init(&ps->tempty[1], 4);
    fence_barrier_init();
  }
  const uint32_t tmem_cols = tmem_cols_for(2 * MT * acc_cols);
  if (warp == 4) tmem_alloc(&ps->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ps->tmem_base;

  if (warp < 4 && use_tma) {
    // ------------------------------ TMA producer (one elected thread) ----------------
    // A: im2col tensor map, one instruction per (tap, 64-channel block); B: bulk copy of the packed weight tile
    if (tid == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int cblocks = k_pitch / kBK;
      B2C_PROF_DECL(p_wait); B2C_PROF_DECL(p_t0); B2C_PROF_START(p_t0);
      for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileInfo ti = decode_tile(ts, d.nclass, t, n_tiles, s_fd[24], MT);
        const b2c_conv_class& cc = d.cls[ti.cls];
        const int32_t* taps = s_taps + tap_off[ti.cls];
        const unsigned Mtot = (unsigned)((long long)d.N * cc.Qt * cc.Qh * cc.Qw);
        int bw[2], bh[2], bt[2], bn_i[2];
        int nvalid = 0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const unsigned m0j = (unsigned)ti.m0 + (unsigned)(j * kTileM);
          bw[j] = bh[j] = bt[j] = bn_i[j] = 0;
          if (j < MT && m0j < Mtot) {
            uint32_t rw, rh, rt;
            uint32_t q = fdivmod(m0j, s_fd[3 * ti.cls], rw);
            q = fdivmod(q, s_fd[3 * ti.cls + 1], rh);
            q = fdivmod(q, s_fd[3 * ti.cls + 2], rt);
            bw[j] = (int)rw * d.si_w + cc.lo_w;
            bh[j] = (int)rh * d.si_h + cc.lo_h;
            bt[j] = (int)rt * d.si_t + cc.lo_t;
            bn_i[j] = (int)q;
            nvalid = j + 1;
          }
        }
        const int nkb = cc.ntaps * cblocks;
        // per-clip weight sets (collapsed decoder tail): the unit's rows all belong to clip bn_i[0] (host-checked)
        const uint8_t* wtile = reinterpret_cast<const uint8_t*>(cc.w) + (size_t)ti.n_idx * nkb * (size_t)b_tile_bytes +
                               (size_t)bn_i[0] * (size_t)d.w_sample_stride;
        const CUtensorMap* map = &maps.a[ti.cls];
        int tp_lo, tp_hi, tj_lo, tj_hi;
        {
          unsigned m_last = (unsigned)ti.m0 + (unsigned)(MT * kTileM) - 1u;
          if (m_last >= Mtot) m_last = Mtot - 1u;
          tap_range(d, cc, taps, (unsigned)ti.m0, m_last, s_fd + 3 * ti.cls, tp_lo, tp_hi, tj_lo, tj_hi);
        }
        const int hbk = cc.h_block < 0 ? -cc.h_block : 0;
        for (int tp = tp_lo; tp < tp_hi; ++tp) {
          if (hbk) {                       // inside-block sub-range (tap columns that are padding for the whole tile)
            const int j = tp % hbk;
            if (j < tj_lo || j >= tj_hi) continue;
          }
          int kb = tp * cblocks;
          const int32_t tv = taps[tp];
          const uint16_t ow = (uint16_t)(tap_dw(tv) - cc.lo_w), oh = (uint16_t)(tap_dh(tv) - cc.lo_h),
                         ot = (uint16_t)(tap_dt(tv) - cc.lo_t);
          for (int cb = 0; cb < cblocks; ++cb, ++kb) {
            B2C_PROF_DECL(w0); B2C_PROF_START(w0);
            mbar_wait(&ps->empty[stage], phase ^ 1, 1);
            B2C_PROF_ACC(p_wait, w0);
            const uint32_t a_st = smem_base + (uint32_t)stage * stage_bytes;
            mbar_arrive_expect_tx(&ps->full[stage], (uint32_t)(nvalid * kATileBytes + b_tile_bytes));
            tma_im2col_5d(a_st, map, &ps->full[stage], cb * kBK, bw[0], bh[0], bt[0], bn_i[0], ow, oh, ot);
            if (nvalid > 1)
              tma_im2col_5d(a_st + kATileBytes, map, &ps->full[stage], cb * kBK, bw[1], bh[1], bt[1], bn_i[1], ow, oh, ot);
            bulk_g2s(a_st + (uint32_t)a_bytes, wtile + (size_t)kb * b_tile_bytes, (uint32_t)b_tile_bytes, &ps->full[stage]);
            if (++stage == stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
      B2C_PROF_PRINT2("producer(tma)", p_t0, p_wait, 0);
    }
  } else if (warp < 4 && !kTF32) {
    // ------------------------------ gather producers (Cin not a multiple of 64; bf16 only) --------
    const int r = tid;
    const bf16* in_n = reinterpret_cast<const bf16*>(d.in) + d.in_c_off;
    const uint32_t a_row_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
    const uint32_t swz = (uint32_t)(r & 7);
    int stage = 0;
    uint32_t phase = 0;
    long long issued = 0;   // K-blocks issued so far by this CTA (all tiles)
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const TileInfo ti = decode_tile(ts, d.nclass, t, n_tiles, s_fd[24], MT);
      const b2c_conv_class& cc = d.cls[ti.cls];
      const int32_t* taps = s_taps + tap_off[ti.cls];
      const unsigned Mtot = (unsigned)((long long)d.N * cc.Qt * cc.Qh * cc.Qw);
      const unsigned m = (unsigned)ti.m0 + (unsigned)r;
      const bool mvalid = m < Mtot;
      int n_i = 0, qt = 0, qh = 0, qw = 0;
      if (mvalid) {
        uint32_t rw, rh, rt;
        uint32_t q = fdivmod(m, s_fd[3 * ti.cls], rw);
        q = fdivmod(q, s_fd[3 * ti.cls + 1], rh);
        q = fdivmod(q, s_fd[3 * ti.cls + 2], rt);
        qw = (int)rw; qh = (int)rh; qt = (int)rt;
        n_i = (int)q;
      }
      const int it0 = qt * d.si_t, ih0 = qh * d.si_h, iw0 = qw * d.si_w;
      const int K = cc.ntaps * d.Cin;
      const int nkb = (K + kBlockK - 1) / kBlockK;
      const uint8_t* wtile = reinterpret_cast<const uint8_t*>(cc.w) + (size_t)ti.n_idx * nkb * (size_t)b_tile_bytes;
      int tap = 0, c = 0;
      const bf16* tap_ptr = nullptr;
      auto load_tap = [&](int tp) {
        tap_ptr = nullptr;
        if (mvalid && tp < cc.ntaps) {
          const int32_t tv = taps[tp];
          const int it = it0 + tap_dt(tv), ih = ih0 + tap_dh(tv), iw = iw0 + tap_dw(tv);
          if ((unsigned)it < (unsigned)d.Ti && (unsigned)ih < (unsigned)d.Hi && (unsigned)iw < (unsigned)d.Wi)
            tap_ptr = in_n + ((((long long)n_i * d.Ti + it) * d.Hi + ih) * d.Wi + iw) * d.in_row_stride;
        }
      };
      load_tap(0);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&ps->empty[stage], phase ^ 1, 1);
        const uint32_t a_st = smem_base + (uint32_t)stage * stage_bytes;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t dst = a_st + a_row_off + (((uint32_t)j ^ swz) << 4);
          const void* src = tap_ptr ? (const void*)(tap_ptr + c) : (const void*)d.in;
          cp_async16(dst, src, tap_ptr ? 16u : 0u);
          c += 8;
          if (c >= d.Cin) {
            c = 0;
            ++tap;
            load_tap(tap);
          }
        }
        if (tid == 0) {
          mbar_arrive_expect_tx(&ps->full[stage], (uint32_t)b_tile_bytes);
          bulk_g2s(a_st + (uint32_t)a_bytes, wtile + (size_t)kb * b_tile_bytes, (uint32_t)b_tile_bytes, &ps->full[stage]);   // MT == 1
        }
        cp_async_commit();
        ++issued;
        if (issued > lag) {   // the K-block issued `lag` iterations ago has landed
          cp_async_wait_dyn(lag);
          fence_proxy_async_smem();
          int s2 = stage - lag;
          if (s2 < 0) s2 += stages;
          mbar_arrive(&ps->full[s2]);
        }
        if (++stage == stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    // drain the last `lag` K-blocks
    cp_async_wait<0>();
    fence_proxy_async_smem();
    {
      long long pending = issued < lag ? issued : lag;
      int s2 = stage - (int)pending;
      if (s2 < 0) s2 += stages;
      for (long long i = 0; i < pending; ++i) {
        mbar_arrive(&ps->full[s2]);
        if (++s2 == stages) s2 = 0;
      }
    }
  } else if (warp == 4) {
    // ------------------------------ MMA issuer -------------------------------------
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    B2C_PROF_DECL(m_wfull); B2C_PROF_DECL(m_wtempty); B2C_PROF_DECL(m_t0); B2C_PROF_START(m_t0);
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const TileInfo ti = decode_tile(ts, d.nclass, t, n_tiles, s_fd[24], MT);
      const b2c_conv_class& cc = d.cls[ti.cls];
      const int n0 = ti.n_idx * d.bn_tile;
      int bn = d.Cout - n0;
      if (bn > d.bn_tile) bn = d.bn_tile;
      const int bn16 = (bn + 15) & ~15;
      const uint32_t idesc = kTF32 ? umma_idesc_tf32(bn16, 0, 0) : umma_idesc_bf16(bn16, 0, 0);
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * MT * acc_cols);
      const unsigned Mtot = (unsigned)((long long)d.N * cc.Qt * cc.Qh * cc.Qw);
      int nkb = (cc.ntaps * k_pitch + kBK - 1) / kBK;
      if (use_tma && cc.h_block != 0) {
        int tp_lo, tp_hi, tj_lo, tj_hi;
        unsigned m_last = (unsigned)ti.m0 + (unsigned)(MT * kTileM) - 1u;
        if (m_last >= Mtot) m_last = Mtot - 1u;
        tap_range(d, cc, s_taps + tap_off[ti.cls], (unsigned)ti.m0, m_last, s_fd + 3 * ti.cls, tp_lo, tp_hi, tj_lo, tj_hi);
        int ntp = tp_hi - tp_lo;
        if (cc.h_block < 0) {
          const int hbk = -cc.h_block;
          const int jh = tj_hi < hbk ? tj_hi : hbk;
          ntp = (ntp / hbk) * (jh - tj_lo);
        }
        nkb = ntp * (k_pitch / kBK);
      }
      const int nvalid = (MT > 1 && (unsigned)ti.m0 + (unsigned)kTileM < Mtot) ? 2 : 1;
      B2C_PROF_DECL(w1); B2C_PROF_START(w1);
      mbar_wait(&ps->tempty[acc], acc_phase ^ 1, 4);   // epilogue has drained this accumulator
      B2C_PROF_ACC(m_wtempty, w1);
      tc_fence_after();
      for (int kb = 0; kb < nkb; ++kb) {
        B2C_PROF_DECL(w2); B2C_PROF_START(w2);
        mbar_wait(&ps->full[stage], phase, 2);
        B2C_PROF_ACC(m_wfull, w2);
        tc_fence_after();
        if (lane == 0) {
          // descriptors: only the 14-bit start-address field changes (16-byte units): +2 per 16-element K step,
          // +1024 per 16 KB A tile -- keep the single-thread issue path short, it paces the long-K layers
          const uint32_t a_st = smem_base + (uint32_t)stage * stage_bytes;
          const uint64_t ad0 = umma_desc_sw128(a_st, 16, 1024);
          const uint64_t bd0 = umma_desc_sw128(a_st + (uint32_t)a_bytes, 16, 1024);
          const uint32_t accf = kb > 0 ? 1u : 0u;
          umma_any<kTF32>(tmem_d, ad0, bd0, idesc, accf);
          umma_any<kTF32>(tmem_d, ad0 + 2, bd0 + 2, idesc, 1u);
          umma_any<kTF32>(tmem_d, ad0 + 4, bd0 + 4, idesc, 1u);
          umma_any<kTF32>(tmem_d, ad0 + 6, bd0 + 6, idesc, 1u);
          if (nvalid > 1) {
            const uint64_t ad1 = ad0 + (kATileBytes >> 4);
            const uint32_t tmem_d1 = tmem_d + (uint32_t)acc_cols;
            umma_any<kTF32>(tmem_d1, ad1, bd0, idesc, accf);
            umma_any<kTF32>(tmem_d1, ad1 + 2, bd0 + 2, idesc, 1u);
            umma_any<kTF32>(tmem_d1, ad1 + 4, bd0 + 4, idesc, 1u);
            umma_any<kTF32>(tmem_d1, ad1 + 6, bd0 + 6, idesc, 1u);
          }
          umma_commit(&ps->empty[stage]);
        }
        __syncwarp();
        if (++stage == stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (lane == 0) umma_commit(&ps->tfull[acc]);
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) { B2C_PROF_PRINT2("mma", m_t0, m_wfull, m_wtempty); }
  } else {
    // ------------------------------ epilogue (warps 5..12) --------------------------
    // Warp w may only touch TMEM lanes 32*(w&3)..+31.  The two warp quartets alternate TILES: quartet h drains
    // accumulator h, so each has two tile periods per tile.  bf16 row outputs go through a 128B-swizzled staging
    // buffer in 64-column chunks and leave with ONE TMA tensor store per warp and chunk (32 rows x 64 columns) when the
    // output rows of a tile are addressable by a tensor map (store_mode 1: rows contiguous; 2: strided classes in
    // `piece`-row pieces), else with coalesced 16-byte stores.  (ncu r01b/c + in-kernel phase timing: with per-lane
    // index arithmetic and per-row stores the epilogue took ~7000 cycles per 128x128 tile against ~1800 of MMA.)
    const int quarter = warp & 3;
    const int half = (warp - 5) >> 2;
    const bool staged = (d.out_fp32 == 0 && !d.accumulate);
    const bool use_store_tma = staged && store_mode != 0;
    // a 64-column box that is only partly owned by this N tile (bn_tile % 64 != 0 with several N tiles) must not be
    // written by TMA (it would spill into the neighbouring tile's columns): such chunks take the coalesced path
    const bool partial_chunks = n_tiles > 1 && (d.bn_tile & (kStageChunkCols - 1)) != 0;
    const bool need_pos = !use_store_tma || partial_chunks || d.scale_nc != nullptr;
    const uint32_t sub = stg_base + (uint32_t)half * (kTileM * 128) + (uint32_t)quarter * 4096;   // this warp's 32 rows
    const uint32_t my_row = sub + (uint32_t)lane * 128;
    const uint32_t swz = (uint32_t)(lane & 7);
    const bool do_relu = d.relu != 0;
    const float relu_lo = d.relu ? 0.f : __int_as_float(0xff800000);     // -inf: max.NaN(x, -inf) = x
    // global (not smem-resident) dropout scale rows also work with the fast body: it only dereferences the pointer
    const bool fast = staged && d.sigmoid_from < 0;
    const int sig_from = d.sigmoid_from >= 0 ? d.sigmoid_from : 0x7fffffff;
    const int n_issue = store_mode == 2 ? 32 / piece : 1;
    const int acc = half;
    B2C_PROF_DECL(e_wait); B2C_PROF_DECL(e_cols); B2C_PROF_DECL(e_t0); B2C_PROF_START(e_t0);
    B2C_PROF_DECL(e_rd); B2C_PROF_DECL(e_ld); B2C_PROF_DECL(e_st);
    long long it = 0;
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      if ((int)(it & 1) != half) continue;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const TileInfo ti = decode_tile(ts, d.nclass, t, n_tiles, s_fd[24], MT);
      const b2c_conv_class& cc = d.cls[ti.cls];
      const int n0 = ti.n_idx * d.bn_tile;
      int bn = d.Cout - n0;
      if (bn > d.bn_tile) bn = d.bn_tile;
      const int bn16 = (bn + 15) & ~15;
      const unsigned Mtot = (unsigned)((long long)d.N * cc.Qt * cc.Qh * cc.Qw);
      B2C_PROF_DECL(w3); B2C_PROF_START(w3);
      mbar_wait(&ps->tfull[acc], acc_phase, 3);
      B2C_PROF_ACC(e_wait, w3);
      B2C_PROF_DECL(w4); B2C_PROF_START(w4);
      tc_fence_after();
      for (int jt = 0; jt < MT; ++jt) {
        const unsigned m0j = (unsigned)ti.m0 + (unsigned)(jt * kTileM);
        if (m0j >= Mtot) break;
        const bool last_j = jt == MT - 1 || m0j + (unsigned)kTileM >= Mtot;   // last 128-row tile of this unit
        const unsigned mw = m0j + (unsigned)(quarter * 32);        // first row of this warp
        const unsigned m = mw + (unsigned)lane;
        const bool mvalid = m < Mtot;
        long long opos = 0;
        int n_i = 0;
        if (need_pos && mvalid) {
          uint32_t rw, rh, rt;
          uint32_t q = fdivmod(m, s_fd[3 * ti.cls], rw);
          q = fdivmod(q, s_fd[3 * ti.cls + 1], rh);
          q = fdivmod(q, s_fd[3 * ti.cls + 2], rt);
          n_i = (int)q;
          opos = (((long long)n_i * d.To + ((int)rt * d.so_t + cc.po_t)) * d.Ho + ((int)rh * d.so_h + cc.po_h)) * d.Wo +
                 ((int)rw * d.so_w + cc.po_w);
        }
        const float* scale_row = nullptr;
        if (d.scale_nc) scale_row = scale_smem ? s_scale + n_i * d.Cout : d.scale_nc + (long long)n_i * d.Cout;
        const uint32_t t_lane = tmem_base + (uint32_t)((acc * MT + jt) * acc_cols) + ((uint32_t)(quarter * 32) << 16);
        for (int ch0 = 0; ch0 < bn16; ch0 += kStageChunkCols) {
          const int cw = bn16 - ch0 < kStageChunkCols ? bn16 - ch0 : kStageChunkCols;
          // out_fold: output column c lives in T-plane (c / out_fold), channel (c % out_fold) (folded stem: all output
          // frames of a pixel are columns of one GEMM row); a 64-column chunk never straddles planes (out_fold % 64 == 0)
          const int fold_t = d.out_fold ? (n0 + ch0) / d.out_fold : 0;
          const int fold_c = d.out_fold ? fold_t * d.out_fold : 0;              // columns to subtract
          const long long fold_pos = (long long)fold_t * d.Ho * d.Wo;
          if (staged) {
            // the previous chunk of this warp has left its staging rows
            B2C_PROF_DECL(w5); B2C_PROF_START(w5);
            if (use_store_tma) {
              if (lane < n_issue) bulk_wait_read0();
            }
            __syncwarp();
            B2C_PROF_ACC(e_rd, w5);
          }
          if (fast) {
            // bf16 staged rows without sigmoid: branch-free body (s_bias / s_scale are padded past Cout; columns >= Cout
            // are never stored)
            for (int c0 = ch0; c0 < ch0 + cw; c0 += 32) {
              const uint32_t ta = t_lane + (uint32_t)c0;
              const float* bp = s_bias + n0 + c0;
              const float* sp = scale_row + n0 + c0;
              const uint32_t u0 = (uint32_t)((c0 - ch0) >> 3);
              if (ch0 + cw - c0 >= 32) {
                if (scale_row) epi_fast<true, 32>(ta, bp, sp, relu_lo, my_row, swz, u0);
                else epi_fast<false, 32>(ta, bp, sp, relu_lo, my_row, swz, u0);
              } else {
                if (scale_row) epi_fast<true, 16>(ta, bp, sp, relu_lo, my_row, swz, u0);
                else epi_fast<false, 16>(ta, bp, sp, relu_lo, my_row, swz, u0);
              }
            }
            if (last_j && ch0 + kStageChunkCols >= bn16) {
              // last TMEM read of this tile is complete: hand the accumulator back before the store phase
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&ps->tempty[acc]);
            }
          } else
          for (int c0 = ch0; c0 < ch0 + cw; c0 += 32) {
            uint32_t r[32];
            const int nh = (ch0 + cw - c0 >= 32) ? 4 : 2;
            if (nh == 4) tmem_ld32_issue(t_lane + (uint32_t)c0, r);
            else tmem_ld16_issue(t_lane + (uint32_t)c0, r);
            float4 bb[8], ss[8];
  #pragma unroll
            for (int h = 0; h < 4; ++h) {
              const int col = n0 + c0 + h * 8;
              const bool on = h < nh && col < d.Cout;
              bb[2 * h] = on ? *reinterpret_cast<const float4*>(s_bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
              bb[2 * h + 1] = on ? *reinterpret_cast<const float4*>(s_bias + col + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
              if (scale_row) {
                ss[2 * h] = on ? *reinterpret_cast<const float4*>(scale_row + col) : make_float4(1.f, 1.f, 1.f, 1.f);
                ss[2 * h + 1] = on ? *reinterpret_cast<const float4*>(scale_row + col + 4) : make_float4(1.f, 1.f, 1.f, 1.f);
              }
            }
            B2C_PROF_DECL(w6); B2C_PROF_START(w6);
            tmem_ld_fence(r);
            B2C_PROF_ACC(e_ld, w6);
            if (last_j && c0 + 32 >= bn16) {
              // last TMEM read of this tile is complete: hand the accumulator back before the (slow) store phase
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&ps->tempty[acc]);
            }
  #pragma unroll
            for (int h = 0; h < 4; ++h) {
              if (h >= nh) break;
              const int col = n0 + c0 + h * 8;
              float vv[8];
              const float4 b0 = bb[2 * h], b1 = bb[2 * h + 1];
              vv[0] = __uint_as_float(r[h * 8 + 0]) + b0.x; vv[1] = __uint_as_float(r[h * 8 + 1]) + b0.y;
              vv[2] = __uint_as_float(r[h * 8 + 2]) + b0.z; vv[3] = __uint_as_float(r[h * 8 + 3]) + b0.w;
              vv[4] = __uint_as_float(r[h * 8 + 4]) + b1.x; vv[5] = __uint_as_float(r[h * 8 + 5]) + b1.y;
              vv[6] = __uint_as_float(r[h * 8 + 6]) + b1.z; vv[7] = __uint_as_float(r[h * 8 + 7]) + b1.w;
              if (scale_row) {
                const float4 s0 = ss[2 * h], s1 = ss[2 * h + 1];
                vv[0] *= s0.x; vv[1] *= s0.y; vv[2] *= s0.z; vv[3] *= s0.w;
                vv[4] *= s1.x; vv[5] *= s1.y; vv[6] *= s1.z; vv[7] *= s1.w;
              }
              if (do_relu) {
  #pragma unroll
                for (int i = 0; i < 8; ++i) vv[i] = fmaxf(vv[i], 0.f);
              }
              if (col >= sig_from) {
  #pragma unroll
                for (int i = 0; i < 8; ++i) vv[i] = sigmoidf_(vv[i]);
              }
              if (staged) {
                const uint32_t u = (uint32_t)((c0 - ch0) >> 3) + (uint32_t)h;          // 16-byte unit within the 128-byte row
                st_shared16(my_row + ((u ^ swz) << 4), pack8(vv));
              } else if (!mvalid || col >= d.Cout) {
                // nothing to write
              } else if (d.out_fp32 == 2) {
                // planar fp32: out[channel][position]; consecutive lanes = consecutive positions -> coalesced
                float* o = reinterpret_cast<float*>(d.out) + (long long)(d.out_c_off + col) * d.out_row_stride + opos;
  #pragma unroll
                for (int i = 0; i < 8; ++i) {
                  if (d.accumulate) vv[i] += o[(long long)i * d.out_row_stride];
                  o[(long long)i * d.out_row_stride] = vv[i];
                }
              } else if (d.out_fp32) {
                float* o = reinterpret_cast<float*>(d.out) + (opos + fold_pos) * d.out_row_stride + d.out_c_off + col - fold_c;
                float4* o4 = reinterpret_cast<float4*>(o);
                if (d.accumulate) {
                  float4 a = o4[0], b = o4[1];
                  vv[0] += a.x; vv[1] += a.y; vv[2] += a.z; vv[3] += a.w;
                  vv[4] += b.x; vv[5] += b.y; vv[6] += b.z; vv[7] += b.w;
                }
                if (kTF32 && d.round_out) {   // the stored activation is the next GEMM's tf32 operand: round, do not truncate
  #pragma unroll
                  for (int i = 0; i < 8; ++i) vv[i] = tf32_rna(vv[i]);
                }
                o4[0] = make_float4(vv[0], vv[1], vv[2], vv[3]);
                o4[1] = make_float4(vv[4], vv[5], vv[6], vv[7]);
              } else {
                bf16* o = reinterpret_cast<bf16*>(d.out) + (opos + fold_pos) * d.out_row_stride + d.out_c_off + col - fold_c;
                uint4* o4 = reinterpret_cast<uint4*>(o);
                if (d.accumulate) {
                  float e[8];
                  unpack8(*o4, e);
  #pragma unroll
                  for (int i = 0; i < 8; ++i) vv[i] += e[i];
                }
                *o4 = pack8(vv);
              }
            }
          }
          if (!staged) continue;
          B2C_PROF_DECL(w7); B2C_PROF_START(w7);
          if (use_store_tma && (cw == kStageChunkCols || n0 + ch0 + kStageChunkCols >= d.Cout)) {
            fence_proxy_async_smem();
            __syncwarp();
            if (store_mode == 1) {
              // rows of the tile are consecutive output rows: one 32-row x 64-column box (clipped at Cout / Mtot by the map)
              if (lane == 0 && mw < Mtot) tma_store_2d(&omaps.o[0], sub, n0 + ch0, (int)mw);
            } else if (lane < n_issue) {
              // strided output class: `piece` consecutive positions never leave a W row
              const unsigned mp = mw + (unsigned)(lane * piece);
              if (mp < Mtot) {
                uint32_t rw, rh, rt;
                uint32_t q = fdivmod(mp, s_fd[3 * ti.cls], rw);
                q = fdivmod(q, s_fd[3 * ti.cls + 1], rh);
                q = fdivmod(q, s_fd[3 * ti.cls + 2], rt);
                tma_store_5d(&omaps.o[ti.cls], sub + (uint32_t)(lane * piece) * 128, n0 + ch0 - fold_c, (int)rw, (int)rh, (int)rt + fold_t, (int)q);
              }
            }
            if (lane < n_issue) bulk_commit();
          } else {
            // coalesced fallback: 8 lanes per row (16-byte units), 4 rows per instruction
            __syncwarp();
            const int ce = bn - ch0 < cw ? bn - ch0 : cw;            // real columns of this chunk
            const int segv = ce > 0 ? ce >> 3 : 0;
            const long long obyte = mvalid ? ((opos + fold_pos) * d.out_row_stride + d.out_c_off + n0 + ch0 - fold_c) * 2 : -1;
            const int vcol = lane & 7, rsub = lane >> 3;
  #pragma unroll
            for (int itr = 0; itr < 8; ++itr) {
              const int rl = itr * 4 + rsub;
              const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)(obyte & 0xffffffffll), rl);
              const int hi = __shfl_sync(0xffffffffu, (int)(obyte >> 32), rl);
              if (hi >= 0 && vcol < segv) {
                const uint4 val = ld_shared16(sub + (uint32_t)rl * 128 + (((uint32_t)vcol ^ (uint32_t)(rl & 7)) << 4));
                uint8_t* o = reinterpret_cast<uint8_t*>(d.out) + (((long long)hi << 32) | (long long)lo) + vcol * 16;
                *reinterpret_cast<uint4*>(o) = val;
              }
            }
          }
          B2C_PROF_ACC(e_st, w7);
        }
      }
      B2C_PROF_ACC(e_cols, w4);
    }
    if (use_store_tma && lane < n_issue) bulk_wait0();   // all tensor stores of this thread have completed before the CTA retires
    if (lane == 0 && (warp == 5 || warp == 9)) { B2C_PROF_PRINT2(warp == 5 ? "epilogue(w5)" : "epilogue(w9)", e_t0, e_wait, e_cols); }
    if (lane == 0 && warp == 5) { B2C_PROF_PRINT2("epi(w5) rd/ld", e_st, e_rd, e_ld); }
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// =====================================================================================
// wgrad kernel
// =====================================================================================
struct alignas(64) WgradMaps {
  CUtensorMap g;   // gathered operand (im2col over the taps), 64 positions x 64 channels per load
  CUtensorMap p;   // plain operand (same traversal, no taps)
};

// kTF32: fp32 tensors read as tf32.  The MN-major swizzle atom is 128 bytes of channels x 8 positions either way, so a
// TMA box holds 64 channels x 64 positions of bf16 or 32 channels x 32 positions of fp32 (the tf32 K-block is 32
// positions): stage bytes, LBO/SBO and the 4 MMAs per K-block stay the same, only the box geometry changes.
template <bool kTF32>
__global__ void __launch_bounds__(kThreads) igemm_wgrad_kernel(const __grid_constant__ b2c_wgrad_desc d,
                                                               const __grid_constant__ WgradMaps maps, int use_tma, int lo_t,
                                                               int lo_h, int lo_w, int stages, int lag, int nsplit, int MT) {
  // MT 128-column (tap, gc) tiles per CTA share ONE plain-operand tile per K-block (in-kernel timing r01: the MMA warp
  // waited on TMA data 52 % of the time at ~80 % of the L2 throughput cap; the plain tile is identical for all taps).
  constexpr int kCB = kTF32 ? 32 : 64;          // channels per box (= per 128-byte swizzle row)
  constexpr int kPK = kTF32 ? 32 : 64;          // positions per K-block
  constexpr int kBoxBytes = kPK * 128;          // one TMA box
  constexpr int kSubs = kTileM / kCB;           // boxes per 128-column (tap, channel) tile
  // TMA path: every tap owns ceil(Cg / kCB) channel blocks; the last one may reach past Cg, where TMA zero-fills
  // (layers whose channel count is not a multiple of the box width; the epilogue skips those columns)
  const int g_pitch = use_tma ? (d.Cg + kCB - 1) / kCB * kCB : d.Cg;
  const int K = d.ntaps * g_pitch;         // GEMM-M extent ((tap, gc) columns of the im2col matrix)
  const int mk0 = blockIdx.x * MT * kTileM;     // first (tap,gc) column of this CTA
  const int n0 = blockIdx.y * d.bn_tile;   // first p-channel
  int bn = d.Cp - n0;
  if (bn > d.bn_tile) bn = d.bn_tile;
  const int bn16 = (bn + 15) & ~15;
  const long long Mtot = (long long)d.N * d.Qt * d.Qh * d.Qw;  // positions = GEMM-K extent
  const long long nkb_all = (Mtot + kPK - 1) / kPK;
  const long long kb_lo = nkb_all * blockIdx.z / nsplit;
  const long long kb_hi = nkb_all * (blockIdx.z + 1) / nsplit;
  const int nkb = (int)(kb_hi - kb_lo);
  if (nkb <= 0) return;
  // per-clip gradients: nsplit is a multiple of N (host), so this CTA's positions belong to clip z / (nsplit / N)
  float* const dw_base = d.dw + (d.dw_sample_stride ? (long long)(blockIdx.z / (nsplit / d.N)) * d.dw_sample_stride : 0ll);
  const int b_atoms = (bn16 + kCB - 1) / kCB;
  const int b_tile_bytes = b_atoms * kBoxBytes;
  const int a_bytes = MT * kATileBytes;
  const int stage_bytes = a_bytes + b_tile_bytes;
  int mt_here = (K - mk0 + kTileM - 1) / kTileM;     // valid tiles of this CTA
  if (mt_here > MT) mt_here = MT;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  PipeSmem* ps = reinterpret_cast<PipeSmem*>(smem_al + (size_t)stages * stage_bytes);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&ps->full[s], use_tma ? 1 : 128);
      mbar_init(&ps->empty[s], 1);
    }
    mbar_init(&ps->accum, 1);
    fence_barrier_init();
  }
  const uint32_t tmem_cols = tmem_cols_for(MT * bn16);
  if (warp == 4) tmem_alloc(&ps->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = ps->tmem_base;

  if (warp < 4 && use_tma) {
    // ---- TMA producer: one thread; per stage 2 im2col boxes (64 positions x 64 channels) of the gathered operand for the
    // two 64-column halves of this CTA's 128 (tap, channel) columns, and bn/64 boxes of the plain operand
    if (tid == 0) {
      const int cblocks = g_pitch / kCB;
      constexpr int kMaxSub = 3 * kSubs;             // MT <= 3 tiles x kSubs boxes
      int sub_c[kMaxSub];
      uint16_t sub_ow[kMaxSub], sub_oh[kMaxSub], sub_ot[kMaxSub];
      bool sub_ok[kMaxSub];
      int n_ok = 0;
#pragma unroll
      for (int h = 0; h < kMaxSub; ++h) {
        const int blk = blockIdx.x * MT * kSubs + h;     // box index over (tap, channel block)
        sub_ok[h] = h < kSubs * MT && blk < d.ntaps * cblocks;
        const int tp = sub_ok[h] ? blk / cblocks : 0;
        sub_c[h] = (blk - tp * cblocks) * kCB;
        const int32_t tv = __ldg(d.taps + tp);
        sub_ow[h] = (uint16_t)(tap_dw(tv) - lo_w);
        sub_oh[h] = (uint16_t)(tap_dh(tv) - lo_h);
        sub_ot[h] = (uint16_t)(tap_dt(tv) - lo_t);
        n_ok += sub_ok[h] ? 1 : 0;
      }
      const int nb = b_atoms;
      const uint32_t tx = (uint32_t)((n_ok + nb) * kBoxBytes);
      long long pos = kb_lo * kPK;
      int n_i, qt, qh, qw;
      {
        unsigned q = (unsigned)pos;
        qw = (int)(q % (unsigned)d.Qw); q /= (unsigned)d.Qw;
        qh = (int)(q % (unsigned)d.Qh); q /= (unsigned)d.Qh;
        qt = (int)(q % (unsigned)d.Qt); q /= (unsigned)d.Qt;
        n_i = (int)q;
      }
      int stage = 0;
      uint32_t phase = 0;
      B2C_PROF_DECL(p_wait); B2C_PROF_DECL(p_t0); B2C_PROF_START(p_t0);
      for (int i = 0; i < nkb; ++i) {
        B2C_PROF_DECL(w0); B2C_PROF_START(w0);
        mbar_wait(&ps->empty[stage], phase ^ 1, 11);
        B2C_PROF_ACC(p_wait, w0);
        const uint32_t a_st = smem_base + (uint32_t)stage * stage_bytes;
        const uint32_t b_st = a_st + (uint32_t)a_bytes;
        mbar_arrive_expect_tx(&ps->full[stage], tx);
        const int gw = qw * d.sg_w + lo_w, gh = qh * d.sg_h + lo_h, gt = qt * d.sg_t + lo_t;
#pragma unroll
        for (int h = 0; h < kMaxSub; ++h)
          if (sub_ok[h])
            tma_im2col_5d(a_st + h * kBoxBytes, &maps.g, &ps->full[stage], sub_c[h], gw, gh, gt, n_i, sub_ow[h], sub_oh[h], sub_ot[h]);
        for (int j = 0; j < nb; ++j) {
          // p_fold: p-channel block c lives in T-plane (c / p_fold) at channel (c % p_fold) (folded stem)
          const int pc = n0 + j * kCB;
          const int ft = d.p_fold ? pc / d.p_fold : 0;
          tma_im2col_5d(b_st + j * kBoxBytes, &maps.p, &ps->full[stage], pc - ft * d.p_fold, qw * d.sp_w + d.pp_w, qh * d.sp_h + d.pp_h,
                        qt * d.sp_t + d.pp_t + ft, n_i, 0, 0, 0);
        }
        qw += kPK;
        while (qw >= d.Qw) {
          qw -= d.Qw;
          if (++qh == d.Qh) {
            qh = 0;
            if (++qt == d.Qt) {
              qt = 0;
              ++n_i;
            }
          }
        }
        if (++stage == stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (blockIdx.y == 0 && blockIdx.z == 0) { B2C_PROF_PRINTW("wgrad producer", p_t0, p_wait, nkb); }
    }
  } else if (warp < 4 && !kTF32) {
    // gather producers (bf16 only): thread -> position row r (0..63) and half (0/1) of the chunk columns
    const int r = tid & 63;
    const int half = tid >> 6;
    // fixed (tap, channel) of this thread's 8 A chunks
    int a_dt[8], a_dh[8], a_dw[8], a_c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kcol = mk0 + (half * 8 + j) * 8;
      if (kcol < K) {
        const int tp = kcol / d.Cg;
        const int32_t tv = __ldg(d.taps + tp);
        a_dt[j] = tap_dt(tv); a_dh[j] = tap_dh(tv); a_dw[j] = tap_dw(tv);
        a_c[j] = kcol - tp * d.Cg;
      } else {
        a_c[j] = -1; a_dt[j] = a_dh[j] = a_dw[j] = 0;
      }
    }
    const bf16* gbase = reinterpret_cast<const bf16*>(d.g) + d.g_c_off;
    const bf16* pbase = reinterpret_cast<const bf16*>(d.p) + d.p_c_off + n0;
    const int nbch = bn16 / 8;          // B chunks per row (even)
    const int bch_per = nbch / 2;       // per half
    const uint32_t row_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
    const uint32_t swz = (uint32_t)(r & 7);

    // position cursor of this thread's row; advanced by 64 positions per stage without divisions
    const long long pos0 = kb_lo * kBlockK + r;
    int n_i, qt, qh, qw;
    {
      long long t = pos0;
      qw = (int)(t % d.Qw); t /= d.Qw;
      qh = (int)(t % d.Qh); t /= d.Qh;
      qt = (int)(t % d.Qt); t /= d.Qt;
      n_i = (int)t;
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(&ps->empty[stage], phase ^ 1, 11);
      const uint32_t a_st = smem_base + (uint32_t)stage * stage_bytes;
      const uint32_t b_st = a_st + (uint32_t)a_bytes;      // gather path: MT == 1
      const bool pvalid = (pos0 + (long long)i * kBlockK) < Mtot;
      const int gt0 = qt * d.sg_t, gh0 = qh * d.sg_h, gw0 = qw * d.sg_w;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t dst = a_st + (uint32_t)(half * 8 * 1024) + row_off + (((uint32_t)j ^ swz) << 4);
        const void* src = d.g;
        uint32_t nb = 0;
        if (pvalid && a_c[j] >= 0) {
          const int it = gt0 + a_dt[j], ih = gh0 + a_dh[j], iw = gw0 + a_dw[j];
          if ((unsigned)it < (unsigned)d.Tg && (unsigned)ih < (unsigned)d.Hg && (unsigned)iw < (unsigned)d.Wg) {
            src = gbase + ((((long long)n_i * d.Tg + it) * d.Hg + ih) * d.Wg + iw) * d.g_row_stride + a_c[j];
            nb = 16;
          }
        }
        cp_async16(dst, src, nb);
      }
      const bf16* prow = nullptr;
      if (pvalid)
        prow = pbase + ((((long long)n_i * d.Tp + (qt * d.sp_t + d.pp_t)) * d.Hp + (qh * d.sp_h + d.pp_h)) * d.Wp +
                        (qw * d.sp_w + d.pp_w)) * d.p_row_stride;
      for (int j = 0; j < bch_per; ++j) {
        const int ch = half * bch_per + j;
        const uint32_t dst = b_st + (uint32_t)((ch >> 3) * 8 * 1024) + row_off + (((uint32_t)(ch & 7) ^ swz) << 4);
        const bool ok = pvalid && (n0 + ch * 8) < d.Cp;
        cp_async16(dst, ok ? (const void*)(prow + ch * 8) : d.p, ok ? 16u : 0u);
      }
      cp_async_commit();
      qw += kBlockK;
      while (qw >= d.Qw) {
        qw -= d.Qw;
        if (++qh == d.Qh) {
          qh = 0;
          if (++qt == d.Qt) {
            qt = 0;
            ++n_i;
          }
        }
      }
      if (i >= lag) {
        cp_async_wait_dyn(lag);
        fence_proxy_async_smem();
        int s2 = stage - lag;
        if (s2 < 0) s2 += stages;
        mbar_arrive(&ps->full[s2]);
      }
      if (++stage == stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    {
      int first = nkb - lag;
      if (first < 0) first = 0;
      for (int i = first; i < nkb; ++i) mbar_arrive(&ps->full[i % stages]);
    }
  }
  if (warp < 4) {
    // epilogue: TMEM lane = (tap,gc) column mk0 + tid ; columns = p channels
    mbar_wait(&ps->accum, 0, 13);
    B2C_PROF_DECL(e_t0); B2C_PROF_START(e_t0);
    tc_fence_after();
    for (int jt = 0; jt < mt_here; ++jt) {
      const int kcol = mk0 + jt * kTileM + tid;
      bool rvalid = kcol < K;
      long long base = 0;
      if (rvalid) {
        const int tp = kcol / g_pitch;
        const int gc = kcol - tp * g_pitch;
        rvalid = gc < d.Cg_real;
        base = (long long)gc * d.s_g + __ldg(d.wtap + tp);
      }
      const uint32_t t_lane = tmem_d + (uint32_t)(jt * bn16) + ((uint32_t)(warp * 32) << 16);
      for (int c0 = 0; c0 < bn16; c0 += 32) {
        float v[32];
        const int nc = (bn16 - c0 >= 32) ? 32 : 16;
        if (nc == 32) tmem_ld32(t_lane + (uint32_t)c0, v);
        else tmem_ld16(t_lane + (uint32_t)c0, v);
        if (!rvalid) continue;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= nc) break;
          const int pc = n0 + c0 + i;
          if (pc < d.Cp_real) {
            float* dst = dw_base + base + (long long)pc * d.s_p;
            if (d.nseg > 1) {      // fused layer: this column's member weight tensor
              int sgm = 0;
#pragma unroll
              for (int q = 1; q < 4; ++q)
                if (q < d.nseg && pc >= d.seg_begin[q]) sgm = q;
              dst = d.seg_dw[sgm] + base + (long long)(pc - d.seg_begin[sgm]) * d.s_p;
            }
            if (d.atomic) atomicAdd(dst, v[i]);
            else *dst = v[i];
          }
        }
      }
    }
    tc_fence_before();
    if (tid == 0 && blockIdx.y == 0 && blockIdx.z == 0) { B2C_PROF_PRINTW("wgrad epilogue", e_t0, 0, nkb); }
  } else {
    const uint32_t idesc = kTF32 ? umma_idesc_tf32(bn16, 1, 1) : umma_idesc_bf16(bn16, 1, 1);
    // descriptor start-address step per MMA (16-byte units): 16 positions x 128 B (bf16, K = 16) or 8 x 128 B (tf32, K = 8)
    constexpr uint32_t kStep = kTF32 ? 64u : 128u;
    int stage = 0;
    uint32_t phase = 0;
    B2C_PROF_DECL(m_wait); B2C_PROF_DECL(m_t0); B2C_PROF_START(m_t0);
    for (int i = 0; i < nkb; ++i) {
      B2C_PROF_DECL(w1); B2C_PROF_START(w1);
      mbar_wait(&ps->full[stage], phase, 12);
      B2C_PROF_ACC(m_wait, w1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_st = smem_base + (uint32_t)stage * stage_bytes;
        const uint32_t b_st = a_st + (uint32_t)a_bytes;
        // MN-major SW128: LBO = stride between MN atoms (= one box), SBO = stride between 8-position groups (1 KB).
        // Only the start-address field (16-byte units) changes: +kStep per MMA K step, +1024 per 16 KB A tile.
        const uint64_t ad0 = umma_desc_sw128(a_st, kBoxBytes, 1024);
        const uint64_t bd0 = umma_desc_sw128(b_st, kBoxBytes, 1024);
        const uint32_t accf = i > 0 ? 1u : 0u;
        for (int jt = 0; jt < mt_here; ++jt) {
          const uint64_t ad = ad0 + (uint64_t)(jt * (kATileBytes >> 4));
          const uint32_t td = tmem_d + (uint32_t)(jt * bn16);
          umma_any<kTF32>(td, ad, bd0, idesc, accf);
          umma_any<kTF32>(td, ad + kStep, bd0 + kStep, idesc, 1u);
          umma_any<kTF32>(td, ad + 2 * kStep, bd0 + 2 * kStep, idesc, 1u);
          umma_any<kTF32>(td, ad + 3 * kStep, bd0 + 3 * kStep, idesc, 1u);
        }
        umma_commit(&ps->empty[stage]);
      }
      __syncwarp();
      if (++stage == stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (lane == 0) umma_commit(&ps->accum);
    __syncwarp();
    if (lane == 0 && blockIdx.y == 0 && blockIdx.z == 0) { B2C_PROF_PRINTW("wgrad mma", m_t0, m_wait, nkb); }
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_d, tmem_cols);
  }
}

// ---------------------------------------------------------------------------------
// fp32 master weights -> bf16 GEMM operand tiles in exactly the shared-memory image the fprop kernel wants:
// [n-tile][k-block][row (bn_tile)][64 K-elements] with the 128-byte swizzle already applied, so a stage's
// weight tile is ONE contiguous bulk copy.  k = t*tap_pitch + col_off + c ; global row = r + r_off.
template <bool kTF32>
__global__ void pack_weights_kernel(const float* __restrict__ w, void* __restrict__ packed, const int32_t* __restrict__ wtap,
                                    int R, int ntaps, int C, int C_real, long long s_r, long long s_c, long long tap_pitch,
                                    long long col_off, int r_off, int bn_tile, int nkb) {
  const long long total = (long long)R * ntaps * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long t2 = i / C;
    const int t = (int)(t2 % ntaps);
    const int r = (int)(t2 / ntaps);
    float v = 0.f;
    if (c < C_real) v = w[(long long)r * s_r + (long long)c * s_c + wtap[t]];
    const int rg = r + r_off;
    const int tile = rg / bn_tile, rr = rg - tile * bn_tile;
    const long long k = (long long)t * tap_pitch + col_off + c;
    pack_store<kTF32>(packed, 0, rr, k, v, bn_tile, nkb, tile);
  }
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeIm2colFn get_encode_im2col() {
  static EncodeIm2colFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(p);
  }
  return fn;
}

// im2col tensor map over a channels-last bf16 view (C, W, H, T, N).  The bounding box of BASE pixels is
// [lo, lo + (Q-1)*stride] per spatial dim (corner arrays in W,H,D order, like CUTLASS); filter taps are passed as
// non-negative offsets (d_tap - lo) in the instruction.  Out-of-tensor reads are zero filled (= padding).
int encode_im2col_map(CUtensorMap* map, const void* base, int C, long long row_stride, int N, int T, int H, int W, int lo_t,
                      int lo_h, int lo_w, int Qt, int Qh, int Qw, int st, int sh, int sw, int channels, int pixels, int esz = 2) {
  EncodeIm2colFn fn = get_encode_im2col();
  if (!fn) return b2c_fail(-2, "cuTensorMapEncodeIm2col entry point not available in this driver");
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)N};
  const cuuint64_t rsb = (cuuint64_t)row_stride * (cuuint64_t)esz;      // bytes per pixel row
  cuuint64_t strides[4] = {rsb, rsb * W, rsb * W * H, rsb * W * H * T};
  int lower[3] = {lo_w, lo_h, lo_t};
  int upper[3] = {lo_w + (Qw - 1) * sw - (W - 1), lo_h + (Qh - 1) * sh - (H - 1), lo_t + (Qt - 1) * st - (T - 1)};
  for (int i = 0; i < 3; ++i)
    if (lower[i] < -16 || lower[i] > 15 || upper[i] < -16 || upper[i] > 15)
      return b2c_fail(-1, "im2col corner out of range: lower %d upper %d (dim %d)", lower[i], upper[i], i);
  cuuint32_t estr[5] = {1, (cuuint32_t)sw, (cuuint32_t)sh, (cuuint32_t)st, 1};
  CUresult r = fn(map, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, lower, upper,
                  (cuuint32_t)channels, (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return b2c_fail(-3, "cuTensorMapEncodeIm2col failed with CUresult %d", (int)r);
  // CUTLASS (copy_traits_sm90_im2col.hpp) clears this descriptor bit for tensors < 128 KiB on drivers <= 13.1
  static int drv = -1;
  if (drv < 0) cudaDriverGetVersion(&drv);
  const unsigned long long bytes = (unsigned long long)rsb * W * H * T * N;
  if (drv <= 13010 && bytes < 131072ull) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// Tiled map of a bf16 output view for the TMA-store epilogue: dims (C, Qw, Qh, Qt, N) with the class strides
// (rank 5), or (C, rows) for outputs whose tile rows are consecutive (rank 2).  128B swizzle, 64-channel boxes.
int encode_out_map(CUtensorMap* map, const bf16* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_tiled();
  if (!fn) return b2c_fail(-2, "cuTensorMapEncodeTiled entry point not available in this driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<bf16*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return b2c_fail(-3, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// All weight re-packing of a step in ONE launch: a device table of jobs (stable pointers: flat parameter buffer
// views -> persistent packed operand buffers); block -> job through a prefix table.
__global__ void pack_weights_batched_kernel(const b2c_pack_job* __restrict__ jobs, const int* __restrict__ block_start, int njobs) {
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {                 // last job whose first block <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if (block_start[mid] <= (int)blockIdx.x) lo = mid;
    else hi = mid - 1;
  }
  const b2c_pack_job J = jobs[lo];
  const int nblk = block_start[lo + 1] - block_start[lo];
  const int bidx = blockIdx.x - block_start[lo];
  const float* __restrict__ w = reinterpret_cast<const float*>(J.w);
  void* __restrict__ packed = J.packed;
  const int32_t* __restrict__ wtap = J.wtap;
  // (host guarantees R * ntaps * C < 2^31: 32-bit index arithmetic)
  const unsigned total = (unsigned)J.R * (unsigned)J.ntaps * (unsigned)J.C;
  const unsigned C = (unsigned)J.C, ntaps = (unsigned)J.ntaps, bn = (unsigned)J.bn_tile;
  if (J.dtype == 0 && (C & 7u) == 0 && (J.tap_pitch & 7) == 0 && (J.col_off & 7) == 0) {
    // bf16 operands, 8 consecutive K elements per thread = one 16-byte unit of the swizzled image: one index split and one
    // store per 8 elements (the element-wise loop below was instruction bound: 0.41 ms for the step's 96 M elements)
    const unsigned C8 = C >> 3, total8 = (unsigned)J.R * ntaps * C8;
    bf16* __restrict__ pk = reinterpret_cast<bf16*>(packed);
    for (unsigned i = (unsigned)bidx * blockDim.x + threadIdx.x; i < total8; i += (unsigned)nblk * blockDim.x) {
      const unsigned t2 = i / C8, c = (i - t2 * C8) << 3;
      const unsigned r = t2 / ntaps, t = t2 - r * ntaps;
      const float* src = w + (long long)r * J.s_r + (long long)c * J.s_c + wtap[t];
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ((int)c + j < J.C_real) ? __ldg(src + (long long)j * J.s_c) : 0.f;
      const unsigned rg = r + (unsigned)J.r_off;
      const unsigned tile = rg / bn, rr = rg - tile * bn;
      const long long k = (long long)t * J.tap_pitch + J.col_off + c;
      const int kb = (int)(k >> 6), kk = (int)(k & 63);
      const long long off = ((long long)tile * J.nkb + kb) * ((long long)J.bn_tile * 64) + (rr >> 3) * 512 + (rr & 7) * 64 +
                            (((kk >> 3) ^ (rr & 7)) << 3);
      *reinterpret_cast<uint4*>(pk + off) = pack8(v);
    }
    return;
  }
  for (unsigned i = (unsigned)bidx * blockDim.x + threadIdx.x; i < total; i += (unsigned)nblk * blockDim.x) {
    const unsigned t2 = i / C, c = i - t2 * C;
    const unsigned r = t2 / ntaps, t = t2 - r * ntaps;
    float v = 0.f;
    if ((int)c < J.C_real) v = __ldg(w + (long long)r * J.s_r + (long long)c * J.s_c + wtap[t]);
    const unsigned rg = r + (unsigned)J.r_off;
    const unsigned tile = rg / bn, rr = rg - tile * bn;
    const long long k = (long long)t * J.tap_pitch + J.col_off + c;
    if (J.dtype) pack_store<true>(packed, 0, (int)rr, k, v, J.bn_tile, J.nkb, (int)tile);
    else pack_store<false>(packed, 0, (int)rr, k, v, J.bn_tile, J.nkb, (int)tile);
  }
}

int pick_bn_tile(int Cout) {
  if (Cout <= 256) return (Cout + 15) & ~15;
  const int nt = (Cout + 255) / 256;
  int t = (Cout + nt - 1) / nt;
  return (t + 15) & ~15;
}

constexpr int kCtaSmemTarget = 72 * 1024;    // ~3 CTAs per SM (2 for 256-wide tiles)

}  // namespace

// 128-row tiles per SM above which two tiles are paired into one scheduling unit (B2C_MT_MIN_TILES_PER_SM overrides)
static int mt_min_tiles_per_sm() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B2C_MT_MIN_TILES_PER_SM");
    v = e ? atoi(e) : 8;
    if (v < 1) v = 1;
  }
  return v;
}

B2C_API int b2c_conv_fprop(const b2c_conv_desc* dh, b2c_stream_t stream) {
  B2C_REQUIRE(dh != nullptr, "conv_fprop: null descriptor");
  b2c_conv_desc d = *dh;
  B2C_REQUIRE(d.in && d.out, "conv_fprop: null tensor");
  B2C_REQUIRE(d.Cin > 0 && d.Cin % 8 == 0, "conv_fprop: Cin=%d must be a positive multiple of 8", d.Cin);
  B2C_REQUIRE(d.Cout > 0 && d.Cout % 8 == 0, "conv_fprop: Cout=%d must be a positive multiple of 8", d.Cout);
  B2C_REQUIRE(d.in_c_off % 8 == 0 && d.out_c_off % 8 == 0, "conv_fprop: channel offsets must be multiples of 8");
  B2C_REQUIRE(d.in_row_stride % 8 == 0 && (d.out_fp32 == 2 || d.out_row_stride % (d.out_fp32 ? 4 : 8) == 0),
              "conv_fprop: row strides unaligned");
  B2C_REQUIRE(d.nclass >= 1 && d.nclass <= 8, "conv_fprop: nclass=%d out of range", d.nclass);
  B2C_REQUIRE(((uintptr_t)d.in & 15) == 0 && ((uintptr_t)d.out & 15) == 0, "conv_fprop: tensors must be 16B aligned");
  if (d.bn_tile <= 0) d.bn_tile = pick_bn_tile(d.Cout);
  B2C_REQUIRE(d.bn_tile % 16 == 0 && d.bn_tile <= 256, "conv_fprop: bn_tile=%d invalid", d.bn_tile);
  int sum_taps = 0;
  long long tiles128 = 0;
  const int n_tiles = (d.Cout + d.bn_tile - 1) / d.bn_tile;
  for (int i = 0; i < d.nclass; ++i) {
    const b2c_conv_class& c = d.cls[i];
    B2C_REQUIRE(c.taps && c.w && c.ntaps > 0, "conv_fprop: class %d incomplete", i);
    B2C_REQUIRE(((uintptr_t)c.w & 15) == 0, "conv_fprop: weights must be 16B aligned");
    const long long m = (long long)d.N * c.Qt * c.Qh * c.Qw;
    B2C_REQUIRE(m < (1LL << 31), "conv_fprop: GEMM-M too large");
    tiles128 += ((m + kTileM - 1) / kTileM) * n_tiles;
    sum_taps += c.ntaps;
  }
  if (tiles128 == 0) return 0;
  const int k_pitch = d.tap_pitch > 0 ? d.tap_pitch : d.Cin;
  const int tf32 = d.dtype == 1 ? 1 : 0;
  B2C_REQUIRE(d.dtype == 0 || d.dtype == 1, "conv_fprop: dtype=%d", d.dtype);
  const int bk = tf32 ? 32 : kBlockK;         // K elements per 128-byte K-block row
  const int esz = tf32 ? 4 : 2;
  B2C_REQUIRE(k_pitch >= d.Cin && (k_pitch == d.Cin || (k_pitch % bk == 0 && k_pitch - d.Cin < bk)),
              "conv_fprop: tap_pitch=%d must be Cin=%d or Cin rounded up to a multiple of %d", d.tap_pitch, d.Cin, bk);
  const int use_tma = (k_pitch % bk == 0) ? 1 : 0;
  B2C_REQUIRE(!tf32 || (use_tma && d.out_fp32 != 0), "conv_fprop: tf32 mode needs tap_pitch %% 32 == 0 and an fp32 output");
  for (int i = 0; i < d.nclass; ++i)
    if (!use_tma || (d.cls[i].h_block != 0 && d.cls[i].ntaps % (d.cls[i].h_block < 0 ? -d.cls[i].h_block : d.cls[i].h_block) != 0))
      d.cls[i].h_block = 0;
  if (d.out_fold != 0) {
    B2C_REQUIRE(d.out_fold % kStageChunkCols == 0 && d.Cout % d.out_fold == 0 && d.bn_tile % d.out_fold == 0 && d.nclass == 1 &&
                    d.out_fp32 != 2 && d.cls[0].Qt == 1 && d.so_t == 1 && d.cls[0].po_t + d.Cout / d.out_fold <= d.To && !d.scale_nc,
                "conv_fprop: out_fold=%d needs 64-column planes, one class with Qt = 1 and Cout / out_fold <= To", d.out_fold);
  }
  if (d.w_sample_stride != 0) {
    B2C_REQUIRE(use_tma && d.w_sample_stride > 0 && d.w_sample_stride % 16 == 0, "conv_fprop: per-clip weights need the TMA path and a 16-byte stride");
    for (int i = 0; i < d.nclass; ++i)
      B2C_REQUIRE(((long long)d.cls[i].Qt * d.cls[i].Qh * d.cls[i].Qw) % (2 * kTileM) == 0,
                  "conv_fprop: per-clip weights need positions per clip to be a multiple of 256 (class %d)", i);
  }
  const int acc_cols = (d.bn_tile + 15) & ~15;
  const int b_tile_bytes_h = ((acc_cols * 128) + 1023) & ~1023;
  const int staging = kStagingBytes + 16;
  const int vecs = (((d.Cout + 3) & ~3) + ((d.scale_nc && d.N * d.Cout <= kMaxScaleSmem) ? d.N * d.Cout : 0)) * 4;
  const int fixed = (int)sizeof(FpropSmem) + sum_taps * 4 + staging + vecs + 1024 + 64;
  // two 128-row tiles per scheduling unit (shared weight tile) when TMEM (2 x 2 x acc_cols <= 512 columns), shared
  // memory (>= 3 stages) and the amount of work (>= 4 units per SM) allow it
  int MT = 1;
  if (use_tma && 4 * acc_cols <= 512 && (224 * 1024 - fixed) / (2 * kATileBytes + b_tile_bytes_h) >= 3 &&
      tiles128 >= (long long)mt_min_tiles_per_sm() * b2c_num_sms())
    MT = 2;
  long long total_tiles = 0;
  TileSched ts;
  memset(&ts, 0, sizeof(ts));
  for (int i = 0; i < d.nclass; ++i) {
    ts.start[i] = (unsigned)total_tiles;
    const b2c_conv_class& c = d.cls[i];
    const long long m = (long long)d.N * c.Qt * c.Qh * c.Qw;
    total_tiles += ((m + (long long)kTileM * MT - 1) / ((long long)kTileM * MT)) * n_tiles;
  }
  for (int i = d.nclass; i < 9; ++i) ts.start[i] = (unsigned)total_tiles;
  B2C_REQUIRE(total_tiles < (1LL << 31), "conv_fprop: too many tiles");
  if (d.nclass > 1 && (d.nclass & (d.nclass - 1)) == 0) {
    bool same = true;
    for (int i = 1; i < d.nclass; ++i) same = same && (ts.start[i + 1] - ts.start[i] == ts.start[1] - ts.start[0]);
    ts.interleave = same ? 1u : 0u;
  }
  const int stage_bytes = MT * kATileBytes + b_tile_bytes_h;
  int stages = (224 * 1024 - fixed) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  B2C_REQUIRE(stages >= 3, "conv_fprop: tile too large for shared memory");
  int lag = stages - 2;
  if (lag > 6) lag = 6;
  const size_t smem = (size_t)stages * stage_bytes + fixed;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(igemm_fprop_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(224 * 1024));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(igemm_fprop_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(224 * 1024));
    if (e != cudaSuccess) return b2c_cuda_check(e, "conv_fprop: cudaFuncSetAttribute");
    configured = true;
  }
  TmaMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (use_tma) {
    for (int i = 0; i < d.nclass; ++i) {
      const b2c_conv_class& c = d.cls[i];
      int rc = encode_im2col_map(&maps.a[i], reinterpret_cast<const uint8_t*>(d.in) + (size_t)d.in_c_off * esz, d.Cin,
                                 d.in_row_stride, d.N, d.Ti, d.Hi, d.Wi, c.lo_t, c.lo_h, c.lo_w, c.Qt, c.Qh, c.Qw, d.si_t, d.si_h,
                                 d.si_w, bk, kTileM, esz);
      if (rc) return rc;
    }
  }
  // TMA-store epilogue for bf16 row outputs (see the kernel's epilogue comment)
  OutMaps omaps;
  memset(&omaps, 0, sizeof(omaps));
  int store_mode = 0, piece = 32;
  if (d.out_fp32 == 0 && !d.accumulate && get_encode_tiled() != nullptr) {
    const bf16* obase = reinterpret_cast<const bf16*>(d.out) + d.out_c_off;
    const b2c_conv_class& c0 = d.cls[0];
    const bool contiguous = d.out_fold == 0 && d.nclass == 1 && d.so_t == 1 && d.so_h == 1 && d.so_w == 1 && c0.po_t == 0 && c0.po_h == 0 &&
                            c0.po_w == 0 && c0.Qt == d.To && c0.Qh == d.Ho && c0.Qw == d.Wo;
    if (contiguous) {
      cuuint64_t dims[2] = {(cuuint64_t)d.Cout, (cuuint64_t)((long long)d.N * d.To * d.Ho * d.Wo)};
      cuuint64_t strides[1] = {(cuuint64_t)d.out_row_stride * 2};
      cuuint32_t box[2] = {(cuuint32_t)kStageChunkCols, 32};
      int rc = encode_out_map(&omaps.o[0], obase, 2, dims, strides, box);
      if (rc) return rc;
      store_mode = 1;
    } else {
      int g = 32;
      for (int i = 0; i < d.nclass; ++i)
        while (d.cls[i].Qw % g) g >>= 1;
      if (g >= 8) {
        for (int i = 0; i < d.nclass; ++i) {
          const b2c_conv_class& c = d.cls[i];
          const long long rs = d.out_row_stride;
          cuuint64_t dims[5] = {(cuuint64_t)(d.out_fold ? d.out_fold : d.Cout), (cuuint64_t)c.Qw, (cuuint64_t)c.Qh,
                                (cuuint64_t)(d.out_fold ? d.To : c.Qt), (cuuint64_t)d.N};
          cuuint64_t strides[4] = {(cuuint64_t)(d.so_w * rs * 2), (cuuint64_t)((long long)d.so_h * d.Wo * rs * 2),
                                   (cuuint64_t)((long long)d.so_t * d.Ho * d.Wo * rs * 2),
                                   (cuuint64_t)((long long)d.To * d.Ho * d.Wo * rs * 2)};
          cuuint32_t box[5] = {(cuuint32_t)kStageChunkCols, (cuuint32_t)g, 1, 1, 1};
          const bf16* cb = obase + (((long long)c.po_t * d.Ho + c.po_h) * d.Wo + c.po_w) * rs;
          int rc = encode_out_map(&omaps.o[i], cb, 5, dims, strides, box);
          if (rc) return rc;
        }
        store_mode = 2;
        piece = g;
      }
    }
  }
  long long grid = b2c_num_sms();
  // class-minor order: an odd grid makes every CTA cycle through all classes (their K lengths differ 8x), instead of
  // alternating between two of them (148 % 8 == 4)
  if (ts.interleave && (grid & 1) == 0) --grid;
  if (grid > total_tiles) grid = total_tiles;
  if (tf32)
    igemm_fprop_kernel<true><<<(unsigned)grid, kFpropThreads, smem, (cudaStream_t)stream>>>(d, maps, omaps, ts, use_tma, stages, lag,
                                                                                            total_tiles, n_tiles, sum_taps, store_mode, piece, MT);
  else
    igemm_fprop_kernel<false><<<(unsigned)grid, kFpropThreads, smem, (cudaStream_t)stream>>>(d, maps, omaps, ts, use_tma, stages, lag,
                                                                                             total_tiles, n_tiles, sum_taps, store_mode, piece, MT);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("conv_fprop launch");
  return 0;
}

// grid size of the position split in waves of CTAs (B2C_WGRAD_WAVES overrides)
static int wgrad_waves() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B2C_WGRAD_WAVES");
    v = e ? atoi(e) : 1;      // measured on the 8+8 step: 1 wave 36.9 ms, 2 waves 37.4, 3 waves 37.8, 4 waves 38.3
    if (v < 1) v = 1;
  }
  return v;
}

B2C_API int b2c_conv_wgrad(const b2c_wgrad_desc* dh, b2c_stream_t stream) {
  B2C_REQUIRE(dh != nullptr, "conv_wgrad: null descriptor");
  b2c_wgrad_desc d = *dh;
  B2C_REQUIRE(d.g && d.p && d.dw && d.taps && d.wtap, "conv_wgrad: null pointer");
  B2C_REQUIRE(d.nseg >= 0 && d.nseg <= 4 && (d.nseg == 0 || d.dw_sample_stride == 0), "conv_wgrad: bad weight-tensor segments");
  for (int i = 0; i < d.nseg; ++i)
    B2C_REQUIRE(d.seg_dw[i] && d.seg_begin[i] >= (i ? d.seg_begin[i - 1] : 0) && (i || d.seg_begin[0] == 0),
                "conv_wgrad: segment %d invalid", i);
  B2C_REQUIRE(d.Cg > 0 && d.Cg % 8 == 0 && d.Cp > 0 && d.Cp % 8 == 0, "conv_wgrad: channels must be multiples of 8");
  B2C_REQUIRE(d.g_c_off % 8 == 0 && d.p_c_off % 8 == 0 && d.g_row_stride % 8 == 0 && d.p_row_stride % 8 == 0,
              "conv_wgrad: unaligned view");
  B2C_REQUIRE(((uintptr_t)d.g & 15) == 0 && ((uintptr_t)d.p & 15) == 0, "conv_wgrad: tensors must be 16B aligned");
  if (d.Cg_real <= 0) d.Cg_real = d.Cg;
  if (d.Cp_real <= 0 || d.Cp_real > d.Cp) d.Cp_real = d.Cp;
  if (d.bn_tile <= 0) d.bn_tile = pick_bn_tile(d.Cp);
  B2C_REQUIRE(d.bn_tile % 16 == 0 && d.bn_tile <= 256, "conv_wgrad: bn_tile=%d invalid", d.bn_tile);
  const long long Mtot = (long long)d.N * d.Qt * d.Qh * d.Qw;
  if (Mtot == 0) return 0;
  // kind::tf32 reads MN-major (position-major) 32-bit operands only from the 128B-swizzle / 32-byte-base shared-memory
  // layout, which the im2col TMA path used here does not produce (measured r02: the kTF32 instantiation returned zeros).
  // The tf32 precision mode therefore computes weight gradients as three bf16 GEMMs on hi / lo splits (b2c_split_bf16).
  B2C_REQUIRE(d.dtype == 0, "conv_wgrad: dtype=%d -- only bf16 operands; tf32 mode passes bf16 hi/lo splits (b2c_split_bf16)", d.dtype);
  const int tf32 = 0;
  const int esz = tf32 ? 4 : 2;
  const int cb = tf32 ? 32 : 64;           // channels per TMA box (128-byte swizzle row)
  const int pk = tf32 ? 32 : 64;           // positions per K-block
  const int box_bytes = pk * 128;
  // TMA path whenever the host tap table is there: channel counts that are not multiples of the box width are handled by
  // TMA's zero fill past the tensor extent (B2C_TMA_TAIL=0 restores the cp.async gather path for those layers)
  static int tail = -1;
  if (tail < 0) {
    const char* e = getenv("B2C_TMA_TAIL");
    tail = (e && atoi(e) == 0) ? 0 : 1;
  }
  const int use_tma = (d.taps_host != nullptr && (tail || tf32 || (d.Cg % cb == 0 && d.Cp % cb == 0))) ? 1 : 0;
  B2C_REQUIRE(!tf32 || use_tma, "conv_wgrad: tf32 mode needs the host tap table (TMA path)");
  const int g_pitch = use_tma ? (d.Cg + cb - 1) / cb * cb : d.Cg;
  const int K = d.ntaps * g_pitch;
  const int mt = (K + kTileM - 1) / kTileM;
  const int nt = (d.Cp + d.bn_tile - 1) / d.bn_tile;
  const long long nkb = (Mtot + pk - 1) / pk;
  const int bn16 = (d.bn_tile + 15) & ~15;
  const int b_tile_bytes = ((bn16 + cb - 1) / cb) * box_bytes;
  // (tap, gc) tiles per CTA sharing one plain-operand tile: as many as TMEM (512 columns) and >= 3 pipeline stages allow
  int MT = 1;
  if (use_tma) {
    for (int cand = 3; cand >= 2; --cand)
      if (cand <= mt && cand * bn16 <= 512 && (200 * 1024) / (cand * kATileBytes + b_tile_bytes) >= 3) {
        MT = cand;
        break;
      }
  }
  const int mtc = (mt + MT - 1) / MT;      // CTAs along the (tap, gc) dimension
  int nsplit = d.nsplit;
  if (nsplit <= 0) {
    // long position loops with a deep pipeline, one CTA per SM: pick the split so the grid is at most one full wave
    const long long tiles = (long long)mtc * nt;
    const long long sms = b2c_num_sms();
    const long long waves = wgrad_waves();
    long long want = (waves * sms + tiles - 1) / tiles;
    if (tiles * want > waves * sms && want > 1) --want;
    long long cap = nkb / 8;
    if (cap < 1) cap = 1;
    nsplit = (int)(want < cap ? want : cap);
    if (nsplit < 1) nsplit = 1;
  }
  if (nsplit > nkb) nsplit = (int)nkb;
  if (d.dw_sample_stride != 0) {
    // per-clip gradients: the position split must not straddle clips
    const long long per = (long long)d.Qt * d.Qh * d.Qw;
    B2C_REQUIRE(d.dw_sample_stride > 0 && per % pk == 0, "conv_wgrad: per-clip dw needs positions per clip %% %d == 0", pk);
    int s = nsplit / d.N;
    if (s < 1) s = 1;
    while (s > 1 && (long long)s * 8 > per / pk) --s;
    nsplit = s * d.N;
  }
  B2C_REQUIRE(nsplit == 1 || d.atomic, "conv_wgrad: nsplit>1 requires atomic accumulation");
  if (d.p_fold != 0)
    B2C_REQUIRE(use_tma && d.p_fold % cb == 0 && d.Cp % d.p_fold == 0 && d.Qt == 1 && d.sp_t == 1 && d.pp_t + d.Cp / d.p_fold <= d.Tp &&
                    ((long long)d.Qh * d.Qw) % pk == 0,
                "conv_wgrad: p_fold=%d needs the TMA path, Qt = 1 and planes of a multiple of %d positions", d.p_fold, pk);
  const int stage_bytes = MT * kATileBytes + b_tile_bytes;
  // latency-bound gather: keep as many K-blocks in flight as shared memory allows (ncu r01: 2 stages -> L2 45 %, tensor 11 %)
  int stages = (200 * 1024) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) stages = 2;
  int lag = stages - 2;
  if (lag < 1) lag = 1;
  if (lag > 6) lag = 6;
  const size_t smem = (size_t)stages * stage_bytes + sizeof(PipeSmem) + 1024 + 64;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(igemm_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024));
    if (e != cudaSuccess) return b2c_cuda_check(e, "conv_wgrad: cudaFuncSetAttribute");
    configured = true;
  }
  // TMA path: both operands 64-channel aligned; min tap offsets = lower corner of the im2col box
  WgradMaps maps;
  memset(&maps, 0, sizeof(maps));
  int lo[3] = {0, 0, 0};
  if (use_tma) {
    for (int k = 0; k < 3; ++k) lo[k] = 127;
    for (int i = 0; i < d.ntaps; ++i) {
      const int32_t tv = d.taps_host[i];
      const int v[3] = {(int)(int8_t)(tv & 0xff), (int)(int8_t)((tv >> 8) & 0xff), (int)(int8_t)((tv >> 16) & 0xff)};
      for (int k = 0; k < 3; ++k)
        if (v[k] < lo[k]) lo[k] = v[k];
    }
    int rc = encode_im2col_map(&maps.g, reinterpret_cast<const uint8_t*>(d.g) + (size_t)d.g_c_off * esz, d.Cg, d.g_row_stride, d.N,
                               d.Tg, d.Hg, d.Wg, lo[0], lo[1], lo[2], d.Qt, d.Qh, d.Qw, d.sg_t, d.sg_h, d.sg_w, cb, pk, esz);
    if (rc) return rc;
    // p_fold: the map covers p_fold channels and every T-plane (the start coordinate selects the plane)
    rc = encode_im2col_map(&maps.p, reinterpret_cast<const uint8_t*>(d.p) + (size_t)d.p_c_off * esz, d.p_fold ? d.p_fold : d.Cp,
                           d.p_row_stride, d.N, d.Tp, d.Hp, d.Wp, d.pp_t, d.pp_h, d.pp_w, d.p_fold ? d.Tp - d.pp_t : d.Qt, d.Qh, d.Qw,
                           d.sp_t, d.sp_h, d.sp_w, cb, pk, esz);
    if (rc) return rc;
  }
  dim3 grid((unsigned)mtc, (unsigned)nt, (unsigned)nsplit);
  igemm_wgrad_kernel<false><<<grid, kThreads, smem, (cudaStream_t)stream>>>(d, maps, use_tma, lo[0], lo[1], lo[2], stages, lag, nsplit, MT);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("conv_wgrad launch");
  return 0;
}

B2C_API int b2c_pack_weights(const float* w, void* packed, const int32_t* wtap, int32_t R, int32_t ntaps, int32_t C,
                             int32_t C_real, int64_t s_r, int64_t s_c, int64_t tap_pitch, int64_t col_off, int32_t r_off,
                             int32_t bn_tile, int32_t nkb, b2c_stream_t stream) {
  B2C_REQUIRE(w && packed && wtap, "pack_weights: null pointer");
  B2C_REQUIRE(R > 0 && ntaps > 0 && C > 0 && C_real > 0 && C_real <= C, "pack_weights: bad dims");
  B2C_REQUIRE(bn_tile > 0 && bn_tile % 16 == 0 && bn_tile <= 256 && nkb > 0, "pack_weights: bad tiling bn_tile=%d nkb=%d", bn_tile, nkb);
  if (tap_pitch <= 0) tap_pitch = C;
  const int tf32 = b2c_precision();       // process-wide: the packed operand type follows the activation precision mode
  const int bk = tf32 ? 32 : 64;
  B2C_REQUIRE(((int64_t)(ntaps - 1) * tap_pitch + col_off + C + bk - 1) / bk <= nkb, "pack_weights: K exceeds nkb");
  const long long total = (long long)R * ntaps * C;
  int blocks = (int)((total + 255) / 256);
  const int cap = b2c_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (tf32)
    pack_weights_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, packed, wtap, R, ntaps, C, C_real, s_r, s_c, tap_pitch,
                                                                       col_off, r_off, bn_tile, nkb);
  else
    pack_weights_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, packed, wtap, R, ntaps, C, C_real, s_r, s_c, tap_pitch,
                                                                        col_off, r_off, bn_tile, nkb);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("pack_weights launch");
  return 0;
}

B2C_API int b2c_pack_weights_batched(const b2c_pack_job* jobs_dev, const int32_t* block_start_dev, int32_t njobs, int32_t nblocks,
                                     b2c_stream_t stream) {
  B2C_REQUIRE(jobs_dev && block_start_dev && njobs > 0 && nblocks > 0, "pack_weights_batched: bad args");
  pack_weights_batched_kernel<<<nblocks, 256, 0, (cudaStream_t)stream>>>(jobs_dev, block_start_dev, njobs);
  b2c_launches_add(1);
  B2C_LAUNCH_CHECK("pack_weights_batched launch");
  return 0;
}
