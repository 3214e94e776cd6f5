/* b200caps C ABI -- the drop-in boundary of the B200-native training step.
 *
 * The reference (AKASH2907/pi-consistency-activity-detection) has no FFI: its seam is the
 * Python module surface.  The Python mirrors of those modules
 * (pi-consistency-activity-detection_b200/{models,utils}) call ONLY the entry points below,
 * through ctypes.  Each entry point names the reference call site it replaces.
 *
 * Conventions
 *   - every function returns int: 0 ok, <0 invalid argument (message via b2c_last_error()),
 *     >0 a cudaError_t.
 *   - all pointers are DEVICE pointers unless the name ends in _host.  The library never
 *     allocates or frees tensor memory; the caller owns every buffer.
 *   - all work is enqueued on the given stream; no internal synchronisation, no host
 *     callbacks (CUDA-graph capturable).
 *   - activations are channels-last (N,T,H,W,C) bf16 "views": base pointer, row stride in
 *     elements (>= C), channel offset -- so concatenations are written in place.
 */
#ifndef B200CAPS_H
#define B200CAPS_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b2c_stream_t; /* cudaStream_t */

const char* b2c_last_error(void);
int b2c_version(void);
/* number of kernels launched by this library since process start (bench.py "gpu_launches") */
long long b2c_launch_count(void);

/* ------------------------------------------------------------------------------------
 * Generalised implicit-GEMM convolution on tcgen05 tensor cores (TMEM accumulators).
 *   out[n, q*so + po, co] = epilogue( sum_{tap, ci} in[n, q*si + d_tap, ci] * w[co][tap][ci] )
 * One descriptor covers: Conv3d/Conv2d fprop (pytorch_i3d.py:115, capsules_ucf101.py:44-45,
 * :490,:497,:501), stride-1 dgrad (flipped taps), ConvTranspose fprop by output-parity class
 * (capsules_ucf101.py:486,495,499,504,509), strided-conv dgrad and ConvTranspose dgrad.
 * taps: int32 per tap, signed bytes (dt | dh<<8 | dw<<16).
 * weights: bf16 tiles produced by b2c_pack_weights with the same bn_tile (pre-swizzled smem image).
 * Requirements: Cin % 8 == 0, Cout % 8 == 0, channel offsets % 8 == 0, row strides % 8 == 0.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  const int32_t* taps; /* [ntaps] */
  const void* w;       /* packed bf16 tiles (b2c_pack_weights) */
  int32_t ntaps;
  int32_t Qt, Qh, Qw;       /* GEMM-M grid of this class */
  int32_t po_t, po_h, po_w; /* output offset of this class */
  int32_t lo_t, lo_h, lo_w; /* per-dimension minimum tap offset (lower corner of the im2col TMA box) */
  int32_t h_block;          /* 0, or > 0: the taps come in consecutive blocks of h_block taps that share (dt, dh), dt the same for
                               all, dh monotonic -- lets a tile skip the blocks whose source rows are padding for all its rows;
                               < 0: blocks of -h_block taps share dt (dt monotonic) and walk the same monotonic dh inside every
                               block: whole blocks are skipped along the T axis and a sub-range inside them along the H axis */
  int32_t pad_;
} b2c_conv_class;

typedef struct {
  const void* in;        /* bf16 */
  void* out;             /* bf16 or fp32 */
  const float* bias;     /* [Cout] or NULL */
  const float* scale_nc; /* [N][Cout] per-sample channel scale (Dropout3d mask) or NULL */
  float* stat_sums;      /* NULL, or fp32 [stat_groups][2][Cout] (zeroed by the caller): the epilogue adds the per-channel sum
                            and sum of squares of the values it stores (BatchNorm batch statistics fused into the GEMM) */
  int64_t in_row_stride, out_row_stride;
  int32_t in_c_off, out_c_off, Cin, Cout;
  int32_t N, Ti, Hi, Wi, To, Ho, Wo;
  int32_t si_t, si_h, si_w, so_t, so_h, so_w;
  int32_t out_fp32;     /* 0: bf16 rows, 1: fp32 rows, 2: fp32 planar out[c][pos] (out_row_stride = plane stride) */
  int32_t relu;         /* apply ReLU */
  int32_t sigmoid_from; /* apply sigmoid to output channels >= this (PrimaryCaps 'a'), <0: none */
  int32_t accumulate;   /* out += result (gradient accumulation) */
  int32_t bn_tile;      /* output-channel tile per CTA: must equal the bn_tile the weights were packed with */
  int32_t tap_pitch;    /* K elements per tap in the packed weights: 0 = Cin; ceil64(Cin) = zero-padded channel tail, which puts
                           layers whose Cin is not a multiple of 64 on the TMA path (TMA zero-fills the channels past Cin) */
  int32_t dtype;        /* 0: bf16 operands (kind::f16), 1: fp32 activations / tf32 operands (kind::tf32; out must be fp32) */
  int32_t stat_groups;  /* statistic groups of stat_sums: the batch N splits into stat_groups equal runs of clips */
  int32_t round_out;    /* dtype 1 only: round the stored fp32 result to tf32 (it is the next GEMM's operand) */
  int32_t nclass;
  int32_t out_fold;     /* 0, or a multiple of 64: output column c is stored in T-plane po_t + c / out_fold at channel c % out_fold
                           (folded stem: the output frames of a pixel are column blocks of one GEMM row; one class, Qt = 1) */
  int64_t w_sample_stride; /* bytes between the packed weight sets of consecutive clips (every class pointer `w` is the set of
                              clip 0); 0 = one set for all clips.  != 0 needs the TMA path and 128-row tiles that do not
                              straddle clips.  Used by the collapsed decoder tail (per-clip composite weights). */
  b2c_conv_class cls[8];
} b2c_conv_desc;

int b2c_conv_fprop(const b2c_conv_desc* desc_host, b2c_stream_t stream);

/* wgrad: dw[pc*s_p + gc*s_g + wtap[tap]] (+)= sum_q g[n, q*sg + d_tap, gc] * p[n, q*sp + pp, pc]
 * (g: tensor that is gathered through the taps, p: tensor read at plain positions).
 * Replaces cuDNN wgrad for every conv / transposed conv of the step (loss.backward(),
 * main_ucf101.py:183).  dw must be zeroed by the caller when nsplit > 1 or accumulate. */
typedef struct {
  const void* g;
  const void* p;
  float* dw;
  const int32_t* taps;      /* [ntaps] device */
  const int32_t* wtap;      /* [ntaps] device: element offset of the tap inside dw */
  const int32_t* taps_host; /* [ntaps] host copy of taps (needed to build the TMA im2col box); NULL = gather path */
  int64_t g_row_stride, p_row_stride, s_p, s_g;
  int32_t g_c_off, p_c_off, Cg, Cp, Cg_real; /* channels >= Cg_real are zero padding */
  int32_t Cp_real;                           /* p channels >= Cp_real are padding (0 = Cp) */
  int32_t N, Tg, Hg, Wg, Tp, Hp, Wp;
  int32_t Qt, Qh, Qw;
  int32_t sg_t, sg_h, sg_w, sp_t, sp_h, sp_w, pp_t, pp_h, pp_w;
  int32_t ntaps;
  int32_t bn_tile; /* 0 = auto */
  int32_t nsplit;  /* 0 = auto */
  int32_t atomic;  /* 1: atomicAdd into dw, 0: plain store (only legal when nsplit==1) */
  int32_t dtype;   /* 0: bf16 operands (the only supported value; tf32 mode passes bf16 hi / lo splits, b2c_split_bf16) */
  int32_t p_fold;  /* 0, or a multiple of 64: p-channel c is read from T-plane pp_t + c / p_fold at channel c % p_fold (folded stem) */
  int32_t pad1_;
  int64_t dw_sample_stride; /* elements between per-clip gradients dw[n]; 0 = one dw summed over all clips.  != 0: the
                               position split is a multiple of N so that no CTA straddles clips */
  /* Fused layers (several convolutions over one input run as one GEMM, e.g. PrimaryCaps pose | activation, the Inception
   * sibling 1x1x1 convolutions): p channels [seg_begin[i], seg_begin[i+1]) belong to the weight tensor seg_dw[i] (same
   * s_p / s_g layout, channel index relative to seg_begin[i]).  nseg = 0: everything goes to dw.  One launch then reads
   * the gathered operand once instead of once per member. */
  float* seg_dw[4];
  int32_t seg_begin[4];
  int32_t nseg;
  int32_t pad2_;
} b2c_wgrad_desc;

int b2c_conv_wgrad(const b2c_wgrad_desc* desc_host, b2c_stream_t stream);

/* fp32 master weights -> bf16 GEMM operand in the fprop kernel's shared-memory image:
 *   tiles [n-tile][k-block][bn_tile rows][64 K-elements], 128-byte swizzle pre-applied, so each pipeline stage
 *   fetches its weight tile with ONE bulk copy.  Element (r, t, c), r<R, t<ntaps, c<C:
 *     value = (c < C_real) ? bf16(w[r*s_r + c*s_c + wtap[t]]) : 0 ; GEMM row = r + r_off ; k = t*tap_pitch + col_off + c.
 *   (tap_pitch = C, col_off = r_off = 0 for a stand-alone layer; other values place sibling layers side by side in
 *   one fused GEMM operand.)  `packed` must be zero-initialised, size ceil(rows/bn_tile)*nkb*bn_tile*64 elements. */
int b2c_pack_weights(const float* w, void* packed, const int32_t* wtap, int32_t R, int32_t ntaps, int32_t C,
                     int32_t C_real, int64_t s_r, int64_t s_c, int64_t tap_pitch, int64_t col_off, int32_t r_off,
                     int32_t bn_tile, int32_t nkb, b2c_stream_t stream);

/* The same packing for MANY layers in one launch (the fused step re-packs all 137 operands after every optimiser
 * step).  jobs_dev: device array; block_start_dev: device int32[njobs+1] prefix of CUDA blocks per job. */
typedef struct {
  const void* w;
  void* packed;
  const int32_t* wtap;
  int64_t s_r, s_c, tap_pitch, col_off;
  int32_t R, ntaps, C, C_real, r_off, bn_tile, nkb;
  int32_t dtype; /* 0: bf16 operand, 1: tf32 (fp32 storage, 32 K-elements per 128-byte row) */
} b2c_pack_job;
int b2c_pack_weights_batched(const b2c_pack_job* jobs_dev, const int32_t* block_start_dev, int32_t njobs, int32_t nblocks,
                             b2c_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Bandwidth kernels (channels-last bf16 views: ptr, rows, C, row_stride, c_off)
 * ---------------------------------------------------------------------------------- */
/* (N,C,T,H,W) fp32 -> (N,T,H,W,Cpad) bf16, zero padded channels.  main_ucf101.py:52-55 cast + layout. */
int b2c_ncdhw_to_ndhwc(const float* in, void* out, int32_t N, int32_t C, int64_t THW, int32_t Cpad, b2c_stream_t s);
/* (rows,C) bf16 view -> (N,C,THW) fp32 and back (module-boundary conversions) */
int b2c_ndhwc_to_ncdhw_f32(const void* in, int64_t in_row_stride, int32_t in_c_off, float* out, int32_t N, int32_t C,
                           int64_t THW, b2c_stream_t s);

/* Input pipeline on the device (ucf_dataloader.py:162-185, main_ucf101.py:52-55): uint8 clips (P,C,T,H,W) as decoded ->
 * channels-last activations (2P, T, H, W, 8) of both forward passes: out[n] = u8 / 255, out[P + n] = the clip mirrored in W
 * (the reference's `aug_data`).  rows = T * H.  The flipped clips and the fp32 copies never cross PCIe. */
int b2c_u8_clip_to_cl(const uint8_t* in, void* out, int32_t P, int32_t C, int64_t rows, int32_t W, b2c_stream_t s);
/* out[i] = in[i] * scale (uint8 segmentation masks -> fp32 targets) */
int b2c_u8_to_f32(const uint8_t* in, float* out, int64_t n, float scale, b2c_stream_t s);

/* Explicit im2col for the few-channel stem (Conv3d_1a_7x7, pytorch_i3d.py:224): x channels-last bf16 (N,T,H,W,Cs) with C
 * real channels -> out (N*To*Ho*Wo, Kpad) bf16, column = tap*C + c, zero padded to Kpad (multiple of 64). */
int b2c_im2col_small(const void* x, void* out, int32_t N, int32_t Cs, int32_t C, int32_t T, int32_t H, int32_t W, int32_t To,
                     int32_t Ho, int32_t Wo, int32_t kt, int32_t kh, int32_t kw, int32_t st, int32_t sh, int32_t sw, int32_t pt,
                     int32_t ph, int32_t pw, int32_t Kpad, b2c_stream_t s);

/* Folded stem: the same layer without the im2col matrix.  The time axis is folded into the channel axis
 * (xs[n][h][w][fp*4 + c] = x[n][fp - pt][h][w][c], fp = padded frame < Tp <= 16, 4 channel slots per frame), which turns the
 * few-channel 3-D convolution into a 2-D convolution over Tp*4 channels that runs on b2c_conv_fprop / b2c_conv_wgrad's TMA
 * path with one output class (own weight set W2[t]) per output frame t.
 *   _fold_input  : x channels-last (N,T,H,W,8) activation precision -> xs (N,H,W,Tp*4)
 *   _fold_weights: w (Cout,Cin,kt,kh,kw) fp32 -> w2 fp32 [To][Cout][kh*kw][Kf = Tp*4], the kernel shifted to frame t's window
 *   _unfold_wgrad: dw += the adjoint of _fold_weights applied to dw2 */
int b2c_stem_fold_input(const void* x, void* xs, int32_t N, int32_t T, int32_t H, int32_t W, int32_t Cs, int32_t pt, int32_t Tp,
                        b2c_stream_t s);
/* the fused step's input path: clips (P,C,T,H,W) fp32 (in_u8 = 0) or uint8 (in_u8 = 1: / 255) straight into the folded
 * layout xs (N',H,W,Tp*4); mirror = 1 also writes the W-mirrored clip (the reference's aug_data) at batch offset P */
int b2c_clips_to_folded(const void* in, int32_t in_u8, void* xs, int32_t P, int32_t C, int32_t T, int32_t H, int32_t W, int32_t pt,
                        int32_t Tp, int32_t mirror, b2c_stream_t s);
int b2c_stem_fold_weights(const float* w, float* w2, int32_t Cout, int32_t Cin, int32_t kt, int32_t khw, int32_t st, int32_t To,
                          int32_t Kf, b2c_stream_t s);
int b2c_stem_unfold_wgrad(const float* dw2, float* dw, int32_t Cout, int32_t Cin, int32_t kt, int32_t khw, int32_t st, int32_t To,
                          int32_t Kf, b2c_stream_t s);

/* BatchNorm3d training statistics (pytorch_i3d.py:80,117).  groups: rows are split evenly into
 * `groups` contiguous segments with independent statistics (two forward passes batched).
 * _sums: ws fp32 [groups][2][C] (zeroed by the caller) += per-channel sum / sum of squares.
 * _finalize: for channels [c_off, c_off+C) of a ws with ws_C channels: mean[g][C], rstd[g][C]; running
 * stats updated with momentum (unbiased variance; groups applied in order) when non-NULL. */
int b2c_bn_sums(const void* x, int64_t rows, int32_t C, int64_t row_stride, int32_t c_off, int32_t groups, float* ws,
                b2c_stream_t s);
int b2c_bn_finalize(const float* ws, int32_t ws_C, int32_t c_off, int32_t C, int32_t groups, int64_t rows_per_group,
                    float* mean, float* rstd, float* running_mean, float* running_var, float momentum, float eps,
                    b2c_stream_t s);
/* _sums and _finalize in one launch: the block that publishes its partial sums last computes mean / rstd / running stats */
int b2c_bn_sums_finalize(const void* x, int64_t rows, int32_t C, int64_t row_stride, int32_t c_off, int32_t groups, float* ws,
                         float* mean, float* rstd, float* running_mean, float* running_var, float momentum, float eps,
                         b2c_stream_t s);
/* y = relu((x-mean)*rstd*gamma+beta) written into a concat slot */
int b2c_bn_relu_apply(const void* x, int64_t rows, int32_t C, int64_t x_row_stride, int32_t x_c_off, int32_t groups,
                      const float* mean, const float* rstd, const float* gamma, const float* beta, void* y,
                      int64_t y_row_stride, int32_t y_c_off, int32_t relu, b2c_stream_t s);
/* backward: pass 1 reduces sum(dyr) and sum(dyr*xhat) (dyr = dy * (y>0)); ws fp32 [groups][2][C] zeroed.
 * y == NULL with relu: both passes recompute the mask y > 0 from x with the forward's own arithmetic (gamma, beta needed)
 * instead of reading y -- a third / a quarter less traffic. */
int b2c_bn_relu_bwd_reduce(const void* dy, int64_t dy_row_stride, int32_t dy_c_off, const void* y, int64_t y_row_stride,
                           int32_t y_c_off, const void* x, int64_t x_row_stride, int32_t x_c_off, int64_t rows,
                           int32_t C, int32_t groups, const float* mean, const float* rstd, const float* gamma,
                           const float* beta, float* ws, int32_t relu, b2c_stream_t s);
/* pass 2: dx = gamma*rstd*(dyr - s1/M - xhat*s2/M); also dgamma += sum_g s2, dbeta += sum_g s1 */
int b2c_bn_relu_bwd_apply(const void* dy, int64_t dy_row_stride, int32_t dy_c_off, const void* y, int64_t y_row_stride,
                          int32_t y_c_off, const void* x, int64_t x_row_stride, int32_t x_c_off, int64_t rows,
                          int32_t C, int32_t groups, const float* mean, const float* rstd, const float* gamma,
                          const float* beta, const float* ws, void* dx, int64_t dx_row_stride, int32_t dx_c_off,
                          float* dgamma, float* dbeta, int32_t relu, b2c_stream_t s);

/* MaxPool3dSamePadding (pytorch_i3d.py:13-45): zero 'same' padding then max; idx = uint8 argmax tap
 * (255 = a padding zero won).  */
int b2c_maxpool_fwd(const void* x, int64_t x_row_stride, int32_t x_c_off, void* y, int64_t y_row_stride, int32_t y_c_off,
                    uint8_t* idx, int32_t N, int32_t C, int32_t Ti, int32_t Hi, int32_t Wi, int32_t To, int32_t Ho,
                    int32_t Wo, int32_t kt, int32_t kh, int32_t kw, int32_t st, int32_t sh, int32_t sw, int32_t pt,
                    int32_t ph, int32_t pw, b2c_stream_t s);
int b2c_maxpool_bwd(const void* dy, int64_t dy_row_stride, int32_t dy_c_off, const uint8_t* idx, void* dx,
                    int64_t dx_row_stride, int32_t dx_c_off, int32_t N, int32_t C, int32_t Ti, int32_t Hi, int32_t Wi,
                    int32_t To, int32_t Ho, int32_t Wo, int32_t kt, int32_t kh, int32_t kw, int32_t st, int32_t sh,
                    int32_t sw, int32_t pt, int32_t ph, int32_t pw, int32_t accumulate, b2c_stream_t s);

/* y[n,pos,c] = x[n,pos,c] * scale[n,c]   (Dropout3d, capsules_ucf101.py:428) ; rows_per_n = T*H*W */
int b2c_channel_scale(const void* x, int64_t x_row_stride, int32_t x_c_off, const float* scale_nc, void* y,
                      int64_t y_row_stride, int32_t y_c_off, int32_t N, int64_t rows_per_n, int32_t C, b2c_stream_t s);
/* dz = dy * (y > 0) (ReLU backward, optional) * scale[n,c] (optional); dbias[c] += sum dz */
int b2c_act_bwd(const void* dy, int64_t dy_row_stride, int32_t dy_c_off, const void* y, int64_t y_row_stride,
                int32_t y_c_off, const float* scale_nc, void* dz, int64_t dz_row_stride, int32_t dz_c_off, float* dbias,
                int32_t N, int64_t rows_per_n, int32_t C, int32_t relu, b2c_stream_t s);
/* out = a + b (bf16 views, gradient fan-in) */
int b2c_add(const void* a, int64_t a_row_stride, int32_t a_c_off, const void* b, int64_t b_row_stride, int32_t b_c_off,
            void* out, int64_t o_row_stride, int32_t o_c_off, int64_t rows, int32_t C, b2c_stream_t s);

/* 'smooth' ConvTranspose3d(128->1,k3,p1) (capsules_ucf101.py:509) second half: 27-tap stencil over the
 * per-tap projections P, PLANAR fp32 [32][rows] -> logits fp32 (N,T,H,W); and its adjoint dP (bf16 rows). */
int b2c_stencil27_fwd(const float* P, float* out, const float* bias, int32_t N, int32_t T, int32_t H, int32_t W,
                      b2c_stream_t s);
/* dP bf16 (rows,cpad), taps 27..cpad-1 zero; dbias[0] += sum(dout) when non-NULL */
int b2c_stencil27_bwd(const float* dout, void* dP, float* dbias, int32_t N, int32_t T, int32_t H, int32_t W, int32_t cpad,
                      b2c_stream_t s);

/* ------------------------------------------------------------------------------------
 * Capsule head
 * ---------------------------------------------------------------------------------- */
/* EM routing, ConvCaps.forward with K=(1,1) (capsules_ucf101.py:290-309; m_step :108-156,
 * e_step :158-182, 3 iterations).  caps: fp32 (b, 32*16 + 32) [poses | activations] as produced by the
 * fused PrimaryCaps GEMM; W fp32 (32, C, 4, 4); out fp32 (b, C*16 + C) [mu | a_out]. C <= 32. */
int b2c_em_routing_fwd(const float* caps, const float* W, const float* beta_u, const float* beta_a, float* out,
                       int64_t b, int32_t C, b2c_stream_t s);
/* backward through the 3 unrolled iterations (recomputes the forward per location).
 * dW/dbeta_u/dbeta_a are accumulated (atomicAdd) -- zero them first. */
int b2c_em_routing_bwd(const float* caps, const float* W, const float* beta_u, const float* beta_a, const float* dout,
                       float* dcaps, float* dW, float* dbeta_u, float* dbeta_a, int64_t b, int32_t C, b2c_stream_t s);
/* Training pair: the forward also saves the per-iteration routing state (assignments, normalisers, means, variances:
 * b2c_em_routing_state_floats() floats per location) and the backward reads it instead of recomputing the three EM
 * iterations.  Same results as the pair above.  The backward CONSUMES the state: its first kernel overwrites the saved
 * assignments with its per-pair coefficients and fills the rows its second kernel reads, so one forward serves one
 * backward. */
int64_t b2c_em_routing_state_floats(void);
int b2c_em_routing_fwd_train(const float* caps, const float* W, const float* beta_u, const float* beta_a, float* out, float* state,
                             int64_t b, int32_t C, b2c_stream_t s);
int b2c_em_routing_bwd_state(const float* caps, const float* W, const float* beta_u, const float* beta_a, const float* dout,
                             float* state, float* dcaps, float* dW, float* dbeta_u, float* dbeta_a, int64_t b, int32_t C,
                             b2c_stream_t s);
/* PrimaryCaps backward prologue (capsules_ucf101.py:43-49 adjoint): g, out fp32 (rows,544); dz bf16 (rows,dz_pitch>=544) =
 * g * (col >= 512 ? a(1-a) : 1); dbias[544] += column sums (first 512: pose bias, last 32: a bias). */
int b2c_primarycaps_bwd_prep(const float* g, const float* out, void* dz, float* dbias, int64_t rows, int32_t dz_pitch,
                             b2c_stream_t s);
/* the same, also writing dz with the clips innermost: dz_rows[(h * Wq + w) * N + n] = dz[(n * Hq + h) * Wq + w] (operand of the
 * rows-major PrimaryCaps dgrad, whose tiles then hold a few columns of one image row of all clips and skip the tap rows /
 * tap columns that are padding for them) */
int b2c_primarycaps_bwd_prep2(const float* g, const float* out, void* dz, void* dz_rows, float* dbias, int32_t N, int32_t Hq,
                              int32_t Wq, int32_t dz_pitch, b2c_stream_t s);
/* Row permutations between the clip-major activation layout (a channel window [c_off, c_off + C) of a (N, H, W, Ctot) tensor
 * with row stride Ctot) and the compact (H, W, N, C) order the rows-major GEMMs work in (C % 8 == 0). */
int b2c_rows_to_clips(const void* in, void* out, int64_t out_row_stride, int32_t out_c_off, int32_t N, int32_t H, int32_t W, int32_t C,
                      b2c_stream_t s);
int b2c_clips_to_rows(const void* in, int64_t in_row_stride, int32_t in_c_off, void* out, int32_t N, int32_t H, int32_t W, int32_t C,
                      b2c_stream_t s);
/* PrimaryCaps epilogue for the K-split forward: the GEMM's K dimension (the 81 taps) runs as `nslice` scheduling classes
 * that write their partial sums to frames 0..nslice-1 of part fp32 (N, nslice, L, 544); out[n][l][c] = sum_s part[n][s][l][c]
 * + bias[c], sigmoid on the 32 activation columns (capsules_ucf101.py:43-49).  Fixed summation order: bit-reproducible. */
int b2c_primarycaps_finish(const float* part, int32_t nslice, const float* bias, float* out, int32_t N, int32_t L, b2c_stream_t s);
/* class activation = mean over the 400 locations (capsules_ucf101.py:450-451); feat is a view of out */
int b2c_class_mean_fwd(const float* rout, float* act, int32_t N, int32_t L, int32_t C, b2c_stream_t s);
/* pose masking (capsules_ucf101.py:455-483): x[n,l,j*16+h] = mu[n,l,j,h] * mask[n,j]  -> bf16 (N,L,C*16) */
int b2c_pose_mask_fwd(const float* rout, const float* mask, void* x, int32_t N, int32_t L, int32_t C, b2c_stream_t s);
/* drout[n,l,:] = [dx * mask | dact[n,j]/L + dfeat[n,l,j]] */
int b2c_caps_head_bwd(const void* dx, const float* mask, const float* dact, const float* dfeat, float* drout, int32_t N,
                      int32_t L, int32_t C, b2c_stream_t s);

/* ------------------------------------------------------------------------------------
 * Losses (utils/losses.py, utils/helpers.py, main_ucf101.py:89-148)
 * ---------------------------------------------------------------------------------- */
/* BCEWithLogits(mean) + Dice over the labeled subset (main_ucf101.py:89-92, losses.py:44-57).
 * logits fp32 (P, V); targets fp32; lab_idx int32 [n_lab] rows of logits; sums: fp64[4] scratch
 * (bce_sum, sum p*t, sum p, sum t; zeroed by the call).  loss out: fp32[2] = (bce, dice). */
int b2c_seg_loss_fwd(const float* logits, const float* targets, const int32_t* lab_idx, int32_t n_lab, int64_t V,
                     double* sums, float* loss, b2c_stream_t s);
/* dlogits[lab rows] += w_bce * dBCE + w_dice * dDice  (other rows untouched) */
int b2c_seg_loss_bwd(const float* logits, const float* targets, const int32_t* lab_idx, int32_t n_lab, int64_t V,
                     const double* sums, float w_bce, float w_dice, float* dlogits, b2c_stream_t s);
/* SpreadLoss (losses.py:14-37) on rows lab_idx of act (P,C); loss fp32[2] (loss, absloss);
 * dact rows (+)= w * dloss */
int b2c_spread_loss(const float* act, const float* target, const int32_t* lab_idx, int32_t n_lab, int32_t C, float m_min,
                    float* loss, float w, float* dact, b2c_stream_t s);
/* Temporal-variance attentive mask, measure_pixelwise_var_v2 (helpers.py:8-67): 14-frame cycle
 * pred[0..7] ++ flip_pred[1..6], cyclic windowed population variance (frames_cnt 3|5), fold to 8 frames,
 * per-clip min-max normalisation.  pred/flip_pred/m: fp32 (P,8,H,W); mm: fp32 (P,2) scratch.
 * pred_tflip / fp_tflip / fp_wmirror apply the caller's torch.flip's (main_ucf101.py:100,114-115) on the
 * fly so the fused step can pass the raw outputs of both forward passes. */
int b2c_bv_mask(const float* pred, const float* flip_pred, float* m, float* mm, int32_t P, int32_t H, int32_t W,
                int32_t frames_cnt, int32_t use_sigmoid, int32_t pred_tflip, int32_t fp_tflip, int32_t fp_wmirror,
                b2c_stream_t s);
/* gradient-smoothness mask, measure_pixelwise_gradient (helpers.py:70-95): sigmoid, optional clamps,
 * np.gradient twice along time, per-clip min-max normalisation.  m: fp32 (P,8,H,W) (no channel dim). */
int b2c_gv_mask(const float* out, float* m, float* mm, int32_t P, int32_t H, int32_t W, float lower, float upper,
                int32_t use_lower, int32_t use_upper, b2c_stream_t s);
/* consistency loss, weighted_mse_loss (losses.py:74-76) as composed in main_ucf101.py:100-148.
 *   d = (mirror ? flipW(flp) : flp) - out ; l2 = mean(d^2) ; lv = mean(w1 d^2) + mean(w2' d^2)
 *   (w2' = time-flipped w2 when w2_tflip) ; lg = mean_thw( mean_j wg_j * mean_i d_i^2 ) -- the
 *   (B,B,8,H,W) broadcast of a (B,8,H,W) weight against (B,1,8,H,W) (main_ucf101.py:130-132).
 * w1/w2/wg may be NULL.  acc: fp64[4] scratch; loss fp32[4] = (cons, l2, lv, lg); mode bit0 bv, bit1 gv.
 * _grad: dout -= g, dflp(+mirror) += g with g = (a_l2 + a_lv (w1+w2')) 2d/(P THW) + a_lg 2 B d/(THW P^2). */
int b2c_cons_reduce(const float* out, const float* flp, const float* w1, const float* w2, const float* wg, double* acc,
                    int32_t P, int32_t H, int32_t W, int32_t mirror, int32_t w2_tflip, b2c_stream_t s);
/* dev_scalars (may be NULL): device float[3] that overrides (wt_ramp, bv_wt, gv_wt) resp. (a_l2, a_lv, a_lg), so the
 * per-epoch schedule values (main_ucf101.py:181,419) can change between replays of one captured CUDA graph. */
int b2c_cons_finish(const double* acc, float* loss, int32_t P, int32_t H, int32_t W, int32_t mode, float wt_ramp,
                    float bv_wt, float gv_wt, const float* dev_scalars, b2c_stream_t s);
int b2c_cons_grad(const float* out, const float* flp, const float* w1, const float* w2, const float* wg, float* dout,
                  float* dflp, int32_t P, int32_t H, int32_t W, int32_t mirror, int32_t w2_tflip, float a_l2, float a_lv,
                  float a_lg, const float* dev_scalars, b2c_stream_t s);

/* ------------------------------------------------------------------------------------
 * Optimiser: Adam(lr, betas, eps=1e-6, wd=0) over one flat fp32 buffer (main_ucf101.py:416,184)
 * ---------------------------------------------------------------------------------- */
/* step_dev: device int32 step counter, incremented by the call (graph replayable bias correction);
 * lr_dev (may be NULL): device float that overrides lr (ReduceLROnPlateau, main_ucf101.py:417,456, between graph replays) */
int b2c_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                  int32_t* step_dev, float grad_scale, const float* lr_dev, b2c_stream_t s);

/* tests: 1 = BatchNorm reductions use one block per statistic group (bit-reproducible activations, slow) */
int b2c_set_deterministic(int32_t on);
/* tests: 1 = the generic max-pool kernels even where a (window, stride)-specialised row kernel exists */
int b2c_set_pool_generic(int32_t on);

/* ------------------------------------------------------------------------------------
 * Collapsed decoder tail: upsample4 (ConvTranspose3d 128->128 k3 s2 p1 op1) -> Dropout3d -> smooth
 * (ConvTranspose3d 128->1 k3 p1), capsules_ucf101.py:504-509, has no non-linearity in between, so the chain equals ONE
 * per-clip stride-2 transposed convolution 128 -> 1 with a 5x5x5 composite kernel
 *     Weff[n][ci][e] = sum_c drop[n][c] sum_{k+m=e} W4[ci][c][k] Ws[c][m]        (per dimension: k, m in 0..2, e in 0..4)
 * except on the index-0 planes (the reference crops upsample4's position -1 before `smooth` reads it), which use a 6th
 * per-dimension column e=2' = e=2 without the (k=0, m=2) term.  6^3 = 216 columns, stored 224 wide.
 *   logits[n][o] = bs + biasfield[n][class(o)] + sum_{i, e: o = 2 i - 2 + e} Y[n][i][col(e, i)],   Y[n][i][:] = x[n][i][:] . Weff[n]
 * The 128-channel (N,8,224,224) tensor (3.3 GB in bf16 at 16+16 clips) is never formed; the per-clip GEMMs
 * Y = x Weff (fprop), dx = dY Weff^T (dgrad) and dWeff[n] = x[n]^T dY[n] (wgrad) run on b2c_conv_fprop / b2c_conv_wgrad
 * with w_sample_stride / dw_sample_stride.  All tensors below are fp32 unless noted.
 * ---------------------------------------------------------------------------------- */
/* composite weights in both packed operand images (fprop: rows = 224 columns, K = 128; dgrad: rows = 128, K = 224 padded to
 * the mode's tap pitch), activation-precision element type, one set per clip `*_stride` ELEMENTS apart (buffers zeroed once by
 * the caller: padding is never written); biasfield[n][27] (border class (t,h,w), 0 = first plane, 1 = interior, 2 = last). */
int b2c_tail_weff(const float* w4, const float* b4, const float* ws, const float* drop_nc, void* packed_fprop,
                  int64_t fprop_stride, void* packed_dgrad, int64_t dgrad_stride, int32_t dgrad_nkb, float* biasfield, int32_t N,
                  b2c_stream_t s);
/* logits (N,2It,2Ih,2Iw) from the planar GEMM output Y[224][N*It*Ih*Iw] */
int b2c_tail_gather_fwd(const float* y_planar, const float* biasfield, const float* bs, float* logits, int32_t N, int32_t It,
                        int32_t Ih, int32_t Iw, b2c_stream_t s);
/* dY rows (N*It*Ih*Iw, 224) in the activation precision from dlogits; also sums[n][27] += per-border-class sums of dlogits
 * (sums zeroed by the caller) */
int b2c_tail_gather_bwd(const float* dlogits, void* dy, float* class_sums, int32_t N, int32_t It, int32_t Ih, int32_t Iw,
                        b2c_stream_t s);
/* chain rule back to the reference's parameters: dweff[n][ci][224] (per-clip wgrad) and class_sums ->
 * dw4 += , db4 += , dws += , dbs += (accumulated: the buffers are the parameters' .grad) */
int b2c_tail_chain_bwd(const float* dweff, const float* class_sums, const float* w4, const float* b4, const float* ws,
                       const float* drop_nc, float* dw4, float* db4, float* dws, float* dbs, int32_t N, b2c_stream_t s);

/* Evaluation consumer (evaluate_ucf101.py:127,151-168): per frame (HW pixels) pred = sigmoid(logit) >= 0.5 against the
 * ground-truth mask -> out[f] = {intersection, union, ground-truth pixels} (int32). */
int b2c_frame_iou_counts(const float* logits, const float* gt, int32_t* out, int64_t frames, int32_t HW, b2c_stream_t s);

/* tf32 mode only: fp32 view (rows, C) -> compact bf16 tensors hi = bf16(x), lo = bf16(x - hi).  The mode's weight
 * gradients are three bf16 GEMMs on these (hi*hi + hi*lo + lo*hi, fp32 accumulate: >= tf32 accuracy). */
int b2c_split_bf16(const float* x, int64_t x_row_stride, int32_t x_c_off, void* hi, void* lo, int64_t rows, int32_t C,
                   b2c_stream_t s);

/* Precision mode of the activation tensors, process-wide.  0 (default): bf16 activations, bf16 GEMM operands
 * (tcgen05.mma kind::f16).  1: fp32 activations and fp32 packed weights read by the tensor core as tf32
 * (tcgen05.mma kind::tf32, fp32 accumulate; weight gradients as 3 x bf16 split GEMMs, see b2c_split_bf16) -- the reference computes in fp32 (main_ucf101.py:52-55, cuDNN convs with
 * allow_tf32 default); this mode is the one its 1e-3 parity bar is asserted in.  Every `void*` activation view of this
 * header is bf16 in mode 0 and fp32 in mode 1; values that feed a later GEMM are rounded to tf32 when stored. */
int b2c_set_precision(int32_t mode);
int b2c_get_precision(void);

/* generic helpers */
int b2c_fill_f32(float* p, int64_t n, float v, b2c_stream_t s);

#ifdef __cplusplus
}
#endif
#endif
