// kernels.h — host-side launchers of every CUDA kernel in libgan_b200.so.
#pragma once
#include "common.cuh"

struct Launch {
  cudaStream_t s;
  uint64_t* count;   // incremented once per kernel launch (bench: gpu_launches)
};

#define STATS_MAX_CHUNKS 592   // 4 x 148 SMs
#define LOSS_SLOTS 16
#define LOSS_BLOCKS 256
#define HEAD_PART_BLOCKS 4096   // per-block bias-gradient partials of the generator-head backward (4 floats each)

// ---- elem.cu --------------------------------------------------------------------------------
void launch_convert(Launch L, int dt, const float* src, int64_t P, int C, void* dst, int pitch, int coff);
// dst[i] = master[idx[i]] (idx < 0 -> 0): small special weight layouts via a device index table
void launch_gather_pack(Launch L, int dt, const float* master, const int* idx_dev, int n, void* dst);
// fp32 NHWC image -> bf16 im2col rows of the 4x4 stride-2 'same' window: dst[m][t*4 + c], 64 per row (slots >= C zero)
void launch_im2col(Launch L, int dt_rows, const float* src, int B, int H, int W, int C, void* dst_rows);
// Same rows from a bf16 image with pixel pitch `pitch` (generator-head gradient dz): G[m][t*4 + c]
// Transposed-conv head as GEMM + col2im: cols[m][(kh*4+kw)*4 + co] (fp32, 64 per input-grid point m) ->
// out[n, 2i+a, 2j+b, co] = tanh(bias[co] + sum of the 4 contributing taps)   (base_gan.py:201-204)
// cols: fp32 rows, or fp16 rows (dt_cols == DT_F16: half the bytes of the largest intermediate of the step)
void launch_col2im_grad(Launch L, int dt, const void* cols, int B, int Hin, int Win, void* out_rows4);
void launch_col2im_tanh(Launch L, int dt_cols, const void* cols, const float* bias, int B, int Hin, int Win, int C, float* out_f32);
// Discriminator head (ZeroPad + Conv2D 4x4 s1, 512 -> 1, bias; base_gan.py:157-161) in cols form:
//   cols[m'][tap*4] = a[m',:] . w[tap,:]  (GEMM over the 31x31 activation grid), then
//   logits[n,oh,ow] = bias + sum_tap cols[(oh+kh-1, ow+kw-1)][tap*4]
void launch_dhead_gather(Launch L, const float* cols, const float* bias, int B, int Hin, int Win, float* logits);
// Gd[m'=(n,i,j)][tap*4] = dlogit[n, i-kh+1, j-kw+1] (0 outside): the stride-1 unfold of the logit gradient
void launch_dhead_unfold(Launch L, const void* dlogit_bf16, int pitch, int B, int Hin, int Win, void* dst_bf16);
void launch_export(Launch L, int dt, const void* src, int pitch, int coff, int64_t P, int C, float* dst);
// dst view = sum over `nslab` fp32 slabs [nslab][P][C] (deterministic split-K reduction)
void launch_sum_slabs(Launch L, int dt, const float* slabs, int nslab, int64_t P, int C, void* dst, int pitch, int coff);

int stats_chunks(int G, int64_t Pg);
size_t stats_ws_floats(int G, int64_t Pg, int C);
// Per-channel (G==1, BatchNorm) or per-(sample,channel) (G==N, InstanceNorm) moments of z, then
// mean/inv/scale/shift (each [G][C]); BN moving statistics updated when mov_mean != nullptr.
void launch_norm_stats(Launch L, int dt, const void* z, int G, int64_t Pg, int C, float* ws, float eps,
                       const float* gamma, const float* beta, float* mean, float* inv, float* scale,
                       float* shift, float* mov_mean, float* mov_var, float momentum);
void launch_norm_stats_finalize(Launch L, const float* ws, int nparts, int64_t n, int C, float eps, const float* gamma,
                                const float* beta, float* mean, float* inv, float* scale, float* shift, float* mov_mean,
                                float* mov_var, float momentum);
// out = act(dropout(z*scale+shift)); scale == nullptr => identity affine (no-norm layers).
// One-kernel BatchNorm / InstanceNorm layer (G groups of P/G pixels) for small groups; returns false (nothing
// launched) when a group's slab does not fit in shared memory.
bool bn_small_fwd_fits(int G, int64_t P);     // the whole-layer kernel will take this layer (no separate statistics needed)
bool launch_bn_small_fwd(Launch L, int dt, const void* z, int64_t P, int G, int HW, int C, float eps, const float* gamma,
                         const float* beta, float* mean, float* inv, float* scale, float* shift, float* mov_mean,
                         float* mov_var, float momentum, int act, DropKey dk, void* out, int out_pitch, int out_coff);
void set_bn_small(bool on);
void launch_norm_apply(Launch L, int dt, const void* z, int64_t P, int64_t Pg, int G, int HW, int C,
                       const float* mean, const float* scale, const float* shift, int act, DropKey dk, void* out,
                       int out_pitch, int out_coff);
struct GradSrc { const void* p; int pitch, coff; };
// Backward of (norm -> dropout -> activation): dz, plus dgamma/dbeta accumulated into the grad buffer.
void launch_norm_bwd(Launch L, int dtz, int dt, const void* z, GradSrc d1, GradSrc d2, int64_t P, int64_t Pg, int G, int HW,
                     int C, int norm, const float* mean, const float* inv, const float* scale, const float* shift,
                     int act, DropKey dk, float* ws, float* c1, float* c2, float* dgamma, float* dbeta, void* dz,
                     unsigned int* counters, int z_pitch = 0, int z_coff = 0);   // z_pitch > 0: `z` is a strided view (no-norm layers:
                                                                                 // the activation stands in for z, same sign)
// Generator head backward: dz = (d1 + d2 + l1_coef*sign(out-ref)) * (1-out^2); dbias += sum(dz).
// head backward written directly as slot-4 rows of the cols operand (bf16 path; dz is never materialised)
void launch_ghead_bwd_cols(Launch L, int dt, const float* out_f32, const float* ref_f32, GradSrc d1, GradSrc d2, float l1_coef, int B,
                           int H, int W, int C, void* gcols, float* dbias, float* part_ws, unsigned int* counters);
void launch_ghead_bwd(Launch L, int dt, const float* out_f32, const float* ref_f32, GradSrc d1, GradSrc d2,
                      float l1_coef, int64_t P, int C, void* dz, int dz_pitch, float* dbias, float* part_ws, unsigned int* counters);
// BCE-from-logits partial sums into loss slot `slot` and (optionally) dz = coef*(sigmoid(x)-label)/n.
void launch_bce(Launch L, int dt, const float* logits, int64_t n, float label, float coef, void* dz, int dz_pitch,
                float* dbias, float* loss_ws, int slot, float* part_ws, unsigned int* counters);
void launch_l1(Launch L, const float* a, const float* b, int64_t n, float* loss_ws, int slot);
// raw[j] = sum(slot j)/denom[j]; out[i] = sum_j mix[i*nraw+j]*raw[j]
struct LossMix { int nraw, nout; float denom[LOSS_SLOTS]; float mix[8 * LOSS_SLOTS]; };
void launch_loss_finalize(Launch L, const float* loss_ws, LossMix mix, float* out);
void launch_adam(Launch L, float* p, const float* g, float* m, float* v, int64_t n, const long long* t_dev, double lr,
                 double b1, double b2, float eps, float gscale);
void launch_bump(Launch L, long long* t64, uint32_t* c32, uint32_t by);   // device-resident step / dropout-call counters
struct PackOp { int ncls; ClassGeom cls[4]; int Kc, Nc, Kr, Nr; int64_t s_tap, s_k, s_n; };   // Kc/Nc padded, Kr/Nr real
void launch_pack(Launch L, int dt, const float* master, void* dst, const PackOp& op);
// All layers / roles of a net in ONE launch: 32x32 tiles, transposed through shared memory when the
// master layout is contiguous along the packed N index.  `tab` is a device array of PackEntry.
struct PackEntry { PackOp op; const float* master; void* dst; int tiles_k, tiles_n, tile_begin, dt16; };   // dt16: DT_F16 | DT_BF16 of a 16-bit destination
void launch_pack_multi(Launch L, int dt, const PackEntry* tab_dev, int nent, int total_tiles);
void launch_scale(Launch L, float* p, int64_t n, float s);
// Fused Keras-Adam + weight packing: one pass over every convolution kernel of a net (32x32 tiles of
// the master [16][A][B] tensors) updates p/m/v in place and writes the new weight into BOTH packed
// copies (forward and data-gradient roles; one directly, one transposed through shared memory), so the
// weights are read once per step (28+4 B/param instead of 28 + 8).  Non-kernel tensors (gamma/beta/
// bias) are updated by launch_adam_ranges.
struct AdamPackEntry {
  long long w_off; int A, B, conv2d, vec;   // vec: full 64x64 tiles with 16-byte aligned rows (vectorised path)
  void* dstF; void* dstD;
  int KcF, KtotF, KcD, KtotD;
  long long boffF[4], boffD[4];
  int8_t invF[16], invD[16];          // master tap index -> (class << 4) | tap-in-class for each role
  int tiles_a, tiles_b, tile_begin, pad1;
};
struct AdamRange { long long off; int n, pad; };
struct AdamArgs {
  float* p; const float* g; float* m; float* v; const long long* t_dev; double lr, b1, b2; float eps, gscale;
  const uint16_t* g16;      // != nullptr: the (all-reduced) gradient as bf16 bits, same indexing as g (data parallel)
};
// fp32 gradient range -> bf16 communication buffer (round to nearest even), n a multiple of 4
void launch_grad_to_bf16(Launch L, const float* g, uint16_t* g16, int64_t n);
void launch_adam_pack(Launch L, int dt_fwd, int dt_dgrad, const AdamArgs& a, const AdamPackEntry* tab_dev, int nent, int total_tiles);
void launch_adam_ranges(Launch L, const AdamArgs& a, const AdamRange* tab_dev, int nranges);
// zero the gamma / beta / bias gradient ranges (the kernels that produce them accumulate)
void launch_zero_ranges(Launch L, float* g, const AdamRange* tab_dev, int nranges);

// ---- conv_ffma.cu ---------------------------------------------------------------------------
void launch_conv_fwd_ffma(Launch L, int dt, const ConvOp& op);
void launch_conv_wgrad_ffma(Launch L, int dt_in, int dt_dy, const ConvOp& op);

// ---- conv_umma.cu ---------------------------------------------------------------------------
struct UmmaPlan;   // opaque: tensor maps + tiling for one (op, batch) pair
bool umma_fwd_supported(const ConvOp& op);
bool umma_wgrad_supported(const ConvOp& op);
// returns the number of BatchNorm-statistics partials the epilogue wrote into op.stats_ws ([part][2][Nc]); 0 = none
int launch_conv_fwd_umma(Launch L, const ConvOp& op);
void launch_conv_wgrad_umma(Launch L, const ConvOp& op);
void umma_init();   // resolves cuTensorMapEncodeTiled, sets kernel attributes

// ---- conv_first.cu --------------------------------------------------------------------------
// First layers (Conv2D 4x4 s2, 1..4 channels per source image, 64 filters, LeakyReLU, no norm) with the im2col rows built
// in shared memory: src[] fp32 NHWC images, wpack = the layer's im2col-order weight pack [64][nsrc*64], a = LeakyReLU(z)
// into the consumer view (the backward pass takes the activation derivative from the sign of a).
struct FirstLayerOp {
  const float* src[2]; int nsrc, C, B, H, W;
  const void* wpack; void* a; int a_pitch, a_coff; int dt;     // z itself is not stored: LeakyReLU keeps the sign
};
void first_init();
bool first_fwd_supported(const FirstLayerOp& op);
void launch_conv_first_fwd(Launch L, const FirstLayerOp& op);
// ... and their weight gradient dW[(kh,kw), source*C + c, co] = sum over output pixels of row[k] * dz[co]: the same rows
// rebuilt in shared memory as the MN-major operand (GEMM-K = pixels), dz tiles by TMA, one fp32 partial tile per CTA
// reduced in CTA order by k_wgrad_reduce (deterministic); no im2col rows in HBM at all.
struct FirstWgradOp {
  const float* src[2]; int nsrc, C, B, H, W;
  const void* dz; int dz_pitch, dz_coff; int dt;               // (B, H/2, W/2, 64) 16-bit
  float* dW; long long s_tap, s_k, s_n; int accumulate;
  float* ws; size_t ws_bytes;                                  // partial-tile workspace
};
bool first_wgrad_supported(const FirstWgradOp& op);
bool first_wgrad_enabled();
void launch_conv_first_wgrad(Launch L, const FirstWgradOp& op);

// second stage of the deterministic weight-gradient reduction (conv_umma.cu)
struct WgradReduceParams {
  const float* slab; float* dW;
  int splits, bn, mblocks, ntiles, ncls, accumulate;
  int Kc, Kr, Nr, im2col_c, n_slot4_c;
  long long s_tap, s_k, s_n;
  int ntaps[4]; int8_t widx[4][16];
};
void launch_wgrad_reduce(Launch L, const WgradReduceParams& R, long long out_tiles);

// on-device input pipeline: gan_image_xform (include/gan_b200.h) plus the float32 resize scales in/out of
// each stage, divided once per image on the host
struct ImageXformDev {
  int src_h, src_w, col0, cols, pre, mid, crop_y, crop_x, flip;
  float sy1, sx1;     // last resize (to mid, or straight to the output size): source extent / target extent
  float sy0, sx0;     // pre > 0: src_h / pre, cols / pre
};
void launch_preprocess(Launch L, const uint8_t* img, int64_t stride, const ImageXformDev* xf_dev, int B, int C, int S, float* out);
