/* gan_b200.h — C-ABI of the B200-native GAN train-step library (libgan_b200.so).
 *
 * The reference (kingjosephm/GAN) has no FFI of its own: its hot path is Python calling
 * Keras/TensorFlow.  This header is the boundary a maintainer binds with ctypes from the
 * reference's Python classes; each entry point cites the reference interface it replaces
 * (file:line into the reference repository).  See INTEGRATION.md for the binding stub.
 *
 * Conventions
 *   - every function returns 0 on success, a negative gan_status on failure; the message is
 *     available from gan_last_error() (thread-local).  Nothing throws or aborts across the ABI.
 *   - image / logit pointers may be HOST or DEVICE pointers (detected with
 *     cudaPointerGetAttributes); layout is NHWC float32 contiguous, values in [-1,1]
 *     (reference: base_gan.py:56-61 normalize()).
 *   - the caller owns every buffer it passes in; the library owns weights, optimizer state,
 *     saved activations, workspaces, streams and the NCCL communicator.
 *   - a gan_ctx and the objects created from it are driven by one host thread at a time.
 *   - the library refuses to run without a CUDA device: there is no CPU fallback.
 */
#ifndef GAN_B200_H
#define GAN_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define GAN_API __attribute__((visibility("default")))
#else
#define GAN_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gan_ctx gan_ctx;     /* device context: stream, workspaces, RNG counters, communicator */
typedef struct gan_net gan_net;     /* one Generator or Discriminator (weights + saved activations)   */
typedef struct gan_adam gan_adam;   /* Keras-Adam state bound to one gan_net                          */

enum gan_status {
  GAN_OK = 0,
  GAN_ERR_INVALID = -1,     /* bad argument */
  GAN_ERR_CUDA = -2,        /* CUDA runtime / driver error */
  GAN_ERR_NO_DEVICE = -3,   /* no usable sm_100 device */
  GAN_ERR_COMM = -4,        /* NCCL error */
  GAN_ERR_UNSUPPORTED = -5
};

enum gan_precision { GAN_FP32 = 0, GAN_BF16 = 1 };         /* fp32: FFMA path, <=1e-4; bf16: tcgen05 path */
enum gan_norm { GAN_NORM_BATCH = 1, GAN_NORM_INSTANCE = 2 };/* base_gan.py:81-85, utils.py:6-30 */
enum gan_engine { GAN_ENGINE_AUTO = -1, GAN_ENGINE_FFMA = 0, GAN_ENGINE_UMMA = 1 };

GAN_API const char* gan_last_error(void);
GAN_API int gan_version(void);

/* ---- context ------------------------------------------------------------------------------ */
/* Replaces the implicit TF runtime/device placement (base_gan.py:16-19). */
GAN_API int gan_ctx_create(int device, int precision, uint64_t seed, gan_ctx** out);
/* Ownership: a context owns every net and optimizer created from it and destroys them with itself;
 * gan_net_destroy also destroys the optimizers bound to that net.  Destroying a handle twice (or a handle
 * whose owner is gone) returns GAN_ERR_INVALID instead of freeing twice. */
GAN_API int gan_ctx_destroy(gan_ctx* ctx);
GAN_API int gan_ctx_sync(gan_ctx* ctx);
/* Dropout(0.5) is always active in the reference (base_gan.py:118, training=True everywhere);
 * enabled=0 exists for tests only. */
GAN_API int gan_ctx_set_dropout(gan_ctx* ctx, int enabled);
/* Dropout counter state: masks are keyed (seed, call counter, layer, global sample, element). */
GAN_API int gan_ctx_set_rng(gan_ctx* ctx, uint64_t seed, uint32_t call_counter);
GAN_API int gan_ctx_get_call_counter(gan_ctx* ctx, uint32_t* out);
/* Convolution engine override for tests (default AUTO: tcgen05 where the layer fits, FFMA else). */
GAN_API int gan_ctx_set_engine(gan_ctx* ctx, int engine);
/* Capture whole train steps into CUDA graphs keyed on (batch, training) (default 0). */
GAN_API int gan_ctx_set_graphs(gan_ctx* ctx, int enabled);
/* Number of kernels this library launched on ctx since creation (bench: gpu_launches). */
GAN_API int gan_ctx_launch_count(gan_ctx* ctx, uint64_t* out);
/* Per-kernel-family timing with CUDA events on the ctx stream (bench.py roofline).  Families:
 * 0 tcgen05 fwd/dgrad conv, 1 tcgen05 wgrad, 2 FFMA fwd/dgrad conv, 3 FFMA wgrad, 4 norm/activation
 * fwd+bwd, 5 Adam, 6 weight packing, 7 other.  work[] is algorithmic FLOPs (0-3) or bytes (4-6). */
GAN_API int gan_ctx_set_profile(gan_ctx* ctx, int enabled);
GAN_API int gan_ctx_profile_read(gan_ctx* ctx, double ms[8], double work[8], int64_t count[8]);
/* The stream every call of this ctx is enqueued on (cudaStream_t as void*), for event timing. */
GAN_API int gan_ctx_stream(gan_ctx* ctx, void** out);

/* ---- data-parallel communicator (new; the reference is single-device, base_gan.py:18-19) --- */
GAN_API int gan_comm_unique_id(void* out128);                      /* ncclGetUniqueId -> 128 bytes */
GAN_API int gan_ctx_comm_init(gan_ctx* ctx, int rank, int world, const void* unique_id128);
/* global index of the first local sample (keys the dropout masks); default rank*B */
GAN_API int gan_ctx_set_sample_offset(gan_ctx* ctx, int64_t sample0);

/* ---- models --------------------------------------------------------------------------------
 * gan_generator_create     replaces GAN.Generator(norm_type, shape)        base_gan.py:168-225
 * gan_discriminator_create replaces GAN.Discriminator(norm_type, target)   base_gan.py:124-166
 * Weights are created zero; the host initialises them (N(0,0.02) kernels etc.,
 * base_gan.py:74,103,132,200; utils.py:17,23) through gan_net_set_tensor so that oracle and
 * device share bit-identical parameters. */
GAN_API int gan_generator_create(gan_ctx* ctx, int norm_type, int height, int width, int channels, gan_net** out);
GAN_API int gan_discriminator_create(gan_ctx* ctx, int norm_type, int channels, int target, gan_net** out);
GAN_API int gan_net_destroy(gan_net* net);

/* model.trainable_variables order (Keras topological order); indices
 * >= gan_net_num_trainable are the BatchNorm moving_mean / moving_variance pairs. */
GAN_API int gan_net_num_tensors(gan_net* net, int* trainable, int* total);
GAN_API int gan_net_tensor_info(gan_net* net, int idx, char* name, int name_cap, int* ndim, int64_t shape[4], int64_t* numel);
GAN_API int gan_net_get_tensor(gan_net* net, int idx, float* host_dst);
GAN_API int gan_net_set_tensor(gan_net* net, int idx, const float* host_src);
GAN_API int gan_net_get_grad(gan_net* net, int idx, float* host_dst);   /* gradient of the last training step */
/* All trainable tensors back to back (float32), same order. */
GAN_API int gan_net_num_params(gan_net* net, int64_t* out);
GAN_API int gan_net_get_params(gan_net* net, float* host_dst);
GAN_API int gan_net_set_params(gan_net* net, const float* host_src);
GAN_API int gan_net_get_grads(gan_net* net, float* host_dst);
/* Saved tensor of the most recent forward in call slot `slot` ("down3.z", "up2.a", ...), as
 * float32 NHWC; tests only. */
GAN_API int gan_net_debug_tensor(gan_net* net, int slot, const char* name, float* host_dst, int64_t cap, int64_t* numel);

/* model(x, training=True)  — pix2pix.py:200,228; cycle_gan.py:186,220-228.  out: (B,H,W,C). */
GAN_API int gan_generator_forward(gan_net* g, const float* x, int batch, float* out);
/* discriminator([inp, tar], training=True) — pix2pix.py:202-203; tar NULL when target=0
 * (cycle_gan.py:230-234).  logits: (B,H/8-2,W/8-2,1). */
GAN_API int gan_discriminator_forward(gan_net* d, const float* inp, const float* tar, int batch, int height, int width,
                              float* logits);

/* ---- optimizer: GAN.optimizer(lr, beta_1, beta_2) -> tf.keras.optimizers.Adam, base_gan.py:247-252 */
GAN_API int gan_adam_create(gan_net* net, double lr, double beta1, double beta2, double eps, gan_adam** out);
GAN_API int gan_adam_destroy(gan_adam* opt);
GAN_API int gan_adam_get_step(gan_adam* opt, int64_t* t);
GAN_API int gan_adam_set_step(gan_adam* opt, int64_t t);
GAN_API int gan_adam_get_state(gan_adam* opt, int which /*0=m,1=v*/, float* host_dst);   /* flat, param order */
GAN_API int gan_adam_set_state(gan_adam* opt, int which, const float* host_src);
/* Replace lr / beta_1 / beta_2 / epsilon (tf.train.Checkpoint.restore brings the saved optimizer
 * hyper-parameters back with the slots, pix2pix.py:400-411).  Captured step graphs are dropped. */
GAN_API int gan_adam_set_hyper(gan_adam* opt, double lr, double beta1, double beta2, double eps);

/* ---- train steps ---------------------------------------------------------------------------
 * gan_pix2pix_train_step  replaces Pix2Pix.train_step(input_image, target, training)
 *     pix2pix.py:190-218 (+ generator_loss pix2pix.py:167-188, discriminator_loss base_gan.py:233-245)
 *     losses = {gen_total_loss, gen_gan_loss, gen_l1_loss, disc_loss}
 * gan_cyclegan_train_step replaces CycleGAN.train_step(real_x, real_y, training)
 *     cycle_gan.py:206-276 (+ cycle_gan.py:154-177)
 *     losses = {gen_g, gen_f, total_cycle, total_gen_g, total_gen_f, disc_x, disc_y}
 * training=0 runs the same forward (still batch statistics + dropout) without updates
 * (pix2pix.py:208).  losses may be NULL: the step is then only enqueued; read the values later
 * with gan_ctx_last_losses (which synchronises). With a communicator, gradients are summed over
 * ranks and divided by world before Adam; losses are the mean over ranks. */
GAN_API int gan_pix2pix_train_step(gan_net* g, gan_net* d, gan_adam* g_opt, gan_adam* d_opt,
                           const float* input_image, const float* target, int batch,
                           float lambda, int training, float losses[4]);
/* Same step with the two weights of the generator objective spelled out: l1_weight multiplies mean|target-G(x)| in
 * both the reported total and the gradient (lambda for the default generator_loss='l1'); gan_grad_scale multiplies
 * the adversarial term of the GENERATOR gradient only.  The reference's generator_loss='ssim' branch
 * (pix2pix.py:182-186) is l1_weight = 0, gan_grad_scale = batch: its total loss is a per-image vector built from
 * SSIM(input, target), a constant of the step, so tape.gradient differentiates the sum over the batch. */
GAN_API int gan_pix2pix_train_step_ex(gan_net* g, gan_net* d, gan_adam* g_opt, gan_adam* d_opt,
                              const float* input_image, const float* target, int batch,
                              float l1_weight, float gan_grad_scale, int training, float losses[4]);
GAN_API int gan_cyclegan_train_step(gan_net* g, gan_net* f, gan_net* dx, gan_net* dy,
                            gan_adam* g_opt, gan_adam* f_opt, gan_adam* dx_opt, gan_adam* dy_opt,
                            const float* real_x, const float* real_y, int batch,
                            float lambda, int training, float losses[7]);
GAN_API int gan_ctx_last_losses(gan_ctx* ctx, float* out, int n);
/* Input prefetch, the role of `dataset.prefetch(AUTOTUNE)` in the reference (pix2pix.py:163): start the
 * host->device copy of the NEXT step's two image batches on a copy stream while the current step
 * computes.  The next train step recognises the same host pointers and uses the device copies. */
GAN_API int gan_ctx_prefetch(gan_ctx* ctx, const float* x_host, const float* y_host, int64_t bytes_each);

/* ---- on-device input pipeline -------------------------------------------------
 * Replaces the per-image tf.data map of the reference: split_img (pix2pix.py:34-55), resize
 * (base_gan.py:45-53, NEAREST_NEIGHBOR), random_crop + flip_left_right (pix2pix.py:57-90,
 * cycle_gan.py:38-62) and normalize (base_gan.py:56-61), as ONE gather kernel over decoded uint8
 * images.  The random draws (crop offset in [0,30], flip) are made by the caller, as the host RNG
 * of tf.data is not part of the arithmetic.
 *
 * out[n,i,j,c] = src_n[row(i), col0 + col(j'), c] / 127.5 - 1,   j' = flip ? S-1-j : j
 *   mid > 0 (train): index (i+crop_y, j'+crop_x) of the mid x mid nearest-resized image,
 *   mid == 0 (val/test/predict): index (i, j') of the S x S nearest-resized image,
 *   pre > 0 (cycle_gan.py load(resize=True)): the image is first nearest-resized to pre x pre.
 * Nearest index (tf.image.resize v2, half-pixel centres, float32): min(floor((o+0.5)*in/out), in-1). */
typedef struct {
  int src_h, src_w;      /* decoded image size in pixels (HWC uint8, src_w * channels bytes per row) */
  int col0, cols;        /* column window that forms the image (split_img: one half) */
  int pre;               /* >0: first resize to pre x pre */
  int mid;               /* >0: resize to mid x mid (img_size + 30), then crop */
  int crop_y, crop_x;    /* crop offset inside the mid x mid image, 0 <= crop <= mid - out_size */
  int flip;              /* mirror left-right */
} gan_image_xform;
/* images: host or device, image n at images + n*image_stride bytes; xf: host array [batch];
 * out: host or device float32 (batch, out_size, out_size, channels). Runs on the context stream. */
GAN_API int gan_preprocess_images(gan_ctx* ctx, const uint8_t* images, int64_t image_stride, int batch, int channels,
                          int out_size, const gan_image_xform* xf, float* out);
/* Prefetching form: uint8 copy + gather for the NEXT step's two batches run on the copy stream while
 * the current step computes (host->device traffic is the uint8 bytes, a quarter of the float32
 * images).  Returns two device pointers; pass exactly those as input_image/target (real_x/real_y) of
 * the next train step.  images_b may equal images_a (Pix2Pix pairs: one copy, two column windows). */
GAN_API int gan_ctx_prefetch_images(gan_ctx* ctx, const uint8_t* images_a, int64_t stride_a, const gan_image_xform* xf_a,
                            const uint8_t* images_b, int64_t stride_b, const gan_image_xform* xf_b,
                            int batch, int channels, int out_size, const float** a_dev, const float** b_dev);

/* ---- single-operator entry points (parity tests of each kernel family) --------------------
 * kind: 0 = Conv2D 4x4 s2 'same', 1 = ZeroPad(1)+Conv2D 4x4 s1 'valid', 2 = Conv2DTranspose 4x4 s2 'same'
 * role: 0 = forward  (a = x (B,H,W,Cin),     b = kernel,            out = y)
 *       1 = dgrad    (a = dy,                 b = kernel,            out = dx (B,H,W,Cin))
 *       2 = wgrad    (a = x,                  b = dy,                out = dkernel)
 * kernel layout is TF's: Conv2D (4,4,Cin,Cout); Conv2DTranspose (4,4,Cout,Cin).  (H,W) is the
 * layer INPUT size.  All pointers host float32; operands are converted to the ctx precision. */
GAN_API int gan_op_conv(gan_ctx* ctx, int kind, int role, int engine, const float* a, const float* b, float* out,
                int batch, int height, int width, int cin, int cout);

#ifdef __cplusplus
}
#endif
#endif /* GAN_B200_H */
