// common.cuh — shared types for the gan_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string>
#include <stdexcept>

typedef __nv_bfloat16 bf16;
typedef __half f16;

// Storage types.  The 16-bit (tcgen05) mode stores activations, gradients and both weight packs as fp16: 10 mantissa
// bits against bf16's 7 (the measured floor of ANY bf16-operand generator forward is L2 8e-3 / max-rel 1.2e-2 against
// the float64 oracle, above the stated 1e-2 tolerance; fp16 operands give 1e-3, scripts/precision_floor.py), at the
// same tcgen05 kind::f16 rate.  Every forward value of these normalised networks is O(1..100); gradients (1e-9..1e-3)
// are kept in range by a static power-of-two LOSS SCALE applied at the loss heads and removed inside Adam
// (engine.cu grad_scale), and every conversion to fp16 saturates instead of producing inf.  tcgen05 kind::f16 needs
// A and B in the SAME format (a f16 x bf16 MMA faults as an illegal instruction on B200 — probed), so activations and
// gradients cannot use different 16-bit formats.  GAN_B200_ACT=bf16 selects all-bf16 storage (round 1) for A/B runs.
enum { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };
enum { K_CONV_S2 = 0, K_CONV_S1P = 1, K_CONVT_S2 = 2 };
enum { NORM_NONE = 0, NORM_BATCH = 1, NORM_INSTANCE = 2 };
enum { ACT_NONE = 0, ACT_LEAKY = 1, ACT_RELU = 2, ACT_TANH = 3 };
enum { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_TANH = 2 };

#define LEAKY_SLOPE 0.3f

// A strided NHWC view: element (n,h,w,c) lives at p[((n*H+h)*W+w)*pitch + coff + c].
struct View {
  void* p;
  int N, H, W, C;
  int pitch;   // elements per pixel in the underlying buffer (>= coff + C)
  int coff;    // channel offset inside the pixel
  int64_t pixels() const { return (int64_t)N * H * W; }
};

static inline View make_view(void* p, int N, int H, int W, int C, int pitch = -1, int coff = 0) {
  View v; v.p = p; v.N = N; v.H = H; v.W = W; v.C = C; v.pitch = pitch < 0 ? C : pitch; v.coff = coff; return v;
}

// One output class of a "tap GEMM" (see DESIGN.md §3): out pixel (mh*so+oa, mw*so+ob) of the
// M-space point (n,mh,mw) accumulates, for every tap t, in[n, mh*si+dh[t], mw*si+dw[t], :] times
// the packed weight slab [Nc][t*Kc .. t*Kc+Kc).
struct ClassGeom {
  int oa, ob, ntaps;
  int8_t dh[16], dw[16], widx[16];   // widx = kh*4+kw of the master 4x4 kernel
  int64_t b_off;                     // element offset of this class's packed weights
};

struct ConvOp {
  const void* in; int in_pitch, in_coff, Hin, Win;
  void* out;      int out_pitch, out_coff, Hout, Wout;
  int N, Hm, Wm;                 // M-space extents
  int si, so;                    // input / output pixel stride of the M-space
  int Kc, Nc;                    // channels per tap (GEMM K = ntaps*Kc), output channels (GEMM N); may be
                                 // zero-padded to 16 in bf16 mode so that small-channel layers fit a tcgen05 tile
  int Kr, Nr;                    // real (unpadded) channel counts: wgrad scatter / bias / fp32 output masks
  int ncls;
  ClassGeom cls[4];
  const void* B;                 // packed weights, dtype = activation dtype
  const float* bias; int epi;    // forward epilogue
  float* out_f32;                // optional fp32 copy of the epilogue output, compact [pix][Nc]
  // wgrad: dW[widx*s_tap + kc*s_k + nc*s_n] += sum_m in(m,t,kc) * out(m,nc)
  float* dW; int64_t s_tap, s_k, s_n;
  // optional fp32 workspace for split-K forward launches (small-M layers): [out pixels][Nc]
  float* splitk_ws; size_t splitk_ws_bytes;
  // im2col first layers (bf16/tcgen05 mode): the op is a 1x1 GEMM whose `ntaps` K-blocks of 64 are
  // separate im2col buffers [M][64] (one per input source); k' = tap16*4 + c inside a block, c < im2col_c.
  const void* in_tap[4]; int im2col_c;
  int n_slot4_c;                 // wgrad: output column n = tap*4 + c (c < n_slot4_c real) maps to master row tap*n_slot4_c + c
  float* out_rows_f32;           // optional: write the raw fp32 accumulators as rows [out pixel][Nc] (no bf16 output)
  // algorithmic (unpadded) GEMM K per class / N for the roofline accounting when the stored operand carries
  // zero slots (im2col / cols rows hold 16 taps x 4 channel slots); 0 = ntaps*Kr / Nr
  int real_k, real_n;
  // storage dtypes (DT_*) of the `in` view (and of the packed weights B, which always match it) and of the `out`
  // view (forward / dgrad: what the epilogue writes; wgrad: the dtype of the output gradient it reads)
  int dt_in, dt_out;
  // optional: fp32 workspace [592][2][Nc] for per-channel sum / sum-of-squares partials produced by the conv epilogue
  // (BatchNorm layers on the CTA-pair tcgen05 kernel); the launcher reports how many partials it wrote
  float* stats_ws;
  // wgrad: 0 = this launch is the first contribution to dW in the step (store), 1 = add to what is there; optional
  // fp32 workspace for the split partial tiles of the deterministic two-stage reduction
  int accumulate;
  float* wgrad_ws; size_t wgrad_ws_bytes;
  long long dW_elems;            // floats of the whole kernel tensor behind dW (the FFMA path zeroes it on a first contribution)
};

// ---- error handling (host) -----------------------------------------------------------------
struct GanError : public std::runtime_error {
  int code;
  GanError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CUDA_CHECK(expr)                                                                         \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      throw GanError(-2, std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " +   \
                             __FILE__ + ":" + std::to_string(__LINE__));                         \
  } while (0)

#define GAN_REQUIRE(cond, msg)                                                                   \
  do {                                                                                           \
    if (!(cond)) throw GanError(-1, std::string(msg) + " (" #cond ") at " + __FILE__ + ":" +    \
                                        std::to_string(__LINE__));                               \
  } while (0)

// ---- device helpers ------------------------------------------------------------------------
#ifdef __CUDACC__
template <typename T> struct VecIO;

template <> struct VecIO<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void load(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

// fp32 -> fp16 with saturation to +-65504 (an overflowing activation must not become inf and poison the
// batch statistics; NaN stays NaN)
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ f16 f16_sat(float x) {
  uint32_t r = pack_f16x2_sat(x, 0.f);
  __half_raw h; h.x = (unsigned short)(r & 0xffffu);
  return f16(h);
}
template <> struct VecIO<f16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void load(const f16* p, float (&v)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  __device__ static __forceinline__ void store(f16* p, const float (&v)[8]) {
    uint4 t;
    t.x = pack_f16x2_sat(v[0], v[1]); t.y = pack_f16x2_sat(v[2], v[3]);
    t.z = pack_f16x2_sat(v[4], v[5]); t.w = pack_f16x2_sat(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

template <> struct VecIO<bf16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void load(const bf16* p, float (&v)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  __device__ static __forceinline__ void store(bf16* p, const float (&v)[8]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

// Thread-private cp.async ring: each thread prefetches the 16-byte vectors it will consume itself
// RING_STAGES-1 iterations ahead into its own shared-memory slots, so the bytes in flight per SM are
// bounded by shared memory (not by registers) and no block-level synchronisation is needed.
#define RING_STAGES 8
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  int sz = pred ? 16 : 0;       // src-size 0: nothing is read, the slot is zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void unpack16(const uint4& t, float (&v)[4], const float*) {
  v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
}
__device__ __forceinline__ void unpack16(const uint4& t, float (&v)[8], const f16*) {
  const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void unpack16(const uint4& t, float (&v)[8], const bf16*) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

// 4-element vector access (16 B fp32 / 8 B bf16): used by the register-heavy normalisation kernels so
// that several independent loads per thread stay in flight without dropping below 3 CTAs/SM.
template <typename T> struct Vec4IO;
template <> struct Vec4IO<float> {
  __device__ static __forceinline__ void load(const float* p, float (&v)[4]) { VecIO<float>::load(p, v); }
  __device__ static __forceinline__ void store(float* p, const float (&v)[4]) { VecIO<float>::store(p, v); }
};
template <> struct Vec4IO<f16> {
  __device__ static __forceinline__ void load(const f16* p, float (&v)[4]) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&t);
    float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  __device__ static __forceinline__ void store(f16* p, const float (&v)[4]) {
    uint2 t;
    t.x = pack_f16x2_sat(v[0], v[1]); t.y = pack_f16x2_sat(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = t;
  }
};
template <> struct Vec4IO<bf16> {
  __device__ static __forceinline__ void load(const bf16* p, float (&v)[4]) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
    float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  __device__ static __forceinline__ void store(bf16* p, const float (&v)[4]) {
    uint2 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
    h[0] = __floats2bfloat162_rn(v[0], v[1]); h[1] = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
__device__ __forceinline__ float to_f(f16 x) { return __half2float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }
template <> __device__ __forceinline__ f16 from_f<f16>(float x) { return f16_sat(x); }

// two fp32 -> one packed 16-bit pair (low half = first value)
template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<bf16>(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
template <> __device__ __forceinline__ uint32_t pack2<f16>(float lo, float hi) { return pack_f16x2_sat(lo, hi); }

// "Last block" pattern for deterministic cross-block sums without a second launch: every block publishes its partial,
// fences, and takes a ticket; exactly one block — the last to arrive, whichever that is — sees all partials and adds
// them up IN INDEX ORDER, so the result does not depend on the arrival order.  The counter resets itself.
__device__ __forceinline__ bool last_block_arrives(unsigned int* counter, unsigned int expected) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(counter, 1u);
    s_last = (t == expected - 1u) ? 1 : 0;
    if (s_last) *counter = 0u;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Philox-4x32-10 (same definition as oracle/gan_oracle.py:philox4x32_10).
__device__ __forceinline__ uint32_t philox_word0(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c0;
}

// Dropout(0.5) keep decision for element `elem` of global sample `sample` (base_gan.py:118).
// The call counter lives in device memory (call = *call_dev + call_off) so that a captured CUDA
// graph draws fresh masks on every replay.
struct DropKey { uint32_t seed_lo, seed_hi; const uint32_t* call_dev; uint32_t call_off, layer; int64_t sample0; int enabled; };
__device__ __forceinline__ bool dropout_keep(const DropKey& k, uint32_t call, int64_t sample_local, uint32_t elem) {
  uint32_t w = philox_word0(elem, (uint32_t)(k.sample0 + sample_local), k.layer, call, k.seed_lo, k.seed_hi);
  return (w >> 31) != 0u;
}
#else
struct DropKey { uint32_t seed_lo, seed_hi; const uint32_t* call_dev; uint32_t call_off, layer; int64_t sample0; int enabled; };
#endif
