// conv_ffma.cu — fp32-accumulate CUDA-core implementation of the "tap GEMM" (DESIGN.md §3).
//
// This is (a) the fp32 precision mode (<=1e-4 against the oracle; TF32/bf16 tensor cores cannot
// hold that through 21 stacked convolutions) and (b) the path for the layers that do not fit a
// tcgen05 tile: the Cin in {1,2,3,6} first layers (base_gan.py:141,180) and the Cout in {1,3}
// heads (base_gan.py:159-161,201-204), which are HBM-bound, not tensor-bound.
//
// One ConvOp covers Conv2D 4x4 s2 'same' (base_gan.py:78), ZeroPad+Conv2D 4x4 s1 (base_gan.py:145-148),
// Conv2DTranspose 4x4 s2 'same' as four parity-class 2x2 correlations (base_gan.py:107), and the
// data-gradients of all three (which are the same shapes with the channel roles swapped).
#include <type_traits>
#include "kernels.h"

#define KLAUNCH(L) (++*(L).count)

namespace {

// Accumulator type: the fp32 precision mode accumulates in double so that the only roundings left
// are those of the stored fp32 operands/results (the step's gradients are badly conditioned: conv
// weight gradients in front of a normalisation layer are small residuals of large cancelling sums);
// the bf16 mode accumulates in fp32 like the tensor cores do.
template <typename T> struct Acc { typedef float type; };
template <> struct Acc<float> { typedef double type; };

constexpr int BM = 64, BN = 64, BK = 16;

template <typename T> __device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void load4<f16>(const f16* p, float (&v)[4]) { Vec4IO<f16>::load(p, v); }
template <> __device__ __forceinline__ void load4<bf16>(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

__device__ __forceinline__ float epilogue(float v, const ConvOp& op, int n) {
  if (op.epi != EPI_NONE) v += op.bias[n];
  if (op.epi == EPI_BIAS_TANH) v = tanhf(v);
  return v;
}

// ---------------------------------------------------------------------------------------------
// Forward-type tap GEMM: out[m, n] = sum_{t,kc} in(m,t,kc) * B[n][t*Kc+kc].
// 64x64 tile, BK=16, 256 threads, 4x4 micro-tile. VEC: Kc%4==0 and 4-element aligned views.
// ---------------------------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) k_conv_fwd(ConvOp op) {
  const ClassGeom& cg = op.cls[blockIdx.z];
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t M = (int64_t)op.N * op.Hm * op.Wm;
  const int K = cg.ntaps * op.Kc;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const T* __restrict__ in = (const T*)op.in;
  const T* __restrict__ Bw = (const T*)op.B + cg.b_off;

  // A-load mapping: one row, 4 consecutive k per thread
  const int arow = tid >> 2, akq = (tid & 3) * 4;
  const int64_t am = m0 + arow;
  const bool am_ok = am < M;
  int a_n = 0, a_h = 0, a_w = 0;
  if (am_ok) { a_w = (int)(am % op.Wm); int64_t r = am / op.Wm; a_h = (int)(r % op.Hm); a_n = (int)(r / op.Hm); }
  // B-load mapping
  const int brow = tid >> 2, bkq = (tid & 3) * 4;
  const int bn = n0 + brow;

  const int tx = tid & 15, ty = tid >> 4;
  typedef typename Acc<T>::type acc_t;
  acc_t acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0;

  for (int k0 = 0; k0 < K; k0 += BK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    const int ka = k0 + akq;
    if (VEC) {
      if (am_ok && ka < K) {
        int t = ka / op.Kc, c = ka - t * op.Kc;
        int ih = a_h * op.si + cg.dh[t], iw = a_w * op.si + cg.dw[t];
        if (ih >= 0 && ih < op.Hin && iw >= 0 && iw < op.Win)
          load4<T>(in + (((int64_t)a_n * op.Hin + ih) * op.Win + iw) * op.in_pitch + op.in_coff + c, av);
      }
      const int kb = k0 + bkq;
      if (bn < op.Nc && kb < K) load4<T>(Bw + (int64_t)bn * K + kb, bv);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int k = ka + e;
        if (am_ok && k < K) {
          int t = k / op.Kc, c = k - t * op.Kc;
          int ih = a_h * op.si + cg.dh[t], iw = a_w * op.si + cg.dw[t];
          if (ih >= 0 && ih < op.Hin && iw >= 0 && iw < op.Win)
            av[e] = to_f(in[(((int64_t)a_n * op.Hin + ih) * op.Win + iw) * op.in_pitch + op.in_coff + c]);
        }
        int kb = k0 + bkq + e;
        if (bn < op.Nc && kb < K) bv[e] = to_f(Bw[(int64_t)bn * K + kb]);
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { As[akq + e][arow] = av[e]; Bs[bkq + e][brow] = bv[e]; }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += (acc_t)a[i] * (acc_t)b[j];
    }
    __syncthreads();
  }

  T* __restrict__ out = (T*)op.out;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    int mw = (int)(m % op.Wm); int64_t r = m / op.Wm; int mh = (int)(r % op.Hm); int n_img = (int)(r / op.Hm);
    int oh = mh * op.so + cg.oa, ow = mw * op.so + cg.ob;
    if (oh >= op.Hout || ow >= op.Wout) continue;
    int64_t pix = ((int64_t)n_img * op.Hout + oh) * op.Wout + ow;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < op.Nc) {
        float v = n < op.Nr ? epilogue((float)acc[i][j], op, n) : 0.f;
        if (out != nullptr) out[pix * op.out_pitch + op.out_coff + n] = from_f<T>(v);
        if (op.out_f32 != nullptr && n < op.Nr) op.out_f32[pix * op.Nr + n] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Skinny-N forward (Nc <= 8: generator head Cout=C, discriminator head Cout=1, D.down1 dgrad):
// one warp per output pixel, lanes stride the channel dimension (coalesced), shuffle reduce.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_conv_fwd_skinny(ConvOp op) {
  const ClassGeom& cg = op.cls[blockIdx.z];
  const int64_t M = (int64_t)op.N * op.Hm * op.Wm;
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  const int K = cg.ntaps * op.Kc;
  const T* __restrict__ in = (const T*)op.in;
  const T* __restrict__ Bw = (const T*)op.B + cg.b_off;
  int mw = (int)(m % op.Wm); int64_t r = m / op.Wm; int mh = (int)(r % op.Hm); int n_img = (int)(r / op.Hm);
  typedef typename Acc<T>::type acc_t;
  acc_t acc[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) acc[n] = 0;
  for (int t = 0; t < cg.ntaps; ++t) {
    int ih = mh * op.si + cg.dh[t], iw = mw * op.si + cg.dw[t];
    if (ih < 0 || ih >= op.Hin || iw < 0 || iw >= op.Win) continue;
    const T* ip = in + (((int64_t)n_img * op.Hin + ih) * op.Win + iw) * op.in_pitch + op.in_coff;
    const T* bp = Bw + (int64_t)t * op.Kc;
    for (int c = lane; c < op.Kc; c += 32) {
      float a = to_f(ip[c]);
#pragma unroll
      for (int n = 0; n < 8; ++n)
        if (n < op.Nc) acc[n] += (acc_t)a * (acc_t)to_f(bp[(int64_t)n * K + c]);
    }
  }
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
  if (lane == 0) {
    int oh = mh * op.so + cg.oa, ow = mw * op.so + cg.ob;
    if (oh < op.Hout && ow < op.Wout) {
      int64_t pix = ((int64_t)n_img * op.Hout + oh) * op.Wout + ow;
      T* out = (T*)op.out;
#pragma unroll
      for (int n = 0; n < 8; ++n)
        if (n < op.Nc) {
          float v = n < op.Nr ? epilogue((float)acc[n], op, n) : 0.f;
          if (out != nullptr) out[pix * op.out_pitch + op.out_coff + n] = from_f<T>(v);
          if (op.out_f32 != nullptr && n < op.Nr) op.out_f32[pix * op.Nr + n] = v;
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Weight gradient: dW[t][kc][nc] += sum_m in(m,t,kc) * dy(m,nc).  64(k) x 64(nc) tile, the pixel
// reduction split over blockIdx.z with fp32 atomics into the (pre-zeroed) gradient buffer.
// ---------------------------------------------------------------------------------------------
template <typename T, typename TD, bool VEC>
__global__ void __launch_bounds__(256) k_conv_wgrad(ConvOp op, int splits) {
  const int ci = blockIdx.z % op.ncls, split = blockIdx.z / op.ncls;
  const ClassGeom& cg = op.cls[ci];
  __shared__ float As[16][BM + 4];
  __shared__ float Ds[16][BN + 4];
  const int tid = threadIdx.x;
  const int64_t M = (int64_t)op.N * op.Hm * op.Wm;
  const int K = cg.ntaps * op.Kc;
  const int k0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int64_t per = ((M + splits - 1) / splits + 15) / 16 * 16;
  const int64_t mbeg = split * per, mend = (mbeg + per < M) ? mbeg + per : M;
  const T* __restrict__ in = (const T*)op.in;
  const TD* __restrict__ dy = (const TD*)op.out;

  const int lm = tid >> 4, lq = (tid & 15) * 4;   // chunk row, 4-wide column group
  // A columns are fixed for the whole loop: precompute tap / channel
  int a_t[4], a_c[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int k = k0 + lq + e;
    if (k < K) { a_t[e] = k / op.Kc; a_c[e] = k - a_t[e] * op.Kc; } else { a_t[e] = -1; a_c[e] = 0; }
  }
  const int tx = tid & 15, ty = tid >> 4;
  typedef typename Acc<T>::type acc_t;
  acc_t acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0;

  for (int64_t mb = mbeg; mb < mend; mb += 16) {
    float av[4] = {0.f, 0.f, 0.f, 0.f}, dv[4] = {0.f, 0.f, 0.f, 0.f};
    int64_t m = mb + lm;
    if (m < mend) {
      int mw = (int)(m % op.Wm); int64_t r = m / op.Wm; int mh = (int)(r % op.Hm); int n_img = (int)(r / op.Hm);
      if (VEC) {
        if (a_t[0] >= 0) {
          int ih = mh * op.si + cg.dh[a_t[0]], iw = mw * op.si + cg.dw[a_t[0]];
          if (ih >= 0 && ih < op.Hin && iw >= 0 && iw < op.Win)
            load4<T>(in + (((int64_t)n_img * op.Hin + ih) * op.Win + iw) * op.in_pitch + op.in_coff + a_c[0], av);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (a_t[e] >= 0) {
            int ih = mh * op.si + cg.dh[a_t[e]], iw = mw * op.si + cg.dw[a_t[e]];
            if (ih >= 0 && ih < op.Hin && iw >= 0 && iw < op.Win)
              av[e] = to_f(in[(((int64_t)n_img * op.Hin + ih) * op.Win + iw) * op.in_pitch + op.in_coff + a_c[e]]);
          }
      }
      int oh = mh * op.so + cg.oa, ow = mw * op.so + cg.ob;
      if (oh < op.Hout && ow < op.Wout) {
        const TD* dp = dy + (((int64_t)n_img * op.Hout + oh) * op.Wout + ow) * op.out_pitch + op.out_coff;
        int n = n0 + lq;
        if (VEC) { if (n < op.Nc) load4<TD>(dp + n, dv); }
        else {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (n + e < op.Nc) dv[e] = to_f(dp[n + e]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { As[lm][lq + e] = av[e]; Ds[lm][lq + e] = dv[e]; }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < 16; ++mm) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[mm][ty * 4 + i]; b[i] = Ds[mm][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += (acc_t)a[i] * (acc_t)b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int k = k0 + ty * 4 + i;
    if (k >= K) continue;
    int t = k / op.Kc, c = k - t * op.Kc;
    if (c >= op.Kr) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < op.Nr) atomicAdd(op.dW + (int64_t)cg.widx[t] * op.s_tap + (int64_t)c * op.s_k + (int64_t)n * op.s_n, (float)acc[i][j]);
    }
  }
}

// Skinny-N weight gradient (Nc <= 4 heads): one thread per input channel, all taps of the class
// kept in registers, pixel range split over blockIdx.y.
template <typename T, typename TD, int NT, int NC>
__global__ void __launch_bounds__(128) k_conv_wgrad_skinny(ConvOp op, int splits) {
  const int ci = blockIdx.z;
  const ClassGeom& cg = op.cls[ci];
  const int c = blockIdx.x * 128 + threadIdx.x;
  const int64_t M = (int64_t)op.N * op.Hm * op.Wm;
  const int64_t per = (M + splits - 1) / splits;
  const int64_t mbeg = blockIdx.y * per, mend = (mbeg + per < M) ? mbeg + per : M;
  const T* __restrict__ in = (const T*)op.in;
  const TD* __restrict__ dy = (const TD*)op.out;
  typedef typename Acc<T>::type acc_t;
  acc_t acc[NT][NC];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int n = 0; n < NC; ++n) acc[t][n] = 0;
  const bool c_ok = c < op.Kc;
  for (int64_t m = mbeg; m < mend; ++m) {
    int mw = (int)(m % op.Wm); int64_t r = m / op.Wm; int mh = (int)(r % op.Hm); int n_img = (int)(r / op.Hm);
    int oh = mh * op.so + cg.oa, ow = mw * op.so + cg.ob;
    if (oh >= op.Hout || ow >= op.Wout) continue;
    const TD* dp = dy + (((int64_t)n_img * op.Hout + oh) * op.Wout + ow) * op.out_pitch + op.out_coff;
    float d[NC];
#pragma unroll
    for (int n = 0; n < NC; ++n) d[n] = to_f(dp[n]);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      int ih = mh * op.si + cg.dh[t], iw = mw * op.si + cg.dw[t];
      if (c_ok && ih >= 0 && ih < op.Hin && iw >= 0 && iw < op.Win) {
        float a = to_f(in[(((int64_t)n_img * op.Hin + ih) * op.Win + iw) * op.in_pitch + op.in_coff + c]);
#pragma unroll
        for (int n = 0; n < NC; ++n) acc[t][n] += (acc_t)a * (acc_t)d[n];
      }
    }
  }
  if (c_ok && c < op.Kr) {
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int n = 0; n < NC; ++n)
        atomicAdd(op.dW + (int64_t)cg.widx[t] * op.s_tap + (int64_t)c * op.s_k + (int64_t)n * op.s_n, (float)acc[t][n]);
  }
}

template <typename T> bool vec_ok(const ConvOp& op, bool wgrad) {
  bool ok = (op.Kc % 4 == 0) && (op.in_pitch % 4 == 0) && (op.in_coff % 4 == 0);
  if (wgrad) ok = ok && (op.Nc % 4 == 0) && (op.out_pitch % 4 == 0) && (op.out_coff % 4 == 0);
  return ok;
}

}  // namespace

void launch_conv_fwd_ffma(Launch L, int dt, const ConvOp& op) {
  const int64_t M = (int64_t)op.N * op.Hm * op.Wm;
  auto run = [&](auto* tag) {
    using T = typename std::remove_pointer<decltype(tag)>::type;
    if (op.Nc <= 8) {
      dim3 grid((unsigned)((M + 7) / 8), 1, op.ncls);
      k_conv_fwd_skinny<T><<<grid, 256, 0, L.s>>>(op);
    } else {
      dim3 grid((unsigned)((M + BM - 1) / BM), (op.Nc + BN - 1) / BN, op.ncls);
      if (vec_ok<T>(op, false)) k_conv_fwd<T, true><<<grid, 256, 0, L.s>>>(op);
      else k_conv_fwd<T, false><<<grid, 256, 0, L.s>>>(op);
    }
  };
  if (dt == DT_F32) run((float*)nullptr); else if (dt == DT_F16) run((f16*)nullptr); else run((bf16*)nullptr);
  KLAUNCH(L);
}

// dt_in: dtype of the layer input x; dt_dy: dtype of the output gradient (always equal: see common.cuh)
void launch_conv_wgrad_ffma(Launch L, int dt_in, int dt_dy, const ConvOp& op) {
  const int64_t M = (int64_t)op.N * op.Hm * op.Wm;
  auto run = [&](auto* tag, auto* dtag) {
    using T = typename std::remove_pointer<decltype(tag)>::type;
    using TD = typename std::remove_pointer<decltype(dtag)>::type;
    const int nt = op.cls[0].ntaps;
    if (op.Nc <= 4 && (nt == 4 || nt == 16) && (op.Nc == 1 || op.Nc == 3 || nt == 4)) {
      int cblocks = (op.Kc + 127) / 128;
      int64_t want = (148 * 8) / (cblocks * op.ncls);
      int splits = (int)(want < 1 ? 1 : (want > M / 64 + 1 ? M / 64 + 1 : want));
      dim3 grid(cblocks, splits, op.ncls);
#define SKINNY(NT_, NC_) k_conv_wgrad_skinny<T, TD, NT_, NC_><<<grid, 128, 0, L.s>>>(op, splits)
      if (nt == 16 && op.Nc == 1) SKINNY(16, 1);
      else if (nt == 16 && op.Nc == 3) SKINNY(16, 3);
      else if (nt == 4 && op.Nc == 1) SKINNY(4, 1);
      else if (nt == 4 && op.Nc == 2) SKINNY(4, 2);
      else if (nt == 4 && op.Nc == 3) SKINNY(4, 3);
      else SKINNY(4, 4);
#undef SKINNY
      return;
    }
    int K = nt * op.Kc;
    int kt = (K + BM - 1) / BM, ntile = (op.Nc + BN - 1) / BN;
    int64_t tiles = (int64_t)kt * ntile * op.ncls;
    int64_t want = (148 * 4 + tiles - 1) / tiles;
    int64_t maxs = (M + 255) / 256;
    int splits = (int)(want < 1 ? 1 : (want > maxs ? maxs : want));
    if (splits < 1) splits = 1;
    dim3 grid(kt, ntile, op.ncls * splits);
    if (vec_ok<T>(op, true)) k_conv_wgrad<T, TD, true><<<grid, 256, 0, L.s>>>(op, splits);
    else k_conv_wgrad<T, TD, false><<<grid, 256, 0, L.s>>>(op, splits);
  };
  if (dt_in == DT_F32) run((float*)nullptr, (float*)nullptr);
  else if (dt_in == DT_F16) run((f16*)nullptr, (f16*)nullptr);
  else run((bf16*)nullptr, (bf16*)nullptr);
  (void)dt_dy;
  KLAUNCH(L);
}
