// elem.cu — memory-bound kernels of the GAN train step: layout conversion, normalisation
// statistics / apply / backward (BatchNorm base_gan.py:83,113,151 and InstanceNormalization
// utils.py:6-30), activations (LeakyReLU 0.3 base_gan.py:87,155; ReLU :120; Dropout 0.5 :118),
// losses (BCE-from-logits base_gan.py:227-245, L1 pix2pix.py:181 / cycle_gan.py:167,176),
// Keras Adam (base_gan.py:252) and weight packing.
//
// All of them are HBM-bound: 128-bit loads/stores, grid sized as a multiple of the 148 SMs,
// fp32 math with double-precision final reductions, no atomics on the value paths except the
// tiny bias-gradient sums.
#include "kernels.h"
#include <cstdlib>

#define NSM 148

template <typename F> static void dispatch_dt(int dt, F&& f) {
  if (dt == DT_F32) f((float*)nullptr); else if (dt == DT_F16) f((f16*)nullptr); else f((bf16*)nullptr);
}
// (activation dtype, gradient dtype) pairs: fp32/fp32, f16/f16 (default 16-bit mode), bf16/bf16
template <typename F> static void dispatch_dt2(int dtz, int dtg, F&& f) {
  GAN_REQUIRE(dtz == dtg, "activations and gradients share one storage format (tcgen05 kind::f16 needs equal A/B formats)");
  if (dtz == DT_F32) f((float*)nullptr, (float*)nullptr);
  else if (dtz == DT_F16) f((f16*)nullptr, (f16*)nullptr);
  else f((bf16*)nullptr, (bf16*)nullptr);
}
#define KLAUNCH(L) (++*(L).count)

static inline int grid_for(int64_t work, int threads, int max_waves = 8) {
  int64_t b = (work + threads - 1) / threads;
  int64_t cap = (int64_t)NSM * max_waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---------------------------------------------------------------------------------------------
// fp32 NHWC (compact) <-> strided view of dtype T
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_convert(const float* __restrict__ src, int64_t total, int C, T* __restrict__ dst, int pitch,
                          int coff) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / C; int c = (int)(i - p * C);
    dst[p * pitch + coff + c] = from_f<T>(src[i]);
  }
}
void launch_convert(Launch L, int dt, const float* src, int64_t P, int C, void* dst, int pitch, int coff) {
  int64_t total = P * C;
  dispatch_dt(dt, [&](auto* tag) {
    using T = typename std::remove_pointer<decltype(tag)>::type;
    k_convert<T><<<grid_for(total, 256), 256, 0, L.s>>>(src, total, C, (T*)dst, pitch, coff);
  });
  KLAUNCH(L);
}

template <typename T>
__global__ void k_export(const T* __restrict__ src, int pitch, int coff, int64_t total, int C, float* __restrict__ dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / C; int c = (int)(i - p * C);
    dst[i] = to_f(src[p * pitch + coff + c]);
  }
}
void launch_export(Launch L, int dt, const void* src, int pitch, int coff, int64_t P, int C, float* dst) {
  int64_t total = P * C;
  dispatch_dt(dt, [&](auto* tag) {
    using T = typename std::remove_pointer<decltype(tag)>::type;
    k_export<T><<<grid_for(total, 256), 256, 0, L.s>>>((const T*)src, pitch, coff, total, C, dst);
  });
  KLAUNCH(L);
}

// ---------------------------------------------------------------------------------------------
// Normalisation kernels.  z is compact [G*Pg][C]; group g covers rows [g*Pg,(g+1)*Pg) (G == 1:
// BatchNorm, G == N: InstanceNorm).  C is a power of two >= 64 for every normalised layer, so a
// thread owns ONE channel vector (V = 8 bf16 / 4 fp32 channels) for its whole lifetime: the
// per-channel parameters sit in registers, index math is shifts, and each thread keeps UNR
// independent 128-bit loads in flight (HBM-bound: bytes in flight per SM is what matters).
// ---------------------------------------------------------------------------------------------
#define NORM_UNR 4
#define NORM_BWD_UNR 4

int stats_chunks(int G, int64_t Pg) {
  int64_t want = (STATS_MAX_CHUNKS + G - 1) / G;         // ~4 CTAs per SM across all groups
  int64_t maxc = (Pg + 63) / 64;                         // at least 64 rows per chunk
  int64_t c = want < maxc ? want : maxc;
  return (int)(c < 1 ? 1 : c);
}
size_t stats_ws_floats(int G, int64_t Pg, int C) { return (size_t)G * stats_chunks(G, Pg) * 2 * C; }

static inline int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// Block-level column reduction of per-thread vectors (threads with equal `col` own the same channels).
template <int V>
__device__ __forceinline__ void block_col_reduce(const float (&s)[V], const float (&q)[V], int cv, int C, float* sh_s,
                                                 float* sh_q, float* __restrict__ out) {
  const int rows_par = 256 / cv;
#pragma unroll
  for (int i = 0; i < V; ++i) { sh_s[threadIdx.x * V + i] = s[i]; sh_q[threadIdx.x * V + i] = q[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float S = 0.f, Q = 0.f;
    for (int rr = 0; rr < rows_par; ++rr) {
      int t = rr * cv + c / V;
      S += sh_s[t * V + c % V]; Q += sh_q[t * V + c % V];
    }
    out[c] = S; out[C + c] = Q;
  }
}

// Stage 1 of the moments: per-chunk (sum, sumsq) per channel in fp32.
template <typename T>
__global__ void __launch_bounds__(256) k_stats_partial(const T* __restrict__ z, uint32_t Pg, int C, int lcv, int nchunk,
                                                       float* __restrict__ ws) {
  constexpr int V = 4;
  __shared__ float sh_s[256 * V];
  __shared__ float sh_q[256 * V];
  const int g = blockIdx.y, chunk = blockIdx.x;
  const int cv = 1 << lcv;
  const uint32_t rows_par = 256u >> lcv;
  const int col = threadIdx.x & (cv - 1);
  const uint32_t r = threadIdx.x >> lcv;
  const uint32_t per = (Pg + nchunk - 1) / nchunk;
  const uint32_t p0 = chunk * per, p1 = min(Pg, p0 + per);
  float s[V], q[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { s[i] = 0.f; q[i] = 0.f; }
  const T* base = z + ((size_t)g * Pg) * C + col * V;
  for (uint32_t p = p0 + r; p < p1; p += rows_par * NORM_UNR) {
    float v[NORM_UNR][V];
#pragma unroll
    for (int u = 0; u < NORM_UNR; ++u) {
      uint32_t pp = p + u * rows_par;
      if (pp < p1) Vec4IO<T>::load(base + (size_t)pp * C, v[u]);
      else {
#pragma unroll
        for (int i = 0; i < V; ++i) v[u][i] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < NORM_UNR; ++u)
#pragma unroll
      for (int i = 0; i < V; ++i) { s[i] += v[u][i]; q[i] = fmaf(v[u][i], v[u][i], q[i]); }
  }
  block_col_reduce<V>(s, q, cv, C, sh_s, sh_q, ws + ((size_t)(g * nchunk + chunk) * 2) * C);
}

// Stage 2: one block per (32 channels, group): 8 chunk-lanes per channel sum the partials in
// double, shared-memory tree over the lanes, then mean / inv-std / fused scale are emitted.
#define FIN_LANES 32
__device__ __forceinline__ void fin_reduce(const float* __restrict__ ws, int g, int nchunk, int C, int c, int lane,
                                           double& S, double& Q) {
  __shared__ double sh[2][FIN_LANES][32];
  double s = 0.0, q = 0.0;
  if (c < C) {
    int k = lane;
    for (; k + 3 * FIN_LANES < nchunk; k += 4 * FIN_LANES) {       // 8 independent loads in flight per thread
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* o = ws + ((size_t)(g * nchunk + k + u * FIN_LANES) * 2) * C;
        a[u] = __ldg(o + c); b[u] = __ldg(o + C + c);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { s += (double)a[u]; q += (double)b[u]; }
    }
    for (; k < nchunk; k += FIN_LANES) {
      const float* o = ws + ((size_t)(g * nchunk + k) * 2) * C;
      s += (double)__ldg(o + c); q += (double)__ldg(o + C + c);
    }
  }
  sh[0][lane][threadIdx.x & 31] = s; sh[1][lane][threadIdx.x & 31] = q;
  __syncthreads();
  S = 0.0; Q = 0.0;
  if (lane == 0) {
#pragma unroll
    for (int l = 0; l < FIN_LANES; ++l) { S += sh[0][l][threadIdx.x & 31]; Q += sh[1][l][threadIdx.x & 31]; }
  }
}

__global__ void __launch_bounds__(32 * FIN_LANES) k_stats_finalize(
    const float* __restrict__ ws, int G, int nchunk, int C, double n, float eps, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* __restrict__ mean, float* __restrict__ inv, float* __restrict__ scale,
    float* __restrict__ shift, float* mov_mean, float* mov_var, float momentum) {
  const int g = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), lane = threadIdx.x >> 5;
  double S, Q;
  fin_reduce(ws, g, nchunk, C, c, lane, S, Q);
  if (lane != 0 || c >= C) return;
  const size_t i = (size_t)g * C + c;
  double m = S / n;
  double var = Q / n - m * m;
  if (var < 0.0) var = 0.0;
  double iv = 1.0 / sqrt(var + (double)eps);
  mean[i] = (float)m; inv[i] = (float)iv;
  scale[i] = (float)((double)gamma[c] * iv);
  shift[i] = beta[c];     // u = (z - mean)*scale + beta: exactly beta when z == mean (n == 1)
  if (mov_mean != nullptr) {
    // Keras BatchNormalization moving averages (momentum 0.99); the fused TF kernel feeds the
    // Bessel-corrected variance.  Never read on the hot path (every call is training=True).
    double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
    mov_mean[c] = momentum * mov_mean[c] + (1.f - momentum) * (float)m;
    mov_var[c] = momentum * mov_var[c] + (1.f - momentum) * (float)unbiased;
  }
}

void launch_norm_stats(Launch L, int dt, const void* z, int G, int64_t Pg, int C, float* ws, float eps,
                       const float* gamma, const float* beta, float* mean, float* inv, float* scale,
                       float* shift, float* mov_mean, float* mov_var, float momentum) {
  int nchunk = stats_chunks(G, Pg);
  dispatch_dt(dt, [&](auto* tag) {
    using T = typename std::remove_pointer<decltype(tag)>::type;
    const int cv = C / 4;
    GAN_REQUIRE((cv & (cv - 1)) == 0 && cv <= 256, "normalised channel count must be a power of two");
    k_stats_partial<T><<<dim3(nchunk, G), 256, 0, L.s>>>((const T*)z, (uint32_t)Pg, C, ilog2(cv), nchunk, ws);
  });
  KLAUNCH(L);
  k_stats_finalize<<<dim3((C + 31) / 32, G), 32 * FIN_LANES, 0, L.s>>>(ws, G, nchunk, C, (double)Pg, eps, gamma, beta, mean,
                                                                      inv, scale, shift, mov_mean, mov_var, momentum);
  KLAUNCH(L);
}

// Stage 2 alone: the partials were produced elsewhere (the CTA-pair conv epilogue writes one [2][C] partial per
// (CTA, TMEM quadrant) from its fp32 accumulators), layout ws[part][2][C].
void launch_norm_stats_finalize(Launch L, const float* ws, int nparts, int64_t n, int C, float eps, const float* gamma,
                                const float* beta, float* mean, float* inv, float* scale, float* shift, float* mov_mean,
                                float* mov_var, float momentum) {
  k_stats_finalize<<<dim3((C + 31) / 32, 1), 32 * FIN_LANES, 0, L.s>>>(ws, 1, nparts, C, (double)n, eps, gamma, beta, mean, inv,
                                                                      scale, shift, mov_mean, mov_var, momentum);
  KLAUNCH(L);
}

// ---------------------------------------------------------------------------------------------
// out = act(dropout((z-mean)*scale + shift)) written into the consumer's (concat-offset) view.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_fwd(float u, int act) {
  if (act == ACT_LEAKY) return u > 0.f ? u : LEAKY_SLOPE * u;
  if (act == ACT_RELU) return u > 0.f ? u : 0.f;
  if (act == ACT_TANH) return tanhf(u);
  return u;
}
__device__ __forceinline__ float act_bwd(float u, int act) {
  if (act == ACT_LEAKY) return u > 0.f ? 1.f : LEAKY_SLOPE;
  if (act == ACT_RELU) return u > 0.f ? 1.f : 0.f;
  return 1.f;
}

template <int V>
struct ChanParams { float mu[V], sc[V], sf[V]; };

template <int V>
__device__ __forceinline__ void load_chan_params(ChanParams<V>& cp, const float* mean, const float* scale,
                                                 const float* shift, size_t off) {
#pragma unroll
  for (int k = 0; k < V; ++k) { cp.mu[k] = __ldg(mean + off + k); cp.sc[k] = __ldg(scale + off + k); cp.sf[k] = __ldg(shift + off + k); }
}

template <typename T, bool DROP>
__global__ void __launch_bounds__(256) k_norm_apply(const T* __restrict__ z, uint32_t P, uint32_t Pg, int G, uint32_t HW,
                                                    int C, int lcv, const float* __restrict__ mean,
                                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                                    int act, DropKey dk, T* __restrict__ out, int out_pitch, int out_coff) {
  constexpr int V = VecIO<T>::N;
  extern __shared__ uint4 ring[];                      // [RING_STAGES][256] thread-private slots
  const uint32_t tid = blockIdx.x * 256u + threadIdx.x;
  const int c0 = (int)(tid & ((1u << lcv) - 1)) * V;
  const uint32_t prow = tid >> lcv, pstride = (gridDim.x * 256u) >> lcv;
  const uint32_t nit = prow < P ? (P - prow + pstride - 1) / pstride : 0;
  const bool affine = scale != nullptr;
  const uint32_t call = DROP ? __ldg(dk.call_dev) + dk.call_off : 0u;
  ChanParams<V> cp;
  int cur_g = -1;
  if (affine && G == 1) { load_chan_params<V>(cp, mean, scale, shift, c0); cur_g = 0; }
  auto issue = [&](uint32_t it) {
    const bool ok = it < nit;
    cp_async16(&ring[(it % RING_STAGES) * 256 + threadIdx.x], z + (ok ? (size_t)(prow + it * pstride) * C + c0 : 0), ok);
    cp_async_commit();
  };
  for (uint32_t it = 0; it < RING_STAGES - 1; ++it) issue(it);
  for (uint32_t it = 0; it < nit; ++it) {
    issue(it + RING_STAGES - 1);
    cp_async_wait<RING_STAGES - 1>();
    const uint32_t pp = prow + it * pstride;
    float v[V];
    unpack16(ring[(it % RING_STAGES) * 256 + threadIdx.x], v, (const T*)nullptr);
    if (affine) {
      if (G != 1) {
        int g = (int)(pp / Pg);
        if (g != cur_g) { load_chan_params<V>(cp, mean, scale, shift, (size_t)g * C + c0); cur_g = g; }
      }
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] = fmaf(v[k] - cp.mu[k], cp.sc[k], cp.sf[k]);
    }
    if (DROP) {
      uint32_t smp = pp / HW, e0 = (pp - smp * HW) * C + c0;
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] = dropout_keep(dk, call, smp, e0 + k) ? 2.f * v[k] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = act_fwd(v[k], act);
    VecIO<T>::store(out + (size_t)pp * out_pitch + out_coff + c0, v);
  }
}

// grid for the row-strided ring kernels: every thread owns one channel vector, rows are strided over
// the grid; `per_sm` CTAs per SM (shared-memory bound), fewer for tiny tensors.
static inline int norm_grid(int64_t P, int cv, int per_sm) {
  int64_t rows_per_block = 256 / cv;
  int64_t want = (P + rows_per_block - 1) / rows_per_block;
  int64_t cap = (int64_t)NSM * per_sm;
  if (want > cap) want = cap;
  return (int)(want < 1 ? 1 : want);
}
template <typename K> static void set_smem(K kern, size_t bytes) {
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

void launch_norm_apply(Launch L, int dt, const void* z, int64_t P, int64_t Pg, int G, int HW, int C,
                       const float* mean, const float* scale, const float* shift, int act, DropKey dk, void* out,
                       int out_pitch, int out_coff) {
  dispatch_dt(dt, [&](auto* tag) {
    using T = typename std::remove_pointer<decltype(tag)>::type;
    const int cv = C / VecIO<T>::N;
    GAN_REQUIRE((cv & (cv - 1)) == 0 && cv <= 256 && cv >= 1, "channel count must be a power of two");
    const size_t smem = (size_t)RING_STAGES * 256 * 16;
    auto kern = dk.enabled ? k_norm_apply<T, true> : k_norm_apply<T, false>;
    static bool once = (set_smem(k_norm_apply<T, true>, smem), set_smem(k_norm_apply<T, false>, smem), true);
    (void)once;
    kern<<<norm_grid(P, cv, 5), 256, smem, L.s>>>((const T*)z, (uint32_t)P, (uint32_t)Pg, G, (uint32_t)HW, C, ilog2(cv),
                                                  mean, scale, shift, act, dk, (T*)out, out_pitch, out_coff);
  });
  KLAUNCH(L);
}

// ---------------------------------------------------------------------------------------------
// Backward of norm -> dropout -> activation.
//   g    = (d1 + d2) * act'(.) * dropout'      (gradient w.r.t. the affine output u = gamma*xhat+beta)
//   dbeta = sum g, dgamma = sum g*xhat,  dz = gamma*inv * (g - mean(g) - xhat*mean(g*xhat))
// ---------------------------------------------------------------------------------------------
// Issue the cp.async copies of one iteration (z row vector + one or two gradient sources).
template <typename TZ, typename T>
__device__ __forceinline__ void ring_issue3(uint4* ring, uint32_t it, bool ok, const TZ* z, size_t zoff, const GradSrc& d1,
                                            const GradSrc& d2, size_t p, int c0) {
  const int narr = d2.p != nullptr ? 3 : 2;
  uint4* slot = ring + ((it % RING_STAGES) * narr) * 256 + threadIdx.x;
  cp_async16(slot, z + (ok ? zoff : 0), ok);
  cp_async16(slot + 256, (const T*)d1.p + (ok ? p * d1.pitch + d1.coff + c0 : 0), ok);
  if (narr == 3) cp_async16(slot + 512, (const T*)d2.p + (ok ? p * d2.pitch + d2.coff + c0 : 0), ok);
  cp_async_commit();
}
template <typename TZ, typename T, int V>
__device__ __forceinline__ void ring_read3(const uint4* ring, uint32_t it, bool has_d2, float (&v)[V], float (&g)[V]) {
  const int narr = has_d2 ? 3 : 2;
  const uint4* slot = ring + ((it % RING_STAGES) * narr) * 256 + threadIdx.x;
  unpack16(slot[0], v, (const TZ*)nullptr);
  unpack16(slot[256], g, (const T*)nullptr);
  if (has_d2) {
    float h[V];
    unpack16(slot[512], h, (const T*)nullptr);
#pragma unroll
    for (int k = 0; k < V; ++k) g[k] += h[k];
  }
}

template <typename TZ, typename T, bool DROP>
__global__ void __launch_bounds__(256) k_bwd_reduce(const TZ* __restrict__ z, GradSrc d1, GradSrc d2, uint32_t Pg, uint32_t HW,
                                                    int C, int lcv, int nchunk, const float* __restrict__ mean,
                                                    const float* __restrict__ inv, const float* __restrict__ scale,
                                                    const float* __restrict__ shift, int act, DropKey dk,
                                                    float* __restrict__ ws) {
  constexpr int V = VecIO<T>::N;
  extern __shared__ uint4 ring[];
  const int g = blockIdx.y, chunk = blockIdx.x;
  const int cv = 1 << lcv;
  const uint32_t rows_par = 256u >> lcv;
  const int c0 = (threadIdx.x & (cv - 1)) * V;
  const uint32_t r = threadIdx.x >> lcv;
  const uint32_t per = (Pg + nchunk - 1) / nchunk;
  const uint32_t p0 = chunk * per, p1 = min(Pg, p0 + per);
  const uint32_t first = p0 + r;
  const uint32_t nit = first < p1 ? (p1 - first + rows_par - 1) / rows_par : 0;
  const bool has_d2 = d2.p != nullptr;
  float s[V], q[V];
  ChanParams<V> cp;
  const uint32_t call = DROP ? __ldg(dk.call_dev) + dk.call_off : 0u;
  load_chan_params<V>(cp, mean, scale, shift, (size_t)g * C + c0);
#pragma unroll
  for (int i = 0; i < V; ++i) { s[i] = 0.f; q[i] = 0.f; }
  auto issue = [&](uint32_t it) {
    const bool ok = it < nit;
    const size_t p = (size_t)g * Pg + first + (size_t)it * rows_par;
    ring_issue3<TZ, T>(ring, it, ok, z, p * C + c0, d1, d2, p, c0);
  };
  for (uint32_t it = 0; it < RING_STAGES - 1; ++it) issue(it);
  for (uint32_t it = 0; it < nit; ++it) {
    issue(it + RING_STAGES - 1);
    cp_async_wait<RING_STAGES - 1>();
    float v[V], gr[V];
    ring_read3<TZ, T, V>(ring, it, has_d2, v, gr);
    const uint32_t pg = g * Pg + first + it * rows_par;
    uint32_t smp = 0, e0 = 0;
    if (DROP) { smp = pg / HW; e0 = (pg - smp * HW) * C + c0; }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float xc = v[k] - cp.mu[k];
      float uu = fmaf(xc, cp.sc[k], cp.sf[k]);
      float gg = gr[k] * act_bwd(uu, act);
      if (DROP) gg = dropout_keep(dk, call, smp, e0 + k) ? 2.f * gg : 0.f;
      s[k] += gg; q[k] = fmaf(gg, xc, q[k]);      // x_hat = xc*inv: inv is applied once after the loop
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) q[i] *= __ldg(inv + (size_t)g * C + c0 + i);
  cp_async_wait<0>();
  __syncthreads();                                   // ring slots of other threads are reused for the reduction
  float* sh_s = reinterpret_cast<float*>(ring);
  float* sh_q = sh_s + 256 * V;
  block_col_reduce<V>(s, q, cv, C, sh_s, sh_q, ws + ((size_t)(g * nchunk + chunk) * 2) * C);
}

__global__ void __launch_bounds__(32 * FIN_LANES) k_bwd_finalize(const float* __restrict__ ws, int G, int nchunk, int C,
                                                                 double n, float* __restrict__ c1, float* __restrict__ c2,
                                                                 float* dgamma, float* dbeta, unsigned int* counters) {
  const int g = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), lane = threadIdx.x >> 5;
  double S, Q;
  fin_reduce(ws, g, nchunk, C, c, lane, S, Q);
  if (lane == 0 && c < C) {
    c1[(size_t)g * C + c] = (float)(S / n); c2[(size_t)g * C + c] = (float)(Q / n);
    // parameter gradients: BatchNorm has one group, so this is the only writer of the channel in this launch (launches
    // of a step are ordered by the stream: deterministic)
    if (G == 1) { dbeta[c] += (float)S; dgamma[c] += (float)Q; }
  }
  if (G == 1) return;
  // InstanceNorm: the last block of this channel strip sums the groups in index order
  if (!last_block_arrives(counters + blockIdx.x, (unsigned int)G)) return;
  if (lane == 0 && c < C) {
    float sb = 0.f, sg = 0.f;
    for (int gg = 0; gg < G; ++gg) { sb += __ldcg(c1 + (size_t)gg * C + c) * (float)n; sg += __ldcg(c2 + (size_t)gg * C + c) * (float)n; }
    dbeta[c] += sb; dgamma[c] += sg;
  }
}

template <typename TZ, typename T, bool DROP>
__global__ void __launch_bounds__(256) k_bwd_apply(const TZ* __restrict__ z, GradSrc d1, GradSrc d2, uint32_t P, uint32_t Pg,
                                                   int G, uint32_t HW, int C, int lcv, int norm,
                                                   const float* __restrict__ mean, const float* __restrict__ inv,
                                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                                   const float* __restrict__ c1, const float* __restrict__ c2, int act,
                                                   DropKey dk, T* __restrict__ dz, int z_pitch, int z_coff) {
  constexpr int V = VecIO<T>::N;
  extern __shared__ uint4 ring[];
  const uint32_t tid = blockIdx.x * 256u + threadIdx.x;
  const int c0 = (int)(tid & ((1u << lcv) - 1)) * V;
  const uint32_t prow = tid >> lcv, pstride = (gridDim.x * 256u) >> lcv;
  const uint32_t nit = prow < P ? (P - prow + pstride - 1) / pstride : 0;
  const bool has_d2 = d2.p != nullptr;
  ChanParams<V> cp;
  float iv[V], k1[V], k2[V];
  int cur_g = -1;
  const uint32_t call = DROP ? __ldg(dk.call_dev) + dk.call_off : 0u;
  auto load_all = [&](int g) {
    size_t off = (size_t)g * C + c0;
    load_chan_params<V>(cp, mean, scale, shift, off);
#pragma unroll
    for (int k = 0; k < V; ++k) { iv[k] = __ldg(inv + off + k); k1[k] = __ldg(c1 + off + k); k2[k] = __ldg(c2 + off + k); }
    cur_g = g;
  };
  if (norm != NORM_NONE && G == 1) load_all(0);
  auto issue = [&](uint32_t it) {
    const bool ok = it < nit;
    const size_t p = (size_t)prow + (size_t)it * pstride;
    ring_issue3<TZ, T>(ring, it, ok, z, p * z_pitch + z_coff + c0, d1, d2, p, c0);
  };
  for (uint32_t it = 0; it < RING_STAGES - 1; ++it) issue(it);
  for (uint32_t it = 0; it < nit; ++it) {
    issue(it + RING_STAGES - 1);
    cp_async_wait<RING_STAGES - 1>();
    const uint32_t pp = prow + it * pstride;
    float v[V], gr[V], o[V];
    ring_read3<TZ, T, V>(ring, it, has_d2, v, gr);
    if (norm == NORM_NONE) {
#pragma unroll
      for (int k = 0; k < V; ++k) o[k] = gr[k] * act_bwd(v[k], act);
    } else {
      if (G != 1) { int g = (int)(pp / Pg); if (g != cur_g) load_all(g); }
      uint32_t smp = 0, e0 = 0;
      if (DROP) { smp = pp / HW; e0 = (pp - smp * HW) * C + c0; }
#pragma unroll
      for (int k = 0; k < V; ++k) {
        float xc = v[k] - cp.mu[k];
        float uu = fmaf(xc, cp.sc[k], cp.sf[k]);
        float gg = gr[k] * act_bwd(uu, act);
        if (DROP) gg = dropout_keep(dk, call, smp, e0 + k) ? 2.f * gg : 0.f;
        o[k] = cp.sc[k] * (gg - k1[k] - xc * iv[k] * k2[k]);
      }
    }
    VecIO<T>::store(dz + (size_t)pp * C + c0, o);
  }
}

// ---------------------------------------------------------------------------------------------
// Small BatchNorm layers (the U-Net bottleneck; every layer of a small per-GPU batch): the three
// launches of the chain above are latency-bound there, so ONE kernel does the whole layer.  A block
// owns one 16-byte channel vector: it pulls its [P][V] slab into shared memory with cp.async (all
// loads in flight at once), reduces in double, and applies from shared memory.
// ---------------------------------------------------------------------------------------------
#define BNS_THREADS 512
#define BNS_MAX_SMEM (192 * 1024)

template <int V>
__device__ __forceinline__ void bns_block_sums(const float (&s)[V], const float (&q)[V], double (*red)[2 * V], double& S,
                                               double& Q) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    float a = warp_sum(s[k]), b = warp_sum(q[k]);
    if (lane == 0) { red[warp][k] = (double)a; red[warp][V + k] = (double)b; }
  }
  __syncthreads();
  S = 0.0; Q = 0.0;
  if (threadIdx.x < V) {
#pragma unroll
    for (int w = 0; w < BNS_THREADS / 32; ++w) { S += red[w][threadIdx.x]; Q += red[w][V + threadIdx.x]; }
  }
}

template <typename T, bool DROP>
__global__ void __launch_bounds__(BNS_THREADS) k_bn_small_fwd(
    const T* __restrict__ z_all, uint32_t P, uint32_t HW, int C, float eps, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* __restrict__ mean, float* __restrict__ inv, float* __restrict__ scale,
    float* __restrict__ shift, float* mov_mean, float* mov_var, float momentum, int act, DropKey dk, T* __restrict__ out_all,
    int out_pitch, int out_coff) {
  // P = pixels of one normalisation group; blockIdx.y = group (BatchNorm: one group = the whole batch;
  // InstanceNorm: one group per sample)
  constexpr int V = VecIO<T>::N;
  extern __shared__ uint4 slab[];                       // [P] vectors of this block's V channels
  __shared__ double red[BNS_THREADS / 32][2 * V];
  __shared__ float par[3 * V];
  const int c0 = blockIdx.x * V;
  const uint32_t g = blockIdx.y;
  const size_t gi = (size_t)g * C;                      // row of this group in the statistics arrays
  const T* z = z_all + (size_t)g * P * C;
  T* out = out_all + (size_t)g * P * out_pitch;
  for (uint32_t p = threadIdx.x; p < P; p += BNS_THREADS) cp_async16(&slab[p], z + (size_t)p * C + c0, true);
  cp_async_commit();
  cp_async_wait<0>();                                   // a thread only ever reads the slots it filled itself
  float s[V], q[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { s[k] = 0.f; q[k] = 0.f; }
  for (uint32_t p = threadIdx.x; p < P; p += BNS_THREADS) {
    float v[V];
    unpack16(slab[p], v, (const T*)nullptr);
#pragma unroll
    for (int k = 0; k < V; ++k) { s[k] += v[k]; q[k] = fmaf(v[k], v[k], q[k]); }
  }
  double S, Q;
  bns_block_sums<V>(s, q, red, S, Q);
  if (threadIdx.x < V) {                                // same arithmetic as k_stats_finalize
    const int c = c0 + threadIdx.x;
    const double n = (double)P;
    double m = S / n;
    double var = Q / n - m * m;
    if (var < 0.0) var = 0.0;
    double iv = 1.0 / sqrt(var + (double)eps);
    float sc = (float)((double)gamma[c] * iv), sf = beta[c];
    mean[gi + c] = (float)m; inv[gi + c] = (float)iv; scale[gi + c] = sc; shift[gi + c] = sf;
    par[threadIdx.x] = (float)m; par[V + threadIdx.x] = sc; par[2 * V + threadIdx.x] = sf;
    if (mov_mean != nullptr) {
      double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
      mov_mean[c] = momentum * mov_mean[c] + (1.f - momentum) * (float)m;
      mov_var[c] = momentum * mov_var[c] + (1.f - momentum) * (float)unbiased;
    }
  }
  __syncthreads();
  float mu[V], sc[V], sf[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { mu[k] = par[k]; sc[k] = par[V + k]; sf[k] = par[2 * V + k]; }
  const uint32_t call = DROP ? __ldg(dk.call_dev) + dk.call_off : 0u;
  for (uint32_t p = threadIdx.x; p < P; p += BNS_THREADS) {
    float v[V];
    unpack16(slab[p], v, (const T*)nullptr);
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = fmaf(v[k] - mu[k], sc[k], sf[k]);
    if (DROP) {
      const uint32_t pgl = g * P + p;
      uint32_t smp = pgl / HW, e0 = (pgl - smp * HW) * C + c0;
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] = dropout_keep(dk, call, smp, e0 + k) ? 2.f * v[k] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = act_fwd(v[k], act);
    VecIO<T>::store(out + (size_t)p * out_pitch + out_coff + c0, v);
  }
}

template <typename TZ, typename T, bool DROP>
__global__ void __launch_bounds__(BNS_THREADS) k_bn_small_bwd(
    const TZ* __restrict__ z_all, GradSrc d1, GradSrc d2, uint32_t P, uint32_t HW, int C, const float* __restrict__ mean,
    const float* __restrict__ inv, const float* __restrict__ scale, const float* __restrict__ shift, int act, DropKey dk,
    float* __restrict__ c1, float* __restrict__ c2, float* dgamma, float* dbeta, T* __restrict__ dz_all, unsigned int* counters) {
  constexpr int V = VecIO<T>::N;
  extern __shared__ uint4 slab[];                       // [narr][P]
  __shared__ double red[BNS_THREADS / 32][2 * V];
  __shared__ float par[2 * V];
  const int c0 = blockIdx.x * V;
  const uint32_t g = blockIdx.y;                        // normalisation group (see k_bn_small_fwd)
  const size_t gi = (size_t)g * C;
  const size_t p_base = (size_t)g * P;
  const TZ* z = z_all + p_base * C;
  T* dz = dz_all + p_base * C;
  const bool has_d2 = d2.p != nullptr;
  for (uint32_t p = threadIdx.x; p < P; p += BNS_THREADS) {
    cp_async16(&slab[p], z + (size_t)p * C + c0, true);
    cp_async16(&slab[P + p], (const T*)d1.p + (p_base + p) * d1.pitch + d1.coff + c0, true);
    if (has_d2) cp_async16(&slab[2 * P + p], (const T*)d2.p + (p_base + p) * d2.pitch + d2.coff + c0, true);
  }
  cp_async_commit();
  float mu[V], sc[V], sf[V], iv[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    mu[k] = __ldg(mean + gi + c0 + k); sc[k] = __ldg(scale + gi + c0 + k); sf[k] = __ldg(shift + gi + c0 + k);
    iv[k] = __ldg(inv + gi + c0 + k);
  }
  const uint32_t call = DROP ? __ldg(dk.call_dev) + dk.call_off : 0u;
  cp_async_wait<0>();
  // g = (d1 + d2) * act'(u) * dropout'
  auto grad_at = [&](uint32_t p, float (&xc)[V], float (&gg)[V]) {
    float v[V], gr[V];
    unpack16(slab[p], v, (const TZ*)nullptr);
    unpack16(slab[P + p], gr, (const T*)nullptr);
    if (has_d2) {
      float h[V];
      unpack16(slab[2 * P + p], h, (const T*)nullptr);
#pragma unroll
      for (int k = 0; k < V; ++k) gr[k] += h[k];
    }
    uint32_t smp = 0, e0 = 0;
    if (DROP) { const uint32_t pgl = g * P + p; smp = pgl / HW; e0 = (pgl - smp * HW) * C + c0; }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      xc[k] = v[k] - mu[k];
      float uu = fmaf(xc[k], sc[k], sf[k]);
      float g = gr[k] * act_bwd(uu, act);
      if (DROP) g = dropout_keep(dk, call, smp, e0 + k) ? 2.f * g : 0.f;
      gg[k] = g;
    }
  };
  float s[V], q[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { s[k] = 0.f; q[k] = 0.f; }
  for (uint32_t p = threadIdx.x; p < P; p += BNS_THREADS) {
    float xc[V], gg[V];
    grad_at(p, xc, gg);
#pragma unroll
    for (int k = 0; k < V; ++k) { s[k] += gg[k]; q[k] = fmaf(gg[k], xc[k], q[k]); }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) q[k] *= iv[k];
  double S, Q;
  bns_block_sums<V>(s, q, red, S, Q);
  if (threadIdx.x < V) {                                // same arithmetic as k_bwd_finalize
    const int c = c0 + threadIdx.x;
    const double n = (double)P;
    float a = (float)(S / n), b = (float)(Q / n);
    c1[gi + c] = a; c2[gi + c] = b;
    par[threadIdx.x] = a; par[V + threadIdx.x] = b;
    if (gridDim.y == 1) { dbeta[c] += (float)S; dgamma[c] += (float)Q; }     // one group: only writer (see k_bwd_finalize)
  }
  __syncthreads();
  float k1[V], k2[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { k1[k] = par[k]; k2[k] = par[V + k]; }
  for (uint32_t p = threadIdx.x; p < P; p += BNS_THREADS) {
    float xc[V], gg[V], o[V];
    grad_at(p, xc, gg);
#pragma unroll
    for (int k = 0; k < V; ++k) o[k] = sc[k] * (gg[k] - k1[k] - xc[k] * iv[k] * k2[k]);
    VecIO<T>::store(dz + (size_t)p * C + c0, o);
  }
  if (gridDim.y > 1) {
    // InstanceNorm: the last group-block of this channel slab sums the groups in index order (deterministic)
    if (!last_block_arrives(counters + blockIdx.x, gridDim.y)) return;
    if (threadIdx.x < V) {
      const int c = c0 + threadIdx.x;
      float sb = 0.f, sg = 0.f;
      for (unsigned int gg = 0; gg < gridDim.y; ++gg) { sb += __ldcg(c1 + (size_t)gg * C + c) * (float)P; sg += __ldcg(c2 + (size_t)gg * C + c) * (float)P; }
      dbeta[c] += sb; dgamma[c] += sg;
    }
  }
}

static bool g_bn_small = [] { const char* e = getenv("GAN_B200_BN_SMALL"); return !(e && e[0] == '0'); }();   // dev A/B switch
void set_bn_small(bool on) { g_bn_small = on; }
static inline bool bn_small_fits(int G, int64_t P, int narr) {     // P = all pixels, G groups of P / G
  return g_bn_small && G >= 1 && G <= 65535 && P >= G && (size_t)(P / G) * narr * 16 <= (size_t)BNS_MAX_SMEM;
}

bool bn_small_fwd_fits(int G, int64_t P) { return bn_small_fits(G, P, 1); }

bool launch_bn_small_fwd(Launch L, int dt, const void* z, int64_t P, int G, int HW, int C, float eps, const float* gamma,
                         const float* beta, float* mean, float* inv, float* scale, float* shift, float* mov_mean,
                         float* mov_var, float momentum, int act, DropKey dk, void* out, int out_pitch, int out_coff) {
  if (!bn_small_fits(G, P, 1)) return false;
  const int64_t Pg = P / G;
  dispatch_dt(dt, [&](auto* tag) {
    using T = typename std::remove_pointer<decltype(tag)>::type;
    constexpr int V = VecIO<T>::N;
    GAN_REQUIRE(C % V == 0, "channel count must be a multiple of the vector width");
    static bool once = (set_smem(k_bn_small_fwd<T, true>, BNS_MAX_SMEM), set_smem(k_bn_small_fwd<T, false>, BNS_MAX_SMEM), true);
    (void)once;
    auto kern = dk.enabled ? k_bn_small_fwd<T, true> : k_bn_small_fwd<T, false>;
    kern<<<dim3(C / V, G), BNS_THREADS, (size_t)Pg * 16, L.s>>>((const T*)z, (uint32_t)Pg, (uint32_t)HW, C, eps, gamma, beta, mean, inv,
                                                      scale, shift, mov_mean, mov_var, momentum, act, dk, (T*)out, out_pitch,
                                                      out_coff);
  });
  KLAUNCH(L);
  return true;
}

void launch_norm_bwd(Launch L, int dtz, int dt, const void* z, GradSrc d1, GradSrc d2, int64_t P, int64_t Pg, int G, int HW,
                     int C, int norm, const float* mean, const float* inv, const float* scale, const float* shift,
                     int act, DropKey dk, float* ws, float* c1, float* c2, float* dgamma, float* dbeta, void* dz,
                     unsigned int* counters, int z_pitch, int z_coff) {
  int nchunk = stats_chunks(G, Pg);
  if (z_pitch <= 0) { z_pitch = C; z_coff = 0; }
  GAN_REQUIRE(norm == NORM_NONE || (z_pitch == C && z_coff == 0), "a strided z view is only supported for layers without normalisation");
  dispatch_dt2(dtz, dt, [&](auto* ztag, auto* tag) {
    using TZ = typename std::remove_pointer<decltype(ztag)>::type;
    using T = typename std::remove_pointer<decltype(tag)>::type;
    static_assert(sizeof(TZ) == sizeof(T), "activation and gradient vectors must have the same width");
    const int cv = C / VecIO<T>::N;
    GAN_REQUIRE((cv & (cv - 1)) == 0 && cv <= 256 && cv >= 1, "channel count must be a power of two");
    const int lcv = ilog2(cv);
    const int narr = d2.p != nullptr ? 3 : 2;
    const size_t smem = (size_t)RING_STAGES * narr * 256 * 16;
    const size_t smem_max = (size_t)RING_STAGES * 3 * 256 * 16;
    static bool once = (set_smem(k_bwd_reduce<TZ, T, true>, smem_max), set_smem(k_bwd_reduce<TZ, T, false>, smem_max),
                        set_smem(k_bwd_apply<TZ, T, true>, smem_max), set_smem(k_bwd_apply<TZ, T, false>, smem_max), true);
    (void)once;
    if (norm != NORM_NONE && bn_small_fits(G, P, narr)) {
      static bool once2 = (set_smem(k_bn_small_bwd<TZ, T, true>, BNS_MAX_SMEM), set_smem(k_bn_small_bwd<TZ, T, false>, BNS_MAX_SMEM), true);
      (void)once2;
      auto kern = dk.enabled ? k_bn_small_bwd<TZ, T, true> : k_bn_small_bwd<TZ, T, false>;
      kern<<<dim3(C / VecIO<T>::N, G), BNS_THREADS, (size_t)Pg * narr * 16, L.s>>>((const TZ*)z, d1, d2, (uint32_t)Pg, (uint32_t)HW, C, mean,
                                                                         inv, scale, shift, act, dk, c1, c2, dgamma, dbeta, (T*)dz, counters);
      KLAUNCH(L);
      return;
    }
    if (norm != NORM_NONE) {
      auto kred = dk.enabled ? k_bwd_reduce<TZ, T, true> : k_bwd_reduce<TZ, T, false>;
      kred<<<dim3(nchunk, G), 256, smem, L.s>>>((const TZ*)z, d1, d2, (uint32_t)Pg, (uint32_t)HW, C, lcv, nchunk, mean, inv,
                                                scale, shift, act, dk, ws);
      KLAUNCH(L);
      k_bwd_finalize<<<dim3((C + 31) / 32, G), 32 * FIN_LANES, 0, L.s>>>(ws, G, nchunk, C, (double)Pg, c1, c2, dgamma, dbeta, counters);
      KLAUNCH(L);
    }
    auto kapp = dk.enabled ? k_bwd_apply<TZ, T, true> : k_bwd_apply<TZ, T, false>;
    kapp<<<norm_grid(P, cv, narr == 3 ? 2 : 3), 256, smem, L.s>>>((const TZ*)z, d1, d2, (uint32_t)P, (uint32_t)Pg, G, (uint32_t)HW,
                                                                 C, lcv, norm, mean, inv, scale, shift, c1, c2, act, dk, (T*)dz, z_pitch, z_coff);
    KLAUNCH(L);
  });
}

// ---------------------------------------------------------------------------------------------
// Generator head backward (tanh output; L1 term pix2pix.py:181 / cycle_gan.py:167,176).
// ---------------------------------------------------------------------------------------------
// dbias[k] += sum over blocks (in index order) of part[block][k], done by the last block to arrive: deterministic
__device__ __forceinline__ void head_bias_finish(const float* part, int C, float* dbias, unsigned int* counter) {
  if (!last_block_arrives(counter, gridDim.x)) return;
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (k >= C || k >= 4) return;
  double s = 0.0;
  for (unsigned int b = lane; b < gridDim.x; b += 32) s += (double)__ldcg(part + (size_t)b * 4 + k);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) dbias[k] += (float)s;
}
template <typename T>
__global__ void __launch_bounds__(256) k_ghead_bwd(const float* __restrict__ out, const float* __restrict__ ref, GradSrc d1,
                                                   GradSrc d2, float l1_coef, int64_t total, int C, T* __restrict__ dz,
                                                   int dz_pitch, float* dbias, float* part, unsigned int* counter) {
  __shared__ float sh[8];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};   // C <= 4
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / C; int c = (int)(i - p * C);
    float o = out[i];
    float d = 0.f;
    if (d1.p != nullptr) d += to_f(((const T*)d1.p)[p * d1.pitch + d1.coff + c]);
    if (d2.p != nullptr) d += to_f(((const T*)d2.p)[p * d2.pitch + d2.coff + c]);
    if (ref != nullptr) {
      float df = o - ref[i];
      d += l1_coef * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f));
    }
    float g = d * (1.f - o * o);
    T gq = from_f<T>(g);
    dz[p * dz_pitch + c] = gq;
    float gf = to_f(gq);
#pragma unroll
    for (int k = 0; k < 4; ++k) if (c == k) bsum[k] += gf;
  }
  for (int k = 0; k < C && k < 4; ++k) {
    float v = warp_sum(bsum[k]);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
      part[(size_t)blockIdx.x * 4 + k] = t;              // per-block partial
    }
  }
  head_bias_finish(part, C, dbias, counter);
}
void launch_ghead_bwd(Launch L, int dt, const float* out_f32, const float* ref_f32, GradSrc d1, GradSrc d2,
                      float l1_coef, int64_t P, int C, void* dz, int dz_pitch, float* dbias, float* part, unsigned int* counters) {
  GAN_REQUIRE(C <= 4, "generator head supports up to 4 output channels");
  int64_t total = P * C;
  dispatch_dt(dt, [&](auto* tag) {
    using T = typename std::remove_pointer<decltype(tag)>::type;
    const int grid = grid_for(total, 256, 4);
    GAN_REQUIRE(grid <= HEAD_PART_BLOCKS, "bias partial workspace too small");
    k_ghead_bwd<T><<<grid, 256, 0, L.s>>>(out_f32, ref_f32, d1, d2, l1_coef, total, C, (T*)dz, dz_pitch, dbias, part, counters);
  });
  KLAUNCH(L);
}

// ---------------------------------------------------------------------------------------------
// Losses.  Each loss kernel writes one fp32 partial per block into loss_ws[slot][blockIdx.x];
// k_loss_finalize sums the partials in double and mixes them into the reported scalars.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_partial_store(float v, float* dst) {
  __shared__ float sh[8];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    *dst = t;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_bce(const float* __restrict__ x, int64_t n, float label, float coef_over_n,
                                             T* dz, int dz_pitch, float* dbias, float* part, unsigned int* counter,
                                             float* __restrict__ loss_slot) {
  float acc = 0.f, bacc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x[i];
    acc += fmaxf(v, 0.f) - v * label + log1pf(expf(-fabsf(v)));
    if (dz != nullptr) {
      float sg = 1.f / (1.f + expf(-v));
      T q = from_f<T>(coef_over_n * (sg - label));
      dz[i * dz_pitch] = q;
      bacc += to_f(q);
    }
  }
  block_partial_store(acc, loss_slot + blockIdx.x);
  if (dz != nullptr && dbias != nullptr) {
    // per-block partial of the bias gradient; the last block adds them in block order (deterministic)
    __syncthreads();
    __shared__ float shb[8];
    float v = warp_sum(bacc);
    if ((threadIdx.x & 31) == 0) shb[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += shb[w];
      part[blockIdx.x] = t;
    }
    if (!last_block_arrives(counter, gridDim.x)) return;
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (unsigned int b = 0; b < gridDim.x; ++b) t += __ldcg(part + b);
      *dbias += t;
    }
  }
}
void launch_bce(Launch L, int dt, const float* logits, int64_t n, float label, float coef, void* dz, int dz_pitch,
                float* dbias, float* loss_ws, int slot, float* part_ws, unsigned int* counters) {
  int blocks = grid_for(n, 256, 1);
  if (blocks > LOSS_BLOCKS) blocks = LOSS_BLOCKS;
  const bool bias = dz != nullptr && dbias != nullptr;
  dispatch_dt(dt, [&](auto* tag) {
    using T = typename std::remove_pointer<decltype(tag)>::type;
    k_bce<T><<<blocks, 256, 0, L.s>>>(logits, n, label, coef / (float)n, (T*)dz, dz_pitch, bias ? dbias : nullptr, part_ws, counters,
                                      loss_ws + slot * LOSS_BLOCKS);
  });
  KLAUNCH(L);
}

__global__ void __launch_bounds__(256) k_l1(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                            float* __restrict__ loss_slot) {
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += fabsf(a[i] - b[i]);
  block_partial_store(acc, loss_slot + blockIdx.x);
}
void launch_l1(Launch L, const float* a, const float* b, int64_t n, float* loss_ws, int slot) {
  int blocks = grid_for(n, 256 * 8, 2);
  if (blocks > LOSS_BLOCKS) blocks = LOSS_BLOCKS;
  k_l1<<<blocks, 256, 0, L.s>>>(a, b, n, loss_ws + slot * LOSS_BLOCKS);
  KLAUNCH(L);
}

__global__ void k_loss_finalize(const float* __restrict__ ws, LossMix mix, float* __restrict__ out) {
  __shared__ double raw[LOSS_SLOTS];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (w < mix.nraw) {
    double s = 0.0;
    for (int i = lane; i < LOSS_BLOCKS; i += 32) s += (double)ws[w * LOSS_BLOCKS + i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) raw[w] = s / (double)mix.denom[w];
  }
  __syncthreads();
  if (threadIdx.x < mix.nout) {
    double v = 0.0;
    for (int j = 0; j < mix.nraw; ++j) v += (double)mix.mix[threadIdx.x * mix.nraw + j] * raw[j];
    out[threadIdx.x] = (float)v;
  }
}
void launch_loss_finalize(Launch L, const float* loss_ws, LossMix mix, float* out) {
  k_loss_finalize<<<1, 32 * LOSS_SLOTS, 0, L.s>>>(loss_ws, mix, out);
  KLAUNCH(L);
}

// ---------------------------------------------------------------------------------------------
// Keras Adam: m += (g-m)(1-b1); v += (g^2-v)(1-b2); p -= lr_t*m/(sqrt(v)+eps)
// One launch per network over the flat parameter buffer: 28 B/param of HBM traffic.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, int64_t n4, int64_t n, const long long* __restrict__ t_dev,
                                              double lr, double b1d, double b2d, float eps, float gscale) {
  // lr_t = lr*sqrt(1-b2^t)/(1-b1^t) from the device-resident step counter (graph-replay safe)
  const double t = (double)(*t_dev);
  const float lr_t = (float)(lr * sqrt(1.0 - pow(b2d, t)) / (1.0 - pow(b1d, t)));
  const float b1 = (float)b1d, b2 = (float)b2d;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 P = reinterpret_cast<float4*>(p)[i], Gr = reinterpret_cast<const float4*>(g)[i];
    float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    float* pp = &P.x; float* gg = &Gr.x; float* mm = &M.x; float* vv = &V.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr = gg[k] * gscale;
      mm[k] += (gr - mm[k]) * (1.f - b1);
      vv[k] += (gr * gr - vv[k]) * (1.f - b2);
      pp[k] -= lr_t * mm[k] / (sqrtf(vv[k]) + eps);
    }
    reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(m)[i] = M; reinterpret_cast<float4*>(v)[i] = V;
  }
  // tail (n not a multiple of 4)
  int64_t tl = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (tl < n) {
    float gr = g[tl] * gscale;
    m[tl] += (gr - m[tl]) * (1.f - b1);
    v[tl] += (gr * gr - v[tl]) * (1.f - b2);
    p[tl] -= lr_t * m[tl] / (sqrtf(v[tl]) + eps);
  }
}
__global__ void k_bump(long long* t64, uint32_t* c32, uint32_t by) {
  if (t64) *t64 += (long long)by;
  if (c32) *c32 += by;
}
void launch_bump(Launch L, long long* t64, uint32_t* c32, uint32_t by) {
  k_bump<<<1, 1, 0, L.s>>>(t64, c32, by);
  KLAUNCH(L);
}
void launch_adam(Launch L, float* p, const float* g, float* m, float* v, int64_t n, const long long* t_dev, double lr,
                 double b1, double b2, float eps, float gscale) {
  int64_t n4 = n / 4;
  k_adam<<<grid_for(n4 > 0 ? n4 : 1, 256, 8), 256, 0, L.s>>>(p, g, m, v, n4, n, t_dev, lr, b1, b2, eps, gscale);
  KLAUNCH(L);
}

__global__ void k_scale(float* p, int64_t n, float s) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] *= s;
}
void launch_scale(Launch L, float* p, int64_t n, float s) {
  k_scale<<<grid_for(n, 256), 256, 0, L.s>>>(p, n, s);
  KLAUNCH(L);
}

// ---------------------------------------------------------------------------------------------
// Weight packing: fp32 master (TF layout) -> [class][Nc][tap*Kc + kc] in the activation dtype.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_pack(const float* __restrict__ master, T* __restrict__ dst, PackOp op) {
  const ClassGeom& cg = op.cls[blockIdx.y];
  const int64_t K = (int64_t)cg.ntaps * op.Kc;
  const int64_t total = K * op.Nc;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = i / K; int64_t k = i - n * K;
    int t = (int)(k / op.Kc); int kc = (int)(k - (int64_t)t * op.Kc);
    float v = (kc < op.Kr && n < op.Nr) ? master[(int64_t)cg.widx[t] * op.s_tap + (int64_t)kc * op.s_k + n * op.s_n] : 0.f;
    dst[cg.b_off + i] = from_f<T>(v);
  }
}
void launch_pack(Launch L, int dt, const float* master, void* dst, const PackOp& op) {
  int64_t total = (int64_t)op.cls[0].ntaps * op.Kc * op.Nc;
  dispatch_dt(dt, [&](auto* tag) {
    using T = typename std::remove_pointer<decltype(tag)>::type;
    k_pack<T><<<dim3(grid_for(total, 256, 4), op.ncls), 256, 0, L.s>>>(master, (T*)dst, op);
  });
  KLAUNCH(L);
}

// 16-bit destinations carry their own format per entry (forward packs: activation dtype, data-gradient packs: bf16)
template <typename T> __device__ __forceinline__ T pack_cvt(float v, int dt16);
template <> __device__ __forceinline__ float pack_cvt<float>(float v, int) { return v; }
template <> __device__ __forceinline__ bf16 pack_cvt<bf16>(float v, int dt16) {
  if (dt16 == DT_F16) { f16 h = f16_sat(v); return *reinterpret_cast<bf16*>(&h); }    // bit pattern of the fp16 value
  return __float2bfloat16_rn(v);
}
template <typename T>
__global__ void __launch_bounds__(256) k_pack_multi(const PackEntry* __restrict__ tab, int nent) {
  __shared__ float tile[32][33];
  __shared__ int s_e;
  if (threadIdx.x == 0) {
    int e = 0;
    while (e + 1 < nent && (int)blockIdx.x >= tab[e + 1].tile_begin) ++e;
    s_e = e;
  }
  __syncthreads();
  const PackEntry& E = tab[s_e];
  const PackOp& op = E.op;
  int lt = blockIdx.x - E.tile_begin;
  const int tn = lt % E.tiles_n; lt /= E.tiles_n;
  const int tk = lt % E.tiles_k; lt /= E.tiles_k;
  const int ntaps = op.cls[0].ntaps;
  const int t = lt % ntaps, ci = lt / ntaps;
  const ClassGeom& cg = op.cls[ci];
  const float* __restrict__ src = E.master + (int64_t)cg.widx[t] * op.s_tap;
  T* __restrict__ dst = (T*)E.dst + cg.b_off;
  const int64_t Ktot = (int64_t)ntaps * op.Kc;
  const int k0 = tk * 32, n0 = tn * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  if (op.s_n == 1) {
    // master contiguous along n: read rows = kc, cols = n; write rows = n, cols = kc (transpose)
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      int kc = k0 + r, n = n0 + tx;
      tile[r][tx] = (kc < op.Kr && n < op.Nr) ? src[(int64_t)kc * op.s_k + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      int n = n0 + r, kc = k0 + tx;
      if (n < op.Nc && kc < op.Kc) dst[(int64_t)n * Ktot + (int64_t)t * op.Kc + kc] = pack_cvt<T>(tile[tx][r], E.dt16);
    }
  } else {
    // master contiguous along kc (s_k == 1): straight tile copy
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      int n = n0 + r, kc = k0 + tx;
      if (n < op.Nc && kc < op.Kc) {
        float v = (kc < op.Kr && n < op.Nr) ? src[(int64_t)kc * op.s_k + (int64_t)n * op.s_n] : 0.f;
        dst[(int64_t)n * Ktot + (int64_t)t * op.Kc + kc] = pack_cvt<T>(v, E.dt16);
      }
    }
  }
}
void launch_pack_multi(Launch L, int dt, const PackEntry* tab_dev, int nent, int total_tiles) {
  // dt: DT_F32 or "16-bit" (the format of each destination is PackEntry::dt16)
  if (dt == DT_F32) k_pack_multi<float><<<total_tiles, 256, 0, L.s>>>(tab_dev, nent);
  else k_pack_multi<bf16><<<total_tiles, 256, 0, L.s>>>(tab_dev, nent);
  KLAUNCH(L);
}

// ---------------------------------------------------------------------------------------------
// Fused Keras-Adam + packing (see kernels.h)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float adam_lr_t(const AdamArgs& a) {
  const double t = (double)(*a.t_dev);
  return (float)(a.lr * sqrt(1.0 - pow(a.b2, t)) / (1.0 - pow(a.b1, t)));
}
// gradient loads: fp32 buffer, or the bf16 communication buffer the data-parallel all-reduce ran on
__device__ __forceinline__ float adam_g1(const AdamArgs& a, long long i) {
  return a.g16 != nullptr ? __uint_as_float((uint32_t)a.g16[i] << 16) : a.g[i];
}
__device__ __forceinline__ float4 adam_g4(const AdamArgs& a, long long i) {      // i % 4 == 0
  if (a.g16 == nullptr) return *reinterpret_cast<const float4*>(a.g + i);
  const uint2 t = *reinterpret_cast<const uint2*>(a.g16 + i);
  return make_float4(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xffff0000u), __uint_as_float(t.y << 16),
                     __uint_as_float(t.y & 0xffff0000u));
}
__global__ void __launch_bounds__(256) k_grad_to_bf16(const float* __restrict__ g, uint16_t* __restrict__ o, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(g) + i);
    uint2 w; w.x = pack2<bf16>(v.x, v.y); w.y = pack2<bf16>(v.z, v.w);
    reinterpret_cast<uint2*>(o)[i] = w;
  }
}
void launch_grad_to_bf16(Launch L, const float* g, uint16_t* g16, int64_t n) {
  GAN_REQUIRE(n % 4 == 0 && ((uintptr_t)g & 15) == 0 && ((uintptr_t)g16 & 7) == 0, "gradient range must be 4-aligned");
  if (n <= 0) return;
  k_grad_to_bf16<<<grid_for(n / 4, 256, 8), 256, 0, L.s>>>(g, g16, n / 4);
  KLAUNCH(L);
}

__device__ __forceinline__ float adam_update(const AdamArgs& a, long long i, float lr_t, float b1, float b2) {
  float gr = adam_g1(a, i) * a.gscale, mm = a.m[i], vv = a.v[i], pp = a.p[i];
  mm += (gr - mm) * (1.f - b1);
  vv += (gr * gr - vv) * (1.f - b2);
  pp -= lr_t * mm / (sqrtf(vv) + a.eps);
  a.m[i] = mm; a.v[i] = vv; a.p[i] = pp;
  return pp;
}

// 64x64 tiles of the master [16][A][B] kernel tensors, 256 threads: thread = (column b, 4 row groups);
// the 4 x 8 rows of a half-tile are loaded first (32 independent loads per thread), then updated.
template <> __device__ __forceinline__ uint32_t pack2<float>(float, float) { return 0u; }   // fp32 mode never packs pairs
#define APT 64
template <typename TF, typename TD>
__global__ void __launch_bounds__(256) k_adam_pack(AdamArgs a, const AdamPackEntry* __restrict__ tab, int nent) {
  __shared__ float tile[APT][APT + 1];
  __shared__ int s_e;
  __shared__ float s_lr;
  if (threadIdx.x == 0) {
    int e = 0;
    while (e + 1 < nent && (int)blockIdx.x >= tab[e + 1].tile_begin) ++e;
    s_e = e;
    s_lr = adam_lr_t(a);
  }
  __syncthreads();
  const AdamPackEntry& E = tab[s_e];
  const float lr_t = s_lr, b1 = (float)a.b1, b2 = (float)a.b2;
  int lt = blockIdx.x - E.tile_begin;
  const int tb = lt % E.tiles_b; lt /= E.tiles_b;
  const int ta = lt % E.tiles_a; const int widx = lt / E.tiles_a;
  const int a0 = ta * APT, b0 = tb * APT;
  const int tx = threadIdx.x & (APT - 1), ty = threadIdx.x >> 6;          // 64 x 4
  const int cF = E.invF[widx] >> 4, tF = E.invF[widx] & 15, cD = E.invD[widx] >> 4, tD = E.invD[widx] & 15;
  TF* __restrict__ dF = (TF*)E.dstF + E.boffF[cF] + (long long)tF * E.KcF;     // + co*KtotF + ci
  TD* __restrict__ dD = (TD*)E.dstD + E.boffD[cD] + (long long)tD * E.KcD;     // + ci*KtotD + co
  const long long base = E.w_off + (long long)widx * E.A * E.B;
  if (E.vec) {
    // full 64x64 tile, 16-byte aligned rows: thread = (4 consecutive b, 4 rows 16 apart); 16 independent 128-bit loads
    // per thread are in flight before the first dependent instruction
    const int c4 = threadIdx.x & 15, rg = threadIdx.x >> 4;
    const int b4 = b0 + 4 * c4;
    float4 P[4], G[4], M[4], V[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long idx = base + (long long)(a0 + rg + 16 * i) * E.B + b4;
      P[i] = *reinterpret_cast<const float4*>(a.p + idx); G[i] = adam_g4(a, idx);
      M[i] = *reinterpret_cast<const float4*>(a.m + idx); V[i] = *reinterpret_cast<const float4*>(a.v + idx);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = rg + 16 * i, ai = a0 + r;
      const long long idx = base + (long long)ai * E.B + b4;
      float* pp = &P[i].x; const float* gg = &G[i].x; float* mm = &M[i].x; float* vv = &V[i].x;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float gr = gg[e] * a.gscale;
        mm[e] += (gr - mm[e]) * (1.f - b1);
        vv[e] += (gr * gr - vv[e]) * (1.f - b2);
        pp[e] -= lr_t * mm[e] / (sqrtf(vv[e]) + a.eps);
        tile[r][4 * c4 + e] = pp[e];
      }
      *reinterpret_cast<float4*>(a.p + idx) = P[i]; *reinterpret_cast<float4*>(a.m + idx) = M[i];
      *reinterpret_cast<float4*>(a.v + idx) = V[i];
      // destination that is contiguous along the master's fast index b: 4 values = one 8-byte store
      if (E.conv2d) {
        uint2 o; o.x = pack2<TD>(pp[0], pp[1]); o.y = pack2<TD>(pp[2], pp[3]);
        *reinterpret_cast<uint2*>(dD + (long long)ai * E.KtotD + b4) = o;             // ci = a, co = b
      } else {
        uint2 o; o.x = pack2<TF>(pp[0], pp[1]); o.y = pack2<TF>(pp[2], pp[3]);
        *reinterpret_cast<uint2*>(dF + (long long)ai * E.KtotF + b4) = o;             // co = a, ci = b
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int bj = b0 + rg + 16 * i, a4 = a0 + 4 * c4;     // row of the transposed tile = b index, 4 consecutive a
      const float v0 = tile[4 * c4][rg + 16 * i], v1 = tile[4 * c4 + 1][rg + 16 * i], v2 = tile[4 * c4 + 2][rg + 16 * i],
                  v3 = tile[4 * c4 + 3][rg + 16 * i];
      if (E.conv2d) {
        uint2 o; o.x = pack2<TF>(v0, v1); o.y = pack2<TF>(v2, v3);
        *reinterpret_cast<uint2*>(dF + (long long)bj * E.KtotF + a4) = o;             // co = b, ci = a
      } else {
        uint2 o; o.x = pack2<TD>(v0, v1); o.y = pack2<TD>(v2, v3);
        *reinterpret_cast<uint2*>(dD + (long long)bj * E.KtotD + a4) = o;             // ci = b, co = a
      }
    }
    return;
  }
  const int bi = b0 + tx;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float P[8], G[8], M[8], V[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ai = a0 + half * 32 + ty + 4 * i;
      if (ai < E.A && bi < E.B) {
        const long long idx = base + (long long)ai * E.B + bi;
        P[i] = a.p[idx]; G[i] = adam_g1(a, idx); M[i] = a.m[idx]; V[i] = a.v[idx];
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = half * 32 + ty + 4 * i, ai = a0 + r;
      float pn = 0.f;
      if (ai < E.A && bi < E.B) {
        const long long idx = base + (long long)ai * E.B + bi;
        float gr = G[i] * a.gscale, mm = M[i], vv = V[i];
        mm += (gr - mm) * (1.f - b1);
        vv += (gr * gr - vv) * (1.f - b2);
        pn = P[i] - lr_t * mm / (sqrtf(vv) + a.eps);
        a.m[idx] = mm; a.v[idx] = vv; a.p[idx] = pn;
        // destination that is contiguous along the master's fast index b
        if (E.conv2d) dD[(long long)ai * E.KtotD + bi] = from_f<TD>(pn);       // ci = a, co = b
        else dF[(long long)ai * E.KtotF + bi] = from_f<TF>(pn);                // co = a, ci = b
      }
      tile[r][tx] = pn;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int r = ty + 4 * i;                 // row of the transposed tile = b index
    const int bj = b0 + r, ai = a0 + tx;
    if (ai < E.A && bj < E.B) {
      if (E.conv2d) dF[(long long)bj * E.KtotF + ai] = from_f<TF>(tile[tx][r]);  // co = b, ci = a
      else dD[(long long)bj * E.KtotD + ai] = from_f<TD>(tile[tx][r]);           // ci = b, co = a
    }
  }
}
void launch_adam_pack(Launch L, int dt_fwd, int dt_dgrad, const AdamArgs& a, const AdamPackEntry* tab_dev, int nent, int total_tiles) {
  dispatch_dt2(dt_fwd, dt_dgrad, [&](auto* ftag, auto* dtag) {
    using TF = typename std::remove_pointer<decltype(ftag)>::type;
    using TD = typename std::remove_pointer<decltype(dtag)>::type;
    k_adam_pack<TF, TD><<<total_tiles, 256, 0, L.s>>>(a, tab_dev, nent);
  });
  KLAUNCH(L);
}

__global__ void __launch_bounds__(256) k_adam_ranges(AdamArgs a, const AdamRange* __restrict__ tab) {
  const AdamRange r = tab[blockIdx.x];
  const float lr_t = adam_lr_t(a), b1 = (float)a.b1, b2 = (float)a.b2;
  for (int i = threadIdx.x; i < r.n; i += 256) adam_update(a, r.off + i, lr_t, b1, b2);
}
void launch_adam_ranges(Launch L, const AdamArgs& a, const AdamRange* tab_dev, int nranges) {
  k_adam_ranges<<<nranges, 256, 0, L.s>>>(a, tab_dev);
  KLAUNCH(L);
}

__global__ void __launch_bounds__(256) k_zero_ranges(float* __restrict__ g, const AdamRange* __restrict__ tab) {
  const AdamRange r = tab[blockIdx.x];
  for (int i = threadIdx.x; i < r.n; i += 256) g[r.off + i] = 0.f;
}
void launch_zero_ranges(Launch L, float* g, const AdamRange* tab_dev, int nranges) {
  if (nranges <= 0) return;
  k_zero_ranges<<<nranges, 256, 0, L.s>>>(g, tab_dev);
  KLAUNCH(L);
}

template <typename T>
__global__ void __launch_bounds__(256) k_sum_slabs(const float* __restrict__ slabs, int nslab, int64_t total, int C,
                                                   T* __restrict__ dst, int pitch, int coff) {
  const int64_t tot4 = total / 4;                 // C % 4 == 0
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < tot4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<const float4*>(slabs)[i];
    for (int k = 1; k < nslab; ++k) {
      float4 b = reinterpret_cast<const float4*>(slabs + (int64_t)k * total)[i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    const int64_t e = i * 4; const int64_t p = e / C; const int c = (int)(e - p * C);
    T* o = dst + p * pitch + coff + c;
    o[0] = from_f<T>(a.x); o[1] = from_f<T>(a.y); o[2] = from_f<T>(a.z); o[3] = from_f<T>(a.w);
  }
}
void launch_sum_slabs(Launch L, int dt, const float* slabs, int nslab, int64_t P, int C, void* dst, int pitch, int coff) {
  const int64_t total = P * C;
  dispatch_dt(dt, [&](auto* tag) {
    using T = typename std::remove_pointer<decltype(tag)>::type;
    k_sum_slabs<T><<<grid_for(total / 4, 256, 4), 256, 0, L.s>>>(slabs, nslab, total, C, (T*)dst, pitch, coff);
  });
  KLAUNCH(L);
}

// ---------------------------------------------------------------------------------------------
// Small special weight layouts (first-layer im2col order, head cols operands): dst[i] = master[idx[i]]
// through a device index table built once on the host (idx < 0 -> 0).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_gather_pack(const float* __restrict__ master, const int* __restrict__ idx, int n, T* __restrict__ dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int j = idx[i];
    dst[i] = from_f<T>(j >= 0 ? master[j] : 0.f);
  }
}
void launch_gather_pack(Launch L, int dt, const float* master, const int* idx_dev, int n, void* dst) {
  dispatch_dt(dt, [&](auto* tag) {
    using T = typename std::remove_pointer<decltype(tag)>::type;
    k_gather_pack<T><<<grid_for(n, 256, 2), 256, 0, L.s>>>(master, idx_dev, n, (T*)dst);
  });
  KLAUNCH(L);
}

// ---------------------------------------------------------------------------------------------
// im2col of a 4x4 stride-2 'same' window over a 1..4-channel image (first layers base_gan.py:141,180;
// unfold of the generator-head gradient): one thread per output-grid point writes one 128-byte row
// [16 taps x 4 channel slots] (slots >= C are zero), entirely from registers.
// ---------------------------------------------------------------------------------------------
template <typename TS, int C, typename TO>
__global__ void __launch_bounds__(256) k_im2col(const TS* __restrict__ src, int pitch, int B, int H, int W,
                                                TO* __restrict__ dst) {
  // 8 threads per output-grid point, each builds one 16-byte chunk (2 taps x 4 channel slots): a warp
  // stores 4 consecutive 128-byte rows, fully coalesced.
  const int Ho = H / 2, Wo = W / 2;
  const int64_t total = (int64_t)B * Ho * Wo * 8;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = g >> 3; const int j = (int)(g & 7);
    const int ow = (int)(m % Wo); const int64_t r = m / Wo; const int oh = (int)(r % Ho); const int n = (int)(r / Ho);
    const int kh = j >> 1, kw0 = (j & 1) * 2;
    const int ih = 2 * oh + kh - 1;
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int iw = 2 * ow + kw0 + e - 1;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
        const TS* sp = src + (((int64_t)n * H + ih) * W + iw) * pitch;
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = to_f(sp[c]);
      }
      w[2 * e] = pack2<TO>(v[0], v[1]); w[2 * e + 1] = pack2<TO>(v[2], v[3]);
    }
    *reinterpret_cast<uint4*>(dst + m * 64 + j * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
// Generator-head backward fused with the unfold of its result (bf16 cols path): computes
// dz = (d1 + d2 + l1_coef*sign(out - ref)) * (1 - out^2) for the 2 taps of this thread's chunk straight from the
// fp32 output / target images and writes the slot-4 row chunk; dz itself never exists in HBM.  Every
// image pixel appears in 4 rows; the bias gradient counts it once, in the row where it is an inner tap.
template <int C, typename T>
__global__ void __launch_bounds__(256) k_ghead_bwd_cols(const float* __restrict__ out, const float* __restrict__ ref,
                                                        GradSrc d1, GradSrc d2, float l1_coef, int B, int H, int W,
                                                        T* __restrict__ dst, float* dbias, float* part, unsigned int* counter) {
  __shared__ float sh[8];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  const int Ho = H / 2, Wo = W / 2;
  const int64_t total = (int64_t)B * Ho * Wo * 8;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = g >> 3; const int j = (int)(g & 7);
    const int ow = (int)(m % Wo); const int64_t r = m / Wo; const int oh = (int)(r % Ho); const int n = (int)(r / Ho);
    const int kh = j >> 1, kw0 = (j & 1) * 2;
    const int ih = 2 * oh + kh - 1;
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int kw = kw0 + e, iw = 2 * ow + kw - 1;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
        const int64_t p = ((int64_t)n * H + ih) * W + iw;
        const bool owner = (kh == 1 || kh == 2) && (kw == 1 || kw == 2);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float o = out[p * C + c];
          float d = 0.f;
          if (d1.p != nullptr) d += to_f(((const T*)d1.p)[p * d1.pitch + d1.coff + c]);
          if (d2.p != nullptr) d += to_f(((const T*)d2.p)[p * d2.pitch + d2.coff + c]);
          if (ref != nullptr) {
            const float df = o - ref[p * C + c];
            d += l1_coef * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f));
          }
          v[c] = to_f(from_f<T>(d * (1.f - o * o)));
          if (owner) bsum[c] += v[c];
        }
      }
      w[2 * e] = pack2<T>(v[0], v[1]); w[2 * e + 1] = pack2<T>(v[2], v[3]);
    }
    *reinterpret_cast<uint4*>(dst + m * 64 + j * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
#pragma unroll
  for (int k = 0; k < C; ++k) {
    float v = warp_sum(bsum[k]);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int wi = 0; wi < (int)(blockDim.x >> 5); ++wi) t += sh[wi];
      part[(size_t)blockIdx.x * 4 + k] = t;              // per-block partial
    }
  }
  head_bias_finish(part, C, dbias, counter);
}
void launch_ghead_bwd_cols(Launch L, int dt, const float* out_f32, const float* ref_f32, GradSrc d1, GradSrc d2, float l1_coef, int B,
                           int H, int W, int C, void* gcols, float* dbias, float* part, unsigned int* counters) {
  GAN_REQUIRE(C >= 1 && C <= 4, "generator head supports up to 4 output channels");
  GAN_REQUIRE(dt == DT_F16 || dt == DT_BF16, "cols path is 16-bit only");
  const int64_t M = (int64_t)B * (H / 2) * (W / 2);
  const int grid = grid_for(M * 8, 256, 16);
  GAN_REQUIRE(grid <= HEAD_PART_BLOCKS, "bias partial workspace too small");
  auto run = [&](auto* tag) {
    using T = typename std::remove_pointer<decltype(tag)>::type;
    T* d = (T*)gcols;
    if (C == 1) k_ghead_bwd_cols<1, T><<<grid, 256, 0, L.s>>>(out_f32, ref_f32, d1, d2, l1_coef, B, H, W, d, dbias, part, counters);
    else if (C == 2) k_ghead_bwd_cols<2, T><<<grid, 256, 0, L.s>>>(out_f32, ref_f32, d1, d2, l1_coef, B, H, W, d, dbias, part, counters);
    else if (C == 3) k_ghead_bwd_cols<3, T><<<grid, 256, 0, L.s>>>(out_f32, ref_f32, d1, d2, l1_coef, B, H, W, d, dbias, part, counters);
    else k_ghead_bwd_cols<4, T><<<grid, 256, 0, L.s>>>(out_f32, ref_f32, d1, d2, l1_coef, B, H, W, d, dbias, part, counters);
  };
  if (dt == DT_F16) run((f16*)nullptr); else run((bf16*)nullptr);
  KLAUNCH(L);
}

template <typename TS, typename TO>
static void im2col_dispatch(Launch L, const TS* src, int pitch, int B, int H, int W, int C, void* dst) {
  GAN_REQUIRE(C >= 1 && C <= 4, "im2col supports 1..4 channels per source");
  const int64_t M = (int64_t)B * (H / 2) * (W / 2);
  const int grid = grid_for(M * 8, 256, 16);
  TO* d = (TO*)dst;
  if (C == 1) k_im2col<TS, 1, TO><<<grid, 256, 0, L.s>>>(src, pitch, B, H, W, d);
  else if (C == 2) k_im2col<TS, 2, TO><<<grid, 256, 0, L.s>>>(src, pitch, B, H, W, d);
  else if (C == 3) k_im2col<TS, 3, TO><<<grid, 256, 0, L.s>>>(src, pitch, B, H, W, d);
  else k_im2col<TS, 4, TO><<<grid, 256, 0, L.s>>>(src, pitch, B, H, W, d);
  KLAUNCH(L);
}
void launch_im2col(Launch L, int dt_rows, const float* src, int B, int H, int W, int C, void* dst_rows) {
  if (dt_rows == DT_F16) im2col_dispatch<float, f16>(L, src, C, B, H, W, C, dst_rows);
  else im2col_dispatch<float, bf16>(L, src, C, B, H, W, C, dst_rows);
}

// col2im of the transposed-conv head: one thread per output pixel gathers its 4 contributing taps
// (one aligned float4 each) from cols[m][tap*4 + co]; taps per output parity as geom_convT4.
__device__ __forceinline__ float4 ld_cols4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld_cols4(const f16* p) {
  const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 ld_cols4(const bf16* p) {
  const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
// The same gather as the data gradient of a 4x4 stride-2 'same' Conv2D whose per-tap products are in cols
// (discriminator first layer, dL/d(input image)): dx[ih, iw, c] = sum over the 4 taps with 2*o - 1 + k = i of
// cols[o][tap*4 + c]; output = compact 4-channel rows (8 bytes per pixel) in the gradient format.
template <typename T>
__global__ void __launch_bounds__(256) k_col2im_grad(const T* __restrict__ cols, int B, int Hin, int Win, T* __restrict__ out) {
  const int Ho = 2 * Hin, Wo = 2 * Win;
  const int64_t total = (int64_t)B * Ho * Wo;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int ow = (int)(q % Wo); const int64_t r = q / Wo; const int oh = (int)(r % Ho); const int n = (int)(r / Ho);
    const int a = oh & 1, b = ow & 1, i = oh >> 1, j = ow >> 1;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int th = 0; th < 2; ++th) {
      const int kh = a ? (th ? 2 : 0) : (th ? 3 : 1), dh = a ? (th ? 0 : 1) : (th ? -1 : 0);
      const int ih = i + dh;
      if (ih < 0 || ih >= Hin) continue;
#pragma unroll
      for (int tw = 0; tw < 2; ++tw) {
        const int kw = b ? (tw ? 2 : 0) : (tw ? 3 : 1), dw = b ? (tw ? 0 : 1) : (tw ? -1 : 0);
        const int iw = j + dw;
        if (iw < 0 || iw >= Win) continue;
        const float4 v = ld_cols4(cols + (((int64_t)n * Hin + ih) * Win + iw) * 64 + (kh * 4 + kw) * 4);
        acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
      }
    }
    *reinterpret_cast<uint2*>(out + q * 4) = make_uint2(pack2<T>(acc[0], acc[1]), pack2<T>(acc[2], acc[3]));
  }
}
void launch_col2im_grad(Launch L, int dt, const void* cols, int B, int Hin, int Win, void* out) {
  GAN_REQUIRE(dt == DT_F16 || dt == DT_BF16, "col2im_grad is 16-bit only");
  const int64_t total = (int64_t)B * Hin * Win * 4;
  if (dt == DT_F16) k_col2im_grad<f16><<<grid_for(total, 256, 16), 256, 0, L.s>>>((const f16*)cols, B, Hin, Win, (f16*)out);
  else k_col2im_grad<bf16><<<grid_for(total, 256, 16), 256, 0, L.s>>>((const bf16*)cols, B, Hin, Win, (bf16*)out);
  KLAUNCH(L);
}

template <typename TC>
__global__ void __launch_bounds__(256) k_col2im_tanh(const TC* __restrict__ cols, const float* __restrict__ bias, int B,
                                                     int Hin, int Win, int C, float* __restrict__ out) {
  const int Ho = 2 * Hin, Wo = 2 * Win;
  const int64_t total = (int64_t)B * Ho * Wo;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int ow = (int)(q % Wo); const int64_t r = q / Wo; const int oh = (int)(r % Ho); const int n = (int)(r / Ho);
    const int a = oh & 1, b = ow & 1, i = oh >> 1, j = ow >> 1;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int th = 0; th < 2; ++th) {
      const int kh = a ? (th ? 2 : 0) : (th ? 3 : 1), dh = a ? (th ? 0 : 1) : (th ? -1 : 0);
      const int ih = i + dh;
      if (ih < 0 || ih >= Hin) continue;
#pragma unroll
      for (int tw = 0; tw < 2; ++tw) {
        const int kw = b ? (tw ? 2 : 0) : (tw ? 3 : 1), dw = b ? (tw ? 0 : 1) : (tw ? -1 : 0);
        const int iw = j + dw;
        if (iw < 0 || iw >= Win) continue;
        const float4 v = ld_cols4(cols + (((int64_t)n * Hin + ih) * Win + iw) * 64 + (kh * 4 + kw) * 4);
        acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
      }
    }
    for (int c = 0; c < C; ++c) out[q * C + c] = tanhf(acc[c] + __ldg(bias + c));
  }
}
void launch_col2im_tanh(Launch L, int dt_cols, const void* cols, const float* bias, int B, int Hin, int Win, int C, float* out_f32) {
  GAN_REQUIRE(C <= 4, "col2im head supports up to 4 channels");
  const int64_t total = (int64_t)B * Hin * Win * 4;
  if (dt_cols == DT_F16) k_col2im_tanh<f16><<<grid_for(total, 256, 16), 256, 0, L.s>>>((const f16*)cols, bias, B, Hin, Win, C, out_f32);
  else k_col2im_tanh<float><<<grid_for(total, 256, 16), 256, 0, L.s>>>((const float*)cols, bias, B, Hin, Win, C, out_f32);
  KLAUNCH(L);
}

__global__ void __launch_bounds__(256) k_dhead_gather(const float* __restrict__ cols, const float* __restrict__ bias, int B,
                                                      int Hin, int Win, float* __restrict__ logits) {
  const int Ho = Hin - 1, Wo = Win - 1;
  const int64_t total = (int64_t)B * Ho * Wo;
  const float b0 = __ldg(bias);
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int ow = (int)(q % Wo); const int64_t r = q / Wo; const int oh = (int)(r % Ho); const int n = (int)(r / Ho);
    float acc = b0;
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const int i = oh + kh - 1;
      if (i < 0 || i >= Hin) continue;
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) {
        const int j = ow + kw - 1;
        if (j < 0 || j >= Win) continue;
        acc += __ldg(cols + (((int64_t)n * Hin + i) * Win + j) * 64 + (kh * 4 + kw) * 4);
      }
    }
    logits[q] = acc;
  }
}
void launch_dhead_gather(Launch L, const float* cols, const float* bias, int B, int Hin, int Win, float* logits) {
  const int64_t total = (int64_t)B * (Hin - 1) * (Win - 1);
  k_dhead_gather<<<grid_for(total, 256, 8), 256, 0, L.s>>>(cols, bias, B, Hin, Win, logits);
  KLAUNCH(L);
}

// (pure data movement of 16-bit elements: the same kernel serves f16 and bf16 gradients)
__global__ void __launch_bounds__(256) k_dhead_unfold(const bf16* __restrict__ dl, int pitch, int B, int Hin, int Win,
                                                      bf16* __restrict__ dst) {
  const int Ho = Hin - 1, Wo = Win - 1;
  const int64_t M = (int64_t)B * Hin * Win;
  const bf16 zero = __float2bfloat16_rn(0.f);
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(m % Win); const int64_t r = m / Win; const int i = (int)(r % Hin); const int n = (int)(r / Hin);
    uint2 row[16];
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const int oh = i - kh + 1;
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) {
        const int ow = j - kw + 1;
        bf16 v = zero;
        if (oh >= 0 && oh < Ho && ow >= 0 && ow < Wo) v = dl[(((int64_t)n * Ho + oh) * Wo + ow) * pitch];
        __nv_bfloat162 lo; lo.x = v; lo.y = zero;
        row[kh * 4 + kw] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), 0u);
      }
    }
    uint4* o = reinterpret_cast<uint4*>(dst + m * 64);
#pragma unroll
    for (int t = 0; t < 8; ++t) o[t] = make_uint4(row[2 * t].x, row[2 * t].y, row[2 * t + 1].x, row[2 * t + 1].y);
  }
}
void launch_dhead_unfold(Launch L, const void* dlogit_bf16, int pitch, int B, int Hin, int Win, void* dst_bf16) {
  const int64_t M = (int64_t)B * Hin * Win;
  k_dhead_unfold<<<grid_for(M, 256, 8), 256, 0, L.s>>>((const bf16*)dlogit_bf16, pitch, B, Hin, Win, (bf16*)dst_bf16);
  KLAUNCH(L);
}

// ---------------------------------------------------------------------------------------------
// Input pipeline: split + nearest resize(s) + crop + flip + normalize as one gather over uint8
// images (base_gan.py:45-61, pix2pix.py:34-112, cycle_gan.py:38-85).  One thread = 4 consecutive
// output floats of one row (16-byte store); HBM-bound on the fp32 output.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int nn_src(int o, float scale, int in) {
  // tf.image.resize(method=NEAREST_NEIGHBOR) of TF2: half-pixel centres, float32 arithmetic;
  // scale = (float)in / (float)out, divided once per image on the host (same IEEE float32 quotient)
  const int i = (int)floorf(__fmul_rn((float)o + 0.5f, scale));
  return i < in - 1 ? i : in - 1;
}
// grid = (quads of a row / 256, S rows, B images): no index divisions; the 256-entry normalize table
// (v / 127.5 - 1, correctly rounded) is built once per block.
template <int C>
__global__ void __launch_bounds__(256) k_preprocess(const uint8_t* __restrict__ img, int64_t stride,
                                                    const ImageXformDev* __restrict__ xf, int S, float* __restrict__ out) {
  __shared__ float lut[256];
  lut[threadIdx.x] = __fsub_rn(__fdiv_rn((float)threadIdx.x, 127.5f), 1.0f);
  __syncthreads();
  const int quads = S * C / 4;
  const int q = blockIdx.x * 256 + threadIdx.x;
  if (q >= quads) return;
  const int i = blockIdx.y, n = blockIdx.z;
  const ImageXformDev x = xf[n];
  const int g1h = x.pre > 0 ? x.pre : x.src_h, g1w = x.pre > 0 ? x.pre : x.cols;
  int r = x.mid > 0 ? nn_src(i + x.crop_y, x.sy1, g1h) : nn_src(i, x.sy1, g1h);
  if (x.pre > 0) r = nn_src(r, x.sy0, x.src_h);
  const uint8_t* row = img + (int64_t)n * stride + ((int64_t)r * x.src_w + x.col0) * C;
  float o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int idx = q * 4 + e; const int j = idx / C, ch = idx - j * C;
    const int jj = x.flip ? S - 1 - j : j;
    int c = x.mid > 0 ? nn_src(jj + x.crop_x, x.sx1, g1w) : nn_src(jj, x.sx1, g1w);
    if (x.pre > 0) c = nn_src(c, x.sx0, x.cols);
    o[e] = lut[row[c * C + ch]];
  }
  reinterpret_cast<float4*>(out)[((int64_t)n * S + i) * quads + q] = make_float4(o[0], o[1], o[2], o[3]);
}
void launch_preprocess(Launch L, const uint8_t* img, int64_t stride, const ImageXformDev* xf_dev, int B, int C, int S, float* out) {
  GAN_REQUIRE((S * C) % 4 == 0, "out_size * channels must be a multiple of 4");
  GAN_REQUIRE(S <= 65535 && B <= 65535, "image size / batch exceed the launch grid");
  const dim3 grid((S * C / 4 + 255) / 256, S, B);
  if (C == 1) k_preprocess<1><<<grid, 256, 0, L.s>>>(img, stride, xf_dev, S, out);
  else if (C == 2) k_preprocess<2><<<grid, 256, 0, L.s>>>(img, stride, xf_dev, S, out);
  else if (C == 3) k_preprocess<3><<<grid, 256, 0, L.s>>>(img, stride, xf_dev, S, out);
  else k_preprocess<4><<<grid, 256, 0, L.s>>>(img, stride, xf_dev, S, out);
  KLAUNCH(L);
}
