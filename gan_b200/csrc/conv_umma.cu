// conv_umma.cu — tcgen05 / TMEM / TMA implicit-GEMM convolution kernels for sm_100a (bf16 in,
// fp32 accumulate in tensor memory).
//
// Forward-type "tap GEMM" (Conv2D 4x4 s2 base_gan.py:78, ZeroPad+Conv2D 4x4 s1 base_gan.py:145-148,
// Conv2DTranspose 4x4 s2 base_gan.py:107 as four parity classes, and the data-gradient of each):
//     D[m, n] = sum_t sum_kc  A_t[m, kc] * B[n, t*Kc + kc]
//   * M = 128 M-space points = one TMA box (64 ch x TW x TH x TN) of the NHWC activation per tap;
//     padding is the TMA out-of-bounds zero fill (negative / overflowing coordinates), stride-2
//     taps read one of four parity sub-lattices of the same buffer (four tensor maps that differ
//     only in base address), so no im2col, space-to-depth or padded copy ever exists in HBM.
//   * A and B tiles land in shared memory as K-major SWIZZLE_128B and are consumed directly by
//     tcgen05.mma.cta_group::1.kind::f16 (M=128, N=64|128, K=16); the accumulator lives in TMEM.
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 2..5 = epilogue
//     (tcgen05.ld -> bf16 -> 16-byte global stores into the consumer's strided NHWC view).
//
// Weight-gradient GEMM (pixel reduction): dW[(t,kc), nc] = sum_m A_t[m, kc] * dY[m, nc]
//   * both operands are the same NHWC boxes, now read as MN-major SWIZZLE_128B operands
//     (K = pixels); split over pixel ranges across CTAs, fp32 red.global.add epilogue into the
//     gradient buffer in the master (TF) weight layout.
#include <cuda.h>
#include <unordered_map>
#include <mutex>
#include <cstring>
#include "kernels.h"

#define KLAUNCH(L) (++*(L).count)

// ---------------------------------------------------------------------------------------------
// driver entry point for cuTensorMapEncodeTiled (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

// A bf16 NHWC view sampled with pixel stride `s` starting at pixel offset (a,b):
// dims (C, W/s, H/s, N), box (64, bw, bh, bn), SWIZZLE_128B, OOB -> 0.
static CUtensorMap make_map4(const void* base, int pitch, int coff, int C, int H, int W, int N, int s, int a, int b,
                             int bw, int bh, int bn) {
  GAN_REQUIRE(g_encode != nullptr, "cuTensorMapEncodeTiled unavailable");
  CUtensorMap m;
  const char* p = (const char*)base + ((int64_t)(a * W + b) * pitch + coff) * 2;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)((W - b + s - 1) / s), (cuuint64_t)((H - a + s - 1) / s), (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)s * pitch * 2, (cuuint64_t)s * W * pitch * 2, (cuuint64_t)H * W * pitch * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)p, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw GanError(-2, "cuTensorMapEncodeTiled(4d) failed: " + std::to_string((int)r));
  return m;
}
// Packed weights [rows][K] K-major: dims (K, rows), box (64, bn).
static CUtensorMap make_map2(const void* base, int64_t K, int64_t rows, int bn) {
  GAN_REQUIRE(g_encode != nullptr, "cuTensorMapEncodeTiled unavailable");
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)bn};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw GanError(-2, "cuTensorMapEncodeTiled(2d) failed: " + std::to_string((int)r));
  return m;
}

static int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar), done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
}  // namespace ptx

// Shared-memory matrix descriptors (cute/arch/mma_sm100_desc.hpp SmemDescriptor bit layout).
// K-major SWIZZLE_128B: rows of 128 B, 8-row atoms of 1024 B => SBO = 1024, LBO unused (1).
__device__ __forceinline__ uint64_t desc_kmajor_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major SWIZZLE_128B: 64 MN elements (128 B) contiguous, 8 k-rows per 1024-B atom => SBO = 1024
// (next 8 k), LBO = byte distance between consecutive 64-element MN blocks.
__device__ __forceinline__ uint64_t desc_mnmajor_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor, kind::f16: D=f32, A=B=bf16 (InstrDescriptor bit layout).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// =============================================================================================
// forward-type kernel
// =============================================================================================
struct alignas(64) UmmaFwdParams {
  CUtensorMap amap[4];
  CUtensorMap bmap;
  int8_t tap_map[4][16], tap_dw[4][16], tap_dh[4][16];
  int ntaps[4], oa[4], ob[4];
  int kchunks;                 // Kc / 64
  int TW, TH, TN, tiles_w, tiles_h, tiles_n;
  bf16* out; int out_pitch, out_coff, Hout, Wout, so;
  int N, Hm, Wm, Nc;
};

constexpr int FWD_STAGES = 3;
constexpr int FWD_THREADS = 192;     // warp0 TMA, warp1 MMA, warps 2..5 epilogue

template <int BN>
__global__ void __launch_bounds__(FWD_THREADS) k_conv_fwd_umma(const __grid_constant__ UmmaFwdParams p) {
  constexpr uint32_t A_BYTES = 128 * 128;          // 128 rows x 64 bf16
  constexpr uint32_t B_BYTES = BN * 128;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + FWD_STAGES * STAGE_BYTES);
  uint64_t* empty = full + FWD_STAGES;
  uint64_t* tmem_full = empty + FWD_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cls = blockIdx.z;
  const int n0 = blockIdx.y * BN;
  int tm = blockIdx.x;
  const int tw_i = tm % p.tiles_w; tm /= p.tiles_w;
  const int th_i = tm % p.tiles_h; tm /= p.tiles_h;
  const int w0 = tw_i * p.TW, h0 = th_i * p.TH, b0 = tm * p.TN;
  const int ntaps = p.ntaps[cls];
  const int nk = ntaps * p.kchunks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.bmap);
    for (int i = 0; i < 4; ++i) ptx::prefetch_tmap(&p.amap[i]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < FWD_STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    ptx::mbar_init(tmem_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc(tmem_slot, BN);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % FWD_STAGES;
        const uint32_t ph = (kb / FWD_STAGES) & 1;
        ptx::mbar_wait(&empty[s], ph ^ 1);
        const int t = kb / p.kchunks, kc = kb - t * p.kchunks;
        uint8_t* sa = smem + s * STAGE_BYTES;
        ptx::mbar_expect_tx(&full[s], STAGE_BYTES);
        ptx::tma_load_4d(sa, &p.amap[p.tap_map[cls][t]], &full[s], kc * 64, w0 + p.tap_dw[cls][t], h0 + p.tap_dh[cls][t], b0);
        ptx::tma_load_2d(sa + A_BYTES, &p.bmap, &full[s], kb * 64, cls * p.Nc + n0);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
    for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % FWD_STAGES;
      const uint32_t ph = (kb / FWD_STAGES) & 1;
      ptx::mbar_wait(&full[s], ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t sa = ptx::smem_u32(smem + s * STAGE_BYTES);
        const uint64_t ad = desc_kmajor_sw128(sa), bd = desc_kmajor_sw128(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)     // 4 x (K=16) per 64-channel block: +32 B inside the swizzle atom
          ptx::umma_bf16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
        ptx::umma_commit(&empty[s]);
        if (kb == nk - 1) ptx::umma_commit(tmem_full);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;                         // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;                    // accumulator row == M-space point inside the tile
    const int wl = r % p.TW, hl = (r / p.TW) % p.TH, nl = r / (p.TW * p.TH);
    const int mw = w0 + wl, mh = h0 + hl, nb = b0 + nl;
    const int oh = mh * p.so + p.oa[cls], ow = mw * p.so + p.ob[cls];
    const bool valid = nb < p.N && mh < p.Hm && mw < p.Wm && oh < p.Hout && ow < p.Wout;
    bf16* dst = p.out + (((int64_t)nb * p.Hout + oh) * p.Wout + ow) * p.out_pitch + p.out_coff + n0;
    ptx::mbar_wait(tmem_full, 0);
    ptx::tc_fence_after();
#pragma unroll
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[32];
      ptx::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, v);
      if (valid) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            h[e] = __floats2bfloat162_rn(__uint_as_float(v[j * 8 + 2 * e]), __uint_as_float(v[j * 8 + 2 * e + 1]));
          *reinterpret_cast<uint4*>(dst + c + j * 8) = o;
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, BN);
}

static size_t fwd_smem_bytes(int BN) { return (size_t)FWD_STAGES * (128 * 128 + BN * 128) + 1024 + 128; }

static bool view_ok(int pitch, int coff, const void* p) {
  return pitch % 8 == 0 && coff % 8 == 0 && ((uintptr_t)p % 16) == 0;
}

bool umma_fwd_supported(const ConvOp& op) {
  if (g_encode == nullptr) return false;
  if (op.Kc % 64 != 0 || op.Nc % 64 != 0) return false;
  if (!view_ok(op.in_pitch, op.in_coff, op.in) || !view_ok(op.out_pitch, op.out_coff, op.out)) return false;
  if (op.epi != EPI_NONE || op.out_f32 != nullptr) return false;
  if (op.si == 2 && (op.Hin % 2 != 0 || op.Win % 2 != 0)) return false;
  return true;
}

static void fill_fwd_params(UmmaFwdParams& P, const ConvOp& op, int bn_tile) {
  memset(&P, 0, sizeof(P));
  int TW = pow2ceil(op.Wm); if (TW > 128) TW = 128;
  int TH = pow2ceil(op.Hm); if (TH > 128 / TW) TH = 128 / TW;
  int TN = 128 / (TW * TH);
  P.TW = TW; P.TH = TH; P.TN = TN;
  P.tiles_w = (op.Wm + TW - 1) / TW; P.tiles_h = (op.Hm + TH - 1) / TH; P.tiles_n = (op.N + TN - 1) / TN;
  if (op.si == 1) {
    P.amap[0] = make_map4(op.in, op.in_pitch, op.in_coff, op.Kc, op.Hin, op.Win, op.N, 1, 0, 0, TW, TH, TN);
    for (int i = 1; i < 4; ++i) P.amap[i] = P.amap[0];
  } else {
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        P.amap[a * 2 + b] = make_map4(op.in, op.in_pitch, op.in_coff, op.Kc, op.Hin, op.Win, op.N, 2, a, b, TW, TH, TN);
  }
  int64_t Ktot = (int64_t)op.cls[0].ntaps * op.Kc;
  P.bmap = make_map2(op.B, Ktot, (int64_t)op.ncls * op.Nc, bn_tile);
  for (int c = 0; c < op.ncls; ++c) {
    const ClassGeom& g = op.cls[c];
    P.ntaps[c] = g.ntaps; P.oa[c] = g.oa; P.ob[c] = g.ob;
    for (int t = 0; t < g.ntaps; ++t) {
      if (op.si == 1) { P.tap_map[c][t] = 0; P.tap_dh[c][t] = g.dh[t]; P.tap_dw[c][t] = g.dw[t]; }
      else {
        int a = ((g.dh[t] % 2) + 2) % 2, b = ((g.dw[t] % 2) + 2) % 2;
        P.tap_map[c][t] = (int8_t)(a * 2 + b);
        P.tap_dh[c][t] = (int8_t)((g.dh[t] - a) / 2); P.tap_dw[c][t] = (int8_t)((g.dw[t] - b) / 2);
      }
    }
  }
  P.kchunks = op.Kc / 64;
  P.out = (bf16*)op.out; P.out_pitch = op.out_pitch; P.out_coff = op.out_coff; P.Hout = op.Hout; P.Wout = op.Wout; P.so = op.so;
  P.N = op.N; P.Hm = op.Hm; P.Wm = op.Wm; P.Nc = op.Nc;
}

void launch_conv_fwd_umma(Launch L, const ConvOp& op) {
  const int BN = (op.Nc % 128 == 0) ? 128 : 64;
  UmmaFwdParams P;
  fill_fwd_params(P, op, BN);
  dim3 grid(P.tiles_w * P.tiles_h * P.tiles_n, op.Nc / BN, op.ncls);
  if (BN == 128) k_conv_fwd_umma<128><<<grid, FWD_THREADS, fwd_smem_bytes(128), L.s>>>(P);
  else k_conv_fwd_umma<64><<<grid, FWD_THREADS, fwd_smem_bytes(64), L.s>>>(P);
  KLAUNCH(L);
}

// =============================================================================================
// weight-gradient kernel
// =============================================================================================
struct alignas(64) UmmaWgradParams {
  CUtensorMap amap[4];     // input activation boxes (64 ch x PW x PH x PN), parity sub-lattices when si == 2
  CUtensorMap dmap[4];     // dY boxes, parity sub-lattices when so == 2
  int8_t tap_map[4][16], tap_dw[4][16], tap_dh[4][16], widx[4][16];
  int ntaps[4], dy_map[4];
  int Kc, kc_blocks;       // channels per tap, Kc/64
  int PW, PH, PN, tiles_w, tiles_h, tiles_n;   // 64-pixel boxes over the M-space
  int splits;
  float* dW; long long s_tap, s_k, s_n;
  int Nc;
};

constexpr int WG_STAGES = 3;
constexpr int WG_PIX = 64;            // pixels (GEMM-K) per stage

template <int BN>
__global__ void __launch_bounds__(FWD_THREADS) k_conv_wgrad_umma(const __grid_constant__ UmmaWgradParams p) {
  constexpr uint32_t BOX_BYTES = WG_PIX * 128;            // 64 pixels x 64 channels bf16
  constexpr uint32_t A_BYTES = 2 * BOX_BYTES;             // 128 (tap,kc) rows
  constexpr uint32_t B_BYTES = (BN / 64) * BOX_BYTES;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + WG_STAGES * STAGE_BYTES);
  uint64_t* empty = full + WG_STAGES;
  uint64_t* tmem_full = empty + WG_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cls = (int)(blockIdx.z / p.splits);
  const int split = blockIdx.z % p.splits;
  const int n0 = blockIdx.y * BN;
  const int kblk0 = blockIdx.x * 2;                        // two 64-row blocks of the (tap,kc) index space
  const int total_boxes = p.tiles_w * p.tiles_h * p.tiles_n;
  const int per = (total_boxes + p.splits - 1) / p.splits;
  const int box_beg = split * per;
  const int box_end = min(total_boxes, box_beg + per);
  const int nk = box_end - box_beg;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) { ptx::prefetch_tmap(&p.amap[i]); ptx::prefetch_tmap(&p.dmap[i]); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    ptx::mbar_init(tmem_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc(tmem_slot, BN);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (nk > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int t_[2], c_[2];
        for (int i = 0; i < 2; ++i) { int k = (kblk0 + i) * 64; t_[i] = k / p.Kc; c_[i] = k - t_[i] * p.Kc; }
        for (int kb = 0; kb < nk; ++kb) {
          const int s = kb % WG_STAGES;
          const uint32_t ph = (kb / WG_STAGES) & 1;
          ptx::mbar_wait(&empty[s], ph ^ 1);
          int bx = box_beg + kb;
          const int w0 = (bx % p.tiles_w) * p.PW; bx /= p.tiles_w;
          const int h0 = (bx % p.tiles_h) * p.PH; bx /= p.tiles_h;
          const int b0 = bx * p.PN;
          uint8_t* sa = smem + s * STAGE_BYTES;
          ptx::mbar_expect_tx(&full[s], STAGE_BYTES);
#pragma unroll
          for (int i = 0; i < 2; ++i)
            ptx::tma_load_4d(sa + i * BOX_BYTES, &p.amap[p.tap_map[cls][t_[i]]], &full[s], c_[i],
                             w0 + p.tap_dw[cls][t_[i]], h0 + p.tap_dh[cls][t_[i]], b0);
#pragma unroll
          for (int i = 0; i < BN / 64; ++i)
            ptx::tma_load_4d(sa + A_BYTES + i * BOX_BYTES, &p.dmap[p.dy_map[cls]], &full[s], n0 + i * 64, w0, h0, b0);
        }
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = make_idesc(128, BN, 1, 1);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % WG_STAGES;
        const uint32_t ph = (kb / WG_STAGES) & 1;
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t sa = ptx::smem_u32(smem + s * STAGE_BYTES);
          const uint64_t ad = desc_mnmajor_sw128(sa, BOX_BYTES), bd = desc_mnmajor_sw128(sa + A_BYTES, BOX_BYTES);
#pragma unroll
          for (int k = 0; k < WG_PIX / 16; ++k)   // 16 pixels (2 atoms of 8 k-rows = 2048 B) per MMA
            ptx::umma_bf16(tmem_base, ad + (uint64_t)(k * 2048 >> 4), bd + (uint64_t)(k * 2048 >> 4), idesc, (kb | k) != 0);
          ptx::umma_commit(&empty[s]);
          if (kb == nk - 1) ptx::umma_commit(tmem_full);
        }
        __syncwarp();
      }
    } else {
      const int q = warp & 3;
      const int r = q * 32 + lane;                   // accumulator row = (tap,kc) index inside the 128 block
      const int k = kblk0 * 64 + r;
      const int t = k / p.Kc, kc = k - t * p.Kc;
      float* dst = p.dW + (long long)p.widx[cls][t] * p.s_tap + (long long)kc * p.s_k + (long long)n0 * p.s_n;
      ptx::mbar_wait(tmem_full, 0);
      ptx::tc_fence_after();
#pragma unroll
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        ptx::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dst + (long long)(c + j) * p.s_n, __uint_as_float(v[j]));
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, BN);
}

static size_t wg_smem_bytes(int BN) { return (size_t)WG_STAGES * ((2 + BN / 64) * WG_PIX * 128) + 1024 + 128; }

bool umma_wgrad_supported(const ConvOp& op) {
  if (g_encode == nullptr) return false;
  if (op.Kc % 64 != 0 || op.Nc % 64 != 0) return false;
  if (!view_ok(op.in_pitch, op.in_coff, op.in) || !view_ok(op.out_pitch, op.out_coff, op.out)) return false;
  if (op.si == 2 && (op.Hin % 2 != 0 || op.Win % 2 != 0)) return false;
  if (op.so == 2 && (op.Hout % 2 != 0 || op.Wout % 2 != 0)) return false;
  return true;
}

void launch_conv_wgrad_umma(Launch L, const ConvOp& op) {
  const int BN = (op.Nc % 128 == 0) ? 128 : 64;
  UmmaWgradParams P; memset(&P, 0, sizeof(P));
  int PW = pow2ceil(op.Wm); if (PW > WG_PIX) PW = WG_PIX;
  int PH = pow2ceil(op.Hm); if (PH > WG_PIX / PW) PH = WG_PIX / PW;
  int PN = WG_PIX / (PW * PH);
  P.PW = PW; P.PH = PH; P.PN = PN;
  P.tiles_w = (op.Wm + PW - 1) / PW; P.tiles_h = (op.Hm + PH - 1) / PH; P.tiles_n = (op.N + PN - 1) / PN;
  if (op.si == 1) {
    P.amap[0] = make_map4(op.in, op.in_pitch, op.in_coff, op.Kc, op.Hin, op.Win, op.N, 1, 0, 0, PW, PH, PN);
    for (int i = 1; i < 4; ++i) P.amap[i] = P.amap[0];
  } else {
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        P.amap[a * 2 + b] = make_map4(op.in, op.in_pitch, op.in_coff, op.Kc, op.Hin, op.Win, op.N, 2, a, b, PW, PH, PN);
  }
  if (op.so == 1) {
    P.dmap[0] = make_map4(op.out, op.out_pitch, op.out_coff, op.Nc, op.Hout, op.Wout, op.N, 1, 0, 0, PW, PH, PN);
    for (int i = 1; i < 4; ++i) P.dmap[i] = P.dmap[0];
  } else {
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        P.dmap[a * 2 + b] = make_map4(op.out, op.out_pitch, op.out_coff, op.Nc, op.Hout, op.Wout, op.N, 2, a, b, PW, PH, PN);
  }
  for (int c = 0; c < op.ncls; ++c) {
    const ClassGeom& g = op.cls[c];
    P.ntaps[c] = g.ntaps;
    P.dy_map[c] = op.so == 1 ? 0 : g.oa * 2 + g.ob;
    for (int t = 0; t < g.ntaps; ++t) {
      P.widx[c][t] = g.widx[t];
      if (op.si == 1) { P.tap_map[c][t] = 0; P.tap_dh[c][t] = g.dh[t]; P.tap_dw[c][t] = g.dw[t]; }
      else {
        int a = ((g.dh[t] % 2) + 2) % 2, b = ((g.dw[t] % 2) + 2) % 2;
        P.tap_map[c][t] = (int8_t)(a * 2 + b);
        P.tap_dh[c][t] = (int8_t)((g.dh[t] - a) / 2); P.tap_dw[c][t] = (int8_t)((g.dw[t] - b) / 2);
      }
    }
  }
  P.Kc = op.Kc; P.kc_blocks = op.Kc / 64;
  P.dW = op.dW; P.s_tap = op.s_tap; P.s_k = op.s_k; P.s_n = op.s_n; P.Nc = op.Nc;
  const int ntaps = op.cls[0].ntaps;
  const int mblocks = ntaps * op.Kc / 128;
  const int ntiles = op.Nc / BN;
  const int total_boxes = P.tiles_w * P.tiles_h * P.tiles_n;
  int64_t ctas = (int64_t)mblocks * ntiles * op.ncls;
  int splits = (int)((148 * 2 + ctas - 1) / ctas);
  if (splits > total_boxes) splits = total_boxes;
  if (splits < 1) splits = 1;
  P.splits = splits;
  dim3 grid(mblocks, ntiles, op.ncls * splits);
  if (BN == 128) k_conv_wgrad_umma<128><<<grid, FWD_THREADS, wg_smem_bytes(128), L.s>>>(P);
  else k_conv_wgrad_umma<64><<<grid, FWD_THREADS, wg_smem_bytes(64), L.s>>>(P);
  KLAUNCH(L);
}

void umma_init() {
  static std::once_flag once;
  std::call_once(once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = (PFN_encodeTiled)fn;
    else cudaGetLastError();
    cudaFuncSetAttribute(k_conv_fwd_umma<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem_bytes(128));
    cudaFuncSetAttribute(k_conv_fwd_umma<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem_bytes(64));
    cudaFuncSetAttribute(k_conv_wgrad_umma<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg_smem_bytes(128));
    cudaFuncSetAttribute(k_conv_wgrad_umma<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg_smem_bytes(64));
  });
}
