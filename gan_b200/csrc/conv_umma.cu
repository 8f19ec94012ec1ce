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
#include <cstdlib>
#include <cstdio>
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
                             int bw, int bh, int bn, int bc = 64) {
  GAN_REQUIRE(g_encode != nullptr, "cuTensorMapEncodeTiled unavailable");
  CUtensorMap m;
  const char* p = (const char*)base + ((int64_t)(a * W + b) * pitch + coff) * 2;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)((W - b + s - 1) / s), (cuuint64_t)((H - a + s - 1) / s), (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)s * pitch * 2, (cuuint64_t)s * W * pitch * 2, (cuuint64_t)H * W * pitch * 2};
  cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)p, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, bc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw GanError(-2, "cuTensorMapEncodeTiled(4d) failed: " + std::to_string((int)r));
  return m;
}
// Packed weights [rows][K] K-major: dims (K, rows), box (64, bn).
static CUtensorMap make_map2(const void* base, int64_t K, int64_t rows, int bn, int bk = 64) {
  GAN_REQUIRE(g_encode != nullptr, "cuTensorMapEncodeTiled unavailable");
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)bn};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// ---- cta_group::2 (CTA pair on one TPC) ------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: the data lands in THIS CTA's shared memory, the transaction bytes are counted on the
// mbarrier of the pair's leader CTA (`bar_cluster_addr`, a shared::cluster address)
__device__ __forceinline__ void tma_load_4d_2sm(void* smem, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[128 rows of each CTA] * B[N/2 rows of each CTA]^T : one M=256 MMA over the pair
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the mbarrier at this smem offset in BOTH CTAs of the pair once all prior MMAs of this thread are done
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
}  // namespace ptx

// Shared-memory matrix descriptors (cute/arch/mma_sm100_desc.hpp SmemDescriptor bit layout):
// start address [0,14) (>>4), LBO [16,30) (>>4), SBO [32,46) (>>4), version=1 [46,48),
// layout type [61,64): 2 = SWIZZLE_128B, 6 = SWIZZLE_32B.
//
// K-major swizzled tile: rows of SW bytes (SW = 128 or 32), 8-row atoms => SBO = 8*SW; LBO unused.
template <int SW>
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
  constexpr uint64_t layout = (SW == 128) ? 2ull : 6ull;
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)((8 * SW) >> 4) << 32) | (1ull << 46) | (layout << 61);
}
// MN-major swizzled tile: SW bytes of MN contiguous per k-row, 8 k-rows per atom => SBO = 8*SW
// (next 8 k), LBO = byte distance between consecutive MN blocks of SW bytes.
template <int SW>
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr, uint32_t lbo_bytes) {
  constexpr uint64_t layout = (SW == 128) ? 2ull : 6ull;
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((8 * SW) >> 4) << 32) | (1ull << 46) | (layout << 61);
}
// Instruction descriptor, kind::f16 (InstrDescriptor bit layout): D format [4,6) = 1 (f32), A format [7,10) and
// B format [10,13) = 0 (f16) | 1 (bf16) — chosen per operand, so an f16 activation tile can meet a bf16 gradient
// tile in one MMA —, A/B major bits 15/16, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major, int a_bf16 = 1, int b_bf16 = 1) {
  return (1u << 4) | ((uint32_t)a_bf16 << 7) | ((uint32_t)b_bf16 << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// two fp32 accumulators -> one packed 16-bit pair of the output dtype
__device__ __forceinline__ uint32_t cvt_pair(float lo, float hi, int out_f16) {
  return out_f16 ? pack2<f16>(lo, hi) : pack2<bf16>(lo, hi);
}

// =============================================================================================
// forward-type kernel.  Template: BN = N tile (128 | 64 | 16), KC = channels per TMA box
// (64: one 128-byte-swizzled box per 64-channel k-block; 16: channel-padded small-Cin layers,
// four 32-byte-swizzled boxes = four taps per k-block).
// =============================================================================================
struct alignas(64) UmmaFwdParams {
  CUtensorMap amap[4];
  CUtensorMap bmap;
  int8_t tap_map[4][16], tap_dw[4][16], tap_dh[4][16];
  int ntaps[4], oa[4], ob[4];
  int kchunks;                 // Kc / 64 (KC == 64)
  int TW, TH, TN, tiles_w, tiles_h, tiles_n;
  bf16* out; int out_pitch, out_coff, Hout, Wout, so;   // 16-bit elements (f16 or bf16 bit patterns, see out_f16)
  int N, Hm, Wm, Nc;
  const float* bias; float* out_f32; int epi, Nr;
  int num_tiles, ncls;
  int f32out;                        // write fp32 rows to ws (slab 0) instead of bf16 (cols of the generator head)
  int ab_bf16;                       // operand format of A and B (both: the weights are packed in the input's dtype)
  int out_f16;                       // 16-bit output format: 1 = f16 (activations), 0 = bf16 (gradients)
  float* ws; int ksplit; long long slab;   // split-K: split ks stores fp32 into ws[ks][pix][Nc] (no atomics)
  int stage_out;                     // one-CTA kernel: 16-bit output rows written through the shared-memory staging (coalesced)
  float* stats_ws;                   // CTA-pair kernel: per-(CTA, quadrant) column sums / sums of squares, or nullptr
};

constexpr int FWD_STAGES = 3;         // BN <= 128: 3 x 32 KB, two CTAs per SM
constexpr int FWD_STAGES_256 = 4;     // BN == 256: 4 x 48 KB, one CTA per SM, all 512 TMEM columns
__host__ __device__ constexpr int fwd_stages(int BN) { return BN == 256 ? FWD_STAGES_256 : FWD_STAGES; }
constexpr int FWD_THREADS = 192;     // warp0 TMA, warp1 MMA, warps 2..5 epilogue
constexpr int F1_STAT_NC = 256;      // epilogue BatchNorm statistics of the one-CTA kernel: up to 256 output channels (8 KB)

__device__ __forceinline__ float epi_apply(float v, int epi, const float* bias, int n) {
  if (epi != EPI_NONE) v += __ldg(bias + n);
  if (epi == EPI_BIAS_TANH) v = tanhf(v);
  return v;
}

// Persistent: each CTA loops over output tiles (tile = blockIdx.x + i*gridDim.x, M-tile index
// fastest so that concurrently running CTAs share the same weight tile in L2).  Two TMEM accumulator
// buffers: the epilogue warps drain tile i while the MMA warp already accumulates tile i+1, and the
// TMA producer runs ahead across tile boundaries through the shared-memory ring.
template <int BN, int KC>
__global__ void __launch_bounds__(FWD_THREADS) k_conv_fwd_umma(const __grid_constant__ UmmaFwdParams p) {
  constexpr int SW = (KC == 64) ? 128 : 32;
  constexpr int SUB = 64 / KC;                       // TMA boxes (taps) per 64-element k-block
  constexpr uint32_t A_SUB = 128 * KC * 2;           // bytes of one A box
  constexpr uint32_t B_SUB = BN * KC * 2;
  constexpr uint32_t A_BYTES = 128 * 128;
  constexpr uint32_t B_BYTES = BN * 128;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t ACC_COLS = BN < 32 ? 32 : BN;   // TMEM columns of one accumulator buffer
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
  constexpr int NSTAGE = fwd_stages(BN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + NSTAGE * STAGE_BYTES);
  uint64_t* empty = full + NSTAGE;
  uint64_t* tmem_full = empty + NSTAGE;          // [2]
  uint64_t* tmem_empty = tmem_full + 2;              // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  float2* sacc = (float2*)(tmem_slot + 4);           // [4 quadrants][F1_STAT_NC] BatchNorm statistics (stats_ws != nullptr)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntn = p.Nc / BN;
  const bool stats = BN >= 32 && p.stats_ws != nullptr;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.bmap);
    for (int i = 0; i < 4; ++i) ptx::prefetch_tmap(&p.amap[i]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&tmem_full[b], 1); ptx::mbar_init(&tmem_empty[b], 4); }
    ptx::fence_barrier_init();
  }
  if (stats)
    for (int i = threadIdx.x; i < 4 * F1_STAT_NC; i += FWD_THREADS) sacc[i] = make_float2(0.f, 0.f);
  if (warp == 2) ptx::tmem_alloc(tmem_slot, TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (class, n-tile, M-space origin)
  // tile order: (class, n-tile) fastest, M tile slowest: the CTAs running at the same time share
  // the same activation boxes (L2 hits instead of HBM re-reads; the weight tiles always fit in L2)
  const int inner = ntn * p.ncls;
  auto decode = [&](int tile, int& cls, int& n0, int& w0, int& h0, int& b0) {
    tile /= p.ksplit;                                  // k-split index is the fastest tile coordinate
    int tm = tile / inner; int r = tile - tm * inner;
    n0 = (r % ntn) * BN; cls = r / ntn;
    w0 = (tm % p.tiles_w) * p.TW; tm /= p.tiles_w;
    h0 = (tm % p.tiles_h) * p.TH; tm /= p.tiles_h;
    b0 = tm * p.TN;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;                               // global k-block counter (ring position)
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int cls, n0, w0, h0, b0;
        decode(tile, cls, n0, w0, h0, b0);
        const int nk = (KC == 64) ? p.ntaps[cls] * p.kchunks : p.ntaps[cls] / SUB;
        const int ks = tile % p.ksplit;
        const int kb0 = ks * nk / p.ksplit, kb1 = (ks + 1) * nk / p.ksplit;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % NSTAGE;
          const uint32_t ph = (it / NSTAGE) & 1;
          ptx::mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = smem + s * STAGE_BYTES;
          ptx::mbar_expect_tx(&full[s], STAGE_BYTES);
          if (KC == 64) {
            const int t = kb / p.kchunks, kc = kb - t * p.kchunks;
            ptx::tma_load_4d(sa, &p.amap[p.tap_map[cls][t]], &full[s], kc * 64, w0 + p.tap_dw[cls][t], h0 + p.tap_dh[cls][t], b0);
            ptx::tma_load_2d(sa + A_BYTES, &p.bmap, &full[s], kb * 64, cls * p.Nc + n0);
          } else {
#pragma unroll
            for (int j = 0; j < SUB; ++j) {
              const int t = kb * SUB + j;
              ptx::tma_load_4d(sa + j * A_SUB, &p.amap[p.tap_map[cls][t]], &full[s], 0, w0 + p.tap_dw[cls][t],
                               h0 + p.tap_dh[cls][t], b0);
              ptx::tma_load_2d(sa + A_BYTES + j * B_SUB, &p.bmap, &full[s], t * KC, cls * p.Nc + n0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(128, BN, 0, 0, p.ab_bf16, p.ab_bf16);
    uint32_t it = 0, li = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++li) {
      const int cls = ((tile / p.ksplit) % inner) / ntn;
      const int nk = (KC == 64) ? p.ntaps[cls] * p.kchunks : p.ntaps[cls] / SUB;
      const int ks = tile % p.ksplit;
      const int kb0 = ks * nk / p.ksplit, kb1 = (ks + 1) * nk / p.ksplit;
      const uint32_t buf = li & 1, use = li >> 1;
      ptx::mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);     // epilogue has drained this accumulator
      ptx::tc_fence_after();
      const uint32_t acc = tmem_base + buf * ACC_COLS;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t sa = ptx::smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // KC==64: +32 B inside the 128-byte swizzle atom per K=16; KC==16: one 32-byte-swizzled box per K=16
            const uint64_t ad = (KC == 64) ? desc_kmajor<SW>(sa) + 2 * k : desc_kmajor<SW>(sa + k * A_SUB);
            const uint64_t bd = (KC == 64) ? desc_kmajor<SW>(sa + A_BYTES) + 2 * k : desc_kmajor<SW>(sa + A_BYTES + k * B_SUB);
            ptx::umma_bf16(acc, ad, bd, idesc, (kb > kb0) || (k != 0));
          }
          ptx::umma_commit(&empty[s]);
          if (kb == kb1 - 1) ptx::umma_commit(&tmem_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;                         // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;                    // accumulator row == M-space point inside the tile
    const int wl = r % p.TW, hl = (r / p.TW) % p.TH, nl = r / (p.TW * p.TH);
    uint32_t li = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++li) {
      int cls, n0, w0, h0, b0;
      decode(tile, cls, n0, w0, h0, b0);
      const uint32_t buf = li & 1, use = li >> 1;
      const int mw = w0 + wl, mh = h0 + hl, nb = b0 + nl;
      const int oh = mh * p.so + p.oa[cls], ow = mw * p.so + p.ob[cls];
      const bool valid = nb < p.N && mh < p.Hm && mw < p.Wm && oh < p.Hout && ow < p.Wout;
      const int64_t pix = ((int64_t)nb * p.Hout + oh) * p.Wout + ow;
      bf16* dst = p.out + pix * p.out_pitch + p.out_coff + n0;
      // Coalesced 16-bit output (launches without epilogue statistics; the 8 KB accumulator region then serves as four
      // warp-private 2 KB staging buffers): a lane holds one output ROW, so a direct 16-byte store per lane touches 32
      // different 128-byte lines per instruction = 32 LSU wavefronts (what bounded the first version of the first-layer
      // kernel).  Staged through shared memory (XOR-swizzled 16-byte slots, conflict-free both ways),
      // four lanes write the 64 contiguous bytes of a row: 8 lines per instruction.
      const bool staged = BN >= 32 && p.stage_out != 0 && !stats && p.ksplit == 1 && !p.f32out;
      bf16* rdst[4]; bool rvalid[4];
      if (staged) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int src = 8 * i + (lane >> 2);
          const unsigned long long a = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)dst, src);
          rdst[i] = reinterpret_cast<bf16*>((uintptr_t)a) + (lane & 3) * 8;
          rvalid[i] = __shfl_sync(0xffffffffu, valid ? 1 : 0, src) != 0;
        }
      }
      uint4* stg = reinterpret_cast<uint4*>(sacc) + q * 128;
      ptx::mbar_wait(&tmem_full[buf], use & 1);
      ptx::tc_fence_after();
      const uint32_t acc = tmem_base + buf * ACC_COLS + ((uint32_t)(q * 32) << 16);
      if (BN >= 32) {
#pragma unroll
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          ptx::tmem_ld32(acc + c, v);
          if (staged) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                o[e] = cvt_pair(__uint_as_float(v[j * 8 + 2 * e]), __uint_as_float(v[j * 8 + 2 * e + 1]), p.out_f16);
              stg[lane * 4 + (j ^ ((lane >> 1) & 3))] = make_uint4(o[0], o[1], o[2], o[3]);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = 8 * i + (lane >> 2), j = lane & 3;
              const uint4 w = stg[r * 4 + (j ^ ((r >> 1) & 3))];
              if (rvalid[i]) *reinterpret_cast<uint4*>(rdst[i] + c) = w;
            }
            __syncwarp();
          } else if (valid && (p.ksplit > 1 || p.f32out)) {
            float4* wdst = reinterpret_cast<float4*>(p.ws + (long long)(tile % p.ksplit) * p.slab + pix * p.Nc + n0 + c);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              wdst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                    __uint_as_float(v[4 * j + 3]));
          } else if (valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                o[e] = cvt_pair(__uint_as_float(v[j * 8 + 2 * e]), __uint_as_float(v[j * 8 + 2 * e + 1]), p.out_f16);
              *reinterpret_cast<uint4*>(dst + c + j * 8) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
          if (stats) {
            // column sums over the warp's 32 rows without shared-memory traffic: a 5-round butterfly in which every
            // lane keeps the half of its values whose column bit equals its own lane bit and hands the other half to
            // lane ^ s; after the last round lane l holds column l (31 shuffles per quantity)
            float x[32], y[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) { x[j] = valid ? __uint_as_float(v[j]) : 0.f; y[j] = x[j] * x[j]; }
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) {
              const bool up = (lane & sft) != 0;
#pragma unroll
              for (int i = 0; i < sft; ++i) {
                const float kx = up ? x[i + sft] : x[i], sx = up ? x[i] : x[i + sft];
                const float ky = up ? y[i + sft] : y[i], sy = up ? y[i] : y[i + sft];
                x[i] = kx + __shfl_xor_sync(0xffffffffu, sx, sft);
                y[i] = ky + __shfl_xor_sync(0xffffffffu, sy, sft);
              }
            }
            float2 a = sacc[q * F1_STAT_NC + n0 + c + lane];
            a.x += x[0]; a.y += y[0];
            sacc[q * F1_STAT_NC + n0 + c + lane] = a;
          }
        }
      } else {
        // N = 16 tile: channel-padded heads (fp32 output with bias / tanh) and padded data gradients
        uint32_t v[16];
        ptx::tmem_ld16(acc, v);
        if (valid) {
          if (p.out_f32 != nullptr) {
            for (int n = 0; n < p.Nr; ++n) p.out_f32[pix * p.Nr + n] = epi_apply(__uint_as_float(v[n]), p.epi, p.bias, n);
          }
          if (p.out != nullptr) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                o[e] = cvt_pair(__uint_as_float(v[j * 8 + 2 * e]), __uint_as_float(v[j * 8 + 2 * e + 1]), p.out_f16);
              *reinterpret_cast<uint4*>(dst + j * 8) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
        }
      }
      // accumulator buffer fully read into registers: hand it back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[buf]);
    }
    if (stats) {
      // one partial per CTA (quadrants summed in a fixed order): ws[cta][0][c] = sum, [1][c] = sum of squares
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float* o = p.stats_ws + ((size_t)blockIdx.x * 2) * p.Nc;
      for (int c = threadIdx.x - 64; c < p.Nc; c += 128) {
        const float2 a0 = sacc[c], a1 = sacc[F1_STAT_NC + c], a2 = sacc[2 * F1_STAT_NC + c], a3 = sacc[3 * F1_STAT_NC + c];
        o[c] = (a0.x + a1.x) + (a2.x + a3.x); o[p.Nc + c] = (a0.y + a1.y) + (a2.y + a3.y);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}


// =============================================================================================
// CTA-pair forward-type kernel (tcgen05.mma.cta_group::2): one 256 x BN output tile per CTA pair.
//   * each CTA of the pair loads ITS 128 M-space rows of A and HALF of the B tile (BN/2 weight rows); the pair's
//     leader issues M=256 MMAs that read both halves, so every weight byte is fetched from L2 once per 256 rows
//     (half the L2->SMEM operand traffic per FLOP of the one-CTA tile, the measured limiter of the N<=256 layers)
//     and each SM's shared-memory operand read per MMA drops from A+B to A+B/2;
//   * 8 epilogue warps (two per TMEM lane quadrant, alternating 32-column chunks) drain a double-buffered
//     accumulator; the deeper TMA ring (5-6 stages) is paid for by the smaller per-CTA stage;
//   * BatchNorm statistics come from the fp32 accumulators: each epilogue warp transposes its 32x32 chunk through
//     shared memory, lane c sums column c, and the per-(CTA, quadrant) sums live in shared memory until the end of
//     the persistent loop (one deterministic partial per CTA and quadrant -> k_stats_finalize), which removes the
//     separate statistics pass over z.
// Restrictions (launch_conv_fwd_umma falls back to the one-CTA kernel otherwise): KC = 64, 16-bit output, no bias
// epilogue, no split-K, no fp32 row output.
// =============================================================================================
constexpr int F2_THREADS = 320;      // warp0 TMA, warp1 MMA (leader CTA only), warps 2..9 epilogue
constexpr int F2_EPI_WARPS = 8;
constexpr int F2_STAT_NC = 512;      // statistics accumulators: up to 512 output channels
__host__ __device__ constexpr int f2_stages(int BN) { return BN == 256 ? 5 : 6; }
__host__ __device__ constexpr uint32_t f2_stage_bytes(int BN) { return 128u * 128u + (uint32_t)(BN / 2) * 128u; }

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(F2_THREADS, 1) k_conv_fwd_umma2(const __grid_constant__ UmmaFwdParams p) {
  constexpr uint32_t A_BYTES = 128 * 128;
  constexpr uint32_t STAGE_BYTES = f2_stage_bytes(BN);
  constexpr uint32_t ACC_COLS = BN;
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;       // 128 | 256 | 512
  constexpr int NSTAGE = f2_stages(BN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + NSTAGE * STAGE_BYTES);     // leader's are used
  uint64_t* empty = full + NSTAGE;                               // own
  uint64_t* tmem_full = empty + NSTAGE;                          // [2] own
  uint64_t* tmem_empty = tmem_full + 2;                          // [2] leader's are used
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  float* scr = (float*)(tmem_slot + 4);                          // [8 warps][32 columns][33]
  float2* sacc = (float2*)(scr + F2_EPI_WARPS * 32 * 33);        // [4 quadrants][F2_STAT_NC]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int ntn = p.Nc / BN;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.bmap);
    for (int i = 0; i < 4; ++i) ptx::prefetch_tmap(&p.amap[i]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&tmem_full[b], 1); ptx::mbar_init(&tmem_empty[b], 2 * F2_EPI_WARPS); }
    ptx::fence_barrier_init();
  }
  if (p.stats_ws != nullptr)
    for (int i = threadIdx.x; i < 4 * F2_STAT_NC; i += F2_THREADS) sacc[i] = make_float2(0.f, 0.f);
  if (warp == 2) ptx::tmem_alloc_2sm(tmem_slot, TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();               // the peer's barriers are initialised before anything signals them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // pair tile -> (class, n-tile, pair of M tiles); (class, n-tile) fastest as in the one-CTA kernel
  const int inner = ntn * p.ncls;
  auto decode = [&](int tile, int& cls, int& n0, int& w0, int& h0, int& b0) {
    int pm = tile / inner; int r = tile - pm * inner;
    n0 = (r % ntn) * BN; cls = r / ntn;
    int tm = 2 * pm + (int)rank;                     // this CTA's M tile (may lie beyond the last: all rows invalid)
    w0 = (tm % p.tiles_w) * p.TW; tm /= p.tiles_w;
    h0 = (tm % p.tiles_h) * p.TH; tm /= p.tiles_h;
    b0 = tm * p.TN;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t full_leader[NSTAGE];
#pragma unroll
      for (int s = 0; s < NSTAGE; ++s) full_leader[s] = ptx::mapa_u32(ptx::smem_u32(&full[s]), 0);
      uint32_t it = 0;
      for (int tile = pair; tile < p.num_tiles; tile += npairs) {
        int cls, n0, w0, h0, b0;
        decode(tile, cls, n0, w0, h0, b0);
        const int nk = p.ntaps[cls] * p.kchunks;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % NSTAGE;
          const uint32_t ph = (it / NSTAGE) & 1;
          ptx::mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = smem + s * STAGE_BYTES;
          if (rank == 0) ptx::mbar_expect_tx(&full[s], 2 * STAGE_BYTES);      // both CTAs' bytes land on the leader's barrier
          const int t = kb / p.kchunks, kc = kb - t * p.kchunks;
          ptx::tma_load_4d_2sm(sa, &p.amap[p.tap_map[cls][t]], full_leader[s], kc * 64, w0 + p.tap_dw[cls][t], h0 + p.tap_dh[cls][t], b0);
          ptx::tma_load_2d_2sm(sa + A_BYTES, &p.bmap, full_leader[s], kb * 64, cls * p.Nc + n0 + (int)rank * (BN / 2));
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = make_idesc(256, BN, 0, 0, p.ab_bf16, p.ab_bf16);
      uint32_t it = 0, li = 0;
      for (int tile = pair; tile < p.num_tiles; tile += npairs, ++li) {
        const int cls = (tile % inner) / ntn;
        const int nk = p.ntaps[cls] * p.kchunks;
        const uint32_t buf = li & 1, use = li >> 1;
        ptx::mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);     // all 16 epilogue warps of the pair have drained it
        ptx::tc_fence_after();
        const uint32_t acc = tmem_base + buf * ACC_COLS;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % NSTAGE;
          const uint32_t ph = (it / NSTAGE) & 1;
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t sa = ptx::smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16_2sm(acc, desc_kmajor<128>(sa) + 2 * k, desc_kmajor<128>(sa + A_BYTES) + 2 * k, idesc, (kb > 0) || (k != 0));
            ptx::umma_commit_2sm(&empty[s]);
            if (kb == nk - 1) ptx::umma_commit_2sm(&tmem_full[buf]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    const int ew = warp - 2;
    const int q = warp & 3;                         // TMEM lane quadrant this warp may access
    const int half = ew >> 2;                       // the two warps of a quadrant alternate 32-column chunks
    const int r = q * 32 + lane;
    const int wl = r % p.TW, hl = (r / p.TW) % p.TH, nl = r / (p.TW * p.TH);
    const uint32_t tmem_empty_leader[2] = {ptx::mapa_u32(ptx::smem_u32(&tmem_empty[0]), 0), ptx::mapa_u32(ptx::smem_u32(&tmem_empty[1]), 0)};
    float* S = scr + ew * 32 * 33;
    float2* A = sacc + q * F2_STAT_NC;
    const bool stats = p.stats_ws != nullptr;
    uint32_t li = 0;
    for (int tile = pair; tile < p.num_tiles; tile += npairs, ++li) {
      int cls, n0, w0, h0, b0;
      decode(tile, cls, n0, w0, h0, b0);
      const uint32_t buf = li & 1, use = li >> 1;
      const int mw = w0 + wl, mh = h0 + hl, nb = b0 + nl;
      const int oh = mh * p.so + p.oa[cls], ow = mw * p.so + p.ob[cls];
      const bool valid = nb < p.N && mh < p.Hm && mw < p.Wm && oh < p.Hout && ow < p.Wout;
      const int64_t pix = ((int64_t)nb * p.Hout + oh) * p.Wout + ow;
      bf16* dst = p.out + pix * p.out_pitch + p.out_coff + n0;
      ptx::mbar_wait(&tmem_full[buf], use & 1);
      ptx::tc_fence_after();
      const uint32_t acc = tmem_base + buf * ACC_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int c = half * 32; c < BN; c += 64) {
        uint32_t v[32];
        ptx::tmem_ld32(acc + c, v);
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              o[e] = cvt_pair(__uint_as_float(v[j * 8 + 2 * e]), __uint_as_float(v[j * 8 + 2 * e + 1]), p.out_f16);
            *reinterpret_cast<uint4*>(dst + c + j * 8) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
        if (stats) {
          // transpose the warp's 32 rows x 32 columns through shared memory: lane = row writes column-major,
          // lane = column reads its 32 rows; both patterns are bank-conflict free with the 33-float pitch
#pragma unroll
          for (int j = 0; j < 32; ++j) S[j * 33 + lane] = valid ? __uint_as_float(v[j]) : 0.f;
          __syncwarp();
          float sm = 0.f, sq = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) { const float x = S[lane * 33 + i]; sm += x; sq = fmaf(x, x, sq); }
          __syncwarp();
          float2 a = A[n0 + c + lane];
          a.x += sm; a.y += sq;
          A[n0 + c + lane] = a;
        }
      }
      // accumulator buffer fully read into registers: hand it back to the leader's MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(tmem_empty_leader[buf]);
    }
    if (stats) {
      // one partial per CTA (the four quadrants summed in a fixed order): ws[cta][0][c] = sum, [1][c] = sum of squares
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int et = threadIdx.x - 64;
      float* o = p.stats_ws + ((size_t)blockIdx.x * 2) * p.Nc;
      for (int c = et; c < p.Nc; c += 256) {
        const float2 a0 = sacc[c], a1 = sacc[F2_STAT_NC + c], a2 = sacc[2 * F2_STAT_NC + c], a3 = sacc[3 * F2_STAT_NC + c];
        o[c] = (a0.x + a1.x) + (a2.x + a3.x); o[p.Nc + c] = (a0.y + a1.y) + (a2.y + a3.y);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();               // nobody leaves while the peer can still signal its barriers / read its smem
  if (warp == 2) ptx::tmem_dealloc_2sm(tmem_base, TMEM_COLS);
}

static size_t f2_smem_bytes(int BN) {
  return (size_t)f2_stages(BN) * f2_stage_bytes(BN) + 1024 + 256 + (size_t)F2_EPI_WARPS * 32 * 33 * 4 + (size_t)4 * F2_STAT_NC * 8;
}

static size_t fwd_smem_bytes(int BN) { return (size_t)fwd_stages(BN) * (128 * 128 + BN * 128) + 1024 + 256 + (size_t)4 * F1_STAT_NC * 8; }

static bool view_ok(int pitch, int coff, const void* p) {
  return pitch % 8 == 0 && coff % 8 == 0 && ((uintptr_t)p % 16) == 0;
}
static bool chan_ok(int c) { return c == 16 || (c >= 64 && c % 64 == 0); }

bool umma_fwd_supported(const ConvOp& op) {
  if (g_encode == nullptr) return false;
  if (!chan_ok(op.Kc) || !chan_ok(op.Nc)) return false;
  if (op.Kc == 16 && op.Nc == 16) return false;
  if (op.Kc == 16 && op.cls[0].ntaps % 4 != 0) return false;
  if (!view_ok(op.in_pitch, op.in_coff, op.in)) return false;
  if (op.out != nullptr && !view_ok(op.out_pitch, op.out_coff, op.out)) return false;
  if (op.out == nullptr && op.out_f32 == nullptr && op.out_rows_f32 == nullptr) return false;
  if ((op.epi != EPI_NONE || op.out_f32 != nullptr) && op.Nc != 16) return false;
  if (op.si == 2 && (op.Hin % 2 != 0 || op.Win % 2 != 0)) return false;
  return true;
}

static void fill_tap_tables(const ConvOp& op, int8_t (*tap_map)[16], int8_t (*tap_dw)[16], int8_t (*tap_dh)[16]) {
  for (int c = 0; c < op.ncls; ++c) {
    const ClassGeom& g = op.cls[c];
    for (int t = 0; t < g.ntaps; ++t) {
      if (op.in_tap[0] != nullptr) { tap_map[c][t] = (int8_t)(t & 3); tap_dh[c][t] = 0; tap_dw[c][t] = 0; }
      else if (op.si == 1) { tap_map[c][t] = 0; tap_dh[c][t] = g.dh[t]; tap_dw[c][t] = g.dw[t]; }
      else {
        int a = ((g.dh[t] % 2) + 2) % 2, b = ((g.dw[t] % 2) + 2) % 2;
        tap_map[c][t] = (int8_t)(a * 2 + b);
        tap_dh[c][t] = (int8_t)((g.dh[t] - a) / 2); tap_dw[c][t] = (int8_t)((g.dw[t] - b) / 2);
      }
    }
  }
}
static void fill_in_maps(CUtensorMap* maps, const ConvOp& op, int bw, int bh, int bn, int bc) {
  if (op.in_tap[0] != nullptr) {
    // im2col first layer: K-block t is its own [M][64] buffer
    for (int t = 0; t < 4; ++t) {
      const void* base = op.in_tap[t] ? op.in_tap[t] : op.in_tap[0];
      maps[t] = make_map4(base, 64, 0, 64, op.Hin, op.Win, op.N, 1, 0, 0, bw, bh, bn, bc);
    }
  } else if (op.si == 1) {
    maps[0] = make_map4(op.in, op.in_pitch, op.in_coff, op.Kc, op.Hin, op.Win, op.N, 1, 0, 0, bw, bh, bn, bc);
    for (int i = 1; i < 4; ++i) maps[i] = maps[0];
  } else {
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        maps[a * 2 + b] = make_map4(op.in, op.in_pitch, op.in_coff, op.Kc, op.Hin, op.Win, op.N, 2, a, b, bw, bh, bn, bc);
  }
}

static void fill_fwd_params(UmmaFwdParams& P, const ConvOp& op, int bn_tile, int kc) {
  memset(&P, 0, sizeof(P));
  int TW = pow2ceil(op.Wm); if (TW > 128) TW = 128;
  int TH = pow2ceil(op.Hm); if (TH > 128 / TW) TH = 128 / TW;
  int TN = 128 / (TW * TH);
  P.TW = TW; P.TH = TH; P.TN = TN;
  P.tiles_w = (op.Wm + TW - 1) / TW; P.tiles_h = (op.Hm + TH - 1) / TH; P.tiles_n = (op.N + TN - 1) / TN;
  fill_in_maps(P.amap, op, TW, TH, TN, kc);
  int64_t Ktot = (int64_t)op.cls[0].ntaps * op.Kc;
  P.bmap = make_map2(op.B, Ktot, (int64_t)op.ncls * op.Nc, bn_tile, kc);
  fill_tap_tables(op, P.tap_map, P.tap_dw, P.tap_dh);
  for (int c = 0; c < op.ncls; ++c) { P.ntaps[c] = op.cls[c].ntaps; P.oa[c] = op.cls[c].oa; P.ob[c] = op.cls[c].ob; }
  P.kchunks = op.Kc / 64;
  P.out = (bf16*)op.out; P.out_pitch = op.out_pitch; P.out_coff = op.out_coff; P.Hout = op.Hout; P.Wout = op.Wout; P.so = op.so;
  P.N = op.N; P.Hm = op.Hm; P.Wm = op.Wm; P.Nc = op.Nc;
  P.bias = op.bias; P.out_f32 = op.out_f32; P.epi = op.epi; P.Nr = op.Nr; P.ncls = op.ncls;
  P.ab_bf16 = op.dt_in == DT_F16 ? 0 : 1; P.out_f16 = op.dt_out == DT_F16 ? 1 : 0;
}

// dev A/B switches (read once): GAN_B200_2CTA=0 keeps every layer on the one-CTA kernel, GAN_B200_EPI_STATS=0 keeps the
// separate statistics pass
static bool env_on(const char* name) { const char* e = getenv(name); return !(e && e[0] == '0'); }
static const bool g_use_2cta = env_on("GAN_B200_2CTA");
static const bool g_epi_stats = env_on("GAN_B200_EPI_STATS");
// smallest MMA N the CTA-pair kernel takes (measured: below N = 256 one CTA pair per TPC loses to two independent CTAs
// per SM, whose k-block round trips overlap)
static const int g_2cta_min_bn = [] { const char* e = getenv("GAN_B200_2CTA_MIN_N"); return e ? atoi(e) : 256; }();
// CTA pairs that can be co-resident (cudaOccupancyMaxActiveClusters, filled by umma_init): a B200 has 148 of the die's
// SMs enabled, and a TPC with one fused-off SM cannot host a pair, so this can be below 74.  The persistent grid must
// not exceed it: a pair that waits for a free TPC would run its whole share of tiles after everybody else.
static int g_f2_pairs[3] = {0, 0, 0};      // BN = 64, 128, 256
static int f2_idx(int BN) { return BN == 64 ? 0 : (BN == 128 ? 1 : 2); }
int umma2_max_pairs(int BN) { return g_f2_pairs[f2_idx(BN)]; }

// CTA-pair path: returns true when it launched (and *stat_parts = number of statistics partials written).
static bool launch_conv_fwd_umma2(Launch L, const ConvOp& op, int* stat_parts) {
  *stat_parts = 0;
  if (!g_use_2cta) return false;
  if (op.Kc % 64 != 0 || op.Nc % 64 != 0 || op.out == nullptr) return false;
  if (op.epi != EPI_NONE || op.out_f32 != nullptr || op.out_rows_f32 != nullptr) return false;
  const int BN = (op.Nc % 256 == 0) ? 256 : (op.Nc % 128 == 0 ? 128 : 64);
  if (BN < g_2cta_min_bn) return false;
  UmmaFwdParams P;
  fill_fwd_params(P, op, BN / 2, 64);                 // each CTA of the pair loads half of the weight tile's rows
  const int mtiles = P.tiles_w * P.tiles_h * P.tiles_n;
  if (mtiles < 148) return false;                     // small-M layers: split-K on the one-CTA kernel fills the chip better
  P.num_tiles = ((mtiles + 1) / 2) * (op.Nc / BN) * op.ncls;     // pair tiles
  P.ksplit = 1; P.ws = nullptr; P.f32out = 0;
  P.stats_ws = (g_epi_stats && op.stats_ws != nullptr && op.Nc <= F2_STAT_NC) ? op.stats_ws : nullptr;
  const int max_pairs = g_f2_pairs[f2_idx(BN)];
  if (max_pairs < 37) return false;                    // clusters not schedulable on (most of) this device
  const int pairs = P.num_tiles < max_pairs ? P.num_tiles : max_pairs;        // persistent: one CTA pair per usable TPC
  dim3 grid(2 * pairs);
  const size_t sm = f2_smem_bytes(BN);
  if (BN == 256) k_conv_fwd_umma2<256><<<grid, F2_THREADS, sm, L.s>>>(P);
  else if (BN == 128) k_conv_fwd_umma2<128><<<grid, F2_THREADS, sm, L.s>>>(P);
  else k_conv_fwd_umma2<64><<<grid, F2_THREADS, sm, L.s>>>(P);
  KLAUNCH(L);
  if (P.stats_ws != nullptr) *stat_parts = (int)grid.x;
  return true;
}

int launch_conv_fwd_umma(Launch L, const ConvOp& op) {
  int stat_parts = 0;
  if (launch_conv_fwd_umma2(L, op, &stat_parts)) return stat_parts;
  int BN = (op.Nc % 128 == 0) ? 128 : (op.Nc % 64 == 0 ? 64 : 16);
  const int KC = op.Kc == 16 ? 16 : 64;
  UmmaFwdParams P;
  fill_fwd_params(P, op, BN, KC);
  // N = 256 tiles (25% less L2->SM operand traffic per FLOP) when there are enough tiles to fill the chip
  if (KC == 64 && op.Nc % 256 == 0 && P.tiles_w * P.tiles_h * P.tiles_n * (op.Nc / 256) * op.ncls >= 148) {
    BN = 256;
    fill_fwd_params(P, op, BN, KC);
  }
  P.num_tiles = P.tiles_w * P.tiles_h * P.tiles_n * (op.Nc / BN) * op.ncls;
  P.ksplit = 1; P.ws = nullptr; P.f32out = 0;
  if (op.out_rows_f32 != nullptr) { P.ws = op.out_rows_f32; P.f32out = 1; P.slab = 0; }
  // split-K for layers with too few output tiles to fill the chip (the 1x1..8x8 bottleneck layers:
  // a serial 128-k-block loop on 4..128 CTAs is pure TMA->MMA latency): k-ranges go to separate CTAs,
  // fp32 atomics into a workspace, one conversion pass to bf16.
  const int nk = (KC == 64) ? op.cls[0].ntaps * (op.Kc / 64) : op.cls[0].ntaps / 4;
  const size_t out_elems = (size_t)op.N * op.Hout * op.Wout * op.Nc;
  if (op.splitk_ws != nullptr && BN >= 64 && BN <= 128 && P.num_tiles <= 74 && op.epi == EPI_NONE && op.out_f32 == nullptr &&
      op.out_rows_f32 == nullptr) {
    int ks = (2 * 148) / P.num_tiles;
    if (ks > nk / 8) ks = nk / 8;
    while (ks >= 2 && out_elems * 4 * ks > op.splitk_ws_bytes) --ks;
    if (ks >= 2) { P.ksplit = ks; P.ws = op.splitk_ws; P.slab = (long long)out_elems; P.num_tiles *= ks; }
  }
  const int per_sm = BN == 256 ? 1 : 2;
  dim3 grid(P.num_tiles < per_sm * 148 ? P.num_tiles : per_sm * 148);     // persistent
  static const bool stage_on = [] { const char* e = getenv("GAN_B200_STAGED_EPI"); return !(e && e[0] == '0'); }();   // dev A/B switch
  P.stage_out = stage_on ? 1 : 0;
  // BatchNorm statistics from the fp32 accumulators (one partial per CTA) when every tile holds complete sums
  P.stats_ws = (g_epi_stats && op.stats_ws != nullptr && op.Nc <= F1_STAT_NC && BN >= 64 && P.ksplit == 1 && !P.f32out &&
                op.epi == EPI_NONE && op.out_f32 == nullptr) ? op.stats_ws : nullptr;
  const size_t sm = fwd_smem_bytes(BN);
  if (KC == 64) {
    if (BN == 256) k_conv_fwd_umma<256, 64><<<grid, FWD_THREADS, sm, L.s>>>(P);
    else if (BN == 128) k_conv_fwd_umma<128, 64><<<grid, FWD_THREADS, sm, L.s>>>(P);
    else if (BN == 64) k_conv_fwd_umma<64, 64><<<grid, FWD_THREADS, sm, L.s>>>(P);
    else k_conv_fwd_umma<16, 64><<<grid, FWD_THREADS, sm, L.s>>>(P);
  } else {
    if (BN == 128) k_conv_fwd_umma<128, 16><<<grid, FWD_THREADS, sm, L.s>>>(P);
    else k_conv_fwd_umma<64, 16><<<grid, FWD_THREADS, sm, L.s>>>(P);
  }
  KLAUNCH(L);
  if (P.ksplit > 1)   // deterministic reduction of the k-split slabs + conversion to bf16
    launch_sum_slabs(L, op.dt_out, op.splitk_ws, P.ksplit, (int64_t)op.N * op.Hout * op.Wout, op.Nc, op.out, op.out_pitch, op.out_coff);
  return P.stats_ws != nullptr ? (int)grid.x : 0;
}

// =============================================================================================
// weight-gradient kernel.  Template: BN = N tile over dY channels (128 | 64 | 16),
// KCA = channels per A box (64 | 16).  A rows = 128 consecutive entries of the (tap, channel)
// index space = 128/KCA boxes; both operands MN-major, GEMM-K = pixels.
// =============================================================================================
struct alignas(64) UmmaWgradParams {
  CUtensorMap amap[4];     // input activation boxes (KCA ch x PW x PH x PN), parity sub-lattices when si == 2
  CUtensorMap dmap[4];     // dY boxes, parity sub-lattices when so == 2
  int8_t tap_map[4][16], tap_dw[4][16], tap_dh[4][16], widx[4][16];
  int ntaps[4], dy_map[4];
  int Kc;                  // (padded) channels per tap
  int PW, PH, PN, tiles_w, tiles_h, tiles_n;   // 64-pixel boxes over the M-space
  int splits;
  float* dW; long long s_tap, s_k, s_n;
  int Nc, Kr, Nr;
  int im2col_c;            // >0: rows are im2col K-blocks (t = source, kc = tap16*4 + channel slot)
  int n_slot4_c;           // >0: columns are (tap16*4 + channel slot) of a cols matrix
  int a_bf16, b_bf16;      // operand formats (always equal: tcgen05 kind::f16 faults on mixed f16 / bf16 operands)
  // Deterministic reduction.  splits > 1: every CTA stores its fp32 partial tile into slab[item][128][BN] (plain
  // stores, item = ((cls*mblocks + mblock)*ntiles + ntile)*splits + split) and k_wgrad_reduce sums the `splits`
  // tiles of an output tile in a fixed order.  splits == 1 (slab == nullptr): the CTA is the only writer of its
  // tile and stores (accumulate == 0) or adds (accumulate == 1, non-atomic read-modify-write) straight into dW.
  float* slab; int accumulate, mblocks, ntiles;
};

constexpr int WG_STAGES = 3;
constexpr int WG_PIX = 64;            // pixels (GEMM-K) per stage

template <int BN, int KCA>
__global__ void __launch_bounds__(FWD_THREADS) k_conv_wgrad_umma(const __grid_constant__ UmmaWgradParams p) {
  constexpr int SWA = (KCA == 64) ? 128 : 32;
  constexpr int SWB = (BN >= 64) ? 128 : 32;
  constexpr int NA = 128 / KCA;                               // A boxes per stage
  constexpr int NB = (BN >= 64) ? BN / 64 : 1;                // dY boxes per stage
  constexpr uint32_t A_BOX = WG_PIX * KCA * 2;
  constexpr uint32_t B_BOX = WG_PIX * ((BN >= 64) ? 64 : 16) * 2;
  constexpr uint32_t A_BYTES = NA * A_BOX;                    // 16 KB
  constexpr uint32_t B_BYTES = NB * B_BOX;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
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
  const int kbase = blockIdx.x * 128;                         // first (tap,channel) index of this M tile
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
  if (warp == 2) ptx::tmem_alloc(tmem_slot, TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (nk > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int t_[NA], c_[NA];
#pragma unroll
        for (int i = 0; i < NA; ++i) { int k = kbase + i * KCA; t_[i] = k / p.Kc; c_[i] = k - t_[i] * p.Kc; }
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
          for (int i = 0; i < NA; ++i)
            ptx::tma_load_4d(sa + i * A_BOX, &p.amap[p.tap_map[cls][t_[i]]], &full[s], c_[i],
                             w0 + p.tap_dw[cls][t_[i]], h0 + p.tap_dh[cls][t_[i]], b0);
#pragma unroll
          for (int i = 0; i < NB; ++i)
            ptx::tma_load_4d(sa + A_BYTES + i * B_BOX, &p.dmap[p.dy_map[cls]], &full[s], n0 + i * 64, w0, h0, b0);
        }
      }
    } else if (warp == 1) {
      const uint32_t idesc = make_idesc(128, BN, 1, 1, p.a_bf16, p.b_bf16);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % WG_STAGES;
        const uint32_t ph = (kb / WG_STAGES) & 1;
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t sa = ptx::smem_u32(smem + s * STAGE_BYTES);
          const uint64_t ad = desc_mnmajor<SWA>(sa, A_BOX), bd = desc_mnmajor<SWB>(sa + A_BYTES, B_BOX);
#pragma unroll
          for (int k = 0; k < WG_PIX / 16; ++k)   // 16 pixels = two 8-row atoms per MMA
            ptx::umma_bf16(tmem_base, ad + (uint64_t)((k * 16 * SWA) >> 4), bd + (uint64_t)((k * 16 * SWB) >> 4), idesc,
                           (kb | k) != 0);
          ptx::umma_commit(&empty[s]);
          if (kb == nk - 1) ptx::umma_commit(tmem_full);
        }
        __syncwarp();
      }
    } else {
      const int q = warp & 3;
      const int r = q * 32 + lane;                   // accumulator row = (tap,channel) index inside the tile
      const int k = kbase + r;
      const int t = k / p.Kc, kc = k - t * p.Kc;
      bool row_ok = kc < p.Kr && t < p.ntaps[cls];
      float* dst = p.dW + (long long)p.widx[cls][t & 15] * p.s_tap + (long long)kc * p.s_k + (long long)n0 * p.s_n;
      if (p.im2col_c > 0) {
        const int ic = p.im2col_c;
        row_ok = t < p.ntaps[cls] && (kc & 3) < ic;
        dst = p.dW + (long long)(kc >> 2) * p.s_tap + (long long)(t * ic + (kc & 3)) * p.s_k + (long long)n0 * p.s_n;
      }
      const bool vec_ok = p.s_n == 1 && p.n_slot4_c == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
      const bool acc_dst = p.accumulate != 0;
      ptx::mbar_wait(tmem_full, 0);
      ptx::tc_fence_after();
      if (p.slab != nullptr) {
        // partial tile -> slab (plain 16-byte stores; every CTA owns its tile)
        const long long item = (((long long)cls * p.mblocks + blockIdx.x) * p.ntiles + blockIdx.y) * p.splits + split;
        float* sl = p.slab + (item * 128 + r) * (BN < 32 ? 16 : BN);
        if (BN >= 32) {
#pragma unroll
          for (int c = 0; c < BN; c += 32) {
            uint32_t v[32];
            ptx::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, v);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<uint4*>(sl + c + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        } else {
          uint32_t v[16];
          ptx::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<uint4*>(sl + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      } else if (BN >= 32) {
#pragma unroll
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          ptx::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, v);
          if (row_ok && vec_ok && n0 + c + 32 <= p.Nr) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              // this CTA is the only writer of these elements in this launch and launches are stream-ordered, so a
              // fire-and-forget reduction is deterministic here (no read latency in the epilogue)
              if (acc_dst)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c + j), "r"(v[j]), "r"(v[j + 1]),
                             "r"(v[j + 2]), "r"(v[j + 3]) : "memory");
              else
                *reinterpret_cast<uint4*>(dst + c + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          } else if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int nn = n0 + c + j;
              float* d1 = nullptr;
              if (p.n_slot4_c > 0) {      // cols column (tap*4 + slot) -> master row tap*C + slot
                if ((nn & 3) < p.n_slot4_c) d1 = p.dW + (long long)kc * p.s_k + (long long)((nn >> 2) * p.n_slot4_c + (nn & 3)) * p.s_n;
              } else if (nn < p.Nr) d1 = dst + (long long)(c + j) * p.s_n;
              if (d1 != nullptr) { if (acc_dst) atomicAdd(d1, __uint_as_float(v[j])); else *d1 = __uint_as_float(v[j]); }
            }
          }
        }
      } else {
        uint32_t v[16];
        ptx::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16), v);
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < p.Nr) { float* d1 = dst + (long long)j * p.s_n; if (acc_dst) atomicAdd(d1, __uint_as_float(v[j])); else *d1 = __uint_as_float(v[j]); }
        }
      }
    }
  } else if (p.slab != nullptr && warp >= 2) {
    // empty pixel range (more splits than boxes cannot happen, kept for safety): the reduction still reads this tile
    const int q = warp & 3, r = q * 32 + lane;
    const long long item = (((long long)cls * p.mblocks + blockIdx.x) * p.ntiles + blockIdx.y) * p.splits + split;
    float* sl = p.slab + (item * 128 + r) * (BN < 32 ? 16 : BN);
    for (int j = 0; j < (BN < 32 ? 16 : BN); ++j) sl[j] = 0.f;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}


// Second stage of the deterministic weight-gradient reduction: one block per output tile sums the `splits` slab tiles
// in a fixed order and scatters the result into the master (TF) weight layout, storing (first contribution of the
// step) or adding (later contributions: the second discriminator call, CycleGAN's three generator calls).
// LANES threads share one float4 of the output: lane l sums the partial tiles l, l + LANES, ... (eight loads in flight),
// the lanes are then combined by a fixed xor-butterfly — the order of every addition is a function of (splits, LANES)
// only, so the result is deterministic — and lane 0 writes.  LANES grows with the number of splits: the first-layer
// and head gradients arrive as 148-296 partial tiles of ONE output tile, which a single lane per output would sum serially
// on 8 blocks (measured: 28-54 us for 32 KB of output).
template <int LANES>
__global__ void __launch_bounds__(256) k_wgrad_reduce(const WgradReduceParams p) {
  // grid = (output tiles) x (sub-tiles of 1024 / LANES elements).  Sub-tiles are row strips when the master layout is
  // contiguous along the output channel, column strips (rows x 8 columns) when it is contiguous along the row index,
  // so that a warp's stores form long runs.
  constexpr int OUT4 = 256 / LANES;                  // float4 outputs per block
  const int bn = p.bn;
  const int nsub = 128 * bn / (4 * OUT4);
  const int sub = blockIdx.x % nsub;
  int b = blockIdx.x / nsub;
  const long long tile_idx = b;
  const int ntile = b % p.ntiles; b /= p.ntiles;
  const int mblock = b % p.mblocks; const int cls = b / p.mblocks;
  const bool along_n = p.s_n == 1;
  const int o = threadIdx.x / LANES, lane = threadIdx.x % LANES;
  int r, c;
  if (along_n) { const int e = (sub * OUT4 + o) * 4; r = e / bn; c = e % bn; }
  else { const int q = sub * OUT4 + o; r = q % 128; c = (q / 128) * 4; }      // consecutive outputs walk down a 4-column strip
  const int total = 128 * bn;
  const float* base = p.slab + (tile_idx * p.splits) * total + (long long)r * bn + c;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int sp = lane; sp < p.splits; sp += 8 * LANES) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      v[u] = sp + u * LANES < p.splits ? __ldcs(reinterpret_cast<const float4*>(base + (long long)(sp + u * LANES) * total)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 8; ++u) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
  }
#pragma unroll
  for (int s = LANES / 2; s >= 1; s >>= 1) {
    a.x += __shfl_xor_sync(0xffffffffu, a.x, s); a.y += __shfl_xor_sync(0xffffffffu, a.y, s);
    a.z += __shfl_xor_sync(0xffffffffu, a.z, s); a.w += __shfl_xor_sync(0xffffffffu, a.w, s);
  }
  if (lane != 0) return;
  const int k = mblock * 128 + r;
  const int t = k / p.Kc, kc = k - t * p.Kc;
  if (t >= p.ntaps[cls]) return;
  const float vals[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int nn = ntile * bn + c + e;
    long long off;
    if (p.im2col_c > 0) {
      if ((kc & 3) >= p.im2col_c || nn >= p.Nr) continue;
      off = (long long)(kc >> 2) * p.s_tap + (long long)(t * p.im2col_c + (kc & 3)) * p.s_k + (long long)nn * p.s_n;
    } else if (p.n_slot4_c > 0) {
      if (kc >= p.Kr || (nn & 3) >= p.n_slot4_c) continue;
      off = (long long)kc * p.s_k + (long long)((nn >> 2) * p.n_slot4_c + (nn & 3)) * p.s_n;
    } else {
      if (kc >= p.Kr || nn >= p.Nr) continue;
      off = (long long)p.widx[cls][t & 15] * p.s_tap + (long long)kc * p.s_k + (long long)nn * p.s_n;
    }
    if (p.accumulate) atomicAdd(p.dW + off, vals[e]); else p.dW[off] = vals[e];      // single writer per element: deterministic
  }
}

void launch_wgrad_reduce(Launch L, const WgradReduceParams& R, long long out_tiles) {
  const long long out4 = out_tiles * 128 * R.bn / 4;                 // float4 outputs
  if (R.splits >= 64) k_wgrad_reduce<32><<<(unsigned)(out4 / 8), 256, 0, L.s>>>(R);
  else if (R.splits >= 8) k_wgrad_reduce<4><<<(unsigned)(out4 / 64), 256, 0, L.s>>>(R);
  else k_wgrad_reduce<1><<<(unsigned)(out4 / 256), 256, 0, L.s>>>(R);
  KLAUNCH(L);
}

static size_t wg_smem_bytes(int BN) {
  size_t b = BN >= 64 ? (size_t)(BN / 64) * WG_PIX * 128 : (size_t)WG_PIX * 32;
  return (size_t)WG_STAGES * (16384 + b) + 1024 + 128;
}

bool umma_wgrad_supported(const ConvOp& op) {
  if (g_encode == nullptr) return false;
  if (!chan_ok(op.Kc) || !chan_ok(op.Nc)) return false;
  if (op.Kc == 16 && op.Nc == 16) return false;
  if ((op.cls[0].ntaps * op.Kc) % 128 != 0 && op.in_tap[0] == nullptr) return false;
  if (!view_ok(op.in_pitch, op.in_coff, op.in) || !view_ok(op.out_pitch, op.out_coff, op.out)) return false;
  if (op.si == 2 && (op.Hin % 2 != 0 || op.Win % 2 != 0)) return false;
  if (op.so == 2 && (op.Hout % 2 != 0 || op.Wout % 2 != 0)) return false;
  return true;
}

void launch_conv_wgrad_umma(Launch L, const ConvOp& op) {
  const int BN = (op.Nc % 128 == 0) ? 128 : (op.Nc % 64 == 0 ? 64 : 16);
  const int KCA = op.Kc == 16 ? 16 : 64;
  UmmaWgradParams P; memset(&P, 0, sizeof(P));
  int PW = pow2ceil(op.Wm); if (PW > WG_PIX) PW = WG_PIX;
  int PH = pow2ceil(op.Hm); if (PH > WG_PIX / PW) PH = WG_PIX / PW;
  int PN = WG_PIX / (PW * PH);
  P.PW = PW; P.PH = PH; P.PN = PN;
  P.tiles_w = (op.Wm + PW - 1) / PW; P.tiles_h = (op.Hm + PH - 1) / PH; P.tiles_n = (op.N + PN - 1) / PN;
  fill_in_maps(P.amap, op, PW, PH, PN, KCA);
  const int bcn = BN >= 64 ? 64 : 16;
  if (op.so == 1) {
    P.dmap[0] = make_map4(op.out, op.out_pitch, op.out_coff, op.Nc, op.Hout, op.Wout, op.N, 1, 0, 0, PW, PH, PN, bcn);
    for (int i = 1; i < 4; ++i) P.dmap[i] = P.dmap[0];
  } else {
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        P.dmap[a * 2 + b] = make_map4(op.out, op.out_pitch, op.out_coff, op.Nc, op.Hout, op.Wout, op.N, 2, a, b, PW, PH, PN, bcn);
  }
  fill_tap_tables(op, P.tap_map, P.tap_dw, P.tap_dh);
  for (int c = 0; c < op.ncls; ++c) {
    const ClassGeom& g = op.cls[c];
    P.ntaps[c] = g.ntaps;
    P.dy_map[c] = op.so == 1 ? 0 : g.oa * 2 + g.ob;
    for (int t = 0; t < g.ntaps; ++t) P.widx[c][t] = g.widx[t];
  }
  P.Kc = op.Kc;
  P.dW = op.dW; P.s_tap = op.s_tap; P.s_k = op.s_k; P.s_n = op.s_n; P.Nc = op.Nc; P.Kr = op.Kr; P.Nr = op.Nr;
  P.im2col_c = op.in_tap[0] != nullptr ? op.im2col_c : 0;
  P.n_slot4_c = op.n_slot4_c;
  P.a_bf16 = op.dt_in == DT_F16 ? 0 : 1; P.b_bf16 = op.dt_out == DT_F16 ? 0 : 1;
  const int ntaps = op.cls[0].ntaps;
  const int mblocks = (ntaps * op.Kc + 127) / 128;
  const int ntiles = op.Nc / BN;
  const int total_boxes = P.tiles_w * P.tiles_h * P.tiles_n;
  // split the pixel reduction so that all CTAs form ONE full wave (2 CTAs/SM): equal-length CTAs
  // make any partial second wave pure tail
  int64_t ctas = (int64_t)mblocks * ntiles * op.ncls;
  int splits = (int)((148 * 2) / ctas);
  // at least 16 pipeline stages per CTA: below that the prologue (barriers, TMEM allocation) and the 64 KB partial-tile
  // epilogue outweigh the main loop (small per-GPU batches)
  if (splits > total_boxes / 16) splits = total_boxes / 16;
  if (splits < 1) splits = 1;
  const int bn_slab = BN < 32 ? 16 : BN;
  if (splits > 1 && (op.wgrad_ws == nullptr || (size_t)ctas * splits * 128 * bn_slab * 4 > op.wgrad_ws_bytes)) splits = 1;
  P.splits = splits;
  P.mblocks = mblocks; P.ntiles = ntiles;
  P.slab = splits > 1 ? op.wgrad_ws : nullptr;
  P.accumulate = op.accumulate;
  dim3 grid(mblocks, ntiles, op.ncls * splits);
  const size_t sm = wg_smem_bytes(BN);
  if (KCA == 64) {
    if (BN == 128) k_conv_wgrad_umma<128, 64><<<grid, FWD_THREADS, sm, L.s>>>(P);
    else if (BN == 64) k_conv_wgrad_umma<64, 64><<<grid, FWD_THREADS, sm, L.s>>>(P);
    else k_conv_wgrad_umma<16, 64><<<grid, FWD_THREADS, sm, L.s>>>(P);
  } else {
    if (BN == 128) k_conv_wgrad_umma<128, 16><<<grid, FWD_THREADS, sm, L.s>>>(P);
    else k_conv_wgrad_umma<64, 16><<<grid, FWD_THREADS, sm, L.s>>>(P);
  }
  KLAUNCH(L);
  if (splits > 1) {
    WgradReduceParams R; memset(&R, 0, sizeof(R));
    R.slab = op.wgrad_ws; R.dW = op.dW; R.splits = splits; R.bn = bn_slab; R.mblocks = mblocks; R.ntiles = ntiles; R.ncls = op.ncls;
    R.accumulate = op.accumulate; R.Kc = op.Kc; R.Kr = op.Kr; R.Nr = op.Nr; R.im2col_c = P.im2col_c; R.n_slot4_c = op.n_slot4_c;
    R.s_tap = op.s_tap; R.s_k = op.s_k; R.s_n = op.s_n;
    for (int c = 0; c < op.ncls; ++c) { R.ntaps[c] = op.cls[c].ntaps; for (int t = 0; t < op.cls[c].ntaps; ++t) R.widx[c][t] = op.cls[c].widx[t]; }
    launch_wgrad_reduce(L, R, ctas);
  }
}

void umma_init() {
  static std::once_flag once;
  std::call_once(once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = (PFN_encodeTiled)fn;
    else cudaGetLastError();
#define SET_SMEM(K, B) cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(B))
    SET_SMEM((k_conv_fwd_umma2<256>), f2_smem_bytes(256));
    SET_SMEM((k_conv_fwd_umma2<128>), f2_smem_bytes(128));
    SET_SMEM((k_conv_fwd_umma2<64>), f2_smem_bytes(64));
    auto query = [](const void* kern, int BN) {
      cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(148); cfg.blockDim = dim3(F2_THREADS); cfg.dynamicSmemBytes = f2_smem_bytes(BN);
      cudaLaunchAttribute at; memset(&at, 0, sizeof(at));
      at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
      cfg.attrs = &at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
      g_f2_pairs[f2_idx(BN)] = n > 74 ? 74 : n;
    };
    query((const void*)k_conv_fwd_umma2<256>, 256);
    query((const void*)k_conv_fwd_umma2<128>, 128);
    query((const void*)k_conv_fwd_umma2<64>, 64);
    if (getenv("GAN_B200_DEBUG")) fprintf(stderr, "[gan_b200] active CTA pairs: N=64 %d, N=128 %d, N=256 %d\n", g_f2_pairs[0], g_f2_pairs[1], g_f2_pairs[2]);
    SET_SMEM((k_conv_fwd_umma<256, 64>), fwd_smem_bytes(256));
    SET_SMEM((k_conv_fwd_umma<128, 64>), fwd_smem_bytes(128));
    SET_SMEM((k_conv_fwd_umma<64, 64>), fwd_smem_bytes(64));
    SET_SMEM((k_conv_fwd_umma<16, 64>), fwd_smem_bytes(16));
    SET_SMEM((k_conv_fwd_umma<128, 16>), fwd_smem_bytes(128));
    SET_SMEM((k_conv_fwd_umma<64, 16>), fwd_smem_bytes(64));
    SET_SMEM((k_conv_wgrad_umma<128, 64>), wg_smem_bytes(128));
    SET_SMEM((k_conv_wgrad_umma<64, 64>), wg_smem_bytes(64));
    SET_SMEM((k_conv_wgrad_umma<16, 64>), wg_smem_bytes(16));
    SET_SMEM((k_conv_wgrad_umma<128, 16>), wg_smem_bytes(128));
    SET_SMEM((k_conv_wgrad_umma<64, 16>), wg_smem_bytes(64));
#undef SET_SMEM
  });
}
