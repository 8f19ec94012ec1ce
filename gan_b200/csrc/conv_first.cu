// conv_first.cu — the image-channel first layers (G.down1 base_gan.py:180, D.down1 base_gan.py:141: Conv2D 4x4 s2
// 'same', Cin in {1,2,3,4} per source image, 64 filters, no normalisation, LeakyReLU 0.3) without im2col rows in HBM.
//
// The 4x4 stride-2 window of a 3-channel image is 48 values per output pixel; as a tcgen05 operand it is one
// 128-byte row k = tap*4 + channel slot (64 x 16 bit, slots >= C zero).  Round 1 materialised those rows in HBM
// (134 MB per image batch at 64 x 256^2, written by k_im2col, read by the GEMM) and then ran a separate activation pass
// over z.  Here a CTA builds the rows in SHARED memory, directly in the SWIZZLE_128B K-major layout the MMA reads:
//   * a group of 4 warps owns one 8 x 16 tile of output pixels: it pulls the 18 x 34 input patch (fp32, coalesced)
//     into shared memory, every thread assembles the row of its output pixel (16 taps x 4 slots -> eight swizzled
//     16-byte stores), fences the async proxy and signals the MMA warp;
//   * warp 8 issues 4 (x sources) tcgen05.mma M=128 x N=64 x K=16 against the stationary weight tile (one TMA load per
//     CTA) into the group's TMEM columns and commits to the group's barrier;
//   * the same 4 warps drain TMEM: a = LeakyReLU(z) goes through shared memory into the consumer view (skip-concat
//     buffer) with 8 lanes per 128-byte row.  z itself is not stored: LeakyReLU keeps the sign, so the backward pass
//     takes the activation derivative from a (launch_norm_bwd with a strided z view).
// Two groups per CTA alternate tiles, so the gather / store phases of one overlap the other's.  HBM traffic per image
// batch: the fp32 image(s) once + a — the roofline of this layer is HBM, and the kernel moves nothing else.
#include <cuda.h>
#include <cstring>
#include "kernels.h"

#define KLAUNCH(L) (++*(L).count)

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar), done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
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
// K-major SWIZZLE_128B shared-memory descriptor (see conv_umma.cu): 8-row atoms of 128-byte rows, SBO = 1024
__device__ __forceinline__ uint64_t desc_k128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t idesc_f16(int M, int N, int ab_bf16) {
  return (1u << 4) | ((uint32_t)ab_bf16 << 7) | ((uint32_t)ab_bf16 << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int FT_W = 16, FT_H = 8;                 // output tile: 8 rows x 16 columns = 128 pixels = the MMA's M
constexpr int PATCH_H = 2 * FT_H + 2, PATCH_W = 2 * FT_W + 2;      // 18 x 34 input pixels
constexpr int FIRST_THREADS = 288;                 // 2 groups x 4 warps + MMA warp

struct alignas(64) FirstFwdParams {
  CUtensorMap bmap;                  // packed weights [64][nsrc*64], K-major
  const float* src[2];               // fp32 NHWC images (B, H, W, C)
  int nsrc, C, B, H, W;              // H, W: input size; output grid H/2 x W/2
  bf16* a; int a_pitch, a_coff;      // LeakyReLU(z) into the consumer view (16-bit)
  int tiles_w, tiles_h, num_tiles;
  int ab_bf16, out_f16;
};

constexpr int PATCH_PITCH = PATCH_W * 8;                            // bytes per patch row: 34 pixels x 4 slots x 16 bit
constexpr int PATCH_BYTES = ((PATCH_H * PATCH_PITCH + 127) / 128) * 128;

template <int NSRC>
__global__ void __launch_bounds__(FIRST_THREADS) k_conv_first_fwd(const __grid_constant__ FirstFwdParams p) {
  constexpr uint32_t A_BYTES = 128 * 128;                         // one source: 128 rows x 64 x 16 bit
  constexpr uint32_t GROUP_A = NSRC * A_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* atile = smem;                                          // [2 groups][NSRC][128][128 B]
  uint8_t* wsm = smem + 2 * GROUP_A;                              // [NSRC][64 rows][128 B]
  uint8_t* patch = wsm + NSRC * 64 * 128;                         // [2 groups][NSRC][18][34][4 slots] 16-bit
  uint64_t* bars = (uint64_t*)(patch + 2 * NSRC * PATCH_BYTES);   // full[2], done[2], wbar
  uint32_t* tmem_slot = (uint32_t*)(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_init(&bars[3], 1); mbar_init(&bars[4], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int Ho = p.H / 2, Wo = p.W / 2;

  if (warp == 8) {
    // ---- MMA issuer: stationary weights, then one k-block (NSRC x 4 MMAs) per tile ----
    if (lane == 0) {
      mbar_expect_tx(&bars[4], NSRC * 64 * 128);
      for (int s = 0; s < NSRC; ++s) tma_load_2d(wsm + s * 64 * 128, &p.bmap, &bars[4], s * 64, 0);
    }
    mbar_wait(&bars[4], 0);
    const uint32_t idesc = idesc_f16(128, 64, p.ab_bf16);
    uint32_t n[2] = {0, 0};
    for (int t = blockIdx.x * 2; t < p.num_tiles; t += gridDim.x * 2) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (t + g >= p.num_tiles) break;
        mbar_wait(&bars[g], n[g] & 1);                              // the group has built its A tile
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t sa = smem_u32(atile + g * GROUP_A), sb = smem_u32(wsm);
#pragma unroll
          for (int s = 0; s < NSRC; ++s)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = desc_k128(sa + s * A_BYTES) + 2 * k, bd = desc_k128(sb + s * 64 * 128) + 2 * k;
              const uint32_t acc = (s | k) != 0;
              asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                           ::"r"(tmem_base + g * 64), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
            }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[2 + g])) : "memory");
        }
        __syncwarp();
        ++n[g];
      }
    }
  } else {
    // ---- gather / epilogue groups ----
    const int g = warp >> 2;                       // group 0: warps 0..3, group 1: warps 4..7
    const int q = warp & 3;                        // TMEM lane quadrant of this warp
    const int pl = (threadIdx.x & 127);            // output pixel inside the tile == accumulator row
    const int olh = pl / FT_W, olw = pl % FT_W;
    uint8_t* my_patch = patch + g * NSRC * PATCH_BYTES;
    uint8_t* my_a = atile + g * GROUP_A;
    const int bar_id = 1 + g;                      // named barrier of the group (128 threads)
    uint32_t n = 0;
    for (int t = blockIdx.x * 2 + g; t < p.num_tiles; t += gridDim.x * 2, ++n) {
      int tt = t;
      const int tw = tt % p.tiles_w; tt /= p.tiles_w;
      const int th = tt % p.tiles_h; const int b = tt / p.tiles_h;
      const int oh0 = th * FT_H, ow0 = tw * FT_W;
      const int ih0 = 2 * oh0 - 1, iw0 = 2 * ow0 - 1;
      // 1. input patch -> shared memory as 16-bit pixels of 4 channel slots (8 bytes; zero outside the image = 'same'
      //    padding).  All of a thread's loads (up to 5 pixels x C channels x NSRC images) are issued before the first
      //    store, so one DRAM round trip covers the whole patch.
      constexpr int NPX = (PATCH_H * PATCH_W + 127) / 128;
      float4 pv[NSRC][NPX];
#pragma unroll
      for (int s = 0; s < NSRC; ++s) {
        const float* img = p.src[s] + (size_t)b * p.H * p.W * p.C;
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
          const int i = pl + j * 128;
          const int r = i / PATCH_W, c = i - r * PATCH_W;
          const int ih = ih0 + r, iw = iw0 + c;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < PATCH_H * PATCH_W && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
            const float* px = img + ((size_t)ih * p.W + iw) * p.C;
            v.x = __ldg(px);
            if (p.C > 1) v.y = __ldg(px + 1);
            if (p.C > 2) v.z = __ldg(px + 2);
            if (p.C > 3) v.w = __ldg(px + 3);
          }
          pv[s][j] = v;
        }
      }
#pragma unroll
      for (int s = 0; s < NSRC; ++s)
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
          const int i = pl + j * 128;
          if (i < PATCH_H * PATCH_W) {
            uint2 o;
            if (p.ab_bf16) { o.x = pack2<bf16>(pv[s][j].x, pv[s][j].y); o.y = pack2<bf16>(pv[s][j].z, pv[s][j].w); }
            else { o.x = pack2<f16>(pv[s][j].x, pv[s][j].y); o.y = pack2<f16>(pv[s][j].z, pv[s][j].w); }
            *reinterpret_cast<uint2*>(my_patch + s * PATCH_BYTES + i * 8) = o;
          }
        }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      // 2. this thread's row: chunk j = taps (kh, kw0), (kh, kw0+1) = two ADJACENT patch pixels = 16 contiguous, 16-byte
      //    aligned bytes of the patch (kw0 is even), copied to chunk (j ^ (row & 7)) of the row (SWIZZLE_128B)
#pragma unroll
      for (int s = 0; s < NSRC; ++s) {
        const uint8_t* ps = my_patch + s * PATCH_BYTES;
        uint8_t* row = my_a + s * A_BYTES + pl * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int kh = j >> 1, kw0 = (j & 1) * 2;
          const uint4 v = *reinterpret_cast<const uint4*>(ps + (2 * olh + kh) * PATCH_PITCH + (2 * olw + kw0) * 8);
          *reinterpret_cast<uint4*>(row + ((j ^ (pl & 7)) << 4)) = v;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (pl == 0) mbar_arrive(&bars[g]);
      // 3. accumulator -> LeakyReLU -> consumer view.  A warp-wide store in which every lane writes into its own pixel
      //    row touches 32 lines = 32 LSU wavefronts (measured: that alone bounded the first version at 182 us per
      //    launch), so the tile is transposed through the group's A buffer (free once the MMA has committed;
      //    XOR-swizzled 16-byte chunks, conflict-free both ways) and written with 8 lanes per 128-byte row.
      mbar_wait(&bars[2 + g], n & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + g * 64 + ((uint32_t)(q * 32) << 16);
      uint4* ast = reinterpret_cast<uint4*>(my_a);                       // [128 rows][8 chunks] (16 KB)
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t ao[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float z0 = __uint_as_float(v[j * 8 + 2 * e]), z1 = __uint_as_float(v[j * 8 + 2 * e + 1]);
            const float a0 = z0 > 0.f ? z0 : LEAKY_SLOPE * z0, a1 = z1 > 0.f ? z1 : LEAKY_SLOPE * z1;
            ao[e] = p.out_f16 ? pack2<f16>(a0, a1) : pack2<bf16>(a0, a1);
          }
          const int ch = (c >> 3) + j;                                   // 16-byte chunk 0..7 of this pixel's row
          ast[pl * 8 + (ch ^ (pl & 7))] = make_uint4(ao[0], ao[1], ao[2], ao[3]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int qi = it * 128 + pl;                                    // linear 16-byte chunk of the tile
        const int r = qi >> 3, ch = qi & 7;                              // tile row (output pixel), chunk
        const int oh = oh0 + r / FT_W, ow = ow0 + r % FT_W;
        const size_t pix = ((size_t)b * Ho + oh) * Wo + ow;
        *reinterpret_cast<uint4*>(p.a + pix * p.a_pitch + p.a_coff + ch * 8) = ast[r * 8 + (ch ^ (r & 7))];
      }
      // the group's patch / A tile / TMEM columns are free again once all 128 threads are here
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
}

size_t first_smem_bytes(int nsrc) {
  return (size_t)2 * nsrc * 128 * 128 + (size_t)nsrc * 64 * 128 + (size_t)2 * nsrc * PATCH_BYTES + 64 + 16 + 1024;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled g_encode = nullptr;
bool g_first_on = true;

}  // namespace

void first_init() {
  static bool done = false;
  if (done) return;
  done = true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
    g_encode = (PFN_encodeTiled)fn;
  else cudaGetLastError();
  cudaFuncSetAttribute(k_conv_first_fwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)first_smem_bytes(1));
  cudaFuncSetAttribute(k_conv_first_fwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)first_smem_bytes(2));
  const char* e = getenv("GAN_B200_FIRST");          // dev A/B switch: 0 = im2col rows in HBM (round 1)
  g_first_on = !(e && e[0] == '0');
}

bool first_fwd_supported(const FirstLayerOp& op) {
  return g_first_on && g_encode != nullptr && op.nsrc >= 1 && op.nsrc <= 2 && op.C >= 1 && op.C <= 4 && op.H % (2 * FT_H) == 0 &&
         op.W % (2 * FT_W) == 0 && op.a_pitch % 8 == 0 && op.a_coff % 8 == 0 && (op.dt == DT_F16 || op.dt == DT_BF16);
}

void launch_conv_first_fwd(Launch L, const FirstLayerOp& op) {
  FirstFwdParams P; memset(&P, 0, sizeof(P));
  cuuint64_t dims[2] = {(cuuint64_t)op.nsrc * 64, 64};
  cuuint64_t strides[1] = {(cuuint64_t)op.nsrc * 64 * 2};
  cuuint32_t box[2] = {64, 64}, es[2] = {1, 1};
  CUresult r = g_encode(&P.bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)op.wpack, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw GanError(-2, "cuTensorMapEncodeTiled(first layer) failed: " + std::to_string((int)r));
  P.src[0] = op.src[0]; P.src[1] = op.src[1]; P.nsrc = op.nsrc; P.C = op.C; P.B = op.B; P.H = op.H; P.W = op.W;
  P.a = (bf16*)op.a; P.a_pitch = op.a_pitch; P.a_coff = op.a_coff;
  P.tiles_w = (op.W / 2) / FT_W; P.tiles_h = (op.H / 2) / FT_H; P.num_tiles = P.tiles_w * P.tiles_h * op.B;
  P.ab_bf16 = op.dt == DT_F16 ? 0 : 1; P.out_f16 = op.dt == DT_F16 ? 1 : 0;
  const int per_sm = 2;                                          // registers (106-118 x 288 threads) allow two CTAs; shared memory 52 / 102 KB
  int grid = (P.num_tiles + 1) / 2;
  if (grid > 148 * per_sm) grid = 148 * per_sm;
  if (op.nsrc == 1) k_conv_first_fwd<1><<<grid, FIRST_THREADS, first_smem_bytes(1), L.s>>>(P);
  else k_conv_first_fwd<2><<<grid, FIRST_THREADS, first_smem_bytes(2), L.s>>>(P);
  KLAUNCH(L);
}
