// conv_first.cu — the image-channel first layers (G.down1 base_gan.py:180, D.down1 base_gan.py:141: Conv2D 4x4 s2
// 'same', Cin in {1,3} per source image (one or two sources), 64 filters, no normalisation, LeakyReLU 0.3) without im2col rows in HBM.
//
// The 4x4 stride-2 window of a 3-channel image is 48 values per output pixel; as a tcgen05 operand it is one
// 128-byte row k = tap*4 + channel slot (64 x 16 bit, slots >= C zero).  Round 1 materialised those rows in HBM
// (134 MB per image batch at 64 x 256^2, written by k_im2col, read by the GEMM) and then ran a separate activation pass
// over z.  Here a CTA builds the rows in SHARED memory, directly in the SWIZZLE_128B K-major layout the MMA reads:
//   * a group of 4 warps owns one 8 x 16 tile of output pixels: its 18 x 34 fp32 input patch arrives by ONE TMA load
//     per source image (out-of-bounds = zero = 'same' padding) through a ring of 2-3 stages per group, so the next
//     tiles are in flight while this one is processed; every thread assembles the row of its output pixel (16 taps x
//     4 slots -> eight swizzled 16-byte stores), fences the async proxy and signals the MMA warp;
//   * the last warp issues 4 (x sources) tcgen05.mma M=128 x N=64 x K=16 against the stationary weight tile (one TMA load per
//     CTA) into the group's TMEM columns and commits to the group's barrier;
//   * the same 4 warps drain TMEM: a = LeakyReLU(z) goes through shared memory into the consumer view (skip-concat
//     buffer) with 8 lanes per 128-byte row.  z itself is not stored: LeakyReLU keeps the sign, so the backward pass
//     takes the activation derivative from a (launch_norm_bwd with a strided z view).
// Two or three groups per CTA work on different tiles, so the gather / store phases of one overlap the others'.  HBM
// traffic per image batch: the fp32 image(s) once + a (ncu: DRAM reads = 1.00 x the image bytes) — the roofline of this
// layer is HBM, and the kernel moves nothing else.  k_conv_first_wgrad (below) is the weight gradient on the same front end.
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
constexpr int PATCH_H = 2 * FT_H + 2;              // 18 input rows
constexpr int PATCH_W = 2 * FT_W + 2;              // 34 input columns, first = 2*ow0 - 1
constexpr int PATCH_LEFT = 3;                      // the staged row starts 3 pixels further left, at column 2*ow0 - 4

// One tile's input: a (37*C rounded up to 4 floats) x 18 box of the fp32 image, fetched by ONE TMA load per source
// (3-D map (W*C, H, B); the box may start at column -4 / row -1 and overhang the right / bottom edge: out-of-bounds
// elements arrive as zeros = the 'same' padding, rows of the neighbouring image are never touched).  The box starts
// at column 2*ow0 - 4 rather than 2*ow0 - 1 because the innermost TMA coordinate has to land on a 16-byte boundary
// ((2*ow0 - 4) * C floats is a multiple of 4 for every C; probed on the device: a box starting at -3 floats raises an
// illegal-instruction fault).
__host__ __device__ constexpr int first_row_floats(int C) { return (((PATCH_W + PATCH_LEFT) * C + 3) / 4) * 4; }    // 112 (C=3), 40 (C=1)
__host__ __device__ constexpr int first_box_bytes(int C) { return PATCH_H * first_row_floats(C) * 4; }
__host__ __device__ constexpr int first_stage_bytes(int C) { return ((first_box_bytes(C) + 127) / 128) * 128; }
__host__ __device__ constexpr int first_threads(int NG) { return NG * 128 + 32; }
__host__ __device__ constexpr int first_tmem_cols(int NG) { return NG <= 2 ? 128 : 256; }

__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

struct TileCoord { int b, oh0, ow0; };
__device__ __forceinline__ TileCoord tile_coord(int t, int tiles_w, int tiles_h) {
  TileCoord c;
  const int tw = t % tiles_w; t /= tiles_w;
  c.ow0 = tw * FT_W; c.oh0 = (t % tiles_h) * FT_H; c.b = t / tiles_h;
  return c;
}

// The 128-byte operand row of output pixel `pl` from the staged fp32 box: chunk j = taps (kh, kw0), (kh, kw0+1) = two
// adjacent input pixels = 2*C consecutive floats -> 2 x 4 channel slots of 16 bit -> 16-byte chunk (j ^ (row & 7)) of the
// row (SWIZZLE_128B, the layout the MMA descriptor reads).  The floats start at an odd word (column offset 3), so C = 3
// reads them as 32 + 64 + 64 + 32 bits.
template <int C>
__device__ __forceinline__ void build_row(const uint8_t* stage_src, uint8_t* row, int pl, int ab_bf16) {
  constexpr int ROWF = first_row_floats(C);
  const int olh = pl / FT_W, olw = pl % FT_W;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int kh = j >> 1, kw0 = (j & 1) * 2;
    const float* f = reinterpret_cast<const float*>(stage_src) + (2 * olh + kh) * ROWF + (2 * olw + kw0 + PATCH_LEFT) * C;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (C == 1) { v[0] = f[0]; v[4] = f[1]; }
    else {
      const float a = f[0]; const float2 b = *reinterpret_cast<const float2*>(f + 1), c = *reinterpret_cast<const float2*>(f + 3);
      const float d = f[5];
      v[0] = a; v[1] = b.x; v[2] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d;
    }
    uint4 o;
    if (ab_bf16) { o.x = pack2<bf16>(v[0], v[1]); o.y = pack2<bf16>(v[2], v[3]); o.z = pack2<bf16>(v[4], v[5]); o.w = pack2<bf16>(v[6], v[7]); }
    else { o.x = pack2<f16>(v[0], v[1]); o.y = pack2<f16>(v[2], v[3]); o.z = pack2<f16>(v[4], v[5]); o.w = pack2<f16>(v[6], v[7]); }
    *reinterpret_cast<uint4*>(row + ((j ^ (pl & 7)) << 4)) = o;
  }
}

struct alignas(64) FirstFwdParams {
  CUtensorMap bmap;                  // packed weights [64][nsrc*64], K-major
  CUtensorMap imap[2];               // fp32 images as (W*C, H, B)
  int nsrc, C, B, H, W;              // H, W: input size; output grid H/2 x W/2
  bf16* a; int a_pitch, a_coff;      // LeakyReLU(z) into the consumer view (16-bit)
  int tiles_w, tiles_h, num_tiles;
  int ab_bf16, out_f16;
};

// NSRC source images of C channels each; NG gather/epilogue groups of 4 warps (+ one MMA warp); NS input stages per group
template <int NSRC, int C, int NG, int NS>
__global__ void __launch_bounds__(first_threads(NG)) k_conv_first_fwd(const __grid_constant__ FirstFwdParams p) {
  constexpr uint32_t A_BYTES = 128 * 128;                         // one source: 128 rows x 64 x 16 bit
  constexpr uint32_t GROUP_A = NSRC * A_BYTES;
  constexpr int STAGE_SRC = first_stage_bytes(C), STAGE = NSRC * STAGE_SRC;
  constexpr uint32_t BOX_BYTES = NSRC * first_box_bytes(C);
  constexpr int MMA_WARP = NG * 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* atile = smem;                                          // [NG groups][NSRC][128][128 B]
  uint8_t* wsm = smem + NG * GROUP_A;                             // [NSRC][64 rows][128 B]
  uint8_t* stage = wsm + NSRC * 64 * 128;                         // [NG groups][NS stages][NSRC][18 rows][ROWF] fp32
  uint64_t* bars = (uint64_t*)(stage + NG * NS * STAGE);          // full[NG], done[NG], wbar, stfull[NG][NS]
  uint64_t* stfull = bars + 2 * NG + 1;
  uint32_t* tmem_slot = (uint32_t*)(stfull + NG * NS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * NG + 1 + NG * NS; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&p.bmap);
    for (int s = 0; s < NSRC; ++s) prefetch_tmap(&p.imap[s]);
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)first_tmem_cols(NG)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int Ho = p.H / 2, Wo = p.W / 2;

  if (warp == MMA_WARP) {
    // ---- MMA issuer: stationary weights, then one k-block (NSRC x 4 MMAs) per tile ----
    if (lane == 0) {
      mbar_expect_tx(&bars[2 * NG], NSRC * 64 * 128);
      for (int s = 0; s < NSRC; ++s) tma_load_2d(wsm + s * 64 * 128, &p.bmap, &bars[2 * NG], s * 64, 0);
    }
    mbar_wait(&bars[2 * NG], 0);
    const uint32_t idesc = idesc_f16(128, 64, p.ab_bf16);
    uint32_t n = 0;
    for (int t = blockIdx.x * NG; t < p.num_tiles; t += gridDim.x * NG, ++n) {
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        if (t + g >= p.num_tiles) break;
        mbar_wait(&bars[g], n & 1);                                 // the group has built its A tile
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
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[NG + g])) : "memory");
        }
        __syncwarp();
      }
    }
  } else {
    // ---- gather / epilogue groups ----
    const int g = warp >> 2;                       // group of this warp
    const int q = warp & 3;                        // TMEM lane quadrant of this warp
    const int pl = (threadIdx.x & 127);            // output pixel inside the tile == accumulator row
    uint8_t* my_stage = stage + g * NS * STAGE;
    uint8_t* my_a = atile + g * GROUP_A;
    uint64_t* my_full = stfull + g * NS;
    const int bar_id = 1 + g;                      // named barrier of the group (128 threads)
    const int tstride = gridDim.x * NG;

    auto fetch = [&](int t, int st) {              // one thread: the tile's input boxes -> stage st
      const TileCoord tc = tile_coord(t, p.tiles_w, p.tiles_h);
      mbar_expect_tx(&my_full[st], BOX_BYTES);
#pragma unroll
      for (int s = 0; s < NSRC; ++s)
        tma_load_3d(my_stage + st * STAGE + s * STAGE_SRC, &p.imap[s], &my_full[st], (2 * tc.ow0 - 1 - PATCH_LEFT) * C, 2 * tc.oh0 - 1, tc.b);
    };

    int t = blockIdx.x * NG + g;
    if (pl == 0)
      for (int st = 0; st < NS; ++st)
        if (t + st * tstride < p.num_tiles) fetch(t + st * tstride, st);
    uint32_t n = 0;
    for (; t < p.num_tiles; t += tstride, ++n) {
      const TileCoord tc = tile_coord(t, p.tiles_w, p.tiles_h);
      const int st = n % NS;
      // 1. this tile's rows have landed (issued NS tiles ago)
      mbar_wait(&my_full[st], (n / NS) & 1);
      // 2. operand rows
#pragma unroll
      for (int s = 0; s < NSRC; ++s)
        build_row<C>(my_stage + st * STAGE + s * STAGE_SRC, my_a + s * A_BYTES + pl * 128, pl, p.ab_bf16);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (pl == 0) {
        mbar_arrive(&bars[g]);
        if (t + NS * tstride < p.num_tiles) fetch(t + NS * tstride, st);     // every thread of the group has read the stage
      }
      // 3. accumulator -> LeakyReLU -> consumer view.  A warp-wide store in which every lane writes into its own pixel
      //    row touches 32 lines = 32 LSU wavefronts (measured: that alone bounded the first version at 182 us per
      //    launch), so the tile is transposed through the group's A buffer (free once the MMA has committed;
      //    XOR-swizzled 16-byte chunks, conflict-free both ways) and written with 8 lanes per 128-byte row.
      mbar_wait(&bars[NG + g], n & 1);
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
        const int oh = tc.oh0 + r / FT_W, ow = tc.ow0 + r % FT_W;
        const size_t pix = ((size_t)tc.b * Ho + oh) * Wo + ow;
        *reinterpret_cast<uint4*>(p.a + pix * p.a_pitch + p.a_coff + ch * 8) = ast[r * 8 + (ch ^ (r & 7))];
      }
      // the group's A tile / TMEM columns are free again once all 128 threads are here
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)first_tmem_cols(NG)) : "memory");
}

size_t first_smem_bytes(int nsrc, int C, int ng, int ns) {
  return (size_t)ng * nsrc * 128 * 128 + (size_t)nsrc * 64 * 128 + (size_t)ng * ns * nsrc * first_stage_bytes(C) + 8 * (2 * ng + 1 + ng * ns) + 16 + 1024;
}

// =============================================================================================
// weight gradient of the first layers.  D[128 = (source, tap*4+slot)][64 filters] += A^T . dz over the tile's 128 pixels:
// A = the same shared-memory rows as the forward kernel, read as the MN-major operand (M = the 64 k-values of a pixel
// row are contiguous, GEMM-K = pixels; both sources side by side = M 128; with one source the upper 64 rows read a
// zeroed buffer), B = the dz tile [128 pixels][64] brought by TMA (double-buffered, one tile ahead), also MN-major.
// ONE accumulator per CTA collects every tile the CTA processes (fixed order); the 128 x 64 fp32 partial goes to a
// slab, k_wgrad_reduce sums the CTAs.
// =============================================================================================
struct alignas(64) FirstWgradParams {
  CUtensorMap dmap;                  // dz (B, Ho, Wo, 64): boxes 64 ch x 16 x 8 x 1, SWIZZLE_128B
  CUtensorMap imap[2];               // fp32 images as (W*C, H, B)
  int nsrc, C, B, H, W;
  int tiles_w, tiles_h, num_tiles;
  int ab_bf16;
  float* slab;                       // [gridDim.x][128][64]
};

// MN-major SWIZZLE_128B descriptor (conv_umma.cu): 128 bytes of MN per k-row, 8 k-rows per atom (SBO = 1024),
// LBO = byte distance between consecutive 64-element MN blocks
__device__ __forceinline__ uint64_t desc_mn128(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}

template <int NSRC, int C, int NG, int NS>
__global__ void __launch_bounds__(first_threads(NG)) k_conv_first_wgrad(const __grid_constant__ FirstWgradParams p) {
  constexpr uint32_t A_BYTES = 128 * 128;
  constexpr uint32_t GROUP_A = NSRC * A_BYTES;
  constexpr int STAGE_SRC = first_stage_bytes(C), STAGE = NSRC * STAGE_SRC;
  constexpr uint32_t BOX_BYTES = NSRC * first_box_bytes(C);
  constexpr int MMA_WARP = NG * 4;
  constexpr int ZBUF = NSRC == 1 ? 1 : 0;                         // one source: 16 KB of zeros stand in for the second
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* atile = smem;                                          // [NG][NSRC][128 pixels][128 B]
  uint8_t* zero = atile + NG * GROUP_A;                           // [ZBUF][16 KB]
  uint8_t* btile = zero + ZBUF * A_BYTES;                         // [NG][2][128 pixels][128 B] dz
  uint8_t* stage = btile + NG * 2 * A_BYTES;                      // [NG][NS][NSRC][18 rows][ROWF] fp32
  uint64_t* bars = (uint64_t*)(stage + NG * NS * STAGE);          // full[NG], done[NG], accfull, dzfull[NG][2], stfull[NG][NS]
  uint64_t* dzfull = bars + 2 * NG + 1;
  uint64_t* stfull = dzfull + 2 * NG;
  uint32_t* tmem_slot = (uint32_t*)(stfull + NG * NS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * NG + 1 + 2 * NG + NG * NS; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&p.dmap);
    for (int s = 0; s < NSRC; ++s) prefetch_tmap(&p.imap[s]);
  }
  if (ZBUF) {
    for (int i = threadIdx.x; i < (int)(A_BYTES / 16); i += first_threads(NG)) reinterpret_cast<uint4*>(zero)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == MMA_WARP) {
    // M = 128 (k-values of both sources), N = 64 (filters), both operands MN-major
    const uint32_t idesc = idesc_f16(128, 64, p.ab_bf16) | (1u << 15) | (1u << 16);
    uint32_t n = 0;
    bool first = true;
    for (int t = blockIdx.x * NG; t < p.num_tiles; t += gridDim.x * NG, ++n) {
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        if (t + g >= p.num_tiles) break;
        mbar_wait(&bars[g], n & 1);                                 // rows built
        mbar_wait(&dzfull[g * 2 + (n & 1)], (n >> 1) & 1);          // dz tile landed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t sa = smem_u32(atile + g * GROUP_A), sb = smem_u32(btile + (g * 2 + (n & 1)) * A_BYTES);
          const uint32_t lbo = NSRC == 2 ? A_BYTES : (uint32_t)(NG - g) * A_BYTES;     // second M block: source 1, or the zeros
          const uint64_t ad = desc_mn128(sa, lbo), bd = desc_mn128(sb, A_BYTES);
#pragma unroll
          for (int k = 0; k < 8; ++k) {                             // 16 pixels (two 8-row atoms) per MMA
            const uint32_t acc = !(first && k == 0);
            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                         ::"r"(tmem_base), "l"(ad + (uint64_t)(k * 128)), "l"(bd + (uint64_t)(k * 128)), "r"(idesc), "r"(acc) : "memory");
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[NG + g])) : "memory");
        }
        __syncwarp();
        first = false;
      }
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[2 * NG])) : "memory");
    __syncwarp();
  } else {
    const int g = warp >> 2, q = warp & 3;
    const int pl = (threadIdx.x & 127);
    uint8_t* my_stage = stage + g * NS * STAGE;
    uint8_t* my_a = atile + g * GROUP_A;
    uint8_t* my_b = btile + g * 2 * A_BYTES;
    uint64_t* my_full = stfull + g * NS;
    const int bar_id = 1 + g;
    const int tstride = gridDim.x * NG;

    auto fetch = [&](int t, int st) {
      const TileCoord tc = tile_coord(t, p.tiles_w, p.tiles_h);
      mbar_expect_tx(&my_full[st], BOX_BYTES);
#pragma unroll
      for (int s = 0; s < NSRC; ++s)
        tma_load_3d(my_stage + st * STAGE + s * STAGE_SRC, &p.imap[s], &my_full[st], (2 * tc.ow0 - 1 - PATCH_LEFT) * C, 2 * tc.oh0 - 1, tc.b);
    };
    auto fetch_dz = [&](int t, int buf) {
      const TileCoord tc = tile_coord(t, p.tiles_w, p.tiles_h);
      mbar_expect_tx(&dzfull[g * 2 + buf], A_BYTES);
      tma_load_4d(my_b + buf * A_BYTES, &p.dmap, &dzfull[g * 2 + buf], 0, tc.ow0, tc.oh0, tc.b);
    };

    int t = blockIdx.x * NG + g;
    if (pl == 0) {
      for (int st = 0; st < NS; ++st)
        if (t + st * tstride < p.num_tiles) fetch(t + st * tstride, st);
      if (t < p.num_tiles) fetch_dz(t, 0);
    }
    uint32_t n = 0;
    for (; t < p.num_tiles; t += tstride, ++n) {
      const int st = n % NS;
      if (n > 0) mbar_wait(&bars[NG + g], (n - 1) & 1);             // the previous tile's MMAs have read the rows and its dz
      if (pl == 0 && t + tstride < p.num_tiles) fetch_dz(t + tstride, (n + 1) & 1);      // dz one tile ahead
      mbar_wait(&my_full[st], (n / NS) & 1);
#pragma unroll
      for (int s = 0; s < NSRC; ++s)
        build_row<C>(my_stage + st * STAGE + s * STAGE_SRC, my_a + s * A_BYTES + pl * 128, pl, p.ab_bf16);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (pl == 0) {
        mbar_arrive(&bars[g]);
        if (t + NS * tstride < p.num_tiles) fetch(t + NS * tstride, st);
      }
    }
    if (g == 0) {
      // the CTA's partial tile -> slab (once per CTA: 32 KB)
      mbar_wait(&bars[2 * NG], 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int r = q * 32 + lane;
      float* sl = p.slab + ((size_t)blockIdx.x * 128 + r) * 64;
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, v);
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(sl + c + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
}

size_t first_wgrad_smem_bytes(int nsrc, int C, int ng, int ns) {
  return (size_t)ng * nsrc * 128 * 128 + (nsrc == 1 ? 128 * 128 : 0) + (size_t)ng * 2 * 128 * 128 + (size_t)ng * ns * nsrc * first_stage_bytes(C) +
         8 * (2 * ng + 1 + 2 * ng + ng * ns) + 16 + 1024;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled g_encode = nullptr;
bool g_first_on = true, g_first_wgrad_on = true;

}  // namespace

// fp32 NHWC image (B, H, W, C) as the 3-D tensor (W*C, H, B); box = one tile's 18 input rows
static void make_image_map(CUtensorMap* m, const float* src, int C, int B, int H, int W) {
  cuuint64_t dims[3] = {(cuuint64_t)W * C, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[3] = {(cuuint32_t)first_row_floats(C), PATCH_H, 1}, es[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw GanError(-2, "cuTensorMapEncodeTiled(first-layer image) failed: " + std::to_string((int)r));
}

// kernel configurations: (sources, channels) -> groups per CTA, input stages per group, CTAs per SM
//   forward : one source 2 groups x 3 stages (86 KB, two CTAs per SM); two sources 3 groups x 2 stages (204 KB)
//   wgrad   : one source 3 groups x 2 stages (205 KB); two sources 2 groups x 3 stages (220 KB)
#define FIRST_FWD_CASES(X) X(1, 1, 2, 3) X(1, 3, 2, 3) X(2, 1, 3, 2) X(2, 3, 3, 2)
#define FIRST_WG_CASES(X) X(1, 1, 3, 2) X(1, 3, 3, 2) X(2, 1, 2, 3) X(2, 3, 2, 3)

void first_init() {
  static bool done = false;
  if (done) return;
  done = true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
    g_encode = (PFN_encodeTiled)fn;
  else cudaGetLastError();
#define X(NSRC, C, NG, NS) cudaFuncSetAttribute(k_conv_first_fwd<NSRC, C, NG, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)first_smem_bytes(NSRC, C, NG, NS));
  FIRST_FWD_CASES(X)
#undef X
#define X(NSRC, C, NG, NS) cudaFuncSetAttribute(k_conv_first_wgrad<NSRC, C, NG, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)first_wgrad_smem_bytes(NSRC, C, NG, NS));
  FIRST_WG_CASES(X)
#undef X
  const char* e = getenv("GAN_B200_FIRST");          // dev A/B switch: 0 = im2col rows in HBM (round 1)
  g_first_on = !(e && e[0] == '0');
  g_first_wgrad_on = !(e && e[0] == '1');            // 1 = forward kernel only (weight gradient on im2col rows)
}

static bool first_geom_ok(const float* const* src, int nsrc, int C, int H, int W) {
  for (int s = 0; s < nsrc && s < 2; ++s)
    if (((uintptr_t)src[s] & 15) != 0) return false;              // TMA: 16-byte aligned base (strides are: W % 32 == 0)
  return g_encode != nullptr && nsrc >= 1 && nsrc <= 2 && (C == 1 || C == 3) && H % (2 * FT_H) == 0 && W % (2 * FT_W) == 0;
}

bool first_fwd_supported(const FirstLayerOp& op) {
  return g_first_on && first_geom_ok(op.src, op.nsrc, op.C, op.H, op.W) && op.a_pitch % 8 == 0 && op.a_coff % 8 == 0 &&
         (op.dt == DT_F16 || op.dt == DT_BF16);
}
bool first_wgrad_enabled() { return g_first_on && g_first_wgrad_on; }

void launch_conv_first_fwd(Launch L, const FirstLayerOp& op) {
  FirstFwdParams P; memset(&P, 0, sizeof(P));
  cuuint64_t dims[2] = {(cuuint64_t)op.nsrc * 64, 64};
  cuuint64_t strides[1] = {(cuuint64_t)op.nsrc * 64 * 2};
  cuuint32_t box[2] = {64, 64}, es[2] = {1, 1};
  CUresult r = g_encode(&P.bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)op.wpack, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw GanError(-2, "cuTensorMapEncodeTiled(first layer) failed: " + std::to_string((int)r));
  for (int s = 0; s < op.nsrc; ++s) make_image_map(&P.imap[s], op.src[s], op.C, op.B, op.H, op.W);
  P.nsrc = op.nsrc; P.C = op.C; P.B = op.B; P.H = op.H; P.W = op.W;
  P.a = (bf16*)op.a; P.a_pitch = op.a_pitch; P.a_coff = op.a_coff;
  P.tiles_w = (op.W / 2) / FT_W; P.tiles_h = (op.H / 2) / FT_H; P.num_tiles = P.tiles_w * P.tiles_h * op.B;
  P.ab_bf16 = op.dt == DT_F16 ? 0 : 1; P.out_f16 = op.dt == DT_F16 ? 1 : 0;
  const int per_sm = op.nsrc == 1 ? 2 : 1;
  bool launched = false;
#define X(NSRC, C_, NG, NS) \
  if (!launched && op.nsrc == NSRC && op.C == C_) { \
    int grid = (P.num_tiles + NG - 1) / NG; \
    if (grid > 148 * per_sm) grid = 148 * per_sm; \
    k_conv_first_fwd<NSRC, C_, NG, NS><<<grid, first_threads(NG), first_smem_bytes(NSRC, C_, NG, NS), L.s>>>(P); \
    launched = true; \
  }
  FIRST_FWD_CASES(X)
#undef X
  if (!launched) throw GanError(-2, "first-layer kernel: unsupported (sources, channels)");
  KLAUNCH(L);
}

bool first_wgrad_supported(const FirstWgradOp& op) {
  return first_wgrad_enabled() && first_geom_ok(op.src, op.nsrc, op.C, op.H, op.W) && op.dz_pitch % 8 == 0 && op.dz_coff % 8 == 0 &&
         (op.dt == DT_F16 || op.dt == DT_BF16) && op.ws != nullptr && op.ws_bytes >= (size_t)148 * 2 * 128 * 64 * 4;
}

void launch_conv_first_wgrad(Launch L, const FirstWgradOp& op) {
  FirstWgradParams P; memset(&P, 0, sizeof(P));
  const int Ho = op.H / 2, Wo = op.W / 2;
  cuuint64_t dims[4] = {64, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)op.B};
  cuuint64_t strides[3] = {(cuuint64_t)op.dz_pitch * 2, (cuuint64_t)Wo * op.dz_pitch * 2, (cuuint64_t)Ho * Wo * op.dz_pitch * 2};
  cuuint32_t box[4] = {64, FT_W, FT_H, 1}, es[4] = {1, 1, 1, 1};
  void* base = (void*)((const uint8_t*)op.dz + (size_t)op.dz_coff * 2);
  CUresult r = g_encode(&P.dmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw GanError(-2, "cuTensorMapEncodeTiled(first-layer dz) failed: " + std::to_string((int)r));
  for (int s = 0; s < op.nsrc; ++s) make_image_map(&P.imap[s], op.src[s], op.C, op.B, op.H, op.W);
  P.nsrc = op.nsrc; P.C = op.C; P.B = op.B; P.H = op.H; P.W = op.W;
  P.tiles_w = Wo / FT_W; P.tiles_h = Ho / FT_H; P.num_tiles = P.tiles_w * P.tiles_h * op.B;
  P.ab_bf16 = op.dt == DT_F16 ? 0 : 1;
  P.slab = op.ws;
  int grid = 0;
#define X(NSRC, C_, NG, NS) \
  if (grid == 0 && op.nsrc == NSRC && op.C == C_) { \
    grid = (P.num_tiles + NG - 1) / NG; \
    if (grid > 148) grid = 148; \
    k_conv_first_wgrad<NSRC, C_, NG, NS><<<grid, first_threads(NG), first_wgrad_smem_bytes(NSRC, C_, NG, NS), L.s>>>(P); \
  }
  FIRST_WG_CASES(X)
#undef X
  if (grid == 0) throw GanError(-2, "first-layer weight gradient: unsupported (sources, channels)");
  KLAUNCH(L);
  // CTA partials -> master layout (kh, kw, source*C + c, co), summed in CTA order
  WgradReduceParams R; memset(&R, 0, sizeof(R));
  R.slab = op.ws; R.dW = op.dW; R.splits = grid; R.bn = 64; R.mblocks = 1; R.ntiles = 1; R.ncls = 1; R.accumulate = op.accumulate;
  R.Kc = 64; R.Kr = 64; R.Nr = 64; R.im2col_c = op.C; R.n_slot4_c = 0; R.s_tap = op.s_tap; R.s_k = op.s_k; R.s_n = op.s_n;
  R.ntaps[0] = op.nsrc; R.widx[0][0] = 0; R.widx[0][1] = 1;
  launch_wgrad_reduce(L, R, 1);
}
