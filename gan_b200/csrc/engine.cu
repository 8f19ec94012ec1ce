// engine.cu — host orchestration of the GAN train step and the C-ABI (include/gan_b200.h).
//
// Restates, as kernel launch sequences on one CUDA stream:
//   GAN.Generator / GAN.Discriminator forward ........ base_gan.py:124-225
//   Pix2Pix.train_step ................................ pix2pix.py:190-218
//   CycleGAN.train_step ............................... cycle_gan.py:206-276
// Autodiff (tf.GradientTape) is replaced by hand-derived backward sweeps over the saved raw
// convolution outputs ("z") and normalisation statistics of each forward call (a Slot).
#include <cstring>
#include <cstdio>
#include <cmath>
#include <mutex>
#include <set>
#include <memory>
#include <algorithm>
#include <functional>
#include "engine.h"

uint64_t g_alloc_epoch = 0;
#include "../../include/gan_b200.h"

static thread_local std::string g_last_error;

#define API_BEGIN try {
#define API_END                                                              \
  return GAN_OK;                                                             \
  } catch (const GanError& e) { g_last_error = e.what(); return e.code; }    \
  catch (const std::exception& e) { g_last_error = e.what(); return GAN_ERR_INVALID; }

static const int DOWN_F[8] = {64, 128, 256, 512, 512, 512, 512, 512};   // base_gan.py:179-188
static const int UP_F[7] = {512, 512, 512, 512, 256, 128, 64};          // base_gan.py:190-198
#define BN_EPS 1e-3f
#define IN_EPS 1e-5f
#define BN_MOMENTUM 0.99f

enum { R_FWD = 0, R_DGRAD = 1, R_WGRAD = 2 };

// ---------------------------------------------------------------------------------------------
// Tap geometry (oracle/direct.py CONVT_TAPS)
// ---------------------------------------------------------------------------------------------
static ClassGeom geom_conv16(int sign) {
  ClassGeom g; memset(&g, 0, sizeof(g));
  g.ntaps = 16;
  for (int kh = 0; kh < 4; ++kh)
    for (int kw = 0; kw < 4; ++kw) {
      int t = kh * 4 + kw;
      g.dh[t] = (int8_t)(sign * (kh - 1)); g.dw[t] = (int8_t)(sign * (kw - 1)); g.widx[t] = (int8_t)t;
    }
  return g;
}
static void geom_convT4(ClassGeom* cls) {
  static const int KH_T[2][2] = {{1, 3}, {0, 2}};
  static const int DH_T[2][2] = {{0, -1}, {1, 0}};
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      ClassGeom g; memset(&g, 0, sizeof(g));
      g.oa = a; g.ob = b; g.ntaps = 4;
      for (int th = 0; th < 2; ++th)
        for (int tw = 0; tw < 2; ++tw) {
          int t = th * 2 + tw;
          g.dh[t] = (int8_t)DH_T[a][th]; g.dw[t] = (int8_t)DH_T[b][tw];
          g.widx[t] = (int8_t)(KH_T[a][th] * 4 + KH_T[b][tw]);
        }
      cls[a * 2 + b] = g;
    }
}
static int fill_geometry(int kind, int role, ClassGeom* cls) {
  bool conv_form = (kind != K_CONVT_S2 && role != R_DGRAD) || (kind == K_CONVT_S2 && role == R_DGRAD) ||
                   (kind == K_CONV_S1P);
  if (conv_form) { cls[0] = geom_conv16((kind == K_CONV_S1P && role == R_DGRAD) ? -1 : 1); return 1; }
  geom_convT4(cls);
  return 4;
}

// Small channel counts are stored zero-padded to 16 in bf16 mode so that every layer fits a
// tcgen05 tile (first layers Cin in {1,2,3,6}: base_gan.py:141,180; heads Cout in {1,3}: :159,:201).
static int pad_c(const gan_ctx* ctx, int c) { return (ctx->dt == DT_BF16 && c < 16) ? 16 : c; }

// Master-weight strides for (tap, K-channel, N-channel) of a role; Kc/Nc are the stored (padded)
// GEMM channel counts, Kr/Nr the real ones.
static void weight_strides(const Layer& ly, int role, int& Kc, int& Nc, int& Kr, int& Nr, int64_t& s_tap, int64_t& s_k,
                           int64_t& s_n) {
  s_tap = (int64_t)ly.Cin * ly.Cout;
  bool transposed_master = (ly.kind == K_CONVT_S2);   // (kh,kw,out,in) instead of (kh,kw,in,out)
  int64_t s_in = transposed_master ? 1 : ly.Cout, s_out = transposed_master ? ly.Cin : 1;
  if (role == R_DGRAD) { Kc = ly.Cout_p; Nc = ly.Cin_p; Kr = ly.Cout; Nr = ly.Cin; s_k = s_out; s_n = s_in; }
  else { Kc = ly.Cin_p; Nc = ly.Cout_p; Kr = ly.Cin; Nr = ly.Cout; s_k = s_in; s_n = s_out; }
}

// x: layer-input-side view (N,Hin,Win,Cin); y: layer-output-side view (N,Hout,Wout,Cout).
static ConvOp make_op(const gan_ctx* ctx, const Layer& ly, int role, View x, View y, const void* wpack) {
  ConvOp op; memset(&op, 0, sizeof(op));
  // forward: activations in, activations out; dgrad: gradients in, gradients out; wgrad: activations x gradients
  op.dt_in = role == R_DGRAD ? ctx->dtG : ctx->dtA;
  op.dt_out = role == R_FWD ? ctx->dtA : ctx->dtG;
  op.ncls = fill_geometry(ly.kind, role, op.cls);
  weight_strides(ly, role, op.Kc, op.Nc, op.Kr, op.Nr, op.s_tap, op.s_k, op.s_n);
  View src = (role == R_DGRAD) ? y : x, dst = (role == R_DGRAD) ? x : y;
  op.in = src.p; op.in_pitch = src.pitch; op.in_coff = src.coff; op.Hin = src.H; op.Win = src.W;
  op.out = dst.p; op.out_pitch = dst.pitch; op.out_coff = dst.coff; op.Hout = dst.H; op.Wout = dst.W;
  op.N = x.N;
  if (role != R_DGRAD) {
    if (ly.kind == K_CONVT_S2) { op.Hm = x.H; op.Wm = x.W; op.si = 1; op.so = 2; }
    else { op.Hm = y.H; op.Wm = y.W; op.si = (ly.kind == K_CONV_S2) ? 2 : 1; op.so = 1; }
  } else {
    if (ly.kind == K_CONV_S2) { op.Hm = y.H; op.Wm = y.W; op.si = 1; op.so = 2; }
    else if (ly.kind == K_CONV_S1P) { op.Hm = x.H; op.Wm = x.W; op.si = 1; op.so = 1; }
    else { op.Hm = x.H; op.Wm = x.W; op.si = 2; op.so = 1; }
  }
  for (int c = 0; c < op.ncls; ++c) op.cls[c].b_off = (int64_t)c * op.Nc * op.cls[c].ntaps * op.Kc;
  op.B = wpack;
  return op;
}

// First layers as a GEMM over im2col rows (bf16 + tcgen05 engine only; the FFMA cross-check and the
// fp32 mode keep the generic tap path).
static bool im2col_on(const gan_ctx* ctx, const Layer& ly) {
  return ly.first && ctx->dt == DT_BF16 && ctx->engine != GAN_ENGINE_FFMA && ly.src_c <= 4 && ly.Cout_p % 64 == 0;
}
// role: R_FWD (y = z written) or R_WGRAD (y = dz read).  M-space = output grid of the layer.
static ConvOp make_op_im2col(const gan_ctx* ctx, const Layer& ly, int role, const Slot& s, View y) {
  ConvOp op; memset(&op, 0, sizeof(op));
  op.dt_in = ctx->dtA; op.dt_out = role == R_FWD ? ctx->dtA : ctx->dtG;
  op.ncls = 1;
  op.cls[0].ntaps = ly.nsrc;
  for (int t = 0; t < ly.nsrc; ++t) { op.cls[0].widx[t] = (int8_t)t; op.in_tap[t] = s.im2col[t]; }
  op.im2col_c = ly.src_c;
  op.in = s.im2col[0]; op.in_pitch = 64; op.in_coff = 0; op.Hin = y.H; op.Win = y.W;
  op.out = y.p; op.out_pitch = y.pitch; op.out_coff = y.coff; op.Hout = y.H; op.Wout = y.W;
  op.N = y.N; op.Hm = y.H; op.Wm = y.W; op.si = 1; op.so = 1;
  op.Kc = 64; op.Kr = 64; op.Nc = ly.Cout_p; op.Nr = ly.Cout;
  op.s_tap = (int64_t)ly.Cin * ly.Cout; op.s_k = ly.Cout; op.s_n = 1;      // Conv2D master (kh,kw,ci,co)
  op.B = ly.wp_im2col.p;
  op.real_k = 16 * ly.Cin;                 // every source contributes 16 taps x src_c real channels
  (void)role;
  return op;
}

// Generator head (Conv2DTranspose Cin -> C<=4, bias, tanh) as single-tap GEMMs (bf16 + tcgen05 only):
//   forward : cols[m][tap*C+co] = x[m,:] . f[tap,co,:]   then col2im + bias + tanh
//   backward: G = im2col(dz) (the 4x4 s2 unfold of the output gradient);  dW = G^T x ;  dx = G f
// so the Cin-channel activation is read once instead of once per (class, tap).
static bool cols_on(const gan_ctx* ctx, const gan_net* n, const Layer& ly) {
  return n->is_gen && ly.head && ctx->dt == DT_BF16 && ctx->engine != GAN_ENGINE_FFMA && ly.Cout <= 4 &&
         ly.Cin % 64 == 0 && ly.wp_cols.p != nullptr;
}
static bool dcols_on(const gan_ctx* ctx, const gan_net* n, const Layer& ly) {
  return !n->is_gen && ly.head && ctx->dt == DT_BF16 && ctx->engine != GAN_ENGINE_FFMA && ly.Cout == 1 &&
         ly.Cin % 64 == 0 && ly.wp_cols.p != nullptr;
}
static ConvOp make_op_1tap(View in, int Kc, View out, int Nc, int Nr, const void* B, int dt_in, int dt_out) {
  ConvOp op; memset(&op, 0, sizeof(op));
  op.dt_in = dt_in; op.dt_out = dt_out;
  op.ncls = 1; op.cls[0].ntaps = 1;
  op.in = in.p; op.in_pitch = in.pitch; op.in_coff = in.coff; op.Hin = in.H; op.Win = in.W;
  op.out = out.p; op.out_pitch = out.pitch; op.out_coff = out.coff; op.Hout = out.H; op.Wout = out.W;
  op.N = in.N; op.Hm = in.H; op.Wm = in.W; op.si = 1; op.so = 1;
  op.Kc = Kc; op.Kr = Kc; op.Nc = Nc; op.Nr = Nr; op.B = B;
  return op;
}

// im2col rows of a step input, computed once per step and shared by every first layer that reads it.
static const void* cached_im2col(gan_ctx* ctx, const float* src, int B, int H, int W, int C) {
  for (auto& e : ctx->im2col_cache)
    if (e.src == src && e.epoch == ctx->step_epoch && e.B == B && e.H == H && e.W == W && e.C == C) return e.buf.p;
  gan_ctx::Im2colEntry& e = ctx->im2col_cache[ctx->im2col_next];
  ctx->im2col_next = (ctx->im2col_next + 1) % 6;
  e.buf.ensure((size_t)B * (H / 2) * (W / 2) * 64 * 2);
  launch_im2col(ctx->L(), ctx->dtA, src, B, H, W, C, e.buf.p);
  e.src = src; e.B = B; e.H = H; e.W = W; e.C = C; e.epoch = ctx->step_epoch;
  return e.buf.p;
}

static void out_dims(int kind, int Hin, int Win, int& Ho, int& Wo) {
  if (kind == K_CONV_S2) { Ho = Hin / 2; Wo = Win / 2; }
  else if (kind == K_CONV_S1P) { Ho = Hin - 1; Wo = Win - 1; }
  else { Ho = Hin * 2; Wo = Win * 2; }
}

// Scoped CUDA-event timer around a group of launches of one kernel family (only when profiling).
struct ProfScope {
  gan_ctx* ctx; ProfEntry e; bool on;
  ProfScope(gan_ctx* c, int fam, double work) : ctx(c), on(c->profile != 0) {
    if (!on) return;
    e.fam = fam; e.work = work; e.a = c->ev_get(); e.b = c->ev_get();
    cudaEventRecord(e.a, c->cs());
  }
  ~ProfScope() { if (on) { cudaEventRecord(e.b, ctx->cs()); ctx->prof.push_back(e); } }
};
// Algorithmic FLOPs of one launch: 2*M*N*K of the UNPADDED layer — real channel counts
// Kr/Nr, and for the slot-4 operands of the image-channel ends the real 16*C columns, not the stored 64.
static double conv_flops(const ConvOp& op) {
  double f = 0;
  const double n_real = op.real_n > 0 ? op.real_n : op.Nr;
  for (int c = 0; c < op.ncls; ++c) {
    const double k_real = op.real_k > 0 ? op.real_k : (double)op.cls[c].ntaps * op.Kr;
    f += 2.0 * op.N * op.Hm * op.Wm * n_real * k_real;
  }
  return f;
}

// ---------------------------------------------------------------------------------------------
// conv dispatch: tcgen05 where the op fits (bf16 mode), FFMA otherwise
// ---------------------------------------------------------------------------------------------
// returns the number of BatchNorm-statistics partials the conv epilogue produced (0: none, run the statistics pass)
static int run_conv_fwd(gan_ctx* ctx, const ConvOp& op_in) {
  ConvOp op = op_in;
  ctx->sc().splitk_ws.ensure((size_t)32 << 20);
  op.splitk_ws = ctx->sc().splitk_ws.as<float>(); op.splitk_ws_bytes = ctx->sc().splitk_ws.bytes;
  bool can = ctx->dt == DT_BF16 && umma_fwd_supported(op);
  if (ctx->engine == GAN_ENGINE_UMMA) GAN_REQUIRE(can, "tcgen05 engine forced but op unsupported");
  const bool um = can && ctx->engine != GAN_ENGINE_FFMA;
  ProfScope ps(ctx, um ? FAM_UMMA_FWD : FAM_FFMA_FWD, conv_flops(op));
  if (um) return launch_conv_fwd_umma(ctx->L(), op);
  launch_conv_fwd_ffma(ctx->L(), op.dt_in, op);
  return 0;
}
// ly: the layer whose kernel gradient is produced.  The first contribution of a step STORES (no zeroing pass over the
// 57 M-float gradient buffer), later ones (second discriminator call, CycleGAN's repeated generator calls) add.
static void run_conv_wgrad(gan_ctx* ctx, Layer& ly, const ConvOp& op_in) {
  ConvOp op = op_in;
  op.accumulate = ly.wgrad_epoch == ctx->step_epoch ? 1 : 0;
  ly.wgrad_epoch = ctx->step_epoch;
  op.dW_elems = 16LL * ly.Cin * ly.Cout;
  ctx->sc().wgrad_ws.ensure((size_t)24 << 20);              // <= 296 partial tiles of 128 x 128 fp32 per launch
  op.wgrad_ws = ctx->sc().wgrad_ws.as<float>(); op.wgrad_ws_bytes = ctx->sc().wgrad_ws.bytes;
  bool can = ctx->dt == DT_BF16 && umma_wgrad_supported(op);
  if (ctx->engine == GAN_ENGINE_UMMA) GAN_REQUIRE(can, "tcgen05 engine forced but op unsupported");
  const bool um = can && ctx->engine != GAN_ENGINE_FFMA;
  ProfScope ps(ctx, um ? FAM_UMMA_WGRAD : FAM_FFMA_WGRAD, conv_flops(op));
  if (um) { launch_conv_wgrad_umma(ctx->L(), op); return; }
  // CUDA-core path (fp32 parity mode, odd shapes): split-M partial sums meet in fp32 atomics, so the tensor is zeroed first
  if (!op.accumulate) CUDA_CHECK(cudaMemsetAsync(op.dW, 0, (size_t)op.dW_elems * 4, ctx->cs()));
  launch_conv_wgrad_ffma(ctx->L(), op.dt_in, op.dt_out, op);
}

// ---------------------------------------------------------------------------------------------
// net construction
// ---------------------------------------------------------------------------------------------
static void add_tensor(gan_net* n, const std::string& name, std::initializer_list<int64_t> shape, int64_t& off,
                       bool trainable) {
  TensorInfo t; t.name = name; t.ndim = (int)shape.size(); t.numel = 1; t.trainable = trainable;
  int i = 0;
  for (auto s : shape) { t.shape[i++] = s; t.numel *= s; }
  for (; i < 4; ++i) t.shape[i] = 1;
  t.off = off; off += t.numel;
  n->tensors.push_back(t);
}

static void add_layer(gan_net* n, const std::string& name, int kind, int Cin, int Cout, int norm, int act, bool bias,
                      bool dropout, int tag, bool head) {
  Layer ly; ly.name = name; ly.kind = kind; ly.Cin = Cin; ly.Cout = Cout; ly.norm = norm; ly.act = act;
  ly.bias = bias; ly.dropout = dropout; ly.tag = tag; ly.head = head;
  ly.Cin_p = pad_c(n->ctx, Cin); ly.Cout_p = pad_c(n->ctx, Cout);
  if (n->layers.empty() && kind == K_CONV_S2) {       // first layer of the net: image channels (x sources)
    ly.first = true; ly.src_c = n->C; ly.nsrc = Cin / n->C;
  }
  ly.w_off = n->nparams;
  if (kind == K_CONVT_S2) add_tensor(n, name + ".kernel", {4, 4, Cout, Cin}, n->nparams, true);
  else add_tensor(n, name + ".kernel", {4, 4, Cin, Cout}, n->nparams, true);
  if (norm != NORM_NONE) {
    ly.g_off = n->nparams; add_tensor(n, name + ".gamma", {Cout}, n->nparams, true);
    ly.b_off = n->nparams; add_tensor(n, name + ".beta", {Cout}, n->nparams, true);
  }
  if (bias) { ly.bias_off = n->nparams; add_tensor(n, name + ".bias", {Cout}, n->nparams, true); }
  n->layers.push_back(std::move(ly));
}

static void finish_net(gan_net* n) {
  n->ntrain = (int)n->tensors.size();
  // BatchNorm moving statistics (non-trainable; bookkeeping only — never read, every call is training=True)
  if (n->norm == NORM_BATCH)
    for (auto& ly : n->layers)
      if (ly.norm != NORM_NONE) {
        ly.mov_off = n->nmov;
        add_tensor(n, ly.name + ".moving_mean", {ly.Cout}, n->nmov, false);
        add_tensor(n, ly.name + ".moving_variance", {ly.Cout}, n->nmov, false);
      }
  // + slack: data-parallel buckets are padded to a multiple of 4*world floats (zeros, never read back)
  n->params.ensure((size_t)(n->nparams + 1028) * 4);
  n->grads.ensure((size_t)(n->nparams + 1028) * 4);
  if (n->nmov > 0) {
    n->mov.ensure((size_t)n->nmov * 4);
    std::vector<float> init((size_t)n->nmov, 0.f);
    for (auto& t : n->tensors)
      if (!t.trainable && t.name.find("moving_variance") != std::string::npos)
        for (int64_t i = 0; i < t.numel; ++i) init[(size_t)(t.off + i)] = 1.f;
    CUDA_CHECK(cudaMemcpy(n->mov.p, init.data(), init.size() * 4, cudaMemcpyHostToDevice));
  }
  n->slots.resize(3);
  n->packed_dirty = true;
}

static void pack_weights(gan_net* n) {
  if (!n->packed_dirty) return;
  gan_ctx* ctx = n->ctx;
  if (n->pack_nent == 0) {
    // build the (static) table once: every layer x {forward, data-gradient} role
    std::vector<PackEntry> tab;
    int tiles = 0;
    for (auto& ly : n->layers) {
      for (int role = R_FWD; role <= R_DGRAD; ++role) {
        if (role == R_DGRAD && !ly.need_dgrad) continue;
        PackEntry e; memset(&e, 0, sizeof(e));
        PackOp& po = e.op;
        po.ncls = fill_geometry(ly.kind, role, po.cls);
        weight_strides(ly, role, po.Kc, po.Nc, po.Kr, po.Nr, po.s_tap, po.s_k, po.s_n);
        for (int c = 0; c < po.ncls; ++c) po.cls[c].b_off = (int64_t)c * po.Nc * po.cls[c].ntaps * po.Kc;
        DevBuf& dst = role == R_FWD ? ly.wp_fwd : ly.wp_dgrad;
        dst.ensure((size_t)16 * ly.Cin_p * ly.Cout_p * ctx->esize());
        e.master = n->params.as<float>() + ly.w_off; e.dst = dst.p;
        e.dt16 = role == R_FWD ? ctx->dtA : ctx->dtG;
        e.tiles_k = (po.Kc + 31) / 32; e.tiles_n = (po.Nc + 31) / 32;
        e.tile_begin = tiles;
        tiles += e.tiles_k * e.tiles_n * po.cls[0].ntaps * po.ncls;
        tab.push_back(e);
      }
    }
    n->pack_tab.ensure(tab.size() * sizeof(PackEntry));
    CUDA_CHECK(cudaMemcpy(n->pack_tab.p, tab.data(), tab.size() * sizeof(PackEntry), cudaMemcpyHostToDevice));
    n->pack_nent = (int)tab.size(); n->pack_tiles = tiles;
    // extra packed copies (bf16 mode) through index tables: the first layer in im2col K order
    // (k = source*64 + tap*4 + channel slot); the generator head as the two single-tap GEMM operands
    // of its cols formulation (n = tap*4 + channel slot)
    auto add_gather = [&](const std::vector<int>& idx, DevBuf& dstbuf, int dt) {
      n->gathers.emplace_back();
      gan_net::GatherTab& gt = n->gathers.back();
      gt.dt = dt;
      dstbuf.ensure(idx.size() * ctx->esize());
      gt.idx.ensure(idx.size() * sizeof(int));
      CUDA_CHECK(cudaMemcpy(gt.idx.p, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice));
      gt.dst = dstbuf.p; gt.n = (int)idx.size();
    };
    Layer& l0 = n->layers[0];
    if (l0.first && ctx->dt == DT_BF16 && l0.src_c <= 4) {
      const int C = l0.src_c, Kt = l0.nsrc * 64;
      std::vector<int> idx((size_t)l0.Cout_p * Kt, -1);
      for (int nn = 0; nn < l0.Cout; ++nn)
        for (int sidx = 0; sidx < l0.nsrc; ++sidx)
          for (int t = 0; t < 16; ++t)
            for (int c = 0; c < C; ++c)
              idx[(size_t)nn * Kt + sidx * 64 + t * 4 + c] = (int)(l0.w_off + ((int64_t)t * l0.Cin + sidx * C + c) * l0.Cout + nn);
      add_gather(idx, l0.wp_im2col, ctx->dtA);
    }
    if (!n->is_gen && l0.first && ctx->dt == DT_BF16 && l0.src_c <= 4 && l0.Cout == 64) {
      // data gradient of the first layer w.r.t. the image the generator produced (the `tar` half when the
      // discriminator takes (input, target)) as per-tap products: B[n = tap*4 + c][k = co] = W[kh, kw, c0 + c, co]
      const int C = l0.src_c, c0 = l0.nsrc == 2 ? C : 0;
      std::vector<int> dc((size_t)64 * 64, -1);
      for (int t = 0; t < 16; ++t)
        for (int c = 0; c < C; ++c)
          for (int co = 0; co < 64; ++co)
            dc[(size_t)(t * 4 + c) * 64 + co] = (int)(l0.w_off + ((int64_t)t * l0.Cin + c0 + c) * l0.Cout + co);
      add_gather(dc, l0.wp_dcols, ctx->dtG);
    }
    Layer& lh = n->layers.back();
    if (n->is_gen && lh.head && ctx->dt == DT_BF16 && lh.Cout <= 4 && lh.Cin % 64 == 0) {
      const int C = lh.Cout, Ci = lh.Cin;           // master (kh,kw,co,ci): row tap*C+co, column ci
      std::vector<int> fc((size_t)64 * Ci, -1), dc((size_t)Ci * 64, -1);
      for (int t = 0; t < 16; ++t)
        for (int co = 0; co < C; ++co)
          for (int ci = 0; ci < Ci; ++ci) {
            const int m = (int)(lh.w_off + ((int64_t)t * C + co) * Ci + ci);
            fc[(size_t)(t * 4 + co) * Ci + ci] = m;      // forward  B[n = tap*4+co][ci]
            dc[(size_t)ci * 64 + t * 4 + co] = m;        // dgrad    B[n = ci][k = tap*4+co]
          }
      add_gather(fc, lh.wp_cols, ctx->dtA);      // forward operand: multiplies activations
      add_gather(dc, lh.wp_dcols, ctx->dtG);     // data-gradient operand: multiplies gradients
    }
    if (!n->is_gen && lh.head && ctx->dt == DT_BF16 && lh.Cout == 1 && lh.Cin % 64 == 0) {
      const int Ci = lh.Cin;                        // master (kh,kw,ci,1): element tap*Ci + ci
      std::vector<int> fc((size_t)64 * Ci, -1), dc((size_t)Ci * 64, -1);
      for (int t = 0; t < 16; ++t)
        for (int ci = 0; ci < Ci; ++ci) {
          const int m = (int)(lh.w_off + (int64_t)t * Ci + ci);
          fc[(size_t)(t * 4) * Ci + ci] = m;        // forward  B[n = tap*4][ci]
          dc[(size_t)ci * 64 + t * 4] = m;          // dgrad    B[n = ci][k = tap*4]
        }
      add_gather(fc, lh.wp_cols, ctx->dtA);      // forward operand: multiplies activations
      add_gather(dc, lh.wp_dcols, ctx->dtG);     // data-gradient operand: multiplies gradients
    }
  }
  launch_pack_multi(ctx->L(), ctx->dt, (const PackEntry*)n->pack_tab.p, n->pack_nent, n->pack_tiles);
  for (auto& gt : n->gathers) launch_gather_pack(ctx->L(), gt.dt, n->params.as<float>(), gt.idx.as<int>(), gt.n, gt.dst);
  n->packed_dirty = false;
}

// ---------------------------------------------------------------------------------------------
// one conv(+norm+act) layer, forward and backward
// ---------------------------------------------------------------------------------------------
// zero-initialised ticket counters of the "last block" sums (self-resetting), one set per stream
static unsigned int* tickets(gan_ctx* ctx) {
  ctx->sc().counters.ensure(4096);
  return ctx->sc().counters.as<unsigned int>();
}
static DropKey drop_key(gan_ctx* ctx, const Layer& ly, const Slot& s) {
  DropKey k;
  k.seed_lo = (uint32_t)(ctx->seed & 0xffffffffu); k.seed_hi = (uint32_t)(ctx->seed >> 32);
  k.call_dev = ctx->call_dev.as<uint32_t>(); k.call_off = s.call_off; k.layer = (uint32_t)ly.tag; k.sample0 = s.sample0;
  k.enabled = (ly.dropout && ctx->dropout_enabled) ? 1 : 0;
  return k;
}

// First layers: rows built in shared memory by k_conv_first_fwd (conv_first.cu), activation fused.
static bool first_kernel_ok(const gan_ctx* ctx, const Layer& ly, const Slot& s, int H, int W) {
  if (!ly.first || ctx->dt != DT_BF16 || ctx->engine == GAN_ENGINE_FFMA || ly.norm != NORM_NONE || ly.act != ACT_LEAKY) return false;
  if (ly.Cout != 64 || ly.wp_im2col.p == nullptr || s.src_f32[0] == nullptr) return false;
  FirstLayerOp op; memset(&op, 0, sizeof(op));
  op.src[0] = s.src_f32[0]; op.src[1] = s.src_f32[1];
  op.nsrc = ly.nsrc; op.C = ly.src_c; op.H = H; op.W = W; op.a_pitch = 8; op.a_coff = 0; op.dt = ctx->dtA;
  if (ly.nsrc == 2 && s.src_f32[1] == nullptr) return false;
  return first_fwd_supported(op);
}

static void layer_forward(gan_net* n, Slot& s, int li, View in, View out) {
  gan_ctx* ctx = n->ctx;
  Layer& ly = n->layers[li];
  int Ho, Wo; out_dims(ly.kind, in.H, in.W, Ho, Wo);
  const int B = in.N;
  s.in_views[li] = in; s.out_views[li] = out;
  if (ly.head && cols_on(ctx, n, ly)) {
    s.used_cols = true;
    // cols rows: fp16 with fp16 storage (rounding 5e-4 of each of the four summed taps; the 268 MB fp32 round trip at
    // batch 64 was the largest intermediate of the step), fp32 with bf16 storage (a bf16 pre-rounding would cost 4e-3)
    const bool cols16 = ctx->dtA == DT_F16;
    s.cols.ensure((size_t)B * in.H * in.W * 64 * 4);
    View cols = make_view(cols16 ? s.cols.p : nullptr, B, in.H, in.W, 64);
    ConvOp cop = make_op_1tap(in, ly.Cin, cols, 64, 64, ly.wp_cols.p, ctx->dtA, ctx->dtA);
    if (!cols16) cop.out_rows_f32 = s.cols.as<float>();
    cop.real_n = 16 * ly.Cout;
    run_conv_fwd(ctx, cop);
    launch_col2im_tanh(ctx->L(), cols16 ? DT_F16 : DT_F32, s.cols.p, n->params.as<float>() + ly.bias_off, B, in.H, in.W, ly.Cout, (float*)out.p);
    return;
  }
  if (ly.head && dcols_on(ctx, n, ly)) {
    // discriminator head in cols form: the 512-channel activation is read once instead of once per tap
    s.used_cols = true;
    s.cols.ensure((size_t)B * in.H * in.W * 64 * 4);
    View cols = make_view(nullptr, B, in.H, in.W, 64);
    ConvOp cop = make_op_1tap(in, ly.Cin, cols, 64, 64, ly.wp_cols.p, ctx->dtA, ctx->dtA);
    cop.out_rows_f32 = s.cols.as<float>();
    cop.real_n = 16 * ly.Cout;
    run_conv_fwd(ctx, cop);
    launch_dhead_gather(ctx->L(), s.cols.as<float>(), n->params.as<float>() + ly.bias_off, B, in.H, in.W, (float*)out.p);
    return;
  }
  if (ly.head) {
    s.used_cols = false;
    // generator head: bias + tanh -> fp32 image; discriminator head: bias -> fp32 logits
    View y = make_view(nullptr, B, Ho, Wo, ly.Cout_p);
    ConvOp op = make_op(ctx, ly, R_FWD, in, y, ly.wp_fwd.p);
    op.bias = n->params.as<float>() + ly.bias_off;
    op.epi = ly.act == ACT_TANH ? EPI_BIAS_TANH : EPI_BIAS;
    op.out_f32 = (float*)out.p;
    run_conv_fwd(ctx, op);
    return;
  }
  int64_t P = (int64_t)B * Ho * Wo;
  if (li == 0) s.z_is_act0 = false;
  s.z[li].ensure((size_t)P * ly.Cout * ctx->esize());
  View z = make_view(s.z[li].p, B, Ho, Wo, ly.Cout);
  if (li == 0 && s.used_im2col && first_kernel_ok(ctx, ly, s, in.H, in.W) && out.pitch % 8 == 0 && out.coff % 8 == 0) {
    FirstLayerOp op; memset(&op, 0, sizeof(op));
    op.src[0] = s.src_f32[0]; op.src[1] = s.src_f32[1]; op.nsrc = ly.nsrc; op.C = ly.src_c; op.B = B; op.H = in.H; op.W = in.W;
    op.wpack = ly.wp_im2col.p; op.a = out.p; op.a_pitch = out.pitch; op.a_coff = out.coff; op.dt = ctx->dtA;
    s.z_is_act0 = true;                    // z is not stored: the backward pass reads the sign from the activation view
    ProfScope ps(ctx, FAM_UMMA_FWD, 2.0 * (double)P * 64.0 * 16.0 * ly.Cin);
    launch_conv_first_fwd(ctx->L(), op);
    return;
  }
  const int G = ly.norm == NORM_BATCH ? 1 : B;
  const int64_t Pg = P / G;
  const size_t gc = (size_t)G * ly.Cout;
  if (ly.norm != NORM_NONE) {
    size_t need = stats_ws_floats(G, Pg, ly.Cout), epi = (size_t)STATS_MAX_CHUNKS * 2 * ly.Cout;
    ctx->sc().stats_ws.ensure((need > epi ? need : epi) * 4);
  }
  ConvOp cop = (li == 0 && s.used_im2col) ? make_op_im2col(ctx, ly, R_FWD, s, z) : make_op(ctx, ly, R_FWD, in, z, ly.wp_fwd.p);
  // BatchNorm statistics straight from the fp32 accumulators (conv epilogue), unless the whole-layer kernel takes the layer
  if (ly.norm == NORM_BATCH && !bn_small_fwd_fits(1, P)) cop.stats_ws = ctx->sc().stats_ws.as<float>();
  const int stat_parts = run_conv_fwd(ctx, cop);
  DropKey dk = drop_key(ctx, ly, s);
  // SURVEY 8d byte model: forward = read z + write activation = 2*s per element (statistics belong to the conv epilogue)
  ProfScope ps(ctx, FAM_NORM, (double)P * ly.Cout * ctx->esize() * 2);
  if (ly.norm == NORM_NONE) {
    launch_norm_apply(ctx->L(), ctx->dtA, z.p, P, P, 1, Ho * Wo, ly.Cout, nullptr, nullptr, nullptr, ly.act, dk, out.p,
                      out.pitch, out.coff);
    return;
  }
  s.stats[li].ensure(gc * 6 * 4);
  float* st = s.stats[li].as<float>();
  float* pr = n->params.as<float>();
  float* mm = (ly.mov_off >= 0) ? n->mov.as<float>() + ly.mov_off : nullptr;
  const float eps = ly.norm == NORM_BATCH ? BN_EPS : IN_EPS;
  if (stat_parts > 0) {
    launch_norm_stats_finalize(ctx->L(), ctx->sc().stats_ws.as<float>(), stat_parts, P, ly.Cout, eps, pr + ly.g_off, pr + ly.b_off,
                               st, st + gc, st + 2 * gc, st + 3 * gc, mm, mm ? mm + ly.Cout : nullptr, BN_MOMENTUM);
  } else {
    if (launch_bn_small_fwd(ctx->L(), ctx->dtA, z.p, P, G, Ho * Wo, ly.Cout, eps, pr + ly.g_off, pr + ly.b_off, st, st + gc,
                            st + 2 * gc, st + 3 * gc, mm, mm ? mm + ly.Cout : nullptr, BN_MOMENTUM, ly.act, dk, out.p,
                            out.pitch, out.coff))
      return;
    launch_norm_stats(ctx->L(), ctx->dtA, z.p, G, Pg, ly.Cout, ctx->sc().stats_ws.as<float>(), eps, pr + ly.g_off, pr + ly.b_off, st,
                      st + gc, st + 2 * gc, st + 3 * gc, mm, mm ? mm + ly.Cout : nullptr, BN_MOMENTUM);
  }
  launch_norm_apply(ctx->L(), ctx->dtA, z.p, P, Pg, G, Ho * Wo, ly.Cout, st, st + 2 * gc, st + 3 * gc, ly.act, dk, out.p,
                    out.pitch, out.coff);
}

// d1/d2: gradient w.r.t. the layer's activated output.  din: where the gradient w.r.t. the layer
// input goes (p == nullptr: not needed).  For head layers d1 is dz itself (already in d1.p, compact).
static void layer_backward(gan_net* n, Slot& s, int li, GradSrc d1, GradSrc d2, View din, bool want_wgrad) {
  gan_ctx* ctx = n->ctx;
  Layer& ly = n->layers[li];
  View in = s.in_views[li];
  int Ho, Wo; out_dims(ly.kind, in.H, in.W, Ho, Wo);
  const int B = in.N;
  const int64_t P = (int64_t)B * Ho * Wo;
  View dz;
  if (ly.head) {
    dz = make_view((void*)d1.p, B, Ho, Wo, ly.Cout_p, d1.pitch, d1.coff);
  } else {
    ctx->sc().dz_scratch.ensure((size_t)P * ly.Cout * ctx->esize());
    dz = make_view(ctx->sc().dz_scratch.p, B, Ho, Wo, ly.Cout);
    const int G = ly.norm == NORM_BATCH ? 1 : (ly.norm == NORM_INSTANCE ? B : 1);
    const int64_t Pg = P / G;
    const size_t gc = (size_t)G * ly.Cout;
    float* st = ly.norm != NORM_NONE ? s.stats[li].as<float>() : nullptr;
    float* gr = want_wgrad ? n->grads.as<float>() : nullptr;
    ctx->sc().junk.ensure(2048 * 4);
    float* dgamma = (gr && ly.norm != NORM_NONE) ? gr + ly.g_off : ctx->sc().junk.as<float>();
    float* dbeta = (gr && ly.norm != NORM_NONE) ? gr + ly.b_off : ctx->sc().junk.as<float>() + 1024;
    if (ly.norm != NORM_NONE) ctx->sc().stats_ws.ensure(stats_ws_floats(G, Pg, ly.Cout) * 4);
    ProfScope ps(ctx, FAM_NORM, (double)P * ly.Cout * ctx->esize() * (ly.norm == NORM_NONE ? 3 : 5));
    const bool from_act = li == 0 && s.z_is_act0;           // first-layer kernel: sign(z) == sign(LeakyReLU(z))
    const View av = s.out_views[li];
    launch_norm_bwd(ctx->L(), ctx->dtA, ctx->dtG, from_act ? av.p : s.z[li].p, d1, d2, P, Pg, G, Ho * Wo, ly.Cout, ly.norm, st,
                    st ? st + gc : nullptr, st ? st + 2 * gc : nullptr, st ? st + 3 * gc : nullptr, ly.act, drop_key(ctx, ly, s),
                    ctx->sc().stats_ws.as<float>(), st ? st + 4 * gc : nullptr, st ? st + 5 * gc : nullptr, dgamma, dbeta, dz.p,
                    tickets(ctx), from_act ? av.pitch : 0, from_act ? av.coff : 0);
  }
  bool first_wgrad_done = false;
  if (want_wgrad && li == 0 && s.used_im2col && first_kernel_ok(ctx, ly, s, in.H, in.W)) {
    // first layer: the rows are rebuilt in shared memory from the fp32 image(s) (conv_first.cu), no im2col rows in HBM
    FirstWgradOp fw; memset(&fw, 0, sizeof(fw));
    fw.src[0] = s.src_f32[0]; fw.src[1] = s.src_f32[1]; fw.nsrc = ly.nsrc; fw.C = ly.src_c; fw.B = B; fw.H = in.H; fw.W = in.W;
    fw.dz = dz.p; fw.dz_pitch = dz.pitch; fw.dz_coff = dz.coff; fw.dt = ctx->dtG;
    fw.dW = n->grads.as<float>() + ly.w_off; fw.s_tap = (long long)ly.Cin * ly.Cout; fw.s_k = ly.Cout; fw.s_n = 1;
    ctx->sc().wgrad_ws.ensure((size_t)24 << 20);
    fw.ws = ctx->sc().wgrad_ws.as<float>(); fw.ws_bytes = ctx->sc().wgrad_ws.bytes;
    if (ctx->dtG == ctx->dtA && first_wgrad_supported(fw)) {
      fw.accumulate = ly.wgrad_epoch == ctx->step_epoch ? 1 : 0;
      ly.wgrad_epoch = ctx->step_epoch;
      ProfScope ps(ctx, FAM_UMMA_WGRAD, 2.0 * (double)P * 16.0 * ly.Cin * ly.Cout);
      launch_conv_first_wgrad(ctx->L(), fw);
      first_wgrad_done = true;
    }
  }
  if (want_wgrad && !first_wgrad_done) {
    if (li == 0 && s.used_im2col) {       // rows of the input image(s) for the weight-gradient GEMM (cached per image and step)
      for (int k = 0; k < ly.nsrc; ++k)
        if (s.src_f32[k] != nullptr) s.im2col[k] = cached_im2col(ctx, s.src_f32[k], B, in.H, in.W, ly.src_c);
    }
    ConvOp op = (li == 0 && s.used_im2col) ? make_op_im2col(ctx, ly, R_WGRAD, s, dz) : make_op(ctx, ly, R_WGRAD, in, dz, nullptr);
    op.dW = n->grads.as<float>() + ly.w_off;
    run_conv_wgrad(ctx, ly, op);
  }
  if (din.p != nullptr) {
    GAN_REQUIRE(ly.need_dgrad, "dgrad weights not packed");
    run_conv_fwd(ctx, make_op(ctx, ly, R_DGRAD, din, dz, ly.wp_dgrad.p));
  }
}

// ---------------------------------------------------------------------------------------------
// generator / discriminator sweeps
// ---------------------------------------------------------------------------------------------
static void slot_prepare(gan_net* n, Slot& s, int B, int H, int W) {
  size_t nl = n->layers.size();
  if (s.z.size() != nl) {
    s.z.resize(nl); s.stats.resize(nl); s.in_views.resize(nl); s.out_views.resize(nl);
    if (n->is_gen) { s.cat.resize(7); s.dcat.resize(7); s.dskip.resize(8); }
    else { s.act.resize(4); s.dact.resize(4); }
  }
  s.B = B; s.H = H; s.W = W;
}

// x_f32: device fp32 (B,H,W,C).  Output: s.out_f32 (B,H,W,C) fp32.
// call_off >= 0: position of this call in the step's generator-call sequence (keys the dropout masks) when the host
// issues the calls in another order than the reference (two-stream CycleGAN step); -1: next in issue order.
static void generator_forward(gan_net* g, int slot, const float* x_f32, int B, int H, int W, int call_off = -1) {
  gan_ctx* ctx = g->ctx;
  GAN_REQUIRE(H % 256 == 0 && W % 256 == 0 && H >= 256 && W >= 256, "generator needs H,W multiples of 256");
  pack_weights(g);
  Slot& s = g->slots[slot];
  slot_prepare(g, s, B, H, W);
  if (call_off < 0) s.call_off = ctx->gen_calls_pending++;
  else { s.call_off = (uint32_t)call_off; if (ctx->gen_calls_pending < (uint32_t)call_off + 1) ctx->gen_calls_pending = (uint32_t)call_off + 1; }
  s.call_id = ctx->call_counter + s.call_off;
  s.sample0 = ctx->sample0_set ? ctx->sample0 : (int64_t)ctx->rank * B;
  const size_t es = ctx->esize();
  const int C = g->C;
  const int Cp = g->Cp;
  s.used_im2col = im2col_on(ctx, g->layers[0]);
  s.xin.ensure((size_t)B * H * W * Cp * es);
  s.src_f32[0] = x_f32; s.src_f32[1] = nullptr;
  if (s.used_im2col) {
    // rows in HBM only where the first-layer kernel (rows built in shared memory) cannot run; the weight-gradient pass
    // fetches them lazily
    if (!first_kernel_ok(ctx, g->layers[0], s, H, W)) s.im2col[0] = cached_im2col(ctx, x_f32, B, H, W, C);
  } else {
    launch_convert(ctx->L(), ctx->dtA, x_f32, (int64_t)B * H * W, C, s.xin.p, Cp, 0);
  }
  // concat buffers: cat[k-1] = [up_k output (UP_F[k-1]) | down_{8-k} output (DOWN_F[7-k])] at H/2^(8-k)
  for (int k = 1; k <= 7; ++k) {
    int hs = H >> (8 - k), ws = W >> (8 - k);
    s.cat[k - 1].ensure((size_t)B * hs * ws * (UP_F[k - 1] + DOWN_F[7 - k]) * es);
  }
  s.d8.ensure((size_t)B * (H >> 8) * (W >> 8) * 512 * es);
  View in = make_view(s.xin.p, B, H, W, Cp);
  for (int j = 1; j <= 8; ++j) {
    int hs = H >> j, ws = W >> j;
    View out;
    if (j < 8) { int k = 8 - j; out = make_view(s.cat[k - 1].p, B, hs, ws, DOWN_F[j - 1], UP_F[k - 1] + DOWN_F[j - 1], UP_F[k - 1]); }
    else out = make_view(s.d8.p, B, hs, ws, 512);
    layer_forward(g, s, j - 1, in, out);
    in = out;
  }
  for (int k = 1; k <= 7; ++k) {
    int hs = H >> (8 - k), ws = W >> (8 - k);
    int pitch = UP_F[k - 1] + DOWN_F[7 - k];
    View out = make_view(s.cat[k - 1].p, B, hs, ws, UP_F[k - 1], pitch, 0);
    layer_forward(g, s, 7 + k, in, out);
    in = make_view(s.cat[k - 1].p, B, hs, ws, pitch, pitch, 0);     // Concatenate([x, skip]) base_gan.py:221
  }
  s.out_f32.ensure((size_t)B * H * W * C * 4);
  layer_forward(g, s, 15, in, make_view(s.out_f32.p, B, H, W, C));
}

// Data parallel: hand the gradient range [from, hi) of a net to the communication stream, hi = the lowest offset
// already handed over in this step (buckets are issued top-down, in backward-completion order).  Offsets are rounded UP
// to the bucket quantum 4*world so that every bucket divides evenly over the ranks; the few floats below a rounded
// boundary travel with the next (later-finishing) bucket, the padding beyond nparams is zeros.  With the sharded
// optimizer the bucket is reduce-scattered and remembered (rank r then owns sub-range r); otherwise it is all-reduced.
static void comm_reduce_bucket(gan_net* n, int64_t from) {
  gan_ctx* ctx = n->ctx;
  if (ctx->world <= 1) return;
  const int64_t q = 4LL * ctx->world;
  GAN_REQUIRE(q <= 1024, "world size above 256 is not supported");
  const int64_t total = (n->nparams + q - 1) / q * q;
  const int64_t hi = n->reduced_from < 0 ? total : n->reduced_from;
  const int64_t lo = from <= 0 ? 0 : (from + q - 1) / q * q;
  if (lo >= hi) return;
  float* gr = n->grads.as<float>();
  if (ctx->shard_optimizer) {
    comm_reducescatter_async(ctx, gr + lo, hi - lo);
    n->bucket_off.push_back(lo); n->bucket_len.push_back(hi - lo);
  } else if (ctx->comm16) {
    // bf16 on the wire: the bucket is rounded once (fp32 range, 8-bit mantissa: below the 16-bit operand noise the
    // gradient already carries) into the communication buffer, summed there, and Adam reads the sum from it
    n->grads16.ensure((size_t)total * 2);
    uint16_t* g16 = n->grads16.as<uint16_t>();
    launch_grad_to_bf16(ctx->L(), gr + lo, g16 + lo, hi - lo);
    comm_allreduce_bf16_async(ctx, g16 + lo, hi - lo);
    if (lo == 0) n->grads16_valid = true;
  } else {
    comm_allreduce_async(ctx, gr + lo, hi - lo);
  }
  n->reduced_from = lo;
}
static void comm_step_begin(gan_net* n) { n->bucket_off.clear(); n->bucket_len.clear(); n->reduced_from = -1; n->grads16_valid = false; }

// Backward through one generator call.  d1/d2: extra gradient sources w.r.t. the tanh output
// (activation dtype); ref/l1_coef: + l1_coef*sign(out-ref).  Accumulates into g->grads.
// `final_call`: this is the last contribution to g->grads in the step, so finished gradient ranges
// can be all-reduced on the communication stream while the rest of the sweep still runs (buckets in
// backward-completion order: [up5..last], [up1..up4], [down5..down8], [down1..down4]).
static void generator_backward(gan_net* g, int slot, GradSrc d1, GradSrc d2, const float* ref_f32, float l1_coef,
                               bool want_input_grad, bool final_call = false) {
  gan_ctx* ctx = g->ctx;
  Slot& s = g->slots[slot];
  const int B = s.B, H = s.H, W = s.W, C = g->C;
  const size_t es = ctx->esize();
  float* gr = g->grads.as<float>();
  // head
  const int Cp = g->Cp;
  ctx->sc().head_part.ensure((size_t)HEAD_PART_BLOCKS * 4 * 4);
  if (s.used_cols) {
    s.gcols.ensure((size_t)B * (H / 2) * (W / 2) * 64 * 2);
    launch_ghead_bwd_cols(ctx->L(), ctx->dtG, s.out_f32.as<float>(), ref_f32, d1, d2, l1_coef, B, H, W, C, s.gcols.p,
                          gr + g->layers[15].bias_off, ctx->sc().head_part.as<float>(), tickets(ctx));
  } else {
    s.dlogit.ensure((size_t)B * H * W * Cp * es);
    launch_ghead_bwd(ctx->L(), ctx->dtG, s.out_f32.as<float>(), ref_f32, d1, d2, l1_coef, (int64_t)B * H * W, C, s.dlogit.p, Cp,
                     gr + g->layers[15].bias_off, ctx->sc().head_part.as<float>(), tickets(ctx));
  }
  for (int k = 1; k <= 7; ++k) {
    int hs = H >> (8 - k), ws = W >> (8 - k);
    s.dcat[k - 1].ensure((size_t)B * hs * ws * (UP_F[k - 1] + DOWN_F[7 - k]) * es);
  }
  {
    int pitch = UP_F[6] + DOWN_F[0];
    View din = make_view(s.dcat[6].p, B, H / 2, W / 2, pitch);
    Layer& lh = g->layers[15];
    if (s.used_cols) {
      // G = im2col(dz): [B*(H/2)*(W/2)][tap*C+co]; dW (kh,kw,co,ci) = G^T x; dx = G f
      View x = s.in_views[15];
      View G = make_view(s.gcols.p, B, H / 2, W / 2, 64);
      ConvOp wg = make_op_1tap(x, lh.Cin, G, 64, 64, nullptr, ctx->dtA, ctx->dtG);
      wg.dW = gr + lh.w_off; wg.s_tap = 0; wg.s_k = 1; wg.s_n = lh.Cin; wg.n_slot4_c = C;
      wg.real_n = 16 * C;
      run_conv_wgrad(ctx, lh, wg);
      ConvOp dg = make_op_1tap(G, 64, din, lh.Cin, lh.Cin, lh.wp_dcols.p, ctx->dtG, ctx->dtG);
      dg.real_k = 16 * C;
      run_conv_fwd(ctx, dg);
    } else {
      layer_backward(g, s, 15, GradSrc{s.dlogit.p, Cp, 0}, GradSrc{nullptr, 0, 0}, din, true);
    }
  }
  s.dd8.ensure((size_t)B * (H >> 8) * (W >> 8) * 512 * es);
  for (int k = 7; k >= 1; --k) {
    int pitch = UP_F[k - 1] + DOWN_F[7 - k];
    GradSrc src{s.dcat[k - 1].p, pitch, 0};
    View din;
    if (k == 1) din = make_view(s.dd8.p, B, H >> 8, W >> 8, 512);
    else { int pp = UP_F[k - 2] + DOWN_F[8 - k]; din = make_view(s.dcat[k - 2].p, B, H >> (9 - k), W >> (9 - k), pp); }
    layer_backward(g, s, 7 + k, src, GradSrc{nullptr, 0, 0}, din, true);
    if (final_call && ctx->world > 1) {
      if (k == 5) comm_reduce_bucket(g, g->layers[12].w_off);       // up5 .. last
      if (k == 1) comm_reduce_bucket(g, g->layers[8].w_off);        // up1 .. up4 (29 M of the 54 M parameters)
    }
  }
  for (int j = 8; j >= 1; --j) {
    GradSrc a, b{nullptr, 0, 0};
    if (j == 8) a = GradSrc{s.dd8.p, 512, 0};
    else {
      int k = 8 - j;
      a = GradSrc{s.dskip[j].p, DOWN_F[j - 1], 0};                              // from down_{j+1} dgrad
      b = GradSrc{s.dcat[k - 1].p, UP_F[k - 1] + DOWN_F[j - 1], UP_F[k - 1]};   // skip half of the concat gradient
    }
    View din = make_view(nullptr, B, H >> (j - 1), W >> (j - 1), j > 1 ? DOWN_F[j - 2] : Cp);
    if (j > 1) {
      s.dskip[j - 1].ensure((size_t)B * din.H * din.W * din.C * es);
      din.p = s.dskip[j - 1].p;
    } else if (want_input_grad) {
      s.dxin.ensure((size_t)B * H * W * Cp * es);
      din.p = s.dxin.p;
    }
    layer_backward(g, s, j - 1, a, b, din, true);
    // down5..down8 hold 16.8 M of the down path's 19.5 M parameters: reduce them under down4..down1
    if (final_call && ctx->world > 1 && j == 5) comm_reduce_bucket(g, g->layers[4].w_off);
  }
  if (final_call && ctx->world > 1) comm_reduce_bucket(g, 0);
}

// inp/tar: device fp32 (B,H,W,C); tar may be nullptr when target == false.
static void discriminator_forward(gan_net* d, int slot, const float* inp, const float* tar, int B, int H, int W) {
  gan_ctx* ctx = d->ctx;
  GAN_REQUIRE(H % 8 == 0 && W % 8 == 0 && H >= 32 && W >= 32, "discriminator needs H,W multiples of 8");
  GAN_REQUIRE((tar != nullptr) == d->target, "discriminator target input mismatch");
  pack_weights(d);
  Slot& s = d->slots[slot];
  slot_prepare(d, s, B, H, W);
  const size_t es = ctx->esize();
  const int C = d->C, C0 = d->Cin0_p;
  s.in0.ensure((size_t)B * H * W * C0 * es);
  s.used_im2col = im2col_on(ctx, d->layers[0]);
  s.src_f32[0] = inp; s.src_f32[1] = tar;
  if (s.used_im2col) {                                   // one im2col K-block per source of concatenate([inp, tar])
    if (!first_kernel_ok(ctx, d->layers[0], s, H, W)) {
      s.im2col[0] = cached_im2col(ctx, inp, B, H, W, C);
      if (tar) s.im2col[1] = cached_im2col(ctx, tar, B, H, W, C);
    }
  } else {
    launch_convert(ctx->L(), ctx->dtA, inp, (int64_t)B * H * W, C, s.in0.p, C0, 0);      // concatenate([inp, tar]) base_gan.py:139
    if (tar) launch_convert(ctx->L(), ctx->dtA, tar, (int64_t)B * H * W, C, s.in0.p, C0, C);
  }
  View in = make_view(s.in0.p, B, H, W, C0);
  int h = H, w = W;
  for (int li = 0; li < 4; ++li) {
    int ho, wo; out_dims(d->layers[li].kind, h, w, ho, wo);
    s.act[li].ensure((size_t)B * ho * wo * d->layers[li].Cout * es);
    View out = make_view(s.act[li].p, B, ho, wo, d->layers[li].Cout);
    layer_forward(d, s, li, in, out);
    in = out; h = ho; w = wo;
  }
  s.logits.ensure((size_t)B * (h - 1) * (w - 1) * 4);
  layer_forward(d, s, 4, in, make_view(s.logits.p, B, h - 1, w - 1, 1));
}

static bool g_in_cols = [] { const char* e = getenv("GAN_B200_DIN_COLS"); return !(e && e[0] == '0'); }();   // dev A/B switch
static bool in_cols_on(const gan_ctx* ctx, const gan_net* d) {
  const Layer& l0 = d->layers[0];
  return g_in_cols && !d->is_gen && ctx->dt == DT_BF16 && ctx->engine != GAN_ENGINE_FFMA && l0.first && l0.kind == K_CONV_S2 &&
         l0.src_c <= 4 && l0.Cout == 64 && l0.norm == NORM_NONE && l0.wp_dcols.p != nullptr;
}

// Backward from s.dlogit (already filled).  want_wgrad: accumulate into d->grads.
static void discriminator_backward(gan_net* d, int slot, bool want_wgrad, bool want_input_grad) {
  gan_ctx* ctx = d->ctx;
  Slot& s = d->slots[slot];
  const size_t es = ctx->esize();
  const int B = s.B;
  for (int li = 4; li >= 0; --li) {
    View in = s.in_views[li];
    View din = make_view(nullptr, B, in.H, in.W, in.C);
    if (li > 0) { s.dact[li - 1].ensure((size_t)B * in.H * in.W * in.C * es); din.p = s.dact[li - 1].p; }
    else if (want_input_grad) { s.din0.ensure((size_t)B * in.H * in.W * in.C * es); din.p = s.din0.p; }
    GradSrc src = (li == 4) ? GradSrc{s.dlogit.p, d->layers[4].Cout_p, 0} : GradSrc{s.dact[li].p, d->layers[li].Cout, 0};
    if (li == 4 && s.used_cols) {
      // Gd = stride-1 unfold of dlogit over the 31x31 activation grid; dW = Gd^T a; da = Gd w
      Layer& lh = d->layers[4];
      s.gcols.ensure((size_t)B * in.H * in.W * 64 * 2);
      launch_dhead_unfold(ctx->L(), s.dlogit.p, lh.Cout_p, B, in.H, in.W, s.gcols.p);
      View G = make_view(s.gcols.p, B, in.H, in.W, 64);
      if (want_wgrad) {
        ConvOp wg = make_op_1tap(in, lh.Cin, G, 64, 64, nullptr, ctx->dtA, ctx->dtG);
        wg.dW = d->grads.as<float>() + lh.w_off; wg.s_tap = 0; wg.s_k = 1; wg.s_n = lh.Cin; wg.n_slot4_c = 1;
        wg.real_n = 16;
        run_conv_wgrad(ctx, lh, wg);
      }
      ConvOp dg = make_op_1tap(G, 64, din, lh.Cin, lh.Cin, lh.wp_dcols.p, ctx->dtG, ctx->dtG);
      dg.real_k = 16;
      run_conv_fwd(ctx, dg);
      continue;
    }
    if (li == 0 && want_input_grad) { s.din0_pitch = d->Cin0_p; s.din0_coff = d->layers[0].nsrc == 2 ? d->C : 0; }
    if (li == 0 && want_input_grad && in_cols_on(ctx, d)) {
      // dL/d(generated image) = per-tap products dz . W[tap] (ONE 1x1 GEMM over dz instead of 4 classes x 4 shifted
      // taps with a 16-channel padded output tile) gathered by k_col2im_grad into compact 4-channel rows
      Layer& l0 = d->layers[0];
      din.p = nullptr;
      layer_backward(d, s, li, src, GradSrc{nullptr, 0, 0}, din, want_wgrad);      // leaves dz in the stream's scratch
      int Ho, Wo; out_dims(l0.kind, in.H, in.W, Ho, Wo);
      View dzv = make_view(ctx->sc().dz_scratch.p, B, Ho, Wo, l0.Cout);
      s.dcols0.ensure((size_t)B * Ho * Wo * 64 * 2);
      View colsv = make_view(s.dcols0.p, B, Ho, Wo, 64);
      ConvOp dg = make_op_1tap(dzv, l0.Cout, colsv, 64, 64, l0.wp_dcols.p, ctx->dtG, ctx->dtG);
      dg.real_n = 16 * l0.src_c;
      run_conv_fwd(ctx, dg);
      s.din0.ensure((size_t)B * in.H * in.W * 4 * 2);
      launch_col2im_grad(ctx->L(), ctx->dtG, s.dcols0.p, B, Ho, Wo, s.din0.p);
      s.din0_pitch = 4; s.din0_coff = 0;
      continue;
    }
    layer_backward(d, s, li, src, GradSrc{nullptr, 0, 0}, din, want_wgrad);
  }
}

static int64_t logits_count(const Slot& s) { return (int64_t)s.B * (s.H / 8 - 2) * (s.W / 8 - 2); }

// BCE on a discriminator call's logits: loss partial into `slot_idx`; when dz: s.dlogit = coef*(sigmoid-label)/n
static void disc_bce(gan_net* d, int slot, float label, float coef, bool make_dz, bool bias_grad, int loss_slot) {
  gan_ctx* ctx = d->ctx;
  Slot& s = d->slots[slot];
  int64_t n = logits_count(s);
  const int dzp = d->layers[4].Cout_p;
  if (make_dz) s.dlogit.ensure((size_t)n * dzp * ctx->esize());
  ctx->sc().head_part.ensure((size_t)HEAD_PART_BLOCKS * 4 * 4);
  launch_bce(ctx->L(), ctx->dtG, s.logits.as<float>(), n, label, coef, make_dz ? s.dlogit.p : nullptr, dzp,
             (make_dz && bias_grad) ? d->grads.as<float>() + d->layers[4].bias_off : nullptr, ctx->loss_ws.as<float>(),
             loss_slot, ctx->sc().head_part.as<float>(), tickets(ctx));
}

// ---------------------------------------------------------------------------------------------
// helpers: staging of caller buffers, adam, all-reduce
// ---------------------------------------------------------------------------------------------
static bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
static const float* stage_in(gan_ctx* ctx, int idx, const float* p, size_t bytes) {
  if (idx < 2 && p == ctx->prefetch_src[idx] && bytes == ctx->prefetch_bytes) {
    // already on the device (gan_ctx_prefetch): wait for that copy, move it into the staging buffer
    // (device-to-device) and release the prefetch buffer for the next batch right away
    CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->prefetch_done, 0));
    ctx->stage[idx].ensure(bytes);
    CUDA_CHECK(cudaMemcpyAsync(ctx->stage[idx].p, ctx->prefetch_buf[idx].p, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_CHECK(cudaEventRecord(ctx->prefetch_consumed, ctx->stream));
    ctx->prefetch_src[idx] = nullptr;
    return ctx->stage[idx].as<float>();
  }
  if (is_device_ptr(p)) return p;
  ctx->stage[idx].ensure(bytes);
  CUDA_CHECK(cudaMemcpyAsync(ctx->stage[idx].p, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return ctx->stage[idx].as<float>();
}
static void copy_out(gan_ctx* ctx, float* dst, const float* src_dev, size_t bytes) {
  if (is_device_ptr(dst)) CUDA_CHECK(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  else {
    CUDA_CHECK(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
}
static void build_adam_tables(gan_net* n);
// Kernel gradients are stored (not accumulated) by the first weight-gradient launch of a step, so only the small
// gamma / beta / bias ranges, which the norm and head kernels accumulate into, are zeroed (one launch per net instead of
// a 229 MB memset per step).
static void zero_grads(gan_net* n) {
  if (n->adam_nent == 0) build_adam_tables(n);
  launch_zero_ranges(n->ctx->L(), n->grads.as<float>(), (const AdamRange*)n->adam_ranges.p, n->adam_nranges);
}
static void build_adam_tables(gan_net* n) {
  pack_weights(n);                                   // makes sure the packed buffers exist
  std::vector<AdamPackEntry> tab;
  std::vector<AdamRange> ranges;
  int tiles = 0;
  for (auto& ly : n->layers) {
    AdamPackEntry e; memset(&e, 0, sizeof(e));
    e.w_off = ly.w_off;
    e.conv2d = ly.kind != K_CONVT_S2;
    e.A = e.conv2d ? ly.Cin : ly.Cout; e.B = e.conv2d ? ly.Cout : ly.Cin;
    e.dstF = ly.wp_fwd.p; e.dstD = ly.wp_dgrad.p;
    ClassGeom cf[4], cd[4];
    int nf = fill_geometry(ly.kind, R_FWD, cf), nd = fill_geometry(ly.kind, R_DGRAD, cd);
    e.KcF = ly.Cin_p; e.KtotF = cf[0].ntaps * ly.Cin_p; e.KcD = ly.Cout_p; e.KtotD = cd[0].ntaps * ly.Cout_p;
    for (int c = 0; c < nf; ++c) {
      e.boffF[c] = (long long)c * ly.Cout_p * e.KtotF;
      for (int t = 0; t < cf[c].ntaps; ++t) e.invF[cf[c].widx[t]] = (int8_t)((c << 4) | t);
    }
    for (int c = 0; c < nd; ++c) {
      e.boffD[c] = (long long)c * ly.Cin_p * e.KtotD;
      for (int t = 0; t < cd[c].ntaps; ++t) e.invD[cd[c].widx[t]] = (int8_t)((c << 4) | t);
    }
    e.tiles_a = (e.A + 63) / 64; e.tiles_b = (e.B + 63) / 64;
    // 8-byte packed stores need 16-bit destinations with even strides; fp32 mode keeps the scalar path
    e.vec = (n->ctx->dt != DT_F32 && e.A % 64 == 0 && e.B % 64 == 0 && ly.w_off % 4 == 0 && e.KtotF % 4 == 0 && e.KtotD % 4 == 0) ? 1 : 0;
    e.tile_begin = tiles;
    tiles += 16 * e.tiles_a * e.tiles_b;
    tab.push_back(e);
    if (ly.g_off >= 0) { ranges.push_back(AdamRange{ly.g_off, ly.Cout, 0}); ranges.push_back(AdamRange{ly.b_off, ly.Cout, 0}); }
    if (ly.bias_off >= 0) ranges.push_back(AdamRange{ly.bias_off, ly.Cout, 0});
  }
  n->adam_tab.ensure(tab.size() * sizeof(AdamPackEntry));
  CUDA_CHECK(cudaMemcpy(n->adam_tab.p, tab.data(), tab.size() * sizeof(AdamPackEntry), cudaMemcpyHostToDevice));
  n->adam_ranges.ensure(ranges.size() * sizeof(AdamRange));
  CUDA_CHECK(cudaMemcpy(n->adam_ranges.p, ranges.data(), ranges.size() * sizeof(AdamRange), cudaMemcpyHostToDevice));
  n->adam_nent = (int)tab.size(); n->adam_tiles = tiles; n->adam_nranges = (int)ranges.size();
}

// `reduced`: the gradient buckets of this net have already been handed to the communication stream (comm_reduce_bucket)
// and joined (comm_join).
static void adam_apply(gan_adam* o, bool reduced = false) {
  gan_net* n = o->net; gan_ctx* ctx = n->ctx;
  if (n->adam_nent == 0) build_adam_tables(n);
  if (ctx->world > 1 && !reduced) { comm_step_begin(n); comm_reduce_bucket(n, 0); comm_join(ctx); }
  o->t += 1;
  launch_bump(ctx->L(), o->t_dev.as<long long>(), nullptr, 1);
  const float gscale = 1.f / ((float)ctx->world * n->grad_scale);     // undo the data-parallel sum and the loss scale
  if (ctx->world > 1 && ctx->shard_optimizer) {
    // ZeRO-1 style: every bucket was reduce-scattered, so this rank holds the summed gradient of its
    // sub-range only; it updates exactly that 1/world of the parameters (and of m, v), the updated master
    // parameters are all-gathered in place, and the 16-bit weight packs are rebuilt from the gathered master.
    ProfScope ps(ctx, FAM_ADAM, 28.0 * (double)n->nparams / ctx->world + 4.0 * (double)n->nparams * (ctx->world - 1) / ctx->world);
    const int nb = (int)n->bucket_off.size();
    for (int i = 0; i < nb; ++i) {
      const int64_t cnt = n->bucket_len[i] / ctx->world, off = n->bucket_off[i] + (int64_t)ctx->rank * cnt;
      launch_adam(ctx->L(), n->params.as<float>() + off, n->grads.as<float>() + off, o->m.as<float>() + off, o->v.as<float>() + off,
                  cnt, o->t_dev.as<long long>(), o->lr, o->b1, o->b2, (float)o->eps, gscale);
    }
    comm_allgather_buckets(ctx, n->params.as<float>(), n->bucket_off.data(), n->bucket_len.data(), nb);
    n->packed_dirty = true;
    pack_weights(n);
    return;
  }
  AdamArgs a{n->params.as<float>(), n->grads.as<float>(), o->m.as<float>(), o->v.as<float>(), o->t_dev.as<long long>(),
             o->lr, o->b1, o->b2, (float)o->eps, gscale, nullptr};
  if (ctx->world > 1 && ctx->comm16 && n->grads16_valid) a.g16 = n->grads16.as<uint16_t>();
  // fused update + repack: 28 B/param of optimizer traffic + 4 B/param for the two packed 16-bit copies
  ProfScope ps(ctx, FAM_ADAM, 32.0 * (double)n->nparams);
  launch_adam_pack(ctx->L(), ctx->dtA, ctx->dtG, a, (const AdamPackEntry*)n->adam_tab.p, n->adam_nent, n->adam_tiles);
  launch_adam_ranges(ctx->L(), a, (const AdamRange*)n->adam_ranges.p, n->adam_nranges);
  // the few small special layouts (first-layer im2col order, head cols operands)
  for (auto& gt : n->gathers) launch_gather_pack(ctx->L(), gt.dt, n->params.as<float>(), gt.idx.as<int>(), gt.n, gt.dst);
  n->packed_dirty = false;
}
static void finish_losses(gan_ctx* ctx, const LossMix& mix, float* losses_host) {
  ctx->loss_out.ensure(16 * 4);
  launch_loss_finalize(ctx->L(), ctx->loss_ws.as<float>(), mix, ctx->loss_out.as<float>());
  if (ctx->world > 1) {
    comm_allreduce_sum(ctx, ctx->loss_out.as<float>(), mix.nout);
    launch_scale(ctx->L(), ctx->loss_out.as<float>(), mix.nout, 1.f / (float)ctx->world);
  }
  ctx->n_losses = mix.nout;
  CUDA_CHECK(cudaMemcpyAsync(ctx->loss_host, ctx->loss_out.p, mix.nout * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (losses_host) {
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    memcpy(losses_host, ctx->loss_host, mix.nout * 4);
  }
}
// Advance the device-resident dropout call counter by the generator forwards run since the last bump.
static void bump_calls(gan_ctx* ctx) {
  if (ctx->gen_calls_pending == 0) return;
  launch_bump(ctx->L(), nullptr, ctx->call_dev.as<uint32_t>(), ctx->gen_calls_pending);
  ctx->call_counter += ctx->gen_calls_pending;
  ctx->gen_calls_pending = 0;
}
static void loss_ws_reset(gan_ctx* ctx) {
  ctx->loss_ws.ensure((size_t)LOSS_SLOTS * LOSS_BLOCKS * 4);
  CUDA_CHECK(cudaMemsetAsync(ctx->loss_ws.p, 0, (size_t)LOSS_SLOTS * LOSS_BLOCKS * 4, ctx->stream));
}

// ---------------------------------------------------------------------------------------------
// Pix2Pix.train_step (pix2pix.py:190-218)
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// Two-stream execution of a step.  The discriminator's pass over the REAL pair (forward, BCE, backward with its weight
// gradients) depends on nothing the generator computes, and the generator's forward spends a third of its launches in
// the 1x1..8x8 bottleneck, where a launch occupies 64..256 of the 296 CTA slots.  The side stream runs that
// discriminator pass meanwhile; the block scheduler fills the idle SMs.  Fork and join are events, so a captured step
// graph gets two parallel branches.  The first-layer rows of the real images are built before the fork (shared), every
// other buffer is private to a (net, call) slot or to the stream's Scratch set.
// ---------------------------------------------------------------------------------------------
static bool side_begin(gan_ctx* ctx) {
  if (!ctx->overlap || ctx->profile) return false;
  if (ctx->side == nullptr) {
    CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_mid, cudaEventDisableTiming));
  }
  CUDA_CHECK(cudaEventRecord(ctx->ev_fork, ctx->stream));
  CUDA_CHECK(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
  ctx->cur = 1;
  return true;
}
static void side_end(gan_ctx* ctx) { ctx->cur = 0; }                       // back to the main stream; the side keeps running
static void side_join(gan_ctx* ctx) {
  CUDA_CHECK(cudaEventRecord(ctx->ev_join, ctx->side));
  CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
}
// intermediate dependency: everything enqueued on the side stream so far (recorded while cur == 1) ...
static void side_mark_mid(gan_ctx* ctx) { CUDA_CHECK(cudaEventRecord(ctx->ev_mid, ctx->side)); }
// ... must have finished before what the main stream enqueues next
static void main_wait_mid(gan_ctx* ctx) { CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_mid, 0)); }
// first-layer rows of an input that several nets / streams read: built on the main stream before a fork
static void share_im2col(gan_ctx* ctx, gan_net* n, const float* img, int B, int H, int W) {
  if (!im2col_on(ctx, n->layers[0])) return;
  // the first-layer kernels rebuild the rows in shared memory (forward and weight gradient): nothing to share
  const Layer& ly = n->layers[0];
  Slot probe; probe.src_f32[0] = img; probe.src_f32[1] = ly.nsrc == 2 ? img : nullptr;
  if (ctx->dtG == ctx->dtA && first_wgrad_enabled() && first_kernel_ok(ctx, ly, probe, H, W)) return;
  cached_im2col(ctx, img, B, H, W, n->C);
}

// Static loss scale for fp16 gradient storage: the largest gradient any loss head emits (`g_head`: weight / number of
// elements of its mean) is brought to ~4 by a power of two.  Measured on this model (scripts/): |dL/dz| over
// all layers spans 2e-9 .. 16x the head value, i.e. 3e-5 .. 64 after scaling — inside fp16's 6e-8 .. 65504 with three
// orders of magnitude of head-room on top, and conversions saturate.  fp32 / bf16 storage: 1.
static float pick_grad_scale(const gan_ctx* ctx, double g_head) {
  if (ctx->dtG != DT_F16 || !(g_head > 0.0)) return 1.f;
  int k = (int)floor(log2(4.0 / g_head));
  if (k < 0) k = 0;
  if (k > 24) k = 24;
  return ldexpf(1.f, k);
}

// gan_scale: factor on the adversarial term of the GENERATOR gradient only (1 for the default 'l1' loss; the
// reference's 'ssim' branch makes the total loss a per-image vector, whose tape.gradient is the gradient of the SUM
// over the batch = B x d(gan_loss), pix2pix.py:182-186,210).
static void pix2pix_step(gan_net* g, gan_net* d, gan_adam* go, gan_adam* dopt, const float* x_in, const float* y_in, int B,
                         float lambda, float gan_scale, int training, float* losses) {
  gan_ctx* ctx = g->ctx;
  GAN_REQUIRE(d->ctx == ctx, "nets belong to different contexts");
  GAN_REQUIRE(g->is_gen && !d->is_gen && d->target, "pix2pix needs a generator and a target=True discriminator");
  GAN_REQUIRE(B >= 1, "batch must be >= 1");
  const int H = g->H, W = g->W, C = g->C;
  const size_t img_bytes = (size_t)B * H * W * C * 4;
  const float* x = stage_in(ctx, 0, x_in, img_bytes);
  const float* y = stage_in(ctx, 1, y_in, img_bytes);
  ctx->prefetch_src[0] = ctx->prefetch_src[1] = nullptr;   // a prefetch this step did not consume must not match later
  ctx->cur = 0;
  loss_ws_reset(ctx);
  ctx->step_epoch++; ctx->im2col_next = 0;
  if (training) { zero_grads(g); zero_grads(d); comm_step_begin(g); comm_step_begin(d); }
  const double n_img_d = (double)B * H * W * C, n_log_d = (double)B * (H / 8 - 2) * (W / 8 - 2);
  const float S = pick_grad_scale(ctx, std::max((double)lambda / n_img_d, std::max((double)gan_scale, 0.5) / n_log_d));
  ctx->grad_scale = S; g->grad_scale = S; d->grad_scale = S;

  pack_weights(g); pack_weights(d);
  share_im2col(ctx, g, x, B, H, W); share_im2col(ctx, d, y, B, H, W);
  const int64_t n_img = (int64_t)B * H * W * C;
  // raw loss slots: 0 BCE(1,fake)  1 L1  2 BCE(1,real)  3 BCE(0,fake)
  const bool forked = side_begin(ctx);
  discriminator_forward(d, 0, x, y, B, H, W);                            // disc_real_output     (:202)
  disc_bce(d, 0, 1.f, 0.5f * S, training, true, 2);
  if (training) discriminator_backward(d, 0, true, false);               // first contribution to D's gradients
  side_end(ctx);
  generator_forward(g, 0, x, B, H, W);                                   // gen_output           (:200)
  const float* gen_out = g->slots[0].out_f32.as<float>();
  if (forked) side_join(ctx);
  discriminator_forward(d, 1, x, gen_out, B, H, W);                      // disc_generated_output(:203)
  launch_l1(ctx->L(), y, gen_out, n_img, ctx->loss_ws.as<float>(), 1);   // gan_loss2 = mean|target-gen_output| (:181)
  disc_bce(d, 1, 0.f, 0.5f * S, training, true, 3);
  if (training) {
    discriminator_backward(d, 1, true, false);
    comm_reduce_bucket(d, 0);                                             // D gradients final: reduce under G's backward
  }
  disc_bce(d, 1, 1.f, gan_scale * S, training, false, 0);
  if (training) {
    discriminator_backward(d, 1, false, true);                           // dL_G/d(gen_output) through D (:210)
    GradSrc dgan{d->slots[1].din0.p, d->slots[1].din0_pitch, d->slots[1].din0_coff};
    generator_backward(g, 0, dgan, GradSrc{nullptr, 0, 0}, y, S * lambda / (float)n_img, false, true);
    comm_join(ctx);
    adam_apply(go, true);                                                // (:213)
    adam_apply(dopt, true);                                              // (:215)
  }
  LossMix mix; memset(&mix, 0, sizeof(mix));
  mix.nraw = 4; mix.nout = 4;
  const float nlog = (float)logits_count(d->slots[0]);
  mix.denom[0] = nlog; mix.denom[1] = (float)n_img; mix.denom[2] = nlog; mix.denom[3] = nlog;
  auto M = [&](int i, int j) -> float& { return mix.mix[i * mix.nraw + j]; };
  M(0, 0) = 1.f; M(0, 1) = lambda;        // gen_total_loss = gan_loss + lambda*l1   (:186)
  M(1, 0) = 1.f;                          // gen_gan_loss
  M(2, 1) = 1.f;                          // gen_gan_loss2 (L1)
  M(3, 2) = 0.5f; M(3, 3) = 0.5f;         // disc_loss = (real+generated)*0.5         (:206)
  bump_calls(ctx);
  finish_losses(ctx, mix, losses);
}

// ---------------------------------------------------------------------------------------------
// CycleGAN.train_step (cycle_gan.py:206-276), single backward sweep
// ---------------------------------------------------------------------------------------------
static void cyclegan_step(gan_net* g, gan_net* f, gan_net* dx, gan_net* dy, gan_adam* og, gan_adam* of, gan_adam* odx,
                          gan_adam* ody, const float* x_in, const float* y_in, int B, float lambda, int training,
                          float* losses) {
  gan_ctx* ctx = g->ctx;
  GAN_REQUIRE(g->is_gen && f->is_gen && !dx->is_gen && !dy->is_gen && !dx->target && !dy->target,
              "cyclegan needs two generators and two target=False discriminators");
  GAN_REQUIRE(B >= 1, "batch must be >= 1");
  const int C = g->C;
  int H = g->H, W = g->W;
  const size_t img_bytes = (size_t)B * H * W * C * 4;
  const float* x = stage_in(ctx, 0, x_in, img_bytes);
  const float* y = stage_in(ctx, 1, y_in, img_bytes);
  ctx->prefetch_src[0] = ctx->prefetch_src[1] = nullptr;   // a prefetch this step did not consume must not match later
  ctx->cur = 0;
  loss_ws_reset(ctx);
  ctx->step_epoch++; ctx->im2col_next = 0;
  if (training) {
    zero_grads(g); zero_grads(f); zero_grads(dx); zero_grads(dy);
    comm_step_begin(g); comm_step_begin(f); comm_step_begin(dx); comm_step_begin(dy);
  }
  const float S = pick_grad_scale(ctx, std::max((double)lambda / ((double)B * H * W * C), 1.0 / ((double)B * (H / 8 - 2) * (W / 8 - 2))));
  ctx->grad_scale = S; g->grad_scale = S; f->grad_scale = S; dx->grad_scale = S; dy->grad_scale = S;

  pack_weights(g); pack_weights(f); pack_weights(dx); pack_weights(dy);
  share_im2col(ctx, g, x, B, H, W); share_im2col(ctx, f, y, B, H, W);
  const int64_t n_img = (int64_t)B * H * W * C;
  float* lw = ctx->loss_ws.as<float>();
  const GradSrc none{nullptr, 0, 0};
  const float lc = S * lambda / (float)n_img;
  // raw: 0 BCE(1,fake_y) 1 BCE(1,fake_x) 2 L1(x,cyc_x) 3 L1(y,cyc_y) 4 L1(y,same_y) 5 L1(x,same_x)
  //      6 BCE(1,real_x) 7 BCE(0,fake_x) 8 BCE(1,real_y) 9 BCE(0,fake_y)
  // Side stream: everything that depends on the REAL images only — both discriminators' real passes and the whole
  // identity branch (same_x = F(x), same_y = G(y), their L1 losses and backward sweeps), which therefore become the
  // FIRST contributions to the four gradient buffers.  The generator calls keep the reference's position in the
  // dropout-mask sequence (cycle_gan.py:220-228: fake_y 0, cycled_x 1, fake_x 2, cycled_y 3, same_x 4, same_y 5).
  const bool forked = side_begin(ctx);
  discriminator_forward(dx, 0, x, nullptr, B, H, W);        // disc_real_x (:230)
  disc_bce(dx, 0, 1.f, 0.5f * S, training, true, 6); if (training) discriminator_backward(dx, 0, true, false);
  discriminator_forward(dy, 0, y, nullptr, B, H, W);        // disc_real_y (:231)
  disc_bce(dy, 0, 1.f, 0.5f * S, training, true, 8); if (training) discriminator_backward(dy, 0, true, false);
  if (forked) side_mark_mid(ctx);
  generator_forward(f, 2, x, B, H, W, 4);  const float* same_x = f->slots[2].out_f32.as<float>();     // (:227)
  generator_forward(g, 2, y, B, H, W, 5);  const float* same_y = g->slots[2].out_f32.as<float>();     // (:228)
  launch_l1(ctx->L(), y, same_y, n_img, lw, 4);
  launch_l1(ctx->L(), x, same_x, n_img, lw, 5);
  if (training) {
    generator_backward(f, 2, none, none, x, 0.5f * lc, false);                           // identity x (:244)
    generator_backward(g, 2, none, none, y, 0.5f * lc, false);                           // identity y (:243)
  }
  side_end(ctx);
  generator_forward(g, 0, x, B, H, W, 0);  const float* fake_y = g->slots[0].out_f32.as<float>();     // (:220)
  generator_forward(f, 0, fake_y, B, H, W, 1); const float* cycled_x = f->slots[0].out_f32.as<float>(); // (:221)
  generator_forward(f, 1, y, B, H, W, 2);  const float* fake_x = f->slots[1].out_f32.as<float>();     // (:223)
  generator_forward(g, 1, fake_x, B, H, W, 3); const float* cycled_y = g->slots[1].out_f32.as<float>(); // (:224)
  launch_l1(ctx->L(), x, cycled_x, n_img, lw, 2);
  launch_l1(ctx->L(), y, cycled_y, n_img, lw, 3);
  if (forked) main_wait_mid(ctx);                           // the real passes precede the fake ones (gradient order, BN statistics)
  discriminator_forward(dx, 1, fake_x, nullptr, B, H, W);   // disc_fake_x (:233)
  discriminator_forward(dy, 1, fake_y, nullptr, B, H, W);   // disc_fake_y (:234)
  disc_bce(dx, 1, 0.f, 0.5f * S, training, true, 7); if (training) { discriminator_backward(dx, 1, true, false); comm_reduce_bucket(dx, 0); }
  disc_bce(dy, 1, 0.f, 0.5f * S, training, true, 9); if (training) { discriminator_backward(dy, 1, true, false); comm_reduce_bucket(dy, 0); }
  disc_bce(dy, 1, 1.f, S, training, false, 0); if (training) discriminator_backward(dy, 1, false, true);
  disc_bce(dx, 1, 1.f, S, training, false, 1); if (training) discriminator_backward(dx, 1, false, true);
  if (forked) side_join(ctx);
  if (training) {
    generator_backward(f, 0, none, none, x, lc, true);                                   // cycle x: through F into fake_y
    generator_backward(g, 1, none, none, y, lc, true);                                   // cycle y: through G into fake_x
    // the last contribution to each generator's gradients: their buckets go to the communication stream as they close
    generator_backward(g, 0, GradSrc{dy->slots[1].din0.p, dy->slots[1].din0_pitch, dy->slots[1].din0_coff}, GradSrc{f->slots[0].dxin.p, f->Cp, 0}, nullptr, 0.f, false, true);
    generator_backward(f, 1, GradSrc{dx->slots[1].din0.p, dx->slots[1].din0_pitch, dx->slots[1].din0_coff}, GradSrc{g->slots[1].dxin.p, g->Cp, 0}, nullptr, 0.f, false, true);
    comm_join(ctx);
    adam_apply(og, true); adam_apply(of, true); adam_apply(odx, true); adam_apply(ody, true);     // (:263-273)
  }
  LossMix mix; memset(&mix, 0, sizeof(mix));
  mix.nraw = 10; mix.nout = 7;
  const float nlog = (float)logits_count(dx->slots[0]);
  for (int j = 0; j < 10; ++j) mix.denom[j] = (j >= 2 && j <= 5) ? (float)n_img : nlog;
  auto M = [&](int i, int j) -> float& { return mix.mix[i * mix.nraw + j]; };
  M(0, 0) = 1.f;                                                       // gen_g_loss
  M(1, 1) = 1.f;                                                       // gen_f_loss
  M(2, 2) = lambda; M(2, 3) = lambda;                                  // total_cycle_loss (:240)
  M(3, 0) = 1.f; M(3, 2) = lambda; M(3, 3) = lambda; M(3, 4) = 0.5f * lambda;   // total_gen_g_loss (:243)
  M(4, 1) = 1.f; M(4, 2) = lambda; M(4, 3) = lambda; M(4, 5) = 0.5f * lambda;   // total_gen_f_loss (:244)
  M(5, 6) = 0.5f; M(5, 7) = 0.5f;                                      // disc_x_loss (:246)
  M(6, 8) = 0.5f; M(6, 9) = 0.5f;                                      // disc_y_loss (:247)
  bump_calls(ctx);
  finish_losses(ctx, mix, losses);
}

// ---------------------------------------------------------------------------------------------
// CUDA-graph execution of a whole train step.  First call with a key: eager (allocates every
// buffer); second call: stream capture + instantiate; later calls: one cudaGraphLaunch.  Inputs
// are always copied into the ctx staging buffers first (their addresses are baked into the graph);
// the Adam step counters and the dropout call counter live in device memory.
// ---------------------------------------------------------------------------------------------
template <typename F>
static void run_step_graphed(gan_ctx* ctx, const std::string& key, const float* x_in, const float* y_in, size_t img_bytes,
                             float* losses_host, int nloss, const std::vector<gan_net*>& nets,
                             const std::vector<gan_adam*>& opts, F&& body) {
  for (gan_net* n : nets) pack_weights(n);      // host-side set_tensor since the last step: repack outside the graph
  ctx->stage[0].ensure(img_bytes); ctx->stage[1].ensure(img_bytes);
  const float* srcs[2] = {x_in, y_in};
  for (int i = 0; i < 2; ++i) {
    const void* src = srcs[i];
    bool pre = false;
    if (src == ctx->prefetch_src[i] && img_bytes == ctx->prefetch_bytes) {   // prefetched: device-to-device
      CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->prefetch_done, 0));
      src = ctx->prefetch_buf[i].p;
      ctx->prefetch_src[i] = nullptr;
      pre = true;
    }
    CUDA_CHECK(cudaMemcpyAsync(ctx->stage[i].p, src, img_bytes, cudaMemcpyDefault, ctx->stream));
    if (pre) CUDA_CHECK(cudaEventRecord(ctx->prefetch_consumed, ctx->stream));   // prefetch buffer free again
  }
  ctx->prefetch_src[0] = ctx->prefetch_src[1] = nullptr;   // a prefetch this step did not consume must not match later
  const float* xs = ctx->stage[0].as<float>(); const float* ys = ctx->stage[1].as<float>();
  gan_ctx::GraphEntry& ge = ctx->graph_cache[key];
  if (ge.exec != nullptr && ge.alloc_epoch != g_alloc_epoch) {
    // a buffer was re-allocated since the capture (a larger batch came by): the graph's pointers may be stale
    cudaGraphExecDestroy(ge.exec);
    ge.exec = nullptr; ge.warm = 0;
  }
  if (ge.exec != nullptr) {
    CUDA_CHECK(cudaGraphLaunch(ge.exec, ctx->stream));
    ctx->launches += ge.launches;
    ctx->call_counter += ge.gen_calls;
    for (gan_adam* o : opts) o->t += 1;
    ctx->n_losses = nloss;
  } else if (ge.warm == 0) {
    body(xs, ys);
    ge.warm = 1;
  } else {
    const uint64_t l0 = ctx->launches;
    const uint32_t c0 = ctx->call_counter;
    cudaGraph_t graph = nullptr;
    CUDA_CHECK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    try { body(xs, ys); }
    catch (...) { cudaStreamEndCapture(ctx->stream, &graph); if (graph) cudaGraphDestroy(graph); throw; }
    CUDA_CHECK(cudaStreamEndCapture(ctx->stream, &graph));
    ge.launches = ctx->launches - l0;
    ge.gen_calls = ctx->call_counter - c0;
    CUDA_CHECK(cudaGraphInstantiate(&ge.exec, graph, 0));
    ge.alloc_epoch = g_alloc_epoch;
    cudaGraphDestroy(graph);
    CUDA_CHECK(cudaGraphLaunch(ge.exec, ctx->stream));   // the capture itself executed nothing
  }
  if (losses_host) {
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    memcpy(losses_host, ctx->loss_host, (size_t)nloss * 4);
  }
}

// =============================================================================================
// C-ABI
// =============================================================================================
extern "C" {

const char* gan_last_error(void) { return g_last_error.c_str(); }
int gan_version(void) { return 100; }

// Ownership: a context owns the nets and optimizers created from it.  Destroying a net also destroys the
// optimizers bound to it; destroying the context destroys everything.  Handles are tracked so that a
// second destroy of the same handle is rejected instead of freeing twice.
static std::mutex g_live_mu;
static std::set<const void*> g_live;
static void live_add(const void* h) { std::lock_guard<std::mutex> l(g_live_mu); g_live.insert(h); }
static bool live_take(const void* h) { std::lock_guard<std::mutex> l(g_live_mu); return g_live.erase(h) == 1; }
static void destroy_adam(gan_adam* o) {
  if (!live_take(o)) return;
  { auto& v = o->net->ctx->adams; v.erase(std::remove(v.begin(), v.end(), o), v.end()); }
  delete o;
}
static void destroy_net(gan_net* n) {
  if (!live_take(n)) return;
  gan_ctx* ctx = n->ctx;
  std::vector<gan_adam*> bound;
  for (gan_adam* o : ctx->adams) if (o->net == n) bound.push_back(o);
  for (gan_adam* o : bound) destroy_adam(o);
  ctx->nets.erase(std::remove(ctx->nets.begin(), ctx->nets.end(), n), ctx->nets.end());
  delete n;
}

int gan_ctx_create(int device, int precision, uint64_t seed, gan_ctx** out) {
  API_BEGIN
  GAN_REQUIRE(out != nullptr, "null out");
  GAN_REQUIRE(precision == GAN_FP32 || precision == GAN_BF16, "bad precision");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    throw GanError(GAN_ERR_NO_DEVICE, "no CUDA device: libgan_b200 has no CPU fallback");
  }
  GAN_REQUIRE(device >= 0 && device < ndev, "bad device index");
  CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) throw GanError(GAN_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major) +
                                                              std::to_string(prop.minor) + ", this library is sm_100a only");
  gan_ctx* c = new gan_ctx();
  c->device = device; c->dt = precision == GAN_FP32 ? DT_F32 : DT_BF16; c->seed = seed;
  { const char* e = getenv("GAN_B200_OVERLAP"); c->overlap = !(e && e[0] == '0'); }      // dev A/B switch: two-stream steps
  if (c->dt == DT_F32) { c->dtA = DT_F32; c->dtG = DT_F32; }
  else {
    // 16-bit mode: fp16 storage with a static loss scale (common.cuh).  GAN_B200_ACT=bf16 selects all-bf16 storage
    // (round 1's numerics: 8x the rounding error, no loss scale needed; kept for A/B measurements).
    const char* e = getenv("GAN_B200_ACT");
    c->dtA = (e && strcmp(e, "bf16") == 0) ? DT_BF16 : DT_F16;
    c->dtG = c->dtA;
  }
  CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaMallocHost((void**)&c->loss_host, 16 * 4));
  c->call_dev.ensure(16);
  umma_init();
  first_init();
  live_add(c);
  *out = c;
  API_END
}
// Captured step graphs are keyed by the addresses of the nets/optimizers they were captured for and hold
// their device pointers: drop them whenever one of those objects (or the seed) goes away.
static void drop_graphs(gan_ctx* ctx) {
  for (auto& kv : ctx->graph_cache) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  ctx->graph_cache.clear();
}

int gan_ctx_destroy(gan_ctx* ctx) {
  API_BEGIN
  if (!ctx) return GAN_OK;
  GAN_REQUIRE(live_take(ctx), "unknown or already destroyed context");
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  drop_graphs(ctx);
  while (!ctx->adams.empty()) destroy_adam(ctx->adams.back());
  while (!ctx->nets.empty()) destroy_net(ctx->nets.back());
  comm_destroy(ctx);
  cudaFreeHost(ctx->loss_host);
  if (ctx->copy_stream) { cudaStreamDestroy(ctx->copy_stream); cudaEventDestroy(ctx->prefetch_done); cudaEventDestroy(ctx->prefetch_consumed); }
  if (ctx->side) { cudaStreamDestroy(ctx->side); cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_join); cudaEventDestroy(ctx->ev_mid); }
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  API_END
}
int gan_ctx_sync(gan_ctx* ctx) {
  API_BEGIN
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  CUDA_CHECK(cudaGetLastError());
  API_END
}
// State that a captured step graph bakes in (kernel template choice, kernel arguments): changing it
// synchronises the stream and drops the captured graphs, so the next step re-captures with the new value.
static void invalidate_graphs(gan_ctx* ctx) {
  if (ctx->graph_cache.empty()) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  drop_graphs(ctx);
}
int gan_ctx_set_dropout(gan_ctx* ctx, int enabled) {
  API_BEGIN
  const int v = enabled ? 1 : 0;
  if (v != ctx->dropout_enabled) invalidate_graphs(ctx);
  ctx->dropout_enabled = v;
  API_END
}
int gan_ctx_set_rng(gan_ctx* ctx, uint64_t seed, uint32_t call_counter) {
  API_BEGIN
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->seed = seed; ctx->call_counter = call_counter; ctx->gen_calls_pending = 0;
  CUDA_CHECK(cudaMemcpy(ctx->call_dev.p, &call_counter, 4, cudaMemcpyHostToDevice));
  drop_graphs(ctx);              // the seed is baked into captured kernels
  API_END
}
int gan_ctx_get_call_counter(gan_ctx* ctx, uint32_t* out) { *out = ctx->call_counter; return GAN_OK; }
int gan_ctx_set_engine(gan_ctx* ctx, int engine) {
  API_BEGIN
  if (engine != ctx->engine) invalidate_graphs(ctx);
  ctx->engine = engine;
  API_END
}
int gan_ctx_set_graphs(gan_ctx* ctx, int enabled) { ctx->graphs = enabled; return GAN_OK; }
int gan_ctx_launch_count(gan_ctx* ctx, uint64_t* out) { *out = ctx->launches; return GAN_OK; }
int gan_ctx_stream(gan_ctx* ctx, void** out) { *out = (void*)ctx->stream; return GAN_OK; }
int gan_ctx_set_profile(gan_ctx* ctx, int enabled) {
  API_BEGIN
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  for (auto& e : ctx->prof) { ctx->ev_pool.push_back(e.a); ctx->ev_pool.push_back(e.b); }
  ctx->prof.clear();
  ctx->profile = enabled;
  API_END
}
int gan_ctx_profile_read(gan_ctx* ctx, double ms[8], double work[8], int64_t count[8]) {
  API_BEGIN
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < FAM_COUNT; ++i) { ms[i] = 0; work[i] = 0; count[i] = 0; }
  for (auto& e : ctx->prof) {
    float t = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&t, e.a, e.b));
    ms[e.fam] += t; work[e.fam] += e.work; count[e.fam] += 1;
    ctx->ev_pool.push_back(e.a); ctx->ev_pool.push_back(e.b);
  }
  ctx->prof.clear();
  API_END
}
int gan_ctx_set_sample_offset(gan_ctx* ctx, int64_t sample0) {
  API_BEGIN
  if (!ctx->sample0_set || sample0 != ctx->sample0) invalidate_graphs(ctx);    // sample0 is a kernel argument
  ctx->sample0 = sample0; ctx->sample0_set = true;
  API_END
}

int gan_comm_unique_id(void* out128) {
  API_BEGIN
  int r = comm_unique_id(out128);
  if (r != 0) throw GanError(GAN_ERR_COMM, "ncclGetUniqueId failed");
  API_END
}
int gan_ctx_comm_init(gan_ctx* ctx, int rank, int world, const void* unique_id128) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(ctx->device));
  // Sharded optimizer (reduce-scatter -> Adam on 1/world of the parameters -> all-gather of the fp32 master): measured
  // SLOWER than all-reduce + the fused full Adam at every world size on 8 x B200 (ms/step at global batch 64:
  // N=2 5.97 vs 5.62, N=4 4.06 vs 3.66, N=8 3.32 vs 3.11): the all-gather of 229 MB sits on the critical path where
  // the all-reduce hides under the backward sweep.  Off unless GAN_B200_SHARD_OPT=1.
  { const char* e = getenv("GAN_B200_SHARD_OPT"); ctx->shard_optimizer = e ? (e[0] != '0') : false; }
  // 16-bit storage mode: gradient buckets are all-reduced as bf16 (GAN_B200_COMM16=0: fp32 on the wire); the fp32
  // parity mode and the sharded optimizer keep fp32
  { const char* e = getenv("GAN_B200_COMM16"); ctx->comm16 = (ctx->dt == DT_BF16 && !ctx->shard_optimizer && world > 1) ? (e ? (e[0] != '0') : 1) : 0; }
  comm_init(ctx, rank, world, unique_id128);
  API_END
}

int gan_generator_create(gan_ctx* ctx, int norm_type, int height, int width, int channels, gan_net** out) {
  API_BEGIN
  GAN_REQUIRE(ctx && out, "null argument");
  GAN_REQUIRE(norm_type == GAN_NORM_BATCH || norm_type == GAN_NORM_INSTANCE, "bad norm type");
  GAN_REQUIRE(channels >= 1 && channels <= 4, "channels must be 1..4");
  GAN_REQUIRE(height >= 256 && width >= 256 && height % 256 == 0 && width % 256 == 0, "image size must be a multiple of 256");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  gan_net* n = new gan_net();
  std::unique_ptr<gan_net> guard(n);          // freed if construction throws
  n->ctx = ctx; n->is_gen = true; n->norm = norm_type; n->H = height; n->W = width; n->C = channels; n->Cin0 = channels;
  n->Cp = pad_c(ctx, channels); n->Cin0_p = n->Cp;
  int cin = channels;
  for (int j = 1; j <= 8; ++j) {            // downsample blocks, first without norm (base_gan.py:179-188)
    add_layer(n, "down" + std::to_string(j), K_CONV_S2, cin, DOWN_F[j - 1], j == 1 ? NORM_NONE : norm_type, ACT_LEAKY, false,
              false, 0, false);
    cin = DOWN_F[j - 1];
  }
  for (int k = 1; k <= 7; ++k) {            // upsample blocks, first three with dropout (base_gan.py:190-198)
    add_layer(n, "up" + std::to_string(k), K_CONVT_S2, cin, UP_F[k - 1], norm_type, ACT_RELU, false, k <= 3, k, false);
    cin = UP_F[k - 1] + DOWN_F[7 - k];
  }
  add_layer(n, "last", K_CONVT_S2, cin, channels, NORM_NONE, ACT_TANH, true, false, 0, true);   // base_gan.py:201-204
  finish_net(n);
  guard.release(); ctx->nets.push_back(n); live_add(n);
  *out = n;
  API_END
}

int gan_discriminator_create(gan_ctx* ctx, int norm_type, int channels, int target, gan_net** out) {
  API_BEGIN
  GAN_REQUIRE(ctx && out, "null argument");
  GAN_REQUIRE(norm_type == GAN_NORM_BATCH || norm_type == GAN_NORM_INSTANCE, "bad norm type");
  GAN_REQUIRE(channels >= 1 && channels <= 4, "channels must be 1..4");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  gan_net* n = new gan_net();
  std::unique_ptr<gan_net> guard(n);          // freed if construction throws
  n->ctx = ctx; n->is_gen = false; n->norm = norm_type; n->C = channels; n->target = target != 0;
  n->Cin0 = target ? 2 * channels : channels;
  n->Cp = pad_c(ctx, channels); n->Cin0_p = pad_c(ctx, n->Cin0);
  add_layer(n, "down1", K_CONV_S2, n->Cin0, 64, NORM_NONE, ACT_LEAKY, false, false, 0, false);   // base_gan.py:141
  add_layer(n, "down2", K_CONV_S2, 64, 128, norm_type, ACT_LEAKY, false, false, 0, false);       // :142
  add_layer(n, "down3", K_CONV_S2, 128, 256, norm_type, ACT_LEAKY, false, false, 0, false);      // :143
  add_layer(n, "conv512", K_CONV_S1P, 256, 512, norm_type, ACT_LEAKY, false, false, 0, false);   // :145-155
  n->tensors[n->tensors.size() - 2].name = "norm.gamma"; n->tensors[n->tensors.size() - 1].name = "norm.beta";
  add_layer(n, "last", K_CONV_S1P, 512, 1, NORM_NONE, ACT_NONE, true, false, 0, true);           // :157-161
  finish_net(n);
  guard.release(); ctx->nets.push_back(n); live_add(n);
  *out = n;
  API_END
}

int gan_net_destroy(gan_net* net) {
  API_BEGIN
  if (net) {
    { std::lock_guard<std::mutex> l(g_live_mu); GAN_REQUIRE(g_live.count(net) == 1, "unknown or already destroyed net"); }
    cudaSetDevice(net->ctx->device); cudaStreamSynchronize(net->ctx->stream); drop_graphs(net->ctx);
    destroy_net(net);
  }
  API_END
}

int gan_net_num_tensors(gan_net* net, int* trainable, int* total) {
  if (trainable) *trainable = net->ntrain;
  if (total) *total = (int)net->tensors.size();
  return GAN_OK;
}
int gan_net_tensor_info(gan_net* net, int idx, char* name, int name_cap, int* ndim, int64_t shape[4], int64_t* numel) {
  API_BEGIN
  GAN_REQUIRE(idx >= 0 && idx < (int)net->tensors.size(), "tensor index out of range");
  const TensorInfo& t = net->tensors[idx];
  if (name && name_cap > 0) { strncpy(name, t.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (ndim) *ndim = t.ndim;
  if (shape) for (int i = 0; i < 4; ++i) shape[i] = t.shape[i];
  if (numel) *numel = t.numel;
  API_END
}
static float* tensor_dev(gan_net* net, int idx, DevBuf* alt = nullptr) {
  GAN_REQUIRE(idx >= 0 && idx < (int)net->tensors.size(), "tensor index out of range");
  const TensorInfo& t = net->tensors[idx];
  if (t.trainable) return (alt ? alt->as<float>() : net->params.as<float>()) + t.off;
  GAN_REQUIRE(alt == nullptr, "non-trainable tensor has no gradient");
  return net->mov.as<float>() + t.off;
}
int gan_net_get_tensor(gan_net* net, int idx, float* host_dst) {
  API_BEGIN
  CUDA_CHECK(cudaStreamSynchronize(net->ctx->stream));
  CUDA_CHECK(cudaMemcpy(host_dst, tensor_dev(net, idx), net->tensors[idx].numel * 4, cudaMemcpyDeviceToHost));
  API_END
}
int gan_net_set_tensor(gan_net* net, int idx, const float* host_src) {
  API_BEGIN
  CUDA_CHECK(cudaStreamSynchronize(net->ctx->stream));
  CUDA_CHECK(cudaMemcpy(tensor_dev(net, idx), host_src, net->tensors[idx].numel * 4, cudaMemcpyHostToDevice));
  net->packed_dirty = true;
  API_END
}
// the gradient buffer holds grad_scale x the gradient (a power of two: the division is exact)
static void unscale_host(float* p, int64_t n, float s) {
  if (s == 1.f) return;
  const float inv = 1.f / s;
  for (int64_t i = 0; i < n; ++i) p[i] *= inv;
}
// gradient range -> host fp32.  Data parallel with bf16 communication: the all-reduced (summed) gradient lives in grads16.
static void grads_to_host(gan_net* net, int64_t off, int64_t numel, float* host_dst) {
  CUDA_CHECK(cudaStreamSynchronize(net->ctx->stream));
  if (net->ctx->world > 1 && net->ctx->comm16 && net->grads16_valid) {
    std::vector<uint16_t> tmp((size_t)numel);
    CUDA_CHECK(cudaMemcpy(tmp.data(), net->grads16.as<uint16_t>() + off, (size_t)numel * 2, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < numel; ++i) { uint32_t b = (uint32_t)tmp[(size_t)i] << 16; memcpy(host_dst + i, &b, 4); }
  } else {
    CUDA_CHECK(cudaMemcpy(host_dst, net->grads.as<float>() + off, (size_t)numel * 4, cudaMemcpyDeviceToHost));
  }
  unscale_host(host_dst, numel, net->grad_scale);
}
int gan_net_get_grad(gan_net* net, int idx, float* host_dst) {
  API_BEGIN
  grads_to_host(net, tensor_dev(net, idx, &net->grads) - net->grads.as<float>(), net->tensors[idx].numel, host_dst);
  API_END
}
int gan_net_num_params(gan_net* net, int64_t* out) { *out = net->nparams; return GAN_OK; }
int gan_net_get_params(gan_net* net, float* host_dst) {
  API_BEGIN
  CUDA_CHECK(cudaStreamSynchronize(net->ctx->stream));
  CUDA_CHECK(cudaMemcpy(host_dst, net->params.p, net->nparams * 4, cudaMemcpyDeviceToHost));
  API_END
}
int gan_net_set_params(gan_net* net, const float* host_src) {
  API_BEGIN
  CUDA_CHECK(cudaStreamSynchronize(net->ctx->stream));
  CUDA_CHECK(cudaMemcpy(net->params.p, host_src, net->nparams * 4, cudaMemcpyHostToDevice));
  net->packed_dirty = true;
  API_END
}
int gan_net_get_grads(gan_net* net, float* host_dst) {
  API_BEGIN
  grads_to_host(net, 0, net->nparams, host_dst);
  API_END
}

int gan_net_debug_tensor(gan_net* net, int slot, const char* name, float* host_dst, int64_t cap, int64_t* numel) {
  API_BEGIN
  gan_ctx* ctx = net->ctx;
  GAN_REQUIRE(slot >= 0 && slot < (int)net->slots.size(), "bad slot");
  Slot& s = net->slots[slot];
  GAN_REQUIRE(s.B > 0, "slot has no forward call yet");
  std::string nm(name);
  const void* src = nullptr; int pitch = 0, coff = 0, C = 0; int64_t P = 0; bool is_f32 = false;
  int src_dt = ctx->dtA;
  if (nm == "out" && net->is_gen) { src = s.out_f32.p; C = net->C; P = (int64_t)s.B * s.H * s.W; pitch = C; is_f32 = true; }
  else if (nm == "logits" && !net->is_gen) { src = s.logits.p; C = 1; P = logits_count(s); pitch = 1; is_f32 = true; }
  else if (nm == "din0" && !net->is_gen) {
    GAN_REQUIRE(s.din0_pitch == net->Cin0_p, "din0 is stored as compact rows of the generated-image channels (GAN_B200_DIN_COLS=0 for the full tensor)");
    src = s.din0.p; C = net->Cin0; P = (int64_t)s.B * s.H * s.W; pitch = net->Cin0_p; src_dt = ctx->dtG;
  }
  else {
    size_t dot = nm.rfind('.');
    GAN_REQUIRE(dot != std::string::npos, "debug tensor name must be <layer>.z or <layer>.a");
    std::string lname = nm.substr(0, dot), what = nm.substr(dot + 1);
    int li = -1;
    for (size_t i = 0; i < net->layers.size(); ++i) if (net->layers[i].name == lname) li = (int)i;
    GAN_REQUIRE(li >= 0 && !net->layers[li].head, "unknown layer");
    View ov = s.out_views[li];
    P = ov.pixels(); C = ov.C;
    if (what == "z") {
      GAN_REQUIRE(!(li == 0 && s.z_is_act0), "down1.z is not stored by the first-layer kernel (set GAN_B200_FIRST=0 to inspect it)");
      src = s.z[li].p; pitch = C; coff = 0;
    }
    else if (what == "a") { src = ov.p; pitch = ov.pitch; coff = ov.coff; }
    else GAN_REQUIRE(false, "debug tensor kind must be z or a");
  }
  if (numel) *numel = P * C;
  GAN_REQUIRE(cap >= P * C, "destination too small");
  if (is_f32) {
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    CUDA_CHECK(cudaMemcpy(host_dst, src, P * C * 4, cudaMemcpyDeviceToHost));
  } else {
    DevBuf tmp; tmp.ensure((size_t)P * C * 4);
    launch_export(ctx->L(), src_dt, src, pitch, coff, P, C, tmp.as<float>());
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    CUDA_CHECK(cudaMemcpy(host_dst, tmp.p, P * C * 4, cudaMemcpyDeviceToHost));
  }
  API_END
}

int gan_generator_forward(gan_net* g, const float* x, int batch, float* out) {
  API_BEGIN
  GAN_REQUIRE(g && g->is_gen && x && out && batch >= 1, "bad argument");
  gan_ctx* ctx = g->ctx;
  CUDA_CHECK(cudaSetDevice(ctx->device));
  size_t bytes = (size_t)batch * g->H * g->W * g->C * 4;
  const float* xd = stage_in(ctx, 0, x, bytes);
  ctx->step_epoch++; ctx->im2col_next = 0;
  generator_forward(g, 0, xd, batch, g->H, g->W);
  bump_calls(ctx);
  copy_out(ctx, out, g->slots[0].out_f32.as<float>(), bytes);
  API_END
}

int gan_discriminator_forward(gan_net* d, const float* inp, const float* tar, int batch, int height, int width,
                              float* logits) {
  API_BEGIN
  GAN_REQUIRE(d && !d->is_gen && inp && logits && batch >= 1, "bad argument");
  gan_ctx* ctx = d->ctx;
  CUDA_CHECK(cudaSetDevice(ctx->device));
  size_t bytes = (size_t)batch * height * width * d->C * 4;
  const float* id = stage_in(ctx, 0, inp, bytes);
  const float* td = tar ? stage_in(ctx, 1, tar, bytes) : nullptr;
  ctx->step_epoch++; ctx->im2col_next = 0;
  discriminator_forward(d, 0, id, td, batch, height, width);
  copy_out(ctx, logits, d->slots[0].logits.as<float>(), (size_t)logits_count(d->slots[0]) * 4);
  API_END
}

int gan_adam_create(gan_net* net, double lr, double beta1, double beta2, double eps, gan_adam** out) {
  API_BEGIN
  GAN_REQUIRE(net && out, "null argument");
  CUDA_CHECK(cudaSetDevice(net->ctx->device));
  gan_adam* o = new gan_adam();
  std::unique_ptr<gan_adam> guard(o);
  o->net = net; o->lr = lr; o->b1 = beta1; o->b2 = beta2; o->eps = eps;
  o->m.ensure((size_t)(net->nparams + 1028) * 4); o->v.ensure((size_t)(net->nparams + 1028) * 4);
  o->t_dev.ensure(16);
  guard.release(); net->ctx->adams.push_back(o); live_add(o);
  *out = o;
  API_END
}
int gan_adam_destroy(gan_adam* opt) {
  API_BEGIN
  if (opt) {
    { std::lock_guard<std::mutex> l(g_live_mu); GAN_REQUIRE(g_live.count(opt) == 1, "unknown or already destroyed optimizer"); }
    cudaSetDevice(opt->net->ctx->device); cudaStreamSynchronize(opt->net->ctx->stream); drop_graphs(opt->net->ctx);
    destroy_adam(opt);
  }
  API_END
}
int gan_adam_get_step(gan_adam* opt, int64_t* t) { *t = opt->t; return GAN_OK; }
int gan_adam_set_step(gan_adam* opt, int64_t t) {
  API_BEGIN
  CUDA_CHECK(cudaStreamSynchronize(opt->net->ctx->stream));
  opt->t = t;
  long long tv = t;
  CUDA_CHECK(cudaMemcpy(opt->t_dev.p, &tv, 8, cudaMemcpyHostToDevice));
  API_END
}
int gan_adam_set_hyper(gan_adam* opt, double lr, double beta1, double beta2, double eps) {
  API_BEGIN
  GAN_REQUIRE(opt != nullptr, "null optimizer");
  GAN_REQUIRE(lr >= 0.0 && beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0, "bad Adam hyper-parameters");
  gan_ctx* ctx = opt->net->ctx;
  CUDA_CHECK(cudaSetDevice(ctx->device));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  drop_graphs(ctx);                       // lr / betas are kernel arguments of the captured Adam launches
  opt->lr = lr; opt->b1 = beta1; opt->b2 = beta2; opt->eps = eps;
  API_END
}
int gan_adam_get_state(gan_adam* opt, int which, float* host_dst) {
  API_BEGIN
  CUDA_CHECK(cudaStreamSynchronize(opt->net->ctx->stream));
  CUDA_CHECK(cudaMemcpy(host_dst, which == 0 ? opt->m.p : opt->v.p, opt->net->nparams * 4, cudaMemcpyDeviceToHost));
  API_END
}
int gan_adam_set_state(gan_adam* opt, int which, const float* host_src) {
  API_BEGIN
  CUDA_CHECK(cudaStreamSynchronize(opt->net->ctx->stream));
  CUDA_CHECK(cudaMemcpy(which == 0 ? opt->m.p : opt->v.p, host_src, opt->net->nparams * 4, cudaMemcpyHostToDevice));
  API_END
}

int gan_pix2pix_train_step_ex(gan_net* g, gan_net* d, gan_adam* g_opt, gan_adam* d_opt, const float* input_image,
                              const float* target, int batch, float l1_weight, float gan_grad_scale, int training,
                              float losses[4]) {
  API_BEGIN
  GAN_REQUIRE(g && d && input_image && target, "null argument");
  GAN_REQUIRE(!training || (g_opt && d_opt && g_opt->net == g && d_opt->net == d), "optimizers do not match the nets");
  gan_ctx* ctx = g->ctx;
  CUDA_CHECK(cudaSetDevice(ctx->device));
  if (!ctx->graphs || ctx->profile) {
    pix2pix_step(g, d, g_opt, d_opt, input_image, target, batch, l1_weight, gan_grad_scale, training, losses);
  } else {
    char key[200];
    snprintf(key, sizeof(key), "p2p:%p:%p:%p:%p:%d:%d:%g:%g", (void*)g, (void*)d, (void*)g_opt, (void*)d_opt, batch, training,
             (double)l1_weight, (double)gan_grad_scale);
    const size_t img_bytes = (size_t)batch * g->H * g->W * g->C * 4;
    run_step_graphed(ctx, key, input_image, target, img_bytes, losses, 4, std::vector<gan_net*>{g, d},
                     training ? std::vector<gan_adam*>{g_opt, d_opt} : std::vector<gan_adam*>{},
                     [&](const float* xs, const float* ys) {
                       pix2pix_step(g, d, g_opt, d_opt, xs, ys, batch, l1_weight, gan_grad_scale, training, nullptr);
                     });
  }
  API_END
}
int gan_pix2pix_train_step(gan_net* g, gan_net* d, gan_adam* g_opt, gan_adam* d_opt, const float* input_image,
                           const float* target, int batch, float lambda, int training, float losses[4]) {
  return gan_pix2pix_train_step_ex(g, d, g_opt, d_opt, input_image, target, batch, lambda, 1.0f, training, losses);
}

int gan_cyclegan_train_step(gan_net* g, gan_net* f, gan_net* dx, gan_net* dy, gan_adam* g_opt, gan_adam* f_opt,
                            gan_adam* dx_opt, gan_adam* dy_opt, const float* real_x, const float* real_y, int batch,
                            float lambda, int training, float losses[7]) {
  API_BEGIN
  GAN_REQUIRE(g && f && dx && dy && real_x && real_y, "null argument");
  GAN_REQUIRE(!training || (g_opt && f_opt && dx_opt && dy_opt && g_opt->net == g && f_opt->net == f &&
                            dx_opt->net == dx && dy_opt->net == dy), "optimizers do not match the nets");
  gan_ctx* ctx = g->ctx;
  CUDA_CHECK(cudaSetDevice(ctx->device));
  if (!ctx->graphs || ctx->profile) {
    cyclegan_step(g, f, dx, dy, g_opt, f_opt, dx_opt, dy_opt, real_x, real_y, batch, lambda, training, losses);
  } else {
    char key[200];
    snprintf(key, sizeof(key), "cyc:%p:%p:%p:%p:%d:%d:%g", (void*)g, (void*)f, (void*)dx, (void*)dy, batch, training, (double)lambda);
    const size_t img_bytes = (size_t)batch * g->H * g->W * g->C * 4;
    run_step_graphed(ctx, key, real_x, real_y, img_bytes, losses, 7, std::vector<gan_net*>{g, f, dx, dy},
                     training ? std::vector<gan_adam*>{g_opt, f_opt, dx_opt, dy_opt} : std::vector<gan_adam*>{},
                     [&](const float* xs, const float* ys) {
                       cyclegan_step(g, f, dx, dy, g_opt, f_opt, dx_opt, dy_opt, xs, ys, batch, lambda, training, nullptr);
                     });
  }
  API_END
}

static void ensure_copy_stream(gan_ctx* ctx);
int gan_ctx_prefetch(gan_ctx* ctx, const float* x_host, const float* y_host, int64_t bytes_each) {
  API_BEGIN
  GAN_REQUIRE(ctx && x_host && y_host && bytes_each > 0, "bad argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  ensure_copy_stream(ctx);
  // the previous batch is moved out of the prefetch buffers at the very start of the step that consumes
  // it; the new copy only has to wait for that device-to-device move, not for the whole step
  CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->prefetch_consumed, 0));
  ctx->prefetch_buf[0].ensure((size_t)bytes_each); ctx->prefetch_buf[1].ensure((size_t)bytes_each);
  CUDA_CHECK(cudaMemcpyAsync(ctx->prefetch_buf[0].p, x_host, (size_t)bytes_each, cudaMemcpyDefault, ctx->copy_stream));
  CUDA_CHECK(cudaMemcpyAsync(ctx->prefetch_buf[1].p, y_host, (size_t)bytes_each, cudaMemcpyDefault, ctx->copy_stream));
  CUDA_CHECK(cudaEventRecord(ctx->prefetch_done, ctx->copy_stream));
  ctx->prefetch_src[0] = x_host; ctx->prefetch_src[1] = y_host; ctx->prefetch_bytes = (size_t)bytes_each;
  API_END
}

static void check_xforms(const gan_image_xform* xf, int batch, int out_size, int64_t stride, int channels) {
  for (int n = 0; n < batch; ++n) {
    const gan_image_xform& x = xf[n];
    GAN_REQUIRE(x.src_h >= 1 && x.src_w >= 1 && x.cols >= 1 && x.col0 >= 0 && x.col0 + x.cols <= x.src_w, "bad image window");
    GAN_REQUIRE((int64_t)x.src_h * x.src_w * channels <= stride, "image does not fit its stride");
    GAN_REQUIRE(x.pre >= 0 && x.mid >= 0 && (x.mid == 0 || x.mid >= out_size), "bad resize sizes");
    GAN_REQUIRE(x.mid == 0 ? (x.crop_y == 0 && x.crop_x == 0)
                           : (x.crop_y >= 0 && x.crop_x >= 0 && x.crop_y + out_size <= x.mid && x.crop_x + out_size <= x.mid),
                "crop window outside the resized image");
  }
}
static void ensure_copy_stream(gan_ctx* ctx) {
  if (ctx->copy_stream) return;
  CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaEventCreateWithFlags(&ctx->prefetch_done, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreateWithFlags(&ctx->prefetch_consumed, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventRecord(ctx->prefetch_consumed, ctx->stream));
}
// uint8 images (host or device) + transforms -> fp32 images at `out` (device), all on stream `st`
static void preprocess_on(gan_ctx* ctx, cudaStream_t st, int slot, const uint8_t* images, int64_t stride, const gan_image_xform* xf,
                          int batch, int channels, int out_size, float* out_dev, const uint8_t* reuse_src, const uint8_t* reuse_dev,
                          const uint8_t** staged) {
  const uint8_t* src = images;
  if (images == reuse_src && reuse_dev != nullptr) src = reuse_dev;            // same host batch already on the device
  else if (!is_device_ptr(images)) {
    ctx->u8_stage[slot].ensure((size_t)batch * stride);
    CUDA_CHECK(cudaMemcpyAsync(ctx->u8_stage[slot].p, images, (size_t)batch * stride, cudaMemcpyHostToDevice, st));
    src = (const uint8_t*)ctx->u8_stage[slot].p;
  }
  if (staged) *staged = src;
  std::vector<ImageXformDev> dev(batch);
  for (int n = 0; n < batch; ++n) {
    const gan_image_xform& x = xf[n];
    ImageXformDev& d = dev[n];
    d.src_h = x.src_h; d.src_w = x.src_w; d.col0 = x.col0; d.cols = x.cols; d.pre = x.pre; d.mid = x.mid;
    d.crop_y = x.crop_y; d.crop_x = x.crop_x; d.flip = x.flip;
    const int g1h = x.pre > 0 ? x.pre : x.src_h, g1w = x.pre > 0 ? x.pre : x.cols, tgt = x.mid > 0 ? x.mid : out_size;
    d.sy1 = (float)g1h / (float)tgt; d.sx1 = (float)g1w / (float)tgt;           // float32 quotients, as TF computes them
    d.sy0 = x.pre > 0 ? (float)x.src_h / (float)x.pre : 1.f; d.sx0 = x.pre > 0 ? (float)x.cols / (float)x.pre : 1.f;
  }
  ctx->xf_dev[slot].ensure((size_t)batch * sizeof(ImageXformDev));
  // pageable source: the runtime has staged the bytes when cudaMemcpyAsync returns
  CUDA_CHECK(cudaMemcpyAsync(ctx->xf_dev[slot].p, dev.data(), (size_t)batch * sizeof(ImageXformDev), cudaMemcpyHostToDevice, st));
  Launch L = ctx->L(); L.s = st;
  launch_preprocess(L, src, stride, (const ImageXformDev*)ctx->xf_dev[slot].p, batch, channels, out_size, out_dev);
}

int gan_preprocess_images(gan_ctx* ctx, const uint8_t* images, int64_t image_stride, int batch, int channels, int out_size,
                          const gan_image_xform* xf, float* out) {
  API_BEGIN
  GAN_REQUIRE(ctx && images && xf && out, "null argument");
  GAN_REQUIRE(batch >= 1 && channels >= 1 && channels <= 4 && out_size >= 1 && image_stride >= 1, "bad argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  check_xforms(xf, batch, out_size, image_stride, channels);
  const size_t out_bytes = (size_t)batch * out_size * out_size * channels * 4;
  float* dst = out;
  if (!is_device_ptr(out)) { ctx->stage[2].ensure(out_bytes); dst = ctx->stage[2].as<float>(); }
  // slot 2: the synchronous path never shares its uint8 staging / transform table with a prefetch in flight
  preprocess_on(ctx, ctx->stream, 2, images, image_stride, xf, batch, channels, out_size, dst, nullptr, nullptr, nullptr);
  // device `out`: asynchronous on the context stream (the pageable transform array has been staged by the
  // runtime when cudaMemcpyAsync returns); host `out`: copied back and synchronised
  if (dst != out) {
    CUDA_CHECK(cudaMemcpyAsync(out, dst, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
  API_END
}

int gan_ctx_prefetch_images(gan_ctx* ctx, const uint8_t* images_a, int64_t stride_a, const gan_image_xform* xf_a,
                            const uint8_t* images_b, int64_t stride_b, const gan_image_xform* xf_b, int batch, int channels,
                            int out_size, const float** a_dev, const float** b_dev) {
  API_BEGIN
  GAN_REQUIRE(ctx && images_a && images_b && xf_a && xf_b && a_dev && b_dev, "null argument");
  GAN_REQUIRE(batch >= 1 && channels >= 1 && channels <= 4 && out_size >= 1 && stride_a >= 1 && stride_b >= 1, "bad argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  check_xforms(xf_a, batch, out_size, stride_a, channels);
  check_xforms(xf_b, batch, out_size, stride_b, channels);
  ensure_copy_stream(ctx);
  const size_t bytes = (size_t)batch * out_size * out_size * channels * 4;
  // the previous prefetched batch leaves these buffers at the start of the step that consumes it
  CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->prefetch_consumed, 0));
  ctx->prefetch_buf[0].ensure(bytes); ctx->prefetch_buf[1].ensure(bytes);
  const uint8_t* staged_a = nullptr;
  preprocess_on(ctx, ctx->copy_stream, 0, images_a, stride_a, xf_a, batch, channels, out_size, ctx->prefetch_buf[0].as<float>(),
                nullptr, nullptr, &staged_a);
  preprocess_on(ctx, ctx->copy_stream, 1, images_b, stride_b, xf_b, batch, channels, out_size, ctx->prefetch_buf[1].as<float>(),
                stride_a == stride_b ? images_a : nullptr, staged_a, nullptr);
  CUDA_CHECK(cudaEventRecord(ctx->prefetch_done, ctx->copy_stream));
  ctx->prefetch_src[0] = ctx->prefetch_buf[0].p; ctx->prefetch_src[1] = ctx->prefetch_buf[1].p; ctx->prefetch_bytes = bytes;
  *a_dev = ctx->prefetch_buf[0].as<float>(); *b_dev = ctx->prefetch_buf[1].as<float>();
  API_END
}

int gan_ctx_last_losses(gan_ctx* ctx, float* out, int n) {
  API_BEGIN
  GAN_REQUIRE(n >= 0 && n <= ctx->n_losses, "more losses requested than the last step produced");
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->loss_host, (size_t)n * 4);
  API_END
}

int gan_op_conv(gan_ctx* ctx, int kind, int role, int engine, const float* a, const float* b, float* out, int batch,
                int height, int width, int cin, int cout) {
  API_BEGIN
  GAN_REQUIRE(ctx && a && b && out, "null argument");
  GAN_REQUIRE(kind >= 0 && kind <= 2 && role >= 0 && role <= 2, "bad kind/role");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  Layer ly; ly.name = "op"; ly.kind = kind; ly.Cin = cin; ly.Cout = cout; ly.norm = NORM_NONE; ly.act = ACT_NONE;
  const int cin_p = pad_c(ctx, cin), cout_p = pad_c(ctx, cout);
  ly.Cin_p = cin_p; ly.Cout_p = cout_p;
  int Ho, Wo; out_dims(kind, height, width, Ho, Wo);
  const int64_t px = (int64_t)batch * height * width, py = (int64_t)batch * Ho * Wo;
  const int64_t nx = px * cin, ny = py * cout, nw = 16LL * cin * cout;
  const size_t es = ctx->esize();
  DevBuf fa, fb, fo, xa, ya, wp;
  Launch L = ctx->L();
  int saved = ctx->engine; ctx->engine = engine;
  try {
    auto up = [&](DevBuf& stage, const float* h, int64_t n) {
      stage.ensure((size_t)n * 4);
      CUDA_CHECK(cudaMemcpyAsync(stage.p, h, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    };
    View x = make_view(nullptr, batch, height, width, cin_p), y = make_view(nullptr, batch, Ho, Wo, cout_p);
    xa.ensure((size_t)px * cin_p * es); ya.ensure((size_t)py * cout_p * es);
    x.p = xa.p; y.p = ya.p;
    if (role != R_WGRAD) {
      // pack the kernel for this role
      up(fb, b, nw);
      PackOp po; memset(&po, 0, sizeof(po));
      po.ncls = fill_geometry(kind, role, po.cls);
      weight_strides(ly, role, po.Kc, po.Nc, po.Kr, po.Nr, po.s_tap, po.s_k, po.s_n);
      for (int c = 0; c < po.ncls; ++c) po.cls[c].b_off = (int64_t)c * po.Nc * po.cls[c].ntaps * po.Kc;
      wp.ensure((size_t)16 * cin_p * cout_p * es);
      launch_pack(L, role == R_FWD ? ctx->dtA : ctx->dtG, fb.as<float>(), wp.p, po);
      if (role == R_FWD) {
        up(fa, a, nx);
        launch_convert(L, ctx->dtA, fa.as<float>(), px, cin, x.p, cin_p, 0);
        run_conv_fwd(ctx, make_op(ctx, ly, R_FWD, x, y, wp.p));
        fo.ensure((size_t)ny * 4);
        launch_export(L, ctx->dtA, y.p, cout_p, 0, py, cout, fo.as<float>());
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        CUDA_CHECK(cudaMemcpy(out, fo.p, ny * 4, cudaMemcpyDeviceToHost));
      } else {
        up(fa, a, ny);
        launch_convert(L, ctx->dtG, fa.as<float>(), py, cout, y.p, cout_p, 0);
        run_conv_fwd(ctx, make_op(ctx, ly, R_DGRAD, x, y, wp.p));
        fo.ensure((size_t)nx * 4);
        launch_export(L, ctx->dtG, x.p, cin_p, 0, px, cin, fo.as<float>());
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        CUDA_CHECK(cudaMemcpy(out, fo.p, nx * 4, cudaMemcpyDeviceToHost));
      }
    } else {
      up(fa, a, nx); up(fb, b, ny);
      launch_convert(L, ctx->dtA, fa.as<float>(), px, cin, x.p, cin_p, 0);
      launch_convert(L, ctx->dtG, fb.as<float>(), py, cout, y.p, cout_p, 0);
      fo.ensure((size_t)nw * 4);
      CUDA_CHECK(cudaMemsetAsync(fo.p, 0, nw * 4, ctx->stream));
      ConvOp op = make_op(ctx, ly, R_WGRAD, x, y, nullptr);
      op.dW = fo.as<float>();
      ly.wgrad_epoch = 0;                      // first contribution: the kernel must define every element itself
      run_conv_wgrad(ctx, ly, op);
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      CUDA_CHECK(cudaMemcpy(out, fo.p, nw * 4, cudaMemcpyDeviceToHost));
    }
    CUDA_CHECK(cudaGetLastError());
  } catch (...) { ctx->engine = saved; throw; }
  ctx->engine = saved;
  API_END
}

}  // extern "C"
