// engine.h — host-side objects behind the C-ABI handles (gan_ctx / gan_net / gan_adam).
#pragma once
#include <string>
#include <vector>
#include <map>
#include "kernels.h"

// Bumped whenever a device buffer is re-allocated: captured CUDA graphs hold raw pointers, so a graph
// captured under an older epoch must not be replayed (run_step_graphed re-captures it).
extern uint64_t g_alloc_epoch;

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
  ~DevBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  // grows only; contents are NOT preserved on growth
  void ensure(size_t n) {
    if (n <= bytes) return;
    if (p) ++g_alloc_epoch;               // an address that a captured graph may have baked in goes away
    release();
    CUDA_CHECK(cudaMalloc(&p, n));
    CUDA_CHECK(cudaMemset(p, 0, n));
    CUDA_CHECK(cudaDeviceSynchronize());   // the memset runs on the legacy stream; ctx streams are non-blocking
    bytes = n;
  }
  template <typename T> T* as() const { return (T*)p; }
};

struct CommApi;   // dlopen'ed NCCL entry points (comm.cu)

// Optional per-kernel-family timing with CUDA events on the ctx stream (bench.py roofline).
enum { FAM_UMMA_FWD = 0, FAM_UMMA_WGRAD = 1, FAM_FFMA_FWD = 2, FAM_FFMA_WGRAD = 3, FAM_NORM = 4, FAM_ADAM = 5,
       FAM_PACK = 6, FAM_OTHER = 7, FAM_COUNT = 8 };
struct ProfEntry { cudaEvent_t a, b; int fam; double work; };

struct gan_ctx {
  int device = 0;
  int dt = DT_F32;            // mode: DT_F32 (FFMA parity mode) or DT_BF16 (= the 16-bit tcgen05 mode)
  int dtA = DT_F32;           // storage dtype of activations / forward weight packs (16-bit mode: DT_F16, see common.cuh)
  int dtG = DT_F32;           // storage dtype of gradients / data-gradient weight packs (always == dtA, see common.cuh)
  float grad_scale = 1.f;     // loss scale of the current step: every gradient buffer holds grad_scale x the true gradient
  cudaStream_t stream = nullptr;
  uint64_t seed = 0;
  uint32_t call_counter = 0;
  int dropout_enabled = 1;
  int engine = -1;
  int graphs = 0;
  int64_t sample0 = 0;
  bool sample0_set = false;
  uint64_t launches = 0;
  // Workspaces that layer code writes, one set per stream: the side stream runs a whole discriminator pass while the
  // main stream runs the generator (engine.cu side_begin / side_join), so they must not share scratch memory.
  struct Scratch { DevBuf stats_ws, dz_scratch, junk, splitk_ws, wgrad_ws, head_part, counters; };   // counters: zeroed tickets of the last-block sums
  Scratch scr[2];
  int cur = 0;                                // 0: main stream, 1: side stream (selects L(), cs(), sc())
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_mid = nullptr;
  int overlap = 1;                            // run independent sub-graphs of a step on the side stream
  Scratch& sc() { return scr[cur]; }
  cudaStream_t cs() const { return cur ? side : stream; }
  DevBuf loss_ws, loss_out;
  DevBuf stage[4];
  // im2col rows of step inputs shared between nets (x feeds G.down1, D(real).down1 and D(fake).down1)
  struct Im2colEntry { const void* src = nullptr; int B = 0, H = 0, W = 0, C = 0; uint64_t epoch = 0; DevBuf buf; };
  Im2colEntry im2col_cache[6];
  // every net / optimizer created from this context (destroyed with it)
  std::vector<struct gan_net*> nets;
  std::vector<struct gan_adam*> adams;
  uint64_t step_epoch = 1;
  int im2col_next = 0;
  // input prefetch (tf.data-style): H2D of the NEXT step's images on a copy stream while this step computes
  DevBuf prefetch_buf[2];
  DevBuf u8_stage[3], xf_dev[3];   // input pipeline: uint8 images and per-image transforms on the device
  const void* prefetch_src[2] = {nullptr, nullptr};
  size_t prefetch_bytes = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t prefetch_done = nullptr, prefetch_consumed = nullptr;
  float* loss_host = nullptr;   // pinned
  int n_losses = 0;
  // device-resident dropout call counter (graph-replay safe) + generator calls since the last bump
  DevBuf call_dev;
  uint32_t gen_calls_pending = 0;
  // captured train steps, keyed on (nets, batch, training)
  struct GraphEntry { cudaGraphExec_t exec = nullptr; uint64_t launches = 0; int warm = 0; uint32_t gen_calls = 0; uint64_t alloc_epoch = 0; };
  std::map<std::string, GraphEntry> graph_cache;
  // profiling
  int profile = 0;
  std::vector<ProfEntry> prof;
  std::vector<cudaEvent_t> ev_pool;
  cudaEvent_t ev_get() {
    if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
    cudaEvent_t e; CUDA_CHECK(cudaEventCreate(&e)); return e;
  }
  // data parallel
  void* comm = nullptr;
  int rank = 0, world = 1;
  cudaStream_t comm_stream = nullptr;     // gradient all-reduces run here, overlapped with backward
  std::vector<cudaEvent_t> comm_events;   // fork/join events (reused round-robin)
  size_t comm_ev_next = 0;
  bool comm_pending = false;
  int comm16 = 0;                         // world > 1, 16-bit mode: gradients travel as bf16 (half the NCCL bytes)
  int shard_optimizer = 1;                // world > 1: reduce-scatter + 1/world Adam + all-gather (0: all-reduce + full Adam)
  Launch L() { return Launch{cs(), &launches}; }
  size_t esize() const { return dt == DT_F32 ? 4 : 2; }
};

struct TensorInfo {
  std::string name;
  int ndim;
  int64_t shape[4];
  int64_t numel;
  int64_t off;        // float offset in params (trainable) or mov (moving stats)
  bool trainable;
};

struct Layer {
  std::string name;
  int kind, Cin, Cout, norm, act;
  int Cin_p = 0, Cout_p = 0;     // channel counts as stored (zero-padded to 16 in bf16 mode when < 16)
  bool bias = false, dropout = false, need_dgrad = true, head = false;
  int tag = 0;
  uint64_t wgrad_epoch = 0;      // step in which this layer's weight gradient was last written (first write stores, later ones add)
  int64_t w_off = -1, g_off = -1, b_off = -1, bias_off = -1, mov_off = -1;
  DevBuf wp_fwd, wp_dgrad;
  // first layers in bf16/tcgen05 mode: GEMM over im2col buffers (one 64-wide K-block per input source)
  bool first = false; int nsrc = 1, src_c = 0;
  DevBuf wp_im2col;
  // generator head in bf16/tcgen05 mode: Conv2DTranspose as GEMM over `cols` + col2im (see engine.cu)
  DevBuf wp_cols, wp_dcols;
};

// Saved state of one forward call of a net (the "tape" of that call).
struct Slot {
  int B = 0, H = 0, W = 0;
  uint32_t call_id = 0, call_off = 0;
  int64_t sample0 = 0;
  std::vector<DevBuf> z;        // raw conv outputs per layer
  std::vector<DevBuf> stats;    // per layer: mean, inv, scale, shift, c1, c2  ([G][C] each)
  std::vector<View> in_views, out_views;   // per layer: input activation view, activated output view
  // generator
  DevBuf xin, d8, out_f32, dd8, dxin;
  std::vector<DevBuf> cat, dcat, dskip;
  // discriminator
  DevBuf in0, logits, dlogit, din0, dcols0;     // dcols0: per-tap products of the first layer's data gradient
  int din0_pitch = 0, din0_coff = 0;            // layout of din0: (Cin0_p, first wanted channel) or compact (4, 0)
  const void* im2col[2] = {nullptr, nullptr};   // first-layer im2col rows per input source (ctx cache entries)
  const float* src_f32[2] = {nullptr, nullptr}; // the fp32 input images of this call (first-layer kernel, lazy rows for wgrad)
  DevBuf cols, gcols;        // generator head: cols = x*W (forward), gcols = im2col(dz) (backward)
  bool used_cols = false;
  bool used_im2col = false;
  bool z_is_act0 = false;     // layer 0 ran on the first-layer kernel: no z stored, the activation view stands in for it
  std::vector<DevBuf> act, dact;
};

struct gan_net {
  gan_ctx* ctx = nullptr;
  bool is_gen = false;
  int norm = NORM_BATCH, H = 0, W = 0, C = 0, Cin0 = 0, Cp = 0, Cin0_p = 0;
  bool target = false;
  std::vector<Layer> layers;
  std::vector<TensorInfo> tensors;
  int ntrain = 0;
  int64_t nparams = 0, nmov = 0;
  DevBuf params, grads, mov;
  DevBuf grads16;             // data parallel: bf16 copy of the gradient buckets, the buffer NCCL all-reduces (ctx->comm16)
  bool grads16_valid = false; // grads16 holds the all-reduced gradient of the last step (the getters read it)
  float grad_scale = 1.f;     // loss scale the contents of `grads` carry (removed by Adam / by the gradient getters)
  // data parallel, sharded optimizer: the gradient buckets reduce-scattered so far in this step (flat offset, length;
  // lengths are multiples of the world size) — rank r owns sub-range r of every bucket
  std::vector<int64_t> bucket_off, bucket_len;
  int64_t reduced_from = -1;  // lowest flat offset already handed to the communication stream (buckets go top-down)
  std::vector<Slot> slots;
  bool packed_dirty = true;
  DevBuf pack_tab;            // device array of PackEntry (all layers x roles), built once
  int pack_nent = 0, pack_tiles = 0;
  // extra packed copies through index tables: first layer in im2col K order, head as cols GEMM operands
  struct GatherTab { DevBuf idx; void* dst = nullptr; int n = 0; int dt = DT_F32; };
  std::vector<GatherTab> gathers;
  DevBuf adam_tab, adam_ranges;   // fused Adam+pack tables (AdamPackEntry / AdamRange)
  int adam_nent = 0, adam_tiles = 0, adam_nranges = 0;
};

struct gan_adam {
  gan_net* net = nullptr;
  double lr, b1, b2, eps;
  int64_t t = 0;             // host mirror of *t_dev
  DevBuf m, v, t_dev;
};

// comm.cu
int comm_unique_id(void* out128);
void comm_init(gan_ctx* ctx, int rank, int world, const void* id128);
void comm_destroy(gan_ctx* ctx);
void comm_allreduce_sum(gan_ctx* ctx, float* buf, int64_t n);
// Fork: all-reduce [buf, buf+n) on the communication stream once everything enqueued so far on the
// compute stream has finished; comm_join makes the compute stream wait for all forked reductions.
void comm_allreduce_async(gan_ctx* ctx, float* buf, int64_t n);
void comm_allreduce_bf16_async(gan_ctx* ctx, void* buf, int64_t n);
void comm_join(gan_ctx* ctx);
void comm_reducescatter_async(gan_ctx* ctx, float* buf, int64_t n);
void comm_allgather_buckets(gan_ctx* ctx, float* base, const int64_t* off, const int64_t* len, int nb);
