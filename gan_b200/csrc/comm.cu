// comm.cu — data-parallel gradient exchange over NCCL (NVLink 5 / NVSwitch).
//
// The reference is single-device (base_gan.py:18-19 only prints the GPU count); data parallelism
// is new work defined by BASELINE.json.  NCCL is resolved at run time with dlopen so that the
// library links against nothing but cudart: one process per GPU loads the same libnccl that
// torch.distributed already mapped; the unique id is exchanged by the Python host through
// torch.distributed (plumbing only).
#include <dlfcn.h>
#include <cstring>
#include "engine.h"

typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void* ncclComm_t_;
typedef int (*fn_getuid)(ncclUniqueId_t*);
typedef int (*fn_initrank)(ncclComm_t_*, int, ncclUniqueId_t, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, ncclComm_t_, cudaStream_t);
typedef int (*fn_reducescatter)(const void*, void*, size_t, int, int, ncclComm_t_, cudaStream_t);
typedef int (*fn_allgather)(const void*, void*, size_t, int, ncclComm_t_, cudaStream_t);
typedef int (*fn_group)(void);
typedef int (*fn_destroy)(ncclComm_t_);
typedef const char* (*fn_errstr)(int);

static struct {
  void* lib = nullptr;
  fn_getuid getuid = nullptr; fn_initrank initrank = nullptr; fn_allreduce allreduce = nullptr;
  fn_destroy destroy = nullptr; fn_errstr errstr = nullptr;
  fn_reducescatter reducescatter = nullptr; fn_allgather allgather = nullptr; fn_group group_start = nullptr, group_end = nullptr;
} g_nccl;

static void nccl_load() {
  if (g_nccl.lib) return;
  const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; names[i] && !g_nccl.lib; ++i) g_nccl.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!g_nccl.lib) throw GanError(-4, std::string("cannot dlopen libnccl: ") + dlerror());
  g_nccl.getuid = (fn_getuid)dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.initrank = (fn_initrank)dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.allreduce = (fn_allreduce)dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.destroy = (fn_destroy)dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.errstr = (fn_errstr)dlsym(g_nccl.lib, "ncclGetErrorString");
  g_nccl.reducescatter = (fn_reducescatter)dlsym(g_nccl.lib, "ncclReduceScatter");
  g_nccl.allgather = (fn_allgather)dlsym(g_nccl.lib, "ncclAllGather");
  g_nccl.group_start = (fn_group)dlsym(g_nccl.lib, "ncclGroupStart");
  g_nccl.group_end = (fn_group)dlsym(g_nccl.lib, "ncclGroupEnd");
  if (!g_nccl.getuid || !g_nccl.initrank || !g_nccl.allreduce || !g_nccl.destroy || !g_nccl.reducescatter || !g_nccl.allgather ||
      !g_nccl.group_start || !g_nccl.group_end)
    throw GanError(-4, "libnccl is missing required symbols");
}
static void nccl_check(int r, const char* what) {
  if (r != 0) throw GanError(-4, std::string(what) + ": " + (g_nccl.errstr ? g_nccl.errstr(r) : "nccl error"));
}

int comm_unique_id(void* out128) {
  try { nccl_load(); } catch (...) { return -1; }
  ncclUniqueId_t id;
  if (g_nccl.getuid(&id) != 0) return -1;
  memcpy(out128, &id, 128);
  return 0;
}

void comm_init(gan_ctx* ctx, int rank, int world, const void* id128) {
  GAN_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank/world");
  ctx->rank = rank; ctx->world = world;
  if (world == 1) return;
  nccl_load();
  ncclUniqueId_t id; memcpy(&id, id128, 128);
  ncclComm_t_ c = nullptr;
  nccl_check(g_nccl.initrank(&c, world, id, rank), "ncclCommInitRank");
  ctx->comm = c;
  CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 16; ++i) {
    cudaEvent_t e; CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->comm_events.push_back(e);
  }
}

void comm_destroy(gan_ctx* ctx) {
  if (ctx->comm && g_nccl.destroy) g_nccl.destroy((ncclComm_t_)ctx->comm);
  ctx->comm = nullptr;
  for (auto e : ctx->comm_events) cudaEventDestroy(e);
  ctx->comm_events.clear();
  if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
  ctx->comm_stream = nullptr;
}

static cudaEvent_t next_event(gan_ctx* ctx) {
  cudaEvent_t e = ctx->comm_events[ctx->comm_ev_next % ctx->comm_events.size()];
  ctx->comm_ev_next++;
  return e;
}

void comm_allreduce_async(gan_ctx* ctx, float* buf, int64_t n) {
  if (ctx->world <= 1 || n <= 0) return;
  GAN_REQUIRE(ctx->comm != nullptr, "communicator not initialised");
  cudaEvent_t e = next_event(ctx);
  CUDA_CHECK(cudaEventRecord(e, ctx->stream));
  CUDA_CHECK(cudaStreamWaitEvent(ctx->comm_stream, e, 0));
  nccl_check(g_nccl.allreduce(buf, buf, (size_t)n, 7, 0, (ncclComm_t_)ctx->comm, ctx->comm_stream), "ncclAllReduce");
  ctx->comm_pending = true;
}

// The same on a bf16 buffer (ncclBfloat16 = 9): half the bytes on NVLink and half the time the NCCL kernels hold SMs.
void comm_allreduce_bf16_async(gan_ctx* ctx, void* buf, int64_t n) {
  if (ctx->world <= 1 || n <= 0) return;
  GAN_REQUIRE(ctx->comm != nullptr, "communicator not initialised");
  cudaEvent_t e = next_event(ctx);
  CUDA_CHECK(cudaEventRecord(e, ctx->stream));
  CUDA_CHECK(cudaStreamWaitEvent(ctx->comm_stream, e, 0));
  nccl_check(g_nccl.allreduce(buf, buf, (size_t)n, 9, 0, (ncclComm_t_)ctx->comm, ctx->comm_stream), "ncclAllReduce(bf16)");
  ctx->comm_pending = true;
}

// Sharded form (ZeRO-1 style): the bucket [buf, buf+n) (n a multiple of world) is reduce-scattered in
// place — rank r ends up with the sum of sub-range r, [buf + r*n/world, +n/world) — on the communication stream.
void comm_reducescatter_async(gan_ctx* ctx, float* buf, int64_t n) {
  if (ctx->world <= 1 || n <= 0) return;
  GAN_REQUIRE(ctx->comm != nullptr, "communicator not initialised");
  GAN_REQUIRE(n % ctx->world == 0, "reduce-scatter bucket must divide evenly over the ranks");
  const int64_t cnt = n / ctx->world;
  cudaEvent_t e = next_event(ctx);
  CUDA_CHECK(cudaEventRecord(e, ctx->stream));
  CUDA_CHECK(cudaStreamWaitEvent(ctx->comm_stream, e, 0));
  nccl_check(g_nccl.reducescatter(buf, buf + (int64_t)ctx->rank * cnt, (size_t)cnt, 7, 0, (ncclComm_t_)ctx->comm, ctx->comm_stream),
             "ncclReduceScatter");
  ctx->comm_pending = true;
}
// All-gather (in place) of the ranks' updated sub-ranges of several buckets as ONE grouped NCCL launch on the ctx stream.
void comm_allgather_buckets(gan_ctx* ctx, float* base, const int64_t* off, const int64_t* len, int nb) {
  if (ctx->world <= 1 || nb <= 0) return;
  GAN_REQUIRE(ctx->comm != nullptr, "communicator not initialised");
  nccl_check(g_nccl.group_start(), "ncclGroupStart");
  for (int i = 0; i < nb; ++i) {
    const int64_t cnt = len[i] / ctx->world;
    nccl_check(g_nccl.allgather(base + off[i] + (int64_t)ctx->rank * cnt, base + off[i], (size_t)cnt, 7, (ncclComm_t_)ctx->comm, ctx->stream),
               "ncclAllGather");
  }
  nccl_check(g_nccl.group_end(), "ncclGroupEnd");
}

void comm_join(gan_ctx* ctx) {
  if (!ctx->comm_pending) return;
  cudaEvent_t e = next_event(ctx);
  CUDA_CHECK(cudaEventRecord(e, ctx->comm_stream));
  CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, e, 0));
  ctx->comm_pending = false;
}

// Sum-all-reduce of an fp32 buffer, enqueued on the ctx stream (ncclFloat32 = 7, ncclSum = 0).
void comm_allreduce_sum(gan_ctx* ctx, float* buf, int64_t n) {
  if (ctx->world <= 1) return;
  GAN_REQUIRE(ctx->comm != nullptr, "communicator not initialised");
  nccl_check(g_nccl.allreduce(buf, buf, (size_t)n, 7, 0, (ncclComm_t_)ctx->comm, ctx->stream), "ncclAllReduce");
}
