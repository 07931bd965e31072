// dp.cu — data parallelism behind the C ABI.  One process per GPU; the only exchange step of the
// training-step path is the sum of the flat gradient arena over ranks, followed by the identical fused optimizer update
// on every rank.  The reference has no distributed path (no Distributed / MPI / NCCL anywhere in the package).
//
// The communicator is this library's own: libnccl is dlopen'ed (no link dependency, so the library still loads on a box
// without NCCL), created from a 128-byte id the host distributes, with its own CTA budget (ncclConfig_t::maxCTAs) and
// its own stream.  Collectives are ordered behind the ctx stream with events; pg_dp_wait / pg_dp_bucket_wait order the
// ctx stream behind them.  Buckets: pg_plan_backward reports finished gradient leaves last-to-first (arena tail first);
// once enough bytes have accumulated their slice is reduced while the remaining backward kernels run.
#include <dlfcn.h>
#include <limits.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"

// from arena.cu
void* pg_arena_base(pg_arena* a);
int pg_arena_dtype(pg_arena* a);
int64_t pg_arena_padded(pg_arena* a);
const std::vector<int64_t>& pg_arena_offsets(pg_arena* a);

namespace {

// ---- the slice of NCCL's public C interface this file uses (types restated so that no NCCL header is needed) ----
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;                      // 0 = success
enum { ncclInt32 = 2, ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat64 = 8, ncclBfloat16 = 9 };
enum { ncclSum = 0 };
struct ncclConfig_v22700 {                     // ncclConfig_t as of NCCL 2.27 (newer libraries accept older, smaller structs: size / version fields)
  size_t size; unsigned int magic; unsigned int version;
  int blocking, cgaClusterSize, minCTAs, maxCTAs;
  const char* netName;
  int splitShare, trafficClass;
  const char* commName;
  int collnetEnable, CTAPolicy, shrinkShare, nvlsCTAs;
};

struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_v22700*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
Nccl g_nccl;

int nccl_load() {
  if (g_nccl.lib) return PG_OK;
  const char* env = getenv("PG_NCCL_LIB");
  void* h = nullptr;
  if (env && env[0]) h = dlopen(env, RTLD_NOW | RTLD_LOCAL);
  // a libnccl that is already part of the process (e.g. the one torch loaded) is shared, not duplicated
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) {
    pg_set_error("cannot load libnccl.so.2 (%s); set PG_NCCL_LIB to its path", dlerror());
    return PG_ERR_UNSUPPORTED;
  }
  g_nccl.lib = h;
#define SYM(field, name) *(void**)(&g_nccl.field) = dlsym(h, name)
  SYM(GetVersion, "ncclGetVersion"); SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommInitRankConfig, "ncclCommInitRankConfig"); SYM(CommDestroy, "ncclCommDestroy"); SYM(AllReduce, "ncclAllReduce");
  SYM(Broadcast, "ncclBroadcast"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.Broadcast || !g_nccl.CommDestroy) {
    pg_set_error("libnccl is missing a required symbol");
    g_nccl.lib = nullptr;
    return PG_ERR_UNSUPPORTED;
  }
  return PG_OK;
}

#define PG_CHECK_NCCL(expr)                                                                                    \
  do {                                                                                                         \
    ncclResult_t _r = (expr);                                                                                  \
    if (_r != 0) {                                                                                             \
      pg_set_error("%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?", __FILE__, __LINE__); \
      return PG_ERR_CUDA;                                                                                      \
    }                                                                                                          \
  } while (0)

int nccl_type(int dtype) {
  switch (dtype) {
    case PG_F32: return ncclFloat32;
    case PG_F64: return ncclFloat64;
    case PG_BF16: return ncclBfloat16;
    case PG_F16: return ncclFloat16;
    case PG_I32: return ncclInt32;
  }
  return -1;
}

struct Bucket { int64_t start, count; cudaEvent_t done; };

}  // namespace

struct pg_dp {
  pg_ctx* ctx = nullptr;
  int rank = 0, world = 1, version = 0;
  ncclComm_t comm = nullptr;
  cudaStream_t stream = nullptr;            // collectives run here
  cudaEvent_t ready = nullptr;              // ctx stream -> comm stream
  cudaEvent_t done = nullptr;               // comm stream -> ctx stream (pg_dp_wait)
  bool outstanding = false;
  // bucketed reduction of one gradient arena
  pg_arena* grads = nullptr;
  int64_t bucket_bytes = 0, hi = 0, lo_ready = 0;
  std::vector<Bucket> buckets;
  std::vector<cudaEvent_t> event_pool;
  int error = PG_OK;                        // first error raised inside the on_ready callback
};

namespace {

int order_after_ctx(pg_dp* dp) {
  PG_CHECK_CUDA(cudaEventRecord(dp->ready, dp->ctx->stream));
  PG_CHECK_CUDA(cudaStreamWaitEvent(dp->stream, dp->ready, 0));
  return PG_OK;
}

int issue_bucket(pg_dp* dp, int64_t lo, int64_t hi) {
  const int dt = pg_arena_dtype(dp->grads);
  const size_t es = dt == PG_F64 ? 8 : dt == PG_F16 ? 2 : 4;
  int rc = order_after_ctx(dp);
  if (rc != PG_OK) return rc;
  char* p = (char*)pg_arena_base(dp->grads) + (size_t)lo * es;
  PG_CHECK_NCCL(g_nccl.AllReduce(p, p, (size_t)(hi - lo), nccl_type(dt), ncclSum, dp->comm, dp->stream));
  Bucket b{lo, hi - lo, nullptr};
  const size_t k = dp->buckets.size();
  if (k >= dp->event_pool.size()) {
    cudaEvent_t e;
    PG_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    dp->event_pool.push_back(e);
  }
  b.done = dp->event_pool[k];
  PG_CHECK_CUDA(cudaEventRecord(b.done, dp->stream));
  dp->buckets.push_back(b);
  return PG_OK;
}

}  // namespace

extern "C" {

int pg_dp_unique_id(void* id128) {
  PG_REQUIRE(id128 != nullptr, "null id buffer");
  int rc = nccl_load();
  if (rc != PG_OK) return rc;
  ncclUniqueId id;
  PG_CHECK_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return PG_OK;
}

int pg_dp_init(pg_ctx* ctx, int rank, int world, const void* id128, int max_ctas, pg_dp** out) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(out != nullptr && id128 != nullptr, "null pointer");
  *out = nullptr;
  PG_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank / world out of range");
  int rc = nccl_load();
  if (rc != PG_OK) return rc;
  PG_CHECK_CUDA(cudaSetDevice(ctx->device));
  pg_dp* dp = new pg_dp();
  dp->ctx = ctx; dp->rank = rank; dp->world = world;
  if (g_nccl.GetVersion) g_nccl.GetVersion(&dp->version);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclResult_t r;
  if (g_nccl.CommInitRankConfig && max_ctas > 0) {
    ncclConfig_v22700 cfg;
    cfg.size = sizeof(cfg); cfg.magic = 0xcafebeef; cfg.version = 22703;
    cfg.blocking = INT_MIN; cfg.cgaClusterSize = INT_MIN; cfg.minCTAs = INT_MIN; cfg.maxCTAs = max_ctas;
    cfg.netName = nullptr; cfg.splitShare = INT_MIN; cfg.trafficClass = INT_MIN; cfg.commName = nullptr;
    cfg.collnetEnable = INT_MIN; cfg.CTAPolicy = INT_MIN; cfg.shrinkShare = INT_MIN; cfg.nvlsCTAs = INT_MIN;
    r = g_nccl.CommInitRankConfig(&dp->comm, world, id, rank, &cfg);
  } else {
    r = g_nccl.CommInitRank(&dp->comm, world, id, rank);
  }
  if (r != 0) {
    pg_set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    delete dp;
    return PG_ERR_CUDA;
  }
  PG_CHECK_CUDA(cudaStreamCreateWithFlags(&dp->stream, cudaStreamNonBlocking));
  PG_CHECK_CUDA(cudaEventCreateWithFlags(&dp->ready, cudaEventDisableTiming));
  PG_CHECK_CUDA(cudaEventCreateWithFlags(&dp->done, cudaEventDisableTiming));
  *out = dp;
  return PG_OK;
}

int pg_dp_destroy(pg_dp* dp) {
  if (!dp) return PG_OK;
  cudaSetDevice(dp->ctx->device);
  if (dp->stream) cudaStreamSynchronize(dp->stream);
  if (dp->comm) g_nccl.CommDestroy(dp->comm);
  for (cudaEvent_t e : dp->event_pool) cudaEventDestroy(e);
  if (dp->ready) cudaEventDestroy(dp->ready);
  if (dp->done) cudaEventDestroy(dp->done);
  if (dp->stream) cudaStreamDestroy(dp->stream);
  delete dp;
  return PG_OK;
}

int pg_dp_info(pg_dp* dp, int* rank, int* world, int* nccl_version, int* p2p) {
  PG_REQUIRE(dp != nullptr, "null dp handle");
  if (rank) *rank = dp->rank;
  if (world) *world = dp->world;
  if (nccl_version) *nccl_version = dp->version;
  if (p2p) *p2p = 0;
  return PG_OK;
}

int pg_dp_allreduce(pg_dp* dp, void* buf, int64_t count, int dtype) {
  PG_REQUIRE(dp != nullptr && buf != nullptr, "null pointer");
  PG_REQUIRE(nccl_type(dtype) >= 0, "unsupported dtype");
  if (count <= 0) return PG_OK;
  int rc = order_after_ctx(dp);
  if (rc != PG_OK) return rc;
  PG_CHECK_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, nccl_type(dtype), ncclSum, dp->comm, dp->stream));
  dp->outstanding = true;
  return PG_OK;
}

int pg_dp_broadcast(pg_dp* dp, void* buf, int64_t count, int dtype, int root) {
  PG_REQUIRE(dp != nullptr && buf != nullptr, "null pointer");
  PG_REQUIRE(nccl_type(dtype) >= 0 && root >= 0 && root < dp->world, "unsupported dtype / bad root");
  if (count <= 0) return PG_OK;
  int rc = order_after_ctx(dp);
  if (rc != PG_OK) return rc;
  PG_CHECK_NCCL(g_nccl.Broadcast(buf, buf, (size_t)count, nccl_type(dtype), root, dp->comm, dp->stream));
  dp->outstanding = true;
  return PG_OK;
}

int pg_dp_wait(pg_dp* dp) {
  PG_REQUIRE(dp != nullptr, "null dp handle");
  if (!dp->outstanding) return PG_OK;
  PG_CHECK_CUDA(cudaEventRecord(dp->done, dp->stream));
  PG_CHECK_CUDA(cudaStreamWaitEvent(dp->ctx->stream, dp->done, 0));
  dp->outstanding = false;
  return PG_OK;
}

int pg_dp_bucket_begin(pg_dp* dp, pg_arena* grads, int64_t bucket_bytes) {
  PG_REQUIRE(dp != nullptr && grads != nullptr, "null pointer");
  dp->grads = grads;
  dp->bucket_bytes = bucket_bytes > 0 ? bucket_bytes : (24 << 20);
  dp->hi = dp->lo_ready = pg_arena_padded(grads);
  dp->buckets.clear();
  dp->error = PG_OK;
  return PG_OK;
}

void pg_dp_bucket_ready(void* user, int first_grad_leaf) {
  pg_dp* dp = (pg_dp*)user;
  if (!dp || !dp->grads || dp->error != PG_OK) return;
  const std::vector<int64_t>& off = pg_arena_offsets(dp->grads);
  if (first_grad_leaf < 0 || first_grad_leaf >= (int)off.size()) { dp->error = PG_ERR_ARG; pg_set_error("pg_dp_bucket_ready: leaf out of range"); return; }
  const int64_t lo = off[first_grad_leaf];
  if (lo > dp->lo_ready) { dp->error = PG_ERR_ARG; pg_set_error("pg_dp_bucket_ready: gradient leaves must complete last-to-first"); return; }
  dp->lo_ready = lo;
  const int dt = pg_arena_dtype(dp->grads);
  const int64_t es = dt == PG_F64 ? 8 : dt == PG_F16 ? 2 : 4;
  if ((dp->hi - lo) * es >= dp->bucket_bytes && lo > 0) {
    dp->error = issue_bucket(dp, lo, dp->hi);
    dp->hi = lo;
  }
}

int pg_dp_bucket_finish(pg_dp* dp, int* n_buckets, int64_t* starts, int64_t* counts, int max_buckets) {
  PG_REQUIRE(dp != nullptr && dp->grads != nullptr, "pg_dp_bucket_finish without pg_dp_bucket_begin");
  if (dp->error != PG_OK) return dp->error;
  if (dp->hi > 0) {
    int rc = issue_bucket(dp, 0, dp->hi);
    if (rc != PG_OK) return rc;
    dp->hi = 0;
  }
  const int n = (int)dp->buckets.size();
  if (n_buckets) *n_buckets = n;
  for (int k = 0; k < n && k < max_buckets; ++k) {
    if (starts) starts[k] = dp->buckets[k].start;
    if (counts) counts[k] = dp->buckets[k].count;
  }
  return PG_OK;
}

// Host-only restatement of the bucket policy above (no GPU, no communicator): given the leaf sizes of an arena and the
// sequence of first-finished leaf indices that pg_plan_backward reports, the (start, count) slices that get reduced, in
// issue order.  The multi-process CPU tests exchange exactly these slices over gloo.
int pg_dp_bucket_plan(int dtype, int n_leaves, const int64_t* leaf_sizes, int64_t bucket_bytes, const int* ready_leaves, int n_ready,
                      int64_t* starts, int64_t* counts, int max_buckets, int* n_buckets) {
  PG_REQUIRE(leaf_sizes != nullptr && n_leaves > 0 && n_buckets != nullptr, "bad arguments");
  std::vector<int64_t> off;
  int64_t total = 0;
  for (int l = 0; l < n_leaves; ++l) { off.push_back(total); total += (leaf_sizes[l] + 31) / 32 * 32; }
  const int64_t es = dtype == PG_F64 ? 8 : dtype == PG_F16 ? 2 : 4;
  if (bucket_bytes <= 0) bucket_bytes = 24 << 20;
  int64_t hi = total, lo_ready = total;
  int n = 0;
  auto emit = [&](int64_t lo, int64_t h) {
    if (n < max_buckets) { if (starts) starts[n] = lo; if (counts) counts[n] = h - lo; }
    ++n;
  };
  for (int k = 0; k < n_ready; ++k) {
    PG_REQUIRE(ready_leaves[k] >= 0 && ready_leaves[k] < n_leaves, "leaf out of range");
    const int64_t lo = off[ready_leaves[k]];
    PG_REQUIRE(lo <= lo_ready, "gradient leaves must complete last-to-first");
    lo_ready = lo;
    if ((hi - lo) * es >= bucket_bytes && lo > 0) { emit(lo, hi); hi = lo; }
  }
  if (hi > 0) emit(0, hi);
  *n_buckets = n;
  return PG_OK;
}

int pg_dp_bucket_wait(pg_dp* dp, int k) {
  PG_REQUIRE(dp != nullptr && k >= 0 && k < (int)dp->buckets.size(), "bucket index out of range");
  PG_CHECK_CUDA(cudaStreamWaitEvent(dp->ctx->stream, dp->buckets[k].done, 0));
  return PG_OK;
}

}  // extern "C"
