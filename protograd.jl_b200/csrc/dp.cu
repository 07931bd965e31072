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
#include "optim.cuh"
#include "tc_ptx.cuh"

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
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
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
  SYM(Broadcast, "ncclBroadcast"); SYM(GetErrorString, "ncclGetErrorString"); SYM(AllGather, "ncclAllGather");
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


// ================================================================================================================
// Fused gradient exchange + optimizer update over NVLink peer memory.
//
// ncclAllReduce + a full-arena Adam pass moves every gradient byte across NVLink twice (reduce-scatter + all-gather),
// runs ~24 CTAs with their own shared memory next to the backward GEMMs and then streams 28 B/param of optimizer
// state through HBM on every rank.  Here each rank OWNS 1/world of every bucket: one kernel
//   1. waits until every peer has finished producing the bucket (epoch flags written into this rank's memory),
//   2. reads its slice of the gradient arena from all peers (P2P loads: each gradient byte crosses NVLink once), sums in
//      rank order (deterministic, identical on every replica by construction),
//   3. applies the optimizer to its slice — optimizer state is touched for 1/world of the arena only,
//   4. stores the new fp32 parameters and their bf16 shadow into EVERY replica (P2P stores; the shadow goes to the half
//      of a double buffer that the step in flight does not read, so a peer's backward GEMMs are never raced),
//   5. signals completion to every peer and waits for theirs, so kernel completion means "my replica is up to date".
// The kernel uses no shared memory and few registers: it co-resides with the persistent GEMM CTAs instead of
// taking SMs away from them.
// ================================================================================================================
constexpr int MAXR = 8, SLOTS = 16;

struct P2PArgs {
  const float* grad[MAXR];
  float* param[MAXR];
  __nv_bfloat16* shadow[MAXR];
  unsigned int* flags[MAXR];
  float* m; float* v;
  unsigned int* block_counter;
  int rank, world, slot;
  unsigned int epoch;
  long long lo, hi;
  unsigned long long* phase_ns; // diagnostics (timing): per slot, accumulated ns of [wait for peers, own slice, wait for completion] + launch count
  int debug;                   // diagnostics (PG_DP_P2P_DEBUG): 1 skip remote fp32 stores, 2 skip remote bf16 stores, 4 skip local state stores
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_sys_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void wait_epoch(const unsigned int* p, unsigned int epoch) {
  if ((int)(ld_acquire_sys(p) - epoch) >= 0) return;
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(p) - epoch) < 0) {
    __nanosleep(200);
    if (clock64() - t0 > 240000000000ll) __trap();     // ~2 min: a peer died (ranks may legitimately be seconds apart on the host side); fail loudly instead of hanging forever
  }
}

// F: per-element update functor, f(p, m, v, g) -> writes p, m, v (GD ignores m, v)
// Launched as clusters of two CTAs: a pair occupies one whole TPC, so the co-running 2-CTA GEMM kernels (which need both
// SMs of a TPC) lose whole pairs, never half of one.
// NVLink reads are latency-bound (a few microseconds per round trip): every thread keeps U float4 loads per peer in
// flight — all loads of an iteration are issued before the first one is consumed (WMAX = 2 / 4 / 8 peers: U = 4 / 2 / 2,
// i.e. 64-224 remote bytes per thread, 0.5-1.8 MB per 16-CTA launch; register budget 128 at 512 threads).
template <typename F, bool STATE, int WMAX, int U>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512) p2p_update_kernel(P2PArgs a, F f) {
  const int W = a.world, R = a.rank;
  unsigned int* mine = a.flags[R];
  unsigned long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    if (a.phase_ns != nullptr && blockIdx.x == 0) t0 = globaltimer_ns();
    if (blockIdx.x == 0) {
      __threadfence_system();                           // this rank's gradient kernels precede in stream order
      for (int p = 0; p < W; ++p) st_release_sys(a.flags[p] + R * SLOTS + 2 * a.slot, a.epoch);
    }
    for (int p = 0; p < W; ++p) wait_epoch(mine + p * SLOTS + 2 * a.slot, a.epoch);
    if (a.phase_ns != nullptr && blockIdx.x == 0) t1 = globaltimer_ns();
  }
  __syncthreads();
  const long long len = a.hi - a.lo;
  long long chunk = (len + W - 1) / W;
  chunk = (chunk + 7) / 8 * 8;          // bf16 slices stay 16-byte aligned
  const long long c0 = a.lo + (long long)R * chunk;
  long long c1 = c0 + chunk;
  if (c1 > a.hi) c1 = a.hi;
  const long long nvec = c1 > c0 ? (c1 - c0) / 4 : 0;   // lo, hi and chunk are multiples of 4
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < nvec; i0 += stride * U) {
    float4 pv[U], mv[U], vv[U], g[U][WMAX];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < nvec) {
        const long long e = c0 + 4 * i;
        pv[u] = *reinterpret_cast<const float4*>(a.param[R] + e);
        if (STATE) { mv[u] = *reinterpret_cast<const float4*>(a.m + e); vv[u] = *reinterpret_cast<const float4*>(a.v + e); }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < nvec) {
        const long long e = c0 + 4 * i;
#pragma unroll
        for (int p = 0; p < WMAX; ++p)
          if (p < W) g[u][p] = ld_sys_f4(a.grad[p] + e);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i >= nvec) continue;
      const long long e = c0 + 4 * i;
      float4 s = g[u][0];
#pragma unroll
      for (int p = 1; p < WMAX; ++p)
        if (p < W) { s.x += g[u][p].x; s.y += g[u][p].y; s.z += g[u][p].z; s.w += g[u][p].w; }      // rank order: deterministic
      float4 P = pv[u], Mv = make_float4(0, 0, 0, 0), Vv = Mv;
      if (STATE) { Mv = mv[u]; Vv = vv[u]; }
      f(P.x, Mv.x, Vv.x, s.x); f(P.y, Mv.y, Vv.y, s.y); f(P.z, Mv.z, Vv.z, s.z); f(P.w, Mv.w, Vv.w, s.w);
      if (STATE) { *reinterpret_cast<float4*>(a.m + e) = Mv; *reinterpret_cast<float4*>(a.v + e) = Vv; }
      __nv_bfloat162 pk[2];
      pk[0] = __floats2bfloat162_rn(P.x, P.y); pk[1] = __floats2bfloat162_rn(P.z, P.w);
      const uint2 sh = *reinterpret_cast<uint2*>(pk);
#pragma unroll
      for (int p = 0; p < WMAX; ++p) {
        if (p < W) {
          *reinterpret_cast<float4*>(a.param[p] + e) = P;
          if (a.shadow[p] != nullptr) *reinterpret_cast<uint2*>(a.shadow[p] + e) = sh;
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    unsigned long long t2 = 0;
    if (a.phase_ns != nullptr && blockIdx.x == 0) {     // (block 0's own share done; the last block closes the kernel)
      t2 = globaltimer_ns();
      a.phase_ns[4 * a.slot + 0] += t1 - t0;
      a.phase_ns[4 * a.slot + 1] += t2 - t1;
      a.phase_ns[4 * a.slot + 3] += 1;
    }
    if (atomicAdd(a.block_counter, 1u) == gridDim.x - 1) {
      *a.block_counter = 0u;
      __threadfence_system();
      const unsigned long long t3 = a.phase_ns != nullptr ? globaltimer_ns() : 0;
      for (int p = 0; p < W; ++p) st_release_sys(a.flags[p] + R * SLOTS + 2 * a.slot + 1, a.epoch);
      for (int p = 0; p < W; ++p) wait_epoch(mine + p * SLOTS + 2 * a.slot + 1, a.epoch);
      if (a.phase_ns != nullptr) atomicAdd(a.phase_ns + 4 * a.slot + 2, globaltimer_ns() - t3);
    }
  }
}

template <typename F, bool STATE>
void launch_p2p_kernel(int world, int grid, cudaStream_t st, const P2PArgs& a, const F& f) {
  if (world <= 2) p2p_update_kernel<F, STATE, 2, 4><<<grid, 512, 0, st>>>(a, f);
  else if (world <= 4) p2p_update_kernel<F, STATE, 4, 2><<<grid, 512, 0, st>>>(a, f);
  else p2p_update_kernel<F, STATE, 8, 2><<<grid, 512, 0, st>>>(a, f);
}

// ---- TMA (bulk-copy) variant: memory-level parallelism from shared memory instead of registers -----------------------
// One elected thread streams tiles of every operand — the W peer gradient slices over NVLink, the local parameter and
// optimizer-state slices from HBM — into a ring of shared-memory stages with cp.async.bulk (completion on an mbarrier);
// the 512 threads of the CTA consume a stage (one float4 each), write the results with plain (posted) stores and hand
// the stage back.  ~170-200 KB in flight per CTA, so a few CTAs saturate NVLink where a register-resident loop is
// latency-bound.
constexpr int P2P_TILE = 2048;                      // floats per stream per stage (8 KB); 512 threads x float4
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

// Results leave the same way they came: every thread writes its float4 (and its bf16 x 4) into an output slot in shared
// memory, and the elected thread sends the slot to every replica with cp.async.bulk stores (posted through the async
// proxy: deep queues, nothing for the SM's load/store units to wait on) — local optimizer state likewise.
template <typename F, bool STATE, int TILE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512) p2p_update_tma_kernel(P2PArgs a, F f, int stages) {
  extern __shared__ __align__(128) uint8_t p2p_smem[];
  const int W = a.world, R = a.rank;
  const int NS = W + (STATE ? 3 : 1);               // input streams per stage: W gradients, p, (m, v)
  constexpr int OUT_BYTES = TILE * 4 + TILE * 2 + (STATE ? 2 * TILE * 4 : 0);     // p fp32, p bf16, (m, v)
  const uint32_t stage_bytes = (uint32_t)NS * TILE * 4;
  const uint32_t smem0 = smem_u32(p2p_smem);
  const uint32_t out0 = smem0 + (uint32_t)stages * stage_bytes;      // two output slots
  const uint32_t bar0 = out0 + 2u * OUT_BYTES;                       // full[stages]
  uint8_t* out_gen = p2p_smem + (size_t)stages * stage_bytes;
  unsigned int* mine = a.flags[R];
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(bar0 + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (blockIdx.x == 0) {
      __threadfence_system();
      for (int p = 0; p < W; ++p) st_release_sys(a.flags[p] + R * SLOTS + 2 * a.slot, a.epoch);
    }
    for (int p = 0; p < W; ++p) wait_epoch(mine + p * SLOTS + 2 * a.slot, a.epoch);
  }
  __syncthreads();
  const long long len = a.hi - a.lo;
  long long chunk = (len + W - 1) / W;
  chunk = (chunk + 7) / 8 * 8;          // bf16 slices stay 16-byte aligned
  const long long c0 = a.lo + (long long)R * chunk;
  long long c1 = c0 + chunk;
  if (c1 > a.hi) c1 = a.hi;
  const long long n_el = c1 > c0 ? c1 - c0 : 0;
  const long long n_tiles = (n_el + TILE - 1) / TILE;
  const long long my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto tile_start = [&](long long k) { return c0 + (blockIdx.x + k * gridDim.x) * TILE; };
  auto tile_elems = [&](long long k) { const long long rem = c1 - tile_start(k); return (uint32_t)(rem < TILE ? rem : TILE); };
  auto issue = [&](long long k) {                    // elected thread: load this CTA's k-th tile into stage k % stages
    const int s = (int)(k % stages);
    const long long e = tile_start(k);
    const uint32_t bytes = tile_elems(k) * 4;
    const uint32_t bar = bar0 + 8u * s, dst = smem0 + (uint32_t)s * stage_bytes;
    mbar_expect_tx(bar, bytes * NS);
    for (int p = 0; p < W; ++p) bulk_load(dst + (uint32_t)p * TILE * 4, a.grad[p] + e, bytes, bar);
    bulk_load(dst + (uint32_t)W * TILE * 4, a.param[R] + e, bytes, bar);
    if (STATE) {
      bulk_load(dst + (uint32_t)(W + 1) * TILE * 4, a.m + e, bytes, bar);
      bulk_load(dst + (uint32_t)(W + 2) * TILE * 4, a.v + e, bytes, bar);
    }
  };
  if (threadIdx.x == 0)
    for (long long k = 0; k < my_tiles && k < stages; ++k) issue(k);
  for (long long k = 0; k < my_tiles; ++k) {
    const int s = (int)(k % stages), o = (int)(k & 1);
    // the bulk stores that last used output slot `o` (tile k-2) must have finished READING it
    if (threadIdx.x == 0 && k >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncthreads();
    mbar_wait(bar0 + 8u * s, (uint32_t)((k / stages) & 1));
    const uint32_t nel = tile_elems(k);
    uint8_t* ob = out_gen + (size_t)o * OUT_BYTES;
    if (4u * threadIdx.x < nel) {
      const float4* st = reinterpret_cast<const float4*>(p2p_smem + (size_t)s * stage_bytes) + threadIdx.x;
      float4 g = st[0];
      for (int p = 1; p < W; ++p) {                  // rank order: deterministic, identical on every replica
        const float4 t = st[p * (TILE / 4)];
        g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
      }
      float4 P = st[W * (TILE / 4)], Mv = make_float4(0, 0, 0, 0), Vv = Mv;
      if (STATE) { Mv = st[(W + 1) * (TILE / 4)]; Vv = st[(W + 2) * (TILE / 4)]; }
      f(P.x, Mv.x, Vv.x, g.x); f(P.y, Mv.y, Vv.y, g.y); f(P.z, Mv.z, Vv.z, g.z); f(P.w, Mv.w, Vv.w, g.w);
      reinterpret_cast<float4*>(ob)[threadIdx.x] = P;
      __nv_bfloat162 pk[2];
      pk[0] = __floats2bfloat162_rn(P.x, P.y); pk[1] = __floats2bfloat162_rn(P.z, P.w);
      reinterpret_cast<uint2*>(ob + TILE * 4)[threadIdx.x] = *reinterpret_cast<uint2*>(pk);
      if (STATE) {
        reinterpret_cast<float4*>(ob + TILE * 6)[threadIdx.x] = Mv;
        reinterpret_cast<float4*>(ob + TILE * 10)[threadIdx.x] = Vv;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk stores
    __syncthreads();                                 // stage s fully consumed, output slot o fully written
    if (threadIdx.x == 0) {
      const long long e = tile_start(k);
      const uint32_t so = out0 + (uint32_t)o * OUT_BYTES;
      for (int p = 0; p < W; ++p) {
        if (p == R || !(a.debug & 1)) bulk_store(a.param[p] + e, so, nel * 4);
        if (a.shadow[p] != nullptr && (p == R || !(a.debug & 2))) bulk_store(a.shadow[p] + e, so + TILE * 4, nel * 2);
      }
      if (STATE && !(a.debug & 4)) { bulk_store(a.m + e, so + TILE * 6, nel * 4); bulk_store(a.v + e, so + TILE * 10, nel * 4); }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (k + stages < my_tiles) issue(k + stages);
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // every store of this CTA has completed
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (atomicAdd(a.block_counter, 1u) == gridDim.x - 1) {
      *a.block_counter = 0u;
      __threadfence_system();
      for (int p = 0; p < W; ++p) st_release_sys(a.flags[p] + R * SLOTS + 2 * a.slot + 1, a.epoch);
      for (int p = 0; p < W; ++p) wait_epoch(mine + p * SLOTS + 2 * a.slot + 1, a.epoch);
    }
  }
}

template <typename F, bool STATE, int TILE>
int launch_p2p_tma_t(int world, int grid, cudaStream_t st, const P2PArgs& a, const F& f) {
  const int NS = world + (STATE ? 3 : 1);
  const int stage_bytes = NS * TILE * 4;
  const int out_bytes = 2 * (TILE * 4 + TILE * 2 + (STATE ? 2 * TILE * 4 : 0));
  int stages = (212 * 1024 - out_bytes) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  const int smem = stages * stage_bytes + out_bytes + 8 * stages + 64;
  static PgOnce attr_set;                      // (per instantiation)
  int dev_ = 0;
  cudaGetDevice(&dev_);
  if (attr_set.need(dev_)) {
    PG_CHECK_CUDA(cudaFuncSetAttribute(p2p_update_tma_kernel<F, STATE, TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  }
  p2p_update_tma_kernel<F, STATE, TILE><<<grid, 512, smem, st>>>(a, f, stages);
  return PG_OK;
}
template <typename F, bool STATE>
int launch_p2p_tma(int world, int grid, cudaStream_t st, const P2PArgs& a, const F& f) {
  // 8 ranks: 11 input streams per stage; 4 KB tiles keep three stages in flight inside 227 KB
  if (world > 4) return launch_p2p_tma_t<F, STATE, 1024>(world, grid, st, a, f);
  return launch_p2p_tma_t<F, STATE, 2048>(world, grid, st, a, f);
}

struct P2PAdamFast {           // AdamFastF of optim.cuh on scalars
  float s, b1, omb1, b2, omb2, rc1, rc2, eps;
  __device__ __forceinline__ void operator()(float& p, float& m, float& v, float g) const {
    m = fmaf(b1, m, omb1 * g);
    v = fmaf(b2, v, omb2 * g * g);
    p = p - s * __fdividef(m * rc1, sqrtf(v * rc2) + eps);
  }
};
struct P2PAdamExact {          // AdamF<float, double>: Float64-promoted per-element math (bit-exact with the reference's broadcast)
  double s, b1, omb1, b2, omb2, c1, c2, eps;
  __device__ __forceinline__ void operator()(float& p, float& m, float& v, float g) const {
    const float in[4] = {p, m, v, g};
    float out[5];
    AdamF<float, double>{s, b1, omb1, b2, omb2, c1, c2, eps, nullptr}(in, out);
    p = out[0]; m = out[1]; v = out[2];
  }
};
struct P2PGd {
  double s; int f64;
  __device__ __forceinline__ void operator()(float& p, float&, float&, float g) const {
    p = f64 ? (float)__dsub_rn((double)p, __dmul_rn(s, (double)g)) : __fsub_rn(p, __fmul_rn((float)s, g));
  }
};

struct Bucket { int64_t start, count; cudaEvent_t done; };

}  // namespace

struct pg_dp {
  pg_ctx* ctx = nullptr;
  int rank = 0, world = 1, version = 0;
  ncclComm_t comm = nullptr;
  cudaStream_t stream = nullptr;            // collectives run here
  cudaStream_t stream2 = nullptr;           // the final (head) bucket of a fused step: runs next to the tail bucket's kernel, not behind it
  cudaEvent_t ready = nullptr;              // ctx stream -> comm stream
  cudaEvent_t done = nullptr;               // comm stream -> ctx stream (pg_dp_wait)
  bool outstanding = false;
  // bucketed reduction of one gradient arena
  pg_arena* grads = nullptr;
  int64_t bucket_bytes = 0, hi = 0, lo_ready = 0;
  std::vector<Bucket> buckets;
  std::vector<cudaEvent_t> event_pool;
  int error = PG_OK;                        // first error raised inside the on_ready callback
  // ---- fused exchange + update over NVLink peer memory (pg_dp_p2p_*) ----
  bool p2p = false;
  void* peer_grad[8] = {nullptr};           // every rank's gradient arena, mapped here (own = local pointer)
  void* peer_param[8] = {nullptr};
  void* peer_shadow[8] = {nullptr};         // bf16 [2][n_padded] (nullable)
  void* peer_flags[8] = {nullptr};
  void* opened[32] = {nullptr}; int n_opened = 0;
  void* flags_local = nullptr;              // [8 ranks][16 slots] epochs written by the peers
  unsigned int* block_counter = nullptr;
  int64_t p2p_padded = 0;
  unsigned int epoch = 0;
  int shadow_half = 0;                      // the half the CURRENT parameters' bf16 copy lives in
  // optimizer of the step in flight (set by pg_dp_p2p_set_adam before pg_plan_backward)
  struct { int kind = -1; void* m = nullptr; void* v = nullptr; double s = 0, b1 = 0, b2 = 0, eps = 0; int64_t it = 1; int flags = 0; } opt;
  int p2p_blocks = 16, p2p_blocks_last = 96;
  bool fused_step = false;
  unsigned long long* phase_ns = nullptr;   // device, [8 slots][4]
  cudaEvent_t tev[16] = {nullptr}; double tsum[8] = {0}; long tcount[8] = {0}; int debug = 0; int timing = 0;
  int reserve_sms = 16;                     // SMs the persistent GEMM grids leave free while a fused bucket is in flight
  bool overlap_next = true;                 // PG_DP_OVERLAP_NEXT=0: every bucket is waited for before the step returns
  bool was_fused = false;                   // the buckets of the last pg_dp_bucket_finish were fused exchange kernels
  bool p2p_tma = false;                     // PG_DP_P2P_TMA=1: cp.async.bulk pipeline instead of register-resident loads (measured slower per CTA)
};

namespace {

int order_after_ctx(pg_dp* dp, cudaStream_t st = nullptr) {
  PG_CHECK_CUDA(cudaEventRecord(dp->ready, dp->ctx->stream));
  PG_CHECK_CUDA(cudaStreamWaitEvent(st ? st : dp->stream, dp->ready, 0));
  return PG_OK;
}

int launch_p2p(pg_dp* dp, int64_t lo, int64_t hi, int slot, bool last, cudaStream_t st) {
  P2PArgs a;
  memset(&a, 0, sizeof(a));
  const bool use_shadow = dp->peer_shadow[dp->rank] != nullptr;
  for (int p = 0; p < dp->world; ++p) {
    a.grad[p] = (const float*)dp->peer_grad[p];
    a.param[p] = (float*)dp->peer_param[p];
    // bf16 copies of the NEW parameters go to the half the step in flight does not read
    a.shadow[p] = use_shadow ? (__nv_bfloat16*)dp->peer_shadow[p] + (size_t)(dp->shadow_half ^ 1) * dp->p2p_padded : nullptr;
    a.flags[p] = (unsigned int*)dp->peer_flags[p];
  }
  a.m = (float*)dp->opt.m; a.v = (float*)dp->opt.v;
  a.block_counter = dp->block_counter + slot;
  a.rank = dp->rank; a.world = dp->world; a.slot = slot; a.epoch = dp->epoch;
  a.lo = lo; a.hi = hi; a.debug = dp->debug;
  if (dp->timing && !dp->phase_ns) { PG_CHECK_CUDA(cudaMalloc((void**)&dp->phase_ns, 32 * sizeof(unsigned long long))); PG_CHECK_CUDA(cudaMemset(dp->phase_ns, 0, 32 * sizeof(unsigned long long))); }
  a.phase_ns = dp->timing ? dp->phase_ns : nullptr;
  if (dp->timing && slot < 8) {
    for (int k = 0; k < 2; ++k)
      if (!dp->tev[2 * slot + k]) PG_CHECK_CUDA(cudaEventCreate(&dp->tev[2 * slot + k]));
    if (dp->tcount[slot] >= 0 && cudaEventQuery(dp->tev[2 * slot + 1]) == cudaSuccess && dp->epoch > 1) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, dp->tev[2 * slot], dp->tev[2 * slot + 1]) == cudaSuccess) { dp->tsum[slot] += ms; dp->tcount[slot]++; }
    }
    cudaGetLastError();
    PG_CHECK_CUDA(cudaEventRecord(dp->tev[2 * slot], st));
  }
  const double s = dp->opt.s, b1 = dp->opt.b1, b2 = dp->opt.b2;
  // buckets that overlap the remaining backward GEMMs run on a few TPCs; the last one has the GPU to itself
  int grid = last ? dp->p2p_blocks_last : dp->p2p_blocks;
  grid = (grid + 1) / 2 * 2;
  if (dp->opt.kind == 0) {
    const double c1 = 1.0 - pow(b1, (double)dp->opt.it), c2 = 1.0 - pow(b2, (double)dp->opt.it);
    if (dp->opt.flags & PG_MATH_FAST) {
      P2PAdamFast f{(float)s, (float)b1, (float)(1.0 - b1), (float)b2, (float)(1.0 - b2), (float)(1.0 / c1), (float)(1.0 / c2), (float)dp->opt.eps};
      if (dp->p2p_tma) { int rc = launch_p2p_tma<P2PAdamFast, true>(dp->world, grid, st, a, f); if (rc != PG_OK) return rc; }
      else launch_p2p_kernel<P2PAdamFast, true>(dp->world, grid, st, a, f);
    } else {
      P2PAdamExact f{s, b1, 1.0 - b1, b2, 1.0 - b2, c1, c2, dp->opt.eps};
      if (dp->p2p_tma) { int rc = launch_p2p_tma<P2PAdamExact, true>(dp->world, grid, st, a, f); if (rc != PG_OK) return rc; }
      else launch_p2p_kernel<P2PAdamExact, true>(dp->world, grid, st, a, f);
    }
  } else {
    P2PGd f{s, (dp->opt.flags & PG_SCALAR_F64) ? 1 : 0};
    if (dp->p2p_tma) { int rc = launch_p2p_tma<P2PGd, false>(dp->world, grid, st, a, f); if (rc != PG_OK) return rc; }
    else launch_p2p_kernel<P2PGd, false>(dp->world, grid, st, a, f);
  }
  dp->ctx->launches++;
  PG_CHECK_CUDA(cudaGetLastError());
  if (dp->timing && slot < 8) PG_CHECK_CUDA(cudaEventRecord(dp->tev[2 * slot + 1], st));
  return PG_OK;
}

int issue_bucket(pg_dp* dp, int64_t lo, int64_t hi) {
  const int dt = pg_arena_dtype(dp->grads);
  const size_t es = dt == PG_F64 ? 8 : dt == PG_F16 ? 2 : 4;
  const bool fused = dp->p2p && dp->opt.kind >= 0;
  // the final (head) bucket of a fused step gets its own stream: it runs NEXT TO the tail bucket's kernel instead of
  // behind it, and the next step's first layers only wait for it (the tail bucket's wait is deferred, see below)
  cudaStream_t st = (fused && lo == 0 && !dp->buckets.empty() && dp->overlap_next) ? dp->stream2 : dp->stream;
  int rc = order_after_ctx(dp, st);
  if (rc != PG_OK) return rc;
  char* p = (char*)pg_arena_base(dp->grads) + (size_t)lo * es;
  if (fused) {
    PG_REQUIRE(dp->buckets.size() < SLOTS / 2, "too many buckets for the fused exchange (8)");
    PG_REQUIRE(pg_arena_base(dp->grads) == dp->peer_grad[dp->rank], "fused exchange: not the gradient arena that was registered");
    rc = launch_p2p(dp, lo, hi, (int)dp->buckets.size(), lo == 0, st);
    if (rc != PG_OK) return rc;
    // The backward kernels enqueued from here on run next to this bucket's kernel: their persistent grids (one CTA per
    // SM, all of its shared memory) leave `reserve_sms` SMs — whole TPCs — free, so the exchange starts at once instead of
    // waiting for a GEMM CTA to retire.  Restored by pg_dp_bucket_finish.
    if (lo > 0 && dp->reserve_sms > 0) dp->ctx->num_sms = dp->ctx->real_sms - dp->reserve_sms;
  } else {
    PG_CHECK_NCCL(g_nccl.AllReduce(p, p, (size_t)(hi - lo), nccl_type(dt), ncclSum, dp->comm, dp->stream));
  }
  Bucket b{lo, hi - lo, nullptr};
  const size_t k = dp->buckets.size();
  if (k >= dp->event_pool.size()) {
    cudaEvent_t e;
    PG_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    dp->event_pool.push_back(e);
  }
  b.done = dp->event_pool[k];
  PG_CHECK_CUDA(cudaEventRecord(b.done, st));
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
  PG_CHECK_CUDA(cudaStreamCreateWithFlags(&dp->stream2, cudaStreamNonBlocking));
  PG_CHECK_CUDA(cudaEventCreateWithFlags(&dp->ready, cudaEventDisableTiming));
  PG_CHECK_CUDA(cudaEventCreateWithFlags(&dp->done, cudaEventDisableTiming));
  *out = dp;
  return PG_OK;
}

int pg_dp_destroy(pg_dp* dp) {
  if (!dp) return PG_OK;
  cudaSetDevice(dp->ctx->device);
  if (dp->timing && dp->rank == 0)
    for (int k = 0; k < 8; ++k)
      if (dp->tcount[k] > 0) fprintf(stderr, "[pg_dp] fused exchange kernel, bucket %d: %.1f us average over %ld launches (includes waiting for the peers)\n", k, 1e3 * dp->tsum[k] / dp->tcount[k], dp->tcount[k]);
  if (dp->timing && dp->phase_ns && dp->rank == 0) {
    unsigned long long h[32];
    if (cudaMemcpy(h, dp->phase_ns, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess)
      for (int k = 0; k < 8; ++k)
        if (h[4 * k + 3] > 0)
          fprintf(stderr, "[pg_dp]   bucket %d phases (rank 0, us, mean of %llu): wait for every peer's gradients %.1f | own slice (CTA 0) %.1f | wait for every peer's stores %.1f\n",
                  k, h[4 * k + 3], 1e-3 * h[4 * k] / h[4 * k + 3], 1e-3 * h[4 * k + 1] / h[4 * k + 3], 1e-3 * h[4 * k + 2] / h[4 * k + 3]);
  }
  if (dp->stream) cudaStreamSynchronize(dp->stream);
  if (dp->comm) g_nccl.CommDestroy(dp->comm);
  for (int k = 0; k < dp->n_opened; ++k) cudaIpcCloseMemHandle(dp->opened[k]);
  if (dp->flags_local) cudaFree(dp->flags_local);
  if (dp->block_counter) cudaFree(dp->block_counter);
  for (cudaEvent_t e : dp->event_pool) cudaEventDestroy(e);
  if (dp->ready) cudaEventDestroy(dp->ready);
  if (dp->done) cudaEventDestroy(dp->done);
  if (dp->stream) cudaStreamDestroy(dp->stream);
  if (dp->stream2) { cudaStreamSynchronize(dp->stream2); cudaStreamDestroy(dp->stream2); }
  delete dp;
  return PG_OK;
}

int pg_dp_info(pg_dp* dp, int* rank, int* world, int* nccl_version, int* p2p) {
  PG_REQUIRE(dp != nullptr, "null dp handle");
  if (rank) *rank = dp->rank;
  if (world) *world = dp->world;
  if (nccl_version) *nccl_version = dp->version;
  if (p2p) *p2p = dp->p2p ? 1 : 0;
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
  pg_flush_deferred(dp->ctx);                 // (also a deferred bucket wait of the last fused step)
  if (!dp->outstanding) return PG_OK;
  PG_CHECK_CUDA(cudaEventRecord(dp->done, dp->stream));
  PG_CHECK_CUDA(cudaStreamWaitEvent(dp->ctx->stream, dp->done, 0));
  dp->outstanding = false;
  return PG_OK;
}


// ---- fused exchange + update over peer memory: setup -------------------------------------------------------------
static int ipc_export(void* ptr, cudaIpcMemHandle_t* h, int64_t* offset) {
  // the handle names the whole allocation: find its base to learn the offset of `ptr` inside it
  typedef int (*GetRange)(unsigned long long*, size_t*, unsigned long long);
  static GetRange get_range = nullptr;
  if (!get_range) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    PG_CHECK_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q));
    get_range = (GetRange)fn;
  }
  unsigned long long base = 0; size_t size = 0;
  if (!get_range || get_range(&base, &size, (unsigned long long)ptr) != 0) { pg_set_error("cuMemGetAddressRange failed"); return PG_ERR_CUDA; }
  PG_CHECK_CUDA(cudaIpcGetMemHandle(h, (void*)base));
  *offset = (int64_t)((unsigned long long)ptr - base);
  return PG_OK;
}

int pg_dp_p2p_enable(pg_dp* dp, pg_arena* params, pg_arena* grads, void* shadow2) {
  PG_REQUIRE(dp != nullptr && params != nullptr && grads != nullptr, "null pointer");
  PG_REQUIRE(dp->world <= MAXR, "the fused exchange supports up to 8 ranks (one NVSwitch domain)");
  PG_REQUIRE(pg_arena_dtype(params) == PG_F32 && pg_arena_dtype(grads) == PG_F32, "the fused exchange needs Float32 arenas");
  PG_REQUIRE(pg_arena_padded(params) == pg_arena_padded(grads), "parameter and gradient arenas differ in layout");
  PG_REQUIRE(g_nccl.AllGather != nullptr, "libnccl has no ncclAllGather");
  if (dp->p2p) return PG_OK;
  PG_CHECK_CUDA(cudaSetDevice(dp->ctx->device));
  const int W = dp->world, R = dp->rank;
  dp->p2p_padded = pg_arena_padded(params);
  // flag block: peers write their epochs here
  PG_CHECK_CUDA(cudaMalloc(&dp->flags_local, MAXR * SLOTS * sizeof(unsigned int)));
  PG_CHECK_CUDA(cudaMemset(dp->flags_local, 0, MAXR * SLOTS * sizeof(unsigned int)));
  PG_CHECK_CUDA(cudaMalloc((void**)&dp->block_counter, SLOTS * sizeof(unsigned int)));       // one per bucket slot: two buckets' kernels may overlap
  PG_CHECK_CUDA(cudaMemset(dp->block_counter, 0, SLOTS * sizeof(unsigned int)));
  struct Rec { cudaIpcMemHandle_t h[4]; int64_t off[4]; int has_shadow; int pad; };
  Rec mine;
  memset(&mine, 0, sizeof(mine));
  void* ptrs[4] = {pg_arena_base(grads), pg_arena_base(params), shadow2, dp->flags_local};
  mine.has_shadow = shadow2 != nullptr;
  for (int k = 0; k < 4; ++k) {
    if (!ptrs[k]) continue;
    int rc = ipc_export(ptrs[k], &mine.h[k], &mine.off[k]);
    if (rc != PG_OK) return rc;
  }
  // exchange the records through the communicator itself
  Rec* dev = nullptr;
  PG_CHECK_CUDA(cudaMalloc((void**)&dev, sizeof(Rec) * (W + 1)));
  PG_CHECK_CUDA(cudaMemcpy(dev + W, &mine, sizeof(Rec), cudaMemcpyHostToDevice));
  PG_CHECK_NCCL(g_nccl.AllGather(dev + W, dev, sizeof(Rec), 0 /* ncclInt8 */, dp->comm, dp->stream));
  PG_CHECK_CUDA(cudaStreamSynchronize(dp->stream));
  std::vector<Rec> all(W);
  PG_CHECK_CUDA(cudaMemcpy(all.data(), dev, sizeof(Rec) * W, cudaMemcpyDeviceToHost));
  PG_CHECK_CUDA(cudaFree(dev));
  void** tables[4] = {dp->peer_grad, dp->peer_param, dp->peer_shadow, dp->peer_flags};
  for (int p = 0; p < W; ++p) {
    PG_REQUIRE(all[p].has_shadow == mine.has_shadow, "ranks disagree about the bf16 shadow");
    for (int k = 0; k < 4; ++k) {
      if (!ptrs[k]) { tables[k][p] = nullptr; continue; }
      if (p == R) { tables[k][p] = ptrs[k]; continue; }
      void* base = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&base, all[p].h[k], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        pg_set_error("cudaIpcOpenMemHandle(rank %d, buffer %d) failed: %s — peer memory is not available; the NCCL all-reduce path stays in use", p, k, cudaGetErrorString(e));
        cudaGetLastError();
        return PG_ERR_UNSUPPORTED;
      }
      dp->opened[dp->n_opened++] = base;
      tables[k][p] = (char*)base + all[p].off[k];
    }
  }
  // measured at 2 ranks: the kernel moves ~76 GB/s per CTA; 2 ranks own half the arena each, so
  // they need more CTAs than 8 ranks (an eighth each) to finish inside the remaining backward GEMMs
  dp->p2p_blocks = dp->reserve_sms = W <= 2 ? 32 : 24;
  const char* nb = getenv("PG_DP_P2P_BLOCKS");
  if (nb && atoi(nb) > 0) dp->p2p_blocks = atoi(nb);
  nb = getenv("PG_DP_P2P_DEBUG");
  if (nb) dp->debug = atoi(nb);
  nb = getenv("PG_DP_P2P_TIMING");
  if (nb && nb[0] == '1') dp->timing = 1;
  nb = getenv("PG_DP_OVERLAP_NEXT");
  if (nb && nb[0] == '0') dp->overlap_next = false;
  nb = getenv("PG_DP_RESERVE_SMS");
  if (nb) dp->reserve_sms = atoi(nb);
  nb = getenv("PG_DP_P2P_TMA");
  if (nb && nb[0] == '0') dp->p2p_tma = false;
  nb = getenv("PG_DP_P2P_BLOCKS_LAST");
  if (nb && atoi(nb) > 0) dp->p2p_blocks_last = atoi(nb);
  dp->p2p = true;
  return PG_OK;
}

/* optimizer of the step whose backward pass comes next: kind 0 = Adam (adam.jl:32-41; `it` before the increment; flags
 * PG_MATH_FAST / PG_SCALAR_F64 as in pg_opt_adam; m, v = this rank's state arenas, of which only its own slices are
 * touched), kind 2 = GradientDescent (gradient_descent.jl:15-17).  kind -1 switches back to plain all-reduce buckets. */
int pg_dp_p2p_set_optimizer(pg_dp* dp, int kind, void* m, void* v, double stepsize, double beta1, double beta2, double epsilon, int64_t it, int flags) {
  PG_REQUIRE(dp != nullptr, "null dp handle");
  PG_REQUIRE(kind == -1 || dp->p2p, "pg_dp_p2p_enable first");
  PG_REQUIRE(kind == -1 || kind == 0 || kind == 2, "kind: 0 Adam, 2 GradientDescent, -1 off");
  PG_REQUIRE(kind != 0 || (m != nullptr && v != nullptr && it >= 1), "Adam needs its state arenas and it >= 1");
  dp->opt.kind = kind; dp->opt.m = m; dp->opt.v = v; dp->opt.s = stepsize; dp->opt.b1 = beta1; dp->opt.b2 = beta2; dp->opt.eps = epsilon;
  dp->opt.it = it; dp->opt.flags = flags;
  return PG_OK;
}

int pg_dp_set_option(pg_dp* dp, const char* name, int value) {
  PG_REQUIRE(dp != nullptr && name != nullptr, "null pointer");
  if (!strcmp(name, "p2p_blocks")) dp->p2p_blocks = value;
  else if (!strcmp(name, "p2p_blocks_last")) dp->p2p_blocks_last = value;
  else if (!strcmp(name, "p2p_tma")) dp->p2p_tma = value != 0;
  else if (!strcmp(name, "reserve_sms")) dp->reserve_sms = value;
  else if (!strcmp(name, "overlap_next")) dp->overlap_next = value != 0;
  else if (!strcmp(name, "debug")) dp->debug = value;
  else if (!strcmp(name, "timing")) dp->timing = value;
  else { pg_set_error("pg_dp_set_option: unknown option %s", name); return PG_ERR_ARG; }
  return PG_OK;
}

/* after pg_dp_bucket_finish + waits of a fused step: which half of the shadow double buffer now holds the bf16 copy of
 * the (new) parameters */
int pg_dp_p2p_shadow_half(pg_dp* dp, int* half) {
  PG_REQUIRE(dp != nullptr && half != nullptr, "null pointer");
  *half = dp->shadow_half;
  return PG_OK;
}

int pg_dp_bucket_begin(pg_dp* dp, pg_arena* grads, int64_t bucket_bytes) {
  PG_REQUIRE(dp != nullptr && grads != nullptr, "null pointer");
  dp->grads = grads;
  dp->bucket_bytes = bucket_bytes > 0 ? bucket_bytes : (24 << 20);
  dp->hi = dp->lo_ready = pg_arena_padded(grads);
  dp->buckets.clear();
  dp->error = PG_OK;
  dp->fused_step = dp->p2p && dp->opt.kind >= 0;
  if (dp->fused_step) dp->epoch += 1;
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
  dp->ctx->num_sms = dp->ctx->real_sms;
  dp->was_fused = dp->fused_step;
  if (dp->fused_step) {                 // the fused kernels wrote the other half of the shadow double buffer
    if (dp->peer_shadow[dp->rank] != nullptr) dp->shadow_half ^= 1;
    dp->fused_step = false;
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
  pg_ctx* ctx = dp->ctx;
  // Fused step with several buckets: the FIRST (tail) bucket holds the parameters of the later layers.  Its wait is
  // deferred: the compute stream picks it up in front of the first kernel that touches those parameters (pg_plan_forward)
  // — or at the next API call that launches anything else — so the next step's first layers overlap the rest of the exchange.
  if (dp->was_fused && dp->overlap_next && k == 0 && dp->buckets.size() > 1 && !ctx->capturing && dp->p2p_padded > 0) {
    if (ctx->deferred_armed) pg_flush_deferred(ctx);
    const char* base = (const char*)dp->peer_param[dp->rank];
    ctx->deferred_event = dp->buckets[0].done;
    ctx->deferred_lo = base + (size_t)dp->buckets[0].start * 4;
    ctx->deferred_hi = base + (size_t)(dp->buckets[0].start + dp->buckets[0].count) * 4;
    ctx->deferred_armed = true;
    return PG_OK;
  }
  PG_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, dp->buckets[k].done, 0));
  return PG_OK;
}

}  // extern "C"
