// common.cuh — context, error convention, small device helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/protograd_b200.h"

struct pg_ctx {
  int device = 0;
  int num_sms = 148;       // SMs used to size grids (pg_set_sm_limit)
  int real_sms = 148;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // H2D prefetch of the next minibatch
  cudaEvent_t switch_event = nullptr;   // orders a newly borrowed stream behind the previous one (pg_set_stream)
  int64_t launches = 0;
  // scratch: split-K partials, block partials for reductions
  void* ws = nullptr;
  size_t ws_bytes = 0;
  // padded bf16 operand copies of narrow (N <= 32) tensor-core layers
  void* aux = nullptr;
  size_t aux_bytes = 0;
  // small persistent scratch (256 words, zeroed): [0..63] reduction / tile-scheduler counters, [64..255] K-split tail flags of the pair GEMM
  unsigned int* counters = nullptr;
  double* result_slot = nullptr;     // device
  double* result_host = nullptr;     // pinned
  bool capturing = false;
  // driver entry point for TMA descriptor encoding (resolved lazily, no libcuda link dependency)
  void* encode_tiled = nullptr;
  // A deferred stream dependency (data-parallel exchange, dp.cu): the tail bucket's fused exchange + update may still be
  // running when the next step's forward starts; the compute stream only has to wait for it before the first kernel
  // that reads parameters in [deferred_lo, deferred_hi).  Every API call that launches work flushes the wait
  // (PG_CHECK_CTX) unless the planner holds it (deferred_hold) while it runs the layers in front of that point.
  cudaEvent_t deferred_event = nullptr;
  const char* deferred_lo = nullptr;
  const char* deferred_hi = nullptr;
  bool deferred_armed = false, deferred_hold = false;
};

// make the compute stream wait for the deferred event now
static inline void pg_flush_deferred(pg_ctx* ctx) {
  if (ctx->deferred_armed) {
    cudaStreamWaitEvent(ctx->stream, ctx->deferred_event, 0);
    ctx->deferred_armed = false;
  }
}
static inline bool pg_deferred_overlaps(const pg_ctx* ctx, const void* p, size_t bytes) {
  return ctx->deferred_armed && p != nullptr && (const char*)p < ctx->deferred_hi && (const char*)p + bytes > ctx->deferred_lo;
}

void pg_set_error(const char* fmt, ...);

// cudaFuncSetAttribute is a per-device setting: a function-static PgOnce remembers it per (call site, device)
struct PgOnce {
  bool done[32] = {};
  bool need(int dev) {
    if (dev < 0 || dev >= 32) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

#define PG_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      pg_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PG_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define PG_REQUIRE(cond, msg)                                     \
  do {                                                            \
    if (!(cond)) {                                                \
      pg_set_error("%s: %s (%s:%d)", __func__, msg, __FILE__, __LINE__); \
      return PG_ERR_ARG;                                          \
    }                                                             \
  } while (0)

// entry-point prologue: also orders the compute stream behind a deferred exchange (see pg_ctx::deferred_*)
#define PG_CHECK_CTX(ctx)                                                      \
  do {                                                                         \
    PG_REQUIRE((ctx) != nullptr, "null context");                              \
    if ((ctx)->deferred_armed && !(ctx)->deferred_hold) pg_flush_deferred(ctx); \
  } while (0)
// plumbing calls that touch no model memory (events, copies of other buffers): no flush
#define PG_CHECK_CTX_NF(ctx) PG_REQUIRE((ctx) != nullptr, "null context")

// call after every kernel launch
#define PG_LAUNCHED(ctx)                         \
  do {                                           \
    (ctx)->launches++;                           \
    PG_CHECK_CUDA(cudaGetLastError());           \
  } while (0)

int pg_ws_ensure(pg_ctx* ctx, size_t bytes);  // grow scratch (fails during graph capture)

// column sums out[n] = sum_m X[m][n]; in_dtype f32/f64 (out same type) or bf16 (out f32)
int pg_colsum_launch(pg_ctx* ctx, int in_dtype, const void* X, void* out, int64_t M, int64_t N);

// shared-memory-tiled direct convolutions (conv_tiled.cu): 1 = handled, 0 = shape outside envelope, <0 = error
template <typename T> int pg_conv_fwd_tiled(pg_ctx* ctx, const T* x, const T* w, const T* b, T* y, int64_t N, int Cin, int H, int W, int Cout, int kH, int kW, int act);
template <typename T> int pg_conv_dgrad_tiled(pg_ctx* ctx, const T* dy, const T* w, T* dx, int64_t N, int Cin, int H, int W, int Cout, int kH, int kW);
template <typename T> int pg_conv_wgrad_tiled(pg_ctx* ctx, const T* x, const T* dy, T* dw, T* db, int64_t N, int Cin, int H, int W, int Cout, int kH, int kW);

// Float16 arenas (arena.cu): the fixed-function algebra entry points lowered onto the expression interpreter
int pg_f16_vec_binary(pg_ctx* ctx, int op, void* out, const void* x, const void* y, int64_t n);
int pg_f16_vec_scalar(pg_ctx* ctx, int op, void* out, const void* x, double c, int flags, int64_t n);
int pg_f16_vec_axpy(pg_ctx* ctx, void* out, const void* x, double s, const void* g, int sign, int flags, int64_t n);
int pg_f16_vec_fill(pg_ctx* ctx, void* out, double c, int64_t n);
int pg_f16_vec_dot(pg_ctx* ctx, const void* x, const void* y, int64_t n, double* result);

static inline int64_t pg_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers ------------------------------------------------------------------------
template <typename T> struct Eps;
template <> struct Eps<float> { static __device__ __forceinline__ float v() { return 1.1920929e-07f; } };
template <> struct Eps<double> { static __device__ __forceinline__ double v() { return 2.220446049250313e-16; } };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum (blockDim.x multiple of 32, <= 1024); result valid in thread 0.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem /* >= 32 */) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) smem[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem[threadIdx.x] : T(0);
  if (w == 0) v = warp_sum(v);
  __syncthreads();
  return v;
}

template <typename T> __device__ __forceinline__ float to_f32(T v) { return (float)v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
