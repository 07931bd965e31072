// loss_kernels.cu — softmax (NNlib, src/layers.jl:87), cross_entropy on probabilities with the
// eps INSIDE the log (src/losses.jl:7-10), mse (src/losses.jl:5), class_error (:12-16), fused so
// that logits are read once and dZ written once: P = softmax(Z); L = mean_b(-sum_c Y log(P+eps));
// dP = -Y/(B (P+eps)); dZ = P .* (dP - sum_c dP.*P).  One sub-warp group of G lanes per sample,
// shuffle reductions inside the group; deterministic last-block reduction for the loss scalar.
#include "common.cuh"

namespace {

template <typename T> __device__ __forceinline__ T t_exp(T x);
template <> __device__ __forceinline__ float t_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double t_exp<double>(double x) { return exp(x); }
template <typename T> __device__ __forceinline__ T t_log(T x);
template <> __device__ __forceinline__ float t_log<float>(float x) { return logf(x); }
template <> __device__ __forceinline__ double t_log<double>(double x) { return log(x); }

template <typename T, int G> __device__ __forceinline__ T group_max(T v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) { T w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
  return v;
}
template <typename T, int G> __device__ __forceinline__ T group_sum(T v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// finish a scalar reduction: per-block partial -> last block sums partials in fixed order
__device__ __forceinline__ void finish_reduction(double block_val, double* partials, unsigned int* counter, double* sm, bool* is_last,
                                                 double scale, void* out, bool out_f64) {
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = block_val;
    __threadfence();
    unsigned int t = atomicAdd(counter, 1u);
    *is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (*is_last) {
    __threadfence();
    double a = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) a += partials[i];
    double tot = block_sum(a, sm);
    if (threadIdx.x == 0) {
      if (out_f64) *(double*)out = tot * scale; else *(float*)out = (float)(tot * scale);
      *counter = 0u;
    }
  }
}

// MODE 0: fused softmax + CE (Z in);  MODE 1: softmax only;  MODE 2: CE on probabilities (P in, dP out)
template <typename T, int G, int MODE, bool IDX>
__global__ void __launch_bounds__(256)
softmax_ce_kernel(const T* __restrict__ Z, const T* __restrict__ Y, const int32_t* __restrict__ labels, T* __restrict__ P, T* __restrict__ dZ,
                  void* loss_out, int64_t B, int C, double inv_batch, double* partials, unsigned int* counter) {
  __shared__ double sm[32];
  __shared__ bool is_last;
  const int lane = threadIdx.x % G;
  const int64_t groups_per_block = blockDim.x / G;
  const int64_t gstride = groups_per_block * gridDim.x;
  double loss_acc = 0.0;
  const T eps = Eps<T>::v();
  const T invB = (T)inv_batch;
  for (int64_t row = (int64_t)blockIdx.x * groups_per_block + threadIdx.x / G; row < (B + gstride - 1) / gstride * gstride; row += gstride) {
    // all lanes of a group share `row`; whole warps iterate together so shuffles stay converged
    const bool live = row < B;
    const T* z = Z + (live ? row : 0) * C;
    const T* y = IDX ? nullptr : Y + (live ? row : 0) * C;
    const int lab = (IDX && live) ? labels[row] : -1;
    T mx = -INFINITY, se = T(0);
    if (MODE != 2) {
      for (int c = lane; c < C; c += G) { T v = z[c]; mx = v > mx ? v : mx; }
      mx = group_max<T, G>(mx);
      for (int c = lane; c < C; c += G) se += t_exp<T>(z[c] - mx);
      se = group_sum<T, G>(se);
    }
    T dot_dp_p = T(0), lrow = T(0);
    if (MODE != 1) {
      for (int c = lane; c < C; c += G) {
        T p = MODE == 2 ? z[c] : t_exp<T>(z[c] - mx) / se;
        T yc = IDX ? (c == lab ? T(1) : T(0)) : y[c];
        T pe = p + eps;
        lrow -= yc * t_log<T>(pe);
        dot_dp_p += (-yc * invB / pe) * p;
      }
      lrow = group_sum<T, G>(lrow);
      dot_dp_p = group_sum<T, G>(dot_dp_p);
    }
    if (live) {
      for (int c = lane; c < C; c += G) {
        T p = MODE == 2 ? z[c] : t_exp<T>(z[c] - mx) / se;
        if (MODE != 2 && P) P[row * C + c] = p;
        if (MODE != 1 && dZ) {
          T yc = IDX ? (c == lab ? T(1) : T(0)) : y[c];
          T dp = -yc * invB / (p + eps);
          dZ[row * C + c] = MODE == 2 ? dp : p * (dp - dot_dp_p);
        }
      }
      if (lane == 0) loss_acc += (double)lrow;
    }
  }
  if (MODE != 1) {
    double bs = block_sum(loss_acc, sm);
    finish_reduction(bs, partials, counter, sm, &is_last, inv_batch, loss_out, sizeof(T) == 8);
  }
}

// dZ = P .* (dP - sum_c dP.*P)   (NNlib ∇softmax)
template <typename T, int G>
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const T* __restrict__ P, const T* __restrict__ dP, T* __restrict__ dZ, int64_t B, int C) {
  const int lane = threadIdx.x % G;
  const int64_t gpb = blockDim.x / G, gstride = gpb * gridDim.x;
  for (int64_t row = (int64_t)blockIdx.x * gpb + threadIdx.x / G; row < (B + gstride - 1) / gstride * gstride; row += gstride) {
    const bool live = row < B;
    const int64_t r = live ? row : 0;
    T s = T(0);
    for (int c = lane; c < C; c += G) s += dP[r * C + c] * P[r * C + c];
    s = group_sum<T, G>(s);
    if (live)
      for (int c = lane; c < C; c += G) dZ[r * C + c] = P[r * C + c] * (dP[r * C + c] - s);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) mse_kernel(const T* __restrict__ Yh, const T* __restrict__ Y, T* __restrict__ dY, void* loss_out, int64_t n,
                                                  double inv_denom, double* partials, unsigned int* counter) {
  __shared__ double sm[32];
  __shared__ bool is_last;
  double acc = 0.0;
  const T two_inv = (T)(2.0 * inv_denom);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    T d = Yh[i] - Y[i];
    acc += (double)(d * d);
    if (dY) dY[i] = two_inv * d;
  }
  double bs = block_sum(acc, sm);
  finish_reduction(bs, partials, counter, sm, &is_last, inv_denom, loss_out, sizeof(T) == 8);
}

template <typename T>
__global__ void __launch_bounds__(256) class_error_kernel(const T* __restrict__ P, const T* __restrict__ Y, int64_t B, int C, double* partials,
                                                          unsigned int* counter, double* result) {
  __shared__ double sm[32];
  __shared__ bool is_last;
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < B; r += stride) {
    int ap = 0, ay = 0;
    T bp = P[r * C], by = Y[r * C];
    for (int c = 1; c < C; ++c) {
      T p = P[r * C + c], y = Y[r * C + c];
      if (p > bp) { bp = p; ap = c; }
      if (y > by) { by = y; ay = c; }
    }
    acc += (ap != ay) ? 1.0 : 0.0;
  }
  double bs = block_sum(acc, sm);
  finish_reduction(bs, partials, counter, sm, &is_last, 1.0 / (double)B, result, true);
}

template <typename T, int MODE, bool IDX>
int launch_sce(pg_ctx* ctx, const void* Z, const void* Y, const int32_t* labels, void* P, void* dZ, void* loss, int64_t B, int C, double inv_batch) {
  if (B <= 0) return PG_OK;
  int G = C <= 4 ? 4 : C <= 8 ? 8 : C <= 16 ? 16 : 32;
  int64_t gpb = 256 / G;
  int64_t blocks = pg_cdiv(B, gpb);
  int64_t cap = (int64_t)ctx->num_sms * 8;
  if (blocks > cap) blocks = cap;
  int rc = pg_ws_ensure(ctx, (size_t)blocks * sizeof(double));
  if (rc != PG_OK) return rc;
  double* part = (double*)ctx->ws;
  unsigned int* cnt = ctx->counters + 1;
#define SCE_LAUNCH(GG) softmax_ce_kernel<T, GG, MODE, IDX><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const T*)Z, (const T*)Y, labels, (T*)P, (T*)dZ, loss, B, C, inv_batch, part, cnt)
  switch (G) {
    case 4: SCE_LAUNCH(4); break;
    case 8: SCE_LAUNCH(8); break;
    case 16: SCE_LAUNCH(16); break;
    default: SCE_LAUNCH(32); break;
  }
#undef SCE_LAUNCH
  PG_LAUNCHED(ctx);
  return PG_OK;
}

}  // namespace

#define PG_FLOAT_DTYPE(dtype) PG_REQUIRE((dtype) == PG_F32 || (dtype) == PG_F64, "dtype must be PG_F32 or PG_F64")

extern "C" {

int pg_softmax_ce_fwd_bwd(pg_ctx* ctx, int dtype, const void* Z, const void* Y, void* P, void* dZ, void* loss_dev, int64_t B, int C, double inv_batch) {
  PG_CHECK_CTX(ctx); PG_FLOAT_DTYPE(dtype);
  PG_REQUIRE(Z && Y && loss_dev && C > 0, "null pointer or C <= 0");
  if (dtype == PG_F32) return launch_sce<float, 0, false>(ctx, Z, Y, nullptr, P, dZ, loss_dev, B, C, inv_batch);
  return launch_sce<double, 0, false>(ctx, Z, Y, nullptr, P, dZ, loss_dev, B, C, inv_batch);
}

int pg_softmax_ce_idx_fwd_bwd(pg_ctx* ctx, int dtype, const void* Z, const int32_t* labels, void* P, void* dZ, void* loss_dev, int64_t B, int C, double inv_batch) {
  PG_CHECK_CTX(ctx); PG_FLOAT_DTYPE(dtype);
  PG_REQUIRE(Z && labels && loss_dev && C > 0, "null pointer or C <= 0");
  if (dtype == PG_F32) return launch_sce<float, 0, true>(ctx, Z, nullptr, labels, P, dZ, loss_dev, B, C, inv_batch);
  return launch_sce<double, 0, true>(ctx, Z, nullptr, labels, P, dZ, loss_dev, B, C, inv_batch);
}

int pg_softmax_fwd(pg_ctx* ctx, int dtype, const void* Z, void* P, int64_t B, int C) {
  PG_CHECK_CTX(ctx); PG_FLOAT_DTYPE(dtype);
  PG_REQUIRE(Z && P && C > 0, "null pointer or C <= 0");
  if (dtype == PG_F32) return launch_sce<float, 1, false>(ctx, Z, nullptr, nullptr, P, nullptr, nullptr, B, C, 0.0);
  return launch_sce<double, 1, false>(ctx, Z, nullptr, nullptr, P, nullptr, nullptr, B, C, 0.0);
}

int pg_cross_entropy_fwd_bwd(pg_ctx* ctx, int dtype, const void* P, const void* Y, void* dP, void* loss_dev, int64_t B, int C, double inv_batch) {
  PG_CHECK_CTX(ctx); PG_FLOAT_DTYPE(dtype);
  PG_REQUIRE(P && Y && loss_dev && C > 0, "null pointer or C <= 0");
  if (dtype == PG_F32) return launch_sce<float, 2, false>(ctx, P, Y, nullptr, nullptr, dP, loss_dev, B, C, inv_batch);
  return launch_sce<double, 2, false>(ctx, P, Y, nullptr, nullptr, dP, loss_dev, B, C, inv_batch);
}

int pg_softmax_bwd(pg_ctx* ctx, int dtype, const void* P, const void* dP, void* dZ, int64_t B, int C) {
  PG_CHECK_CTX(ctx); PG_FLOAT_DTYPE(dtype);
  PG_REQUIRE(P && dP && dZ && C > 0, "null pointer or C <= 0");
  if (B <= 0) return PG_OK;
  int64_t blocks = pg_cdiv(B, 8);
  int64_t cap = (int64_t)ctx->num_sms * 8;
  if (blocks > cap) blocks = cap;
  if (dtype == PG_F32) softmax_bwd_kernel<float, 32><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const float*)P, (const float*)dP, (float*)dZ, B, C);
  else softmax_bwd_kernel<double, 32><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const double*)P, (const double*)dP, (double*)dZ, B, C);
  PG_LAUNCHED(ctx);
  return PG_OK;
}

int pg_mse_fwd_bwd(pg_ctx* ctx, int dtype, const void* Yhat, const void* Y, void* dY, void* loss_dev, int64_t n, double inv_denom) {
  PG_CHECK_CTX(ctx); PG_FLOAT_DTYPE(dtype);
  PG_REQUIRE(Yhat && Y && loss_dev, "null pointer");
  if (n <= 0) return PG_OK;
  int64_t blocks = pg_cdiv(n, 256);
  int64_t cap = (int64_t)ctx->num_sms * 8;
  if (blocks > cap) blocks = cap;
  int rc = pg_ws_ensure(ctx, (size_t)blocks * sizeof(double));
  if (rc != PG_OK) return rc;
  if (dtype == PG_F32)
    mse_kernel<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const float*)Yhat, (const float*)Y, (float*)dY, loss_dev, n, inv_denom, (double*)ctx->ws, ctx->counters + 2);
  else
    mse_kernel<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const double*)Yhat, (const double*)Y, (double*)dY, loss_dev, n, inv_denom, (double*)ctx->ws, ctx->counters + 2);
  PG_LAUNCHED(ctx);
  return PG_OK;
}

int pg_class_error(pg_ctx* ctx, int dtype, const void* P, const void* Y, int64_t B, int C, double* result) {
  PG_CHECK_CTX(ctx); PG_FLOAT_DTYPE(dtype);
  PG_REQUIRE(P && Y && result && C > 0 && B > 0, "null pointer or empty input");
  int64_t blocks = pg_cdiv(B, 256);
  int64_t cap = (int64_t)ctx->num_sms * 8;
  if (blocks > cap) blocks = cap;
  int rc = pg_ws_ensure(ctx, (size_t)blocks * sizeof(double));
  if (rc != PG_OK) return rc;
  if (dtype == PG_F32)
    class_error_kernel<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const float*)P, (const float*)Y, B, C, (double*)ctx->ws, ctx->counters + 3, ctx->result_slot + 1);
  else
    class_error_kernel<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const double*)P, (const double*)Y, B, C, (double*)ctx->ws, ctx->counters + 3, ctx->result_slot + 1);
  PG_LAUNCHED(ctx);
  PG_CHECK_CUDA(cudaMemcpyAsync(ctx->result_host + 1, ctx->result_slot + 1, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PG_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  *result = ctx->result_host[1];
  return PG_OK;
}

}  // extern "C"

// ---- data step next to the path: the reference materialises every minibatch on the
// host as `train_x_flat[:, idx]` / `train_y_onehot[:, idx]` (examples/01_mnist_mlp.jl:35-38).  With the
// dataset resident in HBM the gather becomes one coalesced kernel and labels can stay integers.
namespace {
template <typename T>
__global__ void __launch_bounds__(256) gather_rows_kernel(const T* __restrict__ src, const int32_t* __restrict__ idx, T* __restrict__ dst, int64_t rows, int64_t row_elems) {
  const int64_t total = rows * row_elems;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / row_elems, c = i - r * row_elems;
    dst[i] = src[(int64_t)idx[r] * row_elems + c];
  }
}
template <typename T>
__global__ void __launch_bounds__(256) onehot_kernel(const int32_t* __restrict__ labels, T* __restrict__ out, int64_t B, int C) {
  const int64_t total = B * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (int)(i % C) == labels[i / C] ? T(1) : T(0);
}
}  // namespace

extern "C" {

int pg_gather_rows(pg_ctx* ctx, int dtype, const void* src, const int32_t* idx, void* dst, int64_t rows, int64_t row_elems) {
  PG_CHECK_CTX(ctx); PG_FLOAT_DTYPE(dtype);
  PG_REQUIRE(src && idx && dst, "null pointer");
  const int64_t total = rows * row_elems;
  if (total <= 0) return PG_OK;
  int64_t blocks = pg_cdiv(total, 256);
  const int64_t cap = (int64_t)ctx->num_sms * 16;
  if (blocks > cap) blocks = cap;
  if (dtype == PG_F32) gather_rows_kernel<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const float*)src, idx, (float*)dst, rows, row_elems);
  else gather_rows_kernel<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const double*)src, idx, (double*)dst, rows, row_elems);
  PG_LAUNCHED(ctx);
  return PG_OK;
}

int pg_onehot(pg_ctx* ctx, int dtype, const int32_t* labels, void* out, int64_t B, int C) {
  PG_CHECK_CTX(ctx); PG_FLOAT_DTYPE(dtype);
  PG_REQUIRE(labels && out && C > 0, "null pointer or C <= 0");
  const int64_t total = B * C;
  if (total <= 0) return PG_OK;
  int64_t blocks = pg_cdiv(total, 256);
  const int64_t cap = (int64_t)ctx->num_sms * 16;
  if (blocks > cap) blocks = cap;
  if (dtype == PG_F32) onehot_kernel<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(labels, (float*)out, B, C);
  else onehot_kernel<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(labels, (double*)out, B, C);
  PG_LAUNCHED(ctx);
  return PG_OK;
}

}  // extern "C"
