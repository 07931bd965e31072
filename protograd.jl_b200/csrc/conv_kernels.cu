// conv_kernels.cu — Conv (src/layers.jl:59 -> NNlib.conv: TRUE convolution, kernel flipped in both
// spatial dims, stride 1, no padding), its gradients (NNlib ∇conv_filter / ∇conv_data), max/mean
// pooling with window == stride (layers.jl:93,99; NNlib ∇maxpool routes to the FIRST max, kW
// fastest) and relu (layers.jl:81; relu'(0) = 0).  NCHW row-major == Julia [W,H,C,N].
// The MNIST-size layers (Cin 1/6, Cout 6/16, 5x5) are activation-bandwidth-bound, so these are
// direct CUDA-core kernels with coalesced W-fastest indexing, weights staged in shared memory.
#include "common.cuh"

namespace {

constexpr int COT = 8;  // output channels per thread in the forward kernel

template <typename T>
__global__ void __launch_bounds__(256)
conv_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ b, T* __restrict__ y,
                int64_t N, int Cin, int H, int W, int Cout, int kH, int kW, int relu, int w_in_smem) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* ws = reinterpret_cast<T*>(smem_raw);
  const int Ho = H - kH + 1, Wo = W - kW + 1;
  const int co0 = blockIdx.y * COT;
  const int nco = (Cout - co0) < COT ? (Cout - co0) : COT;
  const int wsz = Cin * kH * kW;
  if (w_in_smem) {
    for (int i = threadIdx.x; i < nco * wsz; i += blockDim.x) ws[i] = w[(int64_t)co0 * wsz + i];
    __syncthreads();
  }
  const T* wt = w_in_smem ? ws : (w + (int64_t)co0 * wsz);
  const int64_t total = N * Ho * Wo;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % Wo);
    const int i = (int)((idx / Wo) % Ho);
    const int64_t n = idx / ((int64_t)Wo * Ho);
    T acc[COT];
#pragma unroll
    for (int c = 0; c < COT; ++c) acc[c] = T(0);
    for (int ci = 0; ci < Cin; ++ci) {
      const T* xp = x + ((n * Cin + ci) * H + i) * W + j;
      for (int a = 0; a < kH; ++a) {
        for (int bb = 0; bb < kW; ++bb) {
          const T xv = xp[(kH - 1 - a) * W + (kW - 1 - bb)];
          const int wo = (ci * kH + a) * kW + bb;
#pragma unroll
          for (int c = 0; c < COT; ++c)
            if (c < nco) acc[c] += xv * wt[c * wsz + wo];
        }
      }
    }
#pragma unroll
    for (int c = 0; c < COT; ++c) {
      if (c < nco) {
        T v = acc[c] + (b ? b[co0 + c] : T(0));
        if (relu) v = v < T(0) ? T(0) : v;
        y[((n * Cout + co0 + c) * Ho + i) * Wo + j] = v;
      }
    }
  }
}

// dx[n,ci,u,v] = sum_{co,a,b} dy[n,co,u-(kH-1-a), v-(kW-1-b)] * w[co,ci,a,b]
template <typename T>
__global__ void __launch_bounds__(256)
conv_dgrad_kernel(const T* __restrict__ dy, const T* __restrict__ w, T* __restrict__ dx, int64_t N, int Cin, int H, int W, int Cout, int kH, int kW) {
  const int Ho = H - kH + 1, Wo = W - kW + 1;
  const int64_t total = N * Cin * H * W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(idx % W);
    const int u = (int)((idx / W) % H);
    const int ci = (int)((idx / ((int64_t)W * H)) % Cin);
    const int64_t n = idx / ((int64_t)W * H * Cin);
    T acc = T(0);
    for (int co = 0; co < Cout; ++co) {
      const T* dyp = dy + (n * Cout + co) * Ho * Wo;
      const T* wp = w + ((int64_t)co * Cin + ci) * kH * kW;
      for (int a = 0; a < kH; ++a) {
        const int i = u - (kH - 1 - a);
        if (i < 0 || i >= Ho) continue;
        for (int bb = 0; bb < kW; ++bb) {
          const int j = v - (kW - 1 - bb);
          if (j < 0 || j >= Wo) continue;
          acc += dyp[i * Wo + j] * wp[a * kW + bb];
        }
      }
    }
    dx[idx] = acc;
  }
}

// dw[co,ci,a,b] = sum_{n,i,j} x[n,ci,i+KS-1-a, j+KS-1-b] * dy[n,co,i,j];  db[co] = sum dy.
// grid (Cout*Cin, S): block reduces its slice of (n,i,j) into partials[s][co][ci][KS*KS].
template <typename T, int KS>
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dw_part, T* __restrict__ db_part,
                  int64_t N, int Cin, int H, int W, int Cout, int64_t pos_per_split) {
  constexpr int KK = KS * KS;
  __shared__ T sm[8][KK + 1];
  const int co = blockIdx.x / Cin, ci = blockIdx.x % Cin;
  const int Ho = H - KS + 1, Wo = W - KS + 1;
  const int64_t total = N * Ho * Wo;
  const int64_t pb = (int64_t)blockIdx.y * pos_per_split;
  const int64_t pe = pb + pos_per_split < total ? pb + pos_per_split : total;
  T acc[KK];
#pragma unroll
  for (int k = 0; k < KK; ++k) acc[k] = T(0);
  T dsum = T(0);
  for (int64_t idx = pb + threadIdx.x; idx < pe; idx += blockDim.x) {
    const int j = (int)(idx % Wo);
    const int i = (int)((idx / Wo) % Ho);
    const int64_t n = idx / ((int64_t)Wo * Ho);
    const T g = dy[((n * Cout + co) * Ho + i) * Wo + j];
    dsum += g;
    const T* xp = x + ((n * Cin + ci) * H + i) * W + j;
#pragma unroll
    for (int a = 0; a < KS; ++a)
#pragma unroll
      for (int bb = 0; bb < KS; ++bb) acc[a * KS + bb] += xp[(KS - 1 - a) * W + (KS - 1 - bb)] * g;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < KK; ++k) {
    T v = warp_sum(acc[k]);
    if (lane == 0) sm[wid][k] = v;
  }
  {
    T v = warp_sum(dsum);
    if (lane == 0) sm[wid][KK] = v;
  }
  __syncthreads();
  if (threadIdx.x <= KK) {
    T s = T(0);
#pragma unroll
    for (int q = 0; q < 8; ++q) s += sm[q][threadIdx.x];
    if (threadIdx.x < KK) dw_part[((int64_t)blockIdx.y * Cout * Cin + blockIdx.x) * KK + threadIdx.x] = s;
    else if (ci == 0 && db_part) db_part[(int64_t)blockIdx.y * Cout + co] = s;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) reduce_splits_kernel(const T* __restrict__ part, T* __restrict__ out, int64_t n, int splits) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T s = T(0);
  for (int z = 0; z < splits; ++z) s += part[(int64_t)z * n + i];
  out[i] = s;
}

template <typename T>
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t NC, int H, int W, int k) {
  const int Ho = (H - k) / k + 1, Wo = (W - k) / k + 1;
  const int64_t total = NC * Ho * Wo;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % Wo), i = (int)((idx / Wo) % Ho);
    const int64_t nc = idx / ((int64_t)Wo * Ho);
    const T* xp = x + (nc * H + (int64_t)i * k) * W + (int64_t)j * k;
    T m = xp[0];
    for (int a = 0; a < k; ++a)
      for (int b = 0; b < k; ++b) { T v = xp[a * W + b]; m = v > m ? v : m; }
    y[idx] = m;
  }
}

// Writes every element of each window (dy at the FIRST max, 0 elsewhere); border rows/cols outside any
// window must be zero-filled beforehand when H or W is not a multiple of k.  relu != 0 additionally
// applies relu' of the pool's input (a fused conv+ReLU output): the routed gradient is kept only where
// the max itself is > 0, which equals relu'(x) at the arg-max.
template <typename T>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, int64_t NC, int H, int W, int k, int relu) {
  const int Ho = (H - k) / k + 1, Wo = (W - k) / k + 1;
  const int64_t total = NC * Ho * Wo;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % Wo), i = (int)((idx / Wo) % Ho);
    const int64_t nc = idx / ((int64_t)Wo * Ho);
    const int64_t base = (nc * H + (int64_t)i * k) * W + (int64_t)j * k;
    const T* xp = x + base;
    T m = xp[0];
    int am = 0;
    for (int a = 0; a < k; ++a)
      for (int b = 0; b < k; ++b) { T v = xp[a * W + b]; if (v > m) { m = v; am = a * W + b; } }  // strict >: first max wins
    const T g = (relu && !(m > T(0))) ? T(0) : dy[idx];
    for (int a = 0; a < k; ++a)
      for (int b = 0; b < k; ++b) dx[base + a * W + b] = (a * W + b == am) ? g : T(0);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) meanpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t NC, int H, int W, int k) {
  const int Ho = (H - k) / k + 1, Wo = (W - k) / k + 1;
  const int64_t total = NC * Ho * Wo;
  const T inv = T(1) / T(k * k);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % Wo), i = (int)((idx / Wo) % Ho);
    const int64_t nc = idx / ((int64_t)Wo * Ho);
    const T* xp = x + (nc * H + (int64_t)i * k) * W + (int64_t)j * k;
    T s = T(0);
    for (int a = 0; a < k; ++a)
      for (int b = 0; b < k; ++b) s += xp[a * W + b];
    y[idx] = s * inv;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) meanpool_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int64_t NC, int H, int W, int k) {
  const int Ho = (H - k) / k + 1, Wo = (W - k) / k + 1;
  const int64_t total = NC * H * W;
  const T inv = T(1) / T(k * k);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(idx % W), u = (int)((idx / W) % H);
    const int64_t nc = idx / ((int64_t)W * H);
    const int i = u / k, j = v / k;
    dx[idx] = (i < Ho && j < Wo) ? dy[(nc * Ho + i) * Wo + j] * inv : T(0);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) relu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T v = x[i];
    y[i] = v < T(0) ? T(0) : v;
  }
}
template <typename T>
__global__ void __launch_bounds__(256) relu_bwd_kernel(const T* __restrict__ y, const T* __restrict__ dy, T* __restrict__ dx, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dx[i] = y[i] > T(0) ? dy[i] : T(0);
}

inline unsigned grid_for(pg_ctx* ctx, int64_t total) {
  int64_t blocks = pg_cdiv(total, 256);
  int64_t cap = (int64_t)ctx->num_sms * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

template <typename T>
int conv_fwd_t(pg_ctx* ctx, const T* x, const T* w, const T* b, T* y, int64_t N, int Cin, int H, int W, int Cout, int kH, int kW, int act) {
  const int Ho = H - kH + 1, Wo = W - kW + 1;
  int64_t total = N * Ho * Wo;
  if (total <= 0) return PG_OK;
  {
    int r = pg_conv_fwd_tiled<T>(ctx, x, w, b, y, N, Cin, H, W, Cout, kH, kW, act);
    if (r != 0) return r < 0 ? r : PG_OK;
  }
  size_t wbytes = (size_t)COT * Cin * kH * kW * sizeof(T);
  int in_smem = wbytes <= 48 * 1024;
  dim3 grid(grid_for(ctx, total), (unsigned)pg_cdiv(Cout, COT));
  conv_fwd_kernel<T><<<grid, 256, in_smem ? wbytes : 0, ctx->stream>>>(x, w, b, y, N, Cin, H, W, Cout, kH, kW, act == PG_ACT_RELU, in_smem);
  PG_LAUNCHED(ctx);
  return PG_OK;
}

template <typename T, int KS>
int conv_wgrad_t(pg_ctx* ctx, const T* x, const T* dy, T* dw, T* db, int64_t N, int Cin, int H, int W, int Cout) {
  const int Ho = H - KS + 1, Wo = W - KS + 1;
  const int64_t total = N * Ho * Wo;
  const int64_t pairs = (int64_t)Cout * Cin;
  int64_t want = pg_cdiv(4 * (int64_t)ctx->num_sms, pairs);
  int64_t maxs = pg_cdiv(total, 1024);
  int S = (int)(want < maxs ? want : maxs);
  if (S < 1) S = 1;
  if (S > 4096) S = 4096;
  int64_t pps = pg_cdiv(total, S);
  S = (int)pg_cdiv(total, pps);
  if (total <= 0) { S = 1; pps = 1; }
  const int64_t nw = pairs * KS * KS;
  size_t need = (size_t)S * (nw + Cout) * sizeof(T);
  int rc = pg_ws_ensure(ctx, need);
  if (rc != PG_OK) return rc;
  T* dw_part = (T*)ctx->ws;
  T* db_part = dw_part + (size_t)S * nw;
  dim3 grid((unsigned)pairs, (unsigned)S);
  conv_wgrad_kernel<T, KS><<<grid, 256, 0, ctx->stream>>>(x, dy, dw_part, db ? db_part : nullptr, N, Cin, H, W, Cout, pps);
  PG_LAUNCHED(ctx);
  if (dw) {
    reduce_splits_kernel<T><<<(unsigned)pg_cdiv(nw, 256), 256, 0, ctx->stream>>>(dw_part, dw, nw, S);
    PG_LAUNCHED(ctx);
  }
  if (db) {
    reduce_splits_kernel<T><<<(unsigned)pg_cdiv(Cout, 256), 256, 0, ctx->stream>>>(db_part, db, Cout, S);
    PG_LAUNCHED(ctx);
  }
  return PG_OK;
}

template <typename T>
int conv_bwd_t(pg_ctx* ctx, const T* x, const T* w, const T* dy, T* dw, T* db, T* dx, int64_t N, int Cin, int H, int W, int Cout, int kH, int kW) {
  int rc = PG_OK;
  bool wgrad_done = false;
  if ((dw || db) && N > 0) {
    int r = pg_conv_wgrad_tiled<T>(ctx, x, dy, dw, db, N, Cin, H, W, Cout, kH, kW);
    if (r < 0) return r;
    wgrad_done = r == 1;
  }
  if ((dw || db) && !wgrad_done) {
    PG_REQUIRE(kH == kW, "conv weight gradient supports square filters only");
    switch (kH) {
      case 1: rc = conv_wgrad_t<T, 1>(ctx, x, dy, dw, db, N, Cin, H, W, Cout); break;
      case 2: rc = conv_wgrad_t<T, 2>(ctx, x, dy, dw, db, N, Cin, H, W, Cout); break;
      case 3: rc = conv_wgrad_t<T, 3>(ctx, x, dy, dw, db, N, Cin, H, W, Cout); break;
      case 4: rc = conv_wgrad_t<T, 4>(ctx, x, dy, dw, db, N, Cin, H, W, Cout); break;
      case 5: rc = conv_wgrad_t<T, 5>(ctx, x, dy, dw, db, N, Cin, H, W, Cout); break;
      case 7: rc = conv_wgrad_t<T, 7>(ctx, x, dy, dw, db, N, Cin, H, W, Cout); break;
      default:
        pg_set_error("pg_conv2d_bwd: filter size %d unsupported (1,2,3,4,5,7)", kH);
        return PG_ERR_UNSUPPORTED;
    }
    if (rc != PG_OK) return rc;
  }
  if (dx) {
    int64_t total = N * Cin * H * W;
    int r = total > 0 ? pg_conv_dgrad_tiled<T>(ctx, dy, w, dx, N, Cin, H, W, Cout, kH, kW) : 1;
    if (r < 0) return r;
    if (total > 0 && r == 0) {
      conv_dgrad_kernel<T><<<grid_for(ctx, total), 256, 0, ctx->stream>>>(dy, w, dx, N, Cin, H, W, Cout, kH, kW);
      PG_LAUNCHED(ctx);
    }
  }
  return PG_OK;
}

}  // namespace

#define PG_FLOAT_DTYPE(dtype) PG_REQUIRE((dtype) == PG_F32 || (dtype) == PG_F64, "dtype must be PG_F32 or PG_F64")
#define DISPATCH_T(call_f32, call_f64) \
  if (dtype == PG_F32) { call_f32; } else { call_f64; }

extern "C" {

int pg_conv2d_fwd(pg_ctx* ctx, int dtype, const void* x, const void* w, const void* b, void* y, int64_t N, int Cin, int H, int W, int Cout, int kH, int kW, int act) {
  PG_CHECK_CTX(ctx); PG_FLOAT_DTYPE(dtype);
  PG_REQUIRE(x && w && y, "null pointer");
  PG_REQUIRE(kH >= 1 && kW >= 1 && H >= kH && W >= kW && Cin >= 1 && Cout >= 1, "bad conv geometry");
  if (dtype == PG_F32) return conv_fwd_t<float>(ctx, (const float*)x, (const float*)w, (const float*)b, (float*)y, N, Cin, H, W, Cout, kH, kW, act);
  return conv_fwd_t<double>(ctx, (const double*)x, (const double*)w, (const double*)b, (double*)y, N, Cin, H, W, Cout, kH, kW, act);
}

int pg_conv2d_bwd(pg_ctx* ctx, int dtype, const void* x, const void* w, const void* dy, void* dw, void* db, void* dx, int64_t N, int Cin, int H, int W, int Cout, int kH, int kW) {
  PG_CHECK_CTX(ctx); PG_FLOAT_DTYPE(dtype);
  PG_REQUIRE(dy != nullptr, "null dy");
  PG_REQUIRE(kH >= 1 && kW >= 1 && H >= kH && W >= kW && Cin >= 1 && Cout >= 1, "bad conv geometry");
  if (dtype == PG_F32) return conv_bwd_t<float>(ctx, (const float*)x, (const float*)w, (const float*)dy, (float*)dw, (float*)db, (float*)dx, N, Cin, H, W, Cout, kH, kW);
  return conv_bwd_t<double>(ctx, (const double*)x, (const double*)w, (const double*)dy, (double*)dw, (double*)db, (double*)dx, N, Cin, H, W, Cout, kH, kW);
}

int pg_maxpool2d_fwd(pg_ctx* ctx, int dtype, const void* x, void* y, int64_t NC, int H, int W, int k) {
  PG_CHECK_CTX(ctx); PG_FLOAT_DTYPE(dtype);
  PG_REQUIRE(x && y && k >= 1 && H >= k && W >= k, "bad pool geometry");
  int64_t total = NC * ((H - k) / k + 1) * ((W - k) / k + 1);
  if (total <= 0) return PG_OK;
  DISPATCH_T((maxpool_fwd_kernel<float><<<grid_for(ctx, total), 256, 0, ctx->stream>>>((const float*)x, (float*)y, NC, H, W, k)),
             (maxpool_fwd_kernel<double><<<grid_for(ctx, total), 256, 0, ctx->stream>>>((const double*)x, (double*)y, NC, H, W, k)));
  PG_LAUNCHED(ctx);
  return PG_OK;
}

int pg_maxpool2d_bwd(pg_ctx* ctx, int dtype, const void* x, const void* dy, void* dx, int64_t NC, int H, int W, int k) {
  return pg_maxpool2d_relu_bwd(ctx, dtype, x, dy, dx, NC, H, W, k, 0);
}

int pg_maxpool2d_relu_bwd(pg_ctx* ctx, int dtype, const void* x, const void* dy, void* dx, int64_t NC, int H, int W, int k, int relu) {
  PG_CHECK_CTX(ctx); PG_FLOAT_DTYPE(dtype);
  PG_REQUIRE(x && dy && dx && k >= 1 && H >= k && W >= k, "bad pool geometry");
  int64_t total = NC * ((H - k) / k + 1) * ((W - k) / k + 1);
  if (total <= 0) return PG_OK;
  if (H % k != 0 || W % k != 0)   // windows do not tile the plane: rows/cols outside any window get zero gradient
    PG_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)NC * H * W * (dtype == PG_F64 ? 8 : 4), ctx->stream));
  DISPATCH_T((maxpool_bwd_kernel<float><<<grid_for(ctx, total), 256, 0, ctx->stream>>>((const float*)x, (const float*)dy, (float*)dx, NC, H, W, k, relu)),
             (maxpool_bwd_kernel<double><<<grid_for(ctx, total), 256, 0, ctx->stream>>>((const double*)x, (const double*)dy, (double*)dx, NC, H, W, k, relu)));
  PG_LAUNCHED(ctx);
  return PG_OK;
}

int pg_meanpool2d_fwd(pg_ctx* ctx, int dtype, const void* x, void* y, int64_t NC, int H, int W, int k) {
  PG_CHECK_CTX(ctx); PG_FLOAT_DTYPE(dtype);
  PG_REQUIRE(x && y && k >= 1 && H >= k && W >= k, "bad pool geometry");
  int64_t total = NC * ((H - k) / k + 1) * ((W - k) / k + 1);
  if (total <= 0) return PG_OK;
  DISPATCH_T((meanpool_fwd_kernel<float><<<grid_for(ctx, total), 256, 0, ctx->stream>>>((const float*)x, (float*)y, NC, H, W, k)),
             (meanpool_fwd_kernel<double><<<grid_for(ctx, total), 256, 0, ctx->stream>>>((const double*)x, (double*)y, NC, H, W, k)));
  PG_LAUNCHED(ctx);
  return PG_OK;
}

int pg_meanpool2d_bwd(pg_ctx* ctx, int dtype, const void* dy, void* dx, int64_t NC, int H, int W, int k) {
  PG_CHECK_CTX(ctx); PG_FLOAT_DTYPE(dtype);
  PG_REQUIRE(dy && dx && k >= 1 && H >= k && W >= k, "bad pool geometry");
  int64_t total = NC * H * W;
  if (total <= 0) return PG_OK;
  DISPATCH_T((meanpool_bwd_kernel<float><<<grid_for(ctx, total), 256, 0, ctx->stream>>>((const float*)dy, (float*)dx, NC, H, W, k)),
             (meanpool_bwd_kernel<double><<<grid_for(ctx, total), 256, 0, ctx->stream>>>((const double*)dy, (double*)dx, NC, H, W, k)));
  PG_LAUNCHED(ctx);
  return PG_OK;
}

int pg_relu_fwd(pg_ctx* ctx, int dtype, const void* x, void* y, int64_t n) {
  PG_CHECK_CTX(ctx); PG_FLOAT_DTYPE(dtype);
  if (n <= 0) return PG_OK;
  DISPATCH_T((relu_fwd_kernel<float><<<grid_for(ctx, n), 256, 0, ctx->stream>>>((const float*)x, (float*)y, n)),
             (relu_fwd_kernel<double><<<grid_for(ctx, n), 256, 0, ctx->stream>>>((const double*)x, (double*)y, n)));
  PG_LAUNCHED(ctx);
  return PG_OK;
}

int pg_relu_bwd(pg_ctx* ctx, int dtype, const void* y, const void* dy, void* dx, int64_t n) {
  PG_CHECK_CTX(ctx); PG_FLOAT_DTYPE(dtype);
  if (n <= 0) return PG_OK;
  DISPATCH_T((relu_bwd_kernel<float><<<grid_for(ctx, n), 256, 0, ctx->stream>>>((const float*)y, (const float*)dy, (float*)dx, n)),
             (relu_bwd_kernel<double><<<grid_for(ctx, n), 256, 0, ctx->stream>>>((const double*)y, (const double*)dy, (double*)dx, n)));
  PG_LAUNCHED(ctx);
  return PG_OK;
}

}  // extern "C"

// ---- dropout (src/layers.jl:109-138) --------------------------------------------------------------
// mask_i = rand_i > p ? T(1/(1-p)) : 0, y = x .* mask; the pullback multiplies by the SAME mask.
// The mask is never stored: it is a pure function of (seed, stream, element index) through
// Philox4x32-10, so the backward pass regenerates it, and data-parallel shards that pass their global
// element offset draw exactly the masks the single-GPU full batch would.
namespace {

__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
  c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
}
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

template <typename T>
__global__ void __launch_bounds__(256) dropout_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n, float p, T scale,
                                                      uint64_t seed, uint32_t stream_id, uint64_t elem_offset) {
  const int64_t ngroups = (n + 3) / 4;
  for (int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gidx < ngroups; gidx += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t e0 = elem_offset + (uint64_t)gidx * 4;     // global index of the group's first element
    const uint64_t ctr = e0 >> 2;
    const uint4 r = philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), stream_id, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = gidx * 4 + j;
      if (i < n) {
        const float u = (float)(rr[j] >> 8) * (1.0f / 16777216.0f);   // uniform [0,1), 24 bits like rand(Float32)
        y[i] = u > p ? x[i] * scale : T(0);
      }
    }
  }
}

}  // namespace

extern "C" int pg_dropout(pg_ctx* ctx, int dtype, const void* x, void* y, int64_t n, double p, uint64_t seed, uint32_t stream_id, uint64_t elem_offset) {
  PG_CHECK_CTX(ctx); PG_FLOAT_DTYPE(dtype);
  PG_REQUIRE(x && y, "null pointer");
  PG_REQUIRE(p >= 0.0 && p < 1.0, "dropout probability must be in [0, 1)");
  PG_REQUIRE((elem_offset & 3) == 0, "element offset must be a multiple of 4");
  if (n <= 0) return PG_OK;
  const double scale = 1.0 / (1.0 - p);                       // T(1 / (1 - p)), src/layers.jl:114
  if (dtype == PG_F32) dropout_kernel<float><<<grid_for(ctx, (n + 3) / 4), 256, 0, ctx->stream>>>((const float*)x, (float*)y, n, (float)p, (float)scale, seed, stream_id, elem_offset);
  else dropout_kernel<double><<<grid_for(ctx, (n + 3) / 4), 256, 0, ctx->stream>>>((const double*)x, (double*)y, n, (float)p, scale, seed, stream_id, elem_offset);
  PG_LAUNCHED(ctx);
  return PG_OK;
}


// ---- dropout / cast / zero-padding copy on pitched, mixed-type tensors (pg_dropout_ex) ---------------------------
namespace {

template <typename T> __device__ __forceinline__ float ld_as_f32(const T* p, int64_t i) { return (float)p[i]; }
template <> __device__ __forceinline__ float ld_as_f32<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }
template <typename T> __device__ __forceinline__ void st_from_f32(T* p, int64_t i, float v) { p[i] = (T)v; }
template <> __device__ __forceinline__ void st_from_f32<__nv_bfloat16>(__nv_bfloat16* p, int64_t i, float v) { p[i] = __float2bfloat16_rn(v); }

// One thread per group of 4 consecutive LOGICAL elements (row-major over [rows][cols]); element e draws lane e & 3 of
// Philox counter e >> 2, exactly like dropout_kernel, so a padded / bf16 tensor gets the mask its fp32 twin would.
// Float64 tensors keep double arithmetic (TI == TO == double).
template <typename TI, typename TO, bool DROP>
__global__ void __launch_bounds__(256) dropout_ex_kernel(const TI* __restrict__ x, TO* __restrict__ y, int64_t rows, int64_t cols, int64_t ipitch, int64_t opitch,
                                                         float p, double scale, uint64_t seed, uint32_t stream_id, uint64_t elem_offset) {
  const int64_t n = rows * cols;
  const int64_t ngroups = (n + 3) / 4;
  for (int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gidx < ngroups; gidx += (int64_t)gridDim.x * blockDim.x) {
    uint32_t rr[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    if (DROP) {
      const uint64_t ctr = (elem_offset + (uint64_t)gidx * 4) >> 2;
      const uint4 r = philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), stream_id, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
      rr[0] = r.x; rr[1] = r.y; rr[2] = r.z; rr[3] = r.w;
    }
    int64_t i = gidx * 4;
    int64_t row = i / cols, col = i - row * cols;
#pragma unroll
    for (int j = 0; j < 4; ++j, ++i) {
      if (i >= n) break;
      bool keep = true;
      if (DROP) keep = (float)(rr[j] >> 8) * (1.0f / 16777216.0f) > p;
      if constexpr (sizeof(TI) == 8) {
        const double v = (double)x[row * ipitch + col];
        y[row * opitch + col] = (TO)(keep ? (DROP ? v * scale : v) : 0.0);
      } else {
        const float v = ld_as_f32<TI>(x, row * ipitch + col);
        st_from_f32<TO>(y, row * opitch + col, keep ? (DROP ? v * (float)scale : v) : 0.0f);
      }
      if (++col == cols) { col = 0; ++row; }
    }
  }
}

template <typename TI, typename TO>
int launch_dropout_ex(pg_ctx* ctx, const void* x, void* y, int64_t rows, int64_t cols, int64_t ip, int64_t op, double p, uint64_t seed, uint32_t stream_id, uint64_t off) {
  const int64_t groups = (rows * cols + 3) / 4;
  const double scale = 1.0 / (1.0 - p);
  if (p > 0.0)
    dropout_ex_kernel<TI, TO, true><<<grid_for(ctx, groups), 256, 0, ctx->stream>>>((const TI*)x, (TO*)y, rows, cols, ip, op, (float)p, scale, seed, stream_id, off);
  else
    dropout_ex_kernel<TI, TO, false><<<grid_for(ctx, groups), 256, 0, ctx->stream>>>((const TI*)x, (TO*)y, rows, cols, ip, op, 0.0f, 1.0, 0, 0, 0);
  PG_LAUNCHED(ctx);
  return PG_OK;
}

}  // namespace

extern "C" int pg_dropout_ex(pg_ctx* ctx, int in_dtype, int out_dtype, const void* x, void* y, int64_t rows, int64_t cols, int64_t in_pitch, int64_t out_pitch,
                             double p, uint64_t seed, uint32_t stream_id, uint64_t elem_offset) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(x && y, "null pointer");
  PG_REQUIRE(p >= 0.0 && p < 1.0, "dropout probability must be in [0, 1)");
  PG_REQUIRE((elem_offset & 3) == 0, "element offset must be a multiple of 4");
  PG_REQUIRE(in_pitch >= cols && out_pitch >= cols, "row pitch smaller than the row");
  if (rows <= 0 || cols <= 0) return PG_OK;
  typedef __nv_bfloat16 bf;
  if (in_dtype == PG_F32 && out_dtype == PG_F32) return launch_dropout_ex<float, float>(ctx, x, y, rows, cols, in_pitch, out_pitch, p, seed, stream_id, elem_offset);
  if (in_dtype == PG_F32 && out_dtype == PG_BF16) return launch_dropout_ex<float, bf>(ctx, x, y, rows, cols, in_pitch, out_pitch, p, seed, stream_id, elem_offset);
  if (in_dtype == PG_BF16 && out_dtype == PG_BF16) return launch_dropout_ex<bf, bf>(ctx, x, y, rows, cols, in_pitch, out_pitch, p, seed, stream_id, elem_offset);
  if (in_dtype == PG_BF16 && out_dtype == PG_F32) return launch_dropout_ex<bf, float>(ctx, x, y, rows, cols, in_pitch, out_pitch, p, seed, stream_id, elem_offset);
  if (in_dtype == PG_F64 && out_dtype == PG_F64) return launch_dropout_ex<double, double>(ctx, x, y, rows, cols, in_pitch, out_pitch, p, seed, stream_id, elem_offset);
  pg_set_error("pg_dropout_ex: unsupported dtype pair (%d -> %d)", in_dtype, out_dtype);
  return PG_ERR_UNSUPPORTED;
}
