// simt_gemm.cu — FP32 / FP64 CUDA-core GEMM with fused epilogues: the "FP32 SIMT mode" of the
// Linear layer (src/layers.jl:28-34 forward; Zygote/ChainRules pullback dW = X^T dY,
// db = sum_batch dY, dX = dY W^T).  Correct for every shape (N = 1, K = 5, ...), both dtypes;
// the bf16 tensor-core path (tc_gemm.cu) covers the large layers.
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

enum { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_RELU = 2, EPI_MASK = 3 };

// C[m][n] = sum_k A(m,k) B(k,n);  A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn].
// gridDim.z > 1: split-K, raw partials to part[z][M][N] (epilogue applied by the reducer).
template <typename T>
__global__ void __launch_bounds__(256)
gemm_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C, T* __restrict__ part,
            int64_t M, int64_t N, int64_t K, int64_t sam, int64_t sak, int64_t sbk, int64_t sbn, int64_t ldc,
            int epi, const T* __restrict__ bias, const T* __restrict__ mask, int64_t k_per_split) {
  __shared__ __align__(16) T As[BK][BM + PAD];
  __shared__ __align__(16) T Bs[BK][BN + PAD];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t kb = (int64_t)blockIdx.z * k_per_split;
  const int64_t ke = (kb + k_per_split < K) ? kb + k_per_split : K;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);

  const bool a_kcontig = (sak == 1);
  const bool b_ncontig = (sbn == 1);
  for (int64_t k0 = kb; k0 < ke; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int mm, kk;
      if (a_kcontig) { kk = t & 15; mm = (t >> 4) + 16 * i; } else { mm = t & 63; kk = (t >> 6) + 4 * i; }
      int64_t gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < ke) ? A[gm * sam + gk * sak] : T(0);
      int nn, kb2;
      if (b_ncontig) { nn = t & 63; kb2 = (t >> 6) + 4 * i; } else { kb2 = t & 15; nn = (t >> 4) + 16 * i; }
      int64_t gn = n0 + nn, gk2 = k0 + kb2;
      Bs[kb2][nn] = (gn < N && gk2 < ke) ? B[gk2 * sbk + gn * sbn] : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      T v = acc[i][j];
      if (split) { part[((int64_t)blockIdx.z * M + gm) * N + gn] = v; continue; }
      if (epi == EPI_BIAS || epi == EPI_BIAS_RELU) {
        if (bias) v += bias[gn];
        if (epi == EPI_BIAS_RELU) v = v < T(0) ? T(0) : v;
      } else if (epi == EPI_MASK) {
        v = mask[gm * ldc + gn] > T(0) ? v : T(0);
      }
      C[gm * ldc + gn] = v;
    }
  }
}


// ---- larger register tile for Float32: 128 x 128 x 16 per CTA, 8 x 8 per thread, 128-bit shared-memory operand loads,
// the next k-tile prefetched into registers while the current one is consumed (double-buffered shared memory: one
// __syncthreads per k-tile).  16 FMA per shared-memory load instead of 2 -> FMA-bound instead of LDS-bound.  Same
// arguments and epilogues as gemm_kernel; used when the output is at least one tile in each direction (FP32-SIMT mode of the
// wide layers; the 64 x 64 kernel keeps the small and Float64 shapes).
constexpr int TM = 128, TN = 128, TK = 16, TPAD = 4;

__global__ void __launch_bounds__(256)
gemm128_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, float* __restrict__ part,
               int64_t M, int64_t N, int64_t K, int64_t sam, int64_t sak, int64_t sbk, int64_t sbn, int64_t ldc,
               int epi, const float* __restrict__ bias, const float* __restrict__ mask, int64_t k_per_split) {
  __shared__ __align__(16) float As[2][TK][TM + TPAD];
  __shared__ __align__(16) float Bs[2][TK][TN + TPAD];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * TM, n0 = (int64_t)blockIdx.x * TN;
  const int64_t kb = (int64_t)blockIdx.z * k_per_split;
  const int64_t ke = (kb + k_per_split < K) ? kb + k_per_split : K;
  const bool a_kcontig = (sak == 1), b_ncontig = (sbn == 1);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
  float ra[8], rb[8];
  auto gload = [&](int64_t k0) {              // this thread's 8 + 8 elements of the k-tile starting at k0
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int mm, kk;
      if (a_kcontig) { kk = t & 15; mm = (t >> 4) + 16 * i; } else { mm = t & 127; kk = (t >> 7) + 2 * i; }
      const int64_t gm = m0 + mm, gk = k0 + kk;
      ra[i] = (gm < M && gk < ke) ? __ldg(A + gm * sam + gk * sak) : 0.0f;
      int nn, k2;
      if (b_ncontig) { nn = t & 127; k2 = (t >> 7) + 2 * i; } else { k2 = t & 15; nn = (t >> 4) + 16 * i; }
      const int64_t gn = n0 + nn, gk2 = k0 + k2;
      rb[i] = (gn < N && gk2 < ke) ? __ldg(B + gk2 * sbk + gn * sbn) : 0.0f;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int mm, kk;
      if (a_kcontig) { kk = t & 15; mm = (t >> 4) + 16 * i; } else { mm = t & 127; kk = (t >> 7) + 2 * i; }
      As[buf][kk][mm] = ra[i];
      int nn, k2;
      if (b_ncontig) { nn = t & 127; k2 = (t >> 7) + 2 * i; } else { k2 = t & 15; nn = (t >> 4) + 16 * i; }
      Bs[buf][k2][nn] = rb[i];
    }
  };
  int buf = 0;
  if (kb < ke) { gload(kb); sstore(0); }
  __syncthreads();
  for (int64_t k0 = kb; k0 < ke; k0 += TK) {
    const bool more = k0 + TK < ke;
    if (more) gload(k0 + TK);                 // global latency hidden behind this tile's FMAs
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4 + 64]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4 + 64]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) sstore(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t gm = m0 + ty * 4 + (i & 3) + (i >> 2) * 64;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t gn = n0 + tx * 4 + (j & 3) + (j >> 2) * 64;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (split) { part[((int64_t)blockIdx.z * M + gm) * N + gn] = v; continue; }
      if (epi == EPI_BIAS || epi == EPI_BIAS_RELU) {
        if (bias) v += bias[gn];
        if (epi == EPI_BIAS_RELU) v = v < 0.0f ? 0.0f : v;
      } else if (epi == EPI_MASK) {
        v = mask[gm * ldc + gn] > 0.0f ? v : 0.0f;
      }
      C[gm * ldc + gn] = v;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const T* __restrict__ part, T* __restrict__ C, int64_t MN, int splits) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= MN) return;
  T s = T(0);
  for (int z = 0; z < splits; ++z) s += part[(int64_t)z * MN + i];
  C[i] = s;
}

// same, with the GEMM's epilogue (bias / ReLU / ReLU' mask) and a pitched output: split-K forward and data-gradient launches
template <typename T>
__global__ void __launch_bounds__(256) splitk_reduce_epi_kernel(const T* __restrict__ part, T* __restrict__ C, int64_t M, int64_t N, int64_t ldc, int splits,
                                                                int epi, const T* __restrict__ bias, const T* __restrict__ mask) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const int64_t gm = i / N, gn = i - gm * N;
  T v = T(0);
  for (int z = 0; z < splits; ++z) v += part[(int64_t)z * M * N + i];
  if (epi == EPI_BIAS || epi == EPI_BIAS_RELU) {
    if (bias) v += bias[gn];
    if (epi == EPI_BIAS_RELU) v = v < T(0) ? T(0) : v;
  } else if (epi == EPI_MASK) {
    v = mask[gm * ldc + gn] > T(0) ? v : T(0);
  }
  C[gm * ldc + gn] = v;
}

template <typename TO, typename TI> __device__ __forceinline__ TO cvt(TI v) {
  if constexpr (std::is_same<TI, __nv_bfloat16>::value) return (TO)__bfloat162float(v);
  else return (TO)v;
}

// column sums: out[n] = sum_m X[m][n].  grid (ceil(N/32), S), block (32, 8).
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) colsum_kernel(const TI* __restrict__ X, TO* __restrict__ out, int64_t M, int64_t N, int64_t rows_per_split) {
  __shared__ TO sm[8][33];
  int64_t n = (int64_t)blockIdx.x * 32 + threadIdx.x;
  int64_t mb = (int64_t)blockIdx.y * rows_per_split;
  int64_t me = mb + rows_per_split < M ? mb + rows_per_split : M;
  TO acc = TO(0);
  if (n < N)
    for (int64_t m = mb + threadIdx.y; m < me; m += 8) acc += cvt<TO>(X[m * N + n]);
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    TO s = TO(0);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sm[i][threadIdx.x];
    out[(int64_t)blockIdx.y * N + n] = s;
  }
}
// vectorised column sums for wide matrices: each thread owns VEC consecutive columns (16-byte loads),
// 8 row-lanes per block; grid (ceil(N / (32*VEC)), S).  HBM-bound: one pass over X.
template <typename TI, typename TO, int VEC>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const TI* __restrict__ X, TO* __restrict__ out, int64_t M, int64_t N, int64_t rows_per_split) {
  __shared__ TO sm[8][32 * VEC + 1];
  const int64_t n0 = ((int64_t)blockIdx.x * 32 + threadIdx.x) * VEC;
  const int64_t mb = (int64_t)blockIdx.y * rows_per_split;
  const int64_t me = mb + rows_per_split < M ? mb + rows_per_split : M;
  TO acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = TO(0);
  if (n0 < N) {
    for (int64_t m = mb + threadIdx.y; m < me; m += 8) {
      const uint4 raw = *reinterpret_cast<const uint4*>(X + m * N + n0);
      const TI* e = reinterpret_cast<const TI*>(&raw);
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] += cvt<TO>(e[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) sm[threadIdx.y][threadIdx.x * VEC + j] = acc[j];
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x;
  for (int c = t; c < 32 * VEC; c += 256) {
    const int64_t n = (int64_t)blockIdx.x * 32 * VEC + c;
    if (n < N) {
      TO s = TO(0);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += sm[i][c];
      out[(int64_t)blockIdx.y * N + n] = s;
    }
  }
}

template <typename T>
int gemm_launch(pg_ctx* ctx, const T* A, const T* B, T* C, int64_t M, int64_t N, int64_t K, int64_t sam, int64_t sak, int64_t sbk,
                int64_t sbn, int64_t ldc, int epi, const T* bias, const T* mask, bool allow_split) {
  if (M <= 0 || N <= 0) return PG_OK;
  // Float32 outputs with at least one 128 x 128 tile per SM: the register-tiled kernel (smaller problems fill the GPU
  // better with 64 x 64 tiles: measured C1 at B = 8192, N = 100: 0.247 ms/step with 64-tiles, 0.323 with 128-tiles)
  const bool big = std::is_same<T, float>::value && pg_cdiv(M, TM) * pg_cdiv(N, TN) >= ctx->num_sms;
  const int bm = big ? TM : BM, bn = big ? TN : BN, bk = big ? TK : BK;
  int64_t gx = pg_cdiv(N, bn), gy = pg_cdiv(M, bm);
  PG_REQUIRE(gy <= 65535, "M too large for the SIMT GEMM grid");
  int splits = 1;
  const int64_t tiles = gx * gy;
  if (allow_split && epi == EPI_NONE && ldc == N) {
    int64_t want = pg_cdiv(2 * (int64_t)ctx->num_sms, tiles);
    int64_t maxs = pg_cdiv(K, 8 * bk);
    splits = (int)(want < maxs ? want : maxs);
    if (splits < 1) splits = 1;
    if (splits > 256) splits = 256;
  } else if (tiles * 8 <= ctx->num_sms && K >= 32 * bk) {
    // a handful of tiles with a deep reduction (forward / data gradient at small batch: 128 x 100 x 784 is 4 CTAs walking 49
    // k-tiles): split K over idle SMs, 4 k-tiles per split at least; the reducer applies the epilogue.  Shorter reductions
    // (K < 512) do not pay for the second launch (measured on the conv net's 256-120-84-10 head at batch 128)
    int64_t want = ctx->num_sms / tiles, maxs = K / (4 * bk);
    splits = (int)(want < maxs ? want : maxs);
    if (splits < 1) splits = 1;
    if (splits > 32) splits = 32;
  }
  int64_t kps = pg_cdiv(pg_cdiv(K, splits), bk) * bk;
  splits = (int)pg_cdiv(K, kps);
  if (K <= 0) { splits = 1; kps = bk; }
  T* part = nullptr;
  if (splits > 1) {
    int rc = pg_ws_ensure(ctx, (size_t)splits * M * N * sizeof(T));
    if (rc != PG_OK) return rc;
    part = (T*)ctx->ws;
  }
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)splits);
  if constexpr (std::is_same<T, float>::value) {
    if (big) gemm128_kernel<<<grid, 256, 0, ctx->stream>>>(A, B, C, part, M, N, K, sam, sak, sbk, sbn, ldc, epi, bias, mask, kps);
    else gemm_kernel<T><<<grid, 256, 0, ctx->stream>>>(A, B, C, part, M, N, K, sam, sak, sbk, sbn, ldc, epi, bias, mask, kps);
  } else {
    gemm_kernel<T><<<grid, 256, 0, ctx->stream>>>(A, B, C, part, M, N, K, sam, sak, sbk, sbn, ldc, epi, bias, mask, kps);
  }
  PG_LAUNCHED(ctx);
  if (splits > 1) {
    int64_t MN = M * N;
    if (epi == EPI_NONE && ldc == N) splitk_reduce_kernel<T><<<(unsigned)pg_cdiv(MN, 256), 256, 0, ctx->stream>>>(part, C, MN, splits);
    else splitk_reduce_epi_kernel<T><<<(unsigned)pg_cdiv(MN, 256), 256, 0, ctx->stream>>>(part, C, M, N, ldc, splits, epi, bias, mask);
    PG_LAUNCHED(ctx);
  }
  return PG_OK;
}

template <typename T>
int linear_fwd_t(pg_ctx* ctx, const T* X, const T* W, const T* b, T* Y, int64_t M, int64_t K, int64_t N, int act) {
  return gemm_launch<T>(ctx, X, W, Y, M, N, K, K, 1, N, 1, N, act == PG_ACT_RELU ? EPI_BIAS_RELU : EPI_BIAS, b, nullptr, false);
}

template <typename T>
int linear_bwd_t(pg_ctx* ctx, const T* X, const T* W, const T* dY, T* dW, T* db, T* dX, const T* mask, int64_t M, int64_t K, int64_t N) {
  int rc;
  if (dW) {  // dW[K][N] = X^T dY : A(m=k,kk=b) = X[b*K+k], B(kk=b,n) = dY[b*N+n]
    rc = gemm_launch<T>(ctx, X, dY, dW, K, N, M, 1, K, N, 1, N, EPI_NONE, nullptr, nullptr, true);
    if (rc != PG_OK) return rc;
  }
  if (db) {
    rc = pg_colsum_launch(ctx, sizeof(T) == 8 ? PG_F64 : PG_F32, dY, db, M, N);
    if (rc != PG_OK) return rc;
  }
  if (dX) {  // dX[M][K] = dY W^T : A(m=b,kk=n) = dY[b*N+n], B(kk=n,n'=k) = W[k*N+n]
    rc = gemm_launch<T>(ctx, dY, W, dX, M, K, N, N, 1, 1, N, K, mask ? EPI_MASK : EPI_NONE, nullptr, mask, false);
    if (rc != PG_OK) return rc;
  }
  return PG_OK;
}

}  // namespace

// shared with tc_gemm.cu (bf16 input, f32 output)
int pg_colsum_launch(pg_ctx* ctx, int in_dtype, const void* X, void* out, int64_t M, int64_t N) {
  if (N <= 0) return PG_OK;
  const int vec = in_dtype == PG_BF16 ? 8 : in_dtype == PG_F32 ? 4 : 2;
  const bool use_vec = N >= 256 && N % vec == 0 && ((uintptr_t)X & 15) == 0;
  int64_t gx = use_vec ? pg_cdiv(N, 32 * vec) : pg_cdiv(N, 32);
  int64_t want = pg_cdiv(4 * (int64_t)ctx->num_sms, gx);
  int64_t maxs = pg_cdiv(M, 64);
  int S = (int)(want < maxs ? want : maxs);
  if (S < 1) S = 1;
  if (S > 1024) S = 1024;
  if (N <= 1024 && S > 32) S = 32;   // the final reduce is one thread per column walking S partials (128 partials of a 120-wide layer: 10 us)
  int64_t rps = pg_cdiv(M, S);
  S = (int)pg_cdiv(M, rps);
  if (M <= 0) { S = 1; rps = 1; }
  size_t osz = in_dtype == PG_F64 ? 8 : 4;
  void* tgt = out;
  if (S > 1) {
    int rc = pg_ws_ensure(ctx, (size_t)S * N * osz);
    if (rc != PG_OK) return rc;
    tgt = ctx->ws;
  }
  dim3 grid((unsigned)gx, (unsigned)S), block(32, 8);
  if (use_vec) {
    if (in_dtype == PG_F32) colsum_vec_kernel<float, float, 4><<<grid, block, 0, ctx->stream>>>((const float*)X, (float*)tgt, M, N, rps);
    else if (in_dtype == PG_F64) colsum_vec_kernel<double, double, 2><<<grid, block, 0, ctx->stream>>>((const double*)X, (double*)tgt, M, N, rps);
    else colsum_vec_kernel<__nv_bfloat16, float, 8><<<grid, block, 0, ctx->stream>>>((const __nv_bfloat16*)X, (float*)tgt, M, N, rps);
  } else {
    if (in_dtype == PG_F32) colsum_kernel<float, float><<<grid, block, 0, ctx->stream>>>((const float*)X, (float*)tgt, M, N, rps);
    else if (in_dtype == PG_F64) colsum_kernel<double, double><<<grid, block, 0, ctx->stream>>>((const double*)X, (double*)tgt, M, N, rps);
    else colsum_kernel<__nv_bfloat16, float><<<grid, block, 0, ctx->stream>>>((const __nv_bfloat16*)X, (float*)tgt, M, N, rps);
  }
  PG_LAUNCHED(ctx);
  if (S > 1) {
    if (in_dtype == PG_F64) splitk_reduce_kernel<double><<<(unsigned)pg_cdiv(N, 256), 256, 0, ctx->stream>>>((const double*)tgt, (double*)out, N, S);
    else splitk_reduce_kernel<float><<<(unsigned)pg_cdiv(N, 256), 256, 0, ctx->stream>>>((const float*)tgt, (float*)out, N, S);
    PG_LAUNCHED(ctx);
  }
  return PG_OK;
}

extern "C" {

int pg_linear_fwd(pg_ctx* ctx, int dtype, const void* X, const void* W, const void* b, void* Y, int64_t M, int64_t K, int64_t N, int act) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(dtype == PG_F32 || dtype == PG_F64, "dtype must be PG_F32 or PG_F64 (bf16: pg_tc_linear_fwd)");
  PG_REQUIRE(X && W && Y, "null pointer");
  PG_REQUIRE(M >= 0 && K >= 0 && N >= 0, "negative dimension");
  if (dtype == PG_F32) return linear_fwd_t<float>(ctx, (const float*)X, (const float*)W, (const float*)b, (float*)Y, M, K, N, act);
  return linear_fwd_t<double>(ctx, (const double*)X, (const double*)W, (const double*)b, (double*)Y, M, K, N, act);
}

int pg_linear_bwd(pg_ctx* ctx, int dtype, const void* X, const void* W, const void* dY, void* dW, void* db, void* dX,
                  const void* relu_mask, int64_t M, int64_t K, int64_t N) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(dtype == PG_F32 || dtype == PG_F64, "dtype must be PG_F32 or PG_F64 (bf16: pg_tc_linear_bwd)");
  PG_REQUIRE(dY != nullptr, "null dY");
  PG_REQUIRE(!dW || X, "dW needs X");
  PG_REQUIRE(!dX || W, "dX needs W");
  if (dtype == PG_F32)
    return linear_bwd_t<float>(ctx, (const float*)X, (const float*)W, (const float*)dY, (float*)dW, (float*)db, (float*)dX, (const float*)relu_mask, M, K, N);
  return linear_bwd_t<double>(ctx, (const double*)X, (const double*)W, (const double*)dY, (double*)dW, (double*)db, (double*)dX, (const double*)relu_mask, M, K, N);
}

}  // extern "C"
