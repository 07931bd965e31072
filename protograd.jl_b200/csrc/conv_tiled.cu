// conv_tiled.cu — shared-memory-tiled, register-blocked direct convolution kernels for the small-
// channel layers of the reference's conv net (src/layers.jl:59 via NNlib.conv / ∇conv_filter /
// ∇conv_data; examples/02_mnist_conv.jl:25-33: 1->6 and 6->16 channels, 5x5).  These layers move
// ~26 KB of activations per sample for ~1.5 Mflop, i.e. they are bandwidth-/issue-bound, not
// tensor-bound.  Design:
//   * every activation is read from HBM once: whole samples are staged in shared memory, coalesced;
//   * each thread owns a row segment x channel block in registers;
//   * forward / data gradient: the (tiny) filter bank lives in __constant__ memory and every weight
//     index is warp-uniform (channel-group loops are block-level), so weights reach the FMA through
//     the uniform datapath / constant cache instead of one shared-memory load per FMA;
//   * weight gradient: a thread owns ALL output channels for one (ci, filter row): dy is a broadcast
//     read shared by the whole block, x is read once per 16 x 5 accumulators.
//   forward     y[co][i][j]   = b[co] + sum_{ci,a,b} x[ci][i+K-1-a][j+K-1-b] w[co][ci][a][b]   (true convolution)
//   data grad   dx[ci][u][v]  = sum_{co,a,b} dy[co][u+a-(K-1)][v+b-(K-1)] w[co][ci][a][b]      (zero outside dy)
//   weight grad dw[co][ci][a][b] = sum_{n,i,j} x[ci][i+K-1-a][j+K-1-b] dy[co][i][j];  db[co] = sum dy
// Generic shapes fall back to the simple kernels in conv_kernels.cu.
#include "common.cuh"

namespace {

constexpr int COB = 8;    // output channels per thread per pass (forward)
constexpr int CIB = 8;    // input channels per thread per pass (data gradient)
constexpr int DJB = 4;    // outputs per thread along W (data gradient)
constexpr int WCO = 16;   // output channels per thread per pass (weight gradient)
constexpr int WJB = 8;    // dy columns per inner step (weight gradient)
constexpr size_t SMEM_CAP = 100 * 1024;
constexpr int CONST_BYTES = 48 * 1024;

// filter bank of the layer being launched (copied device-to-device on the stream before the kernel)
__constant__ __align__(16) unsigned char g_conv_w[CONST_BYTES];

template <typename T, int KS, int JB>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? 2 : 1)
conv_fwd_tiled_kernel(const T* __restrict__ x, const T* __restrict__ b, T* __restrict__ y,
                      int N, int Cin, int H, int W, int Cout, int relu, int SPB, int SEG) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* xs = reinterpret_cast<T*>(smem_raw);        // [SPB][Cin][H][W]
  const T* cw = reinterpret_cast<const T*>(g_conv_w);   // [Cout][Cin][KS][KS]
  const int Ho = H - KS + 1, Wo = W - KS + 1, chw = Cin * H * W;
  const int IPS = Ho * SEG;                      // items per sample: (row, segment)
  const int G = (Cout + COB - 1) / COB;
  for (int n0 = blockIdx.x * SPB; n0 < N; n0 += gridDim.x * SPB) {
    const int ns = min(SPB, N - n0);
    __syncthreads();
    const T* xg = x + (size_t)n0 * chw;
    for (int i = threadIdx.x; i < ns * chw; i += blockDim.x) xs[i] = xg[i];
    __syncthreads();
    for (int item = threadIdx.x; item < ns * IPS; item += blockDim.x) {
      const int s = item / IPS, r = item - s * IPS;
      const int i = r / SEG, j0 = (r - i * SEG) * JB;
      const T* xsamp = xs + s * chw;
      const size_t n = (size_t)(n0 + s);
      for (int g = 0; g < G; ++g) {              // block-uniform: weight indices below are warp-uniform
        T acc[COB][JB];
#pragma unroll
        for (int c = 0; c < COB; ++c)
#pragma unroll
          for (int j = 0; j < JB; ++j) acc[c][j] = T(0);
        for (int ci = 0; ci < Cin; ++ci) {
#pragma unroll
          for (int a = 0; a < KS; ++a) {
            const T* xr = xsamp + (ci * H + i + KS - 1 - a) * W + j0;
            T row[JB + KS - 1];
#pragma unroll
            for (int q = 0; q < JB + KS - 1; ++q) row[q] = (j0 + q < W) ? xr[q] : T(0);
#pragma unroll
            for (int c = 0; c < COB; ++c) {
              const int co = min(g * COB + c, Cout - 1);
              const T* wp = cw + ((co * Cin + ci) * KS + a) * KS;
#pragma unroll
              for (int bb = 0; bb < KS; ++bb) {
                const T wv = wp[bb];
#pragma unroll
                for (int j = 0; j < JB; ++j) acc[c][j] += row[j + KS - 1 - bb] * wv;
              }
            }
          }
        }
#pragma unroll
        for (int c = 0; c < COB; ++c) {
          const int co = g * COB + c;
          if (co < Cout) {
            const T bv = b ? b[co] : T(0);
            T* yp = y + ((n * Cout + co) * Ho + i) * Wo + j0;
#pragma unroll
            for (int j = 0; j < JB; ++j) {
              if (j0 + j < Wo) {
                T v = acc[c][j] + bv;
                if (relu) v = v < T(0) ? T(0) : v;
                yp[j] = v;
              }
            }
          }
        }
      }
    }
  }
}

template <typename T, int KS>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? 2 : 1)
conv_dgrad_tiled_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int Cin, int H, int W, int Cout, int SPB, int SEG) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* ds = reinterpret_cast<T*>(smem_raw);        // [SPB][Cout][Ho][Wo]
  const T* cw = reinterpret_cast<const T*>(g_conv_w);
  const int Ho = H - KS + 1, Wo = W - KS + 1, ych = Cout * Ho * Wo;
  const int IPS = H * SEG;
  const int G = (Cin + CIB - 1) / CIB;
  for (int n0 = blockIdx.x * SPB; n0 < N; n0 += gridDim.x * SPB) {
    const int ns = min(SPB, N - n0);
    __syncthreads();
    const T* dg = dy + (size_t)n0 * ych;
    for (int i = threadIdx.x; i < ns * ych; i += blockDim.x) ds[i] = dg[i];
    __syncthreads();
    for (int item = threadIdx.x; item < ns * IPS; item += blockDim.x) {
      const int s = item / IPS, r = item - s * IPS;
      const int u = r / SEG, v0 = (r - u * SEG) * DJB;
      const T* dsamp = ds + s * ych;
      const size_t n = (size_t)(n0 + s);
      for (int g = 0; g < G; ++g) {
        T acc[CIB][DJB];
#pragma unroll
        for (int c = 0; c < CIB; ++c)
#pragma unroll
          for (int j = 0; j < DJB; ++j) acc[c][j] = T(0);
        for (int co = 0; co < Cout; ++co) {
#pragma unroll
          for (int a = 0; a < KS; ++a) {
            const int ii = u + a - (KS - 1);                 // dy row feeding output row u through filter row a
            if (ii < 0 || ii >= Ho) continue;
            const T* dr = dsamp + (co * Ho + ii) * Wo;
            T row[DJB + KS - 1];
#pragma unroll
            for (int q = 0; q < DJB + KS - 1; ++q) {
              const int jj = v0 + q - (KS - 1);
              row[q] = (jj >= 0 && jj < Wo) ? dr[jj] : T(0);
            }
#pragma unroll
            for (int c = 0; c < CIB; ++c) {
              const int ci = min(g * CIB + c, Cin - 1);
              const T* wp = cw + ((co * Cin + ci) * KS + a) * KS;
#pragma unroll
              for (int bb = 0; bb < KS; ++bb) {
                const T wv = wp[bb];
#pragma unroll
                for (int j = 0; j < DJB; ++j) acc[c][j] += row[j + bb] * wv;
              }
            }
          }
        }
#pragma unroll
        for (int c = 0; c < CIB; ++c) {
          const int ci = g * CIB + c;
          if (ci < Cin) {
            T* xp = dx + ((n * Cin + ci) * H + u) * W + v0;
#pragma unroll
            for (int j = 0; j < DJB; ++j)
              if (v0 + j < W) xp[j] = acc[c][j];
          }
        }
      }
    }
  }
}

// weight gradient: thread item = (sample slot, ci, filter row a, row-group rg); it owns
// dw[co0..co0+WCO)[ci][a][0..KS) (WCO x KS accumulators) over rows i = rg, rg+RG, ... of its slot's samples.
template <typename T, int KS>
__global__ void __launch_bounds__(256)
conv_wgrad_tiled_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dw_part, T* __restrict__ db_part,
                        int N, int Cin, int H, int W, int Cout, int SPB, int RG, int samples_per_block) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Ho = H - KS + 1, Wo = W - KS + 1, chw = Cin * H * W, ych = Cout * Ho * Wo;
  const int nitem = Cin * KS;                   // (ci, a)
  T* xs = reinterpret_cast<T*>(smem_raw);       // [SPB][chw]
  T* dys = xs + SPB * chw;                      // [SPB][ych]
  T* red = dys + SPB * ych;                     // [Cout][Cin][KS][KS] + [Cout]  block result
  const int nred = Cout * Cin * KS * KS;
  for (int i = threadIdx.x; i < nred + Cout; i += blockDim.x) red[i] = T(0);
  const int per_slot = nitem * RG;
  const int slot = threadIdx.x / per_slot;
  const int r = threadIdx.x - slot * per_slot;
  const int it = r / RG, rg = r - it * RG;
  const int ci = it / KS, a = it - ci * KS;
  const bool live = slot < SPB;
  const int nb0 = blockIdx.x * samples_per_block;
  const int nb1 = min(N, nb0 + samples_per_block);
  const int G = (Cout + WCO - 1) / WCO;
  for (int g = 0; g < G; ++g) {                  // channel groups of 16: one pass over the samples each
    T acc[WCO][KS];
    T dbacc[WCO];
#pragma unroll
    for (int c = 0; c < WCO; ++c) {
      dbacc[c] = T(0);
#pragma unroll
      for (int k = 0; k < KS; ++k) acc[c][k] = T(0);
    }
    for (int n0 = nb0; n0 < nb1; n0 += SPB) {
      const int ns = min(SPB, nb1 - n0);
      __syncthreads();
      const T* xg = x + (size_t)n0 * chw;
      for (int i = threadIdx.x; i < ns * chw; i += blockDim.x) xs[i] = xg[i];
      const T* dg = dy + (size_t)n0 * ych;
      for (int i = threadIdx.x; i < ns * ych; i += blockDim.x) dys[i] = dg[i];
      __syncthreads();
      if (live && slot < ns) {
        const T* xsamp = xs + slot * chw + (ci * H + KS - 1 - a) * W;
        const T* dsamp = dys + slot * ych;
        for (int i = rg; i < Ho; i += RG) {
          for (int j0 = 0; j0 < Wo; j0 += WJB) {
            T xr[WJB + KS - 1];
#pragma unroll
            for (int q = 0; q < WJB + KS - 1; ++q) xr[q] = (j0 + q < W) ? xsamp[i * W + j0 + q] : T(0);
#pragma unroll
            for (int c = 0; c < WCO; ++c) {
              const int co = min(g * WCO + c, Cout - 1);
              const T* dp = dsamp + (co * Ho + i) * Wo + j0;     // same address for every thread of the slot: broadcast
              T dr[WJB];
#pragma unroll
              for (int j = 0; j < WJB; ++j) dr[j] = (j0 + j < Wo) ? dp[j] : T(0);
#pragma unroll
              for (int bb = 0; bb < KS; ++bb)
#pragma unroll
                for (int j = 0; j < WJB; ++j) acc[c][bb] += xr[j + KS - 1 - bb] * dr[j];
              if (it == 0) {
#pragma unroll
                for (int j = 0; j < WJB; ++j) dbacc[c] += dr[j];
              }
            }
          }
        }
      }
    }
    __syncthreads();
    if (live) {
#pragma unroll
      for (int c = 0; c < WCO; ++c) {
        const int co = g * WCO + c;
        if (co < Cout) {
#pragma unroll
          for (int k = 0; k < KS; ++k) atomicAdd(&red[((co * Cin + ci) * KS + a) * KS + k], acc[c][k]);
          if (it == 0) atomicAdd(&red[nred + co], dbacc[c]);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nred; i += blockDim.x) dw_part[(size_t)blockIdx.x * nred + i] = red[i];
  if (db_part != nullptr)
    for (int c = threadIdx.x; c < Cout; c += blockDim.x) db_part[(size_t)blockIdx.x * Cout + c] = red[nred + c];
}

// out[i] = sum_z part[z][i]: 64 outputs x 4 z-lanes per block, 4 independent accumulators per thread
template <typename T>
__global__ void __launch_bounds__(256) reduce_blocks_kernel(const T* __restrict__ part, T* __restrict__ out, int64_t n, int blocks) {
  __shared__ T sm[4][64];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int64_t i = (int64_t)blockIdx.x * 64 + tx;
  T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
  if (i < n) {
    int z = ty;
    for (; z + 12 < blocks; z += 16) {
      s0 += part[(int64_t)z * n + i]; s1 += part[(int64_t)(z + 4) * n + i];
      s2 += part[(int64_t)(z + 8) * n + i]; s3 += part[(int64_t)(z + 12) * n + i];
    }
    for (; z < blocks; z += 4) s0 += part[(int64_t)z * n + i];
  }
  sm[ty][tx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ty == 0 && i < n) out[i] = (sm[0][tx] + sm[1][tx]) + (sm[2][tx] + sm[3][tx]);
}

template <typename K>
int set_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024) PG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return PG_OK;
}

template <typename T>
int upload_filters(pg_ctx* ctx, const T* w, size_t count) {
  PG_CHECK_CUDA(cudaMemcpyToSymbolAsync(g_conv_w, w, count * sizeof(T), 0, cudaMemcpyDeviceToDevice, ctx->stream));
  return PG_OK;
}

template <typename T, int KS, int JB>
int fwd_launch(pg_ctx* ctx, const T* x, const T* w, const T* b, T* y, int64_t N, int Cin, int H, int W, int Cout, int act) {
  const int Ho = H - KS + 1, Wo = W - KS + 1;
  const int SEG = (Wo + JB - 1) / JB;
  const int IPS = Ho * SEG;
  const size_t wcount = (size_t)Cout * Cin * KS * KS;
  const size_t per_sample = (size_t)Cin * H * W * sizeof(T);
  if (wcount * sizeof(T) > CONST_BYTES || per_sample > SMEM_CAP / 2) return 0;
  // enough samples per iteration to give every thread one item; small tiles keep two blocks resident
  // per SM so one block's load phase overlaps the other's compute phase
  int SPB = (256 + IPS - 1) / IPS;
  const int cap = (int)((SMEM_CAP / 2) / per_sample);
  if (SPB > cap) SPB = cap;
  if (SPB < 1) SPB = 1;
  const size_t smem = (size_t)SPB * per_sample;
  auto kern = conv_fwd_tiled_kernel<T, KS, JB>;
  int rc = set_smem(kern, smem);
  if (rc != PG_OK) return rc;
  rc = upload_filters<T>(ctx, w, wcount);
  if (rc != PG_OK) return rc;
  int64_t groups = pg_cdiv(N, SPB);
  int blocks = (int)(groups < 4 * (int64_t)ctx->num_sms ? groups : 4 * (int64_t)ctx->num_sms);
  kern<<<blocks, 256, smem, ctx->stream>>>(x, b, y, (int)N, Cin, H, W, Cout, act == PG_ACT_RELU, SPB, SEG);
  ctx->launches++;
  PG_CHECK_CUDA(cudaGetLastError());
  return 1;
}

template <typename T, int KS>
int dgrad_launch(pg_ctx* ctx, const T* dy, const T* w, T* dx, int64_t N, int Cin, int H, int W, int Cout) {
  const int Ho = H - KS + 1, Wo = W - KS + 1;
  const int SEG = (W + DJB - 1) / DJB;
  const int IPS = H * SEG;
  const size_t wcount = (size_t)Cout * Cin * KS * KS;
  const size_t per_sample = (size_t)Cout * Ho * Wo * sizeof(T);
  if (wcount * sizeof(T) > CONST_BYTES || per_sample > SMEM_CAP / 2) return 0;
  int SPB = (256 + IPS - 1) / IPS;
  const int cap = (int)((SMEM_CAP / 2) / per_sample);
  if (SPB > cap) SPB = cap;
  if (SPB < 1) SPB = 1;
  const size_t smem = (size_t)SPB * per_sample;
  auto kern = conv_dgrad_tiled_kernel<T, KS>;
  int rc = set_smem(kern, smem);
  if (rc != PG_OK) return rc;
  rc = upload_filters<T>(ctx, w, wcount);
  if (rc != PG_OK) return rc;
  int64_t groups = pg_cdiv(N, SPB);
  int blocks = (int)(groups < 4 * (int64_t)ctx->num_sms ? groups : 4 * (int64_t)ctx->num_sms);
  kern<<<blocks, 256, smem, ctx->stream>>>(dy, dx, (int)N, Cin, H, W, Cout, SPB, SEG);
  ctx->launches++;
  PG_CHECK_CUDA(cudaGetLastError());
  return 1;
}

template <typename T, int KS>
int wgrad_launch(pg_ctx* ctx, const T* x, const T* dy, T* dw, T* db, int64_t N, int Cin, int H, int W, int Cout) {
  const int Ho = H - KS + 1, Wo = W - KS + 1;
  const int nitem = Cin * KS;
  if (nitem > 256 || N <= 0) return 0;
  const size_t per_sample = ((size_t)Cin * H * W + (size_t)Cout * Ho * Wo) * sizeof(T);
  const size_t nred = (size_t)Cout * Cin * KS * KS;
  const size_t red_bytes = (nred + Cout) * sizeof(T);
  if (per_sample + red_bytes > SMEM_CAP / 2) return 0;
  // threads per slot = nitem * RG: fill the block with row-groups first, then with sample slots
  int RG = 256 / nitem;
  if (RG > Ho) RG = Ho;
  if (RG < 1) RG = 1;
  int SPB = 256 / (nitem * RG);
  if (SPB < 1) SPB = 1;
  const int cap = (int)((SMEM_CAP / 2 - red_bytes) / per_sample);
  if (SPB > cap) SPB = cap;
  const size_t smem = (size_t)SPB * per_sample + red_bytes;
  int blocks = (int)(pg_cdiv(N, SPB) < 2 * (int64_t)ctx->num_sms ? pg_cdiv(N, SPB) : 2 * (int64_t)ctx->num_sms);
  int spb_total = (int)(pg_cdiv(pg_cdiv(N, blocks), SPB) * SPB);   // samples per block, multiple of SPB
  blocks = (int)pg_cdiv(N, spb_total);
  int rc = pg_ws_ensure(ctx, (size_t)blocks * (nred + Cout) * sizeof(T));
  if (rc != PG_OK) return rc;
  T* dw_part = (T*)ctx->ws;
  T* db_part = dw_part + (size_t)blocks * nred;
  auto kern = conv_wgrad_tiled_kernel<T, KS>;
  rc = set_smem(kern, smem);
  if (rc != PG_OK) return rc;
  kern<<<blocks, 256, smem, ctx->stream>>>(x, dy, dw_part, db ? db_part : nullptr, (int)N, Cin, H, W, Cout, SPB, RG, spb_total);
  ctx->launches++;
  PG_CHECK_CUDA(cudaGetLastError());
  if (dw) {
    reduce_blocks_kernel<T><<<(unsigned)pg_cdiv((int64_t)nred, 64), 256, 0, ctx->stream>>>(dw_part, dw, (int64_t)nred, blocks);
    ctx->launches++;
  }
  if (db) {
    reduce_blocks_kernel<T><<<(unsigned)pg_cdiv(Cout, 64), 256, 0, ctx->stream>>>(db_part, db, Cout, blocks);
    ctx->launches++;
  }
  PG_CHECK_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace

// Return 1 when the tiled kernel handled the call, 0 when the shape is outside its envelope (caller
// falls back to the generic kernel), negative on error.
template <typename T>
int pg_conv_fwd_tiled(pg_ctx* ctx, const T* x, const T* w, const T* b, T* y, int64_t N, int Cin, int H, int W, int Cout, int kH, int kW, int act) {
  if (kH != kW || N >= (1ll << 31) / ((int64_t)Cin * H * W + 1)) return 0;
  switch (kH) {
    case 3: return (W - 2 >= 16) ? fwd_launch<T, 3, 8>(ctx, x, w, b, y, N, Cin, H, W, Cout, act) : fwd_launch<T, 3, 4>(ctx, x, w, b, y, N, Cin, H, W, Cout, act);
    case 5: return (W - 4 >= 16) ? fwd_launch<T, 5, 8>(ctx, x, w, b, y, N, Cin, H, W, Cout, act) : fwd_launch<T, 5, 4>(ctx, x, w, b, y, N, Cin, H, W, Cout, act);
  }
  return 0;
}
template <typename T>
int pg_conv_dgrad_tiled(pg_ctx* ctx, const T* dy, const T* w, T* dx, int64_t N, int Cin, int H, int W, int Cout, int kH, int kW) {
  if (kH != kW || N >= (1ll << 31) / ((int64_t)Cin * H * W + (int64_t)Cout * H * W + 1)) return 0;
  switch (kH) {
    case 3: return dgrad_launch<T, 3>(ctx, dy, w, dx, N, Cin, H, W, Cout);
    case 5: return dgrad_launch<T, 5>(ctx, dy, w, dx, N, Cin, H, W, Cout);
  }
  return 0;
}
template <typename T>
int pg_conv_wgrad_tiled(pg_ctx* ctx, const T* x, const T* dy, T* dw, T* db, int64_t N, int Cin, int H, int W, int Cout, int kH, int kW) {
  if (kH != kW || N >= (1ll << 31) / ((int64_t)Cin * H * W + (int64_t)Cout * H * W + 1)) return 0;
  switch (kH) {
    case 3: return wgrad_launch<T, 3>(ctx, x, dy, dw, db, N, Cin, H, W, Cout);
    case 5: return wgrad_launch<T, 5>(ctx, x, dy, dw, db, N, Cin, H, W, Cout);
  }
  return 0;
}

template int pg_conv_fwd_tiled<float>(pg_ctx*, const float*, const float*, const float*, float*, int64_t, int, int, int, int, int, int, int);
template int pg_conv_fwd_tiled<double>(pg_ctx*, const double*, const double*, const double*, double*, int64_t, int, int, int, int, int, int, int);
template int pg_conv_dgrad_tiled<float>(pg_ctx*, const float*, const float*, float*, int64_t, int, int, int, int, int, int);
template int pg_conv_dgrad_tiled<double>(pg_ctx*, const double*, const double*, double*, int64_t, int, int, int, int, int, int);
template int pg_conv_wgrad_tiled<float>(pg_ctx*, const float*, const float*, float*, float*, int64_t, int, int, int, int, int, int);
template int pg_conv_wgrad_tiled<double>(pg_ctx*, const double*, const double*, double*, double*, int64_t, int, int, int, int, int, int);
