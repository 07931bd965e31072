// conv_tiled.cu — shared-memory-tiled, register-blocked direct convolution kernels for the small-
// channel layers of the reference's conv net (src/layers.jl:59 via NNlib.conv / ∇conv_filter /
// ∇conv_data; examples/02_mnist_conv.jl:25-33: 1->6 and 6->16 channels, 5x5).  These layers move
// ~26 KB of activations per sample for ~1.5 Mflop, i.e. they are bandwidth-/issue-bound, not
// tensor-bound, so the design goal is: read every activation from HBM once
// (whole samples staged in shared memory, fully coalesced), and do >= 4 FMAs per shared-memory load
// (each thread owns a row segment x a channel block in registers).
//   forward     y[co][i][j]   = b[co] + sum_{ci,a,b} x[ci][i+K-1-a][j+K-1-b] w[co][ci][a][b]   (true convolution)
//   data grad   dx[ci][u][v]  = sum_{co,a,b} dyp[co][u+a][v+b] w[co][ci][a][b]                 (dyp = dy zero-padded by K-1)
//   weight grad dw[co][ci][a][b] = sum_{n,i,j} x[ci][i+K-1-a][j+K-1-b] dy[co][i][j];  db[co] = sum dy
// The data-gradient kernel reads the (tiny) filter bank from __constant__ memory with warp-uniform
// indices (measured 431 -> 356 us on conv2 at B = 8192); the same trick did not help the forward kernel
// and an all-output-channel register tile made the weight gradient 3x slower (255 registers, one block
// per SM): profiles/r1_launches_c3_b8192_v6_const.csv.  Generic shapes fall back to conv_kernels.cu.
#include "common.cuh"

namespace {

constexpr int JB = 8;    // outputs per thread along W (forward / weight gradient)
constexpr int COB = 8;   // output channels per thread (forward)
constexpr int DJB = 4;   // outputs per thread along W (data gradient)
constexpr int CIB = 8;   // input channels per thread (data gradient)
constexpr size_t SMEM_CAP = 100 * 1024;
constexpr int CONST_BYTES = 48 * 1024;

// filter bank of the layer being launched (copied device-to-device on the stream before the kernel)
__constant__ __align__(16) unsigned char g_conv_w[CONST_BYTES];


template <typename T, int KS>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? 2 : 1)
conv_fwd_tiled_kernel(const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ b, T* __restrict__ y,
                      int N, int Cin, int H, int W, int Cout, int relu, int SPB, int IPS, int SEG, int woff) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* ws = reinterpret_cast<T*>(smem_raw);        // [Cout][Cin][KS][KS]
  T* xs = ws + woff;                             // [SPB][Cin][H][W]
  const int Ho = H - KS + 1, Wo = W - KS + 1, chw = Cin * H * W, wsz = Cin * KS * KS;
  for (int i = threadIdx.x; i < Cout * wsz; i += blockDim.x) ws[i] = w[i];
  for (int n0 = blockIdx.x * SPB; n0 < N; n0 += gridDim.x * SPB) {
    const int ns = min(SPB, N - n0);
    __syncthreads();
    const T* xg = x + (size_t)n0 * chw;
    for (int i = threadIdx.x; i < ns * chw; i += blockDim.x) xs[i] = xg[i];
    __syncthreads();
    for (int item = threadIdx.x; item < ns * IPS; item += blockDim.x) {
      const int s = item / IPS, r = item - s * IPS;
      const int g = r / (Ho * SEG), r2 = r - g * (Ho * SEG);
      const int i = r2 / SEG, j0 = (r2 - i * SEG) * JB;
      T acc[COB][JB];
#pragma unroll
      for (int c = 0; c < COB; ++c)
#pragma unroll
        for (int j = 0; j < JB; ++j) acc[c][j] = T(0);
      const T* xsamp = xs + s * chw;
      for (int ci = 0; ci < Cin; ++ci) {
#pragma unroll
        for (int a = 0; a < KS; ++a) {
          const T* xr = xsamp + (ci * H + i + KS - 1 - a) * W + j0;
          T row[JB + KS - 1];
#pragma unroll
          for (int q = 0; q < JB + KS - 1; ++q) row[q] = (j0 + q < W) ? xr[q] : T(0);
#pragma unroll
          for (int bb = 0; bb < KS; ++bb) {
#pragma unroll
            for (int c = 0; c < COB; ++c) {
              const int co = min(g * COB + c, Cout - 1);
              const T wv = ws[(co * Cin + ci) * KS * KS + a * KS + bb];
#pragma unroll
              for (int j = 0; j < JB; ++j) acc[c][j] += row[j + KS - 1 - bb] * wv;
            }
          }
        }
      }
      const size_t n = (size_t)(n0 + s);
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

// weight gradient: thread item = (co, ci, a, row-group rg); it accumulates dw[co][ci][a][0..KS) over the
// rows i = rg, rg+RG, ... of every sample of its slot; slots and row-groups are summed at the end.
template <typename T, int KS>
__global__ void __launch_bounds__(256)
conv_wgrad_tiled_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dw_part, T* __restrict__ db_part,
                        int N, int Cin, int H, int W, int Cout, int SPB, int RG, int SB, int samples_per_block) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Ho = H - KS + 1, Wo = W - KS + 1, chw = Cin * H * W, ych = Cout * Ho * Wo;
  const int nitem = Cout * Cin * KS;            // (co, ci, a)
  T* xs = reinterpret_cast<T*>(smem_raw);       // [SB][chw]   SB staged samples per iteration
  T* dys = xs + SB * chw;                       // [SB][ych]
  T* red = dys + SB * ych;                     // [nitem][KS + 1]  (block result: the contributing threads add in a fixed order)
  for (int i = threadIdx.x; i < nitem * (KS + 1); i += blockDim.x) red[i] = T(0);
  // thread -> (slot, item, row-group).  nitem <= 256: one item per thread, RG row-groups, SPB sample slots.
  // nitem > 256: RG = SPB = 1 and each thread owns up to NI = ceil(nitem / 256) items (it, it + 256, ...).
  constexpr int NIMAX = 4;
  const int NI = (nitem + 255) / 256;
  const int per_slot = (NI > 1) ? 256 : nitem * RG;
  const int slot = threadIdx.x / per_slot;      // >= SPB: idle thread
  const int r = threadIdx.x - slot * per_slot;
  const int it0 = (NI > 1) ? r : r / RG;
  const int rg = (NI > 1) ? 0 : r - it0 * RG;
  const bool live = slot < SPB;
  T acc[NIMAX][KS];
  T dbacc[NIMAX];
#pragma unroll
  for (int p = 0; p < NIMAX; ++p) {
    dbacc[p] = T(0);
#pragma unroll
    for (int k = 0; k < KS; ++k) acc[p][k] = T(0);
  }
  const int nb0 = blockIdx.x * samples_per_block;
  const int nb1 = min(N, nb0 + samples_per_block);
  for (int n0 = nb0; n0 < nb1; n0 += SB) {
    const int ns = min(SB, nb1 - n0);
    __syncthreads();
    const T* xg = x + (size_t)n0 * chw;
    for (int i = threadIdx.x; i < ns * chw; i += blockDim.x) xs[i] = xg[i];
    const T* dg = dy + (size_t)n0 * ych;
    for (int i = threadIdx.x; i < ns * ych; i += blockDim.x) dys[i] = dg[i];
    __syncthreads();
    if (live)
    for (int sm_i = slot; sm_i < ns; sm_i += SPB) {     // this slot's samples among the staged ones
#pragma unroll
      for (int p = 0; p < NIMAX; ++p) {
        const int it = it0 + p * 256;
        if (p < NI && it < nitem) {
          const int co = it / (Cin * KS), ci = (it / KS) % Cin, a = it % KS;
          const T* xsamp = xs + sm_i * chw + (ci * H + KS - 1 - a) * W;
          const T* dsamp = dys + sm_i * ych + co * Ho * Wo;
          const bool do_db = (ci == 0 && a == 0);
          for (int i = rg; i < Ho; i += RG) {
            for (int j0 = 0; j0 < Wo; j0 += JB) {
              T dr[JB], xr[JB + KS - 1];
#pragma unroll
              for (int j = 0; j < JB; ++j) dr[j] = (j0 + j < Wo) ? dsamp[i * Wo + j0 + j] : T(0);
#pragma unroll
              for (int q = 0; q < JB + KS - 1; ++q) xr[q] = (j0 + q < W) ? xsamp[i * W + j0 + q] : T(0);
#pragma unroll
              for (int bb = 0; bb < KS; ++bb)
#pragma unroll
                for (int j = 0; j < JB; ++j) acc[p][bb] += xr[j + KS - 1 - bb] * dr[j];
              if (do_db) {
#pragma unroll
                for (int j = 0; j < JB; ++j) dbacc[p] += dr[j];
              }
            }
          }
        }
      }
    }
  }
  __syncthreads();
  // an item has SPB * RG contributing threads (one per sample slot and row-group): they add in turn, in a fixed order, so the
  // block result does not depend on scheduling (no floating-point atomics; with several items per thread every item has one owner)
  const int ncontrib = (NI > 1) ? 1 : SPB * RG;
  const int mine = (NI > 1) ? 0 : slot * RG + rg;
  for (int c = 0; c < ncontrib; ++c) {
    if (live && mine == c) {
#pragma unroll
      for (int p = 0; p < NIMAX; ++p) {
        const int it = it0 + p * 256;
        if (p < NI && it < nitem) {
          const int ci = (it / KS) % Cin, a = it % KS;
#pragma unroll
          for (int k = 0; k < KS; ++k) red[it * (KS + 1) + k] += acc[p][k];
          if (ci == 0 && a == 0) red[it * (KS + 1) + KS] += dbacc[p];
        }
      }
    }
    __syncthreads();
  }
  // block partials: dw_part[block][co][ci][a][b], db_part[block][co]
  for (int i = threadIdx.x; i < nitem * KS; i += blockDim.x) dw_part[(size_t)blockIdx.x * nitem * KS + i] = red[(i / KS) * (KS + 1) + (i % KS)];
  if (db_part != nullptr)
    for (int c = threadIdx.x; c < Cout; c += blockDim.x) db_part[(size_t)blockIdx.x * Cout + c] = red[(c * Cin * KS) * (KS + 1) + KS];
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

template <typename T, int KS>
int fwd_launch(pg_ctx* ctx, const T* x, const T* w, const T* b, T* y, int64_t N, int Cin, int H, int W, int Cout, int act) {
  const int Ho = H - KS + 1, Wo = W - KS + 1;
  const int SEG = (Wo + JB - 1) / JB, G = (Cout + COB - 1) / COB;
  const int IPS = G * Ho * SEG;
  const size_t wfl = ((size_t)Cout * Cin * KS * KS + 3) & ~(size_t)3;
  const size_t per_sample = (size_t)Cin * H * W * sizeof(T);
  if (wfl * sizeof(T) + per_sample > SMEM_CAP) return 0;
  // enough samples per iteration to give every thread one item; small tiles keep several blocks
  // resident per SM so one block's load phase overlaps another's compute phase (measured: staging 32
  // samples with one block per SM was 1.3-3x slower, profiles/r1_launches_c3_b8192_v3_fat.csv)
  int SPB = 256 / IPS;
  if (SPB < 1) SPB = 1;
  const int cap = (int)((SMEM_CAP / 2 - wfl * sizeof(T)) / per_sample);
  if (SPB > cap) SPB = cap;
  if (SPB < 1) SPB = 1;
  const size_t smem = wfl * sizeof(T) + (size_t)SPB * per_sample;
  auto kern = conv_fwd_tiled_kernel<T, KS>;
  int rc = set_smem(kern, smem);
  if (rc != PG_OK) return rc;
  int64_t groups = pg_cdiv(N, SPB);
  int blocks = (int)(groups < 4 * (int64_t)ctx->num_sms ? groups : 4 * (int64_t)ctx->num_sms);
  kern<<<blocks, 256, smem, ctx->stream>>>(x, w, b, y, (int)N, Cin, H, W, Cout, act == PG_ACT_RELU, SPB, IPS, SEG, (int)wfl);
  ctx->launches++;
  PG_CHECK_CUDA(cudaGetLastError());
  return 1;
}

template <typename T>
int upload_filters(pg_ctx* ctx, const T* w, size_t count) {
  PG_CHECK_CUDA(cudaMemcpyToSymbolAsync(g_conv_w, w, count * sizeof(T), 0, cudaMemcpyDeviceToDevice, ctx->stream));
  return PG_OK;
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
  const int nitem = Cout * Cin * KS;
  if (nitem > 1024 || N <= 0) return 0;
  const size_t per_sample = ((size_t)Cin * H * W + (size_t)Cout * Ho * Wo) * sizeof(T);
  const size_t red_bytes = (size_t)nitem * (KS + 1) * sizeof(T);
  if (per_sample + red_bytes > SMEM_CAP) return 0;
  // threads per slot = nitem * RG; fill the block with row-groups first, then with sample slots
  int RG = nitem > 256 ? 1 : 256 / nitem;
  if (RG > Ho) RG = Ho;
  if (RG < 1) RG = 1;
  int SPB = nitem > 256 ? 1 : 256 / (nitem * RG);
  if (SPB < 1) SPB = 1;
  int SB = (int)((SMEM_CAP / 2 - red_bytes) / per_sample);     // staged samples per iteration
  if (SB < 1) SB = 1;
  if (SB < SPB) SPB = SB;
  SB = SPB;                                                    // one sample per slot and iteration (several blocks per SM)
  const size_t smem = (size_t)SB * per_sample + red_bytes;
  int blocks = (int)(pg_cdiv(N, SB) < 4 * (int64_t)ctx->num_sms ? pg_cdiv(N, SB) : 4 * (int64_t)ctx->num_sms);
  int spb_total = (int)(pg_cdiv(pg_cdiv(N, blocks), SB) * SB);   // samples per block, multiple of SB
  blocks = (int)pg_cdiv(N, spb_total);
  const size_t nw = (size_t)nitem * KS;
  int rc = pg_ws_ensure(ctx, (size_t)blocks * (nw + Cout) * sizeof(T));
  if (rc != PG_OK) return rc;
  T* dw_part = (T*)ctx->ws;
  T* db_part = dw_part + (size_t)blocks * nw;
  auto kern = conv_wgrad_tiled_kernel<T, KS>;
  rc = set_smem(kern, smem);
  if (rc != PG_OK) return rc;
  kern<<<blocks, 256, smem, ctx->stream>>>(x, dy, dw_part, db ? db_part : nullptr, (int)N, Cin, H, W, Cout, SPB, RG, SB, spb_total);
  ctx->launches++;
  PG_CHECK_CUDA(cudaGetLastError());
  if (dw) {
    reduce_blocks_kernel<T><<<(unsigned)pg_cdiv((int64_t)nw, 64), 256, 0, ctx->stream>>>(dw_part, dw, (int64_t)nw, blocks);
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
    case 3: return fwd_launch<T, 3>(ctx, x, w, b, y, N, Cin, H, W, Cout, act);
    case 5: return fwd_launch<T, 5>(ctx, x, w, b, y, N, Cin, H, W, Cout, act);
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
