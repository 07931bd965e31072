// conv_fused.cu — Conv + bias + ReLU + 2x2 max-pool as ONE forward kernel and ONE pair of backward kernels, for the
// small-channel image layers of the reference's conv example (examples/02_mnist_conv.jl:27-33: Conv((5,5), 1=>6), relu,
// maxpool (2,2), Conv((5,5), 6=>16), relu, maxpool (2,2); src/layers.jl:59,81,93).
//
// These layers move ~10 flop per byte: they are bound by HBM traffic and by instruction issue,
// not by the tensor pipe, so they are NOT reshaped into GEMMs.  What matters here:
//   * HBM traffic: the un-pooled conv output (13.8 KB of the 26 KB a LeNet sample moves in the forward pass) is never
//     written.  The forward kernel keeps a 2x2 window x all output channels in registers, applies bias, max and ReLU
//     there and stores the pooled activation plus one byte per pooled element (which of the four positions won).
//     The backward kernels rebuild the (sparse) conv-output gradient in shared memory from the pooled gradient, that
//     byte and the sign of the pooled value, so max-pool routing, ReLU' and the conv pullbacks are one pass as well.
//   * Issue rate: shared-memory operand traffic is amortised by register tiles — forward: 4 positions x Cout
//     accumulators per thread, weights read as 128-bit broadcasts; weight gradient: a thread owns an input row in
//     registers and K x Cout accumulators (sliding window along the row); data gradient: 4 positions x Cin per thread.
//   * Determinism: the weight/bias gradient is accumulated per CTA in registers over all of its samples, written as
//     per-CTA partials and summed in a fixed order by a second tiny kernel — no floating-point atomics.
// NNlib semantics: true convolution (kernel flipped in both dims), stride 1, no padding; the pooled
// gradient goes to the FIRST maximal element of a window scanning W fastest; relu'(0) = 0.
// Float32 only; geometries are compile-time (see DISPATCH at the bottom); everything else uses conv_tiled.cu.
#include "common.cuh"

#ifndef PG_WGRAD_MINB
#define PG_WGRAD_MINB 1
#endif

namespace {

template <int CIN_, int COUT_, int K_, int H_, int W_>
struct Geo {
  static constexpr int CIN = CIN_, COUT = COUT_, K = K_, H = H_, W = W_;
  static constexpr int HO = H - K + 1, WO = W - K + 1, HP = HO / 2, WP = WO / 2;
  static constexpr int COUTP = (COUT + 3) / 4 * 4;            // channel vectors are read as float4
  static constexpr int CINP = (CIN + 3) / 4 * 4;
  static constexpr int WIN = HP * WP;                          // pooling windows per sample
  static constexpr int XS = CIN * H * W;                       // floats per input sample
  static constexpr int YS = COUT * HP * WP;                    // pooled elements per sample
  static_assert(HO % 2 == 0 && WO % 2 == 0, "conv output must be even for the 2x2 pool");
};

// ---------------------------------------------------------------------------------------------------------------
// forward: one thread = one pooling window of one sample, all output channels
// ---------------------------------------------------------------------------------------------------------------
template <typename G>
struct FwdCfg {
  static constexpr int SPB = (256 + G::WIN - 1) / G::WIN;      // samples per CTA iteration
  static constexpr int THREADS = ((SPB * G::WIN + 31) / 32) * 32;
  static constexpr int SMEM_FLOATS = SPB * G::XS + G::CIN * G::K * G::K * G::COUTP + G::COUTP;
};

template <typename G>
__global__ void __launch_bounds__(FwdCfg<G>::THREADS) conv_relu_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                                const float* __restrict__ b, float* __restrict__ y,
                                                                                uint8_t* __restrict__ amax, int64_t N) {
  using C = FwdCfg<G>;
  constexpr int K = G::K, CIN = G::CIN, COUT = G::COUT, COUTP = G::COUTP;
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                                            // [SPB][CIN][H][W]
  float* ws = xs + C::SPB * G::XS;                             // [CIN][K][K][COUTP]  flipped: ws[ci][u][v][co] = w[co][ci][K-1-u][K-1-v]
  float* bs = ws + CIN * K * K * COUTP;
  for (int i = threadIdx.x; i < CIN * K * K * COUTP; i += blockDim.x) {
    const int co = i % COUTP, t = i / COUTP, v = t % K, u = (t / K) % K, ci = t / (K * K);
    ws[i] = co < COUT ? w[((co * CIN + ci) * K + (K - 1 - u)) * K + (K - 1 - v)] : 0.0f;
  }
  for (int i = threadIdx.x; i < COUTP; i += blockDim.x) bs[i] = i < COUT ? b[i] : 0.0f;
  const int slot = threadIdx.x / G::WIN, win = threadIdx.x % G::WIN;
  const int pi = win / G::WP, pj = win % G::WP;
  const bool worker = threadIdx.x < C::SPB * G::WIN;
  for (int64_t n0 = (int64_t)blockIdx.x * C::SPB; n0 < N; n0 += (int64_t)gridDim.x * C::SPB) {
    __syncthreads();
    const int64_t ns = (N - n0 < C::SPB ? N - n0 : C::SPB);
    // coalesced 128-bit copy of the samples of this iteration (a sample is XS floats; XS % 4 == 0 for every geometry here)
    const float4* src = reinterpret_cast<const float4*>(x + n0 * G::XS);
    for (int i = threadIdx.x; i < ns * (G::XS / 4); i += blockDim.x) reinterpret_cast<float4*>(xs)[i] = src[i];
    __syncthreads();
    if (!worker || slot >= ns) continue;
    float acc[4][COUTP];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int c = 0; c < COUTP; ++c) acc[p][c] = 0.0f;
    const float* xb = xs + slot * G::XS + (2 * pi) * G::W + 2 * pj;
#pragma unroll 1
    for (int ci = 0; ci < CIN; ++ci) {
      float patch[K + 1][K + 1];
#pragma unroll
      for (int r = 0; r <= K; ++r)
#pragma unroll
        for (int c = 0; c <= K; ++c) patch[r][c] = xb[ci * G::H * G::W + r * G::W + c];
      const float4* wv = reinterpret_cast<const float4*>(ws + ci * K * K * COUTP);
#pragma unroll
      for (int u = 0; u < K; ++u)
#pragma unroll
        for (int v = 0; v < K; ++v)
#pragma unroll
          for (int c4 = 0; c4 < COUTP / 4; ++c4) {
            const float4 wq = wv[(u * K + v) * (COUTP / 4) + c4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const float xv = patch[u + (p >> 1)][v + (p & 1)];
              acc[p][4 * c4 + 0] = fmaf(xv, wq.x, acc[p][4 * c4 + 0]);
              acc[p][4 * c4 + 1] = fmaf(xv, wq.y, acc[p][4 * c4 + 1]);
              acc[p][4 * c4 + 2] = fmaf(xv, wq.z, acc[p][4 * c4 + 2]);
              acc[p][4 * c4 + 3] = fmaf(xv, wq.w, acc[p][4 * c4 + 3]);
            }
          }
    }
    const int64_t n = n0 + slot;
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
      // first maximal element, W fastest (positions 0..3 = (0,0) (0,1) (1,0) (1,1)); max commutes with + bias and ReLU
      float m = acc[0][co]; int k = 0;
#pragma unroll
      for (int p = 1; p < 4; ++p)
        if (acc[p][co] > m) { m = acc[p][co]; k = p; }
      m += bs[co];
      const int64_t o = (n * COUT + co) * G::WIN + win;
      y[o] = m > 0.0f ? m : 0.0f;
      if (amax != nullptr) amax[o] = (uint8_t)k;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// weight + bias gradient: thread (ci, u, i) holds input row x[ci][i+u][:] in registers and accumulates
// acc[v][co] += x[i+u][j+v] * dy[i][j][co] while j slides along the row; summed over samples in registers, over the
// HO row-threads in shared memory at the end, written as one partial per CTA.
// The conv-output gradient exists only as a shared-memory tile dy[i][j][co]: every pooled element (held in a register by
// the thread that prefetched it) writes its whole 2x2 window — the winner position gets the relu'-masked pooled gradient,
// the other three get zero — so the tile is complete without a zero-fill pass or a second phase.
// Latency: the NEXT sample's global data (input tile, pooled gradient / value / winner) is prefetched into registers
// while the current one is consumed, and the shared-memory tiles are double-buffered: one __syncthreads per sample.
// ---------------------------------------------------------------------------------------------------------------
template <typename G>
struct WgCfg {
  static constexpr int TPS0 = G::CIN * G::K * G::HO;           // (ci, u, i) row-threads of one sample
  // output channels are split over CS thread groups when that keeps a sample within 512 threads: K x COUTP accumulators
  // plus the input row pushed the kernel to 255 registers with spills and one 8-warp CTA per SM (ncu: 12.5 % warps
  // active, FMA pipe 23-32 % active); half the channels per thread fit 128 registers: twice the warps per SM
  static constexpr int CS = (2 * TPS0 <= 512 && G::COUTP % 8 == 0) ? 2 : 1;
  static constexpr int CT = G::COUTP / CS;                     // output channels per thread
  static constexpr int TPS = CS * TPS0;                        // threads per sample slot
  static constexpr int SLOTS = (256 / TPS) > 0 ? (256 / TPS) : 1;
  static constexpr int THREADS = ((SLOTS * TPS + 31) / 32) * 32;
  static constexpr int NW = G::COUT * G::CIN * G::K * G::K;    // filter elements
  static constexpr int DROW = G::WO * G::COUTP + 4;            // padded row pitch of the conv-output gradient tile dy[i][j][co] (floats): rows land in different banks
  static constexpr int DY = G::HO * DROW;
  static constexpr int XV = (G::XS / 4 + TPS - 1) / TPS;       // float4 of the input tile per thread
  static constexpr int SV = (G::YS + TPS - 1) / TPS;           // pooled elements per thread
  static constexpr int BUF = G::XS + DY;                       // one sample's tiles (floats)
  static constexpr int RED = SLOTS * TPS * G::K * CT;          // final cross-thread reduction buffer (floats)
  static constexpr int STAGE = SLOTS * 2 * BUF;
  static constexpr int SMEM_FLOATS = STAGE > RED ? STAGE : RED;
  static constexpr int MINB = (CS == 2 && THREADS <= 256) ? 2 : PG_WGRAD_MINB;   // CTAs per SM the register budget is sized for
};

template <typename G>
__global__ void __launch_bounds__(WgCfg<G>::THREADS, WgCfg<G>::MINB) conv_relu_pool_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                                 const uint8_t* __restrict__ amax, const float* __restrict__ gy,
                                                                                 float* __restrict__ partial, int64_t N) {
  using C = WgCfg<G>;
  constexpr int K = G::K, CIN = G::CIN, COUT = G::COUT, COUTP = G::COUTP, HO = G::HO, WO = G::WO, CT = C::CT;
  extern __shared__ __align__(16) float smem[];
  const int slot = threadIdx.x / C::TPS, t = threadIdx.x % C::TPS;
  const bool worker = threadIdx.x < C::SLOTS * C::TPS;
  const int i = t % HO, u = (t / HO) % K, ci = C::CS > 1 ? (t / (HO * K)) % CIN : t / (HO * K);
  const int cs = C::CS > 1 ? t / (HO * K * CIN) : 0;           // which group of CT channels
  float* base = smem + slot * 2 * C::BUF;
  float acc[K][CT];
  float dbacc[CT];
#pragma unroll
  for (int v = 0; v < K; ++v)
#pragma unroll
    for (int c = 0; c < CT; ++c) acc[v][c] = 0.0f;
#pragma unroll
  for (int c = 0; c < CT; ++c) dbacc[c] = 0.0f;
  // ---- register prefetch of one sample ----
  float4 xr[C::XV];
  float gr[C::SV];
  uint8_t ar[C::SV];
  auto prefetch = [&](int64_t n) {
    const float4* src = reinterpret_cast<const float4*>(x + n * G::XS);
#pragma unroll
    for (int k = 0; k < C::XV; ++k) {
      const int e = t + k * C::TPS;
      xr[k] = e < G::XS / 4 ? __ldg(src + e) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < C::SV; ++k) {
      const int e = t + k * C::TPS;
      if (e < G::YS) {
        const int64_t o = n * G::YS + e;
        gr[k] = __ldg(y + o) > 0.0f ? __ldg(gy + o) : 0.0f;     // relu'(0) = 0
        ar[k] = __ldg(amax + o);
      }
    }
  };
  auto stash = [&](float* buf) {                                // registers -> this sample's shared-memory tiles
#pragma unroll
    for (int k = 0; k < C::XV; ++k) {
      const int e = t + k * C::TPS;
      if (e < G::XS / 4) reinterpret_cast<float4*>(buf)[e] = xr[k];
    }
    float* dy = buf + G::XS;
#pragma unroll
    for (int k = 0; k < C::SV; ++k) {
      const int e = t + k * C::TPS;
      if (e < G::YS) {
        const int win = e % G::WIN, co = e / G::WIN;
        float* d = dy + (2 * (win / G::WP)) * C::DROW + (2 * (win % G::WP)) * COUTP + co;
        const int a = ar[k];
        const float g = gr[k];
        d[0] = a == 0 ? g : 0.0f;
        d[COUTP] = a == 1 ? g : 0.0f;
        d[C::DROW] = a == 2 ? g : 0.0f;
        d[C::DROW + COUTP] = a == 3 ? g : 0.0f;
      }
    }
  };
  int64_t n = (int64_t)blockIdx.x * C::SLOTS + slot;
  const int64_t stride = (int64_t)gridDim.x * C::SLOTS;
  if (worker && n < N) prefetch(n);
  int it = 0;
  for (int64_t n0 = (int64_t)blockIdx.x * C::SLOTS; n0 < N; n0 += stride, n += stride, ++it) {
    float* buf = base + (it & 1) * C::BUF;
    const bool live = worker && n < N;
    if (live) stash(buf);
    __syncthreads();                                             // (the other buffer is still being read by slower warps: double buffering)
    if (live && n + stride < N) prefetch(n + stride);
    if (!live) continue;
    float row[G::W];
#pragma unroll
    for (int c = 0; c < G::W; ++c) row[c] = buf[(ci * G::H + i + u) * G::W + c];
    const float4* dv = reinterpret_cast<const float4*>(buf + G::XS + i * C::DROW) + cs * (CT / 4);
#pragma unroll
    for (int j = 0; j < WO; ++j) {
#pragma unroll
      for (int c4 = 0; c4 < CT / 4; ++c4) {
        const float4 d = dv[j * (COUTP / 4) + c4];
#pragma unroll
        for (int v = 0; v < K; ++v) {
          acc[v][4 * c4 + 0] = fmaf(row[j + v], d.x, acc[v][4 * c4 + 0]);
          acc[v][4 * c4 + 1] = fmaf(row[j + v], d.y, acc[v][4 * c4 + 1]);
          acc[v][4 * c4 + 2] = fmaf(row[j + v], d.z, acc[v][4 * c4 + 2]);
          acc[v][4 * c4 + 3] = fmaf(row[j + v], d.w, acc[v][4 * c4 + 3]);
        }
      }
      asm volatile("" ::: "memory");     // keep the (fully unrolled) tile loads next to their FMAs: hoisting all of them costs > 128 registers
    }
    if (ci == 0 && u == 0) {            // bias gradient: row i of the tile, by the HO threads of (ci = 0, u = 0) of each channel group
#pragma unroll
      for (int j = 0; j < WO; ++j)
#pragma unroll
        for (int c4 = 0; c4 < CT / 4; ++c4) {
          const float4 d = dv[j * (COUTP / 4) + c4];
          dbacc[4 * c4] += d.x; dbacc[4 * c4 + 1] += d.y; dbacc[4 * c4 + 2] += d.z; dbacc[4 * c4 + 3] += d.w;
        }
    }
  }
  // ---- reduce over the HO row-threads and the sample slots (fixed order), one partial per CTA ----
  __syncthreads();
  float* red = smem;                                             // [SLOTS*TPS][K*CT]
  if (worker) {
#pragma unroll
    for (int v = 0; v < K; ++v)
#pragma unroll
      for (int c = 0; c < CT; ++c) red[threadIdx.x * (K * CT) + v * CT + c] = acc[v][c];
  }
  __syncthreads();
  float* out = partial + (int64_t)blockIdx.x * (C::NW + COUT);
  for (int e = threadIdx.x; e < C::NW; e += blockDim.x) {
    // e = ((co*CIN + ci)*K + a)*K + bq with a, bq the UNFLIPPED filter indices: correlation tap (u, v) = (K-1-a, K-1-bq)
    const int bq = e % K, a = (e / K) % K, ci2 = (e / (K * K)) % CIN, co = e / (K * K * CIN);
    const int u2 = K - 1 - a, v2 = K - 1 - bq;
    const int cs2 = co / CT, cl = co % CT;
    float s = 0.0f;
    for (int sl = 0; sl < C::SLOTS; ++sl)
      for (int r = 0; r < HO; ++r) s += red[(sl * C::TPS + ((cs2 * CIN + ci2) * K + u2) * HO + r) * (K * CT) + v2 * CT + cl];
    out[e] = s;
  }
  __syncthreads();
  if (worker && ci == 0 && u == 0) {
#pragma unroll
    for (int c = 0; c < CT; ++c) red[threadIdx.x * CT + c] = dbacc[c];
  }
  __syncthreads();
  for (int co = threadIdx.x; co < COUT; co += blockDim.x) {
    const int cs2 = co / CT, cl = co % CT;
    float s = 0.0f;
    for (int sl = 0; sl < C::SLOTS; ++sl)
      for (int r = 0; r < HO; ++r) s += red[(sl * C::TPS + cs2 * (CIN * K * HO) + r) * CT + cl];     // threads (cs2, ci = 0, u = 0, i = r)
    out[C::NW + co] = s;
  }
}

// sum of the per-CTA partials in a fixed order: one warp per output element (lane-strided partial sums, then a fixed
// shuffle tree) — deterministic
__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ partial, float* __restrict__ dw, float* __restrict__ db,
                                                           int nw, int nb, int nparts) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (e >= nw + nb) return;
  float s = 0.0f;
  for (int p = lane; p < nparts; p += 32) s += partial[(int64_t)p * (nw + nb) + e];
  s = warp_sum(s);
  if (lane == 0) {
    if (e < nw) { if (dw != nullptr) dw[e] = s; }
    else if (db != nullptr) db[e - nw] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// data gradient: dx[ci][p][q] = sum_{co,a,b} dyp[co][p+a][q+b] * w[co][ci][a][b]  (dyp = conv-output gradient zero-bordered
// by K-1; the double flip of a full correlation with the flipped kernel is the original kernel).  One thread = a 2x2
// block of outputs at even (p, q), all input channels.  With K - 1 even, the (K+1) x (K+1) patch of dyp such a thread
// needs is exactly (K+1)/2 x (K+1)/2 WHOLE pooling windows, each holding one non-zero: the patch is built in registers
// from the pooled gradient (masked by relu') and the winner byte, staged channel-first with a zero border of windows —
// the conv-output gradient is never materialised.  Next sample prefetched into registers, tiles double-buffered.
// ---------------------------------------------------------------------------------------------------------------
template <typename G>
struct DgCfg {
  static_assert(G::H % 2 == 0 && G::W % 2 == 0 && (G::K - 1) % 2 == 0, "even input and odd filter");
  static constexpr int TPS = (G::H / 2) * (G::W / 2);
  static constexpr int SLOTS = (256 / TPS) > 0 ? (256 / TPS) : 1;
  static constexpr int THREADS = ((SLOTS * TPS + 31) / 32) * 32;
  static constexpr int BW = (G::K - 1) / 2;                     // border, in windows
  static constexpr int NWIN = (G::K + 1) / 2;                   // windows per patch side
  static constexpr int PHW = G::HP + 2 * BW, PWW = G::WP + 2 * BW;
  static constexpr int GZ = G::COUT * PHW * PWW;                // bordered pooled gradient (floats)
  static constexpr int AZ = (GZ + 3) / 4;                       // bordered winner bytes (as words)
  static constexpr int SV = (G::YS + TPS - 1) / TPS;
  static constexpr int BUF = GZ + AZ;
  static constexpr int SMEM_FLOATS = SLOTS * 2 * BUF + G::COUT * G::K * G::K * G::CINP;
};

template <typename G>
__global__ void __launch_bounds__(DgCfg<G>::THREADS) conv_relu_pool_dgrad_kernel(const float* __restrict__ w, const float* __restrict__ y,
                                                                                 const uint8_t* __restrict__ amax, const float* __restrict__ gy,
                                                                                 float* __restrict__ dx, int64_t N) {
  using C = DgCfg<G>;
  constexpr int K = G::K, CIN = G::CIN, CINP = G::CINP, COUT = G::COUT, NWIN = C::NWIN;
  extern __shared__ __align__(16) float smem[];
  float* ws = smem + C::SLOTS * 2 * C::BUF;                      // [co][a][b][CINP]
  for (int i = threadIdx.x; i < COUT * K * K * CINP; i += blockDim.x) {
    const int ci = i % CINP, t = i / CINP, bq = t % K, a = (t / K) % K, co = t / (K * K);
    ws[i] = ci < CIN ? w[((co * CIN + ci) * K + a) * K + bq] : 0.0f;
  }
  for (int i = threadIdx.x; i < C::SLOTS * 2 * C::BUF; i += blockDim.x) smem[i] = 0.0f;     // borders stay zero for the whole kernel
  __syncthreads();
  const int slot = threadIdx.x / C::TPS, t = threadIdx.x % C::TPS;
  const bool worker = threadIdx.x < C::SLOTS * C::TPS;
  const int p0 = 2 * (t / (G::W / 2)), q0 = 2 * (t % (G::W / 2));
  float* base = smem + slot * 2 * C::BUF;
  float gr[C::SV];
  uint8_t ar[C::SV];
  auto prefetch = [&](int64_t n) {
#pragma unroll
    for (int k = 0; k < C::SV; ++k) {
      const int e = t + k * C::TPS;
      if (e < G::YS) {
        const int64_t o = n * G::YS + e;
        gr[k] = __ldg(y + o) > 0.0f ? __ldg(gy + o) : 0.0f;     // relu'(0) = 0
        ar[k] = __ldg(amax + o);
      }
    }
  };
  auto stash = [&](float* buf) {
    uint8_t* az = reinterpret_cast<uint8_t*>(buf + C::GZ);
#pragma unroll
    for (int k = 0; k < C::SV; ++k) {
      const int e = t + k * C::TPS;
      if (e < G::YS) {
        const int win = e % G::WIN, co = e / G::WIN;
        const int idx = (co * C::PHW + win / G::WP + C::BW) * C::PWW + win % G::WP + C::BW;
        buf[idx] = gr[k];
        az[idx] = ar[k];
      }
    }
  };
  int64_t n = (int64_t)blockIdx.x * C::SLOTS + slot;
  const int64_t stride = (int64_t)gridDim.x * C::SLOTS;
  if (worker && n < N) prefetch(n);
  int it = 0;
  for (int64_t n0 = (int64_t)blockIdx.x * C::SLOTS; n0 < N; n0 += stride, n += stride, ++it) {
    float* buf = base + (it & 1) * C::BUF;
    const bool live = worker && n < N;
    if (live) stash(buf);
    __syncthreads();
    if (live && n + stride < N) prefetch(n + stride);
    if (!live) continue;
    float acc[4][CINP];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int c = 0; c < CINP; ++c) acc[p][c] = 0.0f;
    const uint8_t* az = reinterpret_cast<const uint8_t*>(buf + C::GZ);
    // patch rows p0 .. p0+K of the bordered gradient = conv rows p0-(K-1) .. p0+1 = windows p0/2 - BW .. p0/2 - BW + NWIN-1,
    // i.e. bordered window rows p0/2 .. p0/2 + NWIN-1
#pragma unroll 1
    for (int co = 0; co < COUT; ++co) {
      float patch[K + 1][K + 1];
      const int wb = (co * C::PHW + (p0 >> 1)) * C::PWW + (q0 >> 1);
#pragma unroll
      for (int wr = 0; wr < NWIN; ++wr)
#pragma unroll
        for (int wc = 0; wc < NWIN; ++wc) {
          const float g = buf[wb + wr * C::PWW + wc];
          const int a = az[wb + wr * C::PWW + wc];
          patch[2 * wr][2 * wc] = a == 0 ? g : 0.0f;
          patch[2 * wr][2 * wc + 1] = a == 1 ? g : 0.0f;
          patch[2 * wr + 1][2 * wc] = a == 2 ? g : 0.0f;
          patch[2 * wr + 1][2 * wc + 1] = a == 3 ? g : 0.0f;
        }
      const float4* wv = reinterpret_cast<const float4*>(ws + co * K * K * CINP);
#pragma unroll
      for (int a = 0; a < K; ++a)
#pragma unroll
        for (int bq = 0; bq < K; ++bq)
#pragma unroll
          for (int c4 = 0; c4 < CINP / 4; ++c4) {
            const float4 wq = wv[(a * K + bq) * (CINP / 4) + c4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const float dv = patch[a + (p >> 1)][bq + (p & 1)];
              acc[p][4 * c4 + 0] = fmaf(dv, wq.x, acc[p][4 * c4 + 0]);
              acc[p][4 * c4 + 1] = fmaf(dv, wq.y, acc[p][4 * c4 + 1]);
              acc[p][4 * c4 + 2] = fmaf(dv, wq.z, acc[p][4 * c4 + 2]);
              acc[p][4 * c4 + 3] = fmaf(dv, wq.w, acc[p][4 * c4 + 3]);
            }
          }
    }
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      float* o = dx + (n * CIN + ci) * (G::H * G::W) + p0 * G::W + q0;
      *reinterpret_cast<float2*>(o) = make_float2(acc[0][ci], acc[1][ci]);
      *reinterpret_cast<float2*>(o + G::W) = make_float2(acc[2][ci], acc[3][ci]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
template <typename G>
int run_fwd(pg_ctx* ctx, const float* x, const float* w, const float* b, float* y, uint8_t* amax, int64_t N) {
  using C = FwdCfg<G>;
  const int smem = C::SMEM_FLOATS * 4;
  static PgOnce attr;
  if (attr.need(ctx->device)) PG_CHECK_CUDA(cudaFuncSetAttribute(conv_relu_pool_fwd_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int64_t blocks = pg_cdiv(N, C::SPB);
  const int64_t cap = (int64_t)ctx->num_sms * 4;
  if (blocks > cap) blocks = cap;
  conv_relu_pool_fwd_kernel<G><<<(unsigned)blocks, C::THREADS, smem, ctx->stream>>>(x, w, b, y, amax, N);
  PG_LAUNCHED(ctx);
  return PG_OK;
}

template <typename G>
int run_bwd(pg_ctx* ctx, const float* x, const float* w, const float* y, const uint8_t* amax, const float* gy, float* dw, float* db, float* dx, int64_t N) {
  if (dw != nullptr || db != nullptr) {
    using C = WgCfg<G>;
    const int smem = C::SMEM_FLOATS * 4;
    static PgOnce attr;
    if (attr.need(ctx->device)) PG_CHECK_CUDA(cudaFuncSetAttribute(conv_relu_pool_wgrad_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int64_t blocks = pg_cdiv(N, C::SLOTS);
    const int64_t cap = (int64_t)ctx->num_sms * 2;
    if (blocks > cap) blocks = cap;
    const int nw = C::NW, nb = G::COUT;
    int rc = pg_ws_ensure(ctx, (size_t)blocks * (nw + nb) * sizeof(float));
    if (rc != PG_OK) return rc;
    conv_relu_pool_wgrad_kernel<G><<<(unsigned)blocks, C::THREADS, smem, ctx->stream>>>(x, y, amax, gy, (float*)ctx->ws, N);
    PG_LAUNCHED(ctx);
    sum_partials_kernel<<<(unsigned)pg_cdiv((int64_t)(nw + nb) * 32, 256), 256, 0, ctx->stream>>>((const float*)ctx->ws, dw, db, nw, nb, (int)blocks);
    PG_LAUNCHED(ctx);
  }
  if (dx != nullptr) {
    using C = DgCfg<G>;
    const int smem = C::SMEM_FLOATS * 4;
    static PgOnce attr;
    if (attr.need(ctx->device)) PG_CHECK_CUDA(cudaFuncSetAttribute(conv_relu_pool_dgrad_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int64_t blocks = pg_cdiv(N, C::SLOTS);
    const int64_t cap = (int64_t)ctx->num_sms * 4;
    if (blocks > cap) blocks = cap;
    conv_relu_pool_dgrad_kernel<G><<<(unsigned)blocks, C::THREADS, smem, ctx->stream>>>(w, y, amax, gy, dx, N);
    PG_LAUNCHED(ctx);
  }
  return PG_OK;
}

// the compiled geometries (Cin, Cout, K, H, W): LeNet on 28x28 (examples/02_mnist_conv.jl) and on 32x32 inputs
typedef Geo<1, 6, 5, 28, 28> G_1_6_28;
typedef Geo<6, 16, 5, 12, 12> G_6_16_12;
typedef Geo<1, 6, 5, 32, 32> G_1_6_32;
typedef Geo<3, 6, 5, 32, 32> G_3_6_32;
typedef Geo<6, 16, 5, 14, 14> G_6_16_14;
#define DISPATCH(CALL)                                                                    \
  if (Cin == 1 && Cout == 6 && kH == 5 && H == 28 && W == 28) return CALL(G_1_6_28);      \
  if (Cin == 6 && Cout == 16 && kH == 5 && H == 12 && W == 12) return CALL(G_6_16_12);    \
  if (Cin == 1 && Cout == 6 && kH == 5 && H == 32 && W == 32) return CALL(G_1_6_32);      \
  if (Cin == 3 && Cout == 6 && kH == 5 && H == 32 && W == 32) return CALL(G_3_6_32);      \
  if (Cin == 6 && Cout == 16 && kH == 5 && H == 14 && W == 14) return CALL(G_6_16_14);

}  // namespace

extern "C" {

int pg_conv2d_relu_pool2_supported(int Cin, int H, int W, int Cout, int kH, int kW) {
  if (kH != kW) return 0;
#define YES(G) 1
  DISPATCH(YES)
#undef YES
  return 0;
}

int pg_conv2d_relu_pool2_fwd(pg_ctx* ctx, const void* x, const void* w, const void* b, void* y_pool, void* argmax,
                             int64_t N, int Cin, int H, int W, int Cout, int kH, int kW) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(x && w && b && y_pool, "null pointer");
  PG_REQUIRE(kH == kW, "square filters only");
  if (N <= 0) return PG_OK;
#define FWD(G) run_fwd<G>(ctx, (const float*)x, (const float*)w, (const float*)b, (float*)y_pool, (uint8_t*)argmax, N)
  DISPATCH(FWD)
#undef FWD
  pg_set_error("pg_conv2d_relu_pool2_fwd: geometry not compiled (see pg_conv2d_relu_pool2_supported)");
  return PG_ERR_UNSUPPORTED;
}

int pg_conv2d_relu_pool2_bwd(pg_ctx* ctx, const void* x, const void* w, const void* y_pool, const void* argmax, const void* dy_pool,
                             void* dw, void* db, void* dx, int64_t N, int Cin, int H, int W, int Cout, int kH, int kW) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(y_pool && argmax && dy_pool, "null pointer");
  PG_REQUIRE((dw == nullptr && db == nullptr) || x != nullptr, "dw / db need x");
  PG_REQUIRE(dx == nullptr || w != nullptr, "dx needs w");
  PG_REQUIRE(kH == kW, "square filters only");
  if (N <= 0) return PG_OK;
#define BWD(G) run_bwd<G>(ctx, (const float*)x, (const float*)w, (const float*)y_pool, (const uint8_t*)argmax, (const float*)dy_pool, (float*)dw, (float*)db, (float*)dx, N)
  DISPATCH(BWD)
#undef BWD
  pg_set_error("pg_conv2d_relu_pool2_bwd: geometry not compiled (see pg_conv2d_relu_pool2_supported)");
  return PG_ERR_UNSUPPORTED;
}

}  // extern "C"
