// mlp_step.cu — one training step of a small MLP in ONE launch: the reference's own example
// (examples/01_mnist_mlp.jl: 784-100-10 at a minibatch of 128, 41 MFLOP per step) is launch-latency-bound as a
// sequence of kernels — 13 dependent launches, 58 us with CUDA-graph replay.  Here a cooperative grid (one CTA per SM)
// walks the phases of the step with grid-wide barriers in between:
//     forward   H_l = relu(H_{l-1} W_l + b_l)  (src/layers.jl:28-34), logits for the last layer
//     loss      P = softmax(Z), L = mean_b(-sum_c Y log(P + eps)), dZ  (src/layers.jl:87, src/losses.jl:7-10)
//     backward  dW_l = H_{l-1}^T dH_l, db_l = sum_b dH_l, dH_{l-1} = (dH_l W_l^T) .* [H_{l-1} > 0]
//     update    GradientDescent (gradient_descent.jl:15-17) or Adam (adam.jl:32-41) over the parameter arena
// Float32 throughout (the 1e-5 contract of the FP32 mode); every sum has a fixed order: results are bit-identical run to run.
// A GEMM phase hands out 16 x 8 output tiles; a CTA stages the tile's A (16 x K) and B (K x 8) panels in shared memory
// with coalesced loads, two threads share one output (K halves).  Envelope: every reduction length (layer widths and
// the batch) <= 1536 (shared memory), 2-4 Linear layers.
#include <cooperative_groups.h>

#include "common.cuh"
#include "optim.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int TM = 16, TN = 8, THREADS = 256, MAX_L = 4, MAX_K = 1536;
constexpr int AP = TM + 1, BP = TN + 1;          // padded panel pitches (bank-conflict-free staging)

struct MlpArgs {
  int L;
  int B;
  int dims[MAX_L + 1];
  const float* x;
  const float* y;
  float* W[MAX_L]; float* b[MAX_L];
  float* gW[MAX_L]; float* gb[MAX_L];
  float* H[MAX_L];                 // H[l]: output of layer l ([B][dims[l+1]]; the last one holds logits, then dZ)
  float* dH[MAX_L];                // dH[l]: gradient w.r.t. H[l] (l < L-1)
  double* row_loss;                // [B]
  float* loss_out;
  double inv_batch;
  int opt;                         // 0 none, 1 gradient descent, 2 Adam
  int flags;                       // PG_SCALAR_F64 / PG_MATH_FAST
  float* p; float* g; float* m; float* v; float* mh; float* vh;
  long long n;
  double s, b1, b2, eps, c1, c2;
};

enum { EPI_BIAS_RELU = 0, EPI_BIAS = 1, EPI_PLAIN = 2, EPI_MASK = 3 };

// debug build only (-DPG_TC_TRACE, tools/mega_trace.py): globaltimer stamps of CTA 0 at the phase boundaries
#ifdef PG_TC_TRACE
__device__ long long g_mlp_trace[16];
#define MLP_TRACE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_mlp_trace[i] = t_; } } while (0)
#else
#define MLP_TRACE(i) do { } while (0)
#endif

// Shared-memory budget of a CTA (floats) and the panel geometry of a macro tile of mi x ni sub-tiles (16 x 8 outputs each):
// A panel [16 mi][KP] row-major (KP = K rounded up to 8 mod 32: the four rows a warp reads land in different banks, rows stay
// 16-byte aligned), B panel [K4][8 ni + 4] (rows 16-byte aligned), 8 x 128 floats for the exchange of the K-slice partials.
constexpr int SMEM_FLOATS = 50 * 1024;            // 200 KB
__host__ __device__ __forceinline__ int kp_of(int K) { return ((K + 3) & ~3) + ((40 - (((K + 3) & ~3) & 31)) & 31); }
__host__ __device__ __forceinline__ int k4_of(int K) { return (K + 3) & ~3; }
__host__ __device__ __forceinline__ bool fits(int K, int mi, int ni) { return (long long)TM * mi * kp_of(K) + 4ll * k4_of(K) + (long long)k4_of(K) * (TN * ni + 4) + 1024 <= SMEM_FLOATS; }
// grow the macro tile until the GEMM is at most `budget` work items (one round of the grid) or shared memory is full
__device__ __forceinline__ void pick_macro(int M, int N, int K, int budget, int& mi, int& ni) {
  mi = 1; ni = 1;
  while (true) {
    const int tm = (M + TM * mi - 1) / (TM * mi), tn = (N + TN * ni - 1) / (TN * ni);
    if (tm * tn <= budget) return;
    if (tm >= tn && tm > 1 && fits(K, mi + 1, ni)) ++mi;
    else if (tn > 1 && fits(K, mi, ni + 1)) ++ni;
    else if (tm > 1 && fits(K, mi + 1, ni)) ++mi;
    else return;
  }
}

__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool pred) {      // pred false: zero fill
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = pred ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst, const float* src, bool pred) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

// one macro tile of C = A B:  A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn]; panels staged once (coalesced, no
// divisions in the loops), then the 16 x 8 sub-tiles one after the other.  The inner product is shared-memory-bound (one
// wavefront per LDS instruction, broadcast or not), so a thread owns 2 x 2 outputs (two float4 of A along k and four float2 of
// B feed 16 FMAs: one wavefront per FMA instruction instead of two) and the eight warps split K; the partials are summed in a
// fixed order through shared memory.
template <int EPI>
__device__ void macro_gemm(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ Bm, long long sbk, long long sbn,
                           int M, int N, int K, int m0, int n0, int mi, int ni, float* __restrict__ C, int ldc, const float* __restrict__ aux, float* smem,
                           [[maybe_unused]] int trace_slot = -1) {
  const int KP = kp_of(K), K4 = k4_of(K), rows = TM * mi, cols = TN * ni, BPn = cols + 4, RP = rows + 4;
  const bool a_kmajor = sak != 1;               // transposed operand: panel stored [K4][RP] instead of [rows][KP]
  float* As = smem;                             // [rows][KP] or [K4][RP]
  float* Bs = smem + (a_kmajor ? (size_t)K4 * RP : (size_t)rows * KP);         // [K4][BPn]
  float* red = Bs + (size_t)K4 * BPn;           // [8][128]
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  // cp.async: every copy of the panels is in flight at once (a load -> store loop ran one latency per element: 9 us for 75 KB)
  if (sak == 1) {                               // rows of A are contiguous in k: warp per row
    const bool vec = (K & 3) == 0 && (sam & 3) == 0 && ((uintptr_t)A & 15) == 0;
    for (int m = warp; m < rows; m += THREADS / 32) {
      const bool live = m0 + m < M;
      const float* src = A + (live ? (long long)(m0 + m) * sam : 0);
      float* dst = As + (size_t)m * KP;
      if (vec) {
        for (int k = 4 * lane; k < K4; k += 128) cp_async16(dst + k, src + k, live);
      } else {
        for (int k = lane; k < K4; k += 32) cp_async4(dst + k, src + (k < K ? k : 0), live && k < K);
      }
    }
  } else {                                      // A is contiguous in m (a transposed operand): k-major panel [K4][rows + 4], 16-byte copies along m
    if ((sak & 3) == 0 && m0 + rows <= M && ((uintptr_t)A & 15) == 0 && sam == 1) {
      const int groups = rows >> 2, per = THREADS / groups, kk = t / groups, g = t - kk * groups;
      if (kk < per)
        for (int k = kk; k < K4; k += per) cp_async16(As + (size_t)k * RP + 4 * g, A + (k < K ? (long long)k * sak + m0 + 4 * g : 0), k < K);
    } else {
      for (int k = warp; k < K4; k += THREADS / 32)
        for (int m = lane; m < rows; m += 32) {
          const bool ok = m0 + m < M && k < K;
          cp_async4(As + (size_t)k * RP + m, A + (ok ? (long long)(m0 + m) * sam + (long long)k * sak : 0), ok);
        }
    }
  }
  if (sbn == 1) {                               // rows of B are contiguous in n
    if ((sbk & 3) == 0 && (n0 & 3) == 0 && n0 + cols <= N && ((uintptr_t)Bm & 15) == 0) {
      const int groups = cols >> 2, per = THREADS / groups, kk = t / groups, g = t - kk * groups;
      if (kk < per)
        for (int k = kk; k < K4; k += per) cp_async16(Bs + (size_t)k * BPn + 4 * g, Bm + (k < K ? (long long)k * sbk + n0 + 4 * g : 0), k < K);
    } else {
      const int per = THREADS / cols, kk = t / cols, n = t - kk * cols;
      if (kk < per)
        for (int k = kk; k < K4; k += per) {
          const bool ok = n0 + n < N && k < K;
          cp_async4(Bs + (size_t)k * BPn + n, Bm + (ok ? (long long)k * sbk + (n0 + n) : 0), ok);
        }
    }
  } else {                                      // B is contiguous in k
    for (int n = 0; n < cols; ++n)
      for (int k = t; k < K4; k += THREADS) {
        const bool ok = n0 + n < N && k < K;
        cp_async4(Bs + (size_t)k * BPn + n, Bm + (ok ? (long long)k * sbk + (long long)(n0 + n) * sbn : 0), ok);
      }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  if (trace_slot >= 0) MLP_TRACE(trace_slot);
  const int rp = lane >> 2, cp = lane & 3;                       // rows rp, rp + 8; columns 2 cp, 2 cp + 1 of the sub-tile
  const int ks = (((K4 + 7) >> 3) + 3) & ~3;                     // K slice of this warp
  const int k0 = warp * ks, k1 = min(K4, k0 + ks);
  for (int si = 0; si < mi; ++si) {
    if (m0 + si * TM >= M) break;
    for (int sj = 0; sj < ni; ++sj) {
      if (n0 + sj * TN >= N) break;
      const float* bc = Bs + sj * TN + 2 * cp;
      float c00 = 0.0f, c01 = 0.0f, c10 = 0.0f, c11 = 0.0f;
      if (a_kmajor) {
        const float* ac = As + si * TM + rp;
#pragma unroll 4
        for (int k = k0; k < k1; ++k) {
          const float lo = ac[(size_t)k * RP], hi = ac[(size_t)k * RP + 8];
          const float2 b0 = *reinterpret_cast<const float2*>(bc + (size_t)k * BPn);
          c00 = fmaf(lo, b0.x, c00); c01 = fmaf(lo, b0.y, c01); c10 = fmaf(hi, b0.x, c10); c11 = fmaf(hi, b0.y, c11);
        }
      } else {
        const float* a_lo = As + (size_t)(si * TM + rp) * KP;
        const float* a_hi = a_lo + (size_t)8 * KP;
#pragma unroll 2
        for (int k = k0; k < k1; k += 4) {
          const float4 lo = *reinterpret_cast<const float4*>(a_lo + k);
          const float4 hi = *reinterpret_cast<const float4*>(a_hi + k);
          const float2 b0 = *reinterpret_cast<const float2*>(bc + (size_t)k * BPn);
          const float2 b1 = *reinterpret_cast<const float2*>(bc + (size_t)(k + 1) * BPn);
          const float2 b2 = *reinterpret_cast<const float2*>(bc + (size_t)(k + 2) * BPn);
          const float2 b3 = *reinterpret_cast<const float2*>(bc + (size_t)(k + 3) * BPn);
          c00 = fmaf(lo.x, b0.x, c00); c01 = fmaf(lo.x, b0.y, c01); c10 = fmaf(hi.x, b0.x, c10); c11 = fmaf(hi.x, b0.y, c11);
          c00 = fmaf(lo.y, b1.x, c00); c01 = fmaf(lo.y, b1.y, c01); c10 = fmaf(hi.y, b1.x, c10); c11 = fmaf(hi.y, b1.y, c11);
          c00 = fmaf(lo.z, b2.x, c00); c01 = fmaf(lo.z, b2.y, c01); c10 = fmaf(hi.z, b2.x, c10); c11 = fmaf(hi.z, b2.y, c11);
          c00 = fmaf(lo.w, b3.x, c00); c01 = fmaf(lo.w, b3.y, c01); c10 = fmaf(hi.w, b3.x, c10); c11 = fmaf(hi.w, b3.y, c11);
        }
      }
      float* rw = red + warp * 128;
      rw[rp * 8 + 2 * cp] = c00; rw[rp * 8 + 2 * cp + 1] = c01;
      rw[(rp + 8) * 8 + 2 * cp] = c10; rw[(rp + 8) * 8 + 2 * cp + 1] = c11;
      __syncthreads();
      if (t < 128) {
        float acc = red[t];
#pragma unroll
        for (int w = 1; w < THREADS / 32; ++w) acc += red[w * 128 + t];
        const int gm = m0 + si * TM + (t >> 3), gn = n0 + sj * TN + (t & 7);
        if (gm < M && gn < N) {
          if (EPI == EPI_BIAS_RELU) { acc += aux[gn]; acc = acc < 0.0f ? 0.0f : acc; }
          else if (EPI == EPI_BIAS) acc += aux[gn];
          else if (EPI == EPI_MASK) acc = aux[(long long)gm * ldc + gn] > 0.0f ? acc : 0.0f;
          C[(long long)gm * ldc + gn] = acc;
        }
      }
      __syncthreads();                          // red is reused by the next sub-tile, the panels by the CTA's next macro tile
    }
  }
}

// db[n] = sum_b dY[b][n] for 8 columns starting at n0: warp c sums column n0 + c, rows strided over the lanes, fixed tree
__device__ void colsum8(const float* __restrict__ dY, int B, int N, int n0, float* __restrict__ db) {
  const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = n0 + c;
  float s = 0.0f;
  if (n < N)
    for (int b = lane; b < B; b += 32) s += dY[(long long)b * N + n];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0 && n < N) db[n] = s;
}

// softmax + cross-entropy of rows [row0, row1) with stride `step` CTAs-worth of warps: one warp per sample (loss_kernels.cu's
// arithmetic); the logits buffer becomes dZ, the sample's loss goes to row_loss
__device__ void softmax_ce_rows(const MlpArgs& a, int first, int end, int cta_stride) {
  const int C = a.dims[a.L];
  float* Z = a.H[a.L - 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float eps = Eps<float>::v();
  const float invB = (float)a.inv_batch;
  // fused form: rows first .. end-1 by this CTA's warps; grid form: first = blockIdx.x, rows blockIdx.x*8 + warp, stride G*8
  const int start = cta_stride == 1 ? first + warp : first * (THREADS / 32) + warp;
  const int stride = cta_stride * (THREADS / 32);
  for (int row = start; row < end; row += stride) {
    float* z = Z + (long long)row * C;
    const float* y = a.y + (long long)row * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.0f;
    for (int c = lane; c < C; c += 32) se += expf(z[c] - mx);
    se = warp_sum(se);
    float dot = 0.0f, lrow = 0.0f;
    for (int c = lane; c < C; c += 32) {
      const float p = expf(z[c] - mx) / se, pe = p + eps;
      lrow -= y[c] * logf(pe);
      dot += (-y[c] * invB / pe) * p;
    }
    lrow = warp_sum(lrow);
    dot = warp_sum(dot);
    for (int c = lane; c < C; c += 32) {
      const float p = expf(z[c] - mx) / se;
      const float dp = -y[c] * invB / (p + eps);
      z[c] = p * (dp - dot);
    }
    if (lane == 0) a.row_loss[row] = (double)lrow;
  }
}

__global__ void __launch_bounds__(THREADS, 1) mlp_step_kernel(const MlpArgs a) {
  extern __shared__ __align__(16) float smem[];
  cg::grid_group grid = cg::this_grid();
  const int L = a.L, B = a.B, G = gridDim.x;
  MLP_TRACE(0);
  // ---- forward; the last layer's macro tiles span all classes, so the CTA that computed a block of logit rows runs their
  //      softmax + cross-entropy right away (no grid barrier in between); the logits buffer becomes dZ ----
  for (int l = 0; l < L; ++l) {
    const float* in = l == 0 ? a.x : a.H[l - 1];
    const int K = a.dims[l], N = a.dims[l + 1];
    const bool last = l + 1 == L;
    int mi, ni;
    pick_macro(B, N, K, G, mi, ni);
    const bool fuse_loss = last && fits(K, 1, (N + TN - 1) / TN);
    if (fuse_loss) { ni = (N + TN - 1) / TN; mi = 1; while ((B + TM * mi - 1) / (TM * mi) > G && fits(K, mi + 1, ni)) ++mi; }
    const int tn = (N + TN * ni - 1) / (TN * ni), nt = ((B + TM * mi - 1) / (TM * mi)) * tn;
    for (int it = blockIdx.x; it < nt; it += G) {
      const int m0 = (it / tn) * TM * mi, n0 = (it % tn) * TN * ni;
      if (!last) macro_gemm<EPI_BIAS_RELU>(in, K, 1, a.W[l], N, 1, B, N, K, m0, n0, mi, ni, a.H[l], N, a.b[l], smem, l == 0 ? 5 : -1);
      else macro_gemm<EPI_BIAS>(in, K, 1, a.W[l], N, 1, B, N, K, m0, n0, mi, ni, a.H[l], N, a.b[l], smem);
      if (fuse_loss) softmax_ce_rows(a, m0, min(B, m0 + TM * mi), 1);
    }
    if (last && !fuse_loss) {
      MLP_TRACE(9);
      grid.sync();
      softmax_ce_rows(a, blockIdx.x, B, G);
    }
    MLP_TRACE(1 + 2 * l);
    grid.sync();
    MLP_TRACE(2 + 2 * l);
  }
  // ---- backward ----
  for (int l = L - 1; l >= 0; --l) {
    const float* in = l == 0 ? a.x : a.H[l - 1];                  // [B][K]
    const float* dY = l == L - 1 ? a.H[L - 1] : a.dH[l];          // [B][N]
    const int K = a.dims[l], N = a.dims[l + 1];
    const int n_b = (N + 7) / 8, n_l = (l == L - 1) ? 1 : 0;
    int mx = 1, nx = 1, mw, nw, n_x = 0, tn_x = 1;
    if (l > 0) {
      pick_macro(B, K, N, max(G / 2, 1), mx, nx);
      tn_x = (K + TN * nx - 1) / (TN * nx);
      n_x = ((B + TM * mx - 1) / (TM * mx)) * tn_x;
    }
    pick_macro(K, N, B, max(G - n_b - n_x - n_l, G / 4), mw, nw);
    const int tn_w = (N + TN * nw - 1) / (TN * nw), n_w = ((K + TM * mw - 1) / (TM * mw)) * tn_w;
    const int total = n_w + n_b + n_x + n_l;
    for (int it = blockIdx.x; it < total; it += G) {
      if (it < n_w) {            // dW[k][n] = sum_b in[b][k] dY[b][n]
        const int m0 = (it / tn_w) * TM * mw, n0 = (it % tn_w) * TN * nw;
        macro_gemm<EPI_PLAIN>(in, 1, K, dY, N, 1, K, N, B, m0, n0, mw, nw, a.gW[l], N, nullptr, smem, l == 0 ? 7 : -1);
      } else if (it < n_w + n_b) {
        colsum8(dY, B, N, (it - n_w) * 8, a.gb[l]);
      } else if (it < n_w + n_b + n_x) {   // dX[b][k] = sum_n dY[b][n] W[k][n], masked by relu'(in)
        const int j = it - n_w - n_b;
        const int m0 = (j / tn_x) * TM * mx, n0 = (j % tn_x) * TN * nx;
        macro_gemm<EPI_MASK>(dY, N, 1, a.W[l], 1, N, B, K, N, m0, n0, mx, nx, a.dH[l - 1], K, in, smem);
      } else {                   // the step's loss: row losses summed in a fixed order by one CTA
        double s = 0.0;
        for (int r = threadIdx.x; r < B; r += THREADS) s += a.row_loss[r];
        double* sd = reinterpret_cast<double*>(smem);
        sd[threadIdx.x] = s;
        __syncthreads();
        for (int o = THREADS / 2; o > 0; o >>= 1) {
          if ((int)threadIdx.x < o) sd[threadIdx.x] += sd[threadIdx.x + o];
          __syncthreads();
        }
        if (threadIdx.x == 0) *a.loss_out = (float)(sd[0] * a.inv_batch);
        __syncthreads();
      }
    }
    MLP_TRACE(11 + (L - 1 - l));
    if (l > 0 || a.opt != 0) grid.sync();
  }
  MLP_TRACE(15);
  // ---- optimizer over the (padded) parameter arena ----
  if (a.opt == 0) return;
  const long long stride = (long long)G * THREADS;
  for (long long i = (long long)blockIdx.x * THREADS + threadIdx.x; i < a.n; i += stride) {
    if (a.opt == 1) {
      // gradient_descent.jl:15-17: variable .-= stepsize .* gradient (Float64-promoted when the stepsize is a Float64)
      if (a.flags & PG_SCALAR_F64) a.p[i] = (float)Ar<double>::sub((double)a.p[i], Ar<double>::mul(a.s, (double)a.g[i]));
      else a.p[i] = Ar<float>::sub(a.p[i], Ar<float>::mul((float)a.s, a.g[i]));
    } else {
      float in[4] = {a.p[i], a.m[i], a.v[i], a.g[i]}, out[5];
      if (a.flags & PG_MATH_FAST)
        AdamFastF{(float)a.s, (float)a.b1, (float)(1.0 - a.b1), (float)a.b2, (float)(1.0 - a.b2), (float)(1.0 / a.c1), (float)(1.0 / a.c2), (float)a.eps, nullptr}(in, out);
      else if (a.flags & PG_SCALAR_F64)
        AdamF<float, double>{a.s, a.b1, 1.0 - a.b1, a.b2, 1.0 - a.b2, a.c1, a.c2, a.eps, nullptr}(in, out);
      else
        AdamF<float, float>{(float)a.s, (float)a.b1, 1.0f - (float)a.b1, (float)a.b2, 1.0f - (float)a.b2, (float)a.c1, (float)a.c2, (float)a.eps, nullptr}(in, out);
      a.p[i] = out[0]; a.m[i] = out[1]; a.v[i] = out[2];
      if (a.mh) a.mh[i] = out[3];
      if (a.vh) a.vh[i] = out[4];
    }
  }
}

}  // namespace

extern "C" {

#ifdef PG_TC_TRACE
int pg_debug_mlp_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_mlp_trace, sizeof(g_mlp_trace)) == cudaSuccess ? PG_OK : PG_ERR_CUDA;
}
#endif

int pg_mlp_step(pg_ctx* ctx, const pg_mlp_desc* d) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(d != nullptr, "null descriptor");
  PG_REQUIRE(d->n_layers >= 2 && d->n_layers <= MAX_L, "2 to 4 Linear layers");
  PG_REQUIRE(d->batch >= 1 && d->batch <= MAX_K, "batch must be in [1, 1536]");
  PG_REQUIRE(d->x && d->labels && d->loss_out, "null pointer");
  PG_REQUIRE(d->opt >= 0 && d->opt <= 2, "optimizer: 0 none, 1 gradient descent, 2 Adam");
  MlpArgs a{};
  a.L = d->n_layers; a.B = (int)d->batch;
  int kmax = a.B;
  size_t act = 0;
  for (int l = 0; l <= a.L; ++l) {
    PG_REQUIRE(d->dims[l] >= 1 && d->dims[l] <= MAX_K, "layer widths must be in [1, 1536]");
    a.dims[l] = (int)d->dims[l];
    if (a.dims[l] > kmax) kmax = a.dims[l];
  }
  for (int l = 0; l < a.L; ++l) {
    PG_REQUIRE(d->W[l] && d->b[l] && d->gW[l] && d->gb[l], "null parameter / gradient pointer");
    a.W[l] = (float*)d->W[l]; a.b[l] = (float*)d->b[l]; a.gW[l] = (float*)d->gW[l]; a.gb[l] = (float*)d->gb[l];
    act += (size_t)a.B * a.dims[l + 1] * (l + 1 < a.L ? 2 : 1);
  }
  // scratch: activations, their gradients, per-sample losses
  const size_t bytes = act * sizeof(float) + 256 + (size_t)a.B * sizeof(double);
  int rc = pg_ws_ensure(ctx, bytes);
  if (rc != PG_OK) return rc;
  float* f = (float*)ctx->ws;
  for (int l = 0; l < a.L; ++l) {
    a.H[l] = f; f += (size_t)a.B * a.dims[l + 1];
    if (l + 1 < a.L) { a.dH[l] = f; f += (size_t)a.B * a.dims[l + 1]; }
  }
  a.row_loss = (double*)(((uintptr_t)f + 255) & ~(uintptr_t)255);
  a.x = (const float*)d->x; a.y = (const float*)d->labels; a.loss_out = (float*)d->loss_out; a.inv_batch = d->loss_scale;
  a.opt = d->opt; a.flags = d->flags;
  a.p = (float*)d->arena_p; a.g = (float*)d->arena_g; a.m = (float*)d->arena_m; a.v = (float*)d->arena_v;
  a.mh = (float*)d->arena_m_hat; a.vh = (float*)d->arena_v_hat; a.n = d->arena_n;
  a.s = d->stepsize; a.b1 = d->beta1; a.b2 = d->beta2; a.eps = d->eps;
  if (a.opt != 0) PG_REQUIRE(a.p && a.g && a.n > 0, "optimizer needs the parameter and gradient arenas");
  if (a.opt == 2) {
    PG_REQUIRE(a.m && a.v && d->it >= 1, "Adam needs m, v and it >= 1 (adam.jl:28)");
    if ((a.flags & (PG_SCALAR_F64 | PG_MATH_FAST)) != 0) { a.c1 = 1.0 - pow(d->beta1, (double)d->it); a.c2 = 1.0 - pow(d->beta2, (double)d->it); }
    else { a.c1 = 1.0f - powf((float)d->beta1, (float)d->it); a.c2 = 1.0f - powf((float)d->beta2, (float)d->it); }   // Float32 hyper-parameters
  }
  (void)kmax;
  const size_t smem = (size_t)SMEM_FLOATS * sizeof(float);
  static PgOnce smem_set;
  if (smem_set.need(ctx->device)) {
    PG_CHECK_CUDA(cudaFuncSetAttribute(mlp_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  void* args[] = {(void*)&a};
  PG_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)mlp_step_kernel, dim3((unsigned)ctx->num_sms), dim3(THREADS), args, smem, ctx->stream));
  ctx->launches++;
  return PG_OK;
}

}  // extern "C"
