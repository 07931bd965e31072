// tc_conv.cu — BF16 tensor-core path of the Conv layer (src/layers.jl:59-67: NNlib `conv`, i.e. a true
// convolution with a flipped kernel, and its pullbacks ∇conv_filter / ∇conv_data): implicit GEMM on
// tcgen05 with TMEM accumulators.  The im2col matrix is never materialised in HBM: producer warps gather
// fp32 NCHW activations, round them to bf16 and write them straight into the SWIZZLE_128B shared-memory
// operand tiles the UMMA descriptors expect (the same tile formats tc_gemm.cu receives from TMA).
//
//   forward    Y[p][co]   = sum_k  A[p][k] Wf[co][k]     p = (n, oy, ox),  k = (ci, i, j),  A = x[n][ci][oy+i][ox+j]
//                                                         Wf[co][(ci,i,j)] = w[co][ci][kH-1-i][kW-1-j]
//   data grad  dX[p][ci]  = sum_k  G[p][k] Wd[ci][k]     p = (n, iy, ix),  k = (co, i, j),  G = dY[n][co][iy+i-(kH-1)][ix+j-(kW-1)] or 0
//                                                         Wd[ci][(co,i,j)] = w[co][ci][i][j]
//   wt. grad   dW[co][k]  = sum_p  dY[p][co] A[p][k]     (k = Ktot is a column of ones: its sum is db[co])
//
// Forward / data gradient: M = 128 positions per tile, N = 16 channels, K in 64-wide blocks; the (tiny) filter
// matrix lives in shared memory for the whole kernel; an 8-deep ring of TMEM accumulators lets the epilogue
// warps (bias + ReLU, coalesced NCHW stores) run several tiles behind the MMA issuer.  Weight gradient: M = 128
// rows of dY^T (only Cout <= 16 are non-zero), N = the im2col width (<= 256), K = 64 positions per block; every
// CTA accumulates its share of the positions in TMEM and writes its partial dW / db to the workspace; a small kernel
// sums the partials over the CTAs in a fixed order (deterministic: no floating-point atomics).
// Tensors stay fp32 in HBM (26 KB of activations per sample at the LeNet shapes: these layers are not
// bandwidth-bound), products are bf16 x bf16 with fp32 accumulation.  Measured behaviour and what bounds it:
// profiles/r1_tc_conv_ablation.md (the stage round trip producer -> MMA -> commit -> producer, not the gathers).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

#include "tc_ptx.cuh"

constexpr int BLOCK_K = 64, UMMA_K = 16;
constexpr int TILE_BYTES = 128 * 128;                  // 128 rows x 64 bf16
constexpr int MAX_K = 512;                             // im2col width limit (lookup table size)
constexpr int NCH = 16;                                // channel tile (UMMA N of the forward / data-gradient products)
constexpr int MAX_STAGES = 8, MAX_SLOTS = 4;

// Warp roles of a CTA: warp 0 issues the MMAs, then EPI_WARPS epilogue warps (groups of four, one warp per TMEM
// lane quarter, the groups taking alternate tiles), one bulk-copy loader warp, PROD_WARPS producer warps
// (8 chunk lanes x HALVES row halves).  The regular geometry runs one CTA per SM and is the only one instantiated; SMALL
// halves everything so that two CTAs — two independent pipelines and MMA issuers — share an SM (measured in round 2:
// no net gain, profiles/r2_tc_conv_small.md; the template parameter is kept, the instantiations are gone).
template <bool SMALL>
struct Geo {
  static constexpr int PROD_WARPS = SMALL ? 8 : 16;
  static constexpr int EPI_WARPS = SMALL ? 4 : 8;
  static constexpr int NACC = SMALL ? 4 : 8;           // TMEM accumulator ring of the forward / data-gradient kernel (32 columns each)
  static constexpr int MAX_POS = SMALL ? 1280 : 2048;  // positions per group (position table size)
  static constexpr int THREADS = 32 * (2 + EPI_WARPS + PROD_WARPS);
  static constexpr int PROD_THREADS = 32 * PROD_WARPS;
  static constexpr int FIRST_PROD_WARP = 2 + EPI_WARPS;
  static constexpr int HALVES = PROD_WARPS / 8;
  static constexpr int CTAS_PER_SM = SMALL ? 2 : 1;
  static constexpr int SMEM_LIMIT = SMALL ? 112 * 1024 : 227 * 1024;
};

struct ConvArgs {
  const float* src;        // im2col source [n][Cs][Hs][Ws]
  const float* w;          // filters [Cout][Cin][kH][kW]
  const float* bias;       // forward only (nullable)
  const float* dy;         // weight gradient: [n][Cout][Hp][Wp]
  float* out;              // forward: y [n][Cout][Hp][Wp]; data gradient: dx [n][Cin][Hp][Wp]
  float* dw;               // weight gradient
  float* db;               // weight gradient (nullable)
  float* part;             // weight gradient: per-CTA partials [grid][Ktot + 1][Cout] (column Ktot = the ones column, i.e. db)
  long long nsamples;
  int G;                   // samples per group: one bulk copy of G whole samples feeds G*Hp*Wp positions
  int Cs, Hs, Ws;          // source channels / height / width
  int Hp, Wp;              // positions per sample
  int kH, kW, pad_h, pad_w;
  int Cin, Cout;
  int Nout;                // output channels of the forward / data-gradient product (Cout resp. Cin)
  int Ktot;                // Cs * kH * kW
  int nkb;                 // 64-wide k-blocks (forward / data gradient), 64-wide im2col boxes (weight gradient)
  int dgrad;               // 1: data gradient (zero padding, unflipped filters)
  int relu;
  int stages, slots;       // operand-tile ring / staged-sample ring depths
  int slot_bytes, dy_off;  // bytes per staged group (source, then dY at dy_off for the weight gradient)
  int padded_bytes;        // data gradient: zero-padded copy of one group (0 otherwise)
  int tmem_cols;
};

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
template <int N>
__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); }

// One 16-byte chunk (8 consecutive k) of one im2col row.  off2[e/2] packs the offsets of taps e, e+1 relative
// to the patch origin (16 bits each, thread-constant registers), `mask` bit e = tap exists, `ones` bit e =
// the constant-1 column.
__device__ __forceinline__ void im2col_chunk(uint32_t dst, const float* __restrict__ staged, int origin, const uint32_t (&off2)[4], uint32_t mask,
                                             uint32_t ones) {
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float x = ((ones >> e) & 1u) ? 1.f : 0.f;
    const int o = (e & 1) ? (int)(off2[e >> 1] >> 16) : (int)(off2[e >> 1] & 0xffffu);
    if ((mask >> e) & 1u) x = staged[origin + o];
    v[e] = x;
  }
  st_shared_v4(dst, pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

// MODE 0: forward / data gradient.  MODE 1: weight gradient.  NKB: 64-wide k-blocks (MODE 0) / im2col boxes (MODE 1).
// Work item = a group of G consecutive samples.  The loader thread bulk-copies the group's source
// activations (and dY for the weight gradient) into a shared-memory slot (cp.async.bulk + mbarrier, a few
// groups ahead); the 8 producer warps expand them into SWIZZLE_128B operand tiles: warp w owns the w-th
// 16-byte chunk of every 64-wide k-block (its 8 tap offsets per block live in registers), lanes are 32
// consecutive positions (conflict-free shared-memory gathers), the patch origin of a position comes from a
// table built once per kernel.  One thread issues the MMAs, the epilogue warps drain TMEM.  The data
// gradient gathers from a zero-padded copy of the group (built by the producers), so it needs no bounds tests.
template <int MODE, int NKB, bool SMALL>
__global__ void __launch_bounds__(Geo<SMALL>::THREADS, Geo<SMALL>::CTAS_PER_SM) tc_conv_kernel(ConvArgs a) {
  using G_ = Geo<SMALL>;
  constexpr int PROD_WARPS = G_::PROD_WARPS, EPI_WARPS = G_::EPI_WARPS, NACC = G_::NACC, MAX_POS = G_::MAX_POS;
  constexpr int THREADS = G_::THREADS, PROD_THREADS = G_::PROD_THREADS, FIRST_PROD_WARP = G_::FIRST_PROD_WARP;
  constexpr int HALVES = G_::HALVES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  // layout: operand stages | [MODE 0: NKB x 2 KB filter tiles] | staged-sample slots | [padded group] | tables | barriers
  const int stages = a.stages, slots = a.slots;
  constexpr int stage_bytes = MODE == 0 ? TILE_BYTES : TILE_BYTES + NKB * (BLOCK_K * 128);
  const uint32_t wm_off = stages * stage_bytes;                              // MODE 0 only
  const uint32_t slot_off = wm_off + (MODE == 0 ? NKB * 2048 : 0);
  const uint32_t pad_off = slot_off + slots * a.slot_bytes;
  const uint32_t tab_off = pad_off + a.padded_bytes;
  int2* tab = reinterpret_cast<int2*>(smem + tab_off);                       // [MAX_K] tap table
  int* origin_tab = reinterpret_cast<int*>(smem + tab_off + MAX_K * 8);      // [MAX_POS] patch origin of a position (source offset)
  int* dybase_tab = origin_tab + MAX_POS;                                    // [MAX_POS] MODE 1: offset of (position, co = 0) in the staged dY
  uint4* offp = reinterpret_cast<uint4*>(smem + tab_off + MAX_K * 8 + 2 * MAX_POS * 4);   // [MAX_K / 8] per 16-byte chunk: 8 tap offsets, 16 bits each
  uint32_t* maskp = reinterpret_cast<uint32_t*>(offp + MAX_K / 8);           // [MAX_K / 8] bits 0-7: tap exists, bits 8-15: constant-1 column
  const uint32_t bar_base = smem_base + tab_off + MAX_K * 8 + 2 * MAX_POS * 4 + (MAX_K / 8) * 20;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                                   // operand tile written
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };                   // operand tile consumed by the MMAs
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * MAX_STAGES + i); };               // accumulator complete
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * MAX_STAGES + NACC + i); };       // accumulator drained
  auto sfull_bar = [&](int i) { return bar_base + 8u * (2 * MAX_STAGES + 2 * NACC + i); };    // staged group landed
  auto sempty_bar = [&](int i) { return bar_base + 8u * (2 * MAX_STAGES + 2 * NACC + MAX_SLOTS + i); };   // staged group expanded
  constexpr int NUM_BARS = 2 * MAX_STAGES + 2 * NACC + 2 * MAX_SLOTS;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + tab_off + MAX_K * 8 + 2 * MAX_POS * 4 + (MAX_K / 8) * 20 + 8 * NUM_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HWp = a.Hp * a.Wp, HWs = a.Hs * a.Ws, kk = a.kH * a.kW;
  // geometry of the buffer the gathers read: the raw staged group, or its zero-padded copy (data gradient)
  const int gW = a.Ws + 2 * a.pad_w, gH = a.Hs + 2 * a.pad_h, gplane = gW * gH, gsample = a.Cs * gplane;
  const long long num_groups = (a.nsamples + a.G - 1) / a.G;
  const int src_per_sample = a.Cs * HWs;                                     // floats
  const int max_pos = a.G * HWp;

  // ---- one-time setup ----
  for (int k = threadIdx.x; k < NKB * 64; k += THREADS) {
    int2 t = make_int2(0, 0);
    if (k < a.Ktot) {
      const int c = k / kk, r = k - c * kk, i = r / a.kW, j = r - i * a.kW;
      t = make_int2(c * gplane + i * gW + j, (i << 16) | j);
    }
    tab[k] = t;
  }
  for (int c8 = threadIdx.x; c8 < NKB * 8; c8 += THREADS) {
    uint32_t w[4] = {0, 0, 0, 0}, m = 0;
    for (int e = 0; e < 8; ++e) {
      const int k = c8 * 8 + e;
      if (k < a.Ktot) {
        const int c = k / kk, r = k - c * kk, i = r / a.kW, j = r - i * a.kW;
        w[e >> 1] |= ((uint32_t)(c * gplane + i * gW + j) & 0xffffu) << (16 * (e & 1));
        m |= 1u << e;
      } else if (MODE == 1 && k == a.Ktot) {
        m |= 1u << (8 + e);
      }
    }
    offp[c8] = make_uint4(w[0], w[1], w[2], w[3]);
    maskp[c8] = m;
  }
  for (int pl = threadIdx.x; pl < max_pos; pl += THREADS) {
    const int gl = pl / HWp, rem = pl - gl * HWp, py = rem / a.Wp, px = rem - py * a.Wp;
    origin_tab[pl] = gl * gsample + py * gW + px;
    dybase_tab[pl] = gl * a.Cout * HWp + rem;
  }
  if (MODE == 0) {
    // filter tiles: [NKB][16 rows][64 k] bf16, K-major SWIZZLE_128B (row r at r*128, 16-byte chunk c at (c ^ (r & 7)) * 16)
    for (int idx = threadIdx.x; idx < NKB * NCH * BLOCK_K; idx += THREADS) {
      const int kb = idx / (NCH * BLOCK_K), r = (idx / BLOCK_K) % NCH, kc = idx % BLOCK_K, k = kb * BLOCK_K + kc;
      float v = 0.f;
      if (r < a.Nout && k < a.Ktot) {
        const int c = k / kk, rr = k - c * kk, i = rr / a.kW, j = rr - i * a.kW;
        v = a.dgrad ? a.w[((c * a.Cin + r) * a.kH + i) * a.kW + j]
                    : a.w[((r * a.Cin + c) * a.kH + (a.kH - 1 - i)) * a.kW + (a.kW - 1 - j)];
      }
      const uint32_t off = kb * 2048 + r * 128 + (((kc >> 3) ^ (r & 7)) << 4) + (kc & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(smem + wm_off + off) = __float2bfloat16_rn(v);
    }
    // zero borders of the padded copy (interiors are rewritten for every group)
    for (int i = threadIdx.x; i < a.padded_bytes / 16; i += THREADS)
      *reinterpret_cast<uint4*>(smem + pad_off + i * 16) = make_uint4(0, 0, 0, 0);
  } else {
    // dY^T operand tiles: rows >= Cout stay zero for the whole kernel
    for (int s = 0; s < stages; ++s)
      for (int i = threadIdx.x; i < TILE_BYTES / 16; i += THREADS)
        *reinterpret_cast<uint4*>(smem + s * stage_bytes + i * 16) = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full_bar(s), PROD_WARPS); mbar_init(empty_bar(s), 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), 4); }
    for (int i = 0; i < slots; ++i) { mbar_init(sfull_bar(i), 1); mbar_init(sempty_bar(i), PROD_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(a.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();                 // generic-proxy writes above (filter tiles, zeroed rows) -> visible to the tensor core
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto group_samples = [&](long long grp) { return (int)min((long long)a.G, a.nsamples - grp * a.G); };
  // 16-wide k slices of k-block kb that contain taps (MODE 0): the MMAs of the remaining slices are skipped
  auto kslices = [&](int kb) { return min(BLOCK_K / UMMA_K, (a.Ktot - kb * BLOCK_K + UMMA_K - 1) / UMMA_K); };

  if (warp == 1 + EPI_WARPS) {
    // ===== loader: one thread bulk-copies whole groups into the slot ring =====
    if (lane == 0) {
      int slot = 0; uint32_t sphase = 0;
      for (long long grp = blockIdx.x; grp < num_groups; grp += gridDim.x) {
        const int gs = group_samples(grp);
        mbar_wait(sempty_bar(slot), sphase ^ 1);
        const uint32_t dst = smem_base + slot_off + slot * a.slot_bytes;
        const uint32_t sbytes = ((uint32_t)(gs * src_per_sample) * 4u + 15u) & ~15u;
        const uint32_t dbytes = MODE == 1 ? (((uint32_t)(gs * a.Cout * HWp) * 4u + 15u) & ~15u) : 0u;
        mbar_expect_tx(sfull_bar(slot), sbytes + dbytes);
        bulk_copy_g2s(dst, a.src + grp * a.G * (long long)src_per_sample, sbytes, sfull_bar(slot));
        if (MODE == 1) bulk_copy_g2s(dst + a.dy_off, a.dy + grp * a.G * (long long)a.Cout * HWp, dbytes, sfull_bar(slot));
        if (++slot == slots) { slot = 0; sphase ^= 1; }
      }
    }
  } else if (warp >= FIRST_PROD_WARP) {
    // ===== producers: warp = (chunk lane cu, row half) =====
    const int pw = warp - FIRST_PROD_WARP;
    const int cu = pw & 7;                                // this warp's 16-byte chunk of every 64-wide k-block
    const int half = pw >> 3;                             // which half of a tile's rows
    const int t = threadIdx.x - 32 * FIRST_PROD_WARP;
    int stage = 0; uint32_t phase = 0;
    int slot = 0; uint32_t sphase = 0;
    for (long long grp = blockIdx.x; grp < num_groups; grp += gridDim.x) {
      const int gs = group_samples(grp);
      const int npos = gs * HWp;
      mbar_wait(sfull_bar(slot), sphase);
      const float* raw = reinterpret_cast<const float*>(smem + slot_off + slot * a.slot_bytes);
      const float* staged = raw;
      if (MODE == 0 && a.padded_bytes != 0) {
        // data gradient: copy the group's rows into the interior of the zero-padded buffer
        float* padded = reinterpret_cast<float*>(smem + pad_off);
        producers_sync<PROD_THREADS>();                                  // everyone has finished gathering the previous group
        const int rows = gs * a.Cs * a.Hs;
        for (int r = t; r < rows; r += PROD_THREADS) {
          const int pc = r / a.Hs, y = r - pc * a.Hs;      // pc = (sample, channel)
          const float* srow = raw + r * a.Ws;
          float* drow = padded + pc * gplane + (y + a.pad_h) * gW + a.pad_w;
          for (int x = 0; x < a.Ws; ++x) drow[x] = srow[x];
        }
        producers_sync<PROD_THREADS>();
        __syncwarp();
        if (lane == 0) mbar_arrive(sempty_bar(slot));      // the raw slot can be refilled already
        staged = padded;
      }
      if (MODE == 0) {
        const int tiles = (npos + 127) / 128;
        constexpr int RQ = 4 / HALVES;                    // 32-row quarters of a tile per producer warp
        for (int tile = 0; tile < tiles; ++tile) {
          int origin[RQ];
#pragma unroll
          for (int ps = 0; ps < RQ; ++ps) {
            const int pl = tile * 128 + (half * RQ + ps) * 32 + lane;
            origin[ps] = pl < npos ? origin_tab[pl] : -1;
          }
#pragma unroll
          for (int kb = 0; kb < NKB; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            // 16-wide k slices beyond the last tap are never read by the MMAs (see the issuer): no stores needed
            if (cu * 8 < kslices(kb) * UMMA_K) {
              const uint32_t sbase = smem_base + stage * stage_bytes;
              const uint4 o4 = offp[kb * 8 + cu];
              const uint32_t off2[4] = {o4.x, o4.y, o4.z, o4.w};
              const uint32_t mk = maskp[kb * 8 + cu];
#pragma unroll
              for (int ps = 0; ps < RQ; ++ps) {
                const int row = (half * RQ + ps) * 32 + lane;
                const uint32_t dst = sbase + row * 128 + ((cu ^ (row & 7)) << 4);
                if (origin[ps] >= 0 && mk != 0) im2col_chunk(dst, staged, origin[ps], off2, mk, 0u);
                else st_shared_v4(dst, 0u, 0u, 0u, 0u);
              }
              fence_async_smem();                          // own writes -> async proxy, then one arrive per warp
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(stage));
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      } else {
        const float* dys = reinterpret_cast<const float*>(smem + slot_off + slot * a.slot_bytes + a.dy_off);
        const int blocks = (npos + 63) / 64;
        for (int pb = 0; pb < blocks; ++pb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + TILE_BYTES;
          // im2col rows (B operand: MN-major boxes of 64 k, rows = positions)
          constexpr int RH = 2 / HALVES;                  // 32-row halves of a block per producer warp
#pragma unroll
          for (int ps = 0; ps < RH; ++ps) {
            const int row = (half * RH + ps) * 32 + lane, pl = pb * 64 + row;
            const int origin = pl < npos ? origin_tab[pl] : -1;
#pragma unroll
            for (int box = 0; box < NKB; ++box) {
              const uint32_t dst = sb + box * (BLOCK_K * 128) + row * 128 + ((cu ^ (row & 7)) << 4);
              const uint4 o4 = offp[box * 8 + cu];
              const uint32_t off2[4] = {o4.x, o4.y, o4.z, o4.w};
              const uint32_t mk = maskp[box * 8 + cu];
              if (origin >= 0 && mk != 0) im2col_chunk(dst, staged, origin, off2, mk & 0xffu, mk >> 8);
              else st_shared_v4(dst, 0u, 0u, 0u, 0u);
            }
          }
          // dY^T rows (A operand): row co, 64 positions of this block, 8 per unit
          for (int u = t; u < a.Cout * 8; u += PROD_THREADS) {
            const int co = u >> 3, chunk = u & 7;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int pe = pb * 64 + chunk * 8 + e;
              v[e] = pe < npos ? dys[dybase_tab[pe] + co * HWp] : 0.f;
            }
            st_shared_v4(sa + co * 128 + ((chunk ^ (co & 7)) << 4), pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(full_bar(stage));
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
      if (!(MODE == 0 && a.padded_bytes != 0)) {
        __syncwarp();
        if (lane == 0) mbar_arrive(sempty_bar(slot));      // this warp no longer reads the staged group
      }
      if (++slot == slots) { slot = 0; sphase ^= 1; }
    }
  } else if (MODE == 0) {
    if (warp == 0) {
      // ===== MMA issuer =====
      if (lane == 0) {
        const uint32_t idesc = make_idesc_mn(false, false, 128, NCH);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (long long grp = blockIdx.x; grp < num_groups; grp += gridDim.x) {
          const int tiles = (group_samples(grp) * HWp + 127) / 128;
          for (int tile = 0; tile < tiles; ++tile) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1);
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 32;
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
              mbar_wait(full_bar(stage), phase);
              tcgen05_fence_after();
              const uint32_t sa = smem_base + stage * stage_bytes, sb = smem_base + wm_off + kb * 2048;
              const int ks = kslices(kb);
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                if (k < ks) umma_bf16(d_tmem, make_desc(sa + k * (UMMA_K * 2), 16, 1024), make_desc(sb + k * (UMMA_K * 2), 16, 1024), idesc, (kb | k) != 0);
              tcgen05_commit(empty_bar(stage));
              if (kb == NKB - 1) tcgen05_commit(tfull_bar(acc));
              if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
          }
        }
      }
    } else {
      // ===== epilogue: one accumulator row (= one position) per thread, 16 channels; the two groups of four
      //       warps take alternate tiles =====
      const int q = warp & 3, parity = (warp - 1) >> 2;
      float bias_r[NCH];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) bias_r[ch] = (a.bias != nullptr && ch < a.Nout) ? __ldg(a.bias + ch) : 0.f;
      int acc = 0; uint32_t acc_phase = 0;
      int tcount = 0;
      for (long long grp = blockIdx.x; grp < num_groups; grp += gridDim.x) {
        const int npos = group_samples(grp) * HWp;
        const int tiles = (npos + 127) / 128;
        int pl = q * 32 + lane;                            // position inside the group
        int gl = pl / HWp, rem = pl - gl * HWp;
        for (int tile = 0; tile < tiles; ++tile, ++tcount) {
          if ((tcount & (EPI_WARPS / 4 - 1)) == parity) {
            mbar_wait(tfull_bar(acc), acc_phase);
            tcgen05_fence_after();
            uint32_t r[NCH];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 32, r);
            tmem_ld_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));   // values are in registers: the accumulator is free again
            if (pl < npos) {
              float* o = a.out + (grp * a.G + gl) * a.Nout * (long long)HWp + rem;
#pragma unroll
              for (int ch = 0; ch < NCH; ++ch) {
                if (ch < a.Nout) {
                  float v = __uint_as_float(r[ch]) + bias_r[ch];
                  if (a.relu) v = fmaxf(v, 0.f);
                  o[ch * HWp] = v;                         // lanes = consecutive positions: coalesced per channel
                }
              }
            }
          }
          if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
          pl += 128; rem += 128;
          while (rem >= HWp) { rem -= HWp; ++gl; }
        }
      }
    }
  } else {
    // ================= weight gradient: MMA issuer / epilogue =================
    constexpr int N = NKB * 64;
    if (warp == 0) {
      if (lane == 0) {
        const uint32_t idesc = make_idesc_mn(false, true, 128, N);
        int stage = 0; uint32_t phase = 0;
        bool first = true;
        for (long long grp = blockIdx.x; grp < num_groups; grp += gridDim.x) {
          const int blocks = (group_samples(grp) * HWp + 63) / 64;
          for (int pb = 0; pb < blocks; ++pb) {
            mbar_wait(full_bar(stage), phase);
            tcgen05_fence_after();
            const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + TILE_BYTES;
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_bf16(tmem_base, make_desc(sa + k * (UMMA_K * 2), 16, 1024), make_desc(sb + k * (UMMA_K * 128), BLOCK_K * 128, 1024), idesc,
                        (first && k == 0) ? 0u : 1u);
            first = false;
            tcgen05_commit(empty_bar(stage));
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
        if (!first) tcgen05_commit(tfull_bar(0));
      }
    } else {
      // ===== epilogue (once): rows 0..Cout-1 of the accumulator live in TMEM lane quarter 0 =====
      if (warp == 4 && (long long)blockIdx.x < num_groups) {
        mbar_wait(tfull_bar(0), 0);
        tcgen05_fence_after();
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + c0, r);
          tmem_ld_wait();
          if (lane < a.Cout) {
            float* part = a.part + (size_t)blockIdx.x * (a.Ktot + 1) * a.Cout + lane;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int k = c0 + j;
              if (k <= a.Ktot) part[(size_t)k * a.Cout] = __uint_as_float(r[j]);
            }
          }
        }
        tcgen05_fence_before();
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

// Sums the weight-gradient partials of all CTAs in CTA order (deterministic) and scatters them into dw (flipped taps: the
// product's column k = (ci, i, j) pairs dY with x[.., oy + i, ox + j], which is the gradient of w[co][ci][kH-1-i][kW-1-j]) and db.
__global__ void tc_conv_wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, float* __restrict__ db, int nparts,
                                            int Ktot, int Cin, int Cout, int kH, int kW) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;       // e = k * Cout + co
  if (e >= (Ktot + 1) * Cout) return;
  const int k = e / Cout, co = e - k * Cout;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(size_t)p * (Ktot + 1) * Cout + e];
  if (k == Ktot) {
    if (db != nullptr) db[co] = s;
    return;
  }
  const int kk = kH * kW, ci = k / kk, t = k - ci * kk, i = t / kW, j = t - i * kW;
  dw[((co * Cin + ci) * kH + (kH - 1 - i)) * kW + (kW - 1 - j)] = s;
}

// samples per group: whole samples, 16-byte aligned group starts, bounded bytes per slot, <= max_pos
// positions, and as few padded rows as possible in the last 128-position tile (64-position block)
int choose_group(int src_floats, int dy_floats, int padded_floats, int HWp, int unit, int max_bytes, int max_pos = Geo<false>::MAX_POS) {
  int best = 0;
  double best_eff = -1.0;
  for (int g = 1; g <= 32; ++g) {
    if ((g * src_floats) % 4 != 0 || (g * dy_floats) % 4 != 0) continue;
    if ((long long)g * (src_floats + dy_floats + padded_floats) * 4 > max_bytes || g * HWp > max_pos) break;
    const int pos = g * HWp;
    const double eff = (double)pos / (double)(((pos + unit - 1) / unit) * unit);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = g; }
  }
  return best;
}

// Sizes the group, the rings and the shared-memory footprint for one geometry; false = does not fit.
template <int MODE, int NKB, bool SMALL>
bool plan_launch(ConvArgs& a, int dy_floats_per_sample, int* bytes_out) {
  using G_ = Geo<SMALL>;
  const int HWp = a.Hp * a.Wp, src_floats = a.Cs * a.Hs * a.Ws;
  const int padded_floats = a.dgrad ? a.Cs * (a.Hs + 2 * a.pad_h) * (a.Ws + 2 * a.pad_w) : 0;
  const int slot_budget = SMALL ? (a.dgrad ? 36 * 1024 : 20 * 1024) : (a.dgrad ? 80 * 1024 : 40 * 1024);
  a.G = choose_group(src_floats, dy_floats_per_sample, padded_floats, HWp, MODE == 0 ? 128 : 64, slot_budget, G_::MAX_POS);
  if (a.G <= 0) return false;
  const int src_bytes = (a.G * src_floats * 4 + 127) & ~127;
  const int dy_bytes = (a.G * dy_floats_per_sample * 4 + 127) & ~127;
  a.dy_off = src_bytes;
  a.slot_bytes = src_bytes + dy_bytes;
  a.padded_bytes = (a.G * padded_floats * 4 + 127) & ~127;
  constexpr int stage_bytes = MODE == 0 ? TILE_BYTES : TILE_BYTES + NKB * (BLOCK_K * 128);
  const int fixed = (MODE == 0 ? NKB * 2048 : 0) + a.padded_bytes + MAX_K * 8 + 2 * G_::MAX_POS * 4 + (MAX_K / 8) * 20 +
                    8 * (2 * MAX_STAGES + 2 * G_::NACC + 2 * MAX_SLOTS) + 16 + 1024;
  a.slots = 2;
  a.stages = MODE == 0 ? 6 : 3;
  while (a.stages > 2 && fixed + a.stages * stage_bytes + a.slots * a.slot_bytes > G_::SMEM_LIMIT) --a.stages;
  while (a.slots < 3 && fixed + a.stages * stage_bytes + (a.slots + 1) * a.slot_bytes <= G_::SMEM_LIMIT) ++a.slots;
  while (a.stages < (MODE == 0 ? MAX_STAGES : 4) && fixed + (a.stages + 1) * stage_bytes + a.slots * a.slot_bytes <= G_::SMEM_LIMIT) ++a.stages;
  *bytes_out = fixed + a.stages * stage_bytes + a.slots * a.slot_bytes;
  a.tmem_cols = G_::NACC * 32;
  if (MODE == 1) { a.tmem_cols = 32; while (a.tmem_cols < NKB * 64) a.tmem_cols *= 2; }
  return *bytes_out <= G_::SMEM_LIMIT;
}

template <int MODE, int NKB, bool SMALL>
int launch_geo(pg_ctx* ctx, const ConvArgs& a, int bytes, int* grid_out) {
  using G_ = Geo<SMALL>;
  static int attr_bytes_dev[32] = {};
  int& attr_bytes = attr_bytes_dev[ctx->device & 31];
  if (bytes > attr_bytes) {
    PG_CHECK_CUDA(cudaFuncSetAttribute(tc_conv_kernel<MODE, NKB, SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    attr_bytes = bytes;
  }
  const long long groups = (a.nsamples + a.G - 1) / a.G;
  const long long ctas = (long long)ctx->num_sms * G_::CTAS_PER_SM;
  const int grid = (int)(groups < ctas ? groups : ctas);
  tc_conv_kernel<MODE, NKB, SMALL><<<grid, G_::THREADS, bytes, ctx->stream>>>(a);
  PG_LAUNCHED(ctx);
  if (grid_out != nullptr) *grid_out = grid;
  return PG_OK;
}

template <int MODE, int NKB>
int launch_nkb(pg_ctx* ctx, ConvArgs& a, int dy_floats_per_sample, int* grid_out) {
  // (the half-size two-CTAs-per-SM geometry of round 1 was measured in round 2 — conv1 forward 138 -> 113 us, conv2
  //  backward 346 -> 435 us at C3 B = 8192, profiles/r2_tc_conv_small.md — and is no longer instantiated)
  int bytes = 0;
  const bool ok = plan_launch<MODE, NKB, false>(a, dy_floats_per_sample, &bytes);
  PG_REQUIRE(a.G > 0, "tensor-core conv: a sample (with its output gradient) does not fit a shared-memory slot");
  PG_REQUIRE(ok, "tensor-core conv: shared-memory budget exceeded");
  return launch_geo<MODE, NKB, false>(ctx, a, bytes, grid_out);
}

template <int MODE>
int launch(pg_ctx* ctx, ConvArgs& a, int dy_floats_per_sample, int* grid_out = nullptr) {
  switch (a.nkb) {
    case 1: return launch_nkb<MODE, 1>(ctx, a, dy_floats_per_sample, grid_out);
    case 2: return launch_nkb<MODE, 2>(ctx, a, dy_floats_per_sample, grid_out);
    case 3: return launch_nkb<MODE, 3>(ctx, a, dy_floats_per_sample, grid_out);
    case 4: return launch_nkb<MODE, 4>(ctx, a, dy_floats_per_sample, grid_out);
  }
  if (MODE == 0) {
    switch (a.nkb) {
      case 5: return launch_nkb<0, 5>(ctx, a, dy_floats_per_sample, grid_out);
      case 6: return launch_nkb<0, 6>(ctx, a, dy_floats_per_sample, grid_out);
      case 7: return launch_nkb<0, 7>(ctx, a, dy_floats_per_sample, grid_out);
    }
  }
  pg_set_error("tensor-core conv: %d k-blocks not instantiated", a.nkb);
  return PG_ERR_ARG;
}

bool supported(int Cin, int H, int W, int Cout, int kH, int kW) {
  if (!(Cin >= 1 && Cout >= 1 && Cin <= NCH && Cout <= NCH && kH >= 1 && kW >= 1 && kH < 256 && kW < 256 && H >= kH && W >= kW &&
        Cin * kH * kW + 1 <= 256 &&         // weight gradient: im2col width (+ ones column) = UMMA N <= 256
        Cout * kH * kW <= 448))             // data gradient: 7 k-blocks of filters in shared memory
    return false;
  const int Ho = H - kH + 1, Wo = W - kW + 1;
  const int padded = Cout * (Ho + 2 * (kH - 1)) * (Wo + 2 * (kW - 1));
  if (Cin * H * W >= 65536 || padded >= 65536) return false;                       // 16-bit tap offsets
  return choose_group(Cin * H * W, 0, 0, Ho * Wo, 128, 40 * 1024) > 0 &&            // forward
         choose_group(Cin * H * W, Cout * Ho * Wo, 0, Ho * Wo, 64, 40 * 1024) > 0 && // weight gradient
         choose_group(Cout * Ho * Wo, 0, padded, H * W, 128, 80 * 1024) > 0;         // data gradient
}

}  // namespace

extern "C" {

int pg_tc_conv2d_supported(int Cin, int H, int W, int Cout, int kH, int kW) { return supported(Cin, H, W, Cout, kH, kW) ? 1 : 0; }

int pg_tc_conv2d_fwd(pg_ctx* ctx, const void* x, const void* w, const void* b, void* y, int64_t N, int Cin, int H, int W, int Cout,
                     int kH, int kW, int act) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(N > 0 && H >= kH && W >= kW, "pg_tc_conv2d_fwd: empty output");
  PG_REQUIRE(supported(Cin, H, W, Cout, kH, kW), "pg_tc_conv2d_fwd: shape outside the tensor-core envelope (pg_tc_conv2d_supported)");
  PG_REQUIRE(act == PG_ACT_NONE || act == PG_ACT_RELU, "pg_tc_conv2d_fwd: unknown activation");
  ConvArgs a{};
  a.src = (const float*)x; a.w = (const float*)w; a.bias = (const float*)b; a.out = (float*)y;
  a.Cs = Cin; a.Hs = H; a.Ws = W; a.Hp = H - kH + 1; a.Wp = W - kW + 1;
  a.kH = kH; a.kW = kW; a.Cin = Cin; a.Cout = Cout; a.Nout = Cout;
  a.Ktot = Cin * kH * kW; a.nkb = (int)pg_cdiv(a.Ktot, BLOCK_K);
  a.nsamples = N; a.relu = act == PG_ACT_RELU;
  return launch<0>(ctx, a, 0);
}

int pg_tc_conv2d_bwd(pg_ctx* ctx, const void* x, const void* w, const void* dy, void* dw, void* db, void* dx, int64_t N, int Cin, int H,
                     int W, int Cout, int kH, int kW) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(N > 0 && H >= kH && W >= kW, "pg_tc_conv2d_bwd: empty output");
  PG_REQUIRE(supported(Cin, H, W, Cout, kH, kW), "pg_tc_conv2d_bwd: shape outside the tensor-core envelope (pg_tc_conv2d_supported)");
  const int Ho = H - kH + 1, Wo = W - kW + 1;
  if (dw != nullptr || db != nullptr) {
    PG_REQUIRE(dw != nullptr, "pg_tc_conv2d_bwd: db without dw is not supported on the tensor-core path");
    ConvArgs a{};
    a.src = (const float*)x; a.dy = (const float*)dy; a.dw = (float*)dw; a.db = (float*)db;
    a.Cs = Cin; a.Hs = H; a.Ws = W; a.Hp = Ho; a.Wp = Wo;
    a.kH = kH; a.kW = kW; a.Cin = Cin; a.Cout = Cout;
    a.Ktot = Cin * kH * kW; a.nkb = (int)pg_cdiv(a.Ktot + 1, 64);
    a.nsamples = N;
    // one partial [Ktot + 1][Cout] per CTA (one CTA per SM at most), summed in CTA order afterwards
    const size_t part_floats = (size_t)(a.Ktot + 1) * Cout;
    int rc = pg_ws_ensure(ctx, (size_t)ctx->real_sms * part_floats * sizeof(float));
    if (rc != PG_OK) return rc;
    a.part = (float*)ctx->ws;
    int grid = 0;
    rc = launch<1>(ctx, a, Cout * Ho * Wo, &grid);
    if (rc != PG_OK) return rc;
    tc_conv_wgrad_reduce_kernel<<<(unsigned)pg_cdiv((int64_t)part_floats, 128), 128, 0, ctx->stream>>>(a.part, a.dw, a.db, grid, a.Ktot, Cin, Cout, kH, kW);
    PG_LAUNCHED(ctx);
  }
  if (dx != nullptr) {
    ConvArgs a{};
    a.src = (const float*)dy; a.w = (const float*)w; a.out = (float*)dx;
    a.Cs = Cout; a.Hs = Ho; a.Ws = Wo; a.Hp = H; a.Wp = W;
    a.kH = kH; a.kW = kW; a.pad_h = kH - 1; a.pad_w = kW - 1; a.Cin = Cin; a.Cout = Cout; a.Nout = Cin;
    a.Ktot = Cout * kH * kW; a.nkb = (int)pg_cdiv(a.Ktot, BLOCK_K); a.dgrad = 1;
    a.nsamples = N;
    int rc = launch<0>(ctx, a, 0);
    if (rc != PG_OK) return rc;
  }
  return PG_OK;
}

}  // extern "C"
