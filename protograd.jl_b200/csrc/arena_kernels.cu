// arena_kernels.cu — Model vector-space algebra, dot and fused optimizer updates over one flat
// parameter arena.  Replaces src/model.jl:103-202 (per-leaf Julia broadcasts / BLAS dot) and
// src/optimizers/*.jl (1–5 broadcast passes per step) with single 128-bit-vectorised passes.
//
// Numerics: Julia's broadcast evaluates each element in the promoted type of the operands and
// rounds once on store.  Kernels therefore take the scalar's Julia type as a
// flag and do the per-element math in double when it was a Float64, using the non-contracting
// __d*_rn / __f*_rn intrinsics so no FMA fusion changes the bits.
#include "common.cuh"
#include "optim.cuh"

namespace {

template <typename T> struct VecOf;
template <> struct VecOf<float> { using type = float4; static constexpr int N = 4; };
template <> struct VecOf<double> { using type = double2; static constexpr int N = 2; };

template <int N> struct Ptrs { const void* p[N > 0 ? N : 1]; };
template <int N> struct MPtrs { void* p[N > 0 ? N : 1]; };

// Generic elementwise kernel: NIN input arrays, NOUT output arrays (null outputs skipped),
// optional bf16 shadow of output 0.  f(in[], out[]) is applied per element.
template <typename T, int NIN, int NOUT, typename F>
__global__ void __launch_bounds__(256) ew_kernel(F f, Ptrs<NIN> in, MPtrs<NOUT> out, __nv_bfloat16* shadow, int64_t n) {
  using V = typename VecOf<T>::type;
  constexpr int VN = VecOf<T>::N;
  const int64_t nvec = n / VN;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    T a[NIN > 0 ? NIN : 1][VN];
#pragma unroll
    for (int k = 0; k < NIN; ++k) {
      V v = reinterpret_cast<const V*>(in.p[k])[i];
      const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
      for (int j = 0; j < VN; ++j) a[k][j] = e[j];
    }
    T r[NOUT][VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) {
      T ein[NIN > 0 ? NIN : 1], eout[NOUT];
#pragma unroll
      for (int k = 0; k < NIN; ++k) ein[k] = a[k][j];
      f(ein, eout);
#pragma unroll
      for (int k = 0; k < NOUT; ++k) r[k][j] = eout[k];
    }
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
      if (out.p[k] != nullptr) reinterpret_cast<V*>(out.p[k])[i] = *reinterpret_cast<V*>(r[k]);
    }
    if (shadow != nullptr) {
      if constexpr (VN == 4) {  // one packed 8-byte store per thread
        __nv_bfloat162 pk[2];
        pk[0] = __floats2bfloat162_rn((float)r[0][0], (float)r[0][1]);
        pk[1] = __floats2bfloat162_rn((float)r[0][2], (float)r[0][3]);
        reinterpret_cast<uint2*>(shadow)[i] = *reinterpret_cast<uint2*>(pk);
      } else {
#pragma unroll
        for (int j = 0; j < VN; ++j) shadow[i * VN + j] = __float2bfloat16_rn((float)r[0][j]);
      }
    }
  }
  // scalar tail
  for (int64_t i = nvec * VN + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    T ein[NIN > 0 ? NIN : 1], eout[NOUT];
#pragma unroll
    for (int k = 0; k < NIN; ++k) ein[k] = reinterpret_cast<const T*>(in.p[k])[i];
    f(ein, eout);
#pragma unroll
    for (int k = 0; k < NOUT; ++k)
      if (out.p[k] != nullptr) reinterpret_cast<T*>(out.p[k])[i] = eout[k];
    if (shadow != nullptr) shadow[i] = __float2bfloat16_rn((float)eout[0]);
  }
}

template <typename T, int NIN, int NOUT, typename F>
int ew_launch(pg_ctx* ctx, F f, Ptrs<NIN> in, MPtrs<NOUT> out, void* shadow, int64_t n) {
  if (n <= 0) return PG_OK;
  for (int k = 0; k < NIN; ++k) PG_REQUIRE(((uintptr_t)in.p[k] & 15) == 0 && in.p[k], "input pointer null or not 16-byte aligned");
  for (int k = 0; k < NOUT; ++k) PG_REQUIRE(((uintptr_t)out.p[k] & 15) == 0, "output pointer not 16-byte aligned");
  constexpr int VN = VecOf<T>::N;
  int64_t work = pg_cdiv(n, VN);
  int64_t blocks = pg_cdiv(work, 256);
  int64_t cap = (int64_t)ctx->num_sms * 8;  // 8 resident CTAs of 256 threads per SM, grid-stride beyond
  if (blocks > cap) blocks = cap;
  ew_kernel<T, NIN, NOUT, F><<<(unsigned)blocks, 256, 0, ctx->stream>>>(f, in, out, (__nv_bfloat16*)shadow, n);
  PG_LAUNCHED(ctx);
  return PG_OK;
}

// ---- functors ------------------------------------------------------------------------------
template <typename T, int OP> struct BinF {
  __device__ __forceinline__ void operator()(const T* in, T* out) const {
    out[0] = OP == PG_ADD ? Ar<T>::add(in[0], in[1]) : Ar<T>::sub(in[0], in[1]);
  }
};
template <typename T, typename P, int OP> struct ScalarF {
  P c;
  __device__ __forceinline__ void operator()(const T* in, T* out) const {
    P x = (P)in[0], r;
    if (OP == PG_X_PLUS_C) r = Ar<P>::add(x, c);
    else if (OP == PG_X_MINUS_C) r = Ar<P>::sub(x, c);
    else if (OP == PG_C_MINUS_X) r = Ar<P>::sub(c, x);
    else if (OP == PG_X_TIMES_C) r = Ar<P>::mul(x, c);
    else if (OP == PG_X_OVER_C) r = Ar<P>::div(x, c);
    else r = Ar<P>::div(c, x);
    out[0] = (T)r;
  }
};
template <typename T, typename P, int SIGN> struct AxpyF {  // out = x ± s*g
  P s;
  __device__ __forceinline__ void operator()(const T* in, T* out) const {
    P t = Ar<P>::mul(s, (P)in[1]);
    out[0] = (T)(SIGN > 0 ? Ar<P>::add((P)in[0], t) : Ar<P>::sub((P)in[0], t));
  }
};
template <typename T> struct FillF {
  T c;
  __device__ __forceinline__ void operator()(const T*, T* out) const { out[0] = c; }
};
// heavy ball: in = {p, d, g}; out = {p, d}
template <typename T, typename PS, typename PM> struct HeavyBallF {
  PS s; PM mu;
  __device__ __forceinline__ void operator()(const T* in, T* out) const {
    PM t1 = Ar<PM>::mul(mu, (PM)in[1]);
    PS t2 = Ar<PS>::mul(s, (PS)in[2]);
    using PP = decltype(t1 + t2);
    T d = (T)Ar<PP>::sub((PP)t1, (PP)t2);
    out[1] = d;
    out[0] = Ar<T>::add(in[0], d);
  }
};
// nesterov: in = {p, g}; out = {p, prev}
template <typename T, typename P> struct NesterovF {
  P s; double mu;
  const double* dev;   // non-null: momentum comes from device scalars (CUDA-graph replay), dev[4]
  __device__ __forceinline__ void operator()(const T* in, T* out) const {
    const double mu = dev ? dev[4] : this->mu;
    T prev = in[0];
    T p1 = (T)Ar<P>::sub((P)in[0], Ar<P>::mul(s, (P)in[1]));
    T diff = Ar<T>::sub(p1, prev);
    out[0] = (T)__dadd_rn((double)p1, __dmul_rn(mu, (double)diff));
    out[1] = prev;
  }
};
// ---- dot -----------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) dot_kernel(const T* __restrict__ x, const T* __restrict__ y, int64_t n,
                                                  double* partials, unsigned int* counter, double* result) {
  using V = typename VecOf<T>::type;
  constexpr int VN = VecOf<T>::N;
  __shared__ double sm[32];
  __shared__ bool is_last;
  double acc = 0.0;
  const int64_t nvec = n / VN;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    V a = reinterpret_cast<const V*>(x)[i], b = reinterpret_cast<const V*>(y)[i];
    const T* ea = reinterpret_cast<const T*>(&a);
    const T* eb = reinterpret_cast<const T*>(&b);
#pragma unroll
    for (int j = 0; j < VN; ++j) acc += (double)ea[j] * (double)eb[j];
  }
  for (int64_t i = nvec * VN + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += (double)x[i] * (double)y[i];
  double bs = block_sum(acc, sm);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = bs;
    __threadfence();
    unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double a2 = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) a2 += partials[i];
    double tot = block_sum(a2, sm);
    if (threadIdx.x == 0) { *result = tot; *counter = 0u; }
  }
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t n) {
  const int64_t nvec = n / 8;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 a = reinterpret_cast<const float4*>(x)[2 * i], b = reinterpret_cast<const float4*>(x)[2 * i + 1];
    __nv_bfloat162 r[4];
    r[0] = __floats2bfloat162_rn(a.x, a.y); r[1] = __floats2bfloat162_rn(a.z, a.w);
    r[2] = __floats2bfloat162_rn(b.x, b.y); r[3] = __floats2bfloat162_rn(b.z, b.w);
    reinterpret_cast<uint4*>(out)[i] = *reinterpret_cast<uint4*>(r);
  }
  for (int64_t i = nvec * 8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = __float2bfloat16_rn(x[i]);
}

// per-step scalars kept on the device so a captured CUDA graph can be replayed unchanged:
// s[0] = Adam `it` (starts at 1), s[1] = 1 - beta1^it, s[2] = 1 - beta2^it  (adam.jl:36-40)
// s[3] = Nesterov t (starts at 1.0), s[4] = momentum (nesterov_method.jl:3-6)
__global__ void opt_tick_kernel(double* s, int kind, double b1, double b2) {
  if (kind == 0) {
    const double it = s[0];
    s[1] = 1.0 - pow(b1, it);
    s[2] = 1.0 - pow(b2, it);
    s[0] = it + 1.0;
  } else {
    const double t = s[3];
    const double t_next = (1.0 + sqrt(1.0 + 4.0 * t * t)) / 2.0;
    s[4] = (t - 1.0) / t_next;
    s[3] = t_next;
  }
}

template <typename T, int OP>
int scalar_dispatch(pg_ctx* ctx, void* out, const void* x, double c, int flags, int64_t n) {
  Ptrs<1> in{{x}}; MPtrs<1> o{{out}};
  if (sizeof(T) == 8 || (flags & PG_SCALAR_F64)) return ew_launch<T, 1, 1>(ctx, ScalarF<T, double, OP>{c}, in, o, nullptr, n);
  return ew_launch<T, 1, 1>(ctx, ScalarF<T, float, OP>{(float)c}, in, o, nullptr, n);
}
template <typename T>
int scalar_op(pg_ctx* ctx, int op, void* out, const void* x, double c, int flags, int64_t n) {
  switch (op) {
    case PG_X_PLUS_C: return scalar_dispatch<T, PG_X_PLUS_C>(ctx, out, x, c, flags, n);
    case PG_X_MINUS_C: return scalar_dispatch<T, PG_X_MINUS_C>(ctx, out, x, c, flags, n);
    case PG_C_MINUS_X: return scalar_dispatch<T, PG_C_MINUS_X>(ctx, out, x, c, flags, n);
    case PG_X_TIMES_C: return scalar_dispatch<T, PG_X_TIMES_C>(ctx, out, x, c, flags, n);
    case PG_X_OVER_C: return scalar_dispatch<T, PG_X_OVER_C>(ctx, out, x, c, flags, n);
    case PG_C_OVER_X: return scalar_dispatch<T, PG_C_OVER_X>(ctx, out, x, c, flags, n);
  }
  pg_set_error("pg_vec_scalar: unknown op %d", op);
  return PG_ERR_ARG;
}

template <typename T, typename P>
int axpy_t(pg_ctx* ctx, void* out, const void* x, double s, const void* g, int sign, int64_t n, void* shadow) {
  Ptrs<2> in{{x, g}}; MPtrs<1> o{{out}};
  if (sign > 0) return ew_launch<T, 2, 1>(ctx, AxpyF<T, P, 1>{(P)s}, in, o, shadow, n);
  return ew_launch<T, 2, 1>(ctx, AxpyF<T, P, -1>{(P)s}, in, o, shadow, n);
}

}  // namespace

#define PG_DTYPE_FLOAT_ONLY(dtype) PG_REQUIRE((dtype) == PG_F32 || (dtype) == PG_F64, "dtype must be PG_F32 or PG_F64")

extern "C" {

int pg_vec_binary(pg_ctx* ctx, int dtype, int op, void* out, const void* x, const void* y, int64_t n) {
  PG_CHECK_CTX(ctx);
  if (dtype == PG_F16) return pg_f16_vec_binary(ctx, op, out, x, y, n);
  PG_DTYPE_FLOAT_ONLY(dtype);
  PG_REQUIRE(op == PG_ADD || op == PG_SUB, "unknown op");
  Ptrs<2> in{{x, y}}; MPtrs<1> o{{out}};
  if (dtype == PG_F32) {
    return op == PG_ADD ? ew_launch<float, 2, 1>(ctx, BinF<float, PG_ADD>{}, in, o, nullptr, n)
                        : ew_launch<float, 2, 1>(ctx, BinF<float, PG_SUB>{}, in, o, nullptr, n);
  }
  return op == PG_ADD ? ew_launch<double, 2, 1>(ctx, BinF<double, PG_ADD>{}, in, o, nullptr, n)
                      : ew_launch<double, 2, 1>(ctx, BinF<double, PG_SUB>{}, in, o, nullptr, n);
}

int pg_vec_scalar(pg_ctx* ctx, int dtype, int op, void* out, const void* x, double c, int flags, int64_t n) {
  PG_CHECK_CTX(ctx);
  if (dtype == PG_F16) return pg_f16_vec_scalar(ctx, op, out, x, c, flags, n);
  PG_DTYPE_FLOAT_ONLY(dtype);
  return dtype == PG_F32 ? scalar_op<float>(ctx, op, out, x, c, flags, n) : scalar_op<double>(ctx, op, out, x, c, flags, n);
}

int pg_vec_axpy(pg_ctx* ctx, int dtype, void* out, const void* x, double s, const void* g, int sign, int flags, int64_t n) {
  PG_CHECK_CTX(ctx);
  if (dtype == PG_F16) return pg_f16_vec_axpy(ctx, out, x, s, g, sign, flags, n);
  PG_DTYPE_FLOAT_ONLY(dtype);
  if (dtype == PG_F64) return axpy_t<double, double>(ctx, out, x, s, g, sign, n, nullptr);
  if (flags & PG_SCALAR_F64) return axpy_t<float, double>(ctx, out, x, s, g, sign, n, nullptr);
  return axpy_t<float, float>(ctx, out, x, s, g, sign, n, nullptr);
}

int pg_vec_fill(pg_ctx* ctx, int dtype, void* out, double c, int64_t n) {
  PG_CHECK_CTX(ctx);
  if (dtype == PG_F16) return pg_f16_vec_fill(ctx, out, c, n);
  PG_DTYPE_FLOAT_ONLY(dtype);
  Ptrs<0> in{}; MPtrs<1> o{{out}};
  if (dtype == PG_F32) return ew_launch<float, 0, 1>(ctx, FillF<float>{(float)c}, in, o, nullptr, n);
  return ew_launch<double, 0, 1>(ctx, FillF<double>{c}, in, o, nullptr, n);
}

int pg_vec_dot(pg_ctx* ctx, int dtype, const void* x, const void* y, int64_t n, double* result) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(result != nullptr, "null result");
  if (dtype == PG_F16) return pg_f16_vec_dot(ctx, x, y, n, result);
  PG_DTYPE_FLOAT_ONLY(dtype);
  PG_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "pointers must be 16-byte aligned");
  if (n <= 0) { *result = 0.0; return PG_OK; }
  int64_t blocks = pg_cdiv(pg_cdiv(n, 4), 256);
  int64_t cap = (int64_t)ctx->num_sms * 8;
  if (blocks > cap) blocks = cap;
  int rc = pg_ws_ensure(ctx, (size_t)blocks * sizeof(double));
  if (rc != PG_OK) return rc;
  if (dtype == PG_F32)
    dot_kernel<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const float*)x, (const float*)y, n, (double*)ctx->ws, ctx->counters, ctx->result_slot);
  else
    dot_kernel<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const double*)x, (const double*)y, n, (double*)ctx->ws, ctx->counters, ctx->result_slot);
  PG_LAUNCHED(ctx);
  PG_CHECK_CUDA(cudaMemcpyAsync(ctx->result_host, ctx->result_slot, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PG_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  // the reference returns the leaf eltype (test/test_models.jl:82): round through T
  *result = dtype == PG_F32 ? (double)(float)ctx->result_host[0] : ctx->result_host[0];
  return PG_OK;
}

// 8-bit pixels -> bf16(float(u8) / divisor): the device half of an input pipeline that keeps images as the bytes they are
// stored in (MNIST's IDX files, examples/01_mnist_mlp.jl:9-12 via MLDatasets) — bit-identical to casting the Float32 array
// the reference builds on the host, at a quarter of its upload bytes
__global__ void __launch_bounds__(256) cast_u8_bf16_kernel(const uint8_t* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t n, float divisor) {
  const int64_t n16 = n >> 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat162 o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[2 * j] = __floats2bfloat162_rn(__fdiv_rn((float)(w[j] & 0xffu), divisor), __fdiv_rn((float)((w[j] >> 8) & 0xffu), divisor));
      o[2 * j + 1] = __floats2bfloat162_rn(__fdiv_rn((float)((w[j] >> 16) & 0xffu), divisor), __fdiv_rn((float)(w[j] >> 24), divisor));
    }
    uint4* op = reinterpret_cast<uint4*>(out) + 2 * i;
    op[0] = *reinterpret_cast<const uint4*>(&o[0]);
    op[1] = *reinterpret_cast<const uint4*>(&o[4]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 15)) {
    const int64_t i = (n16 << 4) + threadIdx.x;
    out[i] = __float2bfloat16_rn(__fdiv_rn((float)x[i], divisor));
  }
}

int pg_vec_cast_u8_bf16(pg_ctx* ctx, void* out, const void* x, int64_t n, double divisor) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0, "pointers must be 16-byte aligned");
  PG_REQUIRE(divisor != 0.0, "zero divisor");
  if (n <= 0) return PG_OK;
  int64_t blocks = pg_cdiv(pg_cdiv(n, 16), 256);
  int64_t cap = (int64_t)ctx->num_sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cast_u8_bf16_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>((const uint8_t*)x, (__nv_bfloat16*)out, n, (float)divisor);
  PG_LAUNCHED(ctx);
  return PG_OK;
}

int pg_vec_cast_bf16(pg_ctx* ctx, void* out, const void* x, int64_t n) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0, "pointers must be 16-byte aligned");
  if (n <= 0) return PG_OK;
  int64_t blocks = pg_cdiv(pg_cdiv(n, 8), 256);
  int64_t cap = (int64_t)ctx->num_sms * 8;
  if (blocks > cap) blocks = cap;
  cast_bf16_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>((const float*)x, (__nv_bfloat16*)out, n);
  PG_LAUNCHED(ctx);
  return PG_OK;
}

int pg_opt_gd(pg_ctx* ctx, int dtype, void* p, const void* g, int64_t n, double s, int flags, void* shadow) {
  PG_CHECK_CTX(ctx); PG_DTYPE_FLOAT_ONLY(dtype);
  PG_REQUIRE(shadow == nullptr || dtype == PG_F32, "bf16 shadow needs an f32 arena");
  if (dtype == PG_F64) return axpy_t<double, double>(ctx, p, p, s, g, -1, n, nullptr);
  if (flags & PG_SCALAR_F64) return axpy_t<float, double>(ctx, p, p, s, g, -1, n, shadow);
  return axpy_t<float, float>(ctx, p, p, s, g, -1, n, shadow);
}

int pg_opt_heavyball(pg_ctx* ctx, int dtype, void* p, void* d, const void* g, int64_t n, double s, double mu, int flags, void* shadow) {
  PG_CHECK_CTX(ctx); PG_DTYPE_FLOAT_ONLY(dtype);
  PG_REQUIRE(shadow == nullptr || dtype == PG_F32, "bf16 shadow needs an f32 arena");
  Ptrs<3> in{{p, d, g}}; MPtrs<2> o{{p, d}};
  if (dtype == PG_F64) return ew_launch<double, 3, 2>(ctx, HeavyBallF<double, double, double>{s, mu}, in, o, nullptr, n);
  bool s64 = flags & PG_SCALAR_F64, m64 = flags & PG_SCALAR2_F64;
  if (s64 && m64) return ew_launch<float, 3, 2>(ctx, HeavyBallF<float, double, double>{s, mu}, in, o, shadow, n);
  if (s64) return ew_launch<float, 3, 2>(ctx, HeavyBallF<float, double, float>{s, (float)mu}, in, o, shadow, n);
  if (m64) return ew_launch<float, 3, 2>(ctx, HeavyBallF<float, float, double>{(float)s, mu}, in, o, shadow, n);
  return ew_launch<float, 3, 2>(ctx, HeavyBallF<float, float, float>{(float)s, (float)mu}, in, o, shadow, n);
}

static int nesterov_impl(pg_ctx* ctx, int dtype, void* p, void* prev, const void* g, int64_t n, double s, double mu, int flags, void* shadow, const double* dev) {
  PG_CHECK_CTX(ctx); PG_DTYPE_FLOAT_ONLY(dtype);
  PG_REQUIRE(shadow == nullptr || dtype == PG_F32, "bf16 shadow needs an f32 arena");
  Ptrs<2> in{{p, g}}; MPtrs<2> o{{p, prev}};
  if (dtype == PG_F64) return ew_launch<double, 2, 2>(ctx, NesterovF<double, double>{s, mu, dev}, in, o, nullptr, n);
  if (flags & PG_SCALAR_F64) return ew_launch<float, 2, 2>(ctx, NesterovF<float, double>{s, mu, dev}, in, o, shadow, n);
  return ew_launch<float, 2, 2>(ctx, NesterovF<float, float>{(float)s, mu, dev}, in, o, shadow, n);
}

int pg_opt_nesterov(pg_ctx* ctx, int dtype, void* p, void* prev, const void* g, int64_t n, double s, double mu, int flags, void* shadow) {
  return nesterov_impl(ctx, dtype, p, prev, g, n, s, mu, flags, shadow, nullptr);
}

int pg_opt_nesterov_dev(pg_ctx* ctx, int dtype, void* p, void* prev, const void* g, int64_t n, double s, double* scalars_dev, int flags, void* shadow) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(scalars_dev != nullptr, "null device scalars");
  opt_tick_kernel<<<1, 1, 0, ctx->stream>>>(scalars_dev, 1, 0.0, 0.0);
  PG_LAUNCHED(ctx);
  return nesterov_impl(ctx, dtype, p, prev, g, n, s, 0.0, flags, shadow, scalars_dev);
}

static int adam_impl(pg_ctx* ctx, int dtype, void* p, void* m, void* v, void* m_hat, void* v_hat, const void* g, int64_t n,
                     double s, double b1, double b2, double eps, int64_t it, int flags, void* shadow, const double* dev) {
  PG_CHECK_CTX(ctx); PG_DTYPE_FLOAT_ONLY(dtype);
  PG_REQUIRE(it >= 1, "Adam iteration counter starts at 1 (adam.jl:28)");
  PG_REQUIRE(shadow == nullptr || dtype == PG_F32, "bf16 shadow needs an f32 arena");
  Ptrs<4> in{{p, m, v, g}}; MPtrs<5> o{{p, m, v, m_hat, v_hat}};
  // Julia: beta^it with an Int exponent (power by squaring); host pow agrees to <= 1 ulp.
  double c1 = 1.0 - pow(b1, (double)it), c2 = 1.0 - pow(b2, (double)it);
  if (dtype == PG_F64)
    return ew_launch<double, 4, 5>(ctx, AdamF<double, double>{s, b1, 1.0 - b1, b2, 1.0 - b2, c1, c2, eps, dev}, in, o, nullptr, n);
  if (flags & PG_MATH_FAST)
    return ew_launch<float, 4, 5>(ctx, AdamFastF{(float)s, (float)b1, (float)(1.0 - b1), (float)b2, (float)(1.0 - b2), (float)(1.0 / c1), (float)(1.0 / c2), (float)eps, dev}, in, o, shadow, n);
  if (flags & PG_SCALAR_F64)
    return ew_launch<float, 4, 5>(ctx, AdamF<float, double>{s, b1, 1.0 - b1, b2, 1.0 - b2, c1, c2, eps, dev}, in, o, shadow, n);
  float fb1 = (float)b1, fb2 = (float)b2;
  // Float32 hyper-parameters: Julia computes 1 - beta and beta^it in Float32
  float fc1 = 1.0f - powf(fb1, (float)it), fc2 = 1.0f - powf(fb2, (float)it);
  return ew_launch<float, 4, 5>(ctx, AdamF<float, float>{(float)s, fb1, 1.0f - fb1, fb2, 1.0f - fb2, fc1, fc2, (float)eps, dev}, in, o, shadow, n);
}

int pg_opt_adam(pg_ctx* ctx, int dtype, void* p, void* m, void* v, void* m_hat, void* v_hat, const void* g, int64_t n,
                double s, double b1, double b2, double eps, int64_t it, int flags, void* shadow) {
  return adam_impl(ctx, dtype, p, m, v, m_hat, v_hat, g, n, s, b1, b2, eps, it, flags, shadow, nullptr);
}

int pg_opt_adam_dev(pg_ctx* ctx, int dtype, void* p, void* m, void* v, void* m_hat, void* v_hat, const void* g, int64_t n,
                    double s, double b1, double b2, double eps, double* scalars_dev, int flags, void* shadow) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(scalars_dev != nullptr, "null device scalars");
  opt_tick_kernel<<<1, 1, 0, ctx->stream>>>(scalars_dev, 0, b1, b2);
  PG_LAUNCHED(ctx);
  return adam_impl(ctx, dtype, p, m, v, m_hat, v_hat, g, n, s, b1, b2, eps, 1, flags, shadow, scalars_dev);
}

}  // extern "C"
