// arena.cu — parameter arenas behind the C ABI (pg_arena_*), the generic Model-broadcast interpreter (pg_vec_expr,
// src/model.jl:148-164 for ANY flattened scalar function f) and Float16 arenas (test/test_models.jl:18 runs the whole
// algebra matrix on Float16): the Float16 forms of pg_vec_binary / scalar / axpy / fill / dot are lowered onto the same
// interpreter kernel.
//
// Numerics of the interpreter: every stack value carries a promotion level — the leaf
// eltype T (Float16 < Float32 < Float64) or the type of the Julia scalar it came from; an op is evaluated at the
// higher level of its operands with IEEE round-to-nearest (no FMA contraction) and the result is rounded to T once,
// on store.  Float16 arithmetic is evaluated in Float32 and rounded to Float16 after every op, which is what Julia
// (and numpy) do and is exact-equivalent to native half arithmetic for + - * / sqrt.
#include <cuda_fp16.h>

#include <vector>

#include "common.cuh"

struct pg_arena {
  pg_ctx* ctx = nullptr;
  int dtype = PG_F32;
  std::vector<int64_t> sizes, offsets;    // elements
  int64_t length = 0, n_padded = 0;
  void* base = nullptr;
  bool owner = true;
};

namespace {

constexpr int PAD = 32;                 // leaf alignment in elements (128 B for Float32: TMA / float4 aligned)
constexpr int MAX_ARGS = 8, MAX_CODE = 96, MAX_CONSTS = 16, STACK = 12;
inline int64_t pad_up(int64_t n) { return (n + PAD - 1) / PAD * PAD; }
inline size_t esize(int dt) { return dt == PG_F64 ? 8 : dt == PG_F16 ? 2 : 4; }

struct Program {
  const void* args[MAX_ARGS];
  int32_t code[MAX_CODE];
  double consts[MAX_CONSTS];
  int n_code;
};

template <typename T> struct Lvl;
template <> struct Lvl<__half> { static constexpr int v = 0; };
template <> struct Lvl<float> { static constexpr int v = 1; };
template <> struct Lvl<double> { static constexpr int v = 2; };

__device__ __forceinline__ double round_to_level(double x, int lvl) {
  if (lvl == 0) return (double)__half2float(__float2half_rn((float)x));   // (exact values only pass through here)
  if (lvl == 1) return (double)(float)x;
  return x;
}
template <typename T> __device__ __forceinline__ double load_as_double(const void* p, int64_t i);
template <> __device__ __forceinline__ double load_as_double<__half>(const void* p, int64_t i) { return (double)__half2float(((const __half*)p)[i]); }
template <> __device__ __forceinline__ double load_as_double<float>(const void* p, int64_t i) { return (double)((const float*)p)[i]; }
template <> __device__ __forceinline__ double load_as_double<double>(const void* p, int64_t i) { return ((const double*)p)[i]; }
template <typename T> __device__ __forceinline__ void store_from_double(void* p, int64_t i, double v);
// double -> half in ONE rounding (through float would round twice)
template <> __device__ __forceinline__ void store_from_double<__half>(void* p, int64_t i, double v) { ((__half*)p)[i] = __double2half(v); }
template <> __device__ __forceinline__ void store_from_double<float>(void* p, int64_t i, double v) { ((float*)p)[i] = (float)v; }
template <> __device__ __forceinline__ void store_from_double<double>(void* p, int64_t i, double v) { ((double*)p)[i] = v; }

__device__ __forceinline__ double binop(int op, double a, double b, int lvl) {
  if (lvl == 2) {
    switch (op) {
      case PG_BC_ADD: return __dadd_rn(a, b);
      case PG_BC_SUB: return __dsub_rn(a, b);
      case PG_BC_MUL: return __dmul_rn(a, b);
      case PG_BC_DIV: return __ddiv_rn(a, b);
      case PG_BC_MAX: return fmax(a, b);
      case PG_BC_MIN: return fmin(a, b);
      default: return pow(a, b);
    }
  }
  const float x = (float)a, y = (float)b;
  float r;
  switch (op) {
    case PG_BC_ADD: r = __fadd_rn(x, y); break;
    case PG_BC_SUB: r = __fsub_rn(x, y); break;
    case PG_BC_MUL: r = __fmul_rn(x, y); break;
    case PG_BC_DIV: r = __fdiv_rn(x, y); break;
    case PG_BC_MAX: r = fmaxf(x, y); break;
    case PG_BC_MIN: r = fminf(x, y); break;
    default: r = powf(x, y); break;
  }
  return lvl == 0 ? (double)__half2float(__float2half_rn(r)) : (double)r;
}
__device__ __forceinline__ double unop(int op, double a, int lvl) {
  if (lvl == 2) {
    switch (op) {
      case PG_BC_NEG: return -a;
      case PG_BC_SQRT: return __dsqrt_rn(a);
      case PG_BC_ABS: return fabs(a);
      case PG_BC_EXP: return exp(a);
      case PG_BC_LOG: return log(a);
      case PG_BC_SQUARE: return __dmul_rn(a, a);
      case PG_BC_RELU: return a < 0.0 ? 0.0 : a;
      default: return __ddiv_rn(1.0, a);
    }
  }
  const float x = (float)a;
  float r;
  switch (op) {
    case PG_BC_NEG: r = -x; break;
    case PG_BC_SQRT: r = __fsqrt_rn(x); break;
    case PG_BC_ABS: r = fabsf(x); break;
    case PG_BC_EXP: r = expf(x); break;
    case PG_BC_LOG: r = logf(x); break;
    case PG_BC_SQUARE: r = __fmul_rn(x, x); break;
    case PG_BC_RELU: r = x < 0.0f ? 0.0f : x; break;
    default: r = __fdiv_rn(1.0f, x); break;
  }
  return lvl == 0 ? (double)__half2float(__float2half_rn(r)) : (double)r;
}

template <typename T>
__global__ void __launch_bounds__(256) expr_kernel(Program pr, void* __restrict__ out, int64_t n) {
  constexpr int TL = Lvl<T>::v;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double st[STACK];
    int lv[STACK];
    int sp = 0;
    for (int pc = 0; pc < pr.n_code; ++pc) {
      const int w = pr.code[pc], op = w & 0xff, k = w >> 8;
      if (op == PG_BC_ARG) { st[sp] = load_as_double<T>(pr.args[k], i); lv[sp++] = TL; }
      else if (op == PG_BC_CONST) { st[sp] = pr.consts[k]; lv[sp++] = 2; }
      else if (op == PG_BC_CONST_T) { st[sp] = round_to_level(pr.consts[k], TL); lv[sp++] = TL; }
      else if (op == PG_BC_CONST_F32) { st[sp] = (double)(float)pr.consts[k]; lv[sp++] = TL > 1 ? TL : 1; }
      else if (op < PG_BC_NEG) {
        const int l = lv[sp - 1] > lv[sp - 2] ? lv[sp - 1] : lv[sp - 2];
        st[sp - 2] = binop(op, st[sp - 2], st[sp - 1], l);
        lv[sp - 2] = l;
        --sp;
      } else {
        st[sp - 1] = unop(op, st[sp - 1], lv[sp - 1]);
      }
    }
    store_from_double<T>(out, i, st[0]);
  }
}

// host-side validation: stack discipline, operand ranges
int check_program(const int32_t* code, int n_code, int n_args, int n_consts) {
  PG_REQUIRE(n_code > 0 && n_code <= MAX_CODE, "program length must be 1..96");
  PG_REQUIRE(n_args >= 0 && n_args <= MAX_ARGS && n_consts >= 0 && n_consts <= MAX_CONSTS, "at most 8 arena operands and 16 constants");
  int sp = 0;
  for (int pc = 0; pc < n_code; ++pc) {
    const int op = code[pc] & 0xff, k = code[pc] >> 8;
    if (op == PG_BC_ARG) { PG_REQUIRE(k >= 0 && k < n_args, "operand index out of range"); ++sp; }
    else if (op == PG_BC_CONST || op == PG_BC_CONST_T || op == PG_BC_CONST_F32) { PG_REQUIRE(k >= 0 && k < n_consts, "constant index out of range"); ++sp; }
    else if (op >= PG_BC_ADD && op <= PG_BC_POW) { PG_REQUIRE(sp >= 2, "stack underflow"); --sp; }
    else if (op >= PG_BC_NEG && op <= PG_BC_INV) { PG_REQUIRE(sp >= 1, "stack underflow"); }
    else { pg_set_error("pg_vec_expr: unknown opcode %d", op); return PG_ERR_ARG; }
    PG_REQUIRE(sp <= STACK, "expression too deep (stack of 12)");
  }
  PG_REQUIRE(sp == 1, "program must leave exactly one value");
  return PG_OK;
}

int zero_padding(pg_ctx* ctx, int dtype, void* base, int n_leaves, const int64_t* sizes) {
  int64_t off = 0;
  const size_t es = esize(dtype);
  for (int l = 0; l < n_leaves; ++l) {
    const int64_t p = pad_up(sizes[l]);
    if (p > sizes[l]) PG_CHECK_CUDA(cudaMemsetAsync((char*)base + (off + sizes[l]) * es, 0, (size_t)(p - sizes[l]) * es, ctx->stream));
    off += p;
  }
  return PG_OK;
}

template <typename T>
__global__ void __launch_bounds__(256) dot_any_kernel(const T* __restrict__ x, const T* __restrict__ y, int64_t n, double* partials, unsigned int* counter, double* result) {
  __shared__ double sm[32];
  __shared__ bool is_last;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += load_as_double<T>(x, i) * load_as_double<T>(y, i);
  const double bs = block_sum(acc, sm);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = bs;
    __threadfence();
    is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double a2 = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) a2 += partials[i];
    const double tot = block_sum(a2, sm);
    if (threadIdx.x == 0) { *result = tot; *counter = 0u; }
  }
}

}  // namespace

extern "C" int pg_vec_expr(pg_ctx* ctx, int dtype, void* out, const void* const* args, int n_args, const int32_t* code, int n_code,
                           const double* consts, int n_consts, int64_t n, int n_leaves, const int64_t* leaf_sizes) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(dtype == PG_F16 || dtype == PG_F32 || dtype == PG_F64, "dtype must be PG_F16, PG_F32 or PG_F64");
  PG_REQUIRE(out != nullptr && code != nullptr, "null pointer");
  int rc = check_program(code, n_code, n_args, n_consts);
  if (rc != PG_OK) return rc;
  if (n <= 0) return PG_OK;
  Program pr;
  memset(&pr, 0, sizeof(pr));
  for (int k = 0; k < n_args; ++k) { PG_REQUIRE(args[k] != nullptr, "null operand"); pr.args[k] = args[k]; }
  for (int k = 0; k < n_code; ++k) pr.code[k] = code[k];
  for (int k = 0; k < n_consts; ++k) pr.consts[k] = consts[k];
  pr.n_code = n_code;
  int64_t blocks = pg_cdiv(n, 256);
  const int64_t cap = (int64_t)ctx->num_sms * 8;
  if (blocks > cap) blocks = cap;
  if (dtype == PG_F16) expr_kernel<__half><<<(unsigned)blocks, 256, 0, ctx->stream>>>(pr, out, n);
  else if (dtype == PG_F32) expr_kernel<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(pr, out, n);
  else expr_kernel<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(pr, out, n);
  PG_LAUNCHED(ctx);
  if (n_leaves > 0 && leaf_sizes != nullptr) return zero_padding(ctx, dtype, out, n_leaves, leaf_sizes);
  return PG_OK;
}

// ---- Float16 forms of the fixed-function algebra entry points (called from arena_kernels.cu) ----------------------
int pg_f16_vec_binary(pg_ctx* ctx, int op, void* out, const void* x, const void* y, int64_t n) {
  const void* args[2] = {x, y};
  const int32_t code[3] = {PG_BC_ARG | (0 << 8), PG_BC_ARG | (1 << 8), op == PG_ADD ? PG_BC_ADD : PG_BC_SUB};
  return pg_vec_expr(ctx, PG_F16, out, args, 2, code, 3, nullptr, 0, n, 0, nullptr);
}
int pg_f16_vec_scalar(pg_ctx* ctx, int op, void* out, const void* x, double c, int flags, int64_t n) {
  const void* args[1] = {x};
  const int32_t cst = ((flags & PG_SCALAR_F64) ? PG_BC_CONST : PG_BC_CONST_T) | (0 << 8), arg = PG_BC_ARG | (0 << 8);
  int32_t code[3];
  switch (op) {
    case PG_X_PLUS_C: code[0] = arg; code[1] = cst; code[2] = PG_BC_ADD; break;
    case PG_X_MINUS_C: code[0] = arg; code[1] = cst; code[2] = PG_BC_SUB; break;
    case PG_C_MINUS_X: code[0] = cst; code[1] = arg; code[2] = PG_BC_SUB; break;
    case PG_X_TIMES_C: code[0] = arg; code[1] = cst; code[2] = PG_BC_MUL; break;
    case PG_X_OVER_C: code[0] = arg; code[1] = cst; code[2] = PG_BC_DIV; break;
    case PG_C_OVER_X: code[0] = cst; code[1] = arg; code[2] = PG_BC_DIV; break;
    default: pg_set_error("pg_vec_scalar: unknown op %d", op); return PG_ERR_ARG;
  }
  return pg_vec_expr(ctx, PG_F16, out, args, 1, code, 3, &c, 1, n, 0, nullptr);
}
int pg_f16_vec_axpy(pg_ctx* ctx, void* out, const void* x, double s, const void* g, int sign, int flags, int64_t n) {
  const void* args[2] = {x, g};
  const int32_t code[5] = {PG_BC_ARG | (0 << 8), ((flags & PG_SCALAR_F64) ? PG_BC_CONST : PG_BC_CONST_T) | (0 << 8), PG_BC_ARG | (1 << 8), PG_BC_MUL,
                           sign > 0 ? PG_BC_ADD : PG_BC_SUB};
  return pg_vec_expr(ctx, PG_F16, out, args, 2, code, 5, &s, 1, n, 0, nullptr);
}
int pg_f16_vec_fill(pg_ctx* ctx, void* out, double c, int64_t n) {
  const int32_t code[1] = {PG_BC_CONST_T | (0 << 8)};
  return pg_vec_expr(ctx, PG_F16, out, nullptr, 0, code, 1, &c, 1, n, 0, nullptr);
}
int pg_f16_vec_dot(pg_ctx* ctx, const void* x, const void* y, int64_t n, double* result) {
  if (n <= 0) { *result = 0.0; return PG_OK; }
  int64_t blocks = pg_cdiv(n, 256);
  const int64_t cap = (int64_t)ctx->num_sms * 8;
  if (blocks > cap) blocks = cap;
  int rc = pg_ws_ensure(ctx, (size_t)blocks * sizeof(double));
  if (rc != PG_OK) return rc;
  dot_any_kernel<__half><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const __half*)x, (const __half*)y, n, (double*)ctx->ws, ctx->counters, ctx->result_slot);
  PG_LAUNCHED(ctx);
  PG_CHECK_CUDA(cudaMemcpyAsync(ctx->result_host, ctx->result_slot, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PG_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  *result = (double)__half2float(__double2half(ctx->result_host[0]));   // the reference returns the leaf eltype (test/test_models.jl:82)
  return PG_OK;
}

// ---- arenas ----------------------------------------------------------------------------------------------------------
extern "C" {

int pg_arena_create(pg_ctx* ctx, int dtype, int n_leaves, const int64_t* leaf_sizes, pg_arena** out) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(out != nullptr, "null out pointer");
  *out = nullptr;
  PG_REQUIRE(dtype == PG_F16 || dtype == PG_F32 || dtype == PG_F64, "arena eltype must be Float16, Float32 or Float64");
  PG_REQUIRE(n_leaves >= 0 && (n_leaves == 0 || leaf_sizes != nullptr), "bad leaf list");
  pg_arena* a = new pg_arena();
  a->ctx = ctx; a->dtype = dtype;
  for (int l = 0; l < n_leaves; ++l) {
    if (leaf_sizes[l] < 0) { delete a; pg_set_error("pg_arena_create: negative leaf size"); return PG_ERR_ARG; }
    a->sizes.push_back(leaf_sizes[l]);
    a->offsets.push_back(a->n_padded);
    a->length += leaf_sizes[l];
    a->n_padded += pad_up(leaf_sizes[l]);
  }
  int rc = pg_malloc(ctx, (size_t)a->n_padded * esize(dtype), &a->base);
  if (rc == PG_OK) rc = pg_memset0(ctx, a->base, (size_t)a->n_padded * esize(dtype));
  if (rc != PG_OK) { delete a; return rc; }
  *out = a;
  return PG_OK;
}

int pg_arena_wrap(pg_ctx* ctx, int dtype, int n_leaves, const int64_t* leaf_sizes, void* base, pg_arena** out) {
  PG_CHECK_CTX(ctx);
  PG_REQUIRE(out != nullptr && base != nullptr, "null pointer");
  *out = nullptr;
  PG_REQUIRE(dtype == PG_F16 || dtype == PG_F32 || dtype == PG_F64, "arena eltype must be Float16, Float32 or Float64");
  PG_REQUIRE(n_leaves > 0 && leaf_sizes != nullptr, "bad leaf list");
  pg_arena* a = new pg_arena();
  a->ctx = ctx; a->dtype = dtype; a->base = base; a->owner = false;
  for (int l = 0; l < n_leaves; ++l) {
    a->sizes.push_back(leaf_sizes[l]);
    a->offsets.push_back(a->n_padded);
    a->length += leaf_sizes[l];
    a->n_padded += pad_up(leaf_sizes[l]);
  }
  *out = a;
  return PG_OK;
}

int pg_arena_destroy(pg_arena* a) {
  if (!a) return PG_OK;
  if (a->base && a->owner) { cudaStreamSynchronize(a->ctx->stream); cudaFree(a->base); }
  delete a;
  return PG_OK;
}

int pg_arena_like(pg_arena* src, int mode, pg_arena** out) {
  PG_REQUIRE(src != nullptr && out != nullptr, "null pointer");
  PG_REQUIRE(mode >= 0 && mode <= 2, "mode: 0 similar, 1 zero, 2 copy");
  int rc = pg_arena_create(src->ctx, src->dtype, (int)src->sizes.size(), src->sizes.data(), out);
  if (rc != PG_OK) return rc;
  if (mode == 2) return pg_copy_d2d(src->ctx, (*out)->base, src->base, (size_t)src->n_padded * esize(src->dtype));
  return PG_OK;
}

int pg_arena_info(pg_arena* a, int* dtype, int* n_leaves, int64_t* length, int64_t* n_padded, void** base) {
  PG_REQUIRE(a != nullptr, "null arena");
  if (dtype) *dtype = a->dtype;
  if (n_leaves) *n_leaves = (int)a->sizes.size();
  if (length) *length = a->length;
  if (n_padded) *n_padded = a->n_padded;
  if (base) *base = a->base;
  return PG_OK;
}

int pg_arena_leaf_ptr(pg_arena* a, int leaf, void** dptr, int64_t* size) {
  PG_REQUIRE(a != nullptr && dptr != nullptr, "null pointer");
  PG_REQUIRE(leaf >= 0 && leaf < (int)a->sizes.size(), "leaf index out of range");
  *dptr = (char*)a->base + (size_t)a->offsets[leaf] * esize(a->dtype);
  if (size) *size = a->sizes[leaf];
  return PG_OK;
}

static int arena_xfer(pg_arena* a, int leaf, void* host, bool up) {
  PG_REQUIRE(a != nullptr && host != nullptr, "null pointer");
  PG_REQUIRE(leaf >= -1 && leaf < (int)a->sizes.size(), "leaf index out of range");
  const size_t es = esize(a->dtype);
  const int l0 = leaf < 0 ? 0 : leaf, l1 = leaf < 0 ? (int)a->sizes.size() : leaf + 1;
  char* h = (char*)host;
  for (int l = l0; l < l1; ++l) {
    char* d = (char*)a->base + (size_t)a->offsets[l] * es;
    const size_t bytes = (size_t)a->sizes[l] * es;
    if (bytes) PG_CHECK_CUDA(cudaMemcpyAsync(up ? (void*)d : (void*)h, up ? (const void*)h : (const void*)d, bytes, up ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, a->ctx->stream));
    h += bytes;
  }
  PG_CHECK_CUDA(cudaStreamSynchronize(a->ctx->stream));
  return PG_OK;
}
int pg_arena_upload(pg_arena* a, int leaf, const void* host) { return arena_xfer(a, leaf, const_cast<void*>(host), true); }
int pg_arena_download(pg_arena* a, int leaf, void* host) { return arena_xfer(a, leaf, host, false); }

static int arena_locate(pg_arena* a, int64_t i, void** p) {
  PG_REQUIRE(a != nullptr, "null arena");
  if (i < 0) i += a->length;                      // m[end-1] style access from the host language
  PG_REQUIRE(i >= 0 && i < a->length, "BoundsError: flat index out of range");
  for (size_t l = 0; l < a->sizes.size(); ++l) {
    if (i < a->sizes[l]) { *p = (char*)a->base + (size_t)(a->offsets[l] + i) * esize(a->dtype); return PG_OK; }
    i -= a->sizes[l];
  }
  return PG_ERR_ARG;
}
int pg_arena_get_scalar(pg_arena* a, int64_t i, double* value) {
  void* p = nullptr;
  int rc = arena_locate(a, i, &p);
  if (rc != PG_OK) return rc;
  PG_REQUIRE(value != nullptr, "null out pointer");
  double buf = 0;
  PG_CHECK_CUDA(cudaMemcpyAsync(&buf, p, esize(a->dtype), cudaMemcpyDeviceToHost, a->ctx->stream));
  PG_CHECK_CUDA(cudaStreamSynchronize(a->ctx->stream));
  if (a->dtype == PG_F64) *value = buf;
  else if (a->dtype == PG_F32) *value = (double)*(float*)&buf;
  else *value = (double)__half2float(*(__half*)&buf);
  return PG_OK;
}
int pg_arena_set_scalar(pg_arena* a, int64_t i, double value) {
  void* p = nullptr;
  int rc = arena_locate(a, i, &p);
  if (rc != PG_OK) return rc;
  double buf = 0;
  if (a->dtype == PG_F64) buf = value;
  else if (a->dtype == PG_F32) *(float*)&buf = (float)value;
  else *(__half*)&buf = __double2half(value);
  PG_CHECK_CUDA(cudaMemcpyAsync(p, &buf, esize(a->dtype), cudaMemcpyHostToDevice, a->ctx->stream));
  PG_CHECK_CUDA(cudaStreamSynchronize(a->ctx->stream));
  return PG_OK;
}

}  // extern "C"

// used by dp.cu
void* pg_arena_base(pg_arena* a) { return a->base; }
int pg_arena_dtype(pg_arena* a) { return a->dtype; }
int64_t pg_arena_padded(pg_arena* a) { return a->n_padded; }
const std::vector<int64_t>& pg_arena_offsets(pg_arena* a) { return a->offsets; }
