/* protograd_b200.h — C ABI of the B200-native ProtoGrad training-step library.
 *
 * The reference (lostella/ProtoGrad.jl) has no FFI: its extension seams are Julia multiple
 * dispatch at four places.  This header is what Julia glue behind those seams
 * `ccall`s (see INTEGRATION.md and julia/ProtoGradB200.jl); each entry point cites the
 * reference code it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *  - every call returns int: 0 = ok, negative = error; pg_last_error() gives the message
 *    (the reference's only error idiom is `error(...)`, src/gradient.jl:1 — the glue rethrows).
 *  - all device work is asynchronous on the context's stream unless the name says otherwise
 *    (pg_sync, pg_download, pg_loss_read, pg_vec_dot are the synchronising calls).
 *  - all arrays are row-major / C order == the bytes Julia holds for its column-major arrays:
 *      Linear: X[M][K] (M = batch), W[K][N] (Julia W[out,in], out contiguous), b[N], Y[M][N]
 *      Conv:   x[N][Cin][H][W] (Julia x[W,H,Cin,N]), w[Cout][Cin][kH][kW], y[N][Cout][H'][W']
 *  - scalars cross the ABI as double; `flags & PG_SCALAR_F64` says the Julia scalar was a
 *    Float64 (per-element math is then done in Float64 and rounded once on store, which is what
 *    Julia's broadcast promotion does — SURVEY.md Appendix A.6); otherwise math stays in T.
 *  - there is NO CPU fallback: without a CUDA device every compute call fails with PG_ERR_CUDA.
 */
#ifndef PROTOGRAD_B200_H
#define PROTOGRAD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pg_ctx pg_ctx;

typedef struct pg_plan pg_plan;
typedef struct pg_arena pg_arena;
typedef struct pg_dp pg_dp;

typedef enum { PG_F32 = 0, PG_F64 = 1, PG_BF16 = 2, PG_F16 = 3, PG_I32 = 4 } pg_dtype;
typedef enum { PG_ACT_NONE = 0, PG_ACT_RELU = 1 } pg_act;
typedef enum { PG_OK = 0, PG_ERR_ARG = -1, PG_ERR_CUDA = -2, PG_ERR_UNSUPPORTED = -3, PG_ERR_NOMEM = -4 } pg_status;
typedef enum { PG_ADD = 0, PG_SUB = 1 } pg_binop;
/* model (op) scalar forms of src/model.jl:168-196 */
typedef enum { PG_X_PLUS_C = 0, PG_X_MINUS_C = 1, PG_C_MINUS_X = 2, PG_X_TIMES_C = 3, PG_X_OVER_C = 4, PG_C_OVER_X = 5 } pg_scalarop;
/* flags of the scalar-taking calls: PG_SCALAR_F64 / PG_SCALAR2_F64 = the first / second Julia scalar was a Float64
 * (per-element math in Float64, one rounding on store); PG_MATH_FAST = fused all-Float32 Adam arithmetic
 * (normwise ~1e-7 of the promoted form) instead of the bit-exact Float64-promoted one. */
enum { PG_SCALAR_F64 = 1, PG_SCALAR2_F64 = 2, PG_MATH_FAST = 4 };

/* ---- lifecycle -------------------------------------------------------------------------- */
const char* pg_last_error(void);
int pg_version(void);
int pg_device_count(int* n);
int pg_init(int device, pg_ctx** ctx);           /* owns one stream + scratch on `device` */
int pg_destroy(pg_ctx* ctx);
int pg_set_stream(pg_ctx* ctx, void* cuda_stream); /* borrow an external stream (NULL = own stream) */
int pg_get_stream(pg_ctx* ctx, void** cuda_stream);
int pg_sync(pg_ctx* ctx);
int pg_launch_count(pg_ctx* ctx, int64_t* n);    /* kernels launched through this ctx so far */
int pg_reserve_workspace(pg_ctx* ctx, size_t bytes); /* pre-size scratch (needed before graph capture) */
int pg_set_sm_limit(pg_ctx* ctx, int n_sms);     /* size persistent grids for n_sms SMs (0 = all, negative = all but |n|): leaves SMs to a concurrent NCCL kernel */

/* ---- device / pinned memory (arena storage; replaces Julia Array allocation for leaves) ---- */
int pg_malloc(pg_ctx* ctx, size_t bytes, void** dptr);
int pg_free(pg_ctx* ctx, void* dptr);
int pg_memset0(pg_ctx* ctx, void* dptr, size_t bytes);
int pg_copy_d2d(pg_ctx* ctx, void* dst, const void* src, size_t bytes);          /* copy(m), model.jl:49-56 */
int pg_upload(pg_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);   /* async H2D */
int pg_download(pg_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes); /* D2H + sync: vec(m), model.jl:33 */
int pg_download_async(pg_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
int pg_host_alloc(size_t bytes, void** p);       /* pinned host memory */
int pg_host_free(void* p);

/* ---- input prefetch: a second (copy) stream + events, so the H2D copy of the next minibatch
 * (the reference's `train_x_flat[:, idx]` gather, examples/01_mnist_mlp.jl:35-38) overlaps the step.
 * which_stream: 0 = the compute stream, 1 = the context's copy stream. ---- */
int pg_event_create(pg_ctx* ctx, void** event);
int pg_event_destroy(pg_ctx* ctx, void* event);
int pg_event_record(pg_ctx* ctx, void* event, int which_stream);
int pg_stream_wait_event(pg_ctx* ctx, int which_stream, void* event);
int pg_event_sync(pg_ctx* ctx, void* event);      /* block the host until the event has completed */
int pg_upload_on(pg_ctx* ctx, int which_stream, void* dst_dev, const void* src_host, size_t bytes);

/* ---- CUDA-graph capture of a whole step (stream capture on the ctx stream) ---------------- */
int pg_graph_begin(pg_ctx* ctx);
int pg_graph_end(pg_ctx* ctx, void** graph_exec);
int pg_graph_launch(pg_ctx* ctx, void* graph_exec);
int pg_graph_destroy(pg_ctx* ctx, void* graph_exec);

/* ---- Model vector-space algebra over a flat arena (src/model.jl:103-202) ------------------ */
/* out = x (+|-) y              model.jl:168-172, broadcast copyto! :148-152 */
int pg_vec_binary(pg_ctx* ctx, int dtype, int op, void* out, const void* x, const void* y, int64_t n);
/* out = x op c  /  c op x      model.jl:174-196 */
int pg_vec_scalar(pg_ctx* ctx, int dtype, int op, void* out, const void* x, double c, int flags, int64_t n);
/* out = x + sign*s*g  (sign = +1 | -1):  `m .= m .- s .* grad`, test/test_models.jl:154, README.md:148 */
int pg_vec_axpy(pg_ctx* ctx, int dtype, void* out, const void* x, double s, const void* g, int sign, int flags, int64_t n);
int pg_vec_fill(pg_ctx* ctx, int dtype, void* out, double c, int64_t n);          /* zero(m), model.jl:49-56 */
/* dot(m1, m2), model.jl:200-202.  Synchronising; result returned as double (exact for T). */
int pg_vec_dot(pg_ctx* ctx, int dtype, const void* x, const void* y, int64_t n, double* result);
/* f32 -> bf16 shadow of an arena (operands of the tensor-core path) */
int pg_vec_cast_bf16(pg_ctx* ctx, void* out_bf16, const void* x_f32, int64_t n);
/* out = bf16(Float32(x_u8) / divisor): 8-bit pixels as the dataset stores them (the reference's examples read MNIST through
 * MLDatasets, /root/reference/examples/01_mnist_mlp.jl:9-12, which yields Float32(pixel / 255)); bit-identical to
 * pg_vec_cast_bf16 of that Float32 array, at a quarter of the upload bytes */
int pg_vec_cast_u8_bf16(pg_ctx* ctx, void* out_bf16, const void* x_u8, int64_t n, double divisor);

/* ---- Optimizers: one fused pass over the arena (src/optimizers/ *.jl) ---------------------- */
/* `shadow_bf16` (nullable): also write bf16(p_new) at the same index (tensor-core operand copy). */
/* gradient_descent.jl:15-17   p -= s*g */
int pg_opt_gd(pg_ctx* ctx, int dtype, void* p, const void* g, int64_t n, double stepsize, int flags, void* shadow_bf16);
/* heavy_ball_method.jl:16-19  d = mu*d - s*g ; p += d */
int pg_opt_heavyball(pg_ctx* ctx, int dtype, void* p, void* d, const void* g, int64_t n, double stepsize, double momentum, int flags, void* shadow_bf16);
/* nesterov_method.jl:33-39    prev = p ; p -= s*g ; p = p + mu*(p - prev)   (mu from the host-side sequence :3-6) */
int pg_opt_nesterov(pg_ctx* ctx, int dtype, void* p, void* prev, const void* g, int64_t n, double stepsize, double mu, int flags, void* shadow_bf16);
/* adam.jl:32-41; `it` is the value BEFORE the increment (starts at 1, adam.jl:28).
 * m_hat / v_hat nullable: when given they are materialised like the reference's state fields. */
int pg_opt_adam(pg_ctx* ctx, int dtype, void* p, void* m, void* v, void* m_hat, void* v_hat, const void* g, int64_t n,
                double stepsize, double beta1, double beta2, double epsilon, int64_t it, int flags, void* shadow_bf16);

/* Graph-replayable forms: the per-step scalars live in `scalars_dev` (5 doubles: Adam it, 1-beta1^it,
 * 1-beta2^it, Nesterov t, momentum; initialise to {1,0,0,1,0}) and are advanced by a one-thread kernel
 * in front of the update, so a captured CUDA graph can be replayed without re-baking kernel arguments. */
int pg_opt_adam_dev(pg_ctx* ctx, int dtype, void* p, void* m, void* v, void* m_hat, void* v_hat, const void* g, int64_t n,
                    double stepsize, double beta1, double beta2, double epsilon, double* scalars_dev, int flags, void* shadow_bf16);
int pg_opt_nesterov_dev(pg_ctx* ctx, int dtype, void* p, void* prev, const void* g, int64_t n, double stepsize, double* scalars_dev, int flags, void* shadow_bf16);

/* ---- Linear layer, FP32/FP64 SIMT mode (src/layers.jl:28-34 and its Zygote pullback) ------- */
/* Y = act(X W + b) */
int pg_linear_fwd(pg_ctx* ctx, int dtype, const void* X, const void* W, const void* b, void* Y,
                  int64_t M, int64_t K, int64_t N, int act);
/* dW = X^T dY ; db = colsum(dY) ; dX = (dY W^T) .* [relu_mask > 0]
 * dW/db/dX nullable (frozen layers / first layer); relu_mask nullable (= X when X came out of a relu). */
int pg_linear_bwd(pg_ctx* ctx, int dtype, const void* X, const void* W, const void* dY, void* dW, void* db, void* dX,
                  const void* relu_mask, int64_t M, int64_t K, int64_t N);

/* ---- Linear layer, BF16 tensor-core mode (tcgen05 + TMEM + TMA) --------------------------- */
/* X bf16 [M][K], W bf16 [K][N], b f32.  y_dtype PG_BF16 (hidden layers) or PG_F32 (logits, N <= 32). */
int pg_tc_linear_fwd(pg_ctx* ctx, const void* X, const void* W, const void* b, void* Y, int y_dtype,
                     int64_t M, int64_t K, int64_t N, int act);
/* dY bf16 (or f32 when N <= 32); dW/db f32 (straight into the gradient arena); dX bf16. */
int pg_tc_linear_bwd(pg_ctx* ctx, const void* X, const void* W, const void* dY, int dy_dtype, void* dW, void* db, void* dX,
                     const void* relu_mask, int64_t M, int64_t K, int64_t N);
/* same, plus dx_colsum (nullable, f32 [K]): column sums of the masked dX, i.e. the bias gradient of the
 * layer that produced X (a deterministic pass over the stored dX; the planner takes bias gradients from the weight-gradient
 * launches instead). */
int pg_tc_linear_bwd_ex(pg_ctx* ctx, const void* X, const void* W, const void* dY, int dy_dtype, void* dW, void* db, void* dX,
                        const void* relu_mask, void* dx_colsum, int64_t M, int64_t K, int64_t N);
/* raw GEMM entry for unit tests: C[M][N] (f32 or bf16) = A op B, bf16 operands, fp32 accumulate.
 * layout 0: A[M][K] B[K][N]   (forward)      1: A[M][K] B[N][K]^T (data gradient)
 * layout 2: A[K][M]^T B[K][N] (weight gradient) */
int pg_tc_gemm(pg_ctx* ctx, int layout, const void* A, const void* B, void* C, int c_dtype, int64_t M, int64_t N, int64_t K);

/* ---- Conv / pooling / activations (src/layers.jl:59,81,93,99; NNlib semantics) ------------- */
int pg_conv2d_fwd(pg_ctx* ctx, int dtype, const void* x, const void* w, const void* b, void* y,
                  int64_t N, int Cin, int H, int W, int Cout, int kH, int kW, int act);
int pg_conv2d_bwd(pg_ctx* ctx, int dtype, const void* x, const void* w, const void* dy, void* dw, void* db, void* dx,
                  int64_t N, int Cin, int H, int W, int Cout, int kH, int kW);
/* Conv layer, BF16 tensor-core mode (tcgen05 implicit GEMM, csrc/tc_conv.cu): same tensors as pg_conv2d_* in
 * Float32 NCHW; operands are rounded to bf16 on the way into shared memory, products accumulate in fp32.
 * Envelope (pg_tc_conv2d_supported returns 1): Cin, Cout <= 16, Cin*kH*kW < 256, Cout*kH*kW <= 448 and one
 * sample (plus its output gradient / zero-padded copy) fits a shared-memory slot; other shapes use pg_conv2d_*.
 * The weight-gradient launch writes one partial dw / db per CTA; a small kernel sums them in CTA order (deterministic). */
int pg_tc_conv2d_supported(int Cin, int H, int W, int Cout, int kH, int kW);
int pg_tc_conv2d_fwd(pg_ctx* ctx, const void* x, const void* w, const void* b, void* y,
                     int64_t N, int Cin, int H, int W, int Cout, int kH, int kW, int act);
int pg_tc_conv2d_bwd(pg_ctx* ctx, const void* x, const void* w, const void* dy, void* dw, void* db, void* dx,
                     int64_t N, int Cin, int H, int W, int Cout, int kH, int kW);
/* Conv + bias + ReLU + 2x2 max-pool in ONE kernel each way (csrc/conv_fused.cu; examples/02_mnist_conv.jl:27-33 chains exactly
 * these three): Float32 NCHW, fp32 math.  y_pool [N][Cout][H'/2][W'/2]; argmax (nullable for inference) one byte per pooled
 * element: which of the window's four positions won (W fastest, first maximum).  The un-pooled conv output never touches HBM.
 * The backward rebuilds the conv-output gradient from (dy_pool, argmax, y_pool > 0) in shared memory; dw / db are summed
 * from per-CTA partials in a fixed order (no floating-point atomics).  Envelope: pg_conv2d_relu_pool2_supported == 1. */
int pg_conv2d_relu_pool2_supported(int Cin, int H, int W, int Cout, int kH, int kW);
int pg_conv2d_relu_pool2_fwd(pg_ctx* ctx, const void* x, const void* w, const void* b, void* y_pool, void* argmax,
                             int64_t N, int Cin, int H, int W, int Cout, int kH, int kW);
int pg_conv2d_relu_pool2_bwd(pg_ctx* ctx, const void* x, const void* w, const void* y_pool, const void* argmax, const void* dy_pool,
                             void* dw, void* db, void* dx, int64_t N, int Cin, int H, int W, int Cout, int kH, int kW);
int pg_maxpool2d_fwd(pg_ctx* ctx, int dtype, const void* x, void* y, int64_t NC, int H, int W, int k);
int pg_maxpool2d_bwd(pg_ctx* ctx, int dtype, const void* x, const void* dy, void* dx, int64_t NC, int H, int W, int k);
/* same, fused with relu' of the pool input (x = a ReLU output): dx is the gradient w.r.t. the pre-activation */
int pg_maxpool2d_relu_bwd(pg_ctx* ctx, int dtype, const void* x, const void* dy, void* dx, int64_t NC, int H, int W, int k, int relu);
int pg_meanpool2d_fwd(pg_ctx* ctx, int dtype, const void* x, void* y, int64_t NC, int H, int W, int k);
int pg_meanpool2d_bwd(pg_ctx* ctx, int dtype, const void* dy, void* dx, int64_t NC, int H, int W, int k);
/* dropout in training mode (layers.jl:109-138): y = x .* mask, mask_i = rand_i > p ? T(1/(1-p)) : 0.
 * The mask is a pure function of (seed, stream_id, elem_offset + i) (Philox4x32-10): call it again on dY
 * with the same arguments for the pullback.  elem_offset: global index of x[0] (data-parallel shards). */
int pg_dropout(pg_ctx* ctx, int dtype, const void* x, void* y, int64_t n, double p, uint64_t seed, uint32_t stream_id, uint64_t elem_offset);
int pg_relu_fwd(pg_ctx* ctx, int dtype, const void* x, void* y, int64_t n);
int pg_relu_bwd(pg_ctx* ctx, int dtype, const void* y, const void* dy, void* dx, int64_t n);

/* ---- softmax + losses (src/layers.jl:87, src/losses.jl:5-10) ------------------------------ */
/* fused: P = softmax(Z); loss = inv_batch * sum_b( -sum_c Y log(P + eps(T)) ); dZ = dL/dZ.
 * P, dZ nullable.  loss_dev: one device scalar of type T.  inv_batch = 1/B_global. */
int pg_softmax_ce_fwd_bwd(pg_ctx* ctx, int dtype, const void* Z, const void* Y, void* P, void* dZ, void* loss_dev,
                          int64_t B, int C, double inv_batch);
/* same with integer class labels (one-hot implied) */
int pg_softmax_ce_idx_fwd_bwd(pg_ctx* ctx, int dtype, const void* Z, const int32_t* labels, void* P, void* dZ, void* loss_dev,
                              int64_t B, int C, double inv_batch);
int pg_softmax_fwd(pg_ctx* ctx, int dtype, const void* Z, void* P, int64_t B, int C);
int pg_softmax_bwd(pg_ctx* ctx, int dtype, const void* P, const void* dP, void* dZ, int64_t B, int C);
/* cross_entropy on probabilities (losses.jl:7-10) and its gradient dP = -Y * inv_batch / (P + eps) */
int pg_cross_entropy_fwd_bwd(pg_ctx* ctx, int dtype, const void* P, const void* Y, void* dP, void* loss_dev,
                             int64_t B, int C, double inv_batch);
/* mse (losses.jl:5): loss = inv_denom * sum (yhat - y)^2 ; dY = 2 inv_denom (yhat - y) */
int pg_mse_fwd_bwd(pg_ctx* ctx, int dtype, const void* Yhat, const void* Y, void* dY, void* loss_dev, int64_t n, double inv_denom);
/* class_error (losses.jl:12-16): fraction of rows whose argmax differs; synchronising */
int pg_class_error(pg_ctx* ctx, int dtype, const void* P, const void* Y, int64_t B, int C, double* result);
/* ---- data step next to the path (examples/01_mnist_mlp.jl:17,35-38) -------------------- */
/* dst[r][:] = src[idx[r]][:]  — on-device minibatch gather from a dataset resident in HBM */
int pg_gather_rows(pg_ctx* ctx, int dtype, const void* src, const int32_t* idx, void* dst, int64_t rows, int64_t row_elems);
/* dense one-hot labels [B][C] from integer classes (Flux.onehotbatch) */
int pg_onehot(pg_ctx* ctx, int dtype, const int32_t* labels, void* out, int64_t B, int C);
/* read a device loss scalar of type T (the step's only host sync) */
int pg_loss_read(pg_ctx* ctx, int dtype, const void* loss_dev, double* result);


/* training-mode dropout on padded / mixed-type tensors: logical [rows][cols] elements, row pitches in elements;
 * in_dtype / out_dtype in {PG_F32, PG_F64, PG_BF16} (a cast, a zero-padding copy and the mask in one pass; p = 0
 * makes it a plain cast / pad copy).  Element (r, c) draws the same number pg_dropout draws for index r*cols + c. */
int pg_dropout_ex(pg_ctx* ctx, int in_dtype, int out_dtype, const void* x, void* y, int64_t rows, int64_t cols, int64_t in_pitch, int64_t out_pitch,
                  double p, uint64_t seed, uint32_t stream_id, uint64_t elem_offset);
/* pg_tc_linear_bwd_ex with a selectable data-gradient type: dx_dtype PG_BF16 (feeds another tensor-core Linear) or
 * PG_F32 (feeds a pooling / conv / fp32 layer). */
int pg_tc_linear_bwd2(pg_ctx* ctx, const void* X, const void* W, const void* dY, int dy_dtype, void* dW, void* db, void* dX, int dx_dtype,
                      const void* relu_mask, int64_t M, int64_t K, int64_t N);

/* ---- parameter arenas: all leaves of a Model in one device block in vec(m) order (src/model.jl:9-24,33) -------------
 * Each leaf start is padded to 32 elements; padding is zero and invisible to length / flat indexing / dot.
 * `leaf` arguments: 0-based leaf index, or -1 = the whole arena in vec(m) order (padding skipped). */
int pg_arena_create(pg_ctx* ctx, int dtype, int n_leaves, const int64_t* leaf_sizes, pg_arena** arena);
/* a non-owning arena handle over device memory the host already laid out with the same rule (leaf starts padded to 32 elements) */
int pg_arena_wrap(pg_ctx* ctx, int dtype, int n_leaves, const int64_t* leaf_sizes, void* base, pg_arena** arena);
int pg_arena_destroy(pg_arena* arena);
/* mode 0 = similar (uninitialised -> zero-filled), 1 = zero, 2 = copy          (model.jl:49-56) */
int pg_arena_like(pg_arena* src, int mode, pg_arena** out);
int pg_arena_info(pg_arena* arena, int* dtype, int* n_leaves, int64_t* length, int64_t* n_padded, void** base);
int pg_arena_leaf_ptr(pg_arena* arena, int leaf, void** dptr, int64_t* size);
int pg_arena_upload(pg_arena* arena, int leaf, const void* host);             /* synchronising */
int pg_arena_download(pg_arena* arena, int leaf, void* host);                 /* synchronising: vec(m), model.jl:33 */
int pg_arena_get_scalar(pg_arena* arena, int64_t flat_index, double* value);  /* m[i], model.jl:79-88 (0-based here) */
int pg_arena_set_scalar(pg_arena* arena, int64_t flat_index, double value);   /* m[i] = v, model.jl:90-99 */

/* ---- generic Model broadcast (src/model.jl:148-164): out .= f(args...) for ANY flattened scalar function f --------
 * `code` is a postfix program evaluated once per element on a small register stack.  Every value on the stack
 * carries Julia's promoted type (the leaf eltype T or Float64): an op on two T values
 * rounds to T, an op involving a Float64 value is done in Float64; the final value is rounded to T on store.
 *   instruction word = opcode | (operand << 8)
 *   PG_BC_ARG k      push arena operand k (type T)          PG_BC_CONST k   push consts[k] as Float64
 *   PG_BC_CONST_T k  push consts[k] rounded to T (an Int / T-typed Julia scalar)    PG_BC_CONST_F32 k  push it as a Float32 scalar
 *   binary: ADD SUB MUL DIV MAX MIN POW      unary: NEG SQRT ABS EXP LOG SQUARE RELU INV
 * dtype in {PG_F16, PG_F32, PG_F64}.  n = padded arena length; padding stays zero when `zero_padding` gives the
 * arena's leaf layout (n_leaves, leaf_sizes) — pass n_leaves = 0 if f(0,...) == 0. */
enum { PG_BC_ARG = 1, PG_BC_CONST = 2, PG_BC_CONST_T = 3, PG_BC_CONST_F32 = 4, PG_BC_ADD = 10, PG_BC_SUB = 11, PG_BC_MUL = 12, PG_BC_DIV = 13, PG_BC_MAX = 14,
       PG_BC_MIN = 15, PG_BC_POW = 16, PG_BC_NEG = 30, PG_BC_SQRT = 31, PG_BC_ABS = 32, PG_BC_EXP = 33, PG_BC_LOG = 34, PG_BC_SQUARE = 35,
       PG_BC_RELU = 36, PG_BC_INV = 37 };
int pg_vec_expr(pg_ctx* ctx, int dtype, void* out, const void* const* args, int n_args, const int32_t* code, int n_code,
                const double* consts, int n_consts, int64_t n, int n_leaves, const int64_t* leaf_sizes);

/* ---- the training-step plan: eval_with_pullback behind the ABI (src/gradient.jl:1-13, ext/ProtoGradZygoteExt.jl:6-13) ----
 * The host language traces `f(m)` once (its layer calls record ops) and hands the recorded graph over; the plan fuses
 * it (Linear/Conv + bias + ReLU, softmax + cross-entropy across the model/loss boundary with dZ produced in the same
 * kernel, ReLU' into the upstream data-gradient epilogue), picks the kernel per layer (FP32/FP64 SIMT or bf16 tcgen05),
 * owns the activation / gradient buffers and launches everything.  Sequential graphs only (every intermediate value
 * has one consumer), which is every model `Compose` builds (src/layers.jl:63-75). */
typedef enum { PG_MATH_FP32 = 0, PG_MATH_BF16 = 1 } pg_math_mode;
typedef enum {
  PG_OP_INPUT = 0,      /* a tensor the closure captured: minibatch, labels (dtype PG_F32/PG_F64, PG_BF16 pre-cast inputs, PG_I32 class indices) */
  PG_OP_PARAM = 1,      /* a Model leaf; grad_leaf >= 0 when it belongs to the differentiated model */
  PG_OP_LINEAR = 2,     /* in = {x, W, b}                 layers.jl:28-34 */
  PG_OP_CONV2D = 3,     /* in = {x, w, b}                 layers.jl:59 */
  PG_OP_RELU = 4,       /* layers.jl:81 */
  PG_OP_MAXPOOL = 5,    /* window k                       layers.jl:93 */
  PG_OP_MEANPOOL = 6,   /* layers.jl:99 */
  PG_OP_RESHAPE = 7,    /* no-op on memory                layers.jl:105, examples/02_mnist_conv.jl:34 */
  PG_OP_DROPOUT = 8,    /* probability p, training mode   layers.jl:118-132 */
  PG_OP_SOFTMAX = 9,    /* over the last (class) axis     layers.jl:87 */
  PG_OP_LOSS_CE = 10,   /* in = {probs, labels}           losses.jl:7-10; denom = global batch (0: this batch) */
  PG_OP_LOSS_MSE = 11   /* in = {output, labels}          losses.jl:5;    denom = divisor (0: number of elements) */
} pg_op_kind;
typedef struct pg_op {
  int32_t kind;
  int32_t dtype;
  int32_t in[3];        /* node indices, -1 = unused */
  int32_t ndim;
  int64_t shape[4];     /* C order */
  int32_t grad_leaf;    /* PG_OP_PARAM: index of its gradient leaf in vec(m) order, -1 = frozen / captured model */
  int32_t k;
  double p;
  double denom;
} pg_op;
/* node flags reported by pg_plan_node_flags */
enum { PG_NODE_NEEDED = 1, PG_NODE_NEEDS_GRAD = 2, PG_NODE_TENSOR_CORE = 4, PG_NODE_FUSED_ACT = 8, PG_NODE_BF16_VALUE = 16, PG_NODE_FUSED_AWAY = 32,
       PG_NODE_FUSED_POOL = 64 /* a Conv whose ReLU and 2x2 max-pool run inside its kernel (pg_conv2d_relu_pool2_*) */ };
typedef void (*pg_ready_fn)(void* user, int first_grad_leaf);

/* ---- one training step of a small MLP in ONE cooperative launch (csrc/mlp_step.cu) -------------------------------------
 * Linear -> ReLU -> ... -> Linear -> softmax -> cross_entropy (the model of /root/reference/examples/01_mnist_mlp.jl:19-29
 * without dropout; forward src/layers.jl:28-34,81,87, loss src/losses.jl:7-10, pullbacks as pg_linear_bwd /
 * pg_softmax_ce_fwd_bwd) followed by the optimizer update over the parameter arena (gradient_descent.jl:15-17 or
 * adam.jl:32-41, same arithmetic and flags as pg_opt_gd / pg_opt_adam).  Float32; widths and batch <= 1536; gradients are
 * written to gW / gb (they stay readable), the loss to loss_out.  Replaces the eval_with_pullback -> pb() -> step! call
 * sequence for models where 13 dependent launches cost more than the arithmetic. */
typedef struct pg_mlp_desc {
  int32_t n_layers;          /* Linear layers, 2..4 */
  int32_t opt;               /* 0: no update, 1: GradientDescent, 2: Adam */
  int64_t batch;
  int64_t dims[5];           /* dims[0] = input width, dims[l+1] = output width of layer l */
  const void* x;             /* [batch][dims[0]] f32 */
  const void* labels;        /* [batch][dims[n_layers]] f32 (one-hot or soft) */
  void* W[4]; void* b[4];    /* W[l]: [dims[l]][dims[l+1]] (Julia's W[out,in] bytes), b[l]: [dims[l+1]] */
  void* gW[4]; void* gb[4];  /* gradients, same shapes */
  void* loss_out;            /* f32 device scalar */
  double loss_scale;         /* 1 / (global) batch */
  void* arena_p; void* arena_g; void* arena_m; void* arena_v; void* arena_m_hat; void* arena_v_hat;  /* m_hat, v_hat nullable */
  int64_t arena_n;           /* padded arena length (elements) */
  double stepsize, beta1, beta2, eps;
  int64_t it;                /* Adam iteration counter of THIS step (starts at 1, adam.jl:28) */
  int32_t flags;             /* PG_SCALAR_F64 | PG_MATH_FAST, as pg_opt_adam */
} pg_mlp_desc;
int pg_mlp_step(pg_ctx* ctx, const pg_mlp_desc* desc);

/* Host-only (callable without a GPU): buffers are allocated by the first pg_plan_forward. */
int pg_plan_create(pg_ctx* ctx, const pg_op* ops, int n_ops, int root, int math_mode, pg_plan** plan);
int pg_plan_destroy(pg_plan* plan);
int pg_plan_node_flags(pg_plan* plan, int node, int* flags);
/* leaf_ptrs[i]: device pointer of INPUT / PARAM node i (ignored for other nodes).  leaf_bf16[i] (array nullable, entries
 * nullable): a VALID bf16 copy of PARAM node i (e.g. written by the fused optimizer pass, pg_opt_* shadow_bf16); NULL lets
 * the plan cast the leaf itself.  (rng_seed, rng_step, sample_offset) key the dropout masks (sample_offset = global index
 * of this process's first sample, so data-parallel shards draw the full batch's masks).  *loss_dev: device scalar of
 * the model dtype (read it with pg_loss_read; slots rotate, a handle stays valid for 64 forwards).  *generation
 * identifies this forward: pg_plan_backward rejects any other (a stale pullback). */
int pg_plan_forward(pg_plan* plan, const void* const* leaf_ptrs, const void* const* leaf_bf16, uint64_t rng_seed, uint64_t rng_step,
                    int64_t sample_offset, void** loss_dev, int64_t* generation);
/* grad_leaf_ptrs[j]: where dL/d(leaf j) is WRITTEN (vec(m) order; normally leaves of one gradient arena).
 * on_ready(user, j) (nullable) fires once every gradient leaf >= j has been enqueued (layers finish last-to-first):
 * the data-parallel caller starts reducing that slice of the arena while the remaining backward kernels run. */
int pg_plan_backward(pg_plan* plan, int64_t generation, void* const* grad_leaf_ptrs, int n_grad_leaves, pg_ready_fn on_ready, void* user);
/* device pointer / dtype of a node's value after pg_plan_forward (inference: `m(test_x)`) */
int pg_plan_value(pg_plan* plan, int node, void** dptr, int* dtype, int64_t* pitch);

/* ---- data parallelism: one process per GPU; the only exchange of the path is the sum of the
 * flat gradient arena over ranks, followed by the identical fused update on every rank ---------------------------------
 * The communicator is this library's own (libnccl is dlopen'ed; PG_NCCL_LIB overrides the path): rank 0 creates an id,
 * the host distributes its 128 bytes (MPI / a file / torch.distributed's store), every rank calls pg_dp_init. */
int pg_dp_unique_id(void* id128);
int pg_dp_init(pg_ctx* ctx, int rank, int world, const void* id128, int max_ctas, pg_dp** dp);
int pg_dp_destroy(pg_dp* dp);
int pg_dp_info(pg_dp* dp, int* rank, int* world, int* nccl_version, int* p2p);
/* in-place sum over ranks of buf[0..count) on the communicator's stream, ordered after everything enqueued on the ctx
 * stream so far; pg_dp_wait makes the ctx stream wait for all outstanding collectives. */
int pg_dp_allreduce(pg_dp* dp, void* buf, int64_t count, int dtype);
int pg_dp_broadcast(pg_dp* dp, void* buf, int64_t count, int dtype, int root);
int pg_dp_wait(pg_dp* dp);
/* bucketed, overlapped reduction of a gradient arena during pg_plan_backward: pass pg_dp_bucket_ready / the bucket
 * handle as on_ready / user.  A bucket is issued once >= bucket_bytes of finished leaves have accumulated. */
int pg_dp_bucket_begin(pg_dp* dp, pg_arena* grads, int64_t bucket_bytes);
void pg_dp_bucket_ready(void* dp, int first_grad_leaf);
/* number of issued buckets; ranges as (start element, count) of the padded arena, in issue order (tail first) */
int pg_dp_bucket_finish(pg_dp* dp, int* n_buckets, int64_t* starts, int64_t* counts, int max_buckets);
int pg_dp_bucket_wait(pg_dp* dp, int bucket);   /* ctx stream waits for bucket k's reduction */
/* Fused gradient exchange + optimizer update over NVLink peer memory (CUDA IPC; csrc/dp.cu): each rank owns 1/world of
 * every bucket; ONE kernel per bucket reads that slice of the gradient arena from all peers (P2P loads, summed in rank
 * order), applies the optimizer to the slice (touching 1/world of the optimizer state) and stores the new fp32 parameters
 * and their bf16 copy into every replica (P2P stores).  It replaces ncclAllReduce + the full-arena pg_opt_* pass of a
 * data-parallel step.  shadow2: bf16 [2][n_padded] double buffer (nullable): the step in flight reads one half, the
 * kernels write the other.  PG_ERR_UNSUPPORTED (peer memory cannot be mapped) leaves the all-reduce path in use. */
int pg_dp_p2p_enable(pg_dp* dp, pg_arena* params, pg_arena* grads, void* shadow2);
/* the optimizer the NEXT bucketed backward applies: kind 0 = Adam (m, v: state arenas; `it` before the increment; flags as
 * pg_opt_adam), 2 = GradientDescent, -1 = off (buckets are plain all-reduces again) */
int pg_dp_p2p_set_optimizer(pg_dp* dp, int kind, void* m, void* v, double stepsize, double beta1, double beta2, double epsilon, int64_t it, int flags);
int pg_dp_p2p_shadow_half(pg_dp* dp, int* half);
/* tuning knobs of the fused exchange: "p2p_blocks" (CTAs of a bucket that overlaps the backward pass), "p2p_blocks_last" (CTAs of
 * the final bucket), "reserve_sms" (SMs the persistent GEMM grids leave free while a bucket is in flight), "p2p_tma" (1: bulk-copy
 * pipeline, 0: register-resident loads), "timing", "debug" */
int pg_dp_set_option(pg_dp* dp, const char* name, int value);
/* the bucket policy as a pure host function (no GPU): slices issued for a sequence of on_ready leaf indices */
int pg_dp_bucket_plan(int dtype, int n_leaves, const int64_t* leaf_sizes, int64_t bucket_bytes, const int* ready_leaves, int n_ready,
                      int64_t* starts, int64_t* counts, int max_buckets, int* n_buckets);

#ifdef __cplusplus
}
#endif
#endif /* PROTOGRAD_B200_H */
