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

typedef enum { PG_F32 = 0, PG_F64 = 1, PG_BF16 = 2 } pg_dtype;
typedef enum { PG_ACT_NONE = 0, PG_ACT_RELU = 1 } pg_act;
typedef enum { PG_OK = 0, PG_ERR_ARG = -1, PG_ERR_CUDA = -2, PG_ERR_UNSUPPORTED = -3, PG_ERR_NOMEM = -4 } pg_status;
typedef enum { PG_ADD = 0, PG_SUB = 1 } pg_binop;
/* model (op) scalar forms of src/model.jl:168-196 */
typedef enum { PG_X_PLUS_C = 0, PG_X_MINUS_C = 1, PG_C_MINUS_X = 2, PG_X_TIMES_C = 3, PG_X_OVER_C = 4, PG_C_OVER_X = 5 } pg_scalarop;
enum { PG_SCALAR_F64 = 1 };

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
int pg_delay_on(pg_ctx* ctx, void* stream, long long ns);   /* enqueue a one-thread kernel that sleeps ns (<= 1 ms) on `stream`: orders a collective behind the GEMM that becomes runnable at the same moment (protograd dp.py) */

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

/* ---- Optimizers: one fused pass over the arena (src/optimizers/*.jl) ---------------------- */
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
 * layer that produced X — accumulated in the data-gradient epilogue instead of a separate pass over dX. */
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
 * The weight-gradient launch adds per-CTA partial sums with fp32 atomics (dw / db are zeroed first). */
int pg_tc_conv2d_supported(int Cin, int H, int W, int Cout, int kH, int kW);
int pg_tc_conv2d_fwd(pg_ctx* ctx, const void* x, const void* w, const void* b, void* y,
                     int64_t N, int Cin, int H, int W, int Cout, int kH, int kW, int act);
int pg_tc_conv2d_bwd(pg_ctx* ctx, const void* x, const void* w, const void* dy, void* dw, void* db, void* dx,
                     int64_t N, int Cin, int H, int W, int Cout, int kH, int kW);
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

#ifdef __cplusplus
}
#endif
#endif /* PROTOGRAD_B200_H */
