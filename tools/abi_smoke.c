/* abi_smoke.c — a plain-C host of the C ABI (SURVEY.md §8b lists "a C smoke binary" among the callers): no Python, no
 * ctypes.  It does what julia/ProtoGradB200.jl does behind the reference's seams for the model of
 * /root/reference/examples/01_mnist_mlp.jl:19-29 (784-100-10, minibatch 128, dropout off):
 *
 *   arenas (vec(m) order, src/model.jl:9-24,33)        pg_arena_create / _like / _upload / _leaf_ptr / _download
 *   eval_with_pullback (src/gradient.jl:1-13)           pg_plan_create / pg_plan_forward / pg_loss_read / pg_plan_backward
 *   step!(::AdamState, grad) (src/optimizers/adam.jl)   pg_opt_adam
 *   dot(m, m) (src/model.jl:200-202)                    pg_vec_dot
 *
 * and checks the first step against a host-side double-precision evaluation of the same forward pass
 * (src/layers.jl:28-34 W*x .+ b, :81 relu, :87 softmax; src/losses.jl:7-10 cross_entropy with eps(Float32) inside the log)
 * and of the last layer's bias gradient (mean over the batch of p - y, up to the eps term), then trains 30 Adam steps on
 * the fixed minibatch and requires the loss to fall.  Exit code 0 and "abi_smoke ok" on success.
 *
 *   gcc -O2 -I include tools/abi_smoke.c -L protograd.jl_b200 -lprotograd_b200 -lm -o abi_smoke
 *   LD_LIBRARY_PATH=protograd.jl_b200 ./abi_smoke
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "protograd_b200.h"

#define CHECK(call)                                                                     \
  do {                                                                                  \
    int rc_ = (call);                                                                   \
    if (rc_ != PG_OK) {                                                                 \
      fprintf(stderr, "%s:%d: %s -> %d: %s\n", __FILE__, __LINE__, #call, rc_, pg_last_error()); \
      return 1;                                                                         \
    }                                                                                   \
  } while (0)

enum { B = 128, D0 = 784, D1 = 100, D2 = 10 };

static unsigned long long lcg_state = 0x9E3779B97F4A7C15ull;
static double uniform01(void) { /* 53 random bits */
  lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull;
  return (double)(lcg_state >> 11) / 9007199254740992.0;
}

static pg_op make_op(int kind, int dtype, int in0, int in1, int in2, int ndim, long long s0, long long s1, int grad_leaf) {
  pg_op o;
  memset(&o, 0, sizeof o);
  o.kind = kind; o.dtype = dtype;
  o.in[0] = in0; o.in[1] = in1; o.in[2] = in2;
  o.ndim = ndim; o.shape[0] = s0; o.shape[1] = s1;
  o.grad_leaf = grad_leaf;
  return o;
}

int main(void) {
  int ndev = 0;
  CHECK(pg_device_count(&ndev));
  if (ndev < 1) { fprintf(stderr, "no CUDA device\n"); return 2; }
  pg_ctx* ctx = NULL;
  CHECK(pg_init(0, &ctx));

  /* ---- parameters in vec(m) order: W1 [in][out] (Julia's W[out,in] bytes), b1, W2, b2; xavier_uniform (src/layers.jl:6-11) ---- */
  const int64_t sizes[4] = {(int64_t)D0 * D1, D1, (int64_t)D1 * D2, D2};
  const int64_t P = sizes[0] + sizes[1] + sizes[2] + sizes[3];
  float* theta = (float*)malloc((size_t)P * sizeof(float));
  float* W1 = theta, *b1 = W1 + sizes[0], *W2 = b1 + sizes[1], *b2 = W2 + sizes[2];
  const double s1 = sqrt(6.0 / (D0 + D1)), s2 = sqrt(6.0 / (D1 + D2));
  for (int64_t i = 0; i < sizes[0]; ++i) W1[i] = (float)(s1 * (2.0 * uniform01() - 1.0));
  for (int64_t i = 0; i < sizes[2]; ++i) W2[i] = (float)(s2 * (2.0 * uniform01() - 1.0));
  for (int i = 0; i < D1; ++i) b1[i] = 0.01f * (float)(i % 7 - 3);
  for (int i = 0; i < D2; ++i) b2[i] = 0.02f * (float)(i % 5 - 2);

  pg_arena *params = NULL, *grads = NULL, *m = NULL, *v = NULL;
  CHECK(pg_arena_create(ctx, PG_F32, 4, sizes, &params));
  CHECK(pg_arena_upload(params, -1, theta));
  CHECK(pg_arena_like(params, 1, &grads));
  CHECK(pg_arena_like(params, 1, &m));
  CHECK(pg_arena_like(params, 1, &v));
  int64_t length = 0, n_padded = 0;
  void *p_base = NULL, *g_base = NULL, *m_base = NULL, *v_base = NULL;
  CHECK(pg_arena_info(params, NULL, NULL, &length, &n_padded, &p_base));
  CHECK(pg_arena_info(grads, NULL, NULL, NULL, NULL, &g_base));
  CHECK(pg_arena_info(m, NULL, NULL, NULL, NULL, &m_base));
  CHECK(pg_arena_info(v, NULL, NULL, NULL, NULL, &v_base));
  if (length != P) { fprintf(stderr, "arena length %lld != %lld\n", (long long)length, (long long)P); return 1; }

  /* ---- a synthetic minibatch: pixels in [0, 1), one-hot labels ---- */
  float* x = (float*)malloc((size_t)B * D0 * sizeof(float));
  float* y = (float*)calloc((size_t)B * D2, sizeof(float));
  for (int i = 0; i < B * D0; ++i) x[i] = (float)uniform01();
  for (int r = 0; r < B; ++r) y[r * D2 + (int)(uniform01() * D2) % D2] = 1.0f;
  void *x_dev = NULL, *y_dev = NULL;
  CHECK(pg_malloc(ctx, (size_t)B * D0 * sizeof(float), &x_dev));
  CHECK(pg_malloc(ctx, (size_t)B * D2 * sizeof(float), &y_dev));
  CHECK(pg_upload(ctx, x_dev, x, (size_t)B * D0 * sizeof(float)));
  CHECK(pg_upload(ctx, y_dev, y, (size_t)B * D2 * sizeof(float)));
  CHECK(pg_sync(ctx));

  /* ---- the traced objective m -> cross_entropy(m(x), y), as the glue records it ---- */
  enum { N_X, N_Y, N_W1, N_B1, N_L1, N_R1, N_W2, N_B2, N_L2, N_SM, N_LOSS, N_OPS };
  pg_op ops[N_OPS];
  ops[N_X] = make_op(PG_OP_INPUT, PG_F32, -1, -1, -1, 2, B, D0, -1);
  ops[N_Y] = make_op(PG_OP_INPUT, PG_F32, -1, -1, -1, 2, B, D2, -1);
  ops[N_W1] = make_op(PG_OP_PARAM, PG_F32, -1, -1, -1, 2, D0, D1, 0);
  ops[N_B1] = make_op(PG_OP_PARAM, PG_F32, -1, -1, -1, 1, D1, 0, 1);
  ops[N_L1] = make_op(PG_OP_LINEAR, PG_F32, N_X, N_W1, N_B1, 2, B, D1, -1);
  ops[N_R1] = make_op(PG_OP_RELU, PG_F32, N_L1, -1, -1, 2, B, D1, -1);
  ops[N_W2] = make_op(PG_OP_PARAM, PG_F32, -1, -1, -1, 2, D1, D2, 2);
  ops[N_B2] = make_op(PG_OP_PARAM, PG_F32, -1, -1, -1, 1, D2, 0, 3);
  ops[N_L2] = make_op(PG_OP_LINEAR, PG_F32, N_R1, N_W2, N_B2, 2, B, D2, -1);
  ops[N_SM] = make_op(PG_OP_SOFTMAX, PG_F32, N_L2, -1, -1, 2, B, D2, -1);
  ops[N_LOSS] = make_op(PG_OP_LOSS_CE, PG_F32, N_SM, N_Y, -1, 0, 0, 0, -1);
  pg_plan* plan = NULL;
  CHECK(pg_plan_create(ctx, ops, N_OPS, N_LOSS, PG_MATH_FP32, &plan));

  const void* leaf_ptrs[N_OPS];
  memset(leaf_ptrs, 0, sizeof leaf_ptrs);
  leaf_ptrs[N_X] = x_dev; leaf_ptrs[N_Y] = y_dev;
  const int param_nodes[4] = {N_W1, N_B1, N_W2, N_B2};
  void* grad_ptrs[4];
  for (int l = 0; l < 4; ++l) {
    void* p = NULL;
    CHECK(pg_arena_leaf_ptr(params, l, &p, NULL));
    leaf_ptrs[param_nodes[l]] = p;
    CHECK(pg_arena_leaf_ptr(grads, l, &grad_ptrs[l], NULL));
  }

  /* ---- host-side double evaluation of the forward pass and of db2 ---- */
  double ref_loss = 0.0, ref_db2[D2];
  memset(ref_db2, 0, sizeof ref_db2);
  const double eps32 = 1.1920928955078125e-07;        /* eps(Float32), src/losses.jl:9 */
  for (int r = 0; r < B; ++r) {
    double h[D1], z[D2], zmax = -1e300, den = 0.0;
    for (int j = 0; j < D1; ++j) {
      double acc = b1[j];
      for (int k = 0; k < D0; ++k) acc += (double)x[r * D0 + k] * (double)W1[k * D1 + j];
      h[j] = acc > 0.0 ? acc : 0.0;
    }
    for (int c = 0; c < D2; ++c) {
      double acc = b2[c];
      for (int j = 0; j < D1; ++j) acc += h[j] * (double)W2[j * D2 + c];
      z[c] = acc;
      if (acc > zmax) zmax = acc;
    }
    for (int c = 0; c < D2; ++c) den += exp(z[c] - zmax);
    double dot_dp_p = 0.0, p[D2], dp[D2];
    for (int c = 0; c < D2; ++c) {
      p[c] = exp(z[c] - zmax) / den;
      ref_loss -= (double)y[r * D2 + c] * log(p[c] + eps32);
      dp[c] = -(double)y[r * D2 + c] / ((p[c] + eps32) * B);      /* d mean(-sum y log(p + eps)) / dp */
      dot_dp_p += dp[c] * p[c];
    }
    for (int c = 0; c < D2; ++c) ref_db2[c] += p[c] * (dp[c] - dot_dp_p);   /* softmax pullback; db2 = column sum of dZ */
  }
  ref_loss /= B;

  /* ---- step 1: eval_with_pullback, pb(), checks ---- */
  void* loss_dev = NULL;
  int64_t gen = 0;
  double loss0 = 0.0;
  CHECK(pg_plan_forward(plan, leaf_ptrs, NULL, 0, 0, 0, &loss_dev, &gen));
  CHECK(pg_loss_read(ctx, PG_F32, loss_dev, &loss0));
  CHECK(pg_plan_backward(plan, gen, grad_ptrs, 4, NULL, NULL));
  float db2[D2];
  CHECK(pg_arena_download(grads, 3, db2));
  double num = 0.0, den = 0.0;
  for (int c = 0; c < D2; ++c) { num += (db2[c] - ref_db2[c]) * (db2[c] - ref_db2[c]); den += ref_db2[c] * ref_db2[c]; }
  const double loss_rel = fabs(loss0 - ref_loss) / fabs(ref_loss), db_rel = sqrt(num / den);
  printf("step 1: loss %.7f (host double %.7f, rel %.2e), db2 rel err %.2e\n", loss0, ref_loss, loss_rel, db_rel);
  if (!(loss_rel <= 1e-5) || !(db_rel <= 1e-5)) { fprintf(stderr, "parity check failed\n"); return 1; }
  /* a stale pullback must be rejected (generation check), not silently run on newer activations */
  if (pg_plan_backward(plan, gen + 7, grad_ptrs, 4, NULL, NULL) == PG_OK) { fprintf(stderr, "stale pullback accepted\n"); return 1; }

  /* dot(grad, grad) over the arena == the sum over the downloaded vec (padding excluded) */
  float* gvec = (float*)malloc((size_t)P * sizeof(float));
  CHECK(pg_arena_download(grads, -1, gvec));
  double gg_host = 0.0, gg_dev = 0.0;
  for (int64_t i = 0; i < P; ++i) gg_host += (double)gvec[i] * (double)gvec[i];
  CHECK(pg_vec_dot(ctx, PG_F32, g_base, g_base, n_padded, &gg_dev));
  if (!(fabs(gg_dev - gg_host) <= 1e-5 * gg_host)) { fprintf(stderr, "dot: %g vs %g\n", gg_dev, gg_host); return 1; }

  /* ---- 30 Adam steps on the fixed minibatch (adam.jl:32-41; `it` starts at 1) ---- */
  double loss = loss0;
  for (int64_t it = 1; it <= 30; ++it) {
    if (it > 1) {
      CHECK(pg_plan_forward(plan, leaf_ptrs, NULL, 0, (uint64_t)it, 0, &loss_dev, &gen));
      CHECK(pg_plan_backward(plan, gen, grad_ptrs, 4, NULL, NULL));
    }
    CHECK(pg_opt_adam(ctx, PG_F32, p_base, m_base, v_base, NULL, NULL, g_base, n_padded, 1e-3, 0.9, 0.999, 1e-8, it, PG_SCALAR_F64, NULL));
  }
  CHECK(pg_plan_forward(plan, leaf_ptrs, NULL, 0, 31, 0, &loss_dev, &gen));
  CHECK(pg_loss_read(ctx, PG_F32, loss_dev, &loss));
  int64_t launches = 0;
  CHECK(pg_launch_count(ctx, &launches));
  printf("after 30 Adam steps: loss %.6f (from %.6f), %lld kernel launches\n", loss, loss0, (long long)launches);
  if (!(loss < 0.8 * loss0) || launches <= 0) { fprintf(stderr, "training did not reduce the loss\n"); return 1; }

  CHECK(pg_plan_destroy(plan));
  CHECK(pg_arena_destroy(params)); CHECK(pg_arena_destroy(grads)); CHECK(pg_arena_destroy(m)); CHECK(pg_arena_destroy(v));
  CHECK(pg_free(ctx, x_dev)); CHECK(pg_free(ctx, y_dev));
  CHECK(pg_destroy(ctx));
  free(theta); free(x); free(y); free(gvec);
  printf("abi_smoke ok\n");
  return 0;
}
