// plan.cu — the training-step plan behind the C ABI: eval_with_pullback's forward and zero-argument pullback
// (src/gradient.jl:1-13, ext/ProtoGradZygoteExt.jl:6-13) for a traced graph of ProtoGrad layer ops.
//
// The reference differentiates an arbitrary closure `f(m)` with Zygote.  Here the host language traces `f(m)` once
// (Linear / Conv / relu / maxpool / reshape / dropout / softmax / cross_entropy / mse calls record nodes) and hands
// the node list to pg_plan_create.  The plan
//   * fuses: Linear/Conv + bias + ReLU in one kernel; softmax + cross-entropy across the model/loss boundary with
//     dZ produced by the same kernel; ReLU' into the upstream data-gradient epilogue (or the max-pool routing);
//   * selects the kernel per layer: FP32/FP64 SIMT (rel 1e-5 contract) or, in PG_MATH_BF16 mode, the tcgen05 GEMM /
//     implicit-GEMM conv kernels.  Widths that are not multiples of 8 (the reference's own examples: 100, 120, 84 —
//     examples/01_mnist_mlp.jl:20, 02_mnist_conv.jl:36-42) run on tensor cores through zero-padded bf16 operand
//     copies (16-byte TMA pitch); a Linear fed by an fp32 tensor (pool output) casts its input and returns an fp32
//     data gradient;
//   * owns activation / gradient buffers (allocated by the first forward, reused by every later step);
//   * keys dropout masks on (seed, step, node, global element index) and regenerates them in the pullback.
// Host-side logic only: every kernel is reached through the per-op C-ABI entry points of this library.
#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace {

constexpr int LOSS_RING = 64;

inline int64_t ceil8(int64_t v) { return (v + 7) / 8 * 8; }
inline size_t dsize(int dt) { return dt == PG_F64 ? 8 : (dt == PG_BF16 || dt == PG_F16) ? 2 : 4; }

struct Buf {
  void* p = nullptr;
  int dtype = PG_F32;
  int64_t rows = 0, cols = 0, pitch = 0;   // logical [rows][cols] with `pitch` elements per row (pad columns are zero)
  int64_t elems() const { return rows * pitch; }
};

struct Lin {                 // tensor-core Linear node: padded geometry and staging buffers
  int64_t M = 0, K = 0, N = 0, Kp = 0, Np = 0;
  bool wide = false;
  void* xbf = nullptr;       // bf16 [M][Kp] copy of an fp32 (or unpadded) input
  void* wbf = nullptr;       // bf16 [Kp][Np] copy of W when it needs padding or no valid shadow was passed
  void* bpad = nullptr;      // f32 [Np] zero-padded bias
  void* dWp = nullptr;       // f32 [Kp][Np] padded weight gradient
  void* dbp = nullptr;       // f32 [Np]
  void* dxf = nullptr;       // f32 [M][Kp] padded data gradient for an fp32 upstream
  const void* x_used = nullptr;   // operand pointers of the current forward (reused by the pullback)
  const void* w_used = nullptr;
};

}  // namespace

struct pg_plan {
  pg_ctx* ctx = nullptr;
  int math = PG_MATH_FP32, root = 0, n = 0, dtype = PG_F32;
  std::vector<pg_op> ops;
  std::vector<int64_t> size;
  std::vector<char> need, ng, sce_softmax, tc_lin, tc_conv;
  std::vector<std::vector<int>> consumers;
  std::vector<int> fused_act, fused_into, fused_sce;
  std::vector<int> fused_pool, pool_of;      // conv id -> the 2x2 max-pool fused into it (conv_fused.cu), and the inverse
  std::vector<void*> amax;                    // per fused pool node: winner byte per pooled element
  std::vector<Buf> vals, grads;
  std::vector<int> vdt;                 // value dtype per node
  std::vector<int64_t> vpitch;          // last-axis pitch of the value
  std::map<int, Lin> lin;
  std::map<int, void*> onehot;          // CE nodes with integer labels that are not fused with a softmax
  std::vector<void*> owned;
  std::vector<const void*> leaf, leaf_bf;
  void* loss_ring = nullptr;
  bool allocated = false;
  int64_t gen = 0;
  uint64_t rng_seed = 0, rng_step = 0;
  int64_t sample_offset = 0;

  // ---- graph helpers ---------------------------------------------------------------------------------------
  int64_t last(int i) const { return ops[i].ndim ? ops[i].shape[ops[i].ndim - 1] : 1; }
  int64_t rows(int i) const { return last(i) ? size[i] / last(i) : 0; }
  bool is(int i, int kind) const { return ops[i].kind == kind; }
  int down(int i) const {               // first consumer that is not a reshape / dropout (-1: none)
    int c = consumers[i].empty() ? -1 : consumers[i][0];
    while (c >= 0 && (is(c, PG_OP_RESHAPE) || is(c, PG_OP_DROPOUT))) c = consumers[c].empty() ? -1 : consumers[c][0];
    return c;
  }
  int down_reshape(int i) const {       // first consumer that is not a reshape
    int c = consumers[i].empty() ? -1 : consumers[i][0];
    while (c >= 0 && is(c, PG_OP_RESHAPE)) c = consumers[c].empty() ? -1 : consumers[c][0];
    return c;
  }
  int base(int i) const {               // node that owns the storage (reshape aliases its input)
    while (is(i, PG_OP_RESHAPE)) i = ops[i].in[0];
    return i;
  }
  int source(int i) const {             // producer behind reshapes and dropouts
    while (is(i, PG_OP_RESHAPE) || is(i, PG_OP_DROPOUT)) i = ops[i].in[0];
    return i;
  }
  // consumers that apply relu' of their input themselves: Linear (data-gradient epilogue mask) and maxpool (routes
  // the gradient only where the window max is > 0).  A dropout in between keeps that true: its output is zero
  // wherever the ReLU output is, so it serves as the mask (see backward()).
  bool premasked(int r) const {
    int c = down(r);
    return c >= 0 && (is(c, PG_OP_LINEAR) || is(c, PG_OP_MAXPOOL));
  }
  bool relu_upstream(int x) const { return is(source(x), PG_OP_RELU); }

  int fail(int code, const char* msg) const { pg_set_error("plan: %s", msg); return code; }
};

namespace {

int analyse(pg_plan& P) {
  const int n = P.n;
  auto& ops = P.ops;
  P.size.assign(n, 1);
  for (int i = 0; i < n; ++i) {
    if (ops[i].ndim < 0 || ops[i].ndim > 4) return P.fail(PG_ERR_ARG, "node rank must be 0..4");
    for (int d = 0; d < ops[i].ndim; ++d) P.size[i] *= ops[i].shape[d];
    for (int k = 0; k < 3; ++k)
      if (ops[i].in[k] >= i) return P.fail(PG_ERR_ARG, "nodes must be in topological order");
  }
  // ---- shape / arity validation: a malformed node list from any host is an error here, never an out-of-bounds kernel ----
  auto bad = [&](int i, const char* what) {
    pg_set_error("plan: node %d (kind %d): %s", i, ops[i].kind, what);
    return PG_ERR_ARG;
  };
  for (int i = 0; i < n; ++i) {
    const pg_op& o = ops[i];
    if (o.kind < PG_OP_INPUT || o.kind > PG_OP_LOSS_MSE) return bad(i, "unknown op kind");
    for (int d = 0; d < o.ndim; ++d)
      if (o.shape[d] < 0) return bad(i, "negative extent");
    static const int arity[] = {0, 0, 3, 3, 1, 1, 1, 1, 1, 1, 2, 2};
    for (int k = 0; k < 3; ++k) {
      if (k < arity[o.kind] ? o.in[k] < 0 : o.in[k] >= 0) return bad(i, "wrong number of inputs for this op");
    }
    const pg_op* a = arity[o.kind] > 0 ? &ops[o.in[0]] : nullptr;
    const int64_t sa = a ? P.size[o.in[0]] : 0;
    switch (o.kind) {
      case PG_OP_LINEAR: {
        const pg_op& w = ops[o.in[1]];
        if (w.ndim != 2 || a->ndim < 1 || o.ndim < 1) return bad(i, "Linear needs x [..., in], W [in][out]");
        const int64_t K = w.shape[0], N = w.shape[1];
        if (K <= 0 || N <= 0 || a->shape[a->ndim - 1] != K) return bad(i, "DimensionMismatch: Linear input width != rows of W");
        if (P.size[o.in[2]] != N) return bad(i, "DimensionMismatch: Linear bias length != columns of W");
        if (o.shape[o.ndim - 1] != N || P.size[i] != sa / K * N) return bad(i, "Linear output shape must be [..., out]");
        break;
      }
      case PG_OP_CONV2D: {
        const pg_op& w = ops[o.in[1]];
        if (a->ndim != 4 || w.ndim != 4 || o.ndim != 4) return bad(i, "Conv needs x [N][Cin][H][W], w [Cout][Cin][kH][kW]");
        if (a->shape[1] != w.shape[1]) return bad(i, "DimensionMismatch: Conv input channels != filter channels");
        if (w.shape[2] < 1 || w.shape[3] < 1 || a->shape[2] < w.shape[2] || a->shape[3] < w.shape[3]) return bad(i, "Conv filter larger than the image");
        if (P.size[o.in[2]] != w.shape[0]) return bad(i, "DimensionMismatch: Conv bias length != output channels");
        if (o.shape[0] != a->shape[0] || o.shape[1] != w.shape[0] || o.shape[2] != a->shape[2] - w.shape[2] + 1 ||
            o.shape[3] != a->shape[3] - w.shape[3] + 1)
          return bad(i, "Conv output shape must be [N][Cout][H-kH+1][W-kW+1]");
        break;
      }
      case PG_OP_MAXPOOL: case PG_OP_MEANPOOL: {
        if (a->ndim != 4 || o.ndim != 4 || o.k < 1) return bad(i, "pooling needs x [N][C][H][W] and a window k >= 1");
        if (a->shape[2] < o.k || a->shape[3] < o.k) return bad(i, "pooling window larger than the image");
        if (o.shape[0] != a->shape[0] || o.shape[1] != a->shape[1] || o.shape[2] != (a->shape[2] - o.k) / o.k + 1 ||
            o.shape[3] != (a->shape[3] - o.k) / o.k + 1)
          return bad(i, "pooling output shape must be [N][C][(H-k)/k+1][(W-k)/k+1]");
        break;
      }
      case PG_OP_RELU: case PG_OP_RESHAPE: case PG_OP_SOFTMAX: case PG_OP_DROPOUT:
        if (P.size[i] != sa) return bad(i, "DimensionMismatch: elementwise / reshape output must have the input's number of elements");
        if (o.kind == PG_OP_SOFTMAX && (o.ndim < 1 || a->ndim < 1 || o.shape[o.ndim - 1] != a->shape[a->ndim - 1]))
          return bad(i, "softmax keeps the class axis");
        if (o.kind == PG_OP_DROPOUT && !(o.p >= 0.0 && o.p < 1.0)) return bad(i, "dropout probability must be in [0, 1)");
        break;
      case PG_OP_LOSS_CE: {
        const pg_op& y = ops[o.in[1]];
        if (a->ndim < 1 || sa <= 0) return bad(i, "cross_entropy of an empty tensor");
        const int64_t C = a->shape[a->ndim - 1];
        const bool idx = y.dtype == PG_I32;
        if (idx ? P.size[o.in[1]] != sa / C : P.size[o.in[1]] != sa)
          return bad(i, "DimensionMismatch: labels must match probs (dense) or hold one Int32 class per sample");
        if (o.ndim != 0) return bad(i, "a loss is a scalar (ndim 0)");
        break;
      }
      case PG_OP_LOSS_MSE:
        if (P.size[o.in[1]] != sa || sa <= 0) return bad(i, "DimensionMismatch: mse output and label sizes differ");
        if (o.ndim != 0) return bad(i, "a loss is a scalar (ndim 0)");
        break;
      default: break;
    }
  }
  P.need.assign(n, 0);
  P.need[P.root] = 1;
  for (int i = n - 1; i >= 0; --i)
    if (P.need[i])
      for (int k = 0; k < 3; ++k)
        if (ops[i].in[k] >= 0) P.need[ops[i].in[k]] = 1;
  P.consumers.assign(n, {});
  for (int i = 0; i < n; ++i)
    if (P.need[i])
      for (int k = 0; k < 3; ++k)
        if (ops[i].in[k] >= 0) P.consumers[ops[i].in[k]].push_back(i);
  for (int i = 0; i < n; ++i) {
    if (!P.need[i]) continue;
    const bool leafnode = P.is(i, PG_OP_INPUT) || P.is(i, PG_OP_PARAM);
    if (!leafnode && P.consumers[i].size() > 1) {
      pg_set_error("the B200 path supports sequential graphs only: an intermediate value is used twice");
      return PG_ERR_UNSUPPORTED;
    }
    if (P.is(i, PG_OP_PARAM) && ops[i].grad_leaf >= 0 && P.consumers[i].size() > 1) {
      pg_set_error("weight sharing (a trainable leaf used by two layers) is not supported");
      return PG_ERR_UNSUPPORTED;
    }
  }
  P.ng.assign(n, 0);
  for (int i = 0; i < n; ++i) {
    if (P.is(i, PG_OP_PARAM)) P.ng[i] = ops[i].grad_leaf >= 0;
    else if (!P.is(i, PG_OP_INPUT))
      for (int k = 0; k < 3; ++k)
        if (ops[i].in[k] >= 0 && P.ng[ops[i].in[k]]) P.ng[i] = 1;
  }
  P.fused_act.assign(n, -1); P.fused_into.assign(n, -1); P.fused_sce.assign(n, -1);
  P.sce_softmax.assign(n, 0);
  for (int i = 0; i < n; ++i) {
    if (!P.need[i]) continue;
    if (P.is(i, PG_OP_RELU)) {
      const int p = ops[i].in[0];
      if ((P.is(p, PG_OP_LINEAR) || P.is(p, PG_OP_CONV2D)) && P.consumers[p].size() == 1) { P.fused_act[p] = i; P.fused_into[i] = p; }
    }
    if (P.is(i, PG_OP_LOSS_CE)) {
      const int p = ops[i].in[0];
      if (P.is(p, PG_OP_SOFTMAX) && P.consumers[p].size() == 1 && p != P.root) { P.fused_sce[i] = p; P.sce_softmax[p] = 1; }
    }
  }
  // Conv + ReLU + 2x2 max-pool chains (examples/02_mnist_conv.jl:27-33) run as one fused kernel each way when the
  // geometry is compiled in (Float32, any math mode: the kernel computes in fp32)
  P.fused_pool.assign(n, -1); P.pool_of.assign(n, -1);
  for (int i = 0; i < n; ++i) {
    if (!P.need[i] || !P.is(i, PG_OP_CONV2D) || ops[i].dtype != PG_F32 || P.fused_act[i] < 0) continue;
    const int r = P.fused_act[i];
    if (P.consumers[r].size() != 1 || r == P.root) continue;
    const int q = P.consumers[r][0];
    if (!P.is(q, PG_OP_MAXPOOL) || ops[q].k != 2) continue;
    const pg_op& w = ops[ops[i].in[1]];
    const pg_op& x = ops[ops[i].in[0]];
    if (x.dtype != PG_F32 || x.ndim != 4) continue;
    if (pg_conv2d_relu_pool2_supported((int)w.shape[1], (int)x.shape[2], (int)x.shape[3], (int)w.shape[0], (int)w.shape[2], (int)w.shape[3]) != 1) continue;
    P.fused_pool[i] = q; P.pool_of[q] = i;
  }
  P.dtype = ops[P.root].dtype;
  return PG_OK;
}

// Which Linear / Conv nodes run on the tcgen05 kernels in PG_MATH_BF16 mode.
//   Linear: Float32 model, batch % 8 == 0 (TMA pitch of the transposed operands).  Its input may be a trace input, an
//   fp32 tensor (cast per step) or the bf16 activation of another tensor-core Linear.  A wide (N > 32) layer stores
//   its activation in bf16 and expects a bf16 gradient back, so its consumer (through reshapes that keep the last
//   axis, and dropouts) must be a tensor-core Linear too; otherwise it stays on the fp32 SIMT GEMM.
//   Conv: Float32 tensors in and out, any position, shapes inside pg_tc_conv2d_supported.
void select_tensor_core(pg_plan& P) {
  const int n = P.n;
  P.tc_lin.assign(n, 0); P.tc_conv.assign(n, 0);
  if (P.math != PG_MATH_BF16) return;
  for (int i = 0; i < n; ++i) {
    if (!P.need[i] || P.ops[i].dtype != PG_F32) continue;
    if (P.is(i, PG_OP_CONV2D)) {
      const pg_op& w = P.ops[P.ops[i].in[1]];
      const pg_op& x = P.ops[P.ops[i].in[0]];
      if (P.fused_pool[i] < 0 && x.dtype == PG_F32 && pg_tc_conv2d_supported((int)w.shape[1], (int)x.shape[2], (int)x.shape[3], (int)w.shape[0], (int)w.shape[2], (int)w.shape[3]) == 1)
        P.tc_conv[i] = 1;
    }
    if (!P.is(i, PG_OP_LINEAR)) continue;
    const int x = P.ops[i].in[0];
    const int64_t M = P.rows(x);
    if (M % 8 != 0 || M <= 0) continue;
    const int xd = P.ops[P.base(x)].dtype;
    if (xd == PG_BF16 && P.last(x) % 8 != 0) continue;          // a pre-cast input must already have a 16-byte pitch
    if (xd != PG_F32 && xd != PG_BF16) continue;
    P.tc_lin[i] = 1;
  }
  bool changed = true;
  while (changed) {               // demote wide layers whose consumer left the tensor-core chain
    changed = false;
    for (int i = n - 1; i >= 0; --i) {
      if (!P.tc_lin[i] || P.last(i) <= 32) continue;
      const int out = P.fused_act[i] >= 0 ? P.fused_act[i] : i;
      bool ok = true;
      int c = P.consumers[out].empty() ? -1 : P.consumers[out][0];
      while (c >= 0 && (P.is(c, PG_OP_RESHAPE) || P.is(c, PG_OP_DROPOUT))) {
        if (P.is(c, PG_OP_RESHAPE) && P.last(c) != P.last(out)) ok = false;    // padded rows cannot be re-flattened
        c = P.consumers[c].empty() ? -1 : P.consumers[c][0];
      }
      // a wide layer that ends the graph (inference of hidden features) or feeds anything but a tensor-core Linear
      if (c < 0 || !P.is(c, PG_OP_LINEAR) || !P.tc_lin[c] || P.ops[c].in[0] == -1) ok = false;
      if (ok && P.source(P.ops[c].in[0]) != out) ok = false;   // must feed the consumer's x operand (not W / b)
      if (!ok) { P.tc_lin[i] = 0; changed = true; }
    }
  }
}

// value type / pitch of every node (bf16 rows of wide tensor-core layers are padded to a multiple of 8 elements)
int infer_layout(pg_plan& P) {
  const int n = P.n;
  P.vdt.assign(n, PG_F32); P.vpitch.assign(n, 0);
  for (int i = 0; i < n; ++i) {
    if (!P.need[i]) continue;
    const pg_op& o = P.ops[i];
    P.vdt[i] = o.dtype; P.vpitch[i] = P.last(i);
    switch (o.kind) {
      case PG_OP_LINEAR:
        if (P.tc_lin[i] && P.last(i) > 32) { P.vdt[i] = PG_BF16; P.vpitch[i] = ceil8(P.last(i)); }
        break;
      case PG_OP_RELU:
        if (P.fused_into[i] >= 0) { P.vdt[i] = P.vdt[o.in[0]]; P.vpitch[i] = P.vpitch[o.in[0]]; }
        break;
      case PG_OP_RESHAPE:
        P.vdt[i] = P.vdt[o.in[0]];
        P.vpitch[i] = (P.vpitch[o.in[0]] != P.last(o.in[0])) ? P.vpitch[o.in[0]] : P.last(i);
        break;
      case PG_OP_DROPOUT: {
        const int x = o.in[0], c = P.down(i);
        P.vdt[i] = P.vdt[x]; P.vpitch[i] = P.vpitch[x];
        // dropout of a trace input in front of a tensor-core Linear (the first layer of both example scripts):
        // mask, cast and pad in one pass
        if (P.is(P.base(x), PG_OP_INPUT) && P.vdt[x] == PG_F32 && c >= 0 && P.is(c, PG_OP_LINEAR) && P.tc_lin[c] && P.down_reshape(i) == c &&
            P.consumers[i].size() == 1 && P.consumers[i][0] == c) {
          P.vdt[i] = PG_BF16; P.vpitch[i] = ceil8(P.last(i));
        }
        break;
      }
      default: break;
    }
  }
  return PG_OK;
}

int dalloc(pg_plan& P, size_t bytes, void** out) {
  void* p = nullptr;
  int rc = pg_malloc(P.ctx, bytes ? bytes : 16, &p);
  if (rc != PG_OK) return rc;
  rc = pg_memset0(P.ctx, p, bytes ? bytes : 16);
  if (rc != PG_OK) return rc;
  P.owned.push_back(p);
  *out = p;
  return PG_OK;
}

int new_buf(pg_plan& P, int node, int dt, int64_t pitch, Buf* b) {
  b->dtype = dt; b->cols = P.last(node); b->rows = P.rows(node); b->pitch = pitch;
  return dalloc(P, (size_t)b->rows * pitch * dsize(dt), &b->p);
}

int allocate(pg_plan& P) {
  if (P.allocated) return PG_OK;
  const int n = P.n;
  int rc;
  P.vals.assign(n, Buf()); P.grads.assign(n, Buf());
  rc = dalloc(P, LOSS_RING * 8 * 2, &P.loss_ring);
  if (rc != PG_OK) return rc;
  for (int i = 0; i < n; ++i) {
    if (!P.need[i]) continue;
    const int k = P.ops[i].kind;
    bool own = false;
    if (k == PG_OP_LINEAR || k == PG_OP_CONV2D) own = P.fused_act[i] < 0;
    else if (k == PG_OP_RELU) own = !(P.fused_into[i] >= 0 && P.fused_pool[P.fused_into[i]] >= 0);   // never materialised when the pool is fused in
    else if (k == PG_OP_MAXPOOL || k == PG_OP_MEANPOOL || k == PG_OP_DROPOUT) own = true;
    else if (k == PG_OP_SOFTMAX) own = !P.sce_softmax[i];
    if (own) {
      rc = new_buf(P, i, P.vdt[i], P.vpitch[i], &P.vals[i]);
      if (rc != PG_OK) return rc;
    }
  }
  for (int i = 0; i < n; ++i) {
    if (!P.need[i] || !P.ng[i]) continue;
    const int k = P.ops[i].kind;
    if (k == PG_OP_INPUT || k == PG_OP_PARAM || k == PG_OP_LOSS_CE || k == PG_OP_LOSS_MSE || k == PG_OP_RESHAPE) continue;
    if ((k == PG_OP_LINEAR || k == PG_OP_CONV2D) && P.fused_act[i] >= 0) continue;   // shares the relu node's gradient buffer
    if (k == PG_OP_SOFTMAX && P.sce_softmax[i]) continue;                              // dZ goes straight to the logits' gradient
    if (k == PG_OP_RELU && P.fused_into[i] >= 0 && P.fused_pool[P.fused_into[i]] >= 0) continue;   // the fused backward starts from the pooled gradient
    rc = new_buf(P, i, P.vdt[i], P.vpitch[i], &P.grads[i]);
    if (rc != PG_OK) return rc;
  }
  P.amax.assign(n, nullptr);
  for (int i = 0; i < n; ++i) {
    if (!P.need[i] || P.pool_of[i] < 0 || !P.ng[P.pool_of[i]]) continue;
    rc = dalloc(P, (size_t)P.size[i], &P.amax[i]);
    if (rc != PG_OK) return rc;
  }
  for (int i = 0; i < n; ++i) {
    if (!P.need[i] || !P.tc_lin[i]) continue;
    Lin L;
    const int x = P.ops[i].in[0], W = P.ops[i].in[1];
    L.K = P.ops[W].shape[0]; L.N = P.ops[W].shape[1]; L.M = P.rows(x);
    L.wide = L.N > 32;
    L.Kp = ceil8(L.K); L.Np = L.wide ? ceil8(L.N) : L.N;
    const bool x_direct = P.vdt[x] == PG_BF16 && P.vpitch[x] == L.Kp;
    if (!x_direct) { rc = dalloc(P, (size_t)L.M * L.Kp * 2, &L.xbf); if (rc != PG_OK) return rc; }
    rc = dalloc(P, (size_t)L.Kp * L.Np * 2, &L.wbf); if (rc != PG_OK) return rc;     // (unused when a valid unpadded shadow is passed)
    const bool padded = L.Kp != L.K || L.Np != L.N;
    if (L.Np != L.N) { rc = dalloc(P, (size_t)L.Np * 4, &L.bpad); if (rc != PG_OK) return rc; }
    if (padded && P.ng[W]) { rc = dalloc(P, (size_t)L.Kp * L.Np * 4, &L.dWp); if (rc != PG_OK) return rc; }
    if (L.Np != L.N && P.ng[P.ops[i].in[2]]) { rc = dalloc(P, (size_t)L.Np * 4, &L.dbp); if (rc != PG_OK) return rc; }
    if (P.ng[x] && P.vdt[x] != PG_BF16 && L.Kp != L.K) { rc = dalloc(P, (size_t)L.M * L.Kp * 4, &L.dxf); if (rc != PG_OK) return rc; }
    P.lin[i] = L;
  }
  for (int i = 0; i < n; ++i) {
    if (!P.need[i] || !P.is(i, PG_OP_LOSS_CE)) continue;
    const int y = P.ops[i].in[1];
    if (P.ops[y].dtype == PG_I32 && P.fused_sce[i] < 0) {
      void* p; rc = dalloc(P, (size_t)P.size[P.ops[i].in[0]] * dsize(P.dtype), &p); if (rc != PG_OK) return rc;
      P.onehot[i] = p;
    }
  }
  P.allocated = true;
  return PG_OK;
}

// value / gradient buffer of a node, following reshape aliases and ReLU fusion
Buf val(const pg_plan& P, int i) {
  const int k = P.ops[i].kind;
  if (k == PG_OP_RESHAPE) {
    Buf b = val(P, P.ops[i].in[0]);
    if (b.pitch == b.cols) { b.cols = b.pitch = P.last(i); b.rows = P.rows(i); }
    return b;
  }
  if ((k == PG_OP_LINEAR || k == PG_OP_CONV2D) && P.fused_act[i] >= 0) return P.vals[P.fused_act[i]];
  if (k == PG_OP_INPUT || k == PG_OP_PARAM) {
    Buf b; b.p = const_cast<void*>(P.leaf[i]); b.dtype = P.ops[i].dtype; b.cols = b.pitch = P.last(i); b.rows = P.rows(i);
    return b;
  }
  return P.vals[i];
}
Buf grad(const pg_plan& P, int i) {
  const int k = P.ops[i].kind;
  if (k == PG_OP_RESHAPE) {
    Buf b = grad(P, P.ops[i].in[0]);
    if (b.pitch == b.cols) { b.cols = b.pitch = P.last(i); b.rows = P.rows(i); }
    return b;
  }
  if ((k == PG_OP_LINEAR || k == PG_OP_CONV2D) && P.fused_act[i] >= 0) return P.grads[P.fused_act[i]];
  return P.grads[i];
}

#define RC(expr) do { int _rc = (expr); if (_rc != PG_OK) return _rc; } while (0)

int need_float(const Buf& b) {
  if (b.dtype == PG_BF16) {
    pg_set_error("plan: a bf16 activation feeds a non-Linear op; use PG_MATH_FP32 for this model");
    return PG_ERR_UNSUPPORTED;
  }
  return PG_OK;
}

uint32_t stream_of(const pg_plan& P, int node) { return (uint32_t)(((P.rng_step & 0xFFFFF) << 12) | ((uint64_t)node & 0xFFF)); }

int run_dropout(pg_plan& P, int i, const Buf& src, const Buf& dst) {
  const int64_t per_sample = P.ops[i].ndim ? P.size[i] / std::max<int64_t>(P.ops[i].shape[0], 1) : 1;
  return pg_dropout_ex(P.ctx, src.dtype, dst.dtype, src.p, dst.p, src.rows, src.cols, src.pitch, dst.pitch, P.ops[i].p,
                       P.rng_seed, stream_of(P, i), (uint64_t)(P.sample_offset * per_sample));
}

int forward_nodes(pg_plan& P);

// A data-parallel exchange of the previous step may still be updating the parameters of the LATER layers (pg_ctx::
// deferred_*): the layers in front of the first one that reads them run meanwhile; the wait is inserted right there.
int forward(pg_plan& P) {
  pg_ctx* ctx = P.ctx;
  ctx->deferred_hold = ctx->deferred_armed && !ctx->capturing;
  const int rc = forward_nodes(P);
  ctx->deferred_hold = false;
  return rc;
}

int forward_nodes(pg_plan& P) {
  pg_ctx* ctx = P.ctx;
  const int n = P.n;
  void* loss_slot = (char*)P.loss_ring + (size_t)(P.gen % LOSS_RING) * 16;
  for (int i = 0; i < n; ++i) {
    if (!P.need[i]) continue;
    const pg_op& o = P.ops[i];
    if (ctx->deferred_armed && (o.kind == PG_OP_LINEAR || o.kind == PG_OP_CONV2D)) {
      const int W = o.in[1], b = o.in[2];
      if (pg_deferred_overlaps(ctx, P.leaf[W], (size_t)P.size[W] * dsize(P.ops[W].dtype)) ||
          pg_deferred_overlaps(ctx, P.leaf[b], (size_t)P.size[b] * dsize(P.ops[b].dtype))) {
        pg_flush_deferred(ctx);
        ctx->deferred_hold = false;
      }
    }
    switch (o.kind) {
      case PG_OP_INPUT: case PG_OP_PARAM:
        if (P.leaf[i] == nullptr) return P.fail(PG_ERR_ARG, "a needed input / parameter node has no device pointer");
        break;
      case PG_OP_RESHAPE: break;
      case PG_OP_LINEAR: {
        const int x = o.in[0], W = o.in[1], b = o.in[2];
        const int64_t K = P.ops[W].shape[0], N = P.ops[W].shape[1], M = P.rows(x);
        const int act = P.fused_act[i] >= 0 ? PG_ACT_RELU : PG_ACT_NONE;
        Buf out = val(P, i), xv = val(P, x);
        if (P.tc_lin[i]) {
          Lin& L = P.lin[i];
          const void* xp = xv.p;
          if (L.xbf) {            // fp32 (or unpadded) source: cast (+ zero-pad) into the staging operand
            if (xv.dtype == PG_F32 && xv.pitch == K && L.Kp == K) RC(pg_vec_cast_bf16(ctx, L.xbf, xv.p, M * K));     // 128-bit vectorised
            else RC(pg_dropout_ex(ctx, xv.dtype, PG_BF16, xv.p, L.xbf, M, K, xv.pitch, L.Kp, 0.0, 0, 0, 0));
            xp = L.xbf;
          }
          const void* wp;
          if (L.Kp == K && L.Np == N && !P.leaf_bf.empty() && P.leaf_bf[W] != nullptr) wp = P.leaf_bf[W];
          else { RC(pg_dropout_ex(ctx, PG_F32, PG_BF16, P.leaf[W], L.wbf, K, N, N, L.Np, 0.0, 0, 0, 0)); wp = L.wbf; }
          const void* bp = P.leaf[b];
          if (L.bpad) { RC(pg_copy_d2d(ctx, L.bpad, P.leaf[b], (size_t)N * 4)); bp = L.bpad; }
          L.x_used = xp; L.w_used = wp;
          RC(pg_tc_linear_fwd(ctx, xp, wp, bp, out.p, out.dtype, M, L.Kp, L.Np, act));
        } else {
          RC(need_float(xv));
          RC(pg_linear_fwd(ctx, o.dtype, xv.p, P.leaf[W], P.leaf[b], out.p, M, K, N, act));
        }
        break;
      }
      case PG_OP_CONV2D: {
        const int x = o.in[0], w = o.in[1], b = o.in[2];
        const pg_op& xs = P.ops[x]; const pg_op& ws = P.ops[w];
        const int act = P.fused_act[i] >= 0 ? PG_ACT_RELU : PG_ACT_NONE;
        Buf xv = val(P, x);
        RC(need_float(xv));
        if (P.fused_pool[i] >= 0) {
          const int q = P.fused_pool[i];
          RC(pg_conv2d_relu_pool2_fwd(ctx, xv.p, P.leaf[w], P.leaf[b], P.vals[q].p, P.amax[q], xs.shape[0], (int)xs.shape[1], (int)xs.shape[2], (int)xs.shape[3],
                                      (int)ws.shape[0], (int)ws.shape[2], (int)ws.shape[3]));
          break;
        }
        Buf out = val(P, i);
        if (P.tc_conv[i])
          RC(pg_tc_conv2d_fwd(ctx, xv.p, P.leaf[w], P.leaf[b], out.p, xs.shape[0], (int)xs.shape[1], (int)xs.shape[2], (int)xs.shape[3], (int)ws.shape[0], (int)ws.shape[2], (int)ws.shape[3], act));
        else
          RC(pg_conv2d_fwd(ctx, o.dtype, xv.p, P.leaf[w], P.leaf[b], out.p, xs.shape[0], (int)xs.shape[1], (int)xs.shape[2], (int)xs.shape[3], (int)ws.shape[0], (int)ws.shape[2], (int)ws.shape[3], act));
        break;
      }
      case PG_OP_RELU: {
        if (P.fused_into[i] >= 0) break;
        Buf xv = val(P, o.in[0]);
        RC(need_float(xv));
        RC(pg_relu_fwd(ctx, o.dtype, xv.p, P.vals[i].p, P.size[i]));
        break;
      }
      case PG_OP_MAXPOOL: case PG_OP_MEANPOOL: {
        if (P.pool_of[i] >= 0) break;              // produced by the fused conv kernel
        const pg_op& xs = P.ops[o.in[0]];
        Buf xv = val(P, o.in[0]);
        RC(need_float(xv));
        if (o.kind == PG_OP_MAXPOOL) RC(pg_maxpool2d_fwd(ctx, o.dtype, xv.p, P.vals[i].p, xs.shape[0] * xs.shape[1], (int)xs.shape[2], (int)xs.shape[3], o.k));
        else RC(pg_meanpool2d_fwd(ctx, o.dtype, xv.p, P.vals[i].p, xs.shape[0] * xs.shape[1], (int)xs.shape[2], (int)xs.shape[3], o.k));
        break;
      }
      case PG_OP_DROPOUT: RC(run_dropout(P, i, val(P, o.in[0]), P.vals[i])); break;
      case PG_OP_SOFTMAX: {
        if (P.sce_softmax[i]) break;
        Buf xv = val(P, o.in[0]);
        RC(need_float(xv));
        const int C = (int)P.last(i);
        RC(pg_softmax_fwd(ctx, o.dtype, xv.p, P.vals[i].p, P.size[i] / C, C));
        break;
      }
      case PG_OP_LOSS_CE: {
        const int p = o.in[0], y = o.in[1];
        const int C = (int)P.last(p);
        const int64_t B = P.size[p] / C;
        const double gb = o.denom > 0 ? o.denom : (double)B;
        const bool idx = P.ops[y].dtype == PG_I32;
        P.vals[i].p = loss_slot; P.vals[i].dtype = o.dtype;
        if (P.fused_sce[i] >= 0) {
          const int z = P.ops[P.fused_sce[i]].in[0];
          Buf zv = val(P, z);
          RC(need_float(zv));
          void* dz = P.ng[z] ? grad(P, z).p : nullptr;
          if (idx) RC(pg_softmax_ce_idx_fwd_bwd(ctx, o.dtype, zv.p, (const int32_t*)P.leaf[y], nullptr, dz, loss_slot, B, C, 1.0 / gb));
          else RC(pg_softmax_ce_fwd_bwd(ctx, o.dtype, zv.p, val(P, y).p, nullptr, dz, loss_slot, B, C, 1.0 / gb));
        } else {
          Buf pv = val(P, p);
          RC(need_float(pv));
          const void* yp = val(P, y).p;
          if (idx) { RC(pg_onehot(ctx, o.dtype, (const int32_t*)P.leaf[y], P.onehot[i], B, C)); yp = P.onehot[i]; }
          void* dp = P.ng[p] ? grad(P, p).p : nullptr;
          RC(pg_cross_entropy_fwd_bwd(ctx, o.dtype, pv.p, yp, dp, loss_slot, B, C, 1.0 / gb));
        }
        break;
      }
      case PG_OP_LOSS_MSE: {
        const int yh = o.in[0], y = o.in[1];
        Buf a = val(P, yh);
        RC(need_float(a));
        const double denom = o.denom > 0 ? o.denom : (double)P.size[yh];
        P.vals[i].p = loss_slot; P.vals[i].dtype = o.dtype;
        void* dy = P.ng[yh] ? grad(P, yh).p : nullptr;
        RC(pg_mse_fwd_bwd(ctx, o.dtype, a.p, val(P, y).p, dy, loss_slot, P.size[yh], 1.0 / denom));
        break;
      }
      default: return P.fail(PG_ERR_ARG, "unknown op kind");
    }
  }
  return PG_OK;
}

int backward(pg_plan& P, void* const* gleaf, int n_leaves, pg_ready_fn on_ready, void* user) {
  pg_ctx* ctx = P.ctx;
  auto leaf_of = [&](int node) -> void* {
    const int gl = P.ops[node].grad_leaf;
    return (P.ng[node] && gl >= 0 && gl < n_leaves) ? gleaf[gl] : nullptr;
  };
  for (int i = P.n - 1; i >= 0; --i) {
    if (!P.need[i] || !P.ng[i]) continue;
    const pg_op& o = P.ops[i];
    switch (o.kind) {
      case PG_OP_INPUT: case PG_OP_PARAM: case PG_OP_RESHAPE: case PG_OP_LOSS_CE: case PG_OP_LOSS_MSE:
        break;                       // losses wrote their input gradient during the forward pass
      case PG_OP_LINEAR: case PG_OP_CONV2D: {
        const int x = o.in[0], W = o.in[1], b = o.in[2];
        if (o.kind == PG_OP_CONV2D && P.fused_pool[i] >= 0) {
          // pooled gradient -> (max-pool routing, ReLU', conv pullbacks) in one pass
          const int q = P.fused_pool[i];
          const pg_op& xs = P.ops[x]; const pg_op& ws = P.ops[W];
          void* gW = leaf_of(W); void* gb = leaf_of(b);
          if (P.ng[W] && !gW) return P.fail(PG_ERR_ARG, "missing gradient leaf pointer");
          Buf gxb = P.ng[x] ? grad(P, x) : Buf();
          if (gxb.p) RC(need_float(gxb));
          RC(pg_conv2d_relu_pool2_bwd(ctx, val(P, x).p, P.leaf[W], P.vals[q].p, P.amax[q], P.grads[q].p, gW, gb, gxb.p, xs.shape[0], (int)xs.shape[1],
                                      (int)xs.shape[2], (int)xs.shape[3], (int)ws.shape[0], (int)ws.shape[2], (int)ws.shape[3]));
          int first_leaf = 1 << 30;
          if (gW) first_leaf = std::min(first_leaf, P.ops[W].grad_leaf);
          if (gb) first_leaf = std::min(first_leaf, P.ops[b].grad_leaf);
          if (on_ready && first_leaf != (1 << 30)) on_ready(user, first_leaf);
          break;
        }
        Buf gy = grad(P, i);
        const int r = P.fused_act[i];
        if (r >= 0 && !P.premasked(r)) {   // the consumer does not apply relu' itself: gy becomes the pre-activation gradient here
          RC(need_float(gy));
          RC(pg_relu_bwd(ctx, o.dtype, P.vals[r].p, gy.p, gy.p, P.size[r]));
        }
        void* gW = leaf_of(W); void* gb = leaf_of(b);
        if (P.ng[W] && !gW) return P.fail(PG_ERR_ARG, "missing gradient leaf pointer");
        Buf gxb = P.ng[x] ? grad(P, x) : Buf();
        int first_leaf = 1 << 30;
        if (gW) first_leaf = std::min(first_leaf, P.ops[W].grad_leaf);
        if (gb) first_leaf = std::min(first_leaf, P.ops[b].grad_leaf);
        if (o.kind == PG_OP_LINEAR) {
          const int64_t K = P.ops[W].shape[0], N = P.ops[W].shape[1], M = P.rows(x);
          if (P.tc_lin[i]) {
            Lin& L = P.lin[i];
            const bool mask = gxb.p != nullptr && P.relu_upstream(x);
            void* dWt = gW ? (L.dWp ? L.dWp : gW) : nullptr;
            void* dbt = gb ? (L.dbp ? L.dbp : gb) : nullptr;
            void* dxt = nullptr; int dxd = PG_BF16;
            if (gxb.p) {
              if (gxb.dtype == PG_BF16) { dxt = gxb.p; dxd = PG_BF16; }
              else { dxt = L.dxf ? L.dxf : gxb.p; dxd = PG_F32; }
            }
            auto finish_params = [&]() -> int {
              if (gW && L.dWp) RC(pg_dropout_ex(ctx, PG_F32, PG_F32, L.dWp, gW, K, N, L.Np, N, 0.0, 0, 0, 0));
              if (gb && L.dbp) RC(pg_copy_d2d(ctx, gb, L.dbp, (size_t)N * 4));
              return PG_OK;
            };
            if (on_ready && dxt && (dWt || dbt)) {
              // data-parallel: parameter gradients first, so their bucket starts reducing while this layer's data
              // gradient (and everything upstream) still computes
              RC(pg_tc_linear_bwd2(ctx, L.x_used, nullptr, gy.p, gy.dtype, dWt, dbt, nullptr, dxd, nullptr, L.M, L.Kp, L.Np));
              RC(finish_params());
              on_ready(user, first_leaf);
              RC(pg_tc_linear_bwd2(ctx, nullptr, L.w_used, gy.p, gy.dtype, nullptr, nullptr, dxt, dxd, mask ? L.x_used : nullptr, L.M, L.Kp, L.Np));
              first_leaf = 1 << 30;
            } else {
              RC(pg_tc_linear_bwd2(ctx, dWt ? L.x_used : nullptr, dxt ? L.w_used : nullptr, gy.p, gy.dtype, dWt, dbt, dxt, dxd,
                                   mask ? L.x_used : nullptr, L.M, L.Kp, L.Np));
              RC(finish_params());
            }
            if (dxt && dxt == L.dxf) RC(pg_dropout_ex(ctx, PG_F32, PG_F32, L.dxf, gxb.p, M, K, L.Kp, gxb.pitch, 0.0, 0, 0, 0));
          } else {
            Buf xv = val(P, x);
            RC(need_float(xv)); RC(need_float(gy));
            const void* mask = (gxb.p && P.relu_upstream(x)) ? xv.p : nullptr;
            if (gxb.p) RC(need_float(gxb));
            RC(pg_linear_bwd(ctx, o.dtype, xv.p, P.leaf[W], gy.p, gW, gb, gxb.p, mask, M, K, N));
          }
        } else {
          const pg_op& xs = P.ops[x]; const pg_op& ws = P.ops[W];
          Buf xv = val(P, x);
          if (gxb.p) RC(need_float(gxb));
          if (P.tc_conv[i] && (gW != nullptr || gb == nullptr))
            RC(pg_tc_conv2d_bwd(ctx, xv.p, P.leaf[W], gy.p, gW, gb, gxb.p, xs.shape[0], (int)xs.shape[1], (int)xs.shape[2], (int)xs.shape[3], (int)ws.shape[0], (int)ws.shape[2], (int)ws.shape[3]));
          else
            RC(pg_conv2d_bwd(ctx, o.dtype, xv.p, P.leaf[W], gy.p, gW, gb, gxb.p, xs.shape[0], (int)xs.shape[1], (int)xs.shape[2], (int)xs.shape[3], (int)ws.shape[0], (int)ws.shape[2], (int)ws.shape[3]));
        }
        if (on_ready && first_leaf != (1 << 30)) on_ready(user, first_leaf);
        break;
      }
      case PG_OP_RELU: {
        if (P.fused_into[i] >= 0) break;     // handled by the producer
        Buf gx = grad(P, o.in[0]);
        RC(need_float(gx));
        if (P.premasked(i)) RC(pg_copy_d2d(ctx, gx.p, P.grads[i].p, (size_t)P.size[i] * dsize(o.dtype)));
        else RC(pg_relu_bwd(ctx, o.dtype, P.vals[i].p, P.grads[i].p, gx.p, P.size[i]));
        break;
      }
      case PG_OP_MAXPOOL: case PG_OP_MEANPOOL: {
        if (P.pool_of[i] >= 0) break;              // handled by the fused conv backward
        const int x = o.in[0];
        const pg_op& xs = P.ops[x];
        Buf gx = grad(P, x);
        RC(need_float(gx));
        if (o.kind == PG_OP_MAXPOOL) {
          const int fuse_relu = P.relu_upstream(x) ? 1 : 0;
          RC(pg_maxpool2d_relu_bwd(ctx, o.dtype, val(P, x).p, P.grads[i].p, gx.p, xs.shape[0] * xs.shape[1], (int)xs.shape[2], (int)xs.shape[3], o.k, fuse_relu));
        } else {
          RC(pg_meanpool2d_bwd(ctx, o.dtype, P.grads[i].p, gx.p, xs.shape[0] * xs.shape[1], (int)xs.shape[2], (int)xs.shape[3], o.k));
        }
        break;
      }
      case PG_OP_DROPOUT:
        RC(run_dropout(P, i, P.grads[i], grad(P, o.in[0])));      // same (seed, step, node, element) key -> same mask
        break;
      case PG_OP_SOFTMAX: {
        if (P.sce_softmax[i]) break;         // dZ was written by the fused loss kernel
        const int C = (int)P.last(i);
        RC(pg_softmax_bwd(ctx, o.dtype, P.vals[i].p, P.grads[i].p, grad(P, o.in[0]).p, P.size[i] / C, C));
        break;
      }
      default: return P.fail(PG_ERR_ARG, "unknown op kind");
    }
  }
  return PG_OK;
}

}  // namespace

extern "C" {

int pg_plan_create(pg_ctx* ctx, const pg_op* ops, int n_ops, int root, int math_mode, pg_plan** out) {
  PG_REQUIRE(out != nullptr && ops != nullptr, "null pointer");
  *out = nullptr;
  PG_REQUIRE(n_ops > 0 && root >= 0 && root < n_ops, "root out of range");
  PG_REQUIRE(math_mode == PG_MATH_FP32 || math_mode == PG_MATH_BF16, "math mode must be PG_MATH_FP32 or PG_MATH_BF16");
  pg_plan* P = new pg_plan();
  P->ctx = ctx; P->math = math_mode; P->root = root; P->n = n_ops;
  P->ops.assign(ops, ops + n_ops);
  int rc = analyse(*P);
  if (rc == PG_OK) { select_tensor_core(*P); rc = infer_layout(*P); }
  if (rc != PG_OK) { delete P; return rc; }
  *out = P;
  return PG_OK;
}

int pg_plan_destroy(pg_plan* P) {
  if (!P) return PG_OK;
  if (P->ctx) {
    if (!P->owned.empty()) cudaStreamSynchronize(P->ctx->stream);
    for (void* p : P->owned) cudaFree(p);
  }
  delete P;
  return PG_OK;
}

int pg_plan_node_flags(pg_plan* P, int node, int* flags) {
  PG_REQUIRE(P && flags && node >= 0 && node < P->n, "bad plan / node");
  int f = 0;
  if (P->need[node]) f |= PG_NODE_NEEDED;
  if (P->ng[node]) f |= PG_NODE_NEEDS_GRAD;
  if (P->tc_lin[node] || P->tc_conv[node]) f |= PG_NODE_TENSOR_CORE;
  if (P->fused_act[node] >= 0) f |= PG_NODE_FUSED_ACT;
  if (P->fused_pool[node] >= 0) f |= PG_NODE_FUSED_POOL;
  if (P->need[node] && P->vdt[node] == PG_BF16 && !P->is(node, PG_OP_INPUT)) f |= PG_NODE_BF16_VALUE;
  if ((P->is(node, PG_OP_SOFTMAX) && P->sce_softmax[node]) || ((P->is(node, PG_OP_LINEAR) || P->is(node, PG_OP_CONV2D)) && P->fused_act[node] >= 0))
    f |= PG_NODE_FUSED_AWAY;
  *flags = f;
  return PG_OK;
}

int pg_plan_forward(pg_plan* P, const void* const* leaf_ptrs, const void* const* leaf_bf16, uint64_t rng_seed, uint64_t rng_step,
                    int64_t sample_offset, void** loss_dev, int64_t* generation) {
  PG_REQUIRE(P != nullptr && leaf_ptrs != nullptr, "null plan / leaf pointers");
  PG_CHECK_CTX(P->ctx);
  for (int i = 0; i < P->n; ++i)
    if (P->need[i] && (P->is(i, PG_OP_INPUT) || P->is(i, PG_OP_PARAM)) && leaf_ptrs[i] == nullptr) {
      pg_set_error("pg_plan_forward: no device pointer bound for %s node %d", P->is(i, PG_OP_INPUT) ? "INPUT" : "PARAM", i);
      return PG_ERR_ARG;
    }
  int rc = allocate(*P);
  if (rc != PG_OK) return rc;
  P->leaf.assign(leaf_ptrs, leaf_ptrs + P->n);
  if (leaf_bf16) P->leaf_bf.assign(leaf_bf16, leaf_bf16 + P->n); else P->leaf_bf.clear();
  P->rng_seed = rng_seed; P->rng_step = rng_step; P->sample_offset = sample_offset;
  P->gen += 1;
  rc = forward(*P);
  if (rc != PG_OK) return rc;
  if (loss_dev) *loss_dev = (P->is(P->root, PG_OP_LOSS_CE) || P->is(P->root, PG_OP_LOSS_MSE)) ? P->vals[P->root].p : nullptr;
  if (generation) *generation = P->gen;
  return PG_OK;
}

int pg_plan_backward(pg_plan* P, int64_t generation, void* const* grad_leaf_ptrs, int n_grad_leaves, pg_ready_fn on_ready, void* user) {
  PG_REQUIRE(P != nullptr && grad_leaf_ptrs != nullptr, "null plan / gradient pointers");
  PG_CHECK_CTX(P->ctx);
  if (!P->allocated || generation != P->gen) {
    pg_set_error("stale pullback: this plan has run forward %lld time(s) since the pullback was created (generation %lld, current %lld); "
                 "call pb() before the next eval_with_pullback of the same objective", (long long)(P->gen - generation), (long long)generation, (long long)P->gen);
    return PG_ERR_ARG;
  }
  return backward(*P, grad_leaf_ptrs, n_grad_leaves, on_ready, user);
}

int pg_plan_value(pg_plan* P, int node, void** dptr, int* dtype, int64_t* pitch) {
  PG_REQUIRE(P && dptr && node >= 0 && node < P->n, "bad plan / node");
  PG_REQUIRE(P->allocated, "pg_plan_value before the first pg_plan_forward");
  Buf b = val(*P, node);
  if (b.p == nullptr) {
    pg_set_error("value of node %d was fused away; evaluate it as the root of its own trace", node);
    return PG_ERR_UNSUPPORTED;
  }
  *dptr = b.p;
  if (dtype) *dtype = b.dtype;
  if (pitch) *pitch = b.pitch;
  return PG_OK;
}

}  // extern "C"
