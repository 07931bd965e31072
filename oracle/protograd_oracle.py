"""CPU oracle for the ProtoGrad training-step hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the arithmetic the reference (lostella/ProtoGrad.jl,
pure Julia, CPU only) performs on its training-step path.  It is the *checker* for the CUDA
path; it is never the thing shipped or measured as the product.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs import it.

Pinning status
--------------
* PINNED against the reference's own known-answer tests (replayed in tests/test_oracle.py):
  model algebra exact-equality matrix (test/test_models.jl:32-125), Linear+mse closed-form
  gradient (test/test_models.jl:150-152), optimizer convergence bounds on the 10-dim quadratic
  (test/test_optimizers.jl:28,40,56) and the Linear 10->1 regression (test/test_optimizers.jl:86).
* PARITY UNPINNED by the reference (it holds no golden vector / known-answer test for these):
  conv, maxpool/meanpool, softmax, cross_entropy, relu'(0), and per-step numerics of
  Adam/Nesterov/HeavyBall.  Their arithmetic lives in third-party Julia packages that are not
  under /root/reference: NNlib (compat "0.9"), Zygote (compat "0.7.3"), ChainRulesCore ("1")
  (Project.toml:18-24; no Manifest is committed, so versions are unpinned upstream too).  For
  these the restatement follows NNlib's published semantics (true convolution with flipped
  kernel, first-argmax pooling, max-shifted softmax) anchored on the reference call sites, and
  is cross-checked against torch-CPU float64 autograd and central finite differences in
  tests/test_oracle.py; conv / maxpool / meanpool forward values are additionally checked on the
  known-answer vectors of NNlib's own test-suite (x = reshape(1:20, 5, 4, 1, 1), w = reshape(1:4, 2, 2, 1, 1);
  quoted from the published package, which cannot be fetched here, and hand-derivable from the
  definitions) — test_nnlib_published_known_answers_conv_and_pooling.  Julia itself is not installed here, so no reference output could be
  generated; `julia/parity.jl` closes the loop on any machine with Julia.

Layout convention: everything is in row-major/C terms, i.e. exactly the bytes Julia holds:
Julia column-major A[d1,d2,...] == C array A[...][d2][d1].  So
    Linear:  X[B][in], W[in][out] (== Julia W[out,in]), b[out], Y[B][out] = X @ W + b
    Conv:    x[N][Cin][H][W] (== Julia x[W,H,Cin,N]), w[Cout][Cin][kH][kW], y[N][Cout][H'][W']
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

# ----------------------------------------------------------------------------------------------
# Layers (src/layers.jl)
# ----------------------------------------------------------------------------------------------


def xavier_uniform(rng: np.random.Generator, dtype, shape, fan_in: int, fan_out: int):
    """src/layers.jl:6-8: sqrt(T(6)/(fan_in+fan_out)) * (2*rand(T,dims...) .- 1).
    The random stream is numpy's, not Julia's Xoshiro: parity runs upload these weights."""
    T = np.dtype(dtype).type
    scale = np.sqrt(T(6) / T(fan_in + fan_out))
    return (scale * (T(2) * rng.random(shape, dtype=np.float64).astype(dtype) - T(1))).astype(dtype)


def linear_fwd(X, W, b):
    """src/layers.jl:28-34: flatten trailing dims to the batch, W*x .+ b, reshape back."""
    X2 = X.reshape(-1, W.shape[0])
    Y = X2 @ W
    Y += b
    return Y.reshape(X.shape[:-1] + (W.shape[1],))


def linear_bwd(X, W, dY, need_dx: bool = True):
    """ChainRules rrule of `*` plus broadcast unbroadcast for `.+ b` (call site layers.jl:31):
    dW = dY * x' (Julia [out,in]) == C dW[in][out] = X^T @ dY; db = sum over batch; dx = W' * dY."""
    X2 = X.reshape(-1, W.shape[0])
    dY2 = dY.reshape(-1, W.shape[1])
    dW = X2.T @ dY2
    db = dY2.sum(axis=0, dtype=dY2.dtype)
    dX = (dY2 @ W.T).reshape(X.shape) if need_dx else None
    return dW, db, dX


def conv2d_fwd(x, w, b):
    """src/layers.jl:59: NNlib.conv(x, w) .+ b.  NNlib's default (flipped=false) is a TRUE
    convolution: y[i,j,co,n] = sum x[i+kW-a, j+kH-b, ci, n] * w[a,b,ci,co] (1-based), stride 1,
    no padding.  In C terms: correlate x with the kernel flipped in both spatial dims."""
    kH, kW = w.shape[2], w.shape[3]
    win = sliding_window_view(x, (kH, kW), axis=(2, 3))  # [N,Cin,H',W',kH,kW]
    wf = w[:, :, ::-1, ::-1]
    y = np.einsum("ncijab,ocab->noij", win, wf, optimize=True).astype(x.dtype, copy=False)
    if b is not None:
        y = y + b.reshape(1, -1, 1, 1)
    return np.ascontiguousarray(y)


def conv2d_bwd(x, w, dy, need_dx: bool = True):
    """NNlib ∇conv_filter / ∇conv_data via NNlib's rrule(conv) (call site layers.jl:59).
    dw[co,ci,a,b] = sum_{n,i,j} x[n,ci,i+kH-1-a, j+kW-1-b] * dy[n,co,i,j] (0-based, flipped);
    dx = full correlation of dy with w (un-flipped); db[co] = sum dy."""
    kH, kW = w.shape[2], w.shape[3]
    win = sliding_window_view(x, (kH, kW), axis=(2, 3))  # [N,Cin,H',W',kH,kW]
    dwf = np.einsum("ncijab,noij->ocab", win, dy, optimize=True)
    dw = np.ascontiguousarray(dwf[:, :, ::-1, ::-1]).astype(x.dtype, copy=False)
    db = dy.sum(axis=(0, 2, 3), dtype=dy.dtype)
    dx = None
    if need_dx:
        # y[i,j] = sum_{p,q} x[i+p, j+q] wf[p,q]  =>  dx[u,v] = sum_{p,q} dy[u-p, v-q] wf[p,q]
        pad = np.pad(dy, ((0, 0), (0, 0), (kH - 1, kH - 1), (kW - 1, kW - 1)))
        winp = sliding_window_view(pad, (kH, kW), axis=(2, 3))  # [N,Cout,H,W,kH,kW]
        # dx[n,ci,u,v] = sum_{co,p,q} pad[u + (kH-1-p), v + (kW-1-q)] wf[co,ci,p,q]
        #             = sum_{co,a,b} winp[n,co,u,v,a,b] * w[co,ci,a,b]   (since wf flipped == w)
        dx = np.einsum("nouvab,ocab->ncuv", winp, w, optimize=True).astype(x.dtype, copy=False)
        dx = np.ascontiguousarray(dx)
    return dw, db.astype(x.dtype), dx


def relu_fwd(x):
    """NNlib relu (layers.jl:81): ifelse(x<0, 0, x)."""
    return np.where(x < 0, x.dtype.type(0), x)


def relu_bwd(y, dy):
    """NNlib broadcast rrule: derivative is [y > 0], so relu'(0) = 0."""
    return np.where(y > 0, dy, dy.dtype.type(0))


def _pool_windows(x, k):
    kH, kW = k
    N, C, H, W = x.shape
    Ho, Wo = (H - kH) // kH + 1, (W - kW) // kW + 1
    xc = x[:, :, : Ho * kH, : Wo * kW]
    return xc.reshape(N, C, Ho, kH, Wo, kW).transpose(0, 1, 2, 4, 3, 5).reshape(N, C, Ho, Wo, kH * kW), Ho, Wo


def maxpool2d_fwd(x, k=(2, 2)):
    """NNlib maxpool(x, k) (layers.jl:93): stride = window, no padding."""
    win, _, _ = _pool_windows(x, k)
    return np.ascontiguousarray(win.max(axis=-1))


def maxpool2d_bwd(x, dy, k=(2, 2)):
    """NNlib ∇maxpool: dY goes to the FIRST window element equal to the max, scanning kW fastest
    then kH (C-order flatten of [kH][kW] == numpy argmax first occurrence)."""
    kH, kW = k
    win, Ho, Wo = _pool_windows(x, k)
    arg = win.argmax(axis=-1)  # first occurrence
    dwin = np.zeros(win.shape, dtype=dy.dtype)
    np.put_along_axis(dwin, arg[..., None], dy[..., None], axis=-1)
    N, C = x.shape[:2]
    dx = np.zeros_like(x, dtype=dy.dtype)
    dx[:, :, : Ho * kH, : Wo * kW] = (
        dwin.reshape(N, C, Ho, Wo, kH, kW).transpose(0, 1, 2, 4, 3, 5).reshape(N, C, Ho * kH, Wo * kW)
    )
    return dx


def meanpool2d_fwd(x, k=(2, 2)):
    """NNlib meanpool (layers.jl:99)."""
    win, _, _ = _pool_windows(x, k)
    return np.ascontiguousarray(win.mean(axis=-1, dtype=x.dtype))


def meanpool2d_bwd(x, dy, k=(2, 2)):
    kH, kW = k
    N, C, H, W = x.shape
    Ho, Wo = dy.shape[2], dy.shape[3]
    dx = np.zeros_like(x, dtype=dy.dtype)
    g = dy / dy.dtype.type(kH * kW)
    dx[:, :, : Ho * kH, : Wo * kW] = np.repeat(np.repeat(g, kH, axis=2), kW, axis=3)
    return dx


def softmax_fwd(z):
    """NNlib softmax over the class dim (layers.jl:87; Julia dims=1 == C last axis)."""
    e = np.exp(z - z.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True, dtype=z.dtype)


def softmax_bwd(p, dp):
    """NNlib ∇softmax: dz = p .* (dp .- sum(dp .* p; dims))."""
    return p * (dp - (dp * p).sum(axis=-1, keepdims=True, dtype=p.dtype))


# ----------------------------------------------------------------------------------------------
# Losses (src/losses.jl)
# ----------------------------------------------------------------------------------------------


def mse(out, label, denom: Optional[float] = None):
    """src/losses.jl:5: mean((out-label).^2) over ALL elements.  `denom` overrides the divisor
    (README.md:86 divides the SUM by the number of samples instead)."""
    d = out - label
    n = out.size if denom is None else denom
    return out.dtype.type((d * d).sum(dtype=out.dtype) / out.dtype.type(n))


def mse_bwd(out, label, denom: Optional[float] = None):
    n = out.size if denom is None else denom
    return (out.dtype.type(2) * (out - label) / out.dtype.type(n)).astype(out.dtype)


def cross_entropy(p, y, batch: Optional[int] = None):
    """src/losses.jl:7-10: mean over batch of -sum_c y*log(p + eps(T)).  `batch` overrides the
    divisor (data-parallel shards divide by the GLOBAL batch so the all-reduce is a pure sum)."""
    T = p.dtype.type
    eps = np.finfo(p.dtype).eps
    s = -(y * np.log(p + T(eps))).sum(axis=-1, dtype=p.dtype)
    B = s.size if batch is None else batch
    return T(s.sum(dtype=p.dtype) / T(B))


def cross_entropy_bwd(p, y, batch: Optional[int] = None):
    """dp = -y / (B (p + eps)) — the eps stays inside the gradient."""
    T = p.dtype.type
    eps = np.finfo(p.dtype).eps
    B = (p.size // p.shape[-1]) if batch is None else batch
    return (-y / (T(B) * (p + T(eps)))).astype(p.dtype)


def class_error(p, y):
    """src/losses.jl:12-16."""
    return float(np.mean(p.argmax(axis=-1) != y.argmax(axis=-1)))


# ----------------------------------------------------------------------------------------------
# A layer-sequence network (Compose, src/layers.jl:63-75) and its training-step gradient
# (eval_with_pullback, src/gradient.jl + ext/ProtoGradZygoteExt.jl:6-13)
# ----------------------------------------------------------------------------------------------
# spec: list of tuples
#   ("linear", in, out [, trainable=True])      params: W[in][out], b[out]
#   ("conv", cin, cout, k [, trainable=True])   params: w[cout][cin][k][k], b[cout]
#   ("relu",) ("maxpool", k) ("meanpool", k) ("flatten",) ("softmax",)


def spec_param_shapes(spec) -> List[Tuple[int, ...]]:
    """vec(m) leaf order (src/model.jl:9-24,33): DFS, Linear=[W,b], Conv=[w,b]."""
    shapes = []
    for op in spec:
        if op[0] == "linear":
            shapes += [(op[1], op[2]), (op[2],)]
        elif op[0] == "conv":
            shapes += [(op[2], op[1], op[3], op[3]), (op[2],)]
    return shapes


def spec_trainable(spec) -> List[bool]:
    out = []
    for op in spec:
        if op[0] == "linear":
            t = op[3] if len(op) > 3 else True
            out += [t, t]
        elif op[0] == "conv":
            t = op[4] if len(op) > 4 else True
            out += [t, t]
    return out


def init_params(spec, dtype=np.float32, seed=0):
    """Linear/Conv constructors (src/layers.jl:20-24, 43-55): xavier weights, zero biases."""
    rng = np.random.default_rng(seed)
    ps = []
    for op in spec:
        if op[0] == "linear":
            ps.append(xavier_uniform(rng, dtype, (op[1], op[2]), op[1], op[2]))
            ps.append(np.zeros(op[2], dtype=dtype))
        elif op[0] == "conv":
            k = op[3]
            ps.append(xavier_uniform(rng, dtype, (op[2], op[1], k, k), k * k * op[1], k * k * op[2]))
            ps.append(np.zeros(op[2], dtype=dtype))
    return ps


def _bf16_stored(spec, ai, linear_round):
    """is activation index `ai` (acts[ai], the output of spec[ai-1]) stored in bf16 by the tensor-core Linear mode?
    True for the output of a wide (N > 32) Linear that is not followed by a ReLU, and for the ReLU output after one."""
    if linear_round is None or ai == 0:
        return False
    op = spec[ai - 1]
    if op[0] == "linear":
        return op[2] > 32 and not (ai < len(spec) and spec[ai][0] == "relu")
    if op[0] == "relu" and ai >= 2 and spec[ai - 2][0] == "linear":
        return spec[ai - 2][2] > 32
    return False


def net_forward(spec, params, x, keep=False, conv_round=None, linear_round=None):
    """conv_round: optional rounding applied to the operands of every conv (x, w) — `bf16_round` restates
    the B200 path's tensor-core conv mode (bf16 operands, exact accumulation, float outputs).
    linear_round: the same for Linear layers in the tensor-core mode: operands (x, W) rounded, products accumulated
    exactly, the activation of a wide (N > 32) layer stored rounded (after its fused ReLU), narrow layers (the logits)
    kept in float."""
    cr = conv_round if conv_round is not None else (lambda a: a)
    lr = linear_round if linear_round is not None else (lambda a: a)
    acts = [x]
    pi = 0
    y = x
    for li, op in enumerate(spec):
        kind = op[0]
        if kind == "linear":
            y = linear_fwd(lr(y), lr(params[pi]), params[pi + 1]); pi += 2
        elif kind == "conv":
            y = conv2d_fwd(cr(y), cr(params[pi]), params[pi + 1]); pi += 2
        elif kind == "relu":
            y = relu_fwd(y)
        elif kind == "maxpool":
            y = maxpool2d_fwd(y, (op[1], op[1]))
        elif kind == "meanpool":
            y = meanpool2d_fwd(y, (op[1], op[1]))
        elif kind == "flatten":
            y = y.reshape(y.shape[0], -1)
        elif kind == "softmax":
            y = softmax_fwd(y)
        else:
            raise ValueError(kind)
        if _bf16_stored(spec, li + 1, linear_round):
            y = lr(y)
        acts.append(y)
    return (y, acts) if keep else y


def net_loss_and_grad(spec, params, x, label, loss="ce", global_batch=None, mse_denom=None, conv_round=None, linear_round=None):
    """One eval_with_pullback(f, m) + pb(): returns (loss, grads) with grads in vec(m) leaf order.
    Non-trainable leaves get a zero gradient and back-propagation stops below the first
    trainable layer (the reference computes-and-discards that dX; values are identical).
    conv_round / linear_round: see net_forward; the pullbacks then see rounded x, w and dy as well, and the gradient
    of a bf16-stored activation is itself stored rounded (after the ReLU' mask)."""
    cr = conv_round if conv_round is not None else (lambda a: a)
    lr = linear_round if linear_round is not None else (lambda a: a)
    out, acts = net_forward(spec, params, x, keep=True, conv_round=conv_round, linear_round=linear_round)
    if loss == "ce":
        val = cross_entropy(out, label, global_batch)
        d = cross_entropy_bwd(out, label, global_batch)
    elif loss == "mse":
        denom = mse_denom
        if denom is None and global_batch is not None:
            denom = out.size // out.shape[0] * global_batch
        val = mse(out, label, denom)
        d = mse_bwd(out, label, denom)
    else:
        raise ValueError(loss)
    train = spec_trainable(spec)
    n_leaves = len(params)
    grads = [None] * n_leaves
    # index of first trainable parametric op: no dX needed at/below it
    pi = n_leaves
    first_train_pi = next((i for i, t in enumerate(train) if t), n_leaves)
    for li in range(len(spec) - 1, -1, -1):
        op = spec[li]
        kind = op[0]
        xin, yout = acts[li], acts[li + 1]
        if kind == "linear":
            pi -= 2
            need_dx = pi > first_train_pi
            dW, _, dx = linear_bwd(lr(xin), lr(params[pi]), lr(d), need_dx)
            db = d.reshape(-1, params[pi].shape[1]).sum(axis=0, dtype=d.dtype)   # the bias gradient sums the STORED dY (float for narrow layers)
            grads[pi], grads[pi + 1] = (dW, db) if train[pi] else (np.zeros_like(params[pi]), np.zeros_like(params[pi + 1]))
            if not need_dx:
                break
            d = dx
            if _bf16_stored(spec, li, linear_round) and spec[li - 1][0] == "linear":
                d = lr(d)
        elif kind == "conv":
            pi -= 2
            need_dx = pi > first_train_pi
            dw, db, dx = conv2d_bwd(cr(xin), cr(params[pi]), cr(d), need_dx)
            grads[pi], grads[pi + 1] = (dw, db) if train[pi] else (np.zeros_like(params[pi]), np.zeros_like(params[pi + 1]))
            if not need_dx:
                break
            d = dx
        elif kind == "relu":
            d = relu_bwd(yout, d)
            if _bf16_stored(spec, li + 1, linear_round):
                d = lr(d)
        elif kind == "maxpool":
            d = maxpool2d_bwd(xin, d, (op[1], op[1]))
        elif kind == "meanpool":
            d = meanpool2d_bwd(xin, d, (op[1], op[1]))
        elif kind == "flatten":
            d = d.reshape(xin.shape)
        elif kind == "softmax":
            d = softmax_bwd(yout, d)
    for i in range(n_leaves):
        if grads[i] is None:
            grads[i] = np.zeros_like(params[i])
    return val, grads


def dp_loss_and_grad(spec, params, x, label, n_ranks, loss="ce"):
    """Data-parallel restatement: shard the batch, scale by 1/B_global, SUM."""
    B = x.shape[0]
    assert B % n_ranks == 0
    s = B // n_ranks
    tot, gsum = None, None
    for r in range(n_ranks):
        v, g = net_loss_and_grad(spec, params, x[r * s:(r + 1) * s], label[r * s:(r + 1) * s], loss, global_batch=B)
        tot = v if tot is None else tot + v
        gsum = g if gsum is None else [a + b for a, b in zip(gsum, g)]
    return tot, gsum


# ----------------------------------------------------------------------------------------------
# Model vector space (src/model.jl)
# ----------------------------------------------------------------------------------------------


def vec(params: Sequence[np.ndarray]) -> np.ndarray:
    """src/model.jl:33: vcat of flattened leaves (memory order)."""
    return np.concatenate([p.reshape(-1) for p in params])


def unvec(v: np.ndarray, shapes) -> List[np.ndarray]:
    out, o = [], 0
    for s in shapes:
        n = int(np.prod(s))
        out.append(v[o:o + n].reshape(s).copy()); o += n
    return out


def dot(p1: Sequence[np.ndarray], p2: Sequence[np.ndarray]):
    """src/model.jl:200-202: sum over leaves of BLAS dot, in the leaf eltype."""
    T = p1[0].dtype.type
    acc = T(0)
    for a, b in zip(p1, p2):
        acc = T(acc + T(np.dot(a.reshape(-1), b.reshape(-1))))
    return acc


def _promote(T, scalar):
    """Julia broadcast promotion: a Python float scalar is a Float64 literal, so the
    per-element math runs in Float64 and is rounded to T on store; an int or a numpy scalar of
    type T keeps the math in T."""
    if isinstance(scalar, (float, np.float64)) and not isinstance(scalar, np.float32):
        return np.float64
    return T


def axpy(a, s, g, sign=-1.0):
    """`a .= a .± s .* g` per leaf (model.jl:127-129; GD gradient_descent.jl:16)."""
    T = a.dtype
    P = _promote(T.type, s)
    if sign > 0:
        r = a.astype(P) + P(s) * g.astype(P)
    else:
        r = a.astype(P) - P(s) * g.astype(P)
    return r.astype(T)


# ----------------------------------------------------------------------------------------------
# Optimizers (src/optimizers/*.jl) on a flat vector (== every leaf, elementwise)
# ----------------------------------------------------------------------------------------------


def nesterov_momentum_sequence():
    """src/optimizers/nesterov_method.jl:3-6."""
    t = 1.0
    while True:
        t_next = (1 + math.sqrt(1 + 4 * t * t)) / 2
        yield (t - 1) / t_next
        t = t_next


@dataclass
class GradientDescent:
    stepsize: float

    def init(self, x):
        """gradient_descent.jl:11-13 (state aliases the variable)."""
        self.x = x
        return self

    def step(self, g):
        """gradient_descent.jl:15-17."""
        self.x[...] = axpy(self.x, self.stepsize, g)


@dataclass
class HeavyBallMethod:
    stepsize: float
    momentum: float

    def init(self, x):
        self.x = x
        self.d = np.zeros_like(x)
        return self

    def step(self, g):
        """heavy_ball_method.jl:16-19: d .= mu.*d .- s.*g ; x .+= d."""
        T = self.x.dtype
        P = np.promote_types(_promote(T.type, self.momentum), _promote(T.type, self.stepsize))
        self.d[...] = (P.type(self.momentum) * self.d.astype(P) - P.type(self.stepsize) * g.astype(P)).astype(T)
        self.x[...] = self.x + self.d


@dataclass
class NesterovMethod:
    stepsize: float

    def init(self, x):
        self.x = x
        self.prev = np.zeros_like(x)
        self.seq = nesterov_momentum_sequence()
        return self

    def step(self, g):
        """nesterov_method.jl:33-39."""
        T = self.x.dtype
        self.prev[...] = self.x
        self.x[...] = axpy(self.x, self.stepsize, g)
        mu = next(self.seq)  # Float64
        diff = self.x - self.prev  # in T
        self.x[...] = (self.x.astype(np.float64) + mu * diff.astype(np.float64)).astype(T)


@dataclass
class Adam:
    stepsize: float
    beta1: float = 0.9
    beta2: float = 0.999
    epsilon: float = 1e-8

    def init(self, x):
        """adam.jl:19-30: four zero state copies, it = 1."""
        self.x = x
        self.m = np.zeros_like(x); self.v = np.zeros_like(x)
        self.m_hat = np.zeros_like(x); self.v_hat = np.zeros_like(x)
        self.it = 1
        return self

    def step(self, g):
        """adam.jl:32-41 with Float64 hyper-parameters (per-element Float64 math, stored in T)."""
        T = self.x.dtype
        f8 = np.float64
        b1, b2 = self.beta1, self.beta2
        self.m[...] = (b1 * self.m.astype(f8) + (1 - b1) * g.astype(f8)).astype(T)
        g2 = g * g  # literal_pow in T
        self.v[...] = (b2 * self.v.astype(f8) + (1 - b2) * g2.astype(f8)).astype(T)
        self.m_hat[...] = (self.m.astype(f8) / (1 - b1 ** self.it)).astype(T)
        self.v_hat[...] = (self.v.astype(f8) / (1 - b2 ** self.it)).astype(T)
        den = np.sqrt(self.v_hat).astype(f8) + self.epsilon  # sqrt in T, then + eps in Float64
        self.x[...] = (self.x.astype(f8) - self.stepsize * (self.m_hat.astype(f8) / den)).astype(T)
        self.it += 1
        return self.it


def make_optimizer(name: str, **kw):
    return {"gd": GradientDescent, "heavyball": HeavyBallMethod, "nesterov": NesterovMethod, "adam": Adam}[name](**kw)


# ----------------------------------------------------------------------------------------------
# A whole training step on a flat parameter vector (what bench.py's CPU arm times)
# ----------------------------------------------------------------------------------------------


class TrainStep:
    """forward -> loss -> backward -> optimizer update, params held as one flat vector whose
    leaves are views (the reference's state.variable aliasing)."""

    def __init__(self, spec, params, optimizer, loss="ce"):
        self.spec, self.loss = spec, loss
        self.shapes = [p.shape for p in params]
        self.flat = vec(params).copy()
        self.leaves, o = [], 0
        for s in self.shapes:
            n = int(np.prod(s))
            self.leaves.append(self.flat[o:o + n].reshape(s)); o += n
        self.opt = optimizer.init(self.flat)

    def step(self, x, label):
        val, grads = net_loss_and_grad(self.spec, self.leaves, x, label, self.loss)
        self.opt.step(vec(grads))
        return val


# ----------------------------------------------------------------------------------------------
# bf16 tensor-core mode restatement (what the B200 path computes when math="bf16")
# ----------------------------------------------------------------------------------------------


def bf16_round(a):
    """round-to-nearest-even to bfloat16, returned as float64 (the value the tensor cores see)"""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    return (((u + r) >> 16) << 16).astype(np.uint32).view(np.float32).astype(np.float64)


def mlp_bf16_loss_and_grad(spec, params, x, label, global_batch=None):
    """The same step with the tensor-core mode's number formats made explicit: operands (inputs,
    weights, hidden activations, hidden-activation gradients) rounded to bf16, products accumulated
    exactly (float64 here, fp32 on the device), wide-layer outputs stored in bf16, the N <= 32
    logits layer / loss / dZ / dW / db kept in float.  Only linear/relu/softmax + cross-entropy.
    The un-rounded float64 oracle differs from this by the bf16 operand rounding AND by ReLU gates
    that flip sign under that rounding (a gate flip is an O(1) change of one gradient entry)."""
    bf = bf16_round
    acts = [bf(x)]
    lin = [op for op in spec if op[0] == "linear"]
    pi = 0
    a = acts[0]
    i = 0
    ops = list(spec)
    layer_inputs, relu_after = [], []
    while i < len(ops):
        op = ops[i]
        if op[0] == "linear":
            W, b = bf(params[pi]), params[pi + 1].astype(np.float64)
            z = a @ W + b
            has_relu = i + 1 < len(ops) and ops[i + 1][0] == "relu"
            if has_relu:
                z = np.maximum(z, 0); i += 1
            wide = op[2] > 32
            layer_inputs.append(a); relu_after.append(has_relu)
            a = bf(z) if wide else z
            pi += 2
        elif op[0] == "softmax":
            a = softmax_fwd(a)
        else:
            raise ValueError("mlp_bf16_loss_and_grad supports linear/relu/softmax only")
        i += 1
    p = a
    eps = np.finfo(np.float32).eps
    B = x.shape[0] if global_batch is None else global_batch
    loss = float((-(label * np.log(p + eps)).sum(-1)).sum() / B)
    dp = -label / (B * (p + eps))
    d = softmax_bwd(p, dp)                                   # dZ (float)
    grads = [None] * len(params)
    for li in range(len(lin) - 1, -1, -1):
        W = bf(params[2 * li])
        xin = layer_inputs[li]
        dq = bf(d)                                           # tensor-core operand copy of dY (bf16)
        grads[2 * li] = xin.T @ dq
        grads[2 * li + 1] = d.sum(0)                         # db sums the stored dY (float for the logits layer)
        if li > 0:
            dx = dq @ W.T
            if relu_after[li - 1]:
                dx = np.where(xin > 0, dx, 0.0)
            d = bf(dx)                                       # hidden gradients are stored in bf16
    return loss, grads
