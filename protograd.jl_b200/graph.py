"""CUDA-graph replay of a whole training step (forward + loss + backward + fused optimizer update).

The reference's loop re-runs Zygote's closures every iteration (examples/01_mnist_mlp.jl:44-55); the
eager B200 path re-traces `f(m)` in Python every step, which bounds small-batch steps at ~0.15-0.3 ms of
host time.  `GraphStep` traces and launches the step ONCE under stream capture and then replays the
captured graph: one `cudaGraphLaunch` per step, no Python tracing, no per-kernel launch latency.

Requirements: the objective closes over FIXED device buffers (write each new minibatch into them, e.g.
`x_dev.upload(batch)` or a `Prefetcher` slot), shapes do not change, and no training-mode dropout (its
counter-based seeds are kernel arguments).  Per-step optimizer scalars (Adam's `it`, Nesterov's momentum)
live in device memory and are advanced by a one-thread kernel inside the graph
(`pg_opt_adam_dev` / `pg_opt_nesterov_dev`).
"""
import ctypes

import numpy as np

from . import capi
from .capi import call
from .device import DeviceArray, dtype_code
from .gradient import Loss, eval_with_pullback
from .model import _scalar_flags
from .optimizers import AdamState, GradientDescentState, HeavyBallMethodState, NesterovState, _f64


class GraphStep:
    def __init__(self, f, model, state, math="fp32", warmup=2):
        if state.variable is not model:
            raise ValueError("the optimizer state must have been created for this model (init(optimizer, model))")
        self.f, self.model, self.state, self.math = f, model, state, math
        self.ctx = model.arena()[0]
        self._scalars = None
        if isinstance(state, (AdamState, NesterovState)):
            self._scalars = DeviceArray.empty(self.ctx, (8,), np.float64)
            self._sync_scalars_to_device()
        # warm-up steps allocate plan buffers / scratch (not allowed during capture); they are real
        # steps, so snapshot and restore the parameters and optimizer state around them
        # bf16 mode on a Float32 arena: the captured optimizer pass also writes the bf16 operand copy of the new parameters, so
        # the captured forward contains no cast pass (as `step(..., shadow=True)` does for eager steps)
        self._shadow = False
        if math == "bf16" and isinstance(state, (AdamState, NesterovState, GradientDescentState, HeavyBallMethodState)):
            leaves = model.arena()[3]
            self._shadow = len(leaves) > 0 and leaves[0].dtype == np.float32 and leaves[0].offset == 0
        snap = self._snapshot()
        for _ in range(max(1, warmup)):
            self._one_step()
        self._restore(snap)
        if self._shadow:
            # the restore invalidated the shadow: one forward (no state change) re-casts it, so that the capture below sees a
            # valid copy and records no cast
            eval_with_pullback(self.f, self.model, math=self.math)
        self.ctx.sync()
        # no Python garbage collection while capturing: a finalizer freeing device memory (cudaFree)
        # in the middle of a stream capture would invalidate it
        import gc
        gc_was_enabled = gc.isenabled()
        gc.collect()
        gc.disable()
        call("pg_graph_begin", self.ctx.h)
        try:
            self._one_step()
        finally:
            ge = ctypes.c_void_p()
            try:
                call("pg_graph_end", self.ctx.h, ctypes.byref(ge))
            finally:
                if gc_was_enabled:
                    gc.enable()
        self._graph = ge
        self.steps = 0

    # ---- pieces --------------------------------------------------------------------------------
    def _arenas(self):
        st = self.state
        ms = [st.variable]
        for name in ("m", "v", "_m_hat", "_v_hat", "variable_prev", "step_"):
            a = getattr(st, name, None)
            if a is not None:
                ms.append(a)
        return ms

    def _snapshot(self):
        snap = [a.copy() for a in self._arenas()]
        it = getattr(self.state, "it", None)
        sc = self._scalars.numpy().copy() if self._scalars is not None else None
        return snap, it, sc

    def _restore(self, snap):
        copies, it, sc = snap
        for a, c in zip(self._arenas(), copies):
            _, pa, n, leaves = a.arena()
            _, pc, _, _ = c.arena()
            call("pg_copy_d2d", self.ctx.h, pa, pc, n * leaves[0].dtype.itemsize)
        if it is not None:
            self.state.it = it
        if sc is not None:
            self._scalars.upload(sc)
        for a in self._arenas():
            a._mutated()

    def _sync_scalars_to_device(self):
        s = np.zeros(8)
        s[0] = float(getattr(self.state, "it", 1))
        s[3] = 1.0
        if isinstance(self.state, NesterovState):
            # replay the host-side momentum iterator position: t after k pops
            t = 1.0
            for _ in range(getattr(self.state, "_pops", 0)):
                t = (1 + np.sqrt(1 + 4 * t * t)) / 2
            s[3] = t
        self._scalars.upload(s)

    def _one_step(self):
        out, pb = eval_with_pullback(self.f, self.model, math=self.math)
        if out._plan.has_dropout:
            raise NotImplementedError("GraphStep cannot replay training-mode dropout: the mask's (seed, step) key is a kernel "
                                      "argument baked in at capture, so every replay would reuse one mask; use eager steps")
        g = pb()
        st = self.state
        ctx, code, p, pg_, n = st._arena(g)
        shadow = None
        if self._shadow:
            from .trace import shadow_target
            shadow = shadow_target(self.model)
        if isinstance(st, AdamState):
            o = st.optimizer
            fl = _f64(st.stepsize) | (capi.MATH_FAST if o.fast else 0)
            call("pg_opt_adam_dev", ctx.h, code, p, st._ptr(st.m), st._ptr(st.v),
                 st._ptr(st._m_hat) if st._m_hat is not None else None, st._ptr(st._v_hat) if st._v_hat is not None else None,
                 pg_, n, float(st.stepsize), float(o.beta1), float(o.beta2), float(o.epsilon), self._scalars.ptr, fl, shadow)
        elif isinstance(st, NesterovState):
            call("pg_opt_nesterov_dev", ctx.h, code, p, st._ptr(st.variable_prev), pg_, n, float(st.stepsize), self._scalars.ptr, _f64(st.stepsize), shadow)
        elif isinstance(st, (GradientDescentState, HeavyBallMethodState)):
            st.step(g, shadow)
        else:
            raise TypeError(f"unsupported optimizer state {type(st).__name__}")
        self._loss_buf = out._buf
        self._params_updated()
        return out

    def _params_updated(self):
        """the parameters changed (version bump); a step that also wrote the bf16 shadow stamps it for the new version"""
        self.model._mutated()
        if self._shadow:
            from .trace import shadow_written
            shadow_written(self.model)

    # ---- replay --------------------------------------------------------------------------------
    def __call__(self):
        """run one training step (graph replay); returns the step's Loss handle"""
        if self._shadow:
            # the captured forward contains no cast: if the parameters were changed behind the graph's back (checkpoint load,
            # assign, indexing) the bf16 copy is stale — refresh it before the replay
            from .trace import _shadow_entry
            leaves = self.model.arena()[3]
            ent = _shadow_entry(self.ctx, leaves[0].buf)
            if ent.version != leaves[0].buf.version:
                call("pg_vec_cast_bf16", self.ctx.h, ent.ptr, leaves[0].buf.ptr, leaves[0].buf.nbytes // 4)
                ent.version = leaves[0].buf.version
        call("pg_graph_launch", self.ctx.h, self._graph)
        self.steps += 1
        self._params_updated()
        if isinstance(self.state, AdamState):
            self.state.it += 1
        elif isinstance(self.state, NesterovState):
            self.state.advance(1)        # keep the host-side momentum iterator in step with the device sequence
        k = self.steps
        # every replay writes the captured loss slot: the handle is valid until the next replay
        return Loss(self._loss_buf, lambda: self.steps == k)

    def __del__(self):
        try:
            capi.lib.pg_graph_destroy(self.ctx.h, self._graph)
        except Exception:
            pass


class MegaStep:
    """One training step of a small Float32 MLP in a SINGLE cooperative launch (csrc/mlp_step.cu, pg_mlp_step).

    The reference's own example (examples/01_mnist_mlp.jl: 784-100-10, minibatch 128) is 41 MFLOP per step: as
    eval_with_pullback -> pb() -> step! it is 13 dependent launches and launch latency is all there is.  `MegaStep(f, m,
    state)` traces `f(m)` once, checks that it is Linear -> relu -> ... -> Linear -> softmax -> cross_entropy over FIXED
    Float32 device buffers (write every new minibatch into them, as with GraphStep) and then runs forward, loss,
    backward and the optimizer update (GradientDescent or Adam, same arithmetic as `step`) in one kernel per call.
    Gradients land in `gradient_container(m)` and stay readable; the return value is the step's Loss handle.
    Envelope: 2-4 Linear layers, widths and batch <= 1536, no dropout; anything else raises (use GraphStep)."""

    def __init__(self, f, model, state):
        from . import trace as T
        from .gradient import gradient_container
        if state.variable is not model:
            raise ValueError("the optimizer state must have been created for this model (init(optimizer, model))")
        ctx, _, _, leaves = model.arena()
        self.ctx, self.model, self.state = ctx, model, state
        tr = T.Trace(ctx, leaves, "fp32")
        prev = T.set_current(tr)
        try:
            out = f(model)
        finally:
            T.set_current(prev)

        def need(cond, what):
            if not cond:
                raise NotImplementedError("MegaStep covers Linear -> relu -> ... -> Linear -> softmax -> cross_entropy on Float32 device "
                                          f"buffers: {what}")
        need(isinstance(out, T.Tensor) and out.op == "cross_entropy", "the objective must end in cross_entropy")
        probs, label = out.inputs
        need(probs.op == "softmax", "cross_entropy must take softmax probabilities")
        need(label.op == "input" and isinstance(label.attrs["src"], DeviceArray) and label.dtype == np.float32, "labels must be a Float32 DeviceArray")
        cur, layers = probs.inputs[0], []
        while True:
            need(cur.op == "linear", f"unexpected op {cur.op}")
            x, Wn, bn = cur.inputs
            need(Wn.attrs["grad_index"] is not None and bn.attrs["grad_index"] is not None, "every Linear must be trainable")
            layers.append((Wn.attrs["leaf"], bn.attrs["leaf"]))
            if x.op == "relu":
                cur = x.inputs[0]
            else:
                need(x.op == "input" and isinstance(x.attrs["src"], DeviceArray) and x.dtype == np.float32 and len(x.shape) == 2,
                     "the input must be a 2-D Float32 DeviceArray")
                xin = x
                break
        layers.reverse()
        flat = [p for Wb in layers for p in Wb]
        need(2 <= len(layers) <= 4, "2 to 4 Linear layers")
        need(len(flat) == len(leaves) and all(a is b for a, b in zip(flat, leaves)), "the Linear layers must be exactly the model's leaves, in order")
        need(all(p.dtype == np.float32 for p in flat), "Float32 parameters")
        B = xin.shape[0]
        dims = [layers[0][0].shape[0]] + [W.shape[1] for W, _ in layers]
        need(xin.shape[1] == dims[0] and label.shape == (B, dims[-1]), "shape mismatch")
        need(max(dims + [B]) <= 1536, "widths and batch <= 1536")
        self._x, self._y = xin.attrs["src"], label.attrs["src"]          # keep the buffers alive
        self._grad = gradient_container(model)
        gl = self._grad.allparams()
        d = capi.MlpDesc()
        d.n_layers, d.batch = len(layers), B
        for i, v in enumerate(dims):
            d.dims[i] = v
        d.x, d.labels = self._x.ptr, self._y.ptr
        for l, (W, b) in enumerate(layers):
            d.W[l], d.b[l], d.gW[l], d.gb[l] = W.ptr, b.ptr, gl[2 * l].ptr, gl[2 * l + 1].ptr
        gb = out.attrs.get("global_batch")
        d.loss_scale = 1.0 / float(gb if gb else B)
        st = state
        _, code, p, pg_, n = st._arena(self._grad)
        d.arena_p, d.arena_g, d.arena_n = p, pg_, n
        d.stepsize = float(st.stepsize)
        if isinstance(st, AdamState):
            o = st.optimizer
            d.opt = 2
            d.arena_m, d.arena_v = st._ptr(st.m), st._ptr(st.v)
            d.arena_m_hat = st._ptr(st._m_hat) if st._m_hat is not None else None
            d.arena_v_hat = st._ptr(st._v_hat) if st._v_hat is not None else None
            d.beta1, d.beta2, d.eps = float(o.beta1), float(o.beta2), float(o.epsilon)
            d.flags = _f64(st.stepsize) | (capi.MATH_FAST if o.fast else 0)
        elif isinstance(st, GradientDescentState):
            d.opt, d.flags = 1, _f64(st.stepsize)
        else:
            raise NotImplementedError(f"MegaStep updates with GradientDescent or Adam, not {type(st).__name__}")
        self._desc = d
        self._loss = DeviceArray.empty(ctx, (64,), np.float32)
        self.steps = 0

    def __call__(self):
        from .trace import _Borrowed
        d, st = self._desc, self.state
        k = self.steps % 64
        d.loss_out = self._loss.ptr + 4 * k
        if d.opt == 2:
            d.it = st.it
        call("pg_mlp_step", self.ctx.h, ctypes.byref(d))
        if d.opt == 2:
            st.it += 1
        self.steps += 1
        self.model._mutated()
        self._grad._mutated()
        n = self.steps
        slot = DeviceArray(_Borrowed(self.ctx, d.loss_out), 0, (1,), np.float32)
        return Loss(slot, lambda: self.steps - n < 64)
