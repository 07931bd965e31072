"""Optimizers — host mirror of src/optimizers/*.jl: `init(optimizer, variable)` / `step(state, grad)`
(`step!` in Julia).  `state.variable` aliases the user's model (gradient_descent.jl:11-13), and each
step is ONE fused kernel over the parameter arena instead of 1-5 broadcast passes per leaf.
Hyper-parameters given as Python floats are Float64 (per-element Float64 math);
pass numpy float32 scalars to mirror `Float32` hyper-parameters.  fast=True selects the
all-Float32 fused Adam (normwise ~1e-7 from the exact form).
"""
import math
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .capi import call
from .device import dtype_code
from .model import Model, _scalar_flags


def _f64(x):
    return capi.SCALAR_F64 if _scalar_flags(x) else 0


class NesterovMomentumSequence:
    """src/optimizers/nesterov_method.jl:1-8."""

    def __iter__(self):
        t = 1.0
        while True:
            t_next = (1 + math.sqrt(1 + 4 * t * t)) / 2
            yield (t - 1) / t_next
            t = t_next


@dataclass
class GradientDescent:
    stepsize: float


@dataclass
class HeavyBallMethod:
    stepsize: float
    momentum: float


@dataclass
class NesterovMethod:
    stepsize: float
    momentum_sequence: object = field(default_factory=NesterovMomentumSequence)


@dataclass
class Adam:
    stepsize: float
    beta1: float = 0.9
    beta2: float = 0.999
    epsilon: float = 1e-8
    fast: bool = False
    keep_hats: bool = False  # True: rewrite m_hat / v_hat every step; False: materialise them on access


class _State:
    """`_range = (start, count)` (padded-arena elements) restricts a step to a slice of the arena:
    the update is elementwise, so a data-parallel caller can update each gradient bucket as soon as
    its all-reduce has landed (see step_range)."""
    _range = None

    def _arena(self, g):
        v = self.variable
        v._check_same(g)
        ctx, p, n, leaves = v.arena()
        _, pg, _, _ = g.arena()
        self._isz = leaves[0].dtype.itemsize
        if self._range is not None:
            start, count = self._range
            p += start * self._isz; pg += start * self._isz; n = count
        return ctx, dtype_code(leaves[0].dtype), p, pg, n

    def _ptr(self, m):
        p = m.arena()[1]
        if self._range is not None:
            p += self._range[0] * self._isz
        return p

    def _updated(self, shadow):
        """parameters changed: bump the arena version; a whole-arena step that also wrote the bf16 shadow stamps it"""
        self.variable._mutated()
        if shadow and self._range is None:
            from .trace import shadow_written
            shadow_written(self.variable)


class GradientDescentState(_State):
    def __init__(self, opt, variable):
        self.optimizer, self.stepsize, self.variable = opt, opt.stepsize, variable

    def step(self, g, shadow=None):
        """gradient_descent.jl:15-17"""
        ctx, code, p, pg, n = self._arena(g)
        call("pg_opt_gd", ctx.h, code, p, pg, n, float(self.stepsize), _f64(self.stepsize), shadow)
        self._updated(shadow)
        return self.variable


class HeavyBallMethodState(_State):
    def __init__(self, opt, variable):
        self.optimizer, self.stepsize, self.variable = opt, opt.stepsize, variable
        self.step_ = variable.zero()      # `step` field of the reference (heavy_ball_method.jl:13-14)

    def step(self, g, shadow=None):
        """heavy_ball_method.jl:16-19"""
        ctx, code, p, pg, n = self._arena(g)
        fl = _f64(self.stepsize) | (capi.SCALAR2_F64 if _scalar_flags(self.optimizer.momentum) else 0)
        call("pg_opt_heavyball", ctx.h, code, p, self._ptr(self.step_), pg, n, float(self.stepsize), float(self.optimizer.momentum), fl, shadow)
        self._updated(shadow)
        return self.variable


class NesterovState(_State):
    def __init__(self, opt, variable):
        self.optimizer, self.stepsize, self.variable = opt, opt.stepsize, variable
        self.variable_prev = variable.zero()
        self.momentum_iterator = iter(opt.momentum_sequence)
        self._pops = 0            # momenta taken from the sequence so far (checkpoints and graph replay restore it)

    def pop_momentum(self):
        self._pops += 1
        return next(self.momentum_iterator)

    def advance(self, k):
        """skip k momenta (graph replays and checkpoint restore keep the host iterator in step with the device)"""
        for _ in range(int(k)):
            self.pop_momentum()

    def step(self, g, shadow=None):
        """nesterov_method.jl:33-39 (momentum popped from the host-side stateful iterator)"""
        ctx, code, p, pg, n = self._arena(g)
        mu = self.pop_momentum()
        call("pg_opt_nesterov", ctx.h, code, p, self._ptr(self.variable_prev), pg, n, float(self.stepsize), float(mu), _f64(self.stepsize), shadow)
        self._updated(shadow)
        return self.variable


class AdamState(_State):
    """adam.jl:8-17.  `m_hat` / `v_hat` are state fields in the reference; they are pure functions
    of (m, v, it), so by default they are NOT rewritten every step (that would be 8 of 36 B/param of
    HBM traffic) but materialised on access with the same arithmetic (m ./ (1 - beta1^it) with the
    `it` of the last step).  Adam(keep_hats=True) restores the eager per-step materialisation."""

    def __init__(self, opt, variable):
        self.optimizer, self.stepsize, self.variable = opt, opt.stepsize, variable
        self.m, self.v = variable.zero(), variable.zero()
        self._m_hat = variable.zero() if opt.keep_hats else None
        self._v_hat = variable.zero() if opt.keep_hats else None
        self.it = 1                                               # adam.jl:28

    def _hat(self, src, beta):
        if self.it <= 1:
            return self.variable.zero()                           # init value (adam.jl:26-27)
        c = 1 - beta ** (self.it - 1)                             # the `it` the last step used
        return src / (float(c) if _scalar_flags(self.stepsize) else np.float32(c))

    @property
    def m_hat(self):
        return self._m_hat if self._m_hat is not None else self._hat(self.m, self.optimizer.beta1)

    @property
    def v_hat(self):
        return self._v_hat if self._v_hat is not None else self._hat(self.v, self.optimizer.beta2)

    def step(self, g, shadow=None):
        """adam.jl:32-41; returns the incremented iteration counter like the reference"""
        ctx, code, p, pg, n = self._arena(g)
        o = self.optimizer
        fl = _f64(self.stepsize) | (capi.MATH_FAST if o.fast else 0)
        call("pg_opt_adam", ctx.h, code, p, self._ptr(self.m), self._ptr(self.v),
             self._ptr(self._m_hat) if self._m_hat is not None else None, self._ptr(self._v_hat) if self._v_hat is not None else None,
             pg, n, float(self.stepsize), float(o.beta1), float(o.beta2), float(o.epsilon), self.it, fl, shadow)
        self._updated(shadow)
        self.it += 1
        return self.it


_STATES = {GradientDescent: GradientDescentState, HeavyBallMethod: HeavyBallMethodState, NesterovMethod: NesterovState, Adam: AdamState}


def init(optimizer, variable):
    """init(optimizer, variable) of src/optimizers/*.jl — the state aliases `variable`."""
    if not isinstance(variable, Model):
        raise TypeError("the B200 optimizers update arena-backed Models")
    return _STATES[type(optimizer)](optimizer, variable)


def step_range(state, gradient, start, count, shadow=None, last=True):
    """step! restricted to arena elements [start, start+count).  Call once per bucket; pass
    last=False for all but the final bucket of a step so per-step scalars (Adam's `it`, Nesterov's
    momentum) advance once per step."""
    if shadow is True:
        from .trace import shadow_target
        shadow = shadow_target(state.variable) + start * 2
    state._range = (int(start), int(count))
    try:
        if isinstance(state, AdamState) and not last:
            it = state.it
            r = state.step(gradient, shadow or None)
            state.it = it
        elif isinstance(state, NesterovState):
            if not hasattr(state, "_mu_cached"):
                state._mu_cached = state.pop_momentum()
            mu = state._mu_cached
            ctx, code, p, pg, n = state._arena(gradient)
            call("pg_opt_nesterov", ctx.h, code, p, state._ptr(state.variable_prev), pg, n, float(state.stepsize), float(mu), _f64(state.stepsize), shadow or None)
            state.variable._mutated()
            if last:
                del state._mu_cached
            r = state.variable
        else:
            r = state.step(gradient, shadow or None)
    finally:
        state._range = None
    if shadow and last:
        # every bucket of this step wrote its slice of the shadow: the whole copy matches the new parameters
        from .trace import shadow_written
        shadow_written(state.variable)
    return r


def step(state, gradient, shadow=None):
    """step!(state, gradient).  shadow=True: the fused pass also writes the bf16 operand copy of the
    updated parameters that the next bf16-mode forward will use (saves that forward's cast pass)."""
    if shadow is True:
        from .trace import shadow_target
        arena_leaves = state.variable.arena()[3]
        full = len(arena_leaves) and arena_leaves[0].offset == 0
        shadow = shadow_target(state.variable) if full else None
    return state.step(gradient, shadow or None)
