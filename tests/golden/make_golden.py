"""Generate the committed golden fixtures from the CPU oracle (fixed seeds).

    python -m tests.golden.make_golden        # rewrites tests/golden/*.npz

No output of the real reference (Julia + Zygote + NNlib) can be produced in this image (no
Julia), so these vectors are ORACLE-generated: they pin the oracle against drift and give the
GPU parity tests fixed inputs/outputs.  Large leaves are stored as (norm, strided sample).
Configs follow SURVEY.md §8 / BASELINE.json: C1 MLP, C2 README least squares (Float64),
C3 conv net, C5 frozen trunk + trainable head.
"""
import os

import numpy as np

from oracle import protograd_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))

C1_SPEC = [("linear", 784, 100), ("relu",), ("linear", 100, 10), ("softmax",)]
C3_SPEC = [("conv", 1, 6, 5), ("relu",), ("maxpool", 2), ("conv", 6, 16, 5), ("relu",), ("maxpool", 2),
           ("flatten",), ("linear", 256, 120), ("relu",), ("linear", 120, 84), ("relu",), ("linear", 84, 10), ("softmax",)]
C5_SPEC = [("conv", 1, 6, 5, False), ("relu",), ("maxpool", 2), ("conv", 6, 16, 5, False), ("relu",), ("maxpool", 2),
           ("flatten",), ("linear", 256, 10, True), ("softmax",)]
STRIDE = 97


def summ(d, key, arr):
    a = np.asarray(arr)
    d[key + "_norm"] = np.array(np.linalg.norm(a.astype(np.float64).ravel()))
    d[key + "_sample"] = a.ravel()[::STRIDE].copy() if a.size > 4096 else a.copy()


def class_data(spec, B, T, seed):
    rng = np.random.default_rng(seed)
    shape = (B, spec[0][1], 28, 28) if spec[0][0] == "conv" else (B, spec[0][1])
    x = rng.random(shape).astype(T)
    lab = np.eye(10, dtype=T)[rng.integers(0, 10, B)]
    return x, lab


def biased_params(spec, T, seed):
    ps = O.init_params(spec, T, seed)
    rng = np.random.default_rng(seed + 100)
    for i in range(1, len(ps), 2):
        ps[i] = (0.05 * rng.standard_normal(ps[i].shape)).astype(T)
    return ps


def _train_case(spec, B, opt_name, opt_kw, steps, seed, T=np.float32):
    x, lab = class_data(spec, B, T, seed)
    p0 = biased_params(spec, T, seed)
    # raw inputs and initial leaves (C order; Julia reads them with reversed dims): julia/parity.jl rebuilds the exact case
    out = {"x": x, "label": lab}
    for i, p in enumerate(p0):
        out[f"p0_{i}"] = p
    v, g = O.net_loss_and_grad(spec, p0, x, lab, "ce")
    v64, g64 = O.net_loss_and_grad(spec, [p.astype(np.float64) for p in p0], x.astype(np.float64), lab.astype(np.float64), "ce")
    out["loss"] = np.array(v); out["loss_f64"] = np.array(v64)
    for i, (a, b) in enumerate(zip(g, g64)):
        summ(out, f"grad{i}", a); summ(out, f"grad64_{i}", b)
    ts = O.TrainStep(spec, p0, O.make_optimizer(opt_name, **opt_kw), "ce")
    losses = [ts.step(x, lab) for _ in range(steps)]
    out["losses"] = np.array(losses, dtype=T)
    for i, leaf in enumerate(ts.leaves):
        summ(out, f"param{i}", leaf)
    return out


def c1_mlp():
    """examples/01_mnist_mlp.jl:19-40 shapes; Adam per BASELINE.json, 3 steps, B=16."""
    return _train_case(C1_SPEC, 16, "adam", dict(stepsize=1e-3), 3, seed=11)


def c3_conv():
    """examples/02_mnist_conv.jl:25-54 shapes; Nesterov per BASELINE.json, 3 steps, B=8."""
    return _train_case(C3_SPEC, 8, "nesterov", dict(stepsize=1e-2), 3, seed=13)


def c5_finetune():
    """test/test_fine_tuning.jl pattern scaled: frozen conv trunk + Linear head, GD 0.1."""
    return _train_case(C5_SPEC, 8, "gd", dict(stepsize=0.1), 3, seed=17)


def c2_inputs():
    rng = np.random.default_rng(7)
    A0 = rng.standard_normal((5, 3)); b0 = rng.standard_normal(3)      # C layout W[in][out]
    x = rng.standard_normal((300, 5))
    y = x @ A0 + b0 + rng.standard_normal((300, 3))
    A = rng.standard_normal((5, 3)); b = rng.standard_normal(3)
    return x, y, A, b


def c2_readme():
    """README.md:27-32,86,96-99,143-150,177-183: Float64 LinearModel 5->3, 300 samples,
    loss = sum((Ax+b-y)^2)/300, 100 manual steps `m .= m .- 0.1 .* grad`, then 100 Adam(0.1)."""
    x, y, A, b = c2_inputs()
    spec = [("linear", 5, 3)]
    out = {"x": x, "y": y, "A": A, "b": b}
    p = [A.copy(), b.copy()]
    v, g = O.net_loss_and_grad(spec, p, x, y, "mse", mse_denom=300)
    out["loss0"] = np.array(v); out["gradA"] = g[0]; out["gradb"] = g[1]
    out["dot_m_msum"] = np.array(O.dot(p, [a + a for a in p]))     # README.md:72
    for _ in range(100):
        v, g = O.net_loss_and_grad(spec, p, x, y, "mse", mse_denom=300)
        p = [O.axpy(a, 0.1, gi) for a, gi in zip(p, g)]
    out["loss_gd100"] = np.array(O.net_loss_and_grad(spec, p, x, y, "mse", mse_denom=300)[0])
    out["A_gd100"] = p[0]; out["b_gd100"] = p[1]
    flat = O.vec([A, b]).copy()
    leaves = [flat[:15].reshape(5, 3), flat[15:]]
    st = O.Adam(stepsize=1e-1).init(flat)
    for _ in range(100):
        v, g = O.net_loss_and_grad(spec, leaves, x, y, "mse", mse_denom=300)
        st.step(O.vec(g))
    out["loss_adam100"] = np.array(O.net_loss_and_grad(spec, leaves, x, y, "mse", mse_denom=300)[0])
    out["vec_adam100"] = flat.copy()
    return out


CASES = {"c1_mlp": c1_mlp, "c2_readme": c2_readme, "c3_conv": c3_conv, "c5_finetune": c5_finetune}

if __name__ == "__main__":
    for name, fn in CASES.items():
        d = fn()
        np.savez(os.path.join(HERE, name + ".npz"), **d)
        print(name, {k: v.shape for k, v in list(d.items())[:4]}, "...")
