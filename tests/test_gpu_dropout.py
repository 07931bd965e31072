"""GPU: training-mode dropout (src/layers.jl:109-138).  Julia's Xoshiro stream is
not reproducible, so parity is: mask values in {0, T(1/(1-p))}, keep rate 1-p, the pullback uses
exactly the forward mask, determinism per (seed, step), identity outside training mode — plus
test/test_models.jl:220-242 (gradient of a parameter-free Dropout model must run)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import protograd_oracle as O
from tests.helpers import rel_err


@pytest.fixture(scope="module")
def pg():
    import protograd_b200 as pg
    return pg


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_dropout_mask_statistics_and_identity(pg, T):
    x = np.ones((512, 1000), dtype=T)
    d = pg.Dropout(0.1)
    assert np.array_equal(d(x).numpy(), x)                      # eval mode: identity (layers.jl:119-121)
    pg.seed(7)
    y = pg.within_training_mode(lambda: d(x).numpy())
    vals = np.unique(y)
    assert len(vals) == 2 and vals[0] == 0 and vals[1] == T(1 / (1 - 0.1))
    keep = (y != 0).mean()
    assert abs(keep - 0.9) < 5e-3
    assert abs(y.mean() - 1.0) < 1e-2                            # inverted scaling keeps the mean
    pg.seed(7)
    y2 = pg.within_training_mode(lambda: d(x).numpy())
    assert np.array_equal(y, y2)                                 # same seed, same step -> same mask
    y3 = pg.within_training_mode(lambda: d(x).numpy())
    assert not np.array_equal(y, y3)                             # next step -> fresh mask


def test_dropout_pullback_uses_forward_mask(pg):
    """model = Dropout -> Linear -> relu -> Dropout -> Linear -> softmax (examples/01_mnist_mlp.jl:22-30)."""
    rng = np.random.default_rng(0)
    spec = [("linear", 64, 32), ("relu",), ("linear", 32, 10), ("softmax",)]
    p0 = O.init_params(spec, np.float32, seed=2)
    x = rng.random((256, 64)).astype(np.float32); lab = np.eye(10, dtype=np.float32)[rng.integers(0, 10, 256)]
    m = pg.Compose(pg.Dropout(0.1), pg.Linear(W=p0[0], b=p0[1]), pg.relu, pg.Dropout(0.1), pg.Linear(W=p0[2], b=p0[3]), pg.softmax)

    def run():
        out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(x), lab), m)
        g = pb()
        return float(out), [p.numpy() for p in g.allparams()]
    pg.seed(11)
    loss_t, g_t = pg.within_training_mode(run)
    assert np.isfinite(loss_t) and all(np.isfinite(a).all() for a in g_t)
    # defining property on one layer: the pullback multiplies by exactly the forward mask
    W = rng.standard_normal((64, 16)).astype(np.float32); b = np.zeros(16, dtype=np.float32)
    lin = pg.Linear(W=W, b=b)
    y_t = rng.standard_normal((256, 16)).astype(np.float32)
    pg.seed(5)
    out, pb = pg.within_training_mode(lambda: pg.eval_with_pullback(lambda mm: pg.mse(pg.dropout(mm(x), 0.25), y_t), lin))
    g = pg.within_training_mode(lambda: pb())
    # reconstruct the mask from the forward value itself: y = (xW+b) .* mask  =>  mask = y / (xW+b)
    pg.seed(5)
    yd = pg.within_training_mode(lambda: pg.dropout(lin(x), 0.25).numpy())
    z64 = x.astype(np.float64) @ W.astype(np.float64) + b
    mask = np.where(yd != 0, 1 / 0.75, 0.0)
    assert abs((mask != 0).mean() - 0.75) < 2e-2
    assert rel_err(yd, z64 * mask) < 1e-5
    r = z64 * mask - y_t
    dz = (2.0 / r.size) * r * mask                                # pullback multiplies by the same mask
    assert abs(float(out) - (r * r).mean()) <= 1e-5 * (r * r).mean()
    assert rel_err(g.W.numpy(), x.astype(np.float64).T @ dz) < 1e-5
    assert rel_err(g.b.numpy(), dz.sum(0)) < 1e-5


def test_ref_dropout_model_gradient_runs(pg):
    """test/test_models.jl:220-242: a parameter-free Dropout model; the update on a field-less model is a no-op."""
    class Wrap(pg.Model):
        __fields__ = ("lin",)

        def __init__(self, lin):
            self.lin = lin
    x = np.random.default_rng(1).standard_normal((8, 4)).astype(np.float32)
    d = pg.Dropout(0.1)
    assert len(d.allparams()) == 0 and len(d) == 0
    y = pg.within_training_mode(lambda: d(x).numpy())
    assert y.shape == x.shape
    d.assign(pg.L(d) - 3.0 * pg.L(d))                                # no-op, must not throw (model.jl:131)
