"""Shared helpers for the GPU parity tests: build a device model from an oracle-style spec."""
import numpy as np


def build_model(pg, spec, params, frozen_as_closure=False):
    """spec as in oracle.protograd_oracle (list of tuples); params: numpy leaves in vec(m) order."""
    layers, pi = [], 0
    for op in spec:
        k = op[0]
        if k == "linear":
            layers.append(pg.Linear(W=params[pi], b=params[pi + 1])); pi += 2
        elif k == "conv":
            layers.append(pg.Conv(w=params[pi], b=params[pi + 1])); pi += 2
        elif k == "relu":
            layers.append(pg.relu)                       # closure form, as in the examples
        elif k == "maxpool":
            layers.append(lambda x, s=op[1]: pg.maxpool(x, (s, s)))
        elif k == "meanpool":
            layers.append(pg.MeanPool((op[1], op[1])))
        elif k == "flatten":
            layers.append(pg.flatten)
        elif k == "softmax":
            layers.append(pg.softmax)
        else:
            raise ValueError(k)
    return pg.Compose(*layers)


def leaves_numpy(m):
    return [p.numpy() for p in m.allparams()]


def rel_err(a, ref):
    a = np.asarray(a, dtype=np.float64).ravel()
    ref = np.asarray(ref, dtype=np.float64).ravel()
    d = np.linalg.norm(ref)
    return np.linalg.norm(a - ref) / (d if d > 0 else 1.0)


class VecModel:
    pass
