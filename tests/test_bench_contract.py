"""CPU checks of bench.py's host-side pieces: the synthetic batch (MNIST's 8-bit value grid, exactly the Float32 the reference's
loader would build), the algorithmic-bytes model behind the sub-lines' HBM rooflines, and the reference arm's JSON line."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_synthetic_batch_is_on_the_8_bit_grid():
    x, lab, u8 = bench.synth(64, 7, with_u8=True)
    assert x.dtype == np.float32 and u8.dtype == np.uint8 and x.shape == (64, bench.D_IN)
    assert np.array_equal(x, u8.astype(np.float32) / np.float32(255))          # what pg_vec_cast_u8_bf16 reproduces on the device
    assert x.min() >= 0.0 and x.max() <= 1.0
    assert np.array_equal(lab.sum(1), np.ones(64, np.float32)) and set(np.unique(lab)) == {0.0, 1.0}
    x2, lab2 = bench.synth(64, 7)
    assert np.array_equal(x, x2) and np.array_equal(lab, lab2)


def test_step_bytes_model():
    # C1 (784-100-10, Adam 28 B/param): inputs read in forward and by dW; hidden activations written, read, read again (ReLU'
    # mask / next dW), their gradients written and read; logits likewise; parameters read twice + gradient + optimizer bytes
    B = 128
    per_sample = 4 * (784 * 2 + 100 * 5 + 10 * 5)
    params = 784 * 100 + 100 + 100 * 10 + 10
    assert bench.step_bytes(bench.C1_SPEC, B, 28, (784,)) == B * per_sample + 4 * 2 * params + params * (4 + 28)
    # frozen trunk (C5): frozen parameters are only read, their activations carry no gradient
    full = bench.step_bytes(bench.C3_SPEC, B, 16, (1, 28, 28))
    frozen = bench.step_bytes(bench.C5_SPEC, B, 12, (1, 28, 28))
    assert frozen < full
    assert bench.FLOP_PER_SAMPLE == 113_754_112


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--batch", "256"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == os.cpu_count()
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"]
