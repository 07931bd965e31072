#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: train samples/sec per step (device-timed) on B200.

Workload (config C4 of SURVEY.md §8 / BASELINE.json configs[3]): wide MLP 784-4096-4096-10,
softmax + cross-entropy, Adam, 8,192 samples per GPU (global batch 65,536 at 8 GPUs -> weak scaling),
synthetic U[0,1) inputs and random one-hot labels, xavier-uniform weights.  A step is
forward -> loss -> backward -> (N>1: NCCL all-reduce of the flat gradient arena) -> fused Adam,
run through the package's public API (eval_with_pullback / pb() / step) in bf16 tensor-core mode
(bf16 operands, fp32 accumulation, fp32 master weights and optimizer state).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Besides the contract keys the JSON line carries
  parity         loss of the first step against the CPU oracle on the same inputs and, at N > 1, the all-reduced
                 gradient of a small global batch against the single-rank full-batch gradient plus the replicas'
                 arena checksums after one update (asserted BEFORE anything is timed);
  roofline       the five wide tcgen05 GEMM launches of a step, event-timed one by one on REAL operands (the
                 step's own activations / gradients / weights), against the measured burst peak;
  other_configs  C1 / C3 / C5 (BASELINE configs[0], [2], [4]) device-timed steps with their HBM roofline (N = 1);
  cpu_baseline   the oracle port of the same step on the host cores at the SAME 8,192-sample batch.

--impl reference times the CPU oracle (numpy/OpenBLAS restatement of the reference step; Julia is
not installed) on 8,192-sample steps of the same workload on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D_IN, D_H, D_OUT = 784, 4096, 10
PER_GPU_BATCH = 8192
SPEC = [("linear", D_IN, D_H), ("relu",), ("linear", D_H, D_H), ("relu",), ("linear", D_H, D_OUT), ("softmax",)]
# algorithmic flop/sample: fwd + dW for all layers + dX for all but the first
FLOP_PER_SAMPLE = 2 * (2 * (D_IN * D_H + D_H * D_H + D_H * D_OUT) + (D_H * D_H + D_H * D_OUT))
C1_SPEC = [("linear", 784, 100), ("relu",), ("linear", 100, 10), ("softmax",)]
C3_SPEC = [("conv", 1, 6, 5), ("relu",), ("maxpool", 2), ("conv", 6, 16, 5), ("relu",), ("maxpool", 2),
           ("flatten",), ("linear", 256, 120), ("relu",), ("linear", 120, 84), ("relu",), ("linear", 84, 10), ("softmax",)]
C5_SPEC = [("conv", 1, 6, 5, False), ("relu",), ("maxpool", 2), ("conv", 6, 16, 5, False), ("relu",), ("maxpool", 2),
           ("flatten",), ("linear", 256, 10, True), ("softmax",)]
FLOPS_OTHER = {"C1": 319_600, "C3": 1_517_040, "C5": 490_240}


def synth(B, seed, with_u8=False):
    """MNIST-shaped synthetic batch: uniform 8-bit pixels scaled to Float32 in [0, 1] exactly like the reference's loader
    (MLDatasets yields Float32(pixel / 255), examples/01_mnist_mlp.jl:9-12) and random one-hot labels"""
    rng = np.random.default_rng(seed)
    u8 = rng.integers(0, 256, (B, D_IN), dtype=np.uint8)
    x = u8.astype(np.float32) / np.float32(255)
    lab = np.zeros((B, D_OUT), dtype=np.float32)
    lab[np.arange(B), rng.integers(0, D_OUT, B)] = 1.0
    return (x, lab, u8) if with_u8 else (x, lab)


def init_params(seed=0):
    rng = np.random.default_rng(seed)
    ps = []
    for (a, b) in ((D_IN, D_H), (D_H, D_H), (D_H, D_OUT)):
        s = np.float32(np.sqrt(6.0 / (a + b)))
        ps.append((s * (2 * rng.random((a, b), dtype=np.float32) - 1)).astype(np.float32))
        ps.append(np.zeros(b, dtype=np.float32))
    return ps


def workload_config(B, world, math, dp_mode, exact_adam):
    return {"workload": "wide MLP 784-4096-4096-10 (BASELINE configs[3]), softmax+cross-entropy, Adam(1e-3), %d samples/GPU" % B,
            "global_batch": B * world, "parallelism": "dp%d" % world, "dp_mode": dp_mode if world > 1 else None, "math": math,
            "adam": "float64-promoted (bit-exact)" if exact_adam else "fused float32 (normwise 1e-7 of the reference form)",
            "inputs": "synthetic: uniform 8-bit pixels / 255 as Float32 (MNIST's own value grid), random one-hot labels",
            "l2": "working set per step (>600 MB of activations, weights, gradients, optimizer state) exceeds the 126 MB L2"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference step) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_port_throughput(steps, warmup, batch=PER_GPU_BATCH, budget_s=150.0):
    """(samples/s, s/step, timed steps actually run): `steps` is cut so the leg fits `budget_s`"""
    from oracle import protograd_oracle as O
    x, lab = synth(batch, 123)
    ts = O.TrainStep(SPEC, init_params(0), O.Adam(stepsize=1e-3), "ce")
    per = None
    for _ in range(max(1, warmup)):
        t0 = time.perf_counter()
        ts.step(x, lab)
        per = time.perf_counter() - t0
    steps = max(1, min(steps, int(budget_s / max(per, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        ts.step(x, lab)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    warm = max(1, min(args.warmup, 5))
    v, per, steps = cpu_port_throughput(args.steps, warm, args.batch)
    cores = os.cpu_count()
    sample = (f"{args.batch}-sample steps (one GPU's shard of the workload) of the same 784-4096-4096-10 Adam step, {steps} timed steps "
              f"after {warm} warm-up, numpy/OpenBLAS float32 oracle (CPU restatement of the reference step, not Julia: no Julia toolchain in "
              f"this image) on all host cores; throughput is per host, independent of N")
    cfg = workload_config(args.batch, world, args.math, args.dp_mode, args.exact_adam)
    line = {
        "impl": "reference", "metric": "train samples/sec per step", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# algorithmic HBM bytes of one training step of a layer spec (the convention DESIGN.md §3 states): every tensor
# of the reference's graph touches HBM once per use — inputs and activations: written once (by their producer,
# fused with bias / ReLU), read once by the next layer, read once more in the backward pass; activation gradients:
# written once, read once; parameters: read in forward and backward, gradient written once, plus the optimizer's
# bytes per parameter (GD 12, Nesterov 16, Adam 28).
# ------------------------------------------------------------------------------------------------
def step_bytes(spec, B, opt_bytes, in_shape, isz=4):
    shape = in_shape
    size = [int(np.prod(shape))]       # tensor i feeds op i -> tensor i+1 (ReLU / flatten / softmax make no tensor)
    has_grad, bwd_read = [False], [False]
    params_t = params_f = 0
    trainable_seen = False
    for op in spec:
        k = op[0]
        if k == "linear":
            tr = op[3] if len(op) > 3 else True
            n = op[1] * op[2] + op[2]
            shape = (op[2],)
        elif k == "conv":
            tr = op[4] if len(op) > 4 else True
            n = op[2] * op[1] * op[3] * op[3] + op[2]
            shape = (op[2], shape[1] - op[3] + 1, shape[2] - op[3] + 1)
        elif k in ("maxpool", "meanpool"):
            tr, n = False, 0
            shape = (shape[0], shape[1] // op[1], shape[2] // op[1])
        elif k == "flatten":
            shape = (int(np.prod(shape)),)
            continue
        elif k in ("relu", "softmax"):
            continue
        else:
            raise ValueError(k)
        if tr:
            params_t += n
            bwd_read[-1] = True            # dW needs the layer's input
            trainable_seen = True
        else:
            params_f += n
        size.append(int(np.prod(shape)))
        has_grad.append(trainable_seen)
        bwd_read.append(trainable_seen)    # ReLU' mask / pool routing / next layer's dW
    per_sample = 0
    for i, n_el in enumerate(size):
        per_sample += n_el * isz * ((1 if i > 0 else 0) + 1 + (1 if bwd_read[i] else 0) + (2 if has_grad[i] else 0))
    return B * per_sample + isz * 2 * (params_t + params_f) + params_t * (4 + opt_bytes)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="samples per GPU")
    ap.add_argument("--math", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg and the oracle loss check")
    ap.add_argument("--no-others", action="store_true", help="skip the C1/C3/C5 sub-lines")
    ap.add_argument("--dp-mode", default="bucket", choices=["bucket", "flat", "none"], help="gradient exchange: layer-bucketed overlapped all-reduce (default), one flat all-reduce, or none (diagnostic)")
    ap.add_argument("--sm-limit", type=int, default=0, help="size persistent grids for this many SMs (0 = all)")
    ap.add_argument("--exact-adam", action="store_true", help="Float64-promoted Adam arithmetic (bit-exact with the reference's broadcast) instead of the fused Float32 form")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = None
    if world > 1 and os.environ.get("PG_BENCH_NO_AFFINITY") is None:
        # one process per GPU: run on the cores next to this rank's GPU, so that the pinned input buffers (first-touch) live in
        # the NUMA node its PCIe root hangs off — with eight ranks uploading at once, remote-socket reads were the e2e bound
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
            numa = sorted(os.sched_getaffinity(0))
        except Exception as e:          # no NVML / restricted container: keep the default placement
            numa = "unavailable: " + repr(e)[:80]
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import protograd_b200 as pg
    from protograd_b200 import dp
    from protograd_b200.capi import call
    from protograd_b200.device import f32_to_bf16

    ctx = pg.Context.get(local)
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    ctx.use_stream(stream.cuda_stream)

    B = args.batch
    GB = B * world
    p0 = init_params(0)
    model = pg.Compose(pg.Linear(W=p0[0], b=p0[1]), pg.relu, pg.Linear(W=p0[2], b=p0[3]), pg.relu, pg.Linear(W=p0[4], b=p0[5]), pg.softmax)
    state = pg.init(pg.Adam(stepsize=1e-3, fast=not args.exact_adam), model)
    if world > 1:
        dp.broadcast_parameters(model, dp=dp.init(ctx))
    x_h, lab_h, u8_h = synth(B, 1000 + rank, with_u8=True)
    x_d, lab_d = pg.DeviceArray.from_numpy(ctx, x_h), pg.DeviceArray.from_numpy(ctx, lab_h)
    # pinned host copies for the end-to-end leg: the minibatch as the host holds it for this path — the 8-bit pixels the
    # dataset is stored in (converted on the device by pg_vec_cast_u8_bf16 to exactly the bf16 values the first GEMM would
    # round the reference's Float32(pixel / 255) array to) and int32 class indices instead of a dense one-hot matrix
    # (pg_softmax_ce_idx_fwd_bwd): 6.5 MB per step instead of 26 MB — at 8 ranks the uploads share the host's PCIe uplinks
    # and 12.9 MB of bf16 pixels per rank and step made the end-to-end leg upload-bound (1.04 vs 0.86 ms)
    if args.math == "bf16":
        cls_h = np.argmax(lab_h, axis=1).astype(np.int32)
        pin_x, pin_l = pg.PinnedBuffer(u8_h.nbytes), pg.PinnedBuffer(cls_h.nbytes)
        x_p, lab_p = pin_x.as_numpy(np.uint8, u8_h.shape), pin_l.as_numpy(np.int32, cls_h.shape)
        x_p[...] = u8_h; lab_p[...] = cls_h
        e2e_specs = [(u8_h.shape, np.uint8), (cls_h.shape, np.int32)]
        xb_dev = pg.DeviceArray.empty(ctx, x_h.shape, pg.BF16)
    else:
        pin_x, pin_l = pg.PinnedBuffer(x_h.nbytes), pg.PinnedBuffer(lab_h.nbytes)
        x_p, lab_p = pin_x.as_numpy(np.float32, x_h.shape), pin_l.as_numpy(np.float32, lab_h.shape)
        x_p[...] = x_h; lab_p[...] = lab_h
        e2e_specs = [(x_h.shape, np.float32), (lab_h.shape, np.float32)]

    dpc = dp.init(ctx) if world > 1 else None        # the library's own NCCL communicator (csrc/dp.cu), id via torch.distributed
    # fused=(model, state): gradient exchange + Adam update in ONE kernel per bucket over NVLink peer memory (csrc/dp.cu);
    # falls back to ncclAllReduce buckets + the fused Adam pass when peer memory cannot be mapped (or PG_DP_FUSED=0)
    buckets = dp.GradientBuckets(pg.gradient_container(model), bucket_bytes=int(os.environ.get("PG_BUCKET_MB", "24")) << 20, dp=dpc,
                                 overlap_update=os.environ.get("PG_DP_OVERLAP_UPDATE", "1") == "1",
                                 fused=(model, state) if args.dp_mode == "bucket" else None) if world > 1 else None
    if args.sm_limit:
        ctx.set_sm_limit(args.sm_limit)
    use_shadow = args.math == "bf16"
    def train_step(x, lab, gb=GB):
        out, pb = pg.eval_with_pullback(lambda m: pg.cross_entropy(m(x), lab, global_batch=gb), model, math=args.math)
        if buckets is None:
            pg.step(state, pb(), shadow=use_shadow)
        elif args.dp_mode == "bucket":
            g = pb(on_ready=buckets.on_ready())     # tail buckets reduce while the remaining backward GEMMs run
            buckets.finish_and_step(state, g, shadow=use_shadow)
        elif args.dp_mode == "flat":
            g = pb()
            dp.allreduce_gradients(g, dpc)
            pg.step(state, g, shadow=use_shadow)
        else:                                   # "none": no exchange (diagnostic only, replicas diverge)
            pg.step(state, pb(), shadow=use_shadow)
        return out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_enqueue = [0.0]

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        host_enqueue[0] = (time.perf_counter() - t0) / steps * 1e3
        if dpc is not None:
            dpc.wait()          # the last step's tail-bucket exchange is part of the timed region (its wait is deferred otherwise)
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- parity, asserted before anything is timed -----------------------------------------------------------
    parity = {}
    if world > 1:
        # (1) a small global batch from identical seeds: every rank draws the same global batch and takes its
        #     shard; the bucketed all-reduce must reproduce the single-rank full-batch gradient (fp32 summation
        #     order is the only difference) ...
        PB = 256
        xg, lg = synth(PB * world, 77)
        xs, ls = xg[rank * PB:(rank + 1) * PB], lg[rank * PB:(rank + 1) * PB]
        out, pb = pg.eval_with_pullback(lambda m: pg.cross_entropy(m(xs), ls, global_batch=PB * world), model, math=args.math)
        g = pb(on_ready=buckets.on_ready(reduce_only=True))
        buckets.finish()
        g_dp = g.vec().astype(np.float64)
        out1, pb1 = pg.eval_with_pullback(lambda m: pg.cross_entropy(m(xg), lg), model, math=args.math)
        g_full = pb1().vec().astype(np.float64)
        err = float(np.linalg.norm(g_dp - g_full) / np.linalg.norm(g_full))
        t = torch.tensor([err], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        parity["dp_grad_vs_full_batch_rel"] = float(t.item())
        assert parity["dp_grad_vs_full_batch_rel"] <= 1e-5, parity
        # (2) ... and after one update from that gradient the replicas' arenas must be bit-identical
        out, pb = pg.eval_with_pullback(lambda m: pg.cross_entropy(m(xs), ls, global_batch=PB * world), model, math=args.math)
        g = pb(on_ready=buckets.on_ready())
        buckets.finish_and_step(state, g, shadow=use_shadow)
        csum, cxor = dp.arena_checksum(model)            # (float64 sum, xor of the raw bits) of the parameter arena
        cs = torch.tensor([int(np.float64(csum).view(np.int64)), cxor], dtype=torch.int64, device=f"cuda:{local}")
        allcs = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(allcs, cs)
        sums = [tuple(c.tolist()) for c in allcs]
        parity["replica_checksums_identical"] = bool(all(s == sums[0] for s in sums))
        assert parity["replica_checksums_identical"], sums
    # first timed-size step: its loss against the CPU oracle's forward pass on the same shard (float32 oracle;
    # bf16 operand rounding bounds the gap at 2e-3, tests/test_gpu_tc.py)
    w_before = model.vec() if (rank == 0 and not args.no_cpu) else None
    loss0 = float(train_step(x_d, lab_d, gb=B) if world == 1 else pg.eval_with_pullback(lambda m: pg.cross_entropy(m(x_d), lab_d), model, math=args.math)[0])
    if w_before is not None:
        from oracle import protograd_oracle as O
        shapes = [p.shape for p in p0]
        probs, _ = O.net_forward(SPEC, O.unvec(w_before, shapes), x_h, keep=True)
        ref = float(O.cross_entropy(probs, lab_h))
        parity["first_step_loss"] = loss0
        parity["first_step_loss_oracle_f32"] = ref
        parity["first_step_loss_rel"] = abs(loss0 - ref) / abs(ref)
        assert parity["first_step_loss_rel"] <= 2e-3, parity
    barrier()                  # rank 0 spent seconds in the CPU oracle: re-align the ranks before any exchange

    # ---- device-resident leg (value) ----
    for _ in range(args.warmup):
        loss = train_step(x_d, lab_d)
    first_loss = float(loss)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = ctx.launch_count()
    ms = timed(lambda: train_step(x_d, lab_d), args.steps)
    launches = ctx.launch_count() - n0
    host_ms = host_enqueue[0]
    clk = clocks.stop() if rank == 0 else None
    value = GB * args.steps / (ms * 1e-3)

    # ---- end-to-end leg: every step's inputs come from pinned HOST memory (H2D inside the timed region,
    #      double-buffered on the copy stream so the copy of step i+1 overlaps step i) and every step's
    #      loss is read back to the host (D2H + sync) ----
    pf = pg.Prefetcher(ctx, e2e_specs, depth=2)

    def e2e_loop(n, lag=2):
        # the host reads EVERY step's loss (D2H copy queued right behind the step), `lag` steps after enqueueing it: with
        # lag = 1 the host can never be more than one step ahead of the GPU, and at 8 ranks — coupled by the exchange of every
        # step — any rank's enqueue jitter stalled all of them
        from collections import deque
        pf.submit(x_p, lab_p)
        pending, last = deque(), None
        for i in range(n):
            xd, ld = pf.next()
            if i + 1 < n:
                pf.submit(x_p, lab_p)
            if args.math == "bf16":        # 8-bit pixels -> bf16(pixel / 255) on the device; the upload slot is free again after it
                call("pg_vec_cast_u8_bf16", ctx.h, xb_dev.ptr, xd.ptr, xd.size, 255.0)
                pf.release()
                out = train_step(xb_dev, ld)
            else:
                out = train_step(xd, ld)
                pf.release()
            out.fetch_async()              # D2H of this step's loss, queued right behind the step
            pending.append(out)
            if len(pending) > lag:
                last = float(pending.popleft())
        while pending:
            last = float(pending.popleft())
        return last
    e2e_loop(3)
    e2e_steps = max(5, args.steps // 2)
    ms_e2e = timed(lambda: e2e_loop(e2e_steps), 1)
    e2e_value = GB * e2e_steps / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel (tcgen05 GEMM): the five wide GEMM launches of one step, event-timed
    #      one by one on REAL operands — U[0,1) inputs, the model's weights, ReLU activations (about half zeros,
    #      as in the step) and dense random hidden gradients ----
    roofline, kernels = None, {}
    if args.math == "bf16" and rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf = pg.BF16
        rng = np.random.default_rng(5)

        def dev_bf16(a):
            d = pg.DeviceArray.empty(ctx, a.shape, bf, zero=False)
            d.upload(f32_to_bf16(a))
            return d
        A0 = dev_bf16(x_h)
        W1, W2 = dev_bf16(model.layers[0].W.numpy()), dev_bf16(model.layers[2].W.numpy())
        A1 = dev_bf16(np.maximum(rng.standard_normal((B, D_H), dtype=np.float32), 0))
        A2 = pg.DeviceArray.empty(ctx, (B, D_H), bf)
        G1 = dev_bf16(1e-3 * rng.standard_normal((B, D_H), dtype=np.float32)); G2 = dev_bf16(1e-3 * rng.standard_normal((B, D_H), dtype=np.float32))
        dW1 = pg.DeviceArray.empty(ctx, (D_IN, D_H), np.float32); dW2 = pg.DeviceArray.empty(ctx, (D_H, D_H), np.float32)
        db = pg.DeviceArray.empty(ctx, (D_H,), np.float32)
        bias = pg.DeviceArray.from_numpy(ctx, (0.01 * rng.standard_normal(D_H)).astype(np.float32))
        gemms = [
            ("fwd_L1", lambda: call("pg_tc_linear_fwd", ctx.h, A0.ptr, W1.ptr, bias.ptr, A2.ptr, 2, B, D_IN, D_H, 1), 2.0 * B * D_IN * D_H),
            ("fwd_L2", lambda: call("pg_tc_linear_fwd", ctx.h, A1.ptr, W2.ptr, bias.ptr, A2.ptr, 2, B, D_H, D_H, 1), 2.0 * B * D_H * D_H),
            ("dgrad_L2", lambda: call("pg_tc_linear_bwd", ctx.h, None, W2.ptr, G2.ptr, 2, None, None, A2.ptr, A1.ptr, B, D_H, D_H), 2.0 * B * D_H * D_H),
            ("wgrad_L2", lambda: call("pg_tc_linear_bwd", ctx.h, A1.ptr, None, G2.ptr, 2, dW2.ptr, db.ptr, None, None, B, D_H, D_H), 2.0 * B * D_H * D_H),
            ("wgrad_L1", lambda: call("pg_tc_linear_bwd", ctx.h, A0.ptr, None, G1.ptr, 2, dW1.ptr, db.ptr, None, None, B, D_IN, D_H), 2.0 * B * D_IN * D_H),
        ]
        reps = 10
        tot_ms, tot_flop = 0.0, 0.0
        for name, fn, flop in gemms:
            fn(); fn()
        torch.cuda.synchronize()
        evs = []
        for _ in range(reps):
            for name, fn, flop in gemms:
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); fn(); b_.record(stream)
                evs.append((name, flop, a, b_))
        torch.cuda.synchronize()
        for name, flop, a, b_ in evs:
            t = a.elapsed_time(b_)
            k = kernels.setdefault(name, {"ms": 0.0, "flop": flop})
            k["ms"] += t / reps
            tot_ms += t
            tot_flop += flop
        for k in kernels.values():
            k["tflops"] = k["flop"] / (k["ms"] * 1e-3) / 1e12
        achieved = tot_flop / (tot_ms * 1e-3) / 1e12
        traffic, traffic_src = None, None
        try:  # DRAM bytes per launch from the committed ncu --set full capture of the same kernels
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))["tc_gemm_kernel"]
            if B == PER_GPU_BATCH:
                traffic, traffic_src = tj["per_launch_bytes"], tj["source"]
        except Exception:
            pass
        burst = peaks.get("bf16_tflops") or 1590.0
        sustained = peaks.get("bf16_tflops_sustained") or 1400.0
        roofline = {"kernel": "tc_gemm2_kernel (tcgen05 bf16 cta_group::2, the 5 wide GEMM launches of a step, real operands)", "bound": "tensor",
                    "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
                    "peak_source": ("MEASURED_PEAKS.json bf16_tflops (burst: each launch is timed alone; of measured)" if peaks
                                    else "fallback 1.59 PFLOP/s burst (of fallback)"),
                    "frac_of_sustained_peak": achieved / sustained, "sustained_peak": sustained,
                    "traffic": traffic, "traffic_source": traffic_src, "avg_launch_ms": tot_ms / (reps * len(gemms)),
                    "share_of_step": (tot_ms / reps) / (ms / args.steps),
                    "whole_step_tflops": FLOP_PER_SAMPLE * B / (ms / args.steps * 1e-3) / 1e12,
                    "whole_step_frac_of_burst": FLOP_PER_SAMPLE * B / (ms / args.steps * 1e-3) / 1e12 / burst}
        del A0, A1, A2, G1, G2, W1, W2, dW1, dW2

    # ---- the same step as a CUDA-graph replay (pg.GraphStep: one cudaGraphLaunch per step, the captured Adam pass writes the
    #      bf16 shadow).  Reported next to the headline, not as the headline: the data-parallel exchange hooks into the eager
    #      pullback (pb(on_ready=...)), so `value` stays on the path that is the same at every N ----
    graph_replay = None
    if world == 1 and args.math == "bf16" and not args.no_others:
        try:
            gs = pg.GraphStep(lambda m: pg.cross_entropy(m(x_d), lab_d), model, state, math="bf16")
            for _ in range(args.warmup):
                out = gs()
            float(out)
            n_g = max(10, args.steps // 2)
            ms_g = timed(gs, n_g)
            graph_replay = {"ms_per_step": ms_g / n_g, "value": GB * n_g / (ms_g * 1e-3), "unit": "samples/s", "steps": n_g,
                            "note": "pg.GraphStep replay of the identical step on device-resident inputs"}
        except Exception as e:          # informational: never take the headline down
            graph_replay = {"error": repr(e)[:200]}

    others = None
    if world == 1 and not args.no_others:
        others = other_configs(pg, ctx, stream, torch)

    if rank == 0:
        line = {
            "metric": "train samples/sec per step", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.math == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(B, world, args.math, args.dp_mode, args.exact_adam),
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(x_p.nbytes + lab_p.nbytes), "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / e2e_steps,
                    "inputs": "uint8 pixels (converted to bf16(pixel / 255) on the device) + int32 class indices from pinned host memory, double-buffered on a copy stream" if args.math == "bf16" else "fp32 pixels + fp32 one-hot labels from pinned host memory"},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_ms, "cpu_affinity_rank0": numa, "clocks": clk, "loss_after_warmup": first_loss,
            "tflops_per_gpu": FLOP_PER_SAMPLE * B / (ms / args.steps * 1e-3) / 1e12, "parity": parity,
        }
        if roofline:
            line["roofline"] = roofline
            line["kernels"] = kernels
        if graph_replay:
            line["graph_replay"] = graph_replay
        if others:
            line["other_configs"] = others
        if world == 1 and not args.no_cpu:
            v, per, n = cpu_port_throughput(3, 1, B)
            line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{B}-sample batches of the same step (the GPU arm's batch), {n} timed steps after 1 warm-up, numpy/OpenBLAS float32 oracle"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# the other BASELINE configs (N = 1): C1 MLP 784-100-10, C3 conv net, C5 frozen trunk + head
# ------------------------------------------------------------------------------------------------
def other_configs(pg, ctx, stream, torch):
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = (peaks.get("hbm_gbs") or 6650.0) * 1e9

    def build(spec, rng):
        layers = []
        for op in spec:
            k = op[0]
            if k == "linear":
                layers.append(pg.Linear(op[1], op[2], rng=rng))                 # xavier_uniform, src/layers.jl:6-8
            elif k == "conv":
                layers.append(pg.Conv(op[1], op[2], (op[3], op[3]), rng=rng))
            elif k == "relu":
                layers.append(pg.relu)
            elif k == "maxpool":
                layers.append(lambda t, s=op[1]: pg.maxpool(t, (s, s)))
            elif k == "flatten":
                layers.append(pg.flatten)
            elif k == "softmax":
                layers.append(pg.softmax)
        return pg.Compose(*layers)

    def run(name, spec, B, opt, opt_bytes, math, graph, steps=30, warmup=5):
        rng = np.random.default_rng(3)
        conv = spec[0][0] == "conv"
        in_shape = (1, 28, 28) if conv else (784,)
        x = rng.random((B,) + in_shape, dtype=np.float32)
        lab = np.eye(10, dtype=np.float32)[rng.integers(0, 10, B)]
        full = build(spec, rng)
        xd, ld = pg.DeviceArray.from_numpy(ctx, x), pg.DeviceArray.from_numpy(ctx, lab)
        if name == "C5":
            head_idx = [i for i, l in enumerate(full.layers) if isinstance(l, pg.Linear)][-1]
            m = full.layers[head_idx]

            def objective(mm):
                layers = list(full.layers); layers[head_idx] = mm
                return pg.cross_entropy(pg.Compose(*layers)(xd), ld)
        else:
            m = full
            objective = lambda mm: pg.cross_entropy(mm(xd), ld)
        st = pg.init(opt, m)
        if graph == "mega":
            stepf = pg.MegaStep(objective, m, st)           # forward + loss + backward + update in ONE cooperative launch
        elif graph:
            stepf = pg.GraphStep(objective, m, st, math=math)
        else:
            def stepf():
                out, pb = pg.eval_with_pullback(objective, m, math=math)
                pg.step(st, pb())
                return out
        for _ in range(warmup):
            out = stepf()
        float(out)
        n0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(steps):
            out = stepf()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        nbytes = step_bytes(spec, B, opt_bytes, in_shape)
        roof_ms = nbytes / hbm * 1e3
        return {"config": name, "batch": B, "optimizer": type(opt).__name__, "math": math,
                "mode": "one cooperative launch (MegaStep, csrc/mlp_step.cu)" if graph == "mega" else "cuda-graph replay" if graph else "eager",
                "ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "gflops": FLOPS_OTHER[name] * B / (ms * 1e-3) / 1e9,
                "launches_per_step": (ctx.launch_count() - n0) / steps if not graph else None,
                "roofline": {"bound": "hbm" if B >= 4096 else "latency (working set fits L2; HBM figure for scale only)",
                             "algorithmic_bytes": nbytes, "roofline_ms": roof_ms, "frac": roof_ms / ms},
                "loss": float(out)}

    res = []
    for args_ in (("C1", C1_SPEC, 128, pg.Adam(1e-3, fast=True), 28, "fp32", "mega"),
                  ("C1", C1_SPEC, 128, pg.Adam(1e-3, fast=True), 28, "fp32", True),
                  # 8,192-sample steps of these small nets are 15-35 launches of a few microseconds each: eager steps are bound by
                  # the host's enqueue rate, so the lines are the CUDA-graph replay of the same step (GraphStep) — plus one eager
                  # line for comparison
                  ("C1", C1_SPEC, 8192, pg.Adam(1e-3, fast=True), 28, "fp32", True),
                  ("C1", C1_SPEC, 8192, pg.Adam(1e-3, fast=True), 28, "bf16", True),
                  ("C3", C3_SPEC, 128, pg.NesterovMethod(1e-2), 16, "fp32", True),
                  ("C3", C3_SPEC, 8192, pg.NesterovMethod(1e-2), 16, "bf16", True),
                  ("C3", C3_SPEC, 8192, pg.NesterovMethod(1e-2), 16, "bf16", False),
                  ("C5", C5_SPEC, 8192, pg.GradientDescent(0.1), 12, "bf16", True)):
        try:
            res.append(run(*args_))
        except Exception as e:      # a sub-line must never take the headline down
            res.append({"config": args_[0], "batch": args_[2], "error": repr(e)[:300]})
    return res


if __name__ == "__main__":
    main()
