#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: train samples/sec per step (device-timed) on B200.

Workload (config C4 of SURVEY.md §8 / BASELINE.json configs[3]): wide MLP 784-4096-4096-10,
softmax + cross-entropy, Adam, 8,192 samples per GPU (global batch 65,536 at 8 GPUs -> weak scaling),
synthetic U[0,1) inputs and random one-hot labels, xavier-uniform weights.  A step is
forward -> loss -> backward -> (N>1: NCCL all-reduce of the flat gradient arena) -> fused Adam,
run through the package's public API (eval_with_pullback / pb() / step) in bf16 tensor-core mode
(bf16 operands, fp32 accumulation, fp32 master weights and optimizer state).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

--impl reference times the CPU oracle (numpy/OpenBLAS restatement of the reference step; Julia is
not installed) on a bounded sample of the same workload on the host cores.
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
CPU_SAMPLE_BATCH = 1024
SPEC = [("linear", D_IN, D_H), ("relu",), ("linear", D_H, D_H), ("relu",), ("linear", D_H, D_OUT), ("softmax",)]
# algorithmic flop/sample: fwd + dW for all layers + dX for all but the first
FLOP_PER_SAMPLE = 2 * (2 * (D_IN * D_H + D_H * D_H + D_H * D_OUT) + (D_H * D_H + D_H * D_OUT))


def synth(B, seed):
    rng = np.random.default_rng(seed)
    x = rng.random((B, D_IN), dtype=np.float32)
    lab = np.zeros((B, D_OUT), dtype=np.float32)
    lab[np.arange(B), rng.integers(0, D_OUT, B)] = 1.0
    return x, lab


def init_params(seed=0):
    rng = np.random.default_rng(seed)
    ps = []
    for (a, b) in ((D_IN, D_H), (D_H, D_H), (D_H, D_OUT)):
        s = np.float32(np.sqrt(6.0 / (a + b)))
        ps.append((s * (2 * rng.random((a, b), dtype=np.float32) - 1)).astype(np.float32))
        ps.append(np.zeros(b, dtype=np.float32))
    return ps


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference step) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_port_throughput(steps, warmup, batch=CPU_SAMPLE_BATCH):
    from oracle import protograd_oracle as O
    x, lab = synth(batch, 123)
    ts = O.TrainStep(SPEC, init_params(0), O.Adam(stepsize=1e-3), "ce")
    for _ in range(warmup):
        ts.step(x, lab)
    t0 = time.perf_counter()
    for _ in range(steps):
        ts.step(x, lab)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 10))
    warm = max(1, min(args.warmup, 2))
    v, per = cpu_port_throughput(steps, warm)
    cores = os.cpu_count()
    sample = f"{CPU_SAMPLE_BATCH}-sample batches of the same 784-4096-4096-10 Adam step, {steps} timed steps after {warm} warm-up, numpy/OpenBLAS float32"
    line = {
        "impl": "reference", "metric": "train samples/sec per step", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "wide MLP 784-4096-4096-10, softmax+cross-entropy, Adam; CPU sample batch %d (GPU arm: 8192/GPU)" % CPU_SAMPLE_BATCH,
                   "note": "CPU restatement of the reference step (not Julia: no Julia toolchain in this image)"},
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import protograd_b200 as pg
    from protograd_b200 import dp
    from protograd_b200.capi import call

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
        dp.broadcast_parameters(model)
    x_h, lab_h = synth(B, 1000 + rank)
    x_d, lab_d = pg.DeviceArray.from_numpy(ctx, x_h), pg.DeviceArray.from_numpy(ctx, lab_h)
    # pinned host copies for the end-to-end leg
    pin_x, pin_l = pg.PinnedBuffer(x_h.nbytes), pg.PinnedBuffer(lab_h.nbytes)
    x_p, lab_p = pin_x.as_numpy(np.float32, x_h.shape), pin_l.as_numpy(np.float32, lab_h.shape)
    x_p[...] = x_h; lab_p[...] = lab_h

    buckets = dp.GradientBuckets(pg.gradient_container(model), bucket_bytes=int(os.environ.get("PG_BUCKET_MB", "24")) << 20,
                                 comm_sms=(int(os.environ["PG_COMM_SMS"]) if "PG_COMM_SMS" in os.environ else None),
                                 delay_us=(float(os.environ["PG_COMM_DELAY_US"]) if "PG_COMM_DELAY_US" in os.environ else None),
                                 overlap_update=(os.environ["PG_DP_OVERLAP_UPDATE"] == "1" if "PG_DP_OVERLAP_UPDATE" in os.environ else None)) if world > 1 else None
    if args.sm_limit:
        ctx.set_sm_limit(args.sm_limit)
    use_shadow = args.math == "bf16"
    timeline = [] if os.environ.get("PG_DP_TIMELINE") else None
    timeline_leaves = []

    klog = {"cur": None, "steps": []}
    if os.environ.get("PG_DP_TIMELINE") == "2":
        # per-call device timeline of the backward + update kernels (diagnostic): an event after every C-ABI call
        from protograd_b200 import trace as _trace, optimizers as _optimizers
        _orig_call = _trace.call

        def _logged_call(name, *a):
            r = _orig_call(name, *a)
            if klog["cur"] is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(stream)
                klog["cur"].append((name, e))
            return r
        _trace.call = _logged_call
        _optimizers.call = _logged_call

    def train_step(x, lab):
        out, pb = pg.eval_with_pullback(lambda m: pg.cross_entropy(m(x), lab, global_batch=GB), model, math=args.math)
        if buckets is None:
            pg.step(state, pb(), shadow=use_shadow)
        elif args.dp_mode == "bucket":
            if timeline is not None:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                leaf_ev = {}
                ev[0].record(stream)

                def hook(i):
                    buckets.ready(i)
                    leaf_ev[i] = torch.cuda.Event(enable_timing=True)
                    leaf_ev[i].record(stream)  # gradient leaf i has been produced
                    if i == 2:
                        ev[1].record(stream)   # parameter gradients of layers 2,3 done; their bucket was just issued
                if os.environ.get("PG_DP_TIMELINE") == "2":
                    klog["cur"] = []
                g = pb(on_ready=hook)
                ev[2].record(stream)           # end of backward compute
                buckets.finish_and_step(state, g, shadow=use_shadow)
                ev[3].record(stream)
                if klog["cur"] is not None:
                    klog["steps"].append((ev[0], klog["cur"]))
                    klog["cur"] = None
                timeline.append(ev)
                timeline_leaves.append((ev[0], leaf_ev))
            else:
                g = pb(on_ready=buckets.ready)      # tail buckets reduce while the remaining backward GEMMs run
                buckets.finish_and_step(state, g, shadow=use_shadow)
        elif args.dp_mode == "flat":
            g = pb()
            dp.allreduce_gradients(g)
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
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

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
    if timeline and rank == 0:
        torch.cuda.synchronize()
        tl = np.array([[e[0].elapsed_time(e[k]) for k in (1, 2, 3)] for e in timeline[-20:]])
        print("DP timeline (ms from step start: bucket issued, backward done, step done):", tl.mean(0).round(4).tolist(), file=sys.stderr)
        if klog["steps"]:
            last = klog["steps"][-20:]
            names = [n for n, _ in last[-1][1]]
            t = np.array([[e0.elapsed_time(e) for _, e in log] for e0, log in last if len(log) == len(names)]).mean(0)
            print("DP kernel timeline (ms from backward start, call finished):", file=sys.stderr)
            prev = 0.0
            for n, x in zip(names, t):
                print(f"   {n:28s} done {x:7.4f}  (+{x - prev:6.4f})", file=sys.stderr)
                prev = x
        keys = sorted(timeline_leaves[-1][1], reverse=True)
        lv = np.array([[e0.elapsed_time(le[k]) for k in keys] for e0, le in timeline_leaves[-20:]])
        print("DP timeline leaves (ms from backward start, gradient leaf done):", dict(zip(keys, lv.mean(0).round(4).tolist())), file=sys.stderr)
    clk = clocks.stop() if rank == 0 else None
    value = GB * args.steps / (ms * 1e-3)

    # ---- end-to-end leg: every step's inputs come from pinned HOST memory (H2D inside the timed region,
    #      double-buffered on the copy stream so the copy of step i+1 overlaps step i) and every step's
    #      loss is read back to the host (D2H + sync) ----
    pf = pg.Prefetcher(ctx, [(x_h.shape, np.float32), (lab_h.shape, np.float32)], depth=2)

    def e2e_loop(n):
        pf.submit(x_p, lab_p)
        last, prev = None, None
        for i in range(n):
            xd, ld = pf.next()
            if i + 1 < n:
                pf.submit(x_p, lab_p)
            out = train_step(xd, ld)
            pf.release()
            out.fetch_async()              # D2H of this step's loss, queued right behind the step
            if prev is not None:
                last = float(prev)         # read step i-1's loss while step i runs
            prev = out
        return float(prev)
    e2e_loop(3)
    e2e_steps = max(5, args.steps // 2)
    ms_e2e = timed(lambda: e2e_loop(e2e_steps), 1)
    e2e_value = GB * e2e_steps / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel (tcgen05 GEMM): every GEMM launch of one step, event-timed ----
    roofline, kernels = None, {}
    if args.math == "bf16":
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf = pg.BF16
        A0 = pg.DeviceArray.empty(ctx, (B, D_IN), bf); A1 = pg.DeviceArray.empty(ctx, (B, D_H), bf); A2 = pg.DeviceArray.empty(ctx, (B, D_H), bf)
        G1 = pg.DeviceArray.empty(ctx, (B, D_H), bf); G2 = pg.DeviceArray.empty(ctx, (B, D_H), bf)
        W1 = pg.DeviceArray.empty(ctx, (D_IN, D_H), bf); W2 = pg.DeviceArray.empty(ctx, (D_H, D_H), bf)
        dW1 = pg.DeviceArray.empty(ctx, (D_IN, D_H), np.float32); dW2 = pg.DeviceArray.empty(ctx, (D_H, D_H), np.float32)
        bias = pg.DeviceArray.empty(ctx, (D_H,), np.float32)
        gemms = [
            ("fwd_L1", lambda: call("pg_tc_linear_fwd", ctx.h, A0.ptr, W1.ptr, bias.ptr, A1.ptr, 2, B, D_IN, D_H, 1), 2.0 * B * D_IN * D_H),
            ("fwd_L2", lambda: call("pg_tc_linear_fwd", ctx.h, A1.ptr, W2.ptr, bias.ptr, A2.ptr, 2, B, D_H, D_H, 1), 2.0 * B * D_H * D_H),
            ("dgrad_L2", lambda: call("pg_tc_linear_bwd", ctx.h, None, W2.ptr, G2.ptr, 2, None, None, G1.ptr, A1.ptr, B, D_H, D_H), 2.0 * B * D_H * D_H),
            ("wgrad_L2", lambda: call("pg_tc_linear_bwd", ctx.h, A1.ptr, None, G2.ptr, 2, dW2.ptr, None, None, None, B, D_H, D_H), 2.0 * B * D_H * D_H),
            ("wgrad_L1", lambda: call("pg_tc_linear_bwd", ctx.h, A0.ptr, None, G1.ptr, 2, dW1.ptr, None, None, None, B, D_IN, D_H), 2.0 * B * D_IN * D_H),
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
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["tc_gemm_kernel"]
            if B == PER_GPU_BATCH:
                traffic, traffic_src = tj["per_launch_bytes"], tj["source"]
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained") or 1400.0
        roofline = {"kernel": "tc_gemm_kernel (tcgen05 bf16, 5 launches/step)", "bound": "tensor", "achieved": achieved, "peak": peak,
                    "unit": "TFLOP/s", "frac": achieved / peak,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)",
                    "traffic": traffic, "traffic_source": traffic_src, "avg_launch_ms": tot_ms / (reps * len(gemms)),
                    "share_of_step": (tot_ms / reps) / (ms / args.steps)}

    if rank == 0:
        line = {
            "metric": "train samples/sec per step", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.math == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": "wide MLP 784-4096-4096-10 (BASELINE configs[3]), softmax+cross-entropy, Adam(1e-3), %d samples/GPU" % B,
                       "global_batch": GB, "parallelism": "dp%d" % world, "dp_mode": args.dp_mode if world > 1 else None, "math": args.math, "adam": "float64-promoted (bit-exact)" if args.exact_adam else "fused float32 (normwise 1e-7 of the reference form)",
                       "l2": "working set per step (>600 MB of activations, weights, gradients, optimizer state) exceeds the 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(x_h.nbytes + lab_h.nbytes), "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / e2e_steps},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_ms, "clocks": clk, "loss_after_warmup": first_loss,
            "tflops_per_gpu": FLOP_PER_SAMPLE * B / (ms / args.steps * 1e-3) / 1e12,
        }
        if roofline:
            line["roofline"] = roofline
            line["kernels"] = kernels
        if world == 1 and not args.no_cpu:
            v, per = cpu_port_throughput(3, 1)
            line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{CPU_SAMPLE_BATCH}-sample batches of the same step, 3 timed steps after 1 warm-up, numpy/OpenBLAS float32 oracle"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
