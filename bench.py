#!/usr/bin/env python
"""Benchmark of the intersected-line robust registration loss on B200 (contract: see the task's bench.py section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload dcp|rpm|fmr|demo|large] [--impl reference]

One "step" = one forward + backward evaluation of the loss over one batch of synthetic pairs per GPU
(forward: dense intersection + sparse Welsch phase; backward: d loss / d points1), plus -- for N > 1 -- the
all-reduce of the scalar loss (batch-sharded workloads) or the line-shard exchange (workload `large`).
Metric: pairs x lines per second, whole job.  Default workload = BASELINE.json configs[1] ("DCP-style ModelNet40
training loss: batch 32 pairs x 1024 points", 15000 lines per pair as in Train_DCP.py:252-255), weak scaling:
every GPU evaluates its own batch of 32 pairs.

`--impl reference` times the reference's CPU path (the eager-PyTorch restatement in oracle/torch_port.py: the
reference is pure Python and its sources do not travel to the GPU box) on the host cores, one pair per step.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (pairs per GPU, triplets per cloud, lines per pair, synth kwargs, description)
    "dcp": (32, 1024, 15000, dict(radius_scale=0.5),
            "BASELINE configs[1]: DCP-style batch, 32 pairs x 1024 triplets x 15000 lines per GPU"),
    "rpm": (64, 2048, 10000, dict(radius_scale=1.0, noise=0.01, outlier_frac=0.1, keep_frac=0.7),
            "BASELINE configs[2]: RPM-Net partial overlap, 64 pairs x 2048 triplets x 10000 lines per GPU"),
    "fmr": (128, 1024, 15000, dict(radius_scale=0.5),
            "BASELINE configs[3]: FMR batch, 128 pairs x 1024 triplets x 15000 lines per GPU"),
    "demo": (1, 1024, 20000, dict(radius_scale=1.0),
             "BASELINE configs[0] shape: one pair, 1024 triplets x 20000 lines"),
    "large": (1, 500000, 100000, dict(radius_scale=0.5),
              "BASELINE configs[4]: one scan pair, 500k triplets x 100k lines, line-sharded across GPUs"),
}
L2_BYTES = 126 << 20


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ---------------------------------------------------------------------------------------------------------
# synthetic inputs
# ---------------------------------------------------------------------------------------------------------
def rigid_variants(base, n_out, seed, keep_first=True):
    """n_out rigidly moved copies of the base pairs (distinct bits, identical statistics)."""
    from oracle import synth
    rng = np.random.default_rng(seed)
    tri1, tri2, lines = [], [], []
    for i in range(n_out):
        p = base[i % len(base)]
        if keep_first and i < len(base):
            R, t = np.eye(3), np.zeros(3)
        else:
            R, t = synth.random_rotation(rng, 180.0), rng.uniform(-0.5, 0.5, 3)
        mv = lambda x: (x.reshape(-1, 3).astype(np.float64) @ R.T + t).astype(np.float32)
        tri1.append(mv(p["tri1"]).reshape(-1, 9))
        tri2.append(mv(p["tri2"]).reshape(-1, 9))
        ln = p["lines"].astype(np.float64)
        lines.append(np.concatenate([ln[:, :3] @ R.T, ln[:, 3:] @ R.T + t], 1).astype(np.float32))
    return np.stack(tri1), np.stack(tri2), np.stack(lines)


def make_inputs(workload, rank, n_sets, n_base=None, nl=None):
    from oracle import synth
    B, nf, nl0, kw, _ = WORKLOADS[workload]
    nl = nl or nl0
    n_base = n_base or min(B, 8 if nf <= 4096 else 1)
    cfg_id = list(WORKLOADS).index(workload) + 2
    base = [synth.make_pair(1000 * cfg_id + 97 * rank + i, nf, nl, **kw) for i in range(n_base)]
    sets = []
    for s in range(n_sets):
        sets.append(rigid_variants(base, B, seed=7919 * (rank + 1) + s, keep_first=(s == 0)))
    return sets


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and clock-event reasons sampled DURING the timed region.

    A step lasts well under a millisecond, so `nvidia-smi -lms 100` would return nothing for a short run: NVML is polled
    in-process from a thread every 2 ms instead (the ctypes calls drop the GIL); nvidia-smi is the fallback."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index, period_s=0.002):
        import threading
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self._stop = threading.Event()
        self._thread, self._smi, self._path = None, None, None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(gpu_index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = [(pynvml.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                    (pynvml.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                    (pynvml.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                    (pynvml.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]

            def poll():
                while not self._stop.is_set():
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        for bit, nme in bits:
                            if r & bit:
                                self.reasons.add(nme)
                        self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                    except Exception:
                        pass
                    self._stop.wait(period_s)
            self._thread = threading.Thread(target=poll, daemon=True)
            self._thread.start()
        except Exception:
            self._thread = None
            try:
                self._path = tempfile.mktemp(prefix="rrl_clocks_", suffix=".csv")
                fields = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                          "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                          "clocks_event_reasons.sw_power_cap")
                self._smi = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + fields,
                                              "--format=csv,noheader,nounits", "-lms", "20"],
                                             stdout=open(self._path, "w"), stderr=subprocess.DEVNULL)
            except Exception:
                self._smi = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
            out["source"] = "nvml polled every 2 ms inside the timed region"
        elif self._smi is not None:
            self._smi.terminate()
            try:
                self._smi.wait(timeout=5)
            except Exception:
                self._smi.kill()
            try:
                for row in open(self._path):
                    f = [x.strip() for x in row.split(",")]
                    if len(f) < 7:
                        continue
                    try:
                        self.samples.append(float(f[0]))
                        self.max_mhz = float(f[1])
                    except ValueError:
                        continue
                    for nme, v in zip(self.NAMES, f[3:7]):
                        if v.lower().startswith("active"):
                            self.reasons.add(nme)
                os.unlink(self._path)
            except Exception:
                pass
            out["source"] = "nvidia-smi -lms 20"
        if self.samples:
            out["sm_mhz"] = float(np.median(self.samples))
            out["sm_min_mhz"] = float(np.min(self.samples))
            out["samples"] = len(self.samples)
        if self.power:
            out["power_w"] = float(np.median(self.power))
        out["sm_max_mhz"] = self.max_mhz
        out["reasons"] = sorted(self.reasons)
        return out


# ---------------------------------------------------------------------------------------------------------
# CPU baseline (the reference's PyTorch CPU path, restated in oracle/torch_port.py)
# ---------------------------------------------------------------------------------------------------------
def cpu_step(tri1, tri2, lines, chunk=512):
    import torch
    from oracle import torch_port as tp
    t1 = torch.from_numpy(tri1).clone().requires_grad_(True)
    loss = tp.loss_pair(t1, torch.from_numpy(tri2), torch.from_numpy(lines), chunk=chunk)
    if loss is not None:
        loss.backward()
    return 0.0 if loss is None else float(loss.item())


def cpu_baseline(workload, budget_s=20.0, inputs=None):
    """pairs x lines / s of the eager-PyTorch CPU path on a bounded sample of the workload (one pair, a prefix of its lines)."""
    import torch
    B, nf, nl, kw, _ = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tri1, tri2, lines = inputs if inputs is not None else make_inputs(workload, 0, 1, n_base=1)[0]
    tri1, tri2, lines = tri1[0], tri2[0], lines[0]
    # the dense phase costs 36*nl*nf bytes per temporary: bound the sample so that one step takes a few seconds
    n_s = nl if nf * nl <= 1024 * 20000 else max(64, int(1024 * 20000 / nf))
    t0 = time.perf_counter()
    cpu_step(tri1, tri2, lines[:n_s])
    warm = time.perf_counter() - t0
    reps = max(1, min(5, int(budget_s / max(warm, 1e-3)) - 1))
    best = warm
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_step(tri1, tri2, lines[:n_s])
        best = min(best, time.perf_counter() - t0)
    return {"value": n_s / best, "unit": "pairs*lines/s", "cores": cores, "kind": "port",
            "sample": "1 pair of the workload (%d triplets per cloud), first %d of %d lines, forward+backward, eager "
                      "PyTorch CPU restatement of code/loss.py (oracle/torch_port.py), best of %d; extrapolation in the "
                      "line count is linear" % (nf, n_s, nl, reps + 1),
            "seconds_per_step": best}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    import torch
    B, nf, nl, kw, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tri1, tri2, lines = make_inputs(args.workload, 0, 1, n_base=1)[0]
    tri1, tri2, lines = tri1[0], tri2[0], lines[0]
    n_s = nl if nf * nl <= 1024 * 20000 else max(64, int(1024 * 20000 / nf))
    t0 = time.perf_counter()
    cpu_step(tri1, tri2, lines[:n_s])
    first = time.perf_counter() - t0
    # keep the whole run (warmup + steps) within ~150 s by shrinking the per-step line sample if needed
    total_steps = args.steps + max(args.warmup - 1, 0)
    if first * total_steps > 150.0:
        n_s = max(64, int(n_s * 150.0 / (first * total_steps)))
    for _ in range(max(args.warmup - 1, 0)):
        cpu_step(tri1, tri2, lines[:n_s])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(tri1, tri2, lines[:n_s])
    dt = (time.perf_counter() - t0) / args.steps
    value = n_s / dt
    sample = ("each step = forward+backward of 1 pair of the workload (%d triplets per cloud) on the first %d of its %d "
              "lines; eager PyTorch CPU restatement of the reference (oracle/torch_port.py), %d threads" % (nf, n_s, nl, cores))
    line = {"impl": "reference", "metric": "loss fwd+bwd evaluations/s (pairs x lines per second)", "value": value,
            "unit": "pairs*lines/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": desc, "name": args.workload, "triplets_per_cloud": nf,
                                            "lines_per_pair": nl},
            "cpu_baseline": {"value": value, "unit": "pairs*lines/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "pairs*lines/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import rrl_b200
    L = rrl_b200._native.lib()

    B, nf, nl, kw, desc = WORKLOADS[args.workload]
    line_sharded = args.workload == "large" and world > 1
    weak_lines = line_sharded and args.large_scaling == "weak"     # every rank keeps nl lines: world * nl lines in total
    bytes_per_set = 4 * (2 * B * nf * 9 + B * nl * 6)
    n_sets = max(2, min(16, math.ceil(1.5 * L2_BYTES / bytes_per_set)))
    nl_total = nl * world if weak_lines else nl                    # lines of the ONE pair all ranks share
    host_sets = make_inputs(args.workload, 0 if line_sharded else rank, n_sets, nl=nl_total)
    if line_sharded:
        lo, hi = rrl_b200.dist.shard_range(nl_total, rank, world)
        host_sets = [(a, b_, c[:, lo:hi]) for a, b_, c in host_sets]
    dev_sets = [tuple(torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in s) for s in host_sets]
    nl_local = dev_sets[0][2].shape[1]

    twist_mode = args.workload == "large"          # the scan pair is evaluated through the se(3) twist
    twist0 = torch.tensor([0.01, -0.02, 0.015, 0.03, -0.01, 0.02], device=dev)

    # --reuse-order (twist-mode workloads = registration loops on a fixed pair): every input set keeps a LossSession, so
    # each step after the first reuses the spatial order of both clouds left by the previous step on that pair
    sessions = [rrl_b200.LossSession() for _ in range(n_sets)] if (twist_mode and args.reuse_order) else None

    def compute(i):
        """the device work of one step on input set i: forward + backward (for the line shard: with its exchange)"""
        t1, t2, ln = dev_sets[i % n_sets]
        if twist_mode:
            tw = twist0.clone().requires_grad_(True)
            sess = sessions[i % n_sets] if sessions else None
            if line_sharded:
                loss, _, _ = rrl_b200.dist.line_sharded_twist_loss(tw, t1[0], t2[0], ln[0], session=sess)
            else:
                tri1 = rrl_b200.se3_apply(tw.reshape(1, 6), t1.reshape(1, -1, 3)).reshape(1, -1, 9)
                loss = rrl_b200.intersected_line_loss(tri1, t2, ln, session=sess)
            total = loss.sum()
            total.backward()
            return total.detach(), tw.grad
        t1 = t1.detach().requires_grad_(True)
        loss = rrl_b200.intersected_line_loss(t1, t2, ln)
        total = loss.sum()
        total.backward()
        return total.detach(), t1.grad

    def exchange(out):
        if world > 1 and not line_sharded:
            # the only exchange of the batch-sharded path: the global loss, a logged scalar that nothing on
            # the device waits for.  It is reduced asynchronously (NCCL's own stream) and collected one step later, so
            # its latency (tens of microseconds at 8 ranks) runs under the next step's kernels; the last one is awaited
            # inside the timed region.
            red = out[0].clone()
            if pending:
                pending.pop().wait()
            pending.append(dist.all_reduce(red, async_op=True))
            return red, out[1]
        return out

    def step(i):
        return exchange(compute(i))

    pending = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # One CUDA graph per input set holds the device work of a step -- (transform,) forward, backward, and for the line shard
    # its five tiny collectives -- so that the host's launch rate (Python + autograd: ~0.25 ms per step, as long as the DCP
    # step itself) does not bound the measurement.  Same kernels, same collectives, same inputs; `--graph 0` issues them
    # eagerly.  The asynchronous all-reduce of the batch shard stays outside the graphs.
    graphs = None
    if args.graph:
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for i in range(max(3, n_sets)):
                    compute(i)
            torch.cuda.current_stream(dev).wait_stream(side)
            barrier()
            graphs = []
            for i in range(n_sets):
                gph = torch.cuda.CUDAGraph()
                n0 = rrl_b200.launch_count()
                with torch.cuda.graph(gph):
                    out = compute(i)
                graphs.append((gph, out, rrl_b200.launch_count() - n0))     # kernels of OURS inside this graph
            barrier()

            def step(i):                                   # noqa: F811
                gph, out, _ = graphs[i % n_sets]
                gph.replay()
                return exchange(out)
        except Exception as exc:                           # capture not available: keep the eager step, say so
            graphs = None
            sys.stderr.write("bench: CUDA graph capture failed (%s: %s); running eagerly\n" % (type(exc).__name__, exc))
            torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = rrl_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        last = step(i)
    while pending:
        pending.pop().wait()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = rrl_b200.launch_count() - launches0
    if graphs:                                            # replayed, not re-issued by the host: count what each graph holds
        launches = sum(graphs[i % n_sets][2] for i in range(args.steps))
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    pairs_lines_per_step = (1 * nl_total) if line_sharded else (world * B * nl)
    value = pairs_lines_per_step / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through the C ABI (H2D of the step's inputs from pinned memory, D2H of loss + status) ----
    e2e = None
    if not line_sharded:
        ctx = C.c_void_p()
        rrl_b200._native.check(L.rrl_host_create(B, nf, nf, nl, local_rank, C.byref(ctx)), "rrl_host_create")
        pinned = [torch.from_numpy(np.ascontiguousarray(np.concatenate([x.reshape(-1) for x in s]))).pin_memory()
                  for s in host_sets]
        n_sub = int(L.rrl_host_subbatches(ctx))
        n1, n2 = B * nf * 9, B * nf * 9
        h_loss = np.zeros(B, np.float32)
        h_status = np.zeros(B, np.int32)

        n_slots = int(L.rrl_host_slots(ctx))
        tickets = [C.c_int(-1) for _ in range(n_slots)]

        def submit(i):
            p = pinned[i % n_sets]
            base = p.data_ptr()
            rrl_b200._native.check(L.rrl_host_submit(ctx, base, base + 4 * n1, base + 4 * (n1 + n2), 1, 1, 5, 5, 0,
                                                     C.byref(tickets[i % n_slots])), "rrl_host_submit")

        def wait(i):
            rrl_b200._native.check(L.rrl_host_wait(ctx, tickets[i % n_slots].value, h_loss.ctypes.data, h_status.ctypes.data, None),
                                   "rrl_host_wait")

        def run(n):
            # the input pipeline of a training loop: the copies of step i+1 are queued before step i's result is awaited,
            # so they run under its kernels; every step still copies its own inputs in and its loss + status out
            submit(0)
            for i in range(n):
                if i + 1 < n:
                    submit(i + 1)
                wait(i)
        run(3)
        barrier()
        t0 = time.perf_counter()
        run(args.steps)
        dt = time.perf_counter() - t0
        # the same steps through the blocking call (no overlap between consecutive steps), reported beside it
        t0 = time.perf_counter()
        for i in range(args.steps):
            p = pinned[i % n_sets]
            rrl_b200._native.check(L.rrl_host_loss_fwd_bwd(ctx, p.data_ptr(), p.data_ptr() + 4 * n1, p.data_ptr() + 4 * (n1 + n2),
                                                           1, 1, 5, 5, h_loss.ctypes.data, h_status.ctypes.data, None),
                                   "rrl_host_loss_fwd_bwd")
        dt_sync = time.perf_counter() - t0
        L.rrl_host_destroy(ctx)
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * nl / (float(t.item()) / args.steps), "unit": "pairs*lines/s",
               "h2d_bytes_per_step": bytes_per_set, "d2h_bytes_per_step": 8 * B,
               "api": "rrl_host_submit + rrl_host_wait (C ABI, host pointers; per step: H2D of the three inputs from pinned "
                      "memory, forward, backward to points1, D2H of loss+status; %d buffer sets in flight, so the copies of "
                      "step i+1 run under the kernels of step i, and each step is cut into %d sub-batches of pairs on "
                      "their own streams)" % (n_slots, n_sub),
               "ms_per_step": float(t.item()) / args.steps * 1e3,
               "blocking_call_ms_per_step": dt_sync / args.steps * 1e3}

    if rank != 0:
        if world > 1:
            dist.barrier()
        return 0

    # ---- roofline of the dominant kernel (dense intersection), timed alone with CUDA events on its own stream ----
    peak, peak_ms = C.c_double(), C.c_double()
    L.rrl_measure_fp32_peak(0, C.byref(peak), C.byref(peak_ms))
    t1, t2, ln = dev_sets[0]
    wsb = L.rrl_workspace_bytes(t1.shape[0], nf, nf, nl_local)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    md, mp = C.c_float(), C.c_float()
    torch.cuda.synchronize()
    rrl_b200._native.check(L.rrl_measure_dense(t1.data_ptr(), t2.data_ptr(), ln.data_ptr(), t1.shape[0], nf, nf, nl_local,
                                               ws.data_ptr(), wsb, 10, C.byref(md), C.byref(mp), None), "rrl_measure_dense")
    alg_flops = 48.0 * t1.shape[0] * nl_local * (nf + nf)              # 16 flops per (line, point) test
    achieved = alg_flops / (md.value * 1e-3) / 1e12
    traffic = None
    prof = os.path.join(ROOT, "profiles", "r01_dense_%s.json" % args.workload)
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "fp32", "kernel": "rrl::dense_kernel", "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
                "frac": achieved / peak.value, "traffic": traffic,
                "peak_source": "measured live: register-resident FFMA loop (rrl_measure_fp32_peak); MEASURED_PEAKS.json has "
                               "no FP32 entry (hbm_gbs 6552.3 applies to the memory side only)",
                "algorithmic_flops_per_launch": alg_flops, "kernel_ms": md.value, "prep_ms": mp.value,
                "kernel_share_of_step": md.value / ms_per_step,
                "note": "achieved counts the ALGORITHMIC 48 flops per (line, triplet) of the reference formulation; the "
                        "kernel decides most (line, triplet) pairs with a conservative bounding-sphere + FMA predicate and "
                        "runs the exact reference-order test only on candidates, so frac can exceed 1 (DESIGN.md)",
                "algorithmic_hbm_bytes_per_launch": 4.0 * t1.shape[0] * (9 * 2 * nf + 6 * nl_local)}
    # the memory side of the same kernel, for the record (north star: "as a fraction of FP32 and HBM peak"): the
    # algorithmic bytes (both clouds and the lines once per launch) against the measured copy bandwidth
    hbm_peak = None
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    hbm_achieved = roofline["algorithmic_hbm_bytes_per_launch"] / (md.value * 1e-3) / 1e9
    roofline["hbm"] = {"achieved": hbm_achieved, "peak": hbm_peak or 6552.3, "unit": "GB/s",
                       "frac": hbm_achieved / (hbm_peak or 6552.3),
                       "peak_source": "MEASURED_PEAKS.json hbm_gbs" if hbm_peak else "fallback: the pool's measured 6552.3 GB/s",
                       "traffic": traffic}

    # ---- lines sampled on the device instead of supplied (SURVEY 8(d) asks for both; the sampler stays out of the
    # roofline figure): device time of rrl_sample_lines for one batch of the workload, and the step rate with it ----
    sampled = None
    try:
        v1 = t1.reshape(t1.shape[0], -1, 9)[:, :, :3].contiguous()           # the clouds = point 0 of every triplet
        v2 = t2.reshape(t2.shape[0], -1, 9)[:, :, :3].contiguous()
        lo2, hi2 = v2.min(1)[0], v2.max(1)[0]
        rad = (hi2 - lo2).norm(dim=1, keepdim=True) * float(kw.get("radius_scale", 0.5))
        cen = v2.mean(1)
        for _ in range(3):
            lines_s, filled = rrl_b200.sample_lines(rad, cen, nl_local, v1, v2, seed=11, offset=0)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s0.record()
        for it in range(20):
            lines_s, filled = rrl_b200.sample_lines(rad, cen, nl_local, v1, v2, seed=11, offset=it + 1)
        s1.record()
        torch.cuda.synchronize()
        smp_ms = s0.elapsed_time(s1) / 20
        sampled = {"sampler_ms_per_batch": smp_ms, "filled_fraction": float(filled.float().mean().item()) / nl_local,
                   "value_with_sampling": pairs_lines_per_step / ((ms_per_step + smp_ms) * 1e-3), "unit": "pairs*lines/s",
                   "note": "rrl_sample_lines (Philox4x32-10, 10 rounds, 12-triangle AABB test of loss.py:265-432) timed alone on "
                           "this rank's batch; value_with_sampling = the step with the sampler in front of it"}
    except Exception as exc:
        sampled = {"error": "%s: %s" % (type(exc).__name__, exc)}

    cpu = cpu_baseline(args.workload, inputs=host_sets[0]) if (world == 1 and not args.no_cpu_baseline) else None
    line = {"metric": "loss fwd+bwd evaluations/s (pairs x lines per second)", "value": value, "unit": "pairs*lines/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong" if (line_sharded and not weak_lines) else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": desc, "name": args.workload, "pairs_per_gpu": B, "triplets_per_cloud": nf,
                       "lines_per_pair": nl_total, "sharding": "lines" if line_sharded else "pairs (batch)",
                       "window": [1, 1, 5, 5],
                       "backward": "d loss / d twist (6) through the fused se(3) transform of cloud 1" if twist_mode
                                   else "d loss / d points1 (B, nf, 9)",
                       "launch": ("one CUDA graph per input set" if graphs else "eager launches"),
                       "reuse_order": bool(sessions),
                       "l2": "inputs rotate over %d distinct sets (%.0f MB > 126 MB L2), no flush needed" %
                             (n_sets, n_sets * bytes_per_set / 2 ** 20)},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "lines_sampled_on_device": sampled,
            "loss_checksum": float(last[0].item())}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dcp", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", type=int, default=1, help="1: replay the device work of a step as a CUDA graph (one per input set); 0: eager launches")
    ap.add_argument("--reuse-order", type=int, default=0,
                    help="twist-mode workloads: 1 = keep the clouds' spatial order from step to step (RRL_REUSE_ORDER)")
    ap.add_argument("--large-scaling", default="strong", choices=["strong", "weak"],
                    help="workload large at N > 1: strong = the 100k lines of the pair are split over the ranks (BASELINE "
                         "configs[4]); weak = every rank keeps 100k lines of a pair with N x 100k lines")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    _, _, world = dist_env()
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
