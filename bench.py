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
    from tools import synth
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
    from tools import synth
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
# shared description of a run (both arms print the SAME config dict)
# ---------------------------------------------------------------------------------------------------------
METRIC = "loss fwd+bwd evaluations/s (pairs x lines per second)"


def n_input_sets(workload):
    B, nf, nl, _, _ = WORKLOADS[workload]
    bytes_per_set = 4 * (2 * B * nf * 9 + B * nl * 6)
    return max(2, min(16, math.ceil(1.5 * L2_BYTES / bytes_per_set))), bytes_per_set


def config_dict(workload, world, large_scaling="strong"):
    B, nf, nl, kw, desc = WORKLOADS[workload]
    line_sharded = workload == "large" and world > 1
    nl_total = nl * world if (line_sharded and large_scaling == "weak") else nl
    n_sets, bytes_per_set = n_input_sets(workload)
    twist = workload in ("large", "fmr", "demo")
    return {"workload": desc, "name": workload, "pairs_per_gpu": B, "triplets_per_cloud": nf, "lines_per_pair": nl_total,
            "sharding": "lines" if line_sharded else "pairs (batch)", "window": [1, 1, 5, 5],
            "backward": {"large": "d loss / d twist (6) through the fused se(3) transform of cloud 1",
                         "demo": "d loss / d twist (6) through the fused se(3) transform of cloud 1",
                         "fmr": "d loss / d twist (B,6) through FMR's Exp (ExpMap gradient) and the fused rigid transform"}.get(
                             workload, "d loss / d points1 (B, nf, 9)"),
            "l2": "GPU arm: inputs rotate over %d distinct pre-resident sets (%.0f MB > 126 MB L2), no flush needed" %
                  (n_sets, n_sets * bytes_per_set / 2 ** 20)}


# ---------------------------------------------------------------------------------------------------------
# CPU arms: the unmodified reference from baseline/_ref (kind "reference") and the eager-torch port (kind "port")
# ---------------------------------------------------------------------------------------------------------
def _load_reference():
    try:
        from baseline import fetch_reference as fr
        return fr.load()
    except Exception:
        return None


def cpu_step_port(tri1, tri2, lines, chunk=512):
    import torch
    from oracle import torch_port as tp
    t1 = torch.from_numpy(tri1).clone().requires_grad_(True)
    loss = tp.loss_pair(t1, torch.from_numpy(tri2), torch.from_numpy(lines), chunk=chunk)
    if loss is not None:
        loss.backward()
    return 0.0 if loss is None else float(loss.item())


def cpu_step_reference(ref, tri1, tri2, lines):
    """the reference's own call (loss.py:170-232) + autograd backward, exactly as its callers make it"""
    import torch
    t1 = torch.from_numpy(tri1).clone().reshape(1, -1, 9).requires_grad_(True)
    out = ref.cal_loss_intersection_batch_whole_median_pts_lines(1, 1, 5, 5, t1, torch.from_numpy(tri2).reshape(1, -1, 9),
                                                                 torch.from_numpy(lines).reshape(1, -1, 6), "cpu")
    if isinstance(out, tuple):
        return 0.0
    out.backward()
    return float(out.item())


def reference_on_this_gpu(torch, host_set, dev, budget_s=8.0):
    """The UNMODIFIED reference (baseline/_ref: code/loss.py) run on THIS GPU through its own `device` argument, one pair per call
    as its DL hooks call it (Train_DCP.py:266-270): the like-for-like datapoint beside the CPU arm.  Outside every timed region;
    None when the reference is not there or a pair does not fit (it materialises (nl, nf, 3, 3) temporaries)."""
    ref = _load_reference()
    if ref is None:
        return None
    tri1, tri2, lines = (x[0] for x in host_set)
    nl = lines.shape[0]
    try:
        t2 = torch.from_numpy(tri2).reshape(1, -1, 9).to(dev)
        ln = torch.from_numpy(lines).reshape(1, -1, 6).to(dev)

        def one():
            t1 = torch.from_numpy(tri1).reshape(1, -1, 9).to(dev).requires_grad_(True)
            out = ref.cal_loss_intersection_batch_whole_median_pts_lines(1, 1, 5, 5, t1, t2, ln, dev)
            if not isinstance(out, tuple):
                out.backward()
            torch.cuda.synchronize(dev)
            return 0.0 if isinstance(out, tuple) else float(out.item())
        loss = one()                                        # warm-up (allocator, kernels)
        times = []
        t_end = time.perf_counter() + budget_s
        while len(times) < 5 and time.perf_counter() < t_end:
            t0 = time.perf_counter()
            one()
            times.append(time.perf_counter() - t0)
        best = min(times)
        return {"value": nl / best, "unit": "pairs*lines/s", "ms_per_pair": best * 1e3, "loss": loss, "calls": len(times),
                "note": "the unmodified reference code/loss.py on this GPU (its own device argument; eager PyTorch CUDA, autograd "
                        "backward), ONE pair per call as its hooks call it, best of %d after a warm-up; a batch of B pairs is B such "
                        "calls" % len(times)}
    except Exception as exc:                               # e.g. out of memory on a pair that does not fit
        return {"error": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
    finally:
        torch.cuda.empty_cache()


def cpu_sample_lines(workload):
    """lines of one pair one CPU step evaluates: the reference materialises 36 * nl * nf bytes per temporary (~10 live)"""
    _, nf, nl, _, _ = WORKLOADS[workload]
    return nl if nf * nl <= 1024 * 20000 else max(64, int(1024 * 20000 / nf))


def cpu_baseline(workload, budget_s=20.0, inputs=None):
    """pairs x lines / s of the reference's CPU path on a bounded sample of the workload (one pair, a prefix of its
    lines): the unmodified reference when baseline/_ref travelled and the pair fits in memory, else the port"""
    import torch
    B, nf, nl, kw, _ = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tri1, tri2, lines = inputs if inputs is not None else make_inputs(workload, 0, 1, n_base=1)[0]
    tri1, tri2, lines = tri1[0], tri2[0], lines[0]
    n_s = cpu_sample_lines(workload)
    ref = _load_reference() if nf <= 8192 else None          # beyond that the dense temporaries do not fit: port, chunked
    out = {}
    for kind, fn in (("reference", (lambda a, b, c: cpu_step_reference(ref, a, b, c)) if ref else None), ("port", cpu_step_port)):
        if fn is None:
            continue
        t0 = time.perf_counter()
        fn(tri1, tri2, lines[:n_s])
        warm = time.perf_counter() - t0
        reps = max(1, min(4, int(0.5 * budget_s / max(warm, 1e-3)) - 1))
        best = warm
        for _ in range(reps):
            t0 = time.perf_counter()
            fn(tri1, tri2, lines[:n_s])
            best = min(best, time.perf_counter() - t0)
        out[kind] = (n_s / best, best, reps + 1)
    kind = "reference" if "reference" in out else "port"
    res = {"value": out[kind][0], "unit": "pairs*lines/s", "cores": cores, "kind": kind,
           "sample": "1 pair of the workload (%d triplets per cloud), first %d of %d lines, forward+backward, %s, best of %d; "
                     "extrapolation in the line count is linear" %
                     (nf, n_s, nl, "the UNMODIFIED reference code/loss.py from baseline/_ref (torch CPU, autograd backward)"
                      if kind == "reference" else "eager PyTorch CPU restatement of code/loss.py (oracle/torch_port.py)", out[kind][2]),
           "seconds_per_step": out[kind][1]}
    if kind == "reference":
        res["port_value"] = out["port"][0]
        res["port_note"] = "oracle/torch_port.py (sparse backward: a faster-than-reference restatement) on the same sample"
    return res


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
    ref = _load_reference() if nf <= 8192 else None
    kind = "reference" if ref is not None else "port"
    step = (lambda n: cpu_step_reference(ref, tri1, tri2, lines[:n])) if ref is not None else (lambda n: cpu_step_port(tri1, tri2, lines[:n]))
    n_s = cpu_sample_lines(args.workload)
    t0 = time.perf_counter()
    step(n_s)
    first = time.perf_counter() - t0
    # keep the whole run (warmup + steps) within ~150 s by shrinking the per-step line sample if needed
    total_steps = args.steps + max(args.warmup - 1, 0)
    if first * total_steps > 150.0:
        n_s = max(64, int(n_s * 150.0 / (first * total_steps)))
    for _ in range(max(args.warmup - 1, 0)):
        step(n_s)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(n_s)
    dt = (time.perf_counter() - t0) / args.steps
    value = n_s / dt
    what = ("the UNMODIFIED reference (code/loss.py from baseline/_ref, torch CPU, autograd backward)" if kind == "reference"
            else "eager PyTorch CPU restatement of the reference (oracle/torch_port.py; baseline/_ref absent or the pair too large for "
                 "the reference's dense temporaries)")
    sample = ("each step = forward+backward of 1 pair of the workload (%d triplets per cloud) on the first %d of its %d lines; %s, "
              "%d threads" % (nf, n_s, nl, what, cores))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs*lines/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong" if (args.workload == "large" and args.gpus > 1 and args.large_scaling == "strong") else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args.workload, args.gpus, args.large_scaling),
            "cpu_baseline": {"value": value, "unit": "pairs*lines/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "pairs*lines/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if kind == "reference":                                  # the port beside it, as a second figure
        t0 = time.perf_counter()
        cpu_step_port(tri1, tri2, lines[:n_s])
        cpu_step_port(tri1, tri2, lines[:n_s])
        line["cpu_baseline"]["port_value"] = n_s / ((time.perf_counter() - t0) / 2)
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def pin_to_local_cpus(local_rank, world):
    """per-rank CPU affinity: each rank keeps to its own slice of the host cores (the ranks' pinned-memory copies and
    launch threads otherwise migrate across sockets)"""
    try:
        cpus = sorted(os.sched_getaffinity(0))
        if world > 1 and len(cpus) >= 2 * world:
            per = len(cpus) // world
            os.sched_setaffinity(0, set(cpus[local_rank * per:(local_rank + 1) * per]))
            return per
    except Exception:
        pass
    return None


class Timer:
    """device time of `steps` calls of step(i), max over ranks"""

    def __init__(self, torch, dist, world, dev):
        self.torch, self.dist, self.world, self.dev = torch, dist, world, dev

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, step, steps, warmup, after=None, repeats=1):
        """`repeats` timed regions of EXACTLY `steps` steps each (barrier + synchronize on both sides, CUDA events, max over
        ranks); returns the median region's ms per step, the last step's outputs and every region's ms per step"""
        torch = self.torch
        for i in range(warmup):
            step(i)
        if after:
            after()
        per = []
        last = None
        for _ in range(repeats):
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                last = step(i)
            if after:
                after()
            e1.record()
            self.barrier()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
            if self.world > 1:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            per.append(float(t.item()) / steps)
        self.last_repeats = per
        return float(np.median(per)), last


def capture_graphs(torch, rrl_b200, compute, n_sets, dev, barrier):
    """one CUDA graph per input set holding the device work of a step; returns [(graph, outputs, launches)] or None"""
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
            graphs.append((gph, out, rrl_b200.launch_count() - n0))
        barrier()
        return graphs
    except Exception as exc:                               # capture not available: keep the eager step, say so
        sys.stderr.write("bench: CUDA graph capture failed (%s: %s); running eagerly\n" % (type(exc).__name__, exc))
        torch.cuda.synchronize()
        return None


def executed_view(workload, kernel_ms, peak_tflops):
    """the executed-instruction side of the dense kernel from the committed ncu capture of the same inputs
    (profiles/r02_dense_<workload>.json, tools/ncu_summary.py): FP32 flops per launch counted from the per-opcode
    thread-instruction counts (FFMA2 x 4, FFMA x 2, FMUL2 x 2, FMUL, FADD), divided by the kernel time measured live"""
    for tag in ("r02", "r01"):
        prof = os.path.join(ROOT, "profiles", "%s_dense_%s.json" % (tag, workload))
        if os.path.exists(prof):
            try:
                d = json.load(open(prof))
            except Exception:
                continue
            out = {"source": "profiles/%s_dense_%s.json (ncu --set full, one launch; counts are per launch and input-determined, "
                             "the time is this run's)" % (tag, workload),
                   "issue_active_pct": d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   "fma_pipe_cycles_active_pct": d.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                   "warps_active_pct": d.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
                   "long_scoreboard_stalls_per_issue": d.get("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
                   "registers_per_thread": d.get("launch__registers_per_thread"),
                   "ncu_kernel_us": d.get("gpu__time_duration.sum"),
                   "traffic": d.get("dram_bytes_per_launch")}
            fl = d.get("fp32_flops_per_launch")
            if fl:
                out["fp32_flops_per_launch"] = fl
                out["achieved_tflops"] = fl / (kernel_ms * 1e-3) / 1e12
                out["frac"] = out["achieved_tflops"] / peak_tflops
            return out
    return None


def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    cpus_per_rank = pin_to_local_cpus(local_rank, world)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import rrl_b200
    L = rrl_b200._native.lib()
    T = Timer(torch, dist, world, dev)

    if args.workload == "large":
        blk = bench_large(args, torch, dist, rrl_b200, T, rank, local_rank, world, dev, headline=True)
        if world > 1:
            dist.barrier()
        return 0 if blk is not None else 1

    B, nf, nl, kw, desc = WORKLOADS[args.workload]
    n_sets, bytes_per_set = n_input_sets(args.workload)
    host_sets = make_inputs(args.workload, rank, n_sets)
    dev_sets = [tuple(torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in s) for s in host_sets]
    twist_mode = args.workload in ("demo", "fmr")
    twist0 = torch.tensor([0.01, -0.02, 0.015, 0.03, -0.01, 0.02], device=dev).repeat(B, 1)
    rows = [s[0].reshape(B, -1, 3) for s in dev_sets] if twist_mode else None
    # --reuse-order 1 (demo): the clouds' spatial order is kept from step to step (LossSession), as in a registration loop
    demo_session = rrl_b200.LossSession() if (args.workload == "demo" and args.reuse_order) else None

    def compute(i):
        """the device work of one step on input set i: forward + backward"""
        t1, t2, ln = dev_sets[i % n_sets]
        if args.workload == "fmr":                           # gradient to the twist (B,6) through Exp
            tw = twist0.clone().requires_grad_(True)
            loss = rrl_b200.hooks.fmr_twist_loss(tw, rows[i % n_sets], t2, ln)
            total = loss.sum()
            total.backward()
            return total.detach(), tw.grad
        if args.workload == "demo":                          # the se(3) optimiser's step (test_demo...py:57-66)
            tw = twist0.clone().requires_grad_(True)
            total = rrl_b200.twist_loss(tw, t1, t2, ln, session=demo_session).sum()
            total.backward()
            return total.detach(), tw.grad
        t1 = t1.detach().requires_grad_(True)
        loss = rrl_b200.intersected_line_loss(t1, t2, ln)
        total = loss.sum()
        total.backward()
        return total.detach(), t1.grad

    pending = []

    def exchange(out):
        if world > 1:
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

    def drain():
        while pending:
            pending.pop().wait()

    graphs = capture_graphs(torch, rrl_b200, compute, n_sets, dev, T.barrier) if args.graph else None
    if graphs:
        def step(i):
            gph, out, _ = graphs[i % n_sets]
            gph.replay()
            return exchange(out)
    else:
        def step(i):
            return exchange(compute(i))

    warm = max(args.warmup, 3)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = rrl_b200.launch_count()
    ms_per_step, last = T.run(step, args.steps, warm, after=drain, repeats=args.repeats)
    main_repeats = list(T.last_repeats)
    launches = rrl_b200.launch_count() - launches0
    if graphs:                                            # replayed, not re-issued by the host: count what each graph holds
        launches = sum(graphs[i % n_sets][2] for i in range(args.steps))
    else:
        launches = launches * args.steps // (args.steps * args.repeats + warm)
    clocks = sampler.stop() if sampler else None
    value = world * B * nl / (ms_per_step * 1e-3)

    e2e = bench_e2e(args, torch, dist, rrl_b200, T, host_sets, B, nf, nl, world, local_rank, dev, bytes_per_set, n_sets)
    e2e["cpus_per_rank"] = cpus_per_rank

    # ---- the unchanged hook loop (B = 1 slices in a Python loop) and the batched hook helper, DCP-style ----
    dropin = hooks = None
    if args.workload in ("dcp", "rpm"):
        try:
            dropin, hooks = bench_hook_paths(args, torch, rrl_b200, T, dev_sets, B, nf, nl, world, n_sets, dev, ms_per_step)
        except Exception as exc:
            dropin = {"error": "%s: %s" % (type(exc).__name__, exc)}

    large = None
    if args.large_block and args.workload == "dcp":
        try:
            large = bench_large(args, torch, dist, rrl_b200, T, rank, local_rank, world, dev, headline=False)
        except Exception as exc:
            large = {"error": "%s: %s" % (type(exc).__name__, exc)}
            torch.cuda.synchronize()

    if rank != 0:
        if world > 1:
            dist.barrier()
        return 0

    roofline = bench_roofline(args, torch, rrl_b200, dev_sets[0], nf, nl, ms_per_step)
    sampled = bench_sampler(torch, rrl_b200, dev_sets[0], nl, kw, world * B * nl, ms_per_step)
    cpu = cpu_baseline(args.workload, inputs=host_sets[0]) if (world == 1 and not args.no_cpu_baseline) else None
    ref_gpu = None
    if world == 1 and not args.no_cpu_baseline and args.workload in ("dcp", "rpm", "fmr", "demo"):
        ref_gpu = reference_on_this_gpu(torch, host_sets[0], dev)
    cfg = config_dict(args.workload, world)
    line = {"metric": METRIC, "value": value, "unit": "pairs*lines/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg,
            "run": {"launch": "one CUDA graph per input set" if graphs else "eager launches", "input_sets": n_sets,
                    "timed_regions_ms_per_step": main_repeats,
                    "statistic": "median of %d timed regions of exactly %d steps each (a single region of %d steps lasts a few "
                                 "milliseconds)" % (args.repeats, args.steps, args.steps),
                    "api": {"demo": "rrl_b200.twist_loss (exp + transform + loss forward, pose-space backward)",
                            "fmr": "rrl_b200.hooks.fmr_twist_loss (Exp -> fused rigid transform -> loss; ExpMap gradient)"}.get(
                                args.workload, "rrl_b200.intersected_line_loss (torch.autograd.Function over the C ABI), forward + backward")},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "lines_sampled_on_device": sampled, "dropin": dropin, "hooks": hooks, "large": large, "reference_on_this_gpu": ref_gpu,
            "loss_checksum": float(last[0].item())}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
    return 0


def bench_e2e(args, torch, dist, rrl_b200, T, host_sets, B, nf, nl, world, local_rank, dev, bytes_per_set, n_sets):
    """host buffers through the C ABI (H2D of the step's inputs from pinned memory, D2H of loss + status): the timed loop is
    repeated and the MEDIAN reported (one ~8 ms wall-clock shot is at the mercy of a single scheduler hiccup)"""
    L = rrl_b200._native.lib()
    ctx = C.c_void_p()
    rrl_b200._native.check(L.rrl_host_create(B, nf, nf, nl, local_rank, C.byref(ctx)), "rrl_host_create")
    pinned = [torch.from_numpy(np.ascontiguousarray(np.concatenate([x.reshape(-1) for x in s]))).pin_memory() for s in host_sets]
    n_sub = int(L.rrl_host_subbatches(ctx))
    n1, n2 = B * nf * 9, B * nf * 9
    h_loss = np.zeros(B, np.float32)
    h_status = np.zeros(B, np.int32)
    n_slots = int(L.rrl_host_slots(ctx))
    tickets = [C.c_int(-1) for _ in range(n_slots)]

    def submit(i):
        base = pinned[i % n_sets].data_ptr()
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

    def timed(fn, n):
        T.barrier()
        t0 = time.perf_counter()
        fn(n)
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n

    def blocking(n):
        for i in range(n):
            p = pinned[i % n_sets]
            rrl_b200._native.check(L.rrl_host_loss_fwd_bwd(ctx, p.data_ptr(), p.data_ptr() + 4 * n1, p.data_ptr() + 4 * (n1 + n2),
                                                           1, 1, 5, 5, h_loss.ctypes.data, h_status.ctypes.data, None),
                                   "rrl_host_loss_fwd_bwd")
    steps = max(args.steps, 20)
    run(5)
    reps = [timed(run, steps) for _ in range(args.e2e_repeats)]
    blk = [timed(blocking, steps) for _ in range(3)]
    L.rrl_host_destroy(ctx)
    med = float(np.median(reps))
    return {"value": world * B * nl / med, "unit": "pairs*lines/s", "h2d_bytes_per_step": bytes_per_set, "d2h_bytes_per_step": 8 * B,
            "api": "rrl_host_submit + rrl_host_wait (C ABI, host pointers; per step: H2D of the three inputs from pinned memory, "
                   "forward, backward to points1, D2H of loss+status; %d buffer sets in flight, so the copies of step i+1 run "
                   "under the kernels of step i, and each step is cut into %d sub-batches of pairs on their own streams)" %
                   (n_slots, n_sub),
            "ms_per_step": med * 1e3, "repeats_ms_per_step": [r * 1e3 for r in reps], "steps_per_repeat": steps,
            "statistic": "median of %d repeats of the timed loop, each the max over ranks" % len(reps),
            "blocking_call_ms_per_step": float(np.median(blk)) * 1e3}


def bench_hook_paths(args, torch, rrl_b200, T, dev_sets, B, nf, nl, world, n_sets, dev, batched_ms):
    """(1) `dropin`: the reference's hook loop replayed literally -- Train_DCP.py:252-297: torch transform of the source
    triplets with the predicted (R, t), transpose + reshape, a Python loop over the pairs calling the drop-in
    cal_loss_intersection_batch_whole_median_pts_lines on B = 1 slices, / 5.0, / batch_size, backward to (R, t) -- once with
    the shim's slice batching (default) and once without (every pair its own forward + backward).
    (2) `hooks`: rrl_b200.hooks.dcp_loss, the batched helper with the fused transform."""
    M = rrl_b200.loss
    gen = torch.Generator().manual_seed(5)
    Rp0 = torch.linalg.qr(torch.randn(B, 3, 3, generator=gen))[0]
    Rp0 = (0.02 * Rp0 + torch.eye(3)).to(dev)               # near-identity predicted transforms (not orthonormal: irrelevant here)
    tp0 = (0.01 * torch.randn(B, 3, generator=gen)).to(dev)
    src_cf = [s[0].reshape(B, -1, 3).transpose(2, 1).contiguous() for s in dev_sets]      # (B, 3, 3nf) as the DCP loader stores it
    tar_cf = [s[1].reshape(B, -1, 3).transpose(2, 1).contiguous() for s in dev_sets]

    def loop_step(i):
        src, tar, ln = src_cf[i % n_sets], tar_cf[i % n_sets], dev_sets[i % n_sets][2]
        Rp, tp = Rp0.clone().requires_grad_(True), tp0.clone().requires_grad_(True)
        tar_faces = tar.transpose(2, 1).reshape(B, -1, 9)
        pred = (torch.matmul(Rp, src) + tp.unsqueeze(2)).transpose(2, 1).reshape(B, -1, 9)           # utils.py:32-37
        acc = torch.zeros(1, device=dev)
        for j in range(B):
            acc = acc + M.cal_loss_intersection_batch_whole_median_pts_lines(1, 1, 5, 5, pred[j:j + 1, :, :], tar_faces[j:j + 1, :, :],
                                                                             ln[j:j + 1, :, :], dev) / 5.0
        out = acc / B
        out.backward()
        return out.detach(), Rp.grad

    def hook_step(i):
        Rp, tp = Rp0.clone().requires_grad_(True), tp0.clone().requires_grad_(True)
        out = rrl_b200.hooks.dcp_loss(src_cf[i % n_sets], Rp, tp, tar_cf[i % n_sets], dev_sets[i % n_sets][2])
        out.backward()
        return out.detach(), Rp.grad

    res = {}
    steps = max(10, min(args.steps, 50))
    M.BATCH_SLICES = True
    ms_b, _ = T.run(loop_step, steps, 3)
    out_b = loop_step(0)
    M.BATCH_SLICES = False
    try:
        ms_u, _ = T.run(loop_step, max(3, steps // 5), 1)
        out_u = loop_step(0)
    finally:
        M.BATCH_SLICES = True
    # what the caller's own Python costs: the same loop with the loss call replaced by a constant (1,) tensor -- slicing,
    # `/ 5.0`, `+=`, the transform and their autograd nodes are the hook's, not the library's
    const = torch.zeros(1, device=dev, requires_grad=True)
    real = M.cal_loss_intersection_batch_whole_median_pts_lines
    M.cal_loss_intersection_batch_whole_median_pts_lines = lambda *a, **k: const * 1.0
    try:
        ms_floor, _ = T.run(loop_step, steps, 3)
    finally:
        M.cal_loss_intersection_batch_whole_median_pts_lines = real
    dropin = {"api": "rrl_b200.loss.cal_loss_intersection_batch_whole_median_pts_lines called on B=1 slices in the reference's "
                     "own Python loop (Train_DCP.py:252-297 replayed literally, eager launches, backward to the predicted R, t)",
              "ms_per_step": ms_b, "value": world * B * nl / (ms_b * 1e-3), "unit": "pairs*lines/s",
              "vs_batched_step": ms_b / batched_ms,
              "callers_own_python_ms_per_step": ms_floor,
              "callers_own_python_note": "the same hook loop with the loss call stubbed by a constant tensor: the 32 x (3 slices, "
                                         "/ 5.0, +=) eager torch ops and autograd nodes of Train_DCP.py:266-270 themselves",
              "library_share_ms_per_step": ms_b - ms_floor,
              "without_slice_batching_ms_per_step": ms_u, "without_slice_batching_value": world * B * nl / (ms_u * 1e-3),
              "loss_equal_bits": bool(out_b[0].item() == out_u[0].item())}
    ms_h, _ = T.run(hook_step, steps, 3)
    hooks = {"api": "rrl_b200.hooks.dcp_loss (one native forward + backward per batch, fused rigid transform, dL/dR + dL/dt)",
             "ms_per_step": ms_h, "value": world * B * nl / (ms_h * 1e-3), "unit": "pairs*lines/s", "vs_batched_step": ms_h / batched_ms}
    return dropin, hooks


def bench_roofline(args, torch, rrl_b200, dev_set, nf, nl, ms_per_step):
    """the dominant kernel (dense intersection), timed alone with CUDA events on its stream"""
    L = rrl_b200._native.lib()
    peak, peak_ms = C.c_double(), C.c_double()
    L.rrl_measure_fp32_peak(0, C.byref(peak), C.byref(peak_ms))
    t1, t2, ln = dev_set
    Bm = t1.shape[0]
    wsb = L.rrl_workspace_bytes(Bm, nf, nf, nl)
    ws = torch.empty(wsb, dtype=torch.uint8, device=t1.device)
    md, mp = C.c_float(), C.c_float()
    torch.cuda.synchronize()
    rrl_b200._native.check(L.rrl_measure_dense(t1.data_ptr(), t2.data_ptr(), ln.data_ptr(), Bm, nf, nf, nl, ws.data_ptr(), wsb, 10,
                                               C.byref(md), C.byref(mp), None), "rrl_measure_dense")
    alg_flops = 48.0 * Bm * nl * (nf + nf)                   # 16 flops per (line, point) test
    alg = alg_flops / (md.value * 1e-3) / 1e12
    # the like-for-like FP32 datapoint: the reference's own formulation, every (line, triplet) tested with the exact
    # non-contracted arithmetic, one thread per line (rrl_debug_set_param(5, 1)); bounded to a sub-batch so it stays short
    brute = None
    try:
        nb = max(1, min(Bm, int(2 ** 33 // max(1, 2 * nf * nl * 48))))
        bd, bp = C.c_float(), C.c_float()
        L.rrl_debug_set_param(5, 1)
        try:
            rrl_b200._native.check(L.rrl_measure_dense(t1.data_ptr(), t2.data_ptr(), ln.data_ptr(), nb, nf, nf, nl, ws.data_ptr(),
                                                       L.rrl_workspace_bytes(nb, nf, nf, nl), 2, C.byref(bd), C.byref(bp), None),
                                   "rrl_measure_dense(bruteforce)")
        finally:
            L.rrl_debug_set_param(5, 0)
        bfl = 48.0 * nb * nl * 2 * nf
        brute = {"kernel": "rrl::bruteforce_kernel", "pairs": nb, "kernel_ms": bd.value, "tflops": bfl / (bd.value * 1e-3) / 1e12,
                 "frac": bfl / (bd.value * 1e-3) / 1e12 / peak.value,
                 "note": "executes the algorithmic 48 non-contracted flops of every (line, triplet) pair (point 0 always, points 1-2 "
                         "when point 0 passes): what 'no culling' costs on this GPU",
                 "filtered_pipeline_speedup": (bd.value / nb) / (md.value / Bm)}
    except Exception as exc:
        brute = {"error": "%s: %s" % (type(exc).__name__, exc)}
    ex = executed_view(args.workload, md.value, peak.value)
    hbm_peak = None
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    alg_bytes = 4.0 * Bm * (9 * 2 * nf + 6 * nl)
    hbm_achieved = alg_bytes / (md.value * 1e-3) / 1e9
    frac = ex["frac"] if ex and ex.get("frac") else None
    return {"bound": "fp32", "kernel": "rrl::dense_kernel + rrl::exact_kernel (the dense stage of one forward)",
            "achieved": ex["achieved_tflops"] if frac else None, "peak": peak.value, "unit": "TFLOP/s", "frac": frac,
            "frac_definition": "EXECUTED FP32 flops of the dense stage (ncu per-opcode thread-instruction counts of this launch: "
                               "FFMA2 x4, FFMA x2, FMUL2 x2, FMUL, FADD) / kernel time measured here / FP32 FMA peak measured here",
            "traffic": ex.get("traffic") if ex else None, "executed": ex,
            "peak_source": "measured live: register-resident FFMA loop (rrl_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 entry",
            "kernel_ms": md.value, "prep_ms": mp.value, "kernel_share_of_step": md.value / ms_per_step,
            "algorithmic": {"flops_per_launch": alg_flops, "tflops": alg, "speedup_vs_bruteforce_at_peak": alg / peak.value,
                            "note": "48 flops per (line, triplet) of the reference formulation (SURVEY 8(d)) / kernel time: the kernel "
                                    "decides most pairs with a conservative bounding-sphere + FMA predicate and runs the exact "
                                    "reference-order test only on candidates, so this exceeds the peak -- it is the strength "
                                    "reduction of the culling, not a utilisation"},
            "bruteforce": brute,
            "hbm": {"achieved": hbm_achieved, "peak": hbm_peak or 6552.3, "unit": "GB/s", "frac": hbm_achieved / (hbm_peak or 6552.3),
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if hbm_peak else "fallback: the pool's measured 6552.3 GB/s"}}


def bench_sampler(torch, rrl_b200, dev_set, nl, kw, pairs_lines_per_step, ms_per_step):
    """lines sampled on the device instead of supplied (SURVEY 8(d) asks for both; the sampler stays out of the roofline)"""
    try:
        t1, t2, _ = dev_set
        v1 = t1.reshape(t1.shape[0], -1, 9)[:, :, :3].contiguous()           # the clouds = point 0 of every triplet
        v2 = t2.reshape(t2.shape[0], -1, 9)[:, :, :3].contiguous()
        lo2, hi2 = v2.min(1)[0], v2.max(1)[0]
        rad = (hi2 - lo2).norm(dim=1, keepdim=True) * float(kw.get("radius_scale", 0.5))
        cen = v2.mean(1)
        for _ in range(3):
            lines_s, filled = rrl_b200.sample_lines(rad, cen, nl, v1, v2, seed=11, offset=0)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s0.record()
        for it in range(20):
            lines_s, filled = rrl_b200.sample_lines(rad, cen, nl, v1, v2, seed=11, offset=it + 1)
        s1.record()
        torch.cuda.synchronize()
        smp_ms = s0.elapsed_time(s1) / 20
        return {"sampler_ms_per_batch": smp_ms, "filled_fraction": float(filled.float().mean().item()) / nl,
                "value_with_sampling": pairs_lines_per_step / ((ms_per_step + smp_ms) * 1e-3), "unit": "pairs*lines/s",
                "note": "rrl_sample_lines (Philox4x32-10, 10 rounds, 12-triangle AABB test of loss.py:265-432) timed alone on "
                        "this rank's batch; value_with_sampling = the step with the sampler in front of it"}
    except Exception as exc:
        return {"error": "%s: %s" % (type(exc).__name__, exc)}


def bench_large(args, torch, dist, rrl_b200, T, rank, local_rank, world, dev, headline):
    """BASELINE configs[4]: one scan pair, 500k triplets per cloud x 100k lines, evaluated through the se(3) twist
    (forward, backward, 6-float gradient).  N = 1: the single-GPU step.  N > 1: the lines are sharded over the ranks --
    strong (the 100k lines split) and weak (100k lines per rank of a pair with N x 100k lines) -- with the exchange step
    done inside the kernels over peer memory (rrl_shard_tail; NCCL protocol as the fallback).  `<mode>_reuse` = the steps of a
    registration loop (LossSession(static_target=True)): the clouds' spatial order is kept from step to step and the fixed
    target's thresholds / records / bounding spheres are not rebuilt (RRL_REUSE_ORDER | RRL_REUSE_TARGET); the source is
    re-transformed, re-thresholded and its spheres rebuilt every step, the lines change every step."""
    L = rrl_b200._native.lib()
    B, nf, nl, kw, desc = WORKLOADS["large"]
    steps = max(10, min(args.steps, 40))
    modes = ["strong", "weak"] if world > 1 else ["single"]
    if headline and world > 1:
        modes = [args.large_scaling]
    twist0 = torch.tensor([0.01, -0.02, 0.015, 0.03, -0.01, 0.02], device=dev)
    nl_max = nl * world if "weak" in modes else nl
    pair = make_inputs("large", 0, 1, nl=nl_max)[0]                             # the ONE pair all ranks share
    t1 = torch.from_numpy(pair[0][0]).to(dev)
    t2 = torch.from_numpy(pair[1][0]).to(dev)
    all_lines = pair[2][0]
    rng = np.random.default_rng(3)
    out = {"workload": desc, "triplets_per_cloud": nf, "steps": steps,
           "backward": "d loss / d twist (6) through the fused se(3) transform of cloud 1"}
    comm = None
    if world > 1:
        comm = rrl_b200.dist.PeerComm.create(nl)                              # collective
    out["exchange"] = ("peer memory inside the kernels (rrl_shard_tail + rrl_comm_allreduce_f64 over CUDA IPC / NVLink)"
                       if comm is not None else ("NCCL protocol (3 all-reduces + twist-gradient all-reduce)" if world > 1 else "none"))
    results = {}
    for mode in modes:
        nl_total = nl * world if mode == "weak" else nl
        lo, hi = rrl_b200.dist.shard_range(nl_total, rank, world)
        # distinct line sets per step (the lines are resampled every step of a registration loop): rotate over 3 permutations
        n_sets = 3
        sets = []
        for s in range(n_sets):
            perm = rng.permutation(nl_total) if s else np.arange(nl_total)
            sets.append(torch.from_numpy(np.ascontiguousarray(all_lines[:nl_total][perm][lo:hi])).to(dev))
        for reuse in (False, True):
            sess = rrl_b200.LossSession(static_target=True) if reuse else None

            def compute(i, sess=sess):
                ln = sets[i % n_sets]
                tw = twist0.clone().requires_grad_(True)
                if world > 1:
                    loss, _, _ = rrl_b200.dist.line_sharded_twist_loss(tw, t1, t2, ln, session=sess, comm=comm)
                else:
                    loss = rrl_b200.twist_loss(tw.reshape(1, 6), t1[None], t2[None], ln[None], session=sess)
                total = loss.sum()
                total.backward()
                return total.detach(), tw.grad
            graphs = capture_graphs(torch, rrl_b200, compute, n_sets, dev, T.barrier) if args.graph else None
            if graphs:
                def step(i, graphs=graphs):
                    graphs[i % n_sets][0].replay()
                    return graphs[i % n_sets][1]
            else:
                step = compute
            ms, last = T.run(step, steps, 5)
            key = mode + ("_reuse" if reuse else "")
            results[key] = {"ms_per_step": ms, "value": nl_total / (ms * 1e-3), "unit": "pairs*lines/s", "lines_total": nl_total,
                            "lines_per_rank": hi - lo, "launch": "one CUDA graph per line set" if graphs else "eager launches",
                            "launches_per_step": graphs[0][2] if graphs else None, "loss": float(last[0].item())}
            del graphs
        # per-step split on this rank (eager, CUDA events): what every rank repeats (prep: thresholds, order, node records),
        # the dense stage of its line shard, and the rest (records, exchange + tail, backward, se(3))
        ln = sets[0]
        wsb = L.rrl_workspace_bytes(1, nf, nf, ln.shape[0])
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        md, mp = C.c_float(), C.c_float()
        torch.cuda.synchronize()
        rrl_b200._native.check(L.rrl_measure_dense(t1.data_ptr(), t2.data_ptr(), ln.data_ptr(), 1, nf, nf, ln.shape[0], ws.data_ptr(),
                                                   wsb, 5, C.byref(md), C.byref(mp), None), "rrl_measure_dense")
        base = results[mode]
        base["split_ms"] = {"replicated_prep": mp.value, "dense": md.value, "records_exchange_tail_backward": base["ms_per_step"] - mp.value - md.value}
        del ws
    if comm is not None:
        out["comm_error"] = comm.error()
        comm.close()
    out["results"] = results
    if world > 1:
        # efficiency against the single-GPU step of the same build is computed by whoever holds both lines; the N = 1 time is
        # measured by the N = 1 run of this bench (block `large.results.single`)
        out["note"] = "strong: lines_total fixed at 100k; weak: 100k lines per rank; value = lines_total / step time (max over ranks)"
    if not headline:
        return out
    if rank == 0:
        mode = modes[0]
        r = results[mode + "_reuse"] if args.reuse_order else results[mode]
        cfg = config_dict("large", world, args.large_scaling)
        line = {"metric": METRIC, "value": r["value"], "unit": "pairs*lines/s", "n_gpus": world, "steps": steps, "warmup": 5,
                "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "strong" if (world > 1 and mode == "strong") else "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": cfg, "run": {"launch": r["launch"], "exchange": out["exchange"]},
                "e2e": None, "gpu_launches": (r["launches_per_step"] or 0) * steps, "large": out, "loss_checksum": r["loss"]}
        if world == 1:
            try:
                line["roofline"] = bench_roofline(args, torch, rrl_b200, (t1[None], t2[None], sets[0][None]), nf, nl, r["ms_per_step"])
                if not args.no_cpu_baseline:
                    line["cpu_baseline"] = cpu_baseline("large", inputs=(pair[0], pair[1], pair[2][:, :nl]))
            except Exception as exc:
                line["roofline"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        print(json.dumps(line), flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dcp", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", type=int, default=1, help="1: replay the device work of a step as a CUDA graph (one per input set); 0: eager launches")
    ap.add_argument("--large-scaling", default="strong", choices=["strong", "weak"],
                    help="workload large at N > 1: strong = the 100k lines of the pair are split over the ranks (BASELINE "
                         "configs[4]); weak = every rank keeps 100k lines of a pair with N x 100k lines")
    ap.add_argument("--reuse-order", type=int, default=0, help="workloads large / demo: the step with the clouds' order kept from step to step (LossSession)")
    ap.add_argument("--large-block", type=int, default=1, help="default workload: also measure BASELINE configs[4] (block `large` of the line)")
    ap.add_argument("--e2e-repeats", type=int, default=5)
    ap.add_argument("--repeats", type=int, default=5, help="timed regions of --steps steps each; the median is reported")
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
