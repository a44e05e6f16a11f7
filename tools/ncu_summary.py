"""Turns ncu captures (gpurun_out/*.ncu-rep, launch-list CSVs) into the small tracked summaries under profiles/.

    python tools/ncu_summary.py rep  gpurun_out/prof_dense_dcp.ncu-rep  profiles/r01_dense_dcp.json
    python tools/ncu_summary.py list gpurun_out/launches_bench.csv     profiles/r01_launches_bench.csv
"""
import csv
import json
import subprocess
import sys
from collections import OrderedDict

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__cycles_active.avg", "smsp__cycles_active.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def opcode_counts(path, metric):
    """per-opcode instance values of a per-opcode metric for every kernel of the report: [{opcode: count}]"""
    import re
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv", "--print-metric-instances", "details", "--metrics", metric],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3 or metric not in rows[0]:
        return []
    col = rows[0].index(metric)
    out = []
    for vals in rows[2:]:
        d = {}
        for name, cnt in re.findall(r"([A-Z0-9_.]+): (\d+)", vals[col]):
            d[name] = int(cnt)
        out.append(d)
    return out


def fp32_flops(thread_ops):
    """executed FP32 flops from per-opcode THREAD-instruction counts: packed FFMA2 = 2 FMAs = 4 flops per thread-instruction"""
    w = {"FFMA2": 4, "FFMA": 2, "FMUL2": 2, "FADD2": 2, "FMUL": 1, "FADD": 1}
    return float(sum(thread_ops.get(k, 0) * v for k, v in w.items()))


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    kernels = []
    for vals in rows[2:]:
        d = OrderedDict()
        d["kernel"] = vals[hdr.index("Kernel Name")]
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[k] = float(vals[i].replace(",", ""))
                except ValueError:
                    d[k] = vals[i]
                d[k + " [unit]"] = units[i]
        def to_bytes(key):
            v, u = d.get(key), d.get(key + " [unit]", "")
            if v is None:
                return None
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        r, w = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
        if r is not None and w is not None:
            d["dram_bytes_per_launch"] = r + w
        kernels.append(d)
    thr = opcode_counts(path, "sass__thread_inst_executed_true_per_opcode")
    wrp = opcode_counts(path, "sass__inst_executed_per_opcode")
    for i, d in enumerate(kernels):
        if i < len(thr) and thr[i]:
            d["fp32_flops_per_launch"] = fp32_flops(thr[i])
            d["fp32_thread_instructions"] = {k: thr[i][k] for k in ("FFMA2", "FFMA", "FMUL2", "FMUL", "FADD") if k in thr[i]}
            d["thread_instructions_total"] = sum(thr[i].values())
        if i < len(wrp) and wrp[i]:
            top = sorted(wrp[i].items(), key=lambda kv: -kv[1])[:16]
            d["warp_instructions_top"] = OrderedDict(top)
            d["warp_instructions_total"] = sum(wrp[i].values())
    res = kernels[0] if len(kernels) == 1 else {"kernels": kernels, "dram_bytes_per_launch": kernels[0].get("dram_bytes_per_launch")}
    res["source"] = path
    json.dump(res, open(out, "w"), indent=1)
    print(out, res.get("dram_bytes_per_launch"))


def launches(path, out):
    """ncu --csv launch list -> per-kernel totals and share of the capture"""
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
        name = r[ik].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("kernel,launches,total_us,share\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.1f,%.4f\n" % (k, n, us, us / tot))
    print(open(out).read())


if __name__ == "__main__":
    {"rep": rep, "list": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
