"""Per-source-line view of an ncu capture: joins `ncu --page source --csv` (SASS rows, in program order) with
`nvdisasm --print-line-info` of the locally built object (same compiler, same source => same instruction sequence).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep a-robust-registration-loss_b200/build/rrl_dense.o dense_kernelILi8ELb1 [top]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, obj, pattern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# NCU_ARGS='-k regex:tail_kernel' selects one kernel of a multi-kernel report
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + os.environ.get("NCU_ARGS", "").split(), capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
body = [r for r in rows[2:] if len(r) == len(hdr)]
col = {n: i for i, n in enumerate(hdr)}

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines, cur, inside = [], None, False
for ln in sass.splitlines():
    if ln.startswith(".text."):
        inside = pattern in ln
        continue
    if not inside:
        continue
    if ln.startswith("\t.section") or ln.startswith("//-----"):
        if lines:
            break
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m:
        lines.append((cur, m.group(2).strip()))
if len(lines) != len(body):
    print("warning: %d SASS rows in the report vs %d in the local object" % (len(body), len(lines)))
n = min(len(lines), len(body))
agg = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
tot_inst = tot_samp = 0
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
for i in range(n):
    key = lines[i][0]
    r = body[i]
    inst = int(r[col["Instructions Executed"]] or 0)
    samp = int(r[col["# Samples"]] or 0)
    a = agg[key]
    a[0] += inst
    a[1] += samp
    a[2] += 1
    for c in stall_cols:
        v = int(r[col[c]] or 0)
        if v:
            a[3][c[6:]] += v
    tot_inst += inst
    tot_samp += samp
print("total warp instructions %d, samples %d, SASS rows %d" % (tot_inst, tot_samp, n))
src_cache = {}


def src(key):
    if key is None:
        return ""
    f, l = key
    if f not in src_cache:
        for root in ("a-robust-registration-loss_b200/csrc", "."):
            p = os.path.join(root, f)
            if os.path.exists(p):
                src_cache[f] = open(p).read().splitlines()
                break
        else:
            src_cache[f] = []
    t = src_cache[f]
    return t[l - 1].strip()[:90] if 0 < l <= len(t) else ""


print("%-22s %7s %7s %5s  %-34s %s" % ("line", "inst%", "samp%", "sass", "top stalls", "source"))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = sorted(a[3].items(), key=lambda kv: -kv[1])[:3]
    print("%-22s %6.2f%% %6.2f%% %5d  %-34s %s" % ("%s:%d" % key if key else "?", 100.0 * a[0] / max(tot_inst, 1),
                                                 100.0 * a[1] / max(tot_samp, 1), a[2],
                                                 " ".join("%s:%d" % (k, v) for k, v in st), src(key)))

# REGIONS="name:lo-hi,name:lo-hi,..." (line ranges of the kernel's own file): executed instructions and samples per
# region; rows whose line info points into an intrinsic header inherit the last own-file line seen before them
if os.environ.get("REGIONS"):
    own = os.environ.get("REGION_FILE", "rrl_dense.cu")
    regs = []
    for part in os.environ["REGIONS"].split(","):
        name, rng = part.split(":")
        lo, hi = rng.split("-")
        regs.append((name, int(lo), int(hi)))
    acc = defaultdict(lambda: [0, 0, 0])
    last = None
    for i in range(n):
        key = lines[i][0]
        if key and key[0] == own:
            last = key[1]
        name = "?"
        for nm, lo, hi in regs:
            if last is not None and lo <= last <= hi:
                name = nm
                break
        r = body[i]
        acc[name][0] += int(r[col["Instructions Executed"]] or 0)
        acc[name][1] += int(r[col["# Samples"]] or 0)
        acc[name][2] += 1
    print("\nregion            inst%   samp%  sass")
    for nm, a in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        print("%-16s %6.2f%% %6.2f%% %5d" % (nm, 100.0 * a[0] / max(tot_inst, 1), 100.0 * a[1] / max(tot_samp, 1), a[2]))
