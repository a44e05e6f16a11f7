"""The bench.py contract that can be checked without a GPU: the reference arm (the unmodified reference from
baseline/_ref when it is there -- it is fetched by __graft_entry__.build() in the build container --, else the
eager-torch port) prints ONE JSON line with the keys the driver reads, with the SAME config dict as the GPU arm, and the
workload table carries the sizes of BASELINE.json's configs."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_workload_table_matches_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    w = bench.WORKLOADS
    assert (w["dcp"][0], w["dcp"][1], w["dcp"][2]) == (32, 1024, 15000)        # Train_DCP.py:252-255: 15000 lines per pair
    assert (w["rpm"][0], w["rpm"][1], w["rpm"][2]) == (64, 2048, 10000)        # Train_RPM.py:218-222
    assert (w["fmr"][0], w["fmr"][1], w["fmr"][2]) == (128, 1024, 15000)       # fmr/model.py:285-288
    assert (w["demo"][0], w["demo"][1], w["demo"][2]) == (1, 1024, 20000)      # test_demo_optimized_Lie_Algebra.py:30-31
    assert (w["large"][0], w["large"][1], w["large"][2]) == (1, 500000, 100000)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "pairs" in base["metric"] and "lines" in base["metric"]


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS=str(min(os.cpu_count() or 1, 8)))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "pairs*lines/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["steps"] == 1 and d["warmup"] == 1 and d["ms_per_step"] > 0 and d["n_gpus"] == 1
    assert "workload" in d["config"] and d["config"]["name"] == "dcp"
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.config_dict("dcp", 1)                        # what the GPU arm prints: same_config
    cb = d["cpu_baseline"]
    have_ref = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "loss.py"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "lines" in cb["sample"]
    if have_ref:
        assert cb["port_value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
