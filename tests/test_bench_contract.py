"""bench.py's reference arm runs on CPU: check the JSON contract of the line it prints (the GPU arm's line has
the same keys plus roofline / clocks and is produced on the B200 box)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-sample", "300"], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["metric"] == "agent-steps/sec" and line["value"] > 0 and line["higher_is_better"] is True
    assert "workload" in line["config"]
    cb = line["cpu_baseline"]
    # oracle/_ref (the unmodified reference learner, staged by oracle/make_ref.py) is there in the build container and on the
    # GPU box; a bare checkout falls back to the numpy port and must say so
    from oracle import make_ref
    assert cb["kind"] == ("reference" if make_ref.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == line["value"] and "N=300" in cb["sample"]
    assert "reference sample N=300" in line["config"]["workload"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_bench_line_has_the_contract_keys():
    """The bench line kept under profiles/ (the run the docs quote) carries every key of the measurement contract, and its
    `roofline.traffic` is the number of the committed ncu capture -- which bench.py reports only while the CUDA sources of
    the step still hash to what the capture was taken from."""
    import bench
    line = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_1gpu.json")))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["metric"] == bench.METRIC and line["unit"] == bench.UNIT and line["n_gpus"] == 1
    assert "workload" in line["config"] and line["gpu_launches"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert 0 < line["e2e"]["value"] < line["value"]
    roof = line["roofline"]
    assert roof["bound"] == "hbm" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    now = bench.ncu_traffic(roof["kernel"])
    assert now is None or now == roof["traffic"]
