"""bench.py's output contract, checked without a GPU: the reference arm (`--impl reference`: the oracle port on the host cores) is run
for one bounded step, and the committed round-end line of the B200 arm (profiles/) is checked for the keys the driver reads."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _line(text):
    lines = [l for l in text.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = _line(out.stdout)
    assert d["impl"] == "reference" and d["metric"].startswith("images/sec @256x256, 50-step DDIM, bs=8") and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0
    # "reference": the unmodified reference modules (/root/reference here, its vendored copy baseline/_ref on the GPU box); "port": the oracle
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["extrapolated"] is True and "workload" in d["config"]
    import bench
    assert d["config"]["workload"] == bench.workload_string("c2", 1.0), "both arms name the workload with the same string"
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # under torchrun only rank 0 works and prints
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_b200_line_has_the_contract_keys():
    d = _line(open(os.path.join(ROOT, "profiles", "r01_bench_final_s4.json.log")).read())
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["vs_baseline"] is None and d["scaling"] == "weak" and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] == 8 * 256 * 256 * 3 and 0 < d["e2e"]["value"] <= d["value"] * 1.01
    assert d["gpu_launches"] > 0 and d["steps"] >= 1 and d["warmup"] >= 3
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["unit"] in ("GB/s", "TFLOP/s") and "traffic" in r
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert abs(d["value"] - 8 * d["n_gpus"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
