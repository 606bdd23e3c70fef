"""bench.py's reference arm (the CPU restatement on the host cores) prints the driver's JSON contract: one line, the keys of
the GPU arm's line for metric / unit / config, `impl: "reference"`, an `e2e` object with zero copy bytes and a
`cpu_baseline` describing the run.  CPU only (the GPU arm is exercised on the GPU box by the driver)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--windows", "4", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "local_ba_lm_iters_per_sec" and j["unit"] == "LM iters/s"
    assert j["higher_is_better"] is True and j["vs_baseline"] is None and j["dtype"] == "f64" and j["data"] == "synthetic"
    assert j["steps"] == 1 and j["warmup"] == 0 and j["n_gpus"] == 1 and j["value"] > 0 and j["ms_per_step"] > 0
    assert set(j["config"]) == {"workload"} and "4 independent windows" in j["config"]["workload"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == j["value"] and cb["cores"] >= 1 and "window solves" in cb["sample"]
