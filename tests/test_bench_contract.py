"""bench.py contract on CPU: the reference arm (oracle port on the host cores) prints ONE JSON line with the keys
the driver reads.  MPCF_BENCH_QUICK shrinks the calibration sample so this stays a few seconds."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    env = dict(os.environ, MPCF_BENCH_QUICK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=300, check=True).stdout.strip().splitlines()
    line = json.loads(out[-1])
    assert line["impl"] == "reference" and line["metric"] == "rollout_steps_per_s_with_jacobians"
    assert line["unit"] == "rollout-steps/s" and line["higher_is_better"] is True and line["dtype"] == "f64"
    assert line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_flop_model_is_the_frozen_one():
    sys.path.insert(0, ROOT)
    import bench
    fm = bench.flop_model(6)
    assert (fm["aba"], fm["step_values"], fm["step_jac"], fm["bytes_values"], fm["bytes_jac"]) == (2067, 8580, 223080, 344, 3944)
    assert bench.flop_model(12)["step_values"] == 19188 and bench.flop_model(37)["step_values"] == 63388
