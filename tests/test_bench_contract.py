"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`: the CPU port of the reference's path on the
host cores, one JSON line, same metric / unit / config keys as our arm) and its behaviour on the non-zero ranks of a torchrun
launch (exit 0 without work or output)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--records", "20000", "--cpu-sample", "20000", "--nodes", "20000", "--haps", "4", "--steps", "2", "--warmup", "1"]


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1"] + SMALL, capture_output=True, text=True,
                       cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.split("\n") if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("GAF records/s") and d["unit"] == "records/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["dtype"] == "int64" and d["data"] == "synthetic" and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="3", WORLD_SIZE="8", LOCAL_RANK="3")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "8"] + SMALL, capture_output=True, text=True, cwd=ROOT,
                       env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
