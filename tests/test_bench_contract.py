"""bench.py's output contract on the leg that runs without a GPU (`--impl reference`, the CPU restatement on the host cores):
stdout is exactly ONE line of JSON with the keys the driver reads; everything else goes to stderr."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line(product_lib, oracle_lib):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--particles", "400"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle_tetra_crossings_per_second" and d["unit"] == "crossings/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["dtype"] == "f64" and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
