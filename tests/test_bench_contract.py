"""CPU tier: the parts of bench.py's contract that need no GPU — the reference arm (the oracle's CPU port on the host
cores) prints exactly one JSON line with the keys the driver reads, pins its BLAS threads even when the launcher
exported OMP_NUM_THREADS=1 (torchrun does), labels its extrapolation, and only rank 0 prints under torchrun."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, args=()):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           "--cpu-sample-rows", "20000", "--batch", "8", *args], capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run({"OMP_NUM_THREADS": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "extrapolated", "rows_timed"):
        assert key in d, key
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["extrapolated"] is True and d["rows_timed"] == 20000
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    try:
        want = len(os.sched_getaffinity(0))
    except Exception:
        want = os.cpu_count() or 1
    assert d["cpu_baseline"]["cores"] == want, "the BLAS pool must not stay at the launcher's OMP_NUM_THREADS=1"


def test_reference_arm_is_silent_on_other_ranks_and_knows_the_target_config():
    out = _run({"RANK": "3", "WORLD_SIZE": "8"}, ("--gpus", "8"))
    assert out.returncode == 0 and out.stdout.strip() == ""
    out = _run({"RANK": "0", "WORLD_SIZE": "8"}, ("--gpus", "8", "--config", "c4t"))
    d = json.loads(out.stdout.strip())
    assert d["scaling"] == "weak" and d["config"]["rows"] == 8 * 12_500_000 and d["config"]["batch"] == 1
