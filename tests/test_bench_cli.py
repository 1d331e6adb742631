"""bench.py contract checks that need no GPU: the reference arm (CPU oracle port) prints one JSON line with the keys the
driver reads, and non-zero ranks of a multi-rank reference launch exit quietly."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cpu-n", "32", "--steps", "1",
                           "--warmup", "1"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.strip().split("\n") if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "particles/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("IC particles/sec") and d["value"] > 0 and d["dtype"] == "f64"
    import oracle
    assert d["cpu_baseline"]["kind"] == ("reference" if oracle.ref_binary() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_without_work():
    r = _run({"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""
