"""bench.py's output contract, checked on the CPU arm (no GPU needed): `--impl reference` prints exactly
ONE line on stdout, a JSON object with the driver's keys, and under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra=None, gpus=1):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", str(gpus),
                        "--steps", "1", "--warmup", "0", "--cpu-sample", "32"],
                       capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = run()
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MLUPS" and d["higher_is_better"] is True
    assert d["metric"].startswith("MLUPS") and d["dtype"] == "f64" and d["data"] == "synthetic"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0


def test_reference_arm_under_torchrun_env_only_rank0_prints_and_uses_all_threads():
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must still use the host's cores
    out0 = run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0", "OMP_NUM_THREADS": "1"}, gpus=2)
    out1 = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1", "OMP_NUM_THREADS": "1"}, gpus=2)
    d = json.loads(out0.strip())
    assert d["n_gpus"] == 2 and d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert out1.strip() == ""
