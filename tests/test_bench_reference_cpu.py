"""CPU tests of bench.py's reference arm (`--impl reference`): the JSON contract, and that under torchrun -- which
exports OMP_NUM_THREADS=1 -- rank 0 alone runs it on ALL host cores while the other ranks exit without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches")


def _json_lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def _check(line, n_gpus):
    assert all(k in line for k in KEYS), [k for k in KEYS if k not in line]
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == n_gpus and line["gpu_launches"] == 0 and line["vs_baseline"] is None
    assert line["value"] > 0 and abs(line["e2e"]["value"] - line["value"]) < 1e-6 * line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["unit"] == "rays/s" and "rows spread evenly" in cb["sample"] and "NOT Embree" in cb["sample"]
    assert cb["cores"] == (os.cpu_count() or 1)
    assert cb["plain_walker_value"] > 0            # the parity checker's own walker, reported beside the faster timing traversal
    assert "cfg2" in line["config"]["workload"] and "1201x1201" in line["config"]["workload"]


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "0", "--ref-seconds", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = _json_lines(out.stdout)
    assert len(lines) == 1
    _check(lines[0], 1)


def test_reference_arm_under_torchrun_uses_all_cores_on_rank_0_only():
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29653", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0", "--ref-seconds", "2"], capture_output=True, text=True, timeout=900,
                         cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = _json_lines(out.stdout)
    assert len(lines) == 1, "exactly one rank prints the reference line"
    _check(lines[0], 2)
