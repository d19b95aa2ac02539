"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the reference's CPU implementation of the
path timed on the host cores) runs here as it does on the GPU box, prints ONE JSON line with the keys the driver reads, and
describes the same workload as our arm; ranks other than 0 print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_the_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["unit"] == "Msamples/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0
    # value is the line's own throughput: samples of the whole C2 capture per step
    assert abs(d["value"] - d["config"]["samples_total"] / (d["ms_per_step"] * 1e-3) / 1e6) < 1e-6 * d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    # the same workload string as our arm (bench.py formats one template for both)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"]["workload"] == bench.WORKLOAD.format(width=bench.SAMPLES_PER_GPU // bench.N_FFT)
    assert d["metric"] == bench.METRIC


def test_reference_arm_is_silent_on_other_ranks():
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}) == []
