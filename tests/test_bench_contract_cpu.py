"""CPU: the reference arm of bench.py (`--impl reference`: the reference's own CPU build, oracle/_ref/ref_build_driver, or the
oracle port where the reference could not be compiled) prints ONE JSON line with the keys the driver reads.  Run at a
tiny scale; the full-size arm is what the driver times."""
import json
import os
import subprocess
import sys

from conftest import REPO


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--scale", "0.004", "--steps", "1", "--warmup", "1"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["unit"] == "k-mers/s" and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["metric"].startswith("canonical k-mers/sec")
    assert d["config"]["workload"].startswith("C2") and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_b200_arm_refuses_to_run_without_a_gpu():
    """no CPU fallback: without a CUDA device the product arm exits non-zero and says why (on the GPU box this test is moot)"""
    import dbg_assembly_b200 as dbg
    if dbg.capi.device_count() > 0:
        return
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1", "--warmup", "0", "--scale", "0.004"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=REPO)
    assert out.returncode != 0
    assert "no CUDA device" in out.stderr and not out.stdout.strip()
