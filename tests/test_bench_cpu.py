"""CPU: bench.py's contract where it can be checked without a GPU -- the reference arm prints ONE JSON line with the keys
the driver reads, and the product arm refuses to run without CUDA (no CPU fallback)."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--resolution", "64", "--batch", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "novel_views_per_sec_fwd_bwd_256x256" and d["unit"] == "views/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_product_arm_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run("--steps", "1", "--warmup", "0", timeout=300)
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)


def test_reference_arm_under_torchrun_only_rank0_prints():
    """N > 1 launch of the reference arm: rank 0 runs and prints the line, the other ranks exit 0 without work."""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", str(ROOT / "bench.py"), "--gpus", "2", "--impl", "reference",
                        "--steps", "1", "--warmup", "0", "--resolution", "64", "--batch", "1"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0


def test_config_presets_are_labelled_by_their_baseline_config():
    """--config N must label the line with configs[N] (round 1 shipped lines whose label named another config)."""
    sys.path.insert(0, str(ROOT))
    import bench
    base = dict(bench.CFG, workload="full")
    assert "BASELINE configs[1]" in bench.workload_name(dict(base, baseline_config=1))
    two = bench.workload_name(dict(base, baseline_config=2, B=2, V_t=3, f=1.2))
    assert "configs[2]" in two and "configs[1]" not in two and "focal 1.2" in two and "V_t=3" in two
    assert "configs[3]" in bench.workload_name(dict(base, baseline_config=3, B=8))
    assert "no BASELINE config" in bench.workload_name(dict(base, baseline_config=None, B=3))
    old = bench.FWD_ONLY
    try:
        bench.FWD_ONLY = True
        assert bench.workload_name(dict(bench.CFG, workload="splat")).startswith("splat_fwd_only")
    finally:
        bench.FWD_ONLY = old
