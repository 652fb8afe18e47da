"""bench.py contract on the CPU side: the reference arm (CPU oracle) prints ONE JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-sample", "2", "--voltages", "4"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "solves/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("GMPNP steady solves/sec") and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    split = d["cpu_baseline"]["per_newton_iteration_ms"]                 # assembly / LU split
    assert split["assembly"] > 0 and split["sparse_lu"] > 0


def test_both_arms_describe_the_same_config():
    """`config` is built by one function for both arms, so the driver's same_config check compares equal dicts."""
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    a = argparse.Namespace(voltages=256, dv=0.75, xtol_path=1.0, xtol=1e-10)
    cfg = mod.base_config(a)
    assert cfg["workload"].startswith("config2") and cfg["points_total_config2"] == 7680
    pts = mod.stratified_points(2, 256)
    assert len(pts) == 60 and len({(p.L_n, p.conc, p.cation) for p in pts}) == 30


def test_cpu_workers_do_not_import_torch():
    """The CPU arm's workers import NumPy/SciPy only (no multi-second `import torch` inside the measurement)."""
    code = ("import sys; sys.path.insert(0, %r); import bench; bench._worker_init(); "
            "r = bench._cpu_solve_point(('K', 0.1, 1e-6, -0.5, 0.75, 1.0, True, 1e-10)); "
            "assert r['ok'] and r['u'].shape == (1091, 7) and 'torch' not in sys.modules; print('ok')") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr[-2000:]


def test_reference_arm_is_silent_on_non_zero_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
