"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "7pt-24", "--steps", "2", "--warmup", "1",
                        "--ref-ranks", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "CSRMatrix apply! GFLOP/s" and d["unit"] == "GFLOP/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["nnz"] == 7 * 24 ** 3 - 6 * 24 ** 2
    assert d["steps"] == 2 and d["warmup"] == 1  # the arm honours --steps / --warmup
    # the config object is the product arm's, key for key
    sys.path.insert(0, ROOT)
    import bench
    spec = bench.workload_spec(type("A", (), {"workload": "7pt-24", "rows": 0})())
    assert d["config"] == bench.bench_config(spec, d["config"]["nnz"], 1, 1)


def test_reference_arm_non_root_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "7pt-8"], capture_output=True,
                       text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under juliapetra.jl_b200/ may import or load it"""
    pkg = os.path.join(ROOT, "juliapetra.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "libjp_oracle" not in src, f
