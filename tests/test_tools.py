"""CPU: the evidence tooling parses the committed profiles (so that the numbers DESIGN.md quotes can be re-derived)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ncu_summary_reproduces_the_committed_traffic_figure(tmp_path):
    raw = os.path.join(ROOT, "profiles", "r01b_k1_pair_ncu_raw.csv")
    out = tmp_path / "t.json"
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), raw, "--traffic-json", str(out)],
                         capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    assert "match_pair_kernel" in res.stdout and "sm__pipe_tensor_subpipe_imma_cycles_active" in res.stdout
    got = json.load(open(out))
    want = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
    assert got["dram_bytes_per_launch"] == want["dram_bytes_per_launch"] > 1e9


def test_final_bench_line_carries_the_contract_keys():
    line = [l for l in open(os.path.join(ROOT, "profiles", "r01b_bench_final.json")) if l.startswith("{")][-1]
    d = json.loads(line)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"])
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"])
