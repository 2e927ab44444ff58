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


def test_round2_evidence_is_consistent():
    """k2_traffic.json is the sum the committed ncu raw page of the final K2 holds; the round-2 bench lines carry the contract keys,
    the K2 roofline object and the sharded runs' per-rank times."""
    import csv
    want = json.load(open(os.path.join(ROOT, "profiles", "k2_traffic.json")))["1329x542000"]
    rows = list(csv.reader(open(os.path.join(ROOT, want["source"].split(" ")[0]), newline="")))      # the capture the figure quotes
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units, vals = rows[hi], rows[hi + 1], rows[hi + 2]
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total = sum(float(d[k]) * scale[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    assert "fused_linearize_kernel" in d["Kernel Name"]
    assert abs(total - want["dram_bytes_per_launch"]) <= 1e-6 * total and total < 2 * 133405246        # <= 2x the algorithmic bytes
    line = json.loads([l for l in open(os.path.join(ROOT, "profiles", "r02_bench_final.json")) if l.startswith("{")][-1])
    for k in ("metric", "value", "unit", "n_gpus", "ms_per_step", "scaling", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in line, k
    k2 = line["roofline"]["k2"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(k2) and k2["unit"] == "GB/s"
    assert line["verify"]["bit_exact"] is True and line["ba_large"]["lm"]["termination"] == 0
    for n in (2, 4, 8):
        ln = json.loads([l for l in open(os.path.join(ROOT, "profiles", f"r02_bench_n{n}_sharded_configs2.json")) if l.startswith("{")][-1])
        assert ln["n_gpus"] == n and len(ln["config"]["per_rank_ms_per_step"]) == n and "1329 images" in ln["config"]["workload"]
        assert ln["ba_large"]["allreduce_ms"] > 0 and ln["ba_large"]["allreduce_message_bytes"] < 13e6


def test_launch_shares_tool_on_the_committed_launch_lists():
    """tools/launch_shares.py on the committed ncu launch lists: K1 dominates the matching pass, the band Cholesky and the fused
    linearisation kernel the BA solve (the shares DESIGN.md / profiles/README.md quote)."""
    import subprocess
    import sys
    def shares(name):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_shares.py"), os.path.join(ROOT, "profiles", name)],
                             capture_output=True, text=True, check=True).stdout
        rows = {}
        for line in out.splitlines():
            if " share " in line:
                rows[line.split()[0]] = float(line.split("share")[1].replace("%", ""))
        return rows
    m = shares("r02_launches_match.csv")
    assert 75.0 < m["k1::match_pair_kernel<0>"] < 90.0 and m["resolve_rows_kernel"] < 15.0
    b = shares("r02_launches_ba_band.csv")
    assert b["band::band_cholesky_kernel"] + b["ba::fused_linearize_kernel<0>"] > 80.0
