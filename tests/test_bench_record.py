"""The committed bench lines (profiles/r1_bench_n1.json, profiles/r2_bench_n1.json, written by
bench.py on a B200) carry every key of the measurement contract, and their derived numbers are
consistent."""
import json
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_committed_bench_line_has_the_contract_keys():
    d = json.loads((ROOT / "profiles" / "r1_bench_n1.json").read_text())
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "pairs/s" and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["unit"] == "pairs/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # achieved = algorithmic bytes per launch / average launch duration
    want = r["pairs_per_launch"] * r["algorithmic_bytes_per_pair"] / (r["launch_ms"] * 1e-3) / 1e9
    assert abs(r["achieved"] - want) / want < 1e-6
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == "pairs/s" and c["sample"]
    # value = pairs of the timed region / its duration
    pairs = d["config"]["pairs_per_step"]
    assert abs(d["value"] - pairs / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_round2_bench_line():
    d = json.loads((ROOT / "profiles" / "r2_bench_n1.json").read_text())
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "parity", "full_matrix"):
        assert k in d, k
    assert d["unit"] == "pairs/s" and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["vs_baseline"] is None
    # the benched workload equals the reference cell for cell
    assert d["parity"]["mismatches"] == 0 and d["parity"]["cells"] > 1000 and d["parity"]["against"] == "reference"
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    want = r["pairs_per_launch"] * r["algorithmic_bytes_per_pair"] / (r["launch_ms"] * 1e-3) / 1e9
    assert abs(r["achieved"] - want) / want < 1e-6
    pairs = d["config"]["pairs_per_step"]
    assert abs(d["value"] - pairs / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    f = d["full_matrix"]
    assert f["rows"] == 3085 and abs(f["pairs_per_s"] - 3085 * 3084 / f["seconds"]) / f["pairs_per_s"] < 1e-6
    assert d["cub_calls"] == 0 and d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # the traffic file belongs to kernel sources by digest
    t = json.loads((ROOT / "profiles" / "walk_traffic.json").read_text())
    assert len(t["kernel_sha"]) == 12 and t["dram_bytes_read"] > 0 and t["pairs_per_launch"] == 3084


def test_round2_lines_agree_across_gpu_counts_and_with_the_traffic_capture():
    """The whole 3085 x 3085 matrix has the same digest at every N (the N > 1 == N = 1 record of the bench), and the
    headline line quotes the DRAM traffic of the committed capture."""
    lines = {n: json.loads((ROOT / "profiles" / f"r2_bench_n{n}.json").read_text()) for n in (1, 2, 4, 8)}
    digests = {d["full_matrix"]["blake2b_of_matrix"] for d in lines.values()}
    assert len(digests) == 1, digests
    for n, d in lines.items():
        assert d["n_gpus"] == n and sum(d["full_matrix"]["rows_per_rank"]) == 3085 and d["full_matrix"]["rows"] == 3085
        if n > 1:  # weak scaling: every rank does the single-GPU step
            assert 0.9 < d["value"] / (n * lines[1]["value"]) < 1.1
    t = json.loads((ROOT / "profiles" / "walk_traffic.json").read_text())
    assert lines[1]["roofline"]["traffic"] == t["dram_bytes_read"] + t["dram_bytes_write"]
    # the lines of the pools that are not like the benchmark are there and are slower than it, not faster
    for name in ("c4_repeats", "512_repeats", "512_outbreak"):
        d = json.loads((ROOT / "profiles" / f"r2_bench_{name}.json").read_text())
        assert d["cub_calls"] == 0 and 0 < d["value"] < lines[1]["value"]
