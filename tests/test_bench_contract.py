"""The benchmark contract, checked without a GPU: bench.py's flags and defaults, and that the JSON lines committed under
profiles/ (produced by tools/run_gpu_checks.sh / run_multi_gpu.sh on B200 boxes) carry every key the contract asks for,
with internally consistent numbers."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

P = os.path.join(ROOT, "profiles")


def _line(name):
    s = open(os.path.join(P, name)).read().strip().splitlines()[-1]
    return json.loads(s)


def test_bench_flags_and_defaults():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--workload"):
        assert flag in out.stdout
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'add_argument("--gpus", type=int, default=1)' in src          # no flags: one GPU, a K / W that finish in minutes
    assert 'add_argument("--warmup", type=int, default=3)' in src        # W >= 3


@pytest.mark.parametrize("name,n", [("r02_bench_two_stage.json", 1), ("r02_bench_two_stage_2gpu.json", 2), ("r02_bench_two_stage_8gpu.json", 8)])
def test_committed_headline_lines_carry_the_contract(name, n):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == n and d["scaling"] == "weak" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["data"] == "synthetic" and d["dtype"] == "f32"
    cfg = d["config"]
    assert cfg["workload"] == "two_stage_morlet" and cfg["traces_per_gpu"] == 8192 and cfg["samples"] == 16384
    # value = units all ranks processed / time
    assert abs(d["value"] - 8192 * 16384 * n / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.5 < r["frac"] <= 1.0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 8192 * 16384 * 4 * n and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    c = d["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n == 1:
        cb = d["cpu_baseline"]
        assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb
        assert d["parity"]["vs_reference"]["tsPWS"] <= 1e-4
        x = d["extra_workloads"]
        assert "error" not in x and {"two_stage_mexhat", "single_stage_w6", "resample_c4", "pairs_c5"} <= set(x)
    else:
        assert d["parity"]["all_ranks_bit_identical"] is True and d["parity"]["peer_vs_nccl_allreduce_tsPWS"] <= 1e-6
        assert d["strong"]["headline_frame"]["parity_vs_reference"]["tsPWS"] <= 1e-4


def test_committed_reference_arm_line():
    d = _line("r02_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == _line("r02_bench_two_stage.json")["metric"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
