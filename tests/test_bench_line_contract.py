"""CPU: the committed bench lines (what `python bench.py` printed on a B200, copied under profiles/) carry every key the
measurement contract names and are self-consistent -- the numbers DESIGN.md / README quote are recomputable from the line."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline")


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["r02s_bench.json", "r02q_bench.json", "r02s_bench_inception.json", "r02g_bench_8gpu.json"])
def test_bench_line_has_the_contract_keys_and_adds_up(name):
    d = _line(name)
    missing = [k for k in REQUIRED if k not in d]
    assert not missing, missing
    assert d["metric"] == "i3d_clips_per_sec" and d["unit"] == "clips/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["data"] == "synthetic" and d["dtype"] == "bf16" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["gpu_launches"] > 0
    cfg = d["config"]
    assert "workload" in cfg and "model" not in cfg
    # value = clips of all ranks per step / device time of a step
    clips_per_step = cfg["clips_per_video"] * cfg["crops"] * cfg["videos_per_step_per_gpu"] * d["n_gpus"]
    assert d["value"] == pytest.approx(clips_per_step / (d["ms_per_step"] / 1e3), rel=1e-6)
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0.5 * d["value"] < e["value"] <= 1.02 * d["value"]           # end to end is never meaningfully above resident
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9) and 0.0 < r["frac"] < 1.0
    c = d["clocks"]
    assert c["sm_mhz"] <= c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_single_gpu_line_carries_traffic_and_cpu_baseline():
    d = _line("r02s_bench.json")
    r = d["roofline"]
    assert r["traffic"] and r["traffic"] > 1e11            # measured DRAM bytes of one step (ncu), ~168 GB
    # whole-step roofline: conv FLOPs of the step over the step's device time
    assert r["achieved"] == pytest.approx(r["flops_per_step"] / (r["avg_step_ms"] / 1e3) / 1e12, rel=1e-6)
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    assert cb["parity"]["max_norm_err"] <= 1e-2 and cb["parity"]["cosine_min"] >= 0.999


def test_reference_arm_line():
    d = _line("r02a_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == "i3d_clips_per_sec" and d["unit"] == "clips/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
