"""The bench lines committed under profiles/ (one per BASELINE config, final code of round 2) carry every key of the bench
contract: metric / value / unit / n_gpus / steps / warmup / ms_per_step / higher_is_better / scaling / vs_baseline / dtype /
data / config.workload, the end-to-end object with its copy sizes, the launch count, the clocks sampled during the timed
region, and the roofline object; the headline line also the CPU baseline."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(p for p in glob.glob(os.path.join(ROOT, "profiles", "r02_bench_*.json"))
               if "refgpu" not in p and "_n2" not in p and "_n4" not in p)


def _line(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def test_every_baseline_config_has_a_line():
    names = {os.path.basename(p)[len("r02_bench_"):-len(".json")] for p in LINES}
    assert {"bair_b64", "bair_b6", "bair_b1", "landscape_b32_fast", "landscape_b32", "dtdb_fire_seq24_b32",
            "iper128_transfer_b64"} <= names


@pytest.mark.parametrize("path", LINES, ids=[os.path.basename(p) for p in LINES])
def test_bench_line_follows_the_contract(path):
    d = _line(path)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "ab"):
        assert k in d, k
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["n_gpus"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    # value and ms_per_step describe the same timed region (frames per step / seconds per step)
    frames = d["value"] * d["ms_per_step"] / 1e3
    assert abs(frames - round(frames)) < 1e-3 * frames and round(frames) % 16 == 0
    e = d["e2e"]
    assert e["unit"] == "frames/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0.8 * d["value"] < e["value"] < 1.05 * d["value"]           # through host buffers: never faster than it can be
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and "traffic" in r
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # the median leg is the one reported
    legs = sorted(d["ab"]["device_ms_per_step"])
    assert abs(legs[(len(legs) - 1) // 2] - d["ms_per_step"]) < 1e-6


def test_headline_line_has_the_cpu_baseline():
    d = _line(os.path.join(ROOT, "profiles", "r02_bench_bair_b64.json"))
    b = d["cpu_baseline"]
    assert b["kind"] == "port" and b["unit"] == "frames/s" and b["cores"] >= 1 and b["value"] > 0 and b["sample"]
    assert d["value"] / b["value"] > 100
