"""CPU: the bench.py JSON contract, checked on the committed line of the last GPU run (profiles/r01_bench_v*.json)
and on the reference arm's line, so a change to bench.py that drops a key is caught without a GPU."""
import glob
import json
import os
import re

from conftest import ROOT


def _latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)),
                   key=lambda p: (int(re.search(r"r(\d+)_", os.path.basename(p)).group(1)), int(re.search(r"_v(\d+)", p).group(1))))
    assert files, pattern
    return json.loads(open(files[-1]).read().strip().splitlines()[-1])


def test_bench_line_has_every_contract_key():
    d = _latest("r0?_bench_v*.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "frames_per_sec_512x512_to_binary_code" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 64 * 512 * 512 * 3 and e["d2h_bytes_per_step"] > 0
    assert abs(e["value"] - d["value"]) > 1e-6                       # measured separately, not a copy of value
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
    assert d["parity"]["flips_outside_band"] == 0


def test_reference_arm_line():
    d = _latest("r0?_bench_ref_v*.json")
    ours = _latest("r0?_bench_v*.json")
    assert d["impl"] == "reference" and d["metric"] == ours["metric"] and d["unit"] == ours["unit"]
    assert d["config"]["workload"] == ours["config"]["workload"] and d["higher_is_better"] == ours["higher_is_better"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference")
