"""bench.py contract (CPU side): the reference arm runs here (it times the oracle port on the host cores) and must print
exactly one JSON line with the contract keys; the committed evidence of the GPU arm carries every key the driver reads."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["value"] > 0


def test_committed_gpu_bench_line_has_every_contract_key():
    d = json.loads((ROOT / "profiles" / "r2_bench_final_default.json").read_text().splitlines()[0])
    assert BASE_KEYS | {"gpu_launches", "clocks", "roofline", "cpu_baseline", "eager_b200", "c4", "infer", "gemm"} <= set(d)
    assert d["eager_b200"]["ms_per_step"] > d["ms_per_step"] and d["c4"]["value"] > 0
    assert d["infer"]["models"] == 7 and d["infer"]["bf16"]["value"] > d["infer"]["fp32"]["value"] > 0
    assert all("frac_of_bf16_peak" in v for k, v in d["gemm"].items() if not k.startswith("_"))
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["dtype"] == "bf16"
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] is not None
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"] * 1.05
    assert "workload" in d["config"] and "model" not in d["config"]
