"""CPU: the reference arm of bench.py (`--impl reference`: the oracle port timed on host cores) runs without a GPU and prints
one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "videomae_vit_s_pretrain_clips_per_sec" and line["unit"] == "clips/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_recorded_gpu_bench_lines_carry_the_contract_keys():
    """The GPU arm cannot run here; the lines it printed on the B200 boxes are kept under profiles/ and must carry every key of the
    bench contract (base keys, clocks, e2e with byte counts, gpu_launches, roofline, cpu_baseline at N = 1)."""
    for name, n in (("r2_bench_1gpu.json", 1), ("r2_bench_2gpu.json", 2), ("r2_bench_8gpu.json", 8)):
        with open(os.path.join(ROOT, "profiles", name)) as f:
            line = json.loads([ln for ln in f.read().splitlines() if ln.startswith("{")][-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                  "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert k in line, (name, k)
        assert line["metric"] == "videomae_vit_s_pretrain_clips_per_sec" and line["unit"] == "clips/s" and line["n_gpus"] == n
        assert line["scaling"] == "weak" and line["dtype"] == "bf16" and line["data"] == "synthetic" and line["warmup"] >= 3
        assert line["gpu_launches"] > 0 and "workload" in line["config"] and "model" not in line["config"]
        assert abs(line["value"] - n * line["config"]["per_gpu_batch"] / (line["ms_per_step"] / 1e3)) < 1e-6 * line["value"]
        e = line["e2e"]
        assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] <= line["value"] * 1.02
        r = line["roofline"]
        assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(line["clocks"])
        assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if n == 1:
            cb = line["cpu_baseline"]
            assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb
