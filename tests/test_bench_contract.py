"""The measurement contract on the committed records (CPU only: nothing here runs a kernel).

The bench line of the round (`profiles/r02al_bench_256cubed_1gpu.json`) carries every key the driver reads, names
BASELINE config 3, and its kernel shares agree with the ncu launch list of the same step
(`profiles/r02at_launches_step_256cubed_final.csv`, summed by `tools/launch_summary.py`): ncu times are serialised
and cold-cache, so it is each kernel's SHARE of the step that must agree, not the absolute."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINE = os.path.join(ROOT, "profiles", "r02al_bench_256cubed_1gpu.json")
LAUNCHES = os.path.join(ROOT, "profiles", "r02at_launches_step_256cubed_final.csv")


def test_bench_line_has_the_contract_keys():
    j = json.load(open(LINE))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in j, key
    assert j["metric"] == "gravity_interactions_per_s" and j["higher_is_better"] is True and j["warmup"] >= 3
    assert "16777216" in j["config"]["workload"] and "model" not in j["config"]
    assert abs(j["value"] - (j["config"]["pc_pairs"] + j["config"]["pp_pairs"]) / (j["ms_per_step"] * 1e-3)) <= 1e-6 * j["value"]
    r = j["roofline"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    e = j["e2e"]
    assert e["h2d_bytes_per_step"] == 40 * 16777216 and e["d2h_bytes_per_step"] == 24 * 16777216
    assert e["value"] < j["value"]  # the copies are inside the timed region
    c = j["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["sample"]
    assert j["gpu_launches"] > 0 and not set(j["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert j["parity"]["within_tolerance"] and j["parity"]["median_da_over_a"] <= 1e-4


def test_launch_list_shares_agree_with_the_bench_line(tmp_path):
    out = tmp_path / "summary.json"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), LAUNCHES, str(out)],
                          stdout=subprocess.DEVNULL)
    s = json.load(open(out))
    j = json.load(open(LINE))
    assert abs(s["step_sum_ms"] - j["ms_per_step"]) <= 0.03 * j["ms_per_step"]  # nothing overlaps, nothing idles
    share = {k["kernel"].split("<")[0]: k["share"] for k in s["kernels"]}
    step = j["ms_per_step"]
    assert abs(share["cell_list_x2_kernel"] - j["kernels"]["pc_ms"] / step) <= 0.02
    assert abs(share["part_list_stream_kernel"] - j["kernels"]["pp_ms"] / step) <= 0.02
    assert abs(share["ewald_slot_kernel"] - j["kernels"]["ewald_ms"] / step) <= 0.02


def test_bench_command_line_is_the_contract():
    h = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, check=True).stdout
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in h, flag


def test_reference_arm_runs_the_reference_own_cpu_gravity():
    """`bench.py --impl reference` on a small box: one JSON line with the arm's keys; kind "reference" (the
    reference's own gravity.h / Ewald.cpp compiled unmodified) where oracle/_ref/libgravity_ref.so exists, and the
    oracle port with --ref-port -- which agree to rounding"""
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libgravity_ref.so"))
    lines = {}
    for flag in ([], ["--ref-port"]):
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "262144", "--steps", "1",
                            "--warmup", "1", "--ref-pairs", "2e7"] + flag, capture_output=True, text=True, check=True, cwd=ROOT)
        out = [l for l in p.stdout.splitlines() if l.startswith("{")]
        assert len(out) == 1
        j = json.loads(out[0])
        assert j["impl"] == "reference" and j["metric"] == "gravity_interactions_per_s" and j["value"] > 0
        assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert j["cpu_baseline"]["value"] == j["value"] and j["cpu_baseline"]["cores"] >= 1
        lines[bool(flag)] = j
    assert lines[True]["cpu_baseline"]["kind"] == "port"
    assert lines[False]["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    if have_ref:
        assert lines[False]["config"]["port_vs_reference_max_da_over_a"] < 1e-9
